python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python profiles/time_raster.py
python profiles/time_raster_res.py 128 256 128
TDS_RASTER_STRIPS=0 python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 256 64 128
TDS_RASTER_STRIPS=1 python profiles/time_raster_res.py 256 64 128
python profiles/bench_configs.py 3 4 5
