# raster at other tile sizes with the default choice of threads per camera (TDS_RASTER_G forces one)
python -m pytest tests/test_gpu_raster.py tests/test_gpu_graph.py tests/test_gpu_goals.py -x -q 2>&1 | tail -3
python profiles/time_raster_res.py 96 256 128
python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 192 128 64
python profiles/time_raster_res.py 256 64 128
python profiles/time_raster_res.py 320 64 64
python profiles/time_raster_res.py 448 32 64
python profiles/time_raster.py
