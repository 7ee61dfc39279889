# raster parity tests + timings at the bench size and at the tile sizes of configs 3 / 4
python -m pytest tests/test_gpu_raster.py tests/test_gpu_graph.py tests/test_gpu_goals.py tests/test_gpu_npc.py -x -q 2>&1 | tail -15
python profiles/time_raster.py
TDS_RASTER_STRIPS=0 python profiles/time_raster.py
python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 256 64 128
