python -m pytest tests/test_gpu_raster.py -x -q 2>&1 | tail -2
python profiles/time_raster.py
