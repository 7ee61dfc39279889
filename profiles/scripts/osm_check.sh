# parity on the map built from its OSM file + config 3 with the three maps
python -m pytest tests/test_gpu_raster.py tests/test_gpu_offroad.py -x -q 2>&1 | tail -3
python profiles/bench_configs.py 3
