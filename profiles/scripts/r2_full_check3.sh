python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python profiles/time_raster.py
python profiles/bench_configs.py 3 4
