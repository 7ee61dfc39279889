python -m pytest tests/test_gpu_raster.py tests/test_gpu_graph.py tests/test_gpu_observations.py -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --kernels-only 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step', d['ms_per_step'], 'raster ms', d['roofline']['raster_ms_per_launch'])
    else: print(l[:300])
"
