# hot-path tests + device-resident bench loop only
python -m pytest tests/test_gpu_offroad.py tests/test_gpu_graph.py -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step', d['ms_per_step'], 'raster ms', d['roofline']['raster_ms_per_launch'], 'value', d['value'])
"
