python -m pytest tests/test_gpu_raster.py tests/test_gpu_offroad.py -x -q 2>&1 | tail -2
for oc in 2 3 4; do TDS_RASTER_CELL=16 TDS_OFFROAD_CELL=$oc python bench.py --steps 10 --warmup 3 --kernels-only 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('offroad cell $oc ms/step', d['ms_per_step'], 'raster ms', d['roofline']['raster_ms_per_launch'])
"; done
