# one iteration of the raster work: parity tests, timing, one ncu capture of the 64x64 kernel
python -m pytest tests/test_gpu_raster.py -x -q 2>&1 | tail -4
python profiles/time_raster.py
python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 256 64 128
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_${1:-r2x} python profiles/time_raster.py > gpurun_out/p_${1:-r2x}.log 2>&1
