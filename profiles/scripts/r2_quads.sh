# sliver quads: parity tests, then the raster timed with and without them (and at the other tile sizes: must not move)
rm -rf gpurun_out/*
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_configs.py -x -q 2>&1 | tail -5
python profiles/time_raster.py
TDS_RASTER_QUADS=0 python profiles/time_raster.py
TDS_RASTER_LEAN=0 python profiles/time_raster.py
python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 256 64 128
if [ "$1" = "cap" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -o gpurun_out/raster_cap python profiles/time_raster.py > gpurun_out/p_cap.log 2>&1
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/raster_cap_source.csv 2> /dev/null
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page raw --csv > gpurun_out/raster_cap_raw.csv 2> /dev/null
fi
du -sh gpurun_out
