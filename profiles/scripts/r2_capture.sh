# one --set full capture of the lean 64x64 raster kernel with source counters; the per-line table is made on the box
rm -rf gpurun_out/*
python profiles/time_raster.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -o gpurun_out/raster_cap python profiles/time_raster.py > gpurun_out/p_cap.log 2>&1
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/raster_cap_source.csv 2> /dev/null
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page source --csv --print-source cuda > gpurun_out/raster_cap_cuda.csv 2> /dev/null
ls -la gpurun_out; du -sh gpurun_out
