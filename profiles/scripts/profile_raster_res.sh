# ncu --set full captures of the raster kernel at the tile sizes of configs 3 and 4 (128x128, 256x256)
python profiles/time_raster_res.py 128 256 128
python profiles/time_raster_res.py 256 64 128
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_r1_res128 python profiles/time_raster_res.py 128 256 128 > gpurun_out/p128.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_r1_res256 python profiles/time_raster_res.py 256 64 128 > gpurun_out/p256.log 2>&1
