rm -rf gpurun_out/*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_finish -s 4 -c 1 -o gpurun_out/finish_cap python profiles/time_raster.py > gpurun_out/p_cap.log 2>&1
timeout 300 ncu -i gpurun_out/finish_cap.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/finish_cap_source.csv 2> /dev/null
timeout 300 ncu -i gpurun_out/finish_cap.ncu-rep --page raw --csv > gpurun_out/finish_cap_raw.csv 2> /dev/null
rm -f gpurun_out/finish_cap.ncu-rep
du -sh gpurun_out
