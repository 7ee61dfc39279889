rm -rf gpurun_out/*
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_configs.py tests/test_gpu_graph.py -x -q 2>&1 | tail -3
python profiles/time_raster.py
TDS_RASTER_TWO_PASS=0 python profiles/time_raster.py
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 180 ncu --metrics $M --clock-control none -k regex:raster -s 12 -c 3 python profiles/time_raster.py 2>&1 | grep -E "raster_|no_instruction|long_score|issue_active|duration|inst_executed|dram__"
if [ "$1" = "cap" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 8 -c 1 -o gpurun_out/raster_cap python profiles/time_raster.py > gpurun_out/p_cap.log 2>&1
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/raster_cap_source.csv 2> /dev/null
timeout 300 ncu -i gpurun_out/raster_cap.ncu-rep --page raw --csv > gpurun_out/raster_cap_raw.csv 2> /dev/null
fi
du -sh gpurun_out
