rm -rf gpurun_out/*
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
python profiles/time_raster.py
for c in 8 12 24; do echo "cell $c"; TDS_RASTER_CELL=$c python profiles/time_raster.py; done
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
timeout 120 ncu --metrics $M --clock-control none -k regex:raster_kernel -s 6 -c 1 python profiles/time_raster.py 2>&1 | grep -E "no_instruction|long_score|issue_active|duration|inst_executed"
