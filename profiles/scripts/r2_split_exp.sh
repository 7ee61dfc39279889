M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
for only in 0 1 2 3; do
  echo "== only $only"
  TDS_RASTER_ONLY=$only python profiles/time_raster.py 2>&1 | grep render
  TDS_RASTER_ONLY=$only ncu --metrics $M --clock-control none -k regex:raster_kernel -s 3 -c 1 python profiles/time_raster.py 2>&1 | grep -E "no_instruction|issue_active|duration|inst_executed"
done
