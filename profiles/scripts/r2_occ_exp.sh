M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
for pad in 0 12000 30000 60000; do
  echo "== pad $pad"
  TDS_RASTER_PAD_SMEM=$pad ncu --metrics $M --clock-control none -k regex:raster_kernel -s 3 -c 1 python profiles/time_raster.py 2>&1 | grep -E "no_instruction|issue_active|warps_active|duration|inst_executed|render"
done
