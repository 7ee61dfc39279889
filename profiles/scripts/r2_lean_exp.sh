M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
for lib in libtds_b200.so libtds_lean.so; do
  echo "== $lib"
  TDS_B200_LIB=$PWD/torchdrivesim_b200/_build/$lib ncu --metrics $M --clock-control none -k regex:raster_kernel -s 3 -c 1 python profiles/time_raster.py 2>&1 | grep -E "no_instruction|issue_active|duration|inst_executed|render"
done
