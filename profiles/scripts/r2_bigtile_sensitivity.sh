# what the border-crossing faces and the resolve cost the one-pass large-tile kernels (experiment builds: wrong / no images)
M=smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum
for l in libtds_b200.so libtds_noclip.so libtds_nores.so libtds_both.so; do
  echo "== $l"
  for cfg in "128 256 128" "256 64 128"; do
    TDS_RASTER_TWO_PASS=0 TDS_B200_LIB=$PWD/torchdrivesim_b200/_build/$l timeout 200 ncu --metrics $M --clock-control none -k regex:raster_kernel -s 6 -c 1 python profiles/time_raster_res.py $cfg 2>&1 | grep -E "no_instruction|barrier|issue_active|duration|inst_executed" | tr -s ' ' | tr '\n' ';'; echo
  done
done
