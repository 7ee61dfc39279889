M=smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread
for l in "$@"; do
  echo "== $l"
  TDS_B200_LIB=$PWD/torchdrivesim_b200/_build/$l python profiles/time_raster.py
  TDS_B200_LIB=$PWD/torchdrivesim_b200/_build/$l timeout 120 ncu --metrics $M --clock-control none -k regex:raster -s 12 -c 2 python profiles/time_raster.py 2>&1 | grep -E "raster_|issue_active|duration|inst_executed|registers"
done
