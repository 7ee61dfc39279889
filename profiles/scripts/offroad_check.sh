# offroad parity tests + kernel time under ncu for agents displaced by sigma off the road
python -m pytest tests/test_gpu_offroad.py -x -q 2>&1 | tail -3
for s in 0 3 10; do
TDS_OFFROAD_SIGMA=$s ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:offroad_fwd -s 5 -c 1 python profiles/exp_offroad.py 2>&1 | grep -E "gpu__time|sigma|warps_active|issue_active|thread_inst"
done
