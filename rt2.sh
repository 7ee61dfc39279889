python -m pytest tests/test_gpu_collision.py tests/test_gpu_npc.py -x -q 2>&1 | tail -5
python profiles/profile_collision.py 2>&1 | head -3
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"allpairs_fwd" -c 1 python profiles/profile_collision.py 2>&1 | grep -E "gpu__time|pipe_fma|issue_active|thread_inst"
