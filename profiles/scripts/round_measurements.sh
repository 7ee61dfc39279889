# the end-of-round measurement run: GPU tests, bench (both arms), ncu launch list, raster capture, collision / kinematic pipe metrics
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_r1_final python profiles/time_raster.py > gpurun_out/p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:offroad_fwd -s 5 -c 1 -o gpurun_out/offroad_r1_final python profiles/exp_offroad.py > gpurun_out/po.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"allpairs|kin_" --csv --log-file gpurun_out/collision_r1.csv python profiles/profile_collision.py > gpurun_out/pc.log 2>&1
cut -c1-1200 gpurun_out/bench_n1.json
