# the round-2 measurement run: GPU tests, bench (both arms), ncu launch list, captures of the raster, offroad and collision kernels.
# Every profiler pass is bounded (timeout) and keeps its report small: gpurun_out/ comes back only below 64 MiB.
set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_ref.err
TDS_BENCH_NO_SAMPLER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_ncu.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 6 -c 1 -o gpurun_out/raster_r2_final python profiles/time_raster.py > gpurun_out/p.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:offroad_fwd -s 5 -c 1 -o gpurun_out/offroad_r2 python profiles/exp_offroad.py > gpurun_out/po.log 2>&1
TDS_PROFILE_ONCE=1 timeout 300 ncu --set full --clock-control none -k regex:"allpairs|kin_" -c 50 -o gpurun_out/collision_r2 python profiles/profile_collision.py > gpurun_out/pc.log 2>&1
# summaries are made on the box; only the raster report (10 MB) travels back (gpurun_out/ must stay below 64 MiB)
python profiles/make_ncu_summary.py raster gpurun_out/raster_r2_final.ncu-rep gpurun_out/r2_raster_ncu_summary.json
python profiles/make_ncu_summary.py kernels gpurun_out/offroad_r2.ncu-rep gpurun_out/r2_offroad_ncu_summary.json
python profiles/make_ncu_summary.py kernels gpurun_out/collision_r2.ncu-rep gpurun_out/r2_collision_ncu_summary.json
rm -f gpurun_out/collision_r2.ncu-rep gpurun_out/offroad_r2.ncu-rep
ls -la gpurun_out | tail -14
cut -c1-1500 gpurun_out/r2_bench_n1.json
