# round 2, final tree: the GPU tests, the two bench arms, the launch list and one --set full capture of the two raster passes
# (no profiler around any bench value; the summaries are made here so that only small files travel back)
rm -rf gpurun_out/*
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster -s 12 -c 2 -o gpurun_out/raster_two_pass python profiles/time_raster.py > gpurun_out/p_cap.log 2>&1
python profiles/make_ncu_summary.py raster2 gpurun_out/raster_two_pass.ncu-rep gpurun_out/r2_raster_ncu_summary.json
python profiles/time_raster.py; TDS_RASTER_TWO_PASS=0 python profiles/time_raster.py
python profiles/time_raster_res.py 128 256 128; python profiles/time_raster_res.py 256 64 128
rm -f gpurun_out/raster_two_pass.ncu-rep
cut -c1-1200 gpurun_out/bench_n1.json
du -sh gpurun_out
