# round 2, final tree: the GPU tests, the two bench arms and the launch list (no profiler around any bench value)
rm -rf gpurun_out/*
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 96 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
cut -c1-1500 gpurun_out/bench_n1.json; cut -c1-600 gpurun_out/bench_reference.json
du -sh gpurun_out
