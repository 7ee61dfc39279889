# bench lines + launch list of the final tree
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 200 gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
cut -c1-300 gpurun_out/bench_n1.json
