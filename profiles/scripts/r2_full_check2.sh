python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -5 gpurun_out/bench_r2a.err; cut -c1-3000 gpurun_out/bench_r2a.json
