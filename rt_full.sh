set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --kernels-only > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 3 -c 1 -o gpurun_out/raster_r1_final python profiles/time_raster.py > gpurun_out/p.log 2>&1
cat gpurun_out/bench_n1.json | cut -c1-1500
