rm -rf gpurun_out/*
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "N=8 exit $?"; tail -c 300 gpurun_out/bench_n8.err; cut -c1-400 gpurun_out/bench_n8.json
