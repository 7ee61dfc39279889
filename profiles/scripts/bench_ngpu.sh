# usage: bash profiles/scripts/bench_ngpu.sh N   (torchrun, one rank per GPU; result in gpurun_out/bench_nN.json)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 300 gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json
