# usage: bash profiles/scripts/r2_scale.sh "2 4 8"   (torchrun, one rank per GPU; results in gpurun_out/bench_nN.json)
rm -rf gpurun_out/*
for N in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N exit $?"; tail -c 300 gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_reference_n2.json 2> gpurun_out/bench_reference_n2.err; echo "ref arm exit $?"; cut -c1-200 gpurun_out/bench_reference_n2.json
