python -m pytest tests/test_gpu_offroad.py -x -q 2>&1 | tail -3
for s in 0 3 10; do
TDS_OFFROAD_SIGMA=$s ncu --metrics gpu__time_duration.sum --clock-control none -k regex:offroad_fwd -s 5 -c 1 python profiles/exp_offroad.py 2>&1 | grep -E "gpu__time|sigma"
done
bash rt.sh
