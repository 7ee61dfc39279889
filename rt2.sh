python -m pytest tests/test_gpu_noise.py tests/test_gpu_observations.py tests/test_gpu_npc.py -x -q 2>&1 | tail -25
