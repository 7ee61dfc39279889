python -m pytest tests/test_gpu_goals.py tests/test_gpu_npc.py tests/test_gpu_graph.py -x -q 2>&1 | tail -25
