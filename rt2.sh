python -m pytest tests/test_gpu_graph.py tests/test_gpu_npc.py -x -q 2>&1 | tail -25
