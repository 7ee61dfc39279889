import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torchdrivesim_b200 as tds
dev = torch.device("cuda:0")
B, A = 1024, 64
state, size, lr, actions = bench.synth_inputs(B, A, 1000, 3)
town = tds.StaticMap.from_npz(bench.map_npz(), offroad_cell=float(os.environ.get("TDS_OFFROAD_CELL", "2")))
import numpy as np
sigma = float(os.environ.get("TDS_OFFROAD_SIGMA", "3"))       # displacement off the road vertices (m)
state[..., :2] += np.random.default_rng(0).normal(0, sigma, state[..., :2].shape).astype(np.float32)
st, sz = torch.tensor(state, device=dev), torch.tensor(size, device=dev)
ms = tds.MapSet([town])
for _ in range(3): tds.ops.offroad(st, sz, ms, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): tds.ops.offroad(st, sz, ms, 0.5)
e1.record(); torch.cuda.synchronize()
print("sigma", sigma, "offroad cell", os.environ.get("TDS_OFFROAD_CELL", "2"), "ms", e0.elapsed_time(e1) / 20)
