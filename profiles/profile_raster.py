"""Small raster-only workload for `ncu --set full` (one GPU, short): B environments x 64 cameras at 64x64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torchdrivesim_b200 as tds  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = bench.AGENTS
dev = torch.device("cuda:0")
state, size, lr, actions = bench.synth_inputs(B, A, 1000, 1)
town = tds.StaticMap.from_npz(bench.map_npz())
km = tds.KinematicBicycle(left_handed=True)
km.set_params(lr=torch.tensor(lr, device=dev))
km.set_state(torch.tensor(state, device=dev))
sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                    tds.TorchDriveConfig(left_handed_coordinates=True))
out = torch.empty(B, A, 3, bench.RES, bench.RES, device=dev)
for _ in range(3):
    sim.step(torch.tensor(actions[0], device=dev))
    sim.render_egocentric(out=out)
    sim.compute_collision()
    sim.compute_offroad()
torch.cuda.synchronize()
print("done", float(out.mean()))
