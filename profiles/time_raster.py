"""Raster-only timing (CUDA events, 65536 cameras at 64x64, bench inputs).  Usage: TDS_B200_LIB=... python profiles/time_raster.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torchdrivesim_b200 as tds  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = bench.AGENTS
dev = torch.device("cuda:0")
state, size, lr, actions = bench.synth_inputs(B, A, 1000, 1)
town = tds.StaticMap.from_npz(bench.map_npz(), raster_cell=float(os.environ.get("TDS_RASTER_CELL", "8")))
km = tds.KinematicBicycle(left_handed=True)
km.set_params(lr=torch.tensor(lr, device=dev))
km.set_state(torch.tensor(state, device=dev))
sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                    tds.TorchDriveConfig(left_handed_coordinates=True))
out = torch.empty(B, A, 3, bench.RES, bench.RES, device=dev)
for _ in range(3):
    sim.render_egocentric(out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    sim.render_egocentric(out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{os.environ.get('TDS_B200_LIB', 'default'):60s} render {ms:.3f} ms  ({B * A * 12 * bench.RES ** 2 / ms / 1e6:.0f} GB/s)  checksum {float(out.sum()):.1f}")
