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
town = tds.StaticMap.from_npz(bench.map_npz(), raster_cell=float(os.environ.get("TDS_RASTER_CELL", "16")))
km = tds.KinematicBicycle(left_handed=True)
km.set_params(lr=torch.tensor(lr, device=dev))
km.set_state(torch.tensor(state, device=dev))
sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                    tds.TorchDriveConfig(left_handed_coordinates=True))
out = torch.empty(B, A, 3, bench.RES, bench.RES, device=dev)
for _ in range(3):
    sim.render_egocentric(out=out)
torch.cuda.synchronize()
# the raster kernel alone: CUDA events recorded around its launch by the library (an eager render call costs ~1.5 ms of
# host time, which would hide any kernel faster than that)
lib = tds._lib.load()
n = 10
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for a, b in evs:
    a.record(); b.record()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for a, b in evs:
    lib.tds_raster_set_timing_events(a.cuda_event, b.cuda_event)
    sim.render_egocentric(out=out)
lib.tds_raster_set_timing_events(None, None)
e1.record()
torch.cuda.synchronize()
ms_call = e0.elapsed_time(e1) / n
ms = sum(a.elapsed_time(b) for a, b in evs) / n
print(f"{os.environ.get('TDS_B200_LIB', 'default'):40s} raster kernel {ms:.3f} ms  ({B * A * 12 * bench.RES ** 2 / ms / 1e6:.0f} GB/s, "
      f"{100 * B * A * 12 * bench.RES ** 2 / ms / 1e6 / 6545.3:.1f} % of measured HBM)  render call {ms_call:.3f} ms  checksum {float(out.sum()):.1f}")
