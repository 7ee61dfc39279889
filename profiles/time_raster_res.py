"""Raster timing at other tile sizes / agent counts (configs 3 and 4 of BASELINE.json).
Usage: python profiles/time_raster_res.py RES B A [FOV]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torchdrivesim_b200 as tds  # noqa: E402

RES, B, A = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
FOV = float(sys.argv[4]) if len(sys.argv) > 4 else 35.0
dev = torch.device("cuda:0")
state, size, lr, actions = bench.synth_inputs(B, A, 1000, 1)
town = tds.StaticMap.from_npz(bench.map_npz())
km = tds.KinematicBicycle(left_handed=True)
km.set_params(lr=torch.tensor(lr, device=dev))
km.set_state(torch.tensor(state, device=dev))
sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                    tds.TorchDriveConfig(left_handed_coordinates=True))
res = tds.Resolution(RES, RES)
out = torch.empty(B, A, 3, RES, RES, device=dev)
for _ in range(3):
    sim.render_egocentric(out=out, res=res, fov=FOV)
torch.cuda.synchronize()
lib = tds._lib.load()
n = 5
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for a, b in evs:
    a.record(); b.record()
torch.cuda.synchronize()
for a, b in evs:
    lib.tds_raster_set_timing_events(a.cuda_event, b.cuda_event)
    sim.render_egocentric(out=out, res=res, fov=FOV)
lib.tds_raster_set_timing_events(None, None)
torch.cuda.synchronize()
ms = sum(a.elapsed_time(b) for a, b in evs) / n
gb = B * A * 12 * RES * RES / 1e9
print(f"res {RES} B {B} A {A} fov {FOV}: raster kernel {ms:.3f} ms, {gb:.2f} GB out, {gb / ms * 1e3:.0f} GB/s "
      f"({100 * gb / ms * 1e3 / 6545.3:.1f} % of measured HBM), {B * A / ms * 1e3 / 1e6:.2f} M cameras/s")
