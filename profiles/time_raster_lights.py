"""Raster timing of a 64x64 scene WITH traffic lights and two agent types (10 active classes: the 5-bit rank kernels):
the scene of BASELINE config 3 (Town01 / Town02 / Town10HD by environment, 96 vehicles + 32 pedestrians) at 64x64.
Usage: python profiles/time_raster_lights.py [B]"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchdrivesim_b200 as tds  # noqa: E402

spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "profiles", "bench_configs.py"))
bc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bc)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A, res = 128, 64
dev = torch.device("cuda:0")
sim, act, _ = bc.build_config3(dev, B, A)
out = torch.empty(B, A, 3, res, res, device=dev)
r = tds.Resolution(res, res)
ms = bc.raster_ms(lambda: sim.render_egocentric(out=out, res=r), 10)
gbs = B * A * 12 * res * res / ms / 1e6
print(f"{B} x {A} cameras of {res}x{res}, traffic lights, 3 maps: raster {ms:.3f} ms = {gbs:.0f} GB/s ({100 * gbs / 6545.3:.1f} % of measured HBM), "
      f"checksum {float(out.sum()):.1f}")
