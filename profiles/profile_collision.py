"""Collision / kinematic workloads for ncu at the sizes of BASELINE configs 3-5 (per GPU of an 8-GPU job):
IoU all-pairs 64 envs x 512 agents (config 4), discs all-pairs 512 x 128 (config 3), discs forward + backward and the
kinematic step forward + backward at 256 x 64 (config 5).  Usage: ncu ... python profiles/profile_collision.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdrivesim_b200 as tds  # noqa: E402
from torchdrivesim_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def boxes(B, A, span):
    xy = rng.uniform(-span / 2, span / 2, (B, A, 2))
    return torch.tensor(np.concatenate([xy, np.tile([4.97, 2.04], (B, A, 1)), rng.uniform(0, 6.28, (B, A, 1))], -1),
                        dtype=torch.float32, device=dev)


def timed(name, fn, n=5):
    if os.environ.get("TDS_PROFILE_ONCE"):      # under `ncu --set full`: every workload once
        fn()
        torch.cuda.synchronize()
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n:.3f} ms")


b4 = boxes(64, 512, 120.0)            # config 4: dense, agents confined to a 120 m box
m4 = torch.ones(64, 512, dtype=torch.bool, device=dev)
timed("iou allpairs 64x512x512 (16.8 M pairs)", lambda: tds.ops.collision_allpairs(b4, b4, m4, _lib.METRIC_IOU))
b3 = boxes(512, 128, 300.0)
m3 = torch.ones(512, 128, dtype=torch.bool, device=dev)
timed("discs allpairs 512x128x128 (8.4 M pairs)", lambda: tds.ops.collision_allpairs(b3, b3, m3, _lib.METRIC_DISCS))
b5 = boxes(256, 64, 200.0).requires_grad_(True)
m5 = torch.ones(256, 64, dtype=torch.bool, device=dev)


def fwd_bwd():
    b5.grad = None
    tds.ops.collision_allpairs(b5, b5, m5, _lib.METRIC_DISCS).sum().backward()


timed("discs allpairs fwd+bwd 256x64x64", fwd_bwd)
km = tds.KinematicBicycle(left_handed=True)
km.set_params(lr=torch.full((256, 64), 1.96, device=dev))
s0 = torch.tensor(rng.uniform(-1, 1, (256, 64, 4)), dtype=torch.float32, device=dev)
act = torch.tensor(rng.uniform(-1, 1, (256, 64, 2)), dtype=torch.float32, device=dev, requires_grad=True)


def rollout():
    act.grad = None
    km.set_state(s0)
    for _ in range(20):
        km.step(act)
    km.get_state().pow(2).sum().backward()


timed("bicycle 20-step rollout fwd+bwd 256x64", rollout)
