"""Configs 3-5 of BASELINE.json at the size ONE GPU of the 8-GPU job holds, timed with CUDA events.
These are parity-test configurations, not bench.py lines; this script records what they cost on a B200.

  config 3: Town01 / Town02 / Town10HD alternating by environment, 512 envs x 128 agents (96 bicycle vehicles + 32 unicycle
            pedestrians), traffic lights cycling every 30 steps, 128x128 birdviews, discs collisions + offroad
  config 4: Town01, 64 envs x 512 agents inside a 120 m box, 256x256 birdviews, IoU collisions + offroad
  config 5: Town01, 256 envs x 64 agents, 20-step rollout, loss = collisions (discs) + offroad + MSE, backward to actions
Usage: python profiles/bench_configs.py [3 4 5]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdrivesim_b200 as tds  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda:0")
VEH, PED = (4.97, 2.04, 1.96), (1.5, 1.5)
HBM = 6545.3


def load(name):
    path = os.path.join(ROOT, "tests", "golden", "maps", name + ".npz")
    d = np.load(path)
    cats = [str(c) for c in d["categories"]]
    return tds.StaticMap.from_npz(path), d["verts"][d["vert_category"] == cats.index("road")]


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def config3():
    rng = np.random.default_rng(3)
    B, A, res = 512, 128, 128
    maps = [load("carla_Town01"), load("carla_Town02"), load("carla_Town10HD")]     # Town10HD is built from its OSM file
    M = len(maps)
    env_map = torch.arange(B, dtype=torch.int32) % M
    xy = np.stack([maps[b % M][1][rng.integers(0, len(maps[b % M][1]), A)] for b in range(B)])
    state = np.concatenate([xy, rng.uniform(0, 6.28, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32)
    types = torch.tensor((np.arange(A) >= 96).astype(np.int64)).expand(B, A).contiguous().to(dev)
    size = torch.where(types[..., None] == 1, torch.tensor(PED, device=dev), torch.tensor(VEH[:2], device=dev))
    km = tds.FusedCompoundKinematicModel(torch.where(types == 1, 2, 0).to(torch.int32), left_handed=True)
    km.set_params(lr=torch.full((B, A), VEH[2], device=dev))
    km.set_state(torch.tensor(state, device=dev))
    poses = [m.traffic_light_poses() for m, _ in maps]
    L = max(len(p) for p in poses)
    pos = np.zeros((B, L, 5), np.float32)
    mask = np.zeros((B, L), bool)
    for b in range(B):
        p = poses[b % M]
        pos[b, :len(p)], mask[b, :len(p)] = p, True
    replay = torch.tensor((np.arange(120)[None, None] // 30 + rng.integers(0, 3, (B, L, 1))) % 3, device=dev)
    tl = tds.TrafficLightControl(torch.tensor(pos, device=dev), replay_states=replay, mask=torch.tensor(mask, device=dev))
    sim = tds.Simulator(tds.MapSet([m for m, _ in maps], env_map.to(dev)), km, size, torch.ones(B, A, dtype=torch.bool, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls={"traffic_light": tl},
                        agent_types=types, agent_type_names=["vehicle", "pedestrian"])
    out = torch.empty(B, A, 3, res, res, device=dev)
    act = torch.tensor(rng.uniform(-1, 1, (B, A, 2)).astype(np.float32), device=dev)
    r = tds.Resolution(res, res)

    def step():
        sim.step(act)
        sim.render_egocentric(out=out, res=r)
        sim.compute_collision()
        sim.compute_offroad()
        sim.compute_traffic_lights_violations()

    ms = timed(step, 5)
    ms_r = timed(lambda: sim.render_egocentric(out=out, res=r), 5)
    report("config 3 (1/8 shard)", B, A, res, ms, ms_r)


def config4():
    rng = np.random.default_rng(4)
    B, A, res = 64, 512, 256
    town, road = load("carla_Town01")
    centre = road[rng.integers(0, len(road), B)]
    near = [road[np.abs(road - c).max(1) < 60.0] for c in centre]
    xy = np.stack([n[rng.integers(0, len(n), A)] for n in near])
    state = np.concatenate([xy, rng.uniform(0, 6.28, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32)
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), VEH[2], device=dev))
    km.set_state(torch.tensor(state, device=dev))
    sim = tds.Simulator(town, km, torch.tensor(VEH[:2], device=dev).expand(B, A, 2).contiguous(),
                        torch.ones(B, A, dtype=torch.bool, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True, collision_metric=tds.CollisionMetric.iou))
    out = torch.empty(B, A, 3, res, res, device=dev)
    act = torch.tensor(rng.uniform(-1, 1, (B, A, 2)).astype(np.float32), device=dev)
    r = tds.Resolution(res, res)

    def step():
        sim.step(act)
        sim.render_egocentric(out=out, res=r)
        sim.compute_collision()
        sim.compute_offroad()

    ms = timed(step, 3)
    ms_r = timed(lambda: sim.render_egocentric(out=out, res=r), 3)
    ms_c = timed(sim.compute_collision, 5)
    report("config 4 (1/8 shard)", B, A, res, ms, ms_r, f", IoU all-pairs {ms_c:.3f} ms ({B * A * A / 1e6:.1f} M pairs)")


def config5():
    rng = np.random.default_rng(5)
    B, A, T = 256, 64, 20
    town, road = load("carla_Town01")
    xy = road[rng.integers(0, len(road), (B, A))]
    state0 = torch.tensor(np.concatenate([xy, rng.uniform(0, 6.28, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32), device=dev)
    target = state0[..., :2] + 5.0
    actions = torch.tensor(rng.uniform(-1, 1, (T, B, A, 2)).astype(np.float32), device=dev, requires_grad=True)
    size = torch.tensor(VEH[:2], device=dev).expand(B, A, 2).contiguous()
    present = torch.ones(B, A, dtype=torch.bool, device=dev)

    def rollout():
        actions.grad = None
        km = tds.KinematicBicycle(left_handed=True)
        km.set_params(lr=torch.full((B, A), VEH[2], device=dev))
        km.set_state(state0)
        sim = tds.Simulator(town, km, size, present, tds.TorchDriveConfig(left_handed_coordinates=True))
        loss = 0.0
        for t in range(T):
            sim.step(actions[t])
            loss = loss + sim.compute_collision().sum() + sim.compute_offroad().sum() + ((sim.get_state()[..., :2] - target) ** 2).mean()
        loss.backward()

    ms = timed(rollout, 3)
    print(f"config 5 (1/8 shard): {B} envs x {A} agents, {T}-step rollout forward + backward {ms:.2f} ms "
          f"= {B * A * T / ms * 1e3 / 1e6:.2f} M differentiable agent-env-steps/s, grad finite: {bool(torch.isfinite(actions.grad).all())}")


def report(name, B, A, res, ms, ms_r, extra=""):
    gb = B * A * 12 * res * res / 1e9
    print(f"{name}: {B} envs x {A} agents, {res}x{res}: step {ms:.2f} ms = {B * A / ms * 1e3 / 1e6:.2f} M agent-env-steps/s "
          f"({100 * B * A * (12 * res * res + 61) / 1e9 / ms * 1e3 / HBM:.1f} % of the HBM roofline); raster {ms_r:.2f} ms = "
          f"{gb / ms_r * 1e3:.0f} GB/s ({100 * gb / ms_r * 1e3 / HBM:.1f} %){extra}")


if __name__ == "__main__":
    for c in (sys.argv[1:] or ["3", "4", "5"]):
        {"3": config3, "4": config4, "5": config5}[c]()
