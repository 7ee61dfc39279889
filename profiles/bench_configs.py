"""Configs 3-5 of BASELINE.json at the size ONE GPU of the 8-GPU job holds, timed with CUDA events.
bench.py imports this module and carries the results in the `configs` block of its JSON line; run directly it prints
them.

  config 3: Town01 / Town02 / Town10HD alternating by environment, 512 envs x 128 agents (96 bicycle vehicles + 32 unicycle
            pedestrians), traffic lights cycling every 30 steps, 128x128 birdviews, discs collisions + offroad + red-light
            violations
  config 4: Town01, 64 envs x 512 agents inside a 120 m box, 256x256 birdviews, IoU collisions + offroad
  config 5: Town01, 256 envs x 64 agents, 20-step rollout, loss = collisions (discs) + offroad + MSE, backward to actions
Usage: python profiles/bench_configs.py [3 4 5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torchdrivesim_b200 as tds  # noqa: E402

VEH, PED = (4.97, 2.04, 1.96), (1.5, 1.5)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def load(name):
    path = os.path.join(ROOT, "tests", "golden", "maps", name + ".npz")
    d = np.load(path)
    cats = [str(c) for c in d["categories"]]
    return tds.StaticMap.from_npz(path), d["verts"][d["vert_category"] == cats.index("road")]


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def raster_ms(render, n):
    """Duration of the raster kernel alone (CUDA events recorded around the launch by the library)."""
    lib = tds._lib.load()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); b.record()
    torch.cuda.synchronize()
    for a, b in evs:
        lib.tds_raster_set_timing_events(a.cuda_event, b.cuda_event)
        render()
    lib.tds_raster_set_timing_events(None, None)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / n


def build_config3(dev, B=512, A=128, seed=3):
    rng = np.random.default_rng(seed)
    maps = [load("carla_Town01"), load("carla_Town02"), load("carla_Town10HD")]     # Town10HD is built from its OSM file
    M = len(maps)
    env_map = torch.arange(B, dtype=torch.int32) % M
    xy = np.stack([maps[b % M][1][rng.integers(0, len(maps[b % M][1]), A)] for b in range(B)])
    state = np.concatenate([xy, rng.uniform(0, 6.28, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32)
    n_ped = A // 4
    types = torch.tensor((np.arange(A) >= A - n_ped).astype(np.int64)).expand(B, A).contiguous().to(dev)
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
    act = torch.tensor(rng.uniform(-1, 1, (B, A, 2)).astype(np.float32), device=dev)
    return sim, act, state


def config3(dev, B=512, A=128, res=128, n=5):
    sim, act, _ = build_config3(dev, B, A)
    out = torch.empty(B, A, 3, res, res, device=dev)
    r = tds.Resolution(res, res)

    def step():
        sim.step(act)
        sim.render_egocentric(out=out, res=r)
        sim.compute_collision()
        sim.compute_offroad()
        sim.compute_traffic_lights_violations()

    ms = timed(step, n)
    ms_r = raster_ms(lambda: sim.render_egocentric(out=out, res=r), n)
    return report("config 3 at the shard of one of 8 GPUs: Town01/Town02/Town10HD by env_map, 96 bicycle + 32 unicycle agents, "
                  "traffic lights, discs + offroad + red-light violations", B, A, res, ms, ms_r)


def build_config4(dev, B=64, A=512, seed=4):
    rng = np.random.default_rng(seed)
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
    act = torch.tensor(rng.uniform(-1, 1, (B, A, 2)).astype(np.float32), device=dev)
    return sim, act, state


def config4(dev, B=64, A=512, res=256, n=3):
    sim, act, _ = build_config4(dev, B, A)
    out = torch.empty(B, A, 3, res, res, device=dev)
    r = tds.Resolution(res, res)

    def step():
        sim.step(act)
        sim.render_egocentric(out=out, res=r)
        sim.compute_collision()
        sim.compute_offroad()

    ms = timed(step, n)
    ms_r = raster_ms(lambda: sim.render_egocentric(out=out, res=r), n)
    ms_c = timed(sim.compute_collision, 5)
    d = report("config 4 at the shard of one of 8 GPUs: Town01, 512 agents on the roads of a 120 m box, IoU collisions + offroad",
               B, A, res, ms, ms_r)
    d["iou_allpairs_ms"] = ms_c
    d["iou_pairs"] = B * A * A
    return d


def config5(dev, B=256, A=64, T=20, n=5):
    rng = np.random.default_rng(5)
    town, road = load("carla_Town01")
    xy = road[rng.integers(0, len(road), (B, A))]
    state0 = torch.tensor(np.concatenate([xy, rng.uniform(0, 6.28, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32), device=dev)
    target = (state0[..., :2] + 5.0).contiguous()
    actions = torch.tensor(rng.uniform(-1, 1, (T, B, A, 2)).astype(np.float32), device=dev)
    size = torch.tensor(VEH[:2], device=dev).expand(B, A, 2).contiguous()
    lr = torch.full((B, A), VEH[2], device=dev)
    present = torch.ones(B, A, dtype=torch.bool, device=dev)

    # eager autograd through the Simulator API (examples/imitation_learning.py-style)
    a_eager = actions.clone().requires_grad_(True)

    def rollout_eager():
        a_eager.grad = None
        km = tds.KinematicBicycle(left_handed=True)
        km.set_params(lr=lr)
        km.set_state(state0)
        sim = tds.Simulator(town, km, size, present, tds.TorchDriveConfig(left_handed_coordinates=True))
        loss = 0.0
        for t in range(T):
            sim.step(a_eager[t])
            loss = loss + sim.compute_collision().sum() + sim.compute_offroad().sum() + ((sim.get_state()[..., :2] - target) ** 2).mean()
        loss.backward()
        return loss

    ms_eager = timed(rollout_eager, 3)
    out = {"workload": f"config 5 at the shard of one of 8 GPUs: Town01, {B} envs x {A} agents, {T}-step rollout, loss = discs "
                       f"collisions + offroad(0.5) + MSE to a target, forward + backward to the actions",
           "envs": B, "agents": A, "steps_per_rollout": T, "eager_ms_per_rollout": ms_eager,
           "eager_agent_env_steps_per_s": B * A * T / ms_eager * 1e3}
    if hasattr(tds, "FusedRollout"):
        ro = tds.FusedRollout(town, state0, size, lr, present, T, offroad_threshold=0.5, left_handed=True, target_xy=target)
        loss_f, grad_f = ro.run(actions)
        loss_e = rollout_eager()
        ms = timed(lambda: ro.run(actions), n)
        out.update({"ms_per_rollout": ms, "agent_env_steps_per_s": B * A * T / ms * 1e3,
                    "loss_rel_diff_vs_eager": float((loss_f - loss_e).abs() / loss_e.abs()),
                    "grad_max_abs_diff_vs_eager": float((grad_f - a_eager.grad).abs().max()),
                    "grad_finite": bool(torch.isfinite(grad_f).all())})
    else:
        out.update({"ms_per_rollout": ms_eager, "agent_env_steps_per_s": B * A * T / ms_eager * 1e3})
    return out


def report(name, B, A, res, ms, ms_r):
    peak = hbm_peak()
    gb = B * A * 12 * res * res / 1e9
    return {"workload": name, "envs": B, "agents": A, "res": res, "ms_per_step": ms, "agent_env_steps_per_s": B * A / ms * 1e3,
            "step_frac_of_hbm_roofline": B * A * (12 * res * res + 61) / 1e9 / ms * 1e3 / peak,
            "raster_ms_per_launch": ms_r, "raster_gbs": gb / ms_r * 1e3, "raster_frac_of_hbm_roofline": gb / ms_r * 1e3 / peak}


def run(which=("3", "4", "5"), dev=None):
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    out = {}
    for c in which:
        out[c] = {"3": config3, "4": config4, "5": config5}[c](dev)
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    for k, v in run(sys.argv[1:] or ["3", "4", "5"]).items():
        print(k, json.dumps(v))
