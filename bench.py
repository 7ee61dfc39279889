#!/usr/bin/env python
"""Benchmark of the TorchDriveSim per-step hot path on B200 (BASELINE.json metric).

One "step" = kinematic step -> egocentric 64x64 birdview for every agent -> pairwise collisions (discs)
-> offroad, over one batch of synthetic input: config 2 of BASELINE.json (carla_Town01 standing in for
the missing Town03 mesh, 1024 environments x 64 agents, bicycle model) PER GPU (weak scaling).

  python bench.py [--gpus N --steps K --warmup W]            our CUDA path  (torchrun for N > 1)
  python bench.py --impl reference [...]                      the CPU port of the reference (oracle), all host cores

Prints ONE JSON line on rank 0 (see the keys in main()).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAP = "carla_Town01"
ENVS_PER_GPU = 1024
AGENTS = 64
RES = 64
FOV = 35.0
VEH = (4.97, 2.04, 1.96)
BYTES_PER_AGENT_STEP = 12 * RES * RES + 61        # fp32 RGB image + state/action/size/mask/outputs
METRIC = "agent-env-steps/sec incl. 64x64 BEV render+collisions"


def map_npz():
    return os.path.join(ROOT, "tests", "golden", "maps", f"{MAP}.npz")


def synth_inputs(B, A, seed, steps):
    """Seeded synthetic inputs of SURVEY.md §8(d): on-road positions, psi ~ U[0,2pi), v ~ U[0,5], vehicles."""
    rng = np.random.default_rng(seed)
    d = np.load(map_npz())
    cats = [str(c) for c in d["categories"]]
    road = d["verts"][d["vert_category"] == cats.index("road")]
    xy = road[rng.integers(0, road.shape[0], (B, A))]
    state = np.concatenate([xy, rng.uniform(0, 2 * np.pi, (B, A, 1)), rng.uniform(0, 5, (B, A, 1))], -1).astype(np.float32)
    size = np.tile(np.array(VEH[:2], np.float32), (B, A, 1))
    lr = np.full((B, A), VEH[2], np.float32)
    actions = rng.uniform(-1, 1, (steps, B, A, 2)).astype(np.float32)
    return state, size, lr, actions


# ------------------------------------------------------------------------------------------------ CPU port
def _cpu_env_step(args):
    """One environment, one step of the hot path with the oracle (the CPU port of the reference)."""
    import torch
    torch.set_num_threads(1)
    from oracle import collision as OC, kinematic as OK, offroad as OO, raster as OR
    state, size, lr, action, m = args
    st = OK.bicycle_step(torch.tensor(state), torch.tensor(action), torch.tensor(lr), 0.1, True).numpy()
    sc = OR.build_scene(m["verts"], m["faces"], m["face_cat"], st, size, ["vehicle"], None, None)
    cam_sc = torch.stack([torch.sin(torch.tensor(st[:, 2])), torch.cos(torch.tensor(st[:, 2]))], -1).numpy()
    chk = 0.0
    for a in range(st.shape[0]):
        img, _ = OR.render_camera(sc, st[a, :2], cam_sc[a], RES, FOV)
        chk += float(img[0, 0, 0])
    box = torch.cat([torch.tensor(st[:, :2]), torch.tensor(size), torch.tensor(st[:, 2:3])], -1)[None]
    coll = OC.collision_allpairs(box, box, torch.ones(1, st.shape[0], dtype=torch.bool))
    off = OO.offroad_loss(st, size, m["verts"], m["faces"], 0.5)
    return float(coll.sum()) + float(off.sum()) + chk


_CPU_MAP = None


def _cpu_map():
    global _CPU_MAP
    if _CPU_MAP is None:
        d = np.load(map_npz())
        cats = [str(c) for c in d["categories"]]
        _CPU_MAP = dict(verts=d["verts"], faces=d["faces"], face_cat=[cats[i] for i in d["vert_category"][d["faces"][:, 0]]])
    return _CPU_MAP


def _cpu_worker(task):
    state, size, lr, action = task
    return _cpu_env_step((state, size, lr, action, _cpu_map()))


def cpu_port_throughput(sample_envs, repeats=1, seed=1):
    """agent-env-steps/s of the oracle on `sample_envs` environments spread over all host cores."""
    import oracle
    oracle.build_clib()
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, sample_envs))
    state, size, lr, actions = synth_inputs(sample_envs, AGENTS, seed, 1)
    tasks = [(state[b], size[b], lr[b], actions[0, b]) for b in range(sample_envs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, tasks[:procs])                       # warm-up: page in libs, build caches
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, tasks, chunksize=1)
            times.append(time.perf_counter() - t0)
    return sample_envs * AGENTS / min(times), procs, times


def reference_python_throughput(steps=5, warmup=1, agents=20):
    """The UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) through its own Simulator API on
    BASELINE config 1: one environment, 20 vehicles, bicycle model, 64x64 birdviews with the cv2 renderer,
    step -> render_egocentric -> compute_collision (discs) -> compute_offroad, on the host cores torch uses.
    A bounded sample: `steps` timed steps of the 100 the config names.  None when the install is not there."""
    try:
        from oracle.ref_harness import import_reference, reference_available
        if not reference_available():
            return None
        import torch
        import_reference()
        from torchdrivesim.simulator import Simulator, TorchDriveConfig
        from torchdrivesim.rendering import CV2RendererConfig
        from torchdrivesim.kinematic import KinematicBicycle
        from torchdrivesim.map import find_map_config
        from torchdrivesim.utils import Resolution
    except Exception as e:      # noqa: BLE001
        return {"unavailable": repr(e)[:200]}
    B, A = 1, agents
    state, size, lr, actions = synth_inputs(B, A, 1, warmup + steps)
    cfgm = find_map_config(MAP)
    km = KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.tensor(lr))
    km.set_state(torch.tensor(state))
    cfg = TorchDriveConfig(left_handed_coordinates=True, renderer=CV2RendererConfig(left_handed_coordinates=True))
    sim = Simulator(cfg=cfg, road_mesh=cfgm.road_mesh.expand(B), kinematic_model=km, agent_size=torch.tensor(size),
                    initial_present_mask=torch.ones(B, A, dtype=torch.bool))
    res = Resolution(RES, RES)
    parts = {"step": 0.0, "render": 0.0, "collision": 0.0, "offroad": 0.0}
    t_all = 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        sim.step(torch.tensor(actions[i]))
        t1 = time.perf_counter()
        sim.render_egocentric(res=res, fov=FOV)
        t2 = time.perf_counter()
        sim.compute_collision()
        t3 = time.perf_counter()
        sim.compute_offroad()
        t4 = time.perf_counter()
        if i >= warmup:
            t_all += t4 - t0
            for k, d in zip(parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                parts[k] += d
    return {"value": B * A * steps / t_all, "unit": "agent-env-steps/s", "cores": int(torch.get_num_threads()), "kind": "reference",
            "sample": f"config 1: {MAP}, 1 environment x {A} vehicles, bicycle, {RES}x{RES} cv2 renderer, discs collisions + offroad; "
                      f"{steps} timed steps of 100 through the unmodified reference Simulator (baseline/_ref)",
            "seconds_per_step": t_all / steps, "seconds_per_step_by_part": {k: v / steps for k, v in parts.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_envs = max(8, min(2 * cores, 256))
    import oracle
    oracle.build_clib()
    procs = max(1, min(cores, sample_envs))
    state, size, lr, actions = synth_inputs(sample_envs, AGENTS, 1, 1)
    tasks = [(state[b], size[b], lr[b], actions[0, b]) for b in range(sample_envs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_worker, tasks[:procs])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_worker, tasks, chunksize=1)
        dt = time.perf_counter() - t0
    value = sample_envs * AGENTS * args.steps / dt
    sample = f"{sample_envs} of {ENVS_PER_GPU} environments x {AGENTS} agents per step, one process per core"
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "agent-env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "agent-env-steps/s", "cores": procs, "kind": "port", "sample": sample,
                         "reference_python": reference_python_throughput()},
        "e2e": {"value": value, "unit": "agent-env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus):
    return {"workload": f"config 2: {MAP} (Town03 mesh is a missing blob), {ENVS_PER_GPU} envs x {AGENTS} agents per GPU, "
                        f"bicycle, {RES}x{RES} BEV fov {FOV} m, discs collisions + offroad(0.5)",
            "envs_per_gpu": ENVS_PER_GPU, "agents": AGENTS, "res": RES, "n_gpus": n_gpus,
            "l2_policy": "each step writes 3.2 GB of images per GPU (>> 126 MB L2), which evicts every input; no explicit flush"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region through NVML (nvidia_ml_py) from a
    background thread; falls back to polling `nvidia-smi` when NVML cannot be imported."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.02):
        self.index, self.period, self.rows, self.stop_flag, self.thread, self.err = index, period, [], False, None, None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.perf_counter(), sm, rs, 0.0))
                time.sleep(self.period)
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.max_mhz = None
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"no samples ({self.err})"]}
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        reasons = [n for n, b in self.REASONS.items() if bits & b]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(rows),
                "period_s": self.period}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torchdrivesim_b200 as tds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    B, A = ENVS_PER_GPU, AGENTS

    # ---- this rank's shard of independent environments (weak scaling: ENVS_PER_GPU each)
    state, size, lr, actions = synth_inputs(B, A, 1000 + rank, W + K)
    town = tds.StaticMap.from_npz(map_npz(), raster_cell=float(os.environ.get("TDS_RASTER_CELL", "16")),
                                  offroad_cell=float(os.environ.get("TDS_OFFROAD_CELL", "2")))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.tensor(lr, device=dev))
    km.set_state(torch.tensor(state, device=dev))
    cfg = tds.TorchDriveConfig(left_handed_coordinates=True)
    sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev), cfg)
    state0 = torch.tensor(state, device=dev)
    act_dev = torch.tensor(actions, device=dev)
    images = torch.empty(B, A, 3, RES, RES, dtype=torch.float32, device=dev)
    metrics = torch.zeros(6, dtype=torch.float64, device=dev)     # distributed.METRIC_NAMES
    lib = tds._lib.load()

    def step(action, out_images, ev=None):
        sim.step(action)
        if ev is not None:
            lib.tds_raster_set_timing_events(ev[0].cuda_event, ev[1].cuda_event)
        img = sim.render_egocentric(out=out_images)
        if ev is not None:
            lib.tds_raster_set_timing_events(None, None)
        coll = sim.compute_collision()
        off = sim.compute_offroad()
        return img, coll, off

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- raster kernel duration in situ (eager launches, CUDA events recorded around the kernel by the library)
    for i in range(W):
        step(act_dev[i], images)
    raster_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for e0, e1 in raster_ev:   # create the CUDA events before they are handed to the library
        e0.record(); e1.record()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        step(act_dev[W + i], images, raster_ev[i])
    ev1.record()
    barrier()
    eager_ms = ev0.elapsed_time(ev1) / K
    raster_ms = sum(a.elapsed_time(b) for a, b in raster_ev) / K

    # ---- device-resident throughput ("value"): the whole step replayed as one CUDA graph
    sim.set_state(state0.clone())
    runner = tds.GraphedHotPath(sim, render=True)
    images = runner.images
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("TDS_BENCH_NO_SAMPLER"):
        sampler.start()
    def full_step(i):
        img, coll, off = runner.run(act_dev[i])
        tds.ops.infraction_metrics(coll, off, None, metrics)        # one launch, accumulated over the steps

    # W warm-up steps identical to the timed ones, continued until the GPU has been busy for 0.5 s so that the
    # timed region starts at sustained clocks (lazy module loading, allocator and caches are warm)
    t_spin = time.perf_counter()
    n_spin = 0
    while time.perf_counter() - t_spin < 0.5 or n_spin < W:
        full_step(n_spin % W)
        torch.cuda.synchronize()
        n_spin += 1
    runner.set_state(state0)
    metrics.zero_()
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        full_step(W + i)
    if world > 1:
        dist.all_reduce(metrics)          # the only collective of the path: aggregate infraction metrics
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    # clocks under load: the samples of the timed region plus the identical warm-up spin right before it (the
    # timed region alone lasts ~30 ms, i.e. one or two NVML samples)
    clocks = sampler.stop(t_spin, t_wall1) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * A * K / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned actions in, state/collision/offroad AND the
    # images out to pinned host memory every step
    h_act = torch.tensor(actions).pin_memory()
    h_img = torch.empty(B, A, 3, RES, RES, dtype=torch.float32).pin_memory()
    h_img_u8 = torch.empty(B, A, 3, RES, RES, dtype=torch.uint8).pin_memory()
    h_img_rank = torch.empty(B, A, RES, RES, dtype=torch.uint8).pin_memory()
    h_state = torch.empty(B, A, 4).pin_memory()
    h_coll = torch.empty(B, A).pin_memory()
    h_off = torch.empty(B, A).pin_memory()

    def e2e_step(i, to_host_images):
        if to_host_images is not False:
            a = h_act[i].to(dev, non_blocking=True)
            sim.step(a)
            # chunks of ~100-400 MB: small enough to overlap the copy with the next chunk's raster, large enough that the
            # host-side cost of an eager render call (~1.5 ms) stays hidden
            chunk = 128 if to_host_images.dtype == torch.float32 else (256 if to_host_images.dim() == 5 else 512)
            sim.render_egocentric_to_host(to_host_images, chunk_envs=chunk)
            h_coll.copy_(sim.compute_collision(), non_blocking=True)
            h_off.copy_(sim.compute_offroad(), non_blocking=True)
            h_state.copy_(sim.get_state(), non_blocking=True)
        else:
            _, coll, off = runner.run(h_act[i])          # pinned host actions -> static device buffer -> graph
            h_coll.copy_(coll, non_blocking=True)
            h_off.copy_(off, non_blocking=True)
            h_state.copy_(runner.state, non_blocking=True)

    e2e = {"host_images": float("nan"), "device_images": float("nan"), "host_images_uint8": float("nan"),
           "host_images_rank": float("nan")}
    legs = (("host_images", h_img), ("host_images_uint8", h_img_u8), ("host_images_rank", h_img_rank), ("device_images", False))
    for name, to_host in (() if args.kernels_only else legs):
        if to_host is not False:
            sim.set_state(state0.clone())
        else:
            runner.set_state(state0)
            sim.kinematic_model.set_state(runner.state)
        for i in range(W):
            e2e_step(i, to_host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            e2e_step(W + i, to_host)
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e[name] = world * B * A * K / (float(tt.item()) * 1e-3)
    small_out = (h_state.numel() + h_coll.numel() + h_off.numel()) * 4

    # ---- the ceiling of the e2e legs: a bare device -> pinned host copy of one step's float32 images, all ranks at
    # once (PCIe per GPU; the host's root complexes and memory when 8 GPUs copy together)
    pcie = None
    if not args.kernels_only:
        for _ in range(2):
            h_img.copy_(images, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            h_img.copy_(images, non_blocking=True)
        c1.record()
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1) / 3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        pcie_ms = float(tc.item())
        pcie = {"d2h_gbs_per_gpu": h_img.numel() * 4 / pcie_ms / 1e6, "d2h_gbs_aggregate": world * h_img.numel() * 4 / pcie_ms / 1e6,
                "ms_per_step_of_images": pcie_ms,
                "e2e_ceiling": world * B * A / (pcie_ms * 1e-3),
                "note": "bare pinned D2H copy of one step's float32 images (3.2 GB per GPU), all ranks concurrently, max over ranks; "
                        "e2e_ceiling = agent-env-steps/s if a step cost nothing but that copy"}

    # ---- BASELINE configs 3-5 at the shard one GPU of the 8-GPU job holds (rank 0 only, after the headline loop)
    configs = None
    if rank == 0 and not args.kernels_only and not os.environ.get("TDS_BENCH_NO_CONFIGS"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "profiles", "bench_configs.py"))
        bc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bc)
        del images, h_img, h_img_u8, h_img_rank
        runner = None
        torch.cuda.empty_cache()
        try:
            configs = bc.run(("3", "4", "5"), dev)
        except Exception as e:      # noqa: BLE001
            configs = {"error": repr(e)[:300]}
    if world > 1:
        dist.barrier()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = B * A * 12 * RES * RES / (raster_ms * 1e-3) / 1e9
        # DRAM bytes of one raster launch of this size, from the committed `ncu --set full` capture (never measured
        # under the timer): profiles/r2_raster_ncu_summary.json
        traffic, traffic_src = None, None
        for ncu_name in ("r2_raster_ncu_summary.json", "r1_raster_ncu_summary.json"):
            ncu_path = os.path.join(ROOT, "profiles", ncu_name)
            if os.path.exists(ncu_path):
                nj = json.load(open(ncu_path))
                if nj.get("algorithmic_bytes_per_launch") == B * A * 12 * RES * RES:
                    traffic, traffic_src = nj["traffic_bytes_per_launch"], f"profiles/{ncu_name} (dram__bytes_read+write.sum)"
                    if nj.get("passes"):
                        traffic_src += "; both raster passes, incl. the bitplanes and face lists handed from the draw to the finish pass"
                    break
        cpu_val, cpu_procs = float("nan"), 0
        if not args.kernels_only:
            cpu_val, cpu_procs, _ = cpu_port_throughput(max(8, min(os.cpu_count() or 1, 128)))
        line = {
            "metric": METRIC, "value": value, "unit": "agent-env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": e2e["host_images"], "unit": "agent-env-steps/s", "h2d_bytes_per_step": int(h_act[0].numel() * 4),
                    "d2h_bytes_per_step": int(B * A * 3 * RES * RES * 4 + small_out),
                    "note": "public API, pinned host buffers; images delivered to host every step (PCIe-bound)"},
            "e2e_uint8": {"value": e2e["host_images_uint8"], "unit": "agent-env-steps/s", "h2d_bytes_per_step": int(h_act[0].numel() * 4),
                          "d2h_bytes_per_step": int(B * A * 3 * RES * RES + small_out),
                          "note": "same, images delivered as uint8 RGB (the pixel values are integers in [0,255]: identical after a cast)"},
            "e2e_rank": {"value": e2e["host_images_rank"], "unit": "agent-env-steps/s", "h2d_bytes_per_step": int(h_act[0].numel() * 4),
                         "d2h_bytes_per_step": int(B * A * RES * RES + small_out),
                         "note": "same, images delivered as one byte per pixel (draw rank; colour table from tds_raster_rank_table)"},
            "pcie": pcie, "configs": configs,
            "e2e_device_images": {"value": e2e["device_images"], "unit": "agent-env-steps/s",
                                  "h2d_bytes_per_step": int(h_act[0].numel() * 4), "d2h_bytes_per_step": int(small_out),
                                  "note": "same, images stay in HBM for a GPU consumer (the reference API returns device tensors)"},
            # per step, all ours (csrc/): kin_fwd, agent_boxes (cameras), dyn_prep, raster draw pass, raster finish pass, raster
            # (general kernel, on the draw pass's redo list: normally empty), agent_boxes (boxes), allpairs_fwd, offroad_fwd;
            # no library kernel (profiles/r2_launches_ncu.csv)
            "gpu_launches": 9 * K, "eager_ms_per_step": eager_ms,
            "roofline": {"bound": "hbm", "kernel": "raster (64x64 two-pass form: raster_kernel<draw> + raster_finish_kernel, timed together)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "raster_ms_per_launch": raster_ms, "algorithmic_bytes_per_launch": B * A * 12 * RES * RES,
                         "step_frac_of_hbm_roofline": value / world * BYTES_PER_AGENT_STEP / 1e9 / peak},
            "cpu_baseline": {"value": cpu_val, "unit": "agent-env-steps/s", "cores": cpu_procs, "kind": "port",
                             "sample": f"{max(8, min(os.cpu_count() or 1, 128))} environments x {AGENTS} agents, one step, "
                                       f"oracle (C + numpy port of the reference path), one process per core",
                             "reference_python": None if args.kernels_only else reference_python_throughput()},
            "infraction_metrics": {"collision_sum": float(metrics[0]), "offroad_sum": float(metrics[1]),
                                   "colliding_agent_steps": float(metrics[2]), "offroad_agent_steps": float(metrics[3])},
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _emit(text: str) -> None:
    """The result line goes to the process's original stdout (see main)."""
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels-only", action="store_true",
                    help="profiling aid: only the device-resident loop (no e2e legs, no CPU baseline)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything else written to file descriptor 1 by this process or
    # its libraries (NCCL prints its version banner there) goes to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
