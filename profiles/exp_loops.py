import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torchdrivesim_b200 as tds
dev = torch.device("cuda:0")
B, A, K = 1024, 64, 20
state, size, lr, actions = bench.synth_inputs(B, A, 1000, K + 5)
town = tds.StaticMap.from_npz(bench.map_npz())
def make():
    km = tds.KinematicBicycle(left_handed=True); km.set_params(lr=torch.tensor(lr, device=dev)); km.set_state(torch.tensor(state, device=dev))
    return tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev), tds.TorchDriveConfig(left_handed_coordinates=True))
act = torch.tensor(actions, device=dev)
images = torch.empty(B, A, 3, 64, 64, device=dev)
sim = make()
def eager(i):
    sim.step(act[i]); sim.render_egocentric(out=images); c = sim.compute_collision(); o = sim.compute_offroad(); return c, o
def timeit(fn, label, n=K):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(i)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{label:40s} gpu {e0.elapsed_time(e1)/n:7.3f} ms/step   cpu-launch {1e3*(t1-t0)/n:7.3f} ms/step")
timeit(eager, "eager")
metrics = torch.zeros(4, dtype=torch.float64, device=dev)
def eager_m(i):
    global metrics
    c, o = eager(i); metrics += torch.stack([c.sum(), o.sum(), (c > 0).sum(), (o > 0).sum()]).double()
timeit(eager_m, "eager + metrics")
sim2 = make(); runner = tds.GraphedHotPath(sim2)
timeit(lambda i: runner.run(act[i]), "graph")
def graph_m(i):
    global metrics
    _, c, o = runner.run(act[i]); metrics += torch.stack([c.sum(), o.sum(), (c > 0).sum(), (o > 0).sum()]).double()
timeit(graph_m, "graph + metrics")
s = bench.ClockSampler(0); s.start(); time.sleep(0.3)
timeit(lambda i: runner.run(act[i]), "graph + NVML sampler 100 ms")
timeit(eager, "eager + NVML sampler 100 ms")
print(s.stop(0, 1e18))
