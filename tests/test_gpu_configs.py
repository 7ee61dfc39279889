"""BASELINE configs 3 and 4 at the SHAPE of the shard one GPU holds (fewer environments, the same agents per
environment, maps, tile sizes and metrics), against the oracle on sampled cameras / environments:

  config 3: Town01 / Town02 / Town10HD by env_map, 128 agents (96 bicycle vehicles + 32 unicycle pedestrians), traffic
            lights, 128x128 birdviews, discs collisions, offroad, red-light violations
  config 4: 512 agents on the roads of a 120 m box, 256x256 birdviews (every camera sees 100-200 agents: the branch that
            walks ALL dynamic primitives), IoU all-pairs collisions at N = 512
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _configs():
    spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "profiles", "bench_configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _sincos(psi):
    p = torch.as_tensor(psi)
    return torch.stack([torch.sin(p), torch.cos(p)], -1).numpy()


def test_config3_shape_against_the_oracle():
    import torchdrivesim_b200 as tds
    from oracle import collision as OC, offroad as OO, traffic
    bc = _configs()
    dev = torch.device("cuda:0")
    B, A, res = 12, 128, 128
    sim, act, _ = bc.build_config3(dev, B=B, A=A)
    names = ["carla_Town01", "carla_Town02", "carla_Town10HD"]
    maps = [util.load_map_np(n) for n in names]
    for _ in range(2):
        sim.step(act)
    img = sim.render_egocentric(res=tds.Resolution(res, res))
    coll, off, viol = sim.compute_collision(), sim.compute_offroad(), sim.compute_traffic_lights_violations()
    torch.cuda.synchronize()
    st = sim.get_state().cpu().numpy()
    size = sim.get_agent_size().cpu().numpy()
    types = sim.get_agent_type().cpu().numpy()
    present = sim.get_present_mask().cpu().numpy()
    tl = sim.traffic_controls["traffic_light"]
    corners, tl_state, tl_mask = tl.corners.cpu().numpy(), tl.state.cpu().numpy(), tl.mask.cpu().numpy()
    assert len(np.unique(tl_state)) == 3 and (types == 1).sum() == B * 32
    rng = np.random.default_rng(33)
    cams = sorted({(int(b), int(c)) for b, c in zip(rng.integers(0, B, 40), rng.integers(0, A, 40))})[:32]
    cam_sc = _sincos(st[..., 2])
    bad = 0
    for (b, c) in cams:
        m = maps[b % 3]
        L = int(tl_mask[b].sum())                       # the lights of this environment's map (the rest is padding)
        ora = util.oracle_render_batch(m, st[b:b + 1], size[b:b + 1], types[b:b + 1], present[b:b + 1], ["vehicle", "pedestrian"],
                                       corners[b:b + 1, :L], tl_state[b:b + 1, :L], st[b:b + 1, :, :2].copy(), cam_sc[b:b + 1], res, 35.0,
                                       cams=[(0, c)])
        bad += int((img[b, c].cpu().numpy() != ora[(0, c)]).any(0).sum())
    assert bad == 0, f"{bad} mismatching pixels over {len(cams)} cameras"
    for b in (0, 1, 2, 7):
        m = maps[b % 3]
        box = torch.cat([torch.tensor(st[b:b + 1, :, :2]), torch.tensor(size[b:b + 1]), torch.tensor(st[b:b + 1, :, 2:3])], -1)
        np.testing.assert_allclose(coll[b:b + 1].cpu().numpy(), OC.collision_allpairs(box, box, torch.tensor(present[b:b + 1])).numpy(),
                                   rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(off[b].cpu().numpy(), OO.offroad_loss(st[b], size[b], m["verts"], m["faces"], 0.5, present[b]),
                                   rtol=1e-5, atol=1e-5)
        L = int(tl_mask[b].sum())
        ref = traffic.tl_violation(box.numpy(), corners[b:b + 1, :L], tl_state[b:b + 1, :L], tl.allowed_states.index("red"), 0.1,
                                   present[b:b + 1])
        assert np.array_equal(viol[b:b + 1].cpu().numpy() != 0, ref)


def test_config4_shape_against_the_oracle():
    import torchdrivesim_b200 as tds
    from oracle import iou as I
    bc = _configs()
    dev = torch.device("cuda:0")
    B, A, res = 4, 512, 256
    sim, act, _ = bc.build_config4(dev, B=B, A=A)
    sim.step(act)
    img = sim.render_egocentric(res=tds.Resolution(res, res))
    coll = sim.compute_collision()
    torch.cuda.synchronize()
    st = sim.get_state().cpu().numpy()
    size = sim.get_agent_size().cpu().numpy()
    m = util.load_map_np("carla_Town01")
    types, present = np.zeros((B, A), np.int64), np.ones((B, A), bool)
    # the cameras that see the most agents
    d = np.abs(st[:, :, None, :2] - st[:, None, :, :2]).max(-1)
    seen = (d < 17.0).sum(-1)
    cams = [(b, int(c)) for b in range(B) for c in np.argsort(-seen[b])[:4]]
    assert min(seen[b, c] for b, c in cams) > 100
    cam_sc = _sincos(st[..., 2])
    ora = util.oracle_render_batch(m, st, size, types, present, ["vehicle"], None, None, st[..., :2].copy(), cam_sc, res, 35.0, cams=cams)
    bad = sum(int((img[b, c].cpu().numpy() != o).any(0).sum()) for (b, c), o in ora.items())
    assert bad == 0, f"{bad} mismatching pixels over {len(ora)} cameras"
    # IoU all-pairs at N = 512 against the float64 restatement of the reference's vertex-sort algorithm, diagonal := 1
    for b in range(B):
        box = np.concatenate([st[b, :, :2], size[b], st[b, :, 2:3]], -1).astype(np.float64)
        ref = I.iou_matrix(box, box)
        np.fill_diagonal(ref, 1.0)
        agg = ref.sum(-1) - ref.max(-1)
        np.testing.assert_allclose(coll[b].cpu().numpy(), agg, rtol=1e-5, atol=A * 2e-6)
    assert float(coll.max()) > 0.1
