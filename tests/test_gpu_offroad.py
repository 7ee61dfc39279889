"""Offroad kernel vs the brute-force oracle (all 30 750 faces of Town01) and the reference golden."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def test_golden_offroad():
    import torchdrivesim_b200 as tds
    g = util.golden("offroad")
    town = tds.StaticMap.from_npz(util.map_path(str(g["map"])))
    st, sz = torch.tensor(g["state"]).cuda(), torch.tensor(g["size"]).cuda()
    for thr, key in ((0.5, "offroad_thr05"), (0.0, "offroad_thr0")):
        out = tds.offroad_infraction_loss(st, sz, town, threshold=thr).cpu().numpy()
        np.testing.assert_allclose(out, g[key], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("mapname", ["carla_Town01", "carla_Town02", "carla_Town10HD"])
def test_vs_oracle_random(mapname):
    from oracle import offroad as OF
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(42)
    m = util.load_map_np(mapname)
    B, A = 2, 96
    state, size, types, present = util.random_scene(m, B, A, rng, spread=30.0, ped_every=4, absent_p=0.1)
    state[0, :6, :2] = np.array([[-80, -60], [600, 500], [200, -300], [197, 164], [0, 0], [1e4, 1e4]], np.float32)
    town = tds.StaticMap.from_npz(util.map_path(mapname))
    for thr in (0.5, 0.0):
        sim_out = tds.ops.offroad(torch.tensor(state).cuda(), torch.tensor(size).cuda(), tds.MapSet([town]), thr,
                                  torch.tensor(present).cuda()).cpu().numpy()
        ref = np.stack([OF.offroad_loss(state[b], size[b], m["verts"], m["faces"], thr, present[b]) for b in range(B)])
        # a corner whose d2 sits within float rounding of the hard threshold may flip; none expected here
        np.testing.assert_allclose(sim_out, ref, rtol=RTOL, atol=ATOL)
    assert (ref > 0).sum() > 10 and (ref == 0).sum() > 10


def test_backward_matches_finite_differences():
    from oracle import offroad as OF
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(1)
    m = util.load_map_np("carla_Town01")
    state, size, _, _ = util.random_scene(m, 1, 48, rng, spread=25.0)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    st = torch.tensor(state).cuda().requires_grad_(True)
    sz = torch.tensor(size).cuda().requires_grad_(True)
    out = tds.offroad_infraction_loss(st, sz, town, threshold=0.5)
    out.sum().backward()
    g = st.grad.cpu().numpy()[0]
    base = OF.offroad_loss(state[0], size[0], m["verts"], m["faces"], 0.5).astype(np.float64)
    checked = 0
    for k, eps in ((0, 2e-2), (1, 2e-2), (2, 2e-3)):
        sp, sm = state[0].copy(), state[0].copy()
        sp[:, k] += eps; sm[:, k] -= eps
        fp = OF.offroad_loss(sp, size[0], m["verts"], m["faces"], 0.5).astype(np.float64)
        fm = OF.offroad_loss(sm, size[0], m["verts"], m["faces"], 0.5).astype(np.float64)
        fd = (fp - fm) / (2 * eps)
        # compare only agents whose set of thresholded corners / nearest features is stable under the step
        ok = (base > 1.0) & (np.abs(fp + fm - 2 * base) < 0.02 * np.maximum(np.abs(fp - fm), 1e-3) + 5e-2)
        checked += int(ok.sum())
        # fp32 finite differences of values up to ~1e3 carry ~0.1 of rounding noise at these step sizes
        np.testing.assert_allclose(g[ok, k], fd[ok], rtol=5e-2, atol=0.25)
    assert checked > 10
    assert float(g[base == 0].__abs__().max()) == 0.0


def test_full_size_properties():
    """Config-2 size (1024 x 64): agents whose 4 corners lie well inside road triangles give exactly 0;
    translating every agent far from the map gives 4 d^2-like growth; present mask zeroes rows."""
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(9)
    m = util.load_map_np("carla_Town01")
    B, A = 1024, 64
    state, size, _, present = util.random_scene(m, B, A, rng, spread=40.0, absent_p=0.2)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    st, sz, pr = torch.tensor(state).cuda(), torch.tensor(size).cuda(), torch.tensor(present).cuda()
    out = tds.ops.offroad(st, sz, tds.MapSet([town]), 0.5, pr)
    assert float(out[~pr].abs().max()) == 0.0
    assert bool((out >= 0).all())
    # sampled rows against the brute-force oracle
    from oracle import offroad as OF
    for b in (0, 511, 1023):
        ref = OF.offroad_loss(state[b], size[b], m["verts"], m["faces"], 0.5, present[b])
        np.testing.assert_allclose(out[b].cpu().numpy(), ref, rtol=RTOL, atol=ATOL)
