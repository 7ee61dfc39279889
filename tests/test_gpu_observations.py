"""Non-visual observations on the GPU (tds_agents_relative) against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _close(got, ref):
    # positions: 1e-5 relative to the distance between the two agents (sin/cos of the origin's heading differ by an ulp
    # between the float64-rounded kernel and numpy's float32 libm) + 2e-5 m; angles may sit on either side of the wrap
    dist = np.linalg.norm(ref[..., :2], axis=-1, keepdims=True)
    assert (np.abs(got[..., :2] - ref[..., :2]) <= 1e-5 * dist + 2e-5).all()
    d = np.abs(got[..., 2] - ref[..., 2])
    assert np.minimum(d, np.abs(d - 2 * np.pi)).max() < 2e-5
    assert np.array_equal(got[..., 3:], ref[..., 3:])


def test_golden_reference_relative():
    import torchdrivesim_b200 as tds
    g = util.golden("relative")
    a = torch.as_tensor(g["absolute"], device="cuda:0")
    for key, excl in (("relative_excl", True), ("relative_all", False)):
        got = tds.ops.agents_relative(a, exclude_self=excl).cpu().numpy()
        assert got.shape == g[key].shape
        _close(got, g[key])


@pytest.mark.parametrize("B,N,A,excl", [(4, 64, 64, True), (2, 33, 7, True), (1, 1, 1, True), (3, 5, 5, False), (2, 300, 300, True)])
def test_vs_oracle(B, N, A, excl):
    import torchdrivesim_b200 as tds
    from oracle import observations
    rng = np.random.default_rng(B * 1000 + N)
    a = np.concatenate([rng.uniform(-300, 300, (B, N, 2)), rng.uniform(-7, 7, (B, N, 1)), rng.uniform(0.5, 5, (B, N, 2)),
                        (rng.uniform(size=(B, N, 1)) > 0.3)], -1).astype(np.float32)
    got = tds.ops.agents_relative(torch.as_tensor(a, device="cuda:0"), A, excl).cpu().numpy()
    ref = observations.agents_relative(a, A, excl)
    assert got.shape == ref.shape
    if got.size:
        _close(got, ref)


def test_simulator_methods():
    import torchdrivesim_b200 as tds
    from oracle import observations
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    m = util.load_map_np("carla_Town01")
    B, A = 2, 12
    state, size, types, present = util.random_scene(m, B, A, rng, absent_p=0.3)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
    km.set_state(torch.tensor(state, device=dev))
    sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.tensor(present, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True))
    absolute = sim.get_all_agents_absolute()
    assert tuple(absolute.shape) == (B, A, 6)
    assert np.array_equal(absolute[..., 5].cpu().numpy() != 0, present)
    rel = sim.get_all_agents_relative()
    assert tuple(rel.shape) == (B, A, A - 1, 6)
    _close(rel.cpu().numpy(), observations.agents_relative(absolute.cpu().numpy()))
    assert tuple(sim.get_all_agents_relative(exclude_self=False).shape) == (B, A, A, 6)
