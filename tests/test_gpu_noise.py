"""Noisy observations on the GPU (tds_sensing_noise, tds_sensing_occlusion, per-origin tds_agents_relative) against
the goldens of the unmodified reference and the oracle."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _rel_close(got, ref):
    dist = np.linalg.norm(ref[..., :2], axis=-1, keepdims=True)
    assert (np.abs(got[..., :2] - ref[..., :2]) <= 1e-5 * dist + 1e-4).all()
    d = np.abs(got[..., 2] - ref[..., 2])
    assert np.minimum(d, np.abs(d - 2 * np.pi)).max() < 2e-5
    assert np.array_equal(got[..., 3:], ref[..., 3:])


def test_golden_reference_noisy_observations():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = util.golden("noise")
    t = lambda k: torch.as_tensor(g[k], device=dev)
    B, A = g["agent_state"].shape[:2]
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
    km.set_state(t("agent_state"))

    class FixedNoise(tds.StandardSensingObservationNoise):       # the reference's deviates, drawn from the same seed
        def sample_noise(self, shape, device):
            assert tuple(shape) == g["eps"].shape
            return t("eps")

    sim = tds.Simulator(tds.StaticMap.from_npz(util.map_path("carla_Town01")), km, t("agent_size"), t("present"),
                        tds.TorchDriveConfig(left_handed_coordinates=True),
                        npc_controller=tds.NPCController(t("npc_size"), t("npc_state"), t("npc_present")),
                        observation_noise_model=FixedNoise())
    np.testing.assert_allclose(sim.get_noisy_state().cpu().numpy(), g["noisy_state"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(sim.get_noisy_present_mask().cpu().numpy(), g["noisy_present"])
    assert np.array_equal(sim.get_noisy_agent_size().cpu().numpy(), g["noisy_size"])
    np.testing.assert_allclose(sim.get_noisy_all_agents_absolute().cpu().numpy(), g["noisy_absolute"], rtol=1e-6, atol=1e-6)
    _rel_close(sim.get_noisy_all_agents_relative().cpu().numpy(), g["noisy_relative"])
    _rel_close(sim.get_noisy_all_agents_relative(exclude_self=False).cpu().numpy(), g["noisy_relative_all"])
    # the base model perceives the truth
    sim.observation_noise_model = tds.ObservationNoise()
    truth = sim.get_all_agents_absolute()
    assert torch.equal(sim.get_noisy_all_agents_absolute(), truth[:, None].expand(-1, A, -1, -1))
    _rel_close(sim.get_noisy_all_agents_relative().cpu().numpy(), sim.get_all_agents_relative().cpu().numpy())


@pytest.mark.parametrize("B,A,N", [(3, 20, 33), (1, 1, 1), (2, 64, 64), (1, 5, 300)])
def test_vs_oracle(B, A, N):
    import torchdrivesim_b200 as tds
    from oracle import observations
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B * 1000 + N)
    state = np.concatenate([rng.normal(0, 45, (B, N, 2)), rng.uniform(-3, 3, (B, N, 2))], -1).astype(np.float32)
    size = rng.uniform(0.5, 5, (B, N, 2)).astype(np.float32)
    base = rng.uniform(size=(B, N)) > 0.2
    eps = rng.normal(size=(B, A, N, 4)).astype(np.float32)
    got = tds.ops.sensing_noise(torch.as_tensor(state, device=dev), A, torch.as_tensor(eps, device=dev)).cpu().numpy()
    np.testing.assert_allclose(got, observations.noisy_state(state, A, eps), rtol=1e-6, atol=1e-6)
    mask = tds.ops.sensing_occlusion(torch.as_tensor(state, device=dev), torch.as_tensor(size, device=dev),
                                     torch.as_tensor(base, device=dev), A).cpu().numpy()
    ref = observations.noisy_present_mask(state, size, base, A)
    # grazing lines sit on a hard threshold: allow (and count) a few flips
    assert (mask != ref).sum() <= max(1, int(2e-4 * mask.size)), f"{(mask != ref).sum()} of {mask.size} differ"
    assert not mask[~np.broadcast_to(base[:, None], mask.shape)].any()
    # random deviates: mean zero, the deviation grows with the distance
    noisy = tds.ops.sensing_noise(torch.as_tensor(state, device=dev), A).cpu().numpy()
    assert noisy.shape == (B, A, N, 4) and np.isfinite(noisy).all()
    if B * A * N > 2000:
        err = np.abs(noisy - state[:, None])[..., 0]
        dist = np.linalg.norm(state[:, :A, None, :2] - state[:, None, :, :2], axis=-1)
        assert err[dist > 100].mean() > err[(dist > 0.5) & (dist < 25)].mean() * 5
