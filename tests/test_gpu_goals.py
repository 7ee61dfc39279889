"""Waypoint goals on the GPU (tds_waypoint_step / tds_waypoint_gather) against the goldens of the unmodified
reference and the oracle."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_golden_reference_rollout_with_goals():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = util.golden("goals")
    t = lambda k: torch.as_tensor(g[k], device=dev)
    B, A = g["state0"].shape[:2]
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=t("lr"))
    km.set_state(t("state0"))
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    sim = tds.Simulator(town, km, t("size"), torch.ones(B, A, dtype=torch.bool, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True),
                        waypoint_goals=tds.WaypointGoal(t("waypoints"), t("mask")))
    for step in range(g["actions"].shape[0]):
        sim.step(torch.as_tensor(g["actions"][step], device=dev))
        assert np.array_equal(sim.get_waypoints_state().cpu().numpy(), g["goal_state"][step]), step
        assert np.array_equal(sim.waypoint_goals.mask.cpu().numpy(), g["goal_mask"][step]), step
        assert np.array_equal(sim.get_waypoints().cpu().numpy(), g["wp1"][step])
        assert np.array_equal(sim.get_waypoints_mask().cpu().numpy(), g["m1"][step])
        assert np.array_equal(sim.get_waypoints(count=2).cpu().numpy(), g["wp2"][step])
        assert np.array_equal(sim.get_waypoints_mask(count=2).cpu().numpy(), g["m2"][step])
    assert sim.get_waypoints_state().dtype == torch.int64 and tuple(sim.get_waypoints_state().shape) == (B, A, 1)
    np.testing.assert_allclose(sim.get_state().cpu().numpy(), g["agent_state"], rtol=1e-5, atol=1e-5)
    img = sim.render_egocentric(n_subsequent_waypoints=2).cpu().numpy()
    assert img.shape == g["image"].shape
    bad = int((img != g["image"]).any(2).sum())
    assert bad <= 0.001 * B * A * 64 * 64, f"{bad} mismatching pixels"
    goal = np.floor(np.array(sim.renderer.color_map["goal_waypoint"], np.float64))
    assert (g["image"].transpose(0, 1, 3, 4, 2) == goal).all(-1).any()       # discs are in the frame


@pytest.mark.parametrize("B,A,N,M", [(3, 17, 5, 4), (1, 1, 1, 1), (2, 300, 3, 2), (2, 4, 2, 0)])
def test_vs_oracle(B, A, N, M):
    import torchdrivesim_b200 as tds
    from oracle import goals
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B + 10 * A + N)
    wp = rng.normal(0, 6, (B, A, N, M, 2)).astype(np.float32)
    mask = rng.uniform(size=(B, A, N, M)) > 0.3
    goal = tds.WaypointGoal(torch.as_tensor(wp, device=dev), torch.as_tensor(mask, device=dev))
    state = np.zeros((B, A, 1), np.int64)
    for step in range(N + 3):
        xy = rng.normal(0, 4, (B, A, 4)).astype(np.float32)
        goal.step(torch.as_tensor(xy, device=dev), threshold=3.0)
        mask, state = goals.waypoint_step(xy[..., :2], wp, mask, state, 3.0)
        assert np.array_equal(goal.state.cpu().numpy(), state) and np.array_equal(goal.mask.cpu().numpy(), mask)
        for count in (1, 3):
            w, m = goals.gather(wp, mask, state, count)
            assert np.array_equal(goal.get_waypoints(count).cpu().numpy(), w)
            assert np.array_equal(goal.get_masks(count).cpu().numpy(), m)
    if M and N > 1 and A > 1:
        assert state.max() > 0
    sub = goal.select_batch_elements(torch.tensor([B - 1], device=dev), in_place=False).extend(2)
    assert sub.waypoints.shape[0] == 2 and torch.equal(sub.state[0], goal.state[B - 1])
    assert torch.equal(torch.as_tensor(wp, device=dev), goal.waypoints)          # never written
