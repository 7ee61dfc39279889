"""Traffic-light violations on the GPU (tds_traffic_light_violation) against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_golden_reference_violations():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = util.golden("traffic")
    box = torch.as_tensor(g["agent_box"], device=dev)
    raw = tds.ops.traffic_light_violation(box, torch.as_tensor(g["tl_corners"], device=dev),
                                          torch.as_tensor(g["tl_state"], device=dev), int(g["red_index"]),
                                          float(g["rear_factor"]))
    assert raw.dtype == torch.bool and np.array_equal(raw.cpu().numpy(), g["violation_raw"])
    masked = tds.ops.traffic_light_violation(box, torch.as_tensor(g["tl_corners"], device=dev),
                                             torch.as_tensor(g["tl_state"], device=dev), int(g["red_index"]),
                                             float(g["rear_factor"]), present=torch.as_tensor(g["present"], device=dev))
    assert np.array_equal(masked.cpu().numpy(), g["violation"])


@pytest.mark.parametrize("B,A", [(3, 50), (1, 1), (2, 300)])
def test_vs_oracle_random_and_control_object(B, A):
    """Random agents around the stop lines of Town02, masked controls, through TrafficLightControl.compute_violation."""
    import torchdrivesim_b200 as tds
    from oracle import traffic
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B * 100 + A)
    m = util.load_map_np("carla_Town02")
    pos, corners, state = util.tl_tensors(m, B, rng)
    L = pos.shape[1]
    mask = rng.uniform(size=(B, L)) > 0.2
    pick = rng.integers(0, L, (B, A))
    xy = np.take_along_axis(pos[..., :2], pick[..., None], 1) + rng.normal(0, 2.5, (B, A, 2))
    box = np.concatenate([xy, np.tile(np.array(util.VEH[:2]), (B, A, 1)), rng.uniform(0, 2 * np.pi, (B, A, 1))], -1).astype(np.float32)
    tl = tds.TrafficLightControl(pos=torch.as_tensor(pos, device=dev), mask=torch.as_tensor(mask, device=dev))
    tl.set_state(torch.as_tensor(state, device=dev))
    got = tl.compute_violation(torch.as_tensor(box, device=dev)).cpu().numpy()
    ref = traffic.tl_violation(box, tl.corners.cpu().numpy(), state, 0, 0.1)
    # area > 0 is a hard threshold: allow (and count) flips of boxes that merely touch
    assert (got != ref).sum() <= max(1, int(0.002 * B * A)), f"{(got != ref).sum()} of {B * A} differ"
    if B * A > 50:
        assert ref.sum() > 0


def test_simulator_method_and_edge_cases():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    m = util.load_map_np("carla_Town01")
    B, A = 2, 16
    state, size, types, present = util.random_scene(m, B, A, rng, absent_p=0.3)
    pos, corners, tl_state = util.tl_tensors(m, B, rng)
    state[:, :8, :2] = pos[:, :8, :2]                         # eight agents on stop lines
    state[:, :8, 2] = pos[:, :8, 4]
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
    km.set_state(torch.tensor(state, device=dev))
    tl = tds.TrafficLightControl(pos=torch.as_tensor(pos, device=dev))
    tl.set_state(torch.zeros(B, pos.shape[1], dtype=torch.long, device=dev))        # all red
    sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.tensor(present, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls={"traffic_light": tl})
    v = sim.compute_traffic_lights_violations().cpu().numpy()
    assert v.shape == (B, A) and v.dtype == np.float32 and set(np.unique(v)) <= {0.0, 1.0}     # the reference's dtype
    v = v != 0
    assert not v[~present].any()
    # an agent centred on a stop line with its heading overlaps it with its rear part only if the line is wide enough:
    # compare with the oracle instead of assuming
    from oracle import traffic
    box = np.concatenate([state[..., :2], size, state[..., 2:3]], -1)
    ref = traffic.tl_violation(box, tl.corners.cpu().numpy(), np.zeros((B, pos.shape[1]), np.int64), 0, 0.1, present)
    assert np.array_equal(v, ref)
    tl.set_state(torch.full((B, pos.shape[1]), 2, dtype=torch.long, device=dev))    # all green
    assert not sim.compute_traffic_lights_violations().any()
    # no lights at all / empty batch
    out = tds.ops.traffic_light_violation(torch.zeros(2, 3, 5, device=dev), torch.zeros(2, 0, 4, 2, device=dev),
                                          torch.zeros(2, 0, dtype=torch.long, device=dev), 0)
    assert out.shape == (2, 3) and not out.any()
    with pytest.raises(tds._lib.TdsError):
        tds.ops.traffic_light_violation(torch.zeros(2, 3, 4, device=dev), torch.zeros(2, 1, 4, 2, device=dev),
                                        torch.zeros(2, 1, dtype=torch.long, device=dev), 0)


def test_unrolled_schedule_steps_the_lights_on_the_device():
    """TrafficLightController.unroll -> replay_states: Simulator.step advances the lights by a device-side gather and
    the red-light violations follow the schedule (traffic_lights.py + traffic_controls.py:127-136 of the reference)."""
    import os
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "light_schedule.npz"))
    ctrl = tds.TrafficLightController.from_json(os.path.join(os.path.dirname(__file__), "golden", "maps",
                                                             "carla_Town02_traffic_light_controller.json"))
    ctrl.set_to([(int(s), float(r)) for s, r in g["start"]])
    steps = 60
    replay = ctrl.unroll(g["ids"].tolist(), 0.1, steps)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town02"))
    B = 2
    pos = torch.tensor(town.traffic_light_poses(), device=dev)[None].expand(B, -1, -1).contiguous()
    L = pos.shape[1]
    assert L == replay.shape[0]
    tl = tds.TrafficLightControl(pos, replay_states=replay[None].expand(B, -1, -1).contiguous().to(dev))
    # one stationary agent per stop line, its rear 10 % on the line (the part compute_violation tests)
    ahead = 0.45 * util.VEH[0] * torch.cat([torch.cos(pos[..., 4:5]), torch.sin(pos[..., 4:5])], -1)
    state = torch.cat([pos[..., :2] + ahead, pos[..., 4:5], torch.zeros(B, L, 1, device=dev)], -1)
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, L), util.VEH[2], device=dev))
    km.set_state(state)
    size = torch.tensor(util.VEH[:2], device=dev).expand(B, L, 2).contiguous()
    sim = tds.Simulator(town, km, size, torch.ones(B, L, dtype=torch.bool, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls={"traffic_light": tl})
    red = tl.allowed_states.index("red")
    from oracle import traffic
    box = torch.cat([state[:1, :, :2], size[:1], state[:1, :, 2:3]], -1).cpu().numpy()
    corners = tl.corners[:1].cpu().numpy()
    seen = set()
    for t in range(1, steps):
        sim.step(torch.zeros(B, L, 2, device=dev))
        assert np.array_equal(tl.state[0].cpu().numpy(), g["lights"][t]), t
        viol = sim.compute_traffic_lights_violations()[0].cpu().numpy() != 0
        assert np.array_equal(viol, traffic.tl_violation(box, corners, g["lights"][t][None], red, 0.1)[0]), t
        on_red = g["lights"][t] == red
        assert viol[on_red].all()                 # the agent whose rear is on a red light's stop line violates it
        seen.update(np.unique(g["lights"][t]).tolist())
    assert seen == {0, 1, 2}
