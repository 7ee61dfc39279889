"""CUDA-graph replay of the hot path equals the eager path, step after step (incl. traffic-light replay)."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _make(B, A, seed, lights, npcs=0, goals=False):
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    m = util.load_map_np("carla_Town02")
    state, size, types, present = util.random_scene(m, B, A, rng, absent_p=0.1)
    town = tds.StaticMap.from_npz(util.map_path("carla_Town02"))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
    km.set_state(torch.tensor(state, device=dev))
    tc = None
    if lights:
        pos = torch.tensor(town.traffic_light_poses(), device=dev)[None].expand(B, -1, -1).contiguous()
        replay = torch.tensor(rng.integers(0, 3, (B, pos.shape[1], 6)), device=dev)
        tc = {"traffic_light": tds.TrafficLightControl(pos, replay_states=replay)}
    ctrl = None
    if npcs:
        log = np.concatenate([state[:, :1, None, :2] + rng.normal(0, 8, (B, npcs, 4, 2)), rng.uniform(0, 5, (B, npcs, 4, 2))], -1)
        spawn = tds.SpawnController(None, torch.tensor(log[:, :, ::-1].copy(), dtype=torch.float32, device=dev).repeat(1, 1, 2, 1),
                                    torch.tensor(rng.uniform(size=(B, npcs, 8)) > 0.5, device=dev))
        ctrl = tds.ReplayController(torch.full((B, npcs, 2), 2.0, device=dev), torch.tensor(log, dtype=torch.float32, device=dev),
                                    torch.tensor(rng.uniform(size=(B, npcs, 4)) > 0.4, device=dev), spawn_controller=spawn)
    wg = None
    if goals:
        # three collections of two waypoints per agent, the first ones within reach of the first steps
        wp = state[:, :, None, None, :2] + rng.normal(0, 1.5, (B, A, 3, 2, 2)) + np.arange(3)[None, None, :, None, None] * 2.0
        wg = tds.WaypointGoal(torch.tensor(wp, dtype=torch.float32, device=dev), torch.tensor(rng.uniform(size=(B, A, 3, 2)) > 0.2, device=dev))
    sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.tensor(present, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls=tc, npc_controller=ctrl,
                        waypoint_goals=wg)
    acts = torch.tensor(rng.uniform(-1, 1, (5, B, A, 2)).astype(np.float32), device=dev)
    return sim, acts


@pytest.mark.parametrize("lights,npcs,goals", [(False, 0, False), (True, 0, False), (True, 3, False), (True, 2, True)])
def test_graph_replay_equals_eager(lights, npcs, goals):
    import torchdrivesim_b200 as tds
    eager, acts = _make(6, 5, 3, lights, npcs, goals)
    graphed_sim, _ = _make(6, 5, 3, lights, npcs, goals)
    runner = tds.GraphedHotPath(graphed_sim)
    for t in range(acts.shape[0]):
        eager.step(acts[t])
        img_e, coll_e, off_e = eager.render_egocentric(), eager.compute_collision(), eager.compute_offroad()
        img_g, coll_g, off_g = runner.run(acts[t])
        torch.cuda.synchronize()
        assert torch.equal(runner.state, eager.get_state()), f"state differs at step {t}"
        assert torch.equal(img_g, img_e), f"image differs at step {t}"
        assert torch.equal(coll_g, coll_e) and torch.equal(off_g, off_e)
        if goals:       # the goals advance inside the graph (simulator.py:860-861) and the discs follow
            assert torch.equal(graphed_sim.waypoint_goals.state, eager.waypoint_goals.state), f"goal state differs at step {t}"
            assert torch.equal(graphed_sim.waypoint_goals.mask, eager.waypoint_goals.mask)
        if npcs:
            assert torch.equal(graphed_sim.get_npc_state(), eager.get_npc_state())
            assert torch.equal(graphed_sim.get_npc_present_mask(), eager.get_npc_present_mask())
    assert graphed_sim.internal_time == eager.internal_time == acts.shape[0]
    if goals:
        assert int(eager.waypoint_goals.state.max()) > 0 and (img_e[:, :, 0] == 139).any()      # goals were reached; discs are drawn
