"""Replayed NPCs with spawning / despawning on the GPU (tds_npc_advance) and NPCs in the render / collision /
observation paths, against the goldens of the unmodified reference and the oracle."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _sim_from_golden(g, dev):
    import torchdrivesim_b200 as tds
    t = lambda k, **kw: torch.as_tensor(g[k], device=dev, **kw)
    ctrl = tds.ReplayController(t("npc_size"), t("replay"), t("replay_present"), npc_types=t("npc_types"),
                                agent_type_names=["vehicle", "pedestrian"],
                                spawn_controller=tds.SpawnController(t("boundary"), t("spawn_states"), t("spawn_masks")))
    B, A = g["agent_state0"].shape[:2]
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=t("lr"))
    km.set_state(t("agent_state0"))
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    return tds.Simulator(town, km, t("agent_size"), torch.ones(B, A, dtype=torch.bool, device=dev),
                         tds.TorchDriveConfig(left_handed_coordinates=True), agent_types=torch.zeros(B, A, dtype=torch.long, device=dev),
                         agent_type_names=["vehicle", "pedestrian"], npc_controller=ctrl)


def test_golden_reference_rollout_with_npcs():
    dev = torch.device("cuda:0")
    g = util.golden("npc")
    sim = _sim_from_golden(g, dev)
    A, Np = sim.agent_count, sim.npc_count
    assert (A, Np) == (g["agent_state0"].shape[1], g["npc_size"].shape[1])
    for step in range(g["actions"].shape[0]):
        sim.step(torch.as_tensor(g["actions"][step], device=dev))
        assert np.array_equal(sim.get_npc_present_mask().cpu().numpy(), g["npc_present"][step]), step
        assert np.array_equal(sim.get_npc_state().cpu().numpy(), g["npc_state"][step]), step
        np.testing.assert_allclose(sim.compute_collision().cpu().numpy(), g["collision"][step], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(sim.get_all_agents_absolute().cpu().numpy(), g["absolute"][step], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sim.get_state().cpu().numpy(), g["agent_state"], rtol=1e-5, atol=1e-5)
    rel = sim.get_all_agents_relative().cpu().numpy()
    assert rel.shape == g["relative"].shape == (2, A, A + Np - 1, 6)
    dist = np.linalg.norm(g["relative"][..., :2], axis=-1, keepdims=True)
    assert (np.abs(rel[..., :2] - g["relative"][..., :2]) <= 1e-5 * dist + 1e-4).all()
    assert np.array_equal(rel[..., 3:], g["relative"][..., 3:])
    # the rendered frame shows agents and NPCs (vehicles and pedestrians); <= 0.1 % of the pixels may differ
    img = sim.render_egocentric().cpu().numpy()
    assert img.shape == g["image"].shape
    bad = int((img != g["image"]).any(2).sum())
    assert bad <= 0.001 * img.shape[0] * img.shape[1] * 64 * 64, f"{bad} mismatching pixels"
    ped = np.array([int(v) for v in np.floor(np.array(sim.renderer.color_map["pedestrian"]))])
    assert (g["image"].transpose(0, 1, 3, 4, 2) == ped).all(-1).any()


@pytest.mark.parametrize("B,Np,V,replay,spawn", [(3, 40, 5, True, True), (2, 7, 0, True, False), (1, 300, 4, False, True),
                                                  (2, 5, 3, False, False), (2, 0, 4, True, True)])
def test_vs_oracle(B, Np, V, replay, spawn):
    import torchdrivesim_b200 as tds
    from oracle import npc
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(B * 100 + Np)
    T, S = 5, 9
    rs = np.concatenate([rng.normal(0, 20, (B, Np, T, 2)), rng.uniform(-3, 3, (B, Np, T, 2))], -1).astype(np.float32)
    rp = rng.uniform(size=(B, Np, T)) > 0.3
    ang = np.sort(rng.uniform(0, 2 * np.pi, (B, max(V, 1))), -1)
    if rng.uniform() < 0.5:
        ang = ang[:, ::-1]                                   # either orientation
    poly = (25.0 * np.stack([np.cos(ang), np.sin(ang)], -1)).astype(np.float32) if V else None
    ss = np.concatenate([rng.normal(0, 10, (B, Np, S, 2)), rng.uniform(-3, 3, (B, Np, S, 2))], -1).astype(np.float32)
    sm = rng.uniform(size=(B, Np, S)) > 0.5
    t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), device=dev)
    sc = tds.SpawnController(t(poly), t(ss) if spawn else None, t(sm) if spawn else None)
    size = torch.ones(B, Np, 2, device=dev)
    if replay:
        ctrl = tds.ReplayController(size, t(rs), t(rp), spawn_controller=sc)
    else:
        ctrl = tds.NPCController(size, t(rs[:, :, 0]), t(rp[:, :, 0]), spawn_controller=sc)

    class _Sim:                                             # the controllers only touch simulator.npc_controller
        npc_controller = ctrl

    state, present = rs[:, :, 0], rp[:, :, 0]
    t_replay = 0
    log = ctrl.npc_states.clone() if replay else None
    for step in range(7):
        ctrl.advance_npcs(_Sim)
        t_replay = (t_replay + 1) % T
        state, present = npc.npc_advance(state, present, rs if replay else None, rp if replay else None, t_replay, poly,
                                         ss if spawn else None, sm if spawn else None, step)
        assert np.array_equal(ctrl.npc_present_mask.cpu().numpy(), present), step
        assert np.array_equal(ctrl.npc_state.cpu().numpy(), state), step
    if replay:
        assert torch.equal(log, ctrl.npc_states)            # the replay log is never written
        assert ctrl.time == 7 % T
    if spawn:
        with pytest.raises(IndexError):                     # the reference indexes past the spawn table
            for _ in range(S):
                ctrl.advance_npcs(_Sim)


def test_batch_plumbing():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = util.golden("npc")
    sim = _sim_from_golden(g, dev)
    one = sim.select_batch_elements(torch.tensor([1], device=dev), in_place=False)
    assert one.npc_count == sim.npc_count and one.get_npc_state().shape[0] == 1
    for step in range(2):
        one.step(torch.as_tensor(g["actions"][step][1:2], device=dev))
    assert np.array_equal(one.get_npc_state().cpu().numpy(), g["npc_state"][1][1:2])
    assert sim.npc_controller.time == 0                    # the copy stepped, the original did not
    big = sim.npc_controller.copy().extend(3)
    assert big.npc_states.shape[0] == 6 and big.spawn_controller.spawn_masks.shape[0] == 6


def test_simulator_extend_repeats_environments():
    """Simulator.extend(n) (simulator.py:443-478): every environment n times in a row, all components included."""
    dev = torch.device("cuda:0")
    g = util.golden("npc")
    sim = _sim_from_golden(g, dev)
    twice = sim.extend(2, in_place=False)
    assert twice.batch_size == 2 * sim.batch_size and sim.batch_size == g["agent_state0"].shape[0]
    for step in range(3):
        act = torch.as_tensor(g["actions"][step], device=dev)
        sim.step(act)
        twice.step(act.repeat_interleave(2, dim=0))
    for got, ref in ((twice.get_state(), sim.get_state()), (twice.get_npc_state(), sim.get_npc_state()),
                     (twice.get_npc_present_mask(), sim.get_npc_present_mask()), (twice.compute_collision(), sim.compute_collision()),
                     (twice.compute_offroad(), sim.compute_offroad()), (twice.render_egocentric(), sim.render_egocentric())):
        assert torch.equal(got, ref.repeat_interleave(2, dim=0))
