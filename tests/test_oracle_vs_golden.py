"""Pins the oracle (CPU restatement) to outputs of the UNMODIFIED reference (tests/golden/*.npz, generated
by tests/golden/make_golden.py in the build container).  The reference's own tests hold no golden vectors
for this path, so these are what parity is anchored on."""
import numpy as np
import torch

from oracle import collision as C, iou as I, kinematic as K, offroad as OF, raster as R
from tests import util


def test_kinematic_trajectories_bit_exact():
    g = util.golden("kinematic")
    for case in g["cases"]:
        q = lambda k: g[f"{case}/{k}"]
        st, lr, acts = torch.tensor(q("state0")), torch.tensor(q("lr")), torch.tensor(q("actions"))
        for t in range(acts.shape[0]):
            st = K.bicycle_step(st, acts[t], lr, 0.1, bool(q("left_handed")), no_reversing=case.startswith("noreverse"))
            assert np.array_equal(st.numpy(), q("traj")[t + 1]), f"{case} step {t}"


def test_compound_matches_single_models():
    gen = torch.Generator().manual_seed(0)
    st = torch.rand(3, 5, 4, generator=gen) * 10
    act = torch.rand(3, 5, 2, generator=gen) * 2 - 1
    lr = 1 + torch.rand(3, 5, generator=gen)
    for mid, fn in ((0, lambda: K.bicycle_step(st, act, lr)), (1, lambda: K.bicycle_step(st, act, lr, no_reversing=True)),
                    (2, lambda: K.unicycle_step(st, act))):
        out = K.compound_step(st, act, lr, torch.full((3, 5), mid))
        assert torch.equal(out, fn())


def test_discs_pairwise_aggregate_and_gradient():
    g = util.golden("collision")
    box, present = torch.tensor(g["box"]), torch.tensor(g["present"])
    assert np.allclose(C.overlap_matrix(box, box).numpy(), g["discs_pair"], rtol=0, atol=1e-7)
    assert np.allclose(C.collision_allpairs(box, box, present).numpy(), g["discs_collision"], rtol=1e-6, atol=1e-6)
    B, A = box.shape[:2]
    st = torch.cat([box[..., :2], box[..., 4:5], torch.zeros(B, A, 1)], -1).requires_grad_(True)
    bx = torch.cat([st[..., :2], box[..., 2:4], st[..., 2:3]], -1)
    C.collision_allpairs(bx, bx, present).sum().backward()
    assert np.allclose(st.grad.numpy(), g["discs_grad_state"], rtol=1e-5, atol=1e-6)


def test_iou_float64_matches_reference_float64():
    g = util.golden("collision")
    iou = np.stack([I.iou_matrix(g["box"][b], g["box"][b]) for b in range(g["box"].shape[0])])
    assert np.abs(iou - g["iou64"]).max() < 1e-12
    assert np.abs(np.diagonal(iou, axis1=1, axis2=2) - 1).max() < 1e-9
    # the reference's own fp32 evaluation is NOT a usable oracle (App. C-8): document the discrepancy
    assert np.abs(g["iou32"] - g["iou64"]).max() > 0.1


def test_offroad_brute_force_bit_exact():
    g = util.golden("offroad")
    m = util.load_map_np(str(g["map"]))
    for thr, key in ((0.5, "offroad_thr05"), (0.0, "offroad_thr0")):
        out = OF.offroad_loss(g["state"][0], g["size"][0], m["verts"], m["faces"], thr)
        assert np.array_equal(out, g[key][0])


def test_render_pixel_exact():
    g = util.golden("render")
    total = 0
    for case in g["cases"]:
        q = lambda k: g[f"{case}/{k}"]
        m = util.load_map_np(str(q("map")))
        st = q("state")
        cam_sc = torch.stack([torch.sin(torch.tensor(st[..., 2])), torch.cos(torch.tensor(st[..., 2]))], -1).numpy()
        ora = util.oracle_render_batch(m, st, q("size"), q("types"), q("present"), [str(s) for s in q("type_names")],
                                       q("tl_corners"), q("tl_state"), st[..., :2], cam_sc, int(q("res")), float(q("fov")))
        for (b, c), img in ora.items():
            assert np.array_equal(img, q("image")[b, c].astype(np.float32)), f"{case} camera {(b, c)}"
            total += 1
    assert total == 12 + 5 + 6 + 2


def test_render_waypoints_pixel_exact():
    """Goal-waypoint discs per camera with a rendering mask (mesh.py:1120-1145) against the unmodified reference."""
    g = util.golden("render_waypoints")
    m = util.load_map_np(str(g["map"]))
    st = g["state"]
    B, A = st.shape[:2]
    cam_sc = torch.stack([torch.sin(torch.tensor(st[..., 2])), torch.cos(torch.tensor(st[..., 2]))], -1).numpy()
    ora = util.oracle_render_batch(m, st, g["size"], np.zeros((B, A), np.int64), np.ones((B, A), bool), ["vehicle"],
                                   g["tl_corners"], g["tl_state"], st[..., :2], cam_sc, int(g["res"]), float(g["fov"]),
                                   waypoints=g["waypoints"], waypoints_mask=g["waypoints_mask"])
    for (b, c), img in ora.items():
        assert np.array_equal(img, g["image"][b, c].astype(np.float32)), f"camera {(b, c)}"
    assert len(ora) == B * A


def test_render_custom_colours_and_static_meshes_pixel_exact():
    """generate(custom_agent_colors=...) (mesh.py:1092-1099) + add_static_meshes (mesh.py:870-883) against the
    unmodified reference: per-camera agent colours, the degenerate face of an absent agent 0, extra static faces."""
    g = util.golden("render_custom")
    m0 = util.load_map_np(str(g["map"]))
    st = g["state"]
    B, A = st.shape[:2]
    cam_sc = torch.stack([torch.sin(torch.tensor(st[..., 2])), torch.cos(torch.tensor(st[..., 2]))], -1).numpy()
    names = [str(s) for s in g["type_names"]]
    n = 0
    for b in range(B):
        m = util.with_extra_static(m0, g["extra_verts"][b], g["extra_faces"][b], "map_boundary")
        ora = util.oracle_render_batch(m, st, g["size"], g["types"], g["present"], names, g["tl_corners"], g["tl_state"],
                                       st[..., :2], cam_sc, int(g["res"]), float(g["fov"]), cams=[(b, c) for c in range(A)],
                                       agent_colors=g["colors"])
        for (bb, c), img in ora.items():
            assert np.array_equal(img, g["image"][bb, c].astype(np.float32)), f"camera {(bb, c)}"
            n += 1
    assert n == B * A


def test_category_ranks_follow_levels():
    r = R.category_ranks()
    assert r["road"] < r["right_lane"] < r["left_lane"] < r["traffic_light_green"] < r["traffic_light_red"] \
        < r["pedestrian"] < r["vehicle"] < r["direction"]


def test_traffic_light_violation_oracle_matches_reference():
    """oracle/traffic.py against TrafficLightControl.compute_violation of the unmodified reference."""
    from oracle import traffic
    g = util.golden("traffic")
    raw = traffic.tl_violation(g["agent_box"], g["tl_corners"], g["tl_state"], int(g["red_index"]), float(g["rear_factor"]))
    assert raw.sum() > 20 and np.array_equal(raw, g["violation_raw"])
    masked = traffic.tl_violation(g["agent_box"], g["tl_corners"], g["tl_state"], int(g["red_index"]),
                                  float(g["rear_factor"]), g["present"])
    assert np.array_equal(masked, g["violation"])


def test_agents_relative_oracle_matches_reference():
    """oracle/observations.py against Simulator.get_all_agents_relative of the unmodified reference."""
    from oracle import observations
    g = util.golden("relative")
    for key, excl in (("relative_excl", True), ("relative_all", False)):
        got = observations.agents_relative(g["absolute"], exclude_self=excl)
        assert got.shape == g[key].shape
        np.testing.assert_allclose(got, g[key], rtol=1e-5, atol=2e-5)
        assert np.array_equal(got[..., 3:], g[key][..., 3:])
    assert np.abs(g["absolute"][..., 2]).max() > np.pi          # the angle wrap is exercised


def test_npc_oracle_matches_reference():
    """oracle/npc.py against ReplayController + SpawnController stepping inside the unmodified reference Simulator."""
    from oracle import npc
    g = util.golden("npc")
    T = g["replay"].shape[2]
    state, present = g["replay"][:, :, 0], g["replay_present"][:, :, 0]
    t_replay = 0
    for step in range(g["npc_state"].shape[0]):
        t_replay = (t_replay + 1) % T
        state, present = npc.npc_advance(state, present, g["replay"], g["replay_present"], t_replay, g["boundary"],
                                         g["spawn_states"], g["spawn_masks"], step)
        assert np.array_equal(present, g["npc_present"][step]), step
        assert np.array_equal(state, g["npc_state"][step]), step
    # despawning and spawning both happen in the fixture
    inside = npc.is_inside_polygon(g["replay"][:, :, 1, :2], g["boundary"])
    assert 0 < inside.sum() < inside.size and g["spawn_masks"].any()


def test_goals_oracle_matches_reference():
    """oracle/goals.py against WaypointGoal stepping inside the unmodified reference Simulator."""
    from oracle import goals, kinematic as OK
    import torch
    g = util.golden("goals")
    mask, state = g["mask"], np.zeros(g["mask"].shape[:2] + (1,), np.int64)
    st = torch.tensor(g["state0"])
    for t in range(g["actions"].shape[0]):
        st = OK.bicycle_step(st, torch.tensor(g["actions"][t]), torch.tensor(g["lr"]), 0.1, True)
        mask, state = goals.waypoint_step(st.numpy()[..., :2], g["waypoints"], mask, state, 2.0)
        assert np.array_equal(state, g["goal_state"][t]), t
        assert np.array_equal(mask, g["goal_mask"][t]), t
        for count, kw, km in ((1, "wp1", "m1"), (2, "wp2", "m2")):
            w, m = goals.gather(g["waypoints"], mask, state, count)
            assert np.array_equal(w, g[kw][t]) and np.array_equal(m, g[km][t])
    assert g["goal_state"][-1].max() == g["mask"].shape[2] - 1 and g["goal_state"][-1].min() < g["mask"].shape[2] - 1


def test_noise_oracle_matches_reference():
    """oracle/observations.py against StandardSensingObservationNoise of the unmodified reference (same normal deviates)."""
    from oracle import observations
    g = util.golden("noise")
    A = g["agent_state"].shape[1]
    all_state = np.concatenate([g["agent_state"], g["npc_state"]], 1)
    all_size = np.concatenate([g["agent_size"], g["npc_size"]], 1)
    base = np.concatenate([g["present"], g["npc_present"]], 1)
    np.testing.assert_allclose(observations.noisy_state(all_state, A, g["eps"]), g["noisy_state"], rtol=1e-6, atol=1e-6)
    mask = observations.noisy_present_mask(all_state, all_size, base, A)
    assert np.array_equal(mask, g["noisy_present"])
    assert (base[:, None] & ~mask).sum() > 10                    # occlusion happens in the fixture
    absolute = np.concatenate([g["noisy_state"][..., :3], g["noisy_size"], g["noisy_present"][..., None]], -1).astype(np.float32)
    np.testing.assert_allclose(absolute, g["noisy_absolute"], rtol=1e-6, atol=1e-6)
