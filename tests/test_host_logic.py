"""Host-side logic of the Python layer on CPU: palette, map loading, generator / simulator plumbing, error
behaviour without a GPU (the product path must fail loudly, never fall back)."""
import os

import numpy as np
import pytest
import torch

import torchdrivesim_b200 as tds
from torchdrivesim_b200 import palette as P
from tests import util


def test_palette_ranks_and_colors():
    pal = P.build_palette(P.get_default_color_map(), P.get_default_rendering_levels(),
                          ["road", "left_lane", "right_lane", "vehicle", "direction", "traffic_light_red",
                           "traffic_light_green", "traffic_light_yellow"], ["vehicle"], True, ["red", "yellow", "green"])
    cid = P.class_id
    rank = lambda n: pal.rank[cid(n)]
    assert rank("road") < rank("right_lane") < rank("left_lane") < rank("traffic_light_green") \
        < rank("traffic_light_yellow") < rank("traffic_light_red") < rank("vehicle") < rank("direction")
    assert tuple(pal.rgb[cid("road")]) == (155, 155, 155) and tuple(pal.rgb[cid("vehicle")]) == (32, 74, 135)
    assert pal.active[cid("road")] == 1 and pal.active[cid("pedestrian")] == 0
    assert pal.agent_type_class[0] == cid("vehicle") and pal.agent_type_class[1] == -1
    assert [pal.tl_state_class[i] for i in range(3)] == [cid("traffic_light_red"), cid("traffic_light_yellow"),
                                                         cid("traffic_light_green")]
    # every default colour survives the cv2 quantisation floor(c/255*0.999*256) unchanged
    for rgb in P.get_default_color_map().values():
        assert P.quantize_color(rgb) == tuple(rgb)


def test_static_map_loading_and_categories():
    m = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    assert m.verts.shape == (46233, 2) and m.faces.shape == (30750, 3) and m.left_handed
    assert sorted(set(m.face_category_names)) == ["left_lane", "right_lane", "road"]
    assert m.traffic_light_poses().shape == (36, 5)
    c = m.world_center
    assert 150 < c[0] < 250 and 100 < c[1] < 220
    with pytest.raises(tds._lib.TdsError):
        m.handle("cpu")


@pytest.mark.needs_reference
def test_mesh_json_loader_matches_npz_fixture():
    root = "/root/reference/torchdrivesim/resources/maps/carla_Town02"
    a = tds.StaticMap.from_mesh_json(os.path.join(root, "carla_Town02_mesh.json"),
                                     os.path.join(root, "carla_Town02_stoplines.json"), left_handed=True)
    b = tds.StaticMap.from_npz(util.map_path("carla_Town02"))
    assert np.array_equal(a.verts, b.verts) and np.array_equal(a.faces, b.faces)
    assert a.face_category_names == b.face_category_names and np.array_equal(a.stoplines, b.stoplines)


def _sim(B=4, A=3, lights=True):
    m = tds.StaticMap.from_npz(util.map_path("carla_Town02"))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), 1.96))
    km.set_state(torch.arange(B * A * 4, dtype=torch.float32).reshape(B, A, 4))
    tc = None
    if lights:
        pos = torch.tensor(m.traffic_light_poses())[None].expand(B, -1, -1).contiguous()
        tc = {"traffic_light": tds.TrafficLightControl(pos, replay_states=torch.arange(3).repeat(B, pos.shape[1], 2))}
    return tds.Simulator(m, km, torch.ones(B, A, 2), torch.ones(B, A, dtype=torch.bool),
                         tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls=tc)


def test_simulator_plumbing_without_gpu():
    sim = _sim()
    assert sim.batch_size == 4 and sim.agent_count == 3 and sim.action_size == 2
    assert sim.kinematic_model.left_handed
    assert sim.get_world_center().shape == (4, 2)
    scene = sim.birdview_mesh_generator.generate(
        3, agent_state=sim.get_state()[:, None].expand(-1, 3, -1, -1),
        present_mask=sim.get_present_mask()[:, None].expand(-1, 3, -1), traffic_lights=sim.traffic_controls["traffic_light"])
    assert scene.agent_state.shape == (4, 3, 4) and scene.present.shape == (4, 3) and scene.tl_corners.shape == (4, 24, 4, 2)
    assert scene.slice(1, 3).agent_state.shape == (2, 3, 4)
    sub = sim.select_batch_elements(torch.tensor([2, 0]), in_place=False)
    assert sub.batch_size == 2 and torch.equal(sub.get_state(), sim.get_state()[[2, 0]])
    assert sub.traffic_controls["traffic_light"].corners.shape[0] == 2 and sim.batch_size == 4
    cp = sim.copy()
    cp.set_state(sim.get_state() + 1)
    assert not torch.equal(cp.get_state(), sim.get_state())
    # traffic light replay advances with the step counter (traffic_controls.py:127-136)
    tl = sim.traffic_controls["traffic_light"]
    tl.step(4)
    assert int(tl.state[0, 0]) == 1
    # there is no CPU fallback: compute on CPU tensors raises
    for call in (lambda: sim.step(torch.zeros(4, 3, 2)), sim.compute_collision, sim.compute_offroad, sim.render_egocentric):
        with pytest.raises(tds._lib.TdsError):
            call()
    with pytest.raises(tds._lib.TdsError):
        sim.step(torch.zeros(4, 2, 2))


def test_generator_rejects_unsupported_inputs():
    sim = _sim(lights=False)
    gen = sim.birdview_mesh_generator
    st = sim.get_state()
    per_cam = st[:, None].repeat(1, 3, 1, 1)                      # materialised per-camera states
    with pytest.raises(tds._lib.TdsError):
        gen.generate(3, agent_state=per_cam, present_mask=None)
    with pytest.raises(tds._lib.TdsError):
        gen.generate(1, agent_state=st[:, None], custom_agent_colors=torch.zeros(4, 2, 6, 3))   # colours per camera: [B,Nc,N,3]
    with pytest.raises(tds._lib.TdsError):
        gen.generate(1, agent_state=st[:, None], waypoints=torch.zeros(4, 2, 2, 2))          # Nc mismatch
    big = gen.expand(2)
    assert big.agent_size.shape[0] == 8 and big.world_center.shape[0] == 8


def test_custom_agent_colours_become_palette_classes():
    """generate(custom_agent_colors=...) (mesh.py:1092-1099): every distinct (agent type, quantized colour) pair is a
    class of the scene's palette at the level of the agent's type; too many distinct colours raise."""
    from torchdrivesim_b200.palette import class_names, class_id
    sim = _sim(lights=False)
    gen = sim.birdview_mesh_generator
    st = sim.get_state()
    B, N = st.shape[0], st.shape[1]
    col = torch.zeros(B, 1, N, 3)
    col[..., 0] = 1.0                    # red everywhere ...
    col[0, 0, 1] = torch.tensor([0.2, 0.9, 0.3])
    scene = gen.generate(1, agent_state=st[:, None], custom_agent_colors=col)
    base = len(class_names())
    assert scene.agent_class.dtype == torch.uint8 and tuple(scene.agent_class.shape) == (B, 1, N)
    assert sorted(set(scene.agent_class.flatten().tolist())) == [base, base + 1]
    assert sorted(c for _, c in scene.custom_classes) == [(51, 230, 76), (255, 0, 0)]
    pal = scene.palette(sim.renderer.color_map, sim.renderer.rendering_levels)
    assert pal.n_classes == base + 2 and pal.rank[base] == pal.rank[class_id("vehicle")] and pal.active[base + 1] == 1
    big = _sim(B=4, A=6, lights=False)          # 24 distinct colours do not fit 32 classes
    with pytest.raises(tds._lib.TdsError):
        big.birdview_mesh_generator.generate(1, agent_state=big.get_state()[:, None], custom_agent_colors=torch.rand(4, 1, 6, 3))


def test_add_static_meshes_extends_every_map():
    sim = _sim(lights=False)
    gen = sim.birdview_mesh_generator
    nf = gen.mapset.maps[0].faces.shape[0]
    extra = tds.StaticMap(np.array([[0, 0], [1, 0], [0, 1]], np.float32), np.array([[0, 1, 2]], np.int32), ["map_boundary"],
                          np.zeros(3, np.int64))
    gen.add_static_meshes([extra])
    m = gen.mapset.maps[0]
    assert m.faces.shape[0] == nf + 1 and m.face_category_names[-1] == "map_boundary" and int(m.faces[-1].min()) == m.verts.shape[0] - 3
    assert "map_boundary" in gen.mapset.static_categories()


def test_mesh_pickle_loader_never_runs_a_nested_pickle(tmp_path):
    """StaticMap.from_mesh_pickle: a payload that smuggles a forbidden global through the legacy
    torch.storage._load_from_bytes hook is rejected instead of being unpickled a second time without restrictions."""
    import io
    import pickle

    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('os').getpid()",))

    inner = io.BytesIO()
    pickle.dump(Evil(), inner)          # what torch.load(weights_only=False) would happily execute

    class Smuggle:
        def __reduce__(self):
            import torch.storage
            return (torch.storage._load_from_bytes, (inner.getvalue(),))

    p = tmp_path / "evil.pkl"
    with open(p, "wb") as f:
        pickle.dump({"verts": Smuggle()}, f)
    with pytest.raises(Exception) as e:
        tds.StaticMap.from_mesh_pickle(str(p))
    assert "eval" in str(e.value) or "Unsupported" in str(e.value) or "weights" in str(e.value).lower() or "not allowed" in str(e.value)


def test_waypoint_triangles_follow_the_reference_layout():
    """generate(waypoints=...) (mesh.py:1120-1145): 10-face discs of radius 2 m, masked ones collapse onto the centre of
    the camera's first waypoint; the disc template equals the oracle's restatement of generate_disc_mesh."""
    from oracle import raster as R
    sim = _sim(lights=False)
    gen = sim.birdview_mesh_generator
    B = sim.batch_size
    wp = torch.arange(B * 2 * 3 * 2, dtype=torch.float32).reshape(B, 2, 3, 2)
    mask = torch.ones(B, 2, 3, dtype=torch.bool)
    mask[0, 1, 2] = False
    tris, cls = gen._waypoint_triangles(2, wp, mask)
    assert tris.shape == (B, 2, 30, 3, 2) and cls.shape == (B, 2, 30) and cls.dtype == torch.int32
    dv, df = R.disc_template()
    assert np.array_equal(tris[1, 0, :10].numpy(), (dv + wp[1, 0, 0].numpy())[df])
    assert np.array_equal(tris[0, 1, 20:].numpy(), np.broadcast_to(wp[0, 1, 0].numpy(), (10, 3, 2)))
    scene = gen.generate(2, agent_state=sim.get_state()[:, None].expand(-1, 2, -1, -1), waypoints=wp,
                         waypoints_rendering_mask=mask)
    assert scene.cam_tris.shape == (B, 2, 30, 3, 2) and scene.slice(1, 3).cam_tri_class.shape[0] == 2


def test_fit_action_inverts_the_oracle_step():
    from oracle import kinematic as K
    gen = torch.Generator().manual_seed(0)
    st = torch.cat([torch.rand(5, 6, 2, generator=gen) * 100, torch.rand(5, 6, 1, generator=gen) * 6 - 3,
                    torch.rand(5, 6, 1, generator=gen) * 5 + 1], -1)
    act = torch.rand(5, 6, 2, generator=gen) * 0.8 - 0.4
    lr = torch.full((5, 6), 1.96)
    for lh in (False, True):
        km = tds.KinematicBicycle(left_handed=lh)
        km.set_params(lr=lr)
        km.set_state(st)
        nxt = K.bicycle_step(st, act, lr, 0.1, lh)
        np.testing.assert_allclose(km.fit_action(nxt).numpy(), act.numpy(), atol=2e-4)
    km2 = km.copy()
    assert torch.equal(km2.get_state(), km.get_state()) and torch.equal(km2.lr, km.lr) and km2.left_handed
    km2.extend(2)
    assert km2.get_state().shape[0] == 10 and km2.lr.shape[0] == 10


def test_traffic_control_corners_and_masking():
    pos = torch.tensor([[[10.0, 5.0, 1.0, 4.0, 0.3], [0.0, 0.0, 2.0, 2.0, 0.0]]])
    tc = tds.TrafficLightControl(pos, mask=torch.tensor([[True, False]]))
    from oracle.raster import box_corners
    np.testing.assert_allclose(tc.corners[0, 0].numpy(), box_corners(pos[0, :1].numpy())[0], rtol=1e-6, atol=1e-6)
    assert float(tc.corners[0, 1].max()) == -1000.0          # masked controls sit at -1000 (traffic_controls.py:33)
    assert tc.allowed_states == ["red", "yellow", "green"]
    assert tc.extend(3, in_place=False).corners.shape[0] == 3


def test_reference_mesh_wire_formats():
    """StaticMap reads the files the reference's BirdviewMesh.pickle / BirdviewMesh.save wrote (no reference needed)."""
    import os
    import pickle
    import tempfile
    import torchdrivesim_b200 as tds
    d = os.path.join(os.path.dirname(__file__), "golden", "maps")
    arrays = np.load(os.path.join(d, "town02_window_arrays.npz"))
    for m in (tds.StaticMap.from_mesh_pickle(os.path.join(d, "town02_window.pkl")),
              tds.StaticMap.from_mesh_json(os.path.join(d, "town02_window_mesh.json"))):
        assert np.array_equal(m.verts, arrays["verts"]) and np.array_equal(m.faces, arrays["faces"])
        assert np.array_equal(m.vert_category, arrays["vert_category"])
        assert list(m.categories) == [str(c) for c in arrays["categories"]]
        assert m.faces.shape[0] > 100 and set(m.face_category_names) == {"road", "left_lane", "right_lane"}
    # anything but tensors inside the mesh object is refused
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump({"verts": os.system}, f)
    try:
        with pytest.raises(pickle.UnpicklingError):
            tds.StaticMap.from_mesh_pickle(f.name)
    finally:
        os.unlink(f.name)


def test_traffic_light_schedule_matches_reference():
    """The light schedule of carla_Town02 (TrafficLightController, traffic_lights.py:159-301) ticked 400 x 0.1 s from
    a fixed state equals the reference tick for tick, and `unroll` produces the same states as a replay tensor."""
    import json
    import os
    import torchdrivesim_b200 as tds
    d = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(d, "light_schedule.npz"))
    path = os.path.join(d, "maps", "carla_Town02_traffic_light_controller.json")
    ctrl = tds.TrafficLightController.from_json(path)
    assert ctrl.get_number_of_light_groups() == g["machine"].shape[1]
    start = [(int(s), float(r)) for s, r in g["start"]]
    ctrl.set_to(start)
    ids = g["ids"].tolist()
    for t in range(g["lights"].shape[0]):
        assert ctrl.state_per_machine == g["machine"][t].tolist(), t
        assert ctrl.time_remaining == g["remaining"][t].tolist(), t           # the same float arithmetic
        assert ctrl.current_state_tensor(ids).tolist() == g["lights"][t].tolist(), t
        ctrl.tick(0.1)
    ctrl.set_to(start)
    replay = ctrl.unroll(ids, 0.1, g["lights"].shape[0])
    assert replay.dtype == torch.int64 and np.array_equal(replay.numpy(), g["lights"].T)
    assert len(np.unique(replay.numpy())) == 3                                 # red, yellow and green all occur
    # the stop lines of the map name the same lights as the schedule
    stop = json.load(open(os.path.join(d, "maps", "carla_Town02_stoplines.json")))
    assert [s["actor_id"] for s in stop if s["agent_type"] == "traffic_light"] == ids


def test_compound_npc_controller_gathers_by_assignment():
    """CompoundNPCController (simulator.py:206-250): NPC tensors are taken from the controller each NPC is assigned to
    and handed back to all controllers (host-side plumbing; the per-step advance is the CUDA path of test_gpu_npc)."""
    B, Np = 2, 5
    mk = lambda v: tds.NPCController(torch.full((B, Np, 2), float(v)), torch.full((B, Np, 4), float(v)),
                                     torch.full((B, Np), bool(v % 2)), torch.full((B, Np), v, dtype=torch.long))
    a, b = mk(1), mk(2)
    idx = torch.tensor([[0, 1, 0, 1, 1], [1, 1, 0, 0, 0]])
    c = tds.CompoundNPCController([a, b], idx)
    want = (idx + 1).float()
    assert torch.equal(c.get_npc_state()[..., 0], want) and torch.equal(c.get_npc_size()[..., 1], want)
    assert torch.equal(c.get_npc_present_mask(), idx == 0) and torch.equal(c.get_npc_types(), idx + 1)
    assert a.npc_state is c.npc_state and b.npc_present_mask is c.npc_present_mask
    one = c.select_batch_elements(torch.tensor([1]), in_place=False).extend(3)
    assert one.npc_state.shape == (3, Np, 4) and one.controller_indices.shape == (3, Np) and c.npc_state.shape == (B, Np, 4)
    assert torch.equal(one.get_npc_state()[0, :, 0], want[1])


def test_torch_library_ops_are_registered_with_shape_functions():
    """torch.ops.tds_b200.* (torchdrivesim_b200/torch_ops.py): registered custom operators whose fake-tensor kernels give
    the output shapes and dtypes without a GPU (what torch.compile / torch.export trace through)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        s, a, lr = torch.empty(2, 3, 4), torch.empty(2, 3, 2), torch.empty(2, 3)
        assert torch.ops.tds_b200.kinematic_step(s, a, lr, None, 0, 0.1, True).shape == (2, 3, 4)
        b, m = torch.empty(2, 3, 5), torch.empty(2, 3, dtype=torch.bool)
        out, arg = torch.ops.tds_b200.collision_allpairs(b, b, m, 0, True)
        assert out.shape == (2, 3) and arg.dtype == torch.int32
        assert torch.ops.tds_b200.collision_pairwise(b, b, 1).shape == (2, 3)
        assert torch.ops.tds_b200.agent_boxes(s, torch.empty(2, 3, 2)).shape == (2, 3, 5)
    with pytest.raises(tds._lib.TdsError):          # and there is no CPU kernel behind them
        torch.ops.tds_b200.agent_boxes(torch.zeros(1, 1, 4), torch.zeros(1, 1, 2))
