"""Parity of the sm_100a birdview raster (through the Python host layer and the C ABI) with the oracle
and with the committed reference goldens.  Integer raster => the bar is pixel-exact equality."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _gpu_render(mapnames, env_map, state, size, types, present, type_names, tl_corners, tl_state, cam_xy, cam_sc, res, fov):
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    maps = [tds.StaticMap.from_npz(util.map_path(n)) for n in mapnames]
    ms = tds.MapSet(maps, None if env_map is None else torch.as_tensor(env_map, device=dev))
    rend = tds.B200Renderer(tds.B200RendererConfig())
    tc = None
    B = state.shape[0]
    gen = tds.B200BirdviewMeshGenerator(ms, rend.color_map, rend.rendering_levels, batch_size=B)
    gen.initialize_actors_mesh(torch.as_tensor(size, device=dev), torch.as_tensor(types, device=dev), type_names)
    tl = None
    if tl_corners is not None:
        tl = tds.TrafficLightControl(pos=torch.zeros(B, tl_corners.shape[1], 5, device=dev))
        tl.corners = torch.as_tensor(tl_corners, device=dev)
        tl.state = torch.as_tensor(tl_state, device=dev)
        gen.initialize_traffic_controls_mesh({"traffic_light": tl})
    Nc = cam_xy.shape[1]
    st = torch.as_tensor(state, device=dev)
    pr = torch.as_tensor(present, device=dev)
    if pr.dim() == 2:
        pr = pr[:, None].expand(-1, Nc, -1)
    scene = gen.generate(Nc, agent_state=st[:, None].expand(-1, Nc, -1, -1), present_mask=pr, traffic_lights=tl)
    img = rend.render_frame(scene, torch.as_tensor(cam_xy, device=dev), torch.as_tensor(cam_sc, device=dev),
                            res=tds.Resolution(res, res), fov=fov)
    torch.cuda.synchronize()
    return img.reshape(B, Nc, 3, res, res).cpu().numpy()


def _sincos_torch(psi):
    p = torch.as_tensor(psi)
    return torch.stack([torch.sin(p), torch.cos(p)], -1).numpy()


def test_golden_reference_images():
    """CUDA raster vs images produced by the unmodified reference (tests/golden/render.npz)."""
    g = util.golden("render")
    for case in g["cases"]:
        q = lambda k: g[f"{case}/{k}"]
        st = q("state")
        res, fov = int(q("res")), float(q("fov"))
        cam_sc = _sincos_torch(st[..., 2])
        img = _gpu_render([str(q("map"))], None, st, q("size"), q("types"), q("present"),
                          [str(s) for s in q("type_names")], q("tl_corners"), q("tl_state"), st[..., :2].copy(), cam_sc,
                          res, fov)
        ref = q("image").astype(np.float32)
        bad = int((img != ref).any(2).sum())
        # budget of the north star: 0.1 % of pixels; measured: 0
        assert bad <= 0.001 * ref.shape[0] * ref.shape[1] * res * res, f"{case}: {bad} mismatching pixels"
        assert bad == 0, f"{case}: {bad} mismatching pixels (expected exact equality)"


@pytest.mark.parametrize("mapname,res,fov,ped_every,absent_p,lights", [
    ("carla_Town01", 64, 35.0, 0, 0.0, True),
    ("carla_Town02", 64, 35.0, 3, 0.3, True),
    ("carla_Town01", 128, 50.0, 4, 0.2, True),
    ("carla_Town01", 256, 35.0, 0, 0.0, False),
    ("carla_Town02", 64, 100.0, 2, 0.1, True),
    ("carla_Town02", 64, 600.0, 0, 0.0, False),      # whole town in view: > 32 grid rows per camera, sub-pixel faces
    ("carla_Town01", 32, 20.0, 0, 0.0, True),
    ("carla_Town02", 48, 35.0, 2, 0.2, True),        # tile sizes that are not a multiple of 32
    ("carla_Town01", 100, 40.0, 3, 0.0, False),
    ("carla_Town02", 320, 70.0, 0, 0.1, True),
    ("carla_Town10HD", 64, 35.0, 3, 0.1, True),      # a map built from its lanelet2 OSM file (joint lane markings)
    ("carla_Town10HD", 128, 60.0, 0, 0.0, True),
    ("carla_Town01", 512, 70.0, 0, 0.0, False),      # beyond the 448 pixels of round 1 (the planes of 5 classes still fit an SM)
])
def test_vs_oracle_random_scenes(mapname, res, fov, ped_every, absent_p, lights):
    rng = np.random.default_rng(res * 7 + int(fov))
    m = util.load_map_np(mapname)
    B, A = 3, 7
    state, size, types, present = util.random_scene(m, B, A, rng, ped_every=ped_every, absent_p=absent_p)
    present[0, 0] = absent_p == 0.0          # exercise the absent-agent-0 stray pixel (App. C-3)
    tl_corners = tl_state = None
    if lights:
        _, tl_corners, tl_state = util.tl_tensors(m, B, rng)
    cam_xy = state[..., :2].copy()
    cam_sc = _sincos_torch(state[..., 2])
    names = ["vehicle", "pedestrian"]
    img = _gpu_render([mapname], None, state, size, types, present, names, tl_corners, tl_state, cam_xy, cam_sc, res, fov)
    ora = util.oracle_render_batch(m, state, size, types, present, names, tl_corners, tl_state, cam_xy, cam_sc, res, fov)
    bad = sum(int((img[b, c] != o).any(0).sum()) for (b, c), o in ora.items())
    assert bad == 0, f"{bad} mismatching pixels over {len(ora)} cameras"
    assert img.max() > 0


def test_per_camera_mask_and_free_cameras():
    """rendering_mask [B,Nc,N] and cameras that are not on agents (Simulator.render, simulator.py:920)."""
    rng = np.random.default_rng(5)
    m = util.load_map_np("carla_Town01")
    B, A, Nc = 2, 6, 4
    state, size, types, _ = util.random_scene(m, B, A, rng)
    present = rng.uniform(size=(B, Nc, A)) > 0.4
    cam_xy = (state[:, :Nc, :2] + rng.normal(0, 5, (B, Nc, 2))).astype(np.float32)
    cam_sc = _sincos_torch(rng.uniform(0, 6.28, (B, Nc)).astype(np.float32))
    img = _gpu_render(["carla_Town01"], None, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, 64, 35.0)
    ora = util.oracle_render_batch(m, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, 64, 35.0)
    assert sum(int((img[b, c] != o).any(0).sum()) for (b, c), o in ora.items()) == 0


def test_mixed_maps_env_map():
    """Heterogeneous batch: environment b uses maps[env_map[b]] (config 3)."""
    rng = np.random.default_rng(6)
    names = ["carla_Town01", "carla_Town02"]
    ms = [util.load_map_np(n) for n in names]
    env_map = np.array([0, 1, 1, 0], np.int32)
    B, A = 4, 5
    parts = [util.random_scene(ms[env_map[b]], 1, A, rng, ped_every=2) for b in range(B)]
    state, size, types, present = (np.concatenate([p[i] for p in parts]) for i in range(4))
    cam_xy, cam_sc = state[..., :2].copy(), _sincos_torch(state[..., 2])
    img = _gpu_render(names, env_map, state, size, types, present, ["vehicle", "pedestrian"], None, None, cam_xy, cam_sc, 64, 35.0)
    bad = 0
    for b in range(B):
        ora = util.oracle_render_batch(ms[env_map[b]], state[b:b + 1], size[b:b + 1], types[b:b + 1], present[b:b + 1],
                                       ["vehicle", "pedestrian"], None, None, cam_xy[b:b + 1], cam_sc[b:b + 1], 64, 35.0)
        bad += sum(int((img[b, c] != o).any(0).sum()) for (_, c), o in ora.items())
    assert bad == 0


def test_empty_and_edge_inputs():
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    # empty map, no agents -> black image; camera far outside the map -> black image
    empty = tds.StaticMap(np.zeros((0, 2), np.float32), np.zeros((0, 3), np.int32), ["road"], np.zeros((0,), np.int64))
    rend = tds.B200Renderer(tds.B200RendererConfig())
    gen = tds.B200BirdviewMeshGenerator(empty, rend.color_map, rend.rendering_levels, batch_size=2)
    scene = gen.generate(3)
    cam = torch.zeros(2, 3, 2, device=dev)
    sc = torch.tensor([0.0, 1.0], device=dev).expand(2, 3, 2).contiguous()
    img = rend.render_frame(scene, cam, sc)
    assert img.shape == (6, 3, 64, 64) and float(img.abs().max()) == 0.0
    town = tds.StaticMap.from_npz(util.map_path("carla_Town01"))
    gen = tds.B200BirdviewMeshGenerator(town, rend.color_map, rend.rendering_levels, batch_size=1)
    far = torch.tensor([[[5000.0, -5000.0]]], device=dev)
    img = rend.render_frame(gen.generate(1), far, sc[:1, :1])
    assert float(img.abs().max()) == 0.0
    with pytest.raises(tds._lib.TdsError):
        rend.render_frame(gen.generate(1), far, sc[:1, :1], res=tds.Resolution(64, 32))
    with pytest.raises(tds._lib.TdsError):
        rend.render_frame(gen.generate(1), far.cpu(), sc[:1, :1].cpu())


def test_full_size_sampled_cameras_and_determinism():
    """BASELINE config 2 at full size (1024 environments x 64 egocentric cameras, 64x64): 48 sampled cameras against the
    oracle, and two launches of the dynamically scheduled persistent grid give identical images."""
    import bench
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    B, A = bench.ENVS_PER_GPU, bench.AGENTS
    state, size, lr, _ = bench.synth_inputs(B, A, 77, 1)
    town = tds.StaticMap.from_npz(bench.map_npz())
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.tensor(lr, device=dev))
    km.set_state(torch.tensor(state, device=dev))
    sim = tds.Simulator(town, km, torch.tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                        tds.TorchDriveConfig(left_handed_coordinates=True))
    img = sim.render_egocentric()
    again = sim.render_egocentric()
    torch.cuda.synchronize()
    assert img.shape == (B, A, 3, 64, 64)
    assert torch.equal(img, again)
    rng = np.random.default_rng(3)
    cams = [(int(b), int(c)) for b, c in zip(rng.integers(0, B, 48), rng.integers(0, A, 48))]
    m = util.load_map_np(bench.MAP)
    types = np.zeros((B, A), np.int64)
    present = np.ones((B, A), bool)
    st = sim.get_state().cpu().numpy()
    cam_sc = _sincos_torch(st[..., 2])
    ora = util.oracle_render_batch(m, st, size, types, present, ["vehicle"], None, None, st[..., :2].copy(), cam_sc, 64,
                                   bench.FOV, cams=cams)
    bad = sum(int((img[b, c].cpu().numpy() != o).any(0).sum()) for (b, c), o in ora.items())
    assert bad == 0, f"{bad} mismatching pixels over {len(ora)} sampled cameras"


def test_goal_waypoints_golden_and_oracle():
    """Goal-waypoint discs per camera with a rendering mask (BirdviewRGBMeshGenerator.generate(waypoints=...),
    mesh.py:1120-1145): images of the unmodified reference, then a larger random case against the oracle at 128x128."""
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")

    def render(mapname, st, size, tl_corners, tl_state, wp, mask, res, fov):
        B, A = st.shape[:2]
        town = tds.StaticMap.from_npz(util.map_path(mapname))
        km = tds.KinematicBicycle(left_handed=True)
        km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
        km.set_state(torch.as_tensor(st, device=dev))
        tc = None
        if tl_corners is not None:
            tl = tds.TrafficLightControl(pos=torch.zeros(B, tl_corners.shape[1], 5, device=dev))
            tl.corners = torch.as_tensor(tl_corners, device=dev)
            tl.set_state(torch.as_tensor(tl_state, device=dev))
            tc = {"traffic_light": tl}
        sim = tds.Simulator(town, km, torch.as_tensor(size, device=dev), torch.ones(B, A, dtype=torch.bool, device=dev),
                            tds.TorchDriveConfig(left_handed_coordinates=True), traffic_controls=tc)
        s = sim.get_state()
        img = sim.render(s[..., :2], s[..., 2:3], res=tds.Resolution(res, res), fov=fov,
                         waypoints=torch.as_tensor(wp, device=dev), waypoints_rendering_mask=torch.as_tensor(mask, device=dev))
        torch.cuda.synchronize()
        return img.cpu().numpy()

    g = util.golden("render_waypoints")
    img = render(str(g["map"]), g["state"], g["size"], g["tl_corners"], g["tl_state"], g["waypoints"], g["waypoints_mask"],
                 int(g["res"]), float(g["fov"]))
    ref = g["image"].astype(np.float32)
    assert int((img != ref).any(2).sum()) == 0
    assert (ref[:, :, 0] == 139).sum() > 100            # the goal colour is really there

    rng = np.random.default_rng(21)
    m = util.load_map_np("carla_Town02")
    B, A, M = 2, 6, 4
    state, size, types, present = util.random_scene(m, B, A, rng)
    wp = (state[:, :, None, :2] + rng.normal(0, 25, (B, A, M, 2))).astype(np.float32)
    mask = rng.uniform(size=(B, A, M)) > 0.3
    img = render("carla_Town02", state, size, None, None, wp, mask, 128, 60.0)
    cam_sc = _sincos_torch(state[..., 2])
    ora = util.oracle_render_batch(m, state, size, types, np.ones((B, A), bool), ["vehicle"], None, None,
                                   state[..., :2].copy(), cam_sc, 128, 60.0, waypoints=wp, waypoints_mask=mask)
    assert sum(int((img[b, c] != o).any(0).sum()) for (b, c), o in ora.items()) == 0


def _sim(mapname, state, size, present, types=None, names=None, town=None):
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    B, A = state.shape[:2]
    town = town or tds.StaticMap.from_npz(util.map_path(mapname))
    km = tds.KinematicBicycle(left_handed=True)
    km.set_params(lr=torch.full((B, A), util.VEH[2], device=dev))
    km.set_state(torch.as_tensor(state, device=dev))
    return tds.Simulator(town, km, torch.as_tensor(size, device=dev), torch.as_tensor(present, device=dev),
                         tds.TorchDriveConfig(left_handed_coordinates=True),
                         agent_types=None if types is None else torch.as_tensor(types, device=dev), agent_type_names=names)


def test_uint8_and_rank_images_equal_the_float_image():
    """The lossless narrow formats: uint8 RGB == the float32 image cast, palette[rank image] == the uint8 image, through
    render_egocentric(dtype=...) and through the chunked pinned-host path, at a warp-per-camera and a CTA-per-camera size."""
    import torchdrivesim_b200 as tds
    rng = np.random.default_rng(12)
    m = util.load_map_np("carla_Town02")
    for res, fov, B, A in ((64, 35.0, 5, 6), (128, 50.0, 3, 5), (100, 40.0, 2, 4)):
        state, size, types, present = util.random_scene(m, B, A, rng, ped_every=3, absent_p=0.2)
        sim = _sim("carla_Town02", state, size, present, types, ["vehicle", "pedestrian"])
        R = tds.Resolution(res, res)
        f32 = sim.render_egocentric(res=R, fov=fov)
        u8 = sim.render_egocentric(res=R, fov=fov, dtype=torch.uint8)
        rk = sim.render_egocentric(res=R, fov=fov, dtype='rank')
        assert u8.dtype == torch.uint8 and u8.shape == f32.shape and rk.shape == (B, A, res, res)
        assert torch.equal(u8, f32.to(torch.uint8)) and float(f32.max()) > 0
        cam = sim.get_state()
        scene = sim._scene(cam[..., :2], None, None, None, None)
        rgb, cls = sim.renderer.rank_table(scene)
        assert torch.equal(rgb.to(rk.device)[rk.long()].permute(0, 1, 4, 2, 3), u8)
        assert int(rk.max()) < rgb.shape[0] and int(cls[0]) == -1
        for dt, shape in ((torch.float32, f32.shape), (torch.uint8, f32.shape), (torch.uint8, rk.shape)):
            host = torch.empty(shape, dtype=dt).pin_memory()
            sim.render_egocentric_to_host(host, chunk_envs=2, res=R, fov=fov)
            torch.cuda.synchronize()
            want = f32 if dt == torch.float32 else (u8 if len(shape) == 5 else rk)
            assert torch.equal(host, want.cpu())


def test_custom_agent_colours_and_static_meshes_golden():
    """generate(custom_agent_colors=...) and add_static_meshes against images of the unmodified reference
    (tests/golden/render_custom.npz), and the same frame without the custom colours against the oracle."""
    import torchdrivesim_b200 as tds
    dev = torch.device("cuda:0")
    g = util.golden("render_custom")
    st, B, A = g["state"], g["state"].shape[0], g["state"].shape[1]
    names = [str(s) for s in g["type_names"]]
    res, fov = int(g["res"]), float(g["fov"])
    imgs = []
    for b in range(B):          # the extra static mesh differs per environment: one map (set) per environment
        sim = _sim(str(g["map"]), st[b:b + 1], g["size"][b:b + 1], g["present"][b:b + 1], g["types"][b:b + 1], names)
        tl = tds.TrafficLightControl(pos=torch.zeros(1, g["tl_corners"].shape[1], 5, device=dev))
        tl.corners = torch.as_tensor(g["tl_corners"][b:b + 1], device=dev)
        tl.set_state(torch.as_tensor(g["tl_state"][b:b + 1], device=dev))
        sim.traffic_controls = {"traffic_light": tl}
        sim.birdview_mesh_generator.initialize_traffic_controls_mesh(sim.traffic_controls)
        m0 = util.load_map_np(str(g["map"]))
        extra = tds.StaticMap(g["extra_verts"][b], g["extra_faces"][b], ["map_boundary"],
                              np.zeros(len(g["extra_verts"][b]), np.int64))
        sim.birdview_mesh_generator.add_static_meshes([extra])
        img = sim.render_egocentric(res=tds.Resolution(res, res), fov=fov,
                                    custom_agent_colors=torch.as_tensor(g["colors"][b:b + 1], device=dev))
        torch.cuda.synchronize()
        imgs.append(img.cpu().numpy())
        ref = g["image"][b:b + 1].astype(np.float32)
        assert int((imgs[-1] != ref).any(2).sum()) == 0, f"environment {b}"
        assert ((ref[:, :, 0] == 255) & (ref[:, :, 1] == 255) & (ref[:, :, 2] == 0)).sum() > 100     # the extra mesh is there
    # too many distinct colours for the palette: a clear error, not a wrong image
    sim = _sim(str(g["map"]), st[:1], g["size"][:1], g["present"][:1], g["types"][:1], names)
    many = torch.rand(1, A, A, 3, device=dev)
    with pytest.raises(tds._lib.TdsError):
        sim.render_egocentric(custom_agent_colors=many)


def test_dense_scene_unlisted_dynamic_branch():
    """BASELINE config 4 density.  A camera lists the dynamic primitives that reach its view quad - up to 128 (warp per
    camera) or 512 (CTA per camera); beyond that it walks ALL agents instead (20 % of them absent, which that branch must
    skip itself).  260 agents in one view: listed at 256x256, unlisted at 64x64; 700 agents: unlisted at 128x128 too."""
    m = util.load_map_np("carla_Town01")
    names = ["vehicle", "pedestrian"]
    for A, spread, need, cases in ((260, 9.0, 130, ((256, 35.0), (64, 35.0))), (700, 6.0, 512, ((128, 35.0),))):
        rng = np.random.default_rng(44 + A)
        B = 2
        state, size, types, present = util.random_scene(m, B, A, rng, spread=spread, ped_every=5, absent_p=0.2)
        present[:, 0] = [True, False]
        ncam = 8 if A == 260 else 3
        cams = [(b, c) for b in range(B) for c in range(ncam)]
        # cameras on the agents nearest to the middle of the crowd
        near = np.argsort(np.abs(state[..., :2] - np.median(state[..., :2], axis=1, keepdims=True)).max(-1), axis=1)[:, :ncam]
        cam_xy = np.take_along_axis(state[..., :2], near[..., None], 1).copy()
        cam_sc = _sincos_torch(np.take_along_axis(state[..., 2], near, 1))
        # how many agents does a camera see?  (35 m view: everything within ~17 m)
        d = np.abs(state[:, None, :, :2] - cam_xy[:, :, None, :]).max(-1)
        assert ((d < 15.0) & present[:, None, :]).sum(-1).min() > need
        for res, fov in cases:
            img = _gpu_render(["carla_Town01"], None, state, size, types, present, names, None, None, cam_xy, cam_sc, res, fov)
            ora = util.oracle_render_batch(m, state, size, types, present, names, None, None, cam_xy, cam_sc, res, fov, cams=cams)
            bad = sum(int((img[b, c] != o).any(0).sum()) for (b, c), o in ora.items())
            assert bad == 0, f"{A} agents, res {res}: {bad} mismatching pixels over {len(ora)} cameras"


def test_extreme_zoom_huge_coordinates():
    """fov <= 1 m at 448x448 (and 0.25 m at 64x64): road triangles and lane-marking strips project to coordinates
    beyond +-8000 pixels, the 64-bit rule of draw_huge - on the device, for single faces and for strips."""
    rng = np.random.default_rng(9)
    m = util.load_map_np("carla_Town01")
    B, A = 2, 3
    state, size, types, present = util.random_scene(m, B, A, rng, spread=1.0)
    # put cameras right on lane-marking vertices and road vertices
    lane = m["verts"][m["vert_category"] != m["categories"].index("road")]
    state[0, :, :2] = lane[rng.integers(0, lane.shape[0], A)] + rng.normal(0, 0.05, (A, 2))
    cam_xy, cam_sc = state[..., :2].copy(), _sincos_torch(state[..., 2])
    for res, fov in ((448, 1.0), (64, 0.25), (256, 0.6)):
        img = _gpu_render(["carla_Town01"], None, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, res, fov)
        ora = util.oracle_render_batch(m, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, res, fov)
        bad = sum(int((img[b, c] != o).any(0).sum()) for (b, c), o in ora.items())
        assert bad == 0, f"res {res} fov {fov}: {bad} mismatching pixels"
        assert img.max() > 0


def test_strip_stage_equals_face_by_face_walk(monkeypatch):
    """Stage 1S (six-vertex strips, vertices plotted at once) against the same records walked face by face
    (TDS_RASTER_STRIPS=0): identical images on thousands of cameras at several tile sizes and zooms."""
    import bench
    rng = np.random.default_rng(31)
    for mapname, res, fov, B, A in (("carla_Town01", 64, 35.0, 24, 64), ("carla_Town10HD", 64, 20.0, 8, 32),
                                    ("carla_Town02", 128, 80.0, 6, 16), ("carla_Town01", 256, 35.0, 2, 16),
                                    ("carla_Town02", 48, 10.0, 8, 16), ("carla_Town01", 96, 120.0, 4, 16)):
        m = util.load_map_np(mapname)
        state, size, types, present = util.random_scene(m, B, A, rng, spread=30.0)
        cam_xy, cam_sc = state[..., :2].copy(), _sincos_torch(state[..., 2])
        monkeypatch.setenv("TDS_RASTER_STRIPS", "1")
        a = _gpu_render([mapname], None, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, res, fov)
        monkeypatch.setenv("TDS_RASTER_STRIPS", "0")
        b = _gpu_render([mapname], None, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, res, fov)
        assert a.max() > 0 and np.array_equal(a, b), f"{mapname} {res} {fov}: {(a != b).any(2).sum()} pixels differ"


def test_two_pass_and_pattern_table_equal_the_one_pass_walk(monkeypatch):
    """The 64x64 path in its product form - draw pass (sliver quads from the coverage-pattern table, triangles inside the
    image as three line walkers) + finish pass (border-crossing faces, resolve) - against the one-pass kernel, against the
    kernel without the pattern table, and against the face-by-face walk: identical images on thousands of cameras, also
    when the hand-over list of a camera holds only 24 border-crossing faces (TDS_RASTER_CLIP_CAP: most cameras then go on
    the redo list and are rendered by the general kernel behind the two passes)."""
    rng = np.random.default_rng(77)
    for mapname, fov, B, A, spread in (("carla_Town01", 35.0, 32, 64, 30.0), ("carla_Town10HD", 35.0, 8, 64, 25.0),
                                       ("carla_Town02", 60.0, 8, 32, 40.0), ("carla_Town01", 90.0, 4, 48, 60.0),
                                       ("carla_Town01", 18.0, 8, 32, 10.0)):
        m = util.load_map_np(mapname)
        state, size, types, present = util.random_scene(m, B, A, rng, spread=spread)
        cam_xy, cam_sc = state[..., :2].copy(), _sincos_torch(state[..., 2])
        imgs = {}
        for name, env in (("product", {}), ("product in rounds of 23 cameras", {"TDS_RASTER_TWO_PASS_KB": "128"}),
                          ("product with a short face list", {"TDS_RASTER_CLIP_CAP": "24"}),
                          ("one pass", {"TDS_RASTER_TWO_PASS": "0"}),
                          ("no table", {"TDS_RASTER_TWO_PASS": "0", "TDS_RASTER_QUADS": "0"}),
                          ("face walk", {"TDS_RASTER_TWO_PASS": "0", "TDS_RASTER_STRIPS": "0"}),
                          ("general kernel", {"TDS_RASTER_LEAN": "0"})):
            for k in ("TDS_RASTER_TWO_PASS", "TDS_RASTER_TWO_PASS_KB", "TDS_RASTER_CLIP_CAP", "TDS_RASTER_QUADS", "TDS_RASTER_STRIPS", "TDS_RASTER_LEAN"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            imgs[name] = _gpu_render([mapname], None, state, size, types, present, ["vehicle"], None, None, cam_xy, cam_sc, 64, fov)
        ref = imgs["face walk"]
        assert ref.max() > 0
        for name, img in imgs.items():
            assert np.array_equal(img, ref), f"{mapname} fov {fov}: {name}: {(img != ref).any(2).sum()} pixels differ from the face walk"
