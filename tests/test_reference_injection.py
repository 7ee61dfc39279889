"""Drop-in seam, checked in the build container (needs /root/reference): our KinematicBicycle, B200Renderer and
B200BirdviewMeshGenerator are injected into the UNMODIFIED reference `Simulator`.  There is no GPU here, so the
two C-ABI calls are replaced by the oracle for the duration of the test (test infrastructure only); what is
verified is the host logic: that the reference's step()/render_egocentric() drive our objects with the
arguments they expect and that the result equals the stock reference (cv2 backend)."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.needs_reference


def test_injected_objects_reproduce_the_stock_reference(monkeypatch):
    from oracle.ref_harness import import_reference
    import_reference()
    from torchdrivesim.simulator import Simulator, TorchDriveConfig
    from torchdrivesim.rendering import CV2RendererConfig
    from torchdrivesim.kinematic import KinematicBicycle as RefBicycle
    from torchdrivesim.map import find_map_config, traffic_controls_from_map_config
    import torchdrivesim_b200 as tds
    from torchdrivesim_b200 import ops
    from oracle import kinematic as OK

    name = "carla_Town02"
    m = util.load_map_np(name)
    rng = np.random.default_rng(3)
    B, A = 2, 4
    state, size, types, present = util.random_scene(m, B, A, rng, absent_p=0.25)
    present[:, 0] = True
    cfgm = find_map_config(name)
    lr = torch.full((B, A), util.VEH[2])
    action = torch.tensor(rng.uniform(-1, 1, (B, A, 2)).astype(np.float32))

    def controls():
        tc = {k: v.extend(B) for k, v in traffic_controls_from_map_config(cfgm).items()}
        tc["traffic_light"].set_state(torch.tensor(rng_states))
        return tc
    rng_states = rng.integers(0, 3, (B, 24))

    # ---- stock reference
    km = RefBicycle(left_handed=True)
    km.set_params(lr=lr)
    km.set_state(torch.tensor(state))
    ref = Simulator(cfg=TorchDriveConfig(left_handed_coordinates=True, renderer=CV2RendererConfig(left_handed_coordinates=True)),
                    road_mesh=cfgm.road_mesh.expand(B), kinematic_model=km, agent_size=torch.tensor(size),
                    initial_present_mask=torch.tensor(present), traffic_controls=controls())
    ref.step(action)
    ref_img = ref.render_egocentric()

    # ---- the C ABI replaced by the oracle (no GPU in this container)
    def fake_kinematic_step(st, act, lr_, model, uniform_model, params, out=None):
        assert uniform_model == tds._lib.MODEL_BICYCLE and params.left_handed == 1
        return OK.bicycle_step(st, act, lr_, params.dt, True)

    def fake_raster(mapset, palette, cam_xy, cam_sc, agent_state, agent_size, agent_type, present_, tl_corners, tl_state,
                    rect_corners, rect_class, res, fov, out=None, workspace=None, cam_tris=None, cam_tri_class=None, **kw):
        assert cam_xy.shape == (B, A, 2) and agent_state.shape == (B, A, 4) and present_.shape == (B, A)
        assert tl_corners.shape == (B, 24, 4, 2) and tl_state.shape == (B, 24)
        imgs = util.oracle_render_batch(m, agent_state.numpy(), agent_size.numpy(), agent_type.numpy(), present_.numpy(),
                                        ["vehicle"], tl_corners.numpy(), tl_state.numpy(), cam_xy.numpy(), cam_sc.numpy(),
                                        res, fov)
        return torch.tensor(np.stack([np.stack([imgs[(b, c)] for c in range(A)]) for b in range(B)]))
    monkeypatch.setattr(ops, "kinematic_step", fake_kinematic_step)
    monkeypatch.setattr(ops, "raster_birdview", fake_raster)
    monkeypatch.setattr(tds._lib, "load", lambda: type("L", (), {"tds_raster_workspace_bytes": staticmethod(lambda *a: 0)})())

    # ---- our objects injected into the unmodified reference Simulator
    town = tds.StaticMap.from_birdview_mesh(cfgm.road_mesh, left_handed=True)
    renderer = tds.B200Renderer(tds.B200RendererConfig(left_handed_coordinates=True))
    km2 = tds.KinematicBicycle(left_handed=True)
    km2.set_params(lr=lr)
    km2.set_state(torch.tensor(state))
    tc = controls()
    gen = tds.B200BirdviewMeshGenerator(town, renderer.color_map, renderer.rendering_levels, batch_size=B)
    gen.initialize_actors_mesh(torch.tensor(size), torch.zeros(B, A, dtype=torch.long), ["vehicle"])
    gen.initialize_traffic_controls_mesh(tc)
    sim = Simulator(cfg=TorchDriveConfig(left_handed_coordinates=True), road_mesh=cfgm.road_mesh.expand(B),
                    kinematic_model=km2, agent_size=torch.tensor(size), initial_present_mask=torch.tensor(present),
                    renderer=renderer, birdview_mesh_generator=gen, traffic_controls=tc)
    sim.step(action)
    img = sim.render_egocentric()
    assert torch.equal(sim.get_state(), ref.get_state())
    assert img.shape == ref_img.shape == (B, A, 3, 64, 64)
    bad = int((img != ref_img).any(2).sum())
    assert bad <= 0.001 * B * A * 64 * 64, f"{bad} mismatching pixels"
    # the reference's own batch plumbing works on our objects
    sub = sim.select_batch_elements(torch.tensor([1]))
    assert sub.get_state().shape == (1, A, 4)
    cp = sim.copy()
    assert cp.kinematic_model is not sim.kinematic_model
