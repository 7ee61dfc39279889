"""The drop-in seam on the B200, behind the REAL reference: INTEGRATION.md section 3 verbatim.  `B200Renderer`,
`B200BirdviewMeshGenerator` and `KinematicBicycle` are injected into the UNMODIFIED reference `Simulator`
(baseline/_ref, installed by baseline/install_ref.sh and shipped with the repository snapshot) running on CUDA, and
compared with the stock reference (cv2 renderer, CPU) on the same box: states, birdviews, and the FastSimulator
overrides of compute_collision / compute_offroad against the reference's own Python loops."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _reference():
    from oracle.ref_harness import import_reference, reference_available
    if not reference_available():
        pytest.skip("the reference is not installed (baseline/install_ref.sh)")
    import_reference()


def test_injected_into_the_reference_simulator_on_cuda():
    _reference()
    from torchdrivesim.simulator import Simulator, TorchDriveConfig
    from torchdrivesim.rendering import CV2RendererConfig
    from torchdrivesim.kinematic import KinematicBicycle as RefBicycle
    from torchdrivesim.map import find_map_config, traffic_controls_from_map_config
    import torchdrivesim_b200 as tds

    dev = torch.device("cuda:0")
    name = "carla_Town01"
    m = util.load_map_np(name)
    rng = np.random.default_rng(17)
    B, A = 2, 8
    state, size, types, present = util.random_scene(m, B, A, rng, spread=8.0, absent_p=0.2)
    present[:, 0] = True
    cfgm = find_map_config(name)
    lr = torch.full((B, A), util.VEH[2])
    actions = torch.tensor(rng.uniform(-1, 1, (3, B, A, 2)).astype(np.float32))
    L = len([t for t in m["stopline_types"] if t == "traffic_light"])
    tl_states = rng.integers(0, 3, (B, L))

    def controls(device):
        tc = {k: v.extend(B) for k, v in traffic_controls_from_map_config(cfgm).items()}
        tc["traffic_light"].set_state(torch.tensor(tl_states))
        return {k: v.to(device) for k, v in tc.items()}

    # ---- the stock reference on the CPU (cv2 renderer)
    km = RefBicycle(left_handed=True)
    km.set_params(lr=lr)
    km.set_state(torch.tensor(state))
    ref = Simulator(cfg=TorchDriveConfig(left_handed_coordinates=True, renderer=CV2RendererConfig(left_handed_coordinates=True)),
                    road_mesh=cfgm.road_mesh.expand(B), kinematic_model=km, agent_size=torch.tensor(size),
                    initial_present_mask=torch.tensor(present), traffic_controls=controls("cpu"))

    # ---- our objects injected into the unmodified reference Simulator, on CUDA
    town = tds.StaticMap.from_birdview_mesh(cfgm.road_mesh, left_handed=True)
    renderer = tds.B200Renderer(tds.B200RendererConfig(left_handed_coordinates=True))
    km2 = tds.KinematicBicycle(left_handed=True)
    km2.set_params(lr=lr.to(dev))
    km2.set_state(torch.tensor(state, device=dev))
    tc = controls(dev)
    agent_size = torch.tensor(size, device=dev)
    gen = tds.B200BirdviewMeshGenerator(town, renderer.color_map, renderer.rendering_levels, batch_size=B)
    gen.initialize_actors_mesh(agent_size, torch.zeros(B, A, dtype=torch.long, device=dev), ["vehicle"])
    gen.initialize_traffic_controls_mesh(tc)

    class FastSimulator(Simulator):
        def compute_collision(self, agent_types=None):
            s, sz = self.get_state(), self.get_agent_size()[..., :2]
            box = torch.cat([s[..., :2], sz, s[..., 2:3]], -1)
            return tds.collision_allpairs(box, box, self.get_all_agent_present_mask(), self.cfg.collision_metric.value)

        def compute_offroad(self):
            return tds.offroad_infraction_loss(self.get_state(), self.get_agent_size(), town,
                                               threshold=self.cfg.offroad_threshold) * self.get_present_mask()

    sim = FastSimulator(road_mesh=cfgm.road_mesh.expand(B).to(dev), kinematic_model=km2, agent_size=agent_size,
                        initial_present_mask=torch.tensor(present, device=dev), cfg=TorchDriveConfig(left_handed_coordinates=True),
                        renderer=renderer, birdview_mesh_generator=gen, traffic_controls=tc)

    total_bad = 0
    for t in range(actions.shape[0]):
        ref.step(actions[t])
        sim.step(actions[t].to(dev))
        np.testing.assert_allclose(sim.get_state().cpu().numpy(), ref.get_state().numpy(), rtol=1e-5, atol=1e-6)
        img_ref = ref.render_egocentric()
        img = sim.render_egocentric()
        assert img.is_cuda and img.shape == img_ref.shape == (B, A, 3, 64, 64)
        bad = int((img.cpu() != img_ref).any(2).sum())
        total_bad += bad
        assert bad <= 0.001 * B * A * 64 * 64, f"step {t}: {bad} mismatching pixels"           # the north star's bound
        coll_ref, off_ref = ref.compute_collision(), ref.compute_offroad()
        np.testing.assert_allclose(sim.compute_collision().cpu().numpy(), coll_ref.numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(sim.compute_offroad().cpu().numpy(), off_ref.numpy(), rtol=1e-5, atol=1e-5)
        # feed the reference's state to both so that 1-ulp sin/cos differences of the step do not accumulate
        sim.set_state(ref.get_state().to(dev))
    assert float(img_ref.max()) > 0 and float(coll_ref.max()) > 0
    print(f"reference injection on CUDA: {total_bad} mismatching pixels over {actions.shape[0] * B * A} cameras")
    # the reference's own batch plumbing works on our objects
    sub = sim.select_batch_elements(torch.tensor([1], device=dev))
    assert sub.get_state().shape == (1, A, 4)
