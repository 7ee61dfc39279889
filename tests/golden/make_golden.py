"""Generates the golden vectors tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on seeded inputs.

Run in the build container only:   python tests/golden/make_golden.py
The reference's own tests hold no known-answer vectors for this path, so these
files are what pins the oracle (tests/test_oracle_vs_golden.py) and, through it, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_harness import import_reference  # noqa: E402

import_reference()
from torchdrivesim.simulator import Simulator, TorchDriveConfig, CollisionMetric  # noqa: E402
from torchdrivesim.rendering import CV2RendererConfig  # noqa: E402
from torchdrivesim.kinematic import KinematicBicycle, BicycleNoReversing  # noqa: E402
from torchdrivesim.map import find_map_config, traffic_controls_from_map_config  # noqa: E402
from torchdrivesim.infractions import (collision_detection_with_discs, iou_differentiable,  # noqa: E402
                                       offroad_infraction_loss)
from torchdrivesim.utils import Resolution  # noqa: E402
from torchdrivesim.mesh import BirdviewMesh  # noqa: E402

VEH = (4.97, 2.04, 1.96)   # behavior/heuristic.py:11-13
PED = (1.5, 1.5)           # examples/imitation_learning.py:80-81


def road_points(mesh: BirdviewMesh, n: int, gen: torch.Generator) -> torch.Tensor:
    road = mesh.separate_by_category()["road"].verts[0]
    return road[torch.randint(0, road.shape[0], (n,), generator=gen)]


def make_sim(map_name, B, A, gen, types=None, type_names=None, with_lights=True, present=None,
             metric=CollisionMetric.discs, jitter=0.0, npc_controller=None):
    cfgm = find_map_config(map_name)
    mesh = cfgm.road_mesh
    xy = road_points(mesh, B * A, gen).reshape(B, A, 2) + jitter * torch.randn(B, A, 2, generator=gen)
    psi = torch.rand(B, A, 1, generator=gen) * 2 * np.pi
    v = torch.rand(B, A, 1, generator=gen) * 5
    state = torch.cat([xy, psi, v], -1)
    size = torch.tensor(VEH[:2]).expand(B, A, 2).clone()
    lr = torch.full((B, A), VEH[2])
    if types is not None:
        size = torch.where(types.unsqueeze(-1) == 1, torch.tensor(PED), size)
    if present is None:
        present = torch.ones(B, A, dtype=torch.bool)
    km = KinematicBicycle(left_handed=True)
    km.set_params(lr=lr)
    km.set_state(state)
    tc = None
    if with_lights:
        tc = {k: v_.extend(B) for k, v_ in traffic_controls_from_map_config(cfgm).items()}
        tl = tc["traffic_light"]
        tl.set_state(torch.randint(0, 3, tl.state.shape, generator=gen))
    cfg = TorchDriveConfig(left_handed_coordinates=True, collision_metric=metric,
                           renderer=CV2RendererConfig(left_handed_coordinates=True))
    sim = Simulator(cfg=cfg, road_mesh=mesh.expand(B), kinematic_model=km, agent_size=size,
                    initial_present_mask=present, traffic_controls=tc, agent_types=types,
                    agent_type_names=type_names, npc_controller=npc_controller)
    return sim, cfgm


def golden_render():
    out = {}
    gen = torch.Generator().manual_seed(101)
    cases = []
    # (name, map, B, A, res, fov, absent pattern, mixed types)
    cases.append(("town01_64", "carla_Town01", 2, 6, 64, 35.0, None, False))
    cases.append(("town01_absent0", "carla_Town01", 1, 5, 64, 35.0, [0, 3], False))
    cases.append(("town02_mixed_128", "carla_Town02", 1, 6, 128, 50.0, [2], True))
    cases.append(("town01_256", "carla_Town01", 1, 2, 256, 35.0, None, False))
    for name, mp, B, A, res, fov, absent, mixed in cases:
        types = names = None
        if mixed:
            types = (torch.arange(A) % 3 == 2).long().expand(B, A).clone()
            names = ["vehicle", "pedestrian"]
        present = torch.ones(B, A, dtype=torch.bool)
        if absent:
            present[:, absent] = False
        # a tight cluster so that agents see each other
        sim, cfgm = make_sim(mp, B, A, gen, types=types, type_names=names, present=present)
        st = sim.get_state().clone()
        st[:, 1:, :2] = st[:, :1, :2] + 8.0 * torch.randn(B, A - 1, 2, generator=gen)
        sim.set_state(st)
        img = sim.render_egocentric(res=Resolution(res, res), fov=fov)       # [B,A,3,H,W]
        assert float((img - img.round()).abs().max()) == 0.0
        tl = sim.traffic_controls["traffic_light"]
        out[name] = dict(map=mp, state=sim.get_state().numpy(), size=sim.get_agent_size().numpy(),
                         present=present.numpy(), types=(types.numpy() if types is not None else np.zeros((B, A), np.int64)),
                         type_names=np.array(names or ["vehicle"]), tl_state=tl.state.numpy(),
                         tl_corners=tl.corners.numpy(), res=res, fov=fov, image=img.numpy().astype(np.uint8))
        print("render", name, img.shape, "nonzero px", int((img.sum(2) > 0).sum()))
    flat = {f"{k}/{kk}": vv for k, v in out.items() for kk, vv in v.items()}
    np.savez_compressed(os.path.join(HERE, "render.npz"), cases=np.array(list(out)), **flat)


def golden_kinematic():
    gen = torch.Generator().manual_seed(202)
    B, A, T = 3, 7, 12
    res = {}
    for name, cls, lh in (("bicycle_rh", KinematicBicycle, False), ("bicycle_lh", KinematicBicycle, True),
                          ("noreverse_lh", BicycleNoReversing, True)):
        km = cls(left_handed=lh) if cls is KinematicBicycle else cls(left_handed=lh)
        state0 = torch.cat([torch.rand(B, A, 2, generator=gen) * 400, torch.rand(B, A, 1, generator=gen) * 6.28 - 3.14,
                            torch.rand(B, A, 1, generator=gen) * 6 - 1], -1)
        lr = 1.0 + 2 * torch.rand(B, A, generator=gen)
        km.set_params(lr=lr)
        km.set_state(state0)
        actions = torch.rand(T, B, A, 2, generator=gen) * 2 - 1
        traj = [state0]
        for t in range(T):
            km.step(actions[t])
            traj.append(km.get_state())
        res[name] = dict(state0=state0.numpy(), lr=lr.numpy(), actions=actions.numpy(),
                         traj=torch.stack(traj).numpy(), left_handed=lh)
    flat = {f"{k}/{kk}": vv for k, v in res.items() for kk, vv in v.items()}
    np.savez_compressed(os.path.join(HERE, "kinematic.npz"), cases=np.array(list(res)), **flat)
    print("kinematic", list(res))


def golden_collision():
    gen = torch.Generator().manual_seed(303)
    B, A = 3, 12
    present = torch.rand(B, A, generator=gen) > 0.2
    types = (torch.arange(A) % 4 == 3).long().expand(B, A).clone()
    sim, _ = make_sim("carla_Town01", B, A, gen, types=types, type_names=["vehicle", "pedestrian"],
                      with_lights=False, present=present)
    st = sim.get_state().clone()
    st[..., :2] = st[:, :1, :2] + 3.5 * torch.randn(B, A, 2, generator=gen)      # dense: many overlaps
    st[0, 1] = st[0, 0]; st[0, 1, 0] += 0.5                                       # near-coincident pair
    st[1, 2, :3] = st[1, 3, :3]; st[1, 2, 1] += 2.04                              # parallel side-by-side (argmin tie)
    sim.set_state(st)
    coll = sim.compute_collision()
    size = sim.get_agent_size()
    box = torch.cat([st[..., :2], size, st[..., 2:3]], -1)
    n = A
    pair = collision_detection_with_discs(box.unsqueeze(2).expand(-1, -1, n, -1).reshape(B, A * n, 5),
                                          box.unsqueeze(1).expand(-1, A, -1, -1).reshape(B, A * n, 5)).reshape(B, A, n)
    # backward through the aggregate, w.r.t. the state
    st_g = st.clone().requires_grad_(True)
    sim.set_state(st_g)
    sim.compute_collision().sum().backward()
    # IoU, evaluated by the reference in float64
    box64 = box.double()
    iou64 = iou_differentiable(box64.unsqueeze(2).expand(-1, -1, n, -1).reshape(B, A * n, 5),
                               box64.unsqueeze(1).expand(-1, A, -1, -1).reshape(B, A * n, 5)).reshape(B, A, n)
    iou32 = iou_differentiable(box.unsqueeze(2).expand(-1, -1, n, -1).reshape(B, A * n, 5).contiguous(),
                               box.unsqueeze(1).expand(-1, A, -1, -1).reshape(B, A * n, 5).contiguous()).reshape(B, A, n)
    np.savez_compressed(os.path.join(HERE, "collision.npz"), box=box.numpy(), present=present.numpy(),
                        discs_pair=pair.numpy(), discs_collision=coll.detach().numpy(),
                        discs_grad_state=st_g.grad.numpy(), iou64=iou64.numpy(), iou32=iou32.numpy())
    print("collision: discs nonzero pairs", int((pair > 0).sum()), "iou64 nonzero", int((iou64 > 0).sum()),
          "fp32 self-iou==1:", float((torch.diagonal(iou32, dim1=1, dim2=2) > 0.999).float().mean()))


def golden_offroad():
    gen = torch.Generator().manual_seed(404)
    B, A = 1, 24
    sim, cfgm = make_sim("carla_Town01", B, A, gen, with_lights=False, jitter=6.0)
    st = sim.get_state().clone()
    st[0, 0, :2] = torch.tensor([-40.0, -30.0])      # far off the map
    st[0, 1, :2] = torch.tensor([200.0, 160.0])      # block interior
    sim.set_state(st)
    off05 = offroad_infraction_loss(sim.get_state(), sim.get_agent_size(), sim.road_mesh, threshold=0.5,
                                    use_pytorch3d=False)
    off0 = offroad_infraction_loss(sim.get_state(), sim.get_agent_size(), sim.road_mesh, threshold=0.0,
                                   use_pytorch3d=False)
    np.savez_compressed(os.path.join(HERE, "offroad.npz"), map="carla_Town01", state=st.numpy(),
                        size=sim.get_agent_size().numpy(), offroad_thr05=off05.numpy(), offroad_thr0=off0.numpy())
    print("offroad: nonzero", int((off05 > 0).sum()), "of", A, "max", float(off05.max()))


def golden_waypoints():
    """Goal-waypoint discs in the birdview (mesh.py:1120-1145, 1243-1271): M waypoints per camera with a mask."""
    gen = torch.Generator().manual_seed(606)
    B, A, M, res, fov = 2, 5, 3, 64, 35.0
    sim, cfgm = make_sim("carla_Town01", B, A, gen)
    st = sim.get_state().clone()
    st[:, 1:, :2] = st[:, :1, :2] + 8.0 * torch.randn(B, A - 1, 2, generator=gen)
    sim.set_state(st)
    wp = st[:, :, None, :2] + 12.0 * torch.randn(B, A, M, 2, generator=gen)            # around every camera
    wp[0, 0, 0] = st[0, 0, :2] + torch.tensor([16.5, 0.0])                             # one cut by the image border
    mask = torch.rand(B, A, M, generator=gen) > 0.3
    mask[0, 1] = False                                                                  # all masked for one camera
    img = sim.render(st[..., :2], st[..., 2:3], res=Resolution(res, res), fov=fov, waypoints=wp,
                     waypoints_rendering_mask=mask)
    img = img.reshape(B, A, 3, res, res)
    assert float((img - img.round()).abs().max()) == 0.0
    tl = sim.traffic_controls["traffic_light"]
    np.savez_compressed(os.path.join(HERE, "render_waypoints.npz"), map="carla_Town01", state=st.numpy(),
                        size=sim.get_agent_size().numpy(), tl_state=tl.state.numpy(), tl_corners=tl.corners.numpy(),
                        waypoints=wp.numpy(), waypoints_mask=mask.numpy(), res=res, fov=fov,
                        image=img.numpy().astype(np.uint8))
    goal = np.array([139, 64, 0]) * 0.999
    print("waypoints: image", img.shape, "goal-coloured px", int((img[:, :, 0] == np.floor(goal[0] / 255 * 256)).sum()))


def golden_custom_colors():
    """generate(custom_agent_colors=...) (mesh.py:1092-1099) and add_static_meshes (mesh.py:870-883): every agent is
    given a colour per camera; a strip of extra static 'map_boundary' triangles is appended to the background."""
    from torchdrivesim.mesh import BaseMesh, rendering_mesh
    gen = torch.Generator().manual_seed(808)
    B, A, res, fov = 2, 6, 64, 35.0
    types = (torch.arange(A) % 3 == 2).long().expand(B, A).clone()
    names = ["vehicle", "pedestrian"]
    present = torch.ones(B, A, dtype=torch.bool)
    present[0, 0] = False          # absent agent 0: the degenerate face takes ITS custom colour
    present[1, 4] = False
    sim, cfgm = make_sim("carla_Town01", B, A, gen, types=types, type_names=names, present=present)
    st = sim.get_state().clone()
    # agents side by side, 7 m apart (no overlaps: rectangles of equal level are drawn in an undefined order)
    offs = torch.stack([7.0 * (torch.arange(A) - A // 2).float(), 3.0 * (torch.arange(A) % 2).float()], -1)
    st[:, :, :2] = st[:, :1, :2] + offs[None]
    sim.set_state(st)
    palette = torch.tensor([[1.0, 0.0, 0.0], [0.2, 0.9, 0.3], [0.5, 0.5, 0.5], [0.9, 0.8, 0.1]])
    pick = torch.randint(0, 4, (B, A, A), generator=gen)
    colors = palette[pick]                                                  # [B,Nc,A,3] in [0,1]
    # a fan of extra static triangles around the first agent
    c = st[:, 0, :2]
    ang = torch.linspace(0, 2 * np.pi, 7)[:-1]
    ring = torch.stack([torch.cos(ang), torch.sin(ang)], -1) * 9.0
    verts = torch.cat([c[:, None], c[:, None] + ring[None]], 1)             # [B,7,2]
    faces = torch.tensor([[0, 1, 2], [0, 3, 4], [0, 5, 6]])[None].expand(B, -1, -1)
    extra = rendering_mesh(BaseMesh(verts=verts, faces=faces), "map_boundary")
    sim.birdview_mesh_generator.add_static_meshes([extra])
    img = sim.render_egocentric(res=Resolution(res, res), fov=fov, custom_agent_colors=colors)
    assert float((img - img.round()).abs().max()) == 0.0
    tl = sim.traffic_controls["traffic_light"]
    np.savez_compressed(os.path.join(HERE, "render_custom.npz"), map="carla_Town01", state=st.numpy(),
                        size=sim.get_agent_size().numpy(), present=present.numpy(), types=types.numpy(),
                        type_names=np.array(names), tl_state=tl.state.numpy(), tl_corners=tl.corners.numpy(),
                        colors=colors.numpy(), extra_verts=verts.numpy(), extra_faces=faces.numpy(), res=res, fov=fov,
                        image=img.numpy().astype(np.uint8))
    print("custom colours: image", img.shape, "red px", int(((img[:, :, 0] == 255) & (img[:, :, 1] == 0)).sum()),
          "boundary px", int(((img[:, :, 0] == 255) & (img[:, :, 1] == 255) & (img[:, :, 2] == 0)).sum()))


def golden_relative():
    """Non-visual observations (simulator.py:730-781): get_all_agents_absolute / get_all_agents_relative."""
    gen = torch.Generator().manual_seed(707)
    B, A = 3, 9
    present = torch.rand(B, A, generator=gen) > 0.2
    sim, _ = make_sim("carla_Town01", B, A, gen, with_lights=False, present=present)
    st = sim.get_state().clone()
    st[..., :2] = st[:, :1, :2] + 30.0 * torch.randn(B, A, 2, generator=gen)
    st[..., 2] = (torch.rand(B, A, generator=gen) - 0.5) * 12.0                     # beyond (-pi, pi): exercises the wrap
    sim.set_state(st)
    np.savez_compressed(os.path.join(HERE, "relative.npz"), absolute=sim.get_all_agents_absolute().numpy(),
                        relative_excl=sim.get_all_agents_relative(exclude_self=True).numpy(),
                        relative_all=sim.get_all_agents_relative(exclude_self=False).numpy())
    print("relative:", tuple(sim.get_all_agents_relative().shape))


def golden_npc():
    """Replayed NPCs with spawning / despawning (simulator.py:54-124, behavior/replay.py:46-60): ReplayController +
    SpawnController inside Simulator.step; states and masks after every step, collisions against all agents,
    absolute / relative observations and one rendered frame."""
    from torchdrivesim.simulator import SpawnController
    from torchdrivesim.behavior.replay import ReplayController
    gen = torch.Generator().manual_seed(808)
    B, A, Np, T, S = 2, 4, 5, 4, 7
    cfgm = find_map_config("carla_Town01")
    base = road_points(cfgm.road_mesh, B, gen).reshape(B, 1, 1, 2)
    xy = base + 12.0 * torch.randn(B, Np, T, 2, generator=gen)
    replay = torch.cat([xy, torch.rand(B, Np, T, 1, generator=gen) * 6.28, torch.rand(B, Np, T, 1, generator=gen) * 5], -1)
    replay_present = torch.rand(B, Np, T, generator=gen) > 0.25
    # convex exit boundary: a rotated rectangle around the replayed positions, some of which fall outside
    ang = torch.tensor([0.3, -0.5]).reshape(B, 1)
    corners = torch.tensor([[14.0, 10.0], [-14.0, 10.0], [-14.0, -10.0], [14.0, -10.0]])
    rot = torch.stack([torch.cos(ang) * corners[:, 0] - torch.sin(ang) * corners[:, 1],
                       torch.sin(ang) * corners[:, 0] + torch.cos(ang) * corners[:, 1]], -1)
    boundary = rot + base.reshape(B, 1, 2)
    spawn_states = torch.cat([base + 5.0 * torch.randn(B, Np, S, 2, generator=gen), torch.rand(B, Np, S, 2, generator=gen)], -1)
    spawn_masks = torch.rand(B, Np, S, generator=gen) > 0.6
    npc_size = torch.tensor(VEH[:2]).expand(B, Np, 2).clone()
    npc_size[:, ::2] = torch.tensor(PED)
    npc_types = (torch.arange(Np) % 2 == 0).long().expand(B, Np).clone()
    ctrl = ReplayController(npc_size, replay, replay_present, npc_types=npc_types, agent_type_names=["vehicle", "pedestrian"],
                            spawn_controller=SpawnController(boundary, spawn_states, spawn_masks))
    types = torch.zeros(B, A, dtype=torch.long)
    sim, _ = make_sim("carla_Town01", B, A, gen, types=types, type_names=["vehicle", "pedestrian"], with_lights=False,
                      npc_controller=ctrl)
    st = sim.get_state().clone()
    st[..., :2] = base.reshape(B, 1, 2) + 6.0 * torch.randn(B, A, 2, generator=gen)
    sim.set_state(st)
    out = dict(replay=replay.numpy(), replay_present=replay_present.numpy(), boundary=boundary.numpy(),
               spawn_states=spawn_states.numpy(), spawn_masks=spawn_masks.numpy(), npc_size=npc_size.numpy(),
               npc_types=npc_types.numpy(), agent_state0=st.numpy(), agent_size=sim.get_agent_size().numpy(),
               lr=sim.kinematic_model.get_params()["lr"].numpy())
    actions = torch.rand(6, B, A, 2, generator=gen) * 2 - 1
    out["actions"] = actions.numpy()
    states, presents, colls, absol = [], [], [], []
    for t in range(6):
        sim.step(actions[t])
        states.append(sim.get_npc_state().numpy().copy())
        presents.append(sim.get_npc_present_mask().numpy().copy())
        colls.append(sim.compute_collision().numpy().copy())
        absol.append(sim.get_all_agents_absolute().numpy().copy())
    out.update(npc_state=np.stack(states), npc_present=np.stack(presents), collision=np.stack(colls), absolute=np.stack(absol),
               relative=sim.get_all_agents_relative().numpy(), agent_state=sim.get_state().numpy(),
               image=sim.render_egocentric().numpy())
    np.savez_compressed(os.path.join(HERE, "npc.npz"), **out)
    print("npc: present per step", np.stack(presents).sum((1, 2)), "collisions", np.stack(colls).sum(), "image", out["image"].shape)


def golden_goals():
    """Waypoint goals (goals.py:11-217) inside Simulator.step (simulator.py:860-861): collections of M waypoints that
    are achieved within 2 m and progressively enabled; state, masks and gathered waypoints after every step, and the
    egocentric frame with the next two collections drawn."""
    from torchdrivesim.goals import WaypointGoal
    gen = torch.Generator().manual_seed(909)
    B, A, N, M, steps = 2, 3, 4, 3, 8
    sim, _ = make_sim("carla_Town01", B, A, gen, with_lights=False)
    st = sim.get_state().clone()
    st[:, 1:, :2] = st[:, :1, :2] + 10.0 * torch.randn(B, A - 1, 2, generator=gen)
    st[..., 3] = 4.0
    sim.set_state(st)
    # collection n of an agent lies about 1.2 n metres ahead of it, so driving straight reaches them one after the other
    ahead = torch.stack([torch.cos(st[..., 2]), torch.sin(st[..., 2])], -1)
    wp = st[:, :, None, None, :2] + ahead[:, :, None, None] * (1.5 + 1.2 * torch.arange(N).reshape(1, 1, N, 1, 1)) \
        + 1.5 * torch.randn(B, A, N, M, 2, generator=gen)
    mask = torch.rand(B, A, N, M, generator=gen) > 0.3
    mask[0, 0, 1] = False                       # a collection that is all padding
    sim.waypoint_goals = WaypointGoal(wp.clone(), mask.clone())
    out = dict(waypoints=wp.numpy(), mask=mask.numpy(), state0=st.numpy(), size=sim.get_agent_size().numpy(),
               lr=sim.kinematic_model.get_params()["lr"].numpy())
    actions = torch.zeros(steps, B, A, 2)
    actions[..., 1] = 0.1 * (torch.rand(steps, B, A, generator=gen) - 0.5)
    out["actions"] = actions.numpy()
    gs, gm, w1, m1, w2, m2 = [], [], [], [], [], []
    for t in range(steps):
        sim.step(actions[t])
        gs.append(sim.get_waypoints_state().numpy().copy())
        gm.append(sim.waypoint_goals.mask.numpy().copy())
        w1.append(sim.get_waypoints().numpy().copy()); m1.append(sim.get_waypoints_mask().numpy().copy())
        w2.append(sim.get_waypoints(count=2).numpy().copy()); m2.append(sim.get_waypoints_mask(count=2).numpy().copy())
    out.update(goal_state=np.stack(gs), goal_mask=np.stack(gm), wp1=np.stack(w1), m1=np.stack(m1), wp2=np.stack(w2),
               m2=np.stack(m2), agent_state=sim.get_state().numpy(),
               image=sim.render_egocentric(n_subsequent_waypoints=2).numpy())
    np.savez_compressed(os.path.join(HERE, "goals.npz"), **out)
    print("goals: final state", np.stack(gs)[-1, ..., 0].tolist(), "masks left", int(np.stack(gm)[-1].sum()), "of", int(mask.sum()))


def golden_noise():
    """Noisy observations (observation_noise.py:69-132, simulator.py:663-679, 740-746, 784-821): distance-dependent
    sensing noise (the normal deviates are re-drawn from the same seed and saved) and line-of-sight occlusion."""
    from torchdrivesim.observation_noise import StandardSensingObservationNoise, StandardSensingObservationNoiseConfig
    from torchdrivesim.simulator import NPCController
    gen = torch.Generator().manual_seed(1010)
    B, A, Np = 2, 6, 4
    npc_state = torch.cat([40.0 * torch.randn(B, Np, 2, generator=gen), torch.rand(B, Np, 2, generator=gen) * 3], -1)
    npc_size = torch.tensor(VEH[:2]).expand(B, Np, 2).clone()
    npc_present = torch.rand(B, Np, generator=gen) > 0.3
    present = torch.rand(B, A, generator=gen) > 0.2
    sim, _ = make_sim("carla_Town01", B, A, gen, with_lights=False, present=present,
                      npc_controller=NPCController(npc_size, npc_state, npc_present))
    st = sim.get_state().clone()
    st[..., :2] = 35.0 * torch.randn(B, A, 2, generator=gen)
    st[0, 1, :2] = st[0, 0, :2] + 0.3                              # closer than 0.5 m: no noise
    st[1, 2, :2] = st[1, 0, :2] * 0.5 + st[1, 1, :2] * 0.5         # exactly between two agents: occludes them
    sim.set_state(st)
    sim.npc_controller.npc_state[..., :2] += st[:, :1, :2] * 0.0
    sim.observation_noise_model = StandardSensingObservationNoise(StandardSensingObservationNoiseConfig())
    E = A + Np
    out = dict(agent_state=st.numpy(), agent_size=sim.get_agent_size().numpy(), present=present.numpy(),
               npc_state=npc_state.numpy(), npc_size=npc_size.numpy(), npc_present=npc_present.numpy())
    torch.manual_seed(5); out["eps"] = torch.randn(B, A, E, 4).numpy()
    torch.manual_seed(5); out["noisy_state"] = sim.get_noisy_state().numpy()
    out["noisy_present"] = sim.get_noisy_present_mask().numpy()
    out["noisy_size"] = sim.get_noisy_agent_size().numpy()
    torch.manual_seed(5); out["noisy_absolute"] = sim.get_noisy_all_agents_absolute().numpy()
    torch.manual_seed(5); out["noisy_relative"] = sim.get_noisy_all_agents_relative().numpy()
    torch.manual_seed(5); out["noisy_relative_all"] = sim.get_noisy_all_agents_relative(exclude_self=False).numpy()
    np.savez_compressed(os.path.join(HERE, "noise.npz"), **out)
    base = np.concatenate([present.numpy(), npc_present.numpy()], -1)[:, None].repeat(A, 1)
    print("noise: occluded", int((base & ~out["noisy_present"]).sum()), "of", int(base.sum()),
          "deviations", np.unique(np.round(np.abs(out["noisy_state"] - np.concatenate([st.numpy(), npc_state.numpy()], 1)[:, None])
                                           / np.maximum(np.abs(out["eps"]), 1e-9), 2)).tolist()[:8])


def golden_mesh_formats():
    """The reference's two mesh wire formats (mesh.py:238-297, 700-719) written by the reference itself for a 50 m
    window of Town02: BirdviewMesh.pickle and BirdviewMesh.save (json), plus the arrays they hold."""
    import dataclasses
    mesh = find_map_config("carla_Town02").road_mesh
    v, f, vc = mesh.verts[0], mesh.faces[0], mesh.vert_category[0]
    centre = v[f[100]].mean(0)
    keep = ((v[f] - centre).abs().amax(-1) < 25.0).all(-1)                  # faces entirely inside the window
    fk = f[keep]
    used, inv = torch.unique(fk, return_inverse=True)
    small = dataclasses.replace(mesh, verts=v[used][None], faces=inv[None], vert_category=vc[used][None])
    os.makedirs(os.path.join(HERE, "maps"), exist_ok=True)
    small.pickle(os.path.join(HERE, "maps", "town02_window.pkl"))
    small.save(os.path.join(HERE, "maps", "town02_window_mesh.json"))
    np.savez_compressed(os.path.join(HERE, "maps", "town02_window_arrays.npz"), verts=small.verts[0].numpy(),
                        faces=small.faces[0].numpy(), vert_category=small.vert_category[0].numpy(),
                        categories=np.array(small.categories))
    print("mesh formats:", tuple(small.verts.shape), tuple(small.faces.shape), small.categories)


def golden_light_schedule():
    """TrafficLightController (traffic_lights.py:159-301) of carla_Town02 ticked 400 times by 0.1 s from a fixed
    state: group states, remaining times and the state index of every traffic light of the map after every tick."""
    import json
    import shutil
    from torchdrivesim.traffic_lights import current_light_state_tensor_from_controller
    cfgm = find_map_config("carla_Town02")
    ctrl = cfgm.traffic_light_controller
    ids = [s.actor_id for s in cfgm.stoplines if s.agent_type == "traffic_light"]
    start = [(i % len(f.states), 0.37 * (i + 1)) for i, f in enumerate(ctrl.traffic_fsms)]
    ctrl.set_to(start)
    machine, remaining, lights = [], [], []
    for t in range(400):
        machine.append(list(ctrl.state_per_machine)); remaining.append(list(ctrl.time_remaining))
        lights.append(current_light_state_tensor_from_controller(ctrl, ids).numpy())
        ctrl.tick(0.1)
    shutil.copy(cfgm.traffic_light_controller_path, os.path.join(HERE, "maps", "carla_Town02_traffic_light_controller.json"))
    os.chmod(os.path.join(HERE, "maps", "carla_Town02_traffic_light_controller.json"), 0o644)
    np.savez_compressed(os.path.join(HERE, "light_schedule.npz"), ids=np.array(ids), start=np.array(start), machine=np.array(machine),
                        remaining=np.array(remaining), lights=np.stack(lights))
    print("light schedule:", len(ids), "lights,", len(ctrl.traffic_fsms), "groups, states seen", np.unique(np.stack(lights)).tolist())


def golden_traffic():
    """TrafficLightControl.compute_violation / Simulator.compute_traffic_lights_violations (traffic_controls.py:152-178,
    simulator.py:1046-1062): agents placed on and around the stop lines of Town01, random light states."""
    gen = torch.Generator().manual_seed(505)
    B, A = 4, 64
    present = torch.rand(B, A, generator=gen) > 0.15
    sim, cfgm = make_sim("carla_Town01", B, A, gen, with_lights=True, present=present)
    tl = sim.get_traffic_controls()["traffic_light"]
    L = tl.pos.shape[1]
    tl.set_state(torch.where(torch.rand(B, L, generator=gen) < 0.5, torch.zeros(B, L, dtype=torch.long), tl.state))
    st = sim.get_state().clone()
    pick = torch.randint(0, L, (B, A), generator=gen)
    centre = torch.gather(tl.pos[..., :2], 1, pick.unsqueeze(-1).expand(-1, -1, 2))
    st[..., :2] = centre + 2.0 * torch.randn(B, A, 2, generator=gen)        # within a car length of a stop line
    st[:, ::5, :2] = centre[:, ::5]                                         # some exactly on it
    sim.set_state(st)
    viol = sim.compute_traffic_lights_violations()
    box = torch.cat([st[..., :2], sim.get_agent_size()[..., :2], st[..., 2:3]], -1)
    raw = tl.compute_violation(box)
    np.savez_compressed(os.path.join(HERE, "traffic.npz"), agent_box=box.numpy(), present=present.numpy(),
                        tl_corners=tl.corners.numpy(), tl_state=tl.state.numpy(), red_index=tl.allowed_states.index("red"),
                        rear_factor=tl.violation_rear_factor, violation_raw=raw.numpy(),
                        violation=np.asarray(viol.numpy() != 0))
    print("traffic: violations", int(raw.sum()), "of", B * A, "red lights", int((tl.state == 0).sum()), "of", B * L)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kinematic", "collision", "offroad", "render", "traffic", "waypoints", "relative", "npc", "goals", "noise", "mesh_formats", "light_schedule", "custom_colors"]
    for w in which:
        globals()["golden_" + w]()
