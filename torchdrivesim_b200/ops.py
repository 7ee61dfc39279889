"""Functional ops over libtds_b200.so with autograd wiring (kinematics, discs collision, offroad).

Every function takes CUDA tensors in the reference's layouts and launches on the current torch
stream.  Nothing here has a CPU implementation; `_lib.ptr` raises on CPU tensors.
"""
import ctypes
import math
from typing import Optional

import torch

from . import _lib
from .maps import MapSet

_F32_HALF_PI = 1.5707963705062866  # float32(pi / 2), the reference's _normalization_factor (kinematic.py:420)


def kinematic_params(dt: float = 0.1, max_acceleration: float = 5.0, max_steering: float = math.pi / 2,
                     max_yaw_rate: float = math.pi / 2, left_handed: bool = False, max_dx: float = 20.0,
                     max_dpsi: float = 10 * math.pi, max_dv: float = 5.0) -> "_lib.KinematicParams":
    return _lib.KinematicParams(dt, max_acceleration, max_steering, max_yaw_rate, 1 if left_handed else 0,
                                max_dx, max_dpsi, max_dv)


# ------------------------------------------------------------------------------------ kinematics
class _KinematicStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, action, lr, model, uniform_model, params):
        lib = _lib.load()
        shape = state.shape
        s = _lib.as_f32(state).reshape(-1, 4)
        adim = action.shape[-1]
        if adim not in (2, 4):
            raise _lib.TdsError("kinematic_step: the action must have 2 or 4 values per agent")
        a = _lib.as_f32(action).reshape(-1, adim)
        l = None if lr is None else _lib.as_f32(lr).reshape(-1)
        m = None if model is None else _lib.as_i32(model).reshape(-1)
        n = s.shape[0]
        if a.shape[0] != n or (l is not None and l.shape[0] != n) or (m is not None and m.shape[0] != n):
            raise _lib.TdsError("kinematic_step: state, action, lr and model must agree on the batch shape")
        out = torch.empty_like(s)
        _lib.check(lib.tds_kinematic_step_fwd(_lib.ptr(s), _lib.ptr(a), adim, _lib.ptr(l), _lib.ptr(m), uniform_model, n,
                                              ctypes.byref(params), _lib.ptr(out), _lib.stream_ptr(s.device)))
        ctx.save_for_backward(s, a, l, m)
        ctx.meta = (uniform_model, params, shape, action.shape, None if lr is None else lr.shape)
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        s, a, l, m = ctx.saved_tensors
        uniform_model, params, sshape, ashape, lshape = ctx.meta
        g = _lib.as_f32(grad_out).reshape(-1, 4)
        n = s.shape[0]
        gs = torch.empty_like(s)
        ga = torch.empty_like(a)
        gl = None if l is None else torch.empty_like(l)
        _lib.check(lib.tds_kinematic_step_bwd(_lib.ptr(s), _lib.ptr(a), a.shape[-1], _lib.ptr(l), _lib.ptr(m), uniform_model, n,
                                              ctypes.byref(params), _lib.ptr(g), _lib.ptr(gs), _lib.ptr(ga), _lib.ptr(gl),
                                              _lib.stream_ptr(s.device)))
        return gs.reshape(sshape), ga.reshape(ashape), (None if gl is None else gl.reshape(lshape)), None, None, None


def kinematic_step(state: torch.Tensor, action: torch.Tensor, lr: Optional[torch.Tensor],
                   model: Optional[torch.Tensor] = None, uniform_model: int = _lib.MODEL_BICYCLE,
                   params: Optional["_lib.KinematicParams"] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """state [...,4], action [...,2|4], lr [...], model [...] int (or None: `uniform_model`) -> new state.
    `out` (contiguous float32, the state's shape; may alias `state`): written directly, no autograd tape."""
    params = params or kinematic_params()
    if out is None:
        return _KinematicStep.apply(state, action, lr, model, uniform_model, params)
    if out.dtype != torch.float32 or not out.is_contiguous() or out.shape != state.shape:
        raise _lib.TdsError("kinematic_step: `out` must be a contiguous float32 tensor of the state's shape")
    lib = _lib.load()
    s = _lib.as_f32(state).reshape(-1, 4)
    adim = action.shape[-1]
    a = _lib.as_f32(action).reshape(-1, adim)
    l = None if lr is None else _lib.as_f32(lr).reshape(-1)
    m = None if model is None else _lib.as_i32(model).reshape(-1)
    _lib.check(lib.tds_kinematic_step_fwd(_lib.ptr(s), _lib.ptr(a), adim, _lib.ptr(l), _lib.ptr(m), uniform_model, s.shape[0],
                                          ctypes.byref(params), _lib.ptr(out), _lib.stream_ptr(s.device)))
    return out


# ------------------------------------------------------------------------------------ collisions
class _Pairwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, box1, box2, metric):
        lib = _lib.load()
        b1 = _lib.as_f32(box1).reshape(-1, 5)
        b2 = _lib.as_f32(box2).reshape(-1, 5)
        out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
        _lib.check(lib.tds_collision_pairwise_fwd(_lib.ptr(b1), _lib.ptr(b2), b1.shape[0], metric,
                                                  _lib.ptr(out), _lib.stream_ptr(b1.device)))
        ctx.save_for_backward(b1, b2)
        ctx.meta = (box1.shape, box2.shape, metric)
        return out.reshape(box1.shape[:-1])

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        b1, b2 = ctx.saved_tensors
        g = _lib.as_f32(grad_out).reshape(-1)
        g1, g2 = torch.empty_like(b1), torch.empty_like(b2)
        _lib.check(lib.tds_collision_pairwise_bwd(_lib.ptr(b1), _lib.ptr(b2), b1.shape[0], ctx.meta[2], _lib.ptr(g),
                                                  _lib.ptr(g1), _lib.ptr(g2), _lib.stream_ptr(b1.device)))
        return g1.reshape(ctx.meta[0]), g2.reshape(ctx.meta[1]), None


def collision_pairwise(box1: torch.Tensor, box2: torch.Tensor, metric: int = _lib.METRIC_DISCS) -> torch.Tensor:
    """Element-wise overlap of box1[...,5] and box2[...,5] (same shape) -> [...], differentiable."""
    if box1.shape != box2.shape or box1.shape[-1] != 5:
        raise _lib.TdsError("collision_pairwise: boxes must have identical shape [...,5]")
    return _Pairwise.apply(box1, box2, metric)


class _AllPairs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ego_box, all_box, mask, metric, ego_is_prefix):
        lib = _lib.load()
        e = _lib.as_f32(ego_box)
        a = _lib.as_f32(all_box)
        m = _lib.as_u8(mask)
        B, A, N = e.shape[0], e.shape[1], a.shape[1]
        out = torch.empty(B, A, dtype=torch.float32, device=e.device)       # the kernel writes every row
        arg = torch.empty(B, A, dtype=torch.int32, device=e.device)
        _lib.check(lib.tds_collision_allpairs_fwd(_lib.ptr(e), _lib.ptr(a), _lib.ptr(m), B, A, N, metric,
                                                  1 if ego_is_prefix else 0, _lib.ptr(out), _lib.ptr(arg),
                                                  _lib.stream_ptr(e.device)))
        ctx.save_for_backward(e, a, m, arg)
        ctx.meta = (metric, 1 if ego_is_prefix else 0)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        e, a, m, arg = ctx.saved_tensors
        B, A, N = e.shape[0], e.shape[1], a.shape[1]
        g = _lib.as_f32(grad_out)
        ge = torch.zeros_like(e)
        ga = torch.zeros_like(a)
        _lib.check(lib.tds_collision_allpairs_bwd(_lib.ptr(e), _lib.ptr(a), _lib.ptr(m), B, A, N, ctx.meta[0], ctx.meta[1],
                                                  _lib.ptr(g), _lib.ptr(arg), _lib.ptr(ge), _lib.ptr(ga),
                                                  _lib.stream_ptr(e.device)))
        return ge, ga, None, None, None


def collision_allpairs(ego_box: torch.Tensor, all_box: torch.Tensor, mask: torch.Tensor,
                       metric: int = _lib.METRIC_DISCS, ego_is_prefix: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused Simulator.compute_collision: ego_box [B,A,5], all_box [B,N,5], mask [B,N] -> [B,A], differentiable.
    `out` (contiguous float32 [B,A]): written directly, no autograd tape."""
    if ego_box.dim() != 3 or all_box.dim() != 3 or ego_box.shape[-1] != 5 or all_box.shape[-1] != 5:
        raise _lib.TdsError("collision_allpairs: expected ego_box [B,A,5] and all_box [B,N,5]")
    if ego_box.shape[0] != all_box.shape[0] or tuple(mask.shape) != tuple(all_box.shape[:2]):
        raise _lib.TdsError("collision_allpairs: batch / mask shape mismatch")
    if metric not in (_lib.METRIC_DISCS, _lib.METRIC_IOU):
        raise _lib.TdsError(f"collision_allpairs: unknown metric {metric}")
    if out is None:
        return _AllPairs.apply(ego_box, all_box, mask, metric, ego_is_prefix)
    e, a, m = _lib.as_f32(ego_box), _lib.as_f32(all_box), _lib.as_u8(mask)
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != tuple(e.shape[:2]):
        raise _lib.TdsError("collision_allpairs: `out` must be a contiguous float32 [B,A] tensor")
    _lib.check(_lib.load().tds_collision_allpairs_fwd(_lib.ptr(e), _lib.ptr(a), _lib.ptr(m), e.shape[0], e.shape[1], a.shape[1], metric,
                                                      1 if ego_is_prefix else 0, _lib.ptr(out), None, _lib.stream_ptr(e.device)))
    return out


class _AgentBoxes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, size):
        lib = _lib.load()
        s = _lib.as_f32(state)
        z = _lib.as_f32(size[..., :2])
        n = s[..., 0].numel()
        box = torch.empty(s.shape[:-1] + (5,), dtype=torch.float32, device=s.device)
        _lib.check(lib.tds_agent_boxes(_lib.ptr(s), _lib.ptr(z), n, _lib.ptr(box), None, None, _lib.stream_ptr(s.device)))
        ctx.meta = (state.shape, size.shape)
        return box

    @staticmethod
    def backward(ctx, g):
        sshape, zshape = ctx.meta
        gs = torch.cat([g[..., 0:2], g[..., 4:5], torch.zeros_like(g[..., 0:1])], dim=-1)
        gz = g[..., 2:4]
        if zshape[-1] > 2:
            gz = torch.cat([gz, torch.zeros(gz.shape[:-1] + (zshape[-1] - 2,), dtype=gz.dtype, device=gz.device)], dim=-1)
        return gs.reshape(sshape), gz.reshape(zshape)


def agent_boxes(state: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """state [...,4] (x, y, psi, v), size [...,2+] (length, width) -> boxes [...,5] (x, y, length, width, psi): the layout of
    compute_collision (simulator.py:1161-1170) in one launch instead of a torch.cat; differentiable."""
    if state.shape[-1] != 4 or size.shape[-1] < 2 or state.shape[:-1] != size.shape[:-1]:
        raise _lib.TdsError("agent_boxes: expected state [...,4] and size [...,2]")
    return _AgentBoxes.apply(state, size)


def egocentric_cameras(state: torch.Tensor):
    """state [...,4] -> (xy [...,2] contiguous, (sin psi, cos psi) [...,2]): the camera of every agent in one launch, the
    heading evaluated like every other heading of the path (float64, rounded) - render_egocentric, simulator.py:961,
    1017.  Not differentiable."""
    lib = _lib.load()
    s = _lib.as_f32(state)
    xy = torch.empty(s.shape[:-1] + (2,), dtype=torch.float32, device=s.device)
    sc = torch.empty_like(xy)
    _lib.check(lib.tds_agent_boxes(_lib.ptr(s), None, s[..., 0].numel(), None, _lib.ptr(sc), _lib.ptr(xy), _lib.stream_ptr(s.device)))
    return xy, sc


def heading_sincos(state: torch.Tensor) -> torch.Tensor:
    return egocentric_cameras(state)[1]


# ------------------------------------------------------------------------------------ traffic lights
def traffic_light_violation(agent_box: torch.Tensor, tl_corners: torch.Tensor, tl_state: torch.Tensor, red_state: int,
                            rear_factor: float = 0.1, present: Optional[torch.Tensor] = None) -> torch.Tensor:
    """agent_box [B,A,5], tl_corners [B,L,4,2], tl_state [B,L] -> bool [B,A]
    (TrafficLightControl.compute_violation, traffic_controls.py:152-178; not differentiable, as the reference)."""
    lib = _lib.load()
    box = _lib.as_f32(agent_box)
    if box.dim() != 3 or box.shape[-1] != 5:
        raise _lib.TdsError("traffic_light_violation: agent_box must be [B,A,5]")
    B, A = box.shape[0], box.shape[1]
    L = tl_corners.shape[1]
    if tl_corners.shape[0] != B or tuple(tl_corners.shape[2:]) != (4, 2) or tuple(tl_state.shape) != (B, L):
        raise _lib.TdsError("traffic_light_violation: tl_corners must be [B,L,4,2] and tl_state [B,L]")
    out = torch.zeros(B, A, dtype=torch.uint8, device=box.device)
    cor = _lib.as_f32(tl_corners) if L > 0 else None
    st = _lib.as_i32(tl_state.to(box.device)) if L > 0 else None
    p = None if present is None else _lib.as_u8(present)
    _lib.check(lib.tds_traffic_light_violation(_lib.ptr(box), _lib.ptr(cor), _lib.ptr(st), _lib.ptr(p), B, A, L,
                                               int(red_state), float(rear_factor), _lib.ptr(out), _lib.stream_ptr(box.device)))
    return out.view(torch.bool)


def agents_relative(absolute: torch.Tensor, n_agents: Optional[int] = None, exclude_self: bool = True) -> torch.Tensor:
    """absolute [B,N,6] (x, y, psi, length, width, present) -> [B,A,N(-1),6]: every agent in the frame of each of the
    first `n_agents` agents (Simulator.get_all_agents_relative, simulator.py:748-781).  With absolute [B,A,N,6] (what
    each agent perceives, get_noisy_all_agents_absolute) row i is relative to its own entry [b,i,i]
    (get_noisy_all_agents_relative, simulator.py:784-821).  The self entry is removed on the device without the
    reference's synchronising boolean-mask index.  Not differentiable."""
    lib = _lib.load()
    a = _lib.as_f32(absolute)
    per_origin = a.dim() == 4
    if a.dim() not in (3, 4) or a.shape[-1] != 6:
        raise _lib.TdsError("agents_relative: absolute must be [B,N,6] or [B,A,N,6]")
    B, N = a.shape[0], a.shape[-2]
    A = a.shape[1] if per_origin else (N if n_agents is None else int(n_agents))
    if not 0 <= A <= N:
        raise _lib.TdsError("agents_relative: n_agents must be in [0, N]")
    M = max(N - 1, 0) if exclude_self else N
    out = torch.empty(B, A, M, 6, dtype=torch.float32, device=a.device)
    _lib.check(lib.tds_agents_relative(_lib.ptr(a), B, A, N, 1 if exclude_self else 0, 1 if per_origin else 0, _lib.ptr(out),
                                       _lib.stream_ptr(a.device)))
    return out


def sensing_noise(all_state: torch.Tensor, n_agents: int, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """all_state [B,N,4] -> [B,A,N,4]: the states the first `n_agents` agents perceive, with noise growing with the
    distance (StandardSensingObservationNoise.get_noisy_state, observation_noise.py:74-89).  `eps` [B,A,N,4] are the
    standard normal deviates (drawn on the device when omitted)."""
    lib = _lib.load()
    s = _lib.as_f32(all_state)
    if s.dim() != 3 or s.shape[-1] != 4:
        raise _lib.TdsError("sensing_noise: all_state must be [B,N,4]")
    B, N = s.shape[0], s.shape[1]
    A = int(n_agents)
    if eps is None:
        eps = torch.randn(B, A, N, 4, dtype=torch.float32, device=s.device)
    eps = _lib.as_f32(eps)
    if tuple(eps.shape) != (B, A, N, 4):
        raise _lib.TdsError("sensing_noise: eps must be [B,A,N,4]")
    out = torch.empty(B, A, N, 4, dtype=torch.float32, device=s.device)
    _lib.check(lib.tds_sensing_noise(_lib.ptr(s), _lib.ptr(eps), B, A, N, _lib.ptr(out), _lib.stream_ptr(s.device)))
    return out


def sensing_occlusion(all_state: torch.Tensor, all_size: torch.Tensor, base_mask: torch.Tensor, n_agents: int) -> torch.Tensor:
    """all_state [B,N,4], all_size [B,N,2], base_mask [B,N] -> bool [B,A,N]: present and in the line of sight of each of
    the first `n_agents` agents (StandardSensingObservationNoise.get_noisy_present_mask, observation_noise.py:91-132)."""
    lib = _lib.load()
    s, z, m = _lib.as_f32(all_state), _lib.as_f32(all_size), _lib.as_u8(base_mask)
    B, N = s.shape[0], s.shape[1]
    if s.dim() != 3 or s.shape[-1] != 4 or tuple(z.shape) != (B, N, 2) or tuple(m.shape) != (B, N):
        raise _lib.TdsError("sensing_occlusion: expected all_state [B,N,4], all_size [B,N,2], base_mask [B,N]")
    A = int(n_agents)
    out = torch.empty(B, A, N, dtype=torch.uint8, device=s.device)
    _lib.check(lib.tds_sensing_occlusion(_lib.ptr(s), _lib.ptr(z), _lib.ptr(m), B, A, N, _lib.ptr(out), _lib.stream_ptr(s.device)))
    return out.view(torch.bool)


def infraction_metrics(collision: torch.Tensor, offroad: torch.Tensor, present: Optional[torch.Tensor] = None,
                       acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[6] float64, accumulated into `acc` when given: collision sum, offroad sum, colliding agents, offroad agents,
    present agents, agent slots (distributed.METRIC_NAMES) - the vector a multi-GPU job all-reduces.  One launch."""
    lib = _lib.load()
    c, o = _lib.as_f32(collision), _lib.as_f32(offroad)
    if c.shape != o.shape:
        raise _lib.TdsError("infraction_metrics: collision and offroad must have the same shape")
    p = None if present is None else _lib.as_u8(present)
    if p is not None and p.shape != c.shape:
        raise _lib.TdsError("infraction_metrics: present must have the shape of collision")
    if acc is None:
        acc = torch.zeros(6, dtype=torch.float64, device=c.device)
    elif acc.dtype != torch.float64 or acc.numel() != 6 or not acc.is_contiguous():
        raise _lib.TdsError("infraction_metrics: acc must be a contiguous float64 tensor of 6 elements")
    _lib.check(lib.tds_infraction_metrics(_lib.ptr(c), _lib.ptr(o), _lib.ptr(p), c.numel(), _lib.ptr(acc), _lib.stream_ptr(c.device)))
    return acc


# ------------------------------------------------------------------------------------ offroad
class _Offroad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, lenwid, present, mapset, threshold):
        lib = _lib.load()
        s = _lib.as_f32(state)
        lw = _lib.as_f32(lenwid)
        p = None if present is None else _lib.as_u8(present)
        B, A = s.shape[0], s.shape[1]
        handles, n_maps = mapset.handles(s.device)
        env_map = mapset.env_map_on(s.device)
        out = torch.empty(B, A, dtype=torch.float32, device=s.device)       # the kernel writes every agent
        face = torch.empty(B, A, 4, dtype=torch.int32, device=s.device)
        _lib.check(lib.tds_offroad_fwd(handles, n_maps, _lib.ptr(env_map), _lib.ptr(s), _lib.ptr(lw), _lib.ptr(p), B, A,
                                       float(threshold), _lib.ptr(out), _lib.ptr(face), _lib.stream_ptr(s.device)))
        ctx.save_for_backward(s, lw, p, face)
        ctx.meta = (mapset, float(threshold), state.shape, lenwid.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        s, lw, p, face = ctx.saved_tensors
        mapset, threshold, sshape, lshape = ctx.meta
        B, A = s.shape[0], s.shape[1]
        handles, n_maps = mapset.handles(s.device)
        env_map = mapset.env_map_on(s.device)
        g = _lib.as_f32(grad_out)
        gs = torch.zeros_like(s)
        gl = torch.zeros_like(lw)
        _lib.check(lib.tds_offroad_bwd(handles, n_maps, _lib.ptr(env_map), _lib.ptr(s), _lib.ptr(lw), _lib.ptr(p), B, A,
                                       threshold, _lib.ptr(face), _lib.ptr(g), _lib.ptr(gs), _lib.ptr(gl),
                                       _lib.stream_ptr(s.device)))
        return gs.reshape(sshape), gl.reshape(lshape), None, None, None


def offroad(state: torch.Tensor, lenwid: torch.Tensor, mapset: MapSet, threshold: float = 0.0,
            present: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """state [B,A,4], lenwid [B,A,2] -> [B,A] sum over corners of thresholded squared distance to the map.
    `out` (contiguous float32 [B,A]): written directly, no autograd tape."""
    if state.dim() != 3 or state.shape[-1] != 4:
        raise _lib.TdsError("offroad: expected state [B,A,4]")
    if lenwid.dim() == 2:
        lenwid = lenwid.unsqueeze(-2).expand(lenwid.shape[0], state.shape[1], lenwid.shape[1])
    if tuple(lenwid.shape) != (state.shape[0], state.shape[1], 2):
        raise _lib.TdsError("offroad: expected lenwid [B,A,2] or [B,2]")
    if out is None:
        return _Offroad.apply(state, lenwid, present, mapset, threshold)
    s, lw = _lib.as_f32(state), _lib.as_f32(lenwid)
    p = None if present is None else _lib.as_u8(present)
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != tuple(s.shape[:2]):
        raise _lib.TdsError("offroad: `out` must be a contiguous float32 [B,A] tensor")
    handles, n_maps = mapset.handles(s.device)
    _lib.check(_lib.load().tds_offroad_fwd(handles, n_maps, _lib.ptr(mapset.env_map_on(s.device)), _lib.ptr(s), _lib.ptr(lw), _lib.ptr(p),
                                           s.shape[0], s.shape[1], float(threshold), _lib.ptr(out), None, _lib.stream_ptr(s.device)))
    return out


# ------------------------------------------------------------------------------------ raster
def raster_birdview(mapset: MapSet, palette: "_lib.Palette", cam_xy: torch.Tensor, cam_sc: torch.Tensor,
                    agent_state: Optional[torch.Tensor], agent_size: Optional[torch.Tensor],
                    agent_type: Optional[torch.Tensor], present: Optional[torch.Tensor],
                    tl_corners: Optional[torch.Tensor], tl_state: Optional[torch.Tensor],
                    rect_corners: Optional[torch.Tensor], rect_class: Optional[torch.Tensor],
                    res: int, fov: float, out: Optional[torch.Tensor] = None,
                    workspace: Optional[torch.Tensor] = None, cam_tris: Optional[torch.Tensor] = None,
                    cam_tri_class: Optional[torch.Tensor] = None, image_format: int = _lib.IMAGE_F32,
                    agent_class: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cam_xy, cam_sc [B,Nc,2] -> images [B,Nc,3,res,res] float32 in [0,255] (not differentiable, like
    the cv2 backend).  present is [B,N] or [B,Nc,N]; cam_tris [B,Nc,Tc,3,2] + cam_tri_class [B,Nc,Tc] are
    triangles seen by one camera each (waypoint discs).  image_format: IMAGE_F32 (the reference's dtype),
    IMAGE_U8 (the same values as uint8 [B,Nc,3,res,res]) or IMAGE_RANK (uint8 [B,Nc,res,res] draw ranks, see
    `raster_rank_table`).  agent_class [B,Nc,N] uint8: palette class of each agent's rectangle per camera (custom colours)."""
    lib = _lib.load()
    dev = cam_xy.device
    cxy, csc = _lib.as_f32(cam_xy), _lib.as_f32(cam_sc)
    B, Nc = cxy.shape[0], cxy.shape[1]
    N = 0 if agent_state is None else agent_state.shape[-2]
    L = 0 if tl_corners is None else tl_corners.shape[1]
    R = 0 if rect_corners is None else rect_corners.shape[1]
    ast = None if N == 0 else _lib.as_f32(agent_state[..., :4])
    asz = None if N == 0 else _lib.as_f32(agent_size)
    aty = None if (N == 0 or agent_type is None) else _lib.as_i32(agent_type)
    per_cam = 0
    pr = None
    if N > 0 and present is not None:
        pr = _lib.as_u8(present)
        per_cam = 1 if pr.dim() == 3 else 0
        if per_cam and tuple(pr.shape) != (B, Nc, N) or (not per_cam and tuple(pr.shape) != (B, N)):
            raise _lib.TdsError("raster: present mask must be [B,N] or [B,Nc,N]")
    tlc = None if L == 0 else _lib.as_f32(tl_corners)
    tls = None if L == 0 else _lib.as_i32(tl_state)
    rc = None if R == 0 else _lib.as_f32(rect_corners)
    rcl = None if R == 0 else _lib.as_i32(rect_class)
    if image_format not in (_lib.IMAGE_F32, _lib.IMAGE_U8, _lib.IMAGE_RANK):
        raise _lib.TdsError(f"raster: unknown image format {image_format}")
    oshape = (B, Nc, res, res) if image_format == _lib.IMAGE_RANK else (B, Nc, 3, res, res)
    odtype = torch.float32 if image_format == _lib.IMAGE_F32 else torch.uint8
    if out is None:
        out = torch.empty(oshape, dtype=odtype, device=dev)
    elif tuple(out.shape) != oshape or out.dtype != odtype or not out.is_contiguous():
        raise _lib.TdsError(f"raster: `out` must be a contiguous {odtype} {list(oshape)} tensor")
    need = lib.tds_raster_workspace_bytes(B, N, L, R)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    Tc, ctr, ccl = 0, None, None
    if cam_tris is not None and cam_tris.shape[2] > 0:
        Tc = cam_tris.shape[2]
        if tuple(cam_tris.shape) != (B, Nc, Tc, 3, 2) or cam_tri_class is None or tuple(cam_tri_class.shape) != (B, Nc, Tc):
            raise _lib.TdsError("raster: cam_tris must be [B,Nc,Tc,3,2] and cam_tri_class [B,Nc,Tc]")
        ctr, ccl = _lib.as_f32(cam_tris), _lib.as_i32(cam_tri_class)
    acl = None
    if agent_class is not None and N > 0:
        if tuple(agent_class.shape) != (B, Nc, N) or agent_class.dtype != torch.uint8:
            raise _lib.TdsError("raster: agent_class must be uint8 [B,Nc,N]")
        acl = agent_class.contiguous()
    handles, n_maps = mapset.handles(dev)
    env_map = mapset.env_map_on(dev)
    _lib.check(lib.tds_raster_birdview_fmt(handles, n_maps, _lib.ptr(env_map), B, Nc, N, _lib.ptr(cxy), _lib.ptr(csc),
                                           _lib.ptr(ast), _lib.ptr(asz), _lib.ptr(aty), _lib.ptr(pr), per_cam,
                                           _lib.ptr(tlc), _lib.ptr(tls), L, _lib.ptr(rc), _lib.ptr(rcl), R,
                                           _lib.ptr(ctr), _lib.ptr(ccl), Tc, ctypes.byref(palette), float(2.0 / fov), int(res),
                                           _lib.ptr(acl), int(image_format), _lib.ptr(out), _lib.ptr(workspace),
                                           _lib.stream_ptr(dev)))
    return out


def raster_rank_table(palette: "_lib.Palette"):
    """(rgb uint8 [K+1,3], class ids int32 [K+1]) of the draw ranks an IMAGE_RANK image holds: row 0 is the
    background, row k the k-th class in draw order.  rgb[rank_image.long()] is the IMAGE_U8 picture (channels last)."""
    lib = _lib.load()
    rgb = (ctypes.c_uint8 * (3 * (_lib.MAX_CLASSES + 1)))()
    cls = (ctypes.c_int32 * (_lib.MAX_CLASSES + 1))()
    n = lib.tds_raster_rank_table(ctypes.byref(palette), rgb, cls)
    return (torch.tensor(list(rgb), dtype=torch.uint8).reshape(-1, 3)[:n].clone(),
            torch.tensor(list(cls), dtype=torch.int32)[:n].clone())
