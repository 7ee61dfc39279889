"""`torch.ops.tds_b200.*`: the pure-tensor entry points of libtds_b200.so registered as PyTorch custom operators
(torch.library), with their hand-written backward kernels wired in through `register_autograd` and shape functions
for fake tensors - so that they are first-class ops for `torch.compile` / `torch.export` graphs, not only Python
`autograd.Function`s.  The ops that take a static map (offroad, raster) stay in `ops.py`: a map is a library handle,
not a tensor.

    torch.ops.tds_b200.kinematic_step(state, action, lr, model, uniform_model, dt, left_handed) -> state'
    torch.ops.tds_b200.collision_pairwise(box1, box2, metric) -> overlap          (metric 0 = discs, 1 = IoU)
    torch.ops.tds_b200.collision_allpairs(ego_box, all_box, mask, metric, ego_is_prefix) -> (loss, argmax)
    torch.ops.tds_b200.agent_boxes(state, size) -> boxes

Reference call sites: kinematic.py:462-523, infractions.py:503-545 / 307-324, simulator.py:1161-1194.
"""
import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops

_NS = "tds_b200"


def _params(dt: float, left_handed: bool):
    return ops.kinematic_params(dt=dt, left_handed=left_handed)


# ------------------------------------------------------------------------------------ kinematic step
@torch.library.custom_op(f"{_NS}::kinematic_step", mutates_args=())
def kinematic_step(state: Tensor, action: Tensor, lr: Tensor, model: Optional[Tensor], uniform_model: int, dt: float,
                   left_handed: bool) -> Tensor:
    lib = _lib.load()
    s = _lib.as_f32(state).reshape(-1, 4)
    a = _lib.as_f32(action).reshape(-1, action.shape[-1])
    l = _lib.as_f32(lr).reshape(-1)
    m = None if model is None else _lib.as_i32(model).reshape(-1)
    out = torch.empty_like(s)
    p = _params(dt, left_handed)
    _lib.check(lib.tds_kinematic_step_fwd(_lib.ptr(s), _lib.ptr(a), a.shape[-1], _lib.ptr(l), _lib.ptr(m), uniform_model, s.shape[0],
                                          ctypes.byref(p), _lib.ptr(out), _lib.stream_ptr(s.device)))
    return out.reshape(state.shape)


@kinematic_step.register_fake
def _(state, action, lr, model, uniform_model, dt, left_handed):
    return torch.empty_like(state, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::kinematic_step_backward", mutates_args=())
def kinematic_step_backward(grad_out: Tensor, state: Tensor, action: Tensor, lr: Tensor, model: Optional[Tensor], uniform_model: int,
                            dt: float, left_handed: bool) -> Tuple[Tensor, Tensor, Tensor]:
    lib = _lib.load()
    s = _lib.as_f32(state).reshape(-1, 4)
    a = _lib.as_f32(action).reshape(-1, action.shape[-1])
    l = _lib.as_f32(lr).reshape(-1)
    m = None if model is None else _lib.as_i32(model).reshape(-1)
    g = _lib.as_f32(grad_out).reshape(-1, 4)
    gs, ga, gl = torch.empty_like(s), torch.empty_like(a), torch.empty_like(l)
    p = _params(dt, left_handed)
    _lib.check(lib.tds_kinematic_step_bwd(_lib.ptr(s), _lib.ptr(a), a.shape[-1], _lib.ptr(l), _lib.ptr(m), uniform_model, s.shape[0],
                                          ctypes.byref(p), _lib.ptr(g), _lib.ptr(gs), _lib.ptr(ga), _lib.ptr(gl), _lib.stream_ptr(s.device)))
    return gs.reshape(state.shape), ga.reshape(action.shape), gl.reshape(lr.shape)


@kinematic_step_backward.register_fake
def _(grad_out, state, action, lr, model, uniform_model, dt, left_handed):
    return (torch.empty_like(state, dtype=torch.float32), torch.empty_like(action, dtype=torch.float32),
            torch.empty_like(lr, dtype=torch.float32))


def _kin_setup(ctx, inputs, output):
    state, action, lr, model, uniform_model, dt, left_handed = inputs
    ctx.save_for_backward(state, action, lr, model)
    ctx.meta = (uniform_model, dt, left_handed)


def _kin_backward(ctx, grad_out):
    state, action, lr, model = ctx.saved_tensors
    gs, ga, gl = kinematic_step_backward(grad_out.contiguous(), state, action, lr, model, *ctx.meta)
    return gs, ga, gl, None, None, None, None


kinematic_step.register_autograd(_kin_backward, setup_context=_kin_setup)


# ------------------------------------------------------------------------------------ collisions
@torch.library.custom_op(f"{_NS}::collision_pairwise", mutates_args=())
def collision_pairwise(box1: Tensor, box2: Tensor, metric: int) -> Tensor:
    lib = _lib.load()
    b1, b2 = _lib.as_f32(box1).reshape(-1, 5), _lib.as_f32(box2).reshape(-1, 5)
    out = torch.empty(b1.shape[0], dtype=torch.float32, device=b1.device)
    _lib.check(lib.tds_collision_pairwise_fwd(_lib.ptr(b1), _lib.ptr(b2), b1.shape[0], metric, _lib.ptr(out), _lib.stream_ptr(b1.device)))
    return out.reshape(box1.shape[:-1])


@collision_pairwise.register_fake
def _(box1, box2, metric):
    return box1.new_empty(box1.shape[:-1], dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::collision_pairwise_backward", mutates_args=())
def collision_pairwise_backward(grad_out: Tensor, box1: Tensor, box2: Tensor, metric: int) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    b1, b2 = _lib.as_f32(box1).reshape(-1, 5), _lib.as_f32(box2).reshape(-1, 5)
    g = _lib.as_f32(grad_out).reshape(-1)
    g1, g2 = torch.empty_like(b1), torch.empty_like(b2)
    _lib.check(lib.tds_collision_pairwise_bwd(_lib.ptr(b1), _lib.ptr(b2), b1.shape[0], metric, _lib.ptr(g), _lib.ptr(g1), _lib.ptr(g2),
                                              _lib.stream_ptr(b1.device)))
    return g1.reshape(box1.shape), g2.reshape(box2.shape)


@collision_pairwise_backward.register_fake
def _(grad_out, box1, box2, metric):
    return torch.empty_like(box1, dtype=torch.float32), torch.empty_like(box2, dtype=torch.float32)


collision_pairwise.register_autograd(
    lambda ctx, g: (*collision_pairwise_backward(g.contiguous(), *ctx.saved_tensors, ctx.metric), None),
    setup_context=lambda ctx, inputs, output: (ctx.save_for_backward(inputs[0], inputs[1]), setattr(ctx, "metric", inputs[2])))


@torch.library.custom_op(f"{_NS}::collision_allpairs", mutates_args=())
def collision_allpairs(ego_box: Tensor, all_box: Tensor, mask: Tensor, metric: int, ego_is_prefix: bool) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    e, a, m = _lib.as_f32(ego_box), _lib.as_f32(all_box), _lib.as_u8(mask)
    B, A, N = e.shape[0], e.shape[1], a.shape[1]
    out = torch.empty(B, A, dtype=torch.float32, device=e.device)
    arg = torch.empty(B, A, dtype=torch.int32, device=e.device)
    _lib.check(lib.tds_collision_allpairs_fwd(_lib.ptr(e), _lib.ptr(a), _lib.ptr(m), B, A, N, metric, 1 if ego_is_prefix else 0,
                                              _lib.ptr(out), _lib.ptr(arg), _lib.stream_ptr(e.device)))
    return out, arg


@collision_allpairs.register_fake
def _(ego_box, all_box, mask, metric, ego_is_prefix):
    shape = ego_box.shape[:2]
    return ego_box.new_empty(shape, dtype=torch.float32), ego_box.new_empty(shape, dtype=torch.int32)


@torch.library.custom_op(f"{_NS}::collision_allpairs_backward", mutates_args=())
def collision_allpairs_backward(grad_out: Tensor, ego_box: Tensor, all_box: Tensor, mask: Tensor, argmax: Tensor, metric: int,
                                ego_is_prefix: bool) -> Tuple[Tensor, Tensor]:
    lib = _lib.load()
    e, a, m = _lib.as_f32(ego_box), _lib.as_f32(all_box), _lib.as_u8(mask)
    B, A, N = e.shape[0], e.shape[1], a.shape[1]
    ge, ga = torch.zeros_like(e), torch.zeros_like(a)
    _lib.check(lib.tds_collision_allpairs_bwd(_lib.ptr(e), _lib.ptr(a), _lib.ptr(m), B, A, N, metric, 1 if ego_is_prefix else 0,
                                              _lib.ptr(_lib.as_f32(grad_out)), _lib.ptr(argmax.contiguous()), _lib.ptr(ge), _lib.ptr(ga),
                                              _lib.stream_ptr(e.device)))
    return ge, ga


@collision_allpairs_backward.register_fake
def _(grad_out, ego_box, all_box, mask, argmax, metric, ego_is_prefix):
    return torch.empty_like(ego_box, dtype=torch.float32), torch.empty_like(all_box, dtype=torch.float32)


def _ap_setup(ctx, inputs, output):
    ego, allb, mask, metric, prefix = inputs
    ctx.save_for_backward(ego, allb, mask, output[1])
    ctx.meta = (metric, prefix)
    ctx.mark_non_differentiable(output[1])


def _ap_backward(ctx, g_out, g_arg):
    ego, allb, mask, arg = ctx.saved_tensors
    ge, ga = collision_allpairs_backward(g_out.contiguous(), ego, allb, mask, arg, *ctx.meta)
    return ge, ga, None, None, None


collision_allpairs.register_autograd(_ap_backward, setup_context=_ap_setup)


# ------------------------------------------------------------------------------------ boxes
@torch.library.custom_op(f"{_NS}::agent_boxes", mutates_args=())
def agent_boxes(state: Tensor, size: Tensor) -> Tensor:
    lib = _lib.load()
    s, z = _lib.as_f32(state), _lib.as_f32(size[..., :2])
    box = torch.empty(s.shape[:-1] + (5,), dtype=torch.float32, device=s.device)
    _lib.check(lib.tds_agent_boxes(_lib.ptr(s), _lib.ptr(z), s[..., 0].numel(), _lib.ptr(box), None, None, _lib.stream_ptr(s.device)))
    return box


@agent_boxes.register_fake
def _(state, size):
    return state.new_empty(state.shape[:-1] + (5,), dtype=torch.float32)


def _box_backward(ctx, g):
    gs = torch.cat([g[..., 0:2], g[..., 4:5], torch.zeros_like(g[..., 0:1])], dim=-1)
    gz = g[..., 2:4]
    if ctx.size_cols > 2:
        gz = torch.cat([gz, g.new_zeros(gz.shape[:-1] + (ctx.size_cols - 2,))], dim=-1)
    return gs, gz


agent_boxes.register_autograd(_box_backward, setup_context=lambda ctx, inputs, output: setattr(ctx, "size_cols", inputs[1].shape[-1]))
