"""Infraction metrics on the GPU, with the reference's function signatures.

collision_detection_with_discs   torchdrivesim/infractions.py:503-545 (+ bbox2discs :378-409)
iou_differentiable               torchdrivesim/infractions.py:307-324 -> _iou_utils.py:344-367
offroad_infraction_loss          torchdrivesim/infractions.py:176-229 (+ point_to_mesh_distance_pt :86-173)
collision_allpairs               fused Simulator.compute_collision (simulator.py:1161-1194)
"""
from typing import Optional, Union

import torch
from torch import Tensor

from . import _lib, ops
from .maps import MapSet, StaticMap


def collision_detection_with_discs(box1: Tensor, box2: Tensor, num_discs: int = 5, backend: str = 'torch') -> Tensor:
    """Differentiable TrafficSim disc collision loss between corresponding boxes.
    box1, box2: BxAx5 (x, y, length, width, orientation) -> BxA in [0,1]."""
    if num_discs != 5:
        raise _lib.TdsError("the B200 kernel implements the reference default of 5 discs")
    if backend != 'torch':
        raise _lib.TdsError("only the torch-tensor interface exists (the numpy backend is CPU code)")
    return ops.collision_pairwise(box1, box2, _lib.METRIC_DISCS)


def iou_differentiable(box1: Tensor, box2: Tensor, fast: bool = True) -> Tensor:
    """Rotated-box IoU between corresponding boxes, BxAx5 -> BxA (differentiable)."""
    if not fast:
        raise _lib.TdsError("only the `fast` IoU of the reference is implemented")
    return ops.collision_pairwise(box1, box2, _lib.METRIC_IOU)


def collision_allpairs(ego_box: Tensor, all_box: Tensor, mask: Tensor, metric: str = "discs",
                       ego_is_prefix: bool = True) -> Tensor:
    """out[b,i] = sum_j o(ego_i, all_j) mask[b,j] - max_j o(ego_i, all_j) mask[b,j]  (one launch)."""
    m = {"discs": _lib.METRIC_DISCS, "iou": _lib.METRIC_IOU}.get(str(metric))
    if m is None:
        raise _lib.TdsError(f"unsupported collision metric '{metric}' (supported: discs, iou)")
    return ops.collision_allpairs(ego_box, all_box, mask, m, ego_is_prefix)


def offroad_infraction_loss(agent_states: Tensor, lenwid: Tensor, driving_surface_mesh: Union[MapSet, StaticMap],
                            threshold: float = 0, use_pytorch3d: Optional[bool] = None) -> Tensor:
    """Sum over the 4 box corners of the thresholded squared distance to the driving surface.
    agent_states BxAx4, lenwid BxAx2 or Bx2, mesh: StaticMap / MapSet -> BxA."""
    if isinstance(driving_surface_mesh, StaticMap):
        driving_surface_mesh = MapSet([driving_surface_mesh])
    if not isinstance(driving_surface_mesh, MapSet):
        from .mesh import _as_mapset
        driving_surface_mesh = _as_mapset(driving_surface_mesh)
    if agent_states.shape[1] == 0 or all(m.faces.shape[0] == 0 for m in driving_surface_mesh.maps):
        return torch.zeros_like(agent_states[..., 0])
    return ops.offroad(agent_states, lenwid, driving_surface_mesh, threshold)
