"""Waypoint goals: `WaypointGoal` of the reference (torchdrivesim/goals.py:11-217) with the per-step bookkeeping
(distance test, mask update, state advance) in ONE kernel and the gathers of `get_waypoints` / `get_masks` in another,
instead of the reference's chains of gather / scatter / where ops.  Same attributes (`waypoints` BxAxNxMx2, `mask`
BxAxNxM, `state` BxAx1 int64, `max_goal_idx`) and batch plumbing.
"""
from typing import Optional

import torch
from torch import Tensor

from . import _lib


class WaypointGoal:
    def __init__(self, waypoints: Tensor, mask: Optional[Tensor] = None):
        if waypoints.dim() != 5 or waypoints.shape[-1] != 2:
            raise _lib.TdsError("waypoints must be [B,A,N,M,2]")
        self.waypoints = waypoints
        self.mask = mask if mask is not None else self._default_mask()
        self.max_goal_idx = self.waypoints.shape[2]
        self.state = self._default_state()

    def _default_mask(self) -> Tensor:
        return torch.ones(*self.waypoints.shape[:-1], dtype=torch.bool, device=self.waypoints.device)

    def _default_state(self) -> Tensor:
        return torch.zeros(*self.waypoints.shape[:2] + (1,), dtype=torch.long, device=self.waypoints.device)

    def _gather(self, count: int):
        lib = _lib.load()
        wp = _lib.as_f32(self.waypoints)
        B, A, N, M = wp.shape[:4]
        out_wp = torch.empty(B, A, count * M, 2, dtype=torch.float32, device=wp.device)
        out_mask = torch.empty(B, A, count * M, dtype=torch.uint8, device=wp.device)
        _lib.check(lib.tds_waypoint_gather(_lib.ptr(wp), _lib.ptr(_lib.as_u8(self.mask)), _lib.ptr(self.state.contiguous()),
                                           B * A, N, M, int(count), _lib.ptr(out_wp), _lib.ptr(out_mask),
                                           _lib.stream_ptr(wp.device)))
        return out_wp, out_mask.view(torch.bool)

    def get_masks(self, count: int = 1) -> Tensor:
        """BxAx(count*M): the masks of the next `count` collections (goals.py:33-68)."""
        return self._gather(count)[1]

    def get_waypoints(self, count: int = 1) -> Tensor:
        """BxAx(count*M)x2: the waypoints of the next `count` collections (goals.py:70-105)."""
        return self._gather(count)[0]

    def get_waypoints_and_masks(self, count: int = 1):
        """Both of the above from one launch."""
        return self._gather(count)

    def step(self, agent_states: Tensor, time: int = 0, threshold: float = 2.0, in_place: bool = False) -> None:
        """goals.py:159-172.  `mask` and `state` are replaced by updated copies, as in the reference; with `in_place`
        the existing tensors are updated instead (they must be a contiguous bool / uint8 mask and a contiguous int64
        state): what a CUDA graph needs, which reads and writes fixed buffers (GraphedHotPath)."""
        lib = _lib.load()
        st = _lib.as_f32(agent_states)
        B, A, N, M = self.waypoints.shape[:4]
        if st.dim() != 3 or st.shape[0] != B or st.shape[1] != A or st.shape[-1] != 4:
            raise _lib.TdsError("agent_states must be [B,A,4] with the batch and agent counts of the waypoints")
        if in_place:
            if self.mask.dtype not in (torch.bool, torch.uint8) or not self.mask.is_contiguous() or \
                    self.state.dtype != torch.int64 or not self.state.is_contiguous():
                raise _lib.TdsError("WaypointGoal.step(in_place=True) needs a contiguous bool mask and a contiguous int64 state")
            mask = self.mask.view(torch.uint8) if self.mask.dtype == torch.bool else self.mask
            state = self.state
        else:
            mask = _lib.as_u8(self.mask).clone()
            state = self.state.to(torch.int64).contiguous().clone()
        _lib.check(lib.tds_waypoint_step(_lib.ptr(st), _lib.ptr(_lib.as_f32(self.waypoints)), _lib.ptr(mask), _lib.ptr(state),
                                         B * A, N, M, float(threshold), _lib.stream_ptr(st.device)))
        if not in_place:
            self.mask = mask.view(torch.bool)
            self.state = state

    # ---- batch plumbing (goals.py:107-157) -------------------------------------------------------
    def copy(self):
        other = self.__class__(waypoints=self.waypoints.clone(), mask=self.mask.clone())
        other.state = self.state.clone()
        return other

    def to(self, device):
        self.waypoints, self.mask, self.state = self.waypoints.to(device), self.mask.to(device), self.state.to(device)
        return self

    def extend(self, n: int, in_place: bool = True):
        if not in_place:
            return self.copy().extend(n, in_place=True)
        grow = lambda x: x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])
        self.waypoints, self.mask, self.state = grow(self.waypoints), grow(self.mask), grow(self.state)
        return self

    def select_batch_elements(self, idx: Tensor, in_place: bool = True):
        if not in_place:
            return self.copy().select_batch_elements(idx, in_place=True)
        self.waypoints, self.mask, self.state = self.waypoints[idx], self.mask[idx], self.state[idx]
        return self
