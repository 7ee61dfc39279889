"""Traffic controls as raster inputs: rectangular stoplines with a discrete state.

Data-only mirror of torchdrivesim/traffic_controls.py:12-178 (`BaseTrafficControl`,
`TrafficLightControl`): `pos` [B,L,5], `corners` [B,L,4,2] (masked ones at -1000), `state` [B,L],
`replay_states` [B,L,T], `step(time)`, `extend`, `select_batch_elements`, `to`, `copy`, and
`TrafficLightControl.compute_violation` (traffic_controls.py:152-178) on the GPU through
`tds_traffic_light_violation`.
"""
from typing import List, Optional

import torch
from torch import Tensor


def box2corners(box: Tensor) -> Tensor:
    """[...,5] (x, y, length, width, psi) -> [...,4,2]; corner order and rotation of box2corners_th
    (torchdrivesim/_iou_utils.py:270-299)."""
    x, y, l, w, a = (box[..., i:i + 1] for i in range(5))
    x4 = torch.tensor([0.5, -0.5, -0.5, 0.5], dtype=box.dtype, device=box.device) * l
    y4 = torch.tensor([0.5, 0.5, -0.5, -0.5], dtype=box.dtype, device=box.device) * w
    s, c = torch.sin(a), torch.cos(a)
    return torch.stack([x4 * c + y4 * (-s) + x, x4 * s + y4 * c + y], dim=-1)


class BaseTrafficControl:
    def __init__(self, pos: Tensor, allowed_states: Optional[List[str]] = None,
                 replay_states: Optional[Tensor] = None, mask: Optional[Tensor] = None):
        self.pos = pos
        self.allowed_states = allowed_states if allowed_states is not None else self._default_allowed_states()
        self.replay_states = replay_states if replay_states is not None else \
            torch.zeros(*pos.shape[:2] + (0,), dtype=torch.long, device=pos.device)
        self.mask = mask if mask is not None else torch.ones(*pos.shape[:2], dtype=torch.bool, device=pos.device)
        m = self.mask.to(pos.dtype).reshape(self.mask.shape[0], self.mask.shape[1], 1, 1)
        self.corners = box2corners(pos) * m + (1 - m) * -1000
        self.state = self.replay_states[..., 0] if self.replay_states.shape[-1] > 0 else \
            torch.zeros(*pos.shape[:2], dtype=torch.long, device=pos.device)

    @classmethod
    def _default_allowed_states(cls) -> List[str]:
        return ['none']

    @property
    def total_replay_time(self) -> int:
        return self.replay_states.shape[-1]

    def copy(self):
        other = self.__class__(pos=self.pos.clone(), allowed_states=list(self.allowed_states),
                               replay_states=self.replay_states.clone(), mask=self.mask.clone())
        other.corners = self.corners.clone()
        other.state = self.state.clone()
        return other

    def to(self, device):
        for k in ("pos", "corners", "replay_states", "mask", "state"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def extend(self, n: int, in_place: bool = True):
        if not in_place:
            return self.copy().extend(n, in_place=True)
        grow = lambda x: x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])
        for k in ("pos", "corners", "replay_states", "mask", "state"):
            setattr(self, k, grow(getattr(self, k)))
        return self

    def select_batch_elements(self, idx: Tensor, in_place: bool = True):
        if not in_place:
            return self.copy().select_batch_elements(idx, in_place=True)
        for k in ("pos", "corners", "replay_states", "mask", "state"):
            setattr(self, k, getattr(self, k)[idx])
        return self

    def set_state(self, state: Tensor) -> None:
        self.state = state

    def compute_state(self, time: int) -> Tensor:
        return self.state

    def compute_violation(self, agent_state: Tensor) -> Tensor:
        """BxAx5 (x, y, length, width, psi) -> BxA bool; the base class reports no violations
        (traffic_controls.py:137-149)."""
        return torch.zeros(agent_state.shape[0], agent_state.shape[1], dtype=torch.bool, device=agent_state.device)

    def step(self, time: int) -> None:
        if time < self.total_replay_time:
            self.set_state(self.replay_states[..., time])
        else:
            self.set_state(self.compute_state(time))


class TrafficLightControl(BaseTrafficControl):
    violation_rear_factor = 0.1

    @classmethod
    def _default_allowed_states(cls) -> List[str]:
        return ['red', 'yellow', 'green']

    def compute_violation(self, agent_state: Tensor) -> Tensor:
        """An agent violates a light iff the light is red and the rear 10 % of its box overlaps the stop line
        (traffic_controls.py:162-178)."""
        from . import ops
        return ops.traffic_light_violation(agent_state, self.corners, self.state, self.allowed_states.index('red'),
                                           self.violation_rear_factor)


class StopSignControl(BaseTrafficControl):
    pass


class YieldControl(BaseTrafficControl):
    pass
