"""Observation noise models: `ObservationNoise` and `StandardSensingObservationNoise` of the reference
(torchdrivesim/observation_noise.py:33-132).  The base model perceives the truth (broadcast views); the standard
sensing model adds distance-dependent Gaussian noise to the perceived states and hides agents that are occluded by a
third agent - O(A (A+Npc)^2) line-circle tests, one kernel each instead of the reference's BxAxExEx2 tensors.
The map-from-log model (noisy lane features / background meshes) is not part of the hot path and is not provided.
"""
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import ops


@dataclass
class ObservationNoiseConfig:
    _type_: str = 'base'


@dataclass
class StandardSensingObservationNoiseConfig:
    _type_: str = 'standard_sensing'


class ObservationNoise:
    def __init__(self, cfg: Optional[ObservationNoiseConfig] = None):
        self.cfg = cfg or ObservationNoiseConfig()

    def get_noisy_state(self, simulator) -> Tensor:
        """BxAx(A+Npc)x4 (observation_noise.py:37-41)."""
        return simulator.get_all_agent_state()[:, None].expand(-1, simulator.agent_count, -1, -1)

    def get_noisy_present_mask(self, simulator) -> Tensor:
        return simulator.get_all_agent_present_mask()[:, None].expand(-1, simulator.agent_count, -1)

    def get_noisy_agent_size(self, simulator) -> Tensor:
        return simulator.get_all_agent_size()[:, None].expand(-1, simulator.agent_count, -1, -1)

    def get_noisy_traffic_controls(self, simulator):
        return simulator.traffic_controls

    def get_noisy_road_mesh(self, simulator):
        return simulator.road_mesh


class StandardSensingObservationNoise(ObservationNoise):
    def __init__(self, cfg: Optional[StandardSensingObservationNoiseConfig] = None):
        super().__init__(cfg or StandardSensingObservationNoiseConfig())

    def sample_noise(self, shape, device) -> Tensor:
        """Standard normal deviates of the perceived states (torch.randn_like in the reference)."""
        return torch.randn(shape, dtype=torch.float32, device=device)

    def get_noisy_state(self, simulator) -> Tensor:
        all_state = simulator.get_all_agent_state().detach()
        B, N = all_state.shape[0], all_state.shape[1]
        A = simulator.agent_count
        return ops.sensing_noise(all_state, A, self.sample_noise((B, A, N, 4), all_state.device))

    def get_noisy_present_mask(self, simulator) -> Tensor:
        return ops.sensing_occlusion(simulator.get_all_agent_state().detach(), simulator.get_all_agent_size()[..., :2],
                                     simulator.get_all_agent_present_mask(), simulator.agent_count)
