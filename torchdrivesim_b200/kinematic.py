"""Kinematic models backed by the fused sm_100a step kernel.

Mirror of the reference's `KinematicModel` interface (torchdrivesim/kinematic.py:20-157) for the models on
the hot path: `KinematicBicycle` (:400-506), `BicycleNoReversing` (:509-523), the unicycle the reference
only names (README.md:16) and `CompoundKinematicModel` (:160-314) re-expressed as a per-agent model id so
that heterogeneous batches step in ONE launch without boolean-mask host syncs.  `step` stays connected to
autograd through a hand-written backward kernel.
"""
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops


class KinematicModel:
    """Base interface (kinematic.py:20-157): state [...,4] = x, y, psi, speed."""
    state_size: int = 4
    action_size: int = 4

    def __init__(self, dt: float = 0.1):
        self.dt = dt
        self.state: Optional[Tensor] = None

    @property
    def batch_size(self) -> int:
        return self.get_state()[..., 0].numel()

    def step(self, action: Tensor, dt: Optional[float] = None, out: Optional[Tensor] = None) -> None:
        """`out`: a float32 tensor of the state's shape that receives the new state and becomes the model's state (it may
        BE the current state: the kernel updates in place); no autograd tape is recorded then - what a CUDA graph uses."""
        raise NotImplementedError

    def fit_action(self, future_state: Tensor, current_state: Optional[Tensor] = None, dt: Optional[float] = None) -> Tensor:
        raise NotImplementedError

    def copy(self, other=None):
        if other is None:
            other = self.__class__(dt=self.dt)
        other.set_params(**self.get_params())
        other.set_state(self.get_state())
        return other

    def to(self, device):
        if self.state is not None:
            self.state = self.state.to(device)
        self.map_param(lambda x: x.to(device))
        return self

    def set_state(self, state: Tensor) -> None:
        self.state = state

    def get_state(self) -> Tensor:
        return self.state

    def get_params(self) -> Dict[str, Tensor]:
        return dict()

    def set_params(self, **kwargs) -> None:
        pass

    def flattening(self, batch_shape) -> None:
        pass

    def unflattening(self, batch_shape) -> None:
        pass

    def map_param(self, f) -> None:
        pass

    def normalize_action(self, action: Tensor) -> Tensor:
        return action

    def denormalize_action(self, action: Tensor) -> Tensor:
        return action

    @staticmethod
    def pack_state(x: Tensor, y: Tensor, psi: Tensor, speed: Tensor) -> Tensor:
        return torch.stack([x, y, psi, speed], dim=-1)

    @staticmethod
    def unpack_state(state: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        return state[..., 0], state[..., 1], state[..., 2], state[..., 3]

    def extend(self, n: int):
        grow = lambda x: x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])
        self.map_param(grow)
        self.set_state(grow(self.get_state()))

    def select_batch_elements(self, idx):
        self.map_param(lambda x: x[idx])
        self.set_state(self.get_state()[idx])


class KinematicBicycle(KinematicModel):
    """Bicycle model with steering applied at the geometric centre (kinematic.py:400-506); `lr` is the
    distance from the centre to the rear axle.  Action = (acceleration, steering) / (max_acceleration,
    max_steering)."""
    action_size: int = 2
    _model_id = _lib.MODEL_BICYCLE

    def __init__(self, max_acceleration=5, max_steering=math.pi / 2, dt=0.1, left_handed=False):
        super().__init__(dt=dt)
        self.max_acceleration = max_acceleration
        self.max_steering = max_steering
        self.left_handed = left_handed
        self._normalization_factor = torch.tensor([self.max_acceleration, self.max_steering])
        self.lr: Optional[Tensor] = None

    def copy(self, other=None):
        if other is None:
            other = self.__class__(max_acceleration=self.max_acceleration, max_steering=self.max_steering, dt=self.dt,
                                   left_handed=self.left_handed)
        other._normalization_factor = self._normalization_factor.clone()
        return super().copy(other)

    def to(self, device):
        super().to(device)
        self._normalization_factor = self._normalization_factor.to(device)
        return self

    def get_params(self):
        return dict(lr=self.lr)

    def set_params(self, **kwargs):
        assert 'lr' in kwargs
        self.lr = kwargs['lr']

    def flattening(self, batch_shape):
        self.lr = self.lr.reshape((int(math.prod(batch_shape)),))

    def unflattening(self, batch_shape):
        self.lr = self.lr.reshape(batch_shape)

    def map_param(self, f):
        assert self.lr is not None
        self.lr = f(self.lr)

    def normalize_action(self, action):
        return action / self._normalization_factor.to(action.device)

    def denormalize_action(self, action):
        return action * self._normalization_factor.to(action.device)

    def _params(self, dt):
        return ops.kinematic_params(self.dt if dt is None else dt, self.max_acceleration, self.max_steering,
                                    self.max_steering, self.left_handed)

    def step(self, action, dt=None, out=None):
        assert action.shape[-1] == 2, "The bicycle model takes as input only acceleration and steering"
        self.set_state(ops.kinematic_step(self.get_state(), action, self.lr, None, self._model_id, self._params(dt), out=out))

    def fit_action(self, future_state, current_state=None, dt=None):
        """Inverse of `step` (kinematic.py:479-506); plain torch, not on the per-step path."""
        dt = self.dt if dt is None else dt
        fx, fy, _, _ = self.unpack_state(future_state)
        cx, cy, cpsi, cv = self.unpack_state(self.get_state() if current_state is None else current_state)
        vx, vy = (fx - cx) / dt, (fy - cy) / dt
        speed = torch.sqrt(vx ** 2 + vy ** 2)
        beta = torch.atan2(vy, vx) - cpsi * torch.sign(torch.abs(speed))
        beta = torch.remainder(beta + math.pi, 2 * math.pi) - math.pi
        reversing = torch.sign(torch.cos(beta)) == -1
        speed = speed * torch.where(reversing, -1, 1)
        beta = torch.where(reversing, beta - math.pi * torch.sign(beta), beta)
        acc = (speed - cv) / dt
        if self.left_handed:
            beta = -beta
        return self.normalize_action(torch.stack([acc, beta], dim=-1))


class BicycleNoReversing(KinematicBicycle):
    """Bicycle that comes to a full stop instead of reversing (kinematic.py:509-523)."""
    _model_id = _lib.MODEL_BICYCLE_NO_REVERSING


class KinematicUnicycle(KinematicModel):
    """Unicycle for pedestrian-like agents.  NOT in the reference (only named in README.md:16); defined
    here as: action (a, omega) scaled by (max_acceleration, max_yaw_rate);
    v' = v + a dt, psi' = psi + omega dt, x' = x + v' cos(psi) dt, y' = y + v' sin(psi) dt."""
    action_size: int = 2

    def __init__(self, max_acceleration=5, max_yaw_rate=math.pi / 2, dt=0.1, left_handed=False):
        super().__init__(dt=dt)
        self.max_acceleration = max_acceleration
        self.max_yaw_rate = max_yaw_rate
        self.left_handed = left_handed
        self._normalization_factor = torch.tensor([self.max_acceleration, self.max_yaw_rate])

    def copy(self, other=None):
        if other is None:
            other = self.__class__(max_acceleration=self.max_acceleration, max_yaw_rate=self.max_yaw_rate, dt=self.dt,
                                   left_handed=self.left_handed)
        return super().copy(other)

    def normalize_action(self, action):
        return action / self._normalization_factor.to(action.device)

    def denormalize_action(self, action):
        return action * self._normalization_factor.to(action.device)

    def step(self, action, dt=None, out=None):
        assert action.shape[-1] == 2
        p = ops.kinematic_params(self.dt if dt is None else dt, self.max_acceleration, math.pi / 2, self.max_yaw_rate,
                                 self.left_handed)
        self.set_state(ops.kinematic_step(self.get_state(), action, None, None, _lib.MODEL_UNICYCLE, p, out=out))

    def fit_action(self, future_state, current_state=None, dt=None):
        dt = self.dt if dt is None else dt
        _, _, fpsi, fv = self.unpack_state(future_state)
        _, _, cpsi, cv = self.unpack_state(self.get_state() if current_state is None else current_state)
        om = (fpsi - cpsi) / dt
        if self.left_handed:
            om = -om
        return self.normalize_action(torch.stack([(fv - cv) / dt, om], dim=-1))


class SimpleKinematicModel(KinematicModel):
    """The action is the time derivative of the state in units of (max_dx, max_dx, max_dpsi, max_dv)
    (kinematic.py:328-376)."""
    action_size: int = 4
    _model_id = _lib.MODEL_SIMPLE

    def __init__(self, max_dx=20, max_dpsi=10 * math.pi, max_dv=5, dt=0.1):
        super().__init__(dt=dt)
        self.max_dx, self.max_dpsi, self.max_dv = max_dx, max_dpsi, max_dv
        self._normalization_factor = torch.tensor([max_dx, max_dx, max_dpsi, max_dv])

    def copy(self, other=None):
        if other is None:
            other = self.__class__(max_dx=self.max_dx, max_dpsi=self.max_dpsi, max_dv=self.max_dv, dt=self.dt)
        return super().copy(other)

    def normalize_action(self, action):
        return action / self._normalization_factor.to(action.device)

    def denormalize_action(self, action):
        return action * self._normalization_factor.to(action.device)

    def step(self, action, dt=None, out=None):
        assert action.shape[-1] == self.action_size
        p = ops.kinematic_params(self.dt if dt is None else dt, max_dx=self.max_dx, max_dpsi=self.max_dpsi,
                                 max_dv=self.max_dv)
        self.set_state(ops.kinematic_step(self.get_state(), action, None, None, self._model_id, p, out=out))

    def fit_action(self, future_state, current_state=None, dt=None):
        dt = self.dt if dt is None else dt
        current_state = self.get_state() if current_state is None else current_state
        return self.normalize_action((future_state - current_state) / dt)


class OrientedKinematicModel(SimpleKinematicModel):
    """Like SimpleKinematicModel with the xy action expressed in the agent frame (kinematic.py:379-397)."""
    _model_id = _lib.MODEL_ORIENTED

    def fit_action(self, future_state, current_state=None, dt=None):
        act = super().fit_action(future_state, current_state=current_state, dt=dt)
        psi = (self.get_state() if current_state is None else current_state)[..., 2:3]
        c, s = torch.cos(-psi), torch.sin(-psi)
        xy = torch.cat([c * act[..., 0:1] - s * act[..., 1:2], s * act[..., 0:1] + c * act[..., 1:2]], dim=-1)
        return torch.cat([xy, act[..., 2:]], dim=-1)


class FusedCompoundKinematicModel(KinematicBicycle):
    """Heterogeneous agents in one launch: `model_assignments` [B,A] holds a model id per agent
    (0 bicycle, 1 no-reversing bicycle, 2 unicycle, 3 simple, 4 oriented).  Replaces CompoundKinematicModel
    (kinematic.py:160-314), whose step splits the batch with boolean masks (a host sync per step).  As there,
    the action is padded to the largest action size (`action_size=4` when simple/oriented agents are present)."""

    def __init__(self, model_assignments: Tensor, max_acceleration=5, max_steering=math.pi / 2,
                 max_yaw_rate=math.pi / 2, dt=0.1, left_handed=False, action_size: int = 2):
        super().__init__(max_acceleration=max_acceleration, max_steering=max_steering, dt=dt, left_handed=left_handed)
        self.max_yaw_rate = max_yaw_rate
        self.model_assignments = model_assignments
        self.action_size = action_size

    def copy(self, other=None):
        if other is None:
            other = self.__class__(self.model_assignments, max_acceleration=self.max_acceleration,
                                   max_steering=self.max_steering, max_yaw_rate=self.max_yaw_rate, dt=self.dt,
                                   left_handed=self.left_handed, action_size=self.action_size)
        return super().copy(other)

    def to(self, device):
        super().to(device)
        self.model_assignments = self.model_assignments.to(device)
        return self

    def extend(self, n: int):
        x = self.model_assignments
        self.model_assignments = x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])
        super().extend(n)

    def select_batch_elements(self, idx):
        self.model_assignments = self.model_assignments[idx]
        super().select_batch_elements(idx)

    def step(self, action, dt=None, out=None):
        assert action.shape[-1] == self.action_size
        p = ops.kinematic_params(self.dt if dt is None else dt, self.max_acceleration, self.max_steering,
                                 self.max_yaw_rate, self.left_handed)
        self.set_state(ops.kinematic_step(self.get_state(), action, self.lr, self.model_assignments, 0, p, out=out))
