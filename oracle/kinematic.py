"""ORACLE (test infrastructure): kinematic state transitions, written in torch so that
autograd supplies the backward reference.

bicycle_step            torchdrivesim/kinematic.py:462-477  KinematicBicycle.step
bicycle_step(no_reversing=True)  kinematic.py:509-523       BicycleNoReversing.step
unicycle_step           NOT in the reference (README.md:16 only); defined by this build, SURVEY.md App. B-1
simple_step / oriented  kinematic.py:362-367, 384-389
"""
import math

import torch

MODEL_BICYCLE = 0
MODEL_BICYCLE_NO_REVERSING = 1
MODEL_UNICYCLE = 2
MODEL_SIMPLE = 3
MODEL_ORIENTED = 4


def bicycle_step(state, action, lr, dt=0.1, left_handed=False, max_acceleration=5.0, max_steering=math.pi / 2,
                 no_reversing=False):
    """state [...,4] (x,y,psi,v), action [...,2] in [-1,1], lr [...]. Returns the new state."""
    norm = torch.tensor([max_acceleration, max_steering], dtype=state.dtype)
    act = action * norm
    if no_reversing:
        acc, beta = act[..., 0], act[..., 1]
        v = state[..., 3]
        reversing = v + acc * dt < 0
        acc = torch.where(reversing, -v / dt, acc)
        act = torch.stack([acc, beta], dim=-1) / norm       # normalize ...
        act = act * norm                                     # ... and denormalize again (kinematic.py:521-523)
    a, beta = act[..., 0], act[..., 1]
    if left_handed:
        beta = -beta
    x, y, psi, v = state[..., 0], state[..., 1], state[..., 2], state[..., 3]
    v = v + a * dt
    x = x + v * torch.cos(psi + beta) * dt
    y = y + v * torch.sin(psi + beta) * dt
    psi = psi + (v / lr) * torch.sin(beta) * dt
    return torch.stack([x, y, psi, v], dim=-1)


def unicycle_step(state, action, dt=0.1, max_acceleration=5.0, max_yaw_rate=math.pi / 2, left_handed=False):
    """Build-defined unicycle: action (a, omega) scaled by (max_acceleration, max_yaw_rate);
    v' = v + a dt; psi' = psi + omega dt; x' = x + v' cos(psi) dt; y' = y + v' sin(psi) dt.
    `left_handed` negates omega, mirroring the bicycle's steering flip (kinematic.py:466-467)."""
    norm = torch.tensor([max_acceleration, max_yaw_rate], dtype=state.dtype)
    act = action * norm
    a, om = act[..., 0], act[..., 1]
    if left_handed:
        om = -om
    x, y, psi, v = state[..., 0], state[..., 1], state[..., 2], state[..., 3]
    v = v + a * dt
    x = x + v * torch.cos(psi) * dt
    y = y + v * torch.sin(psi) * dt
    psi = psi + om * dt
    return torch.stack([x, y, psi, v], dim=-1)


def simple_step(state, action, dt=0.1, max_dx=20.0, max_dpsi=10 * math.pi, max_dv=5.0, oriented=False):
    """kinematic.py:362-367 (SimpleKinematicModel.step) and :384-389 (OrientedKinematicModel.step)."""
    if oriented:
        psi = state[..., 2:3]
        c, s = torch.cos(psi), torch.sin(psi)
        xy = torch.cat([c * action[..., 0:1] + (-s) * action[..., 1:2], s * action[..., 0:1] + c * action[..., 1:2]], -1)
        action = torch.cat([xy, action[..., 2:]], dim=-1)
    norm = torch.tensor([max_dx, max_dx, max_dpsi, max_dv], dtype=state.dtype)
    return state + (action * norm) * dt


def compound_step(state, action, lr, model, dt=0.1, left_handed=False, max_acceleration=5.0,
                  max_steering=math.pi / 2, max_yaw_rate=math.pi / 2):
    """Per-agent model id dispatch (what CompoundKinematicModel.step, kinematic.py:197-201, does by
    boolean-mask splitting).  model [...] int in {0 bicycle, 1 no-reversing bicycle, 2 unicycle}."""
    lr_safe = torch.where(model >= MODEL_UNICYCLE, torch.ones_like(lr), lr)
    full_action, action = action, action[..., :2]
    b0 = bicycle_step(state, action, lr_safe, dt, left_handed, max_acceleration, max_steering, False)
    b1 = bicycle_step(state, action, lr_safe, dt, left_handed, max_acceleration, max_steering, True)
    u = unicycle_step(state, action, dt, max_acceleration, max_yaw_rate, left_handed)
    m = model.unsqueeze(-1)
    out = torch.where(m == MODEL_BICYCLE, b0, torch.where(m == MODEL_BICYCLE_NO_REVERSING, b1, u))
    if full_action.shape[-1] == 4:
        out = torch.where(m == MODEL_SIMPLE, simple_step(state, full_action, dt), out)
        out = torch.where(m == MODEL_ORIENTED, simple_step(state, full_action, dt, oriented=True), out)
    return out
