"""Traffic-light schedules: the finite state machines of the reference (torchdrivesim/traffic_lights.py:27-300) that
cycle groups of lights through timed states, read from `{map}_traffic_light_controller.json`.

The reference ticks these Python objects on the host every step and rebuilds a state tensor from a dict.  Here the
schedule is unrolled ONCE on the host (`TrafficLightController.unroll`) into the `replay_states` tensor of a
`TrafficLightControl`, so that stepping the lights during a rollout is the device-side gather of
`BaseTrafficControl.step` - no host work, no synchronisation, capturable in a CUDA graph.
Same tick arithmetic as the reference (Python floats), so the unrolled states equal what `tick(dt)` produces.
"""
import json
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

LIGHT_STATES = ("none", "green", "yellow", "red")            # TrafficLightState, traffic_lights.py:16-20


@dataclass
class TrafficLightGroupState:
    actor_states: Dict[str, str]        # actor id -> state name
    sequence_number: int
    duration: float                     # seconds
    next_state: int


class TrafficLightStateMachine:
    """One group of lights (traffic_lights.py:37-157)."""

    def __init__(self, group_states: List[TrafficLightGroupState]):
        self._states = group_states
        self.set_to(0, group_states[0].duration)       # the reference starts from a RANDOM state (reset); call set_to

    def set_to(self, state_index: int, time_remaining: float) -> None:
        state = min(max(int(state_index), 0), len(self._states) - 1)
        self._current_state = self._states[state]
        self._duration = self._current_state.duration
        self._time_remaining = time_remaining if time_remaining <= self._duration else self._duration

    def tick(self, dt: float) -> None:
        self._time_remaining -= dt
        while self._time_remaining <= 0:
            next_state = self._current_state.next_state
            next_duration = self._states[next_state].duration
            if self._time_remaining == 0:
                self.set_to(next_state, next_duration)
                break
            elif self._time_remaining + next_duration > 0:
                self._time_remaining += next_duration
                self.set_to(next_state, self._time_remaining)
                break
            else:
                self._time_remaining += next_duration
                self._current_state = self._states[next_state]

    @property
    def states(self) -> List[TrafficLightGroupState]:
        return self._states

    @property
    def current_state(self) -> TrafficLightGroupState:
        return self._current_state

    @property
    def time_remaining(self) -> float:
        return self._time_remaining


class TrafficLightController:
    """All groups of a map (traffic_lights.py:159-292)."""

    def __init__(self, traffic_fsms: List[TrafficLightStateMachine]):
        self.traffic_fsms = traffic_fsms

    @classmethod
    def from_json(cls, json_file_path: str) -> "TrafficLightController":
        with open(json_file_path, "rb") as f:
            items = json.load(f)
        try:
            return cls([TrafficLightStateMachine([
                TrafficLightGroupState(actor_states={k: str(v) for k, v in gs["actor_states"].items()},
                                       sequence_number=int(gs["state"]), duration=float(gs["duration"]),
                                       next_state=int(gs["next_state"])) for gs in sm]) for sm in items])
        except KeyError as e:
            raise ValueError(f"KeyError: {e} in {json_file_path}")

    def tick(self, dt: float) -> None:
        for fsm in self.traffic_fsms:
            fsm.tick(dt)

    def set_to(self, light_states: Sequence[Tuple[float, float]]) -> None:
        """[(state index, time remaining)] per group (traffic_lights.py:240-244)."""
        for fsm, (state, time_remaining) in zip(self.traffic_fsms, light_states):
            fsm.set_to(int(state), time_remaining)

    @property
    def state_per_machine(self) -> List[int]:
        return [fsm.current_state.sequence_number for fsm in self.traffic_fsms]

    @property
    def time_remaining(self) -> List[float]:
        return [fsm.time_remaining for fsm in self.traffic_fsms]

    @property
    def current_state_with_name(self) -> Dict[str, str]:
        out: Dict[str, str] = {}
        for fsm in self.traffic_fsms:
            out.update(fsm.current_state.actor_states)
        return out

    def get_number_of_light_groups(self) -> int:
        return len(self.traffic_fsms)

    def current_state_tensor(self, traffic_light_ids: Sequence[int], allowed_states: Sequence[str] = ("red", "yellow", "green")) -> torch.Tensor:
        """current_light_state_tensor_from_controller (traffic_lights.py:295-301): index of every light's state in the
        allowed states of TrafficLightControl."""
        names = self.current_state_with_name
        return torch.tensor([list(allowed_states).index(names[str(i)]) for i in traffic_light_ids])

    def unroll(self, traffic_light_ids: Sequence[int], dt: float, steps: int,
               allowed_states: Sequence[str] = ("red", "yellow", "green")) -> torch.Tensor:
        """[L, steps] int64: the state of every light before tick 0, 1, ..., steps-1 from the controller's current
        state on (the controller itself is advanced by steps - 1 ticks).  Feed it to
        `TrafficLightControl(pos, replay_states=unrolled[None].expand(B, -1, -1))`."""
        cols = []
        for t in range(steps):
            cols.append(self.current_state_tensor(traffic_light_ids, allowed_states))
            if t + 1 < steps:
                self.tick(dt)
        return torch.stack(cols, dim=-1)
