"""Traffic-light schedules as tables.

The reference drives the lights of a map with one Python state machine per group of lights
(torchdrivesim/traffic_lights.py:27-300, read from `{map}_traffic_light_controller.json`), ticked on the host every
step, and rebuilds the state tensor of `TrafficLightControl` from a dict afterwards.  Here a schedule is three arrays
per map - duration and successor of every (group, phase) and the phase each light shows - advanced in one pass over the groups, and normally not advanced during a rollout at all: `unroll` tabulates the states of the next T steps
once, and that table is the `replay_states` of a `TrafficLightControl`, stepped by a device-side gather
(`BaseTrafficControl.step`): no host work per step, no synchronisation, capturable in a CUDA graph.

The countdown arithmetic is the reference's (Python floats: `remaining -= dt` accumulates its rounding exactly as
there), so the tabulated states equal what its `tick(dt)` produces, tick for tick (tests/golden/light_schedule.npz).
`unroll_controller` tabulates any object with the reference's interface (`tick(dt)`, `current_state_with_name`), e.g.
the reference's own `TrafficLightController`.
"""
import json
from typing import Dict, List, Sequence, Tuple

import torch

LIGHT_STATES = ("none", "green", "yellow", "red")            # TrafficLightState, traffic_lights.py:16-20


def unroll_controller(controller, traffic_light_ids: Sequence[int], dt: float, steps: int,
                      allowed_states: Sequence[str] = ("red", "yellow", "green")) -> torch.Tensor:
    """[L, steps] int64: the state index of every light before tick 0, 1, ..., steps-1 of `controller` (which is
    advanced by steps - 1 ticks).  Works on this module's `TrafficLightController` and on the reference's."""
    allowed = list(allowed_states)
    cols = []
    for t in range(steps):
        names = controller.current_state_with_name
        cols.append(torch.tensor([allowed.index(names[str(i)]) for i in traffic_light_ids]))
        if t + 1 < steps:
            controller.tick(dt)
    return torch.stack(cols, dim=-1)


class TrafficLightController:
    """The schedule of one map: G groups, group g cycling through its phases.

    duration[g][p], successor[g][p]  seconds and next phase of phase p of group g (traffic_lights.py:27-35);
    lights[g][p]                     {light id: state name} shown during that phase;
    phase[g], remaining[g]           the running state: current phase and seconds left in it.
    """

    def __init__(self, duration: List[List[float]], successor: List[List[int]], lights: List[List[Dict[str, str]]],
                 sequence_number: List[List[int]] = None):
        self.duration = [list(map(float, d)) for d in duration]
        self.successor = [list(map(int, s)) for s in successor]
        self.lights = lights
        self.sequence_number = sequence_number or [list(range(len(d))) for d in duration]
        # the reference starts every group from a RANDOM phase (reset); here: phase 0 with its full duration; use set_to
        self.phase = [0 for _ in self.duration]
        self.remaining = [d[0] for d in self.duration]

    @classmethod
    def from_json(cls, json_file_path: str) -> "TrafficLightController":
        """[[{"actor_states": {id: name}, "state": k, "duration": s, "next_state": j}, ...] per group] (traffic_lights.py:181-205)."""
        with open(json_file_path, "rb") as f:
            groups = json.load(f)
        try:
            return cls(duration=[[float(p["duration"]) for p in g] for g in groups],
                       successor=[[int(p["next_state"]) for p in g] for g in groups],
                       lights=[[{str(k): str(v) for k, v in p["actor_states"].items()} for p in g] for g in groups],
                       sequence_number=[[int(p["state"]) for p in g] for g in groups])
        except KeyError as e:
            raise ValueError(f"KeyError: {e} in {json_file_path}")

    # ---- the countdown (traffic_lights.py:107-135), all groups ------------------------------------------------
    def _enter(self, g: int, phase: int, remaining: float) -> None:
        phase = min(max(int(phase), 0), len(self.duration[g]) - 1)
        self.phase[g] = phase
        self.remaining[g] = min(remaining, self.duration[g][phase])

    def tick(self, dt: float) -> None:
        for g in range(len(self.phase)):
            left = self.remaining[g] - dt
            phase = self.phase[g]
            # a phase that has run out hands its deficit to its successors until one of them outlasts it; a countdown
            # that lands on exactly zero starts the successor with its full duration
            while left <= 0:
                nxt = self.successor[g][phase]
                if left == 0:
                    phase, left = nxt, self.duration[g][nxt]
                    break
                left += self.duration[g][nxt]
                phase = nxt
            self.phase[g], self.remaining[g] = phase, left

    def set_to(self, light_states: Sequence[Tuple[float, float]]) -> None:
        """[(phase, seconds remaining)] per group (traffic_lights.py:240-244)."""
        for g, (phase, remaining) in enumerate(light_states):
            if g < len(self.phase):
                self._enter(g, int(phase), remaining)

    # ---- queries ----------------------------------------------------------------------------------------------
    @property
    def state_per_machine(self) -> List[int]:
        return [self.sequence_number[g][p] for g, p in enumerate(self.phase)]

    @property
    def time_remaining(self) -> List[float]:
        return list(self.remaining)

    @property
    def current_state_with_name(self) -> Dict[str, str]:
        out: Dict[str, str] = {}
        for g, p in enumerate(self.phase):
            out.update(self.lights[g][p])
        return out

    def get_number_of_light_groups(self) -> int:
        return len(self.phase)

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
        return unroll_controller(self, traffic_light_ids, dt, steps, allowed_states)
