"""Non-playable agents: NPCController / SpawnController (torchdrivesim/simulator.py:54-203) and ReplayController
(torchdrivesim/behavior/replay.py:46-107) with the per-step update done by ONE kernel (tds_npc_advance) instead of
the reference's chain of index / where / is_inside_polygon ops.

Same constructor arguments, attributes (`npc_size`, `npc_state`, `npc_present_mask`, `npc_types`, `spawn_controller`,
`time`) and batch plumbing (`to`, `copy`, `extend`, `select_batch_elements`) as the reference classes.
"""
from typing import List, Optional

import torch
from torch import Tensor

from . import _lib


def _grow(x: Tensor, n: int) -> Tensor:
    return x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])


class SpawnController:
    """exit_boundary BxVx2 (convex polygon; NPCs outside are despawned), spawn_states BxNpxTx4 and spawn_masks BxNpxT
    (absent NPCs flagged at the current time are spawned at the given state) - simulator.py:54-85."""

    def __init__(self, exit_boundary: Optional[Tensor] = None, spawn_states: Optional[Tensor] = None,
                 spawn_masks: Optional[Tensor] = None):
        self.exit_boundary = exit_boundary
        self.spawn_states = spawn_states
        self.spawn_masks = spawn_masks
        self.time = 0

    def spawn_despawn_npcs(self, simulator) -> None:
        _advance(simulator.npc_controller, self, replay=False)
        self.time += 1

    def to(self, device):
        for k in ("exit_boundary", "spawn_states", "spawn_masks"):
            if getattr(self, k) is not None:
                setattr(self, k, getattr(self, k).to(device))
        return self

    def copy(self):
        return self.__class__(self.exit_boundary, self.spawn_states, self.spawn_masks)

    def extend(self, n: int, in_place: bool = True):
        if not in_place:
            return self.copy().extend(n, in_place=True)
        for k in ("exit_boundary", "spawn_states", "spawn_masks"):
            if getattr(self, k) is not None:
                setattr(self, k, _grow(getattr(self, k), n))
        return self

    def select_batch_elements(self, idx: Tensor, in_place: bool = True):
        if not in_place:
            return self.copy().select_batch_elements(idx, in_place=True)
        for k in ("exit_boundary", "spawn_states", "spawn_masks"):
            if getattr(self, k) is not None:
                setattr(self, k, getattr(self, k)[idx])
        return self


class NPCController:
    """Base class: the NPC state is left unchanged by a step, apart from spawning / despawning (simulator.py:127-203)."""

    def __init__(self, npc_size: Tensor, npc_state: Tensor, npc_present_mask: Optional[Tensor] = None,
                 npc_types: Optional[Tensor] = None, agent_type_names: Optional[List[str]] = None,
                 spawn_controller: Optional[SpawnController] = None):
        self.npc_size = npc_size
        self.npc_state = npc_state
        self.npc_present_mask = npc_present_mask
        if self.npc_present_mask is None:
            self.npc_present_mask = torch.ones_like(npc_state[..., 0], dtype=torch.bool)
        self.npc_types = npc_types
        if self.npc_types is None:
            self.npc_types = torch.zeros_like(self.npc_present_mask).long()
        self.agent_type_names = agent_type_names or ['vehicle']
        self.spawn_controller = spawn_controller or SpawnController()

    def get_npc_state(self) -> Tensor:
        return self.npc_state

    def get_npc_size(self) -> Tensor:
        return self.npc_size

    def get_npc_types(self) -> Tensor:
        return self.npc_types

    def get_npc_present_mask(self) -> Tensor:
        return self.npc_present_mask

    def spawn_despawn_npcs(self, simulator) -> None:
        self.spawn_controller.spawn_despawn_npcs(simulator)

    def advance_npcs(self, simulator) -> None:
        self.spawn_despawn_npcs(simulator)

    _tensors = ("npc_size", "npc_state", "npc_present_mask", "npc_types")

    def to(self, device):
        for k in self._tensors:
            setattr(self, k, getattr(self, k).to(device))
        self.spawn_controller.to(device)
        return self

    def copy(self):
        return self.__class__(self.npc_size, self.npc_state, self.npc_present_mask, self.npc_types, self.agent_type_names,
                              self.spawn_controller.copy())

    def extend(self, n: int, in_place: bool = True):
        if not in_place:
            return self.copy().extend(n, in_place=True)
        for k in self._tensors:
            setattr(self, k, _grow(getattr(self, k), n))
        self.spawn_controller.extend(n, in_place=True)
        return self

    def select_batch_elements(self, idx: Tensor, in_place: bool = True):
        if not in_place:
            return self.copy().select_batch_elements(idx, in_place=True)
        for k in self._tensors:
            setattr(self, k, getattr(self, k)[idx])
        self.spawn_controller.select_batch_elements(idx, in_place=True)
        return self


class CompoundNPCController(NPCController):
    """Several controllers, each responsible for the NPCs that `controller_indices` [B,Np] assigns to it
    (simulator.py:206-250): after every advance the NPC tensors are gathered from the controllers by assignment and
    handed back to all of them."""

    def __init__(self, controllers: List[NPCController], controller_indices: Tensor):
        batch_size, num_agents = controller_indices.shape
        dev = controller_indices.device
        super().__init__(torch.zeros((batch_size, num_agents, 2), device=dev), torch.zeros((batch_size, num_agents, 4), device=dev),
                         torch.zeros((batch_size, num_agents), device=dev, dtype=torch.bool),
                         torch.zeros((batch_size, num_agents), device=dev, dtype=torch.long), controllers[0].agent_type_names)
        self.controllers = controllers
        self.controller_indices = controller_indices
        self.gather_npc_states()

    def gather_npc_states(self) -> None:
        for i, controller in enumerate(self.controllers):
            mask = self.controller_indices == i
            self.npc_size = controller.npc_size.where(mask.unsqueeze(-1), self.npc_size)
            self.npc_state = controller.npc_state.where(mask.unsqueeze(-1), self.npc_state)
            self.npc_present_mask = controller.npc_present_mask.where(mask, self.npc_present_mask)
            self.npc_types = controller.npc_types.where(mask, self.npc_types)
        for controller in self.controllers:          # every controller sees all NPCs
            controller.npc_size, controller.npc_state = self.npc_size, self.npc_state
            controller.npc_present_mask, controller.npc_types = self.npc_present_mask, self.npc_types

    def advance_npcs(self, simulator) -> None:
        for controller in self.controllers:
            controller.advance_npcs(simulator)
        self.gather_npc_states()

    def to(self, device):
        super().to(device)
        self.controller_indices = self.controller_indices.to(device)
        self.controllers = [c.to(device) for c in self.controllers]
        return self

    def copy(self):
        return self.__class__([c.copy() for c in self.controllers], self.controller_indices.clone())

    def extend(self, n: int, in_place: bool = True):
        if not in_place:
            return self.copy().extend(n, in_place=True)
        super().extend(n, in_place=True)
        self.controller_indices = _grow(self.controller_indices, n)
        for c in self.controllers:
            c.extend(n, in_place=True)
        return self

    def select_batch_elements(self, idx: Tensor, in_place: bool = True):
        if not in_place:
            return self.copy().select_batch_elements(idx, in_place=True)
        super().select_batch_elements(idx, in_place=True)
        self.controller_indices = self.controller_indices[idx]
        for c in self.controllers:
            c.select_batch_elements(idx, in_place=True)
        return self


class ReplayController(NPCController):
    """NPCs that replay a log: npc_states BxNpxTx4, npc_present_masks BxNpxT; the log wraps around at its end
    (behavior/replay.py:46-107)."""

    def __init__(self, npc_size, npc_states, npc_present_masks: Optional[Tensor] = None, time: int = 0,
                 npc_types: Optional[Tensor] = None, agent_type_names: Optional[List[str]] = None,
                 spawn_controller: Optional[SpawnController] = None):
        self.time = time
        self.npc_states = npc_states
        self.npc_present_masks = npc_present_masks
        if self.npc_present_masks is None:
            self.npc_present_masks = torch.ones_like(self.npc_states[..., 0], dtype=torch.bool)
        super().__init__(npc_size, self.npc_states[..., self.time, :], self.npc_present_masks[..., self.time], npc_types,
                         agent_type_names, spawn_controller)

    def advance_npcs(self, simulator) -> None:
        self.time += 1
        if self.time == self.npc_states.shape[-2]:
            self.time = 0
        _advance(self, self.spawn_controller, replay=True)
        self.spawn_controller.time += 1

    _tensors = NPCController._tensors + ("npc_states", "npc_present_masks")

    def copy(self):
        obj = self.__class__(self.npc_size, self.npc_states, self.npc_present_masks, self.time, self.npc_types,
                             self.agent_type_names, self.spawn_controller.copy())
        obj.npc_state = self.npc_state.clone()
        obj.npc_present_mask = self.npc_present_mask.clone()
        return obj


def _advance(ctrl: NPCController, spawn: SpawnController, replay: bool) -> None:
    """One launch: replay gather (optional), despawn outside the exit boundary, spawn (tds_npc_advance).
    Two deliberate differences from the reference: the NPC states are float32 (the kernel's arithmetic type; a float64
    log is narrowed), and the spawn / despawn edits land on `ctrl` itself - the reference's SpawnController writes to
    `simulator.npc_controller` (simulator.py:71-85), which is the same object except inside a CompoundNPCController,
    where its edits are overwritten by the gather that follows anyway."""
    lib = _lib.load()
    state = _lib.as_f32(ctrl.npc_state)
    if state.dim() != 3 or state.shape[-1] != 4:
        raise _lib.TdsError("npc_state must be [B,Np,4]")
    B, Np = state.shape[0], state.shape[1]
    dev = state.device
    # the kernel updates in place: never write into the replay log or the caller's tensors
    state = state.clone()
    present = _lib.as_u8(ctrl.npc_present_mask).clone()
    rs = rp = None
    T = t = 0
    if replay:
        rs, rp = _lib.as_f32(ctrl.npc_states), _lib.as_u8(ctrl.npc_present_masks)
        T, t = rs.shape[-2], int(ctrl.time)
        if tuple(rs.shape) != (B, Np, T, 4) or tuple(rp.shape) != (B, Np, T):
            raise _lib.TdsError("npc_states must be [B,Np,T,4] and npc_present_masks [B,Np,T]")
    eb = None
    V = 0
    if spawn.exit_boundary is not None:
        eb = _lib.as_f32(spawn.exit_boundary.to(dev))
        if eb.dim() != 3 or eb.shape[0] != B or eb.shape[-1] != 2:
            raise _lib.TdsError("exit_boundary must be [B,V,2]")
        V = eb.shape[1]
    ss = sm = None
    Ts = ts = 0
    if spawn.spawn_states is not None and spawn.spawn_masks is not None:
        ss, sm = _lib.as_f32(spawn.spawn_states.to(dev)), _lib.as_u8(spawn.spawn_masks.to(dev))
        Ts, ts = ss.shape[-2], int(spawn.time)
        if tuple(ss.shape) != (B, Np, Ts, 4) or tuple(sm.shape) != (B, Np, Ts):
            raise _lib.TdsError("spawn_states must be [B,Np,T,4] and spawn_masks [B,Np,T]")
        if not 0 <= ts < Ts:
            raise IndexError(f"spawn time {ts} is outside the spawn table of {Ts} steps")   # as the reference's indexing
    _lib.check(lib.tds_npc_advance(_lib.ptr(rs), _lib.ptr(rp), T, t, _lib.ptr(eb), V, _lib.ptr(ss), _lib.ptr(sm), Ts, ts,
                                   _lib.ptr(state), _lib.ptr(present), B, Np, _lib.stream_ptr(dev)))
    ctrl.npc_state = state
    ctrl.npc_present_mask = present.view(torch.bool)
