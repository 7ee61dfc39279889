"""ORACLE (test infrastructure): replayed NPCs with spawning / despawning.

Restates  ReplayController.advance_npcs            torchdrivesim/behavior/replay.py:54-60
          SpawnController.spawn_despawn_npcs       torchdrivesim/simulator.py:71-85
          utils.is_inside_polygon                  torchdrivesim/utils.py:99-122
Pinned by tests/golden/npc.npz (unmodified reference, six steps).
"""
import numpy as np


def is_inside_polygon(point, polygon):
    """point [B,P,2], polygon [B,V,2] (convex, either orientation) -> bool [B,P]."""
    p = np.asarray(point, np.float32)
    poly = np.asarray(polygon, np.float32)[:, None]                  # [B,1,V,2]
    nxt = np.roll(poly, -1, axis=-2)
    a = nxt[..., 1] - poly[..., 1]
    b = poly[..., 0] - nxt[..., 0]
    c = -a * poly[..., 0] - b * poly[..., 1]
    right = a * p[..., None, 0] + b * p[..., None, 1] + c >= 0
    return right.all(-1) | (~right).all(-1)


def npc_advance(state, present, replay=None, replay_present=None, t_replay=0, boundary=None, spawn_states=None,
                spawn_masks=None, t_spawn=0):
    """One step.  state [B,Np,4], present [B,Np]; replay [B,Np,T,4] / replay_present [B,Np,T] or None; boundary
    [B,V,2] or None; spawn_states [B,Np,Ts,4] / spawn_masks [B,Np,Ts] or None -> (state, present)."""
    state = np.array(state, np.float32)
    present = np.array(present, bool)
    if replay is not None:
        state = np.array(replay[:, :, t_replay], np.float32)
        present = np.ones(state.shape[:2], bool) if replay_present is None else np.array(replay_present[:, :, t_replay], bool)
    if boundary is not None:
        present = present & is_inside_polygon(state[..., :2], boundary)
    if spawn_states is not None and spawn_masks is not None:
        spawn = np.asarray(spawn_masks[:, :, t_spawn], bool) & ~present
        present = present | spawn
        state = np.where(spawn[..., None], np.asarray(spawn_states[:, :, t_spawn], np.float32), state)
    return state, present
