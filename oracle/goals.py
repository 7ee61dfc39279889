"""ORACLE (test infrastructure): waypoint goals.

Restates  WaypointGoal.step / _update_mask / _advance_state / _agent_waypoint_overlap   torchdrivesim/goals.py:159-217
          WaypointGoal.get_waypoints / get_masks                                        torchdrivesim/goals.py:33-105
Pinned by tests/golden/goals.npz (unmodified reference, eight steps inside Simulator.step).
"""
import numpy as np


def waypoint_step(agent_xy, waypoints, mask, state, threshold=2.0):
    """agent_xy [B,A,2], waypoints [B,A,N,M,2], mask [B,A,N,M] bool, state [B,A,1] int -> (mask, state)."""
    wp = np.asarray(waypoints, np.float32)
    mask = np.array(mask, bool)
    state = np.array(state, np.int64)
    B, A, N, M = mask.shape
    a = np.asarray(agent_xy, np.float32)
    for b in range(B):
        for i in range(A):
            s = int(state[b, i, 0])
            d = wp[b, i, s] - a[b, i]
            dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1], dtype=np.float32)
            if ((dist <= np.float32(threshold)) & mask[b, i, s]).any():
                mask[b, i, s] = False
                state[b, i, 0] = min(s + 1, N - 1)
    return mask, state


def gather(waypoints, mask, state, count=1):
    """-> (waypoints [B,A,count*M,2], mask [B,A,count*M]) of the next `count` collections, zeros past the last one."""
    wp = np.asarray(waypoints, np.float32)
    mask = np.asarray(mask, bool)
    B, A, N, M = mask.shape
    ow = np.zeros((B, A, count, M, 2), np.float32)
    om = np.zeros((B, A, count, M), bool)
    for b in range(B):
        for i in range(A):
            for c in range(count):
                s = int(state[b, i, 0]) + c
                if s < N:
                    ow[b, i, c], om[b, i, c] = wp[b, i, s], mask[b, i, s]
    return ow.reshape(B, A, count * M, 2), om.reshape(B, A, count * M)
