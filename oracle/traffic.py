"""ORACLE (test infrastructure): traffic-light violations.

Restates  TrafficLightControl.compute_violation      torchdrivesim/traffic_controls.py:152-178
          box2corners_with_rear_factor               torchdrivesim/_iou_utils.py:302-341
          Simulator.compute_traffic_lights_violations torchdrivesim/simulator.py:1046-1062
An agent violates a light iff the light is red and the rear `rear_factor` part of its box overlaps the stop
line rectangle with positive area (oriented_box_intersection_2d(...)[0] > 0, the vertex-sort area of oracle/iou.py).
Pinned by tests/golden/traffic.npz (unmodified reference, tests/golden/make_golden.py traffic).
"""
import numpy as np

from . import iou


def rear_corners(box, rear_factor):
    """[P,5] (x, y, length, width, psi) -> [P,4,2]: the part of the box within rear_factor * length of its rear end."""
    box = np.asarray(box)
    x, y, l, w, a = (box[:, i:i + 1] for i in range(5))
    x4 = np.array([0.5, -0.5, -0.5, 0.5], box.dtype) * l * box.dtype.type(rear_factor)
    y4 = np.array([0.5, 0.5, -0.5, -0.5], box.dtype) * w
    s, c = np.sin(a), np.cos(a)
    shift = (l * (1 - box.dtype.type(rear_factor))) / 2          # centre correction along the heading
    cx, cy = x - shift * c, y - shift * s
    return np.stack([x4 * c - y4 * s + cx, x4 * s + y4 * c + cy], -1)


def tl_violation(agent_box, tl_corners, tl_state, red_index=0, rear_factor=0.1, present=None):
    """agent_box [B,A,5], tl_corners [B,L,4,2], tl_state [B,L] -> bool [B,A] (times the present mask if given)."""
    agent_box = np.asarray(agent_box, np.float64)
    tl_corners = np.asarray(tl_corners, np.float64)
    B, A = agent_box.shape[:2]
    L = tl_corners.shape[1]
    out = np.zeros((B, A), bool)
    if B == 0 or A == 0 or L == 0:
        return out
    for b in range(B):
        c1 = rear_corners(agent_box[b], rear_factor)                        # [A,4,2]
        c1 = np.repeat(c1[:, None], L, 1).reshape(A * L, 4, 2)
        c2 = np.repeat(tl_corners[b][None], A, 0).reshape(A * L, 4, 2)
        area = iou.intersection_area(c1, c2).reshape(A, L)
        out[b] = ((area > 0) & (np.asarray(tl_state[b]) == red_index)[None]).any(-1)
    if present is not None:
        out &= np.asarray(present, bool)
    return out
