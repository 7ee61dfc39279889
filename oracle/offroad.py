"""ORACLE (test infrastructure): offroad infraction, brute force over all faces.

torchdrivesim/infractions.py:176-229 offroad_infraction_loss (pure-torch branch) and
infractions.py:86-173 point_to_mesh_distance_pt, implemented in c/offroad_oracle.c.
Aggregation: simulator.py:1035-1044 (multiply by the present mask).
"""
import ctypes

import numpy as np

from . import clib
from .raster import box_corners

f32 = np.float32


def points_mesh_dist2(points, verts, faces):
    """points [P,2] -> (d2 [P] f32 min squared distance over faces, face [P] i32 argmin)."""
    points = np.ascontiguousarray(points, f32).reshape(-1, 2)
    verts = np.ascontiguousarray(verts, f32).reshape(-1, 2)
    faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
    d2 = np.zeros(points.shape[0], f32)
    face = np.zeros(points.shape[0], np.int32)
    clib().oracle_points_mesh_dist2(points.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(points.shape[0]),
                                    verts.ctypes.data_as(ctypes.c_void_p), faces.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_int(faces.shape[0]), d2.ctypes.data_as(ctypes.c_void_p),
                                    face.ctypes.data_as(ctypes.c_void_p))
    return d2, face


def offroad_loss(state, lenwid, verts, faces, threshold=0.0, present=None, return_corners=False):
    """state [A,4], lenwid [A,2] for ONE environment with mesh (verts [V,2], faces [F,3]) -> [A] f32."""
    state = np.asarray(state, f32)
    lenwid = np.asarray(lenwid, f32)
    a = state.shape[0]
    if a == 0 or len(faces) == 0:
        return np.zeros(a, f32)
    box = np.concatenate([state[:, :2], lenwid, state[:, 2:3]], -1)
    corners = box_corners(box)                                   # [A,4,2]
    d2, _ = points_mesh_dist2(corners.reshape(-1, 2), verts, faces)
    d2t = np.where(d2 > f32(threshold), d2, f32(0)).reshape(a, 4)   # F.threshold(x, thr, 0)
    out = ((d2t[:, 0] + d2t[:, 1]) + d2t[:, 2]) + d2t[:, 3]
    if present is not None:
        out = out * np.asarray(present, f32)
    if return_corners:
        return out.astype(f32), d2.reshape(a, 4)
    return out.astype(f32)
