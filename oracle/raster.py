"""ORACLE (test infrastructure): birdview raster, restating the reference cv2 backend.

Scene assembly follows torchdrivesim/mesh.py:1053-1157 (BirdviewRGBMeshGenerator.generate),
the per-camera pipeline torchdrivesim/rendering/cv2.py:27-70 (implemented in c/raster_oracle.c).
"""
import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import clib

f32 = np.float32

# rendering/base.py:234-261 (levels) and :264-292 (colours)
DEFAULT_LEVELS = dict(
    direction=2, ego=3, vehicle=4, bicycle=5, pedestrian=6, map_boundary=7, goal_waypoint=8,
    ground_truth=9, prediction=10, traffic_light=11, traffic_light_green=11, traffic_light_yellow=11,
    traffic_light_red=11, stop_sign=11, yield_sign=11, left_lane=12, joint_lane=13, right_lane=14, road=15,
)
DEFAULT_COLORS = dict(
    background=(0, 0, 0), road=(155, 155, 155), corridor=(0, 155, 0), ego=(255, 0, 0), vehicle=(32, 74, 135),
    bicycle=(24, 104, 225), pedestrian=(173, 127, 168), ground_truth=(196, 188, 165), prediction=(255, 155, 0),
    left_lane=(80, 127, 86), right_lane=(128, 0, 128), joint_lane=(255, 255, 255), direction=(100, 255, 255),
    rear_lights=(255, 255, 0), map_boundary=(255, 255, 0), traffic_light_green=(81, 179, 100),
    traffic_light_yellow=(240, 189, 39), traffic_light_red=(224, 53, 49), yield_sign=(210, 125, 45),
    stop_sign=(72, 60, 50), goal_waypoint=(139, 64, 0),
)
# Draw order among categories of EQUAL z is undefined in the reference (torch.argsort is not
# stable, rendering/cv2.py:47).  The build fixes it: later in this list = drawn later = on top.
TIE_ORDER = ["stop_sign", "yield_sign", "traffic_light", "traffic_light_green", "traffic_light_yellow",
             "traffic_light_red"]


def category_ranks(levels: Dict[str, float] = None) -> Dict[str, int]:
    """rank = draw order (0 drawn first).  Descending z, ties by TIE_ORDER then name."""
    levels = DEFAULT_LEVELS if levels is None else levels
    names = sorted(levels, key=lambda k: (-levels[k], TIE_ORDER.index(k) if k in TIE_ORDER else -1, k))
    return {k: i for i, k in enumerate(names)}


def cr_sincos(psi: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """sin/cos evaluated in float64 and rounded to fp32 (what the CUDA path does as well)."""
    p = np.asarray(psi, dtype=f32).astype(np.float64)
    return np.sin(p).astype(f32), np.cos(p).astype(f32)


def quantize_color(rgb255: Sequence[int]) -> np.ndarray:
    """rendering/cv2.py:50 applied to mesh.py tensor_color(rgb)/255: floor(c/255*0.999*256)."""
    c = np.asarray(rgb255, dtype=f32) / f32(255.0)
    return np.floor(c * f32(1.0 - 1e-3) * f32(256)).astype(np.uint8)


@dataclass
class Scene:
    """Triangle soup of one environment (all cameras of the env share it)."""
    verts: np.ndarray       # [V,2] f32 world
    faces: np.ndarray       # [F,3] i32
    face_rank: np.ndarray   # [F] u8
    face_rgb: np.ndarray    # [F,3] u8


def agent_world_verts(state: np.ndarray, size: np.ndarray) -> np.ndarray:
    """[N,7,2] world vertices of the agent rectangle (4) and direction triangle (3).

    mesh.py:942-951 (corners), :911-940 (direction, size=0.3), utils.py:40-96 (transform).
    """
    state = np.asarray(state, f32)
    size = np.asarray(size, f32)
    l, w = size[:, 0], size[:, 1]
    half = f32(0.5)
    rect = np.stack([np.stack([l, w], -1), np.stack([l, -w], -1), np.stack([-l, -w], -1), np.stack([-l, w], -1)],
                    axis=1) * half                                      # [N,4,2]
    off = l * f32(0.5 - 0.3)                                            # python double 0.2 -> fp32 mult
    tri = np.stack([np.stack([l * f32(0.3) + off, np.zeros_like(l) + f32(0) * l], -1),
                    np.stack([np.zeros_like(l) + off, w * half], -1),
                    np.stack([np.zeros_like(l) + off, -w * half], -1)], axis=1)  # [N,3,2]
    tri[:, 0, 1] = 0.0
    local = np.concatenate([rect, tri], axis=1).astype(f32)             # [N,7,2]
    s, c = cr_sincos(state[:, 2])
    s, c = s[:, None], c[:, None]
    x = (c * local[..., 0]) + ((-s) * local[..., 1])
    y = (s * local[..., 0]) + (c * local[..., 1])
    return np.stack([x + state[:, 0:1], y + state[:, 1:2]], -1).astype(f32)


def box_corners(box: np.ndarray) -> np.ndarray:
    """_iou_utils.py:270-299 box2corners_th for [..,5] (x,y,l,w,psi) -> [..,4,2]."""
    box = np.asarray(box, f32)
    x, y, l, w = box[..., 0:1], box[..., 1:2], box[..., 2:3], box[..., 3:4]
    s, c = cr_sincos(box[..., 4:5])
    x4 = np.array([0.5, -0.5, -0.5, 0.5], f32) * l
    y4 = np.array([0.5, 0.5, -0.5, -0.5], f32) * w
    X = (x4 * c) + (y4 * (-s))
    Y = (x4 * s) + (y4 * c)
    return np.stack([X + x, Y + y], -1).astype(f32)


def disc_template(radius: float = 2.0, num_triangles: int = 10):
    """The waypoint disc of generate_disc_mesh (mesh.py:1243-1271): centre + rim vertices produced by REPEATED
    rotation of (radius, 0) with torch.matmul (utils.py:36-69), faces (0, k, k+1) and finally (0, n, 1).
    Evaluated with the same ATen ops as the reference (the accumulated fp32 rounding is part of the result)."""
    import torch
    step = torch.deg2rad(torch.tensor([[360 / num_triangles]], dtype=torch.float32))
    c, s = torch.cos(step), torch.sin(step)
    rot = torch.stack([torch.cat([c, -s], -1), torch.cat([s, c], -1)], -2)
    verts = [torch.zeros(1, 2), torch.tensor([[radius, 0.0]], dtype=torch.float32)]
    for _ in range(num_triangles - 1):
        verts.append(torch.matmul(rot, verts[-1].unsqueeze(-1)).squeeze(-1))
    faces = [[0, k, k + 1] for k in range(1, num_triangles)] + [[0, num_triangles, 1]]
    return torch.cat(verts, 0).numpy().astype(f32), np.array(faces, np.int32)


def build_scene(static_verts: np.ndarray, static_faces: np.ndarray, static_face_cat: Sequence[str],
                agent_state: Optional[np.ndarray] = None, agent_size: Optional[np.ndarray] = None,
                agent_type_names: Optional[List[str]] = None, agent_types: Optional[np.ndarray] = None,
                present: Optional[np.ndarray] = None,
                tl_corners: Optional[np.ndarray] = None, tl_state: Optional[np.ndarray] = None,
                tl_allowed_states: Sequence[str] = ("red", "yellow", "green"),
                static_controls: Optional[Dict[str, np.ndarray]] = None,
                levels: Dict[str, float] = None, colors: Dict[str, Tuple[int, int, int]] = None,
                waypoints: Optional[np.ndarray] = None, waypoints_mask: Optional[np.ndarray] = None,
                agent_colors: Optional[np.ndarray] = None) -> Scene:
    """Restates generate() for one environment and one rendering mask.

    static_face_cat: per static face category NAME (the category of its first vertex, cv2.py:58).
    present: [N] bool rendering mask; absent agents collapse to a degenerate triangle at actor
    vertex 0 with agent 0's colour / level (mesh.py:1083-1089).
    Concatenation order (mesh.py:1147-1157): background, actors, static controls, traffic lights, waypoints.
    waypoints [M,2] (+ mask [M]) are those of ONE camera (mesh.py:1120-1145): a disc per waypoint, translated (the
    pose has psi = 0, so the rotation is the identity); the faces of a masked waypoint collapse onto vertex 0 of the
    camera's waypoint mesh, the centre of its first waypoint.
    agent_colors [N,3] in [0,1] are the custom colours of ONE camera (mesh.py:1092-1099): they replace the colour of
    the four rectangle vertices of every agent (the degenerate face of an absent agent is at actor vertex 0, so it
    takes agent 0's custom colour); the level stays that of the agent's type, the direction triangle keeps its colour.
    """
    levels = DEFAULT_LEVELS if levels is None else levels
    colors = DEFAULT_COLORS if colors is None else colors
    ranks = category_ranks(levels)
    verts = [np.asarray(static_verts, f32).reshape(-1, 2)]
    faces = [np.asarray(static_faces, np.int32).reshape(-1, 3)]
    cats = list(static_face_cat)
    custom_rgb = {}                       # face index -> quantized colour (rendering/cv2.py:50)
    nv = verts[0].shape[0]
    if agent_state is not None and len(agent_state) > 0:
        n = len(agent_state)
        agent_type_names = agent_type_names or ["vehicle"]
        agent_types = np.zeros(n, np.int64) if agent_types is None else np.asarray(agent_types)
        present = np.ones(n, bool) if present is None else np.asarray(present, bool)
        av = agent_world_verts(agent_state, agent_size).reshape(n * 7, 2)
        verts.append(av)
        qcol = None
        if agent_colors is not None:
            qcol = np.floor(np.asarray(agent_colors, f32) * f32(1.0 - 1e-3) * f32(256)).clip(0, 255).astype(np.uint8)
        for k in range(n):
            b = nv + 7 * k
            if present[k]:
                faces.append(np.array([[b + 0, b + 1, b + 3], [b + 1, b + 3, b + 2], [b + 4, b + 5, b + 6]], np.int32))
                if qcol is not None:
                    custom_rgb[len(cats)] = qcol[k]
                    custom_rgb[len(cats) + 1] = qcol[k]
                cats += [agent_type_names[agent_types[k]]] * 2 + ["direction"]
            else:
                faces.append(np.full((3, 3), nv, np.int32))
                if qcol is not None:
                    for i in range(3):
                        custom_rgb[len(cats) + i] = qcol[0]
                cats += [agent_type_names[agent_types[0]]] * 3
        nv += n * 7
    for name, corners in (static_controls or {}).items():
        corners = np.asarray(corners, f32).reshape(-1, 4, 2)
        for k in range(corners.shape[0]):
            verts.append(corners[k])
            faces.append(np.array([[nv, nv + 1, nv + 3], [nv + 1, nv + 3, nv + 2]], np.int32))
            cats += [name] * 2
            nv += 4
    if tl_corners is not None and len(tl_corners) > 0:
        tl_corners = np.asarray(tl_corners, f32).reshape(-1, 4, 2)
        for k in range(tl_corners.shape[0]):
            verts.append(tl_corners[k])
            faces.append(np.array([[nv, nv + 1, nv + 3], [nv + 1, nv + 3, nv + 2]], np.int32))
            cats += ["traffic_light_" + tl_allowed_states[int(tl_state[k])]] * 2
            nv += 4
    if waypoints is not None and len(waypoints) > 0:
        waypoints = np.asarray(waypoints, f32).reshape(-1, 2)
        wmask = np.ones(len(waypoints), bool) if waypoints_mask is None else np.asarray(waypoints_mask, bool)
        dv, df = disc_template()
        first = nv
        for k in range(waypoints.shape[0]):
            verts.append((dv + waypoints[k][None]).astype(f32))
            faces.append((df + nv) if wmask[k] else np.full_like(df, first))
            cats += ["goal_waypoint"] * df.shape[0]
            nv += dv.shape[0]
    verts = np.concatenate(verts, 0).astype(f32)
    faces = np.concatenate(faces, 0).astype(np.int32)
    lut = {c: quantize_color(colors[c]) for c in set(cats)}
    face_rank = np.array([ranks[c] for c in cats], np.uint8)
    face_rgb = np.stack([lut[c] for c in cats]).astype(np.uint8) if cats else np.zeros((0, 3), np.uint8)
    for i, c in custom_rgb.items():
        face_rgb[i] = c
    return Scene(verts=np.ascontiguousarray(verts), faces=np.ascontiguousarray(faces),
                 face_rank=face_rank, face_rgb=np.ascontiguousarray(face_rgb))


def render_camera(scene: Scene, cam_xy, cam_sc, res: int = 64, fov: float = 35.0) -> Tuple[np.ndarray, int]:
    """One camera -> ([3,H,W] f32 image in [0,255], faces kept by the cull).  cam_sc = (sin, cos)."""
    out = np.zeros((3, res, res), f32)
    lib = clib()
    kept = lib.oracle_render_camera(
        scene.verts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(scene.verts.shape[0]),
        scene.faces.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(scene.faces.shape[0]),
        scene.face_rank.ctypes.data_as(ctypes.c_void_p), scene.face_rgb.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_float(f32(cam_xy[0])), ctypes.c_float(f32(cam_xy[1])),
        ctypes.c_float(f32(cam_sc[0])), ctypes.c_float(f32(cam_sc[1])), ctypes.c_float(f32(2.0 / fov)),
        ctypes.c_int(res), ctypes.c_int(res), out.ctypes.data_as(ctypes.c_void_p))
    return out, kept


def fill_convex_poly(pts: np.ndarray, res_w: int, res_h: int) -> np.ndarray:
    """Coverage mask [H,W] bool of one polygon under the restated cv2.fillConvexPoly rule."""
    img = np.zeros((res_h, res_w), np.int32)
    pts = np.ascontiguousarray(pts, np.int32)
    clib().oracle_fill_convex_poly(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(res_w), ctypes.c_int(res_h),
                                   pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(pts.shape[0]), ctypes.c_int(1))
    return img.astype(bool)
