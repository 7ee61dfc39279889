"""Shared helpers for the tests: fixtures loading, seeded scene generators, oracle drivers."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEH = (4.97, 2.04, 1.96)   # length, width, lr  (reference behavior/heuristic.py:11-13)
PED = (1.5, 1.5)           # reference examples/imitation_learning.py:80-81


def map_path(name):
    return os.path.join(GOLDEN, "maps", f"{name}.npz")


def load_map_np(name):
    d = np.load(map_path(name))
    cats = [str(c) for c in d["categories"]]
    fcat = [cats[i] for i in d["vert_category"][d["faces"][:, 0]]]
    return dict(verts=d["verts"], faces=d["faces"], face_cat=fcat, categories=cats, vert_category=d["vert_category"],
                stoplines=d["stoplines"], stopline_types=[str(s) for s in d["stopline_types"]])


def golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


def road_points(m, n, rng):
    road = m["verts"][m["vert_category"] == m["categories"].index("road")]
    return road[rng.integers(0, road.shape[0], n)]


def random_scene(m, B, A, rng, spread=10.0, ped_every=0, absent_p=0.0):
    """Agents clustered around a random on-road point per environment (so that they see each other)."""
    centre = road_points(m, B, rng)[:, None, :]
    xy = centre + spread * rng.standard_normal((B, A, 2))
    psi = rng.uniform(0, 2 * np.pi, (B, A, 1))
    v = rng.uniform(0, 5, (B, A, 1))
    state = np.concatenate([xy, psi, v], -1).astype(np.float32)
    types = np.zeros((B, A), np.int64)
    if ped_every:
        types[:, ped_every - 1::ped_every] = 1
    size = np.where(types[..., None] == 1, np.array(PED, np.float32), np.array(VEH[:2], np.float32)).astype(np.float32)
    present = rng.uniform(size=(B, A)) >= absent_p
    return state, size, types, present


def tl_tensors(m, B, rng):
    """Traffic-light corners [B,L,4,2] (fp32, via torch like the reference) and random states [B,L]."""
    from torchdrivesim_b200.traffic_controls import box2corners
    idx = [i for i, t in enumerate(m["stopline_types"]) if t == "traffic_light"]
    pos = torch.tensor(m["stoplines"][idx])[None].expand(B, -1, -1).contiguous()
    corners = box2corners(pos).numpy()
    state = rng.integers(0, 3, (B, len(idx)))
    return pos.numpy(), corners, state


def oracle_render_batch(m, state, size, types, present, type_names, tl_corners, tl_state, cam_xy, cam_sc, res, fov,
                        cams=None, waypoints=None, waypoints_mask=None, agent_colors=None):
    """Oracle images for the cameras in `cams` (list of (b, c)); present [B,N] or [B,Nc,N]."""
    from oracle import raster as R
    out = {}
    B = state.shape[0]
    for (b, c) in (cams if cams is not None else [(b, c) for b in range(B) for c in range(cam_xy.shape[1])]):
        pr = present[b, c] if present.ndim == 3 else present[b]
        sc = R.build_scene(m["verts"], m["faces"], m["face_cat"], state[b], size[b], type_names, types[b], pr,
                           tl_corners=None if tl_corners is None else tl_corners[b],
                           tl_state=None if tl_state is None else tl_state[b],
                           waypoints=None if waypoints is None else waypoints[b, c],
                           waypoints_mask=None if waypoints_mask is None else waypoints_mask[b, c],
                           agent_colors=None if agent_colors is None else agent_colors[b, c])
        img, _ = R.render_camera(sc, cam_xy[b, c], cam_sc[b, c], res, fov)
        out[(b, c)] = img
    return out


def with_extra_static(m, extra_verts, extra_faces, category):
    """Map dict `m` + extra static triangles of one category appended behind the background (add_static_meshes)."""
    nv = m["verts"].shape[0]
    cats = list(m["categories"]) + ([category] if category not in m["categories"] else [])
    out = dict(m)
    out["verts"] = np.concatenate([m["verts"][:, :2], np.asarray(extra_verts, np.float32)])
    out["faces"] = np.concatenate([m["faces"], np.asarray(extra_faces, np.int32) + nv]).astype(np.int32)
    out["face_cat"] = list(m["face_cat"]) + [category] * len(extra_faces)
    out["categories"] = cats
    out["vert_category"] = np.concatenate([m["vert_category"], np.full(len(extra_verts), cats.index(category))])
    return out
