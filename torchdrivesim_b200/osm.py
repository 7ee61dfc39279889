"""Lanelet2 OSM maps without the lanelet2 package: OSM file -> road / lane-marking mesh -> `StaticMap`.

Restates, for the maps that ship without a pre-built mesh (carla_Town07, carla_Town10HD: `MapConfig.road_mesh`,
torchdrivesim/map.py:61-74), what the reference does through the lanelet2 library:

  load_lanelet_map               torchdrivesim/lanelet2.py:88-105   lanelet2.io.load with a UtmProjector(Origin(lat, lon))
  road_mesh_from_lanelet_map     torchdrivesim/lanelet2.py:205-250  every lanelet triangulated between its two bounds
  lanelet_map_to_lane_mesh       torchdrivesim/lanelet2.py:286-379  boundary segments -> 6-vertex / 4-face strips
  line_segments_to_mesh          torchdrivesim/lanelet2.py:253-283

Third-party behaviour restated here (lanelet2 is absent from this image; pinned by the meshes the reference ships for
carla_Town01 / carla_Town02, which this module reproduces triangle for triangle, bit for bit - tests/test_osm.py):
  * UtmProjector: transverse Mercator (Krueger series to n^6, WGS84, k0 = 0.9996) in the UTM zone of the ORIGIN, minus
    the origin's own easting / northing;
  * the OSM loader turns the `left` / `right` members of a `type=lanelet` relation into the bounds and, when the
    right bound is not on the right-hand side of the left bound, inverts both (every lanelet of the left-handed CARLA
    maps is in that case).
This is offline data ingestion (host side, numpy); nothing here is on the per-step path.
"""
import gzip
import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_A, _F, _K0 = 6378137.0, 1.0 / 298.257223563, 0.9996


def _utm_forward(lat_deg: np.ndarray, lon_deg: np.ndarray, lon0_deg: float) -> Tuple[np.ndarray, np.ndarray]:
    """Easting (without the 500 km false easting) and northing of WGS84 points on the meridian strip of lon0."""
    n = _F / (2 - _F)
    big_a = _A / (1 + n) * (1 + n ** 2 / 4 + n ** 4 / 64 + n ** 6 / 256)
    alpha = [n / 2 - 2 * n ** 2 / 3 + 5 * n ** 3 / 16 + 41 * n ** 4 / 180 - 127 * n ** 5 / 288 + 7891 * n ** 6 / 37800,
             13 * n ** 2 / 48 - 3 * n ** 3 / 5 + 557 * n ** 4 / 1440 + 281 * n ** 5 / 630 - 1983433 * n ** 6 / 1935360,
             61 * n ** 3 / 240 - 103 * n ** 4 / 140 + 15061 * n ** 5 / 26880 + 167603 * n ** 6 / 181440,
             49561 * n ** 4 / 161280 - 179 * n ** 5 / 168 + 6601661 * n ** 6 / 7257600,
             34729 * n ** 5 / 80640 - 3418889 * n ** 6 / 1995840,
             212378941 * n ** 6 / 319334400]
    phi, lam = np.radians(np.asarray(lat_deg, np.float64)), np.radians(np.asarray(lon_deg, np.float64) - lon0_deg)
    e = math.sqrt(_F * (2 - _F))
    t = np.sinh(np.arctanh(np.sin(phi)) - e * np.arctanh(e * np.sin(phi)))
    xi, eta = np.arctan2(t, np.cos(lam)), np.arctanh(np.sin(lam) / np.sqrt(1 + t * t))
    east, north = eta.copy(), xi.copy()
    for j, a in enumerate(alpha, start=1):
        east += a * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        north += a * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return _K0 * big_a * east, _K0 * big_a * north


class LaneletOsm:
    """The part of a lanelet2 map that the meshes need: projected points and the two bounds of every lanelet."""

    def __init__(self, point_ids: List[int], xy: np.ndarray, lanelets: List[Tuple[int, List[int], List[int]]]):
        self.point_ids = point_ids                   # in file order
        self.xy = xy                                 # [P,2] float32, as the reference's np.float32 vertex array
        self.index = {p: i for i, p in enumerate(point_ids)}
        self.lanelets = lanelets                     # (id, left bound, right bound) as point ids, oriented

    @classmethod
    def load(cls, path: str, origin: Tuple[float, float] = (0.0, 0.0)) -> "LaneletOsm":
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            root = ET.parse(f).getroot()
        ids, lat, lon = [], [], []
        for nd in root.iter("node"):
            ids.append(int(nd.get("id")))
            lat.append(float(nd.get("lat")))
            lon.append(float(nd.get("lon")))
        zone = int(math.floor((origin[1] + 180.0) / 6.0)) % 60 + 1
        lon0 = zone * 6.0 - 183.0
        east, north = _utm_forward(np.array(lat), np.array(lon), lon0)
        e0, n0 = _utm_forward(np.array([origin[0]]), np.array([origin[1]]), lon0)
        xy = np.stack([east - e0[0], north - n0[0]], -1).astype(np.float32)
        ways = {int(w.get("id")): [int(n.get("ref")) for n in w.findall("nd")] for w in root.iter("way")}
        index = {p: i for i, p in enumerate(ids)}
        lanelets = []
        for rel in root.iter("relation"):
            tags = {t.get("k"): t.get("v") for t in rel.findall("tag")}
            if tags.get("type") != "lanelet":
                continue
            bounds = {}
            for role in ("left", "right"):
                pts: List[int] = []
                for m in rel.findall("member"):
                    if m.get("type") == "way" and m.get("role") == role:
                        w = ways[int(m.get("ref"))]
                        pts += w[1:] if pts and pts[-1] == w[0] else w     # consecutive ways share their end point
                bounds[role] = pts
            lb, rb = bounds["left"], bounds["right"]
            if len(lb) >= 2 and len(rb) >= 1:
                # left must be left of right: otherwise the loader inverts both bounds
                a, b, c = xy[index[lb[0]]].astype(np.float64), xy[index[lb[-1]]].astype(np.float64), xy[index[rb[0]]].astype(np.float64)
                if (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) > 0.0:
                    lb, rb = lb[::-1], rb[::-1]
            lanelets.append((int(rel.get("id")), lb, rb))
        return cls(ids, xy, lanelets)


def road_mesh(m: LaneletOsm) -> Tuple[np.ndarray, np.ndarray]:
    """(verts [P,2] f32 = every point of the map, faces [F,3]) - lanelet2.py:205-250."""
    faces = []
    for _, lb, rb in m.lanelets:
        n_faces = len(lb) + len(rb) - 2
        if n_faces < 1:
            continue
        out = np.zeros((n_faces, 3), np.int64)
        i = j = 0
        while i + j < n_faces:
            if i < len(lb) - 1:
                out[i + j] = [m.index[lb[i]], m.index[rb[j]], m.index[lb[i + 1]]]
                i += 1
            if j < len(rb) - 1:
                if i + j < n_faces:              # the reference writes one row past the end here and numpy raises; the
                    out[i + j] = [m.index[lb[i]], m.index[rb[j]], m.index[rb[j + 1]]]   # shipped maps never get there
                j += 1
        faces.append(out)
    return m.xy.copy(), (np.concatenate(faces) if faces else np.zeros((0, 3), np.int64))


def line_segments_to_mesh(points: np.ndarray, line_width: float = 0.3, eps: float = 1e-6) -> Tuple[np.ndarray, np.ndarray]:
    """points [N,2,2] f32 -> (verts [6N,2], faces [4N,3]): lanelet2.py:253-283 in float32, operation by operation."""
    p = np.asarray(points, np.float32).reshape(-1, 2, 2)
    d = p[:, 1] - p[:, 0]
    norm = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1], dtype=np.float32)[:, None]
    d_hat = d / (norm + np.float32(eps))
    perp = np.stack([-d_hat[:, 1], d_hat[:, 0]], -1)[:, None]
    w = np.float32(line_width)
    verts = np.concatenate([p + perp * w, p, p - perp * w], 1).reshape(-1, 2).astype(np.float32)
    faces = (np.array([[0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5]], np.int64)[None] + 6 * np.arange(p.shape[0])[:, None, None]).reshape(-1, 3)
    return verts, faces


def lane_segments(m: LaneletOsm, left_handed: bool = False, join_threshold: float = 0.1) -> Dict[str, np.ndarray]:
    """Boundary segments by marking category, each [N,2,2] f32 - lanelet2.py:309-358.  A left segment whose two end
    points lie within `join_threshold` of the end points of some right segment is a joint marking (kept from the left)."""
    left, right = set(), set()
    for _, lb, rb in m.lanelets:
        for i in range(len(rb) - 1):
            right.add(tuple(sorted([rb[i], rb[i + 1]])))
        for i in range(len(lb) - 1):
            left.add(tuple(sorted([lb[i], lb[i + 1]])))
    seg = lambda s: np.array([[m.xy[m.index[a]], m.xy[m.index[b]]] for a, b in s], np.float32).reshape(-1, 2, 2)
    lp, rp = seg(left), seg(right)
    l_joint, r_joint = np.zeros(len(lp), bool), np.zeros(len(rp), bool)
    if len(lp) and len(rp):
        from scipy.spatial import cKDTree

        def near(a, b):                 # sets of indices of b within the threshold of each point of a (cdist < thr)
            tree = cKDTree(b.astype(np.float64))
            return [set(x) for x in tree.query_ball_point(a.astype(np.float64), join_threshold * (1 - 1e-12))]
        n00, n11, n01, n10 = near(lp[:, 0], rp[:, 0]), near(lp[:, 1], rp[:, 1]), near(lp[:, 0], rp[:, 1]), near(lp[:, 1], rp[:, 0])
        for i in range(len(lp)):
            hits = (n00[i] & n11[i]) | (n01[i] & n10[i])
            if hits:
                l_joint[i] = True
                r_joint[list(hits)] = True
    out = {"joint_lane": lp[l_joint], "left_lane": lp[~l_joint], "right_lane": rp[~r_joint]}
    if left_handed:
        out["left_lane"], out["right_lane"] = out["right_lane"], out["left_lane"]
    return out


def birdview_mesh_arrays(m: LaneletOsm, left_handed: bool = False, join_threshold: float = 0.1,
                         lane_boundary_width: float = 0.275):
    """(verts [V,2] f32, faces [F,3] i32, categories, vert_category [V]) of `lane_mesh.merge(road_mesh)` (map.py:66-72):
    joint, left and right marking strips followed by the road."""
    cats, verts, faces, vcat = [], [], [], []
    offset = 0
    segs = lane_segments(m, left_handed=left_handed, join_threshold=join_threshold)
    parts = [(c, *line_segments_to_mesh(segs[c], lane_boundary_width)) for c in ("joint_lane", "left_lane", "right_lane") if len(segs[c])]
    parts.append(("road", *road_mesh(m)))
    for c, v, f in parts:
        cats.append(c)
        verts.append(v)
        faces.append(f + offset)
        vcat.append(np.full(len(v), len(cats) - 1, np.int64))
        offset += len(v)
    return (np.concatenate(verts).astype(np.float32), np.concatenate(faces).astype(np.int32), cats, np.concatenate(vcat))
