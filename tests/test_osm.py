"""Lanelet2 OSM ingestion without lanelet2 (torchdrivesim_b200/osm.py), pinned by the mesh the reference ships for
carla_Town02 (tests/golden/maps/carla_Town02.npz is that mesh; the OSM file is the reference's too)."""
import collections
import os

import numpy as np

import torchdrivesim_b200 as tds
from torchdrivesim_b200 import osm
from tests import util

MAPS = os.path.join(os.path.dirname(__file__), "golden", "maps")


def _triangles(verts, faces, cat_of_face):
    """{category: [F,6] float32} with the vertex order of every face kept."""
    out = collections.defaultdict(list)
    for f, c in zip(faces, cat_of_face):
        out[c].append(verts[f].ravel())
    return {c: np.array(v, np.float32) for c, v in out.items()}


import pytest


@pytest.mark.parametrize("name", ["carla_Town01", "carla_Town02"])
def test_mesh_from_osm_equals_the_shipped_mesh(name):
    m = tds.StaticMap.from_lanelet_osm(os.path.join(MAPS, name + ".osm.gz"), left_handed=True)
    ref = util.load_map_np(name)
    got = _triangles(m.verts, m.faces, m.face_category_names)
    want = _triangles(ref["verts"], ref["faces"], ref["face_cat"])
    assert set(got) == set(want) == {"road", "left_lane", "right_lane"}
    exact = total = 0
    for c in want:
        assert got[c].shape == want[c].shape
        # the same triangles: every built triangle has a partner among the shipped ones and vice versa, same counts.
        # The UTM series of lanelet2 (GeographicLib) and of osm.py agree to ~1e-9 m: a coordinate may round to the
        # neighbouring float32 (<= 1.6e-5 m at 200 m), which happens in a few triangles per thousand
        from scipy.spatial import cKDTree
        dist, idx = cKDTree(want[c].astype(np.float64)).query(got[c].astype(np.float64))
        back, _ = cKDTree(got[c].astype(np.float64)).query(want[c].astype(np.float64))
        assert dist.max() <= 4e-5 and back.max() <= 4e-5        # (the map repeats a few triangles: no bijection test)
        a, b = got[c], want[c][idx]
        assert np.abs(a - b).max() <= 1.6e-5
        exact += int((a.view(np.uint32) == b.view(np.uint32)).all(1).sum())
        total += len(a)
    # measured: 99.7 % of the triangles of Town02 and 97.9 % of Town01 (larger coordinates, coarser float32) are bit-identical
    assert exact >= 0.97 * total, f"only {exact} of {total} triangles are bit-identical"


def test_projection_and_bound_orientation():
    m = osm.LaneletOsm.load(os.path.join(MAPS, "carla_Town02.osm.gz"))
    ref = util.load_map_np("carla_Town02")
    road = ref["verts"][ref["vert_category"] == ref["categories"].index("road")]
    # every point of the OSM file is a vertex of the shipped road mesh (lanelet2.py:214-222)
    from scipy.spatial import cKDTree
    d, _ = cKDTree(m.xy.astype(np.float64)).query(road.astype(np.float64))
    assert len(road) == len(m.xy) and d.max() <= 1.6e-5
    # after the loader's orientation fix the right bound is on the right of the left bound
    for _, lb, rb in m.lanelets:
        a, b, c = m.xy[m.index[lb[0]]], m.xy[m.index[lb[-1]]], m.xy[m.index[rb[0]]]
        assert (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) <= 0
    # a known point: WGS84 (0, 0) is the origin, one degree north is 110 574 m away along the meridian
    e, n = osm._utm_forward(np.array([0.0, 1.0]), np.array([3.0, 3.0]), 3.0)
    assert abs(e[0]) < 1e-9 and abs(n[0]) < 1e-9 and abs(n[1] - 0.9996 * 110574.3886) < 0.01


def test_line_strips_and_joint_markings():
    pts = np.array([[[0, 0], [2, 0]], [[0, 1], [0, 3]]], np.float32)
    v, f = osm.line_segments_to_mesh(pts, line_width=0.5)
    assert v.shape == (12, 2) and f.shape == (8, 3)
    np.testing.assert_allclose(v[:6], [[0, 0.5], [2, 0.5], [0, 0], [2, 0], [0, -0.5], [2, -0.5]], atol=1e-6)
    assert f[4:].min() == 6 and f[:4].tolist() == [[0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5]]
    # two lanelets side by side whose shared boundary is stored twice (different points, same place): a joint marking
    ids = list(range(1, 13))
    xy = np.array([[0, 0], [5, 0], [0, 3], [5, 3], [0, 3.05], [5, 3.05], [0, 6], [5, 6], [0, 0], [0, 0], [0, 0], [0, 0]], np.float32)
    m = osm.LaneletOsm(ids, xy, [(100, [3, 4], [1, 2]), (101, [7, 8], [5, 6])])
    seg = osm.lane_segments(m)
    assert len(seg["joint_lane"]) == 1 and len(seg["left_lane"]) == 1 and len(seg["right_lane"]) == 1
    np.testing.assert_allclose(seg["joint_lane"][0], [[0, 3], [5, 3]])            # kept from the left bound of lanelet 100
    assert len(osm.lane_segments(m, join_threshold=0.01)["joint_lane"]) == 0
    swapped = osm.lane_segments(m, left_handed=True)
    np.testing.assert_allclose(swapped["left_lane"], seg["right_lane"])
    verts, faces, cats, vcat = osm.birdview_mesh_arrays(m)
    assert cats == ["joint_lane", "left_lane", "right_lane", "road"] and len(faces) == 3 * 4 + 2 * 2
    assert faces.max() < len(verts) == len(vcat)


def test_town10hd_fixture_is_what_the_osm_builds():
    """carla_Town10HD ships without a mesh (map.py:61-74 derives it through lanelet2): the npz used as the input of the
    GPU parity tests on that map is exactly what osm.py builds from the reference's OSM file."""
    m = tds.StaticMap.from_lanelet_osm(os.path.join(MAPS, "carla_Town10HD.osm.gz"), left_handed=True,
                                       stoplines_path=os.path.join(MAPS, "carla_Town10HD_stoplines.json"))
    ref = util.load_map_np("carla_Town10HD")
    assert np.array_equal(m.verts, ref["verts"]) and np.array_equal(m.faces, ref["faces"])
    assert m.categories == ref["categories"] == ["joint_lane", "left_lane", "right_lane", "road"]
    assert np.array_equal(m.stoplines, ref["stoplines"]) and len(m.traffic_light_poses()) > 10
    counts = collections.Counter(m.face_category_names)
    assert counts["joint_lane"] > 1000 and counts["road"] > 10000
