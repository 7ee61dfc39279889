"""The triangle coverage rule shared by the CUDA kernel (torchdrivesim_b200/csrc/tds_raster_tri.h, compiled
here for the host) and the oracle's restatement of cv2.fillConvexPoly, against the live cv2 module
(opencv-python-headless 4.13 in this image; the reference leaves the version unpinned)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import raster as R

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_rule():
    out = os.path.join(HERE, "_build", "libraster_rule_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I",
                           os.path.join(HERE, "..", "torchdrivesim_b200", "csrc"), "-o", out,
                           os.path.join(HERE, "csrc", "raster_rule_host.cpp")])
    return ctypes.CDLL(out)


def _triangles(rng, res, n, max_abs=None):
    for k in range(n):
        mode = k % 6
        if mode == 0:
            pts = rng.integers(0, res, (3, 2))
        elif mode == 1:
            pts = rng.integers(-res // 2, res + res // 2, (3, 2))
        elif mode == 2:
            pts = rng.integers(-20 * res, 20 * res, (3, 2))
        elif mode == 3:
            pts = rng.integers(-4, res + 4, (1, 2)) + rng.integers(-3, 4, (3, 2))
        elif mode == 4:
            pts = rng.integers(-3, res + 3, (1, 2)) + rng.integers(0, 2, (3, 2))          # within 2x2 pixels
        else:
            pts = rng.integers(-3, res + 3, (1, 2)) + np.stack([rng.integers(-6, 7, 3), rng.integers(0, 2, 3)], 1)
        if max_abs is not None:
            pts = pts.clip(-max_abs, max_abs)
        yield np.ascontiguousarray(pts, np.int32)


def _cv2_mask(pts, res):
    img = np.zeros((res, res, 3), np.float32)
    img = cv2.fillConvexPoly(img=img, points=pts, color=[1, 1, 1], shift=0, lineType=cv2.LINE_AA)
    return img[..., 0] > 0


@pytest.mark.parametrize("res", [16, 64, 256])
def test_oracle_fill_matches_cv2(res):
    rng = np.random.default_rng(res)
    for pts in _triangles(rng, res, 3000):
        assert np.array_equal(R.fill_convex_poly(pts, res, res), _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res", [16, 64, 256, 1024])
def test_kernel_rule_generic_matches_cv2(host_rule, res):
    rng = np.random.default_rng(res + 1)
    for pts in _triangles(rng, res, 2500 if res < 1024 else 600):
        m = np.zeros((res, res), np.uint8)
        host_rule.tds_host_draw_triangle(m.ctypes.data_as(ctypes.c_void_p), res, res, pts.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res", [8, 64, 256])
def test_kernel_rule_fast_path_matches_cv2(host_rule, res):
    rng = np.random.default_rng(res + 2)
    for pts in _triangles(rng, res, 4000, max_abs=8000):
        m = np.zeros((res, res), np.uint8)
        host_rule.tds_host_draw_triangle_fast(m.ctypes.data_as(ctypes.c_void_p), res, res,
                                              pts.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res", [8, 64, 256, 448])
def test_kernel_rule_row_runs_match_cv2(host_rule, res):
    """Closed-form LineIterator runs + fixed-point spans per row (tds_raster_rows.h), the rule of the bitplane kernel."""
    rng = np.random.default_rng(res + 3)
    for pts in _triangles(rng, res, 6000, max_abs=8000):
        m = np.zeros((res, res), np.uint8)
        host_rule.tds_host_draw_triangle_rows(m.ctypes.data_as(ctypes.c_void_p), res, res,
                                              pts.ctypes.data_as(ctypes.c_void_p))
        assert m.max() <= 1, pts.tolist()
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res", [8, 64, 100, 256, 448])
def test_kernel_rule_by_parts_matches_cv2(host_rule, res):
    """row_tri_part (tds_raster_rows.h): the runs of each outline edge and the fill spans as four independent parts -
    the form the kernel runs, one part per lane, for the faces that cross the image border."""
    rng = np.random.default_rng(res + 17)
    for pts in _triangles(rng, res, 8000, max_abs=8000):
        m = np.zeros((res, res), np.uint8)
        host_rule.tds_host_draw_triangle_by_parts(m.ctypes.data_as(ctypes.c_void_p), res, res,
                                                  pts.ctypes.data_as(ctypes.c_void_p))
        assert m.max() <= 1, pts.tolist()
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res", [8, 64, 256, 448])
def test_stateless_row_rule_matches_cv2(host_rule, res):
    """tds_raster_rows_at.h: the intervals of a row from the set-up alone, rows taken bottom-up and twice (the building
    block for spreading the rows of a batch of faces over the lanes of a warp; not used by the kernels yet)."""
    rng = np.random.default_rng(res + 11)
    for pts in _triangles(rng, res, 6000, max_abs=8000):
        m = np.zeros((res, res), np.uint8)
        host_rule.tds_host_draw_triangle_rows_at(m.ctypes.data_as(ctypes.c_void_p), res, res,
                                                 pts.ctypes.data_as(ctypes.c_void_p))
        assert m.max() <= 1, pts.tolist()
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()


@pytest.mark.parametrize("res,small", [(8, 1), (64, 1), (128, 1), (64, 0), (256, 0), (448, 0)])
def test_kernel_rule_inside_fast_path_matches_cv2(host_rule, res, small):
    """Triangles with all vertices inside the image: one interval per row (tds_raster_rows.h, FastTri)."""
    rng = np.random.default_rng(res + 5 + small)
    n = 0
    for k in range(30000):
        mode = k % 4
        if mode == 0:
            pts = rng.integers(0, res, (3, 2))
        elif mode == 1:
            pts = rng.integers(0, res, (1, 2)) + rng.integers(-8, 9, (3, 2))
        elif mode == 2:
            pts = rng.integers(0, res, (1, 2)) + np.stack([rng.integers(-res, res, 3), rng.integers(-2, 3, 3)], 1)
        else:
            pts = rng.integers(0, res, (1, 2)) + np.stack([rng.integers(-2, 3, 3), rng.integers(-res, res, 3)], 1)
        pts = np.ascontiguousarray(pts.clip(0, res - 1), np.int32)
        m = np.zeros((res, res), np.uint8)
        n += host_rule.tds_host_draw_triangle_inside(m.ctypes.data_as(ctypes.c_void_p), res, res,
                                                     pts.ctypes.data_as(ctypes.c_void_p), small)
        assert m.max() <= 1, pts.tolist()
        assert np.array_equal(m > 0, _cv2_mask(pts, res)), pts.tolist()
    assert n == 30000


def test_degenerate_and_collinear():
    for pts in ([[5, 5], [5, 5], [5, 5]], [[0, 0], [10, 10], [20, 20]], [[3, 7], [3, 7], [9, 7]], [[-5, -5], [-1, -1], [-3, -9]]):
        pts = np.array(pts, np.int32)
        assert np.array_equal(R.fill_convex_poly(pts, 32, 32), _cv2_mask(pts, 32))
    assert R.fill_convex_poly(np.array([[5, 5]] * 3, np.int32), 32, 32).sum() == 1


def test_quad_patterns_against_cv2(host_rule):
    """Sliver quads (two faces of a road / lane-marking strip inside the image) drawn from the coverage-pattern table, the way
    the kernel does it, vs the two cv2 triangles: random quads at random positions, both namings of the rungs."""
    rng = np.random.default_rng(11)
    res, accepted = 64, 0
    for k in range(8000):
        a = rng.integers(0, res, 2)
        rung = rng.integers(-17, 18, 2)
        r1, r2 = rng.integers(-1 - (k % 7 == 0), 2 + (k % 7 == 0), 2), rng.integers(-1, 2, 2)
        b = a + rung
        c, d = a + r1, b + r2
        pts = np.ascontiguousarray(np.stack([a, b, c, d] if k % 2 else [a, c, b, d]), np.int32)
        img = np.zeros((res, res), np.uint8)
        naming = host_rule.tds_host_draw_quad(img.ctypes.data_as(ctypes.c_void_p), res, res, pts.ctypes.data_as(ctypes.c_void_p))
        inside = bool(((pts >= 0) & (pts < res)).all())
        in_table = inside and np.abs(rung).max() <= 15 and np.abs(r1).max() <= 1
        assert (naming != 0) == in_table or (naming != 0 and inside), (pts.tolist(), naming)
        if not naming:
            continue
        accepted += 1
        want = _cv2_mask(pts[[0, 1, 2]], res) | _cv2_mask(pts[[1, 2, 3]], res)
        assert img.max() <= 1, "pattern contract violated"
        assert np.array_equal(img > 0, want), pts.tolist()
    assert accepted > 1500


def test_quad_pattern_table_entries_against_cv2(host_rule):
    """Entries of the table itself (generated by the product's triangle rule at start-up) against cv2, at a fixed anchor."""
    shape = (ctypes.c_int * 3)()
    host_rule.tds_host_quad_table_shape(shape)
    n_pat, n_rows, R = shape[0], shape[1], shape[2]
    host_rule.tds_host_quad_table.restype = ctypes.POINTER(ctypes.c_uint32)
    table = np.ctypeslib.as_array(host_rule.tds_host_quad_table(), shape=(n_pat, n_rows))
    assert n_pat == (2 * R + 1) ** 2 * 81
    rng = np.random.default_rng(5)
    span = 2 * R + 1
    for idx in rng.integers(0, n_pat, 3000):
        r2, rest = idx % 9, idx // 9
        r1, rest = rest % 9, rest // 9
        gx, gy = rest % span - R, rest // span - R
        a = np.array([24, 24])
        b = a + [gx, gy]
        c = a + [r1 % 3 - 1, r1 // 3 - 1]
        d = b + [r2 % 3 - 1, r2 // 3 - 1]
        pts = np.stack([a, b, c, d]).astype(np.int32)
        want = _cv2_mask(pts[[0, 1, 2]], 64) | _cv2_mask(pts[[1, 2, 3]], 64)
        x0, y0 = pts[:, 0].min(), pts[:, 1].min()
        got = np.zeros((64, 64), bool)
        for r in range(n_rows):
            for j in range(32):
                if (int(table[idx, r]) >> j) & 1:
                    got[y0 + r, x0 + j] = True
        assert np.array_equal(got, want), (idx, pts.tolist())


@pytest.mark.parametrize("res", [32, 64, 128])
def test_three_line_triangles_match_cv2(host_rule, res):
    """Triangles inside the image as three line walkers (the kernel's path for every face inside the image)."""
    rng = np.random.default_rng(res + 5)
    for k in range(4000):
        if k % 3 == 0:
            pts = rng.integers(0, res, (3, 2))
        elif k % 3 == 1:
            pts = (rng.integers(3, res - 3, (1, 2)) + rng.integers(-3, 4, (3, 2)))
        else:
            pts = (rng.integers(8, res - 8, (1, 2)) + np.stack([rng.integers(-8, 9, 3), rng.integers(0, 2, 3)], 1))
        pts = np.ascontiguousarray(pts, np.int32)
        img = np.zeros((res, res), np.uint8)
        assert host_rule.tds_host_draw_triangle_lines3(img.ctypes.data_as(ctypes.c_void_p), res, res, pts.ctypes.data_as(ctypes.c_void_p)) == 1
        assert img.max() <= 1, "row contract violated"
        assert np.array_equal(img > 0, _cv2_mask(pts, res)), pts.tolist()


def test_three_line_triangles_exhaustive_small(host_rule):
    """All 262 144 triangles with vertices in a 8 x 8 window (every degenerate and collinear case included)."""
    host_rule.tds_host_lines3_triangle_sweep.restype = ctypes.c_longlong
    assert host_rule.tds_host_lines3_triangle_sweep(8) == 0
