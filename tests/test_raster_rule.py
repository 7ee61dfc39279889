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
