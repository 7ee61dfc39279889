"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's per-step hot path (kinematic step, cv2 birdview
raster, disc / IoU collisions, point-to-mesh offroad).  It is the checker for the
CUDA path and the CPU baseline of bench.py; it is never the product.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import anything from this package.  The product package (`torchdrivesim_b200`)
never imports it and fails loudly when its CUDA library is missing.

Parity pinning: the reference's own tests hold NO golden vectors for this path
(SURVEY.md §4, §8c), so the oracle is pinned against outputs of the unmodified reference
run in the build container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`) and,
for the third-party raster rule (OpenCV fillConvexPoly), against the live `cv2` module.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_clib(force: bool = False) -> str:
    """Compiles oracle/c/*.c into oracle/_build/liboracle.so (gcc, no contraction)."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, "c", f) for f in ("raster_oracle.c", "offroad_oracle.c")]
    stale = force or not os.path.exists(out) or any(
        os.path.getmtime(s) > os.path.getmtime(out) for s in srcs if os.path.exists(s))
    if stale:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-o", out] + srcs + ["-lm"])
    return out


def clib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_clib())
        _LIB.oracle_point_tri_dist2.restype = ctypes.c_float
        _LIB.oracle_render_camera.restype = ctypes.c_int
    return _LIB
