"""Import the UNMODIFIED reference: /root/reference in the build container, or its offline install under
baseline/_ref (baseline/install_ref.sh; git-ignored, travels to the GPU box with the repository snapshot).

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to generate the committed golden vectors), by
the `needs_reference` tests (the drop-in seam behind the real reference Simulator, on the CPU here and on the
B200 with the real kernels) and by bench.py's reference arm / `cpu_baseline.reference_python` leg.  Nothing on
the product path imports this.

The three stub modules follow SURVEY.md App. D: omegaconf, shapely.geometry and
lanelet2.core are absent from the image and are only touched off the hot path.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    for cand in (os.environ.get("TDS_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "torchdrivesim", "resources", "maps")):
            return cand
    return os.environ.get("TDS_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torchdrivesim"))


def import_reference():
    """Returns the imported `torchdrivesim` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        om.DictConfig = type("DictConfig", (), {})
        om.OmegaConf = type("OmegaConf", (), {})
        om.SCMode = type("SCMode", (), {})
        sys.modules["omegaconf"] = om
    if "shapely" not in sys.modules:
        sh = types.ModuleType("shapely")
        shg = types.ModuleType("shapely.geometry")
        shg.Polygon = object
        sh.geometry = shg
        sys.modules["shapely"] = sh
        sys.modules["shapely.geometry"] = shg
    if "lanelet2" not in sys.modules:
        ll = types.ModuleType("lanelet2")
        ll.core = types.ModuleType("lanelet2.core")
        ll.core.LaneletMap = object
        sys.modules["lanelet2"] = ll
        sys.modules["lanelet2.core"] = ll.core
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torchdrivesim  # noqa: F401
    import torchdrivesim.simulator  # noqa: F401
    import torchdrivesim.rendering  # noqa: F401
    import torchdrivesim.infractions  # noqa: F401
    import torchdrivesim.map  # noqa: F401
    return torchdrivesim
