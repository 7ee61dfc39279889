"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to generate the
committed golden vectors) and by the optional `needs_reference` tests.  Nothing
on the product path, in `-m gpu` tests, smoke() or bench.py may import this:
/root/reference does not exist on the GPU box.

The three stub modules follow SURVEY.md App. D: omegaconf, shapely.geometry and
lanelet2.core are absent from the image and are only touched off the hot path.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TDS_REFERENCE_ROOT", "/root/reference")


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
