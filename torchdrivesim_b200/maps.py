"""Static maps: the triangle mesh of a location plus its device-side acceleration grids.

Loading mirrors the reference's data formats: `*_mesh.json` in the BirdviewMesh schema
(torchdrivesim/mesh.py:700-719), stoplines json (torchdrivesim/map.py:203-229), or the compact npz
fixtures under tests/golden/maps.  A `StaticMap` replaces the batch-expanded `road_mesh`
(`mesh.expand(B)`, simulator.py) by ONE copy per GPU shared by all environments through `env_map`.
"""
import ctypes
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .palette import class_id


class StaticMap:
    """verts [V,2] f32, faces [F,3] i32, per-face category name index; device handles are created lazily
    (one per CUDA device) by `handle(device)`."""

    def __init__(self, verts: np.ndarray, faces: np.ndarray, categories: Sequence[str], vert_category: np.ndarray,
                 name: str = "map", left_handed: bool = False, stoplines: Optional[np.ndarray] = None,
                 stopline_types: Optional[Sequence[str]] = None, raster_cell: float = 16.0, offroad_cell: float = 2.0):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 2)
        self.faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
        self.categories = [str(c) for c in categories]
        self.vert_category = np.ascontiguousarray(vert_category).reshape(-1).astype(np.int64)
        self.name = name
        self.left_handed = bool(left_handed)
        self.stoplines = np.zeros((0, 5), np.float32) if stoplines is None else np.asarray(stoplines, np.float32)
        self.stopline_types = [] if stopline_types is None else [str(s) for s in stopline_types]
        self.raster_cell = raster_cell
        self.offroad_cell = offroad_cell
        # a face takes the colour / level of its FIRST vertex (rendering/cv2.py:44-47, 58)
        if self.faces.shape[0]:
            self.face_category = self.vert_category[self.faces[:, 0]]
        else:
            self.face_category = np.zeros((0,), np.int64)
        self._handles: Dict[int, int] = {}

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def from_npz(cls, path: str, **kw) -> "StaticMap":
        d = np.load(path)
        return cls(d["verts"], d["faces"], [str(c) for c in d["categories"]], d["vert_category"],
                   name=os.path.splitext(os.path.basename(path))[0], left_handed=bool(d["left_handed"]),
                   stoplines=d["stoplines"], stopline_types=[str(s) for s in d["stopline_types"]], **kw)

    @classmethod
    def from_mesh_json(cls, path: str, stoplines_path: Optional[str] = None, left_handed: bool = False, **kw) -> "StaticMap":
        """Reads the reference's `{name}_mesh.json` (BirdviewMesh.serialize, mesh.py:700-719)."""
        with open(path) as f:
            d = json.load(f)
        verts = np.asarray(d["verts"], np.float32)[0][:, :2]
        faces = np.asarray(d["faces"], np.int32)[0]
        vcat = np.asarray(d["vert_category"])[0]
        lines, types = cls._read_stoplines(stoplines_path)
        return cls(verts, faces, d["categories"], vcat, name=os.path.basename(path), left_handed=left_handed,
                   stoplines=lines, stopline_types=types, **kw)

    @classmethod
    def from_birdview_mesh(cls, mesh, batch_index: int = 0, **kw) -> "StaticMap":
        """Accepts a reference `BirdviewMesh` (duck-typed: verts [B,V,2+], faces [B,F,3], categories,
        vert_category [B,V]); takes one batch element."""
        verts = mesh.verts[batch_index][:, :2].detach().cpu().numpy()
        faces = mesh.faces[batch_index].detach().cpu().numpy()
        vcat = mesh.vert_category[batch_index].detach().cpu().numpy()
        return cls(verts, faces, list(mesh.categories), vcat, **kw)

    @classmethod
    def from_lanelet_osm(cls, path: str, origin=(0.0, 0.0), stoplines_path: Optional[str] = None, left_handed: bool = False,
                         **kw) -> "StaticMap":
        """Builds the road + lane-marking mesh of a lanelet2 OSM map (plain or .gz) without the lanelet2 package, as
        `MapConfig.road_mesh` does for the maps that ship without a mesh (map.py:61-74; see osm.py).  `left_handed` is
        the coordinate convention recorded with the map; the markings are derived with left_handed=False, as there."""
        from . import osm
        verts, faces, cats, vcat = osm.birdview_mesh_arrays(osm.LaneletOsm.load(path, origin))
        lines, types = cls._read_stoplines(stoplines_path)
        name = os.path.basename(path).replace(".gz", "").replace(".osm", "")
        return cls(verts, faces, cats, vcat, name=name, left_handed=left_handed, stoplines=lines, stopline_types=types, **kw)

    @staticmethod
    def _read_stoplines(stoplines_path: Optional[str]):
        if stoplines_path is None:
            return None, None
        with open(stoplines_path) as f:
            sl = json.load(f)
        norm = {"traffic-light": "traffic_light", "stop-sign": "stop_sign", "yield-sign": "yield_sign", "yield": "yield_sign"}
        types = [norm.get(s["agent_type"], s["agent_type"]) for s in sl]
        lines = np.array([[s["x"], s["y"], s["length"], s["width"], s["orientation"]] for s in sl], np.float32)
        return lines, types

    @classmethod
    def from_mesh_pickle(cls, path: str, batch_index: int = 0, **kw) -> "StaticMap":
        """Reads a mesh stored by the reference's `BirdviewMesh.pickle` (mesh.py:238-256) WITHOUT the reference
        package: the pickled `torchdrivesim.mesh.*` object is rebuilt as a plain attribute bag, and only tensor /
        array reconstruction helpers are allowed besides it (anything else in the file raises)."""
        import pickle

        class _Bag:
            def __setstate__(self, state):
                self.__dict__.update(state)

        import io

        allowed = {("collections", "OrderedDict"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"),
                   ("torch", "Size"), ("numpy", "ndarray"), ("numpy", "dtype"),
                   ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct")}

        def _storage_from_bytes(b):
            # torch.storage._load_from_bytes is torch.load(..., weights_only=False), i.e. a second, UNRESTRICTED
            # unpickler fed with bytes from the file: never resolve it.  The legacy storage payload is re-read with
            # the weights-only loader, which admits tensors / storages and nothing else.
            return torch.load(io.BytesIO(b), weights_only=True)

        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                if module.startswith("torchdrivesim."):
                    return _Bag
                if (module, name) == ("torch.storage", "_load_from_bytes"):
                    return _storage_from_bytes
                if (module, name) in allowed or (module == "torch" and (name.endswith("Storage") or name in
                                                                         ("float32", "float64", "int64", "int32", "bool", "uint8"))):
                    return super().find_class(module, name)
                raise pickle.UnpicklingError(f"{module}.{name} is not allowed in a mesh pickle")

        with open(path, "rb") as f:
            mesh = _Unpickler(f).load()
        for attr in ("verts", "faces", "categories", "vert_category"):
            if not hasattr(mesh, attr):
                raise _lib.TdsError(f"{path} does not hold a BirdviewMesh (no `{attr}`)")
        return cls.from_birdview_mesh(mesh, batch_index=batch_index, name=os.path.basename(path), **kw)

    def with_extra_meshes(self, meshes, batch_index: int = 0) -> "StaticMap":
        """A new map = this one followed by more static meshes (BirdviewRGBMeshGenerator.add_static_meshes,
        mesh.py:870-883: they are concatenated behind the background mesh).  `meshes`: StaticMap objects or reference
        BirdviewMesh-like objects (verts [B,V,2+], faces [B,F,3], categories, vert_category [B,V])."""
        verts, faces, cats, vcat = [self.verts], [self.faces], list(self.categories), [self.vert_category]
        nv = self.verts.shape[0]
        for m in meshes:
            if not isinstance(m, StaticMap):
                m = StaticMap.from_birdview_mesh(m, batch_index=min(batch_index, m.verts.shape[0] - 1))
            remap = np.zeros(max(len(m.categories), 1), np.int64)
            for i, c in enumerate(m.categories):
                if c not in cats:
                    cats.append(c)
                remap[i] = cats.index(c)
            verts.append(m.verts)
            faces.append(m.faces + nv)
            vcat.append(remap[m.vert_category])
            nv += m.verts.shape[0]
        return StaticMap(np.concatenate(verts), np.concatenate(faces), cats, np.concatenate(vcat), name=self.name,
                         left_handed=self.left_handed, stoplines=self.stoplines, stopline_types=self.stopline_types,
                         raster_cell=self.raster_cell, offroad_cell=self.offroad_cell)

    # ---- queries ------------------------------------------------------------------------------
    @property
    def face_category_names(self) -> List[str]:
        return [self.categories[i] for i in self.face_category]

    def category_verts(self, name: str) -> np.ndarray:
        return self.verts[self.vert_category == self.categories.index(name)]

    @property
    def world_center(self) -> np.ndarray:
        """Centre of the bounding box of the 'road' category (mesh.py:860-868), else of everything."""
        v = self.category_verts("road") if "road" in self.categories else self.verts
        if v.shape[0] == 0:
            return np.zeros(2, np.float32)
        return ((v.min(0) + v.max(0)) / 2).astype(np.float32)

    def traffic_light_poses(self) -> np.ndarray:
        idx = [i for i, t in enumerate(self.stopline_types) if t == "traffic_light"]
        return self.stoplines[idx] if idx else np.zeros((0, 5), np.float32)

    # ---- device ---------------------------------------------------------------------------------
    def handle(self, device) -> int:
        """tds_map_t* for `device` (built on first use)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.TdsError("StaticMap needs a CUDA device: there is no CPU implementation")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._handles:
            lib = _lib.load()
            face_class = np.ascontiguousarray([class_id(self.categories[c]) for c in self.face_category], np.uint8)
            with torch.cuda.device(idx):
                h = lib.tds_map_create(self.verts.ctypes.data_as(ctypes.c_void_p), self.verts.shape[0],
                                       self.faces.ctypes.data_as(ctypes.c_void_p), self.faces.shape[0],
                                       face_class.ctypes.data_as(ctypes.c_void_p), self.raster_cell, self.offroad_cell)
            if not h:
                raise _lib.TdsError(f"tds_map_create failed: {lib.tds_last_error().decode()}")
            self._handles[idx] = h
        return self._handles[idx]

    def info(self, device) -> "_lib.MapInfo":
        out = _lib.MapInfo()
        _lib.check(_lib.load().tds_map_info(self.handle(device), ctypes.byref(out)))
        return out

    def release(self) -> None:
        lib = _lib.load()
        for h in self._handles.values():
            lib.tds_map_destroy(h)
        self._handles.clear()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class MapSet:
    """The maps used by a batch: `maps[env_map[b]]` is the map of environment b."""

    def __init__(self, maps: Sequence[StaticMap], env_map: Optional[torch.Tensor] = None):
        if not 1 <= len(maps) <= _lib.MAX_MAPS:
            raise _lib.TdsError(f"a batch may use 1..{_lib.MAX_MAPS} distinct maps")
        self.maps = list(maps)
        self.env_map = None if env_map is None else env_map.to(torch.int32).contiguous()

    def handles(self, device):
        arr = (ctypes.c_void_p * len(self.maps))(*[m.handle(device) for m in self.maps])
        return arr, len(self.maps)

    def env_map_on(self, device) -> Optional[torch.Tensor]:
        if self.env_map is None:
            return None
        if self.env_map.device != torch.device(device):
            self.env_map = self.env_map.to(device)
        return self.env_map

    def static_categories(self) -> List[str]:
        out: List[str] = []
        for m in self.maps:
            for c in sorted(set(m.face_category_names)):
                if c not in out:
                    out.append(c)
        return out

    def select(self, idx: torch.Tensor) -> "MapSet":
        return MapSet(self.maps, None if self.env_map is None else self.env_map[idx.to(self.env_map.device)])

    def extend(self, n: int) -> "MapSet":
        return MapSet(self.maps, None if self.env_map is None else self.env_map.repeat_interleave(n))
