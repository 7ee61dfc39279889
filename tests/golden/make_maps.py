"""Generates tests/golden/maps/*.npz from the reference's bundled map assets.

Run in the build container only (needs /root/reference).  The npz files are the benchmark /
parity INPUTS (only carla_Town01 and carla_Town02 ship with a mesh); they are
data fixtures, not reference source code.

Schema: verts [V,2] f32, faces [F,3] i32, vert_category [V] u8, categories [C] str,
stoplines [L,5] f32 (x, y, length, width, orientation), stopline_types [L] str,
left_handed bool.
"""
import json
import os
import sys

import numpy as np

REF_MAPS = "/root/reference/torchdrivesim/resources/maps"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")


def convert(name: str) -> None:
    d = json.load(open(os.path.join(REF_MAPS, name, f"{name}_mesh.json")))
    meta = json.load(open(os.path.join(REF_MAPS, name, "metadata.json")))
    stop = json.load(open(os.path.join(REF_MAPS, name, f"{name}_stoplines.json")))
    verts = np.asarray(d["verts"], np.float32)[0]
    faces = np.asarray(d["faces"], np.int32)[0]
    vcat = np.asarray(d["vert_category"], np.uint8)[0]
    # Stopline.__post_init__ (map.py:27-35) normalises the type names
    norm = {"traffic-light": "traffic_light", "stop-sign": "stop_sign", "yield-sign": "yield_sign", "yield": "yield_sign"}
    types = np.array([norm.get(s["agent_type"], s["agent_type"]) for s in stop])
    # traffic_controls_from_map_config (map.py:203-229) builds fp32 tensors from these python floats
    lines = np.array([[s["x"], s["y"], s["length"], s["width"], s["orientation"]] for s in stop], np.float32)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), verts=verts, faces=faces, vert_category=vcat,
                        categories=np.array(d["categories"]), stoplines=lines, stopline_types=types,
                        left_handed=np.array(bool(meta["left_handed_coordinates"])))
    print(name, verts.shape, faces.shape, d["categories"], lines.shape, sorted(set(types)))


def convert_osm(name: str, with_npz: bool) -> None:
    """Maps that the reference derives from their lanelet2 OSM file (map.py:61-74): the OSM (gzipped) and stop lines are
    copied as data fixtures; with_npz also stores the mesh that torchdrivesim_b200.osm builds from them in the schema
    above, as the INPUT of the oracle-vs-GPU parity tests on that map (the reference needs lanelet2 to load it)."""
    import gzip
    import shutil
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import torchdrivesim_b200 as tds
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(REF_MAPS, name, f"{name}.osm"), "rb") as f, gzip.GzipFile(os.path.join(OUT, f"{name}.osm.gz"), "wb", 9, mtime=0) as g:
        shutil.copyfileobj(f, g)
    shutil.copy(os.path.join(REF_MAPS, name, f"{name}_stoplines.json"), os.path.join(OUT, f"{name}_stoplines.json"))
    os.chmod(os.path.join(OUT, f"{name}_stoplines.json"), 0o644)
    if with_npz:
        meta = json.load(open(os.path.join(REF_MAPS, name, "metadata.json")))
        m = tds.StaticMap.from_lanelet_osm(os.path.join(OUT, f"{name}.osm.gz"), origin=tuple(meta["lanelet_map_origin"]),
                                           stoplines_path=os.path.join(OUT, f"{name}_stoplines.json"),
                                           left_handed=bool(meta["left_handed_coordinates"]))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), verts=m.verts, faces=m.faces, vert_category=m.vert_category.astype(np.uint8),
                            categories=np.array(m.categories), stoplines=m.stoplines, stopline_types=np.array(m.stopline_types),
                            left_handed=np.array(m.left_handed))
        print(name, "from osm:", m.verts.shape, m.faces.shape, m.categories, m.stoplines.shape)


if __name__ == "__main__":
    for n in sys.argv[1:] or ["carla_Town01", "carla_Town02"]:
        convert(n)
    if not sys.argv[1:]:
        convert_osm("carla_Town01", with_npz=False)       # pin the OSM path against the shipped meshes
        convert_osm("carla_Town02", with_npz=False)
        convert_osm("carla_Town10HD", with_npz=True)      # a map that ships without a mesh
