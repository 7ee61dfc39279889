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


if __name__ == "__main__":
    for n in sys.argv[1:] or ["carla_Town01", "carla_Town02"]:
        convert(n)
