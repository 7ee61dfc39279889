"""Colour classes of the birdview: category name -> (class id, draw rank, RGB).

Defaults restate torchdrivesim/rendering/base.py:234-292 (get_default_rendering_levels /
get_default_color_map).  Lower level renders on top, so draw rank = descending level.  Categories of
EQUAL level are drawn in an undefined order by the reference (torch.argsort is unstable,
rendering/cv2.py:47); this build fixes the order with TIE_ORDER (later = on top).
"""
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _lib

Color = Tuple[int, int, int]


def get_default_rendering_levels() -> Dict[str, float]:
    return dict(
        direction=2, ego=3, vehicle=4, bicycle=5, pedestrian=6, map_boundary=7, goal_waypoint=8,
        ground_truth=9, prediction=10, traffic_light=11, traffic_light_green=11, traffic_light_yellow=11,
        traffic_light_red=11, stop_sign=11, yield_sign=11, left_lane=12, joint_lane=13, right_lane=14, road=15,
    )


def get_default_color_map() -> Dict[str, Color]:
    return dict(
        background=(0, 0, 0), road=(155, 155, 155), corridor=(0, 155, 0), ego=(255, 0, 0), vehicle=(32, 74, 135),
        bicycle=(24, 104, 225), pedestrian=(173, 127, 168), ground_truth=(196, 188, 165), prediction=(255, 155, 0),
        left_lane=(80, 127, 86), right_lane=(128, 0, 128), joint_lane=(255, 255, 255), direction=(100, 255, 255),
        rear_lights=(255, 255, 0), map_boundary=(255, 255, 0), traffic_light_green=(81, 179, 100),
        traffic_light_yellow=(240, 189, 39), traffic_light_red=(224, 53, 49), yield_sign=(210, 125, 45),
        stop_sign=(72, 60, 50), goal_waypoint=(139, 64, 0),
    )


TIE_ORDER = ["stop_sign", "yield_sign", "traffic_light", "traffic_light_green", "traffic_light_yellow",
             "traffic_light_red"]

# Fixed class ids for the default categories, so that a map built once can be shared by renderers.
_CLASS_NAMES: List[str] = sorted(get_default_rendering_levels())


def class_id(name: str) -> int:
    """Stable class id of a category name (new names are appended, up to 32 classes)."""
    if name not in _CLASS_NAMES:
        if len(_CLASS_NAMES) >= _lib.MAX_CLASSES:
            raise _lib.TdsError(f"more than {_lib.MAX_CLASSES} rendering categories")
        _CLASS_NAMES.append(name)
    return _CLASS_NAMES.index(name)


def class_names() -> List[str]:
    return list(_CLASS_NAMES)


def quantize_color(rgb: Iterable[float]) -> Tuple[int, int, int]:
    """The colour the cv2 backend actually paints: floor(c/255 * 0.999 * 256) (rendering/cv2.py:50)."""
    c = np.asarray(list(rgb), np.float32) / np.float32(255.0)
    q = np.floor(c * np.float32(1.0 - 1e-3) * np.float32(256)).clip(0, 255).astype(np.uint8)
    return int(q[0]), int(q[1]), int(q[2])


def build_palette(color_map: Dict[str, Color], rendering_levels: Dict[str, float], active: Iterable[str],
                  agent_type_names: Optional[List[str]] = None, direction: bool = True,
                  tl_states: Optional[List[str]] = None,
                  custom: Optional[List[Tuple[str, Color]]] = None) -> "_lib.Palette":
    """Fills the C-ABI palette for the categories in `active`.  `custom` = [(category, (r, g, b))]: extra classes of this
    palette only, drawn at the level of `category` with the given, already quantized, colour (custom agent colours);
    they take the class ids len(class_names()), len(class_names()) + 1, ..."""
    active = list(dict.fromkeys(active))
    for name in active:
        class_id(name)
    pal = _lib.Palette()
    names = class_names()
    pal.n_classes = len(names)
    known = [n for n in names if n in rendering_levels]
    draw = sorted(known, key=lambda k: (-rendering_levels[k], TIE_ORDER.index(k) if k in TIE_ORDER else -1, k))
    rank = {k: i for i, k in enumerate(draw)}
    for i, n in enumerate(names):
        pal.active[i] = 1 if n in active else 0
        pal.rank[i] = rank.get(n, 255)
        if n in color_map:
            r, g, b = quantize_color(color_map[n])
            pal.rgb[i][0], pal.rgb[i][1], pal.rgb[i][2] = r, g, b
        if n in active and (n not in rendering_levels or n not in color_map):
            raise _lib.TdsError(f"category '{n}' needs a colour and a rendering level")
    for t in range(_lib.MAX_AGENT_TYPES):
        pal.agent_type_class[t] = -1
    for t, n in enumerate(agent_type_names or []):
        if t >= _lib.MAX_AGENT_TYPES:
            raise _lib.TdsError(f"at most {_lib.MAX_AGENT_TYPES} agent types are supported")
        pal.agent_type_class[t] = class_id(n)
    for k, (cat, rgb) in enumerate(custom or []):
        i = len(names) + k
        if i >= _lib.MAX_CLASSES:
            raise _lib.TdsError(f"too many distinct custom agent colours: {len(custom)} + {len(names)} categories > {_lib.MAX_CLASSES}")
        if cat not in rank:
            raise _lib.TdsError(f"category '{cat}' needs a rendering level")
        pal.n_classes = i + 1
        pal.active[i] = 1
        pal.rank[i] = rank[cat]
        pal.rgb[i][0], pal.rgb[i][1], pal.rgb[i][2] = int(rgb[0]), int(rgb[1]), int(rgb[2])
    pal.direction_class = class_id("direction") if direction else -1
    for s in range(_lib.MAX_TL_STATES):
        pal.tl_state_class[s] = -1
    for s, n in enumerate(tl_states or []):
        if s >= _lib.MAX_TL_STATES:
            raise _lib.TdsError(f"at most {_lib.MAX_TL_STATES} traffic light states are supported")
        pal.tl_state_class[s] = class_id(f"traffic_light_{n}")
    return pal
