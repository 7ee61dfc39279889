"""Renderer backend producing birdview images with the sm_100a raster kernel.

Mirror of the reference's renderer seam (torchdrivesim/rendering/base.py:23-220,
rendering/__init__.py:18-50): a `BirdviewRenderer` with `render_frame(scene, camera_xy, camera_sc, res,
fov) -> [(B*Nc),3,H,W] float32 in [0,255]`, `color_map`, `rendering_levels`, `res`, `scale`, `copy()`,
`get_color()`.  The oracle for the pixels is the reference's cv2 backend (rendering/cv2.py:27-70).
"""
import collections
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops
from .palette import get_default_color_map, get_default_rendering_levels

Resolution = collections.namedtuple('Resolution', ['width', 'height'])


@dataclass
class RendererConfig:
    """Same fields as the reference's RendererConfig (rendering/base.py:23-34)."""
    backend: str = 'default'
    render_agent_direction: bool = True
    left_handed_coordinates: bool = False
    highlight_ego_vehicle: bool = False
    shift_mesh_by_camera_before_rendering: bool = True
    device: Optional[str] = None


@dataclass
class B200RendererConfig(RendererConfig):
    backend: str = 'b200'


class BirdviewRenderer:
    def __init__(self, cfg: RendererConfig, color_map: Optional[Dict[str, Tuple[int, int, int]]] = None,
                 rendering_levels: Optional[Dict[str, float]] = None, res: Resolution = Resolution(64, 64),
                 fov: float = 35):
        self.cfg = cfg
        self.res = res
        self.scale = 2.0 / fov
        self.color_map = get_default_color_map() if color_map is None else color_map
        self.rendering_levels = get_default_rendering_levels() if rendering_levels is None else rendering_levels

    def copy(self):
        other = self.__class__(cfg=self.cfg, color_map=self.color_map.copy(),
                               rendering_levels=self.rendering_levels.copy(), res=self.res)
        other.scale = self.scale
        return other

    def get_color(self, element_type: str) -> Tuple[int, int, int]:
        return self.color_map[element_type]

    def render_frame(self, scene, camera_xy: Tensor, camera_sc: Tensor, res: Optional[Resolution] = None,
                     fov: Optional[float] = None) -> Tensor:
        raise NotImplementedError


class B200Renderer(BirdviewRenderer):
    """Pixel-identical replacement of CV2Renderer running on the GPU.  As in the cv2 backend, the
    left-handed flip is a no-op (it is applied twice there, rendering/cv2.py:63-64 and 68-69) and the
    image is not differentiable."""

    def render_frame(self, scene, camera_xy: Tensor, camera_sc: Tensor, res: Optional[Resolution] = None,
                     fov: Optional[float] = None, out: Optional[Tensor] = None, dtype=None) -> Tensor:
        """dtype: None / torch.float32 = the reference's image ([.,3,H,W] float32 in [0,255]); torch.uint8 = the same
        values as bytes; 'rank' = uint8 [.,H,W] draw ranks (index into `rank_table(scene)`), the most compact form."""
        from .mesh import BirdviewScene
        if not isinstance(scene, BirdviewScene):
            raise _lib.TdsError("B200Renderer renders BirdviewScene objects made by B200BirdviewMeshGenerator.generate; "
                                "pass birdview_mesh_generator=B200BirdviewMeshGenerator(...) to the Simulator")
        res = self.res if res is None else res
        if res.width != res.height:
            raise _lib.TdsError("only square resolutions are supported (as in the reference)")
        fov_m = (2.0 / self.scale) if fov is None else fov
        if camera_xy.dim() == 2:
            camera_xy, camera_sc = camera_xy.unsqueeze(1), camera_sc.unsqueeze(1)
        B, Nc = camera_xy.shape[0], camera_xy.shape[1]
        if B != scene.batch_size:
            raise _lib.TdsError(f"camera batch {B} does not match the scene batch {scene.batch_size}")
        palette = scene.palette(self.color_map, self.rendering_levels)
        fmt = image_format_of(dtype)
        tail = (res.height, res.width) if fmt == _lib.IMAGE_RANK else (3, res.height, res.width)
        if out is not None:
            out = out.view((B, Nc) + tail)
        img = ops.raster_birdview(scene.mapset, palette, camera_xy, camera_sc, scene.agent_state, scene.agent_size,
                                  scene.agent_type, scene.present, scene.tl_corners, scene.tl_state,
                                  scene.rect_corners, scene.rect_class, res.height, fov_m, out=out,
                                  workspace=scene.workspace, cam_tris=scene.cam_tris, cam_tri_class=scene.cam_tri_class,
                                  image_format=fmt, agent_class=scene.agent_class)
        return img.reshape((B * Nc,) + tail)

    def rank_table(self, scene):
        """(rgb uint8 [K+1,3], class ids [K+1]) for images rendered with dtype='rank' from this scene."""
        return ops.raster_rank_table(scene.palette(self.color_map, self.rendering_levels))


def image_format_of(dtype) -> int:
    import torch
    if dtype is None or dtype == torch.float32:
        return _lib.IMAGE_F32
    if dtype == torch.uint8:
        return _lib.IMAGE_U8
    if dtype == 'rank':
        return _lib.IMAGE_RANK
    raise _lib.TdsError(f"unsupported image dtype {dtype}: use torch.float32, torch.uint8 or 'rank'")


def renderer_from_config(cfg: RendererConfig, *args, **kwargs) -> BirdviewRenderer:
    """Only the B200 backend exists here ('default' resolves to it); the reference's cv2 / pytorch3d /
    nvdiffrast backends are what this renderer replaces."""
    if cfg.backend in ('default', 'b200'):
        return B200Renderer(cfg, *args, **kwargs)
    raise ValueError(f"Unrecognized renderer backend for torchdrivesim_b200: {cfg.backend}")
