"""Scene descriptor + generator: what `BirdviewRGBMeshGenerator.generate` (torchdrivesim/mesh.py:761-1157)
produces in the reference, without materialising a mesh per camera.

The reference concatenates background + actors + controls into an RGBMesh of ~1.86 MB PER CAMERA.  Here
`generate` returns a `BirdviewScene`: references to the shared static map(s) and the per-environment
tensors (agent state / size / type / presence, traffic-light corners and state).  The raster kernel
assembles agent rectangles, direction triangles and control rectangles itself, with the reference's
vertex arithmetic (mesh.py:911-1004, utils.py:82-96).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from .maps import MapSet, StaticMap
from .palette import build_palette, class_id


@dataclass
class BirdviewScene:
    mapset: MapSet
    agent_state: Optional[Tensor]      # [B,N,4]
    agent_size: Optional[Tensor]       # [B,N,2]
    agent_type: Optional[Tensor]       # [B,N] int
    present: Optional[Tensor]          # [B,N] or [B,Nc,N] bool
    agent_type_names: List[str]
    render_agent_direction: bool
    tl_corners: Optional[Tensor]       # [B,L,4,2]
    tl_state: Optional[Tensor]         # [B,L] int
    tl_allowed_states: List[str]
    rect_corners: Optional[Tensor]     # [B,R,4,2]
    rect_class: Optional[Tensor]       # [B,R] int32
    rect_categories: List[str]
    batch_size: int
    workspace: Optional[Tensor] = None
    cam_tris: Optional[Tensor] = None          # [B,Nc,Tc,3,2] triangles of one camera each (waypoint discs)
    cam_tri_class: Optional[Tensor] = None     # [B,Nc,Tc] int32 class ids (< 0: skipped)
    agent_class: Optional[Tensor] = None       # [B,Nc,N] uint8: class of each agent's rectangle per camera (custom colours)
    custom_classes: Optional[list] = None      # [(agent type name, (r, g, b))] behind the ids of agent_class

    def slice(self, b0: int, b1: int) -> "BirdviewScene":
        """Environments [b0, b1) of the scene (views, no copies) - used to render in chunks."""
        cut = lambda t: None if t is None else t[b0:b1]
        ms = MapSet(self.mapset.maps, None if self.mapset.env_map is None else self.mapset.env_map[b0:b1])
        return BirdviewScene(ms, cut(self.agent_state), cut(self.agent_size), cut(self.agent_type), cut(self.present),
                             self.agent_type_names, self.render_agent_direction, cut(self.tl_corners), cut(self.tl_state),
                             self.tl_allowed_states, cut(self.rect_corners), cut(self.rect_class), self.rect_categories,
                             b1 - b0, self.workspace, cut(self.cam_tris), cut(self.cam_tri_class), cut(self.agent_class),
                             self.custom_classes)

    def palette(self, color_map, rendering_levels) -> "_lib.Palette":
        active = list(self.mapset.static_categories())
        if self.agent_state is not None:
            active += list(self.agent_type_names)
            if self.render_agent_direction:
                active.append("direction")
        if self.tl_corners is not None:
            active += [f"traffic_light_{s}" for s in self.tl_allowed_states]
        active += self.rect_categories
        if self.cam_tris is not None:
            active.append("goal_waypoint")
        return build_palette(color_map, rendering_levels, active, self.agent_type_names,
                             self.render_agent_direction, self.tl_allowed_states, custom=self.custom_classes)


def _as_mapset(mesh) -> MapSet:
    if isinstance(mesh, MapSet):
        return mesh
    if isinstance(mesh, StaticMap):
        return MapSet([mesh])
    if hasattr(mesh, "verts") and hasattr(mesh, "vert_category"):    # a reference BirdviewMesh
        return MapSet([StaticMap.from_birdview_mesh(mesh)])
    raise _lib.TdsError(f"cannot use {type(mesh)} as a background mesh")


class B200BirdviewMeshGenerator:
    """Drop-in for `birdview_mesh_generator=` of the Simulator (seam described in SURVEY.md §8b)."""

    def __init__(self, background_mesh: Union[StaticMap, MapSet], color_map: Dict[str, Tuple[int, int, int]],
                 rendering_levels: Dict[str, float], world_center: Optional[Tensor] = None,
                 agent_attributes: Optional[Tensor] = None, agent_types: Optional[Tensor] = None,
                 agent_type_names: Optional[List[str]] = None, render_agent_direction: bool = True,
                 traffic_controls: Optional[Dict[str, object]] = None, batch_size: Optional[int] = None):
        self.color_map = color_map
        self.rendering_levels = rendering_levels
        self.mapset = _as_mapset(background_mesh)
        self.background_mesh = background_mesh
        self._batch_size = batch_size
        self._world_center = world_center
        self.agent_size = None
        self.agent_type = None
        self.agent_type_names: List[str] = ["vehicle"]
        self.render_agent_direction = render_agent_direction
        self.tl_corners = None
        self.tl_allowed_states: List[str] = []
        self.rect_corners = None
        self.rect_class = None
        self.rect_categories: List[str] = []
        self._workspace = None
        if agent_attributes is not None:
            self.initialize_actors_mesh(agent_attributes, agent_types, agent_type_names, render_agent_direction)
        if traffic_controls is not None:
            self.initialize_traffic_controls_mesh(traffic_controls)

    # ---- reference-compatible surface ---------------------------------------------------------
    @property
    def world_center(self) -> Tensor:
        if self._world_center is None:
            centers = torch.stack([torch.from_numpy(m.world_center) for m in self.mapset.maps])
            if self.mapset.env_map is not None:
                self._world_center = centers.to(self.mapset.env_map.device)[self.mapset.env_map.long()]
            else:
                self._world_center = centers[:1].expand(self._batch_size or 1, 2)
        return self._world_center

    def initialize_actors_mesh(self, agent_attributes: Tensor, agent_types: Optional[Tensor],
                               agent_type_names: Optional[List[str]], render_agent_direction: bool = True):
        self.agent_size = agent_attributes[..., :2]
        # int32 once, here: the kernel's index type (a conversion per render call would be a library kernel on the hot path)
        self.agent_type = None if agent_types is None else agent_types.to(torch.int32)
        self.agent_type_names = list(agent_type_names) if agent_type_names else ["vehicle"]
        self.render_agent_direction = render_agent_direction
        self._batch_size = agent_attributes.shape[0]

    def initialize_traffic_controls_mesh(self, traffic_controls: Dict[str, object]):
        rects, classes, cats = [], [], []
        for name in ("stop_sign", "yield_sign"):
            el = traffic_controls.get(name)
            if el is not None and el.corners.shape[-3] > 0:
                rects.append(el.corners)
                classes.append(torch.full(el.corners.shape[:2], class_id(name), dtype=torch.int32,
                                          device=el.corners.device))
                cats.append(name)
        self.rect_corners = torch.cat(rects, dim=1) if rects else None
        self.rect_class = torch.cat(classes, dim=1) if rects else None
        self.rect_categories = cats
        tl = traffic_controls.get("traffic_light")
        if tl is not None and tl.corners.shape[-3] > 0:
            self.tl_corners = tl.corners
            self.tl_allowed_states = list(tl.allowed_states)
        else:
            self.tl_corners, self.tl_allowed_states = None, []

    def add_static_meshes(self, meshes) -> None:
        """Includes additional static elements in the background (mesh.py:870-883).  `meshes`: StaticMap objects or
        reference BirdviewMesh objects whose categories have a colour and a rendering level; batch element b of a
        mesh goes to map b of the map set (a mesh of batch size 1 goes to all of them).  The grids of the extended
        maps are rebuilt on first use."""
        meshes = list(meshes)
        if not meshes:
            return
        self.mapset = MapSet([m.with_extra_meshes(meshes, batch_index=i) for i, m in enumerate(self.mapset.maps)],
                             self.mapset.env_map)
        self.background_mesh = self.mapset

    def generate(self, num_cameras: int, agent_state: Optional[Tensor] = None, present_mask: Optional[Tensor] = None,
                 traffic_lights=None, waypoints: Optional[Tensor] = None,
                 waypoints_rendering_mask: Optional[Tensor] = None,
                 custom_agent_colors: Optional[Tensor] = None) -> BirdviewScene:
        """Same arguments as the reference's generate (mesh.py:1053-1075): agent_state [B,Nc,N,4] (one
        copy per camera; must be the broadcast of a [B,N,4] tensor), present_mask [B,Nc,N]."""
        state = size = types = present = None
        if agent_state is not None and self.agent_size is not None:
            if agent_state.dim() == 4:
                if agent_state.shape[1] != num_cameras:
                    raise _lib.TdsError("agent_state must be [B,Nc,N,4]")
                if num_cameras > 1 and agent_state.stride(1) != 0:
                    raise _lib.TdsError("per-camera agent states are not supported: pass state[:, None].expand(...)")
                agent_state = agent_state[:, 0]
            state, size, types = agent_state, self.agent_size, self.agent_type
            if present_mask is not None:
                present = present_mask
                if present.dim() == 3 and (present.shape[1] == 1 or present.stride(1) == 0):
                    present = present[:, 0]
        B = state.shape[0] if state is not None else (self._batch_size or 1)
        tl_corners = tl_state = None
        if traffic_lights is not None and self.tl_corners is not None:
            tl_state = traffic_lights.state
            if tl_state.shape[0] == B * num_cameras and num_cameras > 1:     # `.extend(Nc)`-ed copy
                tl_state = tl_state.reshape(B, num_cameras, -1)[:, 0]
            tl_corners = self.tl_corners
        scene = BirdviewScene(mapset=self.mapset, agent_state=state, agent_size=size, agent_type=types, present=present,
                              agent_type_names=self.agent_type_names, render_agent_direction=self.render_agent_direction,
                              tl_corners=tl_corners, tl_state=tl_state, tl_allowed_states=self.tl_allowed_states,
                              rect_corners=self.rect_corners, rect_class=self.rect_class,
                              rect_categories=self.rect_categories, batch_size=B, workspace=self._workspace)
        if custom_agent_colors is not None and state is not None:
            scene.agent_class, scene.custom_classes = self._custom_agent_classes(num_cameras, custom_agent_colors, types, state)
        if waypoints is not None and waypoints.shape[-2] > 0:
            scene.cam_tris, scene.cam_tri_class = self._waypoint_triangles(num_cameras, waypoints, waypoints_rendering_mask)
        if state is not None:
            need = _lib.load().tds_raster_workspace_bytes(B, state.shape[1], 0 if tl_corners is None else tl_corners.shape[1],
                                                          0 if self.rect_corners is None else self.rect_corners.shape[1])
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != state.device:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=state.device)
            scene.workspace = self._workspace
        return scene

    # ---- custom agent colours (mesh.py:1092-1099, rendering/cv2.py:50) ----------------------------------
    def _custom_agent_classes(self, num_cameras: int, colors: Tensor, types: Optional[Tensor], state: Tensor):
        """colors [B,Nc,N,3] in [0,1] -> (uint8 class per camera and agent, [(type name, rgb)] of those classes).
        The reference paints the rectangle of agent n in camera c with floor(color * (1 - 1e-3) * 256) at the level of
        the agent's type; every distinct (type, colour) pair becomes a class of this scene's palette.  Finding the
        distinct pairs synchronises with the host (torch.unique) - the price of arbitrary colours in a class-based
        raster; where rectangles of EQUAL level overlap, the reference's order is undefined (unstable argsort)."""
        from .palette import class_names
        B, N = state.shape[0], state.shape[1]
        if tuple(colors.shape) != (B, num_cameras, N, 3):
            raise _lib.TdsError("custom_agent_colors must be [B,Nc,N,3]")
        q = (colors.detach().to(torch.float32) * (1.0 - 1e-3) * 256).floor().clamp(0, 255).to(torch.int64)
        ty = torch.zeros(B, N, dtype=torch.int64, device=q.device) if types is None else types.to(q.device).to(torch.int64)
        key = (ty[:, None, :] << 24) | (q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2]
        uniq, inv = torch.unique(key, return_inverse=True)
        base = len(class_names())
        if base + uniq.numel() > _lib.MAX_CLASSES:
            raise _lib.TdsError(f"{uniq.numel()} distinct custom agent colours: at most {_lib.MAX_CLASSES - base} fit the palette")
        custom = []
        for k in uniq.tolist():
            t = k >> 24
            name = self.agent_type_names[t] if t < len(self.agent_type_names) else self.agent_type_names[0]
            custom.append((name, ((k >> 16) & 255, (k >> 8) & 255, k & 255)))
        return (base + inv).to(torch.uint8).contiguous(), custom

    # ---- goal waypoints (mesh.py:885-909, 1120-1145, 1243-1271) -------------------------------------------
    waypoint_radius = 2.0
    waypoint_num_triangles = 10

    def _disc_template(self, device) -> Tuple[Tensor, Tensor]:
        """Centre + rim vertices of the waypoint disc and its faces (0, k, k+1), ..., (0, n, 1).  The rim is made by
        REPEATED rotation of (radius, 0) with torch.matmul on the CPU, like generate_disc_mesh: the accumulated fp32
        rounding of those ATen ops is part of the reference's result."""
        key = (self.waypoint_radius, self.waypoint_num_triangles)
        if getattr(self, "_disc_cache", None) is None or self._disc_cache[0] != key:
            n = self.waypoint_num_triangles
            step = torch.deg2rad(torch.tensor([[360 / n]], dtype=torch.float32))
            c, s = torch.cos(step), torch.sin(step)
            rot = torch.stack([torch.cat([c, -s], -1), torch.cat([s, c], -1)], -2)
            verts = [torch.zeros(1, 2), torch.tensor([[self.waypoint_radius, 0.0]], dtype=torch.float32)]
            for _ in range(n - 1):
                verts.append(torch.matmul(rot, verts[-1].unsqueeze(-1)).squeeze(-1))
            faces = torch.tensor([[0, k, k + 1] for k in range(1, n)] + [[0, n, 1]], dtype=torch.long)
            self._disc_cache = (key, torch.cat(verts, 0), faces, {})
        # one copy per device, made on first use (a host -> device copy cannot be part of a CUDA graph capture)
        on_dev = self._disc_cache[3]
        if str(device) not in on_dev:
            on_dev[str(device)] = (self._disc_cache[1].to(device), self._disc_cache[2].to(device))
        return on_dev[str(device)]

    def _waypoint_triangles(self, num_cameras: int, waypoints: Tensor, mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
        """waypoints [B,Nc,M,2] (+ mask [B,Nc,M]) -> world-space triangles [B,Nc,M*n,3,2] and their class ids.  A disc
        is the template translated to its waypoint (the pose has psi = 0: the rotation is the identity); the faces
        of a masked waypoint collapse onto vertex 0 of the camera's waypoint mesh, the centre of its FIRST waypoint."""
        if waypoints.dim() != 4 or waypoints.shape[1] != num_cameras or waypoints.shape[-1] != 2:
            raise _lib.TdsError("waypoints must be [B,Nc,M,2]")
        wp = waypoints.detach().to(torch.float32)
        B, Nc, M = wp.shape[:3]
        dv, df = self._disc_template(wp.device)
        verts = dv[None, None, None] + wp[..., None, :]                       # [B,Nc,M,V,2]
        tris = verts[:, :, :, df]                                               # [B,Nc,M,n,3,2]
        if mask is not None:
            if tuple(mask.shape) != (B, Nc, M):
                raise _lib.TdsError("waypoints_rendering_mask must be [B,Nc,M]")
            first = verts[:, :, 0, 0]                                           # [B,Nc,2]
            tris = torch.where(mask.to(torch.bool)[..., None, None, None], tris, first[:, :, None, None, None, :])
        n = df.shape[0]
        cls = torch.full((B, Nc, M * n), class_id("goal_waypoint"), dtype=torch.int32, device=wp.device)
        return tris.reshape(B, Nc, M * n, 3, 2).contiguous(), cls

    # ---- batch plumbing -----------------------------------------------------------------------
    def _tensors(self):
        return ("agent_size", "agent_type", "tl_corners", "rect_corners", "rect_class", "_world_center")

    def _clone(self):
        other = self.__class__(self.mapset, self.color_map.copy(), self.rendering_levels.copy(),
                               batch_size=self._batch_size)
        for k in ("agent_type_names", "render_agent_direction", "tl_allowed_states", "rect_categories") + self._tensors():
            setattr(other, k, getattr(self, k))
        other.background_mesh = self.background_mesh
        return other

    def to(self, device):
        for k in self._tensors():
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device))
        if self.mapset.env_map is not None:
            self.mapset.env_map = self.mapset.env_map.to(device)
        self._workspace = None
        return self

    def copy(self):
        return self._clone()

    def expand(self, n: int):
        other = self._clone()
        grow = lambda x: x.unsqueeze(1).expand((x.shape[0], n) + x.shape[1:]).reshape((n * x.shape[0],) + x.shape[1:])
        for k in self._tensors():
            v = getattr(self, k)
            if v is not None:
                setattr(other, k, grow(v))
        other.mapset = self.mapset.extend(n)
        if self._batch_size is not None:
            other._batch_size = self._batch_size * n
        return other

    def select_batch_elements(self, idx: Tensor):
        other = self._clone()
        for k in self._tensors():
            v = getattr(self, k)
            if v is not None:
                setattr(other, k, v[idx])
        other.mapset = self.mapset.select(idx)
        other._batch_size = int(idx.numel())
        return other
