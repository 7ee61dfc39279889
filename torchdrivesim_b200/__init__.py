"""torchdrivesim_b200: the per-step hot path of TorchDriveSim (kinematic step, birdview raster, pairwise
collisions, offroad) as hand-written sm_100a CUDA kernels behind the reference's Python interfaces.

The package needs libtds_b200.so (built by `torchdrivesim_b200._build.build_library()` with nvcc) and a
CUDA device; there is no CPU fallback and every op raises if either is missing.
"""
from . import _lib  # noqa: F401
from .kinematic import (BicycleNoReversing, FusedCompoundKinematicModel, KinematicBicycle,  # noqa: F401
                        KinematicModel, KinematicUnicycle, OrientedKinematicModel, SimpleKinematicModel)
from .maps import MapSet, StaticMap  # noqa: F401
from .mesh import B200BirdviewMeshGenerator, BirdviewScene  # noqa: F401
from .rendering import (B200Renderer, B200RendererConfig, BirdviewRenderer, RendererConfig,  # noqa: F401
                        Resolution, renderer_from_config)
from .infractions import (collision_allpairs, collision_detection_with_discs, iou_differentiable,  # noqa: F401
                          offroad_infraction_loss)
from .goals import WaypointGoal  # noqa: F401
from .npc import CompoundNPCController, NPCController, ReplayController, SpawnController  # noqa: F401
from .observation_noise import (ObservationNoise, ObservationNoiseConfig, StandardSensingObservationNoise,  # noqa: F401
                                StandardSensingObservationNoiseConfig)
from .simulator import CollisionMetric, Simulator, TorchDriveConfig  # noqa: F401
from .graph import GraphedHotPath  # noqa: F401
from .rollout import FusedRollout  # noqa: F401
from . import distributed, ops, torch_ops  # noqa: F401
from .traffic_lights import TrafficLightController, unroll_controller  # noqa: F401
from .traffic_controls import BaseTrafficControl, StopSignControl, TrafficLightControl, YieldControl  # noqa: F401

__version__ = "0.1.0"
