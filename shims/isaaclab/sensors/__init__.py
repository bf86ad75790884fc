from dataclasses import MISSING

from ..utils import configclass
from . import patterns  # noqa: F401


class SensorBase:
    pass


class Camera(SensorBase):
    pass


class TiledCamera(Camera):
    pass


class RayCaster(SensorBase):
    pass


@configclass
class SensorBaseCfg:
    class_type: type = None
    prim_path: str = MISSING
    update_period: float = 0.0
    history_length: int = 0
    debug_vis: bool = False


@configclass
class RayCasterCfg(SensorBaseCfg):
    @configclass
    class OffsetCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)

    mesh_prim_paths: list = MISSING
    offset: OffsetCfg = OffsetCfg()
    attach_yaw_only: bool = MISSING
    pattern_cfg: object = MISSING
    max_distance: float = 1e6
    drift_range: tuple = (0.0, 0.0)


@configclass
class CameraCfg(SensorBaseCfg):
    @configclass
    class OffsetCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)
        convention: str = "ros"

    offset: OffsetCfg = OffsetCfg()
    spawn: object = None
    depth_clipping_behavior: str = "none"
    data_types: list = ["rgb"]
    width: int = MISSING
    height: int = MISSING
    semantic_filter: object = "*:*"
    colorize_semantic_segmentation: bool = True
    colorize_instance_id_segmentation: bool = True
    colorize_instance_segmentation: bool = True


@configclass
class TiledCameraCfg(CameraCfg):
    pass
