from dataclasses import MISSING

from ..utils import configclass


class AssetBase:
    pass


class RigidObject(AssetBase):
    pass


class Articulation(AssetBase):
    pass


@configclass
class AssetBaseCfg:
    @configclass
    class InitialStateCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)

    class_type: type = None
    prim_path: str = MISSING
    spawn: object = None
    init_state: InitialStateCfg = InitialStateCfg()
    collision_group: int = 0
    debug_vis: bool = False


@configclass
class RigidObjectCfg(AssetBaseCfg):
    pass


@configclass
class ArticulationCfg(AssetBaseCfg):
    @configclass
    class InitialStateCfg:
        pos: tuple = (0.0, 0.0, 0.0)
        rot: tuple = (1.0, 0.0, 0.0, 0.0)
        lin_vel: tuple = (0.0, 0.0, 0.0)
        ang_vel: tuple = (0.0, 0.0, 0.0)
        joint_pos: dict = {".*": 0.0}
        joint_vel: dict = {".*": 0.0}

    init_state: InitialStateCfg = InitialStateCfg()
    soft_joint_pos_limit_factor: float = 1.0
    actuators: dict = MISSING
