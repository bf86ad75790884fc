from dataclasses import MISSING

from ..utils import configclass


@configclass
class InteractiveSceneCfg:
    num_envs: int = MISSING
    env_spacing: float = MISSING
    lazy_sensor_update: bool = True
    replicate_physics: bool = True
    filter_collisions: bool = True
