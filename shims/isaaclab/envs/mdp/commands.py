from dataclasses import MISSING

from ...managers import CommandTermCfg
from ...utils import configclass


@configclass
class UniformPose2dCommandCfg(CommandTermCfg):
    asset_name: str = MISSING
    simple_heading: bool = MISSING

    @configclass
    class Ranges:
        pos_x: tuple = MISSING
        pos_y: tuple = MISSING
        heading: tuple = MISSING

    ranges: Ranges = MISSING
    goal_pose_visualizer_cfg: object = None
