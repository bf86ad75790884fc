from dataclasses import MISSING

from ..utils import configclass


class TerrainImporter:
    def __init__(self, cfg):
        self.cfg = cfg


@configclass
class TerrainImporterCfg:
    class_type: type = TerrainImporter
    collision_group: int = -1
    prim_path: str = MISSING
    num_envs: int = 1
    terrain_type: str = "generator"
    terrain_generator: object = None
    usd_path: object = None
    env_spacing: object = None
    visual_material: object = None
    physics_material: object = None
    max_init_terrain_level: object = None
    debug_vis: bool = False
