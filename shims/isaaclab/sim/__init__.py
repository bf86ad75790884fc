"""isaaclab.sim cfg stand-ins (data only)."""
from dataclasses import MISSING

from ..utils import configclass


@configclass
class RigidBodyPropertiesCfg:
    rigid_body_enabled: object = None
    kinematic_enabled: object = None
    disable_gravity: object = None
    linear_damping: object = None
    angular_damping: object = None
    max_linear_velocity: object = None
    max_angular_velocity: object = None
    max_depenetration_velocity: object = None
    max_contact_impulse: object = None
    enable_gyroscopic_forces: object = None
    retain_accelerations: object = None
    solver_position_iteration_count: object = None
    solver_velocity_iteration_count: object = None
    sleep_threshold: object = None
    stabilization_threshold: object = None


@configclass
class ArticulationRootPropertiesCfg:
    articulation_enabled: object = None
    enabled_self_collisions: object = None
    solver_position_iteration_count: object = None
    solver_velocity_iteration_count: object = None
    sleep_threshold: object = None
    stabilization_threshold: object = None
    fix_root_link: object = None


@configclass
class RigidBodyMaterialCfg:
    func: object = None
    static_friction: float = 0.5
    dynamic_friction: float = 0.5
    restitution: float = 0.0
    improve_patch_friction: bool = True
    friction_combine_mode: str = "average"
    restitution_combine_mode: str = "average"
    compliant_contact_stiffness: float = 0.0
    compliant_contact_damping: float = 0.0


@configclass
class SpawnerCfg:
    func: object = None
    visible: bool = True
    semantic_tags: object = None
    copy_from_source: bool = True


@configclass
class UsdFileCfg(SpawnerCfg):
    usd_path: str = MISSING
    variants: object = None
    scale: object = None
    rigid_props: object = None
    collision_props: object = None
    activate_contact_sensors: bool = False
    mass_props: object = None
    articulation_props: object = None
    fixed_tendons_props: object = None
    joint_drive_props: object = None
    deformable_props: object = None
    visual_material_path: str = "material"
    visual_material: object = None


@configclass
class GroundPlaneCfg(SpawnerCfg):
    usd_path: str = "default_environment.usd"
    color: object = (0.0, 0.0, 0.0)
    size: tuple = (100.0, 100.0)
    physics_material: RigidBodyMaterialCfg = RigidBodyMaterialCfg()


@configclass
class DistantLightCfg(SpawnerCfg):
    color: tuple = (1.0, 1.0, 1.0)
    intensity: float = 1.0
    angle: float = 0.53
    exposure: float = 0.0


@configclass
class DomeLightCfg(DistantLightCfg):
    texture_file: object = None


@configclass
class PinholeCameraCfg(SpawnerCfg):
    projection_type: str = "pinhole"
    clipping_range: tuple = (0.01, 1e6)
    focal_length: float = 24.0
    focus_distance: float = 400.0
    f_stop: float = 0.0
    horizontal_aperture: float = 20.955
    vertical_aperture: object = None
    horizontal_aperture_offset: float = 0.0
    vertical_aperture_offset: float = 0.0
    lock_camera: bool = True


@configclass
class PhysxCfg:
    solver_type: int = 1
    min_position_iteration_count: int = 1
    max_position_iteration_count: int = 255
    min_velocity_iteration_count: int = 0
    max_velocity_iteration_count: int = 255
    enable_ccd: bool = False
    enable_stabilization: bool = True
    bounce_threshold_velocity: float = 0.5
    friction_offset_threshold: float = 0.04
    friction_correlation_distance: float = 0.025
    gpu_max_rigid_contact_count: int = 2**23
    gpu_max_rigid_patch_count: int = 5 * 2**15


@configclass
class RenderCfg:
    enable_translucency: bool = False
    antialiasing_mode: str = "DLSS"


@configclass
class SimulationCfg:
    physics_prim_path: str = "/physicsScene"
    device: str = "cuda:0"
    dt: float = 1.0 / 60.0
    render_interval: int = 1
    gravity: tuple = (0.0, 0.0, -9.81)
    enable_scene_query_support: bool = False
    use_fabric: bool = True
    disable_contact_processing: bool = False
    physx: PhysxCfg = PhysxCfg()
    physics_material: RigidBodyMaterialCfg = RigidBodyMaterialCfg()
    render: RenderCfg = RenderCfg()


class SimulationContext:
    _instance = None

    @classmethod
    def instance(cls):
        return cls._instance
