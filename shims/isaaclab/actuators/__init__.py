from dataclasses import MISSING

from ..utils import configclass


@configclass
class ActuatorBaseCfg:
    class_type: type = None
    joint_names_expr: list = MISSING
    effort_limit: object = None
    velocity_limit: object = None
    effort_limit_sim: object = None
    velocity_limit_sim: object = None
    stiffness: object = MISSING
    damping: object = MISSING
    armature: object = None
    friction: object = None


@configclass
class ImplicitActuatorCfg(ActuatorBaseCfg):
    pass


@configclass
class IdealPDActuatorCfg(ActuatorBaseCfg):
    pass


@configclass
class DCMotorCfg(IdealPDActuatorCfg):
    saturation_effort: float = MISSING
