from dataclasses import MISSING

from ..utils import configclass


@configclass
class PatternBaseCfg:
    func: object = None


@configclass
class GridPatternCfg(PatternBaseCfg):
    resolution: float = MISSING
    size: tuple = MISSING
    direction: tuple = (0.0, 0.0, -1.0)
    ordering: str = "xy"
