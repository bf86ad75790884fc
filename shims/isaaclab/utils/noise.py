"""Noise cfgs (data only).  AdditiveGaussian: x + mean + std*randn_like(x) [UPSTREAM-RECALL]."""
from .configclass import configclass


@configclass
class NoiseCfg:
    operation: str = "add"


@configclass
class AdditiveGaussianNoiseCfg(NoiseCfg):
    mean: float = 0.0
    std: float = 1.0


@configclass
class AdditiveUniformNoiseCfg(NoiseCfg):
    n_min: float = -1.0
    n_max: float = 1.0


GaussianNoiseCfg = AdditiveGaussianNoiseCfg
UniformNoiseCfg = AdditiveUniformNoiseCfg
