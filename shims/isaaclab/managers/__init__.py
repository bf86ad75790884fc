"""Manager term cfgs + bases (data carriers; the managers themselves live in wheeledlab_b200)."""
from __future__ import annotations

from dataclasses import MISSING

from ..utils import configclass


@configclass
class SceneEntityCfg:
    name: str = MISSING
    joint_names: object = None
    joint_ids: object = slice(None)
    body_names: object = None
    body_ids: object = slice(None)
    fixed_tendon_names: object = None
    object_collection_names: object = None
    preserve_order: bool = False


@configclass
class ManagerTermBaseCfg:
    func: object = MISSING
    params: dict = dict()


@configclass
class ActionTermCfg:
    class_type: type = MISSING
    asset_name: str = MISSING
    debug_vis: bool = False
    clip: object = None


@configclass
class EventTermCfg(ManagerTermBaseCfg):
    mode: str = MISSING
    interval_range_s: object = None
    is_global_time: bool = False
    min_step_count_between_reset: int = 0


@configclass
class RewardTermCfg(ManagerTermBaseCfg):
    weight: float = MISSING


@configclass
class CurriculumTermCfg(ManagerTermBaseCfg):
    pass


@configclass
class TerminationTermCfg(ManagerTermBaseCfg):
    time_out: bool = False


@configclass
class CommandTermCfg:
    class_type: type = None
    resampling_time_range: tuple = MISSING
    debug_vis: bool = False


@configclass
class ObservationTermCfg(ManagerTermBaseCfg):
    modifiers: object = None
    noise: object = None
    clip: object = None
    scale: object = None
    history_length: int = 0
    flatten_history_dim: bool = True


@configclass
class ObservationGroupCfg:
    concatenate_terms: bool = True
    enable_corruption: bool = False
    history_length: object = None
    flatten_history_dim: bool = True


class ManagerTermBase:
    def __init__(self, cfg, env):
        self.cfg = cfg
        self._env = env

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device

    def reset(self, env_ids=None):
        pass

    def __call__(self, *args):
        raise NotImplementedError


class ActionTerm(ManagerTermBase):
    def __init__(self, cfg, env):
        super().__init__(cfg, env)
        self._asset = self._env.scene[self.cfg.asset_name]

    @property
    def action_dim(self):
        raise NotImplementedError
