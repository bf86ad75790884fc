"""Minimal stand-in for the part of `gymnasium` the UNMODIFIED WheeledLab sources touch: the registry
(`register / make / spec`, wheeledlab_tasks/__init__.py:14-63, scripts/train_rl.py:70, test/create_and_step_env.py:28),
`ActionWrapper` (wheeledlab_rl/utils/clip_action.py:6), `spaces.Box`, `logger`, `wrappers.RecordVideo`.  It is NOT gymnasium:
no env checker, no vector API, no spaces beyond Box.  Used only when the real package is absent (this image)."""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field

from . import spaces, logger, wrappers  # noqa: F401
from .core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401

__version__ = "0.29.1-standin"


@dataclass
class EnvSpec:
    id: str
    entry_point: object = None
    kwargs: dict = field(default_factory=dict)
    disable_env_checker: bool = True
    max_episode_steps: object = None
    order_enforce: bool = False


registry: dict[str, EnvSpec] = {}


def register(id: str, entry_point=None, kwargs=None, disable_env_checker: bool = True, **extra) -> None:
    registry[id] = EnvSpec(id=id, entry_point=entry_point, kwargs=dict(kwargs or {}), disable_env_checker=disable_env_checker,
                           max_episode_steps=extra.get("max_episode_steps"), order_enforce=extra.get("order_enforce", False))


def spec(id: str) -> EnvSpec:
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")          # gymnasium raises error.NameNotFound (a KeyError-like)
    return registry[id]


def _load(entry_point):
    if callable(entry_point):
        return entry_point
    mod, _, attr = str(entry_point).partition(":")
    return getattr(importlib.import_module(mod), attr)


def make(id, **kwargs):
    """gym.make(id, cfg=env_cfg, render_mode=...): registered kwargs are merged under the call's kwargs and passed to the
    entry point, as gymnasium does."""
    s = spec(id) if isinstance(id, str) else id
    merged = dict(s.kwargs)
    merged.update(kwargs)
    env = _load(s.entry_point)(**merged)
    try:
        env.spec = s
    except Exception:
        pass
    return env
