"""``@configclass`` stand-in: class attributes become per-instance deep-copied fields; keyword (or
positional, in declaration order) construction; ``replace / copy / to_dict / from_dict / validate``.
[UPSTREAM-RECALL isaaclab/utils/configclass.py]"""
from __future__ import annotations

import copy
import inspect
import types
from dataclasses import MISSING

_SKIP_TYPES = (types.FunctionType, types.MethodType, property, classmethod, staticmethod)


def _collect_fields(cls):
    names, defaults = [], {}
    for base in reversed(cls.__mro__):
        if base is object:
            continue
        ann = base.__dict__.get("__annotations__", {})
        for key in list(ann.keys()) + [k for k in base.__dict__ if k not in ann]:
            if key.startswith("__") or key in ("_cfg_fields", "_cfg_defaults"):
                continue
            has_val = key in base.__dict__
            val = base.__dict__.get(key, MISSING)
            if has_val and isinstance(val, _SKIP_TYPES):
                continue
            if has_val and isinstance(val, type) and key not in ann:
                continue            # nested class definitions are not fields unless annotated
            if key not in defaults:
                names.append(key)
            defaults[key] = val
    return names, defaults


def _to_dict(obj):
    if hasattr(obj, "_cfg_fields"):
        return {k: _to_dict(getattr(obj, k)) for k in obj._cfg_fields}
    if isinstance(obj, dict):
        return {k: _to_dict(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_dict(v) for v in obj)
    if callable(obj) and not isinstance(obj, type) and hasattr(obj, "__module__") and hasattr(obj, "__name__"):
        return f"{obj.__module__}:{obj.__name__}"
    if isinstance(obj, type):
        return f"{obj.__module__}:{obj.__name__}"
    return obj


def _from_dict(obj, data):
    for k, v in data.items():
        if not hasattr(obj, k):
            raise KeyError(f"[Config]: Key not found under namespace: {k}")
        cur = getattr(obj, k)
        if hasattr(cur, "_cfg_fields") and isinstance(v, dict):
            _from_dict(cur, v)
        elif isinstance(cur, dict) and isinstance(v, dict) and cur and all(hasattr(x, "_cfg_fields") for x in cur.values()):
            for kk, vv in v.items():
                _from_dict(cur[kk], vv) if kk in cur and isinstance(vv, dict) else cur.__setitem__(kk, vv)
        elif callable(cur) and isinstance(v, str) and ":" in v:
            continue                 # function / class references are not overridable from strings here
        else:
            if isinstance(cur, tuple) and isinstance(v, list):
                v = tuple(v)
            setattr(obj, k, v)


def configclass(cls=None, **kwargs):
    def wrap(cls):
        names, defaults = _collect_fields(cls)
        user_post = cls.__dict__.get("__post_init__", None)

        def __init__(self, *args, **kw):
            if len(args) > len(names):
                raise TypeError(f"{cls.__name__}() takes at most {len(names)} positional arguments")
            for k, v in zip(names, args):
                if k in kw:
                    raise TypeError(f"{cls.__name__}() got multiple values for argument '{k}'")
                kw[k] = v
            for k in kw:
                if k not in defaults:
                    raise TypeError(f"{cls.__name__}() got an unexpected keyword argument '{k}'")
            for k in names:
                object.__setattr__(self, k, kw[k] if k in kw else copy.deepcopy(defaults[k]))
            post = getattr(self, "__post_init__", None)
            if post is not None:
                post()

        def replace(self, **kw):
            new = copy.deepcopy(self)
            for k, v in kw.items():
                if k not in new._cfg_fields:
                    raise TypeError(f"{cls.__name__}.replace() got an unexpected field '{k}'")
                setattr(new, k, v)
            return new

        def __repr__(self):
            return f"{cls.__name__}(" + ", ".join(f"{k}={getattr(self, k)!r}" for k in self._cfg_fields) + ")"

        def __eq__(self, other):
            return type(other) is type(self) and all(getattr(self, k) == getattr(other, k) for k in self._cfg_fields)

        if not hasattr(cls, "__post_init__"):
            cls.__post_init__ = lambda self: None          # upstream always installs one (super().__post_init__())
        cls.__init__ = __init__
        cls.replace = replace
        cls.copy = lambda self: copy.deepcopy(self)
        cls.to_dict = lambda self: _to_dict(self)
        cls.from_dict = lambda self, data: _from_dict(self, data)
        cls.validate = lambda self, prefix="": None
        cls.__repr__ = __repr__
        cls.__eq__ = __eq__
        cls.__hash__ = None
        cls._cfg_fields = names
        cls._cfg_defaults = defaults
        # class-level attribute access keeps working (e.g. ``RslRlRunConfig.train.log``)
        return cls

    return wrap(cls) if cls is not None else wrap
