"""PyTorch-extension front of the C-ABI: ``torch.ops.wheeledlab_b200.{step, step_out, observe_out, reset}``.

The ops (csrc/wl_torch_ops.cpp) take tensors, validate them with ``TORCH_CHECK`` (device, dtype, shape, contiguity: a
``RuntimeError`` names the offending argument), and enqueue the same C entry points as the ctypes host
(``include/wheeledlab_b200.h``) on torch's current CUDA stream of the tensors' device.  They are registered for the CUDA
dispatch key only: a CPU tensor raises ``NotImplementedError`` -- the library has no CPU path.

    from wheeledlab_b200 import torch_ops
    ops = torch_ops.load()                                   # builds wheeledlab_b200/_wl_torch_ops.so on first use
    obs, rew, terminated, truncated = ops.step(sim.handle, actions, t)

``WheeledSim.handle`` is the ``wl_sim*`` as an int.  The reference-side meaning of each op is the corresponding
``ManagerBasedRLEnv`` method."""
from __future__ import annotations

import torch

from .build import TORCH_OPS, build_torch_ops

_loaded = False


def load():
    """Build (if stale) and load the op library; returns ``torch.ops.wheeledlab_b200``."""
    global _loaded
    if not _loaded:
        path = build_torch_ops() if not TORCH_OPS.exists() else TORCH_OPS
        torch.ops.load_library(str(path))
        _loaded = True
    return torch.ops.wheeledlab_b200
