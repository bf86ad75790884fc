"""Rollout recording in the format of the reference's playback script.

``play_policy.py:131-165`` appends ``obs`` and ``actions`` every step and saves
``{'observations': [steps, N, D], 'actions': [steps, N, A]}`` with ``torch.save`` as ``<play_name>-rollouts.pt``.  A
rollout slab already holds both stacks, so saving is a view + one file write; ``load_rollouts`` reads either producer's
file.  (The reference appends the env's observation tensor without cloning -- correct there because IsaacLab returns fresh
tensors; slab rows are stable until the next rollout, hence the explicit ``.clone()`` / ``.cpu()`` here.)
"""
from __future__ import annotations

import os

import torch


def save_rollouts(path: str, observations: torch.Tensor, actions: torch.Tensor) -> str:
    """observations [steps, N, D], actions [steps, N, A] (any device) -> torch-saved dict, reference key names."""
    if observations.shape[:2] != actions.shape[:2]:
        raise ValueError("observations and actions must share [steps, num_envs]")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({"observations": observations.detach().cpu().clone(), "actions": actions.detach().cpu().clone()}, path)
    return path


def save_slab(playback_dir: str, play_name: str, slab, steps: int | None = None) -> str:
    """Save the first `steps` rows of a RolloutSlab as ``<playback_dir>/<play_name>-rollouts.pt`` (play_policy.py:160-163)."""
    k = slab.T if steps is None else steps
    return save_rollouts(os.path.join(playback_dir, f"{play_name}-rollouts.pt"), slab.obs[:k], slab.actions[:k])


def load_rollouts(path: str) -> dict:
    data = torch.load(path, map_location="cpu")
    if set(data) != {"observations", "actions"}:
        raise ValueError(f"{path}: not a play_policy rollout file (keys {sorted(data)})")
    return data
