"""sample_space etc. [UPSTREAM-RECALL isaaclab/envs/utils/spaces.py]; unbounded Box dims sample N(0,1)
like gymnasium does."""
import torch


def sample_space(space, device, batch_size=-1, fill_value=None):
    shape = tuple(space.shape)
    if batch_size > 0:
        shape = (batch_size, *shape)
    if fill_value is not None:
        return torch.full(shape, float(fill_value), device=device)
    return torch.randn(shape, device=device)


def replace_env_cfg_spaces_with_strings(env_cfg):
    return env_cfg


def replace_strings_with_env_cfg_spaces(env_cfg):
    return env_cfg
