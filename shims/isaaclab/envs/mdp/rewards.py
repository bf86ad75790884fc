import torch


def is_terminated(env):
    return env.termination_manager.terminated.float()


def is_terminated_term(env, term_keys=".*"):
    """[UPSTREAM-RECALL] sum of the named termination terms, zeroed on time-outs."""
    import re
    tm = env.termination_manager
    keys = [term_keys] if isinstance(term_keys, str) else list(term_keys)
    names = [n for n in tm.active_terms if any(re.fullmatch(k, n) for k in keys)]
    reset_buf = torch.zeros(env.num_envs, device=env.device)
    for name in names:
        reset_buf += tm.get_term(name)
    return (reset_buf * (~tm.time_outs)).float()
