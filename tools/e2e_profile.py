#!/usr/bin/env python
"""Where do the microseconds of one env.step_host() go?  (host wall-clock, 3000 steps, N = 4096)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import wheeledlab_b200 as wl
from wheeledlab_b200._lib import lib

E, K = 4096, 3000
env = wl.make("Isaac-MushrDriftRL-v0", num_envs=E)
env.reset()
h_in = env.host_action_buffer
acts = torch.rand(64, E, 2).mul_(2).sub_(1).pin_memory()
for k in range(200):
    env.step_host(h_in)
torch.cuda.synchronize()
def timeit(f, n=K):
    t0 = time.perf_counter()
    for k in range(n):
        f(k)
    return (time.perf_counter() - t0) / n * 1e6
print("host copy of actions      %.2f us" % timeit(lambda k: h_in.copy_(acts[k % 64])))
print("full env.step_host        %.2f us" % timeit(lambda k: env.step_host(h_in)))
io = env._host_io; sim = env.sim
obs, log, p_obs, p_log = env._ring[0]
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
t = [env.common_step_counter]
def raw(k):
    lib.wl_step_host_zero_copy(sim._h, io["_p_h_action"], p_obs, p_log, io["_p_h_result"], t[0], stream); t[0] += 1
print("raw C call (launch+sync)  %.2f us" % timeit(raw))
def raw_async(k):
    lib.wl_step(sim._h, io["_p_h_action"], p_obs, C.c_void_p(io["d_rew"].data_ptr()), C.c_void_p(io["d_terminated"].data_ptr()),
                C.c_void_p(io["d_truncated"].data_ptr()), p_log, t[0], stream); t[0] += 1
t0 = time.perf_counter()
for k in range(K):
    raw_async(k)
torch.cuda.synchronize()
print("async launches, 1 sync    %.2f us/step" % ((time.perf_counter() - t0) / K * 1e6))
print("current_stream lookup     %.2f us" % timeit(lambda k: torch.cuda.current_stream(sim.device).cuda_stream))
