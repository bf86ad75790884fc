// wl_api.cu -- kernels + the extern "C" ABI declared in include/wheeledlab_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 (see build.py).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>

#include <cuda.h>

#include "wl_step.cuh"

using namespace wl;

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
struct wl_sim {
    wl_config cfg;
    float4* state;          // groups
    wl_globals* globals;    // device
    const float* hf;        // device height-field or null
    size_t state_bytes;
    int64_t launches;
    int64_t last_t;         // counter of the most recent step launched through this handle (-1: none); host mirror
    int64_t base_host;      // host mirror of wl_globals.step_base
    uint8_t* term_bits;     // optional per-env termination-term bits output of wl_step (wl_set_term_bits)
    PeerFan fan;            // peers every output row of wl_step is also stored to (wl_set_peer_fanout); n = 0: none
    const void* hp_host[8]; // wl_step_host_zero_copy: pinned host blocks already resolved to their device alias
    void* hp_dev[8];
    int hp_n;
    int device;             // CUDA device ordinal the handle lives on
    int obs_dim;
    int variant;            // 0 auto, 1 thread-per-env, 4 quad-per-env
    CUtensorMap tmap;       // 2-D tensor map over the height-field (elevation task)
    bool has_tmap;
    int scan_mode;          // ray-cast tile staging: 1 one TMA tile per CTA (default), 2 TMA producer/consumer pipeline, 0 plain loads (A/B)
};

// below this many envs one thread/env cannot fill 148 SMs x 4 schedulers; use 4 lanes per env
#define WL_QUAD_MAX_ENVS (148 * 4 * 32 * 2)

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(WL_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static inline void ensure_device(const wl_sim* sim) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != sim->device) cudaSetDevice(sim->device);   // launches go to the handle's device
}
static inline size_t groups_bytes(int n) { return (size_t)WL_NUM_GROUPS * (size_t)n * 16u; }
static inline size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// ---- per-step bookkeeping without a grid-wide sync (layout + protocol: wl_globals in the header) ---------------------
// step counter of this launch: the host's value, or (negative host value = -1 - k) device base + k
__device__ __forceinline__ uint32_t decode_step(const wl_globals* __restrict__ gl, uint32_t t_arg) {
    return ((int32_t)t_arg < 0) ? __ldcg(&gl->step_base) + (0xFFFFFFFFu - t_arg) : t_arg;
}
// Programmatic dependent launch (sm_90+): a step kernel lets the next launch of the stream start its prologue right away
// (launch_dependents first thing) and itself waits for the previous kernel's memory to be final (wait) before the first
// read of anything a previous launch may have written.  Both are no-ops for a launch without the PDL attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// W(t): the weights step t-1 used, plus increase_reward_weight_over_time (curriculums.py:23-35) for the counter value t
// reached by the step that just ended -- the reference calls it from _reset_idx, i.e. only if >= 1 env reset in step t-1.
// Every thread of the launch evaluates this identically (no thread writes what another one reads here).
__device__ __forceinline__ void load_weights(const wl_config& c, const wl_globals* __restrict__ gl, uint32_t t, float wts[WL_MAX_REW_TERMS]) {
    const float4* wp = reinterpret_cast<const float4*>(gl->rew_weight[(t + 1u) & 1u]);
    const float4 w0 = __ldcg(wp), w1 = __ldcg(wp + 1);
    wts[0] = w0.x; wts[1] = w0.y; wts[2] = w0.z; wts[3] = w0.w; wts[4] = w1.x; wts[5] = w1.y; wts[6] = w1.z; wts[7] = w1.w;
    if (c.curr_n > 0 && t != 0u && (t % (uint32_t)c.max_episode_length) == 0u) {
        if (__ldcg(&gl->acc[(t + 2u) % 3u][8]) > 0.0f && __ldcg(&gl->curr_applied_t) != t) {     // (not already applied in place by the host path)
            const int E = (int)(t / (uint32_t)c.max_episode_length);
            for (int k = 0; k < c.curr_n; ++k) {
                if (E / c.curr_every[k] > c.curr_max[k]) continue;
                if ((E + 1) % c.curr_every[k] == 0) {
                    const int slot = c.curr_slot[k];
#pragma unroll
                    for (int q = 0; q < WL_MAX_REW_TERMS; ++q) if (q == slot) wts[q] += c.curr_inc[k];
                }
            }
        }
    }
}
// extras["log"] row of the step whose accumulators are `row`: means of the episode sums over the reset envs divided by
// max_episode_length_s (RewardManager.reset), then the counts.  Lanes 0..15 of one warp, one element each.
__device__ __forceinline__ void publish_log_row(const wl_config& c, wl_globals* __restrict__ gl, uint32_t row, int lane) {
    const float v = (lane < 16) ? __ldcg(&gl->acc[row][lane]) : 0.0f;
    const float cnt = __shfl_sync(0xffffffffu, v, 8);
    float* lp = gl->log_ptr[row];
    float out = (lane < WL_MAX_REW_TERMS) ? v / (r_max(cnt, 1.0f) * c.episode_length_s) : v;
    if (lane < 16) {
        if (cnt > 0.0f) gl->last_log[lane] = out;            // extras["log"] persists until the next step that resets an env
        else out = gl->last_log[lane];
        if (lp != nullptr) lp[lane] = out;
    }
}
// The janitor: ONE warp of the launch -- the first warp of an EXTRA CTA appended to the grid (blockIdx.x == gridDim.x - 1)
// that owns no envs, so that no env warp starts its dependent chain late.
__device__ __forceinline__ void janitor(const wl_config& c, wl_globals* __restrict__ gl, uint32_t t, const float wts[WL_MAX_REW_TERMS],
                                        float* __restrict__ d_log) {
    const int lane = threadIdx.x & 31;
    if (t != 0u) publish_log_row(c, gl, (t + 2u) % 3u, lane);          // step t-1 (complete: its kernel has finished)
    if (lane < 16) gl->acc[(t + 1u) % 3u][lane] = 0.0f;                // re-arm the row step t+1 will use
    float w = wts[0];
#pragma unroll
    for (int q = 1; q < WL_MAX_REW_TERMS; ++q) w = (lane == q) ? wts[q] : w;
    if (lane < WL_MAX_REW_TERMS) gl->rew_weight[t & 1u][lane] = w;     // W(t), read by step t+1
    if (lane == 0) gl->log_ptr[t % 3u] = d_log;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Finished envs -> this step's accumulator row: warp-shuffle reduce, one RED set per warp (no fence, no ticket: the
// kernel boundary publishes the row to the next launch).  acc layout: [0..7] episode sums, [8] #reset, [9+j] #envs whose
// termination term j fired (tmask bit j).
__device__ __forceinline__ bool log_accumulate(float* __restrict__ acc, bool contrib, uint32_t tmask,
                                               const float sums[WL_MAX_REW_TERMS]) {
    const unsigned any_c = __ballot_sync(0xffffffffu, contrib);
    if (!any_c) return false;
    float vals[16];
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS; ++k) vals[k] = contrib ? sums[k] : 0.0f;
    vals[8] = contrib ? 1.0f : 0.0f;
#pragma unroll
    for (int j = 0; j < WL_MAX_TERM_TERMS; ++j) vals[9 + j] = (contrib && ((tmask >> j) & 1u)) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) vals[k] = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) if (vals[k] != 0.0f) atomicAdd(&acc[k], vals[k]);
    }
    return true;
}

// One thread per env.  TASK selects the MDP + terrain at compile time.
#ifndef WL_STEP_MIN_BLOCKS
#define WL_STEP_MIN_BLOCKS 4      // 4 CTAs x 128 threads / SM (<= 128 registers); see profiles/ for the occupancy A/B
#endif
// STAGE 0: the whole env.step.  STAGE 1 / 2: the same code cut after section E (wl_step_stage_a / wl_step_stage_b), so that
// host-side (Python) reward and termination terms can run between "rewards" and "reset" exactly where the reference's
// managers would call them; the state round-trips through HBM losslessly, so 1 + 2 == 0 bit for bit.
struct StageIO { uint8_t* tmask; const uint8_t* extra_terminated; const uint8_t* extra_truncated; };
template <int TASK, int STAGE = 0>
__global__ void __launch_bounds__(128, WL_STEP_MIN_BLOCKS)
wl_step_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
               const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ rew,
               uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t_arg,
               StageIO sio = StageIO{nullptr, nullptr, nullptr}) {
    constexpr bool ELEV = (TASK == WL_TASK_ELEVATION), VIS = (TASK == WL_TASK_VISUAL);
    pdl_trigger();
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    const uint32_t t = decode_step(gl, t_arg);
    const VisualMap vm = VIS ? visual_map(c, T.hf) : VisualMap{nullptr, nullptr};
    bool done = false;
    uint32_t tmask = 0u;
    EnvState e;
    float wts[WL_MAX_REW_TERMS];            // W(t): issued with the state loads so the latency hides behind the integrator
    load_weights(c, gl, t, wts);
    if (STAGE != 1 && blockIdx.x == gridDim.x - 1) {          // the janitor CTA (appended to the grid, owns no envs)
        if (threadIdx.x < 32) janitor(c, gl, t, wts, d_log);
        return;
    }
    if (STAGE == 2) {
        if (i < n) {
            load_env(st, n, i, e, ELEV, c.dr_wheel_mass_enable != 0, c.d_inv_Iw);
            tmask = sio.tmask[i];
            if (sio.extra_truncated && sio.extra_truncated[i]) tmask |= 1u;            // host-side time-out style term
            if (sio.extra_terminated && sio.extra_terminated[i]) tmask |= 0x80u;       // host-side termination term
            terminated_o[i] = (tmask & ~1u) ? 1 : 0;
            truncated_o[i] = (tmask & 1u) ? 1 : 0;
            done = tmask != 0u;
        }
    } else if (i < n) {
        load_env(st, n, i, e, ELEV, c.dr_wheel_mass_enable != 0, c.d_inv_Iw);
        // A. action manager
        float2 a = action[i];
        e.prev_action[0] = e.action[0]; e.prev_action[1] = e.action[1];
        e.action[0] = a.x; e.action[1] = a.y;
        float wheel_target[4], steer_target[2];
        process_action(c, a.x, a.y, wheel_target, steer_target);
        // B. decimation x (actuators -> integrator)
        Chassis b;
        M3 R = rotmat(e.qw, e.qx, e.qy, e.qz);
        V3 cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
        b.pc = V3{e.p.x + cw.x, e.p.y + cw.y, e.p.z + cw.z};
        b.v = e.v; b.qw = e.qw; b.qx = e.qx; b.qy = e.qy; b.qz = e.qz;
        b.wb = rotT(R, e.w);
        StepConsts kc = make_step_consts<4>(c, e);
        for (int d = 0; d < c.decimation; ++d) {
            float lo[4], hi[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) dc_limits(c, c.dc_effort[w], e.omega[w], lo[w], hi[w]);
            for (int j = 0; j < c.substeps; ++j) physics_substep<TASK, 1>(c, T, e, b, wheel_target, lo, hi, steer_target, kc);
        }
        R = rotmat(b.qw, b.qx, b.qy, b.qz);
        cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
        e.p = V3{b.pc.x - cw.x, b.pc.y - cw.y, b.pc.z - cw.z};
        e.v = b.v; e.qw = b.qw; e.qx = b.qx; e.qy = b.qy; e.qz = b.qz;
        e.w = rot(R, b.wb);
        // C. counters   D. terminations   E. rewards (value = f*w*dt, skipped when w == 0)
        e.ep_len += 1;
        const bool time_out = e.ep_len >= c.max_episode_length;
        float f[WL_MAX_REW_TERMS];
        V3 vb = rotT(R, e.v);
        if (ELEV) {
            tmask = elev_terms(c, e, R, vb, (e.omega[0] + e.omega[1]) + (e.omega[2] + e.omega[3]), time_out, f);
        } else if (VIS) {
            tmask = visual_terms(c, vm, e, vb, time_out, f);
        } else {
            const bool oob = drift_off_track(c, e.p.x, e.p.y);
            drift_reward_terms(c, e.steer[0], e.steer[1], det_atan2(vb.y, vb.x), e.p, vb, b.wb, e.w.z, oob, time_out, f);
            tmask = (time_out ? 1u : 0u) | (oob ? 2u : 0u);
        }
        tmask &= (uint32_t)c.term_enable;
        float total = 0.0f;
#pragma unroll
        for (int k = 0; k < WL_MAX_REW_TERMS; ++k) {
            if (k < c.num_rew_terms) {
                float w = wts[k];
                if (w != 0.0f) { float val = f[k] * w * c.d_step_dt; total += val; e.sums[k] += val; }
            }
        }
        rew[i] = total;
        if (STAGE == 1) {                    // cut: state (with this step's episode sums) back to HBM, term bits for stage 2
            sio.tmask[i] = (uint8_t)tmask;
            store_env(st, n, i, e, ELEV);
            return;
        }
        terminated_o[i] = (tmask & ~1u) ? 1 : 0;
        truncated_o[i] = (tmask & 1u) ? 1 : 0;
        if (STAGE == 0 && sio.tmask != nullptr) sio.tmask[i] = (uint8_t)tmask;      // per-term masks (wl_set_term_bits)
        done = tmask != 0u;
    }
    if (STAGE == 1) return;
    // F. auto-reset + per-step episode log (warp-shuffle reduction over the finished envs)
    log_accumulate(gl->acc[t % 3u], done, tmask, e.sums);
    if (i < n) {
        const uint32_t gid = (uint32_t)(c.env_id_offset + i);
        if (done) {
            if (ELEV) elev_reset_env(c, e, gid, t); else if (VIS) visual_reset_env(c, vm, e, gid, t); else drift_reset_env(c, e, gid, t);
        }
        // G. commands   H. interval events (post-reset state)   I. observations
        if (VIS) {
            float o[8]; visual_proprio(c, e, o);
            const int camf = vis_cam_floats(c);              // the camera floats come first (wl_camera_kernel fills them)
            float4* row = reinterpret_cast<float4*>(obs + (size_t)(WL_OBS_DIM_VISUAL + camf) * i + camf);
            row[0] = make_float4(o[0], o[1], o[2], o[3]); row[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else if (ELEV) {
            elev_command_update(c, e, gid, t, c.d_step_dt);
            float o[13]; elev_proprio(c, e, euler_xyz(e.qw, e.qx, e.qy, e.qz), o);
            float* row = obs + (size_t)WL_OBS_DIM_ELEV * i;
#pragma unroll
            for (int k = 0; k < 13; ++k) row[k] = o[k];
        } else {
            interval_pushes(c, e, gid, t, c.d_step_dt);
            blind_obs(c, e, gid, t, RNG_OBS, 0u, obs + (size_t)WL_OBS_DIM_BLIND * i);
        }
        store_env(st, n, i, e, ELEV);
    }
}

// Four lanes per env (lane = wheel).  Same arithmetic, same results; the per-wheel work runs in parallel and the
// chassis is integrated redundantly in the 4 lanes.  Used when N is too small to fill the chip with one thread/env.
// quad_env_step = sections A..I of ONE env.step() on register-resident state `e` (lane-local view, see load_env_quad).
// lane w of an env's quad computes ONE of the three euler angles (lane 0 roll, 1 pitch, 2 and 3 yaw), wrapped to [0, 2 pi)
__device__ __forceinline__ float euler_lane(float qw, float qx, float qy, float qz, int w) {
    float sin_roll = 2.0f * fm(qw, qx, qy * qz), cos_roll = fm(-2.0f, fm(qx, qx, qy * qy), 1.0f);
    float sin_pitch = 2.0f * fm(qw, qy, -(qz * qx));
    float sin_yaw = 2.0f * fm(qw, qz, qx * qy), cos_yaw = fm(-2.0f, fm(qy, qy, qz * qz), 1.0f);
    float ay = (w == 0) ? sin_roll : (w == 1) ? sin_pitch : sin_yaw;
    float ax = (w == 0) ? cos_roll : (w == 1) ? sqrtf((1.0f - sin_pitch) * (1.0f + sin_pitch)) : cos_yaw;
    float ang = det_atan2(ay, ax);
    if (w == 1 && fabsf(sin_pitch) >= 1.0f) ang = (sin_pitch < 0.0f) ? -1.57079632679489661923f : 1.57079632679489661923f;
    return wrap_2pi(ang);
}

template <int TASK>
__device__ __forceinline__ void quad_env_step(const wl_config& c, const Terrain& T, const VisualMap& vm, float* __restrict__ acc_row,
                                              const float wts[WL_MAX_REW_TERMS], EnvState& e, int i, int w, bool live, uint32_t gid,
                                              unsigned base, uint32_t t, float2 a, const float znoise[4], float* __restrict__ obs_row,
                                              float* __restrict__ rew, uint8_t* __restrict__ terminated_o,
                                              uint8_t* __restrict__ truncated_o, uint8_t* __restrict__ term_bits = nullptr) {
    constexpr bool ELEV = (TASK == WL_TASK_ELEVATION), VIS = (TASK == WL_TASK_VISUAL);
    // A. action manager (redundant in the 4 lanes)
    e.prev_action[0] = e.action[0]; e.prev_action[1] = e.action[1];
    e.action[0] = a.x; e.action[1] = a.y;
    float wheel_target[4], steer_target[2];
    process_action(c, a.x, a.y, wheel_target, steer_target);
    const float my_target = (w == 0) ? wheel_target[0] : (w == 1) ? wheel_target[1] : (w == 2) ? wheel_target[2] : wheel_target[3];
    const float my_effort = (w == 0) ? c.dc_effort[0] : (w == 1) ? c.dc_effort[1] : (w == 2) ? c.dc_effort[2] : c.dc_effort[3];
    float my_steer_target[2] = {(w == 3) ? steer_target[1] : steer_target[0], 0.0f};
    // B. decimation x (actuators -> integrator)
    Chassis b;
    M3 R = rotmat(e.qw, e.qx, e.qy, e.qz);
    V3 cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    b.pc = V3{e.p.x + cw.x, e.p.y + cw.y, e.p.z + cw.z};
    b.v = e.v; b.qw = e.qw; b.qx = e.qx; b.qy = e.qy; b.qz = e.qz;
    b.wb = rotT(R, e.w);
    StepConsts kc = make_step_consts<1>(c, e);
    const float my_targets[4] = {my_target, 0.0f, 0.0f, 0.0f};
    for (int d = 0; d < c.decimation; ++d) {
        float lo[4], hi[4];
        dc_limits(c, my_effort, e.omega[0], lo[0], hi[0]);
        for (int j = 0; j < c.substeps; ++j) physics_substep<TASK, 4>(c, T, e, b, my_targets, lo, hi, my_steer_target, kc);
    }
    R = rotmat(b.qw, b.qx, b.qy, b.qz);
    cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    e.p = V3{b.pc.x - cw.x, b.pc.y - cw.y, b.pc.z - cw.z};
    e.v = b.v; e.qw = b.qw; e.qx = b.qx; e.qy = b.qy; e.qz = b.qz;
    e.w = rot(R, b.wb);
    // C./D./E. counters, terminations, rewards (redundant in the 4 lanes)
    e.ep_len += 1;
    const bool time_out = e.ep_len >= c.max_episode_length;
    V3 vb = rotT(R, e.v);
    uint32_t tmask;
    float f[WL_MAX_REW_TERMS];
    if (ELEV) {
        float o01 = e.omega[0] + __shfl_xor_sync(0xffffffffu, e.omega[0], 1);      // (w0+w1), (w2+w3)
        float osum = o01 + __shfl_xor_sync(0xffffffffu, o01, 2);                   // (w0+w1)+(w2+w3)
        tmask = elev_terms(c, e, R, vb, osum, time_out, f);
    } else if (VIS) {
        tmask = visual_terms(c, vm, e, vb, time_out, f);
    } else {
        const float steer_l = __shfl_sync(0xffffffffu, e.steer[0], base + 2), steer_r = __shfl_sync(0xffffffffu, e.steer[0], base + 3);
        const bool oob = drift_off_track(c, e.p.x, e.p.y);
        drift_reward_terms(c, steer_l, steer_r, det_atan2(vb.y, vb.x), e.p, vb, b.wb, e.w.z, oob, time_out, f);
        tmask = (time_out ? 1u : 0u) | (oob ? 2u : 0u);
    }
    // I'. observation inputs from the post-physics state, formed HERE so that this lane's euler atan2 overlaps the reward
    // terms' atan2 / sqrt chains in one basic block; an env that resets (or gets pushed) redoes them below -- the rare path
    float eu_k = 0.0f;
    V3 wbo{0.0f, 0.0f, 0.0f};
    if (!VIS) eu_k = euler_lane(e.qw, e.qx, e.qy, e.qz, w);
    if (!VIS && !ELEV) wbo = rotT(R, e.w);                        // (R^T of the WORLD rate, as the oracle forms it: not b.wb bit for bit)
    tmask &= (uint32_t)c.term_enable;
    float total = 0.0f;
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS; ++k) {
        if (k < c.num_rew_terms) {
            float wgt = wts[k];
            if (wgt != 0.0f) { float val = f[k] * wgt * c.d_step_dt; total += val; e.sums[k] += val; }
        }
    }
    const bool done = tmask != 0u;
    if (live && w == 0) {
        rew[i] = total; terminated_o[i] = (tmask & ~1u) ? 1 : 0; truncated_o[i] = (tmask & 1u) ? 1 : 0;
        if (term_bits != nullptr) term_bits[i] = (uint8_t)tmask;      // per-term masks for TerminationManager.get_term
    }
    // F. per-step episode log: one contribution per env (lane 0 of each live quad), then auto-reset
    log_accumulate(acc_row, done && live && (w == 0), tmask, e.sums);
    if (done) {
        if (ELEV) elev_reset_env(c, e, gid, t); else if (VIS) visual_reset_env(c, vm, e, gid, t); else drift_reset_env(c, e, gid, t);
        if (!VIS) eu_k = euler_lane(e.qw, e.qx, e.qy, e.qz, w);
        if (!VIS && !ELEV) { R = rotmat(e.qw, e.qx, e.qy, e.qz); vb = rotT(R, e.v); wbo = rotT(R, e.w); }   // (the oracle's own operations: signed zeros)
    }
    if (ELEV) elev_command_update(c, e, gid, t, c.d_step_dt);
    else if (!VIS) {
        // H. interval pushes: never on the step an env resets (fresh timers exceed step_dt), so R is still the rotation of e.q
        if (interval_pushes(c, e, gid, t, c.d_step_dt)) { vb = rotT(R, e.v); wbo = rotT(R, e.w); }
    }
    if (VIS) {      // 8 proprioceptive floats, no noise, no euler: lanes 0 and 1 write one float4 each
        float o[8]; visual_proprio(c, e, o);
        if (live && w < 2) {
            float4* row = reinterpret_cast<float4*>(obs_row + vis_cam_floats(c));
            row[w] = (w == 0) ? make_float4(o[0], o[1], o[2], o[3]) : make_float4(o[4], o[5], o[6], o[7]);
        }
    } else if (ELEV) {
        V3 eu{__shfl_sync(0xffffffffu, eu_k, base + 0), __shfl_sync(0xffffffffu, eu_k, base + 1), __shfl_sync(0xffffffffu, eu_k, base + 2)};
        float o[13]; elev_proprio(c, e, eu, o);
        if (live) {
            // lane w writes o[w], o[w+4], o[w+8] (and lane 0 also o[12]): 4-byte rows, scalar stores
            const float v0 = (w == 0) ? o[0] : (w == 1) ? o[1] : (w == 2) ? o[2] : o[3];
            const float v1 = (w == 0) ? o[4] : (w == 1) ? o[5] : (w == 2) ? o[6] : o[7];
            const float v2 = (w == 0) ? o[8] : (w == 1) ? o[9] : (w == 2) ? o[10] : o[11];
            obs_row[w] = v0; obs_row[w + 4] = v1; obs_row[w + 8] = v2;
            if (w == 0) obs_row[12] = o[12];
        }
    } else {
        // blind_obs_quad shuffles: every lane of the warp calls it, dead quads only skip the stores
        blind_obs_quad(c, e, w, eu_k, vb, wbo, znoise, obs_row, live);
    }
}

template <int TASK>
__global__ void __launch_bounds__(128)
wl_step_quad_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
                    const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ rew,
                    uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t_arg,
                    uint8_t* __restrict__ term_bits) {
    constexpr bool ELEV = (TASK == WL_TASK_ELEVATION), VIS = (TASK == WL_TASK_VISUAL);
    pdl_trigger();
    const VisualMap vm = VIS ? visual_map(c, T.hf) : VisualMap{nullptr, nullptr};
    const int n = c.num_envs;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 2, w = tid & 3;
    const bool live = i < n;                 // a whole quad is live or not (blockDim is a multiple of 4)
    const int ii = live ? i : n - 1;         // dead quads shadow the last env (no stores) so shuffles stay convergent
    const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
    const unsigned base = (threadIdx.x & 31u) & ~3u;
    pdl_wait();
    const uint32_t t = decode_step(gl, t_arg);
    float wts[WL_MAX_REW_TERMS];
    if (blockIdx.x == gridDim.x - 1) {       // the janitor CTA (appended to the grid, owns no envs)
        if (threadIdx.x < 32) { load_weights(c, gl, t, wts); janitor(c, gl, t, wts, d_log); }
        return;
    }
    EnvState e;
    QuadRaw raw;
    quad_issue_loads(st, n, ii, w, raw, ELEV, c.dr_wheel_mass_enable != 0);
    const float2 a = action[ii];
    load_weights(c, gl, t, wts);
    // the observation noise depends on (env id, step) only: drawn here, while the state loads are in flight (nothing above
    // has consumed a loaded value yet)
    float zn[4];
    if (!ELEV && !VIS) quad_obs_noise(c, w, gid, t, RNG_OBS, 0u, zn); else zn[0] = zn[1] = zn[2] = zn[3] = 0.0f;
    quad_unpack(raw, w, e, ELEV, c.dr_wheel_mass_enable != 0, c.d_inv_Iw);
    const int od = ELEV ? WL_OBS_DIM_ELEV : VIS ? WL_OBS_DIM_VISUAL + vis_cam_floats(c) : WL_OBS_DIM_BLIND;
    quad_env_step<TASK>(c, T, vm, gl->acc[t % 3u], wts, e, i, w, live, gid, base, t, a, zn, obs + (size_t)od * ii, rew, terminated_o, truncated_o, term_bits);
    if (live) store_env_quad(st, n, i, w, e, ELEV);
}

// ---------------------------------------------------------------------------------------
// Drift-family step in the latency-optimised geometry for small N: CTA = 2 warps for 8 envs.  Warp 0 (the "env warp")
// is the quad (lane = wheel) and carries ONLY what is on the dependent chain: state loads -> integrator sub-steps ->
// terminations -> reset -> pushes -> observations -> state stores.  Warp 1 (the "aux warp", lane = env) does everything
// that is not on that chain, concurrently on another scheduler: the action map (ready before the first sub-step needs the
// targets), the observation-noise and push draws (counter based: functions of (env, step) only) and -- once the env warp
// has published the post-physics state in shared memory -- reward terms, episode sums, reward / done stores and the
// episode log.  Hand-offs are named barriers (bar.arrive / bar.sync), each used once per launch.  Same arithmetic as
// wl_step_quad_kernel, operation for operation.
// ---------------------------------------------------------------------------------------
#define WL_DUO_ENVS 8
struct DuoShared {
    float tgt[WL_DUO_ENVS][8];      // wheel_target[4], steer_target[2]
    float noise[WL_DUO_ENVS][12];   // observation noise normals (Philox blocks 0..2)
    float push[WL_DUO_ENVS][8];     // hf fired, dvx, dvy, dwz_hf, t_hf_new, lf fired, dwz_lf, t_lf_new
    float fin[WL_DUO_ENVS][16];     // post-physics: p(3) vb(3) wb(3) wz_world steer_l steer_r raw-tmask(bits)
    __align__(16) float obs[WL_DUO_ENVS * WL_OBS_DIM_BLIND];   // the 8 observation rows of the CTA, staged for 128-bit stores
    __align__(16) float rewrow[WL_DUO_ENVS];                    // the CTA's 8 rewards and 2 x 8 done bytes, staged the same way
    __align__(8) uint8_t maskrow[2][WL_DUO_ENVS];
};
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// FAN = 0: no peer: the fan-out / multicast / packed-row code is not in the kernel at all (the dependent stream of one warp per
// scheduler also pays for instruction fetch: 770 static instructions fewer on the single-GPU path, 0.4 us per step).
// FAN = 2: multicast only -- the host has checked that num_envs is a multiple of 8 and the output rows are aligned, so every row
// is a full wide store: local copy + ONE multimem.st, no per-peer loops, no element-wise fallback in the instruction stream.
// FAN = 1: the general fan-out (per-peer stores; multicast for the rows that qualify).
template <int FAN>
__global__ void __launch_bounds__(64)
wl_step_duo_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl,
                   const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ rew,
                   uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t_arg,
                   uint8_t* __restrict__ term_bits, const __grid_constant__ PeerFan pf) {
    __shared__ DuoShared sh;
    const int n = c.num_envs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = decode_step(gl, t_arg);
    if (blockIdx.x == gridDim.x - 1) {       // the janitor CTA (appended to the grid, owns no envs)
        if (threadIdx.x < 32) { float wj[WL_MAX_REW_TERMS]; load_weights(c, gl, t, wj); janitor(c, gl, t, wj, d_log); }
        return;
    }
    const int env0 = blockIdx.x * WL_DUO_ENVS;
    const Terrain T{nullptr};
    if (warp == 1) {
        // ================= aux warp: lane j < 8 <-> env env0 + j =================
        const int j = lane & 7;
        const int i = env0 + j;
        const bool live = (lane < 8) && (i < n);
        const int ii = (i < n) ? i : n - 1;
        const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
        float2 a = make_float2(0.0f, 0.0f);
        float4 s0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), s1 = s0;
        float t_hf = 0.0f, t_lf = 0.0f;
        if (lane < 8) {
            a = action[ii];
            s0 = ldg4(st, WL_G_SUM0, n, ii); s1 = ldg4(st, WL_G_SUM1, n, ii);
            const float* fst = reinterpret_cast<const float*>(st);
            t_hf = fst[((size_t)WL_G_LINVEL * n + ii) * 4 + 3]; t_lf = fst[((size_t)WL_G_ANGVEL * n + ii) * 4 + 3];
        }
        float wts[WL_MAX_REW_TERMS];
        load_weights(c, gl, t, wts);
        // A. action map -> targets of the whole env step
        if (lane < 8) {
            float wt[4], stt[2];
            process_action(c, a.x, a.y, wt, stt);
            sh.tgt[j][0] = wt[0]; sh.tgt[j][1] = wt[1]; sh.tgt[j][2] = wt[2]; sh.tgt[j][3] = wt[3];
            sh.tgt[j][4] = stt[0]; sh.tgt[j][5] = stt[1];
        }
        bar_arrive(1);
        // N. observation noise: lane = (env, Philox block)
        if (lane < 24) {
            const int jn = lane / 3, k = lane - 3 * jn;
            const int in_ = (env0 + jn < n) ? env0 + jn : n - 1;
            float z[4];
            quad_obs_noise(c, k, (uint32_t)(c.env_id_offset + in_), t, RNG_OBS, 0u, z);
            sh.noise[jn][4 * k + 0] = z[0]; sh.noise[jn][4 * k + 1] = z[1]; sh.noise[jn][4 * k + 2] = z[2]; sh.noise[jn][4 * k + 3] = z[3];
        }
        // P. interval pushes of an env that does NOT reset this step (the env warp ignores them otherwise)
        if (lane < 8) {
            float hf = 0.0f, dvx = 0.0f, dvy = 0.0f, dwh = 0.0f, lf = 0.0f, dwl = 0.0f;
            if (c.push_enable) {
                t_hf = t_hf - c.d_step_dt;
                if (t_hf < 1.0e-6f) {
                    uint4 r = philox4x32(c.seed, gid, t, RNG_PUSH_HF, 0u);
                    dvx = uniform(r.x, -c.push_hf_range[0], c.push_hf_range[0]);
                    dvy = uniform(r.y, -c.push_hf_range[1], c.push_hf_range[1]);
                    dwh = uniform(r.z, -c.push_hf_range[2], c.push_hf_range[2]);
                    t_hf = uniform(r.w, c.push_hf_interval[0], c.push_hf_interval[1]);
                    hf = 1.0f;
                }
                t_lf = t_lf - c.d_step_dt;
                if (t_lf < 1.0e-6f) {
                    uint4 r = philox4x32(c.seed, gid, t, RNG_PUSH_LF, 0u);
                    dwl = uniform(r.x, -c.push_lf_yaw, c.push_lf_yaw);
                    t_lf = uniform(r.y, c.push_lf_interval[0], c.push_lf_interval[1]);
                    lf = 1.0f;
                }
            }
            sh.push[j][0] = hf; sh.push[j][1] = dvx; sh.push[j][2] = dvy; sh.push[j][3] = dwh; sh.push[j][4] = t_hf;
            sh.push[j][5] = lf; sh.push[j][6] = dwl; sh.push[j][7] = t_lf;
        }
        bar_arrive(3);
        // R. rewards, episode sums, reward / done stores, episode log -- on the post-physics state the env warp publishes
        bar_sync(2);
        uint32_t tmask = 0u;
        float sums[WL_MAX_REW_TERMS] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        bool done = false;
        if (lane < 8) {
            const float* fi = sh.fin[j];
            const V3 p{fi[0], fi[1], fi[2]}, vb{fi[3], fi[4], fi[5]}, wb{fi[6], fi[7], fi[8]};
            const uint32_t raw = __float_as_uint(fi[12]);
            const bool time_out = (raw & 1u) != 0u, oob = (raw & 2u) != 0u;
            float f[WL_MAX_REW_TERMS];
            drift_reward_terms(c, fi[10], fi[11], det_atan2(vb.y, vb.x), p, vb, wb, fi[9], oob, time_out, f);
            tmask = raw & (uint32_t)c.term_enable;
            float total = 0.0f;
#pragma unroll
            for (int k = 0; k < WL_MAX_REW_TERMS; ++k) {
                if (k < c.num_rew_terms) {
                    float wgt = wts[k];
                    if (wgt != 0.0f) { float val = f[k] * wgt * c.d_step_dt; total += val; sums[k] += val; }
                }
            }
            done = tmask != 0u;
            // fan-out active, full CTA, aligned rows: 32 + 8 + 8 contiguous bytes leave as four wide stores (one NVLink /
            // multicast write each instead of 24 small ones); otherwise element by element (without peers the staging round
            // trip sits at the kernel's tail and costs 0.17 us per step, profiles/r02_kexp_pack.txt)
            const bool packed = (FAN == 2) || ((FAN == 1) && ((pf.n | pf.mc) != 0) && (env0 + WL_DUO_ENVS <= n) && ((reinterpret_cast<uintptr_t>(rew + env0) & 15u) == 0) &&
                                (((reinterpret_cast<uintptr_t>(terminated_o + env0) | reinterpret_cast<uintptr_t>(truncated_o + env0)) & 7u) == 0));
            const uint8_t tb = (uint8_t)((tmask & ~1u) ? 1 : 0), ub = (uint8_t)((tmask & 1u) ? 1 : 0);
            if (packed) {
                sh.rewrow[lane] = total; sh.maskrow[0][lane] = tb; sh.maskrow[1][lane] = ub;
                __syncwarp(0xffu);
                if (lane < 2) fan_store16<FAN == 2>(pf, reinterpret_cast<float4*>(rew + env0) + lane, reinterpret_cast<const float4*>(sh.rewrow)[lane]);
                else if (lane == 2) fan_store8<FAN == 2>(pf, reinterpret_cast<float2*>(terminated_o + env0), *reinterpret_cast<const float2*>(sh.maskrow[0]));
                else if (lane == 3) fan_store8<FAN == 2>(pf, reinterpret_cast<float2*>(truncated_o + env0), *reinterpret_cast<const float2*>(sh.maskrow[1]));
            } else if (live) {
                if (FAN == 1) { fan_store(pf, &rew[i], total); fan_store(pf, &terminated_o[i], tb); fan_store(pf, &truncated_o[i], ub); }
                else { rew[i] = total; terminated_o[i] = tb; truncated_o[i] = ub; }
            }
            if (live && term_bits != nullptr) term_bits[i] = (uint8_t)tmask;
        }
        log_accumulate(gl->acc[t % 3u], done && live, tmask, sums);
        if (live) {
            if (done) {
#pragma unroll
                for (int k = 0; k < WL_MAX_REW_TERMS; ++k) sums[k] = 0.0f;
            }
            stg4(st, WL_G_SUM0, n, i, make_float4(sums[0], sums[1], sums[2], sums[3]));
            stg4(st, WL_G_SUM1, n, i, make_float4(sums[4], sums[5], sums[6], sums[7]));
        }
        return;
    }
    // ================= env warp: the quad (lane = wheel), 8 envs =================
    const int q = lane >> 2, w = lane & 3;
    const int i = env0 + q;
    const bool live = i < n;
    const int ii = live ? i : n - 1;
    const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
    const unsigned base = (unsigned)lane & ~3u;
    const bool wm = c.dr_wheel_mass_enable != 0;
    EnvState e;
    QuadRaw raw;
    quad_issue_loads(st, n, ii, w, raw, false, wm, false);
    const float2 a = action[ii];
    quad_unpack(raw, w, e, false, wm, c.d_inv_Iw);
    // A. action manager: only the bookkeeping; the map itself comes from the aux warp
    e.prev_action[0] = e.action[0]; e.prev_action[1] = e.action[1];
    e.action[0] = a.x; e.action[1] = a.y;
    const float my_effort = (w == 0) ? c.dc_effort[0] : (w == 1) ? c.dc_effort[1] : (w == 2) ? c.dc_effort[2] : c.dc_effort[3];
    // B. decimation x (actuators -> integrator)
    Chassis b;
    M3 R = rotmat(e.qw, e.qx, e.qy, e.qz);
    V3 cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    b.pc = V3{e.p.x + cw.x, e.p.y + cw.y, e.p.z + cw.z};
    b.v = e.v; b.qw = e.qw; b.qx = e.qx; b.qy = e.qy; b.qz = e.qz;
    b.wb = rotT(R, e.w);
    StepConsts kc = make_step_consts<1>(c, e);
    bar_sync(1);
    const float my_targets[4] = {sh.tgt[q][w], 0.0f, 0.0f, 0.0f};
    const float my_steer_target[2] = {sh.tgt[q][(w == 3) ? 5 : 4], 0.0f};
    for (int d = 0; d < c.decimation; ++d) {
        float lo[4], hi[4];
        dc_limits(c, my_effort, e.omega[0], lo[0], hi[0]);
        for (int jj = 0; jj < c.substeps; ++jj) physics_substep<WL_TASK_DRIFT, 4>(c, T, e, b, my_targets, lo, hi, my_steer_target, kc);
    }
    R = rotmat(b.qw, b.qx, b.qy, b.qz);
    cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    e.p = V3{b.pc.x - cw.x, b.pc.y - cw.y, b.pc.z - cw.z};
    e.v = b.v; e.qw = b.qw; e.qx = b.qx; e.qy = b.qy; e.qz = b.qz;
    e.w = rot(R, b.wb);
    // C./D. counters, terminations (the env warp needs `done` for the reset; the reward side lives in the aux warp)
    e.ep_len += 1;
    const bool time_out = e.ep_len >= c.max_episode_length;
    V3 vb = rotT(R, e.v);
    const float steer_l = __shfl_sync(0xffffffffu, e.steer[0], base + 2), steer_r = __shfl_sync(0xffffffffu, e.steer[0], base + 3);
    const bool oob = drift_off_track(c, e.p.x, e.p.y);
    const uint32_t raw_mask = (time_out ? 1u : 0u) | (oob ? 2u : 0u);
    if (w == 0) {
        float* fo = sh.fin[q];
        fo[0] = e.p.x; fo[1] = e.p.y; fo[2] = e.p.z; fo[3] = vb.x; fo[4] = vb.y; fo[5] = vb.z;
        fo[6] = b.wb.x; fo[7] = b.wb.y; fo[8] = b.wb.z; fo[9] = e.w.z; fo[10] = steer_l; fo[11] = steer_r;
        fo[12] = __uint_as_float(raw_mask);
    }
    bar_arrive(2);
    // I'. observation inputs from the post-physics state (an env that resets / gets pushed redoes them below)
    float eu_k = euler_lane(e.qw, e.qx, e.qy, e.qz, w);
    V3 wbo = rotT(R, e.w);
    const bool done = (raw_mask & (uint32_t)c.term_enable) != 0u;
    // F. auto-reset (sums and log: aux warp)
    if (done) {
        drift_reset_env(c, e, gid, t);
        eu_k = euler_lane(e.qw, e.qx, e.qy, e.qz, w);
        R = rotmat(e.qw, e.qx, e.qy, e.qz); vb = rotT(R, e.v); wbo = rotT(R, e.w);
    }
    // H. interval pushes + noise, drawn by the aux warp
    bar_sync(3);
    if (c.push_enable) {
        if (done) { e.t_hf = e.t_hf - c.d_step_dt; e.t_lf = e.t_lf - c.d_step_dt; }       // fresh timers exceed step_dt: no push on a reset step
        else {
            const float* pu = sh.push[q];
            bool fired = false;
            if (pu[0] != 0.0f) { e.v.x = e.v.x + pu[1]; e.v.y = e.v.y + pu[2]; e.w.z = e.w.z + pu[3]; fired = true; }
            e.t_hf = pu[4];
            if (pu[5] != 0.0f) { e.w.z = e.w.z + pu[6]; fired = true; }
            e.t_lf = pu[7];
            if (fired) { vb = rotT(R, e.v); wbo = rotT(R, e.w); }
        }
    }
    const int k4 = (w < 3) ? 4 * w : 0;
    const float zn[4] = {sh.noise[q][k4], sh.noise[q][k4 + 1], sh.noise[q][k4 + 2], sh.noise[q][k4 + 3]};
    // the CTA's 8 observation rows are 448 contiguous bytes: staged in shared memory and written as 28 x 128-bit stores
    // (full sectors towards HBM, NVLink peers and -- zero-copy host transport -- PCIe, instead of 8-byte pieces)
    float* orow0 = obs + (size_t)WL_OBS_DIM_BLIND * env0;
    if ((FAN == 2) || (env0 + WL_DUO_ENVS <= n && (reinterpret_cast<uintptr_t>(orow0) & 15u) == 0)) {
        const float4 v = blind_obs_quad_values(c, e, w, eu_k, vb, wbo, zn);
        float* so = sh.obs + WL_OBS_DIM_BLIND * q + 4 * w;
        so[0] = v.x; so[1] = v.y;
        if (w < 3) { so[2] = v.z; so[3] = v.w; }
        __syncwarp();
        if (lane < (WL_DUO_ENVS * WL_OBS_DIM_BLIND) / 4)
        {
            if (FAN) fan_store16<FAN == 2>(pf, reinterpret_cast<float4*>(orow0) + lane, reinterpret_cast<const float4*>(sh.obs)[lane]);
            else reinterpret_cast<float4*>(orow0)[lane] = reinterpret_cast<const float4*>(sh.obs)[lane];
        }
    } else {
        if (FAN == 1) blind_obs_quad(c, e, w, eu_k, vb, wbo, zn, obs + (size_t)WL_OBS_DIM_BLIND * ii, live, pf);
        else blind_obs_quad(c, e, w, eu_k, vb, wbo, zn, obs + (size_t)WL_OBS_DIM_BLIND * ii, live);
    }
    if (live) store_env_quad(st, n, i, w, e, false, false);
}

// ---------------------------------------------------------------------------------------
// fused policy: actor + critic 64x64 ELU MLPs in front of the quad step (one warp = 8 envs; lane w of a quad owns hidden
// units [16w, 16w+16) of BOTH nets; layer outputs are exchanged through 4 KB of shared memory per warp)
// ---------------------------------------------------------------------------------------
#define WL_HID 64
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct PolicyOffsets { int32_t o[13]; };

static int policy_offsets(int obs_dim, int32_t off[13]) {
    auto al = [](int x) { return (x + 3) & ~3; };
    int p = 0;
    for (int net = 0; net < 2; ++net) {
        const int out = net == 0 ? 2 : 1;
        off[net * 6 + 0] = p; p = al(p + al(obs_dim) * WL_HID);      // rows padded to a multiple of 4 (zeros)
        off[net * 6 + 1] = p; p = al(p + WL_HID);
        off[net * 6 + 2] = p; p = al(p + WL_HID * WL_HID);
        off[net * 6 + 3] = p; p = al(p + WL_HID);
        off[net * 6 + 4] = p; p = al(p + WL_HID * out);
        off[net * 6 + 5] = p; p = al(p + out);
    }
    off[12] = p; p = al(p + 2);
    return p;
}

__device__ __forceinline__ void mbar_init(uint64_t* b) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(b))); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b) {       // phase 0 (each barrier is used once per launch)
    const uint32_t mb = smem_u32(b);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mb), "r"(0u) : "memory");
}
__device__ __forceinline__ float elu(float x) { return x > 0.0f ? x : __expf(x) - 1.0f; }

// one hidden layer for a register tile of 4 units x 4 envs: per 4 inputs, 4 LDS.128 of activations (one per env, a pure
// broadcast: the whole warp shares the env tile) and 4 LDS.128 of input-major weights (conflict-free: the 16 lanes of a
// net read 256 contiguous bytes) feed 64 FMAs; accumulation order per output = bias, then inputs ascending
__device__ __forceinline__ void mlp_tile_layer(const float* W, const float* bias, const float* x0, int xstride, int nin, float acc[4][4]) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias);
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[e][0] = b0.x; acc[e][1] = b0.y; acc[e][2] = b0.z; acc[e][3] = b0.w; }
#pragma unroll 4
    for (int j = 0; j < nin; j += 4) {
        float xs[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 v = *reinterpret_cast<const float4*>(x0 + e * xstride + j);
            xs[e][0] = v.x; xs[e][1] = v.y; xs[e][2] = v.z; xs[e][3] = v.w;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float4 w0 = *reinterpret_cast<const float4*>(W + (j + jj) * WL_HID);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x = xs[e][jj];
                acc[e][0] = fmaf(w0.x, x, acc[e][0]); acc[e][1] = fmaf(w0.y, x, acc[e][1]);
                acc[e][2] = fmaf(w0.z, x, acc[e][2]); acc[e][3] = fmaf(w0.w, x, acc[e][3]);
            }
        }
    }
}

__device__ __forceinline__ void mlp_tile_store_elu(float* h0, int hstride, const float acc[4][4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
        *reinterpret_cast<float4*>(h0 + e * hstride) = make_float4(elu(acc[e][0]), elu(acc[e][1]), elu(acc[e][2]), elu(acc[e][3]));
}

#define WL_ACT_ENVS 32                           // envs per CTA: 4096 envs -> 128 CTAs, one per SM
#define WL_ACT_THREADS 256                       // 8 lanes per env for the MLPs (2 warps per scheduler); threads 0..127 = the env quads
#define WL_ACT_XS 20                             // padded observation row
#define WL_ACT_HS (2 * WL_HID + 4)               // padded hidden row, [actor 64 | critic 64 | pad]

template <int TASK>
__global__ void __launch_bounds__(WL_ACT_THREADS)
wl_act_step_quad_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
                        const float* __restrict__ obs_in, const float* __restrict__ blob, int blob_floats, PolicyOffsets po_off,
                        wl_policy_out po, float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ terminated_o,
                        uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t_arg, int obs_dim) {
    constexpr bool ELEV = (TASK == WL_TASK_ELEVATION), VIS = (TASK == WL_TASK_VISUAL);
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == gridDim.x - 1) {                       // the janitor CTA (appended to the grid, owns no envs)
        if (threadIdx.x < 32) {
            const uint32_t tj = decode_step(gl, t_arg);
            float wj[WL_MAX_REW_TERMS];
            load_weights(c, gl, tj, wj); janitor(c, gl, tj, wj, d_log);
        }
        return;
    }
    float* sw = smem;                                        // the whole weight blob (~42 KB)
    float* sx = smem + blob_floats;                          // [32][WL_ACT_XS]
    float* sh = sx + WL_ACT_ENVS * WL_ACT_XS;                // [32][WL_ACT_HS]
    // ---- bulk async copies (TMA, 1-D) bring the blob from L2 into shared memory while the env state loads fly: the two
    // first-layer blocks (7.5 KB) complete on mbar[0], everything else (34 KB) on mbar[1] and lands during layer 1
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0]); mbar_init(&mbar[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int a0 = po_off.o[0], a1 = po_off.o[2], c0 = po_off.o[6], c1 = po_off.o[8];
        mbar_expect(&mbar[0], (uint32_t)((a1 - a0) + (c1 - c0)) * 4u);
        bulk_g2s(sw + a0, blob + a0, (uint32_t)(a1 - a0) * 4u, &mbar[0]);
        bulk_g2s(sw + c0, blob + c0, (uint32_t)(c1 - c0) * 4u, &mbar[0]);
        mbar_expect(&mbar[1], (uint32_t)((c0 - a1) + (blob_floats - c1)) * 4u);
        bulk_g2s(sw + a1, blob + a1, (uint32_t)(c0 - a1) * 4u, &mbar[1]);
        bulk_g2s(sw + c1, blob + c1, (uint32_t)(blob_floats - c1) * 4u, &mbar[1]);
    }
    const uint32_t t = decode_step(gl, t_arg);
    const VisualMap vm = VIS ? visual_map(c, T.hf) : VisualMap{nullptr, nullptr};
    const int n = c.num_envs;
    const bool stepper = threadIdx.x < 4 * WL_ACT_ENVS;      // warps 0..3 own the env quads; warps 4..7 only help with the MLPs
    const int q = (threadIdx.x >> 2) & (WL_ACT_ENVS - 1), w = threadIdx.x & 3;
    const int i = blockIdx.x * WL_ACT_ENVS + q;
    const bool live = i < n;
    const int ii = live ? i : n - 1;
    const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
    const unsigned base = (threadIdx.x & 31u) & ~3u;
    EnvState e;
    float wts[WL_MAX_REW_TERMS];
    if (stepper) {
        load_env_quad(st, n, ii, w, e, ELEV, c.dr_wheel_mass_enable != 0, c.d_inv_Iw);
        load_weights(c, gl, t, wts);
    }
    for (int k = threadIdx.x; k < WL_ACT_ENVS * 16; k += WL_ACT_THREADS) {
        const int eq = k >> 4, j = k & 15, ei = min(blockIdx.x * WL_ACT_ENVS + eq, n - 1);
        sx[eq * WL_ACT_XS + j] = j < obs_dim ? obs_in[(size_t)ei * obs_dim + j] : 0.0f;
    }
    // the policy noise does not depend on the weights: draw it while the copies are in flight
    float z0 = 0.0f, z1 = 0.0f, zn[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (stepper) {
        const uint4 r = philox4x32(c.seed, gid, t, RNG_POLICY, 0u);
        box_muller(r.x, r.y, z0, z1);
        if (!ELEV && !VIS) quad_obs_noise(c, w, gid, t, RNG_OBS, 0u, zn);
    }
    __syncthreads();                                         // mbarrier init visible to every waiter; obs rows written
    mbar_wait(&mbar[0]);
    // ---- MLPs: warp k takes envs 4k..4k+3 of the CTA; its 32 lanes tile the 128 (actor | critic) hidden columns 4 at a
    // time, so a lane holds 4 units x 4 envs in registers.  Layer exchanges stay warp-local (__syncwarp).  W1t has obs_dim
    // rows; rows obs_dim..nin1-1 read whatever follows in the blob and meet the zero padding of x.
    {
        const int lane = threadIdx.x & 31, net = lane >> 4, ucol = (lane & 15) * 4;
        const int env0 = (threadIdx.x >> 5) * 4;
        const int nin1 = (obs_dim + 3) & ~3;
        const int oW1 = net ? po_off.o[6] : po_off.o[0], oB1 = net ? po_off.o[7] : po_off.o[1];     // (no dynamic indexing of kernel params)
        const int oW2 = net ? po_off.o[8] : po_off.o[2], oB2 = net ? po_off.o[9] : po_off.o[3];
        float acc[4][4];
        float* h0 = sh + env0 * WL_ACT_HS + net * WL_HID;
        mlp_tile_layer(sw + oW1 + ucol, sw + oB1 + ucol, sx + env0 * WL_ACT_XS, WL_ACT_XS, nin1, acc);
        mlp_tile_store_elu(h0 + ucol, WL_ACT_HS, acc);
        __syncwarp();
        mbar_wait(&mbar[1]);                                 // layer-2 / head weights have landed (copy overlapped layer 1)
        mlp_tile_layer(sw + oW2 + ucol, sw + oB2 + ucol, h0, WL_ACT_HS, WL_HID, acc);
        __syncwarp();                                        // every lane has finished reading layer-1 activations
        mlp_tile_store_elu(h0 + ucol, WL_ACT_HS, acc);
    }
    __syncthreads();                                         // layer-2 activations of all 32 envs are in shared memory
    if (stepper) {
        // output heads: lane w of the env's quad takes hidden units [16w, 16w+16) of both nets, then a quad butterfly
        const float* hrow = sh + q * WL_ACT_HS;
        const float* w3a = sw + po_off.o[4]; const float* w3c = sw + po_off.o[10];
        float m0 = 0.0f, m1 = 0.0f, vv = 0.0f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int j = 16 * w + u;
            const float xa = hrow[j], xc = hrow[WL_HID + j];
            m0 = fmaf(w3a[2 * j], xa, m0); m1 = fmaf(w3a[2 * j + 1], xa, m1);
            vv = fmaf(w3c[j], xc, vv);
        }
        m0 += __shfl_xor_sync(0xffffffffu, m0, 1); m0 += __shfl_xor_sync(0xffffffffu, m0, 2);
        m1 += __shfl_xor_sync(0xffffffffu, m1, 1); m1 += __shfl_xor_sync(0xffffffffu, m1, 2);
        vv += __shfl_xor_sync(0xffffffffu, vv, 1); vv += __shfl_xor_sync(0xffffffffu, vv, 2);
        m0 += sw[po_off.o[5]]; m1 += sw[po_off.o[5] + 1]; vv += sw[po_off.o[11]];
        // Gaussian head: a = mean + std * z (rsl_rl ActorCritic.act), log-prob summed over the action dims
        const float s0 = sw[po_off.o[12]], s1 = sw[po_off.o[12] + 1];
        const float2 a = make_float2(fmaf(s0, z0, m0), fmaf(s1, z1, m1));
        if (live && w == 0) {
            reinterpret_cast<float2*>(po.actions)[i] = a;
            reinterpret_cast<float2*>(po.mean)[i] = make_float2(m0, m1);
            po.log_prob[i] = -0.5f * (z0 * z0 + z1 * z1) - __logf(s0) - __logf(s1) - 1.8378770664093453f;   // 2 * 0.5 log(2 pi)
            po.value[i] = vv;
        }
        // ---- the env step on the sampled action
        const int od = ELEV ? WL_OBS_DIM_ELEV : VIS ? WL_OBS_DIM_VISUAL + vis_cam_floats(c) : WL_OBS_DIM_BLIND;
        quad_env_step<TASK>(c, T, vm, gl->acc[t % 3u], wts, e, i, w, live, gid, base, t, a, zn, obs + (size_t)od * ii, rew, terminated_o, truncated_o);
        if (live) store_env_quad(st, n, i, w, e, ELEV);
    }
}

// K consecutive env.step()s in ONE launch (synthetic / scripted-action rollouts): the state stays in registers
// for the whole rollout, every step still writes its observation / reward / done rows ([K,N,...] slab) and its episode-log
// row.  actions == nullptr => U[-1,1]^2 drawn in-kernel from the counter-based generator (and stored to act_out).
// Bit-identical to K calls of wl_step.  The curriculum is applied by the host-side splitter: [t, t+K) never crosses an
// episode boundary of the global counter when curr_n > 0.
template <int TASK>
__global__ void __launch_bounds__(128)
wl_rollout_quad_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
                       const float2* __restrict__ action, float2* __restrict__ act_out, float* __restrict__ obs,
                       float* __restrict__ rew, uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o,
                       float* __restrict__ d_log, uint32_t t_arg, int K, unsigned* __restrict__ ticket) {
    constexpr bool ELEV = (TASK == WL_TASK_ELEVATION), VIS = (TASK == WL_TASK_VISUAL);
    const uint32_t t0 = decode_step(gl, t_arg);
    const VisualMap vm = VIS ? visual_map(c, T.hf) : VisualMap{nullptr, nullptr};
    const int n = c.num_envs;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 2, w = tid & 3;
    const bool live = i < n;
    const int ii = live ? i : n - 1;
    const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
    const unsigned base = (threadIdx.x & 31u) & ~3u;
    EnvState e;
    load_env_quad(st, n, ii, w, e, ELEV, c.dr_wheel_mass_enable != 0, c.d_inv_Iw);
    float wts[WL_MAX_REW_TERMS];             // W(t0) holds for the whole window (it never crosses an episode boundary)
    load_weights(c, gl, t0, wts);
    if (blockIdx.x == 0 && threadIdx.x < 32) janitor(c, gl, t0, wts, nullptr);      // flushes step t0-1's log row
    const int od = ELEV ? WL_OBS_DIM_ELEV : VIS ? WL_OBS_DIM_VISUAL + vis_cam_floats(c) : WL_OBS_DIM_BLIND;
    for (int k = 0; k < K; ++k) {
        const uint32_t t = t0 + (uint32_t)k;
        float2 a;
        if (action != nullptr) a = action[(size_t)k * n + ii];
        else {
            uint4 r = philox4x32(c.seed, gid, t, RNG_ACTION, 0u);
            a = make_float2(2.0f * u01(r.x) - 1.0f, 2.0f * u01(r.y) - 1.0f);
        }
        if (act_out != nullptr && live && w == 0) act_out[(size_t)k * n + i] = a;
        float zn[4];
        if (!ELEV && !VIS) quad_obs_noise(c, w, gid, t, RNG_OBS, 0u, zn); else zn[0] = zn[1] = zn[2] = zn[3] = 0.0f;
        quad_env_step<TASK>(c, T, vm, d_log + (size_t)k * WL_LOG_FLOATS, wts, e, i, w, live, gid, base, t, a, zn,
                            obs + ((size_t)k * n + ii) * od, rew + (size_t)k * n, terminated_o + (size_t)k * n,
                            truncated_o + (size_t)k * n);
    }
    if (live) store_env_quad(st, n, i, w, e, ELEV);
    // last CTA (ticket, amortised over the K steps of the launch): turn the K accumulator rows into means and leave the
    // globals as K single steps would have: W of the last step in its slot, its reset count where the next launch looks for
    // it (the curriculum boundary at t0 + K is evaluated there), nothing pending for the log, the next row re-armed
    if ((threadIdx.x & 31) == 0) __threadfence();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned tk = atomicAdd(ticket, 1u);
    if (tk != gridDim.x - 1) return;
    __threadfence();
    float last_cnt = 0.0f;
    for (int k = 0; k < K; ++k) {
        float* row = d_log + (size_t)k * WL_LOG_FLOATS;
        const float cnt = __ldcg(&row[8]);
        if (cnt > 0.0f) {
            const float denom = cnt * c.episode_length_s;
            for (int q = 0; q < WL_MAX_REW_TERMS; ++q) row[q] = __ldcg(&row[q]) / denom;
            for (int q = 0; q < 16; ++q) gl->last_log[q] = row[q];
        } else {
            for (int q = 0; q < 16; ++q) row[q] = gl->last_log[q];         // no reset in this step: the previous row persists
        }
        last_cnt = cnt;
    }
    const uint32_t tl = t0 + (uint32_t)K - 1u;
    for (int q = 0; q < WL_MAX_REW_TERMS; ++q) gl->rew_weight[tl & 1u][q] = wts[q];
    for (int q = 0; q < 16; ++q) { gl->acc[tl % 3u][q] = (q == 8) ? last_cnt : 0.0f; gl->acc[(tl + 1u) % 3u][q] = 0.0f; }
    gl->log_ptr[tl % 3u] = nullptr;
    *ticket = 0u;
}

// ---------------------------------------------------------------------------------------
// height-scan observation (ray-caster sensor, elevation/mushr_elevation_env_cfg.py:44-48,74-82,132-142):
// one CTA per env.  The yaw-rotated 2.5 m x 2.5 m footprint fits a WL_TILE x WL_TILE window of the height-field;
// one elected thread pulls that window into shared memory with a single TMA (cp.async.bulk.tensor.2d, completion on
// an mbarrier; out-of-raster elements are zero-filled by the hardware and never read), then 128 threads take the
// 676 vertical rays as bilinear samples from shared memory and write obs[13:689] coalesced.
// ---------------------------------------------------------------------------------------
// window = WL_TILE_W x WL_TILE_H samples.  The footprint needs <= 38 samples per axis; TMA wants the innermost box
// coordinate 16-byte aligned (ox % 4 == 0), so the window starts up to 3 samples early and is 44 wide.
#define WL_TILE_W 44
#define WL_TILE_H 40
#define WL_SCAN_THREADS 128


// The 676 rays of one env out of its shared-memory window: thread t takes rays t, t+128, ... (linear order: all 32 lanes busy,
// fully coalesced stores).  The SIX rays of a thread are independent chains, written branch-free (window indices clamped into
// the tile, the miss value selected at the end) and fully unrolled so that the compiler interleaves them: the loop is bound by
// instruction latency (F2I, LDS, dependent FFMAs), not by bytes.  Per-ray arithmetic unchanged (bit-identical outputs).
struct ScanPose { float bx, by, bz, cy, sy, pz; int ox, oy; };
__device__ __forceinline__ void scan_rays(const wl_config& c, const float* __restrict__ tl, const ScanPose& q, float* __restrict__ row, int tid) {
    constexpr int ITER = (WL_SCAN_RAYS + WL_SCAN_THREADS - 1) / WL_SCAN_THREADS;
    const float inv = c.d_inv_hf_cell;
    const float fxmax = (float)(c.hf_nx - 1), fymax = (float)(c.hf_ny - 1);
    const float base = q.pz - c.scan_plane_init;
    float v[ITER], tx[ITER], ty[ITER];
    const float* t0[ITER];
    bool inside[ITER];
#pragma unroll
    for (int j = 0; j < ITER; ++j) {                                              // phase 1: ray -> window address, weights
        int r = tid + WL_SCAN_THREADS * j;
        if (r > WL_SCAN_RAYS - 1) r = WL_SCAN_RAYS - 1;                           // the tail threads recompute the last ray (not stored)
        const int ry = (r * 2521) >> 16, rx = r - ry * WL_SCAN_SIDE;             // r / 26, r % 26 for r < 676
        const float lx = fm((float)rx, c.scan_res, -c.scan_half), ly = fm((float)ry, c.scan_res, -c.scan_half);
        const float wx = fm(q.cy, lx, fm(-q.sy, ly, q.bx)), wy = fm(q.sy, lx, fm(q.cy, ly, q.by));
        const float fx = (wx - c.hf_x0) * inv, fy = (wy - c.hf_y0) * inv;
        inside[j] = (fx >= 0.0f) && (fy >= 0.0f) && (fx <= fxmax) && (fy <= fymax);
        int ix = (int)floorf(fx), iy = (int)floorf(fy);
        ix = min(ix, c.hf_nx - 2); iy = min(iy, c.hf_ny - 2);
        tx[j] = fx - (float)ix; ty[j] = fy - (float)iy;
        // (the six F2I/I2F per ray run on the 16-lane pipe, 48 of its cycles per warp-iteration against 64 issue slots: the loop
        //  is issue-bound, and the 2^23-binade tricks that avoid the conversions cost 16 more issue slots per ray -- not taken)
        const int ux = min(max(ix - q.ox, 0), WL_TILE_W - 2), uy = min(max(iy - q.oy, 0), WL_TILE_H - 2);   // no-op for a hit
        t0[j] = tl + uy * WL_TILE_W + ux;
    }
    float z00[ITER], z10[ITER], z01[ITER], z11[ITER];
    asm volatile("" ::: "memory"); __syncwarp();                                  // (compiler fences: keep the phases apart)
#pragma unroll
    for (int j = 0; j < ITER; ++j) {                                              // phase 2: the 24 shared-memory reads in flight together
        z00[j] = t0[j][0]; z10[j] = t0[j][1]; z01[j] = t0[j][WL_TILE_W]; z11[j] = t0[j][WL_TILE_W + 1];
    }
    asm volatile("" ::: "memory"); __syncwarp();
#pragma unroll
    for (int j = 0; j < ITER; ++j) {                                              // phase 3: bilinear sample, height, clip
        const float za = fm(z10[j] - z00[j], tx[j], z00[j]), zb = fm(z11[j] - z01[j], tx[j], z01[j]);
        const float hit = fm(zb - za, ty[j], za);
        const float hs = q.bz - hit - c.scan_offset;                              // mdp.height_scan
        const float hv = r_clamp(-hs + base, -c.obs_clip, c.obs_clip);           // world_height_map, clip
        v[j] = inside[j] ? hv : c.obs_clip;                                       // miss: +inf clipped
    }
#pragma unroll
    for (int j = 0; j < ITER; ++j) {
        const int r = tid + WL_SCAN_THREADS * j;
        if (r < WL_SCAN_RAYS) row[r] = v[j];
    }
}

template <bool USE_TMA>
__global__ void __launch_bounds__(WL_SCAN_THREADS, 8)
wl_scan_kernel(const __grid_constant__ wl_config c, const __grid_constant__ CUtensorMap tmap, const float4* __restrict__ st,
               const float* __restrict__ hf, float* __restrict__ obs) {
    __shared__ __align__(128) float tile[WL_TILE_W * WL_TILE_H];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ float sp[8];                  // per-env constants of the scan, formed ONCE by thread 0: bx by bz cy sy pz ox oy
    const int n = c.num_envs;
    const int i = blockIdx.x;
    const float inv = c.d_inv_hf_cell;
    if (threadIdx.x == 0) {
        float4 g0 = ldg4(st, WL_G_POS, n, i), g1 = ldg4(st, WL_G_QUAT, n, i);
        EnvState e0; e0.p = V3{g0.x, g0.y, g0.z}; e0.qw = g1.x; e0.qx = g1.y; e0.qy = g1.z; e0.qz = g1.w;
        M3 R = rotmat(e0.qw, e0.qx, e0.qy, e0.qz);
        // ray-caster parent = base_link (root + R (0,0,base_link_z)); rays are yaw-aligned (attach_yaw_only)
        const float bx0 = fm(c.base_link_z, R.r[2], e0.p.x), by0 = fm(c.base_link_z, R.r[5], e0.p.y), bz0 = fm(c.base_link_z, R.r[8], e0.p.z);
        float cy0, sy0; yaw_cs(e0, cy0, sy0);
        // window origin (in samples): covers base +- (half*sqrt2 + 1 cell)
        const float reach = c.scan_half * 1.41421356237f + c.hf_cell;
        const int ox0 = ((int)floorf((bx0 - reach - c.hf_x0) * inv)) & ~3;      // two's complement: rounds toward -inf
        const int oy0 = (int)floorf((by0 - reach - c.hf_y0) * inv);
        sp[0] = bx0; sp[1] = by0; sp[2] = bz0; sp[3] = cy0; sp[4] = sy0; sp[5] = e0.p.z; sp[6] = __int_as_float(ox0); sp[7] = __int_as_float(oy0);
        if (USE_TMA) {
            const uint32_t mb = smem_u32(&mbar), dst = smem_u32(tile);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(WL_TILE_W * WL_TILE_H * 4)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tmap), "r"(ox0), "r"(oy0), "r"(mb) : "memory");
        }
    }
    __syncthreads();                         // constants + mbarrier visible to everyone
    const float bx = sp[0], by = sp[1], bz = sp[2], cy = sp[3], sy = sp[4], pz = sp[5];
    const int ox = __float_as_int(sp[6]), oy = __float_as_int(sp[7]);
    if (USE_TMA) {
        const uint32_t mb = smem_u32(&mbar);     // all threads wait for the bytes (phase 0)
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(mb), "r"(0u) : "memory");
        }
    } else {
        for (int k = threadIdx.x; k < WL_TILE_W * WL_TILE_H; k += WL_SCAN_THREADS) {
            int tx = k % WL_TILE_W, ty = k / WL_TILE_W, gx = ox + tx, gy = oy + ty;
            tile[k] = (gx >= 0 && gy >= 0 && gx < c.hf_nx && gy < c.hf_ny) ? __ldg(hf + (size_t)gy * c.hf_pitch + gx) : 0.0f;
        }
        __syncthreads();
    }
    float* row = obs + (size_t)WL_OBS_DIM_ELEV * i + 13;
    ScanPose q{bx, by, bz, cy, sy, pz, ox, oy};
    scan_rays(c, tile, q, row, threadIdx.x);
}

// ---------------------------------------------------------------------------------------
// The same ray-caster as a TMA producer / consumer PIPELINE (wl_set_scan_tma(2); measured 13.2 us vs 12.4 us for one tile per
// CTA at 4096 envs, 150 vs 135 us at 65536 -- 16 resident one-tile CTAs per SM already hide the set-up + TMA latency, and the
// ray loop is issue-bound -- so it is NOT the default; profiles/r02_scan_ab.jsonl): persistent CTAs (a few per SM) walk the envs; warp 4 is
// the producer -- it reads the next env's pose, forms the per-env constants and issues the cp.async.bulk.tensor.2d of that
// env's window into the other shared-memory stage -- while warps 0..3 take the 676 rays of the current env out of the stage
// that has landed.  full[s] (TMA transaction bytes) and empty[s] (one arrival per consumer warp) are mbarriers; the pose load,
// the constant set-up and the TMA latency of env k+1 hide behind the rays of env k instead of heading every CTA's life.
// Rays are taken in linear order (r = thread + 128 j: all 32 lanes busy, fully coalesced stores).  Same arithmetic.
// ---------------------------------------------------------------------------------------
#define WL_SCAN_STAGES 2
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait_parity(uint64_t* b, uint32_t parity) {
    const uint32_t mb = smem_u32(b);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(WL_SCAN_THREADS + 32)
wl_scan_pipe_kernel(const __grid_constant__ wl_config c, const __grid_constant__ CUtensorMap tmap, const float4* __restrict__ st,
                    float* __restrict__ obs) {
    __shared__ __align__(128) float tile[WL_SCAN_STAGES][WL_TILE_W * WL_TILE_H];
    __shared__ __align__(8) uint64_t full[WL_SCAN_STAGES], empty[WL_SCAN_STAGES];
    __shared__ float sp[WL_SCAN_STAGES][8];                  // bx by bz cy sy pz ox oy of the env in the stage
    const int n = c.num_envs;
    const float inv = c.d_inv_hf_cell;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < WL_SCAN_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(WL_SCAN_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == WL_SCAN_THREADS / 32) {
        // ---------------- producer ----------------
        if (lane != 0) return;
        int k = 0;
        for (int i = blockIdx.x; i < n; i += gridDim.x, ++k) {
            const int s = k % WL_SCAN_STAGES;
            const uint32_t ph = (uint32_t)(k / WL_SCAN_STAGES) & 1u;
            mbar_wait_parity(&empty[s], ph ^ 1u);                        // stage free (passes at once on its first use)
            float4 g0 = ldg4(st, WL_G_POS, n, i), g1 = ldg4(st, WL_G_QUAT, n, i);
            EnvState e0; e0.p = V3{g0.x, g0.y, g0.z}; e0.qw = g1.x; e0.qx = g1.y; e0.qy = g1.z; e0.qz = g1.w;
            M3 R = rotmat(e0.qw, e0.qx, e0.qy, e0.qz);
            const float bx0 = fm(c.base_link_z, R.r[2], e0.p.x), by0 = fm(c.base_link_z, R.r[5], e0.p.y), bz0 = fm(c.base_link_z, R.r[8], e0.p.z);
            float cy0, sy0; yaw_cs(e0, cy0, sy0);
            const float reach = c.scan_half * 1.41421356237f + c.hf_cell;
            const int ox0 = ((int)floorf((bx0 - reach - c.hf_x0) * inv)) & ~3;
            const int oy0 = (int)floorf((by0 - reach - c.hf_y0) * inv);
            float* p = sp[s];
            p[0] = bx0; p[1] = by0; p[2] = bz0; p[3] = cy0; p[4] = sy0; p[5] = e0.p.z; p[6] = __int_as_float(ox0); p[7] = __int_as_float(oy0);
            const uint32_t mb = smem_u32(&full[s]), dst = smem_u32(tile[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(WL_TILE_W * WL_TILE_H * 4)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tmap), "r"(ox0), "r"(oy0), "r"(mb) : "memory");
        }
        return;
    }
    // ---------------- consumers (warps 0..3) ----------------
    int k = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x, ++k) {
        const int s = k % WL_SCAN_STAGES;
        const uint32_t ph = (uint32_t)(k / WL_SCAN_STAGES) & 1u;
        mbar_wait_parity(&full[s], ph);
        const float bx = sp[s][0], by = sp[s][1], bz = sp[s][2], cy = sp[s][3], sy = sp[s][4], pz = sp[s][5];
        const int ox = __float_as_int(sp[s][6]), oy = __float_as_int(sp[s][7]);
        const float* tl = tile[s];
        float* row = obs + (size_t)WL_OBS_DIM_ELEV * i + 13;
        ScanPose q{bx, by, bz, cy, sy, pz, ox, oy};
        scan_rays(c, tl, q, row, threadIdx.x);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);                           // this warp is done with the stage
    }
}

__global__ void wl_startup_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t gid = (uint32_t)(c.env_id_offset + i);
    uint4 r0 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 0u);
    uint4 r1 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 1u);
    uint4 r2 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 2u);
    uint4 r3 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 3u);
    const uint32_t rb[4] = {r0.x, r0.y, r0.z, r0.w}, rk[4] = {r1.x, r1.y, r1.z, r1.w}, rm[4] = {r3.x, r3.y, r3.z, r3.w};
    float D[4], C[4], kd[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t bk = 0;
        if (c.dr_enable && c.dr_num_buckets > 1) bk = __umulhi(rb[w], (uint32_t)c.dr_num_buckets);
        D[w] = c.dr_bucket_D[bk]; C[w] = c.dr_bucket_C[bk];
        kd[w] = c.dc_damping[w];
        if (c.dr_enable && ((c.dr_kd_mask >> w) & 1)) kd[w] = uniform(rk[w], c.dr_kd_range[0], c.dr_kd_range[1]);
    }
    // randomize_rigid_body_mass: "add" -> base_link mass += U; "abs" -> base_link mass := U (mushr_visual_env_cfg.py:280-288);
    // wheel links := U (:290-299), their spin inertia rescaled by the mass ratio
    float mass = c.mass_nominal;
    float inv_Iw[4] = {c.d_inv_Iw, c.d_inv_Iw, c.d_inv_Iw, c.d_inv_Iw};
    if (c.dr_enable) {
        const float u = uniform(r2.x, c.dr_mass_add[0], c.dr_mass_add[1]);
        if (c.dr_mass_mode == 0) mass = mass + u;
        else mass = (mass - c.dr_base_mass_nominal) + u;
        if (c.dr_wheel_mass_enable) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float mw = uniform(rm[w], c.dr_wheel_mass[0], c.dr_wheel_mass[1]);
                mass = mass + (mw - c.wheel_mass_nominal);
                inv_Iw[w] = c.d_inv_Iw * (c.wheel_mass_nominal / mw);
            }
        }
    }
    float inv_mass = 1.0f / mass;
    stg4(st, WL_G_PIW, n, i, make_float4(inv_Iw[0], inv_Iw[1], inv_Iw[2], inv_Iw[3]));
    stg4(st, WL_G_PMASS, n, i, make_float4(mass, inv_mass, 0.0f, 0.0f));
    stg4(st, WL_G_PMU_D, n, i, make_float4(D[0], D[1], D[2], D[3]));
    stg4(st, WL_G_PMU_C, n, i, make_float4(C[0], C[1], C[2], C[3]));
    stg4(st, WL_G_PKD, n, i, make_float4(kd[0], kd[1], kd[2], kd[3]));
    float t_hf = uniform(r2.y, c.push_hf_interval[0], c.push_hf_interval[1]);
    float t_lf = uniform(r2.z, c.push_lf_interval[0], c.push_lf_interval[1]);
    float4 g;
    g = ldg4(st, WL_G_LINVEL, n, i); g.w = t_hf; stg4(st, WL_G_LINVEL, n, i, g);
    g = ldg4(st, WL_G_ANGVEL, n, i); g.w = t_lf; stg4(st, WL_G_ANGVEL, n, i, g);
    stg4(st, WL_G_QUAT, n, i, make_float4(1.0f, 0.0f, 0.0f, 0.0f));
}

__global__ void wl_reset_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, const float* __restrict__ aux,
                                const int64_t* __restrict__ ids, int n_ids, uint32_t t) {
    const int n = c.num_envs;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ids) return;
    const int i = ids ? (int)ids[k] : k;
    if (i < 0 || i >= n) return;
    EnvState e;
    const bool elev = c.task == WL_TASK_ELEVATION;
    load_env(st, n, i, e, elev);
    if (elev) elev_reset_env(c, e, (uint32_t)(c.env_id_offset + i), t);      // command b-frame vector is NOT refreshed by reset()
    else if (c.task == WL_TASK_VISUAL) visual_reset_env(c, visual_map(c, aux), e, (uint32_t)(c.env_id_offset + i), t);
    else drift_reset_env(c, e, (uint32_t)(c.env_id_offset + i), t);
    store_env(st, n, i, e, elev);
}

__global__ void wl_observe_kernel(const __grid_constant__ wl_config c, const float4* __restrict__ st, float* __restrict__ obs,
                                  uint32_t t, uint32_t call_idx) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState e;
    if (c.task == WL_TASK_VISUAL) {
        load_env(st, n, i, e, false);
        float o[8]; visual_proprio(c, e, o);
        const int camf = vis_cam_floats(c);
        float* row = obs + (size_t)(WL_OBS_DIM_VISUAL + camf) * i + camf;
#pragma unroll
        for (int k = 0; k < 8; ++k) row[k] = o[k];
        return;
    }
    if (c.task == WL_TASK_ELEVATION) {
        load_env(st, n, i, e, true);
        float o[13]; elev_proprio(c, e, euler_xyz(e.qw, e.qx, e.qy, e.qz), o);
        float* row = obs + (size_t)WL_OBS_DIM_ELEV * i;
#pragma unroll
        for (int k = 0; k < 13; ++k) row[k] = o[k];
        return;
    }
    load_env(st, n, i, e, false);
    blind_obs(c, e, (uint32_t)(c.env_id_offset + i), t, RNG_OBS_EXTRA, 3u * call_idx, obs + (size_t)WL_OBS_DIM_BLIND * i);
}

// suspension joint pos / vel are not part of the stored state: derive them for the articulation view.
// pos = clamp(compression, +-travel) of the wheel's contact spring, vel = compression rate, wheel order [bl,br,fl,fr].
template <int TASK>
__global__ void wl_suspension_kernel(const __grid_constant__ wl_config c, const float4* __restrict__ st, Terrain T,
                                     float4* __restrict__ pos_o, float4* __restrict__ vel_o) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState e;
    load_env(st, n, i, e, false);
    M3 R = rotmat(e.qw, e.qx, e.qy, e.qz);
    V3 cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    V3 pc{e.p.x + cw.x, e.p.y + cw.y, e.p.z + cw.z};
    V3 vb = rotT(R, e.v), wb = rotT(R, e.w);
    float sp[4], sv[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        V3 rho{((w >= 2) ? c.hub_x_front : c.hub_x_rear) - c.com[0], ((w & 1) ? -c.hub_y : c.hub_y) - c.com[1], c.hub_z - c.com[2]};
        V3 hubw = rot(R, rho);
        hubw.x += pc.x; hubw.y += pc.y; hubw.z += pc.z;
        float zt = 0.0f; V3 nw{0.0f, 0.0f, 1.0f};
        if (TASK == WL_TASK_ELEVATION) heightfield_at(c, T, hubw.x, hubw.y, zt, nw);
        float comp = fm(-(hubw.z - zt), nw.z, c.wheel_radius);
        V3 nb = rotT(R, nw);
        V3 rc = axpy(rho, -c.wheel_radius, nb);
        V3 vc = cross(wb, rc);
        vc.x += vb.x; vc.y += vb.y; vc.z += vb.z;
        sp[w] = r_clamp(comp, -c.susp_travel, c.susp_travel);
        sv[w] = (comp > 0.0f) ? -dot(nb, vc) : 0.0f;
    }
    pos_o[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
    vel_o[i] = make_float4(sv[0], sv[1], sv[2], sv[3]);
}

struct CurrArgs { int32_t n; int32_t slots[WL_MAX_REW_TERMS]; float inc[WL_MAX_REW_TERMS]; uint32_t fire_mask; uint32_t wslot, row; };
__global__ void wl_curriculum_kernel(wl_globals* __restrict__ gl, CurrArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!(gl->acc[a.row][8] > 0.0f)) return;                 // the reference's call site is _reset_idx: only if >= 1 env reset
    for (int t = 0; t < a.n; ++t)
        if ((a.fire_mask >> t) & 1u) gl->rew_weight[a.wslot][a.slots[t]] += a.inc[t];
}
// publish the log row of step t_next - 1 without stepping (idempotent; the next step's janitor would do the same)
__global__ void wl_flush_kernel(const __grid_constant__ wl_config c, wl_globals* __restrict__ gl, uint32_t t_next) {
    if (t_next != 0u) publish_log_row(c, gl, (t_next + 2u) % 3u, threadIdx.x & 31);
}
// counter jump: carry the weights over to the slot the next step reads, clear the accumulators and the pending log
__global__ void wl_rearm_kernel(wl_globals* __restrict__ gl, uint32_t from_slot, uint32_t to_slot) {
    const int lane = threadIdx.x;
    if (lane < WL_MAX_REW_TERMS && from_slot != to_slot) gl->rew_weight[to_slot][lane] = gl->rew_weight[from_slot][lane];
    if (lane < 16) { gl->acc[0][lane] = 0.0f; gl->acc[1][lane] = 0.0f; gl->acc[2][lane] = 0.0f; }
    if (lane < 3) gl->log_ptr[lane] = nullptr;
}
// host-counter path, after a step whose successor counter t_next is a curriculum boundary: apply the terms in place to the
// slot step t_next will read (so the host sees them at once) and mark the boundary as done
__global__ void wl_boundary_kernel(const __grid_constant__ wl_config c, wl_globals* __restrict__ gl, uint32_t t_next) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float wts[WL_MAX_REW_TERMS];
    load_weights(c, gl, t_next, wts);
    for (int q = 0; q < WL_MAX_REW_TERMS; ++q) gl->rew_weight[(t_next + 1u) & 1u][q] = wts[q];
    gl->curr_applied_t = t_next;
}
__global__ void wl_advance_kernel(wl_globals* __restrict__ gl, uint32_t k) { if (threadIdx.x == 0 && blockIdx.x == 0) gl->step_base += k; }

__global__ void wl_synth_actions_kernel(const __grid_constant__ wl_config c, const wl_globals* __restrict__ gl,
                                        float2* __restrict__ action, uint32_t t_arg, int dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.num_envs) return;
    const uint32_t t = decode_step(gl, t_arg);
    uint4 r = philox4x32(c.seed, (uint32_t)(c.env_id_offset + i), t, RNG_ACTION, 0u);
    float a0, a1;
    if (dist == 0) { a0 = 2.0f * u01(r.x) - 1.0f; a1 = 2.0f * u01(r.y) - 1.0f; }
    else { float z0, z1; box_muller(r.x, r.y, z0, z1); a0 = r_clamp(z0, -1.0f, 1.0f); a1 = r_clamp(z1, -1.0f, 1.0f); }
    action[i] = make_float2(a0, a1);
}

// GAE(lambda) over the rollout slab: one thread per env, t = T-1 .. 0; loads are issued WL_GAE_CHUNK steps ahead so the
// reverse scan is bandwidth- rather than latency-bound (17 B read + 8 B written per (t, env)).
#define WL_GAE_CHUNK 8
__global__ void __launch_bounds__(128)
wl_gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ last_val,
              const uint8_t* __restrict__ done, const uint8_t* __restrict__ tout, float gamma, float lam,
              float* __restrict__ ret, float* __restrict__ adv, int T, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float next_v = last_val[i], a = 0.0f;
    for (int t1 = T; t1 > 0; t1 -= WL_GAE_CHUNK) {
        const int t0 = (t1 - WL_GAE_CHUNK > 0) ? t1 - WL_GAE_CHUNK : 0;
        float r[WL_GAE_CHUNK], v[WL_GAE_CHUNK]; uint8_t d[WL_GAE_CHUNK], to[WL_GAE_CHUNK];
#pragma unroll
        for (int k = 0; k < WL_GAE_CHUNK; ++k) {
            const int t = t1 - 1 - k;
            if (t >= t0) {
                const size_t o = (size_t)t * N + i;
                r[k] = rew[o]; v[k] = val[o]; d[k] = done[o]; to[k] = tout ? tout[o] : (uint8_t)0;
            }
        }
#pragma unroll
        for (int k = 0; k < WL_GAE_CHUNK; ++k) {
            const int t = t1 - 1 - k;
            if (t >= t0) {
                const float nt = d[k] ? 0.0f : 1.0f;
                const float rr = to[k] ? r[k] + gamma * v[k] : r[k];           // time-out bootstrap
                const float delta = rr + nt * gamma * next_v - v[k];
                a = delta + nt * gamma * lam * a;
                const size_t o = (size_t)t * N + i;
                ret[o] = a + v[k]; adv[o] = a;
                next_v = v[k];
            }
        }
    }
}

// GAE for SMALL N (the reference's 128 x 4096 rollout): the backward recurrence a_t = delta_t + k_t a_{t+1} is affine, so
// the T steps of an env are cut into segments of WL_GAE_SEG steps, one thread per (env, segment): every thread loads its
// whole segment at once (one exposed memory latency instead of T / WL_GAE_CHUNK), reduces it to (c, p) with
// a_start = c + p * a_in, the S segments of an env are chained through shared memory (S - 1 dependent FMAs), and the
// outputs are formed from registers.  CTA = 32 envs x S segments (warp = segment: every load / store is coalesced over envs).
// Rounding differs from the sequential scan by a few ulp (two-level association).
#define WL_GAE_SEG 16
__global__ void __launch_bounds__(512)
wl_gae_seg_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ last_val,
                  const uint8_t* __restrict__ done, const uint8_t* __restrict__ tout, float gamma, float lam,
                  float* __restrict__ ret, float* __restrict__ adv, int T, int N) {
    __shared__ float s_c[16][33], s_p[16][33], s_in[16][33];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5, S = blockDim.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const bool live = i < N;
    const int ii = live ? i : N - 1;
    const int t0 = seg * WL_GAE_SEG;                       // this thread owns steps [t0, t0 + WL_GAE_SEG) intersected with [0, T)
    float r[WL_GAE_SEG], v[WL_GAE_SEG], kk[WL_GAE_SEG];
    uint8_t d[WL_GAE_SEG], to[WL_GAE_SEG];
#pragma unroll
    for (int k = 0; k < WL_GAE_SEG; ++k) {
        const int t = t0 + k;
        if (t < T) {
            const size_t o = (size_t)t * N + ii;
            r[k] = rew[o]; v[k] = val[o]; d[k] = done[o]; to[k] = tout ? tout[o] : (uint8_t)0;
        } else { r[k] = 0.0f; v[k] = 0.0f; d[k] = 1; to[k] = 0; }
    }
    const int t_end = min(t0 + WL_GAE_SEG, T);             // value that follows the segment: V(t_end) or the bootstrap value
    float next_v = (t_end < T) ? val[(size_t)t_end * N + ii] : last_val[ii];
    // backward over the segment: c = contribution with a_in = 0, p = product of the k_t
    float c = 0.0f, p = 1.0f;
#pragma unroll
    for (int k = WL_GAE_SEG - 1; k >= 0; --k) {
        if (t0 + k < T) {
            const float nt = d[k] ? 0.0f : 1.0f;
            const float rr = to[k] ? r[k] + gamma * v[k] : r[k];           // time-out bootstrap
            const float delta = rr + nt * gamma * next_v - v[k];
            const float kt = nt * gamma * lam;
            c = delta + kt * c; p = kt * p;
            r[k] = c; kk[k] = p;                                           // a_t = r[k] + kk[k] * a_in
            next_v = v[k];
        }
    }
    s_c[seg][lane] = c; s_p[seg][lane] = p;
    __syncthreads();
    if (seg == 0) {                                        // chain the env's segments: a_in of segment s = a at the start of segment s + 1
        float a_in = 0.0f;
        for (int q = S - 1; q >= 0; --q) { s_in[q][lane] = a_in; a_in = s_c[q][lane] + s_p[q][lane] * a_in; }
    }
    __syncthreads();
    const float a_in = s_in[seg][lane];
    if (live) {
#pragma unroll
        for (int k = 0; k < WL_GAE_SEG; ++k) {
            const int t = t0 + k;
            if (t < T) {
                const float a = r[k] + kk[k] * a_in;
                const size_t o = (size_t)t * N + i;
                ret[o] = a + v[k]; adv[o] = a;
            }
        }
    }
}

// Data-parallel learner step for the small rsl_rl networks: gradient all-reduce FUSED with the Adam update in
// one kernel over peer memory.  Every rank keeps its flat gradient in a symmetric (P2P-mapped) buffer; each thread reads
// element i of every rank's gradient (NVLink loads, fixed rank order -> every rank forms the identical mean), and applies
// Adam to its own replica of the parameters.  10.5 K parameters = 42 KB per rank: one launch instead of NCCL all-reduce +
// optimizer kernels.  The caller brackets the launch with barriers (gradients complete / safe to overwrite).
struct GradPeers { int32_t n; int32_t pad; const float* g[WL_MAX_PEERS]; };
__global__ void wl_dp_adam_kernel(float* __restrict__ param, float* __restrict__ m, float* __restrict__ v, const __grid_constant__ GradPeers gp,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = 0.0f;
    for (int r = 0; r < gp.n; ++r) g += __ldcg(gp.g[r] + i);           // (L2-coherent loads: peers have just written)
    g = g / (float)gp.n;
    float p = param[i];
    if (weight_decay != 0.0f) g = fmaf(weight_decay, p, g);
    const float mi = fmaf(beta1, m[i], (1.0f - beta1) * g);
    const float vi = fmaf(beta2, v[i], (1.0f - beta2) * g * g);
    m[i] = mi; v[i] = vi;
    param[i] = p - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);          // torch.optim.Adam's update (bias corrections bc = 1 - beta^t)
}

__global__ void wl_detmath_kernel(int op, const float* __restrict__ in, const float* __restrict__ in2, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = in[i], s, c2, r = 0.0f;
    switch (op) {
        case 0: det_sincos(x, s, c2); r = s; break;
        case 1: det_sincos(x, s, c2); r = c2; break;
        case 2: r = det_atan(x); break;
        case 3: r = det_atan2(in2[i], x); break;
        case 4: r = det_log(x); break;
        case 5: r = det_tan(x); break;
        case 6: r = det_asin(x); break;
        case 7: r = det_exp(x); break;
        case 8: r = det_tanh(x); break;
        case 9: r = fdiv_norm(in2[i], x); break;          // in2 / in
        case 10: r = fsqrt_norm(x); break;
        case 11: r = det_atan_ratio<true>(in2[i], x); break;
        case 12: r = det_atan_ratio<false>(in2[i], x); break;
    }
    out[i] = r;
}
__global__ void wl_null_kernel(int* p) { if (p != nullptr && threadIdx.x == 1024) *p = 0; }
__global__ void wl_null_cfg_kernel(const __grid_constant__ wl_config c, int* p) { if (p != nullptr && threadIdx.x == 1024) *p = c.num_envs; }
__global__ void wl_philox_kernel(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint4* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = philox4x32(seed, c0 + (uint32_t)i, c1, c2, c3);
}

// ---------------------------------------------------------------------------------------
// Visual task camera term (mdp_sensors/observations.py:64-87 on a software pinhole camera): one CTA per env renders the
// kept rows of the 80 x 60 frame against the 2-colour plane mesh, applies ColorJitter on the two class values, the 5 x 5
// Gaussian blur (separable, reflect padding) through shared memory, Grayscale + Normalize(0.5, 0.5), and streams the
// 3200 floats out as float4.  HBM: 12.8 KB written per env, the 250 KB map stays in L2.
// ---------------------------------------------------------------------------------------
#define WL_CAM_THREADS 128
#define WL_CAM_MAX_W 128
#define WL_CAM_MAX_ROWS 64
__global__ void __launch_bounds__(WL_CAM_THREADS)
wl_camera_kernel(const __grid_constant__ wl_config c, const float4* __restrict__ st, const wl_globals* __restrict__ gl,
                 const float* __restrict__ aux, float* __restrict__ obs, uint32_t t_arg, uint32_t stream, uint32_t sub,
                 const float* __restrict__ aug) {
    // Thread (u4, rg) owns 4 consecutive columns 4*u4 .. 4*u4+3 of rows rg, rg + RG, ...: row terms are shared by 4 pixels
    // and every shared-memory access is a vector one.  Jp: white mask (bytes) with 4 pad columns on either side
    // (only the inner 2 are used: reflected border); Hp: horizontal pass, with 2 reflected rows above and below.
    __shared__ __align__(16) uint8_t Jp[WL_CAM_MAX_PIXELS + 8 * WL_CAM_MAX_ROWS];     // the mask is kept as bytes: 9 CTAs per SM
    __shared__ CamAug A_s;
    __shared__ __align__(16) float Hp[WL_CAM_MAX_PIXELS + 4 * WL_CAM_MAX_W];
    __shared__ __align__(16) float CX[WL_CAM_MAX_W], CY[WL_CAM_MAX_W], CZ[WL_CAM_MAX_W];   // per-column part of the ray direction
    __shared__ float yo_s[WL_CAM_MAX_ROWS];
    __shared__ int n_white;
    const int i = blockIdx.x, n = c.num_envs, tid = threadIdx.x;
    const int W = c.vis_cam_w, rows = c.vis_cam_h - c.vis_cam_row0, npix = W * rows, WP = W + 8, W4 = W >> 2;
    const int RG = WL_CAM_THREADS / W4;                      // row groups (threads >= RG * W4 idle in the 2-D phases)
    const int rg = tid / W4, u4 = tid - rg * W4;
    const bool act = rg < RG;
    const uint32_t t = decode_step(gl, t_arg);
    const VisualMap vm = visual_map(c, aux);
    const float4 gp = ldg4(st, WL_G_POS, n, i), gq = ldg4(st, WL_G_QUAT, n, i);
    const M3 R = rotmat(gq.x, gq.y, gq.z, gq.w);
    const V3 off = rot(R, V3{c.vis_cam_pos[0], c.vis_cam_pos[1], c.vis_cam_pos[2]});
    const V3 pc{gp.x + off.x, gp.y + off.y, gp.z + off.z};
    if (tid == 0) n_white = 0;
    // pixel-centre coordinates of the optical frame (x right, y down, z forward) and the column part of R (1, -xo, -yo)
    if (tid < W) {
        const float xo = ((float)tid + 0.5f - c.vis_cam_cx) / c.vis_cam_fx;
        CX[tid] = fm(-R.r[1], xo, R.r[0]); CY[tid] = fm(-R.r[4], xo, R.r[3]); CZ[tid] = fm(-R.r[7], xo, R.r[6]);
    }
    if (tid < rows) yo_s[tid] = ((float)(c.vis_cam_row0 + tid) + 0.5f - c.vis_cam_cy) / c.vis_cam_fy;
    const bool bg = c.vis_cam_bg >= 0.5f, above = pc.z > 0.0f;
    const float cmax = (float)(c.vis_cols - 1), rmax = (float)(c.vis_rows - 1);
    __syncthreads();
    // ---- render: ray through the pixel centre -> plane z = 0 -> face of the coloured mesh (utils/__init__.py:8-89)
    int mine = 0;
    if (act) {
        const float4 cx = reinterpret_cast<const float4*>(CX)[u4], cy = reinterpret_cast<const float4*>(CY)[u4], cz = reinterpret_cast<const float4*>(CZ)[u4];
        const float cxs[4] = {cx.x, cx.y, cx.z, cx.w}, cys[4] = {cy.x, cy.y, cy.z, cy.w}, czs[4] = {cz.x, cz.y, cz.z, cz.w};
        for (int r = rg; r < rows; r += RG) {
            const float yo = yo_s[r];
            uint8_t wv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dz = fm(-R.r[8], yo, czs[j]);
                bool wh = bg;
                if (dz < 0.0f && above) {
                    const float tt = pc.z / (-dz);                                  // depth along the optical axis
                    if (!(tt > 100.0f)) {                                           // clipping_range = (0.01, 1e2)
                        const float hx = fm(tt, fm(-R.r[2], yo, cxs[j]), pc.x), hy = fm(tt, fm(-R.r[5], yo, cys[j]), pc.y);
                        const float fx = floorf((hx - c.vis_mesh_x0) * c.d_vis_mesh_inv_dx), fy = floorf((hy - c.vis_mesh_y0) * c.d_vis_mesh_inv_dy);
                        wh = (fx >= 0.0f) && (fy >= 0.0f) && (fx < cmax) && (fy < rmax) && (__ldg(vm.map + (size_t)(int)fy * c.vis_cols + (int)fx) != 0);
                    }
                }
                wv[j] = wh ? 1 : 0;
                mine += wh ? 1 : 0;
            }
            *reinterpret_cast<uchar4*>(Jp + r * WP + 4 + 4 * u4) = make_uchar4(wv[0], wv[1], wv[2], wv[3]);
        }
    }
    mine = (int)warp_sum((float)mine);           // <= 32 * 64: exact in fp32
    if ((tid & 31) == 0 && mine) atomicAdd(&n_white, mine);
    __syncthreads();
    if (tid == WL_CAM_THREADS - 1)                // one thread draws the frame's ColorJitter / blur parameters for the CTA
        A_s = cam_aug_params(c, t, stream, sub, aug, (float)n_white / (float)npix);
    float4* row4 = reinterpret_cast<float4*>(obs + (size_t)(WL_OBS_DIM_VISUAL + npix) * i);
    if (c.vis_cam != 2) {                        // camera_data_rgb_flattened: grayscale + normalize only
        const float g0 = fm(2.0f, cam_gray(0.0f), -1.0f), g1 = fm(2.0f, cam_gray(1.0f), -1.0f);
        if (act)
            for (int r = rg; r < rows; r += RG) {
                const uchar4 m = *reinterpret_cast<const uchar4*>(Jp + r * WP + 4 + 4 * u4);
                row4[r * W4 + u4] = make_float4(m.x ? g1 : g0, m.y ? g1 : g0, m.z ? g1 : g0, m.w ? g1 : g0);
            }
        return;
    }
    // reflected border columns: -1 -> 1, -2 -> 2, W -> W-2, W+1 -> W-3
    if (tid < rows) {
        uint8_t* jr = Jp + tid * WP + 4;
        jr[-1] = jr[1]; jr[-2] = jr[2]; jr[W] = jr[W - 2]; jr[W + 1] = jr[W - 3];
    }
    __syncthreads();
    const CamAug A = A_s;
    // ---- horizontal 5-tap pass on the mask; padded row rp holds source row reflect(rp - 2)
    if (act)
        for (int rp = rg; rp < rows + 4; rp += RG) {
            int r = rp - 2;
            r = r < 0 ? -r : (r >= rows ? 2 * rows - 2 - r : r);
            const uchar4* j4 = reinterpret_cast<const uchar4*>(Jp + r * WP) + u4;
            const uchar4 Lb = j4[0], Mb = j4[1], Qb = j4[2];
            const float4 L = make_float4(0.0f, 0.0f, (float)Lb.z, (float)Lb.w), M = make_float4((float)Mb.x, (float)Mb.y, (float)Mb.z, (float)Mb.w);
            const float4 Q = make_float4((float)Qb.x, (float)Qb.y, 0.0f, 0.0f);
            float4 o;
            o.x = fm(A.w2, L.z + M.z, fm(A.w1, L.w + M.y, A.w0 * M.x));
            o.y = fm(A.w2, L.w + M.w, fm(A.w1, M.x + M.z, A.w0 * M.y));
            o.z = fm(A.w2, M.x + Q.x, fm(A.w1, M.y + M.w, A.w0 * M.z));
            o.w = fm(A.w2, M.y + Q.y, fm(A.w1, M.z + Q.x, A.w0 * M.w));
            reinterpret_cast<float4*>(Hp + rp * W)[u4] = o;
        }
    __syncthreads();
    // ---- vertical pass, then the ColorJitter class values (blur is affine in the mask: J = v0 + (v1 - v0) w),
    // Grayscale and Normalize((x - 0.5) / 0.5); one STG.128 per thread and row
    if (act) {
        const float dv = A.v1 - A.v0;
        for (int r = rg; r < rows; r += RG) {
            const float4* h = reinterpret_cast<const float4*>(Hp + r * W) + u4;      // padded rows r .. r+4 = source rows r-2 .. r+2
            const float4 a = h[0], b = h[W4], m = h[2 * W4], d = h[3 * W4], e = h[4 * W4];
            float4 o;
            o.x = fm(2.0f, cam_gray(fm(dv, fm(A.w2, a.x + e.x, fm(A.w1, b.x + d.x, A.w0 * m.x)), A.v0)), -1.0f);
            o.y = fm(2.0f, cam_gray(fm(dv, fm(A.w2, a.y + e.y, fm(A.w1, b.y + d.y, A.w0 * m.y)), A.v0)), -1.0f);
            o.z = fm(2.0f, cam_gray(fm(dv, fm(A.w2, a.z + e.z, fm(A.w1, b.z + d.z, A.w0 * m.z)), A.v0)), -1.0f);
            o.w = fm(2.0f, cam_gray(fm(dv, fm(A.w2, a.w + e.w, fm(A.w1, b.w + d.w, A.w0 * m.w)), A.v0)), -1.0f);
            row4[r * W4 + u4] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------
// launch geometry: spread small N over all 148 SMs, use fatter CTAs once the chip is full
// ---------------------------------------------------------------------------------------
// launch with (or without) the programmatic-dependent-launch attribute
// (measured on B200, 4096-env Drift step, back-to-back launches: 9.0 us/step with the attribute vs 6.65 us without --
// profiles/r02_kexp_c_pdl.jsonl -- the dependent grid's early-resident CTAs cost more than the launch gap they hide; the
// attribute is therefore OFF unless WL_PDL=1)
static bool g_pdl = (getenv("WL_PDL") != nullptr) && (atoi(getenv("WL_PDL")) != 0);
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t cs, const Args&... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count");
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)block); lc.dynamicSmemBytes = smem; lc.stream = cs;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = g_pdl ? 1 : 0;
    // parameters by address: the 1.5 KB wl_config is read in place by the runtime, not copied through a by-value call chain.
    // (each argument must already have exactly the kernel's parameter type: checked below)
    static_assert((std::is_same<typename std::decay<KArgs>::type, typename std::decay<Args>::type>::value && ...), "argument types must match the kernel's");
    void* ptrs[] = {const_cast<void*>(static_cast<const void*>(&args))...};
    cudaLaunchKernelExC(&lc, reinterpret_cast<const void*>(kernel), ptrs);
}

static inline int pick_block(int n) {
    if (n >= 148 * 128 * 4) return 128;
    if (n >= 148 * 64 * 2) return 64;
    return 32;
}

static int launch_scan(wl_sim* sim, float* d_obs, cudaStream_t cs) {
    const int n = sim->cfg.num_envs;
    if (sim->scan_mode == 2 && sim->has_tmap) {              // TMA producer/consumer pipeline, persistent CTAs
        static const int per_sm = [] { const char* e = getenv("WL_SCAN_CTAS_PER_SM"); int v = e ? atoi(e) : 6; return v < 1 ? 1 : (v > 12 ? 12 : v); }();
        const int grid = n < 148 * per_sm ? n : 148 * per_sm;    // (WL_SCAN_CTAS_PER_SM: tuning experiments only)
        wl_scan_pipe_kernel<<<grid, WL_SCAN_THREADS + 32, 0, cs>>>(sim->cfg, sim->tmap, sim->state, d_obs);
    } else if (sim->scan_mode >= 1 && sim->has_tmap)
        wl_scan_kernel<true><<<n, WL_SCAN_THREADS, 0, cs>>>(sim->cfg, sim->tmap, sim->state, sim->hf, d_obs);
    else
        wl_scan_kernel<false><<<n, WL_SCAN_THREADS, 0, cs>>>(sim->cfg, sim->tmap, sim->state, sim->hf, d_obs);
    sim->launches++;
    return cuda_check(cudaGetLastError(), "wl_scan_kernel");
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static int launch_camera(wl_sim* sim, float* d_obs, uint32_t t, uint32_t stream_id, uint32_t sub, const float* d_aug, cudaStream_t cs) {
    wl_camera_kernel<<<sim->cfg.num_envs, WL_CAM_THREADS, 0, cs>>>(sim->cfg, sim->state, sim->globals, sim->hf, d_obs, t, stream_id, sub, d_aug);
    sim->launches++;
    return cuda_check(cudaGetLastError(), "wl_camera_kernel");
}

extern "C" {

const char* wl_last_error(void) { return g_err.c_str(); }
const char* wl_build_info(void) { return "wheeledlab_b200 abi=2 arch=sm_100a fmad=false"; }
size_t wl_config_sizeof(void) { return sizeof(wl_config); }

const char* wl_config_describe(void) {
    static std::string s;
    if (s.empty()) {
        char buf[256];
#define WL_XS(type, tag, name)                                                                  \
    snprintf(buf, sizeof buf, "%s:%s:1:%zu;", #name, #tag, offsetof(wl_config, name)); s += buf;
#define WL_XA(type, tag, name, n)                                                               \
    snprintf(buf, sizeof buf, "%s:%s:%d:%zu;", #name, #tag, (int)(n), offsetof(wl_config, name)); s += buf;
        WL_CONFIG_FIELDS(WL_XS, WL_XA)
#undef WL_XS
#undef WL_XA
        snprintf(buf, sizeof buf, "sizeof:%zu", sizeof(wl_config)); s += buf;
    }
    return s.c_str();
}

int wl_set_step_counter(wl_sim* sim, int64_t value, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_set_step_counter: null handle");
    const uint32_t v = (uint32_t)value;
    sim->base_host = value;
    return cuda_check(cudaMemcpyAsync(&sim->globals->step_base, &v, sizeof v, cudaMemcpyHostToDevice, (cudaStream_t)stream),
                      "wl_set_step_counter");
}
int wl_advance_counter(wl_sim* sim, int32_t K, void* stream) {
    if (!sim || K < 0) return fail(WL_EINVAL, "wl_advance_counter: bad argument");
    ensure_device(sim);
    wl_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->globals, (uint32_t)K);
    sim->launches++;
    sim->base_host += K;
    return cuda_check(cudaGetLastError(), "wl_advance_kernel");
}
int wl_note_device_counter(wl_sim* sim, int64_t value) {
    if (!sim) return fail(WL_EINVAL, "wl_note_device_counter: null handle");
    sim->base_host = value;
    sim->last_t = value - 1;          // the replayed graph ended on step value - 1
    return WL_OK;
}
int wl_log_flush(wl_sim* sim, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_log_flush: null handle");
    ensure_device(sim);
    if (sim->last_t < 0) return WL_OK;
    wl_flush_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->cfg, sim->globals, (uint32_t)(sim->last_t + 1));
    sim->launches++;
    return cuda_check(cudaGetLastError(), "wl_flush_kernel");
}
float* wl_reward_weights(wl_sim* sim) { return sim ? sim->globals->rew_weight[(uint32_t)sim->last_t & 1u] : nullptr; }

int wl_set_term_bits(wl_sim* sim, uint8_t* d_term_bits) {
    if (!sim) return fail(WL_EINVAL, "wl_set_term_bits: null handle");
    sim->term_bits = d_term_bits;
    return WL_OK;
}
int wl_set_peer_fanout(wl_sim* sim, int32_t n_peers, const int64_t* byte_deltas) {
    if (!sim || n_peers < 0 || n_peers > WL_MAX_PEERS || (n_peers > 0 && !byte_deltas)) return fail(WL_EINVAL, "wl_set_peer_fanout: bad argument");
    if (n_peers > 0 && sim->cfg.task != WL_TASK_DRIFT) return fail(WL_EUNSUPPORTED, "wl_set_peer_fanout: Drift-family tasks only");
    const int32_t mc = sim->fan.mc; const long long mcd = sim->fan.mc_delta;
    memset(&sim->fan, 0, sizeof sim->fan);
    sim->fan.mc = mc; sim->fan.mc_delta = mcd;
    sim->fan.n = n_peers;
    for (int k = 0; k < n_peers; ++k) sim->fan.delta[k] = (long long)byte_deltas[k];
    return WL_OK;
}
int wl_set_multicast_fanout(wl_sim* sim, int64_t mc_byte_delta) {
    if (!sim) return fail(WL_EINVAL, "wl_set_multicast_fanout: null handle");
    if (mc_byte_delta != 0 && sim->cfg.task != WL_TASK_DRIFT) return fail(WL_EUNSUPPORTED, "wl_set_multicast_fanout: Drift-family tasks only");
    sim->fan.mc = mc_byte_delta != 0 ? 1 : 0;
    sim->fan.mc_delta = (long long)mc_byte_delta;
    return WL_OK;
}
int wl_set_seed(wl_sim* sim, uint64_t seed) {
    if (!sim) return fail(WL_EINVAL, "wl_set_seed: null handle");
    sim->cfg.seed = seed;
    return WL_OK;
}

int wl_config_finalize(wl_config* c) {
    if (!c) return fail(WL_EINVAL, "wl_config_finalize: null");
    if (c->substeps <= 0 || !(c->sim_dt > 0.0f)) return fail(WL_EINVAL, "wl_config_finalize: bad sim timing");
    c->d_h = c->sim_dt / (float)c->substeps;
    c->d_inv_h = 1.0f / c->d_h;
    c->d_step_dt = c->sim_dt * (float)c->decimation;
    c->d_hkp = c->d_h * c->steer_kp;
    c->d_sden = 1.0f / fmaf(c->d_h, c->d_hkp, fmaf(c->d_h, c->steer_kd, c->steer_inertia));
    c->d_inv_Iw = 1.0f / c->wheel_inertia;
    c->d_hI = c->d_h * c->d_inv_Iw;
    c->d_inv_hf_cell = c->hf_cell > 0.0f ? 1.0f / c->hf_cell : 0.0f;
    c->d_fxk = c->tire_mx * c->d_inv_h;
    c->d_fyk = c->tire_my * c->d_inv_h;
    c->d_inv_wheel_radius_cfg = 1.0f / c->wheel_radius_cfg;
    c->d_inv_dc_vel_limit = 1.0f / c->dc_vel_limit;
    c->d_inv_mass_nominal = 1.0f / c->mass_nominal;
    for (int a = 0; a < 3; ++a) c->d_invI_nominal[a] = 1.0f / c->inertia_nominal[a];
    c->d_vis_mesh_inv_dx = c->vis_mesh_dx > 0.0f ? 1.0f / c->vis_mesh_dx : 0.0f;
    c->d_vis_mesh_inv_dy = c->vis_mesh_dy > 0.0f ? 1.0f / c->vis_mesh_dy : 0.0f;
    return WL_OK;
}

size_t wl_globals_offset(int32_t num_envs) { return align256(groups_bytes(num_envs)); }
size_t wl_state_bytes(int32_t num_envs) { return wl_globals_offset(num_envs) + align256(sizeof(wl_globals)); }

int wl_create(const wl_config* cfg, void* d_state, size_t state_bytes, const float* d_heightfield, wl_sim** out) {
    if (!cfg || !d_state || !out) return fail(WL_EINVAL, "wl_create: null argument");
    if (cfg->abi_version != WL_ABI_VERSION) return fail(WL_EINVAL, "wl_create: abi_version mismatch");
    if (cfg->num_envs <= 0) return fail(WL_EINVAL, "wl_create: num_envs must be > 0");
    if (state_bytes < wl_state_bytes(cfg->num_envs)) return fail(WL_EINVAL, "wl_create: state buffer too small");
    if (((uintptr_t)d_state & 255u) != 0) return fail(WL_EINVAL, "wl_create: state buffer must be 256-byte aligned");
    if (cfg->task != WL_TASK_DRIFT && cfg->task != WL_TASK_ELEVATION && cfg->task != WL_TASK_VISUAL)
        return fail(WL_EUNSUPPORTED, "wl_create: unknown task");
    if (cfg->task == WL_TASK_VISUAL) {
        if (!d_heightfield) return fail(WL_EINVAL, "wl_create: the visual task needs the traversability data (see header)");
        if (cfg->vis_rows < 1 || cfg->vis_cols < 1 || cfg->vis_n_trav < 1) return fail(WL_EINVAL, "wl_create: bad traversability map geometry");
        if (((uintptr_t)d_heightfield & 15u) != 0) return fail(WL_EINVAL, "wl_create: traversability data must be 16-byte aligned");
        if (cfg->vis_cam) {
            const int rows = cfg->vis_cam_h - cfg->vis_cam_row0;
            if (cfg->vis_cam < 0 || cfg->vis_cam > 2 || cfg->vis_cam_w < 4 || (cfg->vis_cam_w & 3) || rows < 3 || cfg->vis_cam_row0 < 0 ||
                cfg->vis_cam_w * rows > WL_CAM_MAX_PIXELS || cfg->vis_cam_w > 128 || rows > 64 || !(cfg->vis_cam_fx > 0.0f) || !(cfg->vis_cam_fy > 0.0f) ||
                !(cfg->vis_mesh_dx > 0.0f) || !(cfg->vis_mesh_dy > 0.0f) || (cfg->vis_cam == 2 && !(cfg->vis_aug_sigma[0] > 0.0f)))
                return fail(WL_EINVAL, "wl_create: bad camera geometry (width % 4 == 0 and <= 128, 3..64 kept rows, <= WL_CAM_MAX_PIXELS pixels)");
        }
    }
    if (cfg->task == WL_TASK_ELEVATION) {
        if (!d_heightfield) return fail(WL_EINVAL, "wl_create: the elevation task needs a height-field");
        if (cfg->hf_nx < 2 || cfg->hf_ny < 2 || cfg->hf_pitch < cfg->hf_nx || (cfg->hf_pitch & 3) || !(cfg->hf_cell > 0.0f))
            return fail(WL_EINVAL, "wl_create: bad height-field geometry (hf_pitch must be >= hf_nx and a multiple of 4)");
        if (((uintptr_t)d_heightfield & 15u) != 0) return fail(WL_EINVAL, "wl_create: height-field must be 16-byte aligned");
    }
    if (cfg->bounding != WL_BOUND_CLIP && cfg->bounding != WL_BOUND_NONE && cfg->bounding != WL_BOUND_TANH)
        return fail(WL_EUNSUPPORTED, "wl_create: unknown bounding strategy");
    if (cfg->action_kind != WL_ACT_ACKERMANN && cfg->action_kind != WL_ACT_RWD && cfg->action_kind != WL_ACT_4WD)
        return fail(WL_EUNSUPPORTED, "wl_create: unknown action term kind");
    if (cfg->dr_wheel_mass_enable && !(cfg->dr_wheel_mass[0] > 0.0f && cfg->wheel_mass_nominal > 0.0f))
        return fail(WL_EINVAL, "wl_create: wheel-mass DR needs positive masses");
    if (cfg->decimation <= 0 || cfg->substeps <= 0 || !(cfg->sim_dt > 0.0f)) return fail(WL_EINVAL, "wl_create: bad sim timing");
    if (cfg->num_rew_terms < 0 || cfg->num_rew_terms > WL_MAX_REW_TERMS) return fail(WL_EINVAL, "wl_create: num_rew_terms");
    if (cfg->curr_n < 0 || cfg->curr_n > 4) return fail(WL_EINVAL, "wl_create: curr_n must be in [0,4]");
    for (int k = 0; k < cfg->curr_n; ++k)
        if (cfg->curr_every[k] < 1 || cfg->curr_slot[k] < 0 || cfg->curr_slot[k] >= WL_MAX_REW_TERMS)
            return fail(WL_EINVAL, "wl_create: bad curriculum term");
    if (cfg->num_ref_poses <= 0 || cfg->num_ref_poses > WL_MAX_REF_POSES) return fail(WL_EINVAL, "wl_create: num_ref_poses");
    if (cfg->dr_num_buckets < 1 || cfg->dr_num_buckets > WL_MAX_BUCKETS) return fail(WL_EINVAL, "wl_create: dr_num_buckets");
    int dev_count = 0;
    if (int rc = cuda_check(cudaGetDeviceCount(&dev_count), "cudaGetDeviceCount")) return rc;
    if (dev_count == 0) return fail(WL_ECUDA, "wl_create: no CUDA device (this library has no CPU path)");
    wl_sim* s = new (std::nothrow) wl_sim();
    if (!s) return fail(WL_EINVAL, "wl_create: out of host memory");
    s->cfg = *cfg;
    if (int rc = wl_config_finalize(&s->cfg)) { delete s; return rc; }
    s->state = reinterpret_cast<float4*>(d_state);
    s->globals = reinterpret_cast<wl_globals*>(reinterpret_cast<char*>(d_state) + wl_globals_offset(cfg->num_envs));
    s->hf = d_heightfield;
    s->state_bytes = state_bytes;
    s->launches = 0;
    s->last_t = -1;
    s->base_host = 0;
    s->term_bits = nullptr;
    memset(&s->fan, 0, sizeof s->fan);
    s->hp_n = 0;
    cudaGetDevice(&s->device);
    s->variant = 0;
    s->has_tmap = false;
    s->scan_mode = 1;
    if (cfg->task == WL_TASK_ELEVATION) {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (int rc = cuda_check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres),
                                "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)")) { delete s; return rc; }
        if (!fn || qres != cudaDriverEntryPointSuccess) { delete s; return fail(WL_ECUDA, "cuTensorMapEncodeTiled not available"); }
        cuuint64_t gdim[2] = {(cuuint64_t)cfg->hf_nx, (cuuint64_t)cfg->hf_ny};
        cuuint64_t gstride[1] = {(cuuint64_t)cfg->hf_pitch * sizeof(float)};
        cuuint32_t box[2] = {WL_TILE_W, WL_TILE_H}, estr[2] = {1, 1};
        CUresult cr = ((encode_fn)fn)(&s->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)d_heightfield, gdim, gstride, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { delete s; return fail(WL_ECUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)cr) + ")"); }
        s->has_tmap = true;
    }
    s->obs_dim = (cfg->task == WL_TASK_ELEVATION) ? WL_OBS_DIM_ELEV : (cfg->task == WL_TASK_VISUAL) ? WL_OBS_DIM_VISUAL + (cfg->vis_cam ? cfg->vis_cam_w * (cfg->vis_cam_h - cfg->vis_cam_row0) : 0) : WL_OBS_DIM_BLIND;
    // live reward weights
    for (int slot = 0; slot < 2; ++slot)
        if (int rc = cuda_check(cudaMemcpy(s->globals->rew_weight[slot], cfg->rew_weight, sizeof(float) * WL_MAX_REW_TERMS,
                                           cudaMemcpyHostToDevice), "upload reward weights")) { delete s; return rc; }
    {
        const uint32_t none = 0xFFFFFFFFu;
        if (int rc = cuda_check(cudaMemcpy(&s->globals->curr_applied_t, &none, sizeof none, cudaMemcpyHostToDevice), "init globals")) { delete s; return rc; }
    }
    *out = s;
    return WL_OK;
}

int wl_destroy(wl_sim* sim) { delete sim; return WL_OK; }
int32_t wl_obs_dim(const wl_sim* sim) { return sim ? sim->obs_dim : 0; }
int wl_set_scan_tma(wl_sim* sim, int32_t use_tma) {
    if (!sim) return fail(WL_EINVAL, "wl_set_scan_tma: null handle");
    if (use_tma < 0 || use_tma > 2) return fail(WL_EINVAL, "wl_set_scan_tma: 0, 1 or 2");
    sim->scan_mode = use_tma;
    return WL_OK;
}
int wl_set_kernel_variant(wl_sim* sim, int32_t lanes_per_env) {
    if (!sim) return fail(WL_EINVAL, "wl_set_kernel_variant: null handle");
    if (lanes_per_env != 0 && lanes_per_env != 1 && lanes_per_env != 4 && lanes_per_env != 8)
        return fail(WL_EINVAL, "wl_set_kernel_variant: 0, 1, 4 or 8");
    sim->variant = lanes_per_env;
    return WL_OK;
}
int64_t wl_launch_count(const wl_sim* sim) { return sim ? sim->launches : 0; }

#define WL_LAUNCH_CHECK(sim, what)                                                     \
    do {                                                                               \
        (sim)->launches++;                                                             \
        if (int rc_ = cuda_check(cudaGetLastError(), what)) return rc_;                \
    } while (0)

}   // extern "C"
// Steps are issued with consecutive counters (the janitor protocol, see wl_globals).  Resolve this launch's counter on the host
// mirror; on a jump publish the pending log row, carry the weights to the slot the step will read and clear the rows.
static int prep_step(wl_sim* sim, int64_t step_counter, int32_t n_steps, cudaStream_t cs) {
    ensure_device(sim);
    const int64_t t = step_counter >= 0 ? step_counter : sim->base_host + (-1 - step_counter);
    if (t >= ((int64_t)1 << 31) - n_steps) return fail(WL_EINVAL, "step counter out of range (< 2^31)");
    if (t != sim->last_t + 1) {
        if (sim->last_t >= 0) wl_flush_kernel<<<1, 32, 0, cs>>>(sim->cfg, sim->globals, (uint32_t)(sim->last_t + 1));
        wl_rearm_kernel<<<1, 32, 0, cs>>>(sim->globals, (uint32_t)sim->last_t & 1u, (uint32_t)(t - 1) & 1u);
        sim->launches += 2;
        if (int rc = cuda_check(cudaGetLastError(), "wl_rearm_kernel")) return rc;
    }
    sim->last_t = t + n_steps - 1;
    return WL_OK;
}
// after the step kernel(s): a curriculum boundary reached with a host-supplied counter is applied in place right away
static int post_step(wl_sim* sim, int64_t step_counter, cudaStream_t cs) {
    if (step_counter < 0 || sim->cfg.curr_n <= 0) return WL_OK;
    const int64_t t_next = sim->last_t + 1;
    if (t_next % sim->cfg.max_episode_length != 0) return WL_OK;
    wl_boundary_kernel<<<1, 32, 0, cs>>>(sim->cfg, sim->globals, (uint32_t)t_next);
    sim->launches++;
    return cuda_check(cudaGetLastError(), "wl_boundary_kernel");
}
extern "C" {

int wl_startup(wl_sim* sim, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_startup: null handle");
    ensure_device(sim);
    const int n = sim->cfg.num_envs, bs = 128;
    wl_startup_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state);
    WL_LAUNCH_CHECK(sim, "wl_startup_kernel");
    return WL_OK;
}

int wl_reset(wl_sim* sim, const int64_t* d_env_ids, int32_t n_ids, int64_t step_counter, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_reset: null handle");
    ensure_device(sim);
    const int n = d_env_ids ? n_ids : sim->cfg.num_envs;
    if (n <= 0) return WL_OK;
    const int bs = 128;
    wl_reset_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, sim->hf, d_env_ids, n, (uint32_t)step_counter);
    WL_LAUNCH_CHECK(sim, "wl_reset_kernel");
    return WL_OK;
}

int wl_step(wl_sim* sim, const float* d_action, float* d_obs, float* d_rew, uint8_t* d_terminated, uint8_t* d_truncated,
            float* d_log, int64_t step_counter, void* stream) {
    if (!sim || !d_action || !d_obs || !d_rew || !d_terminated || !d_truncated) return fail(WL_EINVAL, "wl_step: null argument");
    if (((uintptr_t)d_action & 7u) || ((uintptr_t)d_obs & 7u)) return fail(WL_EINVAL, "wl_step: action/obs must be 8-byte aligned");
    const int n = sim->cfg.num_envs;
    Terrain T{sim->hf};
    // 8 = quad + aux warp (Drift family, small N: shortest dependent chain), 4 = quad, 1 = thread per env
    int variant = sim->variant ? sim->variant : ((n <= WL_QUAD_MAX_ENVS) ? ((sim->cfg.task == WL_TASK_DRIFT) ? 8 : 4) : 1);
    if (variant == 8 && sim->cfg.task != WL_TASK_DRIFT) variant = 4;
    const float2* act = reinterpret_cast<const float2*>(d_action);
    cudaStream_t cs = (cudaStream_t)stream;
    const uint32_t t = (uint32_t)step_counter;               // negative (device base + k) stays encoded: see decode_step
    if (int rc = prep_step(sim, step_counter, 1, cs)) return rc;
    const bool elev = sim->cfg.task == WL_TASK_ELEVATION, vis = sim->cfg.task == WL_TASK_VISUAL;
    StageIO sio0{sim->term_bits, nullptr, nullptr};
    if (variant == 8) {
        const int grid = (n + WL_DUO_ENVS - 1) / WL_DUO_ENVS + 1;                            // + the janitor CTA
        const bool rows_ok = (n % WL_DUO_ENVS) == 0 && (((uintptr_t)d_obs | (uintptr_t)d_rew) & 15u) == 0 &&
                             (((uintptr_t)d_terminated | (uintptr_t)d_truncated) & 7u) == 0;
        if (sim->fan.mc && rows_ok)
            launch_k(wl_step_duo_kernel<2>, grid, 64, 0, cs, sim->cfg, sim->state, sim->globals, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits, sim->fan);
        else if (sim->fan.n > 0 || sim->fan.mc)
            launch_k(wl_step_duo_kernel<1>, grid, 64, 0, cs, sim->cfg, sim->state, sim->globals, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits, sim->fan);
        else
            launch_k(wl_step_duo_kernel<0>, grid, 64, 0, cs, sim->cfg, sim->state, sim->globals, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits, sim->fan);
    } else if (sim->fan.n > 0) {
        return fail(WL_EUNSUPPORTED, "wl_step: the peer fan-out is implemented by the Drift-family small-N kernel (variant 8) only");
    } else if (variant == 4) {
#ifndef WL_QUAD_BS
#define WL_QUAD_BS 32
#endif
        const int bs = WL_QUAD_BS, threads = 4 * n, grid = (threads + bs - 1) / bs + 1;      // + the janitor CTA
        if (elev) launch_k(wl_step_quad_kernel<WL_TASK_ELEVATION>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits);
        else if (vis) launch_k(wl_step_quad_kernel<WL_TASK_VISUAL>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits);
        else launch_k(wl_step_quad_kernel<WL_TASK_DRIFT>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sim->term_bits);
    } else {
        const int bs = pick_block(n), grid = (n + bs - 1) / bs + 1;
        if (elev) launch_k(wl_step_kernel<WL_TASK_ELEVATION, 0>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio0);
        else if (vis) launch_k(wl_step_kernel<WL_TASK_VISUAL, 0>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio0);
        else launch_k(wl_step_kernel<WL_TASK_DRIFT, 0>, grid, bs, 0, cs, sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio0);
    }
    WL_LAUNCH_CHECK(sim, "wl_step_kernel");
    if (elev) { if (int rc = launch_scan(sim, d_obs, cs)) return rc; }
    if (vis && sim->cfg.vis_cam) { if (int rc = launch_camera(sim, d_obs, t, RNG_CAM, 0u, nullptr, cs)) return rc; }   // t may be device-relative
    return post_step(sim, step_counter, cs);
}

// ---- env.step cut in two (host-side reward / termination terms run in between) ----------------------------------------
}   // extern "C" (templates have C++ linkage)
template <int STAGE>
static int launch_stage(wl_sim* sim, const float2* act, float* d_obs, float* d_rew, uint8_t* d_terminated, uint8_t* d_truncated,
                        float* d_log, uint32_t t, StageIO sio, cudaStream_t cs) {
    const int n = sim->cfg.num_envs, bs = pick_block(n), grid = (n + bs - 1) / bs + (STAGE == 2 ? 1 : 0);   // stage b: + the janitor CTA
    Terrain T{sim->hf};
    switch (sim->cfg.task) {
    case WL_TASK_ELEVATION:
        wl_step_kernel<WL_TASK_ELEVATION, STAGE><<<grid, bs, 0, cs>>>(sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio);
        break;
    case WL_TASK_VISUAL:
        wl_step_kernel<WL_TASK_VISUAL, STAGE><<<grid, bs, 0, cs>>>(sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio);
        break;
    default:
        wl_step_kernel<WL_TASK_DRIFT, STAGE><<<grid, bs, 0, cs>>>(sim->cfg, sim->state, sim->globals, T, act, d_obs, d_rew, d_terminated, d_truncated, d_log, t, sio);
    }
    WL_LAUNCH_CHECK(sim, STAGE == 1 ? "wl_step_kernel<stage a>" : "wl_step_kernel<stage b>");
    return WL_OK;
}
extern "C" {

int wl_step_stage_a(wl_sim* sim, const float* d_action, float* d_rew, uint8_t* d_term_bits, int64_t step_counter, void* stream) {
    if (!sim || !d_action || !d_rew || !d_term_bits) return fail(WL_EINVAL, "wl_step_stage_a: null argument");
    if ((uintptr_t)d_action & 7u) return fail(WL_EINVAL, "wl_step_stage_a: action must be 8-byte aligned");
    if (step_counter < 0) return fail(WL_EINVAL, "wl_step_stage_a: the staged step takes the host's step counter");
    {   // stage a reads W(t) like a full step: re-arm on a counter jump, but the step is only committed by stage b
        const int64_t keep = sim->last_t;
        if (int rc = prep_step(sim, step_counter, 1, (cudaStream_t)stream)) return rc;
        sim->last_t = (keep == step_counter - 1) ? keep : step_counter - 1;
    }
    return launch_stage<1>(sim, reinterpret_cast<const float2*>(d_action), nullptr, d_rew, nullptr, nullptr, nullptr, (uint32_t)step_counter,
                           StageIO{d_term_bits, nullptr, nullptr}, (cudaStream_t)stream);
}

int wl_step_stage_b(wl_sim* sim, const uint8_t* d_term_bits, const uint8_t* d_extra_terminated, const uint8_t* d_extra_truncated,
                    float* d_obs, uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream) {
    if (!sim || !d_term_bits || !d_obs || !d_terminated || !d_truncated) return fail(WL_EINVAL, "wl_step_stage_b: null argument");
    if (step_counter < 0) return fail(WL_EINVAL, "wl_step_stage_b: the staged step takes the host's step counter");
    if (int rc = prep_step(sim, step_counter, 1, (cudaStream_t)stream)) return rc;
    if (int rc = launch_stage<2>(sim, nullptr, d_obs, nullptr, d_terminated, d_truncated, d_log, (uint32_t)step_counter,
                                 StageIO{const_cast<uint8_t*>(d_term_bits), d_extra_terminated, d_extra_truncated}, (cudaStream_t)stream))
        return rc;
    if (sim->cfg.task == WL_TASK_ELEVATION) { if (int rc = launch_scan(sim, d_obs, (cudaStream_t)stream)) return rc; }
    if (sim->cfg.task == WL_TASK_VISUAL && sim->cfg.vis_cam)
        if (int rc = launch_camera(sim, d_obs, (uint32_t)step_counter, RNG_CAM, 0u, nullptr, (cudaStream_t)stream)) return rc;
    return post_step(sim, step_counter, (cudaStream_t)stream);
}

int wl_rollout(wl_sim* sim, int32_t K, const float* d_actions, float* d_actions_out, float* d_obs, float* d_rew,
               uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream) {
    if (!sim || !d_obs || !d_rew || !d_terminated || !d_truncated || !d_log) return fail(WL_EINVAL, "wl_rollout: null argument");
    if (K < 1) return fail(WL_EINVAL, "wl_rollout: K must be >= 1");
    if (sim->cfg.task == WL_TASK_ELEVATION) return fail(WL_EUNSUPPORTED, "wl_rollout: the elevation step is two kernels (use wl_step)");
    if (sim->cfg.task == WL_TASK_VISUAL && sim->cfg.vis_cam) return fail(WL_EUNSUPPORTED, "wl_rollout: the camera term is a second kernel (use wl_step)");
    if (sim->cfg.curr_n > 0) {
        if (step_counter < 0) return fail(WL_EINVAL, "wl_rollout: with curriculum terms the host must pass the step counter");
        const int64_t L = sim->cfg.max_episode_length;
        // counters after steps 0..K-2 are t+1..t+K-1: none may be an episode boundary (only the last step may end on one)
        if ((step_counter + K - 1) / L != step_counter / L)
            return fail(WL_EINVAL, "wl_rollout: [t, t+K) must end at or before the next episode boundary of the global counter");
    }
    const int n = sim->cfg.num_envs;
    Terrain T{sim->hf};
    const int bs = 32, threads = 4 * n, grid = (threads + bs - 1) / bs;
    cudaStream_t cs = (cudaStream_t)stream;
    if (int rc = cuda_check(cudaMemsetAsync(d_log, 0, (size_t)K * WL_LOG_FLOATS * sizeof(float), cs), "wl_rollout: clear log rows")) return rc;
    if (int rc = prep_step(sim, step_counter, K, cs)) return rc;
    const float2* act = reinterpret_cast<const float2*>(d_actions);
    float2* aout = reinterpret_cast<float2*>(d_actions_out);
    if (sim->cfg.task == WL_TASK_VISUAL)
        wl_rollout_quad_kernel<WL_TASK_VISUAL><<<grid, bs, 0, cs>>>(sim->cfg, sim->state, sim->globals, T, act, aout, d_obs, d_rew, d_terminated, d_truncated, d_log, (uint32_t)step_counter, K, &sim->globals->ticket);
    else
        wl_rollout_quad_kernel<WL_TASK_DRIFT><<<grid, bs, 0, cs>>>(sim->cfg, sim->state, sim->globals, T, act, aout, d_obs, d_rew, d_terminated, d_truncated, d_log, (uint32_t)step_counter, K, &sim->globals->ticket);
    WL_LAUNCH_CHECK(sim, "wl_rollout_quad_kernel");
    return post_step(sim, step_counter, cs);
}

size_t wl_result_bytes(int32_t num_envs) { return (size_t)num_envs * 6u; }

int wl_step_host(wl_sim* sim, const float* h_action, float* d_action, float* d_obs, void* d_result, float* d_log,
                 void* h_result, float* h_obs, int64_t step_counter, void* stream) {
    if (!sim || !h_action || !d_action || !d_obs || !d_result || !h_result) return fail(WL_EINVAL, "wl_step_host: null argument");
    const size_t n = (size_t)sim->cfg.num_envs;
    cudaStream_t cs = (cudaStream_t)stream;
    if (int rc = cuda_check(cudaMemcpyAsync(d_action, h_action, n * 2 * sizeof(float), cudaMemcpyHostToDevice, cs), "H2D actions")) return rc;
    float* d_rew = reinterpret_cast<float*>(d_result);
    uint8_t* d_term = reinterpret_cast<uint8_t*>(d_result) + n * 4;
    uint8_t* d_trunc = d_term + n;
    if (int rc = wl_step(sim, d_action, d_obs, d_rew, d_term, d_trunc, d_log, step_counter, stream)) return rc;
    if (int rc = cuda_check(cudaMemcpyAsync(h_result, d_result, wl_result_bytes((int32_t)n), cudaMemcpyDeviceToHost, cs), "D2H results")) return rc;
    if (h_obs) {
        if (int rc = cuda_check(cudaMemcpyAsync(h_obs, d_obs, n * (size_t)sim->obs_dim * sizeof(float), cudaMemcpyDeviceToHost, cs), "D2H obs")) return rc;
    }
    return cuda_check(cudaStreamSynchronize(cs), "wl_step_host: stream synchronize");
}

// device alias of a pinned host block (cudaHostGetDevicePointer costs ~0.5 us per call: resolved once per block and cached)
static void* host_alias(wl_sim* sim, const void* h) {
    for (int k = 0; k < sim->hp_n; ++k) if (sim->hp_host[k] == h) return sim->hp_dev[k];
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void*>(h), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const int slot = sim->hp_n < 8 ? sim->hp_n++ : (int)(((uintptr_t)h >> 6) & 7u);
    sim->hp_host[slot] = h; sim->hp_dev[slot] = d;
    return d;
}

int wl_step_host_zero_copy(wl_sim* sim, const float* h_action, float* d_obs, float* d_log, void* h_result,
                           int64_t step_counter, void* stream) {
    if (!sim || !h_action || !d_obs || !h_result) return fail(WL_EINVAL, "wl_step_host_zero_copy: null argument");
    const size_t n = (size_t)sim->cfg.num_envs;
    const float* da = reinterpret_cast<const float*>(host_alias(sim, h_action));
    void* dr = host_alias(sim, h_result);
    if (!da || !dr) return fail(WL_EINVAL, "wl_step_host_zero_copy: h_action / h_result must be pinned (device-mapped) host memory");
    float* d_rew = reinterpret_cast<float*>(dr);
    uint8_t* d_term = reinterpret_cast<uint8_t*>(dr) + n * 4;
    uint8_t* d_trunc = d_term + n;
    if (int rc = wl_step(sim, da, d_obs, d_rew, d_term, d_trunc, d_log, step_counter, stream)) return rc;
    return cuda_check(cudaStreamSynchronize((cudaStream_t)stream), "wl_step_host_zero_copy: stream synchronize");
}

int wl_observe(wl_sim* sim, float* d_obs, int64_t step_counter, int32_t call_idx, void* stream) {
    if (!sim || !d_obs) return fail(WL_EINVAL, "wl_observe: null argument");
    ensure_device(sim);
    const int n = sim->cfg.num_envs, bs = 128;
    wl_observe_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, d_obs, (uint32_t)step_counter,
                                                                         (uint32_t)call_idx);
    WL_LAUNCH_CHECK(sim, "wl_observe_kernel");
    if (sim->cfg.task == WL_TASK_ELEVATION) return launch_scan(sim, d_obs, (cudaStream_t)stream);
    if (sim->cfg.task == WL_TASK_VISUAL && sim->cfg.vis_cam)
        return launch_camera(sim, d_obs, (uint32_t)step_counter, RNG_CAM_EXTRA, (uint32_t)call_idx, nullptr, (cudaStream_t)stream);
    return WL_OK;
}

int wl_camera(wl_sim* sim, float* d_obs, int64_t step_counter, const float* d_aug, void* stream) {
    if (!sim || !d_obs) return fail(WL_EINVAL, "wl_camera: null argument");
    ensure_device(sim);
    if (sim->cfg.task != WL_TASK_VISUAL || !sim->cfg.vis_cam) return fail(WL_EUNSUPPORTED, "wl_camera: the handle has no camera term");
    return launch_camera(sim, d_obs, (uint32_t)step_counter, RNG_CAM, 0u, d_aug, (cudaStream_t)stream);
}

int wl_curriculum(wl_sim* sim, int32_t n_terms, const int32_t* slots, const float* increases, uint32_t fire_mask,
                  void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_curriculum: null handle");
    ensure_device(sim);
    if (n_terms < 0 || n_terms > WL_MAX_REW_TERMS) return fail(WL_EINVAL, "wl_curriculum: n_terms");
    if (n_terms == 0 || fire_mask == 0) return WL_OK;
    CurrArgs a; memset(&a, 0, sizeof a);
    a.n = n_terms; a.fire_mask = fire_mask;
    if (sim->last_t < 0) return WL_OK;                       // no step yet: no env has reset
    a.wslot = (uint32_t)sim->last_t & 1u; a.row = (uint32_t)sim->last_t % 3u;
    for (int t = 0; t < n_terms; ++t) {
        if (slots[t] < 0 || slots[t] >= WL_MAX_REW_TERMS) return fail(WL_EINVAL, "wl_curriculum: slot out of range");
        a.slots[t] = slots[t]; a.inc[t] = increases[t];
    }
    wl_curriculum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sim->globals, a);
    WL_LAUNCH_CHECK(sim, "wl_curriculum_kernel");
    return WL_OK;
}

int wl_synth_actions(wl_sim* sim, float* d_action, int64_t step_counter, int32_t dist, void* stream) {
    if (!sim || !d_action) return fail(WL_EINVAL, "wl_synth_actions: null argument");
    ensure_device(sim);
    const int n = sim->cfg.num_envs, bs = 128;
    wl_synth_actions_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->globals, reinterpret_cast<float2*>(d_action),
                                                                               (uint32_t)step_counter, dist);
    WL_LAUNCH_CHECK(sim, "wl_synth_actions_kernel");
    return WL_OK;
}

int wl_derive_suspension(wl_sim* sim, float* d_susp_pos, float* d_susp_vel, void* stream) {
    if (!sim || !d_susp_pos || !d_susp_vel) return fail(WL_EINVAL, "wl_derive_suspension: null argument");
    ensure_device(sim);
    if (((uintptr_t)d_susp_pos & 15u) || ((uintptr_t)d_susp_vel & 15u)) return fail(WL_EINVAL, "wl_derive_suspension: outputs must be 16-byte aligned");
    const int n = sim->cfg.num_envs, bs = 128;
    Terrain T{sim->hf};
    if (sim->cfg.task == WL_TASK_ELEVATION)
        wl_suspension_kernel<WL_TASK_ELEVATION><<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, T, (float4*)d_susp_pos, (float4*)d_susp_vel);
    else
        wl_suspension_kernel<WL_TASK_DRIFT><<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, T, (float4*)d_susp_pos, (float4*)d_susp_vel);
    WL_LAUNCH_CHECK(sim, "wl_suspension_kernel");
    return WL_OK;
}

int32_t wl_policy_blob_floats(int32_t obs_dim, int32_t offsets[13]) {
    int32_t tmp[13];
    return policy_offsets(obs_dim, offsets ? offsets : tmp);
}

int wl_act_step(wl_sim* sim, const float* d_obs_in, const float* d_policy_blob, wl_policy_out out, float* d_obs, float* d_rew,
                uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream) {
    if (!sim || !d_obs_in || !d_policy_blob || !out.actions || !out.mean || !out.log_prob || !out.value || !d_obs || !d_rew ||
        !d_terminated || !d_truncated)
        return fail(WL_EINVAL, "wl_act_step: null argument");
    if (sim->cfg.task == WL_TASK_ELEVATION) return fail(WL_EUNSUPPORTED, "wl_act_step: blind-observation tasks only (obs_dim <= 16)");
    if (sim->obs_dim > 16) return fail(WL_EUNSUPPORTED, "wl_act_step: obs_dim > 16");
    if (((uintptr_t)d_policy_blob & 15u) || ((uintptr_t)out.actions & 7u) || ((uintptr_t)out.mean & 7u))
        return fail(WL_EINVAL, "wl_act_step: policy blob must be 16-byte aligned, actions/mean 8-byte aligned");
    PolicyOffsets po;
    const int blob_floats = policy_offsets(sim->obs_dim, po.o);
    const int n = sim->cfg.num_envs, grid = (n + WL_ACT_ENVS - 1) / WL_ACT_ENVS + 1;      // + the janitor CTA
    const size_t smem = sizeof(float) * ((size_t)blob_floats + WL_ACT_ENVS * (WL_ACT_XS + WL_ACT_HS));
    Terrain T{sim->hf};
    cudaStream_t cs = (cudaStream_t)stream;
    if (int rc = prep_step(sim, step_counter, 1, cs)) return rc;
    static bool attr_set[64] = {false};                       // the attribute is per device
    if (sim->device >= 0 && sim->device < 64 && !attr_set[sim->device]) {
        cudaFuncSetAttribute(wl_act_step_quad_kernel<WL_TASK_VISUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(wl_act_step_quad_kernel<WL_TASK_DRIFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set[sim->device] = true;
    }
    if (sim->cfg.task == WL_TASK_VISUAL)
        launch_k(wl_act_step_quad_kernel<WL_TASK_VISUAL>, grid, WL_ACT_THREADS, smem, cs, sim->cfg, sim->state, sim->globals, T, d_obs_in, d_policy_blob, blob_floats, po, out, d_obs, d_rew, d_terminated, d_truncated, d_log, (uint32_t)step_counter, sim->obs_dim);
    else
        launch_k(wl_act_step_quad_kernel<WL_TASK_DRIFT>, grid, WL_ACT_THREADS, smem, cs, sim->cfg, sim->state, sim->globals, T, d_obs_in, d_policy_blob, blob_floats, po, out, d_obs, d_rew, d_terminated, d_truncated, d_log, (uint32_t)step_counter, sim->obs_dim);
    WL_LAUNCH_CHECK(sim, "wl_act_step_quad_kernel");
    return post_step(sim, step_counter, cs);
}

int wl_gae(const float* d_rewards, const float* d_values, const float* d_last_values, const uint8_t* d_dones,
           const uint8_t* d_time_outs, float gamma, float lam, float* d_returns, float* d_advantages, int32_t T, int32_t N,
           void* stream) {
    if (!d_rewards || !d_values || !d_last_values || !d_dones || !d_returns || !d_advantages) return fail(WL_EINVAL, "wl_gae: null argument");
    if (T < 1 || N < 1) return fail(WL_EINVAL, "wl_gae: T and N must be >= 1");
    const int S = (T + WL_GAE_SEG - 1) / WL_GAE_SEG;
    if (N <= 65536 && S >= 2 && S <= 16)      // small N: latency-bound -> segment-parallel scan; large N: the streaming scan is bandwidth-bound
        wl_gae_seg_kernel<<<(N + 31) / 32, 32 * S, 0, (cudaStream_t)stream>>>(d_rewards, d_values, d_last_values, d_dones, d_time_outs, gamma,
                                                                              lam, d_returns, d_advantages, T, N);
    else
        wl_gae_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_rewards, d_values, d_last_values, d_dones, d_time_outs, gamma,
                                                                         lam, d_returns, d_advantages, T, N);
    return cuda_check(cudaGetLastError(), "wl_gae_kernel");
}

int wl_dp_adam_step(float* d_param, float* d_m, float* d_v, int32_t n_ranks, const float* const* d_grads, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, int32_t n, void* stream) {
    if (!d_param || !d_m || !d_v || !d_grads || n_ranks < 1 || n_ranks > WL_MAX_PEERS || n < 1 || step < 1)
        return fail(WL_EINVAL, "wl_dp_adam_step: bad argument");
    GradPeers gp; memset(&gp, 0, sizeof gp);
    gp.n = n_ranks;
    for (int r = 0; r < n_ranks; ++r) { if (!d_grads[r]) return fail(WL_EINVAL, "wl_dp_adam_step: null gradient pointer"); gp.g[r] = d_grads[r]; }
    const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
    wl_dp_adam_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_param, d_m, d_v, gp, lr, beta1, beta2, eps, weight_decay, bc1, bc2, n);
    return cuda_check(cudaGetLastError(), "wl_dp_adam_kernel");
}

int wl_test_detmath(int32_t op, const float* d_in, const float* d_in2, float* d_out, int32_t n, void* stream) {
    if (n <= 0) return WL_OK;
    wl_detmath_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(op, d_in, d_in2 ? d_in2 : d_in, d_out, n);
    return cuda_check(cudaGetLastError(), "wl_detmath_kernel");
}
int wl_graph_upload(void* graph_exec, void* stream) {
    if (!graph_exec) return fail(WL_EINVAL, "wl_graph_upload: null graph");
    return cuda_check(cudaGraphUpload((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "cudaGraphUpload");
}
int wl_test_null(int32_t grid, int32_t block, void* stream) {
    wl_null_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(nullptr);
    return cuda_check(cudaGetLastError(), "wl_null_kernel");
}
int wl_test_null_cfg(wl_sim* sim, int32_t grid, int32_t block, void* stream) {      /* same, carrying the 1.5 KB wl_config parameter */
    if (!sim) return fail(WL_EINVAL, "wl_test_null_cfg: null handle");
    wl_null_cfg_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(sim->cfg, nullptr);
    return cuda_check(cudaGetLastError(), "wl_null_cfg_kernel");
}
int wl_test_philox(uint64_t seed, uint32_t c0_base, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* d_out, int32_t n,
                   void* stream) {
    if (n <= 0) return WL_OK;
    wl_philox_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, c0_base, c1, c2, c3, reinterpret_cast<uint4*>(d_out), n);
    return cuda_check(cudaGetLastError(), "wl_philox_kernel");
}

}  // extern "C"
