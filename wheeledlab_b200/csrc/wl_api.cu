// wl_api.cu -- kernels + the extern "C" ABI declared in include/wheeledlab_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 (see build.py).
#include <cstdio>
#include <cmath>
#include <cstring>
#include <new>
#include <string>

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
    int obs_dim;
    int variant;            // 0 auto, 1 thread-per-env, 4 quad-per-env
};

// below this many envs one thread/env cannot fill 148 SMs x 4 schedulers; use 4 lanes per env
#define WL_QUAD_MAX_ENVS (148 * 4 * 32 * 2)

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return fail(WL_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static inline size_t groups_bytes(int n) { return (size_t)WL_NUM_GROUPS * (size_t)n * 16u; }
static inline size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-step episode log: warp-shuffle reduce over the finished envs, one atomic set per warp.
__device__ __forceinline__ void log_accumulate(wl_globals* __restrict__ gl, bool contrib, bool terminated, bool time_out,
                                               const float sums[WL_MAX_REW_TERMS]) {
    const unsigned any_c = __ballot_sync(0xffffffffu, contrib);
    if (!any_c) return;
    float vals[WL_MAX_REW_TERMS + 3];
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS; ++k) vals[k] = contrib ? sums[k] : 0.0f;
    vals[WL_MAX_REW_TERMS + 0] = contrib ? 1.0f : 0.0f;
    vals[WL_MAX_REW_TERMS + 1] = (contrib && terminated) ? 1.0f : 0.0f;
    vals[WL_MAX_REW_TERMS + 2] = (contrib && time_out) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS + 3; ++k) vals[k] = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < WL_MAX_REW_TERMS + 3; ++k) atomicAdd(&gl->acc[k], vals[k]);
    }
}
// Last CTA of the launch turns the accumulators into the extras["log"] row and re-arms them.
__device__ __forceinline__ void log_finalize(const wl_config& c, wl_globals* __restrict__ gl, float* __restrict__ d_log) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned tk = atomicAdd(&gl->ticket, 1u);
    if (tk != gridDim.x - 1) return;
    __threadfence();
    float a[WL_MAX_REW_TERMS + 3];
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS + 3; ++k) a[k] = __ldcg(&gl->acc[k]);
    const float cnt = a[WL_MAX_REW_TERMS];
    if (d_log != nullptr) {
        const float denom = r_max(cnt, 1.0f) * c.episode_length_s;
#pragma unroll
        for (int k = 0; k < WL_MAX_REW_TERMS; ++k) d_log[k] = a[k] / denom;
        d_log[8] = a[8]; d_log[9] = a[9]; d_log[10] = a[10];
#pragma unroll
        for (int k = 11; k < WL_LOG_FLOATS; ++k) d_log[k] = 0.0f;
    }
    gl->any_reset_last = (cnt > 0.0f) ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) gl->acc[k] = 0.0f;
    gl->ticket = 0u;
}

// One thread per env.  TASK selects the MDP + terrain at compile time.
template <int TASK>
__global__ void __launch_bounds__(128, 4)
wl_step_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
               const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ rew,
               uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool done = false, terminated = false, time_out = false;
    EnvState e;
    if (i < n) {
        load_env(st, n, i, e, TASK == WL_TASK_ELEVATION);
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
        StepConsts kc = make_step_consts(c, e.mass, e.inv_mass);
        for (int d = 0; d < c.decimation; ++d) {
            float tau[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) tau[w] = dc_motor(c, e.kd[w], c.dc_effort[w], wheel_target[w], e.omega[w]);
            for (int j = 0; j < c.substeps; ++j) physics_substep<TASK, 1>(c, T, e, b, tau, steer_target, kc);
        }
        R = rotmat(b.qw, b.qx, b.qy, b.qz);
        cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
        e.p = V3{b.pc.x - cw.x, b.pc.y - cw.y, b.pc.z - cw.z};
        e.v = b.v; e.qw = b.qw; e.qx = b.qx; e.qy = b.qy; e.qz = b.qz;
        e.w = rot(R, b.wb);
        // C. counters
        e.ep_len += 1;
        // D. terminations
        time_out = e.ep_len >= c.max_episode_length;
        const float step_dt = c.d_step_dt;
        float f[WL_MAX_REW_TERMS];
        V3 vb = rotT(R, e.v);
        if (TASK == WL_TASK_DRIFT) {
            terminated = drift_off_track(c, e.p.x, e.p.y);
            drift_reward_terms(c, e.steer[0], e.steer[1], det_atan2(vb.y, vb.x), e.p, vb, b.wb, e.w.z, terminated, time_out, f);
        }
        // E. rewards: value = f*w*dt, skipped when w == 0
        float total = 0.0f;
#pragma unroll
        for (int k = 0; k < WL_MAX_REW_TERMS; ++k) {
            if (k < c.num_rew_terms) {
                float w = __ldg(&gl->rew_weight[k]);
                if (w != 0.0f) { float val = f[k] * w * step_dt; total += val; e.sums[k] += val; }
            }
        }
        rew[i] = total;
        terminated_o[i] = terminated ? 1 : 0;
        truncated_o[i] = time_out ? 1 : 0;
        done = terminated || time_out;
    }
    // F. auto-reset + per-step episode log (warp-shuffle reduction over the finished envs)
    log_accumulate(gl, done, terminated, time_out, e.sums);
    if (i < n) {
        const uint32_t gid = (uint32_t)(c.env_id_offset + i);
        if (done) {
            if (TASK == WL_TASK_DRIFT) drift_reset_env(c, e, gid, t);
        }
        // H. interval events on the post-reset state
        interval_pushes(c, e, gid, t, c.d_step_dt);
        // I. observations
        if (TASK == WL_TASK_DRIFT) blind_obs(c, e, gid, t, RNG_OBS, 0u, obs + (size_t)WL_OBS_DIM_BLIND * i);
        store_env(st, n, i, e, TASK == WL_TASK_ELEVATION);
    }
    log_finalize(c, gl, d_log);
}

// Four lanes per env (lane = wheel).  Same arithmetic, same results; the per-wheel work runs in parallel and the
// chassis is integrated redundantly in the 4 lanes.  Used when N is too small to fill the chip with one thread/env.
template <int TASK>
__global__ void __launch_bounds__(128)
wl_step_quad_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, wl_globals* __restrict__ gl, Terrain T,
                    const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ rew,
                    uint8_t* __restrict__ terminated_o, uint8_t* __restrict__ truncated_o, float* __restrict__ d_log, uint32_t t) {
    const int n = c.num_envs;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 2, w = tid & 3;
    const bool live = i < n;                 // a whole quad is live or not (blockDim is a multiple of 4)
    const int ii = live ? i : n - 1;         // dead quads shadow the last env (no stores) so shuffles stay convergent
    const uint32_t gid = (uint32_t)(c.env_id_offset + ii);
    const unsigned base = (threadIdx.x & 31u) & ~3u;
    EnvState e;
    load_env_quad(st, n, ii, w, e, TASK == WL_TASK_ELEVATION);
    // A. action manager (redundant in the 4 lanes)
    float2 a = action[ii];
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
    StepConsts kc = make_step_consts(c, e.mass, e.inv_mass);
    for (int d = 0; d < c.decimation; ++d) {
        float tau[4];
        tau[0] = dc_motor(c, e.kd[0], my_effort, my_target, e.omega[0]);
        for (int j = 0; j < c.substeps; ++j) physics_substep<TASK, 4>(c, T, e, b, tau, my_steer_target, kc);
    }
    R = rotmat(b.qw, b.qx, b.qy, b.qz);
    cw = rot(R, V3{c.com[0], c.com[1], c.com[2]});
    e.p = V3{b.pc.x - cw.x, b.pc.y - cw.y, b.pc.z - cw.z};
    e.v = b.v; e.qw = b.qw; e.qx = b.qx; e.qy = b.qy; e.qz = b.qz;
    e.w = rot(R, b.wb);
    // C./D. counters, terminations (redundant)
    e.ep_len += 1;
    const bool time_out = e.ep_len >= c.max_episode_length;
    const float step_dt = c.d_step_dt;
    V3 vb = rotT(R, e.v);
    bool terminated = false;
    float f[WL_MAX_REW_TERMS];
    const float steer_l = __shfl_sync(0xffffffffu, e.steer[0], base + 2), steer_r = __shfl_sync(0xffffffffu, e.steer[0], base + 3);
    if (TASK == WL_TASK_DRIFT) {
        terminated = drift_off_track(c, e.p.x, e.p.y);
        drift_reward_terms(c, steer_l, steer_r, det_atan2(vb.y, vb.x), e.p, vb, b.wb, e.w.z, terminated, time_out, f);
    }
    float total = 0.0f;
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS; ++k) {
        if (k < c.num_rew_terms) {
            float wgt = __ldg(&gl->rew_weight[k]);
            if (wgt != 0.0f) { float val = f[k] * wgt * step_dt; total += val; e.sums[k] += val; }
        }
    }
    const bool done = terminated || time_out;
    if (live && w == 0) { rew[i] = total; terminated_o[i] = terminated ? 1 : 0; truncated_o[i] = time_out ? 1 : 0; }
    // F. per-step episode log: one contribution per env (lane 0 of each live quad)
    const bool contrib = done && live && (w == 0);
    log_accumulate(gl, contrib, terminated, time_out, e.sums);
    if (done) {
        if (TASK == WL_TASK_DRIFT) drift_reset_env(c, e, gid, t);      // redundant in the 4 lanes; joints untouched (Q3)
    }
    interval_pushes(c, e, gid, t, step_dt);
    // I. observations: the three euler angles are three atan2 calls -> one per lane
    {
        float qw = e.qw, qx = e.qx, qy = e.qy, qz = e.qz;
        float sin_roll = 2.0f * fm(qw, qx, qy * qz), cos_roll = fm(-2.0f, fm(qx, qx, qy * qy), 1.0f);
        float sin_pitch = 2.0f * fm(qw, qy, -(qz * qx));
        float sin_yaw = 2.0f * fm(qw, qz, qx * qy), cos_yaw = fm(-2.0f, fm(qy, qy, qz * qz), 1.0f);
        float ay = (w == 0) ? sin_roll : (w == 1) ? sin_pitch : sin_yaw;
        float ax = (w == 0) ? cos_roll : (w == 1) ? sqrtf((1.0f - sin_pitch) * (1.0f + sin_pitch)) : cos_yaw;
        float ang = det_atan2(ay, ax);
        if (w == 1 && fabsf(sin_pitch) >= 1.0f) ang = (sin_pitch < 0.0f) ? -1.57079632679489661923f : 1.57079632679489661923f;
        float eu_k = wrap_2pi(ang);
        if (TASK == WL_TASK_DRIFT) {
            // blind_obs_quad shuffles: every lane of the warp calls it, dead quads only skip the stores
            float* o = obs + (size_t)WL_OBS_DIM_BLIND * ii;
            blind_obs_quad(c, e, w, eu_k, gid, t, RNG_OBS, 0u, o, live);
        }
    }
    if (live) store_env_quad(st, n, i, w, e, TASK == WL_TASK_ELEVATION);
    log_finalize(c, gl, d_log);
}

__global__ void wl_startup_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t gid = (uint32_t)(c.env_id_offset + i);
    uint4 r0 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 0u);
    uint4 r1 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 1u);
    uint4 r2 = philox4x32(c.seed, gid, 0u, RNG_STARTUP, 2u);
    const uint32_t rb[4] = {r0.x, r0.y, r0.z, r0.w}, rk[4] = {r1.x, r1.y, r1.z, r1.w};
    float D[4], C[4], kd[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t bk = 0;
        if (c.dr_enable && c.dr_num_buckets > 1) bk = __umulhi(rb[w], (uint32_t)c.dr_num_buckets);
        D[w] = c.dr_bucket_D[bk]; C[w] = c.dr_bucket_C[bk];
        kd[w] = c.dc_damping[w];
        if (c.dr_enable && ((c.dr_kd_mask >> w) & 1)) kd[w] = uniform(rk[w], c.dr_kd_range[0], c.dr_kd_range[1]);
    }
    float mass = c.mass_nominal;
    if (c.dr_enable) mass = mass + uniform(r2.x, c.dr_mass_add[0], c.dr_mass_add[1]);
    float inv_mass = 1.0f / mass;
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

__global__ void wl_reset_kernel(const __grid_constant__ wl_config c, float4* __restrict__ st, const int64_t* __restrict__ ids,
                                int n_ids, uint32_t t) {
    const int n = c.num_envs;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ids) return;
    const int i = ids ? (int)ids[k] : k;
    if (i < 0 || i >= n) return;
    EnvState e;
    load_env(st, n, i, e, true);
    drift_reset_env(c, e, (uint32_t)(c.env_id_offset + i), t);
    store_env(st, n, i, e, false);
}

__global__ void wl_observe_kernel(const __grid_constant__ wl_config c, const float4* __restrict__ st, float* __restrict__ obs,
                                  uint32_t t, uint32_t call_idx) {
    const int n = c.num_envs;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvState e;
    load_env(st, n, i, e, false);
    blind_obs(c, e, (uint32_t)(c.env_id_offset + i), t, RNG_OBS_EXTRA, 3u * call_idx, obs + (size_t)WL_OBS_DIM_BLIND * i);
}

struct CurrArgs { int32_t n; int32_t slots[WL_MAX_REW_TERMS]; float inc[WL_MAX_REW_TERMS]; uint32_t fire_mask; };
__global__ void wl_curriculum_kernel(wl_globals* __restrict__ gl, CurrArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!gl->any_reset_last) return;
    for (int t = 0; t < a.n; ++t)
        if ((a.fire_mask >> t) & 1u) gl->rew_weight[a.slots[t]] += a.inc[t];
}

__global__ void wl_synth_actions_kernel(const __grid_constant__ wl_config c, float2* __restrict__ action, uint32_t t, int dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.num_envs) return;
    uint4 r = philox4x32(c.seed, (uint32_t)(c.env_id_offset + i), t, RNG_ACTION, 0u);
    float a0, a1;
    if (dist == 0) { a0 = 2.0f * u01(r.x) - 1.0f; a1 = 2.0f * u01(r.y) - 1.0f; }
    else { float z0, z1; box_muller(r.x, r.y, z0, z1); a0 = r_clamp(z0, -1.0f, 1.0f); a1 = r_clamp(z1, -1.0f, 1.0f); }
    action[i] = make_float2(a0, a1);
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
    }
    out[i] = r;
}
__global__ void wl_philox_kernel(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint4* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = philox4x32(seed, c0 + (uint32_t)i, c1, c2, c3);
}

// ---------------------------------------------------------------------------------------
// launch geometry: spread small N over all 148 SMs, use fatter CTAs once the chip is full
// ---------------------------------------------------------------------------------------
static inline int pick_block(int n) {
    if (n >= 148 * 128 * 4) return 128;
    if (n >= 148 * 64 * 2) return 64;
    return 32;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char* wl_last_error(void) { return g_err.c_str(); }
const char* wl_build_info(void) { return "wheeledlab_b200 abi=1 arch=sm_100a fmad=false"; }
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

int wl_config_finalize(wl_config* c) {
    if (!c) return fail(WL_EINVAL, "wl_config_finalize: null");
    if (c->substeps <= 0 || !(c->sim_dt > 0.0f)) return fail(WL_EINVAL, "wl_config_finalize: bad sim timing");
    c->d_h = c->sim_dt / (float)c->substeps;
    c->d_inv_h = 1.0f / c->d_h;
    c->d_step_dt = c->sim_dt * (float)c->decimation;
    c->d_hkp = c->d_h * c->steer_kp;
    c->d_sden = 1.0f / fmaf(c->d_h, c->d_hkp, fmaf(c->d_h, c->steer_kd, c->steer_inertia));
    c->d_inv_Iw = 1.0f / c->wheel_inertia;
    c->d_fxk = c->tire_mx * c->d_inv_h;
    c->d_fyk = c->tire_my * c->d_inv_h;
    c->d_inv_wheel_radius_cfg = 1.0f / c->wheel_radius_cfg;
    c->d_inv_dc_vel_limit = 1.0f / c->dc_vel_limit;
    c->d_inv_mass_nominal = 1.0f / c->mass_nominal;
    for (int a = 0; a < 3; ++a) c->d_invI_nominal[a] = 1.0f / c->inertia_nominal[a];
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
    if (cfg->task != WL_TASK_DRIFT) return fail(WL_EUNSUPPORTED, "wl_create: task not implemented in this build");
    if (cfg->bounding != WL_BOUND_CLIP && cfg->bounding != WL_BOUND_NONE)
        return fail(WL_EUNSUPPORTED, "wl_create: bounding_strategy 'tanh' is not implemented");
    if (cfg->decimation <= 0 || cfg->substeps <= 0 || !(cfg->sim_dt > 0.0f)) return fail(WL_EINVAL, "wl_create: bad sim timing");
    if (cfg->num_rew_terms < 0 || cfg->num_rew_terms > WL_MAX_REW_TERMS) return fail(WL_EINVAL, "wl_create: num_rew_terms");
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
    s->variant = 0;
    s->obs_dim = (cfg->task == WL_TASK_ELEVATION) ? WL_OBS_DIM_ELEV : WL_OBS_DIM_BLIND;
    // live reward weights
    if (int rc = cuda_check(cudaMemcpy(s->globals->rew_weight, cfg->rew_weight, sizeof(float) * WL_MAX_REW_TERMS,
                                       cudaMemcpyHostToDevice), "upload reward weights")) { delete s; return rc; }
    *out = s;
    return WL_OK;
}

int wl_destroy(wl_sim* sim) { delete sim; return WL_OK; }
int32_t wl_obs_dim(const wl_sim* sim) { return sim ? sim->obs_dim : 0; }
int wl_set_kernel_variant(wl_sim* sim, int32_t lanes_per_env) {
    if (!sim) return fail(WL_EINVAL, "wl_set_kernel_variant: null handle");
    if (lanes_per_env != 0 && lanes_per_env != 1 && lanes_per_env != 4) return fail(WL_EINVAL, "wl_set_kernel_variant: 0, 1 or 4");
    sim->variant = lanes_per_env;
    return WL_OK;
}
int64_t wl_launch_count(const wl_sim* sim) { return sim ? sim->launches : 0; }

#define WL_LAUNCH_CHECK(sim, what)                                                     \
    do {                                                                               \
        (sim)->launches++;                                                             \
        if (int rc_ = cuda_check(cudaGetLastError(), what)) return rc_;                \
    } while (0)

int wl_startup(wl_sim* sim, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_startup: null handle");
    const int n = sim->cfg.num_envs, bs = 128;
    wl_startup_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state);
    WL_LAUNCH_CHECK(sim, "wl_startup_kernel");
    return WL_OK;
}

int wl_reset(wl_sim* sim, const int64_t* d_env_ids, int32_t n_ids, int64_t step_counter, void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_reset: null handle");
    const int n = d_env_ids ? n_ids : sim->cfg.num_envs;
    if (n <= 0) return WL_OK;
    const int bs = 128;
    wl_reset_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, d_env_ids, n, (uint32_t)step_counter);
    WL_LAUNCH_CHECK(sim, "wl_reset_kernel");
    return WL_OK;
}

int wl_step(wl_sim* sim, const float* d_action, float* d_obs, float* d_rew, uint8_t* d_terminated, uint8_t* d_truncated,
            float* d_log, int64_t step_counter, void* stream) {
    if (!sim || !d_action || !d_obs || !d_rew || !d_terminated || !d_truncated) return fail(WL_EINVAL, "wl_step: null argument");
    if (((uintptr_t)d_action & 7u) || ((uintptr_t)d_obs & 7u)) return fail(WL_EINVAL, "wl_step: action/obs must be 8-byte aligned");
    const int n = sim->cfg.num_envs;
    Terrain T{sim->hf};
    const int variant = sim->variant ? sim->variant : ((n <= WL_QUAD_MAX_ENVS) ? 4 : 1);
    if (variant == 4) {
        const int bs = 32, threads = 4 * n;
        wl_step_quad_kernel<WL_TASK_DRIFT><<<(threads + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(
            sim->cfg, sim->state, sim->globals, T, reinterpret_cast<const float2*>(d_action), d_obs, d_rew, d_terminated,
            d_truncated, d_log, (uint32_t)step_counter);
    } else {
        const int bs = pick_block(n);
        wl_step_kernel<WL_TASK_DRIFT><<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(
            sim->cfg, sim->state, sim->globals, T, reinterpret_cast<const float2*>(d_action), d_obs, d_rew, d_terminated,
            d_truncated, d_log, (uint32_t)step_counter);
    }
    WL_LAUNCH_CHECK(sim, "wl_step_kernel");
    return WL_OK;
}

int wl_observe(wl_sim* sim, float* d_obs, int64_t step_counter, int32_t call_idx, void* stream) {
    if (!sim || !d_obs) return fail(WL_EINVAL, "wl_observe: null argument");
    const int n = sim->cfg.num_envs, bs = 128;
    wl_observe_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, sim->state, d_obs, (uint32_t)step_counter,
                                                                         (uint32_t)call_idx);
    WL_LAUNCH_CHECK(sim, "wl_observe_kernel");
    return WL_OK;
}

int wl_curriculum(wl_sim* sim, int32_t n_terms, const int32_t* slots, const float* increases, uint32_t fire_mask,
                  void* stream) {
    if (!sim) return fail(WL_EINVAL, "wl_curriculum: null handle");
    if (n_terms < 0 || n_terms > WL_MAX_REW_TERMS) return fail(WL_EINVAL, "wl_curriculum: n_terms");
    if (n_terms == 0 || fire_mask == 0) return WL_OK;
    CurrArgs a; memset(&a, 0, sizeof a);
    a.n = n_terms; a.fire_mask = fire_mask;
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
    const int n = sim->cfg.num_envs, bs = 128;
    wl_synth_actions_kernel<<<(n + bs - 1) / bs, bs, 0, (cudaStream_t)stream>>>(sim->cfg, reinterpret_cast<float2*>(d_action),
                                                                               (uint32_t)step_counter, dist);
    WL_LAUNCH_CHECK(sim, "wl_synth_actions_kernel");
    return WL_OK;
}

int wl_derive_suspension(wl_sim* sim, float* d_susp_pos, float* d_susp_vel, void* stream) {
    (void)sim; (void)d_susp_pos; (void)d_susp_vel; (void)stream;
    return fail(WL_EUNSUPPORTED, "wl_derive_suspension: not implemented yet");
}

int wl_test_detmath(int32_t op, const float* d_in, const float* d_in2, float* d_out, int32_t n, void* stream) {
    if (n <= 0) return WL_OK;
    wl_detmath_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(op, d_in, d_in2 ? d_in2 : d_in, d_out, n);
    return cuda_check(cudaGetLastError(), "wl_detmath_kernel");
}
int wl_test_philox(uint64_t seed, uint32_t c0_base, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* d_out, int32_t n,
                   void* stream) {
    if (n <= 0) return WL_OK;
    wl_philox_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, c0_base, c1, c2, c3, reinterpret_cast<uint4*>(d_out), n);
    return cuda_check(cudaGetLastError(), "wl_philox_kernel");
}

}  // extern "C"
