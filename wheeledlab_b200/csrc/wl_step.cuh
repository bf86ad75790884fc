// wl_step.cuh -- the per-env fused step: action map -> DC-motor / PD actuators ->
// chassis + 4-wheel integrator with Pacejka-style slip forces -> terminations ->
// rewards -> auto-reset -> interval pushes -> observations (+ Gaussian noise).
// One env's whole step lives in registers; HBM is touched once per group on the way in
// and once on the way out (DESIGN.md "Kernels").
//
// Reference behaviour restated here (file:line under /root/reference/source):
//   action term     wheeledlab/wheeledlab/envs/mdp/actions/ackermann_actions.py:119-145,150-201
//                   wheeledlab/wheeledlab/envs/mdp/actions/rc_car_actions.py:12-29,36-64
//   actuators       wheeledlab_assets/wheeledlab_assets/hound.py:4-52
//   terminations    wheeledlab_tasks/wheeledlab_tasks/drifting/mushr_drift_env_cfg.py:201-217,343-362
//   rewards         .../drifting/mushr_drift_env_cfg.py:160-299
//   reset           .../drifting/mdp/events.py:102-133
//   pushes          .../drifting/mushr_drift_env_cfg.py:121-143
//   observations    wheeledlab_tasks/wheeledlab_tasks/common/observations.py:19-56
#pragma once
#include "wl_device.cuh"

namespace wl {

struct EnvState {
    V3 p; float qw, qx, qy, qz; V3 v; V3 w;
    float omega[4]; float steer[2], steer_vel[2];
    float action[2], prev_action[2];
    int ep_len; float t_hf, t_lf;
    float sums[WL_MAX_REW_TERMS];
    float mass, inv_mass, spare0, spare1;
    float D[4], C[4], kd[4];
    float cmd[4];
};

__device__ __forceinline__ void load_env(const float4* __restrict__ st, int n, int i, EnvState& e, bool with_cmd) {
    float4 g;
    g = ldg4(st, WL_G_POS, n, i); e.p = V3{g.x, g.y, g.z}; e.ep_len = __float_as_int(g.w);
    g = ldg4(st, WL_G_QUAT, n, i); e.qw = g.x; e.qx = g.y; e.qy = g.z; e.qz = g.w;
    g = ldg4(st, WL_G_LINVEL, n, i); e.v = V3{g.x, g.y, g.z}; e.t_hf = g.w;
    g = ldg4(st, WL_G_ANGVEL, n, i); e.w = V3{g.x, g.y, g.z}; e.t_lf = g.w;
    g = ldg4(st, WL_G_WHEEL, n, i); e.omega[0] = g.x; e.omega[1] = g.y; e.omega[2] = g.z; e.omega[3] = g.w;
    g = ldg4(st, WL_G_STEER, n, i); e.steer[0] = g.x; e.steer[1] = g.y; e.steer_vel[0] = g.z; e.steer_vel[1] = g.w;
    g = ldg4(st, WL_G_ACTION, n, i); e.action[0] = g.x; e.action[1] = g.y; e.prev_action[0] = g.z; e.prev_action[1] = g.w;
    g = ldg4(st, WL_G_SUM0, n, i); e.sums[0] = g.x; e.sums[1] = g.y; e.sums[2] = g.z; e.sums[3] = g.w;
    g = ldg4(st, WL_G_SUM1, n, i); e.sums[4] = g.x; e.sums[5] = g.y; e.sums[6] = g.z; e.sums[7] = g.w;
    g = ldg4(st, WL_G_PMASS, n, i); e.mass = g.x; e.inv_mass = g.y; e.spare0 = g.z; e.spare1 = g.w;
    g = ldg4(st, WL_G_PMU_D, n, i); e.D[0] = g.x; e.D[1] = g.y; e.D[2] = g.z; e.D[3] = g.w;
    g = ldg4(st, WL_G_PMU_C, n, i); e.C[0] = g.x; e.C[1] = g.y; e.C[2] = g.z; e.C[3] = g.w;
    g = ldg4(st, WL_G_PKD, n, i); e.kd[0] = g.x; e.kd[1] = g.y; e.kd[2] = g.z; e.kd[3] = g.w;
    if (with_cmd) { g = ldg4(st, WL_G_CMD, n, i); e.cmd[0] = g.x; e.cmd[1] = g.y; e.cmd[2] = g.z; e.cmd[3] = g.w; }
}
// dynamic state only (params are read-only in the step)
__device__ __forceinline__ void store_env(float4* __restrict__ st, int n, int i, const EnvState& e, bool with_cmd) {
    stg4(st, WL_G_POS, n, i, make_float4(e.p.x, e.p.y, e.p.z, __int_as_float(e.ep_len)));
    stg4(st, WL_G_QUAT, n, i, make_float4(e.qw, e.qx, e.qy, e.qz));
    stg4(st, WL_G_LINVEL, n, i, make_float4(e.v.x, e.v.y, e.v.z, e.t_hf));
    stg4(st, WL_G_ANGVEL, n, i, make_float4(e.w.x, e.w.y, e.w.z, e.t_lf));
    stg4(st, WL_G_WHEEL, n, i, make_float4(e.omega[0], e.omega[1], e.omega[2], e.omega[3]));
    stg4(st, WL_G_STEER, n, i, make_float4(e.steer[0], e.steer[1], e.steer_vel[0], e.steer_vel[1]));
    stg4(st, WL_G_ACTION, n, i, make_float4(e.action[0], e.action[1], e.prev_action[0], e.prev_action[1]));
    stg4(st, WL_G_SUM0, n, i, make_float4(e.sums[0], e.sums[1], e.sums[2], e.sums[3]));
    stg4(st, WL_G_SUM1, n, i, make_float4(e.sums[4], e.sums[5], e.sums[6], e.sums[7]));
    if (with_cmd) stg4(st, WL_G_CMD, n, i, make_float4(e.cmd[0], e.cmd[1], e.cmd[2], e.cmd[3]));
}

// ---- A. action term ---------------------------------------------------------------
__device__ __forceinline__ void process_action(const wl_config& c, float a0, float a1, float wheel_target[4], float steer_target[2]) {
    if (c.bounding == WL_BOUND_CLIP) { a0 = r_clamp(a0, -1.0f, 1.0f); a1 = r_clamp(a1, -1.0f, 1.0f); }
    float v = a0 * c.act_scale[0] + c.act_offset[0];
    float delta = a1 * c.act_scale[1] + c.act_offset[1];
    if (c.no_reverse) v = r_max(v, 0.0f);
    float tan_d = det_tan(delta);
    float L = c.base_length, W = c.base_width, r = c.wheel_radius_cfg;
    if (c.action_kind == WL_ACT_RWD) {
        float wt = v / r;
        wheel_target[WL_BL] = wt; wheel_target[WL_BR] = wt; wheel_target[WL_FL] = 0.0f; wheel_target[WL_FR] = 0.0f;
        steer_target[0] = tan_d; steer_target[1] = tan_d;      // quirk Q1
        return;
    }
    float Rt = (tan_d == 0.0f) ? 1.0e6f : L / tan_d;
    float hw = W / 2.0f;
    float Rl = Rt - hw, Rr = Rt + hw;
    float Rrl = sqrtf(Rl * Rl + L * L), Rrr = sqrtf(Rr * Rr + L * L);
    float den = Rt * r;
    wheel_target[WL_FL] = v * fabsf(Rrl / den);
    wheel_target[WL_FR] = v * fabsf(Rrr / den);
    wheel_target[WL_BL] = v * fabsf(Rl / den);
    wheel_target[WL_BR] = v * fabsf(Rr / den);
    if (c.action_kind == WL_ACT_4WD) { steer_target[0] = tan_d; steer_target[1] = tan_d; }
    else { steer_target[0] = det_atan(L / Rl); steer_target[1] = det_atan(L / Rr); }
}

// ---- a7 DC motor -------------------------------------------------------------------
__device__ __forceinline__ float dc_motor(const wl_config& c, float kd, float effort_limit, float target, float omega) {
    if (!(effort_limit > 0.0f)) return 0.0f;
    float tau = kd * (target - omega);
    float ratio = omega / c.dc_vel_limit;
    float max_eff = r_clamp(c.dc_saturation * (1.0f - ratio), 0.0f, effort_limit);
    float min_eff = r_clamp(c.dc_saturation * (-1.0f - ratio), -effort_limit, 0.0f);
    return r_clamp(tau, min_eff, max_eff);
}

// ---- terrain -----------------------------------------------------------------------
struct Terrain { const float* __restrict__ hf; };

template <int TASK>
__device__ __forceinline__ void terrain_at(const wl_config& c, const Terrain& T, float x, float y, float& z, V3& n) {
    if (TASK == WL_TASK_ELEVATION) {
        if (T.hf != nullptr) {
            float inv = 1.0f / c.hf_cell;
            float fx = (x - c.hf_x0) * inv, fy = (y - c.hf_y0) * inv;
            if ((fx >= 0.0f) && (fy >= 0.0f) && (fx <= (float)(c.hf_nx - 1)) && (fy <= (float)(c.hf_ny - 1))) {
                int ix = (int)floorf(fx), iy = (int)floorf(fy);
                if (ix > c.hf_nx - 2) ix = c.hf_nx - 2;
                if (iy > c.hf_ny - 2) iy = c.hf_ny - 2;
                float tx = fx - (float)ix, ty = fy - (float)iy;
                const float* row0 = T.hf + (size_t)iy * c.hf_nx + ix;
                const float* row1 = row0 + c.hf_nx;
                float z00 = __ldg(row0), z10 = __ldg(row0 + 1), z01 = __ldg(row1), z11 = __ldg(row1 + 1);
                float za = z00 + (z10 - z00) * tx, zb = z01 + (z11 - z01) * tx;
                z = za + (zb - za) * ty;
                float gx = ((z10 - z00) + ((z11 - z01) - (z10 - z00)) * ty) * inv;
                float gy = (zb - za) * inv;
                float ninv = 1.0f / sqrtf(gx * gx + gy * gy + 1.0f);
                n = V3{-gx * ninv, -gy * ninv, ninv};
                return;
            }
            z = c.hf_outside_z;
        } else {
            z = 0.0f;
        }
    } else {
        z = 0.0f;
    }
    n = V3{0.0f, 0.0f, 1.0f};
}

// ---- a8 integrator sub-step ----------------------------------------------------------
struct Chassis { V3 pc; float qw, qx, qy, qz; V3 v; V3 wb; };
struct StepConsts { float h, inv_h, sden, inv_Iw; float I[3], invI[3]; };

__device__ __forceinline__ StepConsts make_step_consts(const wl_config& c, const EnvState& e) {
    StepConsts k;
    k.h = c.sim_dt / (float)c.substeps;
    k.inv_h = 1.0f / k.h;
    k.sden = 1.0f / (c.steer_inertia + k.h * c.steer_kd + k.h * k.h * c.steer_kp);
    k.inv_Iw = 1.0f / c.wheel_inertia;
    float ms = e.mass / c.mass_nominal;
#pragma unroll
    for (int a = 0; a < 3; ++a) { k.I[a] = c.inertia_nominal[a] * ms; k.invI[a] = 1.0f / k.I[a]; }
    return k;
}

template <int TASK>
__device__ __forceinline__ void physics_substep(const wl_config& c, const Terrain& T, EnvState& e, Chassis& b,
                                                const float tau[4], const float steer_target[2], const StepConsts& k) {
    const float h = k.h;
    M3 R = rotmat(b.qw, b.qx, b.qy, b.qz);
    float sn[2], cs[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float vel = (c.steer_inertia * e.steer_vel[j] + h * c.steer_kp * (steer_target[j] - e.steer[j])) * k.sden;
        vel = r_clamp(vel, -c.steer_vel_limit, c.steer_vel_limit);
        float pos = r_clamp(e.steer[j] + h * vel, -c.steer_pos_limit, c.steer_pos_limit);
        e.steer_vel[j] = vel; e.steer[j] = pos;
        det_sincos(pos, sn[j], cs[j]);
    }
    V3 vb = rotT(R, b.v);
    V3 Fb{0.0f, 0.0f, 0.0f}, Tb{0.0f, 0.0f, 0.0f};
    const float rw = c.wheel_radius, bw = c.wheel_damping;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        V3 rho{((i >= 2) ? c.hub_x_front : c.hub_x_rear) - c.com[0], ((i & 1) ? -c.hub_y : c.hub_y) - c.com[1], c.hub_z - c.com[2]};
        V3 hubw = rot(R, rho);
        hubw.x += b.pc.x; hubw.y += b.pc.y; hubw.z += b.pc.z;
        float zt; V3 nw; terrain_at<TASK>(c, T, hubw.x, hubw.y, zt, nw);
        float comp = rw - (hubw.z - zt) * nw.z;
        // drive torque first, then friction against the resulting slip (implicit stick)
        float om_star = e.omega[i] + h * ((tau[i] - bw * e.omega[i]) * k.inv_Iw);
        float Fx = 0.0f;
        if (comp > 0.0f) {
            V3 nb = rotT(R, nw);
            V3 rc{rho.x - rw * nb.x, rho.y - rw * nb.y, rho.z - rw * nb.z};
            V3 wxr = cross(b.wb, rc);
            V3 vc{vb.x + wxr.x, vb.y + wxr.y, vb.z + wxr.z};
            float sdot = -(nb.x * vc.x + nb.y * vc.y + nb.z * vc.z);
            float Fz = c.susp_k * comp + c.susp_c * sdot;
            if (comp > c.susp_travel) Fz += c.bump_k * (comp - c.susp_travel);
            Fz = r_max(Fz, 0.0f);
            V3 hb;
            if (i >= 2) hb = V3{cs[i - 2], sn[i - 2], 0.0f}; else hb = V3{1.0f, 0.0f, 0.0f};
            float d = hb.x * nb.x + hb.y * nb.y + hb.z * nb.z;
            V3 ft{hb.x - d * nb.x, hb.y - d * nb.y, hb.z - d * nb.z};
            float finv = 1.0f / sqrtf(ft.x * ft.x + ft.y * ft.y + ft.z * ft.z);
            ft.x *= finv; ft.y *= finv; ft.z *= finv;
            V3 lt = cross(nb, ft);
            float vx = vc.x * ft.x + vc.y * ft.y + vc.z * ft.z;
            float vy = vc.x * lt.x + vc.y * lt.y + vc.z * lt.z;
            float vsx = vx - om_star * rw;
            float invden = 1.0f / r_max(fabsf(vx), c.tire_v0);
            float kappa = -vsx * invden, ta = -vy * invden;
            float sigma = sqrtf(kappa * kappa + ta * ta);
            float Fy = 0.0f;
            if (sigma > 1.0e-9f) {
                float sm, cm; det_sincos(e.C[i] * det_atan(c.tire_B * sigma), sm, cm);
                float Fmag = Fz * (e.D[i] * sm) / sigma;
                Fx = Fmag * kappa; Fy = Fmag * ta;
                float fxm = c.tire_mx * fabsf(vsx) * k.inv_h, fym = c.tire_my * fabsf(vy) * k.inv_h;
                Fx = r_clamp(Fx, -fxm, fxm); Fy = r_clamp(Fy, -fym, fym);
            }
            V3 F{Fz * nb.x + Fx * ft.x + Fy * lt.x, Fz * nb.y + Fx * ft.y + Fy * lt.y, Fz * nb.z + Fx * ft.z + Fy * lt.z};
            V3 Tq = cross(rc, F);
            Fb.x += F.x; Fb.y += F.y; Fb.z += F.z;
            Tb.x += Tq.x; Tb.y += Tq.y; Tb.z += Tq.z;
        }
        e.omega[i] = om_star - h * ((rw * Fx) * k.inv_Iw);
    }
    V3 Fw = rot(R, Fb);
    b.v.x = b.v.x + h * (Fw.x * e.inv_mass);
    b.v.y = b.v.y + h * (Fw.y * e.inv_mass);
    b.v.z = b.v.z + h * (Fw.z * e.inv_mass - c.gravity);
    V3 Iw3{k.I[0] * b.wb.x, k.I[1] * b.wb.y, k.I[2] * b.wb.z};
    V3 g = cross(b.wb, Iw3);
    b.wb.x = b.wb.x + h * ((Tb.x - g.x) * k.invI[0]);
    b.wb.y = b.wb.y + h * ((Tb.y - g.y) * k.invI[1]);
    b.wb.z = b.wb.z + h * ((Tb.z - g.z) * k.invI[2]);
    b.pc.x = b.pc.x + h * b.v.x;
    b.pc.y = b.pc.y + h * b.v.y;
    b.pc.z = b.pc.z + h * b.v.z;
    float hh = 0.5f * h;
    float qw = b.qw, qx = b.qx, qy = b.qy, qz = b.qz, ox = b.wb.x, oy = b.wb.y, oz = b.wb.z;
    float nqw = qw - hh * (qx * ox + qy * oy + qz * oz);
    float nqx = qx + hh * (qw * ox + qy * oz - qz * oy);
    float nqy = qy + hh * (qw * oy + qz * ox - qx * oz);
    float nqz = qz + hh * (qw * oz + qx * oy - qy * ox);
    float qinv = 1.0f / sqrtf(nqw * nqw + nqx * nqx + nqy * nqy + nqz * nqz);
    b.qw = nqw * qinv; b.qx = nqx * qinv; b.qy = nqy * qinv; b.qz = nqz * qinv;
}

// ---- drift terminations / rewards -----------------------------------------------------
__device__ __forceinline__ bool drift_off_track(const wl_config& c, float x, float y) {
    float st = c.trk_straight, ro = c.trk_corner_out, ri = c.trk_corner_in;
    bool off, in;
    if (fabsf(y) < st) { off = fabsf(x) > ro; in = fabsf(x) < ri; }
    else if (y > 0.0f) { float d2 = (y - st) * (y - st) + x * x; off = d2 > ro * ro; in = d2 < ri * ri; }
    else { float d2 = (y + st) * (y + st) + x * x; off = d2 > ro * ro; in = d2 < ri * ri; }
    return off || in;
}

__device__ __forceinline__ void drift_reward_terms(const wl_config& c, const EnvState& e, V3 p, V3 vb, V3 wb, float wz_world,
                                                   bool out_of_bounds, bool time_out, float f[WL_MAX_REW_TERMS]) {
    float slip = fabsf(det_atan2(vb.y, vb.x));
    float valid = (fabsf(vb.x) < c.slip_min_vel_x || slip > c.slip_max_thresh) ? 0.0f : slip;
    if (valid < c.slip_min_thresh) valid = 0.0f;
    f[WL_DR_SIDE_SLIP] = valid;
    float gs = sqrtf(vb.x * vb.x + vb.y * vb.y);
    float dv = gs - c.vel_speed_target;
    f[WL_DR_VEL] = dv * dv + c.vel_offset;
    f[WL_DR_PROGRESS] = wz_world;
    float sm = (e.steer[0] + e.steer[1]) / 2.0f;
    float av = r_clamp(wb.z, -c.tlgr_ang_vel_thresh, c.tlgr_ang_vel_thresh);
    float tl = sm * av * -1.0f;
    f[WL_DR_TLGR] = r_max(tl, 0.0f);
    float sp = sqrtf(vb.x * vb.x + vb.y * vb.y + vb.z * vb.z);
    f[WL_DR_TURN_ENERGY] = (fabsf(p.y) > c.energy_straight) ? sp * sp : 0.0f;
    float st = c.trk_straight, tr = c.ctd_track_radius, sq;
    if (fabsf(p.y) < st) { float d = (p.x > 0.0f) ? (p.x - tr) : (p.x + tr); sq = d * d; }
    else { float yy = (p.y > 0.0f) ? (p.y - st) : (p.y + st); float d = sqrtf(yy * yy + p.x * p.x) - tr; sq = d * d; }
    f[WL_DR_CROSS_TRACK] = sqrtf(sq) + c.ctd_offset;
    f[WL_DR_TERM_PENS] = (out_of_bounds && !time_out) ? 1.0f : 0.0f;
    f[7] = 0.0f;
}

// ---- reset / pushes / observations -----------------------------------------------------
__device__ __forceinline__ void sample_interval_timers(const wl_config& c, EnvState& e, uint32_t a, uint32_t b) {
    e.t_hf = uniform(a, c.push_hf_interval[0], c.push_hf_interval[1]);
    e.t_lf = uniform(b, c.push_lf_interval[0], c.push_lf_interval[1]);
}
__device__ __forceinline__ void drift_reset_env(const wl_config& c, EnvState& e, uint32_t gid, uint32_t t) {
    uint4 r = philox4x32(c.seed, gid, t, RNG_RESET, 0u);
    uint32_t idx = __umulhi(r.x, (uint32_t)c.num_ref_poses);
    float nx = (2.0f * u01(r.y) - 1.0f) * c.reset_pos_noise;
    float ny = (2.0f * u01(r.z) - 1.0f) * c.reset_pos_noise;
    float nyaw = (2.0f * u01(r.w) - 1.0f) * c.reset_yaw_noise;
    e.p.x = c.ref_poses[3 * idx + 0] + nx;
    e.p.y = c.ref_poses[3 * idx + 1] + ny;
    e.p.z = 0.0f;
    float yaw = c.ref_poses[3 * idx + 2] * 0.017453292519943295f + nyaw;
    float sh, ch; det_sincos(yaw * 0.5f, sh, ch);
    e.qw = ch; e.qx = 0.0f; e.qy = 0.0f; e.qz = sh;
    e.v = V3{0.0f, 0.0f, 0.0f}; e.w = V3{0.0f, 0.0f, 0.0f};
    e.ep_len = 0;
#pragma unroll
    for (int k = 0; k < WL_MAX_REW_TERMS; ++k) e.sums[k] = 0.0f;
    e.action[0] = e.action[1] = e.prev_action[0] = e.prev_action[1] = 0.0f;
    uint4 r2 = philox4x32(c.seed, gid, t, RNG_RESET, 1u);
    sample_interval_timers(c, e, r2.x, r2.y);
}
__device__ __forceinline__ void interval_pushes(const wl_config& c, EnvState& e, uint32_t gid, uint32_t t, float step_dt) {
    if (!c.push_enable) return;
    e.t_hf = e.t_hf - step_dt;
    if (e.t_hf < 1.0e-6f) {
        uint4 r = philox4x32(c.seed, gid, t, RNG_PUSH_HF, 0u);
        e.v.x = e.v.x + uniform(r.x, -c.push_hf_range[0], c.push_hf_range[0]);
        e.v.y = e.v.y + uniform(r.y, -c.push_hf_range[1], c.push_hf_range[1]);
        e.w.z = e.w.z + uniform(r.z, -c.push_hf_range[2], c.push_hf_range[2]);
        e.t_hf = uniform(r.w, c.push_hf_interval[0], c.push_hf_interval[1]);
    }
    e.t_lf = e.t_lf - step_dt;
    if (e.t_lf < 1.0e-6f) {
        uint4 r = philox4x32(c.seed, gid, t, RNG_PUSH_LF, 0u);
        e.w.z = e.w.z + uniform(r.x, -c.push_lf_yaw, c.push_lf_yaw);
        e.t_lf = uniform(r.y, c.push_lf_interval[0], c.push_lf_interval[1]);
    }
}
// writes 14 floats (obs must be 8-byte aligned: 14 floats = 7 x float2 per env)
__device__ __forceinline__ void blind_obs(const wl_config& c, const EnvState& e, uint32_t gid, uint32_t t, uint32_t stream,
                                          uint32_t sub0, float* __restrict__ obs) {
    M3 R = rotmat(e.qw, e.qx, e.qy, e.qz);
    V3 vb = rotT(R, e.v), wb = rotT(R, e.w), eu = euler_xyz(e.qw, e.qx, e.qy, e.qz);
    float z[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) z[k] = 0.0f;
    if (c.enable_corruption) {
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k) {
            uint4 r = philox4x32(c.seed, gid, t, stream, sub0 + k);
            box_muller(r.x, r.y, z[4 * k + 0], z[4 * k + 1]);
            box_muller(r.z, r.w, z[4 * k + 2], z[4 * k + 3]);
        }
    }
    float o[14];
    o[0] = e.p.x + c.noise_std[0] * z[0]; o[1] = e.p.y + c.noise_std[0] * z[1]; o[2] = e.p.z + c.noise_std[0] * z[2];
    o[3] = eu.x + c.noise_std[1] * z[3]; o[4] = eu.y + c.noise_std[1] * z[4]; o[5] = eu.z + c.noise_std[1] * z[5];
    o[6] = vb.x + c.noise_std[2] * z[6]; o[7] = vb.y + c.noise_std[2] * z[7]; o[8] = vb.z + c.noise_std[2] * z[8];
    o[9] = wb.x + c.noise_std[3] * z[9]; o[10] = wb.y + c.noise_std[3] * z[10]; o[11] = wb.z + c.noise_std[3] * z[11];
    o[12] = r_clamp(e.action[0], -1.0f, 1.0f); o[13] = r_clamp(e.action[1], -1.0f, 1.0f);
    float2* o2 = reinterpret_cast<float2*>(obs);
#pragma unroll
    for (int k = 0; k < 7; ++k) o2[k] = make_float2(o[2 * k], o[2 * k + 1]);
}

}  // namespace wl
