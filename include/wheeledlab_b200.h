/*
 * wheeledlab_b200.h -- C-ABI of the B200-native vectorised wheeled-robot step.
 *
 * This is the drop-in boundary for ONE path of UWRobotLearning/WheeledLab: the
 * object `gym.make(<id>, cfg=env_cfg)` returns (isaaclab.envs:ManagerBasedRLEnv,
 * reference source/wheeledlab_tasks/wheeledlab_tasks/__init__.py:14-63) and its
 * `step()/reset()/observation_manager.compute()` calls (call stack SURVEY.md 3.3).
 * The reference has no FFI of its own (pure Python down to omni.physx), so the
 * entry points below are the ones a Python binding (ctypes)
 * needs to implement that surface:
 *
 *   wl_step      <- ManagerBasedRLEnv.step(action)           (train loop:
 *                   source/wheeledlab_rl/wheeledlab_rl/utils/modified_rsl_rl_runner.py:73)
 *   wl_reset     <- ManagerBasedRLEnv.reset() / _reset_idx   (wheeledlab_tasks/test/create_and_step_env.py:31)
 *   wl_observe   <- observation_manager.compute()            (modified_rsl_rl_runner.py:50)
 *   wl_startup   <- event_manager.apply(mode="startup")      (drifting/mushr_drift_env_cfg.py:95-154)
 *   wl_curriculum<- curriculum_manager.compute()             (wheeledlab/envs/mdp/curriculums.py:10-35)
 *   wl_synth_actions <- sample_space(single_action_space)    (create_and_step_env.py:37-39)
 *
 * Plain pointers and sizes only; every device pointer is caller-owned (the Python
 * host allocates with torch and passes data_ptr()); `stream` is a cudaStream_t
 * passed as void*. All functions return 0 on success, a negative WL_E* code on
 * failure, and wl_last_error() gives the message.  Nothing here falls back to CPU.
 */
#ifndef WHEELEDLAB_B200_H
#define WHEELEDLAB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WL_ABI_VERSION 2

/* tasks (gym ids, wheeledlab_tasks/__init__.py:14-63) */
#define WL_TASK_DRIFT     0  /* Isaac-MushrDriftRL-v0, Isaac-F1TenthDriftRL-v0 */
#define WL_TASK_ELEVATION 1  /* Isaac-MushrElevationRL-v0 */
#define WL_TASK_VISUAL    2  /* Isaac-MushrVisualRL-v0 (physics/reward side + the software camera term, vis_cam) */

/* action term kinds (wheeledlab/envs/mdp/actions/ *.py) */
#define WL_ACT_ACKERMANN 0   /* ackermann_actions.py:150-201 */
#define WL_ACT_RWD       1   /* rc_car_actions.py:12-29 */
#define WL_ACT_4WD       2   /* rc_car_actions.py:36-64 */

/* bounding strategy (ackermann_actions.py:119-133) */
#define WL_BOUND_NONE 0
#define WL_BOUND_CLIP 1
#define WL_BOUND_TANH 2

/* wheel order everywhere: [back_left, back_right, front_left, front_right]
 * (the 4WD action's own order, rc_car_actions.py:62) */
#define WL_BL 0
#define WL_BR 1
#define WL_FL 2
#define WL_FR 3

#define WL_MAX_REW_TERMS  8
#define WL_MAX_REF_POSES  32
#define WL_MAX_BUCKETS    32
#define WL_OBS_DIM_BLIND  14   /* wheeledlab_tasks/common/observations.py:19-56 */
#define WL_SCAN_SIDE      26   /* 26x26 grid, elevation/mushr_elevation_env_cfg.py:132-142 */
#define WL_SCAN_RAYS      676
#define WL_OBS_DIM_ELEV   689  /* 2+3+3+3+2+676, mushr_elevation_env_cfg.py:57-88 */

/* drift reward term slots (drifting/mushr_drift_env_cfg.py:243-299, declaration order) */
#define WL_DR_SIDE_SLIP   0
#define WL_DR_VEL         1
#define WL_DR_PROGRESS    2
#define WL_DR_TLGR        3
#define WL_DR_TURN_ENERGY 4
#define WL_DR_CROSS_TRACK 5
#define WL_DR_TERM_PENS   6
/* elevation reward term slots (mushr_elevation_env_cfg.py:283-305) */
#define WL_ER_GOAL_RATE   0
#define WL_ER_HEIGHT_Z    1
#define WL_ER_FALLING     2
#define WL_ER_TERM_PEN    3
/* visual reward term slots (visual/mushr_visual_env_cfg.py:376-387) */
#define WL_VR_TRAVERSABLE 0
#define WL_VR_FORWARD_VEL 1
#define WL_OBS_DIM_VISUAL 8    /* base_lin_vel(3) base_ang_vel(3) last_action(2); preceded by vis_cam_w*(vis_cam_h-vis_cam_row0)
                                * camera floats when vis_cam != 0 (PolicyCfg order, mushr_visual_env_cfg.py:45-52) */
#define WL_CAM_MAX_PIXELS 4096 /* kept pixels per env (80 x 40 = 3200 in the reference cfg) */

/* error codes */
#define WL_OK          0
#define WL_EINVAL     -1
#define WL_ECUDA      -2
#define WL_EUNSUPPORTED -3

/*
 * wl_config: plain-old-data description of one task instance.  Defined by an
 * X-macro so that wl_config_describe() can hand the exact field list to a
 * foreign-language binding (the ctypes mirror is generated from it).
 * XS(type, tag, name)  scalar;   XA(type, tag, name, n)  fixed array.
 */
#define WL_CONFIG_FIELDS(XS, XA)                                                                   \
    XS(int32_t, i32, abi_version)                                                                  \
    XS(int32_t, i32, task)                                                                         \
    XS(int32_t, i32, num_envs)        /* envs owned by this handle (one GPU) */                    \
    XS(int32_t, i32, env_id_offset)   /* global id of local env 0 (multi-GPU sharding) */          \
    XS(uint64_t, u64, seed)           /* mushr_drift_env_cfg.py:371 */                             \
    /* --- simulation (mushr_drift_env_cfg.py:393-396) --- */                                      \
    XS(float, f32, sim_dt)                                                                         \
    XS(int32_t, i32, decimation)                                                                   \
    XS(int32_t, i32, substeps)        /* integrator sub-steps per sim step (TGS-like) */           \
    XS(int32_t, i32, max_episode_length)                                                           \
    XS(float, f32, episode_length_s)  /* max_episode_length_s, divides the episode log (Appendix B) */ \
    XS(float, f32, gravity)                                                                        \
    /* --- action term (wheeledlab_tasks/common/actions.py) --- */                                 \
    XS(int32_t, i32, action_kind)                                                                  \
    XS(int32_t, i32, bounding)                                                                     \
    XS(int32_t, i32, no_reverse)                                                                   \
    XS(int32_t, i32, _pad0)                                                                        \
    XA(float, f32, act_scale, 2)                                                                   \
    XA(float, f32, act_offset, 2)                                                                  \
    XS(float, f32, base_length)                                                                    \
    XS(float, f32, base_width)                                                                     \
    XS(float, f32, wheel_radius_cfg)  /* action-map radius (0.05), NOT the collider radius (Q2) */ \
    /* --- vehicle --- */                                                      \
    XS(float, f32, mass_nominal)                                                                   \
    XA(float, f32, inertia_nominal, 3)                                                             \
    XA(float, f32, com, 3)            /* COM in root(base_footprint) frame */                      \
    XS(float, f32, hub_x_front)                                                                    \
    XS(float, f32, hub_x_rear)                                                                     \
    XS(float, f32, hub_y)                                                                          \
    XS(float, f32, hub_z)                                                                          \
    XS(float, f32, wheel_radius)      /* collider radius 0.0525 */                                 \
    XS(float, f32, wheel_inertia)                                                                  \
    XS(float, f32, wheel_damping)                                                                  \
    XS(float, f32, susp_k)                                                                         \
    XS(float, f32, susp_c)                                                                         \
    XS(float, f32, susp_travel)                                                                    \
    XS(float, f32, bump_k)                                                                         \
    XS(float, f32, comp_max)          /* compression used for the force is capped (depenetration) */ \
    XS(float, f32, base_link_z)       /* base_link above the root frame (ray-caster parent), A.1 */ \
    /* --- actuators (wheeledlab_assets/hound.py:4-52) --- */                                      \
    XS(float, f32, dc_saturation)                                                                  \
    XS(float, f32, dc_vel_limit)                                                                   \
    XA(float, f32, dc_effort, 4)      /* per wheel; 0 => passive (hound.py:44-51) */               \
    XA(float, f32, dc_damping, 4)     /* nominal kd per wheel (before DR) */                       \
    XS(float, f32, steer_kp)                                                                       \
    XS(float, f32, steer_kd)                                                                       \
    XS(float, f32, steer_inertia)                                                                  \
    XS(float, f32, steer_vel_limit)                                                                \
    XS(float, f32, steer_pos_limit)                                                                \
    /* --- tire / contact --- */                                                                   \
    XS(float, f32, tire_B)                                                                         \
    XS(float, f32, tire_v0)                                                                        \
    XS(float, f32, tire_mx)           /* effective mass for the longitudinal stick clamp */        \
    XS(float, f32, tire_my)                                                                        \
    XS(float, f32, tire_mx_rest)      /* 1/tire_mx - r_w^2/I_w (chassis part; the wheel part follows the wheel-mass DR) */ \
    XS(float, f32, ground_mu_s)       /* mushr_drift_env_cfg.py:45-49 (combine=multiply) */        \
    XS(float, f32, ground_mu_d)                                                                    \
    /* --- startup domain randomisation (mushr_drift_env_cfg.py:95-154) --- */                     \
    XS(int32_t, i32, dr_enable)                                                                    \
    XS(int32_t, i32, dr_num_buckets)                                                               \
    XA(float, f32, dr_bucket_D, WL_MAX_BUCKETS)   /* peak mu of bucket  (mu_s_wheel*mu_s_ground) */ \
    XA(float, f32, dr_bucket_C, WL_MAX_BUCKETS)   /* Pacejka shape s.t. asymptote = mu_d */        \
    XA(float, f32, dr_kd_range, 2)                                                                 \
    XS(int32_t, i32, dr_kd_mask)      /* bit i => wheel i gets randomised damping */               \
    XS(int32_t, i32, _pad1)                                                                        \
    XA(float, f32, dr_mass_add, 2)    /* base-mass range: += U (mode 0) or base_link mass := U (mode 1) */ \
    XS(int32_t, i32, dr_mass_mode)    /* randomize_rigid_body_mass operation: 0 "add" (drift :145-154, elevation :400-407), 1 "abs" on base_link (visual :280-288) */ \
    XS(int32_t, i32, dr_wheel_mass_enable) /* visual :290-299: wheel link masses := U(dr_wheel_mass), inertia rescaled by the mass ratio */ \
    XS(float, f32, dr_base_mass_nominal)   /* base_link mass in the USD (1.0), replaced in mode 1 */       \
    XS(float, f32, wheel_mass_nominal)     /* wheel link mass in the USD (0.1) */                          \
    XA(float, f32, dr_wheel_mass, 2)                                                               \
    /* --- observation noise (common/observations.py:27-45) --- */                                 \
    XS(int32_t, i32, enable_corruption)                                                            \
    XS(int32_t, i32, _pad2)                                                                        \
    XA(float, f32, noise_std, 4)      /* pos, euler, lin vel, ang vel */                           \
    /* --- interval pushes (mushr_drift_env_cfg.py:121-143) --- */                                 \
    XS(int32_t, i32, push_enable)                                                                  \
    XS(int32_t, i32, _pad3)                                                                        \
    XA(float, f32, push_hf_interval, 2)                                                            \
    XA(float, f32, push_hf_range, 3)  /* +-x, +-y, +-yaw */                                        \
    XA(float, f32, push_lf_interval, 2)                                                            \
    XS(float, f32, push_lf_yaw)                                                                    \
    /* --- reset along track (drifting/mdp/events.py:10-133) --- */                                \
    XS(int32_t, i32, num_ref_poses)                                                                \
    XS(float, f32, reset_pos_noise)                                                                \
    XS(float, f32, reset_yaw_noise)                                                                \
    XA(float, f32, ref_poses, 3 * WL_MAX_REF_POSES) /* x, y, yaw_deg */                            \
    /* --- drift terminations / rewards (mushr_drift_env_cfg.py:27-32,160-362) --- */              \
    XS(float, f32, trk_straight)                                                                   \
    XS(float, f32, trk_corner_in)                                                                  \
    XS(float, f32, trk_corner_out)                                                                 \
    XS(float, f32, ctd_track_radius)                                                               \
    XS(float, f32, ctd_offset)                                                                     \
    XS(float, f32, slip_min_thresh)                                                                \
    XS(float, f32, slip_max_thresh)                                                                \
    XS(float, f32, slip_min_vel_x)                                                                 \
    XS(float, f32, vel_speed_target)                                                               \
    XS(float, f32, vel_offset)                                                                     \
    XS(float, f32, tlgr_ang_vel_thresh)                                                            \
    XS(float, f32, energy_straight)                                                                \
    XS(int32_t, i32, term_enable)     /* bit j = termination term j is configured (play cfgs: terminations = None -> 0) */ \
    XS(int32_t, i32, num_rew_terms)                                                                \
    XA(float, f32, rew_weight, WL_MAX_REW_TERMS)  /* initial weights; live copy is on device */    \
    /* --- curriculum: increase_reward_weight_over_time terms (curriculums.py:10-35), evaluated ON DEVICE by the  \
     *     last CTA of every step (no host logic => the step is CUDA-graph replayable) --- */        \
    XS(int32_t, i32, curr_n)                                                                       \
    XS(int32_t, i32, _pad6)                                                                        \
    XA(int32_t, i32, curr_slot, 4)    /* reward slot of term k */                                  \
    XA(int32_t, i32, curr_every, 4)   /* episodes_per_increase */                                  \
    XA(int32_t, i32, curr_max, 4)     /* max_increases (INT32_MAX = inf) */                        \
    XA(float, f32, curr_inc, 4)       /* increase */                                               \
    /* --- elevation task (elevation/mushr_elevation_env_cfg.py) --- */                            \
    XS(int32_t, i32, hf_nx)           /* height-field raster size */                               \
    XS(int32_t, i32, hf_ny)                                                                        \
    XS(int32_t, i32, hf_pitch)        /* row pitch in floats (multiple of 4: TMA needs 16-B strides) */ \
    XS(int32_t, i32, _pad4)                                                                        \
    XS(float, f32, hf_x0)             /* world coord of sample (0,0) */                            \
    XS(float, f32, hf_y0)                                                                          \
    XS(float, f32, hf_cell)                                                                        \
    XS(float, f32, hf_outside_z)      /* physics ground outside the raster (ground plane z=0) */   \
    XS(float, f32, scan_offset)       /* 0.084, :78 */                                             \
    XS(float, f32, scan_plane_init)   /* 0.19,  :79 */                                             \
    XS(float, f32, scan_sensor_dz)    /* +20 m ray start, :135 */                                  \
    XS(float, f32, scan_res)          /* 0.1 */                                                    \
    XS(float, f32, scan_half)         /* 1.25 = (26-1)*res/2, GridPatternCfg size 2.5 */           \
    XS(float, f32, obs_clip)          /* 10, :61-82 */                                             \
    XA(float, f32, cmd_pos_range, 2)  /* +-19 m, :425-435 */                                       \
    XS(float, f32, cmd_resample_s)                                                                 \
    XA(float, f32, elev_reset_xy, 2)  /* +-19 m, :409-419 */                                       \
    XS(float, f32, elev_reset_yaw)    /* 3.14 */                                                   \
    XA(float, f32, elev_reset_vel, 2) /* U(0.1,0.2) on v_x and v_y */                              \
    XS(float, f32, elev_spawn_z)      /* 0.25, :97,147-149 */                                      \
    XS(float, f32, elev_min_height)   /* 0.15, :354-357 */                                         \
    XS(float, f32, elev_stuck_min_vel)                                                             \
    XS(float, f32, elev_stuck_spin)                                                                \
    XS(float, f32, elev_rollover_cos) /* cos(60 deg), :217-222 */                                  \
    XS(float, f32, elev_goal_dist)                                                                 \
    XS(float, f32, elev_fall_vel)                                                                  \
    XS(float, f32, elev_plane_z)      /* 0.19 in higher_elevation, :166-173 */                     \
    /* --- visual task, physics side (visual/mushr_visual_env_cfg.py); the camera term follows --- */ \
    XS(int32_t, i32, vis_rows)        /* traversability map [rows(y), cols(x)], :73-75 */          \
    XS(int32_t, i32, vis_cols)                                                                     \
    XS(int32_t, i32, vis_n_trav)      /* number of traversable cells (spawn candidates) */         \
    XS(int32_t, i32, _pad5)                                                                        \
    XS(float, f32, vis_row_spacing)   /* 0.5, :68-70 */                                            \
    XS(float, f32, vis_col_spacing)                                                                \
    XS(float, f32, vis_width)         /* num_rows*row_spacing = 250, :113-114 */                   \
    XS(float, f32, vis_height)                                                                     \
    XS(float, f32, vis_spawn_z)       /* 0.1, utils/__init__.py:188-202 + InitialPoseCfg */        \
    /* --- visual task, camera term (software pinhole camera over the 2-colour plane) --- */ \
    XS(int32_t, i32, vis_cam)         /* 0 off, 1 camera_data_rgb_flattened, 2 ..._aug (mdp_sensors/observations.py:64-87) */ \
    XS(int32_t, i32, vis_cam_w)       /* 80, :234 */                                                \
    XS(int32_t, i32, vis_cam_h)       /* 60, :233 */                                                \
    XS(int32_t, i32, vis_cam_row0)    /* H // 3: rows [row0, H) are kept, observations.py:67,78 */  \
    XS(float, f32, vis_cam_fx)        /* focal_length * W / horizontal_aperture (pixels), :236-239 */ \
    XS(float, f32, vis_cam_fy)        /* focal_length * H / vertical_aperture */                    \
    XS(float, f32, vis_cam_cx)        /* W / 2 */                                                   \
    XS(float, f32, vis_cam_cy)        /* H / 2 */                                                   \
    XA(float, f32, vis_cam_pos, 3)    /* optical centre in the root frame: camera_link + offset (0.08,0,0), :242 */ \
    XS(float, f32, vis_cam_bg)        /* colour of what is not the plane mesh (sky, black base plane): 0 */ \
    XS(float, f32, vis_mesh_x0)       /* first mesh vertex: -width/2 - row_spacing/2, utils/__init__.py:26-28 */ \
    XS(float, f32, vis_mesh_y0)                                                                    \
    XS(float, f32, vis_mesh_dx)       /* vertex pitch: width / (num_rows - 1) (np.linspace), quirk: != row_spacing */ \
    XS(float, f32, vis_mesh_dy)                                                                    \
    XS(float, f32, vis_aug_brightness) /* ColorJitter(brightness=0.8, contrast=0.2, saturation=0.8, hue=0.5), observations.py:21 */ \
    XS(float, f32, vis_aug_contrast)                                                               \
    XS(float, f32, vis_aug_saturation)                                                             \
    XS(float, f32, vis_aug_hue)                                                                    \
    XA(float, f32, vis_aug_sigma, 2)  /* GaussianBlur(5, sigma=(0.1, 5.0)), observations.py:23 */  \
    /* --- derived constants: filled by wl_config_finalize() (wl_create calls it on its copy); fp32,   \
     *     formed once on the host so the kernels carry no per-step divisions for them --- */        \
    XS(float, f32, d_h)               /* sim_dt / substeps */                                      \
    XS(float, f32, d_inv_h)                                                                        \
    XS(float, f32, d_step_dt)         /* sim_dt * decimation */                                    \
    XS(float, f32, d_hkp)             /* h * steer_kp */                                           \
    XS(float, f32, d_sden)            /* 1 / (J + h kd + h^2 kp) */                                \
    XS(float, f32, d_inv_Iw)                                                                       \
    XS(float, f32, d_hI)              /* h / wheel_inertia */                                      \
    XS(float, f32, d_inv_hf_cell)     /* 1 / hf_cell (0 without a height-field) */                 \
    XS(float, f32, d_fxk)             /* tire_mx / h */                                            \
    XS(float, f32, d_fyk)                                                                          \
    XS(float, f32, d_inv_wheel_radius_cfg)                                                         \
    XS(float, f32, d_inv_dc_vel_limit)                                                             \
    XS(float, f32, d_inv_mass_nominal)                                                             \
    XA(float, f32, d_invI_nominal, 3)                                                            \
    XS(float, f32, d_vis_mesh_inv_dx) /* 1 / vis_mesh_dx (0 when the camera is off) */             \
    XS(float, f32, d_vis_mesh_inv_dy)

typedef struct wl_config {
#define WL_XS(type, tag, name) type name;
#define WL_XA(type, tag, name, n) type name[n];
    WL_CONFIG_FIELDS(WL_XS, WL_XA)
#undef WL_XS
#undef WL_XA
} wl_config;

/*
 * Device state layout.  One caller-owned buffer of wl_state_bytes() bytes holds
 * WL_NUM_GROUPS "groups"; group g is an array float4[num_envs] at byte offset
 * g * num_envs * 16, so env i reads/writes one aligned 128-bit word per group
 * and a warp touches 512 contiguous bytes (coalesced).  Component meaning:
 */
#define WL_G_POS     0  /* root_pos_w x,y,z ; episode_length_buf (int32 bits)          */
#define WL_G_QUAT    1  /* root_quat_w  w,x,y,z                                       */
#define WL_G_LINVEL  2  /* root_lin_vel_w x,y,z ; push_hf time_left                   */
#define WL_G_ANGVEL  3  /* root_ang_vel_w x,y,z ; push_lf time_left                   */
#define WL_G_WHEEL   4  /* wheel spin rate [bl,br,fl,fr] (joint_vel of *_throttle)    */
#define WL_G_STEER   5  /* steer pos L,R ; steer vel L,R                              */
#define WL_G_ACTION  6  /* action_manager.action (2) ; prev_action (2)                */
#define WL_G_SUM0    7  /* reward episode sums, terms 0-3                             */
#define WL_G_SUM1    8  /* reward episode sums, terms 4-7                             */
#define WL_G_PMASS   9  /* total mass, 1/mass, spare, spare              (DR param)   */
#define WL_G_PMU_D  10  /* Pacejka peak D per wheel                      (DR param)   */
#define WL_G_PMU_C  11  /* Pacejka shape C per wheel                     (DR param)   */
#define WL_G_PKD    12  /* DC-motor damping per wheel                    (DR param)   */
#define WL_G_CMD    13  /* elevation: goal x,y (world), heading_w, command time_left  */
#define WL_G_CMDB   14  /* elevation: command in the yaw frame x,y, heading_b, spare  */
#define WL_G_PIW    15  /* 1 / wheel spin inertia per wheel (DR param; read only when dr_wheel_mass_enable) */
#define WL_NUM_GROUPS 16

/* small global (not per-env) device block appended after the groups.  There is NO grid-wide synchronisation in a step:
 * the kernel of step t reads what step t-1 left (weights slot (t-1) & 1, accumulator row (t-1) % 3), finished envs
 * accumulate into row t % 3, and one warp (the "janitor", under the shadow of the state loads) publishes the weights of
 * step t into slot t & 1, turns row (t-1) % 3 into the extras["log"] row of step t-1 at the pointer that step
 * registered, and clears row (t+1) % 3 for the next launch.  Steps must therefore be issued with consecutive counters
 * on one stream; the host side re-arms the block (wl_step does it transparently) when a counter jumps.
 * Curriculum boundaries (counter % max_episode_length == 0): with a host-supplied counter wl_step adds a 1-thread kernel
 * after the boundary step that applies the terms in place, so the host sees the new weights as soon as the step has run
 * (the reference mutates them inside _reset_idx); with the device-resident counter the next step's threads apply them. */
typedef struct wl_globals {
    float rew_weight[2][WL_MAX_REW_TERMS]; /* live reward weights: slot (t & 1) = weights step t used (the curriculum     */
                                           /* mutates them on the way from slot to slot)                               */
    float acc[3][16];                      /* row t % 3, step t: [0..7] sum over reset envs of their episode sums,       */
                                           /* [8] #reset, [9+j] #envs whose termination term j fired                    */
    float* log_ptr[3];                     /* d_log pointer registered by step t (row t % 3); written one launch later   */
    float last_log[16];                    /* the most recent log row of a step that reset >= 1 env: IsaacLab's extras["log"] */
                                           /* persists until the next _reset_idx, so steps without a reset repeat it      */
    uint32_t step_base;                    /* device-resident base of common_step_counter (CUDA-graph replay)            */
    uint32_t ticket;                       /* CTAs finished (wl_rollout only: its K-step launch ends with a last-CTA pass)  */
    uint32_t curr_applied_t;               /* counter value whose curriculum boundary the host path already applied in place  */
    uint32_t _pad[1];
} wl_globals;
#define WL_LOG_FLOATS 16                  /* d_log: [0..7] Episode_Reward means, [8] #reset, [9+j] term counts */
/* termination term order (bit j of the per-env mask; declaration order of the reference cfgs)
 *   drift     (mushr_drift_env_cfg.py:351-362):     0 time_out, 1 out_of_bounds
 *   elevation (mushr_elevation_env_cfg.py:349-376): 0 time_out, 1 cart_out_of_bounds, 2 stuck, 3 rollover, 4 at_goal
 *   visual    (mushr_visual_env_cfg.py:404-409):    0 time_out, 1 out_range */
#define WL_MAX_TERM_TERMS 7

typedef struct wl_sim wl_sim;   /* opaque handle (host memory) */

/* ---- config ---------------------------------------------------------------- */
/* "name:tag:count:offset;..." for every wl_config field, plus "sizeof:<n>". */
const char* wl_config_describe(void);
size_t wl_config_sizeof(void);
/* device-resident counter base (wl_create zeroes it): set it / add K to it (a 1-thread kernel, capturable: the last node
 * of a replayable graph of K steps) / tell the host mirror what the device holds after graph replays the library did
 * not see (pure host call; `value` = base now in device memory). */
int wl_set_step_counter(wl_sim* sim, int64_t value, void* stream);
int wl_advance_counter(wl_sim* sim, int32_t K, void* stream);
int wl_note_device_counter(wl_sim* sim, int64_t value);
/* publish the extras["log"] row of the most recent step now (otherwise the next step's launch does it) */
int wl_log_flush(wl_sim* sim, void* stream);
/* live reward weights (what the next step will use): device pointer to WL_MAX_REW_TERMS floats inside the state buffer */
float* wl_reward_weights(wl_sim* sim);
/* optional per-env output of wl_step / wl_step_host*: one byte of termination-term bits per env (bit j = term j fired this
 * step, order below) -- what TerminationManager.get_term(name) needs; NULL (default) disables it */
int wl_set_term_bits(wl_sim* sim, uint8_t* d_term_bits);
/* Fused rollout-slab fan-out (multi-GPU): every output row wl_step writes (observation, reward, terminated,
 * truncated) is ALSO stored at `pointer + byte_deltas[k]` for k < n_peers -- byte_deltas[k] = base of peer k's symmetric
 * (P2P-mapped, NVLink) gathered buffer minus the base of the local one -- so the learner-facing concat of the rollout is
 * complete when the step kernels are: no separate collective.  n_peers = 0 switches it off.  Drift-family tasks. */
int wl_set_peer_fanout(wl_sim* sim, int32_t n_peers, const int64_t* byte_deltas);
/* NVSwitch multicast form of the fan-out: `mc_byte_delta` = (multicast alias of the symmetric buffer) - (this rank's mapping of
 * it); the 8-envs-per-CTA step kernel (variant 8) then sends every full, aligned output row with ONE multimem.st that the
 * switch replicates into all ranks' buffers (NVLink egress 1x instead of (world-1)x); rows it cannot send that way (partial
 * CTA, unaligned pointers, other kernel variants) still go peer by peer through wl_set_peer_fanout's deltas, so set both.
 * 0 switches it off.  torch: _SymmetricMemory.multicast_ptr. */
int wl_set_multicast_fanout(wl_sim* sim, int64_t mc_byte_delta);
/* ManagerBasedEnv.seed(): re-key the counter-based generator for all later launches (startup draws are not repeated) */
int wl_set_seed(wl_sim* sim, uint64_t seed);
/* fill the d_* derived fields from the primary ones (idempotent). */
int wl_config_finalize(wl_config* cfg);

/* ---- lifetime -------------------------------------------------------------- */
size_t wl_state_bytes(int32_t num_envs);     /* groups + globals, 256-byte padded */
size_t wl_globals_offset(int32_t num_envs);  /* byte offset of wl_globals in the state buffer */
/* d_state: zero-initialised device buffer of wl_state_bytes(cfg->num_envs) bytes.
 * d_heightfield: float[hf_ny*hf_nx] on device (row-major: index iy*hf_nx + ix, 16-byte aligned, hf_nx % 4 == 0)
 * or NULL.  For the elevation task a TMA tensor map over it is built here.
 * Visual task: the same pointer carries the traversability data instead: int32 trav_cells[vis_n_trav] (cell id =
 * row*vis_cols + col, the spawn candidates) followed, at byte offset ((4*vis_n_trav + 15) & ~15), by
 * uint8 map[vis_rows*vis_cols] (1 = traversable). */
int wl_create(const wl_config* cfg, void* d_state, size_t state_bytes, const float* d_heightfield,
              wl_sim** out);
int wl_destroy(wl_sim* sim);
const char* wl_last_error(void);
/* compiled-for architecture string, e.g. "sm_100a" */
const char* wl_build_info(void);

/* ---- hot path -------------------------------------------------------------- */
/* startup events: material buckets / actuator gains / base mass scatter + default joint state
 * + initial interval timers (mushr_drift_env_cfg.py:95-154; EventManager startup). */
int wl_startup(wl_sim* sim, void* stream);
/* reset the listed envs (d_env_ids==NULL => all).  env ids are local int64 indices, as
 * IsaacLab passes them (_reset_idx).  `step_counter` keys the counter-based RNG. */
int wl_reset(wl_sim* sim, const int64_t* d_env_ids, int32_t n_ids, int64_t step_counter, void* stream);
/* one env.step(): action[N,2] f32 -> obs[N,obs_dim] f32, rew[N] f32, terminated[N] u8,
 * truncated[N] u8.  `step_counter` = common_step_counter BEFORE this step (>= 0), or WL_DEVICE_COUNTER_PLUS(k): the
 * counter base kept in device memory plus k -- then the call carries no per-step host value and a captured CUDA graph of
 * K such calls (k = 0..K-1) followed by wl_advance_counter(K) can be replayed.  Auto-resets finished envs (reward
 * belongs to the pre-reset state, obs to the post-reset state).  The cfg's curriculum terms are applied on the device
 * (reference: inside _reset_idx). */
#define WL_DEVICE_COUNTER (-1)
#define WL_DEVICE_COUNTER_PLUS(k) (-1 - (int64_t)(k))
/* d_log: optional float[WL_LOG_FLOATS]: mean over the envs reset in this step of each reward term's episode sum
 * divided by max_episode_length_s (RewardManager.reset -> extras["log"]), then the reset / terminated / time-out
 * counts; a step in which no env reset repeats the row of the last step that did (the reference's extras["log"] is only
 * rebuilt inside _reset_idx).  No extra kernel, no host sync, no grid-wide sync: the row of step t is written by the NEXT launch on the
 * handle (the step after it, or wl_log_flush), i.e. it is valid once that launch has completed. */
int wl_step(wl_sim* sim, const float* d_action, float* d_obs, float* d_rew, uint8_t* d_terminated,
            uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream);
/* K consecutive env.step()s in ONE launch with the state held in registers (synthetic / scripted-action rollouts:
 * wheeledlab_tasks/test/create_and_step_env.py:34-41 without the per-step host round trip).  d_actions: [K,N,2] or NULL
 * (then U[-1,1]^2 is drawn in-kernel from the counter-based generator; d_actions_out, if not NULL, receives them).
 * Outputs are [K,N,...] slabs; d_log is [K, WL_LOG_FLOATS].  Bit-identical to K wl_step calls.  With curriculum terms
 * configured, [step_counter, step_counter+K) must end at or before the next episode boundary of the global counter
 * (the boundary's weight update happens between launches).  Drift / Visual without the camera term only (the Elevation
 * scan and the Visual camera are second kernels). */
int wl_rollout(wl_sim* sim, int32_t K, const float* d_actions, float* d_actions_out, float* d_obs, float* d_rew,
               uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream);
/* env.step() for a HOST-side caller, host buffers in, host buffers out, one call:
 *   H2D  h_action[N,2] (pinned)  ->  d_action
 *   wl_step(...) writing d_obs, and reward / terminated / truncated into ONE device block d_result laid out as
 *        float rew[N] | uint8 terminated[N] | uint8 truncated[N]          (wl_result_bytes(N) bytes)
 *   D2H  d_result -> h_result (pinned, same layout), then cudaStreamSynchronize(stream).
 * Observations stay on the device (the policy lives there); pass h_obs != NULL to copy them back as well. */
size_t wl_result_bytes(int32_t num_envs);
int wl_step_host(wl_sim* sim, const float* h_action, float* d_action, float* d_obs, void* d_result, float* d_log,
                 void* h_result, float* h_obs, int64_t step_counter, void* stream);
/* Same contract, zero-copy transport: h_action / h_result must be PINNED host memory (cudaHostAlloc / torch
 * pin_memory: device-addressable under unified addressing).  The step kernel reads the actions from and writes
 * reward / terminated / truncated to host memory directly over PCIe -- no staging copies, no extra launches -- then
 * the stream is synchronised.  Bytes crossing the bus per step are identical to wl_step_host. */
 /* (h_action may be any pinned [N,2] f32 block of the caller -- it is read in place.) */
int wl_step_host_zero_copy(wl_sim* sim, const float* h_action, float* d_obs, float* d_log, void* h_result,
                           int64_t step_counter, void* stream);
/* observation_manager.compute(): re-samples the noise. call_idx distinguishes
 * repeated calls at the same step_counter. */
int wl_observe(wl_sim* sim, float* d_obs, int64_t step_counter, int32_t call_idx, void* stream);
/* curriculum (curriculums.py:10-35) evaluated on device so that no host sync is needed:
 * for each term t in [0,n): if any env reset during the last step, rew_weight[slot[t]] += inc[t]
 * when the host-evaluated counter conditions in fire_mask bit t hold. */
int wl_curriculum(wl_sim* sim, int32_t n_terms, const int32_t* slots, const float* increases, uint32_t fire_mask,
                  void* stream);
/* synthetic actions: U[-1,1]^2 (dist=0) or clip(N(0,1),-1,1) (dist=1) keyed by
 * (seed, global env id, step_counter); writes d_action[N,2]. */
int wl_synth_actions(wl_sim* sim, float* d_action, int64_t step_counter, int32_t dist, void* stream);
/* derived joint state for the Python articulation view: suspension pos/vel [N,4]x2 */
int wl_derive_suspension(wl_sim* sim, float* d_susp_pos, float* d_susp_vel, void* stream);
/* step-kernel geometry: 0 = auto (by num_envs and task), 1 = one thread per env, 4 = four lanes (one per wheel) per env,
 * 8 = the quad plus an auxiliary warp per 8 envs that takes everything off the dependent chain (Drift family only).
 * Results are bit-identical across variants. */
int wl_set_kernel_variant(wl_sim* sim, int32_t lanes_per_env);
/* height-scan tile staging: 1 = one TMA tile per CTA (default), 2 = TMA producer/consumer pipeline over persistent CTAs,
 * 0 = plain loads (A/B comparison); identical outputs */
int wl_set_scan_tma(wl_sim* sim, int32_t use_tma);
/* observation width for the configured task */
int32_t wl_obs_dim(const wl_sim* sim);
/* number of kernel launches issued through this handle since creation */
int64_t wl_launch_count(const wl_sim* sim);

/* Visual task camera term alone (wl_step / wl_observe launch it themselves when cfg.vis_cam != 0): fills the first
 * vis_cam_w * (vis_cam_h - vis_cam_row0) floats of every observation row from the current state.  d_aug = NULL draws the
 * ColorJitter / GaussianBlur parameters from the generator (one draw per call, like torchvision on a batch); otherwise
 * 9 floats on the device: brightness, contrast, saturation, hue, sigma, order[4] (fn_idx of ColorJitter.get_params). */
int wl_camera(wl_sim* sim, float* d_obs, int64_t step_counter, const float* d_aug, void* stream);

/* ---- env.step() cut in two, for HOST-SIDE (Python) reward / termination terms ----------------------------
 * IsaacLab's managers call user terms between "physics + built-in terms" and "reset" (ManagerBasedRLEnv.step: reward and
 * termination managers run on the post-physics, pre-reset state).  Stage a = sections A-E of wl_step (action, integrator,
 * counters, built-in terminations and rewards): writes the built-in reward [N] and one byte of termination bits per env
 * (bit j = built-in term j, order as listed above) and leaves the pre-reset state in the state buffer.  The caller may now
 * read the state, ADD its own reward terms to d_rew and raise its own termination flags.  Stage b = sections F-I (episode
 * log, auto-reset, commands / interval pushes, observations, last-CTA epilogue) for done = any built-in bit | extras.
 * wl_step_stage_a + wl_step_stage_b with null extras == wl_step, bit for bit. */
int wl_step_stage_a(wl_sim* sim, const float* d_action, float* d_rew, uint8_t* d_term_bits, int64_t step_counter, void* stream);
int wl_step_stage_b(wl_sim* sim, const uint8_t* d_term_bits, const uint8_t* d_extra_terminated, const uint8_t* d_extra_truncated,
                    float* d_obs, uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream);

/* ---- the step before env.step(): the policy ---------------------------------------------------------
 * alg.act(obs) of the rollout loop (modified_rsl_rl_runner.py:72; rsl_rl ActorCritic with the reference's sizes,
 * drifting/config/agents/mushr/rsl_rl_ppo_cfg.py:12-17: actor and critic 64x64 ELU MLPs, Gaussian head with a learned
 * std) FUSED in front of the env step: one launch = value + action mean + sampled action + log-prob + the whole step.
 * Weights are a caller-owned device blob of fp32 in INPUT-MAJOR (transposed) layout, per net:
 *     W1t[ceil4(obs_dim)][64] b1[64]  W2t[64][64] b2[64]  W3t[64][out] b3[out]   (actor: out = 2, critic: out = 1;
 *     rows obs_dim.. of W1t are zero padding)
 * blob = actor net | critic net | std[2]; every block starts 16-byte aligned (wl_policy_blob_floats gives offsets).
 * a = mean + std * z, z ~ N(0,1) from the counter-based generator (stream 9, keyed by global env id and step). */
typedef struct wl_policy_out {
    float* actions;      /* [N,2] sampled actions (also what the env step consumes) */
    float* mean;         /* [N,2] */
    float* log_prob;     /* [N]   sum over action dims */
    float* value;        /* [N]   */
} wl_policy_out;
/* offsets (in floats) of the 13 arrays inside the blob for observation width obs_dim; returns the total float count */
int32_t wl_policy_blob_floats(int32_t obs_dim, int32_t offsets[13]);
int wl_act_step(wl_sim* sim, const float* d_obs_in, const float* d_policy_blob, wl_policy_out out, float* d_obs,
                float* d_rew, uint8_t* d_terminated, uint8_t* d_truncated, float* d_log, int64_t step_counter, void* stream);

/* ---- the step after the rollout: returns / advantages over the [T, N] slab ---------------------------
 * rsl_rl RolloutStorage.compute_returns as called at modified_rsl_rl_runner.py:116 [UPSTREAM-RECALL], with the time-out
 * bootstrap of PPO.process_env_step (rewards += gamma * values * time_outs) folded in when d_time_outs != NULL:
 *   delta_t = r_t + (1 - done_t) * gamma * V_{t+1} - V_t ;  A_t = delta_t + (1 - done_t) * gamma * lam * A_{t+1}
 *   returns_t = A_t + V_t ; advantages_t = A_t          (V_T = d_last_values; no normalisation here)
 * All arrays are [T, N] row-major (env fastest), exactly the rollout-slab layout; one thread per env walks t = T-1..0. */
int wl_gae(const float* d_rewards, const float* d_values, const float* d_last_values, const uint8_t* d_dones,
           const uint8_t* d_time_outs, float gamma, float lam, float* d_returns, float* d_advantages, int32_t T,
           int32_t N, void* stream);

/* ---- the learner's data-parallel step for the small policy networks --------------------------------------
 * Gradient all-reduce fused with the Adam update, one kernel over peer memory: d_grads[r] = rank r's flat fp32 gradient
 * (this rank's own buffer and the P2P-mapped buffers of its peers, e.g. slices of one torch symmetric-memory allocation; the
 * same rank order on every rank, so every replica forms the identical mean); d_param / d_m / d_v = this rank's flat
 * parameters and Adam moments, n floats each.  `step` = 1, 2, ... (bias correction).  The caller brackets the launch with
 * barriers (all gradients written before; none overwritten until every rank has read them).  n_ranks = 1 is plain Adam. */
int wl_dp_adam_step(float* d_param, float* d_m, float* d_v, int32_t n_ranks, const float* const* d_grads, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int32_t step, int32_t n, void* stream);

/* A caller that drives wl_step through a CUDA graph of K steps (how a rollout is meant to be issued: one launch per K steps)
 * pays the upload of the instantiated graph to the device on its FIRST launch (~10 us for 20 kernel nodes); this moves that
 * upload to set-up time.  graph_exec = cudaGraphExec_t (torch: CUDAGraph.raw_cuda_graph_exec()). */
int wl_graph_upload(void* graph_exec, void* stream);

/* ---- test hooks (bit-exactness of the deterministic math vs the oracle) ------ */
/* op: 0 sin,1 cos,2 atan,3 atan2(x=in,y=in2),4 log,5 tan,6 asin,7 exp,8 tanh ; out[n] */
int wl_test_detmath(int32_t op, const float* d_in, const float* d_in2, float* d_out, int32_t n,
                    void* stream);
/* an empty kernel of the given geometry (launch / event-timing floor of the measurement protocol) */
int wl_test_null(int32_t grid, int32_t block, void* stream);
int wl_test_null_cfg(wl_sim* sim, int32_t grid, int32_t block, void* stream);   /* same, with the wl_config kernel parameter */
/* philox4x32-10: out[4*n] for counters (c0_base + i, c1, c2, c3), key from seed */
int wl_test_philox(uint64_t seed, uint32_t c0_base, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* d_out,
                   int32_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WHEELEDLAB_B200_H */
