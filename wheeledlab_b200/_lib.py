"""ctypes binding of the C-ABI in ``include/wheeledlab_b200.h``.

The CUDA library is the ONLY compute path of this package: if it is missing, importing
this module raises (there is no CPU / eager fallback by design).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["WHEELEDLAB_B200_LIB"]) if os.environ.get("WHEELEDLAB_B200_LIB") else _PKG_DIR / "libwheeledlab_b200.so"   # (override: kernel A/B builds)

_TAGS = {"i32": C.c_int32, "u64": C.c_uint64, "f32": C.c_float}


class NativeLibraryMissing(ImportError):
    pass


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found. wheeledlab_b200 has no CPU fallback: build the sm_100a extension first "
            f"(python -c 'import __graft_entry__ as g; g.build()' from the repo root)."
        )
    return C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)


lib = _load()

lib.wl_config_describe.restype = C.c_char_p
lib.wl_config_sizeof.restype = C.c_size_t
lib.wl_last_error.restype = C.c_char_p
lib.wl_build_info.restype = C.c_char_p


def make_config_struct(describe: str, name: str = "WlConfig"):
    """Build a ctypes.Structure from a ``wl_config_describe()`` string and verify offsets."""
    fields, offsets, size = [], {}, None
    for item in describe.split(";"):
        if not item:
            continue
        parts = item.split(":")
        if parts[0] == "sizeof":
            size = int(parts[1])
            continue
        fname, tag, count, off = parts[0], parts[1], int(parts[2]), int(parts[3])
        ctype = _TAGS[tag]
        fields.append((fname, ctype if count == 1 else ctype * count))
        offsets[fname] = off
    cls = type(name, (C.Structure,), {"_fields_": fields})
    for fname, off in offsets.items():
        got = getattr(cls, fname).offset
        if got != off:
            raise RuntimeError(f"wl_config layout mismatch at {fname}: ctypes {got} vs C {off}")
    if size is not None and C.sizeof(cls) != size:
        raise RuntimeError(f"wl_config size mismatch: ctypes {C.sizeof(cls)} vs C {size}")
    return cls


CONFIG_DESCRIBE = lib.wl_config_describe().decode()
WlConfig = make_config_struct(CONFIG_DESCRIBE)


class WlGlobals(C.Structure):
    _fields_ = [
        ("rew_weight", (C.c_float * 8) * 2),
        ("acc", (C.c_float * 16) * 3),
        ("log_ptr", C.c_void_p * 3),
        ("last_log", C.c_float * 16),
        ("step_base", C.c_uint32),
        ("ticket", C.c_uint32),
        ("curr_applied_t", C.c_uint32),
        ("_pad", C.c_uint32 * 1),
    ]


_vp, _i32, _i64, _u32, _u64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_size_t
lib.wl_config_finalize.argtypes = [C.POINTER(WlConfig)]
lib.wl_set_step_counter.argtypes = [_vp, _i64, _vp]
lib.wl_set_seed.argtypes = [_vp, _u64]
lib.wl_set_term_bits.argtypes = [_vp, _vp]
lib.wl_set_peer_fanout.argtypes = [_vp, _i32, C.POINTER(_i64)]
lib.wl_set_multicast_fanout.argtypes = [_vp, C.c_int64]
lib.wl_advance_counter.argtypes = [_vp, _i32, _vp]
lib.wl_note_device_counter.argtypes = [_vp, _i64]
lib.wl_log_flush.argtypes = [_vp, _vp]
lib.wl_reward_weights.restype = _vp
lib.wl_reward_weights.argtypes = [_vp]
lib.wl_state_bytes.restype = _sz
lib.wl_state_bytes.argtypes = [_i32]
lib.wl_globals_offset.restype = _sz
lib.wl_globals_offset.argtypes = [_i32]
lib.wl_create.argtypes = [C.POINTER(WlConfig), _vp, _sz, _vp, C.POINTER(_vp)]
lib.wl_destroy.argtypes = [_vp]
lib.wl_startup.argtypes = [_vp, _vp]
lib.wl_reset.argtypes = [_vp, _vp, _i32, _i64, _vp]
lib.wl_step.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_observe.argtypes = [_vp, _vp, _i64, _i32, _vp]
lib.wl_rollout.argtypes = [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_result_bytes.restype = _sz
lib.wl_result_bytes.argtypes = [_i32]
lib.wl_step_host.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_step_host_zero_copy.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_curriculum.argtypes = [_vp, _i32, C.POINTER(_i32), C.POINTER(C.c_float), _u32, _vp]
lib.wl_synth_actions.argtypes = [_vp, _vp, _i64, _i32, _vp]
lib.wl_derive_suspension.argtypes = [_vp, _vp, _vp, _vp]
lib.wl_set_kernel_variant.argtypes = [_vp, _i32]
lib.wl_set_scan_tma.argtypes = [_vp, _i32]
lib.wl_obs_dim.restype = _i32
lib.wl_obs_dim.argtypes = [_vp]
lib.wl_launch_count.restype = _i64
lib.wl_launch_count.argtypes = [_vp]


class WlPolicyOut(C.Structure):
    _fields_ = [("actions", _vp), ("mean", _vp), ("log_prob", _vp), ("value", _vp)]


lib.wl_policy_blob_floats.restype = _i32
lib.wl_policy_blob_floats.argtypes = [_i32, C.POINTER(_i32)]
lib.wl_act_step.argtypes = [_vp, _vp, _vp, WlPolicyOut, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_camera.argtypes = [_vp, _vp, _i64, _vp, _vp]
lib.wl_step_stage_a.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_step_stage_b.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]
lib.wl_gae.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _i32, _i32, _vp]
lib.wl_dp_adam_step.argtypes = [_vp, _vp, _vp, _i32, C.POINTER(_vp), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i32, _i32, _vp]
lib.wl_test_detmath.argtypes = [_i32, _vp, _vp, _vp, _i32, _vp]
lib.wl_test_philox.argtypes = [_u64, _u32, _u32, _u32, _u32, _vp, _i32, _vp]
lib.wl_test_null.argtypes = [_i32, _i32, _vp]
lib.wl_test_null_cfg.argtypes = [_vp, _i32, _i32, _vp]
lib.wl_graph_upload.argtypes = [_vp, _vp]

EXPORTED_SYMBOLS = [
    "wl_config_describe", "wl_config_sizeof", "wl_config_finalize", "wl_set_step_counter", "wl_advance_counter", "wl_note_device_counter", "wl_log_flush", "wl_reward_weights", "wl_set_seed", "wl_set_term_bits", "wl_set_peer_fanout", "wl_set_multicast_fanout", "wl_state_bytes", "wl_globals_offset", "wl_create", "wl_destroy",
    "wl_last_error", "wl_build_info", "wl_startup", "wl_reset", "wl_step", "wl_step_host", "wl_step_host_zero_copy", "wl_rollout", "wl_result_bytes", "wl_observe", "wl_curriculum",
    "wl_synth_actions", "wl_derive_suspension", "wl_set_kernel_variant", "wl_set_scan_tma", "wl_obs_dim", "wl_launch_count", "wl_policy_blob_floats", "wl_act_step", "wl_step_stage_a", "wl_step_stage_b", "wl_camera", "wl_gae", "wl_dp_adam_step", "wl_graph_upload", "wl_test_detmath", "wl_test_philox", "wl_test_null", "wl_test_null_cfg",
]


class WlError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise WlError(f"{what} failed (code {rc}): {lib.wl_last_error().decode()}")
