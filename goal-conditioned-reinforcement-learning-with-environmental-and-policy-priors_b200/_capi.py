"""ctypes binding of the C ABI in include/twoarmy_b200.h.

There is no CPU fallback: if the shared library is missing, cannot be loaded, or lacks a
symbol the header declares, importing the product fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
REPO_ROOT = PKG_DIR.parent
HEADER = REPO_ROOT / "include" / "twoarmy_b200.h"
LIB_PATH = CSRC / "libtwoarmy_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-lpthread",
]

TA_OK = 0
TA_STEP_AUTORESET = 1
TA_STEP_HOST_DMA = 2
ACT_I32, ACT_U8, ACT_I64 = 0, 1, 2
ENV_ERR_BAD_ACTION, ENV_ERR_NONE_POS, ENV_ERR_OOB_MOVE = 1, 2, 4
F_PONE, F_PATROL, F_UP1, F_RIGHT2, F_UPD_H, F_UPD_L, F_FIRST_ROOM2 = 1, 2, 4, 8, 16, 32, 64


class TwoarmyLibraryError(RuntimeError):
    pass


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HEADER]


def build(force: bool = False, verbose: bool = False) -> Path:
    """nvcc-compile the CUDA library in-tree for sm_100a (cross-compiles without a GPU)."""
    newest = max(p.stat().st_mtime for p in sources())
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= newest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "twoarmy_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise TwoarmyLibraryError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function include/twoarmy_b200.h declares."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        try:
            build()
        except Exception as exc:  # no nvcc on this machine
            raise TwoarmyLibraryError(
                f"{LIB_PATH} is missing and could not be built ({exc}); the product has no CPU path") from exc
    elif LIB_PATH.stat().st_mtime < max(p.stat().st_mtime for p in sources()):
        # a source is newer than the library: never run a stale build (rebuild where nvcc exists, else say so)
        try:
            build()
        except Exception as exc:
            import warnings
            warnings.warn(f"{LIB_PATH} is older than its sources and could not be rebuilt ({exc})")
    try:
        L = C.CDLL(str(LIB_PATH))
    except OSError as exc:
        raise TwoarmyLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise TwoarmyLibraryError(f"{LIB_PATH} lacks symbols declared in the header: {missing}")
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
    L.ta_create.argtypes = [C.POINTER(vp), i32, i64, i32, i32, u64, u64]
    L.ta_destroy.argtypes = [vp]
    L.ta_num_envs.argtypes = [vp]; L.ta_num_envs.restype = i64
    L.ta_view.argtypes = [vp]
    L.ta_version.argtypes = [vp]
    L.ta_reset.argtypes = [vp, vp, i32, vp, vp]
    L.ta_observe_general.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.ta_step.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    L.ta_step_host.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]
    L.ta_step_packed.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp]
    L.ta_decode_packed_host.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp]
    L.ta_step_host_d2h_bytes.argtypes = [vp]; L.ta_step_host_d2h_bytes.restype = i64
    L.ta_rollout.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.ta_state_matrix.argtypes = [vp, vp, vp, vp, vp]
    L.ta_stack_roll.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ta_stack_roll_codes.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ta_stack_push.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.ta_export_state.argtypes = [vp, vp, vp]
    L.ta_import_state.argtypes = [vp, vp, vp]
    L.ta_conv1_fwd.argtypes = [vp, i32, i64, vp, vp, i64, vp, vp]
    L.ta_conv1_fwd_add.argtypes = [vp, i32, i64, vp, vp, i64, vp, i32, vp, vp]
    L.ta_conv1_bwd.argtypes = [vp, i32, i64, vp, vp, i64, vp, vp, vp]
    L.ta_parity_class_weights.argtypes = [vp, i64, i64, i64, i64, i32, i32, i32, vp, vp]
    L.ta_planes_to_dense_relu.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, vp]
    L.ta_conv1_bwd_planes.argtypes = [vp, i32, i64, vp, vp, vp, i32, i64, vp, vp, vp]
    L.ta_conv1_fwd_mask.argtypes = [vp, i32, i64, vp, vp, i64, vp, vp, vp]
    L.ta_conv2_dgrad_prep.argtypes = [vp, i64, i64, i64, i64, vp, vp]
    L.ta_conv2_dgrad_planes.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    L.ta_conv2_dgrad_conv1_bwd.argtypes = [vp, vp, vp, vp, i32, i64, i64, vp, vp, vp]
    L.ta_col2im_s2.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp]
    L.ta_im2col_s2.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp]
    L.ta_render.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
    L.ta_channel_sum_bf16.argtypes = [vp, i64, i32, vp, vp]
    L.ta_her_plan.argtypes = [vp, vp, i32, i64, i32, u64, u64, vp, vp, vp, vp, vp]
    L.ta_gae.argtypes = [vp, vp, vp, vp, vp, f32, f32, i32, i32, i64, vp, vp, vp]
    L.ta_gae_stats.argtypes = [vp, vp, vp, vp, vp, f32, f32, i32, i32, i64, vp, vp, vp, vp]
    L.ta_pred_encoder.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ta_pred_decoder.argtypes = [vp, i64, vp, vp, vp, vp, vp, f32, vp, vp]
    L.ta_lstm_gates.argtypes = [vp, vp, vp, vp, vp, i64, i64, i32, vp]
    L.ta_gae_normalized.argtypes = [vp, vp, vp, vp, vp, f32, f32, i32, i32, i64, vp, vp, vp, vp]
    L.ta_adv_stats.argtypes = [vp, i64, vp, vp]
    L.ta_adv_normalize.argtypes = [vp, i64, vp, vp]
    L.ta_relu_bwd_bias_scratch_floats.argtypes = [i64, i32]; L.ta_relu_bwd_bias_scratch_floats.restype = i64
    L.ta_planes_relu_bwd_bias.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]
    L.ta_relu_bwd_bias.argtypes = [vp, i64, vp, vp, i64, i32, vp, vp, vp]
    L.ta_ppo_actor_loss.argtypes = [vp, vp, vp, vp, i32, f32, f32, vp, vp, vp, vp, vp]
    L.ta_ppo_critic_loss.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.ta_adam_shadow.argtypes = [vp, vp, vp, vp, vp, i64, vp, f32, C.c_double, C.c_double, f32, f32, vp]
    L.ta_tinet_prep.argtypes = [vp, vp]
    L.ta_tinet_grad.argtypes = [vp, vp]
    L.ta_gather_minibatch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ta_strerror.argtypes = [i32]; L.ta_strerror.restype = C.c_char_p
    L.ta_last_cuda_error.restype = C.c_char_p
    L.ta_launch_count.restype = i64
    L.ta_set_timing.argtypes = [vp, i32]
    L.ta_last_step_ms.argtypes = [vp, C.POINTER(f32)]
    L.ta_debug_force_generic_obs.argtypes = [i32]
    L.ta_debug_conv1_tc.argtypes = [i32]
    L.ta_debug_conv1_bwd_tc.argtypes = [i32]
    L.ta_debug_push_tma.argtypes = [i32]
    L.ta_debug_conv1_tc_failed.argtypes = []
    if L.ta_abi_version() != 1:
        raise TwoarmyLibraryError("ABI version mismatch between header and library")
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != TA_OK:
        L = lib()
        msg = L.ta_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.ta_last_cuda_error().decode()
        raise TwoarmyLibraryError(f"{what or 'twoarmy_b200'} failed: {msg}")


def launch_count() -> int:
    return int(lib().ta_launch_count())
