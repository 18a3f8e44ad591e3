"""ctypes binding of the C ABI in include/narde_b200.h (device pointers from torch tensors).

This is the only bridge between the Python host side and the CUDA kernels.  There is no CPU
fallback: `load()` raises if libnarde_b200.so has not been built, and every compute wrapper
raises if CUDA is unavailable or a tensor is not a CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnarde_b200.so")
if os.environ.get("NARDE_B200_DEBUG_HOOKS") == "1":   # measurement tools only (tools/*.py): the -DNARDE_DEBUG_HOOKS build
    LIB_PATH = os.path.join(_HERE, "libnarde_b200_debug.so")

ABI_VERSION = 4
# flags / bits (include/narde_b200.h)
REWARD_MOVER12 = 1
AUTORESET = 2
ACTION_FRACTION = 32
ENUMERATE_ONLY = 64
PACK_RESULT = 128
DEVICE_ADVANCE = 256
HALF_MOVES_ONLY = 4
MAX_HALF_MOVES = 96
NUM_STATS = 9
STAT_NAMES = ("episodes", "white_wins", "black_wins", "mars", "episode_steps", "legal_actions",
              "max_actions", "overflows", "clamped_actions")



def workspace_ints(n):
    """NARDE_WORKSPACE_INTS(n), include/narde_b200.h."""
    return int(n) + 8


_vp, _i64, _u64, _i32, _int = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_int

_SIGNATURES = {
    "narde_abi_version": ([], _int),
    "narde_build_arch": ([], C.c_char_p),
    "narde_reset": ([_vp, _vp, _i64, _i64, _u64, _u64, _vp], _int),
    "narde_reset_masked": ([_vp, _vp, _vp, _i64, _i64, _u64, _u64, _vp], _int),
    "narde_half_moves": ([_vp, _vp, _vp, _i64, _int, _vp, _vp, _vp], _int),
    "narde_step_ref": ([_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_enumerate": ([_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp], _int),
    "narde_enumerate_fast": ([_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_step_full": ([_vp, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _vp, _vp, _i32, _i32, _vp, _vp, _vp], _int),
    "narde_step_full_mirror": ([_vp, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_advance_counter": ([_vp, _vp], _int),
    "narde_mlp_forward": ([_vp, _i64, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_score": ([_vp, _i64, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_forward_states": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_score_states": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_score_states_2sm": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_use_cluster_pair": ([_int], _int),
    "narde_mlp_forward_move2": ([_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_mlp_forward_move2_states": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_action_codes": ([_vp, _vp, _i64, _i32, _vp, _vp], _int),
    "narde_trajectory_append": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp], _int),
    "narde_afterstates": ([_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp], _int),
    "narde_afterstates_scan": ([_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp], _int),
    "narde_gather_overflow": ([_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_scatter_choice": ([_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "narde_step_chosen": ([_vp, _vp, _i64, _i64, _u64, _u64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _i32, _i32, _vp, _vp], _int),
    "narde_segment_argmax": ([_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp], _int),
    "narde_obs198": ([_vp, _vp, _i64, _vp, _vp], _int),
    "narde_obs24": ([_vp, _vp, _i64, _vp, _vp], _int),
    "narde_apply_actions": ([_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp], _int),
    "narde_roll_dice": ([_i64, _i64, _u64, _u64, _vp, _vp], _int),
    "narde_violates_block_rule": ([_vp, _i64, _vp, _vp], _int),
}

_lib = None


class NardeCudaError(RuntimeError):
    pass


def load():
    """Open libnarde_b200.so (built in-tree by gym_narde_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NardeCudaError(
            "gym_narde_b200: CUDA extension %s is missing. Build it with `python -m gym_narde_b200.build` "
            "(needs nvcc; sm_100a only). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.narde_abi_version() != ABI_VERSION:
        raise NardeCudaError("libnarde_b200.so ABI version mismatch")
    _lib = lib
    return lib


def exported_symbols():
    return tuple(_SIGNATURES)


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise NardeCudaError("gym_narde_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    return torch


def _check(rc, what):
    if rc != 0:
        msg = "bad argument" if rc == -1 else "cudaError %d" % rc
        raise NardeCudaError("%s failed: %s" % (what, msg))


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    import torch

    # device memory, or page-locked host memory (mapped into the device address space by CUDA's unified
    # addressing: the kernels read / write it directly over PCIe -- the zero-copy I/O of VecNardeEnv.step_host)
    if not isinstance(t, torch.Tensor) or not (t.is_cuda or t.is_pinned()):
        raise NardeCudaError("%s must be a CUDA tensor or a pinned host tensor (no CPU fallback)" % name)
    if not t.is_contiguous():
        raise NardeCudaError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise NardeCudaError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    # launches go to the CURRENT device's current stream: a tensor of another GPU would be an illegal address there
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise NardeCudaError("%s lives on %s but the current CUDA device is cuda:%d: call torch.cuda.set_device(...) "
                             "or wrap the call in `with torch.cuda.device(...)`" % (name, t.device, torch.cuda.current_device()))
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------------------------------
# thin wrappers (tensors in, tensors mutated in place); all asynchronous on the current stream
# ------------------------------------------------------------------------------------------
def reset(lo, hi, env_base, seed, step, mask=None):
    import torch

    n = lo.shape[0]
    rc = load().narde_reset_masked(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                                   _ptr(mask, torch.uint8, "mask"), n, env_base, seed, step, _stream())
    _check(rc, "narde_reset")


def half_moves(lo, hi, dice4, moves, counts, player_override=0):
    import torch

    rc = load().narde_half_moves(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                                 _ptr(dice4, torch.uint8, "dice"), lo.shape[0], player_override,
                                 _ptr(moves, torch.uint8, "moves"), _ptr(counts, torch.int32, "counts"), _stream())
    _check(rc, "narde_half_moves")


def step_ref(lo, hi, dice, codes, obs24, reward, done, max_episode_steps=0, truncated=None):
    import torch

    rc = load().narde_step_ref(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                               _ptr(dice, torch.uint8, "dice"), _ptr(codes, torch.int32, "codes"), lo.shape[0],
                               max_episode_steps, _ptr(obs24, torch.int32, "obs24"),
                               _ptr(reward, torch.int32, "reward"), _ptr(done, torch.uint8, "done"),
                               _ptr(truncated, torch.uint8, "truncated"), _stream())
    _check(rc, "narde_step_ref")


def enumerate_actions(lo, hi, dice, actions, counts, overflow=None):
    import torch

    cap = actions.shape[1] if actions is not None else 0
    rc = load().narde_enumerate(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                                _ptr(dice, torch.uint8, "dice"), lo.shape[0], cap,
                                _ptr(actions, torch.int64, "actions"), _ptr(counts, torch.int32, "counts"),
                                _ptr(overflow, torch.uint8, "overflow"), _stream())
    _check(rc, "narde_enumerate")


def enumerate_actions_fast(lo, hi, dice, actions, counts, overflow=None, workspace=None):
    import torch

    cap = actions.shape[1] if actions is not None else 0
    if workspace is not None and workspace.numel() < workspace_ints(lo.shape[0]):
        raise NardeCudaError("workspace must hold NARDE_WORKSPACE_INTS(n) = n + 8 int32")
    rc = load().narde_enumerate_fast(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                                     _ptr(dice, torch.uint8, "dice"), lo.shape[0], cap,
                                     _ptr(actions, torch.int64, "actions"), _ptr(counts, torch.int32, "counts"),
                                     _ptr(overflow, torch.uint8, "overflow"), _ptr(workspace, torch.int32, "workspace"),
                                     _stream())
    _check(rc, "narde_enumerate_fast")


def step_full(lo, hi, env_base, seed, step, dice_in=None, action_idx=None, actions=None, counts=None,
              dice_out=None, chosen=None, obs198=None, reward=None, done=None, stats=None, flags=0,
              max_episode_steps=0, truncated=None, workspace=None, step_dev=None, mirror_lo=None, mirror_hi=None):
    import torch

    cap = actions.shape[1] if actions is not None else 0
    if workspace is not None and workspace.numel() < workspace_ints(lo.shape[0]):
        raise NardeCudaError("workspace must hold NARDE_WORKSPACE_INTS(n) = n + 8 int32")
    rc = load().narde_step_full_mirror(
        _ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"), lo.shape[0], env_base, seed, step,
        _ptr(dice_in, torch.uint8, "dice_in"), _ptr(action_idx, torch.int32, "action_idx"), cap,
        _ptr(actions, torch.int64, "actions"), _ptr(counts, torch.int32, "counts"),
        _ptr(dice_out, torch.uint8, "dice_out"), _ptr(chosen, torch.int64, "chosen"),
        _ptr(obs198, torch.float32, "obs198"), _ptr(reward, torch.float32, "reward"),
        _ptr(done, torch.uint8, "done"), _ptr(truncated, torch.uint8, "truncated"), _ptr(stats, torch.int64, "stats"),
        flags, max_episode_steps, _ptr(workspace, torch.int32, "workspace"), _ptr(step_dev, torch.int64, "step_dev"),
        _ptr(mirror_lo, torch.uint8, "mirror_lo"), _ptr(mirror_hi, torch.uint8, "mirror_hi"), _stream())
    _check(rc, "narde_step_full")


def advance_counter(counter):
    import torch

    _check(load().narde_advance_counter(_ptr(counter, torch.int64, "counter"), _stream()), "narde_advance_counter")


def obs198(lo, hi, out):
    import torch

    rc = load().narde_obs198(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"), lo.shape[0],
                             _ptr(out, torch.float32, "obs198"), _stream())
    _check(rc, "narde_obs198")


def obs24(lo, hi, out):
    import torch

    rc = load().narde_obs24(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"), lo.shape[0],
                            _ptr(out, torch.int32, "obs24"), _stream())
    _check(rc, "narde_obs24")


def apply_actions(lo, hi, acts, reward=None, done=None, flags=0):
    import torch

    rc = load().narde_apply_actions(_ptr(lo, torch.uint8, "lo"), _ptr(hi, torch.uint8, "hi"),
                                    _ptr(acts, torch.int64, "acts"), lo.shape[0], flags,
                                    _ptr(reward, torch.float32, "reward"), _ptr(done, torch.uint8, "done"), _stream())
    _check(rc, "narde_apply_actions")


def roll_dice(dice, env_base, seed, step):
    import torch

    rc = load().narde_roll_dice(dice.shape[0], env_base, seed, step, _ptr(dice, torch.uint8, "dice"), _stream())
    _check(rc, "narde_roll_dice")


def violates_block_rule(boards, out):
    import torch

    rc = load().narde_violates_block_rule(_ptr(boards, torch.int8, "boards"), boards.shape[0],
                                          _ptr(out, torch.uint8, "out"), _stream())
    _check(rc, "narde_violates_block_rule")
