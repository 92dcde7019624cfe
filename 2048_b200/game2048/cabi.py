"""ctypes binding of include/b2048.h (libb2048.so).

This is the stub a maintainer of the reference would add: the reference is pure
Python, so its "FFI" is this module.  There is NO CPU fallback: if the library is missing or no CUDA
device is present, every compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(os.path.dirname(_HERE), "libb2048.so")

UPD_ATOMIC, UPD_DETERMINISTIC, UPD_SUM, UPD_MEAN, UPD_SORTED, RUN_STEPWISE, RUN_GENERIC = 0, 1, 0, 2, 4, 8, 16
RUN_SCAN, RUN_LISTS, RUN_EVEN = 32, 64, 128
F_HAVE_STATE, F_DONE, F_OVERFLOW = 1, 2, 4
(CTR_MOVES, CTR_EVALS, CTR_UPDATES, CTR_FINISHED, CTR_SCORE_SUM, CTR_MOVES_SUM, CTR_OVERFLOW, CTR_ACTIVE,
 CTR_LOG, CTR_QUEUE, CTR_FAULT) = range(11)
CTR_COUNT = 16
LUT_ENTRIES = 65536


class B2048Error(RuntimeError):
    pass


class Games(C.Structure):
    """b2048_games_t"""
    _fields_ = [("B", C.c_int64), ("board", C.c_void_p), ("score", C.c_void_p), ("moves", C.c_void_p),
                ("game_id", C.c_void_p), ("state", C.c_void_p), ("old_label", C.c_void_p), ("flags", C.c_void_p),
                ("counters", C.c_void_p), ("tile_hist", C.c_void_p), ("seed", C.c_uint64), ("id_stride", C.c_uint64),
                ("fin_log", C.c_void_p), ("fin_cap", C.c_int64)]


class Replay(C.Structure):
    """b2048_replay_t"""
    _fields_ = [("tile", C.c_void_p), ("pos", C.c_void_p), ("len", C.c_int64)]


MAX_PEERS, PEER_ARRIVE, PEER_DONE, PEER_TICKET, PEER_FAULT, PEER_FLAG_WORDS = 16, 0, 16, 32, 33, 64


class Peers(C.Structure):
    """b2048_peers_t: every rank's weights / w_sync / flag block as mapped into this process"""
    _fields_ = [("w", C.c_void_p * MAX_PEERS), ("w_sync", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS),
                ("world", C.c_int), ("rank", C.c_int)]


_vp, _i64, _u64, _int, _f32, _sz = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_float, C.c_size_t
_GP, _RP, _PP = C.POINTER(Games), C.POINTER(Replay), C.POINTER(Peers)

# name -> (restype, argtypes); mirrors include/b2048.h declaration by declaration
SIGNATURES = {
    "b2048_abi_version": (_int, []),
    "b2048_strerror": (C.c_char_p, [_int]),
    "b2048_num_feat": (_int, [_int]),
    "b2048_table_offset": (_i64, [_int, _int]),
    "b2048_num_weights": (_i64, [_int]),
    "b2048_pack": (_int, [_vp, _vp, _i64, _vp]),
    "b2048_unpack": (_int, [_vp, _vp, _i64, _vp]),
    "b2048_lut_build": (_int, [_vp, _vp]),
    "b2048_move4": (_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b2048_board_stats": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "b2048_spawn_philox": (_int, [_vp, _i64, _u64, _vp, _vp, _vp, _vp]),
    "b2048_spawn_initial": (_int, [_vp, _i64, _u64, _u64, _u64, _vp]),
    "b2048_spawn_replay": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "b2048_sweep": (_int, [_vp, _vp, _i64, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "b2048_features": (_int, [_int, _vp, _i64, _vp, _vp]),
    "b2048_evaluate": (_int, [_int, _vp, _vp, _i64, _vp, _vp]),
    "b2048_td_update_workspace": (_sz, [_int, _i64, _int]),
    "b2048_td_update": (_int, [_int, _vp, _vp, _vp, _vp, _i64, _int, _vp, _sz, _vp]),
    "b2048_games_init": (_int, [_GP, _u64, _int, _vp]),
    "b2048_greedy_play": (_int, [_int, _vp, _vp, _GP, _int, _int, _int, _RP, _vp, _vp, _vp, _i64, _vp]),
    "b2048_td_step": (_int, [_int, _vp, _vp, _vp, _GP, _f32, _int, _vp, _vp, _vp, _sz, _RP, _vp, _vp, _vp, _vp, _i64,
                            _vp]),
    "b2048_td_phase_a": (_int, [_int, _vp, _vp, _GP, _f32, _vp, _vp, _RP, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b2048_td_run": (_int, [_int, _vp, _vp, _vp, _GP, _f32, _int, _int, _vp, _vp, _vp, _sz, _vp]),
    "b2048_td_run_launches": (_i64, [_int, _i64, _int, _int]),
    "b2048_look_forward": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _int, _u64, _vp, _vp]),
    "b2048_expectimax_play": (_int, [_int, _vp, _vp, _GP, _int, _int, _int, _int, _int, _int, _vp, _vp, _i64, _vp]),
    "b2048_delta_pack": (_int, [_vp, _vp, _i64, _vp]),
    "b2048_delta_pack_diff": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "b2048_delta_apply": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "b2048_delta_pack_bits": (_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "b2048_delta_apply_bits": (_int, [_vp, _vp, _vp, _vp, _int, _i64, _vp]),
    "b2048_sync_peers": (_int, [_PP, _i64, C.c_uint32, _int, _vp]),
    "b2048_td_run_peers": (_int, [_int, _vp, _vp, _GP, _f32, _int, _int, _vp, _vp, _vp, _sz, _PP, _int, _int, C.c_uint32,
                                  _vp]),
}

_lib = None


def lib():
    """The loaded libb2048.so.  Raises B2048Error (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B2048Error(f"{SO_PATH} is missing: build it with `python 2048_b200/build.py` "
                             "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.b2048_abi_version() != 1:
            raise B2048Error("libb2048.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().b2048_strerror(rc).decode()
        raise B2048Error(f"{what or 'b2048 call'} failed ({rc}): {msg}")


def num_feat(n):
    f = lib().b2048_num_feat(n)
    if f < 0:
        raise B2048Error(f"unknown n={n}")
    return f


def table_offsets(n):
    L = lib()
    return [L.b2048_table_offset(n, i) for i in range(num_feat(n) + 1)]


def num_weights(n):
    return lib().b2048_num_weights(n)
