"""ctypes binding of libspx.so (include/spx.h).  Fails loudly when the CUDA library is missing:
the product path has no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPX_LIB_PATH") or os.path.join(_HERE, "libspx.so")  # override: experiment builds only

MAX_ACTIONS = 9
GAME_CONNECT4, GAME_TICTACTOE = 0, 1
GAME_DIMS = {GAME_CONNECT4: (7, 6, 7), GAME_TICTACTOE: (3, 3, 9)}  # W, H, A


class SpxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("game", C.c_int32), ("n_games", C.c_int32), ("sims", C.c_int32), ("evaluate", C.c_int32),
                ("strong_play", C.c_int32), ("tie_mode", C.c_int32), ("noise_mode", C.c_int32),
                ("emit_records", C.c_int32), ("max_sims_per_tick", C.c_int32), ("nodes_per_tree", C.c_int32),
                ("move_log", C.c_int32), ("two_nets", C.c_int32), ("opponent_kind", C.c_int32), ("reserved0", C.c_int32),
                ("alpha", C.c_double), ("seed", C.c_uint64),
                ("slot_offset", C.c_int64), ("slot_stride", C.c_int64), ("games_target", C.c_int64),
                ("record_capacity", C.c_int64), ("result_capacity", C.c_int64), ("search_threads", C.c_int32), ("eval_cache_log2", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("own", C.c_uint64), ("opp", C.c_uint64), ("game_index", C.c_uint64),
                ("tree_probs", C.c_float * MAX_ACTIONS), ("q", C.c_float), ("actual_val", C.c_float),
                ("tree", C.c_uint8), ("ply", C.c_uint8), ("pad0", C.c_uint16), ("pad1", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("game_index", C.c_uint64), ("reward", C.c_int8), ("swap_sides", C.c_uint8), ("plies", C.c_uint8),
                ("pad", C.c_uint8 * 5)]


class MoveLog(C.Structure):
    _fields_ = [("tree", C.c_int32), ("ply", C.c_int32), ("action", C.c_int32), ("root_n", C.c_int32),
                ("root_w", C.c_double), ("n", C.c_int32 * MAX_ACTIONS), ("pad", C.c_int32),
                ("w", C.c_double * MAX_ACTIONS), ("noise", C.c_double * MAX_ACTIONS)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("sims", "leaf_evals", "terminal_sims", "path_len_sum", "moves",
                                          "games_finished", "nodes_allocated", "ticks", "records_dropped", "errors", "cache_hits")]


# every symbol include/spx.h declares (checked by tests/test_abi.py)
EXPORTS = ["spx_last_error", "spx_version", "spx_launch_count", "spx_env_step", "spx_env_valid_moves",
           "spx_hashnet_forward", "spx_create", "spx_destroy", "spx_reset", "spx_set_noise_table", "spx_advance",
           "spx_leaf_batch", "spx_root_stats", "spx_drain_records", "spx_drain_results", "spx_read_move_log",
           "spx_counters_read", "spx_all_idle", "spx_device_bytes", "spx_pending_tree", "spx_tower_blob_bytes",
           "spx_tower_create", "spx_tower_destroy", "spx_tower_ncta", "spx_tower_fused_heads", "spx_tower_f16", "spx_tower_version", "spx_set_eval_cache_versions", "spx_softf64_selftest", "spx_tower_load", "spx_tower_forward", "spx_tower_forward_timed", "spx_partition_leaves", "spx_scatter_outputs", "spx_tttnet_blob_floats", "spx_tttnet_create", "spx_tttnet_destroy",
           "spx_tttnet_load", "spx_tttnet_forward",
           "spx_advance_timed", "spx_restart", "spx_set_sims", "spx_set_external_actions", "spx_slot_status", "spx_event_create", "spx_event_destroy", "spx_event_elapsed_ms",
           "spx_replay_create", "spx_replay_destroy", "spx_replay_size", "spx_replay_max_size", "spx_replay_change_size", "spx_replay_reset",
           "spx_drain_records_device", "spx_replay_append", "spx_replay_read", "spx_replay_sample", "spx_replay_deduplicate", "spx_replay_unique", "spx_tick_fused", "spx_tick_fused_balanced",
           "spx_train_create", "spx_train_destroy", "spx_train_param_count", "spx_train_running_count", "spx_train_set_state", "spx_train_get_state",
           "spx_train_step", "spx_train_outputs", "spx_train_debug_planes", "spx_train_debug_wgrad", "spx_train_debug_trace"]

_lib = None


def lib():
    """Loads libspx.so.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is mandatory; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.spx_last_error.restype = C.c_char_p
        L.spx_launch_count.restype = C.c_uint64
        L.spx_device_bytes.restype = C.c_int64
        L.spx_device_bytes.argtypes = [C.c_void_p]
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.spx_env_step.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.spx_env_valid_moves.argtypes = [i32, i64, vp, vp, vp]
        L.spx_hashnet_forward.argtypes = [i32, i64, vp, vp, vp, vp, u64, u64, vp, vp, vp]
        L.spx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.spx_destroy.argtypes = [vp]
        L.spx_reset.argtypes = [vp, vp]
        L.spx_set_noise_table.argtypes = [vp, vp, i64, i64, i32]
        L.spx_advance.argtypes = [vp, vp, vp, vp]
        L.spx_leaf_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.spx_root_stats.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
        L.spx_drain_records.argtypes = [vp, vp, i64, C.POINTER(i64), vp]
        L.spx_drain_results.argtypes = [vp, vp, i64, C.POINTER(i64), vp]
        L.spx_read_move_log.argtypes = [vp, i32, vp, i32, C.POINTER(i32), vp]
        L.spx_counters_read.argtypes = [vp, C.POINTER(Counters), vp]
        L.spx_all_idle.argtypes = [vp, C.POINTER(i32), vp]
        L.spx_pending_tree.argtypes = [vp, vp, vp]
        L.spx_restart.argtypes = [vp, i64, i64, vp]
        L.spx_set_sims.argtypes = [vp, i32]
        L.spx_set_external_actions.argtypes = [vp, vp, vp]
        L.spx_slot_status.argtypes = [vp, vp, vp]
        L.spx_tower_blob_bytes.restype = C.c_int64
        L.spx_tower_blob_bytes.argtypes = [i32, i32]
        L.spx_tower_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.spx_tower_destroy.argtypes = [vp]
        L.spx_tower_ncta.argtypes = [vp]
        L.spx_tower_fused_heads.argtypes = [vp]
        L.spx_tower_f16.argtypes = [vp]
        L.spx_tower_version.argtypes = [vp]
        L.spx_tower_version.restype = C.c_uint32
        L.spx_set_eval_cache_versions.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.spx_softf64_selftest.argtypes = [C.c_uint64, C.c_uint64, vp, vp]
        L.spx_tower_load.argtypes = [vp, vp, i64, vp]
        L.spx_tower_forward.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
        L.spx_tower_forward_timed.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        L.spx_advance_timed.argtypes = [vp, vp, vp, vp, vp, vp]
        L.spx_partition_leaves.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.spx_scatter_outputs.argtypes = [i64, i32, vp, vp, vp, vp, vp, vp, vp]
        L.spx_tttnet_blob_floats.restype = C.c_int64
        L.spx_tttnet_create.argtypes = [C.POINTER(vp)]
        L.spx_tttnet_destroy.argtypes = [vp]
        L.spx_tttnet_load.argtypes = [vp, vp, i64, vp]
        L.spx_tttnet_forward.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
        L.spx_replay_create.argtypes = [i64, i64, C.POINTER(vp)]
        L.spx_replay_destroy.argtypes = [vp]
        L.spx_replay_size.restype = C.c_int64
        L.spx_replay_size.argtypes = [vp]
        L.spx_replay_max_size.restype = C.c_int64
        L.spx_replay_max_size.argtypes = [vp]
        L.spx_replay_change_size.argtypes = [vp, i64]
        L.spx_replay_reset.argtypes = [vp]
        L.spx_drain_records_device.argtypes = [vp, vp, i64, C.POINTER(i64), vp]
        L.spx_replay_append.argtypes = [vp, vp, i64, vp]
        L.spx_replay_read.argtypes = [vp, i64, i64, vp, vp]
        L.spx_replay_sample.argtypes = [vp, i32, i64, u64, u64, vp, vp, vp, vp, vp, vp, vp]
        L.spx_tick_fused.argtypes = [vp, vp, i32, vp, vp, vp]
        L.spx_tick_fused_balanced.argtypes = [vp, vp, i32, vp, vp, vp]
        L.spx_replay_deduplicate.argtypes = [vp, i64, vp]
        L.spx_replay_unique.restype = C.c_int64
        L.spx_replay_unique.argtypes = [vp]
        L.spx_train_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.spx_train_destroy.argtypes = [vp]
        L.spx_train_param_count.restype = C.c_int64
        L.spx_train_param_count.argtypes = [vp]
        L.spx_train_running_count.restype = C.c_int64
        L.spx_train_running_count.argtypes = [vp]
        L.spx_train_set_state.argtypes = [vp, vp, vp, i32, vp]
        L.spx_train_get_state.argtypes = [vp, i32, vp, vp]
        L.spx_train_step.argtypes = [vp, vp, vp, vp, vp, u64, u64, C.c_float, C.c_float, C.c_float, i32, vp, vp]
        L.spx_train_outputs.argtypes = [vp, vp, vp, vp]
        L.spx_train_debug_trace.argtypes = [vp]
        L.spx_train_debug_wgrad.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, i32, vp]
        L.spx_train_debug_planes.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32)]
        L.spx_event_create.argtypes = [C.POINTER(vp)]
        L.spx_event_destroy.argtypes = [vp]
        L.spx_event_elapsed_ms.argtypes = [vp, vp, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def check(rc, what="libspx call"):
    if rc != 0:
        raise SpxError(f"{what} failed ({rc}): {lib().spx_last_error().decode()}")


class Event:
    """A cudaEvent_t owned by libspx (torch.cuda.Event creates its handle lazily, so it cannot be handed to C)."""

    def __init__(self):
        self.h = C.c_void_p()
        check(lib().spx_event_create(C.byref(self.h)), "spx_event_create")

    @property
    def cuda_event(self):
        return self.h.value

    def elapsed_time(self, end):
        ms = C.c_float()
        check(lib().spx_event_elapsed_ms(self.h, end.h, C.byref(ms)), "spx_event_elapsed_ms")
        return ms.value

    def __del__(self):
        try:
            lib().spx_event_destroy(self.h)
        except Exception:
            pass
