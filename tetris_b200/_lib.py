"""ctypes binding of the C ABI in include/tetris_b200.h (csrc/libtetris_b200.so).

There is no CPU fallback: if the CUDA extension cannot be loaded every entry point raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("TB_SO_PATH") or os.path.join(CSRC, "libtetris_b200.so")   # override: experiments only
SOURCES = [os.path.join(CSRC, "tb_kernels.cu"), os.path.join(CSRC, "tb_core.cuh"),
           os.path.join(os.path.dirname(_HERE), "include", "tetris_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

NUM_FEATURES = 8
FLAG_AUTO_RESET = 1
FLAG_ACTION_IS_SLOT = 2
FLAG_INCLUDE_TERMINAL = 4
POLICY_RANDOM = 0
POLICY_GREEDY = 1
STATS = ("placements", "episodes", "lines", "reward", "afterstates",
         "lines0", "lines1", "lines2", "lines3", "lines4",
         "max_ep_lines", "max_ep_steps", "sum_ep_steps", "sum_ep_lines", "reserved0", "reserved1")
STATS_MAX_FIELDS = (10, 11)


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SOURCES)


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not (force or _stale()):
        return SO_PATH
    nvcc = os.environ.get("NVCC") or "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, SOURCES[0]]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO_PATH


_lib = None


def lib():
    """The loaded extension.  Raises (loudly) when it is missing -- there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "tetris_b200: CUDA extension %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc); there is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.tb_version.restype = i32
    L.tb_last_error.restype = C.c_char_p
    L.tb_supported_shape.restype = i32
    L.tb_supported_shape.argtypes = [i32, i32]
    L.tb_state_bytes.restype = C.c_size_t
    L.tb_state_bytes.argtypes = [i32, i32, i64]
    L.tb_num_slots.restype = i32
    L.tb_num_slots.argtypes = [i32, i32]
    L.tb_a_max.restype = i32
    L.tb_a_max.argtypes = [i32, i32]
    L.tb_reset.restype = i32
    L.tb_reset.argtypes = [vp, i32, i32, i64, i64, u64, i32, vp, vp, vp]
    L.tb_afterstates.restype = i32
    L.tb_afterstates.argtypes = [vp, i32, i32, i64, vp, vp, vp, i32, vp, i32, vp]
    L.tb_afterstates_export.restype = i32
    L.tb_afterstates_export.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, i32, vp]
    L.tb_step.restype = i32
    L.tb_step.argtypes = [vp, i32, i32, i64, i64, u64, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.tb_rollout.restype = i32
    L.tb_rollout.argtypes = [vp, i32, i32, i64, i64, u64, i32, i32, i32, vp, vp, vp]
    L.tb_export_boards.restype = i32
    L.tb_export_boards.argtypes = [vp, i32, i32, i64, i64, i64, vp, vp, vp, vp]
    L.tb_import_boards.restype = i32
    L.tb_import_boards.argtypes = [vp, i32, i32, i64, i64, i64, vp, vp, vp]
    L.tb_eval_states.restype = i32
    L.tb_eval_states.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    L.tb_rollout_values.restype = i32
    L.tb_rollout_values.argtypes = [vp, i32, i32, i64, i32, vp, i32, i32, i32, i32, vp, u64, i64, vp, vp, vp, vp]
    L.tb_action_probabilities.restype = i32
    L.tb_action_probabilities.argtypes = [i64, i32, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    L.tb_slot_info.restype = i32
    L.tb_slot_info.argtypes = [i32, i32, i32, vp]
    L.tb_fitness.restype = i32
    L.tb_fitness.argtypes = [i64, vp, vp, vp, vp]
    _lib = L
    return L


EXPORTS = ("tb_version", "tb_last_error", "tb_supported_shape", "tb_state_bytes", "tb_num_slots", "tb_a_max",
           "tb_reset", "tb_afterstates", "tb_afterstates_export", "tb_step", "tb_rollout", "tb_export_boards",
           "tb_import_boards", "tb_eval_states", "tb_slot_info", "tb_fitness", "tb_rollout_values", "tb_action_probabilities")


def check(rc):
    if rc != 0:
        raise RuntimeError("tetris_b200: %s" % lib().tb_last_error().decode())
