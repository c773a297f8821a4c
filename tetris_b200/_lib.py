"""ctypes binding of the C ABI in include/tetris_b200.h (csrc/libtetris_b200.so).

There is no CPU fallback: if the CUDA extension cannot be loaded every entry point raises.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("TB_SO_PATH") or os.path.join(CSRC, "libtetris_b200.so")   # override: experiments only
_INC = os.path.join(os.path.dirname(_HERE), "include", "tetris_b200.h")
SHAPE_DEPS = [os.path.join(CSRC, f) for f in ("tb_shape.cu", "tb_kernels.cuh", "tb_core.cuh", "tb_shape.h")] + [_INC]
ABI_DEPS = [os.path.join(CSRC, f) for f in ("tb_abi.cu", "tb_core.cuh", "tb_shape.h")] + [_INC]
SOURCES = sorted(set(SHAPE_DEPS + ABI_DEPS))
# board shapes linked into libtetris_b200.so (columns, rows); any other 4..16 x 4..27 shape: build_shape() + tb_load_shape
BUILTIN_SHAPES = ((10, 20), (10, 10), (6, 12), (8, 16), (4, 4), (16, 27), (12, 24))
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
BUILD_DIR = os.path.join(CSRC, "build")

NUM_FEATURES = 8
FLAG_AUTO_RESET = 1
FLAG_ACTION_IS_SLOT = 2
FLAG_INCLUDE_TERMINAL = 4
FLAG_VALIDATE_ONLY = 8
FLAG_FEATS_I16 = 16
POLICY_RANDOM = 0
POLICY_GREEDY = 1
STATS = ("placements", "episodes", "lines", "reward", "afterstates",
         "lines0", "lines1", "lines2", "lines3", "lines4",
         "max_ep_lines", "max_ep_steps", "sum_ep_steps", "sum_ep_lines", "reserved0", "reserved1")
STATS_MAX_FIELDS = (10, 11)


def _newer(deps, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _stale():
    return _newer(SOURCES, SO_PATH)


def _code_only(text):
    """C / CUDA source without comments and with white space collapsed: what the compiler sees, for source_hash."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return re.sub(r"\s+", " ", text).strip()


def source_hash():
    """sha256 (first 16 hex digits) over the CODE of the kernel sources (comments and white space do not count):
    profiles/<kernel>_latest.json records the value its ncu capture was taken from, and bench.py flags quoted capture
    figures as stale when the code has changed since."""
    import hashlib
    h = hashlib.sha256()
    for f in ("tb_core.cuh", "tb_kernels.cuh", "tb_shape.h"):
        h.update(_code_only(open(os.path.join(CSRC, f)).read()).encode())
    return h.hexdigest()[:16]


def _nvcc():
    return os.environ.get("NVCC") or "nvcc"


def _shape_obj(c, r):
    return os.path.join(BUILD_DIR, "tb_shape_%dx%d.o" % (c, r))


def _shape_stub(c, r, plugin=False):
    """A two-line translation unit per shape, so that every shape's cubin carries its own name inside the library
    (cuobjdump / nvdisasm / ncu tell them apart): build/tb_shape_<C>x<R>.cu."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    path = os.path.join(BUILD_DIR, "tb_shape_%dx%d%s.cu" % (c, r, "_plugin" if plugin else ""))
    text = "#define TB_C %d\n#define TB_R %d\n%s#include \"tb_shape.cu\"\n" % (
        c, r, "#define TB_SHAPE_PLUGIN 1\n" if plugin else "")
    if not os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)
    return path


def build(force=False, verbose=False, shapes=None):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU): one object per board
    shape (tb_shape.cu -DTB_C -DTB_R), compiled in parallel, plus the ABI object, linked into libtetris_b200.so."""
    shapes = tuple(shapes or BUILTIN_SHAPES)
    if not (force or _stale()):
        return SO_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs = []
    for (c, r) in shapes:
        obj = _shape_obj(c, r)
        if force or _newer(SHAPE_DEPS, obj):
            jobs.append((obj, [_nvcc()] + NVCC_FLAGS + extra + ["-I", CSRC, "-c", "-o", obj, _shape_stub(c, r)]))
    abi_obj = os.path.join(BUILD_DIR, "tb_abi.o")
    # the list of linked-in shapes is a generated header of the ABI object: rewritten only when the list changes
    inc = os.path.join(BUILD_DIR, "tb_builtin_shapes.inc")
    text = "#define TB_BUILTIN_SHAPES(X) " + " ".join("X(%d, %d)" % s for s in shapes) + "\n"
    if not os.path.exists(inc) or open(inc).read() != text:
        open(inc, "w").write(text)
    if force or _newer(ABI_DEPS + [inc], abi_obj):
        jobs.append((abi_obj, [_nvcc()] + NVCC_FLAGS + extra + ["-c", "-o", abi_obj, os.path.join(CSRC, "tb_abi.cu")]))
    procs = []
    max_par = max(1, min(len(jobs), os.cpu_count() or 1))
    pending = list(jobs)
    failed = None
    while pending or procs:
        while pending and len(procs) < max_par:
            obj, cmd = pending.pop(0)
            procs.append((obj, cmd, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        obj, cmd, p = procs.pop(0)
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            failed = failed or (cmd, out)
    if failed:
        raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(failed[0]), failed[1]))
    objs = [_shape_obj(c, r) for (c, r) in shapes] + [abi_obj]
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs + ["-ldl"],
                          cwd=CSRC)
    return SO_PATH


def build_variant(path, defines=(), shapes=((10, 20),)):
    """Experiments: a second library with extra -D flags (A/B runs load it through TB_SO_PATH)."""
    import tempfile
    tmp = tempfile.mkdtemp()
    inc = os.path.join(tmp, "shapes.h")
    open(inc, "w").write("#define TB_BUILTIN_SHAPES(X) " + " ".join("X(%d, %d)" % s for s in shapes) + "\n")
    objs, procs = [], []
    for (c, r) in shapes:
        obj = os.path.join(tmp, "shape_%dx%d.o" % (c, r))
        objs.append(obj)
        procs.append(subprocess.Popen([_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] +
                                      ["-DTB_C=%d" % c, "-DTB_R=%d" % r, "-c", "-o", obj, os.path.join(CSRC, "tb_shape.cu")]))
    abi = os.path.join(tmp, "abi.o")
    procs.append(subprocess.Popen([_nvcc()] + NVCC_FLAGS + ["-include", inc, "-c", "-o", abi, os.path.join(CSRC, "tb_abi.cu")]))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("variant build failed")
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", path] + objs + [abi, "-ldl"])
    return path


def shape_plugin_path(c, r):
    return os.path.join(CSRC, "libtb_shape_%dx%d.so" % (c, r))


def build_shape(c, r, force=False):
    """Compile the kernels of one more board shape into their own in-tree shared object (needs nvcc; ~1 minute)."""
    c, r = int(c), int(r)
    if not (4 <= c <= 16 and 4 <= r <= 27):
        raise ValueError("board shape %dx%d is outside 4..16 columns x 4..27 rows (uint16 row masks, 32-bit column masks)" % (c, r))
    path = shape_plugin_path(c, r)
    if force or _newer(SHAPE_DEPS, path):
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-shared", "-I", CSRC, "-o", path, _shape_stub(c, r, plugin=True)],
                              cwd=CSRC)
    return path


def ensure_shape(c, r):
    """Make board shape c x r available: built in, already loaded, an in-tree shape object, or compiled now (the
    reference's Tetris takes any num_columns / num_rows, game.py:21-31).  Raises when it cannot be provided."""
    L = lib()
    if L.tb_supported_shape(int(c), int(r)):
        return
    path = shape_plugin_path(int(c), int(r))
    if _newer(SHAPE_DEPS, path):
        import shutil
        if shutil.which(_nvcc()) is None:
            raise ValueError("board shape %dx%d is not compiled in and nvcc is not available to build it" % (c, r))
        build_shape(c, r)
    check(L.tb_load_shape(path.encode()))


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
    L.tb_rollout_values.argtypes = [vp, i32, i32, i64, i32, vp, i32, i32, i32, i32, vp, u64, i64, vp, vp, vp, vp, vp]
    L.tb_load_shape.restype = i32
    L.tb_load_shape.argtypes = [C.c_char_p]
    L.tb_set_tuning.restype = i32
    L.tb_set_tuning.argtypes = [C.c_char_p, i32]
    L.tb_get_tuning.restype = i32
    L.tb_get_tuning.argtypes = [C.c_char_p]
    L.tb_combine_stats.restype = i32
    L.tb_combine_stats.argtypes = [vp, i32, vp, vp]
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
           "tb_import_boards", "tb_eval_states", "tb_slot_info", "tb_fitness", "tb_rollout_values", "tb_action_probabilities",
           "tb_load_shape", "tb_set_tuning", "tb_get_tuning", "tb_combine_stats")


def check(rc):
    if rc != 0:
        raise RuntimeError("tetris_b200: %s" % lib().tb_last_error().decode())


def set_tuning(name, value):
    """Experiments / tests: force a tile configuration or cap the grid (see tb_set_tuning in include/tetris_b200.h)."""
    check(lib().tb_set_tuning(name.encode(), int(value)))


def get_tuning(name):
    return lib().tb_get_tuning(name.encode())
