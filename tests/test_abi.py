"""CPU-only: the C-ABI library loads and exports every symbol include/tetris_b200.h declares; host-side
helpers that need no GPU return the reference's numbers."""
import os
import re

from tetris_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "tetris_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|size_t|const char \*)\s*\*?(tb_\w+)\s*\(", hdr, re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_host_queries():
    L = _lib.lib()
    assert L.tb_version() == 200
    assert [L.tb_num_slots(p, 10) for p in range(9)] == [17, 34, 34, 9, 17, 17, 34, 36, 18]
    assert [L.tb_num_slots(p, 6) for p in range(9)] == [9, 18, 18, 5, 9, 9, 18, 20, 10]
    assert L.tb_a_max(10, 0) == 36 and L.tb_a_max(10, 1) == 34
    assert L.tb_supported_shape(10, 20) and L.tb_supported_shape(10, 10) and L.tb_supported_shape(6, 12)
    assert not L.tb_supported_shape(3, 3)
    assert L.tb_state_bytes(10, 20, 1000) == 72 * 1000     # 3 row planes + meta + episode counters
    assert L.tb_state_bytes(10, 10, 1000) == 56 * 1000


def test_tuning_knobs():
    """tb_set_tuning / tb_get_tuning: experiments and tests steer the tile configuration through the ABI (the
    environment is read once, not on every launch)."""
    assert _lib.get_tuning("k1_cfg") == -1 and _lib.get_tuning("max_ctas") == 0 and _lib.get_tuning("small_groups") == 4
    _lib.set_tuning("max_ctas", 3)
    assert _lib.get_tuning("max_ctas") == 3
    _lib.set_tuning("max_ctas", 0)
    assert _lib.lib().tb_set_tuning(b"no_such_knob", 1) != 0


def test_load_shape_rejects_garbage():
    L = _lib.lib()
    assert L.tb_load_shape(b"/nonexistent/libtb_shape_9x9.so") != 0
    assert b"tb_load_shape" in L.tb_last_error()


def test_argument_errors_do_not_touch_the_gpu():
    L = _lib.lib()
    assert L.tb_reset(None, 3, 3, 10, 0, 0, 1, None, None, None) != 0
    assert b"unsupported" in L.tb_last_error()
    assert L.tb_rollout(None, 10, 20, 0, 0, 0, 1, 1, 1, None, None, None) != 0
    assert b"n_env" in L.tb_last_error()
