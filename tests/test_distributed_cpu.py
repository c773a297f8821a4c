"""CPU (gloo, world_size 2): the N > 1 host logic -- env sharding and the end-of-rollout statistics reduction --
and the bench's reference arm under a multi-rank launch."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from tetris_b200 import distributed as D, _lib
rank, world, _ = D.world()
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(100 + rank)
stats = torch.randint(0, 1 << 40, (len(_lib.STATS),), generator=g, dtype=torch.int64)
red = D.reduce_stats(stats)
off, cnt = D.shard_range(1000003, rank, world)
allr = [None] * world
dist.all_gather_object(allr, (off, cnt, stats.tolist(), red.tolist()))
if rank == 0:
    print(json.dumps(allr))
dist.destroy_process_group()
""" % ROOT


def test_shard_range_partitions():
    from tetris_b200 import distributed as D
    for total in (0, 1, 7, 1 << 20, 8 * (1 << 20) + 5):
        for w in (1, 2, 3, 4, 8):
            spans = [D.shard_range(total, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_combine_stats_rule():
    from tetris_b200 import _lib, distributed as D
    a = torch.arange(16, dtype=torch.int64)
    b = torch.arange(16, dtype=torch.int64).flip(0) * 3
    c = D.combine_stats([a, b])
    for i in range(16):
        assert c[i] == (max(a[i], b[i]) if i in _lib.STATS_MAX_FIELDS else a[i] + b[i])
    assert torch.equal(D.reduce_stats(a), a)              # not initialised -> copy


def test_reduce_stats_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29500 + os.getpid() % 2000
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("[[")][-1]
    res = json.loads(line)
    from tetris_b200 import _lib, distributed as D
    assert [r[:2] for r in res] == [list(D.shard_range(1000003, r, 2)) for r in range(2)]
    s0, s1 = np.array(res[0][2]), np.array(res[1][2])
    want = s0 + s1
    for i in _lib.STATS_MAX_FIELDS:
        want[i] = max(s0[i], s1[i])
    assert res[0][3] == want.tolist() and res[1][3] == want.tolist()


def test_bench_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
