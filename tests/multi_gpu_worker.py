"""Worker of tests/test_multi_gpu.py (one process per GPU, NCCL): shard a job over the ranks, run fused rollouts,
all-reduce the episode statistics, and check on rank 0 that boards and statistics equal the unsharded job's."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tetris_b200 import BatchedTetris                      # noqa: E402
from tetris_b200 import distributed as D                   # noqa: E402


def main():
    rank, world, local = D.world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total, seed = 50000 + 7, 99
    for (C, R) in ((10, 20), (6, 12)):
        shard = D.make_shard(C, R, total, piece_set=1, seed=seed)
        shard.rollout(20, "random")
        shard.rollout(15, "greedy")
        red = D.reduce_stats(shard.stats)
        rows = [None] * world
        dist.all_gather_object(rows, (shard.env_offset, shard.rows(), shard.piece))
        if rank == 0:
            whole = BatchedTetris(C, R, total, piece_set=1, seed=seed)
            whole.rollout(20, "random")
            whole.rollout(15, "greedy")
            rows.sort(key=lambda t: t[0])
            assert np.array_equal(np.concatenate([r[1] for r in rows]), whole.rows()), "boards differ from the unsharded job"
            assert np.array_equal(np.concatenate([r[2] for r in rows]), whole.piece)
            assert torch.equal(red.cpu(), whole.stats.cpu()), (red.tolist(), whole.stats.tolist())
            assert int(red[0]) == 35 * total
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
