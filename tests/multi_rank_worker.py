"""Worker of tests/test_multi_rank.py: one gloo rank (RANK / WORLD_SIZE / MASTER_* from the environment)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py as O
    import wga_pkg
    W = wga_pkg.load()
    base = os.path.join(GOLDEN, "cnr2000_head")
    z = np.load(os.path.join(GOLDEN, "cnr2000_head.npz"))
    whole = W.ANSBvGraph.load(base, host_only=True)
    pre = whole.prelude()
    n = whole.num_nodes()
    ranges = W.shard_ranges(pre["pointers"], world)
    plans = [None] * world  # every rank derives the same plan
    dist.all_gather_object(plans, ranges)
    assert all(p == ranges for p in plans)
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[k][1] == ranges[k + 1][0] for k in range(world - 1))
    first, last = ranges[rank]
    lo, hi = W.shard_resident_range(first, last, whole.compression_window())
    shard = W.ANSBvGraph.load(base, shard=(lo, hi), host_only=True)
    assert shard.num_nodes() == n
    # the checker: oracle decode of exactly this node range equals the golden slice
    og = O.OracleGraph.load(base)
    off, succ, _ = og.decode_seq(first, last)
    g_off, g_succ = z["offsets"], z["succ"]
    assert (off == g_off[first:last + 1] - g_off[first]).all()
    assert (succ == g_succ[g_off[first]:g_off[last]]).all()
    # balanced by stream words: compare the stream spans of the ranks
    ptr = np.asarray(pre["pointers"], np.uint64)[::-1].astype(np.int64)  # by node
    words = int(ptr[first] - (ptr[last] if last < n else 0))
    t = torch.tensor([succ.size, last - first, words], dtype=torch.int64)
    mx = t.clone()
    dist.all_reduce(t)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    assert int(t[0]) == g_succ.size and int(t[1]) == n and int(t[2]) == int(ptr[0])
    assert int(mx[2]) <= int(t[2]) // world + 600, (int(mx[2]), int(t[2]))
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok: nodes [%d,%d) %d arcs %d words" % (rank, first, last, succ.size, words))


if __name__ == "__main__":
    main()
