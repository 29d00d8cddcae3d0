"""Worker of tests/test_multi_rank.py::test_two_gpu_ranks_*: one NCCL rank on its own GPU
(RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment)."""
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
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import bench
    import oracle_py as O
    import wga_pkg
    W = wga_pkg.load()
    # ---- node-range shards of ONE graph: every rank opens and decodes only its range (k_halo finds the predecessors)
    base = os.path.join(GOLDEN, "cnr2000_head")
    z = np.load(os.path.join(GOLDEN, "cnr2000_head.npz"))
    pre = W.ANSBvGraph.load(base, host_only=True).prelude()
    first, last = W.shard_ranges(pre["pointers"], world)[rank]
    lo, hi = W.shard_resident_range(first, last, 7)
    g = W.ANSBvGraph.load(base, shard=(lo, hi))
    off, succ = g.decode_range(first, last)
    g_off, g_succ = z["offsets"], z["succ"]
    assert (off.cpu().numpy().astype(np.uint64) == g_off[first:last + 1] - g_off[first]).all()
    assert (succ.cpu().numpy().view(np.uint32) == g_succ[g_off[first]:g_off[last]]).all()
    t = torch.tensor([succ.numel(), last - first], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    assert int(t[0]) == g_succ.size and int(t[1]) == g_off.size - 1
    # ---- model build over the ranks: NCCL all-reduce of the histograms, tables against the oracle on the union
    ok = bench.nrank_model_parity(W, O, rank, world, dist)
    assert ok is None or ok is True
    # ---- the same collective behind the C ABI (wga_model_allreduce on a raw ncclComm_t): identical tables
    kind, n, deg, seed = bench.WORKLOADS["tiny"]
    a, b = bench.node_split(4 * bench.CHUNK_NODES, world)[rank]
    off_r, succ_r = W.synth_graph(kind, 4 * bench.CHUNK_NODES, deg, seed=seed, first=a, last=b, threads=4)
    comps, syms = W.bvcomp_symbols(off_r, succ_r, chunk_nodes=bench.CHUNK_NODES, threads=4, first_node=a, **bench.BVCOMP)
    syms = syms.copy()
    syms[::97] += np.uint64(1) << np.uint64(30 + rank)  # some large raw symbols: they travel in the sparse tail
    m1, m2 = W.ANSModel4EncoderBuilder(), W.ANSModel4EncoderBuilder()
    m1.push_symbols(comps, syms)
    m2.push_symbols(comps, syms)
    m1.all_reduce()
    comm = W.nccl_comm_from_torch()
    m2.all_reduce_nccl(comm)
    t1, t2 = m1.build()[0], m2.build()[0]
    for c in range(9):
        for f in ("frame_size", "radix", "fidelity"):
            assert t1[c][f] == t2[c][f], (c, f)
        assert t1[c]["entries"].size == t2[c]["entries"].size
        for f in ("freq", "cumul_freq", "upperbound"):
            assert (t1[c]["entries"][f] == t2[c]["entries"][f]).all(), (c, f)
    W.nccl_comm_destroy(comm)
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok: nodes [%d,%d) %d arcs" % (rank, first, last, succ.numel()))


if __name__ == "__main__":
    main()
