"""N>1 host logic on CPU: two gloo ranks plan their node-range shards, open them host-only through the C ABI
and check (with the oracle as the checker) that the shards tile the graph.  No GPU needed."""
import os
import sys

import numpy as np
import pytest
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("world", [2, 3])
def test_shard_plan_with_gloo_ranks(world):
    port = 29500 + (os.getpid() % 2000) + world
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_rank_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("ok" in o for o in outs)


def test_shard_ranges_are_balanced_and_cover():
    sys.path.insert(0, ROOT)
    import wga_pkg
    W = wga_pkg.load()
    rng = np.random.default_rng(0)
    words = rng.integers(0, 9, 10000)
    ptr = np.cumsum(words).astype(np.uint64)  # file order: non-decreasing
    for world in (1, 2, 4, 8):
        r = W.shard_ranges(ptr, world)
        assert r[0][0] == 0 and r[-1][1] == 10000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        by_node = ptr[::-1].astype(np.int64)
        sizes = [int(by_node[a] - (by_node[b] if b < 10000 else 0)) if b > a else 0 for a, b in r]
        assert max(sizes) - min(sizes) <= 16 + int(ptr[-1]) // (50 * world)


@pytest.mark.gpu
def test_two_gpu_ranks_decode_their_shards_and_build_one_model():
    """One process per GPU over NCCL (skipped on a one-GPU box): sharded decode of one graph, bit-exact per shard,
    and the 2-rank model build (all-reduced histograms) against the oracle's tables on the union of the symbols."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    port = 29700 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_gpu_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("ok" in o for o in outs)
