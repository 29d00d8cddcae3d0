import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_CNR = "/root/reference/tests/data/cnr-2000/cnr-2000"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def W():
    """The product package (ctypes over libwgans.so)."""
    import wga_pkg
    mod = wga_pkg.load()
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def gpu(W):
    if not W.cuda_available():
        pytest.skip("no CUDA device")
    import torch
    torch.cuda.init()
    return torch


@pytest.fixture(scope="session")
def head():
    """Golden fixture: head of cnr-2000 recompressed by the oracle (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "cnr2000_head.npz"))
    return dict(base=os.path.join(GOLDEN, "cnr2000_head"), offsets=z["offsets"], succ=z["succ"], comps=z["comps"],
                syms=z["syms"])


def random_graph(rng, n, mean_deg, locality=0.5, copy_prob=0.5, run_prob=0.3):
    """Small random graph with copy structure, runs and far links (numpy, seeded)."""
    lists = []
    for v in range(n):
        s = set()
        if lists and rng.random() < copy_prob:
            u = v - 1 - int(rng.integers(0, min(v, 7)))
            for x in lists[u]:
                if rng.random() < 0.8:
                    s.add(int(x))
        k = int(rng.poisson(mean_deg * 0.5))
        for _ in range(k):
            if rng.random() < locality:
                s.add(int(np.clip(v + rng.integers(-50, 50), 0, n - 1)))
            else:
                s.add(int(rng.integers(0, n)))
        if rng.random() < run_prob:
            a = int(rng.integers(0, n))
            for x in range(a, min(n, a + int(rng.integers(2, 12)))):
                s.add(x)
        if rng.random() < 0.1:
            s = set()
        lists.append(np.array(sorted(s), dtype=np.uint32))
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists])
    succ = np.concatenate(lists).astype(np.uint32) if n else np.zeros(0, np.uint32)
    return off, succ


def open_oracle_graph(W, g, host_only=False):
    """Hands an oracle-compressed graph (arrays) to the product through wga_open_mem."""
    inf = g.info()
    states, pointers = g.phases()
    return W.open_mem(g.tables(), g.stream(), inf["state"], inf["n"], inf["window"], inf["min_interval"], inf["arcs"],
                      states, pointers, host_only=host_only)
