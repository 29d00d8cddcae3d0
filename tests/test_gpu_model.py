"""Parity of the GPU model builder (histogram + folding + frame normalisation) against the oracle's
restatement of ANSModel4EncoderBuilder::build.  Frequency tables must be bit-exact."""
import hashlib
import os

import numpy as np
import pytest

from conftest import random_graph

pytestmark = pytest.mark.gpu


def assert_tables_equal(ours, ref):
    for c in range(9):
        a, b = ours[c], ref[c]
        for k in ("frame_size", "radix", "fidelity", "folding_threshold", "folding_offset"):
            assert a[k] == b[k], (c, k, a[k], b[k])
        assert a["entries"].size == b["entries"].size, c
        for f in ("freq", "cumul_freq", "upperbound"):
            assert (a["entries"][f] == b["entries"][f]).all(), (c, f)


def build_both(W, O, comps, syms):
    mb = W.ANSModel4EncoderBuilder()
    mb.push_symbols(comps, syms)
    tables, oc, fc = mb.build()
    og = O.OracleGraph()
    roc, rfc = og.build_model(comps, syms)
    assert_tables_equal(tables, og.tables())
    # costs: CUDA log2 vs libm (<= 1 ulp per term); the approximated cost is summed in symbol-index order on both sides
    assert np.allclose(oc, roc, rtol=1e-9, atol=1e-6) and np.allclose(fc, rfc, rtol=1e-12, atol=1e-9)
    return tables


def test_dummy_sequences(W, O, gpu):
    build_both(W, O, [0] * 10, [1, 1, 1, 2, 2, 2, 3, 3, 4, 5])  # compressor_tests.rs:16
    build_both(W, O, [0] * 3, [1000, 1000, 2000])  # :47 (folding)
    a = [1, 1, 1, 2, 2, 2, 3, 3, 4, 5]
    b = [1, 3, 3, 3, 2, 2, 3, 3, 4, 5]
    build_both(W, O, [0, 2] * 10, [x for p in zip(a, b) for x in p])  # :113-114


@pytest.mark.parametrize("alpha,maxv", [(1.2, 1 << 30), (1.05, 1 << 30), (1.3, 1 << 47), (2.5, 1 << 20)])
def test_zipf_sequences(W, O, gpu, alpha, maxv):
    rng = np.random.default_rng(int(alpha * 100))
    syms = np.minimum(rng.zipf(alpha, 300_000), maxv).astype(np.uint64)
    comps = rng.integers(0, 9, syms.size).astype(np.uint8)
    build_both(W, O, comps, syms)


@pytest.mark.parametrize("seed", range(12))
def test_selection_is_stable_on_sensitive_inputs(W, O, gpu, seed):
    """Inputs built so that several (fidelity, radix, frame) candidates cost almost the same: flat distributions at
    the folding thresholds, frequencies that are ties for the sort, totals just around powers of two.  The chosen
    model and all table fields must equal the oracle's: the cost of every candidate is summed in symbol-index order
    (model4encoder_builder.rs:307-324), so the `ratio <= THETA` and `new_cost >= lowest_cost` decisions agree."""
    rng = np.random.default_rng(1000 + seed)
    parts, comps = [], []
    for c in range(9):
        kind = (seed + c) % 4
        if kind == 0:    # flat over a range that straddles folding thresholds 2^(F+R-1)
            hi = int(rng.choice([15, 16, 17, 127, 128, 129, 1023, 1024, 1025]))
            vals = rng.integers(0, hi + 1, int(rng.integers(2000, 60000)))
        elif kind == 1:  # many symbols with equal frequencies (sort ties), total near a power of two
            k = int(rng.integers(3, 400))
            reps = int(rng.choice([1, 2, 3, 4]))
            vals = np.repeat(rng.choice(1 << 20, k, replace=False), reps)
            vals = np.concatenate([vals, np.zeros(max(0, (1 << int(np.log2(max(2, vals.size)) + 1)) - vals.size - int(rng.integers(0, 3))), np.int64)])
        elif kind == 2:  # two-point mass with a long thin tail
            vals = np.concatenate([np.full(40000, 3), np.full(40000, 5), rng.integers(0, 1 << 30, 2000)])
        else:            # geometric with a scale near a frame boundary
            vals = rng.geometric(1.0 / float(rng.choice([31.5, 32, 63.9, 64.1, 255, 256])), 50000)
        parts.append(np.asarray(vals, np.uint64))
        comps.append(np.full(len(vals), c, np.uint8))
    build_both(W, O, np.concatenate(comps), np.concatenate(parts))


def test_geometric_uniform_and_tiny(W, O, gpu):
    rng = np.random.default_rng(4)
    parts = [rng.geometric(0.01, 100_000), rng.integers(0, 50_000, 100_000), rng.integers(0, 3, 1000),
             np.array([7]), rng.integers(1 << 40, 1 << 44, 5000)]
    syms = np.concatenate(parts).astype(np.uint64)
    comps = np.concatenate([np.full(len(p), c, np.uint8) for c, p in zip((0, 3, 4, 6, 8), parts)])
    build_both(W, O, comps, syms)


def test_graph_symbol_streams(W, O, gpu, head):
    """The (component, symbol) stream of a real BvComp pass (golden head of cnr-2000)."""
    t = build_both(W, O, head["comps"], head["syms"])
    og = O.OracleGraph.load(head["base"])
    assert_tables_equal(t, og.tables())  # == the tables inside the committed .ans


def test_incremental_and_device_input(W, O, gpu):
    import torch
    rng = np.random.default_rng(9)
    syms = np.minimum(rng.zipf(1.15, 400_000), 1 << 35).astype(np.uint64)
    comps = rng.integers(0, 9, syms.size).astype(np.uint8)
    mb = W.ANSModel4EncoderBuilder()
    for a in range(0, syms.size, 70_001):
        if (a // 70_001) % 2:
            mb.push_symbols(comps[a:a + 70_001], syms[a:a + 70_001])
        else:
            mb.push_symbols(torch.from_numpy(comps[a:a + 70_001]).cuda(),
                            torch.from_numpy(syms[a:a + 70_001].view(np.int64)).cuda())
    tables, _, _ = mb.build()
    og = O.OracleGraph()
    og.build_model(comps, syms)
    assert_tables_equal(tables, og.tables())


def test_two_rank_merge_on_one_gpu(W, O, gpu):
    """What the N-rank model build does: dense bins are summed (all-reduce), the sparse tail of large
    symbols is exchanged and merged.  Emulated with two builders on one GPU."""
    rng = np.random.default_rng(10)
    syms = np.minimum(rng.zipf(1.1, 300_000), 1 << 40).astype(np.uint64)
    comps = rng.integers(0, 9, syms.size).astype(np.uint8)
    half = syms.size // 2
    a, b = W.ANSModel4EncoderBuilder(), W.ANSModel4EncoderBuilder()
    a.push_symbols(comps[:half], syms[:half])
    b.push_symbols(comps[half:], syms[half:])
    import ctypes as C
    n = int(W.lib().wga_model_sparse_count(b._h))
    sc, ss, sk = np.zeros(n, np.uint8), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    assert W.lib().wga_model_sparse_export(b._h, sc.ctypes.data_as(C.c_void_p), ss.ctypes.data_as(C.c_void_p),
                                           sk.ctypes.data_as(C.c_void_p)) == 0
    a.bins_tensor().add_(b.bins_tensor())  # the all-reduce
    assert W.lib().wga_model_sparse_merge(a._h, sc.ctypes.data_as(C.c_void_p), ss.ctypes.data_as(C.c_void_p),
                                          sk.ctypes.data_as(C.c_void_p), C.c_uint64(n)) == 0
    tables, oc, _ = a.build()
    og = O.OracleGraph()
    roc, _ = og.build_model(comps, syms)
    assert_tables_equal(tables, og.tables())
    assert np.allclose(oc, roc, rtol=1e-9)


def test_symbol_wider_than_48_bits_is_rejected(W, gpu):
    mb = W.ANSModel4EncoderBuilder()
    with pytest.raises(W.WgaError):  # "Symbol can't be bigger than u48::MAX" (model4encoder_builder.rs:68-70)
        mb.push_symbols([0, 0], [5, 1 << 48])


@pytest.mark.parametrize("params", [(7, 3, 4), (7, 3, 2), (16, 1 << 30, 4)])
def test_store_matches_oracle_store(W, O, gpu, tmp_path, params):
    """ANSBvGraph::store end to end (host BvComp + GPU model build + host ANS encode + writers) produces
    the same .ans/.pointers/.states content as the oracle's restatement, and decodes back on the GPU."""
    off, succ = random_graph(np.random.default_rng(12), 20000, 10)
    base = str(tmp_path / "g")
    W.ANSBvGraph.store_csr(off, succ, base, *params)
    og = O.OracleGraph.store_csr(off, succ, *params)
    lg = O.OracleGraph.load(base)
    assert lg.info() == og.info()
    assert (lg.stream() == og.stream()).all()
    st, pt = og.phases()
    st2, pt2 = lg.phases()
    assert (st == st2).all() and (pt == pt2).all()
    assert_tables_equal(lg.tables(), og.tables())
    g = W.ANSBvGraph.load(base)
    d_off, d_succ = g.decode_range()
    assert (d_off.cpu().numpy().astype(np.uint64) == off).all()
    assert (d_succ.cpu().numpy().view(np.uint32) == succ).all()


def test_store_chunked_parallel_roundtrip(W, O, gpu, tmp_path):
    off, succ = W.synth_graph("web", 120_000, 30.0, seed=5)
    base = str(tmp_path / "w")
    W.ANSBvGraph.store_csr(off, succ, base, 7, 3, 4, chunk_nodes=10_000, threads=8)
    g = W.ANSBvGraph.load(base)
    d_off, d_succ = g.decode_range()
    assert (d_off.cpu().numpy().astype(np.uint64) == off).all()
    assert (d_succ.cpu().numpy().view(np.uint32) == succ).all()
    o_off, o_succ, end = O.OracleGraph.load(base).decode_seq()
    assert (o_succ == succ).all() and end == (0, 65536)
