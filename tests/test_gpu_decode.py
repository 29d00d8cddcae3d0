"""Parity of the CUDA decode path (through the C ABI) against the CPU oracle.  Bit-exact."""
import numpy as np
import pytest

from conftest import open_oracle_graph, random_graph

pytestmark = pytest.mark.gpu


def gpu_csr(g, first=0, last=None):
    off, succ = g.decode_range(first, last)
    return off.cpu().numpy().astype(np.uint64), succ.cpu().numpy().view(np.uint32)


def test_native_library_is_the_one_running(W, gpu):
    assert W.cuda_available()
    before = W.kernel_launches()
    import torch
    assert torch.cuda.get_device_capability(0)[0] >= 10, "sm_100a code needs a Blackwell GPU"
    assert before >= 0


def test_golden_head_full_decode(W, O, gpu, head):
    g = W.ANSBvGraphSeq.load(head["base"])
    k0 = W.kernel_launches()
    off, succ = gpu_csr(g)
    assert W.kernel_launches() > k0
    assert (off == head["offsets"]).all()
    assert (succ == head["succ"]).all()
    # end to end through the host-buffer entry point
    off2, succ2 = g.decode_range_host()
    assert (off2 == head["offsets"]).all() and (succ2 == head["succ"]).all()


def test_sequential_load_from_ans_alone_decodes_bit_exact(W, O, gpu, head, tmp_path):
    """ANSBvGraphSeq::load + iterate with only the .ans present (sequential.rs:29-51)."""
    import shutil
    shutil.copy(head["base"] + ".ans", tmp_path / "only.ans")
    g = W.ANSBvGraphSeq.load(str(tmp_path / "only"))
    off, succ = gpu_csr(g)
    assert (off == head["offsets"]).all() and (succ == head["succ"]).all()


def test_host_entry_point_pipelined_in_chunks(W, O, gpu, head):
    """wga_upload(NULL) + wga_decode_range_host with small chunks: upload / decode / download overlap on three
    streams with double-buffered chunk outputs; result identical to the one-shot decode."""
    import ctypes as C
    g = W.ANSBvGraph.load(head["base"])
    try:
        for chunk in (4096, 7000, 29999):
            W.set_tuning(e2e_chunk=chunk)
            for first, last in ((0, 30000), (1234, 25001)):
                assert W.lib().wga_upload(g._h, C.c_void_p(0)) == 0
                off, succ = g.decode_range_host(first, last)
                assert (off == head["offsets"][first:last + 1] - head["offsets"][first]).all()
                assert (succ == head["succ"][head["offsets"][first]:head["offsets"][last]]).all()
    finally:
        W.set_tuning(reset=1)


def test_packed_tables_equal_reference_decoder_tables(W, O, gpu, head):
    """K3 parity: per slot (freq, cumul_freq, quasi_folded) == ANSModel4Decoder::new (model4decoder.rs:18-68)."""
    g = W.ANSBvGraph.load(head["base"])
    og = O.OracleGraph.load(head["base"])
    for c in range(9):
        ours = g.debug_expand_table(c)
        ref = og.decoder_table(c)
        assert ours.size == ref.size
        for f in ("freq", "cumul_freq", "quasi_folded"):
            assert (ours[f] == ref[f]).all(), (c, f)


@pytest.mark.parametrize("case", ["dummy", "folding", "zipf", "interleaved"])
def test_symbol_decoder_matches_oracle(W, O, gpu, case):
    """ANSDecoder::decode on the device, one symbol at a time (tests/compressor_tests.rs)."""
    rng = np.random.default_rng(0)
    if case == "dummy":
        comps, syms = [0] * 10, [1, 1, 1, 2, 2, 2, 3, 3, 4, 5]
    elif case == "folding":
        comps, syms = [0] * 3, [1000, 1000, 2000]
    elif case == "zipf":
        syms = np.minimum(rng.zipf(1.2, 100_000), 1 << 30)
        comps = np.zeros(syms.size, np.uint8)
    else:
        syms = np.concatenate([np.minimum(rng.zipf(1.3, 30_000), 1 << 30), rng.integers(0, 6, 30_000),
                               np.minimum(rng.zipf(1.1, 30_000), (1 << 47))])
        comps = np.concatenate([np.full(30_000, c, np.uint8) for c in (0, 2, 8)])
        p = rng.permutation(syms.size)
        syms, comps = syms[p], comps[p]
    comps = np.asarray(comps, np.uint8)
    syms = np.asarray(syms, np.uint64)
    og = O.OracleGraph()
    og.build_model(comps, syms)
    og.encode_symbols(comps, syms)
    inf = og.info()
    # a graph handle needs phases: give it a single dummy node
    g = W.open_mem(og.tables(), og.stream(), inf["state"], 1, 0, 0, 0, [inf["state"]], [inf["stream_len"]])
    out, ptr, state = g.debug_decode_symbols(comps[::-1])
    ref, rptr, rstate = og.decode_symbols(comps[::-1])
    assert (out == ref).all() and (out[::-1] == syms).all()
    assert (ptr, state) == (rptr, rstate) == (0, 65536)


GRAPH_CASES = [
    # n, mean_deg, (window, max_ref, min_interval), seed
    (6, 0, (7, 3, 2), 0),          # the 6-node dummy graph of tests/test_bvgraph.rs:23
    (2, 3, (7, 3, 4), 1),
    (2000, 8, (7, 3, 4), 2),       # CLI defaults
    (2000, 8, (7, 3, 2), 3),       # what the reference's tests use
    (3000, 12, (16, 1 << 30, 4), 4),  # "-hc": wide window, unbounded reference chains
    (2000, 8, (0, 3, 4), 5),       # no references
    (2000, 8, (7, 3, 0), 6),       # no intervals
    (2000, 8, (1, 1, 2), 7),
    (40000, 10, (7, 3, 4), 8),
]


def make_case(n, deg, seed):
    if n == 6:
        lists = [[2, 3], [5], [], [0, 1, 2], [1, 2, 3, 4, 5], [0]]
        off = np.cumsum([0] + [len(x) for x in lists]).astype(np.uint64)
        return off, np.array([x for l in lists for x in l], np.uint32)
    return random_graph(np.random.default_rng(seed), n, deg)


@pytest.mark.parametrize("n,deg,params,seed", GRAPH_CASES)
def test_graph_decode_matches_oracle(W, O, gpu, n, deg, params, seed):
    """tests/test_bvgraph.rs: successors(ANS graph) == successors(original), for every node."""
    off, succ = make_case(n, deg, seed)
    og = O.OracleGraph.store_csr(off, succ, *params)
    o_off, o_succ, end = og.decode_seq()
    assert (o_off == off).all() and (o_succ == succ).all()
    g = open_oracle_graph(W, og)
    d_off, d_succ = gpu_csr(g)
    assert (d_off == off).all()
    assert (d_succ == succ).all()


STRESS_TUNINGS = [
    dict(unit=37, k1_blocks=1, refill=1, k2_blocks=3),   # many tiny K1 units on one block; a K2 grid with long shares per block
    dict(unit=5, refill=32, k2_blocks=1, k2_batch=1),
    dict(unit=100000, k1_blocks=2, k2_blocks=5000),
    dict(unit=128, refill=6, k2_batch=32),
    dict(hub_min=1),                                      # every list that needs K2 is merged by a whole warp
    dict(hub_min=9, k2_blocks=2),
]


@pytest.mark.parametrize("tuning", range(len(STRESS_TUNINGS)))
@pytest.mark.parametrize("n,deg,params,seed", [GRAPH_CASES[2], GRAPH_CASES[4], GRAPH_CASES[6], GRAPH_CASES[8]])
def test_graph_decode_under_stress_tunings(W, O, gpu, n, deg, params, seed, tuning):
    """Same parity check with the kernel knobs changed so that small graphs exercise unit boundaries, rows that
    continue in another chunk, lane refill and grid striding."""
    off, succ = make_case(n, deg, seed)
    og = O.OracleGraph.store_csr(off, succ, *params)
    g = open_oracle_graph(W, og)
    try:
        W.set_tuning(**STRESS_TUNINGS[tuning])
        d_off, d_succ = gpu_csr(g)
        a, b = n // 3, n // 3 + n // 2
        s_off, s_succ = gpu_csr(g, a, b)
    finally:
        W.set_tuning(reset=1)
    assert (d_off == off).all()
    assert (d_succ == succ).all()
    assert (s_off == off[a:b + 1] - off[a]).all()
    assert (s_succ == succ[off[a]:off[b]]).all()


def test_long_records_take_the_cooperative_path(W, O, gpu):
    """Records with thousands of successors (power-law hubs): their residual runs are parked in the node's own
    slot (not in the row stream) and merged in place; all must equal the oracle."""
    rng = np.random.default_rng(11)
    n = 60000
    off, succ = make_case(n, 6, 77)
    lists = [succ[off[v]:off[v + 1]] for v in range(n)]

    def hub(k, runs):
        s = set(rng.integers(0, n, k).tolist())
        for _ in range(runs):
            a = int(rng.integers(0, n - 300))
            s.update(range(a, a + int(rng.integers(4, 200))))
        return np.array(sorted(s), np.uint32)

    lists[3] = hub(9000, 40)            # long, many intervals
    lists[4] = lists[3][::2].copy()     # long, will reference node 3
    lists[1000] = hub(30000, 3500)      # more intervals than the shared-memory budget of the cooperative path
    lists[1001] = hub(5000, 0)          # long, residuals only
    lists[n - 1] = hub(4500, 5)
    off2 = np.zeros(n + 1, np.uint64)
    off2[1:] = np.cumsum([len(x) for x in lists])
    succ2 = np.concatenate(lists).astype(np.uint32)
    og = O.OracleGraph.store_csr(off2, succ2, 7, 3, 4)
    g = open_oracle_graph(W, og)
    d_off, d_succ = gpu_csr(g)
    assert (d_off == off2).all()
    assert (d_succ == succ2).all()
    q = np.array([3, 4, 1000, 1001, n - 1, 7])
    q_off, q_succ = g.successors_batch(q)
    exp = np.concatenate([lists[v] for v in q])
    assert (q_succ.cpu().numpy().view(np.uint32) == exp).all()


def test_empty_and_mostly_dangling_graphs(W, O, gpu):
    """Empty graph, and graphs where almost every record is the single Outdegree symbol 0.  (A graph
    with ONLY dangling nodes has zero entropy, which the reference cannot encode: see
    test_zero_entropy_input_is_rejected.)"""
    og = O.OracleGraph.store_csr(np.zeros(1, np.uint64), np.zeros(0, np.uint32), 7, 3, 4)
    g = open_oracle_graph(W, og)
    d_off, d_succ = gpu_csr(g)
    assert d_off.tolist() == [0] and d_succ.size == 0
    for n, owners in ((50, (3, 49)), (5000, (0, 2500, 2501, 4999))):
        lists = [[] for _ in range(n)]
        for k, v in enumerate(owners):
            lists[v] = sorted({(v * 7 + j * 3) % n for j in range(k + 2)})
        off = np.cumsum([0] + [len(x) for x in lists]).astype(np.uint64)
        succ = np.array([x for l in lists for x in l], np.uint32)
        og = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
        g = open_oracle_graph(W, og)
        d_off, d_succ = gpu_csr(g)
        assert (d_off == off).all() and (d_succ == succ).all()


def test_sub_ranges_with_halo(W, O, gpu):
    """Node-range decode (what one rank of a sharded decode does): references leaving the range on the
    left are resolved by re-decoding the predecessor halo."""
    off, succ = make_case(30000, 10, 21)
    for params in ((7, 3, 4), (16, 1 << 30, 4)):
        og = O.OracleGraph.store_csr(off, succ, *params)
        g = open_oracle_graph(W, og)
        rng = np.random.default_rng(2)
        cuts = sorted(set([0, 30000] + list(rng.integers(1, 30000, 9))))
        for a, b in zip(cuts[:-1], cuts[1:]):
            d_off, d_succ = gpu_csr(g, int(a), int(b))
            assert (d_off == off[a:b + 1] - off[a]).all(), (a, b)
            assert (d_succ == succ[off[a]:off[b]]).all(), (a, b)
        d_off, d_succ = gpu_csr(g, 12345, 12346)
        assert (d_succ == succ[off[12345]:off[12346]]).all()


def test_shard_open_decodes_only_its_range(W, O, gpu, head):
    """wga_open_shard: only the stream span / phases of the shard are uploaded."""
    n = 30000
    for a, b in ((0, 7000), (7000, 19000), (19000, n)):
        # widen the resident range to the left so the halo is available
        g = W.ANSBvGraph.load(head["base"], shard=(max(0, a - 64), b))
        d_off, d_succ = gpu_csr(g, a, b)
        assert (d_off == head["offsets"][a:b + 1] - head["offsets"][a]).all()
        assert (d_succ == head["succ"][head["offsets"][a]:head["offsets"][b]]).all()
        assert g.compressed_bytes() < W.ANSBvGraph.load(head["base"], host_only=True).compressed_bytes()


def test_iter_and_successor_api(W, O, gpu, head):
    g = W.ANSBvGraph.load(head["base"])
    assert g.num_nodes() == 30000 and g.num_arcs_hint() == head["succ"].size
    g.ITER_CHUNK_NODES = 7000
    for v, s in g.iter(0, 15000):
        assert (s == head["succ"][head["offsets"][v]:head["offsets"][v + 1]]).all()


@pytest.mark.parametrize("params", [(7, 3, 4), (16, 1 << 30, 4), (0, 3, 4), (7, 3, 0)])
def test_random_access_batch_matches_oracle(W, O, gpu, params):
    """examples/bench_random_access.rs: successors(v) of random nodes == the sequential decode's lists."""
    off, succ = make_case(20000, 10, 41)
    og = O.OracleGraph.store_csr(off, succ, *params)
    g = open_oracle_graph(W, og)
    rng = np.random.default_rng(5)
    for q in (rng.integers(0, 20000, 5000), np.array([0, 19999, 0, 7, 7, 7]), np.arange(100, 160), np.zeros(0, np.int64)):
        d_off, d_succ = g.successors_batch(q)
        d_off = d_off.cpu().numpy().astype(np.uint64)
        d_succ = d_succ.cpu().numpy().view(np.uint32)
        assert d_off.size == q.size + 1 and d_off[0] == 0
        exp = [succ[off[v]:off[v + 1]] for v in q]
        assert (np.diff(d_off) == np.array([e.size for e in exp], np.uint64)).all()
        if q.size:
            assert (d_succ[:int(d_off[-1])] == np.concatenate(exp)).all()
    assert (g.successors(12345) == succ[off[12345]:off[12346]]).all()
    with pytest.raises(W.WgaError):
        g.successors_batch([20000])
    # the host-buffer entry point (numpy in, numpy out), sizing call included
    q = rng.integers(0, 20000, 3000)
    h_off, h_succ = g.successors_batch_host(q)
    exp = [succ[off[v]:off[v + 1]] for v in q]
    assert (np.diff(h_off) == np.array([e.size for e in exp], np.uint64)).all() and (h_succ == np.concatenate(exp)).all()
    with pytest.raises(W.WgaError) as e:
        g.successors_batch_host(q, succ_capacity=max(1, h_succ.size // 2))
    assert e.value.code == -6


@pytest.mark.parametrize("kind,n,deg", [("web", 300_000, 34.3), ("social", 200_000, 35.3)])
def test_synthetic_shapes_roundtrip(W, O, gpu, kind, n, deg):
    """Bench-shaped graphs at a size the oracle finishes in seconds: GPU decode == oracle decode == source."""
    off, succ = W.synth_graph(kind, n, deg, seed=0x5EED0000)
    og = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    g = open_oracle_graph(W, og)
    d_off, d_succ = gpu_csr(g)
    assert (d_off == off).all() and (d_succ == succ).all()
    o_off, o_succ, end = og.decode_seq()
    assert (o_succ == d_succ).all() and end == (0, 65536)


def test_output_buffer_too_small_is_reported_and_nothing_is_written_behind_it(W, O, gpu):
    """A d_succ (or a sub-range halo) smaller than the decoded arcs: WGA_E_WORKSPACE, and no kernel writes past the
    capacity it was given (guard words behind it stay intact)."""
    import ctypes as C
    import torch
    off, succ = make_case(20000, 12, 91)
    og = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    g = open_oracle_graph(W, og)
    arcs = succ.size
    for first, last, cap in ((0, 20000, arcs // 2), (0, 20000, 7), (5000, 15000, 100), (0, 20000, arcs - 1)):
        d_off = torch.zeros(last - first + 1, dtype=torch.int64, device="cuda")
        d_succ = torch.full((arcs + 4096,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        ws = torch.zeros(g.workspace_size(first, last), dtype=torch.uint8, device="cuda")
        rc = W.lib().wga_decode_range(g._h, C.c_uint64(first), C.c_uint64(last), C.c_void_p(d_off.data_ptr()),
                                      C.c_void_p(d_succ.data_ptr()), C.c_uint64(cap), C.c_void_p(ws.data_ptr()),
                                      C.c_uint64(ws.numel()), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert rc == -6, (first, last, cap, rc)
        assert b"need" in W.lib().wga_last_error()
        assert bool((d_succ[cap:] == 0x5A5A5A5A).all()), (first, last, cap)
    # the handle is still usable
    d_off, d_succ = gpu_csr(g)
    assert (d_succ == succ).all()


def test_host_entry_point_survives_a_chunk_much_denser_than_the_average(W, O, gpu):
    """wga_decode_range_host sizes its chunk buffers from the average degree; a chunk with many more arcs is
    retried with larger buffers instead of failing."""
    rng = np.random.default_rng(17)
    n_sparse, n_dense, deg = 150_000, 9_000, 520
    lists = [np.array([(v * 7 + 1) % (n_sparse + n_dense)], np.uint32) for v in range(n_sparse)]
    n = n_sparse + n_dense
    for v in range(n_dense):
        lists.append(np.unique(rng.integers(0, n, deg)).astype(np.uint32))
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum([x.size for x in lists])
    succ = np.concatenate(lists)
    og = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    g = open_oracle_graph(W, og)
    try:
        W.set_tuning(e2e_chunk=10_000)
        h_off, h_succ = g.decode_range_host()
    finally:
        W.set_tuning(reset=1)
    assert (h_off == off).all() and (h_succ == succ).all()


def test_corrupt_stream_is_reported_not_faulted(W, O, gpu):
    """The reference panics on corrupt input (slice index); the kernels set an error flag instead."""
    import torch
    off, succ = make_case(3000, 8, 33)
    og = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    inf = og.info()
    st, pt = og.phases()
    stream = og.stream().copy()
    rng = np.random.default_rng(0)
    stream[rng.integers(0, stream.size, stream.size // 3)] ^= 0xFFFF
    g = W.open_mem(og.tables(), stream, inf["state"], inf["n"], 7, 4, inf["arcs"], st, pt)
    d_off = torch.zeros(3001, dtype=torch.int64, device="cuda")
    d_succ = torch.zeros(succ.size + 1024, dtype=torch.int32, device="cuda")
    ws = torch.zeros(g.workspace_size(0, 3000), dtype=torch.uint8, device="cuda")
    with pytest.raises(W.WgaError) as e:
        g.decode_range_into(0, 3000, d_off, d_succ, ws, want_arcs=True)
    assert e.value.code in (-5, -6, -7)
    # the device is still healthy afterwards
    g2 = open_oracle_graph(W, og)
    o2, s2 = gpu_csr(g2)
    assert (s2 == succ).all()
