"""Host side of the product (formats, BvComp front end, ANS encoder, C-ABI surface) against the oracle
and the reference's golden files.  CPU only -- no compute call touches the GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, REF_CNR, ROOT, open_oracle_graph, random_graph

needs_ref = pytest.mark.skipif(not os.path.exists(REF_CNR + ".graph"), reason="/root/reference not mounted")


def test_cabi_exports_every_declared_symbol(W):
    """include/wga.h is the contract: every function it declares must be exported by libwgans.so."""
    hdr = open(os.path.join(ROOT, "include", "wga.h")).read()
    names = sorted(set(re.findall(r"\b(wga_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 30
    lib = W.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure(W, O, tmp_path):
    """No CPU fallback: without a device the decode / model entry points fail with WGA_E_CUDA."""
    if W.cuda_available():
        pytest.skip("a GPU is present")
    rng = np.random.default_rng(0)
    off, succ = random_graph(rng, 200, 5)
    g = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    with pytest.raises(W.WgaError) as e:
        open_oracle_graph(W, g)  # upload needs CUDA
    assert e.value.code == -4
    h = open_oracle_graph(W, g, host_only=True)
    assert h.num_nodes() == 200 and h.num_arcs_hint() == len(succ)
    o = np.zeros(201, np.uint64)
    s = np.zeros(len(succ) + 1, np.uint32)
    rc = W.lib().wga_decode_range_host(h._h, ctypes.c_uint64(0), ctypes.c_uint64(200), o.ctypes.data_as(ctypes.c_void_p),
                                       s.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(s.size), None)
    assert rc == -4
    with pytest.raises(W.WgaError):
        b = W.ANSModel4EncoderBuilder()
        b.push_symbols([0, 0], [1, 2])
        b.build()


def test_load_errors(W, tmp_path):
    with pytest.raises(W.WgaError) as e:
        W.ANSBvGraph.load(str(tmp_path / "missing"), host_only=True)
    assert e.value.code == -1  # anyhow I/O error in the reference
    bad = tmp_path / "bad"
    for ext in ("ans", "pointers", "states"):
        (tmp_path / f"bad.{ext}").write_bytes(b"not an epserde file at all, definitely not.....")
    with pytest.raises(W.WgaError) as e:
        W.ANSBvGraph.load(str(bad), host_only=True)
    assert e.value.code == -2


def test_golden_head_files_load(W, head):
    g = W.ANSBvGraph.load(head["base"], host_only=True)
    assert g.num_nodes() == 30000 and g.num_arcs_hint() == len(head["succ"])
    assert g.compression_window() == 7 and g.min_interval_length() == 4
    p = g.prelude()
    assert p["pointers"][-1] == p["stream"].size and (np.diff(p["pointers"].astype(np.int64)) >= 0).all()


def test_packed_tables_equal_reference_decoder_tables_on_host(W, O, head):
    """K3 parity without a GPU: the bucket + popcount lookup on the packed tables of a host-only handle gives, per
    slot, the (freq, cumul_freq, quasi_folded) of ANSModel4Decoder::new (model4decoder.rs:18-68)."""
    g = W.ANSBvGraph.load(head["base"], host_only=True)
    og = O.OracleGraph.load(head["base"])
    for c in range(9):
        ours = g.debug_expand_table(c)
        ref = og.decoder_table(c)
        assert ours.size == ref.size
        for f in ("freq", "cumul_freq", "quasi_folded"):
            assert (ours[f] == ref[f]).all(), (c, f)


def test_sequential_load_needs_only_the_ans_file(W, head, tmp_path):
    """ANSBvGraphSeq::load (sequential.rs:29-51): with .pointers and .states deleted, the phases recovered by the
    load-time walk of the stream equal the stored ones."""
    import shutil
    shutil.copy(head["base"] + ".ans", tmp_path / "only.ans")
    g = W.ANSBvGraphSeq.load(str(tmp_path / "only"), host_only=True)
    ref = W.ANSBvGraph.load(head["base"], host_only=True).prelude()
    p = g.prelude()
    assert (p["states"] == ref["states"]).all() and (p["pointers"] == ref["pointers"]).all()
    with pytest.raises(W.WgaError):  # the random-access loader still needs all three files (random_access.rs:52-82)
        W.ANSBvGraph.load(str(tmp_path / "only"), host_only=True)


def test_files_roundtrip_through_oracle_reader(W, O, tmp_path):
    """What the product writes, the oracle's independent epserde/Elias-Fano reader reads back."""
    rng = np.random.default_rng(3)
    off, succ = random_graph(rng, 3000, 8)
    g = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    inf = g.info()
    st, pt = g.phases()
    base = str(tmp_path / "g")
    W.write_files(base, g.tables(), g.stream(), inf["state"], inf["n"], 7, 4, inf["arcs"], st, pt)
    g2 = O.OracleGraph.load(base)
    assert g2.info() == inf
    assert (g2.stream() == g.stream()).all()
    st2, pt2 = g2.phases()
    assert (st2 == st).all() and (pt2 == pt).all()
    for c in range(9):
        a, b = g.table(c), g2.table(c)
        assert (a["entries"] == b["entries"]).all() and {k: a[k] for k in a if k != "entries"} == {k: b[k] for k in b if k != "entries"}
    # and the product's own reader
    h = W.ANSBvGraph.load(base, host_only=True)
    p = h.prelude()
    assert (p["stream"] == g.stream()).all() and (p["states"] == st).all() and (p["pointers"] == pt).all()


@pytest.mark.parametrize("n,step", [(1, 0), (5, 1), (4097, 3), (20000, 1), (9000, 40), (5000, 100000)])
def test_elias_fano_write_read(W, O, tmp_path, n, step):
    """EF writer -> oracle reader (linear scan) and product reader (scan + inventory select)."""
    rng = np.random.default_rng(n)
    vals = np.cumsum(rng.integers(0, 2 * step + 1, n)).astype(np.uint64)
    path = str(tmp_path / "x.ef")
    W.ef_write(path, vals, int(vals[-1]) + 1)
    assert (O.ef_read(path) == vals).all()
    assert (W.ef_read(path) == vals).all()


@needs_ref
def test_elias_fano_writer_reproduces_golden_ef_bytes(W, O, tmp_path):
    """sux 0.4.6 layout incl. the SelectAdaptConst inventory: rebuilding cnr-2000.ef is byte-exact."""
    vals = O.ef_read(REF_CNR + ".ef")
    path = str(tmp_path / "cnr.ef")
    W.ef_write(path, vals, 9318744)
    assert open(path, "rb").read() == open(REF_CNR + ".ef", "rb").read()


@needs_ref
def test_bvgraph_reader_matches_oracle(W, O):
    off, succ = W.bvgraph_read(REF_CNR)
    o_off, o_succ, _ = O.read_bvgraph(REF_CNR)
    assert (off == o_off).all() and (succ == o_succ).all()


@pytest.mark.parametrize("params", [(7, 3, 4), (7, 3, 2), (16, 1000000, 4), (0, 3, 4), (7, 3, 0), (1, 1, 2)])
def test_bvcomp_front_end_and_encoder_match_oracle(W, O, params):
    """Pass 1 (Log2Estimator) -> model1 -> pass 2 (EntropyEstimator) symbols, and the ANS stream/phases,
    equal the oracle's restatement of ANSBvGraph::store."""
    w, r, l = params
    rng = np.random.default_rng(11)
    off, succ = random_graph(rng, 4000, 9)
    c1, s1 = W.bvcomp_symbols(off, succ, w, r, l)
    m1 = O.OracleGraph()
    m1.build_model(c1, s1)
    c2, s2 = W.bvcomp_symbols(off, succ, w, r, l, estimator_tables=m1.tables())
    og = O.OracleGraph.store_csr(off, succ, w, r, l)
    oc, osym = og.trace()
    assert (oc == c2).all() and (osym == s2).all()
    stream, state, states, pointers = W.ans_encode(og.tables(), c2, s2)
    ost, opt = og.phases()
    assert (stream == og.stream()).all() and state == og.info()["state"]
    assert (states == ost).all() and (pointers == opt).all()


def test_bvcomp_chunked_is_a_valid_compression(W, O):
    """Parallel chunks (start_node = chunk start) give a different but valid symbol stream."""
    rng = np.random.default_rng(5)
    off, succ = random_graph(rng, 5000, 9)
    c, s = W.bvcomp_symbols(off, succ, 7, 3, 4, chunk_nodes=700, threads=4)
    m = O.OracleGraph()
    m.build_model(c, s)
    stream, state, states, pointers = W.ans_encode(m.tables(), c, s)
    g = O.OracleGraph.from_arrays(m.tables(), stream, state, 5000, 7, 4, len(succ), states, pointers)
    o2, s2, end = g.decode_seq()
    assert (o2 == off).all() and (s2 == succ).all() and end == (0, 65536)


@pytest.mark.parametrize("kind,deg", [("web", 30.0), ("social", 20.0)])
def test_synthetic_graphs_are_simple_and_range_independent(W, kind, deg):
    n = 20000
    off, succ = W.synth_graph(kind, n, deg, seed=42, threads=3)
    assert off[0] == 0 and off[-1] == succ.size and succ.max() < n
    for v in range(0, n, 371):
        s = succ[off[v]:off[v + 1]].astype(np.int64)
        assert (np.diff(s) > 0).all()
    # any sub-range, any thread count: same lists
    o2, s2 = W.synth_graph(kind, n, deg, seed=42, first=5000, last=9000, threads=1)
    assert (s2 == succ[off[5000]:off[9000]]).all() and (o2 == off[5000:9001] - off[5000]).all()
    mean = succ.size / n
    assert 0.4 * deg < mean < 2.5 * deg
