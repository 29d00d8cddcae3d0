"""Pins the CPU oracle: the reference's own tests (round-trip properties), the golden cnr-2000 files
and the SURVEY.md 8a regression anchors.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, REF_CNR

needs_ref = pytest.mark.skipif(not os.path.exists(REF_CNR + ".graph"), reason="/root/reference not mounted")


def roundtrip(O, comps, syms):
    g = O.OracleGraph()
    g.build_model(comps, syms)
    g.encode_symbols(comps, syms)
    out, ptr, state = g.decode_symbols(np.asarray(comps)[::-1])
    return g, out[::-1], ptr, state


def test_decodes_correctly_single_dummy_sequence(O):  # tests/compressor_tests.rs:14-43
    src = [1, 1, 1, 2, 2, 2, 3, 3, 4, 5]
    g, dec, ptr, state = roundtrip(O, [0] * 10, src)
    assert list(dec) == src
    assert (ptr, state) == (0, 65536)
    t = g.table(0)  # SURVEY 8a anchor
    assert (t["frame_size"], t["fidelity"], t["radix"]) == (7, 1, 3)
    assert list(t["entries"]["freq"]) == [0, 38, 38, 26, 13, 13]
    assert list(t["entries"]["cumul_freq"]) == [0, 0, 38, 76, 102, 115]
    assert g.info()["state"] == 3431671 and list(g.stream()) == [21212]


def test_decodes_correctly_dummy_sequence_with_folding(O):  # tests/compressor_tests.rs:45-76
    src = [1000, 1000, 2000]
    g, dec, ptr, state = roundtrip(O, [0] * 3, src)
    assert list(dec) == src
    t = g.table(0)
    assert (t["frame_size"], t["fidelity"], t["radix"]) == (16, 10, 1)
    assert t["entries"]["freq"][1000] == 43691
    assert (t["entries"]["freq"][1512], t["entries"]["cumul_freq"][1512]) == (21845, 43691)
    assert g.info()["state"] == 699053 and g.stream().size == 0


def zipf_symbols(seed, a, n=200_000, maximum=1 << 30):
    rng = np.random.default_rng(seed)
    x = rng.zipf(a, n)
    return np.minimum(x, maximum).astype(np.uint64)


def test_decoder_decodes_correctly_real_sequence(O):  # tests/compressor_tests.rs:78-109 (Zipf 1.2, up to 2^30)
    src = zipf_symbols(0, 1.2)
    g, dec, ptr, state = roundtrip(O, np.zeros(src.size, np.uint8), src)
    assert (dec == src).all() and (ptr, state) == (0, 65536)


def test_decodes_correctly_dummy_sequences(O):  # tests/compressor_tests.rs:111-152
    a = [1, 1, 1, 2, 2, 2, 3, 3, 4, 5]
    b = [1, 3, 3, 3, 2, 2, 3, 3, 4, 5]
    comps = [0, 2] * 10
    syms = [x for p in zip(a, b) for x in p]
    g, dec, ptr, state = roundtrip(O, comps, syms)
    assert list(dec) == syms
    assert g.info()["state"] == 41903996 and list(g.stream()) == [37316, 54892]  # SURVEY 8a anchor


def test_decodes_correctly_real_interleaved_sequences_with_different_frame_sizes(O):  # :154-213
    rng = np.random.default_rng(7)
    parts = [(0, zipf_symbols(1, 1.3, 60_000)), (2, zipf_symbols(2, 1.2, 60_000)), (4, np.random.default_rng(3).integers(0, 6, 60_000).astype(np.uint64))]
    comps = np.concatenate([np.full(s.size, c, np.uint8) for c, s in parts])
    syms = np.concatenate([s for _, s in parts])
    perm = rng.permutation(syms.size)
    comps, syms = comps[perm], syms[perm]
    g, dec, ptr, state = roundtrip(O, comps, syms)
    assert (dec == syms).all() and (ptr, state) == (0, 65536)
    assert len({g.table(c)["frame_size"] for c in (0, 2, 4)}) > 1


def test_decodes_correctly_dummy_graph(O):  # tests/test_bvgraph.rs:23-101 : BvComp::new(_, 7, 3, 2, 0)
    lists = [[2, 3], [5], [], [0, 1, 2], [1, 2, 3, 4, 5], [0]]
    off = np.cumsum([0] + [len(x) for x in lists]).astype(np.uint64)
    succ = np.array([x for l in lists for x in l], np.uint32)
    g = O.OracleGraph.store_csr(off, succ, 7, 3, 2)
    for v, l in enumerate(lists):
        assert list(g.successors(v)) == l
    o2, s2, end = g.decode_seq()
    assert (o2 == off).all() and (s2 == succ).all() and end == (0, 65536)


def test_zero_entropy_input_is_rejected(O):
    """A graph whose every component is deterministic has original cost 0: `ratio` is NaN
    (model4encoder_builder.rs:166), the 2^16 fallback stores freq 65536 `as u16` == 0 (:223) and the
    reference's encoder then divides by zero (encoder.rs:72-73, UB).  The oracle raises instead."""
    with pytest.raises(RuntimeError):
        O.OracleGraph.store_csr(np.array([0, 1], np.uint64), np.array([0], np.uint32), 7, 3, 4)


def test_golden_head_fixture(O, head):
    """The committed .ans/.pointers/.states decode back to the committed CSR, sequentially and by node."""
    g = O.OracleGraph.load(head["base"])
    off, succ, end = g.decode_seq()
    assert (off == head["offsets"]).all() and (succ == head["succ"]).all() and end == (0, 65536)
    rng = np.random.default_rng(0)
    for v in rng.integers(0, g.info()["n"], 300):
        assert (g.successors(int(v)) == head["succ"][head["offsets"][v]:head["offsets"][v + 1]]).all()
    # decoding node v from its phase ends exactly at the phase of v+1 (SURVEY 8c invariant)
    states, pointers = g.phases()
    n = g.info()["n"]
    assert (np.diff(pointers.astype(np.int64)) >= 0).all() and pointers[-1] == g.info()["stream_len"]
    comps, syms = head["comps"], head["syms"]
    starts = np.flatnonzero(comps == 0)
    for v in (0, 1, 17, n - 2):
        a, b = starts[v], starts[v + 1]
        out, ptr, st = g.decode_symbols(comps[a:b], int(pointers[n - 1 - v]), int(states[n - 1 - v]))
        assert (out == syms[a:b]).all()
        assert (ptr, st) == (int(pointers[n - 2 - v]), int(states[n - 2 - v]))


@needs_ref
def test_cnr2000_golden_graph_and_ef(O):
    """BV reader vs the reference's golden files: arcs/nodes from .properties, bit offsets from .ef."""
    off, succ, bits = O.read_bvgraph(REF_CNR)
    assert len(off) - 1 == 325557 and len(succ) == 3216152
    ef = O.ef_read(REF_CNR + ".ef")
    assert (ef == bits).all()
    anchors = json.load(open(os.path.join(GOLDEN, "cnr2000_full.json")))
    assert hashlib.sha256(off.tobytes() + succ.tobytes()).hexdigest() == anchors["graph"]["csr_sha256"]


@needs_ref
@pytest.mark.parametrize("params", [(7, 3, 4), (7, 3, 2)])
def test_decodes_correctly_sequential_and_random_access_graph(O, params):
    """tests/test_bvgraph.rs:105-154 on cnr-2000 (the reference uses 7/3/2), plus the SURVEY anchors."""
    off, succ, _ = O.read_bvgraph(REF_CNR)
    g = O.OracleGraph.store_csr(off, succ, *params)
    o2, s2, end = g.decode_seq()
    assert (o2 == off).all() and (s2 == succ).all() and end == (0, 65536)
    rng = np.random.default_rng(1)
    for v in rng.integers(0, len(off) - 1, 2000):
        assert (g.successors(int(v)) == succ[off[v]:off[v + 1]]).all()
    anchors = json.load(open(os.path.join(GOLDEN, "cnr2000_full.json")))["w%d_r%d_l%d" % params]
    assert g.info()["stream_len"] * 2 == anchors["stream_bytes"] == {(7, 3, 4): 1040240, (7, 3, 2): 1044012}[params]
    assert hashlib.sha256(g.stream().tobytes()).hexdigest() == anchors["stream_sha256"]
    assert [[t["frame_size"], t["fidelity"], t["radix"], len(t["entries"])] for t in g.tables()] == anchors["models"]
    if params == (7, 3, 4):  # SURVEY 8a "chosen models"
        assert [m[:3] for m in anchors["models"]] == [[16, 5, 2], [6, 1, 3], [12, 3, 1], [13, 1, 6], [11, 1, 3],
                                                      [16, 10, 1], [11, 1, 6], [16, 10, 1], [16, 10, 1]]


# ---------------------------------------------------------------------------------------------------------------
# Two restatements that were written separately must agree: tests/golden/restate_ans.py is a plain-Python
# restatement of the reference's model builder / encoder / decoder made from the Rust sources, not from the oracle.
def _tables_sha(tables):
    parts = []
    for t in tables:
        e = t["entries"]
        a = np.empty((e.size, 3), np.uint32)
        a[:, 0], a[:, 1], a[:, 2] = e["freq"], e["cumul_freq"], e["upperbound"]
        parts.append(a.tobytes())
    return hashlib.sha256(b"".join(parts)).hexdigest()


def _restated():
    return json.load(open(os.path.join(GOLDEN, "cnr2000_restated.json")))


def test_python_restatement_and_oracle_agree_on_cnr2000_anchors():
    """Whole cnr-2000, CLI defaults: models, stream, final state and phases of the oracle's store (cnr2000_full.json,
    written by make_golden.py) equal what the independent Python restatement produced from the same symbol stream
    (cnr2000_restated.json, written by restate_ans.py)."""
    a = json.load(open(os.path.join(GOLDEN, "cnr2000_full.json")))["w7_r3_l4"]
    b = _restated()["w7_r3_l4"]
    for k in ("models", "stream_bytes", "final_state", "stream_sha256", "states_sha256", "pointers_sha256", "symbols"):
        assert a[k] == b[k], k


@needs_ref
def test_oracle_tables_equal_the_python_restatement_on_cnr2000(O):
    off, succ, _ = O.read_bvgraph(REF_CNR)
    g = O.OracleGraph.store_csr(off, succ, 7, 3, 4)
    assert _tables_sha(g.tables()) == _restated()["w7_r3_l4"]["tables_sha256"]


@pytest.mark.parametrize("name", ["dummy", "folding", "zipf"])
def test_oracle_equals_the_python_restatement_on_small_sequences(O, name):
    """The known-answer style inputs of tests/compressor_tests.rs: tables, stream and final state, oracle (C++) against
    the committed output of the Python restatement -- and, for the smallest one, against the restatement run live."""
    rng = np.random.default_rng(7)
    seqs = {"dummy": [1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5] * 3,
            "folding": [int(x) for x in rng.integers(0, 1 << 20, 400)],
            "zipf": [int(x) for x in np.minimum(rng.zipf(1.2, 3000), 1 << 30)]}
    seq = seqs[name]
    exp = _restated()["seq_" + name]
    assert hashlib.sha256(np.array(seq, np.uint64).tobytes()).hexdigest() == exp["input_sha256"]
    comps, syms = np.zeros(len(seq), np.uint8), np.array(seq, np.uint64)
    g = O.OracleGraph()
    g.build_model(comps, syms)
    g.encode_symbols(comps[::-1].copy(), syms[::-1].copy())  # (the store replays the symbols last to first)
    assert [[t["frame_size"], t["fidelity"], t["radix"], len(t["entries"])] for t in g.tables()] == exp["models"]
    assert _tables_sha(g.tables()) == exp["tables_sha256"]
    assert g.info()["state"] == exp["final_state"] and g.info()["stream_len"] * 2 == exp["stream_bytes"]
    assert hashlib.sha256(g.stream().tobytes()).hexdigest() == exp["stream_sha256"]
    if name == "dummy":
        import sys
        sys.path.insert(0, GOLDEN)
        import restate_ans as R
        models, stream, state, _, _ = R.store_symbols([0] * len(seq), seq)
        assert R.digest(models, stream, state, [], [])["stream_sha256"] == exp["stream_sha256"] and state == exp["final_state"]
        dec = R.Decoder(models, stream, state)
        assert [dec.decode(0) for _ in seq] == seq
