"""A second, independent restatement of the reference's ANS layer, in plain Python, written from the Rust sources
(NOT from oracle/wgo.hpp): model builder, encoder, decoder.  Test infrastructure, like the oracle.

    src/utils/ans_utils.rs:4-12                  fold()
    src/utils/data_utils.rs:15-39                scale_freqs()
    src/ans/model4encoder_builder.rs:80-327      build_models()
    src/ans/models/component_model4encoder.rs:28-34   upper bound (u32, wrapping)
    src/ans/encoder.rs:39-86                     Encoder.encode / shrink_state
    src/ans/models/model4decoder.rs:18-68        decoder tables, quasi_fold
    src/ans/decoder.rs:58-100                    Decoder.decode
    src/bvgraph/writers/bvgraph_encoder.rs:159-172   reverse replay, one phase per Outdegree symbol

Purpose: the reference cannot be built in this image (no rustc) and ships no .ans sample, so the oracle's tables and
streams cannot be compared with the reference's bytes.  What can be done is to have TWO restatements that were
written separately agree: `python tests/golden/restate_ans.py` (run in the build container) feeds the symbol stream
of the final BvComp pass of cnr-2000 (the reference's own golden graph) through this file and writes
tests/golden/cnr2000_restated.json; tests/test_oracle_pins.py compares the oracle with it.

Where the reference is not deterministic (it sums the original cost over a HashMap and sorts the symbols with an
unstable sort on the frequency only), this file takes the same decisions as the oracle and the product, which are
listed in DESIGN.md: sums in ascending symbol order, ties broken by symbol index.
"""
import hashlib
import json
import math
import os
import sys
from collections import Counter

MAX_RAW_SYMBOL = (1 << 48) - 1
MAXIMUM_FRAME_SIZE = 1 << 16
THETA = 1.0001
LOWER = 1 << 16
PARAMS = [(1, 3), (2, 2), (3, 1), (1, 4), (2, 3), (3, 2), (4, 1), (1, 5), (2, 4), (3, 3), (4, 2), (5, 1), (1, 6), (2, 5),
          (3, 4), (4, 3), (5, 2), (6, 1), (1, 7), (2, 6), (3, 5), (4, 4), (5, 3), (6, 2), (7, 1), (1, 8), (2, 7), (3, 6),
          (4, 5), (5, 4), (6, 3), (7, 2), (8, 1), (1, 9), (2, 8), (3, 7), (4, 6), (5, 5), (6, 4), (7, 3), (8, 2), (9, 1),
          (1, 10), (2, 9), (3, 8), (4, 7), (5, 6), (6, 5), (7, 4), (8, 3), (9, 2), (10, 1)]  # (fidelity, radix)


def fold(sym, radix, fidelity):
    cuts = (sym.bit_length() - fidelity) // radix
    out = (sym >> (cuts * radix)) + (((1 << radix) - 1) << (fidelity - 1)) * cuts
    if out > 0xFFFF:
        raise OverflowError("folded symbol does not fit 16 bits")
    return out


def scale_freqs(freqs, order, n, m, new_m):
    approx = list(freqs)
    ratio = new_m / m
    for index, s in enumerate(order):
        f = freqs[s]
        second = new_m / m
        scale = (n - index) * ratio / n + index * second / n
        approx[s] = max(1, int(math.floor(0.5 + scale * f)))
        new_m -= approx[s]
        m -= f
        if new_m < 0:
            return None
    return approx


def approx_cost(folded, approx, frame, fidelity, radix):
    threshold = 1 << (fidelity + radix - 1)
    offset = ((1 << radix) - 1) * (1 << (fidelity - 1))
    total = 0.0
    for s, a in enumerate(approx):
        if a == 0:
            continue
        folds = 0.0 if s < threshold else float((s - threshold) // offset + 1)
        total += (-math.log2(a / frame) + folds * radix) * float(folded[s])
    return total


def build_models(hists):
    """hists: 9 Counters raw symbol -> frequency.  -> 9 dicts (table = list of (freq, cumul, upperbound))."""
    totals = [sum(h.values()) for h in hists]
    orig = [sum(-math.log2(h[s] / totals[c]) * h[s] for s in sorted(h)) if h else 0.0 for c, h in enumerate(hists)]
    graph_cost = sum(orig)
    models = []
    for c, h in enumerate(hists):
        if not h:
            models.append(dict(frame_size=0, radix=2, fidelity=2, folding_threshold=10, folding_offset=10, table=[]))
            continue
        best = None  # (distribution, fidelity, radix, frame)
        frame_size, lowest = None, float("inf")
        for fid, rad in PARAMS:
            max_bucket = fold(MAX_RAW_SYMBOL, rad, fid)
            threshold = 1 << (fid + rad - 1)
            folded = [0] * max_bucket
            biggest = 0
            for raw, f in h.items():
                s = raw if raw < threshold else fold(raw, rad, fid)
                folded[s] += f
                biggest = max(biggest, s)
            present = [s for s, f in enumerate(folded) if f > 0]
            n = len(present)
            m = 1
            while m < n:
                m *= 2
            order = sorted(present, key=lambda s: (folded[s], s))
            while m <= MAXIMUM_FRAME_SIZE:
                dist = scale_freqs(folded, order, n, totals[c], m)
                if dist is None:
                    m *= 2
                    continue
                cost = approx_cost(folded, dist, float(m), fid, rad)
                ratio = (graph_cost + (cost - orig[c])) / graph_cost
                if ratio <= THETA:
                    if frame_size is None or m < frame_size:
                        lowest, best, frame_size = cost, (dist[:biggest + 1], fid, rad), m
                elif m == MAXIMUM_FRAME_SIZE:
                    if cost >= lowest:
                        break
                    lowest, best, frame_size = cost, (dist[:biggest + 1], fid, rad), m
                    break
                m *= 2
        assert best is not None
        dist, fid, rad = best
        log_m = frame_size.bit_length() - 1
        k = 16 - log_m if log_m > 0 else 15
        table, cumul = [], 0
        for f in dist:
            f16 = f & 0xFFFF
            table.append((f16, cumul, ((1 << (k + 16)) * f16) & 0xFFFFFFFF))
            cumul = cumul + f16
            if cumul > 0xFFFF:
                cumul = 0  # checked_add(...).unwrap_or(0)
        models.append(dict(frame_size=log_m, radix=rad, fidelity=fid, folding_threshold=1 << (fid + rad - 1),
                           folding_offset=((1 << rad) - 1) * (1 << (fid - 1)), table=table))
    return models


class Encoder:
    def __init__(self, models):
        self.models, self.stream, self.state = models, [], LOWER

    def shrink(self):
        self.stream.append(self.state & 0xFFFF)
        self.state >>= 16

    def encode(self, symbol, c):
        m = self.models[c]
        rad = m["radix"]
        if symbol >= m["folding_threshold"]:
            folds = (symbol.bit_length() - m["fidelity"]) // rad
            for _ in range(folds):
                bits = symbol & ((1 << rad) - 1)
                if 32 - self.state.bit_length() < rad:  # leading_zeros() < radix
                    self.shrink()
                self.state = ((self.state << rad) + bits) & 0xFFFFFFFF
                symbol >>= rad
            symbol += m["folding_offset"] * folds
        freq, cumul, upper = m["table"][symbol]
        if self.state >= upper:
            self.shrink()
        block = self.state // freq
        self.state = ((block << m["frame_size"]) + cumul + (self.state - block * freq)) & 0xFFFFFFFF


def decoder_tables(models):
    out = []
    for m in models:
        slots = [None] * (1 << m["frame_size"])
        last = 0
        for s, (freq, cumul, _) in enumerate(m["table"]):
            if freq == 0:
                continue
            if s < (m["folding_threshold"] & 0xFFFF):
                base, folds = s, 0
            else:
                folds = (s - m["folding_threshold"]) // m["folding_offset"] + 1
                base = (s - m["folding_offset"] * folds) << (folds * m["radix"])
            for slot in range(last, last + freq):
                slots[slot] = (freq, cumul, base, folds)
            last += freq
        out.append(slots)
    return out


class Decoder:
    def __init__(self, models, stream, state, pointer=None):
        self.models, self.tables, self.stream = models, decoder_tables(models), stream
        self.state, self.ptr = state, len(stream) if pointer is None else pointer

    def extend(self):
        self.ptr -= 1
        self.state = ((self.state << 16) | self.stream[self.ptr]) & 0xFFFFFFFF

    def decode(self, c):
        m = self.models[c]
        slot = self.state & ((1 << m["frame_size"]) - 1)
        freq, cumul, base, folds = self.tables[c][slot]
        self.state = (self.state >> m["frame_size"]) * freq + slot - cumul
        if self.state < LOWER:
            self.extend()
        acc, rad = 0, m["radix"]
        for _ in range(folds):
            if self.state < LOWER:
                self.extend()
            acc = (acc << rad) | (self.state & ((1 << rad) - 1))
            self.state >>= rad
            if self.state < LOWER:
                self.extend()
        return base | acc


def store_symbols(comps, syms):
    """The third pass of ANSBvGraph::store on a recorded (component, symbol) stream: model, reverse replay, phases."""
    hists = [Counter() for _ in range(9)]
    for c, s in zip(comps, syms):
        hists[c][s] += 1
    models = build_models(hists)
    enc = Encoder(models)
    states, pointers = [], []
    for c, s in zip(reversed(comps), reversed(syms)):
        enc.encode(s, c)
        if c == 0:  # Outdegree: the first symbol of a record, the last one encoded
            states.append(enc.state)
            pointers.append(len(enc.stream))
    return models, enc.stream, enc.state, states, pointers


def digest(models, stream, state, states, pointers):
    import numpy as np
    return dict(
        models=[[m["frame_size"], m["fidelity"], m["radix"], len(m["table"])] for m in models],
        tables_sha256=hashlib.sha256(b"".join(
            np.array([x for e in m["table"] for x in e], np.uint32).tobytes() for m in models)).hexdigest(),
        stream_bytes=2 * len(stream), final_state=int(state),
        stream_sha256=hashlib.sha256(np.array(stream, np.uint16).tobytes()).hexdigest(),
        states_sha256=hashlib.sha256(np.array(states, np.uint32).tobytes()).hexdigest(),
        pointers_sha256=hashlib.sha256(np.array(pointers, np.uint64).tobytes()).hexdigest())


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(os.path.dirname(here))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import numpy as np
    import oracle_py as O
    out = {}
    # (1) the whole golden graph of the reference, CLI defaults: symbols of the final BvComp pass (from the oracle's
    #     front end -- BvComp is webgraph's, not part of the ANS layer restated here)
    off, succ, _ = O.read_bvgraph("/root/reference/tests/data/cnr-2000/cnr-2000")
    for params in ((7, 3, 4),):
        g = O.OracleGraph.store_csr(off, succ, *params)
        comps, syms = g.trace()
        comps, syms = [int(x) for x in comps], [int(x) for x in syms]
        models, stream, state, states, pointers = store_symbols(comps, syms)
        d = digest(models, stream, state, states, pointers)
        # round trip through this file's decoder: sequential, from the final state
        dec = Decoder(models, stream, state)
        k = 200000
        assert [dec.decode(c) for c in comps[:k]] == syms[:k]
        d["symbols"] = len(syms)
        out["w%d_r%d_l%d" % params] = d
        print(params, d["models"], d["stream_bytes"], d["final_state"])
    # (2) the small known-answer inputs of tests/compressor_tests.rs (every symbol of component 0)
    rng = np.random.default_rng(7)
    small = {"dummy": [1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5] * 3,
             "folding": [int(x) for x in rng.integers(0, 1 << 20, 400)],
             "zipf": [int(x) for x in np.minimum(rng.zipf(1.2, 3000), 1 << 30)]}
    for name, seq in small.items():
        models, stream, state, states, pointers = store_symbols([0] * len(seq), seq)
        out["seq_" + name] = dict(input=seq if len(seq) < 64 else None, input_sha256=hashlib.sha256(
            np.array(seq, np.uint64).tobytes()).hexdigest(), **digest(models, stream, state, states, pointers))
        dec = Decoder(models, stream, state)
        assert [dec.decode(0) for _ in seq] == seq
    with open(os.path.join(here, "cnr2000_restated.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("written", os.path.join(here, "cnr2000_restated.json"))


if __name__ == "__main__":
    main()
