"""ctypes binding of the CPU oracle (oracle/libwgoracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COMPONENT_NAMES = ["Outdegree", "ReferenceOffset", "BlockCount", "Blocks", "IntervalCount",
                   "IntervalStart", "IntervalLen", "FirstResidual", "Residual"]


def build(force=False):
    so = os.path.join(_HERE, "libwgoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("wgo_capi.cpp", "wgo.hpp", "wgo_io.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libwgoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.wgo_last_error.restype = C.c_char_p
        L.wgo_graph_new.restype = C.c_void_p
        L.wgo_bv_read.restype = C.c_void_p
        L.wgo_bv_nodes.restype = C.c_uint64
        L.wgo_bv_arcs.restype = C.c_uint64
        L.wgo_trace_len.restype = C.c_uint64
        L.wgo_graph_decoder_table.restype = C.c_uint64
        L.wgo_successors.restype = C.c_int64
        L.wgo_ef_read.restype = C.c_int64
        L.wgo_fold.restype = C.c_uint16
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().wgo_last_error().decode())


def fold(sym, radix, fidelity):
    return lib().wgo_fold(C.c_uint64(sym), C.c_uint64(radix), C.c_uint64(fidelity))


def ef_read(path):
    n = lib().wgo_ef_read(path.encode(), None, C.c_uint64(0))
    if n < 0:
        raise RuntimeError(lib().wgo_last_error().decode())
    out = np.zeros(n, dtype=np.uint64)
    lib().wgo_ef_read(path.encode(), _p(out), C.c_uint64(n))
    return out


def read_bvgraph(basename):
    """-> (offsets u64[n+1], succ u32[m], bit_offsets u64[n+1]) of a Java/webgraph BVGraph."""
    h = lib().wgo_bv_read(basename.encode())
    if not h:
        raise RuntimeError(lib().wgo_last_error().decode())
    h = C.c_void_p(h)
    n, m = lib().wgo_bv_nodes(h), lib().wgo_bv_arcs(h)
    off = np.zeros(n + 1, np.uint64)
    succ = np.zeros(m, np.uint32)
    bits = np.zeros(n + 1, np.uint64)
    lib().wgo_bv_get(h, _p(off), _p(succ), _p(bits))
    lib().wgo_bv_free(h)
    return off, succ, bits


def synth_graph(kind, n_nodes, mean_degree, seed, first=0, last=None, threads=None):
    """Synthetic benchmark graph (tools/synth/synth_graph.hpp) -> (offsets u64, successors u32) of nodes [first,last)."""
    last = n_nodes if last is None else last
    threads = threads or (os.cpu_count() or 1)
    k = {"web": 0, "social": 1, "coauthor": 2}.get(kind, kind)
    arcs = C.c_uint64(0)
    off = np.zeros(last - first + 1, np.uint64)
    _chk(lib().wgo_synth_graph(k, C.c_uint64(n_nodes), C.c_double(mean_degree), C.c_uint64(seed), C.c_uint64(first),
                               C.c_uint64(last), C.c_int(threads), _p(off), None, C.byref(arcs)))
    succ = np.zeros(max(arcs.value, 1), np.uint32)
    _chk(lib().wgo_synth_graph(k, C.c_uint64(n_nodes), C.c_double(mean_degree), C.c_uint64(seed), C.c_uint64(first),
                               C.c_uint64(last), C.c_int(threads), _p(off), _p(succ), C.byref(arcs)))
    return off, succ[:arcs.value]


def synth_degrees(kind, n_nodes, mean_degree, seed, threads=None):
    """Number of arcs of a synthetic benchmark graph (degrees only)."""
    threads = threads or (os.cpu_count() or 1)
    k = {"web": 0, "social": 1, "coauthor": 2}.get(kind, kind)
    arcs = C.c_uint64(0)
    _chk(lib().wgo_synth_graph(k, C.c_uint64(n_nodes), C.c_double(mean_degree), C.c_uint64(seed), C.c_uint64(0),
                               C.c_uint64(n_nodes), C.c_int(threads), None, None, C.byref(arcs)))
    return arcs.value


class OracleGraph:
    """An ANS graph (Prelude + phases) held by the oracle."""

    def __init__(self):
        self.h = C.c_void_p(lib().wgo_graph_new())

    def __del__(self):
        try:
            lib().wgo_graph_free(self.h)
        except Exception:
            pass

    # ---- construction
    @classmethod
    def load(cls, basename, random_access=True):
        g = cls()
        _chk(lib().wgo_graph_load(g.h, basename.encode(), int(random_access)))
        return g

    @classmethod
    def store_csr(cls, offsets, succ, window, max_ref_count, min_interval):
        g = cls()
        offsets = np.ascontiguousarray(offsets, np.uint64)
        succ = np.ascontiguousarray(succ, np.uint32)
        _chk(lib().wgo_store_csr(g.h, _p(offsets), _p(succ), C.c_uint64(len(offsets) - 1), C.c_uint64(window),
                                 C.c_uint64(max_ref_count), C.c_uint64(min_interval)))
        return g

    @classmethod
    def from_arrays(cls, tables, stream, state, n, window, min_interval, arcs, states=None, pointers=None):
        """tables: list of 9 dicts {entries(u8 view of 8B entries / structured), frame_size, radix, fidelity, thr, off}"""
        g = cls()
        for c, t in enumerate(tables):
            e = np.ascontiguousarray(t["entries"]).view(np.uint8)
            _chk(lib().wgo_graph_set_table(g.h, c, _p(e), C.c_uint64(e.size // 8), C.c_uint64(t["frame_size"]),
                                           C.c_uint64(t["radix"]), C.c_uint64(t["fidelity"]),
                                           C.c_uint64(t["folding_threshold"]), C.c_uint64(t["folding_offset"])))
        stream = np.ascontiguousarray(stream, np.uint16)
        _chk(lib().wgo_graph_set_stream(g.h, _p(stream), C.c_uint64(stream.size), C.c_uint32(state)))
        lib().wgo_graph_set_meta(g.h, C.c_uint64(n), C.c_uint64(window), C.c_uint64(min_interval), C.c_uint64(arcs))
        if states is not None:
            states = np.ascontiguousarray(states, np.uint32)
            pointers = np.ascontiguousarray(pointers, np.uint64)
            _chk(lib().wgo_graph_set_phases(g.h, _p(states), _p(pointers), C.c_uint64(states.size)))
        return g

    def build_model(self, comps, syms):
        comps = np.ascontiguousarray(comps, np.uint8)
        syms = np.ascontiguousarray(syms, np.uint64)
        oc = np.zeros(9)
        fc = np.zeros(9)
        _chk(lib().wgo_model_build(self.h, _p(comps), _p(syms), C.c_uint64(syms.size), _p(oc), _p(fc)))
        return oc, fc

    def build_model_hist(self, comps, syms, counts):
        comps = np.ascontiguousarray(comps, np.uint8)
        syms = np.ascontiguousarray(syms, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint64)
        _chk(lib().wgo_model_build_hist(self.h, _p(comps), _p(syms), _p(counts), C.c_uint64(syms.size)))

    def encode_symbols(self, comps, syms):
        comps = np.ascontiguousarray(comps, np.uint8)
        syms = np.ascontiguousarray(syms, np.uint64)
        _chk(lib().wgo_encode_symbols(self.h, _p(comps), _p(syms), C.c_uint64(syms.size)))

    # ---- inspection
    def info(self):
        o = np.zeros(7, np.uint64)
        lib().wgo_graph_info(self.h, _p(o))
        return dict(n=int(o[0]), arcs=int(o[1]), window=int(o[2]), min_interval=int(o[3]), stream_len=int(o[4]),
                    state=int(o[5]), n_phases=int(o[6]))

    def table(self, c):
        o = np.zeros(6, np.uint64)
        lib().wgo_graph_table_params(self.h, c, _p(o))
        ent = np.zeros(int(o[0]), dtype=np.dtype([("upperbound", "<u4"), ("cumul_freq", "<u2"), ("freq", "<u2")]))
        lib().wgo_graph_table(self.h, c, _p(ent))
        return dict(entries=ent, frame_size=int(o[1]), radix=int(o[2]), fidelity=int(o[3]),
                    folding_threshold=int(o[4]), folding_offset=int(o[5]))

    def tables(self):
        return [self.table(c) for c in range(9)]

    def decoder_table(self, c):
        n = lib().wgo_graph_decoder_table(self.h, c, None)
        ent = np.zeros(n, dtype=np.dtype([("freq", "<u2"), ("cumul_freq", "<u2"), ("pad", "<u4"), ("quasi_folded", "<u8")]))
        lib().wgo_graph_decoder_table(self.h, c, _p(ent))
        return ent

    def stream(self):
        s = np.zeros(self.info()["stream_len"], np.uint16)
        lib().wgo_graph_stream(self.h, _p(s))
        return s

    def phases(self):
        n = self.info()["n_phases"]
        st = np.zeros(n, np.uint32)
        pt = np.zeros(n, np.uint64)
        lib().wgo_graph_phases(self.h, _p(st), _p(pt))
        return st, pt

    def trace(self):
        n = lib().wgo_trace_len(self.h)
        comps = np.zeros(n, np.uint8)
        syms = np.zeros(n, np.uint64)
        lib().wgo_trace_get(self.h, _p(comps), _p(syms))
        return comps, syms

    # ---- decode
    def decode_symbols(self, comps, ptr=None, state=0):
        comps = np.ascontiguousarray(comps, np.uint8)
        out = np.zeros(comps.size, np.uint64)
        ep = C.c_uint64(0)
        es = C.c_uint32(0)
        p = C.c_uint64(0xFFFFFFFFFFFFFFFF if ptr is None else ptr)
        _chk(lib().wgo_decode_symbols(self.h, _p(comps), C.c_uint64(comps.size), p, C.c_uint32(state), _p(out),
                                      C.byref(ep), C.byref(es)))
        return out, ep.value, es.value

    def decode_seq(self, first=0, last=None):
        """-> (offsets relative to `first`, succ u32, (end_ptr, end_state))"""
        inf = self.info()
        last = inf["n"] if last is None else last
        arcs = C.c_uint64(0)
        # count first
        _chk(lib().wgo_decode_seq(self.h, C.c_uint64(first), C.c_uint64(last), None, None, C.c_uint64(0),
                                  C.byref(arcs), None, None))
        off = np.zeros(last - first + 1, np.uint64)
        succ = np.zeros(arcs.value, np.uint32)
        ep = C.c_uint64(0)
        es = C.c_uint32(0)
        _chk(lib().wgo_decode_seq(self.h, C.c_uint64(first), C.c_uint64(last), _p(off), _p(succ),
                                  C.c_uint64(succ.size), C.byref(arcs), C.byref(ep), C.byref(es)))
        return off, succ, (ep.value, es.value)

    def successors(self, v, cap=1 << 22):
        out = np.zeros(cap, np.uint64)
        n = lib().wgo_successors(self.h, C.c_uint64(v), _p(out), C.c_uint64(cap))
        if n < 0:
            raise RuntimeError(lib().wgo_last_error().decode())
        return out[:n].copy()

    def decode_parallel(self, first, last, nthreads):
        arcs = C.c_uint64(0)
        secs = C.c_double(0)
        _chk(lib().wgo_decode_parallel(self.h, C.c_uint64(first), C.c_uint64(last), C.c_int(nthreads),
                                       C.byref(arcs), C.byref(secs)))
        return arcs.value, secs.value

    def decode_parallel_into(self, first, last, nthreads, offsets, succ_out):
        offsets = np.ascontiguousarray(offsets, np.uint64)
        assert succ_out.dtype == np.uint32 and succ_out.flags.c_contiguous
        _chk(lib().wgo_decode_parallel_into(self.h, C.c_uint64(first), C.c_uint64(last), C.c_int(nthreads),
                                            _p(offsets), _p(succ_out)))

    def random_access_bench(self, nodes):
        nodes = np.ascontiguousarray(nodes, np.uint64)
        arcs = C.c_uint64(0)
        secs = C.c_double(0)
        _chk(lib().wgo_random_access_bench(self.h, _p(nodes), C.c_uint64(nodes.size), C.byref(arcs), C.byref(secs)))
        return arcs.value, secs.value
