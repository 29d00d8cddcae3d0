"""webgraph-ans on B200: host-side mirror of the reference's graph API over libwgans.so.

The reference (Rust) exposes

    ANSBvGraph::load(basename)      -> BvGraph     (src/bvgraph/random_access.rs:52)
    ANSBvGraphSeq::load(basename)   -> BvGraphSeq  (src/bvgraph/sequential.rs:29)
    ANSBvGraph::store(basename, new_basename, window, max_ref_count, min_interval_length)  (:91)

and then webgraph's `num_nodes()`, `num_arcs_hint()`, `successors(v)`, `iter()`.  The same names are
kept here.  All compute goes through the C ABI in include/wga.h (ctypes); torch is used only for
device memory, streams and torch.distributed.  There is no CPU fallback: decode and model building
raise when the CUDA library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwgans.so")
_lib = None

COMPONENTS = 9
COMPONENT_NAMES = ["Outdegree", "ReferenceOffset", "BlockCount", "Blocks", "IntervalCount",
                   "IntervalStart", "IntervalLen", "FirstResidual", "Residual"]
CANON_BINS = 20480
ENTRY_DTYPE = np.dtype([("upperbound", "<u4"), ("cumul_freq", "<u2"), ("freq", "<u2")])
DECODER_ENTRY_DTYPE = np.dtype([("freq", "<u2"), ("cumul_freq", "<u2"), ("pad", "<u4"), ("quasi_folded", "<u8")])


class WgaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class _ComponentModel(C.Structure):
    _fields_ = [("table", C.c_void_p), ("table_len", C.c_uint64), ("frame_size", C.c_uint64),
                ("radix", C.c_uint64), ("fidelity", C.c_uint64), ("folding_threshold", C.c_uint64),
                ("folding_offset", C.c_uint64)]


class _PreludeView(C.Structure):
    _fields_ = [("tables", _ComponentModel * COMPONENTS), ("stream", C.c_void_p), ("stream_len", C.c_uint64),
                ("state", C.c_uint32), ("number_of_nodes", C.c_uint64), ("compression_window", C.c_uint64),
                ("min_interval_length", C.c_uint64), ("number_of_arcs", C.c_uint64), ("states", C.c_void_p),
                ("pointers", C.c_void_p)]


def lib():
    """The C-ABI library.  Built by `make -C webgraph-ans-rs_b200/csrc` / __graft_entry__.build()."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise WgaError(-4, f"{_LIB_PATH} is missing: build the CUDA extension first (no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.wga_last_error.restype = C.c_char_p
        for name in ("wga_num_nodes", "wga_num_arcs", "wga_window", "wga_min_interval_length", "wga_stream_len",
                     "wga_compressed_bytes", "wga_decode_workspace_size", "wga_successors_workspace_size",
                     "wga_kernel_launches", "wga_symbols_len", "wga_model_sparse_count", "wga_upload_bytes",
                     "wga_last_halo_nodes"):
            getattr(L, name).restype = C.c_uint64
        L.wga_model_bins.restype = C.c_void_p
        L.wga_symbols_components.restype = C.c_void_p
        L.wga_symbols_values.restype = C.c_void_p
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise WgaError(rc, lib().wga_last_error().decode(errors="replace"))


def _np(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def cuda_available():
    return bool(lib().wga_cuda_available())


def set_tuning(**kw):
    """Decode-kernel tuning knobs (wga_debug_set_tuning); set_tuning(reset=1) restores the defaults."""
    for k, v in kw.items():
        _chk(lib().wga_debug_set_tuning(k.encode(), C.c_uint64(int(v))))


def kernel_launches():
    return int(lib().wga_kernel_launches())


def _torch():
    import torch
    return torch


def _tables_to_view(tables, keep):
    arr = (_ComponentModel * COMPONENTS)()
    for c, t in enumerate(tables):
        e = np.ascontiguousarray(t["entries"]).view(np.uint8)
        keep.append(e)
        arr[c] = _ComponentModel(e.ctypes.data if e.size else None, e.size // 8, t["frame_size"], t["radix"],
                                 t["fidelity"], t["folding_threshold"], t["folding_offset"])
    return arr


def _tables_from_view(arr):
    out = []
    for c in range(COMPONENTS):
        m = arr[c]
        n = int(m.table_len)
        ent = np.zeros(n, ENTRY_DTYPE)
        if n:
            C.memmove(ent.ctypes.data, m.table, n * 8)
        out.append(dict(entries=ent, frame_size=int(m.frame_size), radix=int(m.radix), fidelity=int(m.fidelity),
                        folding_threshold=int(m.folding_threshold), folding_offset=int(m.folding_offset)))
    return out


class BvGraph:
    """What ANSBvGraph::load / ANSBvGraphSeq::load return: an immutable ANS-compressed graph resident
    in HBM.  `successors(v)` / `iter()` mirror webgraph's BvGraph / BvGraphSeq."""

    ITER_CHUNK_NODES = 1 << 20

    def __init__(self, handle):
        self._h = handle
        self._ws = None

    def close(self):
        if self._h is not None:
            lib().wga_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- webgraph's accessors
    def num_nodes(self):
        return int(lib().wga_num_nodes(self._h))

    def num_arcs_hint(self):
        return int(lib().wga_num_arcs(self._h))

    def num_arcs(self):
        return self.num_arcs_hint()

    def compression_window(self):
        return int(lib().wga_window(self._h))

    def min_interval_length(self):
        return int(lib().wga_min_interval_length(self._h))

    def compressed_bytes(self):
        return int(lib().wga_compressed_bytes(self._h))

    def prelude(self):
        """Host view of the loaded files: tables, stream, state, phases."""
        v = _PreludeView()
        _chk(lib().wga_prelude(self._h, C.byref(v)))
        n = int(v.number_of_nodes)
        stream = np.zeros(int(v.stream_len), np.uint16)
        if stream.size:
            C.memmove(stream.ctypes.data, v.stream, stream.size * 2)
        states = pointers = None
        if v.states:
            states = np.zeros(n, np.uint32)
            pointers = np.zeros(n, np.uint64)
            if n:
                C.memmove(states.ctypes.data, v.states, n * 4)
                C.memmove(pointers.ctypes.data, v.pointers, n * 8)
        return dict(tables=_tables_from_view(v.tables), stream=stream, state=int(v.state), number_of_nodes=n,
                    compression_window=int(v.compression_window), min_interval_length=int(v.min_interval_length),
                    number_of_arcs=int(v.number_of_arcs), states=states, pointers=pointers)

    # ---- bulk decode (graph.iter() for a whole node range, on the GPU)
    def workspace_size(self, first, last):
        return int(lib().wga_decode_workspace_size(self._h, C.c_uint64(first), C.c_uint64(last)))

    def decode_range_into(self, first, last, offsets, succ, workspace, stream=None, want_arcs=False):
        """offsets: cuda int64/uint64 tensor [last-first+1]; succ: cuda int32/uint32 tensor; workspace: cuda uint8."""
        torch = _torch()
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        arcs = C.c_uint64(0)
        _chk(lib().wga_decode_range(self._h, C.c_uint64(first), C.c_uint64(last), C.c_void_p(offsets.data_ptr()),
                                    C.c_void_p(succ.data_ptr()), C.c_uint64(succ.numel()),
                                    C.c_void_p(workspace.data_ptr()), C.c_uint64(workspace.numel()),
                                    C.byref(arcs) if want_arcs else None, C.c_void_p(st)))
        return arcs.value if want_arcs else None

    def outdegree_offsets(self, first=0, last=None):
        """Exclusive prefix sum of the outdegrees of [first,last) as a cuda int64 tensor [last-first+1]."""
        torch = _torch()
        last = self.num_nodes() if last is None else last
        ws = torch.empty(self.workspace_size(first, last), dtype=torch.uint8, device="cuda")
        off = torch.empty(last - first + 1, dtype=torch.int64, device="cuda")
        _chk(lib().wga_outdegrees(self._h, C.c_uint64(first), C.c_uint64(last), C.c_void_p(off.data_ptr()),
                                  C.c_void_p(ws.data_ptr()), C.c_uint64(ws.numel()),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return off

    def decode_range(self, first=0, last=None):
        """-> (offsets int64 cuda [n+1], successors int32 cuda [arcs]) of nodes [first,last)."""
        torch = _torch()
        last = self.num_nodes() if last is None else last
        off = self.outdegree_offsets(first, last)
        arcs = int(off[-1].item())
        succ = torch.empty(max(arcs, 1), dtype=torch.int32, device="cuda")
        ws = torch.empty(self.workspace_size(first, last), dtype=torch.uint8, device="cuda")
        self.decode_range_into(first, last, off, succ, ws)
        return off, succ[:arcs]

    def decode_range_host(self, first=0, last=None, succ_capacity=None):
        """End-to-end C-ABI call with HOST buffers -> (offsets u64, successors u32) numpy arrays."""
        last = self.num_nodes() if last is None else last
        cap = self.num_arcs_hint() if succ_capacity is None else succ_capacity
        off = np.zeros(last - first + 1, np.uint64)
        succ = np.zeros(max(cap, 1), np.uint32)
        arcs = C.c_uint64(0)
        _chk(lib().wga_decode_range_host(self._h, C.c_uint64(first), C.c_uint64(last), _np(off), _np(succ),
                                         C.c_uint64(cap), C.byref(arcs)))
        return off, succ[:arcs.value]

    # ---- webgraph's iteration API
    def iter(self, first=0, last=None):
        """Yields (node, successors) in node order, like BvGraphSeq::iter()."""
        last = self.num_nodes() if last is None else last
        a = first
        while a < last:
            b = min(last, a + self.ITER_CHUNK_NODES)
            off, succ = self.decode_range(a, b)
            off = off.cpu().numpy()
            succ = succ.cpu().numpy().view(np.uint32)
            for v in range(a, b):
                yield v, succ[off[v - a]:off[v - a + 1]]
            a = b

    __iter__ = iter

    def successors_batch(self, nodes):
        """-> (offsets int64 cuda [q+1], successors int32 cuda) for a batch of node ids."""
        torch = _torch()
        nodes_t = torch.as_tensor(np.ascontiguousarray(nodes, np.int64), device="cuda")
        q = nodes_t.numel()
        ws_bytes = int(lib().wga_successors_workspace_size(self._h, C.c_uint64(q), C.c_uint64(0)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        off = torch.empty(q + 1, dtype=torch.int64, device="cuda")
        cap = C.c_uint64(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        # first call sizes the result (succ == NULL), second call fills it
        _chk(lib().wga_successors_batch(self._h, C.c_void_p(nodes_t.data_ptr()), C.c_uint64(q),
                                        C.c_void_p(off.data_ptr()), None, C.c_uint64(0),
                                        C.c_void_p(ws.data_ptr()), C.c_uint64(ws_bytes), C.byref(cap), st))
        succ = torch.empty(max(cap.value, 1), dtype=torch.int32, device="cuda")
        ws_bytes = int(lib().wga_successors_workspace_size(self._h, C.c_uint64(q), C.c_uint64(cap.value)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        _chk(lib().wga_successors_batch(self._h, C.c_void_p(nodes_t.data_ptr()), C.c_uint64(q),
                                        C.c_void_p(off.data_ptr()), C.c_void_p(succ.data_ptr()),
                                        C.c_uint64(succ.numel()), C.c_void_p(ws.data_ptr()), C.c_uint64(ws_bytes),
                                        C.byref(cap), st))
        return off, succ[:cap.value]

    def successors_batch_into(self, nodes_t, offsets, succ, workspace, stream=None):
        """wga_successors_batch with caller-owned cuda tensors (no allocation on the timed path); -> arcs."""
        torch = _torch()
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        arcs = C.c_uint64(0)
        _chk(lib().wga_successors_batch(self._h, C.c_void_p(nodes_t.data_ptr()), C.c_uint64(nodes_t.numel()),
                                        C.c_void_p(offsets.data_ptr()), C.c_void_p(succ.data_ptr()),
                                        C.c_uint64(succ.numel()), C.c_void_p(workspace.data_ptr()),
                                        C.c_uint64(workspace.numel()), C.byref(arcs), st))
        return arcs.value

    def successors_batch_host(self, nodes, succ_capacity=None):
        """wga_successors_batch_host: numpy in, numpy out -> (offsets u64[q+1], successors u32).  Without a capacity the
        call is made twice (sizing, then the lists), like a caller that owns its buffers would."""
        nodes = np.ascontiguousarray(nodes, np.uint64)
        off = np.zeros(nodes.size + 1, np.uint64)
        arcs = C.c_uint64(0)
        if succ_capacity is None:
            _chk(lib().wga_successors_batch_host(self._h, _np(nodes), C.c_uint64(nodes.size), _np(off), None,
                                                 C.c_uint64(0), C.byref(arcs)))
            succ_capacity = arcs.value
        succ = np.zeros(max(1, succ_capacity), np.uint32)
        _chk(lib().wga_successors_batch_host(self._h, _np(nodes), C.c_uint64(nodes.size), _np(off), _np(succ),
                                             C.c_uint64(succ_capacity), C.byref(arcs)))
        return off, succ[:arcs.value]

    def successors_workspace_size(self, n_queries, max_total_arcs):
        return int(lib().wga_successors_workspace_size(self._h, C.c_uint64(n_queries), C.c_uint64(max_total_arcs)))

    def successors(self, v):
        """Ascending successors of node v (BvGraph::successors)."""
        off, succ = self.successors_batch([v])
        return succ.cpu().numpy().view(np.uint32)

    # ---- parity hooks
    def debug_expand_table(self, c):
        p = self.prelude()
        n = 1 << p["tables"][c]["frame_size"]
        out = np.zeros(n, DECODER_ENTRY_DTYPE)
        _chk(lib().wga_debug_expand_table(self._h, c, _np(out), C.c_uint64(n)))
        return out

    def debug_decode_symbols(self, comps, ptr=None, state=0):
        comps = np.ascontiguousarray(comps, np.uint8)
        out = np.zeros(comps.size, np.uint64)
        ep, es = C.c_uint64(0), C.c_uint32(0)
        p = C.c_uint64(0xFFFFFFFFFFFFFFFF if ptr is None else ptr)
        _chk(lib().wga_debug_decode_symbols(self._h, _np(comps), C.c_uint64(comps.size), p, C.c_uint32(state),
                                            _np(out), C.byref(ep), C.byref(es)))
        return out, ep.value, es.value


def shard_ranges(pointers, world):
    """Contiguous node ranges [(first, last)] for `world` ranks, balanced by compressed stream words.
    `pointers` is the expanded .pointers array in file order (entry i = node N-1-i, random_access.rs:225-231):
    the record of node v occupies stream words [pointers[N-1-(v+1)], pointers[N-1-v]), so a binary search on
    the (monotone) pointers splits the stream evenly."""
    ptr = np.asarray(pointers, np.uint64)
    n = ptr.size
    if n == 0:
        return [(0, 0)] * world
    total = int(ptr[-1])  # pointer of node 0 = stream length
    by_node = ptr[::-1]   # by_node[v] = start pointer of node v, non-increasing in v
    cuts = [0]
    for k in range(1, world):
        target = total - total * k // world  # words left to the right of the cut
        # first node whose start pointer is <= target
        v = int(np.searchsorted(-by_node.astype(np.int64), -int(target), side="left"))
        cuts.append(min(max(v, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def shard_resident_range(first, last, compression_window, margin_chains=64):
    """Nodes whose inputs a rank must hold to decode [first, last): the range plus a margin on the left for
    the reference chains that leave it (window x chain depth; k_halo finds the exact closure at decode time)."""
    return max(0, first - max(1, compression_window) * margin_chains), last


def _open(basename, flags=0, shard=None):
    h = C.c_void_p()
    if shard is None:
        _chk(lib().wga_open(os.fsencode(basename), flags, C.byref(h)))
    else:
        _chk(lib().wga_open_shard(os.fsencode(basename), C.c_uint64(shard[0]), C.c_uint64(shard[1]), flags,
                                  C.byref(h)))
    return BvGraph(h)


def open_mem(tables, stream, state, number_of_nodes, compression_window, min_interval_length, number_of_arcs,
             states, pointers, host_only=False):
    """Graph from host arrays (what the three files hold)."""
    keep = []
    v = _PreludeView()
    v.tables = _tables_to_view(tables, keep)
    stream = np.ascontiguousarray(stream, np.uint16)
    states = np.ascontiguousarray(states, np.uint32)
    pointers = np.ascontiguousarray(pointers, np.uint64)
    v.stream, v.stream_len, v.state = stream.ctypes.data, stream.size, state
    v.number_of_nodes, v.compression_window = number_of_nodes, compression_window
    v.min_interval_length, v.number_of_arcs = min_interval_length, number_of_arcs
    v.states, v.pointers = states.ctypes.data, pointers.ctypes.data
    h = C.c_void_p()
    _chk(lib().wga_open_mem(C.byref(v), 1 if host_only else 0, C.byref(h)))
    return BvGraph(h)


def write_files(basename, tables, stream, state, number_of_nodes, compression_window, min_interval_length,
                number_of_arcs, states, pointers):
    keep = []
    v = _PreludeView()
    v.tables = _tables_to_view(tables, keep)
    stream = np.ascontiguousarray(stream, np.uint16)
    states = np.ascontiguousarray(states, np.uint32)
    pointers = np.ascontiguousarray(pointers, np.uint64)
    v.stream, v.stream_len, v.state = stream.ctypes.data, stream.size, state
    v.number_of_nodes, v.compression_window = number_of_nodes, compression_window
    v.min_interval_length, v.number_of_arcs = min_interval_length, number_of_arcs
    v.states, v.pointers = states.ctypes.data, pointers.ctypes.data
    _chk(lib().wga_write_files(os.fsencode(basename), C.byref(v)))


class ANSBvGraph:
    """src/bvgraph/random_access.rs:31"""

    @staticmethod
    def load(basename, shard=None, host_only=False):
        return _open(basename, 1 if host_only else 0, shard)

    @staticmethod
    def store(basename, new_basename, compression_window=7, max_ref_count=3, min_interval_length=4):
        _chk(lib().wga_store(os.fsencode(basename), os.fsencode(new_basename), C.c_uint64(compression_window),
                             C.c_uint64(max_ref_count), C.c_uint64(min_interval_length)))

    @staticmethod
    def store_csr(offsets, succ, new_basename, compression_window=7, max_ref_count=3, min_interval_length=4,
                  chunk_nodes=0, threads=1):
        offsets = np.ascontiguousarray(offsets, np.uint64)
        succ = np.ascontiguousarray(succ, np.uint32)
        _chk(lib().wga_store_csr(_np(offsets), _np(succ), C.c_uint64(offsets.size - 1), os.fsencode(new_basename),
                                 C.c_uint64(compression_window), C.c_uint64(max_ref_count),
                                 C.c_uint64(min_interval_length), C.c_uint64(chunk_nodes), C.c_int(threads)))


class ANSBvGraphSeq:
    """src/bvgraph/sequential.rs:21.  Only the .ans is required, as in the reference: when .pointers / .states
    are absent the per-node phases the GPU decoder starts from are recovered at load time by one walk of the
    stream (WGA_OPEN_SEQUENTIAL)."""

    @staticmethod
    def load(basename, host_only=False):
        return _open(basename, 2 | (1 if host_only else 0))


class ANSModel4EncoderBuilder:
    """src/ans/model4encoder_builder.rs:39 -- histograms and normalisation live on the GPU."""

    def __init__(self):
        self._h = C.c_void_p()
        _chk(lib().wga_model_create(C.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                lib().wga_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def push_symbols(self, components, symbols):
        """push_symbol for arrays (numpy -> copied to the device; cuda tensors used in place)."""
        if hasattr(components, "is_cuda"):
            torch = _torch()
            assert components.is_cuda and symbols.is_cuda
            _chk(lib().wga_model_accumulate(self._h, C.c_void_p(components.data_ptr()), C.c_void_p(symbols.data_ptr()),
                                            C.c_uint64(symbols.numel()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            comps = np.ascontiguousarray(components, np.uint8)
            syms = np.ascontiguousarray(symbols, np.uint64)
            _chk(lib().wga_model_accumulate_host(self._h, _np(comps), _np(syms), C.c_uint64(syms.size)))

    def push_symbol(self, symbol, component):
        self.push_symbols([component], [symbol])

    def bins_tensor(self):
        """The dense canonical histogram (9 x 20480 int64) as a cuda tensor aliasing the builder's memory."""
        torch = _torch()
        ptr = lib().wga_model_bins(self._h)

        class _Ext:
            __cuda_array_interface__ = {"shape": (COMPONENTS * CANON_BINS,), "typestr": "<i8", "data": (ptr, False),
                                        "version": 2}
        return torch.as_tensor(_Ext(), device="cuda")

    def all_reduce(self, group=None):
        """Sums the histograms over the ranks of a torch.distributed group: dense bins by all-reduce
        (NCCL over NVLink), the sparse tail of large raw symbols by all-gather + merge."""
        import torch.distributed as dist
        torch = _torch()
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        n = int(lib().wga_model_sparse_count(self._h))
        comps = np.zeros(n, np.uint8)
        syms = np.zeros(n, np.uint64)
        cnts = np.zeros(n, np.uint64)
        _chk(lib().wga_model_sparse_export(self._h, _np(comps), _np(syms), _np(cnts)))
        dist.all_reduce(self.bins_tensor(), group=group)
        mine = (comps, syms, cnts)
        gathered = [None] * dist.get_world_size(group)
        dist.all_gather_object(gathered, mine, group=group)
        me = dist.get_rank(group)
        for r, (c, s, k) in enumerate(gathered):
            if r != me and len(s):
                c = np.ascontiguousarray(c, np.uint8)
                s = np.ascontiguousarray(s, np.uint64)
                k = np.ascontiguousarray(k, np.uint64)
                _chk(lib().wga_model_sparse_merge(self._h, _np(c), _np(s), _np(k), C.c_uint64(s.size)))

    def all_reduce_nccl(self, comm, stream=None):
        """wga_model_allreduce: the same collective behind the C ABI, on a raw ncclComm_t (see nccl_comm_from_torch)."""
        st = C.c_void_p(stream if stream is not None else _torch().cuda.current_stream().cuda_stream)
        _chk(lib().wga_model_allreduce(self._h, C.c_void_p(comm), st))

    def build(self):
        """-> (tables, original_cost[9], final_cost[9])"""
        arr = (_ComponentModel * COMPONENTS)()
        oc = np.zeros(9)
        fc = np.zeros(9)
        _chk(lib().wga_model_build(self._h, arr, _np(oc), _np(fc)))
        return _tables_from_view(arr), oc, fc


def nccl_comm_from_torch(group=None):
    """A raw NCCL communicator over the ranks of a torch.distributed group (the unique id travels through the
    group): -> ncclComm_t as int, for ANSModel4EncoderBuilder.all_reduce_nccl / wga_model_allreduce.  Free it with
    nccl_comm_destroy."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = (C.c_char * 128)()
    if rank == 0:
        _chk(lib().wga_nccl_get_unique_id(uid))
    box = [uid.raw]
    dist.broadcast_object_list(box, src=0, group=group)
    comm = C.c_void_p()
    _chk(lib().wga_nccl_comm_init(C.c_int(world), C.c_int(rank), C.c_char_p(box[0]), C.byref(comm)))
    return comm.value


def nccl_comm_destroy(comm):
    lib().wga_nccl_comm_destroy(C.c_void_p(comm))


# ---- host front end pieces (bvcomp) ---------------------------------------------------------------
def bvcomp_symbols(offsets, succ, compression_window=7, max_ref_count=3, min_interval_length=4,
                   estimator_tables=None, chunk_nodes=0, threads=1, first_node=0, gpu_costing=False):
    """(components u8, symbols u64) that BvComp writes for a CSR graph, choosing references with the
    Log2Estimator (estimator_tables=None) or the EntropyEstimator built from the given tables.
    first_node > 0: the CSR holds nodes [first_node, first_node + n) of a larger graph (one rank's share).
    gpu_costing: the candidate costs and the reference selection run on the GPU (same symbols)."""
    offsets = np.ascontiguousarray(offsets, np.uint64)
    succ = np.ascontiguousarray(succ, np.uint32)
    keep = []
    est = _tables_to_view(estimator_tables, keep) if estimator_tables is not None else None
    h = C.c_void_p()
    fn = lib().wga_bvcomp_symbols_range_gpu if gpu_costing else lib().wga_bvcomp_symbols_range
    _chk(fn(_np(offsets), _np(succ), C.c_uint64(first_node), C.c_uint64(offsets.size - 1),
                                        C.c_uint64(compression_window), C.c_uint64(max_ref_count),
                                        C.c_uint64(min_interval_length), est, C.c_uint64(chunk_nodes), C.c_int(threads),
                                        C.byref(h)))
    n = int(lib().wga_symbols_len(h))
    comps = np.zeros(n, np.uint8)
    vals = np.zeros(n, np.uint64)
    if n:
        C.memmove(comps.ctypes.data, lib().wga_symbols_components(h), n)
        C.memmove(vals.ctypes.data, lib().wga_symbols_values(h), n * 8)
    lib().wga_symbols_free(h)
    return comps, vals


def bvcomp_costs(offsets, succ, compression_window=7, min_interval_length=4, estimator_tables=None, chunk_nodes=0,
                 first_node=0, use_gpu=True):
    """Cost of every (node, reference offset) candidate record -> u64[n, window + 1] (UINT64_MAX: no candidate)."""
    offsets = np.ascontiguousarray(offsets, np.uint64)
    succ = np.ascontiguousarray(succ, np.uint32)
    keep = []
    est = _tables_to_view(estimator_tables, keep) if estimator_tables is not None else None
    n = offsets.size - 1
    out = np.zeros((n, compression_window + 1), np.uint64)
    _chk(lib().wga_debug_bvcomp_costs(_np(offsets), _np(succ), C.c_uint64(first_node), C.c_uint64(n),
                                      C.c_uint64(compression_window), C.c_uint64(min_interval_length), est,
                                      C.c_uint64(chunk_nodes), C.c_int(1 if use_gpu else 0), _np(out)))
    return out


def ans_encode(tables, components, symbols):
    """Serial ANS encode in reverse order -> (stream u16, state, states u32, pointers u64)."""
    keep = []
    arr = _tables_to_view(tables, keep)
    comps = np.ascontiguousarray(components, np.uint8)
    syms = np.ascontiguousarray(symbols, np.uint64)
    ps, pst, ppt = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ns, npn, state = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    _chk(lib().wga_ans_encode(arr, _np(comps), _np(syms), C.c_uint64(syms.size), C.byref(ps), C.byref(ns),
                              C.byref(state), C.byref(pst), C.byref(ppt), C.byref(npn)))
    stream = np.zeros(ns.value, np.uint16)
    states = np.zeros(npn.value, np.uint32)
    pointers = np.zeros(npn.value, np.uint64)
    if ns.value:
        C.memmove(stream.ctypes.data, ps, ns.value * 2)
    if npn.value:
        C.memmove(states.ctypes.data, pst, npn.value * 4)
        C.memmove(pointers.ctypes.data, ppt, npn.value * 8)
    for p in (ps, pst, ppt):
        lib().wga_free(p)
    return stream, int(state.value), states, pointers


def bvgraph_read(basename):
    """Java/webgraph BVGraph -> (offsets u64, successors u32)."""
    n, m = C.c_uint64(0), C.c_uint64(0)
    _chk(lib().wga_bvgraph_read(os.fsencode(basename), C.byref(n), C.byref(m), None, None))
    off = np.zeros(n.value + 1, np.uint64)
    succ = np.zeros(max(m.value, 1), np.uint32)
    _chk(lib().wga_bvgraph_read(os.fsencode(basename), C.byref(n), C.byref(m), _np(off), _np(succ)))
    return off, succ[:m.value]


def ef_write(path, values, u):
    values = np.ascontiguousarray(values, np.uint64)
    _chk(lib().wga_ef_write(os.fsencode(path), _np(values), C.c_uint64(values.size), C.c_uint64(u)))


def ef_read(path):
    n = C.c_uint64(0)
    _chk(lib().wga_ef_read(os.fsencode(path), C.byref(n), None))
    out = np.zeros(n.value, np.uint64)
    _chk(lib().wga_ef_read(os.fsencode(path), C.byref(n), _np(out)))
    return out


def synth_graph(kind, n_nodes, mean_degree, seed, first=0, last=None, threads=None):
    """Synthetic graph of a benchmark shape -> (offsets u64, successors u32) of nodes [first,last)."""
    last = n_nodes if last is None else last
    threads = threads or (os.cpu_count() or 1)
    k = {"web": 0, "social": 1, "coauthor": 2}.get(kind, kind)
    arcs = C.c_uint64(0)
    off = np.zeros(last - first + 1, np.uint64)
    _chk(lib().wga_synth_graph(k, C.c_uint64(n_nodes), C.c_double(mean_degree), C.c_uint64(seed), C.c_uint64(first),
                               C.c_uint64(last), C.c_int(threads), _np(off), None, C.byref(arcs)))
    succ = np.zeros(max(arcs.value, 1), np.uint32)
    _chk(lib().wga_synth_graph(k, C.c_uint64(n_nodes), C.c_double(mean_degree), C.c_uint64(seed), C.c_uint64(first),
                               C.c_uint64(last), C.c_int(threads), _np(off), _np(succ), C.byref(arcs)))
    return off, succ[:arcs.value]
