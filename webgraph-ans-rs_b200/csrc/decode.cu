// =============================================================================
//  decode.cu -- ANS decode of BvGraph components into CSR successor lists (sm_100a)
// =============================================================================
//  Replaces, for whole node ranges at once, what the reference does one symbol at a
//  time on one core:
//    webgraph BvGraphSeq::iter() / BvGraph::successors(v)  (external, un-vendored)
//      -> ANSBVGraphDecoderFactory::new_decoder(v)   src/bvgraph/factories/bvgraph_decoder_factory.rs:46-58
//      -> ANSDecoder::decode(component)              src/ans/decoder.rs:58-100
//  Pipeline (all launches on the caller's stream):
//    K0  k_outdegree      first symbol of every record from (states[N-1-v], pointers[N-1-v]) -> outdegree
//        cub scan         -> CSR offsets
//    K1  k_decode_nodes   entropy-decodes every component of every node (one node per lane).
//                         Residual gaps are prefix-summed and merged with the expanded intervals on the
//                         fly and written straight into the TAIL of the node's final CSR slot; copy-block
//                         lengths go to a small staging arena.  Nodes without a reference are final.
//    K2a k_levels         reference-chain depth of every node (depth[v] = depth[v-r]+1)
//    K2b k_resolve        per depth level: copies the masked blocks of the (finished) referenced list and
//                         merges them with the node's extras, in place.
// =============================================================================
#include <cub/cub.cuh>

#include "graph.hpp"

namespace wga {

std::atomic<uint64_t> g_kernel_launches{0};

namespace {

constexpr int TPB = 128;

struct RangeView {
  uint64_t lo;        // first decoded node (halo start)
  uint64_t first;     // first node the caller asked for
  uint32_t n;         // nodes decoded: last - lo
  uint32_t h;         // halo nodes: first - lo
  uint32_t* outdeg;   // n+1
  uint64_t* offs;     // n+1, relative to lo
  uint64_t* meta;     // n : r (16 bit) | stage offset << 16
  uint32_t* level;    // n
  uint32_t* stage;    // staging arena for copy blocks: [b, len_0, ..., len_{b-1}] per referencing node
  uint64_t stage_cap;
  unsigned long long* cursor;
  uint32_t* maxlevel;
  uint32_t* halo_succ;  // successors of halo nodes
  uint64_t halo_cap;
  uint32_t* succ;       // caller's array: successors of nodes >= first
  uint64_t succ_cap;
  uint32_t* err;
};

__device__ __forceinline__ uint32_t* node_slot(const RangeView& rv, uint32_t t) {
  uint64_t o = rv.offs[t];
  if (t < rv.h) return rv.halo_succ + o;
  return rv.succ + (o - rv.offs[rv.h]);
}

// (state, pointer) of node v: ANSBVGraphDecoderFactory::new_decoder (bvgraph_decoder_factory.rs:46-58)
__device__ __forceinline__ void load_phase(const DevGraph& g, uint64_t v, uint32_t& state, int64_t& ptr,
                                           uint32_t& err) {
  state = g.states[g.top - v];
  uint64_t p = g.ptrs[g.top - v] - g.stream_base;
  if (p > g.stream_words) { err |= ERR_CORRUPT; p = 0; }
  ptr = (int64_t)p;
}

// -------------------------------------------------------------------------------------------- K0
__global__ void __launch_bounds__(TPB) k_outdegree(DevGraph g, uint64_t lo, uint32_t n, uint32_t* outdeg,
                                                   uint32_t* err_out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  if (t == n) { outdeg[n] = 0; return; }
  uint64_t v = lo + t;
  uint32_t state, err = 0;
  int64_t ptr;
  load_phase(g, v, state, ptr, err);
  uint64_t d = ans_decode(g.tb, g.tb.lut, g.tb.ent, Outdegree, state, ptr, g.stream, err);
  if (d > 0xFFFFFFFFull) err |= ERR_SYMBOL_WIDTH;
  outdeg[t] = (uint32_t)d;
  if (err) atomicOr(err_out, err);
}

struct U32ToU64 {
  __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};

// -------------------------------------------------------------------------------------------- halo
// Contiguous closure of references leaving [first, ...) on the left: finds lo <= first such that every
// node in [lo, first + window) references a node >= lo.  One warp; each round decodes (outdegree,
// reference offset) of up to 32 not-yet-inspected nodes.
__global__ void k_halo(DevGraph g, uint64_t first, uint64_t last, uint64_t* lo_out, uint32_t* err_out) {
  const uint32_t lane = threadIdx.x;
  uint64_t lo = first;
  uint64_t chk_lo = first;
  uint64_t chk_hi = first + g.window < last ? first + g.window : last;
  uint32_t err = 0;
  while (chk_lo < chk_hi) {
    uint64_t new_lo = lo;
    for (uint64_t base = chk_lo; base < chk_hi; base += 32) {
      uint64_t v = base + lane;
      uint64_t mine = lo;
      if (v < chk_hi) {
        uint32_t state;
        int64_t ptr;
        load_phase(g, v, state, ptr, err);
        uint64_t d = ans_decode(g.tb, g.tb.lut, g.tb.ent, Outdegree, state, ptr, g.stream, err);
        if (d != 0 && g.window != 0) {
          uint64_t r = ans_decode(g.tb, g.tb.lut, g.tb.ent, ReferenceOffset, state, ptr, g.stream, err);
          if (r > v) err |= ERR_CORRUPT;
          else if (v - r < mine) mine = v - r;
        }
      }
      for (int o = 16; o; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other < mine ? other : mine;
      }
      new_lo = mine < new_lo ? mine : new_lo;
    }
    // next round inspects the newly added nodes [new_lo, lo)
    chk_lo = new_lo;
    chk_hi = lo;
    lo = new_lo;
  }
  if (lane == 0) {
    *lo_out = lo;
    if (err) atomicOr(err_out, err);
  }
}

// -------------------------------------------------------------------------------------------- K1
__global__ void __launch_bounds__(TPB) k_decode_nodes(DevGraph g, RangeView rv) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rv.n) return;
  const uint64_t v = rv.lo + t;
  uint32_t state, err = 0;
  int64_t ptr;
  load_phase(g, v, state, ptr, err);
  const uint16_t* lut = g.tb.lut;
  const uint2* ent = g.tb.ent;
#define DEC(c) ans_decode(g.tb, lut, ent, (c), state, ptr, g.stream, err)
  const uint32_t d = (uint32_t)DEC(Outdegree);
  uint64_t meta = 0;
  if (d != 0) {
    uint32_t r = g.window ? (uint32_t)DEC(ReferenceOffset) : 0u;
    uint32_t copied = 0;
    if (r != 0) {
      if (r > t) {
        err |= ERR_RANGE;
        r = 0;
      } else {
        const uint32_t dref = rv.outdeg[t - r];
        const uint64_t b64 = DEC(BlockCount);
        const uint32_t b = (uint32_t)b64;
        if (b64 > (uint64_t)dref + 1) err |= ERR_CORRUPT;
        unsigned long long so = atomicAdd(rv.cursor, (unsigned long long)b + 1ull);
        bool fits = so + b + 1 <= rv.stage_cap;
        if (!fits) err |= ERR_WORKSPACE;
        if (fits) rv.stage[so] = b;
        uint64_t pos = 0;
        for (uint32_t k = 0; k < b && !(err & ERR_CORRUPT); ++k) {
          uint64_t x = DEC(Blocks);
          uint64_t len = k == 0 ? x : x + 1;
          if (fits) rv.stage[so + 1 + k] = (uint32_t)len;
          if ((k & 1) == 0) copied += (uint32_t)len;
          pos += len;
          if (pos > dref) err |= ERR_CORRUPT;
        }
        if ((b & 1) == 0 && pos <= dref) copied += dref - (uint32_t)pos;
        meta = (uint64_t)r | ((uint64_t)so << 16);
      }
    }
    if (copied > d) { err |= ERR_CORRUPT; copied = d; }
    const uint32_t extras = d - copied;
    if (extras != 0 && !(err & (ERR_CORRUPT | ERR_WORKSPACE))) {
      uint32_t* out = node_slot(rv, t) + copied;
      uint32_t ni = 0;
      if (g.min_interval != 0) {
        uint64_t x = DEC(IntervalCount);
        if (2 * x > extras) { err |= ERR_CORRUPT; x = 0; }
        ni = (uint32_t)x;
      }
      // intervals: (start,len) pairs are parked at the end of the extras region until merged
      uint32_t* park = out + (extras - 2 * ni);
      uint32_t nres = extras;
      int64_t prev_end = 0;
      for (uint32_t k = 0; k < ni; ++k) {
        uint64_t x = DEC(IntervalStart);
        int64_t start = k == 0 ? (int64_t)v + nat2int(x) : prev_end + 1 + (int64_t)x;
        uint64_t len = DEC(IntervalLen) + g.min_interval;
        if (len > nres || len < 2 || start < 0) { err |= ERR_CORRUPT; len = 0; ni = k; break; }
        park[2 * k] = (uint32_t)start;
        park[2 * k + 1] = (uint32_t)len;
        prev_end = start + (int64_t)len;
        nres -= (uint32_t)len;
      }
      uint32_t q = 0, cur = 0;
      uint32_t is = 0xFFFFFFFFu, il = 0;
      if (ni) { is = park[0]; il = park[1]; }
      int64_t prev = 0;
      for (uint32_t k = 0; k < nres; ++k) {
        uint64_t x = k == 0 ? DEC(FirstResidual) : DEC(Residual);
        int64_t val = k == 0 ? (int64_t)v + nat2int(x) : prev + 1 + (int64_t)x;
        prev = val;
        if (val < 0 || val > 0xFFFFFFFFll) { err |= ERR_SYMBOL_WIDTH; break; }
        while (cur < ni && is < (uint32_t)val) {
          for (uint32_t e = 0; e < il; ++e) out[q++] = is + e;
          ++cur;
          if (cur < ni) { is = park[2 * cur]; il = park[2 * cur + 1]; }
        }
        out[q++] = (uint32_t)val;
      }
      while (cur < ni && !(err & ERR_SYMBOL_WIDTH)) {
        for (uint32_t e = 0; e < il; ++e) out[q++] = is + e;
        ++cur;
        if (cur < ni) { is = park[2 * cur]; il = park[2 * cur + 1]; }
      }
    }
  }
#undef DEC
  // a record that failed validation must not be resolved against its reference (K2 trusts the staging)
  rv.meta[t] = err ? 0ull : meta;
  if (err) atomicOr(rv.err, err);
}

// -------------------------------------------------------------------------------------------- K2a
__global__ void __launch_bounds__(TPB) k_levels(RangeView rv) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rv.n) return;
  uint32_t lev = 0, u = t;
  uint32_t r = (uint32_t)(rv.meta[u] & 0xFFFFu);
  while (r) {
    u -= r;
    ++lev;
    r = (uint32_t)(rv.meta[u] & 0xFFFFu);
  }
  rv.level[t] = lev;
  // one atomic per warp
  uint32_t m = lev;
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(rv.maxlevel, m);
}

// -------------------------------------------------------------------------------------------- K2b
__global__ void __launch_bounds__(TPB) k_resolve(RangeView rv, uint32_t lev) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rv.n) return;
  if (rv.level[t] != lev) return;
  const uint64_t meta = rv.meta[t];
  const uint32_t r = (uint32_t)(meta & 0xFFFFu);
  const uint64_t so = meta >> 16;
  const uint32_t u = t - r;
  const uint32_t* ref = node_slot(rv, u);
  const uint32_t dref = rv.outdeg[u];
  const uint32_t d = rv.outdeg[t];
  uint32_t* dst = node_slot(rv, t);
  const uint32_t b = rv.stage[so];
  const uint32_t* bl = rv.stage + so + 1;
  uint32_t copied = 0, pos = 0;
  for (uint32_t k = 0; k < b; ++k) {
    uint32_t len = bl[k];
    if ((k & 1) == 0) copied += len;
    pos += len;
  }
  if ((b & 1) == 0) copied += dref - pos;
  const uint32_t ne = d - copied;
  const uint32_t* ext = dst + copied;
  uint32_t p = 0, e = 0;
  pos = 0;
  for (uint32_t k = 0; k <= b; ++k) {
    uint32_t len;
    if (k < b) len = bl[k];
    else len = dref - pos;  // implicit tail block
    if ((k & 1) == 0) {
      for (uint32_t i = 0; i < len; ++i) {
        uint32_t c = ref[pos + i];
        while (e < ne && ext[e] < c) dst[p++] = ext[e++];
        dst[p++] = c;
      }
    }
    pos += len;
  }
  // remaining extras are already in place (p == copied + e)
}

// -------------------------------------------------------------------------------------------- debug kernels
__global__ void k_expand_table(DevTables tb, int c, uint32_t n_slots, uint4* out) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  uint32_t j = tb.lut[tb.lut_off[c] + (slot >> tb.shift[c])];
  uint2 e = tb.ent[tb.ent_off[c] + j];
  while (slot - (e.x & 0xFFFFu) >= (e.x >> 16)) {
    ++j;
    e = tb.ent[tb.ent_off[c] + j];
  }
  uint32_t folds = e.y >> 16;
  uint4 o;
  if (folds == 0xFFFFu) {  // unused slot: DecoderModelEntry::default()
    o = make_uint4(0, 0, 0, 0);
  } else {
    uint64_t q = ((uint64_t)(e.y & 0xFFFFu) << (folds * tb.R[c])) | ((uint64_t)folds << 48);
    o.x = (e.x >> 16) | ((e.x & 0xFFFFu) << 16);  // u16 freq, u16 cumul
    o.y = 0;
    o.z = (uint32_t)q;
    o.w = (uint32_t)(q >> 32);
  }
  out[slot] = o;
}

__global__ void k_decode_symbols(DevGraph g, const uint8_t* comps, uint64_t n, int64_t ptr, uint32_t state,
                                 uint64_t* out, uint64_t* end) {
  if (threadIdx.x || blockIdx.x) return;
  uint32_t err = 0;
  for (uint64_t i = 0; i < n; ++i)
    out[i] = ans_decode(g.tb, g.tb.lut, g.tb.ent, comps[i], state, ptr, g.stream, err);
  end[0] = (uint64_t)ptr;
  end[1] = state;
  end[2] = err;
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct WorkspacePlan {
  uint64_t off_outdeg, off_offs, off_meta, off_level, off_scalars, off_cub, off_halo, off_stage;
  uint64_t cub_bytes, halo_cap, fixed_bytes;
};

WorkspacePlan plan_workspace(uint64_t n) {
  WorkspacePlan p{};
  uint64_t o = 0;
  p.off_scalars = o; o += 256;
  p.off_outdeg = o; o = align_up(o + 4 * (n + 1), 256);
  p.off_offs = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_meta = o; o = align_up(o + 8 * n, 256);
  p.off_level = o; o = align_up(o + 4 * n, 256);
  size_t cub_bytes = 0;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, it, (uint64_t*)nullptr, (int64_t)(n + 1));
  p.cub_bytes = cub_bytes;
  p.off_cub = o; o = align_up(o + cub_bytes, 256);
  p.halo_cap = 1u << 20;  // successors of halo nodes (u32 each)
  p.off_halo = o; o = align_up(o + 4 * p.halo_cap, 256);
  p.off_stage = o;
  p.fixed_bytes = o;
  return p;
}

}  // namespace

uint64_t decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last) {
  uint64_t n = last - first + 4096;  // room for a halo
  WorkspacePlan p = plan_workspace(n);
  double frac = g->prelude.number_of_nodes ? (double)(last - first) / (double)g->prelude.number_of_nodes : 1.0;
  uint64_t arcs_est = (uint64_t)((double)g->prelude.number_of_arcs * frac * 1.25) + (1u << 20);
  uint64_t stage_cap = n + arcs_est / 2;
  return p.fixed_bytes + 4 * stage_cap;
}

static void check_device_error(wga_graph* g, cudaStream_t st) {
  uint32_t herr = 0;
  WGA_CUDA(cudaMemcpyAsync(&herr, g->d_err, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  if (herr) {
    WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    if (herr & ERR_WORKSPACE) throw Error(WGA_E_WORKSPACE, "decode: block staging arena too small; pass a larger workspace");
    if (herr & ERR_RANGE) throw Error(WGA_E_CORRUPT, "decode: a reference leaves the decoded range");
    if (herr & ERR_SYMBOL_WIDTH) throw Error(WGA_E_UNSUPPORTED, "decode: a decoded value does not fit 32 bits");
    throw Error(WGA_E_CORRUPT, "decode: inconsistent stream or tables");
  }
}

void outdegrees(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, void* ws, uint64_t ws_bytes,
                cudaStream_t st) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (first > last || last > g->res_last || first < g->res_first) throw Error(WGA_E_ARG, "range outside the resident nodes");
  uint64_t n = last - first;
  WorkspacePlan p = plan_workspace(n);
  if (ws_bytes < p.fixed_bytes) throw Error(WGA_E_WORKSPACE, "workspace too small");
  uint8_t* w = (uint8_t*)ws;
  uint32_t* outdeg = (uint32_t*)(w + p.off_outdeg);
  k_outdegree<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, first, (uint32_t)n, outdeg, g->d_err);
  count_launch();
  size_t cb = p.cub_bytes;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(outdeg, U32ToU64());
  WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, d_offsets, (int64_t)(n + 1), st));
  count_launch(2);
  WGA_CUDA(cudaGetLastError());
}

void launch_offsets_rebase(const uint64_t* src, uint64_t base, uint64_t* dst, uint64_t n, cudaStream_t st);

static void mark(wga_graph* g, cudaStream_t st) {
  if (!g->profiling || g->n_ev >= 8) return;
  if (!g->ev[g->n_ev]) cudaEventCreate(&g->ev[g->n_ev]);
  cudaEventRecord(g->ev[g->n_ev++], st);
}

void decode_range(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, uint32_t* d_succ,
                  uint64_t succ_capacity, void* ws, uint64_t ws_bytes, uint64_t* h_arcs, cudaStream_t st) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (first > last || last > g->res_last || first < g->res_first) throw Error(WGA_E_ARG, "range outside the resident nodes");
  if (last - first >= 0xFFFFFFF0ull) throw Error(WGA_E_UNSUPPORTED, "range too large for one call (2^32 nodes)");
  if (first == last) {
    WGA_CUDA(cudaMemsetAsync(d_offsets, 0, 8, st));
    if (h_arcs) *h_arcs = 0;
    return;
  }
  uint8_t* w = (uint8_t*)ws;
  if (ws_bytes < 256) throw Error(WGA_E_WORKSPACE, "workspace too small");
  WGA_CUDA(cudaMemsetAsync(w, 0, 256, st));
  g->n_ev = 0;
  mark(g, st);  // 0: start
  uint64_t* d_lo = (uint64_t*)(w + 16);
  // ---- halo
  uint64_t lo = first;
  if (first > g->res_first && g->prelude.compression_window != 0) {
    k_halo<<<1, 32, 0, st>>>(g->dev, first, last, d_lo, g->d_err);
    count_launch();
    WGA_CUDA(cudaMemcpyAsync(&lo, d_lo, 8, cudaMemcpyDeviceToHost, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    if (lo < g->res_first) throw Error(WGA_E_ARG, "reference chain leaves the resident shard");
  }
  const uint64_t n = last - lo;
  WorkspacePlan p = plan_workspace(n);
  if (ws_bytes < p.fixed_bytes + 4096) throw Error(WGA_E_WORKSPACE, "workspace too small");
  RangeView rv{};
  rv.lo = lo; rv.first = first; rv.n = (uint32_t)n; rv.h = (uint32_t)(first - lo);
  rv.outdeg = (uint32_t*)(w + p.off_outdeg);
  rv.offs = rv.h ? (uint64_t*)(w + p.off_offs) : d_offsets;
  rv.meta = (uint64_t*)(w + p.off_meta);
  rv.level = (uint32_t*)(w + p.off_level);
  rv.stage = (uint32_t*)(w + p.off_stage);
  rv.stage_cap = (ws_bytes - p.off_stage) / 4;
  rv.cursor = (unsigned long long*)(w + 0);
  rv.maxlevel = (uint32_t*)(w + 8);
  rv.halo_succ = (uint32_t*)(w + p.off_halo);
  rv.halo_cap = p.halo_cap;
  rv.succ = d_succ; rv.succ_cap = succ_capacity;
  rv.err = g->d_err;
  const unsigned grid = (unsigned)((n + TPB - 1) / TPB);
  // ---- K0 + scan
  k_outdegree<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, lo, (uint32_t)n, rv.outdeg, g->d_err);
  count_launch();
  {
    size_t cb = p.cub_bytes;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(rv.outdeg, U32ToU64());
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, rv.offs, (int64_t)(n + 1), st));
    count_launch(2);
  }
  mark(g, st);  // 1: outdegrees + scan done
  // capacity check needs the totals: read back (halo arcs, range arcs)
  uint64_t tot[2] = {0, 0};
  WGA_CUDA(cudaMemcpyAsync(&tot[0], rv.offs + rv.h, 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(&tot[1], rv.offs + n, 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  if (tot[0] > rv.halo_cap) throw Error(WGA_E_WORKSPACE, "halo successors exceed the workspace");
  if (tot[1] - tot[0] > succ_capacity)
    throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(tot[1] - tot[0]) + " elements");
  // ---- K1
  k_decode_nodes<<<grid, TPB, 0, st>>>(g->dev, rv);
  count_launch();
  mark(g, st);  // 2: entropy decode done
  // ---- K2
  if (g->prelude.compression_window != 0) {
    k_levels<<<grid, TPB, 0, st>>>(rv);
    count_launch();
    mark(g, st);  // 3: levels done
    uint32_t maxlevel = 0;
    WGA_CUDA(cudaMemcpyAsync(&maxlevel, rv.maxlevel, 4, cudaMemcpyDeviceToHost, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    for (uint32_t lev = 1; lev <= maxlevel; ++lev) {
      k_resolve<<<grid, TPB, 0, st>>>(rv, lev);
      count_launch();
    }
  }
  if (rv.h)  // hand the caller offsets relative to `first`
    launch_offsets_rebase(rv.offs + rv.h, tot[0], d_offsets, last - first + 1, st);
  mark(g, st);  // last: resolve done
  WGA_CUDA(cudaGetLastError());
  check_device_error(g, st);
  if (g->profiling) {
    for (int i = 0; i < 8; ++i) g->stage_ms[i] = 0.f;
    for (int i = 1; i < g->n_ev; ++i) cudaEventElapsedTime(&g->stage_ms[i - 1], g->ev[i - 1], g->ev[i]);
  }
  if (h_arcs) *h_arcs = tot[1] - tot[0];
}

__global__ void k_offsets_rebase(const uint64_t* src, uint64_t base, uint64_t* dst, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - base;
}
void launch_offsets_rebase(const uint64_t* src, uint64_t base, uint64_t* dst, uint64_t n, cudaStream_t st) {
  k_offsets_rebase<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, base, dst, n);
  count_launch();
}

void debug_expand_table(wga_graph* g, int c, void* h_out, uint64_t n_slots) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (c < 0 || c >= WGA_COMPONENTS) throw Error(WGA_E_ARG, "bad component");
  uint64_t want = 1ull << g->prelude.tables[c].frame_size;
  if (n_slots != want) throw Error(WGA_E_ARG, "n_slots must be 2^frame_size");
  uint4* d = nullptr;
  WGA_CUDA(cudaMalloc(&d, n_slots * 16));
  k_expand_table<<<(unsigned)((n_slots + 255) / 256), 256>>>(g->dev.tb, c, (uint32_t)n_slots, d);
  count_launch();
  cudaError_t e = cudaMemcpy(h_out, d, n_slots * 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  WGA_CUDA(e);
}

void debug_decode_symbols(wga_graph* g, const uint8_t* h_comps, uint64_t n, uint64_t ptr, uint32_t state,
                          uint64_t* h_out, uint64_t* h_end_ptr, uint32_t* h_end_state) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (ptr == UINT64_MAX) { ptr = g->prelude.stream.size(); state = g->prelude.state; }
  uint8_t* dc = nullptr; uint64_t* dout = nullptr; uint64_t* dend = nullptr;
  WGA_CUDA(cudaMalloc(&dc, n ? n : 1));
  WGA_CUDA(cudaMalloc(&dout, (n ? n : 1) * 8));
  WGA_CUDA(cudaMalloc(&dend, 24));
  cudaMemcpy(dc, h_comps, n, cudaMemcpyHostToDevice);
  k_decode_symbols<<<1, 1>>>(g->dev, dc, n, (int64_t)(ptr - g->stream_base), state, dout, dend);
  count_launch();
  uint64_t end[3] = {0, 0, 0};
  cudaMemcpy(h_out, dout, n * 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(end, dend, 24, cudaMemcpyDeviceToHost);
  cudaFree(dc); cudaFree(dout); cudaFree(dend);
  WGA_CUDA(e);
  if (end[2]) throw Error(WGA_E_CORRUPT, "decode_symbols: inconsistent stream or tables");
  if (h_end_ptr) *h_end_ptr = end[0] + g->stream_base;
  if (h_end_state) *h_end_state = (uint32_t)end[1];
}

}  // namespace wga
