// =============================================================================
//  decode.cu -- ANS decode of BvGraph components into CSR successor lists (sm_100a)
// =============================================================================
//  Replaces, for whole node ranges at once, what the reference does one symbol at a
//  time on one core:
//    webgraph BvGraphSeq::iter() / BvGraph::successors(v)  (external, un-vendored)
//      -> ANSBVGraphDecoderFactory::new_decoder(v)   src/bvgraph/factories/bvgraph_decoder_factory.rs:46-58
//      -> ANSDecoder::decode(component)              src/ans/decoder.rs:58-100
//  Pipeline (all launches on the caller's stream):
//    K0  k_heads       one lane per node, every lane at the same symbol: the fixed-shape head of every record from
//                      (states[N-1-v], pointers[N-1-v]) -- outdegree, reference offset, block count -- and the
//                      decoder state after it
//        cub scan      outdegrees -> CSR offsets
//    K1  k_entropy     phase one: entropy decode of the rest of every record.  Persistent kernel, one 1024-thread
//                      block per SM with the decoder tables of the six components in SHARED memory (bucket +
//                      popcount lookup, no search).  Warps are independent: each pulls units of consecutive nodes
//                      from a global counter, its lanes take the nodes one by one (ballot-ranked, no atomics) and
//                      run a per-symbol state machine; every busy lane decodes ONE symbol per iteration and the
//                      32 produced words (cumulative copy-block ends, interval count, interval starts / lengths,
//                      prefix-summed residuals) go out as ONE coalesced 128-byte row.  A record is therefore a
//                      column segment (row0, lane, length) of the warp's row stream.
//    K2  k_tile        phase two: one block per tile of consecutive nodes, everything in SHARED memory.  The tile's
//                      rows arrive as one bulk copy; the few predecessor nodes referenced from outside the tile
//                      (found by walking the reference chains of the tile's own nodes) are re-resolved locally, so
//                      tiles are independent.  Nodes are bucketed by (reference-chain depth, outdegree class); per
//                      depth, every lane merges one node -- copied elements of the already finished referenced
//                      list (block mask), expanded intervals, residuals -- from / into shared memory, and the
//                      finished tile leaves as one coalesced copy.
//        k_hard_*      nodes a tile cannot hold (outdegree >= 1024, chains deeper than 8, chains leaving the
//                      look-back window) are resolved afterwards from global memory, one launch per depth.
//    Random access (wga_successors_batch) runs the same kernels on the sorted reference closure of the
//    query nodes (node-list mode) and gathers the query lists.
// =============================================================================
#include <cub/cub.cuh>

#include <mutex>

#include "graph.hpp"

namespace wga {

std::atomic<uint64_t> g_kernel_launches{0};

// run-time tuning (tests change these to exercise tile boundaries, sub-tiling and the hard-node path)
struct Tuning {
  uint32_t tile = 104;        // nodes per K2 tile = nodes per K1 unit (tile + look-back window <= 256 threads)
  uint32_t slotcap = 4608;    // K2: successors a tile keeps in shared memory (words)
  uint32_t rowcap = 96;       // K2: rows of K1 output a tile keeps in shared memory
  uint32_t dbig = 1024;       // outdegree from which a node is resolved from global memory
  uint32_t k1_blocks = 0;     // K1 grid; 0 = one block per SM
  uint32_t refill = 6;        // K1: lanes that must be free before the warp fetches new nodes
  uint32_t seg = 32;          // K2: positions per merge segment
  uint32_t dbg = 0;           // (profiling) 1: merge tasks return at once, 2: after their set-up
  uint32_t e2e_chunk = 1u << 19;  // nodes per chunk of the pipelined host entry point
};
static Tuning g_tuning;

int set_tuning(const char* key, uint64_t value) {
  std::string k(key ? key : "");
  if (k == "tile") g_tuning.tile = (uint32_t)value;
  else if (k == "slotcap") g_tuning.slotcap = (uint32_t)value;
  else if (k == "rowcap") g_tuning.rowcap = (uint32_t)value;
  else if (k == "dbig") g_tuning.dbig = (uint32_t)value;
  else if (k == "k1_blocks") g_tuning.k1_blocks = (uint32_t)value;
  else if (k == "refill") g_tuning.refill = (uint32_t)value;
  else if (k == "seg") g_tuning.seg = (uint32_t)value;
  else if (k == "dbg") g_tuning.dbg = (uint32_t)value;
  else if (k == "e2e_chunk") g_tuning.e2e_chunk = (uint32_t)value;
  else if (k == "reset") g_tuning = Tuning();
  else return WGA_E_ARG;
  return WGA_OK;
}
uint64_t tuning_e2e_chunk() { return g_tuning.e2e_chunk; }

namespace {

constexpr int TPB = 128;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t INF = 0xffffffffu;  // "stream exhausted"; successor ids are <= 0xfffffffe
constexpr uint32_t NOT_FOUND = 0xFFFFFFFFu;

// ---- K1 output: rows ------------------------------------------------------------------------------------
constexpr uint32_t CH_SHIFT = 7, CH = 1u << CH_SHIFT;  // rows per chunk (16 KB)
constexpr uint32_t MAXC = 512;                         // chunks per row stream (one stream per K1 warp)
constexpr int K1_THREADS = 1024;
constexpr uint32_t K1_WARPS = K1_THREADS / 32;
constexpr uint32_t MAX_STREAMS = 8192;
// record word of node t (uint2): x = first row of the record in its stream
//                                y = words (24 bits) | lane << 24 | flags << 29
constexpr uint32_t MF_ERR = 1u, MF_INSLOT = 2u, MF_FINAL = 4u;
constexpr uint32_t NSYM_MAX = (1u << 24) - 1;
// head word of node t (K0): reference offset in node-list positions (12 bits) | block count << 12
constexpr uint32_t RT_BITS = 12, RT_MASK = (1u << RT_BITS) - 1, B_MAX = (1u << (32 - RT_BITS)) - 1;
constexpr uint32_t DSOLO = 1024;  // residual runs at least this long are parked in the node's own slot, not in rows

struct RangeView {
  uint64_t lo;        // first decoded node (halo start)
  uint32_t n;         // nodes decoded: last - lo
  const uint32_t* nodes;  // nullptr: node t is lo + t; else a sorted, duplicate-free list of node ids (random access)
  uint32_t h;         // halo nodes: first - lo
  uint32_t* outdeg;   // n+1
  uint4* nrec;        // n : from K0: decoder (state, stream index) after the record's head, outdegree, head word
  uint64_t* offs;     // n+1, relative to lo
  uint2* meta;        // n : record word of K1
  uint8_t* hardflag;  // n : 0 resolved by its tile, 1 needs the global pass, 2 final without it
  uint32_t* hard_list;  // nodes with hardflag 1
  uint32_t* hard_lev;
  uint32_t* hard_count;
  uint32_t* maxlevel;   // deepest chain among the hard nodes
  uint32_t* rows;       // row storage: chunk c = rows[c*CH*32 ...]
  uint32_t rows_cap;    // chunks
  uint32_t* chunk_ctr;
  uint32_t* stream_chunks;  // [streams][MAXC]
  uint32_t* unit_stream;    // [units]
  uint32_t* unit_ctr;
  uint32_t unit;        // nodes per unit (= K2 tile)
  uint32_t n_units;
  uint32_t* halo_succ;  // successors of halo nodes
  uint64_t halo_cap;
  uint32_t* succ;       // caller's array: successors of nodes >= first
  uint64_t succ_cap;
  uint32_t* err;
  unsigned long long* stats;  // optional per-phase cycle counters of k_tile (WGA_K2_STATS=1), else nullptr
};

// Index of the node referenced by node t with reference offset r (r != 0).  In a sorted duplicate-free
// list the node (id - r) sits at most r positions before t.
__device__ __forceinline__ uint32_t ref_index(const uint32_t* nodes, uint32_t t, uint32_t r) {
  if (!nodes) return r <= t ? t - r : NOT_FOUND;
  const uint32_t id = nodes[t];
  if (r > id) return NOT_FOUND;
  const uint32_t target = id - r;
  uint32_t lo = t >= r ? t - r : 0u, hi = t;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (nodes[mid] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < t && nodes[lo] == target) ? lo : NOT_FOUND;
}

// Final list of node t, or nullptr when it would not fit the destination (error set by the caller).
__device__ __forceinline__ uint32_t* node_slot(const RangeView& rv, uint32_t t) {
  const uint64_t o = rv.offs[t], e = rv.offs[t + 1];
  if (t < rv.h) return e <= rv.halo_cap ? rv.halo_succ + o : nullptr;
  const uint64_t b = rv.offs[rv.h];
  return e - b <= rv.succ_cap ? rv.succ + (o - b) : nullptr;
}

// word k of a record that starts at (row0, lane) of row stream s
__device__ __forceinline__ const uint32_t* row_word(const RangeView& rv, uint32_t s, uint32_t row, uint32_t lane) {
  const uint32_t c = rv.stream_chunks[s * MAXC + (row >> CH_SHIFT)];
  return rv.rows + ((size_t)c * CH + (row & (CH - 1))) * 32 + lane;
}

// (state, pointer) of node v: ANSBVGraphDecoderFactory::new_decoder (bvgraph_decoder_factory.rs:46-58)
__device__ __forceinline__ void load_phase(const DevGraph& g, uint64_t v, Dec& d, uint32_t& err) {
  d.state = g.states[g.top - v];
  uint64_t p = g.ptrs[g.top - v] - g.stream_base;
  if (p > g.stream_words) { err |= ERR_CORRUPT; p = 0; }
  d.sp = (uint32_t)p;  // the resident span has < 2^32 words (checked at upload)
  dec_prime(d, g.stream);
}

// -------------------------------------------------------------------------------------------- K0
// One lane per node, every lane at the same symbol: the fixed-shape head of a record -- outdegree, reference
// offset, block count -- and the decoder state after it.  The block count is validated by K1, which knows the
// outdegree of the referenced node.
__global__ void __launch_bounds__(TPB) k_heads(DevGraph g, uint64_t lo, const uint32_t* nodes, uint32_t n,
                                               uint32_t* outdeg, uint4* nrec, uint32_t* err_out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  if (t == n) { outdeg[n] = 0; return; }
  const GlobalTables tab{g.tb.bkt, g.tb.ent};
  uint64_t v = nodes ? (uint64_t)nodes[t] : lo + t;
  uint32_t err = 0;
  Dec dc;
  load_phase(g, v, dc, err);
  uint64_t d = ans_decode(g.tb, tab, Outdegree, dc, g.stream, err);
  if (d > 0xFFFFFFFEull) { err |= ERR_SYMBOL_WIDTH; d = 0; }
  if (err) d = 0;
  outdeg[t] = (uint32_t)d;
  if (nrec) {
    uint32_t rt = 0, b = 0;
    if (d != 0 && g.window != 0) {
      const uint64_t x = ans_decode(g.tb, tab, ReferenceOffset, dc, g.stream, err);
      if (x > g.window) err |= ERR_CORRUPT;
      else if (x != 0 && !err) {
        const uint32_t ri = ref_index(nodes, t, (uint32_t)x);
        if (ri == NOT_FOUND) err |= ERR_RANGE;  // the referenced node is not part of this decode
        else {
          const uint64_t y = ans_decode(g.tb, tab, BlockCount, dc, g.stream, err);
          if (y > B_MAX) err |= (y > 0xFFFFFFFFull) ? ERR_CORRUPT : ERR_LIMIT;
          else if (!err) { rt = t - ri; b = (uint32_t)y; }
        }
      }
    }
    if (err) { rt = 0; b = 0; }
    nrec[t] = make_uint4(dc.state, dc.sp, (uint32_t)d, rt | (b << RT_BITS));
  }
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
  const GlobalTables tab{g.tb.bkt, g.tb.ent};
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
        Dec dc;
        load_phase(g, v, dc, err);
        uint64_t d = ans_decode(g.tb, tab, Outdegree, dc, g.stream, err);
        if (d != 0 && g.window != 0 && !err) {
          uint64_t r = ans_decode(g.tb, tab, ReferenceOffset, dc, g.stream, err);
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
// Shared-memory tables of the entropy kernel: all buckets of the six components it decodes, and as many
// entries per component as fit (the first ones: small symbols are the frequent ones); the rest is read from
// global memory through the same generic load.
struct K1Tables {
  uint4 cp[WGA_COMPONENTS];       // x = mask | L << 16 | R << 21, y = smem bucket offset, z = smem entry offset, w = hot entries
  uint32_t gent_off[WGA_COMPONENTS];  // global entry offset
  uint32_t bkt_words;             // uint2 elements
  uint32_t ent_words;
};

template <bool ALLHOT>
struct SmemTables {
  const uint2* bkt;   // shared
  const uint2* ent;   // shared (hot prefix of every component)
  const uint2* gent;  // global
  const uint32_t* gent_off;  // shared copy
  const uint32_t* recip_tab;  // shared: floor(65536/R)+1
  __device__ __forceinline__ uint2 bucket(uint32_t i) const { return bkt[i]; }
  __device__ __forceinline__ uint2 entry(const uint4& cp, uint32_t off, uint32_t j) const {
    if (ALLHOT || j < cp.w) return ent[off + j];
    return __ldg(gent + gent_off[(cp.x >> 26)] + j);
  }
  __device__ __forceinline__ uint32_t recip(const uint4& cp) const { return recip_tab[(cp.x >> 21) & 31u]; }
};

// per-lane machine states: the component being decoded (3..8 = BVGraphComponent, mod.rs:46-61) or one of
enum : uint32_t { C_FETCH = 9, C_IDLE = 10 };

// node + nat2int(x) in 32-bit arithmetic (ids are < 2^32, so a valid x is < 2^33); false on leaving [0, 2^32-2]
__device__ __forceinline__ bool add_nat(uint32_t v, uint64_t x, uint32_t& out) {
  const uint32_t half = (uint32_t)(x >> 1);
  const bool neg = (x & 1) != 0;
  out = neg ? v - half - 1u : v + half;
  return (x >> 33) == 0 && (neg ? half < v : (out >= v && out != 0xFFFFFFFFu));
}

// LIST: node t is rv.nodes[t] (random access) instead of rv.lo + t.  ALLHOT: every table entry is in shared memory.
template <bool LIST, bool ALLHOT>
__global__ void __launch_bounds__(K1_THREADS, 1) k_entropy(DevGraph g, RangeView rv, K1Tables kt, uint32_t refill_min) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  uint4* s_cp = reinterpret_cast<uint4*>(k1_smem);                        // 9 x 16 B
  uint32_t* s_goff = reinterpret_cast<uint32_t*>(k1_smem + 9 * 16);       // 9 (+ pad to 12)
  uint32_t* s_recip = s_goff + 12;                                         // 32
  uint2* s_bkt = reinterpret_cast<uint2*>(s_recip + 32);
  uint2* s_ent = s_bkt + kt.bkt_words;
  {
    if (threadIdx.x < WGA_COMPONENTS) {
      uint4 c = kt.cp[threadIdx.x];
      c.x |= threadIdx.x << 26;  // component index for the cold-entry path
      s_cp[threadIdx.x] = c;
      s_goff[threadIdx.x] = kt.gent_off[threadIdx.x];
    }
    if (threadIdx.x < 32) s_recip[threadIdx.x] = 65536u / (threadIdx.x ? threadIdx.x : 1u) + 1u;
    // buckets and hot entries of components 3..8, laid out as kt.cp says
    for (int c = Blocks; c <= Residual; ++c) {
      const uint4 cp = kt.cp[c];
      const uint32_t nb = g.tb.nb[c];
      for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_bkt[cp.y + i] = g.tb.bkt[g.tb.bkt_off[c] + i];
      for (uint32_t i = threadIdx.x; i < cp.w; i += blockDim.x) s_ent[cp.z + i] = g.tb.ent[g.tb.ent_off[c] + i];
    }
  }
  __syncthreads();
  const SmemTables<ALLHOT> tab{s_bkt, s_ent, g.tb.ent, s_goff, s_recip};
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t stream_id = blockIdx.x * K1_WARPS + (threadIdx.x >> 5);
  const uint32_t minint = g.min_interval;
  const uint32_t c_extras = minint ? (uint32_t)IntervalCount : (uint32_t)FirstResidual;
  const uint32_t lo32 = (uint32_t)rv.lo;
  const uint16_t* __restrict__ const stream = g.stream;

  // warp-uniform
  uint32_t nx = 0, ne = 0;      // nodes [nx, ne) of the current unit are not yet handed out
  bool exhausted = false;
  uint32_t row = 0;             // rows written so far by this warp
  uint32_t* rowp = nullptr;     // row `row` of the stream, this lane's word
  bool rows_ok = true;

  // per-lane record state
  uint32_t c = C_FETCH, t = 0, v = 0, prev = 0, d = 0, dref = 0, k = 0, b = 0, copied = 0, pos = 0, extras = 0, ni = 0,
           row0 = 0, ns = 0, flags = 0;
  uint32_t* wp = nullptr;  // residuals parked in the node's own slot (MF_INSLOT)
  Dec dc{0, 0, 0};

  for (;;) {
    uint32_t err = 0;
    bool finish = false;
    // ---------------------------------------------------------------- hand out nodes
    const uint32_t nf = __ballot_sync(FULL, c == C_FETCH);
    if (nf) {
      const uint32_t busy = __ballot_sync(FULL, c < C_FETCH);
      const uint32_t cnt = (uint32_t)__popc(nf);
      if (busy == 0 || cnt >= refill_min) {
        const uint32_t r = (uint32_t)__popc(nf & lt_mask);  // rank among the fetching lanes
        uint32_t my = NOT_FOUND;
        uint32_t taken = 0;
        while (taken < cnt) {  // (uniform) the fetchers may straddle a unit boundary
          if (nx >= ne) {
            if (exhausted) break;
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(rv.unit_ctr, 1u);
            u = __shfl_sync(FULL, u, 0);
            if (u >= rv.n_units) { exhausted = true; break; }
            nx = u * rv.unit;
            ne = min(nx + rv.unit, rv.n);
            if (lane == 0) rv.unit_stream[u] = stream_id;
          }
          const uint32_t take = min(ne - nx, cnt - taken);
          if (r >= taken && r < taken + take) my = nx + (r - taken);
          nx += take;
          taken += take;
        }
        if (c == C_FETCH) {
          if (my == NOT_FOUND) { if (exhausted) c = C_IDLE; }
          else {
            t = my;
            const uint4 p = rv.nrec[t];
            d = p.z;
            v = LIST ? rv.nodes[t] : lo32 + t;
            dc.state = p.x;
            dc.sp = p.y;
            row0 = row;
            ns = 0;
            flags = 0;
            ni = 0;
            wp = nullptr;
            if (d == 0) finish = true;
            else {
              dec_prime(dc, stream);
              const uint32_t rt = p.w & RT_MASK;
              b = p.w >> RT_BITS;
              if (rt == 0) { extras = d; c = c_extras; }
              else {
                flags = 8u;  // (internal) the node has a reference
                dref = rv.outdeg[t - rt];
                if (b > dref && b - dref > 1u) err |= ERR_CORRUPT;  // at most dref + 1 blocks
                else if (b == 0) {
                  if (dref > d) err |= ERR_CORRUPT;
                  else { extras = d - dref; if (extras) c = c_extras; else finish = true; }
                } else { k = b; pos = 0; copied = 0; c = Blocks; }
              }
            }
          }
        }
      }
      if (__all_sync(FULL, c == C_IDLE)) break;
    }
    // ---------------------------------------------------------------- one symbol per busy lane
    const bool decoding = c < C_FETCH && !err && !finish;
    uint32_t val = 0;
    if (decoding) {
      const uint64_t x = ans_decode_cp(s_cp[c], tab, dc, stream, err);
      const uint32_t xl = (uint32_t)x;
      const bool wide = (x >> 32) != 0;  // only nat2int arguments (first residual / interval start) may need 33 bits
      if (c >= FirstResidual) {
        bool ok;
        if (c == FirstResidual) {
          ok = add_nat(v, x, val);
          if (ok && extras >= DSOLO) {  // long run: park it at the tail of the node's own slot
            uint32_t* slot = node_slot(rv, t);
            if (!slot) { err |= ERR_WORKSPACE; ok = true; }
            else {
              wp = slot + (d - extras);
              flags |= MF_INSLOT;
              if (!(flags & 8u) && ni == 0) flags |= MF_FINAL;  // no reference, no interval: the slot is the final list
            }
          }
          c = Residual;
        } else {
          val = prev + 1u + xl;
          ok = !wide && val > prev && val != 0xFFFFFFFFu;
        }
        if (!ok) err |= ERR_SYMBOL_WIDTH;
        prev = val;
        if (--extras == 0) finish = true;
      } else if (c == Blocks) {
        const uint32_t len = xl + ((k != b) ? 1u : 0u);  // first block literal, later ones minus 1
        if (wide || len < xl || len > dref - pos) err |= ERR_CORRUPT;
        else {
          if (((b - k) & 1u) == 0) copied += len;  // blocks before this one: even = copy block
          pos += len;
          val = pos;  // cumulative end
          if (--k == 0) {
            if ((b & 1u) == 0) copied += dref - pos;  // even count: the tail is copied too
            if (copied > d) err |= ERR_CORRUPT;
            else {
              extras = d - copied;
              if (extras) c = c_extras; else finish = true;
            }
          }
        }
      } else if (c == IntervalCount) {
        if (wide || xl > extras || (uint64_t)xl * minint > extras) err |= ERR_CORRUPT;
        else {
          ni = xl;
          val = xl;
          k = 0;
          c = ni ? (uint32_t)IntervalStart : (uint32_t)FirstResidual;
        }
      } else if (c == IntervalStart) {
        bool ok;
        if (k == 0) ok = add_nat(v, x, val);
        else { val = prev + 1u + xl; ok = !wide && val > prev && val != 0xFFFFFFFFu; }  // prev: end of the last one
        if (!ok) err |= ERR_SYMBOL_WIDTH;
        prev = val;
        c = IntervalLen;
      } else {  // IntervalLen
        const uint32_t len = xl + minint;
        const uint32_t end = prev + len;  // one past the end of this interval
        if (wide || len < xl || len > extras || len == 0) err |= ERR_CORRUPT;
        else if (end < prev) err |= ERR_SYMBOL_WIDTH;
        else {
          val = len;
          prev = end;
          extras -= len;
          if (++k == ni) { if (extras) c = FirstResidual; else finish = true; }
          else c = IntervalStart;
        }
      }
    }
    // ---------------------------------------------------------------- one coalesced row
    const bool have = decoding && !err;
    const bool to_row = have && wp == nullptr;
    if (have && wp != nullptr) *wp++ = val;
    if (__any_sync(FULL, to_row)) {
      if ((row & (CH - 1)) == 0) {  // new chunk
        uint32_t cid = 0;
        if (lane == 0) {
          cid = atomicAdd(rv.chunk_ctr, 1u);
          if (cid < rv.rows_cap && (row >> CH_SHIFT) < MAXC) rv.stream_chunks[stream_id * MAXC + (row >> CH_SHIFT)] = cid;
          else { cid = 0xFFFFFFFFu; atomicOr(rv.err, ERR_WORKSPACE); }
        }
        cid = __shfl_sync(FULL, cid, 0);
        rows_ok = cid != 0xFFFFFFFFu;
        rowp = rv.rows + (size_t)(rows_ok ? cid : 0u) * CH * 32 + lane;
      }
      if (to_row) {
        if (rows_ok) *rowp = val;
        ++ns;
      }
      rowp += 32;
      ++row;
    }
    // ---------------------------------------------------------------- end of a record
    if (err) {
      atomicOr(rv.err, err);
      rv.meta[t] = make_uint2(0u, (lane << 24) | (MF_ERR << 29));
      c = C_FETCH;
    } else if (finish) {
      if (ns > NSYM_MAX) { atomicOr(rv.err, ERR_LIMIT); flags |= MF_ERR; }
      rv.meta[t] = make_uint2(row0, (ns & NSYM_MAX) | (lane << 24) | ((flags & 7u) << 29));
      c = C_FETCH;
    }
  }
}

// -------------------------------------------------------------------------------------------- K2: tile kernel
constexpr int K2_NT_MAX = 256;
constexpr uint32_t HMAX = 64;    // look-back window of a tile (nodes before it that it may have to re-resolve)
constexpr uint32_t LMAXT = 8;    // deepest reference chain a tile resolves
constexpr uint32_t NCLS = 8, NBINS = (LMAXT + 1) * NCLS;
constexpr uint32_t HRECCAP = 1536;  // words of records staged compactly in shared memory (look-back nodes, long records)

__device__ __forceinline__ uint32_t d_class(uint32_t d) {  // 0 = largest
  return d > 96 ? 0u : d > 64 ? 1u : d > 48 ? 2u : d > 32 ? 3u : d > 16 ? 4u : d > 8 ? 5u : d > 4 ? 6u : 7u;
}

struct TileCfg {
  uint32_t tile, slotcap, rowcap, dbig, nt, seg, taskcap, dbg;
};

// word k of a record in shared memory (the tile's rows: stride 32; staged look-back records: stride 1)
struct SmemRec {
  const uint32_t* p;
  uint32_t stride;
  __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return p[k * stride]; }
};
// word k of a record in global memory (chunks are not contiguous)
struct GlobalRec {
  const RangeView* rv;
  uint32_t s, r0, ln;
  __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return *row_word(*rv, s, r0 + k, ln); }
};

// One successor list: merge of (copied elements of the finished referenced list, expanded intervals, residuals).
//   rec(k) = word k of the node's K1 record: [cumulative copy-block ends x b][interval count][start,len x ni][residuals]
// inslot: the residuals are not in the record but parked at the tail of `out` (MF_INSLOT); they are consumed before
// the write position reaches them (written <= copied + interval elements + residuals consumed).
// The next element of the copy run and of the residual run is loaded one step ahead, so that the per-element
// dependency chain is min / compare / select only.  PADDED: `ref` has one readable word behind its last element.
template <bool PADDED, class Rec>
__device__ __forceinline__ void merge_list(uint32_t* __restrict__ out, uint32_t d, const uint32_t* __restrict__ ref,
                                           uint32_t dref, uint32_t b, uint32_t ns, uint32_t minint, const Rec rec,
                                           bool inslot) {
  // ---- layout of the record
  uint32_t idx = b, ni = 0;
  if (ns > b && minint) ni = rec(idx++);
  uint32_t ip = idx, iend = idx + 2 * ni;  // interval pairs
  uint32_t rp = iend, rend = ns;           // residuals
  if (ns < b || 2 * (uint64_t)ni > (uint64_t)(ns - idx)) { ni = 0; ip = iend = rp = rend = 0; b = 0; }  // inconsistent record
  const uint32_t* tail = nullptr;
  if (inslot) {  // parked residuals: d - copied - interval elements of them
    uint64_t used = 0;
    for (uint32_t q = 0; q < ni; ++q) used += rec(ip + 2 * q + 1);
    if (ref) {
      if (b == 0) used += dref;
      else {
        uint32_t prev = 0;
        for (uint32_t q = 0; q < b; ++q) { const uint32_t e = rec(q); if ((q & 1u) == 0) used += e - prev; prev = e; }
        if ((b & 1u) == 0) used += dref - prev;
      }
    }
    if (used > d) return;  // K1 checked this
    rend = d;
    rp = (uint32_t)used;  // index into out
    tail = out;
  }
  // ---- heads of the three runs
  uint32_t ci = 0, cend = 0, kb = 0, cval = INF, cnext = INF;
  auto next_copy_block = [&]() {  // after copy block kb (even): skip block kb+1, copy block kb+2
    for (;;) {
      kb += 2;
      if (kb - 1 < b) { ci = rec(kb - 1); cend = kb < b ? rec(kb) : dref; }
      else { ci = cend = dref; cval = INF; return; }
      if (cend > dref) cend = dref;
      if (ci < cend) { cval = ref[ci]; cnext = (PADDED || ci + 1 < dref) ? ref[ci + 1] : INF; return; }
    }
  };
  if (ref) {
    cend = b ? min(rec(0), dref) : dref;
    if (ci < cend) { cval = ref[0]; cnext = (PADDED || 1 < dref) ? ref[1] : INF; } else next_copy_block();
  }
  uint32_t ival = INF, ilim = 0;
  if (ni) { ival = rec(ip); ilim = ival + rec(ip + 1); ip += 2; }
  uint32_t rval = INF, rnext = INF;
  if (rp < rend) rval = tail ? tail[rp] : rec(rp);
  if (!tail && rp + 1 < rend) rnext = rec(rp + 1);
  // ---- merge
  for (uint32_t p = 0; p < d; ++p) {
    const uint32_t mn = min(cval, min(ival, rval));
    if (mn == cval) {
      if (++ci < cend) { cval = cnext; cnext = (PADDED || ci + 1 < dref) ? ref[ci + 1] : INF; } else next_copy_block();
    } else if (mn == rval) {
      ++rp;
      if (tail) rval = rp < rend ? tail[rp] : INF;
      else { rval = rnext; rnext = rp + 1 < rend ? rec(rp + 1) : INF; }
    } else {
      if (++ival == ilim) {
        if (ip < iend) { ival = rec(ip); ilim = ival + rec(ip + 1); ip += 2; } else ival = INF;
      }
    }
    out[p] = mn;  // after the heads moved on: out[p] may be the parked residual that was just read
  }
}

// ---- element-parallel resolve of one level of a tile, everything in shared memory ----------------------------
// A successor list is the sorted union of three sorted runs with distinct values: the kept elements of the
// referenced list (copy-block mask), the interval elements, the residuals.  So the final position of an element is
// its rank in its own run plus the number of elements of the other two runs below it -- no serial merge, no
// dependency between elements.  Work items of a level: chunks of GCH consecutive positions of a referenced list
// (block state and ranks advance incrementally inside a chunk), single residuals, single intervals (their elements
// are consecutive in the output).  All lanes of all warps take items from one flat index space per level.
constexpr uint32_t GCH = 4;

// the K1 record of a node as the tile sees it: word k at s[base + k * stride]
struct TileRec {
  const uint32_t* p;
  uint32_t stride;
  __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return p[k * stride]; }
};

struct NodeCtx {
  uint32_t* out;
  const uint32_t* ref;  // nullptr: no (usable) reference
  uint32_t d, dref, b, ni, ipb, rb0, nres;
  TileRec rec;
  // copy blocks: cumulative ends E(q); block q is a copy block when q is even; after the b listed blocks the tail
  __device__ __forceinline__ uint32_t E(uint32_t q) const { return min(rec(q), dref); }
  // block that holds position j, and the kept elements before j
  __device__ __forceinline__ void block_at(uint32_t j, uint32_t& q, uint32_t& bend, uint32_t& kept) const {
    uint32_t prev = 0;
    q = 0;
    kept = 0;
    while (q < b) {
      const uint32_t e = E(q);
      if (e > j) break;
      if (!(q & 1u)) kept += e - prev;
      prev = e;
      ++q;
    }
    bend = q < b ? E(q) : dref;
    if (!(q & 1u)) kept += j - prev;
  }
  __device__ __forceinline__ uint32_t res_below(uint32_t x) const {  // residuals < x
    uint32_t lo = 0, hi = nres;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (rec(rb0 + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  __device__ __forceinline__ uint32_t int_below(uint32_t x) const {  // interval elements < x
    uint32_t cnt = 0;
    for (uint32_t k = 0; k < ni; ++k) {
      const uint32_t st = rec(ipb + 2 * k), len = rec(ipb + 2 * k + 1);
      if (x > st) cnt += min(x - st, len);
    }
    return cnt;
  }
  __device__ __forceinline__ uint32_t kept_below(uint32_t x) const {  // kept elements of the referenced list < x
    if (!ref) return 0;
    uint32_t lo = 0, hi = dref;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (ref[mid] < x) lo = mid + 1; else hi = mid;
    }
    uint32_t q, bend, kept;
    block_at(lo, q, bend, kept);
    return kept;
  }
};

__device__ __forceinline__ void item_ref_chunk(const NodeCtx& c, uint32_t u) {
  const uint32_t j0 = u * GCH;
  uint32_t q, bend, kept;
  c.block_at(j0, q, bend, kept);
  uint32_t rcnt = c.res_below(c.ref[j0]);
#pragma unroll
  for (uint32_t i = 0; i < GCH; ++i) {
    const uint32_t j = j0 + i;
    if (j >= c.dref) break;
    const uint32_t x = c.ref[j];
    while (j >= bend && q <= c.b) { ++q; bend = q < c.b ? c.E(q) : c.dref; }
    while (rcnt < c.nres && c.rec(c.rb0 + rcnt) < x) ++rcnt;
    if (!(q & 1u)) {
      const uint32_t pos = kept + rcnt + c.int_below(x);
      if (pos < c.d) c.out[pos] = x;
      ++kept;
    }
  }
}
__device__ __forceinline__ void item_residual(const NodeCtx& c, uint32_t i) {
  const uint32_t x = c.rec(c.rb0 + i);
  const uint32_t pos = i + c.kept_below(x) + c.int_below(x);
  if (pos < c.d) c.out[pos] = x;
}
__device__ __forceinline__ void item_interval(const NodeCtx& c, uint32_t k) {
  uint32_t pos = 0;
  for (uint32_t q = 0; q < k; ++q) pos += c.rec(c.ipb + 2 * q + 1);
  const uint32_t st = c.rec(c.ipb + 2 * k), len = c.rec(c.ipb + 2 * k + 1);
  pos += c.kept_below(st) + c.res_below(st);
  for (uint32_t e = 0; e < len && pos + e < c.d; ++e) c.out[pos + e] = st + e;
}

constexpr uint32_t F21 = (1u << 21) - 1;

template <int NT>
__global__ void __launch_bounds__(NT) k_tile(RangeView rv, TileCfg cfg, uint32_t minint, uint32_t lookback) {
  extern __shared__ __align__(16) uint32_t k2_smem[];
  uint32_t* const s_slots = k2_smem;                             // slotcap + 8
  uint32_t* const s_rows = s_slots + cfg.slotcap + 8;            // (rowcap + 1) * 32
  uint32_t* const s_hrec = s_rows + (cfg.rowcap + 1) * 32;       // HRECCAP + 8: records staged compactly
  unsigned long long* const s_pre = reinterpret_cast<unsigned long long*>(s_hrec + HRECCAP + 8);  // NT + 2: item prefix (3 x 21 bits)
  uint32_t* const s_d = reinterpret_cast<uint32_t*>(s_pre + NT + 2);  // per candidate: outdegree
  uint32_t* const s_so = s_d + NT;                               //   slot offset
  uint32_t* const s_rb = s_so + NT;                              //   head word
  uint32_t* const s_row0 = s_rb + NT;                            //   record: first row
  uint32_t* const s_my = s_row0 + NT;                            //   record: words | lane << 24 | flags << 29
  uint32_t* const s_ck = s_my + NT;                              //   chunk of the record's first row | crosses << 31
  uint32_t* const s_recp = s_ck + NT;                            //   staged record: word index in k2_smem | (stride 32) << 31
  uint32_t* const s_st = s_recp + NT;                            //   bit 0 big, 1 hard, 2 part, 3 need, 4 long record ; level << 8
  uint32_t* const s_ni = s_st + NT;                              //   interval count | has-count-word << 31
  uint32_t* const s_order = s_ni + NT;                           // task nodes ordered by level
  uint32_t* const s_bin = s_order + NT;                          // LMAXT + 2 : task nodes per level
  uint32_t* const s_binbase = s_bin + 16;                        // LMAXT + 2
  __shared__ uint32_t s_scal[8];  // 1 rmin, 2 rmax, 3 hrec used, 4 maxlev, 5 holes, 6 first overflowing candidate
  typedef cub::BlockScan<uint32_t, NT> BlockScan;
  typedef cub::BlockScan<unsigned long long, NT> BlockScan64;
  __shared__ union { typename BlockScan::TempStorage a; typename BlockScan64::TempStorage b; } s_scan;
  constexpr uint32_t NW = NT / 32;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t A = blockIdx.x * cfg.tile, B = min(A + cfg.tile, rv.n);
  const uint32_t my_stream = rv.unit_stream[blockIdx.x];
  const uint32_t longrec = max(cfg.rowcap / 3, 1u);  // records longer than this are staged compactly, not with the rows
  long long tk0 = 0;
  auto tick = [&](int phase) {  // debug: cycles per phase, summed over tiles (thread 0 of every block)
    if (rv.stats && tid == 0) {
      const long long now = clock64();
      atomicAdd(&rv.stats[phase], (unsigned long long)(now - tk0));
      tk0 = now;
    }
  };
  if (rv.stats && tid == 0) { tk0 = clock64(); atomicAdd(&rv.stats[15], 1ull); }

  for (uint32_t a = A; a < B;) {
    const uint32_t C0 = a > lookback ? a - lookback : 0u;
    const uint32_t own0 = a - C0;
    uint32_t limit = B;  // owned nodes of this round: [a, limit)
    uint32_t M, total_slots;
    bool fits;
    // -------------------------------------------------------------- descriptors (one candidate per thread)
    uint32_t d = 0, rbw = 0, big = 0, lev = 0, hard = 0;
    uint2 m = make_uint2(0, 0);
    const uint32_t t = C0 + tid;
    if (t < B) {
      const uint4 nr = rv.nrec[t];
      m = rv.meta[t];
      const uint32_t us = rv.unit_stream[t / rv.unit];
      d = nr.z;
      rbw = nr.w;
      const uint32_t fl = m.y >> 29, nsym = m.y & NSYM_MAX;
      big = (d >= cfg.dbig || nsym > HRECCAP / 2 || fl != 0) ? 1u : 0u;
      if (!big && (nsym > longrec || tid < own0)) big |= 16u;  // staged compactly (look-back records always are)
      uint32_t ck = 0;
      if (nsym && !(big & 1u)) {
        ck = rv.stream_chunks[us * MAXC + (m.x >> CH_SHIFT)];
        if ((m.x & (CH - 1)) + nsym > CH) ck |= 0x80000000u;  // the record continues in another chunk
      }
      s_d[tid] = d;
      s_rb[tid] = rbw;
      s_row0[tid] = m.x;
      s_my[tid] = m.y;
      s_ck[tid] = ck;
    }
    s_st[tid] = big;
    if (tid < 8) s_scal[tid] = 0;
    __syncthreads();
    tick(0);
    // -------------------------------------------------------------- reference chains: depth, hardness, needed look-back nodes
    if (t < B) {
      hard = big & 1u;
      uint32_t j = tid;
      while (!hard) {
        const uint32_t rt = s_rb[j] & RT_MASK;
        if (rt == 0) break;
        if (rt > j) { hard = 1; break; }  // the chain leaves the look-back window
        j -= rt;
        if (s_st[j] & 1u) hard = 1;
        else if (++lev > LMAXT) hard = 1;
        else if (tid >= own0) atomicOr(&s_st[j], 8u);
      }
    }
    __syncthreads();
    const uint32_t nsym = m.y & NSYM_MAX;
    const bool longr = (big & 16u) != 0;
    uint32_t hro = INF;
    for (;;) {  // shrink the round until it fits shared memory
      M = limit - C0;
      const bool owned = tid >= own0 && tid < M;
      const bool part = tid < M && !hard && (owned || (s_st[tid] & 8u));
      uint32_t so = 0;
      BlockScan(s_scan.a).ExclusiveSum(part ? d : 0u, so, total_slots);
      if (tid == 0) { s_scal[1] = INF; s_scal[2] = 0; s_scal[3] = 0; s_scal[6] = INF; }
      __syncthreads();
      hro = INF;
      if (part && nsym) {
        if (longr) hro = atomicAdd(&s_scal[3], nsym);
        else {
          atomicMin(&s_scal[1], m.x);
          atomicMax(&s_scal[2], m.x + nsym);
        }
      }
      // first candidate whose slot would not fit
      if (part && so + d > cfg.slotcap) atomicMin(&s_scal[6], tid);
      __syncthreads();
      const uint32_t rmin = s_scal[1], rmax = s_scal[2], over = s_scal[6];
      fits = over == INF && (rmax <= rmin || rmax - rmin <= cfg.rowcap) && s_scal[3] <= HRECCAP;
      if (fits || limit == a + 1) {
        s_so[tid] = so;
        s_st[tid] = (s_st[tid] & 25u) | (hard << 1) | (part ? 4u : 0u) | (lev << 8);
        break;
      }
      // halve the owned range (or cut it at the first slot overflow, whichever is smaller)
      uint32_t nl = a + max(1u, (limit - a) / 2);
      if (over != INF && over > own0 && C0 + over < nl) nl = C0 + over;
      limit = nl;
      __syncthreads();
    }
    // (a single node that does not fit with its ancestors: !fits, the global pass takes it)
    const bool owned = tid >= own0 && tid < M;
    const bool part = fits && tid < M && !hard && (owned || (s_st[tid] & 8u));
    const uint32_t rmin = s_scal[1], rmax = s_scal[2];
    // -------------------------------------------------------------- owned nodes: hard list, flags
    if (owned) {
      const uint32_t fl = m.y >> 29;
      uint32_t hf = 0;
      if (!part) hf = (d == 0 || (fl & (MF_ERR | MF_FINAL))) ? 2u : 1u;
      rv.hardflag[t] = (uint8_t)hf;
      if (hf == 1u) {
        const uint32_t pos = atomicAdd(rv.hard_count, 1u);
        rv.hard_list[pos] = t;
      }
      if (!part && d != 0) s_scal[5] = 1;  // hole in the owned slots: no bulk copy-out
    }
    if (!fits) { __syncthreads(); a = limit; continue; }
    tick(1);
    // -------------------------------------------------------------- stage in: the tile's rows, compact records
    if (rmax > rmin) {  // 16-byte pieces, independent of one another: all the loads are in flight together
      const uint32_t np = (rmax - rmin) * 8;
#pragma unroll 4
      for (uint32_t q = tid; q < np; q += NT) {
        const uint32_t r = rmin + (q >> 3);
        const uint32_t cid = rv.stream_chunks[my_stream * MAXC + (r >> CH_SHIFT)];
        const uint4 v = *reinterpret_cast<const uint4*>(rv.rows + ((size_t)cid * CH + (r & (CH - 1))) * 32 + (q & 7u) * 4);
        *reinterpret_cast<uint4*>(s_rows + (r - rmin) * 32 + (q & 7u) * 4) = v;
      }
    }
    s_recp[tid] = 0;
    if (part && nsym) {
      if (longr) {  // every thread copies its own compact record (a column of the rows: one word per row)
        s_recp[tid] = (uint32_t)(s_hrec - k2_smem) + hro;
        const uint32_t ck = s_ck[tid], ln = (m.y >> 24) & 31u;
        if (!(ck >> 31)) {
          const uint32_t* src = rv.rows + ((size_t)ck * CH + (m.x & (CH - 1))) * 32 + ln;
#pragma unroll 4
          for (uint32_t q = 0; q < nsym; ++q) s_hrec[hro + q] = src[(size_t)q * 32];
        } else {
          const uint32_t s = rv.unit_stream[t / rv.unit];
          for (uint32_t q = 0; q < nsym; ++q) s_hrec[hro + q] = *row_word(rv, s, m.x + q, ln);
        }
      } else s_recp[tid] = ((uint32_t)(s_rows - k2_smem) + (m.x - rmin) * 32 + ((m.y >> 24) & 31u)) | 0x80000000u;
    }
    if (part) atomicMax(&s_scal[4], lev);
    if (tid < 16) s_bin[tid] = 0;
    __syncthreads();
    tick(2);
    // -------------------------------------------------------------- work items per node; nodes ordered by level
    const bool task = part && d != 0;
    const uint32_t rt = rbw & RT_MASK, bcnt = rbw >> RT_BITS;
    uint32_t rank = 0;
    if (task) {
      const TileRec myrec{k2_smem + (s_recp[tid] & 0x7FFFFFFFu), (s_recp[tid] >> 31) ? 32u : 1u};
      const bool has_cnt = nsym > bcnt && minint != 0;
      uint32_t ni = has_cnt ? myrec(bcnt) : 0u, nres = 0;
      const uint32_t hdr = bcnt + (has_cnt ? 1u : 0u);
      if (nsym < hdr || 2ull * ni > (uint64_t)(nsym - hdr)) ni = 0;  // inconsistent record
      else nres = nsym - hdr - 2 * ni;
      s_ni[tid] = ni | (has_cnt ? 0x80000000u : 0u);
      (void)nres;
      rank = atomicAdd(&s_bin[lev], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      for (uint32_t l = 0; l <= LMAXT + 1; ++l) { s_binbase[l] = acc; acc += s_bin[l]; }
    }
    __syncthreads();
    if (task) s_order[s_binbase[lev] + rank] = tid;
    // prefix of the item counts in level order: thread q holds the node at position q
    __syncthreads();
    {
      const uint32_t ntask = s_binbase[LMAXT + 1];
      unsigned long long mine = 0, pre = 0, total = 0;
      if (tid < ntask) {
        // recompute the node's counts from its parsed record (cheaper than passing them through shared memory)
        const uint32_t nn = s_order[tid];
        const uint32_t rbn = s_rb[nn], nsn = s_my[nn] & NSYM_MAX, nin = s_ni[nn];
        const uint32_t rtn = rbn & RT_MASK, bn = rbn >> RT_BITS;
        const uint32_t hdr = bn + (nin >> 31);
        const uint32_t nin_ = nin & 0x7FFFFFFFu;
        const uint32_t nresn = nsn >= hdr + 2 * nin_ ? nsn - hdr - 2 * nin_ : 0u;
        const uint32_t drefn = rtn ? s_d[nn - rtn] : 0u;
        mine = (unsigned long long)((drefn + GCH - 1) / GCH) | ((unsigned long long)nresn << 21) | ((unsigned long long)nin_ << 42);
      }
      BlockScan64(s_scan.b).ExclusiveSum(mine, pre, total);
      if (tid < ntask) s_pre[tid] = pre;
      if (tid == 0) { s_pre[ntask] = total; }
    }
    __syncthreads();
    tick(3);
    // -------------------------------------------------------------- resolve, one level after the other
    const uint32_t maxlev = s_scal[4];
    for (uint32_t l = 0; l <= maxlev; ++l) {
      const uint32_t lb = s_binbase[l], le = s_binbase[l + 1];
      if (le > lb) {
        const unsigned long long P0 = s_pre[lb], P1 = s_pre[le];
        const uint32_t NA = (uint32_t)(P1 & F21) - (uint32_t)(P0 & F21);
        const uint32_t NB = (uint32_t)((P1 >> 21) & F21) - (uint32_t)((P0 >> 21) & F21);
        const uint32_t NC = (uint32_t)((P1 >> 42) & F21) - (uint32_t)((P0 >> 42) & F21);
        for (uint32_t f = tid; f < NA + NB + NC; f += NT) {
          const uint32_t kind = f < NA ? 0u : f < NA + NB ? 1u : 2u;
          const uint32_t sh = kind * 21;
          const uint32_t key = (uint32_t)((P0 >> sh) & F21) + (kind == 0 ? f : kind == 1 ? f - NA : f - NA - NB);
          uint32_t lo = lb, hi = le;  // largest position whose prefix is <= key
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t)((s_pre[mid] >> sh) & F21) <= key) lo = mid; else hi = mid;
          }
          const uint32_t u = key - (uint32_t)((s_pre[lo] >> sh) & F21);
          const uint32_t n = s_order[lo];
          const uint32_t rbn = s_rb[n], nin = s_ni[n], rp = s_recp[n];
          const uint32_t rtn = rbn & RT_MASK;
          NodeCtx c;
          c.out = s_slots + s_so[n];
          c.d = s_d[n];
          c.dref = rtn ? s_d[n - rtn] : 0u;
          c.ref = c.dref ? s_slots + s_so[n - rtn] : nullptr;
          c.b = rbn >> RT_BITS;
          c.ni = nin & 0x7FFFFFFFu;
          c.ipb = c.b + (nin >> 31);
          c.rb0 = c.ipb + 2 * c.ni;
          const uint32_t nsn = s_my[n] & NSYM_MAX;
          c.nres = nsn >= c.rb0 ? nsn - c.rb0 : 0u;
          c.rec = TileRec{k2_smem + (rp & 0x7FFFFFFFu), (rp >> 31) ? 32u : 1u};
          if (kind == 0) item_ref_chunk(c, u);
          else if (kind == 1) item_residual(c, u);
          else item_interval(c, u);
        }
      }
      __syncthreads();
      tick(4 + min(l, 3u));
    }
    // -------------------------------------------------------------- copy out the owned lists
    {
      const uint32_t s_beg = s_so[own0];
      const uint32_t s_end = total_slots;  // owned nodes are the last candidates
      const bool straddle = a < rv.h && limit > rv.h;
      if (!s_scal[5] && !straddle && s_end > s_beg) {
        const uint64_t o0 = rv.offs[a], o1 = rv.offs[limit];
        uint32_t* dst = nullptr;
        if (a < rv.h) { if (o1 <= rv.halo_cap) dst = rv.halo_succ + o0; }
        else { const uint64_t bs = rv.offs[rv.h]; if (o1 - bs <= rv.succ_cap) dst = rv.succ + (o0 - bs); }
        if (!dst) { if (tid == 0) atomicOr(rv.err, ERR_WORKSPACE); }
        else for (uint32_t e = s_beg + tid; e < s_end; e += NT) dst[e - s_beg] = s_slots[e];
      } else {
        for (uint32_t i = own0 + warp; i < M; i += NW) {
          if (!(s_st[i] & 4u) || s_d[i] == 0) continue;
          uint32_t* dst = node_slot(rv, C0 + i);
          if (!dst) { if (lane == 0) atomicOr(rv.err, ERR_WORKSPACE); continue; }
          const uint32_t* srcp = s_slots + s_so[i];
          for (uint32_t e = lane; e < s_d[i]; e += 32) dst[e] = srcp[e];
        }
      }
    }
    __syncthreads();
    tick(8);
    a = limit;
  }
}

// -------------------------------------------------------------------------------------------- K2: global pass
// Nodes the tiles left out (hardflag 1).  Depth among themselves: ancestors that are already final count 0.
__global__ void __launch_bounds__(256) k_hard_levels(RangeView rv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nh = *rv.hard_count;
  uint32_t lev = 0;
  if (i < nh) {
    uint32_t j = rv.hard_list[i];
    for (;;) {
      const uint32_t rt = rv.nrec[j].w & RT_MASK;
      if (rt == 0) break;
      j -= rt;
      if (rv.hardflag[j] != 1) break;
      ++lev;
    }
    rv.hard_lev[i] = lev;
  }
  for (int o = 16; o; o >>= 1) lev = max(lev, __shfl_xor_sync(FULL, lev, o));
  if ((threadIdx.x & 31) == 0 && lev) atomicMax(rv.maxlevel, lev);
}

// One lane per hard node of the given depth: the same merge, from global memory.
__global__ void __launch_bounds__(128) k_hard_resolve(RangeView rv, uint32_t level, uint32_t minint) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *rv.hard_count || rv.hard_lev[i] != level) return;
  const uint32_t t = rv.hard_list[i];
  const uint32_t d = rv.outdeg[t], rbw = rv.nrec[t].w;
  const uint2 m = rv.meta[t];
  const uint32_t rt = rbw & RT_MASK, b = rbw >> RT_BITS, nsym = m.y & NSYM_MAX, fl = m.y >> 29;
  uint32_t* out = node_slot(rv, t);
  if (!out) { atomicOr(rv.err, ERR_WORKSPACE); return; }
  const uint32_t* ref = nullptr;
  uint32_t dref = 0;
  if (rt) {
    ref = node_slot(rv, t - rt);
    dref = rv.outdeg[t - rt];
    if (!ref) { atomicOr(rv.err, ERR_WORKSPACE); return; }
  }
  merge_list<false>(out, d, ref, dref, b, nsym, minint, GlobalRec{&rv, rv.unit_stream[t / rv.unit], m.x, (m.y >> 24) & 31u},
             (fl & MF_INSLOT) != 0);
}

// -------------------------------------------------------------------------------------------- random access
// graph.successors(v) for a batch of query nodes (examples/bench_random_access.rs:30-38).  The reference
// builds one decoder per query (bvgraph_decoder_factory.rs:46-58) and webgraph recurses into the referenced
// node.  Here: the closure of the queries under "referenced node" is built on the device (one round per
// chain level), sorted and de-duplicated, decoded with the same K0/K1/K2 pipeline as a node LIST, and the
// query lists are gathered into the caller's CSR.
__global__ void __launch_bounds__(256) k_query_ids(const uint64_t* q, uint64_t nq, uint64_t res_first, uint64_t res_last,
                                                   uint32_t* out, uint32_t* err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const uint64_t v = q[i];
  if (v < res_first || v >= res_last) { atomicOr(err, ERR_RANGE); out[i] = (uint32_t)res_first; return; }
  out[i] = (uint32_t)v;
}

// frontier -> referenced nodes of the frontier (outdegree + reference offset of every node: two symbols)
__global__ void __launch_bounds__(TPB) k_closure_step(DevGraph g, const uint32_t* in, uint32_t n_in, uint32_t* out,
                                                      uint32_t* count, uint32_t cap, uint64_t res_first, uint32_t* err_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const GlobalTables tab{g.tb.bkt, g.tb.ent};
  uint32_t target = NOT_FOUND, err = 0;
  if (i < n_in && g.window != 0) {
    const uint64_t v = in[i];
    Dec dc;
    load_phase(g, v, dc, err);
    const uint64_t d = ans_decode(g.tb, tab, Outdegree, dc, g.stream, err);
    if (d != 0 && !err) {
      const uint64_t r = ans_decode(g.tb, tab, ReferenceOffset, dc, g.stream, err);
      if (r > g.window) err |= ERR_CORRUPT;
      else if (r != 0 && !err) {
        if (r > v || v - r < res_first) err |= ERR_RANGE;
        else target = (uint32_t)(v - r);
      }
    }
  }
  // warp-aggregated append
  const uint32_t mask = __ballot_sync(FULL, target != NOT_FOUND);
  if (mask) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(FULL, base, __ffs(mask) - 1);
    if (target != NOT_FOUND) {
      const uint32_t pos = base + __popc(mask & ((1u << lane) - 1u));
      if (pos < cap) out[pos] = target; else err |= ERR_WORKSPACE;
    }
  }
  if (err) atomicOr(err_out, err);
}

// query i -> index in the sorted node list U; its outdegree goes to offsets[i] (scanned afterwards)
__global__ void __launch_bounds__(256) k_query_lookup(const uint32_t* qid, uint64_t nq, const uint32_t* U, uint32_t nU,
                                                      const uint64_t* offsU, uint32_t* qidx, uint64_t* offsets) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nq) return;
  if (i == nq) { offsets[nq] = 0; return; }
  const uint32_t v = qid[i];
  uint32_t lo = 0, hi = nU;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (U[mid] < v) lo = mid + 1; else hi = mid;
  }
  qidx[i] = lo;  // present by construction
  offsets[i] = offsU[lo + 1] - offsU[lo];
}

// one warp per query: coalesced copy of its list
__global__ void __launch_bounds__(256) k_query_gather(const uint32_t* qidx, uint64_t nq, const uint64_t* offsU,
                                                      const uint32_t* succU, const uint64_t* offsets, uint32_t* succ,
                                                      uint64_t succ_cap) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
    const uint64_t src = offsU[qidx[i]], dst = offsets[i], d = offsets[i + 1] - dst;
    if (dst + d > succ_cap) continue;  // reported by the host from the total
    for (uint64_t k = lane; k < d; k += 32) succ[dst + k] = succU[src + k];
  }
}

// -------------------------------------------------------------------------------------------- debug kernels
__global__ void k_expand_table(DevTables tb, int c, uint32_t n_slots, uint4* out) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const uint2 bk = tb.bkt[tb.bkt_off[c] + (slot >> 5)];
  const uint32_t j = bk.y + (uint32_t)__popc(bk.x & ((2u << (slot & 31u)) - 1u));
  const uint2 e = tb.ent[tb.ent_off[c] + j];
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

__global__ void k_decode_symbols(DevGraph g, const uint8_t* comps, uint64_t n, uint64_t ptr, uint32_t state,
                                 uint64_t* out, uint64_t* end) {
  if (threadIdx.x || blockIdx.x) return;
  const GlobalTables tab{g.tb.bkt, g.tb.ent};
  uint32_t err = 0;
  Dec dc{state, (uint32_t)ptr, 0};
  dec_prime(dc, g.stream);
  for (uint64_t i = 0; i < n; ++i) out[i] = ans_decode(g.tb, tab, comps[i], dc, g.stream, err);
  end[0] = (uint64_t)dc.sp;
  end[1] = dc.state;
  end[2] = err;
}

__global__ void k_offsets_rebase(const uint64_t* src, uint64_t base, uint64_t* dst, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - base;
}

// scalars the host needs after a decode -> mapped host memory (no DMA copy: see wga_graph::h_pub)
__global__ void k_publish(const uint64_t* tot0, const uint64_t* tot1, const uint32_t* maxlevel, const uint32_t* hard_count,
                          const uint32_t* err, uint64_t* pub) {
  if (threadIdx.x == 0) {
    pub[1] = *tot0;
    pub[2] = *tot1;
    pub[3] = *maxlevel;
    pub[4] = *err;
    pub[5] = *hard_count;
  }
}

__global__ void k_offsets_add(uint64_t* off, uint64_t n, uint64_t base) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) off[i] += base;
}

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Scalars at the head of the workspace (one 256-byte line, cleared per call).
struct Scalars {
  uint64_t lo;                // k_halo result
  uint32_t unit_ctr;
  uint32_t chunk_ctr;
  uint32_t hard_count;
  uint32_t maxlevel;
};
static_assert(NBINS + 1 <= 96, "bin prefix handles three bins per lane");
static_assert(sizeof(Scalars) <= 256, "Scalars must fit the cleared line");

struct WorkspacePlan {
  uint64_t off_outdeg, off_nrec, off_offs, off_meta, off_hflag, off_hlist, off_hlev, off_ustream, off_schunks, off_cub,
      off_halo, off_rows;
  uint64_t cub_bytes, halo_cap, fixed_bytes;
};

WorkspacePlan plan_workspace(uint64_t n, uint32_t unit) {
  WorkspacePlan p{};
  uint64_t o = 0;
  o += 256;  // Scalars
  p.off_outdeg = o; o = align_up(o + 4 * (n + 1), 256);
  p.off_nrec = o; o = align_up(o + 16 * n, 256);
  p.off_offs = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_meta = o; o = align_up(o + 8 * n, 256);
  p.off_hflag = o; o = align_up(o + n, 256);
  p.off_hlist = o; o = align_up(o + 4 * n, 256);
  p.off_hlev = o; o = align_up(o + 4 * n, 256);
  p.off_ustream = o; o = align_up(o + 4 * ((n + unit - 1) / unit + 1), 256);
  p.off_schunks = o; o = align_up(o + 4ull * MAX_STREAMS * MAXC, 256);
  size_t scan_bytes = 0;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, (uint64_t*)nullptr, (int64_t)(n + 1));
  p.cub_bytes = scan_bytes;
  p.off_cub = o; o = align_up(o + p.cub_bytes, 256);
  p.halo_cap = 1u << 20;  // successors of halo nodes (u32 each)
  p.off_halo = o; o = align_up(o + 4 * p.halo_cap, 256);
  p.off_rows = align_up(o, 16384);
  p.fixed_bytes = p.off_rows;
  return p;
}

// threads of a tile block: the look-back window plus the tile's nodes, one candidate per thread
uint32_t tile_threads(const Tuning& tn, uint32_t window) {
  const uint32_t lookback = std::min<uint32_t>(HMAX, window * LMAXT);
  const uint32_t want = (tn.tile ? tn.tile : 1) + lookback;
  return want <= 128 ? 128u : want <= 160 ? 160u : want <= 192 ? 192u : 256u;
}
uint32_t effective_tile(const Tuning& tn, uint32_t window) {
  const uint32_t lookback = std::min<uint32_t>(HMAX, window * LMAXT);
  uint32_t tile = tn.tile ? tn.tile : 1;
  if (tile > K2_NT_MAX - lookback) tile = K2_NT_MAX - lookback;
  return tile;
}

}  // namespace

uint64_t decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last) {
  uint64_t n = last - first + 4096;  // room for a halo
  WorkspacePlan p = plan_workspace(n, effective_tile(g_tuning, (uint32_t)g->prelude.compression_window));
  double frac = g->prelude.number_of_nodes ? (double)(last - first) / (double)g->prelude.number_of_nodes : 1.0;
  uint64_t arcs_est = (uint64_t)((double)g->prelude.number_of_arcs * frac) + (1u << 20);
  // rows: one 4-byte word per decoded symbol, 32 lanes per row of which about three quarters are busy; a record has
  // at most (blocks + 1 + outdegree) words
  uint64_t rows_bytes = 8 * arcs_est + 48 * n + (32ull << 20);
  // every K1 warp that gets work owns at least one chunk
  const uint32_t tile = effective_tile(g_tuning, (uint32_t)g->prelude.compression_window);
  rows_bytes += std::min<uint64_t>((n + tile - 1) / tile, MAX_STREAMS) * (4ull * CH * 32);
  return p.fixed_bytes + rows_bytes;
}

static void check_device_error(wga_graph* g, uint32_t herr, cudaStream_t st) {
  if (herr) {
    WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    if (herr & ERR_WORKSPACE) throw Error(WGA_E_WORKSPACE, "decode: workspace or output buffer too small; pass larger buffers");
    if (herr & ERR_RANGE) throw Error(WGA_E_CORRUPT, "decode: a reference leaves the decoded range");
    if (herr & ERR_SYMBOL_WIDTH) throw Error(WGA_E_UNSUPPORTED, "decode: a decoded value does not fit 32 bits");
    if (herr & ERR_LIMIT) throw Error(WGA_E_UNSUPPORTED, "decode: a record exceeds an implementation limit (2^20 copy blocks / 2^24 words)");
    throw Error(WGA_E_CORRUPT, "decode: inconsistent stream or tables");
  }
}

static uint32_t read_device_error(wga_graph* g, cudaStream_t st) {
  uint32_t herr = 0;
  WGA_CUDA(cudaMemcpyAsync(&herr, g->d_err, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  return herr;
}

void outdegrees(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, void* ws, uint64_t ws_bytes,
                cudaStream_t st) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (first > last || last > g->res_last || first < g->res_first) throw Error(WGA_E_ARG, "range outside the resident nodes");
  uint64_t n = last - first;
  WorkspacePlan p = plan_workspace(n, 192);
  if (ws_bytes < p.off_nrec) throw Error(WGA_E_WORKSPACE, "workspace too small");
  uint8_t* w = (uint8_t*)ws;
  uint32_t* outdeg = (uint32_t*)(w + p.off_outdeg);
  k_heads<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, first, nullptr, (uint32_t)n, outdeg, nullptr, g->d_err);
  count_launch();
  // the scan's temporary storage lives behind the outdegrees (the other arrays of the plan are not used here)
  size_t cb = p.cub_bytes;
  if (ws_bytes < p.off_nrec + cb) throw Error(WGA_E_WORKSPACE, "workspace too small");
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(outdeg, U32ToU64());
  WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_nrec, cb, it, d_offsets, (int64_t)(n + 1), st));
  count_launch(2);
  WGA_CUDA(cudaGetLastError());
}

static void mark(wga_graph* g, cudaStream_t st) {
  if (!g->profiling || g->n_ev >= 8) return;
  if (!g->ev[g->n_ev]) cudaEventCreate(&g->ev[g->n_ev]);
  cudaEventRecord(g->ev[g->n_ev++], st);
}

namespace {

struct DeviceInfo {
  int sms = 148;
  int smem_optin = 227 * 1024;
};
const DeviceInfo& device_info(int dev) {
  static DeviceInfo info[64];
  static bool have[64] = {};
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!have[dev]) {
    cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    have[dev] = true;
  }
  return info[dev];
}

// Shared-memory layout of the entropy kernel's tables.
uint64_t plan_k1_tables(const wga_graph* g, K1Tables& kt, int smem_limit) {
  const PackedTablesData& p = g->packed;
  const uint64_t head = 9 * 16 + 12 * 4 + 32 * 4;
  uint32_t bo = 0;
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    const uint32_t L = p.L[c], R = p.R[c] ? p.R[c] : 1u;
    kt.cp[c] = make_uint4(((1u << L) - 1u) | (L << 16) | (R << 21), 0u, 0u, 0u);
    kt.gent_off[c] = p.ent_off[c];
    if (c >= Blocks) { kt.cp[c].y = bo; bo += p.nb[c]; }
  }
  kt.bkt_words = bo;
  // entries: water-filling of what is left
  int64_t budget = ((int64_t)smem_limit - 1024 - (int64_t)head - 8ll * bo) / 8;
  if (budget < 0) throw Error(WGA_E_UNSUPPORTED, "decoder tables do not fit shared memory");
  uint32_t hot[WGA_COMPONENTS] = {};
  uint32_t left = (uint32_t)budget;
  for (int round = 0; round < 8; ++round) {
    int open = 0;
    for (int c = Blocks; c <= Residual; ++c) if (hot[c] < p.nent[c]) ++open;
    if (!open || !left) break;
    const uint32_t share = std::max<uint32_t>(1, left / open);
    for (int c = Blocks; c <= Residual; ++c) {
      const uint32_t want = std::min<uint32_t>(p.nent[c] - hot[c], std::min(share, left));
      hot[c] += want;
      left -= want;
    }
  }
  uint32_t eo = 0;
  for (int c = Blocks; c <= Residual; ++c) { kt.cp[c].z = eo; kt.cp[c].w = hot[c]; eo += hot[c]; }
  kt.ent_words = eo;
  return head + 8ull * bo + 8ull * eo;
}

uint64_t tile_smem_bytes(const TileCfg& c) {
  return 4ull * ((uint64_t)c.slotcap + 8 + 32ull * (c.rowcap + 1) + HRECCAP + 8 + 2ull * (c.nt + 2) + 10ull * c.nt + 32);
}

}  // namespace

// K0 .. K2 on the nodes described by rv (a contiguous range, or a sorted node list), then one host
// synchronisation that reads back the totals (tot[0] = halo arcs, tot[1] = all arcs), the hard-node count and
// the error word.
static void run_pipeline(wga_graph* g, RangeView& rv, uint8_t* w, const WorkspacePlan& p, Scalars* sc,
                         const Tuning& tn, cudaStream_t st, uint64_t tot[2]) {
  const uint64_t n = rv.n;
  const DeviceInfo& di = device_info(g->device);
  // ---- K0 + scan
  k_heads<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, rv.lo, rv.nodes, (uint32_t)n, rv.outdeg, rv.nrec, g->d_err);
  count_launch();
  {
    size_t cb = p.cub_bytes;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(rv.outdeg, U32ToU64());
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, rv.offs, (int64_t)(n + 1), st));
    count_launch(2);
  }
  mark(g, st);  // 1: heads + scan done
  // ---- K1: entropy decode into rows
  {
    K1Tables kt{};
    const uint64_t smem = plan_k1_tables(g, kt, di.smem_optin);
    uint32_t blocks = tn.k1_blocks ? tn.k1_blocks : (uint32_t)di.sms;
    blocks = std::min<uint32_t>(blocks, (rv.n_units + K1_WARPS - 1) / K1_WARPS);
    blocks = std::max<uint32_t>(1u, std::min<uint32_t>(blocks, MAX_STREAMS / K1_WARPS));
    bool allhot = true;
    for (int c = Blocks; c <= Residual; ++c) allhot = allhot && kt.cp[c].w >= g->packed.nent[c];
    const uint32_t refill = std::max<uint32_t>(1u, std::min<uint32_t>(32u, tn.refill));
    auto launch = [&](auto kern) {
      WGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin - 1024));
      kern<<<blocks, K1_THREADS, smem, st>>>(g->dev, rv, kt, refill);
    };
    if (rv.nodes) { if (allhot) launch(k_entropy<true, true>); else launch(k_entropy<true, false>); }
    else { if (allhot) launch(k_entropy<false, true>); else launch(k_entropy<false, false>); }
    count_launch();
  }
  mark(g, st);  // 2: entropy decode done
  // ---- K2: tiles
  static unsigned long long* d_stats = nullptr;
  static const bool want_stats = getenv("WGA_K2_STATS") != nullptr;
  if (want_stats) {
    if (!d_stats) WGA_CUDA(cudaMalloc(&d_stats, 16 * 8));
    WGA_CUDA(cudaMemsetAsync(d_stats, 0, 16 * 8, st));
    rv.stats = d_stats;
  }
  const uint32_t window = (uint32_t)g->prelude.compression_window;
  const uint32_t lookback = std::min<uint32_t>(HMAX, window * LMAXT);
  {
    TileCfg cfg{rv.unit, (std::max<uint32_t>(64u, tn.slotcap) + 3u) & ~3u, std::max<uint32_t>(3u, tn.rowcap),
                std::min<uint32_t>(std::max<uint32_t>(2u, tn.dbig), 32768u), tile_threads(tn, window),
                std::max<uint32_t>(4u, tn.seg), 0u, tn.dbg};
    // every candidate is at least one task; the referenced lists and the residuals each fit the slots
    cfg.taskcap = (cfg.nt + 2 * (cfg.slotcap / cfg.seg) + 8 + 3) & ~3u;
    const uint64_t smem = tile_smem_bytes(cfg);
    if ((int64_t)smem > (int64_t)di.smem_optin - 2048) throw Error(WGA_E_ARG, "tile tuning exceeds shared memory");
    auto launch = [&](auto kern) {
      WGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<rv.n_units, cfg.nt, smem, st>>>(rv, cfg, g->dev.min_interval, lookback);
    };
    if (cfg.nt == 128) launch(k_tile<128>);
    else if (cfg.nt == 160) launch(k_tile<160>);
    else if (cfg.nt == 192) launch(k_tile<192>);
    else launch(k_tile<256>);
    count_launch();
  }
  mark(g, st);  // 3: tiles done
  // ---- totals, hard nodes, error word
  k_publish<<<1, 32, 0, st>>>(rv.offs + rv.h, rv.offs + n, &sc->maxlevel, &sc->hard_count, g->d_err, g->d_pub);
  count_launch();
  WGA_CUDA(cudaStreamSynchronize(st));
  WGA_CUDA(cudaGetLastError());
  tot[0] = g->h_pub[1];
  tot[1] = g->h_pub[2];
  if (want_stats) {
    unsigned long long hs[16];
    WGA_CUDA(cudaMemcpy(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost));
    const double nb = hs[15] ? (double)hs[15] : 1.0;
    fprintf(stderr, "[k_tile cycles/block] load %.0f plan %.0f stage %.0f tasks %.0f lev0 %.0f lev1 %.0f lev2 %.0f lev3+ %.0f out %.0f (blocks %llu)\n",
            hs[0] / nb, hs[1] / nb, hs[2] / nb, hs[3] / nb, hs[4] / nb, hs[5] / nb, hs[6] / nb, hs[7] / nb, hs[8] / nb, hs[15]);
  }
  uint32_t herr = (uint32_t)g->h_pub[4];
  const uint32_t nhard = (uint32_t)g->h_pub[5];
  if (tot[0] > rv.halo_cap) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "halo successors exceed the workspace");
  }
  if (tot[1] - tot[0] > rv.succ_cap) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(tot[1] - tot[0]) + " elements");
  }
  check_device_error(g, herr, st);
  // ---- nodes the tiles left to the global pass, one launch per depth
  if (nhard) {
    k_hard_levels<<<(nhard + 255) / 256, 256, 0, st>>>(rv);
    k_publish<<<1, 32, 0, st>>>(rv.offs + rv.h, rv.offs + n, &sc->maxlevel, &sc->hard_count, g->d_err, g->d_pub);
    count_launch(2);
    WGA_CUDA(cudaStreamSynchronize(st));
    const uint32_t maxlevel = (uint32_t)g->h_pub[3];
    for (uint32_t l = 0; l <= maxlevel; ++l) {
      k_hard_resolve<<<(nhard + 127) / 128, 128, 0, st>>>(rv, l, g->dev.min_interval);
      count_launch();
    }
    check_device_error(g, read_device_error(g, st), st);
  }
  mark(g, st);  // 4: global pass done
}

static void apply_env_tuning() {
  static std::once_flag once;
  std::call_once(once, [] {  // WGA_TUNING="key=value,key=value": same knobs as wga_debug_set_tuning (profiling runs)
    if (const char* e = getenv("WGA_TUNING")) {
      std::string str(e);
      size_t i = 0;
      while (i < str.size()) {
        size_t j = str.find(',', i);
        if (j == std::string::npos) j = str.size();
        size_t q = str.find('=', i);
        if (q != std::string::npos && q < j) set_tuning(str.substr(i, q - i).c_str(), strtoull(str.c_str() + q + 1, nullptr, 10));
        i = j + 1;
      }
    }
  });
}

static void bind_views(RangeView& rv, uint8_t* w, const WorkspacePlan& p, Scalars* sc, uint64_t ws_bytes, uint32_t unit) {
  rv.outdeg = (uint32_t*)(w + p.off_outdeg);
  rv.nrec = (uint4*)(w + p.off_nrec);
  rv.meta = (uint2*)(w + p.off_meta);
  rv.hardflag = (uint8_t*)(w + p.off_hflag);
  rv.hard_list = (uint32_t*)(w + p.off_hlist);
  rv.hard_lev = (uint32_t*)(w + p.off_hlev);
  rv.hard_count = &sc->hard_count;
  rv.maxlevel = &sc->maxlevel;
  rv.rows = (uint32_t*)(w + p.off_rows);
  const uint64_t chunks = (ws_bytes - p.off_rows) / (4ull * CH * 32);
  rv.rows_cap = (uint32_t)std::min<uint64_t>(chunks, 0xFFFFFFF0ull);
  rv.chunk_ctr = &sc->chunk_ctr;
  rv.stream_chunks = (uint32_t*)(w + p.off_schunks);
  rv.unit_stream = (uint32_t*)(w + p.off_ustream);
  rv.unit_ctr = &sc->unit_ctr;
  rv.unit = unit;
  rv.n_units = (rv.n + unit - 1) / unit;
  rv.halo_succ = (uint32_t*)(w + p.off_halo);
  rv.halo_cap = p.halo_cap;
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
  apply_env_tuning();
  const Tuning tn = g_tuning;
  uint8_t* w = (uint8_t*)ws;
  if (ws_bytes < 256) throw Error(WGA_E_WORKSPACE, "workspace too small");
  WGA_CUDA(cudaMemsetAsync(w, 0, 256, st));
  Scalars* sc = (Scalars*)w;
  g->n_ev = 0;
  mark(g, st);  // 0: start
  // ---- halo
  uint64_t lo = first;
  if (first > g->res_first && g->prelude.compression_window != 0) {
    k_halo<<<1, 32, 0, st>>>(g->dev, first, last, g->d_pub, g->d_err);
    count_launch();
    WGA_CUDA(cudaStreamSynchronize(st));
    lo = g->h_pub[0];
    if (lo < g->res_first) throw Error(WGA_E_ARG, "reference chain leaves the resident shard");
  }
  const uint64_t n = last - lo;
  const uint32_t unit = effective_tile(tn, (uint32_t)g->prelude.compression_window);
  WorkspacePlan p = plan_workspace(n, unit);
  if (ws_bytes < p.fixed_bytes + 4ull * CH * 32) throw Error(WGA_E_WORKSPACE, "workspace too small");
  RangeView rv{};
  rv.lo = lo; rv.n = (uint32_t)n; rv.h = (uint32_t)(first - lo);
  bind_views(rv, w, p, sc, ws_bytes, unit);
  rv.offs = rv.h ? (uint64_t*)(w + p.off_offs) : d_offsets;
  rv.succ = d_succ; rv.succ_cap = d_succ ? succ_capacity : 0;
  rv.err = g->d_err;
  uint64_t tot[2] = {0, 0};
  run_pipeline(g, rv, w, p, sc, tn, st, tot);
  if (rv.h) {  // hand the caller offsets relative to `first`
    const uint64_t cnt = last - first + 1;
    k_offsets_rebase<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(rv.offs + rv.h, tot[0], d_offsets, cnt);
    count_launch();
  }
  WGA_CUDA(cudaGetLastError());
  if (g->profiling) {
    WGA_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) g->stage_ms[i] = 0.f;
    for (int i = 1; i < g->n_ev; ++i) cudaEventElapsedTime(&g->stage_ms[i - 1], g->ev[i - 1], g->ev[i]);
  }
  if (h_arcs) *h_arcs = tot[1] - tot[0];
}

// ---------------------------------------------------------------------------------------------- random access
namespace {
struct BatchPlan {
  uint64_t cap_nodes, cap_arcs;
  uint64_t off_scal, off_qid, off_all[2], off_U, off_qidx, off_cub, off_offsU, off_succU, off_inner;
  uint64_t cub_bytes, inner_bytes, total;
};
BatchPlan plan_batch(uint64_t nq, uint64_t max_total_arcs, uint32_t unit) {
  BatchPlan b{};
  b.cap_nodes = 8 * nq + 4096;       // queries + every node on their reference chains
  if (b.cap_nodes > 0xFFFFFFF0ull) b.cap_nodes = 0xFFFFFFF0ull;
  b.cap_arcs = 4 * max_total_arcs + 65536;
  uint64_t o = 0;
  b.off_scal = o; o += 256;
  b.off_qid = o; o = align_up(o + 4 * (nq + 1), 256);
  for (int i = 0; i < 2; ++i) { b.off_all[i] = o; o = align_up(o + 4 * b.cap_nodes, 256); }
  b.off_U = o; o = align_up(o + 4 * b.cap_nodes, 256);
  b.off_qidx = o; o = align_up(o + 4 * (nq + 1), 256);
  size_t c1 = 0, c2 = 0, c3 = 0;
  cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr);
  cub::DeviceRadixSort::SortKeys(nullptr, c1, dk, (int64_t)b.cap_nodes, 0, 32);
  cub::DeviceSelect::Unique(nullptr, c2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)b.cap_nodes);
  cub::DeviceScan::ExclusiveSum(nullptr, c3, (uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t)(nq + 1));
  b.cub_bytes = std::max(c1, std::max(c2, c3)) + 4096;
  b.off_cub = o; o = align_up(o + b.cub_bytes, 256);
  b.off_offsU = o; o = align_up(o + 8 * (b.cap_nodes + 1), 256);
  b.off_succU = o; o = align_up(o + 4 * b.cap_arcs, 256);
  WorkspacePlan p = plan_workspace(b.cap_nodes, unit);
  b.inner_bytes = p.fixed_bytes + 8 * b.cap_arcs + 48 * b.cap_nodes + (16ull << 20) +
                  std::min<uint64_t>((b.cap_nodes + unit - 1) / unit, MAX_STREAMS) * (4ull * CH * 32);
  b.off_inner = align_up(o, 16384); o = b.off_inner + b.inner_bytes;
  b.total = o;
  return b;
}
}  // namespace

uint64_t successors_workspace_size(const wga_graph* g, uint64_t n_queries, uint64_t max_total_arcs) {
  return plan_batch(n_queries, max_total_arcs, effective_tile(g_tuning, (uint32_t)g->prelude.compression_window)).total;
}

void successors_batch(wga_graph* g, const uint64_t* d_nodes, uint64_t nq, uint64_t* d_offsets, uint32_t* d_succ,
                      uint64_t succ_capacity, void* ws, uint64_t ws_bytes, uint64_t* h_arcs, cudaStream_t st) {
  if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  if (nq >= 0xFFFFFFF0ull) throw Error(WGA_E_UNSUPPORTED, "too many queries for one call");
  if (nq == 0) {
    WGA_CUDA(cudaMemsetAsync(d_offsets, 0, 8, st));
    if (h_arcs) *h_arcs = 0;
    return;
  }
  apply_env_tuning();
  const Tuning tn = g_tuning;
  const uint32_t unit = effective_tile(tn, (uint32_t)g->prelude.compression_window);
  // the caller sizes the workspace with an upper bound of the arcs it expects; recover it from the size
  BatchPlan b = plan_batch(nq, 0, unit);
  if (ws_bytes < b.total) throw Error(WGA_E_WORKSPACE, "workspace too small");
  {  // largest arc bound whose plan fits this workspace (the plan grows by ~50 bytes per arc)
    uint64_t A = (ws_bytes - b.total) / 50;
    BatchPlan b2 = plan_batch(nq, A, unit);
    while (b2.total > ws_bytes && A) { A = A / 16 * 15; b2 = plan_batch(nq, A, unit); }
    if (b2.total <= ws_bytes) b = b2;
  }
  uint8_t* w = (uint8_t*)ws;
  uint32_t* scal = (uint32_t*)(w + b.off_scal);  // [0] closure count, [1] unique count
  WGA_CUDA(cudaMemsetAsync(scal, 0, 256, st));
  uint32_t* qid = (uint32_t*)(w + b.off_qid);
  uint32_t* all = (uint32_t*)(w + b.off_all[0]);
  k_query_ids<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_nodes, nq, g->res_first, g->res_last, qid, g->d_err);
  count_launch();
  if (!d_succ) {  // sizing call: only the outdegrees of the queries (first symbol of each record)
    uint32_t* deg = all;
    k_heads<<<(unsigned)((nq + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, 0, qid, (uint32_t)nq, deg, nullptr, g->d_err);
    size_t cbs = b.cub_bytes;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(deg, U32ToU64());
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + b.off_cub, cbs, it, d_offsets, (int64_t)(nq + 1), st));
    count_launch(3);
    uint64_t arcs0 = 0;
    WGA_CUDA(cudaMemcpyAsync(&arcs0, d_offsets + nq, 8, cudaMemcpyDeviceToHost, st));
    check_device_error(g, read_device_error(g, st), st);
    if (h_arcs) *h_arcs = arcs0;
    return;
  }
  WGA_CUDA(cudaMemcpyAsync(all, qid, 4 * nq, cudaMemcpyDeviceToDevice, st));
  // ---- closure under "referenced node": one round per chain level
  uint64_t total = nq, n_in = nq, in_off = 0;
  while (n_in) {
    if (total >= b.cap_nodes) throw Error(WGA_E_WORKSPACE, "reference closure exceeds the workspace");
    WGA_CUDA(cudaMemsetAsync(scal, 0, 4, st));
    k_closure_step<<<(unsigned)((n_in + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, all + in_off, (uint32_t)n_in, all + total,
                                                                       scal, (uint32_t)(b.cap_nodes - total),
                                                                       g->res_first, g->d_err);
    count_launch();
    uint32_t cnt = 0;
    WGA_CUDA(cudaMemcpyAsync(&cnt, scal, 4, cudaMemcpyDeviceToHost, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    if (total + cnt > b.cap_nodes) throw Error(WGA_E_WORKSPACE, "reference closure exceeds the workspace");
    in_off = total;
    n_in = cnt;
    total += cnt;
  }
  // ---- sort + unique -> U
  cub::DoubleBuffer<uint32_t> dk(all, (uint32_t*)(w + b.off_all[1]));
  size_t cb = b.cub_bytes;
  int end_bit = 1;
  while (end_bit < 32 && (g->prelude.number_of_nodes >> end_bit)) ++end_bit;
  WGA_CUDA(cub::DeviceRadixSort::SortKeys(w + b.off_cub, cb, dk, (int64_t)total, 0, end_bit, st));
  uint32_t* U = (uint32_t*)(w + b.off_U);
  cb = b.cub_bytes;
  WGA_CUDA(cub::DeviceSelect::Unique(w + b.off_cub, cb, dk.Current(), U, scal + 1, (int64_t)total, st));
  count_launch(6);
  uint32_t nU = 0;
  WGA_CUDA(cudaMemcpyAsync(&nU, scal + 1, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  check_device_error(g, read_device_error(g, st), st);
  // ---- decode the node list U into a temporary CSR
  uint8_t* iw = w + b.off_inner;
  WGA_CUDA(cudaMemsetAsync(iw, 0, 256, st));
  WorkspacePlan p = plan_workspace(nU, unit);
  if (p.fixed_bytes + 4ull * CH * 32 > b.inner_bytes) throw Error(WGA_E_WORKSPACE, "workspace too small");
  Scalars* sc = (Scalars*)iw;
  RangeView rv{};
  rv.lo = 0; rv.n = nU; rv.h = 0; rv.nodes = U;
  bind_views(rv, iw, p, sc, b.inner_bytes, unit);
  rv.offs = (uint64_t*)(w + b.off_offsU);
  rv.halo_succ = nullptr; rv.halo_cap = 0;
  rv.succ = (uint32_t*)(w + b.off_succU); rv.succ_cap = b.cap_arcs;
  rv.err = g->d_err;
  g->n_ev = 0;
  uint64_t tot[2] = {0, 0};
  const bool prof = g->profiling;
  g->profiling = false;
  try { run_pipeline(g, rv, iw, p, sc, tn, st, tot); } catch (...) { g->profiling = prof; throw; }
  g->profiling = prof;
  // ---- gather the query lists
  uint32_t* qidx = (uint32_t*)(w + b.off_qidx);
  k_query_lookup<<<(unsigned)((nq + 1 + 255) / 256), 256, 0, st>>>(qid, nq, U, nU, rv.offs, qidx, d_offsets);
  cb = b.cub_bytes;
  WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + b.off_cub, cb, d_offsets, d_offsets, (int64_t)(nq + 1), st));
  uint64_t arcs = 0;
  WGA_CUDA(cudaMemcpyAsync(&arcs, d_offsets + nq, 8, cudaMemcpyDeviceToHost, st));
  k_query_gather<<<148 * 8, 256, 0, st>>>(qidx, nq, rv.offs, rv.succ, d_offsets, d_succ, succ_capacity);
  count_launch(4);
  WGA_CUDA(cudaStreamSynchronize(st));
  WGA_CUDA(cudaGetLastError());
  if (arcs > succ_capacity) throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(arcs) + " elements");
  if (h_arcs) *h_arcs = arcs;
}

void launch_offsets_add(uint64_t* off, uint64_t n, uint64_t base, cudaStream_t st) {
  if (!base || !n) return;
  k_offsets_add<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(off, n, base);
  count_launch();
}

void debug_expand_table(wga_graph* g, int c, void* h_out, uint64_t n_slots) {
  if (c < 0 || c >= WGA_COMPONENTS) throw Error(WGA_E_ARG, "bad component");
  uint64_t want = 1ull << g->prelude.tables[c].frame_size;
  if (n_slots != want) throw Error(WGA_E_ARG, "n_slots must be 2^frame_size");
  if (!g->on_device) {  // host-only handle: the same bucket + popcount lookup on the host copy of the packed tables
    const PackedTablesData& p = g->packed;
    uint32_t* o = (uint32_t*)h_out;
    for (uint32_t slot = 0; slot < n_slots; ++slot, o += 4) {
      const Bkt& bk = p.bkt[p.bkt_off[c] + (slot >> 5)];
      const uint32_t j = bk.j0 + (uint32_t)__builtin_popcount(bk.mask & ((2u << (slot & 31u)) - 1u));
      const Ent& e = p.ent[p.ent_off[c] + j];
      const uint32_t folds = e.bf >> 16;
      if (folds == 0xFFFFu) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
      const uint64_t q = ((uint64_t)(e.bf & 0xFFFFu) << (folds * p.R[c])) | ((uint64_t)folds << 48);
      o[0] = (e.cf >> 16) | ((e.cf & 0xFFFFu) << 16);
      o[1] = 0;
      o[2] = (uint32_t)q;
      o[3] = (uint32_t)(q >> 32);
    }
    return;
  }
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
  k_decode_symbols<<<1, 1>>>(g->dev, dc, n, ptr - g->stream_base, state, dout, dend);
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
