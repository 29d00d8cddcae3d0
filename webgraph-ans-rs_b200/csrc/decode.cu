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
//                      (states[N-1-v], pointers[N-1-v]) -- outdegree, reference offset, block count -- the decoder
//                      state after it, the prefetched next stream word and the phase the record must end at
//        cub scans     outdegrees -> CSR offsets ; (block count + 2 + outdegree) -> record regions
//        k_plan        one thread per node: slot and record pointers, outdegree of the referenced node, head
//                      validation -- so that K1's per-node set-up is three independent loads
//    K1  k_entropy     phase one: entropy decode of the rest of every record.  Persistent kernel, one 1024-thread
//                      block per SM with the decoder tables of the six components in SHARED memory (bucket +
//                      popcount lookup, no search).  Warps are independent: each pulls units of consecutive nodes
//                      from a global counter, its lanes take the nodes one by one (ballot-ranked, no atomics) and
//                      run a per-symbol state machine; every busy lane decodes ONE symbol per iteration and
//                      stores one word: cumulative copy-block ends, interval count, interval starts / lengths into
//                      the node's record, prefix-summed residuals into the tail of the node's own CSR slot.
//                      Reference-free lists without intervals are final after K1.  Every record must end exactly
//                      at the phase of the next node: that is the corruption check.
//    K2  k_levels      phase two, by reference-chain depth: depth of every node that still needs work
//        cub sort      stable sort by level -> one segment per level, node order kept inside a level
//        k_resolve     per level, one node per lane: three-way merge (copied elements of the finished
//                      referenced list, expanded intervals, residuals) into the node's CSR slot, up to four
//                      elements of one run per step, output staged per lane and written as whole sectors
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
  uint32_t unit = 64;         // nodes per K1 unit (what a warp pulls from the global counter)
  uint32_t k1_blocks = 0;     // K1 grid; 0 = one block per SM
  uint32_t refill = 10;       // K1: lanes that must be free before the warp fetches new nodes (swept: 3..16)
  uint32_t k2_blocks = 0;     // K2 grid; 0 = one full wave (SM count x resident blocks per SM): every block gets an
                              // equal share of the level, so a partial second wave would double the time
  uint32_t k2_batch = 14;     // K2: lanes that must be free before the warp sets up new nodes (swept: 4..24)
  uint32_t hub_min = 2048;    // K2: lists at least this long are merged by a whole warp (k_resolve, hub pass)
  uint32_t e2e_chunk = 1u << 19;  // nodes per chunk of the pipelined host entry point
};
static Tuning g_tuning;
static std::mutex g_tuning_mu;  // wga_debug_set_tuning may race with decode calls of other threads
static Tuning tuning_snapshot() {
  std::lock_guard<std::mutex> lk(g_tuning_mu);
  return g_tuning;
}

int set_tuning(const char* key, uint64_t value) {
  std::string k(key ? key : "");
  std::lock_guard<std::mutex> lk(g_tuning_mu);
  if (k == "unit") g_tuning.unit = (uint32_t)value;
  else if (k == "k1_blocks") g_tuning.k1_blocks = (uint32_t)value;
  else if (k == "refill") g_tuning.refill = (uint32_t)value;
  else if (k == "k2_blocks") g_tuning.k2_blocks = (uint32_t)value;
  else if (k == "k2_batch") g_tuning.k2_batch = (uint32_t)value;
  else if (k == "hub_min") g_tuning.hub_min = (uint32_t)value;
  else if (k == "e2e_chunk") g_tuning.e2e_chunk = (uint32_t)value;
  else if (k == "reset") g_tuning = Tuning();
  else return WGA_E_ARG;
  return WGA_OK;
}
uint64_t tuning_e2e_chunk() { return tuning_snapshot().e2e_chunk; }

namespace {

constexpr int TPB = 128;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t INF = 0xffffffffu;  // "stream exhausted"; successor ids are <= 0xfffffffe
constexpr uint32_t NOT_FOUND = 0xFFFFFFFFu;

// ---- K1 output: records --------------------------------------------------------------------------------
// Every node owns a region of (block count + 2 + outdegree) words in the record buffer (offsets: a second scan
// of the K0 heads); K1 writes the node's record there word by word:
//   [cumulative copy-block ends x b][interval count][start,len x ni][number of residuals]
// (2 ni <= outdegree because an interval covers at least two successors).  The residuals themselves (already
// prefix-summed) are parked at the tail of the node's own CSR slot.
constexpr int K1_THREADS = 1024;
constexpr uint32_t K1_WARPS = K1_THREADS / 32;
// record word of node t: words written (24 bits) | flags << 29
constexpr uint32_t MF_ERR = 1u, MF_INSLOT = 2u, MF_FINAL = 4u;
constexpr uint32_t NSYM_MAX = (1u << 24) - 1;
// head word of node t (K0): reference offset in node-list positions (12 bits) | block count << 12
constexpr uint32_t RT_BITS = 12, RT_MASK = (1u << RT_BITS) - 1, B_MAX = (1u << (32 - RT_BITS)) - 1;

struct RangeView {
  uint64_t lo;        // first decoded node (halo start)
  uint32_t n;         // nodes decoded: last - lo
  const uint32_t* nodes;  // nullptr: node t is lo + t; else a sorted, duplicate-free list of node ids (random access)
  uint32_t h;         // halo nodes: first - lo
  uint32_t* outdeg;   // n+1
  uint4* nrec;        // n : from K0: decoder (state, stream index) after the record's head, outdegree, head word
  uint4* nrec2;       // n : from K0: prefetched stream word, phase (state, stream index) the record must end at
                      //     (state 0: not checked); from k_plan: .w = aux (see k_plan)
  uint64_t* offs;     // n+1, relative to lo
  uint32_t* meta;     // n : record word of K1
  uint64_t* roff;     // n+1 : record regions
  uint32_t* recs;     // record buffer
  uint64_t recs_cap;  // words
  uint32_t* maxlevel;   // deepest reference chain seen by k_levels (only tracked from LCAP up)
  uint2* hubs;          // (node, level) of the long lists that K2 merges with a whole warp each
  uint32_t* hub_count;
  uint32_t hub_min;
  uint32_t* unit_ctr;
  uint32_t unit;        // nodes per K1 unit
  uint32_t n_units;
  uint32_t* halo_succ;  // successors of halo nodes
  uint64_t halo_cap;
  uint32_t* succ;       // caller's array: successors of nodes >= first
  uint64_t succ_cap;
  uint32_t* err;
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

// (state, pointer) of node v: ANSBVGraphDecoderFactory::new_decoder (bvgraph_decoder_factory.rs:46-58)
__device__ __forceinline__ void load_phase(const DevGraph& g, uint64_t v, Dec& d, uint32_t& err) {
  d.state = g.states[g.top - v];
  uint64_t p = g.ptrs[g.top - v] - g.stream_base;
  if (p > g.stream_words) { err |= ERR_CORRUPT; p = 0; }
  d.sp = (uint32_t)p;  // the resident span has < 2^32 words (checked at upload)
  dec_prime(d, g.stream);
}

// A record ends where the next one begins: after the last symbol of node v the decoder must stand at the phase of
// node v+1 (the encoder runs through the whole graph with one state, encoder.rs / bvgraph_decoder_factory.rs:46-58),
// which any corruption of the record breaks.  Returns that phase as (state, stream index); state 0 = not checked:
// the last node of a shard (the next phase is not resident) and the last node of a contiguous range (the
// pipelined host entry point decodes a chunk while the phases of the next one are still on their way to the
// device).  The last node of the graph ends at the initial encoder state.
__device__ __forceinline__ uint2 expected_end(const DevGraph& g, bool last_of_range, uint64_t v) {
  if (v + 1 > g.top) return (g.top + 1 == g.N && g.stream_base == 0) ? make_uint2(WGA_LOWER_BOUND, 0u) : make_uint2(0u, 0u);
  if (last_of_range) return make_uint2(0u, 0u);
  uint64_t p = g.ptrs[g.top - (v + 1)] - g.stream_base;
  if (p > g.stream_words) return make_uint2(1u, 0u);  // (never matches)
  uint32_t s = g.states[g.top - (v + 1)];
  // The encoder's upper bound wraps in 32 bits for frames below 2^16 (component_model4encoder.rs:28-34), which can
  // leave a recorded state below 2^16; the decoder arriving there has already extended it (decoder.rs:67-69).
  if (s < WGA_LOWER_BOUND && p != 0) { --p; s = (s << 16) | (uint32_t)g.stream[p]; }
  return make_uint2(s, (uint32_t)p);
}

// -------------------------------------------------------------------------------------------- K0
// One lane per node, every lane at the same symbol: the fixed-shape head of a record -- outdegree, reference
// offset, block count -- and the decoder state after it.  The block count is validated by K1, which knows the
// outdegree of the referenced node.
__global__ void __launch_bounds__(TPB) k_heads(DevGraph g, uint64_t lo, const uint32_t* nodes, uint32_t n,
                                               uint32_t* outdeg, uint4* nrec, uint4* nrec2, uint32_t* err_out) {
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
    const uint2 ee = expected_end(g, !nodes && t + 1 == n, v);
    nrec2[t] = make_uint4(dc.w, ee.x, ee.y, 0u);
  }
  if (err) atomicOr(err_out, err);
}

struct U32ToU64 {
  __host__ __device__ uint64_t operator()(uint32_t x) const { return (uint64_t)x; }
};
// words of a node's record region: the copy blocks, the interval count, two words per interval, the residual count.
// An interval covers at least min_interval_length successors: two words per interval are at most one per successor
// (two when the minimum length is one).
__host__ __device__ inline uint64_t rec_words(uint32_t b, uint32_t d, uint32_t minint) {
  return (uint64_t)b + 2u + (minint == 1u ? 2ull * d : (uint64_t)d);
}
struct RecWords {
  uint32_t minint;
  __host__ __device__ uint64_t operator()(const uint4& r) const { return rec_words(r.w >> RT_BITS, r.z, minint); }
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

// Per-lane machine states of the entropy kernel: where in the record the lane is (the first interval start and the
// first residual are nat-coded differences to the node id, so they get states of their own).
enum : uint32_t { S_BLK = 0, S_ICNT, S_IST0, S_IST, S_ILEN, S_RES0, S_RES, S_FREE, S_COUNT = S_FREE };
__device__ __constant__ uint8_t c_state_comp[S_COUNT] = {Blocks, IntervalCount, IntervalStart, IntervalStart,
                                                         IntervalLen, FirstResidual, Residual};

// What K1 needs to start the record of node t, prepared by k_plan (which sees the scans and the outdegree of the
// referenced node) so that K1's set-up is three independent loads and no dependent one.
struct NodePlan {
  uint32_t* slot_end;  // one past the node's CSR slot
  uint32_t* rec;       // the node's region of the record buffer; nullptr: nothing to decode in K1 (meta is final)
};

// One thread per node, after the scans.  Decides which nodes have symbols left for K1 and validates the head:
//   no successors, or a pure copy of the referenced list (no blocks, same outdegree)  -> meta = 0, nothing for K1
//   block count > outdegree of the referenced node + 1, pure copy longer than the list -> corrupt
//   slot or record region beyond the buffers                                         -> workspace error
// aux (nrec2[t].w) = outdegree of the referenced node (records with copy blocks) or the number of successors that
// are not copied (records without).  Records without K1 work must already stand at the phase of the next node.
__global__ void __launch_bounds__(256) k_plan(DevGraph g, RangeView rv, NodePlan* plan) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rv.n) return;
  const uint4 p = rv.nrec[t];
  const uint32_t d = p.z, rt = p.w & RT_MASK, b = p.w >> RT_BITS;
  NodePlan pl{nullptr, nullptr};
  uint32_t a = 0, err = 0, m = 0;
  bool work = false;
  if (d != 0) {
    uint32_t* slot = node_slot(rv, t);
    const uint64_t ro = rv.roff[t];
    if (!slot || ro + rec_words(b, d, g.min_interval) > rv.recs_cap) err = ERR_WORKSPACE;
    else if (rt == 0) { a = d; work = true; }
    else {
      const uint32_t dref = rv.outdeg[t - rt];
      if (b > dref && b - dref > 1u) err = ERR_CORRUPT;  // at most dref + 1 blocks
      else if (b != 0) { a = dref; work = true; }
      else if (dref > d) err = ERR_CORRUPT;
      else { a = d - dref; work = a != 0; }
    }
    if (work) pl = NodePlan{slot + d, rv.recs + ro};
  }
  if (!work && !err) {
    const uint4 q = rv.nrec2[t];
    if (q.y != 0 && (q.y != p.x || q.z != p.y)) err = ERR_CORRUPT;
  }
  if (err) { atomicOr(rv.err, err); m = MF_ERR << 29; pl = NodePlan{nullptr, nullptr}; }
  if (!pl.rec) rv.meta[t] = m;
  plan[t] = pl;
  rv.nrec2[t].w = a;
}

// LIST: node t is rv.nodes[t] (random access) instead of rv.lo + t.  ALLHOT: every table entry is in shared memory.
// Every busy lane decodes one symbol per iteration.  What follows the table lookup is the same instruction sequence
// for every state -- the value is one of
//   gap  prev + 1 + x     copy-block ends (cumulative: prev starts at -1), later interval starts, residuals
//   nat  v + nat2int(x)   first interval start, first residual
//   len  x (+ min_interval_length)   interval count, interval length
// and the counters and the next state follow from a few selects and one packed transition table.
// Values are not range-checked one by one: what is checked is what memory safety needs (interval count and lengths
// and copied elements within the outdegree) and, at the end of every record, that the decoder stands exactly at
// the phase of the next node -- which any corruption of the record breaks.
template <bool LIST, bool ALLHOT>
__global__ void __launch_bounds__(K1_THREADS, 1) k_entropy(DevGraph g, RangeView rv, K1Tables kt, const NodePlan* plan,
                                                           uint32_t refill_min) {
  extern __shared__ __align__(16) unsigned char k1_smem[];
  uint4* s_cp = reinterpret_cast<uint4*>(k1_smem);                        // one per machine state (8 x 16 B, + 1 spare)
  uint32_t* s_goff = reinterpret_cast<uint32_t*>(k1_smem + 9 * 16);       // 9 (+ pad to 12)
  uint32_t* s_recip = s_goff + 12;                                         // 32
  uint2* s_bkt = reinterpret_cast<uint2*>(s_recip + 32);
  uint2* s_ent = s_bkt + kt.bkt_words;
  {
    if (threadIdx.x < S_COUNT) {
      const uint32_t comp = c_state_comp[threadIdx.x];
      uint4 c = kt.cp[comp];
      c.x |= comp << 26;  // component index for the cold-entry path
      s_cp[threadIdx.x] = c;
    }
    if (threadIdx.x < WGA_COMPONENTS) s_goff[threadIdx.x] = kt.gent_off[threadIdx.x];
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
  const uint32_t minint = g.min_interval;
  const uint32_t s_extras = minint ? (uint32_t)S_ICNT : (uint32_t)S_RES0;  // what follows the copy blocks
  const uint32_t lo32 = (uint32_t)rv.lo;
  const uint16_t* __restrict__ const stream = g.stream;
  // transition table, 4 bits per (state, run counter k != 0): next state | 8 when the record may end here
  //   BLK: k ? BLK : extras-state*   ICNT: k ? IST0 : RES0*   IST0, IST: ILEN   ILEN: k ? IST : RES0*   RES0, RES: RES*
  uint64_t trans = 0;
  {
    const uint32_t e[14] = {s_extras | 8u, S_BLK, S_RES0 | 8u, S_IST0, S_ILEN, S_ILEN, S_ILEN, S_ILEN,
                            S_RES0 | 8u, S_IST, S_RES | 8u, S_RES | 8u, S_RES | 8u, S_RES | 8u};
#pragma unroll
    for (int i = 0; i < 14; ++i) trans |= (uint64_t)e[i] << (4 * i);
  }

  // warp-uniform
  uint32_t nx = 0, ne = 0;      // nodes [nx, ne) of the current unit are not yet handed out
  bool exhausted = false;
  // per-lane record state
  uint32_t c = S_FREE, t = 0, prev = 0, d = 0, k = 0, bb = 0, copied = 0, sgn = 1, extras = 0, ni = 0, ns = 0;
  uint32_t vl = 0;               // LIST: the node id
  uint32_t end_state = 0, end_sp = 0;  // where the decoder must stand after the record (end_state 0: not checked)
  uint32_t* slot_end = nullptr;  // the residuals are parked at the tail of the node's own slot (MF_INSLOT)
  uint32_t* recw = nullptr;      // the node's region of the record buffer
  Dec dc{0, 0, 0};

  for (;;) {
    // ---------------------------------------------------------------- hand out nodes
    const uint32_t busy = __ballot_sync(FULL, c != S_FREE);
    if (!exhausted) {
      const uint32_t cnt = 32u - (uint32_t)__popc(busy);
      if (cnt >= refill_min) {
        const uint32_t r = (uint32_t)__popc(~busy & lt_mask);  // rank among the free lanes
        uint32_t my = NOT_FOUND;
        uint32_t taken = 0;
        while (taken < cnt) {  // (uniform) the free lanes may straddle a unit boundary
          if (nx >= ne) {
            uint32_t u = 0;
            if (lane == 0) u = atomicAdd(rv.unit_ctr, 1u);
            u = __shfl_sync(FULL, u, 0);
            if (u >= rv.n_units) { exhausted = true; break; }
            nx = u * rv.unit;
            ne = min(nx + rv.unit, rv.n);
            // the set-up data of the unit's nodes will be needed within the next few hundred iterations
            for (uint32_t i = nx + 8 * lane; i < ne; i += 256) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rv.nrec + i));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rv.nrec2 + i));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(plan + i));
            }
          }
          const uint32_t take = min(ne - nx, cnt - taken);
          if (r >= taken && r < taken + take) my = nx + (r - taken);
          nx += take;
          taken += take;
        }
        if (c == S_FREE && my != NOT_FOUND) {
          const NodePlan pl = plan[my];
          if (pl.rec != nullptr) {
            const uint4 p = rv.nrec[my];
            const uint4 q = rv.nrec2[my];
            const uint32_t a = q.w;
            t = my;
            if (LIST) vl = rv.nodes[my];
            d = p.z;
            dc.state = p.x;
            dc.sp = p.y;
            dc.w = q.x;
            end_state = q.y;
            end_sp = q.z;
            recw = pl.rec;
            slot_end = pl.slot_end;
            ns = 0;
            ni = 0;
            extras = a;  // (records with copy blocks: the outdegree of the referenced node, until the blocks end)
            const uint32_t b = p.w >> RT_BITS;
            bb = b | ((p.w & RT_MASK) ? 0x80000000u : 0u);
            k = b;
            prev = 0xFFFFFFFFu;
            copied = 0;
            sgn = 1u;
            c = b ? (uint32_t)S_BLK : s_extras;
          }
        }
      }
    } else if (busy == 0) break;
    // ---------------------------------------------------------------- long residual runs
    // When every busy lane of the warp is deep inside a run of residuals (the hubs of power-law graphs, typically
    // alone in their warp at the end) and no refill is due, 32 symbols are decoded in a loop that is only the
    // symbol decode and the running sum: a record is one serial chain, so what counts for it is the latency of
    // one iteration, and the general step below is four times longer.
    {
      const bool inrun = c == S_RES && extras > 32u;
      if (__all_sync(FULL, c == S_FREE || inrun) && (exhausted || 32u - (uint32_t)__popc(busy) < refill_min)) {
        if (inrun) {
          uint32_t err = 0;
          const uint4 cp = s_cp[S_RES];
          uint32_t* const wp = slot_end - extras;
#pragma unroll 1
          for (int i = 0; i < 32; ++i) {
            prev += 1u + (uint32_t)ans_decode_cp(cp, tab, dc, stream, err);
            wp[i] = prev;
          }
          extras -= 32u;
          if (err) {
            atomicOr(rv.err, ERR_CORRUPT);
            rv.meta[t] = MF_ERR << 29;
            c = S_FREE;
          }
        }
        continue;
      }
    }
    // ---------------------------------------------------------------- one symbol per busy lane
    if (c != S_FREE) {
      uint32_t err = 0;
      const uint64_t x = ans_decode_cp(s_cp[c], tab, dc, stream, err);
      const uint32_t xl = (uint32_t)x;
      const uint32_t v = LIST ? vl : lo32 + t;
      const bool isblk = c == S_BLK, isicnt = c == S_ICNT, isilen = c == S_ILEN, isres = c >= S_RES0;
      const bool nat = c == S_IST0 || c == S_RES0;
      const uint32_t half = (uint32_t)(x >> 1);
      const uint32_t natv = (xl & 1u) ? v - half - 1u : v + half;  // v + nat2int(x)
      const uint32_t lenv = xl + (isilen ? minint : 0u);
      const uint32_t val = nat ? natv : (isicnt || isilen) ? lenv : prev + 1u + xl;
      prev = isilen ? prev + lenv : val;  // (interval length: one past the end of the interval)
      bool bad = err != 0;
      if (isblk) { copied += sgn * val; sgn = 0u - sgn; }  // alternating sum of the cumulative ends = copied elements
      if (isicnt) { ni = xl; k = xl; bad = bad || (uint64_t)xl * minint > extras || (x >> 32) != 0; }
      else if (isblk || isilen) --k;
      if (isilen) bad = bad || lenv > extras || lenv < xl || (x >> 32) != 0;
      if (isblk && k == 0) {  // end of the block run: what is left for intervals and residuals
        if ((bb & 1u) == 0) copied += extras;  // even count: the tail of the referenced list is copied too
        bad = bad || copied > d;
        extras = d - copied;
      } else extras -= isres ? 1u : (isilen ? lenv : 0u);
      // ---------------------------------------------------------------- one word of the record
      if (!bad) {
        if (c == S_RES0) recw[ns++] = extras + 1u;  // the record ends with the number of parked residuals
        // residuals: the last words of the node's slot (they are merged in place by K2: the write position never
        // overtakes the unread ones); without reference and intervals they are the final list
        uint32_t* const base = isres ? slot_end : recw;
        const int32_t idx = isres ? -(int32_t)(extras + 1u) : (int32_t)ns;
        base[idx] = val;
        ns += isres ? 0u : 1u;
      }
      // ---------------------------------------------------------------- next state / end of the record
      const uint32_t tr = (uint32_t)(trans >> (8u * c + (k != 0 ? 4u : 0u))) & 15u;
      const bool finish = (tr & 8u) != 0 && extras == 0;
      if (bad || finish) {
        uint32_t m;
        if (bad || ns > NSYM_MAX || (end_state != 0 && (dc.state != end_state || dc.sp != end_sp))) {
          atomicOr(rv.err, (!bad && ns > NSYM_MAX) ? ERR_LIMIT : ERR_CORRUPT);
          m = MF_ERR << 29;
        } else {
          const uint32_t fl = !isres ? 0u : ((bb >> 31) == 0 && ni == 0) ? (MF_INSLOT | MF_FINAL) : MF_INSLOT;
          m = ns | (fl << 29);
        }
        rv.meta[t] = m;
        c = S_FREE;
      } else c = tr & 7u;
    }
  }
}

// -------------------------------------------------------------------------------------------- K2
// Phase two: copy-block resolution + interval expansion + merge, by reference-chain depth.
//   k_levels   depth[v] = ref ? depth[v-ref]+1 : 0 for every node that still needs work
//   cub sort   stable sort by depth: one contiguous segment per level, node order kept inside a level (the lanes of
//              a warp then work on neighbouring nodes, whose records and referenced lists share sectors)
//   k_resolve  one launch per level; one node per LANE, lanes pull nodes from their block's share of the level: a
//              three-way merge of (copied elements of the finished referenced list, expanded intervals, residuals)
//              into the node's CSR slot.  A large pool of nodes per level keeps every warp of the machine busy; a
//              shared-memory tile per block was measured and lost (one busy warp per level under the barriers).
constexpr uint32_t LCAP = 6;       // levels 0..LCAP-1 have their own segment; deeper nodes share segment LCAP
constexpr uint32_t KEY_SKIP = 15;  // level bucket of nodes that are final after K1
constexpr int RES_TPB = 128;
constexpr uint32_t HUB_CAP = 1u << 16;  // entries of the hub list (more hubs than that take the per-lane path)

__global__ void __launch_bounds__(256) k_levels(RangeView rv, uint8_t* keys, uint32_t* vals, uint32_t* lev_out, uint32_t* hist) {
  __shared__ uint32_t s_hist[16];
  if (threadIdx.x < 16) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lev = 0;
  if (t < rv.n) {
    const uint32_t fl = rv.meta[t] >> 29;
    uint32_t lb = KEY_SKIP;
    if (rv.outdeg[t] != 0 && !(fl & (MF_ERR | MF_FINAL))) {
      uint32_t u = t, r = rv.nrec[t].w & RT_MASK;
      while (r) {  // chain of referenced nodes (K0 made sure it stays inside the decoded nodes)
        u -= r;
        ++lev;
        r = rv.nrec[u].w & RT_MASK;
      }
      lb = min(lev, LCAP);
      lev_out[t] = lev;
      if (rv.outdeg[t] >= rv.hub_min) {  // long list: out of the level segments, into the hub list
        const uint32_t slot = atomicAdd(rv.hub_count, 1u);
        if (slot < HUB_CAP) { rv.hubs[slot] = make_uint2(t, lev); lb = KEY_SKIP; }
      }
    }
    keys[t] = (uint8_t)lb;
    vals[t] = t;
    atomicAdd(&s_hist[lb], 1u);
  }
  for (int o = 16; o; o >>= 1) lev = max(lev, __shfl_xor_sync(FULL, lev, o));
  if ((threadIdx.x & 31) == 0 && lev >= LCAP) atomicMax(rv.maxlevel, lev);
  __syncthreads();
  if (threadIdx.x < 16 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
}

// hist[16] -> seg[17] (exclusive prefix): nodes of level bucket l are order[seg[l] .. seg[l+1])
__global__ void k_segments(const uint32_t* hist, uint32_t* seg) {
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int l = 0; l < 16; ++l) { seg[l] = acc; acc += hist[l]; }
    seg[16] = acc;
  }
}

// One level of phase two.  Lane-per-node state machine: every lane holds one node and emits ONE successor per
// step -- the minimum of the three run heads -- so that all lanes of a warp run the same short merge step
// regardless of how their lists are composed.  Lanes that finish a node wait until setup_batch lanes are free and
// then fetch + set up their next nodes together (the set-up is several dependent loads).  Each block owns a
// contiguous share of the level's segment and hands its nodes out in order.
//   record of node t (K1), contiguous in the record buffer:
//     [cumulative copy-block ends x b][interval count][start,len x ni][residuals]
//   MF_INSLOT: the residuals are not in the record but parked at the tail of the node's own slot; they are consumed
//   before the write position reaches them (written <= copied + interval elements + residuals consumed).
constexpr uint32_t HS = 16;  // header words (copy-block ends, interval count, interval pairs) a lane caches in shared memory

// A long list merged by a whole warp (all lanes call this with the same node).  Same three runs as the per-lane
// merge; per round the run with the smallest head emits as many of its next 32 elements as are below the heads of the
// other two runs, with coalesced loads and stores.  A lane-serial merge of a hub of 10^5 successors is 10^5
// dependent steps on one lane while the rest of the level waits for it; hubs have long runs, so this is ~30x fewer.
__device__ void resolve_hub(const RangeView& rv, uint32_t t, uint32_t minint) {
  const uint32_t lane = threadIdx.x & 31;
  const uint4 nr = rv.nrec[t];
  const uint32_t m = rv.meta[t];
  const uint32_t* recp = rv.recs + rv.roff[t];
  const uint32_t rt = nr.w & RT_MASK, ns = m & NSYM_MAX, fl = m >> 29;
  const uint32_t d = nr.z, b = nr.w >> RT_BITS;
  uint32_t* const out = node_slot(rv, t);
  const uint32_t* ref = nullptr;
  uint32_t dref = 0;
  if (rt) {
    ref = node_slot(rv, t - rt);
    dref = rv.outdeg[t - rt];
  }
  if (!out || (rt && !ref)) {
    if (lane == 0) atomicOr(rv.err, ERR_WORKSPACE);
    return;
  }
  uint32_t ni = 0, H = b;
  if (ns > b && minint) { ni = recp[b]; H = b + 1; }
  if (ns < H || 2 * (uint64_t)ni > (uint64_t)(ns - H)) return;  // inconsistent record (K1 reported it)
  uint32_t ip = H;
  const uint32_t iend = H + 2 * ni;
  uint32_t nres = 0;
  if (fl & MF_INSLOT) {
    nres = ns > iend ? recp[iend] : 0u;
    if (nres > d) return;
  }
  const uint32_t* const res = out + (d - nres);
  uint32_t p = 0, rj = 0, kb = 0, ci = 0, cend = 0, ival = INF, ilim = 0;
  if (ni) { ival = recp[ip]; ilim = ival + recp[ip + 1]; ip += 2; }
  auto next_copy_block = [&]() {
    for (;;) {
      kb += 2;
      if (kb - 1 >= b) { ci = cend = dref; return; }
      ci = min(recp[kb - 1], dref);
      cend = kb < b ? min(recp[kb], dref) : dref;
      if (ci < cend) return;
    }
  };
  if (rt) {
    cend = b ? min(recp[0], dref) : dref;
    if (ci >= cend) next_copy_block();
  }
  while (p < d) {
    const uint32_t cval = ci < cend ? ref[ci] : INF;
    const uint32_t rval = rj < nres ? res[rj] : INF;
    const uint32_t mn = min(cval, min(ival, rval));
    if (mn == INF) {  // (corrupt record: the runs end before the list is full)
      for (uint32_t q = p + lane; q < d; q += 32) out[q] = INF;
      return;
    }
    const bool is_c = mn == cval, is_r = !is_c && mn == rval;
    const uint32_t other = is_c ? min(ival, rval) : is_r ? min(cval, ival) : min(cval, rval);
    uint32_t avail, cand = INF;
    if (is_c) { avail = cend - ci; if (lane < avail) cand = ref[ci + lane]; }
    else if (is_r) { avail = nres - rj; if (lane < avail) cand = res[rj + lane]; }
    else { avail = ilim - ival; if (lane < avail) cand = ival + lane; }
    const uint32_t okm = __ballot_sync(FULL, lane == 0 || (lane < avail && cand < other));
    uint32_t n = okm == FULL ? 32u : (uint32_t)__ffs((int)~okm) - 1u;  // leading lanes that go out
    n = min(n, d - p);
    if (lane < n) out[p + lane] = cand;
    p += n;
    if (is_c) {
      ci += n;
      if (ci == cend) next_copy_block();
    } else if (is_r) rj += n;
    else {
      ival += n;
      if (ival == ilim) {
        if (ip < iend) { ival = recp[ip]; ilim = ival + recp[ip + 1]; ip += 2; } else ival = INF;
      }
    }
  }
}

__global__ void __launch_bounds__(RES_TPB) k_resolve(RangeView rv, const uint32_t* order, const uint32_t* seg, uint32_t lb,
                                                     uint32_t exact_level, const uint32_t* lev, uint32_t minint,
                                                     uint32_t setup_batch) {
  __shared__ uint32_t s_hdr[RES_TPB * (HS + 1)];
  __shared__ __align__(16) uint32_t s_stage[8 * RES_TPB];
  __shared__ uint32_t s_next;
  uint32_t* const hdr = s_hdr + threadIdx.x * (HS + 1);  // odd stride: conflict-free
  // Output staging: a ring of 8 words per lane (word i at stg[i * RES_TPB]: conflict-free).  The successors go out
  // as aligned 16-byte stores instead of one 4-byte store each -- every lane writes to its own list, so each scalar
  // store is a separate L2 request, and those requests were what bound this kernel (lts__t_tag_requests 71 %).
  // Only the partial quads at the two ends of a list are written word by word (quads rather than whole 32-byte
  // sectors: the same number of L2 requests, and the word-by-word code, which runs in nearly every step because
  // some lane of the warp starts or ends a list, handles three words instead of seven).
  uint32_t* const stg = s_stage + threadIdx.x;
  uint32_t A = 0;  // word offset of the slot inside its 16-byte quad: list position p lives at quad position A + p
  {  // the long lists of this level first (they take the longest): one warp each
    const uint32_t nh = min(*rv.hub_count, HUB_CAP);
    const uint32_t want = exact_level ? exact_level : lb;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < nh; h += nw) {
      const uint2 e = rv.hubs[h];
      if (e.y == want) resolve_hub(rv, e.x, minint);
    }
  }
  const uint32_t beg = seg[lb], end = seg[lb + 1];
  const uint32_t len = end - beg;
  const uint32_t share = (len + gridDim.x - 1) / gridDim.x;
  const uint32_t cb = beg + min(len, blockIdx.x * share), ce = beg + min(len, (blockIdx.x + 1) * share);
  if (cb >= ce) return;
  if (threadIdx.x == 0) s_next = cb;
  __syncthreads();
  constexpr int STEPS_PER_VOTE = 2;  // merge steps between two scheduling votes
  enum { S_FETCH, S_MERGE, S_IDLE };
  int st = S_FETCH;
  uint32_t* out = nullptr;          // the node's slot; p = successors written, d = outdegree
  const uint32_t* ref = nullptr;    // the referenced list; [ci, cend) = current copy block
  const uint32_t* hp = nullptr;     // header of the record: this lane's shared-memory copy, or the record itself when too long
  const uint32_t* res = nullptr;    // the residuals: in the record, or parked at the tail of `out` (MF_INSLOT)
  uint32_t p = 0, d = 0, ci = 0, cend = 0, dref = 0, rj = 0, nres = 0;
  uint32_t cval = INF, ival = INF, ilim = 0, rval = INF, b = 0, kb = 0, ip = 0, iend = 0;
  // the current copy block is exhausted: skip block kb+1, then copy block kb+2 (block ends are cumulative)
  auto next_copy_block = [&]() {
    cval = INF;
    for (;;) {
      kb += 2;
      if (kb - 1 >= b) { ci = cend = dref; return; }
      ci = min(hp[kb - 1], dref);
      cend = kb < b ? min(hp[kb], dref) : dref;
      if (ci < cend) { cval = ref[ci]; return; }
    }
  };
  for (;;) {
    const uint32_t fetchers = __ballot_sync(FULL, st == S_FETCH);
    const uint32_t mergers = __ballot_sync(FULL, st == S_MERGE);
    if ((fetchers | mergers) == 0) break;
    if (fetchers && (mergers == 0 || __popc(fetchers) >= setup_batch)) {
      if (st == S_FETCH) {
        uint32_t i = atomicAdd(&s_next, 1u);
        uint32_t t = 0;
        bool have = false;
        while (i < ce) {  // (levels deeper than LCAP share a segment: skip nodes of other levels)
          t = order[i];
          if (!exact_level || lev[t] == exact_level) { have = true; break; }
          i = atomicAdd(&s_next, 1u);
        }
        if (!have) st = S_IDLE;
        else {
          const uint4 nr = rv.nrec[t];
          const uint32_t m = rv.meta[t];
          const uint32_t* recp = rv.recs + rv.roff[t];
          const uint32_t rt = nr.w & RT_MASK, ns = m & NSYM_MAX, fl = m >> 29;
          d = nr.z;
          b = nr.w >> RT_BITS;
          out = node_slot(rv, t);
          A = (uint32_t)(reinterpret_cast<uintptr_t>(out) >> 2) & 3u;
          ref = nullptr;
          dref = 0;
          if (rt) {
            ref = node_slot(rv, t - rt);
            dref = rv.outdeg[t - rt];
          }
          if (!out || (rt && !ref)) { atomicOr(rv.err, ERR_WORKSPACE); d = 0; }
          // layout of the record
          uint32_t ni = 0, H = b;
          if (ns > b && minint) { ni = recp[b]; H = b + 1; }
          if (ns < H || 2 * (uint64_t)ni > (uint64_t)(ns - H)) { ni = 0; b = 0; H = 0; d = 0; }  // inconsistent record
          ip = H;
          H += 2 * ni;
          iend = H;
          if (H <= HS) {
            for (uint32_t w = 0; w < H; w += 4) {  // four independent loads per trip
              uint32_t x[4];
#pragma unroll
              for (uint32_t j = 0; j < 4; ++j) x[j] = w + j < H ? recp[w + j] : 0u;
#pragma unroll
              for (uint32_t j = 0; j < 4; ++j) if (w + j < H) hdr[w + j] = x[j];
            }
            hp = hdr;
          } else hp = recp;
          nres = 0;
          if (fl & MF_INSLOT) {  // parked residuals: the last words of the slot; their number ends the record
            nres = ns > H ? recp[H] : 0u;
            if (nres > d) { nres = 0; d = 0; }
          }
          res = out + (d - nres);
          // heads of the three runs
          p = 0;
          rj = 0;
          rval = (nres && d != 0) ? res[0] : INF;
          ival = INF;
          if (ni) { ival = hp[ip]; ilim = ival + hp[ip + 1]; ip += 2; }
          cval = INF;
          ci = cend = 0;
          kb = 0;
          if (rt && d != 0) {  // (d == 0: nothing to do, or the node / its reference does not fit the output buffer)
            cend = b ? min(hp[0], dref) : dref;
            if (ci < cend) cval = ref[0]; else next_copy_block();  // (only the first block can be empty)
          }
          st = (d != 0) ? S_MERGE : S_FETCH;
        }
      }
    }
#pragma unroll
    for (int step = 0; step < STEPS_PER_VOTE; ++step) {
      if (st == S_MERGE) {
        // The smallest head goes out, and with it up to three more elements of the SAME run that are still below
        // the heads of the other two runs: their loads are independent, so a lane inside a long copy block (or a
        // run of residuals) moves four elements per memory round trip instead of one.
        const uint32_t mn = min(cval, min(ival, rval));
        const bool is_c = mn == cval, is_r = !is_c && mn == rval;
        const uint32_t other = is_c ? min(ival, rval) : is_r ? min(cval, ival) : min(cval, rval);
        uint32_t e1 = INF, e2 = INF, e3 = INF, e4 = INF;  // the elements after the head (INF: the run ends before)
        if (is_c || is_r) {
          const uint32_t* src = is_c ? ref + ci : res + rj;  // (in-slot residuals: unread input, see above)
          const uint32_t avail = is_c ? cend - ci : nres - rj;
          if (avail > 1) e1 = src[1];
          if (avail > 2) e2 = src[2];
          if (avail > 3) e3 = src[3];
          if (avail > 4) e4 = src[4];
        } else {
          const uint32_t avail = ilim - ival;
          if (avail > 1) e1 = ival + 1;
          if (avail > 2) e2 = ival + 2;
          if (avail > 3) e3 = ival + 3;
          if (avail > 4) e4 = ival + 4;
        }
        const bool t1 = e1 < other, t2 = t1 && e2 < other, t3 = t2 && e3 < other;
        const uint32_t m = min(1u + (uint32_t)t1 + (uint32_t)t2 + (uint32_t)t3, d - p);
        const uint32_t q0 = A + p;
        stg[(q0 & 7u) * RES_TPB] = mn;
        if (m > 1) stg[((q0 + 1u) & 7u) * RES_TPB] = e1;
        if (m > 2) stg[((q0 + 2u) & 7u) * RES_TPB] = e2;
        if (m > 3) stg[((q0 + 3u) & 7u) * RES_TPB] = e3;
        const uint32_t nh = m == 1 ? e1 : m == 2 ? e2 : m == 3 ? e3 : e4;  // the run's next head
        p += m;
        {
          const uint32_t q1 = A + p, s0 = q0 & ~3u;
          const bool crossed = q1 >= s0 + 4u;
          if (crossed || p == d) {
            uint32_t qa = s0 < A ? A : s0, qb = q1;  // staged and not yet written: [qa, qb)
            if (crossed && s0 >= A) {  // a whole quad of this list
              const uint32_t* r = stg + (s0 & 4u) * RES_TPB;
              *reinterpret_cast<uint4*>(out + (s0 - A)) = make_uint4(r[0], r[RES_TPB], r[2 * RES_TPB], r[3 * RES_TPB]);
              qa = s0 + 4u;
            } else if (crossed) qb = (p == d) ? q1 : s0 + 4u;  // (the list starts inside this quad)
            if (p != d && !(crossed && s0 < A)) qb = qa;        // the open quad stays staged until it is complete
            // (three words at most inside one quad: no loop for them, the lanes of a warp have different counts)
#pragma unroll
            for (uint32_t j = 0; j < 3; ++j)
              if (qa + j < qb) out[qa + j - A] = stg[((qa + j) & 7u) * RES_TPB];
            for (uint32_t q = qa + 3u; q < qb; ++q) out[q - A] = stg[(q & 7u) * RES_TPB];  // (a short list that starts inside one quad and ends in the next)
          }
        }
        if (is_c) {
          ci += m;
          if (ci == cend) next_copy_block(); else cval = nh;
        } else if (is_r) {
          rj += m;
          rval = nh;
        } else {
          ival += m;
          if (ival == ilim) {
            if (ip < iend) { ival = hp[ip]; ilim = ival + hp[ip + 1]; ip += 2; } else ival = INF;
          }
        }
        if (p == d) st = S_FETCH;
      }
    }
  }
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
  for (uint64_t i = 0; i < n; ++i) out[i] = ans_decode<true>(g.tb, tab, comps[i], dc, g.stream, err);
  end[0] = (uint64_t)dc.sp;
  end[1] = dc.state;
  end[2] = err;
}

__global__ void k_offsets_rebase(const uint64_t* src, uint64_t base, uint64_t* dst, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - base;
}

// scalars the host needs after a decode -> mapped host memory (no DMA copy: see wga_graph::h_pub)
__global__ void k_publish(const uint64_t* tot0, const uint64_t* tot1, const uint32_t* maxlevel, const uint32_t* err,
                          uint64_t* pub) {
  if (threadIdx.x == 0) {
    pub[1] = *tot0;
    pub[2] = *tot1;
    pub[3] = *maxlevel;
    pub[4] = *err;
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
  uint32_t maxlevel;
  uint32_t hist[16];          // nodes per level bucket
  uint32_t seg[17];           // exclusive prefix of hist
  uint32_t hub_count;
};
static_assert(sizeof(Scalars) <= 256, "Scalars must fit the cleared line");

struct WorkspacePlan {
  uint64_t off_outdeg, off_nrec, off_nrec2, off_offs, off_roff, off_meta, off_plan, off_hubs, off_lev, off_keys[2], off_vals[2], off_cub, off_halo, off_recs;
  uint64_t cub_bytes, halo_cap, fixed_bytes;
};

WorkspacePlan plan_workspace(uint64_t n) {
  WorkspacePlan p{};
  uint64_t o = 0;
  o += 256;  // Scalars
  p.off_outdeg = o; o = align_up(o + 4 * (n + 1), 256);
  p.off_nrec = o; o = align_up(o + 16 * n, 256);
  p.off_nrec2 = o; o = align_up(o + 16 * n, 256);
  p.off_offs = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_roff = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_meta = o; o = align_up(o + 4 * n, 256);
  p.off_plan = o; o = align_up(o + sizeof(NodePlan) * n, 256);
  p.off_hubs = o; o = align_up(o + 8ull * HUB_CAP, 256);
  p.off_lev = o; o = align_up(o + 4 * n, 256);
  for (int i = 0; i < 2; ++i) { p.off_keys[i] = o; o = align_up(o + n, 256); }
  for (int i = 0; i < 2; ++i) { p.off_vals[i] = o; o = align_up(o + 4 * n, 256); }
  size_t scan_bytes = 0, sort_bytes = 0;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, (uint64_t*)nullptr, (int64_t)(n + 1));
  cub::DoubleBuffer<uint8_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, (int64_t)n, 0, 4);
  p.cub_bytes = std::max(scan_bytes, sort_bytes);
  p.off_cub = o; o = align_up(o + p.cub_bytes, 256);
  p.halo_cap = 1u << 20;  // successors of halo nodes (u32 each)
  p.off_halo = o; o = align_up(o + 4 * p.halo_cap, 256);
  p.off_recs = align_up(o, 256);
  p.fixed_bytes = p.off_recs;
  return p;
}

uint32_t effective_unit(const Tuning& tn) { return std::max<uint32_t>(1u, tn.unit); }

}  // namespace

uint64_t decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last) {
  uint64_t n = last - first + 4096;  // room for a halo
  WorkspacePlan p = plan_workspace(n);
  double frac = g->prelude.number_of_nodes ? (double)(last - first) / (double)g->prelude.number_of_nodes : 1.0;
  uint64_t arcs_est = (uint64_t)((double)g->prelude.number_of_arcs * frac) + (1u << 20);
  // record buffer: (block count + 1 + outdegree) words per node; the block counts sum to about three per node on web
  // graphs and are bounded by (outdegree of the referenced node + 1)
  if (g->prelude.min_interval_length == 1) arcs_est *= 2;  // (two record words per successor in the worst case)
  uint64_t rows_bytes = 4 * (arcs_est + arcs_est / 4 + 8 * n) + (16ull << 20);
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
  WorkspacePlan p = plan_workspace(n);
  if (ws_bytes < p.off_nrec) throw Error(WGA_E_WORKSPACE, "workspace too small");
  uint8_t* w = (uint8_t*)ws;
  uint32_t* outdeg = (uint32_t*)(w + p.off_outdeg);
  k_heads<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, first, nullptr, (uint32_t)n, outdeg, nullptr, nullptr, g->d_err);
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

}  // namespace

// K0 .. K2 on the nodes described by rv (a contiguous range, or a sorted node list), then one host
// synchronisation that reads back the totals (tot[0] = halo arcs, tot[1] = all arcs), the hard-node count and
// the error word.
static void run_pipeline(wga_graph* g, RangeView& rv, uint8_t* w, const WorkspacePlan& p, Scalars* sc,
                         const Tuning& tn, cudaStream_t st, uint64_t tot[2]) {
  const uint64_t n = rv.n;
  const DeviceInfo& di = device_info(g->device);
  // ---- K0 + scan
  k_heads<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, rv.lo, rv.nodes, (uint32_t)n, rv.outdeg, rv.nrec, rv.nrec2, g->d_err);
  count_launch();
  {
    size_t cb = p.cub_bytes;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(rv.outdeg, U32ToU64());
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, rv.offs, (int64_t)(n + 1), st));
    cb = p.cub_bytes;
    cub::TransformInputIterator<uint64_t, RecWords, const uint4*> itr(rv.nrec, RecWords{g->dev.min_interval});
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, itr, rv.roff, (int64_t)n, st));
    count_launch(4);
  }
  NodePlan* const nplan = (NodePlan*)(w + p.off_plan);
  k_plan<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g->dev, rv, nplan);
  count_launch();
  mark(g, st);  // 1: heads + scans + plan done
  // ---- K1: entropy decode into rows
  {
    K1Tables kt{};
    const uint64_t smem = plan_k1_tables(g, kt, di.smem_optin);
    uint32_t blocks = tn.k1_blocks ? tn.k1_blocks : (uint32_t)di.sms;
    blocks = std::min<uint32_t>(blocks, (rv.n_units + K1_WARPS - 1) / K1_WARPS);
    blocks = std::max<uint32_t>(1u, blocks);
    bool allhot = true;
    for (int c = Blocks; c <= Residual; ++c) allhot = allhot && kt.cp[c].w >= g->packed.nent[c];
    const uint32_t refill = std::max<uint32_t>(1u, std::min<uint32_t>(32u, tn.refill));
    auto launch = [&](auto kern) {
      WGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin - 1024));
      kern<<<blocks, K1_THREADS, smem, st>>>(g->dev, rv, kt, nplan, refill);
    };
    if (rv.nodes) { if (allhot) launch(k_entropy<true, true>); else launch(k_entropy<true, false>); }
    else { if (allhot) launch(k_entropy<false, true>); else launch(k_entropy<false, false>); }
    count_launch();
  }
  mark(g, st);  // 2: entropy decode done
  // ---- K2: levels, stable sort by level, one resolve launch per level
  uint32_t* lev = (uint32_t*)(w + p.off_lev);
  cub::DoubleBuffer<uint8_t> dkeys((uint8_t*)(w + p.off_keys[0]), (uint8_t*)(w + p.off_keys[1]));
  cub::DoubleBuffer<uint32_t> dvals((uint32_t*)(w + p.off_vals[0]), (uint32_t*)(w + p.off_vals[1]));
  const bool have_refs = g->prelude.compression_window != 0 || g->prelude.min_interval_length != 0;
  uint32_t grid = tn.k2_blocks;
  const uint32_t k2_batch = std::max<uint32_t>(1u, std::min<uint32_t>(32u, tn.k2_batch));
  if (have_refs) {
    k_levels<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rv, dkeys.Current(), dvals.Current(), lev, sc->hist);
    count_launch();
    size_t cb = p.cub_bytes;
    WGA_CUDA(cub::DeviceRadixSort::SortPairs(w + p.off_cub, cb, dkeys, dvals, (int64_t)n, 0, 4, st));
    count_launch(3);
    k_segments<<<1, 32, 0, st>>>(sc->hist, sc->seg);
    count_launch();
    mark(g, st);  // 3: levels + sort done
    if (!grid) {
      int per_sm = 8;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resolve, RES_TPB, 0);
      per_sm = per_sm > 10 ? 10 : (per_sm > 0 ? per_sm : 1);  // swept: more resident warps thrash L1
      grid = (uint32_t)(di.sms * per_sm);
    }
    const uint32_t nlev = g->prelude.compression_window ? LCAP : 1;  // without references everything is level 0
    for (uint32_t l = 0; l < nlev; ++l) {
      k_resolve<<<grid, RES_TPB, 0, st>>>(rv, dvals.Current(), sc->seg, l, 0, lev, g->dev.min_interval, k2_batch);
      count_launch();
    }
  } else {
    mark(g, st);
  }
  mark(g, st);  // 4: resolve done
  // ---- totals, deepest level, error word
  k_publish<<<1, 32, 0, st>>>(rv.offs + rv.h, rv.offs + n, &sc->maxlevel, g->d_err, g->d_pub);
  count_launch();
  WGA_CUDA(cudaStreamSynchronize(st));
  WGA_CUDA(cudaGetLastError());
  tot[0] = g->h_pub[1];
  tot[1] = g->h_pub[2];
  const uint32_t maxlevel = (uint32_t)g->h_pub[3];
  const uint32_t herr = (uint32_t)g->h_pub[4];
  if (tot[0] > rv.halo_cap) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "the " + std::to_string(rv.h) + " predecessor nodes this range references (halo) have " +
                                     std::to_string(tot[0]) + " successors, the workspace holds " + std::to_string(rv.halo_cap) +
                                     ": decode a range that starts " + std::to_string(rv.h) + " nodes earlier");
  }
  if (tot[1] - tot[0] > rv.succ_cap) {
    g->last_need_succ = tot[1] - tot[0];
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(tot[1] - tot[0]) + " elements");
  }
  check_device_error(g, herr, st);
  // ---- reference chains deeper than LCAP (e.g. graphs compressed with an unbounded max_ref_count): one
  //      launch per extra level over the shared deep segment
  if (have_refs && maxlevel >= LCAP) {
    for (uint32_t l = LCAP; l <= maxlevel; ++l) {
      k_resolve<<<grid, RES_TPB, 0, st>>>(rv, dvals.Current(), sc->seg, LCAP, l, lev, g->dev.min_interval, k2_batch);
      count_launch();
    }
    check_device_error(g, read_device_error(g, st), st);
  }
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

static void bind_views(RangeView& rv, uint8_t* w, const WorkspacePlan& p, Scalars* sc, uint64_t ws_bytes, uint32_t unit,
                       uint32_t hub_min) {
  rv.hub_min = hub_min ? hub_min : 1u;
  rv.outdeg = (uint32_t*)(w + p.off_outdeg);
  rv.nrec = (uint4*)(w + p.off_nrec);
  rv.nrec2 = (uint4*)(w + p.off_nrec2);
  rv.meta = (uint32_t*)(w + p.off_meta);
  rv.maxlevel = &sc->maxlevel;
  rv.hubs = (uint2*)(w + p.off_hubs);
  rv.hub_count = &sc->hub_count;
  rv.roff = (uint64_t*)(w + p.off_roff);
  rv.recs = (uint32_t*)(w + p.off_recs);
  rv.recs_cap = (ws_bytes - p.off_recs) / 4;
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
  const Tuning tn = tuning_snapshot();
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
  g->last_halo_nodes = first - lo;
  const uint64_t n = last - lo;
  const uint32_t unit = effective_unit(tn);
  WorkspacePlan p = plan_workspace(n);
  if (ws_bytes < p.fixed_bytes + 4096)
    throw Error(WGA_E_WORKSPACE, "workspace too small: the range needs " + std::to_string(p.fixed_bytes + 4096) +
                                     " bytes before the record buffer (it references " + std::to_string(first - lo) +
                                     " predecessor nodes; wga_decode_workspace_size allows for 4096)");
  RangeView rv{};
  rv.lo = lo; rv.n = (uint32_t)n; rv.h = (uint32_t)(first - lo);
  bind_views(rv, w, p, sc, ws_bytes, unit, tn.hub_min);
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
  WorkspacePlan p = plan_workspace(b.cap_nodes);
  b.inner_bytes = p.fixed_bytes + 4 * (b.cap_arcs + b.cap_arcs / 4 + 8 * b.cap_nodes) + (16ull << 20);
  b.off_inner = align_up(o, 16384); o = b.off_inner + b.inner_bytes;
  b.total = o;
  return b;
}
}  // namespace

uint64_t successors_workspace_size(const wga_graph* g, uint64_t n_queries, uint64_t max_total_arcs) {
  return plan_batch(n_queries, max_total_arcs, effective_unit(tuning_snapshot())).total;
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
  const Tuning tn = tuning_snapshot();
  const uint32_t unit = effective_unit(tn);
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
    k_heads<<<(unsigned)((nq + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, 0, qid, (uint32_t)nq, deg, nullptr, nullptr, g->d_err);
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
  WorkspacePlan p = plan_workspace(nU);
  if (p.fixed_bytes + 4096 > b.inner_bytes) throw Error(WGA_E_WORKSPACE, "workspace too small");
  Scalars* sc = (Scalars*)iw;
  RangeView rv{};
  rv.lo = 0; rv.n = nU; rv.h = 0; rv.nodes = U;
  bind_views(rv, iw, p, sc, b.inner_bytes, unit, tn.hub_min);
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
