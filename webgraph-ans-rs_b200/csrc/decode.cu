// =============================================================================
//  decode.cu -- ANS decode of BvGraph components into CSR successor lists (sm_100a)
// =============================================================================
//  Replaces, for whole node ranges at once, what the reference does one symbol at a
//  time on one core:
//    webgraph BvGraphSeq::iter() / BvGraph::successors(v)  (external, un-vendored)
//      -> ANSBVGraphDecoderFactory::new_decoder(v)   src/bvgraph/factories/bvgraph_decoder_factory.rs:46-58
//      -> ANSDecoder::decode(component)              src/ans/decoder.rs:58-100
//  Pipeline (all launches on the caller's stream):
//    K0  k_outdegree   one lane per node, every lane at the same symbol (the cheapest way to decode): the
//                      fixed-shape head of every record from (states[N-1-v], pointers[N-1-v]) -- outdegree,
//                      reference offset, block count -- and the decoder state after it for K1
//        cub scan      outdegrees -> CSR offsets
//    K1  k_entropy     phase one: entropy decode of the rest of every record.  One node per LANE: lanes pull
//                      nodes from a per-block counter and run a per-symbol state machine (which component
//                      next / how many left); a warp vote per iteration is the reconvergence point, so the
//                      symbol decode runs with all busy lanes although records differ.  Decoded values are
//                      PARKED inside the node's own final CSR slot: residuals (already prefix-summed) at the
//                      tail, copy-block lengths (u16) and interval (start,len) pairs at the head.  Nodes that
//                      are pure residual lists are final after K1.
//    K2  k_levels      phase two, by reference-chain depth: depth of every node that still needs work
//        cub sort      stable sort by level -> one segment per level, node order kept inside a level
//        k_resolve_big0  long reference-free records with intervals: one block per node
//        k_resolve     per level, one node per lane: three-way merge (copied elements of the finished
//                      referenced list, expanded intervals, residuals) in place into the node's CSR slot
//    Random access (wga_successors_batch) runs the same kernels on the sorted reference closure of the
//    query nodes (node-list mode) and gathers the query lists.
// =============================================================================
#include <cub/cub.cuh>

#include "graph.hpp"

namespace wga {

std::atomic<uint64_t> g_kernel_launches{0};

// run-time tuning (tests change these to exercise span boundaries, grid striding and the overflow paths)
struct Tuning {
  uint32_t k1_span = 2048;    // nodes per K1 block
  uint32_t k1_tpb = 128;      // threads per K1 block
  uint32_t k2_blocks = 0;     // K2 grid; 0 = one full wave (SM count x resident blocks per SM): every block gets an
                              // equal chunk of the level, so a partial second wave would double the time
  uint32_t force_ovf = 0;     // K1: put every header in the overflow arena
  uint32_t e2e_chunk = 1u << 19;  // nodes per chunk of the pipelined host entry point (swept: 31.6 ms at 2^19)
  uint32_t sort_degree = 0;   // K2: 1 = sort key includes the degree bucket; 0 = level only, node order kept
                              // (measured: locality of neighbouring nodes beats equal loop lengths, 2.9 vs 6.6 ms)
};
static Tuning g_tuning;

int set_tuning(const char* key, uint64_t value) {
  std::string k(key ? key : "");
  if (k == "k1_span") g_tuning.k1_span = (uint32_t)value;
  else if (k == "k1_tpb") g_tuning.k1_tpb = (uint32_t)value;
  else if (k == "k2_blocks") g_tuning.k2_blocks = (uint32_t)value;
  else if (k == "force_ovf") g_tuning.force_ovf = (uint32_t)value;
  else if (k == "sort_degree") g_tuning.sort_degree = (uint32_t)value;
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

struct RangeView {
  uint64_t lo;        // first decoded node (halo start)
  uint64_t first;     // first node the caller asked for
  uint32_t n;         // nodes decoded: last - lo
  const uint32_t* nodes;  // nullptr: node t is lo + t; else a sorted, duplicate-free list of node ids (random access)
  uint32_t h;         // halo nodes: first - lo
  uint32_t* outdeg;   // n+1
  uint4* phase1;      // n : from K0: decoder (state, stream index) after the record's head, reference offset, block count
  uint64_t* offs;     // n+1, relative to lo
  uint64_t* meta;     // n : per-node record of K1 (see M_*)
  uint32_t* arena;    // overflow headers (K1) and pass-2 temporaries
  uint64_t arena_cap;
  unsigned long long* cursor;
  uint32_t* maxlevel; // deepest reference chain seen by k_levels (only tracked from LCAP up)
  uint32_t* halo_succ;  // successors of halo nodes
  uint64_t halo_cap;
  uint32_t* succ;       // caller's array: successors of nodes >= first
  uint64_t succ_cap;
  uint32_t* err;
  unsigned long long* stats;  // optional debug counters (16 x u64), nullptr when disabled
};

// meta word written by K1:
//   bits 0-15 reference offset r | bit 16 header in the overflow arena | bit 17 node is final after K1
//   in-slot header: bits 19-33 block count b | 34-47 interval count | 48-63 residual count
//   overflow header: bits 19-63 arena offset of {b, ni, nres, pairs offset, blocks...}
constexpr uint64_t M_OVF = 1ull << 16, M_DIRECT = 1ull << 17;
constexpr uint32_t MAX_B = 1u << 15, MAX_NI = 1u << 14, MAX_NRES = 1u << 16;
constexpr uint32_t HS_WORDS = 16;  // in-slot headers are at most this many words (k_resolve caches them per lane)

constexpr uint32_t NOT_FOUND = 0xFFFFFFFFu;
// Index of the node referenced by node t with reference offset r (r != 0).  In a sorted duplicate-free
// list the node (id - r) sits at most r positions before t.
__device__ __forceinline__ uint32_t ref_index(const RangeView& rv, uint32_t t, uint32_t r) {
  if (!rv.nodes) return r <= t ? t - r : NOT_FOUND;
  const uint32_t id = rv.nodes[t];
  if (r > id) return NOT_FOUND;
  const uint32_t target = id - r;
  uint32_t lo = t >= r ? t - r : 0u, hi = t;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (rv.nodes[mid] < target) lo = mid + 1; else hi = mid;
  }
  return (lo < t && rv.nodes[lo] == target) ? lo : NOT_FOUND;
}

__device__ __forceinline__ uint32_t* node_slot(const RangeView& rv, uint32_t t) {
  uint64_t o = rv.offs[t];
  if (t < rv.h) return rv.halo_succ + o;
  return rv.succ + (o - rv.offs[rv.h]);
}

// Block -> node span [A,B) (indices relative to rv.lo).  Spans never straddle the halo boundary rv.h, so
// that a span's successors are one contiguous piece of either halo_succ or succ.
__host__ __device__ inline uint32_t span_count(uint32_t n, uint32_t h, uint32_t span) {
  return (h + span - 1) / span + (n - h + span - 1) / span;
}
__device__ __forceinline__ void span_range(const RangeView& rv, uint32_t span, uint32_t blk, uint32_t& A, uint32_t& B) {
  const uint32_t nbh = (rv.h + span - 1) / span;
  if (blk < nbh) {
    A = blk * span;
    B = min(A + span, rv.h);
  } else {
    A = rv.h + (blk - nbh) * span;
    B = (rv.n - A > span) ? A + span : rv.n;
  }
}
// true (and error set) when the span's successors would not fit the destination
__device__ __forceinline__ bool span_overflows(const RangeView& rv, uint32_t A, uint32_t B) {
  if (A < rv.h) return rv.offs[B] > rv.halo_cap;
  return rv.offs[B] - rv.offs[rv.h] > rv.succ_cap;
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
// One lane per node, every lane at the same symbol: the cheapest way to decode (about 110 G symbols/s on a B200,
// against 45 G in the general state machine of K1).  So K0 decodes not only the outdegree but the whole
// fixed-shape head of a record -- outdegree, reference offset, block count -- and hands K1 the decoder state
// after it (phase1).  The block count is validated by K1, which knows the outdegree of the referenced node.
__global__ void __launch_bounds__(TPB) k_outdegree(DevGraph g, uint64_t lo, const uint32_t* nodes, uint32_t n,
                                                   uint32_t* outdeg, uint4* phase1, uint32_t* err_out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n) return;
  if (t == n) { outdeg[n] = 0; return; }
  uint64_t v = nodes ? (uint64_t)nodes[t] : lo + t;
  uint32_t state, err = 0;
  int64_t ptr;
  load_phase(g, v, state, ptr, err);
  uint64_t d = ans_decode(g.tb, g.tb.lut, g.tb.ent, Outdegree, state, ptr, g.stream, err);
  if (d > 0xFFFFFFFFull) err |= ERR_SYMBOL_WIDTH;
  outdeg[t] = (uint32_t)d;
  if (phase1) {
    uint32_t r = 0, b = 0;
    if (d != 0 && g.window != 0 && !err) {
      const uint64_t x = ans_decode(g.tb, g.tb.lut, g.tb.ent, ReferenceOffset, state, ptr, g.stream, err);
      if (x > g.window) err |= ERR_CORRUPT;
      r = (uint32_t)x;
      if (r != 0 && !err) {
        const uint64_t y = ans_decode(g.tb, g.tb.lut, g.tb.ent, BlockCount, state, ptr, g.stream, err);
        if (y > 0xFFFFFFFFull) err |= ERR_CORRUPT;
        b = (uint32_t)y;
      }
    }
    phase1[t] = make_uint4(state, (uint32_t)ptr, r, b);
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
// Pseudo components of the per-lane state machine (0..8 are the BVGraphComponent values, mod.rs:46-61).
enum : uint32_t { C_AFTER_BLOCKS = 9, C_FINISH = 10, C_FETCH = 11, C_IDLE = 12 };

// Moves the header of a node (kb block lengths, kp interval pairs already parked in the slot) to a
// record in the overflow arena.  Returns false when the arena is full.
__device__ __noinline__ bool header_to_arena_impl(uint32_t* arena, uint64_t arena_cap, unsigned long long* cursor,
                                                  const uint32_t* slot, uint32_t b, uint32_t kb, uint32_t ni,
                                                  uint32_t kp, uint32_t* ao_out) {
  const unsigned long long need = 4ull + b + 2ull * ni;
  const unsigned long long o = atomicAdd(cursor, need);
  if (o + need > arena_cap || o + need >= 0xFFFFFFFFull) return false;
  const uint32_t ao = (uint32_t)o;
  const uint32_t apo = ao + 4 + b;
  *ao_out = ao;
  uint32_t* rec = arena + ao;
  rec[0] = b;
  rec[1] = ni;
  rec[2] = 0;
  rec[3] = apo;
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(slot);
  for (uint32_t i = 0; i < kb; ++i) rec[4 + i] = s16[i];
  const uint32_t hb = (b + 1) >> 1;
  for (uint32_t i = 0; i < 2 * kp; ++i) arena[apo + i] = slot[hb + i];
  return true;
}
__device__ __forceinline__ bool header_to_arena(const RangeView& rv, const uint32_t* slot, uint32_t b, uint32_t kb,
                                                uint32_t ni, uint32_t kp, uint32_t& ao, uint32_t& apo) {
  uint32_t a = 0;
  const bool ok = header_to_arena_impl(rv.arena, rv.arena_cap, rv.cursor, slot, b, kb, ni, kp, &a);
  ao = a;
  apo = a + 4 + b;
  return ok;
}

constexpr uint32_t SOLO_RUN = 2048;  // residual runs at least this long are finished in a tight loop of their own

// node + nat2int(x) in 32-bit arithmetic (ids are < 2^32, so a valid x is < 2^33); false on leaving [0, 2^32-2]
__device__ __forceinline__ bool add_nat(uint32_t v, uint64_t x, uint32_t& out) {
  const uint32_t half = (uint32_t)(x >> 1);
  const bool neg = (x & 1) != 0;
  out = neg ? v - half - 1u : v + half;
  return (x >> 33) == 0 && (neg ? half < v : (out >= v && out != 0xFFFFFFFFu));
}

// LIST: node t is rv.nodes[t] (random access) instead of rv.lo + t.
template <bool LIST>
__global__ void __launch_bounds__(128) k_entropy(DevGraph g, RangeView rv, uint32_t span, uint32_t force_ovf) {
  __shared__ uint32_t s_next;
  __shared__ uint4 s_cp[WGA_COMPONENTS];
  uint32_t A, Bn;
  span_range(rv, span, blockIdx.x, A, Bn);
  if (span_overflows(rv, A, Bn)) {
    if (threadIdx.x == 0) atomicOr(rv.err, ERR_WORKSPACE);
    return;
  }
  if (threadIdx.x == 0) s_next = A;
  if (threadIdx.x < WGA_COMPONENTS) s_cp[threadIdx.x] = comp_params(g.tb, threadIdx.x);
  __syncthreads();
  const uint16_t* lut = g.tb.lut;
  const uint2* ent = g.tb.ent;
  const uint32_t c_extras = g.min_interval ? (uint32_t)IntervalCount : (uint32_t)FirstResidual;
  const uint32_t minint = g.min_interval;
  const uint32_t window = g.window;
  // slot of node t = base + offs[t] (spans never straddle the halo boundary)
  uint32_t* const slot_base = A < rv.h ? rv.halo_succ : rv.succ - rv.offs[rv.h];
  // phases of node v: states[top - v], ptrs[top - v] (file order is reversed, bvgraph_decoder_factory.rs:49-50)
  const uint32_t* const states_top = g.states + g.top;
  const uint64_t* const ptrs_top = g.ptrs + g.top;
  const uint32_t lo32 = (uint32_t)rv.lo;

  // per-lane record state
  uint32_t c = C_FETCH, t = 0, state = 0, sp = 0, v = 0, prev = 0, d = 0, r = 0, dref = 0, b = 0, k = 0, copied = 0,
           pos = 0, extras = 0, ni = 0, hb = 0, nres = 0, ao = 0, apo = 0;
  uint32_t* slot = nullptr;
  uint32_t* wp = nullptr;
  bool ovf = false, direct = false;

  // Every lane stays in the loop until the whole warp has run out of nodes: the vote at the top is the
  // per-iteration reconvergence point, so that the symbol decode below runs with all busy lanes together.
  // Each case computes an error flag instead of leaving early, which keeps the cases short and single-exit.
  for (;;) {
    uint32_t err = 0;
    if (c == C_FETCH) {
      t = atomicAdd(&s_next, 1u);
      if (t >= Bn) c = C_IDLE;
      else {
        v = LIST ? rv.nodes[t] : lo32 + t;
        const uint4 ph = rv.phase1[t];  // K0 left the decoder after the head: outdegree, reference offset, block count
        state = ph.x;
        sp = ph.y;  // the resident span has < 2^32 words (checked at upload)
        r = ph.z;
        b = ph.w;
        d = rv.outdeg[t];
        extras = d;
        ni = copied = nres = pos = k = 0;
        hb = (b + 1) >> 1;
        ovf = false;
        direct = d == 0;
        if (d == 0) c = C_FINISH;
        else {
          slot = slot_base + rv.offs[t];
          if (r == 0) c = c_extras;
          else {
            const uint32_t ri = LIST ? ref_index(rv, t, r) : (r <= t ? t - r : NOT_FOUND);
            if (ri == NOT_FOUND) err |= ERR_RANGE;  // the referenced node is not part of this decode
            else {
              dref = rv.outdeg[ri];
              if (b > dref && b - dref > 1u) err |= ERR_CORRUPT;  // at most dref + 1 blocks
              else if (b == 0) { copied = dref; c = C_AFTER_BLOCKS; }
              else {
                if (hb > d || hb > HS_WORDS || b >= MAX_B || dref > 0xFFFFu || force_ovf) {
                  if (header_to_arena(rv, slot, b, 0, 0, 0, ao, apo)) ovf = true;
                  else err |= ERR_WORKSPACE;
                }
                c = Blocks;
              }
            }
          }
        }
      }
    }
    if (__all_sync(FULL, c == C_IDLE)) break;
    if (c <= Residual) {
      const uint64_t x = ans_decode_cp(s_cp[c], lut, ent, state, sp, g.stream, err);
      const uint32_t xl = (uint32_t)x;
      const bool wide = (x >> 32) != 0;  // only nat2int arguments (first residual / interval start) may need 33 bits
      if (c >= FirstResidual) {
        // ---- residuals: value = node + nat2int(x) | previous + 1 + x   (most frequent symbols)
        uint32_t val;
        bool ok;
        if (c == FirstResidual) {
          nres = extras;
          direct = (r == 0 && ni == 0);
          if (!direct && !ovf && (nres >= MAX_NRES || hb + 2ull * ni > (uint64_t)(d - nres))) {
            if (header_to_arena(rv, slot, b, b, ni, ni, ao, apo)) ovf = true;
            else err |= ERR_WORKSPACE;
          }
          wp = slot + (d - nres);
          ok = add_nat(v, x, val);
        } else {
          val = prev + 1u + xl;
          ok = !wide && val > prev && val != 0xFFFFFFFFu;
        }
        if (!ok) err |= ERR_SYMBOL_WIDTH;
        if (!err) {
          prev = val;
          *wp++ = val;
          c = --extras ? (uint32_t)Residual : (uint32_t)C_FINISH;
        }
        // A long residual run is one serial chain and ends up as the last thing the kernel waits for (hubs
        // of social graphs: 10^5 gaps): finish it in a loop of its own, whose body is only the symbol decode
        // and the prefix sum, instead of one pass through the whole state machine per gap.
        if (!err && c == Residual && extras >= SOLO_RUN) {
          const uint4 cpr = s_cp[Residual];
          while (extras) {
            uint32_t e2 = 0;
            const uint64_t y = ans_decode_cp(cpr, lut, ent, state, sp, g.stream, e2);
            const uint32_t nv = prev + 1u + (uint32_t)y;
            if (e2 || (y >> 32) || nv <= prev || nv == 0xFFFFFFFFu) { err |= e2 ? e2 : ERR_SYMBOL_WIDTH; break; }
            prev = nv;
            *wp++ = nv;
            --extras;
          }
          if (!err) c = C_FINISH;
        }
      } else if (c == Blocks) {
        const uint32_t len = xl + (k != 0);
        if (wide || len > dref - pos || len < xl) err |= ERR_CORRUPT;
        if (!err) {
          if (ovf) rv.arena[ao + 4 + k] = len;
          else reinterpret_cast<uint16_t*>(slot)[k] = (uint16_t)len;
          if ((k & 1) == 0) copied += len;
          pos += len;
          if (++k == b) {
            if ((b & 1) == 0) copied += dref - pos;
            c = C_AFTER_BLOCKS;
          }
        }
      } else if (c >= IntervalStart) {
        if (c == IntervalStart) {
          uint32_t val;
          bool ok;
          if (k == 0) ok = add_nat(v, x, val);
          else { val = prev + 1u + xl; ok = !wide && val > prev && val != 0xFFFFFFFFu; }  // prev: end of the last one
          if (!ok) err |= ERR_SYMBOL_WIDTH;
          if (!err) {
            prev = val;  // start of this interval
            if (ovf) rv.arena[apo + 2 * k] = val;
            else slot[hb + 2 * k] = val;
            c = IntervalLen;
          }
        } else {
          const uint32_t len = xl + minint;
          if (wide || len < xl || len > extras || len == 0) err |= ERR_CORRUPT;
          const uint32_t end = prev + len;  // one past the end of this interval
          if (end < prev) err |= ERR_SYMBOL_WIDTH;
          if (!err) {
            prev = end;
            if (ovf) rv.arena[apo + 2 * k + 1] = len;
            else slot[hb + 2 * k + 1] = len;
            extras -= len;
            if (++k == ni) c = extras ? (uint32_t)FirstResidual : (uint32_t)C_FINISH;
            else c = IntervalStart;
          }
        }
      } else {  // IntervalCount
        if (wide || xl > extras) err |= ERR_CORRUPT;
        if (!err) {
          ni = xl;
          k = 0;
          if (ni == 0) c = FirstResidual;
          else {
            if (ovf) {  // header already in the arena: the pairs get their own piece
              const unsigned long long o = atomicAdd(rv.cursor, 2ull * ni);
              if (o + 2ull * ni > rv.arena_cap || o + 2ull * ni >= 0xFFFFFFFFull) err |= ERR_WORKSPACE;
              else { apo = (uint32_t)o; rv.arena[ao + 3] = apo; }
            } else if (ni >= MAX_NI || hb + 2ull * ni > d || hb + 2ull * ni > HS_WORDS || force_ovf) {
              if (header_to_arena(rv, slot, b, b, ni, 0, ao, apo)) ovf = true;
              else err |= ERR_WORKSPACE;
            }
            c = IntervalStart;
          }
        }
      }
    }
    if (c == C_AFTER_BLOCKS && !err) {
      if (copied > d) err |= ERR_CORRUPT;
      else {
        extras = d - copied;
        c = extras ? c_extras : (uint32_t)C_FINISH;
      }
    }
    if (err) {  // the record is inconsistent: leave the node out of phase two and report
      atomicOr(rv.err, err);
      rv.meta[t] = M_DIRECT;
      c = C_FETCH;
    } else if (c == C_FINISH) {
      uint64_t m;
      if (direct) m = M_DIRECT;
      else if (ovf) {
        rv.arena[ao + 1] = ni;
        rv.arena[ao + 2] = nres;
        m = (uint64_t)r | M_OVF | ((uint64_t)ao << 19);
      } else {
        m = (uint64_t)r | ((uint64_t)b << 19) | ((uint64_t)ni << 34) | ((uint64_t)nres << 48);
      }
      rv.meta[t] = m;
      c = C_FETCH;
    }
  }
}

// -------------------------------------------------------------------------------------------- K2
// Phase two: copy-block resolution + interval expansion + merge, by reference-chain depth.
//   k_levels   depth[v] = ref ? depth[v-ref]+1 : 0 for every node that still needs work, and a 12-bit sort
//              key (level bucket, degree bucket descending)
//   cub sort   nodes ordered by key: one contiguous segment per level, inside it nodes of similar degree
//              next to each other, so that the 32 lanes of a warp run merge loops of similar length
//   k_resolve  one launch per level; one node per lane: a tight three-way merge of (copied elements of the
//              finished referenced list, expanded intervals, residuals) written in place into the node's
//              CSR slot.  The parked residuals sit at the tail of the slot and are consumed before the write
//              pointer reaches them; the parked header is first copied to shared memory.
constexpr uint32_t HS = 16;       // in-slot header words (u16 block lengths + interval pairs) a lane caches
constexpr uint32_t LCAP = 6;      // levels 0..LCAP-1 have their own segment; deeper nodes share segment LCAP
constexpr uint32_t KEY_SKIP = 15; // level bucket of nodes that are final after K1
constexpr uint32_t KEY_BIG0 = 13; // level bucket of long reference-free records with intervals (k_resolve_big0)
constexpr uint32_t BIG0_DEGREE = 4096;   // outdegree from which a level-0 record takes the cooperative path
constexpr uint32_t BIG0_MAX_NI = 3072;   // intervals the cooperative path keeps in shared memory (3 x 12 KB static)
constexpr int RES_TPB = 128;

__device__ __forceinline__ uint32_t degree_bucket(uint32_t d) {  // monotone, 0..227
  if (d < 128) return d;
  const uint32_t lg = 31u - (uint32_t)__clz((int)d);
  return 128u + (lg - 7u) * 4u + ((d >> (lg - 2u)) & 3u);
}

__global__ void __launch_bounds__(256) k_levels(RangeView rv, uint16_t* keys, uint32_t* vals, uint32_t* lev_out,
                                                uint32_t* hist, uint32_t sort_degree) {
  __shared__ uint32_t s_hist[16];
  if (threadIdx.x < 16) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lev = 0;
  if (t < rv.n) {
    const uint64_t m = rv.meta[t];
    uint32_t lb = KEY_SKIP;
    if (!(m & M_DIRECT)) {
      uint32_t u = t, r = (uint32_t)(m & 0xFFFFu);
      while (r) {  // chain of referenced nodes (a node that is final after K1 has no reference)
        u = ref_index(rv, u, r);
        ++lev;
        r = (uint32_t)(rv.meta[u] & 0xFFFFu);
      }
      lb = min(lev, LCAP);
      lev_out[t] = lev;
      if (lev == 0 && rv.outdeg[t] >= BIG0_DEGREE) {  // one lane would merge this list element by element
        const uint32_t ni = (m & M_OVF) ? rv.arena[(uint32_t)(m >> 19) + 1] : (uint32_t)(m >> 34) & (MAX_NI - 1);
        if (ni <= BIG0_MAX_NI) lb = KEY_BIG0;
      }
    }
    keys[t] = (uint16_t)((lb << 8) | (sort_degree ? 255u - degree_bucket(rv.outdeg[t]) : 0u));
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

// One level of phase two.  Lane-per-node state machine (same shape as K1): every lane holds one node and
// emits ONE successor per iteration -- the minimum of the three stream heads (copied element, interval
// element, residual) -- so that all lanes of a warp run the same short merge step regardless of how their
// lists are composed.  Lanes that finish a node wait until SETUP_BATCH lanes are free and then fetch + set up
// their next nodes together (the set-up is several dependent HBM loads and ~100 instructions).
// Each block owns a contiguous chunk of the level's segment and hands its nodes out in order, so that the
// lanes of a warp work on neighbouring nodes (their records, offsets and referenced lists share sectors).
__global__ void __launch_bounds__(RES_TPB) k_resolve(RangeView rv, const uint32_t* order, const uint32_t* seg,
                                                     uint32_t lb, uint32_t exact_level, const uint32_t* lev) {
  __shared__ uint32_t s_hdr[RES_TPB * (HS + 1)];
  __shared__ uint32_t s_next;
  uint32_t* const hdr = s_hdr + threadIdx.x * (HS + 1);  // odd stride: conflict-free
  const uint32_t beg = seg[lb], end = seg[lb + 1];
  const uint32_t len = end - beg;
  const uint32_t chunk = (len + gridDim.x - 1) / gridDim.x;
  const uint32_t cb = beg + min(len, blockIdx.x * chunk), ce = beg + min(len, (blockIdx.x + 1) * chunk);
  if (cb >= ce) return;
  if (threadIdx.x == 0) s_next = cb;
  __syncthreads();
  constexpr uint32_t SETUP_BATCH = 8;
  constexpr int STEPS_PER_VOTE = 2;  // merge steps between two scheduling votes
  enum { S_FETCH, S_MERGE, S_IDLE };
  int st = S_FETCH;
  // streams of the current node: 32-bit indices against three base pointers
  uint32_t* out = nullptr;          // the node's slot; p = successors written, d = outdegree
  const uint32_t* ref = nullptr;    // the referenced list; [ci, cend) = current copy block
  const uint32_t* res = nullptr;    // the parked residuals (tail of the slot); rj = next one, nres = count
  const uint32_t* blk32 = nullptr;  // block lengths in the overflow arena (else u16 in hdr)
  const uint32_t* pp = nullptr;     // interval pairs (hdr or arena)
  uint32_t p = 0, d = 0, ci = 0, cend = 0, dref = 0, rj = 0, nres = 0;
  uint32_t cval = INF, ival = INF, iend = 0, rval = INF, b = 0, bk = 0, ni = 0, ik = 0;
  uint32_t bw0 = 0, bw1 = 0;        // first four in-slot block lengths (u16 each), kept in registers
  auto block_len = [&](uint32_t k) -> uint32_t {
    if (blk32) return blk32[k];
    if (k < 4) return ((k < 2 ? bw0 : bw1) >> ((k & 1u) * 16u)) & 0xFFFFu;
    return reinterpret_cast<const uint16_t*>(hdr)[k];
  };
  // the current copy block is exhausted: skip block bk, then copy block bk+1 (or the implicit tail)
  auto next_copy_block = [&]() {
    cval = INF;
    if (bk < b) {
      ci += block_len(bk);
      ++bk;
      if (bk < b) { cend = ci + block_len(bk); ++bk; } else cend = dref;
      if (ci < cend) cval = ref[ci];
    }
  };
  for (;;) {
    const uint32_t fetchers = __ballot_sync(FULL, st == S_FETCH);
    const uint32_t mergers = __ballot_sync(FULL, st == S_MERGE);
    if ((fetchers | mergers) == 0) break;
    if (fetchers && (mergers == 0 || __popc(fetchers) >= SETUP_BATCH)) {
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
          const uint64_t m = rv.meta[t];
          const uint32_t r = (uint32_t)(m & 0xFFFFu);
          out = node_slot(rv, t);
          d = (uint32_t)(rv.offs[t + 1] - rv.offs[t]);
          if (m & M_OVF) {
            const uint32_t* rec = rv.arena + (uint32_t)(m >> 19);
            b = rec[0]; ni = rec[1]; nres = rec[2];
            blk32 = rec + 4;
            pp = rv.arena + rec[3];
          } else {
            b = (uint32_t)(m >> 19) & (MAX_B - 1);
            ni = (uint32_t)(m >> 34) & (MAX_NI - 1);
            nres = (uint32_t)(m >> 48);
            const uint32_t hb = (b + 1) >> 1, H = hb + 2 * ni;  // K1 guarantees H <= HS for in-slot headers
            for (uint32_t w = 0; w < H; ++w) hdr[w] = out[w];
            bw0 = hdr[0];
            bw1 = hdr[1];
            blk32 = nullptr;
            pp = hdr + hb;
          }
          p = 0;
          res = out + (d - nres);
          rj = 0;
          rval = nres ? res[0] : INF;
          ik = 0;
          ival = INF;
          if (ni) { ival = pp[0]; iend = ival + pp[1]; }
          cval = INF;
          if (r) {
            const uint32_t tr = ref_index(rv, t, r);  // exists: K1 rejected the record otherwise
            ref = node_slot(rv, tr);
            dref = (uint32_t)(rv.offs[tr + 1] - rv.offs[tr]);
            ci = 0;
            bk = 0;
            cend = dref;
            if (b) { cend = block_len(0); bk = 1; }
            if (ci < cend) cval = ref[ci];
            else next_copy_block();  // empty first copy block
          }
          st = (d != 0) ? S_MERGE : S_FETCH;
        }
      }
    }
#pragma unroll
    for (int step = 0; step < STEPS_PER_VOTE; ++step) {
      if (st == S_MERGE) {
        const uint32_t mn = min(cval, min(ival, rval));
        out[p] = mn;  // (buffering 4 successors for 16-byte stores was measured: no gain, the kernel is issue-bound)
        if (mn == cval) {
          if (++ci == cend) next_copy_block();
          else cval = ref[ci];
        } else if (mn == rval) {
          rval = ++rj < nres ? res[rj] : INF;
        } else {
          if (++ival == iend) {
            if (++ik < ni) { ival = pp[2 * ik]; iend = ival + pp[2 * ik + 1]; } else ival = INF;
          }
        }
        if (++p == d) st = S_FETCH;
      }
    }
  }
}

// Long reference-free records (outdegree >= BIG0_DEGREE, with intervals): one BLOCK per node instead of one
// lane.  The list is the residuals with the expanded intervals inserted, so every element's final position
// is its index in its own run plus the number of elements of the other run below it:
//   residual j   -> j + (total length of the intervals that start below it)
//   interval k,e -> (lengths of the intervals before k) + e + (number of residuals below its start)
// In place: the residuals sit at the tail of the slot and only move towards the front, so they are moved in
// ascending chunks (a chunk is read completely before it is written); the interval elements are filled in
// afterwards.  Social graphs have such records (power-law degrees); one lane would need ~0.2 us per element.
__global__ void __launch_bounds__(256) k_resolve_big0(RangeView rv, const uint32_t* order, const uint32_t* seg) {
  __shared__ uint32_t s_start[BIG0_MAX_NI], s_pl[BIG0_MAX_NI + 1], s_below[BIG0_MAX_NI];
  __shared__ uint32_t s_scan[256];
  const uint32_t beg = seg[KEY_BIG0], end = seg[KEY_BIG0 + 1];
  for (uint32_t i = beg + blockIdx.x; i < end; i += gridDim.x) {
    const uint32_t t = order[i];
    const uint64_t m = rv.meta[t];
    uint32_t* const out = node_slot(rv, t);
    const uint32_t d = (uint32_t)(rv.offs[t + 1] - rv.offs[t]);
    uint32_t ni, nres;
    const uint32_t* pp;
    if (m & M_OVF) {
      const uint32_t* rec = rv.arena + (uint32_t)(m >> 19);
      ni = rec[1]; nres = rec[2];
      pp = rv.arena + rec[3];
    } else {
      ni = (uint32_t)(m >> 34) & (MAX_NI - 1);
      nres = (uint32_t)(m >> 48);
      pp = out;  // reference-free: no block lengths before the pairs
    }
    const uint32_t* const res = out + (d - nres);
    __syncthreads();  // shared arrays of the previous node are no longer read
    // ---- interval starts, exclusive prefix of their lengths, residuals below each start
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ni; base += 256) {
      const uint32_t k = base + threadIdx.x;
      uint32_t len = 0;
      if (k < ni) {
        const uint32_t st = pp[2 * k];
        len = pp[2 * k + 1];
        s_start[k] = st;
        uint32_t lo = 0, hi = nres;  // residuals < st
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (res[mid] < st) lo = mid + 1; else hi = mid;
        }
        s_below[k] = lo;
      }
      s_scan[threadIdx.x] = len;
      __syncthreads();
      for (uint32_t o = 1; o < 256; o <<= 1) {
        const uint32_t y = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0u;
        __syncthreads();
        s_scan[threadIdx.x] += y;
        __syncthreads();
      }
      if (k < ni) s_pl[k] = carry + s_scan[threadIdx.x] - len;
      carry += s_scan[255];
      __syncthreads();
    }
    if (threadIdx.x == 0) s_pl[ni] = carry;
    __syncthreads();
    const uint32_t total_iv = carry;
    // ---- residuals, in ascending chunks
    for (uint32_t base = 0; base < nres; base += 256) {
      const uint32_t j = base + threadIdx.x;
      uint32_t x = 0, shift = 0;
      if (j < nres) {
        x = res[j];
        uint32_t lo = 0, hi = ni;  // intervals that start below x
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_start[mid] < x) lo = mid + 1; else hi = mid;
        }
        shift = s_pl[lo];
      }
      __syncthreads();  // the whole chunk is read before any of it is overwritten
      if (j < nres) out[j + shift] = x;
    }
    __syncthreads();
    // ---- interval elements
    for (uint32_t e = threadIdx.x; e < total_iv; e += 256) {
      uint32_t lo = 0, hi = ni;  // last interval whose prefix is <= e
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_pl[mid + 1] <= e) lo = mid + 1; else hi = mid;
      }
      out[e + s_below[lo]] = s_start[lo] + (e - s_pl[lo]);
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
  uint32_t target = NOT_FOUND, err = 0;
  if (i < n_in && g.window != 0) {
    const uint64_t v = in[i];
    uint32_t state;
    int64_t ptr;
    load_phase(g, v, state, ptr, err);
    const uint64_t d = ans_decode(g.tb, g.tb.lut, g.tb.ent, Outdegree, state, ptr, g.stream, err);
    if (d != 0 && !err) {
      const uint64_t r = ans_decode(g.tb, g.tb.lut, g.tb.ent, ReferenceOffset, state, ptr, g.stream, err);
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
  unsigned long long cursor;  // arena bump pointer
  uint64_t lo;                // k_halo result
  uint32_t maxlevel;
  uint32_t pad;
  uint32_t hist[16];          // nodes per level bucket
  uint32_t seg[17];           // exclusive prefix of hist
};
static_assert(sizeof(Scalars) <= 256, "Scalars must fit the cleared line");

struct WorkspacePlan {
  uint64_t off_outdeg, off_phase1, off_offs, off_meta, off_lev, off_keys[2], off_vals[2], off_cub, off_halo, off_arena;
  uint64_t cub_bytes, halo_cap, fixed_bytes;
};

WorkspacePlan plan_workspace(uint64_t n) {
  WorkspacePlan p{};
  uint64_t o = 0;
  o += 256;  // Scalars
  p.off_outdeg = o; o = align_up(o + 4 * (n + 1), 256);
  p.off_phase1 = o; o = align_up(o + 16 * n, 256);
  p.off_offs = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_meta = o; o = align_up(o + 8 * n, 256);
  p.off_lev = o; o = align_up(o + 4 * n, 256);
  for (int i = 0; i < 2; ++i) { p.off_keys[i] = o; o = align_up(o + 2 * n, 256); }
  for (int i = 0; i < 2; ++i) { p.off_vals[i] = o; o = align_up(o + 4 * n, 256); }
  size_t scan_bytes = 0, sort_bytes = 0;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, it, (uint64_t*)nullptr, (int64_t)(n + 1));
  cub::DoubleBuffer<uint16_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, (int64_t)n, 0, 12);
  p.cub_bytes = std::max(scan_bytes, sort_bytes);
  p.off_cub = o; o = align_up(o + p.cub_bytes, 256);
  p.halo_cap = 1u << 20;  // successors of halo nodes (u32 each)
  p.off_halo = o; o = align_up(o + 4 * p.halo_cap, 256);
  p.off_arena = o;
  p.fixed_bytes = o;
  return p;
}

}  // namespace

uint64_t decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last) {
  uint64_t n = last - first + 4096;  // room for a halo
  WorkspacePlan p = plan_workspace(n);
  double frac = g->prelude.number_of_nodes ? (double)(last - first) / (double)g->prelude.number_of_nodes : 1.0;
  uint64_t arcs_est = (uint64_t)((double)g->prelude.number_of_arcs * frac) + (1u << 20);
  // arena: headers that do not fit their node's slot (rare)
  uint64_t arena_cap = n + arcs_est / 8 + (1u << 20);
  return p.fixed_bytes + 4 * arena_cap;
}

static void check_device_error(wga_graph* g, uint32_t herr, cudaStream_t st) {
  if (herr) {
    WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    if (herr & ERR_WORKSPACE) throw Error(WGA_E_WORKSPACE, "decode: workspace or output buffer too small; pass larger buffers");
    if (herr & ERR_RANGE) throw Error(WGA_E_CORRUPT, "decode: a reference leaves the decoded range");
    if (herr & ERR_SYMBOL_WIDTH) throw Error(WGA_E_UNSUPPORTED, "decode: a decoded value does not fit 32 bits");
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
  if (ws_bytes < p.fixed_bytes) throw Error(WGA_E_WORKSPACE, "workspace too small");
  uint8_t* w = (uint8_t*)ws;
  uint32_t* outdeg = (uint32_t*)(w + p.off_outdeg);
  k_outdegree<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, first, nullptr, (uint32_t)n, outdeg, nullptr, g->d_err);
  count_launch();
  size_t cb = p.cub_bytes;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(outdeg, U32ToU64());
  WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, d_offsets, (int64_t)(n + 1), st));
  count_launch(2);
  WGA_CUDA(cudaGetLastError());
}

static void mark(wga_graph* g, cudaStream_t st) {
  if (!g->profiling || g->n_ev >= 8) return;
  if (!g->ev[g->n_ev]) cudaEventCreate(&g->ev[g->n_ev]);
  cudaEventRecord(g->ev[g->n_ev++], st);
}

static uint32_t resolve_grid(const Tuning& tn) {
  if (tn.k2_blocks) return tn.k2_blocks;
  static uint32_t cached = 0;
  if (!cached) {
    int dev = 0, sms = 148, per_sm = 8;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resolve, RES_TPB, 0);
    // swept on eu-2015-host-shaped: 10 resident blocks per SM (40 warps) beat 12 -- the lane-private streams of
    // more warps no longer fit L1 -- and anything that is not a whole wave loses to the partial second wave
    per_sm = per_sm > 10 ? 10 : (per_sm > 0 ? per_sm : 1);
    cached = (uint32_t)(sms * per_sm);
  }
  return cached;
}

// K0 .. K2 on the nodes described by rv (a contiguous range, or a sorted node list), then one host
// synchronisation that reads back the totals (tot[0] = halo arcs, tot[1] = all arcs), the deepest level and
// the error word.
static void run_pipeline(wga_graph* g, const RangeView& rv, uint8_t* w, const WorkspacePlan& p, Scalars* sc,
                         const Tuning& tn, cudaStream_t st, uint64_t tot[2]) {
  const uint64_t n = rv.n;
  // ---- K0 + scan
  k_outdegree<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, rv.lo, rv.nodes, (uint32_t)n, rv.outdeg, rv.phase1, g->d_err);
  count_launch();
  {
    size_t cb = p.cub_bytes;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(rv.outdeg, U32ToU64());
    WGA_CUDA(cub::DeviceScan::ExclusiveSum(w + p.off_cub, cb, it, rv.offs, (int64_t)(n + 1), st));
    count_launch(2);
  }
  mark(g, st);  // 1: outdegrees + scan done
  // ---- K1: entropy decode (spans that would overflow the output are skipped and reported)
  {
    uint32_t tpb = tn.k1_tpb < 32 ? 32 : (tn.k1_tpb > 128 ? 128 : tn.k1_tpb / 32 * 32);
    uint32_t span = tn.k1_span ? tn.k1_span : 1;
    // small ranges: shrink the spans so that the grid still fills the machine (148 SMs x 32 blocks)
    span = std::min<uint32_t>(span, std::max<uint32_t>(128u, (uint32_t)(n / (148 * 32))));
    if (rv.nodes) k_entropy<true><<<span_count(rv.n, rv.h, span), tpb, 0, st>>>(g->dev, rv, span, tn.force_ovf);
    else k_entropy<false><<<span_count(rv.n, rv.h, span), tpb, 0, st>>>(g->dev, rv, span, tn.force_ovf);
    count_launch();
  }
  mark(g, st);  // 2: entropy decode done
  // ---- K2: levels, sort by (level, degree), one resolve launch per level
  uint32_t* lev = (uint32_t*)(w + p.off_lev);
  cub::DoubleBuffer<uint16_t> dkeys((uint16_t*)(w + p.off_keys[0]), (uint16_t*)(w + p.off_keys[1]));
  cub::DoubleBuffer<uint32_t> dvals((uint32_t*)(w + p.off_vals[0]), (uint32_t*)(w + p.off_vals[1]));
  const bool have_refs = g->prelude.compression_window != 0 || g->prelude.min_interval_length != 0;
  if (have_refs) {
    k_levels<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rv, dkeys.Current(), dvals.Current(), lev, sc->hist, tn.sort_degree);
    count_launch();
    size_t cb = p.cub_bytes;
    WGA_CUDA(cub::DeviceRadixSort::SortPairs(w + p.off_cub, cb, dkeys, dvals, (int64_t)n, tn.sort_degree ? 0 : 8, 12, st));
    count_launch(3);
    k_segments<<<1, 32, 0, st>>>(sc->hist, sc->seg);
    count_launch();
    mark(g, st);  // 3: levels + sort done
    k_resolve_big0<<<296, 256, 0, st>>>(rv, dvals.Current(), sc->seg);  // (empty on graphs without long records)
    count_launch();
    const uint32_t grid = resolve_grid(tn);
    const uint32_t nlev = g->prelude.compression_window ? LCAP : 1;  // without references everything is level 0
    for (uint32_t l = 0; l < nlev; ++l) {
      k_resolve<<<grid, RES_TPB, 0, st>>>(rv, dvals.Current(), sc->seg, l, 0, lev);
      count_launch();
    }
  } else {
    mark(g, st);
  }
  mark(g, st);  // 4: resolve done
  // ---- totals, deepest level, error word
  uint32_t maxlevel = 0, herr = 0;
  k_publish<<<1, 32, 0, st>>>(rv.offs + rv.h, rv.offs + n, &sc->maxlevel, g->d_err, g->d_pub);
  count_launch();
  WGA_CUDA(cudaStreamSynchronize(st));
  WGA_CUDA(cudaGetLastError());
  tot[0] = g->h_pub[1];
  tot[1] = g->h_pub[2];
  maxlevel = (uint32_t)g->h_pub[3];
  herr = (uint32_t)g->h_pub[4];
  if (tot[0] > rv.halo_cap) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "halo successors exceed the workspace");
  }
  if (tot[1] - tot[0] > rv.succ_cap) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(tot[1] - tot[0]) + " elements");
  }
  check_device_error(g, herr, st);
  // ---- reference chains deeper than LCAP (e.g. graphs compressed with an unbounded max_ref_count): one
  //      launch per extra level over the shared deep segment
  if (have_refs && maxlevel >= LCAP) {
    const uint32_t grid = resolve_grid(tn);
    for (uint32_t l = LCAP; l <= maxlevel; ++l) {
      k_resolve<<<grid, RES_TPB, 0, st>>>(rv, dvals.Current(), sc->seg, LCAP, l, lev);
      count_launch();
    }
    check_device_error(g, read_device_error(g, st), st);
  }
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
  static bool env_done = false;
  if (!env_done) {  // WGA_TUNING="key=value,key=value": same knobs as wga_debug_set_tuning (profiling runs)
    env_done = true;
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
  }
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
  WorkspacePlan p = plan_workspace(n);
  if (ws_bytes < p.fixed_bytes + 4096) throw Error(WGA_E_WORKSPACE, "workspace too small");
  RangeView rv{};
  rv.lo = lo; rv.first = first; rv.n = (uint32_t)n; rv.h = (uint32_t)(first - lo);
  rv.outdeg = (uint32_t*)(w + p.off_outdeg);
  rv.phase1 = (uint4*)(w + p.off_phase1);
  rv.offs = rv.h ? (uint64_t*)(w + p.off_offs) : d_offsets;
  rv.meta = (uint64_t*)(w + p.off_meta);
  rv.arena = (uint32_t*)(w + p.off_arena);
  rv.arena_cap = (ws_bytes - p.off_arena) / 4;
  rv.cursor = &sc->cursor;
  rv.maxlevel = &sc->maxlevel;
  rv.halo_succ = (uint32_t*)(w + p.off_halo);
  rv.halo_cap = p.halo_cap;
  rv.succ = d_succ; rv.succ_cap = succ_capacity;
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
BatchPlan plan_batch(uint64_t nq, uint64_t max_total_arcs) {
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
  b.inner_bytes = p.fixed_bytes + 4 * (b.cap_nodes / 4 + b.cap_arcs / 8 + (1u << 20));
  b.off_inner = o; o += b.inner_bytes;
  b.total = o;
  return b;
}
}  // namespace

uint64_t successors_workspace_size(const wga_graph*, uint64_t n_queries, uint64_t max_total_arcs) {
  return plan_batch(n_queries, max_total_arcs).total;
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
  // the caller sizes the workspace with an upper bound of the arcs it expects; recover it from the size
  BatchPlan b = plan_batch(nq, 0);
  if (ws_bytes < b.total) throw Error(WGA_E_WORKSPACE, "workspace too small");
  {  // largest arc bound whose plan fits this workspace (the plan grows by ~18 bytes per arc)
    uint64_t A = (ws_bytes - b.total) / 18;
    BatchPlan b2 = plan_batch(nq, A);
    while (b2.total > ws_bytes && A) { A = A / 16 * 15; b2 = plan_batch(nq, A); }
    if (b2.total <= ws_bytes) b = b2;
  }
  const Tuning tn = g_tuning;
  uint8_t* w = (uint8_t*)ws;
  uint32_t* scal = (uint32_t*)(w + b.off_scal);  // [0] closure count, [1] unique count
  WGA_CUDA(cudaMemsetAsync(scal, 0, 256, st));
  uint32_t* qid = (uint32_t*)(w + b.off_qid);
  uint32_t* all = (uint32_t*)(w + b.off_all[0]);
  k_query_ids<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_nodes, nq, g->res_first, g->res_last, qid, g->d_err);
  count_launch();
  if (!d_succ) {  // sizing call: only the outdegrees of the queries (first symbol of each record)
    uint32_t* deg = all;
    k_outdegree<<<(unsigned)((nq + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, 0, qid, (uint32_t)nq, deg, nullptr, g->d_err);
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
  rv.lo = 0; rv.first = 0; rv.n = nU; rv.h = 0; rv.nodes = U;
  rv.outdeg = (uint32_t*)(iw + p.off_outdeg);
  rv.phase1 = (uint4*)(iw + p.off_phase1);
  rv.offs = (uint64_t*)(w + b.off_offsU);
  rv.meta = (uint64_t*)(iw + p.off_meta);
  rv.arena = (uint32_t*)(iw + p.off_arena);
  rv.arena_cap = (b.inner_bytes - p.off_arena) / 4;
  rv.cursor = &sc->cursor;
  rv.maxlevel = &sc->maxlevel;
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
