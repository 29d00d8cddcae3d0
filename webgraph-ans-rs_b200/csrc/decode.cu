// =============================================================================
//  decode.cu -- ANS decode of BvGraph components into CSR successor lists (sm_100a)
// =============================================================================
//  Replaces, for whole node ranges at once, what the reference does one symbol at a
//  time on one core:
//    webgraph BvGraphSeq::iter() / BvGraph::successors(v)  (external, un-vendored)
//      -> ANSBVGraphDecoderFactory::new_decoder(v)   src/bvgraph/factories/bvgraph_decoder_factory.rs:46-58
//      -> ANSDecoder::decode(component)              src/ans/decoder.rs:58-100
//  Pipeline (all launches on the caller's stream):
//    K0  k_outdegree   first symbol of every record from (states[N-1-v], pointers[N-1-v]) -> outdegree
//        cub scan      -> CSR offsets
//    K1  k_entropy     phase one: entropy decode of every component of every node.  One node per LANE,
//                      lanes pull nodes from a per-block counter and run a uniform per-symbol state
//                      machine (which component next / how many left), so the instruction stream stays
//                      convergent although records differ.  Decoded values are PARKED inside the node's
//                      own final CSR slot: residuals (already prefix-summed) at the tail, copy-block
//                      lengths (u16) and interval (start,len) pairs at the head.  Nodes that are pure
//                      residual lists are final after K1.
//    K2  k_merge       phase two: copy-block resolution + interval expansion + 3-way merge, streamed
//                      through a shared-memory ring that is a sliding window over the output array.
//                      Merge lanes take nodes in order and emit one successor per step; a lane that
//                      copies from a referenced list follows the producer's per-node progress counter
//                      (element-level wavefront), so reference chains do not serialise.  One writer
//                      warp per block retires finished nodes in order with coalesced stores.
//    K2p k_pend_*      the few nodes whose reference leaves the block's span (or that do not fit the
//                      ring) are resolved afterwards, level by level of the remaining chain depth.
// =============================================================================
#include <cub/cub.cuh>

#include "graph.hpp"

namespace wga {

std::atomic<uint64_t> g_kernel_launches{0};

// run-time tuning (tests shrink these to exercise span boundaries, ring wrap and the overflow paths)
struct Tuning {
  uint32_t k1_span = 2048;    // nodes per K1 block
  uint32_t k1_tpb = 64;       // threads per K1 block
  uint32_t k2_span = 4096;    // nodes per K2 block
  uint32_t k2_tpb = 256;      // threads per K2 block (warp 0 = writer)
  uint32_t ring_log2 = 13;    // K2 ring entries (u32) = 1 << ring_log2
  uint32_t force_ovf = 0;     // K1: put every header in the overflow arena
  uint32_t stats = 0;         // collect wait-reason counters (wga_debug_last_stats)
};
static Tuning g_tuning;
static unsigned long long g_last_stats[16];
void last_stats(uint64_t* out16) { for (int i = 0; i < 16; ++i) out16[i] = g_last_stats[i]; }

int set_tuning(const char* key, uint64_t value) {
  std::string k(key ? key : "");
  if (k == "k1_span") g_tuning.k1_span = (uint32_t)value;
  else if (k == "k1_tpb") g_tuning.k1_tpb = (uint32_t)value;
  else if (k == "k2_span") g_tuning.k2_span = (uint32_t)value;
  else if (k == "k2_tpb") g_tuning.k2_tpb = (uint32_t)value;
  else if (k == "ring_log2") g_tuning.ring_log2 = (uint32_t)value;
  else if (k == "force_ovf") g_tuning.force_ovf = (uint32_t)value;
  else if (k == "stats") g_tuning.stats = (uint32_t)value;
  else if (k == "reset") g_tuning = Tuning();
  else return WGA_E_ARG;
  return WGA_OK;
}

namespace {

constexpr int TPB = 128;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t INF = 0xffffffffu;  // "stream exhausted"; successor ids are <= 0xfffffffe

struct RangeView {
  uint64_t lo;        // first decoded node (halo start)
  uint64_t first;     // first node the caller asked for
  uint32_t n;         // nodes decoded: last - lo
  uint32_t h;         // halo nodes: first - lo
  uint32_t* outdeg;   // n+1
  uint64_t* offs;     // n+1, relative to lo
  uint64_t* meta;     // n : per-node record of K1 (see M_*)
  uint32_t* arena;    // overflow headers (K1) and pass-2 temporaries
  uint64_t arena_cap;
  unsigned long long* cursor;
  uint32_t* pend;     // nodes left to pass 2 (index relative to lo)
  uint32_t pend_cap;
  uint32_t* pend_count;
  uint32_t* pend_lev;
  uint32_t* maxlevel;
  uint32_t* halo_succ;  // successors of halo nodes
  uint64_t halo_cap;
  uint32_t* succ;       // caller's array: successors of nodes >= first
  uint64_t succ_cap;
  uint32_t* err;
  unsigned long long* stats;  // optional debug counters (16 x u64), nullptr when disabled
};

// meta word written by K1:
//   bits 0-15 reference offset r | bit 16 header in the overflow arena | bit 17 node is final after K1
//   bit 18 left to pass 2 (set by K2)
//   in-slot header: bits 19-33 block count b | 34-47 interval count | 48-63 residual count
//   overflow header: bits 19-63 arena offset of {b, ni, nres, pairs offset, blocks...}
constexpr uint64_t M_OVF = 1ull << 16, M_DIRECT = 1ull << 17, M_PEND = 1ull << 18;
constexpr uint32_t MAX_B = 1u << 15, MAX_NI = 1u << 14, MAX_NRES = 1u << 16;

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

  // per-lane record state
  uint32_t c = C_FETCH, t = 0, state = 0, d = 0, r = 0, dref = 0, b = 0, k = 0, copied = 0, pos = 0, extras = 0,
           ni = 0, hb = 0, nres = 0, ao = 0, apo = 0;
  int64_t ptr = 0, v = 0, prev = 0;
  uint32_t* slot = nullptr;
  uint32_t* wp = nullptr;
  bool ovf = false, direct = false;

  // Every lane stays in the loop until the whole warp has run out of nodes: the vote at the top is the
  // per-iteration reconvergence point, so that the symbol decode below runs with all busy lanes together.
  // Each case computes a `bad` flag instead of leaving early, which keeps the cases short and single-exit.
  for (;;) {
    uint32_t err = 0;
    if (c == C_FETCH) {
      t = atomicAdd(&s_next, 1u);
      if (t >= Bn) c = C_IDLE;
      else {
        v = (int64_t)(rv.lo + t);
        load_phase(g, (uint64_t)v, state, ptr, err);
        c = Outdegree;
        r = b = ni = copied = hb = nres = 0;
        ovf = direct = false;
      }
    }
    if (__all_sync(FULL, c == C_IDLE)) break;
    if (c <= Residual) {
      const uint64_t x = ans_decode_cp(s_cp[c], lut, ent, state, ptr, g.stream, err);
      const uint32_t xl = (uint32_t)x;
      const bool wide = (x >> 32) != 0;  // no component value of a valid record needs more than 32 bits
      if (wide && !err) err = ERR_SYMBOL_WIDTH;  // (nat2int arguments: ids < 2^32 give x < 2^33, checked below)
      if (c >= FirstResidual) {
        // ---- residuals: value = node + nat2int(x) | previous + 1 + x   (most frequent symbols)
        if (c == FirstResidual) {
          if (err == ERR_SYMBOL_WIDTH && x <= 0x1FFFFFFFFull) err = 0;
          nres = extras;
          direct = (r == 0 && ni == 0);
          if (!direct && !ovf && (nres >= MAX_NRES || hb + 2ull * ni > (uint64_t)(d - nres))) {
            if (header_to_arena(rv, slot, b, b, ni, ni, ao, apo)) ovf = true;
            else err |= ERR_WORKSPACE;
          }
          wp = slot + (d - nres);
          prev = v + nat2int(x);
        } else {
          prev = prev + 1 + (int64_t)xl;
        }
        if (prev < 0 || prev > 0xFFFFFFFEll) err |= ERR_SYMBOL_WIDTH;
        if (!err) {
          *wp++ = (uint32_t)prev;
          c = --extras ? (uint32_t)Residual : (uint32_t)C_FINISH;
        }
      } else if (c == Blocks) {
        const uint32_t len = xl + (k != 0);
        if (len > dref - pos || len < xl) err |= ERR_CORRUPT;
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
          if (err == ERR_SYMBOL_WIDTH && x <= 0x1FFFFFFFFull) err = 0;
          prev = k == 0 ? v + nat2int(x) : prev + 1 + (int64_t)x;  // prev: start of this interval
          if (prev < 0 || prev > 0xFFFFFFFEll) err |= ERR_SYMBOL_WIDTH;
          if (!err) {
            if (ovf) rv.arena[apo + 2 * k] = (uint32_t)prev;
            else slot[hb + 2 * k] = (uint32_t)prev;
            c = IntervalLen;
          }
        } else {
          const uint64_t len = (uint64_t)xl + minint;
          if (len > extras || len == 0) err |= ERR_CORRUPT;
          prev += (int64_t)len;  // prev: one past the end of this interval
          if (prev > 0xFFFFFFFFll) err |= ERR_SYMBOL_WIDTH;
          if (!err) {
            if (ovf) rv.arena[apo + 2 * k + 1] = (uint32_t)len;
            else slot[hb + 2 * k + 1] = (uint32_t)len;
            extras -= (uint32_t)len;
            if (++k == ni) c = extras ? (uint32_t)FirstResidual : (uint32_t)C_FINISH;
            else c = IntervalStart;
          }
        }
      } else if (c == Outdegree) {
        d = xl;
        extras = d;
        if (!err) {
          if (d == 0) { direct = true; c = C_FINISH; }
          else {
            slot = node_slot(rv, t);
            c = window ? (uint32_t)ReferenceOffset : c_extras;
          }
        }
      } else if (c == ReferenceOffset) {
        if (xl > window) err |= ERR_CORRUPT;
        else if (xl > t) err |= ERR_RANGE;
        if (!err) {
          r = xl;
          if (r == 0) c = c_extras;
          else { dref = rv.outdeg[t - r]; c = BlockCount; }
        }
      } else if (c == BlockCount) {
        if (x > (uint64_t)dref + 1) err |= ERR_CORRUPT;
        if (!err) {
          b = xl;
          hb = (b + 1) >> 1;
          pos = 0;
          k = 0;
          if (b == 0) { copied = dref; c = C_AFTER_BLOCKS; }
          else {
            if (hb > d || b >= MAX_B || dref > 0xFFFFu || force_ovf) {
              if (header_to_arena(rv, slot, b, 0, 0, 0, ao, apo)) ovf = true;
              else err |= ERR_WORKSPACE;
            }
            c = Blocks;
          }
        }
      } else {  // IntervalCount
        if (xl > extras) err |= ERR_CORRUPT;
        if (!err) {
          ni = xl;
          k = 0;
          if (ni == 0) c = FirstResidual;
          else {
            if (ovf) {  // header already in the arena: the pairs get their own piece
              const unsigned long long o = atomicAdd(rv.cursor, 2ull * ni);
              if (o + 2ull * ni > rv.arena_cap || o + 2ull * ni >= 0xFFFFFFFFull) err |= ERR_WORKSPACE;
              else { apo = (uint32_t)o; rv.arena[ao + 3] = apo; }
            } else if (ni >= MAX_NI || hb + 2ull * ni > d || force_ovf) {
              if (header_to_arena(rv, slot, b, b, ni, 0, ao, apo)) ovf = true;
              else err |= ERR_WORKSPACE;
            }
            c = IntervalStart;
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
}

// -------------------------------------------------------------------------------------------- K2
// The three sorted streams a successor list is the union of (webgraph BvGraph decode, SURVEY.md 8a
// "BV record order"): copied elements of the referenced list selected by the copy blocks, expanded
// intervals, residuals.  `RefSrc` supplies element i of the referenced list.
struct NodeStreams {
  // header
  const uint32_t* blkw;  // in-slot header: word j = copy length 2j (low half) | skip length 2j+1 (high half)
  const uint32_t* blk32; // overflow header: one u32 per block (nullptr when in-slot)
  const uint32_t* pp;    // interval pairs
  const uint32_t* rp;    // residuals
  uint32_t b, ni, nres, dref;
  // copy stream: [ci, cend) is the current copy block; bk = index of the next (skip) block.
  // cur = header word holding the skip length that follows the current copy block, nxt = the word after it;
  // both are loaded one step ahead so that the merge loop never waits for HBM/L2.
  uint32_t ci, cend, bk, cur, nxt;
  bool cact;
  // interval stream (next pair prefetched)
  uint32_t ik, ival, iend, nis, nil;
  // residual stream (next value prefetched)
  uint32_t rj, rval, rnext;

  // gs = the node's CSR slot (d entries) holding the parked header / residuals
  __device__ __forceinline__ void setup(const RangeView& rv, uint64_t m, const uint32_t* gs, uint32_t d, uint32_t dref_) {
    if (m & M_OVF) {
      const uint32_t* rec = rv.arena + (uint32_t)(m >> 19);
      b = rec[0]; ni = rec[1]; nres = rec[2];
      blk32 = rec + 4; blkw = nullptr;
      pp = rv.arena + rec[3];
    } else {
      b = (uint32_t)(m >> 19) & (MAX_B - 1);
      ni = (uint32_t)(m >> 34) & (MAX_NI - 1);
      nres = (uint32_t)(m >> 48);
      blkw = gs; blk32 = nullptr;
      pp = gs + ((b + 1) >> 1);
    }
    rp = gs + (d - nres);
    rj = 0;
    rval = nres ? rp[0] : INF;
    rnext = nres > 1 ? rp[1] : INF;
    ik = 0;
    ival = iend = nis = nil = INF;
    if (ni) { ival = pp[0]; iend = ival + pp[1]; }
    if (ni > 1) { nis = pp[2]; nil = pp[3]; }
    dref = dref_;
    cact = false;
    ci = cend = bk = cur = nxt = 0;
    if ((uint32_t)(m & 0xFFFFu)) {
      cact = true;
      if (b == 0) cend = dref;
      else if (blk32) { cend = blk32[0]; bk = 1; }
      else {
        cur = blkw[0];
        nxt = b > 2 ? blkw[1] : 0u;
        cend = cur & 0xFFFFu;
        bk = 1;
      }
      if (ci >= cend) next_copy_block();
    }
  }
  // current copy block exhausted: skip block, then the next copy block (explicit, or the implicit tail
  // when the block count is even)
  __device__ __forceinline__ void next_copy_block() {
    if (bk >= b) { cact = false; return; }
    if (blk32) {
      ci += blk32[bk]; ++bk;
      if (bk < b) { cend = ci + blk32[bk]; ++bk; } else cend = dref;
    } else {
      ci += cur >> 16; ++bk;           // skip block bk (odd) lives in the high half of the current word
      cur = nxt;
      if (bk < b) {
        cend = ci + (cur & 0xFFFFu); ++bk;  // copy block bk (even): low half of the next word
        nxt = (bk + 1 < b) ? blkw[(bk + 1) >> 1] : 0u;
      } else cend = dref;
    }
    if (ci >= cend) cact = false;
  }
  __device__ __forceinline__ void take_copy() { if (++ci == cend) next_copy_block(); }
  __device__ __forceinline__ void take_interval() {
    if (++ival == iend) {
      ++ik;
      ival = nis;
      iend = nis + nil;
      if (ik >= ni) ival = INF;
      if (ik + 1 < ni) { nis = pp[2 * ik + 2]; nil = pp[2 * ik + 3]; } else nis = INF;
    }
  }
  __device__ __forceinline__ void take_residual() {
    ++rj;
    rval = rnext;
    rnext = (rj + 1 < nres) ? rp[rj + 1] : INF;
  }
};

constexpr uint32_t FLN = 512;  // per-block window of in-flight nodes (power of two)
constexpr uint32_t ST_RING = 0, ST_DIRECT = 1, ST_POISON = 2;
// flag word of node k (relative to the span): tag (k+1) in bits 16-31 | status in bits 14-15 | progress
// (elements of the list already in the ring) in bits 0-13.  A tag mismatch means "not announced yet".
constexpr uint32_t SPIN_LIMIT = 1u << 24;
// Ordering of the ring protocol.  A producer stores list elements into the ring and then its progress word;
// a consumer loads the progress word and then the elements.  All of these are shared-memory accesses of one
// SM, issued in program order by each thread, and the SM performs one warp's shared-memory accesses in issue
// order, so a compiler barrier (no reordering by nvcc) is all that is needed.  A real fence
// (__threadfence_block = MEMBAR.SC.CTA) would also wait for the thread's outstanding GLOBAL loads/stores --
// the prefetches of the merge lanes, the output stores of the writer -- once per element: measured 2.5k
// cycles per merge step.  Parity tests run with tiny rings / spans to exercise this protocol.
#define SMEM_ORDER() asm volatile("" ::: "memory")
constexpr uint32_t ERR_INTERNAL = 16u;

__device__ __forceinline__ void pend_push(const RangeView& rv, uint32_t t, uint64_t m) {
  const uint32_t i = atomicAdd(rv.pend_count, 1u);
  if (i < rv.pend_cap) rv.pend[i] = t;
  else atomicOr(rv.err, ERR_WORKSPACE);
  rv.meta[t] = m | M_PEND;
}

// Per-block shared state of k_merge.  All per-node arrays are indexed by (k & (FLN-1)), k relative to the span.
struct MergeShared {
  uint64_t meta[FLN];   // K1 record of the node (copied from HBM by the dispatcher, coalesced)
  uint32_t flag[FLN];   // tag | status | progress (see above)
  uint32_t rpos[FLN];   // ring position of the node's list: prefix sum of the degrees of ring-eligible nodes
  uint32_t pos[FLN];    // output position of the node's list relative to the span start
  uint32_t deg[FLN];    // outdegree
  uint32_t next;        // next node to hand to a merge lane
  uint32_t disp;        // nodes whose rpos/pos/deg/meta are valid
  uint32_t flushed;     // nodes retired by the writer
  uint32_t free_rpos;   // ring positions below this one may be overwritten
};

__global__ void __launch_bounds__(256, 4) k_merge(DevGraph g, RangeView rv, uint32_t span, uint32_t ring_mask,
                                                uint32_t dbig) {
  extern __shared__ __align__(16) uint8_t smraw[];
  MergeShared& S = *reinterpret_cast<MergeShared*>(smraw);
  volatile uint32_t* ring = reinterpret_cast<volatile uint32_t*>(smraw + sizeof(MergeShared));
  volatile uint32_t* flag = S.flag;
  volatile uint32_t* v_disp = &S.disp;
  volatile uint32_t* v_flushed = &S.flushed;
  volatile uint32_t* v_free = &S.free_rpos;
  uint32_t A, Bn;
  span_range(rv, span, blockIdx.x, A, Bn);
  if (span_overflows(rv, A, Bn)) {
    if (threadIdx.x == 0) atomicOr(rv.err, ERR_WORKSPACE);
    return;
  }
  const uint32_t nspan = Bn - A;
  for (uint32_t i = threadIdx.x; i < FLN; i += blockDim.x) S.flag[i] = 0;
  if (threadIdx.x == 0) { S.next = 0; S.disp = 0; S.flushed = 0; S.free_rpos = 0; }
  __syncthreads();
  const uint64_t obase = rv.offs[A];
  uint32_t* const gbase = node_slot(rv, A);  // global address of output position 0 of the span
  if (rv.offs[Bn] - obase >= 0xFFFFFFFFull) dbig = 0;  // positions are kept in 32 bits: leave everything to pass 2
  const uint32_t W = dbig ? g.window : 0u;  // nothing is ring-resident when dbig == 0: no list has to be kept
  const uint32_t C = ring_mask + 1;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x < 32) {
    // ------------------------------------------------ warp 0: dispatcher (runs ahead) + writer (in-order retirement)
    uint32_t kd = 0, kw = 0, rbase = 0, spins = 0, w_iter = 0, w_disp = 0, w_retire = 0, w_idle = 0;
    // software-pipelined loads of the next dispatch batch
    uint64_t nm = M_DIRECT, np0 = 0, np1 = 0;
    auto prefetch = [&](uint32_t k0) {
      const uint32_t k = k0 + lane;
      nm = M_DIRECT; np0 = np1 = 0;
      if (k < nspan) { nm = rv.meta[A + k]; np0 = rv.offs[A + k] - obase; np1 = rv.offs[A + k + 1] - obase; }
    };
    prefetch(0);
    while (kw < nspan) {
      bool progressed = false;
      ++w_iter;
      // ---- dispatch one batch of 32 nodes when the window has room for it
      if (kd < nspan && kd + 32 + W + 2 <= kw + FLN) {
        const uint32_t k = kd + lane;
        const bool valid = k < nspan;
        const uint32_t d = (uint32_t)(np1 - np0);
        const bool elig = valid && !(nm & M_DIRECT) && d != 0 && d <= dbig;
        uint32_t sz = elig ? d : 0u, inc = sz;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(FULL, inc, o);
          if ((int)lane >= o) inc += y;
        }
        if (valid) {
          const uint32_t sl = k & (FLN - 1);
          S.meta[sl] = nm;
          S.rpos[sl] = rbase + inc - sz;
          S.pos[sl] = (uint32_t)np0;
          S.deg[sl] = d;
        }
        rbase += __shfl_sync(FULL, inc, 31);
        kd = min(kd + 32, nspan);
        prefetch(kd);
        SMEM_ORDER();
        __syncwarp();
        if (lane == 0) *v_disp = kd;
        progressed = true;
        ++w_disp;
      }
      // ---- retire the longest prefix of finished nodes
      {
        const uint32_t k = kw + lane;
        const bool valid = k < kd;
        const uint32_t sl = k & (FLN - 1);
        uint32_t w = 0, dk = 0;
        bool ok = false;
        if (valid) {
          w = flag[sl];
          dk = S.deg[sl];
          ok = (w >> 16) == ((k + 1) & 0xFFFFu) && (((w >> 14) & 3u) != ST_RING || (w & 0x3FFFu) == dk);
        }
        const uint32_t mk = __ballot_sync(FULL, ok);
        const uint32_t run = (mk == FULL) ? 32u : (uint32_t)__ffs(~mk) - 1u;
        if (run) {
          uint32_t rm = __ballot_sync(FULL, ok && ((w >> 14) & 3u) == ST_RING);
          if (run < 32) rm &= (1u << run) - 1u;
          const uint32_t mypos = valid ? S.pos[sl] : 0u, myrpos = valid ? S.rpos[sl] : 0u;
          while (rm) {  // maximal runs of ring-resident nodes -> flat coalesced copies
            const int s0 = __ffs(rm) - 1;
            const uint32_t y = ~(rm >> s0);
            const int len = y ? __ffs(y) - 1 : 32 - s0;
            const uint32_t ps = __shfl_sync(FULL, mypos, s0);
            const uint32_t rs = __shfl_sync(FULL, myrpos, s0);
            const uint32_t cnt = __shfl_sync(FULL, mypos + dk, s0 + len - 1) - ps;
            for (uint32_t q = lane; q < cnt; q += 32) gbase[ps + q] = ring[(rs + q) & ring_mask];
            rm &= (s0 + len >= 32) ? 0u : ~((1u << (s0 + len)) - 1u);
          }
          kw += run;
          __syncwarp();
          if (lane == 0) {
            // lists of the last W retired nodes stay (they may still be referenced)
            const uint32_t fr = kw > W ? S.rpos[(kw - W) & (FLN - 1)] : 0u;
            SMEM_ORDER();
            *v_free = fr;
            *v_flushed = kw;
          }
          progressed = true;
          ++w_retire;
        }
      }
      if (progressed) spins = 0;
      else {
        ++w_idle;
        __nanosleep(32);
        if (++spins > SPIN_LIMIT) {
          if (lane == 0) atomicOr(rv.err, ERR_INTERNAL);
          return;
        }
      }
    }
    if (rv.stats && lane == 0) {
      atomicAdd(rv.stats + 8, w_iter); atomicAdd(rv.stats + 9, w_disp); atomicAdd(rv.stats + 10, w_retire);
      atomicAdd(rv.stats + 11, w_idle);
    }
    return;
  }

  // ------------------------------------------------------------------ merge lanes
  // Per-lane state machine.  Control steps (fetch a node, wait for the dispatcher / ring space / the
  // referenced node, set the streams up) are batched: the warp runs them only when CTL_BATCH lanes need
  // one (or periodically), so that the common iteration is the short merge step executed by all lanes.
  enum { S_FETCH, S_DISPATCH, S_WAIT, S_MERGE, S_DONE };
  constexpr uint32_t CTL_BATCH = 8;
  int st = S_FETCH;
  uint32_t k = 0, d = 0, r = 0, p = 0, kslot = 0, ktag = 0, jslot = 0, jtag = 0, rb = 0, wb = 0, cval = 0, idle = 0,
           it = 0;
  uint64_t m = 0;
  const uint32_t* refg = nullptr;
  bool havec = false;
  NodeStreams ns;
  uint32_t n_iter = 0, n_ctl = 0, n_wdisp = 0, n_wspace = 0, n_wref = 0, n_mstall = 0, n_mwork = 0, n_done = 0;
  for (;; ++it) {
    const uint32_t ctl = __ballot_sync(FULL, st != S_MERGE && st != S_DONE);
    const uint32_t mrg = __ballot_sync(FULL, st == S_MERGE);
    if ((ctl | mrg) == 0) break;  // every lane is done
    ++n_iter;
    if (st == S_DONE) ++n_done;
    if (__popc(ctl) >= CTL_BATCH || mrg == 0 || (ctl && (it & 7u) == 0)) {
      ++n_ctl;
      if (st == S_FETCH) {
        k = atomicAdd(&S.next, 1u);
        st = k >= nspan ? S_DONE : S_DISPATCH;
        idle = 0;
      }
      if (st == S_DISPATCH) {
        if (k < *v_disp) {
          kslot = k & (FLN - 1);
          m = S.meta[kslot];
          d = S.deg[kslot];
          wb = S.rpos[kslot];
          ktag = ((k + 1) & 0xFFFFu) << 16;
          r = (uint32_t)(m & 0xFFFFu);
          if ((m & M_DIRECT) || d == 0) {
            flag[kslot] = ktag | (ST_DIRECT << 14);
            st = S_FETCH;
          } else if (d > dbig || r > k) {  // does not fit the ring / reference before the span: pass 2
            pend_push(rv, A + k, m);
            flag[kslot] = ktag | (ST_POISON << 14);
            st = S_FETCH;
          } else {
            jslot = (k - r) & (FLN - 1);
            jtag = (k - r + 1) & 0xFFFFu;
            st = S_WAIT;
          }
        } else { ++idle; ++n_wdisp; }  // the dispatcher has not reached this node yet
      }
      if (st == S_WAIT) {
        bool ready = (wb + d - *v_free) <= C;
        uint32_t js = ST_DIRECT;
        if (ready && r) {
          const uint32_t w = flag[jslot];
          ready = (w >> 16) == jtag;
          js = (w >> 14) & 3u;
        }
        if (!ready) { ++idle; if ((wb + d - *v_free) > C) ++n_wspace; else ++n_wref; }
        else if (r && js == ST_POISON) {
          pend_push(rv, A + k, m);
          flag[kslot] = ktag | (ST_POISON << 14);
          st = S_FETCH;
        } else {
          uint32_t dref = 0;
          refg = nullptr;
          if (r) {
            dref = S.deg[jslot];
            if (js == ST_DIRECT) refg = gbase + S.pos[jslot];
            else rb = S.rpos[jslot];
          }
          ns.setup(rv, m, gbase + S.pos[kslot], d, dref);
          p = 0;
          havec = false;
          flag[kslot] = ktag;  // announce: ring-resident, nothing written yet
          st = S_MERGE;
          idle = 0;
        }
      }
    }
    if (st == S_MERGE) {  // one successor per step
      if (ns.cact && !havec) {
        uint32_t avail = ns.dref;
        if (!refg) avail = flag[jslot] & 0x3FFFu;
        if (ns.ci < avail) {
          cval = refg ? refg[ns.ci] : ring[(rb + ns.ci) & ring_mask];
          havec = true;
        }
      }
      if (!ns.cact || havec) {
        const uint32_t cv = ns.cact ? cval : INF;
        const uint32_t mn = min(cv, min(ns.ival, ns.rval));
        ring[(wb + p) & ring_mask] = mn;
        ++p;
        if (mn == INF) {  // cannot happen on records K1 accepted
          atomicOr(rv.err, ERR_CORRUPT);
          p = d;
        } else if (mn == cv) { havec = false; ns.take_copy(); }
        else if (mn == ns.ival) ns.take_interval();
        else ns.take_residual();
        SMEM_ORDER();
        flag[kslot] = ktag | p;
        if (p == d) st = S_FETCH;
        idle = 0;
        ++n_mwork;
      } else { ++idle; ++n_mstall; }
    }
    if (idle > SPIN_LIMIT) {
      atomicOr(rv.err, ERR_INTERNAL);
      st = S_DONE;
    }
  }
  if (rv.stats) {
    atomicAdd(rv.stats + 0, n_iter); atomicAdd(rv.stats + 1, n_ctl); atomicAdd(rv.stats + 2, n_wdisp);
    atomicAdd(rv.stats + 3, n_wspace); atomicAdd(rv.stats + 4, n_wref); atomicAdd(rv.stats + 5, n_mstall);
    atomicAdd(rv.stats + 6, n_mwork); atomicAdd(rv.stats + 7, n_done);
  }
}

// -------------------------------------------------------------------------------------------- K2p
// Chain depth of every pending node inside the pending set (1 = its reference is final).
__global__ void __launch_bounds__(TPB) k_pend_levels(RangeView rv, uint32_t np) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lev = 0;
  if (i < np) {
    uint32_t u = rv.pend[i];
    lev = 1;
    for (;;) {
      const uint32_t r = (uint32_t)(rv.meta[u] & 0xFFFFu);
      if (!r) break;
      u -= r;
      if (!(rv.meta[u] & M_PEND)) break;
      ++lev;
    }
    rv.pend_lev[i] = lev;
  }
  for (int o = 16; o; o >>= 1) lev = max(lev, __shfl_xor_sync(FULL, lev, o));
  if ((threadIdx.x & 31) == 0 && lev) atomicMax(rv.maxlevel, lev);
}

// One pending node per lane: merge from global memory into an arena temporary, then copy over the slot
// (the slot still holds the parked header / residuals while they are being read).
__global__ void __launch_bounds__(TPB) k_pend_resolve(DevGraph g, RangeView rv, uint32_t np, uint32_t lev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np || rv.pend_lev[i] != lev) return;
  const uint32_t t = rv.pend[i];
  const uint64_t m = rv.meta[t];
  const uint32_t r = (uint32_t)(m & 0xFFFFu);
  uint32_t* const gs = node_slot(rv, t);
  const uint32_t d = (uint32_t)(rv.offs[t + 1] - rv.offs[t]);
  const uint32_t* ref = nullptr;
  uint32_t dref = 0;
  if (r) {
    ref = node_slot(rv, t - r);
    dref = (uint32_t)(rv.offs[t - r + 1] - rv.offs[t - r]);
  }
  const unsigned long long o = atomicAdd(rv.cursor, (unsigned long long)d);
  if (o + d > rv.arena_cap) { atomicOr(rv.err, ERR_WORKSPACE); return; }
  uint32_t* tmp = rv.arena + o;
  NodeStreams ns;
  ns.setup(rv, m, gs, d, dref);
  for (uint32_t p = 0; p < d; ++p) {
    const uint32_t cv = ns.cact ? ref[ns.ci] : INF;
    const uint32_t mn = min(cv, min(ns.ival, ns.rval));
    tmp[p] = mn;
    if (mn == INF) { atomicOr(rv.err, ERR_CORRUPT); break; }
    if (mn == cv) ns.take_copy();
    else if (mn == ns.ival) ns.take_interval();
    else ns.take_residual();
  }
  for (uint32_t p = 0; p < d; ++p) gs[p] = tmp[p];
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

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Scalars at the head of the workspace (one 256-byte line, cleared per call).
struct Scalars {
  unsigned long long cursor;  // arena bump pointer
  uint64_t lo;                // k_halo result
  uint32_t maxlevel;
  uint32_t pend_count;
  uint64_t pad[5];
  unsigned long long stats[16];  // offset 64
};

struct WorkspacePlan {
  uint64_t off_outdeg, off_offs, off_meta, off_pend, off_pend_lev, off_cub, off_halo, off_arena;
  uint64_t cub_bytes, halo_cap, pend_cap, fixed_bytes;
};

WorkspacePlan plan_workspace(uint64_t n) {
  WorkspacePlan p{};
  uint64_t o = 0;
  o += 256;  // Scalars
  p.off_outdeg = o; o = align_up(o + 4 * (n + 1), 256);
  p.off_offs = o; o = align_up(o + 8 * (n + 1), 256);
  p.off_meta = o; o = align_up(o + 8 * n, 256);
  p.pend_cap = n / 8 + 65536;
  if (p.pend_cap > n) p.pend_cap = n + 1;
  p.off_pend = o; o = align_up(o + 4 * p.pend_cap, 256);
  p.off_pend_lev = o; o = align_up(o + 4 * p.pend_cap, 256);
  size_t cub_bytes = 0;
  cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, it, (uint64_t*)nullptr, (int64_t)(n + 1));
  p.cub_bytes = cub_bytes;
  p.off_cub = o; o = align_up(o + cub_bytes, 256);
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
  // arena: overflow headers (rare) + pass-2 temporaries (successors of the nodes left to pass 2)
  uint64_t arena_cap = n + arcs_est / 8 + (1u << 20);
  return p.fixed_bytes + 4 * arena_cap;
}

static void check_device_error(wga_graph* g, uint32_t herr, cudaStream_t st) {
  if (herr) {
    WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    if (herr & ERR_INTERNAL) throw Error(WGA_E_CUDA, "decode: internal scheduling error (spin limit reached)");
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
  k_outdegree<<<(unsigned)((n + 1 + TPB - 1) / TPB), TPB, 0, st>>>(g->dev, first, (uint32_t)n, outdeg, g->d_err);
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
    k_halo<<<1, 32, 0, st>>>(g->dev, first, last, &sc->lo, g->d_err);
    count_launch();
    WGA_CUDA(cudaMemcpyAsync(&lo, &sc->lo, 8, cudaMemcpyDeviceToHost, st));
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
  rv.arena = (uint32_t*)(w + p.off_arena);
  rv.arena_cap = (ws_bytes - p.off_arena) / 4;
  rv.cursor = &sc->cursor;
  rv.pend = (uint32_t*)(w + p.off_pend);
  rv.pend_cap = (uint32_t)p.pend_cap;
  rv.pend_count = &sc->pend_count;
  rv.pend_lev = (uint32_t*)(w + p.off_pend_lev);
  rv.maxlevel = &sc->maxlevel;
  rv.halo_succ = (uint32_t*)(w + p.off_halo);
  rv.halo_cap = p.halo_cap;
  rv.succ = d_succ; rv.succ_cap = succ_capacity;
  rv.err = g->d_err;
  rv.stats = tn.stats ? sc->stats : nullptr;
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
  // ---- K1: entropy decode (spans that would overflow the output are skipped and reported)
  {
    uint32_t tpb = tn.k1_tpb < 32 ? 32 : (tn.k1_tpb > 128 ? 128 : tn.k1_tpb / 32 * 32);
    uint32_t span = tn.k1_span ? tn.k1_span : 1;
    // small ranges: shrink the spans so that the grid still fills the machine (148 SMs x 32 blocks)
    span = std::min<uint32_t>(span, std::max<uint32_t>(128u, (uint32_t)(n / (148 * 32))));
    k_entropy<<<span_count(rv.n, rv.h, span), tpb, 0, st>>>(g->dev, rv, span, tn.force_ovf);
    count_launch();
  }
  mark(g, st);  // 2: entropy decode done
  // ---- K2: streamed merge
  {
    uint32_t ring_log2 = tn.ring_log2 < 6 ? 6 : (tn.ring_log2 > 15 ? 15 : tn.ring_log2);
    uint32_t C = 1u << ring_log2;
    uint32_t W = (uint32_t)g->prelude.compression_window;
    uint32_t dbig = C / (W + 1);
    if (dbig > 0x3FFFu) dbig = 0x3FFFu;
    if (W + 2 + 64 >= FLN) dbig = 0;  // window too wide for the in-flight node window: everything goes to pass 2
    uint32_t span = tn.k2_span ? tn.k2_span : 1;
    span = std::min<uint32_t>(span, std::max<uint32_t>(512u, (uint32_t)(n / (148 * 8))));
    if (span > 32768) span = 32768;  // node tags are 16 bits
    uint32_t tpb = tn.k2_tpb < 64 ? 64 : (tn.k2_tpb > 256 ? 256 : tn.k2_tpb / 32 * 32);
    size_t smem = sizeof(MergeShared) + (size_t)C * 4;
    static bool attr_set = false;
    if (!attr_set) {
      WGA_CUDA(cudaFuncSetAttribute(k_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(MergeShared) + (4u << 15))));
      attr_set = true;
    }
    k_merge<<<span_count(rv.n, rv.h, span), tpb, smem, st>>>(g->dev, rv, span, C - 1, dbig);
    count_launch();
  }
  mark(g, st);  // 3: merge done
  // ---- results of the fast path: totals, pending count, error word
  uint64_t tot[2] = {0, 0};
  uint32_t np = 0, herr = 0;
  WGA_CUDA(cudaMemcpyAsync(&tot[0], rv.offs + rv.h, 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(&tot[1], rv.offs + n, 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(&np, &sc->pend_count, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(&herr, g->d_err, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  WGA_CUDA(cudaGetLastError());
  if (tot[0] > rv.halo_cap) { check_device_error(g, herr & ~ERR_WORKSPACE, st); if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st)); throw Error(WGA_E_WORKSPACE, "halo successors exceed the workspace"); }
  if (tot[1] - tot[0] > succ_capacity) {
    if (herr) WGA_CUDA(cudaMemsetAsync(g->d_err, 0, 4, st));
    throw Error(WGA_E_WORKSPACE, "d_succ too small: need " + std::to_string(tot[1] - tot[0]) + " elements");
  }
  if (tn.stats) WGA_CUDA(cudaMemcpy(g_last_stats, sc->stats, sizeof(g_last_stats), cudaMemcpyDeviceToHost));
  check_device_error(g, herr, st);
  // ---- K2p: nodes whose reference left their span, level by level
  if (np) {
    if (np > rv.pend_cap) throw Error(WGA_E_WORKSPACE, "pending list exceeds the workspace");
    const unsigned pgrid = (np + TPB - 1) / TPB;
    k_pend_levels<<<pgrid, TPB, 0, st>>>(rv, np);
    count_launch();
    uint32_t maxlevel = 0;
    WGA_CUDA(cudaMemcpyAsync(&maxlevel, &sc->maxlevel, 4, cudaMemcpyDeviceToHost, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    for (uint32_t lev = 1; lev <= maxlevel; ++lev) {
      k_pend_resolve<<<pgrid, TPB, 0, st>>>(g->dev, rv, np, lev);
      count_launch();
    }
    check_device_error(g, read_device_error(g, st), st);
  }
  if (rv.h) {  // hand the caller offsets relative to `first`
    const uint64_t cnt = last - first + 1;
    k_offsets_rebase<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(rv.offs + rv.h, tot[0], d_offsets, cnt);
    count_launch();
  }
  mark(g, st);  // 4: pass 2 done
  WGA_CUDA(cudaGetLastError());
  if (g->profiling) {
    WGA_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 8; ++i) g->stage_ms[i] = 0.f;
    for (int i = 1; i < g->n_ev; ++i) cudaEventElapsedTime(&g->stage_ms[i - 1], g->ev[i - 1], g->ev[i]);
  }
  if (h_arcs) *h_arcs = tot[1] - tot[0];
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
