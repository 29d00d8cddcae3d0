// Device-side view of an ANS graph and the symbol decoder (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "common.hpp"

namespace wga {

// Packed tables as the kernels see them (global memory; K1 stages them in shared memory).  See common.hpp.
struct DevTables {
  const uint2* bkt;  // {start mask, owner of the first slot} per bucket of 32 slots
  const uint2* ent;  // {cumul | freq << 16, base | folds << 16} per non-zero symbol, + one sentinel per component
  uint32_t bkt_off[WGA_COMPONENTS];
  uint32_t ent_off[WGA_COMPONENTS];
  uint32_t nb[WGA_COMPONENTS];    // buckets
  uint32_t nent[WGA_COMPONENTS];  // entries including the sentinel
  uint8_t L[WGA_COMPONENTS];
  uint8_t R[WGA_COMPONENTS];
};

struct DevGraph {
  DevTables tb;
  const uint16_t* stream;  // word 0 of the resident span
  uint64_t stream_base;    // absolute index of stream[0] (shards)
  uint64_t stream_words;   // resident words
  const uint32_t* states;  // reversed node order (as on disk): node v lives at states[top - v]
  const uint64_t* ptrs;    // same order, absolute u16-word indices into the whole stream
  uint64_t top;            // = (last resident node); whole graph: N-1  (bvgraph_decoder_factory.rs:49-50)
  uint64_t N;
  uint32_t window;
  uint32_t min_interval;
};

// error bits written to the device error word
enum : uint32_t {
  ERR_CORRUPT = 1u,       // stream / table inconsistency (reference would panic or return garbage)
  ERR_WORKSPACE = 2u,     // workspace or output buffer too small
  ERR_RANGE = 4u,         // a reference leaves the decoded range (halo missing)
  ERR_SYMBOL_WIDTH = 8u,  // decoded value does not fit 32 bits
  ERR_LIMIT = 16u,        // a documented implementation limit (copy-block count, record length)
};

#define WGA_LOWER_BOUND 65536u  // INTERVAL_LOWER_BOUND, src/ans/mod.rs:21

// Per-component decode parameters packed for one 16-byte load:
//   x = frame mask (2^L-1) | L << 16 | R << 21     y = bucket offset     z = entry offset     w = floor(65536/R)+1
__host__ __device__ inline uint4 comp_params(const DevTables& tb, int c) {
  const uint32_t R = tb.R[c] ? tb.R[c] : 1u;
  const uint32_t L = tb.L[c];
  return make_uint4(((1u << L) - 1u) | (L << 16) | (R << 21), tb.bkt_off[c], tb.ent_off[c], 65536u / R + 1u);
}

// Table accessors.  Global: both levels through the read-only path.
struct GlobalTables {
  const uint2* bkt;
  const uint2* ent;
  __device__ __forceinline__ uint2 bucket(uint32_t i) const { return __ldg(bkt + i); }
  __device__ __forceinline__ uint2 entry(const uint4&, uint32_t off, uint32_t j) const { return __ldg(ent + off + j); }
  __device__ __forceinline__ uint32_t recip(const uint4& cp) const { return cp.w; }
};

// The decoder of one record: (state, index of the next word below, prefetched next word).
// The next word is loaded right after the previous extend consumed one, so that the 16-bit renormalisation
// (decoder.rs:89-93) never waits for memory: the load overlaps the table lookups of the following symbol.
// The resident span is preceded by 16 readable padding bytes, so the prefetch needs no bounds test (stream[-1]).
struct Dec {
  uint32_t state;
  uint32_t sp;  // words of the resident span below the record's read position
  uint32_t w;   // stream[sp-1]
};

__device__ __forceinline__ void dec_prime(Dec& d, const uint16_t* __restrict__ stream) {
  d.w = (uint32_t)__ldg(stream + (int64_t)d.sp - 1);
}

// One 16-bit extend (decoder.rs:89-93).  false = the stream is exhausted.
__device__ __forceinline__ bool ans_extend(Dec& d, const uint16_t* __restrict__ stream) {
  if (d.sp == 0) return false;
  d.state = (d.state << 16) | d.w;
  --d.sp;
  d.w = (uint32_t)__ldg(stream + (int64_t)d.sp - 1);
  return true;
}

// Order of the R-bit groups of x reversed (n groups, n*R <= 32, or <= 64 for the wide decoder).
__device__ __forceinline__ uint32_t reverse_groups(uint32_t x, uint32_t n, uint32_t R) {
  if (R == 1) return __brev(x) >> (32u - n);
  const uint32_t rmask = (1u << R) - 1u;
  uint32_t out = 0;
  for (; n; --n) { out = (out << R) | (x & rmask); x >>= R; }
  return out;
}
__device__ __forceinline__ uint64_t reverse_groups(uint64_t x, uint32_t n, uint32_t R) {
  const uint64_t rmask = (1ull << R) - 1ull;
  uint64_t out = 0;
  for (; n; --n) { out = (out << R) | (x & rmask); x >>= R; }
  return out;
}

// One ANS symbol.  Restates ANSDecoder::decode (src/ans/decoder.rs:58-87) on the packed tables:
//   slot  = state & (2^L-1)                                   decoder.rs:59
//   entry = owner(slot)                                       decoder.rs:60  (bucket + popcount of the start mask)
//   state = (state >> L)*freq + slot - cumul                  decoder.rs:62-65
//   one conditional 16-bit extend                             decoder.rs:67-69, 89-93
//   folds x { [extend]; fold=(fold<<R)|(state&(2^R-1)); state>>=R; [extend] }   decoder.rs:74-85
//   result = (base << folds*R) | fold                         decoder.rs:86 with quasi_fold (model4decoder.rs:56-68)
// The fold loop of the reference takes R bits per trip (up to 38 trips).  Between two extends the trips only shift
// the state, so they are done in one step per extend: with n = bit length of the state, the next
// t = ceil((n-16)/R) trips need no extend (the state stays >= 2^16 until the t-th shift) and consume the low t*R
// bits.  The bits are collected in the order they leave the state; the reference puts the first chunk highest, so
// the R-bit groups are reversed once at the end (nothing to do for the common single-chunk symbol).
// One loop serves the extend after the state update and the extends between chunk runs, so that the lanes of a warp
// meet at the same instructions whatever their fold counts are.
// The graph kernels collect at most 32 folded bits (every value of a graph with 32-bit node ids has at most 33 bits,
// i.e. folds*R <= 32; more is reported as corrupt); WIDE collects 64 (ANSDecoder::decode on arbitrary symbols).
template <bool WIDE = false, class Tab>
__device__ __forceinline__ uint64_t ans_decode_cp(const uint4 cp, const Tab& tab, Dec& d,
                                                  const uint16_t* __restrict__ stream, uint32_t& err) {
  const uint32_t slot = d.state & (cp.x & 0xFFFFu);
  const uint32_t L = (cp.x >> 16) & 31u;
  const uint2 bk = tab.bucket(cp.y + (slot >> 5));
  const uint32_t j = bk.y + (uint32_t)__popc(bk.x & ((2u << (slot & 31u)) - 1u));
  const uint2 e = tab.entry(cp, cp.z, j);
  const uint32_t folds = e.y >> 16;
  const uint32_t R = (cp.x >> 21) & 31u;
  // sentinel (slot beyond the sum of frequencies: folds = 0xFFFF) and symbols wider than the collector
  if (folds * R > (WIDE ? 48u : 32u)) { err |= ERR_CORRUPT; return 0; }
  d.state = (d.state >> L) * (e.x >> 16) + slot - (e.x & 0xFFFFu);
  using Acc = typename std::conditional<WIDE, uint64_t, uint32_t>::type;
  uint32_t rem = folds, pos = 0;
  Acc acc = 0;
  const uint32_t recip = tab.recip(cp);  // floor(65536/R)+1
  for (;;) {
    if (d.state < WGA_LOWER_BOUND && !ans_extend(d, stream)) { err |= ERR_CORRUPT; return 0; }
    if (rem == 0) break;
    const uint32_t room = 16u - (uint32_t)__clz((int)d.state);  // bit length - 16 (>= 1: state >= 2^16 here)
    const uint32_t t = min(((room + R - 1u) * recip) >> 16, rem);  // ceil(room/R) trips until state < 2^16
    const uint32_t nb = t * R;                                    // <= 16 + R - 1 bits
    acc |= (Acc)(d.state & ((1u << nb) - 1u)) << pos;
    pos += nb;
    d.state >>= nb;
    rem -= t;
  }
  if (folds > 1u) acc = reverse_groups(acc, folds, R);
  return ((uint64_t)(e.y & 0xFFFFu) << pos) | acc;
}

template <bool WIDE = false, class Tab>
__device__ __forceinline__ uint64_t ans_decode(const DevTables& tb, const Tab& tab, int c, Dec& d,
                                               const uint16_t* __restrict__ stream, uint32_t& err) {
  return ans_decode_cp<WIDE>(comp_params(tb, c), tab, d, stream, err);
}

__device__ __forceinline__ int64_t nat2int(uint64_t x) {
  return (x & 1) ? -(int64_t)((x + 1) >> 1) : (int64_t)(x >> 1);
}

}  // namespace wga
