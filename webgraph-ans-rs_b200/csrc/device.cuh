// Device-side view of an ANS graph and the symbol decoder (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

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
struct Dec {
  uint32_t state;
  uint32_t sp;  // words of the resident span below the record's read position
  uint32_t w;   // stream[sp-1] (0 when sp == 0)
};

__device__ __forceinline__ void dec_prime(Dec& d, const uint16_t* __restrict__ stream) {
  d.w = d.sp ? (uint32_t)__ldg(stream + d.sp - 1) : 0u;
}

// One 16-bit extend (decoder.rs:89-93).  false = the stream is exhausted.
__device__ __forceinline__ bool ans_extend(Dec& d, const uint16_t* __restrict__ stream) {
  if (d.sp == 0) return false;
  d.state = (d.state << 16) | d.w;
  --d.sp;
  d.w = d.sp ? (uint32_t)__ldg(stream + d.sp - 1) : 0u;
  return true;
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
// t = ceil((n-16)/R) trips need no extend (the state stays >= 2^16 until the t-th shift), they consume the low t*R
// bits, and the chunks enter `fold` first-taken-highest, i.e. in reversed group order (radix 1: a bit reversal).
// One loop for every folded symbol (usually one or two passes), so that the folded lanes of a warp stay together.
template <class Tab>
__device__ __forceinline__ uint64_t ans_decode_cp(const uint4 cp, const Tab& tab, Dec& d,
                                                  const uint16_t* __restrict__ stream, uint32_t& err) {
  const uint32_t slot = d.state & (cp.x & 0xFFFFu);
  const uint32_t L = (cp.x >> 16) & 31u;
  const uint2 bk = tab.bucket(cp.y + (slot >> 5));
  const uint32_t j = bk.y + (uint32_t)__popc(bk.x & ((2u << (slot & 31u)) - 1u));
  const uint2 e = tab.entry(cp, cp.z, j);
  uint32_t rem = e.y >> 16;  // folds
  if (rem == 0xFFFFu) {  // sentinel: slot beyond the sum of frequencies
    err |= ERR_CORRUPT;
    return 0;
  }
  d.state = (d.state >> L) * (e.x >> 16) + slot - (e.x & 0xFFFFu);
  if (d.state < WGA_LOWER_BOUND && !ans_extend(d, stream)) { err |= ERR_CORRUPT; return 0; }
  uint64_t val = e.y & 0xFFFFu;
  if (rem) {
    const uint32_t R = (cp.x >> 21) & 31u;
    const uint32_t recip = tab.recip(cp);  // floor(65536/R)+1
    const uint32_t rmask = (1u << R) - 1u;
    do {
      const uint32_t room = 16u - (uint32_t)__clz((int)d.state);  // bit length - 16 (>= 1 while state >= 2^16)
      uint32_t t = ((room + R - 1u) * recip) >> 16;                // ceil(room/R): trips until state < 2^16
      t = max(min(t, rem), 1u);  // >= 1 also on corrupt input (state < 2^16 after a failed extend)
      const uint32_t nb = t * R;                                   // <= 31 bits
      uint32_t bits = d.state & ((1u << nb) - 1u);
      d.state >>= nb;
      uint32_t grp;
      if (R == 1) grp = __brev(bits) >> (32u - nb);
      else {
        grp = 0;
        for (uint32_t q = t; q; --q) { grp = (grp << R) | (bits & rmask); bits >>= R; }
      }
      val = (val << nb) | grp;
      rem -= t;
      if (d.state < WGA_LOWER_BOUND && !ans_extend(d, stream)) { err |= ERR_CORRUPT; return 0; }
    } while (rem);
  }
  return val;
}

template <class Tab>
__device__ __forceinline__ uint64_t ans_decode(const DevTables& tb, const Tab& tab, int c, Dec& d,
                                               const uint16_t* __restrict__ stream, uint32_t& err) {
  return ans_decode_cp(comp_params(tb, c), tab, d, stream, err);
}

__device__ __forceinline__ int64_t nat2int(uint64_t x) {
  return (x & 1) ? -(int64_t)((x + 1) >> 1) : (int64_t)(x >> 1);
}

}  // namespace wga
