// Device-side view of an ANS graph and the symbol decoder (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.hpp"

namespace wga {

// Packed tables as the kernels see them (pointers to global memory; kernels may stage them in smem).
struct DevTables {
  const uint16_t* lut;
  const uint2* ent;
  uint32_t lut_off[WGA_COMPONENTS];
  uint32_t ent_off[WGA_COMPONENTS];
  uint8_t L[WGA_COMPONENTS];
  uint8_t R[WGA_COMPONENTS];
  uint8_t shift[WGA_COMPONENTS];
  uint32_t lut_total;  // u16 elements
  uint32_t ent_total;  // uint2 elements
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
  ERR_WORKSPACE = 2u,     // block staging overflow
  ERR_RANGE = 4u,         // a reference leaves the decoded range (halo missing)
  ERR_SYMBOL_WIDTH = 8u,  // decoded value does not fit 32 bits
};

#define WGA_LOWER_BOUND 65536u  // INTERVAL_LOWER_BOUND, src/ans/mod.rs:21

// Per-component decode parameters packed for one 16-byte (shared-memory) load:
//   x = lut_off | L << 16 | shift << 21 | R << 26      y = ent_off     z = floor(65536/R)+1 (division by R)
__host__ __device__ inline uint4 comp_params(const DevTables& tb, int c) {
  const uint32_t R = tb.R[c] ? tb.R[c] : 1u;
  return make_uint4(tb.lut_off[c] | ((uint32_t)tb.L[c] << 16) | ((uint32_t)tb.shift[c] << 21) | (R << 26),
                    tb.ent_off[c], 65536u / R + 1u, 0u);
}

// One 16-bit extend (decoder.rs:89-93).  false = the stream is exhausted.
// StreamT: anything indexable by the word index (a plain pointer, or WindowedStream below).
template <class PtrT, class StreamT>
__device__ __forceinline__ bool ans_extend(uint32_t& state, PtrT& ptr, const StreamT stream) {
  if (ptr <= 0) return false;
  --ptr;
  state = (state << 16) | stream[ptr];
  return true;
}

// One ANS symbol.  Restates ANSDecoder::decode (src/ans/decoder.rs:58-87) on the packed tables:
//   slot  = state & (2^L-1)                                   decoder.rs:59
//   entry = owner(slot)                                       decoder.rs:60  (lut + forward walk)
//   state = (state >> L)*freq + slot - cumul                  decoder.rs:62-65
//   one conditional 16-bit extend                             decoder.rs:67-69, 89-93
//   folds x { [extend]; fold=(fold<<R)|(state&(2^R-1)); state>>=R; [extend] }   decoder.rs:74-85
//   result = (base << folds*R) | fold                         decoder.rs:86 with quasi_fold (model4decoder.rs:56-68)
// The fold loop of the reference takes R bits per trip (up to 38 trips).  Between two extends the trips
// only shift the state, so they are done here in one step per extend: with n = bit length of the state,
// the next j = ceil((n-16)/R) trips need no extend (the state stays >= 2^16 until the j-th shift), they
// consume the low j*R bits, and the chunks enter `fold` first-taken-highest, i.e. in reversed group order.
// LUT / ENT are pointers to the component-indexed packed tables (global or shared memory); cp = comp_params(c).
// Everything after the table lookup: state update, extend, folds.  e = entry that owns `slot`.
// PtrT: index of the next word below in `stream` (int64_t, or uint32_t when the resident span has < 2^32 words).
template <class PtrT, class StreamT>
__device__ __forceinline__ uint64_t ans_apply(const uint4 cp, const uint2 e, const uint32_t slot, uint32_t& state,
                                              PtrT& ptr, const StreamT stream, uint32_t& err) {
  const uint32_t L = (cp.x >> 16) & 31u;
  const uint32_t folds = e.y >> 16;
  if (folds == 0xFFFFu) {  // sentinel: slot beyond the sum of frequencies
    err |= ERR_CORRUPT;
    return 0;
  }
  state = (state >> L) * (e.x >> 16) + slot - (e.x & 0xFFFFu);
  if (state < WGA_LOWER_BOUND && !ans_extend(state, ptr, stream)) { err |= ERR_CORRUPT; return 0; }
  uint64_t sym = e.y & 0xFFFFu;
  if (folds) {
    const uint32_t R = cp.x >> 26;
    const uint32_t rmask = (1u << R) - 1u;
    uint32_t rem = folds;
    uint64_t fold = 0;
    do {
      const uint32_t n = 32u - (uint32_t)__clz((int)state);     // 17..32 (state >= 2^16 here)
      uint32_t t = ((n - 16u + R - 1u) * cp.z) >> 16;            // ceil((n-16)/R): trips until state < 2^16
      t = max(min(t, rem), 1u);  // >= 1 also on corrupt input (state < 2^16 after an extend)
      const uint32_t nb = t * R;                                 // <= 25 bits
      uint32_t bits = state & ((1u << nb) - 1u);
      state >>= nb;
      uint32_t grp;
      if (R == 1) grp = __brev(bits) >> (32u - nb);
      else {
        grp = 0;
        for (uint32_t q = 0; q < t; ++q) { grp = (grp << R) | (bits & rmask); bits >>= R; }
      }
      fold = (fold << nb) | grp;
      rem -= t;
      if (state < WGA_LOWER_BOUND && !ans_extend(state, ptr, stream)) { err |= ERR_CORRUPT; return 0; }
    } while (rem);
    sym = (sym << (folds * R)) | fold;
  }
  return sym;
}

template <class LutPtr, class EntPtr, class PtrT>
__device__ __forceinline__ uint64_t ans_decode_cp(const uint4 cp, LutPtr lut, EntPtr ent, uint32_t& state,
                                                  PtrT& ptr, const uint16_t* __restrict__ stream, uint32_t& err) {
  const uint32_t L = (cp.x >> 16) & 31u;
  const uint32_t slot = state & ((1u << L) - 1u);
  uint32_t j = lut[(cp.x & 0xFFFFu) + (slot >> ((cp.x >> 21) & 31u))];
  const uint32_t eo = cp.y;
  uint2 e = ent[eo + j];
  while (slot - (e.x & 0xFFFFu) >= (e.x >> 16)) {  // rare: several symbols share the bucket
    ++j;
    e = ent[eo + j];
  }
  return ans_apply(cp, e, slot, state, ptr, stream, err);
}

template <class LutPtr, class EntPtr>
__device__ __forceinline__ uint64_t ans_decode(const DevTables& tb, LutPtr lut, EntPtr ent, int c,
                                               uint32_t& state, int64_t& ptr,
                                               const uint16_t* __restrict__ stream, uint32_t& err) {
  return ans_decode_cp(comp_params(tb, c), lut, ent, state, ptr, stream, err);
}

__device__ __forceinline__ int64_t nat2int(uint64_t x) {
  return (x & 1) ? -(int64_t)((x + 1) >> 1) : (int64_t)(x >> 1);
}

}  // namespace wga
