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

// Per-component decode parameters packed for one 8-byte (shared-memory) load:
//   x = lut_off | L << 16 | shift << 21 | R << 26      y = ent_off
__host__ __device__ inline uint2 comp_params(const DevTables& tb, int c) {
  return make_uint2(tb.lut_off[c] | ((uint32_t)tb.L[c] << 16) | ((uint32_t)tb.shift[c] << 21) | ((uint32_t)tb.R[c] << 26),
                    tb.ent_off[c]);
}

// One ANS symbol.  Restates ANSDecoder::decode (src/ans/decoder.rs:58-87) on the packed tables:
//   slot  = state & (2^L-1)                                   decoder.rs:59
//   entry = owner(slot)                                       decoder.rs:60  (lut + forward walk)
//   state = (state >> L)*freq + slot - cumul                  decoder.rs:62-65
//   one conditional 16-bit extend                             decoder.rs:67-69, 89-93
//   folds x { [extend]; fold=(fold<<R)|(state&(2^R-1)); state>>=R; [extend] }   decoder.rs:74-85
//   result = (base << folds*R) | fold                         decoder.rs:86 with quasi_fold (model4decoder.rs:56-68)
// LUT / ENT are pointers to the component-indexed packed tables (global or shared memory); cp = comp_params(c).
template <class LutPtr, class EntPtr>
__device__ __forceinline__ uint64_t ans_decode_cp(const uint2 cp, LutPtr lut, EntPtr ent, uint32_t& state,
                                                  int64_t& ptr, const uint16_t* __restrict__ stream, uint32_t& err) {
  const uint32_t L = (cp.x >> 16) & 31u;
  const uint32_t slot = state & ((1u << L) - 1u);
  uint32_t j = lut[(cp.x & 0xFFFFu) + (slot >> ((cp.x >> 21) & 31u))];
  const uint32_t eo = cp.y;
  uint2 e = ent[eo + j];
  while (slot - (e.x & 0xFFFFu) >= (e.x >> 16)) {  // rare: several symbols share the bucket
    ++j;
    e = ent[eo + j];
  }
  uint32_t folds = e.y >> 16;
  if (folds == 0xFFFFu) {  // sentinel: slot beyond the sum of frequencies
    err |= ERR_CORRUPT;
    return 0;
  }
  state = (state >> L) * (e.x >> 16) + slot - (e.x & 0xFFFFu);
  if (state < WGA_LOWER_BOUND) {
    if (ptr <= 0) { err |= ERR_CORRUPT; return 0; }
    --ptr;
    state = (state << 16) | stream[ptr];
  }
  uint64_t sym = e.y & 0xFFFFu;
  if (folds) {
    const uint32_t R = cp.x >> 26;
    const uint32_t rmask = (1u << R) - 1u;
    uint64_t fold = 0;
    for (uint32_t i = 0; i < folds; ++i) {
      if (state < WGA_LOWER_BOUND) {
        if (ptr <= 0) { err |= ERR_CORRUPT; return 0; }
        --ptr;
        state = (state << 16) | stream[ptr];
      }
      fold = (fold << R) | (uint64_t)(state & rmask);
      state >>= R;
      if (state < WGA_LOWER_BOUND) {
        if (ptr <= 0) { err |= ERR_CORRUPT; return 0; }
        --ptr;
        state = (state << 16) | stream[ptr];
      }
    }
    sym = (sym << (folds * R)) | fold;
  }
  return sym;
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
