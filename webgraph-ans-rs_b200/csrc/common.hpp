// Shared host-side declarations of libwgans (B200-native webgraph-ans hot path).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/wga.h"

namespace wga {

// ---- error plumbing --------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int code, const std::string& m) : std::runtime_error(m), code(code) {}
};
void set_last_error(const std::string& m);
// Runs f, maps exceptions to WGA_E_* codes + thread-local message.
template <class F>
int guarded(F&& f) {
  try {
    f();
    return WGA_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return WGA_E_ARG;
  }
}

// ---- host model of the files (src/ans/mod.rs:31-54, component_model4encoder.rs:37-57) ---------------
struct ComponentModel {
  std::vector<wga_encoder_entry> table;
  uint64_t frame_size = 0;  // log2
  uint64_t radix = 2;
  uint64_t fidelity = 2;
  uint64_t folding_threshold = 10;
  uint64_t folding_offset = 10;
};

struct Prelude {
  ComponentModel tables[WGA_COMPONENTS];
  std::vector<uint16_t> stream;
  uint32_t state = 1u << 16;
  uint64_t number_of_nodes = 0;
  uint64_t compression_window = 0;
  uint64_t min_interval_length = 0;
  uint64_t number_of_arcs = 0;
};

struct Phases {                    // entry i belongs to node N-1-i (random_access.rs:202,225-231)
  std::vector<uint32_t> states;    // .states
  std::vector<uint64_t> pointers;  // .pointers, expanded
};

enum Component : int {
  Outdegree = 0, ReferenceOffset, BlockCount, Blocks, IntervalCount,
  IntervalStart, IntervalLen, FirstResidual, Residual
};

// ---- packed decoder tables (host build; device copies live in the graph handle) ---------------------
// The reference expands every component to 2^frame 16-byte entries (model4decoder.rs:18-54; up to 1 MiB
// per component, L2-resident at best).  Slots are uniformly distributed, so that table cannot be cached.
// We keep instead, per component:
//   bkt[max(1, 2^L / 32)] : per bucket of 32 slots {mask, j0}: j0 = index of the non-zero symbol that owns the
//                           first slot of the bucket, bit s of mask (s >= 1) set iff a symbol's slot range starts
//                           at slot 32*bucket + s.  Owner of a slot = j0 + popc(mask & ((2 << s) - 1)): exact, no walk.
//   ent[nnz+1]            : {cumul | freq<<16 , base | folds<<16}; last entry is a sentinel that owns the unused
//                           slots beyond the sum of frequencies (folds == 0xFFFF)
// <= 16 KB of buckets per component + 8 B per symbol: shared-memory resident in the entropy kernel.
struct Ent {
  uint32_t cf;  // cumul_freq | freq << 16
  uint32_t bf;  // base (symbol - offset*folds) | folds << 16 ; folds == 0xFFFF marks the sentinel
};
struct Bkt {
  uint32_t mask;
  uint32_t j0;
};

struct PackedTablesData {
  std::vector<Bkt> bkt;
  std::vector<Ent> ent;
  uint32_t bkt_off[WGA_COMPONENTS];
  uint32_t ent_off[WGA_COMPONENTS];
  uint32_t nb[WGA_COMPONENTS];
  uint32_t nent[WGA_COMPONENTS];  // including the sentinel
  uint32_t nnz[WGA_COMPONENTS];
  uint8_t L[WGA_COMPONENTS];
  uint8_t R[WGA_COMPONENTS];
};
PackedTablesData pack_tables(const ComponentModel tables[WGA_COMPONENTS]);

}  // namespace wga
