// GPU model builder: device-resident symbol histograms + frame normalisation.
// Replaces ANSModel4EncoderBuilder (src/ans/model4encoder_builder.rs:39-327).
#pragma once
#include <cuda_runtime.h>

#include <map>

#include "common.hpp"

namespace wga {

class ModelBuilder {
 public:
  ModelBuilder();
  ~ModelBuilder();
  ModelBuilder(const ModelBuilder&) = delete;
  ModelBuilder& operator=(const ModelBuilder&) = delete;

  uint64_t* device_bins();
  void accumulate_device(const uint8_t* d_comps, const uint64_t* d_syms, uint64_t n, cudaStream_t st);
  void accumulate_host(const uint8_t* h_comps, const uint64_t* h_syms, uint64_t n);
  uint64_t sparse_count();
  void sparse_export(uint8_t* h_comps, uint64_t* h_syms, uint64_t* h_counts);
  void sparse_merge(const uint8_t* h_comps, const uint64_t* h_syms, const uint64_t* h_counts, uint64_t n);
  void build(ComponentModel out[WGA_COMPONENTS], double* h_original_cost9, double* h_final_cost9);

  ComponentModel result[WGA_COMPONENTS];

 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace wga
