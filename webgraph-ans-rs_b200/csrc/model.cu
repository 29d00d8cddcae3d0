// GPU model builder (stub for the first slice; implemented in model.cu proper).
#include "model.hpp"

#include "graph.hpp"

namespace wga {
struct ModelBuilder::Impl {};
ModelBuilder::ModelBuilder() : impl_(nullptr) {}
ModelBuilder::~ModelBuilder() {}
uint64_t* ModelBuilder::device_bins() { return nullptr; }
void ModelBuilder::accumulate_device(const uint8_t*, const uint64_t*, uint64_t, cudaStream_t) { throw Error(WGA_E_UNSUPPORTED, "model build: not implemented yet"); }
void ModelBuilder::accumulate_host(const uint8_t*, const uint64_t*, uint64_t) { throw Error(WGA_E_UNSUPPORTED, "model build: not implemented yet"); }
uint64_t ModelBuilder::sparse_count() { return 0; }
void ModelBuilder::sparse_export(uint8_t*, uint64_t*, uint64_t*) {}
void ModelBuilder::sparse_merge(const uint8_t*, const uint64_t*, const uint64_t*, uint64_t) {}
void ModelBuilder::build(ComponentModel*, double*, double*) { throw Error(WGA_E_UNSUPPORTED, "model build: not implemented yet"); }

uint64_t successors_workspace_size(const wga_graph*, uint64_t, uint64_t) { return 0; }
void successors_batch(wga_graph*, const uint64_t*, uint64_t, uint64_t*, uint32_t*, uint64_t, void*, uint64_t, uint64_t*, cudaStream_t) {
  throw Error(WGA_E_UNSUPPORTED, "successors_batch: not implemented yet");
}
}  // namespace wga
