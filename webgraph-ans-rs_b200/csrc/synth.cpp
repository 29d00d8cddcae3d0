// wga_synth_graph: synthetic graphs of the benchmark shapes.  The generator itself is workload
// infrastructure shared with the benchmark's reference arm: tools/synth/synth_graph.hpp.
#include "../../tools/synth/synth_graph.hpp"

#include "common.hpp"

namespace wga {
uint64_t synth_graph(int kind, uint64_t N, double mean_degree, uint64_t seed, uint64_t first, uint64_t last,
                     int threads, uint64_t* h_offsets, uint32_t* h_succ) {
  try {
    return wgsynth::synth_graph(kind, N, mean_degree, seed, first, last, threads, h_offsets, h_succ);
  } catch (const std::invalid_argument& e) {
    throw Error(WGA_E_ARG, e.what());
  }
}
}  // namespace wga
