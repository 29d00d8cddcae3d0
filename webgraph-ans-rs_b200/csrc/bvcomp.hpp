// Host front end of ANSBvGraph::store (src/bvgraph/random_access.rs:91-222): webgraph's BvComp
// reference selection driven by the reference's estimators, the (component, symbol) tap that feeds
// the GPU model builder, and the serial ANS encoder that produces stream + phases.
#pragma once
#include <functional>

#include "common.hpp"

namespace wga {

// (component, raw symbol) pairs in write order -- what BVGraphModelBuilder::write_* pushes into the
// model builder (src/bvgraph/writers/bvgraph_model_builder.rs:51-103) and what
// ANSBVGraphEncodeAndEstimate spills (src/bvgraph/writers/bvgraph_encoder.rs:103-156).
struct SymbolStream {
  std::vector<uint8_t> comps;
  std::vector<uint64_t> vals;
  void push(int c, uint64_t v) {
    comps.push_back((uint8_t)c);
    vals.push_back(v);
  }
  size_t size() const { return vals.size(); }
};

// Cost model used by BvComp to pick references: Log2Estimator (log2_estimator.rs:15-49) or
// EntropyEstimator (entropy_estimator.rs:33-113).
class Estimator {
 public:
  Estimator() : log2_(true) {}                                  // Log2Estimator
  explicit Estimator(const ComponentModel tables[WGA_COMPONENTS]);  // EntropyEstimator::new(model, model.get_folding_params())
  // raw view for the GPU candidate costing (bvcomp_gpu.cu)
  bool is_log2() const { return log2_; }
  const std::vector<uint32_t>& table(int c) const { return table_[c]; }
  uint64_t threshold(int c) const { return thr_[c]; }
  uint64_t offset(int c) const { return off_[c]; }
  unsigned fidelity(int c) const { return fid_[c]; }
  unsigned radix(int c) const { return rad_[c]; }
  uint64_t cost(int c, uint64_t value) const {
    if (log2_) return 63u - (unsigned)__builtin_clzll(value + 2);
    uint64_t sym = value;
    if (value >= thr_[c]) {
      unsigned bits = 64u - (unsigned)__builtin_clzll(value);
      uint64_t cuts = (bits - fid_[c]) / rad_[c];
      sym = (value >> (cuts * rad_[c])) + off_[c] * cuts;
    }
    if (sym >= table_[c].size()) throw Error(WGA_E_ARG, "symbol exceeds 48 bits");
    return table_[c][sym];
  }

 private:
  bool log2_;
  std::vector<uint32_t> table_[WGA_COMPONENTS];
  uint64_t thr_[WGA_COMPONENTS], off_[WGA_COMPONENTS];
  unsigned fid_[WGA_COMPONENTS], rad_[WGA_COMPONENTS];
};

struct BvCompParams {
  uint64_t window = 7, max_ref_count = 3, min_interval_length = 4;
};

// Node source: fills `out` with the ascending successors of node v.
using NodeSource = std::function<void(uint64_t v, std::vector<uint64_t>& out)>;

// Runs BvComp over nodes [first,last) with start_node = first (webgraph's parallel-compression
// semantics: no reference crosses `first`) and appends the chosen records' symbols to `out`.
// `choice` (optional): the reference offset already chosen for every node, choice[v - choice_first] (from the GPU
// candidate costing); the estimator is then not consulted.
uint64_t bvcomp_range(const NodeSource& src, uint64_t first, uint64_t last, const BvCompParams& p,
                      const Estimator& est, SymbolStream& out, const uint16_t* choice = nullptr, uint64_t choice_first = 0);

// Whole graph: chunk_nodes == 0 -> one sequential BvComp (== the reference's pass); otherwise
// independent chunks on `threads` host threads, concatenated in node order.
uint64_t bvcomp_graph(const NodeSource& src, uint64_t n_nodes, const BvCompParams& p, const Estimator& est,
                      uint64_t chunk_nodes, int threads, SymbolStream& out);
// Nodes [first,last) of the graph only (one rank's share of a sharded model build): with chunk_nodes > 0 the
// chunks are those of the whole-graph run, so the symbols equal its symbols for these nodes.
uint64_t bvcomp_nodes(const NodeSource& src, uint64_t first, uint64_t last, const BvCompParams& p, const Estimator& est,
                      uint64_t chunk_nodes, int threads, SymbolStream& out, const uint16_t* choice = nullptr);

// GPU candidate costing: for every node v of [first, first + n) and every reference offset
// 0..window the EntropyEstimator / Log2Estimator cost of v's record against node v - offset, one device thread per
// (node, offset) pair, then the reference selection of BvComp (nearest candidate wins ties, reference chains bounded by
// max_ref_count) per chunk.  CSR in host memory (uploaded).  choice[v - first] = chosen offset; costs (optional,
// tests): n * (window + 1) values, UINT64_MAX where there is no candidate.
void bvcomp_choose_gpu(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first, uint64_t n, const BvCompParams& p,
                       const Estimator& est, uint64_t chunk_nodes, std::vector<uint16_t>& choice,
                       std::vector<uint64_t>* costs);
// the same table computed on the host (the costs BvComp computes one by one), for the parity tests
void bvcomp_costs_host(const NodeSource& src, uint64_t first, uint64_t n, const BvCompParams& p, const Estimator& est,
                       uint64_t chunk_nodes, std::vector<uint64_t>& costs);

struct EncodeResult {
  std::vector<uint16_t> stream;
  uint32_t state = 1u << 16;
  Phases phases;  // one per Outdegree symbol, in push order (node N-1 first)
};
// ANSEncoder::encode over the symbols in reverse order, phase after every Outdegree
// (src/ans/encoder.rs:39-103, bvgraph_encoder.rs:159-174).
void ans_encode(const ComponentModel tables[WGA_COMPONENTS], const uint8_t* comps, const uint64_t* vals,
                uint64_t n, EncodeResult& out);

}  // namespace wga
