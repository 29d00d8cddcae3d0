// =============================================================================
//  wgo.hpp -- CPU ORACLE for the webgraph-ans hot path.   TEST INFRASTRUCTURE.
// =============================================================================
//  This header is a plain, single-threaded C++17 restatement of the reference's
//  (ciminilorenzo/webgraph-ans-rs, Rust) algorithms for the ANS decode of BvGraph
//  components and the encoder-side symbol-model construction.  It exists ONLY to
//  check the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline
//  / --impl reference legs).  Nothing under webgraph-ans-rs_b200/ includes,
//  links or executes it.
//
//  Parity status: the reference cannot be compiled here (no rustc/cargo, no
//  vendored crates) and ships no golden .ans/.states/.pointers files nor
//  known-answer vectors for tables or streams, so "the exact bytes the reference
//  writes" are PARITY UNPINNED.  What pins this oracle:
//    * the reference's own tests are round-trip properties (tests/compressor_tests.rs,
//      tests/test_bvgraph.rs); every one of them is restated in tests/ against
//      this oracle;
//    * the golden BV graph tests/data/cnr-2000 (.graph/.properties/.ef) pins the
//      BV record order and the epserde/Elias-Fano byte layout;
//    * SURVEY.md section 8a regression anchors.
//  Tie-break rules where the reference is itself nondeterministic (HashMap
//  iteration order, unstable sort): costs are summed in ascending raw-symbol
//  order; symbols are sorted by (frequency, index) ascending.
//
//  Every function cites the reference file:line it follows
//  (paths relative to the reference repository root).
// =============================================================================
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace wgo {

// ----------------------------------------------------------------- src/lib.rs:11-24
using Symbol = uint16_t;
using RawSymbol = uint64_t;
using State = uint32_t;
using Freq = uint16_t;
constexpr RawSymbol MAX_RAW_SYMBOL = (1ull << 48) - 1;

// ----------------------------------------------------------------- src/ans/mod.rs:18-28
constexpr unsigned B = 16;
constexpr State INTERVAL_LOWER_BOUND = 1u << 16;
constexpr uint32_t NORMALIZATION_MASK = 0xFFFF;
constexpr size_t MAXIMUM_FRAME_SIZE = 1u << 16;

// ----------------------------------------------------------------- src/bvgraph/mod.rs:12-27
enum Component : int {
  Outdegree = 0, ReferenceOffset, BlockCount, Blocks, IntervalCount,
  IntervalStart, IntervalLen, FirstResidual, Residual
};
constexpr int COMPONENTS = 9;

inline unsigned ilog2(uint64_t x) {
  if (x == 0) throw std::domain_error("ilog2(0)");  // Rust panics
  return 63u - (unsigned)__builtin_clzll(x);
}

// ----------------------------------------------------------------- src/utils/ans_utils.rs:4-12
inline Symbol fold_without_streaming_out(RawSymbol sym, size_t radix, size_t fidelity) {
  size_t cuts = ((size_t)ilog2(sym) + 1 - fidelity) / radix;
  size_t bit_to_cut = cuts * radix;
  sym >>= bit_to_cut;
  RawSymbol offset = (((1ull << radix) - 1) * (1ull << (fidelity - 1))) * (RawSymbol)cuts;
  RawSymbol r = sym + offset;
  if (r > 0xFFFF) throw std::overflow_error("Folded symbol is bigger than u16::MAX");
  return (Symbol)r;
}

// ----------------------------------------------------------------- src/utils/data_utils.rs:15-39
// Returns false where the reference bails ("Too many symbols have frequency lower than 1").
// The f64 expression is evaluated exactly as written: left to right, no FMA contraction
// (the oracle is compiled with -ffp-contract=off).
inline bool scale_freqs(const std::vector<size_t>& freqs, const std::vector<size_t>& sorted_indices,
                        size_t n, size_t m, int64_t new_m, std::vector<size_t>& approx) {
  approx = freqs;
  const double ratio = (double)new_m / (double)m;
  for (size_t index = 0; index < sorted_indices.size(); ++index) {
    size_t sym_index = sorted_indices[index];
    size_t sym_freq = freqs[sym_index];
    double second_ratio = (double)new_m / (double)m;
    double scale = (double)(n - index) * ratio / (double)n + (double)index * second_ratio / (double)n;
    double v = std::floor(0.5 + scale * (double)sym_freq);
    size_t a = (size_t)v;            // Rust `as usize` saturates; v >= 0 and small here
    if (a < 1) a = 1;
    approx[sym_index] = a;
    new_m -= (int64_t)a;
    m -= sym_freq;
    if (new_m < 0) return false;
  }
  return true;
}

// ----------------------------------------------------------------- src/ans/models/component_model4encoder.rs:11-70
struct EncoderModelEntry {  // repr(C), 8 bytes
  uint32_t upperbound;
  Freq cumul_freq;
  Freq freq;
};
static_assert(sizeof(EncoderModelEntry) == 8, "layout");

inline EncoderModelEntry make_encoder_entry(uint16_t freq, size_t k, Freq cumul) {  // :28-34
  EncoderModelEntry e;
  e.freq = freq;
  e.upperbound = (uint32_t)((1u << (k + B)) * (uint32_t)freq);  // u32, wrapping in release
  e.cumul_freq = cumul;
  return e;
}

struct ComponentModel4Encoder {
  std::vector<EncoderModelEntry> table;
  size_t frame_size = 0;  // log2 of the frame
  size_t radix = 2;
  size_t fidelity = 2;
  uint64_t folding_threshold = 10;  // Default impl :59-70
  uint64_t folding_offset = 10;
};

using Model4Encoder = std::array<ComponentModel4Encoder, COMPONENTS>;

// ----------------------------------------------------------------- src/ans/model4encoder_builder.rs:28-37
static const std::pair<size_t, size_t> PARAMS_COMBINATIONS[] = {  // (fidelity, radix)
    {1, 3}, {2, 2}, {3, 1},
    {1, 4}, {2, 3}, {3, 2}, {4, 1},
    {1, 5}, {2, 4}, {3, 3}, {4, 2}, {5, 1},
    {1, 6}, {2, 5}, {3, 4}, {4, 3}, {5, 2}, {6, 1},
    {1, 7}, {2, 6}, {3, 5}, {4, 4}, {5, 3}, {6, 2}, {7, 1},
    {1, 8}, {2, 7}, {3, 6}, {4, 5}, {5, 4}, {6, 3}, {7, 2}, {8, 1},
    {1, 9}, {2, 8}, {3, 7}, {4, 6}, {5, 5}, {6, 4}, {7, 3}, {8, 2}, {9, 1},
    {1, 10}, {2, 9}, {3, 8}, {4, 7}, {5, 6}, {6, 5}, {7, 4}, {8, 3}, {9, 2}, {10, 1}};
constexpr size_t N_PARAMS = sizeof(PARAMS_COMBINATIONS) / sizeof(PARAMS_COMBINATIONS[0]);
constexpr double THETA = 1.0001;  // :23

struct ModelBuildInfo {  // what the reference prints with log::info! (:237-266)
  double original_cost[COMPONENTS];
  double final_cost[COMPONENTS];
};

// ----------------------------------------------------------------- src/ans/model4encoder_builder.rs:39-327
class ANSModel4EncoderBuilder {
 public:
  // :67-78 ; returns false where the reference returns Err
  bool push_symbol(RawSymbol symbol, int component) {
    if (symbol > MAX_RAW_SYMBOL) return false;
    total_freqs[component] += 1;
    real_freqs[component][symbol] += 1;
    return true;
  }
  void push_symbol_count(RawSymbol symbol, int component, size_t count) {
    total_freqs[component] += count;
    real_freqs[component][symbol] += count;
  }

  // :275-289 ; summed in ascending raw-symbol order (oracle tie-break rule)
  std::array<double, COMPONENTS> calculate_cost() const {
    std::array<double, COMPONENTS> out{};
    for (int c = 0; c < COMPONENTS; ++c) {
      double s = 0.0;
      for (auto& kv : real_freqs[c]) {
        double prob = (double)kv.second / (double)total_freqs[c];
        s += -std::log2(prob) * (double)kv.second;
      }
      out[c] = s;
    }
    return out;
  }

  // :297-327
  static double calculate_approx_folded_distribution_cost(const std::vector<size_t>& folded_distr,
                                                          const std::vector<size_t>& approx,
                                                          double new_frame_size, size_t fidelity,
                                                          size_t radix) {
    uint64_t folding_threshold = 1ull << (fidelity + radix - 1);
    uint64_t folding_offset = ((1ull << radix) - 1) * (1ull << (fidelity - 1));
    double information_content = 0.0;
    for (size_t symbol = 0; symbol < approx.size(); ++symbol) {
      if (approx[symbol] == 0) continue;
      double freq = (double)folded_distr[symbol];
      double folds = symbol < folding_threshold
                         ? 0.0
                         : (double)((symbol - folding_threshold) / folding_offset + 1);
      double prob = (double)approx[symbol] / new_frame_size;
      information_content += (-std::log2(prob) + (folds * (double)radix)) * freq;
    }
    return information_content;
  }

  // :80-271
  Model4Encoder build(ModelBuildInfo* info = nullptr) const {
    auto original_comp_costs = calculate_cost();
    double original_graph_cost = 0.0;
    for (double c : original_comp_costs) original_graph_cost += c;
    Model4Encoder models;
    std::array<double, COMPONENTS> components_final_cost{};

    for (int component = 0; component < COMPONENTS; ++component) {
      if (real_freqs[component].empty()) {  // :93-97
        models[component] = ComponentModel4Encoder();
        components_final_cost[component] = 0.0;
        continue;
      }
      std::vector<size_t> scaled_distribution;
      size_t fidelity = 0, radix = 0;
      size_t frame_size = SIZE_MAX;
      double lowest_cost = std::numeric_limits<double>::max();

      for (size_t pi = 0; pi < N_PARAMS; ++pi) {
        size_t fid = PARAMS_COMBINATIONS[pi].first, rad = PARAMS_COMBINATIONS[pi].second;
        Symbol max_bucket = fold_without_streaming_out(MAX_RAW_SYMBOL, rad, fid);
        uint64_t folding_threshold = 1ull << (fid + rad - 1);
        // the reference allocates max_bucket slots; +1 here so the (unreachable in practice)
        // raw symbol 2^48-1 does not index out of bounds
        std::vector<size_t> folded_sym_freqs((size_t)max_bucket + 1, 0);
        Symbol biggest_symbol = 0;
        for (auto& kv : real_freqs[component]) {  // :112-120
          Symbol folded = kv.first < folding_threshold ? (Symbol)kv.first
                                                      : fold_without_streaming_out(kv.first, rad, fid);
          folded_sym_freqs[folded] += kv.second;
          biggest_symbol = std::max(biggest_symbol, folded);
        }
        folded_sym_freqs.resize(max_bucket);  // reference length (:108)
        size_t n = 0;
        for (size_t f : folded_sym_freqs) n += f > 0;
        size_t m = 1;
        while (m < n) m <<= 1;  // next_power_of_two (:125-128)

        std::vector<size_t> sorted_indexes;  // :132-138 ; stable => ties by ascending index
        for (size_t i = 0; i < folded_sym_freqs.size(); ++i)
          if (folded_sym_freqs[i] > 0) sorted_indexes.push_back(i);
        std::stable_sort(sorted_indexes.begin(), sorted_indexes.end(), [&](size_t a, size_t b) {
          return folded_sym_freqs[a] < folded_sym_freqs[b];
        });

        while (true) {  // :140-206
          if (m > MAXIMUM_FRAME_SIZE) break;
          std::vector<size_t> new_distribution;
          bool ok = scale_freqs(folded_sym_freqs, sorted_indexes, n, total_freqs[component],
                                (int64_t)m, new_distribution);
          if (ok) {
            double new_cost = calculate_approx_folded_distribution_cost(
                folded_sym_freqs, new_distribution, (double)m, fid, rad);
            double difference = new_cost - original_comp_costs[component];
            double ratio = (original_graph_cost + difference) / original_graph_cost;
            if (ratio <= THETA) {
              if (m < frame_size) {
                lowest_cost = new_cost;
                new_distribution.resize(std::min(new_distribution.size(), (size_t)biggest_symbol + 1));
                scaled_distribution = new_distribution;
                frame_size = m;
                fidelity = fid;
                radix = rad;
              }
            } else if (m == MAXIMUM_FRAME_SIZE) {
              if (new_cost >= lowest_cost) break;
              lowest_cost = new_cost;
              new_distribution.resize(std::min(new_distribution.size(), (size_t)biggest_symbol + 1));
              scaled_distribution = new_distribution;
              frame_size = m;
              fidelity = fid;
              radix = rad;
              break;
            }
            m *= 2;
          } else {
            m *= 2;
          }
        }
      }
      if (frame_size == SIZE_MAX)  // :209-212
        throw std::runtime_error("no approximated distribution with a frame size <= 2^16 for component " +
                                 std::to_string(component));
      components_final_cost[component] = lowest_cost;

      ComponentModel4Encoder cm;  // :216-234
      size_t log_m = ilog2(frame_size);
      size_t k = log_m > 0 ? 16 - log_m : 15;
      uint16_t last_covered_freq = 0;
      for (size_t freq : scaled_distribution) {
        cm.table.push_back(make_encoder_entry((uint16_t)freq, k, last_covered_freq));
        uint32_t s = (uint32_t)last_covered_freq + (uint32_t)(uint16_t)freq;
        last_covered_freq = s > 0xFFFF ? 0 : (uint16_t)s;  // checked_add(..).unwrap_or(0)
      }
      cm.fidelity = fidelity;
      cm.radix = radix;
      cm.folding_threshold = 1ull << (fidelity + radix - 1);
      cm.folding_offset = ((1ull << radix) - 1) * (1ull << (fidelity - 1));
      cm.frame_size = log_m;
      models[component] = std::move(cm);
    }
    if (info) {
      for (int c = 0; c < COMPONENTS; ++c) {
        info->original_cost[c] = original_comp_costs[c];
        info->final_cost[c] = components_final_cost[c];
      }
    }
    return models;
  }

  std::array<std::map<RawSymbol, size_t>, COMPONENTS> real_freqs;  // HashMap in the reference (:41)
  std::array<size_t, COMPONENTS> total_freqs{};
};

// ----------------------------------------------------------------- src/ans/models/component_model4decoder.rs:8-22
struct DecoderModelEntry {  // repr(C) 16 bytes
  Freq freq = 0;
  Freq cumul_freq = 0;
  uint64_t quasi_folded = 0;
};
static_assert(sizeof(DecoderModelEntry) == 16, "layout");

struct ComponentModel4Decoder {
  std::vector<DecoderModelEntry> table;
  size_t frame_size = 0, radix = 0, fidelity = 0;
};

// ----------------------------------------------------------------- src/ans/models/model4decoder.rs:18-68
struct Model4Decoder {
  static constexpr uint64_t BIT_RESERVED_FOR_SYMBOL = 48;
  std::array<ComponentModel4Decoder, COMPONENTS> tables;

  static uint64_t quasi_fold(Symbol sym, uint64_t folding_offset, uint64_t folding_threshold, size_t radix) {
    if (sym < (Symbol)folding_threshold) return sym;  // :57 (threshold cast to u16)
    uint64_t symbol = sym;
    uint64_t folds = (symbol - folding_threshold) / folding_offset + 1;
    uint64_t folds_bits = folds << BIT_RESERVED_FOR_SYMBOL;
    symbol -= folding_offset * folds;
    symbol <<= folds * (uint64_t)radix;
    return symbol | folds_bits;
  }

  explicit Model4Decoder(const Model4Encoder& enc) {  // :18-54
    for (int c = 0; c < COMPONENTS; ++c) {
      const auto& t = enc[c];
      std::vector<DecoderModelEntry> vec((size_t)1 << t.frame_size);
      uint32_t last_slot = 0;
      for (size_t sym = 0; sym < t.table.size(); ++sym) {
        const auto& se = t.table[sym];
        if (se.freq == 0) continue;
        for (uint32_t slot = last_slot; slot < last_slot + se.freq; ++slot) {
          DecoderModelEntry& d = vec.at(slot);  // .get_mut(slot).unwrap()
          d.freq = se.freq;
          d.cumul_freq = se.cumul_freq;
          d.quasi_folded = quasi_fold((Symbol)sym, t.folding_offset, t.folding_threshold, t.radix);
        }
        last_slot += se.freq;
      }
      tables[c].table = std::move(vec);
      tables[c].frame_size = t.frame_size;
      tables[c].radix = t.radix;
      tables[c].fidelity = t.fidelity;
    }
  }
};

// ----------------------------------------------------------------- src/ans/encoder.rs:22-103
struct ANSCompressorPhase {  // src/ans/mod.rs:62-68
  State state;
  size_t stream_pointer;
};

class ANSEncoder {
 public:
  explicit ANSEncoder(const Model4Encoder& m) : model(m), state(INTERVAL_LOWER_BOUND) {}

  void encode(RawSymbol symbol, int component) {  // :39-78
    const auto& cm = model[component];
    if (symbol >= cm.folding_threshold) {
      size_t folds = ((size_t)ilog2(symbol) + 1 - cm.fidelity) / cm.radix;  // :31-34
      for (size_t i = 0; i < folds; ++i) {
        State bits_to_push = (State)(symbol & ((1ull << cm.radix) - 1));
        if ((unsigned)__builtin_clz(state) >= (unsigned)cm.radix) {  // state != 0 always
          state <<= cm.radix;
          state += bits_to_push;
        } else {
          shrink_state();
          state <<= cm.radix;
          state += bits_to_push;
        }
        symbol >>= cm.radix;
      }
      symbol += cm.folding_offset * (RawSymbol)folds;
    }
    const EncoderModelEntry& sd = cm.table.at((Symbol)symbol);  // Index<Symbol> panics when OOB
    if (state >= sd.upperbound) shrink_state();
    // the reference divides by NonZeroU32::unwrap_unchecked(freq) (:72-73): freq == 0 is UB there
    // (e.g. a zero-entropy input makes `ratio` NaN and the single symbol gets freq 65536 as u16 == 0)
    if (sd.freq == 0) throw std::domain_error("ANSEncoder: symbol with frequency 0 (undefined behaviour in the reference)");
    State block = state / (State)sd.freq;
    state = (block << cm.frame_size) + (State)sd.cumul_freq + (state - block * (State)sd.freq);
  }

  ANSCompressorPhase get_current_compressor_phase() const { return {state, stream.size()}; }  // :97-102

  const Model4Encoder& model;
  std::vector<uint16_t> stream;
  State state;

 private:
  void shrink_state() {  // :81-86
    stream.push_back((uint16_t)(state & NORMALIZATION_MASK));
    state >>= B;
  }
};

// ----------------------------------------------------------------- src/ans/decoder.rs:9-100
class ANSDecoder {
 public:
  ANSDecoder(const Model4Decoder& m, const std::vector<uint16_t>& s, State st)  // :27-34
      : model(&m), stream(&s), state(st), stream_pointer(s.size()) {}
  ANSDecoder(const Model4Decoder& m, const std::vector<uint16_t>& s, size_t ptr, State st)  // :41-53
      : model(&m), stream(&s), state(st), stream_pointer(ptr) {}

  RawSymbol decode(int component) {  // :58-87
    const ComponentModel4Decoder& t = model->tables[component];
    State frame_mask = (State)(((uint64_t)1 << t.frame_size) - 1);
    State slot = state & frame_mask;
    const DecoderModelEntry& e = t.table.at((Symbol)slot);
    state = (state >> t.frame_size) * (State)e.freq + slot - (State)e.cumul_freq;
    if (state < INTERVAL_LOWER_BOUND) extend_state();
    uint64_t quasi_unfolded = e.quasi_folded & ((1ull << 48) - 1);  // :96-100
    uint32_t folds = (uint32_t)(e.quasi_folded >> 48);
    uint64_t fold = 0;
    for (uint32_t i = 0; i < folds; ++i) {
      if (state < INTERVAL_LOWER_BOUND) extend_state();
      fold = (fold << t.radix) | ((uint64_t)state & ((1ull << t.radix) - 1));
      state >>= t.radix;
      if (state < INTERVAL_LOWER_BOUND) extend_state();
    }
    return quasi_unfolded | fold;
  }

  const Model4Decoder* model;
  const std::vector<uint16_t>* stream;
  State state;
  size_t stream_pointer;

 private:
  void extend_state() {  // :89-93
    stream_pointer -= 1;
    uint16_t bits = stream->at(stream_pointer);
    state = (state << B) | (State)bits;
  }
};

// ----------------------------------------------------------------- src/ans/mod.rs:31-54 (Prelude) + .states/.pointers
struct ANSGraph {
  Model4Encoder tables;
  std::vector<uint16_t> stream;
  State state = INTERVAL_LOWER_BOUND;
  size_t number_of_nodes = 0;
  size_t compression_window = 0;
  size_t min_interval_length = 0;
  uint64_t number_of_arcs = 0;
  // random access side; entry i belongs to node N-1-i (src/bvgraph/random_access.rs:202,225-231)
  std::vector<State> states;
  std::vector<uint64_t> pointers;
};

inline int64_t nat2int(uint64_t x) { return (x & 1) ? -(int64_t)((x + 1) >> 1) : (int64_t)(x >> 1); }
inline uint64_t int2nat(int64_t x) { return x >= 0 ? (uint64_t)x << 1 : (uint64_t)(-x) * 2 - 1; }

// One BV record, read in webgraph-rs order (SURVEY.md 8a "BV record order"); `ref_list`
// supplies the successors of node v-r.  Mirrors webgraph's BvGraphSeq::Iter /
// BvGraph::successors driving src/ans/decoder.rs:103-139.
template <class RefLookup>
inline void decode_node(ANSDecoder& dec, size_t v, size_t window, size_t min_interval_length,
                        RefLookup&& ref_list, std::vector<uint64_t>& out) {
  out.clear();
  uint64_t degree = dec.decode(Outdegree);
  if (degree == 0) return;
  uint64_t ref_delta = window != 0 ? dec.decode(ReferenceOffset) : 0;
  if (ref_delta != 0) {
    const std::vector<uint64_t>& nb = ref_list(v - ref_delta);
    uint64_t nblocks = dec.decode(BlockCount);
    if (nblocks == 0) {
      out.insert(out.end(), nb.begin(), nb.end());
    } else {
      uint64_t idx = dec.decode(Blocks);
      out.insert(out.end(), nb.begin(), nb.begin() + idx);
      for (uint64_t b = 1; b < nblocks; ++b) {
        uint64_t block = dec.decode(Blocks);
        uint64_t end = idx + block + 1;
        if (b % 2 == 0) out.insert(out.end(), nb.begin() + idx, nb.begin() + end);
        idx = end;
      }
      if ((nblocks & 1) == 0) out.insert(out.end(), nb.begin() + idx, nb.end());
    }
  }
  uint64_t left = degree - out.size();
  if (left != 0 && min_interval_length != 0) {
    uint64_t nint = dec.decode(IntervalCount);
    if (nint != 0) {
      int64_t start = (int64_t)v + nat2int(dec.decode(IntervalStart));
      uint64_t delta = dec.decode(IntervalLen) + min_interval_length;
      for (uint64_t i = 0; i < delta; ++i) out.push_back((uint64_t)start + i);
      start += (int64_t)delta;
      for (uint64_t k = 1; k < nint; ++k) {
        start += 1 + (int64_t)dec.decode(IntervalStart);
        delta = dec.decode(IntervalLen) + min_interval_length;
        for (uint64_t i = 0; i < delta; ++i) out.push_back((uint64_t)start + i);
        start += (int64_t)delta;
      }
    }
  }
  left = degree - out.size();
  if (left != 0) {
    uint64_t prev = (uint64_t)((int64_t)v + nat2int(dec.decode(FirstResidual)));
    out.push_back(prev);
    for (uint64_t k = 1; k < left; ++k) {
      prev = prev + 1 + dec.decode(Residual);
      out.push_back(prev);
    }
  }
  std::sort(out.begin(), out.end());
}

// Sequential decode of the whole graph (src/bvgraph/sequential.rs:29-51 + factory
// bvgraphseq_decoder_factory.rs:29-35): ONE decoder from (stream.len(), prelude.state).
// `sink(v, successors)` is called for every node in order. Returns the final decoder
// (state, pointer) so callers can check the "65536 / 0" invariant.
template <class Sink>
inline ANSCompressorPhase decode_sequential(const ANSGraph& g, const Model4Decoder& model, Sink&& sink,
                                            size_t first = 0, size_t last = SIZE_MAX) {
  const size_t n = g.number_of_nodes;
  const size_t w = g.compression_window;
  if (last > n) last = n;
  std::vector<std::vector<uint64_t>> backrefs(w + 1);
  ANSDecoder dec = (first == 0)
                       ? ANSDecoder(model, g.stream, g.state)
                       : ANSDecoder(model, g.stream, (size_t)g.pointers.at(n - 1 - first), g.states.at(n - 1 - first));
  // when starting mid-graph the window must be primed by decoding up to w*depth earlier nodes;
  // callers that need that use decode_random instead.
  for (size_t v = first; v < last; ++v) {
    std::vector<uint64_t>& cur = backrefs[v % (w + 1)];
    std::vector<uint64_t> tmp;
    decode_node(dec, v, w, g.min_interval_length,
                [&](size_t u) -> const std::vector<uint64_t>& { return backrefs[u % (w + 1)]; }, tmp);
    cur.swap(tmp);
    sink(v, cur);
  }
  return {dec.state, dec.stream_pointer};
}

// Random access (src/bvgraph/random_access.rs:52-82 + bvgraph_decoder_factory.rs:46-58):
// decoder for node v starts at (pointers[N-1-v], states[N-1-v]); references recurse.
inline void successors(const ANSGraph& g, const Model4Decoder& model, size_t v, std::vector<uint64_t>& out) {
  const size_t n = g.number_of_nodes;
  ANSDecoder dec(model, g.stream, (size_t)g.pointers.at(n - 1 - v), g.states.at(n - 1 - v));
  std::vector<uint64_t> ref;
  decode_node(dec, v, g.compression_window, g.min_interval_length,
              [&](size_t u) -> const std::vector<uint64_t>& {
                successors(g, model, u, ref);
                return ref;
              },
              out);
}

// ----------------------------------------------------------------- src/bvgraph/estimators/log2_estimator.rs:15-49
struct Log2Estimator {
  size_t cost(uint64_t value, int) const { return ilog2(value + 2); }
};

// ----------------------------------------------------------------- src/bvgraph/estimators/entropy_estimator.rs:33-113
struct EntropyEstimator {
  std::array<std::vector<size_t>, COMPONENTS> table;
  std::array<std::pair<size_t, size_t>, COMPONENTS> component_args;  // (fidelity, radix)
  std::array<uint64_t, COMPONENTS> folding_thresholds;

  static size_t calculate_symbol_cost(Symbol sym, Freq freq, size_t frame_size, uint16_t folding_offset,
                                      uint16_t folding_threshold, size_t radix) {  // :81-100
    uint16_t folds = sym < folding_threshold ? 0 : (uint16_t)((sym - folding_threshold) / folding_offset + 1);
    double probability = (double)freq / (double)(1ull << frame_size);
    double r = std::round(-std::log2(probability) * (double)(1 << 16));
    size_t shifted = r <= 0 ? 0 : (size_t)r;  // `as usize` saturates at 0
    return shifted + ((size_t)folds * radix) * (1u << 16);
  }

  explicit EntropyEstimator(const Model4Encoder& model) {  // :33-75, args = get_folding_params()
    for (int c = 0; c < COMPONENTS; ++c) {
      size_t fidelity = model[c].fidelity, radix = model[c].radix;
      component_args[c] = {fidelity, radix};
      Symbol max_folded_sym = fold_without_streaming_out(MAX_RAW_SYMBOL, radix, fidelity);
      folding_thresholds[c] = 1ull << (fidelity + radix - 1);
      table[c].resize((size_t)max_folded_sym + 1);
      for (size_t sym = 0; sym <= max_folded_sym; ++sym) {
        Freq f = 1;
        if (sym < model[c].table.size() && model[c].table[sym].freq != 0) f = model[c].table[sym].freq;
        table[c][sym] = calculate_symbol_cost((Symbol)sym, f, model[c].frame_size,
                                              (uint16_t)model[c].folding_offset,
                                              (uint16_t)model[c].folding_threshold, radix);
      }
    }
  }
  size_t cost(uint64_t value, int c) const {  // :103-113
    uint64_t symbol = value < folding_thresholds[c]
                          ? value
                          : fold_without_streaming_out(value, component_args[c].second, component_args[c].first);
    return table[c].at(symbol);
  }
};

// ----------------------------------------------------------------- webgraph-rs BvComp (external, un-vendored; SURVEY.md 8a [MEM])
// Writer concept: size_t write(int component, uint64_t value)  -> estimated bits
//                 Estimator& estimator()
struct Compressor {
  size_t outdegree = 0;
  std::vector<size_t> blocks, extra_nodes, left_interval, len_interval, residuals;
  void clear() {
    outdegree = 0;
    blocks.clear(); extra_nodes.clear(); left_interval.clear(); len_interval.clear(); residuals.clear();
  }
  void diff_comp(const std::vector<size_t>& curr, const std::vector<size_t>& ref) {
    size_t j = 0, k = 0, len = 0;
    bool copying = true;
    while (j < curr.size() && k < ref.size()) {
      if (copying) {
        if (curr[j] > ref[k]) { blocks.push_back(len); copying = false; len = 0; }
        else if (curr[j] < ref[k]) { extra_nodes.push_back(curr[j]); j++; }
        else { j++; k++; len++; }
      } else if (curr[j] < ref[k]) { extra_nodes.push_back(curr[j]); j++; }
      else if (curr[j] > ref[k]) { k++; len++; }
      else { blocks.push_back(len); copying = true; len = 0; }
    }
    if (copying && k < ref.size()) blocks.push_back(len);
    while (j < curr.size()) extra_nodes.push_back(curr[j++]);
    if (!blocks.empty()) blocks[0] += 1;  // so that every block is written as blocks[i]-1
  }
  void intervalize(size_t min_len) {
    size_t vl = extra_nodes.size();
    size_t i = 0;
    while (i < vl) {
      size_t j = 0;
      if (i < vl - 1 && extra_nodes[i] + 1 == extra_nodes[i + 1]) {
        j++;
        while (i + j < vl - 1 && extra_nodes[i + j] + 1 == extra_nodes[i + j + 1]) j++;
        j++;
        if (j >= min_len) {
          left_interval.push_back(extra_nodes[i]);
          len_interval.push_back(j);
          i += j - 1;
        }
      }
      if (j < min_len) residuals.push_back(extra_nodes[i]);
      i++;
    }
  }
  void compress(const std::vector<size_t>& curr, const std::vector<size_t>* ref, size_t min_len) {
    clear();
    outdegree = curr.size();
    if (outdegree != 0) {
      if (ref) diff_comp(curr, *ref);
      else extra_nodes = curr;
      if (!extra_nodes.empty()) {
        if (min_len != 0) intervalize(min_len);
        else residuals = extra_nodes;
      }
    }
  }
  // reference_offset < 0 == None
  template <class W>
  uint64_t write(W& w, size_t curr_node, int64_t reference_offset, size_t min_len) const {
    uint64_t bits = 0;
    bits += w.write(Outdegree, outdegree);
    if (outdegree != 0 && reference_offset >= 0) {
      bits += w.write(ReferenceOffset, (uint64_t)reference_offset);
      if (reference_offset != 0) {
        bits += w.write(BlockCount, blocks.size());
        for (size_t i = 0; i < blocks.size(); ++i) bits += w.write(Blocks, blocks[i] - 1);
      }
    }
    if (!extra_nodes.empty() && min_len != 0) {
      bits += w.write(IntervalCount, left_interval.size());
      if (!left_interval.empty()) {
        bits += w.write(IntervalStart, int2nat((int64_t)left_interval[0] - (int64_t)curr_node));
        bits += w.write(IntervalLen, len_interval[0] - min_len);
        size_t prev = left_interval[0] + len_interval[0];
        for (size_t i = 1; i < left_interval.size(); ++i) {
          bits += w.write(IntervalStart, left_interval[i] - prev - 1);
          bits += w.write(IntervalLen, len_interval[i] - min_len);
          prev = left_interval[i] + len_interval[i];
        }
      }
    }
    if (!residuals.empty()) {
      bits += w.write(FirstResidual, int2nat((int64_t)residuals[0] - (int64_t)curr_node));
      for (size_t i = 1; i < residuals.size(); ++i) bits += w.write(Residual, residuals[i] - residuals[i - 1] - 1);
    }
    return bits;
  }
};

template <class W>
class BvComp {
 public:
  BvComp(W& enc, size_t window, size_t max_ref_count, size_t min_interval_length, size_t start_node)
      : encoder(enc), window(window), max_ref_count(max_ref_count), min_len(min_interval_length),
        curr_node(start_node), start_node(start_node), backrefs(window + 1), ref_counts(window + 1, 0),
        compressors(window + 1) {}

  void push(const std::vector<size_t>& succ) {
    backrefs[curr_node % (window + 1)] = succ;
    const std::vector<size_t>& curr_list = backrefs[curr_node % (window + 1)];
    arcs += curr_list.size();
    compressors[0].compress(curr_list, nullptr, min_len);
    if (window == 0) {
      compressors[0].write(encoder, curr_node, -1, min_len);
      curr_node++;
      return;
    }
    size_t ref_delta = 0;
    auto& est = encoder.estimator();
    uint64_t min_bits = compressors[0].write(est, curr_node, 0, min_len);
    size_t ref_count = 0;
    size_t deltas = 1 + std::min(window, curr_node - start_node);
    for (size_t delta = 1; delta < deltas; ++delta) {
      size_t ref_node = curr_node - delta;
      size_t count = ref_counts[ref_node % (window + 1)];
      if (count >= max_ref_count) continue;
      const std::vector<size_t>& ref_list = backrefs[ref_node % (window + 1)];
      if (ref_list.empty()) continue;
      compressors[delta].compress(curr_list, &ref_list, min_len);
      uint64_t bits = compressors[delta].write(est, curr_node, (int64_t)delta, min_len);
      if (bits < min_bits) {
        min_bits = bits;
        ref_delta = delta;
        ref_count = count + 1;
      }
    }
    compressors[ref_delta].write(encoder, curr_node, (int64_t)ref_delta, min_len);
    ref_counts[curr_node % (window + 1)] = ref_count;
    curr_node++;
  }

  W& encoder;
  size_t window, max_ref_count, min_len, curr_node, start_node;
  std::vector<std::vector<size_t>> backrefs;
  std::vector<size_t> ref_counts;
  std::vector<Compressor> compressors;
  uint64_t arcs = 0;
};

// ----------------------------------------------------------------- src/bvgraph/writers/bvgraph_model_builder.rs:11-112
template <class Est>
struct EstimatorWriter {  // adapts an estimator to the writer concept
  const Est& e;
  size_t write(int c, uint64_t v) { return e.cost(v, c); }
};

template <class Est>
struct BVGraphModelBuilder {
  explicit BVGraphModelBuilder(const Est& e) : mock{e} {}
  size_t write(int c, uint64_t v) {  // :51-103
    builder.push_symbol(v, c);
    return mock.write(c, v);
  }
  EstimatorWriter<Est>& estimator() { return mock; }
  ANSModel4EncoderBuilder builder;
  EstimatorWriter<Est> mock;
};

// ----------------------------------------------------------------- src/bvgraph/writers/bvgraph_encoder.rs:15-179
// The reverse gamma spill buffers (src/utils/rev.rs) are an I/O device for CPU RAM limits;
// the oracle keeps (symbol, component) pairs in memory.
struct ANSBVGraphEncodeAndEstimate {
  ANSBVGraphEncodeAndEstimate(const Model4Encoder& model, const EntropyEstimator& est)
      : mock{est}, encoder(model) {}
  size_t write(int c, uint64_t v) {  // :103-156
    symbols.push_back(v);
    comps.push_back((uint8_t)c);
    return mock.write(c, v);
  }
  EstimatorWriter<EntropyEstimator>& estimator() { return mock; }
  void flush() {  // :159-174
    for (size_t i = symbols.size(); i-- > 0;) {
      encoder.encode(symbols[i], comps[i]);
      if (comps[i] == Outdegree) phases.push_back(encoder.get_current_compressor_phase());
    }
  }
  EstimatorWriter<EntropyEstimator> mock;
  ANSEncoder encoder;
  std::vector<uint64_t> symbols;
  std::vector<uint8_t> comps;
  std::vector<ANSCompressorPhase> phases;
};

// ----------------------------------------------------------------- src/bvgraph/random_access.rs:91-222 (ANSBvGraph::store, in memory)
// `for_each_node(cb)` must call cb(successors) for every node 0..n-1 each time it is invoked.
struct StoreTrace {  // optional: symbols seen by each pass, for parity tests of the product's front end
  std::vector<uint64_t> pass2_symbols;
  std::vector<uint8_t> pass2_comps;
  Model4Encoder model1, model2;
};

inline ANSGraph store(const std::function<void(const std::function<void(const std::vector<size_t>&)>&)>& for_each_node,
                      size_t num_nodes, size_t window, size_t max_ref_count, size_t min_interval_length,
                      StoreTrace* trace = nullptr) {
  // PASS 1 (:105-131)
  Log2Estimator log2;
  BVGraphModelBuilder<Log2Estimator> mb1(log2);
  {
    BvComp<BVGraphModelBuilder<Log2Estimator>> comp(mb1, window, max_ref_count, min_interval_length, 0);
    for_each_node([&](const std::vector<size_t>& s) { comp.push(s); });
  }
  Model4Encoder model1 = mb1.builder.build();
  // PASS 2 (:134-163)
  EntropyEstimator entropy(model1);
  BVGraphModelBuilder<EntropyEstimator> mb2(entropy);
  {
    BvComp<BVGraphModelBuilder<EntropyEstimator>> comp(mb2, window, max_ref_count, min_interval_length, 0);
    for_each_node([&](const std::vector<size_t>& s) { comp.push(s); });
  }
  Model4Encoder model2 = mb2.builder.build();
  // PASS 3 (:166-196)  -- same estimator as pass 2
  ANSBVGraphEncodeAndEstimate enc(model2, entropy);
  uint64_t arcs = 0;
  {
    BvComp<ANSBVGraphEncodeAndEstimate> comp(enc, window, max_ref_count, min_interval_length, 0);
    for_each_node([&](const std::vector<size_t>& s) { comp.push(s); });
    arcs = comp.arcs;
  }
  enc.flush();
  if (trace) {
    trace->pass2_symbols = enc.symbols;
    trace->pass2_comps = enc.comps;
    trace->model1 = model1;
    trace->model2 = model2;
  }
  ANSGraph g;
  g.tables = model2;
  g.stream = enc.encoder.stream;
  g.state = enc.encoder.state;
  g.number_of_nodes = num_nodes;
  g.compression_window = window;
  g.min_interval_length = min_interval_length;
  g.number_of_arcs = arcs;
  for (auto& p : enc.phases) {  // :202, :225-231
    g.states.push_back(p.state);
    g.pointers.push_back(p.stream_pointer);
  }
  return g;
}

}  // namespace wgo
