// Host front end of the recompression: BvComp + estimators + ANS encoder.  See bvcomp.hpp.
#include "bvcomp.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

namespace wga {

// --------------------------------------------------------------------------------- EntropyEstimator
// cost(sym) = round(-log2(freq / 2^frame) * 2^16) + folds * radix * 2^16   (entropy_estimator.rs:81-100)
// freq = table[sym].freq, or 1 when the entry is missing or zero (:49-55).
Estimator::Estimator(const ComponentModel tables[WGA_COMPONENTS]) : log2_(false) {
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    const ComponentModel& m = tables[c];
    const uint64_t F = m.fidelity, R = m.radix;
    if (F == 0 || R == 0 || F + R > 40) throw Error(WGA_E_ARG, "bad fidelity/radix");
    fid_[c] = (unsigned)F;
    rad_[c] = (unsigned)R;
    thr_[c] = 1ull << (F + R - 1);                      // :42
    off_[c] = ((1ull << R) - 1) * (1ull << (F - 1));    // :43
    const uint64_t cuts = (48 - F) / R;                 // fold(2^48-1)
    const uint64_t max_folded = (((1ull << 48) - 1) >> (cuts * R)) + off_[c] * cuts;
    if (max_folded > 0xFFFF) throw Error(WGA_E_ARG, "Folded symbol is bigger than u16::MAX");
    table_[c].resize(max_folded + 1);
    const uint16_t m_thr = (uint16_t)m.folding_threshold, m_off = (uint16_t)m.folding_offset;  // :61-62
    for (uint64_t sym = 0; sym <= max_folded; ++sym) {
      uint16_t freq = 1;
      if (sym < m.table.size() && m.table[sym].freq != 0) freq = m.table[sym].freq;
      uint16_t folds = (uint16_t)sym < m_thr ? 0 : (uint16_t)(((uint16_t)sym - m_thr) / m_off + 1);
      double probability = (double)freq / (double)(1ull << m.frame_size);
      double r = std::round(-std::log2(probability) * 65536.0);
      uint64_t shifted = r <= 0 ? 0 : (uint64_t)r;
      table_[c][sym] = (uint32_t)(shifted + ((uint64_t)folds * R) * 65536ull);
    }
  }
}

namespace {

inline uint64_t int2nat(int64_t x) { return x >= 0 ? (uint64_t)x << 1 : ((uint64_t)(-x) << 1) - 1; }

// One candidate encoding of a successor list (webgraph-rs bvcomp Compressor).
struct Candidate {
  uint64_t outdegree = 0;
  std::vector<uint64_t> blocks, extras, left, len, residuals;

  void reset() {
    outdegree = 0;
    blocks.clear(); extras.clear(); left.clear(); len.clear(); residuals.clear();
  }

  // copy/skip blocks of `cur` against `ref`; the first block carries +1 so every block is emitted as b-1
  void diff(const std::vector<uint64_t>& cur, const std::vector<uint64_t>& ref) {
    size_t j = 0, k = 0;
    uint64_t run = 0;
    bool copying = true;
    const size_t nc = cur.size(), nr = ref.size();
    while (j < nc && k < nr) {
      const uint64_t a = cur[j], b = ref[k];
      if (copying) {
        if (a > b) { blocks.push_back(run); copying = false; run = 0; }
        else if (a < b) { extras.push_back(a); ++j; }
        else { ++j; ++k; ++run; }
      } else {
        if (a < b) { extras.push_back(a); ++j; }
        else if (a > b) { ++k; ++run; }
        else { blocks.push_back(run); copying = true; run = 0; }
      }
    }
    if (copying && k < nr) blocks.push_back(run);
    for (; j < nc; ++j) extras.push_back(cur[j]);
    if (!blocks.empty()) blocks[0] += 1;
  }

  // maximal runs of consecutive ids of length >= L become intervals
  void intervalize(uint64_t L) {
    const size_t n = extras.size();
    size_t i = 0;
    while (i < n) {
      size_t run = 1;
      while (i + run < n && extras[i + run] == extras[i + run - 1] + 1) ++run;
      if (run >= 2 && run >= L) {
        left.push_back(extras[i]);
        len.push_back(run);
      } else {
        for (size_t t = 0; t < run; ++t) residuals.push_back(extras[i + t]);
      }
      i += run;
    }
  }

  void compress(const std::vector<uint64_t>& cur, const std::vector<uint64_t>* ref, uint64_t L) {
    reset();
    outdegree = cur.size();
    if (!outdegree) return;
    if (ref) diff(cur, *ref);
    else extras = cur;
    if (extras.empty()) return;
    if (L) intervalize(L);
    else residuals = extras;
  }

  // emits the record through `emit(component, value)`; returns the sum of what emit returns
  template <class Emit>
  uint64_t write(uint64_t node, int64_t reference /* <0: none */, uint64_t L, Emit&& emit) const {
    uint64_t bits = emit(Outdegree, outdegree);
    if (outdegree && reference >= 0) {
      bits += emit(ReferenceOffset, (uint64_t)reference);
      if (reference) {
        bits += emit(BlockCount, blocks.size());
        for (uint64_t b : blocks) bits += emit(Blocks, b - 1);
      }
    }
    if (!extras.empty() && L) {
      bits += emit(IntervalCount, left.size());
      uint64_t prev_end = 0;
      for (size_t i = 0; i < left.size(); ++i) {
        bits += emit(IntervalStart, i == 0 ? int2nat((int64_t)left[0] - (int64_t)node) : left[i] - prev_end - 1);
        bits += emit(IntervalLen, len[i] - L);
        prev_end = left[i] + len[i];
      }
    }
    for (size_t i = 0; i < residuals.size(); ++i)
      bits += emit(i == 0 ? FirstResidual : Residual,
                   i == 0 ? int2nat((int64_t)residuals[0] - (int64_t)node) : residuals[i] - residuals[i - 1] - 1);
    return bits;
  }
};

}  // namespace

uint64_t bvcomp_range(const NodeSource& src, uint64_t first, uint64_t last, const BvCompParams& p,
                      const Estimator& est, SymbolStream& out, const uint16_t* choice, uint64_t choice_first) {
  const uint64_t W = p.window, L = p.min_interval_length;
  std::vector<std::vector<uint64_t>> lists(W + 1);
  std::vector<uint64_t> ref_counts(W + 1, 0);
  std::vector<Candidate> cand(W + 1);
  uint64_t arcs = 0;
  auto cost = [&](int c, uint64_t v) { return est.cost(c, v); };
  auto tap = [&](int c, uint64_t v) -> uint64_t {
    if (v > (1ull << 48) - 1) throw Error(WGA_E_ARG, "Symbol can't be bigger than u48::MAX");
    out.push(c, v);
    return 0;
  };
  for (uint64_t v = first; v < last; ++v) {
    std::vector<uint64_t>& cur = lists[v % (W + 1)];
    src(v, cur);
    arcs += cur.size();
    cand[0].compress(cur, nullptr, L);
    if (W == 0) {
      cand[0].write(v, -1, L, tap);
      continue;
    }
    if (choice) {  // the reference was chosen on the GPU: only the chosen candidate is compressed
      const uint64_t best = choice[v - choice_first];
      if (best > std::min<uint64_t>(W, v - first)) throw Error(WGA_E_ARG, "bad precomputed reference choice");
      if (best) cand[best].compress(cur, &lists[(v - best) % (W + 1)], L);
      cand[best].write(v, (int64_t)best, L, tap);
      continue;
    }
    uint64_t best_bits = cand[0].write(v, 0, L, cost);
    uint64_t best = 0, best_count = 0;
    const uint64_t deltas = std::min<uint64_t>(W, v - first);
    for (uint64_t delta = 1; delta <= deltas; ++delta) {
      const uint64_t u = v - delta;
      const uint64_t count = ref_counts[u % (W + 1)];
      if (count >= p.max_ref_count) continue;
      const std::vector<uint64_t>& ref = lists[u % (W + 1)];
      if (ref.empty()) continue;
      cand[delta].compress(cur, &ref, L);
      const uint64_t bits = cand[delta].write(v, (int64_t)delta, L, cost);
      if (bits < best_bits) {  // strict: the nearest candidate wins ties
        best_bits = bits;
        best = delta;
        best_count = count + 1;
      }
    }
    cand[best].write(v, (int64_t)best, L, tap);
    ref_counts[v % (W + 1)] = best_count;
  }
  return arcs;
}

uint64_t bvcomp_graph(const NodeSource& src, uint64_t n_nodes, const BvCompParams& p, const Estimator& est,
                      uint64_t chunk_nodes, int threads, SymbolStream& out) {
  return bvcomp_nodes(src, 0, n_nodes, p, est, chunk_nodes, threads, out);
}

void bvcomp_costs_host(const NodeSource& src, uint64_t first, uint64_t n, const BvCompParams& p, const Estimator& est,
                       uint64_t chunk_nodes, std::vector<uint64_t>& costs) {
  const uint64_t W = p.window, L = p.min_interval_length;
  costs.assign(n * (W + 1), ~0ull);
  std::vector<std::vector<uint64_t>> lists(W + 1);
  Candidate cand;
  auto cost = [&](int c, uint64_t v) { return est.cost(c, v); };
  for (uint64_t v = first; v < first + n; ++v) {
    std::vector<uint64_t>& cur = lists[v % (W + 1)];
    src(v, cur);
    const uint64_t a = chunk_nodes ? std::max(first, v / chunk_nodes * chunk_nodes) : first;
    cand.compress(cur, nullptr, L);
    costs[(v - first) * (W + 1)] = cand.write(v, 0, L, cost);
    for (uint64_t delta = 1; delta <= std::min<uint64_t>(W, v - a); ++delta) {
      const std::vector<uint64_t>& ref = lists[(v - delta) % (W + 1)];
      if (ref.empty()) continue;
      cand.compress(cur, &ref, L);
      costs[(v - first) * (W + 1) + delta] = cand.write(v, (int64_t)delta, L, cost);
    }
  }
}

uint64_t bvcomp_nodes(const NodeSource& src, uint64_t first, uint64_t last, const BvCompParams& p, const Estimator& est,
                      uint64_t chunk_nodes, int threads, SymbolStream& out, const uint16_t* choice) {
  if (chunk_nodes == 0) return bvcomp_range(src, first, last, p, est, out, choice, first);
  // chunks are aligned to multiples of chunk_nodes of the WHOLE graph, so that a node range compressed by one
  // rank gives the symbols the whole-graph run gives for those nodes
  const uint64_t c0 = first / chunk_nodes, c1 = (last + chunk_nodes - 1) / chunk_nodes;
  const uint64_t n_chunks = c1 > c0 ? c1 - c0 : 0;
  if (threads < 1) threads = 1;
  std::vector<SymbolStream> parts(n_chunks);
  std::vector<uint64_t> arcs(n_chunks, 0);
  std::vector<std::string> errs(threads);
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t] {
      try {
        for (uint64_t c = next.fetch_add(1); c < n_chunks; c = next.fetch_add(1)) {
          uint64_t a = std::max(first, (c0 + c) * chunk_nodes), b = std::min(last, (c0 + c + 1) * chunk_nodes);
          arcs[c] = bvcomp_range(src, a, b, p, est, parts[c], choice, first);
        }
      } catch (const std::exception& e) {
        errs[t] = e.what();
      }
    });
  }
  for (auto& th : pool) th.join();
  for (auto& e : errs)
    if (!e.empty()) throw Error(WGA_E_ARG, e);
  uint64_t total = 0, total_arcs = 0;
  for (auto& s : parts) total += s.size();
  out.comps.reserve(out.comps.size() + total);
  out.vals.reserve(out.vals.size() + total);
  for (uint64_t c = 0; c < n_chunks; ++c) {
    out.comps.insert(out.comps.end(), parts[c].comps.begin(), parts[c].comps.end());
    out.vals.insert(out.vals.end(), parts[c].vals.begin(), parts[c].vals.end());
    std::vector<uint8_t>().swap(parts[c].comps);
    std::vector<uint64_t>().swap(parts[c].vals);
    total_arcs += arcs[c];
  }
  return total_arcs;
}

// ------------------------------------------------------------------------------------- ANS encoder
void ans_encode(const ComponentModel tables[WGA_COMPONENTS], const uint8_t* comps, const uint64_t* vals,
                uint64_t n, EncodeResult& out) {
  out.stream.clear();
  out.phases.states.clear();
  out.phases.pointers.clear();
  uint32_t state = 1u << 16;  // INTERVAL_LOWER_BOUND, encoder.rs:22-28
  std::vector<uint16_t>& stream = out.stream;
  for (uint64_t i = n; i-- > 0;) {  // bvgraph_encoder.rs:163-172: replay backwards
    const int c = comps[i];
    if (c >= WGA_COMPONENTS) throw Error(WGA_E_ARG, "bad component index");
    const ComponentModel& m = tables[c];
    uint64_t sym = vals[i];
    const unsigned R = (unsigned)m.radix;
    if (sym >= m.folding_threshold) {  // encoder.rs:42-60
      const unsigned bits = 64u - (unsigned)__builtin_clzll(sym);
      const uint64_t folds = (bits - m.fidelity) / R;
      for (uint64_t f = 0; f < folds; ++f) {
        const uint32_t low = (uint32_t)(sym & ((1ull << R) - 1));
        if ((unsigned)__builtin_clz(state) < R) {  // no room: normalise first
          stream.push_back((uint16_t)state);
          state >>= 16;
        }
        state = (state << R) + low;
        sym >>= R;
      }
      sym += m.folding_offset * folds;
    }
    if (sym >= m.table.size()) throw Error(WGA_E_ARG, "symbol has no entry in the model");
    const wga_encoder_entry& e = m.table[sym];
    if (e.freq == 0) throw Error(WGA_E_ARG, "symbol has frequency 0 in the model");
    if (state >= e.upperbound) {  // :63-65
      stream.push_back((uint16_t)state);
      state >>= 16;
    }
    const uint32_t block = state / e.freq;  // :72-77
    state = (block << m.frame_size) + e.cumul_freq + (state - block * e.freq);
    if (c == Outdegree) {
      out.phases.states.push_back(state);
      out.phases.pointers.push_back(stream.size());
    }
  }
  out.state = state;
}

}  // namespace wga
