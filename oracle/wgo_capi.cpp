// =============================================================================
//  wgo_capi.cpp -- C interface of the CPU ORACLE (ctypes-friendly).
//  TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and the
//  cpu_baseline / --impl reference legs of bench.py.  See wgo.hpp.
// =============================================================================
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#include <unordered_map>

#include "../tools/synth/synth_graph.hpp"
#include "wgo_io.hpp"

using namespace wgo;

namespace {
thread_local std::string g_err;
struct Graph {
  ANSGraph g;
  std::unique_ptr<Model4Decoder> dec;
  StoreTrace trace;
  ModelBuildInfo info{};
  const Model4Decoder& model() {
    if (!dec) dec.reset(new Model4Decoder(g.tables));
    return *dec;
  }
};
template <class F>
int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Decode nodes [first,last) from the phase of `first`; references that fall before `first`
// are resolved through random access (bvgraph_decoder_factory.rs:46-58).
template <class Sink>
void decode_range(Graph& G, size_t first, size_t last, Sink&& sink) {
  const ANSGraph& g = G.g;
  const Model4Decoder& model = G.model();
  const size_t n = g.number_of_nodes, w = g.compression_window;
  if (first >= last) return;
  std::vector<std::vector<uint64_t>> back(w + 1);
  std::unordered_map<size_t, std::vector<uint64_t>> halo;
  ANSDecoder dec = first == 0 && g.states.empty()
                       ? ANSDecoder(model, g.stream, g.state)
                       : (first == 0 ? ANSDecoder(model, g.stream, g.state)
                                     : ANSDecoder(model, g.stream, (size_t)g.pointers.at(n - 1 - first),
                                                  g.states.at(n - 1 - first)));
  std::vector<uint64_t> tmp;
  for (size_t v = first; v < last; ++v) {
    decode_node(dec, v, w, g.min_interval_length,
                [&](size_t u) -> const std::vector<uint64_t>& {
                  if (u >= first) return back[u % (w + 1)];
                  auto it = halo.find(u);
                  if (it == halo.end()) {
                    std::vector<uint64_t> s;
                    successors(g, model, u, s);
                    it = halo.emplace(u, std::move(s)).first;
                  }
                  return it->second;
                },
                tmp);
    back[v % (w + 1)].swap(tmp);
    sink(v, back[v % (w + 1)]);
  }
}
}  // namespace

extern "C" {

const char* wgo_last_error() { return g_err.c_str(); }

void* wgo_graph_new() { return new Graph(); }
void wgo_graph_free(void* h) { delete (Graph*)h; }

int wgo_graph_load(void* h, const char* basename, int random_access) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->g = load_ans_graph(basename, random_access != 0);
    G->dec.reset();
  });
}

int wgo_graph_set_table(void* h, int c, const void* entries, uint64_t len, uint64_t frame_size, uint64_t radix,
                        uint64_t fidelity, uint64_t thr, uint64_t off) {
  return guard([&] {
    Graph* G = (Graph*)h;
    auto& t = G->g.tables.at(c);
    t.table.resize(len);
    if (len) std::memcpy(t.table.data(), entries, len * 8);
    t.frame_size = frame_size; t.radix = radix; t.fidelity = fidelity;
    t.folding_threshold = thr; t.folding_offset = off;
    G->dec.reset();
  });
}
int wgo_graph_set_stream(void* h, const uint16_t* s, uint64_t len, uint32_t state) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->g.stream.assign(s, s + len);
    G->g.state = state;
  });
}
int wgo_graph_set_meta(void* h, uint64_t n, uint64_t window, uint64_t min_interval, uint64_t arcs) {
  Graph* G = (Graph*)h;
  G->g.number_of_nodes = n; G->g.compression_window = window;
  G->g.min_interval_length = min_interval; G->g.number_of_arcs = arcs;
  return 0;
}
int wgo_graph_set_phases(void* h, const uint32_t* states, const uint64_t* pointers, uint64_t n) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->g.states.assign(states, states + n);
    G->g.pointers.assign(pointers, pointers + n);
  });
}
// out: n, arcs, window, min_interval, stream_len, state, n_phases
void wgo_graph_info(void* h, uint64_t* out) {
  Graph* G = (Graph*)h;
  out[0] = G->g.number_of_nodes; out[1] = G->g.number_of_arcs; out[2] = G->g.compression_window;
  out[3] = G->g.min_interval_length; out[4] = G->g.stream.size(); out[5] = G->g.state;
  out[6] = G->g.states.size();
}
// out: table_len, frame_size(log2), radix, fidelity, folding_threshold, folding_offset
void wgo_graph_table_params(void* h, int c, uint64_t* out) {
  auto& t = ((Graph*)h)->g.tables.at(c);
  out[0] = t.table.size(); out[1] = t.frame_size; out[2] = t.radix; out[3] = t.fidelity;
  out[4] = t.folding_threshold; out[5] = t.folding_offset;
}
void wgo_graph_table(void* h, int c, void* out) {
  auto& t = ((Graph*)h)->g.tables.at(c);
  if (!t.table.empty()) std::memcpy(out, t.table.data(), t.table.size() * 8);
}
void wgo_graph_stream(void* h, uint16_t* out) {
  auto& s = ((Graph*)h)->g.stream;
  if (!s.empty()) std::memcpy(out, s.data(), s.size() * 2);
}
void wgo_graph_phases(void* h, uint32_t* states, uint64_t* pointers) {
  Graph* G = (Graph*)h;
  if (!G->g.states.empty()) {
    std::memcpy(states, G->g.states.data(), G->g.states.size() * 4);
    std::memcpy(pointers, G->g.pointers.data(), G->g.pointers.size() * 8);
  }
}
// expanded decoder table of component c in the REFERENCE layout (16-byte entries); returns #slots
uint64_t wgo_graph_decoder_table(void* h, int c, void* out) {
  Graph* G = (Graph*)h;
  const auto& t = G->model().tables.at(c).table;
  if (out) std::memcpy(out, t.data(), t.size() * 16);
  return t.size();
}

// ---- model building (src/ans/model4encoder_builder.rs) -------------------------------------
// Builds the 9 tables from (component, raw symbol) pairs and installs them in the graph handle.
int wgo_model_build(void* h, const uint8_t* comps, const uint64_t* syms, uint64_t n, double* orig_cost9,
                    double* final_cost9) {
  return guard([&] {
    Graph* G = (Graph*)h;
    ANSModel4EncoderBuilder b;
    for (uint64_t i = 0; i < n; ++i)
      if (!b.push_symbol(syms[i], comps[i])) throw std::runtime_error("Symbol can't be bigger than u48::MAX");
    G->g.tables = b.build(&G->info);
    G->dec.reset();
    for (int c = 0; c < COMPONENTS; ++c) {
      if (orig_cost9) orig_cost9[c] = G->info.original_cost[c];
      if (final_cost9) final_cost9[c] = G->info.final_cost[c];
    }
  });
}
// same, from a sparse histogram (component, raw symbol, count)
int wgo_model_build_hist(void* h, const uint8_t* comps, const uint64_t* syms, const uint64_t* counts, uint64_t n) {
  return guard([&] {
    Graph* G = (Graph*)h;
    ANSModel4EncoderBuilder b;
    for (uint64_t i = 0; i < n; ++i) b.push_symbol_count(syms[i], comps[i], counts[i]);
    G->g.tables = b.build(&G->info);
    G->dec.reset();
  });
}

uint16_t wgo_fold(uint64_t sym, uint64_t radix, uint64_t fidelity) {
  try { return fold_without_streaming_out(sym, radix, fidelity); } catch (...) { return 0xFFFF; }
}

// ---- encoder / decoder on symbol sequences (tests/compressor_tests.rs) ------------------------
// Encodes symbols IN THE GIVEN ORDER with the graph's tables; stores stream + state in the handle,
// and a phase after every Outdegree symbol (bvgraph_encoder.rs:168-171).
int wgo_encode_symbols(void* h, const uint8_t* comps, const uint64_t* syms, uint64_t n) {
  return guard([&] {
    Graph* G = (Graph*)h;
    ANSEncoder enc(G->g.tables);
    G->g.states.clear(); G->g.pointers.clear();
    for (uint64_t i = 0; i < n; ++i) {
      enc.encode(syms[i], comps[i]);
      if (comps[i] == Outdegree) {
        auto p = enc.get_current_compressor_phase();
        G->g.states.push_back(p.state);
        G->g.pointers.push_back(p.stream_pointer);
      }
    }
    G->g.stream = enc.stream;
    G->g.state = enc.state;
  });
}
// Decodes n symbols of the given components from (state, ptr) [ptr == UINT64_MAX: sequential start].
int wgo_decode_symbols(void* h, const uint8_t* comps, uint64_t n, uint64_t ptr, uint32_t state, uint64_t* out,
                       uint64_t* end_ptr, uint32_t* end_state) {
  return guard([&] {
    Graph* G = (Graph*)h;
    ANSDecoder d = ptr == UINT64_MAX ? ANSDecoder(G->model(), G->g.stream, G->g.state)
                                     : ANSDecoder(G->model(), G->g.stream, (size_t)ptr, state);
    for (uint64_t i = 0; i < n; ++i) out[i] = d.decode(comps[i]);
    if (end_ptr) *end_ptr = d.stream_pointer;
    if (end_state) *end_state = d.state;
  });
}

// ---- graph decode -----------------------------------------------------------------------------
// Sequential decode (ANSBvGraphSeq::load + iter). offsets: last-first+1 entries (relative to first),
// succ: u32 ids, capacity `cap`. Returns #arcs via *arcs and the final decoder phase.
int wgo_decode_seq(void* h, uint64_t first, uint64_t last, uint64_t* offsets, uint32_t* succ, uint64_t cap,
                   uint64_t* arcs, uint64_t* end_ptr, uint32_t* end_state) {
  return guard([&] {
    Graph* G = (Graph*)h;
    uint64_t pos = 0;
    if (offsets) offsets[0] = 0;
    if (first != 0) {
      decode_range(*G, first, last, [&](size_t v, const std::vector<uint64_t>& s) {
        for (uint64_t x : s) {
          if (succ) { if (pos >= cap) throw std::runtime_error("succ buffer too small"); succ[pos] = (uint32_t)x; }
          pos++;
        }
        if (offsets) offsets[v - first + 1] = pos;
      });
      if (arcs) *arcs = pos;
      return;
    }
    ANSCompressorPhase end = decode_sequential(G->g, G->model(), [&](size_t v, const std::vector<uint64_t>& s) {
      for (uint64_t x : s) {
        if (succ) { if (pos >= cap) throw std::runtime_error("succ buffer too small"); succ[pos] = (uint32_t)x; }
        pos++;
      }
      if (offsets) offsets[v + 1] = pos;
    }, 0, last);
    if (arcs) *arcs = pos;
    if (end_ptr) *end_ptr = end.stream_pointer;
    if (end_state) *end_state = end.state;
  });
}
// Random access: successors of one node (ANSBvGraph::load + successors). Returns count or -1.
int64_t wgo_successors(void* h, uint64_t v, uint64_t* out, uint64_t cap) {
  int64_t cnt = -1;
  guard([&] {
    Graph* G = (Graph*)h;
    std::vector<uint64_t> s;
    successors(G->g, G->model(), v, s);
    if (s.size() > cap) throw std::runtime_error("buffer too small");
    std::copy(s.begin(), s.end(), out);
    cnt = (int64_t)s.size();
  });
  return cnt;
}
// Node-range-parallel decode on `nthreads` host threads (each range starts from its phase).
// offsets (n+1, absolute) and succ may be NULL (count only). Returns seconds in *secs.
int wgo_decode_parallel(void* h, uint64_t first, uint64_t last, int nthreads, uint64_t* arcs, double* secs) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->model();
    if (nthreads < 1) nthreads = 1;
    std::vector<uint64_t> counts(nthreads, 0);
    std::vector<std::string> errs(nthreads);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    uint64_t span = last - first;
    for (int t = 0; t < nthreads; ++t) {
      th.emplace_back([&, t] {
        try {
          uint64_t a = first + span * t / nthreads, b = first + span * (t + 1) / nthreads;
          uint64_t c = 0;
          decode_range(*G, a, b, [&](size_t, const std::vector<uint64_t>& s) { c += s.size(); });
          counts[t] = c;
        } catch (const std::exception& e) { errs[t] = e.what(); }
      });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
    uint64_t tot = 0;
    for (auto c : counts) tot += c;
    if (arcs) *arcs = tot;
    if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  });
}
// Node-range-parallel decode INTO caller arrays, for whole-graph verification of a device result:
// `offsets` (last-first+1 entries, relative to first) are given; every node's decoded outdegree must match
// them, and its successors are written to succ[offsets[v-first]..].  Returns 0, or -1 with a message.
int wgo_decode_parallel_into(void* h, uint64_t first, uint64_t last, int nthreads, const uint64_t* offsets,
                             uint32_t* succ) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->model();
    if (nthreads < 1) nthreads = 1;
    std::vector<std::string> errs(nthreads);
    std::vector<std::thread> th;
    uint64_t span = last - first;
    for (int t = 0; t < nthreads; ++t) {
      th.emplace_back([&, t] {
        try {
          uint64_t a = first + span * t / nthreads, b = first + span * (t + 1) / nthreads;
          decode_range(*G, a, b, [&](size_t v, const std::vector<uint64_t>& s) {
            uint64_t o = offsets[v - first];
            if (offsets[v - first + 1] - o != s.size())
              throw std::runtime_error("outdegree mismatch at node " + std::to_string(v));
            for (size_t i = 0; i < s.size(); ++i) succ[o + i] = (uint32_t)s[i];
          });
        } catch (const std::exception& e) { errs[t] = e.what(); }
      });
    }
    for (auto& x : th) x.join();
    for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
  });
}
// Random-access benchmark (examples/bench_random_access.rs:28-41): sum of outdegrees over `n` nodes.
int wgo_random_access_bench(void* h, const uint64_t* nodes, uint64_t n, uint64_t* arcs, double* secs) {
  return guard([&] {
    Graph* G = (Graph*)h;
    G->model();
    std::vector<uint64_t> s;
    uint64_t c = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (uint64_t i = 0; i < n; ++i) {
      successors(G->g, G->model(), nodes[i], s);
      c += s.size();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (arcs) *arcs = c;
    if (secs) *secs = std::chrono::duration<double>(t1 - t0).count();
  });
}

// ---- ANSBvGraph::store, in memory (src/bvgraph/random_access.rs:91-222) -----------------------
int wgo_store_csr(void* h, const uint64_t* offsets, const uint32_t* succ, uint64_t n, uint64_t window,
                  uint64_t max_ref_count, uint64_t min_interval) {
  return guard([&] {
    Graph* G = (Graph*)h;
    auto each = [&](const std::function<void(const std::vector<size_t>&)>& cb) {
      std::vector<size_t> s;
      for (uint64_t v = 0; v < n; ++v) {
        s.assign(succ + offsets[v], succ + offsets[v + 1]);
        cb(s);
      }
    };
    G->g = store(each, n, window, max_ref_count, min_interval, &G->trace);
    G->dec.reset();
  });
}
uint64_t wgo_trace_len(void* h) { return ((Graph*)h)->trace.pass2_symbols.size(); }
void wgo_trace_get(void* h, uint8_t* comps, uint64_t* syms) {
  Graph* G = (Graph*)h;
  std::copy(G->trace.pass2_comps.begin(), G->trace.pass2_comps.end(), comps);
  std::copy(G->trace.pass2_symbols.begin(), G->trace.pass2_symbols.end(), syms);
}

// ---- BV .graph reader (golden input) ------------------------------------------------------------
void* wgo_bv_read(const char* basename) {
  CSR* c = nullptr;
  guard([&] { c = new CSR(read_bvgraph(basename)); });
  return c;
}
void wgo_bv_free(void* c) { delete (CSR*)c; }
uint64_t wgo_bv_nodes(void* c) { return ((CSR*)c)->offsets.size() - 1; }
uint64_t wgo_bv_arcs(void* c) { return ((CSR*)c)->succ.size(); }
void wgo_bv_get(void* c, uint64_t* offsets, uint32_t* succ, uint64_t* bit_offsets) {
  CSR* g = (CSR*)c;
  if (offsets) std::copy(g->offsets.begin(), g->offsets.end(), offsets);
  if (succ) for (size_t i = 0; i < g->succ.size(); ++i) succ[i] = (uint32_t)g->succ[i];
  if (bit_offsets) std::copy(g->bit_offsets.begin(), g->bit_offsets.end(), bit_offsets);
}
// plain values of an epserde Elias-Fano file (.ef / .pointers)
int64_t wgo_ef_read(const char* path, uint64_t* out, uint64_t cap) {
  int64_t n = -1;
  guard([&] {
    auto v = load_elias_fano(path);
    if (out) {
      if (v.size() > cap) throw std::runtime_error("buffer too small");
      std::copy(v.begin(), v.end(), out);
    }
    n = (int64_t)v.size();
  });
  return n;
}

// Synthetic benchmark graphs (tools/synth/synth_graph.hpp: workload infrastructure, shared with the product's test
// API) so that bench.py --impl reference builds its inputs without loading the product library.
int wgo_synth_graph(int kind, uint64_t n_nodes, double mean_degree, uint64_t seed, uint64_t first, uint64_t last,
                    int threads, uint64_t* offsets, uint32_t* succ, uint64_t* n_arcs) {
  try {
    uint64_t a = wgsynth::synth_graph(kind, n_nodes, mean_degree, seed, first, last, threads, offsets, succ);
    if (n_arcs) *n_arcs = a;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
