// extern "C" surface of libwgans (include/wga.h).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "bvcomp.hpp"
#include "formats.hpp"
#include "graph.hpp"
#include "model.hpp"

namespace wga {
const char* last_error_cstr();
uint64_t decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last);
void decode_range(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, uint32_t* d_succ,
                  uint64_t succ_capacity, void* ws, uint64_t ws_bytes, uint64_t* h_arcs, cudaStream_t st);
void outdegrees(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, void* ws, uint64_t ws_bytes,
                cudaStream_t st);
void debug_expand_table(wga_graph* g, int c, void* h_out, uint64_t n_slots);
void debug_decode_symbols(wga_graph* g, const uint8_t* h_comps, uint64_t n, uint64_t ptr, uint32_t state,
                          uint64_t* h_out, uint64_t* h_end_ptr, uint32_t* h_end_state);
int set_tuning(const char* key, uint64_t value);
uint64_t tuning_e2e_chunk();
void launch_offsets_add(uint64_t* off, uint64_t n, uint64_t base, cudaStream_t st);
uint64_t successors_workspace_size(const wga_graph* g, uint64_t n_queries, uint64_t max_total_arcs);
void successors_batch(wga_graph* g, const uint64_t* d_nodes, uint64_t n_queries, uint64_t* d_offsets,
                      uint32_t* d_succ, uint64_t succ_capacity, void* ws, uint64_t ws_bytes, uint64_t* h_arcs,
                      cudaStream_t st);
uint64_t synth_graph(int kind, uint64_t N, double mean_degree, uint64_t seed, uint64_t first, uint64_t last,
                     int threads, uint64_t* h_offsets, uint32_t* h_succ);

static void view_to_models(const wga_component_model* in, ComponentModel* out) {
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    out[c].table.assign(in[c].table, in[c].table + in[c].table_len);
    out[c].frame_size = in[c].frame_size;
    out[c].radix = in[c].radix;
    out[c].fidelity = in[c].fidelity;
    out[c].folding_threshold = in[c].folding_threshold;
    out[c].folding_offset = in[c].folding_offset;
  }
}

static wga_graph* finish_open(std::unique_ptr<wga_graph> g, uint64_t first, uint64_t last, int flags) {
  g->res_first = first;
  g->res_last = last;
  if (!(flags & WGA_OPEN_HOST_ONLY)) g->upload();
  else g->packed = pack_tables(g->prelude.tables);
  return g.release();
}

static std::string with_ext(const char* basename, const char* ext) { return std::string(basename) + "." + ext; }
static bool file_exists(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (f) fclose(f);
  return f != nullptr;
}

// ANSBvGraph::store on a node source (random_access.rs:91-222): pass 1 Log2Estimator -> model1,
// pass 2 EntropyEstimator(model1) -> model2, pass 3 encodes the pass-2 symbols (same estimator, :166-168)
static void store_from_source(const NodeSource& src, uint64_t n_nodes, const char* new_basename,
                              const BvCompParams& p, uint64_t chunk_nodes, int threads) {
  if (!new_basename) throw Error(WGA_E_ARG, "null basename");
  Prelude pre;
  {
    SymbolStream s1;
    bvcomp_graph(src, n_nodes, p, Estimator(), chunk_nodes, threads, s1);
    ModelBuilder mb;
    mb.accumulate_host(s1.comps.data(), s1.vals.data(), s1.size());
    mb.build(pre.tables, nullptr, nullptr);
  }
  SymbolStream s2;
  uint64_t arcs = 0;
  {
    Estimator entropy(pre.tables);
    arcs = bvcomp_graph(src, n_nodes, p, entropy, chunk_nodes, threads, s2);
    ModelBuilder mb;
    mb.accumulate_host(s2.comps.data(), s2.vals.data(), s2.size());
    mb.build(pre.tables, nullptr, nullptr);
  }
  EncodeResult enc;
  ans_encode(pre.tables, s2.comps.data(), s2.vals.data(), s2.size(), enc);
  pre.stream.swap(enc.stream);
  pre.state = enc.state;
  pre.number_of_nodes = n_nodes;
  pre.compression_window = p.window;
  pre.min_interval_length = p.min_interval_length;
  pre.number_of_arcs = arcs;
  if (enc.phases.states.size() != n_nodes) throw Error(WGA_E_ARG, "internal: phases != nodes");
  store_states(with_ext(new_basename, "states"), enc.phases.states);  // random_access.rs:202-204
  uint64_t upper = enc.phases.pointers.empty() ? 0 : enc.phases.pointers.back();
  EliasFano ef = EliasFano::build(enc.phases.pointers.data(), n_nodes, upper + 1);  // :225-236
  write_whole_file(with_ext(new_basename, "pointers"), ef.serialize());
  store_prelude(with_ext(new_basename, "ans"), pre);  // :217-220
}

}  // namespace wga

using namespace wga;

struct wga_symbols {
  SymbolStream s;
};

extern "C" {

const char* wga_last_error(void) { return last_error_cstr(); }

int wga_cuda_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0;
}

uint64_t wga_kernel_launches(void) { return g_kernel_launches.load(); }

// ------------------------------------------------------------------------------------------ load
int wga_open(const char* basename, int flags, wga_graph** out) {
  return guarded([&] {
    if (!basename || !out) throw Error(WGA_E_ARG, "null argument");
    std::unique_ptr<wga_graph> g(new wga_graph());
    load_prelude(with_ext(basename, "ans"), g->prelude);  // random_access.rs:58-59
    if ((flags & WGA_OPEN_SEQUENTIAL) && !file_exists(with_ext(basename, "pointers")) &&
        !file_exists(with_ext(basename, "states"))) {
      // ANSBvGraphSeq::load (sequential.rs:29-51): only the .ans exists; the phases come from one walk of the stream
      bootstrap_phases(g->prelude, pack_tables(g->prelude.tables), g->phases);
      const uint64_t n = g->prelude.number_of_nodes, u = n ? g->phases.pointers.back() + 1 : 1;
      const uint64_t l = (n && u >= n) ? (uint64_t)(63 - __builtin_clzll(u / n)) : 0;
      g->pointers_payload_bytes = (n * l + 7) / 8 + (n + (u >> l) + 1 + 7) / 8;
    } else {
      EliasFano ef = EliasFano::deserialize(read_whole_file(with_ext(basename, "pointers")));  // :62-63
      ef.expand(g->phases.pointers);
      g->pointers_payload_bytes = ef.payload_bytes();
      load_states(with_ext(basename, "states"), g->phases.states);  // :66-67
    }
    uint64_t n = g->prelude.number_of_nodes;
    *out = finish_open(std::move(g), 0, n, flags);
  });
}

int wga_open_shard(const char* basename, uint64_t first, uint64_t last, int flags, wga_graph** out) {
  return guarded([&] {
    if (!basename || !out) throw Error(WGA_E_ARG, "null argument");
    std::unique_ptr<wga_graph> g(new wga_graph());
    load_prelude(with_ext(basename, "ans"), g->prelude);
    EliasFano ef = EliasFano::deserialize(read_whole_file(with_ext(basename, "pointers")));
    ef.expand(g->phases.pointers);
    g->pointers_payload_bytes = ef.payload_bytes();
    load_states(with_ext(basename, "states"), g->phases.states);
    if (first > last || last > g->prelude.number_of_nodes) throw Error(WGA_E_ARG, "bad shard range");
    *out = finish_open(std::move(g), first, last, flags);
  });
}

int wga_open_mem(const wga_prelude_view* v, int flags, wga_graph** out) {
  return guarded([&] {
    if (!v || !out) throw Error(WGA_E_ARG, "null argument");
    std::unique_ptr<wga_graph> g(new wga_graph());
    view_to_models(v->tables, g->prelude.tables);
    g->prelude.stream.assign(v->stream, v->stream + v->stream_len);
    g->prelude.state = v->state;
    g->prelude.number_of_nodes = v->number_of_nodes;
    g->prelude.compression_window = v->compression_window;
    g->prelude.min_interval_length = v->min_interval_length;
    g->prelude.number_of_arcs = v->number_of_arcs;
    if (v->states && v->pointers) {
      g->phases.states.assign(v->states, v->states + v->number_of_nodes);
      g->phases.pointers.assign(v->pointers, v->pointers + v->number_of_nodes);
      // size the .pointers payload as the Elias-Fano the reference would store (random_access.rs:225-236)
      uint64_t n = v->number_of_nodes, u = n ? g->phases.pointers.back() + 1 : 1;
      uint64_t l = (n && u >= n) ? (uint64_t)(63 - __builtin_clzll(u / n)) : 0;
      g->pointers_payload_bytes = (n * l + 7) / 8 + (n + (u >> l) + 1 + 7) / 8;
    }
    uint64_t n = g->prelude.number_of_nodes;
    *out = finish_open(std::move(g), 0, n, flags);
  });
}

void wga_close(wga_graph* g) { delete g; }

uint64_t wga_num_nodes(const wga_graph* g) { return g->prelude.number_of_nodes; }
uint64_t wga_num_arcs(const wga_graph* g) { return g->prelude.number_of_arcs; }
uint64_t wga_window(const wga_graph* g) { return g->prelude.compression_window; }
uint64_t wga_min_interval_length(const wga_graph* g) { return g->prelude.min_interval_length; }
uint64_t wga_stream_len(const wga_graph* g) { return g->prelude.stream.size(); }
uint64_t wga_compressed_bytes(const wga_graph* g) {
  uint64_t n = g->res_last - g->res_first;
  uint64_t N = g->prelude.number_of_nodes;
  uint64_t ptr_bytes = N ? g->pointers_payload_bytes * n / N : 0;
  uint64_t words = g->on_device ? g->stream_words : g->prelude.stream.size();
  return 2 * words + 4 * n + ptr_bytes;
}

int wga_prelude(const wga_graph* g, wga_prelude_view* o) {
  return guarded([&] {
    if (!g || !o) throw Error(WGA_E_ARG, "null argument");
    for (int c = 0; c < WGA_COMPONENTS; ++c) {
      const ComponentModel& t = g->prelude.tables[c];
      o->tables[c] = {t.table.data(), t.table.size(), t.frame_size, t.radix, t.fidelity, t.folding_threshold,
                      t.folding_offset};
    }
    o->stream = g->prelude.stream.data();
    o->stream_len = g->prelude.stream.size();
    o->state = g->prelude.state;
    o->number_of_nodes = g->prelude.number_of_nodes;
    o->compression_window = g->prelude.compression_window;
    o->min_interval_length = g->prelude.min_interval_length;
    o->number_of_arcs = g->prelude.number_of_arcs;
    o->states = g->phases.states.data();
    o->pointers = g->phases.pointers.data();
  });
}

// ---------------------------------------------------------------------------------------- decode
uint64_t wga_decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last) {
  if (!g || first > last) return 0;
  return decode_workspace_size(g, first, last);
}

int wga_decode_range(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, uint32_t* d_succ,
                     uint64_t succ_capacity, void* d_workspace, uint64_t workspace_bytes, uint64_t* h_arcs,
                     void* stream) {
  return guarded([&] {
    if (!g || !d_offsets || !d_workspace) throw Error(WGA_E_ARG, "null argument");
    decode_range(g, first, last, d_offsets, d_succ, succ_capacity, d_workspace, workspace_bytes, h_arcs,
                 (cudaStream_t)stream);
  });
}

int wga_outdegrees(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, void* d_workspace,
                   uint64_t workspace_bytes, void* stream) {
  return guarded([&] {
    if (!g || !d_offsets || !d_workspace) throw Error(WGA_E_ARG, "null argument");
    outdegrees(g, first, last, d_offsets, d_workspace, workspace_bytes, (cudaStream_t)stream);
  });
}

// Host-buffer entry point.  Ranges larger than one chunk are pipelined: the node range is cut into chunks of
// e2e_chunk_nodes nodes; chunk i is decoded on s_dec into one of two device buffers while the results of chunk
// i-1 travel to the host on s_down (and, after wga_upload(g, NULL), while the inputs of later chunks are still
// arriving on s_up).  PCIe is full duplex, so the step costs about max(download, upload + decode).
static void decode_range_host_pipelined(wga_graph* g, uint64_t first, uint64_t last, uint64_t* h_offsets,
                                        uint32_t* h_succ, uint64_t succ_capacity, uint64_t* h_arcs) {
  g->ensure_pipeline();
  const uint64_t CH = g->e2e_chunk_nodes;
  const uint64_t N = g->prelude.number_of_nodes;
  const double avg = N ? (double)g->prelude.number_of_arcs / (double)N : 0.0;
  // first guess of a chunk's arcs; a denser chunk makes its decode fail with "need N" and is retried with larger buffers
  uint64_t chunk_cap = std::max<uint64_t>(g->pipe_succ_n, (uint64_t)(avg * (double)CH * 1.5) + (4u << 20));
  auto ensure_buffers = [&](uint64_t cap) {
    const uint64_t ws_bytes = decode_workspace_size(g, 0, CH) + 8 * cap;  // record buffer sized for the chunk buffer
    if (g->e2e_ws_bytes < ws_bytes) {
      if (g->e2e_ws) cudaFree(g->e2e_ws);
      g->e2e_ws = nullptr; g->e2e_ws_bytes = 0;
      WGA_CUDA(cudaMalloc(&g->e2e_ws, ws_bytes));
      g->e2e_ws_bytes = ws_bytes;
    }
    if (g->pipe_off_n < CH + 1 || g->pipe_succ_n < cap) {
      for (int i = 0; i < 2; ++i) {
        if (g->pipe_off[i]) cudaFree(g->pipe_off[i]);
        if (g->pipe_succ[i]) cudaFree(g->pipe_succ[i]);
        g->pipe_off[i] = nullptr; g->pipe_succ[i] = nullptr;
      }
      g->pipe_off_n = g->pipe_succ_n = 0;
      for (int i = 0; i < 2; ++i) {
        WGA_CUDA(cudaMalloc((void**)&g->pipe_off[i], (CH + 1) * 8));
        WGA_CUDA(cudaMalloc((void**)&g->pipe_succ[i], cap * 4));
      }
      g->pipe_off_n = CH + 1; g->pipe_succ_n = cap;
    }
  };
  ensure_buffers(chunk_cap);
  auto drain = [&]() {  // nothing of this call may still be in flight when it returns, also on errors
    cudaStreamSynchronize(g->s_dec);
    cudaStreamSynchronize(g->s_down);
    g->up_pending = false;
  };
  try {
    uint64_t base = 0;
    int i = 0;
    for (uint64_t a = first; a < last; a += CH, ++i) {
      const uint64_t b = std::min(last, a + CH);
      const int j = i & 1;
      if (g->up_pending) {  // inputs of this chunk (and of the halo just before it) must have arrived
        const uint64_t c1 = (b - 1 - g->res_first) / CH, c0 = a > g->res_first ? (a - 1 - g->res_first) / CH : 0;
        for (uint64_t c = c0; c <= c1 && c < g->up_ev.size(); ++c) WGA_CUDA(cudaStreamWaitEvent(g->s_dec, g->up_ev[c], 0));
      }
      if (i >= 2) WGA_CUDA(cudaStreamWaitEvent(g->s_dec, g->down_done[j], 0));  // buffer j is free again
      uint64_t arcs = 0;
      for (int attempt = 0;; ++attempt) {
        try {
          decode_range(g, a, b, g->pipe_off[j], g->pipe_succ[j], g->pipe_succ_n, g->e2e_ws, g->e2e_ws_bytes, &arcs, g->s_dec);
          break;
        } catch (const Error& e) {
          if (e.code != WGA_E_WORKSPACE || attempt >= 2) throw;
          // a chunk denser than the guess: both chunk buffers must be idle before they are replaced
          WGA_CUDA(cudaStreamSynchronize(g->s_dec));
          WGA_CUDA(cudaStreamSynchronize(g->s_down));
          ensure_buffers(std::max<uint64_t>(2 * g->pipe_succ_n, g->last_need_succ + (1u << 20)));
        }
      }
      if (base + arcs > succ_capacity && h_succ)
        throw Error(WGA_E_WORKSPACE, "h_succ too small: need more than " + std::to_string(base + arcs) + " elements");
      launch_offsets_add(g->pipe_off[j], b - a + 1, base, g->s_dec);
      WGA_CUDA(cudaEventRecord(g->dec_done[j], g->s_dec));
      WGA_CUDA(cudaStreamWaitEvent(g->s_down, g->dec_done[j], 0));
      // offsets: the last entry of a chunk equals the first of the next one
      WGA_CUDA(cudaMemcpyAsync(h_offsets + (a - first), g->pipe_off[j], (b - a + (b == last ? 1 : 0)) * 8,
                               cudaMemcpyDeviceToHost, g->s_down));
      if (arcs && h_succ)
        WGA_CUDA(cudaMemcpyAsync(h_succ + base, g->pipe_succ[j], arcs * 4, cudaMemcpyDeviceToHost, g->s_down));
      WGA_CUDA(cudaEventRecord(g->down_done[j], g->s_down));
      base += arcs;
    }
    WGA_CUDA(cudaStreamSynchronize(g->s_down));
    g->up_pending = false;
    if (h_arcs) *h_arcs = base;
  } catch (...) {
    drain();
    throw;
  }
}

int wga_decode_range_host(wga_graph* g, uint64_t first, uint64_t last, uint64_t* h_offsets, uint32_t* h_succ,
                          uint64_t succ_capacity, uint64_t* h_arcs) {
  return guarded([&] {
    if (!g || !h_offsets) throw Error(WGA_E_ARG, "null argument");
    if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
    if (first > last || last > g->res_last || first < g->res_first) throw Error(WGA_E_ARG, "range outside the resident nodes");
    g->e2e_chunk_nodes = std::max<uint64_t>(1, tuning_e2e_chunk());
    // chunks only pay off when no record is long enough to dominate a chunk (see wga_graph::longest_record)
    if (last - first > g->e2e_chunk_nodes && g->longest_record() < 8192) {
      decode_range_host_pipelined(g, first, last, h_offsets, h_succ, succ_capacity, h_arcs);
      return;
    }
    if (g->up_pending) {  // a chunked upload is in flight: wait for it
      WGA_CUDA(cudaStreamSynchronize(g->s_up));
      g->up_pending = false;
    }
    // grow-only device buffers owned by the handle: no allocation on the steady-state path
    uint64_t ws_bytes = decode_workspace_size(g, first, last);
    if (g->e2e_ws_bytes < ws_bytes) {
      if (g->e2e_ws) cudaFree(g->e2e_ws);
      g->e2e_ws = nullptr; g->e2e_ws_bytes = 0;
      WGA_CUDA(cudaMalloc(&g->e2e_ws, ws_bytes));
      g->e2e_ws_bytes = ws_bytes;
    }
    uint64_t n_off = last - first + 1;
    if (g->e2e_off_n < n_off) {
      if (g->e2e_off) cudaFree(g->e2e_off);
      g->e2e_off = nullptr; g->e2e_off_n = 0;
      WGA_CUDA(cudaMalloc((void**)&g->e2e_off, n_off * 8));
      g->e2e_off_n = n_off;
    }
    uint64_t cap = succ_capacity ? succ_capacity : 1;
    if (g->e2e_succ_n < cap) {
      if (g->e2e_succ) cudaFree(g->e2e_succ);
      g->e2e_succ = nullptr; g->e2e_succ_n = 0;
      WGA_CUDA(cudaMalloc((void**)&g->e2e_succ, cap * 4));
      g->e2e_succ_n = cap;
    }
    uint64_t arcs = 0;
    decode_range(g, first, last, g->e2e_off, g->e2e_succ, succ_capacity, g->e2e_ws, g->e2e_ws_bytes, &arcs, 0);
    WGA_CUDA(cudaMemcpyAsync(h_offsets, g->e2e_off, n_off * 8, cudaMemcpyDeviceToHost, 0));
    if (arcs && h_succ) WGA_CUDA(cudaMemcpyAsync(h_succ, g->e2e_succ, arcs * 4, cudaMemcpyDeviceToHost, 0));
    WGA_CUDA(cudaStreamSynchronize(0));
    if (h_arcs) *h_arcs = arcs;
  });
}

// stream == NULL: chunked upload on the handle's own stream, overlapped by a following wga_decode_range_host
int wga_upload(wga_graph* g, void* stream) {
  return guarded([&] {
    if (!g) throw Error(WGA_E_ARG, "null argument");
    g->e2e_chunk_nodes = std::max<uint64_t>(1, tuning_e2e_chunk());
    if (stream) g->reupload((cudaStream_t)stream);
    else if (g->longest_record() < 8192) g->reupload_chunked();
    else g->reupload(0);
  });
}

uint64_t wga_upload_bytes(const wga_graph* g) {
  uint64_t n = g->res_last - g->res_first;
  return 2 * g->stream_words + 4 * n + 8 * n;
}

uint64_t wga_last_halo_nodes(const wga_graph* g) { return g ? g->last_halo_nodes : 0; }

int wga_set_profiling(wga_graph* g, int on) {
  if (!g) return WGA_E_ARG;
  g->profiling = on != 0;
  return WGA_OK;
}

int wga_last_profile(const wga_graph* g, float* h_stage_ms8) {
  if (!g || !h_stage_ms8) return WGA_E_ARG;
  for (int i = 0; i < 8; ++i) h_stage_ms8[i] = g->stage_ms[i];
  return g->n_ev;
}

// --------------------------------------------------------------------------------- random access
uint64_t wga_successors_workspace_size(const wga_graph* g, uint64_t n_queries, uint64_t max_total_arcs) {
  if (!g) return 0;
  return successors_workspace_size(g, n_queries, max_total_arcs);
}

int wga_successors_batch(wga_graph* g, const uint64_t* d_nodes, uint64_t n_queries, uint64_t* d_offsets,
                         uint32_t* d_succ, uint64_t succ_capacity, void* d_workspace, uint64_t workspace_bytes,
                         uint64_t* h_arcs, void* stream) {
  return guarded([&] {
    if (!g || (!d_nodes && n_queries) || !d_offsets || (!d_workspace && n_queries)) throw Error(WGA_E_ARG, "null argument");
    successors_batch(g, d_nodes, n_queries, d_offsets, d_succ, succ_capacity, d_workspace, workspace_bytes, h_arcs,
                     (cudaStream_t)stream);
  });
}

int wga_successors_batch_host(wga_graph* g, const uint64_t* h_nodes, uint64_t n_queries, uint64_t* h_offsets,
                              uint32_t* h_succ, uint64_t succ_capacity, uint64_t* h_arcs) {
  return guarded([&] {
    if (!g || (!h_nodes && n_queries) || !h_offsets) throw Error(WGA_E_ARG, "null argument");
    if (!g->on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
    if (g->up_pending) {
      WGA_CUDA(cudaStreamSynchronize(g->s_up));
      g->up_pending = false;
    }
    // grow-only device buffers of the handle: queries + offsets share e2e_off, the lists use e2e_succ, e2e_ws
    const uint64_t n_off = 2 * n_queries + 2;
    if (g->e2e_off_n < n_off) {
      if (g->e2e_off) cudaFree(g->e2e_off);
      g->e2e_off = nullptr; g->e2e_off_n = 0;
      WGA_CUDA(cudaMalloc((void**)&g->e2e_off, n_off * 8));
      g->e2e_off_n = n_off;
    }
    uint64_t* d_q = g->e2e_off;
    uint64_t* d_off = g->e2e_off + n_queries + 1;
    auto need_ws = [&](uint64_t bytes) {
      if (g->e2e_ws_bytes < bytes) {
        if (g->e2e_ws) cudaFree(g->e2e_ws);
        g->e2e_ws = nullptr; g->e2e_ws_bytes = 0;
        WGA_CUDA(cudaMalloc(&g->e2e_ws, bytes));
        g->e2e_ws_bytes = bytes;
      }
    };
    if (n_queries) WGA_CUDA(cudaMemcpyAsync(d_q, h_nodes, n_queries * 8, cudaMemcpyHostToDevice, 0));
    uint64_t arcs = 0;
    need_ws(successors_workspace_size(g, n_queries, 0));
    successors_batch(g, d_q, n_queries, d_off, nullptr, 0, g->e2e_ws, g->e2e_ws_bytes, &arcs, 0);  // sizing call
    if (h_arcs) *h_arcs = arcs;
    const bool want = h_succ && succ_capacity >= arcs;
    if (want && arcs) {
      if (g->e2e_succ_n < arcs + 1024) {
        if (g->e2e_succ) cudaFree(g->e2e_succ);
        g->e2e_succ = nullptr; g->e2e_succ_n = 0;
        WGA_CUDA(cudaMalloc((void**)&g->e2e_succ, (arcs + 1024) * 4));
        g->e2e_succ_n = arcs + 1024;
      }
      need_ws(successors_workspace_size(g, n_queries, arcs));
      successors_batch(g, d_q, n_queries, d_off, g->e2e_succ, g->e2e_succ_n, g->e2e_ws, g->e2e_ws_bytes, &arcs, 0);
      WGA_CUDA(cudaMemcpyAsync(h_succ, g->e2e_succ, arcs * 4, cudaMemcpyDeviceToHost, 0));
    }
    WGA_CUDA(cudaMemcpyAsync(h_offsets, d_off, (n_queries + 1) * 8, cudaMemcpyDeviceToHost, 0));
    WGA_CUDA(cudaStreamSynchronize(0));
    if (h_succ && !want) throw Error(WGA_E_WORKSPACE, "h_succ too small: need " + std::to_string(arcs) + " elements");
  });
}

// ----------------------------------------------------------------------------------------- debug
int wga_debug_expand_table(wga_graph* g, int component, void* h_out, uint64_t n_slots) {
  return guarded([&] {
    if (!g || !h_out) throw Error(WGA_E_ARG, "null argument");
    debug_expand_table(g, component, h_out, n_slots);
  });
}

int wga_debug_decode_symbols(wga_graph* g, const uint8_t* h_components, uint64_t n, uint64_t ptr, uint32_t state,
                             uint64_t* h_out, uint64_t* h_end_ptr, uint32_t* h_end_state) {
  return guarded([&] {
    if (!g || (!h_components && n) || (!h_out && n)) throw Error(WGA_E_ARG, "null argument");
    debug_decode_symbols(g, h_components, n, ptr, state, h_out, h_end_ptr, h_end_state);
  });
}

int wga_debug_set_tuning(const char* key, uint64_t value) {
  int rc = set_tuning(key, value);
  if (rc != WGA_OK) set_last_error("unknown tuning key");
  return rc;
}

// ----------------------------------------------------------------------------------------- model
int wga_model_create(wga_model** out) {
  return guarded([&] {
    if (!out) throw Error(WGA_E_ARG, "null argument");
    *out = reinterpret_cast<wga_model*>(new ModelBuilder());
  });
}
void wga_model_destroy(wga_model* m) { delete reinterpret_cast<ModelBuilder*>(m); }
uint64_t* wga_model_bins(wga_model* m) { return reinterpret_cast<ModelBuilder*>(m)->device_bins(); }
int wga_model_accumulate(wga_model* m, const uint8_t* d_components, const uint64_t* d_symbols, uint64_t n,
                         void* stream) {
  return guarded([&] { reinterpret_cast<ModelBuilder*>(m)->accumulate_device(d_components, d_symbols, n, (cudaStream_t)stream); });
}
int wga_model_accumulate_host(wga_model* m, const uint8_t* h_components, const uint64_t* h_symbols, uint64_t n) {
  return guarded([&] { reinterpret_cast<ModelBuilder*>(m)->accumulate_host(h_components, h_symbols, n); });
}
uint64_t wga_model_sparse_count(wga_model* m) {
  uint64_t n = 0;
  guarded([&] { n = reinterpret_cast<ModelBuilder*>(m)->sparse_count(); });
  return n;
}
int wga_model_sparse_export(wga_model* m, uint8_t* h_components, uint64_t* h_symbols, uint64_t* h_counts) {
  return guarded([&] { reinterpret_cast<ModelBuilder*>(m)->sparse_export(h_components, h_symbols, h_counts); });
}
int wga_model_sparse_merge(wga_model* m, const uint8_t* h_components, const uint64_t* h_symbols,
                           const uint64_t* h_counts, uint64_t n) {
  return guarded([&] { reinterpret_cast<ModelBuilder*>(m)->sparse_merge(h_components, h_symbols, h_counts, n); });
}
int wga_model_build(wga_model* m, wga_component_model out_tables[WGA_COMPONENTS], double* h_original_cost9,
                    double* h_final_cost9) {
  return guarded([&] {
    ModelBuilder* mb = reinterpret_cast<ModelBuilder*>(m);
    mb->build(mb->result, h_original_cost9, h_final_cost9);
    for (int c = 0; c < WGA_COMPONENTS; ++c) {
      const ComponentModel& t = mb->result[c];
      out_tables[c] = {t.table.data(), t.table.size(), t.frame_size, t.radix, t.fidelity, t.folding_threshold,
                       t.folding_offset};
    }
  });
}

// ---------------------------------------------------------------------------------------- bvcomp
int wga_bvcomp_symbols_range(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                             uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                             const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                             wga_symbols** out) {
  return guarded([&] {
    if (!h_offsets || !out) throw Error(WGA_E_ARG, "null argument");
    BvCompParams p{compression_window, max_ref_count, min_interval_length};
    std::unique_ptr<wga_symbols> s(new wga_symbols());
    // the CSR arrays hold nodes [first_node, first_node + n_nodes) only; successor ids are global
    NodeSource src = [&](uint64_t v, std::vector<uint64_t>& o) {
      o.assign(h_succ + h_offsets[v - first_node], h_succ + h_offsets[v - first_node + 1]);
    };
    if (estimator_tables) {
      ComponentModel m[WGA_COMPONENTS];
      view_to_models(estimator_tables, m);
      Estimator est(m);
      bvcomp_nodes(src, first_node, first_node + n_nodes, p, est, chunk_nodes, threads, s->s);
    } else {
      bvcomp_nodes(src, first_node, first_node + n_nodes, p, Estimator(), chunk_nodes, threads, s->s);
    }
    *out = s.release();
  });
}
int wga_bvcomp_symbols_range_gpu(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                                 uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                                 const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                                 wga_symbols** out) {
  return guarded([&] {
    if (!h_offsets || !out) throw Error(WGA_E_ARG, "null argument");
    BvCompParams p{compression_window, max_ref_count, min_interval_length};
    std::unique_ptr<wga_symbols> s(new wga_symbols());
    NodeSource src = [&](uint64_t v, std::vector<uint64_t>& o) {
      o.assign(h_succ + h_offsets[v - first_node], h_succ + h_offsets[v - first_node + 1]);
    };
    std::unique_ptr<Estimator> est;
    if (estimator_tables) {
      ComponentModel m[WGA_COMPONENTS];
      view_to_models(estimator_tables, m);
      est.reset(new Estimator(m));
    } else est.reset(new Estimator());
    std::vector<uint16_t> choice;
    bvcomp_choose_gpu(h_offsets, h_succ, first_node, n_nodes, p, *est, chunk_nodes, choice, nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    bvcomp_nodes(src, first_node, first_node + n_nodes, p, *est, chunk_nodes, threads, s->s,
                 compression_window ? choice.data() : nullptr);
    if (getenv("WGA_TIMING"))
      fprintf(stderr, "[wga] host emission of the chosen records: %.3f s\n",
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    *out = s.release();
  });
}

int wga_debug_bvcomp_costs(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                           uint64_t compression_window, uint64_t min_interval_length,
                           const wga_component_model* estimator_tables, uint64_t chunk_nodes, int use_gpu,
                           uint64_t* h_costs) {
  return guarded([&] {
    if (!h_offsets || !h_costs) throw Error(WGA_E_ARG, "null argument");
    BvCompParams p{compression_window, 3, min_interval_length};
    std::unique_ptr<Estimator> est;
    if (estimator_tables) {
      ComponentModel m[WGA_COMPONENTS];
      view_to_models(estimator_tables, m);
      est.reset(new Estimator(m));
    } else est.reset(new Estimator());
    std::vector<uint64_t> costs;
    if (use_gpu) {
      std::vector<uint16_t> choice;
      bvcomp_choose_gpu(h_offsets, h_succ, first_node, n_nodes, p, *est, chunk_nodes, choice, &costs);
    } else {
      NodeSource src = [&](uint64_t v, std::vector<uint64_t>& o) {
        o.assign(h_succ + h_offsets[v - first_node], h_succ + h_offsets[v - first_node + 1]);
      };
      bvcomp_costs_host(src, first_node, n_nodes, p, *est, chunk_nodes, costs);
    }
    std::memcpy(h_costs, costs.data(), costs.size() * 8);
  });
}

int wga_bvcomp_symbols(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t n_nodes,
                       uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                       const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                       wga_symbols** out) {
  return wga_bvcomp_symbols_range(h_offsets, h_succ, 0, n_nodes, compression_window, max_ref_count, min_interval_length,
                                  estimator_tables, chunk_nodes, threads, out);
}
uint64_t wga_symbols_len(const wga_symbols* s) { return s->s.size(); }
const uint8_t* wga_symbols_components(const wga_symbols* s) { return s->s.comps.data(); }
const uint64_t* wga_symbols_values(const wga_symbols* s) { return s->s.vals.data(); }
void wga_symbols_free(wga_symbols* s) { delete s; }

int wga_ans_encode(const wga_component_model tables[WGA_COMPONENTS], const uint8_t* h_components,
                   const uint64_t* h_symbols, uint64_t n, uint16_t** out_stream, uint64_t* out_stream_len,
                   uint32_t* out_state, uint32_t** out_states, uint64_t** out_pointers, uint64_t* out_n_phases) {
  return guarded([&] {
    if (!tables || !out_stream || !out_stream_len || !out_state) throw Error(WGA_E_ARG, "null argument");
    ComponentModel m[WGA_COMPONENTS];
    view_to_models(tables, m);
    EncodeResult r;
    ans_encode(m, h_components, h_symbols, n, r);
    *out_stream = (uint16_t*)std::malloc(r.stream.size() * 2 + 2);
    std::memcpy(*out_stream, r.stream.data(), r.stream.size() * 2);
    *out_stream_len = r.stream.size();
    *out_state = r.state;
    if (out_states && out_pointers && out_n_phases) {
      size_t np = r.phases.states.size();
      *out_states = (uint32_t*)std::malloc(np * 4 + 4);
      *out_pointers = (uint64_t*)std::malloc(np * 8 + 8);
      std::memcpy(*out_states, r.phases.states.data(), np * 4);
      std::memcpy(*out_pointers, r.phases.pointers.data(), np * 8);
      *out_n_phases = np;
    }
  });
}
void wga_free(void* p) { std::free(p); }

int wga_write_files(const char* basename, const wga_prelude_view* v) {
  return guarded([&] {
    if (!basename || !v) throw Error(WGA_E_ARG, "null argument");
    Prelude p;
    view_to_models(v->tables, p.tables);
    p.stream.assign(v->stream, v->stream + v->stream_len);
    p.state = v->state;
    p.number_of_nodes = v->number_of_nodes;
    p.compression_window = v->compression_window;
    p.min_interval_length = v->min_interval_length;
    p.number_of_arcs = v->number_of_arcs;
    if (v->states && v->pointers) {
      std::vector<uint32_t> st(v->states, v->states + v->number_of_nodes);
      store_states(with_ext(basename, "states"), st);
      uint64_t upper = v->number_of_nodes ? v->pointers[v->number_of_nodes - 1] : 0;
      EliasFano ef = EliasFano::build(v->pointers, v->number_of_nodes, upper + 1);
      write_whole_file(with_ext(basename, "pointers"), ef.serialize());
    }
    store_prelude(with_ext(basename, "ans"), p);
  });
}

int wga_store(const char* basename, const char* new_basename, uint64_t compression_window,
              uint64_t max_ref_count, uint64_t min_interval_length) {
  return guarded([&] {
    if (!basename) throw Error(WGA_E_ARG, "null basename");
    // BvGraphSeq::with_basename(..).load() (random_access.rs:101-103); the CSR is kept in host memory
    std::vector<uint64_t> offs(1, 0);
    std::vector<uint32_t> succ;
    read_bvgraph(basename, [&](uint64_t, const std::vector<uint64_t>& s) {
      for (uint64_t x : s) succ.push_back((uint32_t)x);
      offs.push_back(succ.size());
    });
    NodeSource src = [&](uint64_t v, std::vector<uint64_t>& o) {
      o.assign(succ.begin() + offs[v], succ.begin() + offs[v + 1]);
    };
    store_from_source(src, offs.size() - 1, new_basename,
                      BvCompParams{compression_window, max_ref_count, min_interval_length}, 0, 1);
  });
}

int wga_store_csr(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t n_nodes, const char* new_basename,
                  uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                  uint64_t chunk_nodes, int threads) {
  return guarded([&] {
    if (!h_offsets) throw Error(WGA_E_ARG, "null argument");
    NodeSource src = [&](uint64_t v, std::vector<uint64_t>& o) {
      o.assign(h_succ + h_offsets[v], h_succ + h_offsets[v + 1]);
    };
    store_from_source(src, n_nodes, new_basename,
                      BvCompParams{compression_window, max_ref_count, min_interval_length}, chunk_nodes, threads);
  });
}

int wga_bvgraph_read(const char* basename, uint64_t* n_nodes, uint64_t* n_arcs, uint64_t* h_offsets,
                     uint32_t* h_succ) {
  return guarded([&] {
    if (!basename) throw Error(WGA_E_ARG, "null basename");
    uint64_t nodes = 0, arcs = 0;
    if (h_offsets) h_offsets[0] = 0;
    read_bvgraph(basename, [&](uint64_t v, const std::vector<uint64_t>& s) {
      if (h_succ)
        for (size_t i = 0; i < s.size(); ++i) h_succ[arcs + i] = (uint32_t)s[i];
      arcs += s.size();
      nodes = v + 1;
      if (h_offsets) h_offsets[v + 1] = arcs;
    });
    if (n_nodes) *n_nodes = nodes;
    if (n_arcs) *n_arcs = arcs;
  });
}

int wga_ef_write(const char* path, const uint64_t* values, uint64_t n, uint64_t u) {
  return guarded([&] {
    if (!path || (!values && n)) throw Error(WGA_E_ARG, "null argument");
    write_whole_file(path, EliasFano::build(values, n, u).serialize());
  });
}

int wga_ef_read(const char* path, uint64_t* n, uint64_t* h_values) {
  return guarded([&] {
    if (!path) throw Error(WGA_E_ARG, "null path");
    EliasFano ef = EliasFano::deserialize(read_whole_file(path));
    if (n) *n = ef.n;
    if (h_values) {
      std::vector<uint64_t> v;
      ef.expand(v);
      std::memcpy(h_values, v.data(), v.size() * 8);
      // exercise the inventory path on a sample (IndexedSeq::get)
      for (uint64_t i = 0; i < ef.n; i += 997)
        if (ef.get(i) != v[i]) throw Error(WGA_E_FORMAT, "Elias-Fano: inventory select disagrees with the scan");
    }
  });
}

int wga_synth_graph(int kind, uint64_t n_nodes, double mean_degree, uint64_t seed, uint64_t first, uint64_t last,
                    int threads, uint64_t* h_offsets, uint32_t* h_succ, uint64_t* n_arcs) {
  return guarded([&] {
    uint64_t a = synth_graph(kind, n_nodes, mean_degree, seed, first, last, threads, h_offsets, h_succ);
    if (n_arcs) *n_arcs = a;
  });
}

}  // extern "C"
