// The graph handle behind wga_graph*: host copy of the files + device-resident decode inputs.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <vector>

#include "common.hpp"
#include "device.cuh"

namespace wga {

#define WGA_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::wga::Error(WGA_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

// phases of every node from the .ans alone (ANSBvGraphSeq::load, sequential.rs:29-51): see graph.cu
void bootstrap_phases(const Prelude& pre, const PackedTablesData& pk, Phases& out);

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace wga

struct wga_graph {
  wga::Prelude prelude;
  wga::Phases phases;       // full (or shard-local) phases, file order
  wga::PackedTablesData packed;
  // nodes whose inputs are resident: [res_first, res_last); whole graph unless opened as a shard
  uint64_t res_first = 0, res_last = 0;
  bool on_device = false;
  int device = -1;
  // device buffers
  uint16_t* d_stream = nullptr;        // word 0 of the resident span (inside d_stream_alloc, 16 bytes of zero padding in front)
  uint16_t* d_stream_alloc = nullptr;
  uint64_t stream_base = 0, stream_words = 0;
  uint32_t* d_states = nullptr;  // res_last-res_first entries; entry k = node res_last-1-k
  uint64_t* d_ptrs = nullptr;
  uint2* d_bkt = nullptr;
  uint2* d_ent = nullptr;
  uint32_t* d_err = nullptr;  // device error word
  // 64 bytes of mapped pinned host memory: kernels publish the scalars the host needs (halo start, arc totals,
  // deepest level, error word) there, so reading them never queues behind a bulk copy on a DMA engine
  volatile uint64_t* h_pub = nullptr;
  uint64_t* d_pub = nullptr;
  wga::DevGraph dev{};         // view passed to kernels
  uint64_t pointers_payload_bytes = 0;
  bool pinned = false;          // host copies registered with cudaHostRegister (fast re-upload)
  // optional per-stage timing of the last decode_range (bench.py's kernel breakdown)
  bool profiling = false;
  cudaEvent_t ev[8] = {};
  int n_ev = 0;                 // events recorded by the last decode
  float stage_ms[8] = {};
  // grow-only device buffers of the host-buffer entry points
  void* e2e_ws = nullptr; uint64_t e2e_ws_bytes = 0;
  uint64_t* e2e_off = nullptr; uint64_t e2e_off_n = 0;
  uint32_t* e2e_succ = nullptr; uint64_t e2e_succ_n = 0;
  // pipelined host entry point (wga_decode_range_host): node-range chunks flow through upload -> decode ->
  // download on three streams with double-buffered chunk outputs
  uint64_t e2e_chunk_nodes = 1ull << 19;   // tuning knob "e2e_chunk" (tests shrink it)
  cudaStream_t s_up = nullptr, s_dec = nullptr, s_down = nullptr;
  std::vector<cudaEvent_t> up_ev;   // one per upload chunk (chunk c = nodes [res_first + c*CHUNK, ...))
  bool up_pending = false;          // the events above belong to an upload that has not been consumed yet
  cudaEvent_t dec_done[2] = {}, down_done[2] = {};
  uint64_t* pipe_off[2] = {}; uint64_t pipe_off_n = 0;
  uint32_t* pipe_succ[2] = {}; uint64_t pipe_succ_n = 0;
  uint64_t last_halo_nodes = 0;     // halo of the last decode_range
  uint64_t last_need_succ = 0;      // elements the last "d_succ too small" failure asked for
  uint64_t max_record_words = UINT64_MAX;  // largest record of the resident range in stream words (lazy)
  uint64_t longest_record();
  void ensure_pipeline();
  void reupload_chunked();          // H2D in node-range chunks on s_up, one event per chunk
  void reupload_pin();

  ~wga_graph();
  void upload();
  void reupload(cudaStream_t st);  // H2D of stream words, states and pointers only
};
