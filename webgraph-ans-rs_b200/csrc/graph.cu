// Graph handle: packed decoder tables (host build) and upload of the decode inputs to HBM.
#include "graph.hpp"

namespace wga {

static thread_local std::string t_last_error;
void set_last_error(const std::string& m) { t_last_error = m; }
const char* last_error_cstr() { return t_last_error.c_str(); }

// Builds the packed tables from the encoder tables of the Prelude.  Same content as the reference's
// ANSModel4Decoder::new (src/ans/models/model4decoder.rs:18-54) + quasi_fold (:56-68), different layout:
// see common.hpp.  `base` = symbol - folding_offset*folds, so the decoder's
// quasi_folded = (base << folds*radix) | folds << 48.
PackedTablesData pack_tables(const ComponentModel tables[WGA_COMPONENTS]) {
  PackedTablesData p;
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    const ComponentModel& t = tables[c];
    if (t.frame_size > 16) throw Error(WGA_E_FORMAT, "frame size > 2^16");
    if (t.radix == 0 || t.radix > 16 || t.fidelity == 0) throw Error(WGA_E_FORMAT, "bad radix/fidelity");
    const uint32_t L = (uint32_t)t.frame_size;
    p.L[c] = (uint8_t)L;
    p.R[c] = (uint8_t)t.radix;
    p.bkt_off[c] = (uint32_t)p.bkt.size();
    p.ent_off[c] = (uint32_t)p.ent.size();
    const uint32_t frame = 1u << L;
    // non-zero symbols in index order; running sum gives the slot ranges (model4decoder.rs:24-43)
    uint32_t nnz = 0;
    uint32_t last_slot = 0;
    std::vector<uint32_t> starts;  // first slot of every nz symbol
    for (size_t sym = 0; sym < t.table.size(); ++sym) {
      const wga_encoder_entry& e = t.table[sym];
      if (e.freq == 0) continue;
      if (last_slot + e.freq > frame) throw Error(WGA_E_FORMAT, "frequencies exceed the frame");
      uint32_t folds = 0, base = (uint32_t)sym;
      if ((uint16_t)sym >= (uint16_t)t.folding_threshold) {  // quasi_fold, :57
        folds = (uint32_t)(((uint64_t)sym - t.folding_threshold) / t.folding_offset + 1);
        base = (uint32_t)((uint64_t)sym - t.folding_offset * folds);
      }
      if (folds >= 0xFFFF || base > 0xFFFF || folds * t.radix > 48) throw Error(WGA_E_FORMAT, "bad folding parameters");
      Ent en;
      // NOTE the decode formula uses cumul_freq as stored (decoder.rs:65), the slot range uses the running sum
      en.cf = (uint32_t)e.cumul_freq | ((uint32_t)e.freq << 16);
      en.bf = base | (folds << 16);
      if (e.cumul_freq != (uint16_t)last_slot) throw Error(WGA_E_FORMAT, "cumulative frequency does not match the running sum");
      p.ent.push_back(en);
      starts.push_back(last_slot);
      last_slot += e.freq;
      ++nnz;
    }
    // sentinel: owns every slot >= sum of frequencies (unused slots; the reference leaves default entries)
    Ent s;
    s.cf = 0xFFFFu << 16;
    s.bf = 0xFFFFu << 16;
    p.ent.push_back(s);
    starts.push_back(last_slot);
    p.nnz[c] = nnz;
    p.nent[c] = nnz + 1;
    // buckets of 32 slots: owner of the first slot + mask of the slots (1..31) where a symbol's range starts
    const uint32_t nb = frame > 32 ? frame / 32 : 1;
    p.nb[c] = nb;
    uint32_t j = 0;  // owner of the current slot
    for (uint32_t b = 0; b < nb; ++b) {
      const uint32_t s0 = b * 32;
      while (j < nnz && starts[j + 1] <= s0) ++j;
      Bkt k;
      k.j0 = j;
      k.mask = 0;
      uint32_t q = j;
      for (uint32_t sl = 1; sl < 32 && s0 + sl < frame; ++sl) {
        if (q < nnz && starts[q + 1] == s0 + sl) { k.mask |= 1u << sl; ++q; }
      }
      p.bkt.push_back(k);
    }
  }
  return p;
}

// ---- sequential bootstrap: phases from the .ans alone -----------------------------------------------------
// ANSBvGraphSeq::load needs only the .ans (src/bvgraph/sequential.rs:29-51): its decoder starts at
// (stream.len(), prelude.state) (bvgraphseq_decoder_factory.rs:29-35) and walks the one ANS state chain node by
// node.  The GPU decode starts every node from its own phase, so when .pointers / .states are missing the phases
// are recovered here, at load time, by the same walk: every symbol of every record is decoded once in record order
// (host, serial by construction of the format) and the decoder (state, pointer) is noted before each node.
namespace {
struct HostDecoder {
  const PackedTablesData& p;
  const uint16_t* stream;
  uint64_t ptr;
  uint32_t state;
  bool ok = true;
  void extend() {  // decoder.rs:89-93
    if (ptr == 0) { ok = false; return; }
    --ptr;
    state = (state << 16) | stream[ptr];
  }
  uint64_t decode(int c) {  // decoder.rs:58-87 on the packed tables (same lookup as the device code)
    const uint32_t L = p.L[c], R = p.R[c];
    const uint32_t slot = state & ((1u << L) - 1u);
    const Bkt& bk = p.bkt[p.bkt_off[c] + (slot >> 5)];
    const uint32_t j = bk.j0 + (uint32_t)__builtin_popcount(bk.mask & ((2u << (slot & 31u)) - 1u));
    const Ent& e = p.ent[p.ent_off[c] + j];
    const uint32_t folds = e.bf >> 16;
    if (folds == 0xFFFFu) { ok = false; return 0; }
    state = (state >> L) * (e.cf >> 16) + slot - (e.cf & 0xFFFFu);
    if (state < WGA_LOWER_BOUND) extend();
    uint64_t fold = 0;
    for (uint32_t f = 0; f < folds && ok; ++f) {  // decoder.rs:74-85, one trip per fold
      if (state < WGA_LOWER_BOUND) extend();
      fold = (fold << R) | (state & ((1u << R) - 1u));
      state >>= R;
      if (state < WGA_LOWER_BOUND) extend();
    }
    return ((uint64_t)(e.bf & 0xFFFFu) << (folds * R)) | fold;
  }
};
}  // namespace

void bootstrap_phases(const Prelude& pre, const PackedTablesData& pk, Phases& out) {
  const uint64_t N = pre.number_of_nodes;
  const uint64_t window = pre.compression_window, minint = pre.min_interval_length;
  out.states.assign(N, 0);
  out.pointers.assign(N, 0);
  std::vector<uint64_t> outdeg(N, 0);
  HostDecoder dec{pk, pre.stream.data(), pre.stream.size(), pre.state};
  auto fail = [&](uint64_t v) { throw Error(WGA_E_CORRUPT, "sequential bootstrap: inconsistent stream at node " + std::to_string(v)); };
  for (uint64_t v = 0; v < N; ++v) {
    // file order: entry i belongs to node N-1-i (random_access.rs:202, 225-231)
    out.states[N - 1 - v] = dec.state;
    out.pointers[N - 1 - v] = dec.ptr;
    const uint64_t d = dec.decode(Outdegree);
    if (!dec.ok) fail(v);
    outdeg[v] = d;
    if (d == 0) continue;
    uint64_t extras = d;
    if (window != 0) {
      const uint64_t r = dec.decode(ReferenceOffset);
      if (!dec.ok || r > window || r > v) fail(v);
      if (r != 0) {
        const uint64_t dref = outdeg[v - r];
        const uint64_t b = dec.decode(BlockCount);
        if (!dec.ok || b > dref + 1) fail(v);
        uint64_t pos = 0, copied = 0;
        for (uint64_t k = 0; k < b; ++k) {
          const uint64_t len = dec.decode(Blocks) + (k ? 1 : 0);
          if (!dec.ok || len > dref - pos) fail(v);
          if ((k & 1) == 0) copied += len;
          pos += len;
        }
        if ((b & 1) == 0) copied += dref - pos;
        if (copied > d) fail(v);
        extras = d - copied;
      }
    }
    if (extras && minint != 0) {
      const uint64_t ni = dec.decode(IntervalCount);
      if (!dec.ok || ni > extras) fail(v);
      for (uint64_t k = 0; k < ni; ++k) {
        dec.decode(IntervalStart);
        const uint64_t len = dec.decode(IntervalLen) + minint;
        if (!dec.ok || len > extras) fail(v);
        extras -= len;
      }
    }
    if (extras) {
      dec.decode(FirstResidual);
      for (uint64_t k = 1; k < extras; ++k) dec.decode(Residual);
      if (!dec.ok) fail(v);
    }
  }
  // a full sequential decode ends where the encoder started (encoder.rs:22-28)
  if (dec.ptr != 0 || dec.state != WGA_LOWER_BOUND) throw Error(WGA_E_CORRUPT, "sequential bootstrap: the stream does not end at the initial encoder state");
}

}  // namespace wga

using namespace wga;

wga_graph::~wga_graph() {
  if (on_device) {
    cudaFree(d_stream_alloc);
    cudaFree(d_states);
    cudaFree(d_ptrs);
    cudaFree(d_bkt);
    cudaFree(d_ent);
    cudaFree(d_err);
    if (h_pub) cudaFreeHost((void*)h_pub);
    if (e2e_ws) cudaFree(e2e_ws);
    if (e2e_off) cudaFree(e2e_off);
    if (e2e_succ) cudaFree(e2e_succ);
    for (int i = 0; i < 2; ++i) {
      if (pipe_off[i]) cudaFree(pipe_off[i]);
      if (pipe_succ[i]) cudaFree(pipe_succ[i]);
      if (dec_done[i]) cudaEventDestroy(dec_done[i]);
      if (down_done[i]) cudaEventDestroy(down_done[i]);
    }
    for (auto& e : up_ev) if (e) cudaEventDestroy(e);
    if (s_up) cudaStreamDestroy(s_up);
    if (s_dec) cudaStreamDestroy(s_dec);
    if (s_down) cudaStreamDestroy(s_down);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (pinned) {
      if (!prelude.stream.empty()) cudaHostUnregister(prelude.stream.data());
      if (!phases.states.empty()) cudaHostUnregister(phases.states.data());
      if (!phases.pointers.empty()) cudaHostUnregister(phases.pointers.data());
    }
  }
}

// A record is one serial ANS chain (one lane decodes it symbol by symbol), so its length bounds the decode
// time of any range that contains it.  The pipelined host entry point uses this to decide whether cutting
// the range into chunks pays off (every chunk would wait for its own longest record).
uint64_t wga_graph::longest_record() {
  if (max_record_words == UINT64_MAX) {
    const uint64_t N = prelude.number_of_nodes;
    uint64_t mx = 0;
    for (uint64_t v = res_first; v < res_last; ++v) {
      const uint64_t hi = phases.pointers[N - 1 - v], lo = v + 1 < N ? phases.pointers[N - 2 - v] : 0;
      if (hi - lo > mx) mx = hi - lo;
    }
    max_record_words = mx;
  }
  return max_record_words;
}

void wga_graph::ensure_pipeline() {
  if (s_up) return;
  WGA_CUDA(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
  WGA_CUDA(cudaStreamCreateWithFlags(&s_dec, cudaStreamNonBlocking));
  WGA_CUDA(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    WGA_CUDA(cudaEventCreateWithFlags(&dec_done[i], cudaEventDisableTiming));
    WGA_CUDA(cudaEventCreateWithFlags(&down_done[i], cudaEventDisableTiming));
  }
}

// Same bytes as reupload(), but in node-range chunks (lowest nodes first) on the handle's upload stream, with
// one event per chunk, so that wga_decode_range_host can start on chunk 0 while the rest is still in flight.
void wga_graph::reupload_chunked() {
  if (!on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  ensure_pipeline();
  reupload_pin();
  const uint64_t N = prelude.number_of_nodes;
  const uint64_t n_res = res_last - res_first;
  const uint64_t nchunks = (n_res + e2e_chunk_nodes - 1) / e2e_chunk_nodes;
  while (up_ev.size() < nchunks) {
    cudaEvent_t e;
    WGA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    up_ev.push_back(e);
  }
  for (uint64_t c = 0; c < nchunks; ++c) {
    const uint64_t a = res_first + c * e2e_chunk_nodes, b = std::min(res_last, a + e2e_chunk_nodes);
    // node v: file entry N-1-v, device entry res_last-1-v ; its record = stream words [ptr(v+1), ptr(v))
    const uint64_t w_hi = phases.pointers[N - 1 - a], w_lo = b < N ? phases.pointers[N - 1 - b] : 0;
    if (w_hi > w_lo)
      WGA_CUDA(cudaMemcpyAsync(d_stream + (w_lo - stream_base), prelude.stream.data() + w_lo, (w_hi - w_lo) * 2,
                               cudaMemcpyHostToDevice, s_up));
    WGA_CUDA(cudaMemcpyAsync(d_states + (res_last - b), phases.states.data() + (N - b), (b - a) * 4, cudaMemcpyHostToDevice, s_up));
    WGA_CUDA(cudaMemcpyAsync(d_ptrs + (res_last - b), phases.pointers.data() + (N - b), (b - a) * 8, cudaMemcpyHostToDevice, s_up));
    WGA_CUDA(cudaEventRecord(up_ev[c], s_up));
  }
  up_pending = true;
}

void wga_graph::reupload_pin() {
  if (!pinned) {  // pin the host copies once so the copies run at full PCIe speed
    bool ok = true;
    if (!prelude.stream.empty()) ok &= cudaHostRegister(prelude.stream.data(), prelude.stream.size() * 2, cudaHostRegisterDefault) == cudaSuccess;
    if (!phases.states.empty()) ok &= cudaHostRegister(phases.states.data(), phases.states.size() * 4, cudaHostRegisterDefault) == cudaSuccess;
    if (!phases.pointers.empty()) ok &= cudaHostRegister(phases.pointers.data(), phases.pointers.size() * 8, cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();
    pinned = ok;
  }
}

void wga_graph::reupload(cudaStream_t st) {
  if (!on_device) throw Error(WGA_E_CUDA, "graph was opened host-only");
  const uint64_t N = prelude.number_of_nodes;
  const uint64_t n_res = res_last - res_first;
  if (!pinned) {  // pin the host copies once so the copies below run at full PCIe speed
    bool ok = true;
    if (!prelude.stream.empty()) ok &= cudaHostRegister(prelude.stream.data(), prelude.stream.size() * 2, cudaHostRegisterDefault) == cudaSuccess;
    if (!phases.states.empty()) ok &= cudaHostRegister(phases.states.data(), phases.states.size() * 4, cudaHostRegisterDefault) == cudaSuccess;
    if (!phases.pointers.empty()) ok &= cudaHostRegister(phases.pointers.data(), phases.pointers.size() * 8, cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();
    pinned = ok;
  }
  if (stream_words)
    WGA_CUDA(cudaMemcpyAsync(d_stream, prelude.stream.data() + stream_base, stream_words * 2, cudaMemcpyHostToDevice, st));
  if (n_res) {
    WGA_CUDA(cudaMemcpyAsync(d_states, phases.states.data() + (N - res_last), n_res * 4, cudaMemcpyHostToDevice, st));
    WGA_CUDA(cudaMemcpyAsync(d_ptrs, phases.pointers.data() + (N - res_last), n_res * 8, cudaMemcpyHostToDevice, st));
  }
}

// Uploads the decode inputs of nodes [res_first,res_last) to the current device.
void wga_graph::upload() {
  const uint64_t N = prelude.number_of_nodes;
  if (phases.states.size() != N || phases.pointers.size() != N)
    throw Error(WGA_E_FORMAT, "the GPU decode needs .states and .pointers with one phase per node");
  if (prelude.compression_window > 4095) throw Error(WGA_E_UNSUPPORTED, "compression window > 4095");
  if (N >= (1ull << 32)) throw Error(WGA_E_UNSUPPORTED, "graphs with >= 2^32 nodes need 64-bit successors");
  int dev = 0;
  WGA_CUDA(cudaGetDevice(&dev));
  device = dev;
  const uint64_t n_res = res_last - res_first;
  // stream span needed by [res_first,res_last): words [ptr(res_last), ptr(res_first)) ; ptr(N) := 0
  // (node v's record starts at pointer(N-1-v) and reads downwards, src/ans/decoder.rs:89-93)
  uint64_t hi = n_res ? phases.pointers[N - 1 - res_first] : 0;
  uint64_t lo = res_last < N ? phases.pointers[N - 1 - res_last] : 0;
  if (hi > prelude.stream.size() || lo > hi) throw Error(WGA_E_FORMAT, "stream pointers are not monotone");
  stream_base = lo;
  stream_words = hi - lo;
  if (stream_words >= 0xFFFFFFFFull) throw Error(WGA_E_UNSUPPORTED, "resident stream span of 2^32 words or more: open a smaller shard");
  // 8 padding words on both sides: the decoder prefetches the word below its read position without a bounds test
  WGA_CUDA(cudaMalloc(&d_stream_alloc, (stream_words + 16) * 2));
  WGA_CUDA(cudaMemset(d_stream_alloc, 0, (stream_words + 16) * 2));
  d_stream = d_stream_alloc + 8;
  if (stream_words)
    WGA_CUDA(cudaMemcpy(d_stream, prelude.stream.data() + lo, stream_words * 2, cudaMemcpyHostToDevice));
  WGA_CUDA(cudaMalloc(&d_states, (n_res + 1) * 4));
  WGA_CUDA(cudaMalloc(&d_ptrs, (n_res + 1) * 8));
  if (n_res) {
    // file order: entry i = node N-1-i ; resident nodes are entries [N-res_last, N-res_first)
    WGA_CUDA(cudaMemcpy(d_states, phases.states.data() + (N - res_last), n_res * 4, cudaMemcpyHostToDevice));
    WGA_CUDA(cudaMemcpy(d_ptrs, phases.pointers.data() + (N - res_last), n_res * 8, cudaMemcpyHostToDevice));
  }
  packed = pack_tables(prelude.tables);
  WGA_CUDA(cudaMalloc(&d_bkt, packed.bkt.size() * 8 + 16));
  WGA_CUDA(cudaMalloc(&d_ent, packed.ent.size() * 8 + 16));
  WGA_CUDA(cudaMemcpy(d_bkt, packed.bkt.data(), packed.bkt.size() * 8, cudaMemcpyHostToDevice));
  WGA_CUDA(cudaMemcpy(d_ent, packed.ent.data(), packed.ent.size() * 8, cudaMemcpyHostToDevice));
  WGA_CUDA(cudaMalloc(&d_err, 4));
  WGA_CUDA(cudaMemset(d_err, 0, 4));
  {
    void* hp = nullptr;
    WGA_CUDA(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
    h_pub = (volatile uint64_t*)hp;
    void* dp = nullptr;
    WGA_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
    d_pub = (uint64_t*)dp;
  }
  on_device = true;

  this->dev.tb.bkt = d_bkt;
  this->dev.tb.ent = d_ent;
  for (int c = 0; c < WGA_COMPONENTS; ++c) {
    this->dev.tb.bkt_off[c] = packed.bkt_off[c];
    this->dev.tb.ent_off[c] = packed.ent_off[c];
    this->dev.tb.nb[c] = packed.nb[c];
    this->dev.tb.nent[c] = packed.nent[c];
    this->dev.tb.L[c] = packed.L[c];
    this->dev.tb.R[c] = packed.R[c];
  }
  this->dev.stream = d_stream;
  this->dev.stream_base = stream_base;
  this->dev.stream_words = stream_words;
  this->dev.states = d_states;
  this->dev.ptrs = d_ptrs;
  this->dev.top = res_last ? res_last - 1 : 0;
  this->dev.N = N;
  this->dev.window = (uint32_t)prelude.compression_window;
  this->dev.min_interval = (uint32_t)prelude.min_interval_length;
}
