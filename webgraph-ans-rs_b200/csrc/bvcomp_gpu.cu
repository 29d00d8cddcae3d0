// =============================================================================
//  bvcomp_gpu.cu -- candidate-reference costing of BvComp on the GPU (sm_100a)
// =============================================================================
//  What it replaces: in every pass of ANSBvGraph::store (src/bvgraph/random_access.rs:105-163) webgraph's BvComp
//  compresses each successor list against each of the `window` previous lists and asks the estimator
//  (EntropyEstimator::get_symbol_cost, src/bvgraph/estimators/entropy_estimator.rs:102-113, or the Log2Estimator)
//  for the cost of every symbol of every candidate record, one node and one candidate at a time on one core.
//  Here: one device thread per (node, reference offset) pair computes the cost of that candidate record in a single
//  streaming pass over the two sorted lists -- copy/skip blocks, intervals and residuals are costed as they are
//  found, nothing is materialised --, then one thread per BvComp chunk replays the selection rule (strictly smaller
//  cost wins, so the nearest candidate wins ties; chains bounded by max_ref_count).  Costs are integers (16.16 fixed
//  point from the table the host builds), so the chosen references are exactly the host's.
//  The host then compresses only the chosen candidate of every node (bvcomp.cpp, `choice`).
// =============================================================================
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "bvcomp.hpp"
#include "graph.hpp"

namespace wga {

namespace {

struct EstDev {
  const uint32_t* table[WGA_COMPONENTS];
  uint32_t table_len[WGA_COMPONENTS];
  uint64_t thr[WGA_COMPONENTS], off[WGA_COMPONENTS];
  uint32_t fid[WGA_COMPONENTS], rad[WGA_COMPONENTS];
  int log2;
};

constexpr uint64_t NO_COST = ~0ull;

// Estimator::cost (bvcomp.hpp).  A symbol beyond 48 bits makes the host throw; here it poisons the candidate.
__device__ __forceinline__ uint64_t est_cost(const EstDev& e, int c, uint64_t value, bool& bad) {
  if (e.log2) return 63u - (uint32_t)__clzll((long long)(value + 2));
  uint64_t sym = value;
  if (value >= e.thr[c]) {
    const uint32_t bits = 64u - (uint32_t)__clzll((long long)value);
    const uint64_t cuts = (bits - e.fid[c]) / e.rad[c];
    sym = (value >> (cuts * e.rad[c])) + e.off[c] * cuts;
  }
  if (sym >= e.table_len[c]) { bad = true; return 0; }
  return __ldg(e.table[c] + sym);
}

__device__ __forceinline__ uint64_t int2nat(int64_t x) { return x >= 0 ? (uint64_t)x << 1 : ((uint64_t)(-x) << 1) - 1; }

// Cost of node `node`'s record with reference offset `delta` (0: none): Candidate::compress + write with the
// estimator as sink (bvcomp.cpp), fused into one pass.
__device__ uint64_t record_cost(const EstDev& e, uint64_t node, const uint32_t* cur, uint32_t nc, const uint32_t* ref,
                                uint32_t nr, uint32_t delta, uint32_t L, bool& bad) {
  uint64_t bits = est_cost(e, Outdegree, nc, bad);
  if (nc == 0) return bits;
  bits += est_cost(e, ReferenceOffset, delta, bad);
  // the extras (successors that are not copied), consumed as a stream: maximal runs of consecutive ids of length
  // >= max(2, L) are intervals, the rest residuals
  bool have_run = false;
  uint64_t run_start = 0, run_len = 0, n_int = 0, prev_end = 0, n_res = 0, prev_res = 0, n_extras = 0;
  auto flush_run = [&]() {
    if (L != 0 && run_len >= 2 && run_len >= L) {
      bits += est_cost(e, IntervalStart, n_int == 0 ? int2nat((int64_t)run_start - (int64_t)node) : run_start - prev_end - 1, bad);
      bits += est_cost(e, IntervalLen, run_len - L, bad);
      prev_end = run_start + run_len;
      ++n_int;
    } else {
      for (uint64_t t = 0; t < run_len; ++t) {
        const uint64_t r = run_start + t;
        bits += n_res == 0 ? est_cost(e, FirstResidual, int2nat((int64_t)r - (int64_t)node), bad)
                           : est_cost(e, Residual, r - prev_res - 1, bad);
        prev_res = r;
        ++n_res;
      }
    }
  };
  auto push_extra = [&](uint64_t x) {
    ++n_extras;
    if (have_run && x == run_start + run_len) ++run_len;
    else {
      if (have_run) flush_run();
      run_start = x;
      run_len = 1;
      have_run = true;
    }
  };
  if (delta == 0) {
    for (uint32_t j = 0; j < nc; ++j) push_extra(cur[j]);
  } else {
    // copy / skip blocks against the referenced list (Candidate::diff): the first block is emitted as its length,
    // the others as length - 1
    uint32_t j = 0, k = 0;
    uint64_t run = 0, nb = 0;
    bool copying = true;
    auto emit_block = [&]() {
      bits += est_cost(e, Blocks, nb == 0 ? run : run - 1, bad);
      ++nb;
      run = 0;
    };
    while (j < nc && k < nr) {
      const uint32_t a = cur[j], b = ref[k];
      if (copying) {
        if (a > b) { emit_block(); copying = false; }
        else if (a < b) { push_extra(a); ++j; }
        else { ++j; ++k; ++run; }
      } else {
        if (a < b) { push_extra(a); ++j; }
        else if (a > b) { ++k; ++run; }
        else { emit_block(); copying = true; }
      }
    }
    if (copying && k < nr) emit_block();
    for (; j < nc; ++j) push_extra(cur[j]);
    bits += est_cost(e, BlockCount, nb, bad);
  }
  if (have_run) flush_run();
  if (n_extras != 0 && L != 0) bits += est_cost(e, IntervalCount, n_int, bad);
  return bits;
}

// costs[(i - i0) * (W + 1) + delta] for nodes i0 <= i < i1 of the range (node id = first + i)
__global__ void __launch_bounds__(256) k_candidate_costs(EstDev e, const uint64_t* __restrict__ offs,
                                                         const uint32_t* __restrict__ succ, uint64_t first, uint64_t i0,
                                                         uint64_t i1, uint32_t W, uint32_t L, uint64_t chunk_nodes,
                                                         uint64_t* __restrict__ costs, uint32_t* err) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t per = (uint64_t)W + 1;
  if (tid >= (i1 - i0) * per) return;
  // neighbouring threads take the SAME offset of neighbouring nodes (similar lists, similar trip counts)
  const uint64_t nb = i1 - i0;
  const uint32_t delta = (uint32_t)(tid / nb);
  const uint64_t i = i0 + tid % nb;
  const uint64_t v = first + i;
  const uint64_t a = chunk_nodes ? max(first, v / chunk_nodes * chunk_nodes) : first;  // no reference crosses a chunk start
  uint64_t c = NO_COST;
  if (delta <= min((uint64_t)W, v - a)) {
    const uint32_t nc = (uint32_t)(offs[i + 1] - offs[i]);
    const uint32_t nr = delta ? (uint32_t)(offs[i - delta + 1] - offs[i - delta]) : 0u;
    if (delta == 0 || nr != 0) {
      bool bad = false;
      c = record_cost(e, v, succ + offs[i], nc, delta ? succ + offs[i - delta] : nullptr, nr, delta, L, bad);
      if (bad) { atomicOr(err, 1u); c = NO_COST; }
    }
  }
  costs[(i - i0) * per + delta] = c;
}

// One thread per BvComp chunk: the selection loop of bvcomp_range.  counts[i] = length of the reference chain of
// node first + i (kept for the whole range: a batch continues where the previous one stopped).
__global__ void k_choose(const uint64_t* __restrict__ costs, uint64_t first, uint64_t i0, uint64_t i1, uint32_t W,
                         uint64_t max_ref, uint64_t chunk_nodes, uint32_t* counts, uint16_t* choice) {
  const uint64_t per = (uint64_t)W + 1;
  uint64_t lo = i0, hi = i1;
  if (chunk_nodes) {  // this thread's chunk, clipped to the batch
    const uint64_t c0 = (first + i0) / chunk_nodes;
    const uint64_t c = c0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t a = c * chunk_nodes, b = a + chunk_nodes;
    if (a >= first + i1) return;
    lo = max(first + i0, a) - first;
    hi = min(first + i1, b) - first;
  } else if (blockIdx.x || threadIdx.x) return;
  for (uint64_t i = lo; i < hi; ++i) {
    const uint64_t v = first + i;
    const uint64_t a = chunk_nodes ? max(first, v / chunk_nodes * chunk_nodes) : first;
    const uint64_t* row = costs + (i - i0) * per;
    uint64_t best_bits = row[0], best = 0;
    uint32_t best_count = 0;
    const uint64_t deltas = min((uint64_t)W, v - a);
    for (uint64_t delta = 1; delta <= deltas; ++delta) {
      const uint32_t count = counts[i - delta];
      if (count >= max_ref) continue;
      const uint64_t bits = row[delta];
      if (bits == NO_COST) continue;  // empty candidate list
      if (bits < best_bits) { best_bits = bits; best = delta; best_count = count + 1; }
    }
    counts[i] = best_count;
    choice[i] = (uint16_t)best;
  }
}

struct Buf {
  void* p = nullptr;
  ~Buf() { if (p) cudaFree(p); }
  template <class T> T* alloc(uint64_t n) {
    WGA_CUDA(cudaMalloc(&p, std::max<uint64_t>(n, 1) * sizeof(T)));
    return (T*)p;
  }
};

}  // namespace

void bvcomp_choose_gpu(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first, uint64_t n, const BvCompParams& p,
                       const Estimator& est, uint64_t chunk_nodes, std::vector<uint16_t>& choice,
                       std::vector<uint64_t>* costs_out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Error(WGA_E_CUDA, "no CUDA device: the GPU candidate costing has no CPU fallback");
  const bool timing = getenv("WGA_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  double t_upload = 0, t_kernels = 0;
  const uint64_t W = p.window, per = W + 1;
  if (W > 4095) throw Error(WGA_E_UNSUPPORTED, "compression window > 4095");
  choice.assign(n, 0);
  if (costs_out) costs_out->assign(n * per, NO_COST);
  if (n == 0 || W == 0) return;
  const uint64_t arcs = h_offsets[n] - h_offsets[0];
  if (h_offsets[0] != 0) throw Error(WGA_E_ARG, "offsets must start at 0");
  Buf b_off, b_succ, b_costs, b_counts, b_choice, b_err, b_tab[WGA_COMPONENTS];
  uint64_t* d_off = b_off.alloc<uint64_t>(n + 1);
  uint32_t* d_succ = b_succ.alloc<uint32_t>(arcs);
  WGA_CUDA(cudaMemcpy(d_off, h_offsets, (n + 1) * 8, cudaMemcpyHostToDevice));
  if (arcs) WGA_CUDA(cudaMemcpy(d_succ, h_succ, arcs * 4, cudaMemcpyHostToDevice));
  EstDev e{};
  e.log2 = est.is_log2() ? 1 : 0;
  for (int c = 0; c < WGA_COMPONENTS && !est.is_log2(); ++c) {
    const std::vector<uint32_t>& t = est.table(c);
    uint32_t* d = b_tab[c].alloc<uint32_t>(t.size());
    WGA_CUDA(cudaMemcpy(d, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    e.table[c] = d;
    e.table_len[c] = (uint32_t)t.size();
    e.thr[c] = est.threshold(c);
    e.off[c] = est.offset(c);
    e.fid[c] = est.fidelity(c);
    e.rad[c] = est.radix(c);
  }
  if (timing) { cudaDeviceSynchronize(); t_upload = now() - t_start; }
  // batches of nodes (aligned to the chunks): the cost table of a batch is batch * (W + 1) * 8 bytes
  uint64_t batch = std::max<uint64_t>(1, (1ull << 28) / per);  // <= 2 GiB of costs
  if (chunk_nodes) batch = std::max<uint64_t>(chunk_nodes, batch / chunk_nodes * chunk_nodes);
  uint64_t* d_costs = b_costs.alloc<uint64_t>(std::min(n, batch) * per);
  uint32_t* d_counts = b_counts.alloc<uint32_t>(n);
  uint16_t* d_choice = b_choice.alloc<uint16_t>(n);
  uint32_t* d_err = b_err.alloc<uint32_t>(1);
  WGA_CUDA(cudaMemset(d_err, 0, 4));
  WGA_CUDA(cudaMemset(d_counts, 0, n * 4));
  for (uint64_t i0 = 0; i0 < n;) {
    uint64_t i1 = std::min(n, i0 + batch);
    if (chunk_nodes && i1 < n) i1 = std::max(i0 + 1, (first + i1) / chunk_nodes * chunk_nodes - first);  // end on a chunk boundary
    const uint64_t threads = (i1 - i0) * per;
    k_candidate_costs<<<(unsigned)((threads + 255) / 256), 256>>>(e, d_off, d_succ, first, i0, i1, (uint32_t)W,
                                                                 (uint32_t)p.min_interval_length, chunk_nodes, d_costs, d_err);
    const uint64_t n_chunks = chunk_nodes ? (first + i1 - 1) / chunk_nodes - (first + i0) / chunk_nodes + 1 : 1;
    k_choose<<<(unsigned)((n_chunks + 63) / 64), 64>>>(d_costs, first, i0, i1, (uint32_t)W, p.max_ref_count, chunk_nodes,
                                                       d_counts, d_choice);
    count_launch(2);
    WGA_CUDA(cudaGetLastError());
    if (costs_out) WGA_CUDA(cudaMemcpy(costs_out->data() + i0 * per, d_costs, (i1 - i0) * per * 8, cudaMemcpyDeviceToHost));
    i0 = i1;
  }
  uint32_t herr = 0;
  WGA_CUDA(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
  if (timing) {
    t_kernels = now() - t_start - t_upload;
    fprintf(stderr, "[wga] bvcomp_choose_gpu: %llu nodes, upload %.3f s, cost + choose kernels %.3f s\n",
            (unsigned long long)n, t_upload, t_kernels);
  }
  if (herr) throw Error(WGA_E_ARG, "Symbol can't be bigger than u48::MAX");
  WGA_CUDA(cudaMemcpy(choice.data(), d_choice, n * 2, cudaMemcpyDeviceToHost));
}

}  // namespace wga
