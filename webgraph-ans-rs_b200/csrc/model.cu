// =============================================================================
//  model.cu -- encoder-side symbol-model construction on the GPU (sm_100a)
// =============================================================================
//  Replaces ANSModel4EncoderBuilder (src/ans/model4encoder_builder.rs:39-327):
//    K4  k_histogram        push_symbol (:67-78) for a device array of (component, raw symbol) pairs.
//                           Values < 1024 are counted in shared-memory privatised bins (9 x 1024 u32 per
//                           block); larger values go to the canonical global bins with red.global and are
//                           appended to a key buffer for the exact sparse histogram.
//        cub sort + RLE     exact (component, raw symbol) -> count for values >= 1024; only needed for the
//                           raw entropy of :275-289.
//    K5a k_fold             for every (component, (fidelity,radix)) pair: folded histogram (:106-120) from
//                           the canonical bins.  Every folding keeps <= 10 significant bits, so it is a
//                           function of (bit length, top 10 bits) = the canonical bin.
//        cub segmented sort symbols by ascending (frequency, index)  (:132-138; stable tie-break)
//    K5b k_scale            scale_freqs (src/utils/data_utils.rs:15-39) + approximated cost (:297-327, summed in
//                           symbol-index order like the reference) for all 52 foldings x 17 frame sizes of every
//                           component, one serial chain per thread.
//    K5c k_select_emit      the acceptance loop (:140-206), then the winner's table (:216-234).
//  This file is compiled with -fmad=false: scale_freqs is IEEE f64 `* / + floor` evaluated exactly as the
//  reference writes it, so the frequency tables are bit-exact with the CPU.  Costs use CUDA's log2 (<= 1 ulp
//  from libm's) and a parallel sum for the raw entropy; they agree with the CPU to ~1e-12 relative and only
//  feed the `ratio <= THETA` test.
// =============================================================================
#include "model.hpp"

#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>

#include "graph.hpp"

namespace wga {

namespace {

constexpr int NC = WGA_COMPONENTS;
constexpr int NP = 52;                  // PARAMS_COMBINATIONS (:28-37)
constexpr int NF = 17;                  // frame sizes 2^0 .. 2^16
constexpr int CANON = WGA_CANON_BINS;   // 1024 exact + 38 bit lengths x 512
constexpr int FOLD_MAX = 20480;         // >= fold(2^48-1) + 1 for every (F,R)
constexpr uint64_t MAX_RAW = (1ull << 48) - 1;
constexpr double THETA = 1.0001;        // :23

__constant__ uint8_t c_fid[NP];
__constant__ uint8_t c_rad[NP];
const uint8_t h_fid[NP] = {1, 2, 3, 1, 2, 3, 4, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 6, 1, 2, 3, 4, 5, 6, 7, 1,
                           2, 3, 4, 5, 6, 7, 8, 1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
const uint8_t h_rad[NP] = {3, 2, 1, 4, 3, 2, 1, 5, 4, 3, 2, 1, 6, 5, 4, 3, 2, 1, 7, 6, 5, 4, 3, 2, 1, 8,
                           7, 6, 5, 4, 3, 2, 1, 9, 8, 7, 6, 5, 4, 3, 2, 1, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1};

__host__ __device__ inline uint32_t canon_bin(uint64_t sym) {
  if (sym < 1024) return (uint32_t)sym;
  int bl = 64 - (int)
#ifdef __CUDA_ARCH__
      __clzll((long long)sym);
#else
      __builtin_clzll(sym);
#endif
  return 1024u + (uint32_t)(bl - 11) * 512u + (uint32_t)((sym >> (bl - 10)) & 511u);
}
// smallest raw symbol of a canonical bin (any member folds identically)
__host__ __device__ inline uint64_t canon_rep(uint32_t bin) {
  if (bin < 1024) return bin;
  uint32_t k = bin - 1024;
  int bl = 11 + (int)(k / 512);
  return (uint64_t)(512 + (k % 512)) << (bl - 10);
}
// fold_without_streaming_out (src/utils/ans_utils.rs:4-12) for sym >= threshold, else identity (:113-116)
__host__ __device__ inline uint32_t fold_sym(uint64_t sym, int F, int R) {
  if (sym < (1ull << (F + R - 1))) return (uint32_t)sym;
  int bl = 64 - (int)
#ifdef __CUDA_ARCH__
      __clzll((long long)sym);
#else
      __builtin_clzll(sym);
#endif
  int cuts = (bl - F) / R;
  return (uint32_t)((sym >> (cuts * R)) + (uint64_t)(((1u << R) - 1) << (F - 1)) * (uint64_t)cuts);
}

// ------------------------------------------------------------------------------------------ K4
constexpr int HIST_TPB = 256;
__global__ void __launch_bounds__(HIST_TPB) k_histogram(const uint8_t* __restrict__ comps,
                                                         const uint64_t* __restrict__ syms, uint64_t n,
                                                         unsigned long long* __restrict__ bins,
                                                         uint64_t* __restrict__ big_keys,
                                                         unsigned long long* __restrict__ big_count,
                                                         uint32_t* __restrict__ err) {
  __shared__ uint32_t sh[NC * 1024];
  for (int i = threadIdx.x; i < NC * 1024; i += HIST_TPB) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * HIST_TPB;
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t base = (uint64_t)blockIdx.x * HIST_TPB; base < n; base += stride) {
    uint64_t i = base + threadIdx.x;
    bool big = false;
    uint64_t key = 0;
    if (i < n) {
      uint32_t c = comps[i];
      uint64_t s = syms[i];
      if (c >= NC || s > MAX_RAW) {
        atomicOr(err, 1u);  // "Symbol can't be bigger than u48::MAX" (:68-70)
      } else if (s < 1024) {
        atomicAdd(&sh[c * 1024 + (uint32_t)s], 1u);
      } else {
        atomicAdd(&bins[(uint64_t)c * CANON + canon_bin(s)], 1ull);
        big = true;
        key = ((uint64_t)c << 48) | s;
      }
    }
    // warp-aggregated append of the large values
    uint32_t m = __ballot_sync(0xffffffffu, big);
    if (m) {
      unsigned long long pos = 0;
      if (lane == (uint32_t)(__ffs(m) - 1)) pos = atomicAdd(big_count, (unsigned long long)__popc(m));
      pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1);
      if (big) big_keys[pos + __popc(m & ((1u << lane) - 1))] = key;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NC * 1024; i += HIST_TPB) {
    uint32_t v = sh[i];
    if (v) atomicAdd(&bins[(uint64_t)(i / 1024) * CANON + (i % 1024)], (unsigned long long)v);
  }
}

// ------------------------------------------------------------------------------------------ raw entropy
struct RawCostSparse {
  const uint64_t* keys;
  const uint64_t* counts;
  const double* totals;
  int comp;
  __device__ double operator()(uint64_t i) const {
    uint64_t k = keys[i];
    if ((int)(k >> 48) != comp) return 0.0;
    double f = (double)counts[i];
    double prob = f / totals[comp];
    return -log2(prob) * f;
  }
};
struct RawCostDense {
  const unsigned long long* bins;
  const double* totals;
  int comp;
  __device__ double operator()(uint32_t s) const {
    unsigned long long f = bins[(uint64_t)comp * CANON + s];
    if (!f) return 0.0;
    double prob = (double)f / totals[comp];
    return -log2(prob) * (double)f;
  }
};

__global__ void k_totals(const unsigned long long* bins, double* totals, unsigned long long* totals_u) {
  // one block per component
  __shared__ unsigned long long red[256];
  unsigned long long s = 0;
  for (int i = threadIdx.x; i < CANON; i += 256) s += bins[(uint64_t)blockIdx.x * CANON + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[blockIdx.x] = (double)red[0];
    totals_u[blockIdx.x] = red[0];
  }
}

// ------------------------------------------------------------------------------------------ K5a
__global__ void k_fold(const unsigned long long* __restrict__ bins, unsigned long long* __restrict__ folded) {
  const int p = blockIdx.x, c = blockIdx.y;
  const int F = c_fid[p], R = c_rad[p];
  unsigned long long* out = folded + ((uint64_t)c * NP + p) * FOLD_MAX;
  for (int b = threadIdx.x; b < CANON; b += blockDim.x) {
    unsigned long long f = bins[(uint64_t)c * CANON + b];
    if (f) atomicAdd(&out[fold_sym(canon_rep((uint32_t)b), F, R)], f);
  }
}

__global__ void k_keys(const unsigned long long* __restrict__ folded, uint64_t* __restrict__ keys,
                       uint32_t* __restrict__ nnz, uint32_t* __restrict__ biggest) {
  const int p = blockIdx.x, c = blockIdx.y;
  const int F = c_fid[p], R = c_rad[p];
  const uint32_t max_bucket = fold_sym(MAX_RAW, F, R);  // length of the reference's vector (:106-108)
  const uint64_t base = ((uint64_t)c * NP + p) * FOLD_MAX;
  uint32_t cnt = 0, big = 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)FOLD_MAX; i += blockDim.x) {
    unsigned long long f = i < max_bucket ? folded[base + i] : 0ull;
    keys[base + i] = f ? ((uint64_t)f << 15) | i : ~0ull;  // ascending (frequency, index)
    if (f) { ++cnt; big = i; }
  }
  __shared__ uint32_t s_cnt, s_big;
  if (threadIdx.x == 0) { s_cnt = 0; s_big = 0; }
  __syncthreads();
  atomicAdd(&s_cnt, cnt);
  atomicMax(&s_big, big);
  __syncthreads();
  if (threadIdx.x == 0) { nnz[c * NP + p] = s_cnt; biggest[c * NP + p] = s_big; }
}

struct SegBegin {
  __host__ __device__ int64_t operator()(int i) const { return (int64_t)i * FOLD_MAX; }
};
struct SegEnd {
  __host__ __device__ int64_t operator()(int i) const { return (int64_t)(i + 1) * FOLD_MAX; }
};

// scale_freqs (data_utils.rs:15-39) over the symbols in ascending (frequency,index) order.
// `emit` receives (symbol index, real frequency, approximated frequency).  Returns false where the
// reference bails out.
template <class Emit>
__device__ bool scale_freqs_dev(const uint64_t* __restrict__ sorted, uint32_t n, unsigned long long m_total,
                                long long new_m, Emit&& emit) {
  const double ratio = (double)new_m / (double)m_total;
  unsigned long long m = m_total;
  const double dn = (double)n;
  for (uint32_t index = 0; index < n; ++index) {
    uint64_t k = sorted[index];
    unsigned long long f = k >> 15;
    uint32_t sym = (uint32_t)(k & 0x7FFF);
    double second_ratio = (double)new_m / (double)m;
    double scale = (double)(n - index) * ratio / dn + (double)index * second_ratio / dn;
    double v = floor(0.5 + scale * (double)f);
    unsigned long long a = (unsigned long long)v;
    if (a < 1) a = 1;
    emit(sym, f, a);
    new_m -= (long long)a;
    m -= f;
    if (new_m < 0) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------ K5b
// One launch per component c: grid NP, 32 threads: thread k handles frame 2^k.  scale_freqs visits the symbols in
// ascending frequency, the approximated cost (:297-327) is summed in SYMBOL-INDEX order as the reference does
// (a different summation order can flip `ratio <= THETA` / `new_cost >= lowest_cost` on a knife-edge input), so
// the approximated frequencies first go to a scratch slice [p][k][symbol] and are then read back in index order.
__global__ void k_scale(int c, const uint64_t* __restrict__ sorted_keys, const uint32_t* __restrict__ nnz,
                        const uint32_t* __restrict__ biggest, const unsigned long long* __restrict__ folded,
                        const unsigned long long* __restrict__ totals_u, uint32_t* __restrict__ approx_all,
                        double* __restrict__ cost, uint8_t* __restrict__ ok) {
  const int p = blockIdx.x, k = threadIdx.x;
  if (k >= NF) return;
  const int F = c_fid[p], R = c_rad[p];
  const uint32_t n = nnz[c * NP + p];
  const uint64_t idx = ((uint64_t)c * NP + p) * NF + k;
  ok[idx] = 0;
  cost[idx] = 0.0;
  if (n == 0) return;
  const unsigned long long m = 1ull << k;
  if (m < n) return;  // the reference starts at next_power_of_two(n) (:125-128)
  const uint64_t seg = ((uint64_t)c * NP + p) * FOLD_MAX;
  const uint64_t* sorted = sorted_keys + seg;
  uint32_t* approx = approx_all + ((uint64_t)p * NF + k) * FOLD_MAX;
  const bool good = scale_freqs_dev(sorted, n, totals_u[c], (long long)m,
                                    [&](uint32_t sym, unsigned long long, unsigned long long a) { approx[sym] = (uint32_t)a; });
  ok[idx] = good ? 1 : 0;
  if (!good) return;
  const uint32_t thr = 1u << (F + R - 1);
  const uint32_t off = ((1u << R) - 1) << (F - 1);
  const double frame = (double)m;
  const uint32_t len = biggest[c * NP + p] + 1;
  double info = 0.0;
  for (uint32_t sym = 0; sym < len; ++sym) {
    const unsigned long long f = folded[seg + sym];
    if (f == 0) continue;  // approximated frequency 0 <=> folded frequency 0 (:308-310)
    const double folds = sym < thr ? 0.0 : (double)((sym - thr) / off + 1);
    const double prob = (double)approx[sym] / frame;
    info += (-log2(prob) + folds * (double)R) * (double)f;  // :322-323
  }
  cost[idx] = info;
}

struct EmitOut {
  wga_encoder_entry* entries;  // [NC][FOLD_MAX]
  uint32_t* table_len;         // [NC]
  uint32_t* params;            // [NC][3] frame_log2, fidelity, radix ; frame_log2 == 0xFFFFFFFF => failed
  double* final_cost;          // [NC]
};

// ------------------------------------------------------------------------------------------ K5c
// One block per component: thread 0 runs the acceptance loop (:140-206) and the winner's scale_freqs,
// then the block writes the table (:216-234).
__global__ void k_select_emit(const uint64_t* __restrict__ sorted_keys, const uint32_t* __restrict__ nnz,
                              const uint32_t* __restrict__ biggest, const unsigned long long* __restrict__ totals_u,
                              const double* __restrict__ cost, const uint8_t* __restrict__ ok,
                              const double* __restrict__ occ, uint32_t* __restrict__ approx_scratch, EmitOut out) {
  const int c = blockIdx.x;
  __shared__ int s_p, s_logm;
  uint32_t* approx = approx_scratch + (uint64_t)c * FOLD_MAX;
  if (threadIdx.x == 0) {
    s_p = -1;
    s_logm = 0;
    if (totals_u[c] != 0) {
      double ogc = 0.0;
      for (int i = 0; i < NC; ++i) ogc += occ[i];  // :84
      unsigned long long frame_size = ~0ull;
      double lowest = DBL_MAX;
      for (int p = 0; p < NP; ++p) {
        uint32_t n = nnz[c * NP + p];
        unsigned long long m = 1;
        while (m < n) m <<= 1;
        while (m <= 65536ull) {
          int k = 63 - __clzll((long long)m);
          uint64_t idx = ((uint64_t)c * NP + p) * NF + k;
          if (ok[idx]) {
            double new_cost = cost[idx];
            double difference = new_cost - occ[c];
            double ratio = (ogc + difference) / ogc;
            if (ratio <= THETA) {
              if (m < frame_size) { lowest = new_cost; frame_size = m; s_p = p; s_logm = k; }
            } else if (m == 65536ull) {
              if (new_cost >= lowest) break;
              lowest = new_cost; frame_size = m; s_p = p; s_logm = k;
              break;
            }
          }
          m <<= 1;
        }
      }
      out.final_cost[c] = s_p >= 0 ? lowest : 0.0;
    } else {
      out.final_cost[c] = 0.0;
    }
  }
  __syncthreads();
  const int p = s_p;
  if (totals_u[c] == 0) {  // Default model (:93-97, component_model4encoder.rs:59-70)
    if (threadIdx.x == 0) {
      out.table_len[c] = 0;
      out.params[c * 3 + 0] = 0; out.params[c * 3 + 1] = 2; out.params[c * 3 + 2] = 2;
    }
    return;
  }
  if (p < 0) {  // the reference asserts (:209-212)
    if (threadIdx.x == 0) { out.table_len[c] = 0; out.params[c * 3 + 0] = 0xFFFFFFFFu; }
    return;
  }
  const uint32_t len = biggest[c * NP + p] + 1;  // drain(biggest_symbol+1..) (:174)
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) approx[i] = 0;  // folded freq 0 stays 0 (approx = freqs.to_vec())
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint64_t* sorted = sorted_keys + ((uint64_t)c * NP + p) * FOLD_MAX;
    scale_freqs_dev(sorted, nnz[c * NP + p], totals_u[c], 1ll << s_logm,
                    [&](uint32_t sym, unsigned long long, unsigned long long a) { approx[sym] = (uint32_t)a; });
    const uint32_t log_m = (uint32_t)s_logm;
    const uint32_t kk = log_m > 0 ? 16 - log_m : 15;  // :217-218
    uint16_t cumul = 0;
    wga_encoder_entry* e = out.entries + (uint64_t)c * FOLD_MAX;
    for (uint32_t i = 0; i < len; ++i) {
      uint16_t f = (uint16_t)approx[i];
      e[i].freq = f;
      e[i].upperbound = (1u << (kk + 16)) * (uint32_t)f;  // component_model4encoder.rs:28-34
      e[i].cumul_freq = cumul;
      uint32_t s = (uint32_t)cumul + f;
      cumul = s > 0xFFFFu ? 0 : (uint16_t)s;  // checked_add(..).unwrap_or(0) (:224)
    }
    out.table_len[c] = len;
    out.params[c * 3 + 0] = log_m;
    out.params[c * 3 + 1] = c_fid[p];
    out.params[c * 3 + 2] = c_rad[p];
  }
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    WGA_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  template <class T>
  T* as() { return (T*)p; }
};

}  // namespace

struct ModelBuilder::Impl {
  unsigned long long* d_bins = nullptr;  // NC * CANON
  uint32_t* d_err = nullptr;
  // exact sparse histogram of values >= 1024: sorted unique keys (component << 48 | symbol) + counts
  DevBuf sp_keys, sp_counts;
  uint64_t sp_n = 0;
  bool constants_set = false;
  DevBuf k_in, c_in, k_out, c_out, tmp, d_num, big_keys, big_count;  // scratch of accumulate / merge_sparse

  void init() {
    if (d_bins) return;
    WGA_CUDA(cudaMalloc((void**)&d_bins, (size_t)NC * CANON * 8));
    WGA_CUDA(cudaMemset(d_bins, 0, (size_t)NC * CANON * 8));
    WGA_CUDA(cudaMalloc((void**)&d_err, 4));
    WGA_CUDA(cudaMemset(d_err, 0, 4));
    WGA_CUDA(cudaMemcpyToSymbol(c_fid, h_fid, NP));
    WGA_CUDA(cudaMemcpyToSymbol(c_rad, h_rad, NP));
  }
  ~Impl() {
    if (d_bins) cudaFree(d_bins);
    if (d_err) cudaFree(d_err);
  }

  // merges (keys, counts) [n, unsorted, duplicates allowed] into the sparse histogram
  void merge_sparse(const uint64_t* d_keys, const uint64_t* d_counts /* may be null: all ones */, uint64_t n,
                    cudaStream_t st) {
    if (n == 0) return;
    const uint64_t tot = sp_n + n;
    // (grow-only scratch kept in the builder: allocating gigabyte buffers per call costs more than the kernels)
    k_in.ensure(tot * 8); c_in.ensure(tot * 8); k_out.ensure(tot * 8); c_out.ensure(tot * 8);
    d_num.ensure(8);
    if (sp_n) {
      WGA_CUDA(cudaMemcpyAsync(k_in.p, sp_keys.p, sp_n * 8, cudaMemcpyDeviceToDevice, st));
      WGA_CUDA(cudaMemcpyAsync(c_in.p, sp_counts.p, sp_n * 8, cudaMemcpyDeviceToDevice, st));
    }
    WGA_CUDA(cudaMemcpyAsync(k_in.as<uint64_t>() + sp_n, d_keys, n * 8, cudaMemcpyDeviceToDevice, st));
    if (d_counts) {
      WGA_CUDA(cudaMemcpyAsync(c_in.as<uint64_t>() + sp_n, d_counts, n * 8, cudaMemcpyDeviceToDevice, st));
    } else {
      fill_ones(c_in.as<uint64_t>() + sp_n, n, st);
    }
    size_t tb = 0;
    WGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in.as<uint64_t>(), k_out.as<uint64_t>(),
                                             c_in.as<uint64_t>(), c_out.as<uint64_t>(), (int64_t)tot, 0, 52, st));
    tmp.ensure(tb);
    WGA_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k_in.as<uint64_t>(), k_out.as<uint64_t>(),
                                             c_in.as<uint64_t>(), c_out.as<uint64_t>(), (int64_t)tot, 0, 52, st));
    count_launch(4);
    size_t tb2 = 0;
    WGA_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tb2, k_out.as<uint64_t>(), k_in.as<uint64_t>(),
                                            c_out.as<uint64_t>(), c_in.as<uint64_t>(), d_num.as<uint64_t>(),
                                            cub::Sum(), (int64_t)tot, st));
    tmp.ensure(tb2);
    WGA_CUDA(cub::DeviceReduce::ReduceByKey(tmp.p, tb2, k_out.as<uint64_t>(), k_in.as<uint64_t>(),
                                            c_out.as<uint64_t>(), c_in.as<uint64_t>(), d_num.as<uint64_t>(),
                                            cub::Sum(), (int64_t)tot, st));
    count_launch(2);
    uint64_t uniq = 0;
    WGA_CUDA(cudaMemcpyAsync(&uniq, d_num.p, 8, cudaMemcpyDeviceToHost, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    sp_keys.ensure(uniq * 8 + 8);
    sp_counts.ensure(uniq * 8 + 8);
    WGA_CUDA(cudaMemcpyAsync(sp_keys.p, k_in.p, uniq * 8, cudaMemcpyDeviceToDevice, st));
    WGA_CUDA(cudaMemcpyAsync(sp_counts.p, c_in.p, uniq * 8, cudaMemcpyDeviceToDevice, st));
    WGA_CUDA(cudaStreamSynchronize(st));
    sp_n = uniq;
  }
  static void fill_ones(uint64_t* d, uint64_t n, cudaStream_t st);
};

__global__ void k_fill_ones(uint64_t* d, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = 1;
}
void ModelBuilder::Impl::fill_ones(uint64_t* d, uint64_t n, cudaStream_t st) {
  k_fill_ones<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, n);
  count_launch();
}

ModelBuilder::ModelBuilder() : impl_(new Impl()) {}
ModelBuilder::~ModelBuilder() { delete impl_; }

uint64_t* ModelBuilder::device_bins() {
  impl_->init();
  return (uint64_t*)impl_->d_bins;
}

void ModelBuilder::accumulate_device(const uint8_t* d_comps, const uint64_t* d_syms, uint64_t n, cudaStream_t st) {
  impl_->init();
  if (n == 0) return;
  DevBuf& big_keys = impl_->big_keys;
  DevBuf& big_count = impl_->big_count;
  big_keys.ensure(n * 8);
  big_count.ensure(8);
  WGA_CUDA(cudaMemsetAsync(big_count.p, 0, 8, st));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint64_t want = (n + HIST_TPB - 1) / HIST_TPB;
  unsigned grid = (unsigned)std::min<uint64_t>(want, (uint64_t)sms * 8);
  k_histogram<<<grid, HIST_TPB, 0, st>>>(d_comps, d_syms, n, impl_->d_bins, big_keys.as<uint64_t>(),
                                         big_count.as<unsigned long long>(), impl_->d_err);
  count_launch();
  WGA_CUDA(cudaGetLastError());
  uint64_t nbig = 0;
  uint32_t err = 0;
  WGA_CUDA(cudaMemcpyAsync(&nbig, big_count.p, 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(&err, impl_->d_err, 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  if (err) {
    WGA_CUDA(cudaMemset(impl_->d_err, 0, 4));
    throw Error(WGA_E_ARG, "Symbol can't be bigger than u48::MAX");
  }
  impl_->merge_sparse(big_keys.as<uint64_t>(), nullptr, nbig, st);
}

void ModelBuilder::accumulate_host(const uint8_t* h_comps, const uint64_t* h_syms, uint64_t n) {
  impl_->init();
  const uint64_t CH = 1ull << 26;  // 64M symbols per slice
  DevBuf dc, ds;
  for (uint64_t a = 0; a < n; a += CH) {
    uint64_t m = std::min(CH, n - a);
    dc.ensure(m);
    ds.ensure(m * 8);
    WGA_CUDA(cudaMemcpy(dc.p, h_comps + a, m, cudaMemcpyHostToDevice));
    WGA_CUDA(cudaMemcpy(ds.p, h_syms + a, m * 8, cudaMemcpyHostToDevice));
    accumulate_device(dc.as<uint8_t>(), ds.as<uint64_t>(), m, 0);
  }
}

uint64_t ModelBuilder::sparse_count() {
  impl_->init();
  return impl_->sp_n;
}

void ModelBuilder::sparse_export(uint8_t* h_comps, uint64_t* h_syms, uint64_t* h_counts) {
  impl_->init();
  const uint64_t n = impl_->sp_n;
  if (!n) return;
  std::vector<uint64_t> keys(n);
  WGA_CUDA(cudaMemcpy(keys.data(), impl_->sp_keys.p, n * 8, cudaMemcpyDeviceToHost));
  WGA_CUDA(cudaMemcpy(h_counts, impl_->sp_counts.p, n * 8, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < n; ++i) {
    h_comps[i] = (uint8_t)(keys[i] >> 48);
    h_syms[i] = keys[i] & MAX_RAW;
  }
}

void ModelBuilder::sparse_merge(const uint8_t* h_comps, const uint64_t* h_syms, const uint64_t* h_counts,
                                uint64_t n) {
  impl_->init();
  if (!n) return;
  std::vector<uint64_t> keys(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (h_comps[i] >= NC || h_syms[i] > MAX_RAW) throw Error(WGA_E_ARG, "Symbol can't be bigger than u48::MAX");
    keys[i] = ((uint64_t)h_comps[i] << 48) | h_syms[i];
  }
  DevBuf dk, dc;
  dk.ensure(n * 8);
  dc.ensure(n * 8);
  WGA_CUDA(cudaMemcpy(dk.p, keys.data(), n * 8, cudaMemcpyHostToDevice));
  WGA_CUDA(cudaMemcpy(dc.p, h_counts, n * 8, cudaMemcpyHostToDevice));
  impl_->merge_sparse(dk.as<uint64_t>(), dc.as<uint64_t>(), n, 0);
}

void ModelBuilder::build(ComponentModel out[WGA_COMPONENTS], double* h_original_cost9, double* h_final_cost9) {
  impl_->init();
  cudaStream_t st = 0;
  const size_t NSEG = (size_t)NC * NP;
  DevBuf folded, keys_a, keys_b, nnz, biggest, totals, totals_u, cost, ok, occ, approx, approx_all, entries, table_len,
      params, final_cost, tmp;
  folded.ensure(NSEG * FOLD_MAX * 8);
  keys_a.ensure(NSEG * FOLD_MAX * 8);
  keys_b.ensure(NSEG * FOLD_MAX * 8);
  nnz.ensure(NSEG * 4);
  biggest.ensure(NSEG * 4);
  totals.ensure(NC * 8);
  totals_u.ensure(NC * 8);
  cost.ensure(NSEG * NF * 8);
  ok.ensure(NSEG * NF);
  occ.ensure(NC * 8);
  approx.ensure((size_t)NC * FOLD_MAX * 4);
  entries.ensure((size_t)NC * FOLD_MAX * 8);
  table_len.ensure(NC * 4);
  params.ensure(NC * 3 * 4);
  final_cost.ensure(NC * 8);
  WGA_CUDA(cudaMemsetAsync(folded.p, 0, NSEG * FOLD_MAX * 8, st));
  k_totals<<<NC, 256, 0, st>>>(impl_->d_bins, totals.as<double>(), totals_u.as<unsigned long long>());
  count_launch();
  // ---- raw entropy per component (:275-289)
  {
    size_t tb = 0, tb2 = 0;
    for (int c = 0; c < NC; ++c) {
      RawCostDense fd{impl_->d_bins, totals.as<double>(), c};
      cub::TransformInputIterator<double, RawCostDense, cub::CountingInputIterator<uint32_t>> itd(
          cub::CountingInputIterator<uint32_t>(0), fd);
      // dense part: exact bins of values < 1024 ; sparse part: distinct values >= 1024
      DevBuf part;
      part.ensure(16);
      WGA_CUDA(cub::DeviceReduce::Sum(nullptr, tb, itd, part.as<double>(), 1024, st));
      tmp.ensure(tb);
      WGA_CUDA(cub::DeviceReduce::Sum(tmp.p, tb, itd, part.as<double>(), 1024, st));
      double h_part[2] = {0.0, 0.0};
      if (impl_->sp_n) {
        RawCostSparse fs{impl_->sp_keys.as<uint64_t>(), impl_->sp_counts.as<uint64_t>(), totals.as<double>(), c};
        cub::TransformInputIterator<double, RawCostSparse, cub::CountingInputIterator<uint64_t>> its(
            cub::CountingInputIterator<uint64_t>(0), fs);
        WGA_CUDA(cub::DeviceReduce::Sum(nullptr, tb2, its, part.as<double>() + 1, (int64_t)impl_->sp_n, st));
        tmp.ensure(tb2);
        WGA_CUDA(cub::DeviceReduce::Sum(tmp.p, tb2, its, part.as<double>() + 1, (int64_t)impl_->sp_n, st));
        count_launch(2);
      } else {
        WGA_CUDA(cudaMemsetAsync(part.as<double>() + 1, 0, 8, st));
      }
      count_launch(2);
      WGA_CUDA(cudaMemcpyAsync(h_part, part.p, 16, cudaMemcpyDeviceToHost, st));
      WGA_CUDA(cudaStreamSynchronize(st));
      double v = h_part[0] + h_part[1];
      WGA_CUDA(cudaMemcpyAsync(occ.as<double>() + c, &v, 8, cudaMemcpyHostToDevice, st));
      if (h_original_cost9) h_original_cost9[c] = v;
      WGA_CUDA(cudaStreamSynchronize(st));
    }
  }
  // ---- foldings, sort, scale, select
  k_fold<<<dim3(NP, NC), 256, 0, st>>>(impl_->d_bins, folded.as<unsigned long long>());
  k_keys<<<dim3(NP, NC), 256, 0, st>>>(folded.as<unsigned long long>(), keys_a.as<uint64_t>(), nnz.as<uint32_t>(),
                                       biggest.as<uint32_t>());
  count_launch(2);
  {
    cub::TransformInputIterator<int64_t, SegBegin, cub::CountingInputIterator<int>> sb(
        cub::CountingInputIterator<int>(0), SegBegin());
    cub::TransformInputIterator<int64_t, SegEnd, cub::CountingInputIterator<int>> se(
        cub::CountingInputIterator<int>(0), SegEnd());
    size_t tb = 0;
    WGA_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(),
                                                (int64_t)(NSEG * FOLD_MAX), (int64_t)NSEG, sb, se, st));
    tmp.ensure(tb);
    WGA_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp.p, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(),
                                                (int64_t)(NSEG * FOLD_MAX), (int64_t)NSEG, sb, se, st));
    count_launch(3);
  }
  approx_all.ensure((size_t)NP * NF * FOLD_MAX * 4);
  for (int c = 0; c < NC; ++c)
    k_scale<<<NP, 32, 0, st>>>(c, keys_b.as<uint64_t>(), nnz.as<uint32_t>(), biggest.as<uint32_t>(),
                               folded.as<unsigned long long>(), totals_u.as<unsigned long long>(),
                               approx_all.as<uint32_t>(), cost.as<double>(), ok.as<uint8_t>());
  count_launch(NC - 1);
  EmitOut eo{entries.as<wga_encoder_entry>(), table_len.as<uint32_t>(), params.as<uint32_t>(), final_cost.as<double>()};
  k_select_emit<<<NC, 256, 0, st>>>(keys_b.as<uint64_t>(), nnz.as<uint32_t>(), biggest.as<uint32_t>(),
                                    totals_u.as<unsigned long long>(), cost.as<double>(), ok.as<uint8_t>(),
                                    occ.as<double>(), approx.as<uint32_t>(), eo);
  count_launch(2);
  WGA_CUDA(cudaGetLastError());
  // ---- results to the host
  uint32_t h_len[NC], h_params[NC * 3];
  double h_final[NC];
  WGA_CUDA(cudaMemcpyAsync(h_len, table_len.p, NC * 4, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(h_params, params.p, NC * 12, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaMemcpyAsync(h_final, final_cost.p, NC * 8, cudaMemcpyDeviceToHost, st));
  WGA_CUDA(cudaStreamSynchronize(st));
  for (int c = 0; c < NC; ++c) {
    if (h_params[c * 3] == 0xFFFFFFFFu)
      throw Error(WGA_E_ARG, "It's not been possible to approximate the folded distribution for component " +
                                 std::to_string(c) + " with any radix/fidelity and a frame size <= 2^16");
    ComponentModel& m = out[c];
    m.table.resize(h_len[c]);
    if (h_len[c])
      WGA_CUDA(cudaMemcpy(m.table.data(), entries.as<wga_encoder_entry>() + (size_t)c * FOLD_MAX, (size_t)h_len[c] * 8,
                          cudaMemcpyDeviceToHost));
    m.frame_size = h_params[c * 3];
    m.fidelity = h_params[c * 3 + 1];
    m.radix = h_params[c * 3 + 2];
    if (h_len[c] == 0 && m.frame_size == 0 && m.fidelity == 2 && m.radix == 2) {
      m.folding_threshold = 10;  // Default (component_model4encoder.rs:59-70)
      m.folding_offset = 10;
    } else {
      m.folding_threshold = 1ull << (m.fidelity + m.radix - 1);                  // :231
      m.folding_offset = ((1ull << m.radix) - 1) * (1ull << (m.fidelity - 1));   // :232
    }
    if (h_final_cost9) h_final_cost9[c] = h_final[c];
  }
}

}  // namespace wga
