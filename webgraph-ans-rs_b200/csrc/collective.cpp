// Model-build collective behind the C ABI (wga_hist_allreduce): the dense canonical histograms are
// summed with ncclAllReduce, the sparse tail of large raw symbols is exchanged with ncclAllGather and merged, so
// that every rank then builds the identical model (model4encoder_builder.rs:80-271 needs global frequencies).
// NCCL is not a link-time dependency: the process that calls this already has it loaded (torch bundles it, a C
// caller links it to create the communicator), so the few entry points are resolved with dlsym at first use.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.hpp"
#include "../../include/wga.h"

using wga::Error;
using wga::guarded;

namespace {

struct UniqueId { char internal[128]; };  // ncclUniqueId
typedef int (*fn_get_unique_id)(UniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, UniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_comm_count)(void*, int*);
typedef int (*fn_comm_user_rank)(void*, int*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_error_string)(int);
constexpr int NCCL_UINT64 = 5, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t (nccl.h)

struct Nccl {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_comm_count comm_count = nullptr;
  fn_comm_user_rank comm_user_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_error_string error_string = nullptr;
};

const Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    // already loaded by the caller (torch, or the application that made the communicator)?  else the default path
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    n.lib = h;
    n.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    n.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    n.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    n.comm_count = (fn_comm_count)dlsym(h, "ncclCommCount");
    n.comm_user_rank = (fn_comm_user_rank)dlsym(h, "ncclCommUserRank");
    n.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    n.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    n.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
  });
  if (!n.lib || !n.all_reduce || !n.all_gather || !n.comm_count || !n.comm_user_rank)
    throw Error(WGA_E_UNSUPPORTED, "NCCL (libnccl.so.2) is not available in this process");
  return n;
}

void check(int rc, const char* what) {
  if (rc != 0) {
    const Nccl& n = nccl();
    throw Error(WGA_E_CUDA, std::string(what) + ": " + (n.error_string ? n.error_string(rc) : "NCCL error"));
  }
}
void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Error(WGA_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct DevMem {
  void* p = nullptr;
  explicit DevMem(size_t bytes) { cuda_check(cudaMalloc(&p, std::max<size_t>(bytes, 8)), "cudaMalloc"); }
  ~DevMem() { cudaFree(p); }
};

}  // namespace

extern "C" {

int wga_nccl_get_unique_id(void* out128) {
  return guarded([&] {
    if (!out128) throw Error(WGA_E_ARG, "null argument");
    const Nccl& n = nccl();
    if (!n.get_unique_id) throw Error(WGA_E_UNSUPPORTED, "ncclGetUniqueId not found");
    UniqueId id;
    check(n.get_unique_id(&id), "ncclGetUniqueId");
    std::memcpy(out128, &id, sizeof(id));
  });
}

int wga_nccl_comm_init(int n_ranks, int rank, const void* id128, void** out_comm) {
  return guarded([&] {
    if (!id128 || !out_comm) throw Error(WGA_E_ARG, "null argument");
    const Nccl& n = nccl();
    if (!n.comm_init_rank) throw Error(WGA_E_UNSUPPORTED, "ncclCommInitRank not found");
    UniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    check(n.comm_init_rank(out_comm, n_ranks, id, rank), "ncclCommInitRank");
  });
}

void wga_nccl_comm_destroy(void* comm) {
  try {
    if (comm && nccl().comm_destroy) nccl().comm_destroy(comm);
  } catch (...) {
  }
}

int wga_model_allreduce(wga_model* m, void* nccl_comm, void* stream) {
  return guarded([&] {
    if (!m || !nccl_comm) throw Error(WGA_E_ARG, "null argument");
    const Nccl& n = nccl();
    cudaStream_t st = (cudaStream_t)stream;
    int world = 0, me = 0;
    check(n.comm_count(nccl_comm, &world), "ncclCommCount");
    check(n.comm_user_rank(nccl_comm, &me), "ncclCommUserRank");
    if (world <= 1) return;
    // ---- sparse tail of this rank, taken before anything is merged
    const uint64_t mine = wga_model_sparse_count(m);
    std::vector<uint8_t> comps(mine);
    std::vector<uint64_t> syms(mine), cnts(mine);
    if (wga_model_sparse_export(m, comps.data(), syms.data(), cnts.data()) != WGA_OK) throw Error(WGA_E_CUDA, wga_last_error());
    // ---- dense canonical bins: one all-reduce, in place
    uint64_t* bins = wga_model_bins(m);
    check(n.all_reduce(bins, bins, (size_t)WGA_COMPONENTS * WGA_CANON_BINS, NCCL_UINT64, NCCL_SUM, nccl_comm, st), "ncclAllReduce");
    // ---- sparse tail: all-gather of the counts, then of (component, symbol, count) triples padded to the largest
    DevMem d_n(8), d_all_n(8 * (size_t)world);
    cuda_check(cudaMemcpyAsync(d_n.p, &mine, 8, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
    check(n.all_gather(d_n.p, d_all_n.p, 1, NCCL_UINT64, nccl_comm, st), "ncclAllGather");
    std::vector<uint64_t> all_n(world);
    cuda_check(cudaMemcpyAsync(all_n.data(), d_all_n.p, 8 * (size_t)world, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    const uint64_t mx = *std::max_element(all_n.begin(), all_n.end());
    if (mx == 0) return;
    std::vector<uint64_t> send(3 * mx, 0), recv(3 * mx * (size_t)world);
    for (uint64_t i = 0; i < mine; ++i) {
      send[3 * i] = comps[i];
      send[3 * i + 1] = syms[i];
      send[3 * i + 2] = cnts[i];
    }
    DevMem d_send(24 * mx), d_recv(24 * mx * (size_t)world);
    cuda_check(cudaMemcpyAsync(d_send.p, send.data(), 24 * mx, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync");
    check(n.all_gather(d_send.p, d_recv.p, 3 * mx, NCCL_UINT64, nccl_comm, st), "ncclAllGather");
    cuda_check(cudaMemcpyAsync(recv.data(), d_recv.p, 24 * mx * (size_t)world, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    cuda_check(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    for (int r = 0; r < world; ++r) {
      if (r == me || all_n[r] == 0) continue;
      const uint64_t k = all_n[r];
      std::vector<uint8_t> c(k);
      std::vector<uint64_t> s(k), f(k);
      const uint64_t* src = recv.data() + 3 * mx * (size_t)r;
      for (uint64_t i = 0; i < k; ++i) {
        c[i] = (uint8_t)src[3 * i];
        s[i] = src[3 * i + 1];
        f[i] = src[3 * i + 2];
      }
      if (wga_model_sparse_merge(m, c.data(), s.data(), f.data(), k) != WGA_OK) throw Error(WGA_E_CUDA, wga_last_error());
    }
  });
}

}  // extern "C"
