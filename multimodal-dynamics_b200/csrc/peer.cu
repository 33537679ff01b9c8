// Data-parallel gradient exchange fused with the optimizer over NVLink / NVSwitch peer memory (north_star: "the batch
// shards naturally across the 8 GPUs of one box"; the reference itself has no distributed code, its step is
// problems.py:150-155):
//
//   reduce-scatter + Adam + all-gather in ONE kernel.  Rank r owns the r-th contiguous shard of the flat parameter
//   arena.  For every float4 of its shard it loads the gradient from all N ranks' gradient arenas (its own from HBM,
//   the others as peer loads through NVLink), sums them in rank order, applies the Adam update to its local
//   parameter / moment entries, and stores the new parameter into all N ranks' parameter arenas (peer stores).
//   Each parameter is computed by exactly one rank, so replicas stay bit-identical; the moments exist only where
//   they are used.  NVLink traffic per rank and step: (N-1)/N x arena bytes in, the same out, both directions at once.
//
// Synchronisation is part of the kernel (no host round trip, CUDA-graph friendly): every rank owns a block of
// 2N 32-bit flags that its peers write with system-scope stores.
//   ready[p] = e   peer p's gradients of step e are final      (written at the start of p's kernel)
//   done[p]  = e   peer p has finished reading my gradients and writing my parameters for step e
// The epoch e comes from a device-resident counter, so graph replays advance it.  A rank leaves the kernel only
// after all peers reported done: the kernels that follow on its stream (next step's forward, zero_grad) may then
// touch parameters and gradients freely.
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>
#include <cstring>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;
namespace {

constexpr int MAX_PEERS = 8;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin: a lost peer becomes a trap (sticky CUDA error) instead of a GPU that never comes back
__device__ __forceinline__ void wait_flag(const unsigned int* p, unsigned int epoch) {
  const long long t0 = clock64();
  while (static_cast<int>(ld_acquire_sys(p) - epoch) < 0) {
    if (clock64() - t0 > 120000000000LL) __trap();  // ~60 s: ranks may be a graph capture or a profiling pass apart
  }
}

struct PeerArgs {
  float* grad[MAX_PEERS];
  float* param[MAX_PEERS];
  unsigned int* flags[MAX_PEERS];  // flags[p] = rank p's flag block: ready[0..N), done[0..N)
};

__global__ void __launch_bounds__(256)
peer_rs_adam_ag_kernel(const __grid_constant__ PeerArgs pa, float* __restrict__ m, float* __restrict__ v, long long n4,
                       int rank, int world, float lr, float b1, float b2, float eps, float wd,
                       const uint64_t* __restrict__ step_dev, const uint64_t* __restrict__ epoch_dev, float gscale,
                       unsigned int* __restrict__ nonfinite, unsigned int* __restrict__ block_counter) {
  pdl_sync();
  const unsigned int epoch = static_cast<unsigned int>(*epoch_dev);
  // ---- my gradients are final (stream order): tell every peer, then wait for theirs ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pa.flags[threadIdx.x] + rank, epoch);
  }
  if (threadIdx.x < world) wait_flag(pa.flags[rank] + threadIdx.x, epoch);
  __syncthreads();

  const double t = static_cast<double>(*step_dev);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), t));
  const float bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), t)));
  const float step = lr / bc1;
  const long long per = (n4 + world - 1) / world;
  const long long lo = per * rank, hi = min(n4, lo + per);
  bool bad = false;
  for (long long i = lo + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 gp[MAX_PEERS];
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) gp[p] = __ldcg(reinterpret_cast<const float4*>(pa.grad[p]) + i);  // all peer loads in flight
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) {  // fixed rank order: every replica of a run sums identically
        gs.x += gp[p].x; gs.y += gp[p].y; gs.z += gp[p].z; gs.w += gp[p].w;
      }
    float4 pp = reinterpret_cast<float4*>(pa.param[rank])[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa4 = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gs);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // same arithmetic as adam_kernel (elementwise.cu)
      float gr = ga[q] * gscale;
      if (!isfinite(gr)) {
        bad = true;
        continue;
      }
      if (wd != 0.0f) gr = fmaf(wd, pa4[q], gr);
      ma[q] = fmaf(b1, ma[q], (1.0f - b1) * gr);
      va[q] = fmaf(b2, va[q], (1.0f - b2) * gr * gr);
      const float denom = sqrtf(va[q]) / bc2_sqrt + eps;
      pa4[q] -= step * (ma[q] / denom);
    }
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) __stcg(reinterpret_cast<float4*>(pa.param[p]) + i, pp);  // all-gather: every replica gets the entry
  }
  if (bad && nonfinite) atomicOr(nonfinite, 1u);

  // ---- all my peer loads / stores are issued: the last block to get here reports done and waits for the peers ----
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(block_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *block_counter = 0;  // ready for the next launch / graph replay
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(pa.flags[threadIdx.x] + world + rank, epoch);
      wait_flag(pa.flags[rank] + world + threadIdx.x, epoch);
    }
  }
}

}  // namespace
}  // namespace mmdyn

using namespace mmdyn;

extern "C" int mmdyn_enable_peer_access(int peer_device) {
  int cur = 0, can = 0;
  MMDYN_CHECK_CUDA(cudaGetDevice(&cur));
  if (peer_device == cur) return MMDYN_OK;
  MMDYN_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  MMDYN_REQUIRE(can, "enable_peer_access: device %d cannot access device %d (no NVLink / PCIe peer path)", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return MMDYN_OK;
  }
  MMDYN_CHECK_CUDA(e);
  return MMDYN_OK;
}

// CUDA IPC plumbing for the peer-mapped arenas.  The handle names the cudaMalloc allocation that CONTAINS ptr (device
// tensors of the host framework are sub-allocations of larger segments), offset = ptr - allocation base.  The import
// runs with the IMPORTING rank's device current and cudaIpcMemLazyEnablePeerAccess, which maps the peer's memory
// into this device's address space for kernel loads / stores over NVLink (a mapping made under the exporting
// device's index is only good for copies).
extern "C" int mmdyn_ipc_export(const void* ptr, void* handle_out, long long* offset_out) {
  MMDYN_REQUIRE(ptr && handle_out && offset_out, "ipc_export: null pointer");
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_last_error("ipc_export: cuMemGetAddressRange driver entry point not available");
      return MMDYN_ERR_CUDA;
    }
    range_fn = reinterpret_cast<RangeFn>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) {
    set_last_error("ipc_export: cuMemGetAddressRange failed with CUresult %d", static_cast<int>(r));
    return MMDYN_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MMDYN_CHECK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), reinterpret_cast<void*>(base)));
  *offset_out = static_cast<long long>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return MMDYN_OK;
}

extern "C" int mmdyn_ipc_import(const void* handle, long long offset, void** ptr_out) {
  MMDYN_REQUIRE(handle && ptr_out && offset >= 0, "ipc_import: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* base = nullptr;
  MMDYN_CHECK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = static_cast<char*>(base) + offset;
  return MMDYN_OK;
}

extern "C" int mmdyn_peer_rs_adam_ag(float* const* grad_ptrs, float* const* param_ptrs, unsigned int* const* flag_ptrs,
                                     float* m, float* v, long long n, int rank, int world, float lr, float beta1,
                                     float beta2, float eps, float weight_decay, const uint64_t* step_dev,
                                     const uint64_t* epoch_dev, float gscale, unsigned int* nonfinite_flag,
                                     unsigned int* block_counter, void* stream) {
  MMDYN_REQUIRE(grad_ptrs && param_ptrs && flag_ptrs && m && v && step_dev && epoch_dev && block_counter,
                "peer_rs_adam_ag: null pointer");
  MMDYN_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "peer_rs_adam_ag: rank %d of %d (<= %d ranks)",
                rank, world, MAX_PEERS);
  MMDYN_REQUIRE(n > 0 && n % 4 == 0, "peer_rs_adam_ag: n=%lld must be a positive multiple of 4", n);
  PeerArgs pa = {};
  for (int p = 0; p < world; ++p) {
    MMDYN_REQUIRE(grad_ptrs[p] && param_ptrs[p] && flag_ptrs[p] &&
                      ((reinterpret_cast<uintptr_t>(grad_ptrs[p]) | reinterpret_cast<uintptr_t>(param_ptrs[p])) & 15) == 0,
                  "peer_rs_adam_ag: arena of rank %d missing or not 16-byte aligned", p);
    pa.grad[p] = grad_ptrs[p];
    pa.param[p] = param_ptrs[p];
    pa.flags[p] = flag_ptrs[p];
  }
  const long long n4 = n >> 2, per = (n4 + world - 1) / world;
  long long blocks = (per + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  MMDYN_LAUNCH((peer_rs_adam_ag_kernel), static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream), 
      pa, m, v, n4, rank, world, lr, beta1, beta2, eps, weight_decay, step_dev, epoch_dev, gscale, nonfinite_flag,
      block_counter);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}
