// HBM-bound kernels of the cnn-vae / cnn-mvae step: grouped BatchNorm statistics, BN+Swish
// forward/backward, fc tail (Swish + Dropout), ProductOfExperts + reparametrisation + KL,
// BCE / MSE reconstruction losses, weight packing, fused Adam / SGD, Philox RNG.
//
// All of them are coalesced, 16-byte vectorised where the layout allows, reduce with warp
// shuffles -> shared memory -> one atomic per block and channel, and keep fp32 math throughout
// (fp16 is a storage format for activations only).
//
// Reference semantics: mmdyn/pytorch/models/vae.py:52-61, 201-214, 269-276, 311-334 and
// mmdyn/pytorch/problems/problems.py:130-138, 401-458.
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;

namespace {

#define LAUNCHED()                                            \
  do {                                                        \
    g_launch_count.fetch_add(1, std::memory_order_relaxed);   \
    MMDYN_CHECK_CUDA(cudaGetLastError());                     \
  } while (0)

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int RED_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// BatchNorm statistics: per (group, channel) sum and sum of squares over rows.
// grid = (chunks, G); thread -> (8-channel vector, row lane); block partials -> atomics.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
bn_stats_kernel(const __half* __restrict__ x, float* __restrict__ sums, int rows_per_group, int C,
                int rows_per_chunk) {
  pdl_sync();
  extern __shared__ float red[];  // [row_lanes][C][2]
  const int vpr = C >> 3;
  const int row_lanes = RED_THREADS / vpr;
  const int vec = threadIdx.x % vpr, rl = threadIdx.x / vpr;
  const int g = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(rows_per_group, r_begin + rows_per_chunk);
  const __half* xg = x + static_cast<long long>(g) * rows_per_group * C;
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.0f;
  if (rl < row_lanes) {
    for (int r = r_begin + rl; r < r_end; r += row_lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(xg + static_cast<long long>(r) * C + vec * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        ss[i] = fmaf(f[i], f[i], ss[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[(rl * C + vec * 8 + i) * 2] = s[i];
      red[(rl * C + vec * 8 + i) * 2 + 1] = ss[i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 2; idx += RED_THREADS) {
    float a = 0.0f;
    for (int l = 0; l < row_lanes; ++l) a += red[l * C * 2 + idx];
    atomicAdd(sums + static_cast<long long>(g) * C * 2 + idx, a);
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ ab,
                                   float* __restrict__ mean_invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int G, int n, int C, float eps,
                                   float momentum, int stat_repeat, long long* __restrict__ num_batches_tracked) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += static_cast<long long>(G) * stat_repeat;
  if (c >= C) return;
  const float inv_n = 1.0f / static_cast<float>(n);
  const float unbias = n > 1 ? static_cast<float>(n) / static_cast<float>(n - 1) : 1.0f;
  float rm = running_mean ? running_mean[c] : 0.0f;
  float rv = running_var ? running_var[c] : 0.0f;
  const float ga = gamma[c], be = beta[c];
  for (int g = 0; g < G; ++g) {
    const float s = sums[(g * C + c) * 2], ss = sums[(g * C + c) * 2 + 1];
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.0f);
    const float invstd = rsqrtf(var + eps);
    const float a = ga * invstd;
    ab[(g * C + c) * 2] = a;
    ab[(g * C + c) * 2 + 1] = be - mean * a;
    mean_invstd[(g * C + c) * 2] = mean;
    mean_invstd[(g * C + c) * 2 + 1] = invstd;
    for (int q = 0; q < stat_repeat; ++q) {
      rm = (1.0f - momentum) * rm + momentum * mean;
      rv = (1.0f - momentum) * rv + momentum * var * unbias;
    }
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// y = swish(a*x + b), 8 channels per 16-byte vector, 4 vectors in flight per thread
constexpr int EW_UNROLL = 4;

__global__ void __launch_bounds__(256)
bn_swish_fwd_kernel(const __half* __restrict__ x, const float* __restrict__ ab, __half* __restrict__ y,
                    long long n_vec, int rows_per_group, int C) {
  pdl_sync();
  const int vpr = C >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long v0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v0 < n_vec;
       v0 += stride * EW_UNROLL) {
    uint4 in[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long v = v0 + u * stride;
      if (v < n_vec) in[u] = __ldcs(reinterpret_cast<const uint4*>(x) + v);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const long long v = v0 + u * stride;
      if (v >= n_vec) continue;
      float f[8];
      unpack8(in[u], f);
      if (ab) {
        const long long row = v / vpr;
        const int c0 = static_cast<int>(v - row * vpr) * 8;
        const int g = static_cast<int>(row / rows_per_group);
        const float4* p = reinterpret_cast<const float4*>(ab + (static_cast<long long>(g) * C + c0) * 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = __ldg(p + q);
          f[2 * q] = swishf_(fmaf(t.x, f[2 * q], t.y));
          f[2 * q + 1] = swishf_(fmaf(t.z, f[2 * q + 1], t.w));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = swishf_(f[i]);
      }
      reinterpret_cast<uint4*>(y)[v] = pack8(f);
    }
  }
}

// Same op with the thread -> channel mapping fixed (thread owns one 8-channel vector and walks
// rows), so the per-channel scale/shift live in registers: grid = (chunks, G).
__global__ void __launch_bounds__(RED_THREADS)
bn_swish_fwd_rows_kernel(const __half* __restrict__ x, const float* __restrict__ ab, __half* __restrict__ y,
                         int rows_per_group, int C, int rows_per_chunk) {
  pdl_sync();
  const int vpr = C >> 3;
  const int row_lanes = RED_THREADS / vpr;
  const int vec = threadIdx.x % vpr, rl = threadIdx.x / vpr;
  const int g = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(rows_per_group, r_begin + rows_per_chunk);
  const long long gbase = static_cast<long long>(g) * rows_per_group * C + vec * 8;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = ab[(g * C + vec * 8 + i) * 2];
    b[i] = ab[(g * C + vec * 8 + i) * 2 + 1];
  }
  for (int r = r_begin + rl; r < r_end; r += EW_UNROLL * row_lanes) {
    uint4 in[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u)
      if (r + u * row_lanes < r_end)
        in[u] = __ldcs(reinterpret_cast<const uint4*>(x + gbase + static_cast<long long>(r + u * row_lanes) * C));
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      if (r + u * row_lanes >= r_end) continue;
      float f[8];
      unpack8(in[u], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = swishf_(fmaf(a[i], f[i], b[i]));
      *reinterpret_cast<uint4*>(y + gbase + static_cast<long long>(r + u * row_lanes) * C) = pack8(f);
    }
  }
}

// BatchNorm+Swish backward, pass 1 of 2 (read-only): with dU = dY * swish'(a*x+b),
// sums2[g][c] = {sum dU, sum dU * xhat}.  dU is NOT written back — pass 2 (bn_bwd_apply) recomputes it,
// which trades 2 B/element of HBM writes for a second exp per element.
__global__ void __launch_bounds__(RED_THREADS)
bn_swish_bwd_reduce_kernel(const __half* __restrict__ x, const float* __restrict__ ab,
                           const float* __restrict__ mean_invstd, const __half* __restrict__ dY,
                           float* __restrict__ sums2, int rows_per_group, int C, int rows_per_chunk) {
  pdl_sync();
  extern __shared__ float red[];
  const int vpr = C >> 3;
  const int row_lanes = RED_THREADS / vpr;
  const int vec = threadIdx.x % vpr, rl = threadIdx.x / vpr;
  const int g = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(rows_per_group, r_begin + rows_per_chunk);
  const long long gbase = static_cast<long long>(g) * rows_per_group * C + vec * 8;
  float a[8], b[8], s1[8], s2[8];  // s2 accumulates sum dU * x; the xhat form follows from the totals
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = vec * 8 + i;
    a[i] = ab[(g * C + c) * 2];
    b[i] = ab[(g * C + c) * 2 + 1];
    s1[i] = s2[i] = 0.0f;
  }
  constexpr int U = 4;
  for (int r = r_begin + rl; r < r_end; r += U * row_lanes) {
    uint4 ux[U], ud[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u * row_lanes < r_end) {
        const long long off = gbase + static_cast<long long>(r + u * row_lanes) * C;
        ux[u] = __ldg(reinterpret_cast<const uint4*>(x + off));
        ud[u] = __ldg(reinterpret_cast<const uint4*>(dY + off));
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * row_lanes >= r_end) continue;
      float fx[8], fd[8];
      unpack8(ux[u], fx);
      unpack8(ud[u], fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float du = fd[i] * swish_gradf_(fmaf(a[i], fx[i], b[i]));
        s1[i] += du;
        s2[i] = fmaf(du, fx[i], s2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = vec * 8 + i;
    const float mean = mean_invstd[(g * C + c) * 2], invstd = mean_invstd[(g * C + c) * 2 + 1];
    red[(rl * C + c) * 2] = s1[i];
    red[(rl * C + c) * 2 + 1] = (s2[i] - mean * s1[i]) * invstd;  // sum dU * (x - mean) * invstd
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 2; idx += RED_THREADS) {
    float acc = 0.0f;
    for (int l = 0; l < row_lanes; ++l) acc += red[l * C * 2 + idx];
    atomicAdd(sums2 + static_cast<long long>(g) * C * 2 + idx, acc);
  }
}

// plain Swish backward in place: dX = dY * swish'(x)
__global__ void __launch_bounds__(256)
swish_bwd_kernel(const __half* __restrict__ x, __half* __restrict__ dY, long long n_vec) {
  pdl_sync();
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n_vec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float fx[8], fd[8];
    unpack8(reinterpret_cast<const uint4*>(x)[v], fx);
    unpack8(reinterpret_cast<const uint4*>(dY)[v], fd);
#pragma unroll
    for (int i = 0; i < 8; ++i) fd[i] *= swish_gradf_(fx[i]);
    reinterpret_cast<uint4*>(dY)[v] = pack8(fd);
  }
}

// dX = a * (dU - mean(dU) - xhat * mean(dU * xhat)) in place, rewritten per channel as
// dX = A*dU + B*x + K with A = a, B = -a*m2*invstd, K = a*(m2*invstd*mean - m1): the coefficients
// come from a tiny kernel, the streaming kernel then needs 3 vector loads per 8 channels.
__global__ void bn_bwd_coef_kernel(const float* __restrict__ ab, const float* __restrict__ mean_invstd,
                                   const float* __restrict__ sums2, float* __restrict__ coef, int GC,
                                   float inv_n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GC) return;
  const float a = ab[2 * i], mean = mean_invstd[2 * i], invstd = mean_invstd[2 * i + 1];
  const float m1 = sums2[2 * i] * inv_n, m2 = sums2[2 * i + 1] * inv_n;
  coef[4 * i] = a;
  coef[4 * i + 1] = -a * m2 * invstd;
  coef[4 * i + 2] = a * (m2 * invstd * mean - m1);
  coef[4 * i + 3] = ab[2 * i + 1];  // BatchNorm shift b (dU is recomputed from dY in the apply pass)
}

__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_apply_kernel(const __half* __restrict__ x, const float* __restrict__ coef, __half* __restrict__ dU,
                    int rows_per_group, int C, int rows_per_chunk) {
  pdl_sync();
  const int vpr = C >> 3;
  const int row_lanes = RED_THREADS / vpr;
  const int vec = threadIdx.x % vpr, rl = threadIdx.x / vpr;
  const int g = blockIdx.y;
  const int r_begin = blockIdx.x * rows_per_chunk;
  const int r_end = min(rows_per_group, r_begin + rows_per_chunk);
  const long long gbase = static_cast<long long>(g) * rows_per_group * C + vec * 8;
  float ka[8], kb[8], kc[8], ks[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 k = *reinterpret_cast<const float4*>(coef + (static_cast<long long>(g) * C + vec * 8 + i) * 4);
    ka[i] = k.x;
    kb[i] = k.y;
    kc[i] = k.z;
    ks[i] = k.w;
  }
  constexpr int U = 2;
  for (int r = r_begin + rl; r < r_end; r += U * row_lanes) {
    uint4 ix[U], id[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u * row_lanes < r_end) {
        const long long off = gbase + static_cast<long long>(r + u * row_lanes) * C;
        ix[u] = __ldcs(reinterpret_cast<const uint4*>(x + off));
        id[u] = *reinterpret_cast<const uint4*>(dU + off);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * row_lanes >= r_end) continue;
      float fx[8], fd[8];
      unpack8(ix[u], fx);
      unpack8(id[u], fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float du = fd[i] * swish_gradf_(fmaf(ka[i], fx[i], ks[i]));  // dU = dY * swish'(a*x + b)
        fd[i] = fmaf(ka[i], du, fmaf(kb[i], fx[i], kc[i]));
      }
      *reinterpret_cast<uint4*>(dU + gbase + static_cast<long long>(r + u * row_lanes) * C) = pack8(fd);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-streamed variants of the four BatchNorm streaming kernels.  The register-fed versions above
// keep at most 2-4 16-byte loads per thread in flight (~35 KB per SM at 76-79 registers), which is
// about one bandwidth-delay product of HBM3e and left them latency-bound at 0.4-0.6 of the copy peak.
// Here one thread per CTA streams contiguous 16 KB row blocks into a shared-memory ring with TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx), so 96-128 KB per SM are in flight regardless of the
// register budget; the 256 threads consume a stage with conflict-free 16-byte shared loads (same
// thread -> 8-channel-vector mapping as above, so per-channel coefficients stay in registers).
// ---------------------------------------------------------------------------------------------
struct BnFinalizeArgs {  // sums == nullptr: scale/shift come precomputed (ab)
  const float* sums;
  const float* gamma;
  const float* beta;
  float* ab;
  float* mean_invstd;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float eps, momentum;
  int stat_repeat;
};
struct BnBwdFinalArgs {  // sums2 == nullptr: coefficients come precomputed (coef)
  const float* ab;
  const float* mean_invstd;
  const float* sums2;
  float* dgamma;
  float* dbeta;
  float inv_n, unscale;
};

constexpr int BS_STAGE_BYTES = 16384;
constexpr int BS_ROWS_PER_THREAD = 4;  // rows_per_stage <= 4 * row_lanes for every C (8192/C vs 2048/C)

template <int NIN, int STAGES>
struct BulkRows {
  uint32_t smem, bars;
  const uint8_t* src[NIN];
  long long total_bytes;
  int stage_bytes, n_iter;
  __device__ __forceinline__ void issue(int it) const {
    const int s = it % STAGES;
    const long long off = static_cast<long long>(it) * stage_bytes;
    const long long left = total_bytes - off;
    const uint32_t bytes = static_cast<uint32_t>(left < stage_bytes ? left : stage_bytes);
    const uint32_t bar = bars + s * 8;
    mbar_arrive_expect_tx(bar, NIN * bytes);
#pragma unroll
    for (int i = 0; i < NIN; ++i) bulk_load_1d(smem + (s * NIN + i) * BS_STAGE_BYTES, src[i] + off, bytes, bar);
  }
};

#define BS_PROLOGUE(NIN, STAGES)                                                                      \
  extern __shared__ __align__(128) uint8_t bs_smem[];                                                \
  __shared__ __align__(8) uint64_t bs_bar[STAGES];                                                   \
  const int vpr = C >> 3;                                                                            \
  const int row_lanes = RED_THREADS / vpr;                                                           \
  const int vec = threadIdx.x % vpr, rl = threadIdx.x / vpr;                                         \
  const int g = blockIdx.y;                                                                          \
  const int r_begin = blockIdx.x * rows_per_chunk;                                                   \
  const int n_rows = min(rows_per_group, r_begin + rows_per_chunk) - r_begin;                        \
  const int row_bytes = C * 2;                                                                       \
  const int rps = BS_STAGE_BYTES / row_bytes;                                                        \
  const long long base = (static_cast<long long>(g) * rows_per_group + r_begin) * C;                 \
  BulkRows<NIN, STAGES> st;                                                                          \
  st.smem = smem_u32(bs_smem);                                                                       \
  st.bars = smem_u32(bs_bar);                                                                        \
  st.total_bytes = static_cast<long long>(n_rows) * row_bytes;                                       \
  st.stage_bytes = rps * row_bytes;                                                                  \
  st.n_iter = (n_rows + rps - 1) / rps

#define BS_START(STAGES)                                                                              \
  if (threadIdx.x == 0) {                                                                            \
    for (int s_ = 0; s_ < STAGES; ++s_) mbar_init(st.bars + s_ * 8, 1);                              \
    mbar_fence_init();                                                                               \
  }                                                                                                  \
  __syncthreads();                                                                                   \
  if (threadIdx.x == 0)                                                                              \
    for (int it_ = 0; it_ < STAGES && it_ < st.n_iter; ++it_) st.issue(it_)

template <int STAGES>
__global__ void __launch_bounds__(RED_THREADS)
bn_stats_bulk_kernel(const __half* __restrict__ x, float* __restrict__ sums, int rows_per_group, int C,
                     int rows_per_chunk) {
  pdl_sync();
  BS_PROLOGUE(1, STAGES);
  st.src[0] = reinterpret_cast<const uint8_t*>(x + base);
  BS_START(STAGES);
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.0f;
  for (int it = 0; it < st.n_iter; ++it) {
    const int s = it % STAGES;
    mbar_wait(st.bars + s * 8, (it / STAGES) & 1);
    const int nr = min(rps, n_rows - it * rps);
    const uint8_t* sx = bs_smem + s * BS_STAGE_BYTES + vec * 16;
#pragma unroll
    for (int k = 0; k < BS_ROWS_PER_THREAD; ++k) {
      const int row = rl + k * row_lanes;
      if (row < nr) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(sx + row * row_bytes), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[i] += f[i];
          s2[i] = fmaf(f[i], f[i], s2[i]);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && it + STAGES < st.n_iter) st.issue(it + STAGES);
  }
  float* red = reinterpret_cast<float*>(bs_smem);  // ring is idle: every issued stage was consumed
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[(rl * C + vec * 8 + i) * 2] = s1[i];
    red[(rl * C + vec * 8 + i) * 2 + 1] = s2[i];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 2; idx += RED_THREADS) {
    float a = 0.0f;
    for (int l = 0; l < row_lanes; ++l) a += red[l * C * 2 + idx];
    atomicAdd(sums + static_cast<long long>(g) * C * 2 + idx, a);
  }
}

template <int STAGES>
__global__ void __launch_bounds__(RED_THREADS)
bn_swish_fwd_bulk_kernel(const __half* __restrict__ x, const float* __restrict__ ab, __half* __restrict__ y,
                         int rows_per_group, int C, int rows_per_chunk, BnFinalizeArgs fin) {
  pdl_sync();
  BS_PROLOGUE(1, STAGES);
  st.src[0] = reinterpret_cast<const uint8_t*>(x + base);
  BS_START(STAGES);
  float a[8], b[8];
  if (fin.sums) {
    // bn_finalize folded in: every thread derives scale/shift of its 8 channels from the statistics
    // (same arithmetic everywhere, so all CTAs agree bit for bit); chunk 0 of each group publishes
    // ab / mean_invstd for the backward, CTA (0,0) updates the running statistics in group order
    const float inv_n = 1.0f / static_cast<float>(rows_per_group);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = vec * 8 + i;
      const float sm = fin.sums[(g * C + c) * 2], ss = fin.sums[(g * C + c) * 2 + 1];
      const float mean = sm * inv_n;
      const float var = fmaxf(ss * inv_n - mean * mean, 0.0f);
      const float invstd = rsqrtf(var + fin.eps);
      a[i] = fin.gamma[c] * invstd;
      b[i] = fin.beta[c] - mean * a[i];
      if (blockIdx.x == 0 && rl == 0) {
        fin.ab[(g * C + c) * 2] = a[i];
        fin.ab[(g * C + c) * 2 + 1] = b[i];
        fin.mean_invstd[(g * C + c) * 2] = mean;
        fin.mean_invstd[(g * C + c) * 2 + 1] = invstd;
      }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
      if (threadIdx.x == 0 && fin.num_batches_tracked)
        *fin.num_batches_tracked += static_cast<long long>(gridDim.y) * fin.stat_repeat;
      if (fin.running_mean && fin.running_var) {
        const float n = static_cast<float>(rows_per_group);
        const float unbias = rows_per_group > 1 ? n / (n - 1.0f) : 1.0f;
        for (int c = threadIdx.x; c < C; c += RED_THREADS) {
          float rm = fin.running_mean[c], rv = fin.running_var[c];
          for (int gg = 0; gg < static_cast<int>(gridDim.y); ++gg) {
            const float mean = fin.sums[(gg * C + c) * 2] * inv_n;
            const float var = fmaxf(fin.sums[(gg * C + c) * 2 + 1] * inv_n - mean * mean, 0.0f);
            for (int q = 0; q < fin.stat_repeat; ++q) {
              rm = (1.0f - fin.momentum) * rm + fin.momentum * mean;
              rv = (1.0f - fin.momentum) * rv + fin.momentum * var * unbias;
            }
          }
          fin.running_mean[c] = rm;
          fin.running_var[c] = rv;
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = ab[(g * C + vec * 8 + i) * 2];
      b[i] = ab[(g * C + vec * 8 + i) * 2 + 1];
    }
  }
  for (int it = 0; it < st.n_iter; ++it) {
    const int s = it % STAGES;
    mbar_wait(st.bars + s * 8, (it / STAGES) & 1);
    const int nr = min(rps, n_rows - it * rps);
    const uint8_t* sx = bs_smem + s * BS_STAGE_BYTES + vec * 16;
    __half* out = y + base + static_cast<long long>(it) * rps * C + vec * 8;
#pragma unroll
    for (int k = 0; k < BS_ROWS_PER_THREAD; ++k) {
      const int row = rl + k * row_lanes;
      if (row < nr) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(sx + row * row_bytes), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = swishf_(fmaf(a[i], f[i], b[i]));
        *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * C) = pack8(f);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && it + STAGES < st.n_iter) st.issue(it + STAGES);
  }
}

template <int STAGES>
__global__ void __launch_bounds__(RED_THREADS)
bn_swish_bwd_reduce_bulk_kernel(const __half* __restrict__ x, const float* __restrict__ ab,
                                const float* __restrict__ mean_invstd, const __half* __restrict__ dY,
                                float* __restrict__ sums2, int rows_per_group, int C, int rows_per_chunk) {
  pdl_sync();
  BS_PROLOGUE(2, STAGES);
  st.src[0] = reinterpret_cast<const uint8_t*>(x + base);
  st.src[1] = reinterpret_cast<const uint8_t*>(dY + base);
  BS_START(STAGES);
  float a[8], b[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = vec * 8 + i;
    a[i] = ab[(g * C + c) * 2];
    b[i] = ab[(g * C + c) * 2 + 1];
    s1[i] = s2[i] = 0.0f;
  }
  for (int it = 0; it < st.n_iter; ++it) {
    const int s = it % STAGES;
    mbar_wait(st.bars + s * 8, (it / STAGES) & 1);
    const int nr = min(rps, n_rows - it * rps);
    const uint8_t* sx = bs_smem + (s * 2) * BS_STAGE_BYTES + vec * 16;
    const uint8_t* sd = sx + BS_STAGE_BYTES;
#pragma unroll
    for (int k = 0; k < BS_ROWS_PER_THREAD; ++k) {
      const int row = rl + k * row_lanes;
      if (row < nr) {
        float fx[8], fd[8];
        unpack8(*reinterpret_cast<const uint4*>(sx + row * row_bytes), fx);
        unpack8(*reinterpret_cast<const uint4*>(sd + row * row_bytes), fd);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float du = fd[i] * swish_gradf_(fmaf(a[i], fx[i], b[i]));
          s1[i] += du;
          s2[i] = fmaf(du, fx[i], s2[i]);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && it + STAGES < st.n_iter) st.issue(it + STAGES);
  }
  float* red = reinterpret_cast<float*>(bs_smem);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = vec * 8 + i;
    const float mean = mean_invstd[(g * C + c) * 2], invstd = mean_invstd[(g * C + c) * 2 + 1];
    red[(rl * C + c) * 2] = s1[i];
    red[(rl * C + c) * 2 + 1] = (s2[i] - mean * s1[i]) * invstd;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 2; idx += RED_THREADS) {
    float acc = 0.0f;
    for (int l = 0; l < row_lanes; ++l) acc += red[l * C * 2 + idx];
    atomicAdd(sums2 + static_cast<long long>(g) * C * 2 + idx, acc);
  }
}

template <int STAGES>
__global__ void __launch_bounds__(RED_THREADS)
bn_bwd_apply_bulk_kernel(const __half* __restrict__ x, const float* __restrict__ coef, __half* __restrict__ dU,
                         int rows_per_group, int C, int rows_per_chunk, BnBwdFinalArgs fin,
                         __half* __restrict__ out_pad, int lw) {
  pdl_sync();
  BS_PROLOGUE(2, STAGES);
  st.src[0] = reinterpret_cast<const uint8_t*>(x + base);
  st.src[1] = reinterpret_cast<const uint8_t*>(dU + base);
  BS_START(STAGES);
  float ka[8], kb[8], kc[8], ks[8];
  if (fin.sums2) {
    // bn_bwd_coef and bn_param_grad folded in (see bn_bwd_coef_kernel for the algebra)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gc = g * C + vec * 8 + i;
      const float a = fin.ab[2 * gc], mean = fin.mean_invstd[2 * gc], invstd = fin.mean_invstd[2 * gc + 1];
      const float m1 = fin.sums2[2 * gc] * fin.inv_n, m2 = fin.sums2[2 * gc + 1] * fin.inv_n;
      ka[i] = a;
      kb[i] = -a * m2 * invstd;
      kc[i] = a * (m2 * invstd * mean - m1);
      ks[i] = fin.ab[2 * gc + 1];
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && fin.dgamma && fin.dbeta) {
      for (int c = threadIdx.x; c < C; c += RED_THREADS) {
        float s1 = 0.0f, s2 = 0.0f;
        for (int gg = 0; gg < static_cast<int>(gridDim.y); ++gg) {
          s1 += fin.sums2[(gg * C + c) * 2];
          s2 += fin.sums2[(gg * C + c) * 2 + 1];
        }
        fin.dgamma[c] += fin.unscale * s2;
        fin.dbeta[c] += fin.unscale * s1;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 k = *reinterpret_cast<const float4*>(coef + (static_cast<long long>(g) * C + vec * 8 + i) * 4);
      ka[i] = k.x;
      kb[i] = k.y;
      kc[i] = k.z;
      ks[i] = k.w;
    }
  }
  for (int it = 0; it < st.n_iter; ++it) {
    const int s = it % STAGES;
    mbar_wait(st.bars + s * 8, (it / STAGES) & 1);
    const int nr = min(rps, n_rows - it * rps);
    const uint8_t* sx = bs_smem + (s * 2) * BS_STAGE_BYTES + vec * 16;
    const uint8_t* sd = sx + BS_STAGE_BYTES;
    __half* out = dU + base + static_cast<long long>(it) * rps * C + vec * 8;  // in place: rows of this stage only
    // out_pad: dX goes to a second buffer whose image rows of 2^lw pixels carry one extra (zero, never written) pixel on
    // each side — the layout the stride-2 data / weight gradient kernels read as 128-byte pixel pairs (plan.deconv_s2_plan)
    const long long row0 = static_cast<long long>(g) * rows_per_group + r_begin + static_cast<long long>(it) * rps;
#pragma unroll
    for (int k = 0; k < BS_ROWS_PER_THREAD; ++k) {
      const int row = rl + k * row_lanes;
      if (row < nr) {
        float fx[8], fd[8];
        unpack8(*reinterpret_cast<const uint4*>(sx + row * row_bytes), fx);
        unpack8(*reinterpret_cast<const uint4*>(sd + row * row_bytes), fd);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float du = fd[i] * swish_gradf_(fmaf(ka[i], fx[i], ks[i]));
          fd[i] = fmaf(ka[i], du, fmaf(kb[i], fx[i], kc[i]));
        }
        if (out_pad != nullptr) {
          const long long R = row0 + row;
          *reinterpret_cast<uint4*>(out_pad + (R + ((R >> lw) << 1) + 1) * C + vec * 8) = pack8(fd);
        } else {
          *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * C) = pack8(fd);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && it + STAGES < st.n_iter) st.issue(it + STAGES);
  }
}

__global__ void bn_param_grad_kernel(const float* __restrict__ sums2, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int G, int C, float unscale) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.0f, s2 = 0.0f;
  for (int g = 0; g < G; ++g) {
    s1 += sums2[(g * C + c) * 2];
    s2 += sums2[(g * C + c) * 2 + 1];
  }
  dgamma[c] += unscale * s2;
  dbeta[c] += unscale * s1;
}

// ---------------------------------------------------------------------------------------------
// fc tail
// ---------------------------------------------------------------------------------------------
struct MaskPtrs {
  const float* p[8];
};

__global__ void __launch_bounds__(256)
swish_dropout_fwd_kernel(const float* __restrict__ raw, MaskPtrs masks, __half* __restrict__ h,
                         int n_masks, long long n) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float s = swishf_(raw[i]);
    for (int m = 0; m < n_masks; ++m) {
      const float k = masks.p[m] ? masks.p[m][i] : 1.0f;
      h[static_cast<long long>(m) * n + i] = __float2half_rn(s * k);
    }
  }
}

__global__ void __launch_bounds__(256)
swish_dropout_bwd_kernel(const float* __restrict__ raw, MaskPtrs masks, const float* __restrict__ dH,
                         __half* __restrict__ dRaw, int n_masks, long long n) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.0f;
    for (int m = 0; m < n_masks; ++m) {
      const float k = masks.p[m] ? masks.p[m][i] : 1.0f;
      acc = fmaf(k, dH[static_cast<long long>(m) * n + i], acc);
    }
    dRaw[i] = __float2half_rn(acc * swish_gradf_(raw[i]));
  }
}

// ---------------------------------------------------------------------------------------------
// ProductOfExperts + reparametrisation + KL
// ---------------------------------------------------------------------------------------------
constexpr int MAX_EXPERTS = 4;
struct ExpertPtrs {
  const float* mu[MAX_EXPERTS];
  const float* lv[MAX_EXPERTS];
  float* dmu[MAX_EXPERTS];
  float* dlv[MAX_EXPERTS];
  const float* dz[3];
};
constexpr float POE_EPS = 1e-8f;

__device__ __forceinline__ void poe_fwd_body(const ExpertPtrs& ex, int n_experts, int use_prior, int ld,
                                             const float* __restrict__ eps, float* __restrict__ mu_o,
                                             float* __restrict__ lv_o, float* __restrict__ z_o, __half* __restrict__ zh_o,
                                             __half* __restrict__ zh2_o, float* __restrict__ kl_sum, int B, int D) {
  const long long n = static_cast<long long>(B) * D;
  float kl = 0.0f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / D), col = static_cast<int>(i - static_cast<long long>(row) * D);
    const long long ei = static_cast<long long>(row) * ld + col;
    float mu, lv;
    if (!use_prior && n_experts == 1) {
      mu = ex.mu[0][ei];
      lv = ex.lv[0][ei];
    } else {
      // vae.py:311-318 — eps is added twice on the way in and once on the way out
      float st = 0.0f, sm = 0.0f;
      if (use_prior) {
        const float t0 = 1.0f / ((1.0f + POE_EPS) + POE_EPS);
        st = t0;  // mu_0 = 0
      }
      for (int e = 0; e < n_experts; ++e) {
        const float var = expf(ex.lv[e][ei]) + POE_EPS;
        const float t = 1.0f / (var + POE_EPS);
        st += t;
        sm = fmaf(ex.mu[e][ei], t, sm);
      }
      mu = sm / st;
      lv = logf(1.0f / st + POE_EPS);
    }
    const float std = expf(0.5f * lv);
    const float z = fmaf(eps[i], std, mu);
    mu_o[i] = mu;
    lv_o[i] = lv;
    z_o[i] = z;
    if (zh_o) zh_o[i] = __float2half_rn(z);
    if (zh2_o) zh2_o[i] = __float2half_rn(z);
    kl += 1.0f + lv - mu * mu - expf(lv);
  }
  kl = warp_sum(kl);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += wsum[w];
    atomicAdd(kl_sum, -0.5f * t);
  }
}

__global__ void __launch_bounds__(256)
poe_fwd_kernel(ExpertPtrs ex, int n_experts, int use_prior, int ld, const float* __restrict__ eps,
               float* __restrict__ mu_o, float* __restrict__ lv_o, float* __restrict__ z_o,
               __half* __restrict__ zh_o, __half* __restrict__ zh2_o, float* __restrict__ kl_sum, int B, int D) {
  pdl_sync();
  poe_fwd_body(ex, n_experts, use_prior, ld, eps, mu_o, lv_o, z_o, zh_o, zh2_o, kl_sum, B, D);
}

// All sub-sampled passes of the step in ONE launch (blockIdx.y = pass): 7 launches of a few microseconds each sit on
// the critical path between the encoders and the decoders otherwise.
struct PoePassDev {
  ExpertPtrs ex;
  int n_experts;
  const float* eps;
  float* mu_o;
  float* lv_o;
  float* z_o;
  __half* zh_o;
  __half* zh2_o;
  float* kl_sum;
  const float* dmu_in;
  const float* dlv_in;
};
struct PoePassesDev {
  PoePassDev p[MMDYN_MAX_POE_PASSES];
};

__global__ void __launch_bounds__(256)
poe_fwd_multi_kernel(const __grid_constant__ PoePassesDev ps, int use_prior, int ld, int B, int D) {
  pdl_sync();
  const PoePassDev& q = ps.p[blockIdx.y];
  poe_fwd_body(q.ex, q.n_experts, use_prior, ld, q.eps, q.mu_o, q.lv_o, q.z_o, q.zh_o, q.zh2_o, q.kl_sum, B, D);
}

// ATOMIC: several passes may accumulate into the same expert gradient rows concurrently (multi-pass launch)
template <bool ATOMIC>
__device__ __forceinline__ void poe_store(float* p, float v, int accumulate) {
  if (!accumulate) *p = v;
  else if (ATOMIC) atomicAdd(p, v);
  else *p += v;
}

template <bool ATOMIC>
__device__ __forceinline__ void poe_bwd_body(const ExpertPtrs& ex, int n_experts, int use_prior, int ld,
                                             const float* __restrict__ eps, const float* __restrict__ dmu_in,
                                             const float* __restrict__ dlv_in, float kl_coef, int ld_out, int accumulate,
                                             int B, int D) {
  const long long n = static_cast<long long>(B) * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / D), col = static_cast<int>(i - static_cast<long long>(row) * D);
    const long long ei = static_cast<long long>(row) * ld + col;
    const long long oi = static_cast<long long>(row) * ld_out + col;
    float g = 0.0f;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (ex.dz[q]) g += ex.dz[q][i];
    if (!use_prior && n_experts == 1) {
      const float mu = ex.mu[0][ei], lv = ex.lv[0][ei];
      const float dmu = g + kl_coef * mu + (dmu_in ? dmu_in[i] : 0.0f);
      const float dlv = 0.5f * g * eps[i] * expf(0.5f * lv) + 0.5f * kl_coef * (expf(lv) - 1.0f) +
                        (dlv_in ? dlv_in[i] : 0.0f);
      poe_store<ATOMIC>(ex.dmu[0] + oi, dmu, accumulate);
      poe_store<ATOMIC>(ex.dlv[0] + oi, dlv, accumulate);
      continue;
    }
    float st = 0.0f, sm = 0.0f, t_e[MAX_EXPERTS], elv[MAX_EXPERTS];
    if (use_prior) st = 1.0f / ((1.0f + POE_EPS) + POE_EPS);
    for (int e = 0; e < n_experts; ++e) {
      elv[e] = expf(ex.lv[e][ei]);
      t_e[e] = 1.0f / ((elv[e] + POE_EPS) + POE_EPS);
      st += t_e[e];
      sm = fmaf(ex.mu[e][ei], t_e[e], sm);
    }
    const float mu = sm / st;
    const float pvar = 1.0f / st;
    const float lv = logf(pvar + POE_EPS);
    const float dmu = g + kl_coef * mu + (dmu_in ? dmu_in[i] : 0.0f);
    const float dlv = 0.5f * g * eps[i] * expf(0.5f * lv) + 0.5f * kl_coef * (expf(lv) - 1.0f) +
                      (dlv_in ? dlv_in[i] : 0.0f);
    const float dpvar = dlv / (pvar + POE_EPS);
    const float dsm = dmu / st;
    const float dst = -dpvar * pvar * pvar - dmu * mu / st;
    for (int e = 0; e < n_experts; ++e) {
      const float dmu_e = dsm * t_e[e];
      const float dt = fmaf(dsm, ex.mu[e][ei], dst);
      const float dlv_e = -dt * t_e[e] * t_e[e] * elv[e];
      poe_store<ATOMIC>(ex.dmu[e] + oi, dmu_e, accumulate);
      poe_store<ATOMIC>(ex.dlv[e] + oi, dlv_e, accumulate);
    }
  }
}

__global__ void __launch_bounds__(256)
poe_bwd_kernel(ExpertPtrs ex, int n_experts, int use_prior, int ld, const float* __restrict__ eps,
               const float* __restrict__ dmu_in, const float* __restrict__ dlv_in, float kl_coef, int ld_out,
               int accumulate, int B, int D) {
  pdl_sync();
  poe_bwd_body<false>(ex, n_experts, use_prior, ld, eps, dmu_in, dlv_in, kl_coef, ld_out, accumulate, B, D);
}

__global__ void __launch_bounds__(256)
poe_bwd_multi_kernel(const __grid_constant__ PoePassesDev ps, int use_prior, int ld, float kl_coef, int ld_out,
                     int accumulate, int B, int D) {
  pdl_sync();
  const PoePassDev& q = ps.p[blockIdx.y];
  if (q.n_experts == 0) return;
  poe_bwd_body<true>(q.ex, q.n_experts, use_prior, ld, q.eps, q.dmu_in, q.dlv_in, kl_coef, ld_out, accumulate, B, D);
}

// ---------------------------------------------------------------------------------------------
// losses
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float wsum[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += wsum[w];
  return t;  // valid on thread 0
}

// one thread per 4 consecutive pixels: float4 loads from each of the 3 planes (coalesced), four
// 8-byte NHWC4 gradient stores (32 contiguous bytes)
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                  const float* __restrict__ mask, float* __restrict__ loss_sum,
                  __half* __restrict__ dlogits, float gscale, long long n_pix, int HW, int W, int pad) {
  pdl_sync();
  float acc = 0.0f;
  const long long n_quad = n_pix >> 2;  // W is a multiple of 4 (checked by the launcher)
  const int H = HW / W, Wp = W + 2 * pad, Hp = H + 2 * pad;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n_quad;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = q << 2;
    const long long img = p / HW;
    const int hw = static_cast<int>(p - img * HW);
    float g[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k][3] = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long idx = (img * 3 + c) * HW + hw;
      const float4 xv = __ldcs(reinterpret_cast<const float4*>(logits + idx));
      const float4 tv = __ldg(reinterpret_cast<const float4*>(target + idx));
      float4 mv = make_float4(1.f, 1.f, 1.f, 1.f);
      if (mask) mv = __ldg(reinterpret_cast<const float4*>(mask + idx));
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ts[4] = {tv.x, tv.y, tv.z, tv.w}, ms[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float m = ms[k], x = xs[k] * m, t = ts[k] * m;
        // max(x,0) - x*t + log(1 + exp(-|x|))  (torch's stable form)
        const float e = __expf(-fabsf(x));
        acc += fmaxf(x, 0.0f) - x * t + log1pf(e);
        const float sig = x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
        g[k][c] = gscale * (sig - t) * m;
      }
    }
    if (dlogits) {
      // pad > 0: the gradient image carries a zero border of `pad` pixels (never written here)
      const int y = hw / W, x = hw - y * W;
      const long long o = (img * Hp + y + pad) * Wp + x + pad;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        reinterpret_cast<uint2*>(dlogits)[o + k] = make_uint2(pack_h2(g[k][0], g[k][1]), pack_h2(g[k][2], 0.0f));
    }
  }
  const float t = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(loss_sum, t);
}

// generic-layout BCE-with-logits: grid = (chunks, n samples); per-sample and total sums, fp32 gradient
__global__ void __launch_bounds__(256)
bce_flat_kernel(const float* __restrict__ logits, const float* __restrict__ target, const float* __restrict__ mask,
                float* __restrict__ loss_sum, float* __restrict__ per_sample_sum, float* __restrict__ dlogits,
                float gscale, int per_sample) {
  pdl_sync();
  const long long base = static_cast<long long>(blockIdx.y) * per_sample;
  float acc = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += gridDim.x * blockDim.x) {
    float m = mask ? mask[base + i] : 1.0f;
    const float x = logits[base + i] * m, t = target[base + i] * m;
    const float e = expf(-fabsf(x));
    acc += fmaxf(x, 0.0f) - x * t + log1pf(e);
    if (dlogits) dlogits[base + i] = gscale * ((x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e)) - t) * m;
  }
  const float t = block_sum_256(acc);
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, t);
    if (per_sample_sum) atomicAdd(per_sample_sum + blockIdx.y, t);
  }
}

__global__ void mse_rows_kernel(const float* __restrict__ recon, const float* __restrict__ target,
                                float* __restrict__ row_sum, float mult, int n, int d) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float acc = 0.0f;
  for (int j = 0; j < d; ++j) {
    const float df = recon[static_cast<long long>(r) * d + j] - target[static_cast<long long>(r) * d + j];
    acc = fmaf(df, df, acc);
  }
  row_sum[r] += mult * acc;
}

__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ recon, const float* __restrict__ target, float* __restrict__ loss_sum,
           float* __restrict__ drecon, float mult, float gscale, long long n) {
  pdl_sync();
  float acc = 0.0f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = recon[i] - target[i];
    acc = fmaf(d, d, acc);
    if (drecon) drecon[i] = gscale * 2.0f * mult * d;
  }
  const float t = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(loss_sum, mult * t);
}

// ---------------------------------------------------------------------------------------------
// column sums, packing, casts
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, int ld, float scale,
              int rows_per_cta) {
  pdl_sync();
  // thread -> column (coalesced across the row), loop over the CTA's rows
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc = 0.0f;
  for (int r = r0; r < r1; ++r) acc += x[static_cast<long long>(r) * ld + n];
  atomicAdd(out + n, scale * acc);
}

// fp16 [M][N] column sums: block = 32 8-channel vectors x 8 row lanes
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* __restrict__ x, float* __restrict__ out, int M, int N, int ld, float scale,
                  int rows_per_cta) {
  pdl_sync();
  __shared__ float red[8][256 + 1];
  const int vl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + vl) * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < N) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<long long>(r) * ld + c0), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rl][vl * 8 + i] = s[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float a = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) a += red[l][threadIdx.x];
    atomicAdd(out + c, scale * a);
  }
}

__device__ __forceinline__ __half pack_f16_one(const float* __restrict__ src, int32_t j) {
  if (j < 0) return __float2half_rn(0.0f);
  if (j & (1 << 30)) {
    // low part of the two-term fp16 split of an fp32 weight (pose MLP on the tensor cores, mmdyn_split_f16)
    const float w = __ldg(src + (j & ((1 << 30) - 1)));
    return __float2half_rn(w - __half2float(__float2half_rn(w)));
  }
  return __float2half_rn(__ldg(src + j));
}

// 8 packed elements per thread: two 16-byte index loads, 8 gathers from the (L2-resident) fp32 arena, one 16-byte store
__global__ void __launch_bounds__(256)
pack_f16_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, __half* __restrict__ dst,
                long long n) {
  pdl_sync();
  const long long n8 = n >> 3;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n8;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int4 j0 = __ldg(reinterpret_cast<const int4*>(idx) + 2 * v), j1 = __ldg(reinterpret_cast<const int4*>(idx) + 2 * v + 1);
    const int32_t j[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
    __align__(16) __half h[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) h[q] = pack_f16_one(src, j[q]);
    reinterpret_cast<uint4*>(dst)[v] = *reinterpret_cast<const uint4*>(h);
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) dst[i] = pack_f16_one(src, idx[i]);
}

// Two-term fp16 split of an fp32 matrix for fp32-accurate products on the fp16 tensor cores:
//   x = hi + lo (+ O(2^-22 |x|)),  hi = fp16(x),  lo = fp16(x - hi)
//   x . w  ~=  hi_x hi_w + lo_x hi_w + hi_x lo_w     (three fp16 products, fp32 accumulation; lo_x lo_w ~ 2^-22 dropped)
// which one GEMM computes when the three terms are concatenated along the contraction dimension:
//   mode 0 ("activation" side)  out[m] = [hi | lo | hi]      mode 1 ("other" side)  out[m] = [hi | hi | lo]
// (weights are packed the same way by mmdyn_pack_f16, index bit 30 = low part).  Optional on the way:
//   relu      x := max(x, 0)                       (forward: bias was added by the producing GEMM)
//   mask_y    x := x * (mask_y > 0)                (backward through a ReLU, judged by its output)
//   x_out     the fp32 value after relu / mask     (may alias x)
//   colsum0/1 column sums (times colsum_scale) of the fp32 value: columns < n_split into colsum0, the rest into
//             colsum1 (bias gradients of one or two concatenated Linear layers), accumulated with atomics
// thread = column (coalesced along the row), CTA = a slab of rows.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, const float* __restrict__ mask_y, float* __restrict__ x_out,
                 __half* __restrict__ out, int M, int N, int mode, int relu, float* __restrict__ colsum0,
                 float* __restrict__ colsum1, int n_split, float colsum_scale, int rows_per_cta) {
  pdl_sync();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  const int p_lo = mode == 0 ? 1 : 2;  // segment that holds the low part
  float acc = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const long long i = static_cast<long long>(r) * N + n;
    float v = x[i];
    if (relu) v = fmaxf(v, 0.0f);
    if (mask_y && !(mask_y[i] > 0.0f)) v = 0.0f;
    if (x_out) x_out[i] = v;
    acc += v;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    __half* o = out + static_cast<long long>(r) * 3 * N + n;
    o[0] = hi;
    o[N] = p_lo == 1 ? lo : hi;
    o[2 * N] = p_lo == 2 ? lo : hi;
  }
  if (colsum0) {
    if (n < n_split) atomicAdd(colsum0 + n, colsum_scale * acc);
    else if (colsum1) atomicAdd(colsum1 + (n - n_split), colsum_scale * acc);
  }
}

__global__ void __launch_bounds__(256)
gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst,
                  long long n) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int32_t j = idx[i];
    dst[i] = j < 0 ? 0.0f : src[j];
  }
}

__global__ void __launch_bounds__(256)
unpack_add_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, float* __restrict__ dst,
                  long long n) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int32_t j = idx[i];
    if (j >= 0) dst[j] += src[i];
  }
}

// dst[k] += src[inv[k]] (inv[k] < 0: nothing): the scatter above turned inside out — coalesced
// read-modify-write of the gradient arena, gathered reads of the (L2-resident) packed gradients
__global__ void __launch_bounds__(256)
gather_add_kernel(const float* __restrict__ src, const int32_t* __restrict__ inv, float* __restrict__ dst,
                  long long n) {
  pdl_sync();
  // 4 arena elements per thread when inv / dst are 16-byte aligned (they are slices of 16-byte-aligned arena entries)
  const bool vec = ((reinterpret_cast<uintptr_t>(inv) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n4;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int4 i = __ldg(reinterpret_cast<const int4*>(inv) + v);
    float4 d = reinterpret_cast<float4*>(dst)[v];
    if (i.x >= 0) d.x += __ldg(src + i.x);
    if (i.y >= 0) d.y += __ldg(src + i.y);
    if (i.z >= 0) d.z += __ldg(src + i.z);
    if (i.w >= 0) d.w += __ldg(src + i.w);
    reinterpret_cast<float4*>(dst)[v] = d;
  }
  for (long long k = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int32_t i = __ldg(inv + k);
    if (i >= 0) dst[k] += __ldg(src + i);
  }
}

__global__ void __launch_bounds__(256)
f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n, float scale) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = __float2half_rn(scale * src[i]);
}

__global__ void __launch_bounds__(256)
scale_f32_kernel(float* __restrict__ x, long long n, float s) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    x[i] *= s;
}

__global__ void __launch_bounds__(256)
logit_grad_pack_kernel(const float* __restrict__ dl, __half* __restrict__ out, float scale, long long n_pix,
                       int HW, int W, int pad, int cp) {
  pdl_sync();
  const int H = HW / W, Wp = W + 2 * pad, Hp = H + 2 * pad;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n_pix;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = p / HW;
    const int hw = static_cast<int>(p - img * HW);
    float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = scale * dl[(img * 3 + c) * HW + hw];
    const int y = hw / W, x = hw - y * W;
    const long long o = (img * Hp + y + pad) * Wp + x + pad;
    if (cp == 8) reinterpret_cast<uint4*>(out)[o] = pack8(g);
    else reinterpret_cast<uint2*>(out)[o] = make_uint2(pack_h2(g[0], g[1]), pack_h2(g[2], 0.0f));
  }
}

// ---------------------------------------------------------------------------------------------
// optimizers (flat arena, float4 vectorised; 28 B / parameter for Adam)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
            float bc1, float bc2_sqrt, float gscale, const uint64_t* __restrict__ step_dev,
            unsigned int* __restrict__ nonfinite) {
  pdl_sync();
  // `nonfinite` (optional): a gradient entry that is inf / NaN (fp16 overflow somewhere in the backward,
  // or a poisoned input) leaves its parameter and moments untouched and raises the sticky flag, which the
  // host reads together with the loss — the step never trains on a non-finite number.
  bool bad = false;
  if (step_dev) {  // bias corrections from the device-resident step counter
    const double t = static_cast<double>(*step_dev);
    bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), t));
    bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), t)));
  }
  const long long n4 = n >> 2;
  const float step = lr / bc1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float gr = ga[q] * gscale;
      if (!isfinite(gr)) {
        bad = true;
        continue;
      }
      if (wd != 0.0f) gr = fmaf(wd, pa[q], gr);
      ma[q] = fmaf(b1, ma[q], (1.0f - b1) * gr);       // torch: m.lerp_(g, 1-b1)
      va[q] = fmaf(b2, va[q], (1.0f - b2) * gr * gr);  // v.mul_(b2).addcmul_(g, g, 1-b2)
      const float denom = sqrtf(va[q]) / bc2_sqrt + eps;
      pa[q] -= step * (ma[q] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float gr = g[i] * gscale;
    if (!isfinite(gr)) {
      bad = true;
    } else {
      if (wd != 0.0f) gr = fmaf(wd, p[i], gr);
      const float mm = fmaf(b1, m[i], (1.0f - b1) * gr);
      const float vv = fmaf(b2, v[i], (1.0f - b2) * gr * gr);
      m[i] = mm;
      v[i] = vv;
      p[i] -= step * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    }
  }
  if (bad && nonfinite) atomicOr(nonfinite, 1u);
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
           float lr, float momentum, float wd, int first_step, float gscale, unsigned int* __restrict__ nonfinite) {
  pdl_sync();
  bool bad = false;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g0 = g[i] * gscale;
    if (!isfinite(g0)) {  // see adam_kernel: skip the entry, raise the flag
      bad = true;
      continue;
    }
    const float gr = fmaf(wd, p[i], g0);
    const float b = first_step ? gr : fmaf(momentum, buf[i], gr);
    buf[i] = b;
    p[i] -= lr * b;
  }
  if (bad && nonfinite) atomicOr(nonfinite, 1u);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (one 128-bit block -> 4 uniforms)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) {  // (0, 1]
  return (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256)
fill_normal_kernel(float* __restrict__ out, long long n, uint64_t seed, uint64_t offset,
                   const uint64_t* __restrict__ ctr) {
  pdl_sync();
  if (ctr) offset += *ctr;
  const long long n4 = (n + 3) >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint64_t c = offset + static_cast<uint64_t>(i);
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0x6e6f726du, 0),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    float o[4];
    const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    o[0] = r0 * c0; o[1] = r0 * s0; o[2] = r1 * c1; o[3] = r1 * s1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * i + q < n) out[4 * i + q] = o[q];
  }
}

__global__ void __launch_bounds__(256)
fill_dropout_kernel(float* __restrict__ out, long long n, float p_drop, float keep_scale, uint64_t seed,
                    uint64_t offset, const uint64_t* __restrict__ ctr) {
  pdl_sync();
  if (ctr) offset += *ctr;
  const long long n4 = (n + 3) >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint64_t c = offset + static_cast<uint64_t>(i);
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0x64726f70u, 0),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * i + q < n) out[4 * i + q] = (u01(w[q]) > p_drop) ? keep_scale : 0.0f;
  }
}

__global__ void rng_advance_kernel(uint64_t* ctr, uint64_t inc) {
  pdl_sync(); *ctr += inc; }

inline int grid_for(long long n, int threads = 256, int max_blocks = 148 * 8) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}

// bulk-streamed kernels: one balanced wave (every CTA resident, equal row ranges, whole stages)
constexpr int BS_STAGES_1IN = 4, BS_STAGES_2IN = 3;
inline int bulk_chunking(int rows_per_group, int G, int C, int ctas_per_sm, int* rows_per_chunk) {
  const int rps = BS_STAGE_BYTES / (C * 2);
  const int stages_total = (rows_per_group + rps - 1) / rps;
  int chunks = (148 * ctas_per_sm) / G;
  if (chunks > stages_total) chunks = stages_total;
  if (chunks < 1) chunks = 1;
  const int stages_per_chunk = (stages_total + chunks - 1) / chunks;
  *rows_per_chunk = stages_per_chunk * rps;
  return (rows_per_group + *rows_per_chunk - 1) / *rows_per_chunk;
}
inline bool bulk_ok(int C) { return C >= 8 && C <= 2048 && (2048 % C) == 0; }
inline bool use_bulk() {
  static const bool off = getenv("MMDYN_BN_NO_BULK") != nullptr;
  return !off;
}

inline int chunking(int rows_per_group, int G, int C, int* rows_per_chunk) {
  // Many more CTAs than one wave (148 SMs x ~3 resident): ncu showed 1.33 waves with the old 4/SM
  // sizing, i.e. a third of the time spent in a one-third-full tail.  16 CTAs per SM keep the tail
  // under 7 % while every lane still streams >= 16 rows.
  const int row_lanes = RED_THREADS / (C >> 3);
  int chunks = (148 * 16 + G - 1) / G;
  const int max_chunks = (rows_per_group + row_lanes * 16 - 1) / (row_lanes * 16);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  *rows_per_chunk = (rows_per_group + chunks - 1) / chunks;
  return (rows_per_group + *rows_per_chunk - 1) / *rows_per_chunk;
}

}  // namespace
}  // namespace mmdyn

namespace mmdyn {
int elementwise_init() {
#define SET_SMEM(K, BYTES) \
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES))
  SET_SMEM(bn_stats_bulk_kernel<BS_STAGES_1IN>, BS_STAGES_1IN * BS_STAGE_BYTES);
  SET_SMEM(bn_swish_fwd_bulk_kernel<BS_STAGES_1IN>, BS_STAGES_1IN * BS_STAGE_BYTES);
  SET_SMEM(bn_swish_bwd_reduce_bulk_kernel<BS_STAGES_2IN>, 2 * BS_STAGES_2IN * BS_STAGE_BYTES);
  SET_SMEM(bn_bwd_apply_bulk_kernel<BS_STAGES_2IN>, 2 * BS_STAGES_2IN * BS_STAGE_BYTES);
#undef SET_SMEM
  return MMDYN_OK;
}
}  // namespace mmdyn

using namespace mmdyn;
#define ST(s) static_cast<cudaStream_t>(s)

static bool bn_c_ok(int C) { return C >= 8 && C % 8 == 0 && (C >> 3) <= RED_THREADS && (RED_THREADS % (C >> 3)) == 0; }

extern "C" int mmdyn_bn_stats(const void* x, float* sums, int G, int rows_per_group, int C, void* stream) {
  MMDYN_REQUIRE(x && sums && G > 0 && rows_per_group > 0 && bn_c_ok(C), "bn_stats: bad arguments (C=%d)", C);
  int rpc;
  if (use_bulk() && bulk_ok(C)) {
    const int chunks = bulk_chunking(rows_per_group, G, C, 3, &rpc);
    MMDYN_LAUNCH((bn_stats_bulk_kernel<BS_STAGES_1IN>), dim3(chunks, G), RED_THREADS, BS_STAGES_1IN * BS_STAGE_BYTES, ST(stream), 
        reinterpret_cast<const __half*>(x), sums, rows_per_group, C, rpc);
    LAUNCHED();
    return MMDYN_OK;
  }
  const int chunks = chunking(rows_per_group, G, C, &rpc);
  const size_t smem = static_cast<size_t>(RED_THREADS / (C >> 3)) * C * 2 * sizeof(float);
  MMDYN_LAUNCH((bn_stats_kernel), dim3(chunks, G), RED_THREADS, smem, ST(stream), 
      reinterpret_cast<const __half*>(x), sums, rows_per_group, C, rpc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bn_finalize(const float* sums, const float* gamma, const float* beta, float* ab,
                                 float* mean_invstd, float* running_mean, float* running_var, int G,
                                 int rows_per_group, int C, float eps, float momentum, int stat_repeat,
                                 long long* num_batches_tracked, void* stream) {
  MMDYN_REQUIRE(sums && gamma && beta && ab && mean_invstd && G > 0 && C > 0, "bn_finalize: bad arguments");
  MMDYN_LAUNCH((bn_finalize_kernel), (C + 127) / 128, 128, 0, ST(stream), sums, gamma, beta, ab, mean_invstd, running_mean,
                                                               running_var, G, rows_per_group, C, eps, momentum,
                                                               stat_repeat, num_batches_tracked);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bn_swish_fwd(const void* x, const float* ab, void* y, int G, int rows_per_group, int C,
                                  void* stream) {
  MMDYN_REQUIRE(x && y && G > 0 && rows_per_group > 0 && C % 8 == 0, "bn_swish_fwd: bad arguments");
  const long long n_vec = static_cast<long long>(G) * rows_per_group * (C >> 3);
  if (ab && use_bulk() && bulk_ok(C)) {
    int rpc;
    const int chunks = bulk_chunking(rows_per_group, G, C, 3, &rpc);
    BnFinalizeArgs fin = {};
    MMDYN_LAUNCH((bn_swish_fwd_bulk_kernel<BS_STAGES_1IN>), dim3(chunks, G), RED_THREADS, BS_STAGES_1IN * BS_STAGE_BYTES,
                                              ST(stream), reinterpret_cast<const __half*>(x), ab,
                                                            reinterpret_cast<__half*>(y), rows_per_group, C, rpc, fin);
    LAUNCHED();
    return MMDYN_OK;
  }
  if (ab && bn_c_ok(C)) {
    int rpc;
    const int chunks = chunking(rows_per_group, G, C, &rpc);
    MMDYN_LAUNCH((bn_swish_fwd_rows_kernel), dim3(chunks, G), RED_THREADS, 0, ST(stream), 
        reinterpret_cast<const __half*>(x), ab, reinterpret_cast<__half*>(y), rows_per_group, C, rpc);
    LAUNCHED();
    return MMDYN_OK;
  }
  MMDYN_LAUNCH((bn_swish_fwd_kernel), grid_for(n_vec / EW_UNROLL + 1), 256, 0, ST(stream), reinterpret_cast<const __half*>(x), ab,
                                                               reinterpret_cast<__half*>(y), n_vec,
                                                               rows_per_group, C);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bn_finalize_swish_fwd(const void* x, const float* sums, const float* gamma, const float* beta,
                                          float* ab, float* mean_invstd, float* running_mean, float* running_var,
                                          long long* num_batches_tracked, void* y, int G, int rows_per_group, int C,
                                          float eps, float momentum, int stat_repeat, void* stream) {
  MMDYN_REQUIRE(x && sums && gamma && beta && ab && mean_invstd && y && G > 0 && rows_per_group > 0 && C % 8 == 0,
                "bn_finalize_swish_fwd: bad arguments");
  if (use_bulk() && bulk_ok(C)) {
    int rpc;
    const int chunks = bulk_chunking(rows_per_group, G, C, 3, &rpc);
    BnFinalizeArgs fin = {sums, gamma, beta, ab, mean_invstd, running_mean, running_var, num_batches_tracked,
                          eps, momentum, stat_repeat};
    MMDYN_LAUNCH((bn_swish_fwd_bulk_kernel<BS_STAGES_1IN>), dim3(chunks, G), RED_THREADS, BS_STAGES_1IN * BS_STAGE_BYTES,
                                              ST(stream), reinterpret_cast<const __half*>(x), nullptr,
                                                            reinterpret_cast<__half*>(y), rows_per_group, C, rpc, fin);
    LAUNCHED();
    return MMDYN_OK;
  }
  const int rc = mmdyn_bn_finalize(sums, gamma, beta, ab, mean_invstd, running_mean, running_var, G, rows_per_group,
                                   C, eps, momentum, stat_repeat, num_batches_tracked, stream);
  if (rc != MMDYN_OK) return rc;
  return mmdyn_bn_swish_fwd(x, ab, y, G, rows_per_group, C, stream);
}

extern "C" int mmdyn_bn_swish_bwd_reduce(const void* x, const float* ab, const float* mean_invstd, void* dY,
                                         float* sums2, int G, int rows_per_group, int C, void* stream) {
  MMDYN_REQUIRE(x && dY && G > 0 && rows_per_group > 0 && C % 8 == 0, "bn_swish_bwd_reduce: bad arguments");
  if (!ab) {
    const long long n_vec = static_cast<long long>(G) * rows_per_group * (C >> 3);
    MMDYN_LAUNCH((swish_bwd_kernel), grid_for(n_vec), 256, 0, ST(stream), reinterpret_cast<const __half*>(x),
                                                              reinterpret_cast<__half*>(dY), n_vec);
    LAUNCHED();
    return MMDYN_OK;
  }
  MMDYN_REQUIRE(mean_invstd && sums2 && bn_c_ok(C), "bn_swish_bwd_reduce: bad arguments (C=%d)", C);
  int rpc;
  if (use_bulk() && bulk_ok(C)) {
    const int chunks = bulk_chunking(rows_per_group, G, C, 2, &rpc);
    MMDYN_LAUNCH((bn_swish_bwd_reduce_bulk_kernel<BS_STAGES_2IN>), dim3(chunks, G), RED_THREADS, 2 * BS_STAGES_2IN * BS_STAGE_BYTES, ST(stream), 
            reinterpret_cast<const __half*>(x), ab, mean_invstd, reinterpret_cast<const __half*>(dY), sums2,
            rows_per_group, C, rpc);
    LAUNCHED();
    return MMDYN_OK;
  }
  const int chunks = chunking(rows_per_group, G, C, &rpc);
  const size_t smem = static_cast<size_t>(RED_THREADS / (C >> 3)) * C * 2 * sizeof(float);
  MMDYN_LAUNCH((bn_swish_bwd_reduce_kernel), dim3(chunks, G), RED_THREADS, smem, ST(stream), 
      reinterpret_cast<const __half*>(x), ab, mean_invstd, reinterpret_cast<const __half*>(dY), sums2,
      rows_per_group, C, rpc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bn_bwd_apply_padded(const void* x, const float* ab, const float* mean_invstd, const float* sums2,
                                         const void* dU, void* dX_padded, int row_w_log2, float* dgamma, float* dbeta,
                                         int G, int rows_per_group, int C, float grad_unscale, void* stream) {
  MMDYN_REQUIRE(x && ab && mean_invstd && sums2 && dU && dX_padded && G > 0 && C % 8 == 0 && row_w_log2 >= 1 &&
                    row_w_log2 <= 12 && rows_per_group % (1 << row_w_log2) == 0,
                "bn_bwd_apply_padded: bad arguments (rows_per_group must be whole image rows of 2^row_w_log2 pixels)");
  MMDYN_REQUIRE(bulk_ok(C), "bn_bwd_apply_padded: C=%d unsupported", C);
  int rpc;
  const int chunks = bulk_chunking(rows_per_group, G, C, 2, &rpc);
  BnBwdFinalArgs fin = {ab, mean_invstd, sums2, dgamma, dbeta, 1.0f / static_cast<float>(rows_per_group), grad_unscale};
  MMDYN_LAUNCH((bn_bwd_apply_bulk_kernel<BS_STAGES_2IN>), dim3(chunks, G), RED_THREADS, 2 * BS_STAGES_2IN * BS_STAGE_BYTES, ST(stream),
          reinterpret_cast<const __half*>(x), nullptr, const_cast<__half*>(reinterpret_cast<const __half*>(dU)), rows_per_group, C, rpc,
          fin, reinterpret_cast<__half*>(dX_padded), row_w_log2);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bn_bwd_apply(const void* x, const float* ab, const float* mean_invstd, const float* sums2,
                                  void* dU, float* dgamma, float* dbeta, float* coef_scratch, int G,
                                  int rows_per_group, int C, float grad_unscale, void* stream) {
  MMDYN_REQUIRE(x && ab && mean_invstd && sums2 && dU && G > 0 && C % 8 == 0, "bn_bwd_apply: bad arguments");
  MMDYN_REQUIRE(coef_scratch, "bn_bwd_apply: coef_scratch ([G][C][4] floats) is required");
  MMDYN_REQUIRE(bn_c_ok(C), "bn_bwd_apply: C=%d unsupported", C);
  int rpc;
  if (use_bulk() && bulk_ok(C)) {  // coefficients and parameter gradients are derived inside the streaming kernel
    const int chunks = bulk_chunking(rows_per_group, G, C, 2, &rpc);
    BnBwdFinalArgs fin = {ab, mean_invstd, sums2, dgamma, dbeta, 1.0f / static_cast<float>(rows_per_group),
                          grad_unscale};
    MMDYN_LAUNCH((bn_bwd_apply_bulk_kernel<BS_STAGES_2IN>), dim3(chunks, G), RED_THREADS, 2 * BS_STAGES_2IN * BS_STAGE_BYTES, ST(stream), 
            reinterpret_cast<const __half*>(x), nullptr, reinterpret_cast<__half*>(dU), rows_per_group, C, rpc, fin,
            static_cast<__half*>(nullptr), 0);
    LAUNCHED();
    return MMDYN_OK;
  }
  MMDYN_LAUNCH((bn_bwd_coef_kernel), (G * C + 127) / 128, 128, 0, ST(stream), ab, mean_invstd, sums2, coef_scratch, G * C,
                                                                   1.0f / static_cast<float>(rows_per_group));
  LAUNCHED();
  const int chunks = chunking(rows_per_group, G, C, &rpc);
  MMDYN_LAUNCH((bn_bwd_apply_kernel), dim3(chunks, G), RED_THREADS, 0, ST(stream), 
      reinterpret_cast<const __half*>(x), coef_scratch, reinterpret_cast<__half*>(dU), rows_per_group, C, rpc);
  LAUNCHED();
  if (dgamma && dbeta) {
    MMDYN_LAUNCH((bn_param_grad_kernel), (C + 127) / 128, 128, 0, ST(stream), sums2, dgamma, dbeta, G, C, grad_unscale);
    LAUNCHED();
  }
  return MMDYN_OK;
}

extern "C" int mmdyn_swish_dropout_fwd(const float* raw, const float* const* masks, void* h, int n_masks,
                                       int B, int C, void* stream) {
  MMDYN_REQUIRE(raw && h && n_masks >= 1 && n_masks <= 8, "swish_dropout_fwd: bad arguments");
  MaskPtrs mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = (masks && i < n_masks) ? masks[i] : nullptr;
  const long long n = static_cast<long long>(B) * C;
  MMDYN_LAUNCH((swish_dropout_fwd_kernel), grid_for(n), 256, 0, ST(stream), raw, mp, reinterpret_cast<__half*>(h), n_masks, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_swish_dropout_bwd(const float* raw, const float* const* masks, const float* dH, void* dRaw,
                                       int n_masks, int B, int C, void* stream) {
  MMDYN_REQUIRE(raw && dH && dRaw && n_masks >= 1 && n_masks <= 8, "swish_dropout_bwd: bad arguments");
  MaskPtrs mp;
  for (int i = 0; i < 8; ++i) mp.p[i] = (masks && i < n_masks) ? masks[i] : nullptr;
  const long long n = static_cast<long long>(B) * C;
  MMDYN_LAUNCH((swish_dropout_bwd_kernel), grid_for(n), 256, 0, ST(stream), raw, mp, dH, reinterpret_cast<__half*>(dRaw),
                                                                n_masks, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_poe_fwd(const float* const* mu_e, const float* const* lv_e, int n_experts, int use_prior,
                             int ld, const float* eps, float* mu, float* lv, float* z, void* zh, void* zh2,
                             float* kl_sum, int B, int D, void* stream) {
  MMDYN_REQUIRE(n_experts >= 0 && n_experts <= MAX_EXPERTS && (n_experts > 0 || use_prior), "poe_fwd: n_experts=%d", n_experts);
  MMDYN_REQUIRE(eps && mu && lv && z && kl_sum && B > 0 && D > 0, "poe_fwd: null pointer");
  ExpertPtrs ex = {};
  for (int e = 0; e < n_experts; ++e) {
    MMDYN_REQUIRE(mu_e[e] && lv_e[e], "poe_fwd: null expert %d", e);
    ex.mu[e] = mu_e[e];
    ex.lv[e] = lv_e[e];
  }
  const long long n = static_cast<long long>(B) * D;
  MMDYN_LAUNCH((poe_fwd_kernel), grid_for(n), 256, 0, ST(stream), ex, n_experts, use_prior, ld, eps, mu, lv, z,
                                                      reinterpret_cast<__half*>(zh), reinterpret_cast<__half*>(zh2),
                                                      kl_sum, B, D);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_poe_bwd(const float* const* mu_e, const float* const* lv_e, int n_experts, int use_prior,
                             int ld, const float* eps, const float* const* dz, const float* dmu_in,
                             const float* dlv_in, float kl_coef, float* const* dmu_e,
                             float* const* dlv_e, int ld_out, int accumulate, int B, int D, void* stream) {
  MMDYN_REQUIRE(n_experts >= 0 && n_experts <= MAX_EXPERTS, "poe_bwd: n_experts=%d", n_experts);
  if (n_experts == 0) return MMDYN_OK;
  MMDYN_REQUIRE(eps && B > 0 && D > 0, "poe_bwd: null pointer");
  ExpertPtrs ex = {};
  for (int e = 0; e < n_experts; ++e) {
    MMDYN_REQUIRE(mu_e[e] && lv_e[e] && dmu_e[e] && dlv_e[e], "poe_bwd: null expert %d", e);
    ex.mu[e] = mu_e[e];
    ex.lv[e] = lv_e[e];
    ex.dmu[e] = dmu_e[e];
    ex.dlv[e] = dlv_e[e];
  }
  for (int q = 0; q < 3; ++q) ex.dz[q] = dz ? dz[q] : nullptr;
  const long long n = static_cast<long long>(B) * D;
  MMDYN_LAUNCH((poe_bwd_kernel), grid_for(n), 256, 0, ST(stream), ex, n_experts, use_prior, ld, eps, dmu_in, dlv_in, kl_coef,
                                                      ld_out, accumulate, B, D);
  LAUNCHED();
  return MMDYN_OK;
}

static int fill_poe_passes(const mmdyn_poe_pass* passes, int n_passes, bool bwd, PoePassesDev* out, const char* what) {
  MMDYN_REQUIRE(passes && n_passes >= 1 && n_passes <= MMDYN_MAX_POE_PASSES, "%s: n_passes=%d (1..%d)", what, n_passes,
                MMDYN_MAX_POE_PASSES);
  for (int k = 0; k < n_passes; ++k) {
    const mmdyn_poe_pass& a = passes[k];
    PoePassDev& q = out->p[k];
    q = {};
    MMDYN_REQUIRE(a.n_experts >= 0 && a.n_experts <= MAX_EXPERTS && a.eps, "%s: pass %d: n_experts=%d / eps", what, k,
                  a.n_experts);
    q.n_experts = a.n_experts;
    q.eps = a.eps;
    for (int e = 0; e < a.n_experts; ++e) {
      MMDYN_REQUIRE(a.mu_e[e] && a.lv_e[e], "%s: pass %d: null expert %d", what, k, e);
      q.ex.mu[e] = a.mu_e[e];
      q.ex.lv[e] = a.lv_e[e];
      if (bwd) {
        MMDYN_REQUIRE(a.dmu_e[e] && a.dlv_e[e], "%s: pass %d: null expert gradient %d", what, k, e);
        q.ex.dmu[e] = a.dmu_e[e];
        q.ex.dlv[e] = a.dlv_e[e];
      }
    }
    if (bwd) {
      for (int j = 0; j < 3; ++j) q.ex.dz[j] = a.dz[j];
      q.dmu_in = a.dmu_in;
      q.dlv_in = a.dlv_in;
    } else {
      MMDYN_REQUIRE(a.mu && a.lv && a.z && a.kl_sum, "%s: pass %d: null output", what, k);
      q.mu_o = a.mu;
      q.lv_o = a.lv;
      q.z_o = a.z;
      q.zh_o = reinterpret_cast<__half*>(a.zh);
      q.zh2_o = reinterpret_cast<__half*>(a.zh2);
      q.kl_sum = a.kl_sum;
    }
  }
  return MMDYN_OK;
}

extern "C" int mmdyn_poe_fwd_multi(const mmdyn_poe_pass* passes, int n_passes, int use_prior, int ld, int B, int D,
                                   void* stream) {
  MMDYN_REQUIRE(B > 0 && D > 0, "poe_fwd_multi: empty problem");
  PoePassesDev ps = {};
  const int rc = fill_poe_passes(passes, n_passes, false, &ps, "poe_fwd_multi");
  if (rc != MMDYN_OK) return rc;
  for (int k = 0; k < n_passes; ++k)
    MMDYN_REQUIRE(passes[k].n_experts > 0 || use_prior, "poe_fwd_multi: pass %d has no expert and no prior", k);
  const long long n = static_cast<long long>(B) * D;
  int gx = static_cast<int>((n + 255) / 256);
  const int cap = (148 * 8 + n_passes - 1) / n_passes;
  if (gx > cap) gx = cap;
  MMDYN_LAUNCH((poe_fwd_multi_kernel), dim3(gx, n_passes), 256, 0, ST(stream), ps, use_prior, ld, B, D);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_poe_bwd_multi(const mmdyn_poe_pass* passes, int n_passes, int use_prior, int ld, float kl_coef,
                                   int ld_out, int accumulate, int B, int D, void* stream) {
  MMDYN_REQUIRE(B > 0 && D > 0, "poe_bwd_multi: empty problem");
  PoePassesDev ps = {};
  const int rc = fill_poe_passes(passes, n_passes, true, &ps, "poe_bwd_multi");
  if (rc != MMDYN_OK) return rc;
  const long long n = static_cast<long long>(B) * D;
  int gx = static_cast<int>((n + 255) / 256);
  const int cap = (148 * 8 + n_passes - 1) / n_passes;
  if (gx > cap) gx = cap;
  MMDYN_LAUNCH((poe_bwd_multi_kernel), dim3(gx, n_passes), 256, 0, ST(stream), ps, use_prior, ld, kl_coef, ld_out,
               accumulate, B, D);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bce_logits(const float* logits, const float* target, const float* mask, float* loss_sum,
                                void* dlogits_nhwc4, float gscale, int n, int H, int W, int pad, void* stream) {
  MMDYN_REQUIRE(logits && target && loss_sum && n > 0 && H > 0 && W > 0 && W % 4 == 0 && pad >= 0,
                "bce_logits: bad arguments (n=%d H=%d W=%d pad=%d; W must be a multiple of 4)", n, H, W, pad);
  const int HW = H * W;
  const long long n_pix = static_cast<long long>(n) * HW;
  MMDYN_LAUNCH((bce_logits_kernel), grid_for(n_pix >> 2), 256, 0, ST(stream), logits, target, mask, loss_sum,
                                                             reinterpret_cast<__half*>(dlogits_nhwc4), gscale,
                                                             n_pix, HW, W, pad);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_bce_logits_flat(const float* logits, const float* target, const float* mask, float* loss_sum,
                                     float* per_sample_sum, float* dlogits, float gscale, int n, int per_sample,
                                     void* stream) {
  MMDYN_REQUIRE(logits && target && loss_sum && n > 0 && per_sample > 0, "bce_logits_flat: bad arguments");
  int chunks = (148 * 8 + n - 1) / n;
  const int max_chunks = (per_sample + 1023) / 1024;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  MMDYN_LAUNCH((bce_flat_kernel), dim3(chunks, n), 256, 0, ST(stream), logits, target, mask, loss_sum, per_sample_sum, dlogits,
                                                           gscale, per_sample);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_mse_rows(const float* recon, const float* target, float* row_sum, float mult, int n, int d,
                              void* stream) {
  MMDYN_REQUIRE(recon && target && row_sum && n > 0 && d > 0, "mse_rows: bad arguments");
  MMDYN_LAUNCH((mse_rows_kernel), (n + 127) / 128, 128, 0, ST(stream), recon, target, row_sum, mult, n, d);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_mse(const float* recon, const float* target, float* loss_sum, float* drecon, float mult,
                         float gscale, int n, void* stream) {
  MMDYN_REQUIRE(recon && target && loss_sum && n > 0, "mse: bad arguments");
  MMDYN_LAUNCH((mse_kernel), grid_for(n, 256, 64), 256, 0, ST(stream), recon, target, loss_sum, drecon, mult, gscale, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_colsum_f32(const float* x, float* out, int M, int N, int ld, float scale, void* stream) {
  MMDYN_REQUIRE(x && out && M > 0 && N > 0, "colsum: bad arguments");
  int row_ctas = (148 * 2) / ((N + 255) / 256);
  if (row_ctas < 1) row_ctas = 1;
  int rpc = (M + row_ctas - 1) / row_ctas;
  if (rpc < 16) rpc = 16;
  const int gy = (M + rpc - 1) / rpc;
  MMDYN_LAUNCH((colsum_kernel), dim3((N + 255) / 256, gy), 256, 0, ST(stream), x, out, M, N, ld, scale, rpc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_colsum_f16(const void* x, float* out, int M, int N, int ld, float scale, void* stream) {
  MMDYN_REQUIRE(x && out && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "colsum_f16: bad arguments");
  const int gx = (N + 255) / 256;
  int row_ctas = (148 * 2) / gx;
  if (row_ctas < 1) row_ctas = 1;
  int rpc = (M + row_ctas - 1) / row_ctas;
  if (rpc < 32) rpc = 32;
  const int gy = (M + rpc - 1) / rpc;
  MMDYN_LAUNCH((colsum_f16_kernel), dim3(gx, gy), 256, 0, ST(stream), reinterpret_cast<const __half*>(x), out, M, N, ld, scale,
                                                          rpc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_pack_f16(const float* src, const int32_t* idx, void* dst, long long n, void* stream) {
  MMDYN_REQUIRE(src && idx && dst && n > 0, "pack_f16: bad arguments");
  MMDYN_REQUIRE((reinterpret_cast<uintptr_t>(idx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "pack_f16: idx and dst must be 16-byte aligned");
  MMDYN_LAUNCH((pack_f16_kernel), grid_for((n >> 3) + 1), 256, 0, ST(stream), src, idx, reinterpret_cast<__half*>(dst), n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_split_f16(const float* x, const float* mask_y, float* x_out, void* out, int M, int N, int mode,
                               int relu, float* colsum0, float* colsum1, int n_split, float colsum_scale, void* stream) {
  MMDYN_REQUIRE(x && out && M > 0 && N > 0 && (mode == 0 || mode == 1), "split_f16: bad arguments");
  MMDYN_REQUIRE(!colsum1 || (colsum0 && n_split > 0 && n_split < N), "split_f16: colsum1 needs colsum0 and 0 < n_split < N");
  const int gx = (N + 255) / 256;
  int row_ctas = (148 * 4) / gx;
  if (row_ctas < 1) row_ctas = 1;
  int rpc = (M + row_ctas - 1) / row_ctas;
  if (rpc < 8) rpc = 8;
  const int gy = (M + rpc - 1) / rpc;
  MMDYN_LAUNCH((split_f16_kernel), dim3(gx, gy), 256, 0, ST(stream), x, mask_y, x_out, reinterpret_cast<__half*>(out), M, N, mode, relu,
                                                         colsum0, colsum1, colsum0 && !colsum1 ? N : n_split, colsum_scale,
                                                         rpc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_gather_f32(const float* src, const int32_t* idx, float* dst, long long n, void* stream) {
  MMDYN_REQUIRE(src && idx && dst && n > 0, "gather_f32: bad arguments");
  MMDYN_LAUNCH((gather_f32_kernel), grid_for(n), 256, 0, ST(stream), src, idx, dst, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_unpack_add_f32(const float* src, const int32_t* idx, float* dst, long long n, void* stream) {
  MMDYN_REQUIRE(src && idx && dst && n > 0, "unpack_add: bad arguments");
  MMDYN_LAUNCH((unpack_add_kernel), grid_for(n), 256, 0, ST(stream), src, idx, dst, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_gather_add_f32(const float* src, const int32_t* inv, float* dst, long long n, void* stream) {
  MMDYN_REQUIRE(src && inv && dst && n > 0, "gather_add: bad arguments");
  MMDYN_LAUNCH((gather_add_kernel), grid_for((n >> 2) + 1), 256, 0, ST(stream), src, inv, dst, n);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_f32_to_f16(const float* src, void* dst, long long n, float scale, void* stream) {
  MMDYN_REQUIRE(src && dst && n > 0, "f32_to_f16: bad arguments");
  MMDYN_LAUNCH((f32_to_f16_kernel), grid_for(n), 256, 0, ST(stream), src, reinterpret_cast<__half*>(dst), n, scale);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_scale_f32(float* x, long long n, float s, void* stream) {
  MMDYN_REQUIRE(x && n > 0, "scale_f32: bad arguments");
  MMDYN_LAUNCH((scale_f32_kernel), grid_for(n), 256, 0, ST(stream), x, n, s);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_logit_grad_pack(const float* dlogits_nchw, void* out_nhwc, float scale, int n, int H, int W,
                                     int pad, int cp, void* stream) {
  MMDYN_REQUIRE(dlogits_nchw && out_nhwc && n > 0 && H > 0 && W > 0 && pad >= 0 && (cp == 4 || cp == 8),
                "logit_grad_pack: bad arguments (cp=%d must be 4 or 8)", cp);
  const int HW = H * W;
  const long long n_pix = static_cast<long long>(n) * HW;
  MMDYN_LAUNCH((logit_grad_pack_kernel), grid_for(n_pix), 256, 0, ST(stream), dlogits_nchw, reinterpret_cast<__half*>(out_nhwc),
                                                                  scale, n_pix, HW, W, pad, cp);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int step_count, float gscale,
                               void* stream) {
  MMDYN_REQUIRE(p && g && m && v && n > 0 && step_count >= 1, "adam_flat: bad arguments");
  MMDYN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                "adam_flat: arenas must be 16-byte aligned");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step_count);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step_count);
  MMDYN_LAUNCH((adam_kernel), grid_for(n >> 2), 256, 0, ST(stream), p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                        static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
                                                        gscale, nullptr, nullptr);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_adam_flat_guarded(float* p, const float* g, float* m, float* v, long long n, float lr,
                                       float beta1, float beta2, float eps, float weight_decay,
                                       const uint64_t* step_dev, float gscale, unsigned int* nonfinite_flag,
                                       void* stream) {
  MMDYN_REQUIRE(p && g && m && v && n > 0 && step_dev, "adam_flat_devstep: bad arguments");
  MMDYN_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                "adam_flat_devstep: arenas must be 16-byte aligned");
  MMDYN_LAUNCH((adam_kernel), grid_for(n >> 2), 256, 0, ST(stream), p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, 1.0f,
                                                        1.0f, gscale, step_dev, nonfinite_flag);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_adam_flat_devstep(float* p, const float* g, float* m, float* v, long long n, float lr,
                                       float beta1, float beta2, float eps, float weight_decay,
                                       const uint64_t* step_dev, float gscale, void* stream) {
  return mmdyn_adam_flat_guarded(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, gscale, nullptr, stream);
}

extern "C" int mmdyn_sgd_flat(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                              float weight_decay, int first_step, float gscale, void* stream) {
  return mmdyn_sgd_flat_guarded(p, g, buf, n, lr, momentum, weight_decay, first_step, gscale, nullptr, stream);
}

extern "C" int mmdyn_sgd_flat_guarded(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                                      float weight_decay, int first_step, float gscale, unsigned int* nonfinite_flag,
                                      void* stream) {
  MMDYN_REQUIRE(p && g && buf && n > 0, "sgd_flat: bad arguments");
  MMDYN_LAUNCH((sgd_kernel), grid_for(n), 256, 0, ST(stream), p, g, buf, n, lr, momentum, weight_decay, first_step, gscale,
                                                  nonfinite_flag);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_rng_advance(uint64_t* ctr_dev, uint64_t inc, void* stream) {
  MMDYN_REQUIRE(ctr_dev, "rng_advance: null counter");
  MMDYN_LAUNCH((rng_advance_kernel), 1, 1, 0, ST(stream), ctr_dev, inc);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_fill_normal(float* out, long long n, uint64_t seed, uint64_t offset, const uint64_t* ctr_dev,
                                 void* stream) {
  MMDYN_REQUIRE(out && n > 0, "fill_normal: bad arguments");
  MMDYN_LAUNCH((fill_normal_kernel), grid_for((n + 3) >> 2), 256, 0, ST(stream), out, n, seed, offset, ctr_dev);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_fill_dropout_mask(float* out, long long n, float p_drop, uint64_t seed, uint64_t offset,
                                       const uint64_t* ctr_dev, void* stream) {
  MMDYN_REQUIRE(out && n > 0 && p_drop >= 0.0f && p_drop < 1.0f, "fill_dropout_mask: bad arguments");
  MMDYN_LAUNCH((fill_dropout_kernel), grid_for((n + 3) >> 2), 256, 0, ST(stream), out, n, p_drop, 1.0f / (1.0f - p_drop), seed,
                                                                      offset, ctr_dev);
  LAUNCHED();
  return MMDYN_OK;
}
