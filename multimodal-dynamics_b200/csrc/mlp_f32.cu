// fp32 linear layers of the pose MLP expert (Linear(7,512)+ReLU, Linear(512,512), heads, and the
// 256->512->512->7 decoder): mmdyn/pytorch/models/vae.py:14-19, 118-123, 219-222, 282-283.
//
// The pose branch is ~1 MFLOP/sample and feeds the fp32 ProductOfExperts directly, so it stays in
// fp32 on the CUDA cores: a 64x64x16 shared-memory tiled SGEMM (4x4 register micro-tile, 256
// threads), with transposition flags so the same kernel serves y = xW^T, dx = dy W and
// dW = dy^T x (split along the reduction with atomics).
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

// C[m][n] (=|+=) scale * sum_k A(m,k) * B(k,n)   (+ bias[n], optional ReLU)
//   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k];   B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             const float* __restrict__ bias, int M, int N, int K, int lda, int ldb, int ldc, int act,
             int atomic_add, float scale, int k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = t + i * 256;
      int m, k;
      if (!TA) { m = idx >> 4; k = idx & 15; } else { m = idx & 63; k = idx >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.0f;
      if (gm < M && gk < k_end) v = TA ? A[static_cast<long long>(gk) * lda + gm] : A[static_cast<long long>(gm) * lda + gk];
      As[k][m] = v;
      int n, kb;
      if (TB) { n = idx >> 4; kb = idx & 15; } else { n = idx & 63; kb = idx >> 6; }
      const int gn = n0 + n, gkb = k0 + kb;
      float w = 0.0f;
      if (gn < N && gkb < k_end) w = TB ? B[static_cast<long long>(gn) * ldb + gkb] : B[static_cast<long long>(gkb) * ldb + gn];
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = scale * acc[i][j];
      float* o = C + static_cast<long long>(gm) * ldc + gn;
      if (atomic_add) {
        atomicAdd(o, v);
      } else {
        if (bias) v += bias[gn];
        if (act == 1) v = fmaxf(v, 0.0f);
        *o = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
act_grad_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ out, int M,
                int N, int ldy, int act) {
  const long long n = static_cast<long long>(M) * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / N), c = static_cast<int>(i - static_cast<long long>(r) * N);
    const float g = dy[static_cast<long long>(r) * ldy + c];
    out[i] = (act == 1 && y[static_cast<long long>(r) * ldy + c] <= 0.0f) ? 0.0f : g;
  }
}

}  // namespace
}  // namespace mmdyn

using namespace mmdyn;
#define ST(s) static_cast<cudaStream_t>(s)
#define LAUNCHED()                                            \
  do {                                                        \
    g_launch_count.fetch_add(1, std::memory_order_relaxed);   \
    MMDYN_CHECK_CUDA(cudaGetLastError());                     \
  } while (0)

extern "C" int mmdyn_linear_f32_fwd(const float* x, const float* W, const float* b, float* y, int M, int N,
                                    int K, int ldx, int ldy, int act, void* stream) {
  MMDYN_REQUIRE(x && W && y && M > 0 && N > 0 && K > 0, "linear_f32_fwd: bad arguments");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  sgemm_kernel<false, true><<<grid, 256, 0, ST(stream)>>>(x, W, y, b, M, N, K, ldx, K, ldy, act, 0, 1.0f, K);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_linear_f32_bwd(const float* x, const float* W, const float* y, const float* dy,
                                    float* dy_act, float* dx, float* dW, float* db, int M, int N, int K,
                                    int ldx, int ldy, int lddx, int act, int dx_accumulate, float scale,
                                    void* stream) {
  MMDYN_REQUIRE(x && W && y && dy && dy_act && M > 0 && N > 0 && K > 0, "linear_f32_bwd: bad arguments");
  {
    const long long n = static_cast<long long>(M) * N;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    act_grad_kernel<<<static_cast<int>(blocks), 256, 0, ST(stream)>>>(y, dy, dy_act, M, N, ldy, act);
    LAUNCHED();
  }
  if (dx) {  // dx[M][K] = dy_act[M][N] * W[N][K]
    dim3 grid((K + BN - 1) / BN, (M + BM - 1) / BM, 1);
    sgemm_kernel<false, false><<<grid, 256, 0, ST(stream)>>>(dy_act, W, dx, nullptr, M, K, N, N, K, lddx, 0,
                                                             dx_accumulate ? 1 : 0, 1.0f, N);
    LAUNCHED();
  }
  if (dW) {  // dW[N][K] += scale * dy_act^T[N][M] * x[M][K]
    int splits = (148 * 2) / (((K + BN - 1) / BN) * ((N + BM - 1) / BM));
    if (splits < 1) splits = 1;
    int kps = (M + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (M + kps - 1) / kps;
    dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, splits);
    sgemm_kernel<true, false><<<grid, 256, 0, ST(stream)>>>(dy_act, x, dW, nullptr, N, K, M, N, ldx, K, 0, 1,
                                                            scale, kps);
    LAUNCHED();
  }
  if (db) {
    const int rc = mmdyn_colsum_f32(dy_act, db, M, N, N, scale, stream);
    if (rc != MMDYN_OK) return rc;
  }
  return MMDYN_OK;
}
