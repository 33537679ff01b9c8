// fp32 linear layers of the pose MLP expert (Linear(7,512)+ReLU, Linear(512,512), heads, and the
// 256->512->512->7 decoder): mmdyn/pytorch/models/vae.py:14-19, 118-123, 219-222, 282-283.
//
// The pose branch is ~1 MFLOP/sample and feeds the fp32 ProductOfExperts directly, so it stays in
// fp32 on the CUDA cores: a {128,64}x64x16 shared-memory tiled SGEMM (8x4 / 4x4 register micro-tile,
// 256 threads, double-buffered k-tiles, 16-byte global loads), with transposition flags so the same
// kernel serves y = xW^T, dx = dy W and dW = dy^T x (split along the reduction with atomics).
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;
namespace {

constexpr int BN = 64, BK = 16;

// one k-tile of an operand -> registers.  T = false: k contiguous in memory (4 consecutive k of one
// row per float4); T = true: the row (m or n) index is contiguous (4 consecutive rows at one k).
template <bool T, int RB, int NV>
__device__ __forceinline__ void load_tile(const float* __restrict__ X, bool vec, int r0, int R, int ld, int k0,
                                          int k_end, float4 (&regs)[NV]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = t + i * 256;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!T) {
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int gr = r0 + r, gk = k0 + kq;
      if (gr < R) {
        const float* p = X + static_cast<long long>(gr) * ld + gk;
        if (vec && gk + 3 < k_end) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (gk < k_end) v.x = __ldg(p);
          if (gk + 1 < k_end) v.y = __ldg(p + 1);
          if (gk + 2 < k_end) v.z = __ldg(p + 2);
          if (gk + 3 < k_end) v.w = __ldg(p + 3);
        }
      }
    } else {
      const int rq = (idx % (RB / 4)) * 4, k = idx / (RB / 4);
      const int gr = r0 + rq, gk = k0 + k;
      if (gk < k_end) {
        const float* p = X + static_cast<long long>(gk) * ld + gr;
        if (vec && gr + 3 < R) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (gr < R) v.x = __ldg(p);
          if (gr + 1 < R) v.y = __ldg(p + 1);
          if (gr + 2 < R) v.z = __ldg(p + 2);
          if (gr + 3 < R) v.w = __ldg(p + 3);
        }
      }
    }
    regs[i] = v;
  }
}

// C[m][n] (=|+=) scale * sum_k A(m,k) * B(k,n)   (+ bias[n], optional ReLU)
//   A(m,k) = TA ? A[k*lda + m] : A[m*lda + k];   B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
// BM x 64 x 16 tiles (BM = 128: 8x4 register micro-tile, BM = 64: 4x4), 256 threads, operands staged
// k-major in double-buffered shared memory: the next k-tile's global loads (16-byte where the
// operand layout allows) are issued before the FMAs of the current one, one barrier per k-tile.
template <int BM, bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             const float* __restrict__ bias, int M, int N, int K, int lda, int ldb, int ldc, int act,
             int atomic_add, float scale, int k_per_split) {
  pdl_sync();
  constexpr int TM = BM / 16;          // rows per thread
  constexpr int A_V4 = BM * BK / 4 / 256;  // float4 loads of A per thread and k-tile
  constexpr int B_V4 = BN * BK / 4 / 256;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const bool a_vec = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const bool b_vec = (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float4 ra[A_V4], rb[B_V4];
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      const int idx = t + i * 256;
      if (!TA) {
        const int r = idx >> 2, kq = (idx & 3) * 4;
        As[buf][kq][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      } else {
        const int rq = (idx % (BM / 4)) * 4, k = idx / (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][k][rq]) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < B_V4; ++i) {
      const int idx = t + i * 256;
      if (TB) {
        const int r = idx >> 2, kq = (idx & 3) * 4;
        Bs[buf][kq][r] = rb[i].x; Bs[buf][kq + 1][r] = rb[i].y; Bs[buf][kq + 2][r] = rb[i].z; Bs[buf][kq + 3][r] = rb[i].w;
      } else {
        const int rq = (idx % (BN / 4)) * 4, k = idx / (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][rq]) = rb[i];
      }
    }
  };

  load_tile<TA, BM, A_V4>(A, a_vec, m0, M, lda, k_begin, k_end, ra);
  load_tile<!TB, BN, B_V4>(B, b_vec, n0, N, ldb, k_begin, k_end, rb);
  stage(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      load_tile<TA, BM, A_V4>(A, a_vec, m0, M, lda, k0 + BK, k_end, ra);
      load_tile<!TB, BN, B_V4>(B, b_vec, n0, N, ldb, k0 + BK, k_end, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM];
#pragma unroll
      for (int h = 0; h < TM / 4; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][h * 64 + ty * 4]);
        av[4 * h] = a.x; av[4 * h + 1] = a.y; av[4 * h + 2] = a.z; av[4 * h + 3] = a.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stage(buf ^ 1);  // the other buffer was last read before the previous barrier
      __syncthreads();
      buf ^= 1;
    }
  }
  const bool c_vec = !atomic_add && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (gm >= M) continue;
    const int gn = n0 + tx * 4;
    float* o = C + static_cast<long long>(gm) * ldc + gn;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = scale * acc[i][j];
      if (!atomic_add) {
        if (bias && gn + j < N) v[j] += __ldg(bias + gn + j);
        if (act == 1) v[j] = fmaxf(v[j], 0.0f);
      }
    }
    if (c_vec && gn + 3 < N) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gn + j >= N) continue;
        if (atomic_add) atomicAdd(o + j, v[j]);
        else o[j] = v[j];
      }
    }
  }
}

// 128-row tiles when they still fill the machine, 64-row tiles otherwise
template <bool TA, bool TB>
void launch_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                  int ldc, int act, int atomic_add, float scale, int k_per_split, int splits, cudaStream_t st) {
  const int nb = (N + BN - 1) / BN;
  if (static_cast<long long>((M + 127) / 128) * nb * splits >= 148) {
    MMDYN_LAUNCH((sgemm_kernel<128, TA, TB>), dim3(nb, (M + 127) / 128, splits), 256, 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc,
                                                                                   act, atomic_add, scale, k_per_split);
  } else {
    MMDYN_LAUNCH((sgemm_kernel<64, TA, TB>), dim3(nb, (M + 63) / 64, splits), 256, 0, st, A, B, C, bias, M, N, K, lda, ldb, ldc,
                                                                                 act, atomic_add, scale, k_per_split);
  }
}

__global__ void __launch_bounds__(256)
act_grad_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ out, int M,
                int N, int ldy, int act) {
  pdl_sync();
  const long long n = static_cast<long long>(M) * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / N), c = static_cast<int>(i - static_cast<long long>(r) * N);
    const float g = dy[static_cast<long long>(r) * ldy + c];
    out[i] = (act == 1 && y[static_cast<long long>(r) * ldy + c] <= 0.0f) ? 0.0f : g;
  }
}

// ReLU in fp32 (Regressor.out_net behind the tensor-core Linear(512, 256), models.py:56-62)
__global__ void __launch_bounds__(256) relu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_sync();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    v.x = fmaxf(v.x, 0.0f), v.y = fmaxf(v.y, 0.0f), v.z = fmaxf(v.z, 0.0f), v.w = fmaxf(v.w, 0.0f);
    reinterpret_cast<float4*>(y)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) y[n4 * 4 + threadIdx.x] = fmaxf(x[n4 * 4 + threadIdx.x], 0.0f);
}

// ---------------------------------------------------------------------------------------------
// CVAE conditioning (vae.py:231-237, 286-291): torch.cat((x, c), -1) in front of a Linear is the
// Linear on x plus a rank-`cd` term  c . W[:, K0:K0+cd]^T  (cd = 3 shock-force components).  The big
// part stays on the tensor cores; these kernels add the small term in fp32 and produce its weight
// gradient, addressing the weight columns in place (row pitch ldw = K0 + cd) in the arena.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_COND = 8;

// raw[r][n'] (fp16) += sum_j c[r][j] * W[n_idx[n'] * ldw + col0 + j]   (n_idx: arena row of packed column n')
__global__ void __launch_bounds__(256)
cond_add_f16_kernel(__half* __restrict__ raw, const float* __restrict__ c, const float* __restrict__ W,
                    const int32_t* __restrict__ n_idx, int R, int N, int ldw, int col0, int cd) {
  pdl_sync();
  const int nv = N >> 3;
  const long long total = static_cast<long long>(R) * nv;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(v / nv), n0 = static_cast<int>(v - static_cast<long long>(r) * nv) * 8;
    float cv[MAX_COND];
#pragma unroll
    for (int j = 0; j < MAX_COND; ++j) cv[j] = j < cd ? __ldg(c + static_cast<long long>(r) * cd + j) : 0.0f;
    uint4* p = reinterpret_cast<uint4*>(raw + static_cast<long long>(r) * N + n0);
    uint4 u = *p;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      const float* w0 = W + static_cast<long long>(__ldg(n_idx + n0 + 2 * i)) * ldw + col0;
      const float* w1 = W + static_cast<long long>(__ldg(n_idx + n0 + 2 * i + 1)) * ldw + col0;
      for (int j = 0; j < cd; ++j) {
        f.x = fmaf(cv[j], __ldg(w0 + j), f.x);
        f.y = fmaf(cv[j], __ldg(w1 + j), f.y);
      }
      h[i] = __floats2half2_rn(f.x, f.y);
    }
    *p = u;
  }
}

// dW[n_idx[n'] * ldw + col0 + j] += scale * sum_r g[r][n'] * c[r][j];  grid = (ceil(N/256), row chunks)
__global__ void __launch_bounds__(256)
cond_wgrad_f16_kernel(const __half* __restrict__ g, const float* __restrict__ c, float* __restrict__ dW,
                      const int32_t* __restrict__ n_idx, int R, int N, int ldw, int col0, int cd, float scale,
                      int rows_per_cta) {
  pdl_sync();
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  __shared__ float cs[64][MAX_COND];
  float acc[MAX_COND];
#pragma unroll
  for (int j = 0; j < MAX_COND; ++j) acc[j] = 0.0f;
  for (int rb = r0; rb < r1; rb += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * MAX_COND; i += 256) {
      const int rr = i / MAX_COND, j = i - rr * MAX_COND;
      cs[rr][j] = (rb + rr < r1 && j < cd) ? c[static_cast<long long>(rb + rr) * cd + j] : 0.0f;
    }
    __syncthreads();
    if (n < N) {
      const int lim = min(64, r1 - rb);
      for (int rr = 0; rr < lim; ++rr) {
        const float gv = __half2float(g[static_cast<long long>(rb + rr) * N + n]);
#pragma unroll
        for (int j = 0; j < MAX_COND; ++j) acc[j] = fmaf(gv, cs[rr][j], acc[j]);
      }
    }
  }
  if (n < N) {
    float* o = dW + static_cast<long long>(n_idx[n]) * ldw + col0;
    for (int j = 0; j < cd; ++j) atomicAdd(o + j, scale * acc[j]);
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
  launch_sgemm<false, true>(x, W, y, b, M, N, K, ldx, K, ldy, act, 0, 1.0f, K, 1, ST(stream));
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
    MMDYN_LAUNCH((act_grad_kernel), static_cast<int>(blocks), 256, 0, ST(stream), y, dy, dy_act, M, N, ldy, act);
    LAUNCHED();
  }
  if (dx) {  // dx[M][K] = dy_act[M][N] * W[N][K]
    launch_sgemm<false, false>(dy_act, W, dx, nullptr, M, K, N, N, K, lddx, 0, dx_accumulate ? 1 : 0, 1.0f, N, 1,
                               ST(stream));
    LAUNCHED();
  }
  if (dW) {  // dW[N][K] += scale * dy_act^T[N][M] * x[M][K]
    int splits = (148 * 2) / (((K + BN - 1) / BN) * ((N + 127) / 128));
    if (splits < 1) splits = 1;
    int kps = (M + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (M + kps - 1) / kps;
    launch_sgemm<true, false>(dy_act, x, dW, nullptr, N, K, M, N, ldx, K, 0, 1, scale, kps, splits, ST(stream));
    LAUNCHED();
  }
  if (db) {
    const int rc = mmdyn_colsum_f32(dy_act, db, M, N, N, scale, stream);
    if (rc != MMDYN_OK) return rc;
  }
  return MMDYN_OK;
}

// y[M][N] += x[M][K] . W[N][K]^T with explicit pitches (W may be a column slice of a wider matrix)
extern "C" int mmdyn_linear_f32_acc(const float* x, const float* W, float* y, int M, int N, int K, int ldx, int ldw,
                                    int ldy, void* stream) {
  MMDYN_REQUIRE(x && W && y && M > 0 && N > 0 && K > 0 && ldx >= K && ldw >= K && ldy >= N,
                "linear_f32_acc: bad arguments");
  launch_sgemm<false, true>(x, W, y, nullptr, M, N, K, ldx, ldw, ldy, 0, 1, 1.0f, K, 1, ST(stream));
  LAUNCHED();
  return MMDYN_OK;
}

// dW[N][K] (pitch ldw) += scale * dy[M][N]^T . x[M][K]
extern "C" int mmdyn_linear_f32_wgrad(const float* x, const float* dy, float* dW, int M, int N, int K, int ldx,
                                      int lddy, int ldw, float scale, void* stream) {
  MMDYN_REQUIRE(x && dy && dW && M > 0 && N > 0 && K > 0 && ldx >= K && lddy >= N && ldw >= K,
                "linear_f32_wgrad: bad arguments");
  int splits = (148 * 2) / (((K + BN - 1) / BN) * ((N + 127) / 128));
  if (splits < 1) splits = 1;
  int kps = (M + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  splits = (M + kps - 1) / kps;
  launch_sgemm<true, false>(dy, x, dW, nullptr, N, K, M, lddy, ldx, ldw, 0, 1, scale, kps, splits, ST(stream));
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_cond_add_f16(void* raw, const float* c, const float* W, const int32_t* n_idx, int R, int N,
                                  int ldw, int col0, int cd, void* stream) {
  MMDYN_REQUIRE(raw && c && W && n_idx && R > 0 && N > 0 && N % 8 == 0 && cd >= 1 && cd <= MAX_COND,
                "cond_add_f16: bad arguments (N=%d must be a multiple of 8, 1 <= cd=%d <= %d)", N, cd, MAX_COND);
  const long long total = static_cast<long long>(R) * (N >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  MMDYN_LAUNCH((cond_add_f16_kernel), static_cast<int>(blocks), 256, 0, ST(stream), reinterpret_cast<__half*>(raw), c, W, n_idx,
                                                                         R, N, ldw, col0, cd);
  LAUNCHED();
  return MMDYN_OK;
}

extern "C" int mmdyn_cond_wgrad_f16(const void* g, const float* c, float* dW, const int32_t* n_idx, int R, int N,
                                    int ldw, int col0, int cd, float scale, void* stream) {
  MMDYN_REQUIRE(g && c && dW && n_idx && R > 0 && N > 0 && cd >= 1 && cd <= MAX_COND,
                "cond_wgrad_f16: bad arguments (1 <= cd=%d <= %d)", cd, MAX_COND);
  const int gx = (N + 255) / 256;
  int gy = (148 * 4 + gx - 1) / gx;
  int rpc = (R + gy - 1) / gy;
  rpc = (rpc + 63) / 64 * 64;
  gy = (R + rpc - 1) / rpc;
  MMDYN_LAUNCH((cond_wgrad_f16_kernel), dim3(gx, gy), 256, 0, ST(stream), reinterpret_cast<const __half*>(g), c, dW, n_idx, R, N,
                                                               ldw, col0, cd, scale, rpc);
  LAUNCHED();
  return MMDYN_OK;
}

// y = max(x, 0) over n fp32 values (x == y allowed); x, y 16-byte aligned
extern "C" int mmdyn_relu_f32(const float* x, float* y, long long n, void* stream) {
  MMDYN_REQUIRE(x && y && n > 0, "relu_f32: bad arguments");
  MMDYN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "relu_f32: pointers must be 16-byte aligned");
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  MMDYN_LAUNCH((relu_f32_kernel), static_cast<int>(blocks), 256, 0, ST(stream), x, y, n);
  LAUNCHED();
  return MMDYN_OK;
}

// dx[M][N] = dy * act'(y)  (act 1: ReLU through its OUTPUT y; act 0: copy); y, dy at row pitch ldy, dx dense
extern "C" int mmdyn_act_grad_f32(const float* y, const float* dy, float* dx, int M, int N, int ldy, int act,
                                  void* stream) {
  MMDYN_REQUIRE(y && dy && dx && M > 0 && N > 0 && ldy >= N && (act == 0 || act == 1), "act_grad_f32: bad arguments");
  const long long n = static_cast<long long>(M) * N;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  MMDYN_LAUNCH((act_grad_kernel), static_cast<int>(blocks), 256, 0, ST(stream), y, dy, dx, M, N, ldy, act);
  LAUNCHED();
  return MMDYN_OK;
}
