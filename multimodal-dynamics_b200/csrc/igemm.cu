// Implicit-GEMM convolution / transposed convolution / linear kernels for sm_100a.
//
//   igemm_tma_kernel   : out[row, n] = sum_k A_gather[row, k] * W[n, k]        (forward + dgrad form), both operands by TMA:
//                        weights as 2-D boxes, activations as one 4-D box (channel, x, y, image) per filter tap
//   igemm_pair_kernel  : the same contraction with two 128-row tiles per weight stage (the layers bound by L2 -> SM bytes)
//   igemm_patch_kernel : the merged 3x3-tap layers (4 sub-pixel phases of a stride-2 transposed conv as one GEMM): one
//                        activation box per filter column serves three taps, live weight blocks resident in shared memory,
//                        compile-time MMA schedule, BatchNorm statistics / BCE loss in the epilogue
//   wgrad_tma_kernel   : dW[n, k] += sum_row Nat[row, n] * G_gather[row, k]    (weight gradients), both operands MN-major
//   conv1_*            : the Cin = 3 first encoder layer on the fp32 NCHW input
//   igemm_kernel / wgrad_kernel : cp.async (LDGSTS) gather variants, the fallback for geometries without a TMA mode
//
// Mapping to the hardware (B200):
//   * accumulators live in TMEM (tcgen05.alloc, 128 lanes x BLOCK_N fp32 columns, double-buffered);
//   * one ELECTED thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 16..256, K = 16) on UMMA shared-memory
//     descriptors (128B / 64B swizzle or plain core matrices; K-major for igemm, MN-major for wgrad);
//   * operands arrive by TMA (cp.async.bulk.tensor.{2d,4d}, mbarrier transaction counts), hardware zero fill for padding;
//   * shared-memory rings decouple the TMA producer from the MMA issuer; tcgen05.commit frees ring slots and signals the
//     epilogue warps, which read TMEM with tcgen05.ld and write global memory (256-bit stores for fp16 rows) and can
//     reduce the BatchNorm statistics of what they store;
//   * every kernel is launched with programmatic stream serialisation (pdl_sync, common.cuh).
//
// Reference semantics: nn.Conv2d / nn.ConvTranspose2d / nn.Linear in
// mmdyn/pytorch/models/vae.py:198-216, 263-277 and their autograd (problems.py:153).
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>
#include <cstdlib>

namespace mmdyn {

std::atomic<long long> g_launch_count{0};

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int CTA_THREADS = 192;  // 4 producer/epilogue warps + MMA warp + TMA warp
constexpr int A_STAGE_BYTES = TILE_M * 128;

struct __align__(16) RowInfo {
  int32_t a_off;    // element offset of (img, iy0, ix0, 0) in A
  int32_t out_off;  // element offset of the output pixel, -1: row out of range
  int32_t iy0, ix0;
};

template <int BLOCK_N>
struct Cfg {
  static constexpr int STAGES = (BLOCK_N == 128) ? 3 : 4;
  static constexpr int LAG = STAGES - 2;  // cp.async groups kept in flight per producer thread
  static constexpr int B_STAGE_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 contiguous bytes of one thread: ONE 256-bit store (STG.256, sm_100) when the address allows it — a full 32-byte
// sector per request instead of two half-sector writes (the epilogue stores of the N <= 128 layers were the
// longest stage of their tiles, profiles/r2b_patch_ablation.md)
__device__ __forceinline__ void st_global_32B(void* p, const uint4& u0, const uint4& u1) {
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(u0.x), "r"(u0.y), "r"(u0.z),
                 "r"(u0.w), "r"(u1.x), "r"(u1.y), "r"(u1.z), "r"(u1.w)
                 : "memory");
  } else {
    reinterpret_cast<uint4*>(p)[0] = u0;
    reinterpret_cast<uint4*>(p)[1] = u1;
  }
}

// ---------------------------------------------------------------------------------------------
// igemm: forward / dgrad form — persistent, warp-specialised
//   warps 0-3  A producers (im2col gather, cp.async)      warp 4  MMA issuer (+ TMEM owner)
//   warp  5    B producer (TMA)                           warps 6-9 epilogue (TMEM -> global)
// Every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the smem ring and its phases run
// across tile boundaries, and the accumulator is double-buffered in TMEM, so the epilogue of tile
// i overlaps the loads and MMAs of tile i+1 and the per-CTA set-up (barriers, TMEM allocation)
// is paid once instead of once per 128 output rows (these layers have K as small as 128).
// ---------------------------------------------------------------------------------------------
constexpr int IGEMM_THREADS = 320;
constexpr int EPI_WARP0 = 6;

struct TileCoord {
  int row0, tile_pix, n_tile, phase, split;
};

__device__ __forceinline__ TileCoord decode_tile(const mmdyn_igemm_desc& d, int tile, int m_tiles, int n_tiles) {
  TileCoord t;
  const int m = tile % m_tiles;
  int rest = tile / m_tiles;
  t.n_tile = rest % n_tiles;
  rest /= n_tiles;
  t.phase = rest / d.ksplit;
  t.split = rest - t.phase * d.ksplit;
  if (d.row_mode == 0) {
    t.row0 = m * TILE_M;
    t.tile_pix = 0;
  } else {
    const int img_blocks = (d.n_img + TILE_M - 1) / TILE_M;
    t.tile_pix = m / img_blocks;
    t.row0 = (m - t.tile_pix * img_blocks) * TILE_M;
  }
  return t;
}

template <int BLOCK_N>
__global__ void __launch_bounds__(IGEMM_THREADS)
igemm_kernel(const __grid_constant__ mmdyn_igemm_desc d, const __grid_constant__ CUtensorMap tmW, int m_tiles,
             int n_tiles, int total_tiles) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ RowInfo rows[2][TILE_M];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 4 * 32) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), NUM_PRODUCER_THREADS + 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), NUM_PRODUCER_THREADS);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmW);
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&tmem_base_s), 2 * C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();  // the prologue above touched no global memory: it overlaps the previous grid's tail

  const int kb_total = (d.ntaps * d.Cin) >> 6;
  const int kb_per = (kb_total + d.ksplit - 1) / d.ksplit;
  // row_mode 1: all rows of a tile share the virtual pixel, so out-of-image taps are skipped
  auto kb_live = [&](const TileCoord& t, int kb) -> bool {
    if (d.row_mode == 0) return true;
    const int tyv = t.tile_pix / d.OXv, txv = t.tile_pix - tyv * d.OXv;
    const int tap = (kb << 6) / d.Cin;
    const int iy = tyv * d.s_in + d.tap_dy[t.phase][tap];
    const int ix = txv * d.s_in + d.tap_dx[t.phase][tap];
    return (unsigned)iy < (unsigned)d.IH && (unsigned)ix < (unsigned)d.IW;
  };

  if (warp < 4) {
    // ======================= A producers: im2col gather with cp.async =======================
    const __half* A = reinterpret_cast<const __half*>(d.A);
    const int j = threadIdx.x & 7;      // 16-byte chunk inside the 128-byte k-block row
    const int rsub = threadIdx.x >> 3;  // rows rsub + 16*i
    const uint32_t sw = static_cast<uint32_t>((j ^ (rsub & 7)) << 4);
    int it = 0, tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const TileCoord t = decode_tile(d, tile, m_tiles, n_tiles);
      {
        const int r = threadIdx.x;
        int img, p;
        bool valid;
        if (d.row_mode == 0) {
          const int m = t.row0 + r;
          valid = m < d.n_img * d.P;
          img = m / d.P;
          p = m - img * d.P;
        } else {
          img = t.row0 + r;
          valid = img < d.n_img;
          p = t.tile_pix;
        }
        const int yv = p / d.OXv, xv = p - yv * d.OXv;
        RowInfo ri;
        ri.iy0 = valid ? yv * d.s_in : -(1 << 20);
        ri.ix0 = xv * d.s_in;
        ri.a_off = valid ? ((img * d.IH + yv * d.s_in) * d.IW + xv * d.s_in) * d.a_pix_stride : 0;
        ri.out_off = 0;
        rows[tl & 1][r] = ri;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const RowInfo* rw = rows[tl & 1];
      const int kb_begin = t.split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        if (!kb_live(t, kb)) continue;
        const int s = it % C::STAGES;
        mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
        const int k = (kb << 6) + (j << 3);
        const int tap = k / d.Cin;
        const int c = k - tap * d.Cin;
        const int dy = d.tap_dy[t.phase][tap], dx = d.tap_dx[t.phase][tap];
        const int tap_off = (dy * d.IW + dx) * d.a_pix_stride + c;
        const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = rsub + 16 * i;
          const RowInfo ri = rw[r];
          const bool v = (unsigned)(ri.iy0 + dy) < (unsigned)d.IH && (unsigned)(ri.ix0 + dx) < (unsigned)d.IW;
          const __half* src = v ? (A + (ri.a_off + tap_off)) : A;
          cp_async16(a_stage + r * 128 + sw, src, v);
        }
        cp_async_commit();
        if (it >= C::LAG) {
          // the generic->async proxy fence is issued once by the consumer after it acquires the
          // stage (see the MMA issuer): a per-thread fence here drains every in-flight cp.async
          cp_async_wait<C::LAG>();
          mbar_arrive(smem_u32(&full_bar[(it - C::LAG) % C::STAGES]));
        }
        ++it;
      }
    }
    cp_async_wait<0>();
    for (int q = (it > C::LAG ? it - C::LAG : 0); q < it; ++q) mbar_arrive(smem_u32(&full_bar[q % C::STAGES]));
  } else if (warp == 5) {
    // ======================= B producer: TMA loads of the packed weights =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(d, tile, m_tiles, n_tiles);
        const int kb_begin = t.split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!kb_live(t, kb)) continue;
          const int s = it % C::STAGES;
          mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
          const uint32_t bar = smem_u32(&full_bar[s]);
          mbar_arrive_expect_tx(bar, C::B_STAGE_BYTES);
          tma_load_2d(smem_base + s * C::STAGE_BYTES + A_STAGE_BYTES, &tmW, bar, kb << 6,
                      t.phase * d.N + t.n_tile * BLOCK_N);
          ++it;
        }
      }
    }
  } else if (warp == 4) {
    // ======================= MMA issuer ======================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N, 0, 0, 0, 0);
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
        const TileCoord t = decode_tile(d, tile, m_tiles, n_tiles);
        const int acc = tl & 1;
        mbar_wait(smem_u32(&tempty_bar[acc]), ((tl >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS;
        const int kb_begin = t.split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        int first = 1;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!kb_live(t, kb)) continue;
          const int s = it % C::STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (it / C::STAGES) & 1);
          fence_proxy_async_smem();  // producers' cp.async (generic proxy) writes -> tcgen05 (async proxy) reads
          tc_fence_after();
          const uint32_t a_base = smem_base + s * C::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_base, 16, 1024, LAYOUT_SW128);
          const uint64_t bdesc = make_smem_desc(a_base + A_STAGE_BYTES, 16, 1024, LAYOUT_SW128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // 4 x (K = 16) per 64-wide k-block: +32 B per step
            umma_f16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
          first = 0;
          umma_commit(smem_u32(&empty_bar[s]));
          ++it;
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
      }
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> global ============================
    const int q4 = warp & 3;               // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;          // tile row owned by this thread
    int tl = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
      const TileCoord t = decode_tile(d, tile, m_tiles, n_tiles);
      int out_off;
      {
        int img, p;
        bool valid;
        if (d.row_mode == 0) {
          const int m = t.row0 + r;
          valid = m < d.n_img * d.P;
          img = m / d.P;
          p = m - img * d.P;
        } else {
          img = t.row0 + r;
          valid = img < d.n_img;
          p = t.tile_pix;
        }
        const int yv = p / d.OXv, xv = p - yv * d.OXv;
        if (d.out_mode == 3) {
          out_off = valid ? ((img * 3 * d.OH) + 2 * yv) * d.OW + 2 * xv : -1;
        } else {
          const int oy = yv * d.s_out + d.off_y[t.phase], ox = xv * d.s_out + d.off_x[t.phase];
          out_off = valid ? ((img * d.OH + oy) * d.OW + ox) * d.ldc : -1;
        }
      }
      int n_live = 0;
      {
        const int kb_begin = t.split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        if (d.row_mode == 0) n_live = kb_end - kb_begin;
        else
          for (int kb = kb_begin; kb < kb_end; ++kb) n_live += kb_live(t, kb) ? 1 : 0;
      }
      const int acc = tl & 1;
      mbar_wait(smem_u32(&tfull_bar[acc]), (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS + (static_cast<uint32_t>(q4 * 32) << 16);
      const int n_base = t.n_tile * BLOCK_N;
      const bool add_bias = d.bias != nullptr && t.split == 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_d + c0, v);
        tmem_ld_wait(v);
        if (out_off < 0) continue;
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) f[q] = n_live > 0 ? __uint_as_float(v[q]) : 0.0f;
        if (add_bias) {
#pragma unroll
          for (int q = 0; q < 16; ++q) f[q] += __ldg(d.bias + n_base + c0 + q);
        }
        if (d.out_mode == 4) {
          // merged sub-pixel phases, NHWC fp16: this 16-column chunk belongs to one phase
          const int n = n_base + c0, phs = n / d.ldc, co0 = n - phs * d.ldc;
          __half* o = reinterpret_cast<__half*>(d.out) + out_off + ((phs >> 1) * d.OW + (phs & 1)) * d.ldc + co0;
          st_global_32B(o, make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7])),
                        make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]), pack_h2(f[14], f[15])));
        } else if (d.out_mode == 0) {
          __half* o = reinterpret_cast<__half*>(d.out) + out_off + n_base + c0;
          uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]),
                                pack_h2(f[14], f[15]));
          st_global_32B(o, u0, u1);
        } else if (d.out_mode == 1) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) + out_off + n_base + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else if (d.out_mode == 2) {
          float* o = reinterpret_cast<float*>(d.out) + out_off + n_base + c0;  // 16-byte aligned (ldc % 4 == 0)
#pragma unroll
          for (int q = 0; q < 4; ++q) red_add_v4(o + 4 * q, f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else {
          // merged 2x2 sub-pixel phases -> fp32 NCHW planes; n = (ph*2 + pw)*3 + c
          float* o = reinterpret_cast<float*>(d.out) + out_off;
          const int plane = d.OH * d.OW;
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ph = 0; ph < 2; ++ph)
              *reinterpret_cast<float2*>(o + c * plane + ph * d.OW) =
                  make_float2(f[(ph * 2 + 0) * 3 + c], f[(ph * 2 + 1) * 3 + c]);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[acc]));  // accumulator stage free for tile tl + 2
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// igemm, TMA-fed: the activation operand is fetched by the TMA unit as well — one 4-D tiled box
// (channels, x, y, image) per filter tap, element strides for stride-2 gathers, hardware zero
// fill for the padding halo and for partial tiles — so no LSU instruction touches the operands.
//   warp 0  TMA producer (A boxes + weight box per 64-wide k-block, one mbarrier expect_tx)
//   warp 1  MMA issuer (tcgen05.mma, accumulators double-buffered in TMEM)
//   warps 2-5 epilogue (tcgen05.ld -> registers -> global)
// A tile is 128 GEMM rows = a box of bw x bh virtual pixels x bn images (bw*bh*bn = 128), or, for
// the 5x5 / single-pixel layers, one virtual pixel x 128 images ("pixel-major").
// A_MODE: 0 = Cin % 64 == 0 (one 128B-swizzled box per k-block), 1 = Cin == 32 (two 64B-swizzled
// boxes = two taps), 2 = Cin == 8 (eight un-swizzled 16-byte boxes = eight taps), 3 = Cin == 16: the 4-channel
// logit gradient read through 4-pixel (32-byte) windows that advance 16 bytes per output pixel.  The windows
// are never materialised: per row tap the raw padded image rows are fetched twice as 512-byte runs (pairs
// 0..31 and pairs 1..32 of each of the tile's 4 rows), and an un-swizzled K-major descriptor whose two 16-byte
// K chunks point into the two copies (LBO = 2 KB) with 8-row groups 128 bytes apart reads window x as pairs
// x, x + 1 — 32 TMA row requests of 512 bytes per tile instead of 512 requests of 32 bytes.
// ---------------------------------------------------------------------------------------------
constexpr int TMA_THREADS = 192;

struct TileGeom {
  int lbw, lbh;        // log2 of the box width / height in virtual pixels (image-box mode)
  int bn;              // images per tile
  int pixel_major;     // 1: tile = one virtual pixel x 128 images
  int tiles_y;         // image-box mode: OYv / bh
  int img_blocks;      // ceil(n_img / bn)
  int m_tiles, n_tiles, total_tiles;
};

struct TileCoord2 {
  int img0, y0, x0, n_tile, phase, split;
};

__device__ __forceinline__ TileCoord2 decode_tile2(const mmdyn_igemm_desc& d, const TileGeom& g, int tile) {
  TileCoord2 t;
  const int m = tile % g.m_tiles;
  int rest = tile / g.m_tiles;
  t.n_tile = rest % g.n_tiles;
  rest /= g.n_tiles;
  t.phase = rest / d.ksplit;
  t.split = rest - t.phase * d.ksplit;
  if (g.pixel_major) {
    const int pix = m / g.img_blocks;
    t.img0 = (m - pix * g.img_blocks) * TILE_M;
    t.y0 = pix / d.OXv;
    t.x0 = pix - t.y0 * d.OXv;
  } else {
    const int ib = m / g.tiles_y;
    t.img0 = ib * g.bn;
    t.y0 = (m - ib * g.tiles_y) << g.lbh;
    t.x0 = 0;
  }
  return t;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// The per-k-block loops are kept free of integer divisions and table walks: the producer tracks (tap, channel
// block) incrementally, and for pixel-major tiles all three roles derive ONE live-tap bit mask per tile (a tap
// whose input pixel falls outside the image is skipped, not multiplied by zeros) — evaluating that test with
// its divisions per k-block in both the producer and the MMA issuer cost as much as a live k-block's MMAs on
// the 5x5 <-> 8x8 layers (ncu source view, round 2).  Producer and MMA issuer run under elect.sync.
// EG epilogue groups of 4 warps: group e drains the tiles whose accumulator stage is e.  EG = 2 for BLOCK_N = 256 (one CTA
// per SM: its 4 epilogue warps were busy 80 % of the kernel on deconv2.fwd, profiles/r2_stalls_igemm_tma_256_deconv2fwd.txt).
template <int BLOCK_N, int A_MODE, int EG>
__global__ void __launch_bounds__(64 + 128 * EG)
igemm_tma_kernel(const __grid_constant__ mmdyn_igemm_desc d, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmW, const TileGeom g) {
  using C = Cfg<BLOCK_N>;
  const int tile0 = static_cast<int>(blockIdx.x);
  const int tile_step = static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_all[EG][BLOCK_N];
  // BatchNorm statistics of the fp16 outputs (out_mode 0 + bn_sums, the conv / 5x5 -> 8x8 deconv layers): per epilogue
  // group {sum x, sum x^2} per channel, flushed to bn_sums[group] with one atomic per value, CTA and group
  constexpr bool STATS = (A_MODE == 0 || A_MODE == 1) && BLOCK_N >= 64;
  __shared__ float stat_all[STATS ? EG : 1][STATS ? 2 * BLOCK_N : 2];
  if constexpr (STATS)
    for (int i = threadIdx.x; i < EG * 2 * BLOCK_N; i += blockDim.x) (&stat_all[0][0])[i] = 0.0f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 128);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), 2 * C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();  // the prologue above touched no global memory: it overlaps the previous grid's tail

  const int kb_total = (d.ntaps * d.Cin) >> 6;
  const int kb_per = (kb_total + d.ksplit - 1) / d.ksplit;
  const int CB = A_MODE == 0 ? (d.Cin >> 6) : 1;  // 64-channel blocks per tap (A_MODE 0)
  // (linear layers are pixel-major tiles with one tap that is always live, and may split K across CTAs)
  const bool skip_taps = g.pixel_major && A_MODE == 0 && d.ksplit == 1 && d.ntaps > 1;
  // pixel-major tiles share the virtual pixel: bit t = tap t reads a pixel inside the image
  auto live_taps = [&](const TileCoord2& t) -> uint32_t {
    if (!skip_taps) return 0xFFFFFFFFu;
    uint32_t m = 0;
    const int by = t.y0 * d.s_in, bx = t.x0 * d.s_in;
    for (int tap = 0; tap < d.ntaps; ++tap) {
      const int iy = by + d.tap_dy[t.phase][tap], ix = bx + d.tap_dx[t.phase][tap];
      if ((unsigned)iy < (unsigned)d.IH && (unsigned)ix < (unsigned)d.IW) m |= 1u << tap;
    }
    return m;
  };
  // k-blocks of a tile that are actually loaded and multiplied (row_mode 1 implies ksplit 1)
  auto live_kbs = [&](const TileCoord2& t) -> int {
    if (skip_taps) return __popc(live_taps(t)) * CB;
    const int kb_begin = t.split * kb_per;
    return min(kb_total, kb_begin + kb_per) - kb_begin;
  };

  if (warp == 0) {
    // ======================= TMA producer =====================================================
    if (elect_one()) {
      int it = 0;
      for (int tile = tile0; tile < g.total_tiles; tile += tile_step) {
        const TileCoord2 t = decode_tile2(d, g, tile);
        const int kb_begin = t.split * kb_per, kb_end = min(kb_total, kb_begin + kb_per);
        const int wx = t.x0 * (d.s_in_x ? d.s_in_x : d.s_in), wy = t.y0 * d.s_in;
        const uint32_t mask = live_taps(t);
        const int w_row = t.phase * d.N + t.n_tile * BLOCK_N;
        int tap = A_MODE == 0 ? kb_begin / CB : 0, cb = A_MODE == 0 ? kb_begin - tap * CB : 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          bool live = true;
          int tdx = 0, tdy = 0, c0 = 0;
          if (A_MODE == 0) {
            live = (mask >> tap) & 1u;
            tdx = d.tap_dx[t.phase][tap];
            tdy = d.tap_dy[t.phase][tap];
            c0 = cb << 6;
            if (++cb == CB) {
              cb = 0;
              ++tap;
            }
          }
          if (!live) continue;
          const int s = it % C::STAGES;
          mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
          const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
          const uint32_t bar = smem_u32(&full_bar[s]);
          mbar_arrive_expect_tx(bar, A_STAGE_BYTES + C::B_STAGE_BYTES);
          if (A_MODE == 0) {
            tma_load_4d(a_stage, &tmA, bar, c0, wx + tdx, wy + tdy, t.img0);
          } else if (A_MODE == 1) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const int tp = 2 * kb + b;
              tma_load_4d(a_stage + b * 8192, &tmA, bar, 0, wx + d.tap_dx[t.phase][tp], wy + d.tap_dy[t.phase][tp], t.img0);
            }
          } else if (A_MODE == 3) {
#pragma unroll
            for (int b = 0; b < 8; ++b)  // (row tap b / 2, copy b % 2 shifted by one pixel pair = 8 elements)
              tma_load_4d(a_stage + b * 2048, &tmA, bar, (b & 1) * 8, 0, wy + d.tap_dy[t.phase][4 * kb + (b >> 1)], t.img0);
          } else {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              const int tp = 8 * kb + b;
              tma_load_4d(a_stage + b * 2048, &tmA, bar, 0, wx + d.tap_dx[t.phase][tp], wy + d.tap_dy[t.phase][tp], t.img0);
            }
          }
          tma_load_2d(a_stage + A_STAGE_BYTES, &tmW, bar, kb << 6, w_row);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer ======================================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N, 0, 0, 0, 0);
      // descriptor = HI | (address >> 4); B: 128B swizzle, 8-row atoms of 1024 B
      constexpr uint64_t B_HI = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                                (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(LAYOUT_SW128) << 61);
      constexpr uint64_t A_HI = A_MODE == 0 ? B_HI
                                : A_MODE == 1 ? ((static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
                                                 (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(LAYOUT_SW64) << 61))
                                              : ((static_cast<uint64_t>(2048 >> 4) << 16) | (static_cast<uint64_t>(128 >> 4) << 32) |
                                                 (static_cast<uint64_t>(1) << 46));
      int it = 0, tl = 0;
      for (int tile = tile0; tile < g.total_tiles; tile += tile_step, ++tl) {
        const TileCoord2 t = decode_tile2(d, g, tile);
        const int n_kb = live_kbs(t);
        const int acc = tl & 1;
        mbar_wait(smem_u32(&tempty_bar[acc]), ((tl >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS;
        for (int i = 0; i < n_kb; ++i) {
          const int s = it % C::STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (it / C::STAGES) & 1);
          tc_fence_after();
          const uint32_t a16 = (smem_base + s * C::STAGE_BYTES) >> 4;
          const uint32_t b16 = a16 + (A_STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x (K = 16) per 64-wide k-block
            const uint32_t a_off = A_MODE == 0 ? 2 * kk : (A_MODE == 1 ? (kk >> 1) * (8192 >> 4) + (kk & 1) * 2 : kk * (4096 >> 4));  // modes 2, 3: 4 KB per K = 16
            umma_f16(tmem_d, A_HI | static_cast<uint64_t>(a16 + a_off), B_HI | static_cast<uint64_t>(b16 + 2 * kk), idesc,
                     (i | kk) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));
          ++it;
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
      }
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> global ============================
    const int q4 = warp & 3;       // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;  // tile row owned by this thread
    const int eg = EG == 1 ? 0 : (warp - 2) >> 2;  // epilogue group: tiles tl = eg, eg + EG, ...
    const int et = (threadIdx.x - 64) & 127;       // thread index inside the group
    float* bias_s = bias_all[eg];
    int x_l, y_l, n_l;
    if (g.pixel_major) {
      x_l = 0; y_l = 0; n_l = r;
    } else {
      x_l = r & ((1 << g.lbw) - 1);
      y_l = (r >> g.lbw) & ((1 << g.lbh) - 1);
      n_l = r >> (g.lbw + g.lbh);
    }
    // out_mode 5: BCE partial sum of this thread for loss slot bce_cur; a tile lies within one image
    // (checked by the launcher), so slot changes are warp-uniform and rare (<= groups per CTA)
    const bool want_stats = STATS && d.out_mode == 0 && d.bn_sums != nullptr;
    float* stat_s = stat_all[STATS ? eg : 0];
    int stat_grp = -1;
    auto stat_flush = [&]() {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
      for (int i = et; i < 2 * BLOCK_N; i += 128) {
        const float v = stat_s[i];
        if (stat_grp >= 0 && v != 0.0f)
          atomicAdd(d.bn_sums + (static_cast<long long>(stat_grp) * d.N + (i >> 1)) * 2 + (i & 1), v);
        stat_s[i] = 0.0f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
    };
    float bce_acc = 0.0f;
    int bce_cur = -1;
    auto bce_flush = [&]() {
      const float v = warp_sum(bce_acc);
      if (lane == 0 && bce_cur >= 0) atomicAdd(d.bce_loss + bce_cur, v);
      bce_acc = 0.0f;
    };
    for (int tile = tile0 + eg * tile_step, tl = eg; tile < g.total_tiles; tile += EG * tile_step, tl += EG) {
      const TileCoord2 t = decode_tile2(d, g, tile);
      const int img = t.img0 + n_l, yv = t.y0 + y_l, xv = t.x0 + x_l;
      const bool valid = img < d.n_img;
      int out_off;
      if (d.out_mode == 3 || d.out_mode == 5) {
        out_off = valid ? ((img * 3 * d.OH) + 2 * yv) * d.OW + 2 * xv : -1;
      } else {
        const int oy = yv * d.s_out + d.off_y[t.phase], ox = xv * d.s_out + d.off_x[t.phase];
        out_off = valid ? ((img * d.OH + oy) * d.OW + ox) * d.ldc : -1;
      }
      const int n_live = live_kbs(t);
      if (want_stats) {
        // all images of a tile belong to one group (checked by the launcher)
        const int grp = t.img0 / d.bn_rows_per_group;
        if (grp != stat_grp) {
          stat_flush();
          stat_grp = grp;
        }
      }
      // out_mode 5: the targets (and mask) of this thread's 2x2x3 output pixels depend on the tile
      // coordinates only, so they are requested BEFORE waiting for the accumulator: their latency
      // hides behind the tile's MMAs instead of serialising the epilogue
      float2 bce_t[6], bce_m[6];
      int bce_slot_now = -1;
      if constexpr (BLOCK_N == 16) {
        if (d.out_mode == 5 && valid) {
          const int grp = img / d.bce_rows_per_group;
          bce_slot_now = d.bce_slot[grp];
          if (bce_slot_now >= 0) {
            const int plane = d.OH * d.OW;
            const int toff = (((img - grp * d.bce_rows_per_group) * 3) * d.OH + 2 * yv) * d.OW + 2 * xv;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ph = 0; ph < 2; ++ph) {
                bce_t[c * 2 + ph] = __ldg(reinterpret_cast<const float2*>(d.bce_target + toff + c * plane + ph * d.OW));
                bce_m[c * 2 + ph] = d.bce_mask ? __ldg(reinterpret_cast<const float2*>(d.bce_mask + toff + c * plane +
                                                                                       ph * d.OW))
                                               : make_float2(1.0f, 1.0f);
              }
          }
        }
      }
      // bias of this tile's BLOCK_N columns -> shared memory, once per tile and BEFORE the accumulator
      // wait (16 dependent L2 round trips per tile inside the chunk loop made the K = 256 / 512 linear
      // layers epilogue-bound: ncu long_scoreboard 6 warps per issue, tensor pipe 8.6 %)
      if (d.bias != nullptr && t.split == 0) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");  // previous tile's readers are done
        for (int i = et; i < BLOCK_N; i += 128) bias_s[i] = __ldg(d.bias + t.n_tile * BLOCK_N + i);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
      }
      const int acc = tl & 1;
      mbar_wait(smem_u32(&tfull_bar[acc]), (tl >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS + (static_cast<uint32_t>(q4 * 32) << 16);
      const int n_base = t.n_tile * BLOCK_N;
      const bool add_bias = d.bias != nullptr && t.split == 0;
      // one 16-column chunk of this thread's accumulator row -> global memory
      auto emit = [&](const uint32_t (&v)[16], const int c0) {
        if constexpr (STATS) {
          if (want_stats) {
            // fp16 row chunk + its statistics: every lane takes part in the reduction, rows past the last image add zeros
            float f[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] = (n_live > 0 && out_off >= 0) ? __uint_as_float(v[q]) : 0.0f;
            const uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
            const uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]),
                                        pack_h2(f[14], f[15]));
            if (out_off >= 0) st_global_32B(reinterpret_cast<__half*>(d.out) + out_off + n_base + c0, u0, u1);
            // statistics of the values the BatchNorm kernels will read: the fp16-ROUNDED outputs.  32 values (16 sums,
            // 16 sums of squares) over the warp's 32 rows: butterfly reduce-scatter, 31 shuffles
            const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            float w[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
              w[2 * q] = a.x;
              w[2 * q + 1] = a.y;
              w[16 + 2 * q] = a.x * a.x;
              w[16 + 2 * q + 1] = a.y * a.y;
            }
#pragma unroll
            for (int st = 0; st < 5; ++st) {
              const int off = 16 >> st, cnt = 16 >> st;
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < cnt; ++j) {
                const float mine = upper ? w[j + cnt] : w[j];
                const float send = upper ? w[j] : w[j + cnt];
                w[j] = mine + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            atomicAdd(&stat_s[(c0 + (lane & 15)) * 2 + (lane >> 4)], w[0]);
            return;
          }
        }
        if (out_off < 0) return;
#ifdef MMDYN_EXP_NOSTORE  // timing experiment only: how much of the kernel are the fp16 epilogue stores?
        if (d.out_mode == 0 || d.out_mode == 4) return;
#endif
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) f[q] = n_live > 0 ? __uint_as_float(v[q]) : 0.0f;
        if (add_bias) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(&bias_s[c0 + 4 * q]);  // same address in every lane
            f[4 * q] += bq.x; f[4 * q + 1] += bq.y; f[4 * q + 2] += bq.z; f[4 * q + 3] += bq.w;
          }
        }
        if (d.out_mode == 4) {
          // merged sub-pixel phases, NHWC fp16: this 16-column chunk belongs to one phase
          const int n = n_base + c0, phs = n / d.ldc, co0 = n - phs * d.ldc;
          __half* o = reinterpret_cast<__half*>(d.out) + out_off + ((phs >> 1) * d.OW + (phs & 1)) * d.ldc + co0;
          st_global_32B(o, make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7])),
                        make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]), pack_h2(f[14], f[15])));
        } else if (d.out_mode == 0) {
          __half* o = reinterpret_cast<__half*>(d.out) + out_off + n_base + c0;
          uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
          uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]),
                                pack_h2(f[14], f[15]));
          st_global_32B(o, u0, u1);
        } else if (d.out_mode == 1) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(d.out) + out_off + n_base + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else if (d.out_mode == 2) {
          float* o = reinterpret_cast<float*>(d.out) + out_off + n_base + c0;  // 16-byte aligned (ldc % 4 == 0)
#pragma unroll
          for (int q = 0; q < 4; ++q) red_add_v4(o + 4 * q, f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else if (d.out_mode == 5) {
          if constexpr (BLOCK_N == 16) {
            const int plane = d.OH * d.OW;
            if (img >= d.logit_row_lo && img < d.logit_row_hi) {
              float* o = reinterpret_cast<float*>(d.out) + out_off;
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ph = 0; ph < 2; ++ph)
                  *reinterpret_cast<float2*>(o + c * plane + ph * d.OW) =
                      make_float2(f[(ph * 2 + 0) * 3 + c], f[(ph * 2 + 1) * 3 + c]);
            }
            const int slot = bce_slot_now;
            if (slot >= 0) {
              if (slot != bce_cur) {
                bce_flush();
                bce_cur = slot;
              }
              float gq[4][4];
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                  const float2 tv = bce_t[c * 2 + ph], mv = bce_m[c * 2 + ph];
                  const float ts[2] = {tv.x, tv.y}, ms[2] = {mv.x, mv.y};
#pragma unroll
                  for (int pw = 0; pw < 2; ++pw) {
                    const float m = ms[pw], x = f[(ph * 2 + pw) * 3 + c] * m, tt = ts[pw] * m;
                    // torch's stable form max(x,0) - x*t + log(1 + exp(-|x|)) on the SFU (ex2, lg2, rcp):
                    // log(1+e) loses relative accuracy only where the term is < 1e-7 of the O(1) summands,
                    // and the 4 epilogue warps of a CTA cannot afford ~100 instructions per logit
                    const float e = __expf(-fabsf(x));
                    const float r = __fdividef(1.0f, 1.0f + e);
                    bce_acc += fmaxf(x, 0.0f) - x * tt + __logf(1.0f + e);
                    const float sig = x >= 0.0f ? r : e * r;
                    gq[ph * 2 + pw][c] = d.bce_gscale * (sig - tt) * m;
                  }
                }
              if (d.bce_dlogits) {
                __half* go = reinterpret_cast<__half*>(d.bce_dlogits);
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                  const long long pix =
                      (static_cast<long long>(img) * (d.OH + 2) + 2 * yv + ph + 1) * (d.OW + 2) + 2 * xv + 1;
                  uint2* o2 = reinterpret_cast<uint2*>(go + pix * 4);  // NHWC4: 8 bytes per pixel
#pragma unroll
                  for (int pw = 0; pw < 2; ++pw) {
                    const float* q = gq[ph * 2 + pw];
                    o2[pw] = make_uint2(pack_h2(q[0], q[1]), pack_h2(q[2], 0.0f));
                  }
                }
              }
            }
          }
        } else {
          // merged 2x2 sub-pixel phases -> fp32 NCHW planes; n = (ph*2 + pw)*3 + c
          float* o = reinterpret_cast<float*>(d.out) + out_off;
          const int plane = d.OH * d.OW;
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ph = 0; ph < 2; ++ph)
              *reinterpret_cast<float2*>(o + c * plane + ph * d.OW) =
                  make_float2(f[(ph * 2 + 0) * 3 + c], f[(ph * 2 + 1) * 3 + c]);
        }
      };
      // TMEM reads are software-pipelined: the load of chunk c+1 is in flight while chunk c is
      // converted and stored (one tcgen05.ld round trip per 16 columns used to be fully exposed)
      uint32_t va[16], vb[16];
      tmem_ld_x16(tmem_d, va);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld_wait(va);  // ties va to the wait: no read of va may be scheduled before it
        if (c0 + 16 < BLOCK_N) tmem_ld_x16(tmem_d + c0 + 16, vb);
        emit(va, c0);
        if (c0 + 16 < BLOCK_N) {
          tmem_ld_wait(vb);
          if (c0 + 32 < BLOCK_N) tmem_ld_x16(tmem_d + c0 + 32, va);
          emit(vb, c0 + 16);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[acc]));  // accumulator stage free for tile tl + 2
    }
    if (d.out_mode == 5) bce_flush();
    if (want_stats) stat_flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * C::TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------
// igemm, TWO 128-row tiles per weight stage (Cin % 64 == 0, fp16 row output, one N tile, no bias / K split).
// The N >= 128 conv layers run at the L2 -> SM throughput cap (LTS ~ 6.3 KB/clk for the chip = 43 B/clk/SM: every
// 64-wide k-block of a 128 x N tile moves 16 KB of activations + N * 128 B of weights through it; deconv1.fwd
// moves 1.64 GB in 0.115 ms = 14 TB/s), not at the tensor pipe.  Here a CTA owns the tile PAIR (2p, 2p + 1) — same
// virtual pixel or box row, neighbouring image blocks — and each weight k-block is fetched ONCE for both:
// (2 * 16 KB + N * 128 B) per 2 * 128 * N * 64 MACs, 25 % (N = 128) / 33 % (N = 256) fewer L2 bytes per MAC.
//   warp 0 TMA producer, warp 1 MMA issuer (8 MMAs per k-block: 4 per tile, one B descriptor),
//   epilogue group e (4 warps) drains tile e of the pair.  TMEM: NACC stages x 2 tiles x N columns.
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct Cfg2 {
  static constexpr int STAGE_BYTES = 2 * A_STAGE_BYTES + BLOCK_N * 128;
  static constexpr int STAGES = BLOCK_N == 256 ? 3 : 4;
  static constexpr int NACC = BLOCK_N == 256 ? 1 : 2;
  static constexpr int TMEM_COLS = NACC * 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

template <int BLOCK_N, int A_MODE>
__global__ void __launch_bounds__(320)
igemm_pair_kernel(const __grid_constant__ mmdyn_igemm_desc d, const __grid_constant__ CUtensorMap tmA,
                  const __grid_constant__ CUtensorMap tmW, const TileGeom g) {
  using C = Cfg2<BLOCK_N>;
  constexpr int NACC = C::NACC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[NACC];
  __shared__ __align__(8) uint64_t tempty_bar[NACC];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_all[2][2 * BLOCK_N];
  for (int i = threadIdx.x; i < 4 * BLOCK_N; i += blockDim.x) (&stat_all[0][0])[i] = 0.0f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 256);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();

  const int n_pairs = g.total_tiles >> 1;
  const int CB = A_MODE == 0 ? (d.Cin >> 6) : 1;
  const bool skip_taps = A_MODE == 0 && g.pixel_major && d.ntaps > 1;
  // both tiles of a pair share the virtual pixel (pixel-major) or have no out-of-image skipping (box tiles)
  auto live_taps = [&](const TileCoord2& t) -> uint32_t {
    if (!skip_taps) return (d.ntaps >= 32) ? 0xFFFFFFFFu : ((1u << d.ntaps) - 1u);
    uint32_t m = 0;
    const int by = t.y0 * d.s_in, bx = t.x0 * d.s_in;
    for (int tap = 0; tap < d.ntaps; ++tap) {
      const int iy = by + d.tap_dy[0][tap], ix = bx + d.tap_dx[0][tap];
      if ((unsigned)iy < (unsigned)d.IH && (unsigned)ix < (unsigned)d.IW) m |= 1u << tap;
    }
    return m;
  };

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        const TileCoord2 t0 = decode_tile2(d, g, 2 * p), t1 = decode_tile2(d, g, 2 * p + 1);
        const uint32_t mask = live_taps(t0);
        const int wx0 = t0.x0 * d.s_in, wy0 = t0.y0 * d.s_in, wx1 = t1.x0 * d.s_in, wy1 = t1.y0 * d.s_in;
        if constexpr (A_MODE == 1) {  // Cin = 32: two taps (two 64B-swizzled boxes) per k-block and tile, box tiles only
          for (int kb = 0; kb < (d.ntaps >> 1); ++kb) {
            const int s = it % C::STAGES;
            mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
            const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
            const uint32_t bar = smem_u32(&full_bar[s]);
            mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const int tdx = d.tap_dx[0][2 * kb + b], tdy = d.tap_dy[0][2 * kb + b];
              tma_load_4d(a_stage + b * 8192, &tmA, bar, 0, wx0 + tdx, wy0 + tdy, t0.img0);
              tma_load_4d(a_stage + A_STAGE_BYTES + b * 8192, &tmA, bar, 0, wx1 + tdx, wy1 + tdy, t1.img0);
            }
            tma_load_2d(a_stage + 2 * A_STAGE_BYTES, &tmW, bar, kb << 6, 0);
            ++it;
          }
          continue;
        }
        int kb = 0;
        for (int tap = 0; tap < d.ntaps; ++tap) {
          if (!((mask >> tap) & 1u)) {
            kb += CB;
            continue;
          }
          const int tdx = d.tap_dx[0][tap], tdy = d.tap_dy[0][tap];
          for (int cb = 0; cb < CB; ++cb, ++kb) {
            const int s = it % C::STAGES;
            mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
            const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
            const uint32_t bar = smem_u32(&full_bar[s]);
            mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
            tma_load_4d(a_stage, &tmA, bar, cb << 6, wx0 + tdx, wy0 + tdy, t0.img0);
            tma_load_4d(a_stage + A_STAGE_BYTES, &tmA, bar, cb << 6, wx1 + tdx, wy1 + tdy, t1.img0);
            tma_load_2d(a_stage + 2 * A_STAGE_BYTES, &tmW, bar, kb << 6, 0);
            ++it;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N, 0, 0, 0, 0);
      constexpr uint64_t HI = (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                              (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(LAYOUT_SW128) << 61);
      constexpr uint64_t A_HI = A_MODE == 0 ? HI
                                            : ((static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(512 >> 4) << 32) |
                                               (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(LAYOUT_SW64) << 61));
      int it = 0, tl = 0;
      for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++tl) {
        const TileCoord2 t0 = decode_tile2(d, g, 2 * p);
        const int n_kb = A_MODE == 0 ? __popc(live_taps(t0)) * CB : (d.ntaps >> 1);
        const int acc = tl % NACC;
        mbar_wait(smem_u32(&tempty_bar[acc]), ((tl / NACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * BLOCK_N;
        for (int i = 0; i < n_kb; ++i) {
          const int s = it % C::STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (it / C::STAGES) & 1);
          tc_fence_after();
          const uint32_t a16 = (smem_base + s * C::STAGE_BYTES) >> 4;
          const uint32_t b16 = a16 + ((2 * A_STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = HI | static_cast<uint64_t>(b16 + 2 * kk);
            const uint32_t a_off = A_MODE == 0 ? 2 * kk : (kk >> 1) * (8192 >> 4) + (kk & 1) * 2;
            umma_f16(tmem_d, A_HI | static_cast<uint64_t>(a16 + a_off), bdesc, idesc, (i | kk) ? 1u : 0u);
            umma_f16(tmem_d + BLOCK_N, A_HI | static_cast<uint64_t>(a16 + (A_STAGE_BYTES >> 4) + a_off), bdesc, idesc,
                     (i | kk) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[s]));
          ++it;
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
      }
    }
  } else {
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int eg = (warp - 2) >> 2;             // epilogue group = tile of the pair
    const int et = (threadIdx.x - 64) & 127;
    int x_l, y_l, n_l;
    if (g.pixel_major) {
      x_l = 0; y_l = 0; n_l = r;
    } else {
      x_l = r & ((1 << g.lbw) - 1);
      y_l = (r >> g.lbw) & ((1 << g.lbh) - 1);
      n_l = r >> (g.lbw + g.lbh);
    }
    const bool want_stats = d.bn_sums != nullptr;
    float* stat_s = stat_all[eg];
    int stat_grp = -1;
    auto stat_flush = [&]() {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
      for (int i = et; i < 2 * BLOCK_N; i += 128) {
        const float v = stat_s[i];
        if (stat_grp >= 0 && v != 0.0f)
          atomicAdd(d.bn_sums + (static_cast<long long>(stat_grp) * d.N + (i >> 1)) * 2 + (i & 1), v);
        stat_s[i] = 0.0f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
    };
    int tl = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++tl) {
      const TileCoord2 t = decode_tile2(d, g, 2 * p + eg);
      const int img = t.img0 + n_l, yv = t.y0 + y_l, xv = t.x0 + x_l;
      const bool valid = img < d.n_img;
      const int oy = yv * d.s_out + d.off_y[0], ox = xv * d.s_out + d.off_x[0];
      const int out_off = valid ? ((img * d.OH + oy) * d.OW + ox) * d.ldc : -1;
      const bool has_k = live_taps(t) != 0u;
      if (want_stats) {
        const int grp = t.img0 / d.bn_rows_per_group;
        if (grp != stat_grp) {
          stat_flush();
          stat_grp = grp;
        }
      }
      const int acc = tl % NACC;
      mbar_wait(smem_u32(&tfull_bar[acc]), (tl / NACC) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (acc * 2 + eg) * BLOCK_N + (static_cast<uint32_t>(q4 * 32) << 16);
      auto emit = [&](const uint32_t (&v)[16], const int c0) {
        float f[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) f[q] = (has_k && out_off >= 0) ? __uint_as_float(v[q]) : 0.0f;
        const uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
        const uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]), pack_h2(f[14], f[15]));
        if (out_off >= 0) st_global_32B(reinterpret_cast<__half*>(d.out) + out_off + c0, u0, u1);
        if (want_stats) {
          // statistics of the fp16-ROUNDED outputs over the warp's 32 rows: butterfly reduce-scatter, 31 shuffles
          const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          float w[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
            w[2 * q] = a.x;
            w[2 * q + 1] = a.y;
            w[16 + 2 * q] = a.x * a.x;
            w[16 + 2 * q + 1] = a.y * a.y;
          }
#pragma unroll
          for (int st = 0; st < 5; ++st) {
            const int off = 16 >> st, cnt = 16 >> st;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < cnt; ++j) {
              const float mine = upper ? w[j + cnt] : w[j];
              const float send = upper ? w[j] : w[j + cnt];
              w[j] = mine + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          atomicAdd(&stat_s[(c0 + (lane & 15)) * 2 + (lane >> 4)], w[0]);
        }
      };
      uint32_t va[16], vb[16];
      tmem_ld_x16(tmem_d, va);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        tmem_ld_wait(va);
        tmem_ld_x16(tmem_d + c0 + 16, vb);
        emit(va, c0);
        tmem_ld_wait(vb);
        if (c0 + 32 < BLOCK_N) tmem_ld_x16(tmem_d + c0 + 32, va);
        emit(vb, c0 + 16);
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[acc]));
    }
    if (want_stats) stat_flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// igemm, shared-memory PATCH reuse for the merged 3x3-tap layers (the 4 sub-pixel phases of a k4/s2/p1
// transposed conv, or of the dgrad of a k4/s2/p1 conv, as one GEMM over the 3x3 neighbourhood of the
// low-resolution grid: deconv2/3/4 forward, conv2/3 dgrad).  These layers were bound by the L2 -> shared
// memory operand feed: every input pixel was fetched once per tap (9 boxes per 128-row tile), and 20 of the
// 36 (tap, phase) weight blocks are structurally zero.  Here
//   * ONE activation box per filter COLUMN dx and 64-channel block — (bh + 2) image rows x bw pixels x bn
//     images, zero-filled halo — serves the three taps dy = -1, 0, +1: tile row r = (y * bn + img) * bw + x, so
//     a shift by one image row is a shift by bn * bw rows of the box = a whole number of 8-row swizzle atoms,
//     i.e. just another start address of the UMMA descriptor.  9 boxes -> 3 x (bh + 2) / bh box equivalents;
//   * the weight operand is fetched and multiplied per LIVE quarter of the N dimension only (quarter = one
//     sub-pixel phase): corner taps feed 1 phase, edge taps 2, the centre tap 4 -> 16 instead of 36 blocks, and
//     the MMAs run with N = 1, 2 or 4 quarters into the matching TMEM column range.  The centre tap goes first
//     so that its accumulate = 0 MMA initialises every column.
//   warp 0 TMA producer (ring of activation patches; the live weight blocks stay resident), warp 1 MMA issuer,
//   EG groups of 4 epilogue warps: group e drains the tiles whose accumulator stage is e (EG = 2), so the epilogue
//   (fp16 stores + BatchNorm statistics) has two tile times per tile
// CIN_MODE 0: Cin % 64 == 0 (128-byte rows, 128B swizzle); 1: Cin == 32 (64-byte rows, 64B swizzle).
// ---------------------------------------------------------------------------------------------
// Tap schedule of the patch kernel, fixed at compile time so that the MMA issuer is straight-line code (one
// descriptor add per operand and MMA; table look-ups in the issue loop made it issue-bound: ncu source view,
// 27 uniform-datapath instructions incl. 3 dependent constant loads per UTCHMMA).  Order: filter column
// dx = 0, -1, +1; inside a column dy = 0, -1, +1 — the centre tap first (its accumulate = 0 MMA initialises every
// column).  NQ4 (out_mode 4): quarter q = ph*2 + pw of the N dimension is one sub-pixel phase of a k4/s2/p1
// transposed conv; phase parity 0 reads input offsets {0, -1}, parity 1 reads {+1, 0} along that axis, so a tap
// (dy, dx) feeds the quarters listed here.  blk = index of the run's first weight block in the resident region
// (live blocks are stored back to back in schedule order).
struct TapSched {
  int dy, dx, nruns, q0[2], len[2], blk[2];
};
__host__ __device__ constexpr TapSched tap_sched(bool nq4, int ti) {
  if (nq4) {
    switch (ti) {
      case 0: return {0, 0, 1, {0, 0}, {4, 0}, {0, 0}};
      case 1: return {-1, 0, 1, {0, 0}, {2, 0}, {4, 0}};
      case 2: return {1, 0, 1, {2, 0}, {2, 0}, {6, 0}};
      case 3: return {0, -1, 2, {0, 2}, {1, 1}, {8, 9}};
      case 4: return {-1, -1, 1, {0, 0}, {1, 0}, {10, 0}};
      case 5: return {1, -1, 1, {2, 0}, {1, 0}, {11, 0}};
      case 6: return {0, 1, 2, {1, 3}, {1, 1}, {12, 13}};
      case 7: return {-1, 1, 1, {1, 0}, {1, 0}, {14, 0}};
      default: return {1, 1, 1, {3, 0}, {1, 0}, {15, 0}};
    }
  }
  switch (ti) {
    case 0: return {0, 0, 1, {0, 0}, {1, 0}, {0, 0}};
    case 1: return {-1, 0, 1, {0, 0}, {1, 0}, {1, 0}};
    case 2: return {1, 0, 1, {0, 0}, {1, 0}, {2, 0}};
    case 3: return {0, -1, 1, {0, 0}, {1, 0}, {3, 0}};
    case 4: return {-1, -1, 1, {0, 0}, {1, 0}, {4, 0}};
    case 5: return {1, -1, 1, {0, 0}, {1, 0}, {5, 0}};
    case 6: return {0, 1, 1, {0, 0}, {1, 0}, {6, 0}};
    case 7: return {-1, 1, 1, {0, 0}, {1, 0}, {7, 0}};
    default: return {1, 1, 1, {0, 0}, {1, 0}, {8, 0}};
  }
}

// Phase-split variant (PS): layers whose live weight blocks do not fit in shared memory (deconv2.fwd, conv3.dgrad: 256 KB)
// give each CTA ONE output-row parity ph (CTA index parity) and the two sub-pixel phases (ph, pw = 0 / 1) that go with it:
// half of the live blocks (8 per 64-channel block), 6 taps per filter-column patch — dy = 0 and dy = -1 (ph = 0) or +1
// (ph = 1); dx = 0 feeds both pw quarters, dx = -1 only pw = 0, dx = +1 only pw = 1.
struct TapSchedPS {
  int dx, dy_kind, nruns, q0, len, blk;  // dy_kind 0: dy = 0; 1: dy = ph ? +1 : -1
};
__host__ __device__ constexpr TapSchedPS tap_sched_ps(int ti) {
  switch (ti) {
    case 0: return {0, 0, 1, 0, 2, 0};
    case 1: return {0, 1, 1, 0, 2, 2};
    case 2: return {-1, 0, 1, 0, 1, 4};
    case 3: return {-1, 1, 1, 0, 1, 5};
    case 4: return {1, 0, 1, 1, 1, 6};
    default: return {1, 1, 1, 1, 1, 7};
  }
}

struct PatchGeom {
  int lbw, lbn, bh, bn;     // box: bw = 1 << lbw pixels wide (= IW), bh rows, bn = 1 << lbn images
  int tiles_y, img_blocks, total_tiles;
  int a_rows;               // (bh + 2) * bn * bw rows per activation box
  int nq;                   // live-quarter granularity: 4 (out_mode 4: one quarter per sub-pixel phase) or 1
  int w_bytes_cb;           // bytes of the live weight blocks of one 64-channel block (all 9 taps)
  int8_t tap_k[9];          // [dxi * 3 + dyi] -> tap index in the K order of the packed weights
  int8_t tap_dyv[9];        // dy of that tap (-1, 0, 1)
  int8_t tap_dxv[9];
  int8_t nruns[9];          // contiguous runs of live quarters (<= 2)
  int8_t run_q0[9][2], run_len[9][2];
  int32_t run_off[9][2];    // byte offset of the run's first weight block inside the resident weight region
  uint8_t qmask[9];
  int dbg;                  // ablation switches, honoured only by builds with -DMMDYN_PATCH_DBG (env MMDYN_PATCH_DBG):
                            // 1 = epilogue body skipped, 2 = no MMAs, 4 = no activation loads, 8 = no global stores
};
#ifdef MMDYN_PATCH_DBG
#define PATCH_DBG(g, bit) (((g).dbg & (bit)) != 0)
#else
#define PATCH_DBG(g, bit) false
#endif

template <int BLOCK_N, int CIN_MODE, int SA, int W_KB>
struct PatchCfg {
  static constexpr int RB = CIN_MODE == 0 ? 128 : 64;          // bytes per operand row (one pixel, one k-block)
  static constexpr int KK = RB / 32;                            // K = 16 MMAs per k-block
  static constexpr int A_SLOT = CIN_MODE == 0 ? 20 * 1024 : 12 * 1024;  // 160 rows x 128 B / 192 rows x 64 B
  static constexpr int W_BYTES = W_KB * 1024;                   // resident live weight blocks
  static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int SMEM_BYTES = SA * A_SLOT + W_BYTES + 1024;
};

template <int BLOCK_N, int CIN_MODE, int SA, int W_KB, int EG, bool PS>
__global__ void __launch_bounds__(64 + 128 * EG, BLOCK_N == 16 ? 2 : 1)
igemm_patch_kernel(const __grid_constant__ mmdyn_igemm_desc d, const __grid_constant__ CUtensorMap tmA,
                   const __grid_constant__ CUtensorMap tmW, const __grid_constant__ PatchGeom g) {
  using C = PatchCfg<BLOCK_N, CIN_MODE, SA, W_KB>;
  constexpr int RB = C::RB;
  constexpr uint32_t LAYOUT = CIN_MODE == 0 ? LAYOUT_SW128 : LAYOUT_SW64;
  constexpr int NQ_MAX = 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base, a_ring = smem_base + C::W_BYTES;
  __shared__ __align__(8) uint64_t a_full[SA], a_empty[SA], w_full;
  constexpr int NACC = EG < 2 ? 2 : EG;  // accumulator stages in TMEM: epilogue group e drains stage e
  __shared__ __align__(8) uint64_t tfull_bar[NACC], tempty_bar[NACC];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_all[EG][2 * 64];  // per epilogue group: BatchNorm partial sums {sum x, sum x^2} per channel

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    mbar_init(smem_u32(&w_full), 1);
    for (int a = 0; a < NACC; ++a) { mbar_init(smem_u32(&tfull_bar[a]), 1); mbar_init(smem_u32(&tempty_bar[a]), 128); }
    mbar_fence_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  for (int i = threadIdx.x; i < EG * 128; i += blockDim.x) (&stat_all[0][0])[i] = 0.0f;
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), NACC * C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();  // the prologue above touched no global memory: it overlaps the previous grid's tail

  const int CB = CIN_MODE == 0 ? (d.Cin >> 6) : 1;
  const int bw = 1 << g.lbw;
  const int NQ = PS ? BLOCK_N / 2 : BLOCK_N / g.nq;    // columns per quarter
  const uint32_t dy_shift = static_cast<uint32_t>(g.bn * bw * RB);  // one image row of the box, in bytes
  // PS: this CTA's output-row parity; CTAs 2c and 2c + 1 walk the same tiles c, c + gridDim/2, ...
  const int ph = PS ? static_cast<int>(blockIdx.x & 1) : 0;
  const int tile_first = PS ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_stride = PS ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ======================= TMA producer =====================================================
    if (elect_one()) {
      // the live weight blocks of the whole layer stay resident in shared memory: fetched once per CTA
      if constexpr (PS) {
        const uint32_t wbar = smem_u32(&w_full);
        mbar_arrive_expect_tx(wbar, static_cast<uint32_t>(CB * 8 * NQ * RB));
        for (int cb = 0; cb < CB; ++cb)
#pragma unroll
          for (int ti = 0; ti < 6; ++ti) {
            const TapSchedPS ts = tap_sched_ps(ti);
            const int dy = ts.dy_kind == 0 ? 0 : (ph ? 1 : -1);
            const int k = ((dy + 1) * 3 + (ts.dx + 1)) * d.Cin + (cb << 6);
            for (int q = 0; q < ts.len; ++q)
              tma_load_2d(w_base + (cb * 8 + ts.blk + q) * NQ * RB, &tmW, wbar, k, (ph * 2 + ts.q0 + q) * NQ);
          }
      } else {
        const uint32_t wbar = smem_u32(&w_full);
        mbar_arrive_expect_tx(wbar, static_cast<uint32_t>(CB * g.w_bytes_cb));
        for (int cb = 0; cb < CB; ++cb)
          for (int ti = 0; ti < 9; ++ti) {
            const int k = g.tap_k[ti] * d.Cin + (cb << 6);
            uint32_t dst = w_base + cb * g.w_bytes_cb + g.run_off[ti][0];
            const uint32_t mask = g.qmask[ti];
#pragma unroll
            for (int q = 0; q < NQ_MAX; ++q)
              if (mask & (1u << q)) {
                tma_load_2d(dst, &tmW, wbar, k, q * NQ);
                dst += NQ * RB;
              }
          }
      }
      int ia = 0;
      for (int tile = tile_first; tile < g.total_tiles; tile += tile_stride) {
        const int iblk = tile / g.tiles_y;
        const int img0 = iblk * g.bn, y0 = (tile - iblk * g.tiles_y) * g.bh;
        for (int dxi = 0; dxi < 3; ++dxi) {
          const int dx = dxi == 0 ? 0 : (dxi == 1 ? -1 : 1);
          for (int cb = 0; cb < CB; ++cb) {
            const int sa = ia % SA;
            mbar_wait(smem_u32(&a_empty[sa]), ((ia / SA) & 1) ^ 1);
            const uint32_t abar = smem_u32(&a_full[sa]);
            if (PATCH_DBG(g, 4)) {
              mbar_arrive(abar);
            } else {
              mbar_arrive_expect_tx(abar, static_cast<uint32_t>(g.a_rows * RB));
              tma_load_4d(a_ring + sa * C::A_SLOT, &tmA, abar, cb << 6, dx, img0, y0 - 1);  // dims (c, x, img, y)
            }
            ++ia;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer ======================================================
    // one ELECTED thread (elect.sync: the compiler then knows the region is single-threaded and emits plain
    // uniform-datapath code, no ELECT / BRA.U.ANY waterfall around every UTCHMMA); schedule fully unrolled
    if (elect_one()) {
      constexpr bool NQ4 = BLOCK_N >= 64;
      constexpr int NQC = PS ? BLOCK_N / 2 : (NQ4 ? BLOCK_N / 4 : BLOCK_N);   // columns (= weight rows) per quarter
      // descriptor = DESC_HI | (address >> 4): LBO 16 B (unused with swizzled K-major), SBO = 8 rows, version 1
      constexpr uint64_t DESC_HI = (static_cast<uint64_t>(16 >> 4) << 16) | (static_cast<uint64_t>((8 * RB) >> 4) << 32) |
                                   (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(LAYOUT) << 61);
      mbar_wait(smem_u32(&w_full), 0);
      const uint32_t sh0 = 0, sh1 = dy_shift >> 4, sh2 = (2 * dy_shift) >> 4;  // dy = -1, 0, +1 in 16-byte units
      int ia = 0, tl = 0;
      for (int tile = tile_first; tile < g.total_tiles; tile += tile_stride, ++tl) {
        const int acc = tl % NACC;
        mbar_wait(smem_u32(&tempty_bar[acc]), ((tl / NACC) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS;
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi) {
          for (int cb = 0; cb < CB; ++cb) {
            const int sa = ia % SA;
            mbar_wait(smem_u32(&a_full[sa]), (ia / SA) & 1);
            tc_fence_after();
            const uint32_t a16 = (a_ring + sa * C::A_SLOT) >> 4;
            if constexpr (PS) {
              const uint32_t w16 = (w_base + cb * 8 * NQC * RB) >> 4;
#pragma unroll
              for (int dyi = 0; dyi < 2; ++dyi) {
                const TapSchedPS ts = tap_sched_ps(dxi * 2 + dyi);
                const uint32_t a_tap = a16 + (ts.dy_kind == 0 ? sh1 : (ph ? sh2 : sh0));
#pragma unroll
                for (int kk = 0; kk < C::KK; ++kk) {
                  const uint64_t adesc = DESC_HI | static_cast<uint64_t>(a_tap + 2 * kk);
                  const uint64_t bdesc = DESC_HI | static_cast<uint64_t>(w16 + ((ts.blk * NQC * RB + 32 * kk) >> 4));
                  const uint32_t accumulate = (dxi == 0 && dyi == 0 && kk == 0) ? (cb == 0 ? 0u : 1u) : 1u;
                  umma_f16(tmem_d + ts.q0 * NQC, adesc, bdesc, make_idesc_f16(128, ts.len * NQC, 0, 0, 0, 0), accumulate);
                }
              }
              umma_commit(smem_u32(&a_empty[sa]));
              ++ia;
              continue;
            }
            const uint32_t w16 = (w_base + cb * g.w_bytes_cb) >> 4;
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              if (PATCH_DBG(g, 2)) break;
              const TapSched ts = tap_sched(NQ4, dxi * 3 + dyi);
              const uint32_t a_tap = a16 + (ts.dy < 0 ? sh0 : (ts.dy == 0 ? sh1 : sh2));
#pragma unroll
              for (int kk = 0; kk < C::KK; ++kk) {
                const uint64_t adesc = DESC_HI | static_cast<uint64_t>(a_tap + 2 * kk);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                  if (r < ts.nruns) {
                    const uint64_t bdesc = DESC_HI | static_cast<uint64_t>(w16 + ((ts.blk[r] * NQC * RB + 32 * kk) >> 4));
                    const uint32_t accumulate = (dxi == 0 && dyi == 0 && kk == 0) ? (cb == 0 ? 0u : 1u) : 1u;
                    umma_f16(tmem_d + ts.q0[r] * NQC, adesc, bdesc, make_idesc_f16(128, ts.len[r] * NQC, 0, 0, 0, 0),
                             accumulate);
                  }
                }
              }
            }
            umma_commit(smem_u32(&a_empty[sa]));
            ++ia;
          }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
      }
    }
  } else {
    // ======================= epilogue: TMEM -> registers -> global ============================
    const int q4 = warp & 3;       // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;  // tile row owned by this thread: r = (y * bn + img) * bw + x
    const int eg = EG == 1 ? 0 : (warp - 2) >> 2;  // epilogue group: tiles tl = eg, eg + EG, ...
    const int et = (threadIdx.x - 64) & 127;       // thread index inside the group
    float* stat_s = stat_all[eg];
    const int x_l = r & (bw - 1);
    const int n_l = (r >> g.lbw) & (g.bn - 1);
    const int y_l = r >> (g.lbw + g.lbn);
    float bce_acc = 0.0f;
    int bce_cur = -1;
    auto bce_flush = [&]() {
      const float v = warp_sum(bce_acc);
      if (lane == 0 && bce_cur >= 0) atomicAdd(d.bce_loss + bce_cur, v);
      bce_acc = 0.0f;
    };
    // BatchNorm statistics of the raw output (out_mode 4 with bn_sums): the CTA's partial sums live in shared
    // memory and are flushed to bn_sums[group] when the group changes (tiles are walked in image order)
    const bool want_stats = d.out_mode == 4 && d.bn_sums != nullptr;
    int stat_grp = -1;
    auto stat_flush = [&]() {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
      const int i = et;
      if (stat_grp >= 0 && i < 2 * d.ldc) {
        const float v = stat_s[i];
        if (v != 0.0f) atomicAdd(d.bn_sums + (static_cast<long long>(stat_grp) * d.ldc + (i >> 1)) * 2 + (i & 1), v);
        stat_s[i] = 0.0f;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
    };
    for (int tile = tile_first + eg * tile_stride, tl = eg; tile < g.total_tiles; tile += EG * tile_stride, tl += EG) {
      const int iblk = tile / g.tiles_y;
      const int img = iblk * g.bn + n_l, yv = (tile - iblk * g.tiles_y) * g.bh + y_l, xv = x_l;
      const bool valid = img < d.n_img;
      int out_off;
      if (d.out_mode == 3 || d.out_mode == 5) out_off = valid ? ((img * 3 * d.OH) + 2 * yv) * d.OW + 2 * xv : -1;
      else out_off = valid ? ((img * d.OH + 2 * yv) * d.OW + 2 * xv) * d.ldc : -1;
      if (want_stats) {
        // all images of a tile belong to one group (checked by the launcher: rows_per_group % bn == 0)
        const int grp = (iblk * g.bn) / d.bn_rows_per_group;
        if (grp != stat_grp) {
          stat_flush();
          stat_grp = grp;
        }
      }
      float2 bce_t[6], bce_m[6];
      int bce_slot_now = -1;
      if constexpr (BLOCK_N == 16) {
        if (d.out_mode == 5 && valid) {
          const int grp = img / d.bce_rows_per_group;
          bce_slot_now = d.bce_slot[grp];
          if (bce_slot_now >= 0) {
            const int plane = d.OH * d.OW;
            const int toff = (((img - grp * d.bce_rows_per_group) * 3) * d.OH + 2 * yv) * d.OW + 2 * xv;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ph = 0; ph < 2; ++ph) {
                bce_t[c * 2 + ph] = __ldg(reinterpret_cast<const float2*>(d.bce_target + toff + c * plane + ph * d.OW));
                bce_m[c * 2 + ph] = d.bce_mask ? __ldg(reinterpret_cast<const float2*>(d.bce_mask + toff + c * plane +
                                                                                       ph * d.OW))
                                               : make_float2(1.0f, 1.0f);
              }
          }
        }
      }
      const int acc = tl % NACC;
      mbar_wait(smem_u32(&tfull_bar[acc]), (tl / NACC) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::TMEM_COLS + (static_cast<uint32_t>(q4 * 32) << 16);
      if (PATCH_DBG(g, 1)) out_off = -2;
      if (out_off == -2) {
      } else if constexpr (BLOCK_N == 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_d, v);
        tmem_ld_wait(v);
        if (out_off >= 0) {
          float f[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
          const int plane = d.OH * d.OW;
          if (d.out_mode == 3 || (img >= d.logit_row_lo && img < d.logit_row_hi)) {
            float* o = reinterpret_cast<float*>(d.out) + out_off;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ph = 0; ph < 2; ++ph)
                *reinterpret_cast<float2*>(o + c * plane + ph * d.OW) =
                    make_float2(f[(ph * 2 + 0) * 3 + c], f[(ph * 2 + 1) * 3 + c]);
          }
          const int slot = bce_slot_now;
          if (d.out_mode == 5 && slot >= 0) {
            if (slot != bce_cur) {
              bce_flush();
              bce_cur = slot;
            }
            float gq[4][4];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ph = 0; ph < 2; ++ph) {
                const float2 tv = bce_t[c * 2 + ph], mv = bce_m[c * 2 + ph];
                const float ts[2] = {tv.x, tv.y}, ms[2] = {mv.x, mv.y};
#pragma unroll
                for (int pw = 0; pw < 2; ++pw) {
                  const float m = ms[pw], x = f[(ph * 2 + pw) * 3 + c] * m, tt = ts[pw] * m;
                  const float e = __expf(-fabsf(x));
                  const float rr = __fdividef(1.0f, 1.0f + e);
                  bce_acc += fmaxf(x, 0.0f) - x * tt + __logf(1.0f + e);
                  const float sig = x >= 0.0f ? rr : e * rr;
                  gq[ph * 2 + pw][c] = d.bce_gscale * (sig - tt) * m;
                }
              }
            if (d.bce_dlogits) {
              __half* go = reinterpret_cast<__half*>(d.bce_dlogits);
#pragma unroll
              for (int ph = 0; ph < 2; ++ph) {
                const long long pix =
                    (static_cast<long long>(img) * (d.OH + 2) + 2 * yv + ph + 1) * (d.OW + 2) + 2 * xv + 1;
                uint2* o2 = reinterpret_cast<uint2*>(go + pix * 4);  // NHWC4: 8 bytes per pixel
#pragma unroll
                for (int pw = 0; pw < 2; ++pw) {
                  const float* q = gq[ph * 2 + pw];
                  o2[pw] = make_uint2(pack_h2(q[0], q[1]), pack_h2(q[2], 0.0f));
                }
              }
            }
          }
        }
      } else {
        // out_mode 4: n = phase * ldc + c.  Walk 16-channel chunks; inside a chunk the 4 phases, so that the
        // BatchNorm partial sums of a channel chunk are reduced across the warp once per 4 phases.
        const int Cc = d.ldc;
        for (int c0 = 0; c0 < Cc; c0 += 16) {
          float s1[16], s2[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) s1[q] = s2[q] = 0.0f;
          // one 16-channel chunk of sub-pixel phase ph: convert, store 32 bytes, accumulate the statistics
          auto emit = [&](const uint32_t (&v)[16], const int ph) {
            if (out_off < 0) return;
            float f[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
            __half* o = reinterpret_cast<__half*>(d.out) + out_off + ((ph >> 1) * d.OW + (ph & 1)) * Cc + c0;
            const uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
            const uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]),
                                        pack_h2(f[14], f[15]));
            if (!PATCH_DBG(g, 8)) {
              st_global_32B(o, u0, u1);
            }
            if (want_stats) {
              // statistics of the values the BatchNorm kernels will read: the fp16-ROUNDED outputs
              const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
                s1[2 * q] += a.x;
                s2[2 * q] = fmaf(a.x, a.x, s2[2 * q]);
                s1[2 * q + 1] += a.y;
                s2[2 * q + 1] = fmaf(a.y, a.y, s2[2 * q + 1]);
              }
            }
          };
          uint32_t va[16], vb[16];
          tmem_ld_x16(tmem_d + c0, va);
          tmem_ld_wait(va);
          tmem_ld_x16(tmem_d + Cc + c0, vb);
          if constexpr (PS) {  // this CTA holds the phases (ph, 0) and (ph, 1)
            emit(va, ph * 2);
            tmem_ld_wait(vb);
            emit(vb, ph * 2 + 1);
          } else {
            emit(va, 0);
            tmem_ld_wait(vb);
            tmem_ld_x16(tmem_d + 2 * Cc + c0, va);
            emit(vb, 1);
            tmem_ld_wait(va);
            tmem_ld_x16(tmem_d + 3 * Cc + c0, vb);
            emit(va, 2);
            tmem_ld_wait(vb);
            emit(vb, 3);
          }
          if (want_stats) {
            // 32 values (16 sums, 16 sums of squares) over the warp's 32 rows: butterfly reduce-scatter, 31 shuffles;
            // lane l ends with the total of value l (l < 16: sum of channel c0 + l; l >= 16: squares of c0 + l - 16)
            float w[32];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              w[q] = s1[q];
              w[16 + q] = s2[q];
            }
#pragma unroll
            for (int st = 0; st < 5; ++st) {
              const int off = 16 >> st, cnt = 16 >> st;
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int j = 0; j < cnt; ++j) {
                const float mine = upper ? w[j + cnt] : w[j];
                const float send = upper ? w[j] : w[j + cnt];
                w[j] = mine + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            atomicAdd(&stat_s[(c0 + (lane & 15)) * 2 + (lane >> 4)], w[0]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[acc]));
    }
    if (d.out_mode == 5) bce_flush();
    if (want_stats) stat_flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NACC * C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[n][k] += scale * sum_rows Nat[row][n] * G_gather[row][k]
//   D tile = [128 k-columns (M)] x [CN channels (N)], reduction over rows in steps of 64.
//   Both operands sit in smem as rows-of-channels (the layout they have in HBM), which is the
//   MN-major UMMA canonical layout: no transposes anywhere.
// ---------------------------------------------------------------------------------------------
template <int CN>
__global__ void __launch_bounds__(CTA_THREADS)
wgrad_kernel(const __grid_constant__ mmdyn_wgrad_desc d) {
  using C = Cfg<CN>;
  constexpr int STEP_ROWS = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kcol0 = blockIdx.y * 128;
  const int n0 = blockIdx.z * CN;

  if (threadIdx.x == NUM_PRODUCER_THREADS) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), NUM_PRODUCER_THREADS);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();  // the prologue above touched no global memory: it overlaps the previous grid's tail

  const int M = d.n_img * d.P;
  const int steps_total = (M + STEP_ROWS - 1) / STEP_ROWS;
  const int steps_per = (steps_total + d.row_splits - 1) / d.row_splits;
  const int step_begin = blockIdx.x * steps_per;
  const int step_end = min(steps_total, step_begin + steps_per);
  const int n_steps = max(0, step_end - step_begin);

  if (warp < 4) {
    const __half* G = reinterpret_cast<const __half*>(d.G);
    const __half* Nat = reinterpret_cast<const __half*>(d.Nat);
    const int j = threadIdx.x & 7;
    const int rsub = threadIdx.x >> 3;
    const uint32_t sw = static_cast<uint32_t>((j ^ (rsub & 7)) << 4);
    // (tap, channel) of this thread's chunk in each of the two 64-wide k-column blocks
    int tap_off[2], tdy[2], tdx[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int k = kcol0 + b * 64 + j * 8;
      const int tap = k / d.Cg;
      const int c = k - tap * d.Cg;
      tdy[b] = d.tap_dy[tap];
      tdx[b] = d.tap_dx[tap];
      tap_off[b] = (tdy[b] * d.IW + tdx[b]) * d.g_pix_stride + c;
    }
    for (int it = 0; it < n_steps; ++it) {
      const int s = it % C::STAGES;
      mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
      const int m_base = (step_begin + it) * STEP_ROWS;
      const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
      const uint32_t b_stage = a_stage + A_STAGE_BYTES;
      // gathered operand: 2 blocks x [64 rows x 128 B]
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rsub + 16 * i;
        const int m = m_base + r;
        const bool rv = m < M;
        const int img = m / d.P;
        const int p = m - img * d.P;
        const int yv = p / d.OXv, xv = p - yv * d.OXv;
        const int iy0 = yv * d.s_in, ix0 = xv * d.s_in;
        const int a_off = ((img * d.IH + iy0) * d.IW + ix0) * d.g_pix_stride;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const bool v = rv && (unsigned)(iy0 + tdy[b]) < (unsigned)d.IH &&
                         (unsigned)(ix0 + tdx[b]) < (unsigned)d.IW;
          const __half* src = v ? (G + (a_off + tap_off[b])) : G;
          cp_async16(a_stage + b * 8192 + r * 128 + sw, src, v);
        }
      }
      // natural operand: [64 rows x CN channels]
      constexpr int CHUNKS_PER_ROW = CN / 8;
#pragma unroll
      for (int q = 0; q < (STEP_ROWS * CHUNKS_PER_ROW + 127) / 128; ++q) {
        const int idx = q * 128 + threadIdx.x;
        if (idx < STEP_ROWS * CHUNKS_PER_ROW) {
          const int row = idx / CHUNKS_PER_ROW;
          const int cj = idx - row * CHUNKS_PER_ROW;
          const int m = m_base + row;
          const bool v = m < M;
          const __half* src = v ? (Nat + (static_cast<long long>(m) * d.nat_stride + n0 + cj * 8)) : Nat;
          uint32_t off;
          if (CN >= 64) {
            off = (cj >> 3) * 8192 + row * 128 + (((cj & 7) ^ (row & 7)) << 4);
          } else if (CN == 32) {
            off = swz<2>(row * 64 + cj * 16);
          } else {
            off = swz<1>(row * 32 + cj * 16);
          }
          cp_async16(b_stage + off, src, v);
        }
      }
      cp_async_commit();
      if (it >= C::LAG) {
        cp_async_wait<C::LAG>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&full_bar[(it - C::LAG) % C::STAGES]));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (int q = (n_steps > C::LAG ? n_steps - C::LAG : 0); q < n_steps; ++q)
      mbar_arrive(smem_u32(&full_bar[q % C::STAGES]));
  } else if (warp == 4) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, CN, 0, 0, 1, 1);
      constexpr uint32_t b_layout = CN >= 64 ? LAYOUT_SW128 : (CN == 32 ? LAYOUT_SW64 : LAYOUT_SW32);
      constexpr uint32_t b_sbo = CN >= 64 ? 1024 : (CN == 32 ? 512 : 256);  // 8 rows of the tile
      for (int it = 0; it < n_steps; ++it) {
        const int s = it % C::STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (it / C::STAGES) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_base + s * C::STAGE_BYTES;
        const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 16 reduction rows per MMA
          const uint64_t adesc = make_smem_desc(a_base + kk * 2048, 8192, 1024, LAYOUT_SW128);
          const uint64_t bdesc = make_smem_desc(b_base + kk * 2 * b_sbo, 8192, b_sbo, b_layout);
          umma_f16(tmem_base, adesc, bdesc, idesc, (it | kk) != 0);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(&accum_bar));
    }
  }

  if (warp < 4 && n_steps > 0) {
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    float* o = d.dW + static_cast<long long>(n0) * d.ldw + kcol0 + threadIdx.x;
#pragma unroll 1
    for (int c0 = 0; c0 < CN; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait(v);
#pragma unroll
      for (int q = 0; q < 16; ++q)
        atomicAdd(o + static_cast<long long>(c0 + q) * d.ldw, d.scale * __uint_as_float(v[q]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad, TMA-fed: both operands of dW += Nat^T * gather(G) arrive as 4-D TMA boxes of 64 reduction
// rows (bw x bh virtual pixels x bn images, or one pixel x 64 images); they land in shared memory
// as rows-of-channels, which IS the MN-major UMMA layout, so nothing is transposed or touched by
// the LSU.  G_MODE: 0 = Cg % 64 == 0 (2 boxes of 64 k-columns, 128B swizzle), 1 = Cg == 32 (4 taps,
// 64B swizzle), 2 = Cg == 8 (16 taps, un-swizzled 16-byte rows), 3 = Cg == 16 with ntaps * Cg = 64: the 4-channel
// logit gradient (or the padded 4-channel input image) through 4-pixel windows — per row tap the raw rows are
// fetched twice as 512-byte runs (pixel pairs 0..31 and 1..32), which land as the two 16-byte column chunks of
// that tap in G_MODE 2's layout: 16 TMA row requests per step instead of 256.  The upper 64 accumulator rows
// multiply whatever the stage held before and are never read.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (fp32 atomics into dW)
// ---------------------------------------------------------------------------------------------
struct WgradGeomDev {
  int lbw, lbh, bn, pixel_major, tiles_y, img_blocks, total_steps;
};

// NB = 128-column blocks of dW per CTA.  These kernels run at the L2 -> SM throughput cap (~43 B/clk/SM: every
// 64-row step moves NB * 16 KB of gathered gradient + CN * 128 B of the natural operand for NB * 128 * CN * 64
// MACs), so NB = 2 shares each natural-operand stage between two accumulators: 25 % fewer L2 bytes per MAC at
// CN = 128.  Stage count: what fits two CTAs per SM (>= 2), TMEM = NB * CN columns.
template <int CN, int NB>
struct WCfg {
  static constexpr int STAGE_BYTES = NB * A_STAGE_BYTES + CN * 128;
  static constexpr int TMEM_COLS = NB * CN < 32 ? 32 : NB * CN;
  static constexpr int FIT = (TMEM_COLS > 256 ? 200 * 1024 : 110 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = FIT < 2 ? 2 : (FIT > 4 ? 4 : FIT);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

template <int CN, int G_MODE, int NB>
__global__ void __launch_bounds__(TMA_THREADS)
wgrad_tma_kernel(const __grid_constant__ mmdyn_wgrad_desc d, const __grid_constant__ CUtensorMap tmG,
                 const __grid_constant__ CUtensorMap tmN, const WgradGeomDev g) {
  using C = WCfg<CN, NB>;
  static_assert(NB == 1 || G_MODE == 0 || G_MODE == 1, "two column blocks per CTA: G_MODE 0 / 1 only");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kcol0 = blockIdx.y * (128 * NB);
  const int n0 = blockIdx.z * CN;
  constexpr int G_BYTES = NB * A_STAGE_BYTES;  // gathered operand of one stage; the natural operand follows it

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmN);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_sync();  // the prologue above touched no global memory: it overlaps the previous grid's tail

  const int steps_per = (g.total_steps + d.row_splits - 1) / d.row_splits;
  const int step_begin = blockIdx.x * steps_per;
  const int step_end = min(g.total_steps, step_begin + steps_per);
  const int n_steps = max(0, step_end - step_begin);
  constexpr int NAT_BYTES = 64 * CN * 2;

  if (warp == 0) {
    if (elect_one()) {
      // operand-box coordinates that do not depend on the step: taps / channel offsets of this CTA's k-columns
      int tdx[16], tdy[16], tc0[2 * NB];
#pragma unroll
      for (int b = 0; b < 16; ++b) tdx[b] = tdy[b] = 0;
      if (G_MODE == 0) {
#pragma unroll
        for (int b = 0; b < 2 * NB; ++b) {
          const int k = kcol0 + b * 64;
          const int tap = k / d.Cg;
          tc0[b] = k - tap * d.Cg;
          tdx[b] = d.tap_dx[tap];
          tdy[b] = d.tap_dy[tap];
        }
      } else if (G_MODE == 1) {
#pragma unroll
        for (int b = 0; b < 4 * NB; ++b) {
          tdx[b] = d.tap_dx[(kcol0 >> 5) + b];
          tdy[b] = d.tap_dy[(kcol0 >> 5) + b];
        }
      } else if (G_MODE == 3) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          tdx[b] = d.tap_dx[b];
          tdy[b] = d.tap_dy[b];
        }
      } else {
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          tdx[b] = d.tap_dx[(kcol0 >> 3) + b];
          tdy[b] = d.tap_dy[(kcol0 >> 3) + b];
        }
      }
      // step -> (outer, inner) without a division per step: pixel-major (pixel, image block), else (image block, row tile)
      const int inner_n = g.pixel_major ? g.img_blocks : g.tiles_y;
      int outer = step_begin / inner_n, inner = step_begin - outer * inner_n;
      int py = 0, px = 0;
      if (g.pixel_major) {
        py = outer / d.OXv;
        px = outer - py * d.OXv;
      }
      for (int it = 0; it < n_steps; ++it) {
        const int s = it % C::STAGES;
        mbar_wait(smem_u32(&empty_bar[s]), ((it / C::STAGES) & 1) ^ 1);
        int img0, y0, x0;
        if (g.pixel_major) {
          img0 = inner * 64;
          y0 = py;
          x0 = px;
        } else {
          img0 = outer * g.bn;
          y0 = inner << g.lbh;
          x0 = 0;
        }
        if (++inner == inner_n) {
          inner = 0;
          ++outer;
          if (g.pixel_major && ++px == d.OXv) {
            px = 0;
            ++py;
          }
        }
        const uint32_t bar = smem_u32(&full_bar[s]);
        const uint32_t a_stage = smem_base + s * C::STAGE_BYTES;
        const uint32_t b_stage = a_stage + G_BYTES;
        mbar_arrive_expect_tx(bar, (G_MODE == 3 ? 8 * 1024 : G_BYTES) + NAT_BYTES);
        const int wx = x0 * (d.s_in_x ? d.s_in_x : d.s_in), wy = y0 * d.s_in;
        if (G_MODE == 3) {
#pragma unroll
          for (int b = 0; b < 8; ++b) tma_load_4d(a_stage + b * 1024, &tmG, bar, (b & 1) * 8, 0, wy + tdy[b >> 1], img0);
        } else if (G_MODE == 0) {
#pragma unroll
          for (int b = 0; b < 2 * NB; ++b) tma_load_4d(a_stage + b * 8192, &tmG, bar, tc0[b], wx + tdx[b], wy + tdy[b], img0);
        } else if (G_MODE == 1) {
#pragma unroll
          for (int b = 0; b < 4 * NB; ++b) tma_load_4d(a_stage + b * 4096, &tmG, bar, 0, wx + tdx[b], wy + tdy[b], img0);
        } else {
#pragma unroll
          for (int b = 0; b < 16; ++b) tma_load_4d(a_stage + b * 1024, &tmG, bar, 0, wx + tdx[b], wy + tdy[b], img0);
        }
        if (CN >= 64) {
#pragma unroll
          for (int b = 0; b < CN / 64; ++b) tma_load_4d(b_stage + b * 8192, &tmN, bar, n0 + b * 64, x0, y0, img0);
        } else {
          tma_load_4d(b_stage, &tmN, bar, n0, x0, y0, img0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(128, CN, 0, 0, 1, 1);
      constexpr uint32_t b_layout = CN >= 64 ? LAYOUT_SW128 : (CN == 32 ? LAYOUT_SW64 : LAYOUT_SW32);
      constexpr uint32_t b_sbo = CN >= 64 ? 1024 : (CN == 32 ? 512 : 256);  // 8 reduction rows of the tile
      for (int it = 0; it < n_steps; ++it) {
        const int s = it % C::STAGES;
        mbar_wait(smem_u32(&full_bar[s]), (it / C::STAGES) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_base + s * C::STAGE_BYTES;
        const uint32_t b_base = a_base + G_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 16 reduction rows per MMA
          const uint64_t bdesc = make_smem_desc(b_base + kk * 2 * b_sbo, 8192, b_sbo, b_layout);
#pragma unroll
          for (int h = 0; h < NB; ++h) {  // column blocks: same natural-operand slice, own accumulator
            const uint32_t a_h = a_base + h * A_STAGE_BYTES;
            uint64_t adesc;
            if (G_MODE == 0) adesc = make_smem_desc(a_h + kk * 2048, 8192, 1024, LAYOUT_SW128);
            else if (G_MODE == 1) adesc = make_smem_desc(a_h + kk * 1024, 4096, 512, LAYOUT_SW64);
            else adesc = make_smem_desc(a_h + kk * 256, 128, 1024, 0);
            umma_f16(tmem_base + h * CN, adesc, bdesc, idesc, (it | kk) != 0);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(&accum_bar));
    }
  } else if (n_steps > 0) {
    const int q4 = warp & 3;
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const int c_end = (G_MODE == 3 && q4 >= 2) ? 0 : CN;  // G_MODE 3: 64 live k-columns
#pragma unroll 1
    for (int h = 0; h < NB; ++h) {
      float* o = d.dW + static_cast<long long>(n0) * d.ldw + kcol0 + h * 128 + q4 * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < c_end; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + h * CN + c0, v);
        tmem_ld_wait(v);
#pragma unroll
        for (int q = 0; q < 16; ++q)
          atomicAdd(o + static_cast<long long>(c0 + q) * d.ldw, d.scale * __uint_as_float(v[q]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// conv1: Conv2d(3, 32, k4, s2, p1) on the fp32 NCHW input, 64x64 -> 32x32, fp16 NHWC out.
//   One tile = 4 output rows x 32 columns of one image (128 GEMM rows), K = 48 (+16 zero pad),
//   N = 32: a single 64-wide k-block, so no ring: gather -> 4 MMAs -> epilogue.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv1_fwd_kernel(const float* __restrict__ x, const __half* __restrict__ Wp, __half* __restrict__ out,
                 __half* __restrict__ act, int n_img) {
  pdl_sync();
  __shared__ __align__(1024) uint8_t a_tile[TILE_M * 128];
  __shared__ __align__(1024) uint8_t b_tile[32 * 128];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int r = threadIdx.x;
  const int tile = blockIdx.x;  // n_img * 8 tiles
  const int img = tile >> 3;
  const int oy = ((tile & 7) << 2) + (r >> 5);
  const int ox = r & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&accum_bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 32);
    tmem_relinquish();
  }
  // weights: 32 rows x 128 B -> swizzled K-major tile (256 16-byte chunks, 2 per thread)
  for (int idx = threadIdx.x; idx < 256; idx += 128) {
    const int n = idx >> 3, cj = idx & 7;
    const uint4 w = reinterpret_cast<const uint4*>(Wp)[idx];
    *reinterpret_cast<uint4*>(b_tile + n * 128 + ((cj ^ (n & 7)) << 4)) = w;
  }
  // patch gather: k = (ci, kh, kw); chunk cj holds (ci = cj/2, kh = 2*(cj&1) + {0,1}, kw = 0..3)
  const float* xi = x + static_cast<long long>(img) * 3 * 64 * 64;
#pragma unroll
  for (int cj = 0; cj < 8; ++cj) {
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = 0.0f;
    if (cj < 6) {
      const int ci = cj >> 1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int iy = 2 * oy - 1 + 2 * (cj & 1) + h;
        if ((unsigned)iy < 64u) {
          const float* row = xi + (ci * 64 + iy) * 64;
#pragma unroll
          for (int kw = 0; kw < 4; ++kw) {
            const int ix = 2 * ox - 1 + kw;
            if ((unsigned)ix < 64u) f[h * 4 + kw] = __ldg(row + ix);
          }
        }
      }
    }
    const uint4 u = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]),
                               pack_h2(f[6], f[7]));
    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((cj ^ (r & 7)) << 4)) = u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, 32, 0, 0, 0, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(a_tile), 16, 1024, LAYOUT_SW128);
    const uint64_t bdesc = make_smem_desc(smem_u32(b_tile), 16, 1024, LAYOUT_SW128);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
    umma_commit(smem_u32(&accum_bar));
  }
  mbar_wait(smem_u32(&accum_bar), 0);
  tc_fence_after();
  __half* o = out + ((static_cast<long long>(img) * 32 + oy) * 32 + ox) * 32;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait(v);
    float f[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) f[q] = __uint_as_float(v[q]);
    const uint4 u0 = make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7]));
    const uint4 u1 = make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]), pack_h2(f[14], f[15]));
    st_global_32B(o + c0, u0, u1);
    if (act != nullptr) {
      // Swish of the fp16-ROUNDED conv output (what the stand-alone mmdyn_bn_swish_fwd pass would read back): vae.py:199
      const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      uint32_t a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&uu[q]));
        a[q] = pack_h2(swishf_(v.x), swishf_(v.y));
      }
      st_global_32B(act + (o - out) + c0, make_uint4(a[0], a[1], a[2], a[3]), make_uint4(a[4], a[5], a[6], a[7]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// conv1 weight gradient on CUDA cores: dW[co][k] += scale * sum_pix dRaw[pix][co] * patch[pix][k].
// 1536 outputs, HBM-bound (reads the fp32 input once and the fp16 dRaw once); 256 threads, each
// owns 6 (co, k) outputs; pixels are staged through smem 64 at a time.
__global__ void __launch_bounds__(256)
conv1_wgrad_kernel(const float* __restrict__ x, const __half* __restrict__ dRaw, float* __restrict__ dW,
                   int n_img, float scale, int pix_per_cta) {
  pdl_sync();
  constexpr int PIX = 64;
  __shared__ float patch[PIX][49];
  __shared__ float dy[PIX][33];
  const int t = threadIdx.x;
  const long long total_pix = static_cast<long long>(n_img) * 1024;
  const long long p_begin = static_cast<long long>(blockIdx.x) * pix_per_cta;
  const long long p_end = min(total_pix, p_begin + pix_per_cta);
  // outputs owned: co = t & 31, k = (t >> 5) * 6 + q, q < 6  (8 groups x 6 = 48)
  const int co = t & 31, kg = (t >> 5) * 6;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  for (long long p0 = p_begin; p0 < p_end; p0 += PIX) {
    // stage patches: 64 pixels x 48 taps
    for (int idx = t; idx < PIX * 48; idx += 256) {
      const int pp = idx / 48, k = idx - pp * 48;
      const long long pix = p0 + pp;
      float v = 0.0f;
      if (pix < p_end) {
        const int img = static_cast<int>(pix >> 10);
        const int oy = static_cast<int>(pix >> 5) & 31, ox = static_cast<int>(pix) & 31;
        const int ci = k >> 4, kh = (k >> 2) & 3, kw = k & 3;
        const int iy = 2 * oy - 1 + kh, ix = 2 * ox - 1 + kw;
        if ((unsigned)iy < 64u && (unsigned)ix < 64u)
          v = __ldg(x + ((static_cast<long long>(img) * 3 + ci) * 64 + iy) * 64 + ix);
      }
      patch[pp][k] = v;
    }
    for (int idx = t; idx < PIX * 32; idx += 256) {
      const int pp = idx >> 5, c = idx & 31;
      const long long pix = p0 + pp;
      dy[pp][c] = pix < p_end ? __half2float(dRaw[pix * 32 + c]) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < PIX; ++pp) {
      const float g = dy[pp][co];
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[q] = fmaf(g, patch[pp][kg + q], acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) atomicAdd(dW + co * 48 + kg + q, scale * acc[q]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int g_sm_count = 148;
int g_igemm_occ[5] = {1, 1, 1, 1, 1};  // resident CTAs per SM of igemm_kernel<16,32,64,128,256>
int g_tma_occ[5] = {1, 1, 1, 1, 1};    // same for igemm_tma_kernel

template <int BLOCK_N, int A_MODE>
int launch_igemm_tma(const mmdyn_igemm_desc* d, const CUtensorMap& tmA, const CUtensorMap& tmW, const TileGeom& g,
                     int occ, cudaStream_t st) {
  int grid = g_sm_count * occ;
  if (grid > g.total_tiles) grid = g.total_tiles;
  constexpr int EG = BLOCK_N == 256 ? 2 : 1;
  MMDYN_LAUNCH((igemm_tma_kernel<BLOCK_N, A_MODE, EG>), grid, 64 + 128 * EG, Cfg<BLOCK_N>::SMEM_BYTES, st, *d, tmA, tmW, g);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

template <int BLOCK_N, int A_MODE>
int launch_igemm_pair(const mmdyn_igemm_desc* d, const CUtensorMap& tmA, const CUtensorMap& tmW, const TileGeom& g,
                      cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_pair_kernel<BLOCK_N, A_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg2<BLOCK_N>::SMEM_BYTES));
    configured = true;
  }
  int grid = g_sm_count;
  if (grid > g.total_tiles / 2) grid = g.total_tiles / 2;
  constexpr int smem_bytes = Cfg2<BLOCK_N>::SMEM_BYTES;
  MMDYN_LAUNCH((igemm_pair_kernel<BLOCK_N, A_MODE>), grid, 320, smem_bytes, st, *d, tmA, tmW, g);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

template <int BLOCK_N>
int dispatch_amode(int a_mode, const mmdyn_igemm_desc* d, const CUtensorMap& tmA, const CUtensorMap& tmW,
                   const TileGeom& g, int occ, cudaStream_t st) {
  if (a_mode == 0) return launch_igemm_tma<BLOCK_N, 0>(d, tmA, tmW, g, occ, st);
  if (a_mode == 1) return launch_igemm_tma<BLOCK_N, 1>(d, tmA, tmW, g, occ, st);
  if (a_mode == 3) {
    if constexpr (BLOCK_N == 32) return launch_igemm_tma<32, 3>(d, tmA, tmW, g, occ, st);
    MMDYN_REQUIRE(false, "igemm: Cin = 16 is built for N = 32 only (N=%d)", d->N);
  }
  return launch_igemm_tma<BLOCK_N, 2>(d, tmA, tmW, g, occ, st);
}

template <int CN, int G_MODE, int NB>
int launch_wgrad_tma(const mmdyn_wgrad_desc* d, const CUtensorMap& tmG, const CUtensorMap& tmN,
                     const WgradGeomDev& g, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    MMDYN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel<CN, G_MODE, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          WCfg<CN, NB>::SMEM_BYTES));
    configured = true;
  }
  if (NB == 2) {  // half as many column blocks: split the reduction twice as finely to keep the CTA count
    grid.y /= 2;
    grid.x = static_cast<unsigned>(std::min<long long>(2LL * grid.x, g.total_steps));
  }
  mmdyn_wgrad_desc dd = *d;
  dd.row_splits = static_cast<int>(grid.x);
  constexpr int smem_bytes = WCfg<CN, NB>::SMEM_BYTES;
  MMDYN_LAUNCH((wgrad_tma_kernel<CN, G_MODE, NB>), grid, TMA_THREADS, smem_bytes, st, dd, tmG, tmN, g);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

template <int CN>
int dispatch_gmode(int g_mode, const mmdyn_wgrad_desc* d, const CUtensorMap& tmG, const CUtensorMap& tmN,
                   const WgradGeomDev& g, dim3 grid, cudaStream_t st) {
  static const bool nb1 = getenv("MMDYN_WGRAD_NB1") != nullptr;
  // measured (tools/bench_layers.py): +10 % on the 5x5 <-> 8x8 layer at 4096 rows, neutral at Cn = 64, a loss when a CTA
  // has only ~10 reduction steps to amortise the second accumulator's epilogue
  const bool two = !nb1 && (d->ntaps * d->Cg) % 256 == 0 && CN >= 128 && g.total_steps >= 48LL * grid.x;
  if constexpr (CN >= 128) {
    if (two && g_mode == 0) return launch_wgrad_tma<CN, 0, 2>(d, tmG, tmN, g, grid, st);
    if (two && g_mode == 1) return launch_wgrad_tma<CN, 1, 2>(d, tmG, tmN, g, grid, st);
  }
  if (g_mode == 0) return launch_wgrad_tma<CN, 0, 1>(d, tmG, tmN, g, grid, st);
  if (g_mode == 1) return launch_wgrad_tma<CN, 1, 1>(d, tmG, tmN, g, grid, st);
  if (g_mode == 3) {
    if constexpr (CN == 32) return launch_wgrad_tma<32, 3, 1>(d, tmG, tmN, g, grid, st);
    MMDYN_REQUIRE(false, "wgrad: Cg = 16 is built for Cn = 32 only (Cn=%d)", d->Cn);
  }
  return launch_wgrad_tma<CN, 2, 1>(d, tmG, tmN, g, grid, st);
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

template <int BLOCK_N, int CIN_MODE, int SA, int W_KB, int EG, bool PS = false>
int launch_patch(const mmdyn_igemm_desc* d, const CUtensorMap& tmA, const CUtensorMap& tmW, const PatchGeom& g,
                 cudaStream_t st) {
  using C = PatchCfg<BLOCK_N, CIN_MODE, SA, W_KB>;
  const int w_need = (d->Cin >= 64 ? d->Cin / 64 : 1) * (PS ? 8 * (BLOCK_N / 2) * C::RB : g.w_bytes_cb);
  MMDYN_REQUIRE(w_need <= C::W_BYTES, "igemm patch_mode: %d bytes of live weights do not fit the resident region (%d)",
                w_need, C::W_BYTES);
  static bool configured = false;
  if (!configured) {
    MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_patch_kernel<BLOCK_N, CIN_MODE, SA, W_KB, EG, PS>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  int occ = (228 * 1024) / (C::SMEM_BYTES + 1024 + 1024);
  constexpr int NACC = EG < 2 ? 2 : EG;
  if (occ * NACC * C::TMEM_COLS > 512) occ = 512 / (NACC * C::TMEM_COLS);
  if (occ > 6) occ = 6;
  if (occ < 1) occ = 1;
  int grid = g_sm_count * occ;
  if (PS) {  // two CTAs (output-row parities) per tile
    grid &= ~1;
    if (grid > 2 * g.total_tiles) grid = 2 * g.total_tiles;
  } else if (grid > g.total_tiles) {
    grid = g.total_tiles;
  }
  MMDYN_LAUNCH((igemm_patch_kernel<BLOCK_N, CIN_MODE, SA, W_KB, EG, PS>), grid, 64 + 128 * EG, C::SMEM_BYTES, st, *d, tmA, tmW, g);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

// epilogue groups (= TMEM accumulator stages) of the one-CTA-per-SM patch variants: a tile's accumulator is handed
// MMA -> epilogue -> MMA through two mbarrier round trips, so its stage is busy for (round trips + epilogue time);
// with 4 stages / 4 groups of 4 warps that latency is spread over 4 tiles in flight instead of 2
constexpr int PEG = 4;

// patch_mode launch: geometry checks, the two tensor maps (activations with dims (c, x, img, y); weights with a
// box of one N-quarter), the tap schedule (centre tap first) and the live-quarter runs of every tap
int igemm_patch(const mmdyn_igemm_desc* d, cudaStream_t st) {
  const int IW = d->IW, IH = d->IH;
  MMDYN_REQUIRE(d->s_in == 1 && d->n_phases == 1 && d->ksplit == 1 && d->row_mode == 0 && d->ntaps >= 9 &&
                    d->OXv == IW && d->P == IH * IW && (IW == 8 || IW == 16 || IW == 32) && d->N == d->block_n &&
                    (d->out_mode == 3 || d->out_mode == 4 || d->out_mode == 5) && d->bias == nullptr &&
                    d->a_row_stride == 0 && d->a_img_stride == 0 && d->a_pix_stride == d->Cin,
                "igemm patch_mode: unsupported geometry (s_in=%d ntaps=%d OXv=%d IW=%d IH=%d N=%d out_mode=%d)", d->s_in,
                d->ntaps, d->OXv, IW, IH, d->N, d->out_mode);
  for (int t = 0; t < 9; ++t)
    MMDYN_REQUIRE(d->tap_dy[0][t] == t / 3 - 1 && d->tap_dx[0][t] == t % 3 - 1,
                  "igemm patch_mode: taps 0..8 must be the 3x3 neighbourhood in row-major order");
  const int cin_mode = (d->Cin % 64 == 0) ? 0 : (d->Cin == 32 ? 1 : -1);
  MMDYN_REQUIRE(cin_mode >= 0, "igemm patch_mode: Cin=%d (multiple of 64, or 32)", d->Cin);
  PatchGeom g = {};
  static const int dbg = getenv("MMDYN_PATCH_DBG") ? atoi(getenv("MMDYN_PATCH_DBG")) : 0;
  g.dbg = dbg;
  const int bw = IW;
  g.bh = 128 / bw < IH ? 128 / bw : IH;
  g.bn = 128 / (bw * g.bh);
  MMDYN_REQUIRE(g.bn >= 1 && bw * g.bh * g.bn == 128 && IH % g.bh == 0 && (g.bn & (g.bn - 1)) == 0,
                "igemm patch_mode: tile %d x %d x %d", bw, g.bh, g.bn);
  g.lbw = ilog2(bw);
  g.lbn = ilog2(g.bn);
  g.tiles_y = IH / g.bh;
  g.img_blocks = (d->n_img + g.bn - 1) / g.bn;
  g.total_tiles = g.img_blocks * g.tiles_y;
  g.a_rows = (g.bh + 2) * g.bn * bw;
  const int rb = cin_mode == 0 ? 128 : 64;
  MMDYN_REQUIRE(g.a_rows * rb <= (cin_mode == 0 ? 20 * 1024 : 12 * 1024), "igemm patch_mode: activation box of %d rows", g.a_rows);
  g.nq = d->out_mode == 4 ? 4 : 1;
  MMDYN_REQUIRE(d->N != 256 || (d->Cin == 128 && d->out_mode == 4), "igemm patch_mode: N = 256 is the phase-split layer (Cin = 128)");
  MMDYN_REQUIRE(d->out_mode != 4 || (d->N == 4 * d->ldc && d->ldc % 16 == 0 && d->N >= 64),
                "igemm patch_mode: out_mode 4 needs N = 4*ldc >= 64");
  MMDYN_REQUIRE(d->out_mode == 4 || d->N == 16, "igemm patch_mode: out_mode 3 / 5 need N = 16");
  if (d->bn_sums) {
    MMDYN_REQUIRE(d->out_mode == 4 && d->ldc <= 64 && d->bn_rows_per_group > 0 && d->bn_rows_per_group % g.bn == 0,
                  "igemm patch_mode: bn_sums needs out_mode 4, <= 64 channels and rows_per_group %% %d == 0", g.bn);
  }
  if (d->out_mode == 5) {
    MMDYN_REQUIRE(d->bce_target && d->bce_loss && d->bce_rows_per_group > 0 &&
                      d->n_img <= MMDYN_MAX_GROUPS * d->bce_rows_per_group && g.bn == 1,
                  "igemm patch_mode: out_mode 5 arguments");
  }
  // tap schedule (tap_sched: shared with the kernel's unrolled MMA loop)
  const int nq_bytes = (d->N / g.nq) * rb;  // one weight block: N / nq rows of one k-block
  int w_off = 0;
  for (int ti = 0; ti < 9; ++ti) {
    const TapSched ts = tap_sched(g.nq == 4, ti);
    g.tap_k[ti] = static_cast<int8_t>((ts.dy + 1) * 3 + (ts.dx + 1));
    g.tap_dyv[ti] = static_cast<int8_t>(ts.dy);
    g.tap_dxv[ti] = static_cast<int8_t>(ts.dx);
    g.nruns[ti] = static_cast<int8_t>(ts.nruns);
    unsigned mask = 0;
    for (int r = 0; r < ts.nruns; ++r) {
      g.run_q0[ti][r] = static_cast<int8_t>(ts.q0[r]);
      g.run_len[ti][r] = static_cast<int8_t>(ts.len[r]);
      g.run_off[ti][r] = ts.blk[r] * nq_bytes;
      for (int q = 0; q < ts.len[r]; ++q) mask |= 1u << (ts.q0[r] + q);
      w_off += ts.len[r] * nq_bytes;
    }
    g.qmask[ti] = static_cast<uint8_t>(mask);
  }
  g.w_bytes_cb = w_off;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("igemm: cuTensorMapEncodeTiled driver entry point not available");
    return MMDYN_ERR_CUDA;
  }
  CUtensorMap tmA, tmW;
  {
    const cuuint64_t dim[4] = {static_cast<cuuint64_t>(d->Cin), static_cast<cuuint64_t>(IW),
                               static_cast<cuuint64_t>(d->n_img), static_cast<cuuint64_t>(IH)};
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->Cin) * 2;
    const cuuint64_t str[3] = {pix_b, pix_b * IW * IH, pix_b * IW};  // x, image, y
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(cin_mode == 0 ? 64 : 32), static_cast<cuuint32_t>(bw),
                               static_cast<cuuint32_t>(g.bn), static_cast<cuuint32_t>(g.bh + 2)};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->A), dim, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           cin_mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("igemm patch_mode: cuTensorMapEncodeTiled(A) failed with CUresult %d", static_cast<int>(r));
      return MMDYN_ERR_CUDA;
    }
  }
  {
    const int ktot = d->ntaps * d->Cin;
    const cuuint64_t dim[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(d->N)};
    const cuuint64_t str[1] = {static_cast<cuuint64_t>(ktot) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(cin_mode == 0 ? 64 : 32), static_cast<cuuint32_t>(d->N / g.nq)};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->W), dim, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           cin_mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("igemm patch_mode: cuTensorMapEncodeTiled(W) failed with CUresult %d", static_cast<int>(r));
      return MMDYN_ERR_CUDA;
    }
  }
  if (cin_mode == 1) {
    MMDYN_REQUIRE(d->N == 16, "igemm patch_mode: Cin = 32 is the logits layer (N = 16)");
    return launch_patch<16, 1, 6, 9, 2>(d, tmA, tmW, g, st);   // 9 KB of weights, 6 x 12 KB patches, 2 CTAs per SM x 8 epilogue warps
  }
  switch (d->N) {
    case 64: return launch_patch<64, 0, 6, 64, PEG>(d, tmA, tmW, g, st);
    case 128: return launch_patch<128, 0, 7, 64, PEG>(d, tmA, tmW, g, st);  // 64 KB of weights + 7 x 20 KB patches, 1 CTA per SM
    case 256:  // phase-split: each CTA owns one output-row parity = 128 columns, 128 KB of weights + 4 x 20 KB patches
      return launch_patch<128, 0, 4, 128, PEG, true>(d, tmA, tmW, g, st);
    default: break;
  }
  set_last_error("igemm patch_mode: N=%d unsupported", d->N);
  return MMDYN_ERR_ARG;
}

template <int BLOCK_N>
int launch_igemm(const mmdyn_igemm_desc* d, const CUtensorMap& tm, int m_tiles, int n_tiles, int total_tiles,
                 int occ, cudaStream_t st) {
  int grid = g_sm_count * occ;
  if (grid > total_tiles) grid = total_tiles;
  MMDYN_LAUNCH((igemm_kernel<BLOCK_N>), grid, IGEMM_THREADS, Cfg<BLOCK_N>::SMEM_BYTES, st, *d, tm, m_tiles, n_tiles,
                                                                                 total_tiles);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

template <int CN>
int launch_wgrad(const mmdyn_wgrad_desc* d, dim3 grid, cudaStream_t st) {
  MMDYN_LAUNCH((wgrad_kernel<CN>), grid, CTA_THREADS, Cfg<CN>::SMEM_BYTES, st, *d);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

}  // namespace

int igemm_init() {
#define SET_SMEM(K, BYTES) \
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES))
  SET_SMEM(igemm_kernel<16>, Cfg<16>::SMEM_BYTES);
  SET_SMEM(igemm_kernel<32>, Cfg<32>::SMEM_BYTES);
  SET_SMEM(igemm_kernel<64>, Cfg<64>::SMEM_BYTES);
  SET_SMEM(igemm_kernel<128>, Cfg<128>::SMEM_BYTES);
  SET_SMEM(igemm_kernel<256>, Cfg<256>::SMEM_BYTES);
  SET_SMEM(wgrad_kernel<16>, Cfg<16>::SMEM_BYTES);
  SET_SMEM(wgrad_kernel<32>, Cfg<32>::SMEM_BYTES);
  SET_SMEM(wgrad_kernel<64>, Cfg<64>::SMEM_BYTES);
  SET_SMEM(wgrad_kernel<128>, Cfg<128>::SMEM_BYTES);
  SET_SMEM(wgrad_kernel<256>, Cfg<256>::SMEM_BYTES);
#undef SET_SMEM
  int dev = 0;
  MMDYN_CHECK_CUDA(cudaGetDevice(&dev));
  MMDYN_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
#define OCC(I, K, BYTES)                                                                                \
  MMDYN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_igemm_occ[I], K, IGEMM_THREADS, BYTES)); \
  if (g_igemm_occ[I] < 1) g_igemm_occ[I] = 1
  OCC(0, igemm_kernel<16>, Cfg<16>::SMEM_BYTES);
  OCC(1, igemm_kernel<32>, Cfg<32>::SMEM_BYTES);
  OCC(2, igemm_kernel<64>, Cfg<64>::SMEM_BYTES);
  OCC(3, igemm_kernel<128>, Cfg<128>::SMEM_BYTES);
  OCC(4, igemm_kernel<256>, Cfg<256>::SMEM_BYTES);
#undef OCC
#define SET_TMA(I, BN)                                                                                        \
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_tma_kernel<BN, 0, (BN == 256 ? 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        Cfg<BN>::SMEM_BYTES));                                               \
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_tma_kernel<BN, 1, (BN == 256 ? 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        Cfg<BN>::SMEM_BYTES));                                               \
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_tma_kernel<BN, 2, (BN == 256 ? 2 : 1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        Cfg<BN>::SMEM_BYTES));                                               \
  MMDYN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_tma_occ[I], igemm_tma_kernel<BN, 0, (BN == 256 ? 2 : 1)>,      \
                                                                TMA_THREADS, Cfg<BN>::SMEM_BYTES));          \
  if (g_tma_occ[I] < 1) g_tma_occ[I] = 1
  SET_TMA(0, 16);
  SET_TMA(1, 32);
  SET_TMA(2, 64);
  SET_TMA(3, 128);
  SET_TMA(4, 256);
#undef SET_TMA
  MMDYN_CHECK_CUDA(cudaFuncSetAttribute(igemm_tma_kernel<32, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<32>::SMEM_BYTES));
  // cudaOccupancyMaxActiveBlocksPerMultiprocessor under-reports these kernels (it returned <= 1 for every
  // variant on B200 although ncu shows a shared-memory limit of 2-3 CTAs), so the persistent grids
  // are sized from the resources directly: 228 KB shared memory per SM, 1 KB reserved per CTA.
  const int smem_dyn[5] = {Cfg<16>::SMEM_BYTES, Cfg<32>::SMEM_BYTES, Cfg<64>::SMEM_BYTES, Cfg<128>::SMEM_BYTES,
                           Cfg<256>::SMEM_BYTES};
  for (int i = 0; i < 5; ++i) {
    const int by_smem_tma = (228 * 1024) / (smem_dyn[i] + 1024 + 256);
    const int by_smem_cpa = (228 * 1024) / (smem_dyn[i] + 1024 + 4608);
    g_tma_occ[i] = by_smem_tma < 1 ? 1 : (by_smem_tma > 8 ? 8 : by_smem_tma);
    g_igemm_occ[i] = by_smem_cpa < 1 ? 1 : (by_smem_cpa > 6 ? 6 : by_smem_cpa);
  }
  // TMEM: 512 columns per SM, two accumulator stages per CTA
  const int cols[5] = {64, 64, 128, 256, 512};
  for (int i = 0; i < 5; ++i) {
    if (g_igemm_occ[i] * cols[i] > 512) g_igemm_occ[i] = 512 / cols[i];
    if (g_tma_occ[i] * cols[i] > 512) g_tma_occ[i] = 512 / cols[i];
  }
  // experiment knob: cap the resident CTAs per SM of the persistent GEMM grids (leaves shared memory for
  // co-resident CTAs of the HBM-bound BatchNorm kernels of the other modality's branch)
  if (const char* cap = std::getenv("MMDYN_GEMM_OCC_CAP")) {
    const int c = std::atoi(cap);
    if (c >= 1)
      for (int i = 0; i < 5; ++i) {
        if (g_tma_occ[i] > c) g_tma_occ[i] = c;
        if (g_igemm_occ[i] > c) g_igemm_occ[i] = c;
      }
  }
  return MMDYN_OK;
}

}  // namespace mmdyn

using namespace mmdyn;

extern "C" int mmdyn_igemm(const mmdyn_igemm_desc* d, void* stream) {
  MMDYN_REQUIRE(d && d->A && d->W && d->out, "igemm: null pointer");
  MMDYN_REQUIRE(d->Cin > 0 && d->Cin % 8 == 0, "igemm: Cin=%d must be a positive multiple of 8", d->Cin);
  MMDYN_REQUIRE(d->ntaps >= 1 && d->ntaps <= MMDYN_MAX_TAPS, "igemm: ntaps=%d", d->ntaps);
  MMDYN_REQUIRE((d->ntaps * d->Cin) % 64 == 0, "igemm: ntaps*Cin=%d must be a multiple of 64",
                d->ntaps * d->Cin);
  MMDYN_REQUIRE(d->n_phases >= 1 && d->n_phases <= MMDYN_MAX_PHASES, "igemm: n_phases=%d", d->n_phases);
  MMDYN_REQUIRE(d->block_n == 16 || d->block_n == 32 || d->block_n == 64 || d->block_n == 128 ||
                    d->block_n == 256,
                "igemm: block_n=%d", d->block_n);
  MMDYN_REQUIRE(d->N > 0 && d->N % d->block_n == 0, "igemm: N=%d not a multiple of block_n=%d", d->N,
                d->block_n);
  MMDYN_REQUIRE(d->ksplit >= 1 && (d->ksplit == 1 || d->out_mode == 2),
                "igemm: ksplit=%d needs out_mode 2", d->ksplit);
  MMDYN_REQUIRE(d->out_mode >= 0 && d->out_mode <= 5, "igemm: out_mode=%d", d->out_mode);
  if (d->out_mode == 5) {
    MMDYN_REQUIRE(d->block_n == 16 && d->N == 16 && d->bce_target && d->bce_loss && d->bce_rows_per_group > 0 &&
                      d->n_img <= MMDYN_MAX_GROUPS * d->bce_rows_per_group && d->P >= 128 && d->P % 128 == 0 &&
                      d->ksplit == 1,
                  "igemm: out_mode 5 needs N=16, bce_target, bce_loss, rows_per_group > 0, <= %d groups, "
                  "P a multiple of 128", MMDYN_MAX_GROUPS);
  }
  MMDYN_REQUIRE(d->out_mode != 4 || (d->ldc % 16 == 0 && d->N == 4 * d->ldc && d->s_out == 2 && d->n_phases == 1),
                "igemm: out_mode 4 needs N = 4*ldc, ldc %% 16 == 0, s_out = 2");
  MMDYN_REQUIRE(d->out_mode != 3 || (d->block_n == 16 && d->N == 16), "igemm: out_mode 3 needs N=16");
  MMDYN_REQUIRE((d->out_mode != 1 && d->out_mode != 2) || d->ldc % 4 == 0,
                "igemm: fp32 outputs (out_mode 1 / 2) are written as 16-byte vectors: ldc=%d must be a multiple of 4",
                d->ldc);
  MMDYN_REQUIRE(d->row_mode == 0 || (d->row_mode == 1 && d->ksplit == 1 && d->Cin % 64 == 0),
                "igemm: row_mode=%d (row_mode 1 needs ksplit 1 and Cin %% 64 == 0)", d->row_mode);
  MMDYN_REQUIRE(d->n_img > 0 && d->P > 0 && d->OXv > 0, "igemm: empty problem");
  MMDYN_REQUIRE((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 && d->a_pix_stride % 8 == 0,
                "igemm: operands must be 16-byte aligned");
  const long long rows = static_cast<long long>(d->n_img) * d->P;
  MMDYN_REQUIRE(rows * d->ldc < (1LL << 31) &&
                    static_cast<long long>(d->n_img) * d->IH * d->IW * d->a_pix_stride < (1LL << 31),
                "igemm: tensor too large for 32-bit offsets");
  const bool custom_strides = d->a_row_stride != 0 || d->a_img_stride != 0 || d->a_pix_stride < d->Cin;
  MMDYN_REQUIRE(d->a_row_stride % 8 == 0 && d->a_img_stride % 8 == 0 && d->a_row_stride >= 0 && d->a_img_stride >= 0,
                "igemm: a_row_stride / a_img_stride must be non-negative multiples of 8 elements");

  if (d->patch_mode) return igemm_patch(d, static_cast<cudaStream_t>(stream));
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("igemm: cuTensorMapEncodeTiled driver entry point not available");
    return MMDYN_ERR_CUDA;
  }
  const int ktot = d->ntaps * d->Cin;
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(d->n_phases) * d->N};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ktot) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(d->block_n)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->W), gdim, gstr, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_last_error("igemm: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(cr));
    return MMDYN_ERR_CUDA;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_tiles = d->N / d->block_n;
  const int occ_idx = d->block_n == 16 ? 0 : d->block_n == 32 ? 1 : d->block_n == 64 ? 2 : d->block_n == 128 ? 3 : 4;

  // ---- TMA-fed path: tile = box of bw x bh virtual pixels x bn images, or pixel-major -------------
  const int a_mode = (d->Cin % 64 == 0) ? 0 : (d->Cin == 32 ? 1 : (d->Cin == 8 ? 2 : (d->Cin == 16 ? 3 : -1)));
  static const bool legacy = getenv("MMDYN_IGEMM_LEGACY") != nullptr;
  if (a_mode >= 0 && !legacy && (a_mode != 1 || d->ntaps % 2 == 0) && (a_mode != 2 || d->ntaps % 8 == 0) &&
      (a_mode != 3 || d->ntaps % 4 == 0)) {
    TileGeom g = {};
    const int OYv = d->P / d->OXv;
    const int bw = d->OXv;
    bool box_ok = d->row_mode == 0 && bw <= TILE_M && (TILE_M % bw) == 0 && d->P > 1;
    int bh = 1, bn = TILE_M;
    if (box_ok) {
      bh = TILE_M / bw < OYv ? TILE_M / bw : OYv;
      box_ok = (OYv % bh) == 0 && (bh & (bh - 1)) == 0 && (TILE_M % (bw * bh)) == 0;
      bn = TILE_M / (bw * bh);
    }
    // pixel-major tiles need every k-block to belong to one tap when taps can be skipped
    g.pixel_major = box_ok ? 0 : 1;
    if (g.pixel_major) {
      bh = 1;
      bn = TILE_M;
    }
    g.lbw = g.pixel_major ? 0 : ilog2(bw);
    g.lbh = g.pixel_major ? 0 : ilog2(bh);
    g.bn = bn;
    g.tiles_y = g.pixel_major ? 1 : OYv / bh;
    g.img_blocks = (d->n_img + bn - 1) / bn;
    const long long m_tiles_ll = g.pixel_major ? static_cast<long long>(d->P) * g.img_blocks
                                               : static_cast<long long>(g.img_blocks) * g.tiles_y;
    const long long total_ll = m_tiles_ll * n_tiles * d->n_phases * d->ksplit;
    MMDYN_REQUIRE(total_ll < (1LL << 31), "igemm: too many tiles");
    g.m_tiles = static_cast<int>(m_tiles_ll);
    g.n_tiles = n_tiles;
    g.total_tiles = static_cast<int>(total_ll);

    CUtensorMap tmA;
    const int kc = a_mode == 0 ? 64 : (a_mode == 1 ? 32 : (a_mode == 3 ? 16 : 8));
    const int es = g.pixel_major ? 1 : d->s_in;
    const int esx = g.pixel_major ? 1 : (d->s_in_x ? d->s_in_x : d->s_in);
    const cuuint64_t adim[4] = {static_cast<cuuint64_t>(d->Cin), static_cast<cuuint64_t>(d->IW),
                                static_cast<cuuint64_t>(d->IH), static_cast<cuuint64_t>(d->n_img)};
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->a_pix_stride) * 2;
    const cuuint64_t row_b = d->a_row_stride ? static_cast<cuuint64_t>(d->a_row_stride) * 2 : pix_b * d->IW;
    const cuuint64_t img_b = d->a_img_stride ? static_cast<cuuint64_t>(d->a_img_stride) * 2 : row_b * d->IH;
    const cuuint64_t astr[3] = {pix_b, row_b, img_b};
    const cuuint32_t abox[4] = {static_cast<cuuint32_t>(kc),
                                static_cast<cuuint32_t>(g.pixel_major ? 1 : bw * esx),
                                static_cast<cuuint32_t>(g.pixel_major ? 1 : bh * es), static_cast<cuuint32_t>(bn)};
    const cuuint32_t aes[4] = {1, static_cast<cuuint32_t>(esx), static_cast<cuuint32_t>(es), 1};
    const CUtensorMapSwizzle asw = a_mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : (a_mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
    CUresult ar;
    if (a_mode == 3) {
      // raw rows: (elements of a padded row, 1, rows, images); box = 256 elements x bh rows (see the kernel header)
      MMDYN_REQUIRE(!g.pixel_major && bw == 32 && bh == 4 && bn == 1 && d->a_pix_stride == 8 && esx == 1 &&
                        d->a_row_stride >= (d->IW + 1) * 8 && d->a_row_stride >= 264,
                    "igemm: Cin = 16 is the 4-pixel window form (OXv = 32, pixel pairs of 8 elements, s_in_x = 1)");
      for (int t_ = 0; t_ < d->ntaps; ++t_) MMDYN_REQUIRE(d->tap_dx[0][t_] == 0, "igemm: Cin = 16 takes row taps only");
      const cuuint64_t rdim[4] = {static_cast<cuuint64_t>(d->a_row_stride), 1, static_cast<cuuint64_t>(d->IH),
                                  static_cast<cuuint64_t>(d->n_img)};
      const cuuint64_t rstr[3] = {row_b, row_b, img_b};
      const cuuint32_t rbox[4] = {256, 1, static_cast<cuuint32_t>(bh * es), 1};
      const cuuint32_t res[4] = {1, 1, static_cast<cuuint32_t>(es), 1};
      ar = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->A), rdim, rstr, rbox, res,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      ar = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->A), adim, astr, abox, aes,
               CU_TENSOR_MAP_INTERLEAVE_NONE, asw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (ar != CUDA_SUCCESS) {
      set_last_error("igemm: cuTensorMapEncodeTiled(A) failed with CUresult %d (Cin=%d IW=%d IH=%d box=%u,%u,%u,%u es=%d)",
                     static_cast<int>(ar), d->Cin, d->IW, d->IH, abox[0], abox[1], abox[2], abox[3], es);
      return MMDYN_ERR_CUDA;
    }
    MMDYN_REQUIRE(d->out_mode != 5 || (g.bn == 1 && !g.pixel_major),
                  "igemm: out_mode 5 needs tiles that lie within one image (OXv=%d P=%d)", d->OXv, d->P);
    if (d->bn_sums) {
      MMDYN_REQUIRE(d->out_mode == 0 && (a_mode == 0 || a_mode == 1) && d->block_n >= 64 && n_tiles == 1 && d->ksplit == 1 &&
                        d->n_phases == 1 && d->bn_rows_per_group > 0 && d->bn_rows_per_group % bn == 0,
                    "igemm: bn_sums needs out_mode 0, Cin %% 64 == 0 or 32, 64 <= N <= 256, one phase, no K split and "
                    "rows_per_group %% %d == 0 (images per tile)", bn);
    }
    // tile pairs sharing each weight stage (igemm_pair_kernel): the L2-bound conv layers with enough work per SM
    static const int pair_mode = getenv("MMDYN_IGEMM_PAIR") ? atoi(getenv("MMDYN_IGEMM_PAIR")) : 1;
    if (pair_mode && d->out_mode == 0 && d->bias == nullptr && d->ksplit == 1 && n_tiles == 1 && d->n_phases == 1 &&
        (g.total_tiles & 1) == 0 && (!g.pixel_major || (g.img_blocks & 1) == 0) &&
        (pair_mode == 2 || g.total_tiles >= 4 * g_sm_count) && d->ntaps <= 32) {
      if (a_mode == 0 && d->block_n == 128) return launch_igemm_pair<128, 0>(d, tmA, tm, g, st);
      if (a_mode == 0 && d->block_n == 256) return launch_igemm_pair<256, 0>(d, tmA, tm, g, st);
      if (a_mode == 0 && d->block_n == 64) return launch_igemm_pair<64, 0>(d, tmA, tm, g, st);
      if (a_mode == 1 && d->block_n == 64 && !g.pixel_major && (pair_mode >= 2 || pair_mode == 1))
        return launch_igemm_pair<64, 1>(d, tmA, tm, g, st);
    }
    const int occ = g_tma_occ[occ_idx];
    switch (d->block_n) {
      case 16: return dispatch_amode<16>(a_mode, d, tmA, tm, g, occ, st);
      case 32: return dispatch_amode<32>(a_mode, d, tmA, tm, g, occ, st);
      case 64: return dispatch_amode<64>(a_mode, d, tmA, tm, g, occ, st);
      case 128: return dispatch_amode<128>(a_mode, d, tmA, tm, g, occ, st);
      default: return dispatch_amode<256>(a_mode, d, tmA, tm, g, occ, st);
    }
  }

  // ---- cp.async gather path (any dense geometry) -------------------------------------------------
  MMDYN_REQUIRE(d->out_mode != 5, "igemm: out_mode 5 needs the TMA path");
  MMDYN_REQUIRE(d->bn_sums == nullptr, "igemm: bn_sums needs the TMA path");
  MMDYN_REQUIRE(!custom_strides, "igemm: a_row_stride / a_img_stride / overlapping windows need the TMA path "
                "(Cin=%d ntaps=%d)", d->Cin, d->ntaps);
  long long m_tiles_ll;
  if (d->row_mode == 0)
    m_tiles_ll = (rows + TILE_M - 1) / TILE_M;
  else
    m_tiles_ll = static_cast<long long>(d->P) * ((d->n_img + TILE_M - 1) / TILE_M);
  const long long total_ll = m_tiles_ll * n_tiles * d->n_phases * d->ksplit;
  MMDYN_REQUIRE(total_ll < (1LL << 31), "igemm: too many tiles");
  const int m_tiles = static_cast<int>(m_tiles_ll), total = static_cast<int>(total_ll);
  switch (d->block_n) {
    case 16: return launch_igemm<16>(d, tm, m_tiles, n_tiles, total, g_igemm_occ[0], st);
    case 32: return launch_igemm<32>(d, tm, m_tiles, n_tiles, total, g_igemm_occ[1], st);
    case 64: return launch_igemm<64>(d, tm, m_tiles, n_tiles, total, g_igemm_occ[2], st);
    case 128: return launch_igemm<128>(d, tm, m_tiles, n_tiles, total, g_igemm_occ[3], st);
    default: return launch_igemm<256>(d, tm, m_tiles, n_tiles, total, g_igemm_occ[4], st);
  }
}

// diagnostics (not part of the public header): resident-CTA estimates used to size persistent grids
extern "C" int mmdyn_debug_occ(int which, int idx) {
  if (idx < 0 || idx > 4) return -1;
  return which == 0 ? g_igemm_occ[idx] : (which == 1 ? g_tma_occ[idx] : g_sm_count);
}

extern "C" int mmdyn_wgrad(const mmdyn_wgrad_desc* d, void* stream) {
  MMDYN_REQUIRE(d && d->G && d->Nat && d->dW, "wgrad: null pointer");
  MMDYN_REQUIRE(d->Cg > 0 && d->Cg % 8 == 0 && ((d->ntaps * d->Cg) % 128 == 0 || (d->Cg == 16 && d->ntaps == 4)),
                "wgrad: Cg=%d ntaps=%d (ntaps*Cg must be a multiple of 128, or the 4 x 16 window form)", d->Cg, d->ntaps);
  MMDYN_REQUIRE(d->ntaps >= 1 && d->ntaps <= MMDYN_MAX_TAPS, "wgrad: ntaps=%d", d->ntaps);
  const int cn_tile = d->Cn >= 256 ? 256 : d->Cn;
  MMDYN_REQUIRE(cn_tile == 16 || cn_tile == 32 || cn_tile == 64 || cn_tile == 128 || cn_tile == 256,
                "wgrad: Cn=%d unsupported", d->Cn);
  MMDYN_REQUIRE(d->Cn % cn_tile == 0, "wgrad: Cn=%d not a multiple of %d", d->Cn, cn_tile);
  MMDYN_REQUIRE(d->row_splits >= 1, "wgrad: row_splits=%d", d->row_splits);
  MMDYN_REQUIRE((reinterpret_cast<uintptr_t>(d->G) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->Nat) & 15) == 0 &&
                    d->g_pix_stride % 8 == 0 && d->nat_stride % 8 == 0,
                "wgrad: operands must be 16-byte aligned");
  MMDYN_REQUIRE(static_cast<long long>(d->n_img) * d->IH * d->IW * d->g_pix_stride < (1LL << 31),
                "wgrad: tensor too large for 32-bit offsets");
  const bool custom_strides = d->g_row_stride != 0 || d->g_img_stride != 0 || d->g_pix_stride < d->Cg;
  MMDYN_REQUIRE(d->g_row_stride % 8 == 0 && d->g_img_stride % 8 == 0 && d->g_row_stride >= 0 && d->g_img_stride >= 0,
                "wgrad: g_row_stride / g_img_stride must be non-negative multiples of 8 elements");
  dim3 grid(d->row_splits, (d->ntaps * d->Cg + 127) / 128, d->Cn / cn_tile);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // ---- TMA-fed path ---------------------------------------------------------------------------------
  // Cg == 8 (16-byte pixels: logit gradients, repacked input) stays on the cp.async gather: sixteen
  // 1 KB boxes of 16-byte rows per step are slower through TMA than 16-byte LDGSTS (measured)
  static const bool g8_tma = getenv("MMDYN_WGRAD_G8_TMA") != nullptr;
  const int g_mode = (d->Cg % 64 == 0) ? 0
                     : d->Cg == 32     ? 1
                     : (d->Cg == 16 && d->ntaps == 4) ? 3
                     : (d->Cg == 8 && g8_tma)         ? 2
                                                      : -1;
  static const bool legacy = getenv("MMDYN_WGRAD_LEGACY") != nullptr;
  EncodeTiledFn enc = get_encode_fn();
  if (g_mode >= 0 && !legacy && enc) {
    WgradGeomDev g = {};
    const int OYv = d->P / d->OXv, bw = d->OXv;
    bool box_ok = d->P > 1 && bw <= 64 && (64 % bw) == 0;
    int bh = 1, bn = 64;
    if (box_ok) {
      bh = 64 / bw < OYv ? 64 / bw : OYv;
      box_ok = (OYv % bh) == 0 && (bh & (bh - 1)) == 0 && (64 % (bw * bh)) == 0;
      bn = 64 / (bw * bh);
    }
    g.pixel_major = box_ok ? 0 : 1;
    if (g.pixel_major) {
      bh = 1;
      bn = 64;
    }
    g.lbw = g.pixel_major ? 0 : ilog2(bw);
    g.lbh = g.pixel_major ? 0 : ilog2(bh);
    g.bn = bn;
    g.tiles_y = g.pixel_major ? 1 : OYv / bh;
    g.img_blocks = (d->n_img + bn - 1) / bn;
    const long long steps = g.pixel_major ? static_cast<long long>(d->P) * g.img_blocks
                                          : static_cast<long long>(g.img_blocks) * g.tiles_y;
    MMDYN_REQUIRE(steps < (1LL << 31), "wgrad: too many steps");
    g.total_steps = static_cast<int>(steps);
    if (static_cast<long long>(grid.x) > steps) grid.x = static_cast<unsigned>(steps);

    const int es = g.pixel_major ? 1 : d->s_in;
    const int esx = g.pixel_major ? 1 : (d->s_in_x ? d->s_in_x : d->s_in);
    const int kc = g_mode == 0 ? 64 : (g_mode == 1 ? 32 : (g_mode == 3 ? 16 : 8));
    CUtensorMap tmG, tmN;
    {
      const cuuint64_t dim[4] = {static_cast<cuuint64_t>(d->Cg), static_cast<cuuint64_t>(d->IW),
                                 static_cast<cuuint64_t>(d->IH), static_cast<cuuint64_t>(d->n_img)};
      const cuuint64_t pb = static_cast<cuuint64_t>(d->g_pix_stride) * 2;
      const cuuint64_t rb = d->g_row_stride ? static_cast<cuuint64_t>(d->g_row_stride) * 2 : pb * d->IW;
      const cuuint64_t ib = d->g_img_stride ? static_cast<cuuint64_t>(d->g_img_stride) * 2 : rb * d->IH;
      const cuuint64_t str[3] = {pb, rb, ib};
      const cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(g.pixel_major ? 1 : bw * esx),
                                 static_cast<cuuint32_t>(g.pixel_major ? 1 : bh * es), static_cast<cuuint32_t>(bn)};
      const cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(esx), static_cast<cuuint32_t>(es), 1};
      const CUtensorMapSwizzle sw = g_mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                : (g_mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
      CUresult r;
      if (g_mode == 3) {
        // raw rows: (elements of a padded row, 1, rows, images); box = 256 elements x bh rows (see the kernel header)
        MMDYN_REQUIRE(!g.pixel_major && bw == 32 && bh == 2 && bn == 1 && d->g_pix_stride == 8 && esx == 1 &&
                          d->g_row_stride >= (d->IW + 1) * 8 && d->g_row_stride >= 264,
                      "wgrad: Cg = 16 is the 4-pixel window form (OXv = 32, pixel pairs of 8 elements, s_in_x = 1)");
        for (int t_ = 0; t_ < d->ntaps; ++t_) MMDYN_REQUIRE(d->tap_dx[t_] == 0, "wgrad: Cg = 16 takes row taps only");
        const cuuint64_t rdim[4] = {static_cast<cuuint64_t>(d->g_row_stride), 1, static_cast<cuuint64_t>(d->IH),
                                    static_cast<cuuint64_t>(d->n_img)};
        const cuuint64_t rstr[3] = {rb, rb, ib};
        const cuuint32_t rbox[4] = {256, 1, static_cast<cuuint32_t>(bh * es), 1};
        const cuuint32_t res[4] = {1, 1, static_cast<cuuint32_t>(es), 1};
        r = enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->G), rdim, rstr, rbox, res,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        r = enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->G), dim, str, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) {
        set_last_error("wgrad: cuTensorMapEncodeTiled(G) failed with CUresult %d", static_cast<int>(r));
        return MMDYN_ERR_CUDA;
      }
    }
    {
      const cuuint64_t dim[4] = {static_cast<cuuint64_t>(d->Cn), static_cast<cuuint64_t>(d->OXv),
                                 static_cast<cuuint64_t>(OYv), static_cast<cuuint64_t>(d->n_img)};
      const cuuint64_t pb = static_cast<cuuint64_t>(d->nat_stride) * 2;
      const cuuint64_t str[3] = {pb, pb * d->OXv, pb * d->P};
      const int cnb = cn_tile >= 64 ? 64 : cn_tile;
      const cuuint32_t box[4] = {static_cast<cuuint32_t>(cnb), static_cast<cuuint32_t>(g.pixel_major ? 1 : bw),
                                 static_cast<cuuint32_t>(g.pixel_major ? 1 : bh), static_cast<cuuint32_t>(bn)};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      const CUtensorMapSwizzle sw = cn_tile >= 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (cn_tile == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      const CUresult r = enc(&tmN, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->Nat), dim, str, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_last_error("wgrad: cuTensorMapEncodeTiled(Nat) failed with CUresult %d", static_cast<int>(r));
        return MMDYN_ERR_CUDA;
      }
    }
    switch (cn_tile) {
      case 16: return dispatch_gmode<16>(g_mode, d, tmG, tmN, g, grid, st);
      case 32: return dispatch_gmode<32>(g_mode, d, tmG, tmN, g, grid, st);
      case 64: return dispatch_gmode<64>(g_mode, d, tmG, tmN, g, grid, st);
      case 128: return dispatch_gmode<128>(g_mode, d, tmG, tmN, g, grid, st);
      default: return dispatch_gmode<256>(g_mode, d, tmG, tmN, g, grid, st);
    }
  }

  MMDYN_REQUIRE(!custom_strides, "wgrad: g_row_stride / g_img_stride / overlapping windows need the TMA path "
                "(Cg=%d ntaps=%d)", d->Cg, d->ntaps);
  switch (cn_tile) {
    case 16: return launch_wgrad<16>(d, grid, st);
    case 32: return launch_wgrad<32>(d, grid, st);
    case 64: return launch_wgrad<64>(d, grid, st);
    case 128: return launch_wgrad<128>(d, grid, st);
    default: return launch_wgrad<256>(d, grid, st);
  }
}

extern "C" int mmdyn_conv1_fwd(const float* x_nchw, const void* Wp, void* out, void* act_out, int n_img, void* stream) {
  MMDYN_REQUIRE(x_nchw && Wp && out && n_img > 0, "conv1_fwd: bad arguments");
  MMDYN_LAUNCH((conv1_fwd_kernel), n_img * 8, 128, 0, static_cast<cudaStream_t>(stream), 
      x_nchw, reinterpret_cast<const __half*>(Wp), reinterpret_cast<__half*>(out), reinterpret_cast<__half*>(act_out), n_img);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}

extern "C" int mmdyn_conv1_wgrad(const float* x_nchw, const void* dRaw, float* dW, int n_img, float scale,
                                 int row_splits, void* stream) {
  MMDYN_REQUIRE(x_nchw && dRaw && dW && n_img > 0 && row_splits > 0, "conv1_wgrad: bad arguments");
  const long long total = static_cast<long long>(n_img) * 1024;
  long long per = (total + row_splits - 1) / row_splits;
  per = (per + 63) / 64 * 64;
  const int grid = static_cast<int>((total + per - 1) / per);
  MMDYN_LAUNCH((conv1_wgrad_kernel), grid, 256, 0, static_cast<cudaStream_t>(stream), 
      x_nchw, reinterpret_cast<const __half*>(dRaw), dW, n_img, scale, static_cast<int>(per));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}
