// Device-side input pipeline in front of the step: uint8 frames resident in
// HBM -> the fp32 NCHW batch the encoders read, in one kernel.
//
// Replaces, per frame, transforms.Compose([Resize(input_size), ToTensor()])
// (mmdyn/pytorch/utils/datasets.py:23-31, applied in _parse_list_data :382-392), and per batch the
// row selection of seq_collate_fn + parse_input (datasets.py:395-404, problems.py:634-673) through a
// frame index.  The arithmetic is Pillow's (src/libImaging/Resample.c, BILINEAR = antialiased
// triangle filter in 22-bit fixed point, horizontal pass then vertical pass, each rounded to uint8),
// followed by ToTensor's uint8 / 255: the result is bit-identical to the reference's CPU loader.
//
// The coefficient tables depend on the four sizes only: mmdyn_resize_table builds them on the host
// in double precision exactly as precompute_coeffs / normalize_coeffs_8bpc do; the caller keeps them
// in device memory.  HBM-bound byte work: one CTA per (output row, frame) stages the <= ksize_y
// source rows it needs in shared memory with 16-byte loads, filters horizontally into shared
// memory, then vertically, and writes three 256-byte plane rows.
#include "common.cuh"
#include "../../include/mmdyn_b200.h"

#include <atomic>
#include <cmath>
#include <cstdlib>

namespace mmdyn {
extern std::atomic<long long> g_launch_count;
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

// table layout (int32): [0] ksize_x  [1] ksize_y  [2] in_h  [3] in_w  [4] out_h  [5] out_w  [6..7] reserved
//   then bounds_x[out_w][2], kk_x[out_w][ksize_x], bounds_y[out_h][2], kk_y[out_h][ksize_y]
constexpr int TABLE_HEADER = 8;

int axis_ksize(int in_size, int out_size) {
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  return static_cast<int>(std::ceil(support)) * 2 + 1;
}

double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc for one axis (box = whole axis)
void axis_coeffs(int in_size, int out_size, int ksize, int32_t* bounds, int32_t* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  double w[64];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bilinear_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < ksize; ++x) {
      double k = 0.0;
      if (x < xmax) k = (ww != 0.0) ? w[x] / ww : w[x];
      kk[xx * ksize + x] = k < 0 ? static_cast<int>(-0.5 + k * (1 << PRECISION_BITS))
                                 : static_cast<int>(0.5 + k * (1 << PRECISION_BITS));
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// grid = (out_h, n_frames); block = 256 threads; dynamic smem = ksize_y * (in_w*3 rounded up to 16 + out_w*3) bytes
__global__ void __launch_bounds__(256)
frames_u8_to_f32_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ index,
                        const int32_t* __restrict__ table, float* __restrict__ dst, long long frame_stride) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t fr_smem[];
  const int ksx = table[0], ksy = table[1], in_h = table[2], in_w = table[3], out_h = table[4], out_w = table[5];
  const int32_t* bx = table + TABLE_HEADER;
  const int32_t* kx = bx + out_w * 2;
  const int32_t* by = kx + out_w * ksx;
  const int32_t* ky = by + out_h * 2;
  const int yy = blockIdx.x;
  const long long frame = index ? index[blockIdx.y] : blockIdx.y;
  const uint8_t* img = src + frame * frame_stride;
  const int row_bytes = in_w * 3;
  const int row_pitch = (row_bytes + 15) & ~15;
  const bool vpass = in_h != out_h, hpass = in_w != out_w;
  const int ymin = vpass ? by[yy * 2] : yy;
  const int ycnt = vpass ? by[yy * 2 + 1] : 1;
  uint8_t* rows = fr_smem;                      // [ycnt][row_pitch] source rows
  uint8_t* tmp = fr_smem + ksy * row_pitch;     // [ycnt][out_w*3] horizontally filtered rows

  // stage the source rows (rows are contiguous in memory: one flat range of ycnt*row_bytes bytes)
  {
    const uint8_t* g = img + static_cast<long long>(ymin) * row_bytes;
    const int total = ycnt * row_bytes;
    if ((row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const uint4* g4 = reinterpret_cast<const uint4*>(g);
      uint4* s4 = reinterpret_cast<uint4*>(rows);
      for (int i = threadIdx.x; i < (total >> 4); i += blockDim.x) s4[i] = __ldg(g4 + i);
    } else {
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / row_bytes, b = i - r * row_bytes;
        rows[r * row_pitch + b] = __ldg(g + i);
      }
    }
  }
  __syncthreads();
  const int ow3 = out_w * 3;
  // horizontal pass: tmp[r][xx][c] = clip8(2^21 + sum_x rows[r][xmin + x][c] * kx[xx][x])
  for (int i = threadIdx.x; i < ycnt * ow3; i += blockDim.x) {
    const int r = i / ow3, j = i - r * ow3, xx = j / 3, c = j - xx * 3;
    const uint8_t* p = rows + r * row_pitch;
    int v;
    if (hpass) {
      const int xmin = bx[xx * 2], xmax = bx[xx * 2 + 1];
      const int32_t* k = kx + xx * ksx;
      int ss = 1 << (PRECISION_BITS - 1);
      for (int x = 0; x < xmax; ++x) ss += static_cast<int>(p[(xmin + x) * 3 + c]) * __ldg(k + x);
      v = clip8(ss);
    } else {
      v = p[xx * 3 + c];
    }
    tmp[i] = static_cast<uint8_t>(v);
  }
  __syncthreads();
  // vertical pass + ToTensor: dst[n][c][yy][xx] = clip8(...) / 255
  const long long plane = static_cast<long long>(out_h) * out_w;
  float* o = dst + static_cast<long long>(blockIdx.y) * 3 * plane + static_cast<long long>(yy) * out_w;
  for (int j = threadIdx.x; j < ow3; j += blockDim.x) {
    const int c = j / out_w, xx = j - c * out_w;  // plane-major so that a warp writes one contiguous row segment
    int v;
    if (vpass) {
      const int32_t* k = ky + yy * ksy;
      int ss = 1 << (PRECISION_BITS - 1);
      for (int y = 0; y < ycnt; ++y) ss += static_cast<int>(tmp[y * ow3 + xx * 3 + c]) * __ldg(k + y);
      v = clip8(ss);
    } else {
      v = tmp[xx * 3 + c];
    }
    o[c * plane + xx] = __fdiv_rn(static_cast<float>(v), 255.0f);
  }
}

// R output rows of one frame per CTA: the source rows of consecutive output rows overlap (a 256 -> 64 resize reads 9 source
// rows per output row, advancing by 4), so the one-row kernel above fetches every source row ~2.25x (from L2) and filters it
// horizontally 2.25x.  Here the union of the R rows' windows — (R - 1) * scale + ksize_y rows — is staged and filtered ONCE
// (37 instead of 72 row filterings for R = 8), the horizontal coefficients sit in shared memory, and the vertical pass reads
// the filtered rows it needs.  Same arithmetic, same order: bit-identical to the one-row kernel and to Pillow.
// grid = (ceil(out_h / R), n_frames); dynamic smem = rows_max * (row_pitch + out_w * 3) + out_w * (2 + ksize_x) * 4
__global__ void __launch_bounds__(256)
frames_u8_to_f32_rows_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ index,
                             const int32_t* __restrict__ table, float* __restrict__ dst, long long frame_stride, int R,
                             int rows_max) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t fr_smem[];
  const int ksx = table[0], ksy = table[1], in_h = table[2], in_w = table[3], out_h = table[4], out_w = table[5];
  const int32_t* bx = table + TABLE_HEADER;
  const int32_t* kx = bx + out_w * 2;
  const int32_t* by = kx + out_w * ksx;
  const int32_t* ky = by + out_h * 2;
  const int yy0 = blockIdx.x * R, yy1 = min(out_h, yy0 + R);
  const long long frame = index ? index[blockIdx.y] : blockIdx.y;
  const uint8_t* img = src + frame * frame_stride;
  const int row_bytes = in_w * 3;
  const int row_pitch = (row_bytes + 15) & ~15;
  const bool vpass = in_h != out_h, hpass = in_w != out_w;
  const int ymin0 = vpass ? by[yy0 * 2] : yy0;
  const int ylast = vpass ? by[(yy1 - 1) * 2] + by[(yy1 - 1) * 2 + 1] : yy1;  // one past the last source row needed
  const int nrows = ylast - ymin0;
  const int ow3 = out_w * 3;
  uint8_t* rows = fr_smem;                                            // [nrows][row_pitch] source rows
  uint8_t* tmp = fr_smem + static_cast<size_t>(rows_max) * row_pitch;  // [nrows][ow3] horizontally filtered rows
  int32_t* sbx = reinterpret_cast<int32_t*>(fr_smem + ((static_cast<size_t>(rows_max) * (row_pitch + ow3) + 15) & ~size_t(15)));
  int32_t* skx = sbx + out_w * 2;
  if (hpass)
    for (int i = threadIdx.x; i < out_w * (2 + ksx); i += blockDim.x) sbx[i] = __ldg(bx + i);  // bounds_x then kk_x, contiguous
  {
    const uint8_t* g = img + static_cast<long long>(ymin0) * row_bytes;
    const int total = nrows * row_bytes;
    if ((row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const uint4* g4 = reinterpret_cast<const uint4*>(g);
      uint4* s4 = reinterpret_cast<uint4*>(rows);
      for (int i = threadIdx.x; i < (total >> 4); i += blockDim.x) s4[i] = __ldg(g4 + i);
    } else {
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / row_bytes, b = i - r * row_bytes;
        rows[r * row_pitch + b] = __ldg(g + i);
      }
    }
  }
  __syncthreads();
  // horizontal pass over every staged row: tmp[r][xx][c] = clip8(2^21 + sum_x rows[r][xmin + x][c] * kx[xx][x])
  constexpr int KMAX = 12;
  if (hpass && ksx <= KMAX && out_w <= static_cast<int>(blockDim.x)) {
    // thread = (output column, row group): the column's coefficients stay in registers, the three channels of a pixel
    // share them, and the tap loop is unrolled — the table-driven loop below spent most of its issue slots on
    // coefficient loads and index arithmetic (the kernel is integer-MAC bound, not traffic bound)
    const int xx = threadIdx.x % out_w, rg = threadIdx.x / out_w, ngroups = blockDim.x / out_w;
    if (rg < ngroups) {
      const int xmin = sbx[xx * 2], xmax = sbx[xx * 2 + 1];
      int kc[KMAX];
#pragma unroll
      for (int x = 0; x < KMAX; ++x) kc[x] = x < xmax ? skx[xx * ksx + x] : 0;
      for (int r = rg; r < nrows; r += ngroups) {
        const uint8_t* p = rows + r * row_pitch + xmin * 3;
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
#pragma unroll
        for (int x = 0; x < KMAX; ++x)
          if (x < xmax) {
            s0 += static_cast<int>(p[3 * x]) * kc[x];
            s1 += static_cast<int>(p[3 * x + 1]) * kc[x];
            s2 += static_cast<int>(p[3 * x + 2]) * kc[x];
          }
        uint8_t* o = tmp + r * ow3 + xx * 3;
        o[0] = static_cast<uint8_t>(clip8(s0));
        o[1] = static_cast<uint8_t>(clip8(s1));
        o[2] = static_cast<uint8_t>(clip8(s2));
      }
    }
  } else {
    for (int i = threadIdx.x; i < nrows * ow3; i += blockDim.x) {
      const int r = i / ow3, j = i - r * ow3, xx = j / 3, c = j - xx * 3;
      const uint8_t* p = rows + r * row_pitch;
      int v;
      if (hpass) {
        const int xmin = sbx[xx * 2], xmax = sbx[xx * 2 + 1];
        const int32_t* k = skx + xx * ksx;
        int ss = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < xmax; ++x) ss += static_cast<int>(p[(xmin + x) * 3 + c]) * k[x];
        v = clip8(ss);
      } else {
        v = p[xx * 3 + c];
      }
      tmp[i] = static_cast<uint8_t>(v);
    }
  }
  __syncthreads();
  // vertical pass + ToTensor for the R output rows
  const long long plane = static_cast<long long>(out_h) * out_w;
  const int nout = (yy1 - yy0) * ow3;
  for (int i = threadIdx.x; i < nout; i += blockDim.x) {
    const int ry = i / ow3, j = i - ry * ow3;
    const int c = j / out_w, xx = j - c * out_w;  // plane-major so that a warp writes contiguous row segments
    const int yy = yy0 + ry;
    int v;
    if (vpass) {
      const int y0 = by[yy * 2] - ymin0, ycnt = by[yy * 2 + 1];
      const int32_t* k = ky + yy * ksy;
      int ss = 1 << (PRECISION_BITS - 1);
      for (int y = 0; y < ycnt; ++y) ss += static_cast<int>(tmp[(y0 + y) * ow3 + xx * 3 + c]) * __ldg(k + y);
      v = clip8(ss);
    } else {
      v = tmp[ry * ow3 + xx * 3 + c];
    }
    dst[static_cast<long long>(blockIdx.y) * 3 * plane + c * plane + static_cast<long long>(yy) * out_w + xx] =
        __fdiv_rn(static_cast<float>(v), 255.0f);
  }
}

// Same-size frames (the renders are already at the model's resolution, e.g. the uint8 host batches of
// bench.py's e2e leg): Pillow's resize is the identity there, so the kernel is ToTensor alone — HWC uint8 ->
// planar fp32 / 255.  Pure HBM byte work (3 B read + 12 B written per pixel): one CTA stages 1024 pixels
// with 16-byte loads and writes three 4 KB plane segments with 16-byte stores.
constexpr int TT_PIX = 1024;
__global__ void __launch_bounds__(256)
frames_u8_to_tensor_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ index,
                           float* __restrict__ dst, int hw, int chunks_per_frame) {
  pdl_sync();
  __shared__ __align__(16) uint8_t px[TT_PIX * 3];
  const int fi = blockIdx.x / chunks_per_frame, ch = blockIdx.x - fi * chunks_per_frame;
  const long long frame = index ? index[fi] : fi;
  const int p0 = ch * TT_PIX, np = min(TT_PIX, hw - p0);
  const uint8_t* g = src + (frame * hw + p0) * 3;
  if (np == TT_PIX && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    for (int i = threadIdx.x; i < TT_PIX * 3 / 16; i += 256)
      reinterpret_cast<uint4*>(px)[i] = __ldg(reinterpret_cast<const uint4*>(g) + i);
  } else {
    for (int i = threadIdx.x; i < np * 3; i += 256) px[i] = __ldg(g + i);
  }
  __syncthreads();
  float* o = dst + static_cast<long long>(fi) * 3 * hw + p0;
  for (int j = threadIdx.x; j < 3 * (TT_PIX / 4); j += 256) {
    const int c = j / (TT_PIX / 4), q = (j - c * (TT_PIX / 4)) * 4;
    if (q + 3 < np && ((hw | p0) & 3) == 0) {
      float4 v;
      v.x = __fdiv_rn(static_cast<float>(px[(q + 0) * 3 + c]), 255.0f);
      v.y = __fdiv_rn(static_cast<float>(px[(q + 1) * 3 + c]), 255.0f);
      v.z = __fdiv_rn(static_cast<float>(px[(q + 2) * 3 + c]), 255.0f);
      v.w = __fdiv_rn(static_cast<float>(px[(q + 3) * 3 + c]), 255.0f);
      *reinterpret_cast<float4*>(o + static_cast<long long>(c) * hw + q) = v;
    } else {
      for (int e = q; e < min(q + 4, np); ++e)
        o[static_cast<long long>(c) * hw + e] = __fdiv_rn(static_cast<float>(px[e * 3 + c]), 255.0f);
    }
  }
}

}  // namespace
}  // namespace mmdyn

using namespace mmdyn;

extern "C" int mmdyn_resize_table_ints(int in_h, int in_w, int out_h, int out_w) {
  if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return -1;
  return TABLE_HEADER + out_w * (2 + axis_ksize(in_w, out_w)) + out_h * (2 + axis_ksize(in_h, out_h));
}

extern "C" int mmdyn_resize_table(int in_h, int in_w, int out_h, int out_w, int32_t* table_host, int n_ints) {
  MMDYN_REQUIRE(table_host && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "resize_table: bad arguments");
  const int ksx = axis_ksize(in_w, out_w), ksy = axis_ksize(in_h, out_h);
  MMDYN_REQUIRE(ksx <= 64 && ksy <= 64, "resize_table: down-scaling by more than 31x is not supported (ksize %d/%d)",
                ksx, ksy);
  MMDYN_REQUIRE(n_ints >= mmdyn_resize_table_ints(in_h, in_w, out_h, out_w), "resize_table: table too small");
  table_host[0] = ksx; table_host[1] = ksy; table_host[2] = in_h; table_host[3] = in_w;
  table_host[4] = out_h; table_host[5] = out_w; table_host[6] = 0; table_host[7] = 0;
  int32_t* bx = table_host + TABLE_HEADER;
  int32_t* kx = bx + out_w * 2;
  int32_t* by = kx + out_w * ksx;
  int32_t* ky = by + out_h * 2;
  axis_coeffs(in_w, out_w, ksx, bx, kx);
  axis_coeffs(in_h, out_h, ksy, by, ky);
  return MMDYN_OK;
}

extern "C" int mmdyn_frames_u8_to_f32(const void* frames_u8, const long long* index, const int32_t* table_dev,
                                      float* out_nchw, int n, int in_h, int in_w, int out_h, int out_w,
                                      void* stream) {
  MMDYN_REQUIRE(frames_u8 && table_dev && out_nchw && n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0,
                "frames_u8_to_f32: bad arguments");
  MMDYN_REQUIRE(n <= 65535, "frames_u8_to_f32: at most 65535 frames per call (n=%d)", n);
  if (in_h == out_h && in_w == out_w) {  // no resampling: ToTensor only
    const int hw = in_h * in_w, chunks = (hw + TT_PIX - 1) / TT_PIX;
    MMDYN_LAUNCH((frames_u8_to_tensor_kernel), n * chunks, 256, 0, static_cast<cudaStream_t>(stream), 
        static_cast<const uint8_t*>(frames_u8), index, out_nchw, hw, chunks);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    MMDYN_CHECK_CUDA(cudaGetLastError());
    return MMDYN_OK;
  }
  const int ksy = axis_ksize(in_h, out_h);
  const int row_pitch = (in_w * 3 + 15) & ~15;
  {
    // R output rows per CTA (see frames_u8_to_f32_rows_kernel): the largest R in {8, 4, 2} whose staging fits (16 measured
    // slower: 0.92 vs 0.77 ms for 4096 frames — fewer resident CTAs)
    static const bool one_row = std::getenv("MMDYN_FRAMES_ONE_ROW") != nullptr;
    const int ksx = axis_ksize(in_w, out_w);
    const double scale = in_h > out_h ? static_cast<double>(in_h) / out_h : 1.0;
    for (int R = 8; R >= 2 && !one_row; R >>= 1) {
      if (R > out_h) continue;
      const int rows_max = in_h != out_h ? static_cast<int>(std::ceil((R - 1) * scale)) + ksy + 1 : R;
      const size_t smem_r = ((static_cast<size_t>(rows_max) * (row_pitch + out_w * 3) + 15) & ~size_t(15)) +
                            static_cast<size_t>(out_w) * (2 + ksx) * 4;
      if (smem_r > 96 * 1024) continue;
      static size_t configured_r = 0;
      if (smem_r > 48 * 1024 && smem_r > configured_r) {
        MMDYN_CHECK_CUDA(cudaFuncSetAttribute(frames_u8_to_f32_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(smem_r)));
        configured_r = smem_r;
      }
      MMDYN_LAUNCH((frames_u8_to_f32_rows_kernel), dim3((out_h + R - 1) / R, n), 256, smem_r, static_cast<cudaStream_t>(stream),
                   static_cast<const uint8_t*>(frames_u8), index, table_dev, out_nchw, static_cast<long long>(in_h) * in_w * 3, R,
                   rows_max);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      MMDYN_CHECK_CUDA(cudaGetLastError());
      return MMDYN_OK;
    }
  }
  const size_t smem = static_cast<size_t>(ksy) * (row_pitch + out_w * 3);
  MMDYN_REQUIRE(smem <= 200 * 1024, "frames_u8_to_f32: %zu bytes of shared memory needed (image too wide)", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MMDYN_CHECK_CUDA(cudaFuncSetAttribute(frames_u8_to_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    configured = smem;
  }
  MMDYN_LAUNCH((frames_u8_to_f32_kernel), dim3(out_h, n), 256, smem, static_cast<cudaStream_t>(stream), 
      static_cast<const uint8_t*>(frames_u8), index, table_dev, out_nchw,
      static_cast<long long>(in_h) * in_w * 3);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  MMDYN_CHECK_CUDA(cudaGetLastError());
  return MMDYN_OK;
}
