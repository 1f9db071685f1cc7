// Layout and elementwise helpers around the NHWC bf16 activation format, plus the two "thin" convolutions
// (3-channel stem, 3/6-channel head) that are too narrow for a UMMA tile and run on CUDA cores.
// All of these are HBM/L2-bound: 128-bit accesses, channel-contiguous thread mapping.
#include <cuda_bf16.h>

#include "host_common.h"

namespace pddm {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __low2float(h[i]);
    f[2 * i + 1] = __high2float(h[i]);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

static int grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  const long long cap = static_cast<long long>(device_info().sm_count > 0 ? device_info().sm_count : 148) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
#define GRID_STRIDE(i, n)                                                                        \
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (n);     \
       i += static_cast<long long>(gridDim.x) * blockDim.x)

// ---------------------------------------------------------------------------------- NCHW <-> NHWC
template <typename TOut>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int C, int HW) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + static_cast<size_t>(b) * C * HW;
  TOut* d = dst + static_cast<size_t>(b) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? s[static_cast<size_t>(c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) d[static_cast<size_t>(p) * C + c] = static_cast<TOut>(tile[threadIdx.x][i]);
  }
}
template <typename TIn>
__global__ void nhwc_to_nchw_kernel(const TIn* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn* s = src + static_cast<size_t>(b) * C * HW;
  float* d = dst + static_cast<size_t>(b) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? static_cast<float>(s[static_cast<size_t>(p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (p < HW && c < C) d[static_cast<size_t>(c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------- channel copy / add
__global__ void copy_channels_kernel(const bf16* __restrict__ src, int ld_src, bf16* __restrict__ dst, int ld_dst,
                                     long long M, int C8) {
  pdl_entry();
  const long long n = M * C8;
  GRID_STRIDE(i, n) {
    const long long m = i / C8;
    const int c = static_cast<int>(i - m * C8) * 8;
    *reinterpret_cast<uint4*>(dst + m * ld_dst + c) = *reinterpret_cast<const uint4*>(src + m * ld_src + c);
  }
}
__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ y,
                                long long n8) {
  pdl_entry();
  GRID_STRIDE(i, n8) {
    float fa[8], fb[8];
    unpack8(a[i], fa);
    unpack8(b[i], fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    y[i] = pack8(fa);
  }
}

// ---------------------------------------------------------------------------------- upsample / phases
__global__ void upsample2x_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int H, int W, int C8) {
  pdl_entry();
  const long long n = static_cast<long long>(B) * 2 * H * 2 * W * C8;
  GRID_STRIDE(i, n) {
    const int c = static_cast<int>(i % C8);
    long long r = i / C8;
    const int x = static_cast<int>(r % (2 * W)); r /= 2 * W;
    const int y = static_cast<int>(r % (2 * H));
    const int b = static_cast<int>(r / (2 * H));
    reinterpret_cast<uint4*>(dst)[i] =
        reinterpret_cast<const uint4*>(src)[((static_cast<long long>(b) * H + (y >> 1)) * W + (x >> 1)) * C8 + c];
  }
}
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ gd, bf16* __restrict__ gs, int B, int H, int W,
                                      int C8) {
  pdl_entry();
  const long long n = static_cast<long long>(B) * H * W * C8;
  GRID_STRIDE(i, n) {
    const int c = static_cast<int>(i % C8);
    long long r = i / C8;
    const int x = static_cast<int>(r % W); r /= W;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        unpack8(reinterpret_cast<const uint4*>(
                    gd)[((static_cast<long long>(b) * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C8 + c], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    reinterpret_cast<uint4*>(gs)[i] = pack8(acc);
  }
}
// split: dst[p][b][i][j] = src[b][2i + (p>>1)][2j + (p&1)] ; merge is the inverse (iterate over the full-res side)
template <bool SPLIT>
__global__ void phase_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int H, int W, int C8) {
  pdl_entry();
  const int H2 = H / 2, W2 = W / 2;
  const long long n = static_cast<long long>(B) * H * W * C8;
  GRID_STRIDE(i, n) {  // i indexes the full-resolution tensor
    const int c = static_cast<int>(i % C8);
    long long r = i / C8;
    const int x = static_cast<int>(r % W); r /= W;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    const int p = ((y & 1) << 1) | (x & 1);
    const long long j = (((static_cast<long long>(p) * B + b) * H2 + (y >> 1)) * W2 + (x >> 1)) * C8 + c;
    if (SPLIT) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(src)[i];
    else reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[j];
  }
}

// ---------------------------------------------------------------------------------- SiLU on vectors
__global__ void silu_kernel(const float* __restrict__ x, void* __restrict__ y, int y_dtype, long long n) {
  pdl_entry();
  GRID_STRIDE(i, n) {
    const float v = x[i];
    const float s = v / (1.f + expf(-v));
    if (y_dtype == PDDM_BF16) static_cast<bf16*>(y)[i] = __float2bfloat16(s);
    else static_cast<float*>(y)[i] = s;
  }
}
__global__ void silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                long long n) {
  pdl_entry();
  GRID_STRIDE(i, n) {
    const float v = x[i];
    const float sg = 1.f / (1.f + expf(-v));
    dx[i] = dy[i] * sg * (1.f + v * (1.f - sg));
  }
}

// SiLU on bf16 feature maps (the reference's stand-alone nn.SiLU between GroupNorm and conv, src/modules/unet.py:146-150;
// this framework's own UNet fuses it into GroupNorm): 8 values per thread, fp32 math.  n8 = number of 16-byte groups.
__global__ void silu_map_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8) {
  pdl_entry();
  GRID_STRIDE(i, n8) {
    uint4 v = x[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h[k]);
      f.x = f.x / (1.f + __expf(-f.x));
      f.y = f.y / (1.f + __expf(-f.y));
      h[k] = __floats2bfloat162_rn(f.x, f.y);
    }
    y[i] = v;
  }
}
__global__ void silu_map_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                    long long n8) {
  pdl_entry();
  GRID_STRIDE(i, n8) {
    uint4 v = x[i], g = dy[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]), d = __bfloat1622float2(gh[k]);
      const float s0 = 1.f / (1.f + __expf(-f.x)), s1 = 1.f / (1.f + __expf(-f.y));
      h[k] = __floats2bfloat162_rn(d.x * s0 * (1.f + f.x * (1.f - s0)), d.y * s1 * (1.f + f.y * (1.f - s1)));
    }
    dx[i] = v;
  }
}

// ---------------------------------------------------------------------------------- dtype conversion
template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
  pdl_entry();
  GRID_STRIDE(i, n) y[i] = static_cast<TO>(static_cast<float>(x[i]));
}

// ---------------------------------------------------------------------------------- column sums
// Deterministic two-stage reduction (no atomics): stage 1, grid (gx, segs): a block of C/8 vector columns x nlanes
// row lanes sums its slab of rows, folds the lanes through shared memory in lane order and writes one partial row;
// stage 2 folds the gx partial rows of each segment in index order.  grid.y = independent row segments (per-sample
// sums).  With gx == 1 stage 1 writes the result directly.
__global__ void colsum_partial_kernel(const bf16* __restrict__ x, int ld, long long rows_per_seg, long long rows_per_blk,
                                      int C8, float* __restrict__ part, float* __restrict__ out, int out_ld,
                                      int accumulate) {
  pdl_entry();
  extern __shared__ float sh[];  // [nlanes][C]
  const int seg = blockIdx.y, C = C8 * 8;
  const bf16* xs = x + static_cast<size_t>(seg) * rows_per_seg * ld;
  const int cv = threadIdx.x % C8, lane_r = threadIdx.x / C8, nlanes = blockDim.x / C8;
  const long long r0 = blockIdx.x * rows_per_blk, r1 = min(rows_per_seg, r0 + rows_per_blk);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0 + lane_r; r < r1; r += 4LL * nlanes) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * nlanes < r1) v[u] = *reinterpret_cast<const uint4*>(xs + (r + u * nlanes) * ld + cv * 8);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * nlanes < r1) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
  }
  float* d = sh + lane_r * C + cv * 8;
  *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = 0.f;
    for (int l = 0; l < nlanes; ++l) v += sh[l * C + c];
    if (gridDim.x == 1) {
      float* o = out + static_cast<size_t>(seg) * out_ld + c;
      *o = accumulate ? *o + v : v;
    } else {
      part[(static_cast<size_t>(seg) * gridDim.x + blockIdx.x) * C + c] = v;
    }
  }
}
// stage 2: out[seg][c] (+)= sum_k part[seg][k][c], k in index order; block = 32 columns x 8 partial lanes
__global__ void colsum_fold_kernel(const float* __restrict__ part, int gx, int C, float* __restrict__ out, int out_ld,
                                   int accumulate) {
  pdl_entry();
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, ly = threadIdx.y, seg = blockIdx.y;
  float a = 0.f;
  if (c < C)
    for (int k = ly; k < gx; k += 8) a += part[(static_cast<size_t>(seg) * gx + k) * C + c];
  sh[ly][threadIdx.x] = a;
  __syncthreads();
  if (ly == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += sh[k][threadIdx.x];
    float* o = out + static_cast<size_t>(seg) * out_ld + c;
    *o = accumulate ? *o + a : a;
  }
}

// ---------------------------------------------------------------------------------- weight packing
__global__ void pack_w_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int ntaps,
                              int mode, int Cout_p, int Cin_p) {
  pdl_entry();
  const long long n = static_cast<long long>(Cout_p) * Cin_p * ntaps;
  GRID_STRIDE(i, n) {  // i indexes dst (padded); padding rows / columns are written as zeros
    int co, ci, tap;
    if (mode == 0) {  // dst[co][tap][ci]
      ci = static_cast<int>(i % Cin_p);
      tap = static_cast<int>((i / Cin_p) % ntaps);
      co = static_cast<int>(i / (static_cast<long long>(Cin_p) * ntaps));
    } else {  // dst[ci][ntaps-1-tap][co]
      co = static_cast<int>(i % Cout_p);
      tap = ntaps - 1 - static_cast<int>((i / Cout_p) % ntaps);
      ci = static_cast<int>(i / (static_cast<long long>(Cout_p) * ntaps));
    }
    const float v = (co < Cout && ci < Cin) ? w[(static_cast<long long>(co) * Cin + ci) * ntaps + tap] : 0.f;
    dst[i] = __float2bfloat16(v);
  }
}

// all weights of a model in ONE launch: block i packs tile blocks[i].y of tensor blocks[i].x.  A tile is 32 output
// channels x 32 input channels x all taps: the fp32 source rows are read coalesced into shared memory (for a fixed
// output channel, 32 input channels x ntaps floats are contiguous), then gathered out of shared memory (odd strides:
// conflict-free) into 16-byte bf16 stores in either pack order.  (The first version read the source with a stride of
// ntaps floats and wrote single bf16 values: one LSU wavefront per element, 413 us per step for the CIFAR UNet.)
struct PackDesc {
  const float* src;
  bf16* dst;
  int Cout, Cin, ntaps, mode, Cout_p, Cin_p;
  int ld_dst, reserved;  // ld_dst: elements between destination rows (0 = dense)
  bf16* dst2;            // optional pack of the other mode (same paddings, dense) from the same source tile
};
constexpr int kPackTile = 32;
__global__ void __launch_bounds__(256) pack_multi_kernel(const PackDesc* __restrict__ descs,
                                                         const int2* __restrict__ blocks) {
  pdl_entry();
  extern __shared__ float tile[];  // [32][32 * ntaps + 1]
  const int2 blk = blocks[blockIdx.x];
  const PackDesc d = descs[blk.x];
  const int nt = d.ntaps;
  const int n_ci_blk = (d.Cin_p + kPackTile - 1) / kPackTile;
  const int co0 = (blk.y / n_ci_blk) * kPackTile, ci0 = (blk.y % n_ci_blk) * kPackTile;
  const int row_len = kPackTile * nt, pitch = row_len + 1;
  const int valid_len = max(0, min(kPackTile, d.Cin - ci0)) * nt;  // contiguous floats per source row
  // the 32 source rows of the tile are contiguous runs of 32 * ntaps floats: 16-byte loads (nine per thread for a
  // 3x3 kernel, all independent) keep enough bytes in flight to stream at HBM rate; rows that are not 16-byte
  // aligned or are cut by the channel count (stem, padded operands) take the scalar path
  const bool vec4 = (row_len & 3) == 0 && valid_len == row_len && ((static_cast<long long>(d.Cin) * nt) & 3) == 0 &&
                    ((reinterpret_cast<uintptr_t>(d.src) & 15) == 0);
  if (vec4) {
    const int row4 = row_len >> 2;
    for (int idx = threadIdx.x; idx < kPackTile * row4; idx += blockDim.x) {
      const int r = idx / row4, k = (idx - r * row4) * 4;
      const int co = co0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < d.Cout)
        v = __ldg(reinterpret_cast<const float4*>(d.src + (static_cast<long long>(co) * d.Cin + ci0) * nt + k));
      float* t = tile + r * pitch + k;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else {
    for (int idx = threadIdx.x; idx < kPackTile * row_len; idx += blockDim.x) {
      const int r = idx / row_len, k = idx - r * row_len;
      const int co = co0 + r;
      tile[r * pitch + k] =
          (co < d.Cout && k < valid_len) ? d.src[(static_cast<long long>(co) * d.Cin + ci0) * nt + k] : 0.f;
    }
  }
  __syncthreads();
  // work item = one 16-byte store: (row of the destination tile, tap, piece of 8 channels)
  const int items = kPackTile * nt * (kPackTile / 8);
  for (int pass = 0; pass < 2; ++pass) {
    bf16* dst = pass == 0 ? d.dst : d.dst2;
    if (!dst) break;
    const int mode = pass == 0 ? d.mode : 1 - d.mode;
    const int ld = pass == 0 ? d.ld_dst : 0;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int piece = it & 3, tap = (it >> 2) % nt, r = (it >> 2) / nt;
      float v[8];
      long long dst_off;
      if (mode == 0) {  // dst[co][tap][ci]: r = output channel, piece = 8 input channels
        const int co = co0 + r, ci = ci0 + piece * 8;
        if (co >= d.Cout_p || ci >= d.Cin_p) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tile[r * pitch + (piece * 8 + j) * nt + tap];
        dst_off = ld ? static_cast<long long>(co) * ld + tap * d.Cin_p + ci
                     : (static_cast<long long>(co) * nt + tap) * d.Cin_p + ci;
      } else {  // dst[ci][ntaps-1-tap][co]: r = input channel, piece = 8 output channels
        const int ci = ci0 + r, co = co0 + piece * 8;
        if (ci >= d.Cin_p || co >= d.Cout_p) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tile[(piece * 8 + j) * pitch + r * nt + tap];
        dst_off = ld ? static_cast<long long>(ci) * ld + (nt - 1 - tap) * d.Cout_p + co
                     : (static_cast<long long>(ci) * nt + (nt - 1 - tap)) * d.Cout_p + co;
      }
      *reinterpret_cast<uint4*>(dst + dst_off) = pack8(v);
    }
  }
}
// out[c] = sum_m x[m, c] for a small fp32 matrix (fixed order); block = 32 columns x 8 row lanes
__global__ void colsum_f32_kernel(const float* __restrict__ x, int M, int C, float* __restrict__ out) {
  pdl_entry();
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, rl = threadIdx.y;
  float a = 0.f;
  if (c < C)
    for (int m = rl; m < M; m += 8) a += x[static_cast<size_t>(m) * C + c];
  sh[rl][threadIdx.x] = a;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) a += sh[k][threadIdx.x];
    out[c] = a;
  }
}

// out[b, c] (+)= sum_hw x[b, hw, c]: grid (ceil(C/64), B), 256 threads = 32 row lanes x 8 units of 8 channels
__global__ void __launch_bounds__(256) colsum_rows_kernel(const bf16* __restrict__ x, int ldx, int HW, int C,
                                                          float* __restrict__ out, int ld_out, int accumulate) {
  pdl_entry();
  __shared__ float sh[32][65];
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int u = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c = c0 + u * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < C) {
    const bf16* xp = x + (static_cast<size_t>(b) * HW) * ldx + c;
    for (int r = lane; r < HW; r += 128) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r + 32 * k < HW) v[k] = *reinterpret_cast<const uint4*>(xp + static_cast<size_t>(r + 32 * k) * ldx);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r + 32 * k < HW) {
          float f[8];
          unpack8(v[k], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[lane][u * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float v = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) v += sh[l][threadIdx.x];
    float* o = out + static_cast<size_t>(b) * ld_out + c0 + threadIdx.x;
    *o = accumulate ? *o + v : v;
  }
}
// dst[j] = sum_b ps[b, src_of[j]] (fixed order, 8 loads in flight per thread)
__global__ void batch_fold_kernel(const float* __restrict__ ps, int ld, int B, const int* __restrict__ src_of, int n,
                                  float* __restrict__ dst) {
  pdl_entry();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int col = src_of ? src_of[j] : j;
  const float* p = ps + col;
  float s = 0.f;
  int b = 0;
  for (; b + 8 <= B; b += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(p + static_cast<size_t>(b + k) * ld);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; b < B; ++b) s += __ldg(p + static_cast<size_t>(b) * ld);
  dst[j] = s;
}
__global__ void convert_rows_kernel(const float* __restrict__ src, int ld_src, bf16* __restrict__ dst, int ld_dst,
                                    int rows, int cols8) {
  pdl_entry();
  const long long n = static_cast<long long>(rows) * cols8;
  GRID_STRIDE(i, n) {
    const int r = static_cast<int>(i / cols8), c = static_cast<int>(i % cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * ld_src + c);
    const float4 b = *reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * ld_src + c + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * ld_dst + c) = pack8(v);
  }
}

// ---------------------------------------------------------------------------------- thin-conv helpers
// The 3-channel stem and the 3/6-channel head are run on the tensor cores as well: the stem through an im2col of
// the (tiny) model input to a [pixels, 32] bf16 patch matrix, the head by zero-padding its output channels.
// out[b,h,w, ci*9 + tap] = x[b,ci,h+dh,w+dw]  (NCHW fp32 -> [B,H,W,Kp] bf16, zero outside the image and for k >= Cin*9)
__global__ void im2col3x3_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                 int Kp) {
  pdl_entry();
  const int K8 = Kp / 8;
  const long long n = static_cast<long long>(B) * H * W * K8;
  GRID_STRIDE(i, n) {
    const int kg = static_cast<int>(i % K8);
    long long r = i / K8;
    const int xw = static_cast<int>(r % W); r /= W;
    const int yh = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kg * 8 + j;
      const int ci = k / 9, t = k - ci * 9;
      const int hh = yh + t / 3 - 1, ww = xw + t % 3 - 1;
      f[j] = (ci < Cin && hh >= 0 && hh < H && ww >= 0 && ww < W)
                 ? __ldg(x + ((static_cast<size_t>(b) * Cin + ci) * H + hh) * W + ww)
                 : 0.f;
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
}
// NCHW fp32 [B,C,HW] -> NHWC bf16 [B,HW,Cp] with channels >= C zero (Cp % 8 == 0)
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int B, int C, int HW,
                                        int Cp) {
  pdl_entry();
  const int C8 = Cp / 8;
  const long long n = static_cast<long long>(B) * HW * C8;
  GRID_STRIDE(i, n) {
    const int cg = static_cast<int>(i % C8);
    const long long r = i / C8;
    const int pix = static_cast<int>(r % HW), b = static_cast<int>(r / HW);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      f[j] = c < C ? __ldg(src + (static_cast<size_t>(b) * C + c) * HW + pix) : 0.f;
    }
    reinterpret_cast<uint4*>(dst)[i] = pack8(f);
  }
}
// first C channels of NHWC fp32 [B,HW,ld] -> NCHW fp32 [B,C,HW]
__global__ void nhwc_slice_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW,
                                          int ld) {
  pdl_entry();
  const long long n = static_cast<long long>(B) * C * HW;
  GRID_STRIDE(i, n) {
    const int pix = static_cast<int>(i % HW);
    const long long r = i / HW;
    const int c = static_cast<int>(r % C), b = static_cast<int>(r / C);
    dst[i] = src[(static_cast<size_t>(b) * HW + pix) * ld + c];
  }
}

}  // namespace pddm

using namespace pddm;
#define S(s_) static_cast<cudaStream_t>(s_)

extern "C" int pddm_nchw_to_nhwc(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t C, int32_t HW,
                                 pddm_stream_t s) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0) return PDDM_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  if (dst_dtype == PDDM_BF16) PdlLaunch(grid, block, 0, S(s))(nchw_to_nhwc_kernel<bf16>, src, static_cast<bf16*>(dst), C, HW);
  else PdlLaunch(grid, block, 0, S(s))(nchw_to_nhwc_kernel<float>, src, static_cast<float*>(dst), C, HW);
  return launch_status();
}
extern "C" int pddm_nhwc_to_nchw(const void* src, int32_t src_dtype, float* dst, int32_t B, int32_t C, int32_t HW,
                                 pddm_stream_t s) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0) return PDDM_ERR_BAD_ARG;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  if (src_dtype == PDDM_BF16) PdlLaunch(grid, block, 0, S(s))(nhwc_to_nchw_kernel<bf16>, static_cast<const bf16*>(src), dst, C, HW);
  else PdlLaunch(grid, block, 0, S(s))(nhwc_to_nchw_kernel<float>, static_cast<const float*>(src), dst, C, HW);
  return launch_status();
}
extern "C" int pddm_copy_channels(const void* src, int32_t ld_src, int32_t src_off, void* dst, int32_t ld_dst,
                                  int32_t dst_off, int64_t M, int32_t C, pddm_stream_t s) {
  if (!src || !dst || M <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8 || ld_src % 8 || ld_dst % 8 || src_off % 8 || dst_off % 8 || !aligned16(src) || !aligned16(dst))
    return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(M * (C / 8), 256), 256, 0, S(s))(copy_channels_kernel, static_cast<const bf16*>(src) + src_off, ld_src,
                                                                      static_cast<bf16*>(dst) + dst_off, ld_dst, M, C / 8);
  return launch_status();
}
extern "C" int pddm_add_bf16(const void* a, const void* b, void* y, int64_t n, pddm_stream_t s) {
  if (!a || !b || !y || n <= 0) return PDDM_ERR_BAD_ARG;
  if (n % 8 || !aligned16(a) || !aligned16(b) || !aligned16(y)) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(n / 8, 256), 256, 0, S(s))(add_bf16_kernel, static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                          static_cast<uint4*>(y), n / 8);
  return launch_status();
}
extern "C" int pddm_upsample2x(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t s) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * 4 * H * W * (C / 8), 256), 256, 0, S(s))(upsample2x_kernel, 
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), B, H, W, C / 8);
  return launch_status();
}
extern "C" int pddm_upsample2x_bwd(const void* gd, void* gs, int32_t B, int32_t H, int32_t W, int32_t C,
                                   pddm_stream_t s) {
  if (!gd || !gs || B <= 0 || H <= 0 || W <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * H * W * (C / 8), 256), 256, 0, S(s))(upsample2x_bwd_kernel, 
      static_cast<const bf16*>(gd), static_cast<bf16*>(gs), B, H, W, C / 8);
  return launch_status();
}
extern "C" int pddm_phase_split(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t s) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8 || H % 2 || W % 2) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * H * W * (C / 8), 256), 256, 0, S(s))(phase_kernel<true>, 
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), B, H, W, C / 8);
  return launch_status();
}
extern "C" int pddm_phase_merge(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t s) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8 || H % 2 || W % 2) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * H * W * (C / 8), 256), 256, 0, S(s))(phase_kernel<false>, 
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), B, H, W, C / 8);
  return launch_status();
}
extern "C" int pddm_convert(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, pddm_stream_t s) {
  if (!x || !y || n <= 0) return PDDM_ERR_BAD_ARG;
  const int g = grid_for(n, 256);
  if (x_dtype == PDDM_F32 && y_dtype == PDDM_BF16)
    PdlLaunch(g, 256, 0, S(s))(convert_kernel<float, bf16>, static_cast<const float*>(x), static_cast<bf16*>(y), n);
  else if (x_dtype == PDDM_BF16 && y_dtype == PDDM_F32)
    PdlLaunch(g, 256, 0, S(s))(convert_kernel<bf16, float>, static_cast<const bf16*>(x), static_cast<float*>(y), n);
  else
    return PDDM_ERR_UNSUPPORTED;
  return launch_status();
}
extern "C" int pddm_silu(const float* x, void* y, int32_t y_dtype, int64_t n, pddm_stream_t s) {
  if (!x || !y || n <= 0) return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(n, 256), 256, 0, S(s))(silu_kernel, x, y, y_dtype, n);
  return launch_status();
}
extern "C" int pddm_silu_map(const void* x, void* y, int64_t n, pddm_stream_t s) {
  if (!x || !y || n <= 0 || n % 8 != 0 || !aligned16(x) || !aligned16(y)) return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(n / 8, 256), 256, 0, S(s))(silu_map_kernel, static_cast<const uint4*>(x), static_cast<uint4*>(y),
                                                static_cast<long long>(n / 8));
  return launch_status();
}
extern "C" int pddm_silu_map_bwd(const void* x, const void* dy, void* dx, int64_t n, pddm_stream_t s) {
  if (!x || !dy || !dx || n <= 0 || n % 8 != 0 || !aligned16(x) || !aligned16(dy) || !aligned16(dx))
    return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(n / 8, 256), 256, 0, S(s))(silu_map_bwd_kernel, static_cast<const uint4*>(x),
                                                static_cast<const uint4*>(dy), static_cast<uint4*>(dx),
                                                static_cast<long long>(n / 8));
  return launch_status();
}
extern "C" int pddm_silu_bwd(const float* x, const float* dy, float* dx, int64_t n, pddm_stream_t s) {
  if (!x || !dy || !dx || n <= 0) return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(n, 256), 256, 0, S(s))(silu_bwd_kernel, x, dy, dx, n);
  return launch_status();
}

struct ColsumGeom {
  int C8, threads, nlanes, gx;
  long long rows_per_blk;
};
static int colsum_geom(long long rows_per_seg, int segs, int C, ColsumGeom* g) {
  if (C % 8 || C > 8192 || C <= 0 || rows_per_seg <= 0 || segs <= 0) return PDDM_ERR_UNSUPPORTED;
  g->C8 = C / 8;
  g->nlanes = 256 / g->C8 > 0 ? 256 / g->C8 : 1;
  g->threads = g->C8 * g->nlanes;
  if (g->threads > 1024) return PDDM_ERR_UNSUPPORTED;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  long long gx = (4LL * sms + segs - 1) / segs;                                  // ~4 blocks per SM overall
  const long long max_gx = (rows_per_seg + 4LL * g->nlanes - 1) / (4LL * g->nlanes);  // >= 4 rows per lane
  if (gx > max_gx) gx = max_gx;
  if (gx < 1) gx = 1;
  g->rows_per_blk = (rows_per_seg + gx - 1) / gx;
  g->gx = static_cast<int>((rows_per_seg + g->rows_per_blk - 1) / g->rows_per_blk);
  return PDDM_OK;
}
static size_t colsum_ws_bytes(long long rows_per_seg, int segs, int C) {
  ColsumGeom g;
  const int cw = C > 4096 ? 4096 : C;  // wide matrices are reduced in column panels that reuse the workspace
  if (colsum_geom(rows_per_seg, segs, cw, &g) != PDDM_OK) return 0;
  return g.gx > 1 ? static_cast<size_t>(segs) * g.gx * cw * sizeof(float) : 16;
}
static int colsum_launch(const void* x, int ld, long long rows_per_seg, int segs, int C, float* out, int out_ld,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ld % 8 || !aligned16(x)) return PDDM_ERR_UNSUPPORTED;
  ColsumGeom g;
  int rc = colsum_geom(rows_per_seg, segs, C, &g);
  if (rc) return rc;
  const size_t smem = static_cast<size_t>(g.nlanes) * C * sizeof(float);
  if (g.gx > 1 && (!ws || ws_bytes < static_cast<size_t>(segs) * g.gx * C * sizeof(float))) return PDDM_ERR_WORKSPACE;
  if (smem > 48 * 1024) {
    if (ensure_smem_optin(reinterpret_cast<const void*>(colsum_partial_kernel))) return PDDM_ERR_CUDA;
    if (smem > 64 * 1024) return PDDM_ERR_UNSUPPORTED;
  }
  PdlLaunch(dim3(g.gx, segs), g.threads, smem, s)(colsum_partial_kernel, static_cast<const bf16*>(x), ld, rows_per_seg,
                                                   g.rows_per_blk, g.C8, static_cast<float*>(ws), out, out_ld,
                                                   accumulate);
  if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  if (g.gx > 1)
    PdlLaunch(dim3((C + 31) / 32, segs), dim3(32, 8), 0, s)(colsum_fold_kernel, static_cast<const float*>(ws), g.gx, C,
                                                            out, out_ld, accumulate);
  return launch_status();
}
extern "C" size_t pddm_colsum_workspace(int64_t rows_per_segment, int32_t segments, int32_t C) {
  return colsum_ws_bytes(rows_per_segment, segments, C);
}
extern "C" int pddm_colsum(const void* x, int32_t ld, int64_t M, int32_t C, float* out, int32_t accumulate,
                           void* workspace, size_t workspace_bytes, pddm_stream_t s) {
  if (!x || !out || M <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  // wide matrices (e.g. the batched timestep-embedding projection) are reduced in column panels
  for (int c0 = 0; c0 < C; c0 += 4096) {
    const int cw = C - c0 < 4096 ? C - c0 : 4096;
    int rc = colsum_launch(static_cast<const bf16*>(x) + c0, ld, M, 1, cw, out + c0, cw, accumulate, workspace,
                           workspace_bytes, S(s));
    if (rc) return rc;
  }
  return PDDM_OK;
}
extern "C" int pddm_colsum_per_sample(const void* x, int32_t B, int32_t HW, int32_t C, float* out, void* workspace,
                                      size_t workspace_bytes, pddm_stream_t s) {
  if (!x || !out || B <= 0 || HW <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  return colsum_launch(x, C, HW, B, C, out, C, 0, workspace, workspace_bytes, S(s));
}
extern "C" int pddm_pack_conv_weight(const float* w, void* dst, int32_t Cout, int32_t Cin, int32_t ntaps, int32_t mode,
                                     int32_t Cout_pad, int32_t Cin_pad, pddm_stream_t s) {
  if (!w || !dst || Cout <= 0 || Cin <= 0 || ntaps <= 0 || mode < 0 || mode > 1 || Cout_pad < Cout || Cin_pad < Cin)
    return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(static_cast<long long>(Cout_pad) * Cin_pad * ntaps, 256), 256, 0, S(s))(pack_w_kernel, 
      w, static_cast<bf16*>(dst), Cout, Cin, ntaps, mode, Cout_pad, Cin_pad);
  return launch_status();
}
extern "C" int pddm_pack_weights_multi(const void* descs, const void* blocks, int32_t nblocks, int32_t max_ntaps,
                                       pddm_stream_t s) {
  if (!descs || !blocks || nblocks <= 0 || max_ntaps <= 0 || max_ntaps > PDDM_MAX_TAPS) return PDDM_ERR_BAD_ARG;
  const size_t smem = static_cast<size_t>(kPackTile) * (kPackTile * max_ntaps + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    if (ensure_smem_optin(reinterpret_cast<const void*>(pack_multi_kernel))) return PDDM_ERR_CUDA;
    if (smem > 96 * 1024) return PDDM_ERR_UNSUPPORTED;
  }
  PdlLaunch(nblocks, 256, smem, S(s))(pack_multi_kernel, static_cast<const PackDesc*>(descs),
                                       static_cast<const int2*>(blocks));
  return launch_status();
}
extern "C" int pddm_colsum_f32(const float* x, int32_t M, int32_t C, float* out, pddm_stream_t s) {
  if (!x || !out || M <= 0 || C <= 0) return PDDM_ERR_BAD_ARG;
  PdlLaunch((C + 31) / 32, dim3(32, 8), 0, S(s))(colsum_f32_kernel, x, M, C, out);
  return launch_status();
}
extern "C" int pddm_colsum_rows(const void* x, int32_t ldx, int32_t B, int32_t HW, int32_t C, float* out,
                                int32_t ld_out, int32_t accumulate, pddm_stream_t s) {
  if (!x || !out || B <= 0 || HW <= 0 || C <= 0 || ldx < C || ld_out < C) return PDDM_ERR_BAD_ARG;
  if (C % 8 || ldx % 8 || !aligned16(x)) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(dim3((C + 63) / 64, B), 256, 0, S(s))(colsum_rows_kernel, static_cast<const bf16*>(x), ldx, HW, C, out,
                                                  ld_out, accumulate);
  return launch_status();
}
extern "C" int pddm_batch_fold(const float* ps, int32_t ld, int32_t B, const int32_t* src_of, int32_t n, float* dst,
                               pddm_stream_t s) {
  if (!ps || !dst || ld <= 0 || B <= 0 || n <= 0) return PDDM_ERR_BAD_ARG;
  PdlLaunch((n + 127) / 128, 128, 0, S(s))(batch_fold_kernel, ps, ld, B, src_of, n, dst);
  return launch_status();
}
extern "C" int pddm_convert_rows(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows,
                                 int32_t cols, pddm_stream_t s) {
  if (!src || !dst || rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < cols) return PDDM_ERR_BAD_ARG;
  if (cols % 8 || ld_src % 4 || ld_dst % 8 || !aligned16(src) || !aligned16(dst)) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(rows) * (cols / 8), 256), 256, 0, S(s))(
      convert_rows_kernel, src, ld_src, static_cast<bf16*>(dst), ld_dst, rows, cols / 8);
  return launch_status();
}
extern "C" int pddm_im2col3x3(const float* x_nchw, void* out, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Kp,
                              pddm_stream_t s) {
  if (!x_nchw || !out || B <= 0 || Cin <= 0 || H <= 0 || W <= 0) return PDDM_ERR_BAD_ARG;
  if (Kp % 8 || Kp < Cin * 9) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * H * W * (Kp / 8), 256), 256, 0, S(s))(im2col3x3_kernel, 
      x_nchw, static_cast<bf16*>(out), B, Cin, H, W, Kp);
  return launch_status();
}
extern "C" int pddm_nchw_to_nhwc_padded(const float* src, void* dst, int32_t B, int32_t C, int32_t HW, int32_t Cp,
                                        pddm_stream_t s) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0) return PDDM_ERR_BAD_ARG;
  if (Cp % 8 || Cp < C) return PDDM_ERR_UNSUPPORTED;
  PdlLaunch(grid_for(static_cast<long long>(B) * HW * (Cp / 8), 256), 256, 0, S(s))(nchw_to_nhwc_pad_kernel, 
      src, static_cast<bf16*>(dst), B, C, HW, Cp);
  return launch_status();
}
// Sample post-processing (src/modules/fid_score.py:15-27 -> src/datasets/data.py:108-128 unnormalize(clip=True) ->
// the image writer's (255 * A).astype(uint8)): fp32 NCHW model-space samples -> uint8 NHWC pixels in ONE pass, so a
// mini-batch leaves the device as B*H*W*C bytes instead of 4x that in fp32.  The affine runs in double like the
// reference's numpy expression (float32 image x float64 std/mean arrays), truncation like astype(uint8).
__global__ void images_to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int Cc, int HW,
                                       const double* __restrict__ mean, const double* __restrict__ stdv) {
  pdl_entry();
  const long long n = static_cast<long long>(B) * HW * Cc;
  GRID_STRIDE(i, n) {  // i indexes the NHWC output
    const int c = static_cast<int>(i % Cc);
    const long long p = i / Cc;
    const int hw = static_cast<int>(p % HW);
    const long long b = p / HW;
    double v = static_cast<double>(x[(b * Cc + c) * HW + hw]);
    if (mean) v = v * stdv[c] + mean[c];
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    out[i] = static_cast<uint8_t>(255.0 * v);
  }
}
extern "C" int pddm_images_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int32_t B, int32_t C, int32_t HW,
                                    const double* mean, const double* stdv, pddm_stream_t s) {
  if (!x_nchw || !out_nhwc || B <= 0 || C <= 0 || HW <= 0 || ((mean == nullptr) != (stdv == nullptr)))
    return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(static_cast<long long>(B) * C * HW, 256), 256, 0, S(s))(images_to_uint8_kernel, x_nchw, out_nhwc, B,
                                                                           C, HW, mean, stdv);
  return launch_status();
}
extern "C" int pddm_nhwc_slice_to_nchw(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, int32_t ld,
                                       pddm_stream_t s) {
  if (!src || !dst || B <= 0 || C <= 0 || HW <= 0 || ld < C) return PDDM_ERR_BAD_ARG;
  PdlLaunch(grid_for(static_cast<long long>(B) * C * HW, 256), 256, 0, S(s))(nhwc_slice_to_nchw_kernel, src, dst, B, C, HW, ld);
  return launch_status();
}
