// QKVAttention (src/modules/unet.py:237-256) as fused tcgen05 kernels.
//
// Sequence length T = H*W <= 256, so one head's whole K and V sit in shared memory and a 128-query tile's score
// matrix sits in tensor memory: a single pass, no online-softmax rescaling.
//   forward  (CTA = sample x head x 128-query tile):  S = Q K^T (TMEM) -> fp32 softmax in registers, thread per
//            row -> P (bf16, swizzled smem) -> O = P V (TMEM) -> normalise, store; log-sum-exp saved.
//   backward (CTA = sample x head): per query tile S -> P, dP = dO V^T, dV += P^T dO, dS = P*(dP - D)*scale,
//            dQ = dS K, dK += dS^T Q; every product is a tcgen05.mma with accumulators in TMEM.
// Each [tokens x d] operand tile is loaded once by TMA as 64- (or 32-) channel swizzled chunks and is used both
// as a K-major operand (contraction over channels) and, re-described, as an MN-major operand (contraction over
// tokens); the P / dS tile likewise serves as A (K-major) and A^T (MN-major).
#include <cuda_bf16.h>

#include "host_common.h"
#include "ptx.cuh"

namespace pddm {

constexpr int kAttnParts = 2;  // threads per query row in the elementwise phases (each warp stays in its TMEM lane
                               // quadrant); measured on B200: 1 -> 2 is -17% (bwd) / -13% (fwd), 4 is slightly worse than 2

typedef __nv_bfloat16 bf16;

struct AttnArgs {
  const bf16* out;   // bwd: forward output [B,T,C]
  const bf16* dout;  // bwd
  bf16* y;           // fwd out [B,T,C]  /  bwd dqkv [B,T,3C]
  float* lse;        // [B, heads, T]
  int B, T, heads, d, C;
  int Tp;         // keys padded to a multiple of 32
  int cw, nck;    // channel chunk width (64 -> SW128, 32 -> SW64), chunks per head
  int nqt;        // query tiles of 128
  float scale_log2e, scale;
  uint32_t tmem_cols;
  uint32_t off_q, off_do, off_k, off_v, off_p, off_bar;  // smem offsets
  int pass;  // backward only: 0 = dQ, dK, dV in one launch; 1 = dQ + dV; 2 = dK (two launches when TMEM is short)
};

// ---- descriptor helpers -------------------------------------------------------------------------
// tile = nck chunks of [rows x cw] bf16, row pitch cw*2 bytes, chunk stride rows*cw*2 bytes.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int rows, int cw, int kk) {
  // contraction over channels: k-step kk covers channels [16kk, 16kk+16)
  const int ch = kk * 16;
  const uint32_t addr = tile + (ch / cw) * (rows * cw * 2) + (ch % cw) * 2;
  return make_smem_desc(addr, 16, 8 * cw * 2, cw == 64 ? kLayoutSW128 : kLayoutSW64);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int rows, int cw, int kk, int mn_chunk0) {
  // contraction over token rows: k-step kk covers rows [16kk, 16kk+16); MN runs over channels from chunk mn_chunk0
  const uint32_t addr = tile + mn_chunk0 * (rows * cw * 2) + kk * 16 * (cw * 2);
  return make_smem_desc(addr, rows * cw * 2, 8 * cw * 2, cw == 64 ? kLayoutSW128 : kLayoutSW64);
}

// write 32 fp32 values of row r, columns [c0, c0+32) into the [128 x Tp] bf16 tile of 64-column SW128 chunks
__device__ __forceinline__ void store_p32(uint8_t* ptile, int r, int c0, const float* v) {
  uint8_t* chunk = ptile + (c0 >> 6) * (128 * 128) + r * 128;
  const int u0 = (c0 & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint4 o;
    o.x = pack_bf16(v[u * 8 + 0], v[u * 8 + 1]);
    o.y = pack_bf16(v[u * 8 + 2], v[u * 8 + 3]);
    o.z = pack_bf16(v[u * 8 + 4], v[u * 8 + 5]);
    o.w = pack_bf16(v[u * 8 + 6], v[u * 8 + 7]);
    *reinterpret_cast<uint4*>(chunk + (((u0 + u) ^ (r & 7)) << 4)) = o;
  }
}
__device__ __forceinline__ void load_p32(const uint8_t* ptile, int r, int c0, float* v) {
  const uint8_t* chunk = ptile + (c0 >> 6) * (128 * 128) + r * 128;
  const int u0 = (c0 & 63) >> 3;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 o = *reinterpret_cast<const uint4*>(chunk + (((u0 + u) ^ (r & 7)) << 4));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[u * 8 + 2 * i] = __low2float(h[i]);
      v[u * 8 + 2 * i + 1] = __high2float(h[i]);
    }
  }
}
// P tile as K-major A (M = rows, K = columns): k-step kk covers columns [16kk, 16kk+16)
__device__ __forceinline__ uint64_t desc_p_kmajor(uint32_t ptile, int kk) {
  return make_smem_desc(ptile + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024, kLayoutSW128);
}
// P tile as MN-major A (M = columns [128*mt, +128), K = rows): k-step kk covers rows [16kk, 16kk+16)
__device__ __forceinline__ uint64_t desc_p_mnmajor(uint32_t ptile, int kk, int mt) {
  return make_smem_desc(ptile + (2 * mt) * (128 * 128) + kk * 2048, 128 * 128, 1024, kLayoutSW128);
}

// store one accumulator row (d columns at TMEM column `col`) scaled by `mul` as bf16 to dst[0..d)
__device__ __forceinline__ void store_row_from_tmem(uint32_t taddr_row, int col, int d, float mul, bf16* dst,
                                                    bool valid) {
  for (int c = 0; c < d; c += 32) {
    uint32_t r[32];
    tmem_ld32(taddr_row + col + c, r);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[u * 8 + 0]) * mul, __uint_as_float(r[u * 8 + 1]) * mul);
        o.y = pack_bf16(__uint_as_float(r[u * 8 + 2]) * mul, __uint_as_float(r[u * 8 + 3]) * mul);
        o.z = pack_bf16(__uint_as_float(r[u * 8 + 4]) * mul, __uint_as_float(r[u * 8 + 5]) * mul);
        o.w = pack_bf16(__uint_as_float(r[u * 8 + 6]) * mul, __uint_as_float(r[u * 8 + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + c + u * 8) = o;
      }
    }
  }
}

// =================================================================================================== forward
// 128*kAttnParts threads: kAttnParts threads per query row split the columns of the softmax and of the output drain (see the
// backward kernel); the row maximum and the row sum are combined through a small shared-memory exchange.
__global__ void __launch_bounds__(128 * kAttnParts)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x % a.nqt;
  const int bh = blockIdx.x / a.nqt;
  const int h = bh % a.heads, b = bh / a.heads;
  const int t0 = qt * 128;

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_launch_dependents();  // after the TMEM allocation (see conv_fwd.cu)
  pdl_wait();
  const uint32_t sQ = smem_u32(smem + a.off_q), sK = smem_u32(smem + a.off_k), sV = smem_u32(smem + a.off_v);
  const uint32_t sP = smem_u32(smem + a.off_p);

  if (tid == 0) {
    const uint32_t bytes = (128 + 2 * a.Tp) * a.d * 2;
    mbar_expect_tx(bar_load, bytes);
    const int cq = h * 3 * a.d;
    for (int c = 0; c < a.nck; ++c) {
      tma_load_4d(smem + a.off_q + c * (128 * a.cw * 2), &tmQ, bar_load, cq + c * a.cw, t0, b, 0);
      tma_load_4d(smem + a.off_k + c * (a.Tp * a.cw * 2), &tmKV, bar_load, cq + a.d + c * a.cw, 0, b, 0);
      tma_load_4d(smem + a.off_v + c * (a.Tp * a.cw * 2), &tmKV, bar_load, cq + 2 * a.d + c * a.cw, 0, b, 0);
    }
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, a.Tp, 0, 0);
    for (int kk = 0; kk < a.d / 16; ++kk)
      umma_bf16(tmem, desc_kmajor(sQ, 128, a.cw, kk), desc_kmajor(sK, a.Tp, a.cw, kk), idesc, kk > 0);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- softmax, two threads per query row (half = column range) ----
  const int row = tid & 127, part = tid >> 7;
  const uint32_t trow = tmem + (static_cast<uint32_t>(row) << 16);  // lane = row (warp w owns lanes 32(w%4)..)
  float* xch = reinterpret_cast<float*>(smem + a.off_bar + 64);  // [max | sum][part][128 rows]
  const int nch_s = a.Tp >> 5, nch_d = a.d >> 5;
  const int cs0 = nch_s * part / kAttnParts * 32, cs1 = nch_s * (part + 1) / kAttnParts * 32;
  const int cd0 = nch_d * part / kAttnParts * 32, cd1 = nch_d * (part + 1) / kAttnParts * 32;
  const int t = t0 + row;
  float mx = -INFINITY;
  for (int c = cs0; c < cs1; c += 32) {
    uint32_t r[32];
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < a.T) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  xch[part * 128 + row] = mx;
  __syncthreads();
  mx = xch[row];
#pragma unroll
  for (int k = 1; k < kAttnParts; ++k) mx = fmaxf(mx, xch[k * 128 + row]);
  float sum = 0.f;
  const float mxs = mx * a.scale_log2e;
  for (int c = cs0; c < cs1; c += 32) {
    uint32_t r[32];
    float p[32];
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      p[j] = (c + j < a.T) ? exp2f(__uint_as_float(r[j]) * a.scale_log2e - mxs) : 0.f;
      sum += p[j];
    }
    store_p32(smem + a.off_p, row, c, p);  // aliases Q|K, which the S MMA has finished reading
  }
  xch[(kAttnParts + part) * 128 + row] = sum;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  sum = xch[kAttnParts * 128 + row];
#pragma unroll
  for (int k = 1; k < kAttnParts; ++k) sum += xch[(kAttnParts + k) * 128 + row];
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, a.d, 0, 1);
    for (int kk = 0; kk < a.Tp / 16; ++kk)
      umma_bf16(tmem, desc_p_kmajor(sP, kk), desc_mnmajor(sV, a.Tp, a.cw, kk, 0), idesc, kk > 0);
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  const bool valid = t < a.T;
  bf16* dst = a.y + (static_cast<size_t>(b) * a.T + (valid ? t : 0)) * a.C + h * a.d;
  if (cd0 < cd1) store_row_from_tmem(trow, cd0, cd1 - cd0, 1.f / sum, dst + cd0, valid);
  if (valid && a.lse && part == 0) a.lse[(static_cast<size_t>(b) * a.heads + h) * a.T + t] = mx * a.scale + logf(sum);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, a.tmem_cols);
  }
}

// =================================================================================================== backward
// 128*kAttnParts threads: kAttnParts threads per query row (thread = row + 128*part) split the COLUMNS of every elementwise phase
// (P = exp(S - lse), dS = P*(dP - D), the dQ/dK/dV drains); a warp may only touch the TMEM lanes of its quadrant
// (warp % 4), which is exactly row / 32 for both halves.  With 128 threads these phases, not the five GEMMs, set the
// kernel's time.
__global__ void __launch_bounds__(128 * kAttnParts)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* bar_q = bar_kv + 1;
  uint64_t* bar_mma = bar_kv + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_kv + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7;
  const int h = blockIdx.x % a.heads, b = blockIdx.x / a.heads;
  const int ns = (a.Tp + 127) / 128;  // key tiles of 128
  const int r0w = a.Tp > a.d ? a.Tp : a.d;
  const bool do_v = a.pass != 2, do_k = a.pass != 1;  // dQ is produced together with dV
  const int col_dv = r0w, col_dk = (a.pass == 0) ? r0w + ns * a.d : r0w;

  if (tid == 0) {
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  pdl_launch_dependents();  // after the TMEM allocation (see conv_fwd.cu)
  pdl_wait();
  const uint32_t sQ = smem_u32(smem + a.off_q), sDO = smem_u32(smem + a.off_do), sK = smem_u32(smem + a.off_k),
                 sV = smem_u32(smem + a.off_v), sP = smem_u32(smem + a.off_p);
  const uint32_t trow = tmem + (static_cast<uint32_t>(row) << 16);
  // column ranges of this half: 32-column chunks of the [128 x Tp] score tile and of the d-wide accumulators
  const int nch_s = a.Tp >> 5, nch_d = a.d >> 5;
  const int cs0 = nch_s * part / kAttnParts * 32, cs1 = nch_s * (part + 1) / kAttnParts * 32;
  const int cd0 = nch_d * part / kAttnParts * 32, cd1 = nch_d * (part + 1) / kAttnParts * 32;
  const int cq = h * 3 * a.d;
  uint32_t mma_phase = 0;

  if (tid == 0) {
    mbar_expect_tx(bar_kv, 2 * a.Tp * a.d * 2);
    for (int c = 0; c < a.nck; ++c) {
      tma_load_4d(smem + a.off_k + c * (a.Tp * a.cw * 2), &tmKV, bar_kv, cq + a.d + c * a.cw, 0, b, 0);
      tma_load_4d(smem + a.off_v + c * (a.Tp * a.cw * 2), &tmKV, bar_kv, cq + 2 * a.d + c * a.cw, 0, b, 0);
    }
  }

  for (int qt = 0; qt < a.nqt; ++qt) {
    const int t0 = qt * 128;
    const int t = t0 + row;
    const bool valid = t < a.T;
    if (tid == 0) {
      mbar_expect_tx(bar_q, 2 * 128 * a.d * 2);
      for (int c = 0; c < a.nck; ++c) {
        tma_load_4d(smem + a.off_q + c * (128 * a.cw * 2), &tmQ, bar_q, cq + c * a.cw, t0, b, 0);
        tma_load_4d(smem + a.off_do + c * (128 * a.cw * 2), &tmDO, bar_q, h * a.d + c * a.cw, t0, b, 0);
      }
      if (qt == 0) mbar_wait(bar_kv, 0);
      mbar_wait(bar_q, qt & 1);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, a.Tp, 0, 0);
      for (int kk = 0; kk < a.d / 16; ++kk)
        umma_bf16(tmem, desc_kmajor(sQ, 128, a.cw, kk), desc_kmajor(sK, a.Tp, a.cw, kk), idesc, kk > 0);
      umma_commit(bar_mma);
    }
    // D[t] = sum_c dO[t,c] * O[t,c]  (overlaps the S MMA)
    float Dt = 0.f, lse = 0.f;
    if (valid) {
      const bf16* op = a.out + (static_cast<size_t>(b) * a.T + t) * a.C + h * a.d;
      const bf16* dp = a.dout + (static_cast<size_t>(b) * a.T + t) * a.C + h * a.d;
      for (int c = 0; c < a.d; c += 8) {
        const uint4 ov = *reinterpret_cast<const uint4*>(op + c);
        const uint4 dv = *reinterpret_cast<const uint4*>(dp + c);
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
        const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          Dt += __low2float(oh[i]) * __low2float(dh[i]) + __high2float(oh[i]) * __high2float(dh[i]);
      }
      lse = a.lse[(static_cast<size_t>(b) * a.heads + h) * a.T + t];
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // P = exp(S*scale - lse)
    const float lse2 = lse * 1.4426950408889634f;
    for (int c = cs0; c < cs1; c += 32) {
      uint32_t r[32];
      float p[32];
      tmem_ld32(trow + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        p[j] = (valid && c + j < a.T) ? exp2f(__uint_as_float(r[j]) * a.scale_log2e - lse2) : 0.f;
      store_p32(smem + a.off_p, row, c, p);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, a.Tp, 0, 0);
      for (int kk = 0; kk < a.d / 16; ++kk)  // dP = dO V^T
        umma_bf16(tmem, desc_kmajor(sDO, 128, a.cw, kk), desc_kmajor(sV, a.Tp, a.cw, kk), idesc_s, kk > 0);
      const uint32_t idesc_t = make_idesc_bf16(128, a.d, 1, 1);
      for (int j = 0; j < ns && do_v; ++j)  // dV_j += P_j^T dO
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(tmem + col_dv + j * a.d, desc_p_mnmajor(sP, kk, j), desc_mnmajor(sDO, 128, a.cw, kk, 0), idesc_t,
                    (qt > 0 || kk > 0) ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // dS = P * (dP - D) * scale, in place over P
    for (int c = cs0; c < cs1; c += 32) {
      uint32_t r[32];
      float p[32];
      tmem_ld32(trow + c, r);
      tmem_ld_wait();
      load_p32(smem + a.off_p, row, c, p);
#pragma unroll
      for (int j = 0; j < 32; ++j) p[j] = p[j] * (__uint_as_float(r[j]) - Dt) * a.scale;
      store_p32(smem + a.off_p, row, c, p);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      const uint32_t idesc_q = make_idesc_bf16(128, a.d, 0, 1);
      for (int kk = 0; kk < a.Tp / 16 && do_v; ++kk)  // dQ = dS K
        umma_bf16(tmem, desc_p_kmajor(sP, kk), desc_mnmajor(sK, a.Tp, a.cw, kk, 0), idesc_q, kk > 0);
      const uint32_t idesc_t = make_idesc_bf16(128, a.d, 1, 1);
      for (int j = 0; j < ns && do_k; ++j)  // dK_j += dS_j^T Q
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(tmem + col_dk + j * a.d, desc_p_mnmajor(sP, kk, j), desc_mnmajor(sQ, 128, a.cw, kk, 0), idesc_t,
                    (qt > 0 || kk > 0) ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    bf16* dq = a.y + (static_cast<size_t>(b) * a.T + (valid ? t : 0)) * (3 * a.C) + cq;
    if (do_v && cd0 < cd1) store_row_from_tmem(trow, cd0, cd1 - cd0, 1.f, dq + cd0, valid);
    tc_fence_before();
    __syncthreads();  // R0 and the Q/dO tiles are free for the next query tile
    tc_fence_after();
  }
  for (int j = 0; j < ns; ++j) {
    const int s = j * 128 + row;
    const bool valid = s < a.T;
    bf16* base = a.y + (static_cast<size_t>(b) * a.T + (valid ? s : 0)) * (3 * a.C) + cq;
    if (cd0 < cd1) {
      if (do_k) store_row_from_tmem(trow, col_dk + j * a.d + cd0, cd1 - cd0, 1.f, base + a.d + cd0, valid);
      if (do_v) store_row_from_tmem(trow, col_dv + j * a.d + cd0, cd1 - cd0, 1.f, base + 2 * a.d + cd0, valid);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, a.tmem_cols);
  }
}

static int attn_common(int B, int T, int heads, int d, AttnArgs* a) {
  if (B <= 0 || T <= 0 || heads <= 0 || d <= 0) return PDDM_ERR_BAD_ARG;
  if (T > 256 || d % 32 != 0 || d > 128) return PDDM_ERR_UNSUPPORTED;
  a->B = B; a->T = T; a->heads = heads; a->d = d; a->C = heads * d;
  a->Tp = (T + 31) / 32 * 32;
  a->cw = d % 64 == 0 ? 64 : 32;
  a->nck = d / a->cw;
  a->nqt = (T + 127) / 128;
  a->scale = 1.0f / sqrtf(static_cast<float>(d));  // (d^-1/4)^2: the reference scales q and k separately
  a->scale_log2e = a->scale * 1.4426950408889634f;
  return PDDM_OK;
}

static int make_tok_map(CUtensorMap* m, const void* base, int channels, int T, int B, int cw, int rows) {
  // [channels (inner), T, B, 1]: the box never crosses a sample, rows >= T are zero-filled
  const uint64_t dims[4] = {static_cast<uint64_t>(channels), static_cast<uint64_t>(T), static_cast<uint64_t>(B), 1};
  const uint64_t str[3] = {static_cast<uint64_t>(channels) * 2, static_cast<uint64_t>(T) * channels * 2,
                           static_cast<uint64_t>(B) * T * channels * 2};
  const uint32_t box[4] = {static_cast<uint32_t>(cw), static_cast<uint32_t>(rows), 1, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box, cw * 2);
}

}  // namespace pddm

using namespace pddm;

extern "C" int pddm_attn_fwd(const pddm_attn_fwd_params* p, pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->qkv || !p->out) return PDDM_ERR_BAD_ARG;
  if (!device_info().ok) return PDDM_ERR_ARCH;
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  int rc = attn_common(p->B, p->T, p->heads, p->d, &a);
  if (rc) return rc;
  if (!aligned16(p->qkv) || !aligned16(p->out)) return PDDM_ERR_BAD_ARG;
  a.y = static_cast<bf16*>(p->out);
  a.lse = p->lse;
  const uint32_t q_bytes = 128 * a.d * 2, kv_bytes = a.Tp * a.d * 2;
  const uint32_t p_bytes = ((a.Tp + 63) / 64) * 128 * 128;
  const uint32_t region_a = (q_bytes + kv_bytes) > p_bytes ? (q_bytes + kv_bytes) : p_bytes;
  a.off_q = 0; a.off_k = q_bytes; a.off_p = 0; a.off_v = (region_a + 1023) / 1024 * 1024;
  a.off_bar = a.off_v + (kv_bytes + 1023) / 1024 * 1024;
  const size_t smem = a.off_bar + 64 + 2 * kAttnParts * 512 + 1024;  // barriers, softmax exchange, alignment slack
  uint32_t cols = 32;
  const uint32_t need = a.Tp > a.d ? a.Tp : a.d;
  while (cols < need) cols <<= 1;
  a.tmem_cols = cols;
  CUtensorMap tmQ, tmKV;
  if ((rc = make_tok_map(&tmQ, p->qkv, 3 * a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmKV, p->qkv, 3 * a.C, a.T, a.B, a.cw, a.Tp))) return rc;
  if (smem > static_cast<size_t>(device_info().max_smem_optin)) return PDDM_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           device_info().max_smem_optin) != cudaSuccess)
    return PDDM_ERR_CUDA;
  PdlLaunch(a.B * a.heads * a.nqt, 128 * kAttnParts, smem, s)(attn_fwd_kernel, tmQ, tmKV, a);
  return launch_status();
}

extern "C" int pddm_attn_bwd(const pddm_attn_bwd_params* p, pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->qkv || !p->out || !p->dout || !p->lse || !p->dqkv) return PDDM_ERR_BAD_ARG;
  if (!device_info().ok) return PDDM_ERR_ARCH;
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  int rc = attn_common(p->B, p->T, p->heads, p->d, &a);
  if (rc) return rc;
  if (!aligned16(p->qkv) || !aligned16(p->out) || !aligned16(p->dout) || !aligned16(p->dqkv)) return PDDM_ERR_BAD_ARG;
  a.out = static_cast<const bf16*>(p->out);
  a.dout = static_cast<const bf16*>(p->dout);
  a.y = static_cast<bf16*>(p->dqkv);
  a.lse = const_cast<float*>(p->lse);
  const int ns = (a.Tp + 127) / 128;
  uint32_t need = (a.Tp > a.d ? a.Tp : a.d) + 2 * ns * a.d;
  const bool two_pass = need > 512;  // e.g. d = 96, T = 256: dV and dK accumulators do not fit together
  if (two_pass) need = (a.Tp > a.d ? a.Tp : a.d) + ns * a.d;
  if (need > 512) return PDDM_ERR_UNSUPPORTED;
  uint32_t cols = 32;
  while (cols < need) cols <<= 1;
  a.tmem_cols = cols;
  const uint32_t q_bytes = 128 * a.d * 2, kv_bytes = a.Tp * a.d * 2;
  uint32_t p_bytes = ((a.Tp + 63) / 64) * 128 * 128;
  if (p_bytes < 2 * 128 * 128) p_bytes = 2 * 128 * 128;  // the transposed view always spans two 64-column chunks
  auto up = [](uint32_t v) { return (v + 1023) / 1024 * 1024; };
  a.off_q = 0;
  a.off_do = up(q_bytes);
  a.off_k = a.off_do + up(q_bytes);
  a.off_v = a.off_k + up(kv_bytes);
  a.off_p = a.off_v + up(kv_bytes);
  a.off_bar = a.off_p + p_bytes;
  const size_t smem = a.off_bar + 64 + 1024;
  if (smem > static_cast<size_t>(device_info().max_smem_optin)) return PDDM_ERR_UNSUPPORTED;
  CUtensorMap tmQ, tmKV, tmDO;
  if ((rc = make_tok_map(&tmQ, p->qkv, 3 * a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmKV, p->qkv, 3 * a.C, a.T, a.B, a.cw, a.Tp))) return rc;
  if ((rc = make_tok_map(&tmDO, p->dout, a.C, a.T, a.B, a.cw, 128))) return rc;
  if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           device_info().max_smem_optin) != cudaSuccess)
    return PDDM_ERR_CUDA;
  if (!two_pass) {
    a.pass = 0;
    PdlLaunch(a.B * a.heads, 128 * kAttnParts, smem, s)(attn_bwd_kernel, tmQ, tmKV, tmDO, a);
  } else {
    a.pass = 1;
    PdlLaunch(a.B * a.heads, 128 * kAttnParts, smem, s)(attn_bwd_kernel, tmQ, tmKV, tmDO, a);
    if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
    a.pass = 2;
    PdlLaunch(a.B * a.heads, 128 * kAttnParts, smem, s)(attn_bwd_kernel, tmQ, tmKV, tmDO, a);
  }
  return launch_status();
}
