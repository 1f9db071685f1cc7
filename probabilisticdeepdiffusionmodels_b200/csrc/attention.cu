// QKVAttention (src/modules/unet.py:237-256) as fused tcgen05 kernels.
//
// Sequence length T = H*W <= 256, so one head's whole K and V sit in shared memory and a 128-query tile's score
// matrix sits in tensor memory: a single pass, no online-softmax rescaling.
//   forward  (CTA = sample x head x 128-query tile):  S = Q K^T (TMEM) -> fp32 softmax in registers -> P (bf16,
//            swizzled smem) -> O = P V (TMEM) -> normalise, store; log-sum-exp saved.
//   backward (CTA = sample x head x 128-key tile; two key tiles = a cluster of two CTAs): per query tile S -> P,
//            dP = dO V^T, dV += P^T dO, dS = P*(dP - D), dQ = dS K, dK += dS^T Q; every product is a tcgen05.mma
//            with accumulators in TMEM; the two key tiles' dQ products meet through an fp32 workspace.
// Each [tokens x d] operand tile is loaded once by TMA as 64- (or 32-) channel swizzled chunks and is used both
// as a K-major operand (contraction over channels) and, re-described, as an MN-major operand (contraction over
// tokens); the P / dS tile likewise serves as A (K-major) and A^T (MN-major).
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

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
  int Tk, ns;     // backward: padded keys per key tile (min(128, Tp)), key tiles (= cluster size, 1 or 2)
  float* ws;      // backward, ns == 2: fp32 [B*heads, 2, d/4, 128, 4] partial dQ products
  long long* dbg; // PDDM_ATTN_DBG=1: per-CTA clock64 stamps at the phase boundaries (debug builds of the timeline)
};
#define ATTN_STAMP(k) do { if (a.dbg && tid == 0) a.dbg[blockIdx.x * 24 + (k)] = clock64(); } while (0)

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

// P tile as K-major A (M = rows, K = columns): k-step kk covers columns [16kk, 16kk+16)
__device__ __forceinline__ uint64_t desc_p_kmajor(uint32_t ptile, int kk) {
  return make_smem_desc(ptile + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024, kLayoutSW128);
}
// P tile as MN-major A (M = columns [128*mt, +128), K = rows): k-step kk covers rows [16kk, 16kk+16)
__device__ __forceinline__ uint64_t desc_p_mnmajor(uint32_t ptile, int kk, int mt) {
  return make_smem_desc(ptile + (2 * mt) * (128 * 128) + kk * 2048, 128 * 128, 1024, kLayoutSW128);
}

// ---- MMA issue loops with the descriptors hoisted: one 64-bit add per k-step (the address field counts 16-byte units)
// K-major x K-major over the d channels of two [rows x cw]-chunked tiles (S = Q K^T, dP = dO V^T)
__device__ __forceinline__ void issue_kk(uint32_t tmem_d, uint32_t tileA, int rowsA, uint32_t tileB, int rowsB, int cw,
                                         int nck, uint32_t idesc) {
  const uint64_t a0 = desc_kmajor(tileA, rowsA, cw, 0), b0 = desc_kmajor(tileB, rowsB, cw, 0);
  const uint32_t stepA = (rowsA * cw * 2) >> 4, stepB = (rowsB * cw * 2) >> 4;
  const int spc = cw >> 4;
  uint32_t acc = 0;
  for (int c = 0; c < nck; ++c) {
    uint64_t da = a0 + static_cast<uint64_t>(c) * stepA, db = b0 + static_cast<uint64_t>(c) * stepB;
    for (int q = 0; q < spc; ++q) {
      umma_bf16(tmem_d, da, db, idesc, acc);
      acc = 1;
      da += 2;
      db += 2;
    }
  }
}
// P (K-major A over `nk` 16-column steps) x MN-major [rows x cw]-chunked tile (O = P V, dQ = dS K)
__device__ __forceinline__ void issue_p_k(uint32_t tmem_d, uint32_t ptile, uint32_t tileB, int rowsB, int cw, int nk,
                                          uint32_t idesc) {
  const uint64_t a0 = desc_p_kmajor(ptile, 0);
  uint64_t db = desc_mnmajor(tileB, rowsB, cw, 0, 0);
  const uint32_t stepB = (16 * cw * 2) >> 4;
  for (int kk = 0; kk < nk; ++kk) {
    umma_bf16(tmem_d, a0 + static_cast<uint64_t>((kk >> 2) * 1024 + (kk & 3) * 2), db, idesc, kk > 0);
    db += stepB;
  }
}
// P^T (MN-major A, 128 rows = 8 k-steps) x MN-major [128 x cw]-chunked tile (dV += P^T dO, dK += dS^T Q)
__device__ __forceinline__ void issue_pt(uint32_t tmem_d, uint32_t ptile, uint32_t tileB, int cw, uint32_t idesc,
                                         uint32_t acc) {
  uint64_t da = desc_p_mnmajor(ptile, 0, 0), db = desc_mnmajor(tileB, 128, cw, 0, 0);
  const uint32_t stepB = (16 * cw * 2) >> 4;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    umma_bf16(tmem_d, da, db, idesc, acc);
    acc = 1;
    da += 2048 >> 4;
    db += stepB;
  }
}

// ---- output staging: accumulators leave through a shared-memory tile laid out exactly like a TMA-loaded operand
// ([rows x cw]-channel chunks, 128 B / 64 B swizzle) and one bulk tensor store per chunk; rows beyond T are clipped
// by the tensor map.  (Per-thread 16-byte row stores cost one 32-byte sector each: 6k cycles for the dK/dV drain.)
__device__ __forceinline__ void stage_row32(uint8_t* tile, int rows, int cw, int row, int c0, const uint32_t* pk) {
  if (cw == 64) {
    uint8_t* base = tile + (c0 >> 6) * (rows * 128) + row * 128;
    const int u0 = (c0 & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(base + (((u0 + u) ^ (row & 7)) << 4)) =
          make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  } else {  // 32-channel chunks, 64-byte rows: the 16-byte unit index is XORed with address bits [7, 9)
    uint8_t* base = tile + (c0 >> 5) * (rows * 64) + row * 64;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(base + ((u ^ ((row >> 1) & 3)) << 4)) =
          make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  }
}
// drain `ncol` columns of this thread's TMEM lane from column `col`, scaled, as bf16 to tile columns [c0, c0 + ncol)
__device__ __forceinline__ void stage_from_tmem(uint32_t taddr_row, int col, int c0, int ncol, float mul, uint8_t* tile,
                                                int rows, int cw, int row) {
  for (int c = 0; c < ncol; c += 32) {
    uint32_t r[32], pk[16];
    tmem_ld32(taddr_row + col + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(r[2 * j]) * mul, __uint_as_float(r[2 * j + 1]) * mul);
    if (row < rows) stage_row32(tile, rows, cw, row, c0 + c, pk);
  }
}

// ---- small helpers of the elementwise phases ----------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU.EX2, no range fix-up; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 packed bf16 pairs (32 values of row r, columns [c0, c0+32)) -> the [128 x cols] tile of 64-column SW128 chunks
__device__ __forceinline__ void store_p32_packed(uint8_t* ptile, int r, int c0, const uint32_t* pk) {
  stage_row32(ptile, 128, 64, r, c0, pk);
}
// P = exp2(S*sl2e - sub) for 32 columns -> 16 packed bf16 pairs; MASK: columns >= nvalid are zero
template <bool MASK>
__device__ __forceinline__ float exp_chunk(const uint32_t (&s)[32], float sl2e, float sub, int nvalid, uint32_t* pk) {
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(s[j]), sl2e, -sub));
    float p1 = ex2_approx(fmaf(__uint_as_float(s[j + 1]), sl2e, -sub));
    if (MASK) {
      p0 = j < nvalid ? p0 : 0.f;
      p1 = j + 1 < nvalid ? p1 : 0.f;
    }
    sum += p0 + p1;
    pk[j >> 1] = pack_bf16(p0, p1);
  }
  return sum;
}

// =================================================================================================== forward
// CTA = (sample, head, 128-query tile), 128*kAttnParts threads: kAttnParts threads per query row split the columns of
// the softmax and of the output drain; the row maximum and the row sum are combined through a small shared-memory
// exchange.  TMA and MMA instructions are issued from warp-uniform code by one elected lane of warp 0 (a divergent
// `tid == 0` branch makes every descriptor operand a uniform-register waterfall loop, ~150 cycles per MMA issued).
// The elementwise loops carry no per-element bounds checks: padded key columns (T % 32 != 0) are masked only in the
// one chunk that straddles T, padded query rows produce finite garbage that the output tensor map clips.
__global__ void __launch_bounds__(128 * kAttnParts, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int T = a.T, Tp = a.Tp, d = a.d, cw = a.cw, nck = a.nck;
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
  uint8_t* ptile = smem + a.off_p;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_load, (128 + 2 * Tp) * d * 2);
      const int cq = h * 3 * d;
      for (int c = 0; c < nck; ++c) {
        tma_load_4d(smem + a.off_q + c * (128 * cw * 2), &tmQ, bar_load, cq + c * cw, t0, b, 0);
        tma_load_4d(smem + a.off_k + c * (Tp * cw * 2), &tmKV, bar_load, cq + d + c * cw, 0, b, 0);
        tma_load_4d(smem + a.off_v + c * (Tp * cw * 2), &tmKV, bar_load, cq + 2 * d + c * cw, 0, b, 0);
      }
    }
    __syncwarp();
    mbar_wait(bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      issue_kk(tmem, sQ, 128, sK, Tp, cw, nck, make_idesc_bf16(128, Tp, 0, 0));
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- softmax, kAttnParts threads per query row (part = column range) ----
  const int row = tid & 127, part = tid >> 7;
  const uint32_t trow = tmem + (static_cast<uint32_t>(row) << 16);  // lane = row (warp w owns lanes 32(w%4)..)
  float* xch = reinterpret_cast<float*>(smem + a.off_bar + 64);     // [max | sum][part][128 rows]
  const int nch_s = Tp >> 5, nch_d = d >> 5;
  const int cs0 = nch_s * part / kAttnParts * 32, cs1 = nch_s * (part + 1) / kAttnParts * 32;
  const int cd0 = nch_d * part / kAttnParts * 32, cd1 = nch_d * (part + 1) / kAttnParts * 32;
  const int t = t0 + row;
  const float sl2e = a.scale_log2e;
  float mx = -INFINITY;
  for (int c = cs0; c < cs1; c += 32) {
    uint32_t r[32];
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
    if (c + 32 <= T) {
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c + j < T) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
  }
  xch[part * 128 + row] = mx;
  __syncthreads();
  mx = xch[row];
#pragma unroll
  for (int k = 1; k < kAttnParts; ++k) mx = fmaxf(mx, xch[k * 128 + row]);
  float sum = 0.f;
  const float mxs = mx * sl2e;
  for (int c = cs0; c < cs1; c += 32) {
    uint32_t r[32], pk[16];
    tmem_ld32(trow + c, r);
    tmem_ld_wait();
    if (c + 32 <= T)
      sum += exp_chunk<false>(r, sl2e, mxs, 32, pk);
    else
      sum += exp_chunk<true>(r, sl2e, mxs, T - c, pk);
    store_p32_packed(ptile, row, c, pk);  // aliases Q|K, which the S MMA has finished reading
  }
  xch[(kAttnParts + part) * 128 + row] = sum;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  sum = xch[kAttnParts * 128 + row];
#pragma unroll
  for (int k = 1; k < kAttnParts; ++k) sum += xch[(kAttnParts + k) * 128 + row];
  if (warp == 0) {
    if (elect_one()) {
      issue_p_k(tmem, sP, sV, Tp, cw, Tp / 16, make_idesc_bf16(128, d, 0, 1));
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  // the P tile is free: stage the normalised output there and store it with one bulk tensor copy per chunk
  if (cd0 < cd1) stage_from_tmem(trow, cd0, cd0, cd1 - cd0, 1.f / sum, ptile, 128, cw, row);
  if (t < T && a.lse && part == 0) a.lse[(static_cast<size_t>(b) * a.heads + h) * T + t] = mx * a.scale + logf(sum);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      for (int c = 0; c < nck; ++c) tma_store_4d(&tmO, ptile + c * (128 * cw * 2), h * d + c * cw, t0, b, 0);
      bulk_commit_group();
      bulk_wait_group_read<0>();
    }
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, a.tmem_cols);
  }
}

// =================================================================================================== backward
// CTA = (sample, head, key tile of <= 128 keys); a head with two key tiles (128 < T <= 256) is a CLUSTER of two CTAs.
// The CTA keeps K_j, V_j and the dK_j / dV_j accumulators (TMEM) and walks over the query tiles:
//   S = Q_i K_j^T -> P = exp2(S*scale*log2e - lse) (bf16, swizzled smem + packed registers)
//   dP = dO_i V_j^T, dV_j += P^T dO_i -> dS = P * (dP - D) in place over P -> dQ_i^(j) = dS K_j, dK_j += dS^T Q_i
// so that 96 KB of shared memory and 256 TMEM columns suffice at d = 64: two CTAs per SM, whose load / MMA /
// elementwise phases overlap.  dQ_i is the sum of the two key tiles' products: CTA r first works on the query tile its
// PEER owns and parks the fp32 partial product in the workspace, then on its own tile, whose product it completes
// with the peer's partial (cluster barrier, release / acquire) -- two addends, so the sum does not depend on timing.
// 128*kAttnParts threads: kAttnParts threads per row split the columns of every elementwise phase; a warp touches
// only the TMEM lanes of its quadrant (warp % 4 == row / 32 for both halves).  The softmax scale of dS is applied
// when dQ and dK are drained.  TMA / MMA issue and output staging as in the forward kernel.
__global__ void __launch_bounds__(128 * kAttnParts, 2)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDKV,
                const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* bar_q = bar_kv + 1;
  uint64_t* bar_mma = bar_kv + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_kv + 3);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, part = tid >> 7;
  const int T = a.T, d = a.d, cw = a.cw, nck = a.nck, Tk = a.Tk, ns = a.ns;
  ATTN_STAMP(0);
  const int r = blockIdx.x % ns;  // key tile == rank in the cluster
  const int bh = blockIdx.x / ns;
  const int h = bh % a.heads, b = bh / a.heads;
  const int k0 = r * 128, kvalid = T - k0;  // keys [k0, k0 + Tk) of which the first `kvalid` exist
  const int col_dv = Tk > d ? Tk : d, col_dk = col_dv + d;

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
  ATTN_STAMP(1);
  const uint32_t sQ = smem_u32(smem + a.off_q), sDO = smem_u32(smem + a.off_do), sK = smem_u32(smem + a.off_k),
                 sV = smem_u32(smem + a.off_v), sP = smem_u32(smem + a.off_p);
  uint8_t* ptile = smem + a.off_p;
  const uint32_t trow = tmem + (static_cast<uint32_t>(row) << 16);
  // column ranges of this thread: 32-column chunks of the [128 x Tk] score tile and of the d-wide accumulators
  const int nch_s = Tk >> 5, nch_d = d >> 5;
  const int cs0 = nch_s * part / kAttnParts * 32, cs1 = nch_s * (part + 1) / kAttnParts * 32;  // <= 2 chunks
  const int cd0 = nch_d * part / kAttnParts * 32, cd1 = nch_d * (part + 1) / kAttnParts * 32;
  const int cq = h * 3 * d;
  const float sl2e = a.scale_log2e, scale = a.scale;
  uint32_t mma_phase = 0;
  // fp32 partial dQ products cross the cluster through the workspace as [writer rank][d/4][128 rows] float4
  float4* ws_mine = reinterpret_cast<float4*>(a.ws) + (static_cast<size_t>(bh) * 2 + r) * (d / 4) * 128;
  const float4* ws_peer = reinterpret_cast<const float4*>(a.ws) + (static_cast<size_t>(bh) * 2 + (1 - r)) * (d / 4) * 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_kv, 2 * Tk * d * 2);
      for (int c = 0; c < nck; ++c) {
        tma_load_4d(smem + a.off_k + c * (Tk * cw * 2), &tmKV, bar_kv, cq + d + c * cw, k0, b, 0);
        tma_load_4d(smem + a.off_v + c * (Tk * cw * 2), &tmKV, bar_kv, cq + 2 * d + c * cw, k0, b, 0);
      }
      const int t0 = (ns == 2 ? 1 - r : 0) * 128;
      mbar_expect_tx(bar_q, 3 * 128 * d * 2);
      for (int c = 0; c < nck; ++c) {
        tma_load_4d(smem + a.off_q + c * (128 * cw * 2), &tmQ, bar_q, cq + c * cw, t0, b, 0);
        tma_load_4d(smem + a.off_do + c * (128 * cw * 2), &tmDO, bar_q, h * d + c * cw, t0, b, 0);
        tma_load_4d(ptile + c * (128 * cw * 2), &tmO, bar_q, h * d + c * cw, t0, b, 0);  // O_i, for D (see below)
      }
    }
    __syncwarp();
  }
  // lse of this thread's row in both query tiles (padded rows: +inf -> P = 0, dS = 0)
  float lq[2] = {INFINITY, INFINITY};
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (i < ns && i * 128 + row < T)
      lq[i] = a.lse[(static_cast<size_t>(b) * a.heads + h) * T + i * 128 + row] * 1.4426950408889634f;
  float* xch = reinterpret_cast<float*>(smem + a.off_bar + 64);  // [part][128 rows] partial D
  ATTN_STAMP(2);

  for (int it = 0; it < ns; ++it) {
    const int i = ns == 2 ? (it == 0 ? 1 - r : r) : 0;  // the peer's query tile first, then our own
    const float lse2 = i == 0 ? lq[0] : lq[1];
    const bool last = it + 1 == ns;
    mbar_wait(bar_q, it & 1);
    if (warp == 0) {
      if (it == 0) mbar_wait(bar_kv, 0);
      tc_fence_after();
      ATTN_STAMP(3 + it * 10);
      if (elect_one()) {
        issue_kk(tmem, sQ, 128, sK, Tk, cw, nck, make_idesc_bf16(128, Tk, 0, 0));  // S = Q K^T
        umma_commit(bar_mma);
      }
      __syncwarp();
      ATTN_STAMP(4 + it * 10);
    }
    // ---- D[t] = sum_c dO[t,c] * O[t,c] from the two TMA-loaded tiles (O_i sits in the P region until P is written):
    // each of the kAttnParts threads of a row takes its share of the 16-byte units, the shares meet in shared memory.
    float Dt = 0.f;
    {
      const uint8_t* otile = ptile;
      const uint8_t* dtile = smem + a.off_do;
      const int units = d >> 3;  // 16-byte units per row
      for (int u = part; u < units; u += kAttnParts) {
        const int ch = u * 8, c = ch / cw, uu = (ch - c * cw) >> 3;
        const int off = c * (128 * cw * 2) + row * (cw * 2) + ((cw == 64 ? (uu ^ (row & 7)) : (uu ^ ((row >> 1) & 3))) << 4);
        const uint4 ov = *reinterpret_cast<const uint4*>(otile + off);
        const uint4 dv = *reinterpret_cast<const uint4*>(dtile + off);
        const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
        const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 of = __bfloat1622float2(oh[q]), df = __bfloat1622float2(dh[q]);
          Dt = fmaf(of.x, df.x, Dt);
          Dt = fmaf(of.y, df.y, Dt);
        }
      }
      xch[part * 128 + row] = Dt;
      __syncthreads();  // also: every thread has read O_i before the first P store below
      Dt = xch[row];
#pragma unroll
      for (int k = 1; k < kAttnParts; ++k) Dt += xch[k * 128 + row];
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    ATTN_STAMP(5 + it * 10);
    // ---- P = exp2(S*scale*log2e - lse*log2e): bf16 pairs kept in registers for the dS phase and stored for the MMAs
    uint32_t pk[32];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = cs0 + 32 * q;
      if (c < cs1) {
        uint32_t sreg[32];
        tmem_ld32(trow + c, sreg);
        tmem_ld_wait();
        if (c + 32 <= kvalid)
          exp_chunk<false>(sreg, sl2e, lse2, 32, pk + q * 16);
        else
          exp_chunk<true>(sreg, sl2e, lse2, kvalid - c, pk + q * 16);
        store_p32_packed(ptile, row, c, pk + q * 16);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    ATTN_STAMP(6 + it * 10);
    if (warp == 0) {
      if (elect_one()) {
        issue_kk(tmem, sDO, 128, sV, Tk, cw, nck, make_idesc_bf16(128, Tk, 0, 0));                // dP = dO V^T
        issue_pt(tmem + col_dv, sP, sDO, cw, make_idesc_bf16(128, d, 1, 1), it > 0 ? 1u : 0u);  // dV_j += P^T dO
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
    ATTN_STAMP(7 + it * 10);
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    ATTN_STAMP(8 + it * 10);
    // ---- dS = P * (dP - D), in place over P (the scale rides with the dQ / dK drains)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = cs0 + 32 * q;
      if (c < cs1) {
        uint32_t g[32];
        tmem_ld32(trow + c, g);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const uint32_t u = pk[q * 16 + (j >> 1)];
          const float p0 = __uint_as_float(u << 16), p1 = __uint_as_float(u & 0xffff0000u);
          pk[q * 16 + (j >> 1)] = pack_bf16(p0 * (__uint_as_float(g[j]) - Dt), p1 * (__uint_as_float(g[j + 1]) - Dt));
        }
        store_p32_packed(ptile, row, c, pk + q * 16);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    ATTN_STAMP(9 + it * 10);
    if (warp == 0) {
      if (elect_one()) {
        issue_p_k(tmem, sP, sK, Tk, cw, Tk / 16, make_idesc_bf16(128, d, 0, 1));               // dQ_i^(j) = dS K_j
        issue_pt(tmem + col_dk, sP, sQ, cw, make_idesc_bf16(128, d, 1, 1), it > 0 ? 1u : 0u);  // dK_j += dS^T Q_i
        umma_commit(bar_mma);
      }
      __syncwarp();
    }
    ATTN_STAMP(10 + it * 10);
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    ATTN_STAMP(11 + it * 10);
    if (!last) {
      // the Q / dO tiles are free: fetch our own query tile while the partial product drains to the workspace
      if (warp == 0) {
        if (elect_one()) {
          mbar_expect_tx(bar_q, 3 * 128 * d * 2);
          for (int c = 0; c < nck; ++c) {
            tma_load_4d(smem + a.off_q + c * (128 * cw * 2), &tmQ, bar_q, cq + c * cw, r * 128, b, 0);
            tma_load_4d(smem + a.off_do + c * (128 * cw * 2), &tmDO, bar_q, h * d + c * cw, r * 128, b, 0);
            tma_load_4d(ptile + c * (128 * cw * 2), &tmO, bar_q, h * d + c * cw, r * 128, b, 0);
          }
        }
        __syncwarp();
      }
      for (int c = cd0; c < cd1; c += 32) {
        uint32_t g[32];
        tmem_ld32(trow + c, g);
        tmem_ld_wait();
        float4* dst = ws_mine + (c >> 2) * 128 + row;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          __stcg(dst + u * 128, make_float4(__uint_as_float(g[4 * u]), __uint_as_float(g[4 * u + 1]),
                                            __uint_as_float(g[4 * u + 2]), __uint_as_float(g[4 * u + 3])));
      }
    } else {
      if (ns == 2) {  // the partials were written an iteration ago: the release fence finds them already performed
        cluster_arrive_release();
        cluster_wait_acquire();
      }
      for (int c = cd0; c < cd1; c += 32) {
        uint32_t g[32], o[16];
        tmem_ld32(trow + c, g);
        tmem_ld_wait();
        if (ns == 2) {
          const float4* src = ws_peer + (c >> 2) * 128 + row;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 v = __ldcg(src + u * 128);
            g[4 * u] = __float_as_uint(__uint_as_float(g[4 * u]) + v.x);
            g[4 * u + 1] = __float_as_uint(__uint_as_float(g[4 * u + 1]) + v.y);
            g[4 * u + 2] = __float_as_uint(__uint_as_float(g[4 * u + 2]) + v.z);
            g[4 * u + 3] = __float_as_uint(__uint_as_float(g[4 * u + 3]) + v.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
          o[j] = pack_bf16(__uint_as_float(g[2 * j]) * scale, __uint_as_float(g[2 * j + 1]) * scale);
        stage_row32(smem + a.off_q, 128, cw, row, c, o);  // the Q tile is free: dQ leaves through it
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();  // TMEM columns [0, Tk) are free for the next query tile; the staged dQ tile is complete
    tc_fence_after();
    ATTN_STAMP(12 + it * 10);
  }
  if (warp == 0) {
    if (elect_one()) {
      for (int c = 0; c < nck; ++c) tma_store_4d(&tmDQ, smem + a.off_q + c * (128 * cw * 2), cq + c * cw, r * 128, b, 0);
      bulk_commit_group();
    }
    __syncwarp();
  }
  // dK_j, dV_j leave through the K and V tiles (every MMA that read them has completed)
  if (cd0 < cd1) {
    stage_from_tmem(trow, col_dk + cd0, cd0, cd1 - cd0, scale, smem + a.off_k, Tk, cw, row);
    stage_from_tmem(trow, col_dv + cd0, cd0, cd1 - cd0, 1.f, smem + a.off_v, Tk, cw, row);
  }
  fence_proxy_async();
  ATTN_STAMP(23);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      for (int c = 0; c < nck; ++c) {
        tma_store_4d(&tmDKV, smem + a.off_k + c * (Tk * cw * 2), cq + d + c * cw, k0, b, 0);
        tma_store_4d(&tmDKV, smem + a.off_v + c * (Tk * cw * 2), cq + 2 * d + c * cw, k0, b, 0);
      }
      bulk_commit_group();
      bulk_wait_group_read<0>();
    }
    __syncwarp();
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
  CUtensorMap tmQ, tmKV, tmO;
  if ((rc = make_tok_map(&tmQ, p->qkv, 3 * a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmKV, p->qkv, 3 * a.C, a.T, a.B, a.cw, a.Tp))) return rc;
  if ((rc = make_tok_map(&tmO, p->out, a.C, a.T, a.B, a.cw, 128))) return rc;
  if (smem > static_cast<size_t>(device_info().max_smem_optin)) return PDDM_ERR_UNSUPPORTED;
  if ((rc = ensure_smem_optin(reinterpret_cast<const void*>(attn_fwd_kernel)))) return rc;
  PdlLaunch(a.B * a.heads * a.nqt, 128 * kAttnParts, smem, s)(attn_fwd_kernel, tmQ, tmKV, tmO, a);
  return launch_status();
}

extern "C" int64_t pddm_attn_bwd_workspace_bytes(int32_t B, int32_t T, int32_t heads, int32_t d) {
  if (B <= 0 || T <= 128 || heads <= 0 || d <= 0) return 0;
  return static_cast<int64_t>(B) * heads * 2 * 128 * d * 4;
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
  a.ns = (a.Tp + 127) / 128;
  a.Tk = a.Tp < 128 ? a.Tp : 128;
  if (a.ns == 2) {
    if (!p->ws || !aligned16(p->ws) || p->ws_bytes < pddm_attn_bwd_workspace_bytes(p->B, p->T, p->heads, p->d))
      return PDDM_ERR_BAD_ARG;
    a.ws = static_cast<float*>(p->ws);
  }
  const uint32_t need = (a.Tk > a.d ? a.Tk : a.d) + 2 * a.d;  // S | dP | dQ share columns; dV_j; dK_j
  uint32_t cols = 32;
  while (cols < need) cols <<= 1;
  a.tmem_cols = cols;
  const uint32_t q_bytes = 128 * a.d * 2, kv_bytes = a.Tk * a.d * 2;
  const uint32_t p_bytes = 2 * 128 * 128;  // [128 x 128] bf16: the transposed view always spans two 64-column chunks
  auto up = [](uint32_t v) { return (v + 1023) / 1024 * 1024; };
  a.off_q = 0;
  a.off_do = up(q_bytes);
  a.off_k = a.off_do + up(q_bytes);
  a.off_v = a.off_k + up(kv_bytes);
  a.off_p = a.off_v + up(kv_bytes);
  a.off_bar = a.off_p + p_bytes;
  const size_t smem = a.off_bar + 64 + kAttnParts * 512 + 1024;  // barriers, D exchange, alignment slack
  if (smem > static_cast<size_t>(device_info().max_smem_optin)) return PDDM_ERR_UNSUPPORTED;
  CUtensorMap tmQ, tmKV, tmDO, tmO, tmDQ, tmDKV;
  if ((rc = make_tok_map(&tmO, p->out, a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmQ, p->qkv, 3 * a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmKV, p->qkv, 3 * a.C, a.T, a.B, a.cw, a.Tk))) return rc;
  if ((rc = make_tok_map(&tmDO, p->dout, a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmDQ, p->dqkv, 3 * a.C, a.T, a.B, a.cw, 128))) return rc;
  if ((rc = make_tok_map(&tmDKV, p->dqkv, 3 * a.C, a.T, a.B, a.cw, a.Tk))) return rc;
  if ((rc = ensure_smem_optin(reinterpret_cast<const void*>(attn_bwd_kernel)))) return rc;
  if (env_knobs().attn_dbg > 0) {  // debug only: synchronous, prints the phase timeline of a few CTAs
    const int nb = a.B * a.heads * a.ns;
    cudaMalloc(&a.dbg, static_cast<size_t>(nb) * 24 * 8);
    cudaMemset(a.dbg, 0, static_cast<size_t>(nb) * 24 * 8);
    PdlLaunch(nb, 128 * kAttnParts, smem, s, a.ns == 2 ? 2 : 0)(attn_bwd_kernel, tmQ, tmKV, tmDO, tmO, tmDQ, tmDKV, a);
    cudaStreamSynchronize(s);
    long long* hbuf = static_cast<long long*>(malloc(static_cast<size_t>(nb) * 24 * 8));
    cudaMemcpy(hbuf, a.dbg, static_cast<size_t>(nb) * 24 * 8, cudaMemcpyDeviceToHost);
    const int picks[4] = {0, 1, nb / 2, nb - 1};
    for (int q = 0; q < 4; ++q) {
      const long long* t = hbuf + static_cast<size_t>(picks[q]) * 24;
      fprintf(stderr, "attn_bwd dbg T=%d d=%d cta %d:", a.T, a.d, picks[q]);
      for (int k = 1; k < 24; ++k)
        if (t[k]) fprintf(stderr, " [%d]%lld", k, t[k] - t[0]);
      fprintf(stderr, "\n");
    }
    free(hbuf);
    cudaFree(a.dbg);
    return launch_status();
  }
  PdlLaunch(a.B * a.heads * a.ns, 128 * kAttnParts, smem, s, a.ns == 2 ? 2 : 0)(attn_bwd_kernel, tmQ, tmKV, tmDO, tmO, tmDQ, tmDKV, a);
  return launch_status();
}
