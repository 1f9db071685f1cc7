// GroupNorm(32) [+ per-sample scale/shift] [+ SiLU], forward and backward, on NHWC activations.
//
// HBM-bound elementwise + reduction work.  One sample is processed by a thread-block CLUSTER of S CTAs
// (S in {1,2,4,8}); each CTA owns a slab of pixels.  A thread owns 4 consecutive channels (8 B of bf16 / 16 B of
// fp32) and walks down its slab, so every warp touches one contiguous span and the register footprint stays small
// enough for 3-4 resident CTAs per SM.  Partial sums are folded lanes -> channels -> (groups) in shared memory and
// exchanged between the CTAs of the cluster through distributed shared memory in rank order: deterministic, no
// global staging, and the whole op is ONE kernel (statistics pass from HBM, exchange, normalisation pass re-reading
// the slab from L2).  The backward stores dzn = dL/d(gamma*xhat+beta) in the dx buffer during its first pass so the
// second pass needs neither dy nor the activation derivative again.
#include <cuda_bf16.h>
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"

namespace pddm {

typedef __nv_bfloat16 bf16;
constexpr int CH = 4;  // channels per thread

__device__ __forceinline__ void load4(const void* base, int dtype, size_t elem_off, float* f) {
  if (dtype == PDDM_BF16) {
    const uint2 v = *reinterpret_cast<const uint2*>(static_cast<const bf16*>(base) + elem_off);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  } else {
    const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  }
}
__device__ __forceinline__ void store4(void* base, int dtype, size_t elem_off, const float* f) {
  if (dtype == PDDM_BF16) {
    uint2 v;
    *reinterpret_cast<__nv_bfloat162*>(&v.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&v.y) = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(static_cast<bf16*>(base) + elem_off) = v;
  } else {
    *reinterpret_cast<float4*>(static_cast<float*>(base) + elem_off) = make_float4(f[0], f[1], f[2], f[3]);
  }
}
__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

struct GnGeom {
  int B, HW, C, G, cpg, CV, nlanes, S, rows_per_cta;
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_peer_smem(const float* local_ptr, int rank) {
  uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr)), raddr;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
}

// sum of one shared-memory float over the S CTAs of the cluster, in rank order; the loads are issued back to back
// (a dependent add after each ld.shared::cluster would serialise ~S remote round trips)
__device__ __forceinline__ float cluster_fold(const float* local_ptr, int S) {
  float t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < S) t[k] = ld_peer_smem(local_ptr, k);
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < S) v += t[k];
  return v;
}

// ------------------------------------------------------------------------------------------------ forward
template <bool SILU, bool SS>
__global__ void __launch_bounds__(256, 4) gn_fwd_cluster_kernel(pddm_gn_fwd_params p, GnGeom g) {
  pdl_entry();
  extern __shared__ float sh[];  // scratch[2][nlanes][C] | part[G][2] | stat[G][2]
  const int b = blockIdx.y, s = blockIdx.x;  // s = rank in the cluster
  const int cv = threadIdx.x % g.CV, lr = threadIdx.x / g.CV;
  float* sh_sum = sh;
  float* sh_sq = sh + g.nlanes * g.C;
  float* part = sh + 2 * g.nlanes * g.C;
  float* stat = part + 2 * g.G;
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  const size_t base = static_cast<size_t>(b) * g.HW * g.C + cv * CH;
  float sum[CH], sq[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) sum[j] = sq[j] = 0.f;
  for (int r = r0 + lr; r < r1; r += 8 * g.nlanes) {  // 8 independent vector loads in flight per thread
    float f[8][CH];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u * g.nlanes < r1) load4(p.x, p.x_dtype, base + static_cast<size_t>(r + u * g.nlanes) * g.C, f[u]);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u * g.nlanes < r1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          sum[j] += f[u][j];
          sq[j] += f[u][j] * f[u][j];
        }
      }
  }
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    sh_sum[lr * g.C + cv * CH + j] = sum[j];
    sh_sq[lr * g.C + cv * CH + j] = sq[j];
  }
  __syncthreads();
  for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int c = gi * g.cpg; c < (gi + 1) * g.cpg; ++c)
      for (int l = 0; l < g.nlanes; ++l) {
        a += sh_sum[l * g.C + c];
        q += sh_sq[l * g.C + c];
      }
    part[gi * 2] = a;
    part[gi * 2 + 1] = q;
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();
  for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
    float a = 0.f, q = 0.f;
    if (g.S > 1) {
      for (int k = 0; k < g.S; ++k) {
        a += ld_peer_smem(part + gi * 2, k);
        q += ld_peer_smem(part + gi * 2 + 1, k);
      }
    } else {
      a = part[gi * 2];
      q = part[gi * 2 + 1];
    }
    const float n = static_cast<float>(g.cpg) * g.HW;
    const float mean = a / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    stat[gi] = mean;
    stat[g.G + gi] = rstd;
    if (s == 0) {
      p.mean[b * g.G + gi] = mean;
      p.rstd[b * g.G + gi] = rstd;
    }
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();  // peers are done reading this CTA's partials
  // y = ((x - mean) * rstd * gamma + beta) [* (1 + scale) + shift]  ==  x * a + c  with per-channel a, c
  float ca[CH], cc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int c = cv * CH + j;
    const float mu = stat[c / g.cpg], rs = stat[g.G + c / g.cpg];
    float a = rs * p.gamma[c];
    float o = p.beta[c] - mu * a;
    if (SS) {
      const float sc = 1.f + p.scale[static_cast<size_t>(b) * p.ld_ss + c];
      a *= sc;
      o = o * sc + p.shift[static_cast<size_t>(b) * p.ld_ss + c];
    }
    ca[j] = a;
    cc[j] = o;
  }
  for (int r = r0 + lr; r < r1; r += 8 * g.nlanes) {
    float f[8][CH];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u * g.nlanes < r1) load4(p.x, p.x_dtype, base + static_cast<size_t>(r + u * g.nlanes) * g.C, f[u]);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u * g.nlanes < r1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const float z = f[u][j] * ca[j] + cc[j];
          f[u][j] = SILU ? z * fast_sigmoid(z) : z;
        }
        store4(p.y, PDDM_BF16, base + static_cast<size_t>(r + u * g.nlanes) * g.C, f[u]);
      }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// With xh = (x-mean)*rstd, zn = gamma*xh + beta, z = zn*(1+scale)+shift, y = act(z):
//   dz  = dy * act'(z);  dzn = dz*(1+scale);  dshift = sum_hw dz;  dscale = sum_hw dz*zn
//   A_c = sum_hw dzn;  Bq_c = sum_hw dzn*xh;  Xh_c = sum_hw xh
//   S1_g = sum_{c in g} gamma_c A_c;  S2_g = sum_{c in g} gamma_c Bq_c;  n = cpg*HW
//   dx = rstd*(gamma_c*dzn - (S1_g + xh*S2_g)/n);  dgamma_c = sum_b Bq_c;  dbeta_c = sum_b A_c
//   sum_hw dx = rstd*(gamma_c*A_c - (HW*S1_g + S2_g*Xh_c)/n)
// Rank 0 of each cluster writes the per-sample channel totals tot[b][5][C] = {A, Bq, Xh, dshift, dscale}.
// Pass 1 parks dzn in the dx buffer (rounded to dx's dtype); pass 2 turns it into dx in place.
template <bool SILU, bool SS>
__global__ void __launch_bounds__(256, 3) gn_bwd_cluster_kernel(pddm_gn_bwd_params p, float* __restrict__ tot, GnGeom g) {
  pdl_entry();
  extern __shared__ float sh[];  // scratch[NQ][nlanes][C] | part[5][C] | sA[C] sB[C] | S1[G] S2[G]
  constexpr int NQ = SS ? 5 : 3;
  const int b = blockIdx.y, s = blockIdx.x;
  const int cv = threadIdx.x % g.CV, lr = threadIdx.x / g.CV;
  const int LC = g.nlanes * g.C;
  float* part = sh + NQ * LC;
  float* sA = part + 5 * g.C;
  float* sB = sA + g.C;
  float* sS1 = sB + g.C;
  float* sS2 = sS1 + g.G;
  const size_t base = static_cast<size_t>(b) * g.HW * g.C + cv * CH;
  // xh = x*xa + xc ;  z = xh*za + zc (za = gamma*(1+scale), zc = beta*(1+scale)+shift)
  float xa[CH], xc[CH], za[CH], zc[CH], sc1[CH], gam[CH], bet[CH];
  float acc[NQ][CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int c = cv * CH + j;
    const float mu = p.mean[b * g.G + c / g.cpg], rs = p.rstd[b * g.G + c / g.cpg];
    xa[j] = rs;
    xc[j] = -mu * rs;
    gam[j] = p.gamma[c];
    bet[j] = p.beta[c];
    sc1[j] = SS ? 1.f + p.scale[static_cast<size_t>(b) * p.ld_ss + c] : 1.f;
    za[j] = gam[j] * sc1[j];
    zc[j] = bet[j] * sc1[j] + (SS ? p.shift[static_cast<size_t>(b) * p.ld_ss + c] : 0.f);
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q][j] = 0.f;
  }
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  for (int r = r0 + lr; r < r1; r += 4 * g.nlanes) {
    float f[4][CH], d[4][CH];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * g.nlanes < r1) {
        const size_t off = base + static_cast<size_t>(r + u * g.nlanes) * g.C;
        load4(p.x, p.x_dtype, off, f[u]);
        load4(p.dy, PDDM_BF16, off, d[u]);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * g.nlanes < r1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const float xh = f[u][j] * xa[j] + xc[j];
          float dz = d[u][j];
          if (SILU) {
            const float z = xh * za[j] + zc[j];
            const float sg = fast_sigmoid(z);
            dz *= sg * (1.f + z * (1.f - sg));
          }
          const float dzn = dz * sc1[j];
          acc[0][j] += dzn;
          acc[1][j] += dzn * xh;
          acc[2][j] += xh;
          if (SS) {
            acc[3][j] += dz;
            acc[4][j] += dz * (xh * gam[j] + bet[j]);
          }
          d[u][j] = dzn;
        }
        store4(p.dx, p.dx_dtype, base + static_cast<size_t>(r + u * g.nlanes) * g.C, d[u]);
      }
  }
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int c = lr * g.C + cv * CH + j;
#pragma unroll
    for (int q = 0; q < NQ; ++q) sh[q * LC + c] = acc[q][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 5 * g.C; i += blockDim.x) {
    const int q = i / g.C, c = i - q * g.C;
    float v = 0.f;
    if (q < NQ)
      for (int l = 0; l < g.nlanes; ++l) v += sh[q * LC + l * g.C + c];
    part[i] = v;
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();
  for (int i = threadIdx.x; i < 5 * g.C; i += blockDim.x) {
    float v = 0.f;
    if (g.S > 1) {
      for (int k = 0; k < g.S; ++k) v += ld_peer_smem(part + i, k);
    } else {
      v = part[i];
    }
    if (i < g.C) sA[i] = v;
    else if (i < 2 * g.C) sB[i - g.C] = v;
    if (s == 0) tot[static_cast<size_t>(b) * 5 * g.C + i] = v;
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();
  for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = gi * g.cpg; c < (gi + 1) * g.cpg; ++c) {
      s1 += p.gamma[c] * sA[c];
      s2 += p.gamma[c] * sB[c];
    }
    sS1[gi] = s1;
    sS2[gi] = s2;
  }
  __syncthreads();
  // dx = rs*gamma*dzn - rs*(s1 + xh*s2)/n  =  dzn*da + x*db + dc
  const float inv_n = 1.f / (static_cast<float>(g.cpg) * g.HW);
  float da[CH], db[CH], dc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int gi = (cv * CH + j) / g.cpg;
    const float s1 = sS1[gi] * inv_n, s2 = sS2[gi] * inv_n, rs = xa[j];
    da[j] = rs * gam[j];
    db[j] = -rs * s2 * xa[j];
    dc[j] = -rs * (s1 + s2 * xc[j]);
  }
  for (int r = r0 + lr; r < r1; r += 4 * g.nlanes) {
    float f[4][CH], d[4][CH];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * g.nlanes < r1) {
        const size_t off = base + static_cast<size_t>(r + u * g.nlanes) * g.C;
        load4(p.x, p.x_dtype, off, f[u]);
        load4(p.dx, p.dx_dtype, off, d[u]);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r + u * g.nlanes < r1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) d[u][j] = d[u][j] * da[j] + f[u][j] * db[j] + dc[j];
        store4(p.dx, p.dx_dtype, base + static_cast<size_t>(r + u * g.nlanes) * g.C, d[u]);
      }
  }
}

// pass 3a (tiny): per-sample by-products from tot[b][5][C]; grid = (C/128, B)
__global__ void gn_bwd_per_sample_kernel(pddm_gn_bwd_params p, const float* __restrict__ tot, GnGeom g) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= g.C) return;
  const int gi = c / g.cpg;
  const float inv_n = 1.f / (static_cast<float>(g.cpg) * g.HW);
  const float* t = tot + static_cast<size_t>(b) * 5 * g.C;
  if (p.dx_colsum) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = gi * g.cpg; k < (gi + 1) * g.cpg; ++k) {
      s1 += p.gamma[k] * t[k];
      s2 += p.gamma[k] * t[g.C + k];
    }
    p.dx_colsum[static_cast<size_t>(b) * g.C + c] =
        p.rstd[b * g.G + gi] * (p.gamma[c] * t[c] - (g.HW * s1 + s2 * t[2 * g.C + c]) * inv_n);
  }
  if (p.dshift) p.dshift[static_cast<size_t>(b) * p.ld_ss + c] = t[3 * g.C + c];
  if (p.dscale) p.dscale[static_cast<size_t>(b) * p.ld_ss + c] = t[4 * g.C + c];
}
// pass 3b (tiny): dgamma / dbeta = fixed-order sums over the batch; block = 32 channels x 32 batch lanes
__global__ void gn_bwd_finalize_kernel(pddm_gn_bwd_params p, const float* __restrict__ tot, GnGeom g) {
  pdl_entry();
  __shared__ float sg[32][33], sb[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, bl = threadIdx.y;
  float dg = 0.f, db = 0.f;
  if (c < g.C) {
#pragma unroll 4
    for (int b = bl; b < g.B; b += 32) {
      const float* t = tot + static_cast<size_t>(b) * 5 * g.C;
      db += t[c];
      dg += t[g.C + c];
    }
  }
  sg[bl][threadIdx.x] = dg;
  sb[bl][threadIdx.x] = db;
  __syncthreads();
  if (bl == 0 && c < g.C) {
#pragma unroll
    for (int k = 1; k < 32; ++k) {
      dg += sg[k][threadIdx.x];
      db += sb[k][threadIdx.x];
    }
    p.dgamma[c] = dg;
    p.dbeta[c] = db;
  }
}

// ------------------------------------------------------------------------------------------------ slab-resident path
// The common case (bf16 activations, no scale-shift, one sample's slab of rows small enough for shared memory):
// each CTA of the cluster pulls its slab [rows, C] into shared memory ONCE with 1-D bulk copies (cp.async.bulk +
// mbarrier complete_tx, chunked so the statistics pass starts on the first chunk while the rest is in flight), and
// both passes then run out of shared memory.  HBM traffic is the algorithmic minimum -- forward: read x, write y
// (4 B/element); backward: read x and dy, write dx (6 B/element) -- instead of 6 and 12 B/element for the
// register-streaming kernels above, and the deep bulk-copy queue keeps enough bytes in flight per SM that the
// kernel no longer depends on occupancy to cover HBM latency.  A thread owns 8 consecutive channels (one 16-byte
// shared-memory word per row) and the same rows in both passes, so the slab needs no intra-CTA synchronisation.
constexpr int CH8 = 8;
constexpr int kSlabChunks = 8;

struct SlabGeom {
  int B, HW, C, G, cpg, CV, nlanes, S, rows_per_cta, chunk_rows;
  int slab_bytes;  // bytes reserved per slab (rows_per_cta * C * 2, rounded up to 128)
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __low2float(h);
    f[2 * j + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]);
  v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]);
  v.w = pack_bf16(f[6], f[7]);
  return v;
}
__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// issue the chunked bulk copies of this CTA's slab(s); called by one thread after the barriers are initialised
__device__ __forceinline__ void slab_issue_loads(const SlabGeom& g, int nrows, uint64_t* bars, uint8_t* slab0,
                                                 const bf16* src0, uint8_t* slab1, const bf16* src1) {
  const int row_bytes = g.C * 2;
  for (int ch = 0, r = 0; r < nrows; ++ch, r += g.chunk_rows) {
    const int rows = min(g.chunk_rows, nrows - r);
    const uint32_t bytes = static_cast<uint32_t>(rows) * row_bytes;
    mbar_expect_tx(&bars[ch], src1 ? 2 * bytes : bytes);
    bulk_load_1d(slab0 + static_cast<size_t>(r) * row_bytes, src0 + static_cast<size_t>(r) * g.C, bytes, &bars[ch]);
    if (src1)
      bulk_load_1d(slab1 + static_cast<size_t>(r) * row_bytes, src1 + static_cast<size_t>(r) * g.C, bytes, &bars[ch]);
  }
}

__device__ __forceinline__ void ld8(const float* sm, float* r) {
  const float4 a = *reinterpret_cast<const float4*>(sm), b = *reinterpret_cast<const float4*>(sm + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}

template <bool SILU>
__global__ void __launch_bounds__(256, 2) gn_fwd_slab_kernel(pddm_gn_fwd_params p, SlabGeom g) {
  pdl_entry();
  extern __shared__ __align__(128) uint8_t smraw[];
  uint8_t* slab = smraw;
  float* scratch = reinterpret_cast<float*>(smraw + g.slab_bytes);  // [2][nlanes][C]
  const int LC = g.nlanes * g.C;
  float* csum = scratch + 2 * LC;  // [2][C] per-channel sums of this CTA; later the per-channel coefficients a, c
  float* part = csum + 2 * g.C;    // [G][2] per-group sums of this CTA (read by the cluster peers)
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * g.G);
  const int b = blockIdx.y, s = blockIdx.x;
  const int cv = threadIdx.x % g.CV, lr = threadIdx.x / g.CV;
  const int r0 = s * g.rows_per_cta, nrows = min(g.rows_per_cta, g.HW - r0);
  const size_t gbase = (static_cast<size_t>(b) * g.HW + r0) * g.C;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      for (int i = 0; i < kSlabChunks; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
      slab_issue_loads(g, nrows, bars, slab, static_cast<const bf16*>(p.x) + gbase, nullptr, nullptr);
    }
    __syncwarp();
  }
  __syncthreads();
  float sum[CH8], sq[CH8];
#pragma unroll
  for (int j = 0; j < CH8; ++j) sum[j] = sq[j] = 0.f;
  for (int ch = 0, c0 = 0; c0 < nrows; ++ch, c0 += g.chunk_rows) {
    mbar_wait(&bars[ch], 0);
    const int c1 = min(c0 + g.chunk_rows, nrows);
    for (int r = c0 + lr; r < c1; r += g.nlanes) {
      float f[CH8];
      unpack8(*reinterpret_cast<const uint4*>(slab + (static_cast<size_t>(r) * g.C + cv * CH8) * 2), f);
#pragma unroll
      for (int j = 0; j < CH8; ++j) {
        sum[j] += f[j];
        sq[j] += f[j] * f[j];
      }
    }
  }
  {
    float* d = scratch + lr * g.C + cv * CH8;
    *reinterpret_cast<float4*>(d) = make_float4(sum[0], sum[1], sum[2], sum[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(sum[4], sum[5], sum[6], sum[7]);
    *reinterpret_cast<float4*>(d + LC) = make_float4(sq[0], sq[1], sq[2], sq[3]);
    *reinterpret_cast<float4*>(d + LC + 4) = make_float4(sq[4], sq[5], sq[6], sq[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.C; i += blockDim.x) {  // fold the row lanes, fixed order
    const int q = i / g.C, c = i - q * g.C;
    float v = 0.f;
    for (int l = 0; l < g.nlanes; ++l) v += scratch[q * LC + l * g.C + c];
    csum[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * g.G; i += blockDim.x) {  // fold the channels of a group
    const int gi = i >> 1, q = i & 1;
    float v = 0.f;
    for (int c = gi * g.cpg; c < (gi + 1) * g.cpg; ++c) v += csum[q * g.C + c];
    part[i] = v;
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();
  // 2G threads fold the group partials over the cluster (rank order) into mean / rstd ...
  float* gstat = scratch;  // [G][2] (the lane scratch is dead by now)
  for (int i = threadIdx.x; i < g.G; i += blockDim.x) {
    const float a = g.S > 1 ? cluster_fold(part + i * 2, g.S) : part[i * 2];
    const float q = g.S > 1 ? cluster_fold(part + i * 2 + 1, g.S) : part[i * 2 + 1];
    const float n = static_cast<float>(g.cpg) * g.HW;
    const float mean = a / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    gstat[i * 2] = mean;
    gstat[i * 2 + 1] = rstd;
    if (s == 0) {
      p.mean[b * g.G + i] = mean;
      p.rstd[b * g.G + i] = rstd;
    }
  }
  __syncthreads();
  // ... and every channel thread derives y = x*a + c
  for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
    const int gi = c / g.cpg;
    const float ca = gstat[gi * 2 + 1] * p.gamma[c];
    csum[c] = ca;
    csum[g.C + c] = p.beta[c] - gstat[gi * 2] * ca;
  }
  __syncthreads();
  float ca[CH8], cc[CH8];
  ld8(csum + cv * CH8, ca);
  ld8(csum + g.C + cv * CH8, cc);
  bf16* y = static_cast<bf16*>(p.y) + gbase;
  for (int r = lr; r < nrows; r += g.nlanes) {
    const size_t off = static_cast<size_t>(r) * g.C + cv * CH8;
    float f[CH8];
    unpack8(*reinterpret_cast<const uint4*>(slab + off * 2), f);
#pragma unroll
    for (int j = 0; j < CH8; ++j) {
      const float z = f[j] * ca[j] + cc[j];
      f[j] = SILU ? z * fast_sigmoid(z) : z;
    }
    *reinterpret_cast<uint4*>(y + off) = pack8(f);
  }
  if (g.S > 1) cluster_sync_all();  // no CTA may exit while a peer can still read its partials
}

// backward: slabs of x and dy; pass 1 overwrites the dy slab with dzn (rounded to bf16, as the streaming kernel
// parks it in dx), pass 2 writes dx.  tot[b][5][C] as above (rows 3,4 are zero without scale-shift).
template <bool SILU>
__global__ void __launch_bounds__(256, 2)
gn_bwd_slab_kernel(pddm_gn_bwd_params p, float* __restrict__ tot, SlabGeom g) {
  pdl_entry();
  extern __shared__ __align__(128) uint8_t smraw[];
  uint8_t* slab_x = smraw;
  uint8_t* slab_d = smraw + g.slab_bytes;
  float* scratch = reinterpret_cast<float*>(smraw + 2 * g.slab_bytes);  // [3][nlanes][C]
  const int LC = g.nlanes * g.C;
  float* part = scratch + 3 * LC;  // [3][C] per-channel sums of this CTA (read by the cluster peers)
  float* coef = part + 3 * g.C;    // [4][C]: xa, xc, gamma, beta ; later da, db, dc
  float* sAB = coef + 4 * g.C;     // [2][C]: gamma*A, gamma*Bq over the whole sample
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAB + 2 * g.C);
  const int b = blockIdx.y, s = blockIdx.x;
  const int cv = threadIdx.x % g.CV, lr = threadIdx.x / g.CV;
  const int r0 = s * g.rows_per_cta, nrows = min(g.rows_per_cta, g.HW - r0);
  const size_t gbase = (static_cast<size_t>(b) * g.HW + r0) * g.C;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      for (int i = 0; i < kSlabChunks; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
      slab_issue_loads(g, nrows, bars, slab_x, static_cast<const bf16*>(p.x) + gbase, slab_d,
                       static_cast<const bf16*>(p.dy) + gbase);
    }
    __syncwarp();
  }
  for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
    const int gi = c / g.cpg;
    const float mu = p.mean[b * g.G + gi], rs = p.rstd[b * g.G + gi];
    coef[c] = rs;
    coef[g.C + c] = -mu * rs;
    coef[2 * g.C + c] = p.gamma[c];
    coef[3 * g.C + c] = p.beta[c];
  }
  __syncthreads();
  float xa[CH8], xc[CH8], gam[CH8], bet[CH8];
  ld8(coef + cv * CH8, xa);
  ld8(coef + g.C + cv * CH8, xc);
  ld8(coef + 2 * g.C + cv * CH8, gam);
  ld8(coef + 3 * g.C + cv * CH8, bet);
  float acc[3][CH8];
#pragma unroll
  for (int j = 0; j < CH8; ++j) acc[0][j] = acc[1][j] = acc[2][j] = 0.f;
  for (int ch = 0, c0 = 0; c0 < nrows; ++ch, c0 += g.chunk_rows) {
    mbar_wait(&bars[ch], 0);
    const int c1 = min(c0 + g.chunk_rows, nrows);
    for (int r = c0 + lr; r < c1; r += g.nlanes) {
      const size_t off = (static_cast<size_t>(r) * g.C + cv * CH8) * 2;
      float f[CH8], d[CH8];
      unpack8(*reinterpret_cast<const uint4*>(slab_x + off), f);
      unpack8(*reinterpret_cast<const uint4*>(slab_d + off), d);
#pragma unroll
      for (int j = 0; j < CH8; ++j) {
        const float xh = f[j] * xa[j] + xc[j];
        float dz = d[j];
        if (SILU) {
          const float z = xh * gam[j] + bet[j];
          const float sg = fast_sigmoid(z);
          dz *= sg * (1.f + z * (1.f - sg));
        }
        acc[0][j] += dz;
        acc[1][j] += dz * xh;
        acc[2][j] += xh;
        d[j] = dz;
      }
      *reinterpret_cast<uint4*>(slab_d + off) = pack8(d);
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float* d = scratch + q * LC + lr * g.C + cv * CH8;
    *reinterpret_cast<float4*>(d) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * g.C; i += blockDim.x) {
    const int q = i / g.C, c = i - q * g.C;
    float v = 0.f;
    for (int l = 0; l < g.nlanes; ++l) v += scratch[q * LC + l * g.C + c];
    part[i] = v;
  }
  if (g.S > 1) cluster_sync_all(); else __syncthreads();
  for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
    float v[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) v[q] = g.S > 1 ? cluster_fold(part + q * g.C + c, g.S) : part[q * g.C + c];
    const float gm = coef[2 * g.C + c];
    sAB[c] = gm * v[0];
    sAB[g.C + c] = gm * v[1];
    if (s == 0) {
      float* t = tot + static_cast<size_t>(b) * 5 * g.C + c;
      t[0] = v[0];
      t[g.C] = v[1];
      t[2 * g.C] = v[2];
      t[3 * g.C] = 0.f;
      t[4 * g.C] = 0.f;
    }
  }
  __syncthreads();
  // dx = rs*gamma*dzn - rs*(s1 + xh*s2)/n  =  dzn*da + x*db + dc     (coef rows 0..2 are rewritten in place: each
  // channel is read and written by the same thread, and the per-thread copies xa/xc/gam/bet are already in registers)
  const float inv_n = 1.f / (static_cast<float>(g.cpg) * g.HW);
  for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
    const int gi = c / g.cpg;
    float s1 = 0.f, s2 = 0.f;
    for (int k = gi * g.cpg; k < (gi + 1) * g.cpg; ++k) {
      s1 += sAB[k];
      s2 += sAB[g.C + k];
    }
    s1 *= inv_n;
    s2 *= inv_n;
    const float rs = coef[c], xcc = coef[g.C + c], gm = coef[2 * g.C + c];
    if (s == 0 && p.dx_colsum) {  // sum_hw dx = rstd*(gamma*A - (HW*S1 + S2*Xh)/n), from the sample totals
      const float* t = tot + static_cast<size_t>(b) * 5 * g.C + c;
      p.dx_colsum[static_cast<size_t>(b) * g.C + c] = rs * (sAB[c] - (g.HW * s1 + s2 * t[2 * g.C]));
    }
    coef[c] = rs * gm;
    coef[g.C + c] = -rs * s2 * rs;
    coef[2 * g.C + c] = -rs * (s1 + s2 * xcc);
  }
  __syncthreads();
  float da[CH8], db[CH8], dc[CH8];
  ld8(coef + cv * CH8, da);
  ld8(coef + g.C + cv * CH8, db);
  ld8(coef + 2 * g.C + cv * CH8, dc);
  bf16* dx = static_cast<bf16*>(p.dx) + gbase;
  for (int r = lr; r < nrows; r += g.nlanes) {
    const size_t off = static_cast<size_t>(r) * g.C + cv * CH8;
    float f[CH8], d[CH8];
    unpack8(*reinterpret_cast<const uint4*>(slab_x + off * 2), f);
    unpack8(*reinterpret_cast<const uint4*>(slab_d + off * 2), d);
#pragma unroll
    for (int j = 0; j < CH8; ++j) d[j] = d[j] * da[j] + f[j] * db[j] + dc[j];
    *reinterpret_cast<uint4*>(dx + off) = pack8(d);
  }
  if (g.S > 1) cluster_sync_all();  // no CTA may exit while a peer can still read its partials
}

// geometry of the slab path; returns false when the shape does not qualify (caller falls back to streaming)
static bool make_slab_geom(int B, int HW, int C, int G, int nslabs, int nscratch, SlabGeom* g, int* threads,
                           size_t* smem) {
  if (C % CH8 || C % G || C / CH8 > 256) return false;
  g->B = B; g->HW = HW; g->C = C; g->G = G; g->cpg = C / G; g->CV = C / CH8;
  g->nlanes = 256 / g->CV;
  if (g->nlanes > HW) g->nlanes = HW;
  *threads = g->CV * g->nlanes;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const size_t budget = static_cast<size_t>(device_info().max_smem_optin) - 1024;
  // Cluster size: measured over the model's shapes (profiles/r1_bench_gn_*.log, PDDM_GN_S sweeps), the kernel is
  // fastest when a CTA holds about 64 KB of slabs (all of its slabs together): smaller CTAs pay their fixed cost
  // (barriers, folds, cluster exchange) too often, larger ones leave one CTA per SM.  Among the feasible sizes pick
  // the one closest to that; PDDM_GN_S forces an upper bound for experiments.
  const int smax = env_knobs().gn_s > 0 ? env_knobs().gn_s : 8;
  int best_S = 0;
  long long best_dist = 0;
  for (int S = 8; S >= 1; S /= 2) {
    const int rows = (HW + S - 1) / S;
    // a CTA wants >= 4 rows per lane to amortise its fixed cost
    if (S > 1 && (S > smax || rows < 4 * g->nlanes || (S - 1) * rows >= HW || B * (S / 2) >= 8 * sms)) continue;
    const int slab = (rows * C * 2 + 127) / 128 * 128;
    const size_t need = static_cast<size_t>(nslabs) * slab +
                        (static_cast<size_t>(nscratch) * g->nlanes * C + 9 * C + 4 * G) * sizeof(float) +
                        kSlabChunks * sizeof(uint64_t);
    if (need > budget) continue;
    long long dist = static_cast<long long>(nslabs) * slab - 64 * 1024;
    if (dist < 0) dist = -dist;
    if (best_S == 0 || dist < best_dist) {
      best_S = S;
      best_dist = dist;
    }
  }
  if (best_S == 0) return false;
  {
    const int S = best_S;
    const int rows = (HW + S - 1) / S;
    const int slab = (rows * C * 2 + 127) / 128 * 128;
    g->S = S;
    g->rows_per_cta = rows;
    g->slab_bytes = slab;
    int cr = (rows + kSlabChunks - 1) / kSlabChunks;
    cr = (cr + g->nlanes - 1) / g->nlanes * g->nlanes;
    g->chunk_rows = cr;
    *smem = static_cast<size_t>(nslabs) * slab +
            (static_cast<size_t>(nscratch) * g->nlanes * C + 9 * C + 4 * G) * sizeof(float) +
            kSlabChunks * sizeof(uint64_t);
    return true;
  }
  return false;
}

static int make_geom(int B, int HW, int C, int G, GnGeom* g, int* threads) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8 || C % G || C / CH > 256) return PDDM_ERR_UNSUPPORTED;  // C <= 1024
  g->B = B; g->HW = HW; g->C = C; g->G = G; g->cpg = C / G; g->CV = C / CH;
  g->nlanes = 256 / g->CV > 0 ? 256 / g->CV : 1;
  if (g->nlanes > HW) g->nlanes = HW;
  *threads = g->CV * g->nlanes;
  // cluster of S CTAs per sample (portable cluster sizes only): enough CTAs for ~4 per SM, >= 4 rows per lane
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int S = 1;
  while (S < 8 && B * S < 4 * sms && HW / (2 * S) >= g->nlanes * 4) S *= 2;
  g->rows_per_cta = (HW + S - 1) / S;
  g->S = S;
  if ((S - 1) * g->rows_per_cta >= HW) {  // keep every rank non-empty
    g->S = 1;
    g->rows_per_cta = HW;
  }
  return PDDM_OK;
}

// groupnorm_pipe.cu
int gn_fwd_pipe(const pddm_gn_fwd_params* p, cudaStream_t s);
int gn_bwd_pipe(const pddm_gn_bwd_params* p, float* part_dgamma, float* part_dbeta, int ld_part, cudaStream_t s);

static int launch_cluster(const void* func, dim3 grid, int threads, size_t smem, int S, cudaStream_t s, void** args) {
  PdlLaunch l(grid, dim3(threads), smem, s, S);
  return l.launch_c(func, args) == cudaSuccess ? PDDM_OK : PDDM_ERR_CUDA;
}

}  // namespace pddm

using namespace pddm;

// Scratch queries.  The forward exchanges its partial sums through DSMEM and needs none (the argument is kept
// so callers need not special-case it); the backward needs B*5*C floats for the per-sample channel totals.
extern "C" size_t pddm_gn_silu_fwd_workspace(int32_t B, int32_t G) {
  (void)B;
  (void)G;
  return 16;
}
extern "C" size_t pddm_gn_silu_bwd_workspace(int32_t B, int32_t C) {
  return static_cast<size_t>(B) * 5 * C * sizeof(float);
}

extern "C" int pddm_gn_silu_fwd(const pddm_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                                pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  (void)workspace;
  (void)workspace_bytes;
  if (!p || !p->x || !p->y || !p->gamma || !p->beta || !p->mean || !p->rstd) return PDDM_ERR_BAD_ARG;
  if ((p->scale == nullptr) != (p->shift == nullptr)) return PDDM_ERR_BAD_ARG;
  GnGeom g;
  int threads;
  int rc = make_geom(p->B, p->HW, p->C, p->G, &g, &threads);
  if (rc) return rc;
  if (!aligned16(p->x) || !aligned16(p->y)) return PDDM_ERR_BAD_ARG;
  if (!env_knobs().gn_nopipe && !env_knobs().gn_stream) {
    rc = gn_fwd_pipe(p, s);  // persistent bulk-tensor kernel (groupnorm_pipe.cu); 1 = shape does not qualify
    if (rc != 1) return rc;
  }
  // the cluster kernels below take one contiguous tensor only
  if (p->x2 || (p->ldx && p->ldx != p->C) || (p->ldy && p->ldy != p->C)) return PDDM_ERR_UNSUPPORTED;
  if (!p->scale && p->x_dtype == PDDM_BF16 && !env_knobs().gn_stream) {
    SlabGeom sg;
    int st;
    size_t ssm;
    if (make_slab_geom(p->B, p->HW, p->C, p->G, 1, 2, &sg, &st, &ssm)) {
      const void* fn = p->silu ? reinterpret_cast<const void*>(gn_fwd_slab_kernel<true>)
                               : reinterpret_cast<const void*>(gn_fwd_slab_kernel<false>);
      if (ensure_smem_optin(fn)) return PDDM_ERR_CUDA;
      pddm_gn_fwd_params pp = *p;
      void* args[2] = {&pp, &sg};
      rc = launch_cluster(fn, dim3(sg.S, sg.B), st, ssm, sg.S, s, args);
      return rc ? rc : launch_status();
    }
  }
  const size_t smem = (2 * static_cast<size_t>(g.nlanes) * g.C + 4 * g.G) * sizeof(float);
  pddm_gn_fwd_params pp = *p;
  void* args[2] = {&pp, &g};
  const bool ss = p->scale != nullptr;
  const void* fn = p->silu ? (ss ? reinterpret_cast<const void*>(gn_fwd_cluster_kernel<true, true>)
                                 : reinterpret_cast<const void*>(gn_fwd_cluster_kernel<true, false>))
                           : (ss ? reinterpret_cast<const void*>(gn_fwd_cluster_kernel<false, true>)
                                 : reinterpret_cast<const void*>(gn_fwd_cluster_kernel<false, false>));
  rc = launch_cluster(fn, dim3(g.S, g.B), threads, smem, g.S, s, args);
  return rc ? rc : launch_status();
}

extern "C" int pddm_gn_silu_bwd(const pddm_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                                pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->x || !p->dy || !p->dx || !p->gamma || !p->beta || !p->mean || !p->rstd) return PDDM_ERR_BAD_ARG;
  if ((!p->dgamma || !p->dbeta) && (!p->part_dgamma || !p->part_dbeta)) return PDDM_ERR_BAD_ARG;
  if ((p->dgamma || p->dbeta) && !workspace) return PDDM_ERR_BAD_ARG;
  if ((p->scale == nullptr) != (p->shift == nullptr)) return PDDM_ERR_BAD_ARG;
  GnGeom g;
  int threads;
  int rc = make_geom(p->B, p->HW, p->C, p->G, &g, &threads);
  if (rc) return rc;
  if (!aligned16(p->x) || !aligned16(p->dy) || !aligned16(p->dx)) return PDDM_ERR_BAD_ARG;
  const size_t need = static_cast<size_t>(g.B) * 5 * g.C * sizeof(float);
  const bool want_final = p->dgamma && p->dbeta;  // else the caller reduces the per-sample partials itself
  if (want_final && workspace_bytes < need) return PDDM_ERR_WORKSPACE;
  float* tot = static_cast<float*>(workspace);
  if (!env_knobs().gn_nopipe && !env_knobs().gn_stream) {
    // persistent bulk-tensor kernel (groupnorm_pipe.cu).  Per-sample partials go to the caller's matrices, or to
    // tot[b][0][c] (dbeta) / tot[b][1][c] (dgamma), which the batch fold below reads.
    float* pg = p->part_dgamma ? p->part_dgamma : tot + g.C;
    float* pb = p->part_dbeta ? p->part_dbeta : tot;
    const int ldp = p->part_dgamma ? (p->ld_part > 0 ? p->ld_part : g.C) : 5 * g.C;
    rc = gn_bwd_pipe(p, pg, pb, ldp, s);
    if (rc == PDDM_OK) {
      if (want_final) {
        if (p->part_dgamma) return PDDM_ERR_BAD_ARG;  // either final sums or partials, not both
        PdlLaunch((g.C + 31) / 32, dim3(32, 32), 0, s)(gn_bwd_finalize_kernel, *p, tot, g);
      }
      return launch_status();
    }
    if (rc != 1) return rc;
  }
  // the cluster kernels below take one contiguous tensor and none of the extensions
  if (p->x2 || p->gres || p->dx2 || p->part_dgamma || p->dx_colsum2 || p->dx_accumulate || p->colsum_accumulate ||
      (p->ldx && p->ldx != p->C) || (p->lddy && p->lddy != p->C) || (p->ld_dx && p->ld_dx != p->C) ||
      (p->ld_colsum && p->ld_colsum != p->C))
    return PDDM_ERR_UNSUPPORTED;
  const bool ss = p->scale != nullptr;
  bool launched = false;
  if (!ss && p->x_dtype == PDDM_BF16 && p->dx_dtype == PDDM_BF16 && !env_knobs().gn_stream) {
    SlabGeom sg;
    int st;
    size_t ssm;
    if (make_slab_geom(p->B, p->HW, p->C, p->G, 2, 3, &sg, &st, &ssm)) {
      const void* fn = p->silu ? reinterpret_cast<const void*>(gn_bwd_slab_kernel<true>)
                               : reinterpret_cast<const void*>(gn_bwd_slab_kernel<false>);
      if (ensure_smem_optin(fn)) return PDDM_ERR_CUDA;
      pddm_gn_bwd_params pp = *p;
      void* args[3] = {&pp, &tot, &sg};
      rc = launch_cluster(fn, dim3(sg.S, sg.B), st, ssm, sg.S, s, args);
      if (rc) return rc;
      if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
      launched = true;
    }
  }
  const int nq = ss ? 5 : 3;
  const size_t smem = (nq * static_cast<size_t>(g.nlanes) * g.C + 7 * g.C + 2 * g.G) * sizeof(float);
  const void* fn = p->silu ? (ss ? reinterpret_cast<const void*>(gn_bwd_cluster_kernel<true, true>)
                                 : reinterpret_cast<const void*>(gn_bwd_cluster_kernel<true, false>))
                           : (ss ? reinterpret_cast<const void*>(gn_bwd_cluster_kernel<false, true>)
                                 : reinterpret_cast<const void*>(gn_bwd_cluster_kernel<false, false>));
  if (!launched && smem > 48 * 1024) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return PDDM_ERR_UNSUPPORTED;
  }
  if (!launched) {
    pddm_gn_bwd_params pp = *p;
    void* args[3] = {&pp, &tot, &g};
    rc = launch_cluster(fn, dim3(g.S, g.B), threads, smem, g.S, s, args);
    if (rc) return rc;
    if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  }
  if (!launched && (p->dx_colsum || p->dshift || p->dscale)) {
    PdlLaunch(dim3((g.C + 127) / 128, g.B), 128, 0, s)(gn_bwd_per_sample_kernel, *p, tot, g);
    if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  }
  PdlLaunch((g.C + 31) / 32, dim3(32, 32), 0, s)(gn_bwd_finalize_kernel, *p, tot, g);
  return launch_status();
}
