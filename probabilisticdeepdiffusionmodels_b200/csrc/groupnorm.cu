// GroupNorm(32) [+ per-sample scale/shift] [+ SiLU], forward and backward, on NHWC activations.
//
// HBM-bound.  Thread mapping: a thread owns one 8-channel vector (16 B of bf16 / 32 B of fp32) and walks down
// the pixels of its sample slab, so a warp always touches one fully contiguous span of memory.  Statistics are
// accumulated per channel in registers, folded to groups through shared memory, and exchanged between the
// CTAs of one sample through a small [B, S, ...] partials array that the second pass re-reads (the tensor
// itself is re-read from L2: at B=128 every activation of the CIFAR config fits in the 126 MB L2).
#include <cuda_bf16.h>

#include "host_common.h"

namespace pddm {

typedef __nv_bfloat16 bf16;
constexpr int kMaxSplit = 16;

__device__ __forceinline__ void load8(const void* base, int dtype, size_t elem_off, float* f) {
  if (dtype == PDDM_BF16) {
    const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const bf16*>(base) + elem_off);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __low2float(h[i]);
      f[2 * i + 1] = __high2float(h[i]);
    }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}
__device__ __forceinline__ void store8(void* base, int dtype, size_t elem_off, const float* f) {
  if (dtype == PDDM_BF16) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(static_cast<bf16*>(base) + elem_off) = v;
  } else {
    float* p = static_cast<float*>(base) + elem_off;
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

struct GnGeom {
  int B, HW, C, G, cpg, C8, nlanes, S, rows_per_cta;
};

// ------------------------------------------------------------------------------------------------ forward
// pass 1: per (sample, split) partial group sums -> part[b][s][g][2]
__global__ void gn_stats_kernel(const void* __restrict__ x, int x_dtype, float* __restrict__ part, GnGeom g) {
  extern __shared__ float sh[];  // [2][nlanes][C]
  const int b = blockIdx.y, s = blockIdx.x;
  const int cv = threadIdx.x % g.C8, lr = threadIdx.x / g.C8;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  for (int r = r0 + lr; r < r1; r += g.nlanes) {
    float f[8];
    load8(x, x_dtype, (static_cast<size_t>(b) * g.HW + r) * g.C + cv * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sum[j] += f[j];
      sq[j] += f[j] * f[j];
    }
  }
  // fixed-order (deterministic) fold: lanes -> channels -> groups
  float* sh_sum = sh;                       // [nlanes][C]
  float* sh_sq = sh + g.nlanes * g.C;       // [nlanes][C]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh_sum[lr * g.C + cv * 8 + j] = sum[j];
    sh_sq[lr * g.C + cv * 8 + j] = sq[j];
  }
  __syncthreads();
  for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int c = gi * g.cpg; c < (gi + 1) * g.cpg; ++c)
      for (int l = 0; l < g.nlanes; ++l) {
        a += sh_sum[l * g.C + c];
        q += sh_sq[l * g.C + c];
      }
    float* o = part + ((static_cast<size_t>(b) * g.S + s) * g.G + gi) * 2;
    o[0] = a;
    o[1] = q;
  }
}

// pass 2: normalise + affine (+ scale/shift) (+ SiLU)
__global__ void gn_apply_kernel(pddm_gn_fwd_params p, const float* __restrict__ part, GnGeom g) {
  extern __shared__ float sh[];  // mean[G], rstd[G]
  const int b = blockIdx.y, s = blockIdx.x;
  for (int gi = threadIdx.x; gi < g.G; gi += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int k = 0; k < g.S; ++k) {
      const float* o = part + ((static_cast<size_t>(b) * g.S + k) * g.G + gi) * 2;
      a += o[0];
      q += o[1];
    }
    const float n = static_cast<float>(g.cpg) * g.HW;
    const float mean = a / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    sh[gi] = mean;
    sh[g.G + gi] = rstd;
    if (s == 0) {
      p.mean[b * g.G + gi] = mean;
      p.rstd[b * g.G + gi] = rstd;
    }
  }
  __syncthreads();
  const int cv = threadIdx.x % g.C8, lr = threadIdx.x / g.C8;
  float ga[8], be[8], mu[8], rs[8], sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    ga[j] = p.gamma[c];
    be[j] = p.beta[c];
    mu[j] = sh[c / g.cpg];
    rs[j] = sh[g.G + c / g.cpg];
    sc[j] = p.scale ? 1.f + p.scale[static_cast<size_t>(b) * p.ld_ss + c] : 1.f;
    sf[j] = p.shift ? p.shift[static_cast<size_t>(b) * p.ld_ss + c] : 0.f;
  }
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  for (int r = r0 + lr; r < r1; r += g.nlanes) {
    const size_t off = (static_cast<size_t>(b) * g.HW + r) * g.C + cv * 8;
    float f[8];
    load8(p.x, p.x_dtype, off, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z = (f[j] - mu[j]) * rs[j] * ga[j] + be[j];
      z = z * sc[j] + sf[j];
      f[j] = p.silu ? z / (1.f + __expf(-z)) : z;
    }
    store8(p.y, PDDM_BF16, off, f);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// With xh = (x-mean)*rstd, zn = gamma*xh + beta, z = zn*(1+scale)+shift, y = act(z):
//   dz  = dy * act'(z);  dzn = dz*(1+scale);  dshift = sum_hw dz;  dscale = sum_hw dz*zn
//   A_c = sum_hw dzn;  Bq_c = sum_hw dzn*xh;  Xh_c = sum_hw xh
//   S1_g = sum_{c in g} gamma_c A_c;  S2_g = sum_{c in g} gamma_c Bq_c;  n = cpg*HW
//   dx = rstd*(gamma_c*dzn - (S1_g + xh*S2_g)/n);  dgamma_c = sum_b Bq_c;  dbeta_c = sum_b A_c
//   sum_hw dx = rstd*(gamma_c*A_c - (HW*S1_g + S2_g*Xh_c)/n)
// pass 1 writes part[b][s][5][C] = {A, Bq, Xh, dshift, dscale}
__global__ void gn_bwd_stats_kernel(pddm_gn_bwd_params p, float* __restrict__ part, GnGeom g) {
  extern __shared__ float sh[];  // [5][nlanes][C]
  const int b = blockIdx.y, s = blockIdx.x;
  const int cv = threadIdx.x % g.C8, lr = threadIdx.x / g.C8;
  float ga[8], be[8], mu[8], rs[8], sc[8], sf[8];
  float A[8], Bq[8], Xh[8], Ds[8], Dc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    ga[j] = p.gamma[c];
    be[j] = p.beta[c];
    mu[j] = p.mean[b * g.G + c / g.cpg];
    rs[j] = p.rstd[b * g.G + c / g.cpg];
    sc[j] = p.scale ? 1.f + p.scale[static_cast<size_t>(b) * p.ld_ss + c] : 1.f;
    sf[j] = p.shift ? p.shift[static_cast<size_t>(b) * p.ld_ss + c] : 0.f;
    A[j] = Bq[j] = Xh[j] = Ds[j] = Dc[j] = 0.f;
  }
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  for (int r = r0 + lr; r < r1; r += g.nlanes) {
    const size_t off = (static_cast<size_t>(b) * g.HW + r) * g.C + cv * 8;
    float f[8], d[8];
    load8(p.x, p.x_dtype, off, f);
    load8(p.dy, PDDM_BF16, off, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu[j]) * rs[j];
      const float zn = xh * ga[j] + be[j];
      const float z = zn * sc[j] + sf[j];
      float dz = d[j];
      if (p.silu) {
        const float sg = 1.f / (1.f + __expf(-z));
        dz *= sg * (1.f + z * (1.f - sg));
      }
      const float dzn = dz * sc[j];
      A[j] += dzn;
      Bq[j] += dzn * xh;
      Xh[j] += xh;
      Ds[j] += dz;
      Dc[j] += dz * zn;
    }
  }
  // fixed-order (deterministic) fold over the pixel lanes: sh[q][lane][C]
  const int LC = g.nlanes * g.C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lr * g.C + cv * 8 + j;
    sh[c] = A[j];
    sh[LC + c] = Bq[j];
    sh[2 * LC + c] = Xh[j];
    sh[3 * LC + c] = Ds[j];
    sh[4 * LC + c] = Dc[j];
  }
  __syncthreads();
  float* o = part + (static_cast<size_t>(b) * g.S + s) * 5 * g.C;
  for (int i = threadIdx.x; i < 5 * g.C; i += blockDim.x) {
    const int q = i / g.C, c = i - q * g.C;
    float v = 0.f;
    for (int l = 0; l < g.nlanes; ++l) v += sh[q * LC + l * g.C + c];
    o[i] = v;
  }
}

// pass 2: dx; CTA s == 0 of each sample also emits the per-sample channel totals tot[b][5][C]
__global__ void gn_bwd_apply_kernel(pddm_gn_bwd_params p, const float* __restrict__ part, float* __restrict__ tot,
                                    GnGeom g) {
  extern __shared__ float sh[];  // A[C], Bq[C], S1[G], S2[G]
  const int b = blockIdx.y, s = blockIdx.x;
  float* sA = sh;
  float* sB = sh + g.C;
  float* sS1 = sh + 2 * g.C;
  float* sS2 = sh + 2 * g.C + g.G;
  for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
    float v[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < g.S; ++k) {
      const float* o = part + (static_cast<size_t>(b) * g.S + k) * 5 * g.C;
#pragma unroll
      for (int q = 0; q < 5; ++q) v[q] += o[q * g.C + c];
    }
    sA[c] = v[0];
    sB[c] = v[1];
    if (s == 0) {
#pragma unroll
      for (int q = 0; q < 5; ++q) tot[(static_cast<size_t>(b) * 5 + q) * g.C + c] = v[q];
    }
  }
  __syncthreads();
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
  const int cv = threadIdx.x % g.C8, lr = threadIdx.x / g.C8;
  const float inv_n = 1.f / (static_cast<float>(g.cpg) * g.HW);
  float ga[8], be[8], mu[8], rs[8], sc[8], sf[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j, gi = c / g.cpg;
    ga[j] = p.gamma[c];
    be[j] = p.beta[c];
    mu[j] = p.mean[b * g.G + gi];
    rs[j] = p.rstd[b * g.G + gi];
    sc[j] = p.scale ? 1.f + p.scale[static_cast<size_t>(b) * p.ld_ss + c] : 1.f;
    sf[j] = p.shift ? p.shift[static_cast<size_t>(b) * p.ld_ss + c] : 0.f;
    s1[j] = sS1[gi] * inv_n;
    s2[j] = sS2[gi] * inv_n;
  }
  const int r0 = s * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.HW);
  for (int r = r0 + lr; r < r1; r += g.nlanes) {
    const size_t off = (static_cast<size_t>(b) * g.HW + r) * g.C + cv * 8;
    float f[8], d[8];
    load8(p.x, p.x_dtype, off, f);
    load8(p.dy, PDDM_BF16, off, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (f[j] - mu[j]) * rs[j];
      const float z = (xh * ga[j] + be[j]) * sc[j] + sf[j];
      float dz = d[j];
      if (p.silu) {
        const float sg = 1.f / (1.f + __expf(-z));
        dz *= sg * (1.f + z * (1.f - sg));
      }
      f[j] = rs[j] * (ga[j] * dz * sc[j] - (s1[j] + xh * s2[j]));
    }
    store8(p.dx, p.dx_dtype, off, f);
  }
}

// pass 3 (tiny): batch reductions and per-sample by-products from tot[b][5][C]
__global__ void gn_bwd_finalize_kernel(pddm_gn_bwd_params p, const float* __restrict__ tot, GnGeom g) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.C) return;
  const int gi = c / g.cpg;
  const float inv_n = 1.f / (static_cast<float>(g.cpg) * g.HW);
  const float gam = p.gamma[c];
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < g.B; ++b) {
    const float* t = tot + static_cast<size_t>(b) * 5 * g.C;
    const float A = t[c], Bq = t[g.C + c];
    dg += Bq;
    db += A;
    if (p.dx_colsum) {
      float s1 = 0.f, s2 = 0.f;
      for (int k = gi * g.cpg; k < (gi + 1) * g.cpg; ++k) {
        s1 += p.gamma[k] * t[k];
        s2 += p.gamma[k] * t[g.C + k];
      }
      p.dx_colsum[static_cast<size_t>(b) * g.C + c] =
          p.rstd[b * g.G + gi] * (gam * A - (g.HW * s1 + s2 * t[2 * g.C + c]) * inv_n);
    }
    if (p.dshift) p.dshift[static_cast<size_t>(b) * p.ld_ss + c] = t[3 * g.C + c];
    if (p.dscale) p.dscale[static_cast<size_t>(b) * p.ld_ss + c] = t[4 * g.C + c];
  }
  p.dgamma[c] = dg;
  p.dbeta[c] = db;
}

static int make_geom(int B, int HW, int C, int G, GnGeom* g, int* threads) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0) return PDDM_ERR_BAD_ARG;
  if (C % 8 || C % G || C / 8 > 512) return PDDM_ERR_UNSUPPORTED;
  g->B = B; g->HW = HW; g->C = C; g->G = G; g->cpg = C / G; g->C8 = C / 8;
  g->nlanes = 256 / g->C8 > 0 ? 256 / g->C8 : 1;
  if (g->nlanes > HW) g->nlanes = HW;
  *threads = g->C8 * g->nlanes;
  // split a sample over S CTAs so that ~2 waves of CTAs exist and each CTA still sees >= 4 rows per lane
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  int S = (2 * sms + B - 1) / B;
  const int max_s = HW / (g->nlanes * 4) > 0 ? HW / (g->nlanes * 4) : 1;
  if (S > max_s) S = max_s;
  if (S > kMaxSplit) S = kMaxSplit;
  if (S < 1) S = 1;
  g->rows_per_cta = (HW + S - 1) / S;
  g->S = (HW + g->rows_per_cta - 1) / g->rows_per_cta;
  return PDDM_OK;
}

}  // namespace pddm

using namespace pddm;

// scratch for the cross-CTA partial sums: forward B*S*G*2 floats, backward (B*S + B)*5*C floats (S <= 16)
extern "C" size_t pddm_gn_silu_fwd_workspace(int32_t B, int32_t G) {
  return static_cast<size_t>(B) * kMaxSplit * G * 2 * sizeof(float);
}
extern "C" size_t pddm_gn_silu_bwd_workspace(int32_t B, int32_t C) {
  return static_cast<size_t>(B) * (kMaxSplit + 1) * 5 * C * sizeof(float);
}

extern "C" int pddm_gn_silu_fwd(const pddm_gn_fwd_params* p, void* workspace, size_t workspace_bytes,
                                   pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->x || !p->y || !p->gamma || !p->beta || !p->mean || !p->rstd || !workspace) return PDDM_ERR_BAD_ARG;
  if ((p->scale == nullptr) != (p->shift == nullptr)) return PDDM_ERR_BAD_ARG;
  GnGeom g;
  int threads;
  int rc = make_geom(p->B, p->HW, p->C, p->G, &g, &threads);
  if (rc) return rc;
  if (!aligned16(p->x) || !aligned16(p->y)) return PDDM_ERR_BAD_ARG;
  if (workspace_bytes < static_cast<size_t>(g.B) * g.S * g.G * 2 * sizeof(float)) return PDDM_ERR_WORKSPACE;
  float* part = static_cast<float*>(workspace);
  dim3 grid(g.S, g.B);
  gn_stats_kernel<<<grid, threads, 2 * g.nlanes * g.C * sizeof(float), s>>>(p->x, p->x_dtype, part, g);
  if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  gn_apply_kernel<<<grid, threads, 2 * g.G * sizeof(float), s>>>(*p, part, g);
  return launch_status();
}

extern "C" int pddm_gn_silu_bwd(const pddm_gn_bwd_params* p, void* workspace, size_t workspace_bytes,
                                pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->x || !p->dy || !p->dx || !p->gamma || !p->beta || !p->mean || !p->rstd || !p->dgamma || !p->dbeta ||
      !workspace)
    return PDDM_ERR_BAD_ARG;
  if ((p->scale == nullptr) != (p->shift == nullptr)) return PDDM_ERR_BAD_ARG;
  GnGeom g;
  int threads;
  int rc = make_geom(p->B, p->HW, p->C, p->G, &g, &threads);
  if (rc) return rc;
  if (!aligned16(p->x) || !aligned16(p->dy) || !aligned16(p->dx)) return PDDM_ERR_BAD_ARG;
  const size_t need = (static_cast<size_t>(g.B) * g.S + g.B) * 5 * g.C * sizeof(float);
  if (workspace_bytes < need) return PDDM_ERR_WORKSPACE;
  float* part = static_cast<float*>(workspace);
  float* tot = part + static_cast<size_t>(g.B) * g.S * 5 * g.C;
  dim3 grid(g.S, g.B);
  gn_bwd_stats_kernel<<<grid, threads, 5 * g.nlanes * g.C * sizeof(float), s>>>(*p, part, g);
  if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  gn_bwd_apply_kernel<<<grid, threads, (2 * g.C + 2 * g.G) * sizeof(float), s>>>(*p, part, tot, g);
  if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  gn_bwd_finalize_kernel<<<(g.C + 127) / 128, 128, 0, s>>>(*p, tot, g);
  return launch_status();
}
