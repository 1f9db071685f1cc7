// GroupNorm(32) (+SiLU) forward / backward as a PERSISTENT, bulk-tensor-pipelined kernel (bf16 NHWC, no scale-shift).
//
// HBM-bound work.  The cluster kernels in groupnorm.cu process one sample per cluster and serialise
// load -> statistics -> exchange -> apply -> store inside every CTA (0.3-0.4 of the HBM roofline measured).  Here:
//
//   * work item = (sample b, chunk of `cc` channels): `cc` is a multiple of the group width, so every group of the
//     chunk is complete inside one item -- no cross-CTA exchange, no cluster, no DSMEM;
//   * one CTA per SM walks its items; a dedicated producer warp keeps `nslots` items in flight with 3-D bulk-tensor
//     loads (box = [cc channels, <=256 rows, 1 sample]; mbarrier complete_tx) and writes results back with
//     bulk-tensor STORES (or reduce-adds) straight from shared memory, so the 16 compute warps only ever touch
//     shared memory and the copy engine sees loads of item j+2 and the store of item j-1 while item j is computed;
//   * the compute warps form NG (2) independent groups that take alternate items, each with its own named barrier
//     and scratch: the serial part of an item (sum fold, group statistics, barriers) of one group overlaps the
//     streaming passes of the other instead of idling the SM;
//   * both passes of an item (statistics, apply) run from the shared-memory slab: HBM traffic is the algorithmic
//     minimum (forward 4 B/element, backward 6 B/element, +2 with a fused residual-gradient add);
//   * optional two-source input (channels [0,C_a) from x, the rest from x2) replaces the materialised th.cat of the
//     UNet's skip connections (src/modules/unet.py:492); the backward writes the two halves of dx to two tensors and
//     can add the gradient of the residual branch (src/modules/unet.py:201,234) on the way out;
//   * per-sample channel sums (dgamma / dbeta partials, sum_hw dx) go to caller-provided [B, ld] matrices: fixed
//     summation order everywhere -> bitwise reproducible.
#include <cuda_bf16.h>

#include "host_common.h"
#include "ptx.cuh"

namespace pddm {

constexpr int kPipeCompute = 512;                // 16 compute warps, split into NG groups
constexpr int kPipeThreads = kPipeCompute + 32;  // + the producer warp
constexpr int kPipeMaxSlots = 4;

struct PipeGeom {
  int B, HW, C, G, cpg;
  int cc, nchunk, chunkA;  // channels per item; items per sample; items of a sample that live in segment A
  int NG, TG;              // compute groups, threads per group
  int U, lanes, active;    // 16-byte units per row; row lanes of a group; lanes * U threads of a group own rows
  int rows_box, nbox, box_bytes, slab_bytes, slot_bytes, ntens;
  int nslots, nitems, pitch;
  int off_grp, grp_bytes;                 // per-group float block: scratch | csum | coef | scr3 | csum3
  int o_csum, o_coef, o_scr3, o_csum3;    // float offsets inside a group block
  int off_bars;
  int dbg;  // PDDM_GN_DBG experiment bits: 1 = no pass 1, 2 = no pass 2, 4 = no sum fold, 8 = no stores
  float inv_n;
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void gbar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// dynamic shared memory, 1024-byte aligned, WITHOUT laundering the pointer through an integer (the compiler must
// keep seeing a shared-memory address, or every slab access becomes a generic LD/ST)
__device__ __forceinline__ uint8_t* pipe_smem(uint8_t* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}

__device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]);
  v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]);
  v.w = pack_bf16(f[6], f[7]);
  return v;
}
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// tanh.approx.f32 (one special-function operation, max relative error 2^-11).  With hz = z/2:
//   silu(z) = z * sigmoid(z) = hz * (1 + tanh(hz)) = fma(hz, tanh(hz), hz)          -- ONE SFU op per element
// against two (ex2 + rcp) for the textbook form; the SFU pipe (16 results/clk/SM) is what bounds the apply pass
// (measured: 88 % busy with 1.5 ops/element).  Error of y: <= |hz| * 2^-11 absolute, i.e. ~1 % of the bf16 output
// rounding in the rms sense (it exceeds the rounding only for z < -2, where |y| < 0.24).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Fold `pairs` columns of a group's reduction scratch [warps of the group][pairs] over its rows into csum[pairs]
// (fixed order).  Called by all threads of the group after they have parked their sums.
__device__ __forceinline__ void fold_rows(const float* scratch, float* csum, int pairs, int t, int TG, int bar_id) {
  gbar(bar_id, TG);
  const int wrows = TG >> 5;
  for (int i = t; i < pairs; i += TG) {
    float v = 0.f;
    for (int r = 0; r < wrows; r += 4) {
      const float a = scratch[(r + 0) * pairs + i], b = scratch[(r + 1) * pairs + i];
      const float c = scratch[(r + 2) * pairs + i], d = scratch[(r + 3) * pairs + i];
      v += a; v += b; v += c; v += d;
    }
    csum[i] = v;
  }
  gbar(bar_id, TG);
}

// Per-thread sums acc[NV][8] (8 channels of unit u = t % U over this thread's rows) -> one scratch row per warp.
// Inside a warp the threads that own the same unit sit U lanes apart; they are summed with a shuffle-down tree in
// steps of U, 2U, 4U ... lanes (fixed order; any U, not only powers of two), after which lanes 0..U-1 hold the warp's
// totals for units (t & ~31 + lane) % U.  Threads without rows carry zeros.  (TG is a multiple of 32.)
template <int NV>
__device__ __forceinline__ void park_sums(float (&acc)[NV][8], float* scratch, const PipeGeom& g, int t) {
  const int pairs = NV * g.cc, lane = t & 31;
  // (the step loop is the OUTER one: the NV*8 shuffles of a step are independent and pipeline; a per-value loop
  //  would serialise NV*8 chains of dependent shuffles)
  for (int off = g.U; off < 32; off <<= 1) {
    const bool ok = lane + off < 32;
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = __shfl_down_sync(0xffffffffu, acc[q][j], off);
        if (ok) acc[q][j] += o;
      }
  }
  if (lane < g.U) {
    float* d = scratch + (t >> 5) * pairs + (t % g.U) * 8;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      *reinterpret_cast<float4*>(d + q * g.cc) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
      *reinterpret_cast<float4*>(d + q * g.cc + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
    }
  }
}

struct PipeItem {
  int b, segB, c_seg, c_virt;  // sample, segment flag, first channel inside the segment / the virtual tensor
};
__device__ __forceinline__ PipeItem pipe_item(const PipeGeom& g, int j) {
  const int id = blockIdx.x + j * gridDim.x;
  PipeItem it;
  it.b = id / g.nchunk;
  const int k = id - it.b * g.nchunk;
  it.segB = k >= g.chunkA;
  it.c_virt = k * g.cc;
  it.c_seg = it.segB ? (k - g.chunkA) * g.cc : it.c_virt;
  return it;
}
__device__ __forceinline__ int pipe_my_items(const PipeGeom& g) {
  return (g.nitems - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
}
__device__ __forceinline__ void pipe_init_barriers(const PipeGeom& g, uint64_t* full, uint64_t* done) {
  for (int s = 0; s < g.nslots; ++s) {
    mbar_init(&full[s], 1);
    mbar_init(&done[s], g.TG / 32);
  }
  fence_mbar_init();
}

// ------------------------------------------------------------------------------------------------ forward
struct PipeFwdArgs {
  const float* gamma;
  const float* beta;
  float* mean;
  float* rstd;
  float eps;
};

template <bool SILU>
__global__ void __launch_bounds__(kPipeThreads, 1)
gn_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmXb,
                   const __grid_constant__ CUtensorMap tmY, const __grid_constant__ PipeFwdArgs a,
                   const __grid_constant__ PipeGeom g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = pipe_smem(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.off_bars);
  uint64_t* done = full + kPipeMaxSlots;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) pipe_init_barriers(g, full, done);
  if (warp == kPipeCompute / 32 && (tid & 31) == 0) {
    tma_prefetch_desc(&tmXa);
    tma_prefetch_desc(&tmXb);
    tma_prefetch_desc(&tmY);
  }
  __syncthreads();
  pdl_entry();
  const int n_my = pipe_my_items(g);

  if (warp == kPipeCompute / 32) {
    // ------------------------------------------------------------ producer: bulk-tensor loads and stores
    if ((tid & 31) == 0) {
      auto load = [&](int j) {
        const PipeItem it = pipe_item(g, j);
        const int s = j % g.nslots;
        uint8_t* dst = smem + s * g.slot_bytes;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(g.nbox * g.box_bytes));
        const CUtensorMap* m = it.segB ? &tmXb : &tmXa;
        for (int bx = 0; bx < g.nbox; ++bx)
          tma_load_3d(dst + bx * g.box_bytes, m, &full[s], it.c_seg, bx * g.rows_box, it.b);
      };
      const int pre = n_my < g.nslots - 1 ? n_my : g.nslots - 1;
      for (int j = 0; j < pre; ++j) load(j);
      for (int j = 0; j < n_my; ++j) {
        const int nxt = j + g.nslots - 1;
        if (nxt < n_my) {
          if (j >= 1) bulk_wait_group_read<0>();  // the store of item j-1 has finished reading that slot
          load(nxt);
        }
        const int s = j % g.nslots;
        mbar_wait(&done[s], (j / g.nslots) & 1);
        const PipeItem it = pipe_item(g, j);
        const uint8_t* src = smem + s * g.slot_bytes;
        if (!(g.dbg & 8))
          for (int bx = 0; bx < g.nbox; ++bx) tma_store_3d(&tmY, src + bx * g.box_bytes, it.c_virt, bx * g.rows_box, it.b);
        bulk_commit_group();
      }
      bulk_wait_group<0>();
    }
    return;
  }

  // -------------------------------------------------------------- compute groups (alternate items)
  const int grp = tid / g.TG, t = tid - grp * g.TG, bar_id = 1 + grp, TG = g.TG;
  float* gf = reinterpret_cast<float*>(smem + g.off_grp + grp * g.grp_bytes);
  float* scratch = gf;
  float* csum = gf + g.o_csum;
  float* coef = gf + g.o_coef;
  const int u = t % g.U, lane_row = t / g.U;
  const bool owner = t < g.active;
  const int pitch = g.pitch, HW = g.HW, lanes = g.lanes, cc = g.cc;
  for (int j = grp; j < n_my; j += g.NG) {
    const PipeItem it = pipe_item(g, j);
    const int s = j % g.nslots;
    uint8_t* slab = smem + s * g.slot_bytes + u * 16;
    // affine parameters of the channel this thread finalises (loaded before the slab is awaited: latency hidden)
    float p_gamma = 0.f, p_beta = 0.f;
    if (t < cc) {
      p_gamma = __ldg(a.gamma + it.c_virt + t);
      p_beta = __ldg(a.beta + it.c_virt + t);
    }
    mbar_wait(&full[s], (j / g.nslots) & 1);
    // ---- pass 1: per-channel sum and sum of squares of this thread's rows
    float acc[2][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[0][q] = acc[1][q] = 0.f;
    if (owner && !(g.dbg & 1)) {
      int r = lane_row;
      for (; r + 3 * lanes < HW; r += 4 * lanes) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const uint4*>(slab + (r + k * lanes) * pitch);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float f[8];
          unpack8f(v[k], f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            acc[0][q] += f[q];
            acc[1][q] = fmaf(f[q], f[q], acc[1][q]);
          }
        }
      }
      for (; r < HW; r += lanes) {
        float f[8];
        unpack8f(*reinterpret_cast<const uint4*>(slab + r * pitch), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc[0][q] += f[q];
          acc[1][q] = fmaf(f[q], f[q], acc[1][q]);
        }
      }
    }
    if (!(g.dbg & 4)) {
      park_sums<2>(acc, scratch, g, t);
      fold_rows(scratch, csum, 2 * cc, t, TG, bar_id);
    }
    // ---- group statistics -> per-channel y = x * ca + cb   (cc <= 256 <= TG: one channel per thread)
    if (t < cc) {
      const int g0 = (t / g.cpg) * g.cpg;
      float sa = 0.f, sq = 0.f;
      for (int k = 0; k < g.cpg; ++k) {
        sa += csum[g0 + k];
        sq += csum[cc + g0 + k];
      }
      const float mean = sa * g.inv_n;
      const float var = fmaxf(sq * g.inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      const float ca = rstd * p_gamma, cb = p_beta - mean * ca;
      coef[t] = SILU ? 0.5f * ca : ca;  // with SiLU the apply pass works on hz = z/2 (see tanh_approx)
      coef[cc + t] = SILU ? 0.5f * cb : cb;
      if (t == g0) {
        const int gi = it.b * g.G + (it.c_virt + t) / g.cpg;
        a.mean[gi] = mean;
        a.rstd[gi] = rstd;
      }
    }
    gbar(bar_id, TG);
    // ---- pass 2: normalise (+SiLU) in place
    if (owner && !(g.dbg & 2)) {
      float ca[8], cb[8];
      ld8f(coef + u * 8, ca);
      ld8f(coef + cc + u * 8, cb);
      for (int r = lane_row; r < HW; r += 2 * lanes) {
        const bool two = r + lanes < HW;
        uint4 v0 = *reinterpret_cast<const uint4*>(slab + r * pitch), v1 = v0;
        if (two) v1 = *reinterpret_cast<const uint4*>(slab + (r + lanes) * pitch);
        float f[8], h[8];
        unpack8f(v0, f);
        unpack8f(v1, h);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          f[q] = fmaf(f[q], ca[q], cb[q]);
          h[q] = fmaf(h[q], ca[q], cb[q]);
        }
        if (SILU) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            f[q] = fmaf(f[q], tanh_approx(f[q]), f[q]);
            h[q] = fmaf(h[q], tanh_approx(h[q]), h[q]);
          }
        }
        *reinterpret_cast<uint4*>(slab + r * pitch) = pack8f(f);
        if (two) *reinterpret_cast<uint4*>(slab + (r + lanes) * pitch) = pack8f(h);
      }
    }
    fence_proxy_async();  // generic-proxy writes to the slab -> visible to the bulk-tensor store
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&done[s]);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// With xh = (x-mean)*rstd, z = gamma*xh + beta, y = act(z)  (notation of groupnorm.cu):
//   dz = dy*act'(z);  A_c = sum_hw dz;  Bq_c = sum_hw dz*xh;  Xh_c = sum_hw xh
//   S1_g = sum_{c in g} gamma_c A_c;  S2_g = sum_{c in g} gamma_c Bq_c;  n = cpg*HW
//   dx = rstd*(gamma_c*dz - (S1_g + xh*S2_g)/n) [+ gres];   dbeta_c = sum_b A_c;  dgamma_c = sum_b Bq_c
//   sum_hw dx = rstd*(gamma_c*A_c - (HW*S1_g + S2_g*Xh_c)/n)      (without gres: analytic, no extra pass)
struct PipeBwdArgs {
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* rstd;
  float* part_dgamma;
  float* part_dbeta;
  float* csA;
  float* csB;
  int ld_part, ld_csA, ld_csB, acc_csA, acc_csB;
  int dxb_c0;  // channel coordinate offset of segment B inside the tmDb tensor (C_a when dx is one tensor)
  int dx_accA, dx_accB;
};

template <bool SILU, bool GRES>
__global__ void __launch_bounds__(kPipeThreads, 1)
gn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmXa, const __grid_constant__ CUtensorMap tmXb,
                   const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmGR,
                   const __grid_constant__ CUtensorMap tmDa, const __grid_constant__ CUtensorMap tmDb,
                   const __grid_constant__ PipeBwdArgs a, const __grid_constant__ PipeGeom g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = pipe_smem(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.off_bars);
  uint64_t* done = full + kPipeMaxSlots;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) pipe_init_barriers(g, full, done);
  if (warp == kPipeCompute / 32 && (tid & 31) == 0) {
    tma_prefetch_desc(&tmXa);
    tma_prefetch_desc(&tmXb);
    tma_prefetch_desc(&tmDY);
    if (GRES) tma_prefetch_desc(&tmGR);
    tma_prefetch_desc(&tmDa);
    tma_prefetch_desc(&tmDb);
  }
  __syncthreads();
  pdl_entry();
  const int n_my = pipe_my_items(g);

  if (warp == kPipeCompute / 32) {
    if ((tid & 31) == 0) {
      auto load = [&](int j) {
        const PipeItem it = pipe_item(g, j);
        const int s = j % g.nslots;
        uint8_t* dst = smem + s * g.slot_bytes;
        mbar_expect_tx(&full[s], static_cast<uint32_t>(g.ntens * g.nbox * g.box_bytes));
        const CUtensorMap* m = it.segB ? &tmXb : &tmXa;
        for (int bx = 0; bx < g.nbox; ++bx) {
          tma_load_3d(dst + bx * g.box_bytes, m, &full[s], it.c_seg, bx * g.rows_box, it.b);
          tma_load_3d(dst + g.slab_bytes + bx * g.box_bytes, &tmDY, &full[s], it.c_virt, bx * g.rows_box, it.b);
          if (GRES)
            tma_load_3d(dst + 2 * g.slab_bytes + bx * g.box_bytes, &tmGR, &full[s], it.c_virt, bx * g.rows_box, it.b);
        }
      };
      const int pre = n_my < g.nslots - 1 ? n_my : g.nslots - 1;
      for (int j = 0; j < pre; ++j) load(j);
      for (int j = 0; j < n_my; ++j) {
        const int nxt = j + g.nslots - 1;
        if (nxt < n_my) {
          if (j >= 1) bulk_wait_group_read<0>();
          load(nxt);
        }
        const int s = j % g.nslots;
        mbar_wait(&done[s], (j / g.nslots) & 1);
        const PipeItem it = pipe_item(g, j);
        const uint8_t* src = smem + s * g.slot_bytes + g.slab_bytes;  // dx was written over the dy slab
        const CUtensorMap* m = it.segB ? &tmDb : &tmDa;
        const int c0 = it.segB ? it.c_seg + a.dxb_c0 : it.c_seg;
        const bool accum = it.segB ? a.dx_accB : a.dx_accA;
        if (!(g.dbg & 8))
          for (int bx = 0; bx < g.nbox; ++bx) {
            if (accum) tma_reduce_add_3d(m, src + bx * g.box_bytes, c0, bx * g.rows_box, it.b);
            else tma_store_3d(m, src + bx * g.box_bytes, c0, bx * g.rows_box, it.b);
          }
        bulk_commit_group();
      }
      bulk_wait_group<0>();
    }
    return;
  }

  const int grp = tid / g.TG, t = tid - grp * g.TG, bar_id = 1 + grp, TG = g.TG;
  float* gf = reinterpret_cast<float*>(smem + g.off_grp + grp * g.grp_bytes);
  float* scratch = gf;
  float* csum = gf + g.o_csum;
  float* coef = gf + g.o_coef;
  float* scr3 = gf + g.o_scr3;
  float* csum3 = gf + g.o_csum3;
  const int u = t % g.U, lane_row = t / g.U;
  const bool owner = t < g.active;
  const int pitch = g.pitch, HW = g.HW, lanes = g.lanes, cc = g.cc, cpg = g.cpg;
  for (int j = grp; j < n_my; j += g.NG) {
    const PipeItem it = pipe_item(g, j);
    const int s = j % g.nslots;
    const uint8_t* slab_x = smem + s * g.slot_bytes + u * 16;
    uint8_t* slab_d = smem + s * g.slot_bytes + g.slab_bytes + u * 16;
    const uint8_t* slab_g = smem + s * g.slot_bytes + 2 * g.slab_bytes + u * 16;
    // per-channel constants of this thread's 8 channels (global loads issued before the wait on the slab)
    float xa[8], xc[8], gam[8], bet[8];
    {
      const int cv = it.c_virt + u * 8;
      ld8f(a.gamma + cv, gam);
      ld8f(a.beta + cv, bet);
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // hz = z/2 = xh * (gamma/2) + beta/2
        gam[q] *= 0.5f;
        bet[q] *= 0.5f;
      }
      int gi = cv / cpg, rem = cv - gi * cpg;
      gi += it.b * g.G;
      float mu = __ldg(a.mean + gi), rs = __ldg(a.rstd + gi);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        xa[q] = rs;
        xc[q] = -mu * rs;
        if (++rem == cpg && q < 7) {
          rem = 0;
          ++gi;
          mu = __ldg(a.mean + gi);
          rs = __ldg(a.rstd + gi);
        }
      }
    }
    // the channel this thread finalises after the fold
    float p_gamma = 0.f, p_mu = 0.f, p_rs = 0.f;
    if (t < cc) {
      const int cv = it.c_virt + t, gi = it.b * g.G + cv / cpg;
      p_gamma = __ldg(a.gamma + cv);
      p_mu = __ldg(a.mean + gi);
      p_rs = __ldg(a.rstd + gi);
    }
    mbar_wait(&full[s], (j / g.nslots) & 1);
    // ---- pass 1: dz = dy * act'(z) parked over dy (bf16), channel sums A, Bq, Xh
    float acc[3][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[0][q] = acc[1][q] = acc[2][q] = 0.f;
    if (owner && !(g.dbg & 1)) {
      // one row: dz = dy * act'(z) parked over dy, sums A += dz, Bq += dz*xh, Xh += xh
      //   act'(z) = s*(1 + z*(1-s)) with s = (1+t)/2, t = tanh(z/2), hz = z/2:  act' = s + s*hz*(1-t)
      auto row = [&](int r) {
        float f[8], e[8];
        unpack8f(*reinterpret_cast<const uint4*>(slab_x + r * pitch), f);
        unpack8f(*reinterpret_cast<const uint4*>(slab_d + r * pitch), e);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float xh = fmaf(f[q], xa[q], xc[q]);
          float dz = e[q];
          if (SILU) {
            const float hz = fmaf(xh, gam[q], bet[q]);
            const float th = tanh_approx(hz);
            const float sg = fmaf(0.5f, th, 0.5f);
            dz *= fmaf(sg, fmaf(-hz, th, hz), sg);
          }
          acc[0][q] += dz;
          acc[1][q] = fmaf(dz, xh, acc[1][q]);
          acc[2][q] += xh;
          e[q] = dz;
        }
        *reinterpret_cast<uint4*>(slab_d + r * pitch) = pack8f(e);
      };
      int r = lane_row;
      for (; r + lanes < HW; r += 2 * lanes) {  // two independent rows per iteration
        row(r);
        row(r + lanes);
      }
      if (r < HW) row(r);
    }
    if (!(g.dbg & 4)) {
      park_sums<3>(acc, scratch, g, t);
      fold_rows(scratch, csum, 3 * cc, t, TG, bar_id);
    }
    // ---- group sums -> dx = dz*da + x*db + dc ; per-sample partial sums out
    if (t < cc) {
      const int g0 = (t / cpg) * cpg;
      const int cv = it.c_virt + t;
      float s1 = 0.f, s2 = 0.f;
      for (int k = 0; k < cpg; ++k) {
        const float gm = __ldg(a.gamma + it.c_virt + g0 + k);
        s1 = fmaf(gm, csum[g0 + k], s1);
        s2 = fmaf(gm, csum[cc + g0 + k], s2);
      }
      s1 *= g.inv_n;
      s2 *= g.inv_n;
      const float A = csum[t], Bq = csum[cc + t], Xh = csum[2 * cc + t];
      coef[t] = p_rs * p_gamma;
      coef[cc + t] = -p_rs * s2 * p_rs;
      coef[2 * cc + t] = -p_rs * (s1 - s2 * p_mu * p_rs);
      if (a.part_dbeta) a.part_dbeta[static_cast<size_t>(it.b) * a.ld_part + cv] = A;
      if (a.part_dgamma) a.part_dgamma[static_cast<size_t>(it.b) * a.ld_part + cv] = Bq;
      if (!GRES) {
        float* cs = it.segB ? a.csB : a.csA;
        if (cs) {
          float* d = cs + static_cast<size_t>(it.b) * (it.segB ? a.ld_csB : a.ld_csA) + it.c_seg + t;
          const float v = p_rs * (p_gamma * A - (static_cast<float>(HW) * s1 + s2 * Xh));
          *d = (it.segB ? a.acc_csB : a.acc_csA) ? *d + v : v;
        }
      }
    }
    gbar(bar_id, TG);
    // ---- pass 2
    float cs_acc[1][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) cs_acc[0][q] = 0.f;
    if (owner && !(g.dbg & 2)) {
      float da[8], db[8], dc[8];
      ld8f(coef + u * 8, da);
      ld8f(coef + cc + u * 8, db);
      ld8f(coef + 2 * cc + u * 8, dc);
      for (int r = lane_row; r < HW; r += lanes) {
        float f[8], d[8];
        unpack8f(*reinterpret_cast<const uint4*>(slab_x + r * pitch), f);
        unpack8f(*reinterpret_cast<const uint4*>(slab_d + r * pitch), d);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = fmaf(d[q], da[q], fmaf(f[q], db[q], dc[q]));
        if (GRES) {
          float gr[8];
          unpack8f(*reinterpret_cast<const uint4*>(slab_g + r * pitch), gr);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            d[q] += gr[q];
            cs_acc[0][q] += d[q];
          }
        }
        *reinterpret_cast<uint4*>(slab_d + r * pitch) = pack8f(d);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive(&done[s]);
    if (GRES) {
      // sum_hw (dx + gres) has no closed form: fold the pass-2 sums (own scratch region: the group's next item may
      // already be parking its pass-1 sums in `scratch` while slower warps are still here)
      float* cs = it.segB ? a.csB : a.csA;
      if (cs) {  // uniform over the CTA
        park_sums<1>(cs_acc, scr3, g, t);
        fold_rows(scr3, csum3, cc, t, TG, bar_id);
        if (t < cc) {
          float* d = cs + static_cast<size_t>(it.b) * (it.segB ? a.ld_csB : a.ld_csA) + it.c_seg + t;
          *d = (it.segB ? a.acc_csB : a.acc_csA) ? *d + csum3[t] : csum3[t];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// Geometry of the persistent kernels for `ntens` resident tensors per item and `nq` summed quantities per channel
// (`extra_q` = 1 reserves the second scratch region of the backward's gres path).  Returns false if the shape does
// not qualify (the caller then uses the cluster kernels of groupnorm.cu).
bool make_pipe_geom(int B, int HW, int C, int G, int C_a, int ntens, int nq, int extra_q, PipeGeom* g, size_t* smem) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) return false;
  const int cpg = C / G;
  const int base = cpg / gcd_i(cpg, 8) * 8;  // lcm(cpg, 8)
  if (C % base) return false;
  if (C_a <= 0 || C_a >= C) C_a = C;
  if (C_a % base || (C - C_a) % base) return false;
  const int budget = device_info().max_smem_optin > 0 ? device_info().max_smem_optin - 1024 : 232448 - 1024;
  const int rows_box = HW < 256 ? HW : 256;
  const int nbox = (HW + rows_box - 1) / rows_box;
  const int target = ntens == 1 ? 64 * 1024 : 32 * 1024;
  bool have = false;
  double best_score = 0;
  for (int cc = base; cc <= 256 && cc <= C; cc += base) {
    if (C % cc || C_a % cc) continue;
    if (env_knobs().gn_cc > 0 && cc != env_knobs().gn_cc) continue;
    PipeGeom t;
    t.B = B; t.HW = HW; t.C = C; t.G = G; t.cpg = cpg;
    t.cc = cc; t.nchunk = C / cc; t.chunkA = C_a / cc;
    t.U = cc / 8;
    t.NG = env_knobs().gn_ng == 1 ? 1 : 2;
    t.TG = kPipeCompute / t.NG;
    if (cc > t.TG) continue;
    t.lanes = t.TG / t.U;
    if (t.lanes > HW) t.lanes = HW;
    t.active = t.lanes * t.U;
    t.rows_box = rows_box; t.nbox = nbox;
    t.box_bytes = rows_box * cc * 2;
    if (nbox > 1 && t.box_bytes % 128) continue;
    t.slab_bytes = (nbox * t.box_bytes + 127) / 128 * 128;
    t.ntens = ntens;
    t.slot_bytes = (ntens * t.slab_bytes + 1023) / 1024 * 1024;
    t.pitch = cc * 2;
    const int pairs = nq * cc, wrows = t.TG / 32;
    const int fl_scratch = wrows * pairs, fl_csum = pairs, fl_coef = 4 * cc;
    const int fl_scr3 = extra_q ? wrows * cc : 0, fl_csum3 = extra_q ? cc : 0;
    t.o_csum = fl_scratch;
    t.o_coef = t.o_csum + fl_csum;
    t.o_scr3 = t.o_coef + fl_coef;
    t.o_csum3 = t.o_scr3 + fl_scr3;
    t.grp_bytes = (t.o_csum3 + fl_csum3) * 4;
    const int fixed = t.NG * t.grp_bytes + 2 * kPipeMaxSlots * 8 + 64;
    int nslots = (budget - fixed) / t.slot_bytes;
    if (nslots > kPipeMaxSlots) nslots = kPipeMaxSlots;
    if (nslots < 2) continue;
    t.nslots = nslots;
    t.nitems = B * t.nchunk;
    t.inv_n = 1.f / (static_cast<float>(cpg) * HW);
    t.dbg = env_knobs().gn_dbg > 0 ? env_knobs().gn_dbg : 0;
    int off = nslots * t.slot_bytes;
    t.off_grp = off; off += t.NG * t.grp_bytes;
    off = (off + 15) / 16 * 16;
    t.off_bars = off; off += 2 * kPipeMaxSlots * 8;
    // score: three or more slots first, then rows of >= 64 B, then a slab close to the target size
    double ratio = static_cast<double>(t.slab_bytes) / target;
    if (ratio < 1) ratio = 1 / ratio;
    const double score = (nslots >= 3 ? 100 : 0) + (cc * 2 >= 64 ? 10 : 0) - ratio;
    if (!have || score > best_score) {
      have = true;
      best_score = score;
      *g = t;
      *smem = static_cast<size_t>(off) + 1024;
    }
  }
  return have;
}

static int make_map3(CUtensorMap* m, const void* base, int Cseg, int HW, int B, int ld, int cc, int rows_box) {
  const uint64_t dims[3] = {static_cast<uint64_t>(Cseg), static_cast<uint64_t>(HW), static_cast<uint64_t>(B)};
  const uint64_t str[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(HW) * ld * 2};
  const uint32_t box[3] = {static_cast<uint32_t>(cc), static_cast<uint32_t>(rows_box), 1u};
  return make_tmap_bf16(m, base, 3, dims, str, box, 0);
}

static int pipe_grid(const PipeGeom& g) {
  const int sms = launch_sms();
  return g.nitems < sms ? g.nitems : sms;
}

// returns PDDM_OK / an error, or 1 if the shape does not qualify for the persistent kernel
int gn_fwd_pipe(const pddm_gn_fwd_params* p, cudaStream_t s) {
  if (p->scale || p->x_dtype != PDDM_BF16) return 1;
  const int C_a = (p->x2 && p->C_a > 0 && p->C_a < p->C) ? p->C_a : p->C;
  const int ldx = p->ldx > 0 ? p->ldx : C_a, ldx2 = p->ldx2 > 0 ? p->ldx2 : p->C - C_a, ldy = p->ldy > 0 ? p->ldy : p->C;
  if (ldx % 8 || ldx2 % 8 || ldy % 8 || ldx < C_a || ldy < p->C) return PDDM_ERR_UNSUPPORTED;
  PipeGeom g;
  size_t smem;
  if (!make_pipe_geom(p->B, p->HW, p->C, p->G, C_a, 1, 2, 0, &g, &smem)) return 1;
  if (!aligned16(p->x) || !aligned16(p->y) || (p->x2 && !aligned16(p->x2)) || !aligned16(p->gamma) || !aligned16(p->beta))
    return PDDM_ERR_BAD_ARG;
  CUtensorMap tmXa, tmXb, tmY;
  int rc = make_map3(&tmXa, p->x, C_a, p->HW, p->B, ldx, g.cc, g.rows_box);
  if (rc) return rc;
  if (C_a < p->C) rc = make_map3(&tmXb, p->x2, p->C - C_a, p->HW, p->B, ldx2, g.cc, g.rows_box);
  else tmXb = tmXa;
  if (rc) return rc;
  rc = make_map3(&tmY, p->y, p->C, p->HW, p->B, ldy, g.cc, g.rows_box);
  if (rc) return rc;
  PipeFwdArgs a;
  a.gamma = p->gamma; a.beta = p->beta; a.mean = p->mean; a.rstd = p->rstd; a.eps = p->eps;
  auto fn = p->silu ? gn_fwd_pipe_kernel<true> : gn_fwd_pipe_kernel<false>;
  if (ensure_smem_optin(reinterpret_cast<const void*>(fn))) return PDDM_ERR_CUDA;
  PdlLaunch(pipe_grid(g), kPipeThreads, smem, s)(fn, tmXa, tmXb, tmY, a, g);
  return launch_status();
}

int gn_bwd_pipe(const pddm_gn_bwd_params* p, float* part_dgamma, float* part_dbeta, int ld_part, cudaStream_t s) {
  if (p->scale || p->x_dtype != PDDM_BF16 || p->dx_dtype != PDDM_BF16) return 1;
  const int C = p->C;
  const int C_a = (p->x2 && p->C_a > 0 && p->C_a < C) ? p->C_a : C;
  const int ldx = p->ldx > 0 ? p->ldx : C_a, ldx2 = p->ldx2 > 0 ? p->ldx2 : C - C_a, lddy = p->lddy > 0 ? p->lddy : C;
  const int ld_gres = p->ld_gres > 0 ? p->ld_gres : C;
  const bool split_dx = p->dx2 != nullptr && C_a < C;
  const int ld_dx = p->ld_dx > 0 ? p->ld_dx : (split_dx ? C_a : C), ld_dx2 = p->ld_dx2 > 0 ? p->ld_dx2 : C - C_a;
  if (ldx % 8 || ldx2 % 8 || lddy % 8 || ld_gres % 8 || ld_dx % 8 || ld_dx2 % 8) return PDDM_ERR_UNSUPPORTED;
  const bool gres = p->gres != nullptr;
  PipeGeom g;
  size_t smem;
  if (!make_pipe_geom(p->B, p->HW, C, p->G, C_a, gres ? 3 : 2, 3, gres ? 1 : 0, &g, &smem)) return 1;
  if (!aligned16(p->x) || !aligned16(p->dy) || !aligned16(p->dx) || (p->x2 && !aligned16(p->x2)) ||
      (gres && !aligned16(p->gres)) || (p->dx2 && !aligned16(p->dx2)) || !aligned16(p->gamma) || !aligned16(p->beta))
    return PDDM_ERR_BAD_ARG;
  CUtensorMap tmXa, tmXb, tmDY, tmGR, tmDa, tmDb;
  int rc = make_map3(&tmXa, p->x, C_a, p->HW, p->B, ldx, g.cc, g.rows_box);
  if (rc) return rc;
  if (C_a < C) rc = make_map3(&tmXb, p->x2, C - C_a, p->HW, p->B, ldx2, g.cc, g.rows_box);
  else tmXb = tmXa;
  if (rc) return rc;
  rc = make_map3(&tmDY, p->dy, C, p->HW, p->B, lddy, g.cc, g.rows_box);
  if (rc) return rc;
  if (gres) rc = make_map3(&tmGR, p->gres, C, p->HW, p->B, ld_gres, g.cc, g.rows_box);
  else tmGR = tmDY;
  if (rc) return rc;
  PipeBwdArgs a;
  if (split_dx) {
    rc = make_map3(&tmDa, p->dx, C_a, p->HW, p->B, ld_dx, g.cc, g.rows_box);
    if (rc) return rc;
    rc = make_map3(&tmDb, p->dx2, C - C_a, p->HW, p->B, ld_dx2, g.cc, g.rows_box);
    if (rc) return rc;
    a.dxb_c0 = 0;
    a.dx_accB = p->dx2_accumulate;
  } else {
    rc = make_map3(&tmDa, p->dx, C, p->HW, p->B, ld_dx, g.cc, g.rows_box);
    if (rc) return rc;
    tmDb = tmDa;
    a.dxb_c0 = C_a;
    a.dx_accB = p->dx_accumulate;
  }
  a.dx_accA = p->dx_accumulate;
  a.gamma = p->gamma; a.beta = p->beta; a.mean = p->mean; a.rstd = p->rstd;
  a.part_dgamma = part_dgamma; a.part_dbeta = part_dbeta; a.ld_part = ld_part;
  a.csA = p->dx_colsum;
  a.ld_csA = p->ld_colsum > 0 ? p->ld_colsum : (p->dx_colsum2 ? C_a : C);
  a.acc_csA = p->colsum_accumulate;
  if (p->dx_colsum2) {
    a.csB = p->dx_colsum2;
    a.ld_csB = p->ld_colsum2 > 0 ? p->ld_colsum2 : C - C_a;
    a.acc_csB = p->colsum2_accumulate;
  } else {
    a.csB = p->dx_colsum ? p->dx_colsum + C_a : nullptr;
    a.ld_csB = a.ld_csA;
    a.acc_csB = p->colsum_accumulate;
  }
  void (*fn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, PipeBwdArgs, PipeGeom) =
      p->silu ? (gres ? gn_bwd_pipe_kernel<true, true> : gn_bwd_pipe_kernel<true, false>)
              : (gres ? gn_bwd_pipe_kernel<false, true> : gn_bwd_pipe_kernel<false, false>);
  if (ensure_smem_optin(reinterpret_cast<const void*>(fn))) return PDDM_ERR_CUDA;
  PdlLaunch(pipe_grid(g), kPipeThreads, smem, s)(fn, tmXa, tmXb, tmDY, tmGR, tmDa, tmDb, a, g);
  return launch_status();
}

}  // namespace pddm

extern "C" int pddm_gn_pipe_slots(int32_t B, int32_t HW, int32_t C, int32_t G, int32_t C_a, int32_t ntens) {
  pddm::PipeGeom g;
  size_t smem;
  if (ntens < 1 || ntens > 3) return 0;
  const int nq = ntens == 1 ? 2 : 3;
  if (!pddm::make_pipe_geom(B, HW, C, G, C_a, ntens, nq, ntens == 3 ? 1 : 0, &g, &smem)) return 0;
  return g.nslots;
}
