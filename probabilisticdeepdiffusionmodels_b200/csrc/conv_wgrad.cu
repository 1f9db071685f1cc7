// Convolution weight gradient on tcgen05 tensor cores:
//
//   dW[n, tap, c] = sum_{pixels p} dY[p, n] * X[p + offset(tap), c]
//
// A GEMM whose reduction dimension is the pixel index, so both operands are "MN-major": the NHWC tensors are
// loaded as [pixel rows][64 channels = 128 B] boxes by TMA (128B swizzle) and consumed by UMMA with the
// transpose bits set in the instruction descriptor.  Work item = (tap, 128-wide Cout tile, <=256-wide Cin tile,
// K-split); each item accumulates its pixel range in tensor memory and writes an fp32 partial tile; a second
// kernel reduces the splits in a fixed order (deterministic, no atomics) into the parameter layout.
#include "host_common.h"
#include "ptx.cuh"

namespace pddm {

struct WgradKArgs {
  float* partial;  // [splits][Cout][ntaps][Cin]
  int B, H, W, Cin, Cout;
  int BW, BH, BB, tiles_w, tiles_h, k_blocks;  // pixel boxes of <= 64 rows
  int block_n, n_chunks, m_tiles, n_tiles, splits, kb_per_split, ntaps;
  int tap_db[PDDM_MAX_TAPS], tap_dh[PDDM_MAX_TAPS], tap_dw[PDDM_MAX_TAPS];
  int stages, b_bytes, tx_bytes;
  uint32_t idesc, tmem_cols;
};

constexpr int kWgThreads = 256;
constexpr int kWgMaxStages = 8;
constexpr int kChunkBytes = 64 * 128;  // 64 pixel rows x 64 channels bf16

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ WgradKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 2 * kChunkBytes;
  const int stage_bytes = a_bytes + a.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgMaxStages;
  uint64_t* tmem_full = bars + 2 * kWgMaxStages;
  uint64_t* tmem_empty = bars + 2 * kWgMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = a.ntaps * a.m_tiles * a.n_tiles * a.splits;

  // Zero the operand ring once: pixel boxes with fewer than 64 rows leave the tail rows untouched, and those
  // rows must contribute 0 to the reduction.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = a.stages * stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // PDL: barriers, TMEM and the tensor-map prefetch were set up under the previous kernel's tail.  The trigger
  // comes after our TMEM allocation so a dependent CTA can never take tensor memory this grid still needs.
  pdl_launch_dependents();
  pdl_wait();

  // item -> (split, n_tile, m_tile, tap): consecutive CTAs share the dY tile (same m_tile/split) across taps
  auto decode = [&](int item, int& tap, int& m_tile, int& n_tile, int& split) {
    tap = item % a.ntaps;
    item /= a.ntaps;
    n_tile = item % a.n_tiles;
    item /= a.n_tiles;
    m_tile = item % a.m_tiles;
    split = item / a.m_tiles;
  };

  if (warp >= 1 && warp <= 3) {
    // TMA producers: a K-block needs 2 + n_chunks bulk-tensor copies (~200 cycles of issue latency each on one
    // thread), far more than the 512 cycles of tensor work it feeds, so K-blocks are dealt round-robin to three
    // single-lane producers, each walking its own (item, K-block, stage, parity) counters.
    // (warp-uniform loop, issue predicated on one elected lane: see conv_fwd.cu)
    const int pid = warp - 1;
    const int nstages = a.stages, n_chunks = a.n_chunks, block_n = a.block_n;
    const int tiles_w = a.tiles_w, tiles_h = a.tiles_h, BB = a.BB, BH = a.BH, BW = a.BW;
    const uint32_t tx_bytes = a.tx_bytes;
    int stage = pid % nstages;
    uint32_t phase = (pid / nstages) & 1;
    int skip = pid;  // K-blocks of the current item that belong to the other producers before ours
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int tap, m_tile, n_tile, split;
      decode(item, tap, m_tile, n_tile, split);
      const int kb0 = split * a.kb_per_split;
      const int kb1 = min(kb0 + a.kb_per_split, a.k_blocks);
      const int dw = a.tap_dw[tap], dh = a.tap_dh[tap], db = a.tap_db[tap];
      int kb = kb0 + skip;
      for (; kb < kb1; kb += 3) {
        const int tw = kb % tiles_w;
        const int th = (kb / tiles_w) % tiles_h;
        const int tb = kb / (tiles_w * tiles_h);
        const int b0 = tb * BB, h0 = th * BH, w0 = tw * BW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          uint8_t* sa = smem + stage * stage_bytes;
          tma_load_4d(sa, &tmDY, &full_bar[stage], m_tile * 128, w0, h0, b0);
          tma_load_4d(sa + kChunkBytes, &tmDY, &full_bar[stage], m_tile * 128 + 64, w0, h0, b0);
          for (int c = 0; c < n_chunks; ++c)
            tma_load_4d(sa + a_bytes + c * kChunkBytes, &tmX, &full_bar[stage], n_tile * block_n + c * 64, w0 + dw,
                        h0 + dh, b0 + db);
        }
        __syncwarp();
        stage += 3;
        while (stage >= nstages) {
          stage -= nstages;
          phase ^= 1;
        }
      }
      skip = kb - kb1;  // carry the round-robin position into the next item
    }
  } else if (warp == 0) {
    // MMA issuer: loop state hoisted into registers (see conv_fwd.cu)
    const int nstages = a.stages;
    // MN-major SW128: LBO = distance between 64-channel chunks, SBO = 8 pixel rows * 128 B
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem), kChunkBytes, 1024, kLayoutSW128);
    const uint32_t stage_d = static_cast<uint32_t>(stage_bytes) >> 4, b_off_d = static_cast<uint32_t>(a_bytes) >> 4;
    const uint32_t idesc = a.idesc, block_n = a.block_n;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int tap, m_tile, n_tile, split;
      decode(item, tap, m_tile, n_tile, split);
      const int kb0 = split * a.kb_per_split;
      const int kb1 = min(kb0 + a.kb_per_split, a.k_blocks);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * block_n;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * stage_d);
          const uint64_t bdesc = adesc + b_off_d;
          // 16 pixel rows per MMA = 2048 B = 128 in the (addr >> 4) field
          umma_bf16(d_tmem, adesc, bdesc, idesc, kb > kb0 ? 1u : 0u);
          umma_bf16(d_tmem, adesc + 128, bdesc + 128, idesc, 1u);
          umma_bf16(d_tmem, adesc + 256, bdesc + 256, idesc, 1u);
          umma_bf16(d_tmem, adesc + 384, bdesc + 384, idesc, 1u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int tap, m_tile, n_tile, split;
      decode(item, tap, m_tile, n_tile, split);
      const int co = m_tile * 128 + row;
      const bool valid = co < a.Cout;
      float* dst = a.partial + ((static_cast<size_t>(split) * a.Cout + co) * a.ntaps + tap) * a.Cin;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * a.block_n;
      const int nchunks = a.block_n >> 5;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_tile * a.block_n + c * 32;
        if (valid) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int n = n0 + g * 4;
            if (n < a.Cin)
              *reinterpret_cast<float4*>(dst + n) =
                  make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]), __uint_as_float(r[g * 4 + 2]),
                              __uint_as_float(r[g * 4 + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// dw[layout] (+)= sum_s partial[s][n][tap][c]   (fixed summation order: deterministic; 4 channels per thread)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int Cout,
                                    int ntaps, int Cin, int layout, int accumulate, int ldc, int c0) {
  pdl_entry();
  const size_t total4 = static_cast<size_t>(Cout) * ntaps * Cin / 4;
  for (size_t i4 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i4 < total4;
       i4 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    // eight independent loads in flight, then the adds in split order (1x1 layers have up to ~150 splits and only
    // a few thousand outputs: a dependent load-add chain would cost splits x DRAM latency)
    for (int k0 = 0; k0 < splits; k0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < splits) v[u] = __ldg(reinterpret_cast<const float4*>(partial) + (k0 + u) * total4 + i4);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < splits) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    const size_t i = i4 * 4;
    if (layout == 0) {
      float4* o = reinterpret_cast<float4*>(dw) + i4;
      if (accumulate) { const float4 d = *o; s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w; }
      *o = s;
    } else {
      const int c = static_cast<int>(i % Cin);
      const int tap = static_cast<int>((i / Cin) % ntaps);
      const int n = static_cast<int>(i / (static_cast<size_t>(Cin) * ntaps));
      float* o = dw + (static_cast<size_t>(n) * ldc + c0 + c) * ntaps + tap;
      const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j * ntaps] = accumulate ? o[j * ntaps] + v[j] : v[j];
    }
  }
}

struct WgradPlan {
  WgradKArgs a;
  size_t smem_bytes;
  int grid;
};

static int plan_wgrad(const pddm_wgrad_params* p, WgradPlan* plan) {
  if (!p) return PDDM_ERR_BAD_ARG;
  if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0 || p->ntaps <= 0 ||
      p->ntaps > PDDM_MAX_TAPS || p->x_NB < p->B)
    return PDDM_ERR_BAD_ARG;
  if (p->Cin % 8 != 0 || p->Cout % 8 != 0 || p->ldx % 8 != 0 || p->lddy % 8 != 0 || p->ldx < p->Cin ||
      p->lddy < p->Cout)
    return PDDM_ERR_UNSUPPORTED;
  WgradKArgs& a = plan->a;
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cin = p->Cin; a.Cout = p->Cout;
  a.BW = p->W < 64 ? p->W : 64;
  a.BH = 64 / a.BW < p->H ? 64 / a.BW : p->H;
  if (a.BH < 1) a.BH = 1;
  a.BB = 64 / (a.BW * a.BH) < p->B ? 64 / (a.BW * a.BH) : p->B;
  if (a.BB < 1) a.BB = 1;
  a.tiles_w = (p->W + a.BW - 1) / a.BW;
  a.tiles_h = (p->H + a.BH - 1) / a.BH;
  const int tiles_b = (p->B + a.BB - 1) / a.BB;
  a.k_blocks = a.tiles_w * a.tiles_h * tiles_b;
  const int cin64 = (p->Cin + 63) / 64 * 64;
  a.block_n = cin64 <= 256 ? cin64 : (cin64 % 256 == 0 ? 256 : (cin64 % 192 == 0 ? 192 : 128));
  a.n_chunks = a.block_n / 64;
  a.n_tiles = (p->Cin + a.block_n - 1) / a.block_n;
  a.m_tiles = (p->Cout + 127) / 128;
  a.ntaps = p->ntaps;
  for (int i = 0; i < PDDM_MAX_TAPS; ++i) {
    a.tap_db[i] = i < p->ntaps ? p->tap_db[i] : 0;
    a.tap_dh[i] = i < p->ntaps ? p->tap_dh[i] : 0;
    a.tap_dw[i] = i < p->ntaps ? p->tap_dw[i] : 0;
  }
  const int base_items = a.ntaps * a.m_tiles * a.n_tiles;
  const int sms = launch_sms();
  int splits = sms / base_items;  // one wave of work items: every extra split costs a full fp32 partial tile
  const int max_splits = (a.k_blocks + 7) / 8;           // but keep >= 8 K-blocks per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.k_blocks + splits - 1) / splits;
  a.splits = (a.k_blocks + a.kb_per_split - 1) / a.kb_per_split;
  a.b_bytes = a.n_chunks * kChunkBytes;
  const int box_bytes = a.BW * a.BH * a.BB * 128;
  a.tx_bytes = (2 + a.n_chunks) * box_bytes;
  a.idesc = make_idesc_bf16(128, a.block_n, 1, 1);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * a.block_n)) cols <<= 1;
  a.tmem_cols = cols;
  const int stage_bytes = 2 * kChunkBytes + a.b_bytes;
  const int max_smem = device_info().max_smem_optin > 0 ? device_info().max_smem_optin : 232448;
  int stages = (max_smem - 1024 - 512) / stage_bytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  if (stages < 2) return PDDM_ERR_UNSUPPORTED;
  a.stages = stages;
  plan->smem_bytes = static_cast<size_t>(stages) * stage_bytes + 1024 + 512;
  const int total = base_items * a.splits;
  plan->grid = total < sms ? total : sms;
  return PDDM_OK;
}

}  // namespace pddm

using namespace pddm;

extern "C" size_t pddm_conv2d_wgrad_workspace(const pddm_wgrad_params* p) {
  WgradPlan plan;
  if (plan_wgrad(p, &plan) != PDDM_OK) return 0;
  return static_cast<size_t>(plan.a.splits) * p->Cout * p->ntaps * p->Cin * sizeof(float);
}

extern "C" int pddm_conv2d_wgrad(const pddm_wgrad_params* p, void* workspace, size_t workspace_bytes,
                                 pddm_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->x || !p->dy || !p->dw || !workspace) return PDDM_ERR_BAD_ARG;
  if (!device_info().ok) return PDDM_ERR_ARCH;
  WgradPlan plan;
  int rc = plan_wgrad(p, &plan);
  if (rc) return rc;
  if (!aligned16(p->x) || !aligned16(p->dy) || !aligned16(workspace)) return PDDM_ERR_BAD_ARG;
  const size_t need = static_cast<size_t>(plan.a.splits) * p->Cout * p->ntaps * p->Cin * sizeof(float);
  if (workspace_bytes < need) return PDDM_ERR_WORKSPACE;
  plan.a.partial = static_cast<float*>(workspace);

  CUtensorMap tmDY, tmX;
  const uint32_t box[4] = {64u, static_cast<uint32_t>(plan.a.BW), static_cast<uint32_t>(plan.a.BH),
                           static_cast<uint32_t>(plan.a.BB)};
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(p->Cout), static_cast<uint64_t>(p->W), static_cast<uint64_t>(p->H),
                              static_cast<uint64_t>(p->B)};
    const uint64_t str[3] = {static_cast<uint64_t>(p->lddy) * 2, static_cast<uint64_t>(p->W) * p->lddy * 2,
                             static_cast<uint64_t>(p->H) * p->W * p->lddy * 2};
    rc = make_tmap_bf16(&tmDY, p->dy, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(p->Cin), static_cast<uint64_t>(p->W), static_cast<uint64_t>(p->H),
                              static_cast<uint64_t>(p->x_NB)};
    const uint64_t str[3] = {static_cast<uint64_t>(p->ldx) * 2, static_cast<uint64_t>(p->W) * p->ldx * 2,
                             static_cast<uint64_t>(p->H) * p->W * p->ldx * 2};
    rc = make_tmap_bf16(&tmX, p->x, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  if (ensure_smem_optin(reinterpret_cast<const void*>(conv_wgrad_kernel))) return PDDM_ERR_CUDA;
  PdlLaunch(plan.grid, kWgThreads, plan.smem_bytes, stream)(conv_wgrad_kernel, tmDY, tmX, plan.a);
  if (cudaPeekAtLastError() != cudaSuccess) return PDDM_ERR_CUDA;
  const size_t total = static_cast<size_t>(p->Cout) * p->ntaps * p->Cin;
  int blocks = static_cast<int>((total / 4 + 255) / 256);
  if (blocks > 2368) blocks = 2368;
  if (blocks < 1) blocks = 1;
  if (p->dw_ldc && (p->dw_layout != 1 || p->dw_c0 < 0 || p->dw_c0 + p->Cin > p->dw_ldc)) return PDDM_ERR_BAD_ARG;
  PdlLaunch(blocks, 256, 0, stream)(wgrad_reduce_kernel, plan.a.partial, p->dw, plan.a.splits, p->Cout, p->ntaps, p->Cin,
                                    p->dw_layout, p->accumulate, p->dw_ldc ? p->dw_ldc : p->Cin,
                                    p->dw_ldc ? p->dw_c0 : 0);
  return launch_status();
}
