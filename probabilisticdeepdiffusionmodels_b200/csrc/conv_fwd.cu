// Tap-GEMM convolution / linear forward (and data gradient) on tcgen05 tensor cores.
//
//   D[m, n] = sum_tap sum_c  X[pixel(m) + offset(tap), c] * Wp[n, tap, c]        (fp32 accumulate in TMEM)
//
// One persistent CTA per SM, warp-specialised:
//   warps 1-3   TMA producers (K-blocks dealt round-robin): per K-block one 4-D box load of the (shifted) NHWC
//               activation tile -- out-of-bounds coordinates are zero-filled by the TMA unit, which is the conv's
//               padding -- and one 2-D box load of the packed weight tile; both land in 128B/64B-swizzled K-major
//               shared memory.
//   warp 0      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, bf16 -> fp32),
//               accumulators double-buffered in tensor memory so the epilogue of tile i overlaps the main loop of
//               tile i+1.
//   warps 4..7  epilogue: tcgen05.ld 32 lanes x 32 columns, + bias + per-sample broadcast (timestep embedding)
//               + residual, convert, 128-bit stores.
// Pipelines: smem full/empty mbarrier ring (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue).
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace pddm {

struct ConvKArgs {
  const float* bias;
  const float* bcast;
  const void* residual;
  void* y;
  int ld_bcast, res_dtype, y_dtype;
  int B, H, W, Cout;
  int BW, BH, BB, tiles_w, tiles_h, m_tiles, n_tiles;
  int block_n, bk, kblocks_per_tap, ntaps;
  int tap_db[PDDM_MAX_TAPS], tap_dh[PDDM_MAX_TAPS], tap_dw[PDDM_MAX_TAPS], tap_w[PDDM_MAX_TAPS];
  int out_H, out_W, out_sh, out_sw, out_oh, out_ow;
  int stages, a_slot_bytes, a_tx_bytes, b_bytes;
  int kc_split;  // K-blocks [kc_split, kblocks_per_tap) of every tap are read from the second source tensor
  int mt;    // 128-row M sub-tiles per CTA tile sharing one weight tile (1 or 2): 8 UMMAs per barrier round trip
  int swap;  // 1: operand roles swapped (conv_fwd_swap_kernel): M = 128 output channels, N = 256 pixels
  int nacc;  // accumulator stages in TMEM (2 = epilogue overlaps the next tile's main loop)
  int dbg;   // PDDM_CONV_DBG experiment bits: 1 = empty epilogue, 2 = no A loads, 4 = no MMAs, 8 = no B loads
  uint32_t idesc, layout_type, sbo_bytes, tmem_cols;
  int epi_tma;             // 1: bulk-tensor epilogue (conv_fwd_kernel): output leaves and the residual arrives by TMA
  uint32_t off_out, off_res;  // byte offsets of the 2 x 16 KB output / residual staging buffers
};

constexpr int kConvThreads = 384;  // warp 0 MMA, warps 1-3 TMA producers, warps 4-11 epilogue (two per TMEM lane quadrant)
constexpr int kMaxStages = 8;
constexpr int kNumProducers = 3;
constexpr int kMaxAddRows = 8;                   // tile rows may span up to this many samples with a staged bias/bcast
constexpr int kAddBytes = kMaxAddRows * 256 * 4;  // smem for the staged bias + per-sample broadcast vector

__global__ void __launch_bounds__(kConvThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ CUtensorMap tmR, const __grid_constant__ ConvKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int mt = a.mt;
  const int stage_bytes = mt * a.a_slot_bytes + a.b_bytes;
  const int nstages = a.stages, nacc = a.nacc, dbg = a.dbg;
  const int m_tiles = (a.m_tiles + mt - 1) / mt;  // CTA tiles along M (mt sub-tiles each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_bar = bars;                      // [stages]
  uint64_t* empty_bar = bars + kMaxStages;        // [stages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;    // [2]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  uint64_t* res_full = bars + 2 * kMaxStages + 6;  // [2 groups][2] residual staging buffers (bulk-tensor epilogue)
  float* smem_add = reinterpret_cast<float*>(smem + nstages * stage_bytes + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = m_tiles * a.n_tiles;
  const int total_kb = a.ntaps * a.kblocks_per_tap;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], a.epi_tma ? 8 : 4);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&res_full[s], 1);
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

  if (warp >= 1 && warp <= 3) {
    // ---------------------------------------------------------------- TMA producers (3 single-lane issuers)
    // Issuing one K-block (barrier wait + expect_tx + a 4-D and a 2-D bulk-tensor copy) costs several hundred
    // cycles of issue latency on one thread, more than the 256-512 cycles of tensor work it feeds; K-blocks are
    // therefore dealt round-robin to three producer warps, each walking its own (tile, tap, channel-block,
    // stage, parity) counters.  The loops run warp-uniformly and only the issue is predicated on one elected
    // lane: inside a divergent `if (lane == 0)` region the compiler wraps every UTMALDG / SYNCS operand in a
    // uniform-register waterfall loop (~100-200 cycles per instruction).
    const int pid = warp - 1;
    const int kpt = a.kblocks_per_tap, bk = a.bk, block_n = a.block_n;
    const int tiles_w = a.tiles_w, tiles_h = a.tiles_h, BW = a.BW, BH = a.BH, BB = a.BB;
    const uint32_t tx_bytes = ((dbg & 2) ? 0 : mt * a.a_tx_bytes) + ((dbg & 8) ? 0 : a.b_bytes);
    const uint32_t a_slot = a.a_slot_bytes;
    int stage = pid % nstages;
    uint32_t phase = (pid / nstages) & 1;
    int kb = pid;  // K-block index inside the current tile
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = (tile % m_tiles) * mt, n_tile = tile / m_tiles;
      const int b0 = (m_tile / (tiles_w * tiles_h)) * BB;
      const int h0 = ((m_tile / tiles_w) % tiles_h) * BH;
      const int w0 = (m_tile % tiles_w) * BW;
      // second sub-tile (mt == 2); past the last tile its coordinates fall outside the tensor -> zero fill
      const int b1 = ((m_tile + 1) / (tiles_w * tiles_h)) * BB;
      const int h1 = (((m_tile + 1) / tiles_w) % tiles_h) * BH;
      const int w1 = ((m_tile + 1) % tiles_w) * BW;
      const int n0 = n_tile * block_n;
      int tap = kb / kpt, kc = kb - tap * kpt;
      while (kb < total_kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          if (!(dbg & 2)) {
            // two-source input (concat-free skip connection): the trailing K-blocks of a tap come from tmA2
            const bool second = kc >= a.kc_split;
            const CUtensorMap* ma = second ? &tmA2 : &tmA;
            const int ck = (second ? kc - a.kc_split : kc) * bk;
            tma_load_4d(sa, ma, &full_bar[stage], ck, w0 + a.tap_dw[tap], h0 + a.tap_dh[tap], b0 + a.tap_db[tap]);
            if (mt == 2)
              tma_load_4d(sa + a_slot, ma, &full_bar[stage], ck, w1 + a.tap_dw[tap], h1 + a.tap_dh[tap],
                          b1 + a.tap_db[tap]);
          }
          if (!(dbg & 8))
            tma_load_2d(sa + mt * a_slot, &tmB, &full_bar[stage], (a.tap_w[tap] * kpt + kc) * bk, n0);
        }
        __syncwarp();
        kb += kNumProducers;
        kc += kNumProducers;
        while (kc >= kpt) {
          kc -= kpt;
          ++tap;
        }
        stage += kNumProducers;
        while (stage >= nstages) {
          stage -= nstages;
          phase ^= 1;
        }
      }
      kb -= total_kb;  // carry the round-robin position into the next tile
    }
  } else if (warp == 0) {
    // ---------------------------------------------------------------- MMA issuer (warp-uniform loop, one lane issues)
    // Everything the loop needs is hoisted into registers: per K-block it is one barrier wait, one 64-bit add per
    // operand descriptor, the UMMAs and the commit.  (A single thread's dependent-instruction latency, not the
    // tensor pipe, bounds this loop when it is written carelessly: ~70 SASS instructions per K-block measured
    // 0.32 us against 0.135 us of tensor work at N=128.)
    const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, a.sbo_bytes, a.layout_type);
    const uint32_t stage_d = static_cast<uint32_t>(stage_bytes) >> 4;  // descriptor address field counts 16 B
    const uint32_t a_slot_d = static_cast<uint32_t>(a.a_slot_bytes) >> 4;
    const uint32_t b_off_d = mt * a_slot_d;
    const uint32_t idesc = a.idesc;
    const bool k64 = a.bk == 64;
    const bool no_mma = dbg & 4;
    const uint32_t block_n = a.block_n;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * mt * block_n;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = desc0 + static_cast<uint64_t>(stage * stage_d);
          const uint64_t bdesc = adesc + b_off_d;
          if (!no_mma) {
            // advance 16 elements (32 B) along K inside the swizzled row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc, bdesc, idesc, kb != 0);
            umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
            if (k64) {
              umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
              umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
            }
            if (mt == 2) {  // the second sub-tile reuses the weight tile already in shared memory
              const uint64_t adesc1 = adesc + a_slot_d;
              umma_bf16(d_tmem + block_n, adesc1, bdesc, idesc, kb != 0);
              umma_bf16(d_tmem + block_n, adesc1 + 2, bdesc + 2, idesc, 1);
              if (k64) {
                umma_bf16(d_tmem + block_n, adesc1 + 4, bdesc + 4, idesc, 1);
                umma_bf16(d_tmem + block_n, adesc1 + 6, bdesc + 6, idesc, 1);
              }
            }
          }
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
      if (++acc == nacc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int etid = threadIdx.x - 128;  // 0..255 (0..127 on the direct-store path, which uses warps 4-7 only)
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rows_per_b = a.BW * a.BH;
    const int nchunks = a.block_n >> 5;
    const bool has_add = (a.bias != nullptr) || (a.bcast != nullptr);
    const bool stage_add = has_add && a.BB <= kMaxAddRows;
    if (a.epi_tma) {
      // -------------------------------------------------------------- bulk-tensor epilogue (8 warps)
      // A single epilogue warp per scheduler drains a 128 x 256 accumulator in ~4 us (every dependent-instruction
      // stall is exposed), far longer than the 1.2 us main loop of a 1x1 convolution -- so TWO warps share each TMEM
      // lane quadrant and take alternate 32-channel slices.  Each thread converts its pixel row (TMEM lane) of a slice
      // and writes the 64 bytes into its group's double-buffered staging tile, laid out like a TMA operand (64 B rows,
      // 64 B swizzle); one elected thread per group hands the slice to the TMA unit (one bulk tensor store, rows
      // outside the tensor clipped) and the group goes on.  The residual arrives the same way, two slices ahead --
      // the first of a tile while its main loop is still running -- so the drain never waits on a global load, and
      // per-thread 16-byte stores to 128 different rows (one sector each, LSU-bound) are gone.  Requires 128-pixel
      // tiles, bf16 output, identity output map, Cout % 32 == 0 (host-checked).
      const int grp = (warp - 4) >> 2, eg = etid & 127;
      uint8_t* out_stage = smem + a.off_out + grp * (2 * 128 * 64);
      uint8_t* res_stage = smem + a.off_res + grp * (2 * 128 * 64);
      uint64_t* rfull = res_full + grp * 2;
      const bool res = a.residual != nullptr;
      const int sw = (row >> 1) & 3;  // 64-byte swizzle: 16-byte unit index ^ address bits [7, 9)
      uint32_t sl = 0;  // slices this group has drained: staging buffer = sl & 1, residual barrier parity = (sl >> 1) & 1
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile / m_tiles;
        const int m_tile = tile % m_tiles;
        const int tw = m_tile % a.tiles_w;
        const int th = (m_tile / a.tiles_w) % a.tiles_h;
        const int tb = m_tile / (a.tiles_w * a.tiles_h);
        const int w0 = tw * a.BW, h0 = th * a.BH, b0 = tb * a.BB;
        const int nbase = n_tile * a.block_n;
        int nsl = (a.Cout - nbase) >> 5;
        if (nsl > (a.block_n >> 5)) nsl = a.block_n >> 5;
        const int bb = row / rows_per_b;
        if (stage_add && !(dbg & 32)) {
          asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's readers are done
          const int total = a.BB * a.block_n;
          for (int i = etid; i < total; i += 256) {
            const int sb = i / a.block_n, n = nbase + (i - sb * a.block_n);
            float v = 0.f;
            if (n < a.Cout) {
              if (a.bias) v = __ldg(a.bias + n);
              const int gb = b0 + sb;
              if (a.bcast && gb < a.B) v += __ldg(a.bcast + static_cast<size_t>(gb) * a.ld_bcast + n);
            }
            smem_add[i] = v;
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const float* addrow = smem_add + (bb < a.BB ? bb : 0) * a.block_n;
        if (res && q == 0) {
          if (elect_one()) {
            for (int k = 0; k < 2; ++k) {
              const int s2 = grp + 2 * k;
              if (s2 < nsl) {
                const uint32_t buf = (sl + k) & 1;
                mbar_expect_tx(&rfull[buf], 128 * 64);
                tma_load_4d(res_stage + buf * (128 * 64), &tmR, &rfull[buf], nbase + 32 * s2, w0, h0, b0);
              }
            }
          }
          __syncwarp();
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * a.block_n;
        if (grp >= nsl) {  // nothing to drain for this group (32-channel tile): hand the accumulator back at once
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        for (int s2 = grp; s2 < nsl; s2 += 2, ++sl) {
          const uint32_t buf = sl & 1;
          uint32_t r0[32];
          tmem_ld32(taddr + s2 * 32, r0);
          if (res) mbar_wait(&rfull[buf], (sl >> 1) & 1);
          tmem_ld_wait();
          if (s2 + 2 >= nsl) {  // every TMEM read of this warp has completed: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          uint8_t* orow = out_stage + buf * (128 * 64) + row * 64;
          const uint8_t* rrow = res_stage + buf * (128 * 64) + row * 64;
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // 16-byte units of the row: channels [8u, 8u + 8) of the slice
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r0[u * 8 + j]);
            if (has_add) {
              if (stage_add) {
                const float4 a0 = *reinterpret_cast<const float4*>(addrow + s2 * 32 + u * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(addrow + s2 * 32 + u * 8 + 4);
                v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
              } else {
                const int n = nbase + s2 * 32 + u * 8;
                const int gb = b0 + bb < a.B ? b0 + bb : a.B - 1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (a.bias) v[j] += __ldg(a.bias + n + j);
                  if (a.bcast) v[j] += __ldg(a.bcast + static_cast<size_t>(gb) * a.ld_bcast + n + j);
                }
              }
            }
            const int su = (u ^ sw) << 4;
            if (res) {
              const uint4 rv = *reinterpret_cast<const uint4*>(rrow + su);
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                v[2 * j] += __low2float(h2);
                v[2 * j + 1] += __high2float(h2);
              }
            }
            *reinterpret_cast<uint4*>(orow + su) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
          fence_proxy_async();
          // The staging buffer the group's NEXT slice will fill was handed to the TMA unit one slice ago by the warp
          // whose turn it was: once that store has read it (it has had a whole slice of time), the group may overwrite
          // it after the barrier.  The issuing role rotates over the group's four warps and is taken by an elected
          // lane in warp-uniform code, so the ~300 cycles of TMA issue land on a different warp every slice.
          if (q == ((sl + 3) & 3)) {
            if (elect_one()) bulk_wait_group_read<0>();
            __syncwarp();
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
          if (q == (sl & 3)) {
            if (elect_one()) {
              if (!(dbg & 64)) tma_store_4d(&tmY, out_stage + buf * (128 * 64), nbase + 32 * s2, w0, h0, b0);
              bulk_commit_group();
              if (res && s2 + 4 < nsl) {  // this residual buffer has been read by the group: refill it two slices ahead
                mbar_expect_tx(&rfull[buf], 128 * 64);
                tma_load_4d(res_stage + buf * (128 * 64), &tmR, &rfull[buf], nbase + 32 * (s2 + 4), w0, h0, b0);
              }
            }
            __syncwarp();
          }
        }
        if (++acc == nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (elect_one()) bulk_wait_group<0>();  // every warp's issuing lane: its stores have completed
      __syncwarp();
    } else if (warp < 8)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
     const int n_tile = tile / m_tiles;
     for (int hf = 0; hf < mt; ++hf) {
      const int m_tile = (tile % m_tiles) * mt + hf;
      const int tw = m_tile % a.tiles_w;
      const int th = (m_tile / a.tiles_w) % a.tiles_h;
      const int tb = m_tile / (a.tiles_w * a.tiles_h);
      const int bb = row / rows_per_b;
      const int rr = row - bb * rows_per_b;
      const int hh = rr / a.BW;
      const int ww = rr - hh * a.BW;
      const int b = tb * a.BB + bb, h = th * a.BH + hh, w = tw * a.BW + ww;
      const bool valid = (bb < a.BB) && (b < a.B) && (h < a.H) && (w < a.W);
      const size_t opix = (static_cast<size_t>(b) * a.out_H + (h * a.out_sh + a.out_oh)) * a.out_W +
                          (w * a.out_sw + a.out_ow);
      const int nbase = n_tile * a.block_n;
      // Stage bias[n] + bcast[b, n] for this tile in shared memory while the MMA main loop is still running,
      // so the accumulator drain below never waits on a global load.
      if (stage_add) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers are done
        const int total = a.BB * a.block_n;
        for (int i = etid; i < total; i += 128) {
          const int sb = i / a.block_n, n = nbase + (i - sb * a.block_n);
          float v = 0.f;
          if (n < a.Cout) {
            if (a.bias) v = __ldg(a.bias + n);
            const int gb = tb * a.BB + sb;
            if (a.bcast && gb < a.B) v += __ldg(a.bcast + static_cast<size_t>(gb) * a.ld_bcast + n);
          }
          smem_add[i] = v;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const float* addrow = smem_add + (bb < a.BB ? bb : 0) * a.block_n;
      const bool res_bf16 = a.residual && a.res_dtype == PDDM_BF16;
      const __nv_bfloat16* res_b = reinterpret_cast<const __nv_bfloat16*>(a.residual) + opix * a.Cout;
      // prefetch the first residual chunk before blocking on the accumulator
      uint4 rres[2][4];
      if (res_bf16 && valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (nbase + g * 8 < a.Cout) rres[0][g] = __ldg(reinterpret_cast<const uint4*>(res_b + nbase + g * 8));
      }
      if (hf == 0) mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc * mt + hf) * a.block_n;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
      for (int cp = 0; cp < nchunks; cp += 2) {
#pragma unroll
       for (int ci = 0; ci < 2; ++ci) {
        const int c = cp + ci;
        if (c < nchunks) {
          tmem_ld_wait();
          if (c + 1 < nchunks) {
            tmem_ld32(taddr + (c + 1) * 32, r[(ci + 1) & 1]);  // overlaps with the math/stores of chunk c
            if (res_bf16 && valid) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (nbase + (c + 1) * 32 + g * 8 < a.Cout)
                  rres[(ci + 1) & 1][g] = __ldg(reinterpret_cast<const uint4*>(res_b + nbase + (c + 1) * 32 + g * 8));
            }
          }
          const int n0 = nbase + c * 32;
          if (valid && n0 < a.Cout && !(dbg & 1)) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[ci][j]);
            if (stage_add) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 av = *reinterpret_cast<const float4*>(addrow + c * 32 + j);
                v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
              }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = n0 + g * 8;
              if (n < a.Cout) {
                if (has_add && !stage_add) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (a.bias) v[g * 8 + j] += __ldg(a.bias + n + j);
                    if (a.bcast) v[g * 8 + j] += __ldg(a.bcast + static_cast<size_t>(b) * a.ld_bcast + n + j);
                  }
                }
                if (a.residual) {
                  if (res_bf16) {
                    const uint4 rv = rres[ci][g];
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                      v[g * 8 + 2 * j] += __low2float(h2);
                      v[g * 8 + 2 * j + 1] += __high2float(h2);
                    }
                  } else {
                    const float* rp = reinterpret_cast<const float*>(a.residual) + opix * a.Cout + n;
                    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
                    const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
                    v[g * 8 + 0] += r0.x; v[g * 8 + 1] += r0.y; v[g * 8 + 2] += r0.z; v[g * 8 + 3] += r0.w;
                    v[g * 8 + 4] += r1.x; v[g * 8 + 5] += r1.y; v[g * 8 + 6] += r1.z; v[g * 8 + 7] += r1.w;
                  }
                }
                if (a.y_dtype == PDDM_BF16) {
                  uint4 o;
                  o.x = pack_bf16(v[g * 8 + 0], v[g * 8 + 1]);
                  o.y = pack_bf16(v[g * 8 + 2], v[g * 8 + 3]);
                  o.z = pack_bf16(v[g * 8 + 4], v[g * 8 + 5]);
                  o.w = pack_bf16(v[g * 8 + 6], v[g * 8 + 7]);
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + opix * a.Cout + n) = o;
                } else {
                  float* yp = reinterpret_cast<float*>(a.y) + opix * a.Cout + n;
                  *reinterpret_cast<float4*>(yp) = make_float4(v[g * 8 + 0], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
                  *reinterpret_cast<float4*>(yp + 4) =
                      make_float4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
                }
              }
            }
          }
        }
       }
      }
     }
      // all of this warp's TMEM reads have completed (last tmem_ld_wait): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == nacc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ swapped operands
// Layers with <= 128 output channels (the 32x32 level of the UNet: 40 % of its conv FLOPs) would run 128x128x16
// UMMAs, which cost ~105 cycles each against ~144 for 128x256x16 (shared-memory operand reads dominate at N=128).
// Here the roles are swapped: the WEIGHT tile [128 channels x 64] is the M operand and a tile of 256 PIXELS is the N
// operand, so every UMMA is 128x256x16.  The accumulator then holds D^T (lane = output channel, column = pixel); the
// epilogue adds bias + per-sample broadcast per lane, transposes 32x32 blocks through a padded shared-memory stage
// and finishes pixel-major: residual add from coalesced 16-byte loads, bf16 pack, 16-byte stores.
// Requirements (checked by the host): Cout <= 128, bf16 output (and residual), identity output map, W a power of two,
// (pixels per sample in a tile) % 32 == 0.
constexpr int kSwapPitch = 144;                       // bytes per staged pixel row: 32 fp32 + 16 B pad
constexpr int kSwapEpiWarps = 8;                      // two warps per TMEM lane quadrant, four pixel chunks each
constexpr int kSwapThreads = 128 + 32 * kSwapEpiWarps;
constexpr int kSwapStageBytes = kSwapEpiWarps * 32 * kSwapPitch;  // one 32x32 fp32 block per epilogue warp

__global__ void __launch_bounds__(kSwapThreads, 1)
conv_fwd_swap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ ConvKArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kWBytes = 128 * 128;  // weight tile: 128 channels x 64 bf16
  const int stage_bytes = kWBytes + a.b_bytes;  // + 256 pixels x 64 bf16
  const int nstages = a.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  uint8_t* smem_stage = smem + nstages * stage_bytes + 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles;  // pixel tiles (one channel tile)
  const int total_kb = a.ntaps * a.kblocks_per_tap;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kSwapEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp >= 1 && warp <= 3) {
    const int pid = warp - 1;
    const int kpt = a.kblocks_per_tap, bk = a.bk;
    const uint32_t tx_bytes = kWBytes + a.a_tx_bytes;  // a_tx_bytes: bytes of the (clipped) pixel box
    int stage = pid % nstages;
    uint32_t phase = (pid / nstages) & 1;
    int kb = pid;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b0 = (tile / (a.tiles_w * a.tiles_h)) * a.BB;
      const int h0 = ((tile / a.tiles_w) % a.tiles_h) * a.BH;
      const int w0 = (tile % a.tiles_w) * a.BW;
      int tap = kb / kpt, kc = kb - tap * kpt;
      while (kb < total_kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sw = smem + stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_2d(sw, &tmB, &full_bar[stage], (a.tap_w[tap] * kpt + kc) * bk, 0);
          tma_load_4d(sw + kWBytes, &tmA, &full_bar[stage], kc * bk, w0 + a.tap_dw[tap], h0 + a.tap_dh[tap],
                      b0 + a.tap_db[tap]);
        }
        __syncwarp();
        kb += kNumProducers;
        kc += kNumProducers;
        while (kc >= kpt) {
          kc -= kpt;
          ++tap;
        }
        stage += kNumProducers;
        while (stage >= nstages) {
          stage -= nstages;
          phase ^= 1;
        }
      }
      kb -= total_kb;
    }
  } else if (warp == 0) {
    const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, a.sbo_bytes, a.layout_type);
    const uint32_t stage_d = static_cast<uint32_t>(stage_bytes) >> 4;
    const uint32_t p_off_d = static_cast<uint32_t>(kWBytes) >> 4;
    const uint32_t idesc = a.idesc;  // M = 128, N = 256
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t wdesc = desc0 + static_cast<uint64_t>(stage * stage_d);  // M operand: weights
          const uint64_t pdesc = wdesc + p_off_d;                                  // N operand: pixels
          umma_bf16(d_tmem, wdesc, pdesc, idesc, kb != 0);
          umma_bf16(d_tmem, wdesc + 2, pdesc + 2, idesc, 1);
          umma_bf16(d_tmem, wdesc + 4, pdesc + 4, idesc, 1);
          umma_bf16(d_tmem, wdesc + 6, pdesc + 6, idesc, 1);
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
  } else {
    const int q = warp & 3;
    const int co = q * 32 + lane;  // this thread's output channel (TMEM lane)
    const int eg = (warp - 4) >> 2;  // which half of the tile's 8 pixel chunks this warp drains
    uint8_t* stg = smem_stage + (warp - 4) * (32 * kSwapPitch);
    const int rpb = a.BW * a.BH;   // pixels of one sample inside a tile (multiple of 32)
    const int wshift = 31 - __clz(a.BW), wmask = a.BW - 1;
    const bool res = a.residual != nullptr;
    const float bias_v = (a.bias && co < a.Cout) ? __ldg(a.bias + co) : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tw = tile % a.tiles_w, th = (tile / a.tiles_w) % a.tiles_h, tb = tile / (a.tiles_w * a.tiles_h);
      const int c_lo = eg * (8 / (kSwapEpiWarps / 4)), c_hi = c_lo + 8 / (kSwapEpiWarps / 4);
      // pixel-major role of this lane: pixel i*8 + lane/4 of a 32-pixel chunk, 8 channels from n
      const int piece = lane & 3;
      const int n = q * 32 + piece * 8;
      // element offset of that pixel (chunk c, group i) in the NHWC output / residual, or -1 outside the tensor
      auto pix_off = [&](int c, int i) -> long long {
        const int p = c * 32 + i * 8 + (lane >> 2);
        const int bb = p / rpb, rr = p - bb * rpb;
        const int hh = rr >> wshift, ww = rr & wmask;
        const int b = tb * a.BB + bb, h = th * a.BH + hh, w = tw * a.BW + ww;
        const bool valid = (bb < a.BB) && (b < a.B) && (h < a.H) && (w < a.W) && (n < a.Cout);
        return valid ? ((static_cast<long long>(b) * a.H + h) * a.W + w) * a.Cout + n : -1ll;
      };
      // The residual of chunk c+1 is fetched while chunk c is transposed and stored, and the first chunk's while the
      // main loop of this tile is still running: the 16-byte loads never sit on the drain's critical path.
      uint4 rres[2][4];
      if (res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long off = pix_off(c_lo, i);
          if (off >= 0) rres[0][i] = __ldg(reinterpret_cast<const uint4*>(
                            reinterpret_cast<const __nv_bfloat16*>(a.residual) + off));
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      uint32_t r[2][32];
      tmem_ld32(taddr + c_lo * 32, r[0]);
#pragma unroll 1
      for (int cp = c_lo; cp < c_hi; cp += 2) {
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = cp + ci;
          tmem_ld_wait();
          if (c + 1 < c_hi) {
            tmem_ld32(taddr + (c + 1) * 32, r[ci ^ 1]);
            if (res) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const long long off = pix_off(c + 1, i);
                if (off >= 0) rres[ci ^ 1][i] = __ldg(reinterpret_cast<const uint4*>(
                                  reinterpret_cast<const __nv_bfloat16*>(a.residual) + off));
              }
            }
          }
          // ---- channel-major part: + bias + per-sample broadcast, park the 32 pixels of this channel
          const int bs = tb * a.BB + (c * 32) / rpb;  // the 32 pixels of a chunk belong to one sample
          float add = bias_v;
          if (a.bcast && co < a.Cout && bs < a.B) add += __ldg(a.bcast + static_cast<size_t>(bs) * a.ld_bcast + co);
          __syncwarp();  // the previous chunk's pixel-major readers are done with the stage
#pragma unroll
          for (int j = 0; j < 32; ++j)
            *reinterpret_cast<float*>(stg + j * kSwapPitch + lane * 4) = __uint_as_float(r[ci][j]) + add;
          __syncwarp();
          // ---- pixel-major part: residual, pack, store
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pl = i * 8 + (lane >> 2);
            const long long off = pix_off(c, i);
            const float4 v0 = *reinterpret_cast<const float4*>(stg + pl * kSwapPitch + piece * 32);
            const float4 v1 = *reinterpret_cast<const float4*>(stg + pl * kSwapPitch + piece * 32 + 16);
            if (off >= 0) {
              float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
              if (res) {
                const uint4 rv = rres[ci][i];
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[k]);
                  v[2 * k] += __low2float(h2);
                  v[2 * k + 1] += __high2float(h2);
                }
              }
              uint4 o;
              o.x = pack_bf16(v[0], v[1]);
              o.y = pack_bf16(v[2], v[3]);
              o.z = pack_bf16(v[4], v[5]);
              o.w = pack_bf16(v[6], v[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + off) = o;
            }
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
    tmem_dealloc(tmem_base, 512);
  }
}

static int pick_block_n(int cout) {
  const int c32 = (cout + 31) / 32 * 32;
  if (c32 <= 256) return c32;
  int best = 256, best_pad = 1 << 30;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int pad = (cout + cands[i] - 1) / cands[i] * cands[i];
    if (pad < best_pad) {
      best_pad = pad;
      best = cands[i];
    }
  }
  return best;
}

}  // namespace pddm

using namespace pddm;

extern "C" int pddm_conv2d_fwd(const pddm_conv_params* p, pddm_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->x || !p->w || !p->y) return PDDM_ERR_BAD_ARG;
  if (!device_info().ok) return PDDM_ERR_ARCH;
  if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->Cin <= 0 || p->Cout <= 0 || p->ntaps <= 0 ||
      p->ntaps > PDDM_MAX_TAPS || p->x_NB < p->B || p->w_ntaps < 1 || p->w_ntaps > PDDM_MAX_TAPS)
    return PDDM_ERR_BAD_ARG;
  if (p->Cin % 32 != 0 || p->Cout % 8 != 0 || p->ldx % 8 != 0) return PDDM_ERR_UNSUPPORTED;
  if (p->x2) {
    const int kb = (p->Cin % 64 == 0) ? 64 : 32;
    if (p->Cin_a <= 0 || p->Cin_a >= p->Cin || p->Cin_a % kb != 0 || p->ldx < p->Cin_a || p->ldx2 % 8 != 0 ||
        p->ldx2 < p->Cin - p->Cin_a || !aligned16(p->x2))
      return PDDM_ERR_UNSUPPORTED;
  } else if (p->ldx < p->Cin) {
    return PDDM_ERR_UNSUPPORTED;
  }
  if (!aligned16(p->x) || !aligned16(p->w) || !aligned16(p->y) || (p->residual && !aligned16(p->residual)) ||
      (p->bias && !aligned16(p->bias)) || (p->bcast && (!aligned16(p->bcast) || p->ld_bcast % 4 != 0)))
    return PDDM_ERR_BAD_ARG;

  ConvKArgs a;
  a.bias = p->bias;
  a.bcast = p->bcast;
  a.residual = p->residual;
  a.y = p->y;
  a.ld_bcast = p->ld_bcast;
  a.res_dtype = p->res_dtype;
  a.y_dtype = p->y_dtype;
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cout = p->Cout;
  a.swap = 0;
  {
    // swapped-operand path (conv_fwd_swap_kernel): <= 128 output channels on a large pixel count
    const int PW = p->W < 256 ? p->W : 256;
    const int PH = 256 / PW < p->H ? 256 / PW : p->H;
    const int PB = 256 / (PW * PH) < p->B ? 256 / (PW * PH) : p->B;
    const int tiles = ((p->W + PW - 1) / PW) * ((p->H + PH - 1) / PH) * ((p->B + PB - 1) / PB);
    const bool pow2w = (p->W & (p->W - 1)) == 0;
    if (!env_knobs().conv_noswap && !p->x2 && p->Cout <= 128 && p->Cin % 64 == 0 && p->y_dtype == PDDM_BF16 &&
        (!p->residual || p->res_dtype == PDDM_BF16) && p->out_sh == 1 && p->out_sw == 1 && p->out_oh == 0 &&
        p->out_ow == 0 && p->out_H == p->H && p->out_W == p->W && pow2w && PW * PH * PB == 256 &&
        (PW * PH) % 32 == 0 &&
        ((tiles >= 2 * launch_sms() && p->ntaps * (p->Cin / 64) >= 16) || env_knobs().conv_swap_force)) {
      a.swap = 1;
      a.BW = PW; a.BH = PH; a.BB = PB;
      a.tiles_w = (p->W + PW - 1) / PW;
      a.tiles_h = (p->H + PH - 1) / PH;
      a.m_tiles = tiles;
      a.n_tiles = 1;
      a.block_n = 256;
      a.bk = 64;
      a.kblocks_per_tap = p->Cin / 64;
      a.ntaps = p->ntaps;
      for (int i = 0; i < PDDM_MAX_TAPS; ++i) {
        a.tap_db[i] = i < p->ntaps ? p->tap_db[i] : 0;
        a.tap_dh[i] = i < p->ntaps ? p->tap_dh[i] : 0;
        a.tap_dw[i] = i < p->ntaps ? p->tap_dw[i] : 0;
        a.tap_w[i] = i < p->ntaps ? p->tap_w[i] : 0;
        if (i < p->ntaps && (p->tap_w[i] < 0 || p->tap_w[i] >= p->w_ntaps)) return PDDM_ERR_BAD_ARG;
      }
      a.out_H = p->out_H; a.out_W = p->out_W; a.out_sh = 1; a.out_sw = 1; a.out_oh = 0; a.out_ow = 0;
      a.layout_type = kLayoutSW128;
      a.sbo_bytes = 1024;
      a.a_slot_bytes = 128 * 128;
      a.a_tx_bytes = PW * PH * PB * 128;
      a.b_bytes = 256 * 128;
      a.idesc = make_idesc_bf16(128, 256, 0, 0);
      a.mt = 1; a.nacc = 2; a.dbg = 0; a.tmem_cols = 512;
      const int stage_bytes = 128 * 128 + a.b_bytes;
      int stages = (device_info().max_smem_optin - 1024 - 512 - kSwapStageBytes) / stage_bytes;
      if (stages > kMaxStages) stages = kMaxStages;
      if (stages < 3) return PDDM_ERR_UNSUPPORTED;
      a.stages = stages;
      const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + 1024 + 512 + kSwapStageBytes;
      CUtensorMap tmA, tmB;
      {
        const uint64_t dims[4] = {static_cast<uint64_t>(p->Cin), static_cast<uint64_t>(p->W),
                                  static_cast<uint64_t>(p->H), static_cast<uint64_t>(p->x_NB)};
        const uint64_t str[3] = {static_cast<uint64_t>(p->ldx) * 2, static_cast<uint64_t>(p->W) * p->ldx * 2,
                                 static_cast<uint64_t>(p->H) * p->W * p->ldx * 2};
        const uint32_t box[4] = {64, static_cast<uint32_t>(PW), static_cast<uint32_t>(PH), static_cast<uint32_t>(PB)};
        int rc = make_tmap_bf16(&tmA, p->x, 4, dims, str, box, 128);
        if (rc) return rc;
      }
      {
        const uint64_t ktot = static_cast<uint64_t>(p->w_ntaps) * p->Cin;
        const uint64_t dims[2] = {ktot, static_cast<uint64_t>(p->Cout)};
        const uint64_t str[1] = {ktot * 2};
        const uint32_t box[2] = {64, 128};
        int rc = make_tmap_bf16(&tmB, p->w, 2, dims, str, box, 128);
        if (rc) return rc;
      }
      if (ensure_smem_optin(reinterpret_cast<const void*>(conv_fwd_swap_kernel))) return PDDM_ERR_CUDA;
      const int grid = tiles < launch_sms() ? tiles : launch_sms();
      PdlLaunch(grid, kSwapThreads, smem_bytes, stream)(conv_fwd_swap_kernel, tmA, tmB, a);
      return launch_status();
    }
  }
  a.BW = p->W < 128 ? p->W : 128;
  a.BH = 128 / a.BW < p->H ? 128 / a.BW : p->H;
  if (a.BH < 1) a.BH = 1;
  a.BB = 128 / (a.BW * a.BH) < p->B ? 128 / (a.BW * a.BH) : p->B;
  if (a.BB < 1) a.BB = 1;
  a.tiles_w = (p->W + a.BW - 1) / a.BW;
  a.tiles_h = (p->H + a.BH - 1) / a.BH;
  const int tiles_b = (p->B + a.BB - 1) / a.BB;
  a.m_tiles = a.tiles_w * a.tiles_h * tiles_b;
  a.block_n = pick_block_n(p->Cout);
  // small spatial extents: trade B-operand reuse for enough tiles to occupy every SM
  const int min_bn = env_knobs().conv_min_bn > 0 ? env_knobs().conv_min_bn : 32;
  while (a.block_n % 64 == 0 && a.block_n / 2 >= min_bn &&
         a.m_tiles * ((p->Cout + a.block_n / 2 - 1) / (a.block_n / 2)) <= launch_sms())
    a.block_n /= 2;
  a.n_tiles = (p->Cout + a.block_n - 1) / a.block_n;
  a.bk = (p->Cin % 64 == 0) ? 64 : 32;
  a.kblocks_per_tap = p->Cin / a.bk;
  a.ntaps = p->ntaps;
  for (int i = 0; i < PDDM_MAX_TAPS; ++i) {
    a.tap_db[i] = i < p->ntaps ? p->tap_db[i] : 0;
    a.tap_dh[i] = i < p->ntaps ? p->tap_dh[i] : 0;
    a.tap_dw[i] = i < p->ntaps ? p->tap_dw[i] : 0;
    a.tap_w[i] = i < p->ntaps ? p->tap_w[i] : 0;
    if (i < p->ntaps && (p->tap_w[i] < 0 || p->tap_w[i] >= p->w_ntaps)) return PDDM_ERR_BAD_ARG;
  }
  a.out_H = p->out_H; a.out_W = p->out_W; a.out_sh = p->out_sh; a.out_sw = p->out_sw;
  a.out_oh = p->out_oh; a.out_ow = p->out_ow;
  const int swz = a.bk * 2;  // 128 or 64 bytes per K-major row
  a.layout_type = swz == 128 ? kLayoutSW128 : kLayoutSW64;
  a.sbo_bytes = 8 * swz;
  a.a_slot_bytes = 128 * swz;
  a.a_tx_bytes = a.BW * a.BH * a.BB * swz;
  a.b_bytes = a.block_n * swz;
  a.idesc = make_idesc_bf16(128, a.block_n, 0, 0);
  // Two 128-row sub-tiles can share one weight tile (8 UMMAs per barrier round trip, 25% fewer bytes into smem).
  // Measured on B200 it does not pay: a 128x128x16 UMMA costs ~105 cycles against ~144 for 128x256x16 whatever the
  // issue pattern, and the 256-row tiles lose more to wave quantisation than they gain.  Kept as an experiment knob.
  a.mt = 1;
  if (env_knobs().conv_mt == 2 && a.block_n <= 128) a.mt = 2;
  a.nacc = (2 * a.mt * a.block_n <= 512) ? 2 : 1;
  a.dbg = env_knobs().conv_dbg > 0 ? env_knobs().conv_dbg : 0;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(a.nacc * a.mt * a.block_n)) cols <<= 1;
  a.tmem_cols = cols;
  const int stage_bytes = a.mt * a.a_slot_bytes + a.b_bytes;
  // bulk-tensor epilogue (see the kernel): 2 x 16 KB output staging (+ 2 x 16 KB residual staging)
  a.epi_tma = (!(env_knobs().conv_dbg > 0 && (env_knobs().conv_dbg & 16)) && p->y_dtype == PDDM_BF16 &&
               p->out_sh == 1 && p->out_sw == 1 && p->out_oh == 0 && p->out_ow == 0 && p->out_H == p->H &&
               p->out_W == p->W && p->Cout % 32 == 0 && a.block_n % 32 == 0 && a.mt == 1 &&
               a.BW * a.BH * a.BB == 128 && (!p->residual || p->res_dtype == PDDM_BF16)) ? 1 : 0;
  const int epi_bytes = a.epi_tma ? (p->residual ? 8 : 4) * 128 * 64 + 1024 : 0;  // [group][2] x 8 KB (+ residual)
  const int budget = device_info().max_smem_optin - 1024 - 512 - kAddBytes - epi_bytes;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (env_knobs().conv_stages > 0 && env_knobs().conv_stages < stages) stages = env_knobs().conv_stages;
  if (stages < 2) return PDDM_ERR_UNSUPPORTED;
  a.stages = stages;
  a.off_out = (static_cast<uint32_t>(stages) * stage_bytes + 512 + kAddBytes + 1023) / 1024 * 1024;
  a.off_res = a.off_out + 4 * 128 * 64;
  const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + 1024 + 512 + kAddBytes + epi_bytes;

  CUtensorMap tmA, tmA2, tmB;
  const int cin_a = p->x2 ? p->Cin_a : p->Cin;
  a.kc_split = cin_a / a.bk;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(cin_a), static_cast<uint64_t>(p->W), static_cast<uint64_t>(p->H),
                              static_cast<uint64_t>(p->x_NB)};
    const uint64_t str[3] = {static_cast<uint64_t>(p->ldx) * 2, static_cast<uint64_t>(p->W) * p->ldx * 2,
                             static_cast<uint64_t>(p->H) * p->W * p->ldx * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(a.bk), static_cast<uint32_t>(a.BW), static_cast<uint32_t>(a.BH),
                             static_cast<uint32_t>(a.BB)};
    int rc = make_tmap_bf16(&tmA, p->x, 4, dims, str, box, swz);
    if (rc) return rc;
    tmA2 = tmA;
    if (p->x2) {
      const uint64_t dims2[4] = {static_cast<uint64_t>(p->Cin - cin_a), dims[1], dims[2], dims[3]};
      const uint64_t str2[3] = {static_cast<uint64_t>(p->ldx2) * 2, static_cast<uint64_t>(p->W) * p->ldx2 * 2,
                                static_cast<uint64_t>(p->H) * p->W * p->ldx2 * 2};
      rc = make_tmap_bf16(&tmA2, p->x2, 4, dims2, str2, box, swz);
      if (rc) return rc;
    }
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(p->w_ntaps) * p->Cin;
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(p->Cout)};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(a.bk), static_cast<uint32_t>(a.block_n)};
    int rc = make_tmap_bf16(&tmB, p->w, 2, dims, str, box, swz);
    if (rc) return rc;
  }
  CUtensorMap tmY = tmA, tmR = tmA;
  if (a.epi_tma) {
    const uint64_t dims[4] = {static_cast<uint64_t>(p->Cout), static_cast<uint64_t>(p->W), static_cast<uint64_t>(p->H),
                              static_cast<uint64_t>(p->B)};
    const uint64_t str[3] = {static_cast<uint64_t>(p->Cout) * 2, static_cast<uint64_t>(p->W) * p->Cout * 2,
                             static_cast<uint64_t>(p->H) * p->W * p->Cout * 2};
    const uint32_t box[4] = {32, static_cast<uint32_t>(a.BW), static_cast<uint32_t>(a.BH), static_cast<uint32_t>(a.BB)};
    int rc = make_tmap_bf16(&tmY, p->y, 4, dims, str, box, 64);
    if (rc) return rc;
    tmR = tmY;
    if (p->residual && (rc = make_tmap_bf16(&tmR, p->residual, 4, dims, str, box, 64))) return rc;
  }
  if (ensure_smem_optin(reinterpret_cast<const void*>(conv_fwd_kernel))) return PDDM_ERR_CUDA;
  const int total_tiles = ((a.m_tiles + a.mt - 1) / a.mt) * a.n_tiles;
  const int grid = total_tiles < launch_sms() ? total_tiles : launch_sms();
  PdlLaunch(grid, kConvThreads, smem_bytes, stream)(conv_fwd_kernel, tmA, tmA2, tmB, tmY, tmR, a);
  return launch_status();
}
