// Host-side helpers shared by the C-ABI entry points: status mapping, device query, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pddm.h"

namespace pddm {

inline int launch_status() { return cudaPeekAtLastError() == cudaSuccess ? PDDM_OK : PDDM_ERR_CUDA; }

struct DeviceInfo {
  int ok;  // 1 if sm_100
  int sm_count;
  int max_smem_optin;
};
const DeviceInfo& device_info();
// SMs the persistent kernels (one CTA per SM: tap-GEMMs, weight gradients, pipelined GroupNorm) size their grids for:
// the device's SM count minus the reserve set through pddm_set_sm_reserve (SMs left to a concurrently running
// collective, see include/pddm.h).  A launch-configuration knob: results never depend on it beyond summation order.
int launch_sms();

// Experiment knobs (environment variables), read ONCE per process -- the launch paths themselves are stateless
// and never call getenv.  -1 = not set.
struct EnvKnobs {
  int pdl;              // PDDM_PDL=1: programmatic dependent launch
  int conv_noswap;      // PDDM_CONV_NOSWAP: never use the swapped-operand conv kernel
  int conv_swap_force;  // PDDM_CONV_SWAP_FORCE: use it whenever the shape allows
  int conv_mt;          // PDDM_CONV_MT
  int conv_dbg;         // PDDM_CONV_DBG bit mask
  int conv_stages;      // PDDM_CONV_STAGES upper bound
  int gn_stream;        // PDDM_GN_STREAM: register-streaming GroupNorm kernels only
  int gn_nopipe;        // PDDM_GN_NOPIPE: skip the persistent bulk-tensor GroupNorm kernels
  int gn_s;             // PDDM_GN_S: upper bound on the GroupNorm cluster size
  int gn_dbg;           // PDDM_GN_DBG experiment bit mask (groupnorm_pipe.cu)
  int gn_ng;            // PDDM_GN_NG=1: one compute group instead of two in the persistent GroupNorm kernels
  int gn_cc;            // PDDM_GN_CC: force the channel-chunk width of the persistent GroupNorm kernels
  int attn_dbg;         // PDDM_ATTN_DBG
  int conv_min_bn;      // PDDM_CONV_MIN_BN: smallest channel tile the small-extent heuristic may pick (default 32)
};
const EnvKnobs& env_knobs();

// Opt `fn` in to the device's maximum dynamic shared memory.  The attribute is per (device, function); it is set on
// first use and remembered in a mutex-protected cache (like device_info), so launch paths stay re-entrant and work
// with several devices in one process.  Returns PDDM_OK or PDDM_ERR_CUDA.
int ensure_smem_optin(const void* fn);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

// bf16 tensor map, rank <= 4.  dims/box innermost-first; strides_bytes[i] = stride of dim i+1 (rank-1 entries).
// swizzle_bytes in {32, 64, 128}.  Returns PDDM_OK or PDDM_ERR_TMA.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Every kernel of this library starts with griddepcontrol.launch_dependents (the next kernel in the stream may
// be scheduled as soon as all of OUR CTAs have started) and executes griddepcontrol.wait before its first global
// memory access (blocks until every kernel it depends on has completed and flushed).  Launched with the
// programmatic-stream-serialization attribute, the launch latency and prologue (barrier init, TMEM allocation,
// tensor-map prefetch) of kernel N+1 overlap the tail of kernel N -- ~1200 launches per training step.
// Opt-in with PDDM_PDL=1: inside replayed CUDA graphs (the product path) it measured neutral, so the default
// launches without the attribute, which turns both instructions into no-ops.
bool pdl_enabled();

struct PdlLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x = 0) {
    cfg = cudaLaunchConfig_t();
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    int n = 0;
    if (cluster_x > 0) {
      attr[n].id = cudaLaunchAttributeClusterDimension;
      attr[n].val.clusterDim.x = cluster_x;
      attr[n].val.clusterDim.y = 1;
      attr[n].val.clusterDim.z = 1;
      ++n;
    }
    if (pdl_enabled()) {
      attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
  }
#ifdef __CUDACC__
  template <typename... KArgs, typename... Args>
  cudaError_t operator()(void (*kernel)(KArgs...), Args&&... args) {
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
  }
#endif
  cudaError_t launch_c(const void* fn, void** args) { return cudaLaunchKernelExC(&cfg, fn, args); }
};

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_launch_dependents();
  pdl_wait();
}
#endif

}  // namespace pddm
