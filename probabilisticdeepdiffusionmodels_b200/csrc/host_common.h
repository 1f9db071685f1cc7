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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

// bf16 tensor map, rank <= 4.  dims/box innermost-first; strides_bytes[i] = stride of dim i+1 (rank-1 entries).
// swizzle_bytes in {32, 64, 128}.  Returns PDDM_OK or PDDM_ERR_TMA.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pddm
