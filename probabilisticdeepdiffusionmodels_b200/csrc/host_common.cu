#include <stdlib.h>
#include "host_common.h"

#include <atomic>
#include <mutex>

namespace pddm {

static int env_int(const char* name) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : -1;
}
static int env_set(const char* name) { return getenv(name) ? 1 : 0; }

const EnvKnobs& env_knobs() {
  static EnvKnobs k;
  static std::once_flag once;
  std::call_once(once, [] {
    k.pdl = env_int("PDDM_PDL");
    k.conv_noswap = env_set("PDDM_CONV_NOSWAP");
    k.conv_swap_force = env_set("PDDM_CONV_SWAP_FORCE");
    k.conv_mt = env_int("PDDM_CONV_MT");
    k.conv_dbg = env_int("PDDM_CONV_DBG");
    k.conv_stages = env_int("PDDM_CONV_STAGES");
    k.gn_stream = env_set("PDDM_GN_STREAM");
    k.gn_nopipe = env_set("PDDM_GN_NOPIPE");
    k.gn_s = env_int("PDDM_GN_S");
    k.gn_dbg = env_int("PDDM_GN_DBG");
    k.gn_cc = env_int("PDDM_GN_CC");
    k.gn_ng = env_int("PDDM_GN_NG");
    k.attn_dbg = env_int("PDDM_ATTN_DBG");
    k.conv_min_bn = env_int("PDDM_CONV_MIN_BN");
  });
  return k;
}

// opt-in: measured neutral inside CUDA graphs (17.0 vs 16.8 ms/step)
bool pdl_enabled() { return env_knobs().pdl == 1; }


const DeviceInfo& device_info() {
  // Per-process, per-current-device cache (one process per GPU is the deployment model).
  static DeviceInfo info[64];
  static bool have[64] = {false};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    static DeviceInfo bad = {0, 0, 0};
    return bad;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    cudaDeviceProp prop;
    DeviceInfo d = {0, 0, 0};
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
      d.ok = (prop.major == 10);
      d.sm_count = prop.multiProcessorCount;
      d.max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    }
    info[dev] = d;
    have[dev] = true;
  }
  return info[dev];
}

static std::atomic<int> g_sm_reserve{0};
int launch_sms() {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int r = g_sm_reserve.load();
  return sms - r > 0 ? sms - r : sms;
}

int ensure_smem_optin(const void* fn) {
  static std::mutex mu;
  static const void* seen_fn[256];
  static int seen_dev[256];
  static int nseen = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PDDM_ERR_CUDA;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < nseen; ++i)
    if (seen_fn[i] == fn && seen_dev[i] == dev) return PDDM_OK;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, device_info().max_smem_optin) != cudaSuccess)
    return PDDM_ERR_CUDA;
  if (nseen < 256) {
    seen_fn[nseen] = fn;
    seen_dev[nseen] = dev;
    ++nseen;
  }
  return PDDM_OK;
}

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return PDDM_ERR_TMA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PDDM_OK : PDDM_ERR_TMA;
}

}  // namespace pddm

extern "C" {

int pddm_version(void) { return 100; }

const char* pddm_strerror(int status) {
  switch (status) {
    case PDDM_OK: return "ok";
    case PDDM_ERR_BAD_ARG: return "bad argument (null/misaligned pointer or non-positive size)";
    case PDDM_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernel";
    case PDDM_ERR_WORKSPACE: return "workspace too small";
    case PDDM_ERR_CUDA: return "CUDA launch error";
    case PDDM_ERR_ARCH: return "device is not sm_100 (B200); this library has no fallback path";
    case PDDM_ERR_TMA: return "cuTensorMapEncodeTiled failed";
    default: return "unknown pddm status";
  }
}

int pddm_check_device(void) { return pddm::device_info().ok ? PDDM_OK : PDDM_ERR_ARCH; }
int pddm_sm_count(void) { return pddm::device_info().sm_count; }
int pddm_set_sm_reserve(int32_t n) {
  const int sms = pddm::device_info().sm_count;
  if (n < 0 || (sms > 0 && n > sms / 2)) return PDDM_ERR_BAD_ARG;
  pddm::g_sm_reserve.store(n);
  return PDDM_OK;
}
int pddm_get_sm_reserve(void) { return pddm::g_sm_reserve.load(); }

}  // extern "C"
