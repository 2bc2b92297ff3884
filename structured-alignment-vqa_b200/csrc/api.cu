// Library plumbing: error reporting, device check, TMA descriptor construction (driver entry point
// resolved at run time so the .so links without libcuda and loads on a CPU-only box).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace savqa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  (void)cudaGetLastError();  // do not let a reported (non-sticky) error leak into the next call's cudaGetLastError()
  return SAVQA_ERR_CUDA;
}

int ensure_dynamic_smem(const void* kernel, size_t bytes, const char* what) {
  static std::mutex mu;
  static std::unordered_map<const void*, size_t> granted;
  std::lock_guard<std::mutex> lk(mu);
  auto it = granted.find(kernel);
  if (it != granted.end() && it->second >= bytes) return SAVQA_OK;
  cudaFuncAttributes fa;
  SAVQA_CHECK_CUDA(cudaFuncGetAttributes(&fa, kernel));
  const size_t limit = 227 * 1024;
  if (bytes + fa.sharedSizeBytes > limit) {
    set_error("%s needs %zu bytes of dynamic + %zu bytes of static shared memory; the sm_100 limit is %zu per CTA", what, bytes,
              static_cast<size_t>(fa.sharedSizeBytes), limit);
    return SAVQA_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024) SAVQA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  granted[kernel] = bytes;
  return SAVQA_OK;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SAVQA_PDL");
    on = (e && e[0] == '1') ? 1 : 0;  // measured on the training step: no gain over plain launches (two streams already overlap tails)
  }
  return on == 1;
}

static std::atomic<long long> g_launch_counts[LK_COUNT];
void count_launch(int kind) {
  if (kind >= 0 && kind < LK_COUNT) g_launch_counts[kind].fetch_add(1, std::memory_order_relaxed);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) cached = n;
  }
  return cached > 0 ? cached : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t w[12];
  bool operator==(const MapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.w) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, bool swizzle128) {
  return make_tensor_map(out, base, false, rank, dims, strides_bytes, box, swizzle128);
}

int make_tensor_map(CUtensorMap* out, const void* base, bool is_f32, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return SAVQA_ERR_CUDA;
  }
  if (rank < 2 || rank > 3) {
    set_error("tensor map rank %d unsupported", rank);
    return SAVQA_ERR_BAD_ARGUMENT;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return SAVQA_ERR_BAD_ARGUMENT;
  }
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uint64_t>(base);
  key.w[1] = static_cast<uint64_t>(rank) | (swizzle128 ? 256u : 0u) | (is_f32 ? 512u : 0u);
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[8 + i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) {
    key.w[5 + i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) {
      set_error("TMA global stride %llu bytes is not a multiple of 16", static_cast<unsigned long long>(strides_bytes[i]));
      return SAVQA_ERR_BAD_ARGUMENT;
    }
  }
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return SAVQA_OK;
    }
  }
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bdim[3];
  cuuint32_t estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  alignas(64) CUtensorMap m;
  CUresult r = fn(&m, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", static_cast<int>(r), rank,
              static_cast<unsigned long long>(dims[0]), static_cast<unsigned long long>(dims[1]),
              static_cast<unsigned long long>(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return SAVQA_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return SAVQA_OK;
}

}  // namespace savqa

extern "C" {

int savqa_abi_version(void) { return SAVQA_ABI_VERSION; }

const char* savqa_last_error(void) { return savqa::g_err; }

int savqa_launch_counts(int64_t* out, int n) {
  static const char* names = "gemm_pair,gemm_single,attn_fwd_tc,attn_bwd_tc_shared,attn_bwd_tc,attn_fwd_simt,attn_bwd_simt,attn_row1_fwd,"
                             "attn_row1_bwd,rowln_gemm,mil_nce";
  (void)names;
  for (int i = 0; i < n; ++i) out[i] = i < savqa::LK_COUNT ? static_cast<int64_t>(savqa::g_launch_counts[i].load(std::memory_order_relaxed)) : 0;
  return savqa::LK_COUNT;
}

int savqa_device_check(int* sm_count_out) {
  int dev = 0;
  SAVQA_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0, sms = 0;
  SAVQA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SAVQA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  SAVQA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count_out) *sm_count_out = sms;
  if (major != 10) {
    savqa::set_error("savqa_b200 needs an sm_100a device (B200); found compute capability %d.%d", major, minor);
    return SAVQA_ERR_UNSUPPORTED;
  }
  return SAVQA_OK;
}

}  // extern "C"
