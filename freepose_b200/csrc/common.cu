// Host-side support: error channel, device query, TMA descriptor creation.
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

namespace fp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at: %s", int(e), cudaGetErrorString(e), what);
  return -2;
}

int sm_count() {
  static std::atomic<int> cache[64];   // per device ordinal (0 = not queried yet)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

static std::mutex g_smem_mu;
bool dyn_smem_needed(DynSmemOnce& once, int* device) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  *device = dev;
  if (dev < 0 || dev >= 64) return true;
  std::lock_guard<std::mutex> lk(g_smem_mu);
  return !((once.mask >> dev) & 1ull);
}
void dyn_smem_done(DynSmemOnce& once, int device) {
  if (device < 0 || device >= 64) return;
  std::lock_guard<std::mutex> lk(g_smem_mu);
  once.mask |= 1ull << device;
}

// ---------------------------------------------------------------------------------------------- profiler
namespace {
struct ProfRec { cudaEvent_t a, b; int kind; double work; int launches; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_event_pool;
std::atomic<long long> g_launches{0};
const char* kProfNames[PROF_NUM_KINDS] = {"gemm_qkv", "gemm_proj", "gemm_fc1_gelu", "gemm_fc2", "gemm_patch_embed",
                                          "attention", "layernorm", "token_prep", "score_topk", "raster", "geometry"};
cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(int kind, double work, int launches, cudaStream_t s) : idx(-1), stream(s) {
  g_launches += launches;
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{get_event(), get_event(), kind, work, launches};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  idx = int(g_recs.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_recs[idx].b, stream);
}

void prof_enable(int on) { std::lock_guard<std::mutex> lk(g_prof_mu); g_prof_on = on != 0; }
void prof_reset() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_recs) { g_event_pool.push_back(r.a); g_event_pool.push_back(r.b); }
  g_recs.clear();
}
const char* prof_name(int kind) { return (kind >= 0 && kind < PROF_NUM_KINDS) ? kProfNames[kind] : ""; }
long long launch_count() { return g_launches.load(); }
// Sums the event-timed durations of `kind` (synchronises on the recorded events).
int prof_collect(int kind, double* total_ms, double* total_work, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0, work = 0;
  long long n = 0;
  for (auto& r : g_recs) {
    if (r.kind != kind) continue;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize(profile)");
    float t = 0.f;
    e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime(profile)");
    ms += t; work += r.work; n += r.launches;
  }
  *total_ms = ms; *total_work = work; *launches = n;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  FP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  FP_REQUIRE((row_stride_elems * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  FP_REQUIRE(box_cols * 2 == 128, "swizzle-128B tiles are 64 bf16 wide");
  FP_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
  return 0;
}

int make_tmap_2d_bf16_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t row_stride_elems, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  FP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  FP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  FP_REQUIRE((row_stride_elems * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed with CUresult %d", int(r));
  return 0;
}

}  // namespace fp
