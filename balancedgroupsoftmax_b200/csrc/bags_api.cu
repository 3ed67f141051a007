// C-ABI entry points of libbags_b200.so (see include/bags_b200.h).
// Host side only: argument validation, TMA descriptor encoding, kernel launches.
// No torch, no allocations, no device synchronisation.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/bags_b200.h"
#include "bags_gemm.cuh"
#include "bags_fused_fwd.cuh"
#include "bags_bwd_fused.cuh"
#include "bags_kernels.cuh"
#include "bags_allreduce.cuh"
#include "bags_nms.cuh"

using namespace bags;

// ----------------------------------------------------------------------------
// error handling
// ----------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define BAGS_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(BAGS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                  __FILE__, __LINE__);                                                    \
  } while (0)

#define BAGS_REQUIRE(cond, ...)                              \
  do {                                                       \
    if (!(cond)) return fail(BAGS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

extern "C" int bags_abi_version(void) { return BAGS_ABI_VERSION; }
extern "C" const char* bags_last_error(void) { return g_last_error.c_str(); }

// ----------------------------------------------------------------------------
// per-process device info (SM count, arch check) and driver entry point
// ----------------------------------------------------------------------------
struct DeviceInfo {
  int num_sms = 0;
  int cc_major = 0;
};
static std::mutex g_mutex;
static DeviceInfo g_dev[64];
static bool g_dev_ok[64] = {false};

// test hook: device buffer [ctas][8] int64 that the next launches stamp with %globaltimer values
static long long* g_timing = nullptr;
static int g_dbg = 0;
extern "C" int bags_debug_set_timing(void* dev_ptr) {
  g_timing = reinterpret_cast<long long*>(dev_ptr);
  g_dbg = 0;
  if (const char* v = getenv("BAGS_DBG")) g_dbg = atoi(v);   // test hook, not on the per-step path
  return BAGS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

// Library-owned counter pairs for the merged backward kernel's in-kernel grid synchronisation: one 256-byte slot
// per (device, stream) that has used it, zeroed once here and left zero by every launch.
static unsigned int* g_sync_pool[64] = {nullptr};
static cudaStream_t g_sync_streams[64][64];
static int g_sync_count[64] = {0};
static unsigned int* sync_slot(int dev, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_mutex);
  if (g_sync_pool[dev] == nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      (void)cudaGetLastError();
      return nullptr;   // cannot allocate while capturing: the caller falls back to the separate preparation kernel
    }
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, 64 * 256) != cudaSuccess || cudaMemset(ptr, 0, 64 * 256) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    g_sync_pool[dev] = static_cast<unsigned int*>(ptr);
  }
  for (int i = 0; i < g_sync_count[dev]; ++i)
    if (g_sync_streams[dev][i] == stream) return g_sync_pool[dev] + i * 64;
  if (g_sync_count[dev] >= 64) return nullptr;
  const int i = g_sync_count[dev]++;
  g_sync_streams[dev][i] = stream;
  return g_sync_pool[dev] + i * 64;
}

static int device_info(DeviceInfo& out) {
  int dev = 0;
  BAGS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(BAGS_ERR_INVALID, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_mutex);
  if (!g_dev_ok[dev]) {
    DeviceInfo d;
    BAGS_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    BAGS_CUDA(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    g_dev[dev] = d;
    g_dev_ok[dev] = true;
  }
  out = g_dev[dev];
  if (out.cc_major != 10)
    return fail(BAGS_ERR_ARCH, "libbags_b200 requires an sm_100 (B200) device; found compute capability %d.x",
                out.cc_major);
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    BAGS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return fail(BAGS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return BAGS_OK;
}

// 2-D row-major tensor [outer, inner] (inner contiguous), 128B-swizzled boxes.
static int make_tmap(CUtensorMap* tm, const void* ptr, int dtype, long long inner, long long outer,
                     long long ld_elems, int box_inner, int box_outer, bool atom32 = false, bool swz64 = false) {
  const int elt = (dtype == BAGS_DTYPE_BF16) ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return fail(BAGS_ERR_INVALID, "operand pointer %p is not 16-byte aligned", ptr);
  if (((ld_elems * elt) & 15) != 0)
    return fail(BAGS_ERR_INVALID, "operand row stride %lld bytes is not a multiple of 16", ld_elems * elt);
  if (box_inner * elt != (swz64 ? 64 : 128)) return fail(BAGS_ERR_INVALID, "internal: box inner must span the swizzle width");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems * elt)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, dtype == BAGS_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swz64 ? CU_TENSOR_MAP_SWIZZLE_64B : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(BAGS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%lld outer=%lld ld=%lld box=%dx%d)",
                (int)r, inner, outer, ld_elems, box_inner, box_outer);
  return BAGS_OK;
}

static int make_group_table(GroupTable& gt, const int32_t* slices_host, int G, int C) {
  BAGS_REQUIRE(slices_host != nullptr, "slices_host is NULL");
  BAGS_REQUIRE(G >= 1 && G <= kMaxG, "G=%d outside [1,%d]", G, kMaxG);
  gt.G = G;
  int prev_end = 0;
  for (int g = 0; g < kMaxG; ++g) {
    gt.start[g] = 0;
    gt.len[g] = 0;
    if (g < G) {
      const int s = slices_host[2 * g], l = slices_host[2 * g + 1];
      BAGS_REQUIRE(s >= prev_end && l >= 1 && s + l <= C,
                   "pred_slice row %d = (%d,%d) is not an ascending, in-range slice of %d logits", g, s, l, C);
      gt.start[g] = s;
      gt.len[g] = l;
      prev_end = s + l;
    }
  }
  return BAGS_OK;
}

// Tuning / experiment switches come from the environment.  They are read ONCE per name (no getenv on the per-call path);
// bags_reload_env() drops the cache (tests flip switches inside one process).
struct EnvEntry { const char* name; bool set; int value; };
static EnvEntry g_env[96];
static int g_env_n = 0;
static std::mutex g_env_mutex;
static int env_int(const char* name, int dflt) {
  std::lock_guard<std::mutex> lk(g_env_mutex);
  for (int i = 0; i < g_env_n; ++i)
    if (g_env[i].name == name || strcmp(g_env[i].name, name) == 0) return g_env[i].set ? g_env[i].value : dflt;
  const char* v = getenv(name);
  EnvEntry e{name, v && *v, (v && *v) ? atoi(v) : 0};
  if (g_env_n < 96) g_env[g_env_n++] = e;
  return e.set ? e.value : dflt;
}
extern "C" int bags_reload_env(void) {
  std::lock_guard<std::mutex> lk(g_env_mutex);
  g_env_n = 0;
  return BAGS_OK;
}

// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may begin while its predecessor
// in the stream is still running; every kernel of this library guards its first access to predecessor-produced
// data (and its first write) with griddepcontrol.wait, so stream semantics are preserved.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("BAGS_PDL", 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ----------------------------------------------------------------------------
// GEMM launcher
// ----------------------------------------------------------------------------
struct GemmArgs {
  const void* a; long long lda; bool a_mn;
  const void* b; long long ldb; bool b_mn;
  int M, N, K;
  int dtype;
  int splits;
  GemmParams p;  // epilogue fields pre-filled (out, ldo, bias, gscale, ...)
};

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TF32, int STAGES>
static int launch_gemm(const GemmArgs& ga, const DeviceInfo& di, cudaStream_t stream, bool pdl = false) {
  using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN, EPI, TF32, STAGES>;
  auto kernel = bags_gemm_kernel<BLOCK_N, A_MN, B_MN, EPI, TF32, STAGES>;
  CUtensorMap ta, tb;
  int rc;
  // A
  if (A_MN) rc = make_tmap(&ta, ga.a, ga.dtype, ga.M, ga.K, ga.lda, Cfg::SLAB, Cfg::BLOCK_K, TF32);
  else      rc = make_tmap(&ta, ga.a, ga.dtype, ga.K, ga.M, ga.lda, Cfg::BLOCK_K, Cfg::BLOCK_M);
  if (rc) return rc;
  if (B_MN) rc = make_tmap(&tb, ga.b, ga.dtype, ga.N, ga.K, ga.ldb, Cfg::SLAB, Cfg::BLOCK_K, TF32);
  else      rc = make_tmap(&tb, ga.b, ga.dtype, ga.K, ga.N, ga.ldb, Cfg::BLOCK_K, Cfg::UMMA_N);
  if (rc) return rc;

  GemmParams p = ga.p;
  p.timing = g_timing;
  p.dbg = g_timing ? g_dbg : 0;
  p.M = ga.M; p.N = ga.N; p.K = ga.K;
  p.num_m_tiles = (ga.M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  p.num_n_tiles = (ga.N + BLOCK_N - 1) / BLOCK_N;
  p.kblocks_total = (ga.K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;
  int splits = ga.splits < 1 ? 1 : ga.splits;
  if (splits > p.kblocks_total) splits = p.kblocks_total;
  if (EPI != EPI_RED_F32) splits = 1;
  p.num_splits = splits;
  const int units = p.num_m_tiles * p.num_n_tiles * splits;
  if (units == 0) return BAGS_OK;
  const int grid = units < di.num_sms ? units : di.num_sms;

  BAGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && env_int("BAGS_PDL", 1)) ? 1 : 0;
  BAGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, p));
  return BAGS_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int EPI, bool TF32, int STAGES>
static int launch_gemm_pdl(const GemmArgs& ga, const DeviceInfo& di, cudaStream_t stream) {
  return launch_gemm<BLOCK_N, A_MN, B_MN, EPI, TF32, STAGES>(ga, di, stream, true);
}

static int pick_splits(int tiles, int kblocks, int num_sms) {
  int s = num_sms / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
}

// forward tile width: fewest (waves x tile width)
static int pick_fwd_block_n(int N, int C, int num_sms) {
  const int forced = env_int("BAGS_FWD_BN", 0);
  if (forced == 256 || forced == 320) return forced;
  const int mt = (N + 127) / 128;
  auto cost = [&](int bn) {
    const int tiles = mt * ((C + bn - 1) / bn);
    return ((tiles + num_sms - 1) / num_sms) * bn;
  };
  return cost(320) < cost(256) ? 320 : 256;
}

// ----------------------------------------------------------------------------
// public entry points
// ----------------------------------------------------------------------------
extern "C" size_t bags_workspace_bytes(void) { return 256 + 4096 * kMaxG * sizeof(float); }

extern "C" int bags_linear_fwd(const void* x, long long ldx, const void* w, long long ldw,
                               const float* bias, float* out, long long ldo, int N, int K, int C,
                               int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(N >= 0 && K >= 1 && C >= 1, "bags_linear_fwd: bad shape N=%d K=%d C=%d", N, K, C);
  BAGS_REQUIRE(N == 0 || (x && w && out), "bags_linear_fwd: NULL operand");
  BAGS_REQUIRE(dtype == BAGS_DTYPE_F32 || dtype == BAGS_DTYPE_BF16, "bags_linear_fwd: bad dtype %d", dtype);
  BAGS_REQUIRE(ldx >= K && ldw >= K && ldo >= C, "bags_linear_fwd: leading dimension smaller than row");
  if (N == 0) return BAGS_OK;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (bias != nullptr) BAGS_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bias must be 16-byte aligned");
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
  GemmArgs ga{};
  ga.a = x; ga.lda = ldx; ga.a_mn = false;
  ga.b = w; ga.ldb = ldw; ga.b_mn = false;
  ga.M = N; ga.N = C; ga.K = K; ga.dtype = dtype; ga.splits = 1;
  ga.p.out = out; ga.p.ldo = ldo; ga.p.bias = bias; ga.p.gscale = nullptr; ga.p.G = 0;
  ga.p.colsum_in = nullptr; ga.p.colsum_out = nullptr;
  const int bn = pick_fwd_block_n(N, C, di.num_sms);
  if (dtype == BAGS_DTYPE_BF16) {
    return bn == 320 ? launch_gemm<320, false, false, EPI_STORE_F32, false, 3>(ga, di, stream)
                     : launch_gemm<256, false, false, EPI_STORE_F32, false, 4>(ga, di, stream);
  }
  return bn == 320 ? launch_gemm<320, false, false, EPI_STORE_F32, true, 3>(ga, di, stream)
                   : launch_gemm<256, false, false, EPI_STORE_F32, true, 4>(ga, di, stream);
}

// y = act(x W^T + b): the head's shared FCs (ReLU) and fc_reg (identity) on the same tcgen05 pipeline
// (convfc_bbox_head.py:138-143,167: nn.Linear + ReLU through cuBLAS / ATen in the reference)
extern "C" int bags_linear_act_fwd(const void* x, long long ldx, const void* w, long long ldw, const float* bias,
                                   void* out, long long ldo, int N, int K, int C, int dtype, int out_dtype, int relu,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(N >= 0 && K >= 1 && C >= 1, "bags_linear_act_fwd: bad shape N=%d K=%d C=%d", N, K, C);
  BAGS_REQUIRE(N == 0 || (x && w && out), "bags_linear_act_fwd: NULL operand");
  BAGS_REQUIRE(dtype == BAGS_DTYPE_F32 || dtype == BAGS_DTYPE_BF16, "bags_linear_act_fwd: bad dtype %d", dtype);
  BAGS_REQUIRE(out_dtype == BAGS_DTYPE_F32 || (out_dtype == BAGS_DTYPE_BF16 && dtype == BAGS_DTYPE_BF16),
               "bags_linear_act_fwd: output dtype %d not available for operand dtype %d", out_dtype, dtype);
  BAGS_REQUIRE(ldx >= K && ldw >= K && ldo >= C, "bags_linear_act_fwd: leading dimension smaller than row");
  if (N == 0) return BAGS_OK;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (bias != nullptr && out_dtype == BAGS_DTYPE_F32)
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bags_linear_act_fwd: bias must be 16-byte aligned");
  GemmArgs ga{};
  ga.a = x; ga.lda = ldx; ga.a_mn = false;
  ga.b = w; ga.ldb = ldw; ga.b_mn = false;
  ga.M = N; ga.N = C; ga.K = K; ga.dtype = dtype; ga.splits = 1;
  ga.p.out = out; ga.p.ldo = ldo; ga.p.bias = bias; ga.p.relu = relu ? 1 : 0;
  if (dtype == BAGS_DTYPE_BF16)
    return out_dtype == BAGS_DTYPE_BF16 ? launch_gemm<256, false, false, EPI_STORE_BF16, false, 4>(ga, di, stream)
                                        : launch_gemm<256, false, false, EPI_STORE_F32, false, 4>(ga, di, stream);
  return launch_gemm<256, false, false, EPI_STORE_F32, true, 4>(ga, di, stream);
}

// Split-K factor worth using for a forward layer (1 = the single-pass kernel): few output tiles and a long contraction --
// shared_fcs.0 at 2 x 512 RoIs is 8 x 4 tiles of 128 x 256 with K = 12544, i.e. 32 busy SMs out of 148 without it.
extern "C" int bags_linear_act_splits(int N, int K, int C, int dtype) {
  DeviceInfo di;
  if (device_info(di)) return 1;
  const int tiles = ((N + 127) / 128) * ((C + 255) / 256);
  const int kblocks = (K + (dtype == BAGS_DTYPE_BF16 ? 63 : 31)) / (dtype == BAGS_DTYPE_BF16 ? 64 : 32);
  if (tiles <= 0 || tiles * 2 > di.num_sms || kblocks < 16) return 1;
  int s = di.num_sms / tiles;
  if (s > kblocks / 4) s = kblocks / 4;
  return s < 2 ? 1 : (s > 16 ? 16 : s);
}

// Two-pass forward for such layers: split-K GEMM with red.add into a zeroed fp32 workspace [N, ldws], then
// out = act(ws + bias) in the output dtype.
extern "C" int bags_linear_act_fwd_splitk(const void* x, long long ldx, const void* w, long long ldw, const float* bias,
                                          void* out, long long ldo, int N, int K, int C, int dtype, int out_dtype, int relu,
                                          float* ws, long long ldws, int splits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(N >= 0 && K >= 1 && C >= 1, "bags_linear_act_fwd_splitk: bad shape N=%d K=%d C=%d", N, K, C);
  BAGS_REQUIRE(N == 0 || (x && w && out && ws), "bags_linear_act_fwd_splitk: NULL operand");
  BAGS_REQUIRE(dtype == BAGS_DTYPE_F32 || dtype == BAGS_DTYPE_BF16, "bags_linear_act_fwd_splitk: bad dtype %d", dtype);
  BAGS_REQUIRE(out_dtype == BAGS_DTYPE_F32 || out_dtype == BAGS_DTYPE_BF16, "bags_linear_act_fwd_splitk: bad output dtype");
  BAGS_REQUIRE(ldx >= K && ldw >= K && ldo >= C && ldws >= C && (ldws % 4) == 0 && splits >= 2,
               "bags_linear_act_fwd_splitk: bad leading dimension / split count");
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "bags_linear_act_fwd_splitk: workspace must be 16-byte aligned");
  if (N == 0) return BAGS_OK;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  BAGS_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * static_cast<size_t>(N) * static_cast<size_t>(ldws), stream));
  GemmArgs ga{};
  ga.a = x; ga.lda = ldx; ga.a_mn = false;
  ga.b = w; ga.ldb = ldw; ga.b_mn = false;
  ga.M = N; ga.N = C; ga.K = K; ga.dtype = dtype; ga.splits = splits;
  ga.p.out = ws; ga.p.ldo = ldws;
  int rc = (dtype == BAGS_DTYPE_BF16) ? launch_gemm<256, false, false, EPI_RED_F32, false, 4>(ga, di, stream)
                                      : launch_gemm<256, false, false, EPI_RED_F32, true, 4>(ga, di, stream);
  if (rc) return rc;
  const long long quads = static_cast<long long>(N) * ((C + 3) / 4);
  long long grid = (quads + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (out_dtype == BAGS_DTYPE_BF16)
    bias_act_store_kernel<true><<<static_cast<unsigned>(grid), 256, 0, stream>>>(ws, ldws, bias, out, ldo, N, C, relu ? 1 : 0);
  else
    bias_act_store_kernel<false><<<static_cast<unsigned>(grid), 256, 0, stream>>>(ws, ldws, bias, out, ldo, N, C, relu ? 1 : 0);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

// g = (y > 0 ? dy : 0) in the operand dtype of the backward contractions (y == NULL: g = dy, i.e. a cast)
extern "C" int bags_act_bwd(const void* dy, long long lddy, int dy_dtype, const void* y, long long ldy, int y_dtype,
                            void* g, long long ldg, int g_dtype, int rows, int cols, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(rows >= 0 && cols >= 0 && (cols % 4) == 0, "bags_act_bwd: cols must be a multiple of 4");
  if (rows == 0 || cols == 0) return BAGS_OK;
  BAGS_REQUIRE(dy && g, "bags_act_bwd: NULL argument");
  BAGS_REQUIRE(lddy >= cols && ldg >= cols && (ldg % 4) == 0 && (y == nullptr || ldy >= cols), "bags_act_bwd: bad leading dimension");
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "bags_act_bwd: g must be 16-byte aligned");
  const bool f_dy = dy_dtype == BAGS_DTYPE_F32, f_y = y_dtype == BAGS_DTYPE_F32, f_g = g_dtype == BAGS_DTYPE_F32;
  const long long quads = static_cast<long long>(rows) * (cols / 4);
  long long grid = (quads + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  const dim3 gr(static_cast<unsigned>(grid)), bl(256);
  typedef __nv_bfloat16 bf;
#define BAGS_ACT(TDY, TY, OB) act_bwd_kernel<TDY, TY, OB><<<gr, bl, 0, stream>>>(                                   \
      static_cast<const TDY*>(dy), lddy, static_cast<const TY*>(y), ldy, g, ldg, rows, cols)
  if (f_dy && f_y && f_g) BAGS_ACT(float, float, false);
  else if (f_dy && f_y) BAGS_ACT(float, float, true);
  else if (f_dy && f_g) BAGS_ACT(float, bf, false);
  else if (f_dy) BAGS_ACT(float, bf, true);
  else if (f_y && f_g) BAGS_ACT(bf, float, false);
  else if (f_y) BAGS_ACT(bf, float, true);
  else if (f_g) BAGS_ACT(bf, bf, false);
  else BAGS_ACT(bf, bf, true);
#undef BAGS_ACT
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

extern "C" int bags_sample_others_step(const int64_t* labels, const int32_t* label2bin, int N, int G,
                                       int classes, double ratio, uint64_t seed, const uint64_t* seed_step,
                                       uint8_t* wmask, float* avg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(labels && label2bin && wmask && avg, "bags_sample_others: NULL argument");
  BAGS_REQUIRE(G >= 1 && G <= kMaxG && classes >= 1 && N >= 0, "bags_sample_others: bad shape");
  BAGS_REQUIRE(ratio >= 0.0, "bags_sample_others: negative ratio");
  const long long* lab = reinterpret_cast<const long long*>(labels);
  const unsigned long long sd = static_cast<unsigned long long>(seed);
  const unsigned long long* st = reinterpret_cast<const unsigned long long*>(seed_step);
  if (N <= 4096)       BAGS_CUDA(launch_pdl(sample_others_kernel<4>, dim3(G), dim3(1024), 0, stream, lab, label2bin, classes, G, N, ratio, sd, wmask, avg, st));
  else if (N <= 16384) BAGS_CUDA(launch_pdl(sample_others_kernel<16>, dim3(G), dim3(1024), 0, stream, lab, label2bin, classes, G, N, ratio, sd, wmask, avg, st));
  else                 BAGS_CUDA(launch_pdl(sample_others_kernel<0>, dim3(G), dim3(1024), 0, stream, lab, label2bin, classes, G, N, ratio, sd, wmask, avg, st));
  return BAGS_OK;
}

extern "C" int bags_sample_others(const int64_t* labels, const int32_t* label2bin, int N, int G,
                                  int classes, double ratio, uint64_t seed, uint8_t* wmask,
                                  float* avg, void* stream_) {
  return bags_sample_others_step(labels, label2bin, N, G, classes, ratio, seed, nullptr, wmask, avg, stream_);
}

extern "C" int bags_mask_avg(const uint8_t* wmask, int N, int G, float* avg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(wmask && avg, "bags_mask_avg: NULL argument");
  BAGS_REQUIRE(G >= 1 && G <= kMaxG && N >= 0, "bags_mask_avg: bad shape");
  mask_avg_kernel<<<G, 256, 0, stream>>>(wmask, N, avg);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

template <int NV>
static int launch_group_ce(const float* logits, long long ldz, const int64_t* labels,
                           const int32_t* label2bin, const GroupTable& gt, const uint8_t* wmask,
                           const float* avg, int N, int C, int classes, float* loss, float* lse,
                           void* dz, long long ldd, int dz_dtype, float* colsum, void* workspace,
                           int num_sms, cudaStream_t stream, bool wf = false) {
  const int smem = 8 * NV * 128 * (int)sizeof(float);
  // persistent CTAs (2 per SM: 126 regs x 256 threads): every CTA walks several row octets so the
  // bias-gradient column sums are reduced in registers/smem and hit global atomics once per CTA
  int per_sm = (200 * 1024) / (smem + 2048);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = (N + 7) / 8;
  if (grid > num_sms * per_sm) grid = num_sms * per_sm;
  if (grid > 4096) grid = 4096;
  if (grid < 1) grid = 1;
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  if (wf) {   // fp32 per-RoI weights (reweight head variant)
    auto k = (dz_dtype == BAGS_DTYPE_F32) ? group_ce_kernel<NV, true, true> : group_ce_kernel<NV, false, true>;
    BAGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 256, smem, stream>>>(logits, ldz, lab, label2bin, classes, gt, wmask, avg, N, C, loss, lse, dz, ldd,
                                   colsum, part, counter);
  } else if (dz_dtype == BAGS_DTYPE_F32) {
    auto k = group_ce_kernel<NV, true>;
    BAGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 256, smem, stream>>>(logits, ldz, lab, label2bin, classes, gt, wmask, avg, N, C, loss, lse, dz, ldd,
                                   colsum, part, counter);
  } else {
    auto k = group_ce_kernel<NV, false>;
    BAGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<<<grid, 256, smem, stream>>>(logits, ldz, lab, label2bin, classes, gt, wmask, avg, N, C, loss, lse, dz, ldd,
                                   colsum, part, counter);
  }
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

static int group_ce_impl(const float* logits, long long ldz, const int64_t* labels,
                         const int32_t* label2bin, const int32_t* slices_host,
                         const uint8_t* wmask, bool wf, const float* avg, int N, int C, int G,
                         int classes, float* loss, float* lse, void* dz, long long ldd,
                         int dz_dtype, float* colsum, void* workspace, size_t workspace_bytes,
                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(label2bin && loss && workspace && (N == 0 || (logits && labels)), "bags_group_ce: NULL argument");
  BAGS_REQUIRE(workspace_bytes >= bags_workspace_bytes(), "bags_group_ce: workspace too small (%zu < %zu)",
               workspace_bytes, bags_workspace_bytes());
  BAGS_REQUIRE(N >= 0 && C >= 4 && (C % 4) == 0 && C <= 4096, "bags_group_ce: C=%d must be a multiple of 4 in [4,4096]", C);
  BAGS_REQUIRE((ldz % 4) == 0 && ldz >= C, "bags_group_ce: ldz=%lld must be a multiple of 4 and >= C", ldz);
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "bags_group_ce: logits not 16-byte aligned");
  if (N == 0) dz = nullptr;
  BAGS_REQUIRE(dz_dtype == BAGS_DTYPE_F32 || dz_dtype == BAGS_DTYPE_BF16, "bags_group_ce: bad dz dtype");
  if (dz != nullptr) {
    BAGS_REQUIRE(ldd >= C && (ldd % 8) == 0, "bags_group_ce: ldd=%lld must be a multiple of 8 and >= C", ldd);
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(dz) & 15) == 0, "bags_group_ce: dz not 16-byte aligned");
  }
  if (colsum != nullptr)
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(colsum) & 15) == 0, "bags_group_ce: colsum not 16-byte aligned");
  GroupTable gt;
  if (int rc = make_group_table(gt, slices_host, G, C)) return rc;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (colsum != nullptr) BAGS_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * C, stream));
  if (N == 0) {
    BAGS_CUDA(cudaMemsetAsync(loss, 0, sizeof(float) * G, stream));
    return BAGS_OK;
  }
  const int nv = (C / 4 + 31) / 32;
  if (nv <= 10)
    return launch_group_ce<10>(logits, ldz, labels, label2bin, gt, wmask, avg, N, C, classes, loss, lse, dz, ldd,
                               dz_dtype, colsum, workspace, di.num_sms, stream, wf);
  if (nv <= 16)
    return launch_group_ce<16>(logits, ldz, labels, label2bin, gt, wmask, avg, N, C, classes, loss, lse, dz, ldd,
                               dz_dtype, colsum, workspace, di.num_sms, stream, wf);
  return launch_group_ce<32>(logits, ldz, labels, label2bin, gt, wmask, avg, N, C, classes, loss, lse, dz, ldd,
                             dz_dtype, colsum, workspace, di.num_sms, stream, wf);
}

extern "C" int bags_group_ce(const float* logits, long long ldz, const int64_t* labels,
                             const int32_t* label2bin, const int32_t* slices_host,
                             const uint8_t* wmask, const float* avg, int N, int C, int G,
                             int classes, float* loss, float* lse, void* dz, long long ldd,
                             int dz_dtype, float* colsum, void* workspace, size_t workspace_bytes,
                             void* stream_) {
  return group_ce_impl(logits, ldz, labels, label2bin, slices_host, wmask, false, avg, N, C, G, classes, loss, lse, dz,
                       ldd, dz_dtype, colsum, workspace, workspace_bytes, stream_);
}

extern "C" int bags_group_ce_w(const float* logits, long long ldz, const int64_t* labels,
                               const int32_t* label2bin, const int32_t* slices_host,
                               const float* wfloat, const float* avg, int N, int C, int G,
                               int classes, float* loss, float* lse, void* dz, long long ldd,
                               int dz_dtype, float* colsum, void* workspace, size_t workspace_bytes,
                               void* stream_) {
  return group_ce_impl(logits, ldz, labels, label2bin, slices_host, reinterpret_cast<const uint8_t*>(wfloat),
                       wfloat != nullptr, avg, N, C, G, classes, loss, lse, dz, ldd, dz_dtype, colsum, workspace,
                       workspace_bytes, stream_);
}


// ----------------------------------------------------------------------------
// fused forward (GEMM + grouped softmax-CE in one kernel)
// ----------------------------------------------------------------------------
static bool fused_eligible(const int32_t* slices_host, int G, int C) {
  if (slices_host == nullptr || G < 1 || G > FusedCfg<false>::MAXG || C > 4 * FusedCfg<false>::BLOCK_N || (C % 4) != 0)
    return false;
  int end = 0;
  for (int g = 0; g < G; ++g) {   // bins must tile [0, C) contiguously
    if (slices_host[2 * g] != end || slices_host[2 * g + 1] < 1) return false;
    end += slices_host[2 * g + 1];
  }
  if (end != C) return false;
  for (int c0 = 0; c0 < C; c0 += 16) {   // at most two bins per 16-column chunk
    int cnt = 0;
    for (int g = 0; g < G; ++g) {
      const int s = slices_host[2 * g], e = s + slices_host[2 * g + 1];
      if (s < c0 + 16 && e > c0) ++cnt;
    }
    if (cnt > 2) return false;
  }
  return env_int("BAGS_FUSED", 1) != 0;
}

extern "C" int bags_fused_eligible(const int32_t* slices_host, int G, int C) {
  return fused_eligible(slices_host, G, C) ? 1 : 0;
}

template <bool TF32>
static int launch_fused_fwd(const void* x, long long ldx, const void* w, long long ldw, const FusedFwdParams& p0,
                            void* dz, long long ldd, const DeviceInfo& di, cudaStream_t stream, bool wf = false) {
  using Cfg = FusedCfg<TF32>;
  const int dtype = TF32 ? BAGS_DTYPE_F32 : BAGS_DTYPE_BF16;
  CUtensorMap tx, tw;
  int rc = make_tmap(&tx, x, dtype, p0.K, p0.N, ldx, Cfg::BLOCK_K, Cfg::BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tw, w, dtype, p0.K, p0.C, ldw, Cfg::BLOCK_K, Cfg::UMMA_N);
  if (rc) return rc;
  if (dz != nullptr && (reinterpret_cast<uintptr_t>(dz) & 15) != 0)
    return fail(BAGS_ERR_INVALID, "bags_fwd: dz must be 16-byte aligned");
  FusedFwdParams p = p0;
  p.dz = dz;
  p.ldd = ldd;
  p.kblocks = (p.K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;
  p.want_dz = dz != nullptr ? 1 : 0;
  p.timing = g_timing;
  p.dbg = g_timing ? g_dbg : 0;
  auto kernel = wf ? bags_fwd_fused_kernel<TF32, true> : bags_fwd_fused_kernel<TF32, false>;
  BAGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = Cfg::CLUSTER * ((p.N + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M);
  (void)di;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  // PDL INVARIANT (also the sampler's): this kernel reads x, W, bias and labels BEFORE its griddepcontrol.wait (only the
  // sampler's masks / avg and every global write come after it).  That is correct as long as those tensors are not
  // produced by the immediately preceding kernel of the stream with an early launch_dependents trigger -- true for torch
  // kernels (they never trigger early) and for this library's own chain (the predecessor is the sampler / the previous
  // step's backward or exchange, none of which writes them).  A future producer that triggers early must be followed by
  // a non-PDL launch (BAGS_PDL=0) or move these reads behind the wait.
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap with the sampler (pdl_wait in-kernel)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("BAGS_PDL", 1) ? 1 : 0;
  BAGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, tx, tw, p));
  return BAGS_OK;
}

static int fwd_impl(const void* x, long long ldx, const void* w, long long ldw,
                    const float* bias, const int64_t* labels, const int32_t* label2bin,
                    const int32_t* slices_host, const uint8_t* wmask, bool wf, const float* avg, int N,
                        int K, int C, int G, int classes, int dtype, float* logits, long long ldz,
                        float* loss, float* lse, void* dz, long long ldd, float* colsum,
                        int colsum_tiles, void* workspace, size_t workspace_bytes, void* stream_,
                        void* clear = nullptr, size_t clear_bytes = 0) {
  if (clear != nullptr) {
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(clear) & 15) == 0 && (clear_bytes % 16) == 0,
                 "bags_fwd_ex: the buffer to clear must be 16-byte aligned and a multiple of 16 bytes");
    if (logits != nullptr || N == 0) {   // not the fused kernel: a plain memset
      BAGS_CUDA(cudaMemsetAsync(clear, 0, clear_bytes, static_cast<cudaStream_t>(stream_)));
      clear = nullptr;
    }
  }
  if (colsum != nullptr)
    BAGS_REQUIRE(colsum_tiles >= 1 && (logits != nullptr || colsum_tiles == (N + 127) / 128 || N == 0),
                 "bags_fwd: colsum must hold ceil(N/128) = %d row tiles of C floats (got %d)", (N + 127) / 128, colsum_tiles);
  if (logits != nullptr) {
    // caller wants materialised logits: GEMM with fp32 store, then the stand-alone grouped CE (tile 0 holds
    // the column sums, the other tiles are zero)
    if (colsum != nullptr && colsum_tiles > 1)
      BAGS_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * C * colsum_tiles, static_cast<cudaStream_t>(stream_)));
    if (int rc = bags_linear_fwd(x, ldx, w, ldw, bias, logits, ldz, N, K, C, dtype, stream_)) return rc;
    return group_ce_impl(logits, ldz, labels, label2bin, slices_host, wmask, wf, avg, N, C, G, classes, loss,
                         lse, dz, ldd, dtype, colsum, workspace, workspace_bytes, stream_);
  }
  // fused path: logits stay in TMEM
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(fused_eligible(slices_host, G, C),
               "bags_fwd: logits == NULL requests the fused kernel, but this bin table / C=%d / G=%d is not eligible "
               "(see bags_fused_eligible); pass a logits workspace", C, G);
  BAGS_REQUIRE(dtype == BAGS_DTYPE_F32 || dtype == BAGS_DTYPE_BF16, "bags_fwd: bad dtype %d", dtype);
  BAGS_REQUIRE(label2bin && loss && workspace && (N == 0 || (x && w && labels)), "bags_fwd: NULL argument");
  BAGS_REQUIRE(workspace_bytes >= bags_workspace_bytes(), "bags_fwd: workspace too small");
  BAGS_REQUIRE(N >= 0 && K >= 1, "bags_fwd: bad shape");
  if (dz != nullptr) BAGS_REQUIRE(ldd >= C && (ldd % 8) == 0, "bags_fwd: ldd=%lld must be a multiple of 8 and >= C", ldd);
  GroupTable gt;
  if (int rc = make_group_table(gt, slices_host, G, C)) return rc;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (N == 0) {
    BAGS_CUDA(cudaMemsetAsync(loss, 0, sizeof(float) * G, stream));
    if (colsum != nullptr) BAGS_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * C * colsum_tiles, stream));
    return BAGS_OK;
  }
  if (bias != nullptr) BAGS_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "bags_fwd: bias must be 16-byte aligned");
  const int grid = 4 * ((N + 127) / 128);
  BAGS_REQUIRE(grid <= 4096, "bags_fwd: N=%d too large for the fused path's loss workspace", N);
  FusedFwdParams p{};
  p.N = N; p.C = C; p.K = K; p.gt = gt; p.bias = bias;
  p.labels = reinterpret_cast<const long long*>(labels);
  p.l2b = label2bin; p.classes = classes; p.wmask = wmask; p.avg = avg;
  p.loss = loss; p.lse = lse; p.colsum = colsum;
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  if (N == 0) dz = nullptr;
  p.clear = reinterpret_cast<float4*>(clear);
  p.clear_vecs = static_cast<long long>(clear_bytes / 16);
  return dtype == BAGS_DTYPE_BF16 ? launch_fused_fwd<false>(x, ldx, w, ldw, p, dz, ldd, di, stream, wf)
                                  : launch_fused_fwd<true>(x, ldx, w, ldw, p, dz, ldd, di, stream, wf);
}

extern "C" int bags_fwd(const void* x, long long ldx, const void* w, long long ldw,
                        const float* bias, const int64_t* labels, const int32_t* label2bin,
                        const int32_t* slices_host, const uint8_t* wmask, const float* avg, int N,
                        int K, int C, int G, int classes, int dtype, float* logits, long long ldz,
                        float* loss, float* lse, void* dz, long long ldd, float* colsum,
                        int colsum_tiles, void* workspace, size_t workspace_bytes, void* stream_) {
  return fwd_impl(x, ldx, w, ldw, bias, labels, label2bin, slices_host, wmask, false, avg, N, K, C, G, classes, dtype,
                  logits, ldz, loss, lse, dz, ldd, colsum, colsum_tiles, workspace, workspace_bytes, stream_);
}

extern "C" int bags_fwd_ex(const void* x, long long ldx, const void* w, long long ldw,
                           const float* bias, const int64_t* labels, const int32_t* label2bin,
                           const int32_t* slices_host, const uint8_t* wmask, const float* avg, int N,
                           int K, int C, int G, int classes, int dtype, float* logits, long long ldz,
                           float* loss, float* lse, void* dz, long long ldd, float* colsum,
                           int colsum_tiles, void* workspace, size_t workspace_bytes, void* clear, size_t clear_bytes,
                           void* stream_) {
  return fwd_impl(x, ldx, w, ldw, bias, labels, label2bin, slices_host, wmask, false, avg, N, K, C, G, classes, dtype,
                  logits, ldz, loss, lse, dz, ldd, colsum, colsum_tiles, workspace, workspace_bytes, stream_, clear,
                  clear_bytes);
}

extern "C" int bags_fwd_w(const void* x, long long ldx, const void* w, long long ldw,
                          const float* bias, const int64_t* labels, const int32_t* label2bin,
                          const int32_t* slices_host, const float* wfloat, const float* avg, int N,
                          int K, int C, int G, int classes, int dtype, float* logits, long long ldz,
                          float* loss, float* lse, void* dz, long long ldd, float* colsum,
                          int colsum_tiles, void* workspace, size_t workspace_bytes, void* stream_) {
  return fwd_impl(x, ldx, w, ldw, bias, labels, label2bin, slices_host, reinterpret_cast<const uint8_t*>(wfloat),
                  wfloat != nullptr, avg, N, K, C, G, classes, dtype, logits, ldz, loss, lse, dz, ldd, colsum,
                  colsum_tiles, workspace, workspace_bytes, stream_);
}

extern "C" int bags_reweight(const int64_t* labels, const int32_t* label2bin, const uint8_t* wmask,
                             const float* cls_weight, int wstride, int N, int G, int classes, float* wfloat,
                             float* avg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(label2bin && cls_weight && wfloat && avg && (N == 0 || labels), "bags_reweight: NULL argument");
  BAGS_REQUIRE(G >= 1 && G <= kMaxG && classes >= 1 && N >= 0 && wstride >= 1, "bags_reweight: bad shape");
  reweight_kernel<<<G, 256, 0, stream>>>(reinterpret_cast<const long long*>(labels), label2bin, classes, N, wmask,
                                         cls_weight, wstride, wfloat, avg);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

__global__ void __launch_bounds__(256)
scale_colsum_kernel(const float* __restrict__ colsum, int tiles, float* __restrict__ db, int C, GroupTable gt,
                    const float* __restrict__ gout) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float s = (gout == nullptr) ? 1.f : 0.f;
  if (gout != nullptr)
    for (int g = 0; g < gt.G; ++g)
      if (c >= gt.start[g] && c < gt.start[g] + gt.len[g]) s = __ldg(gout + g);
  float cs = 0.f;
  for (int t = 0; t < tiles; ++t) cs += colsum[(long long)t * C + c];
  db[c] = s * cs;
}


template <bool TF32, int MT>
static int launch_bwd_merged(const void* dz, long long ldd, const void* x, long long ldx, const void* wb, const void* w,
                             long long ldw, const BwdFusedParams& p0, const DeviceInfo& di, cudaStream_t stream) {
  using Cfg = BwdCfg<TF32, MT>;
  const int dtype = TF32 ? BAGS_DTYPE_F32 : BAGS_DTYPE_BF16;
  CUtensorMap t_dzT, t_xT, t_dz, t_wT, t_w;
  int rc;
  if ((rc = make_tmap(&t_w, w, dtype, p0.Kf, p0.C, ldw, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  if ((rc = make_tmap(&t_dzT, dz, dtype, p0.C, p0.Nr, ldd, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  if ((rc = make_tmap(&t_xT, x, dtype, p0.Kf, p0.Nr, ldx, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  if ((rc = make_tmap(&t_dz, dz, dtype, p0.C, p0.Nr, ldd, Cfg::BLOCK_K, Cfg::BLOCK_M))) return rc;
  if ((rc = make_tmap(&t_wT, wb, dtype, p0.Kf, p0.C, ldw, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  BwdFusedParams p = p0;
  p.timing = g_timing ? g_timing + 2048 * 8 : nullptr;   // rows [2048, ..): the forward of the same step uses [0, 2048)
  if (p.dw_units > 0 && p.dw_splits == 1 && p.prep_jobs > 0) {
    // no split-K: every dW unit owns its output tile -> plain stores, and the zeroing job (the first z_ctas tickets)
    // disappears.  (Only with in-kernel preparation: a separate preparation kernel has already been launched.)
    p.dw_store = 1;
    p.prep_jobs -= p.prep.z_ctas;
    p.prep.z_ctas = 0;
    if (p.prep.c_ctas == 0) p.dw_needs_prep = 0;
  }
  // preparation jobs handed out by an atomic ticket (no co-residency requirement); BAGS_BWD_TICKET=0: static assignment
  auto kernel = (p.prep_jobs > 0 && env_int("BAGS_BWD_TICKET", 1)) ? bags_bwd_fused_kernel<TF32, MT, true>
                                                                  : bags_bwd_fused_kernel<TF32, MT, false>;
  BAGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int units = p.dw_units + p.dx_units;
  const int grid = units < di.num_sms ? units : di.num_sms;
  BAGS_CUDA(launch_pdl(kernel, dim3(grid), dim3(Cfg::NUM_THREADS), Cfg::SMEM_BYTES, stream, t_dzT, t_xT, t_dz, t_wT, t_w, p));
  return BAGS_OK;
}

// CTA pairs that can be co-resident (GPCs with an odd number of usable SMs leave one SM without a partner)
template <bool TF32>
static int pair_capacity(const DeviceInfo& di) {
  static int cached[2] = {0, 0};
  int& c = cached[TF32 ? 1 : 0];
  if (c > 0) return c;
  using Cfg = BwdPairCfg<TF32>;
  auto kernel = bags_bwd_pair_kernel<TF32>;
  int n = di.num_sms / 2;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (di.num_sms / 2));
    cfg.blockDim = dim3(Cfg::NUM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, kernel, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
    else (void)cudaGetLastError();
  }
  const int forced = env_int("BAGS_PAIRS", 0);
  if (forced > 0 && forced < n) n = forced;
  c = n;
  return c;
}

// split-K factor of the dW units for the pair kernel: fewest (rounds x longest unit), units costed in k-blocks
static int pick_pair_splits(int dw_tiles, int dw_kblocks, int dx_units, int dx_kblocks, int pairs) {
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= 16 && s <= dw_kblocks; ++s) {
    const int units = dw_tiles * s + dx_units;
    const int rounds = (units + pairs - 1) / pairs;
    const int dwk = (dw_kblocks + s - 1) / s;
    const long unit = ((dx_units > 0 && dx_kblocks > dwk) ? dx_kblocks : dwk) + 4;   // + epilogue, in k-block equivalents
    const long cost = rounds * unit;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// CTA-pair (cta_group::2) variant: cluster of 2, units are 256-row pairs
template <bool TF32>
static int launch_bwd_pair(const void* dz, long long ldd, const void* x, long long ldx, const void* wb, long long ldw,
                           const BwdFusedParams& p0, const DeviceInfo& di, cudaStream_t stream) {
  using Cfg = BwdPairCfg<TF32>;
  const int dtype = TF32 ? BAGS_DTYPE_F32 : BAGS_DTYPE_BF16;
  CUtensorMap t_dzT, t_xT, t_dz, t_wT;
  int rc;
  if ((rc = make_tmap(&t_dzT, dz, dtype, p0.C, p0.Nr, ldd, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  if ((rc = make_tmap(&t_xT, x, dtype, p0.Kf, p0.Nr, ldx, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  if ((rc = make_tmap(&t_dz, dz, dtype, p0.C, p0.Nr, ldd, Cfg::BLOCK_K, Cfg::BLOCK_M))) return rc;
  if ((rc = make_tmap(&t_wT, wb, dtype, p0.Kf, p0.C, ldw, Cfg::SLAB, Cfg::BLOCK_K, TF32))) return rc;
  BwdFusedParams p = p0;
  p.timing = g_timing ? g_timing + 2048 * 8 : nullptr;
  auto kernel = bags_bwd_pair_kernel<TF32>;
  BAGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int units = p.dw_units + p.dx_units;
  const int max_pairs = pair_capacity<TF32>(di);
  const int pairs = units < max_pairs ? units : max_pairs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("BAGS_PDL", 1) ? 2 : 1;
  BAGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, t_dzT, t_xT, t_dz, t_wT, p));
  return BAGS_OK;
}

static constexpr int kColsumTiles = 8;   // row groups of the bias-gradient partial sums made by bwd_prep

extern "C" size_t bags_bwd_scratch_bytes(int C, long long ldw, int dtype) {
  const size_t elt = (dtype == BAGS_DTYPE_BF16) ? 2 : 4;
  const size_t wbytes = (static_cast<size_t>(C) * static_cast<size_t>(ldw) * elt + 255) & ~static_cast<size_t>(255);
  return wbytes + static_cast<size_t>(kColsumTiles) * C * sizeof(float);
}

extern "C" int bags_bwd_ex(const void* dz, long long ldd, const void* x, long long ldx, const void* w,
                           long long ldw, const float* gout, const int32_t* slices_host,
                           const float* colsum, int colsum_tiles, float* dW, long long lddw, float* db, void* dX,
                           long long lddx, void* wscratch, size_t wscratch_bytes, int N, int K, int C, int G,
                           int dtype, int flags, void* stream_);

extern "C" int bags_bwd(const void* dz, long long ldd, const void* x, long long ldx, const void* w,
                        long long ldw, const float* gout, const int32_t* slices_host,
                        const float* colsum, int colsum_tiles, float* dW, long long lddw, float* db, void* dX,
                        long long lddx, void* wscratch, size_t wscratch_bytes, int N, int K, int C, int G,
                        int dtype, void* stream_) {
  return bags_bwd_ex(dz, ldd, x, ldx, w, ldw, gout, slices_host, colsum, colsum_tiles, dW, lddw, db, dX, lddx, wscratch,
                     wscratch_bytes, N, K, C, G, dtype, 0, stream_);
}

extern "C" int bags_bwd_ex(const void* dz, long long ldd, const void* x, long long ldx, const void* w,
                           long long ldw, const float* gout, const int32_t* slices_host,
                           const float* colsum, int colsum_tiles, float* dW, long long lddw, float* db, void* dX,
                           long long lddx, void* wscratch, size_t wscratch_bytes, int N, int K, int C, int G,
                           int dtype, int flags, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool prezeroed = (flags & BAGS_BWD_DW_PREZEROED) != 0;   // the caller (bags_fwd_ex's clear hook) zeroed dW
  BAGS_REQUIRE(dz != nullptr || N == 0, "bags_bwd: dz is NULL");
  BAGS_REQUIRE(dtype == BAGS_DTYPE_F32 || dtype == BAGS_DTYPE_BF16, "bags_bwd: bad dtype %d", dtype);
  BAGS_REQUIRE(N >= 0 && K >= 1 && C >= 1, "bags_bwd: bad shape");
  GroupTable gt;
  if (int rc = make_group_table(gt, slices_host, G, C)) return rc;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  const bool bf = dtype == BAGS_DTYPE_BF16;
  const int vecw = bf ? 8 : 4;

  const bool want_scale = (dX != nullptr && N > 0 && gout != nullptr);
  const bool want_colpart = (db != nullptr && colsum == nullptr && N > 0);
  if (db != nullptr && colsum == nullptr && N == 0) {   // no rows: zero bias gradient
    BAGS_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, stream));
    db = nullptr;
  }
  if (want_scale || want_colpart) {
    BAGS_REQUIRE(wscratch != nullptr && wscratch_bytes >= bags_bwd_scratch_bytes(C, ldw, dtype),
                 "bags_bwd: wscratch must provide bags_bwd_scratch_bytes() = %zu bytes (got %zu)",
                 bags_bwd_scratch_bytes(C, ldw, dtype), wscratch_bytes);
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(wscratch) & 255) == 0, "bags_bwd: wscratch must be 256-byte aligned");
  }
  float* colpart = nullptr;
  if (wscratch != nullptr) {
    const size_t elt = bf ? 2 : 4;
    const size_t wbytes = (static_cast<size_t>(C) * static_cast<size_t>(ldw) * elt + 255) & ~static_cast<size_t>(255);
    colpart = reinterpret_cast<float*>(reinterpret_cast<char*>(wscratch) + wbytes);
  }
  if (dW != nullptr)
    BAGS_REQUIRE((x != nullptr || N == 0) && lddw >= K && (lddw % 4) == 0 && (K % 4) == 0 &&
                     (reinterpret_cast<uintptr_t>(dW) & 15) == 0,
                 "bags_bwd: dW must be 16-byte aligned with K and lddw multiples of 4");
  if (want_scale)
    BAGS_REQUIRE(w != nullptr && (K % vecw) == 0 && (ldw % vecw) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
                 "bags_bwd: w must be 16-byte aligned with K and ldw multiples of %d", vecw);
  if (want_colpart)
    BAGS_REQUIRE(((ldd * (bf ? 2 : 4)) % 16) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0,
                 "bags_bwd: dz rows must be 16-byte aligned");

  // ---- preparation: zero dW, W' = gout-scaled W, bias-gradient partial column sums.  On the merged single-CTA path
  // these jobs run inside the GEMM kernel (no extra launch in the dependent chain); otherwise one small kernel. ----
  // dW (+ db) alone also runs on the merged kernel (no dX units): in-kernel preparation, 256 x 256 units when the
  // problem is large enough -- the launch the split in-step schedule (dW -> exchange || dX) puts on its critical path
  // ... unless there is nothing to prepare (dW zeroed and the bias-gradient partials supplied by the forward): then the
  // plain split-K GEMM (4-stage ring, no job machinery) is the faster dW-only launch -- 52.9 vs 53.7 us per
  // forward + dW-only + dX-only sequence, and 54.6 with the jobs in the backward (profiles/r02_dwonly_path.log)
  const bool nothing_to_prepare = prezeroed && (db == nullptr || colsum != nullptr);
  const bool dw_only_merged = dW != nullptr && dX == nullptr && w != nullptr && N > 0 && (K % 8) == 0 &&
                              (reinterpret_cast<uintptr_t>(w) & 15) == 0 && !env_int("BAGS_BWD_PAIR", 0) &&
                              env_int("BAGS_BWD_DW_MERGED", nothing_to_prepare ? 0 : 1);
  // dX alone with per-bin upstream gradients: the merged kernel decides ON THE DEVICE whether they are uniform (then it
  // reads W and scales in the epilogue: no scaled copy W', no preparation launch); with gout == NULL the plain GEMM is used
  const bool dx_only_merged = dW == nullptr && db == nullptr && dX != nullptr && gout != nullptr && x != nullptr &&
                              w != nullptr && N > 0 && (K % 8) == 0 && env_int("BAGS_BWD_DX_MERGED", 1) &&
                              !env_int("BAGS_BWD_PAIR", 0);
  const bool merged_path = ((dW != nullptr && dX != nullptr && N > 0 && (K % 8) == 0) || dw_only_merged || dx_only_merged) &&
                           env_int("BAGS_BWD_MERGED", 1);
  unsigned int* sync = nullptr;
  if (merged_path && !env_int("BAGS_BWD_PAIR", 0) && env_int("BAGS_BWD_INKERNEL_PREP", 1)) {
    int dev = 0;
    BAGS_CUDA(cudaGetDevice(&dev));
    sync = sync_slot(dev, stream);
  }
  BwdPrepParams pp{};
  int prep_jobs = 0;
  bool prep_launched = false;
  if (dW != nullptr || want_scale || want_colpart) {
    pp.dW = dW; pp.lddw = lddw; pp.C = C; pp.K = K;
    pp.w = w; pp.wscr = wscratch; pp.ldw = ldw; pp.gout = gout; pp.gt = gt;
    pp.dz = dz; pp.ldd = ldd; pp.N = N; pp.colpart = colpart; pp.ctiles = kColsumTiles;
    pp.z_ctas = (dW != nullptr && !prezeroed) ? di.num_sms : 0;
    pp.s_ctas = want_scale ? di.num_sms : 0;
    pp.c_ctas = want_colpart ? ((C + 63) / 64) * kColsumTiles : 0;
    const int grid = pp.z_ctas + pp.s_ctas + pp.c_ctas;
    if (grid == 0) { /* nothing to prepare */ }
    else if (sync != nullptr) prep_jobs = grid;
    else if (bf) { BAGS_CUDA(launch_pdl(bwd_prep_kernel<false>, dim3(grid), dim3(256), 0, stream, pp)); prep_launched = true; }
    else         { BAGS_CUDA(launch_pdl(bwd_prep_kernel<true>, dim3(grid), dim3(256), 0, stream, pp)); prep_launched = true; }
  }
  const float* cs_in = (colsum != nullptr) ? colsum : colpart;
  const int cs_tiles = (colsum != nullptr) ? colsum_tiles : kColsumTiles;
  if (db != nullptr) BAGS_REQUIRE(cs_in != nullptr && cs_tiles >= 1, "bags_bwd: db requested but no column sums available");

  // ---- both contractions in one persistent launch when both are requested ----
  if (merged_path) {
    BAGS_REQUIRE(w != nullptr, "bags_bwd: w is NULL but dX requested");
    if (dX != nullptr)
      BAGS_REQUIRE((reinterpret_cast<uintptr_t>(dX) & 15) == 0 && ((lddx * (bf ? 2 : 4)) % 16) == 0,
                   "bags_bwd: dX rows must be 16-byte aligned");
    const int bk = bf ? 64 : 32;
    BwdFusedParams bp{};
    bp.C = C; bp.Kf = K; bp.Nr = N;
    bp.dw_m_tiles = (C + 127) / 128; bp.dw_n_tiles = (K + 255) / 256; bp.dw_kblocks = (N + bk - 1) / bk;
    bp.dw_splits = env_int("BAGS_DW_SPLITS", pick_splits(bp.dw_m_tiles * bp.dw_n_tiles, bp.dw_kblocks, di.num_sms));
    if (bp.dw_splits > bp.dw_kblocks) bp.dw_splits = bp.dw_kblocks;
    bp.dx_m_tiles = (N + 127) / 128; bp.dx_n_tiles = (K + 255) / 256; bp.dx_kblocks = (C + bk - 1) / bk;
    bp.dw_units = (dW != nullptr) ? bp.dw_m_tiles * bp.dw_n_tiles * bp.dw_splits : 0;
    bp.dx_units = (dX != nullptr) ? bp.dx_m_tiles * bp.dx_n_tiles : 0;
    bp.dW = dW; bp.lddw = lddw; bp.dX = dX; bp.lddx = lddx;
    bp.gscale = gout; bp.G = (gout != nullptr) ? gt.G : 0;
    for (int g = 0; g < kMaxGroups; ++g) { bp.gstart[g] = gt.start[g]; bp.glen[g] = gt.len[g]; }
    bp.colsum_in = (db != nullptr) ? cs_in : nullptr;
    bp.colsum_tiles = cs_tiles;
    bp.db = db;
    bp.prep = pp; bp.prep_jobs = prep_jobs; bp.sync = sync;
    // the first operand load must wait for the producer of dz unless a preparation kernel sits in between (its own wait
    // covers it); the dW epilogues only depend on the zeroing / column-sum jobs
    bp.wait_dz = (prep_jobs > 0 || !prep_launched) ? 1 : 0;
    bp.dw_needs_prep = (pp.z_ctas > 0 || pp.c_ctas > 0 || prep_launched) ? 1 : 0;
    const void* wb = want_scale ? wscratch : w;
    if (env_int("BAGS_BWD_PAIR", 0)) {
      // units are 256-row CTA pairs
      bp.dw_m_tiles = (bp.dw_m_tiles + 1) / 2;
      bp.dx_m_tiles = (bp.dx_m_tiles + 1) / 2;
      const int cap = bf ? pair_capacity<false>(di) : pair_capacity<true>(di);
      bp.dw_splits = env_int("BAGS_DW_SPLITS", pick_pair_splits(bp.dw_m_tiles * bp.dw_n_tiles, bp.dw_kblocks,
                                                                bp.dx_m_tiles * bp.dx_n_tiles, bp.dx_kblocks, cap));
      if (bp.dw_splits > bp.dw_kblocks) bp.dw_splits = bp.dw_kblocks;
      bp.dw_units = bp.dw_m_tiles * bp.dw_n_tiles * bp.dw_splits;
      bp.dx_units = bp.dx_m_tiles * bp.dx_n_tiles;
      const int only = env_int("BAGS_PAIR_ONLY", 0);   // timing experiments: 1 = dW units only, 2 = dX units only
      if (only == 1) bp.dx_units = 0;
      if (only == 2) bp.dw_units = 0;
      return bf ? launch_bwd_pair<false>(dz, ldd, x, ldx, wb, ldw, bp, di, stream)
                : launch_bwd_pair<true>(dz, ldd, x, ldx, wb, ldw, bp, di, stream);
    }
    bp.dx_uniform_ok = (want_scale && env_int("BAGS_DX_UNIFORM", 1)) ? 1 : 0;
    bp.prep.skip_scale_if_uniform = (bp.dx_uniform_ok && prep_jobs > 0) ? 1 : 0;
    // 256 x 256 units (two accumulator sub-tiles sharing the B tile) once the problem fills the machine with them
    int mt_auto = (bp.dx_units + bp.dw_m_tiles * bp.dw_n_tiles >= di.num_sms) ? 2 : 1;
    if (dW == nullptr) mt_auto = (bp.dx_units / 2 >= di.num_sms) ? 2 : 1;   // dX alone: 128-row units fill the machine sooner
    // dW alone: 128 x 256 units.  Filling the machine with 256 x 256 units takes a 7-way split at the benchmark shape, i.e.
    // 35 MB of red.add for a 5 MB result -- the L2 atomic rate (~3 TB/s) then costs more than the larger tile saves
    if (dX == nullptr) mt_auto = env_int("BAGS_DW_ONLY_MT", 1);
    if (env_int("BAGS_BWD_MT", mt_auto) == 2) {
      bp.dw_m_tiles = (C + 255) / 256;
      bp.dx_m_tiles = (N + 255) / 256;
      bp.dx_units = (dX != nullptr) ? bp.dx_m_tiles * bp.dx_n_tiles : 0;
      bp.dw_splits = env_int("BAGS_DW_SPLITS", pick_pair_splits(bp.dw_m_tiles * bp.dw_n_tiles, bp.dw_kblocks, bp.dx_units,
                                                                bp.dx_kblocks, di.num_sms));
      if (bp.dw_splits > bp.dw_kblocks) bp.dw_splits = bp.dw_kblocks;
      bp.dw_units = (dW != nullptr) ? bp.dw_m_tiles * bp.dw_n_tiles * bp.dw_splits : 0;
      return bf ? launch_bwd_merged<false, 2>(dz, ldd, x, ldx, wb, w, ldw, bp, di, stream)
                : launch_bwd_merged<true, 2>(dz, ldd, x, ldx, wb, w, ldw, bp, di, stream);
    }
    return bf ? launch_bwd_merged<false, 1>(dz, ldd, x, ldx, wb, w, ldw, bp, di, stream)
              : launch_bwd_merged<true, 1>(dz, ldd, x, ldx, wb, w, ldw, bp, di, stream);
  }

  if (dW != nullptr && N > 0) {
    GemmArgs ga{};
    ga.a = dz; ga.lda = ldd; ga.a_mn = true;   // A = dz^T : [C, N_roi], stored [N_roi, C]
    ga.b = x;  ga.ldb = ldx; ga.b_mn = true;   // B = x^T  : [K, N_roi], stored [N_roi, K]
    ga.M = C; ga.N = K; ga.K = N; ga.dtype = dtype;
    const int tiles = ((C + 127) / 128) * ((K + 255) / 256);
    const int kblocks = (N + (bf ? 63 : 31)) / (bf ? 64 : 32);
    ga.splits = env_int("BAGS_DW_SPLITS", pick_splits(tiles, kblocks, di.num_sms));
    ga.p.out = dW; ga.p.ldo = lddw; ga.p.bias = nullptr;
    ga.p.gscale = gout; ga.p.G = (gout != nullptr) ? gt.G : 0;
    for (int g = 0; g < kMaxGroups; ++g) { ga.p.gstart[g] = gt.start[g]; ga.p.glen[g] = gt.len[g]; }
    ga.p.colsum_in = (db != nullptr) ? cs_in : nullptr;
    ga.p.colsum_tiles = cs_tiles;
    ga.p.colsum_out = (db != nullptr) ? db : nullptr;
    ga.p.pdl_wait_epilogue = 1;   // dW zeroing + column-sum partials come from bwd_prep; the mainloop overlaps it
    ga.p.pdl_wait_producer = prep_launched ? 0 : 1;   // no preparation kernel in between: dz comes from the preceding kernel
    int rc = bf ? launch_gemm_pdl<256, true, true, EPI_RED_F32, false, 4>(ga, di, stream)
                : launch_gemm_pdl<256, true, true, EPI_RED_F32, true, 4>(ga, di, stream);
    if (rc) return rc;
  } else if (db != nullptr) {
    scale_colsum_kernel<<<(C + 255) / 256, 256, 0, stream>>>(cs_in, cs_tiles, db, C, gt, gout);
    BAGS_CUDA(cudaGetLastError());
  }

  if (dX != nullptr && N > 0) {
    BAGS_REQUIRE(w != nullptr, "bags_bwd: w is NULL but dX requested");
    BAGS_REQUIRE((reinterpret_cast<uintptr_t>(dX) & 15) == 0, "bags_bwd: dX not 16-byte aligned");
    GemmArgs ga{};
    ga.a = dz; ga.lda = ldd; ga.a_mn = false;                 // A = dz : [N_roi, C]
    ga.b = want_scale ? wscratch : w; ga.ldb = ldw; ga.b_mn = true;   // B = W'^T: [K, C], stored [C, K]
    ga.M = N; ga.N = K; ga.K = C; ga.dtype = dtype; ga.splits = 1;
    ga.p.out = dX; ga.p.ldo = lddx; ga.p.bias = nullptr; ga.p.gscale = nullptr; ga.p.G = 0;
    ga.p.colsum_in = nullptr; ga.p.colsum_out = nullptr;
    ga.p.pdl_wait_producer = 1;   // W' is produced by bwd_prep (two kernels back; the chain of waits covers it)
    int rc = bf ? launch_gemm_pdl<256, false, true, EPI_STORE_BF16, false, 4>(ga, di, stream)
                : launch_gemm_pdl<256, false, true, EPI_STORE_F32, true, 4>(ga, di, stream);
    if (rc) return rc;
  }
  return BAGS_OK;
}

// ----------------------------------------------------------------------------
// gradient exchange over NVLink peer memory (one kernel; see bags_allreduce.cuh)
// ----------------------------------------------------------------------------
extern "C" size_t bags_grad_allreduce_flag_bytes(int world) {
  if (world < 1) world = 1;
  // slots [kArMaxBlocks][world] + epochs [kArMaxBlocks] + one status word, padded to 64 bytes
  return (static_cast<size_t>(kArMaxBlocks) * static_cast<size_t>(world) + kArMaxBlocks + 16) * sizeof(uint32_t);
}

extern "C" long long bags_grad_allreduce_status_offset(int world) {
  if (world < 1) world = 1;
  return (static_cast<long long>(kArMaxBlocks) * world + kArMaxBlocks) * static_cast<long long>(sizeof(uint32_t));
}

extern "C" int bags_grad_allreduce(void* const* peer_bufs_host, void* mc_buf, long long flag_off_bytes,
                                   long long count, int rank, int world, float scale, int max_blocks,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(peer_bufs_host != nullptr, "bags_grad_allreduce: peer_bufs_host is NULL");
  BAGS_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "bags_grad_allreduce: bad rank %d / world %d (max %d ranks)", rank, world, kMaxRanks);
  BAGS_REQUIRE(count >= 0 && (count % 4) == 0, "bags_grad_allreduce: count=%lld must be a multiple of 4 floats", count);
  BAGS_REQUIRE(flag_off_bytes >= count * 4 && (flag_off_bytes % 16) == 0,
               "bags_grad_allreduce: the flag words must follow the data (flag_off_bytes=%lld, data bytes=%lld)",
               flag_off_bytes, count * 4);
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  AllReduceParams p{};
  for (int r = 0; r < world; ++r) {
    BAGS_REQUIRE(peer_bufs_host[r] != nullptr && (reinterpret_cast<uintptr_t>(peer_bufs_host[r]) & 15) == 0,
                 "bags_grad_allreduce: peer buffer %d is NULL or not 16-byte aligned", r);
    p.peer[r] = static_cast<float*>(peer_bufs_host[r]);
  }
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(mc_buf) & 15) == 0, "bags_grad_allreduce: multicast buffer not 16-byte aligned");
  p.mc = static_cast<float*>(mc_buf);
  p.flag_off = flag_off_bytes;
  p.count = count;
  p.rank = rank; p.world = world; p.scale = scale;
  // max_blocks < 0: soft-failure mode of the construction-time self test (grid = default)
  p.trap_on_timeout = (max_blocks < 0) ? 0 : 1;
  if (max_blocks < 0) max_blocks = 0;
  p.timeout_ns = 1000000LL * env_int("BAGS_AR_TIMEOUT_MS", p.trap_on_timeout ? 30000 : 3000);
  p.timing = g_timing ? g_timing + 4096 * 8 : nullptr;   // rows [4096, ..): after the forward's and the backward's
  p.mode = env_int("BAGS_AR_MODE", 3);   // relaxed first-barrier store + relaxed polling: -3..4 us per exchange
  if (count == 0) return BAGS_OK;
  // enough threads to keep one vector per thread and unroll slot in flight, at most kArMaxBlocks blocks;
  // every rank must launch the same grid: it depends only on (count, world, max_blocks) and the environment
  int threads = env_int("BAGS_AR_THREADS", 256);
  if (threads != 128 && threads != 256) threads = 256;
  if (max_blocks > kArMaxBlocks) max_blocks = kArMaxBlocks;
  const long long per_rank = (count / 4 + world - 1) / world;
  long long blocks = (per_rank + static_cast<long long>(threads) * 4 - 1) / (static_cast<long long>(threads) * 4);
  if (blocks < 1) blocks = 1;
  const bool mm = p.mc != nullptr && !env_int("BAGS_AR_NO_MULTIMEM", 0);
  // Default grid (measured inside the step at 2 ranks, profiles/r02_exchange_2gpu.md): one block per SM on the multimem
  // path (a 16-block grid, right for an exchange hidden under the NEXT step in round 1, costs 2x inside the step); the
  // plain peer path keeps one vector per thread in flight.
  if (max_blocks <= 0) max_blocks = env_int("BAGS_AR_MAX_BLOCKS", mm ? 148 : kArMaxBlocks);
  if (max_blocks <= 0 || max_blocks > kArMaxBlocks) max_blocks = kArMaxBlocks;
  if (blocks > max_blocks) blocks = max_blocks;
  const bool epoch = env_int("BAGS_AR_EPOCH", 1) != 0;
  const dim3 grid(static_cast<unsigned>(blocks)), block(static_cast<unsigned>(threads));
  if (mm && epoch)       BAGS_CUDA(launch_pdl(bags_grad_allreduce_kernel<true, true>, grid, block, 0, stream, p));
  else if (mm)           BAGS_CUDA(launch_pdl(bags_grad_allreduce_kernel<true, false>, grid, block, 0, stream, p));
  else if (epoch)        BAGS_CUDA(launch_pdl(bags_grad_allreduce_kernel<false, true>, grid, block, 0, stream, p));
  else                   BAGS_CUDA(launch_pdl(bags_grad_allreduce_kernel<false, false>, grid, block, 0, stream, p));
  return BAGS_OK;
}

// ----------------------------------------------------------------------------
// class-aware batched NMS (see bags_nms.cuh)
// ----------------------------------------------------------------------------
extern "C" int bags_class_nms_dense(const float* boxes, int box_cols, const int32_t* order, const int32_t* counts,
                                    int num_classes_fg, int n, float iou_thr, uint8_t* keep, int32_t* overflow,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(num_classes_fg >= 0 && n >= 0, "bags_class_nms_dense: bad sizes");
  if (num_classes_fg == 0 || n == 0) return BAGS_OK;
  BAGS_REQUIRE(boxes && order && counts && keep && overflow, "bags_class_nms_dense: NULL argument");
  BAGS_REQUIRE(box_cols == 4 || box_cols == 4 * (num_classes_fg + 1),
               "bags_class_nms_dense: boxes must have 4 or 4 * (classes incl. background) = %d columns (got %d)",
               4 * (num_classes_fg + 1), box_cols);
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "bags_class_nms_dense: boxes must be 16-byte aligned");
  const int seg = n < kNmsMaxSeg ? n : kNmsMaxSeg;
  const int words = (seg + 31) / 32;
  const size_t smem = static_cast<size_t>((seg + 1) & ~1) * sizeof(float4) + static_cast<size_t>(seg) * words * sizeof(uint32_t);
  BAGS_CUDA(cudaFuncSetAttribute(class_nms_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  class_nms_dense_kernel<<<num_classes_fg, kNmsThreads, smem, stream>>>(boxes, box_cols, order, counts, n, iou_thr, keep, overflow);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

// test hook: a kernel of `blocks` x `threads` that only waits `micros` microseconds (stands in for a latency-bound
// collective when probing how side-stream work co-schedules with the GEMM kernels)
__global__ void bags_debug_spin_kernel(long long ns) {
  const long long t0 = global_timer_ns();
  while (global_timer_ns() - t0 < ns) __nanosleep(200);
}
extern "C" int bags_debug_spin(int blocks, int threads, int micros, void* stream_) {
  BAGS_REQUIRE(blocks >= 1 && threads >= 32 && threads <= 1024 && micros >= 0, "bags_debug_spin: bad arguments");
  bags_debug_spin_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream_)>>>(1000LL * micros);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

// test hook: how many clusters of `cluster` CTAs (each `threads` threads + `smem_bytes` dynamic shared memory, i.e. one CTA
// per SM for the GEMM-sized value) can be resident at once -- the GPC layout decides (148 SMs do not split evenly)
extern "C" int bags_debug_max_clusters(int cluster, int threads, int smem_bytes) {
  if (cluster < 1 || cluster > 16 || threads < 32 || threads > 1024 || smem_bytes < 0) return -1;
  DeviceInfo di;
  if (device_info(di)) return -1;
  if (cudaFuncSetAttribute(bags_debug_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  if (cluster > 8) (void)cudaFuncSetAttribute(bags_debug_spin_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(cluster * ((2 * di.num_sms) / cluster)));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, bags_debug_spin_kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  return n;
}

extern "C" int bags_merge_scores(const float* logits, long long ldz, const int32_t* slices_host,
                                 const int32_t* cls2col, int N, int C, int G, int classes,
                                 float* scores, long long lds, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(cls2col && (N == 0 || (logits && scores)), "bags_merge_scores: NULL argument");
  BAGS_REQUIRE(N >= 0 && C >= 4 && (C % 4) == 0 && C <= 4096, "bags_merge_scores: C=%d must be a multiple of 4 in [4,4096]", C);
  BAGS_REQUIRE((ldz % 4) == 0 && ldz >= C && lds >= classes, "bags_merge_scores: bad leading dimension");
  BAGS_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "bags_merge_scores: logits not 16-byte aligned");
  GroupTable gt;
  if (int rc = make_group_table(gt, slices_host, G, C)) return rc;
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  if (N == 0) return BAGS_OK;
  const int nv = (C / 4 + 31) / 32;
  int grid = (N + 7) / 8;
  if (grid > di.num_sms * 4) grid = di.num_sms * 4;
#define BAGS_LAUNCH_MERGE(NV)                                                                              \
  do {                                                                                                     \
    const int smem = 8 * NV * 128 * (int)sizeof(float);                                                    \
    auto k = merge_scores_kernel<NV>;                                                                      \
    BAGS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));               \
    k<<<grid, 256, smem, stream>>>(logits, ldz, gt, cls2col, N, C, classes, scores, lds);                  \
  } while (0)
  if (nv <= 10) BAGS_LAUNCH_MERGE(10);
  else if (nv <= 16) BAGS_LAUNCH_MERGE(16);
  else BAGS_LAUNCH_MERGE(32);
#undef BAGS_LAUNCH_MERGE
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

extern "C" int bags_cast_bf16(const float* src, long long lds, void* dst, long long ldd, int rows,
                              int cols, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(src && dst, "bags_cast_bf16: NULL argument");
  BAGS_REQUIRE(rows >= 0 && cols >= 0 && (cols % 4) == 0 && (lds % 4) == 0 && (ldd % 4) == 0,
               "bags_cast_bf16: cols and leading dims must be multiples of 4");
  if (rows == 0 || cols == 0) return BAGS_OK;
  const long long total = (long long)rows * (cols / 4);
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  cast_bf16_kernel<<<(int)grid, 256, 0, stream>>>(src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  BAGS_CUDA(cudaGetLastError());
  return BAGS_OK;
}

extern "C" int bags_gemm_probe(const void* a, long long lda, int a_mn, const void* b, long long ldb,
                               int b_mn, void* out, long long ldo, int M, int N, int K, int dtype,
                               int block_n, int splits, int epi, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BAGS_REQUIRE(a && b && out, "bags_gemm_probe: NULL argument");
  DeviceInfo di;
  if (int rc = device_info(di)) return rc;
  GemmArgs ga{};
  ga.a = a; ga.lda = lda; ga.a_mn = a_mn != 0;
  ga.b = b; ga.ldb = ldb; ga.b_mn = b_mn != 0;
  ga.M = M; ga.N = N; ga.K = K; ga.dtype = dtype; ga.splits = splits;
  ga.p.out = out; ga.p.ldo = ldo; ga.p.bias = nullptr; ga.p.gscale = nullptr; ga.p.G = 0;
  ga.p.colsum_in = nullptr; ga.p.colsum_out = nullptr;
  const bool bf = dtype == BAGS_DTYPE_BF16;
  BAGS_REQUIRE(bf || dtype == BAGS_DTYPE_F32, "bags_gemm_probe: bad dtype");
  // the instantiations the product uses
  if (!a_mn && !b_mn && epi == 0 && block_n == 320)
    return bf ? launch_gemm<320, false, false, EPI_STORE_F32, false, 3>(ga, di, stream)
              : launch_gemm<320, false, false, EPI_STORE_F32, true, 3>(ga, di, stream);
  if (!a_mn && !b_mn && epi == 0 && block_n == 256)
    return bf ? launch_gemm<256, false, false, EPI_STORE_F32, false, 4>(ga, di, stream)
              : launch_gemm<256, false, false, EPI_STORE_F32, true, 4>(ga, di, stream);
  if (!a_mn && b_mn && block_n == 256 && ((bf && epi == 1) || (!bf && epi == 0)))
    return bf ? launch_gemm<256, false, true, EPI_STORE_BF16, false, 4>(ga, di, stream)
              : launch_gemm<256, false, true, EPI_STORE_F32, true, 4>(ga, di, stream);
  if (a_mn && b_mn && block_n == 256 && epi == 2)
    return bf ? launch_gemm<256, true, true, EPI_RED_F32, false, 4>(ga, di, stream)
              : launch_gemm<256, true, true, EPI_RED_F32, true, 4>(ga, di, stream);
  return fail(BAGS_ERR_INVALID, "bags_gemm_probe: configuration a_mn=%d b_mn=%d block_n=%d epi=%d dtype=%d not instantiated",
              a_mn, b_mn, block_n, epi, dtype);
}
