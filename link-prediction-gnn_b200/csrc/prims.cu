// prims.cu - device-wide building blocks: exclusive scan and stable LSD radix sort.
// Hand-written (no CUB/Thrust): both are HBM-bound integer passes; tiles are 2048 items per CTA and
// grids are plain ceil-div (far more CTAs than SMs at the sizes that matter).
#include <stdarg.h>

#include "common.cuh"

namespace twowl {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ------------------------------------------------------------------------------------------------
// exclusive scan, three passes: tile sums -> spine (single CTA) -> downsweep
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = CTA sum
template <typename T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* THREADS/32 + 1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = warp_inclusive_scan(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < THREADS / 32) ? smem[lane] : T(0);
    T winc = warp_inclusive_scan(w, lane);
    if (lane < THREADS / 32) smem[lane] = winc - w;
    if (lane == THREADS / 32 - 1) smem[THREADS / 32] = winc;
  }
  __syncthreads();
  T res = smem[warp] + inc - v;
  *total = smem[THREADS / 32];
  __syncthreads();
  return res;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const T* __restrict__ in, int64_t n,
                                                                 T* __restrict__ tile_sums) {
  __shared__ T sm[kScanThreads / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  T s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  T total;
  block_exclusive_scan<T, kScanThreads>(s, &total, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_spine(T* __restrict__ tile_sums, int64_t ntiles) {
  __shared__ T sm[1024 / 32 + 1];
  const int64_t per = (ntiles + 1023) / 1024;
  const int64_t b = (int64_t)threadIdx.x * per;
  const int64_t e = b + per < ntiles ? b + per : ntiles;
  T s = 0;
  for (int64_t i = b; i < e; ++i) s += tile_sums[i];
  T total;
  T run = block_exclusive_scan<T, 1024>(s, &total, sm);
  for (int64_t i = b; i < e; ++i) {
    T v = tile_sums[i];
    tile_sums[i] = run;
    run += v;
  }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_down(const T* in, T* out, int64_t n,
                                                            const T* __restrict__ tile_sums) {
  __shared__ T sm[kScanThreads / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  T v[kScanItems];
  T s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : T(0);
    s += v[i];
  }
  T total;
  T run = block_exclusive_scan<T, kScanThreads>(s, &total, sm) + tile_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) {
      out[base + i] = run;
      run += v[i];
      if (base + i == n - 1) out[n] = run;
    }
  }
}

template <typename T>
static size_t scan_ws_bytes_t(int64_t n) {
  return align_up((size_t)(cdiv(n, kScanTile) + 1) * sizeof(T));
}

template <typename T>
static int scan_exclusive_t(const T* in, T* out, int64_t n, void* ws, cudaStream_t s) {
  if (n <= 0) {
    TW_CUDA(cudaMemsetAsync(out, 0, sizeof(T), s));
    return 0;
  }
  const int64_t ntiles = cdiv(n, kScanTile);
  T* tile_sums = static_cast<T*>(ws);
  k_scan_tile_sums<T><<<(unsigned)ntiles, kScanThreads, 0, s>>>(in, n, tile_sums);
  k_scan_spine<T><<<1, 1024, 0, s>>>(tile_sums, ntiles);
  k_scan_down<T><<<(unsigned)ntiles, kScanThreads, 0, s>>>(in, out, n, tile_sums);
  TW_LAUNCH_CHECK();
  return 0;
}

size_t scan_workspace_bytes(int64_t n) { return scan_ws_bytes_t<int64_t>(n); }
int scan_exclusive_i64(const int64_t* in, int64_t* out, int64_t n, void* ws, cudaStream_t s) {
  return scan_exclusive_t<int64_t>(in, out, n, ws, s);
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits. Per pass: tile histogram -> scan of the digit-major table ->
// stable scatter (ballot/match-based in-warp ranking, warps own consecutive runs of the tile).
// ------------------------------------------------------------------------------------------------
constexpr int kRadixThreads = 256;
constexpr int kRadixItems = 8;
constexpr int kRadixTile = kRadixThreads * kRadixItems;
constexpr int kRadixWarps = kRadixThreads / 32;

__global__ void __launch_bounds__(kRadixThreads) k_radix_hist(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                                              uint32_t* __restrict__ hist, int64_t ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRadixTile;
#pragma unroll
  for (int i = 0; i < kRadixItems; ++i) {
    const int64_t p = base + i * kRadixThreads + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kRadixThreads) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ keys_out,
                                                                 uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                 const uint32_t* __restrict__ hist, int64_t ntiles) {
  __shared__ uint32_t wh[kRadixWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRadixWarps * 256; i += kRadixThreads) (&wh[0][0])[i] = 0;
  __syncthreads();
  // warp w owns items [w*32*ITEMS, (w+1)*32*ITEMS) of the tile; round r covers 32 consecutive items
  const int64_t wbase = (int64_t)blockIdx.x * kRadixTile + (int64_t)warp * 32 * kRadixItems;
  uint32_t key[kRadixItems], val[kRadixItems], rank[kRadixItems];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRadixItems; ++r) {
    const int64_t p = wbase + r * 32 + lane;
    const bool ok = p < n;
    key[r] = ok ? keys_in[p] : 0u;
    val[r] = ok ? vals_in[p] : 0u;
    // digit 256 is a private class for out-of-range lanes so that match_any groups them apart
    const uint32_t d = ok ? ((key[r] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = 0;
    if (ok) before = wh[warp][d];
    __syncwarp();
    rank[r] = before + __popc(peers & lt);
    if (ok && (peers & lt) == 0) wh[warp][d] = before + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // 256 threads <-> 256 digits
    uint32_t run = hist[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRadixWarps; ++w) {
      const uint32_t c = wh[w][d];
      wh[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRadixItems; ++r) {
    const int64_t p = wbase + r * 32 + lane;
    if (p < n) {
      const uint32_t pos = wh[warp][(key[r] >> shift) & 255u] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

size_t radix_workspace_bytes(int64_t n) {
  const int64_t ntiles = cdiv(n > 0 ? n : 1, kRadixTile);
  return align_up((size_t)(256 * ntiles + 1) * sizeof(uint32_t)) + scan_ws_bytes_t<uint32_t>(256 * ntiles);
}

int radix_sort_pairs(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n, int bits,
                     void* ws, cudaStream_t s) {
  if (n <= 0) return 0;
  const int64_t ntiles = cdiv(n, kRadixTile);
  Carver c(ws);
  uint32_t* hist = c.take<uint32_t>(256 * ntiles + 1);
  void* scan_ws = c.take<char>(scan_ws_bytes_t<uint32_t>(256 * ntiles));
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  uint32_t *ki = keys_in, *vi = vals_in, *ko = keys_out, *vo = vals_out;
  for (int p = 0; p < passes; ++p) {
    k_radix_hist<<<(unsigned)ntiles, kRadixThreads, 0, s>>>(ki, n, p * 8, hist, ntiles);
    int rc = scan_exclusive_t<uint32_t>(hist, hist, 256 * ntiles, scan_ws, s);
    if (rc) return rc;
    k_radix_scatter<<<(unsigned)ntiles, kRadixThreads, 0, s>>>(ki, vi, ko, vo, n, p * 8, hist, ntiles);
    TW_LAUNCH_CHECK();
    uint32_t* t;
    t = ki, ki = ko, ko = t;
    t = vi, vi = vo, vo = t;
  }
  // after the loop the result lives in (ki, vi); make sure it ends in the caller's *_out buffers
  if (ki != keys_out) {
    TW_CUDA(cudaMemcpyAsync(keys_out, ki, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    TW_CUDA(cudaMemcpyAsync(vals_out, vi, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

// ---------------------------------------------------------------- TMA tensor maps ----------------
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tmap_encode_fn>(sym);
  }
  return fn;
}
int make_tmap_2d_f32(CUtensorMap* tm, const float* base, int64_t rows, int cols, int box_cols, int box_rows,
                     CUtensorMapSwizzle swizzle, int64_t ld) {
  tmap_encode_fn enc = tmap_encoder();
  TW_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)(ld > 0 ? ld : cols) * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TW_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace twowl

extern "C" int twowl_version(void) { return 100; }
extern "C" const char* twowl_last_error(void) { return twowl::get_error(); }
