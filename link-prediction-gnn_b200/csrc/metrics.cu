// metrics.cu - ROC-AUC on the device, replacing the per-step host round trip of the reference's train / test loop
// (TwoWL/model/train.py:41-43 and :61-66: pred.sigmoid().cpu().numpy() -> sklearn.metrics.roc_auc_score).
//
// AUC = (sum of the average ranks of the positives - n_pos (n_pos + 1) / 2) / (n_pos n_neg)   (Mann-Whitney U with ties
// sharing their average rank = the area under the trapezoidal ROC curve sklearn integrates). Scores are sorted with the
// library's stable LSD radix sort on an order-preserving 32-bit key; each element finds its tie group by binary search
// in the sorted keys; sums are integer / double with a fixed two-level order: deterministic, no atomics.
#include "common.cuh"

namespace twowl {

constexpr int kAucThreads = 256;
constexpr int kAucMaxCtas = kNumSMs * 4;

__device__ __forceinline__ uint32_t auc_key(float x) {
  // ascending order-preserving map of IEEE-754 floats to uint32 (-0 and +0 share a key; NaNs sort last)
  if (x == 0.f) x = 0.f;
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kAucThreads) k_auc_keys(const float* __restrict__ score, int64_t n, uint32_t* __restrict__ keys,
                                                          uint32_t* __restrict__ vals) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = auc_key(score[i]);
    vals[i] = (uint32_t)i;
  }
}

// per CTA: (sum over positives of 2 * average rank, number of positives) as (double, int64)
__global__ void __launch_bounds__(kAucThreads) k_auc_ranks(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                           const float* __restrict__ label, int64_t n, double* __restrict__ part) {
  __shared__ double s_r[kAucThreads];
  __shared__ double s_p[kAucThreads];
  double r2 = 0, np = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    if (!(label[vals[j]] > 0.5f)) continue;
    const uint32_t k = keys[j];
    int64_t lo = 0, hi = j;          // first position with key == k
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    const int64_t s = lo;
    lo = j + 1, hi = n;               // first position with key > k
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] <= k) lo = mid + 1; else hi = mid;
    }
    r2 += (double)(s + lo + 1);       // ranks s+1 .. lo: twice their average
    np += 1.0;
  }
  s_r[threadIdx.x] = r2, s_p[threadIdx.x] = np;
  __syncthreads();
  for (int d = kAucThreads / 2; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) s_r[threadIdx.x] += s_r[threadIdx.x + d], s_p[threadIdx.x] += s_p[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[2 * blockIdx.x] = s_r[0], part[2 * blockIdx.x + 1] = s_p[0];
}

__global__ void k_auc_final(const double* __restrict__ part, int nparts, int64_t n, double* __restrict__ out) {
  double r2 = 0, np = 0;
  for (int b = 0; b < nparts; ++b) r2 += part[2 * b], np += part[2 * b + 1];
  const double nn = (double)n - np;
  out[1] = np, out[2] = nn;
  out[0] = (np > 0 && nn > 0) ? (0.5 * r2 - 0.5 * np * (np + 1.0)) / (np * nn) : nan("");
}

}  // namespace twowl

using namespace twowl;

extern "C" size_t twowl_auc_workspace_bytes(int64_t n) {
  const size_t m = (size_t)(n > 0 ? n : 1);
  return 4 * align_up(m * sizeof(uint32_t)) + align_up(radix_workspace_bytes(n)) + align_up((size_t)kAucMaxCtas * 2 * sizeof(double));
}

extern "C" int twowl_auc(const float* score, const float* label, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(n >= 0 && n < 0x7fffffffLL, "auc: n out of range");
  TW_CHECK_ARG(score && label && out, "auc: null pointer");
  TW_CHECK_WS(ws_bytes, twowl_auc_workspace_bytes(n));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t m = (size_t)(n > 0 ? n : 1);
  Carver c(ws);
  uint32_t* k0 = c.take<uint32_t>(m);
  uint32_t* v0 = c.take<uint32_t>(m);
  uint32_t* k1 = c.take<uint32_t>(m);
  uint32_t* v1 = c.take<uint32_t>(m);
  void* rws = c.take<char>(radix_workspace_bytes(n));
  double* part = c.take<double>((size_t)kAucMaxCtas * 2);
  const int grid = n > 0 ? grid_for(n, kAucThreads, 4) : 1;
  if (n > 0) {
    k_auc_keys<<<grid, kAucThreads, 0, s>>>(score, n, k0, v0);
    TW_LAUNCH_CHECK();
    if (int rc = radix_sort_pairs(k0, v0, k1, v1, n, 32, rws, s)) return rc;
  }
  k_auc_ranks<<<grid, kAucThreads, 0, s>>>(k1, v1, label, n, part);
  k_auc_final<<<1, 1, 0, s>>>(part, grid, n, out);
  TW_LAUNCH_CHECK();
  return 0;
}
