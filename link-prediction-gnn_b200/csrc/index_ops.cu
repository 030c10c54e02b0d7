// index_ops.cu - the integer graph operators of TwoWL/utils.py:8-90 as sm_100a kernels.
// All of them are bit-exact against the reference: integer arithmetic, order-preserving layouts
// obtained from prefix sums (never from atomics on positions).
#include "common.cuh"

namespace twowl {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- degree / histogram -----------
__global__ void __launch_bounds__(kThreads) k_hist_i64(const int64_t* __restrict__ keys, int64_t stride, int64_t n,
                                                       int64_t key_xor, int64_t num, unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = -1;
    if (i < n) k = keys[i * stride] ^ key_xor;
    const bool ok = (i < n) && k >= 0 && k < num;
    // warp-aggregated increment: one atomic per distinct key per warp (hubs of power-law graphs)
    const unsigned long long kk = ok ? (unsigned long long)k : ~0ull;
    const unsigned peers = __match_any_sync(0xffffffffu, kk);
    if (ok && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&out[k], (unsigned long long)__popc(peers));
  }
}

// ---------------------------------------------------------------- csr build ----------------------
__global__ void __launch_bounds__(kThreads) k_csr_keys(const int64_t* __restrict__ keys, int64_t stride, int64_t n,
                                                       int64_t key_xor, int64_t num_keys, uint32_t* __restrict__ k32,
                                                       uint32_t* __restrict__ v32) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = keys[i * stride] ^ key_xor;
    k32[i] = (k >= 0 && k < num_keys) ? (uint32_t)k : (uint32_t)num_keys;  // sentinel bucket sorts last
    v32[i] = (uint32_t)i;
  }
}

static int bits_for(int64_t max_value) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) <= max_value) ++b;
  return b;
}

__global__ void __launch_bounds__(kThreads) k_gather_cols(const int32_t* __restrict__ ids, int64_t n,
                                                          const int64_t* __restrict__ vals, int64_t stride,
                                                          int64_t val_xor, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)(vals[(int64_t)ids[i] * stride] ^ val_xor);
}

// ---------------------------------------------------------------- get_ei2 ------------------------
__global__ void __launch_bounds__(kThreads) k_ei2_seg(const int64_t* __restrict__ in_ptr, const int64_t* __restrict__ out_ptr,
                                                      int64_t n_node, int64_t* __restrict__ seg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_node; i += (int64_t)gridDim.x * blockDim.x)
    seg[i] = (in_ptr[i + 1] - in_ptr[i]) * (out_ptr[i + 1] - out_ptr[i]);
}

// largest i in [lo, hi] with off[i] <= t   (off is non-decreasing; zero-length segments are skipped
// because the LAST index with off[i] <= t is the one whose segment contains t)
__device__ __forceinline__ int64_t seg_search(const int64_t* off, int64_t lo, int64_t hi, int64_t t) {
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (off[mid] <= t)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

constexpr int kFillTile = 2048;   // wedges per CTA
constexpr int kFillCache = 2048;  // segment offsets staged in shared memory per CTA

// ROWS = false: interleaved (a, b) pairs, the [T,2] buffer get_ei2 returns transposed; ROWS = true: two contiguous rows
// out_ab[0 .. n) = a, out_ab[n .. 2n) = b (n = t_end - t_begin), the fresh [2,T'] tensor blockei2's boolean indexing returns.
template <bool ROWS>
__global__ void __launch_bounds__(kThreads) k_ei2_fill(const int64_t* __restrict__ in_ptr, const int32_t* __restrict__ in_ids,
                                                       const int64_t* __restrict__ out_ptr, const int32_t* __restrict__ out_ids,
                                                       const int64_t* __restrict__ off, int64_t n_node, int64_t t_begin,
                                                       int64_t t_end, longlong2* __restrict__ out_ab) {
  __shared__ int64_t s_off[kFillCache + 1];
  __shared__ int64_t s_range[2];
  const int64_t t0 = t_begin + (int64_t)blockIdx.x * kFillTile;
  const int64_t t1 = (t0 + kFillTile < t_end) ? t0 + kFillTile : t_end;
  if (threadIdx.x == 0) s_range[0] = seg_search(off, 0, n_node - 1, t0);
  if (threadIdx.x == 32) s_range[1] = seg_search(off, 0, n_node - 1, t1 - 1);
  __syncthreads();
  const int64_t n_lo = s_range[0], n_hi = s_range[1];
  const bool cached = (n_hi - n_lo + 1) <= kFillCache;
  if (cached)
    for (int64_t i = threadIdx.x; i <= n_hi - n_lo + 1; i += blockDim.x) s_off[i] = off[n_lo + i];
  __syncthreads();
  // a tile inside ONE node's segment (hubs: cin*cout wedges >> tile) needs no search and its divisor is block-uniform
  const bool one = n_lo == n_hi;
  int64_t node1 = n_lo, base1 = 0, ob1 = 0, cout1 = 1, ib1 = 0;
  if (one) {
    base1 = cached ? s_off[0] : off[n_lo];
    ob1 = out_ptr[n_lo];
    cout1 = out_ptr[n_lo + 1] - ob1;
    ib1 = in_ptr[n_lo];
  }
  auto emit = [&](int64_t t, int32_t a, int32_t b) {
    if (ROWS) {
      int64_t* __restrict__ o = reinterpret_cast<int64_t*>(out_ab);
      __stcs(o + (t - t_begin), (long long)a);
      __stcs(o + (t_end - t_begin) + (t - t_begin), (long long)b);
    } else {
      __stcs(out_ab + (t - t_begin), make_longlong2((long long)a, (long long)b));  // one 128-bit streaming store per wedge
    }
  };
  if (one) {
    // (ia, ib) = divmod(t - base, cout) once per thread, then stepped by the block stride: the kernel was issue-bound (80 % of the
    // issue slots, ncu) on a divide per wedge. cout < 2^31 (pair rows), ia < cin < 2^31.
    const uint32_t cout = (uint32_t)cout1;
    const uint32_t qs = (uint32_t)blockDim.x / cout, rs = (uint32_t)blockDim.x % cout;
    int64_t t = t0 + threadIdx.x;
    if (t < t1) {
      const int64_t local = t - base1;
      uint32_t ia = (uint32_t)(local / (int64_t)cout), ib = (uint32_t)(local - (int64_t)ia * (int64_t)cout);
      const int32_t* __restrict__ ina = in_ids + ib1;
      const int32_t* __restrict__ outb = out_ids + ob1;
      for (; t < t1; t += blockDim.x) {
        emit(t, __ldg(ina + ia), __ldg(outb + ib));
        ia += qs, ib += rs;
        if (ib >= cout) ib -= cout, ++ia;
      }
    }
    return;
  }
  for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
    int64_t node, base;
    if (cached) {
      const int64_t j = seg_search(s_off, 0, n_hi - n_lo, t);
      node = n_lo + j;
      base = s_off[j];
    } else {
      node = seg_search(off, n_lo, n_hi, t);
      base = off[node];
    }
    const int64_t ob = out_ptr[node];
    const int64_t cout = out_ptr[node + 1] - ob;
    const int64_t local = t - base;
    int64_t ia, ib;
    if ((uint64_t)local < 0x100000000ull && (uint64_t)cout < 0x100000000ull) {   // 32-bit divide: ~10x cheaper than the 64-bit one
      const uint32_t q = (uint32_t)local / (uint32_t)cout;
      ia = q, ib = (int64_t)((uint32_t)local - q * (uint32_t)cout);
    } else {
      ia = local / cout, ib = local - ia * cout;
    }
    emit(t, __ldg(in_ids + in_ptr[node] + ia), __ldg(out_ids + ob + ib));
  }
}

// ---------------------------------------------------------------- masks / selection --------------
__global__ void __launch_bounds__(kThreads) k_mask_scatter(const int64_t* __restrict__ idx, int64_t k, uint8_t* __restrict__ mask,
                                                           int64_t num) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = idx[i];
    if (v < 0) v += num;  // python-style negative index, as `mask[idx] = True` would accept
    if (v >= 0 && v < num) mask[v] = 1;
  }
}

// x[idx] semantics of the reference's advanced indexing (model.py:78, utils.py:55): an id in [-num, 0) wraps, anything else
// outside [0, num) is an error. out[i] = the wrapped id (0 for an invalid one, so that no consumer can read out of bounds) and
// bad[0] counts the invalid ids - the caller turns a non-zero count into a device-side assertion / an IndexError.
__global__ void __launch_bounds__(kThreads) k_index_guard(const int64_t* __restrict__ idx, int64_t stride, int64_t k, int64_t num,
                                                          int64_t* __restrict__ out, int32_t* __restrict__ bad) {
  int nbad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = idx[i * stride];
    if (v < 0) v += num;
    const bool ok = v >= 0 && v < num;
    nbad += !ok;
    out[i] = ok ? v : 0;
  }
  nbad = __reduce_add_sync(0xffffffffu, nbad);
  if ((threadIdx.x & 31) == 0 && nbad) atomicAdd(bad, nbad);
}

// tile = 2048 columns = 8 rounds of 256 threads (round r, thread i -> column base + r*256 + i: coalesced). All rounds' keys are
// requested before any mask byte is looked up, and the order-preserving ranks come from ONE block-wide prefix over the 8 x 8
// (round, warp) ballot counts - two barriers per tile instead of three per round.
constexpr int kSelRounds = TWOWL_SELECT_TILE / kThreads;
constexpr int kSelWarps = kThreads / 32;

__device__ __forceinline__ void select_flags(const int64_t* __restrict__ row0, int64_t s0, int64_t T, const uint8_t* __restrict__ mask,
                                             int64_t mask_len, int mode, int64_t base, bool (&keep)[kSelRounds]) {
  int64_t key[kSelRounds];
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r) {
    const int64_t t = base + r * kThreads + threadIdx.x;
    key[r] = t < T ? (mode == 0 ? t : __ldg(row0 + t * s0)) : -1;
  }
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r)
    keep[r] = (base + r * kThreads + threadIdx.x < T) && (key[r] < 0 || key[r] >= mask_len || __ldg(mask + key[r]) == 0);
}

__global__ void __launch_bounds__(kThreads) k_select_count(const int64_t* __restrict__ row0, int64_t s0, int64_t T,
                                                           const uint8_t* __restrict__ mask, int64_t mask_len, int mode,
                                                           int64_t* __restrict__ tile_cnt) {
  __shared__ int s_cnt[kSelWarps];
  const int64_t base = (int64_t)blockIdx.x * TWOWL_SELECT_TILE;
  bool keep[kSelRounds];
  select_flags(row0, s0, T, mask, mask_len, mode, base, keep);
  int c = 0;
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r) c += __popc(__ballot_sync(0xffffffffu, keep[r]));
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kSelWarps; ++w) tot += s_cnt[w];
    tile_cnt[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(kThreads) k_select_fill(const int64_t* __restrict__ row0, int64_t s0,
                                                          const int64_t* __restrict__ row1, int64_t s1, int64_t T,
                                                          const uint8_t* __restrict__ mask, int64_t mask_len, int mode,
                                                          const int64_t* __restrict__ tile_off, int64_t* __restrict__ out0,
                                                          int64_t* __restrict__ out1) {
  __shared__ int s_cnt[kSelRounds * kSelWarps];   // [round][warp]: the tile order is round-major, then warp, then lane
  __shared__ int s_pre[kSelRounds * kSelWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * TWOWL_SELECT_TILE;
  const int64_t obase = tile_off[blockIdx.x];
  bool keep[kSelRounds];
  select_flags(row0, s0, T, mask, mask_len, mode, base, keep);
  // the payload loads do not depend on the ranks: request them now (row0 again: an L1/L2 hit of the key load above)
  int64_t v0[kSelRounds], v1[kSelRounds];
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r) {
    const int64_t t = base + r * kThreads + threadIdx.x;
    v0[r] = keep[r] ? row0[t * s0] : 0;
    v1[r] = keep[r] ? row1[t * s1] : 0;
  }
  unsigned bal[kSelRounds];
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r) {
    bal[r] = __ballot_sync(0xffffffffu, keep[r]);
    if (lane == 0) s_cnt[r * kSelWarps + warp] = __popc(bal[r]);
  }
  __syncthreads();
  if (threadIdx.x < kSelRounds * kSelWarps) {      // exclusive prefix of the 64 counts by one thread each (64 adds at most)
    int pre = 0;
    for (int i = 0; i < (int)threadIdx.x; ++i) pre += s_cnt[i];
    s_pre[threadIdx.x] = pre;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSelRounds; ++r) {
    if (keep[r]) {
      const int64_t o = obase + s_pre[r * kSelWarps + warp] + __popc(bal[r] & ((1u << lane) - 1u));
      __stcs(out0 + o, v0[r]);
      __stcs(out1 + o, v1[r]);
    }
  }
}

// ---------------------------------------------------------------- check_in_set -------------------
__global__ void __launch_bounds__(kThreads) k_set_hist(const int64_t* __restrict__ set, int64_t m, int64_t range,
                                                       int32_t* __restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = set[i];
    if (v >= 0 && v < range) atomicAdd(&counts[v], 1);
  }
}
__global__ void __launch_bounds__(kThreads) k_set_lookup(const int64_t* __restrict__ target, int64_t st, int64_t n,
                                                         int64_t range, const int32_t* __restrict__ counts,
                                                         int64_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = target[i * st];
    out[i] = (v >= 0 && v < range) ? (int64_t)counts[v] : 0;
  }
}

// ---------------------------------------------------------------- elementwise index ops ----------
__global__ void __launch_bounds__(kThreads) k_reverse(const int64_t* __restrict__ row0, int64_t s0,
                                                      const int64_t* __restrict__ row1, int64_t s1, int64_t T,
                                                      int64_t* __restrict__ edge, int64_t* __restrict__ edge_r) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = row0[t * s0], b = row1[t * s1];
    edge[t] = a ^ 1;  // utils.py:72-75: +1 for even ids, -1 for odd ids
    edge[T + t] = b;
    edge_r[t] = a;
    edge_r[T + t] = b ^ 1;
  }
}

// Two wedges per thread with 128-bit accesses. PAIRED: ei2 is the transposed view of a [T,2] buffer (what get_ei2 returns,
// utils.py:45), one longlong2 = (a, b) of a wedge; otherwise two contiguous rows. Needs T even and 16-byte aligned pointers.
template <bool PAIRED>
__global__ void __launch_bounds__(kThreads) k_reverse_v2(const int64_t* __restrict__ row0, const int64_t* __restrict__ row1, int64_t T,
                                                         int64_t* __restrict__ edge, int64_t* __restrict__ edge_r) {
  const int64_t half = T >> 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (int64_t)gridDim.x * blockDim.x) {
    longlong2 a, b;
    if (PAIRED) {
      const longlong2 w0 = __ldcs(reinterpret_cast<const longlong2*>(row0) + 2 * i);
      const longlong2 w1 = __ldcs(reinterpret_cast<const longlong2*>(row0) + 2 * i + 1);
      a = make_longlong2(w0.x, w1.x), b = make_longlong2(w0.y, w1.y);
    } else {
      a = __ldcs(reinterpret_cast<const longlong2*>(row0) + i);
      b = __ldcs(reinterpret_cast<const longlong2*>(row1) + i);
    }
    __stcs(reinterpret_cast<longlong2*>(edge) + i, make_longlong2(a.x ^ 1, a.y ^ 1));   // utils.py:72-75
    __stcs(reinterpret_cast<longlong2*>(edge + T) + i, b);
    __stcs(reinterpret_cast<longlong2*>(edge_r) + i, a);
    __stcs(reinterpret_cast<longlong2*>(edge_r + T) + i, make_longlong2(b.x ^ 1, b.y ^ 1));
  }
}

__global__ void __launch_bounds__(kThreads) k_double_edges(const int64_t* __restrict__ r, int64_t sr,
                                                           const int64_t* __restrict__ c, int64_t sc, int64_t M,
                                                           int64_t* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = r[k * sr], b = c[k * sc];
    out[2 * k] = a;
    out[2 * k + 1] = b;
    out[2 * M + 2 * k] = b;
    out[2 * M + 2 * k + 1] = a;
  }
}

__global__ void __launch_bounds__(kThreads) k_double_index(const int64_t* __restrict__ x, int64_t sx, int64_t B,
                                                           int64_t* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < B; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = x[k * sx];
    out[2 * k] = 2 * v;
    out[2 * k + 1] = 2 * v + 1;
  }
}

__global__ void __launch_bounds__(kThreads) k_set_mul(const int64_t* __restrict__ a, int64_t p, const int64_t* __restrict__ b,
                                                      int64_t q, longlong2* __restrict__ out) {
  const int64_t n = p * q;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    longlong2 v;
    v.x = a[t / q];
    v.y = b[t % q];
    out[t] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_narrow(const int64_t* __restrict__ in, int64_t stride, int64_t n,
                                                     int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)in[i * stride];
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_degree(const int64_t* keys, int64_t stride, int64_t n, int64_t num_node, int64_t* out, void* stream) {
  TW_CHECK_ARG(n >= 0 && num_node >= 0, "degree: negative size");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_node > 0) TW_CUDA(cudaMemsetAsync(out, 0, (size_t)num_node * sizeof(int64_t), s));
  if (n > 0 && num_node > 0) {
    k_hist_i64<<<grid_for(n, kThreads), kThreads, 0, s>>>(keys, stride, n, 0, num_node, (unsigned long long*)out);
    TW_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" size_t twowl_csr_build_workspace_bytes(int64_t n, int64_t num_keys) {
  const size_t nn = (size_t)(n > 0 ? n : 1);
  return 4 * align_up(nn * sizeof(uint32_t)) + radix_workspace_bytes(n) + scan_workspace_bytes(num_keys + 1) + 256;
}

extern "C" int twowl_csr_build(const int64_t* keys, int64_t stride, int64_t n, int64_t key_xor, int64_t num_keys,
                               int64_t* ptr, int32_t* ids, void* ws, size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(n >= 0 && n < 0x7fffffffLL, "csr_build: n=%lld out of int32 range", (long long)n);
  TW_CHECK_ARG(num_keys >= 0 && num_keys < 0x7fffffffLL, "csr_build: num_keys=%lld out of range", (long long)num_keys);
  TW_CHECK_WS(ws_bytes, twowl_csr_build_workspace_bytes(n, num_keys));
  cudaStream_t s = (cudaStream_t)stream;
  if (ptr) TW_CUDA(cudaMemsetAsync(ptr, 0, (size_t)(num_keys + 1) * sizeof(int64_t), s));
  if (n == 0) return 0;
  Carver c(ws);
  const size_t nn = (size_t)n;
  uint32_t* k0 = c.take<uint32_t>(nn);
  uint32_t* v0 = c.take<uint32_t>(nn);
  uint32_t* k1 = c.take<uint32_t>(nn);
  uint32_t* v1 = c.take<uint32_t>(nn);
  void* radix_ws = c.take<char>(radix_workspace_bytes(n));
  void* scan_ws = c.take<char>(scan_workspace_bytes(num_keys + 1));
  k_csr_keys<<<grid_for(n, kThreads), kThreads, 0, s>>>(keys, stride, n, key_xor, num_keys, k0, v0);
  // histogram of the in-range keys into ptr[0..num_keys), then exclusive scan in place -> ptr[num_keys] = total
  if (ptr && num_keys > 0) {
    k_hist_i64<<<grid_for(n, kThreads), kThreads, 0, s>>>(keys, stride, n, key_xor, num_keys, (unsigned long long*)ptr);
    int rc = scan_exclusive_i64(ptr, ptr, num_keys, scan_ws, s);
    if (rc) return rc;
  }
  TW_LAUNCH_CHECK();
  int rc = radix_sort_pairs(k0, v0, k1, v1, n, bits_for(num_keys), radix_ws, s);
  if (rc) return rc;
  TW_CUDA(cudaMemcpyAsync(ids, v1, nn * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int twowl_gather_cols(const int32_t* ids, int64_t n, const int64_t* vals, int64_t stride, int64_t val_xor,
                                 int32_t* out, void* stream) {
  if (n <= 0) return 0;
  k_gather_cols<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(ids, n, vals, stride, val_xor, out);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t twowl_ei2_count_workspace_bytes(int64_t n_node) { return scan_workspace_bytes(n_node) + 256; }

extern "C" int twowl_ei2_count(const int64_t* in_ptr, const int64_t* out_ptr, int64_t n_node, int64_t* off, void* ws,
                               size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(n_node >= 0, "ei2_count: negative n_node");
  TW_CHECK_WS(ws_bytes, twowl_ei2_count_workspace_bytes(n_node));
  cudaStream_t s = (cudaStream_t)stream;
  if (n_node > 0) {
    k_ei2_seg<<<grid_for(n_node, kThreads), kThreads, 0, s>>>(in_ptr, out_ptr, n_node, off);
    TW_LAUNCH_CHECK();
  }
  return scan_exclusive_i64(off, off, n_node, ws, s);
}

extern "C" int twowl_ei2_fill(const int64_t* in_ptr, const int32_t* in_ids, const int64_t* out_ptr, const int32_t* out_ids,
                              const int64_t* off, int64_t n_node, int64_t t_begin, int64_t t_end, int64_t* out_ab,
                              void* stream) {
  TW_CHECK_ARG(t_begin >= 0 && t_end >= t_begin, "ei2_fill: bad range [%lld,%lld)", (long long)t_begin, (long long)t_end);
  if (t_end == t_begin || n_node == 0) return 0;
  TW_CHECK_ARG(aligned16(out_ab), "ei2_fill: out_ab must be 16-byte aligned");
  const int64_t tiles = cdiv(t_end - t_begin, kFillTile);
  TW_CHECK_ARG(tiles < 0x7fffffffLL, "ei2_fill: range too large for one launch; split [t_begin,t_end)");
  k_ei2_fill<false><<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(in_ptr, in_ids, out_ptr, out_ids, off, n_node, t_begin,
                                                                           t_end, (longlong2*)out_ab);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_ei2_fill_rows(const int64_t* in_ptr, const int32_t* in_ids, const int64_t* out_ptr, const int32_t* out_ids,
                                   const int64_t* off, int64_t n_node, int64_t t_begin, int64_t t_end, int64_t* out_rows,
                                   void* stream) {
  TW_CHECK_ARG(t_begin >= 0 && t_end >= t_begin, "ei2_fill_rows: bad range [%lld,%lld)", (long long)t_begin, (long long)t_end);
  if (t_end == t_begin || n_node == 0) return 0;
  const int64_t tiles = cdiv(t_end - t_begin, kFillTile);
  TW_CHECK_ARG(tiles < 0x7fffffffLL, "ei2_fill_rows: range too large for one launch; split [t_begin,t_end)");
  k_ei2_fill<true><<<(unsigned)tiles, kThreads, 0, (cudaStream_t)stream>>>(in_ptr, in_ids, out_ptr, out_ids, off, n_node, t_begin,
                                                                          t_end, (longlong2*)out_rows);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_mask_from_idx(const int64_t* idx, int64_t k, uint8_t* mask, int64_t num, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (num > 0) TW_CUDA(cudaMemsetAsync(mask, 0, (size_t)num, s));
  if (k > 0 && num > 0) {
    k_mask_scatter<<<grid_for(k, kThreads), kThreads, 0, s>>>(idx, k, mask, num);
    TW_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int twowl_index_guard(const int64_t* idx, int64_t stride, int64_t k, int64_t num, int64_t* out, int32_t* bad,
                                 void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  TW_CHECK_ARG(bad != nullptr && (k == 0 || (idx && out)), "index_guard: null argument");
  TW_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), s));
  if (k > 0) {
    k_index_guard<<<grid_for(k, kThreads), kThreads, 0, s>>>(idx, stride, k, num, out, bad);
    TW_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" size_t twowl_select_workspace_bytes(int64_t T) {
  return scan_workspace_bytes(cdiv(T > 0 ? T : 1, TWOWL_SELECT_TILE)) + 256;
}

extern "C" int twowl_select_count(const int64_t* row0, int64_t s0, int64_t T, const uint8_t* mask, int64_t mask_len, int mode,
                                  int64_t* tile_off, void* ws, size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(T >= 0 && (mode == 0 || mode == 1), "select_count: bad arguments");
  TW_CHECK_WS(ws_bytes, twowl_select_workspace_bytes(T));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t ntiles = cdiv(T, TWOWL_SELECT_TILE);
  if (ntiles > 0) {
    k_select_count<<<(unsigned)ntiles, kThreads, 0, s>>>(row0, s0, T, mask, mask_len, mode, tile_off);
    TW_LAUNCH_CHECK();
  }
  return scan_exclusive_i64(tile_off, tile_off, ntiles, ws, s);
}

extern "C" int twowl_select_fill(const int64_t* row0, int64_t s0, const int64_t* row1, int64_t s1, int64_t T,
                                 const uint8_t* mask, int64_t mask_len, int mode, const int64_t* tile_off, int64_t* out0,
                                 int64_t* out1, void* stream) {
  const int64_t ntiles = cdiv(T, TWOWL_SELECT_TILE);
  if (ntiles <= 0) return 0;
  k_select_fill<<<(unsigned)ntiles, kThreads, 0, (cudaStream_t)stream>>>(row0, s0, row1, s1, T, mask, mask_len, mode,
                                                                       tile_off, out0, out1);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_check_in_set(const int64_t* target, int64_t st, int64_t n, const int64_t* set, int64_t m, int64_t range,
                                  int32_t* counts_ws, int64_t* out, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (range > 0) TW_CUDA(cudaMemsetAsync(counts_ws, 0, (size_t)range * sizeof(int32_t), s));
  if (m > 0 && range > 0) k_set_hist<<<grid_for(m, kThreads), kThreads, 0, s>>>(set, m, range, counts_ws);
  if (n > 0) k_set_lookup<<<grid_for(n, kThreads), kThreads, 0, s>>>(target, st, n, range, counts_ws, out);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_reverse(const int64_t* row0, int64_t s0, const int64_t* row1, int64_t s1, int64_t T, int64_t* edge,
                             int64_t* edge_r, void* stream) {
  if (T <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = (T & 1) == 0 && aligned16(row0) && aligned16(edge) && aligned16(edge_r);
  if (vec && s0 == 2 && s1 == 2 && row1 == row0 + 1) {
    k_reverse_v2<true><<<grid_for(T / 2, kThreads), kThreads, 0, s>>>(row0, row1, T, edge, edge_r);
  } else if (vec && s0 == 1 && s1 == 1 && aligned16(row1)) {
    k_reverse_v2<false><<<grid_for(T / 2, kThreads), kThreads, 0, s>>>(row0, row1, T, edge, edge_r);
  } else
    k_reverse<<<grid_for(T, kThreads), kThreads, 0, s>>>(row0, s0, row1, s1, T, edge, edge_r);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_double_edges(const int64_t* r, int64_t sr, const int64_t* c, int64_t sc, int64_t M, int64_t* out,
                                  void* stream) {
  if (M <= 0) return 0;
  k_double_edges<<<grid_for(M, kThreads), kThreads, 0, (cudaStream_t)stream>>>(r, sr, c, sc, M, out);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_double_index(const int64_t* x, int64_t sx, int64_t B, int64_t* out, void* stream) {
  if (B <= 0) return 0;
  k_double_index<<<grid_for(B, kThreads), kThreads, 0, (cudaStream_t)stream>>>(x, sx, B, out);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_set_mul(const int64_t* a, int64_t p, const int64_t* b, int64_t q, int64_t* out, void* stream) {
  if (p <= 0 || q <= 0) return 0;
  TW_CHECK_ARG(aligned16(out), "set_mul: out must be 16-byte aligned");
  k_set_mul<<<grid_for(p * q, kThreads), kThreads, 0, (cudaStream_t)stream>>>(a, p, b, q, (longlong2*)out);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_narrow_i32(const int64_t* in, int64_t stride, int64_t n, int32_t* out, void* stream) {
  if (n <= 0) return 0;
  k_narrow<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(in, stride, n, out);
  TW_LAUNCH_CHECK();
  return 0;
}
