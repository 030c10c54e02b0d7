// linear.cu - the dense per-row linear layer inside GCNConv (PyG Linear, no bias; reference call site
// TwoWL/model/model.py:37 `GCNConv(insize, outsize)`): Z = X W^T, dX = dZ W, dW = dZ^T X, fp32 in / fp32 out.
//
// impl 0 (this file): SIMT FFMA register-tiled GEMM, exact fp32 accumulation - the parity baseline.
// The shapes are tall-skinny (M = rows of the pair table, up to 6e7; Ci, Co <= 256) so the op is HBM-bound
// up to C ~ 64 on FFMA and needs tensor cores (3xTF32 on tcgen05, the kernel of pair_conv.cu) beyond that.
#include "common.cuh"

namespace twowl {

constexpr int kGemmThreads = 256;
constexpr int BM = 128, BN = 64, BK = 16;  // CTA tile; thread tile 8 x 4

// C[M,N] = A[M,K] * B   with A row-major (K contiguous) and
//   B_KMAJOR = true : B given as [N,K] row-major (K contiguous)  -> C = A * B^T   (forward, B = W[Co,Ci])
//   B_KMAJOR = false: B given as [K,N] row-major (N contiguous)  -> C = A * B     (backward input, B = W[Co,Ci])
template <bool B_KMAJOR>
__global__ void __launch_bounds__(kGemmThreads) k_gemm_rows(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ Cm, int64_t M, int N, int K) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads; thread tile rows ty*8.., cols tx*4..
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK floats = 512 float4 (along K), 2 per thread, stored k-major
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * kGemmThreads;  // 0..511
      const int r = f / (BK / 4), kq = f % (BK / 4);
      float4 v = f4_zero();
      if (m0 + r < M && k0 + kq * 4 < K) v = ldg_stream(reinterpret_cast<const float4*>(A + (m0 + r) * K + k0 + kq * 4));
      As[kq * 4 + 0][r] = v.x, As[kq * 4 + 1][r] = v.y, As[kq * 4 + 2][r] = v.z, As[kq * 4 + 3][r] = v.w;
    }
    if (B_KMAJOR) {
      // B tile: BN x BK floats = 256 float4 along K, 1 per thread
      const int r = tid / (BK / 4), kq = tid % (BK / 4);
      float4 v = f4_zero();
      if (n0 + r < N && k0 + kq * 4 < K) v = ldg_cached(reinterpret_cast<const float4*>(B + (int64_t)(n0 + r) * K + k0 + kq * 4));
      Bs[kq * 4 + 0][r] = v.x, Bs[kq * 4 + 1][r] = v.y, Bs[kq * 4 + 2][r] = v.z, Bs[kq * 4 + 3][r] = v.w;
    } else {
      // B tile: BK x BN floats = 256 float4 along N, 1 per thread
      const int kk = tid / (BN / 4), nq = tid % (BN / 4);
      float4 v = f4_zero();
      if (k0 + kk < K && n0 + nq * 4 < N) v = ldg_cached(reinterpret_cast<const float4*>(B + (int64_t)(k0 + kk) * N + n0 + nq * 4));
      *reinterpret_cast<float4*>(&Bs[kk][nq * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty * 8 + i;
      if (m < M) stg_stream(reinterpret_cast<float4*>(Cm + m * N + n), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

// dW[Co,Ci] = sum_m dZ[m,co] * X[m,ci]: 64x64 output tile per CTA, split over M; each split writes its
// partial tile to part[split][Co][Ci]; k_dw_final adds the splits in order (deterministic).
constexpr int WT = 64, WK = 16;
__global__ void __launch_bounds__(kGemmThreads) k_dw_partial(const float* __restrict__ dZ, const float* __restrict__ rs,
                                                             const float* __restrict__ X, int64_t M, int Ci, int Co,
                                                             int64_t rows_per_split, float* __restrict__ part) {
  __shared__ __align__(16) float As[WK][WT];
  __shared__ __align__(16) float Bs[WK][WT];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int co0 = blockIdx.y * WT, ci0 = blockIdx.x * WT;
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = (mb + rows_per_split < M) ? mb + rows_per_split : M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid / (WT / 4), lq = tid % (WT / 4);  // tile loader: row lr (0..15), float4 column lq (0..15)
  for (int64_t m0 = mb; m0 < me; m0 += WK) {
    float4 a = f4_zero(), b = f4_zero();
    if (m0 + lr < me) {
      if (co0 + lq * 4 < Co) {
        a = ldg_stream(reinterpret_cast<const float4*>(dZ + (m0 + lr) * Co + co0 + lq * 4));
        if (rs) {
          const float sc = __ldg(rs + m0 + lr);
          a.x *= sc, a.y *= sc, a.z *= sc, a.w *= sc;
        }
      }
      if (ci0 + lq * 4 < Ci) b = ldg_stream(reinterpret_cast<const float4*>(X + (m0 + lr) * Ci + ci0 + lq * 4));
    }
    *reinterpret_cast<float4*>(&As[lr][lq * 4]) = a;
    *reinterpret_cast<float4*>(&Bs[lr][lq * 4]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* o = part + (size_t)blockIdx.z * Co * Ci;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    const int ci = ci0 + tx * 4;
    if (co < Co && ci < Ci) *reinterpret_cast<float4*>(o + (size_t)co * Ci + ci) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

__global__ void k_dw_final(const float* __restrict__ part, int splits, int n, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int k = 0; k < splits; ++k) s += (double)part[(size_t)k * n + i];
  dW[i] = (float)s;
}

static int dw_splits(int64_t M, int Ci, int Co) {
  const int64_t tiles = cdiv(Ci, WT) * cdiv(Co, WT);
  int64_t want = cdiv((int64_t)kNumSMs * 4, tiles);        // ~4 CTAs per SM in total
  const int64_t max_by_rows = cdiv(M, 4 * WK);             // at least 64 rows per split
  if (want > max_by_rows) want = max_by_rows;
  if (want < 1) want = 1;
  return (int)want;
}

static int check_lin(const char* op, int64_t M, int Ci, int Co) {
  TW_CHECK_ARG(M >= 0 && Ci >= 4 && Co >= 4 && (Ci & 3) == 0 && (Co & 3) == 0 && Ci <= 1024 && Co <= 1024,
               "%s: Ci=%d Co=%d must be multiples of 4 in [4,1024]", op, Ci, Co);
  return 0;
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_linear_fwd(const float* X, const float* W, int64_t M, int32_t Ci, int32_t Co, float* Z, int32_t impl,
                                void* stream) {
  if (int rc = check_lin("linear_fwd", M, Ci, Co)) return rc;
  TW_CHECK_ARG(aligned16(X) && aligned16(W) && aligned16(Z), "linear_fwd: 16-byte alignment required");
  TW_CHECK_ARG(impl >= 0 && impl <= 2, "linear_fwd: impl %d unknown (0 = SIMT FFMA, 1 = tcgen05 3xTF32, 2 = auto)", impl);
  if (M == 0) return 0;
  if (impl == 1 || (impl == 2 && linear_tc_supported(Ci, Co))) return linear_tc(X, W, Z, M, Ci, Co, 0, (cudaStream_t)stream);
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(Co, BN));
  k_gemm_rows<true><<<grid, kGemmThreads, 0, (cudaStream_t)stream>>>(X, W, Z, M, Co, Ci);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_linear_bwd_input(const float* dZ, const float* W, int64_t M, int32_t Ci, int32_t Co, float* dX,
                                      int32_t impl, void* stream) {
  if (int rc = check_lin("linear_bwd_input", M, Ci, Co)) return rc;
  TW_CHECK_ARG(aligned16(dZ) && aligned16(W) && aligned16(dX), "linear_bwd_input: 16-byte alignment required");
  TW_CHECK_ARG(impl >= 0 && impl <= 2, "linear_bwd_input: impl %d unknown (0 = SIMT FFMA, 1 = tcgen05 3xTF32, 2 = auto)", impl);
  if (M == 0) return 0;
  if (impl == 1 || (impl == 2 && linear_tc_supported(Co, Ci))) return linear_tc(dZ, W, dX, M, Co, Ci, 1, (cudaStream_t)stream);
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(Ci, BN));
  k_gemm_rows<false><<<grid, kGemmThreads, 0, (cudaStream_t)stream>>>(dZ, W, dX, M, Ci, Co);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t twowl_linear_bwd_weight_workspace_bytes(int64_t M, int32_t Ci, int32_t Co) {
  return align_up((size_t)dw_splits(M > 0 ? M : 1, Ci, Co) * (size_t)Ci * Co * sizeof(float));
}

extern "C" int twowl_linear_bwd_weight_scaled(const float* dZ, const float* row_scale, const float* X, int64_t M, int32_t Ci,
                                              int32_t Co, float* dW, void* ws, size_t ws_bytes, void* stream);
extern "C" int twowl_linear_bwd_weight(const float* dZ, const float* X, int64_t M, int32_t Ci, int32_t Co, float* dW, void* ws,
                                       size_t ws_bytes, void* stream) {
  return twowl_linear_bwd_weight_scaled(dZ, nullptr, X, M, Ci, Co, dW, ws, ws_bytes, stream);
}

extern "C" int twowl_linear_bwd_weight_scaled(const float* dZ, const float* row_scale, const float* X, int64_t M, int32_t Ci,
                                              int32_t Co, float* dW, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_lin("linear_bwd_weight", M, Ci, Co)) return rc;
  TW_CHECK_ARG(aligned16(dZ) && aligned16(X) && aligned16(dW), "linear_bwd_weight: 16-byte alignment required");
  TW_CHECK_WS(ws_bytes, twowl_linear_bwd_weight_workspace_bytes(M, Ci, Co));
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0) {
    TW_CUDA(cudaMemsetAsync(dW, 0, (size_t)Ci * Co * sizeof(float), s));
    return 0;
  }
  const int splits = dw_splits(M, Ci, Co);
  int64_t rows_per_split = cdiv(M, splits);
  rows_per_split = cdiv(rows_per_split, WK) * WK;
  dim3 grid((unsigned)cdiv(Ci, WT), (unsigned)cdiv(Co, WT), (unsigned)splits);
  k_dw_partial<<<grid, kGemmThreads, 0, s>>>(dZ, row_scale, X, M, Ci, Co, rows_per_split, (float*)ws);
  k_dw_final<<<(int)cdiv((int64_t)Ci * Co, 256), 256, 0, s>>>((const float*)ws, splits, Ci * Co, dW);
  TW_LAUNCH_CHECK();
  return 0;
}
