// nodefeat.cu - the node-attribute input branch of LocalWLNet (TwoWL/model/model.py:47-51,71, use_node_feat=True):
//   x = Dropout(dp_lin1)( LayerNorm(c, elementwise_affine=False)( Linear(F, c)( Dropout(dp_lin0)(node_feat) ) ) )
// The Linear's GEMM runs on the package's linear kernels (linear.cu / pair_conv.cu); this file holds what surrounds it:
// the input dropout (elementwise) and bias + LayerNorm + dropout fused into one row pass, forward and backward.
// HBM-bound: one warp per row, the row lives in registers (C <= 1024), statistics by the stable two-pass form.
#include "common.cuh"

namespace twowl {

constexpr int kNfThreads = 256;
constexpr int kNfVec = 8;   // float4 per lane: C <= 32 * 8 * 4 = 1024

__global__ void __launch_bounds__(kNfThreads) k_dropout(const float* __restrict__ x, int64_t n, uint32_t thresh, float inv_keep,
                                                        uint64_t seed, float* __restrict__ out) {
  seed = resolve_seed(seed);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] * drop_scale(seed, (uint64_t)i, thresh, inv_keep);
}

// y[m] = dropout( (u - mean(u)) / sqrt(var(u) + eps) ), u = z[m] + bias; stats[m] = (mean, inv_std)
__global__ void __launch_bounds__(kNfThreads) k_bias_ln_fwd(const float* __restrict__ z, const float* __restrict__ bias, int64_t M, int C,
                                                            float eps, uint32_t thresh, float inv_keep, uint64_t seed,
                                                            float* __restrict__ y, float* __restrict__ stats) {
  seed = resolve_seed(seed);
  const int lane = threadIdx.x & 31, cv = C >> 2;
  const int64_t warp0 = ((int64_t)blockIdx.x * kNfThreads + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kNfThreads) >> 5;
  const float4* __restrict__ b4 = reinterpret_cast<const float4*>(bias);
  for (int64_t m = warp0; m < M; m += nwarps) {
    const float4* __restrict__ z4 = reinterpret_cast<const float4*>(z) + m * cv;
    float4 u[kNfVec];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < kNfVec; ++v) {
      const int c4 = lane + v * 32;
      u[v] = f4_zero();
      if (c4 < cv) {
        u[v] = ldg_stream(z4 + c4);
        if (bias) f4_add(u[v], __ldg(b4 + c4));
        s += (u[v].x + u[v].y) + (u[v].z + u[v].w);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < kNfVec; ++v) {
      if (lane + v * 32 < cv) {
        const float a = u[v].x - mean, b = u[v].y - mean, c = u[v].z - mean, d = u[v].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    const float inv = rsqrtf_exact(q / (float)C + eps);
    float4* __restrict__ y4 = reinterpret_cast<float4*>(y) + m * cv;
#pragma unroll
    for (int v = 0; v < kNfVec; ++v) {
      const int c4 = lane + v * 32;
      if (c4 < cv) {
        float4 o = make_float4((u[v].x - mean) * inv, (u[v].y - mean) * inv, (u[v].z - mean) * inv, (u[v].w - mean) * inv);
        if (thresh) {
          const uint64_t e0 = (uint64_t)m * C + (uint64_t)c4 * 4;
          o.x *= drop_scale(seed, e0, thresh, inv_keep), o.y *= drop_scale(seed, e0 + 1, thresh, inv_keep);
          o.z *= drop_scale(seed, e0 + 2, thresh, inv_keep), o.w *= drop_scale(seed, e0 + 3, thresh, inv_keep);
        }
        stg_stream(y4 + c4, o);
      }
    }
    if (lane == 0) stats[2 * m] = mean, stats[2 * m + 1] = inv;
  }
}

// du[m] = inv_std * (g' - mean(g') - n * mean(g' * n)),  g' = g * dropout mask,  n = (u - mean) * inv_std
__global__ void __launch_bounds__(kNfThreads) k_bias_ln_bwd(const float* __restrict__ g, const float* __restrict__ z,
                                                            const float* __restrict__ bias, const float* __restrict__ stats, int64_t M,
                                                            int C, uint32_t thresh, float inv_keep, uint64_t seed,
                                                            float* __restrict__ dz) {
  seed = resolve_seed(seed);
  const int lane = threadIdx.x & 31, cv = C >> 2;
  const int64_t warp0 = ((int64_t)blockIdx.x * kNfThreads + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kNfThreads) >> 5;
  const float4* __restrict__ b4 = reinterpret_cast<const float4*>(bias);
  for (int64_t m = warp0; m < M; m += nwarps) {
    const float4* __restrict__ z4 = reinterpret_cast<const float4*>(z) + m * cv;
    const float4* __restrict__ g4 = reinterpret_cast<const float4*>(g) + m * cv;
    const float mean = stats[2 * m], inv = stats[2 * m + 1];
    float4 n[kNfVec], gp[kNfVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < kNfVec; ++v) {
      const int c4 = lane + v * 32;
      n[v] = gp[v] = f4_zero();
      if (c4 < cv) {
        float4 u = ldg_stream(z4 + c4);
        if (bias) f4_add(u, __ldg(b4 + c4));
        n[v] = make_float4((u.x - mean) * inv, (u.y - mean) * inv, (u.z - mean) * inv, (u.w - mean) * inv);
        gp[v] = ldg_stream(g4 + c4);
        if (thresh) {
          const uint64_t e0 = (uint64_t)m * C + (uint64_t)c4 * 4;
          gp[v].x *= drop_scale(seed, e0, thresh, inv_keep), gp[v].y *= drop_scale(seed, e0 + 1, thresh, inv_keep);
          gp[v].z *= drop_scale(seed, e0 + 2, thresh, inv_keep), gp[v].w *= drop_scale(seed, e0 + 3, thresh, inv_keep);
        }
        s1 += (gp[v].x + gp[v].y) + (gp[v].z + gp[v].w);
        s2 += (gp[v].x * n[v].x + gp[v].y * n[v].y) + (gp[v].z * n[v].z + gp[v].w * n[v].w);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    const float m1 = s1 / (float)C, m2 = s2 / (float)C;
    float4* __restrict__ d4 = reinterpret_cast<float4*>(dz) + m * cv;
#pragma unroll
    for (int v = 0; v < kNfVec; ++v) {
      const int c4 = lane + v * 32;
      if (c4 < cv)
        stg_stream(d4 + c4, make_float4(inv * (gp[v].x - m1 - n[v].x * m2), inv * (gp[v].y - m1 - n[v].y * m2),
                                        inv * (gp[v].z - m1 - n[v].z * m2), inv * (gp[v].w - m1 - n[v].w * m2)));
    }
  }
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_dropout(const float* x, int64_t n, float p, uint64_t seed, float* out, void* stream) {
  TW_CHECK_ARG(n >= 0 && p >= 0.f && p < 1.f, "dropout: n >= 0 and p in [0,1) required (p=%f)", p);
  if (n == 0) return 0;
  k_dropout<<<grid_for(n, kNfThreads), kNfThreads, 0, (cudaStream_t)stream>>>(x, n, p > 0.f ? drop_thresh(p) : 0u, 1.f / (1.f - p), seed, out);
  TW_LAUNCH_CHECK();
  return 0;
}

static int check_ln(const char* op, int64_t M, int C) {
  TW_CHECK_ARG(M >= 0 && C >= 4 && (C & 3) == 0 && C <= 1024, "%s: C=%d must be a multiple of 4 in [4,1024]", op, C);
  return 0;
}

extern "C" int twowl_bias_layernorm_fwd(const float* z, const float* bias, int64_t M, int32_t C, float eps, float p, uint64_t seed,
                                        float* y, float* stats, void* stream) {
  if (int rc = check_ln("bias_layernorm_fwd", M, C)) return rc;
  TW_CHECK_ARG(aligned16(z) && aligned16(bias) && aligned16(y) && stats && p >= 0.f && p < 1.f, "bias_layernorm_fwd: bad arguments");
  if (M == 0) return 0;
  k_bias_ln_fwd<<<grid_for(M, kNfThreads / 32, 8), kNfThreads, 0, (cudaStream_t)stream>>>(z, bias, M, C, eps, p > 0.f ? drop_thresh(p) : 0u,
                                                                                       1.f / (1.f - p), seed, y, stats);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_bias_layernorm_bwd(const float* g, const float* z, const float* bias, const float* stats, int64_t M, int32_t C,
                                        float p, uint64_t seed, float* dz, void* stream) {
  if (int rc = check_ln("bias_layernorm_bwd", M, C)) return rc;
  TW_CHECK_ARG(aligned16(g) && aligned16(z) && aligned16(bias) && aligned16(dz) && stats && p >= 0.f && p < 1.f,
               "bias_layernorm_bwd: bad arguments");
  if (M == 0) return 0;
  k_bias_ln_bwd<<<grid_for(M, kNfThreads / 32, 8), kNfThreads, 0, (cudaStream_t)stream>>>(g, z, bias, stats, M, C,
                                                                                       p > 0.f ? drop_thresh(p) : 0u, 1.f / (1.f - p), seed, dz);
  TW_LAUNCH_CHECK();
  return 0;
}
