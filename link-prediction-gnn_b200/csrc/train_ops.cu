// train_ops.cu - the loss and the optimiser of the reference's train step as ONE kernel each:
//   F.binary_cross_entropy_with_logits(pred, y) forward + backward   (TwoWL/model/train.py:37-38)
//   torch.optim.Adam.step() over a FLAT parameter buffer              (TwoWL/model/train.py:39; Adam built at TwoWL_work.py:100)
// For the launch-bound small configurations (fb-pages-food, Cora scale: < 1 ms of GPU work per step) stock torch spends ~15
// launches on the loss (forward, mean, backward chain) and the foreach-Adam; here it is 2 + 1. Both are capturable in a CUDA
// graph: the Adam step count lives in device memory.
#include "common.cuh"

namespace twowl {

constexpr int kTrThreads = 256;
constexpr int kBceMaxCtas = 256;

// per element: l = max(x, 0) - x*y + log1p(exp(-|x|)) (the numerically stable form torch uses); d l / d x = sigmoid(x) - y
__global__ void __launch_bounds__(kTrThreads) k_bce_fwd_bwd(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                                                            float inv_n, float* __restrict__ dx, float* __restrict__ prob,
                                                            double* __restrict__ part) {
  __shared__ double s_sum[kTrThreads / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[i], yi = y[i];
    const float e = expf(-fabsf(xi));
    acc += (double)(fmaxf(xi, 0.f) - xi * yi + log1pf(e));
    // sigmoid and its complement both without cancellation; (1 - y) s - y (1 - s) = s - y, exact also where s rounds to 1
    const float s = xi >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    const float c = xi >= 0.f ? e / (1.f + e) : 1.f / (1.f + e);
    if (dx) dx[i] = ((1.f - yi) * s - yi * c) * inv_n;
    if (prob) prob[i] = s;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kTrThreads / 32; ++w) t += s_sum[w];   // fixed order
    part[blockIdx.x] = t;
  }
}
__global__ void k_bce_final(const double* __restrict__ part, int nparts, double inv_n, float* __restrict__ loss) {
  double t = 0.0;
  for (int b = 0; b < nparts; ++b) t += part[b];               // fixed order: deterministic
  loss[0] = (float)(t * inv_n);
}

// torch.optim.Adam (amsgrad = False, maximize = False): step t is read from / written to device memory
__global__ void __launch_bounds__(kTrThreads) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                                                     float weight_decay, const float* __restrict__ gscale, int64_t* __restrict__ step) {
  // the scalars as torch computes them on the host: in double, rounded to fp32 where they meet the tensors
  const int64_t t = step[0] + 1;
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step_size = (float)((double)lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
  const float gs = gscale ? gscale[0] : 1.f;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float m0 = m[i];
    const float mi = fmaf(w1, gi - m0, m0);                       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(w2 * gi, gi, v[i] * beta2);             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    m[i] = mi, v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / sqrt_bc2 + eps));  // param.addcdiv_(exp_avg, sqrt(v) / sqrt(bc2) + eps, -step_size)
  }
}
__global__ void k_adam_tick(int64_t* step) { step[0] += 1; }

}  // namespace twowl

using namespace twowl;

extern "C" size_t twowl_bce_logits_workspace_bytes(int64_t n) {
  (void)n;
  return align_up((size_t)kBceMaxCtas * sizeof(double));
}

extern "C" int twowl_bce_logits(const float* logits, const float* labels, int64_t n, float* loss, float* dlogits, float* prob, void* ws,
                                size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(n > 0 && logits && labels && loss, "bce_logits: needs n > 0 and non-null logits / labels / loss");
  TW_CHECK_WS(ws_bytes, twowl_bce_logits_workspace_bytes(n));
  cudaStream_t s = (cudaStream_t)stream;
  int grid = (int)cdiv(n, kTrThreads * 4);
  grid = grid < 1 ? 1 : (grid > kBceMaxCtas ? kBceMaxCtas : grid);
  k_bce_fwd_bwd<<<grid, kTrThreads, 0, s>>>(logits, labels, n, 1.f / (float)n, dlogits, prob, (double*)ws);
  k_bce_final<<<1, 1, 0, s>>>((const double*)ws, grid, 1.0 / (double)n, loss);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, const float* grad_scale, int64_t* step, void* stream) {
  TW_CHECK_ARG(n >= 0 && step && (n == 0 || (param && grad && exp_avg && exp_avg_sq)), "adam_step: null argument");
  TW_CHECK_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_step: bad hyper-parameters");
  cudaStream_t s = (cudaStream_t)stream;
  if (n > 0) k_adam<<<grid_for(n, kTrThreads), kTrThreads, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                 grad_scale, step);
  k_adam_tick<<<1, 1, 0, s>>>(step);
  TW_LAUNCH_CHECK();
  return 0;
}
