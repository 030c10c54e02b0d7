// agg.cu - the segmented gather-reduce every TwoWL aggregation runs through, plus the gcn_norm degree.
//
// Replaces, on the hot path, PyG's gather -> scale -> scatter_add_ inside GCNConv.propagate
// (reference call sites TwoWL/model/model.py:73 node level, :77 pair level, both directions) and the
// autograd transposes of the same ops. HBM/L2-bound fp32 work: no tensor cores here.
//
// Layout: CSR by OUTPUT row (ptr int64[M+1], col int32[nnz]); features fp32 row-major [rows, C], C % 4 == 0.
// A row is owned by a group of G lanes (G = pow2 >= C/4, <= 32); each lane keeps VEC float4 accumulators.
// The group loads G col entries at once (coalesced), then broadcasts them one by one with shuffles and
// issues the 128-bit row gathers 4 deep. Sums run in CSR order (= the reference's column order, because the
// CSR is built by a STABLE sort) in registers: deterministic, no atomics.
#include "common.cuh"

namespace twowl {

constexpr int kAggThreads = 256;

struct SegParams {
  const int64_t* ptr;
  const int32_t* col;
  int64_t M;
  const float* X;
  int C;
  int flip;
  int row_flip;
  const float* src_scale;
  const uint8_t* skip_mask;
  const uint8_t* row_skip_mask;
  int skip_self;
  int self_mode;
  const float* dst_scale;
  const float* bias;
  const float* X2;
  const int32_t* mul_idx;
  float* out;
  int accumulate;
};

template <int G, int VEC>
__global__ void __launch_bounds__(kAggThreads) k_seg_reduce(const SegParams p) {
  constexpr int kGroupsPerCta = kAggThreads / G;
  const int lane = threadIdx.x & 31;
  const int gl = threadIdx.x % G;                       // lane inside the group
  const int gbase = lane - gl;                          // first warp lane of the group
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const int cv = p.C >> 2;                              // float4 per row
  const int64_t group0 = (int64_t)blockIdx.x * kGroupsPerCta + threadIdx.x / G;
  const int64_t ngroups = (int64_t)gridDim.x * kGroupsPerCta;
  const float4* __restrict__ X4 = reinterpret_cast<const float4*>(p.X);
  const float4* __restrict__ X24 = reinterpret_cast<const float4*>(p.X2);

  // rows are dealt round-robin to groups: neighbouring hub rows land on different warps
  for (int64_t m = group0; m < p.M; m += ngroups) {
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = f4_zero();
    const int64_t r = m ^ (int64_t)p.row_flip;
    int64_t kb = 0, ke = 0;
    if (r < p.M && !(p.row_skip_mask && p.row_skip_mask[r])) {
      kb = p.ptr[r];
      ke = p.ptr[r + 1];
    }
    for (int64_t k0 = kb; k0 < ke; k0 += G) {
      // stage G entries: source row (or -1 = dropped), its scale, optional second-factor row
      int s_mine = -1, m2_mine = 0;
      float w_mine = 0.f;
      if (k0 + gl < ke) {
        const int c = __ldg(p.col + k0 + gl);
        const int s = c ^ p.flip;
        const bool keep = !(p.skip_self && (int64_t)s == m) && !(p.skip_mask && p.skip_mask[c]);
        if (keep) {
          s_mine = s;
          w_mine = p.src_scale ? __ldg(p.src_scale + s) : 1.f;
          if (p.X2) m2_mine = __ldg(p.mul_idx + c);
        }
      }
      const int cnt = (ke - k0 < G) ? (int)(ke - k0) : G;
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const int s = __shfl_sync(gmask, s_mine, gbase + j);
        const float w = __shfl_sync(gmask, w_mine, gbase + j);
        const int m2 = __shfl_sync(gmask, m2_mine, gbase + j);
        if (s >= 0) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int c4 = gl + v * G;
            if (c4 < cv) {
              float4 x = ldg_cached(X4 + (int64_t)s * cv + c4);
              if (p.X2) x = f4_mul(x, ldg_cached(X24 + (int64_t)m2 * cv + c4));
              f4_fma(acc[v], w, x);
            }
          }
        }
      }
    }
    const float ds = p.dst_scale ? p.dst_scale[m] : 1.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c4 = gl + v * G;
      if (c4 < cv) {
        float4 o = make_float4(ds * acc[v].x, ds * acc[v].y, ds * acc[v].z, ds * acc[v].w);
        if (p.self_mode == 1) f4_fma(o, ds * ds, ldg_cached(X4 + m * cv + c4));
        if (p.bias) f4_add(o, __ldg(reinterpret_cast<const float4*>(p.bias) + c4));
        float4* dst = reinterpret_cast<float4*>(p.out) + m * cv + c4;
        if (p.accumulate) f4_add(o, *dst);
        *dst = o;
      }
    }
  }
}

template <int G, int VEC>
static void launch_seg(const SegParams& p, cudaStream_t s) {
  constexpr int kGroupsPerCta = kAggThreads / G;
  const int grid = grid_for(p.M, kGroupsPerCta, 8);
  k_seg_reduce<G, VEC><<<grid, kAggThreads, 0, s>>>(p);
}

// deg[m] = 1 + #{k in row (m^row_flip) : (col[k]^flip) != m, !skip_mask[col[k]]}  ->  dinv = deg^-1/2
// one warp per row; integer counting, exact.
__global__ void __launch_bounds__(kAggThreads) k_gcn_dinv(const int64_t* __restrict__ ptr, const int32_t* __restrict__ col,
                                                          int64_t M, int flip, int row_flip,
                                                          const uint8_t* __restrict__ skip_mask,
                                                          const uint8_t* __restrict__ row_skip_mask, float* __restrict__ dinv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kAggThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kAggThreads) >> 5;
  for (int64_t m = warp0; m < M; m += nwarps) {
    const int64_t r = m ^ (int64_t)row_flip;
    int64_t kb = 0, ke = 0;
    if (r < M && !(row_skip_mask && row_skip_mask[r])) {
      kb = ptr[r];
      ke = ptr[r + 1];
    }
    int64_t c = 0;
    for (int64_t k = kb + lane; k - lane < ke; k += 32) {
      bool keep = false;
      if (k < ke) {
        const int cc = __ldg(col + k);
        keep = ((int64_t)(cc ^ flip) != m) && !(skip_mask && skip_mask[cc]);
      }
      c += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) dinv[m] = rsqrtf_exact((float)(c + 1));
  }
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_gcn_dinv(const int64_t* ptr, const int32_t* col, int64_t M, int32_t flip, int32_t row_flip,
                              const uint8_t* skip_mask, const uint8_t* row_skip_mask, float* dinv, void* stream) {
  TW_CHECK_ARG(M >= 0, "gcn_dinv: negative M");
  TW_CHECK_ARG(!(row_flip && (M & 1)), "gcn_dinv: row_flip needs an even row count");
  if (M == 0) return 0;
  k_gcn_dinv<<<grid_for(M, kAggThreads / 32, 8), kAggThreads, 0, (cudaStream_t)stream>>>(ptr, col, M, flip, row_flip, skip_mask,
                                                                                      row_skip_mask, dinv);
  TW_LAUNCH_CHECK();
  return 0;
}

extern "C" int twowl_seg_reduce(const twowl_seg_args* a, void* stream) {
  TW_CHECK_ARG(a != nullptr, "seg_reduce: null args");
  TW_CHECK_ARG(a->M >= 0 && a->C > 0 && (a->C & 3) == 0 && a->C <= 1024, "seg_reduce: C=%d must be a multiple of 4 in [4,1024]",
               a->C);
  TW_CHECK_ARG(aligned16(a->X) && aligned16(a->out) && aligned16(a->bias) && aligned16(a->X2),
               "seg_reduce: feature pointers must be 16-byte aligned");
  TW_CHECK_ARG(!(a->row_flip && (a->M & 1)), "seg_reduce: row_flip needs an even row count");
  TW_CHECK_ARG((a->X2 == nullptr) == (a->mul_idx == nullptr), "seg_reduce: X2 and mul_idx go together");
  if (a->M == 0) return 0;
  SegParams p;
  p.ptr = a->ptr, p.col = a->col, p.M = a->M, p.X = a->X, p.C = a->C, p.flip = a->flip, p.row_flip = a->row_flip;
  p.src_scale = a->src_scale, p.skip_mask = a->skip_mask, p.row_skip_mask = a->row_skip_mask;
  p.skip_self = a->skip_self, p.self_mode = a->self_mode, p.dst_scale = a->dst_scale, p.bias = a->bias;
  p.X2 = a->X2, p.mul_idx = a->mul_idx, p.out = a->out, p.accumulate = a->accumulate;
  cudaStream_t s = (cudaStream_t)stream;
  const int cv = a->C >> 2;
  if (cv <= 4) launch_seg<4, 1>(p, s);
  else if (cv <= 8) launch_seg<8, 1>(p, s);
  else if (cv <= 16) launch_seg<16, 1>(p, s);
  else if (cv <= 32) launch_seg<32, 1>(p, s);
  else if (cv <= 64) launch_seg<32, 2>(p, s);
  else if (cv <= 128) launch_seg<32, 4>(p, s);
  else launch_seg<32, 8>(p, s);
  TW_LAUNCH_CHECK();
  return 0;
}
