// pair_conv.cu - the pair-level GCNConv of TwoWL/model/model.py:77 with its dense linear layer on tcgen05 tensor
// cores and the structured ("factorised") aggregation fused into the GEMM's epilogue:
//
//   out[r, :] = sum_{s < nsrc} rs_s[r] * (A_s[r, :] * B_s^T)  +  sum_{g < ngather} coef_g[r] * T_g[idx_g[r], :]  +  bias
//
//   forward, one direction d :  A = H, rs = selfw_d, B = W_d, gather (S_d, centre_d, dinv_d), bias_d, + column
//                               statistics of `out` for the GraphNorm that follows (no extra pass over out)
//   backward w.r.t. H        :  A = (dO_f, dO_r), rs = (selfw_f, selfw_r), B = (W_f^T, W_r^T), gathers
//                               ((dS_f W_f), node_f, dinv_f) and ((dS_r W_r), node_r, dinv_r)
//   plain linear             :  nsrc = 1, no scale, no gather
//
// One persistent CTA per SM, 12 warps, warp-specialised:
//   warp  0    TMA producer: one lane streams raw fp32 128-row tiles of A_s into a ring of shared-memory stages with
//                            cp.async.bulk.tensor (SWIZZLE_128B tensor map = the UMMA canonical K-major layout, L2
//                            evict-first), 2-3 tiles in flight per SM
//   warps 2-3  split       : lo = x - trunc_tf32(x) of a landed tile into a second buffer (smem -> smem). The tensor
//                            core TRUNCATES fp32 operands to tf32 (measured, tools/mma_probe.cu), so the raw tile IS
//                            the `hi` operand of the 3xTF32 scheme and is never rewritten
//   warp  1    MMA issuer  : one lane issues 3 x Kd/8 tcgen05.mma kind::tf32 per source (lo*hi, hi*lo, hi*hi) into that
//                            source's TMEM accumulator (double buffered); tcgen05.commit releases the stages / publishes
//   warps 4-11 epilogue    : (4 TMEM lane quarters x 2 column halves) gathered rows requested before the accumulator
//                            is waited for; tcgen05.ld -> row scale, source sum -> per-warp smem transpose -> coalesced
//                            128-byte row segments: gather terms, bias, store, column sums (fp32 per tile -> double per
//                            CTA, fixed order)
// HBM-bound by design: 4*M*(nsrc*Kd + Nd) bytes + gathers for 6*M*Kd*Nd*nsrc tensor flop.
#include "common.cuh"

namespace twowl {

constexpr int kPcSplitWarps = 2;
constexpr int kPcEpilogueWarps = 8;
constexpr int kPcFirstSplit = 2;
constexpr int kPcFirstEpi = kPcFirstSplit + kPcSplitWarps;                    // 4
constexpr int kPcThreads = (kPcFirstEpi + kPcEpilogueWarps) * 32;            // 384: 12 warps -> up to 168 registers per thread
constexpr int kPcTileM = 128;
constexpr int kPcMaxStages = 4;

struct ConvParams {
  const float* rs[2];
  const float* W[2];
  int w_kn[2];
  int nsrc;
  int64_t M;
  int Nd;
  const float* T[2];
  const int32_t* tidx[2];
  const float* tcoef[2];
  int ngather;
  const float* bias;
  float* out;
  double* stats_part;  // [gridDim.x][2][Nd] column (sum, sum of squares) of `out`, or NULL
  int tmem_cols;
  int stages;          // raw-tile ring depth (2..kPcMaxStages)
};

__device__ __forceinline__ uint32_t pc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pc_desc(uint32_t saddr) {
  // K-major SWIZZLE_128B, 8-row groups 1024 B apart, descriptor version 1 (sm_100)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void pc_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(pc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void pc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pc_tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          pc_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(pc_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void pc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void pc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void pc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void pc_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void pc_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void pc_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pc_sw128(int r, int c) { return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ r) & 7) << 4)); }
__device__ __forceinline__ float pc_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void pc_split(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x, lo.y = v.y - hi.y, lo.z = v.z - hi.z, lo.w = v.w - hi.w;
}

template <int KD, int NG>
__global__ void __launch_bounds__(kPcThreads, 1) k_pair_conv(const ConvParams p, const __grid_constant__ CUtensorMap tmA0,
                                                             const __grid_constant__ CUtensorMap tmA1) {
  constexpr int KB = KD / 32;
  constexpr int K4 = KD / 4;
  constexpr uint32_t kTile = (uint32_t)kPcTileM * KD * 4u;  // one raw (or lo) tile
  extern __shared__ uint8_t pc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int Nd = p.Nd;
  const int S = p.stages;
  const uint32_t kBMat = (uint32_t)KD * Nd * 4u;  // one of hi / lo of one source
  uint8_t* Bs = smem;                                   // [nsrc][hi, lo]
  uint8_t* As = Bs + (size_t)p.nsrc * 2 * kBMat;        // [S] raw tiles
  uint8_t* Ls = As + (size_t)S * kTile;                 // lo tile
  uint8_t* Es = Ls + kTile;                             // [epilogue warp] 32 rows x 32 cols fp32 transpose buffer
  uint64_t* bars = reinterpret_cast<uint64_t*>(Es + kPcEpilogueWarps * 4096);
  uint64_t* raw_full = bars;                       // [S]  TMA -> split, MMA
  uint64_t* raw_empty = bars + kPcMaxStages;       // [S]  MMA -> TMA
  uint64_t* lo_full = bars + 2 * kPcMaxStages;     //      split -> MMA
  uint64_t* lo_empty = lo_full + 1;                //      MMA -> split
  uint64_t* tfull = lo_full + 2;                   // [2]  MMA -> epilogue
  uint64_t* tempty = lo_full + 4;                  // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_full + 6);
  double* red = reinterpret_cast<double*>(lo_full + 8);   // [epilogue warps][2][Nd] end-of-kernel reduction

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (p.M + kPcTileM - 1) / kPcTileM;
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t nuse = my_tiles * p.nsrc;  // stage uses, in order (tile, src)

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pc_smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kPcMaxStages; ++s) {
      pc_mbar_init(&raw_full[s], 1);
      pc_mbar_init(&raw_empty[s], 1);
    }
    pc_mbar_init(lo_full, kPcSplitWarps * 32);
    pc_mbar_init(lo_empty, 1);
    for (int a = 0; a < 2; ++a) {
      pc_mbar_init(&tfull[a], 1);
      pc_mbar_init(&tempty[a], kPcEpilogueWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA0)) : "memory");
    if (p.nsrc > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA1)) : "memory");
  }
  // resident weights: hi/lo split in the canonical layout (slab = 32 k-values, Nd rows of 128 bytes)
  for (int s = 0; s < p.nsrc; ++s) {
    const float* __restrict__ W = (s == 0) ? p.W[0] : p.W[1];
    const int w_kn = (s == 0) ? p.w_kn[0] : p.w_kn[1];
    uint8_t* Bhi = Bs + (size_t)s * 2 * kBMat;
    uint8_t* Blo = Bhi + kBMat;
    for (int i = tid; i < Nd * K4; i += kPcThreads) {
      const int n = i / K4, k4 = i % K4;
      float4 v;
      if (!w_kn) {
        v = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * KD + k4 * 4));
      } else {
        v.x = __ldg(W + (size_t)(k4 * 4 + 0) * Nd + n);
        v.y = __ldg(W + (size_t)(k4 * 4 + 1) * Nd + n);
        v.z = __ldg(W + (size_t)(k4 * 4 + 2) * Nd + n);
        v.w = __ldg(W + (size_t)(k4 * 4 + 3) * Nd + n);
      }
      float4 hi, lo;
      pc_split(v, hi, lo);
      const uint32_t off = (uint32_t)(k4 >> 3) * (uint32_t)Nd * 128u + pc_sw128(n, k4 & 7);
      *reinterpret_cast<float4*>(Bhi + off) = hi;
      *reinterpret_cast<float4*>(Blo + off) = lo;
    }
  }
  if (p.stats_part)
    for (int i = tid; i < kPcEpilogueWarps * 2 * Nd; i += kPcThreads) red[i] = 0.0;
  pc_proxy_fence();
  pc_fence_before();
  __syncthreads();
  pc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (int64_t u = 0; u < nuse; ++u) {
        const int64_t tile = blockIdx.x + (u / p.nsrc) * gridDim.x;
        const int s = (int)(u % p.nsrc);
        const int st = (int)(u % S);
        pc_mbar_wait(&raw_empty[st], (uint32_t)(((u / S) & 1) ^ 1));
        pc_mbar_expect_tx(&raw_full[st], kTile);
        const CUtensorMap* tm = s ? &tmA1 : &tmA0;
        uint8_t* dst = As + (size_t)st * kTile;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) pc_tma_load_2d(dst + (size_t)kb * kPcTileM * 128, tm, kb * 32, (int)(tile * kPcTileM), &raw_full[st], policy);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nd >> 3) << 17) | ((uint32_t)(kPcTileM >> 4) << 24);
      int64_t u = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        const int a = (int)(ti & 1);
        pc_mbar_wait(&tempty[a], (uint32_t)(((ti >> 1) & 1) ^ 1));
        pc_fence_after();
        for (int s = 0; s < p.nsrc; ++s, ++u) {
          const int st = (int)(u % S);
          const uint32_t tacc = tmem_base + (uint32_t)(a * (p.tmem_cols >> 1) + s * Nd);
          pc_mbar_wait(&raw_full[st], (uint32_t)((u / S) & 1));
          pc_mbar_wait(lo_full, (uint32_t)(u & 1));
          pc_fence_after();
          const uint8_t* Ahi = As + (size_t)st * kTile;
          const uint8_t* Alo = Ls;
          const uint8_t* Bhi = Bs + (size_t)s * 2 * kBMat;
          const uint8_t* Blo = Bhi + kBMat;
          uint32_t acc = 0;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {  // lo*hi, hi*lo, hi*hi: small terms first
            const uint8_t* Ap = (pass == 0) ? Alo : Ahi;
            const uint8_t* Bp = (pass == 1) ? Blo : Bhi;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                pc_mma(tacc, pc_desc(pc_smem_u32(Ap + (size_t)kb * kPcTileM * 128) + k * 32),
                       pc_desc(pc_smem_u32(Bp + (size_t)kb * Nd * 128) + k * 32), idesc, acc);
                acc = 1;
              }
            }
            if (pass == 0) pc_commit(lo_empty);  // the lo tile is free once the first pass has read it
          }
          pc_commit(&raw_empty[st]);  // stage reusable once these MMAs have read it
        }
        pc_commit(&tfull[a]);         // accumulators complete
      }
    }
    __syncwarp();
  } else if (warp < kPcFirstEpi) {
    // ===================================================== split: lo = x - trunc_tf32(x), same (swizzled) offsets
    const int t = tid - kPcFirstSplit * 32;  // 0..63
    constexpr int kSplitThreads = kPcSplitWarps * 32;
    constexpr int kBatch = 16;                                       // float4 per thread per batch
    constexpr int kBatches = (int)(kTile / 16u) / (kSplitThreads * kBatch);
    for (int64_t u = 0; u < nuse; ++u) {
      const int st = (int)(u % S);
      pc_mbar_wait(&raw_full[st], (uint32_t)((u / S) & 1));
      const float4* __restrict__ src = reinterpret_cast<const float4*>(As + (size_t)st * kTile);
      float4* __restrict__ dst = reinterpret_cast<float4*>(Ls);
      float4 x[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) x[j] = src[j * kSplitThreads + t];
      pc_mbar_wait(lo_empty, (uint32_t)((u & 1) ^ 1));
#pragma unroll
      for (int b = 0; b < kBatches; ++b) {
#pragma unroll
        for (int j = 0; j < kBatch; ++j)
          dst[(b * kBatch + j) * kSplitThreads + t] = make_float4(pc_lo(x[j].x), pc_lo(x[j].y), pc_lo(x[j].z), pc_lo(x[j].w));
        if (b + 1 < kBatches) {
#pragma unroll
          for (int j = 0; j < kBatch; ++j) x[j] = src[((b + 1) * kBatch + j) * kSplitThreads + t];
        }
      }
      pc_proxy_fence();
      pc_mbar_arrive(lo_full);
    }
  } else {
    // ===================================================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves
    const int ew = warp - kPcFirstEpi;
    const int quarter = warp & 3, half = ew >> 2;   // a warp may only touch TMEM lanes 32*(warp%4)..+31
    uint8_t* Et = Es + ew * 4096;
    const int lrow = lane >> 3, lchunk = lane & 7;   // coalesced phase: 4 rows x 8 chunks of 16 bytes per pass
    constexpr int NGA = NG > 0 ? NG : 1;
    int ix[NGA][8], ixn[NGA][8];
    float cf[NGA][8], cfn[NGA][8];
    float rsc[2], rscn[2];
    auto load_idx = [&](int64_t tile_, int (&ixx)[NGA][8], float (&cff)[NGA][8], float (&rss)[2]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = tile_ * kPcTileM + quarter * 32 + i * 4 + lrow;
#pragma unroll
        for (int g = 0; g < NGA; ++g) {
          const bool ok = g < NG && row < p.M;
          ixx[g][i] = ok ? __ldg(p.tidx[g] + row) : -1;
          cff[g][i] = ok ? __ldg(p.tcoef[g] + row) : 0.f;
        }
      }
      const int64_t myrow = tile_ * kPcTileM + quarter * 32 + lane;   // TMEM phase: lane = row
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float* __restrict__ rsv = s ? p.rs[1] : p.rs[0];
        rss[s] = (s < p.nsrc && rsv && myrow < p.M) ? __ldg(rsv + myrow) : 1.f;
      }
    };
    if (blockIdx.x < ntiles) load_idx(blockIdx.x, ix, cf, rsc);
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      const int64_t tile = blockIdx.x + ti * gridDim.x;
      const int a = (int)(ti & 1);
      const uint32_t taddr = tmem_base + (uint32_t)(a * (p.tmem_cols >> 1)) + ((uint32_t)(quarter * 32) << 16);
      const int64_t wrow0 = tile * kPcTileM + quarter * 32;
      bool waited = false;
      for (int c0 = half * 32; c0 < Nd; c0 += 64) {
        const int col = c0 + lchunk * 4;
        // gathered rows first: they do not depend on the accumulator, so their latency hides behind the wait for it
        float4 gv[NGA][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int g = 0; g < NGA; ++g)
            gv[g][i] = (NG > g && ix[g][i] >= 0 && col < Nd)
                           ? ldg_cached(reinterpret_cast<const float4*>(p.T[g] + (size_t)ix[g][i] * Nd + col)) : f4_zero();
        }
        if (!waited) {
          // next tile's indices / coefficients / row scales - one more load latency off the chain (NG <= 1: registers)
          if (NG <= 1 && ti + 1 < my_tiles) load_idx(tile + gridDim.x, ixn, cfn, rscn);
          pc_mbar_wait(&tfull[a], (uint32_t)((ti >> 1) & 1));
          pc_fence_after();
          waited = true;
        }
        // TMEM -> registers (lane = row), row scale and source sum, transpose through smem: 16-byte chunk q of row
        // `lane` goes to chunk q ^ (lane & 7)
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) {
          uint32_t v0[8], v1[8];
          pc_tmem_ld8(taddr + c0 + q2 * 8, v0);
          if (p.nsrc > 1) pc_tmem_ld8(taddr + Nd + c0 + q2 * 8, v1);
          pc_tmem_wait_ld();
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = rsc[0] * __uint_as_float(v0[e]);
          if (p.nsrc > 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(rsc[1], __uint_as_float(v1[e]), o[e]);
          }
          *reinterpret_cast<float4*>(Et + lane * 128 + (((q2 * 2) ^ (lane & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(Et + lane * 128 + (((q2 * 2 + 1) ^ (lane & 7)) << 4)) = make_float4(o[4], o[5], o[6], o[7]);
        }
        __syncwarp();
        float4 bsum = f4_zero(), bsq = f4_zero();
        const float4 bias4 = (p.bias && col < Nd) ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : f4_zero();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + lrow;
          const int64_t row = wrow0 + rl;
          float4 o = *reinterpret_cast<const float4*>(Et + rl * 128 + ((lchunk ^ (rl & 7)) << 4));
          if (row < p.M && col < Nd) {
#pragma unroll
            for (int g = 0; g < NGA; ++g)
              if (NG > g) f4_fma(o, cf[g][i], gv[g][i]);
            f4_add(o, bias4);
            *reinterpret_cast<float4*>(p.out + row * Nd + col) = o;
            f4_add(bsum, o);
            bsq.x = fmaf(o.x, o.x, bsq.x), bsq.y = fmaf(o.y, o.y, bsq.y), bsq.z = fmaf(o.z, o.z, bsq.z), bsq.w = fmaf(o.w, o.w, bsq.w);
          }
        }
        if (p.stats_part) {
          // lanes with the same chunk (lane & 7) hold partial sums of the same 4 columns: fixed-order xor tree,
          // then one lane per chunk adds the 32-row partial to this warp's double accumulators in smem
          float ps[8] = {bsum.x, bsum.y, bsum.z, bsum.w, bsq.x, bsq.y, bsq.z, bsq.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            ps[e] += __shfl_xor_sync(0xffffffffu, ps[e], 8);
            ps[e] += __shfl_xor_sync(0xffffffffu, ps[e], 16);
          }
          if (lrow == 0 && col < Nd) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              red[(ew * 2 + 0) * Nd + col + e] += (double)ps[e];
              red[(ew * 2 + 1) * Nd + col + e] += (double)ps[4 + e];
            }
          }
        }
        __syncwarp();
      }
      if (!waited) {  // this warp has no column block (Nd <= 32 and half == 1): still take part in the handshake
        if (NG <= 1 && ti + 1 < my_tiles) load_idx(tile + gridDim.x, ixn, cfn, rscn);
        pc_mbar_wait(&tfull[a], (uint32_t)((ti >> 1) & 1));
        pc_fence_after();
      }
      pc_fence_before();
      pc_mbar_arrive(&tempty[a]);
      if (NG <= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ix[0][i] = ixn[0][i], cf[0][i] = cfn[0][i];
        rsc[0] = rscn[0], rsc[1] = rscn[1];
      } else if (ti + 1 < my_tiles) {
        load_idx(tile + gridDim.x, ix, cf, rsc);
      }
    }
  }
  pc_fence_before();
  __syncthreads();
  if (p.stats_part) {
    for (int i = tid; i < 2 * Nd; i += kPcThreads) {
      const int v = i / Nd, c = i % Nd;
      double s = 0;
      for (int w = 0; w < kPcEpilogueWarps; ++w) s += red[(w * 2 + v) * Nd + c];
      p.stats_part[((size_t)blockIdx.x * 2 + v) * Nd + c] = s;
    }
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

// mean / inv_std of GraphNorm from per-CTA (sum, sum of squares) double partials (norm.cu semantics)
__global__ void k_pc_stats_final(const double* __restrict__ part, int nparts, int64_t M, int C, const float* __restrict__ mean_scale,
                                 float eps, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int b = 0; b < nparts; ++b) {
    s += part[((size_t)b * 2 + 0) * C + c];
    q += part[((size_t)b * 2 + 1) * C + c];
  }
  const double mean = s / (double)M;
  double var = q / (double)M - mean * mean;
  if (var < 0) var = 0;
  const double a = (double)mean_scale[c];
  stats[c] = (float)mean;
  stats[C + c] = (float)(1.0 / sqrt(var + (1.0 - a) * (1.0 - a) * mean * mean + (double)eps));
}

static size_t pc_smem_bytes(int Kd, int Nd, int nsrc, bool with_stats, int stages) {
  return (size_t)nsrc * 2 * Kd * Nd * 4 + (size_t)(stages + 1) * kPcTileM * Kd * 4 + kPcEpilogueWarps * 4096 + (2 * kPcMaxStages + 8) * 8 +
         (with_stats ? (size_t)kPcEpilogueWarps * 2 * Nd * 8 : 0) + 1024;
}
// deepest raw-tile ring (<= kPcMaxStages) that fits the 227 KB of shared memory, 0 if not even 2 stages fit
static int pc_stages(int Kd, int Nd, int nsrc, bool with_stats) {
  for (int st = kPcMaxStages; st >= 2; --st)
    if (pc_smem_bytes(Kd, Nd, nsrc, with_stats, st) <= 227 * 1024) return st;
  return 0;
}
static bool pc_supported(int Kd, int Nd, int nsrc) {
  return (Kd == 32 || Kd == 64) && Nd >= 16 && Nd <= 256 && (Nd % 16) == 0 && 2 * nsrc * Nd <= 512 && pc_stages(Kd, Nd, nsrc, nsrc == 1) >= 2;
}
static int pc_grid(int64_t M) {
  const int64_t ntiles = cdiv(M, kPcTileM);
  return (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
}

// row-major fp32 [M, Kd] -> boxes of 32 columns x 128 rows, SWIZZLE_128B (the UMMA K-major canonical layout)
static int pc_make_tmap(CUtensorMap* tm, const float* A, int64_t M, int Kd) {
  return make_tmap_2d_f32(tm, A, M, Kd, 32, kPcTileM, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int KD, int NG>
static int pc_launch(const ConvParams& p, const CUtensorMap& t0, const CUtensorMap& t1, cudaStream_t s) {
  const size_t smem = pc_smem_bytes(KD, p.Nd, p.nsrc, p.stats_part != nullptr, p.stages);
  TW_CUDA(cudaFuncSetAttribute(k_pair_conv<KD, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pair_conv<KD, NG><<<pc_grid(p.M), kPcThreads, smem, s>>>(p, t0, t1);
  TW_LAUNCH_CHECK();
  return 0;
}
template <int KD>
static int pc_launch_ng(const ConvParams& p, const CUtensorMap& t0, const CUtensorMap& t1, cudaStream_t s) {
  return p.ngather == 0 ? pc_launch<KD, 0>(p, t0, t1, s) : p.ngather == 1 ? pc_launch<KD, 1>(p, t0, t1, s) : pc_launch<KD, 2>(p, t0, t1, s);
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_pair_conv_supported(int32_t Kd, int32_t Nd, int32_t nsrc) { return pc_supported(Kd, Nd, nsrc) ? 1 : 0; }

extern "C" size_t twowl_pair_conv_workspace_bytes(int64_t M, int32_t Nd) {
  (void)M;
  return align_up((size_t)kNumSMs * 2 * (size_t)Nd * sizeof(double));
}

extern "C" int twowl_pair_conv(const twowl_conv_args* a, void* ws, size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(a != nullptr && a->nsrc >= 1 && a->nsrc <= 2 && a->ngather >= 0 && a->ngather <= 2, "pair_conv: bad nsrc/ngather");
  TW_CHECK_ARG(pc_supported(a->Kd, a->Nd, a->nsrc), "pair_conv: Kd=%d Nd=%d nsrc=%d unsupported (Kd in {32,64}, Nd %% 16, smem)",
               a->Kd, a->Nd, a->nsrc);
  TW_CHECK_ARG(a->M >= 0, "pair_conv: negative M");
  ConvParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < a->nsrc; ++s) {
    TW_CHECK_ARG(aligned16(a->A[s]) && aligned16(a->W[s]), "pair_conv: A/W must be 16-byte aligned");
    p.rs[s] = a->row_scale[s], p.W[s] = a->W[s], p.w_kn[s] = a->w_kn[s];
  }
  for (int g = 0; g < a->ngather; ++g) {
    TW_CHECK_ARG(a->T[g] && a->tidx[g] && a->tcoef[g] && aligned16(a->T[g]), "pair_conv: incomplete gather term %d", g);
    p.T[g] = a->T[g], p.tidx[g] = a->tidx[g], p.tcoef[g] = a->tcoef[g];
  }
  TW_CHECK_ARG(aligned16(a->out) && aligned16(a->bias), "pair_conv: out/bias must be 16-byte aligned");
  p.nsrc = a->nsrc, p.ngather = a->ngather, p.M = a->M, p.Nd = a->Nd, p.bias = a->bias, p.out = a->out;
  int cols = 32;
  while (cols < 2 * a->nsrc * a->Nd) cols <<= 1;
  p.tmem_cols = cols;
  cudaStream_t s = (cudaStream_t)stream;
  const bool want_stats = a->stats != nullptr;
  if (want_stats) {
    TW_CHECK_ARG(a->mean_scale != nullptr && a->M > 0, "pair_conv: stats need mean_scale and M > 0");
    TW_CHECK_WS(ws_bytes, twowl_pair_conv_workspace_bytes(a->M, a->Nd));
    p.stats_part = (double*)ws;
  }
  if (a->M == 0) return 0;
  p.stages = pc_stages(a->Kd, a->Nd, a->nsrc, want_stats);
  TW_CHECK_ARG(p.stages >= 2, "pair_conv: shared memory does not hold Kd=%d Nd=%d nsrc=%d%s", a->Kd, a->Nd, a->nsrc,
               want_stats ? " with statistics" : "");
  CUtensorMap tm[2];
  memset(tm, 0, sizeof(tm));
  for (int i = 0; i < a->nsrc; ++i) {
    const int rc_t = pc_make_tmap(&tm[i], a->A[i], a->M, a->Kd);
    if (rc_t) return rc_t;
  }
  int rc = (a->Kd == 32) ? pc_launch_ng<32>(p, tm[0], tm[1], s) : pc_launch_ng<64>(p, tm[0], tm[1], s);
  if (rc) return rc;
  if (want_stats) {
    k_pc_stats_final<<<(int)cdiv(a->Nd, 128), 128, 0, s>>>(p.stats_part, pc_grid(a->M), a->M, a->Nd, a->mean_scale, a->eps, a->stats);
    TW_LAUNCH_CHECK();
  }
  return 0;
}
