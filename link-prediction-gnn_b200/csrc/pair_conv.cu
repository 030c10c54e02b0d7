// pair_conv.cu - the pair-level GCNConv of TwoWL/model/model.py:77 with its dense linear layer on tcgen05 tensor
// cores and the structured ("factorised") aggregation fused into the GEMM's prologue / epilogue:
//
//   out[r, :] = sum_{s < nsrc} (rs_s[r] * A_s[r, :]) * B_s^T  +  sum_{g < ngather} coef_g[r] * T_g[idx_g[r], :]  +  bias
//
//   forward, one direction d :  A = H, rs = selfw_d, B = W_d, gather (S_d, centre_d, dinv_d), bias_d, + column
//                               statistics of `out` for the GraphNorm that follows (no extra pass over out)
//   backward w.r.t. H        :  A = (dO_f, dO_r), rs = (selfw_f, selfw_r), B = (W_f^T, W_r^T), gathers
//                               ((dS_f W_f), node_f, dinv_f) and ((dS_r W_r), node_r, dinv_r)
//   plain linear             :  nsrc = 1, no scale, no gather
//
// One persistent CTA per SM, 13 warps, warp-specialised:
//   warps 0-3  producers : coalesced 128-bit global loads of a 128-row tile (the next tile's loads are in flight
//                          while the current one is staged), row scale, 3xTF32 hi/lo split, swizzled st.shared into a 2-stage ring
//                          (UMMA canonical K-major SWIZZLE_128B), fence.proxy.async, mbarrier arrive
//   warp  12   MMA issuer: one lane issues 3 x Kd/8 tcgen05.mma kind::tf32 per source into one of two TMEM
//                          accumulators; tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 4-11 epilogue  : (4 TMEM lane quarters x 2 column halves) gathered rows requested before the accumulator
//                          is waited for; tcgen05.ld -> per-warp smem transpose -> coalesced 128-byte row segments: gather terms,
//                          bias, store, column sums (fp32 per tile -> double per CTA, fixed order)
// HBM-bound by design: 4*M*(nsrc*Kd + Nd) bytes + gathers for 6*M*Kd*Nd*nsrc tensor flop.
#include "common.cuh"

namespace twowl {

constexpr int kPcProducerWarps = 4;
constexpr int kPcEpilogueWarps = 8;
constexpr int kPcThreads = (kPcProducerWarps + kPcEpilogueWarps + 1) * 32;  // 416
constexpr int kPcTileM = 128;
constexpr int kPcStages = 2;

struct ConvParams {
  const float* A[2];
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
__device__ __forceinline__ void pc_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pc_sw128(int r, int c) { return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ r) & 7) << 4)); }
__device__ __forceinline__ void pc_split(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x, lo.y = v.y - hi.y, lo.z = v.z - hi.z, lo.w = v.w - hi.w;
}

template <int KD, int NG>
__global__ void __launch_bounds__(kPcThreads, 1) k_pair_conv(const ConvParams p) {
  constexpr int KB = KD / 32;
  constexpr int K4 = KD / 4;
  constexpr int kChunks = kPcTileM * K4 / (kPcProducerWarps * 32);  // 16-byte chunks per producer thread per stage
  constexpr uint32_t kAStage = 2u * kPcTileM * KD * 4u;             // hi + lo
  extern __shared__ uint8_t pc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int Nd = p.Nd;
  const uint32_t kBMat = (uint32_t)KD * Nd * 4u;  // one of hi / lo of one source
  uint8_t* Bs = smem;                                   // [nsrc][hi, lo]
  uint8_t* As = Bs + (size_t)p.nsrc * 2 * kBMat;        // [stage][hi, lo]
  uint8_t* Es = As + (size_t)kPcStages * kAStage;       // [epilogue warp] 32 rows x 32 cols fp32 transpose buffer
  uint64_t* bars = reinterpret_cast<uint64_t*>(Es + kPcEpilogueWarps * 4096);
  uint64_t* full = bars;           // [stages]  producers -> MMA
  uint64_t* empty = bars + 2;      // [stages]  MMA -> producers
  uint64_t* tfull = bars + 4;      // [2]       MMA -> epilogue
  uint64_t* tempty = bars + 6;     // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  double* red = reinterpret_cast<double*>(bars + 10);   // [epilogue warps][2][Nd] end-of-kernel reduction

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (p.M + kPcTileM - 1) / kPcTileM;

  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pc_smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kPcStages; ++s) {
      pc_mbar_init(&full[s], kPcProducerWarps * 32);
      pc_mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      pc_mbar_init(&tfull[a], 1);
      pc_mbar_init(&tempty[a], kPcEpilogueWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

  if (warp < kPcProducerWarps) {
    // ===================================================== producers
    const int ptid = tid;  // 0..127
    float4 regs[kChunks];
    const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t nit = my_tiles * p.nsrc;  // stage uses, in order (tile, src)
    auto issue = [&](int64_t it) {
      const int64_t tile = blockIdx.x + (it / p.nsrc) * gridDim.x;
      const int s = (int)(it % p.nsrc);
      const float4* __restrict__ A4 = reinterpret_cast<const float4*>(s == 0 ? p.A[0] : p.A[1]) + tile * kPcTileM * K4;
      const int64_t rows_left = p.M - tile * kPcTileM;
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {
        const int idx = j * (kPcProducerWarps * 32) + ptid;
        regs[j] = (idx / K4 < rows_left) ? ldg_stream(A4 + idx) : f4_zero();
      }
    };
    auto stage = [&](int64_t it) {
      const int64_t tile = blockIdx.x + (it / p.nsrc) * gridDim.x;
      const int s = (int)(it % p.nsrc);
      const int st = (int)(it % kPcStages);
      const uint32_t ph = (uint32_t)((it / kPcStages) & 1);
      pc_mbar_wait(&empty[st], ph ^ 1u);
      uint8_t* Ahi = As + (size_t)st * kAStage;
      uint8_t* Alo = Ahi + kAStage / 2;
      const float* __restrict__ rsv = (s == 0) ? p.rs[0] : p.rs[1];
      const int64_t row0 = tile * kPcTileM;
#pragma unroll
      for (int j = 0; j < kChunks; ++j) {
        const int idx = j * (kPcProducerWarps * 32) + ptid;
        const int rr = idx / K4, k4 = idx % K4;
        float4 v = regs[j];
        if (rsv) {
          const float sc = (row0 + rr < p.M) ? __ldg(rsv + row0 + rr) : 0.f;
          v.x *= sc, v.y *= sc, v.z *= sc, v.w *= sc;
        }
        float4 hi, lo;
        pc_split(v, hi, lo);
        const uint32_t off = (uint32_t)(k4 >> 3) * (kPcTileM * 128u) + pc_sw128(rr, k4 & 7);
        *reinterpret_cast<float4*>(Ahi + off) = hi;
        *reinterpret_cast<float4*>(Alo + off) = lo;
      }
      pc_proxy_fence();
      pc_mbar_arrive(&full[st]);
    };
    // the loads of use it+1 are in flight while use it is staged and while the wait for its smem slot lasts
    if (nit > 0) issue(0);
    for (int64_t it = 0; it < nit; ++it) {
      stage(it);
      if (it + 1 < nit) issue(it + 1);
    }
  } else if (warp == 12) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nd >> 3) << 17) | ((uint32_t)(kPcTileM >> 4) << 24);
      int64_t it = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const int a = (int)(ti & 1);
        pc_mbar_wait(&tempty[a], (uint32_t)(((ti >> 1) & 1) ^ 1));
        pc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * (p.tmem_cols >> 1));
        uint32_t acc = 0;
        for (int s = 0; s < p.nsrc; ++s, ++it) {
          const int st = (int)(it % kPcStages);
          pc_mbar_wait(&full[st], (uint32_t)((it / kPcStages) & 1));
          pc_fence_after();
          const uint8_t* Ahi = As + (size_t)st * kAStage;
          const uint8_t* Alo = Ahi + kAStage / 2;
          const uint8_t* Bhi = Bs + (size_t)s * 2 * kBMat;
          const uint8_t* Blo = Bhi + kBMat;
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
          }
          pc_commit(&empty[st]);   // stage reusable once these MMAs have read it
        }
        pc_commit(&tfull[a]);      // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves
    const int ew = warp - kPcProducerWarps;
    const int quarter = ew & 3, half = ew >> 2;
    uint8_t* Et = Es + ew * 4096;
    const int lrow = lane >> 3, lchunk = lane & 7;   // coalesced phase: 4 rows x 8 chunks of 16 bytes per pass
    constexpr int NGA = NG > 0 ? NG : 1;
    int ix[NGA][8], ixn[NGA][8];
    float cf[NGA][8], cfn[NGA][8];
    auto load_idx = [&](int64_t tile_, int (&ixx)[NGA][8], float (&cff)[NGA][8]) {
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
    };
    if (NG > 0 && blockIdx.x < ntiles) load_idx(blockIdx.x, ix, cf);
    int64_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
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
          // next tile's indices / coefficients (NG <= 1 only: register budget) - one more load latency off the chain
          if (NG == 1 && tile + gridDim.x < ntiles) load_idx(tile + gridDim.x, ixn, cfn);
          pc_mbar_wait(&tfull[a], (uint32_t)((ti >> 1) & 1));
          pc_fence_after();
          waited = true;
        }
        uint32_t v[32];
        pc_tmem_ld32(taddr + c0, v);
        // transpose through smem: lane = row; 16-byte chunk q of row `lane` goes to chunk q ^ (lane & 7)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(Et + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
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
        pc_mbar_wait(&tfull[a], (uint32_t)((ti >> 1) & 1));
        pc_fence_after();
      }
      pc_fence_before();
      pc_mbar_arrive(&tempty[a]);
      if (NG == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ix[0][i] = ixn[0][i], cf[0][i] = cfn[0][i];
      } else if (NG == 2 && tile + gridDim.x < ntiles) {
        load_idx(tile + gridDim.x, ix, cf);
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
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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

static size_t pc_smem_bytes(int Kd, int Nd, int nsrc, bool with_stats = true) {
  return (size_t)nsrc * 2 * Kd * Nd * 4 + (size_t)kPcStages * 2 * kPcTileM * Kd * 4 + kPcEpilogueWarps * 4096 + 10 * 8 +
         (with_stats ? (size_t)kPcEpilogueWarps * 2 * Nd * 8 : 0) + 1024;
}
static bool pc_supported(int Kd, int Nd, int nsrc) {
  return (Kd == 32 || Kd == 64) && Nd >= 16 && Nd <= 256 && (Nd % 16) == 0 && 2 * Nd <= 512 && pc_smem_bytes(Kd, Nd, nsrc, nsrc == 1) <= 227 * 1024;
}
static int pc_grid(int64_t M) {
  const int64_t ntiles = cdiv(M, kPcTileM);
  return (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
}

template <int KD, int NG>
static int pc_launch(const ConvParams& p, cudaStream_t s) {
  const size_t smem = pc_smem_bytes(KD, p.Nd, p.nsrc, p.stats_part != nullptr);
  TW_CHECK_ARG(smem <= 227 * 1024, "pair_conv: %zu bytes of shared memory needed (statistics only with one source)", smem);
  TW_CUDA(cudaFuncSetAttribute(k_pair_conv<KD, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pair_conv<KD, NG><<<pc_grid(p.M), kPcThreads, smem, s>>>(p);
  TW_LAUNCH_CHECK();
  return 0;
}
template <int KD>
static int pc_launch_ng(const ConvParams& p, cudaStream_t s) {
  return p.ngather == 0 ? pc_launch<KD, 0>(p, s) : p.ngather == 1 ? pc_launch<KD, 1>(p, s) : pc_launch<KD, 2>(p, s);
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
    p.A[s] = a->A[s], p.rs[s] = a->row_scale[s], p.W[s] = a->W[s], p.w_kn[s] = a->w_kn[s];
  }
  for (int g = 0; g < a->ngather; ++g) {
    TW_CHECK_ARG(a->T[g] && a->tidx[g] && a->tcoef[g] && aligned16(a->T[g]), "pair_conv: incomplete gather term %d", g);
    p.T[g] = a->T[g], p.tidx[g] = a->tidx[g], p.tcoef[g] = a->tcoef[g];
  }
  TW_CHECK_ARG(aligned16(a->out) && aligned16(a->bias), "pair_conv: out/bias must be 16-byte aligned");
  p.nsrc = a->nsrc, p.ngather = a->ngather, p.M = a->M, p.Nd = a->Nd, p.bias = a->bias, p.out = a->out;
  int cols = 32;
  while (cols < 2 * a->Nd) cols <<= 1;
  p.tmem_cols = cols;
  cudaStream_t s = (cudaStream_t)stream;
  const bool want_stats = a->stats != nullptr;
  if (want_stats) {
    TW_CHECK_ARG(a->mean_scale != nullptr && a->M > 0, "pair_conv: stats need mean_scale and M > 0");
    TW_CHECK_WS(ws_bytes, twowl_pair_conv_workspace_bytes(a->M, a->Nd));
    p.stats_part = (double*)ws;
  }
  if (a->M == 0) return 0;
  int rc = (a->Kd == 32) ? pc_launch_ng<32>(p, s) : pc_launch_ng<64>(p, s);
  if (rc) return rc;
  if (want_stats) {
    k_pc_stats_final<<<(int)cdiv(a->Nd, 128), 128, 0, s>>>(p.stats_part, pc_grid(a->M), a->M, a->Nd, a->mean_scale, a->eps, a->stats);
    TW_LAUNCH_CHECK();
  }
  return 0;
}
