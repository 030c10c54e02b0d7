// linear_tc.cu - the dense per-row linear layer of GCNConv (reference TwoWL/model/model.py:37) on the
// 5th-generation tensor cores: tcgen05.mma kind::tf32 with 3xTF32 split operands (fp32-accurate), the
// accumulator in TMEM, operands staged in shared memory in the UMMA canonical K-major SWIZZLE_128B layout.
//
// Shape: C[M,Nd] = A[M,Kd] * B[Nd,Kd]^T with M = rows of the pair table (up to 6e7) and Kd, Nd <= 256: the op
// moves 4*M*(Kd+Nd) bytes for 2*M*Kd*Nd flop, i.e. it is HBM-bound once the math runs on tensor cores.
//   * B (the weight, <= 64 KB) is split ONCE per CTA into tf32 hi/lo parts and stays resident in smem.
//   * A persistent CTA (128 threads) walks 128-row tiles: coalesced 128-bit global loads -> hi/lo split in
//     registers -> swizzled st.shared -> fence.proxy.async -> one thread issues 3 x Kd/8 tcgen05.mma
//     (lo*hi, hi*lo, hi*hi: small terms first) -> tcgen05.commit on an mbarrier -> the next tile's global
//     loads are issued while the MMAs run -> tcgen05.ld of the 128 x Nd accumulator -> st.global.
//   * 2-3 CTAs per SM overlap one CTA's load phase with another's MMA / epilogue.
// A is never a TMA candidate here because every element has to pass through registers for the split.
#include "common.cuh"

namespace twowl {

constexpr int kTcThreads = 128;
constexpr int kTileM = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major (or MN-major) operands
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int Mdim, int Ndim, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(Ndim >> 3) << 17) | ((uint32_t)(Mdim >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of 16-byte chunk c (0..7) of row r inside one 128-byte-wide SWIZZLE_128B slab
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ r) & 7) << 4)); }

__device__ __forceinline__ void split_tf32(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x, lo.y = v.y - hi.y, lo.z = v.z - hi.z, lo.w = v.w - hi.w;
}

// KD = reduction width (multiple of 32, <= 128). Nd = output width (multiple of 16, <= 256).
// w_kn = 0: B[n][k] = W[n*KD + k] (forward, W = [Co,Ci]);  w_kn = 1: B[n][k] = W[k*Nd + n] (backward input).
template <int KD>
__global__ void __launch_bounds__(kTcThreads) k_linear_tc(const float* __restrict__ A, const float* __restrict__ W,
                                                          float* __restrict__ Cm, int64_t M, int Nd, int w_kn, int tmem_cols) {
  constexpr int KB = KD / 32;        // 128-byte K slabs
  constexpr int K4 = KD / 4;         // 16-byte chunks per row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Bhi = smem;
  uint8_t* Blo = Bhi + (size_t)KD * Nd * 4;
  uint8_t* Ahi = Blo + (size_t)KD * Nd * 4;
  uint8_t* Alo = Ahi + (size_t)kTileM * KD * 4;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Alo + (size_t)kTileM * KD * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + kTileM - 1) / kTileM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident B: hi/lo split of the weight in the canonical layout (slab kb = 32 k-values, Nd rows of 128 B)
  for (int i = tid; i < Nd * K4; i += kTcThreads) {
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
    split_tf32(v, hi, lo);
    const uint32_t off = (uint32_t)(k4 >> 3) * (uint32_t)Nd * 128u + sw128(n, k4 & 7);
    *reinterpret_cast<float4*>(Bhi + off) = hi;
    *reinterpret_cast<float4*>(Blo + off) = lo;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = umma_idesc_tf32(kTileM, Nd, 0, 0);

  float4 pre[K4];  // this thread's 16-byte chunks of the next A tile: chunk index j*128 + tid
  auto prefetch = [&](int64_t tile) {
    const float4* __restrict__ A4 = reinterpret_cast<const float4*>(A) + tile * kTileM * K4;
    const int64_t rows_left = M - tile * kTileM;
#pragma unroll
    for (int j = 0; j < K4; ++j) {
      const int idx = j * kTcThreads + tid;
      pre[j] = (idx / K4 < rows_left) ? ldg_stream(A4 + idx) : f4_zero();
    }
  };

  uint32_t parity = 0;
  int64_t tile = blockIdx.x;
  if (tile < ntiles) prefetch(tile);
  for (; tile < ntiles; tile += gridDim.x) {
    // stage A: split and store swizzled (the previous tile's MMAs were waited for, so the buffers are free)
#pragma unroll
    for (int j = 0; j < K4; ++j) {
      const int idx = j * kTcThreads + tid;
      const int r = idx / K4, k4 = idx % K4;
      float4 hi, lo;
      split_tf32(pre[j], hi, lo);
      const uint32_t off = (uint32_t)(k4 >> 3) * (kTileM * 128u) + sw128(r, k4 & 7);
      *reinterpret_cast<float4*>(Ahi + off) = hi;
      *reinterpret_cast<float4*>(Alo + off) = lo;
    }
    proxy_fence();       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();   // also orders this CTA's tcgen05.ld of the previous tile before the next MMAs
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint8_t* As = (pass == 0) ? Alo : Ahi;
        const uint8_t* Bs = (pass == 1) ? Blo : Bhi;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle atom
            const uint64_t ad = umma_desc(smem_u32(As + (size_t)kb * kTileM * 128) + k * 32, 16, 1024);
            const uint64_t bd = umma_desc(smem_u32(Bs + (size_t)kb * Nd * 128) + k * 32, 16, 1024);
            umma_tf32(tmem_d, ad, bd, idesc, acc);
            acc = 1;
          }
        }
      }
      umma_commit(mbar);  // arrives when every MMA above has finished reading smem and writing TMEM
    }
    const int64_t next = tile + gridDim.x;
    if (next < ntiles) prefetch(next);  // global loads in flight while the tensor core works
    mbar_wait(mbar, parity);
    parity ^= 1;
    tc_fence_after();
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows of the tile; each thread streams its own row
    const int64_t row = tile * kTileM + warp * 32 + lane;
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < Nd; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      if (row < M) {
        float4* dst = reinterpret_cast<float4*>(Cm + row * Nd + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (c0 + q * 4 < Nd) dst[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
}

static int tmem_cols_for(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}
static size_t tc_smem_bytes(int Kd, int Nd) { return 2 * (size_t)Kd * Nd * 4 + 2 * (size_t)kTileM * Kd * 4 + 64 + 1024; }

template <int KD>
static int launch_tc(const float* A, const float* W, float* C, int64_t M, int Nd, int w_kn, cudaStream_t s) {
  const size_t smem = tc_smem_bytes(KD, Nd);
  TW_CUDA(cudaFuncSetAttribute(k_linear_tc<KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = cdiv(M, kTileM);
  int64_t grid = (int64_t)kNumSMs * per_sm;
  if (grid > ntiles) grid = ntiles;
  k_linear_tc<KD><<<(unsigned)grid, kTcThreads, smem, s>>>(A, W, C, M, Nd, w_kn, tmem_cols_for(Nd));
  TW_LAUNCH_CHECK();
  return 0;
}

// C[M,Nd] = A[M,Kd] * B^T on tcgen05 (3xTF32). Returns TWOWL_EINVAL for shapes this kernel does not cover.
int linear_tc(const float* A, const float* W, float* C, int64_t M, int Kd, int Nd, int w_kn, cudaStream_t s) {
  TW_CHECK_ARG(linear_tc_supported(Kd, Nd), "linear (tcgen05): Kd=%d Nd=%d unsupported (Kd%%32, Nd%%16, smem)", Kd, Nd);
  if (M == 0) return 0;
  switch (Kd) {
    case 32: return launch_tc<32>(A, W, C, M, Nd, w_kn, s);
    case 64: return launch_tc<64>(A, W, C, M, Nd, w_kn, s);
    case 96: return launch_tc<96>(A, W, C, M, Nd, w_kn, s);
    case 128: return launch_tc<128>(A, W, C, M, Nd, w_kn, s);
  }
  return TWOWL_EINVAL;
}

bool linear_tc_supported(int Kd, int Nd) {
  return (Kd == 32 || Kd == 64 || Kd == 96 || Kd == 128) && Nd >= 16 && Nd <= 256 && (Nd % 16) == 0 &&
         tc_smem_bytes(Kd, Nd) <= 220 * 1024;
}

}  // namespace twowl
