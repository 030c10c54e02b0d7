// dw_tc.cu - weight gradients of the two pair-level linear layers in ONE pass on tcgen05 tensor cores:
//
//   [dW_f ; dW_r] (2C x C)  =  [ selfw_f * dO_f | selfw_r * dO_r ]^T (2C x M)  *  H (M x C)
//
// (reference: autograd of GCNConv.lin at TwoWL/model/model.py:37 for conv2s[i] and conv2s_r[i], which share
// their input H - model.py:77). The reduction runs over M = rows of the pair table, so both operands are used
// in their natural row-major form as MN-MAJOR UMMA operands (SWIZZLE_128B_BASE32B, the one layout a transposed tf32
// operand may use; a 16-byte chunk of a row is stored as loaded): a 64-row tile of dO_f/dO_r is the A operand [128 x 64k],
// the same rows of H are the B operand [C x 64k]; tcgen05.mma kind::tf32, 3xTF32 split.
// H is read once for both directions: 3 reads of [M,C] in total.
//
// One persistent warp-specialised CTA per SM (same roles as pair_conv.cu). Accuracy of the long reduction: the
// TMEM accumulator is drained into fp32 registers every kFlush tiles (1024 rows), CTA partials are written to
// global memory and added in double in CTA order by a second kernel: deterministic.
#include "common.cuh"

namespace twowl {

constexpr int kDwProducerWarps = 8;
constexpr int kDwThreads = (kDwProducerWarps + 4 + 1) * 32;  // 416
constexpr int kDwTileK = 64;                                 // rows of the pair table per stage
constexpr int kDwStages = 2;
constexpr int kDwFlush = 16;                                 // tiles accumulated in TMEM between drains

struct DwParams {
  const float* dOf;
  const float* dOr;
  const float* rsf;
  const float* rsr;
  const float* H;
  int64_t M;
  float* part;  // [gridDim.x][128][C]
};

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// MN-major tf32 operands: on sm_100a the tensor core accepts ONLY layout type 1 (SWIZZLE_128B_BASE32B) for a transposed
// tf32 operand - every other layout type makes tcgen05.mma write zeros (measured with tools/mma_probe.cu). The layout,
// as the hardware reads it (same probe): element (mn, k) of an operand lives at byte
//   (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4
// i.e. 512-byte atoms of 4 k-rows x 32 mn-elements whose 32-byte units are XOR-swizzled by the k-row (Swizzle<2,5,2>).
// A 16-byte chunk of a row-major global row is stored as loaded; 8 consecutive chunks of one row fill one 128-byte line,
// so the staging stores of a quarter-warp are bank-conflict free.
constexpr uint32_t kDwSbo = 512u;                                   // next group of 4 k-rows
constexpr uint32_t kDwLbo = (kDwTileK / 4) * kDwSbo;                // next group of 32 mn-elements (8 KB)
constexpr uint32_t kDwKStep = 2u * kDwSbo;                          // one tf32 MMA consumes 8 k-rows
__device__ __forceinline__ uint64_t dw_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((kDwLbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((kDwSbo >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of the 16-byte chunk cc (mn elements 4cc..4cc+3) of k-row rr inside one operand buffer
__device__ __forceinline__ uint32_t dw_off(int rr, int cc) {
  return (uint32_t)(cc >> 3) * kDwLbo + (uint32_t)(rr >> 2) * kDwSbo + (uint32_t)(rr & 3) * 128u +
         (uint32_t)((((cc & 7) >> 1) ^ (rr & 3)) << 5) + (uint32_t)((cc & 1) << 4);
}
__device__ __forceinline__ void dw_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dw_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void dw_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(dw_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void dw_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dw_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dw_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dw_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dw_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void dw_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void dw_split(const float4& v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x, lo.y = v.y - hi.y, lo.z = v.z - hi.z, lo.w = v.w - hi.w;
}

// C = width of dO_f, dO_r and H (32 or 64).
//   A stage: 4 groups of 32 M-elements, hi then lo  (M = 128 = [f cols | r cols | zero padding when C = 32])
//   B stage: C/32 groups hi + C/32 lo
template <int C>
__global__ void __launch_bounds__(kDwThreads, 1) k_dw_tc(const DwParams p) {
  constexpr int C4 = C / 4;                        // 16-byte chunks per row
  constexpr uint32_t kAHalf = 4u * kDwLbo;         // hi (or lo) part of the A stage: 128 M elements = 4 groups
  constexpr uint32_t kBHalf = (uint32_t)(C / 32) * kDwLbo;
  constexpr uint32_t kStage = 2u * kAHalf + 2u * kBHalf;
  constexpr int kChunksSrc = kDwTileK * C4 / (kDwProducerWarps * 32);  // chunks per producer thread per source
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  uint64_t* tfull = bars + 4;
  uint64_t* tempty = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (p.M + kDwTileK - 1) / kDwTileK;
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dw_smem_u32(tmem_slot)), "r"(2 * 64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      dw_mbar_init(&full[s], kDwProducerWarps * 32);
      dw_mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      dw_mbar_init(&tfull[a], 1);
      dw_mbar_init(&tempty[a], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the padding slabs of A (C = 32 only): written once, never touched by the producers
  if (2 * C < 128) {
    for (int st = 0; st < kDwStages; ++st)
      for (int half = 0; half < 2; ++half) {
        float4* z = reinterpret_cast<float4*>(smem + st * kStage + half * kAHalf + (2 * C / 32) * kDwLbo);
        for (int i = tid; i < (int)((4 - 2 * C / 32) * kDwLbo / 16); i += kDwThreads) z[i] = f4_zero();
      }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kDwProducerWarps) {
    // ===================================================== producers
    float4 regs[2][3][kChunksSrc];
    auto issue = [&](int64_t it, float4 (&r)[3][kChunksSrc]) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t rows_left = p.M - tile * kDwTileK;
      const float4* __restrict__ s0 = reinterpret_cast<const float4*>(p.dOf) + tile * kDwTileK * C4;
      const float4* __restrict__ s1 = reinterpret_cast<const float4*>(p.dOr) + tile * kDwTileK * C4;
      const float4* __restrict__ s2 = reinterpret_cast<const float4*>(p.H) + tile * kDwTileK * C4;
#pragma unroll
      for (int j = 0; j < kChunksSrc; ++j) {
        const int idx = j * (kDwProducerWarps * 32) + tid;
        const bool ok = idx / C4 < rows_left;
        r[0][j] = ok ? ldg_stream(s0 + idx) : f4_zero();
        r[1][j] = ok ? ldg_stream(s1 + idx) : f4_zero();
        r[2][j] = ok ? ldg_stream(s2 + idx) : f4_zero();
      }
    };
    auto stage = [&](int64_t it, float4 (&r)[3][kChunksSrc]) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int st = (int)(it % kDwStages);
      dw_mbar_wait(&empty[st], (uint32_t)(((it / kDwStages) & 1) ^ 1));
      uint8_t* Ahi = smem + st * kStage;
      uint8_t* Alo = Ahi + kAHalf;
      uint8_t* Bhi = Alo + kAHalf;
      uint8_t* Blo = Bhi + kBHalf;
      const int64_t row0 = tile * kDwTileK;
#pragma unroll
      for (int j = 0; j < kChunksSrc; ++j) {
        const int idx = j * (kDwProducerWarps * 32) + tid;
        const int rr = idx / C4, c4 = idx % C4;
        const bool ok = row0 + rr < p.M;
        const float scf = ok ? __ldg(p.rsf + row0 + rr) : 0.f;
        const float scr = ok ? __ldg(p.rsr + row0 + rr) : 0.f;
        const uint32_t off = dw_off(rr, c4), off_r = dw_off(rr, c4 + C4);
        float4 v = r[0][j], hi, lo;
        v.x *= scf, v.y *= scf, v.z *= scf, v.w *= scf;
        dw_split(v, hi, lo);
        *reinterpret_cast<float4*>(Ahi + off) = hi;
        *reinterpret_cast<float4*>(Alo + off) = lo;
        v = r[1][j];
        v.x *= scr, v.y *= scr, v.z *= scr, v.w *= scr;
        dw_split(v, hi, lo);
        *reinterpret_cast<float4*>(Ahi + off_r) = hi;
        *reinterpret_cast<float4*>(Alo + off_r) = lo;
        dw_split(r[2][j], hi, lo);
        *reinterpret_cast<float4*>(Bhi + off) = hi;
        *reinterpret_cast<float4*>(Blo + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      dw_mbar_arrive(&full[st]);
    };
    if (my_tiles > 0) issue(0, regs[0]);
    for (int64_t it = 0; it < my_tiles; it += 2) {
      if (it + 1 < my_tiles) issue(it + 1, regs[1]);
      stage(it, regs[0]);
      if (it + 2 < my_tiles) issue(it + 2, regs[0]);
      if (it + 1 < my_tiles) stage(it + 1, regs[1]);
    }
  } else if (warp == 12) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      // D[128, C] (+)= A[128 x 8] * B[C x 8]^T per K-step; A and B both MN-major
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t grp = it / kDwFlush;
        const int a = (int)(grp & 1);
        if (it % kDwFlush == 0) {
          dw_mbar_wait(&tempty[a], (uint32_t)(((grp >> 1) & 1) ^ 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int st = (int)(it % kDwStages);
        dw_mbar_wait(&full[st], (uint32_t)((it / kDwStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint8_t* Ahi = smem + st * kStage;
        const uint8_t* Alo = Ahi + kAHalf;
        const uint8_t* Bhi = Alo + kAHalf;
        const uint8_t* Blo = Bhi + kBHalf;
        uint32_t acc = (it % kDwFlush == 0) ? 0u : 1u;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint8_t* Ap = (pass == 0) ? Alo : Ahi;
          const uint8_t* Bp = (pass == 1) ? Blo : Bhi;
#pragma unroll
          for (int k = 0; k < kDwTileK / 8; ++k) {  // one group of 8 k-rows per K-step
            dw_mma(tmem_base + (uint32_t)(a * 64), dw_desc(dw_smem_u32(Ap) + k * kDwKStep), dw_desc(dw_smem_u32(Bp) + k * kDwKStep), idesc, acc);
            acc = 1;
          }
        }
        dw_commit(&empty[st]);
        if ((it + 1) % kDwFlush == 0 || it + 1 == my_tiles) dw_commit(&tfull[a]);
      }
    }
    __syncwarp();
  } else {
    // ===================================================== drain: warp ew <-> TMEM lanes (= rows of [dW_f; dW_r]) 32ew..32ew+31
    const int ew = warp - kDwProducerWarps;
    float acc[C];
#pragma unroll
    for (int i = 0; i < C; ++i) acc[i] = 0.f;
    const int64_t ngroups = (my_tiles + kDwFlush - 1) / kDwFlush;
    for (int64_t grp = 0; grp < ngroups; ++grp) {
      const int a = (int)(grp & 1);
      dw_mbar_wait(&tfull[a], (uint32_t)((grp >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (uint32_t)(a * 64) + ((uint32_t)(ew * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t v[32];
        dw_tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c0 + i] += __uint_as_float(v[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      dw_mbar_arrive(&tempty[a]);
    }
    float4* dst = reinterpret_cast<float4*>(p.part + ((size_t)blockIdx.x * 128 + ew * 32 + lane) * C);
#pragma unroll
    for (int q = 0; q < C4; ++q) dst[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * 64) : "memory");
}

// dW_f[co][ci] = sum_cta part[cta][co][ci], dW_r = rows C..2C-1, added in CTA order in double
__global__ void k_dw_tc_final(const float* __restrict__ part, int nparts, int C, float* __restrict__ dWf, float* __restrict__ dWr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C * C) return;
  const int row = i / C, col = i % C;
  double s = 0;
  for (int b = 0; b < nparts; ++b) s += (double)part[((size_t)b * 128 + row) * C + col];
  if (row < C) dWf[row * C + col] = (float)s;
  else dWr[(row - C) * C + col] = (float)s;
}

static int dw_grid(int64_t M) {
  const int64_t ntiles = cdiv(M, kDwTileK);
  return (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
}
template <int C>
static size_t dw_smem() {
  const size_t stage = 2 * 4 * (size_t)kDwLbo + 2 * (C / 32) * (size_t)kDwLbo;
  return (size_t)kDwStages * stage + 128 + 1024;
}

template <int C>
static int dw_launch(const DwParams& p, cudaStream_t s) {
  const size_t smem = dw_smem<C>();
  TW_CUDA(cudaFuncSetAttribute(k_dw_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_dw_tc<C><<<dw_grid(p.M), kDwThreads, smem, s>>>(p);
  TW_LAUNCH_CHECK();
  return 0;
}

}  // namespace twowl

using namespace twowl;

extern "C" int twowl_pair_dw_supported(int32_t C) { return (C == 32 || C == 64) ? 1 : 0; }

extern "C" size_t twowl_pair_dw_workspace_bytes(int64_t M, int32_t C) {
  (void)M;
  return align_up((size_t)kNumSMs * 128 * (size_t)C * sizeof(float));
}

extern "C" int twowl_pair_dw(const float* dOf, const float* dOr, const float* rsf, const float* rsr, const float* H, int64_t M,
                             int32_t C, float* dWf, float* dWr, void* ws, size_t ws_bytes, void* stream) {
  TW_CHECK_ARG(C == 32 || C == 64, "pair_dw: C=%d unsupported (32 or 64)", C);
  TW_CHECK_ARG(M > 0, "pair_dw: needs M > 0");
  TW_CHECK_ARG(aligned16(dOf) && aligned16(dOr) && aligned16(H) && rsf && rsr, "pair_dw: bad pointers");
  TW_CHECK_WS(ws_bytes, twowl_pair_dw_workspace_bytes(M, C));
  DwParams p{dOf, dOr, rsf, rsr, H, M, (float*)ws};
  cudaStream_t s = (cudaStream_t)stream;
  int rc = (C == 32) ? dw_launch<32>(p, s) : dw_launch<64>(p, s);
  if (rc) return rc;
  k_dw_tc_final<<<(int)cdiv(2 * C * C, 256), 256, 0, s>>>((const float*)ws, dw_grid(M), C, dWf, dWr);
  TW_LAUNCH_CHECK();
  return 0;
}
