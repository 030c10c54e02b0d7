// mma_probe.cu - standalone probe of how tcgen05.mma kind::tf32 addresses an MN-major shared-memory operand.
// The operand under test is filled with float(word index); the other operand is a K-major SWIZZLE_128B identity
// block, so D reveals, for every (mn, k), WHICH 32-bit word of shared memory the tensor core read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mma_probe tools/mma_probe.cu && gpurun_out/mma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct Variant {
  int test_b;          // 0: A is MN-major under test (B identity K-major); 1: B under test (A identity K-major)
  uint32_t lbo, sbo;   // bytes
  uint32_t layout;     // descriptor layout_type (0 none, 2 = 128B, 4 = 64B, 6 = 32B)
  int N;               // MMA N
  const char* name;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_probe(int test_b, uint32_t lbo, uint32_t sbo, uint32_t layout, int N, int hi_code, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  float* T = reinterpret_cast<float*>(smem + 17408);               // operand under test: 16 KB, value = word index (mod 2048)
  float* I = reinterpret_cast<float*>(smem);       // identity operand, K-major SW128: 128 rows x 128 B = 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16384 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152; i += 128) T[i] = hi_code == 2 ? (1.0f + 0x1p-11f + 0x1p-12f) * ((i & 1) ? -1.f : 1.f) : (float)(hi_code ? (i / 2047) + 1 : (i % 2047) + 1);
  for (int i = tid; i < 4096; i += 128) I[i] = 0.f;
  __syncthreads();
  if (tid < 8) {
    // row r = tid, element k = tid (k < 8): 16-byte chunk c = k/4 stored at chunk (c ^ (r & 7))
    const int r = tid, k = tid, c = k >> 2;
    I[r * 32 + ((c ^ (r & 7)) << 2) + (k & 3)] = 1.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint64_t dT = (uint64_t)((s32(T) & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
                        ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
    const uint64_t dI = (uint64_t)((s32(I) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
                        ((uint64_t)2 << 61);
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (test_b == 1 || test_b == 3) idesc |= (1u << 16);
    if (test_b == 0) idesc |= (1u << 15);
    const uint64_t da = (test_b == 1) ? dI : dT, db = (test_b == 1) ? dT : dI;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar)),
      "r"(0u)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D[128][N]: lane = row, column = n
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[tid * N + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
  const Variant vs[] = {
      {2, 16, 1024, 2, 16, "CONTROL A K-major sw128"},
      {0, 1024, 1024, 1, 16, "A MN sw128_base32 lbo=1024 sbo=1024"},
      {0, 1024, 1024, 4, 16, "A MN sw64 lbo=1024 sbo=1024"},
      {0, 1024, 1024, 6, 16, "A MN sw32 lbo=1024 sbo=1024"},
      {3, 1024, 1024, 2, 16, "A K-major sw128 but B flagged MN (identity symmetric-ish)"},
      {0, 4096, 128, 0, 16, "A MN none lbo=4096 sbo=128"},
      {0, 128, 4096, 0, 16, "A MN none lbo=128 sbo=4096"},
      {0, 256, 128, 0, 16, "A MN none lbo=256 sbo=128"},
      {0, 1024, 1024, 2, 16, "A MN sw128 lbo=1024 sbo=1024"},
      {0, 1024, 2048, 2, 16, "A MN sw128 lbo=1024 sbo=2048"},
      {0, 2048, 1024, 2, 16, "A MN sw128 lbo=2048 sbo=1024"},
      {1, 4096, 128, 0, 64, "B MN none lbo=4096 sbo=128"},
      {1, 128, 4096, 0, 64, "B MN none lbo=128 sbo=4096"},
      {1, 1024, 1024, 2, 64, "B MN sw128 lbo=1024 sbo=1024"},
      {1, 2048, 1024, 2, 64, "B MN sw128 lbo=2048 sbo=1024"},
  };
  float* d_out;
  cudaMalloc(&d_out, 128 * 256 * sizeof(float));
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 216000);
  for (const Variant& v : vs) {
    std::vector<float> h(128 * v.N), h2(128 * v.N);
    for (int pass = 0; pass < 2; ++pass) {
      cudaMemset(d_out, 0, 128 * 256 * sizeof(float));
      k_probe<<<1, 128, 216000>>>(v.test_b, v.lbo, v.sbo, v.layout, v.N, pass, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (pass == 0) printf("== %s : %s\n", v.name, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      cudaMemcpy(pass ? h2.data() : h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    }
    // D[m][n]: test A -> D[m][k<8] = A_hw(m,k);  test B -> D[k<8][n] = B_hw(n,k). Print the BYTE offset read for (mn,k).
    if (v.test_b == 2) {
      k_probe<<<1, 128, 216000>>>(v.test_b, v.lbo, v.sbo, v.layout, v.N, 2, d_out);
      cudaDeviceSynchronize();
      float t[8];
      cudaMemcpy(t, d_out, sizeof(t), cudaMemcpyDeviceToHost);
      printf("  operand 1+2^-11+2^-12 (0.75 ulp_tf32 above 1), sign alternating: D[0][0..3] = %.10f %.10f %.10f %.10f (1.0 = truncation, 1.0009765625 = rounding)\n", t[0], t[1], t[2], t[3]);
    }
    int nz = 0; for (float x : h) nz += (x != 0.f);
    printf("  nonzeros in D: %d\n", nz);
    const int mns[] = {0, 1, 2, 3, 4, 5, 7, 8, 12, 16, 31, 32, 33, 36, 63};
    for (int mn : mns) {
      if (v.test_b == 1 && mn >= v.N) continue;
      printf("  mn=%3d  k0..7 byte offsets:", mn);
      for (int k = 0; k < 8; ++k) {
        const int at = (v.test_b == 1) ? k * v.N + mn : mn * v.N + k;
        const int lo = (int)h[at], hi = (int)h2[at];
        printf(" %6d", (lo == 0 || hi == 0) ? -1 : ((hi - 1) * 2047 + (lo - 1)) * 4);
      }
      printf("\n");
    }
  }
  return 0;
}
