// Experiment: which smem-descriptor start addresses does tcgen05.mma accept for SWIZZLE_128B K-major and
// SWIZZLE_128B_BASE32B MN-major TF32 operands when the operand window is a SHIFTED view (by whole 128-byte rows)
// of a larger TMA-written tile?  Decides whether a halo tile can be loaded once and reused for every tap.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// generic descriptor
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= layout << 61;
  return d;
}

struct Case { int mode; int shift; int sbo; int use_base; };   // mode 0: K-major A shifted (SW128). mode 1: MN-major A shifted along K (SW128_BASE32B)

#define AROWS 224
// One CTA, 128 threads.  A tile: AROWS rows x 32 floats (TMA, 2 boxes of 112 rows).  B tile: 16 rows x 32 floats (K-major) or
// for mode 1: B MN-major [K = 64 pixels][32 channels].
__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                  const __grid_constant__ CUtensorMap mapA2, const Case* cases, int ncases, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // AROWS * 128 B = 28 KB
  uint8_t* sB = smem + 32 * 1024;              // 16 x 128 B (K-major) ; mode 1: 64 x 128 B
  uint8_t* sA2 = smem + 48 * 1024;             // mode 1 A: AROWS pixel rows x 128 B written with SWIZZLE_128B_ATOM_32B
  uint8_t* sB2 = smem + 80 * 1024;             // mode 1 B: 64 pixel rows x 128 B (ATOM_32B)
  uint64_t* bars = (uint64_t*)(smem + 96 * 1024);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t fb = smem_u32(&bars[0]);
    mbar_expect_tx(fb, AROWS * 128 * 2 + 16 * 128 + 112 * 128);
    tma_load_2d(smem_u32(sA), &mapA, fb, 0, 0);
    tma_load_2d(smem_u32(sA) + 112 * 128, &mapA, fb, 0, 112);
    tma_load_2d(smem_u32(sB), &mapB, fb, 0, 0);
    tma_load_2d(smem_u32(sA2), &mapA2, fb, 0, 0);
    tma_load_2d(smem_u32(sA2) + 112 * 128, &mapA2, fb, 0, 112);
    tma_load_2d(smem_u32(sB2), &mapA2, fb, 0, 0);       // first 64 rows (box is 112 rows; we only use 64) -- same source matrix
  }
  // note: sB2 receives 112 rows (14 KB) -> fits before bars (80K + 14K < 96K)
  mbar_wait(smem_u32(&bars[0]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t ph = 0;
  for (int c = 0; c < ncases; ++c) {
    const Case cs = cases[c];
    if (threadIdx.x == 0) {
      if (cs.mode == 0) {
        // D[128][16] = A[rows shift..][32] * B[16][32]^T ; K-major both, SW128, 4 K-steps of 8
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sA) + cs.shift * 128;
        for (int k = 0; k < 4; ++k) {
          const uint32_t aa = a0 + k * 32;
          const uint32_t bo = cs.use_base ? ((aa >> 7) & 7u) : 0u;
          umma_tf32(tmem, make_desc(aa, 16, cs.sbo, bo, 2), make_desc(smem_u32(sB) + k * 32, 16, 1024, 0, 2), idesc, k > 0);
        }
      } else {
        // D[128 (M = 32 channels x ... only 32 valid)][32] : MN-major.  A = sA2 window of 64 pixel rows starting at `shift`
        // (K = pixels), M = 32 channels (one atom; M=128 needs 4 atoms: we point all 4 atoms at the same data via LBO = 0?) ->
        // use M = 64 instead?  keep M = 128 with LBO = 0 so atoms alias: rows 32..127 replicate rows 0..31.
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sA2) + cs.shift * 128;
        for (int k = 0; k < 8; ++k) {                  // 64 pixels = 8 K-steps of 8 pixels (1024 B each)
          const uint32_t aa = a0 + k * 1024;
          const uint32_t bo = cs.use_base ? ((aa >> 7) & 7u) : 0u;
          umma_tf32(tmem, make_desc(aa, 0, cs.sbo, bo, 1), make_desc(smem_u32(sB2) + k * 1024, 0, 512, 0, 1), idesc, k > 0);
        }
      }
      umma_commit(smem_u32(&bars[1]));
    }
    mbar_wait(smem_u32(&bars[1]), ph);
    ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const int ncol = cs.mode == 0 ? 16 : 32;
    for (int c0 = 0; c0 < ncol; c0 += 16) {
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[((size_t)c * 128 + threadIdx.x) * 32 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

#define CKC(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  g_encode = (EncodeTiledFn)fn;
  // A source: AROWS x 32 small integers; B source: 16 x 32
  std::vector<float> hA(AROWS * 32), hB(16 * 32);
  for (int r = 0; r < AROWS; ++r) for (int k = 0; k < 32; ++k) hA[r * 32 + k] = (float)(((r * 7 + k * 3) % 13) - 6);
  for (int n = 0; n < 16; ++n) for (int k = 0; k < 32; ++k) hB[n * 32 + k] = (float)(((n * 5 + k) % 7) - 3);
  float *dA, *dB, *dOut;
  CKC(cudaMalloc(&dA, hA.size() * 4)); CKC(cudaMalloc(&dB, hB.size() * 4));
  CKC(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  std::vector<Case> cases;
  for (int shift : {0, 1, 2, 3, 5, 8, 9, 13}) for (int ub : {0, 1}) cases.push_back({0, shift, 1024, ub});
  for (int shift : {0, 1, 3}) for (int ub : {0, 1}) cases.push_back({0, shift, 1280, ub});     // 8-row groups every 10 rows
  for (int shift : {0, 2}) for (int ub : {0, 1}) cases.push_back({0, shift, 2048, ub});        // every 16 rows
  for (int shift : {0, 1, 2, 3, 4, 5, 8}) for (int ub : {0, 1}) cases.push_back({1, shift, 512, ub});
  Case* dC;
  CKC(cudaMalloc(&dC, cases.size() * sizeof(Case)));
  CKC(cudaMemcpy(dC, cases.data(), cases.size() * sizeof(Case), cudaMemcpyHostToDevice));
  CKC(cudaMalloc(&dOut, cases.size() * 128 * 32 * 4));
  CKC(cudaMemset(dOut, 0, cases.size() * 128 * 32 * 4));
  CUtensorMap mapA, mapB, mapA2;
  const cuuint32_t ones[2] = {1, 1};
  {
    cuuint64_t dims[2] = {32, AROWS}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {32, 112};
    CUresult r = g_encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; }
    r = g_encode(&mapA2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A2 failed %d\n", (int)r); return 1; }
    cuuint64_t dimsb[2] = {32, 16}; cuuint32_t boxb[2] = {32, 16};
    r = g_encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsb, strides, boxb, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  CKC(cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  exp_kernel<<<1, 128, 98 * 1024>>>(mapA, mapB, mapA2, dC, (int)cases.size(), dOut);
  CKC(cudaDeviceSynchronize());
  std::vector<float> hOut(cases.size() * 128 * 32);
  CKC(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < cases.size(); ++c) {
    const Case cs = cases[c];
    int bad = 0, total = 0;
    if (cs.mode == 0) {
      // expected: row m = group g (m/8), r (m%8): source row = shift + g*(sbo/128) + r
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        int src = cs.shift + (m / 8) * (cs.sbo / 128) + (m % 8);
        if (src >= AROWS) continue;
        float ref = 0;
        for (int k = 0; k < 32; ++k) ref += hA[src * 32 + k] * hB[n * 32 + k];
        ++total;
        if (hOut[(c * 128 + m) * 32 + n] != ref) ++bad;
      }
    } else {
      // D[m][n] = sum_{pix<64} A[shift + pix][m % 32 ...] * B2[pix][n]; B2 = first 64 rows of the A source matrix
      for (int m = 0; m < 32; ++m) for (int n = 0; n < 32; ++n) {
        float ref = 0;
        for (int p = 0; p < 64; ++p) ref += hA[(cs.shift + p) * 32 + m] * hA[p * 32 + n];
        ++total;
        if (hOut[(c * 128 + m) * 32 + n] != ref) ++bad;
      }
    }
    printf("mode %d shift %2d sbo %4d base_offset %s : %s (%d / %d wrong)\n", cs.mode, cs.shift, cs.sbo, cs.use_base ? "set " : "zero", bad ? "WRONG" : "ok", bad, total);
  }
  return 0;
}
