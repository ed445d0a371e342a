// kernels_tc.cu -- tcgen05 / TMA / TMEM implicit-GEMM convolutions for sm_100a (DCGANSR_FAST_TF32).
//
// The tap-list geometry of common.h as a tensor-core GEMM, one CTA per 128-pixel x BN-channel tile:
//
//   D[128 pixels][BN couts] (fp32, TMEM) += A[128 pixels][KB channels of tap t] (smem, K-major, TMA)
//                                         * B[BN couts][KB]                     (smem, K-major, TMA)
//
//   * A is never materialised (no im2col): the 128 rows of a tile are a TB x TH x TW box of the output
//     grid, so for one tap the needed input pixels are ONE TMA box of the NHWC activation tensor at a
//     shifted coordinate; out-of-image taps (padding) are the TMA's zero fill.  Stride-2 gathers use a
//     5-D view [N][H/2][2][W/2][2*C] of the same tensor, so the box stays dense.
//   * warp roles: warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer + TMEM owner,
//     warps 2..5 = epilogue (tcgen05.ld -> activation -> coalesced 16-byte NHWC stores).
//   * smem ring of NSTAGE {A,B} stages with full/empty mbarriers; tcgen05.commit releases a stage.
//   * operands are fp32 in HBM and read by the tensor core as TF32 (kind::tf32); weights are rounded
//     to TF32 (cvt.rna) when packed, accumulation is fp32 in TMEM.
//
// Replaces THCUNN SpatialConvolutionMM / SpatialFullConvolution (im2col + SGEMM) and the cuDNN path of
// cudnn.convert (train.lua:174-179) for nn.SpatialConvolution / nn.SpatialFullConvolution forward and
// updateGradInput (train.lua:99-111).
#include "common.h"

#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#define NSM 148

// ------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no -lcuda: resolved through the runtime)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

bool tc_init(std::string* err) {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
    if (err) *err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
    return false;
  }
  g_encode = (EncodeTiledFn)fn;
  return true;
}

static inline int kb_of(int C) { return C % 32 == 0 ? 32 : (C % 16 == 0 ? 16 : (C % 8 == 0 ? 8 : 0)); }
static inline CUtensorMapSwizzle swz_of(int kb) {
  return kb == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (kb == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

bool tc_tapconv_supported(const TapGeom& g) {
  if (!g_encode) return false;
  if (kb_of(g.Ci) == 0 || g.ntaps < 1) return false;
  if (g.si != 1 && g.si != 2) return false;
  if (g.si == 2 && (g.Hi % 2 || g.Wi % 2)) return false;
  if (g.Hg < 1 || g.Wg < 1) return false;
  return true;
}

// ------------------------------------------------------------------------------------------
// weight packing for the tensor-core path: Bp[b][t*A + a] = tf32(master[a*sa + b*sb + tapidx[t]])
// ------------------------------------------------------------------------------------------
size_t tc_packed_elems(int ntaps, int A, int B) { return (size_t)ntaps * A * B; }

__global__ void pack_taps_tc_kernel(const float* __restrict__ master, float* __restrict__ bp, int ntaps,
                                    const int* __restrict__ tapidx, int A, int B, int64_t sa, int64_t sb) {
  int64_t total = (int64_t)ntaps * A * B;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int a = (int)(i % A);
    int64_t r = i / A;
    int t = (int)(r % ntaps);
    int b = (int)(r / ntaps);
    float v = master[a * sa + b * sb + tapidx[t]];
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    bp[i] = __uint_as_float(u);
  }
}
void k_pack_taps_tc(St st, const float* master, float* bp, int ntaps, const int* tapidx_dev, int A, int B, int64_t sa, int64_t sb) {
  int64_t total = (int64_t)ntaps * A * B;
  int64_t blocks = (total + 255) / 256;
  if (blocks > NSM * 8) blocks = NSM * 8;
  if (blocks < 1) blocks = 1;
  pack_taps_tc_kernel<<<(int)blocks, 256, 0, st.s>>>(master, bp, ntaps, tapidx_dev, A, B, sa, sb);
  DSR_LAUNCHED(st, "pack_taps_tc", 8.0 * total, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (sm_100 "version 1"): rows of KB floats (= the swizzle span),
// 8-row swizzle atoms stacked along M/N every SBO bytes; LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, int kb) {
  uint32_t sbo = 8u * (uint32_t)kb * 4u;                                      // 8 rows * row bytes
  uint64_t layout = kb == 32 ? 2ull : (kb == 16 ? 4ull : 6ull);              // SWIZZLE_128B / 64B / 32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                                                     // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;                                                     // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ float act_apply_t(float v, int act, float neg) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * neg;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
struct TcParams {
  int N, Hg, Wg, Ho, Wo, Co;
  int so, oy0, ox0, si, Ci;
  int TW, TH, TB, tiles_x, tiles_y;
  int ntaps, KB, kchunks, BN, nstage;
  int a_stage_bytes, b_stage_bytes, tmem_cols;
  int act;
  float neg;
  // per tap: box origin offsets (si = 1: dy,dx; si = 2: floor(dy/2), floor(dx/2) and the parities)
  short oy[DSR_MAX_TAPS], ox[DSR_MAX_TAPS], py[DSR_MAX_TAPS], px[DSR_MAX_TAPS];
};

#define TC_THREADS 192

__global__ void __launch_bounds__(TC_THREADS) tapconv_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB, const TcParams p,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.nstage * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.nstage * p.b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstage;
  uint64_t* tmem_full = bars + 2 * p.nstage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.nstage + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.ntaps * p.kchunks;

  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y; tile /= p.tiles_y;
  const int b0 = tile * p.TB, gy0 = ty * p.TH, gx0 = tx * p.TW;
  const int n0 = blockIdx.y * p.BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < p.nstage; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    mbar_init(smem_u32(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(128 * p.KB * 4 + p.BN * p.KB * 4);
      int kb = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        for (int c = 0; c < p.kchunks; ++c, ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, bytes);
          const uint32_t da = smem_u32(sA + (size_t)s * p.a_stage_bytes);
          if (p.si == 1)
            tma_load_4d(da, &mapA, fb, c * p.KB, gx0 + p.ox[t], gy0 + p.oy[t], b0);
          else
            tma_load_5d(da, &mapA, fb, p.px[t] * p.Ci + c * p.KB, gx0 + p.ox[t], p.py[t], gy0 + p.oy[t], b0);
          tma_load_2d(smem_u32(sB + (size_t)s * p.b_stage_bytes), &mapB, fb, (t * p.kchunks + c) * p.KB, n0);
          if (++s == p.nstage) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = make_kmajor_desc(smem_u32(sA + (size_t)s * p.a_stage_bytes), p.KB);
        const uint64_t bd = make_kmajor_desc(smem_u32(sB + (size_t)s * p.b_stage_bytes), p.KB);
        const int ksteps = p.KB >> 3;
        for (int k = 0; k < ksteps; ++k)
          umma_tf32(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);   // +32 B per K step
        umma_commit(smem_u32(&empty[s]));
        if (kb == nk - 1) umma_commit(smem_u32(tmem_full));
      }
      __syncwarp();
      if (++s == p.nstage) { s = 0; ph ^= 1; }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> activation -> NHWC global =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                  // tile row = pixel
    const int w = r % p.TW;
    const int h = (r / p.TW) % p.TH;
    const int b = r / (p.TW * p.TH);
    const int n = b0 + b, gy = gy0 + h, gx = gx0 + w;
    const bool valid = n < p.N && gy < p.Hg && gx < p.Wg;
    float* orow = nullptr;
    if (valid) orow = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0) * p.Wo + gx * p.so + p.ox0) * p.Co;
    mbar_wait(smem_u32(tmem_full), 0);
    tc_fence_after();
    const bool vec = (p.Co & 3) == 0;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (valid) {
        const int co = n0 + c0;
        if (vec) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (co + j < p.Co) {
              float4 o;
              o.x = act_apply_t(__uint_as_float(v[j]), p.act, p.neg);
              o.y = act_apply_t(__uint_as_float(v[j + 1]), p.act, p.neg);
              o.z = act_apply_t(__uint_as_float(v[j + 2]), p.act, p.neg);
              o.w = act_apply_t(__uint_as_float(v[j + 3]), p.act, p.neg);
              *reinterpret_cast<float4*>(orow + co + j) = o;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (co + j < p.Co) orow[co + j] = act_apply_t(__uint_as_float(v[j]), p.act, p.neg);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pow2_ge(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

bool k_tapconv_tc(St st, const TapGeom& g, const float* in, const float* bp, float* out, int act, float negval, std::string* err) {
  if (!tc_tapconv_supported(g)) { if (err) *err = "geometry not supported by the tcgen05 path"; return false; }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hg = g.Hg; p.Wg = g.Wg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.so = g.so; p.oy0 = g.oy0; p.ox0 = g.ox0; p.si = g.si; p.Ci = g.Ci;
  p.KB = kb_of(g.Ci);
  p.kchunks = g.Ci / p.KB;
  p.ntaps = g.ntaps;
  p.TW = std::min(pow2_ge(g.Wg), 128);
  p.TH = std::min(pow2_ge(g.Hg), 128 / p.TW);
  p.TB = 128 / (p.TW * p.TH);
  p.tiles_x = (g.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (g.Hg + p.TH - 1) / p.TH;
  const int tiles_b = (g.N + p.TB - 1) / p.TB;
  const int co16 = (g.Co + 15) / 16 * 16;
  p.BN = std::min(co16, 128);
  const int ntiles_n = (g.Co + p.BN - 1) / p.BN;
  p.a_stage_bytes = 128 * p.KB * 4;
  p.b_stage_bytes = (p.BN * p.KB * 4 + 1023) / 1024 * 1024;
  const int stage = p.a_stage_bytes + p.b_stage_bytes;
  p.nstage = std::max(2, std::min(8, (96 * 1024) / stage));
  p.tmem_cols = std::max(32, pow2_ge(p.BN));
  p.act = act; p.neg = negval;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.si == 1) { p.oy[t] = (short)g.dy[t]; p.ox[t] = (short)g.dx[t]; p.py[t] = 0; p.px[t] = 0; }
    else {
      int fy = floordiv2(g.dy[t]), fx = floordiv2(g.dx[t]);
      p.oy[t] = (short)fy; p.ox[t] = (short)fx; p.py[t] = (short)(g.dy[t] - 2 * fy); p.px[t] = (short)(g.dx[t] - 2 * fx);
    }
  }
  // ---- tensor maps ----
  CUtensorMap mapA, mapB;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  if (g.si == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.KB, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = g_encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swz_of(p.KB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Ci, (cuuint64_t)g.Wi / 2, 2, (cuuint64_t)g.Hi / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)2 * g.Wi * g.Ci * 4,
                             (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[5] = {(cuuint32_t)p.KB, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = g_encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swz_of(p.KB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); return false; }
  {
    const cuuint64_t ktot = (cuuint64_t)g.ntaps * g.Ci;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)g.Co};
    cuuint64_t strides[1] = {ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)p.KB, (cuuint32_t)p.BN};
    r = g_encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bp, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swz_of(p.KB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r); return false; }

  const size_t smem = 1024 + (size_t)p.nstage * stage + (2 * p.nstage + 1) * sizeof(uint64_t) + 16;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem) failed";
      return false;
    }
    configured = true;
  }
  dim3 grid((unsigned)(tiles_b * p.tiles_y * p.tiles_x), (unsigned)ntiles_n);
  tapconv_tc_kernel<<<grid, TC_THREADS, smem, st.s>>>(mapA, mapB, p, out);
  DSR_LAUNCHED(st, "tapconv_tc", 2.0 * g.N * g.Hg * g.Wg * g.ntaps * g.Ci * g.Co, WORK_FLOPS);
  return true;
}
