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

#include "tc_ptx.cuh"

static EncodeTiledFn g_encode = nullptr;
EncodeTiledFn tc_encode_fn() { return g_encode; }

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

// K block (floats per TMA row): 32 / 16 when the channel count is a multiple; other multiples of 8 from 24 up take 32 with a
// partial last block per tap, of which only the K = 8 steps that hold real channels are issued (24 -> 3 steps, 40 -> 4 + 1);
// (48 as 32 + 16 instead of 3 x 16 was measured: slower)
static inline int kb_of(int C) { return C % 32 == 0 ? 32 : (C % 16 == 0 ? 16 : (C % 8 ? 0 : (C >= 24 ? 32 : 8))); }
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
// N tile of the per-tap kernel: up to 128 couts.  (256-cout tiles -- one 128 x 256 accumulator, one CTA per SM, the A tile fetched
// once per 256 couts -- were measured on B200 and are slower: D conv 128->256 forward 81 -> 93 us, C 128->256 at C1b 611 -> 639 us;
// two resident 128-wide CTAs overlap one's epilogue with the other's MMAs.  DCGANSR_TC_BN256=1 re-enables them for experiments.)
static int tc_max_bn() { static int v = getenv("DCGANSR_TC_BN256") ? 256 : 128; return v; }
int tc_bt_rows(int B) { const int b16 = (B + 15) / 16 * 16; return b16 >= 256 ? tc_max_bn() : std::min(b16, 128); }
size_t tc_bt_elems(int ntaps, int A, int B) {
  const int bn = tc_bt_rows(B);
  return (size_t)((B + bn - 1) / bn) * ((size_t)ntaps * A / 32) * bn * 32;
}

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
// the kernel
// ------------------------------------------------------------------------------------------
#define TC_MAXCLS 4
struct TcParams {
  int N, Hg, Wg, Ho, Wo, Co;
  int so, si, Ci;
  int TW, TH, TB, tiles_x, tiles_y;
  int KB, kchunks, BN, nstage;
  int ncls, oy0[TC_MAXCLS], ox0[TC_MAXCLS], ntaps[TC_MAXCLS];      // sub-pixel classes: blockIdx.z / ksplit
  int ksplit;                                                      // K splits per tile: blockIdx.z % ksplit
  int cs;                                                          // cluster size along blockIdx.x (1, 2 or 4): the CTAs of a cluster work on
                                                                   // different pixel tiles and share every weight tile by TMA multicast
  const float* bt[TC_MAXCLS];                                      // pre-tiled weight images (bulk-copy source) or null -> TMA 2-D
  int a_stage_bytes, b_stage_bytes, tmem_cols;
  int act;
  float neg;
  // per tap: box origin offsets (si = 1: dy,dx; si = 2: floor(dy/2), floor(dx/2) and the parities)
  short oy[TC_MAXCLS][DSR_MAX_TAPS], ox[TC_MAXCLS][DSR_MAX_TAPS], py[TC_MAXCLS][DSR_MAX_TAPS], px[TC_MAXCLS][DSR_MAX_TAPS];
};
struct TcMaps { CUtensorMap b[TC_MAXCLS]; };

#define TC_THREADS 192

__global__ void __launch_bounds__(TC_THREADS) tapconv_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ TcMaps mapsB, const TcParams p,
                                                                float* __restrict__ out, float* __restrict__ part, int* __restrict__ counters) {
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
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = blockIdx.z / p.ksplit, split = blockIdx.z % p.ksplit;
  const int ntaps = p.ntaps[cls];
  const int nk_all = ntaps * p.kchunks;
  // this CTA's range of K blocks (split-K: small grids with a long contraction use all SMs; the partial tiles are summed in
  // fixed order by the last CTA of the tile to arrive)
  const int kb_beg = (int)((long long)nk_all * split / p.ksplit), kb_end = (int)((long long)nk_all * (split + 1) / p.ksplit);
  const int nk = kb_end - kb_beg;

  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y; tile /= p.tiles_y;
  const int b0 = tile * p.TB, gy0 = ty * p.TH, gx0 = tx * p.TW;
  const int n0 = blockIdx.y * p.BN;

  // cluster of cs CTAs along x (cs == 1: plain launch): a stage of THIS CTA is written by its own TMA (A tile, 1/cs of the weight
  // tile) and by the peers' multicasts (the other parts of the weight tile), so it is free again only when the MMAs of EVERY CTA
  // of the cluster have read that stage: each issuer's tcgen05.commit arrives on the empty barrier of all cs CTAs
  uint32_t crank = 0;
  if (p.cs > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint16_t cmask = (uint16_t)((1u << p.cs) - 1u);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapsB.b[cls]) : "memory");
    for (int s = 0; s < p.nstage; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), (uint32_t)p.cs); }
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
  if (p.cs > 1) {          // every CTA's barriers are initialised before any peer multicasts into this CTA or signals them
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (uint32_t)(128 * p.KB * 4 + p.BN * p.KB * 4);
      const uint32_t bpart = (uint32_t)(p.BN / p.cs) * (uint32_t)p.KB * 4u;       // this CTA's share of a weight tile (rows crank*BN/cs ...)
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        const int t = kb / p.kchunks, c = kb - t * p.kchunks;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&full[s]);
        mbar_expect_tx(fb, bytes);
        const uint32_t da = smem_u32(sA + (size_t)s * p.a_stage_bytes);
        if (p.si == 1)
          tma_load_4d(da, &mapA, fb, c * p.KB, gx0 + p.ox[cls][t], gy0 + p.oy[cls][t], b0);
        else
          tma_load_5d(da, &mapA, fb, p.px[cls][t] * p.Ci + c * p.KB, gx0 + p.ox[cls][t], p.py[cls][t], gy0 + p.oy[cls][t], b0);
        const uint32_t db = smem_u32(sB + (size_t)s * p.b_stage_bytes);
        if (p.cs > 1) {
          // 1/cs of the weight tile, multicast to the same shared-memory offset (and the same full barrier) of every CTA of the cluster
          if (p.bt[cls])
            bulk_load_mc(db + crank * bpart, p.bt[cls] + ((size_t)blockIdx.y * nk_all + kb) * (p.BN * 32) + (size_t)crank * (bpart / 4), bpart, fb, cmask);
          else
            tma_load_2d_mc(db + crank * bpart, &mapsB.b[cls], fb, t * p.Ci + c * p.KB, n0 + (int)crank * (p.BN / p.cs), cmask);
        } else if (p.bt[cls])    // one contiguous 1-D bulk copy of the whole BN x 32 tile image instead of BN 128-byte TMA rows
          bulk_load(db, p.bt[cls] + ((size_t)blockIdx.y * nk_all + kb) * (p.BN * 32), (uint32_t)(p.BN * 128), fb);
        else
          tma_load_2d(db, &mapsB.b[cls], fb, t * p.Ci + c * p.KB, n0);
        if (++s == p.nstage) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    int kc = kb_beg % p.kchunks;                                              // K chunk inside the tap (no division in the issue loop)
    const int ksteps_full = p.KB >> 3, ksteps_last = (p.Ci - (p.kchunks - 1) * p.KB) >> 3;
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      const int ksteps = kc == p.kchunks - 1 ? ksteps_last : ksteps_full;        // the last K chunk of a tap may be partial
      if (++kc == p.kchunks) kc = 0;
      if (elect_one()) {
        const uint64_t ad = make_kmajor_desc(smem_u32(sA + (size_t)s * p.a_stage_bytes), p.KB);
        const uint64_t bd = make_kmajor_desc(smem_u32(sB + (size_t)s * p.b_stage_bytes), p.KB);
        for (int k = 0; k < ksteps; ++k)
          umma_tf32(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);   // +32 B per K step
        if (p.cs > 1) umma_commit_mc(smem_u32(&empty[s]), cmask);
        else umma_commit(smem_u32(&empty[s]));
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
    bool valid_any = true;
    float* orow = nullptr;
    if (valid) orow = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0[cls]) * p.Wo + gx * p.so + p.ox0[cls]) * p.Co;
    mbar_wait(smem_u32(tmem_full), 0);
    tc_fence_after();
    const bool vec = (p.Co & 3) == 0, vec8 = (p.Co & 7) == 0;
    const float* psrc = nullptr;                 // split-K: the tile's partials, read back by the last CTA
    if (p.ksplit > 1) {
      const long long tile_lin = ((long long)cls * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      float* prow = part + ((tile_lin * p.ksplit + split) * 128 + r) * p.BN;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          __stcg(reinterpret_cast<float4*>(prow + c0 + j), make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                                        __uint_as_float(v[j + 3])));
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const int old = atomicAdd(&counters[tile_lin], 1);
        const int last = old == p.ksplit - 1;
        if (last) counters[tile_lin] = 0;          // self-resetting: ready for the next launch
        *s_last = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!*s_last) valid_any = false;
      else { __threadfence(); psrc = part + (tile_lin * p.ksplit * 128 + r) * p.BN; }
    }
    for (int c0 = 0; valid_any && c0 < p.BN; c0 += 16) {
      uint32_t v[16];
      if (psrc) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = 0.f;
        for (int sp = 0; sp < p.ksplit; ++sp) {     // fixed order: deterministic
          const float* src = psrc + (long long)sp * 128 * p.BN + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(src + j));
            a[j] += x.x; a[j + 1] += x.y; a[j + 2] += x.z; a[j + 3] += x.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(a[j]);
      } else {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      }
      if (valid) {
        const int co = n0 + c0;
        if (vec8) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            if (co + j < p.Co)
              st_global_v8(orow + co + j, act_apply_t(__uint_as_float(v[j]), p.act, p.neg), act_apply_t(__uint_as_float(v[j + 1]), p.act, p.neg),
                           act_apply_t(__uint_as_float(v[j + 2]), p.act, p.neg), act_apply_t(__uint_as_float(v[j + 3]), p.act, p.neg),
                           act_apply_t(__uint_as_float(v[j + 4]), p.act, p.neg), act_apply_t(__uint_as_float(v[j + 5]), p.act, p.neg),
                           act_apply_t(__uint_as_float(v[j + 6]), p.act, p.neg), act_apply_t(__uint_as_float(v[j + 7]), p.act, p.neg));
          }
        } else if (vec) {
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
  if (p.cs > 1) {          // no CTA leaves while a peer may still multicast into its shared memory or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
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

// All classes share the input / output tensors and the iterated grid (Hg, Wg); they differ in taps and output offset.
bool tc_tapconv_multi_ok(const TapGeom* cls, int ncls) {
  if (ncls < 1 || ncls > TC_MAXCLS) return false;
  for (int i = 0; i < ncls; ++i) {
    if (!tc_tapconv_supported(cls[i])) return false;
    if (cls[i].Hg != cls[0].Hg || cls[i].Wg != cls[0].Wg || cls[i].Ci != cls[0].Ci || cls[i].Co != cls[0].Co || cls[i].si != cls[0].si ||
        cls[i].so != cls[0].so)
      return false;
  }
  return true;
}

bool k_tapconv_tc_multi(St st, const TapGeom* cls, int ncls, const float* const* bp, const float* in, float* out, int act, float negval,
                        std::string* err, const float* const* bt) {
  if (!tc_tapconv_multi_ok(cls, ncls)) { if (err) *err = "geometry not supported by the tcgen05 path"; return false; }
  const TapGeom& g = cls[0];
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hg = g.Hg; p.Wg = g.Wg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.so = g.so; p.si = g.si; p.Ci = g.Ci; p.ncls = ncls;
  p.KB = kb_of(g.Ci);
  p.kchunks = (g.Ci + p.KB - 1) / p.KB;
  p.TW = std::min(pow2_ge(g.Wg), 128);
  p.TH = std::min(pow2_ge(g.Hg), 128 / p.TW);
  p.TB = 128 / (p.TW * p.TH);
  p.tiles_x = (g.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (g.Hg + p.TH - 1) / p.TH;
  const int tiles_b = (g.N + p.TB - 1) / p.TB;
  const int co16 = (g.Co + 15) / 16 * 16;
  p.BN = tc_bt_rows(g.Co);
  const int ntiles_n = (g.Co + p.BN - 1) / p.BN;
  p.a_stage_bytes = 128 * p.KB * 4;
  p.b_stage_bytes = (p.BN * p.KB * 4 + 1023) / 1024 * 1024;
  const int stage = p.a_stage_bytes + p.b_stage_bytes;
  // small grids (<= one CTA per SM) take the whole shared memory for a deeper TMA ring; larger ones keep two CTAs per SM
  const long long nctas = (long long)tiles_b * p.tiles_y * p.tiles_x * ntiles_n * ncls;
  const int budget = (nctas <= NSM || p.BN > 128) ? 200 * 1024 : 96 * 1024;
  p.nstage = std::max(2, std::min(8, budget / stage));
  p.tmem_cols = std::max(32, pow2_ge(p.BN));
  // split-K when the grid leaves most SMs idle and the contraction is long
  p.ksplit = 1;
  {
    int nk_min = 1 << 30;
    for (int i = 0; i < ncls; ++i) nk_min = std::min(nk_min, cls[i].ntaps * p.kchunks);
    // (measured on the discriminator layers: pays at <= 1/4 of the SMs, loses at 1/2 -- the partial-tile round trip)
    if (st.ws && st.ws->part && nctas * 4 <= NSM && nk_min >= 16 && !getenv("DCGANSR_NO_SPLITK")) {
      int ks = (int)std::min<long long>(NSM / nctas, 8);
      ks = std::min(ks, nk_min / 8);
      while (ks > 1 && ((size_t)nctas * ks * 128 * p.BN * sizeof(float) > st.ws->part_bytes || nctas > st.ws->ncounters)) --ks;
      p.ksplit = std::max(1, ks);
    }
  }
  // Clusters of 2 / 4 CTAs along x can share the weight tiles by TMA multicast (a cluster of cs moves 16 + 16/cs KB per
  // 128 x 128 x 32 MMA group instead of 32 KB from L2).  Measured on B200 (tests/test_gpu_tiles.py covers the path): no gain --
  // D conv 64->128 forward 113 / 117 / 120 us and C1b FC 256->128 forward 786 / 782 / 826 us at cs = 1 / 2 / 4 -- so the per-tap
  // kernel is not bound by L2 bandwidth but by the bytes it can keep in flight in shared memory (6 stages x 32 KB per SM cover
  // ~0.8 us of L2 latency at full tensor rate).  Off by default; DCGANSR_TC_CLUSTER=2|4 enables it.
  p.cs = 1;
  {
    const long long gx = (long long)tiles_b * p.tiles_y * p.tiles_x;
    int want = 1;
    if (const char* e = getenv("DCGANSR_TC_CLUSTER")) want = atoi(e);
    if (want > 1 && p.ksplit == 1 && gx >= 2 * want && nctas >= 2 * NSM && p.BN % (16 * want) == 0) p.cs = want == 4 ? 4 : 2;
  }
  p.act = act; p.neg = negval;
  for (int i = 0; i < ncls; ++i) {
    p.bt[i] = (bt && bt[i] && p.KB == 32 && p.BN == tc_bt_rows(g.Co)) ? bt[i] : nullptr;
    p.oy0[i] = cls[i].oy0; p.ox0[i] = cls[i].ox0; p.ntaps[i] = cls[i].ntaps;
    for (int t = 0; t < cls[i].ntaps; ++t) {
      if (g.si == 1) { p.oy[i][t] = (short)cls[i].dy[t]; p.ox[i][t] = (short)cls[i].dx[t]; p.py[i][t] = 0; p.px[i][t] = 0; }
      else {
        int fy = floordiv2(cls[i].dy[t]), fx = floordiv2(cls[i].dx[t]);
        p.oy[i][t] = (short)fy; p.ox[i][t] = (short)fx; p.py[i][t] = (short)(cls[i].dy[t] - 2 * fy); p.px[i][t] = (short)(cls[i].dx[t] - 2 * fx);
      }
    }
  }
  // ---- tensor maps ----
  CUtensorMap mapA;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
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
  double flops = 0;
  for (int i = 0; i < ncls; ++i) {
    const cuuint64_t ktot = (cuuint64_t)cls[i].ntaps * g.Ci;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)g.Co};
    cuuint64_t strides[1] = {ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)p.KB, (cuuint32_t)(p.BN / p.cs)};
    r = g_encode(&maps.b[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bp[i], dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 swz_of(p.KB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r); return false; }
    flops += 2.0 * g.N * g.Hg * g.Wg * cls[i].ntaps * g.Ci * g.Co;
  }

  const size_t smem = 1024 + (size_t)p.nstage * stage + (2 * p.nstage + 1) * sizeof(uint64_t) + 32;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem) failed";
      return false;
    }
    configured = true;
  }
  const unsigned gx = (unsigned)(tiles_b * p.tiles_y * p.tiles_x);
  dim3 grid((gx + p.cs - 1) / p.cs * p.cs, (unsigned)ntiles_n, (unsigned)(ncls * p.ksplit));      // padded to whole clusters: the extra CTAs'
                                                                                                 // pixels are out of range (loads zero-filled, stores masked)
  if (p.cs > 1) {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid; lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st.s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    float* partp = st.ws ? st.ws->part : nullptr;
    int* cntp = st.ws ? st.ws->counters : nullptr;
    if (cudaLaunchKernelEx(&lc, tapconv_tc_kernel, mapA, maps, p, out, partp, cntp) != cudaSuccess) {
      if (err) *err = std::string("cluster launch failed: ") + cudaGetErrorString(cudaGetLastError());
      return false;
    }
  } else
    tapconv_tc_kernel<<<grid, TC_THREADS, smem, st.s>>>(mapA, maps, p, out, st.ws ? st.ws->part : nullptr, st.ws ? st.ws->counters : nullptr);
  DSR_LAUNCHED(st, "tapconv_tc", flops, WORK_FLOPS);
  return true;
}

bool k_tapconv_tc(St st, const TapGeom& g, const float* in, const float* bp, float* out, int act, float negval, std::string* err) {
  return k_tapconv_tc_multi(st, &g, 1, &bp, in, out, act, negval, err);
}


// ==========================================================================================
// wgrad on tensor cores.
//
//   acc[t][cp][cq] = sum_pix P[pix][cp] * Q[shift_t(pix)][cq]
//
// per tap t a GEMM  D_t[M = 128 cp][N = cq] += A[K = pixels][M]^T * B_t[K = pixels][N]  with BOTH operands
// MN-major in shared memory: the TMA boxes of the NHWC tensors ([pixels][channels], channels contiguous)
// are used as they land, no transpose.  One CTA owns a 128-channel slice of cp, a slice of cq, a group of
// TG taps (TG * N columns of TMEM, one accumulator per tap) and a range of pixel tiles (split-K); the P tile
// of a pixel tile stays resident in smem while the TG shifted Q tiles stream through a second ring.
// Partials go to scratch[split][cp][t*Cq + cq]; k_wgrad_reduce adds them into the Torch7-layout master
// gradient in a fixed order (deterministic).
// ==========================================================================================
struct WgParams {
  int Cp, Cq, ntaps, s;
  int TW, TH, TB, tiles_x, tiles_y, ntiles, tiles_per_split;
  int KPIX, kbp, kbq, atoms_p, atoms_q, n_mma, TG, tmem_cols;
  int p_stage_bytes, q_stage_bytes, np_stage, nq_stage;
  int q_tiles;                 // number of cq slices (blockIdx.z = mtile * q_tiles + qtile)
  long long split_stride;      // Cp * ntaps * Cq
  short oy[DSR_MAX_TAPS], ox[DSR_MAX_TAPS], py[DSR_MAX_TAPS], px[DSR_MAX_TAPS];
};

// MN-major TF32 shared-memory matrix descriptor.  The only layout the tensor core accepts for MN-major 32-bit
// operands is SWIZZLE_128B_BASE32B (32-byte chunks swizzled inside 128-byte rows, period 4 rows) -- what
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces.  Rows (= K = pixels) are 128 B = 32 channels; 4-row groups
// every SBO = 512 B; the next 32 channels (MN atom) every `lbo` bytes.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, int kb, uint32_t lbo) {
  (void)kb;
  uint32_t sbo = 4u * 128u;
  uint64_t layout = 1ull;                                                     // SWIZZLE_128B_BASE32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(TC_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapP,
                                                              const __grid_constant__ CUtensorMap mapQ, const WgParams p,
                                                              float* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sP = smem;
  uint8_t* sQ = smem + (size_t)p.np_stage * p.p_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + (size_t)p.nq_stage * p.q_stage_bytes);
  uint64_t* p_full = bars;
  uint64_t* p_empty = p_full + p.np_stage;
  uint64_t* q_full = p_empty + p.np_stage;
  uint64_t* q_empty = q_full + p.nq_stage;
  uint64_t* tmem_full = q_empty + p.nq_stage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int t0 = blockIdx.y * p.TG;
  const int tg_n = min(p.TG, p.ntaps - t0);
  const int mtile = blockIdx.z / p.q_tiles, qtile = blockIdx.z % p.q_tiles;
  const int m0 = mtile * 128, q0 = qtile * p.n_mma;
  const int tile_beg = split * p.tiles_per_split;
  const int tile_end = min(p.ntiles, tile_beg + p.tiles_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    for (int i = 0; i < p.np_stage; ++i) { mbar_init(smem_u32(&p_full[i]), 1); mbar_init(smem_u32(&p_empty[i]), 1); }
    for (int i = 0; i < p.nq_stage; ++i) { mbar_init(smem_u32(&q_full[i]), 1); mbar_init(smem_u32(&q_empty[i]), 1); }
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
    if (elect_one()) {
      int ps = 0, qs = 0;
      uint32_t pph = 0, qph = 0;
      const uint32_t p_atom = (uint32_t)(p.KPIX * p.kbp * 4), q_atom = (uint32_t)(p.KPIX * p.kbq * 4);
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int tt = tile;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y; tt /= p.tiles_y;
        const int b0 = tt * p.TB, gy0 = ty * p.TH, gx0 = tx * p.TW;
        mbar_wait(smem_u32(&p_empty[ps]), pph ^ 1);
        const uint32_t pf = smem_u32(&p_full[ps]);
        mbar_expect_tx(pf, p_atom * p.atoms_p);
        for (int a = 0; a < p.atoms_p; ++a)
          tma_load_4d(smem_u32(sP + (size_t)ps * p.p_stage_bytes) + a * p_atom, &mapP, pf, m0 + a * p.kbp, gx0, gy0, b0);
        if (++ps == p.np_stage) { ps = 0; pph ^= 1; }
        for (int tg = 0; tg < tg_n; ++tg) {
          const int t = t0 + tg;
          mbar_wait(smem_u32(&q_empty[qs]), qph ^ 1);
          const uint32_t qf = smem_u32(&q_full[qs]);
          mbar_expect_tx(qf, q_atom * p.atoms_q);
          const uint32_t dq = smem_u32(sQ + (size_t)qs * p.q_stage_bytes);
          for (int a = 0; a < p.atoms_q; ++a) {
            if (p.s == 1)
              tma_load_4d(dq + a * q_atom, &mapQ, qf, q0 + a * p.kbq, gx0 + p.ox[t], gy0 + p.oy[t], b0);
            else
              tma_load_5d(dq + a * q_atom, &mapQ, qf, p.px[t] * p.Cq + q0 + a * p.kbq, gx0 + p.ox[t], p.py[t], gy0 + p.oy[t], b0);
          }
          if (++qs == p.nq_stage) { qs = 0; qph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // D = f32, A = B = tf32, both MN-major (bits 15, 16), N = n_mma, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_mma >> 3) << 17) | ((128u >> 4) << 24);
    int ps = 0, qs = 0;
    uint32_t pph = 0, qph = 0;
    const uint32_t lbo_p = (uint32_t)(p.KPIX * p.kbp * 4), lbo_q = (uint32_t)(p.KPIX * p.kbq * 4);
    const uint32_t kstep_p = (uint32_t)(8 * p.kbp * 4), kstep_q = (uint32_t)(8 * p.kbq * 4);
    const int ksteps = p.KPIX >> 3;
    for (int tile = tile_beg; tile < tile_end; ++tile) {
      mbar_wait(smem_u32(&p_full[ps]), pph);
      const uint32_t pa = smem_u32(sP + (size_t)ps * p.p_stage_bytes);
      for (int tg = 0; tg < tg_n; ++tg) {
        mbar_wait(smem_u32(&q_full[qs]), qph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t qa = smem_u32(sQ + (size_t)qs * p.q_stage_bytes);
          for (int k = 0; k < ksteps; ++k)
            umma_tf32(tmem_base + (uint32_t)(tg * p.n_mma), make_mnmajor_desc(pa + k * kstep_p, p.kbp, lbo_p),
                      make_mnmajor_desc(qa + k * kstep_q, p.kbq, lbo_q), idesc, (tile > tile_beg || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&q_empty[qs]));
          if (tg == tg_n - 1) {
            umma_commit(smem_u32(&p_empty[ps]));
            if (tile == tile_end - 1) umma_commit(smem_u32(tmem_full));
          }
        }
        __syncwarp();
        if (++qs == p.nq_stage) { qs = 0; qph ^= 1; }
      }
      if (++ps == p.np_stage) { ps = 0; pph ^= 1; }
    }
  } else if (tile_end > tile_beg) {
    const int q = warp & 3;
    const int cp = m0 + q * 32 + lane;
    mbar_wait(smem_u32(tmem_full), 0);
    tc_fence_after();
    const int Ntot = p.ntaps * p.Cq;
    float* drow = scratch + (long long)split * p.split_stride + (long long)cp * Ntot;
    for (int tg = 0; tg < tg_n; ++tg) {
      for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tg * p.n_mma + c0), v);
        tmem_ld_wait();
        if (cp < p.Cp) {
          const int cq = q0 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (cq + j + 3 < p.Cq) {
              *reinterpret_cast<float4*>(drow + (t0 + tg) * p.Cq + cq + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (cq + j + e < p.Cq) drow[(t0 + tg) * p.Cq + cq + j + e] = __uint_as_float(v[j + e]);
            }
          }
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

struct WgTcCfg { WgParams p; int S, tap_groups, z; size_t smem; };

static bool wg_tc_cfg(const WgradGeom& g, WgTcCfg& c) {
  if (!g_encode) return false;
  WgParams& p = c.p;
  memset(&p, 0, sizeof(p));
  // 32-channel atoms always; channel tails are TMA zero fill (only 16-byte row strides are required)
  p.kbp = 32; p.kbq = 32;
  if (g.Cp % 4 || g.Cq % 4 || g.Cp < 8 || g.Cq < 8 || g.ntaps < 1) return false;
  if (g.s != 1 && g.s != 2) return false;
  if (g.s == 2 && (g.Hq % 2 || g.Wq % 2)) return false;
  p.Cp = g.Cp; p.Cq = g.Cq; p.ntaps = g.ntaps; p.s = g.s;
  p.KPIX = 64;
  p.TW = std::min(pow2_ge(g.Wp), p.KPIX);
  p.TH = std::min(pow2_ge(g.Hp), p.KPIX / p.TW);
  p.TB = p.KPIX / (p.TW * p.TH);
  p.tiles_x = (g.Wp + p.TW - 1) / p.TW;
  p.tiles_y = (g.Hp + p.TH - 1) / p.TH;
  const int tiles_b = (g.N + p.TB - 1) / p.TB;
  p.ntiles = tiles_b * p.tiles_y * p.tiles_x;
  p.atoms_p = 128 / p.kbp;
  const int cq32 = (g.Cq + 31) / 32 * 32;
  const int mtiles = (g.Cp + 127) / 128;
  // Work decomposition: (cq slice n_mma) x (taps per CTA TG) x (pixel splits S).  Splitting the OUTPUT (smaller n_mma / TG)
  // costs re-reads of the P tile through L2; splitting the pixels costs S partial copies of the gradient that
  // wgrad_reduce has to stream.  A tf32 MMA costs max(64, N/2) tensor cycles, so N = 128 slices lose nothing.
  double best = 1e30;
  int best_n = 0, best_tg = 0, best_s = 1;
  for (int n_mma = std::min(cq32, 256); n_mma >= 32; n_mma >>= 1) {
    if (n_mma > cq32) continue;
    const int q_tiles = (g.Cq + n_mma - 1) / n_mma;
    for (int tg = std::max(1, std::min(g.ntaps, 512 / n_mma)); tg >= 1; tg >>= 1) {
      const int other = ((g.ntaps + tg - 1) / tg) * mtiles * q_tiles;
      for (int S = 1; S <= p.ntiles; S <<= 1) {
        const double ctas = (double)other * S;
        const double waves = std::max(1.0, ctas / NSM);
        const double tiles_cta = (double)(p.ntiles + S - 1) / S;
        const double mma = tiles_cta * tg * (p.KPIX / 8) * std::max(64.0, n_mma / 2.0);                    // tensor cycles per CTA
        const double l2 = tiles_cta * (128.0 * p.KPIX * 4 + (double)tg * n_mma * p.KPIX * 4) / 48.0;       // ~48 B/clk/SM from L2
        const double red = (S + 2.0) * g.Cp * g.Cq * g.ntaps * 4.0 / 2000.0;                               // ~2 KB/clk chip-wide
        const double est = std::max(mma, l2) * waves + 3000.0 + red;
        if (est < best) { best = est; best_n = n_mma; best_tg = tg; best_s = S; }
        if (ctas >= 2 * NSM) break;
      }
    }
  }
  if (best_n == 0) return false;
  p.n_mma = best_n;
  p.q_tiles = (g.Cq + p.n_mma - 1) / p.n_mma;
  p.atoms_q = (p.n_mma + p.kbq - 1) / p.kbq;
  p.TG = best_tg;
  c.tap_groups = (g.ntaps + p.TG - 1) / p.TG;
  p.tmem_cols = std::max(32, pow2_ge(p.TG * p.n_mma));
  p.p_stage_bytes = p.atoms_p * p.KPIX * p.kbp * 4;          // 128 channels x KPIX pixels x 4 B = 32 KB
  p.q_stage_bytes = (p.atoms_q * p.KPIX * p.kbq * 4 + 1023) / 1024 * 1024;
  p.np_stage = 2;
  p.nq_stage = std::max(2, std::min(8, (120 * 1024) / p.q_stage_bytes));
  c.z = mtiles * p.q_tiles;
  int S = std::min(best_s, p.ntiles);
  p.tiles_per_split = (p.ntiles + S - 1) / S;
  c.S = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.split_stride = (long long)g.Cp * g.ntaps * g.Cq;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.s == 1) { p.oy[t] = (short)g.dy[t]; p.ox[t] = (short)g.dx[t]; }
    else {
      int fy = floordiv2(g.dy[t]), fx = floordiv2(g.dx[t]);
      p.oy[t] = (short)fy; p.ox[t] = (short)fx; p.py[t] = (short)(g.dy[t] - 2 * fy); p.px[t] = (short)(g.dx[t] - 2 * fx);
    }
  }
  c.smem = 1024 + (size_t)p.np_stage * p.p_stage_bytes + (size_t)p.nq_stage * p.q_stage_bytes +
           (2 * p.np_stage + 2 * p.nq_stage + 1) * sizeof(uint64_t) + 16;
  return c.smem <= 200 * 1024;
}

bool tc_wgrad_supported(const WgradGeom& g) { WgTcCfg c; return wg_tc_cfg(g, c); }
size_t wgrad_tc_scratch_bytes(const WgradGeom& g) {
  WgTcCfg c;
  if (!wg_tc_cfg(g, c)) return 0;
  return (size_t)c.S * g.Cp * g.ntaps * g.Cq * sizeof(float);
}

bool k_wgrad_tc(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                std::string* err) {
  WgTcCfg c;
  if (!wg_tc_cfg(g, c)) { if (err) *err = "wgrad geometry not supported by the tcgen05 path"; return false; }
  if ((size_t)c.S * g.Cp * g.ntaps * g.Cq * sizeof(float) > scratch_bytes) { if (err) *err = "wgrad scratch too small"; return false; }
  const WgParams& p = c.p;
  CUtensorMap mapP, mapQ;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.Wp, (cuuint64_t)g.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 4, (cuuint64_t)g.Wp * g.Cp * 4, (cuuint64_t)g.Hp * g.Wp * g.Cp * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.kbp, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = g_encode(&mapP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)P, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(P) failed: " + std::to_string((int)r); return false; }
  if (g.s == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cq, (cuuint64_t)g.Wq, (cuuint64_t)g.Hq, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.kbq, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = g_encode(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Cq, (cuuint64_t)g.Wq / 2, 2, (cuuint64_t)g.Hq / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)2 * g.Wq * g.Cq * 4,
                             (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[5] = {(cuuint32_t)p.kbq, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = g_encode(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(Q) failed: " + std::to_string((int)r); return false; }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem) failed";
      return false;
    }
    configured = true;
  }
  dim3 grid((unsigned)c.S, (unsigned)c.tap_groups, (unsigned)c.z);
  wgrad_tc_kernel<<<grid, TC_THREADS, c.smem, st.s>>>(mapP, mapQ, p, scratch);
  DSR_LAUNCHED(st, "wgrad_tc", 2.0 * g.N * g.Hp * g.Wp * g.Cp * g.Cq * g.ntaps, WORK_FLOPS);
  k_wgrad_reduce(st, scratch, c.S, g.Cp, g.Cq, g.ntaps, grad_master);
  return true;
}
