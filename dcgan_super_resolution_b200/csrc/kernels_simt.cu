// kernels_simt.cu -- strict-fp32 (FFMA) implicit-GEMM convolutions for sm_100a.
//
// One kernel family serves nn.SpatialConvolution forward / updateGradInput and
// nn.SpatialFullConvolution forward / updateGradInput (train.lua:99-111) through the tap-list
// geometry of common.h, and one serves both accGradParameters.  This is the DCGANSR_STRICT_FP32
// engine (<= 1e-5 vs the float64 oracle); the tcgen05 engine in kernels_tc.cu shares the geometry.
//
//   tapconv : C[M = N*Hg*Wg pixels][Co] = A[M][K = ntaps*Ci] * W[K][Co]   (A gathered on the fly, NHWC)
//   wgrad   : acc[Cp][ntaps*Cq]         = P^T[Cp][pixels] * Q_shift[pixels][ntaps*Cq], split over pixels,
//             reduced in fixed order (deterministic), added into the Torch7-layout master gradient.
#include "common.h"

#include <algorithm>

#define NSM 148

// ------------------------------------------------------------------------------------------
// weight packing: Wp[t][a][b] = master[a*sa + b*sb + tapidx[t]]
// ------------------------------------------------------------------------------------------
__global__ void pack_taps_kernel(const float* __restrict__ master, float* __restrict__ wp, int ntaps,
                                 const int* __restrict__ tapidx, int A, int B, int64_t sa, int64_t sb) {
  int64_t total = (int64_t)ntaps * A * B;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int b = (int)(i % B);
    int64_t r = i / B;
    int a = (int)(r % A);
    int t = (int)(r / A);
    wp[i] = master[a * sa + b * sb + tapidx[t]];
  }
}
void k_pack_taps(St st, const float* master, float* wp, int ntaps, const int* tapidx_dev, int A, int B,
                 int64_t sa, int64_t sb) {
  int64_t total = (int64_t)ntaps * A * B;
  int64_t blocks = (total + 255) / 256;
  if (blocks > NSM * 8) blocks = NSM * 8;
  if (blocks < 1) blocks = 1;
  pack_taps_kernel<<<(int)blocks, 256, 0, st.s>>>(master, wp, ntaps, tapidx_dev, A, B, sa, sb);
  DSR_LAUNCHED(st, "pack_taps", 8.0 * total, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// all weight packs of a net in ONE launch (after every Adam step): jobs[j] covers flat elements [begin_j, begin_{j+1})
// ------------------------------------------------------------------------------------------
#define PACK_MAXJOBS 128
// x / d and x % d with a shift when d is a power of two (the DCGAN widths 64 ... 512, 16 taps): the runtime divisions were most
// of this kernel's instructions
__device__ __forceinline__ uint32_t fast_divmod(uint32_t x, uint32_t d, uint32_t& rem) {
  if ((d & (d - 1)) == 0) { rem = x & (d - 1); return x >> (__ffs((int)d) - 1); }
  const uint32_t q = x / d;
  rem = x - q * d;
  return q;
}
__global__ void pack_all_kernel(const PackJob* __restrict__ jobs, int njobs, int64_t total) {
  __shared__ int64_t sbeg[PACK_MAXJOBS + 1];
  for (int i = threadIdx.x; i <= njobs && i <= PACK_MAXJOBS; i += blockDim.x) sbeg[i] = jobs[i].begin;     // jobs[njobs] = sentinel
  __syncthreads();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {                       // last job with begin <= i
      int mid = (lo + hi + 1) >> 1;
      if (sbeg[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const PackJob& j = jobs[lo];
    const uint32_t l = (uint32_t)(i - sbeg[lo]);        // a job has < 2^31 elements: 32-bit index arithmetic
    const uint32_t A = (uint32_t)j.A, B = (uint32_t)j.B, T = (uint32_t)j.ntaps;
    const uint32_t sa = (uint32_t)j.sa, sb = (uint32_t)j.sb;      // a master tensor has < 2^31 elements
    if (j.tc == 2) {                        // pre-tiled, pre-swizzled smem images: [n tile][k block][bn rows][32 floats]
      const uint32_t BNr = (uint32_t)j.bn, nk = T * A / 32;
      uint32_t r, kb, a;
      const uint32_t c = l & 31, q = fast_divmod(l >> 5, BNr, r), nt = fast_divmod(q, nk, kb);
      const uint32_t kcol = ((((c >> 2) ^ (r & 7)) << 2) | (c & 3));       // SWIZZLE_128B: 16-byte chunk index ^ (row & 7)
      const uint32_t k = kb * 32 + kcol, t = fast_divmod(k, A, a), n = nt * BNr + r;
      float v = n < B ? j.src[a * sa + n * sb + (uint32_t)j.tapidx[t]] : 0.f;
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      j.dst[l] = __uint_as_float(u);
    } else if (j.tc) {                      // K-major tensor-core pack bp[b][t*A + a], TF32-rounded
      uint32_t a, t;
      const uint32_t r = fast_divmod(l, A, a), b = fast_divmod(r, T, t);
      float v = j.src[a * sa + b * sb + (uint32_t)j.tapidx[t]];
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      j.dst[l] = __uint_as_float(u);
    } else {                                // SIMT / streaming pack wp[t][a][b]
      uint32_t b, a;
      const uint32_t r = fast_divmod(l, B, b), t = fast_divmod(r, A, a);
      j.dst[l] = j.src[a * sa + b * sb + (uint32_t)j.tapidx[t]];
    }
  }
}
void k_pack_all(St st, const PackJob* jobs_dev, int njobs, int64_t total) {
  if (njobs <= 0 || total <= 0 || njobs > PACK_MAXJOBS) return;
  int64_t blocks = (total + 255) / 256;
  if (blocks > NSM * 8) blocks = NSM * 8;
  pack_all_kernel<<<(int)blocks, 256, 0, st.s>>>(jobs_dev, njobs, total);
  DSR_LAUNCHED(st, "pack_all", 8.0 * total, WORK_BYTES);
}

__device__ __forceinline__ float act_apply_s(float v, int act, float neg) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * neg;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------
// tapconv: 256 threads, CTA tile BM pixels x BN couts, K step 16, thread tile TM x TN.
// ------------------------------------------------------------------------------------------
#define TC_BK 16

template <int BM, int BN, int TM, int TN, bool VECA>
__global__ void __launch_bounds__(256) tapconv_simt_kernel(TapGeom g, const float* __restrict__ in, const float* __restrict__ wp,
                                                           float* __restrict__ out, int act, float neg) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int LDA = BM + 4;
  __shared__ __align__(16) float As[TC_BK][LDA];
  __shared__ __align__(16) float Bs[TC_BK][BN];
  __shared__ int4 rowinfo[BM];   // {n or -1, iy0, ix0, output pixel index}

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)g.N * g.Hg * g.Wg;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = g.ntaps * g.Ci;

  for (int r = tid; r < BM; r += 256) {
    int64_t m = m0 + r;
    int4 ri;
    if (m < M) {
      int gx = (int)(m % g.Wg);
      int64_t q = m / g.Wg;
      int gy = (int)(q % g.Hg);
      int n = (int)(q / g.Hg);
      ri.x = n; ri.y = gy * g.si; ri.z = gx * g.si;
      ri.w = (n * g.Ho + gy * g.so + g.oy0) * g.Wo + gx * g.so + g.ox0;
    } else {
      ri.x = -1; ri.y = 0; ri.z = 0; ri.w = 0;
    }
    rowinfo[r] = ri;
  }
  __syncthreads();

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int kv = tid & 3;              // which float4 (4 consecutive k) of the 16-wide K step
  const int rbase = tid >> 2;          // first row handled by this thread; then += 64
  const bool vecB = (g.Co % 4 == 0);

  for (int k0 = 0; k0 < K; k0 += TC_BK) {
    // ---- A tile: gather BM rows x 16 k ----
    {
      const int kk = k0 + kv * 4;
      if (VECA) {
        int t = kk / g.Ci, ci = kk - t * g.Ci;
        const bool kvalid = kk < K;
        const int ddy = kvalid ? g.dy[t] : 0, ddx = kvalid ? g.dx[t] : 0;
#pragma unroll 4
        for (int r = rbase; r < BM; r += 64) {
          int4 ri = rowinfo[r];
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          int iy = ri.y + ddy, ix = ri.z + ddx;
          if (kvalid && ri.x >= 0 && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            v = __ldg(reinterpret_cast<const float4*>(in + ((int64_t)(ri.x * g.Hi + iy) * g.Wi + ix) * g.Ci + ci));
          As[kv * 4 + 0][r] = v.x; As[kv * 4 + 1][r] = v.y; As[kv * 4 + 2][r] = v.z; As[kv * 4 + 3][r] = v.w;
        }
      } else {
        int tt[4], cc[4], dyy[4], dxx[4];
        bool kval[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int k = kk + e;
          kval[e] = k < K;
          tt[e] = kval[e] ? k / g.Ci : 0;
          cc[e] = k - tt[e] * g.Ci;
          dyy[e] = g.dy[tt[e]]; dxx[e] = g.dx[tt[e]];
        }
        for (int r = rbase; r < BM; r += 64) {
          int4 ri = rowinfo[r];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = 0.f;
            int iy = ri.y + dyy[e], ix = ri.z + dxx[e];
            if (kval[e] && ri.x >= 0 && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
              v = __ldg(in + ((int64_t)(ri.x * g.Hi + iy) * g.Wi + ix) * g.Ci + cc[e]);
            As[kv * 4 + e][r] = v;
          }
        }
      }
    }
    // ---- B tile: 16 k x BN couts from Wp[K][Co] ----
    if (vecB) {
      for (int s = tid; s < TC_BK * BN / 4; s += 256) {
        int kr = s / (BN / 4), c4 = s % (BN / 4);
        int k = k0 + kr, co = n0 + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K && co < g.Co) v = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)k * g.Co + co));
        *reinterpret_cast<float4*>(&Bs[kr][c4 * 4]) = v;
      }
    } else {
      for (int s = tid; s < TC_BK * BN; s += 256) {
        int kr = s / BN, c = s % BN;
        int k = k0 + kr, co = n0 + c;
        Bs[kr][c] = (k < K && co < g.Co) ? __ldg(wp + (int64_t)k * g.Co + co) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TC_BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int4 ri = rowinfo[ty * TM + i];
    if (ri.x < 0) continue;
    float* o = out + (int64_t)ri.w * g.Co;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      int co = n0 + tx * TN + j;
      if (vecB && co + 3 < g.Co) {
        float4 v;
        v.x = act_apply_s(acc[i][j], act, neg); v.y = act_apply_s(acc[i][j + 1], act, neg);
        v.z = act_apply_s(acc[i][j + 2], act, neg); v.w = act_apply_s(acc[i][j + 3], act, neg);
        *reinterpret_cast<float4*>(o + co) = v;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (co + e < g.Co) o[co + e] = act_apply_s(acc[i][j + e], act, neg);
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
static void launch_tapconv(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float neg) {
  int64_t M = (int64_t)g.N * g.Hg * g.Wg;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((g.Co + BN - 1) / BN));
  if (g.Ci % 4 == 0)
    tapconv_simt_kernel<BM, BN, TM, TN, true><<<grid, 256, 0, st.s>>>(g, in, wp, out, act, neg);
  else
    tapconv_simt_kernel<BM, BN, TM, TN, false><<<grid, 256, 0, st.s>>>(g, in, wp, out, act, neg);
  DSR_LAUNCHED(st, "tapconv_simt", 2.0 * M * g.ntaps * g.Ci * g.Co, WORK_FLOPS);
}

void k_tapconv_simt(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float negval) {
  int64_t M = (int64_t)g.N * g.Hg * g.Wg;
  if (M <= 0) return;
  if (g.Co <= 16) launch_tapconv<256, 16, 4, 4>(st, g, in, wp, out, act, negval);
  else if (g.Co <= 32) launch_tapconv<256, 32, 8, 4>(st, g, in, wp, out, act, negval);
  else if (g.Co <= 64 || M >= 128 * NSM * 2) launch_tapconv<128, 64, 8, 4>(st, g, in, wp, out, act, negval);
  else launch_tapconv<64, 128, 8, 4>(st, g, in, wp, out, act, negval);
}

// ------------------------------------------------------------------------------------------
// wgrad.  CTA tile 64 (cp) x 64 (flattened (t,cq)), 16 pixels per K step, thread tile 4x4.
// grid = (ceil(T*Cq/64), ceil(Cp/64), S); split s handles pixels [s*chunk, (s+1)*chunk).
// ------------------------------------------------------------------------------------------
struct WgCfg { int S; int64_t chunk; int tiles_n, tiles_m, bm, bn; };

static WgCfg wg_cfg(const WgradGeom& g) {
  WgCfg c;
  int64_t npix = (int64_t)g.N * g.Hp * g.Wp;
  // CTA tile (cp x (t,cq)): 64 x 64, or the narrow variants for few grid-tensor channels (DCGANSR_SIMT_NARROW=0 disables them)
  static const bool narrow = !(getenv("DCGANSR_SIMT_NARROW") && atoi(getenv("DCGANSR_SIMT_NARROW")) == 0);
  c.bm = 64; c.bn = 64;
  // (measured at C3b: Cp = 12: 22.0 ms at 64 rows, 14.1 at 32, 8.5 at 16; but Cp = 48 / 96 lose with narrower tiles -- 22.3 -> 28.3 ->
  //  40.5 ms -- their operand loads are shared by fewer FMAs: narrow tiles only where a 64-row tile would be more than half empty)
  if (narrow && g.Cp <= 16) { c.bm = 16; c.bn = 256; }
  else if (narrow && g.Cp <= 32) { c.bm = 32; c.bn = 128; }
  if (const char* e = getenv("DCGANSR_SIMT_BM")) { const int v = atoi(e); if (v == 16 || v == 32 || v == 64) { c.bm = v; c.bn = 4096 / v; } }
  // 64 x 128 tile with 8 x 4 thread tiles (twice the FMAs per shared-memory read) for the wide layers
  if (c.bm == 64 && c.bn == 64 && g.ntaps * g.Cq >= 128 && !(getenv("DCGANSR_SIMT_BIG") && atoi(getenv("DCGANSR_SIMT_BIG")) == 0)) c.bn = 128;
  c.tiles_n = (g.ntaps * g.Cq + c.bn - 1) / c.bn;
  c.tiles_m = (g.Cp + c.bm - 1) / c.bm;
  int64_t tiles = (int64_t)c.tiles_n * c.tiles_m;
  int64_t want = (NSM * 4 + tiles - 1) / tiles;          // aim at ~4 CTAs per SM
  int64_t max_s = (npix + 255) / 256;                   // at least 256 pixels per split
  if (want > max_s) want = max_s;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  int64_t chunk = (npix + want - 1) / want;
  chunk = (chunk + 15) / 16 * 16;
  c.chunk = chunk;
  c.S = (int)((npix + chunk - 1) / chunk);
  if (c.S < 1) c.S = 1;
  return c;
}

size_t wgrad_simt_scratch_bytes(const WgradGeom& g) {
  WgCfg c = wg_cfg(g);
  return (size_t)c.S * g.Cp * g.ntaps * g.Cq * sizeof(float);
}

template <bool VECP, bool VECQ>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradGeom g, const float* __restrict__ P, const float* __restrict__ Q,
                                                         float* __restrict__ scratch, int64_t chunk) {
  __shared__ __align__(16) float As[16][64];   // [pixel][cp]
  __shared__ __align__(16) float Bs[16][64];   // [pixel][(t,cq)]
  const int tid = threadIdx.x;
  const int Ntot = g.ntaps * g.Cq;
  const int n0 = blockIdx.x * 64, mbase = blockIdx.y * 64;
  const int64_t npix = (int64_t)g.N * g.Hp * g.Wp;
  const int64_t pbeg = (int64_t)blockIdx.z * chunk;
  const int64_t pend = pbeg + chunk < npix ? pbeg + chunk : npix;

  const int lk = tid >> 4;       // pixel slot inside the K step (0..15)
  const int lv = tid & 15;       // float4 slot along the 64-wide tile
  // (t, cq) of this thread's 4 B columns (fixed over the whole loop)
  int bt[4], bc[4];
  bool bval[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int n = n0 + lv * 4 + e;
    bval[e] = n < Ntot;
    bt[e] = bval[e] ? n / g.Cq : 0;
    bc[e] = n - bt[e] * g.Cq;
  }
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = pbeg; p0 < pend; p0 += 16) {
    int64_t p = p0 + lk;
    const bool pv = p < pend;
    int gx = 0, gy = 0, n = 0;
    if (pv) {
      gx = (int)(p % g.Wp);
      int64_t q = p / g.Wp;
      gy = (int)(q % g.Hp);
      n = (int)(q / g.Hp);
    }
    // A: P[p][mbase + lv*4 ..]
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int cp = mbase + lv * 4;
      if (pv) {
        const float* src = P + p * g.Cp + cp;
        if (VECP) {
          if (cp < g.Cp) v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (cp + 0 < g.Cp) v.x = __ldg(src + 0);
          if (cp + 1 < g.Cp) v.y = __ldg(src + 1);
          if (cp + 2 < g.Cp) v.z = __ldg(src + 2);
          if (cp + 3 < g.Cp) v.w = __ldg(src + 3);
        }
      }
      *reinterpret_cast<float4*>(&As[lk][lv * 4]) = v;
    }
    // B: Q[n, gy*s+dy[t], gx*s+dx[t], cq]
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pv) {
        if (VECQ) {
          if (bval[0]) {
            int qy = gy * g.s + g.dy[bt[0]], qx = gx * g.s + g.dx[bt[0]];
            if (qy >= 0 && qy < g.Hq && qx >= 0 && qx < g.Wq)
              v = __ldg(reinterpret_cast<const float4*>(Q + ((int64_t)(n * g.Hq + qy) * g.Wq + qx) * g.Cq + bc[0]));
          }
        } else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (bval[e]) {
              int qy = gy * g.s + g.dy[bt[e]], qx = gx * g.s + g.dx[bt[e]];
              if (qy >= 0 && qy < g.Hq && qx >= 0 && qx < g.Wq)
                e4[e] = __ldg(Q + ((int64_t)(n * g.Hq + qy) * g.Wq + qx) * g.Cq + bc[e]);
            }
          }
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      *reinterpret_cast<float4*>(&Bs[lk][lv * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // partial tile -> scratch[split][cp][n]
  float* dst = scratch + (int64_t)blockIdx.z * g.Cp * Ntot;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int cp = mbase + ty * 4 + i;
    if (cp >= g.Cp) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < Ntot) dst[(int64_t)cp * Ntot + n] = acc[i][j];
    }
  }
}

// Narrow-M variant for few grid-tensor channels (Cp <= 32: the generator's last layers, e.g. C 24->12 with P = dy of 12 channels):
// CTA tile BM (cp) x BN ((t,cq)) with BM * BN = 4096 -- 16 x 256 or 32 x 128 -- instead of 64 x 64, which ran that layer with 19 %
// of its rows in use (22 ms, 3.5 TFLOP/s at C3b).  Same thread tile (4 x 4), K step (16 pixels) and partial layout.
template <int BM, int BN, int TM, bool VECQ>
__global__ void __launch_bounds__(256) wgrad_simt_narrow_kernel(WgradGeom g, const float* __restrict__ P, const float* __restrict__ Q,
                                                                float* __restrict__ scratch, int64_t chunk) {
  static_assert(BM * BN == 1024 * TM && BM % 4 == 0 && BN % 64 == 0 && TM % 4 == 0, "256 threads x (TM x 4)");
  constexpr int NB = BN / 64;                                   // B float4 slots per thread and K step
  constexpr int NA = (16 * (BM / 4) + 255) / 256;               // A float4 slots per thread and K step
  __shared__ __align__(16) float As[16][BM];   // [pixel][cp]
  __shared__ __align__(16) float Bs[16][BN];   // [pixel][(t,cq)]
  const int tid = threadIdx.x;
  const int Ntot = g.ntaps * g.Cq;
  const int n0 = blockIdx.x * BN, mbase = blockIdx.y * BM;
  const int64_t npix = (int64_t)g.N * g.Hp * g.Wp;
  const int64_t pbeg = (int64_t)blockIdx.z * chunk;
  const int64_t pend = pbeg + chunk < npix ? pbeg + chunk : npix;
  // B: slot = tid + j * 256 -> pixel slot / (BN / 4), column group slot % (BN / 4); the (t, cq) of its 4 columns are fixed
  int lkB[NB], lvB[NB], bt[NB][4], bc[NB][4];
  bool bval[NB][4];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int slot = tid + j * 256;
    lkB[j] = slot / (BN / 4); lvB[j] = slot % (BN / 4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n0 + lvB[j] * 4 + e;
      bval[j][e] = n < Ntot;
      bt[j][e] = bval[j][e] ? n / g.Cq : 0;
      bc[j][e] = n - bt[j][e] * g.Cq;
    }
  }
  const int tx = tid % (BN / 4), ty = tid / (BN / 4);
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = pbeg; p0 < pend; p0 += 16) {
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const int slot = tid + j * 256;
      if (slot < 16 * (BM / 4)) {
        const int lkA = slot / (BM / 4), lvA = slot % (BM / 4);
        const int64_t p = p0 + lkA;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int cp = mbase + lvA * 4;
        if (p < pend) {
          const float* src = P + p * g.Cp + cp;
          if (g.Cp % 4 == 0) { if (cp < g.Cp) v = __ldg(reinterpret_cast<const float4*>(src)); }
          else {
            if (cp + 0 < g.Cp) v.x = __ldg(src + 0);
            if (cp + 1 < g.Cp) v.y = __ldg(src + 1);
            if (cp + 2 < g.Cp) v.z = __ldg(src + 2);
            if (cp + 3 < g.Cp) v.w = __ldg(src + 3);
          }
        }
        *reinterpret_cast<float4*>(&As[lkA][lvA * 4]) = v;
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int64_t p = p0 + lkB[j];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < pend) {
        const int gx = (int)(p % g.Wp);
        const int64_t q = p / g.Wp;
        const int gy = (int)(q % g.Hp), n = (int)(q / g.Hp);
        if (VECQ) {
          if (bval[j][0]) {
            const int qy = gy * g.s + g.dy[bt[j][0]], qx = gx * g.s + g.dx[bt[j][0]];
            if (qy >= 0 && qy < g.Hq && qx >= 0 && qx < g.Wq)
              v = __ldg(reinterpret_cast<const float4*>(Q + ((int64_t)(n * g.Hq + qy) * g.Wq + qx) * g.Cq + bc[j][0]));
          }
        } else {
          float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (bval[j][e]) {
              const int qy = gy * g.s + g.dy[bt[j][e]], qx = gx * g.s + g.dx[bt[j][e]];
              if (qy >= 0 && qy < g.Hq && qx >= 0 && qx < g.Wq)
                e4[e] = __ldg(Q + ((int64_t)(n * g.Hq + qy) * g.Wq + qx) * g.Cq + bc[j][e]);
            }
          }
          v = make_float4(e4[0], e4[1], e4[2], e4[3]);
        }
      }
      *reinterpret_cast<float4*>(&Bs[lkB[j]][lvB[j] * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
        av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = scratch + (int64_t)blockIdx.z * g.Cp * Ntot;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int cp = mbase + ty * TM + i;
    if (cp >= g.Cp) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Ntot) dst[(int64_t)cp * Ntot + n] = acc[i][j];
    }
  }
}

// grad_master[cp*(Cq*T) + cq*T + t] += sum_s scratch[s][cp][t*Cq + cq]          (fixed order: deterministic)
// A block owns one cp and a chunk of its T*Cq row.  The S partials are summed by 4 split groups (s = g, g+4, ...) over 64
// element lanes, combined through shared memory in fixed order, and the (t,cq) -> (cq,t) transposition to the Torch7
// layout happens there too, so both the scratch reads and the master-gradient read-modify-write are coalesced.
#define WR_SMEM_FLOATS 10240                              // shared-memory floats: G split groups x (cq chunk x odd-padded taps)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ scratch, int S, int Cp, int Cq, int T, int ncq, int G,
                                                           float* __restrict__ grad_master) {
  __shared__ float part[WR_SMEM_FLOATS];
  const int row = T * Cq;                               // elements per cp
  const int nchunk = (Cq + ncq - 1) / ncq;
  const int cp = blockIdx.x / nchunk, cq0 = (blockIdx.x % nchunk) * ncq;
  const int nc = min(ncq, Cq - cq0);                    // cq range of this block
  const int n = nc * T;
  const int TP = T | 1;                                 // odd row stride: the transposed stores spread over the banks
  const int gstride = ncq * TP;
  // G split groups x E element lanes (few elements and hundreds of partials -> many groups)
  const int E = 256 / G;
  const int lane = threadIdx.x % E, grp = threadIdx.x / E;
  const int64_t total = (int64_t)Cp * row;
  // read order (t major, cq minor): coalesced along cq; stored transposed ([cq][t]) in shared memory.  Partials are added
  // in ascending split order within a group (4 loads in flight), groups in ascending order afterwards: deterministic.
  float* mine = part + grp * gstride;
  if (((nc | cq0 | Cq) & 3) == 0) {
    const int nc4 = nc >> 2, n4 = nc4 * T;
    const int64_t total4 = total >> 2;
    for (int e = lane; e < n4; e += E) {
      const int t = e / nc4, c = (e - t * nc4) << 2;
      const float4* src = reinterpret_cast<const float4*>(scratch + (int64_t)cp * row + (int64_t)t * Cq + cq0 + c);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = grp;
      for (; s + 3 * G < S; s += 4 * G) {
        const float4 x0 = __ldcg(src + (int64_t)s * total4), x1 = __ldcg(src + (int64_t)(s + G) * total4);
        const float4 x2 = __ldcg(src + (int64_t)(s + 2 * G) * total4), x3 = __ldcg(src + (int64_t)(s + 3 * G) * total4);
        a.x += x0.x; a.y += x0.y; a.z += x0.z; a.w += x0.w;
        a.x += x1.x; a.y += x1.y; a.z += x1.z; a.w += x1.w;
        a.x += x2.x; a.y += x2.y; a.z += x2.z; a.w += x2.w;
        a.x += x3.x; a.y += x3.y; a.z += x3.z; a.w += x3.w;
      }
      for (; s < S; s += G) {
        const float4 x = __ldcg(src + (int64_t)s * total4);
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      }
      float* d = mine + c * TP + t;
      d[0] = a.x; d[TP] = a.y; d[2 * TP] = a.z; d[3 * TP] = a.w;
    }
  } else {
    for (int e = lane; e < n; e += E) {
      const int t = e / nc, c = e - t * nc;
      const float* src = scratch + (int64_t)cp * row + (int64_t)t * Cq + cq0 + c;
      float a = 0.f;
      for (int s = grp; s < S; s += G) a += src[(int64_t)s * total];
      mine[c * TP + t] = a;
    }
  }
  __syncthreads();
  float* dst = grad_master + (int64_t)cp * row + (int64_t)cq0 * T;
  for (int o = threadIdx.x; o < n; o += 256) {
    const int c = o / T, i = c * TP + (o - c * T);
    float a = part[i];
    for (int g = 1; g < G; ++g) a += part[g * gstride + i];
    dst[o] += a;
  }
}

void k_wgrad_simt(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master,
                  float* scratch, size_t scratch_bytes) {
  (void)scratch_bytes;
  WgCfg c = wg_cfg(g);
  dim3 grid(c.tiles_n, c.tiles_m, c.S);
  bool vp = g.Cp % 4 == 0, vq = g.Cq % 4 == 0;
  if (c.bm == 16) {
    if (vq) wgrad_simt_narrow_kernel<16, 256, 4, true><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
    else wgrad_simt_narrow_kernel<16, 256, 4, false><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  } else if (c.bm == 32) {
    if (vq) wgrad_simt_narrow_kernel<32, 128, 4, true><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
    else wgrad_simt_narrow_kernel<32, 128, 4, false><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  } else if (c.bn == 128) {
    if (vq) wgrad_simt_narrow_kernel<64, 128, 8, true><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
    else wgrad_simt_narrow_kernel<64, 128, 8, false><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  } else if (vp && vq) wgrad_simt_kernel<true, true><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  else if (vp) wgrad_simt_kernel<true, false><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  else if (vq) wgrad_simt_kernel<false, true><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  else wgrad_simt_kernel<false, false><<<grid, 256, 0, st.s>>>(g, P, Q, scratch, c.chunk);
  DSR_LAUNCHED(st, "wgrad_simt", 2.0 * g.N * g.Hp * g.Wp * g.Cp * g.Cq * g.ntaps, WORK_FLOPS);
  k_wgrad_reduce(st, scratch, c.S, g.Cp, g.Cq, g.ntaps, grad_master);
}

void k_wgrad_reduce(St st, const float* scratch, int S, int Cp, int Cq, int T, float* grad_master) {
  const int64_t row = (int64_t)T * Cq, total = row * Cp;
  if (total <= 0) return;
  const int TP = T | 1;
  const bool v4 = (Cq & 3) == 0;
  // cq values per block: start from everything that fits one split group's share of shared memory at G = 4, then halve
  // while there are fewer blocks than the machine wants (thin layers: few elements, hundreds of partials)
  int ncq = std::max(1, std::min(Cq, WR_SMEM_FLOATS / 4 / TP));
  if (v4 && ncq >= 4) ncq &= ~3;
  while (ncq > 4 && (int64_t)Cp * ((Cq + ncq - 1) / ncq) < 2 * NSM) ncq = v4 ? (((ncq + 1) / 2 + 3) & ~3) : (ncq + 1) / 2;
  const int64_t nchunk = (Cq + ncq - 1) / ncq;
  // element lanes per block (float4 lanes when vectorised) -> split groups: as many as the threads, the partial count and
  // shared memory allow
  const int lanes = std::max(1, (v4 && (ncq & 3) == 0 ? ncq / 4 : ncq) * T);
  int E = 16;
  while (E < 256 && E < lanes) E <<= 1;
  int G = 256 / E;
  while (G > 1 && (G > S || (int64_t)G * ncq * TP > WR_SMEM_FLOATS)) G >>= 1;
  wgrad_reduce_kernel<<<(unsigned)(Cp * nchunk), 256, 0, st.s>>>(scratch, S, Cp, Cq, T, ncq, G, grad_master);
  DSR_LAUNCHED(st, "wgrad_reduce", 4.0 * total * (S + 2), WORK_BYTES);
}
