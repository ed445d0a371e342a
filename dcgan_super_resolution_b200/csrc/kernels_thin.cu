// kernels_thin.cu -- bandwidth kernels for the "thin" convolutions of the DCGAN-SR graphs: layers where one side
// has 1..4 channels (the gray / RGB image side: train-gray.lua:104-116 FC 1->64 and C 16->1, train.lua:99,111,121
// FC 3->96, C 12->3, C 3->64, and D's final 512->1 conv train.lua:133).  A GEMM has nothing to offer there
// (N or K of 1..4): the work is one streaming pass over the FAT tensor, so these are plain coalesced
// float4 kernels in exact fp32 (used by both STRICT_FP32 and FAST_TF32).
//
//   thin wgrad:  acc[t][cp][cq] = sum_pix P[pix][cp] * Q[shift_t(pix)][cq]   with min(Cp, Cq) <= 4.
//     The kernel walks the FAT tensor once (rows of one sub-pixel class per thread-slot, channels across
//     lanes as float4) and, per fat pixel, the (tap, thin pixel) pairs that touch it:
//       fat = Q (P thin, full-conv 1->64 / conv 16->1): s*s classes of T/(s*s) pairs,  thin = ((qy - dy_t)/s, ..)
//       fat = P (Q thin, conv 1->64):                   one class of T pairs,          thin = (gy*s + dy_t, ..)
//     Block partials go to scratch[block][cp][t*Cq + cq]; k_wgrad_reduce adds them into the Torch7-layout
//     master gradient in fixed order (deterministic, accumulating like accGradParameters).
#include "common.h"

#include <string.h>

#include <algorithm>

#define NSM 148
#define THIN_MAXP 16      // (tap, thin pixel) pairs per class
#define THIN_MAXCLS 4

struct ThinWg {
  int N, Hf, Wf, Cf;          // fat tensor (NHWC)
  int Ht, Wt, Ct;             // thin tensor
  int cs, ts;                 // fat pixels of a class: (a*cs + cy, b*cs + cx); thin pixel = (a*ts + oy, b*ts + ox)
  int ncls, Ha, Wa;           // class grid extents (upper bound; fy < Hf / fx < Wf is checked)
  int V, PS;                  // float4 lanes per pixel (pow2 >= Cf/4), pixel slots per block
  int T, Cp, Cq, fat_is_p;
  int npairs[THIN_MAXCLS];
  short cy[THIN_MAXCLS], cx[THIN_MAXCLS];
  short oy[THIN_MAXCLS][THIN_MAXP], ox[THIN_MAXCLS][THIN_MAXP], tap[THIN_MAXCLS][THIN_MAXP];
};

static bool thin_wg_cfg(const WgradGeom& g, ThinWg& w) {
  memset(&w, 0, sizeof(w));
  const bool p_thin = g.Cp <= 4, q_thin = g.Cq <= 4;
  if (!p_thin && !q_thin) return false;
  w.T = g.ntaps; w.Cp = g.Cp; w.Cq = g.Cq; w.N = g.N;
  if (g.s < 1 || g.s > 2) return false;
  if (p_thin && (!q_thin || g.Cq >= g.Cp)) {
    // fat = Q (shifted tensor), thin = P (grid tensor)
    w.fat_is_p = 0;
    w.Hf = g.Hq; w.Wf = g.Wq; w.Cf = g.Cq; w.Ht = g.Hp; w.Wt = g.Wp; w.Ct = g.Cp;
    w.cs = g.s; w.ts = 1; w.ncls = g.s * g.s;
    for (int c = 0; c < w.ncls; ++c) {
      const int cy = c / g.s, cx = c % g.s;
      w.cy[c] = (short)cy; w.cx[c] = (short)cx;
      int np = 0;
      for (int t = 0; t < g.ntaps; ++t) {
        const int vy = cy - g.dy[t], vx = cx - g.dx[t];
        if (((vy % g.s) + g.s) % g.s || ((vx % g.s) + g.s) % g.s) continue;
        if (np >= THIN_MAXP) return false;
        // exact division (vy, vx are multiples of s; may be negative)
        w.oy[c][np] = (short)(vy / g.s); w.ox[c][np] = (short)(vx / g.s); w.tap[c][np] = (short)t;
        ++np;
      }
      w.npairs[c] = np;
    }
  } else {
    // fat = P (grid tensor), thin = Q (shifted tensor)
    w.fat_is_p = 1;
    w.Hf = g.Hp; w.Wf = g.Wp; w.Cf = g.Cp; w.Ht = g.Hq; w.Wt = g.Wq; w.Ct = g.Cq;
    w.cs = 1; w.ts = g.s; w.ncls = 1;
    if (g.ntaps > THIN_MAXP) return false;
    for (int t = 0; t < g.ntaps; ++t) { w.oy[0][t] = (short)g.dy[t]; w.ox[0][t] = (short)g.dx[t]; w.tap[0][t] = (short)t; }
    w.npairs[0] = g.ntaps;
  }
  if (w.Cf % 4 || w.Cf < 4 || w.Cf > 1024 || w.Ct < 1 || w.Ct > 4) return false;
  int v = 1;
  while (v < w.Cf / 4) v <<= 1;
  w.V = v; w.PS = 256 / v;
  w.Ha = (w.Hf + w.cs - 1) / w.cs; w.Wa = (w.Wf + w.cs - 1) / w.cs;
  // block reduction buffer: (pairs per group <= 16 / Ct ... 16) * Ct * V float4 <= 40 KB
  const int grp = w.Ct == 1 ? 16 : (w.Ct == 2 ? 8 : 4);
  if ((size_t)grp * w.Ct * w.V * 4 * sizeof(float) > 40 * 1024) return false;
  return true;
}

static int thin_wg_blocks(const ThinWg& w) {
  const int64_t rows = (int64_t)w.N * w.Ha;
  int64_t nb = (rows + w.PS - 1) / w.PS;
  const int64_t cap = std::max(1, NSM * 4 / w.ncls);
  return (int)std::max<int64_t>(1, std::min(nb, cap));
}

bool thin_wgrad_supported(const WgradGeom& g) { ThinWg w; return thin_wg_cfg(g, w); }
size_t thin_wgrad_scratch_bytes(const WgradGeom& g) {
  ThinWg w;
  if (!thin_wg_cfg(g, w)) return 0;
  return (size_t)thin_wg_blocks(w) * g.Cp * g.Cq * g.ntaps * sizeof(float);
}

template <int CT, int MAXP>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const ThinWg w, const float* __restrict__ fat,
                                                          const float* __restrict__ thin, float* __restrict__ scratch) {
  __shared__ __align__(16) float4 red4[MAXP * CT * 256 > 2560 ? 2560 : MAXP * CT * 256];   // <= 40 KB: [pair][ct][V lanes]
  const int cls = blockIdx.y;
  const int j0 = blockIdx.z * MAXP;                       // this block's group of (tap, thin pixel) pairs
  const int np = min(MAXP, w.npairs[cls] - j0);
  if (np <= 0) return;
  const int lane4 = threadIdx.x % w.V;            // float4 lane along the fat channels
  const int slot = threadIdx.x / w.V;             // pixel slot
  const bool lane_ok = lane4 * 4 < w.Cf;
  const int cy = w.cy[cls], cx = w.cx[cls];
  const int nb = (w.Wf - cx + w.cs - 1) / w.cs;   // class pixels per fat row

  float4 acc[MAXP][CT];
#pragma unroll
  for (int j = 0; j < MAXP; ++j)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[j][c] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int64_t rows = (int64_t)w.N * w.Ha;
  for (int64_t row = (int64_t)blockIdx.x * w.PS + slot; row < rows; row += (int64_t)gridDim.x * w.PS) {
    const int n = (int)(row / w.Ha), a = (int)(row % w.Ha);
    const int fy = a * w.cs + cy;
    if (fy >= w.Hf || !lane_ok) continue;
    const float* frow = fat + ((int64_t)(n * w.Hf + fy) * w.Wf + cx) * w.Cf + lane4 * 4;
    const int64_t fstep = (int64_t)w.cs * w.Cf;
    // thin row pointers (null when the thin row is outside the image: zero padding)
    const float* trow[MAXP];
    int tx0[MAXP];
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      trow[j] = nullptr;
      tx0[j] = 0;
      if (j < np) {
        const int ty = a * w.ts + w.oy[cls][j0 + j];
        tx0[j] = w.ox[cls][j0 + j];
        if (ty >= 0 && ty < w.Ht) trow[j] = thin + ((int64_t)(n * w.Ht + ty) * w.Wt) * CT;
      }
    }
    constexpr int UNR = 4;
    for (int b0 = 0; b0 < nb; b0 += UNR) {
      float4 f[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        f[u] = b0 + u < nb ? __ldg(reinterpret_cast<const float4*>(frow + (int64_t)(b0 + u) * fstep)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < MAXP; ++j) {
        if (j < np && trow[j]) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int tx = (b0 + u) * w.ts + tx0[j];
            if (tx >= 0 && tx < w.Wt) {
#pragma unroll
              for (int c = 0; c < CT; ++c) {
                const float tv = __ldg(trow[j] + tx * CT + c);
                acc[j][c].x = fmaf(tv, f[u].x, acc[j][c].x); acc[j][c].y = fmaf(tv, f[u].y, acc[j][c].y);
                acc[j][c].z = fmaf(tv, f[u].z, acc[j][c].z); acc[j][c].w = fmaf(tv, f[u].w, acc[j][c].w);
              }
            }
          }
        }
      }
    }
  }

  // ---- block reduction over the pixel slots (fixed order), then one partial per block ----
  for (int k = 0; k < w.PS; ++k) {
    if (slot == k) {
#pragma unroll
      for (int j = 0; j < MAXP; ++j)
        if (j < np) {
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            float4* d = red4 + (j * CT + c) * w.V + lane4;
            if (k == 0) *d = acc[j][c];
            else { float4 o = *d; o.x += acc[j][c].x; o.y += acc[j][c].y; o.z += acc[j][c].z; o.w += acc[j][c].w; *d = o; }
          }
        }
    }
    __syncthreads();
  }
  // scratch[block][cp][t*Cq + cq]
  const float* red = reinterpret_cast<const float*>(red4);
  float* dst = scratch + (int64_t)blockIdx.x * w.Cp * w.T * w.Cq;
  const int total = np * CT * w.Cf;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int cf = i % w.Cf;
    const int c = (i / w.Cf) % CT;
    const int j = i / (w.Cf * CT);
    const float v = red[((j * CT + c) * w.V) * 4 + cf];
    const int t = w.tap[cls][j0 + j];
    const int cp = w.fat_is_p ? cf : c, cq = w.fat_is_p ? c : cf;
    dst[(int64_t)cp * w.T * w.Cq + t * w.Cq + cq] = v;
  }
}

bool k_wgrad_thin(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch,
                  size_t scratch_bytes) {
  ThinWg w;
  if (!thin_wg_cfg(g, w)) return false;
  const int nb = thin_wg_blocks(w);
  if ((size_t)nb * g.Cp * g.Cq * g.ntaps * sizeof(float) > scratch_bytes) return false;
  const float* fat = w.fat_is_p ? P : Q;
  const float* thin = w.fat_is_p ? Q : P;
  int maxp = 0;
  for (int c = 0; c < w.ncls; ++c) maxp = std::max(maxp, w.npairs[c]);
  // accumulators: MAXP * CT float4 per thread, kept <= 16 (64 registers); more pairs -> pair groups on blockIdx.z
#define THIN_LAUNCH(CT, MP)                                                     \
  do {                                                                          \
    dim3 grid(nb, w.ncls, (maxp + (MP) - 1) / (MP));                            \
    thin_wgrad_kernel<CT, MP><<<grid, 256, 0, st.s>>>(w, fat, thin, scratch);   \
  } while (0)
  switch (w.Ct) {
    case 1: if (maxp <= 4) THIN_LAUNCH(1, 4); else THIN_LAUNCH(1, 16); break;
    case 2: if (maxp <= 4) THIN_LAUNCH(2, 4); else THIN_LAUNCH(2, 8); break;
    case 3: THIN_LAUNCH(3, 4); break;
    default: THIN_LAUNCH(4, 4); break;
  }
#undef THIN_LAUNCH
  DSR_LAUNCHED(st, "wgrad_thin", 4.0 * ((double)g.N * w.Hf * w.Wf * w.Cf + (double)g.N * w.Ht * w.Wt * w.Ct), WORK_BYTES);
  k_wgrad_reduce(st, scratch, nb, g.Cp, g.Cq, g.ntaps, grad_master);
  return true;
}
