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

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define NSM 148
template <int A> struct ActC { static constexpr int value = A; };   // compile-time int carried through a generic lambda
#define THIN_MAXP 16      // (tap, thin pixel) pairs per class
#define THIN_MAXCLS 4

struct ThinWg {
  int N, Hf, Wf, Cf;          // fat tensor (NHWC)
  int Ht, Wt, Ct;             // thin tensor
  int cs, ts;                 // fat pixels of a class: (a*cs + cy, b*cs + cx); thin pixel = (a*ts + oy, b*ts + ox)
  int ncls, Ha, Wa;           // class grid extents (upper bound; fy < Hf / fx < Wf is checked)
  int V, PS;                  // float4 lanes per pixel (pow2 >= Cf/4), pixel slots per block
  int seg, nseg;              // a work item = `seg` class pixels of one fat row (rows are split when there are few of them)
  int T, Cp, Cq, fat_is_p;
  int nzg;                    // pair groups per class (set by the launcher)
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
  if ((int64_t)w.N * ((w.Hf + w.cs - 1) / w.cs) * 64 >= (1ll << 31)) return false;      // 32-bit item index (rows x segments)
  int v = 1;
  while (v < w.Cf / 4) v <<= 1;
  w.V = v; w.PS = 256 / v;
  w.Ha = (w.Hf + w.cs - 1) / w.cs; w.Wa = (w.Wf + w.cs - 1) / w.cs;
  // split rows into segments while there are fewer items than thread slots on the machine
  // split rows into segments while there are fewer work items than thread slots on the machine (3 resident blocks per SM)
  // (splitting rows into segments to fill the machine was measured twice, before and after the interior fast path: slower both
  //  times -- C 16->1 wgrad 34 -> 41 us, D conv 1->64 wgrad 27 -> 36 us; narrow layers get deeper load batches instead, UNR below)
  w.seg = w.Wa; w.nseg = 1;
  // (giving the slots of one warp segments of the SAME row, for L1 locality of the thin loads, was slower still: the per-item
  //  setup of up to 16 pair pointers is paid per segment)
  // block reduction buffer: (pairs per group <= 16 / Ct ... 16) * Ct * V float4 <= 48 KB (the static shared-memory limit)
  const int grp = w.Ct == 1 ? 16 : (w.Ct == 2 ? 8 : 4);
  if ((size_t)grp * w.Ct * w.V * 4 * sizeof(float) > 48 * 1024) return false;      // (FC 3->1024 of the C5 generator: exactly 48 KB)
  return true;
}

// blocks per class: every SM full (3 resident blocks at <= 80 registers), never more items than there are
static int thin_wg_blocks(const ThinWg& w) {
  const int64_t items = (int64_t)w.N * w.Ha * w.nseg;
  int64_t nb = (items + w.PS - 1) / w.PS;
  const int resident = w.Ct == 1 ? 3 : (w.Ct <= 3 ? 2 : 1);        // blocks per SM the launch bounds allow (THIN_WG_MINB)
  const int64_t cap = std::max(1, NSM * resident / w.ncls);
  if (nb <= cap) return (int)std::max<int64_t>(1, nb);
  const int64_t rounds = (nb + cap - 1) / cap;             // every block walks the same number of item rounds
  return (int)((nb + rounds - 1) / rounds);
}

bool thin_wgrad_supported(const WgradGeom& g) { ThinWg w; return thin_wg_cfg(g, w); }
size_t thin_wgrad_scratch_bytes(const WgradGeom& g) {
  ThinWg w;
  if (!thin_wg_cfg(g, w)) return 0;
  // + the 3-channel thin tensor padded to 4 channels (one 16-byte load per thin pixel instead of three 4-byte ones)
  const size_t part = ((size_t)thin_wg_blocks(w) * g.Cp * g.Cq * g.ntaps * sizeof(float) + 255) / 256 * 256;
  return part + (w.Ct == 3 ? (size_t)w.N * w.Ht * w.Wt * 4 * sizeof(float) : 0);
}

// 3 -> 4 channel padding of a small NHWC tensor (the RGB side of a thin wgrad)
__global__ void __launch_bounds__(256) pad34_kernel(const float* __restrict__ in, float4* __restrict__ out, int64_t npix) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < npix; i += (int64_t)gridDim.x * 256)
    out[i] = make_float4(__ldg(in + 3 * i), __ldg(in + 3 * i + 1), __ldg(in + 3 * i + 2), 0.f);
}

// 3 resident blocks per SM for the 4-accumulator variant (80 registers, no spills): occupancy is what hides the HBM latency
// of the fat-tensor stream; the bigger variants would spill under that cap.
#define THIN_WG_MINB(CT, MAXP, UNR) ((CT) * (MAXP) <= 4 ? ((UNR) <= 4 ? 3 : 2) : ((CT) * (MAXP) <= 12 ? 2 : 1))
// UNR = fat pixels loaded per thread before they are used
// CTS = floats between thin pixels in memory: CT, or 4 for the padded 3-channel tensor (vector loads: the kernel is bound by the
// L1's sector rate -- 82 % l1tex throughput on C 12->3 wgrad -- and the three 4-byte loads per thin pixel were 86 % of its sectors)
template <int CT, int MAXP, int UNR, int CTS>
__global__ void __launch_bounds__(256, THIN_WG_MINB(CT, MAXP, UNR)) thin_wgrad_kernel(const ThinWg w, const float* __restrict__ fat,
                                                          const float* __restrict__ thin, float* __restrict__ scratch) {
  __shared__ __align__(16) float4 red4[MAXP * CT * 256 > 3072 ? 3072 : MAXP * CT * 256];   // <= 48 KB: [pair][ct][V lanes]
  // blockIdx.x = block * nzg + pair group: the groups of one block index walk the SAME fat rows, so they are neighbours in the
  // launch order and run at the same time -- the fat tensor comes from HBM once and from L2 for the other groups (as blockIdx.z
  // the groups ran one after the other: D conv 3->64 wgrad at C3b read its 268 MB four times, 412 us)
  const int cls = blockIdx.y;
  const int nzg = w.nzg;                                   // pair groups folded into blockIdx.x
  const int bx = blockIdx.x / nzg, nbx = gridDim.x / nzg;
  const int j0 = (int)(blockIdx.x % nzg) * MAXP;           // this block's group of (tap, thin pixel) pairs
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

  const int items = w.N * w.Ha * w.nseg;          // < 2^31 (checked by thin_wg_cfg): 32-bit index arithmetic per item
  auto run = [&](auto tsc) {
  constexpr int TS = decltype(tsc)::value;        // thin pixels per class pixel along x (1 or 2), compile time: immediate offsets
  for (int item = bx * w.PS + slot; item < items; item += nbx * w.PS) {
    const int row = item / w.nseg;
    const int bbeg = (item - row * w.nseg) * w.seg, bend = min(nb, bbeg + w.seg);
    const int n = row / w.Ha, a = row - n * w.Ha;
    const int fy = a * w.cs + cy;
    if (fy >= w.Hf || !lane_ok) continue;
    const float* frow = fat + ((int64_t)(n * w.Hf + fy) * w.Wf + cx) * w.Cf + lane4 * 4;
    const int64_t fstep = (int64_t)w.cs * w.Cf;
    // thin row pointers (null when the thin row is outside the image: zero padding); [bi0, bi1) = class pixels whose thin
    // pixel is inside the row for every live pair (no bounds checks there)
    const float* trow[MAXP];
    int tx0[MAXP];
    int bi0 = bbeg, bi1 = bend;
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      trow[j] = nullptr;
      tx0[j] = 0;
      if (j < np) {
        const int ty = a * w.ts + w.oy[cls][j0 + j];
        tx0[j] = w.ox[cls][j0 + j];
        if (ty >= 0 && ty < w.Ht) {
          trow[j] = thin + ((int64_t)(n * w.Ht + ty) * w.Wt) * CTS;
          bi0 = max(bi0, tx0[j] < 0 ? (-tx0[j] + TS - 1) / TS : 0);
          bi1 = min(bi1, w.Wt - 1 - tx0[j] >= 0 ? (w.Wt - 1 - tx0[j]) / TS + 1 : 0);
        }
      }
    }
    for (int b0 = bbeg; b0 < bend; b0 += UNR) {
      float4 f[UNR];
      if (b0 >= bi0 && b0 + UNR <= bi1) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) f[u] = __ldg(reinterpret_cast<const float4*>(frow + (int64_t)(b0 + u) * fstep));
#pragma unroll
        for (int j = 0; j < MAXP; ++j) {
          if (j < np && trow[j]) {
            const float* tp = trow[j] + (b0 * TS + tx0[j]) * CTS;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
              float tvv[4];
              if (CTS == 4) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(tp + u * TS * 4));
                tvv[0] = t4.x; tvv[1] = t4.y; tvv[2] = t4.z; tvv[3] = t4.w;
              } else {
#pragma unroll
                for (int c = 0; c < CT; ++c) tvv[c] = __ldg(tp + u * TS * CT + c);
              }
#pragma unroll
              for (int c = 0; c < CT; ++c) {
                const float tv = tvv[c];
                acc[j][c].x = fmaf(tv, f[u].x, acc[j][c].x); acc[j][c].y = fmaf(tv, f[u].y, acc[j][c].y);
                acc[j][c].z = fmaf(tv, f[u].z, acc[j][c].z); acc[j][c].w = fmaf(tv, f[u].w, acc[j][c].w);
              }
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        f[u] = b0 + u < bend ? __ldg(reinterpret_cast<const float4*>(frow + (int64_t)(b0 + u) * fstep)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < MAXP; ++j) {
        if (j < np && trow[j]) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int tx = (b0 + u) * TS + tx0[j];
            if (tx >= 0 && tx < w.Wt) {
#pragma unroll
              for (int c = 0; c < CT; ++c) {
                const float tv = __ldg(trow[j] + tx * CTS + c);
                acc[j][c].x = fmaf(tv, f[u].x, acc[j][c].x); acc[j][c].y = fmaf(tv, f[u].y, acc[j][c].y);
                acc[j][c].z = fmaf(tv, f[u].z, acc[j][c].z); acc[j][c].w = fmaf(tv, f[u].w, acc[j][c].w);
              }
            }
          }
        }
      }
    }
  }
  };
  if (w.ts == 1) run(ActC<1>{}); else run(ActC<2>{});

  // ---- block reduction over the pixel slots (fixed order), then one partial per block ----
  // slots that share a warp first (xor butterfly over the slot bits of the lane id), then unit by unit through shared memory
  if (w.V < 32) {
    for (int off = 16; off >= w.V; off >>= 1) {
#pragma unroll
      for (int j = 0; j < MAXP; ++j)
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          acc[j][c].x += __shfl_xor_sync(0xffffffffu, acc[j][c].x, off); acc[j][c].y += __shfl_xor_sync(0xffffffffu, acc[j][c].y, off);
          acc[j][c].z += __shfl_xor_sync(0xffffffffu, acc[j][c].z, off); acc[j][c].w += __shfl_xor_sync(0xffffffffu, acc[j][c].w, off);
        }
    }
  }
  const int units = w.V < 32 ? 8 : w.PS;
  const int unit = w.V < 32 ? (int)(threadIdx.x >> 5) : slot;
  const bool writer = w.V < 32 ? (int)(threadIdx.x & 31) < w.V : true;
  for (int k = 0; k < units; ++k) {
    if (unit == k && writer) {
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
  float* dst = scratch + (int64_t)bx * w.Cp * w.T * w.Cq;
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
  const size_t part = ((size_t)nb * g.Cp * g.Cq * g.ntaps * sizeof(float) + 255) / 256 * 256;
  const char* pe = getenv("DCGANSR_THIN_PAD34");
  // measured in one process (scripts/exp/thin_ab.py): C 12->3 wgrad 460 -> 225 us, FC 3->96 wgrad 303 -> 299 us (fat = the shifted
  // tensor), but D conv 3->64 wgrad 382 -> 439 us (fat = the grid tensor, 16 pairs in 4 groups): padded only in the first form
  const size_t padb = (w.Ct == 3 && !w.fat_is_p && !(pe && atoi(pe) == 0)) ? (size_t)w.N * w.Ht * w.Wt * 4 * sizeof(float) : 0;
  if (part + padb > scratch_bytes) return false;
  const float* fat = w.fat_is_p ? P : Q;
  const float* thin = w.fat_is_p ? Q : P;
  if (padb) {
    float4* tp4 = reinterpret_cast<float4*>(reinterpret_cast<char*>(scratch) + part);
    const int64_t npix = (int64_t)w.N * w.Ht * w.Wt;
    pad34_kernel<<<(unsigned)std::min<int64_t>((npix + 255) / 256, NSM * 8), 256, 0, st.s>>>(thin, tp4, npix);
    thin = reinterpret_cast<const float*>(tp4);
  }
  int maxp = 0;
  for (int c = 0; c < w.ncls; ++c) maxp = std::max(maxp, w.npairs[c]);
  // accumulators: MAXP * CT float4 per thread, kept <= 16 (64 registers); more pairs -> pair groups folded into blockIdx.x
#define THIN_LAUNCH(CT, MP, UNR)                                                    \
  do {                                                                              \
    w.nzg = (maxp + (MP) - 1) / (MP);                                               \
    dim3 grid(nb * w.nzg, w.ncls, 1);                                               \
    if (padb) thin_wgrad_kernel<CT, MP, UNR, 4><<<grid, 256, 0, st.s>>>(w, fat, thin, scratch);  \
    else thin_wgrad_kernel<CT, MP, UNR, CT><<<grid, 256, 0, st.s>>>(w, fat, thin, scratch);      \
  } while (0)
  switch (w.Ct) {
    case 1: THIN_LAUNCH(1, 4, 4); break;        // (deeper load batches, UNR 8 / 16, measured on the 16-channel layer: no change)
    case 2: if (maxp <= 4) THIN_LAUNCH(2, 4, 4); else THIN_LAUNCH(2, 8, 4); break;
    case 3: THIN_LAUNCH(3, 4, 4); break;
    default: THIN_LAUNCH(4, 4, 4); break;
  }
#undef THIN_LAUNCH
  DSR_LAUNCHED(st, "wgrad_thin", 4.0 * ((double)g.N * w.Hf * w.Wf * w.Cf + (double)g.N * w.Ht * w.Wt * w.Ct), WORK_BYTES);
  k_wgrad_reduce(st, scratch, nb, g.Cp, g.Cq, g.ntaps, grad_master);
  return true;
}

// =====================================================================================================================
// thin-INPUT convolution (Ci <= 4): FC 1->64 forward, conv 16->1 dgrad, D's first conv (train-gray.lua:105,116, train.lua:99,121)
//   out[n, gy*so+oy0, gx*so+ox0, co] = act( sum_t sum_ci in[n, gy*si+dy_t, gx*si+dx_t, ci] * W[t][ci][co] )
// Pure output-write streaming: a thread owns one output pixel x 4 couts, the <= 16 x Ci input scalars come through L1,
// the (tiny) weights live in shared memory.  All sub-pixel classes in one launch (blockIdx.y).
// =====================================================================================================================
struct ThinInCls { int Hg, Wg, oy0, ox0, ntaps, gx_lo, gx_hi; short dy[16], dx[16]; const float* wp; };
struct ThinIn {
  int N, Hi, Wi, Ci, Ho, Wo, Co, si, so, ncls, act, vpad, seg, nseg;
  float neg;
  ThinInCls c[4];
};

__device__ __forceinline__ float thin_act(float v, int act, float neg) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * neg;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// activation with the id known at compile time (A >= 0) or at run time (A == -1: tanh / sigmoid, rare)
template <int A>
__device__ __forceinline__ float4 thin_act4(float4 a, int act, float neg) {
  if (A == ACT_NONE) return a;
  if (A == ACT_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); return a; }
  if (A == ACT_LRELU) {
    a.x = a.x > 0.f ? a.x : a.x * neg; a.y = a.y > 0.f ? a.y : a.y * neg;
    a.z = a.z > 0.f ? a.z : a.z * neg; a.w = a.w > 0.f ? a.w : a.w * neg;
    return a;
  }
  a.x = thin_act(a.x, act, neg); a.y = thin_act(a.y, act, neg); a.z = thin_act(a.z, act, neg); a.w = thin_act(a.w, act, neg);
  return a;
}

#ifndef THIN_IN_UNR
#define THIN_IN_UNR 4
#endif
template <int CI, int NT>
__global__ void __launch_bounds__(256) thin_in_kernel(const ThinIn p, const float* __restrict__ in, float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];                  // [t][ci][Co]
  const ThinInCls& c = p.c[blockIdx.y];
  const int wn = c.ntaps * CI * p.Co;
  for (int i = threadIdx.x; i < wn; i += 256) sw[i] = c.wp[i];
  __syncthreads();
  // threads = [row slot][float4 lane over couts]; a slot walks one grid row (n, gy) at a time along gx
  const int V = p.Co >> 2;
  const int VP = p.vpad;                                        // pow2 >= V, divides 256
  const int v = threadIdx.x % VP, slot = threadIdx.x / VP, PS = 256 / VP;
  if (v >= V) return;
  const float4* wv = reinterpret_cast<const float4*>(sw) + v;  // + (t*CI + ci) * V
  // this thread's 4 couts of every tap: in registers when they are few (<= 4 float4), else read from shared memory per use
  constexpr bool WREG = NT * CI <= 4;      // (16 float4 of weights in registers was measured: 157 registers, slower)
  float4 wr[WREG ? NT * CI : 1];
  if (WREG) {
#pragma unroll
    for (int k = 0; k < NT * CI; ++k) wr[k] = k < c.ntaps * CI ? wv[k * V] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // interior columns: every tap inside the image, no bounds checks
  const int gx_lo = c.gx_lo, gx_hi = c.gx_hi;
  const int items = p.N * c.Hg * p.nseg;                       // (row, gx segment)
  auto run = [&](auto actc) {
  constexpr int A = decltype(actc)::value;
  for (int item = blockIdx.x * PS + slot; item < items; item += gridDim.x * PS) {
    const int row = item / p.nseg, gx0 = (item - row * p.nseg) * p.seg, gx1 = min(c.Wg, gx0 + p.seg);
    const int n = row / c.Hg, gy = row - n * c.Hg;
    const float* ir[NT];
    int dxs[NT];
    bool clean = c.ntaps == NT;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      ir[t] = nullptr;
      dxs[t] = 0;
      if (t < c.ntaps) {
        const int iy = gy * p.si + c.dy[t];
        dxs[t] = c.dx[t];
        if (iy >= 0 && iy < p.Hi) ir[t] = in + ((int64_t)(n * p.Hi + iy) * p.Wi) * CI;
        else clean = false;
      }
    }
    const int ostep = p.so * p.Co;
    float* o = out + ((int64_t)(n * p.Ho + gy * p.so + c.oy0) * p.Wo + c.ox0) * p.Co + v * 4 + (int64_t)gx0 * ostep;
    // [gx0, ia) and [ib, gx1): columns where some tap falls outside the image (checked); [ia, ib): interior, every tap inside
    const int ia = clean ? min(max(gx0, gx_lo), gx1) : gx1;
    const int ib = clean ? max(min(gx1, gx_hi + 1), ia) : gx1;
    int gx = gx0;
    auto edge = [&](int gxe) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int ix = gxe * p.si + dxs[t];
        if (ir[t] && ix >= 0 && ix < p.Wi) {
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) {
            const float x = __ldg(ir[t] + ix * CI + ci);
            const float4 w = WREG ? wr[t * CI + ci] : wv[(t * CI + ci) * V];
            acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
          }
        }
      }
      return acc;
    };
    for (; gx < ia; ++gx, o += ostep) *reinterpret_cast<float4*>(o) = thin_act4<A>(edge(gx), p.act, p.neg);
    if (gx < ib) {
      const float* ip[NT];                                      // walking input pointers, one per tap (clean: all NT taps live)
#pragma unroll
      for (int t = 0; t < NT; ++t) ip[t] = ir[t] + (gx * p.si + dxs[t]) * CI;
      const int istep = p.si * CI;
      constexpr int kUnr = NT <= 4 ? THIN_IN_UNR : 1;      // independent accumulator chains in flight per thread (few taps only:
                                                           // 9 / 16 taps double the registers and lose occupancy)
#pragma unroll kUnr
      for (; gx < ib; ++gx, o += ostep) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) {
            const float x = __ldg(ip[t] + ci);
            const float4 w = WREG ? wr[t * CI + ci] : wv[(t * CI + ci) * V];
            acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
          }
          ip[t] += istep;
        }
        *reinterpret_cast<float4*>(o) = thin_act4<A>(acc, p.act, p.neg);
      }
    }
    for (; gx < gx1; ++gx, o += ostep) *reinterpret_cast<float4*>(o) = thin_act4<A>(edge(gx), p.act, p.neg);
  }
  };
  // one uniform branch per thread instead of a jump table per output value
  if (p.act == ACT_NONE) run(ActC<ACT_NONE>{});
  else if (p.act == ACT_RELU) run(ActC<ACT_RELU>{});
  else if (p.act == ACT_LRELU) run(ActC<ACT_LRELU>{});
  else run(ActC<-1>{});
}

// Pixel-per-thread variant for few-channel inputs with more than one channel (RGB: D's first conv 3->64, FC 3->96, the dgrad of
// C 12->3; train.lua:99,111,121).  One thread = one output pixel of one sub-pixel class x CH consecutive couts: the CH weights of a
// (tap, ci) are one conflict-free BROADCAST read of shared memory per warp (every lane the same address) feeding CH FMAs per
// lane, where thin_in_kernel spends one 512-byte shared-memory read per 4 FMAs per lane (it is shared-memory bound at CI = 3:
// D conv 3->64 615 us for 318 MB at C3b).  Lanes are consecutive pixels of a row: coalesced 12 B-strided input reads through L1,
// 4*CH-byte output segments.
template <int CI, int CH, int PX>
__global__ void __launch_bounds__(256) thin_in_px_kernel(const ThinIn p, const float* __restrict__ in, float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];                  // [t][ci][CH] of this (class, cout chunk)
  const int nchunk = p.Co / CH;
  const int cls = blockIdx.y / nchunk, chunk = blockIdx.y - cls * nchunk;
  const ThinInCls& c = p.c[cls];
  for (int i = threadIdx.x; i < c.ntaps * CI * CH; i += 256) {
    const int tc = i / CH, j = i - tc * CH;
    sw[i] = c.wp[(int64_t)tc * p.Co + chunk * CH + j];
  }
  __syncthreads();
  // PX = 2: a thread owns two horizontally adjacent output pixels (every broadcast weight read feeds 2 x 4 FMAs); the class
  // grid's row is walked in pixel pairs (Wg even, checked by the launcher)
  const int Wp = c.Wg / PX;
  const int64_t npix = (int64_t)p.N * c.Hg * Wp;
  auto run = [&](auto actc) {
  constexpr int A = decltype(actc)::value;
  for (int64_t pix = (int64_t)blockIdx.x * 256 + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * 256) {
    int64_t q = pix;
    const int gx = (int)(q % Wp) * PX; q /= Wp;
    const int gy = (int)(q % c.Hg);
    const int n = (int)(q / c.Hg);
    float4 acc[PX][CH / 4];
#pragma unroll
    for (int u = 0; u < PX; ++u)
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) acc[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < c.ntaps; ++t) {
      const int iy = gy * p.si + c.dy[t];
      if (iy < 0 || iy >= p.Hi) continue;
      const float* irow = in + ((int64_t)(n * p.Hi + iy) * p.Wi) * CI;
      const float4* w = reinterpret_cast<const float4*>(sw + t * CI * CH);
      float x[PX][CI];
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const int ix = (gx + u) * p.si + c.dx[t];
        const bool ok = ix >= 0 && ix < p.Wi;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) x[u][ci] = ok ? __ldg(irow + ix * CI + ci) : 0.f;
      }
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
        for (int j = 0; j < CH / 4; ++j) {
          const float4 wv = w[ci * (CH / 4) + j];
#pragma unroll
          for (int u = 0; u < PX; ++u) {
            acc[u][j].x = fmaf(x[u][ci], wv.x, acc[u][j].x); acc[u][j].y = fmaf(x[u][ci], wv.y, acc[u][j].y);
            acc[u][j].z = fmaf(x[u][ci], wv.z, acc[u][j].z); acc[u][j].w = fmaf(x[u][ci], wv.w, acc[u][j].w);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      float* o = out + ((int64_t)(n * p.Ho + gy * p.so + c.oy0) * p.Wo + (gx + u) * p.so + c.ox0) * p.Co + chunk * CH;
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) *reinterpret_cast<float4*>(o + 4 * j) = thin_act4<A>(acc[u][j], p.act, p.neg);
    }
  }
  };
  if (p.act == ACT_NONE) run(ActC<ACT_NONE>{});
  else if (p.act == ACT_RELU) run(ActC<ACT_RELU>{});
  else if (p.act == ACT_LRELU) run(ActC<ACT_LRELU>{});
  else run(ActC<-1>{});
}

bool thin_in_supported(const TapGeom* cls, int ncls) {
  if (ncls < 1 || ncls > 4) return false;
  const TapGeom& g = cls[0];
  if (g.Ci < 1 || g.Ci > 4 || g.Co % 4 || g.Co < 4 || g.Co > 1024) return false;
  for (int i = 0; i < ncls; ++i)
    if (cls[i].ntaps > 16 || (size_t)cls[i].ntaps * g.Ci * g.Co * sizeof(float) > 96 * 1024) return false;
  return true;
}

// wp[i]: SIMT pack [t][ci][co] of class i
bool k_tapconv_thin_in(St st, const TapGeom* cls, int ncls, const float* const* wp, const float* in, float* out, int act, float neg) {
  if (!thin_in_supported(cls, ncls)) return false;
  const TapGeom& g = cls[0];
  ThinIn p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hi = g.Hi; p.Wi = g.Wi; p.Ci = g.Ci; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co; p.si = g.si; p.so = g.so; p.ncls = ncls;
  p.act = act; p.neg = neg;
  size_t smem = 0;
  int64_t maxtot = 0;
  double bytes = 0;
  for (int i = 0; i < ncls; ++i) {
    p.c[i].Hg = cls[i].Hg; p.c[i].Wg = cls[i].Wg; p.c[i].oy0 = cls[i].oy0; p.c[i].ox0 = cls[i].ox0; p.c[i].ntaps = cls[i].ntaps;
    int dxmin = 1 << 20, dxmax = -(1 << 20);
    for (int t = 0; t < cls[i].ntaps; ++t) {
      p.c[i].dy[t] = (short)cls[i].dy[t]; p.c[i].dx[t] = (short)cls[i].dx[t];
      dxmin = std::min(dxmin, cls[i].dx[t]); dxmax = std::max(dxmax, cls[i].dx[t]);
    }
    // gx with every tap column inside [0, Wi): gx*si + dxmin >= 0 and gx*si + dxmax <= Wi - 1
    p.c[i].gx_lo = dxmin < 0 ? (-dxmin + g.si - 1) / g.si : 0;
    p.c[i].gx_hi = (g.Wi - 1 - dxmax) >= 0 ? (g.Wi - 1 - dxmax) / g.si : -1;
    p.c[i].wp = wp[i];
    smem = std::max(smem, (size_t)cls[i].ntaps * g.Ci * g.Co * sizeof(float));
    maxtot = std::max<int64_t>(maxtot, (int64_t)g.N * cls[i].Hg * cls[i].Wg * (g.Co / 4));
    bytes += 4.0 * g.N * cls[i].Hg * cls[i].Wg * g.Co;
  }
  if (maxtot <= 0) return true;
  bytes += 4.0 * g.N * g.Hi * g.Wi * g.Ci;
  // Long per-output contractions (taps x Ci >= 16: the stride-2 convs, where thin_in_kernel is bound by its shared-memory weight
  // reads) and few-cout RGB layers take the pixel-per-thread kernel; measured on B200 at C3b / C2 sizes: D conv 3->64 forward
  // 312 -> 171 us, C 12->3 dgrad 460 -> 314 us, D conv 1->64 25 -> 17 us.  Short contractions with many couts stay on
  // thin_in_kernel, whose lanes run along the couts (FC 3->96 forward 287 us vs 471, FC 1->64 48 vs 99: output-write bound).
  // DCGANSR_THIN_IN_PX=0 / 1 forces the choice for A/B runs.
  {
    int maxt1 = 1;
    for (int i = 0; i < ncls; ++i) maxt1 = std::max(maxt1, cls[i].ntaps);
    const char* e = getenv("DCGANSR_THIN_IN_PX");
    const bool px = e ? atoi(e) != 0 : (maxt1 * g.Ci >= 16 || (g.Ci >= 2 && g.Co <= 16));
    const int CH = g.Co % 16 == 0 ? 16 : (g.Co % 12 == 0 ? 12 : (g.Co % 8 == 0 ? 8 : 4));
    if (px) {
      int64_t maxpix = 0;
      int maxt2 = 1;
      for (int i = 0; i < ncls; ++i) { maxpix = std::max<int64_t>(maxpix, (int64_t)g.N * cls[i].Hg * cls[i].Wg); maxt2 = std::max(maxt2, cls[i].ntaps); }
      // two pixels per thread for the stride-2 gathers when every class grid has an even width (measured: D conv 3->64 forward
      // 323 -> 300 us, D conv 1->64 29 -> 25 us; not for the stride-1 classes: C 12->3 dgrad 309 -> 407 us).  Also measured and
      // dropped: the cout chunks of a pixel on consecutive threads instead of on different blocks (whole-row writes, but four
      // distinct weight addresses per warp: 304 -> 765 us); a 4-channel padded copy of the 3-channel input for 16-byte loads (what
      // doubled the thin wgrad): D conv 3->64 forward 302 -> 341 us, C 12->3 dgrad 304 -> 319 us
      bool even = true;
      for (int i = 0; i < ncls; ++i) even = even && cls[i].Wg % 2 == 0;
      const char* e2 = getenv("DCGANSR_THIN_PX2");
      const int PXv = (even && g.si == 2 && !(e2 && atoi(e2) == 0)) ? 2 : 1;
      dim3 grid2((unsigned)std::min<int64_t>((maxpix / PXv + 255) / 256, NSM * 8), (unsigned)(ncls * (g.Co / CH)));
      const size_t smem2 = (size_t)maxt2 * g.Ci * CH * sizeof(float);
#define THIN_PX_LAUNCH(CI_, CH_)                                                                       \
      do {                                                                                             \
        if (PXv == 2) thin_in_px_kernel<CI_, CH_, 2><<<grid2, 256, smem2, st.s>>>(p, in, out);         \
        else thin_in_px_kernel<CI_, CH_, 1><<<grid2, 256, smem2, st.s>>>(p, in, out);                  \
      } while (0)
#define THIN_PX_CH(CI_)                                         \
      do {                                                      \
        if (CH == 16) THIN_PX_LAUNCH(CI_, 16);                  \
        else if (CH == 12) THIN_PX_LAUNCH(CI_, 12);             \
        else if (CH == 8) THIN_PX_LAUNCH(CI_, 8);               \
        else THIN_PX_LAUNCH(CI_, 4);                            \
      } while (0)
      switch (g.Ci) {
        case 1: THIN_PX_CH(1); break;
        case 2: THIN_PX_CH(2); break;
        case 3: THIN_PX_CH(3); break;
        default: THIN_PX_CH(4); break;
      }
#undef THIN_PX_CH
#undef THIN_PX_LAUNCH
      DSR_LAUNCHED(st, "tapconv_thin_in", bytes, WORK_BYTES);
      return true;
    }
  }
  int vp = 1;
  while (vp < g.Co / 4) vp <<= 1;
  p.vpad = vp;
  int maxrows = 1, maxt = 1;
  for (int i = 0; i < ncls; ++i) { maxrows = std::max(maxrows, g.N * cls[i].Hg); maxt = std::max(maxt, cls[i].ntaps); }
  const int PS = 256 / vp;
  // split rows into gx segments while there are fewer work items than thread slots on the machine
  int maxw = 1;
  for (int i = 0; i < ncls; ++i) maxw = std::max(maxw, cls[i].Wg);
  p.seg = maxw; p.nseg = 1;
  while (p.seg > 4 && (int64_t)maxrows * p.nseg * ncls < (int64_t)NSM * 8 * PS) { p.seg = (p.seg + 1) / 2; p.nseg = (maxw + p.seg - 1) / p.seg; }
  dim3 grid((unsigned)std::min((maxrows * p.nseg + PS - 1) / PS, NSM * 8), (unsigned)ncls);
#define THIN_IN_LAUNCH(CI, NT)                                                                                     \
  do {                                                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(thin_in_kernel<CI, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
    thin_in_kernel<CI, NT><<<grid, 256, smem, st.s>>>(p, in, out);                                                 \
  } while (0)
#define THIN_IN_NT(CI)                                    \
  do {                                                    \
    if (maxt <= 4) THIN_IN_LAUNCH(CI, 4);                 \
    else if (maxt <= 9) THIN_IN_LAUNCH(CI, 9);            \
    else THIN_IN_LAUNCH(CI, 16);                          \
  } while (0)
  switch (g.Ci) {
    case 1: THIN_IN_NT(1); break;
    case 2: THIN_IN_NT(2); break;
    case 3: THIN_IN_NT(3); break;
    default: THIN_IN_NT(4); break;
  }
#undef THIN_IN_NT
#undef THIN_IN_LAUNCH
  DSR_LAUNCHED(st, "tapconv_thin_in", bytes, WORK_BYTES);
  return true;
}

// =====================================================================================================================
// thin-OUTPUT convolution (Co <= 4): D's final 512->1 conv, G's last conv 16->1 / 12->3 (train.lua:111,133), and the dgrad
// of every thin-input layer.  One pass over the input: WPP warps per output pixel split the K = ntaps*Ci contraction
// (float4 over ci, coalesced), warp shuffle + shared-memory reduction, activation fused.
// =====================================================================================================================
struct ThinOut {
  int N, Hi, Wi, Ci, Ho, Wo, Co, Hg, Wg, si, so, oy0, ox0, ntaps, act, wpp;
  float neg;
  short dy[DSR_MAX_TAPS], dx[DSR_MAX_TAPS];
};

template <int CO>
__global__ void __launch_bounds__(256) thin_out_kernel(const ThinOut p, const float* __restrict__ in, const float* __restrict__ wp,
                                                        float* __restrict__ out) {
  __shared__ float red[8][CO];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ppb = 8 / p.wpp;                                    // pixels per block
  const int sub = warp % p.wpp;                                 // this warp's slice of K
  const int64_t npix = (int64_t)p.N * p.Hg * p.Wg;
  const int V = p.Ci >> 2;                                      // float4 per tap
  const int KV = p.ntaps * V;
  for (int64_t base = (int64_t)blockIdx.x * ppb; base < npix; base += (int64_t)gridDim.x * ppb) {
    const int64_t pix = base + warp / p.wpp;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    int gx = 0, gy = 0, n = 0;
    const bool pv = pix < npix;
    if (pv) {
      int64_t q = pix;
      gx = (int)(q % p.Wg); q /= p.Wg;
      gy = (int)(q % p.Hg);
      n = (int)(q / p.Hg);
      for (int kv = sub * 32 + lane; kv < KV; kv += 32 * p.wpp) {
        const int t = kv / V, c4 = kv - t * V;
        const int iy = gy * p.si + p.dy[t], ix = gx * p.si + p.dx[t];
        if (iy < 0 || iy >= p.Hi || ix < 0 || ix >= p.Wi) continue;
        const float4 x = __ldg(reinterpret_cast<const float4*>(in + ((int64_t)(n * p.Hi + iy) * p.Wi + ix) * p.Ci + c4 * 4));
        const float* w = wp + (int64_t)(t * p.Ci + c4 * 4) * CO;      // [k][co]
#pragma unroll
        for (int c = 0; c < CO; ++c)
          acc[c] = fmaf(x.x, __ldg(w + c), fmaf(x.y, __ldg(w + CO + c), fmaf(x.z, __ldg(w + 2 * CO + c), fmaf(x.w, __ldg(w + 3 * CO + c), acc[c]))));
      }
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < CO; ++c) red[warp][c] = acc[c];
    }
    __syncthreads();
    if (pv && sub == 0 && lane < CO) {
      float s = 0.f;
      for (int j = 0; j < p.wpp; ++j) s += red[warp + j][lane];
      out[((int64_t)(n * p.Ho + gy * p.so + p.oy0) * p.Wo + gx * p.so + p.ox0) * p.Co + lane] = thin_act(s, p.act, p.neg);
    }
    __syncthreads();
  }
}

// Many pixels, short contraction (G's last conv 12 -> 3 of train.lua:111 when the tensor-core kernels cannot take it: 12 is not a
// multiple of 8): one thread per output pixel, the whole K = ntaps * Ci contraction in registers, weights in shared memory as
// [tap][4-channel chunk][co] float4 so that a chunk costs one 16-byte input load + CO broadcast shared loads + 4*CO FMAs.
template <int CO>
__global__ void __launch_bounds__(256) thin_out_px_kernel(const ThinOut p, const float* __restrict__ in, const float* __restrict__ wp,
                                                           float* __restrict__ out) {
  extern __shared__ __align__(16) float4 sw4[];                 // [t][c4][co] = w[t][4*c4 .. 4*c4+3][co]
  const int V = p.Ci >> 2;
  for (int i = threadIdx.x; i < p.ntaps * V * CO; i += 256) {
    const int co = i % CO, tc = i / CO;                         // tc = t*V + c4 -> first input channel row (t*Ci + 4*c4)
    const float* w = wp + (int64_t)tc * 4 * CO + co;
    sw4[i] = make_float4(w[0], w[CO], w[2 * CO], w[3 * CO]);
  }
  __syncthreads();
  const int64_t npix = (int64_t)p.N * p.Hg * p.Wg;
  for (int64_t pix = (int64_t)blockIdx.x * 256 + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * 256) {
    int64_t q = pix;
    const int gx = (int)(q % p.Wg); q /= p.Wg;
    const int gy = (int)(q % p.Hg);
    const int n = (int)(q / p.Hg);
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    for (int t = 0; t < p.ntaps; ++t) {
      const int iy = gy * p.si + p.dy[t], ix = gx * p.si + p.dx[t];
      if (iy < 0 || iy >= p.Hi || ix < 0 || ix >= p.Wi) continue;
      const float4* ip = reinterpret_cast<const float4*>(in + ((int64_t)(n * p.Hi + iy) * p.Wi + ix) * p.Ci);
      const float4* w = sw4 + t * V * CO;
      for (int c4 = 0; c4 < V; ++c4) {
        const float4 x = __ldg(ip + c4);
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          const float4 wv = w[c4 * CO + c];
          acc[c] = fmaf(x.x, wv.x, fmaf(x.y, wv.y, fmaf(x.z, wv.z, fmaf(x.w, wv.w, acc[c]))));
        }
      }
    }
    float* o = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0) * p.Wo + gx * p.so + p.ox0) * p.Co;
#pragma unroll
    for (int c = 0; c < CO; ++c) o[c] = thin_act(acc[c], p.act, p.neg);
  }
}
bool thin_out_px_supported(const TapGeom& g) {
  return g.Co >= 1 && g.Co <= 4 && g.Ci % 4 == 0 && g.Ci >= 4 && g.Ci <= 64 && g.ntaps >= 1 && g.ntaps <= DSR_MAX_TAPS &&
         (int64_t)g.N * g.Hg * g.Wg > NSM * 32;
}
bool k_tapconv_thin_out_px(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float neg) {
  if (!thin_out_px_supported(g)) return false;
  ThinOut p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hi = g.Hi; p.Wi = g.Wi; p.Ci = g.Ci; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co; p.Hg = g.Hg; p.Wg = g.Wg;
  p.si = g.si; p.so = g.so; p.oy0 = g.oy0; p.ox0 = g.ox0; p.ntaps = g.ntaps; p.act = act; p.neg = neg;
  for (int t = 0; t < g.ntaps; ++t) { p.dy[t] = (short)g.dy[t]; p.dx[t] = (short)g.dx[t]; }
  const int64_t npix = (int64_t)g.N * g.Hg * g.Wg;
  const size_t smem = (size_t)g.ntaps * (g.Ci / 4) * g.Co * sizeof(float4);
  const unsigned nb = (unsigned)std::min<int64_t>((npix + 255) / 256, NSM * 16);
  switch (g.Co) {
    case 1: thin_out_px_kernel<1><<<nb, 256, smem, st.s>>>(p, in, wp, out); break;
    case 2: thin_out_px_kernel<2><<<nb, 256, smem, st.s>>>(p, in, wp, out); break;
    case 3: thin_out_px_kernel<3><<<nb, 256, smem, st.s>>>(p, in, wp, out); break;
    default: thin_out_px_kernel<4><<<nb, 256, smem, st.s>>>(p, in, wp, out); break;
  }
  DSR_LAUNCHED(st, "tapconv_thin_out_px", 4.0 * ((double)g.N * g.Hi * g.Wi * g.Ci + (double)npix * g.Co), WORK_BYTES);
  return true;
}

// only where the contraction is long and the pixels are few (D's final 512 -> 1 conv: 64 outputs of K = 8192); with many
// pixels the GEMM kernels' tiling of the input wins (a warp-per-pixel reduction re-reads every input pixel per tap)
bool thin_out_supported(const TapGeom& g) {
  return g.Co >= 1 && g.Co <= 4 && g.Ci % 4 == 0 && g.Ci >= 4 && g.ntaps >= 1 && (int64_t)g.N * g.Hg * g.Wg <= NSM * 32;
}

bool k_tapconv_thin_out(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float neg) {
  if (!thin_out_supported(g)) return false;
  ThinOut p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hi = g.Hi; p.Wi = g.Wi; p.Ci = g.Ci; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co; p.Hg = g.Hg; p.Wg = g.Wg;
  p.si = g.si; p.so = g.so; p.oy0 = g.oy0; p.ox0 = g.ox0; p.ntaps = g.ntaps; p.act = act; p.neg = neg;
  for (int t = 0; t < g.ntaps; ++t) { p.dy[t] = (short)g.dy[t]; p.dx[t] = (short)g.dx[t]; }
  const int64_t npix = (int64_t)g.N * g.Hg * g.Wg;
  if (npix <= 0) return true;
  const int KV = g.ntaps * (g.Ci / 4);
  // warps per pixel: enough lanes for K, and enough blocks for the machine when there are few pixels
  int wpp = 1;
  while (wpp < 8 && (KV > 64 * wpp) && (npix * wpp < (int64_t)NSM * 64)) wpp <<= 1;
  p.wpp = wpp;
  const int ppb = 8 / wpp;
  int64_t nb = std::min<int64_t>((npix + ppb - 1) / ppb, NSM * 16);
  switch (g.Co) {
    case 1: thin_out_kernel<1><<<(unsigned)nb, 256, 0, st.s>>>(p, in, wp, out); break;
    case 2: thin_out_kernel<2><<<(unsigned)nb, 256, 0, st.s>>>(p, in, wp, out); break;
    case 3: thin_out_kernel<3><<<(unsigned)nb, 256, 0, st.s>>>(p, in, wp, out); break;
    default: thin_out_kernel<4><<<(unsigned)nb, 256, 0, st.s>>>(p, in, wp, out); break;
  }
  DSR_LAUNCHED(st, "tapconv_thin_out", 4.0 * ((double)g.N * g.Hi * g.Wi * g.Ci + (double)npix * g.Co), WORK_BYTES);
  return true;
}
