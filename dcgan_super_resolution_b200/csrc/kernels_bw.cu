// kernels_bw.cu -- the HBM-bound kernels of the DCGAN-SR training step (sm_100a).
//
// Everything here is bandwidth work: layout import/export, BatchNorm statistics / apply /
// backward, activations, nearest up-sampling, the 2x2 box down-sample, the criteria, the
// per-sample pixel MSE and the fused Adam update.  Activations are NHWC fp32; the channel is
// the fastest dimension so every kernel reads and writes 16-byte vectors when C % 4 == 0.
// Reductions are deterministic: per-thread double accumulators -> shared-memory tree ->
// one partial row per CTA -> fixed-order column sum (no atomics).
//
// Reference semantics restated:
//   nn.SpatialBatchNormalization  train.lua:100        nn.ReLU/LeakyReLU/Tanh/Sigmoid  train.lua:100-134
//   nn.SpatialUpSamplingNearest   train-gray.lua:104   2x2 box down-sample loop        train.lua:225-230
//   nn.BCECriterion / MSECriterion train-gray-patch.lua:113 / train.lua:142
//   calMSE loop                   train.lua:193-195,237-239     optim.adam   train.lua:280
#include "common.h"
#include <math.h>
#include <algorithm>

#define NSM 148

static inline int cdiv64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
// layout import / export: per image a [R][Cc] -> [Cc][R] transpose through shared memory
// ------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc, int N) {
  __shared__ float tile[32][33];
  int tiles_r = (R + 31) / 32;
  for (int n = blockIdx.z; n < N; n += gridDim.z) {
    const float* src = in + (int64_t)n * R * Cc;
    float* dst = out + (int64_t)n * R * Cc;
    for (int tr = blockIdx.y; tr < tiles_r; tr += gridDim.y) {
      int c0 = blockIdx.x * 32, r0 = tr * 32;
      for (int j = threadIdx.y; j < 32; j += 8) {
        int r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[j][threadIdx.x] = src[(int64_t)r * Cc + c];
      }
      __syncthreads();
      for (int j = threadIdx.y; j < 32; j += 8) {
        int c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < Cc) dst[(int64_t)c * R + r] = tile[threadIdx.x][j];
      }
      __syncthreads();
    }
  }
}

static void transpose_batched(St st, const float* in, float* out, int N, int R, int Cc) {
  if (R == 1 || Cc == 1) {
    cudaMemcpyAsync(out, in, (size_t)N * R * Cc * sizeof(float), cudaMemcpyDeviceToDevice, st.s);
    return;
  }
  int gy = (R + 31) / 32;
  if (gy > 4096) gy = 4096;
  dim3 grid((Cc + 31) / 32, gy, N > 1024 ? 1024 : N);
  transpose_kernel<<<grid, dim3(32, 8), 0, st.s>>>(in, out, R, Cc, N);
  DSR_LAUNCHED(st, "transpose", 8.0 * N * R * Cc, WORK_BYTES);
}

// per image: [C][HW] -> [HW][C]
void k_nchw_to_nhwc(St st, const float* in, float* out, int N, int C, int H, int W) {
  transpose_batched(st, in, out, N, C, H * W);
}
// per image: [HW][C] -> [C][HW]
void k_nhwc_to_nchw(St st, const float* in, float* out, int N, int C, int H, int W) {
  transpose_batched(st, in, out, N, H * W, C);
}

// ------------------------------------------------------------------------------------------
// activations
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act, float neg) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LRELU: return v > 0.f ? v : v * neg;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
// gradient through the activation given its OUTPUT y (in-place Torch7 modules test the output)
__device__ __forceinline__ float act_grad(float y, float dy, int act, float neg) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? dy : 0.f;
    case ACT_LRELU: return y > 0.f ? dy : dy * neg;
    case ACT_TANH: return dy * (1.f - y * y);
    case ACT_SIGMOID: return dy * y * (1.f - y);
    default: return dy;
  }
}

__global__ void act_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t count, int act, float neg, int vec) {
  int64_t n4 = vec ? count >> 2 : 0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (int64_t k = i; k < n4; k += stride) {
    float4 v = x4[k];
    v.x = act_apply(v.x, act, neg); v.y = act_apply(v.y, act, neg);
    v.z = act_apply(v.z, act, neg); v.w = act_apply(v.w, act, neg);
    y4[k] = v;
  }
  for (int64_t k = (n4 << 2) + i; k < count; k += stride) y[k] = act_apply(x[k], act, neg);
}

__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx,
                               int64_t count, int act, float neg, int vec) {
  int64_t n4 = vec ? count >> 2 : 0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const float4* d4 = reinterpret_cast<const float4*>(dy);
  float4* o4 = reinterpret_cast<float4*>(dx);
  for (int64_t k = i; k < n4; k += stride) {
    float4 a = y4[k], d = d4[k], o;
    o.x = act_grad(a.x, d.x, act, neg); o.y = act_grad(a.y, d.y, act, neg);
    o.z = act_grad(a.z, d.z, act, neg); o.w = act_grad(a.w, d.w, act, neg);
    o4[k] = o;
  }
  for (int64_t k = (n4 << 2) + i; k < count; k += stride) dx[k] = act_grad(y[k], dy[k], act, neg);
}

static inline int ew_grid(int64_t count) {
  int64_t want = (count / 4 + 255) / 256;
  int64_t cap = NSM * 16;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

void k_act(St st, const float* x, float* y, int64_t count, int act, float negval) {
  if (count <= 0) return;
  // float4 path only for 16-byte aligned tensors (a sample-group offset into a 1-channel tensor is not)
  const int vec = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  act_kernel<<<ew_grid(count), 256, 0, st.s>>>(x, y, count, act, negval, vec);
  DSR_LAUNCHED(st, "act", 8.0 * count, WORK_BYTES);
}
void k_act_bwd(St st, const float* y, const float* dy, float* dx, int64_t count, int act, float negval) {
  if (count <= 0) return;
  const int vec = (((uintptr_t)y | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0;
  act_bwd_kernel<<<ew_grid(count), 256, 0, st.s>>>(y, dy, dx, count, act, negval, vec);
  DSR_LAUNCHED(st, "act_bwd", 12.0 * count, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// BatchNorm.  x is [P][C] (P = N*H*W pixels).  Thread layout inside a 256-thread CTA:
// tx in [0,TX) walks channel vectors, ty in [0,TY) walks pixels; a CTA owns a contiguous
// pixel range, so each warp reads whole contiguous runs of the tensor.
// ------------------------------------------------------------------------------------------
struct BnCfg { int vec, CV, TX, TY, nb; int64_t ppb; };

static BnCfg bn_cfg(int64_t P, int C) {
  BnCfg c;
  c.vec = (C % 4 == 0) ? 4 : 1;
  c.CV = C / c.vec;
  int tx = 1;
  while (tx < c.CV && tx < 256) tx <<= 1;
  c.TX = tx;
  c.TY = 256 / tx;
  int64_t min_ppb = (int64_t)c.TY * 8;                 // at least 8 pixels per thread row
  int64_t nb = (P + min_ppb - 1) / min_ppb;
  int64_t cap = NSM * 4;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  c.ppb = (P + nb - 1) / nb;
  c.nb = (int)((P + c.ppb - 1) / c.ppb);
  if (c.nb < 1) c.nb = 1;
  return c;
}

int bn_partial_rows(int64_t P, int C) { return bn_cfg(P, C).nb; }

// Reduce per-thread accumulators a[2*VEC] over ty and write one partial row.
template <int VEC>
__device__ __forceinline__ void bn_block_reduce_store(double (&a)[2 * VEC], int tx, int ty, int TX, int TY, int cv, int CV,
                                                      int C, double* __restrict__ prow) {
  __shared__ double red[256 * 2 * VEC];
  int tid = ty * TX + tx;
#pragma unroll
  for (int j = 0; j < 2 * VEC; ++j) red[j * 256 + tid] = a[j];
  __syncthreads();
  for (int off = TY >> 1; off > 0; off >>= 1) {
    if (ty < off) {
#pragma unroll
      for (int j = 0; j < 2 * VEC; ++j) red[j * 256 + tid] += red[j * 256 + tid + off * TX];
    }
    __syncthreads();
  }
  if (ty == 0 && cv < CV) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      prow[cv * VEC + j] = red[j * 256 + tid];
      prow[C + cv * VEC + j] = red[(VEC + j) * 256 + tid];
    }
  }
  __syncthreads();
}

#define BN_CHUNK 16
template <int VEC>
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int64_t P, int C, int TX, int TY,
                                                       int64_t ppb, double* __restrict__ partials) {
  int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int CV = C / VEC;
  x += (int64_t)blockIdx.y * P * C;                               // sample group
  int64_t p0 = (int64_t)blockIdx.x * ppb;
  int64_t p1 = p0 + ppb < P ? p0 + ppb : P;
  double* prow = partials + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C;
  for (int cv0 = 0; cv0 < CV; cv0 += TX) {
    int cv = cv0 + tx;
    double a[2 * VEC];
#pragma unroll
    for (int j = 0; j < 2 * VEC; ++j) a[j] = 0.0;
    if (cv < CV) {
      // two-level accumulation: BN_CHUNK pixels in fp32 (the f32 -> f64 conversions and DADDs of a per-element double sum
      // bound the kernel at ~4.9 TB/s), chunk sums in double -- the error of a 16-term fp32 sum is a few 1e-8 relative
      for (int64_t pc = p0 + ty; pc < p1; pc += (int64_t)TY * BN_CHUNK) {
        float f[2 * VEC];
#pragma unroll
        for (int j = 0; j < 2 * VEC; ++j) f[j] = 0.f;
        const int cnt = (int)((p1 - pc + TY - 1) / TY < BN_CHUNK ? (p1 - pc + TY - 1) / TY : BN_CHUNK);
        auto one = [&](const float4& v) {
          f[0] += v.x; f[1] += v.y; f[2] += v.z; f[3] += v.w;
          f[VEC + 0] = fmaf(v.x, v.x, f[VEC + 0]); f[VEC + 1] = fmaf(v.y, v.y, f[VEC + 1]);
          f[VEC + 2] = fmaf(v.z, v.z, f[VEC + 2]); f[VEC + 3] = fmaf(v.w, v.w, f[VEC + 3]);
        };
        if (VEC == 4 && cnt == BN_CHUNK) {
          // full chunk: loads batched four deep, no bounds checks
          const float* xp = x + pc * C + cv * 4;
          const int64_t step = (int64_t)TY * C;
#pragma unroll
          for (int u0 = 0; u0 < BN_CHUNK; u0 += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(xp + (int64_t)(u0 + u) * step);
#pragma unroll
            for (int u = 0; u < 4; ++u) one(v[u]);
          }
        } else {
          for (int u = 0; u < cnt; ++u) {
            const int64_t p = pc + (int64_t)u * TY;
            if (VEC == 4) one(*reinterpret_cast<const float4*>(x + p * C + cv * 4));
            else { const float d0 = x[p * C + cv]; f[0] += d0; f[VEC] = fmaf(d0, d0, f[VEC]); }
          }
        }
#pragma unroll
        for (int j = 0; j < 2 * VEC; ++j) a[j] += (double)f[j];
      }
    }
    bn_block_reduce_store<VEC>(a, tx, ty, TX, TY, cv, CV, C, prow);
  }
}

// sums[j] = sum_b partials[b][j]   (fixed order: deterministic)
__global__ void colsum_kernel(const double* __restrict__ partials, int nb, int ncol, double* __restrict__ sums) {
  __shared__ double red[8][33];
  int j = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0;
  if (j < ncol)
    for (int b = threadIdx.y; b < nb; b += 8) a += partials[(int64_t)b * ncol + j];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < ncol) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
    sums[j] = s;
  }
}

void k_bn_stats(St st, const float* x, int64_t P, int C, double* partials, double* sums) {
  BnCfg c = bn_cfg(P, C);
  if (c.vec == 4) bn_stats_kernel<4><<<c.nb, 256, 0, st.s>>>(x, P, C, c.TX, c.TY, c.ppb, partials);
  else bn_stats_kernel<1><<<c.nb, 256, 0, st.s>>>(x, P, C, c.TX, c.TY, c.ppb, partials);
  DSR_LAUNCHED(st, "bn_stats", 4.0 * P * C, WORK_BYTES);
  colsum_kernel<<<(2 * C + 31) / 32, dim3(32, 8), 0, st.s>>>(partials, c.nb, 2 * C, sums);
  DSR_LAUNCHED(st, "bn_colsum", 16.0 * c.nb * C, WORK_BYTES);
}

// mean / invstd / running statistics (biased variance normalises, unbiased goes to running_var)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double n_total, float eps, float momentum,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = sums[c] / n_total;
  double var = sums[C + c] / n_total - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = 1.0 / sqrt(var + (double)eps);
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)invstd;
  if (running_mean) {
    double unb = n_total > 1.0 ? var * (n_total / (n_total - 1.0)) : var;
    running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
    running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unb);
  }
}

void k_bn_finalize(St st, const double* sums, int C, double n_total, float eps, float momentum,
                   float* save_mean, float* save_invstd, float* running_mean, float* running_var) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st.s>>>(sums, C, n_total, eps, momentum, save_mean, save_invstd,
                                                         running_mean, running_var);
  DSR_LAUNCHED(st, "bn_finalize", 32.0 * C, WORK_BYTES);
}

// BatchNorm pre-activation gamma * xhat + beta with every rounding spelled out: the forward (bn_apply_kernel) and the backward
// (which re-derives the ReLU / LeakyReLU mask from x instead of reading the output tensor) must agree bit for bit, whatever the
// compiler contracts elsewhere.
__device__ __forceinline__ float bn_pre(float x, float m, float s, float g, float b) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(x, m), s), g, b);
}
// y = act(gamma * (x - mean) * invstd + beta)
template <int VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total_v, int CV,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       int act, float neg, int64_t sstride) {
  x += (int64_t)blockIdx.y * total_v * VEC; y += (int64_t)blockIdx.y * total_v * VEC;      // sample group
  mean += (int64_t)blockIdx.y * sstride; invstd += (int64_t)blockIdx.y * sstride;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_v; i += stride) {
    int cv = (int)(i % CV);
    if (VEC == 4) {
      float4 v = reinterpret_cast<const float4*>(x)[i];
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + cv);
      float4 b = __ldg(reinterpret_cast<const float4*>(beta) + cv);
      float4 m = __ldg(reinterpret_cast<const float4*>(mean) + cv);
      float4 s = __ldg(reinterpret_cast<const float4*>(invstd) + cv);
      v.x = act_apply(bn_pre(v.x, m.x, s.x, g.x, b.x), act, neg);
      v.y = act_apply(bn_pre(v.y, m.y, s.y, g.y, b.y), act, neg);
      v.z = act_apply(bn_pre(v.z, m.z, s.z, g.z, b.z), act, neg);
      v.w = act_apply(bn_pre(v.w, m.w, s.w, g.w, b.w), act, neg);
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      float v = x[i];
      y[i] = act_apply(bn_pre(v, __ldg(mean + cv), __ldg(invstd + cv), __ldg(gamma + cv), __ldg(beta + cv)), act, neg);
    }
  }
}

void k_bn_apply_act(St st, const float* x, float* y, int64_t P, int C, const float* gamma, const float* beta,
                    const float* mean, const float* invstd, int act, float negval) {
  if (C % 4 == 0) {
    int64_t tv = P * (C / 4);
    bn_apply_kernel<4><<<ew_grid(tv * 4), 256, 0, st.s>>>(x, y, tv, C / 4, gamma, beta, mean, invstd, act, negval, 0);
  } else {
    int64_t tv = P * C;
    bn_apply_kernel<1><<<ew_grid(tv * 4), 256, 0, st.s>>>(x, y, tv, C, gamma, beta, mean, invstd, act, negval, 0);
  }
  DSR_LAUNCHED(st, "bn_apply_act", 8.0 * P * C, WORK_BYTES);
}

// g = dy * act'(.) of one element.  y != nullptr: the activation OUTPUT is read (Torch7 in-place modules test the output; needed
// when the parameters changed since the cached forward -- fGx's updateGradInput through D, train.lua:268 -- and for Tanh /
// Sigmoid).  y == nullptr (ReLU / LeakyReLU only): output > 0 <=> pre-activation > 0, recomputed from x: one tensor less to read.
__device__ __forceinline__ float bn_gval(float dy, const float* y, int64_t o, float x, float m, float s, float ga, float be, int act, float neg) {
  if (act == ACT_NONE) return dy;
  if (y) return act_grad(y[o], dy, act, neg);
  const bool pos = bn_pre(x, m, s, ga, be) > 0.f;
  return pos ? dy : (act == ACT_LRELU ? dy * neg : 0.f);
}

// backward reductions: g = dy * act'(y) ; sum g, sum g * xhat.  Reads dy, x (and y when given); writes only the partial sums.
template <int VEC>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, int64_t P,
                                                            int C, int TX, int TY, int64_t ppb, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, int act, float neg,
                                                            double* __restrict__ partials, int64_t sstride) {
  int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  int CV = C / VEC;
  {
    const int64_t go = (int64_t)blockIdx.y * P * C;               // sample group
    dy += go; x += go;
    if (y) y += go;
    mean += (int64_t)blockIdx.y * sstride; invstd += (int64_t)blockIdx.y * sstride;
  }
  int64_t p0 = (int64_t)blockIdx.x * ppb;
  int64_t p1 = p0 + ppb < P ? p0 + ppb : P;
  double* prow = partials + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C;
  for (int cv0 = 0; cv0 < CV; cv0 += TX) {
    int cv = cv0 + tx;
    double a[2 * VEC];
#pragma unroll
    for (int j = 0; j < 2 * VEC; ++j) a[j] = 0.0;
    if (cv < CV) {
      if (VEC == 4) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + cv);
        const float4 s = __ldg(reinterpret_cast<const float4*>(invstd) + cv);
        float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act != ACT_NONE && !y) { ga = __ldg(reinterpret_cast<const float4*>(gamma) + cv); be = __ldg(reinterpret_cast<const float4*>(beta) + cv); }
        auto gof = [&](const float4& d, const float4& xv, int64_t o) {
          float4 g = d;
          if (act != ACT_NONE) {
            if (y) {
              const float4 yv = *reinterpret_cast<const float4*>(y + o);
              g.x = act_grad(yv.x, d.x, act, neg); g.y = act_grad(yv.y, d.y, act, neg);
              g.z = act_grad(yv.z, d.z, act, neg); g.w = act_grad(yv.w, d.w, act, neg);
            } else {
              g.x = bn_gval(d.x, nullptr, 0, xv.x, m.x, s.x, ga.x, be.x, act, neg);
              g.y = bn_gval(d.y, nullptr, 0, xv.y, m.y, s.y, ga.y, be.y, act, neg);
              g.z = bn_gval(d.z, nullptr, 0, xv.z, m.z, s.z, ga.z, be.z, act, neg);
              g.w = bn_gval(d.w, nullptr, 0, xv.w, m.w, s.w, ga.w, be.w, act, neg);
            }
          }
          return g;
        };
        for (int64_t pc = p0 + ty; pc < p1; pc += (int64_t)TY * BN_CHUNK) {
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // fp32 over a chunk of pixels, double across chunks (see bn_stats_kernel)
          auto one = [&](const float4& d, const float4& xv, int64_t o) {
            const float4 g = gof(d, xv, o);
            f[0] += g.x; f[1] += g.y; f[2] += g.z; f[3] += g.w;
            f[4] = fmaf(g.x, (xv.x - m.x) * s.x, f[4]); f[5] = fmaf(g.y, (xv.y - m.y) * s.y, f[5]);
            f[6] = fmaf(g.z, (xv.z - m.z) * s.z, f[6]); f[7] = fmaf(g.w, (xv.w - m.w) * s.w, f[7]);
          };
          const int cnt = (int)((p1 - pc + TY - 1) / TY < BN_CHUNK ? (p1 - pc + TY - 1) / TY : BN_CHUNK);
          const int64_t o0 = pc * C + cv * 4, step = (int64_t)TY * C;
          if (cnt == BN_CHUNK) {
#pragma unroll
            for (int u0 = 0; u0 < BN_CHUNK; u0 += 4) {
              float4 d[4], xv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                d[u] = *reinterpret_cast<const float4*>(dy + o0 + (int64_t)(u0 + u) * step);
                xv[u] = *reinterpret_cast<const float4*>(x + o0 + (int64_t)(u0 + u) * step);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) one(d[u], xv[u], o0 + (int64_t)(u0 + u) * step);
            }
          } else {
            for (int u = 0; u < cnt; ++u) {
              const int64_t o = o0 + (int64_t)u * step;
              one(*reinterpret_cast<const float4*>(dy + o), *reinterpret_cast<const float4*>(x + o), o);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] += (double)f[j];
        }
      } else {
        const float m = __ldg(mean + cv), s = __ldg(invstd + cv);
        const float ga = (act != ACT_NONE && !y) ? __ldg(gamma + cv) : 1.f, be = (act != ACT_NONE && !y) ? __ldg(beta + cv) : 0.f;
        for (int64_t pc = p0 + ty; pc < p1; pc += (int64_t)TY * BN_CHUNK) {
          float f0 = 0.f, f1 = 0.f;
#pragma unroll 4
          for (int u = 0; u < BN_CHUNK; ++u) {
            const int64_t p = pc + (int64_t)u * TY;
            if (p >= p1) break;
            const int64_t o = p * C + cv;
            const float xv = x[o];
            const float g = bn_gval(dy[o], y, o, xv, m, s, ga, be, act, neg);
            f0 += g;
            f1 = fmaf(g, (xv - m) * s, f1);
          }
          a[0] += (double)f0;
          a[VEC] += (double)f1;
        }
      }
    }
    bn_block_reduce_store<VEC>(a, tx, ty, TX, TY, cv, CV, C, prow);
  }
}

static inline double bn_bwd_reduce_bytes(int64_t P, int C, int act, const float* y) { return ((act != ACT_NONE && y) ? 12.0 : 8.0) * P * C; }
static inline double bn_bwd_apply_bytes(int64_t P, int C, int act, const float* y) { return ((act != ACT_NONE && y) ? 16.0 : 12.0) * P * C; }

void k_bn_bwd_reduce(St st, const float* dy, const float* y, const float* x, int64_t P, int C, const float* gamma, const float* beta,
                     const float* mean, const float* invstd, int act, float negval, double* partials, double* sums) {
  BnCfg c = bn_cfg(P, C);
  if (c.vec == 4)
    bn_bwd_reduce_kernel<4><<<c.nb, 256, 0, st.s>>>(dy, y, x, P, C, c.TX, c.TY, c.ppb, gamma, beta, mean, invstd, act, negval, partials, 0);
  else
    bn_bwd_reduce_kernel<1><<<c.nb, 256, 0, st.s>>>(dy, y, x, P, C, c.TX, c.TY, c.ppb, gamma, beta, mean, invstd, act, negval, partials, 0);
  DSR_LAUNCHED(st, "bn_bwd_reduce", bn_bwd_reduce_bytes(P, C, act, y), WORK_BYTES);
  colsum_kernel<<<(2 * C + 31) / 32, dim3(32, 8), 0, st.s>>>(partials, c.nb, 2 * C, sums);
  DSR_LAUNCHED(st, "bn_colsum", 16.0 * c.nb * C, WORK_BYTES);
}

__global__ void bn_bwd_param_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] += (float)sums[c];
  if (dgamma) dgamma[c] += (float)sums[C + c];
}
void k_bn_bwd_param(St st, const double* sums_local, int C, float* dgamma, float* dbeta) {
  bn_bwd_param_kernel<<<(C + 127) / 128, 128, 0, st.s>>>(sums_local, C, dgamma, dbeta);
  DSR_LAUNCHED(st, "bn_bwd_param", 32.0 * C, WORK_BYTES);
}

// dx = (g - sum_g/n - xhat * sum_gxhat/n) * gamma * invstd with g = dy * act'(.) re-derived per element (see bn_gval); dx may
// alias dy (in place: every element is read before it is written by the same thread).
template <int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* dy, const float* __restrict__ y, const float* __restrict__ x,
                                                           float* dx /* may alias dy */, int64_t total_v, int CV, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd, int act,
                                                           float neg, const float* __restrict__ fmeans, int64_t sstride) {
  // fmeans[2C] per group: (float)(sum g / n), (float)(sum g*xhat / n), written once by the tail kernel (the double divisions
  // used to sit in this loop: 8 per float4, which bound the kernel at ~4 TB/s)
  dy += (int64_t)blockIdx.y * total_v * VEC; x += (int64_t)blockIdx.y * total_v * VEC; dx += (int64_t)blockIdx.y * total_v * VEC;
  if (y) y += (int64_t)blockIdx.y * total_v * VEC;
  mean += (int64_t)blockIdx.y * sstride; invstd += (int64_t)blockIdx.y * sstride; fmeans += (int64_t)blockIdx.y * 2 * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_v; i += stride) {
    int cv = (int)(i % CV);
    if (VEC == 4) {
      float4 gv = reinterpret_cast<const float4*>(dy)[i];
      float4 xv = reinterpret_cast<const float4*>(x)[i];
      float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cv);
      float4 m = __ldg(reinterpret_cast<const float4*>(mean) + cv);
      float4 s = __ldg(reinterpret_cast<const float4*>(invstd) + cv);
      if (act != ACT_NONE) {
        if (y) {
          float4 yv = reinterpret_cast<const float4*>(y)[i];
          gv.x = act_grad(yv.x, gv.x, act, neg); gv.y = act_grad(yv.y, gv.y, act, neg);
          gv.z = act_grad(yv.z, gv.z, act, neg); gv.w = act_grad(yv.w, gv.w, act, neg);
        } else {
          float4 be = __ldg(reinterpret_cast<const float4*>(beta) + cv);
          gv.x = bn_gval(gv.x, nullptr, 0, xv.x, m.x, s.x, ga.x, be.x, act, neg);
          gv.y = bn_gval(gv.y, nullptr, 0, xv.y, m.y, s.y, ga.y, be.y, act, neg);
          gv.z = bn_gval(gv.z, nullptr, 0, xv.z, m.z, s.z, ga.z, be.z, act, neg);
          gv.w = bn_gval(gv.w, nullptr, 0, xv.w, m.w, s.w, ga.w, be.w, act, neg);
        }
      }
      const float4 mg = __ldg(reinterpret_cast<const float4*>(fmeans) + cv);
      const float4 mx = __ldg(reinterpret_cast<const float4*>(fmeans + C) + cv);
      float4 o;
      o.x = (gv.x - mg.x - (xv.x - m.x) * s.x * mx.x) * (ga.x * s.x);
      o.y = (gv.y - mg.y - (xv.y - m.y) * s.y * mx.y) * (ga.y * s.y);
      o.z = (gv.z - mg.z - (xv.z - m.z) * s.z * mx.z) * (ga.z * s.z);
      o.w = (gv.w - mg.w - (xv.w - m.w) * s.w * mx.w) * (ga.w * s.w);
      reinterpret_cast<float4*>(dx)[i] = o;
    } else {
      const float mg = __ldg(fmeans + cv), mx = __ldg(fmeans + C + cv);
      float s = __ldg(invstd + cv), m = __ldg(mean + cv), ga = __ldg(gamma + cv);
      const float xv = x[i];
      const float g = bn_gval(dy[i], y, i, xv, m, s, ga, (act != ACT_NONE && !y) ? __ldg(beta + cv) : 0.f, act, neg);
      dx[i] = (g - mg - (xv - m) * s * mx) * (ga * s);
    }
  }
}

__global__ void bn_fmeans_kernel(const double* __restrict__ sums, int n2c, double n_total, float* __restrict__ fmeans) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2c) fmeans[i] = (float)(sums[i] / n_total);
}
void k_bn_bwd_apply(St st, const float* dy, const float* y, const float* x, float* dx, int64_t P, int C, const float* gamma,
                    const float* beta, const float* mean, const float* invstd, int act, float negval, const double* sums_total,
                    double n_total, float* fmeans) {
  bn_fmeans_kernel<<<(2 * C + 127) / 128, 128, 0, st.s>>>(sums_total, 2 * C, n_total, fmeans);
  DSR_LAUNCHED(st, "bn_fmeans", 24.0 * C, WORK_BYTES);
  if (C % 4 == 0) {
    int64_t tv = P * (C / 4);
    bn_bwd_apply_kernel<4><<<ew_grid(tv * 4), 256, 0, st.s>>>(dy, y, x, dx, tv, C / 4, C, gamma, beta, mean, invstd, act, negval, fmeans, 0);
  } else {
    int64_t tv = P * C;
    bn_bwd_apply_kernel<1><<<ew_grid(tv * 4), 256, 0, st.s>>>(dy, y, x, dx, tv, C, C, gamma, beta, mean, invstd, act, negval, fmeans, 0);
  }
  DSR_LAUNCHED(st, "bn_bwd_apply", bn_bwd_apply_bytes(P, C, act, y), WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// Grouped, fused BatchNorm (the fused training step; no cross-rank statistics): `groups` independent minibatches of P
// pixels each sit back to back in x.  Forward = 3 launches for all groups (partials, column sums + finalize, apply),
// backward = 3 (reduce, column sums + parameter gradients, apply) instead of 4 per group.  All sums in fixed order.
// ------------------------------------------------------------------------------------------
#define BN_TAIL_Y 32      // row threads per 32-channel block: the tails are latency-bound sums over <= 592 partial rows
__global__ void bn_fwd_tail_kernel(const double* __restrict__ partials, int nb, int C, int groups, double n_total, float eps,
                                   float momentum, float* __restrict__ save_mean, float* __restrict__ save_invstd, int64_t sstride,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  __shared__ double red[2][BN_TAIL_Y][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int g = 0; g < groups; ++g) {                    // sequential: the running statistics see group 0 first
    const double* pg = partials + (int64_t)g * nb * 2 * C;
    double a1 = 0.0, a2 = 0.0;
    if (c < C)
      for (int b = threadIdx.y; b < nb; b += BN_TAIL_Y) { a1 += pg[(int64_t)b * 2 * C + c]; a2 += pg[(int64_t)b * 2 * C + C + c]; }
    red[0][threadIdx.y][threadIdx.x] = a1;
    red[1][threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int r = 0; r < BN_TAIL_Y; ++r) { s1 += red[0][r][threadIdx.x]; s2 += red[1][r][threadIdx.x]; }
      double mean = s1 / n_total;
      double var = s2 / n_total - mean * mean;
      if (var < 0.0) var = 0.0;
      double invstd = 1.0 / sqrt(var + (double)eps);
      save_mean[(int64_t)g * sstride + c] = (float)mean;
      save_invstd[(int64_t)g * sstride + c] = (float)invstd;
      if (running_mean) {
        double unb = n_total > 1.0 ? var * (n_total / (n_total - 1.0)) : var;
        running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unb);
      }
    }
    __syncthreads();
  }
}

void k_bn_fwd_grouped(St st, const float* x, float* y, int64_t P, int C, int groups, const float* gamma, const float* beta,
                      float* save_mean, float* save_invstd, int64_t sstride, float* running_mean, float* running_var, float eps,
                      float momentum, int act, float negval, double* partials) {
  BnCfg c = bn_cfg(P, C);
  dim3 grid(c.nb, groups);
  if (c.vec == 4) bn_stats_kernel<4><<<grid, 256, 0, st.s>>>(x, P, C, c.TX, c.TY, c.ppb, partials);
  else bn_stats_kernel<1><<<grid, 256, 0, st.s>>>(x, P, C, c.TX, c.TY, c.ppb, partials);
  DSR_LAUNCHED(st, "bn_stats", 4.0 * P * C * groups, WORK_BYTES);
  bn_fwd_tail_kernel<<<(C + 31) / 32, dim3(32, BN_TAIL_Y), 0, st.s>>>(partials, c.nb, C, groups, (double)P, eps, momentum, save_mean, save_invstd,
                                                               sstride, running_mean, running_var);
  DSR_LAUNCHED(st, "bn_fwd_tail", 16.0 * c.nb * C * groups, WORK_BYTES);
  if (C % 4 == 0) {
    int64_t tv = P * (C / 4);
    bn_apply_kernel<4><<<dim3(ew_grid(tv * 4), groups), 256, 0, st.s>>>(x, y, tv, C / 4, gamma, beta, save_mean, save_invstd, act, negval, sstride);
  } else {
    int64_t tv = P * C;
    bn_apply_kernel<1><<<dim3(ew_grid(tv * 4), groups), 256, 0, st.s>>>(x, y, tv, C, gamma, beta, save_mean, save_invstd, act, negval, sstride);
  }
  DSR_LAUNCHED(st, "bn_apply_act", 8.0 * P * C * groups, WORK_BYTES);
}

void k_bn_fwd_from_partials(St st, const float* x, float* y, int64_t P, int C, int nb, const float* gamma, const float* beta, float* save_mean,
                            float* save_invstd, float* running_mean, float* running_var, float eps, float momentum, int act, float negval,
                            const double* partials) {
  bn_fwd_tail_kernel<<<(C + 31) / 32, dim3(32, BN_TAIL_Y), 0, st.s>>>(partials, nb, C, 1, (double)P, eps, momentum, save_mean, save_invstd, 0,
                                                               running_mean, running_var);
  DSR_LAUNCHED(st, "bn_fwd_tail", 16.0 * nb * C, WORK_BYTES);
  k_bn_apply_act(st, x, y, P, C, gamma, beta, save_mean, save_invstd, act, negval);
}

// sums[g][2C] = column sums of group g's partial rows; dbeta += sum_g sums[g][c], dgamma += sum_g sums[g][C + c]
__global__ void bn_bwd_tail_kernel(const double* __restrict__ partials, int nb, int C, int groups, double* __restrict__ sums,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ fmeans, double n_total) {
  __shared__ double red[2][BN_TAIL_Y][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int g = 0; g < groups; ++g) {
    const double* pg = partials + (int64_t)g * nb * 2 * C;
    double a1 = 0.0, a2 = 0.0;
    if (c < C)
      for (int b = threadIdx.y; b < nb; b += BN_TAIL_Y) { a1 += pg[(int64_t)b * 2 * C + c]; a2 += pg[(int64_t)b * 2 * C + C + c]; }
    red[0][threadIdx.y][threadIdx.x] = a1;
    red[1][threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int r = 0; r < BN_TAIL_Y; ++r) { s1 += red[0][r][threadIdx.x]; s2 += red[1][r][threadIdx.x]; }
      sums[(int64_t)g * 2 * C + c] = s1;
      sums[(int64_t)g * 2 * C + C + c] = s2;
      fmeans[(int64_t)g * 2 * C + c] = (float)(s1 / n_total);            // what the apply pass subtracts
      fmeans[(int64_t)g * 2 * C + C + c] = (float)(s2 / n_total);
      // the reference accumulates group after group in fp32 (two backward calls): same order here
      if (dbeta) dbeta[c] += (float)s1;
      if (dgamma) dgamma[c] += (float)s2;
    }
    __syncthreads();
  }
}

void k_bn_bwd_grouped(St st, const float* dy, const float* y, const float* x, float* dx, int64_t P, int C, int groups,
                      const float* gamma, const float* beta, const float* save_mean, const float* save_invstd, int64_t sstride, int act,
                      float negval, double* partials, double* sums, float* dgamma, float* dbeta, float* fmeans) {
  BnCfg c = bn_cfg(P, C);
  dim3 grid(c.nb, groups);
  if (c.vec == 4)
    bn_bwd_reduce_kernel<4><<<grid, 256, 0, st.s>>>(dy, y, x, P, C, c.TX, c.TY, c.ppb, gamma, beta, save_mean, save_invstd, act, negval, partials, sstride);
  else
    bn_bwd_reduce_kernel<1><<<grid, 256, 0, st.s>>>(dy, y, x, P, C, c.TX, c.TY, c.ppb, gamma, beta, save_mean, save_invstd, act, negval, partials, sstride);
  DSR_LAUNCHED(st, "bn_bwd_reduce", bn_bwd_reduce_bytes(P, C, act, y) * groups, WORK_BYTES);
  bn_bwd_tail_kernel<<<(C + 31) / 32, dim3(32, BN_TAIL_Y), 0, st.s>>>(partials, c.nb, C, groups, sums, dgamma, dbeta, fmeans, (double)P);
  DSR_LAUNCHED(st, "bn_bwd_tail", 16.0 * c.nb * C * groups, WORK_BYTES);
  if (C % 4 == 0) {
    int64_t tv = P * (C / 4);
    bn_bwd_apply_kernel<4><<<dim3(ew_grid(tv * 4), groups), 256, 0, st.s>>>(dy, y, x, dx, tv, C / 4, C, gamma, beta, save_mean, save_invstd, act, negval, fmeans, sstride);
  } else {
    int64_t tv = P * C;
    bn_bwd_apply_kernel<1><<<dim3(ew_grid(tv * 4), groups), 256, 0, st.s>>>(dy, y, x, dx, tv, C, C, gamma, beta, save_mean, save_invstd, act, negval, fmeans, sstride);
  }
  DSR_LAUNCHED(st, "bn_bwd_apply", bn_bwd_apply_bytes(P, C, act, y) * groups, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// nearest up-sampling, 2x2 box down-sample (NHWC, one thread per output element)
// ------------------------------------------------------------------------------------------
__global__ void upnearest_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int s) {
  int Ho = H * s, Wo = W * s;
  int64_t total = (int64_t)N * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int ox = (int)(p % Wo); p /= Wo;
    int oy = (int)(p % Ho);
    int n = (int)(p / Ho);
    y[i] = x[(((int64_t)n * H + oy / s) * W + ox / s) * C + c];
  }
}
__global__ void upnearest_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int s) {
  int Ho = H * s, Wo = W * s;
  int64_t total = (int64_t)N * H * W * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int ix = (int)(p % W); p /= W;
    int iy = (int)(p % H);
    int n = (int)(p / H);
    float a = 0.f;
    for (int u = 0; u < s; ++u)
      for (int v = 0; v < s; ++v) a += dy[(((int64_t)n * Ho + iy * s + u) * Wo + ix * s + v) * C + c];
    dx[i] = a;
  }
}
// (a[2i,2j] + a[2i+1,2j] + a[2i,2j+1] + a[2i+1,2j+1]) / 4 -- the association order of train.lua:227-229
__global__ void avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2;
  int64_t total = (int64_t)N * Ho * Wo * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int ox = (int)(p % Wo); p /= Wo;
    int oy = (int)(p % Ho);
    int n = (int)(p / Ho);
    const float* b = x + (((int64_t)n * H + 2 * oy) * W + 2 * ox) * C + c;
    float a00 = b[0], a10 = b[(int64_t)W * C], a01 = b[C], a11 = b[(int64_t)W * C + C];
    y[i] = (((a00 + a10) + a01) + a11) / 4.f;
  }
}
static inline int sc_grid(int64_t total) {
  int64_t want = (total + 255) / 256;
  int64_t cap = NSM * 16;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}
void k_upnearest_fwd(St st, const float* x, float* y, int N, int H, int W, int C, int scale) {
  upnearest_fwd_kernel<<<sc_grid((int64_t)N * H * W * C * scale * scale), 256, 0, st.s>>>(x, y, N, H, W, C, scale);
  DSR_LAUNCHED(st, "upnearest_fwd", 4.0 * N * H * W * C * (1 + scale * scale), WORK_BYTES);
}
void k_upnearest_bwd(St st, const float* dy, float* dx, int N, int H, int W, int C, int scale) {
  upnearest_bwd_kernel<<<sc_grid((int64_t)N * H * W * C), 256, 0, st.s>>>(dy, dx, N, H, W, C, scale);
  DSR_LAUNCHED(st, "upnearest_bwd", 4.0 * N * H * W * C * (1 + scale * scale), WORK_BYTES);
}
void k_avgpool2(St st, const float* x, float* y, int N, int H, int W, int C) {
  avgpool2_kernel<<<sc_grid((int64_t)N * (H / 2) * (W / 2) * C), 256, 0, st.s>>>(x, y, N, H, W, C);
  DSR_LAUNCHED(st, "avgpool2", 5.0 * N * H * W * C, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// criteria (count = nElement of D's output: small) -- one CTA, deterministic, evaluated in double
// like THNN's BCECriterion (eps 1e-12) / MSECriterion with sizeAverage = true.
// ------------------------------------------------------------------------------------------
// Small outputs (the DCGAN-64 discriminator: B values) take one CTA; patch discriminators (B x 25 x 25 values, C1b / C4) take up to
// one CTA per SM: block partials in double, summed in block order by the last CTA to arrive (fixed order: deterministic; the
// arrival counter resets itself).
__global__ void __launch_bounds__(256) loss_kernel(int kind, const float* __restrict__ x, int64_t count,
                                                   const float* __restrict__ label_vec, int64_t per, float label_const,
                                                   double n_total, float* __restrict__ loss_out, float* __restrict__ dx,
                                                   double* __restrict__ part, int* __restrict__ counter) {
  __shared__ double red[256];
  __shared__ int s_last;
  const double eps = 1e-12;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
    double t = label_vec ? (double)label_vec[i / per] : (double)label_const;
    double xd = (double)x[i];
    double l, g;
    if (kind == LOSS_BCE) {
      l = -(t * log(xd + eps) + (1.0 - t) * log(1.0 - xd + eps));
      g = -(t - xd) / ((1.0 - xd + eps) * (xd + eps)) / n_total;
    } else {
      double d = xd - t;
      l = d * d;
      g = 2.0 * d / n_total;
    }
    acc += l;
    if (dx) dx[i] = (float)g;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) loss_out[0] = (float)(red[0] / n_total);
    return;
  }
  if (threadIdx.x == 0) {
    part[blockIdx.x] = red[0];
    __threadfence();
    const int old = atomicAdd(counter, 1);
    s_last = old == (int)gridDim.x - 1;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double sum = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) sum += __ldcg(part + b);
    loss_out[0] = (float)(sum / n_total);
  }
}
void k_loss(St st, int kind, const float* x, int64_t count, const float* label_vec, int64_t per, float label_const,
            double n_total, float* loss_out, float* dx) {
  int nb = 1;
  if (st.ws && st.ws->lpart && count > 4096) nb = (int)std::min<int64_t>((count + 1023) / 1024, 148);
  loss_kernel<<<nb, 256, 0, st.s>>>(kind, x, count, label_vec, per > 0 ? per : 1, label_const, n_total, loss_out, dx,
                                    st.ws ? st.ws->lpart : nullptr, st.ws ? st.ws->lcounter : nullptr);
  DSR_LAUNCHED(st, "criterion", 8.0 * count, WORK_BYTES);
}

// per-sample sum((real - fake)^2) / div  -> becomes D's target on the fake pass (train.lua:237-245)
__global__ void __launch_bounds__(256) pixel_mse_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                                        float* __restrict__ out, int64_t per_sample, float div) {
  __shared__ double red[256];
  const float* r = real + (int64_t)blockIdx.x * per_sample;
  const float* f = fake + (int64_t)blockIdx.x * per_sample;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < per_sample; i += 256) {
    float d = r[i] - f[i];
    acc += (double)d * (double)d;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(red[0] / (double)div);
}
void k_pixel_mse(St st, const float* real, const float* fake, float* out, int n, int64_t per_sample, float div) {
  pixel_mse_kernel<<<n, 256, 0, st.s>>>(real, fake, out, per_sample, div);
  DSR_LAUNCHED(st, "pixel_mse", 8.0 * n * per_sample, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// optim.adam (Torch7 form): x -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)
// 28 B/param: read p,g,m,v ; write p,m,v -- one pass (the reference issues 7 tensor ops).
// ------------------------------------------------------------------------------------------
__global__ void adam_prep_kernel(int64_t* t, float* step, double lr, double b1, double b2) {
  int64_t tt = *t + 1;
  *t = tt;
  double bc1 = 1.0 - pow(b1, (double)tt);
  double bc2 = 1.0 - pow(b2, (double)tt);
  *step = (float)(lr * sqrt(bc2) / bc1);
}
void k_adam_prep(St st, int64_t* dev_t, float* dev_step, double lr, double beta1, double beta2) {
  adam_prep_kernel<<<1, 1, 0, st.s>>>(dev_t, dev_step, lr, beta1, beta2);
  DSR_LAUNCHED(st, "adam_prep", 16.0, WORK_BYTES);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t count, const float* __restrict__ step_p,
                                                   float b1, float b2, float omb1, float omb2, float eps) {
  const float step = __ldg(step_p);
  int64_t n4 = count >> 2;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    mv.x = mv.x * b1 + omb1 * gv.x; mv.y = mv.y * b1 + omb1 * gv.y;
    mv.z = mv.z * b1 + omb1 * gv.z; mv.w = mv.w * b1 + omb1 * gv.w;
    vv.x = vv.x * b2 + omb2 * gv.x * gv.x; vv.y = vv.y * b2 + omb2 * gv.y * gv.y;
    vv.z = vv.z * b2 + omb2 * gv.z * gv.z; vv.w = vv.w * b2 + omb2 * gv.w * gv.w;
    pv.x -= step * (mv.x / (sqrtf(vv.x) + eps)); pv.y -= step * (mv.y / (sqrtf(vv.y) + eps));
    pv.z -= step * (mv.z / (sqrtf(vv.z) + eps)); pv.w -= step * (mv.w / (sqrtf(vv.w) + eps));
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (n4 << 2) + i0; i < count; i += stride) {
    float gg = g[i];
    float mm = m[i] * b1 + omb1 * gg;
    float vv = v[i] * b2 + omb2 * gg * gg;
    p[i] -= step * (mm / (sqrtf(vv) + eps));
    m[i] = mm;
    v[i] = vv;
  }
}
void k_adam(St st, float* p, const float* g, float* m, float* v, int64_t count, const float* dev_step,
              double beta1, double beta2, double eps) {
  if (count <= 0) return;
  adam_kernel<<<ew_grid(count), 256, 0, st.s>>>(p, g, m, v, count, dev_step, (float)beta1, (float)beta2,
                                                 (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps);
  DSR_LAUNCHED(st, "adam", 28.0 * count, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
__global__ void fill_kernel(float* __restrict__ p, int64_t count, float v) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) p[i] = v;
}
// stage = [rmean * inv_world (nbn) | rvar * inv_world (nbn) | losses (3, as they are) | pad]: what a net's last gradient bucket carries
__global__ void stage_pack_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, int nbn, const float* __restrict__ losses,
                                  float* __restrict__ stage, float inv_world) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbn) { stage[i] = rmean[i] * inv_world; stage[nbn + i] = rvar[i] * inv_world; }
  if (i < 4) stage[2 * nbn + i] = (losses && i < 3) ? losses[i] : 0.f;
}
__global__ void stage_unpack_kernel(const float* __restrict__ stage, float* __restrict__ rmean, float* __restrict__ rvar, int nbn,
                                    float* __restrict__ losses) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nbn && rmean) { rmean[i] = stage[i]; rvar[i] = stage[nbn + i]; }
  if (i < 3 && losses) losses[i] = stage[2 * nbn + i];
}
void k_stage_pack(St st, const float* rmean, const float* rvar, int nbn, const float* losses, float* stage, float inv_world) {
  stage_pack_kernel<<<(std::max(nbn, 4) + 255) / 256, 256, 0, st.s>>>(rmean, rvar, nbn, losses, stage, inv_world);
  DSR_LAUNCHED(st, "stage_pack", 16.0 * nbn + 32, WORK_BYTES);
}
void k_stage_unpack(St st, const float* stage, float* rmean, float* rvar, int nbn, float* losses) {
  stage_unpack_kernel<<<(std::max(nbn, 4) + 255) / 256, 256, 0, st.s>>>(stage, rmean, rvar, nbn, losses);
  DSR_LAUNCHED(st, "stage_unpack", 16.0 * nbn + 32, WORK_BYTES);
}
__global__ void dacc_kernel(double* __restrict__ dst, const double* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
// dst[i] += src[i] (whole-batch BatchNorm sums accumulated over micro-batches)
void k_dacc(St st, double* dst, const double* src, int n) {
  dacc_kernel<<<(n + 255) / 256, 256, 0, st.s>>>(dst, src, n);
  DSR_LAUNCHED(st, "bn_dacc", 24.0 * n, WORK_BYTES);
}
void k_fill(St st, float* p, int64_t count, float v) {
  if (count <= 0) return;
  fill_kernel<<<sc_grid(count), 256, 0, st.s>>>(p, count, v);
  DSR_LAUNCHED(st, "fill", 4.0 * count, WORK_BYTES);
}
__global__ void scale_kernel(float* __restrict__ p, int64_t count, float s) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) p[i] *= s;
}
void k_scale(St st, float* p, int64_t count, float s) {
  if (count <= 0) return;
  scale_kernel<<<sc_grid(count), 256, 0, st.s>>>(p, count, s);
  DSR_LAUNCHED(st, "scale", 8.0 * count, WORK_BYTES);
}
// L2 flush: write a buffer larger than L2 (bench hygiene)
__global__ void flush_kernel(float4* __restrict__ p, int64_t n4) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
void k_flush(St st, float* buf, int64_t count) {
  flush_kernel<<<NSM * 8, 256, 0, st.s>>>(reinterpret_cast<float4*>(buf), count / 4);
  DSR_LAUNCHED(st, "l2_flush", 4.0 * count, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// Patch extraction / re-assembly (train-gray-patch.lua:267-273,588-595, train-gray-patch-batch.lua:258-264,
// 434-442, train-gray-patch-batch-overlap.lua:393-399): the step right before the training path in the patch configs, a
// scalar Lua triple loop per pixel in the reference.  Single-channel images [K][H][W]; patch i of image k takes
//     patch[k*nper + i][a][b] = image[k][(i / line) * stride + a][(i % line) * stride + b]
// (line, stride) = (patchSize, patchSize) in the non-overlapping scripts -- the reference divides by patchSize, which is
// the number of patches per row only when fineSize / patchSize == patchSize -- and (overlapPatchLine, overlap) in the
// overlapping one.  Assembly is the inverse scatter of the non-overlapping form.
// ------------------------------------------------------------------------------------------
__global__ void extract_patches_kernel(const float* __restrict__ img, float* __restrict__ patches, int K, int H, int W, int p,
                                       int line, int nper, int stride) {
  const int64_t total = (int64_t)K * nper * p * p;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int b = (int)(idx % p);
    int64_t r = idx / p;
    const int a = (int)(r % p); r /= p;
    const int i = (int)(r % nper);
    const int k = (int)(r / nper);
    const int y = (i / line) * stride + a, x = (i % line) * stride + b;
    patches[idx] = (y < H && x < W) ? img[((int64_t)k * H + y) * W + x] : 0.f;
  }
}
__global__ void assemble_patches_kernel(const float* __restrict__ patches, float* __restrict__ img, int K, int H, int W, int p,
                                        int line, int nper, int stride) {
  // one thread per IMAGE pixel (gather form of the scatter: deterministic).  Where patches overlap (stride < p) the
  // reference's sequential loop lets the LAST patch in index order win; the gather picks that same patch.
  const int64_t total = (int64_t)K * H * W;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int x = (int)(idx % W);
    int64_t r = idx / W;
    const int y = (int)(r % H);
    const int k = (int)(r / H);
    // last (row block, col block) whose patch covers the pixel
    int rb = y / stride, cb = x / stride;
    const int rows = (nper + line - 1) / line;
    if (rb > rows - 1) rb = rows - 1;
    if (cb > line - 1) cb = line - 1;
    float v = img[idx];
    bool found = false;
    for (int rr = rb; rr >= 0 && !found && rr * stride + p > y; --rr)
      for (int cc = cb; cc >= 0 && cc * stride + p > x; --cc) {
        const int i = rr * line + cc;
        if (i < nper) { v = patches[(((int64_t)k * nper + i) * p + (y - rr * stride)) * p + (x - cc * stride)]; found = true; break; }
      }
    img[idx] = v;
  }
}
void k_extract_patches(St st, const float* img, float* patches, int K, int H, int W, int p, int line, int nper, int stride) {
  const int64_t total = (int64_t)K * nper * p * p;
  if (total <= 0) return;
  extract_patches_kernel<<<sc_grid(total), 256, 0, st.s>>>(img, patches, K, H, W, p, line, nper, stride);
  DSR_LAUNCHED(st, "extract_patches", 8.0 * total, WORK_BYTES);
}
void k_assemble_patches(St st, const float* patches, float* img, int K, int H, int W, int p, int line, int nper, int stride) {
  const int64_t total = (int64_t)K * H * W;
  if (total <= 0) return;
  assemble_patches_kernel<<<sc_grid(total), 256, 0, st.s>>>(patches, img, K, H, W, p, line, nper, stride);
  DSR_LAUNCHED(st, "assemble_patches", 8.0 * total, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// Overlap stitching by a minimum-error boundary cut (train-gray-patch-batch-overlap.lua:457-694).
// The reference walks the L x L patches in index order and every patch rewrites its whole p x p footprint, so the value of
// an image pixel is decided by the LAST patch covering it, (x, y) = (min(L-1, r/ov), min(L-1, c/ov)), and by that patch's
// last write: the left seam when y > 0, else the top seam when x > 0, else a plain copy.  All seams read generated pixels
// only, never stitched output, so they are independent: one CTA per image, one thread per seam for the (p x ov)-cell
// programme (float64 tables like the reference's DoubleTensors -> identical cuts), then one gather per pixel.
// Top-seam costs are taken against patch i-1 (reference quirk, :557) unless flags & 1.
// ------------------------------------------------------------------------------------------
#define STITCH_MAXP 32
#define STITCH_MAXOV 16
__global__ void __launch_bounds__(256) stitch_overlap_kernel(const float* __restrict__ patches, float* __restrict__ img, int H, int W,
                                                             int p, int L, int ov, int flags) {
  extern __shared__ unsigned char s_idx[];                 // [L*L][p]: pixels taken from the neighbour on each line of the seam
  const int n = L * L;
  const float* P = patches + (int64_t)blockIdx.x * n * p * p;
  double path[STITCH_MAXP * STITCH_MAXOV];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int x = i / L, y = i % L;
    if (i == 0) continue;
    const bool left = y != 0;
    const float* cur = P + (int64_t)i * p * p;
    const float* nb = P + (int64_t)((left || !(flags & 1)) ? i - 1 : i - L) * p * p;
    // line = the direction the cut runs along (rows for a left seam, columns for a top seam); cross = across the strip
    const int sl = left ? p : 1, sc = left ? 1 : p, noff = left ? p - ov : (p - ov) * p;
    for (int b = 0; b < ov; ++b) path[b] = fabs((double)nb[noff + b * sc] - (double)cur[b * sc]);
    for (int a = 1; a < p; ++a)
      for (int b = 0; b < ov; ++b) {
        const double* pr = path + (a - 1) * ov;
        double m = pr[b];
        if (b > 0) m = fmin(m, pr[b - 1]);
        if (b < ov - 1) m = fmin(m, pr[b + 1]);
        path[a * ov + b] = fabs((double)nb[noff + a * sl + b * sc] - (double)cur[a * sl + b * sc]) + m;
      }
    unsigned char* idx = s_idx + i * p;
    int k = 0;
    {
      const double* pr = path + (p - 1) * ov;
      double m = pr[0];
      for (int b = 1; b < ov; ++b) m = fmin(m, pr[b]);
      for (int b = 0; b < ov; ++b) if (pr[b] == m) k = b;            // the last minimum wins
    }
    idx[p - 1] = (unsigned char)(k + 1);
    for (int a = p - 2; a >= 0; --a) {
      const double* pr = path + a * ov;
      if (k == 0) k = (pr[0] <= pr[1]) ? 0 : 1;
      else if (k == ov - 1) k = (pr[ov - 1] <= pr[ov - 2]) ? ov - 1 : ov - 2;
      else {
        const double m3 = fmin(pr[k], fmin(pr[k - 1], pr[k + 1]));
        k = (pr[k] == m3) ? k : (pr[k + 1] == m3) ? k + 1 : k - 1;
      }
      idx[a] = (unsigned char)(k + 1);
    }
    (void)x;
  }
  __syncthreads();
  float* out = img + (int64_t)blockIdx.x * H * W;
  const int cover = (L - 1) * ov + p;
  for (int px = threadIdx.x; px < H * W; px += blockDim.x) {
    const int r = px / W, c = px % W;
    if (r >= cover || c >= cover) continue;                 // no patch reaches here: keeps the caller's value
    const int x = min(L - 1, r / ov), y = min(L - 1, c / ov);
    const int a = r - x * ov, b = c - y * ov, i = x * L + y;
    const float* cur = P + (int64_t)i * p * p;
    float v;
    if (y != 0) v = (b < s_idx[i * p + a]) ? P[(int64_t)(i - 1) * p * p + a * p + (p - ov + b)] : cur[a * p + b];
    else if (x != 0) v = (a < s_idx[i * p + b]) ? P[(int64_t)(i - L) * p * p + (p - ov + a) * p + b] : cur[a * p + b];
    else v = cur[a * p + b];
    out[px] = v;
  }
}
bool stitch_overlap_supported(int p, int L, int ov) {
  return p <= STITCH_MAXP && ov <= STITCH_MAXOV && ov >= 2 && p > ov && L >= 1 && (size_t)L * L * p <= 48 * 1024;
}
void k_stitch_overlap(St st, const float* patches, float* img, int K, int H, int W, int p, int L, int ov, int flags) {
  if (K <= 0) return;
  stitch_overlap_kernel<<<K, 256, (size_t)L * L * p, st.s>>>(patches, img, H, W, p, L, ov, flags);
  DSR_LAUNCHED(st, "stitch_overlap", 4.0 * K * ((double)H * W + (double)L * L * p * p), WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// Evaluation metrics of the eval sweeps: calPSNR (train-gray-3.lua:143-151) and calSSIM (:156-221) on
// batches of single-channel H x W images.  One CTA per image pair, deterministic.
//   PSNR: MSE = sum((a-b)^2) / (H*W);  10*log10(1/MSE), 99 when MSE == 0.
//   SSIM: images mapped to [0,255] by (x+1)/2*255, 11x11 Gaussian (sigma 1.5, normalised), 'full' convolution (zero padded,
//         (H+10) x (W+10) map), K1 = 0.01, K2 = 0.03, L = 255, mean of the SSIM map.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) psnr_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                   int64_t per) {
  __shared__ double red[256];
  const float* pa = a + (int64_t)blockIdx.x * per;
  const float* pb = b + (int64_t)blockIdx.x * per;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < per; i += 256) { const float d = pa[i] - pb[i]; acc += (double)(d * d); }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) {
    const double mse = red[0] / (double)per;
    out[blockIdx.x] = mse > 0.0 ? (float)(10.0 * log(1.0 / mse) / log(10.0)) : 99.f;
  }
}
void k_psnr(St st, const float* a, const float* b, float* out, int n, int64_t per) {
  if (n <= 0) return;
  psnr_kernel<<<n, 256, 0, st.s>>>(a, b, out, per);
  DSR_LAUNCHED(st, "psnr", 8.0 * n * per, WORK_BYTES);
}

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                   int H, int W) {
  __shared__ float g[11];
  __shared__ double red[256];
  if (threadIdx.x < 11) {
    // image.gaussian(11, 1.5/11): exp(-((i - 6) / 1.5)^2 / 2), i = 1..11; 2-D window = outer product / sum = (g/sum g) x (g/sum g)
    float s = 0.f;
    for (int i = 0; i < 11; ++i) { const float d = ((float)(i + 1) - 6.f) / 1.5f; s += expf(-d * d / 2.f); }
    const float d = ((float)(threadIdx.x + 1) - 6.f) / 1.5f;
    g[threadIdx.x] = expf(-d * d / 2.f) / s;
  }
  __syncthreads();
  const float* pa = a + (int64_t)blockIdx.x * H * W;
  const float* pb = b + (int64_t)blockIdx.x * H * W;
  const int Ho = H + 10, Wo = W + 10;
  const float C1 = (0.01f * 255.f) * (0.01f * 255.f), C2 = (0.03f * 255.f) * (0.03f * 255.f);
  double acc = 0.0;
  for (int o = threadIdx.x; o < Ho * Wo; o += 256) {
    const int oy = o / Wo, ox = o - oy * Wo;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int ky = 0; ky < 11; ++ky) {
      const int y = oy - ky;                      // full convolution: out[oy][ox] = sum_k w[k] * img[oy - ky][ox - kx]
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < 11; ++kx) {
        const int x = ox - kx;
        if (x < 0 || x >= W) continue;
        const float w = g[ky] * g[kx];
        const float u = (pa[y * W + x] + 1.f) * 0.5f * 255.f, v = (pb[y * W + x] + 1.f) * 0.5f * 255.f;
        m1 = fmaf(w, u, m1); m2 = fmaf(w, v, m2);
        s11 = fmaf(w, u * u, s11); s22 = fmaf(w, v * v, s22); s12 = fmaf(w, u * v, s12);
      }
    }
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
    acc += (double)(((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2)));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(red[0] / (double)(Ho * Wo));
}
void k_ssim(St st, const float* a, const float* b, float* out, int n, int H, int W) {
  if (n <= 0) return;
  ssim_kernel<<<n, 256, 0, st.s>>>(a, b, out, H, W);
  DSR_LAUNCHED(st, "ssim", 8.0 * n * H * W, WORK_BYTES);
}

// ------------------------------------------------------------------------------------------
// image.scale(..., 'bilinear') for enlarging sizes: the eval sweeps' baseline (train-gray-3.lua:399).  Separable, end points
// aligned, width pass then height pass with a float32 intermediate -- evaluated here per output pixel with the same
// operations in the same order (no fused multiply-add: the CPU code has none).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coord(int d, int src_len, int dst_len, int& i, float& f, bool& copy) {
  if (dst_len == src_len || src_len == 1) { i = (src_len == 1) ? 0 : d; f = 0.f; copy = true; return; }
  if (d == dst_len - 1) { i = src_len - 1; f = 0.f; copy = true; return; }
  const float scale = __fdiv_rn((float)(src_len - 1), (float)(dst_len - 1));
  const float sf = __fmul_rn((float)d, scale);
  i = (int)sf;
  f = __fsub_rn(sf, (float)i);
  copy = false;
}
__device__ __forceinline__ float lin_mix(float a, float b, float f) {
  return __fadd_rn(__fmul_rn(__fsub_rn(1.f, f), a), __fmul_rn(f, b));
}
__global__ void scale_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W, int DH, int DW) {
  const int64_t total = (int64_t)N * DH * DW;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
    const int x = (int)(idx % DW);
    const int64_t r = idx / DW;
    const int y = (int)(r % DH);
    const float* s = src + (r / DH) * H * W;
    int xi, yi; float fx, fy; bool cx, cy;
    lin_coord(x, W, DW, xi, fx, cx);
    lin_coord(y, H, DH, yi, fy, cy);
    const float t0 = cx ? s[yi * W + xi] : lin_mix(s[yi * W + xi], s[yi * W + xi + 1], fx);
    float v = t0;
    if (!cy) {
      const float t1 = cx ? s[(yi + 1) * W + xi] : lin_mix(s[(yi + 1) * W + xi], s[(yi + 1) * W + xi + 1], fx);
      v = lin_mix(t0, t1, fy);
    }
    dst[idx] = v;
  }
}
void k_scale_bilinear(St st, const float* src, float* dst, int N, int H, int W, int DH, int DW) {
  const int64_t total = (int64_t)N * DH * DW;
  if (total <= 0) return;
  scale_bilinear_kernel<<<sc_grid(total), 256, 0, st.s>>>(src, dst, N, H, W, DH, DW);
  DSR_LAUNCHED(st, "scale_bilinear", 4.0 * (total + (double)N * H * W), WORK_BYTES);
}
