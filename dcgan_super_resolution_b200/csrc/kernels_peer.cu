// One-shot all-reduce over NVLink peer memory, fused into the kernels that need the result.
//
// What it replaces: with sync_bn=1 every BatchNorm of a data-parallel step needs the cross-rank sums of 2*C doubles before it
// can normalise (forward: sum x, sum x^2; backward: sum g, sum g*xhat) -- big-batch semantics of SpatialBatchNormalization
// (train.lua:100-109) over a sharded minibatch.  Through NCCL that is one latency-bound collective kernel + one finalize kernel
// per BatchNorm and direction, all on the critical path.  Here the kernel that finalises the statistics does the exchange
// itself: every rank PUSHES its 2*C sums into a receive slot it owns in every peer's memory (plain stores through the NVSwitch
// to cudaIpc-mapped buffers), releases a per-(slot, source) flag with a system-scope store, waits for the flags of all sources
// in its own memory, and adds the W contributions in rank order (the same order on every rank: results are bit-identical
// across ranks, which the replicated parameters need).  No rank ever READS remote memory, so nothing waits for a round trip.
//
// Slots: call k uses slot k % PEER_SLOTS and flag value k (64-bit, never reset).  A rank can finish call k only after every
// peer has entered call k (it needs their data), so ranks are never more than one call apart and two slots would do; four are
// kept.  The call counter lives in device memory and is advanced by the kernel, so a captured CUDA graph replays correctly.
// A rank that waits longer than PEER_TIMEOUT_NS raises *err (host-mapped) and goes on: a lost peer shows up as an error of the
// next API call instead of a hung GPU.
#include "common.h"

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PEER_TIMEOUT_NS 4000000000ull

// in[n] (local sums) -> out[n] (sums over all ranks).  Whole block; n <= p.nmax.  in / out may alias.
__device__ void peer_allreduce_block(const PeerAR& p, const double* in, double* out, int n) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *p.seq + 1;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int slot = (int)(seq % PEER_SLOTS);
  const size_t mine = ((size_t)slot * PEER_MAXW + p.rank) * p.nmax;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = in[i];
    for (int r = 0; r < p.world; ++r) p.data[r][mine + i] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.flags[threadIdx.x] + slot * PEER_MAXW + p.rank, seq);
    const unsigned long long* f = p.flags[p.rank] + slot * PEER_MAXW + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(f) < seq) {
      if (globaltimer_ns() - t0 > PEER_TIMEOUT_NS) { *p.err = 1; break; }
      __nanosleep(20);
    }
  }
  __syncthreads();
  const double* own = p.data[p.rank] + (size_t)slot * PEER_MAXW * p.nmax;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < p.world; ++r) s += __ldcv(own + (size_t)r * p.nmax + i);
    out[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.seq = seq;
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerAR p, const double* __restrict__ in, double* __restrict__ out, int n) {
  peer_allreduce_block(p, in, out, n);
}

// BatchNorm statistics tail of the sync_bn forward: cross-rank sums, then mean / invstd / running statistics (the arithmetic of
// bn_finalize_kernel, kernels_bw.cu) in the same kernel.
__global__ void __launch_bounds__(512) bn_finalize_peer_kernel(const PeerAR p, double* __restrict__ sums, int C, double n_total, float eps,
                                                               float momentum, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                                               float* __restrict__ running_mean, float* __restrict__ running_var) {
  peer_allreduce_block(p, sums, sums, 2 * C);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
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
}

}  // namespace

void k_peer_allreduce(St st, const PeerAR& p, const double* in, double* out, int n) {
  peer_allreduce_kernel<<<1, 512, 0, st.s>>>(p, in, out, n);
  DSR_LAUNCHED(st, "peer_allreduce", 8.0 * n * (p.world + 1), WORK_BYTES);
}

void k_bn_finalize_peer(St st, const PeerAR& p, double* sums, int C, double n_total, float eps, float momentum, float* save_mean,
                        float* save_invstd, float* running_mean, float* running_var) {
  bn_finalize_peer_kernel<<<1, 512, 0, st.s>>>(p, sums, C, n_total, eps, momentum, save_mean, save_invstd, running_mean, running_var);
  DSR_LAUNCHED(st, "bn_finalize_peer", 8.0 * 2 * C * (p.world + 1) + 32.0 * C, WORK_BYTES);
}
