// kernels_tc2.cu -- persistent, wide-tile per-tap tcgen05 implicit GEMM for the tensor-bound convolutions (FAST_TF32):
// nn.SpatialConvolution / nn.SpatialFullConvolution forward and updateGradInput of the discriminator's inner layers
// (train.lua:124-131) and of the generator's inner layers at ngf >= 64 (train-gray-patch.lua:60-70), whose weights do not
// fit in shared memory (so the weights-resident halo kernel cannot take them).
//
// What bounds kernels_tc.cu's one-tile-per-CTA kernel on these layers is the L2 -> SM path: a 128 pixel x 128 cout x 32
// channel step moves 32 KB of fp32 operands into shared memory for 4 tcgen05.mma (271 tensor cycles) = 118 B/clk/SM, against
// ~43 B/clk/SM the L2 delivers chip-wide (B300_MICROARCH: LTS cap ~6300 B/clk), which only the L2's merging of concurrent
// identical requests lifts to the measured 45-56 % of the TF32 peak.  Thread-block-cluster multicast of the weight tile does not
// help (measured; the L2 already merges what a cluster of <= 4 would share).  Fewer bytes per flop does:
//
//   * a work item is 256 x 128 or 128 x 256 (pixels x couts): two pixel tiles share one weight tile, or one pixel tile
//     meets a 256-cout weight tile -- 48 KB per 542 tensor cycles = 88 B/clk/SM, 25 % less operand traffic per flop;
//   * one persistent CTA per SM owns all 512 TMEM columns as two accumulator buffers: the epilogue of item i (tcgen05.ld ->
//     activation -> 256-bit NHWC stores) overlaps the TMA + MMA of item i + 1 (the one-tile kernel needed a second resident
//     CTA for that, which halved its ring);
//   * the whole shared memory is one 4-stage ring of {A tile(s), weight tile} stages (192 KB in flight per SM).
//
// Operands as in kernels_tc.cu: A = one TMA box of the NHWC activation per (pixel tile, tap, 32-channel chunk) at a shifted
// coordinate (zero fill = padding; stride-2 gathers through the 5-D parity view), B = the pre-tiled, pre-swizzled weight
// images (1-D bulk copies), both K-major SWIZZLE_128B, kind::tf32, fp32 accumulation in TMEM.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define NSM 148
#define TC2_MAXCLS 4
#define TC2_THREADS 192

struct Tc2Params {
  int N, Hg, Wg, Ho, Wo, Co;
  int so, si, Ci;
  int TW, TH, TB, tiles_x, tiles_y, ntiles_m;
  int kchunks, ksteps_last;
  int MT, BN, ngroups_m, ntiles_n, nwork, nstage;
  int n_fast;                      // work order: 0 = pixel groups fastest (neighbouring CTAs stream the same weight tiles), 1 = cout tiles fastest
  int bt_rows, bt_per_n;           // weight images of bt_rows couts; bt_per_n images make one BN-cout stage
  int a_tile_bytes, stage_bytes, acc_cols, tmem_cols;
  int ncls, oy0[TC2_MAXCLS], ox0[TC2_MAXCLS], ntaps[TC2_MAXCLS];
  const float* bt[TC2_MAXCLS];
  int act;
  float neg;
  short oy[TC2_MAXCLS][DSR_MAX_TAPS], ox[TC2_MAXCLS][DSR_MAX_TAPS], py[TC2_MAXCLS][DSR_MAX_TAPS], px[TC2_MAXCLS][DSR_MAX_TAPS];
};

__device__ __forceinline__ void tc2_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

struct TcMapsW { CUtensorMap w[TC2_MAXCLS]; };      // pair kernel: the weight image buffers as 2-D maps of 128-byte rows
struct Tc2Item { int cls, ntile, mg; };
__device__ __forceinline__ Tc2Item tc2_item(const Tc2Params& p, int w) {
  Tc2Item it;
  if (p.n_fast) { it.ntile = w % p.ntiles_n; w /= p.ntiles_n; it.mg = w % p.ngroups_m; it.cls = w / p.ngroups_m; }
  else { it.mg = w % p.ngroups_m; w /= p.ngroups_m; it.ntile = w % p.ntiles_n; it.cls = w / p.ntiles_n; }
  return it;
}

template <int ACT>
__device__ __forceinline__ void tc2_epilogue(const Tc2Params& p, float* __restrict__ out, uint32_t tmem_base, uint64_t* acc_full,
                                             uint64_t* acc_empty, int warp, int lane) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;                  // tile row = pixel
  const int w = r % p.TW;
  const int h = (r / p.TW) % p.TH;
  const int b = r / (p.TW * p.TH);
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  int it = 0;
  for (int wi = blockIdx.x; wi < p.nwork; wi += gridDim.x, ++it) {
    const Tc2Item item = tc2_item(p, wi);
    const int buf = it & 1;
    const uint32_t aph = (uint32_t)(it >> 1) & 1u;
    const int n0 = item.ntile * p.BN;
    mbar_wait(smem_u32(&acc_full[buf]), aph);
    tc_fence_after();
    for (int m = 0; m < p.MT; ++m) {
      int tile = item.mg * p.MT + m;
      if (tile >= p.ntiles_m) break;
      const int tx = tile % p.tiles_x; tile /= p.tiles_x;
      const int ty = tile % p.tiles_y; tile /= p.tiles_y;
      const int n = tile * p.TB + b, gy = ty * p.TH + h, gx = tx * p.TW + w;
      const bool valid = n < p.N && gy < p.Hg && gx < p.Wg;
      float* orow = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0[item.cls]) * p.Wo + gx * p.so + p.ox0[item.cls]) * p.Co + n0;
      const uint32_t cbase = lane_base + (uint32_t)(buf * p.acc_cols + m * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(cbase + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) store_row<ACT, 32>(orow + c0, v, n0 + c0, p.Co, p.neg);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) tc2_arrive(smem_u32(&acc_empty[buf]));
  }
}

__global__ void __launch_bounds__(TC2_THREADS, 1) tapconv_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const Tc2Params p,
                                                                     float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstage * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstage;
  uint64_t* acc_full = bars + 2 * p.nstage;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    for (int s = 0; s < p.nstage; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 4); }
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
  const uint32_t b_off = (uint32_t)(p.MT * p.a_tile_bytes);          // weight tile inside a stage

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t img_bytes = (uint32_t)p.bt_rows * 128u;
      const uint32_t bytes = (uint32_t)(p.MT * 128 * 128) + (uint32_t)p.bt_per_n * img_bytes;
      for (int wi = blockIdx.x; wi < p.nwork; wi += gridDim.x) {
        const Tc2Item item = tc2_item(p, wi);
        const int cls = item.cls;
        const int nk = p.ntaps[cls] * p.kchunks;
        int b0[2], gy0[2], gx0[2];
        for (int m = 0; m < p.MT; ++m) {
          int tile = item.mg * p.MT + m;                            // tiles past the end: batch coordinate out of range -> zero fill
          const int tx = tile % p.tiles_x; tile /= p.tiles_x;
          const int ty = tile % p.tiles_y; tile /= p.tiles_y;
          b0[m] = tile * p.TB; gy0[m] = ty * p.TH; gx0[m] = tx * p.TW;
        }
        const float* wsrc = p.bt[cls] + (size_t)item.ntile * p.bt_per_n * nk * (p.bt_rows * 32);
        int t = 0, c = 0;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full[s]);
          mbar_expect_tx(fb, bytes);
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          for (int m = 0; m < p.MT; ++m) {
            if (p.si == 1)
              tma_load_4d(st + m * p.a_tile_bytes, &mapA, fb, c * 32, gx0[m] + p.ox[cls][t], gy0[m] + p.oy[cls][t], b0[m]);
            else
              tma_load_5d(st + m * p.a_tile_bytes, &mapA, fb, p.px[cls][t] * p.Ci + c * 32, gx0[m] + p.ox[cls][t], p.py[cls][t],
                          gy0[m] + p.oy[cls][t], b0[m]);
          }
          for (int j = 0; j < p.bt_per_n; ++j)
            bulk_load(st + b_off + j * img_bytes, wsrc + ((size_t)j * nk + kb) * (p.bt_rows * 32), img_bytes, fb);
          if (++c == p.kchunks) { c = 0; ++t; }
          if (++s == p.nstage) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int wi = blockIdx.x; wi < p.nwork; wi += gridDim.x, ++it) {
      const Tc2Item item = tc2_item(p, wi);
      const int nk = p.ntaps[item.cls] * p.kchunks;
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), aph ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem_base + (uint32_t)(buf * p.acc_cols);
      int kc = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        tc_fence_after();
        const int ksteps = kc == p.kchunks - 1 ? p.ksteps_last : 4;          // the last K chunk of a tap may be partial
        if (++kc == p.kchunks) kc = 0;
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          const uint64_t bd = make_kmajor_desc(st + b_off, 32);
          for (int m = 0; m < p.MT; ++m) {
            const uint64_t ad = make_kmajor_desc(st + m * p.a_tile_bytes, 32);
            for (int k = 0; k < ksteps; ++k)
              umma_tf32(dbase + (uint32_t)(m * p.BN), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty[s]));
          if (kb == nk - 1) umma_commit(smem_u32(&acc_full[buf]));
        }
        __syncwarp();
        if (++s == p.nstage) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> activation -> NHWC global =====================
    switch (p.act) {
      case ACT_RELU: tc2_epilogue<ACT_RELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane); break;
      case ACT_LRELU: tc2_epilogue<ACT_LRELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane); break;
      case ACT_TANH: tc2_epilogue<ACT_TANH>(p, out, tmem_base, acc_full, acc_empty, warp, lane); break;
      case ACT_SIGMOID: tc2_epilogue<ACT_SIGMOID>(p, out, tmem_base, acc_full, acc_empty, warp, lane); break;
      default: tc2_epilogue<ACT_NONE>(p, out, tmem_base, acc_full, acc_empty, warp, lane); break;
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
// CTA-pair variant (cta_group::2) for 256-cout work items: the two SMs of a TPC run ONE 256 pixel x 256 cout x 8 MMA.
// Each CTA loads its own 128-pixel A tile and HALF of the weight tile (one 128-cout image): 32 KB per 542 tensor cycles per
// SM = 59 B/clk instead of 88 -- the L2 -> SM path (measured: ~13.5 TB/s chip-wide, lts__t_bytes of the single-CTA kernel equals
// its operand traffic: the L2 merges nothing) is what bounds these layers.  Protocol:
//   * both CTAs' TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address with
//     the cluster-rank bit cleared); the leader arms it with the bytes of both;
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2; its tcgen05.commit multicasts to the empty / acc_full
//     barriers of both CTAs, so each producer and each epilogue waits locally;
//   * accumulator rows 0..127 live in the leader's TMEM, 128..255 in the peer's; each CTA's epilogue drains its own half and
//     arrives on the leader's acc_empty barrier (the peer through mapa + a remote arrive).
// ------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void tc2p_epilogue(const Tc2Params& p, float* __restrict__ out, uint32_t tmem_base, uint64_t* acc_full,
                                              uint64_t* acc_empty, int warp, int lane, uint32_t crank, int pair, int npairs) {
  const int q = warp & 3;
  const int r = q * 32 + lane;
  const int w = r % p.TW;
  const int h = (r / p.TW) % p.TH;
  const int b = r / (p.TW * p.TH);
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  int it = 0;
  for (int wi = pair; wi < p.nwork; wi += npairs, ++it) {
    const Tc2Item item = tc2_item(p, wi);
    const int buf = it & 1;
    const uint32_t aph = (uint32_t)(it >> 1) & 1u;
    const int n0 = item.ntile * p.BN;
    mbar_wait(smem_u32(&acc_full[buf]), aph);
    tc_fence_after();
    int tile = item.mg * 2 + (int)crank;
    if (tile < p.ntiles_m) {
      const int tx = tile % p.tiles_x; tile /= p.tiles_x;
      const int ty = tile % p.tiles_y; tile /= p.tiles_y;
      const int n = tile * p.TB + b, gy = ty * p.TH + h, gx = tx * p.TW + w;
      const bool valid = n < p.N && gy < p.Hg && gx < p.Wg;
      float* orow = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0[item.cls]) * p.Wo + gx * p.so + p.ox0[item.cls]) * p.Co + n0;
      const uint32_t cbase = lane_base + (uint32_t)(buf * p.acc_cols);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(cbase + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) store_row<ACT, 32>(orow + c0, v, n0 + c0, p.Co, p.neg);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (crank == 0) tc2_arrive(smem_u32(&acc_empty[buf]));
      else mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0);
    }
  }
}

__global__ void __launch_bounds__(TC2_THREADS, 1) tapconv_tc2_pair_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                          const __grid_constant__ TcMapsW mapsW, const Tc2Params p,
                                                                          float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstage * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstage;
  uint64_t* acc_full = bars + 2 * p.nstage;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    for (int s = 0; s < p.nstage; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // both CTAs' barriers are initialised before any TMA / commit / remote arrive targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t b_off = (uint32_t)p.a_tile_bytes;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion on the leader's barrier) =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes_pair = 2u * (uint32_t)p.stage_bytes;
      for (int wi = pair; wi < p.nwork; wi += npairs) {
        const Tc2Item item = tc2_item(p, wi);
        const int cls = item.cls;
        const int nk = p.ntaps[cls] * p.kchunks;
        int tile = item.mg * 2 + (int)crank;
        const int tx = tile % p.tiles_x; tile /= p.tiles_x;
        const int ty = tile % p.tiles_y; tile /= p.tiles_y;
        const int b0 = tile * p.TB, gy0 = ty * p.TH, gx0 = tx * p.TW;
        const int wrow0 = ((item.ntile * 2 + (int)crank) * nk) * 128;         // first row of this CTA's image sequence in the image buffer
        int t = 0, c = 0;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t fb_local = smem_u32(&full[s]);
          const uint32_t fb = fb_local & 0xFEFFFFFFu;                          // the leader's barrier (cluster rank bit cleared)
          if (crank == 0) mbar_expect_tx(fb_local, bytes_pair);
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          if (p.si == 1)
            tma_load_4d_2sm(st, &mapA, fb, c * 32, gx0 + p.ox[cls][t], gy0 + p.oy[cls][t], b0);
          else
            tma_load_5d_2sm(st, &mapA, fb, p.px[cls][t] * p.Ci + c * 32, gx0 + p.ox[cls][t], p.py[cls][t], gy0 + p.oy[cls][t], b0);
          tma_load_2d_2sm(st + b_off, &mapsW.w[cls], fb, 0, wrow0 + kb * 128);
          if (++c == p.kchunks) { c = 0; ++t; }
          if (++s == p.nstage) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ===================== MMA issuer (leader only) =====================
    // instruction descriptor: D = f32, A = B = tf32, both K-major, N = 256, M = 256 (128 rows per CTA)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((256u >> 4) << 24);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int wi = pair; wi < p.nwork; wi += npairs, ++it) {
      const Tc2Item item = tc2_item(p, wi);
      const int nk = p.ntaps[item.cls] * p.kchunks;
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), aph ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem_base + (uint32_t)(buf * p.acc_cols);
      int kc = 0;
      for (int kb = 0; kb < nk; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        tc_fence_after();
        const int ksteps = kc == p.kchunks - 1 ? p.ksteps_last : 4;
        if (++kc == p.kchunks) kc = 0;
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
          const uint64_t ad = make_kmajor_desc(st, 32);
          const uint64_t bd = make_kmajor_desc(st + b_off, 32);
          for (int k = 0; k < ksteps; ++k)
            umma_tf32_2sm(dbase, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(smem_u32(&empty[s]));
          if (kb == nk - 1) umma_commit_2sm(smem_u32(&acc_full[buf]));
        }
        __syncwarp();
        if (++s == p.nstage) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 2) {
    switch (p.act) {
      case ACT_RELU: tc2p_epilogue<ACT_RELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_LRELU: tc2p_epilogue<ACT_LRELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_TANH: tc2p_epilogue<ACT_TANH>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_SIGMOID: tc2p_epilogue<ACT_SIGMOID>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      default: tc2p_epilogue<ACT_NONE>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA frees its TMEM / leaves while the pair's MMAs, commits or remote arrives may still target it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pow2_ge2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int floordiv2_2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

struct Tc2Cfg { Tc2Params p; size_t smem; int grid; double flops; bool pair; };

// co-resident CTA pairs of the pair kernel at its shared-memory footprint (74 on a full B200; fewer if a TPC has a disabled SM)
static int tc2_max_clusters() {
  static int v = -1;
  if (v >= 0) return v;
  v = 0;
  if (cudaFuncSetAttribute(tapconv_tc2_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return v; }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(2 * NSM); lc.blockDim = dim3(TC2_THREADS); lc.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tapconv_tc2_pair_kernel, &lc) != cudaSuccess) { cudaGetLastError(); return v; }
  v = std::min(n, NSM / 2);
  return v;
}

static bool tc2_cfg(const TapGeom* cls, int ncls, const float* const* bt, Tc2Cfg& c) {
  if (!tc_encode_fn() || !tc_tapconv_multi_ok(cls, ncls) || !bt) return false;
  const char* me = getenv("DCGANSR_TC2");
  const int mode = me ? atoi(me) : 1;     // 0: off, 1: where it pays, 2: wherever it runs
  if (mode == 0) return false;
  const TapGeom& g = cls[0];
  if (g.Ci % 32 != 0 && !(g.Ci % 8 == 0 && g.Ci >= 24)) return false;                    // 32-float K blocks (a partial last block per tap)
  for (int i = 0; i < ncls; ++i)
    if (!bt[i] || cls[i].ntaps < 1 || cls[i].ntaps > DSR_MAX_TAPS) return false;
  if (g.Ci % 32 != 0) return false;                                                       // the weight images exist for whole blocks only
  Tc2Params& p = c.p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hg = g.Hg; p.Wg = g.Wg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.so = g.so; p.si = g.si; p.Ci = g.Ci; p.ncls = ncls;
  p.kchunks = (g.Ci + 31) / 32;
  p.ksteps_last = (g.Ci - (p.kchunks - 1) * 32) >> 3;
  p.TW = std::min(pow2_ge2(g.Wg), 128);
  p.TH = std::min(pow2_ge2(g.Hg), 128 / p.TW);
  p.TB = 128 / (p.TW * p.TH);
  p.tiles_x = (g.Wg + p.TW - 1) / p.TW;
  p.tiles_y = (g.Hg + p.TH - 1) / p.TH;
  p.ntiles_m = ((g.N + p.TB - 1) / p.TB) * p.tiles_y * p.tiles_x;
  p.bt_rows = tc_bt_rows(g.Co);
  if (p.bt_rows < 64 || p.bt_rows > 128 || p.bt_rows % 16) return false;
  const int n_img = (g.Co + p.bt_rows - 1) / p.bt_rows;
  if (n_img >= 2 && n_img % 2 == 0 && p.bt_rows == 128) { p.BN = 256; p.bt_per_n = 2; p.MT = 1; }
  else { p.BN = p.bt_rows; p.bt_per_n = 1; p.MT = 2; }
  if (const char* e = getenv("DCGANSR_TC2_SHAPE")) {                   // experiments: "1x128" one pixel tile x one image (the old tile)
    if (!strcmp(e, "1x128")) { p.BN = p.bt_rows; p.bt_per_n = 1; p.MT = 1; }
    if (!strcmp(e, "2x128")) { p.BN = p.bt_rows; p.bt_per_n = 1; p.MT = 2; }
  }
  p.ntiles_n = n_img / p.bt_per_n;
  p.ngroups_m = (p.ntiles_m + p.MT - 1) / p.MT;
  p.nwork = ncls * p.ntiles_n * p.ngroups_m;
  p.n_fast = getenv("DCGANSR_TC2_NFAST") ? 1 : 0;
  p.a_tile_bytes = 128 * 128;
  p.stage_bytes = p.MT * p.a_tile_bytes + p.BN * 128;
  // CTA pairs (cta_group::2) for the 256-cout items: a pair owns two pixel tiles, each CTA stages its tile + half the weight tile.
  // Measured (scripts/exp/tc2_pair_sweep.py): 1.2 - 1.5x over the single-CTA items from 32 pair items up (C 128->256 at C1b size
  // 422 -> 296 us = 929 TFLOP/s; C5's FC 512->256 996 TFLOP/s = 0.91 of the 1.09 PFLOP/s TF32 pipe peak)
  c.pair = false;
  {
    const char* pe = getenv("DCGANSR_TC2_PAIR");
    const int pmode = pe ? atoi(pe) : 1;
    const int nwork2 = ncls * p.ntiles_n * ((p.ntiles_m + 1) / 2);
    const int maxcl = pmode ? tc2_max_clusters() : 0;
    if (pmode && p.BN == 256 && p.bt_per_n == 2 && p.MT == 1 && maxcl >= 32 && (pmode == 2 || nwork2 >= 32)) {
      c.pair = true;
      p.ngroups_m = (p.ntiles_m + 1) / 2;
      p.nwork = nwork2;
      p.stage_bytes = p.a_tile_bytes + 128 * 128;
    }
  }
  p.nstage = std::max(2, std::min(8, (200 * 1024) / p.stage_bytes));
  p.acc_cols = c.pair ? p.BN : p.MT * p.BN;
  p.tmem_cols = std::max(32, pow2_ge2(2 * p.acc_cols));
  if (p.tmem_cols > 512) return false;
  // one CTA per SM walks ceil(nwork / 148) items.  Measured per layer and batch against the one-tile kernel (scripts/exp/tc2_sweep.py):
  // 1.15 - 1.6x from ~128 items up whatever the wave count, even at 32 - 64 items, slower below (C 256->512 at 8 x 8, batch 64:
  // 16 items, 0.72x -- the one-tile kernel's split-K fills the machine there)
  c.grid = c.pair ? 2 * std::min(p.nwork, tc2_max_clusters()) : std::min(p.nwork, NSM);
  if (mode == 1 && !c.pair && p.nwork < 96) return false;
  c.flops = 0;
  for (int i = 0; i < ncls; ++i) {
    p.bt[i] = bt[i];
    p.oy0[i] = cls[i].oy0; p.ox0[i] = cls[i].ox0; p.ntaps[i] = cls[i].ntaps;
    for (int t = 0; t < cls[i].ntaps; ++t) {
      if (g.si == 1) { p.oy[i][t] = (short)cls[i].dy[t]; p.ox[i][t] = (short)cls[i].dx[t]; }
      else {
        const int fy = floordiv2_2(cls[i].dy[t]), fx = floordiv2_2(cls[i].dx[t]);
        p.oy[i][t] = (short)fy; p.ox[i][t] = (short)fx; p.py[i][t] = (short)(cls[i].dy[t] - 2 * fy); p.px[i][t] = (short)(cls[i].dx[t] - 2 * fx);
      }
    }
    c.flops += 2.0 * g.N * g.Hg * g.Wg * cls[i].ntaps * g.Ci * g.Co;
  }
  c.smem = 1024 + (size_t)p.nstage * p.stage_bytes + (2 * p.nstage + 4) * sizeof(uint64_t) + 16;
  return c.smem <= 227 * 1024;
}

bool tc2_tapconv_supported(const TapGeom* cls, int ncls, const float* const* bt) {
  Tc2Cfg c;
  return tc2_cfg(cls, ncls, bt, c);
}

bool k_tapconv_tc2(St st, const TapGeom* cls, int ncls, const float* const* bt, const float* in, float* out, int act, float negval,
                   std::string* err) {
  Tc2Cfg c;
  if (!tc2_cfg(cls, ncls, bt, c)) { if (err) *err = "geometry not taken by the wide-tile tcgen05 kernel"; return false; }
  Tc2Params& p = c.p;
  p.act = act; p.neg = negval;
  const TapGeom& g = cls[0];
  CUtensorMap mapA;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  if (g.si == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = tc_encode_fn()(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Ci, (cuuint64_t)g.Wi / 2, 2, (cuuint64_t)g.Hi / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)2 * g.Wi * g.Ci * 4,
                             (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = tc_encode_fn()(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); return false; }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem) failed";
      return false;
    }
    configured = true;
  }
  if (c.pair) {
    TcMapsW mw;
    memset(&mw, 0, sizeof(mw));
    for (int i = 0; i < ncls; ++i) {
      const cuuint64_t rows = (cuuint64_t)((g.Co + 127) / 128) * cls[i].ntaps * p.kchunks * 128;
      cuuint64_t dims[2] = {32, rows};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {32, 128};
      r = tc_encode_fn()(&mw.w[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bt[i], dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(W images) failed: " + std::to_string((int)r); return false; }
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)c.grid); lc.blockDim = dim3(TC2_THREADS); lc.dynamicSmemBytes = c.smem; lc.stream = st.s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (cudaLaunchKernelEx(&lc, tapconv_tc2_pair_kernel, mapA, mw, p, out) != cudaSuccess) {
      if (err) *err = std::string("pair-kernel launch failed: ") + cudaGetErrorString(cudaGetLastError());
      return false;
    }
    DSR_LAUNCHED(st, "tapconv_tc2_pair", c.flops, WORK_FLOPS);
    return true;
  }
  tapconv_tc2_kernel<<<c.grid, TC2_THREADS, c.smem, st.s>>>(mapA, p, out);
  DSR_LAUNCHED(st, "tapconv_tc2", c.flops, WORK_FLOPS);
  return true;
}

// ==========================================================================================
// wgrad on CTA pairs (cta_group::2):  acc[t][cp][cq] = sum_pix P[pix][cp] * Q[shift_t(pix)][cq]
//
// The per-tap formulation of kernels_tc.cu:wgrad_tc_kernel (both operands MN-major as the TMA boxes of the NHWC tensors land,
// one TMEM accumulator per tap of the CTA's tap group, split-K over pixel tiles, partials -> scratch -> k_wgrad_reduce) with
// M = 256: the pair owns 256 consecutive cp channels (each CTA stages the P tile of its 128) and every shifted Q tile is
// staged half by each CTA (n_mma / 2 channels).  Per tap and 64-pixel tile a CTA moves 32 KB / TG + n_mma * 128 B instead of
// 32 KB / TG + n_mma * 256 B, and an N = 256 MMA runs at the full tensor rate (an M = 128, N = 128 one is bound by its A read).
// ==========================================================================================
struct WgpParams {
  int Cp, Cq, ntaps, s;
  int TW, TH, TB, tiles_x, tiles_y, ntiles, tiles_per_split;
  int n_mma, atoms_q_half, TG, tmem_cols, tap_groups, q_tiles, mpairs, S;
  int q_stage_bytes, nq_stage;
  long long split_stride;
  short oy[DSR_MAX_TAPS], ox[DSR_MAX_TAPS], py[DSR_MAX_TAPS], px[DSR_MAX_TAPS];
};

__device__ __forceinline__ uint64_t wgp_desc_mn(uint32_t saddr, uint32_t lbo) {      // MN-major SWIZZLE_128B_BASE32B (see kernels_tc.cu)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

#define WGP_KPIX 64
#define WGP_P_STAGE (4 * WGP_KPIX * 32 * 4)      // 128 channels x 64 pixels x 4 B = 32 KB
__global__ void __launch_bounds__(TC2_THREADS, 1) wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap mapP,
                                                                       const __grid_constant__ CUtensorMap mapQ, const WgpParams p,
                                                                       float* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sP = smem;                                        // 2 stages
  uint8_t* sQ = smem + 2 * WGP_P_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + (size_t)p.nq_stage * p.q_stage_bytes);
  uint64_t* p_full = bars;           // [2]
  uint64_t* p_empty = bars + 2;      // [2]
  uint64_t* q_full = bars + 4;       // [8]
  uint64_t* q_empty = bars + 12;     // [8]
  uint64_t* tmem_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  // item = (split, tap group, cp pair, cq tile)
  int item = blockIdx.x >> 1;
  const int qtile = item % p.q_tiles; item /= p.q_tiles;
  const int mpair = item % p.mpairs; item /= p.mpairs;
  const int tgi = item % p.tap_groups; item /= p.tap_groups;
  const int split = item;
  const int t0 = tgi * p.TG;
  const int tg_n = min(p.TG, p.ntaps - t0);
  const int m0 = mpair * 256 + (int)crank * 128, q0 = qtile * p.n_mma + (int)crank * (p.n_mma / 2);
  const int tile_beg = split * p.tiles_per_split;
  const int tile_end = min(p.ntiles, tile_beg + p.tiles_per_split);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&p_full[i]), 1); mbar_init(smem_u32(&p_empty[i]), 1); }
    for (int i = 0; i < p.nq_stage; ++i) { mbar_init(smem_u32(&q_full[i]), 1); mbar_init(smem_u32(&q_empty[i]), 1); }
    mbar_init(smem_u32(tmem_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t atom = (uint32_t)(WGP_KPIX * 32 * 4);        // 32 channels x 64 pixels

  if (warp == 0) {
    if (elect_one()) {
      int ps = 0, qs = 0;
      uint32_t pph = 0, qph = 0;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int tt = tile;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y; tt /= p.tiles_y;
        const int b0 = tt * p.TB, gy0 = ty * p.TH, gx0 = tx * p.TW;
        mbar_wait(smem_u32(&p_empty[ps]), pph ^ 1u);
        const uint32_t pf_local = smem_u32(&p_full[ps]);
        const uint32_t pf = pf_local & 0xFEFFFFFFu;
        if (crank == 0) mbar_expect_tx(pf_local, 2u * 4u * atom);
        for (int a = 0; a < 4; ++a)
          tma_load_4d_2sm(smem_u32(sP + (size_t)ps * WGP_P_STAGE) + a * atom, &mapP, pf, m0 + a * 32, gx0, gy0, b0);
        if (++ps == 2) { ps = 0; pph ^= 1u; }
        for (int tg = 0; tg < tg_n; ++tg) {
          const int t = t0 + tg;
          mbar_wait(smem_u32(&q_empty[qs]), qph ^ 1u);
          const uint32_t qf_local = smem_u32(&q_full[qs]);
          const uint32_t qf = qf_local & 0xFEFFFFFFu;
          if (crank == 0) mbar_expect_tx(qf_local, 2u * (uint32_t)p.atoms_q_half * atom);
          const uint32_t dq = smem_u32(sQ + (size_t)qs * p.q_stage_bytes);
          for (int a = 0; a < p.atoms_q_half; ++a) {
            if (p.s == 1)
              tma_load_4d_2sm(dq + a * atom, &mapQ, qf, q0 + a * 32, gx0 + p.ox[t], gy0 + p.oy[t], b0);
            else
              tma_load_5d_2sm(dq + a * atom, &mapQ, qf, p.px[t] * p.Cq + q0 + a * 32, gx0 + p.ox[t], p.py[t], gy0 + p.oy[t], b0);
          }
          if (++qs == p.nq_stage) { qs = 0; qph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // D = f32, A = B = tf32, both MN-major (bits 15, 16), N = n_mma, M = 256
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_mma >> 3) << 17) | ((256u >> 4) << 24);
    int ps = 0, qs = 0;
    uint32_t pph = 0, qph = 0;
    const uint32_t kstep = 8u * 32u * 4u;                     // 8 pixels of a 32-channel atom
    for (int tile = tile_beg; tile < tile_end; ++tile) {
      mbar_wait(smem_u32(&p_full[ps]), pph);
      const uint32_t pa = smem_u32(sP + (size_t)ps * WGP_P_STAGE);
      for (int tg = 0; tg < tg_n; ++tg) {
        mbar_wait(smem_u32(&q_full[qs]), qph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t qa = smem_u32(sQ + (size_t)qs * p.q_stage_bytes);
#pragma unroll
          for (int k = 0; k < WGP_KPIX / 8; ++k)
            umma_tf32_2sm(tmem_base + (uint32_t)(tg * p.n_mma), wgp_desc_mn(pa + k * kstep, atom), wgp_desc_mn(qa + k * kstep, atom), idesc,
                          (tile > tile_beg || k > 0) ? 1u : 0u);
          umma_commit_2sm(smem_u32(&q_empty[qs]));
          if (tg == tg_n - 1) {
            umma_commit_2sm(smem_u32(&p_empty[ps]));
            if (tile == tile_end - 1) umma_commit_2sm(smem_u32(tmem_full));
          }
        }
        __syncwarp();
        if (++qs == p.nq_stage) { qs = 0; qph ^= 1u; }
      }
      if (++ps == 2) { ps = 0; pph ^= 1u; }
    }
  } else if (warp >= 2 && tile_end > tile_beg) {
    const int q = warp & 3;
    const int cp = m0 + q * 32 + lane;
    const int cq0 = qtile * p.n_mma;                          // the accumulator columns are the whole n_mma slice (both CTAs' halves)
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
          const int cq = cq0 + c0;
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
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

static int wgp_max_clusters() {
  static int v = -1;
  if (v >= 0) return v;
  v = 0;
  if (cudaFuncSetAttribute(wgrad_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return v; }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(2 * NSM); lc.blockDim = dim3(TC2_THREADS); lc.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, wgrad_tc_pair_kernel, &lc) != cudaSuccess) { cudaGetLastError(); return v; }
  v = std::min(n, NSM / 2);
  return v;
}

struct WgpCfg { WgpParams p; size_t smem; int grid; };

static bool wgp_cfg(const WgradGeom& g, WgpCfg& c) {
  if (!tc_encode_fn()) return false;
  const char* e = getenv("DCGANSR_WGRAD_PAIR");
  const int mode = e ? atoi(e) : 1;        // 0 off, 1 where it pays, 2 wherever it runs
  if (mode == 0) return false;
  WgpParams& p = c.p;
  memset(&p, 0, sizeof(p));
  if (g.Cp % 32 || g.Cq % 32 || g.Cp < 256 || g.Cq < 64 || g.ntaps < 1 || g.ntaps > DSR_MAX_TAPS) return false;
  if (g.s != 1 && g.s != 2) return false;
  if (g.s == 2 && (g.Hq % 2 || g.Wq % 2)) return false;
  const int maxcl = wgp_max_clusters();
  if (maxcl < 32) return false;
  p.Cp = g.Cp; p.Cq = g.Cq; p.ntaps = g.ntaps; p.s = g.s;
  p.TW = std::min(pow2_ge2(g.Wp), WGP_KPIX);
  p.TH = std::min(pow2_ge2(g.Hp), WGP_KPIX / p.TW);
  p.TB = WGP_KPIX / (p.TW * p.TH);
  p.tiles_x = (g.Wp + p.TW - 1) / p.TW;
  p.tiles_y = (g.Hp + p.TH - 1) / p.TH;
  p.ntiles = ((g.N + p.TB - 1) / p.TB) * p.tiles_y * p.tiles_x;
  // cq slice per item: 256 columns when the layer has them (an N = 256 pair MMA runs at the full tensor rate), else 128 / 64
  p.n_mma = g.Cq >= 256 ? 256 : (g.Cq >= 128 ? 128 : 64);
  p.q_tiles = (g.Cq + p.n_mma - 1) / p.n_mma;
  p.atoms_q_half = p.n_mma / 64;
  p.TG = std::max(1, std::min(g.ntaps, 512 / p.n_mma));
  p.tap_groups = (g.ntaps + p.TG - 1) / p.TG;
  p.mpairs = (g.Cp + 255) / 256;
  p.tmem_cols = std::max(32, pow2_ge2(p.TG * p.n_mma));
  p.q_stage_bytes = p.atoms_q_half * WGP_KPIX * 32 * 4;
  p.nq_stage = std::max(2, std::min(8, (200 * 1024 - 2 * WGP_P_STAGE) / p.q_stage_bytes));
  // pixel splits: every item accumulates over its own range of pixel tiles.  The items run in waves of `maxcl` pairs, so S is
  // chosen to fill whole waves (19 splits of 4 tap groups = 76 items on 74 pairs would run two half-empty waves: measured 0.91x)
  const int other = p.tap_groups * p.mpairs * p.q_tiles;
  int best_s = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= p.ntiles && S * other <= 4 * maxcl; ++S) {
    const int tps = (p.ntiles + S - 1) / S;
    const int s_eff = (p.ntiles + tps - 1) / tps;              // splits that actually hold tiles
    const long long items = (long long)s_eff * other;
    const long long waves = (items + maxcl - 1) / maxcl;
    const double eff = (double)items / (double)(waves * maxcl);
    if (eff > best_eff + 0.03) { best_eff = eff; best_s = S; }   // a later (larger) S must be clearly better: more partials to reduce
  }
  p.tiles_per_split = (p.ntiles + best_s - 1) / best_s;
  p.S = (p.ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.split_stride = (long long)g.Cp * g.ntaps * g.Cq;
  for (int t = 0; t < g.ntaps; ++t) {
    if (g.s == 1) { p.oy[t] = (short)g.dy[t]; p.ox[t] = (short)g.dx[t]; }
    else {
      const int fy = floordiv2_2(g.dy[t]), fx = floordiv2_2(g.dx[t]);
      p.oy[t] = (short)fy; p.ox[t] = (short)fx; p.py[t] = (short)(g.dy[t] - 2 * fy); p.px[t] = (short)(g.dx[t] - 2 * fx);
    }
  }
  c.grid = 2 * p.S * other;
  c.smem = 1024 + 2 * WGP_P_STAGE + (size_t)p.nq_stage * p.q_stage_bytes + 24 * sizeof(uint64_t);
  // measured (scripts/exp/wgrad_pair_sweep.py): 1.27 - 1.63x over wgrad_tc from 256 pixel tiles up (C1b FC 256->128 566 -> 352 us,
  // C5 FC 512->256 1811 -> 1111 us = 990 TFLOP/s), 0.91 - 0.95x on the small discriminator layers at batch 128 (128 tiles)
  if (mode == 1 && (p.ntiles < 256 || p.tiles_per_split < 4)) return false;
  return c.smem <= 227 * 1024;
}

bool wgrad_tc_pair_supported(const WgradGeom& g) { WgpCfg c; return wgp_cfg(g, c); }
size_t wgrad_tc_pair_scratch_bytes(const WgradGeom& g) {
  WgpCfg c;
  if (!wgp_cfg(g, c)) return 0;
  return (size_t)c.p.S * g.Cp * g.ntaps * g.Cq * sizeof(float);
}

bool k_wgrad_tc_pair(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                     std::string* err) {
  WgpCfg c;
  if (!wgp_cfg(g, c)) { if (err) *err = "wgrad geometry not taken by the pair kernel"; return false; }
  const WgpParams& p = c.p;
  if ((size_t)p.S * g.Cp * g.ntaps * g.Cq * sizeof(float) > scratch_bytes) { if (err) *err = "wgrad scratch too small"; return false; }
  CUtensorMap mapP, mapQ;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.Wp, (cuuint64_t)g.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 4, (cuuint64_t)g.Wp * g.Cp * 4, (cuuint64_t)g.Hp * g.Wp * g.Cp * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = tc_encode_fn()(&mapP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)P, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(P) failed: " + std::to_string((int)r); return false; }
  if (g.s == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cq, (cuuint64_t)g.Wq, (cuuint64_t)g.Hq, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = tc_encode_fn()(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Cq, (cuuint64_t)g.Wq / 2, 2, (cuuint64_t)g.Hq / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)2 * g.Wq * g.Cq * 4,
                             (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
    r = tc_encode_fn()(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(Q) failed: " + std::to_string((int)r); return false; }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)c.grid); lc.blockDim = dim3(TC2_THREADS); lc.dynamicSmemBytes = c.smem; lc.stream = st.s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  if (cudaLaunchKernelEx(&lc, wgrad_tc_pair_kernel, mapP, mapQ, p, scratch) != cudaSuccess) {
    if (err) *err = std::string("wgrad pair launch failed: ") + cudaGetErrorString(cudaGetLastError());
    return false;
  }
  DSR_LAUNCHED(st, "wgrad_tc_pair", 2.0 * g.N * g.Hp * g.Wp * g.Cp * g.Cq * g.ntaps, WORK_FLOPS);
  k_wgrad_reduce(st, scratch, p.S, g.Cp, g.Cq, g.ntaps, grad_master);
  return true;
}
