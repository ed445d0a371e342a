// kernels_tc3.cu -- halo-tile A, streamed weights, CTA pairs: the tensor-bound convolutions with <= 128 couts per class
// (FAST_TF32): the generator's inner layers at ngf >= 64 (train-gray-patch.lua:60-70: FC 256->128 forward, C 128->256
// updateGradInput, FC 128->64 / C 64->128), the discriminator's updateGradInput of train.lua:124-131 and the patch
// discriminator's 3 x 3 convolutions (train-gray-patch.lua:96-104).
//
// With 128 couts the per-tap kernels (kernels_tc.cu, kernels_tc2.cu) cannot share an A tile between cout tiles, so their
// L2 -> SM traffic is dominated by re-fetching the input pixels once per tap: 88 B/clk/SM against the ~50 B/clk/SM the L2
// delivers (measured: lts__t_bytes = operand bytes, 13.5 TB/s), i.e. 45 % tensor-pipe activity.  Here
//
//   * the input pixels of an 8 x 16 tile are loaded ONCE per 32-channel chunk as a halo tile (one TMA box per stride-parity
//     plane; zero fill = padding) and every tap is a shifted START ADDRESS of that tile (the UMMA swizzle is a function of the
//     absolute shared-memory address, scripts/exp/exp_desc.cu): 20 KB instead of 4 x 16 KB for a 2 x 2-tap class;
//   * the weights are streamed tap by tap through a second ring (they do not fit shared memory: 2 MB for FC 256->128);
//   * a CTA pair (cta_group::2) runs one 256 pixel x BN cout MMA: each CTA stages its own halo tile and HALF of every weight
//     tile (BN/2 rows of the pre-tiled image) -- 52 KB per 1088 tensor cycles = 48 B/clk/SM for a 2 x 2-tap class at BN = 128;
//   * TMEM: two BN-column accumulator buffers per CTA, the epilogue of item i overlaps the MMAs of item i + 1.
//
// Barrier protocol as in kernels_tc2.cu's pair kernel (loads complete on the leader's barriers, multicast commits, remote
// accumulator release).  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM owner, warps 2..9 epilogue.
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define NSM 148
#define TC3_MAXCLS 4
#define TC3_MAXPL 4
#define TC3_THREADS 320      // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter: the outputs of these
                             // layers are large against their flops, four warps drain an accumulator more slowly than the MMAs fill it)
#define TC3_TW 8
#define TC3_TH 16

struct Tc3MapsW { CUtensorMap w[TC3_MAXCLS]; };

struct Tc3Params {
  int N, Hg, Wg, Ho, Wo, Co;
  int so, si, Ci;
  int tiles_x, tiles_y, ntiles_m;
  int kchunks;
  int BN, bn_half, img_rows, img_per_n;       // couts per item; weight rows per CTA; rows of a pre-tiled image; images per item
  int ntiles_n, ngroups_m, nwork;
  int PW, PH, plane_bytes, plane_tx, b_stage_bytes, na_stage, nb_stage;      // na_stage: PLANE buffers in the A ring
  int acc_cols, tmem_cols;
  int ncls, oy0[TC3_MAXCLS], ox0[TC3_MAXCLS], ntaps[TC3_MAXCLS], nplanes[TC3_MAXCLS];
  short pl_x[TC3_MAXCLS][TC3_MAXPL], pl_y[TC3_MAXCLS][TC3_MAXPL], pl_py[TC3_MAXCLS][TC3_MAXPL], pl_c[TC3_MAXCLS][TC3_MAXPL];
  // taps in issue order: sorted by plane (plane pl owns taps [pl_tbeg[pl], pl_tbeg[pl + 1])); per tap the byte offset of its window
  // inside the plane buffer and its index in the weight images (original tap number)
  short pl_tbeg[TC3_MAXCLS][TC3_MAXPL + 1], tap_w[TC3_MAXCLS][DSR_MAX_TAPS];
  int tap_aoff[TC3_MAXCLS][DSR_MAX_TAPS];
  int act;
  float neg;
};

__device__ __forceinline__ void tc3_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// K-major descriptor with explicit SBO (8-pixel groups one halo row apart), SWIZZLE_128B
__device__ __forceinline__ uint64_t tc3_desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Tc3Item { int cls, ntile, mg; };
__device__ __forceinline__ Tc3Item tc3_item(const Tc3Params& p, int w) {
  Tc3Item it;
  it.mg = w % p.ngroups_m; w /= p.ngroups_m;
  it.ntile = w % p.ntiles_n;
  it.cls = w / p.ntiles_n;
  return it;
}

template <int ACT>
__device__ __forceinline__ void tc3_epilogue(const Tc3Params& p, float* __restrict__ out, uint32_t tmem_base, uint64_t* acc_full,
                                             uint64_t* acc_empty, int warp, int lane, uint32_t crank, int pair, int npairs) {
  const int q = warp & 3;
  const int half = (warp - 2) >> 2;            // which half of the item's couts this warp stores
  const int r = q * 32 + lane;                 // tile row = pixel (x fastest, 8 wide)
  const int w = r % TC3_TW, h = r / TC3_TW;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  int it = 0;
  for (int wi = pair; wi < p.nwork; wi += npairs, ++it) {
    const Tc3Item item = tc3_item(p, wi);
    const int buf = it & 1;
    const uint32_t aph = (uint32_t)(it >> 1) & 1u;
    const int n0 = item.ntile * p.BN;
    mbar_wait(smem_u32(&acc_full[buf]), aph);
    tc_fence_after();
    int tile = item.mg * 2 + (int)crank;
    if (tile < p.ntiles_m) {
      const int tx = tile % p.tiles_x; tile /= p.tiles_x;
      const int ty = tile % p.tiles_y; tile /= p.tiles_y;
      const int n = tile, gy = ty * TC3_TH + h, gx = tx * TC3_TW + w;
      const bool valid = gy < p.Hg && gx < p.Wg;
      float* orow = out + ((int64_t)(n * p.Ho + gy * p.so + p.oy0[item.cls]) * p.Wo + gx * p.so + p.ox0[item.cls]) * p.Co + n0;
      const uint32_t cbase = lane_base + (uint32_t)(buf * p.acc_cols);
      const int cbeg = half * (p.BN >> 1), cend = cbeg + (p.BN >> 1);
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(cbase + (uint32_t)c0, v);
        tmem_ld_wait();
        if (valid) store_row<ACT, 32>(orow + c0, v, n0 + c0, p.Co, p.neg);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (crank == 0) tc3_arrive(smem_u32(&acc_empty[buf]));
      else mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0);
    }
  }
}

__global__ void __launch_bounds__(TC3_THREADS, 1) tapconv_tc3_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                     const __grid_constant__ Tc3MapsW mapsW, const Tc3Params p,
                                                                     float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.na_stage * p.plane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.nb_stage * p.b_stage_bytes);
  uint64_t* a_full = bars;                     // [8]
  uint64_t* a_empty = bars + 8;                // [8]
  uint64_t* b_full = bars + 16;                // [32]
  uint64_t* b_empty = bars + 48;               // [32]
  uint64_t* acc_full = bars + 80;              // [2]
  uint64_t* acc_empty = bars + 82;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 84);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    for (int s = 0; s < p.na_stage; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < p.nb_stage; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&acc_full[s]), 1); mbar_init(smem_u32(&acc_empty[s]), 16); }
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

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion on the leader's barriers) =====================
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t b_tx_pair = 2u * (uint32_t)p.bn_half * 128u;
      const int half_row = ((int)crank * p.bn_half) % p.img_rows, half_img = ((int)crank * p.bn_half) / p.img_rows;
      for (int wi = pair; wi < p.nwork; wi += npairs) {
        const Tc3Item item = tc3_item(p, wi);
        const int cls = item.cls;
        const int npl = p.nplanes[cls];
        const int nk = p.ntaps[cls] * p.kchunks;
        int tile = item.mg * 2 + (int)crank;                      // past the end: image index out of range -> zero fill
        const int tx = tile % p.tiles_x; tile /= p.tiles_x;
        const int ty = tile % p.tiles_y; tile /= p.tiles_y;
        const int n = tile, gy0 = ty * TC3_TH, gx0 = tx * TC3_TW;
        const int wrow0 = ((item.ntile * p.img_per_n + half_img) * nk) * p.img_rows + half_row;
        for (int c = 0; c < p.kchunks; ++c) {
          for (int pl = 0; pl < npl; ++pl) {
            mbar_wait(smem_u32(&a_empty[as]), aph ^ 1u);
            const uint32_t af_local = smem_u32(&a_full[as]);
            const uint32_t af = af_local & 0xFEFFFFFFu;
            if (crank == 0) mbar_expect_tx(af_local, 2u * (uint32_t)p.plane_tx);
            const uint32_t sa = smem_u32(sA + (size_t)as * p.plane_bytes);
            if (p.si == 1)
              tma_load_4d_2sm(sa, &mapA, af, c * 32, gx0 + p.pl_x[cls][pl], gy0 + p.pl_y[cls][pl], n);
            else
              tma_load_5d_2sm(sa, &mapA, af, p.pl_c[cls][pl] + c * 32, gx0 + p.pl_x[cls][pl], p.pl_py[cls][pl], gy0 + p.pl_y[cls][pl], n);
            if (++as == p.na_stage) { as = 0; aph ^= 1u; }
            for (int t = p.pl_tbeg[cls][pl]; t < p.pl_tbeg[cls][pl + 1]; ++t) {
              mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1u);
              const uint32_t bf_local = smem_u32(&b_full[bs]);
              if (crank == 0) mbar_expect_tx(bf_local, b_tx_pair);
              tma_load_2d_2sm(smem_u32(sB + (size_t)bs * p.b_stage_bytes), &mapsW.w[cls], bf_local & 0xFEFFFFFFu, 0,
                              wrow0 + (p.tap_w[cls][t] * p.kchunks + c) * p.img_rows);
              if (++bs == p.nb_stage) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ===================== MMA issuer (leader only): D[256 pixels][BN] += A(tap window)[256][8] * W[BN][8] =====================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t a_sbo = (uint32_t)p.PW * 128u;
    int as = 0, bs = 0, it = 0;
    uint32_t aph = 0, bph = 0;
    for (int wi = pair; wi < p.nwork; wi += npairs, ++it) {
      const Tc3Item item = tc3_item(p, wi);
      const int cls = item.cls;
      const int npl = p.nplanes[cls];
      const int buf = it & 1;
      const uint32_t cph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), cph ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem_base + (uint32_t)(buf * p.acc_cols);
      for (int c = 0; c < p.kchunks; ++c) {
        for (int pl = 0; pl < npl; ++pl) {
          mbar_wait(smem_u32(&a_full[as]), aph);
          tc_fence_after();
          const uint32_t sa = smem_u32(sA + (size_t)as * p.plane_bytes);
          const int tend = p.pl_tbeg[cls][pl + 1];
          for (int t = p.pl_tbeg[cls][pl]; t < tend; ++t) {
            mbar_wait(smem_u32(&b_full[bs]), bph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad = tc3_desc(sa + (uint32_t)p.tap_aoff[cls][t], a_sbo);
              const uint64_t bd = make_kmajor_desc(smem_u32(sB + (size_t)bs * p.b_stage_bytes), 32);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_2sm(dbase, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (c > 0 || t > 0 || k > 0) ? 1u : 0u);
              umma_commit_2sm(smem_u32(&b_empty[bs]));
              if (t == tend - 1) {
                umma_commit_2sm(smem_u32(&a_empty[as]));
                if (c == p.kchunks - 1 && pl == npl - 1) umma_commit_2sm(smem_u32(&acc_full[buf]));
              }
            }
            __syncwarp();
            if (++bs == p.nb_stage) { bs = 0; bph ^= 1u; }
          }
          if (++as == p.na_stage) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp >= 2) {
    switch (p.act) {
      case ACT_RELU: tc3_epilogue<ACT_RELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_LRELU: tc3_epilogue<ACT_LRELU>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_TANH: tc3_epilogue<ACT_TANH>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      case ACT_SIGMOID: tc3_epilogue<ACT_SIGMOID>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
      default: tc3_epilogue<ACT_NONE>(p, out, tmem_base, acc_full, acc_empty, warp, lane, crank, pair, npairs); break;
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pow2_ge3(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int fdiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

static int tc3_max_clusters() {
  static int v = -1;
  if (v >= 0) return v;
  v = 0;
  if (cudaFuncSetAttribute(tapconv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return v; }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(2 * NSM); lc.blockDim = dim3(TC3_THREADS); lc.dynamicSmemBytes = 210 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tapconv_tc3_kernel, &lc) != cudaSuccess) { cudaGetLastError(); return v; }
  v = std::min(n, NSM / 2);
  return v;
}

struct Tc3Cfg { Tc3Params p; size_t smem; int grid; double flops; };

static bool tc3_cfg(const TapGeom* cls, int ncls, const float* const* bt, Tc3Cfg& c) {
  if (!tc_encode_fn() || !tc_tapconv_multi_ok(cls, ncls) || !bt || ncls > TC3_MAXCLS) return false;
  const char* me = getenv("DCGANSR_TC3");
  const int mode = me ? atoi(me) : 1;      // 0: off, 1: where it pays, 2: wherever it runs
  if (mode == 0) return false;
  const TapGeom& g = cls[0];
  if (g.Ci % 32 != 0) return false;
  for (int i = 0; i < ncls; ++i)
    if (!bt[i] || cls[i].ntaps < 1 || cls[i].ntaps > DSR_MAX_TAPS) return false;
  if (g.Wg < TC3_TW || g.Hg < 8) return false;
  Tc3Params& p = c.p;
  memset(&p, 0, sizeof(p));
  p.N = g.N; p.Hg = g.Hg; p.Wg = g.Wg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.so = g.so; p.si = g.si; p.Ci = g.Ci; p.ncls = ncls;
  p.kchunks = g.Ci / 32;
  p.tiles_x = (g.Wg + TC3_TW - 1) / TC3_TW;
  p.tiles_y = (g.Hg + TC3_TH - 1) / TC3_TH;
  p.ntiles_m = g.N * p.tiles_y * p.tiles_x;
  // weight images: img_rows couts each (kernels_tc.cu:tc_bt_rows); an item takes BN couts = 1 or 2 images, a CTA half of them
  p.img_rows = tc_bt_rows(g.Co);
  if (p.img_rows != 128 && p.img_rows != 64) return false;
  const int n_img = (g.Co + p.img_rows - 1) / p.img_rows;
  if (p.img_rows == 128 && n_img % 2 == 0) { p.BN = 256; p.img_per_n = 2; }
  else { p.BN = p.img_rows; p.img_per_n = 1; }
  p.bn_half = p.BN / 2;
  p.ntiles_n = n_img / p.img_per_n;
  p.ngroups_m = (p.ntiles_m + 1) / 2;
  p.nwork = ncls * p.ntiles_n * p.ngroups_m;
  // planes (stride-parity planes of the input that the taps of a class read) and the tap windows inside them
  int ext_x = 0, ext_y = 0, max_pl = 0;
  struct Pl { int py, px, ymin, ymax, xmin, xmax; };
  Pl pls[TC3_MAXCLS][TC3_MAXPL];
  int tap_pl[TC3_MAXCLS][DSR_MAX_TAPS], tap_oy[TC3_MAXCLS][DSR_MAX_TAPS], tap_ox[TC3_MAXCLS][DSR_MAX_TAPS];
  for (int i = 0; i < ncls; ++i) {
    int npl = 0;
    for (int t = 0; t < cls[i].ntaps; ++t) {
      int oy = cls[i].dy[t], ox = cls[i].dx[t], py = 0, px = 0;
      if (g.si == 2) { oy = fdiv2(cls[i].dy[t]); ox = fdiv2(cls[i].dx[t]); py = cls[i].dy[t] - 2 * oy; px = cls[i].dx[t] - 2 * ox; }
      int f = -1;
      for (int k = 0; k < npl; ++k) if (pls[i][k].py == py && pls[i][k].px == px) f = k;
      if (f < 0) { if (npl == TC3_MAXPL) return false; f = npl++; pls[i][f] = Pl{py, px, oy, oy, ox, ox}; }
      pls[i][f].ymin = std::min(pls[i][f].ymin, oy); pls[i][f].ymax = std::max(pls[i][f].ymax, oy);
      pls[i][f].xmin = std::min(pls[i][f].xmin, ox); pls[i][f].xmax = std::max(pls[i][f].xmax, ox);
      tap_pl[i][t] = f; tap_oy[i][t] = oy; tap_ox[i][t] = ox;
    }
    p.nplanes[i] = npl;
    max_pl = std::max(max_pl, npl);
    for (int k = 0; k < npl; ++k) { ext_x = std::max(ext_x, pls[i][k].xmax - pls[i][k].xmin); ext_y = std::max(ext_y, pls[i][k].ymax - pls[i][k].ymin); }
  }
  p.PW = TC3_TW + ext_x; p.PH = TC3_TH + ext_y;
  if (p.PW > 256 || p.PH > 256) return false;
  p.plane_tx = p.PW * p.PH * 128;
  p.plane_bytes = (p.plane_tx + 1023) / 1024 * 1024;
  p.b_stage_bytes = std::max(1024, p.bn_half * 128);
  for (int i = 0; i < ncls; ++i) {
    p.oy0[i] = cls[i].oy0; p.ox0[i] = cls[i].ox0; p.ntaps[i] = cls[i].ntaps;
    int k2 = 0;
    for (int k = 0; k < p.nplanes[i]; ++k) {
      p.pl_x[i][k] = (short)pls[i][k].xmin; p.pl_y[i][k] = (short)pls[i][k].ymin; p.pl_py[i][k] = (short)pls[i][k].py;
      p.pl_c[i][k] = (short)(pls[i][k].px * g.Ci);
      p.pl_tbeg[i][k] = (short)k2;
      for (int t = 0; t < cls[i].ntaps; ++t)
        if (tap_pl[i][t] == k) {
          p.tap_w[i][k2] = (short)t;
          p.tap_aoff[i][k2] = ((tap_oy[i][t] - pls[i][k].ymin) * p.PW + (tap_ox[i][t] - pls[i][k].xmin)) * 128;
          ++k2;
        }
    }
    p.pl_tbeg[i][p.nplanes[i]] = (short)k2;
  }
  // shared memory: both rings must cover the L2 latency under load at the tensor rate (the per-tap pair kernel's 6 x 544 cycles
  // do; 3 x 544 measured here: 60 % instead of 89 % tensor-pipe activity).  Weight ring first: ~3600 tensor cycles of taps
  // (4 MMAs of max(64, BN / 2) cycles each), the rest holds plane buffers
  const int budget = 212 * 1024;
  const int tap_clk = 4 * std::max(64, p.BN / 2);
  p.nb_stage = std::max(4, std::min(32, (3600 + tap_clk - 1) / tap_clk));
  while (p.nb_stage > 3 && budget - p.nb_stage * p.b_stage_bytes < 2 * p.plane_bytes) --p.nb_stage;
  p.na_stage = std::min(8, (budget - p.nb_stage * p.b_stage_bytes) / p.plane_bytes);
  if (p.na_stage < 2) return false;
  p.nb_stage = std::min(32, (budget - p.na_stage * p.plane_bytes) / p.b_stage_bytes);
  if (const char* e = getenv("DCGANSR_TC3_NA")) p.na_stage = std::max(2, std::min(p.na_stage, atoi(e)));      // ring-depth experiments
  if (const char* e = getenv("DCGANSR_TC3_NB")) p.nb_stage = std::max(2, std::min(p.nb_stage, atoi(e)));
  p.acc_cols = p.BN;
  p.tmem_cols = std::max(32, pow2_ge3(2 * p.acc_cols));
  const int maxcl = tc3_max_clusters();
  if (maxcl < 32) return false;
  c.grid = 2 * std::min(p.nwork, maxcl);
  if (mode == 1) {
    // Measured per layer and batch against the per-tap kernels (scripts/exp/tc3_sweep.py): 1.15 - 1.4x for <= 128-cout items
    // (45 -> 58-60 % tensor-pipe activity: with N = 128 the MMA is bound by its A-operand read from shared memory, not by L2
    // any more) and 1.1x for 256-cout items on full tiles (C1b C 128->256 forward 296 -> 265 us = 1.04 PFLOP/s TF32, 95 % of the
    // pipe peak); half-empty tiles (class grids under 12 rows, 16 for the 256-cout items the per-tap pair kernel already runs at
    // 89 %) waste the MMAs they save; small grids stay on the one-tile kernel with its split-K
    if (p.nwork < 32 || g.Hg < (p.BN > 128 ? 16 : 12)) return false;
  }
  c.flops = 0;
  for (int i = 0; i < ncls; ++i) c.flops += 2.0 * g.N * g.Hg * g.Wg * cls[i].ntaps * g.Ci * g.Co;
  c.smem = 1024 + (size_t)p.na_stage * p.plane_bytes + (size_t)p.nb_stage * p.b_stage_bytes + 88 * sizeof(uint64_t);
  return c.smem <= 227 * 1024;
}

bool tc3_tapconv_supported(const TapGeom* cls, int ncls, const float* const* bt) {
  Tc3Cfg c;
  return tc3_cfg(cls, ncls, bt, c);
}

bool k_tapconv_tc3(St st, const TapGeom* cls, int ncls, const float* const* bt, const float* in, float* out, int act, float negval,
                   std::string* err) {
  Tc3Cfg c;
  if (!tc3_cfg(cls, ncls, bt, c)) { if (err) *err = "geometry not taken by the halo-tile pair kernel"; return false; }
  Tc3Params& p = c.p;
  p.act = act; p.neg = negval;
  const TapGeom& g = cls[0];
  CUtensorMap mapA;
  Tc3MapsW mw;
  memset(&mw, 0, sizeof(mw));
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  if (g.si == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1};
    r = tc_encode_fn()(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Ci, (cuuint64_t)g.Wi / 2, 2, (cuuint64_t)g.Hi / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)2 * g.Wi * g.Ci * 4,
                             (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)p.PW, 1, (cuuint32_t)p.PH, 1};
    r = tc_encode_fn()(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(A halo) failed: " + std::to_string((int)r); return false; }
  for (int i = 0; i < ncls; ++i) {
    const cuuint64_t rows = (cuuint64_t)((g.Co + p.img_rows - 1) / p.img_rows) * cls[i].ntaps * p.kchunks * p.img_rows;
    cuuint64_t dims[2] = {32, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)p.bn_half};
    r = tc_encode_fn()(&mw.w[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bt[i], dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(W images) failed: " + std::to_string((int)r); return false; }
  }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)c.grid); lc.blockDim = dim3(TC3_THREADS); lc.dynamicSmemBytes = c.smem; lc.stream = st.s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  if (cudaLaunchKernelEx(&lc, tapconv_tc3_kernel, mapA, mw, p, out) != cudaSuccess) {
    if (err) *err = std::string("halo-tile pair kernel launch failed: ") + cudaGetErrorString(cudaGetLastError());
    return false;
  }
  DSR_LAUNCHED(st, "tapconv_tc3", c.flops, WORK_FLOPS);
  return true;
}
