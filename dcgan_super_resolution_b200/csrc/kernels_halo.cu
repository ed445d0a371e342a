// kernels_halo.cu -- weights-resident, halo-tile tcgen05 implicit-GEMM convolution for the spatially large, thin
// layers of the generator (FAST_TF32): nn.SpatialFullConvolution / nn.SpatialConvolution forward and updateGradInput
// (train.lua:99-111, train-gray.lua:105-116) where the activations (hundreds of MB) dominate and the whole
// weight tensor fits in shared memory.
//
// What bounds these layers is HBM, and what bounded the per-tap kernel (kernels_tc.cu) was the L2 -> SM path: it
// re-fetches the input pixels once per tap (4..16x) and the weights once per 128-pixel tile.  Here
//
//   * a persistent CTA loads its slice of the packed weights ONCE (TMA, K-major, swizzled) and keeps it in smem;
//   * per 16 x 8 tile of the output grid it loads the input pixels ONCE: a halo tile, one TMA box per 32-channel
//     plane (zero fill = padding).  The planes of consecutive tiles go through a RING of plane buffers: the MMAs of a
//     tile are issued plane by plane, so a buffer is free again as soon as the MMAs that read it have completed -- the
//     ring can be shorter than two whole tiles (resident weights of 100..150 KB leave room for 3..6 planes) and the
//     freed shared memory takes the TMA-store staging buffers;
//   * every tap is then just another START ADDRESS of the same smem tile: the UMMA descriptor's swizzle is a
//     function of the absolute shared-memory address, so a window shifted by whole pixel rows (and stepping
//     `pitch` bytes between 8-pixel groups, SBO = halo row pitch) is a valid K-major operand
//     (verified on B200 by scripts/exp/exp_desc.cu);
//   * all stride^2 sub-pixel classes of a full-conv forward / conv dgrad are computed by the same CTA from the
//     same halo tile into separate TMEM column ranges, so the input is read once, not once per class;
//   * accumulators are double buffered in TMEM: the epilogue of tile i (tcgen05.ld -> activation -> NHWC stores)
//     overlaps the TMA + MMA of tile i + 1.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define NSM 148
#define HALO_TH 16
#define HALO_TW 8
#define HALO_MAXCLS 4
#define HALO_MAXTAPS 16
#define HALO_MAXPLANES 8
#define HALO_MAXGRP 4
#define HALO_MAXSLOT 8
#define HALO_MAXRING 16

struct HaloMaps { CUtensorMap b[HALO_MAXCLS]; CUtensorMap bh[HALO_MAXCLS]; };      // bh: boxes of Npad / 2 rows (CTA-pair mode, single-class MMAs)
// CTA-pair mode: the resident weights of CTA `rank` as a list of TMA boxes {class | half-box flag << 8, K coordinate, cout offset,
// destination >> 4}: for an MMA over m classes a CTA holds rows [rank * N/2, (rank + 1) * N/2) of the m * Npad stacked weight rows
#define HALO_MAXWL 160
struct HaloWList { int n; uint4 e[2][HALO_MAXWL]; };
#define HALO_MAXMMA 256
// MMA issue table (kernel parameter = constant bank, so the issuing warp reads it through the uniform datapath):
// one entry per tcgen05.mma of a tile {A offset >> 4, W offset >> 4, TMEM column | accumulate << 31, instruction descriptor}:
// everything the issuing thread would otherwise have to compute per MMA.
struct HaloTab { uint4 e[HALO_MAXMMA]; };

struct HaloParams {
  int N, Hg, Wg, tiles_x, tiles_y, ntiles;
  int Ho, Wo, Co, so, ncls, si, Ci;
  int row_bytes, a_layout;           // A pixel-row bytes (128 / 64) and UMMA layout code (2 = SW128, 4 = SW64)
  int KBw, kchunks, ksteps;          // weight rows: KBw floats; chunks per tap; K = 8 steps per chunk
  int nplanes, plane_bytes, plane_tx, PH, PW, pitch_bytes;
  short pl_c[HALO_MAXPLANES], pl_x[HALO_MAXPLANES], pl_py[HALO_MAXPLANES], pl_y[HALO_MAXPLANES];
  int Npad, wtile_bytes, w_bytes, w_tx;
  int ntaps[HALO_MAXCLS];
  short coy[HALO_MAXCLS], cox[HALO_MAXCLS];
  unsigned short tap_plane[HALO_MAXCLS][HALO_MAXTAPS], tap_wtile[HALO_MAXCLS][HALO_MAXTAPS], tap_wstride[HALO_MAXCLS][HALO_MAXTAPS];
  int tap_aoff[HALO_MAXCLS][HALO_MAXTAPS];
  int nring, acc_cols, nacc, tmem_cols, nmma;   // nring: plane buffers in the ring
  int ngrp, gbeg[HALO_MAXGRP + 1];      // MMA issuer warps and their table ranges
  unsigned short pend[HALO_MAXGRP][HALO_MAXPLANES];   // group g, plane pl: its table entries end at pend[g][pl] (entries sorted by plane)
  int nslots;                            // accumulator slots (Npad TMEM columns each)
  int cls_nsl[HALO_MAXCLS];              // class c = sum of slots cls_sl[c][0 .. cls_nsl[c])
  unsigned char cls_sl[HALO_MAXCLS][4];
  int act;
  float neg;
  int tstore, st_bytes, st_cls;          // TMA-store epilogue: rows of st_row floats staged in shared memory, one box per class and
                                         // tile; bytes of one staging buffer (ncls class blocks) and of a class block (128 pixels x
                                         // st_row*4 B, rounded to 1 KB); st_nbuf buffers after the plane ring
  int st_nbuf, st_row, st_xor;           // st_xor: 16-byte chunk index ^= (128-byte line index & st_xor) -- the pattern of the store
                                         // map's swizzle mode (7 = 128B rows, 3 = 64B rows, 1 = 32B rows, 0 = none: dense rows)
  int st_nbox;                           // TMA boxes per tile: one per class, or one per output-row parity when the two column
                                         // classes of a row parity are staged side by side (rows of 2*Co floats: wider global segments)
  short st_cblk[HALO_MAXCLS], st_coff[HALO_MAXCLS];      // class -> staging block, float offset inside the staged row
  short st_bc0[HALO_MAXCLS], st_bpy[HALO_MAXCLS];        // box -> channel coordinate, row-parity coordinate
  double* stats;                         // != nullptr: BatchNorm statistics of the OUTPUT fused into the epilogue: row (stats_row0 + blockIdx.x) of
  int stats_row0;                        // [rows][2*Co] doubles receives this CTA's (sum y, sum y^2) per output channel (its cout slice)
  int pair;                              // CTA pairs (cta_group::2): one MMA covers the tiles of both CTAs (M = 256), each CTA holds half of
                                         // every weight block; loads complete on the leader's barriers, commits are multicast
  int pf_dist;                           // L2 prefetch distance in tiles (0: no prefetch); DCGANSR_HALO_PFDIST overrides the default
  int dbg;                               // DCGANSR_HALO_DBG (timing experiments only): 1 skip MMAs, 2 skip stores, 4 skip the halo TMA loads
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// K-major descriptor with explicit SBO and layout (start address may be any 16-byte aligned window of the tile)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr, uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// 8 epilogue warps: two per TMEM lane quarter (`half` 0 / 1), which split the tile's (class, 16-column chunk) items between
// them.  Per item all TMEM loads (one per accumulator slot of the class) are issued before a single wait.
template <int ACT>
__device__ __forceinline__ void halo_epilogue(const HaloParams& p, float* __restrict__ out, uint32_t tmem_base, uint64_t* acc_full,
                                              uint64_t* acc_empty, int warp, int half, int lane, int n0, const CUtensorMap* mapO,
                                              uint32_t sO, float* sred, uint32_t crank) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;                  // tile row = pixel
  const int w = r % HALO_TW, h = r / HALO_TW;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const int chunks = p.Npad >> 4;
  const int nitems = p.ncls * chunks;
  // fused BatchNorm statistics (chunks <= 2: every item of this warp half is the same 16-column chunk): per-thread fp32 sums over
  // this thread's pixel row of all its tiles, reduced once at the end
  const bool dostats = p.stats != nullptr;
  float s1[16], s2[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  int it = 0;
  // pair mode: the pair walks tile pairs, CTA `crank` owns the second tile of each (one past the end when ntiles is odd)
  const int tstep = p.pair ? (int)gridDim.x : (int)gridDim.x, tfirst = p.pair ? (int)(blockIdx.x & ~1u) + (int)crank : (int)blockIdx.x;
  const int tlimit = p.pair ? ((p.ntiles + 1) & ~1) : p.ntiles;
  for (int tile = tfirst; tile < tlimit; tile += tstep, ++it) {
    const int buf = it % p.nacc;
    const uint32_t aph = (uint32_t)(it / p.nacc) & 1u;
    int tt = tile;
    const int tx = tt % p.tiles_x; tt /= p.tiles_x;
    const int ty = tt % p.tiles_y; tt /= p.tiles_y;
    const int n = tt, gy = ty * HALO_TH + h, gx = tx * HALO_TW + w;
    const bool valid = tile < p.ntiles && gy < p.Hg && gx < p.Wg && !(p.dbg & 2);
    float* pix = out + ((int64_t)(n * p.Ho + gy * p.so) * p.Wo + gx * p.so) * p.Co + n0;
    const uint32_t cbase = lane_base + (uint32_t)(buf * p.acc_cols);
    // TMA-store mode: with two staging buffers, buffer it & 1 was the source of the store group issued two tiles ago; with
    // one, of the previous tile's group
    const bool leader = warp == p.ngrp + 1 && lane == 0;
    const uint32_t stg = sO + (uint32_t)((p.st_nbuf == 2 ? (it & 1) : 0) * p.st_bytes);
    if (p.tstore) {
      if (leader) { if (p.st_nbuf == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>(); }
      named_bar_sync(2, 256);
    }
    mbar_wait(smem_u32(&acc_full[buf]), aph);
    tc_fence_after();
    for (int item = half; item < nitems; item += 2) {
      const int c = item / chunks, c0 = (item % chunks) << 4;
      float* orow = pix + ((int64_t)p.coy[c] * p.Wo + p.cox[c]) * p.Co + c0;
      const int nsl = p.cls_nsl[c];
      uint32_t v[16];
      tmem_ld16(cbase + (uint32_t)(p.cls_sl[c][0] * p.Npad + c0), v);
      if (nsl == 1) {
        tmem_ld_wait();
      } else if (nsl == 2) {
        uint32_t u[16];
        tmem_ld16(cbase + (uint32_t)(p.cls_sl[c][1] * p.Npad + c0), u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
      } else {
        uint32_t u[16], x[16], y[16];
        tmem_ld16(cbase + (uint32_t)(p.cls_sl[c][1] * p.Npad + c0), u);
        tmem_ld16(cbase + (uint32_t)(p.cls_sl[c][2] * p.Npad + c0), x);
        if (nsl > 3) tmem_ld16(cbase + (uint32_t)(p.cls_sl[c][3] * p.Npad + c0), y);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float sum = (__uint_as_float(v[j]) + __uint_as_float(u[j])) + __uint_as_float(x[j]);
          if (nsl > 3) sum += __uint_as_float(y[j]);
          v[j] = __float_as_uint(sum);
        }
      }
      if (dostats && tile < p.ntiles && gy < p.Hg && gx < p.Wg) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float y = __uint_as_float(v[j]); s1[j] += y; s2[j] = fmaf(y, y, s2[j]); }
      }
      if (p.tstore) {
        // row r of class c: st_row floats (this CTA's cout slice), dense, inside the (1 KB aligned) class block.  A swizzled
        // store map (rows of exactly 128 / 64 / 32 bytes) expects byte offset o at o ^ (((o >> 7) & st_xor) << 4): the 16-byte
        // chunk index XORed with the 128-byte line index -- which also makes the 32 lanes' stores conflict-free; other row
        // widths use a non-swizzled map (st_xor = 0, dense rows)
        const uint32_t blk = stg + (uint32_t)(p.st_cblk[c] * p.st_cls);
        const int coff = p.st_coff[c];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c0 + j < p.Npad && n0 + c0 + j < p.Co) {
            const uint32_t o = (uint32_t)(r * p.st_row + coff + c0 + j) * 4u;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (o ^ (((o >> 7) & (uint32_t)p.st_xor) << 4))),
                         "f"(act_c<ACT>(__uint_as_float(v[j]), p.neg)), "f"(act_c<ACT>(__uint_as_float(v[j + 1]), p.neg)),
                         "f"(act_c<ACT>(__uint_as_float(v[j + 2]), p.neg)), "f"(act_c<ACT>(__uint_as_float(v[j + 3]), p.neg))
                         : "memory");
          }
        }
      } else if (valid) store_row<ACT, 16>(orow, v, n0 + c0, p.Co, p.neg);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (crank == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
      else mbar_arrive_remote(smem_u32(&acc_empty[buf]), 0);        // the leader's MMA issuers own the accumulator ring of both CTAs
    }
    if (p.tstore) {
      fence_proxy_async_smem();
      named_bar_sync(3, 256);
      if (leader && tile < p.ntiles && !(p.dbg & 2)) {
        const int gy0 = ty * HALO_TH, gx0 = tx * HALO_TW;
        // so == 2: output seen as [N][Ho/2][2][Wo/2][2*Co], class (coy, cox) = row parity coy, channel offset cox*Co;
        // so == 1: the same 5-D view [N][Ho][1][Wo][Co] with a unit parity dimension
        for (int b = 0; b < p.st_nbox; ++b)
          tma_store_5d(mapO, stg + (uint32_t)(b * p.st_cls), p.st_bc0[b] + n0, gx0, p.st_bpy[b], gy0, n);
        tma_store_commit();
      }
    }
  }
  if (p.tstore && warp == p.ngrp + 1 && lane == 0) tma_store_wait_all();
  if (dostats) {
    // rows (lanes) of a warp, then the four lane-quarter warps of this half: one (sum, sum^2) per channel of the half's chunk
#pragma unroll
    for (int j = 0; j < 16; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o); }
    }
    const int ew = half * 4 + q;                 // 0..7
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { sred[ew * 32 + j] = s1[j]; sred[ew * 32 + 16 + j] = s2[j]; }
    }
    named_bar_sync(4, 256);
    const int t = (half * 4 + q) * 32 + lane;    // 256 epilogue threads
    if (t < 64) {
      // t = h2*32 + k: half h2, k < 16 -> sum of channel k of that half's chunk, k >= 16 -> sum of squares
      const int h2 = t >> 5, k = t & 31;
      const int chunk = chunks == 1 ? 0 : h2;
      if (!(chunks == 1 && h2 == 1) ) {
        float a = 0.f;
        if (chunks == 1) { for (int w8 = 0; w8 < 8; ++w8) a += sred[w8 * 32 + k]; }       // both halves worked on chunk 0
        else { for (int w4 = 0; w4 < 4; ++w4) a += sred[(h2 * 4 + w4) * 32 + k]; }
        const int ch = n0 + chunk * 16 + (k & 15);
        if (ch < p.Co) p.stats[((size_t)(p.stats_row0 + blockIdx.x) * 2 + (k >> 4)) * p.Co + ch] = (double)a;
      }
    }
  }
}

// PAIR: a separate instantiation, because a kernel that contains cta_group::2 instructions cannot be launched without a cluster
// ("cluster misconfiguration"), whatever path it takes at run time
template <bool PAIR>
__global__ void __launch_bounds__(32 * (9 + HALO_MAXGRP), 1) tapconv_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ HaloMaps mapsB,
                                                                    const __grid_constant__ HaloTab tab,
                                                                    const __grid_constant__ CUtensorMap mapO,
                                                                    const __grid_constant__ HaloWList wl,
                                                                    const HaloParams p, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + p.w_bytes;
  uint8_t* sOut = sA + (size_t)p.nring * p.plane_bytes;                  // TMA-store staging (st_nbuf buffers) when p.tstore
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + (size_t)(p.tstore ? p.st_nbuf * p.st_bytes : 0));
  uint64_t* w_full = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* acc_empty = bars + 5;
  uint64_t* a_full = bars + 9;                                           // [HALO_MAXRING]
  uint64_t* a_empty = a_full + HALO_MAXRING;                             // [HALO_MAXRING]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + HALO_MAXRING);
  float* sred = reinterpret_cast<float*>(tmem_slot + 4);                 // 8 warps x 32 floats (fused BatchNorm statistics)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.Npad;
  uint32_t crank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < p.nring; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), (uint32_t)p.ngrp);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&acc_full[s]), (uint32_t)p.ngrp);
      mbar_init(smem_u32(&acc_empty[s]), PAIR ? 16 : 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // TMEM owner
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // both CTAs' barriers exist before any load / commit / remote arrive targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tstep = (int)gridDim.x, tfirst = PAIR ? (int)(blockIdx.x & ~1u) + (int)crank : (int)blockIdx.x;
  const int tlimit = PAIR ? ((p.ntiles + 1) & ~1) : p.ntiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // resident weights: every (class, tap, chunk) tile of this CTA's cout slice
      const uint32_t wf = smem_u32(w_full);
      if (PAIR) {
        // this CTA's half of every weight block; both CTAs' boxes complete on the leader's barrier
        if (crank == 0) mbar_expect_tx(wf, 2u * (uint32_t)p.w_tx);
        const uint32_t wfl = wf & 0xFEFFFFFFu;
        for (int i = 0; i < wl.n; ++i) {
          const uint4 e = wl.e[crank][i];
          const int c = (int)(e.x & 0xFFu);
          tma_load_2d_2sm(smem_u32(sW) + (e.w << 4), (e.x >> 8) ? &mapsB.bh[c] : &mapsB.b[c], wfl, (int)e.y, n0 + (int)e.z);
        }
      } else {
        mbar_expect_tx(wf, (uint32_t)p.w_tx);
        for (int c = 0; c < p.ncls; ++c)
          for (int t = 0; t < p.ntaps[c]; ++t)
            for (int q = 0; q < p.kchunks; ++q)
              tma_load_2d(smem_u32(sW) + (uint32_t)(p.tap_wtile[c][t] + q * p.tap_wstride[c][t]) * p.wtile_bytes, &mapsB.b[c], wf,
                          t * p.Ci + q * p.KBw, n0);
      }
      // plane ring: plane pl of the CTA's it-th tile lives in buffer (it * nplanes + pl) % nring
      int rs = 0;
      uint32_t rph = 0;
      const int pf_dist = p.pf_dist;
      for (int tile = tfirst; tile < tlimit; tile += tstep) {
        int tt = tile;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y; tt /= p.tiles_y;
        const int n = tt, gy0 = ty * HALO_TH, gx0 = tx * HALO_TW;
        // L2 prefetch of the halo tile `pf_dist` tiles ahead: the smem ring is only 1-2 tiles deep (the weights take
        // most of the shared memory), which alone does not keep enough DRAM reads in flight
        {
          const int ptile = tile + pf_dist * tstep;
          if (pf_dist > 0 && ptile < p.ntiles) {
            int pt = ptile;
            const int ptx = pt % p.tiles_x; pt /= p.tiles_x;
            const int pty = pt % p.tiles_y; pt /= p.tiles_y;
            for (int pl = 0; pl < p.nplanes; ++pl) {
              if (p.si == 1)
                tma_prefetch_4d(&mapA, p.pl_c[pl], ptx * HALO_TW + p.pl_x[pl], pty * HALO_TH + p.pl_y[pl], pt);
              else
                tma_prefetch_5d(&mapA, p.pl_c[pl], ptx * HALO_TW + p.pl_x[pl], p.pl_py[pl], pty * HALO_TH + p.pl_y[pl], pt);
            }
          }
        }
        for (int pl = 0; pl < p.nplanes; ++pl) {
          mbar_wait(smem_u32(&a_empty[rs]), rph ^ 1u);
          const uint32_t fb = smem_u32(&a_full[rs]);
          const uint32_t dst = smem_u32(sA + (size_t)rs * p.plane_bytes);
          if (PAIR) {
            // a tile past the end (odd tile count): image index out of range, the box is zero-filled and still counts its bytes
            if (crank == 0) mbar_expect_tx(fb, 2u * (uint32_t)p.plane_tx);
            const uint32_t fbl = fb & 0xFEFFFFFFu;
            if (p.si == 1)
              tma_load_4d_2sm(dst, &mapA, fbl, p.pl_c[pl], gx0 + p.pl_x[pl], gy0 + p.pl_y[pl], n);
            else
              tma_load_5d_2sm(dst, &mapA, fbl, p.pl_c[pl], gx0 + p.pl_x[pl], p.pl_py[pl], gy0 + p.pl_y[pl], n);
          } else if (p.dbg & 4) mbar_arrive(fb);
          else {
            mbar_expect_tx(fb, (uint32_t)p.plane_tx);
            if (p.si == 1)
              tma_load_4d(dst, &mapA, fb, p.pl_c[pl], gx0 + p.pl_x[pl], gy0 + p.pl_y[pl], n);
            else
              tma_load_5d(dst, &mapA, fb, p.pl_c[pl], gx0 + p.pl_x[pl], p.pl_py[pl], gy0 + p.pl_y[pl], n);
          }
          if (++rs == p.nring) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp <= p.ngrp && crank != 0) {
    // pair mode, second CTA: the leader's issuers run the MMAs of both tiles
  } else if (warp <= p.ngrp) {
    // ===================== MMA issuers =====================
    // The MMAs of these thin layers are small (N = 16..64: 8..32 tensor cycles each) while issuing one costs a single
    // thread ~40 cycles (descriptor arithmetic + 5 R2UR), so the tile's MMA list is split over up to 4 issuer warps,
    // each accumulating into its own TMEM slot(s): a class per issuer when there are sub-pixel classes, a range of
    // taps otherwise (the epilogue then adds the partial accumulators).
    const int grp = warp - 1;
    const uint64_t wl = p.KBw == 32 ? 2ull : (p.KBw == 16 ? 4ull : 6ull);
    const uint32_t w_sbo = 8u * (uint32_t)p.KBw * 4u;
    const int ibeg = p.gbeg[grp];
    mbar_wait(smem_u32(w_full), 0);
    const uint64_t bdesc = make_desc_k(smem_u32(sW), w_sbo, wl);
    int it = 0, rs = 0;
    uint32_t rph = 0;
    for (int tile = tfirst; tile < tlimit; tile += tstep, ++it) {
      const int buf = it % p.nacc;
      const uint32_t aph = (uint32_t)(it / p.nacc) & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), aph ^ 1u);
      const uint32_t dbase = tmem_base + (uint32_t)(buf * p.acc_cols);
      int i = ibeg;
      // plane by plane (the table is sorted by plane): every issuer waits for every plane and releases it after its own MMAs
      // on it (possibly none) -- a plane buffer is refilled once all issuers have let go of it
      for (int pl = 0; pl < p.nplanes; ++pl) {
        mbar_wait(smem_u32(&a_full[rs]), rph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_desc_k(smem_u32(sA + (size_t)rs * p.plane_bytes), (uint32_t)p.pitch_bytes, (uint64_t)p.a_layout);
          const int iend = (p.dbg & 1) ? min(i + 1, (int)p.pend[grp][pl]) : (int)p.pend[grp][pl];
          if (PAIR) {
#pragma unroll 4
            for (int k = i; k < iend; ++k) {
              const uint4 e = tab.e[k];
              umma_tf32_2sm(dbase + (e.z & 0xFFFFu), adesc + (uint64_t)e.x, bdesc + (uint64_t)e.y, e.w, (uint32_t)((int)e.z < 0 ? 0 : 1));
            }
            umma_commit_2sm(smem_u32(&a_empty[rs]));
            if (pl == p.nplanes - 1) umma_commit_2sm(smem_u32(&acc_full[buf]));
          } else {
#pragma unroll 4
            for (int k = i; k < iend; ++k) {
              const uint4 e = tab.e[k];
              umma_tf32(dbase + (e.z & 0xFFFFu), adesc + (uint64_t)e.x, bdesc + (uint64_t)e.y, e.w, (uint32_t)((int)e.z < 0 ? 0 : 1));
            }
            umma_commit(smem_u32(&a_empty[rs]));
            if (pl == p.nplanes - 1) umma_commit(smem_u32(&acc_full[buf]));
          }
        }
        __syncwarp();
        i = p.pend[grp][pl];
        if (++rs == p.nring) { rs = 0; rph ^= 1u; }
      }
    }
  } else if (warp <= p.ngrp + 8) {
    const int half = (warp - p.ngrp - 1) >> 2;
    // ===================== epilogue: TMEM -> registers -> activation -> NHWC global =====================
    switch (p.act) {
      case ACT_RELU: halo_epilogue<ACT_RELU>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut), sred, crank); break;
      case ACT_LRELU: halo_epilogue<ACT_LRELU>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut), sred, crank); break;
      case ACT_TANH: halo_epilogue<ACT_TANH>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut), sred, crank); break;
      case ACT_SIGMOID: halo_epilogue<ACT_SIGMOID>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut), sred, crank); break;
      default: halo_epilogue<ACT_NONE>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut), sred, crank); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();      // neither CTA frees its TMEM / leaves while the pair's MMAs, commits or remote arrives may target it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int pow2_ge_h(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int floordiv2_h(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

struct HaloCfg { HaloParams p; HaloTab tab; HaloWList wl; int nsplit; size_t smem; int grid_x; };

#define HALO_SMEM_MAX 232448      // 227 KB: the sm_100 per-block dynamic shared memory limit

static int halo_max_clusters();
static bool halo_cfg_mode(const TapGeom* cls, int ncls, HaloCfg& c, bool pair);

// CTA pairs first (DCGANSR_HALO_PAIR: 0 never, 1 where the geometry allows, default), the single-CTA form otherwise.  The choice
// depends on the geometry only, never on the batch: the plan-time decisions of dcgansr.cu (halo_mode, weight packs) and the
// run-time configuration must agree.
static bool halo_cfg(const TapGeom* cls, int ncls, HaloCfg& c) {
  // Measured per layer (C2 / C3b / C4, scripts/bench_layers.py with DCGANSR_HALO_PAIR=2): a cta_group::2 MMA with N <= 128 costs
  // ~110 tensor cycles against 64 for the single-CTA M = 128 one, so pairs do not cut the tensor time of these thin layers
  // (FC 48->24 forward 1258 -> 1374 us, C2 FC 64->32 dgrad 143 -> 210 us).  What they do is CAPACITY: each CTA holds half of every
  // weight block, so a class group whose weights do not fit one CTA runs as ONE launch with merged classes (FC 96->48 forward,
  // 295 KB: four per-class launches 1056 us -> 679 us) and a cout-split layer whose plane ring is starved by its weights reads
  // its input once (FC 96->48 dgrad 958 -> 681 us).  DCGANSR_HALO_PAIR: 0 never, 1 that policy (default), 2 wherever possible.
  const char* e = getenv("DCGANSR_HALO_PAIR");
  const int pm = e ? atoi(e) : 1;
  const bool can_pair = pm && halo_max_clusters() >= 32;
  if (pm == 2 && can_pair && halo_cfg_mode(cls, ncls, c, true)) return true;
  const bool single = halo_cfg_mode(cls, ncls, c, false);
  // (single-CTA form kept when it runs in one or two cout slices with a plane ring of at least one whole tile)
  if (!can_pair || (single && (c.nsplit == 1 || (c.nsplit == 2 && c.p.nring >= c.p.nplanes)))) return single;
  HaloCfg c2;
  if (halo_cfg_mode(cls, ncls, c2, true) && (!single || c2.nsplit < c.nsplit)) { c = c2; return true; }
  return single;
}

static bool halo_cfg_mode(const TapGeom* cls, int ncls, HaloCfg& c, bool pair) {
  if (!tc_encode_fn() || ncls < 1 || ncls > HALO_MAXCLS) return false;
  HaloParams& p = c.p;
  memset(&p, 0, sizeof(p));
  memset(&c.tab, 0, sizeof(c.tab));
  c.wl.n = 0;
  p.pair = pair ? 1 : 0;
  const TapGeom& g0 = cls[0];
  const int Ci = g0.Ci;
  // channel counts: multiples of 32 or 16 map exactly; for si == 1 any multiple of 4 (the reference's ngf = 12 gives 48 / 24 / 12)
  // is taken with the last K chunk zero-filled by TMA (out-of-bounds channels read as 0, so whatever the weight tile holds
  // there contributes nothing) and only the K = 8 steps that hold real channels issued
  // For si == 2 (parity view, rows of 2*Ci floats cut into 32-float planes) a tap's channels start at px*Ci floats into the
  // row: any multiple of 8 works, the K = 8 steps are addressed individually.
  const bool exactC = Ci % 32 == 0 || Ci == 16;
  if (!exactC && getenv("DCGANSR_HALO_EXACT_C")) return false;
  if (!exactC && !(g0.si == 1 ? (Ci % 4 == 0 && Ci >= 8) : (Ci % 8 == 0 && Ci >= 8))) return false;
  if (g0.si != 1 && g0.si != 2) return false;
  if (g0.si == 2 && (ncls != 1 || g0.so != 1 || g0.Hi % 2 || g0.Wi % 2)) return false;
  if (g0.Hg < 8 || g0.Wg < HALO_TW) return false;            // small images: the per-tap kernel tiles several images
  for (int i = 0; i < ncls; ++i) {
    const TapGeom& g = cls[i];
    if (g.Hg != g0.Hg || g.Wg != g0.Wg || g.Ci != Ci || g.Co != g0.Co || g.si != g0.si || g.so != g0.so || g.Hi != g0.Hi ||
        g.Wi != g0.Wi || g.Ho != g0.Ho || g.Wo != g0.Wo)
      return false;
    if (g.ntaps < 1 || g.ntaps > HALO_MAXTAPS) return false;
  }
  p.N = g0.N; p.Hg = g0.Hg; p.Wg = g0.Wg; p.Ho = g0.Ho; p.Wo = g0.Wo; p.Co = g0.Co; p.so = g0.so; p.ncls = ncls; p.si = g0.si; p.Ci = Ci;
  p.tiles_x = (p.Wg + HALO_TW - 1) / HALO_TW;
  p.tiles_y = (p.Hg + HALO_TH - 1) / HALO_TH;
  p.ntiles = p.N * p.tiles_y * p.tiles_x;
  // weight rows (and, for si == 1, the A planes): 32-float chunks, or 16-float ones when that tiles Ci exactly (48 = 3 x 16
  // instead of 32 + a half-empty 32: a quarter less shared memory for the resident weights)
  p.KBw = (Ci >= 32 && !(Ci % 32 == 16 && !getenv("DCGANSR_HALO_KB32"))) ? 32 : 16;
  p.kchunks = (Ci + p.KBw - 1) / p.KBw;
  p.ksteps = p.KBw / 8;
  // K = 8 steps of chunk q that hold real channels, and their total per (class, tap)
  auto ks_of = [&](int q) { return (std::min(p.KBw, Ci - q * p.KBw) + 7) / 8; };
  int ks_tap = 0;
  for (int q = 0; q < p.kchunks; ++q) ks_tap += ks_of(q);
  int tap_kb[HALO_MAXTAPS] = {0};        // si == 2: where a tap's channels start inside the wide pixel row (floats)
  // ---- A planes ----
  if (p.si == 1) {
    p.row_bytes = p.KBw * 4;
    p.a_layout = p.KBw == 32 ? 2 : 4;
    int ymin = 1 << 20, ymax = -(1 << 20), xmin = 1 << 20, xmax = -(1 << 20);
    for (int i = 0; i < ncls; ++i)
      for (int t = 0; t < cls[i].ntaps; ++t) {
        ymin = std::min(ymin, cls[i].dy[t]); ymax = std::max(ymax, cls[i].dy[t]);
        xmin = std::min(xmin, cls[i].dx[t]); xmax = std::max(xmax, cls[i].dx[t]);
      }
    p.PH = HALO_TH + ymax - ymin; p.PW = HALO_TW + xmax - xmin;
    p.nplanes = p.kchunks;
    if (p.nplanes > HALO_MAXPLANES) return false;
    for (int q = 0; q < p.nplanes; ++q) { p.pl_c[q] = (short)(q * p.KBw); p.pl_x[q] = (short)xmin; p.pl_y[q] = (short)ymin; p.pl_py[q] = 0; }
    p.pitch_bytes = p.PW * p.row_bytes;
    for (int i = 0; i < ncls; ++i)
      for (int t = 0; t < cls[i].ntaps; ++t) {
        p.tap_plane[i][t] = 0;
        p.tap_aoff[i][t] = ((cls[i].dy[t] - ymin) * p.PW + (cls[i].dx[t] - xmin)) * p.row_bytes;
      }
  } else {
    // parity view [N][H/2][2][W/2][2*Ci]: rows of 2*Ci floats; planes = (row parity) x (32-float chunk of the wide pixel)
    p.row_bytes = 128;
    p.a_layout = 2;
    const int J = std::max(1, (2 * Ci + 31) / 32);
    // x range per column parity when the two parities live in different planes (Ci a multiple of 32); shared otherwise
    int ymin[2] = {1 << 20, 1 << 20}, ymax[2] = {-(1 << 20), -(1 << 20)}, xmin[2] = {1 << 20, 1 << 20}, xmax[2] = {-(1 << 20), -(1 << 20)};
    const TapGeom& g = cls[0];
    const bool split_px = Ci >= 32 && Ci % 32 == 0;
    for (int t = 0; t < g.ntaps; ++t) {
      const int fy = floordiv2_h(g.dy[t]), py = g.dy[t] - 2 * fy, fx = floordiv2_h(g.dx[t]), px = split_px ? g.dx[t] - 2 * fx : 0;
      ymin[py] = std::min(ymin[py], fy); ymax[py] = std::max(ymax[py], fy);
      xmin[px] = std::min(xmin[px], fx); xmax[px] = std::max(xmax[px], fx);
    }
    int ext = 0, extx = 0, pyidx[2] = {-1, -1}, np = 0;
    for (int py = 0; py < 2; ++py)
      if (ymin[py] <= ymax[py]) { ext = std::max(ext, ymax[py] - ymin[py]); pyidx[py] = np++; }
    for (int px = 0; px < 2; ++px)
      if (xmin[px] <= xmax[px]) extx = std::max(extx, xmax[px] - xmin[px]);
    p.PH = HALO_TH + ext; p.PW = HALO_TW + extx;
    p.nplanes = np * J;
    if (p.nplanes > HALO_MAXPLANES) return false;
    for (int py = 0; py < 2; ++py) {
      if (pyidx[py] < 0) continue;
      for (int j = 0; j < J; ++j) {
        const int pl = pyidx[py] * J + j;
        const int px = split_px ? j / (Ci / 32) : 0;
        p.pl_c[pl] = (short)(j * 32); p.pl_py[pl] = (short)py; p.pl_y[pl] = (short)ymin[py];
        p.pl_x[pl] = (short)(xmin[px] <= xmax[px] ? xmin[px] : 0);
      }
    }
    p.pitch_bytes = p.PW * p.row_bytes;
    for (int t = 0; t < g.ntaps; ++t) {
      const int fy = floordiv2_h(g.dy[t]), py = g.dy[t] - 2 * fy, fx = floordiv2_h(g.dx[t]), px = g.dx[t] - 2 * fx;
      p.tap_plane[0][t] = (unsigned short)(pyidx[py] * J);          // first plane of the row parity; the K step picks the chunk
      p.tap_aoff[0][t] = ((fy - ymin[py]) * p.PW + (fx - xmin[split_px ? px : 0])) * p.row_bytes;
      tap_kb[t] = px * Ci;                                           // floats into the wide (2*Ci) pixel
    }
  }
  p.plane_tx = p.PH * p.PW * p.row_bytes;
  p.plane_bytes = (p.plane_tx + 1023) / 1024 * 1024;
  // ---- cout slice so that the resident weights + >= 1 stage fit ----
  int ttot = 0;
  for (int i = 0; i < ncls; ++i) {
    p.ntaps[i] = cls[i].ntaps; p.coy[i] = (short)cls[i].oy0; p.cox[i] = (short)cls[i].ox0;
    ttot += cls[i].ntaps;
  }
  // 2 x 2 sub-pixel classes: taps of different classes that read the SAME shifted window are merged into one MMA
  // (N = 2 or 4 classes) -- the A operand is read from shared memory once per shift instead of once per (class, tap),
  // which is what bounds these N = 16..64 MMAs.  See plan_merged().
  int ring[5] = {-1, -1, -1, -1, -1};
  bool try_merge = ncls == 4 && p.so == 2 && !getenv("DCGANSR_HALO_NOMERGE");
  if (try_merge) {
    int ci[2][2] = {{-1, -1}, {-1, -1}};
    for (int i = 0; i < 4; ++i)
      if (cls[i].oy0 >= 0 && cls[i].oy0 < 2 && cls[i].ox0 >= 0 && cls[i].ox0 < 2) ci[cls[i].oy0][cls[i].ox0] = i;
    if (ci[0][0] < 0 || ci[0][1] < 0 || ci[1][0] < 0 || ci[1][1] < 0) try_merge = false;
    else { ring[0] = ci[0][0]; ring[1] = ci[0][1]; ring[2] = ci[1][1]; ring[3] = ci[1][0]; ring[4] = ci[0][0]; }
  }
  const int co16 = (p.Co + 15) / 16 * 16;
  const size_t fixed = 1024 + (16 + 2 * HALO_MAXRING) * sizeof(uint64_t) + 1024;      // alignment slack, barriers, statistics scratch
  auto set_npad = [&](int npad) {
    p.Npad = npad;
    p.wtile_bytes = npad * p.KBw * 4;            // a multiple of 1024 (npad % 16 == 0, KBw >= 16)
    p.w_bytes = ttot * p.kchunks * p.wtile_bytes / (pair ? 2 : 1);       // pair mode: each CTA holds half of every weight block
    p.w_tx = p.w_bytes;
    p.acc_cols = (try_merge ? 5 : (ncls > 1 ? ncls : std::min(HALO_MAXGRP, cls[0].ntaps))) * npad;
  };
  // Largest cout slice whose resident weights leave room for a useful plane ring: two whole tiles when they fit, never fewer
  // than min_ring planes (the issuer works on one plane while the next ones are in flight; L2 prefetch covers the DRAM latency).
  const int full_ring = std::min(HALO_MAXRING, 2 * p.nplanes);
  const int min_ring = p.nplanes == 1 ? 2 : std::min(full_ring, 3);
  auto ring_for = [&](size_t extra) {
    const size_t used = fixed + (size_t)p.w_bytes + extra;
    if (used >= HALO_SMEM_MAX) return 0;
    return (int)std::min<size_t>(full_ring, (HALO_SMEM_MAX - used) / p.plane_bytes);
  };
  bool found = false;
  for (int npad = std::min(co16, 256); npad >= 16 && !found; npad -= 16) {
    if (co16 % npad && npad != std::min(co16, 256)) continue;       // equal slices only
    if (try_merge && 4 * npad > 256) continue;                      // a merged MMA spans up to 4 slots (N <= 256)
    set_npad(npad);
    if (p.acc_cols > 512) continue;
    p.nring = ring_for(0);
    if (p.nring >= min_ring) found = true;
  }
  if (!found) return false;
  if (const char* e = getenv("DCGANSR_HALO_NRING")) p.nring = std::max(min_ring, std::min(p.nring, atoi(e)));      // timing experiments
  if (p.w_bytes > (1 << 20) || p.plane_bytes > (1 << 20)) return false;      // 16-bit (>> 4) offsets in the table

  struct Op { int a_plane, a_off, m, slot0, cl[4], tp[4]; bool init; };
  std::vector<Op> ops;
  bool merged = false;
  if (try_merge) {
    // group (class, tap) by shift
    struct Sh { int dy, dx, n, cl[4], tp[4]; };
    std::vector<Sh> shifts;
    for (int i = 0; i < 4; ++i)
      for (int t = 0; t < cls[i].ntaps; ++t) {
        Sh* f = nullptr;
        for (auto& sh : shifts) if (sh.dy == cls[i].dy[t] && sh.dx == cls[i].dx[t]) f = &sh;
        if (!f) { shifts.push_back(Sh{cls[i].dy[t], cls[i].dx[t], 0, {0, 0, 0, 0}, {0, 0, 0, 0}}); f = &shifts.back(); }
        if (f->n < 4) { f->cl[f->n] = i; f->tp[f->n] = t; ++f->n; } else try_merge = false;   // a class twice on one shift
      }
    auto tap_of = [&](const Sh& sh, int c) { for (int j = 0; j < sh.n; ++j) if (sh.cl[j] == c) return sh.tp[j]; return -1; };
    auto single = [&](const Sh& sh, int j, int slot) {
      Op o; memset(&o, 0, sizeof(o));
      o.a_plane = p.tap_plane[sh.cl[j]][sh.tp[j]]; o.a_off = p.tap_aoff[sh.cl[j]][sh.tp[j]]; o.m = 1; o.slot0 = slot; o.cl[0] = sh.cl[j]; o.tp[0] = sh.tp[j];
      return o;
    };
    int full_idx = -1;
    for (auto& sh : shifts) {
      if (!try_merge) break;
      int run0 = -1, runm = 0;
      if (sh.n == 4) { run0 = 0; runm = 4; }
      else if (sh.n == 2)
        for (int r = 0; r < 4; ++r)
          if ((ring[r] == sh.cl[0] && ring[r + 1] == sh.cl[1]) || (ring[r] == sh.cl[1] && ring[r + 1] == sh.cl[0])) { run0 = r; runm = 2; }
      if (run0 >= 0) {
        Op o; memset(&o, 0, sizeof(o));
        o.a_plane = p.tap_plane[sh.cl[0]][sh.tp[0]]; o.a_off = p.tap_aoff[sh.cl[0]][sh.tp[0]]; o.m = runm; o.slot0 = run0;
        for (int j = 0; j < runm; ++j) { o.cl[j] = ring[run0 + j]; o.tp[j] = tap_of(sh, ring[run0 + j]); }
        if (runm == 4) full_idx = (int)ops.size();
        ops.push_back(o);
      } else {
        for (int j = 0; j < sh.n; ++j) {
          int slot = 0;
          for (int r = 0; r < 4; ++r) if (ring[r] == sh.cl[j]) slot = r;
          ops.push_back(single(sh, j, slot));
        }
      }
    }
    if (try_merge && full_idx >= 0) {
      // slot 4 duplicates class ring[0] so that the pair {ring[3], ring[0]} is a contiguous run; it needs an op of its own
      // to initialise it: a single-class op of ring[0] is moved there.  Without one, that pair is split into singles.
      bool uses4 = false;
      for (auto& o : ops) if (o.slot0 + o.m > 4) uses4 = true;
      int init4 = -1;
      if (uses4)
        for (size_t i = 0; i < ops.size(); ++i) if (ops[i].m == 1 && ops[i].cl[0] == ring[0]) { init4 = (int)i; break; }
      if (uses4 && init4 < 0) {
        std::vector<Op> o2;
        for (auto& o : ops) {
          if (o.slot0 + o.m > 4) {
            Op a = o, b = o;
            a.m = 1; a.slot0 = 3; a.cl[0] = o.cl[0]; a.tp[0] = o.tp[0];
            b.m = 1; b.slot0 = 0; b.cl[0] = o.cl[1]; b.tp[0] = o.tp[1];
            b.a_plane = p.tap_plane[b.cl[0]][b.tp[0]]; b.a_off = p.tap_aoff[b.cl[0]][b.tp[0]];
            o2.push_back(a); o2.push_back(b);
          } else o2.push_back(o);
        }
        ops.swap(o2);
        uses4 = false;
        for (size_t i = 0; i < ops.size(); ++i) if (ops[i].m == 4) full_idx = (int)i;
      }
      if (init4 >= 0) ops[init4].slot0 = 4;
      // order: the all-class op first (initialises slots 0..3), then slot 4's initialiser, then the rest
      std::vector<Op> ord;
      ops[full_idx].init = true;
      ord.push_back(ops[full_idx]);
      if (init4 >= 0) { ops[init4].init = true; ord.push_back(ops[init4]); }
      for (size_t i = 0; i < ops.size(); ++i) if ((int)i != full_idx && (int)i != init4) ord.push_back(ops[i]);
      ops.swap(ord);
      merged = true;
      p.ngrp = 1; p.nslots = uses4 ? 5 : 4;
      for (int r = 0; r < 4; ++r) { p.cls_nsl[ring[r]] = 1; p.cls_sl[ring[r]][0] = (unsigned char)r; }
      if (uses4) { p.cls_nsl[ring[0]] = 2; p.cls_sl[ring[0]][1] = 4; }
      p.gbeg[0] = 0;
    }
  }
  if (!merged) {
    // issuer groups and accumulator slots: one class per issuer, or (single class) contiguous tap ranges with partial sums
    ops.clear();
    p.ngrp = ncls > 1 ? ncls : std::min(HALO_MAXGRP, p.ntaps[0]);
    p.nslots = p.ngrp;
    for (int cl = 0; cl < ncls; ++cl) p.cls_nsl[cl] = 0;
    for (int g = 0; g < p.ngrp; ++g) {
      p.gbeg[g] = (int)ops.size() * ks_tap;
      const int cl = ncls > 1 ? g : 0;
      const int tb = ncls > 1 ? 0 : g * p.ntaps[0] / p.ngrp, te = ncls > 1 ? p.ntaps[cl] : (g + 1) * p.ntaps[0] / p.ngrp;
      p.cls_sl[cl][p.cls_nsl[cl]++] = (unsigned char)g;
      for (int t = tb; t < te; ++t) {
        Op o; memset(&o, 0, sizeof(o));
        o.a_plane = p.tap_plane[cl][t]; o.a_off = p.tap_aoff[cl][t]; o.m = 1; o.slot0 = g; o.cl[0] = cl; o.tp[0] = t; o.init = t == tb;
        ops.push_back(o);
      }
    }
  }
  p.nmma = (int)ops.size() * ks_tap;
  if (p.nmma > HALO_MAXMMA) return false;
  p.gbeg[p.ngrp] = p.nmma;
  {
    // one entry per MMA, then a stable sort by (issuer group, plane): the issuers consume the plane ring in plane order
    struct Ent { int grp, plane; uint4 e; };
    std::vector<Ent> ents;
    int i = 0, wt = 0;       // wt: running weight-tile index (pair mode: in half tiles of Npad / 2 rows)
    const uint32_t hb = (uint32_t)p.wtile_bytes / 2;
    for (auto& o : ops) {
      int grp = 0;
      while (grp + 1 < p.ngrp && i >= p.gbeg[grp + 1]) ++grp;
      for (int j = 0; j < o.m; ++j) { p.tap_wtile[o.cl[j]][o.tp[j]] = (unsigned short)(wt + j); p.tap_wstride[o.cl[j]][o.tp[j]] = (unsigned short)o.m; }
      if (pair) {
        // rows [rank * N/2, (rank + 1) * N/2) of the o.m stacked class tiles: m = 1 half a tile, m = 2 one class each, m = 4 two each
        if (o.m == 3) return false;
        for (int q = 0; q < p.kchunks; ++q)
          for (int rank = 0; rank < 2; ++rank) {
            const uint32_t dst = (uint32_t)(wt + q * o.m) * hb;
            const int nl = o.m == 4 ? 2 : 1;
            for (int l = 0; l < nl; ++l) {
              const int j = o.m == 1 ? 0 : (o.m == 2 ? rank : 2 * rank + l);
              const int idx = c.wl.n + (rank == 0 ? l : l);      // both ranks fill the same slots
              if (idx >= HALO_MAXWL) return false;
              c.wl.e[rank][idx] = make_uint4((uint32_t)o.cl[j] | (o.m == 1 ? 0x100u : 0u), (uint32_t)(o.tp[j] * Ci + q * p.KBw),
                                            o.m == 1 ? (uint32_t)(rank * (p.Npad / 2)) : 0u, (dst + (uint32_t)l * p.wtile_bytes) >> 4);
            }
            if (rank == 1) c.wl.n += nl;
          }
      }
      for (int q = 0; q < p.kchunks; ++q)
        for (int k = 0; k < ks_of(q); ++k, ++i) {
          // si == 1: plane = K chunk.  si == 2: the tap's K step sits f floats into the wide pixel row -> plane f / 32
          const int f = tap_kb[o.tp[0]] + q * p.KBw + k * 8;
          const int plane = p.si == 1 ? o.a_plane + q : o.a_plane + f / 32;
          const uint32_t aoff = p.si == 1 ? (uint32_t)o.a_off + k * 32 : (uint32_t)o.a_off + (f % 32) * 4;
          const uint32_t woff = pair ? (uint32_t)(wt + q * o.m) * hb + k * 32 : (uint32_t)(wt + q * o.m) * p.wtile_bytes + k * 32;
          const uint32_t accum = (o.init && q == 0 && k == 0) ? 0u : 1u;
          // D = f32, A = B = tf32, K-major, M = 128 (pair mode: 256, 128 rows per CTA), N = m * Npad
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((o.m * p.Npad) >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24);
          if (plane < 0 || plane >= p.nplanes) return false;
          ents.push_back(Ent{grp, plane, make_uint4(aoff >> 4, woff >> 4, (uint32_t)(o.slot0 * p.Npad) | (accum ? 0u : 0x80000000u), idesc)});
        }
      wt += o.m * p.kchunks;
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.grp != b.grp ? a.grp < b.grp : a.plane < b.plane; });
    if (!merged) {
      // every issuer group owns one accumulator slot: its first MMA in ISSUE order initialises it
      int prev = -1;
      for (auto& en : ents) {
        if (en.grp != prev) { en.e.z |= 0x80000000u; prev = en.grp; }
        else en.e.z &= 0x7FFFFFFFu;
      }
    } else {
      // merged classes: planes are the K chunks and the generation order (all-class op, slot 4's initialiser, rest) is kept
      // inside every plane, so the initialising MMAs (chunk 0, K step 0) still come first
      for (size_t k = 1; k < ents.size(); ++k) if (ents[k].plane < ents[k - 1].plane) return false;
    }
    for (size_t k = 0; k < ents.size(); ++k) c.tab.e[k] = ents[k].e;
    size_t k = 0;
    for (int g = 0; g < p.ngrp; ++g)
      for (int pl = 0; pl < p.nplanes; ++pl) {
        while (k < ents.size() && ents[k].grp == g && ents[k].plane == pl) ++k;
        p.pend[g][pl] = (unsigned short)k;
      }
    if (k != ents.size()) return false;
  }
  p.acc_cols = p.nslots * p.Npad;
  if (p.acc_cols > 512) return false;
  c.nsplit = (p.Co + p.Npad - 1) / p.Npad;
  p.nacc = std::max(1, std::min(4, 512 / p.acc_cols));      // accumulator ring in TMEM
  if (const char* e = getenv("DCGANSR_HALO_NACC")) p.nacc = std::max(1, std::min(p.nacc, atoi(e)));
  p.tmem_cols = std::max(32, pow2_ge_h(p.nacc * p.acc_cols));
  if (p.tmem_cols > 512) return false;
  // TMA-store epilogue.  The direct epilogue (32-byte vector stores from 8 warps) tops out near 4 TB/s; staging the tile in
  // shared memory and leaving the write-out to the TMA engine took C 32->16 dgrad from 148 to 104 us.  Output rows of this
  // CTA's cout slice (st_row = min(Npad, Co) floats) are staged densely; rows of exactly 128 / 64 / 32 bytes use the matching
  // swizzled store map (conflict-free staging writes), any other width a non-swizzled map.  The output is seen as
  // [N][Ho/so][so][Wo/so][so*Co]: a class of a tile is one 5-D box.  Taken when the staging buffers (two, else one) fit
  // beside the resident weights and a plane ring of at least min_ring.
  p.tstore = 0;
  p.st_row = std::min(p.Npad, p.Co);
  p.st_nbox = ncls;
  for (int i = 0; i < ncls; ++i) { p.st_cblk[i] = (short)i; p.st_coff[i] = 0; p.st_bc0[i] = (short)(p.cox[i] * p.Co); p.st_bpy[i] = p.coy[i]; }
  // stride-2 outputs whose rows are not a multiple of 128 bytes: the two column classes (cox = 0, 1) of an output-row parity
  // are neighbours in memory ([..][Wo/2][2*Co]), so they are staged side by side and leave as ONE box of 2*Co-float rows --
  // 96-byte segments with 96-byte gaps become contiguous 192-byte ones (C 24->12 dgrad: 965 us with per-class boxes, 828 direct)
  if (p.so == 2 && ncls == 4 && c.nsplit == 1 && p.Co % 32 != 0 && 2 * p.Co <= 256 && !getenv("DCGANSR_HALO_NO_PAIR")) {
    bool ok = true;
    for (int i = 0; i < ncls; ++i) ok = ok && p.coy[i] >= 0 && p.coy[i] < 2 && p.cox[i] >= 0 && p.cox[i] < 2;
    if (ok) {
      p.st_row = 2 * p.Co;
      p.st_nbox = 2;
      for (int i = 0; i < ncls; ++i) { p.st_cblk[i] = p.coy[i]; p.st_coff[i] = (short)(p.cox[i] * p.Co); }
      for (int b = 0; b < 2; ++b) { p.st_bc0[b] = 0; p.st_bpy[b] = (short)b; }
    }
  }
  p.st_cls = (128 * p.st_row * 4 + 1023) / 1024 * 1024;
  p.st_bytes = p.st_nbox * p.st_cls;
  p.st_xor = p.st_row == 32 ? 7 : (p.st_row == 16 ? 3 : (p.st_row == 8 ? 1 : 0));
  p.st_nbuf = 0;
  {
    bool ok = (p.so == 1 || p.so == 2) && p.Ho % p.so == 0 && p.Wo % p.so == 0 && p.Co % 4 == 0 && p.Co % p.Npad % 4 == 0 &&
              (p.Co % p.Npad == 0 || c.nsplit == 1) && !getenv("DCGANSR_HALO_NO_TSTORE");
    // measured on B200 (same box, per-layer A/B): the staged epilogue wins where the four classes leave as 128-byte swizzled rows
    // (C 32->16 dgrad 150 -> 140 us) and for the per-class launches of a stride-2 output (FC 96->48 forward 1200 -> 852 us at
    // C3b); it loses with non-swizzled rows shared by four classes (C 24->12 dgrad 828 -> 965 us), with 64-byte rows (FC 32->16
    // forward 687 -> 701 us) and on single-class stride-1 outputs (FC 64->32 dgrad 153 -> 214 us): those keep the direct epilogue
    if (!(p.st_xor == 7 && ncls == 4 && p.st_nbox == 4) && !(ncls == 1 && p.so == 2) && !getenv("DCGANSR_HALO_TSTORE_ALL")) ok = false;
    for (int i = 0; i < ncls; ++i) ok = ok && p.coy[i] >= 0 && p.coy[i] < p.so && p.cox[i] >= 0 && p.cox[i] < p.so;
    if (ok) {
      for (int nb = 2; nb >= 1 && !p.tstore; --nb) {
        const int r = ring_for((size_t)nb * p.st_bytes);
        // two buffers only if the ring keeps its full depth or at least min_ring + 1; one buffer whenever min_ring still fits
        if (r >= min_ring && (nb == 1 || r >= std::min(full_ring, min_ring + 1))) { p.tstore = 1; p.st_nbuf = nb; p.nring = std::min(p.nring, r); }
      }
    }
  }
  c.smem = fixed + (size_t)p.w_bytes + (size_t)p.nring * p.plane_bytes + (size_t)p.st_nbuf * p.st_bytes;
  // one persistent CTA per SM (13 warps x ~100 registers; the 8 epilogue warps provide the memory-level parallelism)
  c.grid_x = std::max(1, std::min(p.ntiles, (NSM + c.nsplit - 1) / c.nsplit));
  if (pair) {
    if ((p.Npad / 2) % 8) return false;                                  // N / 2 rows per CTA, whole 8-row groups
    const int pairs = std::max(1, std::min((p.ntiles + 1) / 2, std::min(halo_max_clusters(), NSM / 2) / c.nsplit));
    c.grid_x = 2 * pairs;
  }
  return true;
}

// co-resident CTA pairs of the halo kernel at its shared-memory footprint (74 on a full B200)
static int halo_max_clusters() {
  static int v = -1;
  if (v >= 0) return v;
  v = 0;
  if (cudaFuncSetAttribute(tapconv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { cudaGetLastError(); return v; }
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3(2 * NSM); lc.blockDim = dim3(32 * (9 + HALO_MAXGRP)); lc.dynamicSmemBytes = 220 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tapconv_halo_kernel<true>, &lc) != cudaSuccess) { cudaGetLastError(); return v; }
  v = std::min(n, NSM / 2);
  return v;
}

// Supported AND expected to beat the per-tap kernel: the activations are read once only if the cout slices are few, and
// TMA / MMA overlap needs the double-buffered halo tile.  DCGANSR_HALO_ALL=1 (tests) takes every supported geometry.
bool halo_tapconv_supported(const TapGeom* classes, int ncls, bool allow_pair) {
  HaloCfg c;
  if (!halo_cfg(classes, ncls, c)) return false;
  if (!allow_pair && c.p.pair) return false;
  if (getenv("DCGANSR_HALO_ALL")) return true;
  // few cout slices (the input is read once per slice), and either the full two-tile ring or a spatially large layer (a short
  // ring leans on the L2 prefetch, which needs many tiles per CTA to pay off)
  const int full_ring = std::min(HALO_MAXRING, 2 * c.p.nplanes);
  return c.nsplit <= 2 && (c.p.nring >= full_ring || (int64_t)c.p.Hg * c.p.Wg >= 4096);
}

// rows of the partial-statistics matrix a launch on this geometry would write (0: the fused statistics are not available)
int halo_stats_rows(const TapGeom* classes, int ncls) {
  HaloCfg c;
  if (getenv("DCGANSR_NO_FUSED_STATS") || !halo_cfg(classes, ncls, c)) return 0;
  return c.p.Npad <= 32 ? c.grid_x : 0;
}

bool k_tapconv_halo(St st, const TapGeom* classes, int ncls, const float* const* bp, const float* in, float* out, int act,
                    float negval, std::string* err, double* stats, int stats_row0) {
  HaloCfg c;
  if (!halo_cfg(classes, ncls, c)) { if (err) *err = "geometry not supported by the halo kernel"; return false; }
  HaloParams& p = c.p;
  p.act = act; p.neg = negval;
  p.stats = (stats && p.Npad <= 32) ? stats : nullptr; p.stats_row0 = stats_row0;
  // L2 prefetch policy (measured on the C3b layers, scripts/exp/halo_pf_ab.py): the prefetch doubles the read requests the L2 serves
  // (lts__t_sectors of FC 48->24 forward: 71 M sectors of TMA loads + 140 M of prefetches), which costs more than it hides once the
  // plane ring itself holds at least one tile (C 24->12 dgrad 937 -> 895 us, FC 48->24 dgrad 1046 -> 994, C 24->12 forward 811 ->
  // 794, FC 48->24 forward unchanged); CTA pairs (big resident weights, starved ring) keep it, one tile ahead (FC 96->48 forward
  // 677 -> 649 us at distance 1 instead of 2; its dgrad 678 at 1, 732 at 0 or 2)
  // a ring of exactly one tile (single-CTA form): C2's FC 64->32 dgrad 183 -> 169 us, C4's C 32->64 forward 54.7 -> 50.0 without prefetch
  p.pf_dist = p.pair ? 1 : (p.nring >= p.nplanes ? 0 : p.nring / p.nplanes + 1);
  if (const char* d = getenv("DCGANSR_HALO_PFDIST")) p.pf_dist = std::max(0, std::min(8, atoi(d)));      // timing experiments
  if (const char* d = getenv("DCGANSR_HALO_DBG")) p.dbg = atoi(d);
  const TapGeom& g = classes[0];
  EncodeTiledFn enc = tc_encode_fn();
  CUtensorMap mapA;
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  const CUtensorMapSwizzle aswz = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (p.si == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[4] = {(cuuint32_t)p.KBw, (cuuint32_t)p.PW, (cuuint32_t)p.PH, 1};
    r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, aswz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Ci, (cuuint64_t)g.Wi / 2, 2, (cuuint64_t)g.Hi / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Ci * 4, (cuuint64_t)g.Wi * g.Ci * 4, (cuuint64_t)2 * g.Wi * g.Ci * 4,
                             (cuuint64_t)g.Hi * g.Wi * g.Ci * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)p.PW, 1, (cuuint32_t)p.PH, 1};
    r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, aswz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(A, halo) failed: " + std::to_string((int)r); return false; }
  const CUtensorMapSwizzle wswz = p.KBw == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int i = 0; i < ncls; ++i) {
    const cuuint64_t ktot = (cuuint64_t)classes[i].ntaps * g.Ci;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)g.Co};
    cuuint64_t strides[1] = {ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)p.KBw, (cuuint32_t)p.Npad};
    r = enc(&maps.b[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bp[i], dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, wswz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(B, halo) failed: " + std::to_string((int)r); return false; }
    if (p.pair) {
      cuuint32_t boxh[2] = {(cuuint32_t)p.KBw, (cuuint32_t)(p.Npad / 2)};
      r = enc(&maps.bh[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)bp[i], dims, strides, boxh, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, wswz,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(B half, halo) failed: " + std::to_string((int)r); return false; }
    }
  }
  CUtensorMap mapO;
  memset(&mapO, 0, sizeof(mapO));
  if (p.tstore) {
    const cuuint64_t so = (cuuint64_t)p.so;
    cuuint64_t dims[5] = {so * g.Co, (cuuint64_t)g.Wo / so, so, (cuuint64_t)g.Ho / so, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {so * g.Co * 4, (cuuint64_t)g.Wo * g.Co * 4, so * g.Wo * g.Co * 4, (cuuint64_t)g.Ho * g.Wo * g.Co * 4};
    cuuint32_t box[5] = {(cuuint32_t)p.st_row, HALO_TW, 1, HALO_TH, 1};
    const CUtensorMapSwizzle oswz = p.st_xor == 7 ? CU_TENSOR_MAP_SWIZZLE_128B : p.st_xor == 3 ? CU_TENSOR_MAP_SWIZZLE_64B
                                    : p.st_xor == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    r = enc(&mapO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, oswz,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(out, halo) failed: " + std::to_string((int)r); return false; }
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(tapconv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem, halo) failed";
      return false;
    }
    configured = true;
  }
  dim3 grid((unsigned)c.grid_x, (unsigned)c.nsplit);
  if (p.pair) {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid; lc.blockDim = dim3(32 * (9 + p.ngrp)); lc.dynamicSmemBytes = c.smem; lc.stream = st.s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    if (cudaLaunchKernelEx(&lc, tapconv_halo_kernel<true>, mapA, maps, c.tab, mapO, c.wl, p, out) != cudaSuccess) {
      if (err) *err = std::string("halo pair launch failed: ") + cudaGetErrorString(cudaGetLastError());
      return false;
    }
  } else
    tapconv_halo_kernel<false><<<grid, 32 * (9 + p.ngrp), c.smem, st.s>>>(mapA, maps, c.tab, mapO, c.wl, p, out);
  // HBM-bound by construction (input and output once, weights resident): the roofline work of a launch is its
  // algorithmic bytes = input + output + weights (fp32)
  double wbytes = 0;
  for (int i = 0; i < ncls; ++i) wbytes += 4.0 * classes[i].ntaps * g.Ci * g.Co;
  const double bytes = 4.0 * ((double)g.N * g.Hi * g.Wi * g.Ci + (double)g.N * g.Ho * g.Wo * g.Co) + wbytes;
  DSR_LAUNCHED(st, "tapconv_halo", bytes, WORK_BYTES);
  return true;
}
