// kernels_halo.cu -- weights-resident, halo-tile tcgen05 implicit-GEMM convolution for the spatially large, thin
// layers of the generator (FAST_TF32): nn.SpatialFullConvolution / nn.SpatialConvolution forward and updateGradInput
// (train.lua:99-111, train-gray.lua:105-116) where the activations (hundreds of MB) dominate and the whole
// weight tensor fits in shared memory.
//
// What bounds these layers is HBM, and what bounded the per-tap kernel (kernels_tc.cu) was the L2 -> SM path: it
// re-fetches the input pixels once per tap (4..16x) and the weights once per 128-pixel tile.  Here
//
//   * a persistent CTA loads its slice of the packed weights ONCE (TMA, K-major, swizzled) and keeps it in smem;
//   * per 16 x 8 tile of the output grid it loads the input pixels ONCE: a halo tile (one TMA box per 32-channel
//     plane, zero fill = padding), double buffered;
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

struct HaloMaps { CUtensorMap b[HALO_MAXCLS]; };
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
  int nstage, stage_bytes, acc_cols, nacc, tmem_cols, nmma;
  int ngrp, gbeg[HALO_MAXGRP + 1];      // MMA issuer warps and their table ranges
  int nslots;                            // accumulator slots (Npad TMEM columns each)
  int cls_nsl[HALO_MAXCLS];              // class c = sum of slots cls_sl[c][0 .. cls_nsl[c])
  unsigned char cls_sl[HALO_MAXCLS][4];
  int act;
  float neg;
  int tstore, st_bytes, st_cls;          // TMA-store epilogue: rows of Co floats staged in shared memory (SWIZZLE_128B pattern), one box
                                         // per class and tile; bytes of one staging buffer (ncls class blocks) and of a class block
                                         // (128 pixels x Co*4 B, rounded to 1 KB); two buffers after the A stages
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
                                              uint32_t sO) {
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;                  // tile row = pixel
  const int w = r % HALO_TW, h = r / HALO_TW;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const int chunks = p.Npad >> 4;
  const int nitems = p.ncls * chunks;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int buf = it % p.nacc;
    const uint32_t aph = (uint32_t)(it / p.nacc) & 1u;
    int tt = tile;
    const int tx = tt % p.tiles_x; tt /= p.tiles_x;
    const int ty = tt % p.tiles_y; tt /= p.tiles_y;
    const int n = tt, gy = ty * HALO_TH + h, gx = tx * HALO_TW + w;
    const bool valid = gy < p.Hg && gx < p.Wg && !(p.dbg & 2);
    float* pix = out + ((int64_t)(n * p.Ho + gy * p.so) * p.Wo + gx * p.so) * p.Co + n0;
    const uint32_t cbase = lane_base + (uint32_t)(buf * p.acc_cols);
    // TMA-store mode: staging buffer it & 1 was the source of the store group issued two tiles ago
    const bool leader = warp == p.ngrp + 1 && lane == 0;
    const uint32_t stg = sO + (uint32_t)((it & 1) * p.st_bytes);
    if (p.tstore) {
      if (leader) tma_store_wait_read<1>();
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
      if (p.tstore) {
        // row r of class c: Co floats at pitch Co*4 inside the (1 KB aligned) class block; the SWIZZLE_128B store map expects
        // byte offset o at o ^ (((o >> 7) & 7) << 4) (16-byte chunk index XOR 128-byte line index, mod 8)
        const uint32_t blk = stg + (uint32_t)(c * p.st_cls);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c0 + j < p.Co) {
            const uint32_t o = (uint32_t)(r * p.Co + c0 + j) * 4u;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (o ^ (((o >> 7) & 7u) << 4))),
                         "f"(act_c<ACT>(__uint_as_float(v[j]), p.neg)), "f"(act_c<ACT>(__uint_as_float(v[j + 1]), p.neg)),
                         "f"(act_c<ACT>(__uint_as_float(v[j + 2]), p.neg)), "f"(act_c<ACT>(__uint_as_float(v[j + 3]), p.neg))
                         : "memory");
          }
        }
      } else if (valid) store_row<ACT, 16>(orow, v, n0 + c0, p.Co, p.neg);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
    if (p.tstore) {
      fence_proxy_async_smem();
      named_bar_sync(3, 256);
      if (leader && !(p.dbg & 2)) {
        const int gy0 = ty * HALO_TH, gx0 = tx * HALO_TW;
        for (int c = 0; c < p.ncls; ++c)      // output seen as [N][Ho/2][2][Wo/2][2*Co]: class (coy, cox) = row parity, channel offset cox*Co
          tma_store_5d(mapO, stg + (uint32_t)(c * p.st_cls), p.cox[c] * p.Co, gx0, p.coy[c], gy0, n);
        tma_store_commit();
      }
    }
  }
  if (p.tstore && warp == p.ngrp + 1 && lane == 0) tma_store_wait_all();
}

__global__ void __launch_bounds__(32 * (9 + HALO_MAXGRP), 1) tapconv_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ HaloMaps mapsB,
                                                                    const __grid_constant__ HaloTab tab,
                                                                    const __grid_constant__ CUtensorMap mapO,
                                                                    const HaloParams p, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + p.w_bytes;
  uint8_t* sOut = sA + (size_t)p.nstage * p.stage_bytes;                 // TMA-store staging (2 buffers) when p.tstore
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + (size_t)(p.tstore ? 2 * p.st_bytes : 0));
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 3;
  uint64_t* acc_full = bars + 5;
  uint64_t* acc_empty = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.Npad;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), (uint32_t)p.ngrp);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&acc_full[s]), (uint32_t)p.ngrp);
      mbar_init(smem_u32(&acc_empty[s]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // TMEM owner
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
      // resident weights: every (class, tap, chunk) tile of this CTA's cout slice
      const uint32_t wf = smem_u32(w_full);
      mbar_expect_tx(wf, (uint32_t)p.w_tx);
      for (int c = 0; c < p.ncls; ++c)
        for (int t = 0; t < p.ntaps[c]; ++t)
          for (int q = 0; q < p.kchunks; ++q)
            tma_load_2d(smem_u32(sW) + (uint32_t)(p.tap_wtile[c][t] + q * p.tap_wstride[c][t]) * p.wtile_bytes, &mapsB.b[c], wf,
                        t * p.Ci + q * p.KBw, n0);
      int it = 0;
      const int pf_dist = p.nstage + 1;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int s = it % p.nstage;
        const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
        int tt = tile;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y; tt /= p.tiles_y;
        const int n = tt, gy0 = ty * HALO_TH, gx0 = tx * HALO_TW;
        // L2 prefetch of the halo tile `pf_dist` tiles ahead: the smem ring is only 1-2 tiles deep (the weights take
        // most of the shared memory), which alone does not keep enough DRAM reads in flight
        {
          const int ptile = tile + pf_dist * (int)gridDim.x;
          if (ptile < p.ntiles) {
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
        mbar_wait(smem_u32(&a_empty[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&a_full[s]);
        if (p.dbg & 4) { mbar_arrive(fb); continue; }
        mbar_expect_tx(fb, (uint32_t)(p.nplanes * p.plane_tx));
        const uint32_t dst0 = smem_u32(sA + (size_t)s * p.stage_bytes);
        for (int pl = 0; pl < p.nplanes; ++pl) {
          const uint32_t dst = dst0 + (uint32_t)pl * p.plane_bytes;
          if (p.si == 1)
            tma_load_4d(dst, &mapA, fb, p.pl_c[pl], gx0 + p.pl_x[pl], gy0 + p.pl_y[pl], n);
          else
            tma_load_5d(dst, &mapA, fb, p.pl_c[pl], gx0 + p.pl_x[pl], p.pl_py[pl], gy0 + p.pl_y[pl], n);
        }
      }
    }
  } else if (warp <= p.ngrp) {
    // ===================== MMA issuers =====================
    // The MMAs of these thin layers are small (N = 16..64: 8..32 tensor cycles each) while issuing one costs a single
    // thread ~40 cycles (descriptor arithmetic + 5 R2UR), so the tile's MMA list is split over up to 4 issuer warps,
    // each accumulating into its own TMEM slot(s): a class per issuer when there are sub-pixel classes, a range of
    // taps otherwise (the epilogue then adds the partial accumulators).
    const int grp = warp - 1;
    const uint64_t wl = p.KBw == 32 ? 2ull : (p.KBw == 16 ? 4ull : 6ull);
    const uint32_t w_sbo = 8u * (uint32_t)p.KBw * 4u;
    const int ibeg = p.gbeg[grp], iend = p.gbeg[grp + 1];
    mbar_wait(smem_u32(w_full), 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.nstage;
      const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
      const int buf = it % p.nacc;
      const uint32_t aph = (uint32_t)(it / p.nacc) & 1u;
      mbar_wait(smem_u32(&acc_empty[buf]), aph ^ 1u);
      mbar_wait(smem_u32(&a_full[s]), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_desc_k(smem_u32(sA + (size_t)s * p.stage_bytes), (uint32_t)p.pitch_bytes, (uint64_t)p.a_layout);
        const uint64_t bdesc = make_desc_k(smem_u32(sW), w_sbo, wl);
        const uint32_t dbase = tmem_base + (uint32_t)(buf * p.acc_cols);
#pragma unroll 4
        for (int i = ibeg; i < ((p.dbg & 1) ? ibeg + 1 : iend); ++i) {
          const uint4 e = tab.e[i];
          umma_tf32(dbase + (e.z & 0xFFFFu), adesc + (uint64_t)e.x, bdesc + (uint64_t)e.y, e.w, (uint32_t)((int)e.z < 0 ? 0 : 1));
        }
        umma_commit(smem_u32(&a_empty[s]));
        umma_commit(smem_u32(&acc_full[buf]));
      }
      __syncwarp();
    }
  } else if (warp <= p.ngrp + 8) {
    const int half = (warp - p.ngrp - 1) >> 2;
    // ===================== epilogue: TMEM -> registers -> activation -> NHWC global =====================
    switch (p.act) {
      case ACT_RELU: halo_epilogue<ACT_RELU>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut)); break;
      case ACT_LRELU: halo_epilogue<ACT_LRELU>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut)); break;
      case ACT_TANH: halo_epilogue<ACT_TANH>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut)); break;
      case ACT_SIGMOID: halo_epilogue<ACT_SIGMOID>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut)); break;
      default: halo_epilogue<ACT_NONE>(p, out, tmem_base, acc_full, acc_empty, warp, half, lane, n0, &mapO, smem_u32(sOut)); break;
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
static inline int pow2_ge_h(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int floordiv2_h(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

struct HaloCfg { HaloParams p; HaloTab tab; int nsplit; size_t smem; int grid_x; };

#define HALO_SMEM_MAX 232448      // 227 KB: the sm_100 per-block dynamic shared memory limit

static bool halo_cfg(const TapGeom* cls, int ncls, HaloCfg& c) {
  if (!tc_encode_fn() || ncls < 1 || ncls > HALO_MAXCLS) return false;
  HaloParams& p = c.p;
  memset(&p, 0, sizeof(p));
  memset(&c.tab, 0, sizeof(c.tab));
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
  p.KBw = Ci >= 32 ? 32 : 16;
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
  p.stage_bytes = p.nplanes * p.plane_bytes;
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
  const size_t fixed = 1024 + 16 * sizeof(uint64_t);
  auto set_npad = [&](int npad) {
    p.Npad = npad;
    p.wtile_bytes = npad * p.KBw * 4;            // a multiple of 1024 (npad % 16 == 0, KBw >= 16)
    p.w_bytes = ttot * p.kchunks * p.wtile_bytes;
    p.w_tx = p.w_bytes;
    p.acc_cols = (try_merge ? 5 : (ncls > 1 ? ncls : std::min(HALO_MAXGRP, cls[0].ntaps))) * npad;
  };
  // largest cout slice whose weights leave room for a double-buffered halo tile; failing that, a single stage
  bool found = false;
  for (int want = 2; want >= 1 && !found; --want)
    for (int npad = std::min(co16, 256); npad >= 16; npad -= 16) {
      if (co16 % npad && npad != std::min(co16, 256)) continue;       // equal slices only
      if (try_merge && 4 * npad > 256) continue;                      // a merged MMA spans up to 4 slots (N <= 256)
      set_npad(npad);
      if (p.acc_cols <= 512 && p.w_bytes + (size_t)want * p.stage_bytes + fixed <= HALO_SMEM_MAX) { p.nstage = want; found = true; break; }
    }
  if (!found) return false;
  if (const char* e = getenv("DCGANSR_HALO_NSTAGE")) p.nstage = std::max(1, std::min(p.nstage, atoi(e)));      // timing experiments
  if (p.w_bytes > (1 << 20) || p.stage_bytes > (1 << 20)) return false;      // 16-bit (>> 4) offsets in the table

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
    int i = 0, wt = 0;       // wt: running weight-tile index
    for (auto& o : ops) {
      for (int j = 0; j < o.m; ++j) { p.tap_wtile[o.cl[j]][o.tp[j]] = (unsigned short)(wt + j); p.tap_wstride[o.cl[j]][o.tp[j]] = (unsigned short)o.m; }
      for (int q = 0; q < p.kchunks; ++q)
        for (int k = 0; k < ks_of(q); ++k, ++i) {
          // si == 1: plane = K chunk.  si == 2: the tap's K step sits f floats into the wide pixel row -> plane f / 32
          const int f = tap_kb[o.tp[0]] + q * p.KBw + k * 8;
          const uint32_t aoff = p.si == 1 ? (uint32_t)(o.a_plane + q) * p.plane_bytes + (uint32_t)o.a_off + k * 32
                                          : (uint32_t)(o.a_plane + f / 32) * p.plane_bytes + (uint32_t)o.a_off + (f % 32) * 4;
          const uint32_t woff = (uint32_t)(wt + q * o.m) * p.wtile_bytes + k * 32;
          const uint32_t accum = (o.init && q == 0 && k == 0) ? 0u : 1u;
          // D = f32, A = B = tf32, K-major, M = 128, N = m * Npad
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((o.m * p.Npad) >> 3) << 17) | ((128u >> 4) << 24);
          c.tab.e[i] = make_uint4(aoff >> 4, woff >> 4, (uint32_t)(o.slot0 * p.Npad) | (accum ? 0u : 0x80000000u), idesc);
        }
      wt += o.m * p.kchunks;
    }
  }
  p.acc_cols = p.nslots * p.Npad;
  if (p.acc_cols > 512) return false;
  c.nsplit = (p.Co + p.Npad - 1) / p.Npad;
  p.nacc = std::max(1, std::min(4, 512 / p.acc_cols));      // accumulator ring in TMEM
  if (const char* e = getenv("DCGANSR_HALO_NACC")) p.nacc = std::max(1, std::min(p.nacc, atoi(e)));
  p.tmem_cols = std::max(32, pow2_ge_h(p.nacc * p.acc_cols));
  if (p.tmem_cols > 512) return false;
  c.smem = 1024 + (size_t)p.w_bytes + (size_t)p.nstage * p.stage_bytes + 16 * sizeof(uint64_t);
  // TMA-store epilogue where an output row fits one 128-byte swizzle row (Co <= 32, one cout slice), the four sub-pixel classes
  // of a stride-2 output are present (the output viewed as [N][Ho/2][2][Wo/2][2*Co] makes a class one 5-D box per tile) and
  // two staging buffers still fit beside the resident weights and the halo stages.  The direct epilogue (32-byte vector stores
  // from 8 warps) tops out near 4 TB/s; the staged one leaves the write-out to the TMA engine (C 32->16 dgrad 148 -> 104 us).
  p.tstore = 0;
  p.st_cls = (128 * p.Co * 4 + 1023) / 1024 * 1024;
  p.st_bytes = ncls * p.st_cls;
  // (rows narrower than 128 bytes -- Co = 24 / 16 -- were tried with a Co-float box: wrong results, TMA does not pack such rows the
  //  way the staging writer assumed; only full 128-byte rows are taken)
  if (p.so == 2 && ncls == 4 && p.Co == 32 && p.Npad == 32 && c.nsplit == 1 && p.Ho % 2 == 0 && p.Wo % 2 == 0 &&
      c.smem + 2 * (size_t)p.st_bytes <= HALO_SMEM_MAX && !getenv("DCGANSR_HALO_NO_TSTORE")) {
    bool ok = true;
    for (int i = 0; i < ncls; ++i) ok = ok && p.coy[i] >= 0 && p.coy[i] < 2 && p.cox[i] >= 0 && p.cox[i] < 2;
    if (ok) { p.tstore = 1; c.smem += 2 * (size_t)p.st_bytes; }
  }
  // one persistent CTA per SM (13 warps x ~100 registers; the 8 epilogue warps provide the memory-level parallelism)
  c.grid_x = std::max(1, std::min(p.ntiles, (NSM + c.nsplit - 1) / c.nsplit));
  return true;
}

// Supported AND expected to beat the per-tap kernel: the activations are read once only if the cout slices are few, and
// TMA / MMA overlap needs the double-buffered halo tile.  DCGANSR_HALO_ALL=1 (tests) takes every supported geometry.
bool halo_tapconv_supported(const TapGeom* classes, int ncls) {
  HaloCfg c;
  if (!halo_cfg(classes, ncls, c)) return false;
  if (getenv("DCGANSR_HALO_ALL")) return true;
  return c.nsplit <= 2 && c.p.nstage == 2;
}

bool k_tapconv_halo(St st, const TapGeom* classes, int ncls, const float* const* bp, const float* in, float* out, int act,
                    float negval, std::string* err) {
  HaloCfg c;
  if (!halo_cfg(classes, ncls, c)) { if (err) *err = "geometry not supported by the halo kernel"; return false; }
  HaloParams& p = c.p;
  p.act = act; p.neg = negval;
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
  }
  CUtensorMap mapO;
  memset(&mapO, 0, sizeof(mapO));
  if (p.tstore) {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Co, (cuuint64_t)g.Wo / 2, 2, (cuuint64_t)g.Ho / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Co * 4, (cuuint64_t)g.Wo * g.Co * 4, (cuuint64_t)2 * g.Wo * g.Co * 4,
                             (cuuint64_t)g.Ho * g.Wo * g.Co * 4};
    cuuint32_t box[5] = {(cuuint32_t)g.Co, HALO_TW, 1, HALO_TH, 1};
    r = enc(&mapO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)out, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(out, halo) failed: " + std::to_string((int)r); return false; }
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tapconv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem, halo) failed";
      return false;
    }
    configured = true;
  }
  dim3 grid((unsigned)c.grid_x, (unsigned)c.nsplit);
  tapconv_halo_kernel<<<grid, 32 * (9 + p.ngrp), c.smem, st.s>>>(mapA, maps, c.tab, mapO, p, out);
  // HBM-bound by construction (input and output once, weights resident): the roofline work of a launch is its
  // algorithmic bytes = input + output + weights (fp32)
  double wbytes = 0;
  for (int i = 0; i < ncls; ++i) wbytes += 4.0 * classes[i].ntaps * g.Ci * g.Co;
  const double bytes = 4.0 * ((double)g.N * g.Hi * g.Wi * g.Ci + (double)g.N * g.Ho * g.Wo * g.Co) + wbytes;
  DSR_LAUNCHED(st, "tapconv_halo", bytes, WORK_BYTES);
  return true;
}
