// kernels_wgrad_halo.cu -- accGradParameters of nn.SpatialConvolution / nn.SpatialFullConvolution (train.lua:99-111)
// for the spatially large layers, as a halo-tile tcgen05 kernel (FAST_TF32).
//
//   acc[t][cp][cq] = sum_{n,gy,gx} P[n,gy,gx,cp] * Q[n, gy*s + dy_t, gx*s + dx_t, cq]          (P on the grid, Q shifted)
//
// A tcgen05.mma with M = 128 costs >= 64 tensor cycles whatever N <= 128 is (the A operand is read from shared memory
// at 64 B/clk), so the per-tap formulation of kernels_tc.cu (one N = Cq MMA per tap and 8 pixels, half of M empty when
// Cp = 64) runs the tensor pipe at ~12 % and re-fetches Q once per tap through L2.  Here, for the taps of one
// stride-parity plane of Q (k = 4, s = 2: 2 x 2 taps per plane; s = 1: the k x k taps of the single plane)
//
//     acc_{oy,ox}[cp][cq] = sum_g P[g][cp] QP[g + (oy,ox)][cq] = sum_g' P[g' - (oy,0)][cp] QP[g' + (0,ox)][cq]
//
// the ROW shift is moved to P and the COLUMN shift stays on QP, and both are stacked inside ONE MMA:
//   * A (MN-major) = P rows g'-oy for all ny row shifts: M = ny * Cp.  The P tile is loaded (one 5-D TMA box) as
//     [row][32-channel chunk][8 pixels] so that consecutive M atoms (shift, chunk) are a constant LBO = 1 KB apart;
//   * B (MN-major) = QP columns +ox for all nx column shifts: N = nx * 32, consecutive atoms LBO = 128 B (one pixel)
//     apart -- windows of the same plane tile;
//   * every operand window is just a start address inside a TMA-written tile: the UMMA swizzle is a function of the
//     absolute shared-memory address (scripts/exp/exp_desc.cu), so shifted windows are valid operands.
// One MMA per (plane, 8 pixels) instead of one per (tap, 8 pixels); P and Q are read from HBM exactly once.
// Split-K over persistent CTAs; partials -> scratch[cta][cp][t*Cq + cq]; k_wgrad_reduce adds them to the Torch7-layout
// master gradient in fixed order (deterministic).
#include "tc_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define NSM 148
#define WH_MAXPL 4
#define WH_MAXMMA 64
#define WH_THREADS 192

struct WhTab { uint4 e[WH_MAXMMA]; };       // {A offset >> 4, B offset >> 4, TMEM column, 0} per MMA of a K tile

struct WhParams {
  int N, Hg, Wg, Cp, Cq, T, s;
  int tiles_x, tiles_y, ntiles, tiles_per_cta;
  int ny, nx, npl, Qs, mch, q0_stride;     // row / column shifts per plane, planes, 32-channel chunks of Q, chunks of P per M tile
  int a5d;                                 // P tile through the 5-D [y][chunk][x] map (Cp >= 32) or a single zero-padded chunk
  int cstride, nch;                        // channels between the starts of consecutive 32-channel chunks of P (32; less when Cp
                                           // is not a multiple of 32: the chunks then OVERLAP so that none reads past the pixel),
                                           // and the number of chunks
  int RA, PW, a_bytes, plane_bytes, stage_bytes, nstage, nmma, ncols, tmem_cols;
  short pl_py[WH_MAXPL], pl_px[WH_MAXPL], pl_oy0[WH_MAXPL], pl_ox0[WH_MAXPL];
  short tap_of[WH_MAXPL][4][4][2];         // [plane][row shift i][col shift i2][16-column half] -> tap index (or -1)
  int packed;                              // Cq = 16, s = 2: a 128-byte row of the parity view holds BOTH column parities, so an N atom
                                           // = (px, 16 channels) and the planes are the row parities only
  long long split_stride;
};

// MN-major TF32 descriptor (SWIZZLE_128B_BASE32B): K rows of 128 B, 4-row groups every 512 B, MN atoms every `lbo` bytes
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

__global__ void __launch_bounds__(WH_THREADS) wgrad_halo_kernel(const __grid_constant__ CUtensorMap mapP,
                                                                const __grid_constant__ CUtensorMap mapQ,
                                                                const __grid_constant__ WhTab tab, const WhParams p,
                                                                float* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstage * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* done = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.y;                              // M tile: chunks [mt*mch, mt*mch + mch) of P
  const int tile_beg = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(p.ntiles, tile_beg + p.tiles_per_cta);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
    for (int i = 0; i < p.nstage; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(done), 1);
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
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)(p.RA * p.mch * 1024 + p.npl * p.Qs * 8 * p.PW * 128);
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int tt = tile;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y; tt /= p.tiles_y;
        const int n = tt, gy0 = ty * 8, gx0 = tx * 8;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&full[s]);
        mbar_expect_tx(fb, tx_bytes);
        const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
        // P rows gy0 - (ny - 1) .. gy0 + 7
        if (p.a5d) tma_load_5d(st, &mapP, fb, 0, gx0, mt * p.mch, gy0 - (p.ny - 1), n);
        else tma_load_4d(st, &mapP, fb, 0, gx0, gy0 - (p.ny - 1), n);
        for (int pl = 0; pl < p.npl; ++pl)
          for (int qs = 0; qs < p.Qs; ++qs) {
            const uint32_t dst = st + (uint32_t)p.a_bytes + (uint32_t)(pl * p.Qs + qs) * p.plane_bytes;
            if (p.s == 1) tma_load_4d(dst, &mapQ, fb, qs * 32, gx0 + p.pl_ox0[pl], gy0 + p.pl_oy0[pl], n);
            else tma_load_5d(dst, &mapQ, fb, p.packed ? 0 : p.pl_px[pl] * p.Cq + qs * 32, gx0 + p.pl_ox0[pl], p.pl_py[pl], gy0 + p.pl_oy0[pl], n);
          }
        if (++s == p.nstage) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D = f32, A = B = tf32, both MN-major, M = 128, N = nx * 32
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((p.nx * 32) >> 3) << 17) | ((128u >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    for (int tile = tile_beg; tile < tile_end; ++tile) {
      mbar_wait(smem_u32(&full[s]), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint64_t adesc = make_desc_mn(st, 1024u);
        const uint64_t bdesc = make_desc_mn(st + (uint32_t)p.a_bytes, 128u);
        const uint32_t first = tile == tile_beg ? 1u : 0u;
#pragma unroll 4
        for (int i = 0; i < p.nmma; ++i) {
          const uint4 e = tab.e[i];
          umma_tf32(tmem_base + e.z, adesc + (uint64_t)e.x, bdesc + (uint64_t)e.y, idesc, (first & e.w) ? 0u : 1u);
        }
        umma_commit(smem_u32(&empty[s]));
        if (tile == tile_end - 1) umma_commit(smem_u32(done));
      }
      __syncwarp();
      if (++s == p.nstage) { s = 0; ph ^= 1u; }
    }
  } else if (tile_end > tile_beg) {
    // ===================== epilogue (once): TMEM -> scratch[cta][cp][t*Cq + cq] =====================
    const int q = warp & 3;                                  // TMEM lane quarter = M atom (row shift, chunk)
    const int ip = q / p.mch, ch = q % p.mch;                // stacked row-shift slot i', chunk inside the M tile
    const int chunk = mt * p.mch + ch;
    const int cp = chunk * p.cstride + lane;
    const int ish = p.ny - 1 - ip;                           // plane-local row shift
    mbar_wait(smem_u32(done), 0);
    tc_fence_after();
    // overlapping chunks: a channel is written by the first chunk that holds it
    const bool row_ok = ip < p.ny && cp < p.Cp && chunk < p.nch && (chunk == 0 || lane >= 32 - p.cstride);
    const int Ntot = p.T * p.Cq;
    float* drow = scratch + (long long)blockIdx.x * p.split_stride + (long long)cp * Ntot;
    for (int pl = 0; pl < p.npl; ++pl)
      for (int qs = 0; qs < p.Qs; ++qs)
        for (int i2 = 0; i2 < p.nx; ++i2) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((pl * p.Qs + qs) * p.nx + i2) * 32), v);
          tmem_ld_wait();
          if (p.packed) {
            // columns = (px, 16 channels): each half belongs to a different tap (or to none)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int t = row_ok ? p.tap_of[pl][ish][i2][hf] : -1;
              if (t >= 0) {
                float* d = drow + t * p.Cq;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(v[hf * 16 + j]), __uint_as_float(v[hf * 16 + j + 1]),
                                                                  __uint_as_float(v[hf * 16 + j + 2]), __uint_as_float(v[hf * 16 + j + 3]));
              }
            }
            continue;
          }
          const int t = row_ok ? p.tap_of[pl][ish][i2][0] : -1;
          if (t >= 0) {
            float* d = drow + t * p.Cq + qs * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (qs * 32 + j < p.Cq)
                *reinterpret_cast<float4*>(d + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                                __uint_as_float(v[j + 3]));
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline int fdiv(int v, int s) { return v >= 0 ? v / s : -((-v + s - 1) / s); }
static inline int pow2_ge_w(int x) { int p = 1; while (p < x) p <<= 1; return p; }

struct WhCfg { WhParams p; WhTab tab; int S, mtiles; size_t smem; };

static bool wh_cfg(const WgradGeom& g, WhCfg& c) {
  if (!tc_encode_fn()) return false;
  WhParams& p = c.p;
  memset(&p, 0, sizeof(p));
  memset(&c.tab, 0, sizeof(c.tab));
  if (g.s != 1 && g.s != 2) return false;
  if (g.s == 2 && (g.Hq % 2 || g.Wq % 2)) return false;
  const bool packed = g.Cq == 16 && g.s == 2;
  // Q channels: 32-float chunks, the last one may run past Cq (into the other column parity or into TMA zero fill): those
  // accumulator columns are never written out.  P channels: chunks of 32 from the 5-D map, overlapping when Cp % 32 != 0.
  if (g.Cq % 4 || g.Cq < 8 || g.Cp % 4 || g.Cp < 8 || g.ntaps < 1 || g.ntaps > DSR_MAX_TAPS) return false;
  if (getenv("DCGANSR_HALO_EXACT_C") && ((g.Cq % 32 && !packed) || !(g.Cp % 32 == 0 || g.Cp <= 32))) return false;
  p.nch = (g.Cp + 31) / 32;
  p.cstride = 32;
  if (g.Cp > 32 && g.Cp % 32) {
    const int st = (g.Cp - 32) / (p.nch - 1);
    if ((g.Cp - 32) % (p.nch - 1) || st % 4) return false;
    p.cstride = st;
  }
  if (g.Hp < 8 || g.Wp < 8) return false;
  p.N = g.N; p.Hg = g.Hp; p.Wg = g.Wp; p.Cp = g.Cp; p.Cq = g.Cq; p.T = g.ntaps; p.s = g.s;
  // planes and their tap rectangles
  struct Pl { int py, px, oymin, oymax, oxmin, oxmax, cnt; };
  Pl pls[WH_MAXPL];
  int npl = 0;
  p.packed = packed ? 1 : 0;
  for (int t = 0; t < g.ntaps; ++t) {
    const int oy = fdiv(g.dy[t], g.s), py = g.dy[t] - oy * g.s, ox = fdiv(g.dx[t], g.s);
    const int px = packed ? 0 : g.dx[t] - ox * g.s;         // packed: both column parities share the plane
    int f = -1;
    for (int i = 0; i < npl; ++i) if (pls[i].py == py && pls[i].px == px) f = i;
    if (f < 0) { if (npl == WH_MAXPL) return false; f = npl++; pls[f] = Pl{py, px, oy, oy, ox, ox, 0}; }
    pls[f].oymin = std::min(pls[f].oymin, oy); pls[f].oymax = std::max(pls[f].oymax, oy);
    pls[f].oxmin = std::min(pls[f].oxmin, ox); pls[f].oxmax = std::max(pls[f].oxmax, ox);
    ++pls[f].cnt;
  }
  p.npl = npl;
  p.ny = pls[0].oymax - pls[0].oymin + 1; p.nx = pls[0].oxmax - pls[0].oxmin + 1;
  for (int i = 0; i < npl; ++i) {
    if (pls[i].oymax - pls[i].oymin + 1 != p.ny || pls[i].oxmax - pls[i].oxmin + 1 != p.nx) return false;   // equal rectangles only
    p.pl_py[i] = (short)pls[i].py; p.pl_px[i] = (short)pls[i].px; p.pl_oy0[i] = (short)pls[i].oymin; p.pl_ox0[i] = (short)pls[i].oxmin;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) { p.tap_of[i][a][b][0] = -1; p.tap_of[i][a][b][1] = -1; }
  }
  if (p.ny > 4 || p.nx > 4) return false;
  for (int t = 0; t < g.ntaps; ++t) {
    const int oy = fdiv(g.dy[t], g.s), py = g.dy[t] - oy * g.s, ox = fdiv(g.dx[t], g.s), pxr = g.dx[t] - ox * g.s;
    const int px = packed ? 0 : pxr;
    for (int i = 0; i < npl; ++i)
      if (pls[i].py == py && pls[i].px == px) p.tap_of[i][oy - pls[i].oymin][ox - pls[i].oxmin][packed ? pxr : 0] = (short)t;
  }
  p.Qs = packed ? 1 : (g.Cq + 31) / 32;
  p.a5d = g.Cp >= 32 ? 1 : 0;
  const int Qp = p.nch;
  p.mch = std::max(1, std::min(Qp, 4 / p.ny));           // M = ny * mch * 32 <= 128
  if (p.ny * p.mch > 4) return false;
  c.mtiles = (Qp + p.mch - 1) / p.mch;
  p.ncols = npl * p.Qs * p.nx * 32;
  if (p.ncols > 512 || p.nx * 32 > 256) return false;
  p.tmem_cols = std::max(32, pow2_ge_w(p.ncols));
  p.RA = 8 + p.ny - 1; p.PW = 8 + p.nx - 1;
  p.a_bytes = p.RA * p.mch * 1024;
  p.plane_bytes = (8 * p.PW * 128 + 1023) / 1024 * 1024;
  p.stage_bytes = p.a_bytes + npl * p.Qs * p.plane_bytes;
  p.nstage = std::min(8, (int)((232448 - 2048) / p.stage_bytes));
  if (p.nstage < 2) return false;
  p.nmma = npl * p.Qs * 8;
  if (p.nmma > WH_MAXMMA) return false;
  p.tiles_x = (p.Wg + 7) / 8;
  p.tiles_y = (p.Hg + p.ny - 1 + 7) / 8;                 // g' = g + (plane-local row shift) runs over Hg + ny - 1 rows
  p.ntiles = p.N * p.tiles_y * p.tiles_x;
  // MMA table: plane pl, chunk qs, 8-pixel row j of the K tile.  A window starts at P-tile row j (atom i' adds one row
  // per stacked shift); B window = plane row j
  int i = 0;
  for (int j = 0; j < 8; ++j)
    for (int pl = 0; pl < npl; ++pl)
      for (int qs = 0; qs < p.Qs; ++qs, ++i) {
        const uint32_t aoff = (uint32_t)(j * p.mch) * 1024u;
        const uint32_t boff = (uint32_t)(pl * p.Qs + qs) * p.plane_bytes + (uint32_t)(j * p.PW) * 128u;
        c.tab.e[i] = make_uint4(aoff >> 4, boff >> 4, (uint32_t)((pl * p.Qs + qs) * p.nx * 32), j == 0 ? 1u : 0u);
      }
  int S = std::max(1, NSM / c.mtiles);
  S = std::min(S, p.ntiles);
  p.tiles_per_cta = (p.ntiles + S - 1) / S;
  c.S = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.split_stride = (long long)g.Cp * g.ntaps * g.Cq;
  c.smem = 1024 + (size_t)p.nstage * p.stage_bytes + 32 * sizeof(uint64_t);
  return true;
}

// taken only where it pays: many pixels per weight (the per-tap kernel stays for the small discriminator layers)
bool wgrad_halo_supported(const WgradGeom& g) {
  WhCfg c;
  if (!wh_cfg(g, c)) return false;
  if (getenv("DCGANSR_NO_WGRAD_HALO")) return false;
  if (getenv("DCGANSR_HALO_ALL")) return true;
  return c.p.ntiles >= 4 * c.S;
}
size_t wgrad_halo_scratch_bytes(const WgradGeom& g) {
  WhCfg c;
  if (!wh_cfg(g, c)) return 0;
  return (size_t)c.S * g.Cp * g.ntaps * g.Cq * sizeof(float);
}

bool k_wgrad_halo(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                  std::string* err) {
  WhCfg c;
  if (!wh_cfg(g, c)) { if (err) *err = "wgrad geometry not supported by the halo kernel"; return false; }
  if ((size_t)c.S * g.Cp * g.ntaps * g.Cq * sizeof(float) > scratch_bytes) { if (err) *err = "wgrad scratch too small"; return false; }
  const WhParams& p = c.p;
  EncodeTiledFn enc = tc_encode_fn();
  CUtensorMap mapP, mapQ;
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r;
  if (p.a5d) {
    // [n][y][chunk][x][32]: the box lands as [row][chunk][8 pixels] so that M atoms are 1 KB apart
    cuuint64_t dims[5] = {32, (cuuint64_t)g.Wp, (cuuint64_t)p.nch, (cuuint64_t)g.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)g.Cp * 4, (cuuint64_t)p.cstride * 4, (cuuint64_t)g.Wp * g.Cp * 4, (cuuint64_t)g.Hp * g.Wp * g.Cp * 4};
    cuuint32_t box[5] = {32, 8, (cuuint32_t)p.mch, (cuuint32_t)p.RA, 1};
    r = enc(&mapP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)P, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cp, (cuuint64_t)g.Wp, (cuuint64_t)g.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cp * 4, (cuuint64_t)g.Wp * g.Cp * 4, (cuuint64_t)g.Hp * g.Wp * g.Cp * 4};
    cuuint32_t box[4] = {32, 8, (cuuint32_t)p.RA, 1};
    r = enc(&mapP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)P, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(P, wgrad halo) failed: " + std::to_string((int)r); return false; }
  if (g.s == 1) {
    cuuint64_t dims[4] = {(cuuint64_t)g.Cq, (cuuint64_t)g.Wq, (cuuint64_t)g.Hq, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.PW, 8, 1};
    r = enc(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)2 * g.Cq, (cuuint64_t)g.Wq / 2, 2, (cuuint64_t)g.Hq / 2, (cuuint64_t)g.N};
    cuuint64_t strides[4] = {(cuuint64_t)2 * g.Cq * 4, (cuuint64_t)g.Wq * g.Cq * 4, (cuuint64_t)2 * g.Wq * g.Cq * 4,
                             (cuuint64_t)g.Hq * g.Wq * g.Cq * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)p.PW, 1, 8, 1};
    r = enc(&mapQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)Q, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { if (err) *err = "cuTensorMapEncodeTiled(Q, wgrad halo) failed: " + std::to_string((int)r); return false; }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      if (err) *err = "cudaFuncSetAttribute(smem, wgrad halo) failed";
      return false;
    }
    configured = true;
  }
  dim3 grid((unsigned)c.S, (unsigned)c.mtiles);
  wgrad_halo_kernel<<<grid, WH_THREADS, c.smem, st.s>>>(mapP, mapQ, c.tab, p, scratch);
  // HBM-bound by construction (P and Q read once): algorithmic bytes = both activations + the gradient
  DSR_LAUNCHED(st, "wgrad_halo", 4.0 * ((double)g.N * g.Hp * g.Wp * g.Cp + (double)g.N * g.Hq * g.Wq * g.Cq + (double)g.Cp * g.Cq * g.ntaps),
               WORK_BYTES);
  k_wgrad_reduce(st, scratch, c.S, g.Cp, g.Cq, g.ntaps, grad_master);
  return true;
}
