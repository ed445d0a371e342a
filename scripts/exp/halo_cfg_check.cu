// Host-only consistency check of the halo kernel's configuration / MMA table (no GPU needed):
//   nvcc -std=c++17 -I<csrc> halo_cfg_check.cu -o halo_cfg_check && ./halo_cfg_check
#include "kernels_halo.cu"
EncodeTiledFn tc_encode_fn() { return (EncodeTiledFn)(void*)1; }
#include <map>
#include <set>

static void sub_pixel(std::vector<TapGeom>& v, int k, int s, int p, int Hi, int Wi, int Ci, int Ho, int Wo, int Co) {
  for (int ry = 0; ry < s; ++ry)
    for (int rx = 0; rx < s; ++rx) {
      TapGeom g; memset(&g, 0, sizeof(g));
      g.Hi = Hi; g.Wi = Wi; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.si = 1; g.so = s; g.oy0 = ry; g.ox0 = rx;
      g.Hg = Ho > ry ? (Ho - ry + s - 1) / s : 0; g.Wg = Wo > rx ? (Wo - rx + s - 1) / s : 0;
      int nt = 0;
      for (int ky = 0; ky < k; ++ky) { int vy = ry + p - ky; if (((vy % s) + s) % s) continue;
        for (int kx = 0; kx < k; ++kx) { int vx = rx + p - kx; if (((vx % s) + s) % s) continue; g.dy[nt] = vy / s; g.dx[nt] = vx / s; ++nt; } }
      g.ntaps = nt; g.N = 4; v.push_back(g);
    }
}
static void direct(std::vector<TapGeom>& v, int k, int s, int p, int Hi, int Wi, int Ci, int Ho, int Wo, int Co) {
  TapGeom g; memset(&g, 0, sizeof(g));
  g.Hi = Hi; g.Wi = Wi; g.Ci = Ci; g.Ho = Ho; g.Wo = Wo; g.Co = Co; g.si = s; g.so = 1; g.Hg = Ho; g.Wg = Wo; g.N = 4;
  int nt = 0;
  for (int ky = 0; ky < k; ++ky) for (int kx = 0; kx < k; ++kx) { g.dy[nt] = ky - p; g.dx[nt] = kx - p; ++nt; }
  g.ntaps = nt; v.push_back(g);
}

static int check(const char* name, std::vector<TapGeom>& cls) {
  HaloCfg c;
  if (!halo_cfg(cls.data(), (int)cls.size(), c)) { printf("%-28s unsupported\n", name); return 0; }
  HaloParams& p = c.p;
  int bad = 0;
  // per group: slots initialised before accumulated, column ranges of groups disjoint, plane order, window inside the plane
  std::set<int> allcols;
  for (int g = 0; g < p.ngrp; ++g) {
    std::set<int> written;
    int i = p.gbeg[g];
    for (int pl = 0; pl < p.nplanes; ++pl) {
      for (; i < p.pend[g][pl]; ++i) {
        uint4 e = c.tab.e[i];
        int col = e.z & 0xFFFF, N = ((e.w >> 17) & 0x3F) << 3;
        bool init = (int)e.z < 0;
        for (int k = col; k < col + N; ++k) {
          if (init) { if (written.count(k)) { printf("  %s: re-init of col %d\n", name, k); ++bad; } written.insert(k); }
          else if (!written.count(k)) { printf("  %s: accumulate into unwritten col %d (grp %d entry %d)\n", name, k, g, i); ++bad; }
        }
        size_t aoff = (size_t)e.x << 4;
        if ((aoff % p.row_bytes) + 32 > (size_t)p.row_bytes || aoff / p.row_bytes * p.row_bytes + (size_t)(HALO_TH - 1) * p.pitch_bytes + HALO_TW * p.row_bytes > (size_t)p.plane_tx) { printf("  %s: A window leaves the plane (aoff %zu)\n", name, aoff); ++bad; }
        size_t woff = (size_t)e.y << 4;
        if (woff + 32 > (size_t)p.w_bytes) { printf("  %s: W offset out of range\n", name); ++bad; }
      }
    }
    if (i != p.gbeg[g + 1]) { printf("  %s: group %d entries %d != %d\n", name, g, i, p.gbeg[g + 1]); ++bad; }
    for (int k : written) { if (allcols.count(k)) { printf("  %s: groups share col %d\n", name, k); ++bad; } allcols.insert(k); }
  }
  // every accumulator slot a class reads in the epilogue has been written
  for (int cl = 0; cl < p.ncls; ++cl)
    for (int j = 0; j < p.cls_nsl[cl]; ++j)
      for (int k = 0; k < p.Npad; ++k) if (!allcols.count(p.cls_sl[cl][j] * p.Npad + k)) { printf("  %s: class %d slot %d col %d never written\n", name, cl, j, k); ++bad; break; }
  printf("%-28s Npad %3d nsplit %d KBw %2d planes %d ring %2d (plane %5d B) w %6d B grp %d mma %3d acc %d x %3d tstore %d nbuf %d row %2d xor %d smem %6zu %s\n", name,
         p.Npad, c.nsplit, p.KBw, p.nplanes, p.nring, p.plane_bytes, p.w_bytes, p.ngrp, p.nmma, p.nacc, p.acc_cols, p.tstore, p.st_nbuf, p.st_row, p.st_xor, c.smem,
         bad ? "BAD" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  struct L { const char* name; bool full; int cin, cout, k, s, p, H; };
  L layers[] = {{"FC 64->32 (C2)", true, 64, 32, 4, 2, 1, 128}, {"C 32->16 (C2)", false, 32, 16, 4, 2, 1, 256},
                {"FC 96->48 (C3)", true, 96, 48, 4, 2, 1, 128}, {"FC 48->24 (C3)", true, 48, 24, 4, 2, 1, 256}, {"C 24->12 (C3)", false, 24, 12, 4, 2, 1, 512},
                {"FC 32->16 (C4)", true, 32, 16, 4, 2, 1, 64}, {"C 16->32 (C4)", false, 16, 32, 4, 2, 1, 128}, {"C 32->64 (C4)", false, 32, 64, 4, 2, 1, 64},
                {"FC 256->128 (C1b)", true, 256, 128, 4, 2, 1, 64}, {"FC 128->64 (C1b)", true, 128, 64, 4, 2, 1, 128}, {"C 64->128 (C1b)", false, 64, 128, 4, 2, 1, 256},
                {"C 128->256 (C1b)", false, 128, 256, 4, 2, 1, 128}, {"D C 64->128 s2", false, 64, 128, 4, 2, 1, 64}, {"conv3x3 32->24", false, 32, 24, 3, 1, 1, 40},
                {"patchD 64->128 k3", false, 64, 128, 3, 1, 0, 30}};
  for (auto& l : layers) {
    int Ho = l.full ? (l.H - 1) * l.s - 2 * l.p + l.k : (l.H + 2 * l.p - l.k) / l.s + 1;
    std::vector<TapGeom> fwd, dg;
    if (l.full) { sub_pixel(fwd, l.k, l.s, l.p, l.H, l.H, l.cin, Ho, Ho, l.cout); direct(dg, l.k, l.s, l.p, Ho, Ho, l.cout, l.H, l.H, l.cin); }
    else { direct(fwd, l.k, l.s, l.p, l.H, l.H, l.cin, Ho, Ho, l.cout); sub_pixel(dg, l.k, l.s, l.p, Ho, Ho, l.cout, l.H, l.H, l.cin); }
    char nm[64];
    snprintf(nm, sizeof nm, "%s fwd", l.name); bad += check(nm, fwd);
    snprintf(nm, sizeof nm, "%s dgrad", l.name); bad += check(nm, dg);
    if (fwd.size() == 4) { std::vector<TapGeom> one(fwd.begin(), fwd.begin() + 1); snprintf(nm, sizeof nm, "%s fwd/class", l.name); bad += check(nm, one); }
    if (dg.size() == 4) { std::vector<TapGeom> one(dg.begin() + 3, dg.begin() + 4); snprintf(nm, sizeof nm, "%s dgrad/class", l.name); bad += check(nm, one); }
  }
  printf(bad ? "FAILED (%d)\n" : "all ok\n", bad);
  return bad != 0;
}
