// dcgansr.cu -- context, convolution plans, the nn.Sequential executor, the fused training step
// and the C ABI of libdcgansr.so (include/dcgansr.h).
//
// What is restated from the reference (file:line under /root/reference):
//   netG / netD builders        train.lua:97-139, train-gray.lua:102-137, train-gray-patch.lua:54-109
//   Module:getParameters order  train.lua:202-203
//   fDx / fGx / loop order      train.lua:208-283   (stale-activation G step, two D passes per step)
//   optim.adam                  train.lua:280,283
// Design: activations NHWC fp32 in HBM, parameters / gradients / Adam state in one flat master
// vector per net in the Torch7 layout (export/import is a memcpy, Adam is one fused pass), packed
// per-tap weight copies refreshed after every update.  One CUDA stream per context; no host
// synchronisation inside a step unless the caller asks for the loss values.
#pragma GCC visibility push(default)
#include "../../include/dcgansr.h"
#pragma GCC visibility pop
#include "common.h"

#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string t_err;
#define NSM_WS 160      // split-K workspace: one 128 x 128 partial tile per resident CTA

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  const void *G = nullptr, *D = nullptr, *real = nullptr;
  int batch = 0;
  dcgansr_step_cfg cfg;
  int64_t launches = 0;
};

struct dcgansr_ctx {
  dcgansr_cfg cfg;
  cudaStream_t stream = nullptr, comm_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_c2m = nullptr, ev_m2c = nullptr;
  int64_t launches = 0;
  std::string err;
  float* flush_buf = nullptr;
  int64_t flush_count = 0;
  float* d_losses = nullptr;   // errD_real, errD_fake, errG, spare
  float* h_losses = nullptr;   // pinned mirror
  std::vector<float*> slots;   // staged batches (NHWC)
  std::vector<size_t> slot_bytes;
  float* tmp = nullptr;        // device scratch for layout import/export
  size_t tmp_bytes = 0;
  float* label_vec = nullptr;  // per-sample pixel MSE (train.lua:237-245)
  size_t label_cap = 0;
  float* lr_buf = nullptr;     // 2x2 box down-sampled batch (train.lua:225-230)
  size_t lr_cap = 0;
  float* fake_buf = nullptr;   // generator output of the whole batch when G runs micro-batched and D is not paired
  size_t fake_cap = 0;
  NcclApi nccl;
  ncclComm_t comm = nullptr;
  // CUDA graphs of the step (cfg.use_graph), one per (nets, staged batch, batch size, step cfg)
  std::vector<GraphEntry> graphs;
  bool tc_failed = false;
  cudaEvent_t graph_ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> bucket_ev;      // one per gradient bucket in flight (main -> comm stream hand-over)
  int buckets_in_flight = 0;               // all-reduces issued on comm_stream and not yet joined
  uint64_t graph_seq = 0;
  Prof prof;
  TcWorkspace tcws;
  std::vector<dcgansr_net*> nets;          // live nets created on this ctx (dcgansr_ctx_destroy detaches them)
  // one-shot all-reduce over NVLink peer memory (kernels_peer.cu): set up by dcgansr_comm_init when every rank can map every peer
  PeerAR peer{};
  bool peer_ok = false;
  void* peer_base = nullptr;
  std::vector<void*> peer_maps;
  int* peer_err_h = nullptr;
  St st() { return St{stream, &launches, &prof, &tcws}; }
  int world() const { return cfg.world_size > 1 && comm ? cfg.world_size : 1; }
};

// A captured step graph bakes in device pointers (net buffers, the staged batch, lr_buf, label_vec): every entry that could
// reference a buffer about to be freed is destroyed first.  net == nullptr: all of them.
static void graphs_invalidate(dcgansr_ctx* ctx, const void* net) {
  if (!ctx || ctx->graphs.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  size_t k = 0;
  for (size_t i = 0; i < ctx->graphs.size(); ++i) {
    GraphEntry& g = ctx->graphs[i];
    if (!net || g.G == net || g.D == net) { if (g.exec) cudaGraphExecDestroy(g.exec); }
    else ctx->graphs[k++] = g;
  }
  ctx->graphs.resize(k);
}

static int fail(dcgansr_ctx* ctx, int code, const std::string& msg) {
  t_err = msg;
  if (ctx) ctx->err = msg;
  return code;
}
#define CK(ctx, call)                                                                                     \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(ctx, DCGANSR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)
#define CKLAST(ctx)                                                                                       \
  do {                                                                                                    \
    if ((ctx) && (ctx)->tc_failed) { (ctx)->tc_failed = false; t_err = (ctx)->err; return DCGANSR_ERR_CUDA; } \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess)                                                                                \
      return fail(ctx, DCGANSR_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_));        \
  } while (0)
#define CKN(ctx, call)                                                                                    \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess)                                                                                \
      return fail(ctx, DCGANSR_ERR_NCCL, std::string(#call) + ": " + (ctx)->nccl.GetErrorString(r_));     \
  } while (0)

struct Arena {
  std::vector<void*> ptrs;
  ~Arena() { for (void* p : ptrs) cudaFree(p); }
  float* f(int64_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)std::max<int64_t>(n, 4) * sizeof(float)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return (float*)p;
  }
  void* bytes(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 16)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return p;
  }
};

static int ensure(dcgansr_ctx* ctx, float** buf, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return 0;
  if (*buf) { graphs_invalidate(ctx, nullptr); cudaStreamSynchronize(ctx->stream); cudaFree(*buf); *buf = nullptr; *cap = 0; }
  CK(ctx, cudaMalloc((void**)buf, bytes));
  *cap = bytes;
  return 0;
}

// host NCHW -> device NHWC (dst); uses ctx->tmp for the transpose source
static int upload_nchw(dcgansr_ctx* ctx, const float* host, int N, int C, int H, int W, float* dst) {
  size_t bytes = (size_t)N * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (C == 1 || H * W == 1) {
    CK(ctx, cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
  }
  if (int rc = ensure(ctx, &ctx->tmp, &ctx->tmp_bytes, bytes)) return rc;
  CK(ctx, cudaMemcpyAsync(ctx->tmp, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_nchw_to_nhwc(ctx->st(), ctx->tmp, dst, N, C, H, W);
  return 0;
}
// device NHWC (src) -> host NCHW; synchronises the stream
static int download_nchw(dcgansr_ctx* ctx, const float* src, int N, int C, int H, int W, float* host) {
  size_t bytes = (size_t)N * C * H * W * sizeof(float);
  if (bytes == 0) return 0;
  if (C == 1 || H * W == 1) {
    CK(ctx, cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    if (int rc = ensure(ctx, &ctx->tmp, &ctx->tmp_bytes, bytes)) return rc;
    k_nhwc_to_nchw(ctx->st(), src, ctx->tmp, N, C, H, W);
    CK(ctx, cudaMemcpyAsync(host, ctx->tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CKLAST(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------
// convolution plan: geometry classes + packed weights of one conv / full-conv module
// ------------------------------------------------------------------------------------------
struct TapClass {
  TapGeom g;
  std::vector<int> tapidx;
  int* tapidx_dev = nullptr;
  float* wp = nullptr;       // SIMT pack  [t][A][B]
  float* bp = nullptr;       // tensor-core pack (K-major, tf32)
  float* bt = nullptr;       // the same, pre-tiled + pre-swizzled for the per-tap kernel's bulk-copy loads (instead of bp)
  int A = 0, B = 0;
  int64_t sa = 0, sb = 0;
};

struct ConvPlan {
  bool full = false;
  int cin = 0, cout = 0, k = 0, s = 1, p = 0, adj = 0;
  int Hin = 0, Win = 0, Hout = 0, Wout = 0, T = 0;
  std::vector<TapClass> fwd, dgrad;
  int hmode_fwd = -1, hmode_dgrad = -1;      // halo_mode() of the two class groups, decided once in alloc_device (host cost per launch)
  WgradGeom wg;
  bool device = false;

  int64_t weight_count() const { return (int64_t)cin * cout * k * k; }

  // classes of the "transposed" form: out[gy*s+ry] = sum over ky with (ry+p-ky) % s == 0 of in[gy + (ry+p-ky)/s]
  static void sub_pixel_classes(std::vector<TapClass>& v, int k, int s, int p, int Hi, int Wi, int Ci, int Ho, int Wo, int Co) {
    for (int ry = 0; ry < s; ++ry)
      for (int rx = 0; rx < s; ++rx) {
        TapClass c;
        memset(&c.g, 0, sizeof(c.g));
        c.g.Hi = Hi; c.g.Wi = Wi; c.g.Ci = Ci; c.g.Ho = Ho; c.g.Wo = Wo; c.g.Co = Co;
        c.g.si = 1; c.g.so = s; c.g.oy0 = ry; c.g.ox0 = rx;
        c.g.Hg = Ho > ry ? (Ho - ry + s - 1) / s : 0;
        c.g.Wg = Wo > rx ? (Wo - rx + s - 1) / s : 0;
        int nt = 0;
        for (int ky = 0; ky < k; ++ky) {
          int vy = ry + p - ky;
          if (((vy % s) + s) % s != 0) continue;
          for (int kx = 0; kx < k; ++kx) {
            int vx = rx + p - kx;
            if (((vx % s) + s) % s != 0) continue;
            c.g.dy[nt] = vy / s; c.g.dx[nt] = vx / s;     // exact (divisible)
            c.tapidx.push_back(ky * k + kx);
            ++nt;
          }
        }
        c.g.ntaps = nt;
        v.push_back(c);
      }
  }
  static void direct_class(std::vector<TapClass>& v, int k, int s, int p, int Hi, int Wi, int Ci, int Ho, int Wo, int Co) {
    TapClass c;
    memset(&c.g, 0, sizeof(c.g));
    c.g.Hi = Hi; c.g.Wi = Wi; c.g.Ci = Ci; c.g.Ho = Ho; c.g.Wo = Wo; c.g.Co = Co;
    c.g.si = s; c.g.so = 1; c.g.oy0 = 0; c.g.ox0 = 0; c.g.Hg = Ho; c.g.Wg = Wo;
    int nt = 0;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        c.g.dy[nt] = ky - p; c.g.dx[nt] = kx - p;
        c.tapidx.push_back(ky * k + kx);
        ++nt;
      }
    c.g.ntaps = nt;
    v.push_back(c);
  }

  // returns "" or an error text
  std::string build(bool full_, int cin_, int cout_, int k_, int s_, int p_, int adj_, int Hin_, int Win_) {
    full = full_; cin = cin_; cout = cout_; k = k_; s = s_; p = p_; adj = adj_; Hin = Hin_; Win = Win_;
    T = k * k;
    if (cin <= 0 || cout <= 0 || k <= 0 || s <= 0 || p < 0) return "bad convolution parameters";
    if (T > DSR_MAX_TAPS) return "kernel larger than 5x5 is not supported";
    if (full) { Hout = (Hin - 1) * s - 2 * p + k + adj; Wout = (Win - 1) * s - 2 * p + k + adj; }
    else {
      if (Hin + 2 * p < k || Win + 2 * p < k) return "input smaller than kernel";
      Hout = (Hin + 2 * p - k) / s + 1; Wout = (Win + 2 * p - k) / s + 1;
    }
    if (Hout <= 0 || Wout <= 0) return "empty convolution output";
    const int64_t Ti = T;
    if (!full) {
      direct_class(fwd, k, s, p, Hin, Win, cin, Hout, Wout, cout);
      for (auto& c : fwd) { c.A = cin; c.B = cout; c.sa = Ti; c.sb = (int64_t)cin * Ti; }       // w[co][ci][t]
      sub_pixel_classes(dgrad, k, s, p, Hout, Wout, cout, Hin, Win, cin);
      for (auto& c : dgrad) { c.A = cout; c.B = cin; c.sa = (int64_t)cin * Ti; c.sb = Ti; }
      memset(&wg, 0, sizeof(wg));
      wg.Hp = Hout; wg.Wp = Wout; wg.Cp = cout; wg.Hq = Hin; wg.Wq = Win; wg.Cq = cin; wg.s = s; wg.ntaps = T;
    } else {
      sub_pixel_classes(fwd, k, s, p, Hin, Win, cin, Hout, Wout, cout);
      for (auto& c : fwd) { c.A = cin; c.B = cout; c.sa = (int64_t)cout * Ti; c.sb = Ti; }     // w[ci][co][t]
      direct_class(dgrad, k, s, p, Hout, Wout, cout, Hin, Win, cin);
      for (auto& c : dgrad) { c.A = cout; c.B = cin; c.sa = Ti; c.sb = (int64_t)cout * Ti; }
      memset(&wg, 0, sizeof(wg));
      wg.Hp = Hin; wg.Wp = Win; wg.Cp = cin; wg.Hq = Hout; wg.Wq = Wout; wg.Cq = cout; wg.s = s; wg.ntaps = T;
    }
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) { wg.dy[ky * k + kx] = ky - p; wg.dx[ky * k + kx] = kx - p; }
    return "";
  }

  // how the weights-resident halo kernel takes a class group: 2 = all sub-pixel classes in one launch, 1 = one launch per class
  // (the group's weights do not fit shared memory together, each class's do: the input is read once per class, still far less
  // than once per tap), 0 = not at all
  static int halo_mode(const TapGeom* gs, int n) {
    if (n < 1 || n > 4 || getenv("DCGANSR_NO_HALO")) return 0;
    if (halo_tapconv_supported(gs, n)) return 2;
    if (n == 1 || getenv("DCGANSR_NO_HALO_PER_CLASS")) return 0;
    // only for spatially large classes (>= 16 tiles of 16 x 8 per image): every CTA reloads its class's weights, which a small
    // discriminator-sized grid does not amortise (D conv 64->128 dgrad, 16 x 16 per class: 41 us against 27 us per-tap)
    for (int i = 0; i < n; ++i) {
      if (gs[i].Hg <= 0 || gs[i].Wg <= 0) continue;
      if ((int64_t)gs[i].Hg * gs[i].Wg < 16 * 128 && !getenv("DCGANSR_HALO_ALL")) return 0;
      if (!halo_tapconv_supported(gs + i, 1, false)) return 0;
    }
    return 1;
  }
  int alloc_device(dcgansr_ctx* ctx) {
    const bool fast = ctx->cfg.precision == DCGANSR_FAST_TF32;
    this->fast = fast;
    for (auto* v : {&fwd, &dgrad}) {
      // the halo kernel also takes contraction widths that are not a multiple of 8 (zero-filled K tail, e.g. the 12 channels
      // of train.lua's ngf = 12): such a class group gets the TF32 pack too, and is then always run by the halo kernel
      bool halo_grp = false;
      if (fast && !v->empty() && v->size() <= 4 && !getenv("DCGANSR_NO_HALO")) {
        TapGeom gs[4];
        bool wide = true;
        for (size_t i = 0; i < v->size(); ++i) { gs[i] = (*v)[i].g; gs[i].N = 1; wide = wide && (*v)[i].A > 4 && (*v)[i].B > 4; }
        const int hm = halo_mode(gs, (int)v->size());
        (v == &fwd ? hmode_fwd : hmode_dgrad) = hm;
        halo_grp = wide && hm != 0;
      }
      for (auto& c : *v) {
        if (fast) {
          TapGeom g = c.g;
          g.N = 1;
          if (tc_tapconv_supported(g) || halo_grp)
            CK(ctx, cudaMalloc((void**)&c.bp, std::max<size_t>(tc_packed_elems(c.g.ntaps, c.A, c.B), 4) * sizeof(float)));
        }
        size_t n = std::max<size_t>(c.tapidx.size(), 1);
        CK(ctx, cudaMalloc((void**)&c.tapidx_dev, n * sizeof(int)));
        if (!c.tapidx.empty())
          CK(ctx, cudaMemcpy(c.tapidx_dev, c.tapidx.data(), c.tapidx.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(ctx, cudaMalloc((void**)&c.wp, std::max<size_t>((size_t)c.g.ntaps * c.A * c.B, 4) * sizeof(float)));
      }
    }
    // class groups the weights-resident halo kernel will NOT take run on the per-tap kernel: give them the pre-tiled
    // weight images as well (decided once, here: it only depends on the geometry, not on the batch)
    if (fast && !getenv("DCGANSR_NO_BT"))
      for (auto* v : {&fwd, &dgrad}) {
        if (v->empty() || v->size() > 4) continue;
        TapGeom gs[4];
        bool all_tc = true;
        for (size_t i = 0; i < v->size(); ++i) { gs[i] = (*v)[i].g; gs[i].N = 1; all_tc = all_tc && (*v)[i].bp && (*v)[i].A > 4; }
        if (!all_tc || (*v)[0].A % 32 || !tc_tapconv_multi_ok(gs, (int)v->size()) || (v == &fwd ? hmode_fwd : hmode_dgrad) > 0) continue;
        for (auto& c : *v)
          CK(ctx, cudaMalloc((void**)&c.bt, std::max<size_t>(tc_bt_elems(c.g.ntaps, c.A, c.B), 4) * sizeof(float)));
      }
    device = true;
    return 0;
  }
  void free_device() {
    for (auto* v : {&fwd, &dgrad})
      for (auto& c : *v) {
        if (c.tapidx_dev) cudaFree(c.tapidx_dev);
        if (c.wp) cudaFree(c.wp);
        if (c.bp) cudaFree(c.bp);
        if (c.bt) cudaFree(c.bt);
        c.tapidx_dev = nullptr; c.wp = nullptr; c.bp = nullptr; c.bt = nullptr;
      }
    device = false;
  }
  // single-class repack of the pre-tiled images (layer-level ops; nets use the fused job list)
  static void pack_bt_one(St st, const float* master, TapClass& c) {
    PackJob j[2];
    j[0].src = master; j[0].dst = c.bt; j[0].tapidx = c.tapidx_dev; j[0].ntaps = c.g.ntaps; j[0].A = c.A; j[0].B = c.B; j[0].tc = 2;
    j[0].bn = tc_bt_rows(c.B); j[0].sa = c.sa; j[0].sb = c.sb; j[0].begin = 0;
    j[1] = j[0]; j[1].begin = (int64_t)tc_bt_elems(c.g.ntaps, c.A, c.B);
    PackJob* dj = nullptr;
    if (cudaMalloc((void**)&dj, sizeof(j)) != cudaSuccess) return;
    cudaMemcpyAsync(dj, j, sizeof(j), cudaMemcpyHostToDevice, st.s);
    k_pack_all(st, dj, 1, j[1].begin);
    cudaStreamSynchronize(st.s);
    cudaFree(dj);
  }
  void pack(St st, const float* master) {
    for (auto* v : {&fwd, &dgrad})
      for (auto& c : *v)
        if (c.g.ntaps > 0) {
          // the streaming kernels (1..4-channel side) read the [t][a][b] pack; a thin-OUTPUT class keeps both packs
          // (few pixels -> streaming reduction, many pixels -> tensor-core kernels)
          if (c.bp && c.A > 4) k_pack_taps_tc(st, master, c.bp, c.g.ntaps, c.tapidx_dev, c.A, c.B, c.sa, c.sb);
          if (c.bt) pack_bt_one(st, master, c);
          if (!c.bp || c.A <= 4 || c.B <= 4) k_pack_taps(st, master, c.wp, c.g.ntaps, c.tapidx_dev, c.A, c.B, c.sa, c.sb);
        }
  }
  void collect_jobs(std::vector<PackJob>& jobs, const float* master, int64_t& total) {
    for (auto* v : {&fwd, &dgrad})
      for (auto& c : *v)
        if (c.g.ntaps > 0) {
          PackJob j;
          j.src = master; j.tapidx = c.tapidx_dev; j.ntaps = c.g.ntaps; j.A = c.A; j.B = c.B; j.sa = c.sa; j.sb = c.sb;
          j.bn = 0;
          if (c.bt) {          // per-tap kernel: pre-tiled images only (the TMA 2-D pack is not read)
            j.tc = 2; j.dst = c.bt; j.bn = tc_bt_rows(c.B); j.begin = total;
            total += (int64_t)tc_bt_elems(c.g.ntaps, c.A, c.B);
            jobs.push_back(j);
          } else if (c.bp && c.A > 4) {
            j.tc = 1; j.dst = c.bp; j.begin = total;
            total += (int64_t)c.g.ntaps * c.A * c.B;
            jobs.push_back(j);
          }
          if (!c.bp || c.A <= 4 || c.B <= 4) {
            j.tc = 0; j.dst = c.wp; j.begin = total;
            total += (int64_t)c.g.ntaps * c.A * c.B;
            jobs.push_back(j);
          }
        }
  }
  // stats / stats_rows: see k_tapconv_halo; *stats_rows stays 0 unless the single-launch halo path ran with fused statistics
  static void run_classes(dcgansr_ctx* ctx, std::vector<TapClass>& v, int hmode, const float* in, float* out, int N, int act, float neg,
                          double* stats = nullptr, int* stats_rows = nullptr) {
    // 1..4-channel side: streaming fp32 kernels (both precisions)
    if (!v.empty() && v.size() <= 4 && !getenv("DCGANSR_NO_THIN")) {
      TapGeom gs[4];
      const float* wps[4];
      for (size_t i = 0; i < v.size(); ++i) { gs[i] = v[i].g; gs[i].N = N; wps[i] = v[i].wp; }
      if (thin_in_supported(gs, (int)v.size()) && k_tapconv_thin_in(ctx->st(), gs, (int)v.size(), wps, in, out, act, neg)) return;
      if (thin_out_supported(gs[0])) {
        for (size_t i = 0; i < v.size(); ++i)
          if (gs[i].Hg > 0 && gs[i].Wg > 0) k_tapconv_thin_out(ctx->st(), gs[i], in, wps[i], out, act, neg);
        return;
      }
      // many pixels, 1..4 output channels, and no tensor-core pack (strict mode, or a contraction width the TC kernels do not
      // take): one thread per output pixel instead of the generic FFMA tile kernel
      if (thin_out_px_supported(gs[0]) && !v[0].bp) {
        for (size_t i = 0; i < v.size(); ++i)
          if (gs[i].Hg > 0 && gs[i].Wg > 0) k_tapconv_thin_out_px(ctx->st(), gs[i], in, wps[i], out, act, neg);
        return;
      }
    }
    // spatially large thin layers: one weights-resident halo launch for all sub-pixel classes
    if (!v.empty() && v.size() <= 4 && !getenv("DCGANSR_NO_HALO")) {
      TapGeom gs[4];
      const float* bps[4];
      bool all_tc = true;
      for (size_t i = 0; i < v.size(); ++i) { gs[i] = v[i].g; gs[i].N = N; bps[i] = v[i].bp; all_tc = all_tc && v[i].bp && v[i].A > 4; }
      const int hm = (all_tc && !v[0].bt) ? (hmode >= 0 ? hmode : halo_mode(gs, (int)v.size())) : 0;    // (groups with pre-tiled images belong to the per-tap kernel)
      if (hm == 2) {
        std::string e;
        const int rows = (stats && stats_rows) ? halo_stats_rows(gs, (int)v.size()) : 0;
        if (k_tapconv_halo(ctx->st(), gs, (int)v.size(), bps, in, out, act, neg, &e, rows > 0 ? stats : nullptr, 0)) {
          if (rows > 0) *stats_rows = rows;
          return;
        }
        ctx->err = "tcgen05 halo path: " + e;
        ctx->tc_failed = true;
        return;
      }
      if (hm == 1) {
        for (size_t i = 0; i < v.size(); ++i) {
          if (gs[i].Hg <= 0 || gs[i].Wg <= 0) continue;
          std::string e;
          if (!k_tapconv_halo(ctx->st(), gs + i, 1, bps + i, in, out, act, neg, &e)) {
            ctx->err = "tcgen05 halo path: " + e;
            ctx->tc_failed = true;
            return;
          }
        }
        return;
      }
      const float* bts[4];
      for (size_t i = 0; i < v.size(); ++i) bts[i] = v[i].bt;
      if (all_tc && tc3_tapconv_supported(gs, (int)v.size(), bts)) {
        std::string e;
        if (k_tapconv_tc3(ctx->st(), gs, (int)v.size(), bts, in, out, act, neg, &e)) return;
        ctx->err = "tcgen05 halo-tile pair path: " + e;
        ctx->tc_failed = true;
        return;
      }
      if (all_tc && tc2_tapconv_supported(gs, (int)v.size(), bts)) {
        std::string e;
        if (k_tapconv_tc2(ctx->st(), gs, (int)v.size(), bts, in, out, act, neg, &e)) return;
        ctx->err = "tcgen05 wide-tile path: " + e;
        ctx->tc_failed = true;
        return;
      }
      if (all_tc && tc_tapconv_multi_ok(gs, (int)v.size())) {
        std::string e;
        if (k_tapconv_tc_multi(ctx->st(), gs, (int)v.size(), bps, in, out, act, neg, &e, bts)) return;
        ctx->err = "tcgen05 path: " + e;
        ctx->tc_failed = true;
        return;
      }
    }
    for (auto& c : v) {
      if (c.g.Hg <= 0 || c.g.Wg <= 0) continue;
      TapGeom g = c.g;
      g.N = N;
      if (c.bp && c.A > 4) {
        std::string e;
        if (k_tapconv_tc(ctx->st(), g, in, c.bp, out, act, neg, &e)) continue;
        ctx->err = "tcgen05 path: " + e;       // surfaced by the caller's CKLAST / status
        ctx->tc_failed = true;
        continue;
      }
      k_tapconv_simt(ctx->st(), g, in, c.wp, out, act, neg);
    }
  }
  void forward(dcgansr_ctx* ctx, const float* in, float* out, int N, int act, float neg, double* stats = nullptr, int* stats_rows = nullptr) {
    run_classes(ctx, fwd, hmode_fwd, in, out, N, act, neg, stats, stats_rows);
  }
  void dgrad_run(dcgansr_ctx* ctx, const float* dy, float* dx, int N) { run_classes(ctx, dgrad, hmode_dgrad, dy, dx, N, ACT_NONE, 0.f); }
  bool fast = false;       // FAST_TF32: tensor-core wgrad when the geometry allows
  size_t wscratch_bytes(int N) const {
    WgradGeom g = wg;
    g.N = N;
    size_t b = wgrad_simt_scratch_bytes(g);
    b = std::max(b, thin_wgrad_scratch_bytes(g));
    if (fast && tc_wgrad_supported(g)) b = std::max(b, wgrad_tc_scratch_bytes(g));
    if (fast) b = std::max(b, wgrad_tc_pair_scratch_bytes(g));
    if (fast) b = std::max(b, wgrad_halo_scratch_bytes(g));
    return b;
  }
  // x: module input, dy: gradient w.r.t. module output
  void wgrad_run(dcgansr_ctx* ctx, const float* x, const float* dy, float* grad_master, int N, float* scratch, size_t scratch_bytes) {
    WgradGeom g = wg;
    g.N = N;
    const float* Pp = full ? x : dy;
    const float* Qp = full ? dy : x;
    if (k_wgrad_thin(ctx->st(), g, Pp, Qp, grad_master, scratch, scratch_bytes)) return;   // 1..4-channel side: streaming kernel
    if (fast && wgrad_halo_supported(g)) {
      std::string e;
      if (k_wgrad_halo(ctx->st(), g, Pp, Qp, grad_master, scratch, scratch_bytes, &e)) return;
      ctx->err = "tcgen05 halo wgrad: " + e;
      ctx->tc_failed = true;
      return;
    }
    if (fast && wgrad_tc_pair_supported(g)) {
      std::string e;
      if (k_wgrad_tc_pair(ctx->st(), g, Pp, Qp, grad_master, scratch, scratch_bytes, &e)) return;
      ctx->err = "tcgen05 pair wgrad: " + e;
      ctx->tc_failed = true;
      return;
    }
    if (fast && tc_wgrad_supported(g)) {
      std::string e;
      if (k_wgrad_tc(ctx->st(), g, Pp, Qp, grad_master, scratch, scratch_bytes, &e)) return;
      ctx->err = "tcgen05 wgrad: " + e;
      ctx->tc_failed = true;
      return;
    }
    k_wgrad_simt(ctx->st(), g, Pp, Qp, grad_master, scratch, scratch_bytes);
  }
};

// ------------------------------------------------------------------------------------------
// nn.Sequential
// ------------------------------------------------------------------------------------------
struct Mod {
  int kind = 0;
  dcgansr_layer L;
  int cin = 0, hin = 0, win = 0, cout = 0, hout = 0, wout = 0;
  ConvPlan* conv = nullptr;
  int64_t p_off = -1, p_cnt = 0;
  int64_t bn_off = -1;
  int bucket = -1;               // conv modules: index of the gradient bucket [this conv, following BN ...)
  float* out = nullptr;
  bool owns_out = false;
  int fused_act = ACT_NONE;      // activation fused into this CONV / BN module
  float fused_neg = 0.f;
  bool fused_into_prev = false;  // ACT module executed by its producer
  int act = ACT_NONE;            // ACT modules: own kind
  float *save_mean = nullptr, *save_invstd = nullptr;
  int stats_rows = 0;            // BN modules: rows of net->bn_partials the producing convolution's epilogue has filled (0: none)
};

struct dcgansr_net {
  dcgansr_ctx* ctx = nullptr;
  std::vector<Mod> mods;
  int in_c = 0, in_h = 0, in_w = 0, max_batch = 0;
  int64_t nparams = 0, nbn = 0;
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  // data parallel: [stage | grads] is ONE allocation.  stage = (running_mean, running_var) / world [+ the 3 step losses], packed right
  // before the backward walk, so that the all-reduce of the walk's LAST gradient bucket (the first conv's, at offset 0) carries
  // them too: one collective instead of five latency-bound ones per step
  float* grads_base = nullptr;
  int64_t stage_len = 0;
  bool stage_live = false;                            // the stage holds packed values waiting for their all-reduce
  int64_t* adam_t = nullptr;
  float* adam_step = nullptr;
  float *bn_rmean = nullptr, *bn_rvar = nullptr, *bn_save = nullptr;
  float* in_buf = nullptr;
  float* gbuf[2] = {nullptr, nullptr};
  int64_t gelems = 0;
  double *bn_partials = nullptr, *bn_sums = nullptr, *bn_sums_total = nullptr;
  double *mb_fsum = nullptr, *mb_bsum = nullptr;       // micro-batched execution: whole-batch BN sums (forward / backward), 2 x nbn each
  // micro-batched execution: the output of ONE convolution kept for every micro-batch of the step (ck_B samples), so that the
  // re-forwards of the later BatchNorm passes start behind it instead of at the net input (mb_ensure_ckpt picks the module)
  int ck_mod = -1, ck_B = 0;
  float *ck_buf = nullptr, *ck_home = nullptr;         // ck_home: the module's own (micro-batch sized) output buffer
  int64_t ck_elems = 0;                                // floats per micro-batch
  bool ck_valid = false;
  float* bn_fmeans = nullptr;                         // (float)(sum / n) of the BN backward reductions, 2 groups x 2C
  float* wscratch = nullptr;
  size_t wscratch_bytes = 0;
  int first_param_mod = -1;
  int last_batch = 0;
  const float* last_out = nullptr;
  // parameter version at the last update / at the cached forward: while they agree, the BatchNorm backward may re-derive the
  // activation mask from x and the current gamma / beta instead of reading the cached output (kernels_bw.cu:bn_gval)
  uint64_t params_ver = 1, fwd_ver = 0;
  int out_c = 0, out_h = 0, out_w = 0;
  std::vector<std::pair<int64_t, int64_t>> buckets;   // gradient buckets (offset, count), forward order
  PackJob* pack_jobs = nullptr;                       // fused weight repack (one launch per net)
  int n_pack_jobs = 0;
  int64_t pack_total = 0;

  float* own(const float* cur) { return (cur == gbuf[0] || cur == gbuf[1]) ? const_cast<float*>(cur) : gbuf[0]; }
  float* other(const float* cur) { return cur == gbuf[0] ? gbuf[1] : gbuf[0]; }
};

// Frees everything the net holds on the device and detaches it from its context (the handle stays valid as a plan-only net).
static void net_release_device(dcgansr_net* net) {
  if (!net || !net->ctx) return;
  dcgansr_ctx* c = net->ctx;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  graphs_invalidate(c, net);          // a re-created net often gets the same heap / device addresses: no stale replay
  c->nets.erase(std::remove(c->nets.begin(), c->nets.end(), net), c->nets.end());
  if (net->ck_mod >= 0 && net->ck_buf) net->mods[net->ck_mod].out = net->ck_home;       // micro-batch checkpoint: back to the module's own buffer
  if (net->ck_buf) cudaFree(net->ck_buf);
  net->ck_buf = nullptr; net->ck_mod = -1; net->ck_B = 0; net->ck_valid = false;
  for (auto& m : net->mods) {
    if (m.conv) m.conv->free_device();
    if (m.owns_out && m.out) cudaFree(m.out);
    m.out = nullptr; m.owns_out = false; m.save_mean = nullptr; m.save_invstd = nullptr;
  }
  void** ptrs[] = {(void**)&net->params, (void**)&net->grads_base, (void**)&net->adam_m, (void**)&net->adam_v, (void**)&net->adam_t,
                   (void**)&net->adam_step, (void**)&net->bn_rmean, (void**)&net->bn_rvar, (void**)&net->bn_save, (void**)&net->in_buf,
                   (void**)&net->gbuf[0], (void**)&net->gbuf[1], (void**)&net->bn_partials, (void**)&net->bn_sums,
                   (void**)&net->bn_sums_total, (void**)&net->mb_fsum, (void**)&net->mb_bsum, (void**)&net->bn_fmeans, (void**)&net->wscratch, (void**)&net->pack_jobs};
  for (void** p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
  net->grads = nullptr;
  net->last_out = nullptr; net->last_batch = 0;
  net->ctx = nullptr;
}

static int act_of_kind(int kind) {
  switch (kind) {
    case DCGANSR_RELU: return ACT_RELU;
    case DCGANSR_LRELU: return ACT_LRELU;
    case DCGANSR_TANH: return ACT_TANH;
    case DCGANSR_SIGMOID: return ACT_SIGMOID;
    default: return ACT_NONE;
  }
}

static int nccl_allreduce(dcgansr_ctx* ctx, void* buf, size_t count, ncclDataType_t dt, cudaStream_t s) {
  if (ctx->world() <= 1) return 0;
  CKN(ctx, ctx->nccl.AllReduce(buf, buf, count, dt, ncclSum, ctx->comm, s));
  ++ctx->launches;
  return 0;
}

static void net_pack_all(dcgansr_net* net) {
  if (net->pack_jobs) { k_pack_all(net->ctx->st(), net->pack_jobs, net->n_pack_jobs, net->pack_total); return; }
  for (auto& m : net->mods)
    if (m.conv) m.conv->pack(net->ctx->st(), net->params + m.p_off);
}

// forward on device buffers (in: NHWC).  Caches every module output: last forward wins.
// groups > 1: the batch holds `groups` independent minibatches of B samples each (the fused step runs D(real) and D(fake)
// as ONE pass over 2B samples): convolutions see groups*B samples, every BatchNorm normalises each group with its own
// statistics and updates the running statistics group after group -- exactly what `groups` separate forwards would do.
// stop >= 0: run the modules [0, stop) only and return the input of module `stop` in *stop_in (micro-batched statistics passes).
// bn_frozen: BatchNorm modules normalise with the statistics already in save_mean / save_invstd (whole-batch statistics of a
// micro-batched step) and leave the running statistics alone.
static int net_forward_dev(dcgansr_net* net, const float* in, int B, int groups = 1, int stop = -1, bool bn_frozen = false,
                           const float** stop_in = nullptr, int start = 0) {
  dcgansr_ctx* ctx = net->ctx;
  St st = ctx->st();
  // start > 0: module start - 1 already holds its output for this batch (micro-batch checkpoint)
  const float* cur = start > 0 ? net->mods[start - 1].out : in;
  const bool sync = ctx->cfg.sync_bn && ctx->world() > 1;
  const int NB = B * groups;
  for (size_t mi = (size_t)start; mi < net->mods.size(); ++mi) {
    Mod& m = net->mods[mi];
    if ((int)mi == stop) { if (stop_in) *stop_in = cur; CKLAST(ctx); return 0; }
    switch (m.kind) {
      case DCGANSR_UPNEAREST:
        k_upnearest_fwd(st, cur, m.out, NB, m.hin, m.win, m.cin, m.L.scale);
        cur = m.out;
        break;
      case DCGANSR_CONV:
      case DCGANSR_FULLCONV: {
        // a BatchNorm right behind a halo-kernel convolution gets its batch sums from that kernel's epilogue (one read of the
        // tensor less); single sample group, local statistics, statistics actually wanted (not a frozen micro-batch pass)
        Mod* nb = mi + 1 < net->mods.size() && net->mods[mi + 1].kind == DCGANSR_BN ? &net->mods[mi + 1] : nullptr;
        int rows = 0;
        if (nb && groups == 1 && !sync && !bn_frozen && (int)mi + 1 != stop && m.fused_act == ACT_NONE)
          m.conv->forward(ctx, cur, m.out, NB, m.fused_act, m.fused_neg, net->bn_partials, &rows);
        else
          m.conv->forward(ctx, cur, m.out, NB, m.fused_act, m.fused_neg);
        if (nb) nb->stats_rows = rows;
        cur = m.out;
        break;
      }
      case DCGANSR_BN: {
        int64_t P = (int64_t)B * m.hin * m.win;
        int C = m.cin;
        if (bn_frozen) {
          k_bn_apply_act(st, cur, m.out, P * groups, C, net->params + m.p_off, net->params + m.p_off + C, m.save_mean, m.save_invstd,
                         m.fused_act, m.fused_neg);
          cur = m.out;
          break;
        }
        if (!sync && m.stats_rows > 0 && groups == 1) {
          k_bn_fwd_from_partials(st, cur, m.out, P, C, m.stats_rows, net->params + m.p_off, net->params + m.p_off + C, m.save_mean, m.save_invstd,
                                 net->bn_rmean + m.bn_off, net->bn_rvar + m.bn_off, m.L.eps, m.L.momentum, m.fused_act, m.fused_neg,
                                 net->bn_partials);
          m.stats_rows = 0;
          cur = m.out;
          break;
        }
        if (!sync) {
          k_bn_fwd_grouped(st, cur, m.out, P, C, groups, net->params + m.p_off, net->params + m.p_off + C, m.save_mean, m.save_invstd,
                           2 * net->nbn, net->bn_rmean + m.bn_off, net->bn_rvar + m.bn_off, m.L.eps, m.L.momentum, m.fused_act, m.fused_neg,
                           net->bn_partials);
          cur = m.out;
          break;
        }
        for (int g = 0; g < groups; ++g) {
          const float* xg = cur + (int64_t)g * P * C;
          float* smean = m.save_mean + (int64_t)g * 2 * net->nbn;
          float* sinv = m.save_invstd + (int64_t)g * 2 * net->nbn;
          k_bn_stats(st, xg, P, C, net->bn_partials, net->bn_sums);
          double n_total = (double)P;
          if (sync) n_total *= ctx->world();
          if (sync && ctx->peer_ok && 2 * C <= ctx->peer.nmax) {
            // the statistics tail does the cross-rank exchange itself (pushes over NVLink, kernels_peer.cu)
            k_bn_finalize_peer(st, ctx->peer, net->bn_sums, C, n_total, m.L.eps, m.L.momentum, smean, sinv, net->bn_rmean + m.bn_off,
                               net->bn_rvar + m.bn_off);
          } else {
            if (sync)
              if (int rc = nccl_allreduce(ctx, net->bn_sums, 2 * C, ncclDouble, ctx->stream)) return rc;
            k_bn_finalize(st, net->bn_sums, C, n_total, m.L.eps, m.L.momentum, smean, sinv, net->bn_rmean + m.bn_off, net->bn_rvar + m.bn_off);
          }
          k_bn_apply_act(st, xg, m.out + (int64_t)g * P * C, P, C, net->params + m.p_off, net->params + m.p_off + C, smean, sinv,
                         m.fused_act, m.fused_neg);
        }
        cur = m.out;
        break;
      }
      case DCGANSR_RELU: case DCGANSR_LRELU: case DCGANSR_TANH: case DCGANSR_SIGMOID:
        if (!m.fused_into_prev) {
          k_act(st, cur, m.out, (int64_t)NB * m.cin * m.hin * m.win, m.act, m.L.negval);
          cur = m.out;
        }
        break;
      default: break;   // VIEW
    }
  }
  net->last_out = cur;
  net->last_batch = NB;
  net->fwd_ver = net->params_ver;
  CKLAST(ctx);
  return 0;
}

// Gradient buckets overlapped with the backward walk (data parallel): as soon as a convolution's wgrad (and the BatchNorm
// that follows it in the net, already visited by the backward walk) has produced its slice of the flat gradient vector,
// that slice is all-reduced on the communication stream while the main stream goes on with the next dgrad / wgrad.
static int bucket_allreduce_async(dcgansr_ctx* ctx, dcgansr_net* net, int bucket) {
  if (bucket < 0 || bucket >= (int)net->buckets.size() || net->buckets[bucket].second <= 0) return 0;
  const int k = ctx->buckets_in_flight;
  while ((int)ctx->bucket_ev.size() <= k) {
    cudaEvent_t e;
    CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->bucket_ev.push_back(e);
  }
  CK(ctx, cudaEventRecord(ctx->bucket_ev[k], ctx->stream));
  CK(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->bucket_ev[k], 0));
  float* ptr = net->grads + net->buckets[bucket].first;
  int64_t cnt = net->buckets[bucket].second;
  if (bucket == 0 && net->stage_live && net->buckets[0].first == 0) { ptr -= net->stage_len; cnt += net->stage_len; net->stage_live = false; }
  if (int rc = nccl_allreduce(ctx, ptr, cnt, ncclFloat, ctx->comm_stream)) return rc;
  ++ctx->buckets_in_flight;
  return 0;
}
// main stream waits for every bucket issued so far
static int bucket_join(dcgansr_ctx* ctx) {
  if (ctx->buckets_in_flight == 0) return 0;
  CK(ctx, cudaEventRecord(ctx->ev_c2m, ctx->comm_stream));
  CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_c2m, 0));
  ctx->buckets_in_flight = 0;
  return 0;
}

// backward walk.  acc: accumulate parameter gradients (net:backward) or not (net:updateGradInput).
// Returns the gradient w.r.t. the net input in *dx_out (nullptr when need_dx is false).
// groups / goff: the cached activations hold several B-sample groups (see net_forward_dev); this walk covers groups
// [goff, goff + groups): `in` and `dy` point at the first sample of group goff, cached module outputs are offset likewise.
// Micro-batched walk (mb != nullptr): BatchNorm modules use the whole-batch sums in net->mb_bsum (n = b_total * H * W), exactly
// like the sync_bn path uses the all-reduced ones.  stop_mod >= 0: the walk ends at that BatchNorm module after adding this
// micro-batch's (sum g, sum g*xhat) to its whole-batch sums.  param_grads: add dgamma / dbeta (once per step, from the totals).
struct MbWalk { int b_total; int stop_mod; bool param_grads; };
static int net_backward_dev(dcgansr_net* net, const float* in, const float* dy, int B, bool acc, bool need_dx,
                            const float** dx_out, int groups = 1, int goff = 0, bool reduce_buckets = false, const MbWalk* mb = nullptr) {
  dcgansr_ctx* ctx = net->ctx;
  St st = ctx->st();
  const bool sync = ctx->cfg.sync_bn && ctx->world() > 1;
  const float* cur = dy;
  const int NB = B * groups;
  if (dx_out) *dx_out = nullptr;
  for (int i = (int)net->mods.size() - 1; i >= 0; --i) {
    Mod& m = net->mods[i];
    const float* inp = i > 0 ? net->mods[i - 1].out + (int64_t)goff * B * m.cin * m.hin * m.win : in;
    const float* mout = m.out ? m.out + (int64_t)goff * B * m.cout * m.hout * m.wout : nullptr;
    switch (m.kind) {
      case DCGANSR_UPNEAREST: {
        float* t = net->other(cur);
        k_upnearest_bwd(st, cur, t, NB, m.hin, m.win, m.cin, m.L.scale);
        cur = t;
        break;
      }
      case DCGANSR_CONV:
      case DCGANSR_FULLCONV: {
        if (m.fused_act != ACT_NONE) {
          float* t = net->own(cur);
          k_act_bwd(st, mout, cur, t, (int64_t)NB * m.cout * m.hout * m.wout, m.fused_act, m.fused_neg);
          cur = t;
        }
        if (acc) m.conv->wgrad_run(ctx, inp, cur, net->grads + m.p_off, NB, net->wscratch, net->wscratch_bytes);
        if (acc && reduce_buckets && m.bucket >= 0)
          if (int rc = bucket_allreduce_async(ctx, net, m.bucket)) return rc;
        if (i == net->first_param_mod && !need_dx) { CKLAST(ctx); return 0; }
        float* t = net->other(cur);
        m.conv->dgrad_run(ctx, cur, t, NB);
        cur = t;
        break;
      }
      case DCGANSR_BN: {
        int64_t P = (int64_t)B * m.hin * m.win;
        int C = m.cin;
        float* gall = net->own(cur);
        const float* gamma = net->params + m.p_off;
        const float* beta = gamma + C;
        // the activation mask: re-derived from x while the parameters are those of the cached forward (one tensor less to
        // read), else from the cached output (fGx walks D with post-Adam weights and pre-Adam activations, train.lua:264-270)
        const bool remask = net->fwd_ver == net->params_ver && (m.fused_act == ACT_RELU || m.fused_act == ACT_LRELU) && !getenv("DCGANSR_BN_READ_Y");
        const float* yact = (m.fused_act != ACT_NONE && !remask) ? mout : nullptr;
        if (mb) {
          double* tot = net->mb_bsum + 2 * m.bn_off;
          if (i == mb->stop_mod) {
            k_bn_bwd_reduce(st, cur, yact, inp, P, C, gamma, beta, m.save_mean, m.save_invstd, m.fused_act, m.fused_neg, net->bn_partials, net->bn_sums);
            k_dacc(st, tot, net->bn_sums, 2 * C);
            CKLAST(ctx);
            return 0;
          }
          if (acc && mb->param_grads) k_bn_bwd_param(st, tot, C, net->grads + m.p_off, net->grads + m.p_off + C);
          k_bn_bwd_apply(st, cur, yact, inp, gall, P, C, gamma, beta, m.save_mean, m.save_invstd, m.fused_act, m.fused_neg, tot,
                         (double)mb->b_total * m.hin * m.win, net->bn_fmeans);
          cur = gall;
          break;
        }
        if (!sync) {
          k_bn_bwd_grouped(st, cur, yact, inp, gall, P, C, groups, gamma, beta,
                           m.save_mean + (int64_t)goff * 2 * net->nbn, m.save_invstd + (int64_t)goff * 2 * net->nbn, 2 * net->nbn,
                           m.fused_act, m.fused_neg, net->bn_partials, net->bn_sums, acc ? net->grads + m.p_off : nullptr,
                           acc ? net->grads + m.p_off + C : nullptr, net->bn_fmeans);
          cur = gall;
          break;
        }
        for (int gi = 0; gi < groups; ++gi) {
          const int64_t off = (int64_t)gi * P * C;
          const float* smean = m.save_mean + (int64_t)(goff + gi) * 2 * net->nbn;
          const float* sinv = m.save_invstd + (int64_t)(goff + gi) * 2 * net->nbn;
          k_bn_bwd_reduce(st, cur + off, yact ? yact + off : nullptr, inp + off, P, C, gamma, beta, smean, sinv, m.fused_act, m.fused_neg,
                          net->bn_partials, net->bn_sums);
          if (acc) k_bn_bwd_param(st, net->bn_sums, C, net->grads + m.p_off, net->grads + m.p_off + C);
          const double* tot = net->bn_sums;
          double n_total = (double)P;
          if (sync) {
            if (ctx->peer_ok && 2 * C <= ctx->peer.nmax) {
              k_peer_allreduce(st, ctx->peer, net->bn_sums, net->bn_sums_total, 2 * C);
            } else {
              CK(ctx, cudaMemcpyAsync(net->bn_sums_total, net->bn_sums, 2 * C * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
              if (int rc = nccl_allreduce(ctx, net->bn_sums_total, 2 * C, ncclDouble, ctx->stream)) return rc;
            }
            tot = net->bn_sums_total;
            n_total *= ctx->world();
          }
          k_bn_bwd_apply(st, cur + off, yact ? yact + off : nullptr, inp + off, gall + off, P, C, gamma, beta, smean, sinv, m.fused_act,
                         m.fused_neg, tot, n_total, net->bn_fmeans);
        }
        cur = gall;
        break;
      }
      case DCGANSR_RELU: case DCGANSR_LRELU: case DCGANSR_TANH: case DCGANSR_SIGMOID:
        if (!m.fused_into_prev) {
          float* t = net->own(cur);
          k_act_bwd(st, mout, cur, t, (int64_t)NB * m.cin * m.hin * m.win, m.act, m.L.negval);
          cur = t;
        }
        break;
      default: break;   // VIEW
    }
  }
  if (dx_out) *dx_out = cur;
  CKLAST(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------
// Micro-batched execution with EXACT whole-batch BatchNorm (BASELINE config C5: the activations of 128 samples of the
// ngf = 128 generator do not fit 180 GB).  A net created for max_batch = b runs a batch B = k*b as k micro-batches; what
// couples the samples of a batch is only BatchNorm, whose batch sums are accumulated over the micro-batches before they are
// used -- the same structure as sync_bn, with time in the place of ranks.  Activations are recomputed instead of kept:
//   forward : for every BatchNorm j in turn, all micro-batches run up to its input (earlier BatchNorms frozen at their final
//             statistics) and add (sum x, sum x^2) to its whole-batch sums; then one more pass produces the outputs;
//   backward: for every BatchNorm j from the last to the first, all micro-batches are re-forwarded and walked back down to j
//             (later BatchNorms use their whole-batch sums) adding (sum g, sum g*xhat); the final pass walks all the way,
//             accumulating the parameter gradients.
// Cost: about (J + 1) forward and backward passes for J BatchNorms; result = the single-batch step up to summation order.
// ------------------------------------------------------------------------------------------
// Points the checkpointed module's output (and the fused modules that alias it) at micro-batch mi's slice, or back home (mi < 0)
static void mb_select(dcgansr_net* net, int mi) {
  if (net->ck_mod < 0 || !net->ck_buf) return;
  net->mods[net->ck_mod].out = mi < 0 ? net->ck_home : net->ck_buf + (int64_t)mi * net->ck_elems;
  for (size_t i = net->ck_mod + 1; i < net->mods.size() && !net->mods[i].owns_out; ++i) net->mods[i].out = net->mods[i - 1].out;
}

// Chooses and allocates the checkpoint for a step of B samples (never inside a capture: called by step_run before it).  A
// checkpoint at convolution i serves every pass that runs past i: the candidate with the largest (flops up to i) x 2 x (BatchNorms
// behind i) whose B-sample output fits the free memory is taken.  C5: FC 1024->512's output (64 GiB) -- the re-forwards of
// the BN3 / BN4 / final passes and of three of the five backward passes start behind the two first layers.
static void mb_ensure_ckpt(dcgansr_ctx* ctx, dcgansr_net* net, int B) {
  if (net->ck_B == B || B <= net->max_batch) return;
  cudaStreamSynchronize(ctx->stream);
  graphs_invalidate(ctx, net);
  mb_select(net, -1);
  if (net->ck_buf) cudaFree(net->ck_buf);
  net->ck_buf = nullptr; net->ck_mod = -1; net->ck_B = B; net->ck_valid = false;
  if (getenv("DCGANSR_NO_MB_CKPT")) return;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return; }
  const size_t margin = (size_t)4 << 30;
  double prefix = 0.0, best = 0.0;
  int best_i = -1;
  for (size_t i = 0; i + 1 < net->mods.size(); ++i) {
    Mod& m = net->mods[i];
    if (!m.conv) continue;
    prefix += 2.0 * m.cin * m.cout * m.L.kh * m.L.kw * (m.kind == DCGANSR_FULLCONV ? (double)m.hin * m.win : (double)m.hout * m.wout);
    int bn_after = 0;
    for (size_t j = i + 1; j < net->mods.size(); ++j) bn_after += net->mods[j].kind == DCGANSR_BN;
    if (!m.owns_out || bn_after == 0) continue;
    const size_t bytes = (size_t)B * m.cout * m.hout * m.wout * sizeof(float);
    if (bytes + margin > free_b) continue;
    const double score = prefix * 2.0 * bn_after;      // bn_after forward passes (the later statistics passes + the final one) and as many backward passes start behind it
    if (score > best) { best = score; best_i = (int)i; }
  }
  if (best_i < 0) return;
  Mod& m = net->mods[best_i];
  net->ck_elems = (int64_t)net->max_batch * m.cout * m.hout * m.wout;
  if (cudaMalloc((void**)&net->ck_buf, (size_t)B * m.cout * m.hout * m.wout * sizeof(float)) != cudaSuccess) { cudaGetLastError(); net->ck_buf = nullptr; return; }
  net->ck_mod = best_i;
  net->ck_home = m.out;
  if (getenv("DCGANSR_MB_DEBUG"))
    fprintf(stderr, "[dcgansr] micro-batch checkpoint: module %d (%d -> %d channels, %d x %d), %.1f GiB for %d samples, %.1f GiB were free\n", best_i,
            m.cin, m.cout, m.hout, m.wout, (double)B * m.cout * m.hout * m.wout * 4 / (1 << 30), B, (double)free_b / (1 << 30));
}

static int net_forward_mb(dcgansr_net* net, const float* in_all, int B, float* out_all) {
  dcgansr_ctx* ctx = net->ctx;
  St st = ctx->st();
  const int b = net->max_batch, k = B / b;
  const int64_t ie = (int64_t)b * net->in_c * net->in_h * net->in_w, oe = (int64_t)b * net->out_c * net->out_h * net->out_w;
  if (net->nbn > 0) CK(ctx, cudaMemsetAsync(net->mb_fsum, 0, 2 * net->nbn * sizeof(double), ctx->stream));
  const bool ck = net->ck_buf && net->ck_B == B;
  net->ck_valid = false;            // the parameters changed since the last step
  for (size_t j = 0; j < net->mods.size(); ++j) {
    Mod& m = net->mods[j];
    if (m.kind != DCGANSR_BN) continue;
    const int C = m.cin;
    const int64_t P = (int64_t)b * m.hin * m.win;
    const bool from_ck = ck && net->ck_valid && (int)j > net->ck_mod;
    for (int mi = 0; mi < k; ++mi) {
      const float* x = nullptr;
      if (ck) mb_select(net, mi);
      if (int rc = net_forward_dev(net, in_all + mi * ie, b, 1, (int)j, true, &x, from_ck ? net->ck_mod + 1 : 0)) return rc;
      k_bn_stats(st, x, P, C, net->bn_partials, net->bn_sums);
      k_dacc(st, net->mb_fsum + 2 * m.bn_off, net->bn_sums, 2 * C);
    }
    // this pass ran the checkpointed convolution with the final statistics of every BatchNorm before it: its slices are final
    if (ck && (int)j > net->ck_mod) net->ck_valid = true;
    k_bn_finalize(st, net->mb_fsum + 2 * m.bn_off, C, (double)P * k, m.L.eps, m.L.momentum, m.save_mean, m.save_invstd,
                  net->bn_rmean + m.bn_off, net->bn_rvar + m.bn_off);
  }
  for (int mi = 0; mi < k; ++mi) {
    if (ck) mb_select(net, mi);
    if (int rc = net_forward_dev(net, in_all + mi * ie, b, 1, -1, true, nullptr, (ck && net->ck_valid) ? net->ck_mod + 1 : 0)) return rc;
    CK(ctx, cudaMemcpyAsync(out_all + mi * oe, net->last_out, oe * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (ck) { net->ck_valid = true; mb_select(net, -1); }
  net->last_out = nullptr;           // the cached activations are those of the LAST micro-batch only
  net->last_batch = 0;
  CKLAST(ctx);
  return 0;
}

static int net_backward_mb(dcgansr_net* net, const float* in_all, const float* dy_all, int B, bool reduce_buckets) {
  dcgansr_ctx* ctx = net->ctx;
  const int b = net->max_batch, k = B / b;
  const int64_t ie = (int64_t)b * net->in_c * net->in_h * net->in_w, oe = (int64_t)b * net->out_c * net->out_h * net->out_w;
  if (net->nbn > 0) CK(ctx, cudaMemsetAsync(net->mb_bsum, 0, 2 * net->nbn * sizeof(double), ctx->stream));
  const float* dxd = nullptr;
  for (int j = (int)net->mods.size() - 1; j >= 0; --j) {
    if (net->mods[j].kind != DCGANSR_BN) continue;
    MbWalk w{B, j, false};
    // the walk of this pass stops at BatchNorm j: it needs the activations from module j - 1 on; behind the checkpoint they are
    // rebuilt from it (the forward pass of this step left every slice final)
    const bool from_ck = net->ck_buf && net->ck_B == B && net->ck_valid && j > net->ck_mod;
    for (int mi = 0; mi < k; ++mi) {
      if (net->ck_buf && net->ck_B == B) mb_select(net, mi);
      if (int rc = net_forward_dev(net, in_all + mi * ie, b, 1, -1, true, nullptr, from_ck ? net->ck_mod + 1 : 0)) return rc;
      CK(ctx, cudaMemcpyAsync(net->gbuf[0], dy_all + mi * oe, oe * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      if (int rc = net_backward_dev(net, in_all + mi * ie, net->gbuf[0], b, false, false, &dxd, 1, 0, false, &w)) return rc;
    }
  }
  for (int mi = 0; mi < k; ++mi) {
    MbWalk w{B, -1, mi == 0};
    if (net->ck_buf && net->ck_B == B) mb_select(net, mi);
    if (int rc = net_forward_dev(net, in_all + mi * ie, b, 1, -1, true)) return rc;
    CK(ctx, cudaMemcpyAsync(net->gbuf[0], dy_all + mi * oe, oe * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    if (int rc = net_backward_dev(net, in_all + mi * ie, net->gbuf[0], b, true, false, &dxd, 1, 0, reduce_buckets && mi == k - 1, &w)) return rc;
  }
  mb_select(net, -1);
  CKLAST(ctx);
  return 0;
}

static int net_adam_dev(dcgansr_net* net, double lr, double b1, double b2, double eps) {
  dcgansr_ctx* ctx = net->ctx;
  St st = ctx->st();
  k_adam_prep(st, net->adam_t, net->adam_step, lr, b1, b2);
  k_adam(st, net->params, net->grads, net->adam_m, net->adam_v, net->nparams, net->adam_step, b1, b2, eps);
  ++net->params_ver;
  net_pack_all(net);
  CKLAST(ctx);
  return 0;
}

// ------------------------------------------------------------------------------------------
// C ABI: lifecycle
// ------------------------------------------------------------------------------------------
#pragma GCC visibility push(default)
extern "C" {

int dcgansr_version(void) { return 100; }

const char* dcgansr_last_error(dcgansr_ctx* ctx) { return ctx ? ctx->err.c_str() : t_err.c_str(); }

int dcgansr_ctx_create(const dcgansr_cfg* cfg, dcgansr_ctx** out) {
  if (!cfg || !out) return fail(nullptr, DCGANSR_ERR_INVALID, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, DCGANSR_ERR_CUDA, std::string("no CUDA device: libdcgansr has no CPU fallback (") +
                                               cudaGetErrorString(e) + ")");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, DCGANSR_ERR_INVALID, "bad device ordinal");
  if (cfg->precision != DCGANSR_STRICT_FP32 && cfg->precision != DCGANSR_FAST_TF32)
    return fail(nullptr, DCGANSR_ERR_INVALID, "bad precision mode");
  CK(nullptr, cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(nullptr, DCGANSR_ERR_UNSUPPORTED, std::string("libdcgansr is built for sm_100a only; device is ") + prop.name);
  if (cfg->precision == DCGANSR_FAST_TF32) {
    std::string e;
    if (!tc_init(&e)) return fail(nullptr, DCGANSR_ERR_UNSUPPORTED, e);
  }
  dcgansr_ctx* ctx = new dcgansr_ctx();
  ctx->cfg = *cfg;
  if (cfg->precision == DCGANSR_FAST_TF32) {
    ctx->tcws.part_bytes = (size_t)NSM_WS * 128 * 128 * sizeof(float);
    ctx->tcws.ncounters = 4096;
    CK(ctx, cudaMalloc((void**)&ctx->tcws.part, ctx->tcws.part_bytes));
    CK(ctx, cudaMalloc((void**)&ctx->tcws.counters, ctx->tcws.ncounters * sizeof(int)));
    CK(ctx, cudaMemset(ctx->tcws.counters, 0, ctx->tcws.ncounters * sizeof(int)));
  }
  CK(ctx, cudaMalloc((void**)&ctx->tcws.lpart, 256 * sizeof(double)));
  CK(ctx, cudaMalloc((void**)&ctx->tcws.lcounter, sizeof(int)));
  CK(ctx, cudaMemset(ctx->tcws.lcounter, 0, sizeof(int)));
  if (ctx->cfg.world_size < 1) ctx->cfg.world_size = 1;
  CK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CK(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  CK(ctx, cudaEventCreate(&ctx->ev0));
  CK(ctx, cudaEventCreate(&ctx->ev1));
  CK(ctx, cudaEventCreateWithFlags(&ctx->ev_c2m, cudaEventDisableTiming));
  CK(ctx, cudaEventCreateWithFlags(&ctx->ev_m2c, cudaEventDisableTiming));
  CK(ctx, cudaMalloc((void**)&ctx->d_losses, 4 * sizeof(float)));
  CK(ctx, cudaMemset(ctx->d_losses, 0, 4 * sizeof(float)));
  CK(ctx, cudaMallocHost((void**)&ctx->h_losses, 4 * sizeof(float)));
  *out = ctx;
  return 0;
}

// ---- peer-memory exchange area of the one-shot all-reduce (kernels_peer.cu) ------------------
// The exchange areas are pooled per process and device and never freed: freeing memory a peer still has mapped is undefined
// behaviour, and a barrier inside a destructor could hang behind a lost rank.  A context takes a free area of its device (1 MB
// each) and hands it back; every dcgansr_comm_init zeroes it again behind the ranks' rendezvous.
struct PeerArea { int device; void* ptr; bool in_use; };
static std::mutex g_peer_mu;
static std::vector<PeerArea> g_peer_pool;
static void* peer_area_acquire(int device, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_peer_mu);
  for (auto& a : g_peer_pool)
    if (a.device == device && !a.in_use) { a.in_use = true; return a.ptr; }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  g_peer_pool.push_back(PeerArea{device, p, true});
  return p;
}
static void peer_area_release(void* ptr) {
  std::lock_guard<std::mutex> lk(g_peer_mu);
  for (auto& a : g_peer_pool)
    if (a.ptr == ptr) a.in_use = false;
}
static void peer_teardown(dcgansr_ctx* ctx) {
  for (void* m : ctx->peer_maps) if (m) cudaIpcCloseMemHandle(m);
  ctx->peer_maps.clear();
  if (ctx->peer_base) peer_area_release(ctx->peer_base);
  ctx->peer_base = nullptr;
  if (ctx->peer_err_h) cudaFreeHost(ctx->peer_err_h);
  ctx->peer_err_h = nullptr;
  ctx->peer_ok = false;
}
// Every rank allocates its receive area, the cudaIpc handles travel through one NCCL all-reduce (a sum over zero-filled slots is
// a gather), every rank maps every peer, and a second all-reduce makes the outcome unanimous: either all ranks use the peer
// path or all stay on ncclAllReduce (different nodes, no peer access, DCGANSR_PEER_AR=0).
static int peer_setup(dcgansr_ctx* ctx) {
  const char* e = getenv("DCGANSR_PEER_AR");
  const int W = ctx->cfg.world_size, me = ctx->cfg.rank;
  const bool want = !(e && atoi(e) == 0) && W <= PEER_MAXW;
  const int nmax = 4096;
  const size_t data_bytes = (size_t)PEER_SLOTS * PEER_MAXW * nmax * sizeof(double);
  const size_t flag_bytes = (size_t)PEER_SLOTS * PEER_MAXW * sizeof(unsigned long long);
  struct Rec { cudaIpcMemHandle_t h; unsigned char ok; unsigned char pad[63]; };
  static_assert(sizeof(Rec) == 128, "one record per rank");
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  peer_teardown(ctx);
  if (want && (ctx->peer_base = peer_area_acquire(ctx->cfg.device, data_bytes + flag_bytes + 256)) != nullptr) {
    CK(ctx, cudaMemsetAsync(ctx->peer_base, 0, data_bytes + flag_bytes + 256, ctx->stream));
    if (cudaIpcGetMemHandle(&mine.h, ctx->peer_base) == cudaSuccess) mine.ok = 1;
  }
  cudaGetLastError();
  unsigned char* d_rec = nullptr;
  CK(ctx, cudaMalloc((void**)&d_rec, (size_t)W * sizeof(Rec)));
  std::vector<Rec> all((size_t)W);
  auto gather = [&]() -> int {
    CK(ctx, cudaMemsetAsync(d_rec, 0, (size_t)W * sizeof(Rec), ctx->stream));
    CK(ctx, cudaMemcpyAsync(d_rec + (size_t)me * sizeof(Rec), &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
    CKN(ctx, ctx->nccl.AllReduce(d_rec, d_rec, (size_t)W * sizeof(Rec), ncclUint8, ncclSum, ctx->comm, ctx->stream));
    CK(ctx, cudaMemcpyAsync(all.data(), d_rec, (size_t)W * sizeof(Rec), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
  };
  int rc = gather();
  bool ok = rc == 0;
  for (int r = 0; ok && r < W; ++r) ok = all[r].ok == 1;
  std::vector<void*> base((size_t)W, nullptr);
  if (ok) {
    for (int r = 0; r < W; ++r) {
      if (r == me) { base[r] = ctx->peer_base; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      ctx->peer_maps.push_back(ptr);
      base[r] = ptr;
    }
  }
  if (ok && cudaHostAlloc((void**)&ctx->peer_err_h, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); ok = false; }
  // unanimity (also the barrier behind which every rank's flags are known to be zeroed)
  mine.ok = ok ? 1 : 0;
  if (rc == 0) rc = gather();
  for (int r = 0; ok && r < W; ++r) ok = all[r].ok == 1;
  cudaFree(d_rec);
  if (rc != 0 || !ok) { peer_teardown(ctx); return rc; }
  *ctx->peer_err_h = 0;
  PeerAR& p = ctx->peer;
  for (int r = 0; r < W; ++r) {
    p.data[r] = (double*)base[r];
    p.flags[r] = (unsigned long long*)((char*)base[r] + data_bytes);
  }
  p.seq = (unsigned long long*)((char*)ctx->peer_base + data_bytes + flag_bytes);
  int* derr = nullptr;
  CK(ctx, cudaHostGetDevicePointer((void**)&derr, ctx->peer_err_h, 0));
  p.err = derr;
  p.rank = me; p.world = W; p.nmax = nmax;
  ctx->peer_ok = true;
  return 0;
}

void dcgansr_ctx_destroy(dcgansr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaDeviceSynchronize();
  for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
  ctx->graphs.clear();
  // nets that outlive their context (LuaJIT / Python finalizers run in no particular order): release their device memory now
  // and leave plan-only handles behind, so a later dcgansr_net_destroy never touches this ctx
  {
    std::vector<dcgansr_net*> live;
    live.swap(ctx->nets);
    for (dcgansr_net* n : live) net_release_device(n);
  }
  peer_teardown(ctx);
  if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
  for (float* p : ctx->slots) if (p) cudaFree(p);
  if (ctx->tcws.part) cudaFree(ctx->tcws.part);
  if (ctx->tcws.counters) cudaFree(ctx->tcws.counters);
  if (ctx->tcws.lpart) cudaFree(ctx->tcws.lpart);
  if (ctx->tcws.lcounter) cudaFree(ctx->tcws.lcounter);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  if (ctx->tmp) cudaFree(ctx->tmp);
  if (ctx->label_vec) cudaFree(ctx->label_vec);
  if (ctx->lr_buf) cudaFree(ctx->lr_buf);
  if (ctx->fake_buf) cudaFree(ctx->fake_buf);
  if (ctx->d_losses) cudaFree(ctx->d_losses);
  if (ctx->h_losses) cudaFreeHost(ctx->h_losses);
  for (cudaEvent_t e : ctx->prof.pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->graph_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->bucket_ev) cudaEventDestroy(e);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_c2m) cudaEventDestroy(ctx->ev_c2m);
  if (ctx->ev_m2c) cudaEventDestroy(ctx->ev_m2c);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
}

int dcgansr_synchronize(dcgansr_ctx* ctx) {
  if (!ctx) return fail(nullptr, DCGANSR_ERR_INVALID, "null ctx");
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CKLAST(ctx);
  if (ctx->peer_err_h && *(volatile int*)ctx->peer_err_h)
    return fail(ctx, DCGANSR_ERR_NCCL, "peer-memory all-reduce: a rank did not arrive within the time-out");
  return 0;
}
int dcgansr_timer_begin(dcgansr_ctx* ctx) {
  if (!ctx) return fail(nullptr, DCGANSR_ERR_INVALID, "null ctx");
  CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}
int dcgansr_timer_end(dcgansr_ctx* ctx, float* ms_out) {
  if (!ctx || !ms_out) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CK(ctx, cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return 0;
}
int dcgansr_launch_count(dcgansr_ctx* ctx, int64_t* out) {
  if (!ctx || !out) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  *out = ctx->launches;
  return 0;
}
int dcgansr_flush_l2(dcgansr_ctx* ctx) {
  if (!ctx) return fail(nullptr, DCGANSR_ERR_INVALID, "null ctx");
  if (!ctx->flush_buf) {
    ctx->flush_count = (int64_t)64 << 20;   // 256 MiB of floats > 126 MB L2
    CK(ctx, cudaMalloc((void**)&ctx->flush_buf, ctx->flush_count * sizeof(float)));
  }
  k_flush(ctx->st(), ctx->flush_buf, ctx->flush_count);
  CKLAST(ctx);
  return 0;
}

// ---- per-launch event profiler (bench.py roofline leg) ------------------------------------------
int dcgansr_profile_begin(dcgansr_ctx* ctx) {
  if (!ctx) return fail(nullptr, DCGANSR_ERR_INVALID, "null ctx");
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  Prof& p = ctx->prof;
  if (p.pool.empty()) {
    cudaEvent_t e;
    CK(ctx, cudaEventCreate(&e));
    p.pool.push_back(e);
  }
  p.recs.clear();
  CK(ctx, cudaEventRecord(p.pool[0], ctx->stream));
  p.on = true;
  return 0;
}
int dcgansr_profile_end(dcgansr_ctx* ctx, char* json_out, int64_t cap) {
  if (!ctx || !json_out || cap < 64) return fail(ctx, DCGANSR_ERR_INVALID, "bad argument");
  Prof& p = ctx->prof;
  p.on = false;
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  struct Agg { std::string name; double work; int kind; int64_t n; double ms; };
  std::vector<Agg> aggs;
  for (size_t i = 0; i < p.recs.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.pool[i], p.pool[i + 1]) != cudaSuccess) continue;
    const ProfRec& r = p.recs[i];
    Agg* a = nullptr;
    for (auto& x : aggs)
      if (x.name == r.name && x.work == r.work && x.kind == r.kind) { a = &x; break; }
    if (!a) { aggs.push_back(Agg{r.name, r.work, r.kind, 0, 0.0}); a = &aggs.back(); }
    a->n += 1; a->ms += ms;
  }
  std::sort(aggs.begin(), aggs.end(), [](const Agg& a, const Agg& b) { return a.ms > b.ms; });
  std::string js = "[";
  char buf[256];
  for (size_t i = 0; i < aggs.size(); ++i) {
    snprintf(buf, sizeof(buf), "%s{\"name\":\"%s\",\"work\":%.6e,\"kind\":\"%s\",\"launches\":%lld,\"ms\":%.6f}",
             i ? "," : "", aggs[i].name.c_str(), aggs[i].work, aggs[i].kind == WORK_FLOPS ? "flops" : "bytes",
             (long long)aggs[i].n, aggs[i].ms);
    if ((int64_t)(js.size() + strlen(buf) + 2) >= cap) break;
    js += buf;
  }
  js += "]";
  memcpy(json_out, js.c_str(), js.size() + 1);
  p.recs.clear();
  return 0;
}

// ---- communicator ---------------------------------------------------------------------------
static int nccl_load(dcgansr_ctx* ctx) {
  if (ctx->nccl.h) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(ctx, DCGANSR_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
  NcclApi& a = ctx->nccl;
  a.h = h;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
  a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
  a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString)
    return fail(ctx, DCGANSR_ERR_NCCL, "libnccl lacks required symbols");
  return 0;
}
int dcgansr_comm_get_unique_id(dcgansr_ctx* ctx, void* unique_id_128) {
  if (!ctx || !unique_id_128) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (int rc = nccl_load(ctx)) return rc;
  ncclUniqueId id;
  CKN(ctx, ctx->nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(unique_id_128, &id, 128);
  return 0;
}
int dcgansr_comm_init(dcgansr_ctx* ctx, const void* unique_id_128) {
  if (!ctx || !unique_id_128) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (ctx->cfg.world_size <= 1) return 0;
  if (int rc = nccl_load(ctx)) return rc;
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  ncclUniqueId id;
  memcpy(&id, unique_id_128, 128);
  CKN(ctx, ctx->nccl.CommInitRank(&ctx->comm, ctx->cfg.world_size, id, ctx->cfg.rank));
  return peer_setup(ctx);
}

int dcgansr_comm_peer_enabled(dcgansr_ctx* ctx) { return ctx && ctx->peer_ok ? 1 : 0; }

// ---- net description --------------------------------------------------------------------------
void dcgansr_net_destroy(dcgansr_net* net) {
  if (!net) return;
  net_release_device(net);
  for (auto& m : net->mods)
    if (m.conv) { delete m.conv; m.conv = nullptr; }
  delete net;
}

int dcgansr_net_create(dcgansr_ctx* ctx, const dcgansr_layer* layers, int n_layers, int in_c, int in_h, int in_w,
                       int max_batch, dcgansr_net** out) {
  if (!layers || !out || n_layers <= 0 || in_c <= 0 || in_h <= 0 || in_w <= 0 || max_batch <= 0)
    return fail(ctx, DCGANSR_ERR_INVALID, "bad net description");
  *out = nullptr;
  dcgansr_net* net = new dcgansr_net();
  net->ctx = ctx;
  net->in_c = in_c; net->in_h = in_h; net->in_w = in_w; net->max_batch = max_batch;
  int c = in_c, h = in_h, w = in_w;
  int64_t poff = 0, bnoff = 0;
  std::string err;
  for (int i = 0; i < n_layers && err.empty(); ++i) {
    Mod m;
    m.kind = layers[i].kind;
    m.L = layers[i];
    m.cin = c; m.hin = h; m.win = w;
    switch (m.kind) {
      case DCGANSR_CONV:
      case DCGANSR_FULLCONV: {
        const dcgansr_layer& L = layers[i];
        if (L.cin != c) { err = "layer " + std::to_string(i) + ": cin does not match the incoming channel count"; break; }
        if (L.kh != L.kw || L.sh != L.sw || L.ph != L.pw || L.adjh != L.adjw) { err = "only square kernels / strides / pads"; break; }
        m.conv = new ConvPlan();
        err = m.conv->build(m.kind == DCGANSR_FULLCONV, L.cin, L.cout, L.kh, L.sh > 0 ? L.sh : 1, L.ph, L.adjh, h, w);
        if (!err.empty()) { delete m.conv; m.conv = nullptr; err = "layer " + std::to_string(i) + ": " + err; break; }
        m.p_off = poff; m.p_cnt = m.conv->weight_count(); poff += m.p_cnt;
        c = L.cout; h = m.conv->Hout; w = m.conv->Wout;
        if (net->first_param_mod < 0) net->first_param_mod = i;
        break;
      }
      case DCGANSR_BN:
        if (layers[i].cout != c) { err = "layer " + std::to_string(i) + ": BN channel count mismatch"; break; }
        m.p_off = poff; m.p_cnt = 2 * (int64_t)c; poff += m.p_cnt;
        m.bn_off = bnoff; bnoff += c;
        if (m.L.eps <= 0.f) m.L.eps = 1e-5f;
        if (net->first_param_mod < 0) net->first_param_mod = i;
        break;
      case DCGANSR_RELU: case DCGANSR_LRELU: case DCGANSR_TANH: case DCGANSR_SIGMOID:
        m.act = act_of_kind(m.kind);
        break;
      case DCGANSR_UPNEAREST:
        if (m.L.scale <= 0) m.L.scale = 2;
        h *= m.L.scale; w *= m.L.scale;
        break;
      case DCGANSR_VIEW: break;
      default: err = "layer " + std::to_string(i) + ": unknown kind";
    }
    m.cout = c; m.hout = h; m.wout = w;
    net->mods.push_back(m);
  }
  if (!err.empty()) { dcgansr_net_destroy(net); return fail(ctx, DCGANSR_ERR_INVALID, err); }
  net->nparams = poff; net->nbn = bnoff;
  net->out_c = c; net->out_h = h; net->out_w = w;
  // activation fusion: an ACT right after CONV / BN runs in that module's epilogue
  for (size_t i = 1; i < net->mods.size(); ++i) {
    Mod& m = net->mods[i];
    Mod& pm = net->mods[i - 1];
    if (m.act != ACT_NONE && (pm.kind == DCGANSR_CONV || pm.kind == DCGANSR_FULLCONV || pm.kind == DCGANSR_BN) &&
        pm.fused_act == ACT_NONE) {
      pm.fused_act = m.act; pm.fused_neg = m.L.negval; m.fused_into_prev = true;
    }
  }
  // gradient buckets: [conv, following BN ...) ranges
  {
    int64_t start = 0;
    bool seen = false;
    Mod* prev_conv = nullptr;
    for (auto& m : net->mods) {
      if (m.conv) {
        if (seen) { prev_conv->bucket = (int)net->buckets.size(); net->buckets.push_back({start, m.p_off - start}); }
        start = m.p_off; seen = true; prev_conv = &m;
      }
    }
    if (net->nparams > start || !seen) {
      if (prev_conv) prev_conv->bucket = (int)net->buckets.size();
      net->buckets.push_back({start, net->nparams - start});
    }
  }
  if (!ctx) { *out = net; return 0; }

  // ---- device memory ----
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  int64_t np4 = (net->nparams + 3) / 4 * 4 + 4;
  auto dalloc = [&](float** p, int64_t n) -> cudaError_t {
    cudaError_t e = cudaMalloc((void**)p, (size_t)std::max<int64_t>(n, 4) * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, (size_t)std::max<int64_t>(n, 4) * sizeof(float));
    return e;
  };
  cudaError_t e = cudaSuccess;
  auto A = [&](float** p, int64_t n) { if (e == cudaSuccess) e = dalloc(p, n); };
  net->stage_len = (2 * net->nbn + 4 + 3) / 4 * 4;
  A(&net->params, np4); A(&net->grads_base, np4 + net->stage_len); A(&net->adam_m, np4); A(&net->adam_v, np4);
  if (e == cudaSuccess) net->grads = net->grads_base + net->stage_len;
  A(&net->adam_step, 4);
  A(&net->bn_rmean, net->nbn); A(&net->bn_rvar, net->nbn); A(&net->bn_save, 4 * net->nbn);          // (mean, invstd) x 2 sample groups
  if (e == cudaSuccess) e = cudaMalloc((void**)&net->adam_t, sizeof(int64_t));
  if (e == cudaSuccess) e = cudaMemset(net->adam_t, 0, sizeof(int64_t));
  int64_t B = max_batch;
  int64_t gmax = B * in_c * in_h * in_w;
  int maxC = 1;
  int64_t maxP = 1;
  size_t wsc = 0;
  for (auto& m : net->mods) {
    int64_t oe = B * m.cout * m.hout * m.wout;
    gmax = std::max(gmax, std::max(oe, B * m.cin * m.hin * m.win));
    bool needs_buf = m.kind == DCGANSR_CONV || m.kind == DCGANSR_FULLCONV || m.kind == DCGANSR_BN ||
                     m.kind == DCGANSR_UPNEAREST || (m.act != ACT_NONE && !m.fused_into_prev);
    if (needs_buf) { A(&m.out, oe); m.owns_out = true; }
    if (m.kind == DCGANSR_BN) {
      maxC = std::max(maxC, m.cin);
      maxP = std::max(maxP, B * m.hin * m.win);
      m.save_mean = net->bn_save + m.bn_off;                      // group g: + g * 2 * nbn
      m.save_invstd = net->bn_save + net->nbn + m.bn_off;
    }
    if (m.conv) {
      if (e == cudaSuccess && m.conv->alloc_device(ctx) != 0) e = cudaErrorMemoryAllocation;
      wsc = std::max(wsc, m.conv->wscratch_bytes((int)B));
    }
  }
  // aliases for fused ACT / VIEW modules
  for (size_t i = 0; i < net->mods.size(); ++i) {
    Mod& m = net->mods[i];
    if (!m.owns_out) m.out = i > 0 ? net->mods[i - 1].out : nullptr;
  }
  net->gelems = gmax;
  A(&net->in_buf, B * in_c * in_h * in_w);
  A(&net->gbuf[0], gmax); A(&net->gbuf[1], gmax);
  {
    int64_t rows = 0;
    for (auto& m : net->mods)
      if (m.kind == DCGANSR_BN)      // rows of the BN kernels' own partials, or one row per persistent CTA of a producing halo-kernel convolution
        rows = std::max<int64_t>(rows, (int64_t)std::max(bn_partial_rows(B * m.hin * m.win, m.cin), 160) * 2 * m.cin);
    float* tmpf = nullptr;
    A(&tmpf, std::max<int64_t>(rows, 2) * 2 * 2); net->bn_partials = (double*)tmpf; tmpf = nullptr;     // x 2 sample groups
    A(&tmpf, (int64_t)maxC * 4 * 2 + 4); net->bn_sums = (double*)tmpf; tmpf = nullptr;
    A(&tmpf, (int64_t)maxC * 4 + 4); net->bn_sums_total = (double*)tmpf;
    A(&net->bn_fmeans, (int64_t)maxC * 4 + 4);
    A(&tmpf, net->nbn * 4 + 4); net->mb_fsum = (double*)tmpf; tmpf = nullptr;
    A(&tmpf, net->nbn * 4 + 4); net->mb_bsum = (double*)tmpf; tmpf = nullptr;
  }
  net->wscratch_bytes = wsc;
  A(&net->wscratch, (int64_t)(wsc / sizeof(float)) + 4);
  if (e != cudaSuccess) {
    std::string msg = std::string("device allocation failed: ") + cudaGetErrorString(e);
    dcgansr_net_destroy(net);
    return fail(ctx, DCGANSR_ERR_NOMEM, msg);
  }
  {
    std::vector<PackJob> jobs;
    int64_t total = 0;
    for (auto& m : net->mods)
      if (m.conv) m.conv->collect_jobs(jobs, net->params + m.p_off, total);
    if (!jobs.empty() && jobs.size() <= 128) {
      PackJob sentinel = jobs.back();
      sentinel.begin = total;
      jobs.push_back(sentinel);
      if (cudaMalloc((void**)&net->pack_jobs, jobs.size() * sizeof(PackJob)) == cudaSuccess &&
          cudaMemcpy(net->pack_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice) == cudaSuccess) {
        net->n_pack_jobs = (int)jobs.size() - 1;
        net->pack_total = total;
      } else if (net->pack_jobs) { cudaFree(net->pack_jobs); net->pack_jobs = nullptr; }
    }
  }
  // BN running_var starts at 1 (Torch7 init)
  if (net->nbn > 0) k_fill(ctx->st(), net->bn_rvar, net->nbn, 1.f);
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->nets.push_back(net);
  *out = net;
  return 0;
}

int dcgansr_net_out_shape(dcgansr_net* net, int* c, int* h, int* w) {
  if (!net) return fail(nullptr, DCGANSR_ERR_INVALID, "null net");
  if (c) *c = net->out_c;
  if (h) *h = net->out_h;
  if (w) *w = net->out_w;
  return 0;
}
int dcgansr_net_num_params(dcgansr_net* net, int64_t* out) {
  if (!net || !out) return fail(nullptr, DCGANSR_ERR_INVALID, "null argument");
  *out = net->nparams;
  return 0;
}
int dcgansr_net_num_bn_channels(dcgansr_net* net, int64_t* out) {
  if (!net || !out) return fail(nullptr, DCGANSR_ERR_INVALID, "null argument");
  *out = net->nbn;
  return 0;
}

#define NEED_DEV(net)                                                                             \
  if (!(net)) return fail(nullptr, DCGANSR_ERR_INVALID, "null net");                               \
  if (!(net)->ctx) return fail(nullptr, DCGANSR_ERR_INVALID, "plan-only net: no CUDA context");     \
  dcgansr_ctx* ctx = (net)->ctx;                                                                   \
  CK(ctx, cudaSetDevice(ctx->cfg.device));

int dcgansr_net_set_params(dcgansr_net* net, const float* host_flat) {
  NEED_DEV(net);
  if (!host_flat) return fail(ctx, DCGANSR_ERR_INVALID, "null params");
  CK(ctx, cudaMemcpyAsync(net->params, host_flat, net->nparams * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  ++net->params_ver;
  net_pack_all(net);
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CKLAST(ctx);
  return 0;
}
static int d2h(dcgansr_ctx* ctx, float* host, const float* dev, int64_t n) {
  if (n <= 0 || !host) return 0;
  CK(ctx, cudaMemcpyAsync(host, dev, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CKLAST(ctx);
  return 0;
}
static int h2d(dcgansr_ctx* ctx, float* dev, const float* host, int64_t n) {
  if (n <= 0 || !host) return 0;
  CK(ctx, cudaMemcpyAsync(dev, host, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
int dcgansr_net_get_params(dcgansr_net* net, float* host_flat) { NEED_DEV(net); return d2h(ctx, host_flat, net->params, net->nparams); }
int dcgansr_net_get_grads(dcgansr_net* net, float* host_flat) { NEED_DEV(net); return d2h(ctx, host_flat, net->grads, net->nparams); }
int dcgansr_net_get_bn_running(dcgansr_net* net, float* mean, float* var) {
  NEED_DEV(net);
  if (int rc = d2h(ctx, mean, net->bn_rmean, net->nbn)) return rc;
  return d2h(ctx, var, net->bn_rvar, net->nbn);
}
int dcgansr_net_set_bn_running(dcgansr_net* net, const float* mean, const float* var) {
  NEED_DEV(net);
  if (int rc = h2d(ctx, net->bn_rmean, mean, net->nbn)) return rc;
  return h2d(ctx, net->bn_rvar, var, net->nbn);
}
int dcgansr_net_get_adam_state(dcgansr_net* net, float* m, float* v, int64_t* t) {
  NEED_DEV(net);
  if (int rc = d2h(ctx, m, net->adam_m, net->nparams)) return rc;
  if (int rc = d2h(ctx, v, net->adam_v, net->nparams)) return rc;
  if (t) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(t, net->adam_t, sizeof(int64_t), cudaMemcpyDeviceToHost));
  }
  return 0;
}
int dcgansr_net_set_adam_state(dcgansr_net* net, const float* m, const float* v, int64_t t) {
  NEED_DEV(net);
  if (int rc = h2d(ctx, net->adam_m, m, net->nparams)) return rc;
  if (int rc = h2d(ctx, net->adam_v, v, net->nparams)) return rc;
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CK(ctx, cudaMemcpy(net->adam_t, &t, sizeof(int64_t), cudaMemcpyHostToDevice));
  return 0;
}

// ---- Torch7-shaped net ops -------------------------------------------------------------------
int dcgansr_net_forward(dcgansr_net* net, const float* x, int batch, float* y) {
  NEED_DEV(net);
  if (!x || batch <= 0 || batch > net->max_batch) return fail(ctx, DCGANSR_ERR_INVALID, "bad batch / null input");
  if (int rc = upload_nchw(ctx, x, batch, net->in_c, net->in_h, net->in_w, net->in_buf)) return rc;
  if (int rc = net_forward_dev(net, net->in_buf, batch)) return rc;
  if (y) return download_nchw(ctx, net->last_out, batch, net->out_c, net->out_h, net->out_w, y);
  return 0;
}

static int net_backward_host(dcgansr_net* net, const float* x, const float* dy, int batch, float* dx, bool acc) {
  NEED_DEV(net);
  if (!dy || batch <= 0 || batch > net->max_batch) return fail(ctx, DCGANSR_ERR_INVALID, "bad batch / null gradOutput");
  if (batch != net->last_batch) return fail(ctx, DCGANSR_ERR_INVALID, "backward batch differs from the cached forward");
  if (x)
    if (int rc = upload_nchw(ctx, x, batch, net->in_c, net->in_h, net->in_w, net->in_buf)) return rc;
  if (int rc = upload_nchw(ctx, dy, batch, net->out_c, net->out_h, net->out_w, net->gbuf[0])) return rc;
  const float* dxd = nullptr;
  if (int rc = net_backward_dev(net, net->in_buf, net->gbuf[0], batch, acc, dx != nullptr, &dxd)) return rc;
  if (dx) {
    if (!dxd) return fail(ctx, DCGANSR_ERR_INVALID, "no input gradient produced");
    return download_nchw(ctx, dxd, batch, net->in_c, net->in_h, net->in_w, dx);
  }
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  CKLAST(ctx);
  return 0;
}
int dcgansr_net_backward(dcgansr_net* net, const float* x, const float* dy, int batch, float* dx) {
  return net_backward_host(net, x, dy, batch, dx, true);
}
int dcgansr_net_update_grad_input(dcgansr_net* net, const float* x, const float* dy, int batch, float* dx) {
  return net_backward_host(net, x, dy, batch, dx, false);
}
int dcgansr_net_zero_grads(dcgansr_net* net) {
  NEED_DEV(net);
  CK(ctx, cudaMemsetAsync(net->grads, 0, net->nparams * sizeof(float), ctx->stream));
  return 0;
}
int dcgansr_net_adam(dcgansr_net* net, double lr, double beta1, double beta2, double eps) {
  NEED_DEV(net);
  return net_adam_dev(net, lr, beta1, beta2, eps);
}

// ---- the fused step ---------------------------------------------------------------------------
static bool overlap_on(dcgansr_ctx* ctx) { return ctx->world() > 1 && !getenv("DCGANSR_NO_OVERLAP"); }

// Stage = what rides along with the gradient exchange of a net (local-batch-statistics mode: the running statistics averaged over
// the ranks; for G also the three step losses).  Packed on the main stream before the backward walk whose last bucket carries it.
static void stage_pack(dcgansr_ctx* ctx, dcgansr_net* net, const float* losses) {
  if (ctx->world() <= 1) return;
  k_stage_pack(ctx->st(), net->bn_rmean, net->bn_rvar, (int)net->nbn, losses, net->grads_base, 1.f / ctx->world());
  net->stage_live = true;
}
// bucketed == true: the backward walk already issued every bucket on the communication stream -> just join
static int allreduce_grads(dcgansr_ctx* ctx, dcgansr_net* net, bool bucketed, float* losses) {
  if (ctx->world() <= 1) return 0;
  const bool staged = net->stage_live || bucketed;       // bucketed: bucket 0 took the stage with it (stage_live already cleared)
  if (bucketed) {
    if (int rc = bucket_join(ctx)) return rc;
    if (net->stage_live) {      // no bucket 0 at offset 0 (should not happen with these graphs): the stage travels alone
      if (int rc = nccl_allreduce(ctx, net->grads_base, net->stage_len, ncclFloat, ctx->stream)) return rc;
      net->stage_live = false;
    }
  } else {
    float* ptr = net->grads;
    int64_t cnt = net->nparams;
    if (net->stage_live) { ptr = net->grads_base; cnt += net->stage_len; net->stage_live = false; }
    if (int rc = nccl_allreduce(ctx, ptr, cnt, ncclFloat, ctx->stream)) return rc;
  }
  if (staged) k_stage_unpack(ctx->st(), net->grads_base, ctx->cfg.sync_bn ? nullptr : net->bn_rmean, ctx->cfg.sync_bn ? nullptr : net->bn_rvar,
                             (int)net->nbn, losses);
  return 0;
}

static int step_body(dcgansr_ctx* ctx, dcgansr_net* G, dcgansr_net* D, const dcgansr_step_cfg* cfg, const float* real, int B) {
  St st = ctx->st();
  const int world = ctx->world();
  const int64_t dcount = (int64_t)B * D->out_c * D->out_h * D->out_w;   // nElement of D's output
  const int64_t per = dcount / B;
  const double n_total = (double)dcount * world;
  const int lossk = cfg->loss == DCGANSR_LOSS_BCE ? LOSS_BCE : LOSS_MSE;
  const float* dxd = nullptr;
  const bool gmb = B > G->max_batch;                  // generator micro-batched with exact whole-batch BatchNorm (config C5)

  // ---------------- fDx (train.lua:208-253) ----------------
  CK(ctx, cudaMemsetAsync(D->grads, 0, D->nparams * sizeof(float), ctx->stream));
  const bool paired = D->max_batch >= 2 * B && !getenv("DCGANSR_NO_PAIRED_D");
  const float* fake = nullptr;
  int fake_group = 0;
  if (!paired) {
    if (int rc = net_forward_dev(D, real, B)) return rc;
    k_loss(st, lossk, D->last_out, dcount, nullptr, per, cfg->real_label, n_total, ctx->d_losses + 0, D->gbuf[0]);
    if (int rc = net_backward_dev(D, real, D->gbuf[0], B, true, false, &dxd)) return rc;

    k_avgpool2(st, real, ctx->lr_buf, B, D->in_h, D->in_w, D->in_c);                  // train.lua:225-230
    if (gmb) {
      if (int rc = net_forward_mb(G, ctx->lr_buf, B, ctx->fake_buf)) return rc;
      fake = ctx->fake_buf;
    } else {
      if (int rc = net_forward_dev(G, ctx->lr_buf, B)) return rc;                     // :233-234
      fake = G->last_out;
    }
    const float* lvec = nullptr;
    if (cfg->pixel_label) {                                                           // :237-239,245
      k_pixel_mse(st, real, fake, ctx->label_vec, B, (int64_t)D->in_c * D->in_h * D->in_w, cfg->pixel_div);
      lvec = ctx->label_vec;
    }
    if (int rc = net_forward_dev(D, fake, B)) return rc;                              // :242-243
    k_loss(st, lossk, D->last_out, dcount, lvec, per, cfg->fake_label, n_total, ctx->d_losses + 1, D->gbuf[0]);
    stage_pack(ctx, D, nullptr);
    if (int rc = net_backward_dev(D, fake, D->gbuf[0], B, true, false, &dxd, 1, 0, overlap_on(ctx))) return rc;      // second pass: grads final
  } else {
    // D(real) and D(fake) as ONE pass over [real; fake] (D was created for >= 2B samples): the generator forward does not
    // depend on D, so it moves first; convolutions then run once on 2B samples (these layers are latency / L2-bound at
    // B samples), BatchNorm keeps per-minibatch statistics (groups = 2) and updates the running statistics real-then-fake,
    // gradients accumulate over both halves exactly as the two reference backward calls do (train.lua:218-248).
    k_avgpool2(st, real, ctx->lr_buf, B, D->in_h, D->in_w, D->in_c);
    const int64_t isz = (int64_t)B * D->in_c * D->in_h * D->in_w;
    if (real != D->in_buf) CK(ctx, cudaMemcpyAsync(D->in_buf, real, isz * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    if (gmb) {
      if (int rc = net_forward_mb(G, ctx->lr_buf, B, D->in_buf + isz)) return rc;     // micro-batches write straight into D's fake half
    } else {
      if (int rc = net_forward_dev(G, ctx->lr_buf, B)) return rc;
      CK(ctx, cudaMemcpyAsync(D->in_buf + isz, G->last_out, isz * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    fake = D->in_buf + isz;
    fake_group = 1;
    const float* lvec = nullptr;
    if (cfg->pixel_label) {
      k_pixel_mse(st, D->in_buf, fake, ctx->label_vec, B, (int64_t)D->in_c * D->in_h * D->in_w, cfg->pixel_div);
      lvec = ctx->label_vec;
    }
    if (int rc = net_forward_dev(D, D->in_buf, B, 2)) return rc;
    k_loss(st, lossk, D->last_out, dcount, nullptr, per, cfg->real_label, n_total, ctx->d_losses + 0, D->gbuf[0]);
    k_loss(st, lossk, D->last_out + dcount, dcount, lvec, per, cfg->fake_label, n_total, ctx->d_losses + 1, D->gbuf[0] + dcount);
    stage_pack(ctx, D, nullptr);
    if (int rc = net_backward_dev(D, D->in_buf, D->gbuf[0], B, true, false, &dxd, 2, 0, overlap_on(ctx))) return rc;
  }
  if (int rc = allreduce_grads(ctx, D, overlap_on(ctx), nullptr)) return rc;
  if (int rc = net_adam_dev(D, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps)) return rc;   // optim.adam(fDx) :280

  // ---------------- fGx (train.lua:256-272) ----------------
  CK(ctx, cudaMemsetAsync(G->grads, 0, G->nparams * sizeof(float), ctx->stream));
  // stale netD.output (pre-Adam forward on fake), post-Adam weights in the dgrad walk
  const float* dout_fake = D->last_out + (int64_t)fake_group * dcount;
  k_loss(st, lossk, dout_fake, dcount, nullptr, per, cfg->gen_label, n_total, ctx->d_losses + 2, D->gbuf[0]);
  if (int rc = net_backward_dev(D, fake, D->gbuf[0], B, false, true, &dxd, 1, fake_group)) return rc;  // netD:updateGradInput :268
  const float* dummy = nullptr;
  stage_pack(ctx, G, ctx->d_losses);                  // the three loss scalars (each rank's share of the global mean) ride with G's gradients
  if (gmb) {
    if (int rc = net_backward_mb(G, ctx->lr_buf, dxd, B, overlap_on(ctx))) return rc;
  } else if (int rc = net_backward_dev(G, ctx->lr_buf, dxd, B, true, false, &dummy, 1, 0, overlap_on(ctx))) return rc; // netG:backward :270
  if (int rc = allreduce_grads(ctx, G, overlap_on(ctx), ctx->d_losses)) return rc;
  if (int rc = net_adam_dev(G, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps)) return rc;   // optim.adam(fGx) :283
  (void)world;
  return 0;
}

static int check_step_args(dcgansr_ctx* ctx, dcgansr_net* G, dcgansr_net* D, const dcgansr_step_cfg* cfg, int B) {
  if (!ctx || !G || !D || !cfg) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (G->ctx != ctx || D->ctx != ctx) return fail(ctx, DCGANSR_ERR_INVALID, "nets belong to another context");
  if (B <= 0 || B > D->max_batch) return fail(ctx, DCGANSR_ERR_INVALID, "batch exceeds max_batch");
  if (B > G->max_batch) {      // the generator runs micro-batched (net_forward_mb / net_backward_mb)
    if (B % G->max_batch) return fail(ctx, DCGANSR_ERR_INVALID, "batch must be a multiple of netG's max_batch (its micro-batch size)");
    if (ctx->cfg.sync_bn && ctx->cfg.world_size > 1) return fail(ctx, DCGANSR_ERR_UNSUPPORTED, "micro-batched generator with sync_bn");
  }
  if (D->in_h % 2 || D->in_w % 2) return fail(ctx, DCGANSR_ERR_INVALID, "D input must have even spatial size");
  if (G->in_c != D->in_c || G->in_h * 2 != D->in_h || G->in_w * 2 != D->in_w)
    return fail(ctx, DCGANSR_ERR_INVALID, "G input must be the 2x2 down-sampled D input");
  if (G->out_c != D->in_c || G->out_h != D->in_h || G->out_w != D->in_w)
    return fail(ctx, DCGANSR_ERR_INVALID, "G output shape must equal D input shape");
  if (cfg->loss != DCGANSR_LOSS_BCE && cfg->loss != DCGANSR_LOSS_MSE) return fail(ctx, DCGANSR_ERR_INVALID, "bad loss kind");
  if (ctx->peer_err_h && *(volatile int*)ctx->peer_err_h)
    return fail(ctx, DCGANSR_ERR_NCCL, "peer-memory all-reduce: a rank did not arrive within the time-out (earlier step)");
  if (ctx->cfg.world_size > 1 && !ctx->comm)
    return fail(ctx, DCGANSR_ERR_NCCL, "world_size > 1 but dcgansr_comm_init was never called: the ranks would train unsynchronised replicas");
  return 0;
}

static int step_run(dcgansr_ctx* ctx, dcgansr_net* G, dcgansr_net* D, const dcgansr_step_cfg* cfg, const float* real_dev, int B,
                    float* out_losses) {
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  size_t lr_bytes = (size_t)B * G->in_c * G->in_h * G->in_w * sizeof(float);
  if (int rc = ensure(ctx, &ctx->lr_buf, &ctx->lr_cap, std::max<size_t>(lr_bytes, 16))) return rc;
  if (int rc = ensure(ctx, &ctx->label_vec, &ctx->label_cap, std::max<size_t>((size_t)B * sizeof(float), 16))) return rc;
  if (B > G->max_batch) {
    if (int rc = ensure(ctx, &ctx->fake_buf, &ctx->fake_cap, (size_t)B * D->in_c * D->in_h * D->in_w * sizeof(float))) return rc;
    mb_ensure_ckpt(ctx, G, B);
  }
  if (ctx->cfg.use_graph && !ctx->prof.on) {       // the per-launch event profiler needs eager launches
    GraphEntry* ge = nullptr;
    for (auto& g : ctx->graphs)
      if (g.G == G && g.D == D && g.real == real_dev && g.batch == B && memcmp(&g.cfg, cfg, sizeof(*cfg)) == 0) ge = &g;
    if (!ge) {
      if (ctx->graphs.size() >= 64) {
        for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
        ctx->graphs.clear();
      }
      cudaGraph_t graph = nullptr;
      int64_t before = ctx->launches;
      CK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      int rc = step_body(ctx, G, D, cfg, real_dev, B);
      cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
      int64_t nodes = ctx->launches - before;
      ctx->launches = before;
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return fail(ctx, DCGANSR_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
      GraphEntry ne;
      e = cudaGraphInstantiate(&ne.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return fail(ctx, DCGANSR_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
      ne.G = G; ne.D = D; ne.real = real_dev; ne.batch = B; ne.cfg = *cfg; ne.launches = nodes;
      ctx->graphs.push_back(ne);
      ge = &ctx->graphs.back();
    }
    if (ctx->world() > 1) {
      // graphs that contain NCCL collectives: one replay in flight (measured on 2 x B200: with the host running many
      // replays ahead a step takes 12 ms, with two in flight 2.96 ms, with one 2.54 ms)
      cudaEvent_t& ev = ctx->graph_ev[0];
      if (!ev) CK(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      else CK(ctx, cudaEventSynchronize(ev));
      CK(ctx, cudaGraphLaunch(ge->exec, ctx->stream));
      CK(ctx, cudaEventRecord(ev, ctx->stream));
      ++ctx->graph_seq;
    } else {
      CK(ctx, cudaGraphLaunch(ge->exec, ctx->stream));
    }
    ctx->launches += ge->launches;
  } else {
    if (int rc = step_body(ctx, G, D, cfg, real_dev, B)) return rc;
  }
  if (out_losses) {
    CK(ctx, cudaMemcpyAsync(ctx->h_losses, ctx->d_losses, 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CKLAST(ctx);
    out_losses[0] = ctx->h_losses[0]; out_losses[1] = ctx->h_losses[1]; out_losses[2] = ctx->h_losses[2];
  }
  return 0;
}

int dcgansr_stage_batch(dcgansr_ctx* ctx, dcgansr_net* netD, const float* real_host, int local_batch, int slot) {
  if (!ctx || !netD || !real_host || local_batch <= 0 || slot < 0 || slot > 63)
    return fail(ctx, DCGANSR_ERR_INVALID, "bad stage_batch argument");
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  if ((int)ctx->slots.size() <= slot) { ctx->slots.resize(slot + 1, nullptr); ctx->slot_bytes.resize(slot + 1, 0); }
  size_t bytes = (size_t)local_batch * netD->in_c * netD->in_h * netD->in_w * sizeof(float);
  if (int rc = ensure(ctx, &ctx->slots[slot], &ctx->slot_bytes[slot], bytes)) return rc;
  return upload_nchw(ctx, real_host, local_batch, netD->in_c, netD->in_h, netD->in_w, ctx->slots[slot]);
}
int dcgansr_train_step_staged(dcgansr_ctx* ctx, dcgansr_net* netG, dcgansr_net* netD, const dcgansr_step_cfg* cfg, int slot,
                              int local_batch, float* out_losses) {
  if (int rc = check_step_args(ctx, netG, netD, cfg, local_batch)) return rc;
  if (slot < 0 || slot >= (int)ctx->slots.size() || !ctx->slots[slot]) return fail(ctx, DCGANSR_ERR_INVALID, "empty batch slot");
  if (ctx->slot_bytes[slot] < (size_t)local_batch * netD->in_c * netD->in_h * netD->in_w * sizeof(float))
    return fail(ctx, DCGANSR_ERR_INVALID, "staged batch smaller than local_batch");
  return step_run(ctx, netG, netD, cfg, ctx->slots[slot], local_batch, out_losses);
}
int dcgansr_train_step(dcgansr_ctx* ctx, dcgansr_net* netG, dcgansr_net* netD, const dcgansr_step_cfg* cfg,
                       const float* real_host, int local_batch, float* out_losses) {
  if (int rc = check_step_args(ctx, netG, netD, cfg, local_batch)) return rc;
  if (!real_host) return fail(ctx, DCGANSR_ERR_INVALID, "null batch");
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  if (int rc = upload_nchw(ctx, real_host, local_batch, netD->in_c, netD->in_h, netD->in_w, netD->in_buf)) return rc;
  return step_run(ctx, netG, netD, cfg, netD->in_buf, local_batch, out_losses);
}
#define NEED_CTX(ctx)                                                          \
  if (!(ctx)) return fail(nullptr, DCGANSR_ERR_INVALID, "null ctx");            \
  CK(ctx, cudaSetDevice((ctx)->cfg.device));

// ---- patch extraction / re-assembly ----------------------------------------------------------
static int patch_args_ok(dcgansr_ctx* ctx, int k, int h, int w, int patch, int line, int nper, int stride) {
  if (k <= 0 || h <= 0 || w <= 0 || patch <= 0 || line <= 0 || nper <= 0 || stride <= 0)
    return fail(ctx, DCGANSR_ERR_INVALID, "bad patch geometry");
  const int rows = (nper + line - 1) / line;
  if ((rows - 1) * stride + patch > h || (line - 1) * stride + patch > w)
    return fail(ctx, DCGANSR_ERR_INVALID, "patches reach outside the image");
  return 0;
}
int dcgansr_extract_patches(dcgansr_ctx* ctx, const float* images, float* patches, int k, int h, int w, int patch, int line, int nper,
                            int stride) {
  NEED_CTX(ctx);
  if (!images || !patches) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (int rc = patch_args_ok(ctx, k, h, w, patch, line, nper, stride)) return rc;
  Arena ar;
  const int64_t ni = (int64_t)k * h * w, np = (int64_t)k * nper * patch * patch;
  float *dI = ar.f(ni), *dP = ar.f(np);
  if (!dI || !dP) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dI, images, ni * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_extract_patches(ctx->st(), dI, dP, k, h, w, patch, line, nper, stride);
  return d2h(ctx, patches, dP, np);
}
int dcgansr_assemble_patches(dcgansr_ctx* ctx, const float* patches, float* images, int k, int h, int w, int patch, int line, int nper,
                             int stride) {
  NEED_CTX(ctx);
  if (!images || !patches) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (int rc = patch_args_ok(ctx, k, h, w, patch, line, nper, stride)) return rc;
  Arena ar;
  const int64_t ni = (int64_t)k * h * w, np = (int64_t)k * nper * patch * patch;
  float *dI = ar.f(ni), *dP = ar.f(np);
  if (!dI || !dP) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dI, images, ni * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));     // uncovered pixels keep their value
  CK(ctx, cudaMemcpyAsync(dP, patches, np * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_assemble_patches(ctx->st(), dP, dI, k, h, w, patch, line, nper, stride);
  return d2h(ctx, images, dI, ni);
}
// minimum-error boundary cut stitching of overlapping generated patches (train-gray-patch-batch-overlap.lua:457-694)
int dcgansr_stitch_overlap(dcgansr_ctx* ctx, const float* patches, float* images, int k, int h, int w, int patch, int overlap, int flags) {
  NEED_CTX(ctx);
  if (!images || !patches) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (k < 1 || h < 1 || w < 1 || patch < 1 || overlap < 2 || patch <= overlap || h != w)
    return fail(ctx, DCGANSR_ERR_INVALID, "bad stitch geometry (square images, 2 <= overlap < patch)");
  if ((h - overlap) % (patch - overlap) != 0) return fail(ctx, DCGANSR_ERR_INVALID, "(fineSize - overlap) must be a multiple of (patchSize - overlap)");
  const int line = (h - overlap) / (patch - overlap);                 // overlapPatchLine (:387)
  if (line < 1 || (line - 1) * overlap + patch > h) return fail(ctx, DCGANSR_ERR_INVALID, "patches reach outside the image");
  if (!stitch_overlap_supported(patch, line, overlap)) return fail(ctx, DCGANSR_ERR_UNSUPPORTED, "stitch: patch <= 32, overlap <= 16, line*line*patch <= 48K");
  Arena ar;
  const int64_t ni = (int64_t)k * h * w, np = (int64_t)k * line * line * patch * patch;
  float *dI = ar.f(ni), *dP = ar.f(np);
  if (!dI || !dP) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dI, images, ni * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));     // uncovered pixels keep their value
  CK(ctx, cudaMemcpyAsync(dP, patches, np * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_stitch_overlap(ctx->st(), dP, dI, k, h, w, patch, line, overlap, flags);
  return d2h(ctx, images, dI, ni);
}
// image.scale(src, dw, dh) in its default bilinear mode, enlarging only (train-gray-3.lua:399)
int dcgansr_scale_bilinear(dcgansr_ctx* ctx, const float* src, float* dst, int n, int h, int w, int dh, int dw) {
  NEED_CTX(ctx);
  if (!src || !dst || n < 1 || h < 1 || w < 1) return fail(ctx, DCGANSR_ERR_INVALID, "bad scale_bilinear argument");
  if (dh < h || dw < w) return fail(ctx, DCGANSR_ERR_UNSUPPORTED, "scale_bilinear only enlarges (the reference's use)");
  Arena ar;
  const int64_t ns = (int64_t)n * h * w, nd = (int64_t)n * dh * dw;
  float *dS = ar.f(ns), *dD = ar.f(nd);
  if (!dS || !dD) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dS, src, ns * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_scale_bilinear(ctx->st(), dS, dD, n, h, w, dh, dw);
  return d2h(ctx, dst, dD, nd);
}
// images -> device -> the patches become the staged batch of `slot` (k * nper samples of 1 x patch x patch): the patch
// scripts' real_none without the per-pixel host loop (train-gray-patch.lua:267-275)
int dcgansr_stage_patches(dcgansr_ctx* ctx, dcgansr_net* netD, const float* images_host, int k, int h, int w, int patch, int line, int nper,
                          int stride, int slot) {
  if (!ctx || !netD || !images_host || slot < 0 || slot > 63) return fail(ctx, DCGANSR_ERR_INVALID, "bad stage_patches argument");
  if (int rc = patch_args_ok(ctx, k, h, w, patch, line, nper, stride)) return rc;
  if (netD->in_c != 1 || netD->in_h != patch || netD->in_w != patch)
    return fail(ctx, DCGANSR_ERR_INVALID, "netD input must be 1 x patch x patch");
  CK(ctx, cudaSetDevice(ctx->cfg.device));
  if ((int)ctx->slots.size() <= slot) { ctx->slots.resize(slot + 1, nullptr); ctx->slot_bytes.resize(slot + 1, 0); }
  const size_t pbytes = (size_t)k * nper * patch * patch * sizeof(float), ibytes = (size_t)k * h * w * sizeof(float);
  if (int rc = ensure(ctx, &ctx->slots[slot], &ctx->slot_bytes[slot], pbytes)) return rc;
  if (int rc = ensure(ctx, &ctx->tmp, &ctx->tmp_bytes, ibytes)) return rc;
  CK(ctx, cudaMemcpyAsync(ctx->tmp, images_host, ibytes, cudaMemcpyHostToDevice, ctx->stream));
  k_extract_patches(ctx->st(), ctx->tmp, ctx->slots[slot], k, h, w, patch, line, nper, stride);
  CKLAST(ctx);
  return 0;
}

// ---- evaluation metrics: calPSNR / calSSIM of train-gray-3.lua:143-221 on n single-channel image pairs ----
static int metric_op(dcgansr_ctx* ctx, int which, const float* a, const float* b, float* out, int n, int h, int w) {
  NEED_CTX(ctx);
  if (!a || !b || !out || n <= 0 || h <= 0 || w <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "bad argument");
  Arena ar;
  const int64_t ni = (int64_t)n * h * w;
  float *dA = ar.f(ni), *dB = ar.f(ni), *dO = ar.f(n);
  if (!dA || !dB || !dO) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dA, a, ni * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dB, b, ni * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (which == 0) k_psnr(ctx->st(), dA, dB, dO, n, (int64_t)h * w);
  else k_ssim(ctx->st(), dA, dB, dO, n, h, w);
  return d2h(ctx, out, dO, n);
}
int dcgansr_psnr(dcgansr_ctx* ctx, const float* a, const float* b, float* out, int n, int h, int w) { return metric_op(ctx, 0, a, b, out, n, h, w); }
int dcgansr_ssim(dcgansr_ctx* ctx, const float* a, const float* b, float* out, int n, int h, int w) { return metric_op(ctx, 1, a, b, out, n, h, w); }

int dcgansr_generate(dcgansr_ctx* ctx, dcgansr_net* netG, const float* lr_host, int batch, float* sr_host) {
  if (!ctx || !netG || netG->ctx != ctx) return fail(ctx, DCGANSR_ERR_INVALID, "bad argument");
  return dcgansr_net_forward(netG, lr_host, batch, sr_host);
}

// ---- layer-level ops (parity tests) -----------------------------------------------------------

static int conv_op(dcgansr_ctx* ctx, bool full, int what, const float* a, const float* b, float* outp, int n, int cin, int h,
                   int wd, int cout, int k, int s, int p) {
  NEED_CTX(ctx);
  if (!a || !b || !outp || n <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  ConvPlan plan;
  std::string err = plan.build(full, cin, cout, k, s, p, 0, h, wd);
  if (!err.empty()) return fail(ctx, DCGANSR_ERR_INVALID, err);
  struct Guard { ConvPlan& p; ~Guard() { p.free_device(); } } guard{plan};
  if (int rc = plan.alloc_device(ctx)) return rc;
  Arena ar;
  int64_t xin = (int64_t)n * cin * h * wd, yout = (int64_t)n * cout * plan.Hout * plan.Wout, wn = plan.weight_count();
  float* dX = ar.f(xin);
  float* dY = ar.f(yout);
  float* dW = ar.f(wn);
  if (!dX || !dY || !dW) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  int rc = 0;
  if (what == 0) {          // fwd: a = x, b = w, out = y
    if ((rc = upload_nchw(ctx, a, n, cin, h, wd, dX))) return rc;
    CK(ctx, cudaMemcpyAsync(dW, b, wn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    plan.pack(ctx->st(), dW);
    plan.forward(ctx, dX, dY, n, ACT_NONE, 0.f);
    return download_nchw(ctx, dY, n, cout, plan.Hout, plan.Wout, outp);
  } else if (what == 1) {   // dgrad: a = dy, b = w, out = dx
    if ((rc = upload_nchw(ctx, a, n, cout, plan.Hout, plan.Wout, dY))) return rc;
    CK(ctx, cudaMemcpyAsync(dW, b, wn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    plan.pack(ctx->st(), dW);
    plan.dgrad_run(ctx, dY, dX, n);
    return download_nchw(ctx, dX, n, cin, h, wd, outp);
  } else {                  // wgrad: a = x, b = dy, out = dw
    if ((rc = upload_nchw(ctx, a, n, cin, h, wd, dX))) return rc;
    if ((rc = upload_nchw(ctx, b, n, cout, plan.Hout, plan.Wout, dY))) return rc;
    CK(ctx, cudaMemsetAsync(dW, 0, wn * sizeof(float), ctx->stream));
    size_t sb = plan.wscratch_bytes(n);
    float* sc = (float*)ar.bytes(sb);
    if (!sc) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
    plan.wgrad_run(ctx, dX, dY, dW, n, sc, sb);
    return d2h(ctx, outp, dW, wn);
  }
}
int dcgansr_conv2d_fwd(dcgansr_ctx* ctx, const float* x, const float* w, float* y, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, false, 0, x, w, y, n, cin, h, wd, cout, k, s, p);
}
int dcgansr_conv2d_dgrad(dcgansr_ctx* ctx, const float* dy, const float* w, float* dx, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, false, 1, dy, w, dx, n, cin, h, wd, cout, k, s, p);
}
int dcgansr_conv2d_wgrad(dcgansr_ctx* ctx, const float* x, const float* dy, float* dw, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, false, 2, x, dy, dw, n, cin, h, wd, cout, k, s, p);
}
int dcgansr_fullconv2d_fwd(dcgansr_ctx* ctx, const float* x, const float* w, float* y, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, true, 0, x, w, y, n, cin, h, wd, cout, k, s, p);
}
int dcgansr_fullconv2d_dgrad(dcgansr_ctx* ctx, const float* dy, const float* w, float* dx, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, true, 1, dy, w, dx, n, cin, h, wd, cout, k, s, p);
}
int dcgansr_fullconv2d_wgrad(dcgansr_ctx* ctx, const float* x, const float* dy, float* dw, int n, int cin, int h, int wd, int cout, int k, int s, int p) {
  return conv_op(ctx, true, 2, x, dy, dw, n, cin, h, wd, cout, k, s, p);
}

// Device-resident timing of one convolution op (what: 0 fwd, 1 dgrad, 2 wgrad) at batch n: `iters` launches between two
// CUDA events on the library stream after 2 warm-ups, tensors filled on the device (no host traffic).  Used by
// scripts/bench_layers.py to put every layer of a config against its roofline.
int dcgansr_bench_conv(dcgansr_ctx* ctx, int full, int what, int n, int cin, int h, int wd, int cout, int k, int s, int p,
                       int iters, float* ms_out) {
  NEED_CTX(ctx);
  if (!ms_out || n <= 0 || iters <= 0 || what < 0 || what > 2) return fail(ctx, DCGANSR_ERR_INVALID, "bad argument");
  ConvPlan plan;
  std::string err = plan.build(full != 0, cin, cout, k, s, p, 0, h, wd);
  if (!err.empty()) return fail(ctx, DCGANSR_ERR_INVALID, err);
  struct Guard { ConvPlan& p; ~Guard() { p.free_device(); } } guard{plan};
  if (int rc = plan.alloc_device(ctx)) return rc;
  Arena ar;
  int64_t xin = (int64_t)n * cin * h * wd, yout = (int64_t)n * cout * plan.Hout * plan.Wout, wn = plan.weight_count();
  float *dX = ar.f(xin), *dY = ar.f(yout), *dW = ar.f(wn);
  size_t sb = plan.wscratch_bytes(n);
  float* sc = (float*)ar.bytes(sb);
  if (!dX || !dY || !dW || !sc) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  St st = ctx->st();
  k_fill(st, dX, xin, 0.37f); k_fill(st, dY, yout, -0.21f); k_fill(st, dW, wn, 0.02f);
  plan.pack(st, dW);
  for (int it = -2; it < iters; ++it) {
    if (it == 0) CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (what == 0) plan.forward(ctx, dX, dY, n, ACT_NONE, 0.f);
    else if (what == 1) plan.dgrad_run(ctx, dY, dX, n);
    else plan.wgrad_run(ctx, dX, dY, dW, n, sc, sb);
  }
  CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CK(ctx, cudaEventSynchronize(ctx->ev1));
  CKLAST(ctx);
  float ms = 0.f;
  CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_out = ms / iters;
  return 0;
}

int dcgansr_bn_fwd_train(dcgansr_ctx* ctx, const float* x, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float* y, float* save_mean, float* save_invstd, int n, int c, int h, int wd,
                         float eps, float momentum) {
  NEED_CTX(ctx);
  if (!x || !gamma || !beta || !y || n <= 0 || c <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  Arena ar;
  int64_t E = (int64_t)n * c * h * wd, P = (int64_t)n * h * wd;
  float *dX = ar.f(E), *dY = ar.f(E), *dG = ar.f(c), *dB = ar.f(c), *dM = ar.f(c), *dI = ar.f(c), *dRM = ar.f(c), *dRV = ar.f(c);
  double* part = (double*)ar.bytes((size_t)bn_partial_rows(P, c) * 2 * c * sizeof(double));
  double* sums = (double*)ar.bytes((size_t)2 * c * sizeof(double));
  if (!dX || !dY || !dG || !dB || !dM || !dI || !dRM || !dRV || !part || !sums) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  if (int rc = upload_nchw(ctx, x, n, c, h, wd, dX)) return rc;
  CK(ctx, cudaMemcpyAsync(dG, gamma, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dB, beta, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  bool run = running_mean && running_var;
  if (run) {
    CK(ctx, cudaMemcpyAsync(dRM, running_mean, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(dRV, running_var, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  St st = ctx->st();
  k_bn_stats(st, dX, P, c, part, sums);
  k_bn_finalize(st, sums, c, (double)P, eps, momentum, dM, dI, run ? dRM : nullptr, run ? dRV : nullptr);
  k_bn_apply_act(st, dX, dY, P, c, dG, dB, dM, dI, ACT_NONE, 0.f);
  if (int rc = download_nchw(ctx, dY, n, c, h, wd, y)) return rc;
  if (int rc = d2h(ctx, save_mean, dM, c)) return rc;
  if (int rc = d2h(ctx, save_invstd, dI, c)) return rc;
  if (run) {
    if (int rc = d2h(ctx, running_mean, dRM, c)) return rc;
    if (int rc = d2h(ctx, running_var, dRV, c)) return rc;
  }
  return 0;
}

int dcgansr_bn_bwd(dcgansr_ctx* ctx, const float* x, const float* dy, const float* gamma, const float* save_mean,
                   const float* save_invstd, float* dx, float* dgamma, float* dbeta, int n, int c, int h, int wd) {
  NEED_CTX(ctx);
  if (!x || !dy || !gamma || !save_mean || !save_invstd || n <= 0 || c <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  Arena ar;
  int64_t E = (int64_t)n * c * h * wd, P = (int64_t)n * h * wd;
  float *dX = ar.f(E), *dDY = ar.f(E), *dGm = ar.f(E), *dG = ar.f(c), *dM = ar.f(c), *dI = ar.f(c), *dDG = ar.f(c), *dDB = ar.f(c);
  double* part = (double*)ar.bytes((size_t)bn_partial_rows(P, c) * 2 * c * sizeof(double));
  double* sums = (double*)ar.bytes((size_t)2 * c * sizeof(double));
  if (!dX || !dDY || !dGm || !dG || !dM || !dI || !dDG || !dDB || !part || !sums) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  if (int rc = upload_nchw(ctx, x, n, c, h, wd, dX)) return rc;
  if (int rc = upload_nchw(ctx, dy, n, c, h, wd, dDY)) return rc;
  CK(ctx, cudaMemcpyAsync(dG, gamma, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dM, save_mean, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dI, save_invstd, c * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemsetAsync(dDG, 0, c * sizeof(float), ctx->stream));
  CK(ctx, cudaMemsetAsync(dDB, 0, c * sizeof(float), ctx->stream));
  St st = ctx->st();
  k_bn_bwd_reduce(st, dDY, nullptr, dX, P, c, dG, nullptr, dM, dI, ACT_NONE, 0.f, part, sums);
  k_bn_bwd_param(st, sums, c, dDG, dDB);
  float* dFm = ar.f(2 * c);
  if (!dFm) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  k_bn_bwd_apply(st, dDY, nullptr, dX, dGm, P, c, dG, nullptr, dM, dI, ACT_NONE, 0.f, sums, (double)P, dFm);
  if (dx) if (int rc = download_nchw(ctx, dGm, n, c, h, wd, dx)) return rc;
  if (int rc = d2h(ctx, dgamma, dDG, c)) return rc;
  return d2h(ctx, dbeta, dDB, c);
}

int dcgansr_act_fwd(dcgansr_ctx* ctx, const float* x, float* y, int64_t count, int kind, float negval) {
  NEED_CTX(ctx);
  if (!x || !y || count < 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (count == 0) return 0;
  Arena ar;
  float* d = ar.f(count);
  if (!d) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(d, x, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_act(ctx->st(), d, d, count, kind, negval);
  return d2h(ctx, y, d, count);
}
int dcgansr_act_bwd(dcgansr_ctx* ctx, const float* y, const float* dy, float* dx, int64_t count, int kind, float negval) {
  NEED_CTX(ctx);
  if (!y || !dy || !dx || count < 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  if (count == 0) return 0;
  Arena ar;
  float *a = ar.f(count), *b = ar.f(count);
  if (!a || !b) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(a, y, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(b, dy, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_act_bwd(ctx->st(), a, b, b, count, kind, negval);
  return d2h(ctx, dx, b, count);
}

static int resample_op(dcgansr_ctx* ctx, int what, const float* a, float* outp, int n, int c, int h, int wd) {
  NEED_CTX(ctx);
  if (!a || !outp || n <= 0 || c <= 0 || h <= 0 || wd <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  // what: 0 upnearest fwd (in h x wd -> 2h x 2wd), 1 upnearest bwd (in 2h x 2wd -> h x wd), 2 avgpool2 (in h x wd -> h/2 x wd/2)
  int ih = what == 1 ? 2 * h : h, iw = what == 1 ? 2 * wd : wd;
  int oh = what == 0 ? 2 * h : (what == 1 ? h : h / 2), ow = what == 0 ? 2 * wd : (what == 1 ? wd : wd / 2);
  if (what == 2 && (h % 2 || wd % 2)) return fail(ctx, DCGANSR_ERR_INVALID, "avgpool2 needs even spatial dims");
  Arena ar;
  float *dI = ar.f((int64_t)n * c * ih * iw), *dO = ar.f((int64_t)n * c * oh * ow);
  if (!dI || !dO) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  if (int rc = upload_nchw(ctx, a, n, c, ih, iw, dI)) return rc;
  if (what == 0) k_upnearest_fwd(ctx->st(), dI, dO, n, h, wd, c, 2);
  else if (what == 1) k_upnearest_bwd(ctx->st(), dI, dO, n, h, wd, c, 2);
  else k_avgpool2(ctx->st(), dI, dO, n, h, wd, c);
  return download_nchw(ctx, dO, n, c, oh, ow, outp);
}
int dcgansr_upnearest2_fwd(dcgansr_ctx* ctx, const float* x, float* y, int n, int c, int h, int wd) { return resample_op(ctx, 0, x, y, n, c, h, wd); }
int dcgansr_upnearest2_bwd(dcgansr_ctx* ctx, const float* dy, float* dx, int n, int c, int h, int wd) { return resample_op(ctx, 1, dy, dx, n, c, h, wd); }
int dcgansr_avgpool2_fwd(dcgansr_ctx* ctx, const float* x, float* y, int n, int c, int h, int wd) { return resample_op(ctx, 2, x, y, n, c, h, wd); }

static int loss_op(dcgansr_ctx* ctx, int kind, const float* x, const float* label, int64_t count, float* loss, float* dx) {
  NEED_CTX(ctx);
  if (!x || !label || count <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  Arena ar;
  float *dX = ar.f(count), *dL = ar.f(count), *dD = ar.f(count), *dS = ar.f(4);
  if (!dX || !dL || !dD || !dS) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(dX, x, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dL, label, count * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_loss(ctx->st(), kind, dX, count, dL, 1, 0.f, (double)count, dS, dD);
  if (int rc = d2h(ctx, loss, dS, 1)) return rc;
  return d2h(ctx, dx, dD, count);
}
int dcgansr_bce(dcgansr_ctx* ctx, const float* x, const float* label, int64_t count, float* loss, float* dx) { return loss_op(ctx, LOSS_BCE, x, label, count, loss, dx); }
int dcgansr_mse(dcgansr_ctx* ctx, const float* x, const float* label, int64_t count, float* loss, float* dx) { return loss_op(ctx, LOSS_MSE, x, label, count, loss, dx); }

int dcgansr_pixel_mse_per_sample(dcgansr_ctx* ctx, const float* real, const float* fake, float* out, int n, int64_t per_sample, float div) {
  NEED_CTX(ctx);
  if (!real || !fake || !out || n <= 0 || per_sample <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  Arena ar;
  float *a = ar.f(n * per_sample), *b = ar.f(n * per_sample), *o = ar.f(n);
  if (!a || !b || !o) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  CK(ctx, cudaMemcpyAsync(a, real, n * per_sample * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(b, fake, n * per_sample * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  k_pixel_mse(ctx->st(), a, b, o, n, per_sample, div);
  return d2h(ctx, out, o, n);
}

int dcgansr_adam_step(dcgansr_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count, int64_t t, double lr,
                      double beta1, double beta2, double eps) {
  NEED_CTX(ctx);
  if (!p || !g || !m || !v || count <= 0) return fail(ctx, DCGANSR_ERR_INVALID, "null argument");
  Arena ar;
  float *dp = ar.f(count), *dg = ar.f(count), *dm = ar.f(count), *dv = ar.f(count), *ds = ar.f(4);
  int64_t* dt = (int64_t*)ar.bytes(sizeof(int64_t));
  if (!dp || !dg || !dm || !dv || !ds || !dt) return fail(ctx, DCGANSR_ERR_NOMEM, "device allocation failed");
  size_t nb = count * sizeof(float);
  CK(ctx, cudaMemcpyAsync(dp, p, nb, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dg, g, nb, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dm, m, nb, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dv, v, nb, cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dt, &t, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  k_adam_prep(ctx->st(), dt, ds, lr, beta1, beta2);
  k_adam(ctx->st(), dp, dg, dm, dv, count, ds, beta1, beta2, eps);
  if (int rc = d2h(ctx, p, dp, count)) return rc;
  if (int rc = d2h(ctx, m, dm, count)) return rc;
  return d2h(ctx, v, dv, count);
}

}  // extern "C"
#pragma GCC visibility pop
