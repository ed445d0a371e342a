// common.h -- internal declarations shared by the translation units of libdcgansr.so.
// Nothing here is part of the C ABI (include/dcgansr.h is).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#define DSR_MAX_TAPS 32

// Stream handle + launch accounting (gpu_launches in bench.py is read from here) + the optional
// per-launch event profiler (dcgansr_profile_begin/end): every launcher reports its kernel name and
// the ALGORITHMIC work of the launch (flops for the convolution GEMMs, bytes for bandwidth kernels).
#include <vector>
enum { WORK_FLOPS = 0, WORK_BYTES = 1 };
struct ProfRec { const char* name; double work; int kind; };
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> pool;     // pool[0] = begin marker, pool[i+1] recorded after launch i
  std::vector<ProfRec> recs;
};
// per-context device workspace of the tensor-core kernels (split-K partial tiles + self-resetting arrival counters)
struct TcWorkspace {
  float* part = nullptr; size_t part_bytes = 0; int* counters = nullptr; int ncounters = 0;
  double* lpart = nullptr; int* lcounter = nullptr;     // criterion kernel: block partials (256 doubles) + arrival counter (both precisions)
};
struct St {
  cudaStream_t s;
  int64_t* launches;
  Prof* prof;
  TcWorkspace* ws = nullptr;
};
static inline void dsr_launched(const St& st, const char* name, double work, int kind) {
  if (st.launches) ++*st.launches;
  if (st.prof && st.prof->on) {
    Prof* p = st.prof;
    size_t i = p->recs.size() + 1;
    if (i >= p->pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      p->pool.push_back(e);
    }
    cudaEventRecord(p->pool[i], st.s);
    p->recs.push_back(ProfRec{name, work, kind});
  }
}
#define DSR_LAUNCHED(st, name, work, kind) dsr_launched((st), (name), (double)(work), (kind))

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };
enum { LOSS_BCE = 0, LOSS_MSE = 1 };

// ---------------------------------------------------------------------------------------
// Generic "tap-list" convolution geometry (one sub-pixel class per launch).
//
//   out[n, gy*so + oy0, gx*so + ox0, co] = act( sum_t sum_ci in[n, gy*si + dy[t], gx*si + dx[t], ci] * W[t][ci][co] )
//
// for (gy,gx) in [0,Hg)x[0,Wg); reads outside the input are zero.  It covers
//   conv fwd            (si = stride, so = 1, taps = k*k,          dy = ky - p)
//   full-conv dgrad     (same form with dy as the input)
//   full-conv fwd       (si = 1, so = stride, one launch per output parity class)
//   conv dgrad          (same sub-pixel form with dy as the input)
// ---------------------------------------------------------------------------------------
struct TapGeom {
  int N, Hi, Wi, Ci;        // input tensor NHWC
  int Ho, Wo, Co;           // output tensor NHWC (full dims)
  int Hg, Wg;               // iterated grid for this class
  int si, so, oy0, ox0;
  int ntaps;
  int dy[DSR_MAX_TAPS], dx[DSR_MAX_TAPS];
};

// ---------------------------------------------------------------------------------------
// Generic weight-gradient geometry:
//   acc[t][cp][cq] = sum_{n,gy,gx} P[n,gy,gx,cp] * Q[n, gy*s + dy[t], gx*s + dx[t], cq]
// conv:      P = dy (Cp = cout), Q = x  (Cq = cin), master index cp*(Cq*T) + cq*T + t = w[co][ci][ky][kx]
// full-conv: P = x  (Cp = cin),  Q = dy (Cq = cout),                              = w[ci][co][ky][kx]
// ---------------------------------------------------------------------------------------
struct WgradGeom {
  int N, Hp, Wp, Cp;        // grid tensor
  int Hq, Wq, Cq;           // shifted tensor
  int s;
  int ntaps;
  int dy[DSR_MAX_TAPS], dx[DSR_MAX_TAPS];
};

// ---- kernels_bw.cu : bandwidth kernels ----------------------------------------------------
void k_nchw_to_nhwc(St st, const float* in, float* out, int N, int C, int H, int W);
void k_nhwc_to_nchw(St st, const float* in, float* out, int N, int C, int H, int W);
// BN forward statistics.  partials: bn_partial_rows(P,C) x 2C doubles; sums[2C] = (sum x, sum x^2).
int  bn_partial_rows(int64_t P, int C);
void k_bn_stats(St st, const float* x, int64_t P, int C, double* partials, double* sums);
void k_bn_finalize(St st, const double* sums, int C, double n_total, float eps, float momentum,
                   float* save_mean, float* save_invstd, float* running_mean, float* running_var);
void k_bn_apply_act(St st, const float* x, float* y, int64_t P, int C, const float* gamma, const float* beta,
                    const float* mean, const float* invstd, int act, float negval);
void k_act(St st, const float* x, float* y, int64_t count, int act, float negval);
void k_act_bwd(St st, const float* y, const float* dy, float* dx, int64_t count, int act, float negval);
// BN backward: g = dy*act'(.) is re-derived per element in both passes (nothing is written but the sums / dx).  y == nullptr
// (allowed for ACT_NONE / RELU / LRELU): the mask is recomputed from x with the CURRENT gamma / beta -- only valid while the
// parameters are those of the cached forward; otherwise pass the cached activation output y.  sums[2C] = (sum g, sum g*xhat)
void k_bn_bwd_reduce(St st, const float* dy, const float* y, const float* x, int64_t P, int C, const float* gamma, const float* beta,
                     const float* mean, const float* invstd, int act, float negval, double* partials, double* sums);
// dbeta += sums_local[0..C), dgamma += sums_local[C..2C)   (either pointer may be null)
void k_bn_bwd_param(St st, const double* sums_local, int C, float* dgamma, float* dbeta);
void k_bn_bwd_apply(St st, const float* dy, const float* y, const float* x, float* dx, int64_t P, int C, const float* gamma,
                    const float* beta, const float* mean, const float* invstd, int act, float negval, const double* sums_total,
                    double n_total, float* fmeans /* scratch, 2C floats */);
// grouped + fused BatchNorm of the training step (no cross-rank statistics): see kernels_bw.cu
void k_bn_fwd_grouped(St st, const float* x, float* y, int64_t P, int C, int groups, const float* gamma, const float* beta,
                      float* save_mean, float* save_invstd, int64_t sstride, float* running_mean, float* running_var, float eps,
                      float momentum, int act, float negval, double* partials);
void k_bn_bwd_grouped(St st, const float* dy, const float* y, const float* x, float* dx, int64_t P, int C, int groups,
                      const float* gamma, const float* beta, const float* save_mean, const float* save_invstd, int64_t sstride, int act,
                      float negval, double* partials, double* sums, float* dgamma, float* dbeta, float* fmeans /* groups x 2C floats */);
void k_upnearest_fwd(St st, const float* x, float* y, int N, int H, int W, int C, int scale);
void k_upnearest_bwd(St st, const float* dy, float* dx, int N, int H, int W, int C, int scale);
void k_extract_patches(St st, const float* img, float* patches, int K, int H, int W, int p, int line, int nper, int stride);
void k_assemble_patches(St st, const float* patches, float* img, int K, int H, int W, int p, int line, int nper, int stride);
bool stitch_overlap_supported(int p, int L, int ov);
void k_scale_bilinear(St st, const float* src, float* dst, int N, int H, int W, int DH, int DW);
void k_stitch_overlap(St st, const float* patches, float* img, int K, int H, int W, int p, int L, int ov, int flags);
void k_psnr(St st, const float* a, const float* b, float* out, int n, int64_t per);
void k_ssim(St st, const float* a, const float* b, float* out, int n, int H, int W);
void k_avgpool2(St st, const float* x, float* y, int N, int H, int W, int C);
// criterion forward+backward fused.  label_vec (per-sample, `per` outputs each) or constant.
// loss_out[0] = sum/n_total ; dx = dL/dx with 1/n_total scaling.  dx may be null.
void k_loss(St st, int kind, const float* x, int64_t count, const float* label_vec, int64_t per, float label_const,
            double n_total, float* loss_out, float* dx);
void k_pixel_mse(St st, const float* real, const float* fake, float* out, int n, int64_t per_sample, float div);
// Adam: step size is read from dev_step[0] (written by k_adam_prep) so the launch is graph-replayable.
void k_adam_prep(St st, int64_t* dev_t, float* dev_step, double lr, double beta1, double beta2);
void k_adam(St st, float* p, const float* g, float* m, float* v, int64_t count, const float* dev_step,
            double beta1, double beta2, double eps);
void k_stage_pack(St st, const float* rmean, const float* rvar, int nbn, const float* losses, float* stage, float inv_world);
void k_stage_unpack(St st, const float* stage, float* rmean, float* rvar, int nbn, float* losses);
void k_dacc(St st, double* dst, const double* src, int n);
void k_fill(St st, float* p, int64_t count, float v);
void k_scale(St st, float* p, int64_t count, float s);
void k_flush(St st, float* buf, int64_t count);

// ---- kernels_peer.cu : one-shot all-reduce over NVLink peer memory (sync_bn statistics) ------
#define PEER_MAXW 8       // ranks of one NVSwitch domain
#define PEER_SLOTS 4
struct PeerAR {
  double* data[PEER_MAXW];               // rank r's receive area [PEER_SLOTS][PEER_MAXW][nmax] (own: local pointer, peers: cudaIpc mappings)
  unsigned long long* flags[PEER_MAXW];  // rank r's flags [PEER_SLOTS][PEER_MAXW]
  unsigned long long* seq;               // calls completed by this rank (device memory: advanced by the kernels, graph-replayable)
  int* err;                              // host-mapped: set when a peer did not arrive within the time-out
  int rank, world, nmax;
};
void k_peer_allreduce(St st, const PeerAR& p, const double* in, double* out, int n);
void k_bn_finalize_peer(St st, const PeerAR& p, double* sums, int C, double n_total, float eps, float momentum, float* save_mean,
                        float* save_invstd, float* running_mean, float* running_var);

// ---- kernels_simt.cu : strict fp32 FFMA convolutions -----------------------------------------
// Packed weights for the tap-list kernels: Wp[t][a][b] = master[a*sa + b*sb + tapidx[t]]
void k_pack_taps(St st, const float* master, float* wp, int ntaps, const int* tapidx_dev, int A, int B,
                 int64_t sa, int64_t sb);
// one fused pack launch per net: job j repacks master weights into dst (tc: K-major TF32 pack, else [t][a][b])
// tc: 0 = [t][a][b] streaming pack, 1 = K-major TF32 pack bp[b][t*A + a] (TMA 2-D source), 2 = the same values pre-tiled and
// pre-swizzled as the [n tile][k block][bn rows][32 floats] shared-memory images of the per-tap kernel (bulk-copy source)
struct PackJob { const float* src; float* dst; const int* tapidx; int ntaps, A, B, tc, bn; int64_t sa, sb, begin; };
void k_pack_all(St st, const PackJob* jobs_dev, int njobs, int64_t total);
void k_tapconv_simt(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float negval);
// wgrad: accumulates into grad_master (+=) through a deterministic split-K reduction in `scratch`
size_t wgrad_simt_scratch_bytes(const WgradGeom& g);
void k_wgrad_simt(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master,
                  float* scratch, size_t scratch_bytes);
// grad_master[cp*(Cq*T) + cq*T + t] += sum_s scratch[s][cp][t*Cq + cq]   (fixed order)
void k_wgrad_reduce(St st, const float* scratch, int S, int Cp, int Cq, int T, float* grad_master);

// ---- kernels_thin.cu : streaming fp32 kernels for layers with a 1..4-channel side -------------------
bool thin_wgrad_supported(const WgradGeom& g);
size_t thin_wgrad_scratch_bytes(const WgradGeom& g);
// false when the geometry is not covered (caller falls through to the GEMM kernels)
bool k_wgrad_thin(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes);

// thin-input (Ci <= 4, all classes in one launch; wp[i] = SIMT pack [t][ci][co]) and thin-output (Co <= 4) convolutions
bool thin_in_supported(const TapGeom* classes, int ncls);
bool k_tapconv_thin_in(St st, const TapGeom* classes, int ncls, const float* const* wp, const float* in, float* out, int act, float neg);
bool thin_out_supported(const TapGeom& g);
bool k_tapconv_thin_out(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float neg);
bool thin_out_px_supported(const TapGeom& g);
bool k_tapconv_thin_out_px(St st, const TapGeom& g, const float* in, const float* wp, float* out, int act, float neg);

// ---- kernels_tc.cu : tcgen05 / TMA / TMEM implicit-GEMM convolutions (FAST_TF32) --------------
bool tc_init(std::string* err);                       // resolves cuTensorMapEncodeTiled
bool tc_tapconv_supported(const TapGeom& g);
// K-major packed weights for the tensor-core path: Bp[b (N, padded)][t*A + a] (tf32-rounded)
size_t tc_packed_elems(int ntaps, int A, int B);
// pre-tiled weight images for k_tapconv_tc_multi (needs A % 32 == 0): tile rows bn = tc_bt_rows(B)
int tc_bt_rows(int B);
size_t tc_bt_elems(int ntaps, int A, int B);
void k_pack_taps_tc(St st, const float* master, float* bp, int ntaps, const int* tapidx_dev, int A, int B,
                    int64_t sa, int64_t sb);
// returns false (with err) if the launch could not be configured
bool k_tapconv_tc(St st, const TapGeom& g, const float* in, const float* bp, float* out, int act, float negval,
                  std::string* err);
// ---- kernels_halo.cu : weights-resident halo-tile kernel (all sub-pixel classes of a module in one launch) ----
// allow_pair = false: only the single-CTA form counts (per-class launches of a class group are not worth a CTA pair: the
// halo-tile pair kernel of kernels_tc3.cu runs those layers faster -- C1b FC 256->128 forward 0.43 ms against 4 x 0.17 ms)
bool halo_tapconv_supported(const TapGeom* classes, int ncls, bool allow_pair = true);
// stats != nullptr (and halo_stats_rows(...) > 0): the epilogue also writes per-CTA BatchNorm partial sums of the output,
// rows [stats_row0, stats_row0 + halo_stats_rows) of a [rows][2*Co] double matrix (sum y | sum y^2)
int halo_stats_rows(const TapGeom* classes, int ncls);
bool k_tapconv_halo(St st, const TapGeom* classes, int ncls, const float* const* bp, const float* in, float* out, int act,
                    float negval, std::string* err, double* stats = nullptr, int stats_row0 = 0);
// BatchNorm forward from partial sums some other kernel produced (nb rows of [2C]): column sums + finalize, then apply
void k_bn_fwd_from_partials(St st, const float* x, float* y, int64_t P, int C, int nb, const float* gamma, const float* beta, float* save_mean,
                            float* save_invstd, float* running_mean, float* running_var, float eps, float momentum, int act, float negval,
                            const double* partials);
// ---- kernels_wgrad_halo.cu : stacked-shift halo-tile wgrad ----
bool wgrad_halo_supported(const WgradGeom& g);
size_t wgrad_halo_scratch_bytes(const WgradGeom& g);
bool k_wgrad_halo(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                  std::string* err);
// all sub-pixel classes of a module in one launch (blockIdx.z); classes must share the iterated grid
bool tc_tapconv_multi_ok(const TapGeom* classes, int ncls);
bool k_tapconv_tc_multi(St st, const TapGeom* classes, int ncls, const float* const* bp, const float* in, float* out, int act,
                        float negval, std::string* err, const float* const* bt = nullptr);
// ---- kernels_tc2.cu : persistent wide-tile per-tap kernel (256 x 128 / 128 x 256 work items, double-buffered TMEM) ----
// takes a class group when it has pre-tiled weight images, fills the machine and wastes little in its last wave
bool tc2_tapconv_supported(const TapGeom* classes, int ncls, const float* const* bt);
bool k_tapconv_tc2(St st, const TapGeom* classes, int ncls, const float* const* bt, const float* in, float* out, int act, float negval,
                   std::string* err);
// ---- kernels_tc3.cu : halo-tile A + streamed weights on CTA pairs (class groups with <= 128 couts) ----
bool tc3_tapconv_supported(const TapGeom* classes, int ncls, const float* const* bt);
bool k_tapconv_tc3(St st, const TapGeom* classes, int ncls, const float* const* bt, const float* in, float* out, int act, float negval,
                   std::string* err);
// wgrad on CTA pairs (kernels_tc2.cu): Cp >= 256, 256 cp x up to 256 cq per MMA
bool wgrad_tc_pair_supported(const WgradGeom& g);
size_t wgrad_tc_pair_scratch_bytes(const WgradGeom& g);
bool k_wgrad_tc_pair(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                     std::string* err);
bool tc_wgrad_supported(const WgradGeom& g);
size_t wgrad_tc_scratch_bytes(const WgradGeom& g);
bool k_wgrad_tc(St st, const WgradGeom& g, const float* P, const float* Q, float* grad_master, float* scratch, size_t scratch_bytes,
                std::string* err);
