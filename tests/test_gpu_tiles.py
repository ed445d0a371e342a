"""Per-layer parity at BASELINE spatial sizes: every persistent CTA of the halo-tile kernels walks several tiles (>= 4 x 148
tiles per launch), so the stage / accumulator ring wrap and the TMA-store staging-buffer reuse are checked per layer, not only
through the step tests.  Strict fp32 <= 1e-5 and FAST_TF32 <= 2e-3 (max-norm) against the float64 oracle."""
import numpy as np
import pytest

from dcgan_super_resolution_b200 import _lib as L
from oracle import ops
from util import FAST_TOL, STRICT_TOL, ptr, rel_err, rng, t64

pytestmark = pytest.mark.gpu

# (kind, n, cin, h, w, cout, k, s, p) -- tiles are 16 x 8 pixels of the iterated grid (per sub-pixel class group)
BIG_SHAPES = [
    ("full", 8, 64, 128, 128, 32, 4, 2, 1),     # C2 G layer 3 (FC 64->32, 128^2 -> 256^2): 1024 tiles; dgrad = two cout slices
    ("conv", 8, 32, 256, 256, 16, 4, 2, 1),     # C2 G layer 4 (C 32->16, 256^2 -> 128^2): dgrad through the TMA-store epilogue
    ("full", 4, 48, 256, 256, 24, 4, 2, 1),     # train.lua G layer 3 at C3b size (FC 48->24, 256^2 -> 512^2): 96-byte output rows
    ("conv", 4, 24, 512, 512, 12, 4, 2, 1),     # train.lua G layer 4 at C3b size (C 24->12): fwd 48-byte rows, dgrad 96-byte rows
    ("full", 6, 96, 128, 128, 48, 4, 2, 1),     # train.lua G layer 2 at C3b size (FC 96->48): one launch per sub-pixel class
]


@pytest.mark.parametrize("mode", ["strict", "tf32"])
@pytest.mark.parametrize("shape", BIG_SHAPES)
def test_big_layer_fwd_dgrad_wgrad(ctx, ctx_fast, shape, mode):
    c = ctx if mode == "strict" else ctx_fast
    tol = STRICT_TOL if mode == "strict" else FAST_TOL
    kind, n, cin, h, w, cout, k, s, p = shape
    full = kind == "full"
    r = rng(hash(shape[1:]) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (0.1 * r.standard_normal((cin, cout, k, k) if full else (cout, cin, k, k))).astype(np.float32)
    ho, wo = ((h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hc = c.lib, c.h
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    args = (n, cin, h, w, cout, k, s, p)
    X, W, DY = t64(x), t64(wt), t64(dy)
    y = np.empty((n, cout, ho, wo), np.float32)
    L.check(getattr(lib, pre + "fwd")(hc, ptr(x), ptr(wt), ptr(y), *args), hc)
    ref = (ops.fullconv2d_fwd(X, W, s, p) if full else ops.conv2d_fwd(X, W, s, p)).numpy()
    assert rel_err(y, ref) <= tol, ("fwd", rel_err(y, ref))
    # every image and every tile row must be right, not only the largest element: per-image check as well
    for i in range(n):
        assert rel_err(y[i], ref[i]) <= 2 * tol, ("fwd image", i)
    del y, ref
    dx = np.empty_like(x)
    L.check(getattr(lib, pre + "dgrad")(hc, ptr(dy), ptr(wt), ptr(dx), *args), hc)
    ref = (ops.fullconv2d_dgrad(DY, W, s, p) if full else ops.conv2d_dgrad(DY, W, x.shape, s, p)).numpy()
    assert rel_err(dx, ref) <= tol, ("dgrad", rel_err(dx, ref))
    for i in range(n):
        assert rel_err(dx[i], ref[i]) <= 2 * tol, ("dgrad image", i)
    del dx, ref
    dw = np.empty_like(wt)
    L.check(getattr(lib, pre + "wgrad")(hc, ptr(x), ptr(dy), ptr(dw), *args), hc)
    ref = (ops.fullconv2d_wgrad(X, DY, wt.shape, s, p) if full else ops.conv2d_wgrad(X, DY, wt.shape, s, p)).numpy()
    assert rel_err(dw, ref) <= tol, ("wgrad", rel_err(dw, ref))


# grids of >= 2 x 148 CTAs on the per-tap kernel: thread-block clusters of 2 / 4 CTAs share every weight tile by TMA multicast
CLUSTER_SHAPES = [
    ("conv", 40, 64, 64, 64, 128, 4, 2, 1),     # D layer 2 (train.lua:124) at 64^2: TMA 2-D weight boxes (stride-2 parity view), dgrad 4 classes
    ("full", 16, 128, 32, 32, 64, 4, 2, 1),     # C1b G layer 3 shape class: pre-tiled weight images (1-D bulk multicast)
    ("conv", 37, 128, 32, 32, 256, 4, 2, 1),    # two N tiles, odd tile count (the grid is padded to whole clusters)
]


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("shape", CLUSTER_SHAPES)
def test_per_tap_kernel_clusters(ctx_fast, shape, cluster, monkeypatch):
    monkeypatch.setenv("DCGANSR_TC_CLUSTER", str(cluster))
    monkeypatch.setenv("DCGANSR_NO_HALO", "1")          # keep these shapes on the per-tap kernel
    kind, n, cin, h, w, cout, k, s, p = shape
    full = kind == "full"
    r = rng(hash(shape[1:]) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (0.1 * r.standard_normal((cin, cout, k, k) if full else (cout, cin, k, k))).astype(np.float32)
    ho, wo = ((h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hc = ctx_fast.lib, ctx_fast.h
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    args = (n, cin, h, w, cout, k, s, p)
    X, W, DY = t64(x), t64(wt), t64(dy)
    y = np.empty((n, cout, ho, wo), np.float32)
    L.check(getattr(lib, pre + "fwd")(hc, ptr(x), ptr(wt), ptr(y), *args), hc)
    ref = (ops.fullconv2d_fwd(X, W, s, p) if full else ops.conv2d_fwd(X, W, s, p)).numpy()
    assert rel_err(y, ref) <= FAST_TOL, ("fwd", rel_err(y, ref))
    for i in range(0, n, 7):
        assert rel_err(y[i], ref[i]) <= 2 * FAST_TOL, ("fwd image", i)
    dx = np.empty_like(x)
    L.check(getattr(lib, pre + "dgrad")(hc, ptr(dy), ptr(wt), ptr(dx), *args), hc)
    ref = (ops.fullconv2d_dgrad(DY, W, s, p) if full else ops.conv2d_dgrad(DY, W, x.shape, s, p)).numpy()
    assert rel_err(dx, ref) <= FAST_TOL, ("dgrad", rel_err(dx, ref))
    for i in range(0, n, 7):
        assert rel_err(dx[i], ref[i]) <= 2 * FAST_TOL, ("dgrad image", i)


# the persistent wide-tile per-tap kernel (kernels_tc2.cu): 256 x 128 and 128 x 256 work items, several items per CTA (stage ring
# and TMEM double-buffer wrap), odd tile counts (a pixel group whose second tile is past the end), stride-1 and stride-2 gathers
WIDE_SHAPES = [
    ("conv", 75, 64, 64, 64, 128, 4, 2, 1),     # D layer 2 (train.lua:124): 2 pixel tiles x 128 couts, 300 items; dgrad 4 classes x 64 couts
    ("conv", 37, 128, 32, 32, 256, 4, 2, 1),    # D layer 3: 128 pixels x 256 couts, odd tile counts; dgrad 2 x 128
    ("conv", 36, 256, 16, 16, 512, 4, 2, 1),    # D layer 4: two 256-cout tiles per pixel tile; dgrad 128 x 256
    ("full", 19, 128, 32, 32, 64, 4, 2, 1),     # C1b G layer 3 shape class (FC 128->64): 64-cout images; dgrad 16 taps, stride-2 gather
    ("conv", 11, 64, 30, 30, 128, 3, 1, 0),     # C1b patch-D layer 2 (train-gray-patch.lua:97): 9 taps, 28 x 28 output (partial tiles)
]


@pytest.mark.parametrize("mode", ["1", "2", "2p"])
@pytest.mark.parametrize("shape", WIDE_SHAPES)
def test_wide_tile_kernel(ctx_fast, shape, mode, monkeypatch):
    # 1: the production policy, 2: every geometry the kernel can run (single-CTA items only), 2p: the same with the 256-cout
    # items forced onto CTA pairs (cta_group::2)
    monkeypatch.setenv("DCGANSR_TC2_PAIR", "2" if mode == "2p" else ("0" if mode == "2" else "1"))
    mode = mode[0]
    monkeypatch.setenv("DCGANSR_TC2", mode)
    monkeypatch.setenv("DCGANSR_TC3", "0")             # (the halo-tile pair kernel would take most of these shapes first)
    kind, n, cin, h, w, cout, k, s, p = shape
    full = kind == "full"
    r = rng(hash(shape[1:]) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (0.1 * r.standard_normal((cin, cout, k, k) if full else (cout, cin, k, k))).astype(np.float32)
    ho, wo = ((h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hc = ctx_fast.lib, ctx_fast.h
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    args = (n, cin, h, w, cout, k, s, p)
    X, W, DY = t64(x), t64(wt), t64(dy)
    ctx_fast.profile_begin()
    y = np.empty((n, cout, ho, wo), np.float32)
    L.check(getattr(lib, pre + "fwd")(hc, ptr(x), ptr(wt), ptr(y), *args), hc)
    dx = np.empty_like(x)
    L.check(getattr(lib, pre + "dgrad")(hc, ptr(dy), ptr(wt), ptr(dx), *args), hc)
    names = [k_["name"] for k_ in ctx_fast.profile_end()]
    if mode == "2":
        assert ("tapconv_tc2" in names or "tapconv_tc2_pair" in names) and "tapconv_tc" not in names, names
    ref = (ops.fullconv2d_fwd(X, W, s, p) if full else ops.conv2d_fwd(X, W, s, p)).numpy()
    assert rel_err(y, ref) <= FAST_TOL, ("fwd", rel_err(y, ref))
    for i in range(n):
        assert rel_err(y[i], ref[i]) <= 2 * FAST_TOL, ("fwd image", i)
    ref = (ops.fullconv2d_dgrad(DY, W, s, p) if full else ops.conv2d_dgrad(DY, W, x.shape, s, p)).numpy()
    assert rel_err(dx, ref) <= FAST_TOL, ("dgrad", rel_err(dx, ref))
    for i in range(n):
        assert rel_err(dx[i], ref[i]) <= 2 * FAST_TOL, ("dgrad image", i)


# the halo-tile pair kernel (kernels_tc3.cu): 8 x 16 pixel tiles with a halo, every tap a shifted window, streamed weight halves,
# cta_group::2.  Shapes: 2 x 2-tap sub-pixel classes (full-conv forward / conv dgrad), the 16-tap stride-2 gather through the four
# parity planes (conv forward / full-conv dgrad), 3 x 3 stride-1 taps, odd tile counts, partial tiles, 64 / 128 / 256 couts
HALO_PAIR_SHAPES = [
    ("full", 5, 256, 32, 32, 128, 4, 2, 1),     # C1b G layer 2 shape class (FC 256->128): fwd 4 classes x 128 couts; dgrad 256 couts, 4 planes
    ("conv", 7, 64, 64, 64, 128, 4, 2, 1),      # D layer 2 (train.lua:124): fwd through the parity planes, dgrad 4 classes x 64 couts
    ("conv", 9, 128, 32, 48, 256, 4, 2, 1),     # D layer 3, non-square: dgrad 4 classes x 128 couts on 16 x 24 class grids
    ("conv", 6, 64, 30, 30, 128, 3, 1, 0),      # patch-D layer 2 (train-gray-patch.lua:97): 9 taps, 28 x 28 grid (partial tiles)
    ("full", 3, 128, 24, 40, 64, 4, 2, 1),      # 64-cout images (32-row weight halves)
]


@pytest.mark.parametrize("shape", HALO_PAIR_SHAPES)
def test_halo_pair_kernel(ctx_fast, shape, monkeypatch):
    monkeypatch.setenv("DCGANSR_TC3", "2")             # every geometry the kernel can run
    monkeypatch.setenv("DCGANSR_NO_HALO_PER_CLASS", "1")
    kind, n, cin, h, w, cout, k, s, p = shape
    full = kind == "full"
    r = rng(hash(shape[1:]) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wt = (0.1 * r.standard_normal((cin, cout, k, k) if full else (cout, cin, k, k))).astype(np.float32)
    ho, wo = ((h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hc = ctx_fast.lib, ctx_fast.h
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    args = (n, cin, h, w, cout, k, s, p)
    X, W, DY = t64(x), t64(wt), t64(dy)
    ctx_fast.profile_begin()
    y = np.empty((n, cout, ho, wo), np.float32)
    L.check(getattr(lib, pre + "fwd")(hc, ptr(x), ptr(wt), ptr(y), *args), hc)
    dx = np.empty_like(x)
    L.check(getattr(lib, pre + "dgrad")(hc, ptr(dy), ptr(wt), ptr(dx), *args), hc)
    prof = {}
    for k_ in ctx_fast.profile_end():
        prof[k_["name"]] = prof.get(k_["name"], 0) + k_["launches"]
    assert prof.get("tapconv_tc3", 0) == 2 and "tapconv_tc" not in prof and "tapconv_tc2" not in prof, prof
    ref = (ops.fullconv2d_fwd(X, W, s, p) if full else ops.conv2d_fwd(X, W, s, p)).numpy()
    assert rel_err(y, ref) <= FAST_TOL, ("fwd", rel_err(y, ref))
    for i in range(n):
        assert rel_err(y[i], ref[i]) <= 2 * FAST_TOL, ("fwd image", i)
    ref = (ops.fullconv2d_dgrad(DY, W, s, p) if full else ops.conv2d_dgrad(DY, W, x.shape, s, p)).numpy()
    assert rel_err(dx, ref) <= FAST_TOL, ("dgrad", rel_err(dx, ref))
    for i in range(n):
        assert rel_err(dx[i], ref[i]) <= 2 * FAST_TOL, ("dgrad image", i)


# accGradParameters on CTA pairs (kernels_tc2.cu:wgrad_tc_pair_kernel): 256 cp per MMA, cq slices of 64 / 128 / 256, tap groups,
# stride-1 and stride-2 shifted tensors, several pixel splits
WGRAD_PAIR_SHAPES = [
    ("conv", 12, 128, 32, 32, 256, 4, 2, 1),    # D layer 3 (train.lua:127): P = dy (256), Q = x (128), 16 taps in groups of 4
    ("conv", 10, 256, 16, 16, 512, 4, 2, 1),    # D layer 4: two cp pairs, 256-column cq slice, groups of 2 taps
    ("full", 6, 256, 24, 24, 128, 4, 2, 1),     # C1b G layer 2 shape class (FC 256->128): P = x (256), Q = dy (128) at twice the size
    ("conv", 9, 128, 28, 28, 256, 3, 1, 0),     # patch-D layer 3 (train-gray-patch.lua:100): 9 taps, stride 1
    ("full", 5, 384, 16, 16, 64, 4, 2, 1),      # cp not a multiple of 256 (second pair half empty), 64-column cq slice
]


@pytest.mark.parametrize("shape", WGRAD_PAIR_SHAPES)
def test_wgrad_pair_kernel(ctx_fast, shape, monkeypatch):
    monkeypatch.setenv("DCGANSR_WGRAD_PAIR", "2")
    monkeypatch.setenv("DCGANSR_NO_WGRAD_HALO", "1")
    kind, n, cin, h, w, cout, k, s, p = shape
    full = kind == "full"
    r = rng(hash(shape[1:]) % 2**31)
    x = r.standard_normal((n, cin, h, w)).astype(np.float32)
    wshape = (cin, cout, k, k) if full else (cout, cin, k, k)
    ho, wo = ((h - 1) * s - 2 * p + k, (w - 1) * s - 2 * p + k) if full else ((h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
    dy = r.standard_normal((n, cout, ho, wo)).astype(np.float32)
    lib, hc = ctx_fast.lib, ctx_fast.h
    pre = "dcgansr_fullconv2d_" if full else "dcgansr_conv2d_"
    dw = np.empty(wshape, np.float32)
    ctx_fast.profile_begin()
    L.check(getattr(lib, pre + "wgrad")(hc, ptr(x), ptr(dy), ptr(dw), n, cin, h, w, cout, k, s, p), hc)
    names = [k_["name"] for k_ in ctx_fast.profile_end()]
    assert "wgrad_tc_pair" in names and "wgrad_tc" not in names, names
    X, DY = t64(x), t64(dy)
    ref = (ops.fullconv2d_wgrad(X, DY, wshape, s, p) if full else ops.conv2d_wgrad(X, DY, wshape, s, p)).numpy()
    assert rel_err(dw, ref) <= FAST_TOL, ("wgrad", rel_err(dw, ref))
