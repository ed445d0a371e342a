"""Layer-spec builders: the netG / netD graphs of the reference scripts, as plain dict lists.

The same spec list feeds the product (``nn.Sequential.from_specs`` -> ``dcgansr_layer`` records)
and the test oracle, so parity tests build both sides from one description.
FC = SpatialFullConvolution(.,.,4,4,2,2,1,1), C = SpatialConvolution(.,.,4,4,2,2,1,1).
"""
from __future__ import annotations


def _fc(cin, cout, k=4, s=2, p=1):
    return dict(kind="fullconv", cin=cin, cout=cout, k=k, s=s, p=p)


def _c(cin, cout, k=4, s=2, p=1):
    return dict(kind="conv", cin=cin, cout=cout, k=k, s=s, p=p)


def _bn(c):
    return dict(kind="bn", c=c)


RELU = dict(kind="relu")
LRELU = dict(kind="lrelu", negval=0.2)


def dcgan64_D(nc, ndf):
    """netD of train.lua:119-136 (and train-gray*.lua): DCGAN-64 discriminator."""
    return [_c(nc, ndf), dict(LRELU),
            _c(ndf, ndf * 2), _bn(ndf * 2), dict(LRELU),
            _c(ndf * 2, ndf * 4), _bn(ndf * 4), dict(LRELU),
            _c(ndf * 4, ndf * 8), _bn(ndf * 8), dict(LRELU),
            _c(ndf * 8, 1, k=4, s=1, p=0), dict(kind="sigmoid"), dict(kind="view")]


def patch_D(ndf, nc=1):
    """patch discriminator of train-gray-patch.lua:94-108: 3x3, 3x3, 3x3, 2x2 valid convs."""
    return [_c(nc, ndf, k=3, s=1, p=0), dict(LRELU),
            _c(ndf, ndf * 2, k=3, s=1, p=0), _bn(ndf * 2), dict(LRELU),
            _c(ndf * 2, ndf * 4, k=3, s=1, p=0), _bn(ndf * 4), dict(LRELU),
            _c(ndf * 4, 1, k=2, s=1, p=0), dict(kind="sigmoid"), dict(kind="view")]


def train_lua_G(nc, ngf):
    """netG of train.lua:97-113."""
    return [_fc(nc, ngf * 8), _bn(ngf * 8), dict(RELU),
            _fc(ngf * 8, ngf * 4), _bn(ngf * 4), dict(RELU),
            _fc(ngf * 4, ngf * 2), _bn(ngf * 2), dict(RELU),
            _c(ngf * 2, ngf), _bn(ngf), dict(LRELU),
            _c(ngf, nc), dict(kind="tanh")]


def train_gray_G(ngf, nc=1):
    """netG of train-gray.lua:102-117: no BN, no inner activations."""
    return [dict(kind="upnearest", scale=2), _fc(nc, ngf * 4), _fc(ngf * 4, ngf * 2), _c(ngf * 2, ngf), _c(ngf, nc),
            dict(kind="tanh")]


def train_gray_2_G(ngf, nc=1):
    """netG of train-gray-2.lua:65-76: three nearest up-samplings then two stride-2 convs."""
    up = dict(kind="upnearest", scale=2)
    return [dict(up), dict(up), dict(up), _c(nc, ngf), _bn(ngf), dict(RELU), _c(ngf, nc), dict(kind="sigmoid")]


def train_gray_3_G(ngf, nc=1):
    """netG of train-gray-3.lua:52-73 == train-gray-patch.lua:54-75 == ...overlap.lua:76-102."""
    return [dict(kind="upnearest", scale=2),
            _fc(nc, ngf * 4), _bn(ngf * 4), dict(RELU),
            _fc(ngf * 4, ngf * 2), _bn(ngf * 2), dict(RELU),
            _fc(ngf * 2, ngf), _bn(ngf), dict(RELU),
            _c(ngf, ngf * 2), _bn(ngf * 2), dict(RELU),
            _c(ngf * 2, ngf * 4), _bn(ngf * 4), dict(RELU),
            _c(ngf * 4, nc), dict(kind="sigmoid")]


def patch_batch_G(ngf, nc=1):
    """netG of train-gray-patch-batch.lua:55-77: a 4th full-conv instead of the nearest up-sampling."""
    return [_fc(nc, ngf * 8), _bn(ngf * 8), dict(RELU),
            _fc(ngf * 8, ngf * 4), _bn(ngf * 4), dict(RELU),
            _fc(ngf * 4, ngf * 2), _bn(ngf * 2), dict(RELU),
            _fc(ngf * 2, ngf), _bn(ngf), dict(RELU),
            _c(ngf, ngf * 2), _bn(ngf * 2), dict(RELU),
            _c(ngf * 2, ngf * 4), _bn(ngf * 4), dict(RELU),
            _c(ngf * 4, nc), dict(kind="sigmoid")]


# ---- the BASELINE.json configurations (SURVEY.md 8(d) "shape decisions") -------------------------
# name -> dict(G=specs, D=specs, nc, hr (D input size), batch, step=dict(...), data_range)
def _bce_step():
    return dict(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0, pixel_label=False, pixel_div=1.0)


def config(name: str):
    """Return the named workload.  C1a/C2/C3a are reference-exact geometries; C1b/C3b/C4/C5 are the
    BASELINE-worded scale-ups of the same graphs."""
    if name == "C1a":   # train-gray-patch.lua defaults: 64 x (1x8x8), ngf 16, ndf 64, BCE
        return dict(G=train_gray_3_G(16), D=patch_D(64), nc=1, hr=8, batch=64, step=_bce_step(), data_range=(0.0, 1.0))
    if name == "C1b":   # BASELINE configs[0]: 32x32 patches, ngf = ndf = 64
        return dict(G=train_gray_3_G(64), D=patch_D(64), nc=1, hr=32, batch=64, step=_bce_step(), data_range=(0.0, 1.0))
    if name == "C2":    # BASELINE configs[1]: train-gray.lua, 64x64 gray, batch 64, MSE family
        return dict(G=train_gray_G(16), D=dcgan64_D(1, 64), nc=1, hr=64, batch=64,
                    step=dict(family="mse", real_label=0.001, fake_label=0.0, gen_label=0.0, pixel_label=True,
                              pixel_div=64.0 * 64.0), data_range=(-1.0, 1.0))
    if name in ("C3a", "C3b"):   # train.lua: RGB, MSE family, pixel divisor 4*C*H*W (train.lua:194)
        hr = 64 if name == "C3a" else 128
        return dict(G=train_lua_G(3, 12), D=dcgan64_D(3, 64), nc=3, hr=hr, batch=128,
                    step=dict(family="mse", real_label=0.0, fake_label=0.0, gen_label=0.0, pixel_label=True,
                              pixel_div=4.0 * 3 * hr * hr), data_range=(-1.0, 1.0))
    if name == "C4":    # ...overlap.lua training step: 32x32 patches, ngf 16, ndf 64, B = 512 global
        return dict(G=train_gray_3_G(16), D=patch_D(64), nc=1, hr=32, batch=512, step=_bce_step(), data_range=(0.0, 1.0))
    if name == "C4a":   # same with the reference's own 8x8 patches
        return dict(G=train_gray_3_G(16), D=patch_D(64), nc=1, hr=8, batch=512, step=_bce_step(), data_range=(0.0, 1.0))
    if name == "C5":    # scaled synthetic RGB 128 -> 256, ngf = ndf = 128
        # 128 samples of this generator's activations (244 GiB fp32) do not fit one B200: netG is created for 16-sample
        # micro-batches and the library runs the batch with exact whole-batch BatchNorm (dcgansr.cu:net_forward_mb)
        return dict(G=train_lua_G(3, 128), D=dcgan64_D(3, 128), nc=3, hr=256, batch=128, g_microbatch=16,
                    step=dict(family="mse", real_label=0.0, fake_label=0.0, gen_label=0.0, pixel_label=True,
                              pixel_div=4.0 * 3 * 256 * 256), data_range=(-1.0, 1.0))
    raise KeyError(name)


def conv_flops(specs, in_c, in_h, in_w, batch):
    """Forward FLOPs per conv layer: [(index, flops)], SURVEY.md 8(d) formulae."""
    out = []
    c, h, w = in_c, in_h, in_w
    for i, s in enumerate(specs):
        k = s["kind"]
        if k == "conv":
            ho = (h + 2 * s["p"] - s["k"]) // s["s"] + 1
            wo = (w + 2 * s["p"] - s["k"]) // s["s"] + 1
            out.append((i, 2.0 * batch * ho * wo * s["cout"] * s["cin"] * s["k"] ** 2))
            c, h, w = s["cout"], ho, wo
        elif k == "fullconv":
            out.append((i, 2.0 * batch * h * w * s["cin"] * s["cout"] * s["k"] ** 2))
            h = (h - 1) * s["s"] - 2 * s["p"] + s["k"] + s.get("adj", 0)
            w = (w - 1) * s["s"] - 2 * s["p"] + s["k"] + s.get("adj", 0)
            c = s["cout"]
        elif k == "upnearest":
            h *= s.get("scale", 2)
            w *= s.get("scale", 2)
    return out


def step_flops(cfg, batch=None):
    """Algorithmic FLOPs of one training step: 7 F_D - 2 F_D,L1 + 3 F_G - F_G,L1."""
    b = batch or cfg["batch"]
    fd = conv_flops(cfg["D"], cfg["nc"], cfg["hr"], cfg["hr"], b)
    fg = conv_flops(cfg["G"], cfg["nc"], cfg["hr"] // 2, cfg["hr"] // 2, b)
    FD, FG = sum(f for _, f in fd), sum(f for _, f in fg)
    return 7 * FD - 2 * fd[0][1] + 3 * FG - fg[0][1]
