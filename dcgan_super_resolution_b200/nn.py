"""Host-side mirror of the Torch7 surface the reference scripts use (train.lua:97-283), on top of the
C ABI of libdcgansr.so.  Same names and argument order as the Lua modules so the parity tests read
like the reference: ``nn.Sequential():add(nn.SpatialFullConvolution(nc, ngf*8, 4, 4, 2, 2, 1, 1))``.

Everything that computes goes through the CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


class Context:
    """``require 'cunn'; cutorch.setDevice(n)`` (train.lua:168-169) -> dcgansr_ctx."""

    def __init__(self, device=0, precision="strict", world_size=1, rank=0, sync_bn=False, use_graph=False):
        self.lib = L.load()
        prec = {"strict": L.STRICT_FP32, "fp32": L.STRICT_FP32, "tf32": L.FAST_TF32, "fast": L.FAST_TF32}[precision]
        cfg = L.Cfg(device, prec, world_size, rank, int(sync_bn), int(use_graph))
        h = C.c_void_p()
        L.check(self.lib.dcgansr_ctx_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.world_size, self.rank = world_size, rank

    def close(self):
        if getattr(self, "h", None):
            self.lib.dcgansr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        L.check(self.lib.dcgansr_synchronize(self.h), self.h)

    def timer_begin(self):
        L.check(self.lib.dcgansr_timer_begin(self.h), self.h)

    def timer_end(self) -> float:
        ms = C.c_float()
        L.check(self.lib.dcgansr_timer_end(self.h, C.byref(ms)), self.h)
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(self.lib.dcgansr_launch_count(self.h, C.byref(n)), self.h)
        return n.value

    def profile_begin(self):
        L.check(self.lib.dcgansr_profile_begin(self.h), self.h)

    def profile_end(self):
        """-> list of dicts {name, work, kind, launches, ms} sorted by time (work = algorithmic per launch)."""
        import json
        buf = C.create_string_buffer(1 << 16)
        L.check(self.lib.dcgansr_profile_end(self.h, buf, len(buf)), self.h)
        return json.loads(buf.value.decode())

    def flush_l2(self):
        L.check(self.lib.dcgansr_flush_l2(self.h), self.h)

    # -- data-parallel communicator (NCCL; the id travels over torch.distributed / any host channel)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        L.check(self.lib.dcgansr_comm_get_unique_id(self.h, buf), self.h)
        return buf.raw

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        L.check(self.lib.dcgansr_comm_init(self.h, buf), self.h)

    def comm_peer_enabled(self) -> bool:
        """True when the sync_bn statistics go through the one-shot NVLink peer-memory all-reduce (kernels_peer.cu)."""
        return bool(self.lib.dcgansr_comm_peer_enabled(self.h))


# ---- module descriptors: constructor signatures of torch/nn ------------------------------------
class Module:
    spec: dict


class SpatialConvolution(Module):
    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=0):
        assert kW == kH and dW == dH and padW == padH, "square kernels only"
        self.spec = dict(kind="conv", cin=nIn, cout=nOut, k=kW, s=dW, p=padW)


class SpatialFullConvolution(Module):
    def __init__(self, nIn, nOut, kW, kH, dW=1, dH=1, padW=0, padH=0, adjW=0, adjH=0):
        assert kW == kH and dW == dH and padW == padH and adjW == adjH, "square kernels only"
        self.spec = dict(kind="fullconv", cin=nIn, cout=nOut, k=kW, s=dW, p=padW, adj=adjW)


class SpatialBatchNormalization(Module):
    def __init__(self, nFeature, eps=1e-5, momentum=0.1, affine=True):
        assert affine, "the reference only uses affine BN"
        self.spec = dict(kind="bn", c=nFeature, eps=eps, momentum=momentum)


class ReLU(Module):
    def __init__(self, inplace=False):
        self.spec = dict(kind="relu")


class LeakyReLU(Module):
    def __init__(self, negval=0.01, inplace=False):
        self.spec = dict(kind="lrelu", negval=negval)


class Tanh(Module):
    def __init__(self):
        self.spec = dict(kind="tanh")


class Sigmoid(Module):
    def __init__(self):
        self.spec = dict(kind="sigmoid")


class SpatialUpSamplingNearest(Module):
    def __init__(self, scale):
        self.spec = dict(kind="upnearest", scale=scale)


class View(Module):
    def __init__(self, *sizes):
        self.spec = dict(kind="view")

    def setNumInputDims(self, n):
        return self


_KIND = {"conv": L.CONV, "fullconv": L.FULLCONV, "bn": L.BN, "relu": L.RELU, "lrelu": L.LRELU, "tanh": L.TANH,
         "sigmoid": L.SIGMOID, "upnearest": L.UPNEAREST, "view": L.VIEW}


def specs_to_layers(specs):
    """dict specs -> ctypes array of dcgansr_layer (include/dcgansr.h)."""
    arr = (L.Layer * len(specs))()
    for i, s in enumerate(specs):
        l = arr[i]
        l.kind = _KIND[s["kind"]]
        if s["kind"] in ("conv", "fullconv"):
            l.cin, l.cout = s["cin"], s["cout"]
            l.kh = l.kw = s["k"]
            l.sh = l.sw = s.get("s", 1)
            l.ph = l.pw = s.get("p", 0)
            l.adjh = l.adjw = s.get("adj", 0)
        elif s["kind"] == "bn":
            l.cin = l.cout = s["c"]
            l.eps, l.momentum = s.get("eps", 1e-5), s.get("momentum", 0.1)
        elif s["kind"] == "lrelu":
            l.negval = s.get("negval", 0.2)
        elif s["kind"] == "upnearest":
            l.scale = s.get("scale", 2)
    return arr


class Sequential:
    """nn.Sequential (train.lua:97,119).  ``cuda()`` realises it as a dcgansr_net on the device."""

    def __init__(self):
        self.specs = []
        self.ctx = None
        self.h = None
        self.in_shape = None
        self.max_batch = 0
        self.output = None

    @classmethod
    def from_specs(cls, specs):
        s = cls()
        s.specs = [dict(x) for x in specs]
        return s

    def add(self, m: Module):
        assert self.h is None, "net already realised"
        self.specs.append(dict(m.spec))
        return self

    # net:cuda() (train.lua:180).  ctx=None gives a plan-only net (shapes / parameter counts).
    def cuda(self, ctx, in_shape, max_batch):
        lib = L.load()
        self.lib = lib
        layers = specs_to_layers(self.specs)
        h = C.c_void_p()
        c, hh, w = in_shape
        L.check(lib.dcgansr_net_create(ctx.h if ctx is not None else None, layers, len(self.specs), c, hh, w, max_batch,
                                       C.byref(h)), ctx.h if ctx is not None else None)
        self.ctx, self.h, self.in_shape, self.max_batch = ctx, h, tuple(in_shape), max_batch
        return self

    def close(self):
        if self.h:
            self.lib.dcgansr_net_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        L.check(rc, self.ctx.h if self.ctx is not None else None)

    def out_shape(self):
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.lib.dcgansr_net_out_shape(self.h, C.byref(c), C.byref(h), C.byref(w)))
        return c.value, h.value, w.value

    def num_params(self) -> int:
        n = C.c_int64()
        self._chk(self.lib.dcgansr_net_num_params(self.h, C.byref(n)))
        return n.value

    def num_bn_channels(self) -> int:
        n = C.c_int64()
        self._chk(self.lib.dcgansr_net_num_bn_channels(self.h, C.byref(n)))
        return n.value

    # Module:getParameters() (train.lua:202-203): flat vectors in module order
    def getParameters(self):
        return self.get_params(), self.get_grads()

    def get_params(self):
        out = np.empty(self.num_params(), np.float32)
        self._chk(self.lib.dcgansr_net_get_params(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def get_grads(self):
        out = np.empty(self.num_params(), np.float32)
        self._chk(self.lib.dcgansr_net_get_grads(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_params(self, flat):
        a, p = _f32(flat)
        assert a.size == self.num_params()
        self._chk(self.lib.dcgansr_net_set_params(self.h, p))

    def get_bn_running(self):
        n = self.num_bn_channels()
        m, v = np.empty(n, np.float32), np.empty(n, np.float32)
        self._chk(self.lib.dcgansr_net_get_bn_running(self.h, m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)))
        return m, v

    def set_bn_running(self, mean, var):
        a, pa = _f32(mean)
        b, pb = _f32(var)
        self._chk(self.lib.dcgansr_net_set_bn_running(self.h, pa, pb))

    def get_adam_state(self):
        n = self.num_params()
        m, v, t = np.empty(n, np.float32), np.empty(n, np.float32), C.c_int64()
        self._chk(self.lib.dcgansr_net_get_adam_state(self.h, m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                                      C.byref(t)))
        return m, v, t.value

    def set_adam_state(self, m, v, t):
        a, pa = _f32(m)
        b, pb = _f32(v)
        self._chk(self.lib.dcgansr_net_set_adam_state(self.h, pa, pb, int(t)))

    # net:forward(x) (train.lua:218)
    def forward(self, x):
        a, pa = _f32(x)
        B = a.shape[0]
        assert tuple(a.shape[1:]) == self.in_shape, (a.shape, self.in_shape)
        c, h, w = self.out_shape()
        y = np.empty((B, c, h, w), np.float32)
        self._chk(self.lib.dcgansr_net_forward(self.h, pa, B, y.ctypes.data_as(C.c_void_p)))
        self.output = y
        return y

    # net:backward(x, dy) (train.lua:222): returns gradInput
    def backward(self, x, dy, need_dx=True):
        return self._bwd(self.lib.dcgansr_net_backward, x, dy, need_dx)

    # net:updateGradInput(x, dy) (train.lua:268)
    def updateGradInput(self, x, dy):
        return self._bwd(self.lib.dcgansr_net_update_grad_input, x, dy, True)

    def _bwd(self, fn, x, dy, need_dx):
        a, pa = _f32(x)
        g, pg = _f32(dy)
        B = a.shape[0]
        dx = np.empty_like(a) if need_dx else None
        self._chk(fn(self.h, pa, pg, B, dx.ctypes.data_as(C.c_void_p) if need_dx else None))
        return dx

    def zeroGradParameters(self):
        self._chk(self.lib.dcgansr_net_zero_grads(self.h))


class optim:
    """optim.adam(feval, x, state) (train.lua:280): the update runs on the net's flat device vectors."""

    @staticmethod
    def adam(net: Sequential, config: dict):
        net._chk(net.lib.dcgansr_net_adam(net.h, config.get("learningRate", 1e-3), config.get("beta1", 0.9),
                                          config.get("beta2", 0.999), config.get("epsilon", 1e-8)))


def make_step_cfg(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0, pixel_label=False, pixel_div=1.0,
                  lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
    return L.StepCfg(L.LOSS_BCE if family == "bce" else L.LOSS_MSE, real_label, fake_label, gen_label, int(pixel_label),
                     pixel_div, lr, beta1, beta2, eps)


def train_step(ctx: Context, netG: Sequential, netD: Sequential, cfg: L.StepCfg, real, want_losses=True):
    """One iteration of the reference loop (fDx -> adam(D) -> fGx -> adam(G), train.lua:208-283) on this
    rank's shard ``real`` (host NCHW fp32).  Returns (errD_real, errD_fake, errG) or None."""
    a, pa = _f32(real)
    losses = (C.c_float * 3)()
    L.check(ctx.lib.dcgansr_train_step(ctx.h, netG.h, netD.h, C.byref(cfg), pa, a.shape[0],
                                       losses if want_losses else None), ctx.h)
    return tuple(losses) if want_losses else None


def train_step_ptr(ctx: Context, netG, netD, cfg, host_ptr: int, batch: int, losses_out=None):
    """Same, from a raw (pinned) host pointer -- the e2e leg of bench.py."""
    L.check(ctx.lib.dcgansr_train_step(ctx.h, netG.h, netD.h, C.byref(cfg), C.c_void_p(host_ptr), batch, losses_out), ctx.h)


def stage_batch(ctx: Context, netD: Sequential, real, slot: int):
    a, pa = _f32(real)
    L.check(ctx.lib.dcgansr_stage_batch(ctx.h, netD.h, pa, a.shape[0], slot), ctx.h)


def train_step_staged(ctx: Context, netG, netD, cfg, slot: int, batch: int, want_losses=False):
    losses = (C.c_float * 3)()
    L.check(ctx.lib.dcgansr_train_step_staged(ctx.h, netG.h, netD.h, C.byref(cfg), slot, batch,
                                              losses if want_losses else None), ctx.h)
    return tuple(losses) if want_losses else None


# ---- patch extraction / re-assembly (train-gray-patch*.lua) ----------------------------------------
def _patch_geom(fine, patch, overlap=0):
    """(line, nper, stride) of the reference scripts: non-overlapping -> the reference's (patchSize, (fine/patch)^2, patchSize)
    (train-gray-patch.lua:267-273), overlapping -> (overlapPatchLine, overlapPatchLine^2, overlap) (…-overlap.lua:387-399)."""
    if overlap:
        line = (fine - overlap) // (patch - overlap)
        return line, line * line, overlap
    return patch, (fine // patch) ** 2, patch


def extract_patches(ctx: Context, images, patch: int, line: int, nper: int, stride: int):
    a, pa = _f32(images)
    k, h, w = a.shape
    out = np.empty((k * nper, patch, patch), np.float32)
    L.check(ctx.lib.dcgansr_extract_patches(ctx.h, pa, out.ctypes.data_as(C.c_void_p), k, h, w, patch, line, nper, stride), ctx.h)
    return out


def assemble_patches(ctx: Context, patches, images, patch: int, line: int, nper: int, stride: int):
    p_, pp = _f32(patches)
    out = np.array(images, dtype=np.float32, order="C", copy=True)
    k, h, w = out.shape
    L.check(ctx.lib.dcgansr_assemble_patches(ctx.h, pp, out.ctypes.data_as(C.c_void_p), k, h, w, patch, line, nper, stride), ctx.h)
    return out


def stitch_overlap(ctx: Context, patches, fine: int, patch: int, overlap: int, fix_top_cost: bool = False):
    """Minimum-error boundary cut stitching (train-gray-patch-batch-overlap.lua:457-694).  patches: [k*L*L][patch][patch] with
    L = (fine - overlap) / (patch - overlap); returns [k][fine][fine] (zeros where no patch reaches, like the reference)."""
    p_, pp = _f32(patches)
    line = (fine - overlap) // (patch - overlap) if patch > overlap else 0
    if line < 1 or p_.ndim != 3 or p_.shape[0] % (line * line) != 0:
        raise ValueError("patches must be [k*L*L][patch][patch]")
    k = p_.shape[0] // (line * line)
    out = np.zeros((k, fine, fine), np.float32)
    L.check(ctx.lib.dcgansr_stitch_overlap(ctx.h, pp, out.ctypes.data_as(C.c_void_p), k, fine, fine, patch, overlap,
                                           1 if fix_top_cost else 0), ctx.h)
    return out


def stage_patches(ctx: Context, netD: Sequential, images, patch: int, line: int, nper: int, stride: int, slot: int):
    a, pa = _f32(images)
    k, h, w = a.shape
    L.check(ctx.lib.dcgansr_stage_patches(ctx.h, netD.h, pa, k, h, w, patch, line, nper, stride, slot), ctx.h)
    return k * nper


# ---- evaluation metrics (train-gray-3.lua:143-221) ---------------------------------------------------
def _metric(ctx: Context, fn, a, b):
    a_, pa = _f32(a)
    b_, pb = _f32(b)
    n, h, w = a_.shape
    out = np.empty(n, np.float32)
    L.check(fn(ctx.h, pa, pb, out.ctypes.data_as(C.c_void_p), n, h, w), ctx.h)
    return out


def psnr(ctx: Context, a, b):
    """calPSNR per image pair; a, b: [n][h][w]."""
    return _metric(ctx, ctx.lib.dcgansr_psnr, a, b)


def ssim(ctx: Context, a, b):
    """calSSIM per image pair; a, b: [n][h][w] in [-1, 1]."""
    return _metric(ctx, ctx.lib.dcgansr_ssim, a, b)


def scale_bilinear(ctx: Context, images, dst_h: int, dst_w: int):
    """image.scale(img, dst_w, dst_h) (bilinear, enlarging) per image; images: [n][h][w]  (train-gray-3.lua:399)."""
    a_, pa = _f32(images)
    n, h, w = a_.shape
    out = np.empty((n, dst_h, dst_w), np.float32)
    L.check(ctx.lib.dcgansr_scale_bilinear(ctx.h, pa, out.ctypes.data_as(C.c_void_p), n, h, w, dst_h, dst_w), ctx.h)
    return out
