"""Torch7 `.t7` serialisation: byte-level known answers built by hand from the format, shared-storage views and
back-references as `getParameters()` produces them, round trips of every net family, error paths."""
import struct

import numpy as np
import pytest

from dcgan_super_resolution_b200 import init, models, t7
from util import rng


def _s(x):
    b = x.encode()
    return struct.pack("<i", len(b)) + b


def _torch_head(index, cls, version="V 1"):
    return struct.pack("<ii", 4, index) + (_s(version) if version else b"") + _s(cls)


def _float_tensor_bytes(index, sizes, strides, offset, storage_bytes):
    nd = len(sizes)
    return (_torch_head(index, "torch.FloatTensor") + struct.pack("<i", nd) + struct.pack("<%dq" % nd, *sizes)
            + struct.pack("<%dq" % nd, *strides) + struct.pack("<q", offset) + storage_bytes)


def _float_storage_bytes(index, values):
    return _torch_head(index, "torch.FloatStorage") + struct.pack("<q", len(values)) + np.asarray(values, "<f4").tobytes()


def test_known_bytes_scalar_types_and_tensor():
    assert t7.save(None, None) == struct.pack("<i", 0)
    assert t7.save(None, 2.5) == struct.pack("<id", 1, 2.5)
    assert t7.save(None, "abc") == struct.pack("<i", 2) + _s("abc")
    assert t7.save(None, True) == struct.pack("<ii", 5, 1)
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    want = _float_tensor_bytes(1, (2, 3), (3, 1), 1, _float_storage_bytes(2, a.reshape(-1)))
    assert t7.save(None, a) == want
    assert np.array_equal(t7.load(want), a)
    # table {1: 7, "k": false}: type 3, index 1, 2 pairs
    tb = struct.pack("<iii", 3, 1, 2) + struct.pack("<id", 1, 1.0) + struct.pack("<id", 1, 7.0) + struct.pack("<i", 2) + _s("k") + struct.pack("<ii", 5, 0)
    assert t7.load(tb) == {1: 7.0, "k": False}
    assert t7.save(None, {1: 7, "k": False}) == tb
    # empty tensor (what clearState leaves in .output): ndim 0, offset 1, nil storage
    assert t7.save(None, np.zeros((0,), np.float32)) == _torch_head(1, "torch.FloatTensor") + struct.pack("<iq", 0, 1) + struct.pack("<i", 0)


def test_shared_storage_views_and_back_references():
    """After getParameters() (train.lua:202-203) every weight is a view into ONE storage: the first tensor carries the storage,
    later ones refer to it by index; offsets are 1-based; strides need not be contiguous."""
    flat = np.arange(20, dtype=np.float32)
    st = _float_storage_bytes(3, flat)
    t_a = _float_tensor_bytes(2, (2, 3), (3, 1), 1, st)                                   # elements 0..5
    t_b = _float_tensor_bytes(4, (3, 2), (1, 3), 7, struct.pack("<ii", 4, 3))              # transposed view of 6..11, storage by reference
    t_b_again = struct.pack("<ii", 4, 4)                                                   # the tensor itself by reference
    table = (struct.pack("<iii", 3, 1, 3) + struct.pack("<i", 2) + _s("a") + t_a + struct.pack("<i", 2) + _s("b") + t_b
             + struct.pack("<i", 2) + _s("c") + t_b_again)
    o = t7.load(table)
    assert np.array_equal(o["a"], flat[:6].reshape(2, 3))
    assert np.array_equal(o["b"], flat[6:12].reshape(2, 3).T)
    assert o["c"] is o["b"]
    # a view reaching past its storage is an error, not a read out of bounds
    bad = _float_tensor_bytes(1, (4, 4), (4, 1), 10, _float_storage_bytes(2, flat))
    with pytest.raises(t7.T7Error):
        t7.load(bad)


def _norm(specs):
    out = []
    for s in specs:
        d = {k: v for k, v in s.items()}
        if d["kind"] in ("conv", "fullconv"):
            d.setdefault("s", 1), d.setdefault("p", 0)
            if d["kind"] == "fullconv":
                d.setdefault("adj", 0)
        if d["kind"] == "bn":
            d.setdefault("eps", 1e-5), d.setdefault("momentum", 0.1)
        if d["kind"] == "upnearest":
            d.setdefault("scale", 2)
        out.append(d)
    return out


FAMILIES = [models.train_lua_G(3, 12), models.dcgan64_D(3, 16), models.train_gray_G(8), models.train_gray_2_G(8),
            models.train_gray_3_G(4), models.patch_batch_G(4), models.patch_D(8)]


@pytest.mark.parametrize("specs", FAMILIES, ids=["train_G", "dcgan64_D", "gray_G", "gray2_G", "gray3_G", "patch_batch_G", "patch_D"])
def test_net_round_trip(specs, tmp_path):
    params = init.weights_init(specs, 99)
    nbn = sum(s["c"] for s in specs if s["kind"] == "bn")
    r = rng(5)
    running = (r.normal(0, 1, nbn).astype(np.float32), r.uniform(0.5, 2, nbn).astype(np.float32))
    path = tmp_path / "net.t7"
    t7.save_net(path, specs, params, running)
    specs2, params2, (rm, rv) = t7.load_net(path)
    assert _norm(specs2) == _norm(specs)
    assert np.array_equal(params2, params) and np.array_equal(rm, running[0]) and np.array_equal(rv, running[1])
    # the module table carries what Torch7's constructors set (train.lua:99-136)
    o = t7.load(path)
    assert o.classname == "nn.Sequential" and len(o["modules"]) == len(specs)
    first_conv = next(m for m in o["modules"].values() if "Convolution" in m.classname)
    assert first_conv["kW"] == first_conv["kH"] and first_conv.get("bias") is None and first_conv["weight"].ndim == 4


def test_dialects_cudnn_cuda_legacy_running_std():
    w = rng(1).normal(0, 0.02, (4, 2, 3, 3)).astype(np.float32)
    g, b = np.ones(4, np.float32), np.zeros(4, np.float32)
    var = np.array([0.5, 1.0, 2.0, 4.0], np.float32)
    conv = t7.TorchObject("cudnn.SpatialConvolution", dict(nInputPlane=2, nOutputPlane=4, kW=3, kH=3, dW=1, dH=1, padW=0, padH=0, weight=w))
    bn = t7.TorchObject("cudnn.SpatialBatchNormalization", dict(eps=1e-5, momentum=0.1, affine=True, weight=g, bias=b,
                                                                running_mean=b, running_std=(1.0 / np.sqrt(var.astype(np.float64) + 1e-5))))
    seq = t7.TorchObject("nn.Sequential", dict(modules=[conv, bn, t7.TorchObject("cudnn.ReLU", dict(inplace=True))]))
    raw = t7.save(None, seq)
    specs, params, (rm, rv) = t7.load_net(raw)
    assert [s["kind"] for s in specs] == ["conv", "bn", "relu"] and params.size == w.size + 8
    assert np.allclose(rv, var, rtol=1e-5)
    # torch.CudaTensor payloads are float32 like FloatTensor
    cuda_t = (_torch_head(1, "torch.CudaTensor") + struct.pack("<i", 1) + struct.pack("<q", 3) + struct.pack("<q", 1) + struct.pack("<q", 1)
              + _torch_head(2, "torch.CudaStorage") + struct.pack("<q", 3) + np.array([1, 2, 3], "<f4").tobytes())
    assert t7.load(cuda_t).tolist() == [1.0, 2.0, 3.0]
    # legacy stream: no "V 1" record, the class name comes first
    legacy = struct.pack("<ii", 4, 1) + _s("torch.DoubleStorage") + struct.pack("<q", 2) + np.array([1.5, 2.5], "<f8").tobytes()
    assert t7.load(legacy).data.tolist() == [1.5, 2.5]
    # nn.View's LongStorage survives
    v = t7.load(t7.save(None, t7.net_to_object([dict(kind="view")], np.zeros(0, np.float32))))
    assert v["modules"][1]["size"].data.tolist() == [1]


def test_errors():
    with pytest.raises(t7.T7Error):
        t7.load(struct.pack("<i", 6))                                       # serialised Lua function
    with pytest.raises(t7.T7Error):
        t7.load(t7.save(None, np.arange(4, dtype=np.float32))[:-3])         # truncated
    conv = t7.TorchObject("nn.SpatialConvolution", dict(nInputPlane=1, nOutputPlane=1, kW=2, kH=2, weight=np.zeros((1, 1, 2, 2), np.float32),
                                                        bias=np.zeros(1, np.float32)))
    with pytest.raises(t7.T7Error):                                         # the path is bias-free (train.lua:46)
        t7.net_from_object(t7.TorchObject("nn.Sequential", dict(modules=[conv])))
    with pytest.raises(t7.T7Error):
        t7.net_from_object(t7.TorchObject("nn.Sequential", dict(modules=[t7.TorchObject("nn.Dropout", {})])))
    with pytest.raises(t7.T7Error):
        t7.net_to_object(models.patch_D(8), np.zeros(3, np.float32))


@pytest.mark.gpu
def test_gpu_checkpoint_resume_is_bit_exact(ctx, tmp_path):
    """Train a few steps, save G and D as .t7, rebuild from the files: forward outputs are identical."""
    import dcgan_super_resolution_b200 as dsr
    specsG, specsD = models.train_gray_3_G(4), models.patch_D(8)
    B = 16
    G = dsr.Sequential.from_specs(specsG).cuda(ctx, (1, 4, 4), B)
    D = dsr.Sequential.from_specs(specsD).cuda(ctx, (1, 8, 8), 2 * B)
    G.set_params(init.weights_init(specsG, 1))
    D.set_params(init.weights_init(specsD, 2))
    step = dsr.make_step_cfg(family="bce")
    r = rng(3)
    for _ in range(3):
        dsr.train_step(ctx, G, D, step, r.uniform(0, 1, (B, 1, 8, 8)).astype(np.float32))
    x = r.uniform(0, 1, (B, 1, 4, 4)).astype(np.float32)
    for net, specs, shape, name in ((G, specsG, (1, 4, 4), "G"), (D, specsD, (1, 8, 8), "D")):
        path = tmp_path / f"net_{name}.t7"
        t7.save_net(path, specs, net.get_params(), net.get_bn_running())
        specs2, params2, running2 = t7.load_net(path)
        net2 = dsr.Sequential.from_specs(specs2).cuda(ctx, shape, 2 * B)
        net2.set_params(params2)
        net2.set_bn_running(*running2)
        inp = x if name == "G" else r.uniform(0, 1, (B, 1, 8, 8)).astype(np.float32)
        assert np.array_equal(net.forward(inp), net2.forward(inp))
        assert all(np.array_equal(a, b) for a, b in zip(net.get_bn_running(), net2.get_bn_running()))
