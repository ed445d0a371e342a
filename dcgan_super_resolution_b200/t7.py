"""Torch7 binary serialisation (`.t7`) reader / writer for the nets of the path.

The reference would checkpoint with `torch.save(path, netG:clearState())` (train.lua:297-299, commented out there) and caches
its file index as `.t7` too (data/donkey_folder.lua:74-92).  This module reads and writes that on-disk format so weights can
move between a Torch7 installation and libdcgansr: `load_net` turns a serialised `nn.Sequential` of the modules the scripts
use (train.lua:97-136 and the gray / patch variants, plain `nn.*` or `cudnn.*` after `cudnn.convert`, Float or Cuda tensors)
into this package's layer specs + the flat parameter vector in `getParameters()` order + BN running statistics; `save_net`
writes the same structure back.  Pure host-side byte handling: nothing here touches the device or the oracle.

Format (torch7 `File.lua` writeObject / readObject, binary mode, little endian, 8-byte longs -- un-vendored upstream, restated
from its published layout; no `.t7` file exists in /root/reference, so this is PARITY UNPINNED like the rest):
  object      := int32 type, payload
  type 0 nil | 1 number: float64 | 2 string: int32 n, n bytes | 5 boolean: int32
  type 3 table: int32 index; if the index was seen before nothing follows (back-reference), else int32 npairs, npairs x (object key,
         object value)
  type 4 torch object: int32 index (back-reference rule as above), string version ("V 1"; legacy files put the class name
         here), string class name, then
           *Tensor : int32 ndim, ndim x int64 sizes, ndim x int64 strides, int64 storage offset (1-based), object storage
           *Storage: int64 n, n raw elements
           any other class (nn modules): one object = the table of its fields
"""
from __future__ import annotations

import io
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5

_ELEM = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Short": np.int16, "Char": np.int8,
         "Byte": np.uint8, "Cuda": np.float32, "CudaDouble": np.float64, "CudaLong": np.int64, "CudaInt": np.int32,
         "CudaByte": np.uint8}
_NAME_OF = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long", np.dtype(np.int32): "Int",
            np.dtype(np.int16): "Short", np.dtype(np.int8): "Char", np.dtype(np.uint8): "Byte"}


class T7Error(ValueError):
    pass


class TorchObject:
    """A serialised torch class instance that is not a tensor / storage (e.g. nn.SpatialConvolution): class name + fields."""

    def __init__(self, classname, fields=None):
        self.classname = classname
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return f"TorchObject({self.classname}, {sorted(map(str, self.fields))})"


class Storage:
    """torch.*Storage: kept distinct from tensors so LongStorage fields (nn.View.size) round-trip."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data).reshape(-1)


def _split_class(name):
    """'torch.FloatTensor' -> ('Float', 'Tensor'); None if not a tensor / storage class."""
    if not name.startswith("torch."):
        return None
    body = name[6:]
    for kind in ("Tensor", "Storage"):
        if body.endswith(kind) and body[:-len(kind)] in _ELEM:
            return body[:-len(kind)], kind
    return None


# ---------------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, f):
        self.f = f
        self.seen = {}

    def _rd(self, fmt):
        n = struct.calcsize(fmt)
        b = self.f.read(n)
        if len(b) != n:
            raise T7Error("truncated .t7 stream")
        return struct.unpack(fmt, b)[0]

    def _str(self):
        n = self._rd("<i")
        if n < 0:
            raise T7Error("negative string length")
        b = self.f.read(n)
        if len(b) != n:
            raise T7Error("truncated .t7 stream")
        return b.decode("latin-1")

    def obj(self):
        t = self._rd("<i")
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            return self._rd("<d")
        if t == TYPE_STRING:
            return self._str()
        if t == TYPE_BOOLEAN:
            return self._rd("<i") == 1
        if t == TYPE_TABLE:
            idx = self._rd("<i")
            if idx in self.seen:
                return self.seen[idx]
            out = {}
            self.seen[idx] = out
            for _ in range(self._rd("<i")):
                k = self.obj()
                v = self.obj()
                if isinstance(k, float) and k.is_integer():
                    k = int(k)
                out[k] = v
            return out
        if t == TYPE_TORCH:
            idx = self._rd("<i")
            if idx in self.seen:
                return self.seen[idx]
            ver = self._str()
            cls = self._str() if ver.startswith("V ") else ver          # legacy files: no version record
            sc = _split_class(cls)
            if sc is None:
                o = TorchObject(cls)
                self.seen[idx] = o
                fields = self.obj()
                if not isinstance(fields, dict):
                    raise T7Error(f"{cls}: field table expected")
                o.fields = fields
                return o
            elem, kind = sc
            if kind == "Storage":
                n = self._rd("<q")
                dt = np.dtype(_ELEM[elem]).newbyteorder("<")
                raw = self.f.read(n * dt.itemsize)
                if n < 0 or len(raw) != n * dt.itemsize:
                    raise T7Error("truncated storage")
                o = Storage(np.frombuffer(raw, dt).astype(_ELEM[elem]))
                self.seen[idx] = o
                return o
            nd = self._rd("<i")
            sizes = [self._rd("<q") for _ in range(nd)]
            strides = [self._rd("<q") for _ in range(nd)]
            off = self._rd("<q") - 1
            st = self.obj()
            if st is None or nd == 0:
                arr = np.zeros((0,), _ELEM[elem])
            else:
                if not isinstance(st, Storage):
                    raise T7Error("tensor without a storage object")
                need = off + sum((s - 1) * k for s, k in zip(sizes, strides)) + 1 if all(s > 0 for s in sizes) else 0
                if off < 0 or need > st.data.size or any(k < 0 for k in strides):
                    raise T7Error("tensor view reaches outside its storage")
                item = st.data.dtype.itemsize
                arr = np.lib.stride_tricks.as_strided(st.data[off:], shape=sizes, strides=[k * item for k in strides]).copy()
            self.seen[idx] = arr
            return arr
        raise T7Error(f"unsupported .t7 object type {t} (functions are not data)")


def load(path_or_bytes):
    """Read one object from a `.t7` file (binary mode).  Tables -> dict, tensors -> numpy arrays (copies), storages -> Storage,
    other torch classes -> TorchObject."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _Reader(io.BytesIO(bytes(path_or_bytes))).obj()
    with open(path_or_bytes, "rb") as f:
        return _Reader(f).obj()


# ---------------------------------------------------------------------------------------------------------------- writer
class _Writer:
    def __init__(self, f):
        self.f = f
        self.n = 0

    def _w(self, fmt, v):
        self.f.write(struct.pack(fmt, v))

    def _str(self, s):
        b = s.encode("latin-1")
        self._w("<i", len(b))
        self.f.write(b)

    def _head(self, cls):
        self._w("<i", TYPE_TORCH)
        self.n += 1
        self._w("<i", self.n)
        self._str("V 1")
        self._str(cls)

    def _storage(self, data):
        self._head("torch.%sStorage" % _NAME_OF[data.dtype])
        self._w("<q", data.size)
        self.f.write(data.astype(data.dtype.newbyteorder("<")).tobytes())

    def obj(self, o):
        if o is None:
            self._w("<i", TYPE_NIL)
        elif isinstance(o, (bool, np.bool_)):
            self._w("<i", TYPE_BOOLEAN)
            self._w("<i", 1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self._w("<i", TYPE_NUMBER)
            self._w("<d", float(o))
        elif isinstance(o, str):
            self._w("<i", TYPE_STRING)
            self._str(o)
        elif isinstance(o, (dict, list, tuple)):
            items = list(o.items()) if isinstance(o, dict) else [(i + 1, v) for i, v in enumerate(o)]
            self._w("<i", TYPE_TABLE)
            self.n += 1
            self._w("<i", self.n)
            self._w("<i", len(items))
            for k, v in items:
                self.obj(k)
                self.obj(v)
        elif isinstance(o, Storage):
            self._storage(o.data)
        elif isinstance(o, np.ndarray):
            if o.dtype not in _NAME_OF:
                raise T7Error(f"no torch tensor type for dtype {o.dtype}")
            a = np.ascontiguousarray(o)
            self._head("torch.%sTensor" % _NAME_OF[a.dtype])
            nd = a.ndim if a.size else 0
            self._w("<i", nd)
            for s in a.shape[:nd]:
                self._w("<q", s)
            for k in (np.array(a.strides[:nd]) // a.itemsize):
                self._w("<q", int(k))
            self._w("<q", 1)
            if nd:
                self._storage(a.reshape(-1))
            else:
                self._w("<i", TYPE_NIL)
        elif isinstance(o, TorchObject):
            self._head(o.classname)
            self.obj(o.fields)
        else:
            raise T7Error(f"cannot serialise {type(o).__name__}")


def save(path, o):
    """Write one object as a `.t7` file (binary mode): the inverse of `load`.  path=None returns the bytes."""
    if path is None:
        b = io.BytesIO()
        _Writer(b).obj(o)
        return b.getvalue()
    with open(path, "wb") as f:
        _Writer(f).obj(o)


# ------------------------------------------------------------------------------------------- nn.Sequential <-> layer specs
def _i(v):
    return int(round(float(v)))


def _module_to_spec(m):
    """One serialised module -> (spec, [param arrays in getParameters order], (running_mean, running_var) or None)."""
    if not isinstance(m, TorchObject):
        raise T7Error("nn module expected")
    name = m.classname.split(".")[-1]
    g = m.get
    if name in ("SpatialConvolution", "SpatialConvolutionMM", "SpatialFullConvolution"):
        full = name == "SpatialFullConvolution"
        if _i(g("kW")) != _i(g("kH")) or _i(g("dW", 1)) != _i(g("dH", 1)) or _i(g("padW", 0)) != _i(g("padH", 0)):
            raise T7Error(f"{m.classname}: only square kernels / strides / paddings are on the path")
        if g("bias") is not None and np.asarray(g("bias")).size:
            raise T7Error(f"{m.classname}: the path's convolutions are bias-free (weights_init calls noBias, train.lua:46)")
        cin, cout, k = _i(g("nInputPlane")), _i(g("nOutputPlane")), _i(g("kW"))
        spec = dict(kind="fullconv" if full else "conv", cin=cin, cout=cout, k=k, s=_i(g("dW", 1)), p=_i(g("padW", 0)))
        if full:
            spec["adj"] = _i(g("adjW", 0))
        w = np.asarray(g("weight"), np.float32)
        want = (cin, cout, k, k) if full else (cout, cin, k, k)       # fullconv is IOHW, conv OIHW (or O x I*k*k for the MM class)
        if w.size != int(np.prod(want)):
            raise T7Error(f"{m.classname}: weight has {w.size} elements, expected {want}")
        return spec, [w.reshape(want)], None
    if name in ("SpatialBatchNormalization", "BatchNormalization"):
        if g("affine") is False or g("weight") is None:
            raise T7Error("only affine batch normalisation is on the path")
        w, b = np.asarray(g("weight"), np.float32), np.asarray(g("bias"), np.float32)
        rm = np.asarray(g("running_mean"), np.float32)
        if g("running_var") is not None:
            rv = np.asarray(g("running_var"), np.float32)
        elif g("running_std") is not None:                               # pre-2016 nn kept 1/sqrt(var + eps)
            rv = (1.0 / np.square(np.asarray(g("running_std"), np.float64)) - float(g("eps", 1e-5))).astype(np.float32)
        else:
            raise T7Error("batch normalisation without running statistics")
        return dict(kind="bn", c=w.size, eps=float(g("eps", 1e-5)), momentum=float(g("momentum", 0.1))), [w, b], (rm, rv)
    if name == "ReLU" or (name == "Threshold" and float(g("threshold", 0)) == 0 and float(g("val", 0)) == 0):
        return dict(kind="relu"), [], None
    if name == "LeakyReLU":
        return dict(kind="lrelu", negval=float(g("negval", 0.01))), [], None
    if name == "Tanh":
        return dict(kind="tanh"), [], None
    if name == "Sigmoid":
        return dict(kind="sigmoid"), [], None
    if name == "SpatialUpSamplingNearest":
        return dict(kind="upnearest", scale=_i(g("scale_factor", 2))), [], None
    if name == "View":
        return dict(kind="view"), [], None
    raise T7Error(f"module {m.classname} is not on the path")


def net_from_object(o):
    """Serialised nn.Sequential -> (specs, flat params, (running_mean, running_var)), parameters in getParameters() order."""
    if not isinstance(o, TorchObject) or o.classname.split(".")[-1] != "Sequential":
        raise T7Error("nn.Sequential expected")
    mods = o.get("modules") or {}
    specs, params, rms, rvs = [], [], [], []
    for i in range(1, len(mods) + 1):
        if i not in mods:
            raise T7Error("modules table is not a sequence")
        s, ps, run = _module_to_spec(mods[i])
        specs.append(s)
        params += [p.reshape(-1) for p in ps]
        if run is not None:
            rms.append(run[0])
            rvs.append(run[1])
    cat = lambda xs: np.concatenate(xs).astype(np.float32) if xs else np.zeros(0, np.float32)
    return specs, cat(params), (cat(rms), cat(rvs))


def load_net(path_or_bytes):
    return net_from_object(load(path_or_bytes))


def net_to_object(specs, flat_params, running=None):
    """Layer specs + flat parameter vector (+ BN running statistics) -> the TorchObject tree of `netX:clearState()`."""
    flat = np.asarray(flat_params, np.float32).reshape(-1)
    rm, rv = (np.asarray(running[0], np.float32), np.asarray(running[1], np.float32)) if running is not None else (None, None)
    pos = rpos = 0
    empty = lambda: np.zeros((0,), np.float32)

    def take(n):
        nonlocal pos
        if pos + n > flat.size:
            raise T7Error("parameter vector too short for the specs")
        out = flat[pos:pos + n].copy()
        pos += n
        return out

    mods = []
    for s in specs:
        base = {"_type": "torch.FloatTensor", "output": empty(), "gradInput": empty(), "train": True}
        kind = s["kind"]
        if kind in ("conv", "fullconv"):
            cin, cout, k = s["cin"], s["cout"], s["k"]
            shape = (cin, cout, k, k) if kind == "fullconv" else (cout, cin, k, k)
            w = take(int(np.prod(shape))).reshape(shape)
            f = dict(base, nInputPlane=cin, nOutputPlane=cout, kW=k, kH=k, dW=s.get("s", 1), dH=s.get("s", 1), padW=s.get("p", 0),
                     padH=s.get("p", 0), weight=w, gradWeight=np.zeros_like(w))
            if kind == "fullconv":
                f.update(adjW=s.get("adj", 0), adjH=s.get("adj", 0))
            mods.append(TorchObject("nn.SpatialFullConvolution" if kind == "fullconv" else "nn.SpatialConvolution", f))
        elif kind == "bn":
            c = s["c"]
            w, b = take(c), take(c)
            mean = rm[rpos:rpos + c].copy() if rm is not None else np.zeros(c, np.float32)
            var = rv[rpos:rpos + c].copy() if rv is not None else np.ones(c, np.float32)
            rpos += c
            mods.append(TorchObject("nn.SpatialBatchNormalization", dict(
                base, nDim=4, eps=s.get("eps", 1e-5), momentum=s.get("momentum", 0.1), affine=True, weight=w, bias=b,
                gradWeight=np.zeros_like(w), gradBias=np.zeros_like(b), running_mean=mean, running_var=var)))
        elif kind == "relu":
            mods.append(TorchObject("nn.ReLU", dict(base, threshold=0, val=0, inplace=True)))
        elif kind == "lrelu":
            mods.append(TorchObject("nn.LeakyReLU", dict(base, negval=s.get("negval", 0.2), inplace=True)))
        elif kind == "tanh":
            mods.append(TorchObject("nn.Tanh", dict(base)))
        elif kind == "sigmoid":
            mods.append(TorchObject("nn.Sigmoid", dict(base)))
        elif kind == "upnearest":
            mods.append(TorchObject("nn.SpatialUpSamplingNearest", dict(base, scale_factor=s.get("scale", 2))))
        elif kind == "view":
            mods.append(TorchObject("nn.View", dict(base, size=Storage(np.array([1], np.int64)), numElements=1, numInputDims=3)))
        else:
            raise T7Error(f"unknown layer kind {kind}")
    if pos != flat.size:
        raise T7Error("parameter vector longer than the specs need")
    return TorchObject("nn.Sequential", {"_type": "torch.FloatTensor", "output": empty(), "gradInput": empty(), "train": True,
                                         "modules": mods})


def save_net(path, specs, flat_params, running=None):
    return save(path, net_to_object(specs, flat_params, running))
