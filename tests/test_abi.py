"""CPU tests of the drop-in boundary: libdcgansr.so loads, exports every symbol include/dcgansr.h declares,
plan-only nets (no GPU) reproduce the reference's shapes and parameter counts, and compute entry points fail
loudly without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import _lib as L
from dcgan_super_resolution_b200 import init, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dcgansr.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcgansr_[a-z0-9_]+)\s*\(", src)))


def test_header_parses_as_plain_c():
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ffi_cdef_block_parses_and_binds():
    """The block lua/dcgansr.lua hands to LuaJIT ffi.cdef, parsed by cffi (same C-declaration grammar) and bound
    to the built library in ABI mode -- the Lua shim itself cannot run here (no LuaJIT in the image)."""
    import cffi
    src = open(HEADER).read()
    b = src.index("\n", src.index("FFI-CDEF-BEGIN")) + 1
    e = src.index("/* FFI-CDEF-END")
    ffi = cffi.FFI()
    ffi.cdef(src[b:e])
    lib = ffi.dlopen(L.LIB_PATH)
    assert lib.dcgansr_version() >= 100
    cfg = ffi.new("dcgansr_step_cfg*", dict(loss=1, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8))
    assert cfg.beta1 == 0.5 and ffi.sizeof("dcgansr_layer") == C.sizeof(L.Layer)
    # plan-only net through the cffi binding: the call sequence of dcgansr.lua's Sequential:cuda()
    layers = ffi.new("dcgansr_layer[2]")
    layers[0].kind, layers[0].cin, layers[0].cout = 1, 3, 8
    layers[0].kh = layers[0].kw = 4
    layers[0].sh = layers[0].sw = 2
    layers[0].ph = layers[0].pw = 1
    layers[1].kind = 5
    out = ffi.new("dcgansr_net*[1]")
    assert lib.dcgansr_net_create(ffi.NULL, layers, 2, 3, 16, 16, 4, out) == 0
    n = ffi.new("int64_t[1]")
    assert lib.dcgansr_net_num_params(out[0], n) == 0 and n[0] == 8 * 3 * 16
    lib.dcgansr_net_destroy(out[0])


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_functions()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dcgansr.h but not exported"
    assert sorted(L.symbols()) == names, "ctypes binding and header disagree"
    assert lib.dcgansr_version() >= 100


def test_struct_layouts_match_header():
    # sizes the C compiler gives the three boundary structs
    code = '#include <stdio.h>\n#include "dcgansr.h"\nint main(){printf("%zu %zu %zu", sizeof(dcgansr_cfg), sizeof(dcgansr_layer), sizeof(dcgansr_step_cfg));return 0;}'
    exe = "/tmp/_dsr_sizes"
    r = subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=code, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split()))
    assert sizes == [C.sizeof(L.Cfg), C.sizeof(L.Layer), C.sizeof(L.StepCfg)]


@pytest.mark.parametrize("name,gparams,dparams", [("C3a", 102312, 2765568), ("C1a", 84384, 371008), ("C2", 42240, 2763520),
                                                  ("C1b", 1320576, 371008), ("C5", 11069184, 11036160)])
def test_plan_only_nets_match_survey_counts(name, gparams, dparams):
    cfg = models.config(name)
    hr, nc = cfg["hr"], cfg["nc"]
    G = dsr.Sequential.from_specs(cfg["G"]).cuda(None, (nc, hr // 2, hr // 2), 4)
    D = dsr.Sequential.from_specs(cfg["D"]).cuda(None, (nc, hr, hr), 4)
    assert G.num_params() == gparams and D.num_params() == dparams
    assert G.out_shape() == (nc, hr, hr)
    assert init.weights_init(cfg["G"], 1).size == gparams
    G.close(); D.close()


def test_d_output_shapes_at_baseline_sizes():
    # at BASELINE sizes D no longer ends at 1x1
    for name, want in (("C1a", (1, 1, 1)), ("C1b", (1, 25, 25)), ("C3a", (1, 1, 1)), ("C3b", (1, 5, 5)), ("C5", (1, 13, 13))):
        cfg = models.config(name)
        D = dsr.Sequential.from_specs(cfg["D"]).cuda(None, (cfg["nc"], cfg["hr"], cfg["hr"]), 2)
        assert D.out_shape() == want, name
        D.close()


def test_bad_descriptions_are_rejected():
    bad = [dict(kind="conv", cin=3, cout=8, k=4, s=2, p=1), dict(kind="bn", c=16)]           # BN channel mismatch
    with pytest.raises(dsr.DcgansrError):
        dsr.Sequential.from_specs(bad).cuda(None, (3, 8, 8), 2)
    with pytest.raises(dsr.DcgansrError):
        dsr.Sequential.from_specs([dict(kind="conv", cin=3, cout=8, k=9, s=1, p=0)]).cuda(None, (3, 16, 16), 2)   # 9x9 kernel
    with pytest.raises(dsr.DcgansrError):
        dsr.Sequential.from_specs([dict(kind="conv", cin=4, cout=8, k=3, s=1, p=0)]).cuda(None, (3, 16, 16), 2)   # cin mismatch


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: this checks the no-GPU behaviour")
    with pytest.raises(dsr.DcgansrError, match="no CPU fallback"):
        dsr.Context(device=0)
    net = dsr.Sequential.from_specs(models.patch_D(8)).cuda(None, (1, 8, 8), 2)
    with pytest.raises(dsr.DcgansrError):
        net.forward(np.zeros((2, 1, 8, 8), np.float32))       # plan-only net cannot compute
    net.close()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dcgan_super_resolution_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_lua_shim_only_calls_declared_entry_points():
    """Every lib.dcgansr_* the LuaJIT shim calls is declared inside the header's ffi.cdef block (the shim cannot be executed
    here: no LuaJIT), with the argument count the declaration has."""
    src = open(HEADER).read()
    b = src.index("\n", src.index("FFI-CDEF-BEGIN")) + 1
    e = src.index("/* FFI-CDEF-END")
    cdef = re.sub(r"/\*.*?\*/", "", src[b:e], flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(dcgansr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", cdef, flags=re.S)}
    lua = open(os.path.join(ROOT, "lua", "dcgansr.lua")).read()
    lua = re.sub(r"--[^\n]*", "", lua)
    used = set(re.findall(r"lib\.(dcgansr_[a-z0-9_]+)", lua))
    assert len(used) >= 20
    for name in sorted(used):
        assert name in decl, f"lua/dcgansr.lua calls {name}, which the ffi.cdef block does not declare"

    def nargs_decl(params):
        params = params.strip()
        return 0 if params in ("", "void") else params.count(",") + 1

    def call_args(text, pos):          # argument count of the call whose '(' is at pos
        depth, n, i, seen = 0, 0, pos, False
        while i < len(text):
            ch = text[i]
            if ch in "({":
                depth += 1
            elif ch in ")}":
                depth -= 1
                if depth == 0:
                    return n + 1 if seen else 0
            elif ch == "," and depth == 1:
                n += 1
            elif depth >= 1 and not ch.isspace():
                seen = True
            i += 1
        raise AssertionError("unbalanced call")

    for m in re.finditer(r"lib\.(dcgansr_[a-z0-9_]+)\s*\(", lua):
        assert call_args(lua, m.end() - 1) == nargs_decl(decl[m.group(1)]), f"argument count of {m.group(1)} in lua/dcgansr.lua"


LUA_HOSTS = ["train.lua", "train-gray.lua", "train-gray-patch.lua", "train-gray-patch-batch-overlap.lua"]


@pytest.mark.parametrize("host", LUA_HOSTS)
def test_lua_hosts_use_only_the_shim_surface(host):
    """The Lua hosts of the BASELINE-named scripts (they cannot be executed here: no LuaJIT) only use names the shim exports,
    and build the same layer lists as the Python spec builders the parity tests run (so the graphs the GPU tests cover ARE
    the graphs these hosts create)."""
    shim = re.sub(r"--[^\n]*", "", open(os.path.join(ROOT, "lua", "dcgansr.lua")).read())
    exported = set(re.findall(r"function M\.([A-Za-z_][A-Za-z0-9_]*)", shim)) | {"nn", "optim", "lib"}
    nn_exported = set(re.findall(r"function M\.nn\.([A-Za-z_][A-Za-z0-9_]*)", shim))
    src = open(os.path.join(ROOT, "lua", host)).read()
    src = re.sub(r"--\[\[.*?\]\]", "", src, flags=re.S)
    src = re.sub(r"--[^\n]*", "", src)
    for name in set(re.findall(r"\bdsr\.([A-Za-z_][A-Za-z0-9_]*)", src)):
        assert name in exported, f"lua/{host} uses dsr.{name}, which lua/dcgansr.lua does not export"
    for name in set(re.findall(r"\bnn\.([A-Za-z_][A-Za-z0-9_]*)", src)):
        assert name in nn_exported, f"lua/{host} uses nn.{name}, which lua/dcgansr.lua does not export"
    # layer lists: count the constructor calls per net and compare with the spec builders
    want = {"train.lua": (models.train_lua_G(3, 12), models.dcgan64_D(3, 64)),
            "train-gray.lua": (models.train_gray_G(16), models.dcgan64_D(1, 64)),
            "train-gray-patch.lua": (models.train_gray_3_G(16), models.patch_D(64)),
            "train-gray-patch-batch-overlap.lua": (models.train_gray_3_G(16), models.patch_D(64))}[host]
    lua_kind = {"SpatialConvolution": "conv", "SpatialFullConvolution": "fullconv", "SpatialBatchNormalization": "bn", "ReLU": "relu",
                "LeakyReLU": "lrelu", "Tanh": "tanh", "Sigmoid": "sigmoid", "SpatialUpSamplingNearest": "upnearest", "View": "view"}
    for var, specs in zip(("netG", "netD"), want):
        kinds = []
        for ln in src.splitlines():
            if re.match(rf"\s*{var}:add\(", ln):
                kinds += [lua_kind[k] for k in re.findall(r"(?:nn\.)?\b(SpatialFullConvolution|SpatialConvolution|SpatialBatchNormalization|"
                                                          r"SpatialUpSamplingNearest|LeakyReLU|ReLU|Tanh|Sigmoid|View)\(", ln)]
        assert kinds == [s["kind"] for s in specs], (host, var, kinds)


def test_build_compiles_every_cuda_source():
    """Every .cu under csrc/ is in the build list (a kernel file left out would only show up as an unresolved symbol at link time,
    or -- with a weak fallback -- not at all)."""
    import os

    from dcgan_super_resolution_b200 import build as b
    here = os.path.dirname(os.path.abspath(b.__file__))
    on_disk = sorted(f for f in os.listdir(os.path.join(here, "csrc")) if f.endswith(".cu"))
    assert sorted(b.SOURCES) == on_disk, (sorted(b.SOURCES), on_disk)
    assert "arch=compute_100a,code=sm_100a" in " ".join(b.NVCC_FLAGS) and "-lineinfo" in b.NVCC_FLAGS


def test_comm_peer_enabled_is_zero_without_a_communicator():
    """dcgansr_comm_peer_enabled answers 0 for a NULL context (no device needed): the peer-memory exchange of the sync_bn
    statistics (kernels_peer.cu) only exists behind dcgansr_comm_init on a multi-rank context."""
    from dcgan_super_resolution_b200 import _lib as L
    lib = L.load()
    assert lib.dcgansr_comm_peer_enabled(None) == 0
