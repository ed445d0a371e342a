"""Object-lifetime hazards of the C ABI: the CUDA-graph cache of the fused step must never replay a graph that holds
pointers of a destroyed net or of a re-allocated context buffer, and contexts / nets may be destroyed in either order
(LuaJIT and Python finalizers are unordered)."""
import numpy as np
import pytest

import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import init, models
from util import rng, smooth_images

pytestmark = pytest.mark.gpu

SPECS_G, SPECS_D = models.train_gray_3_G(4), models.patch_D(8)
STEP = dict(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0)


def _nets(ctx, B, hr=8):
    G = dsr.Sequential.from_specs(SPECS_G).cuda(ctx, (1, hr // 2, hr // 2), B)
    D = dsr.Sequential.from_specs(SPECS_D).cuda(ctx, (1, hr, hr), 2 * B)
    G.set_params(init.weights_init(SPECS_G, 4321))
    D.set_params(init.weights_init(SPECS_D, 8765))
    return G, D


def _run(ctx, G, D, batches, staged=True):
    cfg = dsr.make_step_cfg(**STEP)
    out = []
    for x in batches:
        if staged:
            dsr.stage_batch(ctx, D, x, 0)
            out.append(dsr.train_step_staged(ctx, G, D, cfg, 0, x.shape[0], want_losses=True))
        else:
            out.append(dsr.train_step(ctx, G, D, cfg, x))
    return out, G.get_params(), D.get_params()


def test_graph_cache_survives_net_recreation():
    """Destroy and re-create the nets on ONE graph-replaying context (the allocator commonly hands back the same addresses):
    every generation must give the results of a fresh eager context, bit for bit."""
    r = rng(7)
    batches = [smooth_images(r, (16, 1, 8, 8), 0.0, 1.0) for _ in range(3)]
    ectx = dsr.Context(device=0, precision="strict", use_graph=False)
    G, D = _nets(ectx, 16)
    want = _run(ectx, G, D, batches)
    G.close(); D.close(); ectx.close()
    gctx = dsr.Context(device=0, precision="strict", use_graph=True)
    for generation in range(4):
        G, D = _nets(gctx, 16)
        got = _run(gctx, G, D, batches)
        assert np.array_equal(np.array(got[0]), np.array(want[0])), generation
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), generation
        G.close(); D.close()
    gctx.close()


def test_graph_cache_survives_batch_growth():
    """Batch sizes small, big, small on one graph-replaying context: the big step re-allocates the context's low-resolution /
    label buffers and the staged slot, which the small step's cached graph had captured."""
    r = rng(8)
    small = [smooth_images(r, (8, 1, 8, 8), 0.0, 1.0) for _ in range(2)]
    big = [smooth_images(r, (32, 1, 8, 8), 0.0, 1.0) for _ in range(2)]

    def seq(use_graph):
        ctx = dsr.Context(device=0, precision="strict", use_graph=use_graph)
        G, D = _nets(ctx, 32)
        res = []
        for batches in (small, big, small, big):
            res.append(_run(ctx, G, D, batches))
        G.close(); D.close(); ctx.close()
        return res

    for a, b in zip(seq(False), seq(True)):
        assert np.array_equal(np.array(a[0]), np.array(b[0]))
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_context_destroyed_before_its_nets():
    ctx = dsr.Context(device=0, precision="strict")
    G, D = _nets(ctx, 4)
    x = smooth_images(rng(9), (4, 1, 8, 8), 0.0, 1.0)
    dsr.train_step(ctx, G, D, dsr.make_step_cfg(**STEP), x)
    ctx.close()                      # releases the nets' device memory, leaves plan-only handles
    assert G.num_params() > 0        # shape queries still work
    with pytest.raises(dsr.DcgansrError):
        G.forward(np.zeros((4, 1, 4, 4), np.float32))
    G.close(); D.close()             # must not touch the destroyed context


def test_world_size_without_communicator_is_an_error():
    ctx = dsr.Context(device=0, precision="strict", world_size=2, rank=0)
    G, D = _nets(ctx, 4)
    x = smooth_images(rng(10), (4, 1, 8, 8), 0.0, 1.0)
    with pytest.raises(dsr.DcgansrError, match="comm_init"):
        dsr.train_step(ctx, G, D, dsr.make_step_cfg(**STEP), x)
    G.close(); D.close(); ctx.close()
