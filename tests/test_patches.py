"""Patch extraction / re-assembly: oracle properties on the CPU, bit-exact GPU parity through the C ABI."""
import numpy as np
import pytest

from oracle import patches as op
from util import rng

# (K images, fine, patch, overlap): the reference's own geometry (64, 8, none / overlap 4) and smaller relatives
GEOMS = [(3, 64, 8, 0), (2, 64, 8, 4), (2, 16, 4, 0), (1, 16, 4, 2), (2, 36, 6, 0)]


def _geom(fine, patch, overlap):
    from dcgan_super_resolution_b200.nn import _patch_geom
    return _patch_geom(fine, patch, overlap)


def test_reference_index_formula_only_tiles_when_fine_is_patch_squared():
    """train-gray-patch.lua:267-273 divides the patch index by patchSize: with fine = 64, patch = 8 that is the patches per
    row and the 64 patches tile the image exactly (assemble(extract(x)) == x); with fine = 36, patch = 6 likewise; the
    oracle reproduces the loop as written."""
    for fine, patch in ((64, 8), (36, 6)):
        line, nper, stride = _geom(fine, patch, 0)
        x = rng(1).uniform(0, 1, (2, fine, fine)).astype(np.float32)
        p = op.extract(x, patch, line, nper, stride)
        assert p.shape == (2 * nper, patch, patch)
        assert np.array_equal(op.assemble(p, np.zeros_like(x), patch, line, nper, stride), x)
    # patch index -> origin, as in the comment of train-gray-patch-batch-overlap.lua:391 ("2 -> (0,4), 16 -> (4,0), 255 -> (56,56)")
    line, nper, stride = _geom(64, 8, 4)
    assert (line, nper, stride) == (15, 225, 4)
    x = np.arange(64 * 64, dtype=np.float32).reshape(1, 64, 64)
    p = op.extract(x, 8, line, nper, stride)
    assert p[1, 0, 0] == x[0, 0, 4] and p[15, 0, 0] == x[0, 4, 0] and p[224, 0, 0] == x[0, 56, 56]


def test_overlap_assembly_last_patch_wins():
    line, nper, stride = _geom(16, 4, 2)
    x = rng(2).uniform(0, 1, (1, 16, 16)).astype(np.float32)
    p = op.extract(x, 4, line, nper, stride)
    p2 = p + np.arange(p.shape[0], dtype=np.float32)[:, None, None]        # tag each patch
    out = op.assemble(p2, np.full_like(x, -1.0), 4, line, nper, stride)
    # pixel (3, 3) is covered by patches 0, 1, line, line+1: the highest index wins
    assert out[0, 3, 3] == p2[line + 1, 1, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOMS)
def test_gpu_extract_assemble_bit_exact(ctx, geom):
    import dcgan_super_resolution_b200 as dsr
    K, fine, patch, overlap = geom
    line, nper, stride = _geom(fine, patch, overlap)
    x = rng(hash(geom) % 2**31).uniform(0, 1, (K, fine, fine)).astype(np.float32)
    p = dsr.extract_patches(ctx, x, patch, line, nper, stride)
    assert np.array_equal(p, op.extract(x, patch, line, nper, stride))
    tagged = p + np.arange(p.shape[0], dtype=np.float32)[:, None, None]
    base = np.full_like(x, -1.0)
    assert np.array_equal(dsr.assemble_patches(ctx, tagged, base, patch, line, nper, stride),
                          op.assemble(tagged, base, patch, line, nper, stride))


@pytest.mark.gpu
def test_gpu_stage_patches_feeds_the_step(ctx):
    """stage_patches == stage_batch of the oracle's patches: the fused step sees the same batch (train-gray-patch.lua:264-275)."""
    import dcgan_super_resolution_b200 as dsr
    from dcgan_super_resolution_b200 import init, models
    specsG, specsD = models.train_gray_3_G(4), models.patch_D(8)
    line, nper, stride = _geom(64, 8, 0)
    img = rng(7).uniform(0, 1, (1, 64, 64)).astype(np.float32)
    B = nper
    step = dsr.make_step_cfg(family="bce", real_label=1.0, fake_label=0.0, gen_label=1.0)
    res = []
    for use_patches in (False, True):
        G = dsr.Sequential.from_specs(specsG).cuda(ctx, (1, 4, 4), B)
        D = dsr.Sequential.from_specs(specsD).cuda(ctx, (1, 8, 8), 2 * B)
        G.set_params(init.weights_init(specsG, 4321))
        D.set_params(init.weights_init(specsD, 8765))
        if use_patches:
            assert dsr.stage_patches(ctx, D, img, 8, line, nper, stride, 0) == B
        else:
            dsr.stage_batch(ctx, D, op.extract(img, 8, line, nper, stride)[:, None], 0)
        res.append((dsr.train_step_staged(ctx, G, D, step, 0, B, want_losses=True), G.get_params()))
        G.close(); D.close()
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
