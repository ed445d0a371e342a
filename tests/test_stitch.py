"""Overlap stitching by minimum-error boundary cut: oracle properties on the CPU, bit-exact GPU parity."""
import numpy as np
import pytest

from oracle import patches as op
from oracle import stitch as ost
from util import rng

# (fine, patch, overlap): the reference's geometry (train-gray-patch-batch-overlap.lua:19-21) and smaller relatives
GEOMS = [(64, 8, 4), (16, 4, 2), (24, 8, 4), (36, 12, 6)]


def _patches_of(img, patch, overlap):
    fine = img.shape[-1]
    L = ost.line_count(fine, patch, overlap)
    return op.extract(img[None], patch, L, L * L, overlap), L


def _noisy_patches(seed, fine, patch, overlap, noise=0.05, k=1):
    """Overlapping patches of smooth images with per-patch noise (so neighbouring patches disagree on their shared strip,
    as independently generated patches do)."""
    g = rng(seed)
    out = []
    for _ in range(k):
        img = g.uniform(0, 1, (fine, fine)).astype(np.float32)
        p, _ = _patches_of(img, patch, overlap)
        out.append(p + g.normal(0, noise, p.shape).astype(np.float32))
    return np.concatenate(out)


@pytest.mark.parametrize("geom", GEOMS)
def test_consistent_patches_stitch_back_to_the_image(geom):
    """If the patches agree on every overlap strip (cut costs all zero) any cut reproduces the image."""
    fine, patch, overlap = geom
    img = rng(3).uniform(0, 1, (fine, fine)).astype(np.float32)
    p, L = _patches_of(img, patch, overlap)
    assert np.array_equal(ost.stitch(p, fine, patch, overlap), img)


def test_seam_follows_the_cheapest_path_and_tie_rules():
    # one obvious valley: the cut must run through it
    delta = np.ones((6, 4))
    valley = [2, 2, 1, 1, 2, 3]
    for a, b in enumerate(valley):
        delta[a, b] = 0.0
    assert list(ost._seam(delta)) == [b + 1 for b in valley]
    # all-equal costs: the start is the LAST minimum of the final line (the reference's loop keeps overwriting) and the
    # back-track stays in the same column
    assert list(ost._seam(np.ones((5, 4)))) == [4, 4, 4, 4, 4]
    # every written pixel is a copy of a generated pixel, and column 1 of a left seam always comes from the neighbour
    p = _noisy_patches(5, 16, 4, 2)
    out = ost.stitch(p, 16, 4, 2)
    assert np.isin(out, p).all()
    L = ost.line_count(16, 4, 2)
    assert out[0, 2] == p[0, 0, 2] and out[0, 2 * (L - 1)] == p[L - 2, 0, 2]


def test_interior_top_seams_are_dead_and_top_cost_quirk_matters_only_in_column_one():
    """An interior patch's left-seam write covers its whole footprint again (train-gray-patch-batch-overlap.lua:684-691), so
    only the first column of patches shows a top seam; the reference takes that seam's cost against patch i-1 (:557)."""
    fine, patch, overlap = 24, 8, 4
    p = _noisy_patches(11, fine, patch, overlap, noise=0.2)
    a = ost.stitch(p, fine, patch, overlap)
    b = ost.stitch(p, fine, patch, overlap, fix_top_cost=True)
    assert np.array_equal(a[:, overlap:], b[:, overlap:])      # beyond the first `overlap` columns nothing can differ
    assert not np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("fix", [False, True])
def test_gpu_stitch_bit_exact(ctx, geom, fix):
    import dcgan_super_resolution_b200 as dsr
    fine, patch, overlap = geom
    k = 3
    p = _noisy_patches(hash(geom) % 2**31, fine, patch, overlap, noise=0.1, k=k)
    got = dsr.stitch_overlap(ctx, p, fine, patch, overlap, fix_top_cost=fix)
    n = p.shape[0] // k
    want = np.stack([ost.stitch(p[j * n:(j + 1) * n], fine, patch, overlap, fix_top_cost=fix) for j in range(k)])
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_stitch_ties_and_generated_patches(ctx):
    """Quantised patches make exact cost ties common (tie-breaking parity); consistent patches give the image back; bad
    geometry is an error, not a crash."""
    import dcgan_super_resolution_b200 as dsr
    fine, patch, overlap = 64, 8, 4
    p = np.round(_noisy_patches(21, fine, patch, overlap, noise=0.2, k=2) * 4) / 4
    n = p.shape[0] // 2
    want = np.stack([ost.stitch(p[j * n:(j + 1) * n], fine, patch, overlap) for j in range(2)])
    assert np.array_equal(dsr.stitch_overlap(ctx, p, fine, patch, overlap), want)
    img = rng(9).uniform(0, 1, (fine, fine)).astype(np.float32)
    pc, _ = _patches_of(img, patch, overlap)
    assert np.array_equal(dsr.stitch_overlap(ctx, pc, fine, patch, overlap)[0], img)
    with pytest.raises(dsr.DcgansrError):
        dsr.stitch_overlap(ctx, np.zeros((25, 8, 8), np.float32), 31, 8, 3)      # (31-3) % (8-3) != 0
