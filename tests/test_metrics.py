"""calPSNR / calSSIM: oracle known answers on the CPU, GPU parity through the C ABI."""
import numpy as np
import pytest

from oracle import metrics as om
from util import rng


def test_oracle_known_answers():
    x = rng(1).uniform(-1, 1, (64, 64))
    assert om.psnr(x, x) == 99.0                                   # MSE == 0 branch (train-gray-3.lua:148)
    assert abs(om.psnr(x, x + 0.1) - 20.0) < 1e-9                   # MSE = 0.01 -> 20 dB
    assert abs(om.ssim(x, x) - 1.0) < 1e-12                         # "If img1 = img2, then mssim = 1" (:180)
    w = om.gaussian_window()
    assert w.shape == (11, 11) and abs(w.sum() - 1) < 1e-12 and np.allclose(w, w.T) and w[5, 5] == w.max()
    # sigma: w[5][6] / w[5][5] = exp(-(1/1.5)^2 / 2)
    assert abs(w[5, 6] / w[5, 5] - np.exp(-(1 / 1.5) ** 2 / 2)) < 1e-12
    y = np.clip(x + rng(2).normal(0, 0.2, x.shape), -1, 1)
    assert 0.0 < om.ssim(x, y) < 1.0
    # 'full' convolution: (H + 10) x (W + 10) map, total mass preserved
    assert om.convolve_full(np.ones((4, 5)), w).shape == (14, 15) and abs(om.convolve_full(np.ones((4, 5)), w).sum() - 20) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 64, 64), (2, 32, 48), (1, 8, 8)])
def test_gpu_psnr_ssim(ctx, shape):
    import dcgan_super_resolution_b200 as dsr
    r = rng(hash(shape) % 2**31)
    a = r.uniform(-1, 1, shape).astype(np.float32)
    b = np.clip(a + r.normal(0, 0.15, shape), -1, 1).astype(np.float32)
    b[0] = a[0]                                                    # one identical pair: PSNR 99, SSIM 1
    p, s = dsr.psnr(ctx, a, b), dsr.ssim(ctx, a, b)
    for i in range(shape[0]):
        assert abs(p[i] - om.psnr(a[i], b[i])) <= 1e-4 * max(1.0, abs(om.psnr(a[i], b[i])))
        assert abs(s[i] - om.ssim(a[i], b[i])) <= 2e-5
    assert p[0] == 99.0 and abs(s[0] - 1.0) <= 2e-6


def test_oracle_scale_bilinear_properties():
    """End points aligned, exact on linear ramps, identity at equal size, agrees with torch's align_corners=True bilinear
    (an independent implementation of the same definition) to float32 rounding."""
    import torch
    import torch.nn.functional as F
    x = rng(3).uniform(0, 1, (32, 32)).astype(np.float32)
    y = om.scale_bilinear(x, 64, 64)
    assert y.shape == (64, 64) and y.dtype == np.float32
    assert y[0, 0] == x[0, 0] and y[-1, -1] == x[-1, -1] and y[0, -1] == x[0, -1] and y[-1, 0] == x[-1, 0]
    assert np.array_equal(om.scale_bilinear(x, 32, 32), x)
    ref = F.interpolate(torch.from_numpy(x)[None, None].double(), size=(64, 64), mode="bilinear", align_corners=True)[0, 0].numpy()
    assert np.abs(y - ref).max() < 1e-5
    ramp = np.tile(np.linspace(0, 1, 8, dtype=np.float32), (4, 1))
    up = om.scale_bilinear(ramp, 4, 15)
    assert np.abs(up - np.tile(np.linspace(0, 1, 15), (4, 1))).max() < 1e-6
    assert om.scale_bilinear(np.full((1, 1), 0.25, np.float32), 3, 5).tolist() == [[0.25] * 5] * 3


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [(2, 32, 32, 64, 64), (1, 4, 4, 8, 8), (3, 5, 7, 9, 20), (1, 6, 6, 6, 11), (2, 1, 3, 4, 3)])
def test_gpu_scale_bilinear_bit_exact(ctx, geom):
    import dcgan_super_resolution_b200 as dsr
    n, h, w, dh, dw = geom
    x = rng(hash(geom) % 2**31).uniform(-1, 1, (n, h, w)).astype(np.float32)
    got = dsr.scale_bilinear(ctx, x, dh, dw)
    want = np.stack([om.scale_bilinear(x[i], dh, dw) for i in range(n)])
    assert np.array_equal(got, want)
    with pytest.raises(dsr.DcgansrError):
        dsr.scale_bilinear(ctx, x, h - 1 if h > 1 else 0, dw)
