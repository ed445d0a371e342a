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
