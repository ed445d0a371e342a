"""weights_init (train.lua:42-51) on the host: conv / full-conv weights ~ N(0, 0.02) and bias-free,
BatchNorm gamma ~ N(1, 0.02), beta = 0, returned as the flat vector of Module:getParameters()
(train.lua:202-203).  The reference seeds from an unseeded RNG (train.lua:30-32), so the generator
is ours: numpy Philox, values rounded to float32."""
from __future__ import annotations

import numpy as np


def param_shapes(specs):
    """[(name, shape)] in getParameters() order."""
    out = []
    for i, s in enumerate(specs):
        if s["kind"] == "conv":
            out.append((f"{i}.weight", (s["cout"], s["cin"], s["k"], s["k"])))
        elif s["kind"] == "fullconv":
            out.append((f"{i}.weight", (s["cin"], s["cout"], s["k"], s["k"])))
        elif s["kind"] == "bn":
            out.append((f"{i}.weight", (s["c"],)))
            out.append((f"{i}.bias", (s["c"],)))
    return out


def weights_init(specs, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    parts = []
    for name, shape in param_shapes(specs):
        if name.endswith(".bias"):
            parts.append(np.zeros(shape, np.float32))
        elif len(shape) == 1:
            parts.append(rng.normal(1.0, 0.02, size=shape).astype(np.float32))
        else:
            parts.append(rng.normal(0.0, 0.02, size=shape).astype(np.float32))
    return np.concatenate([p.reshape(-1) for p in parts]) if parts else np.zeros(0, np.float32)
