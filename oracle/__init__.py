"""CPU oracle for the DCGAN-SR training step  --  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU float64 / float32 functional ops plus
an independent naive numpy implementation) of the Torch7 semantics the reference's
``train*.lua`` scripts rely on.  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``dcgan_super_resolution_b200`` and
``libdcgansr.so``) never imports, links or executes anything in here and fails loudly
when its CUDA library is missing.

PARITY UNPINNED.  The reference (/root/reference, 100 % Lua / Torch7) cannot be run in
this environment (no Lua, LuaJIT or Torch7) and it ships no tests, golden vectors or
recorded outputs for the hot path.  The arithmetic lives
in un-vendored, un-pinned third-party packages: torch/torch7, torch/nn (THNN),
torch/cunn (THCUNN), torch/optim (adam.lua), soumith/cudnn.torch.  This oracle restates
their published algorithms and follows the reference's own call
sites: ``train.lua:97-152,208-283`` and the variants listed in SURVEY.md App. A.
The golden vectors under ``tests/golden`` are produced by *this* oracle
(``tests/golden/make_golden.py``), not by Torch7.
"""
from . import ops, nets, step, naive  # noqa: F401
