import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import dcgan_super_resolution_b200 as dsr
from test_gpu_microbatch import CASES, _run
from util import rel_err
case=CASES['bce_patch']
cs=dsr.Context(device=0, precision='strict'); cf=dsr.Context(device=0, precision='tf32')
for steps in (1,2):
    ref=_run(cs, case, 16, False, steps)
    for nm,c,gb in (('strict k4',cs,4),('tf32 plain',cf,16),('tf32 k2',cf,8),('tf32 k4',cf,4)):
        g=_run(c, case, gb, False, steps)
        print(steps, nm, {k: float('%.2e'%rel_err(g[k], ref[k])) for k in ('gG','pG','pD','bnG')})
