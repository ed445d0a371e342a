"""2-GPU data-parallel checks (run under torchrun, one rank per GPU):
  (1) gradient buckets overlapped with the backward walk give bit-identical parameters to one all-reduce per net;
  (2) sync_bn=1 over 2 shards of B == a single-GPU step on the 2B batch (exact big-batch semantics) within 1e-5.
Prints one line per check on rank 0."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dcgan_super_resolution_b200 as dsr
from dcgan_super_resolution_b200 import init, models, parallel

rank, local_rank, world = parallel.env_rank()
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
cfg = models.config("C2")
B = 8
specsG, specsD = models.train_gray_G(8), models.dcgan64_D(1, 8)
step = dsr.make_step_cfg(**cfg["step"])
rng = np.random.Generator(np.random.Philox(99))
full = [rng.uniform(-1, 1, (world * B, 1, 64, 64)).astype(np.float32) for _ in range(3)]


def run(world_size, precision, sync_bn, env=None, batch=B, shard=True, graph=False):
    for k, v in (env or {}).items():
        os.environ[k] = v
    ctx = dsr.Context(device=local_rank, precision=precision, world_size=world_size, rank=rank if world_size > 1 else 0, sync_bn=sync_bn,
                      use_graph=graph)
    if world_size > 1:
        parallel.exchange_unique_id(ctx, dist, device=torch.device("cuda", local_rank))
    G = dsr.Sequential.from_specs(specsG).cuda(ctx, (1, 32, 32), batch)
    D = dsr.Sequential.from_specs(specsD).cuda(ctx, (1, 64, 64), 2 * batch)
    G.set_params(init.weights_init(specsG, 4321)); D.set_params(init.weights_init(specsD, 8765))
    losses = []
    for x in full:
        xb = x[rank * batch:(rank + 1) * batch] if shard else x
        losses.append(dsr.train_step(ctx, G, D, step, xb))
    out = (G.get_params(), D.get_params(), losses)
    G.close(); D.close(); ctx.close()
    for k in (env or {}):
        os.environ.pop(k, None)
    return out


for prec in ("strict", "tf32"):
    a = run(world, prec, False)
    b = run(world, prec, False, env={"DCGANSR_NO_OVERLAP": "1"})
    c = run(world, prec, False, graph=True)
    if rank == 0:
        print(f"[{prec}] overlap vs single all-reduce: G identical {np.array_equal(a[0], b[0])}, D identical {np.array_equal(a[1], b[1])}; "
              f"graph replay identical {np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])}", flush=True)
dist.barrier()
a = run(world, "strict", True)
if rank == 0:
    ref = run(1, "strict", False, batch=world * B, shard=False)
    eG = np.max(np.abs(a[0] - ref[0])) / np.max(np.abs(ref[0]))
    eD = np.max(np.abs(a[1] - ref[1])) / np.max(np.abs(ref[1]))
    print(f"[strict] sync_bn dp{world} vs single GPU on the {world * B}-batch: rel err G {eG:.2e} D {eD:.2e}; losses {a[2][-1]} vs {ref[2][-1]}", flush=True)
dist.barrier()
dist.destroy_process_group()
os._exit(0)
