"""N >= 2 GPUs under torchrun: bench.py's dp_check alone (bucket overlap, graph replay, sync_bn vs the single-GPU step, and the
one-shot peer-memory all-reduce of the sync_bn statistics against ncclAllReduce, with the sync_bn step time both ways).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/peer_ar_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

env = bench.Env()
res = bench.dp_check(env, quick="--quick" in sys.argv)
env.close()
if env.rank == 0:
    print(json.dumps(res))
