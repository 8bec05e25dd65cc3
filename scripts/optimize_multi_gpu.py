"""optimize() under torchrun: one process per GPU, chains sharded over the ranks, parallel-tempering ladders strided
across the GPUs, every rank ends with the same global best structure.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/optimize_multi_gpu.py"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
from so_b200 import LSC_cat, optimize
import helpers as Hp
from test_gpu_api import duck_surrogate
p = Hp.golden_surrogate(gated=True)
cat = LSC_cat(12, device=local)
random.seed(0)
cat.randomize(coverage=0.5)
out = optimize(cat, surrogate=duck_surrogate(p), n_cycles=50, T_0=0.05, n_chains=1024, seed=3, ladder=8, gather_every=1440,
               exact_guard=False, return_stats=True)
x = torch.tensor(cat.variable_occs, device="cuda", dtype=torch.int32)
xs = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
dist.all_gather(xs, x)
assert all(bool((v == xs[0]).all()) for v in xs), "ranks disagree on the best structure"
print("rank %d/%d: best OF %.6f from global chain %d, %d temperature swaps on this rank, kernels %.1f ms" % (
    out["rank"], out["world"], out["OF"], out["best_chain"], out["temperature_swaps"], out["kernel_ms"]), flush=True)
dist.destroy_process_group()
