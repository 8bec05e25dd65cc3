"""GPU, TWO devices (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the cross-GPU paths on real hardware over NCCL -- parallel-tempering swap rounds with ladders strided across the ranks
against oracle/exchange.py, the best-structure all-gather, and optimize(ladder=...) with seed=None (every rank must end
with the same structure and the same temperature table)."""
import os
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from so_b200 import engine, parallel as P, LSC_cat, optimize
    from oracle import lattice as L, surrogate as S, sa as SA, exchange as OX
    import random
    d, N, n, ladder, T_ref = 12, 144, 64, 8, 0.4
    p = S.SurrogateParams.random_init(L.local_offsets(3), H=20, seed=11)
    lat = engine.Lattice(d, "LSC", device=rank)
    tab = engine.SurrogateTables(lat, p.offsets, p.W1, p.b1, p.w2, p.b2, p.y_mean, p.y_scale)
    ch = engine.Chains(lat, n)
    ch.randomize(seed=50 + rank, coverage=0.5)
    ch.attach(tab)
    pt = P.ReplicaExchange(ch, ladder=ladder, seed=99, rank=rank, world=world)
    gather = P.BestGather(ch, rank * n, world)
    ts0 = pt.tscale_all.cpu().numpy().copy()
    log = []
    for r in range(6):
        ch.anneal(np.full(60, T_ref), step0=60 * r, seed=99, chain_id0=rank * n)
        before = pt.tscale_all.cpu().numpy().copy()
        pt.round(r, t_ref=T_ref)
        of_all = pt.of_all.cpu().numpy()
        after = pt.tscale_all.cpu().numpy()
        want = OX.exchange_round(of_all, before, world, n, ladder, r & 1, 99, r, T_ref)
        assert np.array_equal(want, after), "swap decisions differ from oracle/exchange.py in round %d" % r
        assert np.array_equal(ch.get_tscale(), after[rank * n:(rank + 1) * n])
        assert np.array_equal(of_all[rank * n:(rank + 1) * n], ch.objective())
        log.append(after.copy())
    assert sorted(log[-1].tolist()) == sorted(ts0.tolist())          # temperatures are permuted, never lost
    assert not np.array_equal(log[-1], ts0)                            # ... and some did move across the two GPUs
    best = gather.gather()
    of_local = ch.objective()
    # optimize() through the drop-in call, seed=None: rank 0's draw is broadcast
    random.seed(1000 + rank)
    cat = LSC_cat(d, device=rank)
    random.seed(7)
    cat.randomize(coverage=0.5)
    random.seed(2000 + rank)                                           # the ranks' generators disagree from here on

    class Sur(object):
        pass
    sur = Sur()
    sur.predictor = Sur()
    sur.predictor.coefs_, sur.predictor.intercepts_ = [p.W1, p.w2.reshape(-1, 1)], [p.b1, np.array([p.b2])]
    sur.predictor.out_activation_ = 'identity'
    sur.Y_scaler, sur.classifier = None, None
    st = optimize(cat, surrogate=sur, n_cycles=4, T_0=0.4, n_chains=32, ladder=8, exchange_every=72, gather_every=144,
                  exact_guard=False, return_stats=True)
    out[rank] = dict(tables=[t.tobytes() for t in log], best=(best[0], best[1], best[2].tobytes()), of_max=float(of_local.max()),
                     x=list(cat.variable_occs), seed=st["seed"], OF=st["OF"], best_chain=st["best_chain"],
                     swaps=st["temperature_swaps"])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_replica_exchange_and_best_gather_across_two_gpus():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29631, out), nprocs=world, join=True)
        r0, r1 = dict(out[0]), dict(out[1])
    assert r0["tables"] == r1["tables"]                               # same temperature table on both ranks, every round
    assert r0["best"] == r1["best"] and r0["best"][0] == max(r0["of_max"], r1["of_max"])
    assert r0["seed"] == r1["seed"]                                   # one Philox key although each rank passed seed=None
    assert r0["x"] == r1["x"] and r0["OF"] == r1["OF"] and r0["best_chain"] == r1["best_chain"]
    assert r0["swaps"] + r1["swaps"] > 0
