"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (chain-id layout, ladder temperatures,
record gathering and the deterministic global pick) -- the device kernels are covered by the -m gpu tests."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from so_b200 import parallel as P
    n_local, words, ladder = 8, 5, 4
    # each rank fabricates its best-structure record, as so_best_record_dev would write it
    rng = np.random.default_rng(100 + rank)
    of_local = rng.random(n_local)
    best = int(np.argmax(of_local))
    rec = np.zeros(P.record_bytes(words), dtype=np.uint8)
    rec[0:8] = np.frombuffer(np.float64(of_local[best]).tobytes(), dtype=np.uint8)
    rec[8:16] = np.frombuffer(np.int64(rank * n_local + best).tobytes(), dtype=np.uint8)
    rec[16:] = np.frombuffer(np.full(words, rank + 1, dtype=np.uint32).tobytes(), dtype=np.uint8)
    gathered = torch.zeros((world, len(rec)), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(rec))
    of, cid, bits = P.parse_records(gathered.numpy(), words)
    pick = P.pick_best(of, cid, bits)
    # objectives all-gather in rank-major layout + interleaved global ids
    of_all = torch.zeros(world * n_local, dtype=torch.float64)
    dist.all_gather_into_tensor(of_all, torch.from_numpy(of_local))
    g = P.global_chain_id(rank, world, np.arange(n_local))
    pos = P.gathered_position(g, world, n_local)
    assert (of_all.numpy()[pos] == of_local).all()
    out[rank] = (pick[0], pick[1], pick[2].tolist(), P.ladder_tscale(g, ladder).tolist())
    dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29613, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[:3] == r1[:3]                                           # every rank picks the same global best
    ofs = [np.random.default_rng(100 + r).random(8) for r in range(world)]
    best_rank = int(np.argmax([o.max() for o in ofs]))
    assert r0[0] == max(o.max() for o in ofs) and r0[1] == best_rank * 8 + int(np.argmax(ofs[best_rank]))
    assert r0[2] == [best_rank + 1] * 5
    # ladders of 4 consecutive GLOBAL ids are strided across the two ranks: rank r holds rungs r, r+2 of each ladder
    np.testing.assert_allclose(r0[3][:2], [1.0, np.exp(-5 * 2 / 3.0)])
    np.testing.assert_allclose(r1[3][:2], [np.exp(-5 * 1 / 3.0), np.exp(-5.0)])


def test_layout_helpers():
    from so_b200 import parallel as P
    world, n_local = 4, 6
    g = np.concatenate([P.global_chain_id(r, world, np.arange(n_local)) for r in range(world)])
    assert sorted(g.tolist()) == list(range(world * n_local))
    assert (P.gathered_position(g, world, n_local) == np.arange(world * n_local)).all()
    of = np.array([1.0, 3.0, 3.0, 2.0])
    cid = np.array([9, 7, 4, 1])
    assert P.pick_best(of, cid, np.arange(8).reshape(4, 2))[:2] == (3.0, 4)      # tie -> lowest chain id


def _seed_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    from so_b200 import optimize_SA as O
    random.seed(1000 + rank)                                  # the ranks' own generators disagree ...
    drawn = O._shared_seed(None, rank, world)                 # ... the key they anneal and swap with must not
    given = O._shared_seed(5 + rank, rank, world)             # a per-rank argument is overridden by rank 0's as well
    out[rank] = (drawn, given, O._rank_world())
    dist.destroy_process_group()


def test_every_rank_uses_the_same_philox_key():
    """Parallel-tempering swap decisions are recomputed on every rank from the run's seed: with seed=None each process
    would draw its own; optimize() broadcasts rank 0's."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_seed_worker, args=(world, 29617, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] == r1[0] and r0[1] == r1[1] == 5
    assert r0[2] == (0, 2) and r1[2] == (1, 2)
    import random
    random.seed(1000)
    assert r0[0] == random.getrandbits(63)                    # rank 0's draw from the global generator, as upstream seeds


def test_exchange_oracle_ladder_subset_equals_full_round():
    from oracle import exchange as OX
    rng = np.random.default_rng(4)
    world, n_local, ladder = 2, 32, 8
    of = rng.normal(size=world * n_local)
    ts = np.exp(-5.0 * ((np.arange(world * n_local) // 1) % ladder) / (ladder - 1))
    full = OX.exchange_round(of, ts, world, n_local, ladder, 1, 77, 3, 0.05)
    part = OX.exchange_round(of, ts, world, n_local, ladder, 1, 77, 3, 0.05, ladders=[1, 5])
    pos = [OX.gathered_position(l * ladder + i, world, n_local) for l in (1, 5) for i in range(ladder)]
    assert (full[pos] == part[pos]).all() and sorted(full.tolist()) == sorted(ts.tolist())
    rest = np.setdiff1d(np.arange(world * n_local), pos)
    assert (part[rest] == ts[rest]).all()
