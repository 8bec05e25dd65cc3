"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Chains shard across
ranks with NO data-path collective; the only exchanges are latency-bound:

  * BestGather        periodic best-structure all-gather: one {f64 OF, i64 chain id, u32 bits[W]} record per rank
  * ReplicaExchange   parallel-tempering swap round: all-gather of the per-chain objectives (8 B per chain), then
                      every rank takes the same deterministic swap decisions on temperature, never moving states

Neither exists in the reference (it only scatters Python job lists over MPI ranks,
drivers/generate_init_structures.py:154-160); semantics are defined in SURVEY 8(e) and restated in
oracle/exchange.py.  The layout helpers below are pure host logic (tested at world_size 2 over gloo on CPU).
"""
import ctypes as C
import math
import numpy as np
from . import _lib as B


# ---- layout (pure host logic) -----------------------------------------------------------------------------
def global_chain_id(rank, world, local):
    """PT layout: global chain g = local*world + rank, so a ladder of consecutive ids is strided across ranks."""
    return np.asarray(local) * world + rank


def gathered_position(g, world, n_local):
    """Index of global chain g in an all-gathered (rank-major) [world][n_local] array."""
    g = np.asarray(g)
    return (g % world) * n_local + g // world


def ladder_tscale(global_ids, ladder, span=5.0):
    """Geometric temperature ladder in [T_ref*e^-span, T_ref]: rung = g % ladder, tscale = exp(-span*rung/(R-1))."""
    rung = np.asarray(global_ids) % ladder
    return np.exp(-span * rung / float(ladder - 1))


def record_bytes(words):
    return 16 + 4 * words


def parse_records(buf, words):
    """[world, record_bytes] uint8 -> (of[world] f64, chain_id[world] i64, bits[world, words] u32)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1, record_bytes(words))
    of = buf[:, 0:8].copy().view(np.float64).reshape(-1)
    cid = buf[:, 8:16].copy().view(np.int64).reshape(-1)
    bits = buf[:, 16:].copy().view(np.uint32).reshape(-1, words)
    return of, cid, bits


def pick_best(of, cid, bits):
    """Global arg-max; ties go to the lowest chain id so every rank agrees."""
    order = np.lexsort((cid, -of))
    i = int(order[0])
    return float(of[i]), int(cid[i]), bits[i].copy()


# ---- device-side drivers -----------------------------------------------------------------------------------
def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class BestGather(object):
    def __init__(self, chains, chain_id0, world, group=None):
        import torch
        self.torch = torch
        self.ch, self.chain_id0, self.world, self.group = chains, int(chain_id0), int(world), group
        self.words = chains.lattice.W
        dev = torch.device("cuda", chains.lattice.device)
        self.rec = torch.zeros(record_bytes(self.words), dtype=torch.uint8, device=dev)
        self.all = torch.zeros((self.world, record_bytes(self.words)), dtype=torch.uint8, device=dev)
        self.best = (-math.inf, -1, None)        # best-so-far over all gathers
        # overlapped form (gather_async / finish): a side stream for the collective and the copy to a pinned host buffer
        self.side = torch.cuda.Stream(device=dev)
        self.host = torch.zeros((self.world, record_bytes(self.words)), dtype=torch.uint8).pin_memory()
        self.done = torch.cuda.Event()
        self.pending = False

    def gather_async(self):
        """The same exchange off the critical path (collective on a side stream): the arg-max kernel runs now
        (it reads the chains, so the next sweep must not start before it: the host waits for it, a few tens of us), the
        all-gather and the device-to-host copy are queued on the side stream and overlap the next sweep; finish() returns the
        round's global best.  Rank skew is absorbed by the side stream instead of stalling the sweep."""
        torch = self.torch
        if self.pending:
            self.finish()
        with torch.cuda.stream(self.side):
            B.check(self.ch.lib.so_best_record_dev(self.ch.h, self.chain_id0, C.c_void_p(self.rec.data_ptr()),
                                                   C.c_void_p(self.side.cuda_stream)))
            self.side.synchronize()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.all.view(-1), self.rec, group=self.group)
            else:
                self.all[0].copy_(self.rec)
            self.host.copy_(self.all, non_blocking=True)
            self.done.record(self.side)
        self.pending = True

    def finish(self):
        """Wait for the pending gather_async() and fold its result into the best-so-far; returns the round's global best."""
        if not self.pending:
            return None
        self.done.synchronize()
        self.pending = False
        of, cid, bits = parse_records(self.host.numpy(), self.words)
        cur = pick_best(of, cid, bits)
        if cur[0] > self.best[0]:
            self.best = cur
        return cur

    def gather(self):
        """Arg-max kernel writes straight into the NCCL send buffer; returns the global best of this round."""
        torch = self.torch
        B.check(self.ch.lib.so_best_record_dev(self.ch.h, self.chain_id0, C.c_void_p(self.rec.data_ptr()), _stream_ptr()))
        torch.cuda.nvtx.range_push("best-structure all-gather (NCCL)")
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.all.view(-1), self.rec, group=self.group)
        else:
            self.all[0].copy_(self.rec)
        torch.cuda.nvtx.range_pop()
        of, cid, bits = parse_records(self.all.cpu().numpy(), self.words)
        cur = pick_best(of, cid, bits)
        if cur[0] > self.best[0]:
            self.best = cur
        return cur


class ReplicaExchange(object):
    def __init__(self, chains, ladder, seed, rank, world, span=5.0, group=None):
        import torch
        self.torch = torch
        self.ch, self.ladder, self.seed, self.rank, self.world, self.group = chains, int(ladder), int(seed), rank, world, group
        n = chains.n
        if (n * world) % ladder:
            raise ValueError("world*n_chains must be a multiple of the ladder size")
        dev = torch.device("cuda", chains.lattice.device)
        self.of_local = torch.zeros(n, dtype=torch.float64, device=dev)
        self.of_all = torch.zeros(n * world, dtype=torch.float64, device=dev)
        g_all = np.concatenate([global_chain_id(r, world, np.arange(n)) for r in range(world)])   # rank-major
        ts = ladder_tscale(g_all, ladder, span)
        self.tscale_all = torch.from_numpy(ts).to(dev)
        chains.set_tscale(ts[rank * n:(rank + 1) * n])
        self.swapped = 0
        self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def round(self, r, t_ref):
        """One swap round (parity = r mod 2).  Returns the device time of gather + decision kernels in ms."""
        torch = self.torch
        ch = self.ch
        self.ev0.record()
        B.check(ch.lib.so_objectives_dev(ch.h, C.c_void_p(self.of_local.data_ptr()), _stream_ptr()))
        torch.cuda.nvtx.range_push("replica-exchange objectives all-gather (NCCL)")
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.of_all, self.of_local, group=self.group)
        else:
            self.of_all.copy_(self.of_local)
        torch.cuda.nvtx.range_pop()
        ns = C.c_int64()
        B.check(ch.lib.so_exchange(ch.h, self.rank, self.world, self.ladder, int(r) & 1, self.seed, int(r), float(t_ref),
                                   C.c_void_p(self.of_all.data_ptr()), C.c_void_p(self.tscale_all.data_ptr()),
                                   _stream_ptr(), C.byref(ns)))
        self.ev1.record()
        self.ev1.synchronize()
        self.swapped += ns.value
        return self.ev0.elapsed_time(self.ev1)
