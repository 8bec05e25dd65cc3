"""
optimize -- drop-in for OML/optimize_SA.py:11-187: simulated annealing that MAXIMISES the objective by single-site
flips.  Reference keyword arguments keep their meaning; keyword-only extras (n_chains, seed, proposals, uniforms,
exact_guard, ...) default to the reference's behaviour of one chain.  The hot loop (:59-132) runs in so_anneal():
per-flip delta of the surrogate, Metropolis acceptance and commit for every chain, many steps per launch; the host
only cuts the run into segments at the trajectory-record points (:29-34,126-130).
"""
import os
import random
import time
import numpy as np
from . import engine
from .train_surrogate import tables_for


def cooling_schedule(cooling, T_0, total_steps, py2_linear=False):
    """Temperature of every step, OML/optimize_SA.py:61-68 (np.exp / np.log as upstream)."""
    step = np.arange(total_steps, dtype=np.float64)
    tau = total_steps / 5.0
    if cooling == 'log':
        return T_0 / np.log(step + 2)
    if cooling == 'exp':
        return T_0 * np.exp(-step / tau)
    if cooling == 'linear':
        if py2_linear:          # upstream runs under Python 2: (step+1)/total_steps is an integer division
            return T_0 * (1 - ((np.arange(total_steps) + 1) // total_steps)).astype(np.float64)
        return T_0 * (1 - (step + 1) / total_steps)
    if cooling == 'quench':
        return np.zeros(total_steps)
    raise NameError('Unrecognized cooling schedule')


def default_guard_tol(tables, p_w2_abs_sum, y_scale):
    """|delta - threshold| below which a decision is re-taken with the original fp64 weights: a generous multiple of
    the weight-quantisation error of delta (grid 2^e_h on the first layer)."""
    return 8.0 * np.sqrt(tables.K) * np.ldexp(1.0, tables.e_h) * abs(y_scale) * p_w2_abs_sum


def optimize(cat, mode='surrogate', surrogate=None, syms=None, n_cycles=100, T_0=0.05, cooling='exp', fldr=None,
             n_record=100, prefix='', *, n_chains=1, seed=None, proposals=None, uniforms=None, exact_guard=True,
             guard_tol=None, gate=True, descriptor=None, py2_linear=False, ladder=None, exchange_every=None, gather_every=None,
             return_stats=False):
    """
    :param cat:        catalyst (LSC_cat / NH3_cat); its occupancy is the start and is replaced by the result
    :param mode:       'surrogate' or 'analytical' (LSC only)
    :param surrogate:  fitted surrogate (attributes predictor / classifier / Y_scaler), surrogate mode only
    :param syms:       upstream: the translation matrix of the start structure; accepted and checked, not needed
    :param n_cycles:   total Metropolis steps = n_cycles * number of variable sites
    :param T_0, cooling: schedule, 'exp' | 'log' | 'linear' | 'quench'
    Extras: n_chains independent chains from the same start (the best final one is kept); seed of the Philox streams
    (default: drawn from Python's global `random`, the generator upstream uses); proposals/uniforms [n_chains, steps]
    to inject the random sequence; exact_guard re-decides near-ties with the fp64 weights.
    Parallel tempering: `ladder=R` groups the chains into ladders of R temperatures, geometric in
    [T*e^-5, T] around the schedule's T; every `exchange_every` steps (default: one sweep) adjacent rungs swap
    temperatures.  Under torch.distributed (one process per GPU) each rank anneals its own n_chains chains, ladders are
    strided over the ranks, and every rank ends with the same global best structure (NCCL all-gather of one record per
    rank); `gather_every` steps additionally tracks the best structure seen so far.
    """
    N = len(cat.variable_occs)
    total_steps = n_cycles * N
    step_rec = np.zeros(n_record + 1)
    OF_rec = np.zeros(n_record + 1)
    temps_rec = np.zeros(n_record + 1)
    steps_to_record = np.linspace(0, total_steps, n_record + 1).astype(int)
    if mode not in ('surrogate', 'analytical'):
        raise NameError('Unrecognized mode for evaluating objective function')
    temps = cooling_schedule(cooling, T_0, total_steps, py2_linear=py2_linear)
    if mode == 'analytical':
        if (proposals is None) != (uniforms is None):
            raise ValueError('proposals and uniforms must be given together')
        return _optimize_analytical(cat, temps, T_0, steps_to_record, step_rec, OF_rec, temps_rec, fldr, prefix,
                                    proposals, uniforms, n_chains=n_chains, seed=seed, return_stats=return_stats)
    if surrogate is None:
        raise NameError('surrogate mode needs a fitted surrogate')

    x0 = cat._occs_array()
    if syms is not None and np.asarray(syms).shape[0] != N:
        raise ValueError('syms must have one row per variable site')
    lattice = cat._lattice
    tables = surrogate if isinstance(surrogate, engine.SurrogateTables) else tables_for(
        surrogate, lattice, gate=gate, descriptor=descriptor if descriptor is not None else getattr(surrogate, 'descriptor', None))
    chains = engine.Chains(lattice, n_chains)
    chains.set_occupancy(np.repeat(x0[None, :], n_chains, axis=0))
    chains.attach(tables)
    rank, world = _rank_world()
    seed = _shared_seed(seed, rank, world)
    chain_id0 = rank * n_chains
    pt = gather = None
    if ladder is not None or gather_every is not None or world > 1:
        from . import parallel
        if ladder is not None:
            pt = parallel.ReplicaExchange(chains, ladder=int(ladder), seed=seed, rank=rank, world=world)
            exchange_every = N if exchange_every is None else int(exchange_every)
            if exchange_every < 1:
                raise ValueError('exchange_every must be a positive number of steps')
        gather = parallel.BestGather(chains, chain_id0, world)
    if (proposals is None) != (uniforms is None):
        raise ValueError('proposals and uniforms must be given together')
    if proposals is not None:
        proposals = np.ascontiguousarray(proposals, dtype=np.int32).reshape(n_chains, total_steps)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float32).reshape(n_chains, total_steps)
    if exact_guard and guard_tol is None:
        mlp = getattr(surrogate, 'predictor', None)
        w2sum = float(np.abs(mlp.coefs_[1]).sum()) if mlp is not None else float(tables.H)
        scaler = getattr(surrogate, 'Y_scaler', None)
        guard_tol = default_guard_tol(tables, w2sum, float(scaler.scale_[0]) if scaler is not None else 1.0)

    try:
        return _anneal_and_record(cat, chains, tables, temps, T_0, total_steps, steps_to_record, step_rec, OF_rec, temps_rec,
                                  proposals, uniforms, exact_guard, guard_tol, seed, chain_id0, rank, world, pt, gather,
                                  exchange_every, gather_every, fldr, prefix, return_stats, N)
    finally:
        chains.close()                                        # also when a launch raises: no device state outlives the call


def _anneal_and_record(cat, chains, tables, temps, T_0, total_steps, steps_to_record, step_rec, OF_rec, temps_rec, proposals,
                       uniforms, exact_guard, guard_tol, seed, chain_id0, rank, world, pt, gather, exchange_every,
                       gather_every, fldr, prefix, return_stats, N):
    OF = chains.objective()
    record_ind = 0
    step_rec[record_ind] = 0
    OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
    temps_rec[record_ind] = T_0
    record_ind += 1
    stats = dict(flips=0, accepted=0, guard_fired=0, kernel_ms=0.0, seed=seed)
    t_start = time.time()
    done = 0
    record_at = set(int(v) for v in steps_to_record if v > 0)
    cuts = set(record_at)
    if pt is not None:
        cuts.update(range(exchange_every, total_steps, exchange_every))
    if gather_every is not None:
        cuts.update(range(int(gather_every), total_steps, int(gather_every)))
    swap_round = 0
    for b in sorted(cuts):
        seg = slice(done, b)
        st = chains.anneal(temps[seg], step0=done, seed=seed, chain_id0=chain_id0,
                           proposals=None if proposals is None else proposals[:, seg],
                           uniforms=None if uniforms is None else uniforms[:, seg],
                           exact_guard=exact_guard, guard_tol=guard_tol or 0.0)
        for k in ('flips', 'accepted', 'guard_fired', 'kernel_ms'):
            stats[k] += st[k]
        done = b
        if pt is not None and b % exchange_every == 0 and b < total_steps:
            pt.round(swap_round, t_ref=max(float(temps[b - 1]), 1e-300))
            swap_round += 1
        if gather_every is not None and b % int(gather_every) == 0:
            gather.gather()
        if b in record_at:
            OF = chains.objective()
            step_rec[record_ind] = b                          # (step + 1) of the last step of the segment
            OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
            temps_rec[record_ind] = temps[b - 1]
            record_ind += 1
    stats['wall_s'] = time.time() - t_start
    OF = chains.objective()
    best = int(np.argmax(OF)) if len(OF) else 0
    x = chains.get_occupancy(best, 1)[0]
    stats['best_chain'], stats['OF'] = chain_id0 + best, float(OF[best])
    if gather is not None:
        # every rank ends with the same structure: the global best of the final states, or an earlier one that was better
        gather.gather()
        of_b, cid_b, bits_b = gather.best
        x = engine.unpack_bits(bits_b[None, :], N)[0]
        stats['best_chain'], stats['OF'] = int(cid_b), float(of_b)
        stats['world'], stats['rank'] = world, rank
    if pt is not None:
        stats['temperature_swaps'] = int(pt.swapped)
    stats['trajectory'] = np.array([step_rec, OF_rec])
    stats['temperatures'] = temps_rec
    cat.assign_occs([int(v) for v in x])                      # OML/optimize_SA.py:138
    _save_trajectory(fldr, prefix, step_rec, OF_rec, temps_rec)
    return stats if return_stats else None


def _shared_seed(seed, rank, world):
    """The Philox key of a run.  seed=None draws one from Python's global `random` (the generator upstream uses); under
    torch.distributed EVERY rank must hold the same key -- the parallel-tempering swap decisions are recomputed on each
    rank from it -- so rank 0's value (given or drawn) is broadcast."""
    if seed is None:
        seed = random.getrandbits(63)
    if world > 1:
        import torch.distributed as dist
        box = [int(seed) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        seed = box[0]
    return int(seed)


def _rank_world():
    """(rank, world) of torch.distributed when this process is one of several (one per GPU), else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def _save_trajectory(fldr, prefix, step_rec, OF_rec, temps_rec):
    """OML/optimize_SA.py:144-175: trajectory .npy always, the two PNGs when matplotlib is importable."""
    if fldr is None:
        return
    trajectory = np.array([step_rec, OF_rec])
    np.save(os.path.join(fldr, prefix + 'sim_anneal_trajectory.npy'), trajectory)
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except Exception:
        return
    for y, label, name in ((trajectory[1, :], 'Structure rate', 'sim_anneal_trajectory'),
                           (temps_rec, 'Temperature', 'temperature_profile')):
        plt.figure()
        plt.plot(trajectory[0, :], y, '-')
        plt.xlabel('Metropolis step', size=24)
        plt.ylabel(label, size=24)
        plt.xlim([trajectory[0, 0], trajectory[0, -1]])
        plt.ylim([0, None])
        plt.tight_layout()
        plt.savefig(os.path.join(fldr, prefix + name), dpi=600)
        plt.close()


def _optimize_analytical(cat, temps, T_0, steps_to_record, step_rec, OF_rec, temps_rec, fldr, prefix, proposals,
                         uniforms, n_chains=1, seed=None, return_stats=False):
    """mode='analytical' (LSC, OML/optimize_SA.py:43-44,103-104): every step's objective is the analytical steady
    state of the toy model, evaluated by so_anneal_analytical (typing + dense solve in shared memory per step)."""
    x0 = cat._occs_array()
    N = len(x0)
    total_steps = len(temps)
    chains = engine.Chains(cat._lattice, n_chains)
    chains.set_occupancy(np.repeat(x0[None, :], n_chains, axis=0))
    if seed is None:
        seed = random.getrandbits(63)
    if proposals is not None:
        proposals = np.ascontiguousarray(proposals, dtype=np.int32).reshape(n_chains, total_steps)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float32).reshape(n_chains, total_steps)
    OF, _ = chains.lsc_analytical(site_rates=False)
    record_ind = 0
    OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
    temps_rec[record_ind] = T_0
    record_ind += 1
    stats = dict(flips=0, accepted=0, guard_fired=0, kernel_ms=0.0, seed=seed)
    done = 0
    for b in sorted(set(int(v) for v in steps_to_record if v > 0)):
        seg = slice(done, b)
        st = chains.anneal_analytical(temps[seg], step0=done, seed=seed,
                                      proposals=None if proposals is None else proposals[:, seg],
                                      uniforms=None if uniforms is None else uniforms[:, seg])
        for k in ('flips', 'accepted', 'kernel_ms'):
            stats[k] += st[k]
        OF = st['OF']
        done = b
        step_rec[record_ind] = b
        OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
        temps_rec[record_ind] = temps[b - 1]
        record_ind += 1
    best = int(np.argmax(OF)) if len(OF) else 0
    x = chains.get_occupancy(best, 1)[0]
    stats['best_chain'], stats['OF'] = best, float(OF[best])
    stats['trajectory'] = np.array([step_rec, OF_rec])
    stats['temperatures'] = temps_rec
    cat.assign_occs([int(v) for v in x])
    chains.close()
    _save_trajectory(fldr, prefix, step_rec, OF_rec, temps_rec)
    return stats if return_stats else None
