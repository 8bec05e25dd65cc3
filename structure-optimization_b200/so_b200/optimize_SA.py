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
             guard_tol=None, gate=True, py2_linear=False, return_stats=False):
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
    tables = surrogate if isinstance(surrogate, engine.SurrogateTables) else tables_for(surrogate, lattice, gate=gate)
    chains = engine.Chains(lattice, n_chains)
    chains.set_occupancy(np.repeat(x0[None, :], n_chains, axis=0))
    chains.attach(tables)
    if seed is None:
        seed = random.getrandbits(63)
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

    OF = chains.objective()
    record_ind = 0
    step_rec[record_ind] = 0
    OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
    temps_rec[record_ind] = T_0
    record_ind += 1
    stats = dict(flips=0, accepted=0, guard_fired=0, kernel_ms=0.0, seed=seed)
    t_start = time.time()
    done = 0
    for b in sorted(set(int(v) for v in steps_to_record if v > 0)):
        seg = slice(done, b)
        st = chains.anneal(temps[seg], step0=done, seed=seed,
                           proposals=None if proposals is None else proposals[:, seg],
                           uniforms=None if uniforms is None else uniforms[:, seg],
                           exact_guard=exact_guard, guard_tol=guard_tol or 0.0)
        for k in ('flips', 'accepted', 'guard_fired', 'kernel_ms'):
            stats[k] += st[k]
        done = b
        OF = chains.objective()
        step_rec[record_ind] = b                              # (step + 1) of the last step of the segment
        OF_rec[record_ind] = OF.max() / cat.atoms_per_layer
        temps_rec[record_ind] = temps[b - 1]
        record_ind += 1
    stats['wall_s'] = time.time() - t_start
    best = int(np.argmax(OF)) if len(OF) else 0
    x = chains.get_occupancy(best, 1)[0]
    stats['best_chain'], stats['OF'] = best, float(OF[best])
    stats['trajectory'] = np.array([step_rec, OF_rec])
    stats['temperatures'] = temps_rec
    cat.assign_occs([int(v) for v in x])                      # OML/optimize_SA.py:138
    chains.close()
    _save_trajectory(fldr, prefix, step_rec, OF_rec, temps_rec)
    return stats if return_stats else None


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
