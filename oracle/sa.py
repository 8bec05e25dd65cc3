"""
ORACLE (test infrastructure, NOT product code) -- the simulated-annealing loop.

Reference followed:
  OML/optimize_SA.py:11-138   optimize(): schedules (:61-68), proposal (:75-83), incremental descriptor
                              update (:86-92), objective (:102), Metropolis (:108-123), record (:126-130)
Defects/ambiguities resolved as in SURVEY 9.5: proposals and uniforms are indexed by step (prop[s], u[s]; u[s]
is simply unused when the move improves or T == 0), `print step` dropped, py2 integer division made explicit,
tie (delta == 0) goes to the uphill branch.

  optimize_faithful()  literal restatement: O(N*K) copy of the descriptor matrix, per-row update, full
                       surrogate.eval_rate() per step, fp64.  This is the "reference CPU path" that bench.py times.
  anneal_fixed()       same loop in the CUDA path's exact fixed-point arithmetic (see oracle/surrogate.py);
                       its decisions/occupancies/sums must be reproduced bit-for-bit by the GPU.
"""
import math
import numpy as np
from . import lattice as L
from . import surrogate as S


def schedule(cooling, T_0, total_steps, py2_linear=False):
    """Temperature at every step, optimize_SA.py:61-68 (np.exp / np.log as upstream)."""
    step = np.arange(total_steps, dtype=np.float64)
    tau = total_steps / 5.0
    if cooling == "log":
        return T_0 / np.log(step + 2)
    if cooling == "exp":
        return T_0 * np.exp(-step / tau)
    if cooling == "linear":
        if py2_linear:       # py2: (step+1)/total_steps is integer division -> 0 until the last step
            return T_0 * (1 - ((np.arange(total_steps) + 1) // total_steps)).astype(np.float64)
        return T_0 * (1 - (step + 1) / total_steps)
    if cooling == "quench":
        return np.zeros(total_steps)
    if cooling == "const":   # not in the reference: fixed-temperature rungs of parallel tempering
        return np.full(total_steps, float(T_0))
    raise NameError("Unrecognized cooling schedule")


def steps_to_record(total_steps, n_record):
    return np.linspace(0, total_steps, n_record + 1).astype(int)        # optimize_SA.py:33-34


def optimize_faithful(x0, d, surrogate, offsets, temps, prop, u, n_record=100, syms=None):
    """Literal optimize() (surrogate mode) for ONE chain.  `surrogate` has eval_rate(matrix).
    Returns dict(x, OF, accept[steps] bool, delta[steps], traj (step, OF/N, T))."""
    N = d * d
    total_steps = len(temps)
    rec = steps_to_record(total_steps, n_record)
    x = [int(v) for v in x0]
    if syms is None:
        syms = L.descriptor_rows(np.asarray(x), offsets, d)
    K = syms.shape[1]
    OF = surrogate.eval_rate(syms)
    traj = [(0, OF / N, temps[0] if total_steps else 0.0)]
    accept_log = np.zeros(total_steps, dtype=bool)
    deltas = np.zeros(total_steps)
    rec_set = set(int(v) for v in rec)
    o1, o2 = offsets[:, 0], offsets[:, 1]
    for step in range(total_steps):
        T = temps[step]
        x_new = [i for i in x]
        ind = int(prop[step])
        if x_new[ind] == 1:
            x_new[ind] = 0
        elif x_new[ind] == 0:
            x_new[ind] = 1
        else:
            raise NameError("Invalid occupancy")
        syms_new = np.copy(syms)
        f1, f2 = ind % d, ind // d
        for k in range(K):                      # row r sees the flipped site in column k iff r + off_k = f
            r = ((f2 - o2[k]) % d) * d + (f1 - o1[k]) % d
            syms_new[r, k] = x_new[ind]
        OF_new = surrogate.eval_rate(syms_new)
        delta = OF_new - OF
        if delta > 0:
            accept = True
        elif T > 0:
            accept = np.exp((OF_new - OF) / T) > float(u[step])
        else:
            accept = False
        if accept:
            x, syms, OF = x_new, syms_new, OF_new
        accept_log[step] = accept
        deltas[step] = delta
        if (step + 1) in rec_set:
            traj.append((step + 1, OF / N, T))
    return dict(x=np.asarray(x, dtype=np.uint8), OF=float(OF), accept=accept_log, delta=deltas,
                traj=np.asarray(traj))


def anneal_fixed(p, q, d, x0, temps, prop, u, tscale=1.0, margins=False):
    """One chain in exact fixed-point arithmetic.  Returns dict(x, OF, hi, lo, n_act, accept, delta)."""
    N = d * d
    x = np.asarray(x0, dtype=np.int64).copy()
    hq, _ = S.preact_fixed(p, q, x, d)
    hq = hq.astype(np.int64)
    _, hi, lo, n_act, _ = S.score_fixed(p, q, x, d)
    w2q = q["w2q"].astype(np.int64)
    W1q = q["W1q"].astype(np.int64)
    o1, o2 = p.offsets[:, 0].astype(np.int64), p.offsets[:, 1].astype(np.int64)
    mask = (1 << S.SPLIT_BITS) - 1
    total = len(temps)
    accept_log = np.zeros(total, dtype=bool)
    deltas = np.zeros(total)
    margin = np.full(total, np.inf)
    gated = p.tree is not None
    r_all = np.arange(N)
    for s in range(total):
        T = np.float64(tscale) * np.float64(temps[s])
        f = int(prop[s])
        f1, f2 = f % d, f // d
        sgn = 1 if x[f] == 0 else -1
        rows = ((f2 - o2) % d) * d + (f1 - o1) % d                 # [K]
        h_old = hq[rows]
        h_new = h_old + sgn * W1q
        if not gated:
            dA = int(((np.maximum(h_new, 0) - np.maximum(h_old, 0)) * w2q).sum())
            dn = 0
            dhi, dlo = dA >> S.SPLIT_BITS, dA & mask
        else:
            A_old = (np.maximum(h_old, 0) * w2q).sum(1)
            A_new = (np.maximum(h_new, 0) * w2q).sum(1)
            x_new = x.copy()
            x_new[f] ^= 1
            rr1, rr2 = rows % d, rows // d
            idx = ((rr2[:, None] + o2[None, :]) % d) * d + (rr1[:, None] + o1[None, :]) % d
            g_old = S.tree_predict(p.tree, x[idx])
            g_new = S.tree_predict(p.tree, x_new[idx])
            dhi = sum(int(a) >> S.SPLIT_BITS for a in A_new[g_new]) - sum(int(a) >> S.SPLIT_BITS for a in A_old[g_old])
            dlo = sum(int(a) & mask for a in A_new[g_new]) - sum(int(a) & mask for a in A_old[g_old])
            dn = int(g_new.sum()) - int(g_old.sum())
        delta = float(S.real_from_sums(p, q, dhi, dlo, dn))
        if delta > 0:
            acc = True
        elif T > 0:
            pr = math.exp(delta / T)
            acc = pr > float(u[s])
            margin[s] = abs(pr - float(u[s]))
        else:
            acc = False
        if acc:
            x[f] ^= 1
            hq[rows] = h_new
            hi += dhi
            lo += dlo
            hi, lo = hi + (lo >> S.SPLIT_BITS), lo & mask
            n_act += dn
        accept_log[s] = acc
        deltas[s] = delta
    out = dict(x=x.astype(np.uint8), OF=float(S.real_from_sums(p, q, hi, lo, n_act)), hi=hi, lo=lo,
               n_act=n_act, accept=accept_log, delta=deltas, hq=hq.astype(np.int32))
    if margins:
        out["margin"] = margin
    return out


def optimize_analytical(x0, d, temps, prop, u):
    """optimize(mode='analytical') for ONE chain (OML/optimize_SA.py:43-44,103-116 with OML/LSC_cat.py:248-316)."""
    from . import lsc_analytical as LA
    x = np.asarray(x0, dtype=np.uint8).copy()
    OF = LA.eval_struc_rate_anal(x, d)
    accept_log = np.zeros(len(temps), dtype=bool)
    margin = np.full(len(temps), np.inf)
    for s in range(len(temps)):
        f = int(prop[s])
        x[f] ^= 1
        OF_new = LA.eval_struc_rate_anal(x, d)
        delta = OF_new - OF
        if delta > 0:
            acc = True
        elif temps[s] > 0:
            pr = math.exp(delta / temps[s])
            acc = pr > float(u[s])
            margin[s] = abs(pr - float(u[s]))
        else:
            acc = False
        if acc:
            OF = OF_new
        else:
            x[f] ^= 1
        accept_log[s] = acc
    return dict(x=x, OF=float(OF), accept=accept_log, margin=margin)
