"""The reference's online-learning loop (drivers/Online_learn_main.py:105-230) end to end on the GPU, with the toy
model's analytical site rates standing in for the external Zacros KMC run (the reference's own stand-in:
drivers/toy_cat_optimize.py evaluates LSC structures analytically):

    database of random island structures  ->  site rates  ->  surrogate.train (tree on the host, regressor on the GPU)
      ->  optimize (simulated annealing on the GPU, gated MLP)  ->  evaluate the optimum  ->  add it to the database

Usage: python scripts/active_learning_loop.py [iterations=3] [n_structures=16] [n_cycles=100]"""
import os, sys, time, random, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "structure-optimization_b200"))
import numpy as np
from OML.LSC_cat import LSC_cat                     # the reference's module paths
from OML.train_surrogate import surrogate
from OML.optimize_SA import optimize
from so_b200.database import make_random_structure


def main(iterations=3, n_strucs=16, n_cycles=100, verbose=True):
    warnings.simplefilter("ignore")
    cat = LSC_cat()
    N = cat.atoms_per_layer
    occs, sym_rows, rates = [], [], []

    def add(x):
        cat.assign_occs([int(v) for v in x])
        occs.append(list(cat.variable_occs))
        sym_rows.append(cat.generate_all_translations_and_rotations())      # 3N x N, rotations of every translation
        rates.append(cat.compute_site_rates_lsc())                          # "KMC" labels: analytical site rates
        return float(np.sum(rates[-1]))

    for i in range(n_strucs):
        add(make_random_structure(i, n_strucs, cat))
    history = []
    for it in range(iterations):
        t0 = time.perf_counter()
        surr = surrogate()
        surr.train(np.vstack(sym_rows), np.array(rates), max_iter=300, random_state=it)
        t_train = time.perf_counter() - t0
        best = int(np.argmax([np.sum(r) for r in rates]))
        cat.assign_occs(list(occs[best]))                                   # start from the best structure so far
        random.seed(it)
        t0 = time.perf_counter()
        out = optimize(cat, surrogate=surr, syms=cat.generate_all_translations(), n_cycles=n_cycles,
                       T_0=float(np.max(rates)), n_chains=64, seed=it, return_stats=True)
        t_opt = time.perf_counter() - t0
        predicted = out["OF"]
        actual = add(cat.variable_occs)
        history.append((predicted, actual, max(np.sum(r) for r in rates[:-1])))
        if verbose:
            print("iteration %d: train %.2f s (regressor %d epochs on the GPU), anneal %.2f s (%d flip-evaluations); surrogate "
                  "predicts %.3f, model gives %.3f; best in database before: %.3f"
                  % (it, t_train, surr.predictor.n_iter_, t_opt, out["flips"], predicted, actual, history[-1][2]), flush=True)
    return history


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
