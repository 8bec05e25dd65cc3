"""
TEST INFRASTRUCTURE -- runs the reference's OWN source files (read from /root/reference at run time, never copied
into this repo) under Python 3, so that golden vectors come from the reference's code and not from a restatement.

Only tests/golden/make_ref_golden.py (fixture generation, build container) and tests/test_ref_golden.py::*_live
(skipped when /root/reference is absent, e.g. on the GPU box) use this module.  The product never imports it.

The reference is Python 2 and depends on ase / zacros_wrapper / matplotlib / mpi4py, none of which exist here.
What is done to make its code execute, exhaustively:

  Source transformations (mechanical, applied to the text/AST of every loaded file; no line is hand-edited):
    T1  py2 `print <expr>` statements -> `print(<expr>)`                      (regex, line based)
    T2  name `xrange` -> `range`                                              (AST)
    T3  py2 semantics of `/`: every ast.Div becomes `__py2_div__(a, b)` = floor division when both operands
        are integers, true division otherwise (OML/dynamic_cat.py:363 `var_ind / self.dim1`,
        OML/optimize_SA.py:66 `(step + 1) / total_steps`)                      (AST)
  Shims (objects placed next to the code; the code itself is untouched):
    S1  stub modules: `ase.build.fcc111`, `ase.neighborlist.NeighborList`, `ase.io`, `zacros_wrapper.Lattice`,
        `matplotlib`, `matplotlib.pyplot`, `mpi4py`, `zacros_wrapper` -- written here from the published behaviour of
        those libraries (ASE's non-orthogonal fcc111 builder + radius neighbour list; Zacros-Wrapper's Lattice
        container with Build_neighbor_list over the self/north/northeast/east/southeast cells).  The geometry they
        produce is pinned by the reference's own fixture zacros_input/lattice_input.dat (tests/test_oracle_lattice.py).
    S2  networkx 1.x API: `Graph.node` -> `Graph.nodes` (subclass seen by the reference through its `nx` name)
    S3  `np.float` (removed in NumPy 1.24) -> `float`, through a numpy proxy in train_surrogate's namespace
    S4  train_surrogate defects: `Y_scaler` is a bare global in eval_structure_rate -> injected
        into the module namespace (with a 1-D tolerant `inverse_transform`, modern sklearn wants 2-D);
        `eval_rate` (the name optimize() calls) aliased to `eval_structure_rate`; the broken `train()` is not called,
        its three steps are (partition_data_set, train_classifier, train_regressor)
    S5  `random` in optimize_SA's / dynamic_cat's namespace can be replaced by a REPLAY object so that
        `random.choice(range(N))` returns prop[step] and `random.random()` returns u[step]
    S6  `print` -> no-op inside the loaded namespaces (optimize() prints every step)
    S7  NH3_cat.flip_atom (OML/NH3_cat.py:115-155) references undefined names and cannot run; it is never called.
"""
import ast
import builtins
import contextlib
import copy as _copy
import math
import os
import re
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "OML", "optimize_SA.py"))


# ---------------------------------------------------------------------------------------------------------
# T1-T3: source -> code object
# ---------------------------------------------------------------------------------------------------------
_PRINT_RE = re.compile(r"^(\s*)print[ \t]+(?!\()(.+?)\s*$")


def _py2_div(a, b):
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


class _Py2Semantics(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2_div__", ctx=ast.Load()),
                                              args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_Name(self, node):
        if node.id == "xrange":
            return ast.copy_location(ast.Name(id="range", ctx=node.ctx), node)
        return node


def load_tree(relpath):
    """Parse a reference file into a py3 AST with py2 semantics restored (T1-T3)."""
    path = os.path.join(REF_ROOT, relpath)
    lines = open(path).read().split("\n")
    lines = [_PRINT_RE.sub(r"\1print(\2)", ln) for ln in lines]
    tree = ast.parse("\n".join(lines), filename=path)
    tree = _Py2Semantics().visit(tree)
    ast.fix_missing_locations(tree)
    return tree, path


# ---------------------------------------------------------------------------------------------------------
# S1: stub third-party modules
# ---------------------------------------------------------------------------------------------------------
_Z = {"Pt": 78, "Au": 79, "Ni": 28, "Cu": 29}
_SYM = {v: k for k, v in _Z.items()}


class _Atom(object):
    def __init__(self, symbol, position, index):
        self.symbol, self.position, self.index = symbol, position, index


class Atoms(object):
    """The part of ase.Atoms the reference touches."""

    def __init__(self, numbers, positions, cell, pbc):
        self.numbers = np.array(numbers, dtype=int)
        self.positions = np.array(positions, dtype=float)
        self.cell = np.array(cell, dtype=float)
        self.pbc = np.array(pbc, dtype=bool)

    def __len__(self):
        return len(self.numbers)

    def __iter__(self):
        for i in range(len(self)):
            yield _Atom(_SYM[int(self.numbers[i])], self.positions[i], i)

    def __delitem__(self, sel):
        sel = np.asarray(sel)
        keep = ~sel if sel.dtype == bool else ~np.isin(np.arange(len(self)), sel)
        self.numbers, self.positions = self.numbers[keep], self.positions[keep]

    def get_positions(self):
        return self.positions.copy()

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def set_atomic_numbers(self, z):
        self.numbers = np.array(z, dtype=int)

    def get_chemical_symbols(self):
        return [_SYM[int(z)] for z in self.numbers]

    def set_chemical_symbols(self, s):
        self.numbers = np.array([_Z[v] for v in s], dtype=int)

    def set_pbc(self, pbc):
        self.pbc = np.array(pbc, dtype=bool)

    def get_cell(self):
        return self.cell.copy()

    def set_cell(self, cell):
        self.cell = np.array(cell, dtype=float)

    def get_scaled_positions(self, wrap=True):
        frac = np.linalg.solve(self.cell.T, self.positions.T).T
        if wrap:
            for i in range(3):
                if self.pbc[i]:
                    frac[:, i] %= 1.0
                    frac[:, i] %= 1.0
        return frac


def fcc111(symbol, size, a=None, vacuum=None, orthogonal=False):
    """ASE's fcc(111) slab builder, non-orthogonal cell: atoms on an integer grid [layer][i2][i1], the second layer
    from the top (and every third below) shifted by (-1/3, 2/3), the third by (1/3, 1/3) in units of the in-plane
    basis a/sqrt(2)*{(1,0),(1/2,sqrt(3)/4)}, layers a/sqrt(3) apart, then centred with `vacuum` on both sides."""
    if orthogonal:
        raise NotImplementedError
    n1, n2, n3 = size
    grid = np.empty((n3, n2, n1, 3))
    grid[..., 0] = np.arange(n1).reshape((1, 1, -1))
    grid[..., 1] = np.arange(n2).reshape((1, -1, 1))
    grid[..., 2] = np.arange(n3).reshape((-1, 1, 1))
    grid[-2::-3, ..., :2] += (-1.0 / 3, 2.0 / 3)
    grid[-3::-3, ..., :2] += (1.0 / 3, 1.0 / 3)
    grid = grid.reshape((-1, 3))
    dnn = a / math.sqrt(2.0)
    surface_cell = dnn * np.array([(1.0, 0.0), (0.5, math.sqrt(0.75))])
    pos = np.zeros_like(grid)
    pos[:, :2] = grid[:, :2] @ surface_cell
    pos[:, 2] = grid[:, 2] * dnn * math.sqrt(2.0 / 3.0)
    cell = np.zeros((3, 3))
    cell[0, :2], cell[1, :2] = surface_cell[0] * n1, surface_cell[1] * n2
    thickness = pos[:, 2].max() - pos[:, 2].min()
    if vacuum is not None:
        pos[:, 2] += vacuum - pos[:, 2].min()
        cell[2, 2] = thickness + 2 * vacuum
    else:
        cell[2, 2] = n3 * dnn * math.sqrt(2.0 / 3.0)
    return Atoms([_Z[symbol]] * len(pos), pos, cell, (True, True, False))


def fcc100(*a, **k):
    raise NotImplementedError("fcc100 is outside the tested path")


class NeighborList(object):
    """ase.neighborlist.NeighborList: atoms i, j are neighbours iff |r_j + image - r_i| < (c_i+skin) + (c_j+skin),
    images over the periodic axes; with bothways=False every pair is listed once."""

    def __init__(self, cutoffs, skin=0.3, sorted=False, self_interaction=True, bothways=False):
        self.cutoffs = np.asarray(cutoffs, dtype=float) + skin
        self.self_interaction, self.bothways = self_interaction, bothways
        self.neighbors = None

    def build(self, atoms):
        pos, cell, n = atoms.positions, atoms.cell, len(atoms)
        rng = [(-1, 0, 1) if p else (0,) for p in atoms.pbc]
        images = [np.array((i, j, k)) @ cell for i in rng[0] for j in rng[1] for k in rng[2]]
        nb = [[] for _ in range(n)]
        for i in range(n):
            for im in images:
                dist = np.sqrt(((pos + im - pos[i]) ** 2).sum(1))
                for j in np.nonzero(dist < self.cutoffs + self.cutoffs[i])[0]:
                    is_self = (j == i and not im.any())
                    if is_self and not self.self_interaction:
                        continue
                    if self.bothways or j > i or (j == i and is_self):
                        nb[i].append(int(j))
        self.neighbors = [np.array(sorted_(v), dtype=int) for v in nb]

    update = build


def sorted_(v):
    return builtins.sorted(set(v))


class Lattice(object):
    """zacros_wrapper.Lattice.Lattice: container of KMC sites; Build_neighbor_list(cut) lists site pairs closer
    than `cut` in the same cell ('self', site_1 < site_2) or with site_2 displaced into the north (+b), northeast
    (+a+b), east (+a) or southeast (+a-b) periodic image."""
    CELLS = ("self", "north", "northeast", "east", "southeast")

    def __init__(self):
        self.text_only = True
        self.lattice_matrix = None
        self.repeat = [1, 1]
        self.site_type_names = []
        self.site_type_inds = []
        self.frac_coords = []
        self.cart_coords = []
        self.neighbor_list = []
        self.cell_list = []

    def set_cart_coords(self, cart_coords):
        self.cart_coords = np.array(cart_coords, dtype=float).reshape(-1, 2)
        self.frac_coords = self.cart_coords @ np.linalg.inv(np.asarray(self.lattice_matrix, dtype=float))

    def set_frac_coords(self, frac_coords):
        self.frac_coords = np.array(frac_coords, dtype=float)
        self.cart_coords = self.frac_coords @ np.asarray(self.lattice_matrix, dtype=float)

    def Build_neighbor_list(self, cut=3.0):
        a, b = np.asarray(self.lattice_matrix, dtype=float)
        shift = {"self": 0 * a, "north": b, "northeast": a + b, "east": a, "southeast": a - b}
        self.neighbor_list, self.cell_list = [], []
        n = len(self.site_type_inds)
        for s1 in range(n):
            for s2 in range(n):
                for cell in self.CELLS:
                    if cell == "self" and s1 >= s2:
                        continue
                    if np.linalg.norm(self.cart_coords[s2] + shift[cell] - self.cart_coords[s1]) < cut:
                        self.neighbor_list.append([s1, s2])
                        self.cell_list.append(cell)


class _Anything(object):
    """matplotlib / mpi4py stand-in: every attribute is a callable that returns another stand-in."""
    rcParams = {}

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _stub_modules():
    anything = _Anything()
    ase = _module("ase")
    ase.build = _module("ase.build", fcc111=fcc111, fcc100=fcc100)
    ase.io = _module("ase.io", read=anything, write=anything)
    ase.neighborlist = _module("ase.neighborlist", NeighborList=NeighborList)
    zw = _module("zacros_wrapper")
    zw.Lattice = _module("zacros_wrapper.Lattice", Lattice=Lattice)
    mat = _module("matplotlib", rcParams={})
    mat.pyplot = _module("matplotlib.pyplot")
    mat.pyplot.__getattr__ = lambda name: anything          # PEP 562
    return {"ase": ase, "ase.build": ase.build, "ase.io": ase.io, "ase.neighborlist": ase.neighborlist,
            "zacros_wrapper": zw, "zacros_wrapper.Lattice": zw.Lattice, "matplotlib": mat,
            "matplotlib.pyplot": mat.pyplot, "mpi4py": _module("mpi4py", MPI=anything)}


@contextlib.contextmanager
def _patched_sys_modules(extra):
    saved = {k: sys.modules.get(k) for k in extra}
    sys.modules.update(extra)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---------------------------------------------------------------------------------------------------------
# S2, S3: networkx / numpy proxies
# ---------------------------------------------------------------------------------------------------------
def _nx_proxy():
    import networkx as nx

    class Graph(nx.Graph):
        @property
        def node(self):                      # networkx 1.x spelling used throughout the reference
            return self.nodes

    class _NX(object):
        def __getattr__(self, name):
            return Graph if name == "Graph" else getattr(nx, name)
    return _NX()


class _NumpyProxy(object):
    float = float

    def __getattr__(self, name):
        return getattr(np, name)


# ---------------------------------------------------------------------------------------------------------
# the loaded reference
# ---------------------------------------------------------------------------------------------------------
class ReplayRandom(object):
    """S5: stands in for the `random` module inside optimize(): choice() returns the injected proposal of the
    current step, random() the injected uniform of the same step (consumed or not, as upstream decides)."""

    def __init__(self, proposals, uniforms):
        self.prop = [int(v) for v in proposals]
        self.u = [float(v) for v in uniforms]
        self.step = -1
        self.u_used = []

    def choice(self, seq):
        self.step += 1
        return seq[self.prop[self.step]]

    def random(self):
        self.u_used.append(self.step)
        return self.u[self.step]


class Reference(object):
    """The reference's modules, executed from source.  Attributes: dynamic_cat, LSC_cat, NH3_cat, surrogate (classes),
    optimize, dyn_cat_dist, make_random_structure (functions), and the module namespaces in .ns[...]"""

    def __init__(self, quiet=True):
        if not available():
            raise RuntimeError("reference sources not found under %s" % REF_ROOT)
        self.ns = {}
        stubs = _stub_modules()
        oml = _module("OML")
        oml.__path__ = []
        stubs["OML"] = oml
        with _patched_sys_modules(stubs):
            for name in ("dynamic_cat", "LSC_cat", "NH3_cat", "train_surrogate", "optimize_SA"):
                tree, path = load_tree(os.path.join("OML", name + ".py"))
                mod = _module("OML." + name, __file__=path, __py2_div__=_py2_div)
                if quiet:
                    mod.__dict__["print"] = lambda *a, **k: None                 # S6
                sys.modules["OML." + name] = mod
                setattr(oml, name, mod)
                exec(compile(tree, path, "exec"), mod.__dict__)
                self.ns[name] = mod.__dict__
            for name in ("dynamic_cat", "LSC_cat", "NH3_cat", "train_surrogate", "optimize_SA"):
                sys.modules.pop("OML." + name, None)
        nxp = _nx_proxy()
        for name in ("LSC_cat", "NH3_cat"):
            self.ns[name]["nx"] = nxp                                            # S2
        self.ns["train_surrogate"]["np"] = _NumpyProxy()                         # S3
        self.ns["train_surrogate"]["os"] = os                                    # S4 (train() uses os un-imported)
        self.dynamic_cat = self.ns["dynamic_cat"]["dynamic_cat"]
        self.dyn_cat_dist = self.ns["dynamic_cat"]["dyn_cat_dist"]
        self.LSC_cat = self.ns["LSC_cat"]["LSC_cat"]
        self.NH3_cat = self.ns["NH3_cat"]["NH3_cat"]
        self.surrogate = self.ns["train_surrogate"]["surrogate"]
        self.surrogate.eval_rate = self.surrogate.eval_structure_rate            # S4
        self.optimize = self.ns["optimize_SA"]["optimize"]
        # drivers/generate_init_structures.py: only the function make_random_structure is taken (the module body
        # imports mpi4py and cluster paths and runs main()).
        tree, path = load_tree(os.path.join("drivers", "generate_init_structures.py"))
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "make_random_structure"]
        gen_ns = {"np": np, "__py2_div__": _py2_div, "write_structure_files": lambda cat, fldr: None}
        exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), gen_ns)
        self.ns["generate_init_structures"] = gen_ns
        self.make_random_structure = gen_ns["make_random_structure"]

    # -- helpers around the reference objects (observation only) -------------------------------------------
    def set_random(self, module, obj):
        """S5: replace the `random` name seen by OML/<module>.py; returns the previous object."""
        old = self.ns[module]["random"]
        self.ns[module]["random"] = obj
        return old

    def set_y_scaler(self, scaler):
        """S4: eval_structure_rate reads a module-global Y_scaler."""
        class _Scaler1D(object):
            def inverse_transform(self_inner, y):
                return scaler.inverse_transform(np.asarray(y).reshape(-1, 1))
        self.ns["train_surrogate"]["Y_scaler"] = _Scaler1D()

    def node_types(self, cat):
        """site_type of every graph node after graph_to_KMClattice(), None -> 0."""
        n_at = len(cat.atoms_template)
        return np.array([cat.defected_graph.nodes[i]["site_type"] or 0 for i in range(n_at)], dtype=np.uint8)

    def train_surrogate(self, all_syms, site_rates, seed=0):
        """partition_data_set + train_classifier + train_regressor as train() intends (train_surrogate.py:61-69).
        MLPRegressor has no random_state upstream: the global NumPy stream is seeded here instead."""
        ns = self.ns["train_surrogate"]
        sur = self.surrogate()
        sur.all_syms = all_syms
        sur.partition_data_set(site_rates)
        sur.train_classifier()
        captured = {}
        real_scaler = ns["StandardScaler"]

        class CapturingScaler(real_scaler):              # Y_scaler is a LOCAL of train_regressor: observe it
            def fit(self_inner, *a, **k):
                captured["scaler"] = self_inner
                return real_scaler.fit(self_inner, *a, **k)

            def inverse_transform(self_inner, y, *a, **k):
                return real_scaler.inverse_transform(self_inner, np.asarray(y).reshape(-1, 1)).reshape(-1)
        ns["StandardScaler"] = CapturingScaler
        real_mlp = ns["MLPRegressor"]
        ns["MLPRegressor"] = lambda **kw: real_mlp(**dict(kw, verbose=False))
        np.random.seed(seed)
        try:
            sur.train_regressor()
        finally:
            ns["StandardScaler"], ns["MLPRegressor"] = real_scaler, real_mlp
        sur.Y_scaler = captured["scaler"]
        self.set_y_scaler(sur.Y_scaler)
        return sur

    def run_optimize(self, cat, x0, proposals, uniforms, surrogate=None, mode="surrogate", **kw):
        """optimize() with injected draws.  Returns dict(x, accept[steps], OF_new[steps], OF0, u_used, traj)."""
        cat.variable_occs = [int(v) for v in x0]
        seen = []
        replay = ReplayRandom(proposals, uniforms)
        old = self.set_random("optimize_SA", replay)
        if mode == "surrogate":
            inner = surrogate.eval_rate

            class Observed(object):
                def eval_rate(self_inner, syms):
                    v = inner(syms)
                    seen.append((np.asarray(syms[0]).astype(np.uint8).copy(), float(v)))
                    return v
            target, kwargs = Observed(), dict(surrogate=None)
            kwargs["surrogate"] = target
        else:
            inner = cat.eval_struc_rate_anal

            def observed(x):
                v = inner(x)
                seen.append((np.asarray(x, dtype=np.uint8).copy(), float(v)))
                return v
            cat.eval_struc_rate_anal = observed
            kwargs = {}
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            try:
                self.optimize(cat, mode=mode, fldr=tmp, **kwargs, **kw)
            finally:
                self.set_random("optimize_SA", old)
                if mode != "surrogate":
                    del cat.eval_struc_rate_anal
            traj = np.load(os.path.join(tmp, kw.get("prefix", "") + "sim_anneal_trajectory.npy"))
        x_final = np.array(cat.variable_occs, dtype=np.uint8)
        steps = len(seen) - 1
        # row 0 of the evaluated matrix is the candidate x_new itself (translation by (0,0)); a step was accepted
        # iff the next candidate differs from it in exactly the next proposed site
        cand = [s[0] for s in seen[1:]]
        accept = np.zeros(steps, dtype=bool)
        for s in range(steps):
            if s + 1 < steps:
                nxt = cand[s + 1].copy()
                nxt[replay.prop[s + 1]] ^= 1
            else:
                nxt = x_final
            accept[s] = bool((nxt == cand[s]).all())
        return dict(x=x_final, accept=accept, OF_new=np.array([s[1] for s in seen[1:]]), OF0=seen[0][1],
                    u_used=np.array(replay.u_used, dtype=np.int64), traj=traj)


_REF = None


def reference():
    global _REF
    if _REF is None:
        _REF = Reference()
    return _REF
