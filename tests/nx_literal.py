"""
Test helper: evaluate the reference's site-typing *patterns* literally with NetworkX VF2 on the full 5N-node
slab graph, so the oracle's closed forms (oracle/site_types.py) can be cross-checked against the graph
semantics of OML/LSC_cat.py:164-245 and OML/NH3_cat.py:158-426.  The rules are held as a data table and run
by one generic matcher loop; the graph comes from distances (oracle.lattice.neighbour_sets_bruteforce), i.e.
what ASE's NeighborList produces upstream.
"""
import networkx as nx
import networkx.algorithms.isomorphism as iso
import numpy as np
from oracle import lattice as L

TRI = [("A", "B"), ("B", "C"), ("C", "A")]
TET = TRI + [("A", "D"), ("B", "D"), ("C", "D")]


def build_graph(x, d):
    N = d * d
    G = nx.Graph()
    G.add_nodes_from(range(5 * N))
    for i in range(5 * N):
        layer = i // N
        G.nodes[i]["element"] = "vacuum" if layer == 4 else "Pt"
        G.nodes[i]["site_type"] = None
    for v in range(N):
        G.nodes[3 * N + v]["element"] = "Ni" if x[v] else "vacancy"
    for i, s in enumerate(L.neighbour_sets_bruteforce(d)):
        for j in s:
            G.add_edge(i, j)
    return G


def _n_ni(G, i):
    return sum(1 for j in G.neighbors(i) if G.nodes[j]["element"] == "Ni")


def _match(G, edges, attr, labels, default):
    P = nx.Graph()
    P.add_edges_from(edges)
    for k, v in labels.items():
        P.nodes[k][attr] = v
    GM = iso.GraphMatcher(G, P, node_match=iso.categorical_node_match(attr, default))
    for m in GM.subgraph_isomorphisms_iter():
        yield {v: k for k, v in m.items()}


def lsc_types_nx(x, d):
    G = build_graph(x, d)
    for i in G.nodes:
        if G.nodes[i]["element"] == "Ni" and _n_ni(G, i) == 6:
            G.nodes[i]["site_type"] = 1
    for m in _match(G, TRI, "element", {"A": "Ni", "B": "vacancy", "C": "vacancy"}, "Ni"):
        if _n_ni(G, m["A"]) == 4:
            G.nodes[m["A"]]["site_type"] = 2
    return np.array([G.nodes[i]["site_type"] or 0 for i in range(5 * d * d)], dtype=np.uint8)


# (pattern attr, {pattern node: label}, {pattern node: type to assign})
NH3_ELEMENT_RULES_A = [
    ({"A": "Ni", "B": "Ni", "C": "Ni", "D": "vacuum"}, {"D": 1}),
    ({"A": "Ni", "B": "Ni", "C": "Ni", "D": "Pt"}, {"D": 2}),
]
NH3_ELEMENT_RULES_B = [
    ({"A": "vacancy", "B": "vacancy", "C": "vacancy", "D": "Pt"}, {"D": 8}),
]
NH3_ELEMENT_RULES_C = [
    ({"A": "Ni", "B": "vacancy", "C": "vacancy", "D": "vacuum"}, {"B": 10, "C": 10, "D": 12}),
    ({"A": "Ni", "B": "Ni", "C": "vacancy", "D": "vacuum"}, {"C": 11, "D": 14}),
    ({"A": "Ni", "B": "Ni", "C": "vacancy", "D": "Pt"}, {"D": 15}),
]


def nh3_types_nx(x, d):
    G = build_graph(x, d)
    N = d * d
    pos = L.slab_positions(d)
    T = lambda i: G.nodes[i]["site_type"]

    def put(i, t):
        G.nodes[i]["site_type"] = t

    def run_element_rules(rules):
        for labels, assign in rules:
            for m in _match(G, TET, "element", labels, "Ni"):
                for k, t in assign.items():
                    put(m[k], t)

    for i in G.nodes:
        if G.nodes[i]["element"] == "Ni":
            put(i, 4)
    for i in G.nodes:
        if G.nodes[i]["element"] == "Ni" and _n_ni(G, i) == 6:
            put(i, 3)
    for m in _match(G, TRI, "element", {"A": "Ni", "B": "vacancy", "C": "vacancy"}, "Ni"):
        if _n_ni(G, m["A"]) == 4:
            put(m["A"], 5)
    run_element_rules(NH3_ELEMENT_RULES_A)
    for i in G.nodes:
        if G.nodes[i]["element"] == "vacancy":
            put(i, 6)
    run_element_rules(NH3_ELEMENT_RULES_B)
    for m in _match(G, TET, "site_type", {"A": 8, "B": 8, "C": 8, "D": None}, 8):
        put(m["D"], 7)
    run_element_rules(NH3_ELEMENT_RULES_C)
    for i in list(G.nodes):
        if T(i) == 1:
            n_edges = sum(1 for j in G.neighbors(i) if T(j) in (4, 5))
            if n_edges == 1:
                put(i, 17)
            elif n_edges >= 2:
                put(i, 16)
    for m in _match(G, TET, "site_type", {"A": 3, "B": 3, "C": 5, "D": 2}, 8):
        put(m["D"], 18)
    for m in _match(G, TET, "site_type", {"A": 3, "B": 5, "C": 5, "D": 2}, 8):
        put(m["D"], 19)
    s1_nodes = [j for j in G.nodes if T(j) == 15]
    for i in list(G.nodes):
        if T(i) == 6:
            if any(np.linalg.norm(pos[i, :2] - pos[j, :2]) < 2.77 * 2 for j in s1_nodes):
                put(i, 9)
    for m in _match(G, TET, "site_type", {"A": 8, "B": 8, "C": None, "D": None}, 8):
        if pos[m["D"], 2] - pos[m["A"], 2] < -0.1:
            put(m["D"], 13)
    return np.array([T(i) or 0 for i in range(5 * N)], dtype=np.uint8)
