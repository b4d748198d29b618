"""ESCAPE34/opf.jl:36-286 — two-stage stochastic AC optimal power flow (BASELINE.json configs[3]).

The reference downloads its grid from pglib-opf at run time (opf.jl:15-18), which is impossible
offline: ``case3()`` embeds a 3-bus / 3-branch / 3-generator case modelled on pglib_opf_case3_lmbd
(values restated from memory of that file — the model STRUCTURE is what matters here), and
``synthetic_grid(nbus)`` builds ring-plus-chords grids for the many-generator regime.
PowerModels' preprocessing (per-unit scaling, standardize_cost_terms!, calc_thermal_limits!,
calc_branch_y / calc_branch_t, build_ref) is restated in ``build_ref``.
"""
from __future__ import annotations

import numpy as np

from . import infopt as io
from .infopt import InfiniteModel, cos, sin


def case3():
    base = 100.0
    bus = {1: dict(pd=110.0, qd=40.0, vmin=0.9, vmax=1.1, ref=True),
           2: dict(pd=110.0, qd=40.0, vmin=0.9, vmax=1.1, ref=False),
           3: dict(pd=95.0, qd=50.0, vmin=0.9, vmax=1.1, ref=False)}
    gen = {1: dict(bus=1, pmin=0.0, pmax=2000.0, qmin=-1000.0, qmax=1000.0, cost=(0.11, 5.0, 0.0)),
           2: dict(bus=2, pmin=0.0, pmax=2000.0, qmin=-1000.0, qmax=1000.0, cost=(0.085, 1.2, 0.0)),
           3: dict(bus=3, pmin=0.0, pmax=0.0, qmin=-1000.0, qmax=1000.0, cost=(0.0, 0.0, 0.0))}
    branch = {1: dict(f=1, t=3, r=0.065, x=0.62, b=0.45, rate_a=9000.0),
              2: dict(f=3, t=2, r=0.025, x=0.75, b=0.7, rate_a=50.0),
              3: dict(f=1, t=2, r=0.042, x=0.9, b=0.3, rate_a=9000.0)}
    return dict(baseMVA=base, bus=bus, gen=gen, branch=branch)


def synthetic_grid(nbus: int, seed: int = 0):
    """ring + random chords, one generator at every third bus"""
    rng = np.random.default_rng(seed)
    bus = {i: dict(pd=float(rng.uniform(20, 120)), qd=float(rng.uniform(5, 40)), vmin=0.9, vmax=1.1, ref=(i == 1))
           for i in range(1, nbus + 1)}
    gen = {}
    for g, b in enumerate(range(1, nbus + 1, 3), start=1):
        gen[g] = dict(bus=b, pmin=0.0, pmax=float(rng.uniform(300, 900)), qmin=-400.0, qmax=400.0,
                      cost=(float(rng.uniform(0.01, 0.2)), float(rng.uniform(1, 10)), 0.0))
    branch = {}
    edges = [(i, i % nbus + 1) for i in range(1, nbus + 1)]
    edges += [(int(a), int(b)) for a, b in rng.integers(1, nbus + 1, size=(nbus // 3, 2)) if a != b]
    for l, (f, t) in enumerate(edges, start=1):
        branch[l] = dict(f=f, t=t, r=float(rng.uniform(0.01, 0.08)), x=float(rng.uniform(0.2, 0.9)),
                         b=float(rng.uniform(0.1, 0.6)), rate_a=900.0)
    return dict(baseMVA=100.0, bus=bus, gen=gen, branch=branch)


def build_ref(case):
    """per-unit ref like ``PowerModels.build_ref`` (opf.jl:29-34) + the branch tuples of opf.jl:52-78"""
    base = case["baseMVA"]
    bus = {i: dict(b, pd=b["pd"] / base, qd=b["qd"] / base) for i, b in case["bus"].items()}
    gen = {i: dict(g, pmin=g["pmin"] / base, pmax=g["pmax"] / base, qmin=g["qmin"] / base, qmax=g["qmax"] / base,
                   cost=(g["cost"][0] * base ** 2, g["cost"][1] * base, g["cost"][2])) for i, g in case["gen"].items()}
    branches = []
    for l, br in case["branch"].items():
        z2 = br["r"] ** 2 + br["x"] ** 2
        g_, b_ = br["r"] / z2, -br["x"] / z2
        tr, ti = 1.0, 0.0
        branches.append(dict(l=l, f_bus=br["f"], t_bus=br["t"], f_idx=(l, br["f"], br["t"]), t_idx=(l, br["t"], br["f"]),
                             g=g_, b=b_, tr=tr, ti=ti, ttm=tr ** 2 + ti ** 2, g_fr=0.0, b_fr=br["b"] / 2, g_to=0.0,
                             b_to=br["b"] / 2, angmin=-np.pi / 6, angmax=np.pi / 6, rate_a=br["rate_a"] / base))
    arcs = [b["f_idx"] for b in branches] + [b["t_idx"] for b in branches]
    bus_arcs = {i: [a for a in arcs if a[1] == i] for i in bus}
    bus_gens = {i: [g for g, d in gen.items() if d["bus"] == i] for i in bus}
    rate = {b["l"]: b["rate_a"] for b in branches}
    return dict(bus=bus, gen=gen, branch=branches, arcs=arcs, bus_arcs=bus_arcs, bus_gens=bus_gens, rate=rate,
                ref_buses=[i for i, b in bus.items() if b["ref"]])


def opf(case=None, num_supports: int = 100, seed: int = 0) -> InfiniteModel:
    ref = build_ref(case or case3())
    nbus = len(ref["bus"])
    buses = sorted(ref["bus"])
    bpos = {b: i for i, b in enumerate(buses)}
    rng = np.random.default_rng(seed)
    sd = 0.1 * np.array([ref["bus"][i]["pd"] for i in buses] + [ref["bus"][i]["qd"] for i in buses])
    supports = rng.normal(0.0, 1.0, size=(2 * nbus, num_supports)) * sd[:, None]      # MvNormal(0, diag(sd²)) (opf.jl:48-50,112)

    m = InfiniteModel()
    B, G, A = ref["bus"], ref["gen"], ref["arcs"]
    # first stage (finite) variables — opf.jl:84-110
    va0 = {i: m.variable() for i in buses}
    vm0 = {i: m.variable(lb=B[i]["vmin"], ub=B[i]["vmax"], start=1.0) for i in buses}
    pg0 = {i: m.variable(lb=G[i]["pmin"], ub=G[i]["pmax"]) for i in G}
    qg0 = {i: m.variable(lb=G[i]["qmin"], ub=G[i]["qmax"]) for i in G}
    p0 = {a: m.variable(lb=-ref["rate"][a[0]], ub=ref["rate"][a[0]]) for a in A}
    q0 = {a: m.variable(lb=-ref["rate"][a[0]], ub=ref["rate"][a[0]]) for a in A}
    # second stage — opf.jl:112-140
    th = m.dependent_parameters(supports)
    va = {i: m.variable(*th) for i in buses}
    vm = {i: m.variable(*th, lb=B[i]["vmin"], ub=B[i]["vmax"], start=1.0) for i in buses}
    pg = {i: m.variable(*th, lb=G[i]["pmin"], ub=G[i]["pmax"]) for i in G}
    qg = {i: m.variable(*th, lb=G[i]["qmin"], ub=G[i]["qmax"]) for i in G}
    p = {a: m.variable(*th, lb=-ref["rate"][a[0]], ub=ref["rate"][a[0]]) for a in A}
    q = {a: m.variable(*th, lb=-ref["rate"][a[0]], ub=ref["rate"][a[0]]) for a in A}

    m.objective("Min", sum(g["cost"][0] * pg0[i] ** 2 + g["cost"][1] * pg0[i] + g["cost"][2] for i, g in G.items()))

    def stage(va, vm, pg, qg, p, q, theta=None):
        for i in ref["ref_buses"]:
            m.constraint(va[i], "==", 0)
        for br in ref["branch"]:
            f, t = br["f_bus"], br["t_bus"]
            m.constraint(p[br["f_idx"]], "==",
                         (br["g"] + br["g_fr"]) / br["ttm"] * vm[f] ** 2
                         + (-br["g"] * br["tr"] + br["b"] * br["ti"]) / br["ttm"] * (vm[f] * vm[t] * cos(va[f] - va[t]))
                         + (-br["b"] * br["tr"] - br["g"] * br["ti"]) / br["ttm"] * (vm[f] * vm[t] * sin(va[f] - va[t])))
        for br in ref["branch"]:
            f, t = br["f_bus"], br["t_bus"]
            m.constraint(q[br["f_idx"]], "==",
                         -(br["b"] + br["b_fr"]) / br["ttm"] * vm[f] ** 2
                         - (-br["b"] * br["tr"] - br["g"] * br["ti"]) / br["ttm"] * (vm[f] * vm[t] * cos(va[f] - va[t]))
                         + (-br["g"] * br["tr"] + br["b"] * br["ti"]) / br["ttm"] * (vm[f] * vm[t] * sin(va[f] - va[t])))
        for br in ref["branch"]:
            f, t = br["f_bus"], br["t_bus"]
            m.constraint(p[br["t_idx"]], "==",
                         (br["g"] + br["g_to"]) * vm[t] ** 2
                         + (-br["g"] * br["tr"] - br["b"] * br["ti"]) / br["ttm"] * (vm[t] * vm[f] * cos(va[t] - va[f]))
                         + (-br["b"] * br["tr"] + br["g"] * br["ti"]) / br["ttm"] * (vm[t] * vm[f] * sin(va[t] - va[f])))
        for br in ref["branch"]:
            f, t = br["f_bus"], br["t_bus"]
            m.constraint(q[br["t_idx"]], "==",
                         -(br["b"] + br["b_to"]) * vm[t] ** 2
                         - (-br["b"] * br["tr"] + br["g"] * br["ti"]) / br["ttm"] * (vm[t] * vm[f] * cos(va[t] - va[f]))
                         + (-br["g"] * br["tr"] - br["b"] * br["ti"]) / br["ttm"] * (vm[t] * vm[f] * sin(va[t] - va[f])))
        for br in ref["branch"]:
            m.constraint(va[br["f_bus"]] - va[br["t_bus"]], (br["angmin"], br["angmax"]))
        for br in ref["branch"]:
            m.constraint(p[br["f_idx"]] ** 2 + q[br["f_idx"]] ** 2, "<=", br["rate_a"])
        for br in ref["branch"]:
            m.constraint(p[br["t_idx"]] ** 2 + q[br["t_idx"]] ** 2, "<=", br["rate_a"])
        for i in buses:
            lhs_p = sum(p[a] for a in ref["bus_arcs"][i])
            lhs_q = sum(q[a] for a in ref["bus_arcs"][i])
            rhs_p = sum(pg[g] for g in ref["bus_gens"][i]) - B[i]["pd"] - 0.0 * vm[i] ** 2
            rhs_q = sum(qg[g] for g in ref["bus_gens"][i]) - B[i]["qd"] + 0.0 * vm[i] ** 2
            if theta is not None:
                rhs_p = theta[bpos[i]] + rhs_p
                rhs_q = theta[nbus + bpos[i]] + rhs_q
            m.constraint(lhs_p, "==", rhs_p)
            m.constraint(lhs_q, "==", rhs_q)

    stage(va0, vm0, pg0, qg0, p0, q0)
    stage(va, vm, pg, qg, p, q, theta=th)
    # ramping constraints — opf.jl:282-283
    for i, g in G.items():
        d = 0.1 * (g["pmax"] - g["pmin"])
        m.constraint(pg0[i] - pg[i], (-d, d))
    for i, g in G.items():
        d = 0.1 * (g["qmax"] - g["qmin"])
        m.constraint(qg0[i] - qg[i], (-d, d))
    return m
