"""oracle/grid.py -- TEST INFRASTRUCTURE ONLY (CPU oracle): independent restatement of the reference's mesh generator
and parameter scaling for the PRODUCTION geometry, so that the oracle (and bench.py's CPU arm) can run without the
product library.

Restates, with numpy and closed-form index arithmetic instead of a quadtree:
  * reference source/Grid.cpp:44-133   make_grids: trapezoid coarse cells (bulk + boundary layer per subdomain),
                                       merge_triangulations, refine_global(g), l refinements of the boundary-layer
                                       cells by material id
  * reference source/Grid.cpp:340-461  boundary tagging (Dirichlet / Neumann / Schottky by EXACT coordinate tests,
                                       including the radius-one quirk on the bottom edge)
  * reference include/Parameters.hpp:181-242  singular-perturbation scaling of the .prm values
  * deal.II conventions: active cells level-major, children in lexicographic order (so the cells of
    a refined coarse cell come in Morton order, x bit lowest), vertices v0=(0,0) v1=(1,0) v2=(0,1) v3=(1,1),
    faces 0: xi=0, 1: xi=1, 2: eta=0, 3: eta=1, new vertices at edge / cell midpoints.

Scope: local refinements 0 or 1 (the default input and cfg3; no 2:1 smoothing arises).  For l >= 2 the tests keep
feeding the oracle the product's tables.  tests/test_oracle_grid.py compares these tables with the product's
(own quadtree, host/Triangulation.cpp) entry by entry and with the counts of SURVEY App. D.
"""
import numpy as np

# reference include/Grid.hpp material ids and boundary ids
SEMICONDUCTOR, SEMI_LAYER, ELECTROLYTE, ELEC_LAYER = 0, 1, 2, 3
INTERFACE, DIRICHLET, NEUMANN, SCHOTTKY = 0, 1, 2, 3
FACE_SAME_LEVEL, FACE_BOUNDARY, FACE_HAS_CHILDREN, FACE_COARSER = 0, 1, 2, 3

DEFAULT_PRM = {  # reference input_file.prm
    "end time": 5e1, "global refinements": 4, "local refinements": 1, "time stamps": 100, "time step size": 0.05,
    "electron mobility": 1350.0, "electron recombination velocity": 3e5, "electron transfer rate": 1e-19,
    "hole mobility": 480.0, "hole recombination velocity": 2.9e5, "hole transfer rate": 1e-14,
    "boundary layer": 0.1, "mesh height": 1.0, "mesh length": 1.0, "radius one": 0.3, "radius two": 0.6,
    "oxidant mobility": 1.0, "reductant mobility": 1.0,
    "absorption coefficient": 1.74974e5, "applied bias": 0.0, "built in bias": 0.41, "characteristic density": 1.0e16,
    "characteristic length": 1.0e-4, "characteristic time": 1.0e-12, "electrolyte permittivity": 1000.0,
    "illumination status": True, "insulated": True, "intrinsic density": 2.564e9, "photon flux": 1.2e17,
    "schottky bias": 0.0, "schottky status": True, "semiconductor permittivity": 11.9,
    "electron recombination time": 5e-5, "hole recombination time": 5e-5,
    "srh recombination": False,  # not a key of the reference: its SRH_Recombination returns 0.0 (SolarCell.hpp:86-98)
}

# include/pecs_b200.h PECS_P_* slots of the 32-double parameter block the oracle's C API takes
PARAM_SLOTS = ["delta_t", "penalty", "mu_n", "mu_p", "mu_r", "mu_o", "eps_s", "eps_e", "lambda2", "k_et", "k_ht", "v_n",
               "v_p", "gen_flux", "gen_alpha", "gen_location", "rho_n_e", "rho_p_e", "rho_r_e", "rho_o_e", "phi_bi",
               "phi_app", "phi_sch", "sch_location", "transient", "srh", "n_intrinsic", "tau_n", "tau_p"]


def scaled_parameters(prm=None):
    """reference include/Parameters.hpp:181-242 (+ SolarCell.cpp:15-92 for where each value goes)"""
    p = dict(DEFAULT_PRM)
    p.update(prm or {})
    U_T, q, eps0 = 0.02585, 1.62e-19, 8.85e-14  # PhysicalConstants, Parameters.hpp:9-11
    L, T, C = p["characteristic length"], p["characteristic time"], p["characteristic density"]
    mob = T * U_T / (L * L)
    lit = bool(p["illumination status"])
    v = {
        "delta_t": p["time step size"], "penalty": 1.0,
        "mu_n": p["electron mobility"] * mob, "mu_p": p["hole mobility"] * mob,
        "mu_r": p["reductant mobility"] * mob, "mu_o": p["oxidant mobility"] * mob,
        "eps_s": p["semiconductor permittivity"], "eps_e": p["electrolyte permittivity"],
        "lambda2": (U_T * eps0) / (q * C * L * L),
        "k_et": p["electron transfer rate"] * (T * C / L), "k_ht": p["hole transfer rate"] * (T * C / L),
        "v_n": p["electron recombination velocity"] * (T / L), "v_p": p["hole recombination velocity"] * (T / L),
        "gen_flux": p["photon flux"] * (T / C) if lit else 0.0,
        "gen_alpha": p["absorption coefficient"] * L if lit else 0.0,
        "gen_location": p["mesh height"] if lit else 0.0,
        "rho_n_e": 2.0, "rho_p_e": 0.0, "rho_r_e": 30.0, "rho_o_e": 29.0,  # InitialConditions.cpp:20,43,61,79
        "phi_bi": p["built in bias"] / U_T, "phi_app": p["applied bias"] / U_T, "phi_sch": p["schottky bias"] / U_T,
        "sch_location": p["mesh height"], "transient": 1.0,
        "srh": 1.0 if p["srh recombination"] else 0.0, "n_intrinsic": p["intrinsic density"] / C,
        "tau_n": p["electron recombination time"] / T, "tau_p": p["hole recombination time"] / T,
    }
    out = np.zeros(32)
    for k, name in enumerate(PARAM_SLOTS):
        out[k] = v[name]
    return out


def _refine_lattice(X):
    """one isotropic bisection of a structured vertex lattice: new vertices at edge midpoints and cell centres"""
    n = X.shape[0]
    Z = np.empty((2 * n - 1, 2 * n - 1))
    Z[::2, ::2] = X
    Z[1::2, ::2] = 0.5 * (X[:-1, :] + X[1:, :])
    Z[::2, 1::2] = 0.5 * (X[:, :-1] + X[:, 1:])
    Z[1::2, 1::2] = 0.5 * (Z[1::2, :-2:2] + Z[1::2, 2::2])
    return Z


def _morton(i, j, depth):
    z = np.zeros_like(i)
    for k in range(depth):
        z |= ((i >> k) & 1) << (2 * k)
        z |= ((j >> k) & 1) << (2 * k + 1)
    return z


class _Block:
    """one coarse cell refined `depth` times: lattice[i, j] = vertex (i along xi), cells in Morton order"""

    def __init__(self, corners, material, depth):
        (x00, y00), (x10, y10), (x01, y01), (x11, y11) = corners  # v0 v1 v2 v3
        X = np.array([[x00, x01], [x10, x11]], float)
        Y = np.array([[y00, y01], [y10, y11]], float)
        for _ in range(depth):
            X, Y = _refine_lattice(X), _refine_lattice(Y)
        self.X, self.Y, self.depth, self.material, self.n = X, Y, depth, material, 1 << depth
        n = self.n
        ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        self.z = _morton(ii, jj, depth)       # z[i, j] = position of cell (i, j) inside the block
        order = np.argsort(self.z.ravel())
        self.i_of, self.j_of = ii.ravel()[order], jj.ravel()[order]
        self.first = 0                        # index of the block's first cell in the active order

    def cell(self, i, j):
        return self.first + self.z[i, j]

    def diameter(self, i, j):
        X, Y = self.X, self.Y
        d1 = np.hypot(X[i + 1, j + 1] - X[i, j], Y[i + 1, j + 1] - Y[i, j])
        d2 = np.hypot(X[i, j + 1] - X[i + 1, j], Y[i, j + 1] - Y[i + 1, j])
        return np.maximum(d1, d2)


def _tables(blocks, joins):
    """blocks: list of _Block in coarse-cell order; joins: (left block index, right block index) sharing the edge
    xi=1 of the left one and xi=0 of the right one.  Active order: level-major, then coarse-cell order."""
    order = sorted(range(len(blocks)), key=lambda b: (blocks[b].depth, b))
    n_cells = 0
    for b in order:
        blocks[b].first = n_cells
        n_cells += blocks[b].n ** 2
    m = {"n_cells": n_cells, "vertices": np.zeros((n_cells, 4, 2)), "material_id": np.zeros(n_cells, np.int32),
         "level": np.zeros(n_cells, np.int32), "face_kind": np.full((n_cells, 4), FACE_BOUNDARY, np.int32),
         "neighbor": np.full((n_cells, 4), -1, np.int32), "neighbor2": np.full((n_cells, 4), -1, np.int32),
         "boundary_id": np.full((n_cells, 4), -1, np.int32), "nb_parent_diameter": np.zeros((n_cells, 4))}
    for B in blocks:
        i, j = B.i_of, B.j_of
        c = B.first + np.arange(B.n ** 2)
        for a, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
            m["vertices"][c, a, 0] = B.X[i + di, j + dj]
            m["vertices"][c, a, 1] = B.Y[i + di, j + dj]
        m["material_id"][c] = B.material
        m["level"][c] = B.depth
        for f, (di, dj) in enumerate(((-1, 0), (1, 0), (0, -1), (0, 1))):
            ni, nj = i + di, j + dj
            inside = (ni >= 0) & (ni < B.n) & (nj >= 0) & (nj < B.n)
            m["face_kind"][c[inside], f] = FACE_SAME_LEVEL
            m["neighbor"][c[inside], f] = B.first + B.z[ni[inside], nj[inside]]
    for left, right in joins:
        A, B = blocks[left], blocks[right]
        if A.depth == B.depth:
            j = np.arange(A.n)
            ca, cb = A.cell(A.n - 1, j), B.cell(0, j)
            m["face_kind"][ca, 1] = FACE_SAME_LEVEL
            m["neighbor"][ca, 1] = cb
            m["face_kind"][cb, 0] = FACE_SAME_LEVEL
            m["neighbor"][cb, 0] = ca
            continue
        # one level apart: C = coarse block, F = fine block; fc / ff = the faces that meet
        (C, ic, fc), (F, i_f, ff) = ((A, A.n - 1, 1), (B, 0, 0)) if A.depth < B.depth else ((B, 0, 0), (A, A.n - 1, 1))
        assert F.depth == C.depth + 1, "only one level of local refinement is restated here"
        j = np.arange(C.n)
        cc = C.cell(ic, j)
        m["face_kind"][cc, fc] = FACE_HAS_CHILDREN
        m["neighbor"][cc, fc] = F.cell(i_f, 2 * j)
        m["neighbor2"][cc, fc] = F.cell(i_f, 2 * j + 1)
        # diameter of the refined (inactive) neighbour: the fine block's cell one level up
        parent = _Block([(F.X[0, 0], F.Y[0, 0]), (F.X[-1, 0], F.Y[-1, 0]), (F.X[0, -1], F.Y[0, -1]),
                         (F.X[-1, -1], F.Y[-1, -1])], F.material, C.depth)
        m["nb_parent_diameter"][cc, fc] = parent.diameter(np.full(C.n, 0 if i_f == 0 else parent.n - 1), j)
        jf = np.arange(F.n)
        cf = F.cell(i_f, jf)
        m["face_kind"][cf, ff] = FACE_COARSER
        m["neighbor"][cf, ff] = C.cell(ic, jf // 2)
        m["neighbor2"][cf, ff] = jf % 2
    m["boundary_id"][m["face_kind"] == FACE_BOUNDARY] = 0
    return m


def _face_centers(m):
    v = m["vertices"]
    pairs = ((0, 2), (1, 3), (0, 1), (2, 3))  # vertices of faces 0..3
    return np.stack([0.5 * (v[:, a] + v[:, b]) for a, b in pairs], axis=1)  # [n][4][2]


def _tag(m, rule):
    fc = _face_centers(m)
    ids = rule(fc[..., 0], fc[..., 1])
    sel = (m["face_kind"] == FACE_BOUNDARY) & (ids >= 0)
    m["boundary_id"][sel] = ids[sel]


def make_grids(prm=None, full_system=True):
    """reference Grid.cpp:44-133: returns (semiconductor, electrolyte, Poisson) mesh tables in the layout of
    oracle.binding.Oracle.set_mesh"""
    p = dict(DEFAULT_PRM)
    p.update(prm or {})
    g, l = int(p["global refinements"]), int(p["local refinements"])
    assert l in (0, 1), "oracle/grid.py restates local refinements 0 and 1 only"
    H, L, r1, r2 = p["mesh height"], p["mesh length"], p["radius one"], p["radius two"]
    bl = p["boundary layer"] if l > 0 else 0.0  # Grid.cpp:24-40
    layer = l > 0 and bl > 0

    def semi_blocks():
        xb, xt = r2 - bl, r1 - bl  # Grid.cpp:143-149, 176-182
        out = [_Block([(0, 0), (xb, 0), (0, H), (xt, H)], SEMICONDUCTOR, g)]
        if layer:
            out.append(_Block([(xb, 0), (r2, 0), (xt, H), (r1, H)], SEMI_LAYER, g + l))
        return out

    def elec_blocks():
        xb, xt = r2 + bl, r1 + bl  # Grid.cpp:239-246, 269-276
        bulk = _Block([(xb, 0), (L, 0), (xt, H), (L, H)], ELECTROLYTE, g)
        if not layer:
            return [bulk]
        return [_Block([(r2, 0), (xb, 0), (r1, H), (xt, H)], ELEC_LAYER, g + l), bulk]

    chain = lambda blocks: [(k, k + 1) for k in range(len(blocks) - 1)]  # noqa: E731
    sb, eb = semi_blocks(), elec_blocks()
    semi = _tables(sb, chain(sb))
    elec = _tables(eb, chain(eb))
    pb = semi_blocks() + elec_blocks() if full_system else semi_blocks()
    poisson = _tables(pb, chain(pb))

    def dirichlet(x, y):  # Grid.cpp:340-372: every outer boundary face, exact coordinate tests
        return np.where((x == 0.0) | (x == L) | (y == 0.0) | (y == H), DIRICHLET, -1)

    def neumann(x, y):  # Grid.cpp:374-428, bottom edge tested against radius ONE
        ids = np.full(x.shape, -1)
        ids = np.where(y == H, np.where(x > r1, NEUMANN, DIRICHLET), ids)
        ids = np.where(y == 0.0, np.where(x > r1, NEUMANN, DIRICHLET), ids)
        return np.where(x == 0.0, NEUMANN, ids)

    def schottky(x, y):  # Grid.cpp:430-461: the whole top edge
        return np.where(y == H, SCHOTTKY, -1)

    for m in (semi, elec, poisson):
        _tag(m, dirichlet)
    if p["insulated"]:
        for m in (poisson, semi, elec):
            _tag(m, neumann)
    if p["schottky status"]:
        _tag(semi, schottky)
        _tag(poisson, schottky)
    return semi, elec, poisson


def make_oracle(prm=None):
    """an Oracle of the full production system built WITHOUT the product library: own grid, own parameter scaling"""
    from .binding import Oracle
    semi, elec, poisson = make_grids(prm, True)
    o = Oracle(scaled_parameters(prm), True)
    o.set_mesh(0, semi)
    o.set_mesh(1, elec)
    o.set_mesh(2, poisson)
    return o
