"""CPU oracle of the output path -- test infrastructure only (tests/ may import it, the product never does).

Restates, in numpy, what the reference computes per time stamp in print_results (source/SolarCell.cpp:1826-1858):
  * dealii::DataOut::build_patches with the default of one subdivision evaluates the solution at the four vertices of
    every cell (one patch per cell, deal.II lexicographic vertex order);
  * PostProcessor::compute_derived_quantities_vector (source/PostProcessor.cpp:80-123) multiplies the vector part by
    scale_current (carriers) or scale_elec_field (Poisson), leaves the densities unscaled and multiplies the potential
    by scale_potential; the scales are set in the constructor (source/PostProcessor.cpp:14-18).
deal.II is not available here: the finite elements are the closed forms of SURVEY App. B --
DGQ1 nodal values sit at the vertices; the RT0 field is J psihat / det J with psihat_f = (1-xi,0), (xi,0), (0,1-eta),
(0,eta).  Parity for this path is unpinned against the reference binary like everything else."""
import numpy as np


def scales(characteristic_length, characteristic_density, characteristic_time):
    """{potential, field, density, current}: reference source/PostProcessor.cpp:14-18"""
    potential = 0.02585
    field = 0.2585 / characteristic_length
    density = characteristic_density
    current = 1.6e-19 * density * characteristic_length / characteristic_time
    return np.array([potential, field, density, current])


def carrier_patches(solution, scale_current):
    """solution: [Jx | Jy | rho] blocks, 4 nodal values per cell -> (current [4n,3], density [4n])"""
    u = np.asarray(solution)
    n4 = u.size // 3
    current = np.zeros((n4, 3))
    current[:, 0] = scale_current * u[:n4]
    current[:, 1] = scale_current * u[n4:2 * n4]
    return current, u[2 * n4:].copy()


def poisson_patches(vertices, face_dof, n_rt, solution, scale_field, scale_potential):
    """vertices [n,4,2], face_dof [n,4] -> (field [4n,3], potential [4n]) at the patch vertices"""
    v = np.asarray(vertices)
    X = np.asarray(solution)
    n = v.shape[0]
    Xf = X[np.asarray(face_dof)]  # [n,4]
    field = np.zeros((n, 4, 3))
    for a in range(4):
        xi, eta = float(a & 1), float(a >> 1)
        xxi = (v[:, 1, 0] - v[:, 0, 0]) * (1 - eta) + (v[:, 3, 0] - v[:, 2, 0]) * eta
        yxi = (v[:, 1, 1] - v[:, 0, 1]) * (1 - eta) + (v[:, 3, 1] - v[:, 2, 1]) * eta
        xeta = (v[:, 2, 0] - v[:, 0, 0]) * (1 - xi) + (v[:, 3, 0] - v[:, 1, 0]) * xi
        yeta = (v[:, 2, 1] - v[:, 0, 1]) * (1 - xi) + (v[:, 3, 1] - v[:, 1, 1]) * xi
        det = xxi * yeta - xeta * yxi
        dhx = Xf[:, 0] * (1 - xi) + Xf[:, 1] * xi
        dhy = Xf[:, 2] * (1 - eta) + Xf[:, 3] * eta
        field[:, a, 0] = scale_field * (xxi * dhx + xeta * dhy) / det
        field[:, a, 1] = scale_field * (yxi * dhx + yeta * dhy) / det
    potential = np.repeat(scale_potential * X[n_rt:n_rt + n], 4)
    return field.reshape(-1, 3), potential


_GAUSS_X = np.array([0.5 - np.sqrt(0.15), 0.5, 0.5 + np.sqrt(0.15)])
_GAUSS_W = np.array([5.0, 8.0, 5.0]) / 18.0


def _face_point(f, t):
    return (float(f), t) if f < 2 else (t, float(f - 2))


def _shape(xi, eta):
    return np.array([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta])


def interface_currents(semi_vertices, pairs, states, k_et, k_ht, rho_n_e, rho_p_e):
    """I-V post-processing (no counterpart in the reference; integrals of its two interface terms,
    source/SolarCell.cpp:1265-1347): (int k_et (rho_n - rho_n^e) rho_o ds, int k_ht (rho_p - rho_p^e) rho_r ds) over the
    interface faces, QGauss(3) per face, the same quadrature index on both sides.
    pairs = (semi_cell, semi_face, elec_cell, elec_face) arrays; states = four [Jx|Jy|rho] vectors."""
    v = np.asarray(semi_vertices)
    rho = [np.asarray(s)[2 * (s.size // 3):].reshape(-1, 4) for s in states]
    i_et = i_ht = 0.0
    for cs, fs, ce, fe in zip(*pairs):
        a, b = {0: (0, 2), 1: (1, 3), 2: (0, 1), 3: (2, 3)}[int(fs)]  # the two vertices of face fs
        ds = np.hypot(*(v[cs, b] - v[cs, a]))
        for t, w in zip(_GAUSS_X, _GAUSS_W):
            n_s, n_e = _shape(*_face_point(int(fs), t)), _shape(*_face_point(int(fe), t))
            i_et += k_et * (n_s @ rho[0][cs] - rho_n_e) * (n_e @ rho[3][ce]) * ds * w
            i_ht += k_ht * (n_s @ rho[1][cs] - rho_p_e) * (n_e @ rho[2][ce]) * ds * w
    return np.array([i_et, i_ht])
