"""CPU study behind DESIGN.md section 5b ("is 12.85 GB of factor traffic per step necessary?").

Would a preconditioned Krylov method on the Schur-reduced density systems S = S_uu - G_uq A_qq^-1 G_qu beat the direct
solve in bytes?  Measures, on the default input at refinements g (and l = 1):
  * the spectrum of the 4x4-block-Jacobi-preconditioned S of every species (dense, g = 2) -- VERDICT r1 expected
    kappa ~ 1.0-1.3 for the redox species (mobility 2.6e-6 against 1/dt = 20);
  * PCG iteration counts to 1e-11 with block-Jacobi, and with a two-level method (block-Jacobi + correction on the
    conforming Q1 subspace, exact and inexact coarse solves; additive and symmetric multiplicative).
Result (profiles/r02_iterative_solver_study.txt): kappa is 380 at g = 2 and grows like 1/h^2 -- the LDG penalty
tau/h is NOT scaled by the mobility (reference source/LDG.cpp:455-618: sigma = penalty / h on every interior face),
so S is "jump penalty + small mass", not "mass + small stiffness"; block-Jacobi PCG needs 350-600 iterations at
g = 3-4 and the h-independent two-level method 37-74 -- at 59 MB of S per iteration more bytes and far more latency
than one pass over the 2.8 GB factor.  What IS free: the reductant and oxidant matrices are bitwise identical
(equal mobilities), so one factorisation serves both right-hand sides (cuda/context.cu: shared_pair).

    python scripts/iterative_solver_study.py 3 4
"""
import sys
import warnings

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

warnings.filterwarnings("ignore")
sys.path.insert(0, ".")
import pecs_b200 as pecs  # noqa: E402  (host library only: mesh tables and the constant matrices)


def schur(A):
    A = A.tocsr()
    n = A.shape[0] // 12
    nq = 8 * n
    Aqq, Gqu, Guq, Suu = A[:nq, :nq].tocsc(), A[:nq, nq:].tocsc(), A[nq:, :nq].tocsr(), A[nq:, nq:]
    T2 = sp.csr_matrix(spl.spsolve(Aqq, Gqu))
    S = (Suu - Guq @ T2).tocsr()
    S.data[np.abs(S.data) < 1e-13 * np.abs(S.data).max()] = 0
    S.eliminate_zeros()
    return S


def study(g, l=1):
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l))
    prob.setup_full_system_host()
    print(f"== global refinements {g}, local {l}")
    same = (prob.matrix(2) != prob.matrix(3)).nnz == 0
    print(f"  reductant and oxidant matrices identical: {same}")
    for s in (0, 2):
        V = np.asarray(prob.mesh(s // 2)["vertices"]).reshape(-1, 4, 2)
        n = V.shape[0]
        keys, rows, cols = {}, [], []
        for c in range(n):
            for a in range(4):
                k = (round(V[c, a, 0] * 2 ** 40), round(V[c, a, 1] * 2 ** 40))
                rows.append(4 * c + a)
                cols.append(keys.setdefault(k, len(keys)))
        P = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(4 * n, len(keys)))  # conforming Q1 -> DG nodal
        S = schur(prob.matrix(s))
        nu = 4 * n
        Dinv = sp.block_diag([sp.csr_matrix(np.linalg.inv(S[4 * c:4 * c + 4, 4 * c:4 * c + 4].toarray()))
                              for c in range(n)]).tocsr()
        if nu <= 400:
            ev = np.linalg.eigvals((Dinv @ S).toarray()).real
            print(f"  species {s}: block-Jacobi-preconditioned spectrum [{ev.min():.5f}, {ev.max():.5f}], kappa {ev.max() / ev.min():.0f}")
        Sc = (P.T @ S @ P).tocsc()
        lu_c = spl.splu(Sc)
        dc = 1.0 / Sc.diagonal()
        b = S @ np.random.default_rng(0).standard_normal(nu)

        def run(prec, name):
            its = [0]
            spl.cg(S, b, rtol=1e-11, maxiter=5000, M=spl.LinearOperator((nu, nu), matvec=prec),
                   callback=lambda xk: its.__setitem__(0, its[0] + 1))
            print(f"  species {s} ({nu} unknowns, {S.nnz / nu:.1f} nnz/row): {name}: {its[0]} iterations to 1e-11")

        run(lambda r: Dinv @ r, "PCG, 4x4 block-Jacobi")
        run(lambda r: Dinv @ r + P @ lu_c.solve(P.T @ r), "PCG, additive two-level (exact coarse solve)")
        run(lambda r: Dinv @ r + P @ spl.cg(Sc, P.T @ r, rtol=1e-30, maxiter=6, M=sp.diags(dc))[0],
            "PCG, additive two-level (6 Jacobi-PCG coarse iterations)")

        def mult(r):
            z = 0.7 * (Dinv @ r)
            z = z + P @ lu_c.solve(P.T @ (r - S @ z))
            return z + 0.7 * (Dinv @ (r - S @ z))
        run(mult, "PCG, symmetric multiplicative two-level")
    prob.close()


if __name__ == "__main__":
    for g in [int(a) for a in sys.argv[1:]] or [2, 3, 4]:
        study(g)
