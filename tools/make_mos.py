"""Generate the cached fixture MO matrices (qmctorch_b200/data/*_mos.json).

One-electron (core-Hamiltonian) orbitals: solve  H_core C = S C e  with
H_core = T + V_ne evaluated in closed form (McMurchie-Davidson Hermite
expansion) over the contracted Cartesian Gaussians of
qmctorch_b200.molecules.build_basis, normalised per primitive like the
reference (qmctorch/wavefunction/orbitals/norm_orbital.py:136-161).
Run once, offline:  python tools/make_mos.py [key ...]
"""

import json
import math
import os
import sys
from functools import lru_cache

import numpy as np
from scipy.linalg import eigh
from scipy.special import hyp1f1

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qmctorch_b200.molecules import Molecule, fixture_spec, build_basis, _parse_atoms, _Z  # noqa


def dfact(n):
    return 1.0 if n <= 0 else float(np.prod(np.arange(n, 0, -2)))


def gnorm(a, b, c, alpha):
    pref = (2 * alpha / math.pi) ** 0.75
    out = pref
    for k in (a, b, c):
        out *= (4 * alpha) ** (k / 2) / math.sqrt(dfact(2 * k - 1))
    return out


def hermite_E(i, j, t, Q, a, b):
    p = a + b
    q = a * b / p

    @lru_cache(maxsize=None)
    def E(i, j, t):
        if t < 0 or t > i + j:
            return 0.0
        if i == j == t == 0:
            return math.exp(-q * Q * Q)
        if j == 0:
            return (E(i - 1, j, t - 1) / (2 * p) - q * Q / a * E(i - 1, j, t)
                    + (t + 1) * E(i - 1, j, t + 1))
        return (E(i, j - 1, t - 1) / (2 * p) + q * Q / b * E(i, j - 1, t)
                + (t + 1) * E(i, j - 1, t + 1))
    return E(i, j, t)


def overlap(a, la, A, b, lb, B):
    p = a + b
    s = 1.0
    for d in range(3):
        s *= hermite_E(la[d], lb[d], 0, A[d] - B[d], a, b)
    return s * (math.pi / p) ** 1.5


def kinetic(a, la, A, b, lb, B):
    l2, m2, n2 = lb
    t0 = b * (2 * (l2 + m2 + n2) + 3) * overlap(a, la, A, b, lb, B)
    t1 = -2 * b * b * (overlap(a, la, A, b, (l2 + 2, m2, n2), B)
                       + overlap(a, la, A, b, (l2, m2 + 2, n2), B)
                       + overlap(a, la, A, b, (l2, m2, n2 + 2), B))
    t2 = 0.0
    if l2 > 1:
        t2 += l2 * (l2 - 1) * overlap(a, la, A, b, (l2 - 2, m2, n2), B)
    if m2 > 1:
        t2 += m2 * (m2 - 1) * overlap(a, la, A, b, (l2, m2 - 2, n2), B)
    if n2 > 1:
        t2 += n2 * (n2 - 1) * overlap(a, la, A, b, (l2, m2, n2 - 2), B)
    return t0 + t1 - 0.5 * t2


def boys(n, T):
    return hyp1f1(n + 0.5, n + 1.5, -T) / (2 * n + 1)


def nuclear(a, la, A, b, lb, B, C):
    p = a + b
    P = (a * np.asarray(A) + b * np.asarray(B)) / p
    PC = P - np.asarray(C)
    RPC2 = float(PC @ PC)
    L = sum(la) + sum(lb)
    F = [boys(n, p * RPC2) for n in range(L + 1)]

    @lru_cache(maxsize=None)
    def R(t, u, v, n):
        if t < 0 or u < 0 or v < 0:
            return 0.0
        if t == u == v == 0:
            return (-2 * p) ** n * F[n]
        if t > 0:
            return (t - 1) * R(t - 2, u, v, n + 1) + PC[0] * R(t - 1, u, v, n + 1)
        if u > 0:
            return (u - 1) * R(t, u - 2, v, n + 1) + PC[1] * R(t, u - 1, v, n + 1)
        return (v - 1) * R(t, u, v - 2, n + 1) + PC[2] * R(t, u, v - 1, n + 1)

    Ex = [hermite_E(la[0], lb[0], t, A[0] - B[0], a, b) for t in range(la[0] + lb[0] + 1)]
    Ey = [hermite_E(la[1], lb[1], t, A[1] - B[1], a, b) for t in range(la[1] + lb[1] + 1)]
    Ez = [hermite_E(la[2], lb[2], t, A[2] - B[2], a, b) for t in range(la[2] + lb[2] + 1)]
    val = 0.0
    for t, ex in enumerate(Ex):
        for u, ey in enumerate(Ey):
            for v, ez in enumerate(Ez):
                val += ex * ey * ez * R(t, u, v, 0)
    return val * 2 * math.pi / p


def core_mos(mol):
    b = mol.basis
    nb = len(b.bas_exp)
    atom_of = np.repeat(np.arange(mol.natom), b.nshells)
    prim = []
    for ip in range(nb):
        l = (int(b.bas_kx[ip]), int(b.bas_ky[ip]), int(b.bas_kz[ip]))
        al = float(b.bas_exp[ip])
        prim.append((al, l, tuple(mol.atom_coords[atom_of[ip]]),
                     float(b.bas_coeffs[ip]) * gnorm(*l, al)))
    ao_prims = [[] for _ in range(b.nao)]
    for ip in range(nb):
        ao_prims[b.index_ctr[ip]].append(prim[ip])
    n = b.nao
    S = np.zeros((n, n))
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            s = h = 0.0
            for (a, la, A, ca) in ao_prims[i]:
                for (bb, lb, B, cb) in ao_prims[j]:
                    w = ca * cb
                    s += w * overlap(a, la, A, bb, lb, B)
                    t = kinetic(a, la, A, bb, lb, B)
                    v = 0.0
                    for C, Z in zip(mol.atom_coords, mol.atomic_number):
                        v -= Z * nuclear(a, la, A, bb, lb, B, C)
                    h += w * (t + v)
            S[i, j] = S[j, i] = s
            H[i, j] = H[j, i] = h
    # drop the s-type contaminant of cartesian d shells like cart2sph would:
    # canonical orthogonalisation with a threshold keeps the matrix well conditioned
    e, C = eigh(H, S)
    return C, e, S


def main(keys):
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                           "qmctorch_b200", "data")
    os.makedirs(out_dir, exist_ok=True)
    for key in keys:
        spec = fixture_spec(key)
        names, coords = _parse_atoms(spec["atom"], spec["unit"])
        nao = build_basis(names, coords, spec["basis"]).nao
        mol = Molecule(mos=np.eye(nao), **spec)
        C, e, S = core_mos(mol)
        # fixed sign convention: largest-magnitude coefficient of each column positive
        for k in range(C.shape[1]):
            if C[np.argmax(np.abs(C[:, k])), k] < 0:
                C[:, k] = -C[:, k]
        path = os.path.join(out_dir, "%s_%s_mos.json" % (mol.name, spec["basis"].lower()))
        with open(path, "w") as f:
            json.dump({"key": key, "how": "core-Hamiltonian eigenvectors, tools/make_mos.py",
                       "orbital_energies": [float(x) for x in e],
                       "mos": [[float("%.15e" % x) for x in row] for row in C]}, f)
        print(key, "nao", nao, "lowest eps", e[:6], "S diag", np.diag(S)[:4], "->", path)


if __name__ == "__main__":
    main(sys.argv[1:] or ["h2", "lih_sto3g", "lih", "h2o", "c4h6"])
