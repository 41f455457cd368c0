"""Per-primitive normalisation constants, computed once on the host at construction
(qmctorch/wavefunction/orbitals/norm_orbital.py:8-161; frozen afterwards, see
atomic_orbitals.py:90-94)."""
import math

import numpy as np


def _odd_factorial(n):
    """(n)!! with (-1)!! = 0!! = 1, utils/algebra_utils.py:45-55."""
    n = int(n)
    out = 1.0
    while n > 1:
        out *= n
        n -= 2
    return out


def _norm_spherical(basis):
    """norm_slater_spherical / norm_gaussian_spherical (norm_orbital.py:45-93): functions of the radial
    power ``bas_n`` and the exponent only."""
    n = np.asarray(basis.bas_n, dtype=np.float64)
    alpha = np.asarray(basis.bas_exp, dtype=np.float64)
    if basis.radial_type.startswith("sto"):
        nfact = np.array([float(math.factorial(int(2 * k))) for k in n])
        return (2.0 * alpha) ** n * np.sqrt(2.0 * alpha / nfact)
    if basis.radial_type.startswith("gto"):
        n1 = n + 1.0
        a = alpha ** (0.25 * (2.0 * n1 + 1.0))
        b = 2.0 ** (2.0 * n1 + 1.5)
        c = np.array([_odd_factorial(2 * int(k) - 1) for k in n1]) * math.pi ** 0.5
        return np.sqrt(b / c) * a
    raise ValueError("%s is not a valid radial_type" % basis.radial_type)


def atomic_orbital_norm(basis):
    if basis.harmonics_type == "sph":
        return _norm_spherical(basis)
    kx = np.asarray(basis.bas_kx).astype(int)
    ky = np.asarray(basis.bas_ky).astype(int)
    kz = np.asarray(basis.bas_kz).astype(int)
    alpha = np.asarray(basis.bas_exp, dtype=np.float64)
    out = np.empty_like(alpha)
    if basis.radial_type.startswith("gto"):
        for i, (a, b, c, z) in enumerate(zip(kx, ky, kz, alpha)):
            v = (2.0 * z / math.pi) ** 0.75
            for k in (a, b, c):
                v *= (4.0 * z) ** (k / 2.0) / math.sqrt(_odd_factorial(2 * k - 1))
            out[i] = v
    elif basis.radial_type.startswith("sto"):
        kr = np.asarray(basis.bas_kr).astype(int)
        for i, (a, b, c, n, z) in enumerate(zip(kx, ky, kz, kr, alpha)):
            big_l = a + b + c + n + 1
            pref = 4.0 * math.pi * math.factorial(2 * big_l) / (2.0 * z) ** (2 * big_l + 1)
            num = _odd_factorial(2 * a - 1) * _odd_factorial(2 * b - 1) * _odd_factorial(2 * c - 1)
            den = _odd_factorial(2 * (a + b + c) + 1)
            out[i] = math.sqrt(1.0 / (pref * num / den))
    else:
        raise ValueError("%s is not a valid radial_type" % basis.radial_type)
    return out
