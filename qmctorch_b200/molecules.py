"""Fixture molecules: the ``mol`` / ``mol.basis`` data contract without pyscf.

The reference obtains ``mol.basis`` from an SCF front end (pyscf / ADF) that is
not available offline (``qmctorch/scf/calculator/pyscf.py:104-252`` builds the
namespace; ``qmctorch/wavefunction/orbitals/atomic_orbitals.py:27-94`` consumes
it).  The hot path only needs the *data contract*, so this module emits the very
same namespace from basis tables typed in here:

* primitives are expanded per Cartesian component exactly as
  ``pyscf.py:145-205`` does (``radial_type="gto_pure"``, ``harmonics_type="cart"``,
  ``bas_kr=0``, one ``index_ctr`` entry per primitive, ``nshells`` = number of
  flat primitives per atom);
* molecular orbitals come from a one-electron (core-Hamiltonian) diagonalisation
  computed with closed-form Cartesian-Gaussian integrals (`tools/make_mos.py`),
  cached in ``qmctorch_b200/data/*.json`` and column-normalised like
  ``scf/calculator/calculator_base.py:35-46``.  They are orthonormal, have the
  right nodal structure and are deterministic; they are NOT SCF quality and the
  basis exponents are typed from memory of the public tables ("approximate").
  Parity and throughput depend on neither.

Any object with these attributes (e.g. a real ``qmctorch.scf.Molecule``) can be
handed to :class:`qmctorch_b200.SlaterJastrow` instead.
"""

import json
import os
from types import SimpleNamespace

import numpy as np

ANGS2BOHR = 1.8897259886

_Z = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9}

# cartesian component order, pyscf.py:118-121
_KX = {0: [0], 1: [1, 0, 0], 2: [2, 1, 1, 0, 0, 0]}
_KY = {0: [0], 1: [0, 1, 0], 2: [0, 1, 0, 2, 1, 0]}
_KZ = {0: [0], 1: [0, 0, 1], 2: [0, 0, 1, 0, 1, 2]}

# basis tables: element -> list of shells (l, [exponents], [coefficients])
_BASIS = {
    "sto-3g": {
        "H": [
            (0, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454]),
        ],
        "Li": [
            (0, [16.1195750, 2.9362007, 0.7946505], [0.15432897, 0.53532814, 0.44463454]),
            (0, [0.6362897, 0.1478601, 0.0480887], [-0.09996723, 0.39951283, 0.70011547]),
            (1, [0.6362897, 0.1478601, 0.0480887], [0.15591627, 0.60768372, 0.39195739]),
        ],
    },
    "6-31g": {
        "H": [
            (0, [18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733]),
            (0, [0.1612778], [1.0]),
        ],
        "Li": [
            (
                0,
                [642.4189200, 96.7985150, 22.0911210, 6.2010703, 1.9351177, 0.6367358],
                [0.0021426, 0.0162089, 0.0773156, 0.2457860, 0.4701890, 0.3454708],
            ),
            (0, [2.3249184, 0.6324306, 0.0790534], [-0.0350917, -0.1912328, 1.0839878]),
            (1, [2.3249184, 0.6324306, 0.0790534], [0.0089415, 0.1410095, 0.9453637]),
            (0, [0.0359620], [1.0]),
            (1, [0.0359620], [1.0]),
        ],
    },
    "cc-pvdz": {
        "H": [
            (0, [13.01, 1.962, 0.4446, 0.122], [0.019685, 0.137977, 0.478148, 0.501240]),
            (0, [0.122], [1.0]),
            (1, [0.727], [1.0]),
        ],
        "O": [
            (
                0,
                [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013, 0.3023],
                [0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458,
                 -0.002585],
            ),
            (
                0,
                [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013, 0.3023],
                [-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368,
                 0.572759],
            ),
            (0, [0.3023], [1.0]),
            (1, [17.70, 3.854, 1.046, 0.2753], [0.043018, 0.228913, 0.508728, 0.460531]),
            (1, [0.2753], [1.0]),
            (2, [1.185], [1.0]),
        ],
    },
    # DZP without its polarisation shells (second point of the C4H6 basis-size sweep, BASELINE config 5)
    "dz": {
        "H": [
            (0, [19.2406, 2.8992, 0.6534], [0.032828, 0.231208, 0.817238]),
            (0, [0.1776], [1.0]),
        ],
        "C": [
            (
                0,
                [4232.61, 634.882, 146.097, 42.4974, 14.1892, 1.9666],
                [0.002029, 0.015535, 0.075411, 0.257121, 0.596555, 0.242517],
            ),
            (0, [5.1477], [1.0]),
            (0, [0.4962], [1.0]),
            (0, [0.1533], [1.0]),
            (1, [18.1557, 3.9864, 1.1429, 0.3594], [0.018534, 0.115442, 0.386206, 0.640089]),
            (1, [0.1146], [1.0]),
        ],
    },
    "dzp": {
        "H": [
            (0, [19.2406, 2.8992, 0.6534], [0.032828, 0.231208, 0.817238]),
            (0, [0.1776], [1.0]),
            (1, [1.0], [1.0]),
        ],
        "C": [
            (
                0,
                [4232.61, 634.882, 146.097, 42.4974, 14.1892, 1.9666],
                [0.002029, 0.015535, 0.075411, 0.257121, 0.596555, 0.242517],
            ),
            (0, [5.1477], [1.0]),
            (0, [0.4962], [1.0]),
            (0, [0.1533], [1.0]),
            (1, [18.1557, 3.9864, 1.1429, 0.3594], [0.018534, 0.115442, 0.386206, 0.640089]),
            (1, [0.1146], [1.0]),
            (2, [0.75], [1.0]),
        ],
    },
}

# uncontracted (ADF-style, scf/calculator/adf.py:189-296) tables: element -> list of (kx, ky, kz, kr, zeta).
# One AO per entry, bas_coeffs = 1, index_ctr = arange(nao).  Exponents are plausible double-zeta
# Slater values typed by hand ("approximate"); they exercise the sto / sto_pure / gto radial forms.
_P = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
_UNCONTRACTED = {
    "sto-dz": {
        "H": [(0, 0, 0, 0, 0.76), (0, 0, 0, 0, 1.28)] + [(a, b, c, 0, 1.25) for a, b, c in _P],
        "Li": [(0, 0, 0, 0, 2.45), (0, 0, 0, 0, 3.90), (0, 0, 0, 1, 0.65), (0, 0, 0, 1, 1.05)]
              + [(a, b, c, 0, 0.70) for a, b, c in _P],
    },
}

# spherical-harmonics Slater / Gaussian bases (harmonics_type "sph": radial r^n exp(-zeta r) or r^n exp(-zeta r^2)
# times the real Y_lm of spherical_harmonics.py:352-702): element -> list of (n, l, zeta); every m of a shell
# is one AO.  Hand-built, "approximate" exponents: the reference reaches this path only through hand-built or
# loaded bases (both of its calculators emit cartesian functions).
_SPHERICAL = {
    "sph-dz": {
        "H": [(0, 0, 1.24), (1, 0, 0.90), (1, 1, 1.20)],
        "Li": [(0, 0, 2.70), (1, 0, 0.65), (1, 1, 0.70), (2, 2, 0.90)],
    },
}

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _parse_atoms(atom, unit):
    names, coords = [], []
    conv = ANGS2BOHR if unit == "angs" else 1.0
    if unit not in ("angs", "bohr"):
        raise ValueError("unit should be angs or bohr")
    for a in atom.split(";"):
        d = a.split()
        if not d:
            continue
        names.append(d[0])
        coords.append([float(d[1]) * conv, float(d[2]) * conv, float(d[3]) * conv])
    return names, coords


def build_basis(atoms, atom_coords, basis_name):
    """Flat-primitive basis namespace in the layout of ``pyscf.py:145-252``."""
    table = _BASIS[basis_name.lower()]
    b = SimpleNamespace()
    b.radial_type = "gto_pure"
    b.harmonics_type = "cart"
    bas_coeff, bas_exp, index_ctr = [], [], []
    bas_kx, bas_ky, bas_kz, bas_l = [], [], [], []
    nshells = [0] * len(atoms)
    iao = 0
    for iat, el in enumerate(atoms):
        for (lval, exps, coefs) in table[el]:
            nprim = len(exps)
            ncomp = len(_KX[lval])
            bas_coeff += list(coefs) * ncomp
            bas_exp += list(exps) * ncomp
            bas_l += [lval] * nprim * ncomp
            nshells[iat] += nprim * ncomp
            for _ in range(ncomp):
                index_ctr += [iao] * nprim
                iao += 1
            for k in _KX[lval]:
                bas_kx += [k] * nprim
            for k in _KY[lval]:
                bas_ky += [k] * nprim
            for k in _KZ[lval]:
                bas_kz += [k] * nprim
    b.nao = iao
    b.nshells = nshells
    b.index_ctr = index_ctr
    edges = np.concatenate(([0], np.cumsum(nshells)))
    b.nao_per_atom = [
        len(np.unique(index_ctr[edges[i]: edges[i + 1]])) for i in range(len(atoms))
    ]
    b.nctr_per_ao = np.bincount(np.asarray(index_ctr), minlength=iao)
    b.bas_coeffs = np.array(bas_coeff, dtype=np.float64)
    b.bas_exp = np.array(bas_exp, dtype=np.float64)
    b.bas_l = bas_l
    b.bas_kr = np.zeros_like(b.bas_exp)
    b.bas_kx = np.array(bas_kx)
    b.bas_ky = np.array(bas_ky)
    b.bas_kz = np.array(bas_kz)
    b.bas_n = [l + 1 for l in bas_l]
    b.atom_coords_internal = [list(c) for c in atom_coords]
    b.TotalEnergy = 0.0
    return b


def build_basis_uncontracted(atoms, atom_coords, basis_name, radial_type):
    """ADF-style namespace: every AO is a single Slater/Gaussian primitive x^kx y^ky z^kz r^kr."""
    table = _UNCONTRACTED[basis_name.lower()]
    b = SimpleNamespace()
    b.radial_type = radial_type
    b.harmonics_type = "cart"
    rows, nshells = [], []
    for el in atoms:
        rows += table[el]
        nshells.append(len(table[el]))
    n = len(rows)
    b.nao = n
    b.nshells = nshells
    b.nao_per_atom = nshells
    b.index_ctr = np.arange(n)
    b.nctr_per_ao = np.ones(n)
    b.bas_kx = np.array([r[0] for r in rows])
    b.bas_ky = np.array([r[1] for r in rows])
    b.bas_kz = np.array([r[2] for r in rows])
    b.bas_kr = np.array([0 if radial_type.endswith("pure") else r[3] for r in rows])
    b.bas_exp = np.array([r[4] for r in rows], dtype=np.float64)
    b.bas_coeffs = np.ones(n)
    b.bas_n = (b.bas_kx + b.bas_ky + b.bas_kz + b.bas_kr + 1).tolist()
    b.atom_coords_internal = [list(c) for c in atom_coords]
    b.TotalEnergy = 0.0
    return b


def build_basis_spherical(atoms, atom_coords, basis_name, radial_type):
    """Namespace of a spherical-harmonics basis (the fields atomic_orbitals.py:27-94 reads for
    harmonics_type == "sph": bas_n, bas_l, bas_m instead of the cartesian powers)."""
    table = _SPHERICAL[basis_name.lower()]
    b = SimpleNamespace()
    b.radial_type = radial_type
    b.harmonics_type = "sph"
    n, lq, mq, zeta, nshells = [], [], [], [], []
    for el in atoms:
        cnt = 0
        for (ni, li, zi) in table[el]:
            for m in range(-li, li + 1):
                n.append(ni); lq.append(li); mq.append(m); zeta.append(zi)
                cnt += 1
        nshells.append(cnt)
    nb = len(n)
    b.nao = nb
    b.nshells = nshells
    b.nao_per_atom = nshells
    b.index_ctr = np.arange(nb)
    b.nctr_per_ao = np.ones(nb)
    import torch
    # (torch tensors: the reference's norm_slater_spherical mixes them with tensors, norm_orbital.py:61-67)
    b.bas_n = torch.tensor(n, dtype=torch.int64)
    b.bas_l = np.array(lq)
    b.bas_m = np.array(mq)
    b.bas_exp = torch.tensor(zeta, dtype=torch.float64)
    b.bas_coeffs = np.ones(nb)
    b.atom_coords_internal = [list(c) for c in atom_coords]
    b.TotalEnergy = 0.0
    return b


def _seeded_mos(n, seed):
    """Deterministic, well-conditioned, diagonally dominant orthonormal matrix (test fixtures only)."""
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(np.eye(n) + 0.35 * rng.standard_normal((n, n)))
    return q * np.sign(np.diag(q))


class Molecule:
    """Duck-typed stand-in for ``qmctorch.scf.Molecule`` (``scf/molecule.py:21``).

    Same attribute names (``nelec nup ndown spin natom atoms atom_coords
    atomic_number atomic_nelec hdf5file basis``) and the same ``domain(method)``
    contract (``scf/molecule.py:176-213``).
    """

    def __init__(self, atom=None, basis="sto-3g", unit="bohr", charge=0, spin=0, name=None,
                 calculator="builtin", mos=None, radial_type=None, load=None):
        if load is not None:      # scf/molecule.py:96-100: everything comes from the file
            self.charge, self.spin = charge, spin
            self._load(load)
            return
        self.atoms_str = atom
        self.unit = unit
        self.charge = charge
        self.spin = spin
        self.basis_name = basis
        self.calculator_name = calculator
        self.max_angular = 2
        names, coords = _parse_atoms(atom, unit)
        self.atoms = np.array(names)
        self.atom_coords = coords
        self.atomic_number = [_Z[n] for n in names]
        self.atomic_nelec = [_Z[n] for n in names]
        self.natom = len(names)
        self.nelec = sum(self.atomic_nelec) + charge
        if (self.nelec - spin) % 2 != 0:
            raise ValueError(
                "%d electrons and spin %d doesn't make sense" % (self.nelec, spin))
        self.nup = int((self.nelec - spin) / 2) + spin
        self.ndown = int((self.nelec - spin) / 2)
        self.name = name or "".join(names)
        self.hdf5file = "_".join([self.name, calculator, basis]) + ".hdf5"
        if basis.lower() in _UNCONTRACTED:
            self.basis = build_basis_uncontracted(names, coords, basis, radial_type or "sto")
            if mos is None:
                mos = _seeded_mos(self.basis.nao, 17)
        elif basis.lower() in _SPHERICAL:
            self.basis = build_basis_spherical(names, coords, basis, radial_type or "sto")
            if mos is None:
                mos = _seeded_mos(self.basis.nao, 23)
        else:
            self.basis = build_basis(names, coords, basis)
        if mos is None:
            mos = _load_cached_mos(self.name, basis, self.basis.nao)
        mos = np.asarray(mos, dtype=np.float64)
        self.basis.mos = mos / np.sqrt((mos ** 2).sum(0))
        self.basis.nmo = self.basis.mos.shape[1]

    def _load(self, path):
        """``Molecule(load=...)`` (``scf/molecule.py:394-402`` + ``utils/hdf5_utils.py:27-98``): every
        dataset of the ``molecule`` group becomes an attribute, sub-groups become namespaces.
        ``path`` is an HDF5 file written by QMCTorch (read by ``utils/hdf5_min.py``) or the JSON
        dump of one (``tools/hdf5_to_fixture.py``).  MOs are taken verbatim (no renormalisation)."""
        if path.endswith(".json"):
            with open(path) as f:
                tree = _tree_from_json(json.load(f))
        else:
            from .utils.hdf5_min import read_hdf5
            tree = read_hdf5(path)
            if "molecule" not in tree:
                raise KeyError("no 'molecule' group in %s" % path)
            tree = tree["molecule"]

        def fill(obj, d):
            for k, v in d.items():
                if isinstance(v, dict):
                    ns = SimpleNamespace()
                    fill(ns, v)
                    setattr(obj, k, ns)
                else:
                    setattr(obj, k, v.item() if isinstance(v, np.generic) else v)
        fill(self, tree)
        self.hdf5file = path
        self.atoms = np.array([str(a) for a in self.atoms])
        self.atom_coords = [list(map(float, c)) for c in np.asarray(self.atom_coords)]
        self.atomic_number = [int(z) for z in self.atomic_number]
        self.atomic_nelec = [int(z) for z in self.atomic_nelec]
        b = self.basis
        if b.harmonics_type != "cart":
            raise ValueError("Harmonics type should be cart here (spherical harmonics are not "
                             "supported) but %s was found in %s" % (b.harmonics_type, path))
        b.nshells = [int(x) for x in b.nshells]
        b.nao_per_atom = [int(x) for x in b.nao_per_atom]
        b.bas_n = (np.asarray(b.bas_kx) + b.bas_ky + b.bas_kz + b.bas_kr + 1).tolist()
        b.atom_coords_internal = [list(map(float, c)) for c in np.asarray(b.atom_coords_internal)]
        b.mos = np.asarray(b.mos, dtype=np.float64)

    def domain(self, method):
        d = dict(method=method)
        ac = np.asarray(self.atom_coords)
        if method == "center":
            d["center"] = np.mean(ac, 0)
        elif method == "uniform":
            d["min"] = np.min(ac) - 0.5
            d["max"] = np.max(ac) + 0.5
        elif method == "normal":
            d["mean"] = np.mean(ac, 0)
            d["sigma"] = np.diag(np.std(ac, 0) + 0.25)
        elif method == "atomic":
            d["atom_coords"] = self.atom_coords
            d["atom_num"] = self.atomic_number
            d["atom_nelec"] = self.atomic_nelec
        else:
            raise ValueError("Method to initialize the walkers not recognized")
        return d

    def get_total_energy(self):
        return self.basis.TotalEnergy


def _tree_to_json(tree):
    """Nested dict of numpy arrays / scalars / str -> JSON-serialisable (floats round-trip exactly)."""
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out[k] = _tree_to_json(v)
        elif isinstance(v, str):
            out[k] = v
        else:
            a = np.asarray(v)
            out[k] = {"dtype": "str" if a.dtype.kind == "U" else str(a.dtype), "shape": list(a.shape),
                      "data": a.ravel().tolist()}
    return out


def _tree_from_json(js):
    out = {}
    for k, v in js.items():
        if isinstance(v, str):
            out[k] = v
        elif "dtype" in v and "data" in v:
            a = np.array(v["data"], dtype=None if v["dtype"] == "str" else v["dtype"]).reshape(v["shape"])
            out[k] = a if a.ndim else a[()]
        else:
            out[k] = _tree_from_json(v)
    return out


def _load_cached_mos(name, basis, nao):
    path = os.path.join(_DATA_DIR, "%s_%s_mos.json" % (name, basis.lower()))
    if not os.path.isfile(path):
        raise FileNotFoundError(
            "no cached MO matrix %s; run tools/make_mos.py or pass mos=" % path)
    with open(path) as f:
        mos = np.array(json.load(f)["mos"], dtype=np.float64)
    if mos.shape[0] != nao:
        raise ValueError("cached MO matrix %s does not match the basis" % path)
    return mos


# --- the five BASELINE.json configurations -------------------------------------------

_BUTADIENE = None


def _butadiene_atoms():
    """planar s-trans 1,3-butadiene, angstrom (C=C 1.34, C-C 1.46, C-H 1.09, 123 deg)."""
    global _BUTADIENE
    if _BUTADIENE is None:
        import math
        a = math.radians(123.0)
        c2 = np.array([0.73, 0.0, 0.0])
        c3 = -c2
        d12 = np.array([-math.cos(a), math.sin(a), 0.0])  # direction C2->C1
        c1 = c2 + 1.34 * d12
        c4 = -c1

        def rot(v, ang):
            c, s = math.cos(ang), math.sin(ang)
            return np.array([c * v[0] - s * v[1], s * v[0] + c * v[1], 0.0])
        u23 = np.array([-1.0, 0.0, 0.0])
        h2 = c2 + 1.09 * rot(u23, -math.radians(119.0))
        u21 = -d12
        h1a = c1 + 1.09 * rot(u21, math.radians(121.0))
        h1b = c1 + 1.09 * rot(u21, -math.radians(121.0))
        pts = [("C", c1), ("C", c2), ("C", c3), ("C", c4), ("H", h1a), ("H", h1b),
               ("H", h2), ("H", -h2), ("H", -h1a), ("H", -h1b)]
        _BUTADIENE = "; ".join("%s %.8f %.8f %.8f" % (n, p[0], p[1], p[2]) for n, p in pts)
    return _BUTADIENE


_SPECS = {
    "h2": dict(atom="H 0 0 -0.69; H 0 0 0.69", unit="bohr", basis="sto-3g", name="H2"),
    "lih_sto3g": dict(atom="Li 0 0 0; H 0 0 3.015", unit="bohr", basis="sto-3g", name="LiH"),
    "lih": dict(atom="Li 0 0 0; H 0 0 3.015", unit="bohr", basis="6-31g", name="LiH"),
    # ADF-style uncontracted Slater / Gaussian-with-r^n bases (radial_slater, radial_slater_pure,
    # radial_gaussian of radial_functions.py); seeded orthonormal MOs
    "lih_sto": dict(atom="Li 0 0 0; H 0 0 3.015", unit="bohr", basis="sto-dz", name="LiH", radial_type="sto"),
    "lih_sto_pure": dict(atom="Li 0 0 0; H 0 0 3.015", unit="bohr", basis="sto-dz", name="LiH",
                         radial_type="sto_pure"),
    "lih_gto_kr": dict(atom="Li 0 0 0; H 0 0 3.015", unit="bohr", basis="sto-dz", name="LiH", radial_type="gto"),
    # real spherical harmonics up to l = 2 (d shell on Li), Slater and Gaussian radial parts with r^n
    "lih_sph": dict(atom="Li 0 0 0; H 0.3 -0.2 3.015", unit="bohr", basis="sph-dz", name="LiH", radial_type="sto"),
    "lih_sph_gto": dict(atom="Li 0 0 0; H 0.3 -0.2 3.015", unit="bohr", basis="sph-dz", name="LiH", radial_type="gto"),
    "h2o": dict(atom="O 0 0 0; H 0.757 0.587 0; H -0.757 0.587 0", unit="angs",
                basis="cc-pvdz", name="H2O"),
    "c4h6": dict(atom=None, unit="angs", basis="dzp", name="C4H6"),
    "c4h6_dz": dict(atom=None, unit="angs", basis="dz", name="C4H6"),
}

# Test inputs that are dumps of the reference's own files (tests/data/*_adf_*.json, made by
# tools/hdf5_to_fixture.py from the reference's tests/hdf5/*.hdf5) live with the tests, not in the
# package: keys lih_adf | h2_adf | co2_adf resolve there when that directory exists (a source checkout).
_TEST_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data")
for _k, _f in (("lih_adf", "LiH_adf_dz.json"), ("h2_adf", "H2_adf_dzp.json"), ("co2_adf", "CO2_adf_dzp.json")):
    if os.path.isfile(os.path.join(_TEST_DATA, _f)):
        _SPECS[_k] = dict(load=os.path.join(_TEST_DATA, _f))


def fixture_spec(key):
    spec = dict(_SPECS[key])
    if key.startswith("c4h6"):
        spec["atom"] = _butadiene_atoms()
    return spec


def fixture_molecule(key, mos=None):
    """``h2`` | ``lih_sto3g`` | ``lih`` (6-31G, BASELINE config 2/3) | ``h2o`` | ``c4h6``."""
    return Molecule(mos=mos, **fixture_spec(key))
