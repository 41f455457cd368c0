"""CPU oracle: a restatement of QMCTorch's Slater-Jastrow hot path (torch, CPU, FP64).

TEST INFRASTRUCTURE ONLY - this file is the *checker*.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; nothing under ``qmctorch_b200/`` does, and the product
path raises if its CUDA library is missing instead of falling back to this.

Why torch and not numpy: the reference's results depend on ATen's rounding in two
places - the Gram-form electron-electron distance (``torch.bmm`` with K=3,
``electron_electron_distance.py:177-190``) and the LU behind ``torch.det`` /
``torch.inverse`` (``slater_pooling.py:106-107,847-848``).  Using the same library
calls on the same host keeps the oracle bit-identical to the reference there, so
the 1e-10 parity bar is tested against the reference's own rounding.  Everything
else is elementwise arithmetic written out from the formulas.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(``oracle/ref_shim.py``) on the fixture molecules in the build container and
(a) asserts this file reproduces it to <= 1e-13 relative (bit-identical for
ground-state single determinants), (b) writes ``tests/golden/*.npz`` which
``tests/test_oracle_golden.py`` re-checks wherever the reference is absent.

Each function cites the reference lines (relative to ``qmctorch/``) it restates.
Shapes: W walkers, Ne electrons, Nb flat primitives, Na AOs, Nm MOs, Nc configs.
"""

import math
from types import SimpleNamespace

import numpy as np
import torch

F64 = torch.float64


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------

def _dfact(n):
    n = int(n)
    return 1.0 if n <= 0 else float(np.prod(np.arange(n, 0, -2, dtype=np.float64)))


def primitive_norms(basis):
    """wavefunction/orbitals/norm_orbital.py:8-161 (cartesian harmonics only)."""
    a = np.asarray(basis.bas_kx, dtype=np.float64)
    b = np.asarray(basis.bas_ky, dtype=np.float64)
    c = np.asarray(basis.bas_kz, dtype=np.float64)
    ex = np.asarray(basis.bas_exp, dtype=np.float64)
    if basis.harmonics_type != "cart":
        raise NotImplementedError("oracle covers cartesian harmonics")
    if basis.radial_type.startswith("gto"):
        out = (2 * ex / np.pi) ** 0.75
        for k in (a, b, c):
            df = np.array([_dfact(2 * ki - 1) for ki in k])
            out = out * ((4 * ex) ** (k / 2) / np.sqrt(df))
        return out
    if basis.radial_type.startswith("sto"):
        n = np.asarray(basis.bas_kr, dtype=np.float64)
        lv = a + b + c + n + 1.0
        lfact = np.array([float(math.factorial(int(2 * i))) for i in lv])
        pref = 4 * np.pi * lfact / ((2 * ex) ** (2 * lv + 1))
        num = np.array([_dfact(2 * i - 1) * _dfact(2 * j - 1) * _dfact(2 * k - 1)
                        for i, j, k in zip(a, b, c)])
        den = np.array([_dfact(2 * (i + j + k) + 1) for i, j, k in zip(a, b, c)])
        return np.sqrt(1.0 / (pref * num / den))
    raise ValueError("radial_type")


# Real spherical harmonics of the reference, l <= 2 (orbitals/spherical_harmonics.py:352-702; Y00 is the
# truncated literal 0.2820948 of :363), written as sums of cartesian monomials over r^l:
#   (l, m) -> [(coefficient, kx, ky, kz)]
_SPH_MONOMIALS = {
    (0, 0): [(0.2820948, 0, 0, 0)],
    (1, -1): [(0.4886025119029199, 0, 1, 0)], (1, 0): [(0.4886025119029199, 0, 0, 1)],
    (1, 1): [(0.4886025119029199, 1, 0, 0)],
    (2, -2): [(1.0925484305920792, 1, 1, 0)], (2, -1): [(1.0925484305920792, 0, 1, 1)],
    (2, 1): [(1.0925484305920792, 1, 0, 1)],
    (2, 2): [(0.5462742152960396, 2, 0, 0), (-0.5462742152960396, 0, 2, 0)],
    (2, 0): [(-0.31539156525252005, 2, 0, 0), (-0.31539156525252005, 0, 2, 0), (2 * 0.31539156525252005, 0, 0, 2)],
}


def spherical_as_cartesian(b):
    """A spherical-harmonics basis (harmonics_type "sph": AO = N(n, alpha) r^n e^{-alpha r | r^2} Y_lm,
    atomic_orbitals.py:56-64, radial_functions.py:6-238, norm_orbital.py:45-93) restated as the contracted
    cartesian basis the rest of this oracle evaluates: r^n Y_lm = sum_t c_t x^a y^b z^c r^(n-l).  Returns
    (cartesian namespace, per-primitive norm with the monomial coefficient folded in)."""
    n = np.asarray(b.bas_n).astype(int)
    lq, mq = np.asarray(b.bas_l).astype(int), np.asarray(b.bas_m).astype(int)
    ex = np.asarray(b.bas_exp, dtype=np.float64)
    if b.radial_type.startswith("sto"):
        nfact = np.array([float(math.factorial(int(2 * k))) for k in n])
        norm = (2 * ex) ** n * np.sqrt(2 * ex / nfact)
    else:
        n1 = n + 1.0
        norm = np.sqrt(2 ** (2 * n1 + 1.5) / (np.array([_dfact(2 * int(k) - 1) for k in n1]) * np.pi ** 0.5)) \
            * ex ** (0.25 * (2 * n1 + 1))
    atom = np.repeat(np.arange(len(b.nshells)), np.asarray(b.nshells))
    rows = []
    for i in range(len(n)):
        for c, a, bb, cc in _SPH_MONOMIALS[(int(lq[i]), int(mq[i]))]:
            rows.append((i, c, a, bb, cc, n[i] - lq[i]))
    ix = np.array([r[0] for r in rows])
    out = SimpleNamespace(**vars(b))
    out.harmonics_type = "cart"
    out.nshells = [int((atom[ix] == a).sum()) for a in range(len(b.nshells))]
    out.index_ctr = np.asarray(b.index_ctr)[ix]
    out.bas_coeffs = np.asarray(b.bas_coeffs, dtype=np.float64)[ix]
    out.bas_exp = ex[ix]
    out.bas_kx = np.array([r[2] for r in rows]); out.bas_ky = np.array([r[3] for r in rows])
    out.bas_kz = np.array([r[4] for r in rows]); out.bas_kr = np.array([r[5] for r in rows])
    return out, norm[ix] * np.array([r[1] for r in rows]), ix


def make_params(mol, configs, jastrow_weight=1.0, en_weight=None):
    """Collect every tensor the path reads.  Trainable leaves are plain tensors."""
    b = mol.basis
    sph_norm = None
    if b.harmonics_type == "sph":
        b, sph_norm, _ = spherical_as_cartesian(b)
    p = SimpleNamespace()
    p.nelec, p.nup, p.ndown = mol.nelec, mol.nup, mol.ndown
    p.atom_coords = torch.tensor(np.asarray(b.atom_coords_internal), dtype=F64)
    p.atomic_number = [float(z) for z in mol.atomic_number]
    p.nshells = torch.as_tensor(np.asarray(b.nshells))
    p.index_ctr = torch.as_tensor(np.asarray(b.index_ctr)).long()
    p.contract = len(np.unique(np.asarray(b.index_ctr))) != len(b.index_ctr)
    p.nao = int(b.nao)
    p.bas_coeffs = torch.tensor(np.asarray(b.bas_coeffs), dtype=F64)
    p.bas_exp = torch.tensor(np.asarray(b.bas_exp), dtype=F64)
    p.bas_n = torch.tensor(np.asarray(b.bas_kr), dtype=F64)
    p.bas_k = torch.tensor(np.stack([b.bas_kx, b.bas_ky, b.bas_kz], 1)).long()
    p.radial_type = b.radial_type
    p.norm = torch.tensor(primitive_norms(b) if sph_norm is None else sph_norm, dtype=F64)
    if sph_norm is not None:
        p.contract = True      # the monomials of an AO are summed by the contraction (index_add over index_ctr)
    p.mo_scf = torch.tensor(np.asarray(b.mos), dtype=F64)
    p.mo_modifier = torch.ones_like(p.mo_scf)
    p.configs = (torch.as_tensor(configs[0]).long(), torch.as_tensor(configs[1]).long())
    nci = p.configs[0].shape[0]
    p.ci = torch.zeros(1, nci, dtype=F64)
    p.ci[0, 0] = 1.0
    p.jastrow_weight = None if jastrow_weight is None else torch.tensor([jastrow_weight], dtype=F64)
    p.en_weight = None if en_weight is None else torch.tensor([en_weight], dtype=F64)
    # three-body Boys-Handy term: dict(num [1,2,nterm], denom [1,2,nterm], fc [1,nterm]) or None
    p.een = None
    return p


# --------------------------------------------------------------------------------------
# a1-a4: atomic orbitals
# --------------------------------------------------------------------------------------

def _ipow(x, k):
    """x**k for integer tensors k, following utils/torch_utils.py:24-59."""
    if int(k.max()) < 3:
        out = x.clone()
        out.masked_fill_(k == 0, 1.0)
        if int(k.max()) > 1:
            m2 = k == 2
            out[..., m2] = out[..., m2] * out[..., m2]
        return out
    return x ** k


def ao_all(p, pos):
    """AO value, gradient, Laplacian.  orbitals/atomic_orbitals.py:578-669,
    radial_functions.py:6-406, spherical_harmonics.py:102-199.

    Returns ao [W,Ne,Na], dao [W,Ne,Na,3], d2ao [W,Ne,Na]."""
    W = pos.shape[0]
    Ne = pos.shape[1] // 3
    xyz_at = pos.view(W, Ne, 1, 3) - p.atom_coords[None, None]
    r_at = torch.sqrt((xyz_at * xyz_at).sum(3))
    xyz = xyz_at.repeat_interleave(p.nshells, dim=2)          # [W,Ne,Nb,3]
    r = r_at.repeat_interleave(p.nshells, dim=2)              # [W,Ne,Nb]
    al, n = p.bas_exp, p.bas_n
    rt = p.radial_type
    # radial part and derivatives
    if rt in ("gto", "gto_pure"):
        r2 = r * r
        er = torch.exp(-al * r2)
        aer = al * er
        d_er = -2 * aer.unsqueeze(-1) * xyz
        lap_er = al * er * (4 * al * r2 - 6)
    else:
        er = torch.exp(-al * r)
        aer = al * er
        d_er = -aer.unsqueeze(-1) * xyz / r.unsqueeze(-1)
        lap_er = aer * (al - 2.0 / r)
    if rt in ("gto", "sto"):
        rn = _ipow(r, n.long()) if float(n.max()) < 3 else r ** n
        nrnm2 = n * r ** (n - 2)
        d_rn = nrnm2.unsqueeze(-1) * xyz
        lap_rn = nrnm2 * (n + 1)
        R = rn * er
        dR = d_rn * er.unsqueeze(-1) + rn.unsqueeze(-1) * d_er
        d2R = lap_rn * er + 2 * (d_rn * d_er).sum(3) + rn * lap_er
    else:
        R, dR, d2R = er, d_er, lap_er
    # cartesian harmonics
    k = p.bas_k
    xk = _ipow(xyz, k)
    Y = xk.prod(-1)
    km1 = (k - 1).clamp(min=0)
    km2 = (k - 2).clamp(min=0)
    xkm1 = _ipow(xyz, km1)
    xkm2 = _ipow(xyz, km2)
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    dY = torch.stack((kx * xkm1[..., 0] * xk[..., 1] * xk[..., 2],
                      ky * xk[..., 0] * xkm1[..., 1] * xk[..., 2],
                      kz * xk[..., 0] * xk[..., 1] * xkm1[..., 2]), dim=-1)
    d2Y = (kx * (kx - 1) * xkm2[..., 0] * xk[..., 1] * xk[..., 2]
           + ky * (ky - 1) * xk[..., 0] * xkm2[..., 1] * xk[..., 2]
           + kz * (kz - 1) * xk[..., 0] * xk[..., 1] * xkm2[..., 2])
    # products + contraction  (atomic_orbitals.py:236-249,330-356,397-423,654-669)
    v = p.norm * R * Y
    g = dR * Y.unsqueeze(-1) + R.unsqueeze(-1) * dY
    g = p.norm.unsqueeze(-1) * p.bas_coeffs.unsqueeze(-1) * g
    l = p.norm * (d2R * Y + 2.0 * (dR * dY).sum(3) + R * d2Y)
    if p.contract:
        ao = torch.zeros(W, Ne, p.nao, dtype=F64).index_add_(2, p.index_ctr, p.bas_coeffs * v)
        dao = torch.zeros(W, Ne, p.nao, 3, dtype=F64).index_add_(2, p.index_ctr, g)
        d2ao = torch.zeros(W, Ne, p.nao, dtype=F64).index_add_(2, p.index_ctr, p.bas_coeffs * l)
    else:
        ao, dao, d2ao = v, g, l
    return ao, dao, d2ao


def ao_values(p, pos):
    """AO values only (atomic_orbitals.py:221-249), same arithmetic as ao_all."""
    W = pos.shape[0]
    Ne = pos.shape[1] // 3
    xyz_at = pos.view(W, Ne, 1, 3) - p.atom_coords[None, None]
    r_at = torch.sqrt((xyz_at * xyz_at).sum(3))
    xyz = xyz_at.repeat_interleave(p.nshells, dim=2)
    r = r_at.repeat_interleave(p.nshells, dim=2)
    al, n = p.bas_exp, p.bas_n
    if p.radial_type in ("gto", "gto_pure"):
        er = torch.exp(-al * (r * r))
    else:
        er = torch.exp(-al * r)
    if p.radial_type in ("gto", "sto"):
        er = (_ipow(r, n.long()) if float(n.max()) < 3 else r ** n) * er
    Y = _ipow(xyz, p.bas_k).prod(-1)
    v = p.norm * er * Y
    if p.contract:
        return torch.zeros(W, Ne, p.nao, dtype=F64).index_add_(2, p.index_ctr, p.bas_coeffs * v)
    return v


# --------------------------------------------------------------------------------------
# a5: molecular orbitals
# --------------------------------------------------------------------------------------

def mo_weight(p):
    """orbitals/molecular_orbitals.py:91."""
    return p.mo_scf * p.mo_modifier


def ao2mo(p, ao):
    """orbitals/molecular_orbitals.py:79-95."""
    w = mo_weight(p)
    return ao @ w.reshape(1, *w.shape)


# --------------------------------------------------------------------------------------
# a6-a8: distances and Pade Jastrow factors
# --------------------------------------------------------------------------------------

def _tri_up(Ne):
    idx = [(i, j) for i in range(Ne - 1) for j in range(i + 1, Ne)]
    return (torch.tensor([i for i, _ in idx]).long(), torch.tensor([j for _, j in idx]).long())


def ee_distance_gram(pos3):
    """r_ij through the Gram expansion, jastrows/distance/electron_electron_distance.py:47-190
    (get_distance_quadratic :177-190, safe_sqrt :104-110), eps = 1e-16 in FP64."""
    norm = (pos3 ** 2).sum(-1).unsqueeze(-1)
    d2 = norm + norm.transpose(1, 2) - 2.0 * torch.bmm(pos3, pos3.transpose(1, 2))
    Ne = pos3.shape[1]
    eye = torch.eye(Ne, dtype=F64)
    diag = torch.diag_embed(torch.diagonal(d2, dim1=-1, dim2=-2))
    return torch.sqrt(d2 - diag + 1e-16 * eye)


def ee_static_weight(nup, ndown):
    """elec_elec/kernels/pade_jastrow_kernel.py:34-66: 0.25 same spin block, 0.5 otherwise."""
    Ne = nup + ndown
    w = torch.full((Ne, Ne), 0.5, dtype=F64)
    w[:nup, :nup] = 0.25
    w[nup:, nup:] = 0.25
    row, col = _tri_up(Ne)
    return w[row, col]


def jastrow_ee(p, pos, derivative=True):
    """Pade electron-electron Jastrow.  jastrow_factor_electron_electron.py:124-260,
    kernels/pade_jastrow_kernel.py:68-153.
    Returns J [W,1] (, dJ [W,3,Ne], d2J [W,Ne])."""
    W = pos.shape[0]
    Ne = p.nelec
    w = p.jastrow_weight
    pos3 = pos.view(W, Ne, 3)
    row, col = _tri_up(Ne)
    rfull = ee_distance_gram(pos3)
    r = rfull[:, row, col]                                   # [W,Np]
    w0 = ee_static_weight(p.nup, p.ndown)
    kern = w0 * r / (1.0 + w * r)
    J = torch.exp(kern.sum(-1)).unsqueeze(-1)
    if not derivative:
        return J
    diff = pos3[:, row, :] - pos3[:, col, :]                  # [W,Np,3]
    dr = (diff / r.unsqueeze(-1)).transpose(1, 2)             # [W,3,Np]
    sq = diff ** 2
    d2r = torch.stack((sq[..., 1] + sq[..., 2], sq[..., 2] + sq[..., 0],
                       sq[..., 0] + sq[..., 1]), dim=1) / (r ** 3).unsqueeze(1)
    r_ = r.unsqueeze(1)
    den = 1.0 / (1.0 + w * r_)
    dk = w0 * dr * den + (-w0 * w * r_ * dr * den ** 2)
    den2 = den ** 2
    dr2 = dr * dr
    d2k = (w0 * d2r * den + (-2 * w0 * w * dr2 * den2)
           + (-w0 * w * r_ * d2r * den2) + 2 * w0 * w ** 2 * r_ * dr2 * den ** 3)
    dJ = torch.zeros(W, 3, Ne, dtype=F64)
    dJ.index_add_(-1, row, dk * J.unsqueeze(-1))
    dJ.index_add_(-1, col, -(dk * J.unsqueeze(-1)))
    d2s = d2k.sum(-2)
    h = torch.zeros(W, Ne, dtype=F64)
    h.index_add_(-1, row, d2s)
    h.index_add_(-1, col, d2s)
    g = torch.zeros(W, 3, Ne, dtype=F64)
    g.index_add_(-1, row, dk)
    g.index_add_(-1, col, -dk)
    d2J = (h + (g ** 2).sum(-2)) * J
    return J, dJ, d2J


def _jastrow_atoms(p):
    """The Jastrow factors keep their OWN copy of the atom positions, a plain tensor outside the autograd graph
    (elec_nuclei/jastrow_factor_electron_nuclei.py:40-41, elec_elec_nuclei/...:46-47): derivatives w.r.t.
    ao.atom_coords (forces) do not pass through them."""
    return p.atom_coords.detach()


def jastrow_en(p, pos, derivative=True):
    """Pade electron-nucleus Jastrow.  elec_nuclei/jastrow_factor_electron_nuclei.py:60-161,
    elec_nuclei/kernels/pade_jastrow_kernel.py:36-116, distance/electron_nuclei_distance.py:53-162."""
    W = pos.shape[0]
    Ne = p.nelec
    w = p.en_weight
    pos3 = pos.view(W, Ne, 3)
    atoms = _jastrow_atoms(p)
    diff = pos3.unsqueeze(2) - atoms[None, None]              # [W,Ne,Nat,3]
    # Gram-form distance like the reference (electron_nuclei_distance.py:153-162)
    nrm = (pos3 ** 2).sum(-1).unsqueeze(-1)
    nrm_at = (atoms ** 2).sum(-1).unsqueeze(-1).T
    r = torch.sqrt(nrm + nrm_at - 2.0 * pos3 @ atoms.T)       # [W,Ne,Nat]
    kern = r / (1.0 + w * r)                                  # w0 = 1
    J = torch.exp(kern.sum((-1, -2))).unsqueeze(-1)
    if not derivative:
        return J
    eps = 1e-16
    invr = 1.0 / (r + eps)
    dr = (diff * invr.unsqueeze(-1)).permute(0, 3, 1, 2)      # [W,3,Ne,Nat]
    invr3 = 1.0 / (r ** 3 + eps)
    sq = diff ** 2
    d2r = torch.stack((sq[..., 1] + sq[..., 2], sq[..., 2] + sq[..., 0],
                       sq[..., 0] + sq[..., 1]), dim=1) * invr3.unsqueeze(1)
    r_ = r.unsqueeze(1)
    den = 1.0 / (1.0 + w * r_)
    dk = dr * den + (-w * r_ * dr * den ** 2)
    den2 = den ** 2
    dr2 = dr * dr
    d2k = (d2r * den + (-2 * w * dr2 * den2) + (-w * r_ * d2r * den2)
           + 2 * w ** 2 * r_ * dr2 * den ** 3)
    g = dk.sum(-1)                                            # [W,3,Ne]
    dJ = g * J.unsqueeze(-1)
    d2J = (d2k.sum(-1).sum(1) + (g ** 2).sum(1)) * J
    return J, dJ, d2J


def _een_kernel(p, r):
    """Boys-Handy kernel, elec_elec_nuclei/kernels/boys_handy_jastrow_kernel.py:50-93 with exponents 1:
    r [..., 3] = (r_iA, r_jA, r_ij) -> sum_mu c_mu f_mu(r_iA) f_mu(r_jA) g_mu(r_ij)."""
    num, den, fc = p.een["num"], p.een["denom"], p.een["fc"]
    rep = torch.tensor([2, 1])
    shp = list(r.shape)[:-1] + [1]
    x = r.reshape(-1, 3, 1)
    wn = num.repeat_interleave(rep, dim=1)
    wd = den.repeat_interleave(rep, dim=1)
    x = (wn * x) / (1.0 + wd * x)
    x = x ** torch.ones(3, num.shape[-1], dtype=F64)
    x = x.prod(1)
    return (x @ fc.t()).reshape(*shp)


def _een_logj(p, pos):
    """sum_A sum_{i<j} K, distances assembled like jastrow_factor_electron_electron_nuclei.py:126-143
    (Gram-form r_ij and r_iA)."""
    W = pos.shape[0]
    Ne = p.nelec
    pos3 = pos.view(W, Ne, 3)
    row, col = _tri_up(Ne)
    ree = ee_distance_gram(pos3)[:, row, col]                          # [W,Np]
    atoms = _jastrow_atoms(p)
    nrm = (pos3 ** 2).sum(-1).unsqueeze(-1)
    nrm_at = (atoms ** 2).sum(-1).unsqueeze(-1).T
    ren = torch.sqrt(nrm + nrm_at - 2.0 * pos3 @ atoms.T)              # [W,Ne,Nat]
    nat = atoms.shape[0]
    r = torch.stack((ren[:, row, :].transpose(1, 2), ren[:, col, :].transpose(1, 2),
                     ree.unsqueeze(1).expand(W, nat, ree.shape[1])), dim=-1)   # [W,Nat,Np,3]
    return _een_kernel(p, r).reshape(W, -1).sum(-1)


def jastrow_een(p, pos, derivative=True):
    """Three-body factor exp(sum K) with gradient / per-electron Laplacian by autograd, which is how
    the reference obtains them (jastrow_factor_electron_electron_nuclei.py:188-251,385-439)."""
    if not derivative:
        return torch.exp(_een_logj(p, pos)).unsqueeze(-1)
    W = pos.shape[0]
    Ne = p.nelec
    with torch.enable_grad():
        x = pos.detach().clone().requires_grad_(True)
        J = torch.exp(_een_logj(p, x))
        (jac,) = torch.autograd.grad(J.sum(), x, create_graph=True)
        hess = torch.zeros_like(jac)
        for i in range(x.shape[1]):
            (h,) = torch.autograd.grad(jac[:, i].sum(), x, retain_graph=True)
            hess[:, i] = h[:, i]
    dJ = jac.detach().view(W, Ne, 3).permute(0, 2, 1).contiguous()      # [W,3,Ne]
    d2J = torch.nan_to_num(hess.detach().view(W, Ne, 3).sum(2), nan=0.0)
    return J.detach().unsqueeze(-1), dJ, d2J


def jastrow_all(p, pos):
    """Product of the active Jastrow terms, jastrows/combine_jastrow.py:33-195.
    Returns J [W,1], dJ [W,3,Ne], d2J [W,Ne]; None when no Jastrow is configured."""
    terms = []
    if p.jastrow_weight is not None:
        terms.append(jastrow_ee(p, pos))
    if p.en_weight is not None:
        terms.append(jastrow_en(p, pos))
    if getattr(p, "een", None) is not None:
        terms.append(jastrow_een(p, pos))
    if not terms:
        return None
    J, dJ, d2J = terms[0]
    for (Jb, dJb, d2Jb) in terms[1:]:      # product rule, pairwise
        d2J = d2J * Jb + d2Jb * J + 2 * (dJ * dJb).sum(1)
        dJ = dJ * Jb.unsqueeze(-1) + dJb * J.unsqueeze(-1)
        J = J * Jb
    return J, dJ, d2J


def jastrow_value(p, pos):
    J = None
    if p.jastrow_weight is not None:
        J = jastrow_ee(p, pos, derivative=False)
    if p.en_weight is not None:
        Jn = jastrow_en(p, pos, derivative=False)
        J = Jn if J is None else J * Jn
    if getattr(p, "een", None) is not None:
        Jt = jastrow_een(p, pos, derivative=False)
        J = Jt if J is None else J * Jt
    return J


# --------------------------------------------------------------------------------------
# a11-a14: Slater determinants, trace trick, psi, kinetic energy
# --------------------------------------------------------------------------------------

def _spin_blocks(p, mat, cup, cdown):
    return mat[..., : p.nup, :][..., cup], mat[..., p.nup:, :][..., cdown]


def slater_dets(p, mo):
    """D_up * D_down per configuration (explicit route), pooling/slater_pooling.py:96-111."""
    out = []
    for cup, cdown in zip(*p.configs):
        au, ad = _spin_blocks(p, mo, cup, cdown)
        out.append(torch.det(au) * torch.det(ad))
    return torch.stack(out, dim=-1)                           # [W,Nc]


def slater_trace(p, mo, bop):
    """Tr(Aup^-1 Bup) + Tr(Adown^-1 Bdown) per configuration, slater_pooling.py:348-387.
    bop may carry a leading operator dimension."""
    out = []
    for cup, cdown in zip(*p.configs):
        au, ad = _spin_blocks(p, mo, cup, cdown)
        bu, bd = _spin_blocks(p, bop, cup, cdown)
        tu = torch.diagonal(torch.inverse(au) @ bu, dim1=-2, dim2=-1).sum(-1)
        td = torch.diagonal(torch.inverse(ad) @ bd, dim1=-2, dim2=-1).sum(-1)
        out.append(tu + td)
    return torch.stack(out, dim=-1)


def psi(p, pos):
    """wavefunction/slater_jastrow.py:243-286."""
    mo = ao2mo(p, ao_values(p, pos))
    dets = slater_dets(p, mo)
    val = dets @ p.ci.t()
    J = jastrow_value(p, pos)
    return val if J is None else J * val


def kinetic_operator(p, pos, ao, dao, d2ao, mo):
    """B_kin, slater_jastrow.py:449-482."""
    bkin = ao2mo(p, d2ao)
    jast = jastrow_all(p, pos)
    if jast is not None:
        J, dJ, d2J = jast
        dJ = dJ.transpose(1, 2) / J.unsqueeze(-1)
        d2J = d2J / J
        dmo = ao2mo(p, dao.transpose(2, 3)).transpose(2, 3)
        bkin = bkin + 2 * (dJ.unsqueeze(2) * dmo).sum(-1) + d2J.unsqueeze(-1) * mo
    return -0.5 * bkin


def kinetic_energy(p, pos):
    """Jacobi-trick kinetic energy, slater_jastrow.py:312-344."""
    ao, dao, d2ao = ao_all(p, pos)
    mo = ao2mo(p, ao)
    bkin = kinetic_operator(p, pos, ao, dao, d2ao, mo)
    kin = slater_trace(p, mo, bkin)
    dets = slater_dets(p, mo)
    return ((kin * dets) @ p.ci.t()) / (dets @ p.ci.t())


# --------------------------------------------------------------------------------------
# a15: potentials and local energy
# --------------------------------------------------------------------------------------

def potentials(p, pos):
    """wavefunction/wf_base.py:49-116 - sequential accumulation in the reference's order."""
    W = pos.shape[0]
    Ne = p.nelec
    vee = torch.zeros(W, dtype=F64)
    for i in range(Ne - 1):
        a = pos[:, 3 * i: 3 * i + 3]
        for j in range(i + 1, Ne):
            b = pos[:, 3 * j: 3 * j + 3]
            vee += 1.0 / torch.sqrt(((a - b) ** 2).sum(1))
    ven = torch.zeros(W, dtype=F64)
    for i in range(Ne):
        a = pos[:, 3 * i: 3 * i + 3]
        for A in range(p.atom_coords.shape[0]):
            ven += -p.atomic_number[A] / torch.sqrt(((a - p.atom_coords[A]) ** 2).sum(1))
    vnn = 0.0
    nat = p.atom_coords.shape[0]
    for A in range(nat - 1):
        for B in range(A + 1, nat):
            rnn = torch.sqrt(((p.atom_coords[A] - p.atom_coords[B]) ** 2).sum())
            vnn = vnn + p.atomic_number[A] * p.atomic_number[B] / rnn
    return ven.view(-1, 1), vee.view(-1, 1), vnn


def local_energy(p, pos):
    """wavefunction/wf_base.py:184-215."""
    ven, vee, vnn = potentials(p, pos)
    return kinetic_energy(p, pos) + ven + vee + vnn


# --------------------------------------------------------------------------------------
# a16: analytic gradient of psi w.r.t. electron coordinates
# --------------------------------------------------------------------------------------

def grad_psi(p, pos, pdf=False):
    """slater_jastrow.py:346-447 -> [W, 3 Ne] (electron-major, xyz fastest)."""
    W = pos.shape[0]
    Ne = p.nelec
    ao, dao, _ = ao_all(p, pos)
    mo = ao2mo(p, ao)
    dmo = ao2mo(p, dao.transpose(2, 3)).transpose(2, 3)       # [W,Ne,Nm,3]
    dmo = dmo.permute(3, 0, 1, 2)                             # [3,W,Ne,Nm]
    eye = torch.eye(Ne, dtype=F64)
    op = dmo.unsqueeze(2) * eye.unsqueeze(-1)                 # [3,W,Ne(sel),Ne,Nm]
    op = op.permute(2, 0, 1, 3, 4).reshape(-1, W, Ne, mo.shape[-1])
    gd = slater_trace(p, mo, op)                              # [3Ne,W,Nc]
    dets = slater_dets(p, mo)
    out = ((gd * dets) @ p.ci.t()).transpose(0, 1).squeeze(-1)
    sig = dets @ p.ci.t()
    J = jastrow_value(p, pos)
    if J is not None:
        _, dJ, _ = jastrow_all(p, pos)
        out = J * out + (dJ.permute(0, 2, 1) * sig.unsqueeze(-1)).reshape(W, -1)
    if pdf:
        out = 2 * out * sig
        if J is not None:
            out = out * J
    return out


# --------------------------------------------------------------------------------------
# a17: Metropolis step
# --------------------------------------------------------------------------------------

def metropolis_step(p, pos, fx, displacement, tau, eps=1e-16):
    """One all-electron move with injected draws, sampler/metropolis.py:134-160,279-298.
    Returns (new_pos, new_fx, accept[bool W], fxn)."""
    xn = pos + displacement
    fxn = (psi(p, xn) ** 2).reshape(-1)
    fxn[fxn == 0.0] = eps
    df = fxn / fx
    df[df > 1] = 1.0
    acc = (df - tau >= 0).reshape(-1)
    new_pos = pos.clone()
    new_fx = fx.clone()
    new_pos[acc, :] = xn[acc, :]
    new_fx[acc] = fxn[acc]
    new_fx[new_fx == 0] = eps
    return new_pos, new_fx, acc, fxn


def proposal_sigma(step_size):
    """covariance passed to MultivariateNormal, sampler/metropolis.py:207-212."""
    return step_size / (2 * math.sqrt(2 * math.log(2.0)))


# --------------------------------------------------------------------------------------
# f2: gradient-driven samplers, restated with the analytic density gradient of a16
# --------------------------------------------------------------------------------------

def _rho_and_grad(p, x):
    return (psi(p, x) ** 2).reshape(-1), grad_psi(p, x, pdf=True)


def generalized_metropolis(p, pos0, nstep, step_size, ntherm=-1, ndecor=1):
    """sampler/generalized_metropolis.py:47-219 with the torch CPU generator called in the same
    order (LongTensor.random_, MultivariateNormal.sample, torch.rand per step).  As coded there:
    proposals start from the INITIAL positions (:127-140), the transition uses the plain norm
    (:185-186), the proposal covariance is sqrt(step_size) I (:165-167)."""
    from torch.distributions import MultivariateNormal
    W, Ne = pos0.shape[0], p.nelec
    if ntherm < 0:
        ntherm = nstep + ntherm
    start = pos0.clone()
    xi = pos0.clone()
    rhoi, g = _rho_and_grad(p, xi)
    drifti = 0.5 * g / rhoi.view(-1, 1)
    rhoi[rhoi == 0] = 1e-16
    kept, idecor = [], 0

    def trans(xf, x0, drift):
        return torch.exp(-0.5 * (xf - x0 - drift * step_size).norm(dim=1) / step_size)
    for istep in range(nstep):
        xf = start.clone().view(W, Ne, 3)
        index = torch.LongTensor(W).random_(0, Ne)
        mv = MultivariateNormal(torch.zeros(3), math.sqrt(step_size) * torch.eye(3))
        d = drifti.view(W, Ne, 3)
        xf[range(W), index, :] += step_size * d[range(W), index, :] + mv.sample((W, 1)).squeeze()
        xf = xf.view(W, 3 * Ne)
        rhof, g = _rho_and_grad(p, xf)
        driftf = 0.5 * g / rhof.view(-1, 1)
        rhof[rhof == 0.0] = 1e-16
        P = (trans(xi, xf, driftf) * rhof) / (trans(xf, xi, drifti) * rhoi).double()
        P[P > 1] = 1.0
        tau = torch.rand(W, dtype=F64)
        acc = (P - tau >= 0).reshape(-1)
        xi[acc, :] = xf[acc, :]
        rhoi[acc] = rhof[acc]
        rhoi[rhoi == 0] = 1e-16
        drifti[acc, :] = driftf[acc, :]
        if istep >= ntherm:
            if idecor % ndecor == 0:
                kept.append(xi.clone())
            idecor += 1
    return torch.cat(kept)


def hamiltonian(p, pos0, nstep, step_size, L, ntherm=-1, ndecor=1):
    """sampler/hamiltonian.py:80-201: leapfrog on U = -log psi^2 with grad U = -grad rho / rho,
    torch.randn momenta and torch.rand accept draws in the reference's order."""
    if ntherm < 0:
        ntherm = nstep + ntherm

    def U(x):
        return -torch.log((psi(p, x) ** 2).reshape(-1))

    def dU(x):
        rho, g = _rho_and_grad(p, x)
        return -g / rho.view(-1, 1)
    q_cur = pos0.clone()
    kept, idecor = [], 0
    for istep in range(nstep):
        q = q_cur.clone()
        mom = torch.randn(q.shape)
        e_init = U(q) + 0.5 * (mom * mom).sum(1)
        mom -= 0.5 * step_size * dU(q)
        for _ in range(L - 1):
            q += step_size * mom
            mom -= step_size * dU(q)
        q += step_size * mom
        mom -= 0.5 * step_size * dU(q)
        mom = -mom
        e_new = U(q) + 0.5 * (mom * mom).sum(1)
        eps = torch.rand(e_new.shape)
        rejected = torch.exp(e_init - e_new) < eps
        q[rejected] = q_cur[rejected]
        q_cur = q
        if istep >= ntherm:
            if idecor % ndecor == 0:
                kept.append(q_cur)
            idecor += 1
    return torch.cat(kept)


# --------------------------------------------------------------------------------------
# a18-a19: psi-weighted parameter gradients, statistics
# --------------------------------------------------------------------------------------

def param_grads(p, pos, eloc=None, names=("jastrow_weight", "mo_modifier", "ci", "bas_exp",
                                          "bas_coeffs")):
    """solver/solver.py:372-431 with clip_loss off: psi.backward(2/N (E_L - <E_L>)/psi).
    Returns dict name -> gradient."""
    if eloc is None:
        with torch.no_grad():
            eloc = local_energy(p, pos)
    leaves = {}
    for nme in names:
        if nme.startswith("een_"):
            if p.een is None:
                continue
            t = p.een[nme[4:]].detach().clone().requires_grad_(True)
            p.een[nme[4:]] = t
            leaves[nme] = t
            continue
        t = getattr(p, nme)
        if t is None:
            continue
        t = t.detach().clone().requires_grad_(True)
        setattr(p, nme, t)
        leaves[nme] = t
    val = psi(p, pos)
    norm = 1.0 / len(val)
    weight = (eloc - torch.mean(eloc)) / val.detach()
    weight = weight * (2.0 * norm)
    val.backward(weight)
    out = {k: v.grad.detach().clone() for k, v in leaves.items()}
    for k, v in leaves.items():
        if k.startswith("een_"):
            p.een[k[4:]] = v.detach()
        else:
            setattr(p, k, v.detach())
    return out, eloc


def local_energy_adjoint(p, pos, w_eloc=None, w_psi=None):
    """sum_w w_eloc d E_L / d theta + w_psi d psi / d theta by autograd through this restatement: what the
    reference's loss.backward() (solver/solver.py:352-370, grad="auto") and compute_forces (:433-519) obtain.
    Leaves: atom_coords (through the AOs and potentials only - the Jastrow factors hold a constant copy,
    _jastrow_atoms), bas_exp, bas_coeffs, mo_modifier, ci and the Pade weights.  Without the three-body term
    (its derivatives are detached here, partially detached in the reference).  Returns dict name -> gradient."""
    assert getattr(p, "een", None) is None
    names = ["atom_coords", "bas_exp", "bas_coeffs", "mo_modifier", "ci"]
    if p.jastrow_weight is not None:
        names.append("jastrow_weight")
    if p.en_weight is not None:
        names.append("en_weight")
    old = {n: getattr(p, n) for n in names}
    for n in names:
        setattr(p, n, old[n].detach().clone().requires_grad_(True))
    total = 0.0
    if w_eloc is not None:
        total = total + (local_energy(p, pos).reshape(-1) * w_eloc).sum()
    if w_psi is not None:
        total = total + (psi(p, pos).reshape(-1) * w_psi).sum()
    gr = torch.autograd.grad(total, [getattr(p, n) for n in names], allow_unused=True)
    for n in names:
        setattr(p, n, old[n])
    ren = {"jastrow_weight": "jee_w", "en_weight": "jen_w"}
    return {ren.get(n, n): v for n, v in zip(names, gr)}


def energy_stats(eloc):
    """wf_base.py:217-229, solver_base.py:371: mean, unbiased variance, sqrt(var/N)."""
    e = eloc.mean()
    v = eloc.var()
    return e, v, torch.sqrt(v / eloc.shape[0])
