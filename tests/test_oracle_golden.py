"""CPU: the oracle (oracle/sj_oracle.py) against the committed golden vectors, which were
produced by the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import _cases as C
import sj_oracle as orc

TOL = 1e-11   # the oracle reproduced the reference to <= 1e-12 when the fixtures were made


@pytest.mark.parametrize("name", C.CASES)
def test_oracle_reproduces_reference(name):
    g = C.load(name)
    mol, P = C.oracle_params(g)
    pos = torch.tensor(g["pos"])
    assert C.rel_err(orc.psi(P, pos), g["psi"]) < TOL
    assert C.rel_err(orc.kinetic_energy(P, pos), g["ekin"]) < TOL
    assert C.rel_err(orc.local_energy(P, pos), g["eloc"]) < TOL
    ns = g["ao"].shape[0]
    ao, dao, d2ao = orc.ao_all(P, pos[:ns])
    assert C.scaled_err(ao, g["ao"]) < TOL
    assert C.scaled_err(dao, g["dao"]) < TOL
    assert C.scaled_err(d2ao, g["d2ao"]) < TOL
    if g["jastrow"] != "None":
        J, dJ, d2J = orc.jastrow_all(P, pos)
        assert C.rel_err(J, g["J"]) < TOL
        assert C.scaled_err(dJ, g["dJ"]) < TOL
        assert C.scaled_err(d2J, g["d2J"]) < TOL
    assert C.scaled_err(orc.grad_psi(P, pos), g["gpsi"]) < TOL
    assert C.scaled_err(orc.grad_psi(P, pos, pdf=True), g["gpdf"]) < TOL


@pytest.mark.parametrize("name", ["h2_single22", "lih_ground", "lih_cas24", "lih_een", "h2o_cas44", "lih_sd22_een3"])
def test_oracle_parameter_gradients(name):
    g = C.load(name)
    mol, P = C.oracle_params(g)
    pos = torch.tensor(g["pos"])
    names = [k[5:] for k in g if k.startswith("grad_")]
    og, _ = orc.param_grads(P, pos, names=tuple(names))
    for k in names:
        if k == "ci" and g["ci"].shape[1] == 1:
            continue   # sum_w (E_L - <E_L>) = 0: pure rounding noise
        ref = torch.tensor(g["grad_" + k])
        err = float((og[k] - ref).abs().max() / max(float(ref.abs().max()), 1e-6))
        assert err < 1e-9, (k, err)


@pytest.mark.parametrize("name", ["h2_single22", "lih_ground", "h2o_cas44"])
def test_oracle_metropolis_decisions_bit_exact(name):
    g = C.load(name)
    mol, P = C.oracle_params(g)
    nstep = g["mh_disp"].shape[0]
    for it in range(nstep):
        cur = torch.tensor(g["mh_pos"][it])
        fx = (orc.psi(P, cur) ** 2).reshape(-1)
        fx[fx == 0] = 1e-16
        npos, nfx, acc, fxn = orc.metropolis_step(P, cur, fx, torch.tensor(g["mh_disp"][it]),
                                                  torch.tensor(g["mh_tau"][it]))
        assert np.array_equal(acc.numpy(), g["mh_acc"][it])
        assert np.array_equal(npos.numpy(), g["mh_pos"][it + 1])
        assert C.rel_err(fxn, g["mh_fxn"][it]) < TOL


def test_oracle_local_energy_matches_autograd_laplacian():
    """The reference's own known-answer procedure (tests/wavefunction/base_test_cases.py:92-105):
    Jacobi kinetic energy == autograd Laplacian."""
    g = C.load("lih_ground")
    mol, P = C.oracle_params(g)
    pos = torch.tensor(g["pos"][:6]).requires_grad_(True)
    val = orc.psi(P, pos)
    (jac,) = torch.autograd.grad(val.sum(), pos, create_graph=True)
    lap = torch.zeros(pos.shape[0], dtype=torch.float64)
    for i in range(pos.shape[1]):
        (h,) = torch.autograd.grad(jac[:, i].sum(), pos, retain_graph=True)
        lap += h[:, i]
    ekin_auto = -0.5 * lap.view(-1, 1) / val.detach()
    assert torch.allclose(ekin_auto, orc.kinetic_energy(P, pos.detach()), rtol=1e-8, atol=1e-9)
    assert torch.allclose(jac.detach(), orc.grad_psi(P, pos.detach()), rtol=1e-9, atol=1e-12)


def test_oracle_antisymmetry():
    """tests/wavefunction/base_test_cases.py:24-57."""
    g = C.load("lih_ground")
    mol, P = C.oracle_params(g)
    pos = torch.tensor(g["pos"][:16])
    swapped = pos.clone()
    swapped[:, 0:3], swapped[:, 3:6] = pos[:, 3:6], pos[:, 0:3]     # two spin-up electrons
    assert torch.allclose(orc.psi(P, pos), -orc.psi(P, swapped), rtol=1e-11, atol=0)


def _gradient_sampler_case():
    import os
    from qmctorch_b200.molecules import fixture_molecule
    f = dict(np.load(os.path.join(C.GOLDEN, "gradient_samplers.npz")))
    g = C.load("lih_ground")
    mol = fixture_molecule("lih")
    P = orc.make_params(mol, (g["cfg_up"], g["cfg_down"]), jastrow_weight=float(f["jw"][0]))
    return f, mol, P


def test_oracle_gradient_samplers_reproduce_reference_chains(double_default):
    """the reference's GeneralizedMetropolis / Hamiltonian chains (autograd drift,
    fixtures from oracle/make_golden.py) are reproduced by the restatements that use the analytic
    density gradient, from the same seed."""
    f, mol, P = _gradient_sampler_case()
    start = torch.tensor(f["start"])
    nstep, ntherm, ndecor = (int(v) for v in f["gm_cfg"])
    torch.manual_seed(int(f["gm_seed"][0]))
    out = orc.generalized_metropolis(P, start.clone(), nstep, float(f["gm_step"][0]), ntherm=ntherm, ndecor=ndecor)
    assert out.shape == f["gm_pos"].shape and C.scaled_err(out, f["gm_pos"]) < 1e-10
    nstep, ntherm, ndecor, L = (int(v) for v in f["hm_cfg"])
    torch.manual_seed(int(f["hm_seed"][0]))
    out = orc.hamiltonian(P, start.clone(), nstep, float(f["hm_step"][0]), L, ntherm=ntherm, ndecor=ndecor)
    assert out.shape == f["hm_pos"].shape and C.scaled_err(out, f["hm_pos"]) < 1e-10


def _sph_case(key):
    """Oracle parameters of a spherical-harmonics fixture of tests/golden/sph.npz."""
    import os
    from qmctorch_b200.molecules import fixture_molecule
    g = np.load(os.path.join(C.GOLDEN, "sph.npz"))
    mol = fixture_molecule(key)
    P = orc.make_params(mol, (g[key + "_cfg_up"], g[key + "_cfg_down"]), jastrow_weight=0.8)
    P.mo_modifier = torch.tensor(g[key + "_mo_modifier"])
    P.ci = torch.tensor(g[key + "_ci"])
    return g, mol, P


@pytest.mark.parametrize("key", ["lih_sph", "lih_sph_gto"])
def test_oracle_spherical_harmonics_reproduce_reference(key):
    """Real spherical harmonics l <= 2 (spherical_harmonics.py:352-702) with Slater / Gaussian radial parts
    r^n e^{...}: the oracle evaluates them as sums of cartesian monomials with radial power n - l and
    reproduces the reference's AO values, psi, E_L (the reference's autograd kinetic energy - its Jacobi
    path cannot run on spherical harmonics), the manual gradient estimator and the accept decisions."""
    g, mol, P = _sph_case(key)
    pos = torch.tensor(g[key + "_pos"])
    assert C.scaled_err(orc.ao_values(P, pos[:8]), g[key + "_ao"]) < TOL
    assert C.rel_err(orc.psi(P, pos), g[key + "_psi"]) < TOL
    el = torch.tensor(g[key + "_eloc"])
    assert float(((orc.local_energy(P, pos) - el).abs() / el.abs().clamp(min=1.0)).max()) < 1e-10
    og, _ = orc.param_grads(P, pos, eloc=el, names=("mo_modifier", "ci", "jastrow_weight"))
    for k, v in og.items():
        ref = torch.tensor(g[key + "_grad_" + k])
        assert float((v - ref).abs().max() / ref.abs().max()) < 1e-10, k
    for it in range(g[key + "_mh_disp"].shape[0]):
        cur = torch.tensor(g[key + "_mh_pos"][it])
        fx = (orc.psi(P, cur) ** 2).reshape(-1)
        npos, nfx, acc, fxn = orc.metropolis_step(P, cur, fx, torch.tensor(g[key + "_mh_disp"][it]),
                                                  torch.tensor(g[key + "_mh_tau"][it]))
        assert np.array_equal(acc.numpy().astype(bool), g[key + "_mh_acc"][it])
        assert np.array_equal(npos.numpy(), g[key + "_mh_pos"][it + 1])


VJP_CASES = ["h2_single22", "lih_cas24", "lih_een", "h2o_cas44", "lih_sto", "lih_gto_kr", "lih_adf_sd22"]


@pytest.mark.parametrize("name", VJP_CASES)
def test_oracle_local_energy_adjoint(name):
    """Autograd through the oracle's E_L and psi reproduces the reference's (tests/golden/vjp.npz, written by
    oracle/make_golden_vjp.py from the unmodified reference): pins the oracle for grad="auto" and the forces."""
    import os
    gold = np.load(os.path.join(C.GOLDEN, "vjp.npz"))
    g = C.load(name)
    mol, P = C.oracle_params(g)
    n = int(gold[name + "/n"][0])
    pos = torch.tensor(g["pos"][:n])
    gE = orc.local_energy_adjoint(P, pos, w_eloc=torch.tensor(gold[name + "/wE"]))
    gP = orc.local_energy_adjoint(P, pos, w_psi=torch.tensor(gold[name + "/wP"]))
    checked = 0
    for leaf in ("atom_coords", "bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w"):
        for tag, got in (("gE_", gE), ("gP_", gP)):
            key = name + "/" + tag + leaf
            if key not in gold.files or got.get(leaf) is None:
                continue
            ref = torch.tensor(gold[key])
            err = float((got[leaf] - ref).abs().max() / ref.abs().max().clamp(min=1e-4))
            assert err < 1e-10, (key, err)
            checked += 1
    assert checked >= 10
