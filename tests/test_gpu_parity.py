"""GPU: parity of the CUDA path (through the C ABI) with the reference's results.

Tolerances (BASELINE.json north_star): psi, E_L, gradients within 1e-10 relative in FP64;
Metropolis accept/reject decisions bit-exact under teacher forcing."""
import os

import numpy as np
import pytest
import torch

import _cases as C
import sj_oracle as orc
from qmctorch_b200.molecules import fixture_molecule

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _dev(x):
    return torch.as_tensor(x).cuda()


@pytest.mark.parametrize("name", C.CASES)
def test_golden_psi_energy_gradients(name):
    g = C.load(name)
    mol, wf = C.build_wf(g)
    pos = _dev(g["pos"])
    assert C.rel_err(wf(pos), g["psi"]) < RTOL
    assert C.rel_err(wf.local_energy(pos), g["eloc"]) < RTOL
    assert C.rel_err(wf.kinetic_energy(pos), g["ekin"]) < RTOL
    assert C.rel_err(wf.pdf(pos), g["psi"].reshape(-1) ** 2) < 2 * RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos), g["gpsi"]) < RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos, pdf=True), g["gpdf"]) < RTOL
    assert C.scaled_err(wf.pdf(pos, return_grad=True), g["gpdf"]) < RTOL


@pytest.mark.parametrize("name", C.CASES)
def test_golden_operators(name):
    """Sub-module parity: AtomicOrbitals, MolecularOrbitals, Jastrow factor, SlaterPooling."""
    g = C.load(name)
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    ns = g["ao"].shape[0]
    pos = _dev(g["pos"][:ns])
    ao, dao, d2ao = wf.ao(pos, derivative=[0, 1, 2])
    assert ao.shape == g["ao"].shape and dao.shape == g["dao"].shape and d2ao.shape == g["d2ao"].shape
    assert C.scaled_err(ao, g["ao"]) < RTOL
    assert C.scaled_err(dao, g["dao"]) < RTOL
    assert C.scaled_err(d2ao, g["d2ao"]) < RTOL
    assert C.scaled_err(wf.ao(pos), g["ao"]) < RTOL
    assert C.scaled_err(wf.ao(pos, derivative=1), g["dao"].sum(-1)) < RTOL
    assert C.scaled_err(wf.ao(pos, derivative=1, sum_grad=False), g["dao"]) < RTOL
    assert C.scaled_err(wf.ao(pos, derivative=2), g["d2ao"]) < RTOL
    # MO projection and Slater pooling against the oracle on the same AO values
    ao_ref = torch.tensor(g["ao"])
    mo_ref = orc.ao2mo(P, ao_ref)
    mo = wf.mo(_dev(g["ao"]))
    assert C.scaled_err(mo, mo_ref) < RTOL
    assert C.scaled_err(wf.pos2mo(pos), mo_ref) < RTOL
    dets_ref = orc.slater_dets(P, mo_ref)
    assert C.scaled_err(wf.pool(mo), dets_ref) < 1e-9
    bop = torch.randn(3, *mo_ref.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    tr_ref = orc.slater_trace(P, mo_ref, bop)
    assert C.scaled_err(wf.pool.operator(mo, bop.cuda()), tr_ref) < 1e-8
    assert C.scaled_err(wf.pool.operator(mo, bop[0].cuda()), tr_ref[0]) < 1e-8
    if g["jastrow"] != "None":
        posw = _dev(g["pos"])
        J, dJ, d2J = wf.jastrow(posw, derivative=[0, 1, 2], sum_grad=False)
        assert J.shape == g["J"].shape and dJ.shape == g["dJ"].shape and d2J.shape == g["d2J"].shape
        assert C.rel_err(J, g["J"]) < RTOL
        assert C.scaled_err(dJ, g["dJ"]) < RTOL
        assert C.scaled_err(d2J, g["d2J"]) < RTOL
        assert C.rel_err(wf.jastrow(posw), g["J"]) < RTOL
        assert C.scaled_err(wf.jastrow(posw, derivative=1), g["dJ"].sum(1)) < RTOL


@pytest.mark.parametrize("name", ["h2_single22", "lih_ground", "lih_cas24", "lih_een", "h2o_cas44", "c4h6_ground",
                                  "lih_sd22_een3", "h2o_cas44_een", "lih_adf_sd22", "co2_adf_ground"])
def test_metropolis_decisions_bit_exact_teacher_forced(name):
    """Same state, same proposal and uniform draws as the reference -> identical decisions,
    identical new positions, psi^2 within tolerance (sampler/metropolis.py:134-160,279-298)."""
    from qmctorch_b200 import _lib
    g = C.load(name)
    mol, wf = C.build_wf(g)
    L = _lib.lib()
    nstep = g["mh_disp"].shape[0]
    for it in range(nstep):
        x = _dev(g["mh_pos"][it]).clone()
        W = x.shape[0]
        fx = wf(x).reshape(-1) ** 2
        fx[fx == 0] = 1e-16
        acc = torch.zeros(W, dtype=torch.uint8, device="cuda")
        nacc = torch.zeros(1, dtype=torch.int64, device="cuda")
        disp, tau = _dev(g["mh_disp"][it]), _dev(g["mh_tau"][it])    # keep alive across the call
        fx = fx.detach().contiguous()
        _lib.check(L.qmcb_metropolis_step(
            wf._handle.plan(), _lib.ptr(x), _lib.ptr(fx), W, _lib.ptr(disp), _lib.ptr(tau), None, -1, 1, 1.0,
            1e-16, 0, 0, _lib.ptr(acc), _lib.ptr(nacc), _lib.stream_ptr(x.device)), "qmcb_metropolis_step")
        assert np.array_equal(acc.cpu().numpy().astype(bool), g["mh_acc"][it])
        assert int(nacc) == int(g["mh_acc"][it].sum())
        assert np.array_equal(x.cpu().numpy(), g["mh_pos"][it + 1])          # bit-exact positions
        a = g["mh_acc"][it]
        if a.any():
            assert C.rel_err(fx[torch.as_tensor(a).cuda()], g["mh_fxn"][it][a]) < 2 * RTOL


def _thermalised(wf, mol, nw, nstep=60, step=0.3, seed=11):
    from qmctorch_b200.sampler import Metropolis
    torch.manual_seed(seed)
    s = Metropolis(nwalkers=nw, nstep=nstep, step_size=step, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                   move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=seed, keep_on_device=True)
    return s(wf.pdf, with_tqdm=False).detach(), s


@pytest.mark.parametrize("name,nw", [("lih_ground", 20000), ("h2_single22", 20000), ("lih_een", 4000),
                                     ("h2o_cas44", 1500), ("lih_sd22_een3", 1500), ("h2o_cas44_een", 400)])
def test_fresh_walkers_against_oracle(name, nw):
    """Seeded ensembles the fixtures have never seen, CUDA vs oracle, per-walker relative error."""
    g = C.load(name)
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    pos, _ = _thermalised(wf, mol, nw)
    cpu = pos.cpu()
    psi_o, el_o = orc.psi(P, cpu), orc.local_energy(P, cpu)
    assert C.rel_err(wf(pos), psi_o) < RTOL
    e = wf.local_energy(pos).cpu()
    # relative to max(|E_L|, 1): E_L crosses zero for a few walkers
    assert float(((e - el_o).abs() / el_o.abs().clamp(min=1.0)).max()) < RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos[:512]), orc.grad_psi(P, cpu[:512])) < RTOL


_GJ_SYSTEMS = {
    # key: (atoms [angstrom], spin) - DZP carbon / hydrogen tables, seeded orthonormal MOs
    # C2H4: 16 electrons, 8x8 blocks (padded order 8); triplet: 9x9 / 7x7 (padded order 12, the smaller
    # block idles through two steps); C2H6: 18 electrons, 9x9 blocks (padded order 12)
    "c2h4": ("C 0 0 0.667; C 0 0 -0.667; H 0 0.923 1.238; H 0 -0.923 1.238; H 0 0.923 -1.238; H 0 -0.923 -1.238", 0),
    "c2h4_triplet": ("C 0 0 0.667; C 0 0 -0.667; H 0 0.923 1.238; H 0 -0.923 1.238; H 0 0.923 -1.238; H 0 -0.923 -1.238", 2),
    "c2h6": ("C 0 0 0.765; C 0 0 -0.765; H 1.019 0 1.158; H -0.510 0.883 1.158; H -0.510 -0.883 1.158; "
             "H -1.019 0 -1.158; H 0.510 0.883 -1.158; H 0.510 -0.883 -1.158", 0),
}


@pytest.mark.parametrize("jit", ["0", "2"])
@pytest.mark.parametrize("key", sorted(_GJ_SYSTEMS))
def test_half_warp_gauss_jordan_orders(key, jit, monkeypatch):
    """Spin blocks of order 7..9 (the row-owner half-warp Gauss-Jordan with padded orders 8 and 12,
    unequal blocks in one warp; C4H6 covers 15 -> 16): psi, E_L, grad psi and one Metropolis
    decision against the oracle on thermalised walkers - on the generic CTA-tile kernels (QMCB_JIT=0)
    and on the structure-specialised warp-tile kernels (spec_tile.cuh)."""
    monkeypatch.setenv("QMCB_JIT", jit)
    from qmctorch_b200.molecules import Molecule, _seeded_mos, build_basis, _parse_atoms
    from qmctorch_b200.wavefunction import SlaterJastrow
    from qmctorch_b200.wavefunction.pooling import OrbitalConfigurations
    atoms, spin = _GJ_SYSTEMS[key]
    names, coords = _parse_atoms(atoms, "angs")
    nao = build_basis(names, coords, "dzp").nao
    mol = Molecule(atoms, basis="dzp", unit="angs", spin=spin, name=key, mos=_seeded_mos(nao, 5))
    wf = SlaterJastrow(mol, configs="ground_state", cuda=True)
    assert max(mol.nup, mol.ndown) in (8, 9) and wf._handle.info(15) == (0 if jit == "0" else 2)
    P = orc.make_params(mol, OrbitalConfigurations(mol).get_configs("ground_state"), jastrow_weight=1.0)
    pos, _ = _thermalised(wf, mol, 777, nstep=40, step=0.1)
    cpu = pos.cpu()
    psi_o, el_o = orc.psi(P, cpu), orc.local_energy(P, cpu)
    assert C.rel_err(wf(pos), psi_o) < RTOL
    e = wf.local_energy(pos).cpu()
    assert float(((e - el_o).abs() / el_o.abs().clamp(min=1.0)).max()) < RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos[:128]), orc.grad_psi(P, cpu[:128])) < RTOL
    # one teacher-forced Metropolis step: identical decisions
    gen = torch.Generator().manual_seed(3)
    disp = 0.05 * torch.randn(pos.shape, generator=gen, dtype=torch.float64)
    tau = torch.rand(pos.shape[0], generator=gen, dtype=torch.float64)
    fx = (psi_o ** 2).reshape(-1)
    x_o, fx_o, acc_o, fxn_o = orc.metropolis_step(P, cpu, fx, disp, tau)
    from qmctorch_b200 import _lib
    x1, fx1, d_disp, d_tau = pos.clone(), fx.cuda().contiguous(), disp.cuda(), tau.cuda()
    acc1 = torch.zeros(pos.shape[0], dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().qmcb_metropolis_step(
        wf._handle.plan(), _lib.ptr(x1), _lib.ptr(fx1), pos.shape[0], _lib.ptr(d_disp), _lib.ptr(d_tau), None, -1, 1,
        1.0, 1e-16, 0, 0, _lib.ptr(acc1), None, _lib.stream_ptr(x1.device)), "qmcb_metropolis_step")
    torch.cuda.synchronize()
    # (a decision whose margin is below the 1e-10 tolerance of psi^2 itself is not determined)
    margin = ((fxn_o / fx).clamp(max=1.0) - tau).abs()
    sure = margin > 1e-8
    assert int(sure.sum()) >= pos.shape[0] - 2
    assert torch.equal(acc1.cpu().bool()[sure], acc_o.bool().reshape(-1)[sure])
    assert torch.equal(x1.cpu()[sure], x_o[sure])
    assert 0.05 < float(acc_o.float().mean()) < 0.95


def test_sampler_replays_reference_draw_sequence():
    """rng='torch' makes the same generator calls in the same order as the reference
    (Appendix C): the whole trajectory equals the oracle's step by step, bit for bit."""
    from qmctorch_b200.sampler import Metropolis
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    nw, nstep, step = 300, 12, 0.3
    torch.manual_seed(21)
    s = Metropolis(nwalkers=nw, nstep=nstep, step_size=step, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                   move={"type": "all-elec", "proba": "normal"}, cuda=True, rng="torch")
    out = s(wf.pdf, with_tqdm=False)
    assert out.device.type == "cpu" and out.requires_grad and out.shape == (nw, 3 * wf.nelec)
    torch.manual_seed(21)
    from torch.distributions import MultivariateNormal
    d = mol.domain("normal")
    pos = MultivariateNormal(torch.as_tensor(d["mean"]), torch.as_tensor(d["sigma"])).sample((nw, wf.nelec))
    pos = pos.type(torch.float64).view(nw, -1)
    fx = (orc.psi(P, pos) ** 2).reshape(-1)
    for _ in range(nstep):
        disp = s.multiVariate.sample((nw, wf.nelec)).view(nw, -1)
        tau = torch.rand(nw, dtype=torch.float64)
        pos, fx, acc, _ = orc.metropolis_step(P, pos, fx, disp, tau)
    assert torch.equal(out.detach(), pos)


def test_full_size_properties_lih_1m():
    """BASELINE config 2 size (1e6 walkers): size-independent properties."""
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    nw = 1_000_000
    pos, sampler = _thermalised(wf, mol, nw, nstep=30)
    assert pos.shape == (nw, 12) and 0.05 < sampler.acceptance_rate < 0.99
    psi = wf(pos)
    e, p2, k = wf._eloc(pos, want_psi=True, want_ekin=True)
    assert torch.isfinite(e).all() and torch.isfinite(psi).all()
    assert C.rel_err(p2, psi) < 1e-12                            # both kernels, same psi
    # antisymmetry under exchange of the two spin-up / the two spin-down electrons
    sw = pos.clone()
    sw[:, 0:3], sw[:, 3:6] = pos[:, 3:6], pos[:, 0:3]
    assert float(((wf(sw) + psi).abs() / psi.abs()).max()) < 1e-9
    assert float(((wf.local_energy(sw) - e).abs() / e.abs().clamp(min=1.0)).max()) < 1e-7
    sw = pos.clone()
    sw[:, 6:9], sw[:, 9:12] = pos[:, 9:12], pos[:, 6:9]
    assert float(((wf(sw) + psi).abs() / psi.abs()).max()) < 1e-9
    # energy is physical: LiH ground state is about -8.07 Ha; this trial function gives ~ -7.9
    mean = float(e.mean())
    assert -8.3 < mean < -7.5
    # Metropolis: tau=0 accepts everything that is finite, tau=2 accepts nothing
    from qmctorch_b200 import _lib
    L = _lib.lib()
    x = pos[:100000].clone()
    fx = (wf(x).reshape(-1) ** 2).detach()
    disp = 0.05 * torch.randn_like(x)
    acc = torch.zeros(x.shape[0], dtype=torch.uint8, device="cuda")
    for tau_val, expect_all in ((0.0, True), (2.0, False)):
        y, fy = x.clone(), fx.clone()
        tau = torch.full((x.shape[0],), tau_val, dtype=torch.float64, device="cuda")
        _lib.check(L.qmcb_metropolis_step(wf._handle.plan(), _lib.ptr(y), _lib.ptr(fy), y.shape[0], _lib.ptr(disp),
                                          _lib.ptr(tau), None, -1, 1, 1.0, 1e-16, 0, 0, _lib.ptr(acc), None,
                                          _lib.stream_ptr(y.device)), "qmcb_metropolis_step")
        if expect_all:
            assert bool(acc.all()) and torch.equal(y, x + disp)
            assert C.rel_err(fy, wf(x + disp).reshape(-1) ** 2) < 1e-12
        else:
            assert not bool(acc.any()) and torch.equal(y, x) and torch.equal(fy, fx)


def test_edge_cases_empty_single_ragged_nan():
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    pos = _dev(g["pos"])
    tw = wf._handle.info(6)
    assert wf(pos[:0]).shape == (0, 1) and wf.local_energy(pos[:0]).shape == (0, 1)
    ref_psi, ref_e = wf(pos), wf.local_energy(pos)
    for n in (1, 2, tw - 1, tw, tw + 1, 2 * tw + 3):
        n = min(n, pos.shape[0])
        assert torch.equal(wf(pos[:n]), ref_psi[:n])             # results do not depend on the tiling
        assert torch.equal(wf.local_energy(pos[:n]), ref_e[:n])
    bad = pos[:8].clone()
    bad[3, 4] = float("nan")
    e = wf.local_energy(bad)
    assert torch.isnan(e[3]).all() and torch.isfinite(e[[0, 1, 2, 4, 5, 6, 7]]).all()
    # coincident electrons: 1/r_ij diverges; must not poison neighbours
    bad = pos[:8].clone()
    bad[2, 0:3] = bad[2, 6:9]
    e = wf.local_energy(bad)
    assert not torch.isfinite(e[2]).all() or float(e[2].abs()) > 1e6
    assert torch.isfinite(e[[0, 1, 3]]).all()
    with pytest.raises(ValueError):
        wf(pos[:, :6])
    # non-contiguous / float32 / CPU inputs are converted, not misread
    assert torch.equal(wf(pos.cpu()), ref_psi)
    assert torch.equal(wf(pos.t().contiguous().t()), ref_psi)


def test_parameter_update_refreshes_device_tables():
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    pos = _dev(g["pos"][:64])
    p0 = wf(pos).clone()
    with torch.no_grad():
        wf.jastrow.jastrow_kernel.weight.mul_(1.1)
        wf.mo.mo_modifier[0, 0] *= 1.01
    p1 = wf(pos)
    assert not torch.equal(p0, p1)
    mol_o, P = C.oracle_params(g)
    P.jastrow_weight = P.jastrow_weight * 1.1
    P.mo_modifier[0, 0] *= 1.01
    assert C.rel_err(p1, orc.psi(P, pos.cpu())) < RTOL
    sd = {k: v.clone() for k, v in wf.state_dict().items()}
    mol2, wf2 = C.build_wf(g)
    wf2.load_state_dict(sd)
    assert torch.equal(wf2(pos), p1)


def test_solver_single_point_h2_config1():
    """BASELINE config 1 (H2 STO-3G, single(2,2), 1000 walkers x 2000 steps, step 0.5): the VMC
    energy must be consistent with the reference's (-1.11 +/- 0.02 Ha measured here with the
    reference itself on these orbitals; tests/solver/test_h2_pyscf_metropolis.py declares
    -1.146 for SCF orbitals)."""
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    g = C.load("h2_single22")
    mol, wf = C.build_wf(g)
    torch.manual_seed(0)
    sampler = Metropolis(nwalkers=1000, nstep=2000, step_size=0.5, ndim=wf.ndim, nelec=wf.nelec,
                         init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True,
                         seed=1)
    solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.Adam(wf.parameters(), lr=0.01))
    obs = solver.single_point(with_tqdm=False)
    assert obs.pos.shape == (1000, 6) and obs.local_energy.shape == (1000, 1)
    mol_o, P = C.oracle_params(g)
    assert C.rel_err(obs.local_energy, orc.local_energy(P, obs.pos.detach().cpu())) < RTOL
    assert abs(float(obs.energy) - (-1.13)) < 0.12
    assert float(obs.error) < 0.05


def _manual_grads(wf, pos):
    """Solver.evaluate_grad_manual (solver/solver.py:372-431) through the public API."""
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    mol = wf.mol
    sampler = Metropolis(nwalkers=pos.shape[0], nstep=2, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                         cuda=True)
    opt = torch.optim.SGD(wf.parameters(), lr=0.0)
    solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
    solver.configure(track=["local_energy"], loss="energy", grad="manual")
    opt.zero_grad()
    wf.ao.bas_coeffs.grad = None      # plain tensor, not reached by zero_grad (solver.py:119-120)
    solver.evaluate_grad_manual(pos)
    out = {"mo_modifier": wf.mo.mo_modifier.grad, "ci": wf.fc.weight.grad, "bas_exp": wf.ao.bas_exp.grad,
           "bas_coeffs": wf.ao.bas_coeffs.grad}
    if wf._jee is not None:
        out["jastrow_weight"] = wf._jee.jastrow_kernel.weight.grad
    if wf._jen is not None:
        out["en_weight"] = wf._jen.jastrow_kernel.weight.grad
    if wf._jeen is not None:
        bh = wf._jeen.jastrow_kernel
        out["een_num"], out["een_denom"], out["een_fc"] = bh.weight_num.grad, bh.weight_denom.grad, bh.fc.weight.grad
    return out


@pytest.mark.parametrize("bwd_spec", ["1", "0"])
@pytest.mark.parametrize("name", C.CASES)
def test_parameter_gradients_match_reference(name, bwd_spec, monkeypatch):
    """psi.backward(weight) of the reference solver (golden) vs qmcb_psi_backward: through the
    structure-specialised backward where the structure has one (spec_backward_all: one walker per thread,
    register accumulators) and through the DMMA tile kernel (QMCB_BWD_SPEC=0)."""
    monkeypatch.setenv("QMCB_BWD_SPEC", bwd_spec)
    g = C.load(name)
    mol, wf = C.build_wf(g)
    grads = _manual_grads(wf, _dev(g["pos"]))
    for k in [k[5:] for k in g if k.startswith("grad_")]:
        if k == "ci" and g["ci"].shape[1] == 1:
            continue            # sum_w (E_L - <E_L>) = 0: rounding noise in the reference itself
        ref = torch.tensor(g["grad_" + k])
        got = grads[k].detach().cpu().reshape(ref.shape)
        err = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-6))
        assert err < RTOL, (k, err)


def test_backward_is_deterministic_and_matches_oracle_on_fresh_walkers():
    g = C.load("lih_sd22")
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    pos, _ = _thermalised(wf, mol, 3000)
    a = _manual_grads(wf, pos)
    a = {k: v.clone() for k, v in a.items()}
    b = _manual_grads(wf, pos)
    for k in a:
        assert torch.equal(a[k], b[k]), k                     # bitwise reproducible
    og, _ = orc.param_grads(P, pos.cpu(), names=("jastrow_weight", "mo_modifier", "ci", "bas_exp", "bas_coeffs"))
    for k, ref in og.items():
        got = a[k].detach().cpu().reshape(ref.shape)
        err = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-6))
        assert err < 1e-9, (k, err)


def test_solver_optimisation_lowers_energy_and_matches_oracle_step():
    """BASELINE config 3 in miniature: LiH, Jastrow + MO coefficients (freeze ci, ao)."""
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    torch.manual_seed(5)
    sampler = Metropolis(nwalkers=20000, nstep=200, step_size=0.3, nelec=wf.nelec, ndim=3,
                         init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=5)
    opt = torch.optim.SGD(wf.parameters(), lr=0.05)
    solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
    solver.configure(track=["local_energy", "parameters"], freeze=["ci", "ao"], loss="energy", grad="manual",
                     resampling={"mode": "update", "resample_every": 1, "nstep_update": 30})
    assert not wf.fc.weight.requires_grad and not wf.ao.bas_exp.requires_grad
    obs = solver.run(6, tqdm=False)
    # initial sampling stores the energy only (solver_base.py:192-201), then one entry per epoch
    assert len(obs.energy) == 7 and len(obs.local_energy) == 6
    assert all(np.isfinite(obs.energy))
    assert wf.ao.bas_exp.grad is None and wf.mo.mo_modifier.grad is not None
    assert np.mean(obs.energy[-2:]) < obs.energy[0] + 0.02          # energy does not go up
    assert hasattr(obs.models, "best") and "mo.mo_modifier" in obs.models.best


def test_solver_tracked_local_energies_are_staged_behind_the_resampling():
    """Large ensembles: the per-epoch local energies reach the host through a pinned ring while the resampling
    kernels run (Solver._stage / _flush_observables, Metropolis.host_work).  The observable must be what the
    blocking path stores: one float64 numpy array per epoch whose mean is the tracked energy of that epoch."""
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    sampler = Metropolis(nwalkers=70000, nstep=60, step_size=0.3, nelec=wf.nelec, ndim=3,
                         init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=2)
    solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.SGD(wf.parameters(), lr=0.01))
    solver.configure(track=["local_energy"], freeze=["ci", "ao"], loss="energy", grad="manual",
                     resampling={"mode": "update", "resample_every": 1, "nstep_update": 10})
    obs = solver.run(4, tqdm=False)
    assert solver._pending is None and sampler.host_work is None and not solver._defer_observables
    assert len(obs.local_energy) == 4 and len(obs.energy) == 5
    for k, e in enumerate(obs.local_energy):
        assert isinstance(e, np.ndarray) and e.dtype == np.float64 and e.shape == (70000, 1)
        assert abs(float(e.mean()) - obs.energy[k + 1]) < 1e-9 * abs(obs.energy[k + 1])
    assert not np.array_equal(obs.local_energy[0], obs.local_energy[1])          # ring slots were copied out


@pytest.mark.parametrize("move_type,proba", [("one-elec", "normal"), ("all-elec-iter", "normal"),
                                             ("all-elec", "uniform"), ("one-elec", "uniform")])
def test_sampler_move_types_replay_reference_draws(move_type, proba):
    """Every move type / proposal of sampler/metropolis.py:179-277 with rng='torch': same generator
    calls in the same order as the reference, so the trajectory equals the oracle's bit for bit."""
    from qmctorch_b200.sampler import Metropolis
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    nw, nstep, step, ne = 200, 6, 0.4, wf.nelec
    torch.manual_seed(77)
    s = Metropolis(nwalkers=nw, nstep=nstep, step_size=step, nelec=ne, ndim=3, init=mol.domain("normal"),
                   move={"type": move_type, "proba": proba}, cuda=True, rng="torch")
    out = s(wf.pdf, with_tqdm=False).detach()
    # oracle replay
    torch.manual_seed(77)
    from torch.distributions import MultivariateNormal
    d = mol.domain("normal")
    pos = MultivariateNormal(torch.as_tensor(d["mean"]), torch.as_tensor(d["sigma"])).sample((nw, ne))
    pos = pos.type(torch.float64).view(nw, -1)
    fx = (orc.psi(P, pos) ** 2).reshape(-1)

    def draw(n):
        if proba == "uniform":
            return step * (2.0 * torch.rand((nw, n, 3), dtype=torch.float64) - 1.0)
        return s.multiVariate.sample((nw, n)).to(torch.float64)

    for _ in range(nstep):
        for id_elec in (range(ne) if move_type == "all-elec-iter" else [None]):
            if move_type == "all-elec":
                disp = draw(ne).view(nw, -1)
            else:
                idx = torch.LongTensor(nw).random_(0, ne) if id_elec is None else torch.full((nw,), id_elec)
                full = torch.zeros(nw, ne, 3, dtype=torch.float64)
                full[torch.arange(nw), idx, :] = draw(1).view(nw, 3)
                disp = full.view(nw, -1)
            tau = torch.rand(nw, dtype=torch.float64)
            pos, fx, acc, _ = orc.metropolis_step(P, pos, fx, disp, tau)
    assert torch.equal(out, pos)
    assert 0.0 < s.acceptance_rate <= 1.0


def test_philox_sampler_is_tiling_independent_and_samples_psi2():
    """In-kernel Philox draws are functions of (seed, step, global index): the same seed gives the
    same ensemble whatever the number of walkers processed together; the sampled energy agrees
    with the torch-draw sampler within statistical error."""
    from qmctorch_b200.sampler import Metropolis
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    torch.manual_seed(3)
    init = Metropolis(nwalkers=4096, nstep=1, nelec=wf.nelec, ndim=3, init=mol.domain("normal"), cuda=True)
    init.walkers.initialize()
    start = init.walkers.pos.clone()

    def run(pos, seed, rng):
        s = Metropolis(nwalkers=pos.shape[0], nstep=150, step_size=0.3, nelec=wf.nelec, ndim=3,
                       init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=seed,
                       rng=rng, keep_on_device=True)
        return s(wf.pdf, pos=pos.clone(), with_tqdm=False).detach()
    a = run(start, 9, "philox")
    b = run(start, 9, "philox")
    assert torch.equal(a, b)
    head = run(start[:1000], 9, "philox")
    assert torch.equal(head, a[:1000])                      # independent of the ensemble size / tiling
    assert not torch.equal(run(start, 10, "philox"), a)
    e_philox = wf.local_energy(a)
    torch.manual_seed(4)
    e_torch = wf.local_energy(run(start, 0, "torch"))
    err = float((e_philox.var() / len(e_philox) + e_torch.var() / len(e_torch)).sqrt())
    assert abs(float(e_philox.mean() - e_torch.mean())) < 6 * err


def test_solver_api_batching_trajectory_checkpoint(tmp_path):
    """Solver.single_point(batchsize), sampling_traj, save/load checkpoint (solver_base.py:316-471)."""
    import os
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    g = C.load("lih_sd22")
    mol, wf = C.build_wf(g)
    torch.manual_seed(2)
    sampler = Metropolis(nwalkers=500, nstep=40, step_size=0.3, ntherm=30, ndecor=5, nelec=wf.nelec, ndim=3,
                         init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=2)
    assert sampler.get_sampling_size() == 500 * 2
    opt = torch.optim.Adam(wf.parameters(), lr=0.01)
    solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
    full = solver.single_point(with_tqdm=False)
    assert full.pos.shape == (1000, 12)
    with torch.no_grad():
        whole = wf.local_energy(full.pos)
        parts = torch.cat([wf.local_energy(full.pos[i:i + 300]) for i in range(0, 1000, 300)])
    assert torch.equal(whole, parts)                       # batching does not change a single bit
    traj = solver.sampling_traj(pos=full.pos.detach(), with_tqdm=False)
    assert traj.local_energy.shape == (2, 500)
    assert np.allclose(traj.local_energy.reshape(-1), whole.cpu().numpy().reshape(-1), rtol=0, atol=0)
    # checkpoint round trip keeps the device tables in sync
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with torch.no_grad():
            wf.jastrow.jastrow_kernel.weight.fill_(0.9)
        before = wf(full.pos[:64]).detach().clone()
        solver.save_checkpoint(3, 0.5)
        with torch.no_grad():
            wf.jastrow.jastrow_kernel.weight.fill_(1.7)
        assert not torch.equal(wf(full.pos[:64]).detach(), before)
        epoch, loss = solver.load_checkpoint("checkpoint_epoch3.pth")
        assert epoch == 3 and loss == 0.5
        assert torch.equal(wf(full.pos[:64]).detach(), before)
    finally:
        os.chdir(cwd)


def test_one_electron_ao_and_update():
    """AtomicOrbitals.forward(one_elec=True) and .update (atomic_orbitals.py:131-218,671-695)."""
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    pos = _dev(g["pos"][:32])
    ao = wf.ao(pos)
    one = wf.ao(pos[:, 3:6], one_elec=True)
    assert one.shape == (32, 1, ao.shape[-1])
    assert torch.equal(one[:, 0], ao[:, 1])
    moved = pos.clone()
    moved[:, 3:6] += 0.1
    upd = wf.ao.update(ao, moved, 1)
    assert torch.equal(upd, wf.ao(moved))
    assert torch.equal(upd[:, 0], ao[:, 0]) and not torch.equal(upd[:, 1], ao[:, 1])


def _mh_philox(wf, x, tau, seed, offset, scale=0.3, move_elec=-1):
    """One in-kernel-Philox Metropolis move through the C ABI; returns (new pos, accept mask)."""
    from qmctorch_b200 import _lib
    L = _lib.lib()
    x = x.clone()
    W = x.shape[0]
    fx = (wf(x).reshape(-1) ** 2).detach().contiguous()
    acc = torch.zeros(W, dtype=torch.uint8, device="cuda")
    _lib.check(L.qmcb_metropolis_step(
        wf._handle.plan(), _lib.ptr(x), _lib.ptr(fx), W, None, _lib.ptr(tau) if tau is not None else None, None,
        move_elec, 1, scale, 1e-16, seed, offset, _lib.ptr(acc), None, _lib.stream_ptr(x.device)),
        "qmcb_metropolis_step")
    torch.cuda.synchronize()
    return x, acc.bool()


# structures of the warp-tile kernels (spec_tile.cuh): CI expansions over 5 x 5 blocks, 11 x 11 and
# 15 x 15 blocks, Slater radial functions with d shells, e-n and three-body Jastrow factors
_TILE_CASES = ["h2o_ground", "h2o_cas44", "c4h6_ground", "co2_adf_ground", "lih_sd22_een3", "h2o_cas44_een"]
# small structures FORCED onto the warp-tile kernels (QMCB_SPEC_KIND=tile): 8 and 16 walkers per warp,
# closed-form determinants, e-n Jastrow, Slater radial functions
_FORCED_TILE = ["lih_een@tile", "lih_cas24@tile", "h2_single22@tile", "lih_sto@tile", "lih_nojastrow@tile"]


@pytest.mark.parametrize("name", ["lih_ground", "h2_single22", "lih_sd22", "lih_cas24", "lih_nojastrow", "h2_ground",
                                  "lih_sto", "lih_sto_pure", "lih_gto_kr", "lih_adf_sd22", "lih_een"] + _TILE_CASES
                         + _FORCED_TILE)
def test_specialised_kernels_match_generic(name, monkeypatch):
    """The NVRTC structure-specialised kernels (spec_kernel.cuh: one walker per thread; spec_tile.cuh:
    warp tiles) against the generic interpreter kernels (fused_impl.cuh, QMCB_JIT=0) on the same
    walkers: psi, E_L, E_kin to rounding, identical Philox proposals, identical accept decisions."""
    tile = name in _TILE_CASES or name.endswith("@tile")
    if name.endswith("@tile"):
        name = name[:-5]
        monkeypatch.setenv("QMCB_SPEC_KIND", "tile")
    g = C.load(name)
    monkeypatch.setenv("QMCB_JIT", "0")
    mol, wf0 = C.build_wf(g)
    step = float(g["step"])
    scale = (step / (2.0 * np.sqrt(2.0 * np.log(2.0)))) ** 0.5     # proposal std of the reference (metropolis.py:207-212)
    pos, _ = _thermalised(wf0, mol, 1003 if mol.nelec > 20 else 5003, step=step)   # ragged: not a multiple of the tile size
    assert wf0._handle.info(13) == 0
    monkeypatch.setenv("QMCB_JIT", "2")              # 2: a missing / failing NVRTC is an error, not a fallback
    mol, wf1 = C.build_wf(g)
    assert wf1._handle.info(14) == (2 if tile else 1) and wf1._handle.info(13) == 1
    assert wf1._handle.info(15) == (2 if tile else 1)
    for f in ("__call__", "local_energy", "kinetic_energy"):
        a, b = getattr(wf0, f)(pos), getattr(wf1, f)(pos)
        # psi to rounding; energies can pass through zero for a walker, which inflates the
        # element-wise relative error of both kernels alike
        assert C.rel_err(b, a) < (1e-12 if f == "__call__" else RTOL), f
    x0, a0 = _mh_philox(wf0, pos, None, 17, 3, scale=scale)
    x1, a1 = _mh_philox(wf1, pos, None, 17, 3, scale=scale)
    assert torch.equal(a0, a1) and torch.equal(x0, x1)
    assert 0.02 < float(a1.float().mean()) < 0.95
    for me in (-2, 1):                              # one random electron / electron 1 only
        x0, a0 = _mh_philox(wf0, pos, None, 5, 8, scale=scale, move_elec=me)
        x1, a1 = _mh_philox(wf1, pos, None, 5, 8, scale=scale, move_elec=me)
        assert torch.equal(a0, a1) and torch.equal(x0, x1)
        moved = ((x1 - pos).reshape(len(pos), -1, 3).abs().sum(-1) > 0).sum(-1)
        assert int(moved.max()) <= 1
    # uniform proposals and injected displacements / acceptance draws take the other coordinate path
    gen = torch.Generator().manual_seed(9)
    disp = (scale * torch.randn(pos.shape, generator=gen, dtype=torch.float64)).cuda()
    tau = torch.rand(pos.shape[0], generator=gen, dtype=torch.float64).cuda()
    outs = []
    for wf in (wf0, wf1):
        x, fx = pos.clone(), (wf(pos).reshape(-1) ** 2).detach().contiguous()
        acc = torch.zeros(pos.shape[0], dtype=torch.uint8, device="cuda")
        from qmctorch_b200 import _lib
        _lib.check(_lib.lib().qmcb_metropolis_step(
            wf._handle.plan(), _lib.ptr(x), _lib.ptr(fx), pos.shape[0], _lib.ptr(disp), _lib.ptr(tau), None, -1, 1,
            1.0, 1e-16, 0, 0, _lib.ptr(acc), None, _lib.stream_ptr(x.device)), "qmcb_metropolis_step")
        x2 = pos.clone()
        _lib.check(_lib.lib().qmcb_metropolis_step(
            wf._handle.plan(), _lib.ptr(x2), _lib.ptr(fx.clone()), pos.shape[0], None, None, None, -1, 0,
            step, 1e-16, 3, 1, None, None, _lib.stream_ptr(x.device)), "qmcb_metropolis_step")
        torch.cuda.synchronize()
        outs.append((x, acc.clone(), x2))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][2], outs[1][2])


def test_specialised_kernel_follows_parameter_updates():
    """An optimiser step rewrites the parameter block of the specialised kernel (no recompilation):
    results track the generic oracle after in-place parameter changes."""
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    _, P = C.oracle_params(g)
    pos = _dev(g["pos"])
    assert wf._handle.info(13) == 1
    with torch.no_grad():
        wf.mo.mo_modifier.mul_(1.0 + 0.03 * torch.rand_like(wf.mo.mo_modifier))
        wf.ao.bas_exp.mul_(1.02)
        wf.jastrow.jastrow_kernel.weight.fill_(0.6)
    P.mo_modifier = wf.mo.mo_modifier.detach().cpu().clone()
    P.bas_exp = wf.ao.bas_exp.detach().cpu().clone()
    P.jastrow_weight = torch.tensor([0.6], dtype=torch.float64)
    assert C.rel_err(wf(pos), orc.psi(P, pos.cpu())) < RTOL
    assert C.rel_err(wf.local_energy(pos), orc.local_energy(P, pos.cpu())) < RTOL
    assert wf._handle.info(13) == 1
    # the s and p primitives of the 6-31G SP shells share one exponential in the generated program
    # as long as their exponents are bitwise equal; an update that separates them must select a
    # program without the sharing
    with torch.no_grad():
        torch.manual_seed(3)
        wf.ao.bas_exp.mul_(1.0 + 0.05 * torch.rand_like(wf.ao.bas_exp))
    P.bas_exp = wf.ao.bas_exp.detach().cpu().clone()
    assert C.rel_err(wf(pos), orc.psi(P, pos.cpu())) < RTOL
    assert C.rel_err(wf.local_energy(pos), orc.local_energy(P, pos.cpu())) < RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos), orc.grad_psi(P, pos.cpu())) < RTOL
    assert wf._handle.info(13) == 1


def test_philox_normal_draws_are_standard_symmetric_and_tiling_free():
    """In-kernel proposal draws (philox.cuh: FP32 Box-Muller on the SFU, four normals per Philox
    call): with tau = 0 every move is accepted, so (x' - x)/scale exposes the draws.  Moments of a
    standard normal, exact sign symmetry of the pair construction, and element g of the ensemble
    gets the same draw whatever slice of walkers is processed."""
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    pos, _ = _thermalised(wf, mol, 200_000, nstep=5)
    tau = torch.zeros(len(pos), dtype=torch.float64, device="cuda")
    x, acc = _mh_philox(wf, pos, tau, 123, 7, scale=0.25)
    assert bool(acc.all())
    z = ((x - pos) / 0.25).reshape(-1)
    n = z.numel()
    assert abs(float(z.mean())) < 5 / n ** 0.5
    assert abs(float(z.var()) - 1.0) < 5 * (2.0 / n) ** 0.5
    assert abs(float((z ** 3).mean())) < 5 * (15.0 / n) ** 0.5
    assert abs(float((z ** 4).mean()) - 3.0) < 5 * (96.0 / n) ** 0.5
    assert 4.0 < float(z.abs().max()) < 5.9
    # every coordinate of every electron is drawn independently: lag-1 correlation ~ 0
    assert abs(float((z[:-1] * z[1:]).mean())) < 5 / n ** 0.5
    # a slice of the ensemble starting at a walker boundary that is not a multiple of 4 elements
    part, _ = _mh_philox(wf, pos[:777], tau[:777], 123, 7, scale=0.25)
    assert torch.equal(part, x[:777])
    other, _ = _mh_philox(wf, pos, tau, 123, 8, scale=0.25)
    assert not torch.equal(other, x)


@pytest.mark.parametrize("name,jit", [("lih_ground", "2"), ("lih_ground", "0"), ("h2o_ground", "1")])
def test_local_energy_stats_one_call(name, jit, monkeypatch):
    """qmcb_local_energy_stats (E_L + [sum, sum sq, n finite, n non-finite] in one call, fused into the
    specialised kernel / two extra kernels behind the generic one) against qmcb_local_energy +
    torch sums; deterministic; non-finite walkers are counted, not summed."""
    monkeypatch.setenv("QMCB_JIT", jit)
    g = C.load(name)
    mol, wf = C.build_wf(g)
    pos, _ = _thermalised(wf, mol, 30011, step=g["step"])
    pos[17, 0] = float("nan")
    e_ref = wf.local_energy(pos)
    e, out4 = wf.local_energy_stats(pos)
    e2, out4b = wf.local_energy_stats(pos)
    assert torch.equal(out4, out4b) and torch.equal(torch.nan_to_num(e, nan=7.0), torch.nan_to_num(e2, nan=7.0))
    ok = torch.isfinite(e_ref.reshape(-1))
    assert torch.equal(e[ok.reshape(-1, 1)], e_ref[ok.reshape(-1, 1)])
    assert int(out4[2]) == int(ok.sum()) and int(out4[3]) == 1
    good = e_ref.reshape(-1)[ok]
    assert abs(float(out4[0]) - float(good.sum())) < 1e-9 * float(good.abs().sum())
    assert abs(float(out4[1]) - float((good * good).sum())) < 1e-9 * float((good * good).sum())


def test_gradient_samplers_replay_reference_chains(double_default):
    """GeneralizedMetropolis and Hamiltonian on the analytic grad psi^2 of qmcb_grad_psi
    (instead of autograd) reproduce the reference's chains from the same torch seed: the committed
    fixture holds the positions the unmodified reference samplers returned."""
    import os
    from qmctorch_b200.sampler import GeneralizedMetropolis, Hamiltonian
    f = dict(np.load(os.path.join(C.GOLDEN, "gradient_samplers.npz")))
    g = C.load("lih_ground")
    g = dict(g, mo_modifier=np.ones_like(g["mo_modifier"]), jw=f["jw"])
    mol, wf = C.build_wf(g)
    start = _dev(f["start"])
    nstep, ntherm, ndecor = (int(v) for v in f["gm_cfg"])
    s = GeneralizedMetropolis(nwalkers=len(start), nstep=nstep, step_size=float(f["gm_step"][0]), ntherm=ntherm,
                              ndecor=ndecor, nelec=wf.nelec, ndim=3, init=mol.domain("normal"), cuda=True)
    torch.manual_seed(int(f["gm_seed"][0]))
    out = s(wf.pdf, pos=start.clone(), with_tqdm=False)
    assert out.requires_grad and out.shape == f["gm_pos"].shape
    assert C.scaled_err(out.detach(), f["gm_pos"]) < 1e-9
    assert 0.0 < s.acceptance_rate <= 1.0
    nstep, ntherm, ndecor, L = (int(v) for v in f["hm_cfg"])
    s = Hamiltonian(nwalkers=len(start), nstep=nstep, step_size=float(f["hm_step"][0]), L=L, ntherm=ntherm,
                    ndecor=ndecor, nelec=wf.nelec, ndim=3, init=mol.domain("normal"), cuda=True)
    torch.manual_seed(int(f["hm_seed"][0]))
    out = s(wf.pdf, pos=start.clone(), with_tqdm=False)
    assert out.shape == f["hm_pos"].shape and C.scaled_err(out.detach(), f["hm_pos"]) < 1e-9
    # a plain callable is differentiated with autograd, like the reference does
    s = Hamiltonian(nwalkers=64, nstep=3, step_size=0.1, L=3, nelec=1, ndim=3, init={"min": -1, "max": 1}, cuda=True)
    torch.manual_seed(0)
    out = s(lambda x: torch.exp(-(x * x).sum(1)), with_tqdm=False)
    assert out.shape == (64, 3) and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("name", ["lih_ground", "h2_single22", "lih_sd22", "lih_nojastrow", "lih_sto", "lih_sto_pure",
                                  "lih_gto_kr", "lih_adf_sd22"])
def test_specialised_gradient_kernel(name, monkeypatch):
    """spec_grad_psi (generated inverses / CI weights / electron loop) against the generic
    fused_kernel<MODE_GRAD> and the oracle: grad psi and grad psi^2 (slater_jastrow.py:346-447)."""
    g = C.load(name)
    monkeypatch.setenv("QMCB_JIT", "0")
    mol, wf0 = C.build_wf(g)
    pos, _ = _thermalised(wf0, mol, 3001)
    g0, p0 = wf0.gradients_jacobi(pos), wf0.gradients_jacobi(pos, pdf=True)
    monkeypatch.setenv("QMCB_JIT", "2")
    mol, wf1 = C.build_wf(g)
    assert wf1._handle.info(13) == 1
    g1, p1 = wf1.gradients_jacobi(pos), wf1.gradients_jacobi(pos, pdf=True)
    assert C.scaled_err(g1, g0) < 1e-12 and C.scaled_err(p1, p0) < 1e-12
    _, P = C.oracle_params(g)
    assert C.scaled_err(g1[:128], orc.grad_psi(P, pos[:128].cpu())) < RTOL
    assert C.scaled_err(p1[:128], orc.grad_psi(P, pos[:128].cpu(), pdf=True)) < RTOL
    # autograd w.r.t. positions goes through the same kernel
    x = pos[:64].clone().requires_grad_(True)
    wf1(x).sum().backward()
    assert C.scaled_err(x.grad, g1[:64]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["lih", "h2"])
def test_gto2sto_wave_function_on_the_cuda_path(key, double_default):
    """The wave function returned by gto2sto() (slater_jastrow.py:649-733: single-zeta sto_pure basis fitted
    to the Gaussian AOs) evaluates on the CUDA kernels and matches the reference's own gto2sto() object
    on the same walkers (tests/golden/gto2sto.npz)."""
    from qmctorch_b200.wavefunction import SlaterJastrow
    g = np.load(os.path.join(C.GOLDEN, "gto2sto.npz"))
    wf = SlaterJastrow(fixture_molecule(key), configs="ground_state", cuda=True).gto2sto()
    assert wf.cuda and wf.ao.radial_type == "sto_pure"
    with torch.no_grad():
        wf.jastrow.jastrow_kernel.weight.fill_(0.8)
    pos = torch.tensor(g[key + "_pos"]).cuda()
    assert C.rel_err(wf(pos), g[key + "_psi"]) < RTOL
    assert C.rel_err(wf.local_energy(pos), g[key + "_eloc"]) < RTOL
    assert C.scaled_err(wf.gradients_jacobi(pos, sum_grad=False).reshape(len(pos), -1), g[key + "_gpsi"]) < RTOL


def test_c_program_evaluates_psi_on_the_device(tmp_path):
    """The pure-C client of include/qmcb.h (tests/c/abi_smoke.c) on a device: H2 STO-3G psi against the
    closed form evaluated in the C program itself, E_L finite."""
    import subprocess
    from test_host import _build_abi_smoke
    r = subprocess.run([_build_abi_smoke(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "ABI_OK device" in r.stdout


@pytest.mark.parametrize("name", ["lih_ground", "h2_single22", "lih_sd22", "lih_een", "lih_sto", "lih_cas24",
                                  "lih_nojastrow"])
def test_specialised_backward_matches_tile_backward_and_oracle(name):
    """spec_backward (one walker per thread, register accumulators: the Jastrow / MO / CI gradients of
    BASELINE config 3) against backward_kernel (DMMA tile kernel, taken when basis-parameter gradients
    are asked for too) and against the oracle's autograd on fresh walkers; bitwise reproducible."""
    g = C.load(name)
    mol, wf = C.build_wf(g)
    mol_o, P = C.oracle_params(g)
    pos, _ = _thermalised(wf, mol, 3001)
    torch.manual_seed(2)
    wgt = torch.randn(pos.shape[0], dtype=torch.float64, device="cuda")
    want = {"mo_modifier", "ci", "jee_w", "jen_w"}
    a = wf._psi_backward(pos, wgt, want)
    b = wf._psi_backward(pos, wgt, want)
    all_spec = wf._psi_backward(pos, wgt, None)             # every parameter: spec_backward_all
    all_spec2 = wf._psi_backward(pos, wgt, None)
    os.environ["QMCB_BWD_SPEC"] = "0"
    try:
        full = wf._psi_backward(pos, wgt, None)             # the DMMA tile kernel
    finally:
        del os.environ["QMCB_BWD_SPEC"]
    keys = ["mo_modifier", "ci"] + (["jee_w"] if wf._jee is not None else []) + (["jen_w"] if wf._jen is not None else [])
    for k in keys:
        assert torch.equal(a[k], b[k]), k
        ref = full[k]
        err = float((a[k] - ref).abs().max() / max(float(ref.abs().max()), 1e-300))
        assert err < 1e-11, (k, err)
    for k in keys + ["bas_exp", "bas_coeffs"]:
        assert torch.equal(all_spec[k], all_spec2[k]), k     # bitwise reproducible
        ref = full[k]
        err = float((all_spec[k] - ref).abs().max() / max(float(ref.abs().max()), 1e-300))
        assert err < 1e-11, (k, err)
    # the oracle: psi.backward(weight) by autograd on the CPU restatement
    leaves = {}
    for nme in ("mo_modifier", "ci", "jastrow_weight", "en_weight"):
        t = getattr(P, nme, None)
        if t is not None:
            t = t.detach().clone().requires_grad_(True)
            setattr(P, nme, t)
            leaves[nme] = t
    orc.psi(P, pos.cpu()).backward(wgt.cpu().reshape(-1, 1))
    for k, ok in (("mo_modifier", "mo_modifier"), ("ci", "ci"), ("jee_w", "jastrow_weight"), ("jen_w", "en_weight")):
        if ok in leaves and k in keys:
            ref = leaves[ok].grad.reshape(a[k].shape)
            err = float((a[k].cpu() - ref).abs().max() / max(float(ref.abs().max()), 1e-300))
            assert err < RTOL, (k, err)
    assert wf._handle.info(15) == 1


@pytest.mark.parametrize("key", ["lih_sph", "lih_sph_gto"])
def test_spherical_harmonics_basis_against_reference(key, monkeypatch):
    """harmonics_type = "sph" (real spherical harmonics up to l = 2, d shell on Li; Slater and Gaussian radial
    parts with r^n): AO values, psi, E_L, grad psi, accept decisions and the Jastrow / MO / CI gradients on the
    CUDA path against the reference (tests/golden/sph.npz; its E_L is the autograd-Hessian kinetic energy - the
    reference's Jacobi path raises on spherical harmonics) and against the oracle; generic and specialised
    kernels; basis-parameter gradients go through the flat-primitive adjoint kernel (an AO is a sum of several monomials)."""
    from test_oracle_golden import _sph_case
    from qmctorch_b200.wavefunction import SlaterJastrow
    g, mol, P = _sph_case(key)
    pos = _dev(g[key + "_pos"])
    el_ref = torch.tensor(g[key + "_eloc"])
    for jit in ("0", "2"):
        monkeypatch.setenv("QMCB_JIT", jit)
        wf = SlaterJastrow(mol, configs="single_double(2,2)", cuda=True)
        with torch.no_grad():
            wf.mo.mo_modifier.copy_(torch.tensor(g[key + "_mo_modifier"]))
            wf.fc.weight.copy_(torch.tensor(g[key + "_ci"]))
            wf.jastrow.jastrow_kernel.weight.fill_(0.8)
        assert wf._handle.info(13) == (0 if jit == "0" else 1)
        assert C.scaled_err(wf.ao(pos[:8]), g[key + "_ao"]) < RTOL
        ao, dao, d2ao = wf.ao(pos[:8], derivative=[0, 1, 2])
        o_ao, o_dao, o_d2 = orc.ao_all(P, pos[:8].cpu())
        assert C.scaled_err(ao, o_ao) < RTOL and C.scaled_err(dao, o_dao) < RTOL and C.scaled_err(d2ao, o_d2) < RTOL
        assert C.rel_err(wf(pos), g[key + "_psi"]) < RTOL
        e = wf.local_energy(pos).cpu()
        assert float(((e - el_ref).abs() / el_ref.abs().clamp(min=1.0)).max()) < 1e-9       # reference: autograd Hessian
        eo = orc.local_energy(P, pos.cpu())
        assert float(((e - eo).abs() / eo.abs().clamp(min=1.0)).max()) < RTOL               # oracle: Jacobi
        assert C.scaled_err(wf.gradients_jacobi(pos), orc.grad_psi(P, pos.cpu())) < RTOL
        # teacher-forced Metropolis decisions
        L = _lib_mod().lib()
        for it in range(g[key + "_mh_disp"].shape[0]):
            x = _dev(g[key + "_mh_pos"][it])
            disp, tau = _dev(g[key + "_mh_disp"][it]), _dev(g[key + "_mh_tau"][it])     # (kept alive until the sync)
            fx = (wf(x).reshape(-1) ** 2).detach().contiguous()
            acc = torch.zeros(x.shape[0], dtype=torch.uint8, device="cuda")
            _lib_mod().check(L.qmcb_metropolis_step(
                wf._handle.plan(), _lib_mod().ptr(x), _lib_mod().ptr(fx), x.shape[0], _lib_mod().ptr(disp),
                _lib_mod().ptr(tau), None, -1, 1, 1.0, 1e-16, 0, 0, _lib_mod().ptr(acc), None,
                _lib_mod().stream_ptr(x.device)), "qmcb_metropolis_step")
            torch.cuda.synchronize()
            assert np.array_equal(acc.cpu().numpy().astype(bool), g[key + "_mh_acc"][it])
            assert np.array_equal(x.cpu().numpy(), g[key + "_mh_pos"][it + 1])
        # psi.backward(weight) with the reference's weights
        psi_ref = torch.tensor(g[key + "_psi"])
        wgt = (2.0 / len(psi_ref) * (el_ref - el_ref.mean()) / psi_ref).reshape(-1).cuda()
        got = wf._psi_backward(pos, wgt, {"mo_modifier", "ci", "jee_w"})
        for k, ref in (("mo_modifier", g[key + "_grad_mo_modifier"]), ("ci", g[key + "_grad_ci"]),
                       ("jee_w", g[key + "_grad_jastrow_weight"])):
            ref = torch.tensor(ref)
            err = float((got[k].cpu().reshape(ref.shape) - ref).abs().max() / ref.abs().max())
            assert err < 1e-9, (k, err)
        # every gradient at once: the basis-parameter part of a multi-monomial basis comes from the flat-primitive
        # adjoint kernel (checked against the oracle in tests/test_gpu_vjp.py), the rest must not change
        full = wf._psi_backward(pos, wgt, None)
        assert full["bas_exp"].shape == wf.ao.bas_exp.shape and bool(torch.isfinite(full["bas_exp"]).all())
        for k in ("mo_modifier", "ci", "jee_w"):
            assert torch.equal(full[k], got[k])


def _lib_mod():
    from qmctorch_b200 import _lib
    return _lib


@pytest.mark.parametrize("name,nw,jastrow", [("c4h6_ground", 4_000_000, None), ("h2o_cas44_een", 250_000, "een")])
def test_full_size_properties_large_systems(name, nw, jastrow):
    """BASELINE configs 5 (C4H6 DZP, 4e6 walkers) and 4 (H2O cas(4,4) + e-e-n Jastrow, 2.5e5 walkers) at
    full size through the warp-tile kernels: size-independent properties - finite results, the E_L kernel's
    psi equals the psi kernel's, antisymmetry under exchange of two same-spin electrons (psi flips sign, E_L
    is unchanged), tiling independence (a ragged slice evaluated alone is BITWISE the slice of the full run),
    bitwise determinism, and the oracle on a sample."""
    g = C.load(name)
    mol, wf = C.build_wf(g)
    _, P = C.oracle_params(g)
    assert wf._handle.info(15) == 2
    pos, sampler = _thermalised(wf, mol, nw, nstep=12, step=float(g["step"]))
    ne3 = 3 * wf.nelec
    assert pos.shape == (nw, ne3)
    e, p2, k = wf._eloc(pos, want_psi=True, want_ekin=True)
    psi = wf(pos)
    assert bool(torch.isfinite(e).all()) and bool(torch.isfinite(psi).all())
    assert C.rel_err(p2, psi) < 1e-12
    e2, _, _ = wf._eloc(pos)
    assert torch.equal(e, e2)                                        # deterministic
    lo, n = 12346, 1001                                              # ragged, and shifted against the warp tiles
    es, ps, _ = wf._eloc(pos[lo:lo + n].contiguous(), want_psi=True)
    assert torch.equal(es, e[lo:lo + n]) and torch.equal(ps, p2[lo:lo + n])
    sub = pos[:200_000]
    sw = sub.clone()
    sw[:, 0:3], sw[:, 3:6] = sub[:, 3:6], sub[:, 0:3]                # two spin-up electrons
    assert float(((wf(sw) + psi[:200_000]).abs() / psi[:200_000].abs()).max()) < 1e-8
    d = (wf.local_energy(sw) - e[:200_000]).abs() / e[:200_000].abs().clamp(min=1.0)
    assert float(d.max()) < 1e-6
    ns = 24 if jastrow is None else 48
    cpu = pos[:ns].cpu()
    assert C.rel_err(psi[:ns], orc.psi(P, cpu)) < RTOL
    eo = orc.local_energy(P, cpu)
    assert float(((e[:ns].cpu() - eo).abs() / eo.abs().clamp(min=1.0)).max()) < RTOL


def test_fused_statistics_on_concurrent_streams_share_a_plan():
    """qmcb_local_energy_stats finishes its reduction in the CTA that arrives last (one kernel per E_L +
    statistics step); the arrival counter is per (plan, stream), so two streams may drive ONE plan
    concurrently (own workspace / outputs each): sums are bitwise those of sequential calls."""
    from qmctorch_b200 import _lib
    g = C.load("lih_ground")
    mol, wf = C.build_wf(g)
    L = _lib.lib()
    plan = wf._handle.plan()
    W = 300_000
    ens = [_thermalised(wf, mol, W, nstep=5, seed=s)[0] for s in (1, 2)]
    ref = []
    for x in ens:
        e, o4 = wf.local_energy_stats(x)
        torch.cuda.synchronize()
        ref.append(o4.clone())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bufs = [(torch.empty(W, dtype=torch.float64, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda"),
             torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8, device="cuda")) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(25):
        for k in range(2):
            e, o4, ws = bufs[k]
            _lib.check(L.qmcb_local_energy_stats(plan, _lib.ptr(ens[k]), W, _lib.ptr(e), None, None, _lib.ptr(o4),
                                                 _lib.ptr(ws), ctypes_stream(streams[k])), "local_energy_stats")
        torch.cuda.synchronize()
        for k in range(2):
            assert torch.equal(bufs[k][1], ref[k]), (rep, k)


def ctypes_stream(s):
    import ctypes
    return ctypes.c_void_p(s.cuda_stream)
