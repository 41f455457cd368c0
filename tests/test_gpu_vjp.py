"""GPU: the adjoint of the local energy (qmcb_local_energy_backward) against the UNMODIFIED reference's autograd
through WaveFunction.local_energy / psi (tests/golden/vjp.npz, written by oracle/make_golden_vjp.py), and the two
reference features built on it: Solver.configure(grad="auto") and Solver.compute_forces.

Tolerance (BASELINE.json north_star): 1e-10 relative in FP64, taken against the largest entry of each gradient
array (the arrays carry structural zeros)."""
import os

import numpy as np
import pytest
import torch

import _cases as C

pytestmark = pytest.mark.gpu
RTOL = 1e-10

GOLD = dict(np.load(os.path.join(C.GOLDEN, "vjp.npz")))
CASES = sorted({k.split("/")[0] for k in GOLD})
LEAVES = ("atom_coords", "bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w")


def _leaf(wf, name):
    return {"atom_coords": wf.ao.atom_coords, "bas_exp": wf.ao.bas_exp, "bas_coeffs": wf.ao.bas_coeffs,
            "mo_modifier": wf.mo.mo_modifier, "ci": wf.fc.weight,
            "jee_w": wf._jee.jastrow_kernel.weight if wf._jee is not None else None,
            "jen_w": wf._jen.jastrow_kernel.weight if wf._jen is not None else None}[name]


def _err(a, ref, floor=1e-4):
    """max |a - ref| / max |ref|; ``floor`` bounds the denominator from below for gradients that vanish
    analytically (the CI coefficient of a one-determinant expansion: two O(10) terms cancel to round-off)."""
    a = torch.as_tensor(a).detach().cpu().double().reshape(-1)
    ref = torch.as_tensor(ref).detach().cpu().double().reshape(-1)
    return float((a - ref).abs().max() / ref.abs().max().clamp(min=floor))


def _setup(name):
    g = C.load(name)
    mol, wf = C.build_wf(g)
    n = int(GOLD[name + "/n"][0])
    pos = torch.as_tensor(g["pos"][:n]).cuda()
    return g, wf, pos


@pytest.mark.parametrize("name", CASES)
def test_adjoint_of_local_energy_and_psi_matches_reference_autograd(name):
    """sum_w wE dE_L/dtheta and sum_w wP dpsi/dtheta for every leaf, straight through the C ABI."""
    g, wf, pos = _setup(name)
    wE = torch.as_tensor(GOLD[name + "/wE"]).cuda()
    wP = torch.as_tensor(GOLD[name + "/wP"]).cuda()
    assert C.rel_err(wf.local_energy(pos), GOLD[name + "/eloc"].reshape(-1, 1)) < RTOL
    want = {n for n in LEAVES if name + "/gE_" + n in GOLD}
    if not wf.ao.contract:
        # an uncontracted basis: the reference multiplies bas_coeffs into the gradient channel only
        # (atomic_orbitals.py:344 vs :236-249), a quirk the plan refuses to mirror (coefficients must be 1)
        want.discard("bas_coeffs")
    gE = wf._eloc_backward(pos, wE, None, want)
    gP = wf._eloc_backward(pos, None, wP, want)
    both = wf._eloc_backward(pos, wE, wP, want)
    for n in sorted(want):
        refE, refP = GOLD[name + "/gE_" + n], GOLD[name + "/gP_" + n]
        assert _err(gE[n], refE) < RTOL, (n, _err(gE[n], refE))
        assert _err(gP[n], refP) < RTOL, (n, _err(gP[n], refP))
        assert _err(both[n], refE + refP) < 2 * RTOL, n
    # deterministic: bitwise the same result on a second call
    again = wf._eloc_backward(pos, wE, wP, want)
    for n in want:
        assert torch.equal(again[n], both[n])


@pytest.mark.parametrize("name", ["lih_een", "h2o_cas44"])
def test_local_energy_is_differentiable_through_autograd(name):
    """wf.local_energy carries an autograd node (what grad='auto' and compute_forces rely on); the psi node
    agrees with qmcb_psi_backward for the parameters both kernels serve."""
    g, wf, pos = _setup(name)
    wE = torch.as_tensor(GOLD[name + "/wE"]).cuda()
    wP = torch.as_tensor(GOLD[name + "/wP"]).cuda()
    wf.ao.bas_coeffs.requires_grad = True
    wf.ao.atom_coords.requires_grad = True
    wf.atom_coords_grad = True
    names = [n for n in LEAVES if name + "/gE_" + n in GOLD]
    leaves = [_leaf(wf, n) for n in names]
    e = wf.local_energy(pos)
    assert e.requires_grad
    gr = torch.autograd.grad((e.reshape(-1) * wE).sum(), leaves)
    for n, a in zip(names, gr):
        assert a.shape == _leaf(wf, n).shape
        assert _err(a, GOLD[name + "/gE_" + n]) < RTOL, n
    psi = wf(pos)
    gp = torch.autograd.grad((psi.reshape(-1) * wP).sum(), leaves)
    for n, a in zip(names, gp):
        assert _err(a, GOLD[name + "/gP_" + n]) < RTOL, n
    with torch.no_grad():
        assert not wf.local_energy(pos).requires_grad


@pytest.mark.parametrize("key", ["lih_sph", "lih_sph_gto"])
def test_spherical_harmonics_adjoint_against_oracle(key):
    """Real spherical harmonics (l = 2: an AO is a sum of several monomials).  The reference's Jacobi kinetic energy
    raises on such bases, so the checker is autograd through the oracle (pinned on the cartesian cases by
    test_oracle_local_energy_adjoint); exponent derivatives of the monomials of one primitive add up.  The same
    kernel serves the basis-parameter gradients of psi.backward for these bases (qmcb_psi_backward)."""
    import sj_oracle as orc
    from test_oracle_golden import _sph_case
    from qmctorch_b200.wavefunction import SlaterJastrow
    g, mol, P = _sph_case(key)
    pos = torch.as_tensor(g[key + "_pos"][:12]).cuda()
    wf = SlaterJastrow(mol, configs="single_double(2,2)", cuda=True)
    with torch.no_grad():
        wf.mo.mo_modifier.copy_(torch.tensor(g[key + "_mo_modifier"]))
        wf.fc.weight.copy_(torch.tensor(g[key + "_ci"]))
        wf.jastrow.jastrow_kernel.weight.fill_(0.8)
    gen = torch.Generator().manual_seed(5)
    wE = torch.rand(12, generator=gen, dtype=torch.float64) - 0.3
    wP = torch.rand(12, generator=gen, dtype=torch.float64) - 0.3
    want = {"atom_coords", "bas_exp", "mo_modifier", "ci", "jee_w"}
    ix = torch.as_tensor(wf.ao.expand_index, dtype=torch.long)

    def fold(d):
        d = dict(d)
        d["bas_exp"] = torch.zeros(wf.ao.nbas, dtype=torch.float64).index_add_(0, ix, d["bas_exp"])
        return d
    for we, wp in ((wE, None), (None, wP), (wE, wP)):
        got = wf._eloc_backward(pos, None if we is None else we.cuda(), None if wp is None else wp.cuda(), want)
        ref = fold(orc.local_energy_adjoint(P, pos.cpu(), w_eloc=we, w_psi=wp))
        for n in sorted(want):
            assert got[n].shape == _leaf(wf, n).shape
            assert _err(got[n], ref[n]) < RTOL, (n, _err(got[n], ref[n]))
    # psi.backward with the basis exponents trainable
    wf.ao.bas_exp.requires_grad = True
    psi = wf(pos)
    psi.backward(wP.cuda().reshape(psi.shape))
    ref = fold(orc.local_energy_adjoint(P, pos.cpu(), w_psi=wP))
    assert _err(wf.ao.bas_exp.grad, ref["bas_exp"]) < RTOL
    assert _err(wf.mo.mo_modifier.grad, ref["mo_modifier"]) < RTOL


@pytest.mark.parametrize("name,nw", [("lih_een", 300_001), ("h2o_cas44", 20_011)])
def test_adjoint_is_additive_over_walkers_at_full_size(name, nw):
    """Size-independent property at sizes the oracle cannot reach: the adjoint is a sum over walkers, so any
    split of a large ensemble (several Jastrow-operator chunks of 131072 walkers, ragged tails, several trips of
    the persistent walker loop, partially filled lane groups) must add up to the adjoint of the whole - and the
    first walkers, evaluated alone, must reproduce the reference-pinned small case."""
    g, wf, pos0 = _setup(name)
    n0 = pos0.shape[0]
    gen = torch.Generator().manual_seed(11)
    reps = -(-nw // g["pos"].shape[0])
    base = torch.as_tensor(g["pos"]).repeat(reps, 1)[:nw]
    pos = (base + 0.05 * torch.randn(base.shape, generator=gen, dtype=torch.float64)).cuda()
    pos[:n0] = pos0
    wE = (torch.rand(nw, generator=gen, dtype=torch.float64) - 0.3).cuda()
    wP = (torch.rand(nw, generator=gen, dtype=torch.float64) - 0.3).cuda()
    wE[:n0] = torch.as_tensor(GOLD[name + "/wE"]).cuda()
    wP[:n0] = torch.as_tensor(GOLD[name + "/wP"]).cuda()
    want = {"atom_coords", "bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w"}
    whole = wf._eloc_backward(pos, wE, wP, want)
    cuts = [0, n0, 131072 + 5, nw] if nw > 140000 else [0, n0, 7777, nw]
    parts = [wf._eloc_backward(pos[a:b].contiguous(), wE[a:b].contiguous(), wP[a:b].contiguous(), want)
             for a, b in zip(cuts[:-1], cuts[1:])]
    for n in sorted(want):
        total = sum(p[n] for p in parts)
        scale = sum(p[n].abs() for p in parts).max().clamp(min=1e-300)
        assert bool(torch.isfinite(whole[n]).all()), n
        assert float((whole[n] - total).abs().max() / scale) < 1e-12, n
        ref = GOLD[name + "/gE_" + n] + GOLD[name + "/gP_" + n]
        assert _err(parts[0][n], ref) < 2 * RTOL, n
    again = wf._eloc_backward(pos, wE, wP, want)
    for n in want:
        assert torch.equal(again[n], whole[n])


def test_adjoint_entry_point_error_behaviour_and_empty_input():
    """C ABI contract of qmcb_local_energy_backward: both weights NULL or no workspace -> QMCB_EINVAL with a
    message, outputs untouched; W = 0 -> success and zero sums (V_nn contributes sum(w_eloc) = 0)."""
    import ctypes
    from qmctorch_b200 import _lib
    g, wf, pos = _setup("lih_een")
    L = _lib.lib()
    plan = wf._handle.plan()
    W = pos.shape[0]
    ws = torch.empty(int(L.qmcb_local_energy_backward_workspace_bytes(plan, W)), dtype=torch.uint8, device="cuda")
    wgt = torch.ones(W, dtype=torch.float64, device="cuda")
    out = torch.full((wf.natom, 3), 7.0, dtype=torch.float64, device="cuda")
    sp = _lib.stream_ptr(pos.device)
    args = lambda we, wp, n, w_: (plan, _lib.ptr(pos), we, wp, n, None, None, None, None, None, None, _lib.ptr(out),
                                  w_, sp)
    assert L.qmcb_local_energy_backward(*args(None, None, W, _lib.ptr(ws))) == -1
    assert b"qmcb_local_energy_backward" in L.qmcb_last_error()
    assert L.qmcb_local_energy_backward(*args(_lib.ptr(wgt), None, W, None)) == -1
    torch.cuda.synchronize()
    assert float(out.min()) == 7.0 and float(out.max()) == 7.0
    assert L.qmcb_local_energy_backward(*args(_lib.ptr(wgt), None, 0, _lib.ptr(ws))) == 0
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
    assert L.qmcb_local_energy_backward_workspace_bytes(plan, 0) > 0
    with pytest.raises(RuntimeError):
        wf._eloc_backward(pos, None, None, {"ci"})


def _solver(wf, mol, nw):
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    sampler = Metropolis(nwalkers=nw, nstep=10, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                         move={"type": "all-elec", "proba": "normal"}, cuda=True)
    return Solver(wf=wf, sampler=sampler, optimizer=torch.optim.SGD(wf.parameters(), lr=1e-3))


@pytest.mark.parametrize("name", ["lih_een", "lih_cas24"])
def test_solver_forces_and_grad_auto_match_reference(name):
    g = C.load(name)
    mol, wf = C.build_wf(g)
    n = int(GOLD[name + "/n"][0])
    pos = torch.as_tensor(g["pos"][:n]).cuda()
    solver = _solver(wf, mol, n)
    f = solver.compute_forces(pos)
    assert f.shape == wf.ao.atom_coords.shape
    assert _err(f, GOLD[name + "/forces"]) < RTOL
    assert _err(solver.compute_forces(pos, batch_size=n // 2, clip=2), GOLD[name + "/forces_clip"]) < RTOL
    assert not wf.atom_coords_grad and not wf.ao.atom_coords.requires_grad
    for loss in ("energy", "variance"):
        solver.configure(track=["local_energy"], loss=loss, grad="auto",
                         resampling={"mode": "update", "resample_every": 1, "nstep_update": 5})
        wf.zero_grad()
        val, eloc = solver.evaluate_gradient(pos)
        assert abs(float(val) - float(GOLD[name + "/auto_%s_loss" % loss][0])) < 1e-9 * max(1.0, abs(float(val)))
        seen = 0
        for pn, p in wf.named_parameters():
            key = name + "/auto_%s/%s" % (loss, pn)
            if key in GOLD:
                assert p.grad is not None, pn
                assert _err(p.grad, GOLD[key]) < RTOL, (loss, pn, _err(p.grad, GOLD[key]))
                seen += 1
        assert seen >= 4
    # sampling weights: walkers kept for two epochs - the second loss is weighted by (psi / psi0)^2 / sum and its
    # gradient passes through psi (qmcb_psi_backward) as well as through E_L (qmcb_local_energy_backward)
    solver.configure(track=["local_energy"], loss="energy", grad="auto",
                     resampling={"mode": "update", "resample_every": 2, "nstep_update": 5})
    assert solver.loss.use_weight
    wf.zero_grad()
    solver.evaluate_gradient(pos)
    solver.opt.step()
    wf.zero_grad()
    val, _ = solver.evaluate_gradient(pos)
    ref = float(GOLD[name + "/auto_weighted_loss"][0])
    assert abs(float(val) - ref) < 1e-9 * max(1.0, abs(ref))
    seen = 0
    for pn, p in wf.named_parameters():
        key = name + "/auto_weighted/%s" % pn
        if key in GOLD:
            assert _err(p.grad, GOLD[key]) < 10 * RTOL, (pn, _err(p.grad, GOLD[key]))
            seen += 1
    assert seen >= 4


@pytest.mark.parametrize("loss", ["energy", "variance"])
def test_grad_auto_is_the_derivative_of_the_loss(loss):
    """grad='auto' on a FIXED ensemble of 4096 LiH walkers: the loss is a deterministic function of the
    parameters, so a small step against the gradient the Solver left in .grad must change the loss by
    -eps |g| to first order (central finite difference of the loss along the gradient direction)."""
    g = C.load("lih_een")
    mol, wf = C.build_wf(g)
    from qmctorch_b200.sampler import Metropolis
    torch.manual_seed(0)
    sampler = Metropolis(nwalkers=4096, nstep=300, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                         move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=3)
    pos = sampler(wf.pdf, with_tqdm=False).detach()
    solver = _solver(wf, mol, 4096)
    solver.configure(track=["local_energy"], loss=loss, grad="auto",
                     resampling={"mode": "never", "resample_every": 1, "nstep_update": 25})
    wf.zero_grad()
    solver.evaluate_gradient(pos)
    params = [p for p in wf.parameters() if p.grad is not None]
    assert len(params) >= 4
    grads = [p.grad.detach().clone() for p in params]
    norm = float(torch.sqrt(sum((x ** 2).sum() for x in grads)))
    eps = 1e-5

    def loss_at(step):
        with torch.no_grad():
            for p, x in zip(params, grads):
                p.add_(x, alpha=step / norm)
            val = float(solver.loss(pos, no_grad=True)[0])
            for p, x in zip(params, grads):
                p.sub_(x, alpha=step / norm)
        return val

    slope = (loss_at(eps) - loss_at(-eps)) / (2 * eps)
    assert abs(slope - norm) < 1e-5 * norm, (slope, norm)


def test_geometry_step_through_the_solver():
    """The geometry-optimisation step of the reference's ASE optimiser (ase/optimizer/torch_optim.py:126-133):
    set_params_requires_grad(wf_params=False, geo_params=True) + evaluate_grad_auto leaves d<E_L>/dR in
    ao.atom_coords.grad and nothing else; an SGD step moves the atoms and the device tables follow."""
    g = C.load("lih_een")
    mol, wf = C.build_wf(g)
    n = int(GOLD["lih_een/n"][0])
    pos = torch.as_tensor(g["pos"][:n]).cuda()
    solver = _solver(wf, mol, n)
    solver.configure(track=["local_energy"], loss="energy", grad="auto",
                     resampling={"mode": "update", "resample_every": 1, "nstep_update": 5})
    solver.set_params_requires_grad(wf_params=False, geo_params=True)
    assert wf.atom_coords_grad and wf.ao.atom_coords.requires_grad and not wf.mo.mo_modifier.requires_grad
    opt = torch.optim.SGD(wf.parameters(), lr=1e-2)
    solver.opt = opt
    wf.zero_grad()
    loss, eloc = solver.evaluate_gradient(pos)
    grad = wf.ao.atom_coords.grad.clone()
    ref = wf._eloc_backward(pos, torch.full((n,), 1.0 / n, dtype=torch.float64, device="cuda"), None,
                            {"atom_coords"})["atom_coords"]
    assert _err(grad, ref) < 1e-12 and wf.mo.mo_modifier.grad is None
    e0 = wf.local_energy(pos).detach().clone()
    before = wf.geometry(None)
    opt.step()
    after = wf.geometry(None, convert_to_angs=True)
    assert before != wf.geometry(None) and abs(after[0][2] - wf.geometry(None)[0][2] * 0.529177) < 1e-12
    with torch.no_grad():
        e1 = wf.local_energy(pos)
    assert float((e1 - e0).abs().max()) > 1e-6          # the plan follows the moved atoms
    solver.set_params_requires_grad(wf_params=True, geo_params=False)
    assert not wf.atom_coords_grad


def test_three_body_weights_are_refused_by_grad_auto():
    g = C.load("lih_sd22_een3")
    mol, wf = C.build_wf(g)
    solver = _solver(wf, mol, 8)
    solver.configure(track=["local_energy"], loss="energy", grad="auto",
                     resampling={"mode": "update", "resample_every": 1, "nstep_update": 5})
    with pytest.raises(NotImplementedError):
        solver.evaluate_gradient(torch.as_tensor(g["pos"][:8]).cuda())
