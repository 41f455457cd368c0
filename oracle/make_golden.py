"""Pin the oracle against the unmodified reference and write tests/golden/*.npz.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
For every case: build the reference SlaterJastrow on a fixture molecule, sample
walkers with the reference Metropolis, evaluate the hot path with the reference,
assert oracle/sj_oracle.py agrees, and store inputs + reference outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import ref_shim  # noqa: E402

ref_shim.load_reference()

from qmctorch.wavefunction import SlaterJastrow  # noqa: E402
from qmctorch.sampler import Metropolis  # noqa: E402
from qmctorch.solver import Solver  # noqa: E402
from qmctorch.wavefunction.jastrows.elec_elec import (  # noqa: E402
    JastrowFactor as JastrowEE, PadeJastrowKernel as PadeEE)
from qmctorch.wavefunction.jastrows.elec_nuclei import (  # noqa: E402
    JastrowFactor as JastrowEN, PadeJastrowKernel as PadeEN)
from qmctorch.wavefunction.jastrows.elec_elec_nuclei import (  # noqa: E402
    JastrowFactor as JastrowEEN, BoysHandyJastrowKernel as BoysHandy)

import sj_oracle as orc  # noqa: E402
from qmctorch_b200.molecules import fixture_molecule  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

# name, molecule key, configs, jastrow, nwalkers, thermalisation steps, step size, init
CASES = [
    ("h2_single22", "h2", "single(2,2)", "ee", 256, 200, 0.5, "normal"),
    ("h2_ground", "h2", "ground_state", "ee", 128, 100, 0.5, "normal"),
    ("lih_ground", "lih", "ground_state", "ee", 512, 200, 0.3, "normal"),
    ("lih_nojastrow", "lih", "ground_state", None, 64, 100, 0.3, "normal"),
    ("lih_sd22", "lih", "single_double(2,2)", "ee", 128, 100, 0.3, "normal"),
    ("lih_cas24", "lih", "cas(2,4)", "ee", 128, 100, 0.3, "normal"),
    ("lih_een", "lih", "ground_state", "ee+en", 128, 100, 0.3, "normal"),
    ("h2o_ground", "h2o", "ground_state", "ee", 96, 100, 0.15, "atomic"),
    ("h2o_cas44", "h2o", "cas(4,4)", "ee+en", 48, 100, 0.15, "atomic"),
    ("c4h6_ground", "c4h6", "ground_state", "ee", 16, 60, 0.05, "atomic"),
    # the other radial forms (ADF-style uncontracted bases)
    ("lih_sto", "lih_sto", "single_double(2,2)", "ee", 96, 100, 0.3, "normal"),
    ("lih_sto_pure", "lih_sto_pure", "ground_state", "ee", 64, 100, 0.3, "normal"),
    ("lih_gto_kr", "lih_gto_kr", "ground_state", "ee+en", 64, 100, 0.3, "normal"),
    # real ADF results read from the reference's tests/hdf5/*.hdf5 (tools/hdf5_to_fixture.py)
    ("lih_adf_sd22", "lih_adf", "single_double(2,2)", "ee", 96, 100, 0.3, "normal"),
    ("co2_adf_ground", "co2_adf", "ground_state", "ee", 16, 60, 0.05, "atomic"),
    # three-body Boys-Handy term (BASELINE config 4: CAS + e-e-n Jastrow)
    ("lih_sd22_een3", "lih", "single_double(2,2)", "ee+en+een", 96, 100, 0.3, "normal"),
    ("h2o_cas44_een", "h2o", "cas(4,4)", "ee+een", 32, 100, 0.15, "atomic"),
]


def rel(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float(((a - b).abs() / b.abs().clamp(min=1e-300)).max())


def relmax(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))


def build(case):
    name, key, configs, jast, nw, ntherm, step, init = case
    mol = fixture_molecule(key)
    if jast == "ee":
        j = "default"
    elif jast is None:
        j = None
    elif jast == "ee+en":
        j = [JastrowEE(mol, PadeEE), JastrowEN(mol, PadeEN)]
    elif jast == "ee+en+een":
        j = [JastrowEE(mol, PadeEE), JastrowEN(mol, PadeEN), JastrowEEN(mol, BoysHandy)]
    else:
        j = [JastrowEE(mol, PadeEE), JastrowEEN(mol, BoysHandy)]
    wf = SlaterJastrow(mol, configs=configs, jastrow=j, include_all_mo=True)
    return mol, wf


def main(only=None):
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        name, key, configs, jast, nw, ntherm, step, init = case
        if only and name not in only:
            continue
        torch.manual_seed(1234)
        np.random.seed(1234)
        mol, wf = build(case)
        sampler = Metropolis(nwalkers=nw, nstep=ntherm, step_size=step, nelec=wf.nelec,
                             ndim=3, init=mol.domain(init),
                             move={"type": "all-elec", "proba": "normal"})
        pos = sampler(wf.pdf, with_tqdm=False).detach().clone()
        # perturb the trainable parameters so gradients/values are not at a symmetric point
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            wf.mo.mo_modifier.mul_(1 + 0.05 * (torch.rand(wf.mo.mo_modifier.shape, generator=g) - 0.5))
            if wf.nci > 1:
                wf.fc.weight.add_(0.2 * (torch.rand(wf.fc.weight.shape, generator=g) - 0.5))
            if jast == "ee":
                wf.jastrow.jastrow_kernel.weight.fill_(0.8)
            elif jast is not None:
                wf.jastrow.jastrow_terms[0].jastrow_kernel.weight.fill_(0.8)
                if "+en" in jast:
                    wf.jastrow.jastrow_terms[1].jastrow_kernel.weight.data.fill_(1.3)
                if jast.endswith("een"):
                    bh = wf.jastrow.jastrow_terms[-1].jastrow_kernel
                    bh.weight_num.data.copy_(0.05 + 0.3 * torch.rand(1, 2, 5, generator=g))
                    bh.weight_denom.data.copy_(0.5 + torch.rand(1, 2, 5, generator=g))
                    bh.fc.weight.data.copy_(torch.rand(1, 5, generator=g) - 0.5)

        has_en = jast is not None and "+en" in jast
        has_een = jast is not None and jast.endswith("een")
        P = orc.make_params(mol, wf.configs,
                            jastrow_weight=None if jast is None else 0.8,
                            en_weight=1.3 if has_en else None)
        if has_een:
            bh = wf.jastrow.jastrow_terms[-1].jastrow_kernel
            P.een = dict(num=bh.weight_num.detach().clone(), denom=bh.weight_denom.detach().clone(),
                         fc=bh.fc.weight.detach().clone())
        P.mo_modifier = wf.mo.mo_modifier.detach().clone()
        P.ci = wf.fc.weight.detach().clone()

        out = dict(pos=pos.numpy(), mo_modifier=P.mo_modifier.numpy(), ci=P.ci.numpy(),
                   cfg_up=wf.configs[0].numpy(), cfg_down=wf.configs[1].numpy(),
                   jw=np.array([0.8 if jast else np.nan]),
                   enw=np.array([1.3 if has_en else np.nan]))
        if has_een:
            out.update(bh_num=P.een["num"].numpy(), bh_denom=P.een["denom"].numpy(), bh_fc=P.een["fc"].numpy())
        # the reference differentiates the three-body term by autograd: positions must carry grad
        pos_in = pos
        if has_een:
            pos = pos.clone().requires_grad_(True)
        with (torch.enable_grad() if has_een else torch.no_grad()):
            ao, dao, d2ao = wf.ao(pos, derivative=[0, 1, 2])
            psi = wf(pos)
            if jast is not None:
                J, dJ, d2J = wf.jastrow(pos, derivative=[0, 1, 2], sum_grad=False)
            ekin = wf.kinetic_energy(pos)
            eloc = wf.local_energy(pos)
            gpsi = wf.gradients_jacobi(pos)
            gpdf = wf.gradients_jacobi(pos, pdf=True)
        pos = pos_in
        ao, dao, d2ao, psi, ekin, eloc, gpsi, gpdf = [t.detach() for t in (ao, dao, d2ao, psi, ekin, eloc, gpsi, gpdf)]
        if jast is not None:
            J, dJ, d2J = J.detach(), dJ.detach(), d2J.detach()
        # --- oracle vs reference
        o_ao, o_dao, o_d2ao = orc.ao_all(P, pos)
        o_psi = orc.psi(P, pos)
        o_ekin = orc.kinetic_energy(P, pos)
        o_eloc = orc.local_energy(P, pos)
        o_g = orc.grad_psi(P, pos)
        o_gp = orc.grad_psi(P, pos, pdf=True)
        errs = dict(ao=relmax(o_ao, ao), dao=relmax(o_dao, dao), d2ao=relmax(o_d2ao, d2ao),
                    psi=rel(o_psi, psi), ekin=rel(o_ekin, ekin), eloc=rel(o_eloc, eloc),
                    gpsi=relmax(o_g, gpsi), gpdf=relmax(o_gp, gpdf))
        if jast is not None:
            oJ, odJ, od2J = orc.jastrow_all(P, pos)
            errs.update(J=rel(oJ, J), dJ=relmax(odJ, dJ), d2J=relmax(od2J, d2J))
            out.update(J=J.numpy(), dJ=dJ.numpy(), d2J=d2J.numpy())
        # keep AO tensors only for a slice (size)
        ns = min(8, pos.shape[0])
        out.update(ao=ao[:ns].numpy(), dao=dao[:ns].numpy(), d2ao=d2ao[:ns].numpy(),
                   psi=psi.numpy(), ekin=ekin.numpy(), eloc=eloc.numpy(),
                   gpsi=gpsi.numpy(), gpdf=gpdf.numpy())

        # --- parameter gradients through the reference solver (grad="manual")
        opt = torch.optim.SGD(wf.parameters(), lr=0.0)
        solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
        solver.configure(track=["local_energy"], loss="energy", grad="manual")
        opt.zero_grad()
        solver.evaluate_grad_manual(pos.clone().requires_grad_(True) if has_een else pos.clone())
        ref_g = dict(mo_modifier=wf.mo.mo_modifier.grad.clone(), ci=wf.fc.weight.grad.clone(),
                     bas_exp=wf.ao.bas_exp.grad.clone())
        if wf.ao.bas_coeffs.grad is not None:      # uncontracted bases never use the coefficients in psi
            ref_g["bas_coeffs"] = wf.ao.bas_coeffs.grad.clone()
        if jast == "ee":
            ref_g["jastrow_weight"] = wf.jastrow.jastrow_kernel.weight.grad.clone()
        elif jast is not None:
            ref_g["jastrow_weight"] = wf.jastrow.jastrow_terms[0].jastrow_kernel.weight.grad.clone()
        if has_een:
            bh = wf.jastrow.jastrow_terms[-1].jastrow_kernel
            ref_g["een_num"] = bh.weight_num.grad.clone()
            ref_g["een_denom"] = bh.weight_denom.grad.clone()
            ref_g["een_fc"] = bh.fc.weight.grad.clone()
        names = tuple(ref_g.keys()) + (("en_weight",) if has_en else ())
        og, _ = orc.param_grads(P, pos, names=names)
        for k, v in ref_g.items():
            if k == "ci" and wf.nci == 1:
                continue
            # (sum_w (E_L - <E_L>) = 0 makes the single-determinant ci gradient pure rounding noise)
            errs["g_" + k] = float((og[k] - v).abs().max() / max(float(v.abs().max()), 1e-6))
            out["grad_" + k] = v.numpy()
        if "en_weight" in og:
            out["grad_en_weight"] = og["en_weight"].numpy()   # reference keeps this one off the graph

        # --- Metropolis decisions, teacher forced, injected draws
        with torch.no_grad():
            gen = torch.Generator().manual_seed(99)
            sig = orc.proposal_sigma(step)
            nstep = 4
            cur = pos.clone()
            fx = wf.pdf(cur)
            fx[fx == 0] = 1e-16
            tr_disp, tr_tau, tr_acc, tr_fxn, tr_pos = [], [], [], [], [cur.numpy().copy()]
            ofx = (orc.psi(P, cur) ** 2).reshape(-1)
            for it in range(nstep):
                disp = torch.randn(cur.shape, generator=gen, dtype=torch.float64) * np.sqrt(sig)
                xn = cur + disp
                fxn = wf.pdf(xn)
                fxn[fxn == 0.0] = 1e-16
                df = fxn / fx
                torch.manual_seed(1000 + it)
                idx = sampler._accept(df.clone())
                torch.manual_seed(1000 + it)
                tau = torch.rand_like(df)
                # oracle from the same state
                npos, nfx, oacc, ofxn = orc.metropolis_step(P, cur, fx.clone(), disp, tau)
                assert bool((oacc == idx).all()), "oracle accept mismatch"
                errs["mh_fxn"] = max(errs.get("mh_fxn", 0.0), rel(ofxn, fxn))
                cur[idx, :] = xn[idx, :]
                fx[idx] = fxn[idx]
                fx[fx == 0] = 1e-16
                assert torch.equal(npos, cur)
                tr_disp.append(disp.numpy()); tr_tau.append(tau.numpy())
                tr_acc.append(idx.numpy()); tr_fxn.append(fxn.numpy())
                tr_pos.append(cur.numpy().copy())
            out.update(mh_disp=np.stack(tr_disp), mh_tau=np.stack(tr_tau), mh_acc=np.stack(tr_acc),
                       mh_fxn=np.stack(tr_fxn), mh_pos=np.stack(tr_pos))
        out["meta"] = np.array([name, key, configs, str(jast), str(step)])
        print("%-14s W=%4d " % (name, pos.shape[0])
              + " ".join("%s=%.1e" % kv for kv in errs.items()))
        bad = {k: v for k, v in errs.items() if not v < 5e-11}
        assert not bad, "oracle does not reproduce the reference: %r" % bad
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def spherical_fixture():
    """Spherical-harmonics bases (l <= 2, Slater and Gaussian radial parts with r^n; fixtures lih_sph,
    lih_sph_gto).  The reference's Jacobi kinetic energy cannot run on them (spherical_harmonics.py:226-245
    raises for derivative lists), so its wave function is built with kinetic="auto" (autograd Hessian of psi,
    wf_base.py:142-182): psi, AO values, E_L, the manual-gradient estimator and Metropolis decisions are the
    reference's; the oracle (spherical harmonics restated as cartesian monomials) is asserted against them
    -> tests/golden/sph.npz."""
    out = {}
    for key, nw in (("lih_sph", 96), ("lih_sph_gto", 64)):
        torch.manual_seed(2024)
        np.random.seed(2024)
        mol = fixture_molecule(key)
        wf = SlaterJastrow(mol, configs="single_double(2,2)", kinetic="auto", include_all_mo=True)
        sampler = Metropolis(nwalkers=nw, nstep=120, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                             move={"type": "all-elec", "proba": "normal"})
        pos = sampler(wf.pdf, with_tqdm=False).detach().clone()
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            wf.mo.mo_modifier.mul_(1 + 0.05 * (torch.rand(wf.mo.mo_modifier.shape, generator=g) - 0.5))
            wf.fc.weight.add_(0.2 * (torch.rand(wf.fc.weight.shape, generator=g) - 0.5))
            wf.jastrow.jastrow_kernel.weight.fill_(0.8)
        P = orc.make_params(mol, wf.configs, jastrow_weight=0.8)
        P.mo_modifier = wf.mo.mo_modifier.detach().clone()
        P.ci = wf.fc.weight.detach().clone()
        x = pos.clone().requires_grad_(True)
        psi = wf(x).detach()
        eloc = wf.local_energy(x).detach()
        with torch.no_grad():
            ao = wf.ao(pos)
        errs = dict(ao=relmax(orc.ao_values(P, pos), ao), psi=rel(orc.psi(P, pos), psi),
                    eloc=float(((orc.local_energy(P, pos) - eloc).abs() / eloc.abs().clamp(min=1.0)).max()))
        # manual energy-gradient estimator, solver.py:414-429 applied by hand: Solver.evaluate_grad_manual wraps
        # local_energy in no_grad, which kinetic="auto" cannot run under, so E_L above is reused
        wf.zero_grad()
        val = wf(pos)
        weight = 2.0 / len(val) * (eloc - eloc.mean()) / val.detach()
        val.backward(weight)
        ref_g = dict(mo_modifier=wf.mo.mo_modifier.grad.clone(), ci=wf.fc.weight.grad.clone(),
                     jastrow_weight=wf.jastrow.jastrow_kernel.weight.grad.clone())
        og, _ = orc.param_grads(P, pos, eloc=eloc, names=tuple(ref_g))
        for k, v in ref_g.items():
            errs["g_" + k] = float((og[k] - v).abs().max() / max(float(v.abs().max()), 1e-6))
        # Metropolis decisions, teacher forced
        with torch.no_grad():
            gen = torch.Generator().manual_seed(99)
            sig = orc.proposal_sigma(0.3)
            cur = pos.clone()
            fx = wf.pdf(cur)
            tr_disp, tr_tau, tr_acc, tr_pos = [], [], [], [cur.numpy().copy()]
            for it in range(3):
                disp = torch.randn(cur.shape, generator=gen, dtype=torch.float64) * np.sqrt(sig)
                xn = cur + disp
                fxn = wf.pdf(xn)
                df = fxn / fx
                torch.manual_seed(1000 + it)
                idx = sampler._accept(df.clone())
                torch.manual_seed(1000 + it)
                tau = torch.rand_like(df)
                npos, nfx, oacc, ofxn = orc.metropolis_step(P, cur, fx.clone(), disp, tau)
                assert bool((oacc == idx).all()), "oracle accept mismatch"
                cur[idx, :] = xn[idx, :]
                fx[idx] = fxn[idx]
                assert torch.equal(npos, cur)
                tr_disp.append(disp.numpy()); tr_tau.append(tau.numpy()); tr_acc.append(idx.numpy())
                tr_pos.append(cur.numpy().copy())
        print("%-12s W=%3d " % (key, nw) + " ".join("%s=%.1e" % kv for kv in errs.items()))
        bad = {k: v for k, v in errs.items() if not v < (2e-9 if k == "eloc" or k.startswith("g_") else 5e-11)}
        assert not bad, "oracle does not reproduce the reference: %r" % bad
        out.update({key + "_pos": pos.numpy(), key + "_psi": psi.numpy(), key + "_eloc": eloc.numpy(),
                    key + "_ao": ao[:8].numpy(), key + "_mo_modifier": P.mo_modifier.numpy(), key + "_ci": P.ci.numpy(),
                    key + "_cfg_up": wf.configs[0].numpy(), key + "_cfg_down": wf.configs[1].numpy(),
                    key + "_mh_disp": np.stack(tr_disp), key + "_mh_tau": np.stack(tr_tau),
                    key + "_mh_acc": np.stack(tr_acc), key + "_mh_pos": np.stack(tr_pos)})
        for k, v in ref_g.items():
            out[key + "_grad_" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "sph.npz"), **out)


def walker_init_fixture():
    """Initial ensembles of the reference's ``Walkers.initialize`` (sampler/walkers.py:41-150) for
    the four ``Molecule.domain`` methods, seeds 5/5 -> tests/golden/walkers_init.npz."""
    from qmctorch.sampler.walkers import Walkers
    out = {}
    for key in ("lih", "h2o"):
        mol = fixture_molecule(key)
        for method in ("center", "uniform", "normal", "atomic"):
            torch.manual_seed(5)
            np.random.seed(5)
            w = Walkers(nwalkers=9, nelec=mol.nelec, ndim=3, init=mol.domain(method))
            w.initialize()
            out["%s_%s" % (key, method)] = w.pos.double().numpy()
    np.savez_compressed(os.path.join(OUT, "walkers_init.npz"), **out)
    print("walkers_init  %d ensembles" % len(out))


def gradient_sampler_fixture():
    """Chains of the reference's GeneralizedMetropolis and Hamiltonian samplers (autograd drift,
    sampler/generalized_metropolis.py, sampler/hamiltonian.py) on LiH 6-31G, and the check that the
    oracle restatements with the ANALYTIC density gradient reproduce them under the same seed
    -> tests/golden/gradient_samplers.npz."""
    from qmctorch.sampler import GeneralizedMetropolis, Hamiltonian
    case = [c for c in CASES if c[0] == "lih_ground"][0]
    torch.manual_seed(1234)
    mol, wf = build(case)
    with torch.no_grad():
        wf.jastrow.jastrow_kernel.weight.fill_(0.8)
    P = orc.make_params(mol, (wf.configs[0], wf.configs[1]), jastrow_weight=0.8)
    torch.manual_seed(77)
    start = Metropolis(nwalkers=48, nstep=60, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                       move={"type": "all-elec", "proba": "normal"})(wf.pdf, with_tqdm=False).detach().clone()
    out = dict(start=start.numpy(), jw=np.array([0.8]))
    # generalized Metropolis: 12 steps, keep the last 4
    gm = GeneralizedMetropolis(nwalkers=48, nstep=12, step_size=0.2, ntherm=8, ndecor=1, nelec=wf.nelec, ndim=3,
                               init=mol.domain("normal"))
    torch.manual_seed(5)
    ref = gm(wf.pdf, pos=start.clone(), with_tqdm=False).detach()
    torch.manual_seed(5)
    mine = orc.generalized_metropolis(P, start.clone(), 12, 0.2, ntherm=8)
    err = relmax(mine, ref)
    print("generalized_metropolis  chain err %.1e  moved %.2f" % (err, float((ref[-48:] != start).any(1).float().mean())))
    assert err < 1e-9
    out.update(gm_pos=ref.numpy(), gm_seed=np.array([5]), gm_cfg=np.array([12, 8, 1]), gm_step=np.array([0.2]))
    # Hamiltonian: 4 trajectories of 6 leapfrog steps, keep the last 2
    hm = Hamiltonian(nwalkers=48, nstep=4, step_size=0.05, L=6, ntherm=2, ndecor=1, nelec=wf.nelec, ndim=3,
                     init=mol.domain("normal"))
    torch.manual_seed(6)
    ref = hm(wf.pdf, pos=start.clone(), with_tqdm=False).detach()
    torch.manual_seed(6)
    mine = orc.hamiltonian(P, start.clone(), 4, 0.05, 6, ntherm=2)
    err = relmax(mine, ref)
    print("hamiltonian             chain err %.1e  moved %.2f" % (err, float((ref[-48:] != start).any(1).float().mean())))
    assert err < 1e-9
    out.update(hm_pos=ref.numpy(), hm_seed=np.array([6]), hm_cfg=np.array([4, 2, 1, 6]), hm_step=np.array([0.05]))
    np.savez_compressed(os.path.join(OUT, "gradient_samplers.npz"), **out)


def gto2sto_fixture():
    """`SlaterJastrow.gto2sto()` of the reference (slater_jastrow.py:649-733) on LiH 6-31G and H2 STO-3G:
    the fitted single-zeta Slater basis, and psi / E_L / grad psi of the returned wave function on
    walkers sampled from it -> tests/golden/gto2sto.npz.  The oracle is asserted on the new basis."""
    out = {}
    for key in ("lih", "h2"):
        torch.manual_seed(4321)
        mol = fixture_molecule(key)
        wf = SlaterJastrow(mol, configs="ground_state", include_all_mo=True).gto2sto()
        with torch.no_grad():
            wf.jastrow.jastrow_kernel.weight.fill_(0.8)
        b = wf.mol.basis
        assert b.radial_type == "sto_pure"
        pos = Metropolis(nwalkers=64, nstep=100, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                         move={"type": "all-elec", "proba": "normal"})(wf.pdf, with_tqdm=False).detach().clone()
        psi = wf(pos).detach()
        eloc = wf.local_energy(pos).detach()
        gpsi = wf.gradients_jacobi(pos, sum_grad=False).detach().reshape(len(pos), -1)
        P = orc.make_params(wf.mol, wf.configs, jastrow_weight=0.8)
        e_psi, e_el = rel(orc.psi(P, pos), psi), rel(orc.local_energy(P, pos), eloc)
        print("gto2sto %-4s exps %s  oracle psi=%.1e eloc=%.1e" % (key, np.round(b.bas_exp, 4), e_psi, e_el))
        assert e_psi < 1e-12 and e_el < 1e-11
        out.update({key + "_bas_exp": np.asarray(b.bas_exp, dtype=np.float64),
                    key + "_bas_norm": np.asarray(b.bas_norm, dtype=np.float64),
                    key + "_bas_k": np.stack([b.bas_kx, b.bas_ky, b.bas_kz, b.bas_kr]).astype(np.int64),
                    key + "_nshells": np.asarray(b.nshells, dtype=np.int64),
                    key + "_norm_cst": wf.ao.norm_cst.detach().numpy(),
                    key + "_pos": pos.numpy(), key + "_psi": psi.numpy(), key + "_eloc": eloc.numpy(),
                    key + "_gpsi": gpsi.numpy()})
    np.savez_compressed(os.path.join(OUT, "gto2sto.npz"), **out)


if __name__ == "__main__":
    special = {"walkers_init": walker_init_fixture, "gradient_samplers": gradient_sampler_fixture,
               "gto2sto": gto2sto_fixture, "sph": spherical_fixture}
    args = sys.argv[1:]
    for name, fn in special.items():
        if not args or name in args:
            fn()
    rest = [a for a in args if a not in special]
    if not args or rest:
        main(rest)
