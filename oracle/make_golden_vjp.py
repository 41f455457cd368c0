"""Golden vectors for the adjoint of the local energy (qmcb_local_energy_backward): the UNMODIFIED reference
back-propagates through WaveFunction.local_energy and through psi.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_vjp.py

For the golden cases below (walkers and parameters of tests/golden/<case>.npz) it stores, in
tests/golden/vjp.npz,

    gE_<leaf> = d/d leaf  sum_w wE_w E_L(R_w)         (torch.autograd.grad through wf.local_energy)
    gP_<leaf> = d/d leaf  sum_w wP_w psi(R_w)         (through wf.forward)

for the leaves ao.atom_coords, ao.bas_exp, ao.bas_coeffs, mo.mo_modifier, fc.weight and the Pade Jastrow
weights, with seeded random weights wE, wP; the result of the reference's Solver.compute_forces; and the
parameter gradients its Solver leaves in .grad after evaluate_grad_auto for the energy and the variance loss.
It also checks that autograd through oracle/sj_oracle.py (the CPU restatement) reproduces the E_L adjoint, which
pins the oracle for this row.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))

import make_golden as mg  # noqa: E402  (loads the reference behind the stubs)
import sj_oracle as orc  # noqa: E402
import _cases as C  # noqa: E402

from qmctorch.solver import Solver  # noqa: E402
from qmctorch.sampler import Metropolis  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "vjp.npz")

# case -> walkers used
CASES = {"h2_single22": 24, "lih_ground": 24, "lih_cas24": 16, "lih_een": 16, "h2o_cas44": 8, "lih_sto": 16,
         "lih_sto_pure": 16, "lih_gto_kr": 16, "lih_adf_sd22": 16, "lih_sd22_een3": 8, "c4h6_ground": 4}
SOLVER_CASES = ["lih_een", "lih_cas24"]


def reference_wf(name):
    case = [c for c in mg.CASES if c[0] == name][0]
    g = C.load(name)
    mol, wf = mg.build(case)
    jt = g["jastrow"]
    with torch.no_grad():
        wf.mo.mo_modifier.copy_(torch.tensor(g["mo_modifier"]))
        wf.fc.weight.copy_(torch.tensor(g["ci"]))
        if jt == "ee":
            wf.jastrow.jastrow_kernel.weight.fill_(float(g["jw"][0]))
        elif jt != "None":
            wf.jastrow.jastrow_terms[0].jastrow_kernel.weight.fill_(float(g["jw"][0]))
            if "+en" in jt.replace("+een", ""):
                wf.jastrow.jastrow_terms[1].jastrow_kernel.weight.fill_(float(g["enw"][0]))
            if jt.endswith("een"):
                bh = wf.jastrow.jastrow_terms[-1].jastrow_kernel
                bh.weight_num.copy_(torch.tensor(g["bh_num"]))
                bh.weight_denom.copy_(torch.tensor(g["bh_denom"]))
                bh.fc.weight.copy_(torch.tensor(g["bh_fc"]))
    return g, mol, wf


def leaves_of(wf, jt):
    wf.ao.atom_coords.requires_grad = True
    wf.ao.bas_coeffs.requires_grad = True
    L = {"atom_coords": wf.ao.atom_coords, "bas_exp": wf.ao.bas_exp, "bas_coeffs": wf.ao.bas_coeffs,
         "mo_modifier": wf.mo.mo_modifier, "ci": wf.fc.weight}
    if jt == "ee":
        L["jee_w"] = wf.jastrow.jastrow_kernel.weight
    elif jt != "None":
        L["jee_w"] = wf.jastrow.jastrow_terms[0].jastrow_kernel.weight
        if "+en" in jt.replace("+een", ""):
            L["jen_w"] = wf.jastrow.jastrow_terms[1].jastrow_kernel.weight
    return L


def oracle_adjoint(g, pos, wE):
    """The same adjoint by autograd through the oracle (no three-body term: the oracle detaches it)."""
    mol, P = C.oracle_params(g)
    return orc.local_energy_adjoint(P, pos, w_eloc=wE)


def main():
    out = {}
    gen = torch.Generator().manual_seed(2024)
    for name, nw in CASES.items():
        g, mol, wf = reference_wf(name)
        jt = g["jastrow"]
        pos = torch.tensor(g["pos"][:nw])
        wE = torch.rand(nw, generator=gen, dtype=torch.float64) - 0.3
        wP = torch.rand(nw, generator=gen, dtype=torch.float64) - 0.3
        L = leaves_of(wf, jt)
        names = list(L)
        x = pos.clone()
        if jt.endswith("een"):
            x.requires_grad_(True)          # the reference differentiates the three-body term by autograd
        e = wf.local_energy(x)
        gE = torch.autograd.grad((e.reshape(-1) * wE).sum(), [L[n] for n in names], allow_unused=True)
        psi = wf(pos)
        gP = torch.autograd.grad((psi.reshape(-1) * wP).sum(), [L[n] for n in names], allow_unused=True)
        out[name + "/n"] = np.array([nw])
        out[name + "/wE"] = wE.numpy()
        out[name + "/wP"] = wP.numpy()
        out[name + "/eloc"] = e.detach().reshape(-1).numpy()
        for n, a, b in zip(names, gE, gP):
            if a is not None:
                out[name + "/gE_" + n] = a.detach().numpy()
            if b is not None:
                out[name + "/gP_" + n] = b.detach().numpy()
        msg = ""
        if not jt.endswith("een") and name != "c4h6_ground":
            go = oracle_adjoint(g, pos, wE)
            worst = 0.0
            for n, a in zip(names, gE):
                if a is None or go.get(n) is None:
                    continue
                err = float((go[n] - a).abs().max() / a.abs().max().clamp(min=1e-6))   # (a one-determinant ci gradient is noise around 0)
                if err > 1e-10:
                    print("   ", name, n, "%.2e" % err, "max|g| %.2e" % float(a.abs().max()))
                worst = max(worst, err)
            assert worst < 1e-7, (name, worst)
            msg = "  oracle autograd vs reference %.1e" % worst
        print("%-16s W=%d leaves=%s%s" % (name, nw, ",".join(n for n, a in zip(names, gE) if a is not None), msg))

    # end to end through the reference's Solver: forces and grad="auto"
    for name in SOLVER_CASES:
        g, mol, wf = reference_wf(name)
        nw = CASES[name]
        pos = torch.tensor(g["pos"][:nw])
        sampler = Metropolis(nwalkers=nw, nstep=10, step_size=0.3, nelec=wf.nelec, ndim=3,
                             init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"})
        opt = torch.optim.SGD(wf.parameters(), lr=1e-3)
        solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
        out[name + "/forces"] = solver.compute_forces(pos.clone()).detach().numpy()
        out[name + "/forces_clip"] = solver.compute_forces(pos.clone(), batch_size=nw // 2, clip=2).detach().numpy()
        for loss in ("energy", "variance"):
            solver.configure(track=["local_energy"], loss=loss, grad="auto",
                             resampling={"mode": "update", "resample_every": 1, "nstep_update": 5})
            wf.zero_grad()
            val, _ = solver.evaluate_gradient(pos.clone())
            out[name + "/auto_%s_loss" % loss] = np.array([float(val)])
            for pn, p in wf.named_parameters():
                if p.grad is not None:
                    out[name + "/auto_%s/%s" % (loss, pn)] = p.grad.detach().numpy().copy()
        # sampling weights (loss.py:112-153): walkers kept for two epochs -> the second loss is weighted by
        # (psi / psi0)^2 / sum, and its gradient passes through psi as well as through E_L
        solver.configure(track=["local_energy"], loss="energy", grad="auto",
                         resampling={"mode": "update", "resample_every": 2, "nstep_update": 5})
        assert solver.loss.use_weight
        wf.zero_grad()
        solver.evaluate_gradient(pos.clone())          # records psi0, weights = 1
        opt.step()                                     # SGD, lr = 1e-3: psi now differs from psi0
        wf.zero_grad()
        val, _ = solver.evaluate_gradient(pos.clone())
        out[name + "/auto_weighted_loss"] = np.array([float(val)])
        for pn, p in wf.named_parameters():
            if p.grad is not None:
                out[name + "/auto_weighted/%s" % pn] = p.grad.detach().numpy().copy()
        print("%-16s solver: forces, grad=auto (energy, variance, weighted energy)" % name)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "%.0f KB" % (os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
