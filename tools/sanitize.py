"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): a few hundred walkers
through every fused entry point of the kernel kinds in use - one walker per thread (LiH: spec_* and
spec_backward), warp tiles (spect_*: H2O cas(4,4) with thread-per-block determinants, H2O + three-body
factor tables, C4H6 / triplet C2H4 with the half-warp Gauss-Jordan and the pair-once Jastrow, a
spherical-harmonics basis), the adjoint kernel of the local energy on each of them - and, with QMCB_JIT=0, the generic CTA-tile kernels of the same systems.

    compute-sanitizer --tool racecheck python tools/sanitize.py
"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200 import _lib
from qmctorch_b200.molecules import Molecule, _parse_atoms, _seeded_mos, build_basis, fixture_molecule
from qmctorch_b200.wavefunction import SlaterJastrow

C2H4 = "C 0 0 0.667; C 0 0 -0.667; H 0 0.923 1.238; H 0 -0.923 1.238; H 0 0.923 -1.238; H 0 -0.923 -1.238"


def systems():
    yield "lih", fixture_molecule("lih"), "ground_state", 700
    yield "h2o", fixture_molecule("h2o"), "cas(4,4)", 130
    yield "c4h6", fixture_molecule("c4h6"), "ground_state", 50
    yield "lih_sph", fixture_molecule("lih_sph"), "single_double(2,2)", 300
    yield "h2o_een", fixture_molecule("h2o"), "cas(2,2)", 100
    names, coords = _parse_atoms(C2H4, "angs")
    nao = build_basis(names, coords, "dzp").nao
    yield "c2h4_triplet", Molecule(C2H4, basis="dzp", unit="angs", spin=2, name="c2h4t", mos=_seeded_mos(nao, 5)), \
        "ground_state", 70


for name, mol, cfg, W in systems():
    jast = "default"
    if name == "h2o_een":
        from qmctorch_b200.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
        from qmctorch_b200.wavefunction.jastrows.elec_elec_nuclei import JastrowFactor as JEEN, BoysHandyJastrowKernel as BH
        jast = [JEE(mol, PEE, cuda=True), JEEN(mol, BH, cuda=True)]
    wf = SlaterJastrow(mol, configs=cfg, jastrow=jast, cuda=True)
    g = torch.Generator().manual_seed(1)
    mean = torch.as_tensor(mol.domain("normal")["mean"])
    sig = torch.as_tensor(mol.domain("normal")["sigma"]).diagonal()
    pos = (mean + sig * torch.randn(W, mol.nelec, 3, generator=g, dtype=torch.float64)).view(W, -1).cuda()
    L = _lib.lib()
    plan = wf._handle.plan()
    sp = _lib.stream_ptr(pos.device)
    psi = wf(pos)
    e, s4 = wf.local_energy_stats(pos)
    gr = wf.gradients_jacobi(pos)
    fx = (psi.reshape(-1) ** 2).contiguous()
    acc = torch.zeros(W, dtype=torch.uint8, device="cuda")
    x = pos.clone()
    _lib.check(L.qmcb_metropolis_step(plan, _lib.ptr(x), _lib.ptr(fx), W, None, None, None, -1, 1, 0.2, 1e-16, 3, 0,
                                      _lib.ptr(acc), None, sp), "mh")
    wgt = torch.randn(W, dtype=torch.float64, device="cuda")
    gb = wf._psi_backward(pos, wgt, {"mo_modifier", "ci", "jee_w"})
    # adjoint of the local energy (eloc_vjp.cu): every leaf, E_L and psi weights, lane groups / shared or global
    # work areas as the launch heuristic picks them for this structure
    nv = min(W, 96)
    want = {"mo_modifier", "ci", "bas_exp", "bas_coeffs", "jee_w", "atom_coords"}
    if not wf.ao.contract:
        want.discard("bas_coeffs")
    ga = wf._eloc_backward(pos[:nv].contiguous(), wgt[:nv].contiguous(), wgt[:nv].contiguous(), want)
    torch.cuda.synchronize()
    print(name, "adjoint dR", float(ga["atom_coords"].abs().sum()), "dexp", float(ga["bas_exp"].abs().sum()), flush=True)
    print(name, "psi", float(psi.abs().mean()), "E", float(s4[0] / s4[2]), "grad", float(gr.abs().mean()),
          "acc", float(acc.float().mean()), "dW", float(gb["mo_modifier"].abs().sum()), "kind", wf._handle.info(15), flush=True)
