"""BASELINE config 4 (H2O cc-pVDZ, CAS + e-e + e-e-n Jastrow): oracle parity on a sample of the
walkers and kernel-only timings (CUDA events) for cas(2,2), cas(4,4), cas(6,6)."""
import os
import sys
import time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qmctorch_b200 import _lib
from qmctorch_b200.molecules import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.wavefunction import SlaterJastrow
from qmctorch_b200.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
from qmctorch_b200.wavefunction.jastrows.elec_elec_nuclei import JastrowFactor as JEEN, BoysHandyJastrowKernel as BH
from oracle import sj_oracle as orc      # checker only (tools/ is not the product path)

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
mol = fixture_molecule("h2o")
L = _lib.lib()
for cfg in sys.argv[2:] or ["cas(2,2)", "cas(4,4)", "cas(6,6)"]:
    torch.manual_seed(3)
    wf = SlaterJastrow(mol, configs=cfg, jastrow=[JEE(mol, PEE), JEEN(mol, BH)], cuda=True)
    with torch.no_grad():
        wf.fc.weight.add_(0.05 * torch.rand_like(wf.fc.weight))
    s = Metropolis(nwalkers=nw, nstep=20, step_size=0.15, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                   move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=0, keep_on_device=True)
    pos = s(wf.pdf, with_tqdm=False).detach()
    plan = wf._handle.plan()
    sp = _lib.stream_ptr(pos.device)
    W = pos.shape[0]
    e = torch.empty(W, dtype=torch.float64, device=pos.device)
    p = torch.empty_like(e)

    def run():
        _lib.check(L.qmcb_local_energy(plan, _lib.ptr(pos), W, _lib.ptr(e), _lib.ptr(p), None, sp), "eloc")

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    ns = 64
    t0 = time.time()
    P = orc.make_params(mol, (wf.configs[0].cpu(), wf.configs[1].cpu()),
                        jastrow_weight=float(wf.jastrow.jastrow_terms[0].jastrow_kernel.weight))
    P.ci = wf.fc.weight.detach().cpu().clone()
    k = wf.jastrow.jastrow_terms[1].jastrow_kernel
    P.een = dict(num=k.weight_num.detach().cpu(), denom=k.weight_denom.detach().cpu(), fc=k.fc.weight.detach().cpu())
    eo = orc.local_energy(P, pos[:ns].cpu()).reshape(-1)
    po = orc.psi(P, pos[:ns].cpu()).reshape(-1)
    rel_e = float(((e[:ns].cpu() - eo).abs() / eo.abs()).max())
    rel_p = float(((p[:ns].cpu() - po).abs() / po.abs()).max())
    info = [wf._handle.info(i) for i in range(10)]
    print("h2o %s nci=%d nuniq=%d/%d tile=%d thr=%d smem=%d | eloc %.3f ms  %.3e evals/s | rel err E_L %.1e psi %.1e (oracle %.1fs)"
          % (cfg, wf.nci, info[4], info[5], info[6], info[7], info[8], ms, W / ms * 1e3, rel_e, rel_p,
             time.time() - t0), flush=True)
