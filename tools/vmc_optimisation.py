"""BASELINE config 3: LiH wave-function optimisation (Jastrow + MO coefficients, manual energy
gradient) through the public Solver API; wall clock per epoch with a breakdown by CUDA events."""
import os
import sys
import time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200 import set_torch_double_precision
from qmctorch_b200.scf import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.solver import Solver
from qmctorch_b200.wavefunction import SlaterJastrow

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nepoch = int(sys.argv[2]) if len(sys.argv) > 2 else 10
set_torch_double_precision()
mol = fixture_molecule("lih")
wf = SlaterJastrow(mol, configs="ground_state", cuda=True)
sampler = Metropolis(nwalkers=nw, nstep=200, step_size=0.3, ntherm=-1, ndecor=1, nelec=wf.nelec, ndim=3,
                     init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=1)
opt = torch.optim.Adam(wf.parameters(), lr=5e-3)
solver = Solver(wf=wf, sampler=sampler, optimizer=opt)
solver.configure(track=["local_energy"], freeze=["ci", "ao"], loss="energy", grad="manual",
                 resampling={"mode": "update", "resample_every": 1, "nstep_update": 25, "ntherm_update": -1})
torch.cuda.synchronize()
t0 = time.time()
solver.prepare_optimization(None, None)
torch.cuda.synchronize()
print("prepare (initial sampling 200 steps + observables): %.3f s" % (time.time() - t0))
solver.run_epochs(2)                      # warm-up (NVRTC build, allocator)
torch.cuda.synchronize()
t0 = time.time()
solver.run_epochs(nepoch)
torch.cuda.synchronize()
dt = (time.time() - t0) / nepoch
e = solver.observable.local_energy
print("LiH optimisation, %d walkers: %.1f ms / epoch (E_L + psi fwd/bwd + Adam + 25 Metropolis steps); "
      "energy first %.5f -> last %.5f" % (nw, dt * 1e3, float(e[2].mean()), float(e[-1].mean())))
# breakdown
pos = solver.dataloader.dataset
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); t = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t) / n * 1e3
print("  evaluate_grad_manual %.2f ms | store_observable %.2f ms | resample(25 steps) %.2f ms | opt.step %.2f ms"
      % (timed(lambda: (opt.zero_grad(), solver.evaluate_grad_manual(pos))),
         timed(lambda: solver.store_observable(pos, local_energy=wf.local_energy(pos), ibatch=0)),
         timed(lambda: solver.resample(1, pos)), timed(lambda: opt.step())))
print("  sampler.nstep=%d ntherm=%d ndecor=%d keep_on_device=%s rng=%s" % (sampler.nstep, sampler.ntherm, sampler.ndecor,
      sampler.keep_on_device, sampler.rng))
for k in range(3):
    print("  resample again: %.2f ms" % timed(lambda: solver.resample(1, pos), n=3))
print("  25 raw qmcb_metropolis_step-equivalent sampler call: %.2f ms" % timed(lambda: sampler(wf.pdf, pos=pos, with_tqdm=False), n=3))
def touch_and_plan():
    with torch.no_grad():
        wf.mo.mo_modifier.mul_(1.0)
    wf._handle.plan()
print("  parameter sync (D2H of the parameters + qmcb_plan_update): %.2f ms" % timed(touch_and_plan))
