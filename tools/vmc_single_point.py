"""End-to-end VMC single point through the public API (Solver.single_point): sampling with the fused
Metropolis kernel, local energy, statistics.  Wall clock, everything included."""
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

key = sys.argv[1] if len(sys.argv) > 1 else "lih"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nstep = int(sys.argv[3]) if len(sys.argv) > 3 else 500
set_torch_double_precision()
mol = fixture_molecule(key)
cfg = {"lih": "ground_state", "h2": "single(2,2)", "h2o": "ground_state", "c4h6": "ground_state"}[key]
step = {"lih": 0.3, "h2": 0.5, "h2o": 0.15, "c4h6": 0.05}[key]
wf = SlaterJastrow(mol, configs=cfg, cuda=True)
for trial in range(2):       # first pass includes the NVRTC build and CUDA context creation
    sampler = Metropolis(nwalkers=nw, nstep=nstep, step_size=step, ntherm=-1, ndecor=1, nelec=wf.nelec, ndim=3,
                         init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=trial)
    solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.Adam(wf.parameters(), lr=0.01))
    torch.cuda.synchronize()
    t0 = time.time()
    obs = solver.single_point(with_tqdm=False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("%s single point: %d walkers x %d Metropolis steps + E_L: %.3f s wall  (%.3e walker-steps/s)  "
          "E = %.5f +- %.5f  acceptance %.2f  specialised=%d"
          % (key, nw, nstep, dt, nw * nstep / dt, float(obs.energy), float(obs.error), sampler.acceptance_rate,
             wf._handle.info(13)), flush=True)
