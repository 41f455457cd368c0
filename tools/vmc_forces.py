"""End-to-end use of Solver.compute_forces on the CUDA path: H2 (STO-3G, core-Hamiltonian MOs, Pade Jastrow) at
three bond lengths, 10^6 walkers.  Prints d<E>/dR_A (what compute_forces returns; the force is its negative) and
the VMC energy.  Symmetry requires equal and opposite z components on the two nuclei and vanishing x, y components;
the sign must follow the energy curve (stretching a compressed bond lowers the energy, and vice versa).

    python tools/vmc_forces.py [walkers]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200 import set_torch_double_precision  # noqa: E402
from qmctorch_b200.molecules import Molecule  # noqa: E402
from qmctorch_b200.sampler import Metropolis  # noqa: E402
from qmctorch_b200.solver import Solver  # noqa: E402
from qmctorch_b200.wavefunction import SlaterJastrow  # noqa: E402

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
set_torch_double_precision()
for bond in (1.0, 1.4, 2.2):
    mol = Molecule(atom="H 0 0 %.6f; H 0 0 %.6f" % (-bond / 2, bond / 2), basis="sto-3g", unit="bohr", name="H2")
    wf = SlaterJastrow(mol, configs="ground_state", cuda=True)
    sampler = Metropolis(nwalkers=nw, nstep=400, step_size=0.4, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                         move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=1, init_rng="philox")
    solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.SGD(wf.parameters(), lr=1e-3))
    pos = sampler(wf.pdf, with_tqdm=False).detach()
    torch.cuda.synchronize()
    t0 = time.time()
    f = solver.compute_forces(pos)
    torch.cuda.synchronize()
    dt = time.time() - t0
    with torch.no_grad():
        e = wf.local_energy(pos)
    print("R = %.2f bohr  E = %.5f +- %.5f  dE/dR_A = [%+.4f %+.4f %+.4f]  dE/dR_B = [%+.4f %+.4f %+.4f]  "
          "(compute_forces: %.1f ms, acceptance %.2f)"
          % (bond, float(e.mean()), float(e.std() / nw ** 0.5), *f[0].tolist(), *f[1].tolist(), dt * 1e3,
             sampler.acceptance_rate), flush=True)
