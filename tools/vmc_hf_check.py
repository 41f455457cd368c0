"""End-to-end physics check on real SCF orbitals: the VMC energy of the bare Hartree-Fock determinant
(no Jastrow factor) must reproduce the SCF total energy stored in the molecule file within the
statistical error.  Molecules: the ADF results read from the reference's HDF5 files (lih_adf, h2_adf,
co2_adf; `Molecule(load=...)`).

    python tools/vmc_hf_check.py lih_adf 1000000 1000 [step_size] [all-elec|one-elec]
"""
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

key = sys.argv[1] if len(sys.argv) > 1 else "lih_adf"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nstep = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
step = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
move = sys.argv[5] if len(sys.argv) > 5 else "all-elec"
set_torch_double_precision()
mol = fixture_molecule(key)
wf = SlaterJastrow(mol, configs="ground_state", jastrow=None, cuda=True)
sampler = Metropolis(nwalkers=nw, nstep=nstep, step_size=step, ntherm=-1, ndecor=1, nelec=wf.nelec, ndim=3,
                     init=mol.domain("atomic"), move={"type": move, "proba": "normal"}, cuda=True, seed=3)
solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.Adam(wf.parameters(), lr=0.01))
t0 = time.time()
obs = solver.single_point(with_tqdm=False)
torch.cuda.synchronize()
e, err, scf = float(obs.energy), float(obs.error), float(mol.get_total_energy())
print("%s HF determinant: E_VMC = %.5f +- %.5f   E_SCF(file) = %.5f   diff = %+.5f (%.1f sigma)  var = %.4f  "
      "acceptance %.2f  %d walkers x %d steps in %.2f s"
      % (key, e, err, scf, e - scf, abs(e - scf) / err, float(obs.variance), sampler.acceptance_rate, nw, nstep,
         time.time() - t0), flush=True)
