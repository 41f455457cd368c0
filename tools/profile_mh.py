"""Small driver for ncu: a few fused Metropolis moves (spec_mh / spect_mh) on a fixture."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200.molecules import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.wavefunction import SlaterJastrow
key = sys.argv[1] if len(sys.argv) > 1 else "lih"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cfg = {"lih": "ground_state", "h2": "single(2,2)", "h2o": "cas(4,4)", "c4h6": "ground_state"}[key]
mol = fixture_molecule(key)
wf = SlaterJastrow(mol, configs=cfg, cuda=True)
s = Metropolis(nwalkers=nw, nstep=8, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
               move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=0, keep_on_device=True, init_rng="philox")
pos = s(wf.pdf, with_tqdm=False).detach()
torch.cuda.synchronize()
print(s.acceptance_rate)
