import os, sys, torch
sys.path.insert(0, os.getcwd())
from qmctorch_b200.molecules import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.wavefunction import SlaterJastrow
key, nw = sys.argv[1], int(sys.argv[2])
cfg = {"h2o": "cas(4,4)", "c4h6": "ground_state"}[key]
mol = fixture_molecule(key)
wf = SlaterJastrow(mol, configs=cfg, cuda=True)
s = Metropolis(nwalkers=nw, nstep=5, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
               move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=0, keep_on_device=True)
pos = s(wf.pdf, with_tqdm=False).detach()
for _ in range(3):
    g = wf.gradients_jacobi(pos)
torch.cuda.synchronize()
print(float(g.abs().mean()), [wf._handle.info(i) for i in (6, 7, 8, 9)])
