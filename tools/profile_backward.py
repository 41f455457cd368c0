"""Driver for ncu: a few launches of the backward (parameter-gradient) kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200.molecules import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.wavefunction import SlaterJastrow
key = sys.argv[1] if len(sys.argv) > 1 else "lih"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
mol = fixture_molecule(key)
wf = SlaterJastrow(mol, configs={"lih": "ground_state", "h2o": "cas(4,4)", "c4h6": "ground_state"}[key], cuda=True)
s = Metropolis(nwalkers=nw, nstep=5, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"), cuda=True,
               seed=0, keep_on_device=True)
pos = s(wf.pdf, with_tqdm=False).detach()
w = torch.randn(nw, device=pos.device, dtype=torch.float64) / nw
for _ in range(3):
    g = wf._psi_backward(pos, w)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g = wf._psi_backward(pos, w)
b.record()
torch.cuda.synchronize()
print("backward %s W=%d (all gradients): %.3f ms/call (%.3e walkers/s)" % (key, nw, a.elapsed_time(b) / 5, nw / (a.elapsed_time(b) / 5) * 1e3))
# Jastrow + MO only (BASELINE config 3: freeze ci, ao)
from qmctorch_b200 import _lib
import ctypes
L = _lib.lib(); plan = wf._handle.plan(); dev = pos.device
nao, nmo = wf.mo.mo_scf.shape
g_mo = torch.empty(nao, nmo, dtype=torch.float64, device=dev); g_j = torch.empty(1, dtype=torch.float64, device=dev)
ws = torch.empty(int(L.qmcb_backward_workspace_bytes(plan, nw)), dtype=torch.uint8, device=dev)
def call():
    _lib.check(L.qmcb_psi_backward(plan, _lib.ptr(pos), _lib.ptr(w), nw, _lib.ptr(g_mo), None, None, None, _lib.ptr(g_j),
                                   None, None, _lib.ptr(ws), _lib.stream_ptr(dev)), "bwd")
for _ in range(3): call()
torch.cuda.synchronize(); a.record()
for _ in range(5): call()
b.record(); torch.cuda.synchronize()
print("backward %s W=%d (Jastrow + MO only): %.3f ms/call (%.3e walkers/s)" % (key, nw, a.elapsed_time(b) / 5, nw / (a.elapsed_time(b) / 5) * 1e3))
assert torch.allclose(g_mo * wf.mo.mo_scf, g["mo_modifier"], rtol=1e-9, atol=1e-14)
