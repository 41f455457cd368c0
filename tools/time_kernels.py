"""Kernel-only timings (CUDA events) of the fused entry points for one workload."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200 import _lib
from qmctorch_b200.molecules import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.wavefunction import SlaterJastrow

key = sys.argv[1] if len(sys.argv) > 1 else "lih"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cfg = {"lih": "ground_state", "h2": "single(2,2)", "h2o": "cas(4,4)", "c4h6": "ground_state"}.get(key, "ground_state")
step = {"lih": 0.3, "h2": 0.5, "h2o": 0.15, "c4h6": 0.05}.get(key, 0.3)
mol = fixture_molecule(key)
wf = SlaterJastrow(mol, configs=cfg, cuda=True)
s = Metropolis(nwalkers=nw, nstep=10, step_size=step, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
               move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=0, keep_on_device=True)
pos = s(wf.pdf, with_tqdm=False).detach()
L = _lib.lib()
plan = wf._handle.plan()
dev = pos.device
sp = _lib.stream_ptr(dev)
W = pos.shape[0]
e = torch.empty(W, dtype=torch.float64, device=dev)
g = torch.empty(W, 3 * wf.nelec, dtype=torch.float64, device=dev)
fx = (wf(pos).reshape(-1) ** 2).detach().contiguous()
x = pos.clone()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t_e = timeit(lambda: L.qmcb_local_energy(plan, _lib.ptr(pos), W, _lib.ptr(e), None, None, sp))
t_p = timeit(lambda: L.qmcb_psi(plan, _lib.ptr(pos), W, _lib.ptr(e), sp))
t_g = timeit(lambda: L.qmcb_grad_psi(plan, _lib.ptr(pos), W, 0, _lib.ptr(g), sp))
cnt = [0]


def mh():
    L.qmcb_metropolis_step(plan, _lib.ptr(x), _lib.ptr(fx), W, None, None, None, -1, 1, 0.3, 1e-16, 7, cnt[0],
                           None, None, sp)
    cnt[0] += 1


t_m = timeit(mh)
info = [wf._handle.info(i) for i in range(10)]
print("specialised kernels: %d (%s)" % (wf._handle.info(13), L.qmcb_last_error().decode()[:200] if not wf._handle.info(13) else "NVRTC"))
print("%s lib=%s W=%d tile=%d thr=%d smem=%d | eloc %.3f ms (%.3e/s) psi %.3f ms (%.3e/s) grad %.3f ms mh %.3f ms (%.3e/s)"
      % (key, os.path.basename(_lib.LIB_PATH), W, info[6], info[7], info[8], t_e, W / t_e * 1e3, t_p, W / t_p * 1e3,
         t_g, t_m, W / t_m * 1e3), flush=True)
