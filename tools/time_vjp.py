"""Times qmcb_local_energy_backward (the adjoint of E_L: grad="auto", forces) on the fixture systems.

    python tools/time_vjp.py [workload walkers] ...      e.g.  lih 1000000 h2o 100000 c4h6 20000

Prints ms per call for (a) every wave-function parameter, (b) Jastrow + MO + CI only (BASELINE config 3's set),
(c) the atom coordinates only (forces), next to one E_L evaluation of the same walkers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qmctorch_b200.molecules import fixture_molecule  # noqa: E402
from qmctorch_b200.wavefunction import SlaterJastrow  # noqa: E402

CONFIGS = {"lih": "ground_state", "h2": "single(2,2)", "h2o": "cas(4,4)", "c4h6": "ground_state", "lih_sto": "ground_state"}


def timeit(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main(argv):
    jobs = list(zip(argv[0::2], argv[1::2])) or [("lih", "1000000"), ("h2o", "100000"), ("c4h6", "20000")]
    for key, w in jobs:
        W = int(w)
        mol = fixture_molecule(key)
        wf = SlaterJastrow(mol, configs=CONFIGS.get(key, "ground_state"), cuda=True)
        torch.manual_seed(0)
        pos = torch.randn(W, wf.nelec * 3, dtype=torch.float64, device="cuda") * 0.8
        at = wf.ao.atom_coords.detach()
        pos = (pos.view(W, wf.nelec, 3) + at[torch.arange(wf.nelec) % at.shape[0]][None]).reshape(W, -1).contiguous()
        wE = torch.randn(W, dtype=torch.float64, device="cuda")
        sets = {"all parameters": {"bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w"},
                "jastrow+mo+ci": {"mo_modifier", "ci", "jee_w"}, "atom coordinates": {"atom_coords"},
                "parameters+atoms": {"bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "atom_coords"}}
        with torch.no_grad():
            t_e = timeit(lambda: wf.local_energy(pos))
        out = ["%s W=%d | E_L %.3f ms" % (key, W, t_e)]
        for name, want in sets.items():
            t = timeit(lambda: wf._eloc_backward(pos, wE, None, want), n=5, warm=1)
            out.append("%s %.3f ms (%.3g walkers/s)" % (name, t, W / t * 1e3))
        print(" | ".join(out), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
