"""Diagnostic dump (not a test): every golden case, every quantity, CUDA path vs golden."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import _cases as C  # noqa


def main():
    torch.set_default_dtype(torch.float64)
    for name in C.CASES:
        g = C.load(name)
        try:
            mol, wf = C.build_wf(g)
            pos = torch.tensor(g["pos"]).cuda()
            res = {}
            res["psi"] = C.rel_err(wf(pos), g["psi"])
            res["eloc"] = C.rel_err(wf.local_energy(pos), g["eloc"])
            res["ekin"] = C.rel_err(wf.kinetic_energy(pos), g["ekin"])
            ns = g["ao"].shape[0]
            ao, dao, d2ao = wf.ao(pos[:ns], derivative=[0, 1, 2])
            res["ao"] = C.scaled_err(ao, g["ao"]); res["dao"] = C.scaled_err(dao, g["dao"])
            res["d2ao"] = C.scaled_err(d2ao, g["d2ao"])
            if wf.use_jastrow:
                J, dJ, d2J = wf.jastrow(pos, derivative=[0, 1, 2], sum_grad=False)
                res["J"] = C.rel_err(J, g["J"]); res["dJ"] = C.scaled_err(dJ, g["dJ"])
                res["d2J"] = C.scaled_err(d2J, g["d2J"])
            res["gpsi"] = C.scaled_err(wf.gradients_jacobi(pos), g["gpsi"])
            res["gpdf"] = C.scaled_err(wf.gradients_jacobi(pos, pdf=True), g["gpdf"])
            print("%-14s " % name + " ".join("%s=%.1e" % kv for kv in res.items()), flush=True)
        except Exception:
            print(name, "FAILED")
            traceback.print_exc()


if __name__ == "__main__":
    main()
