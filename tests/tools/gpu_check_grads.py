"""Diagnostic: parameter gradients of every golden case, CUDA path vs reference golden."""
import os, sys, traceback
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import _cases as C  # noqa
from test_gpu_parity import _manual_grads  # noqa

for name in C.CASES:
    g = C.load(name)
    try:
        mol, wf = C.build_wf(g)
        grads = _manual_grads(wf, torch.tensor(g["pos"]).cuda())
        res = {}
        for k in [k[5:] for k in g if k.startswith("grad_")]:
            ref = torch.tensor(g["grad_" + k])
            got = grads[k].detach().cpu().reshape(ref.shape)
            res[k] = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-6))
        print("%-14s " % name + " ".join("%s=%.1e" % kv for kv in res.items()), flush=True)
    except Exception:
        print(name, "FAILED"); traceback.print_exc()
