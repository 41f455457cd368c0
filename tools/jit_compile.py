"""Compile the structure-specialised (NVRTC) kernels of a fixture OFFLINE (host-only plan, no GPU) and
print ptxas' resource usage; with QMCB_JIT_DUMP the source / PTX / cubin are kept for cuobjdump.

    python tools/jit_compile.py h2o "cas(4,4)" [ee|ee+een] [dump_prefix]
"""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
key = sys.argv[1] if len(sys.argv) > 1 else "h2o"
cfg = sys.argv[2] if len(sys.argv) > 2 else "ground_state"
jast = sys.argv[3] if len(sys.argv) > 3 else "ee"
dump = sys.argv[4] if len(sys.argv) > 4 else None
if dump:
    os.environ["QMCB_JIT_DUMP"] = dump
from qmctorch_b200 import _lib  # noqa: E402
from qmctorch_b200.molecules import fixture_molecule  # noqa: E402
from qmctorch_b200.wavefunction import SlaterJastrow  # noqa: E402

mol = fixture_molecule(key)
j = "default"
if jast == "ee+een":
    from qmctorch_b200.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
    from qmctorch_b200.wavefunction.jastrows.elec_elec_nuclei import JastrowFactor as JEEN, BoysHandyJastrowKernel as BH
    j = [JEE(mol, PEE), JEEN(mol, BH)]
wf = SlaterJastrow(mol, configs=cfg, jastrow=j, cuda=False)
L = _lib.lib()
arrays = wf._handle._system()
p = ctypes.c_void_p()
_lib.check(L.qmcb_plan_create(ctypes.byref(arrays.struct), -1, ctypes.byref(p)), "qmcb_plan_create")
t0 = time.time()
kind = L.qmcb_plan_info(p, 15)
print("%s %s %s: specialised kind = %d (1 thread, 2 tile) in %.1f s; %s" % (
    key, cfg, jast, kind, time.time() - t0, L.qmcb_last_error().decode()[:600] if not kind else "ok"))
L.qmcb_plan_destroy(p)
if dump and os.path.isfile(dump + ".cubin"):
    out = subprocess.run(["cuobjdump", "-res-usage", dump + ".cubin"], capture_output=True, text=True).stdout
    print(out)
