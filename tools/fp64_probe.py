"""FP64 pipe probes: DFMA (vector) and DMMA (mma.sync m8n8k4) throughput on the whole GPU."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qmctorch_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda", 0)
sink = torch.zeros(1, dtype=torch.float64, device=dev)
fl = ctypes.c_double(0.0)
sp = _lib.stream_ptr(dev)
for kind, name in ((0, "DFMA"), (1, "DMMA m8n8k4"), (2, "mixed 4 DMMA + 8 DFMA")):
    best = 0.0
    for rep in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.qmcb_fp64_probe(kind, 40000, _lib.ptr(sink), ctypes.byref(fl), sp)
        b.record()
        torch.cuda.synchronize()
        best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    print("%-12s %.2f TFLOP/s" % (name, best))
