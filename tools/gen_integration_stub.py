"""Regenerates the `_System` ctypes stub of INTEGRATION.md from the binding the package itself uses
(qmctorch_b200/_lib.py: QmcbSystem), so the documented struct cannot drift from include/qmcb.h
(tests/test_host.py checks header, binding and document against each other).

    python tools/gen_integration_stub.py          # rewrites the block between the two markers
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200._lib import QmcbSystem  # noqa: E402

BEGIN, END = "# >>> qmcb_system fields (generated: tools/gen_integration_stub.py)", "# <<< qmcb_system fields"
NAMES = {C.c_int32: "C.c_int32", C.c_double: "C.c_double", C.c_void_p: "C.c_void_p"}


def block():
    rows, line = [], "    _fields_ = ["
    for name, typ in QmcbSystem._fields_:
        item = '("%s", %s), ' % (name, NAMES[typ])
        if len(line) + len(item) > 108:
            rows.append(line.rstrip())
            line = "                "
        line += item
    rows.append(line.rstrip().rstrip(",") + "]")
    return "\n".join([BEGIN] + rows + [END])


def main():
    path = os.path.join(ROOT, "INTEGRATION.md")
    text = open(path).read()
    a, b = text.index(BEGIN), text.index(END) + len(END)
    new = text[:a] + block() + text[b:]
    if new != text:
        open(path, "w").write(new)
        print("INTEGRATION.md: stub regenerated")


if __name__ == "__main__":
    main()
