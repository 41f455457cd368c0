"""Import the UNMODIFIED reference (QMCTorch v0.4.0) behind stub modules.

TEST INFRASTRUCTURE ONLY.  The reference's third-party imports (twiggy, h5py,
matplotlib, mendeleev, pints, pyscf, scm.plams, ase) are absent from this image,
and none of them carries hot-path arithmetic.  This shim
registers inert stubs for exactly those names, puts the reference root on
``sys.path`` and returns the ``qmctorch`` package.  It is used in the build
container (where ``/root/reference`` is mounted) by ``oracle/make_golden.py`` to
pin ``oracle/sj_oracle.py`` and to write ``tests/golden/*.npz``.  Nothing under
``qmctorch_b200/`` imports it; ``bench.py`` uses it only for the CPU legs (the reference arm and
``cpu_baseline``), from ``baseline/_ref`` on the GPU box.
"""

import os
import sys
import types

# /root/reference: the build container.  baseline/_ref: `pip install --no-deps --target baseline/_ref` of the
# unmodified reference - git-ignored, but it travels to the GPU box with the snapshot,
# so bench.py's CPU legs can time the REAL reference there (cpu_baseline.kind = "reference").
REF_ROOTS = ["/root/reference", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (_Anything,), {})


class _Anything(metaclass=_AnyMeta):
    """Instances/classes that swallow every call and attribute access."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        cls = _AnyMeta(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


_STUBS = [
    "twiggy", "h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mendeleev", "pints", "pyscf",
    "scm", "scm.plams", "plams", "ase", "ase.calculators",
    "ase.calculators.calculator", "ase.optimize", "ase.optimize.optimize",
    "ase.io", "ase.units",
]


def available():
    return any(os.path.isdir(os.path.join(r, "qmctorch")) for r in REF_ROOTS)


def load_reference():
    """Returns the reference ``qmctorch`` package (FP64 default dtype set)."""
    root = next((r for r in REF_ROOTS if os.path.isdir(os.path.join(r, "qmctorch"))), None)
    if root is None:
        raise RuntimeError("reference tree not found (expected /root/reference or baseline/_ref)")
    for name in _STUBS:
        if name not in sys.modules:
            mod = _StubModule(name)
            mod.__path__ = []
            sys.modules[name] = mod
    for name in _STUBS:
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    tw = sys.modules["twiggy"]
    tw.quick_setup = lambda *a, **k: None
    tw.log = _Anything()
    tw.levels = _Anything()
    sys.modules["ase.calculators.calculator"].Calculator = object
    sys.modules["ase.calculators.calculator"].all_changes = []
    sys.modules["ase.optimize.optimize"].Optimizer = object
    sys.modules["pints"].LogPDF = object
    if root not in sys.path:
        sys.path.insert(0, root)
    import qmctorch  # noqa
    from qmctorch.utils import set_torch_double_precision
    set_torch_double_precision()
    import qmctorch.solver.solver_base as sb
    import qmctorch.solver.solver as ss
    for m in (sb, ss):
        if hasattr(m, "dump_to_hdf5"):
            m.dump_to_hdf5 = lambda *a, **k: "grp"
        if hasattr(m, "add_group_attr"):
            m.add_group_attr = lambda *a, **k: None
    return qmctorch
