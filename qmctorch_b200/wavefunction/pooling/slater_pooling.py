"""SlaterPooling operator (qmctorch/wavefunction/pooling/slater_pooling.py:18-387).

``forward(mo)`` -> D_up * D_down per configuration, ``operator(mo, bop)`` ->
Tr(A_up^-1 B_up) + Tr(A_down^-1 B_down) per configuration.  Every configuration is
evaluated through its (deduplicated) explicit spin determinants in ``qmcb_slater``;
``tests/wavefunction/pooling/test_slater.py:44-78`` of the reference asserts that the
explicit and the single/double update routes agree, so one route suffices."""
import operator as _op

import torch
from torch import nn

from ... import _lib


class SlaterPooling(nn.Module):
    def __init__(self, config_method, configs, mol, cuda=False):
        super().__init__()
        self.config_method = config_method
        self.configs = configs
        self.nconfs = len(configs[0])
        self.nmo = mol.basis.nmo
        self.nup = mol.nup
        self.ndown = mol.ndown
        self.nelec = self.nup + self.ndown
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self._handle = None

    def _run(self, mo, bop):
        if self._handle is None:
            raise RuntimeError("SlaterPooling must be attached to a SlaterJastrow")
        dev = self._handle.ao.atom_coords.device
        x = mo.detach().to(device=dev, dtype=torch.float64).contiguous()
        W = x.shape[0]
        nmo = x.shape[-1]
        if nmo != self._handle.mo.mo_scf.shape[1]:
            raise ValueError("mo has %d columns, expected %d" % (nmo, self._handle.mo.mo_scf.shape[1]))
        L = _lib.lib()
        dets = torch.empty(W, self.nconfs, dtype=torch.float64, device=dev)
        trace = None
        b = None
        nop = 0
        if bop is not None:
            b = bop.detach().to(device=dev, dtype=torch.float64).contiguous()
            nop = 1 if b.dim() == 3 else b.shape[0]
            trace = torch.empty(nop, W, self.nconfs, dtype=torch.float64, device=dev)
        _lib.check(L.qmcb_slater(self._handle.plan(), _lib.ptr(x), _lib.ptr(b), nop, W, _lib.ptr(dets),
                                 _lib.ptr(trace), _lib.stream_ptr(dev)), "qmcb_slater")
        if trace is not None and bop.dim() == 3:
            trace = trace[0]
        return dets, trace

    def forward(self, input):
        """mo [W, nelec, nmo] -> [W, nconf]  (slater_pooling.py:64-80)."""
        return self._run(input, None)[0]

    def det_explicit(self, input):
        return self.forward(input)

    def det_single_double(self, input):
        return self.forward(input)

    def operator(self, mo, bop, op=_op.add, op_squared=False, inv_mo=None):
        """slater_pooling.py:262-306 with op=operator.add (the only use outside backflow)."""
        if op is not _op.add or op_squared or inv_mo is not None:
            raise NotImplementedError(
                "only op=operator.add without op_squared/inv_mo is on the CUDA path "
                "(the other variants serve backflow, slater_jastrow.py:484-580)")
        return self._run(mo, bop)[1]

    def operator_explicit(self, mo, bkin, op=_op.add, op_squared=False):
        return self.operator(mo, bkin, op, op_squared)

    def operator_single_double(self, mo, bop, op=_op.add, op_squared=False, inv_mo=None):
        return self.operator(mo, bop, op, op_squared, inv_mo)
