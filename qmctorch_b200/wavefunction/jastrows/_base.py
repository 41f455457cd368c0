"""Shared forward logic of the Jastrow factor operators (qmcb_jastrow)."""
import torch

from ... import _lib
from .._plan import as_walkers


def jastrow_forward(module, handle, which, pos, derivative, sum_grad):
    """Returns J [W,1] | dJ [W,nelec] or [W,3,nelec] | d2J [W,nelec] | (J, dJ, d2J) with the
    reference's shapes (jastrow_factor_electron_electron.py:124-175)."""
    if derivative not in (0, 1, 2, [0, 1, 2]):
        raise ValueError("derivative not understood")
    dev = handle.ao.atom_coords.device
    nelec = module.nelec
    x = as_walkers(pos, 3 * nelec, dev)
    W = x.shape[0]
    L = _lib.lib()
    J = torch.empty(W, dtype=torch.float64, device=dev)
    dJ = d2J = None
    if derivative != 0:
        dJ = torch.empty(W, 3, nelec, dtype=torch.float64, device=dev)
        d2J = torch.empty(W, nelec, dtype=torch.float64, device=dev)
    _lib.check(L.qmcb_jastrow(handle.plan(), _lib.ptr(x), W, which, _lib.ptr(J), _lib.ptr(dJ),
                              _lib.ptr(d2J), _lib.stream_ptr(dev)), "qmcb_jastrow")
    J = J.unsqueeze(-1)
    if derivative == 0:
        return J
    if derivative == 1:
        return dJ.sum(1) if sum_grad else dJ
    if derivative == 2:
        return d2J
    return J, (dJ.sum(1) if sum_grad else dJ), d2J
