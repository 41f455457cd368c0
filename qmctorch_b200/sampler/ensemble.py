"""The walker ensemble of a sampler: positions ``pos [nwalkers, nelec*ndim]`` and their initial
distribution.

Behavioural contract taken from qmctorch/sampler/walkers.py:41-150: the start distribution is
selected by the keys of the ``init`` dictionary that ``mol.domain(method)`` returns
("center" | "min"/"max" | "mean"/"sigma" | "atom_coords"), and every draw is made with the same
torch / numpy generator calls in the same order, so a given ``torch.manual_seed`` yields the
reference's initial ensemble.  The table below maps a distinguishing key to its generator.
"""
import numpy as np
import torch
from torch.distributions import MultivariateNormal

F64 = torch.float64


def _around_center(w):
    """all electrons within +-1e-3 of the centre: one ``torch.rand`` of the full shape."""
    half_width = 1e-3
    u = torch.rand(w.nwalkers, w.nelec * w.ndim)
    return 2 * half_width * u - half_width


def _uniform_box(w):
    """uniform in [min, max]^d: one ``torch.rand`` of the full shape."""
    lo, hi = w.init_domain["min"], w.init_domain["max"]
    u = torch.rand(w.nwalkers, w.nelec * w.ndim)
    u *= hi - lo
    u += lo
    return u


def _gaussian_cloud(w):
    """each electron ~ N(mean, sigma): one MultivariateNormal sample of shape (W, Ne)."""
    law = MultivariateNormal(torch.as_tensor(w.init_domain["mean"]), torch.as_tensor(w.init_domain["sigma"]))
    return law.sample((w.nwalkers, w.nelec)).type(F64).reshape(w.nwalkers, w.nelec * w.ndim)


def _around_atoms(w):
    """electrons distributed over the atoms, shell by shell.  Per walker: one ``torch.randperm``
    to shuffle the atom of each electron, then one ``numpy.random.normal`` triple per electron
    with a width set by how many electrons that atom already carries."""
    dom = w.init_domain
    owner = np.repeat(np.arange(len(dom["atom_nelec"])), dom["atom_nelec"])
    nuclei = torch.as_tensor(np.asarray(dom["atom_coords"]), dtype=F64)
    charge = dom["atom_num"]
    out = torch.zeros(w.nwalkers, w.nelec * w.ndim, dtype=F64)
    for iw in range(w.nwalkers):
        order = torch.as_tensor(owner)[torch.randperm(len(owner))]
        xyz = nuclei[order].clone()
        filled = np.zeros(len(charge), dtype=int)
        for k, a in enumerate(order.tolist()):
            if filled[a] == 0:
                width = 1.0 / charge[a]
            elif filled[a] < 5:
                width = 2.0 / (charge[a] - 2)
            else:
                width = 3.0 / (charge[a] - 3)
            xyz[k] += torch.as_tensor(np.random.normal(scale=width, size=(1, 3)))[0]
            filled[a] += 1
        out[iw] = xyz.reshape(-1)[: w.nelec * w.ndim]
    return out


# first key of the init dictionary that identifies the start distribution
_START = (("center", _around_center), ("min", _uniform_box), ("mean", _gaussian_cloud),
          ("atom_coords", _around_atoms))


class Walkers:
    def __init__(self, nwalkers=100, nelec=1, ndim=3, init=None, cuda=False):
        self.nwalkers, self.nelec, self.ndim = nwalkers, nelec, ndim
        self.init_domain = init
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.pos = None
        self.status = None

    def initialize(self, pos=None):
        """Start from ``pos`` (its last ``nwalkers`` rows) or draw a fresh ensemble."""
        if pos is not None:
            self.pos = pos[-self.nwalkers:, :] if len(pos) > self.nwalkers else pos
            return
        for key, make in _START:
            if key in self.init_domain:
                self.pos = make(self).type(F64).to(self.device)
                return
        raise ValueError("Init walkers not recognized")
