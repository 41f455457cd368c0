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


# ---- the same four start distributions drawn ON THE DEVICE (init_rng="philox": torch's CUDA
# generator is Philox4x32-10).  Same laws as above, different draws: an opt-in for large ensembles,
# where the reference-identical CPU draws dominate a run (1e6 LiH walkers: 0.25 s of a 0.30 s
# single point).  g = a torch.Generator on w.device.
def _dev_center(w, g):
    return 2e-3 * torch.rand(w.nwalkers, w.nelec * w.ndim, dtype=F64, device=w.device, generator=g) - 1e-3


def _dev_box(w, g):
    lo, hi = w.init_domain["min"], w.init_domain["max"]
    return lo + (hi - lo) * torch.rand(w.nwalkers, w.nelec * w.ndim, dtype=F64, device=w.device, generator=g)


def _dev_gauss(w, g):
    mean = torch.as_tensor(w.init_domain["mean"], dtype=F64, device=w.device)
    chol = torch.linalg.cholesky(torch.as_tensor(w.init_domain["sigma"], dtype=F64)).to(w.device)
    z = torch.randn(w.nwalkers, w.nelec, w.ndim, dtype=F64, device=w.device, generator=g)
    return (mean + z @ chol.T).reshape(w.nwalkers, w.nelec * w.ndim)


def _dev_atoms(w, g):
    """walkers.py:115-150 without the per-walker Python loop: a random permutation of the electron ->
    atom assignment per walker (argsort of uniforms), the number of electrons the atom already
    carries by a cumulative one-hot count, then one normal triple per electron."""
    dom = w.init_domain
    dev = w.device
    owner = torch.as_tensor(np.repeat(np.arange(len(dom["atom_nelec"])), dom["atom_nelec"]), device=dev)
    nuclei = torch.as_tensor(np.asarray(dom["atom_coords"]), dtype=F64, device=dev)
    charge = torch.as_tensor(np.asarray(dom["atom_num"], dtype=np.float64), device=dev)
    ne = len(owner)
    perm = torch.rand(w.nwalkers, ne, device=dev, generator=g).argsort(dim=1)
    order = owner[perm]                                                  # [W, ne] atom of each electron
    hot = torch.nn.functional.one_hot(order, len(charge))
    filled = ((hot.cumsum(1) - hot) * hot).sum(-1)                      # electrons already on that atom
    z = charge[order]
    width = torch.where(filled == 0, 1.0 / z, torch.where(filled < 5, 2.0 / (z - 2), 3.0 / (z - 3)))
    xyz = nuclei[order] + width[..., None] * torch.randn(w.nwalkers, ne, 3, dtype=F64, device=dev, generator=g)
    return xyz.reshape(w.nwalkers, -1)[:, : w.nelec * w.ndim]


_START_DEV = (("center", _dev_center), ("min", _dev_box), ("mean", _dev_gauss), ("atom_coords", _dev_atoms))


class Walkers:
    def __init__(self, nwalkers=100, nelec=1, ndim=3, init=None, cuda=False, init_rng="torch"):
        self.nwalkers, self.nelec, self.ndim = nwalkers, nelec, ndim
        self.init_domain = init
        self.cuda = cuda
        if init_rng not in ("torch", "philox"):
            raise ValueError("init_rng should be 'torch' (the reference's CPU generator calls) or 'philox'")
        self.init_rng = init_rng
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.pos = None
        self.status = None

    def initialize(self, pos=None):
        """Start from ``pos`` (its last ``nwalkers`` rows) or draw a fresh ensemble."""
        if pos is not None:
            self.pos = pos[-self.nwalkers:, :] if len(pos) > self.nwalkers else pos
            return
        from ..solver.distributed import is_distributed, rank_seed
        if self.cuda and self.device.type != "cuda":      # Solver switches .cuda on after construction
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.init_rng == "philox":
            g = torch.Generator(device=self.device)
            g.manual_seed(rank_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF))
            for key, make in _START_DEV:
                if key in self.init_domain:
                    self.pos = make(self, g).contiguous()
                    return
            raise ValueError("Init walkers not recognized")
        for key, make in _START:
            if key in self.init_domain:
                if is_distributed():
                    # same manual_seed on every rank (usual practice) must not give identical shards
                    with torch.random.fork_rng(devices=[]):
                        torch.manual_seed(rank_seed(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF))
                        state = np.random.get_state()
                        np.random.seed(rank_seed(torch.initial_seed()) % (2 ** 32))
                        self.pos = make(self).type(F64).to(self.device)
                        np.random.set_state(state)
                else:
                    self.pos = make(self).type(F64).to(self.device)
                return
        raise ValueError("Init walkers not recognized")
