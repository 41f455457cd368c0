"""Walker initialisation (qmctorch/sampler/walkers.py:8-150): same torch generator calls in
the same order, so ``torch.manual_seed`` reproduces the reference's initial ensemble."""
import numpy as np
import torch
from torch.distributions import MultivariateNormal


class Walkers:
    def __init__(self, nwalkers=100, nelec=1, ndim=3, init=None, cuda=False):
        self.nwalkers = nwalkers
        self.ndim = ndim
        self.nelec = nelec
        self.init_domain = init
        self.pos = None
        self.status = None
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")

    def initialize(self, pos=None):
        if pos is not None:
            if len(pos) > self.nwalkers:
                pos = pos[-self.nwalkers:, :]
            self.pos = pos
            return
        dom = self.init_domain
        if "center" in dom:
            self.pos = self._init_center()
        elif "min" in dom:
            self.pos = self._init_uniform()
        elif "mean" in dom:
            self.pos = self._init_multivar()
        elif "atom_coords" in dom:
            self.pos = self._init_atomic()
        else:
            raise ValueError("Init walkers not recognized")

    def _init_center(self):
        eps = 1e-3
        pos = -eps + 2 * eps * torch.rand(self.nwalkers, self.nelec * self.ndim)
        return pos.type(torch.float64).to(device=self.device)

    def _init_uniform(self):
        pos = torch.rand(self.nwalkers, self.nelec * self.ndim)
        pos *= self.init_domain["max"] - self.init_domain["min"]
        pos += self.init_domain["min"]
        return pos.type(torch.float64).to(device=self.device)

    def _init_multivar(self):
        multi = MultivariateNormal(torch.as_tensor(self.init_domain["mean"]),
                                   torch.as_tensor(self.init_domain["sigma"]))
        pos = multi.sample((self.nwalkers, self.nelec)).type(torch.float64)
        return pos.view(self.nwalkers, self.nelec * self.ndim).to(device=self.device)

    def _init_atomic(self):
        """walkers.py:115-150 (host loop; runs once per sampler)."""
        dom = self.init_domain
        pos = torch.zeros(self.nwalkers, self.nelec * self.ndim, dtype=torch.float64)
        idx_ref = []
        for iat, n in enumerate(dom["atom_nelec"]):
            idx_ref += [iat] * n
        ntot = len(idx_ref)
        coords = torch.as_tensor(np.asarray(dom["atom_coords"]), dtype=torch.float64)
        for iw in range(self.nwalkers):
            placed = [0] * len(dom["atom_nelec"])
            idx = torch.as_tensor(idx_ref)[torch.randperm(ntot)]
            xyz = coords[idx, :].clone()
            for ie in range(ntot):
                a = int(idx[ie])
                if placed[a] == 0:
                    s = 1.0 / dom["atom_num"][a]
                elif placed[a] < 5:
                    s = 2.0 / (dom["atom_num"][a] - 2)
                else:
                    s = 3.0 / (dom["atom_num"][a] - 3)
                xyz[ie, :] += torch.as_tensor(np.random.normal(scale=s, size=(3,)))
                placed[a] += 1
            pos[iw, :] = xyz.view(-1)[: self.nelec * self.ndim]
        return pos.to(device=self.device)
