"""Samplers that need the gradient of the density: drift-diffusion (generalized) Metropolis and
Hamiltonian Monte Carlo.

The reference differentiates ``pdf`` with autograd at every proposal
(sampler/generalized_metropolis.py:188-205, sampler/hamiltonian.py:45-66).  Here the gradient comes
from ``pdf(x, return_grad=True)`` = the analytic grad psi^2 of ``qmcb_grad_psi`` (slater_jastrow.py:346-447,
row a16) whenever ``pdf`` is the method of a wave function; any other callable is differentiated
with autograd exactly like the reference does.

Both samplers reproduce the reference's update rules *as coded*, including the ones that differ
from the textbook, so that the same ``torch.manual_seed`` gives the reference's chain:

* ``GeneralizedMetropolis.move`` proposes from the ensemble's INITIAL positions at every step
  (``self.walkers.pos`` is only rebound after the loop, generalized_metropolis.py:127-140,105);
* its transition density uses the norm, not the squared norm (``:185-186``);
* its proposal covariance is ``sqrt(step_size) * I`` (``:165-167``).

Random draws are made on the CPU generator in the reference's order and shipped to the device of
the walkers.

Deliberate restatements: ``GeneralizedMetropolis.__call__/move/_move/trans`` and ``Hamiltonian.__call__`` track
generalized_metropolis.py:60-186 and hamiltonian.py:68-150 variable for variable for that reason; what is new
here is ``_density_and_gradient`` (the analytic gradient through the CUDA path).
"""
import numpy as np
import torch
from torch.distributions import MultivariateNormal
from tqdm import tqdm

from .metropolis import SamplerBase


def _density_and_gradient(pdf, x):
    """(rho [W], grad rho [W, D]) - analytic when ``pdf`` belongs to a wave function."""
    owner = getattr(pdf, "__self__", None)
    if owner is not None and hasattr(owner, "gradients_jacobi"):
        with torch.no_grad():
            return pdf(x).reshape(-1), pdf(x, return_grad=True)
    with torch.enable_grad():
        xg = x.detach().clone().requires_grad_(True)
        rho = pdf(xg).reshape(-1)
        (g,) = torch.autograd.grad(rho, xg, grad_outputs=torch.ones_like(rho))
    return rho.detach(), g.detach()


class GeneralizedMetropolis(SamplerBase):
    def __init__(self, nwalkers=100, nstep=1000, step_size=3, ntherm=-1, ndecor=1, nelec=1, ndim=1,
                 init={"type": "uniform", "min": -5, "max": 5}, cuda=False):
        """sampler/generalized_metropolis.py:13-45 (same arguments and defaults)."""
        SamplerBase.__init__(self, nwalkers, nstep, step_size, ntherm, ndecor, nelec, ndim, init, cuda)
        self.acceptance_rate = None

    def __call__(self, pdf, pos=None, with_tqdm=True):
        """sampler/generalized_metropolis.py:47-123."""
        with torch.no_grad():
            if self.ntherm < 0:
                self.ntherm = self.nstep + self.ntherm
            self.walkers.initialize(pos=pos)
            xi = self.walkers.pos.detach().clone()
            rhoi = pdf(xi).reshape(-1).clone()
            drifti = self.get_drift(pdf, xi)
            rhoi[rhoi == 0] = 1e-16
            kept, rate, idecor = [], 0.0, 0
            for istep in tqdm(range(self.nstep), desc="INFO:QMCTorch|  Sampling", disable=not with_tqdm):
                xf = self.move(drifti)
                rhof = pdf(xf).reshape(-1).clone()
                driftf = self.get_drift(pdf, xf)
                rhof[rhof == 0.0] = 1e-16
                t_if = self.trans(xi, xf, driftf)
                t_fi = self.trans(xf, xi, drifti)
                index = self._accept((t_if * rhof) / (t_fi * rhoi).double())
                rate += float(index.sum()) / self.walkers.nwalkers
                xi[index, :] = xf[index, :]
                rhoi[index] = rhof[index]
                rhoi[rhoi == 0] = 1e-16
                drifti[index, :] = driftf[index, :]
                if istep >= self.ntherm:
                    if idecor % self.ndecor == 0:
                        kept.append(xi.clone().detach())
                    idecor += 1
            self.acceptance_rate = rate / max(self.nstep, 1)
            self.walkers.pos = xi
        return torch.cat(kept).requires_grad_()

    def move(self, drift):
        """One random electron per walker, displaced from the walkers' positions as stored in
        ``self.walkers.pos`` (generalized_metropolis.py:125-143)."""
        nw = self.walkers.nwalkers
        new_pos = self.walkers.pos.clone().view(nw, self.nelec, self.ndim)
        index = torch.LongTensor(nw).random_(0, self.nelec)
        rows = torch.arange(nw, device=new_pos.device)
        new_pos[rows, index.to(new_pos.device), :] += self._move(drift, index)
        return new_pos.view(nw, self.nelec * self.ndim)

    def _move(self, drift, index):
        """step_size * drift_e + N(0, sqrt(step_size) I) (generalized_metropolis.py:145-171)."""
        nw = self.walkers.nwalkers
        d = drift.view(nw, self.nelec, self.ndim)
        mv = MultivariateNormal(torch.zeros(self.ndim), np.sqrt(self.step_size) * torch.eye(self.ndim))
        noise = mv.sample((nw, 1)).squeeze().to(d.device, d.dtype)
        rows = torch.arange(nw, device=d.device)
        return self.step_size * d[rows, index.to(d.device), :] + noise

    def trans(self, xf, xi, drifti):
        """exp(-|xf - xi - step drift| / (2 step)) (generalized_metropolis.py:173-186)."""
        a = (xf - xi - drifti * self.step_size).norm(dim=1)
        return torch.exp(-0.5 * a / self.step_size)

    def get_drift(self, pdf, x):
        """drift velocity grad rho / (2 rho) (generalized_metropolis.py:188-205)."""
        rho, g = _density_and_gradient(pdf, x)
        return 0.5 * g / rho.view(-1, 1)

    def _accept(self, P):
        """generalized_metropolis.py:207-219."""
        P[P > 1] = 1.0
        tau = torch.rand(self.walkers.nwalkers, dtype=torch.float64).to(P.device)
        return (P - tau >= 0).reshape(-1)


class Hamiltonian(SamplerBase):
    def __init__(self, nwalkers=100, nstep=100, step_size=0.2, L=10, ntherm=-1, ndecor=1, nelec=1, ndim=3,
                 init={"min": -5, "max": 5}, cuda=False):
        """sampler/hamiltonian.py:10-43 (same arguments and defaults)."""
        SamplerBase.__init__(self, nwalkers, nstep, step_size, ntherm, ndecor, nelec, ndim, init, cuda)
        self.traj_length = L
        self.acceptance_rate = None

    @staticmethod
    def log_func(func):
        """U = -log pdf (hamiltonian.py:68-78)."""
        return lambda x: -torch.log(func(x))

    @staticmethod
    def _potential_gradient(pdf, q):
        """grad U = -grad rho / rho (what hamiltonian.py:45-66 gets from autograd on -log pdf)."""
        rho, g = _density_and_gradient(pdf, q)
        return -g / rho.view(-1, 1)

    def __call__(self, pdf, pos=None, with_tqdm=True):
        """hamiltonian.py:80-139."""
        if self.ntherm < 0:
            self.ntherm = self.nstep + self.ntherm
        self.walkers.initialize(pos=pos)
        self.walkers.pos = self.walkers.pos.detach().clone()
        kept, rate, idecor = [], 0.0, 0
        with torch.no_grad():
            for istep in tqdm(range(self.nstep), desc="INFO:QMCTorch|  Sampling", disable=not with_tqdm):
                self.walkers.pos, r = self._step(pdf, self.step_size, self.traj_length, self.walkers.pos)
                rate += r
                if istep >= self.ntherm:
                    if idecor % self.ndecor == 0:
                        kept.append(self.walkers.pos)
                    idecor += 1
        self.acceptance_rate = rate / max(self.nstep, 1)
        return torch.cat(kept).requires_grad_()

    @classmethod
    def _step(cls, pdf, epsilon, L, q_init):
        """One leapfrog trajectory + accept test (hamiltonian.py:141-201)."""
        U = cls.log_func(pdf)
        q = q_init.clone()
        p = torch.randn(q.shape).to(q.device, q.dtype)
        e_init = U(q).reshape(-1) + 0.5 * (p * p).sum(1)
        p -= 0.5 * epsilon * cls._potential_gradient(pdf, q)
        for _ in range(L - 1):
            q += epsilon * p
            p -= epsilon * cls._potential_gradient(pdf, q)
        q += epsilon * p
        p -= 0.5 * epsilon * cls._potential_gradient(pdf, q)
        p = -p
        e_new = U(q).reshape(-1) + 0.5 * (p * p).sum(1)
        eps = torch.rand(e_new.shape).to(q.device, q.dtype)
        rejected = torch.exp(e_init - e_new) < eps
        q[rejected] = q_init[rejected]
        return q, 1.0 - float(rejected.sum()) / rejected.shape[0]
