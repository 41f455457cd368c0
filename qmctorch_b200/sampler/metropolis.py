"""Metropolis sampler - drop-in for qmctorch.sampler.Metropolis
(qmctorch/sampler/metropolis.py:10-298, sampler_base.py:8-91).

When ``pdf`` is the ``pdf`` method of a qmctorch_b200 SlaterJastrow on a CUDA device, each
move is ONE fused kernel (propose, psi, accept, in-place update: ``qmcb_metropolis_step``)
and the ensemble never leaves the GPU.  Draws come from an in-kernel Philox generator
(``rng="philox"``, default) or, for parity runs, from the same torch generator calls the
reference makes (``rng="torch"``: MultivariateNormal on the CPU generator, ``rand`` for the
acceptance draw).  For any other callable the reference's generic torch loop is used.

Deliberate restatements of the reference (they fix the ORDER of the generator calls, which is what makes
``rng="torch"`` chains and the arbitrary-callable fallback reproduce the reference's from the same seed):
``configure_move``, ``move``, ``_move``, ``_accept`` and ``_call_generic`` follow metropolis.py:179-298
statement by statement, print strings included.  The fused path inside ``__call__`` and ``_torch_draws`` are original.
"""
import math
from time import time

import torch
from torch.distributions import MultivariateNormal
from tqdm import tqdm

from .. import _lib
from .ensemble import Walkers


class SamplerBase:
    def __init__(self, nwalkers, nstep, step_size, ntherm, ndecor, nelec, ndim, init, cuda, init_rng="torch"):
        self.nelec = nelec
        self.ndim = ndim
        self.nstep = nstep
        self.step_size = step_size
        self.ntherm = ntherm
        self.ndecor = ndecor
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.walkers = Walkers(nwalkers=nwalkers, nelec=nelec, ndim=ndim, init=init, cuda=cuda, init_rng=init_rng)

    def __call__(self, pdf, *args, **kwargs):
        raise NotImplementedError("Sampler must have a __call__ method")

    def __repr__(self):
        return self.__class__.__name__ + " sampler with  %d walkers" % self.walkers.nwalkers

    def get_sampling_size(self):
        """sampler_base.py:86-91."""
        if self.ntherm == -1:
            return self.walkers.nwalkers
        return self.walkers.nwalkers * int((self.nstep - self.ntherm) / self.ndecor)


class Metropolis(SamplerBase):
    def __init__(self, nwalkers=100, nstep=1000, step_size=0.2, ntherm=-1, ndecor=1, nelec=1, ndim=3,
                 init={"min": -5, "max": 5}, move={"type": "all-elec", "proba": "normal"}, logspace=False,
                 symmetry=None, cuda=False, rng="philox", seed=None, keep_on_device=False, init_rng="torch"):
        SamplerBase.__init__(self, nwalkers, nstep, step_size, ntherm, ndecor, nelec, ndim, init, cuda, init_rng)
        self.logspace = logspace
        self.configure_move(move)
        self.symmetry = (lambda x: x) if symmetry is None else symmetry
        if rng not in ("philox", "torch"):
            raise ValueError("rng should be 'philox' or 'torch'")
        self.rng = rng
        self.seed = seed
        self.keep_on_device = keep_on_device
        self._step_counter = 0
        self.acceptance_rate = None

    host_work = None        # optional callable run between the launch of the moves and the first device wait

    def configure_move(self, move):
        """metropolis.py:179-225."""
        self.movedict = move
        if "type" not in self.movedict:
            print("Metroplis : Set 1 electron move by default")
            self.movedict["type"] = "one-elec"
        if "proba" not in self.movedict:
            print("Metroplis : Set uniform trial move probability")
            self.movedict["proba"] = "uniform"
        if self.movedict["proba"] == "normal":
            # NB: the reference hands _sigma to MultivariateNormal as the COVARIANCE
            # (metropolis.py:207-212), so the proposal std is sqrt(_sigma); kept.
            self._sigma = self.step_size / (2 * math.sqrt(2 * math.log(2.0)))
            _sigma = self.step_size / (2 * torch.sqrt(2 * torch.log(torch.as_tensor(2.0))))
            self.multiVariate = MultivariateNormal(torch.zeros(self.ndim), _sigma * torch.eye(self.ndim))
        self._move_per_iter = 1
        if self.movedict["type"] not in ["one-elec", "all-elec", "all-elec-iter"]:
            raise ValueError(" 'type' in move should be 'one-elec','all-elec', 'all-elec-iter'")
        if self.movedict["type"] == "all-elec-iter":
            self.fixed_id_elec_list = range(self.nelec)
            self._move_per_iter = self.nelec
        else:
            self.fixed_id_elec_list = [None]

    # ---------------------------------------------------------------------------------------
    def __call__(self, pdf, pos=None, with_tqdm=True):
        """metropolis.py:85-177.  Returns the kept walker positions [W*nkept, 3*nelec]
        (CPU tensor with requires_grad, like the reference; device tensor if
        keep_on_device=True)."""
        eps = 1e-16
        if self.ntherm >= self.nstep:
            raise ValueError("Thermalisation longer than trajectory")
        wf = getattr(pdf, "__self__", None)
        fused = (wf is not None and getattr(pdf, "__name__", "") == "pdf" and hasattr(wf, "_handle")
                 and not self.logspace and wf.ao.atom_coords.device.type == "cuda")
        if not fused:
            return self._call_generic(pdf, pos, with_tqdm, eps)
        with torch.no_grad():
            if self.ntherm < 0:
                self.ntherm = self.nstep + self.ntherm
            self.walkers.initialize(pos=pos)
            dev = wf.ao.atom_coords.device
            x = self.walkers.pos.detach().to(device=dev, dtype=torch.float64).contiguous().clone()
            self.walkers.pos = x
            W = x.shape[0]
            fx = wf._psi(x).reshape(-1) ** 2
            fx[fx == 0] = eps
            naccept = torch.zeros(1, dtype=torch.int64, device=dev)
            L = _lib.lib()
            plan = wf._handle.plan()
            normal = self.movedict["proba"] == "normal"
            scale = math.sqrt(self._sigma) if normal else self.step_size
            # the kernel indexes its Philox draws by the rank-LOCAL walker index: fold the rank into the
            # seed so that shards are independent even when every rank called the same manual_seed
            from ..solver.distributed import rank_seed
            seed = rank_seed(self.seed if self.seed is not None else int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF))
            kept, idecor = [], 0
            tstart = time()
            for istep in tqdm(range(self.nstep), desc="INFO:QMCTorch|  Sampling", disable=not with_tqdm):
                for id_elec in self.fixed_id_elec_list:
                    disp = tau = eidx = None
                    if self.movedict["type"] == "all-elec" or self.nelec == 1:
                        move_elec = -1
                    elif id_elec is None:
                        move_elec = -2
                    else:
                        move_elec = int(id_elec)
                    if self.rng == "torch":
                        disp, tau, eidx = self._torch_draws(W, move_elec, dev)
                    _lib.check(L.qmcb_metropolis_step(
                        plan, _lib.ptr(x), _lib.ptr(fx), W, _lib.ptr(disp), _lib.ptr(tau), _lib.ptr(eidx),
                        move_elec, int(normal), scale, eps, seed, self._step_counter, None,
                        _lib.ptr(naccept), _lib.stream_ptr(dev)), "qmcb_metropolis_step")
                    self._step_counter += 1
                if istep >= self.ntherm:
                    if idecor % self.ndecor == 0:
                        kept.append(x.clone() if self.keep_on_device else x.to("cpu"))
                    idecor += 1
            if self.host_work is not None:
                # host-side work of the caller (Solver: turning the previous E_L batch into its numpy observable)
                # while the moves queued above run on the device; the counter read-back below is the first wait
                self.host_work()
            self.acceptance_rate = float(naccept.item()) / max(W * self._move_per_iter * max(self.nstep, 1), 1)
            self.sampling_time = time() - tstart
        out = self.symmetry(torch.cat(kept))
        return out.requires_grad_()

    def _torch_draws(self, W, move_elec, dev):
        """Same generator calls, same order as the reference (metropolis.py:247,266-275,295)."""
        nmove = self.nelec if move_elec == -1 else 1
        eidx = None
        if move_elec == -2:
            eidx = torch.LongTensor(W).random_(0, self.nelec).to(torch.int32).to(dev)
        # all draws on the CPU generator: replays the reference's CPU run draw for draw
        if self.movedict["proba"] == "uniform":
            d = torch.rand((W, nmove, self.ndim), dtype=torch.float64).view(W, nmove * self.ndim)
            d = (self.step_size * (2.0 * d - 1.0)).to(dev)
        else:
            d = self.multiVariate.sample((W, nmove)).to(torch.float64).view(W, nmove * self.ndim).to(dev)
        if nmove != self.nelec:
            full = torch.zeros(W, self.nelec, self.ndim, device=dev, dtype=torch.float64)
            if move_elec >= 0:
                full[:, move_elec, :] = d
            else:
                full[torch.arange(W, device=dev), eidx.long(), :] = d
            d = full.view(W, -1)
        tau = torch.rand(W, dtype=torch.float64).to(dev)
        return d.contiguous(), tau, eidx

    # ---------------------------------------------------------------------------------------
    def _call_generic(self, pdf, pos, with_tqdm, eps):
        """Reference algorithm for an arbitrary ``pdf`` callable (torch ops on its device)."""
        with torch.no_grad():
            if self.ntherm < 0:
                self.ntherm = self.nstep + self.ntherm
            self.walkers.initialize(pos=pos)
            logf = (lambda x: torch.log(pdf(x))) if self.logspace else pdf
            fx = logf(self.walkers.pos)
            if not self.logspace:
                fx[fx == 0] = eps
            kept, rate, idecor = [], 0.0, 0
            for istep in tqdm(range(self.nstep), desc="INFO:QMCTorch|  Sampling", disable=not with_tqdm):
                for id_elec in self.fixed_id_elec_list:
                    xn = self.move(pdf, id_elec)
                    fxn = logf(xn)
                    if self.logspace:
                        df = fxn - fx
                    else:
                        fxn[fxn == 0.0] = eps
                        df = fxn / fx
                    index = self._accept(df)
                    rate += float(index.sum()) / (self.walkers.nwalkers * self._move_per_iter)
                    self.walkers.pos[index, :] = xn[index, :]
                    fx[index] = fxn[index]
                    if not self.logspace:
                        fx[fx == 0] = eps
                if istep >= self.ntherm:
                    if idecor % self.ndecor == 0:
                        kept.append(self.walkers.pos.to("cpu").clone())
                    idecor += 1
            self.acceptance_rate = rate / max(self.nstep, 1)
        return self.symmetry(torch.cat(kept)).requires_grad_()

    def move(self, pdf, id_elec):
        """metropolis.py:227-254."""
        nw = self.walkers.nwalkers
        if self.nelec == 1 or self.movedict["type"] == "all-elec":
            return self.walkers.pos + self._move(self.nelec)
        new_pos = self.walkers.pos.clone().view(nw, self.nelec, self.ndim)
        if id_elec is None:
            index = torch.LongTensor(nw).random_(0, self.nelec)
        else:
            index = torch.LongTensor(nw).fill_(id_elec)
        new_pos[range(nw), index, :] += self._move(1)
        return new_pos.view(nw, self.nelec * self.ndim)

    def _move(self, num_elec):
        """metropolis.py:256-277."""
        nw = self.walkers.nwalkers
        dev = self.walkers.pos.device
        if self.movedict["proba"] == "uniform":
            d = torch.rand((nw, num_elec, self.ndim), device=dev).view(nw, num_elec * self.ndim)
            return self.step_size * (2.0 * d - 1.0)
        d = self.multiVariate.sample((nw, num_elec)).to(dev)
        return d.view(nw, num_elec * self.ndim)

    def _accept(self, proba):
        """metropolis.py:279-298."""
        if self.logspace:
            proba[proba > 0] = 0.0
            tau = torch.log(torch.rand_like(proba))
        else:
            proba[proba > 1] = 1.0
            tau = torch.rand_like(proba)
        return (proba - tau >= 0).reshape(-1).type(torch.bool)
