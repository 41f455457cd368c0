"""Solver - drop-in for qmctorch.solver.Solver (qmctorch/solver/solver.py:15-431,
solver_base.py:12-546): single-point energies and wave-function optimisation with the
low-variance ("manual") energy-gradient estimator.

Orchestration stays Python; ``wf.local_energy``, ``wf(pos)`` and its backward are the fused
CUDA kernels.  Under torch.distributed every rank owns a shard of the walkers and the
statistics / gradients are summed with one all-reduce each (solver/distributed.py).
Results are returned and - when an ``output`` file is named - dumped in the layout of the reference's
``utils/hdf5_utils.py:dump_to_hdf5`` by the pure-Python writer ``utils/hdf5_write.py`` (rank 0 only).
"""
import os
from math import ceil
from time import time
from types import SimpleNamespace

import numpy as np
import torch

from .. import _lib
from . import distributed as D


class Loss:
    """qmctorch/solver/loss.py:7-153: energy or variance of the (sampling-weighted, clipped) local energies.
    ``wf.local_energy`` and ``wf(pos)`` carry autograd nodes backed by qmcb_local_energy_backward /
    qmcb_psi_backward, so ``loss.backward()`` (grad="auto") works like the reference's."""

    def __init__(self, wf, method="energy", clip=False, clip_threshold=5):
        if method not in ("energy", "variance", "weighted-energy", "weighted-variance"):
            raise ValueError("loss method should be energy, variance, weighted-energy or weighted-variance")
        self.wf = wf
        self.method = method
        self.clip = clip
        self.clip_num_std = clip_threshold
        self.use_weight = False
        # loss.py:43: only "energy" and "variance" index the reference's table (the weighted names raise a
        # KeyError there); the weighted variants are the same reductions with use_weight switched on
        self.loss_fn = torch.var if method.endswith("variance") else torch.mean
        self.weight = {"psi": None, "psi0": None}

    def __call__(self, pos, no_grad=False, deactivate_weight=False):
        """loss.py:49-82 -> (loss, local energies)."""
        with (torch.no_grad() if no_grad else torch.enable_grad()):
            local_energies = self.wf.local_energy(pos)
            mask = self.get_clipping_mask(local_energies)
            weight = self.get_sampling_weights(pos, deactivate_weight)
            loss = self.loss_fn((weight * local_energies)[mask])
        return loss, local_energies

    forward = __call__

    def get_sampling_weights(self, pos, deactivate_weight):
        """loss.py:112-153: (psi / psi0)^2, normalised, when the walkers are not resampled every epoch."""
        if not (self.use_weight and not deactivate_weight):
            return torch.tensor(1.0, dtype=torch.float64, device=self.wf.ao.atom_coords.device)
        self.weight["psi"] = self.wf(pos)
        if self.weight["psi0"] is None:
            self.weight["psi0"] = self.weight["psi"].detach().clone()
            return torch.ones_like(self.weight["psi"])
        w = (self.weight["psi"] / self.weight["psi0"]) ** 2
        return w / w.sum()

    def get_clipping_mask(self, eloc):
        if not self.clip:
            return torch.ones_like(eloc).type(torch.bool)
        if D.is_distributed():
            raise NotImplementedError("clip_loss needs a global median; keep it off under torch.distributed")
        median = torch.median(eloc)
        std = torch.std(eloc)
        return torch.abs((eloc - median) / std) < self.clip_num_std


class _Loader:
    """utils/torch_utils.py:156-207 (batches are views; walkers stay where they are)."""

    def __init__(self, data, batch_size):
        self.dataset = data
        self.batch_size = batch_size

    def __iter__(self):
        n = len(self.dataset)
        for i in range(ceil(n / self.batch_size)):
            yield self.dataset[i * self.batch_size: (i + 1) * self.batch_size]


class Solver:
    def __init__(self, wf=None, sampler=None, optimizer=None, scheduler=None, output=None, rank=0):
        self.wf = wf
        self.sampler = sampler
        self.opt = optimizer
        self.scheduler = scheduler
        self.rank = rank
        self.cuda = bool(wf.cuda)
        self.device = wf.ao.atom_coords.device
        self.dataloader = None
        self.loss = None
        self.obs_dict = None
        self.qmctorch_version = "qmctorch_b200"
        if self.opt is not None and "lpos_needed" not in self.opt.__dict__:
            self.opt.lpos_needed = False
        self.save_model = "model.pth"
        if self.cuda:
            self.sampler.cuda = True
            self.sampler.walkers.cuda = True
            # the ensemble stays on the device between sampling, E_L, backward and
            # resampling; the reference copies it to the host at every kept step (metropolis.py:164)
            if hasattr(self.sampler, "keep_on_device"):
                self.sampler.keep_on_device = True
        self.hdf5file = output
        # the reference always dumps to <molecule>_QMCTorch.hdf5 in the working directory; here a file is
        # written only when the caller names one (solver.write_hdf5 can be switched on afterwards)
        self.write_hdf5 = output is not None
        if output is None:
            base = os.path.basename(getattr(wf.mol, "hdf5file", "mol.hdf5")).split(".")[0]
            self.hdf5file = base + "_QMCTorch.hdf5"
        self._stats_ws = None
        self.set_params_requires_grad()
        self.configure(track=["local_energy"], freeze=None, loss="energy", grad="manual", ortho_mo=False,
                       clip_loss=False,
                       resampling={"mode": "update", "resample_every": 1, "nstep_update": 25})

    # -- configuration (solver.py:50-178) ------------------------------------------------------
    def configure(self, track=None, freeze=None, loss=None, grad=None, ortho_mo=None, clip_loss=False,
                  clip_threshold=5, resampling=None):
        self.set_params_requires_grad()
        self.freeze_params_list = freeze
        self.freeze_parameters(freeze)
        if track is not None:
            self.track_observable(track)
        if grad is not None:
            if grad not in ("manual", "auto"):
                raise ValueError("grad should be 'auto' or 'manual'")
            self.grad_method = grad
            self.evaluate_gradient = {"auto": self.evaluate_grad_auto, "manual": self.evaluate_grad_manual}[grad]
        if resampling is not None:
            self.configure_resampling(**resampling)
        if loss is not None:
            self.loss = Loss(self.wf, method=loss, clip=clip_loss, clip_threshold=clip_threshold)
            self.loss.use_weight = self.resampling_options.resample_every > 1
        self.ortho_mo = ortho_mo

    def set_params_requires_grad(self, wf_params=True, geo_params=False):
        self.wf.ao.bas_exp.requires_grad = wf_params
        self.wf.ao.bas_coeffs.requires_grad = wf_params
        for p in self.wf.mo.parameters():
            p.requires_grad = wf_params
        self.wf.fc.weight.requires_grad = wf_params
        if getattr(self.wf, "jastrow", None) is not None:
            for p in self.wf.jastrow.parameters():
                p.requires_grad = wf_params
        self.wf.ao.atom_coords.requires_grad = geo_params
        # geometry optimisation (ase/optimizer/torch_optim.py:126-133: geo_params=True + evaluate_grad_auto): the atom
        # coordinates join the autograd graph of E_L / psi (qmcb_local_energy_backward) only when asked for
        self.wf.atom_coords_grad = bool(geo_params)

    def freeze_parameters(self, freeze):
        if freeze is None:
            return
        if not isinstance(freeze, list):
            freeze = [freeze]
        for name in freeze:
            low = name.lower()
            if low == "ci":
                self.wf.fc.weight.requires_grad = False
            elif low == "mo":
                for p in self.wf.mo.parameters():
                    p.requires_grad = False
            elif low == "ao":
                self.wf.ao.bas_exp.requires_grad = False
                self.wf.ao.bas_coeffs.requires_grad = False
            elif low == "jastrow":
                for p in self.wf.jastrow.parameters():
                    p.requires_grad = False
            elif low == "backflow":
                pass
            else:
                raise ValueError("Valid arguments for freeze are :", ["ci", "mo", "ao", "jastrow", "backflow"])

    def configure_resampling(self, mode="update", resample_every=1, nstep_update=25, ntherm_update=-1,
                             increment={"every": None, "factor": None}):
        if mode not in ["never", "full", "update"]:
            raise ValueError(mode, "not a valid update method : ", ["never", "full", "update"])
        self.resampling_options = SimpleNamespace(mode=mode, resample_every=resample_every,
                                                  ntherm_update=ntherm_update, nstep_update=nstep_update,
                                                  increment=increment)

    def track_observable(self, obs_name):
        if not isinstance(obs_name, list):
            obs_name = list(obs_name)
        valid = ["energy", "local_energy", "geometry", "parameters", "gradients"]
        for name in obs_name:
            if name not in valid and not hasattr(self.wf, name):
                raise ValueError("Observable not recognized")
        self.observable = SimpleNamespace()
        self.observable.qmctorch_version = self.qmctorch_version
        obs_name = list(obs_name)
        for extra in ("energy", "geometry"):
            if extra not in obs_name:
                obs_name.append(extra)
        for k in obs_name:
            if k == "parameters":
                for key, p in self.wf.named_parameters():
                    if p.requires_grad:
                        setattr(self.observable, key, [])
            elif k == "gradients":
                for key, p in self.wf.named_parameters():
                    if p.requires_grad:
                        setattr(self.observable, key + ".grad", [])
            else:
                setattr(self.observable, k, [])
        self.observable.models = SimpleNamespace()

    def store_observable(self, pos, local_energy=None, ibatch=None, **kwargs):
        """solver_base.py:166-248."""
        if pos.device != self.device:
            pos = pos.to(self.device)
        for obs in list(self.observable.__dict__.keys()):
            if obs in ("qmctorch_version", "models"):
                continue
            if obs == "energy":
                if local_energy is None:
                    local_energy = self.wf.local_energy(pos)
                m = float(local_energy.mean())
                if ibatch is None or ibatch == 0:
                    self.observable.energy.append(m)
                else:
                    self.observable.energy[-1] *= ibatch / (ibatch + 1)
                    self.observable.energy[-1] += m / (ibatch + 1)
            elif obs == "local_energy":
                if local_energy is None:
                    continue
                if (ibatch is None or ibatch == 0) and self._defer_observables and local_energy.is_cuda \
                        and local_energy.numel() >= 65536:
                    # large device tensor inside run_epochs: start the D2H into a pinned ring slot now, turn it
                    # into the numpy array the list holds while the GPU runs the resampling kernels
                    # (Metropolis calls _flush_observables before it waits for the acceptance counter)
                    self._flush_observables()
                    self.observable.local_energy.append(None)
                    self._stage(local_energy, self.observable.local_energy, len(self.observable.local_energy) - 1)
                    continue
                self._flush_observables()
                data = self._to_numpy(local_energy)
                if ibatch is None or ibatch == 0:
                    self.observable.local_energy.append(data)
                else:
                    self.observable.local_energy[-1] = np.append(self.observable.local_energy[-1], data)
            elif obs.endswith(".grad"):
                continue
            elif obs in self.wf.state_dict():
                p = dict(self.wf.named_parameters()).get(obs)
                getattr(self.observable, obs).append(self.wf.state_dict()[obs].detach().cpu().numpy().copy())
                if obs + ".grad" in self.observable.__dict__:
                    g = p.grad if p is not None and p.grad is not None else torch.zeros_like(p.data)
                    getattr(self.observable, obs + ".grad").append(g.detach().cpu().numpy().copy())
            elif hasattr(self.wf, obs):
                data = getattr(self.wf, obs)(pos)
                if isinstance(data, torch.Tensor):
                    data = data.detach().cpu().numpy()
                if isinstance(data, list):
                    data = np.array(data)
                if ibatch is None or ibatch == 0:
                    getattr(self.observable, obs).append(data)
                else:
                    getattr(self.observable, obs)[-1] = np.append(getattr(self.observable, obs)[-1], data)

    # -- device -> numpy for tracked observables ------------------------------------------------
    _defer_observables = False       # set by run_epochs for the duration of the epoch loop
    _pending = None

    def _stage(self, t, target, index):
        """Asynchronous D2H of ``t`` into a slot of a two-deep pinned ring; ``target[index]`` receives the numpy
        array when _flush_observables runs."""
        ring = getattr(self, "_ring", None)
        if ring is None or ring[0].numel() < t.numel() or ring[0].dtype != t.dtype:
            ring = self._ring = [torch.empty(t.numel(), dtype=t.dtype).pin_memory() for _ in range(2)]
            self._ring_next = 0
        k = self._ring_next
        self._ring_next = 1 - k
        view = ring[k][: t.numel()].view(t.shape)
        view.copy_(t.detach(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        self._pending = (ev, view, target, index)

    def _flush_observables(self):
        """Completes a staged observable (waits for its copy, which finished long ago when this runs behind the
        resampling kernels, and copies it out of the ring)."""
        if self._pending is None:
            return
        ev, view, target, index = self._pending
        self._pending = None
        ev.synchronize()
        target[index] = view.numpy().copy()

    def _to_numpy(self, t):
        """Tracked observables are numpy arrays (solver_base.py:166-248).  Large device tensors go
        through a reused pinned staging buffer: one DMA + one host memcpy instead of a pageable D2H
        (8 MB of local energies per epoch: 4.3 -> 1.4 ms)."""
        t = t.detach()
        if t.device.type != "cuda" or t.numel() < 65536:
            return t.cpu().numpy().copy()
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = self._pinned = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
        view = buf[: t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return view.numpy().copy()

    # -- statistics -------------------------------------------------------------------------
    def _stats(self, eloc):
        """[sum, sum sq, n finite, n non-finite] on the device (qmcb_energy_stats), summed over
        ranks -> mean, variance, error."""
        L = _lib.lib()
        dev = eloc.device
        flat = eloc.detach().reshape(-1).contiguous()
        if self._stats_ws is None or self._stats_ws.device != dev:
            self._stats_ws = torch.empty(int(L.qmcb_stats_workspace_bytes(flat.numel())), dtype=torch.uint8,
                                         device=dev)
        out4 = torch.empty(4, dtype=torch.float64, device=dev)
        _lib.check(L.qmcb_energy_stats(_lib.ptr(flat), flat.numel(), _lib.ptr(out4), _lib.ptr(self._stats_ws),
                                       _lib.stream_ptr(dev)), "qmcb_energy_stats")
        return D.global_stats(out4, None, None)

    # -- single point (solver_base.py:316-387) -------------------------------------------------
    def single_point(self, with_tqdm=True, batchsize=None, hdf5_group="single_point"):
        with torch.no_grad():
            pos = self.sampler(self.wf.pdf, with_tqdm=with_tqdm)
            if pos.device != self.device:
                pos = pos.to(self.device)
            out4 = None
            if batchsize is None:
                eloc, out4 = self.wf.local_energy_stats(pos)     # E_L and its sums in one pass
            else:
                eloc = torch.cat([self.wf.local_energy(pos[i: i + batchsize])
                                  for i in range(0, len(pos), batchsize)])
            # mean / unbiased variance / standard error from the four device sums (one all-reduce of
            # four doubles under torch.distributed); solver_base.py:371, wf_base.py:217-229
            mean, var, err, n, nbad = D.global_stats(out4, None, None) if out4 is not None else self._stats(eloc)
            dt = dict(dtype=torch.float64, device=eloc.device)
            e, s, er = torch.tensor(mean, **dt), torch.tensor(var, **dt), torch.tensor(err, **dt)
            res = SimpleNamespace(pos=pos, local_energy=eloc, energy=e, variance=s, error=er)
        self._dump("single_point", hdf5_group, res)
        return res

    def _dump(self, kind, group, obj):
        """HDF5 dump of a result + its ``type`` attribute (solver_base.py:381-385,466-470, solver.py:255-263);
        see utils/hdf5_write.py.  Rank 0 writes."""
        if not getattr(self, "write_hdf5", False) or self.rank != 0:
            return None
        from ..utils.hdf5_write import add_group_attr, dump_to_hdf5
        grp = dump_to_hdf5(obj, self.hdf5file, group)
        add_group_attr(self.hdf5file, grp, {"type": kind})
        return grp

    # -- optimisation (solver.py:186-431) --------------------------------------------------------
    def save_sampling_parameters(self):
        self.sampler._nstep_save = self.sampler.nstep
        self.sampler._ntherm_save = self.sampler.ntherm
        if self.resampling_options.mode == "update":
            self.sampler.ntherm = self.resampling_options.ntherm_update
            self.sampler.nstep = self.resampling_options.nstep_update

    def restore_sampling_parameters(self):
        self.sampler.nstep = self.sampler._nstep_save
        self.sampler.ntherm = self.sampler._ntherm_save

    def run(self, nepoch, batchsize=None, hdf5_group="wf_opt", chkpt_every=None, tqdm=False):
        self.prepare_optimization(batchsize, chkpt_every, tqdm)
        self.run_epochs(nepoch)
        self.restore_sampling_parameters()
        self.observable.models.last = dict(self.wf.state_dict())
        self._dump("opt", hdf5_group, self.observable)
        return self.observable

    def prepare_optimization(self, batchsize, chkpt_every, tqdm=False):
        pos = self.sampler(self.wf.pdf, with_tqdm=tqdm)
        pos = pos.detach().to(self.device)
        if batchsize is None:
            batchsize = len(pos)
        self.save_sampling_parameters()
        self.dataloader = _Loader(pos, batchsize)
        if D.is_distributed():
            # every rank must run the same number of batches: each batch carries one all-reduce of
            # (sum E_L, n) for the global mean, and ranks with different counts would hang
            nb = torch.tensor([float(ceil(len(pos) / batchsize)), -float(ceil(len(pos) / batchsize))],
                              dtype=torch.float64, device=self.device)
            D.allreduce_max_(nb)
            if nb[0] != -nb[1]:
                raise ValueError("torch.distributed: ranks would run different numbers of batches "
                                 "(%d..%d); choose nwalkers / batchsize so that every rank has the same "
                                 "count" % (int(-nb[1]), int(nb[0])))
        with torch.no_grad():
            for ibatch, data in enumerate(self.dataloader):
                self.store_observable(data, ibatch=ibatch)
        self.chkpt_every = chkpt_every

    def run_epochs(self, nepoch, with_tqdm=False, verbose=True):
        self._defer_observables = True
        if hasattr(self.sampler, "host_work"):
            self.sampler.host_work = self._flush_observables
        try:
            return self._run_epochs(nepoch)
        finally:
            self._defer_observables = False
            if hasattr(self.sampler, "host_work"):
                self.sampler.host_work = None
            self._flush_observables()

    def _run_epochs(self, nepoch):
        cumulative_loss = 0
        min_loss = 0
        for n in range(nepoch):
            tstart = time()
            cumulative_loss = 0
            self.opt.zero_grad()
            self.wf.zero_grad()
            for ibatch, data in enumerate(self.dataloader):
                lpos = data.to(self.device)
                # gradients of the batches accumulate locally in .grad; ONE all-reduce per epoch below
                loss, eloc = self.evaluate_gradient(lpos, allreduce=False)
                cumulative_loss += float(loss)
                if torch.isnan(eloc).any():
                    return cumulative_loss
                self.store_observable(lpos, local_energy=eloc, ibatch=ibatch)
            if self.grad_method == "auto":
                self._average_auto_gradients()
            else:
                D.allreduce_gradients(self._trainable())
            self.optimization_step(lpos)
            if n == 0 or cumulative_loss < min_loss:
                min_loss = cumulative_loss
                self.observable.models.best = {k: v.clone() for k, v in self.wf.state_dict().items()}
            if self.chkpt_every is not None and n > 0 and n % self.chkpt_every == 0:
                self.save_checkpoint(n, cumulative_loss)
            self.dataloader.dataset = self.resample(n, self.dataloader.dataset)
            if self.scheduler is not None:
                self.scheduler.step()
            self.epoch_time = time() - tstart
        return cumulative_loss

    def evaluate_grad_auto(self, lpos, allreduce=True):
        """solver.py:352-370: loss.backward() through the local energies.  The backward of E_L is one call of
        qmcb_local_energy_backward (adjoint of the Jacobi kinetic energy, csrc/eloc_vjp.cu).  Under
        torch.distributed the loss of a rank is the loss of its shard; the gradients are averaged."""
        bh = self.wf._jeen.jastrow_kernel if getattr(self.wf, "_jeen", None) is not None else None
        if bh is not None and any(p.requires_grad for p in bh.parameters()):
            raise NotImplementedError(
                "grad='auto' with trainable three-body (Boys-Handy) weights: the reference's own graph drops the "
                "dependence of the Jastrow Laplacian on them (jastrow_factor_electron_electron_nuclei.py:411-431, "
                "create_graph=False), so its gradient is not the derivative of the loss; freeze ['jastrow'] or "
                "use grad='manual'")
        loss, eloc = self.loss(lpos)
        loss.backward()
        if allreduce:
            self._average_auto_gradients()
        return loss.detach(), eloc.detach()

    def _average_auto_gradients(self):
        """grad="auto": every rank differentiated the loss of its own shard -> mean over the ranks."""
        if not D.is_distributed():
            return
        D.allreduce_gradients(self._trainable())
        for p in self._trainable():
            if p.grad is not None:
                p.grad /= D.world()[1]

    def compute_forces(self, lpos, batch_size=None, clip=None):
        """F = -< grad_A E_L + (E_L - <E_L>) grad_A log psi^2 >  as returned (without the sign) by
        solver.py:433-519: both gradients w.r.t. ``wf.ao.atom_coords`` come from qmcb_local_energy_backward
        (the reference back-propagates through the local energy and through log pdf)."""
        wf = self.wf
        original_requires_grad = wf.ao.atom_coords.requires_grad
        original_flag = wf.atom_coords_grad
        wf.ao.atom_coords.requires_grad = True
        wf.atom_coords_grad = True
        try:
            lpos = lpos.to(self.device)
            if batch_size is None:
                batch_size = lpos.shape[0]
            nbatch = lpos.shape[0] // batch_size
            forces = torch.zeros_like(wf.ao.atom_coords).requires_grad_(False)
            for ibatch in range(nbatch):
                batch = lpos[ibatch * batch_size: (ibatch + 1) * batch_size].detach()
                with torch.enable_grad():
                    local_energy = wf.local_energy(batch)
                    if clip is not None:
                        median = torch.median(local_energy)
                        std = torch.std(local_energy)
                        clip_mask = (torch.abs((local_energy - median) / std) < clip).to(local_energy.dtype)
                    else:
                        clip_mask = torch.ones_like(local_energy)
                    grad_eloc = torch.autograd.grad(local_energy, wf.ao.atom_coords, grad_outputs=clip_mask)[0]
                    proba = torch.log(wf.pdf(batch))
                    grad_outputs = ((local_energy - local_energy.mean()) * clip_mask).detach().squeeze()
                    grad_proba = torch.autograd.grad(proba, wf.ao.atom_coords,
                                                     grad_outputs=grad_outputs.reshape(proba.shape))[0]
                forces += 1.0 / batch_size * (grad_eloc + grad_proba)
        finally:
            wf.ao.atom_coords.requires_grad = original_requires_grad
            wf.atom_coords_grad = original_flag
        return forces

    def evaluate_grad_manual(self, lpos, allreduce=True):
        """dE/dk = < (dpsi/dk)/psi (E_L - <E_L>) > * 2   (solver.py:372-431); the mean and the
        normalisation are GLOBAL over all ranks.  ``allreduce=True`` (a direct call) sums the
        accumulated ``.grad`` over the ranks before returning - call it once per zero_grad();
        ``run_epochs`` passes False and all-reduces once per epoch after the batch loop."""
        if self.loss.method not in ["energy", "weighted-energy"]:
            raise ValueError("Manual gradient only for energy minimization")
        # ONE E_L launch yields E_L and psi; psi enters the autograd graph without a second launch
        eloc, psi = self.wf.local_energy_and_psi(lpos)
        if D.is_distributed():
            buf = torch.stack([eloc.sum(), torch.tensor(float(len(psi)), dtype=torch.float64, device=eloc.device)])
            D.allreduce_sum_(buf)
            ntot = buf[1]              # stays on the device: no host read-back between the two collectives
            mean = buf[0] / ntot
        else:
            ntot = float(len(psi))
            mean = torch.mean(eloc)
        weight = eloc.clone()
        weight -= mean
        weight /= psi.detach().clone()
        weight *= 2.0 / ntot
        if self.loss.clip:          # (no mask, and no host read-back of it, when clipping is off)
            weight = weight * self.loss.get_clipping_mask(eloc)
        psi.backward(weight)
        if allreduce:
            D.allreduce_gradients(self._trainable())
        return mean, eloc

    def _trainable(self):
        ps = [p for p in self.wf.parameters() if p.requires_grad]
        if self.wf.ao.bas_coeffs.requires_grad:
            ps.append(self.wf.ao.bas_coeffs)
        return ps

    def optimization_step(self, lpos):
        if self.opt.lpos_needed:
            self.opt.step(lpos)
        else:
            self.opt.step()

    def resample(self, n, pos):
        """solver_base.py:273-314 - walkers stay on the device between epochs."""
        if self.resampling_options.mode != "never":
            if n % self.resampling_options.resample_every == 0:
                if self.resampling_options.mode == "update":
                    pos = pos.clone().detach()[: self.sampler.walkers.nwalkers].to(self.device)
                else:
                    pos = None
                inc = self.resampling_options.increment
                if inc["every"] is not None and n % inc["every"] == 0:
                    self.sampler.nstep += inc["factor"] * self.sampler.ndecor
                pos = self.sampler(self.wf.pdf, pos=pos, with_tqdm=False).detach().to(self.device)
                self.dataloader.dataset = pos
            if self.loss.use_weight:
                self.loss.weight["psi0"] = None
        return pos

    def sampling_traj(self, pos=None, with_tqdm=True, hdf5_group="sampling_trajectory"):
        """solver_base.py:435-471."""
        if pos is None:
            pos = self.sampler(self.wf.pdf, with_tqdm=with_tqdm)
        ndim = pos.shape[-1]
        p = pos.view(-1, self.sampler.walkers.nwalkers, ndim)
        el = []
        with torch.no_grad():
            for ip in p:
                el.append(self.wf.local_energy(ip.to(self.device)).cpu().numpy())
        el = np.array(el).squeeze(-1)
        obs = SimpleNamespace(local_energy=el, pos=pos)
        self._dump("sampling_traj", hdf5_group, obs)
        return obs

    # -- small host-side helpers of the reference Solver (solver_base.py:250-271,473-517) --------------------
    def print_observable(self, cumulative_loss, verbose=False):
        """solver_base.py:250-271 (plain print instead of the twiggy logger)."""
        self._flush_observables()
        for k in self.observable.__dict__.keys():
            if k == "local_energy" and self.observable.local_energy:
                eloc = self.observable.local_energy[-1]
                e, v = np.mean(eloc), np.var(eloc)
                print("  energy   : %f +/- %f" % (e, np.sqrt(v / len(eloc))))
                print("  variance : %f" % np.sqrt(v))
            elif verbose and k not in ("qmctorch_version", "models") and getattr(self.observable, k):
                print(k + " : ", getattr(self.observable, k)[-1])
                print("loss %f" % cumulative_loss)

    def print_parameters(self, grad=False):
        """solver_base.py:473-484."""
        for p in self.wf.parameters():
            if p.requires_grad:
                print(p.grad if grad else p)

    def save_traj(self, fname, obs):
        """xyz trajectory of a geometry optimisation (solver_base.py:498-517; ``obs.geometry`` in bohr)."""
        nm2bohr = 1.88973
        with open(fname, "w") as f:
            for snap in obs.geometry:
                f.write("%d \n\n" % len(snap))
                for i, pos in enumerate(snap):
                    f.write("%s % 7.5f % 7.5f %7.5f\n" % (self.wf.atoms[i][0], pos[0] / nm2bohr, pos[1] / nm2bohr,
                                                        pos[2] / nm2bohr))
                f.write("\n")

    def log_data(self):
        pass

    def save_checkpoint(self, epoch, loss):
        """solver_base.py:389-405 (key spelling kept so checkpoints interchange)."""
        torch.save({"epoch": epoch, "model_state_dict": self.wf.state_dict(),
                    "optimzier_state_dict": self.opt.state_dict(), "loss": loss},
                   "checkpoint_epoch%d.pth" % epoch)

    def load_checkpoint(self, filename):
        data = torch.load(filename)
        self.wf.load_state_dict(data["model_state_dict"])
        self.opt.load_state_dict(data["optimzier_state_dict"])
        return data["epoch"], data["loss"]
