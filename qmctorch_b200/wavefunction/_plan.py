"""Host-side plan handle: flattens the modules' parameters into a ``qmcb_system`` and keeps
the device tables of libqmcb.so in sync with them.

The handle is shared by a SlaterJastrow and its sub-modules (``ao``, ``mo``, ``pool``,
``jastrow``) so that operator-level calls and the fused path see the same tables.
"""
import ctypes as C
import weakref

import numpy as np
import torch

from .. import _lib


def _cpu(t):
    return t.detach().to("cpu", torch.float64).contiguous().numpy()


class PlanHandle:
    def __init__(self, ao, mo=None, configs=None, fc=None, jastrow_ee=None, jastrow_en=None,
                 nup=None, ndown=None, device=None, jastrow_een=None):
        self.ao = ao
        self.mo = mo
        self.configs = configs
        self.fc = fc
        self.jee = jastrow_ee
        self.jen = jastrow_en
        self.jeen = jastrow_een
        self.nup = ao.nup if nup is None else nup
        self.ndown = ao.ndown if ndown is None else ndown
        self.device = device
        self._plan = C.c_void_p()
        self._sig = None
        self._content = None
        self._arrays = None
        self._finalizer = None
        # "content": every plan() call compares the parameter VALUES with the ones the device tables
        # were built from (a few small device ops + one 1-byte read-back, ~50 us), so in-place edits
        # through ``.data`` - which change neither data_ptr nor Tensor._version - are seen.
        # "version": (data_ptr, _version) only - no read-back, for callers that never edit ``.data``
        # (or call invalidate() after doing so) and want fully asynchronous launches.
        self.param_check = "content"

    # -- parameters that feed the device tables
    def _tracked(self):
        ts = [self.ao.atom_coords, self.ao.bas_exp, self.ao.bas_coeffs]
        if self.mo is not None:
            ts += [self.mo.mo_modifier, self.mo.mo_scf]
        if self.fc is not None:
            ts.append(self.fc.weight)
        if self.jee is not None:
            ts.append(self.jee.jastrow_kernel.weight)
        if self.jen is not None:
            ts.append(self.jen.jastrow_kernel.weight)
        if self.jeen is not None:
            k = self.jeen.jastrow_kernel
            ts += [k.weight_num, k.weight_denom, k.fc.weight]
        return ts

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in self._tracked())

    def _flat(self):
        return torch.cat([t.detach().reshape(-1).to(torch.float64) for t in self._tracked()])

    def invalidate(self):
        """Forces the next plan() call to rebuild the device tables from the current parameters."""
        self._sig = None
        self._content = None

    def _host(self, flat=None):
        """Every tracked tensor on the host from ONE device-to-host copy (a plan update after opt.step()
        otherwise pays one synchronising copy per tensor): list of numpy arrays in _tracked() order."""
        ts = self._tracked()
        if flat is None:
            flat = self._flat()
        h = flat.cpu().numpy()
        out, off = [], 0
        for t in ts:
            n = t.numel()
            out.append(h[off:off + n].reshape(tuple(t.shape)))
            off += n
        return out

    def _system(self, flat=None):
        ao = self.ao
        nelec = self.nup + self.ndown
        nao = ao.norb
        host = iter(self._host(flat))
        h_atom, h_exp, h_coef = next(host), next(host), next(host)
        if self.mo is not None:
            h_mod, h_scf = next(host), next(host)
            w = h_scf * h_mod
        else:
            w = np.eye(nao)
        h_ci = next(host) if self.fc is not None else None
        h_jee = next(host) if self.jee is not None else None
        h_jen = next(host) if self.jen is not None else None
        h_een = [next(host), next(host), next(host)] if self.jeen is not None else None
        if getattr(self, "_norm_host", None) is None:
            self._norm_host = _cpu(ao.norm_cst)            # frozen at construction (atomic_orbitals.py:90-94)
        h_norm = self._norm_host
        if getattr(ao, "expand_index", None) is not None:
            # spherical harmonics: one plan primitive per (primitive, cartesian monomial); the monomial's
            # coefficient rides on the norm (an uncontracted basis keeps bas_coeffs = 1)
            ix = ao.expand_index
            h_exp, h_coef, h_norm = h_exp[ix], h_coef[ix], h_norm[ix] * ao.expand_scale
        nmo = w.shape[1]
        if self.configs is not None:
            cu = np.asarray(self.configs[0].cpu().numpy(), dtype=np.int32).reshape(-1, max(self.nup, 0))
            cd = np.asarray(self.configs[1].cpu().numpy(), dtype=np.int32).reshape(-1, max(self.ndown, 0))
        else:
            cu = np.arange(self.nup, dtype=np.int32)[None]
            cd = np.arange(self.ndown, dtype=np.int32)[None]
        nconf = cu.shape[0]
        if self.fc is not None:
            ci = h_ci.reshape(-1)
        else:
            ci = np.zeros(nconf)
            ci[0] = 1.0
        # Gram-form dot product: ATen's small-matrix bmm path (unfused) is taken when
        # 3*Ne*Ne < 400, MKL (fma chain) otherwise - SURVEY.md section 7, hard part 1.
        gram_fma = 0 if 3 * nelec * nelec < 400 else 1
        return _lib.SystemArrays(
            nelec=nelec, nup=self.nup, ndown=self.ndown, natom=ao.natoms, nbas=len(ao.index_ctr_np), nao=nao,
            nmo=nmo, radial_type=_lib.RADIAL[ao.radial_type], contract=int(ao.contract),
            atom_coords=h_atom, atomic_number=np.asarray(ao.atomic_number, dtype=np.float64),
            bas_atom=ao.bas_atom_np, bas_exp=h_exp, bas_coeffs=h_coef,
            bas_norm=h_norm, bas_kx=ao.bas_kx_np, bas_ky=ao.bas_ky_np, bas_kz=ao.bas_kz_np,
            bas_kr=ao.bas_kr_np, index_ctr=ao.index_ctr_np, mo=w, nconf=nconf, cfg_up=cu, cfg_down=cd,
            ci=ci,
            use_jee=int(self.jee is not None),
            jee_w=float(h_jee.reshape(-1)[0]) if self.jee is not None else 0.0,
            use_jen=int(self.jen is not None),
            jen_w=float(h_jen.reshape(-1)[0]) if self.jen is not None else 0.0,
            gram_fma=gram_fma, **self._een_arrays(h_een))

    def _een_arrays(self, h_een):
        if self.jeen is None:
            return dict(een_nterm=0)
        k = self.jeen.jastrow_kernel
        return dict(een_nterm=k.nterm, een_num=h_een[0].reshape(2, k.nterm),
                    een_denom=h_een[1].reshape(2, k.nterm), een_fc=h_een[2].reshape(-1))

    def plan(self):
        """Returns the (up to date) ``qmcb_plan*``; rebuilds the tables if a parameter changed."""
        dev = self.ao.atom_coords.device
        if dev.type != "cuda":
            raise RuntimeError(
                "qmctorch_b200 runs on CUDA devices only (construct the wave function with "
                "cuda=True); there is no CPU path")
        sig = self._signature()
        flat = None
        if sig == self._sig:
            if self.param_check != "content":
                return self._plan
            flat = self._flat()
            if self._content is not None and self._content.shape == flat.shape and \
                    self._content.device == flat.device and torch.equal(self._content, flat):
                return self._plan
        L = _lib.lib()
        if flat is None:
            flat = self._flat()
        arrays = self._system(flat)
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if not self._plan:
            # make sure the primary context exists before the library's runtime touches it
            torch.cuda.current_stream(dev)
            _lib.check(L.qmcb_plan_create(C.byref(arrays.struct), index, C.byref(self._plan)),
                       "qmcb_plan_create")
            handle = self._plan.value
            self._finalizer = weakref.finalize(self, L.qmcb_plan_destroy, C.c_void_p(handle))
        else:
            torch.cuda.current_stream(dev).synchronize()
            _lib.check(L.qmcb_plan_update(self._plan, C.byref(arrays.struct)), "qmcb_plan_update")
        self._arrays = arrays
        self._sig = sig
        if self.param_check == "content":
            self._content = flat.clone()
        return self._plan

    def host_plan_info(self):
        """Host-only plan (no device) -> dict of grouping/tiling figures; used by CPU tests."""
        L = _lib.lib()
        arrays = self._system()
        p = C.c_void_p()
        _lib.check(L.qmcb_plan_create(C.byref(arrays.struct), -1, C.byref(p)), "qmcb_plan_create(host)")
        names = ["nshell", "nprim", "ncomp", "nmo_used", "nuniq_up", "nuniq_down", "tw_eloc",
                 "threads_eloc", "smem_eloc", "tw_psi"]
        out = {n: L.qmcb_plan_info(p, i) for i, n in enumerate(names)}
        L.qmcb_plan_destroy(p)
        return out

    def info(self, what):
        return _lib.lib().qmcb_plan_info(self.plan(), what)


def as_walkers(pos, ncols, device):
    """Contiguous FP64 [W, ncols] tensor on the wave function's device (one H2D copy if needed)."""
    if pos.dim() != 2 or pos.shape[1] != ncols:
        raise ValueError("positions must have shape [nwalkers, %d], got %s" % (ncols, tuple(pos.shape)))
    return pos.detach().to(device=device, dtype=torch.float64).contiguous()
