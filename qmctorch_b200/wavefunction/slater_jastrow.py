"""SlaterJastrow - drop-in for qmctorch.wavefunction.SlaterJastrow
(qmctorch/wavefunction/slater_jastrow.py:30-482) on top of libqmcb.so.

Same constructor, attributes, parameter names and method surface.  ``forward``,
``local_energy``, ``kinetic_energy``, ``gradients_jacobi`` and ``pdf`` each run one fused
sm_100a kernel per call; ``forward`` is differentiable w.r.t. the wave-function
parameters (``psi.backward(weight)`` as used by ``Solver.evaluate_grad_manual``) through
``qmcb_psi_backward``; ``local_energy`` is differentiable w.r.t. the parameters and the atom
coordinates (``grad="auto"``, ``Solver.compute_forces``) through ``qmcb_local_energy_backward``.
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib
from ._plan import PlanHandle, as_walkers
from .jastrows import (CombineJastrow, JastrowFactorElectronElectron, JastrowFactorElectronNuclei,
                       JastrowFactorElectronElectronNuclei)
from .jastrows.elec_elec import PadeJastrowKernel
from .orbitals import AtomicOrbitals, MolecularOrbitals
from .pooling import OrbitalConfigurations, SlaterPooling
from .wf_base import WaveFunction


class _PsiFunction(torch.autograd.Function):
    """psi(pos; theta) with analytic parameter gradients."""

    @staticmethod
    def forward(ctx, wf, x, bas_exp, bas_coeffs, mo_modifier, ci, jee_w, jen_w, een_num, een_denom, een_fc,
                psi_value=None, atom_coords=None):
        ctx.wf = wf
        ctx.save_for_backward(x)
        # psi_value: psi of exactly these walkers, already produced by the E_L launch (out1)
        return wf._psi(x) if psi_value is None else psi_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        wf = ctx.wf
        (x,) = ctx.saved_tensors
        need = ctx.needs_input_grad
        # parameter gradients only when some parameter asks for them (position-only autograd,
        # e.g. drift / Hamiltonian samplers, just needs qmcb_grad_psi)
        names = ("bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w", "een", "een", "een")
        want = {n for n, flag in zip(names, need[2:]) if flag}
        g = wf._psi_backward(x, grad_out.reshape(-1).contiguous(), want) if want else {}
        gx = None
        if need[1]:
            gx = wf._grad_psi(x, pdf=False) * grad_out.reshape(-1, 1)
        g = {k: g.get(k) for k in ("bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w", "een_num",
                                   "een_denom", "een_fc")}
        g_atom = None
        if need[12]:
            # d psi / d atom_coords (Solver.compute_forces): the adjoint kernel of the local energy with a psi weight
            g_atom = wf._eloc_backward(x, None, grad_out.reshape(-1).contiguous(), {"atom_coords"})["atom_coords"]
        return (None, gx,
                g["bas_exp"] if need[2] else None,
                # uncontracted bases: the reference never multiplies bas_coeffs into psi
                # (atomic_orbitals.py:236-249), so its .grad stays None
                g["bas_coeffs"] if (need[3] and wf.ao.contract) else None,
                g["mo_modifier"] if need[4] else None,
                g["ci"] if need[5] else None,
                g["jee_w"] if need[6] else None,
                g["jen_w"] if need[7] else None,
                g["een_num"] if need[8] else None,
                g["een_denom"] if need[9] else None,
                g["een_fc"] if need[10] else None,
                None, g_atom)


class _ElocFunction(torch.autograd.Function):
    """E_L(pos; theta) differentiable w.r.t. the parameters and the atom coordinates: the backward of
    WaveFunction.local_energy that Solver.evaluate_grad_auto (solver/solver.py:352-370) and
    Solver.compute_forces (:433-519) run through autograd, here one call of qmcb_local_energy_backward."""

    NAMES = ("atom_coords", "bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w", "jen_w")

    @staticmethod
    def forward(ctx, wf, x, atom_coords, bas_exp, bas_coeffs, mo_modifier, ci, jee_w, jen_w):
        ctx.wf = wf
        ctx.save_for_backward(x)
        return wf._eloc(x)[0]

    @staticmethod
    def backward(ctx, grad_out):
        wf = ctx.wf
        (x,) = ctx.saved_tensors
        need = ctx.needs_input_grad
        if need[1]:
            raise NotImplementedError("the local energy is not differentiable w.r.t. the walker positions here")
        want = {n for n, flag in zip(_ElocFunction.NAMES, need[2:]) if flag}
        if not wf.ao.contract:
            want.discard("bas_coeffs")      # as for psi: an uncontracted basis never multiplies bas_coeffs in
        g = wf._eloc_backward(x, grad_out.reshape(-1).contiguous(), None, want) if want else {}
        return (None, None) + tuple(g.get(n) for n in _ElocFunction.NAMES)


class SlaterJastrow(WaveFunction):
    def __init__(self, mol, jastrow="default", backflow=None, configs="ground_state", kinetic="jacobi",
                 cuda=False, include_all_mo=True, mix_mo=False, orthogonalize_mo=False):
        super().__init__(mol.nelec, 3, kinetic, cuda)
        if self.cuda and not torch.cuda.is_available():
            raise ValueError("Cuda not available, use cuda=False")
        if not include_all_mo and isinstance(configs, str) and configs.startswith("cas("):
            raise ValueError("CAS calculation only possible with include_all_mo=True")
        if backflow is not None:
            raise NotImplementedError(
                "backflow orbitals (slater_jastrow.py:484-580) are outside the fused hot path")
        if kinetic != "jacobi":
            raise NotImplementedError(
                "kinetic='auto' (autograd Hessian, wf_base.py:142-182) is not provided; the CUDA "
                "path implements the Jacobi formula")
        self.mol = mol
        self.atoms = mol.atoms
        self.natom = mol.natom
        # configurations (slater_jastrow.py:158-169)
        self.orb_confs = OrbitalConfigurations(mol)
        self.configs_method = configs if isinstance(configs, str) else "explicit"
        self.configs = self.orb_confs.get_configs(configs)
        self.nci = len(self.configs[0])
        self.highest_occ_mo = int(max(self.configs[0].max(), self.configs[1].max())) + 1
        # operators
        self.use_backflow = False
        self.ao = AtomicOrbitals(mol, cuda)
        self.include_all_mo = include_all_mo
        self.nmo_opt = mol.basis.nmo if include_all_mo else self.highest_occ_mo
        self.mo = MolecularOrbitals(mol, include_all_mo, self.highest_occ_mo, mix_mo, orthogonalize_mo, cuda)
        self.pool = SlaterPooling(self.configs_method, self.configs, mol, cuda)
        self.fc = nn.Linear(self.nci, 1, bias=False).to(torch.float64)
        self.fc.weight.data.fill_(0.0)
        self.fc.weight.data[0][0] = 1.0
        if self.cuda:
            self.fc = self.fc.to(self.device)
        self._init_jastrow(jastrow)
        self.kinetic_method = kinetic
        self.gradients = self.gradients_jacobi
        self.kinetic_energy = self.kinetic_energy_jacobi
        # shared device tables
        self._handle = PlanHandle(self.ao, self.mo, self.configs, self.fc, self._jee, self._jen,
                                  nup=mol.nup, ndown=mol.ndown, jastrow_een=self._jeen)
        for m in (self.ao, self.mo, self.pool, self._jee, self._jen, self._jeen,
                  self.jastrow if isinstance(self.jastrow, CombineJastrow) else None):
            if m is not None:
                m._handle = self._handle
        self._ws = {}

    # -- construction helpers -------------------------------------------------------------
    def _init_jastrow(self, jastrow):
        """slater_jastrow.py:193-222."""
        # plain attributes (not registered sub-modules: the parameters already live under
        # ``jastrow.`` and state_dict names must match the reference)
        self.__dict__["_jee"] = None
        self.__dict__["_jen"] = None
        self.__dict__["_jeen"] = None
        if jastrow is None:
            self.jastrow = None
            self.use_jastrow = False
            return
        self.use_jastrow = True
        if isinstance(jastrow, str) and jastrow == "default":
            self.jastrow = JastrowFactorElectronElectron(self.mol, PadeJastrowKernel, cuda=self.cuda)
        elif isinstance(jastrow, list):
            self.jastrow = CombineJastrow(jastrow)
        elif isinstance(jastrow, nn.Module):
            self.jastrow = jastrow
        else:
            raise TypeError("Jastrow factor not supported.")
        if isinstance(self.jastrow, JastrowFactorElectronElectron):
            self.__dict__["_jee"] = self.jastrow
        elif isinstance(self.jastrow, JastrowFactorElectronNuclei):
            self.__dict__["_jen"] = self.jastrow
        elif isinstance(self.jastrow, JastrowFactorElectronElectronNuclei):
            self.__dict__["_jeen"] = self.jastrow
        elif isinstance(self.jastrow, CombineJastrow):
            self.__dict__["_jee"], self.__dict__["_jen"] = self.jastrow.ee, self.jastrow.en
            self.__dict__["_jeen"] = self.jastrow.een
        else:
            raise NotImplementedError(
                "only the Pade e-e / e-n and Boys-Handy e-e-n Jastrow factors (and their product) are "
                "fused into the CUDA path")
        self.jastrow_type = self.jastrow.__repr__()
        if self.cuda:
            self.jastrow = self.jastrow.to(self.device)

    def set_combined_jastrow(self, jastrow):
        raise NotImplementedError("rebuild the wave function with jastrow=[...] instead")

    # -- raw kernel calls -----------------------------------------------------------------
    def _dev(self):
        return self.ao.atom_coords.device

    def _x(self, pos):
        return as_walkers(pos, self.ndim_tot, self._dev())

    def _psi(self, x):
        W = x.shape[0]
        out = torch.empty(W, 1, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().qmcb_psi(self._handle.plan(), _lib.ptr(x), W, _lib.ptr(out),
                                       _lib.stream_ptr(x.device)), "qmcb_psi")
        return out

    def _eloc(self, x, want_psi=False, want_ekin=False):
        W = x.shape[0]
        e = torch.empty(W, 1, dtype=torch.float64, device=x.device)
        p = torch.empty(W, 1, dtype=torch.float64, device=x.device) if want_psi else None
        k = torch.empty(W, 1, dtype=torch.float64, device=x.device) if want_ekin else None
        _lib.check(_lib.lib().qmcb_local_energy(self._handle.plan(), _lib.ptr(x), W, _lib.ptr(e), _lib.ptr(p),
                                                _lib.ptr(k), _lib.stream_ptr(x.device)), "qmcb_local_energy")
        return e, p, k

    def _grad_psi(self, x, pdf):
        W = x.shape[0]
        g = torch.empty(W, self.ndim_tot, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().qmcb_grad_psi(self._handle.plan(), _lib.ptr(x), W, int(pdf), _lib.ptr(g),
                                            _lib.stream_ptr(x.device)), "qmcb_grad_psi")
        return g

    def _psi_backward(self, x, weight, want=None):
        """want: names of the gradients to form (None = all).  Outputs that are not wanted are passed
        as NULL, which lets qmcb_psi_backward skip the basis-parameter contractions (4 x cheaper when
        only Jastrow / MO / CI gradients are needed, BASELINE config 3)."""
        L = _lib.lib()
        plan = self._handle.plan()
        dev = x.device
        W = x.shape[0]
        nao, nmo = self.mo.mo_scf.shape
        nbas = len(self.ao.index_ctr_np)         # plan primitives (one per cartesian monomial)
        g_mo = torch.empty(nao, nmo, dtype=torch.float64, device=dev)
        g_ci = torch.empty(1, self.nci, dtype=torch.float64, device=dev)
        g_exp = torch.empty(nbas, dtype=torch.float64, device=dev)
        g_cf = torch.empty(nbas, dtype=torch.float64, device=dev)
        g_jee = torch.empty(1, dtype=torch.float64, device=dev)
        g_jen = torch.empty(1, dtype=torch.float64, device=dev)
        nt = self._jeen.jastrow_kernel.nterm if self._jeen is not None else 0
        g_een = torch.zeros(max(5 * nt, 1), dtype=torch.float64, device=dev)
        nbytes = L.qmcb_backward_workspace_bytes(plan, W)
        ws = self._ws.get("bwd")
        if ws is None or ws.numel() < nbytes or ws.device != dev:
            ws = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)
            self._ws["bwd"] = ws
        def out(name, t):
            return _lib.ptr(t) if (want is None or name in want) else None
        _lib.check(L.qmcb_psi_backward(plan, _lib.ptr(x), _lib.ptr(weight), W, out("mo_modifier", g_mo),
                                       out("ci", g_ci), out("bas_exp", g_exp), out("bas_coeffs", g_cf),
                                       out("jee_w", g_jee), out("jen_w", g_jen),
                                       out("een", g_een) if nt else None, _lib.ptr(ws), _lib.stream_ptr(dev)),
                   "qmcb_psi_backward")
        g_exp, g_cf = self._fold_primitives(g_exp), self._fold_primitives(g_cf)
        return {"mo_modifier": g_mo * self.mo.mo_scf, "ci": g_ci, "bas_exp": g_exp, "bas_coeffs": g_cf,
                "jee_w": g_jee, "jen_w": g_jen,
                "een_num": g_een[: 2 * nt].view(1, 2, nt) if nt else None,
                "een_denom": g_een[2 * nt: 4 * nt].view(1, 2, nt) if nt else None,
                "een_fc": g_een[4 * nt: 5 * nt].view(1, nt) if nt else None}

    def _fold_primitives(self, g):
        """Plan primitives -> the reference's flat primitives: the cartesian monomials of a spherical
        harmonic share one (exponent, coefficient), so their derivatives add up."""
        ix = getattr(self.ao, "expand_index", None)
        if ix is None:
            return g
        ix = torch.as_tensor(ix, dtype=torch.long, device=g.device)
        return torch.zeros(self.ao.nbas, dtype=g.dtype, device=g.device).index_add_(0, ix, g)

    def _eloc_backward(self, x, w_eloc, w_psi, want):
        """sum_w w_eloc d E_L / d theta + w_psi d psi / d theta for the names in ``want`` (a subset of
        _ElocFunction.NAMES) through qmcb_local_energy_backward; returns {name: tensor shaped like the leaf}."""
        L = _lib.lib()
        plan = self._handle.plan()
        dev = x.device
        W = x.shape[0]
        nao, nmo = self.mo.mo_scf.shape
        nbas = len(self.ao.index_ctr_np)
        new = lambda *shape: torch.empty(*shape, dtype=torch.float64, device=dev)
        bufs = {"mo_modifier": new(nao, nmo), "ci": new(1, self.nci), "bas_exp": new(nbas), "bas_coeffs": new(nbas),
                "jee_w": new(1), "jen_w": new(1), "atom_coords": new(self.natom, 3)}
        if self._jee is None:
            want = set(want) - {"jee_w"}
        if self._jen is None:
            want = set(want) - {"jen_w"}
        nbytes = L.qmcb_local_energy_backward_workspace_bytes(plan, W)
        ws = self._ws.get("vjp")
        if ws is None or ws.numel() < nbytes or ws.device != dev:
            ws = self._ws["vjp"] = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)
        out = lambda name: _lib.ptr(bufs[name]) if name in want else None
        _lib.check(L.qmcb_local_energy_backward(
            plan, _lib.ptr(x), _lib.ptr(w_eloc), _lib.ptr(w_psi), W, out("mo_modifier"), out("ci"), out("bas_exp"),
            out("bas_coeffs"), out("jee_w"), out("jen_w"), out("atom_coords"), _lib.ptr(ws), _lib.stream_ptr(dev)),
            "qmcb_local_energy_backward")
        g = {n: bufs[n] for n in want}
        if "mo_modifier" in g:
            g["mo_modifier"] = g["mo_modifier"] * self.mo.mo_scf
        for n in ("bas_exp", "bas_coeffs"):
            if n in g:
                g[n] = self._fold_primitives(g[n])
        return g

    # atom coordinates join the autograd graphs of psi and E_L only on request (Solver.compute_forces switches
    # this on): ao.atom_coords requires grad by default like the reference's, and every psi.backward() of an
    # optimisation would otherwise pay for a derivative nobody reads
    atom_coords_grad = False

    # -- public API (reference signatures) ---------------------------------------------------
    def forward(self, x, ao=None, _psi_value=None):
        """psi(R) [W,1]  (slater_jastrow.py:243-286)."""
        if ao is not None:
            raise NotImplementedError("forward(x, ao=...) only serves the one-electron update sampler path")
        xd = self._x(x)
        jw = self._jee.jastrow_kernel.weight if self._jee is not None else None
        nw = self._jen.jastrow_kernel.weight if self._jen is not None else None
        bh = self._jeen.jastrow_kernel if self._jeen is not None else None
        leaves = [self.ao.bas_exp, self.ao.bas_coeffs, self.mo.mo_modifier, self.fc.weight, jw, nw,
                  bh.weight_num if bh is not None else None, bh.weight_denom if bh is not None else None,
                  bh.fc.weight if bh is not None else None]
        atom = self.ao.atom_coords if (self.atom_coords_grad and self.ao.atom_coords.requires_grad) else None
        track = torch.is_grad_enabled() and (
            any(t is not None and t.requires_grad for t in leaves) or x.requires_grad or atom is not None)
        if not track:
            return self._psi(xd) if _psi_value is None else _psi_value
        if x.requires_grad:
            xd = x if (x.device == xd.device and x.dtype == torch.float64 and x.is_contiguous()) else \
                x.to(device=xd.device, dtype=torch.float64).contiguous()
        return _PsiFunction.apply(self, xd, *leaves, _psi_value, atom)

    def local_energy_and_psi(self, pos):
        """(E_L [W,1] without graph, psi [W,1] in the autograd graph of the parameters) from ONE
        qmcb_local_energy launch: the kernel returns psi alongside E_L (out1), and the autograd node
        of psi is built around that value instead of a second forward launch.  This is the pair
        Solver.evaluate_grad_manual needs (solver.py:410-414)."""
        x = self._x(pos)
        with torch.no_grad():
            eloc, psi, _ = self._eloc(x, want_psi=True)
        return eloc, self.forward(pos if pos.requires_grad else x, _psi_value=psi)

    def ao2mo(self, ao):
        return self.mo(ao)

    def pos2mo(self, x, derivative=0, sum_grad=True):
        """slater_jastrow.py:292-310."""
        return self.ao2mo(self.ao(x, derivative=derivative, sum_grad=sum_grad))

    def local_energy(self, pos):
        """E_L [W,1]  (wf_base.py:184-215 + slater_jastrow.py:312-344).  A host tensor is streamed
        to the device in chunks so that the H2D copy of chunk k+1 overlaps the kernel on chunk k."""
        streamed = pos.device.type == "cpu" and pos.shape[0] >= self.host_chunk_min
        if torch.is_grad_enabled() and not streamed:
            # differentiable w.r.t. the parameters (grad="auto", forces): one autograd node around the E_L kernel
            # (a large HOST ensemble is streamed through the device in chunks and carries no graph: its device
            # copy does not outlive the call)
            jw = self._jee.jastrow_kernel.weight if self._jee is not None else None
            nw = self._jen.jastrow_kernel.weight if self._jen is not None else None
            atom = self.ao.atom_coords if self.atom_coords_grad else None
            leaves = [atom, self.ao.bas_exp, self.ao.bas_coeffs, self.mo.mo_modifier, self.fc.weight, jw, nw]
            if any(t is not None and t.requires_grad for t in leaves):
                return _ElocFunction.apply(self, self._x(pos), *leaves)
        if streamed:
            return self._eloc_from_host(pos)
        return self._eloc(self._x(pos))[0]

    def local_energy_stats(self, pos):
        """(E_L [W,1], device tensor [sum E_L, sum E_L^2, n finite, n non-finite]) from ONE call of
        qmcb_local_energy_stats: the energy step of Solver.single_point (solver_base.py:355-371)
        without a second pass over E_L.  The four sums are what ranks all-reduce."""
        x = self._x(pos)
        W = x.shape[0]
        L = _lib.lib()
        ws = self._ws.get("stats")
        if ws is None or ws.device != x.device:
            ws = self._ws["stats"] = torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8,
                                                 device=x.device)
        e = torch.empty(W, 1, dtype=torch.float64, device=x.device)
        out4 = torch.empty(4, dtype=torch.float64, device=x.device)
        _lib.check(L.qmcb_local_energy_stats(self._handle.plan(), _lib.ptr(x), W, _lib.ptr(e), None, None,
                                             _lib.ptr(out4), _lib.ptr(ws), _lib.stream_ptr(x.device)),
                   "qmcb_local_energy_stats")
        return e, out4

    host_chunk_min = 65536      # walkers; below this one copy + one launch is cheaper
    host_chunks = 8

    def _eloc_from_host(self, pos):
        dev = self._dev()
        if pos.dim() != 2 or pos.shape[1] != self.ndim_tot:
            raise ValueError("positions must have shape [nwalkers, %d], got %s" % (self.ndim_tot, tuple(pos.shape)))
        src = pos.detach()
        if src.dtype != torch.float64 or not src.is_contiguous():
            src = src.to(torch.float64).contiguous()
        W = src.shape[0]
        L = _lib.lib()
        plan = self._handle.plan()
        main = torch.cuda.current_stream(dev)
        side = self._ws.get("side")
        if side is None:
            side = self._ws["side"] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        xbuf = self._ws.get("xbuf")
        if xbuf is None or xbuf.shape[0] < W or xbuf.device != dev:
            xbuf = self._ws["xbuf"] = torch.empty(W, self.ndim_tot, dtype=torch.float64, device=dev)
        e = torch.empty(W, 1, dtype=torch.float64, device=dev)
        step = -(-W // self.host_chunks)
        for s_ in side:
            s_.wait_stream(main)
        for k, lo in enumerate(range(0, W, step)):
            hi = min(lo + step, W)
            st = side[k % 2]
            with torch.cuda.stream(st):
                xbuf[lo:hi].copy_(src[lo:hi], non_blocking=True)
                _lib.check(L.qmcb_local_energy(plan, _lib.ptr(xbuf[lo:hi]), hi - lo, _lib.ptr(e[lo:hi]), None, None,
                                               C.c_void_p(st.cuda_stream)), "qmcb_local_energy")
        for s_ in side:
            e.record_stream(s_)
            main.wait_stream(s_)
        return e

    def kinetic_energy_jacobi(self, x, **kwargs):
        """slater_jastrow.py:312-344."""
        return self._eloc(self._x(x), want_ekin=True)[2]

    def gradients_jacobi(self, x, sum_grad=False, pdf=False):
        """slater_jastrow.py:346-447 -> [W, 3*nelec]."""
        return self._grad_psi(self._x(x), pdf)

    def get_kinetic_operator(self, x, ao, dao, d2ao, mo):
        raise NotImplementedError("B_kin is formed inside the fused kernel and never materialised")

    def nuclear_potential(self, pos):
        """wf_base.py:72-95."""
        x = self._x(pos).view(-1, self.nelec, 1, 3)
        z = torch.as_tensor(self.ao.atomic_number, dtype=torch.float64, device=x.device)
        r = (x - self.ao.atom_coords.detach()[None, None]).norm(dim=-1)
        return (-z / r).sum((1, 2)).view(-1, 1)

    def electronic_potential(self, pos):
        """wf_base.py:49-70."""
        x = self._x(pos).view(-1, self.nelec, 3)
        iu = torch.triu_indices(self.nelec, self.nelec, 1, device=x.device)
        r = (x[:, iu[0]] - x[:, iu[1]]).norm(dim=-1)
        return (1.0 / r).sum(1).view(-1, 1)

    def nuclear_repulsion(self):
        """wf_base.py:97-116."""
        c = self.ao.atom_coords.detach()
        vnn = 0.0
        for a in range(self.natom - 1):
            for b in range(a + 1, self.natom):
                vnn = vnn + self.ao.atomic_number[a] * self.ao.atomic_number[b] / (c[a] - c[b]).norm()
        return vnn

    def geometry(self, pos=None, convert_to_angs=False):
        """slater_jastrow.py:629-647."""
        d = []
        convert = 0.529177 if convert_to_angs else 1
        for iat in range(self.natom):
            xyz = (self.ao.atom_coords[iat, :].detach().cpu().numpy() * convert).tolist()
            d.append(xyz)
        return d

    def gto2sto(self, plot=False):
        """Fit every contracted Gaussian AO with ONE Slater function ``norm * exp(-alpha |x|)`` and
        return the wave function on that single-zeta ``sto_pure`` basis (slater_jastrow.py:649-733).

        Host-side set-up code like the reference's: the radial profiles ``sum_p c_p N_p exp(-a_p x^2)``
        are tabulated on ``torch.linspace(-5, 5, 501)`` (default dtype, as there) and fitted
        with ``scipy.optimize.curve_fit``; norms of the new basis are recomputed by ``AtomicOrbitals``
        (``bas_norm`` is ignored, atomic_orbitals.py:90).  The returned object evaluates on the CUDA
        path like any other.  (A deliberate restatement of slater_jastrow.py:649-733: grid, dtype and fit call must
        be the reference's for the fitted exponents to come out bit-identical.)"""
        from copy import deepcopy
        import numpy as np
        from scipy.optimize import curve_fit
        assert self.ao.radial_type.startswith("gto")
        assert self.ao.harmonics_type == "cart"
        if plot:
            raise NotImplementedError("plot=True needs matplotlib, which is not a dependency here")

        def sto(x, norm, alpha):
            return norm * np.exp(-alpha * np.abs(x))

        nao = self.mol.basis.nao
        new_mol = deepcopy(self.mol)
        basis = deepcopy(self.mol.basis)
        basis.radial_type = "sto_pure"
        basis.nshells = self.ao.nao_per_atom.detach().cpu().numpy()
        basis.index_ctr = np.arange(nao)
        basis.bas_coeffs = np.ones(nao)
        basis.bas_exp = np.zeros(nao)
        basis.bas_norm = np.zeros(nao)
        basis.bas_kr = np.zeros(nao)
        basis.bas_kx = np.zeros(nao)
        basis.bas_ky = np.zeros(nao)
        basis.bas_kz = np.zeros(nao)

        x = torch.linspace(-5, 5, 501)
        pos = x.reshape(-1, 1).repeat(1, self.ao.nbas)
        norm = self.ao.norm_cst.detach().cpu()
        gto = norm * torch.exp(-self.ao.bas_exp.detach().cpu() * pos ** 2)
        idx = self.ao.index_ctr.cpu().long()
        ao = torch.zeros(len(x), nao, dtype=gto.dtype)
        ao.index_add_(1, idx, gto * self.ao.bas_coeffs.detach().cpu())      # atomic_orbitals.py:654-669
        ao = ao.numpy()
        xdata = x.numpy()
        for iorb in range(nao):
            popt, _ = curve_fit(sto, xdata, ao[:, iorb])
            basis.bas_norm[iorb], basis.bas_exp[iorb] = popt[0], popt[1]
            sel = (idx == iorb).numpy()
            for name, k in (("bas_kx", self.ao.bas_kx_np), ("bas_ky", self.ao.bas_ky_np), ("bas_kz", self.ao.bas_kz_np)):
                ks = np.unique(k[sel])
                if len(ks) != 1:
                    raise ValueError("primitives of one AO with different powers")
                getattr(basis, name)[iorb] = ks[0]
        new_mol.basis = basis
        return self.__class__(new_mol, self.jastrow, backflow=None, configs=self.configs_method,
                              kinetic=self.kinetic_method, cuda=self.cuda, include_all_mo=self.include_all_mo)

    def log_data(self):
        pass
