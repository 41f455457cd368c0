"""Benchmark of the hot path: local-energy evaluations per second (walkers x steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lih|h2|h2o|c4h6|lih-opt] [--walkers M]
                    [--configs all|none|h2,lih-opt,...]
    python bench.py --impl reference ...      # the reference's own CPU implementation on host cores

Headline line (BASELINE config 2): a "step" is ONE local-energy evaluation of every walker of this rank
(LiH 6-31G single determinant + Pade Jastrow, 1e6 walkers per GPU) followed by the energy statistics
(fused into the kernel; one 4-double all-reduce when N > 1).  Walkers are thermalised by the fused
Metropolis kernel beforehand and are resident in HBM when the timed region starts; four independent
ensembles (4 x 96 MB > 126 MB L2) are cycled so that no step re-reads an ensemble that is still in L2.

The same JSON line carries
  * ``configs``: one entry per BASELINE.json configuration (H2 single point, LiH, LiH optimisation step
    with its two all-reduces, H2O CAS + three-body Jastrow, C4H6 with a walker/basis sweep), each with
    its own throughput, kernel time, roofline fractions and CPU leg;
  * ``strong``: BASELINE config 2 as written (1e6 walkers in TOTAL, split over the N GPUs), stepped
    from a CUDA graph so that the 18 us kernels at N = 8 are not launch-bound;
  * ``e2e``: the public call on pinned HOST walkers (H2D + kernel + statistics + D2H per step) next to
    the raw H2D copy ceiling measured in the same run.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BASELINE_CONFIGS = [
    "H2 sto-3g (pyscf SCF) SlaterJastrow + PadeJastrow, Metropolis 1000 walkers, single-point VMC energy on CPU",
    "LiH 6-31G single-determinant SlaterJastrow, 1M walkers, VMC energy + Jacobi kinetic on 8xB200",
    "LiH wavefunction optimisation (Jastrow + MO coeffs), gradient allreduce over NVLink",
    "H2O cc-pVDZ CAS multi-determinant CI expansion with elec-elec-nuc Jastrow",
    "Butadiene C4H6 dzp single-point VMC, 4M walkers, large-basis AO/MO throughput sweep",
]

# name -> workload.  walkers = per GPU; cpu_walkers = bounded sample of the CPU leg; therm = Metropolis
# moves per ensemble before timing; nbuf = ensembles cycled (inputs > L2 between timed iterations)
WORKLOADS = {
    "h2": dict(baseline=0, label="H2 STO-3G", fixture="h2", configs="single(2,2)", jastrow="ee",
               walkers=1_000_000, step=0.5, cpu_walkers=50_000, therm=100, nbuf=4),
    "lih": dict(baseline=1, label="LiH 6-31G", fixture="lih", configs="ground_state", jastrow="ee",
                walkers=1_000_000, step=0.3, cpu_walkers=50_000, therm=100, nbuf=4),
    "lih-opt": dict(baseline=2, label="LiH 6-31G optimisation step", fixture="lih", configs="ground_state",
                    jastrow="ee", walkers=1_000_000, step=0.3, cpu_walkers=20_000, therm=100, nbuf=1, kind="opt"),
    "h2o": dict(baseline=3, label="H2O cc-pVDZ", fixture="h2o", configs="cas(4,4)", jastrow="ee+een",
                walkers=250_000, step=0.15, cpu_walkers=512, therm=60, nbuf=4),
    "c4h6": dict(baseline=4, label="C4H6 DZP", fixture="c4h6", configs="ground_state", jastrow="ee",
                 walkers=4_000_000, step=0.05, cpu_walkers=2000, therm=20, nbuf=1,
                 sweep=[("c4h6", 100_000), ("c4h6", 1_000_000), ("c4h6_dz", 1_000_000)]),
}

# sm__inst_executed_pipe_fp64 (% of peak) of the dominant kernel from the `ncu --set full` captures under
# profiles/ (static: a number measured under a profiler is evidence, never a bench value), and the DRAM
# traffic per launch of the same capture: (workload, walkers) -> (pipe %, bytes, file)
NCU = {}      # (workload, walkers) -> (FP64 pipe % under ncu, DRAM bytes per launch, file); profiles/ncu_summary.json
try:
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as _f:
        for _k, _v in json.load(_f).items():
            _name, _w = _k.rsplit("@", 1)
            NCU[(_name, int(_w))] = (_v.get("pipe_fp64_pct"), _v.get("dram_bytes"), _v.get("file"))
except Exception:
    pass


# ------------------------------------------------------------------------------------------------
# flop conventions: add/mul = 1, fma = 2, div/sqrt/exp = 1 each
# ------------------------------------------------------------------------------------------------
def algorithmic_flops(mol, wf, info, een_nterm=0):
    """Flops per local-energy evaluation of the REFERENCE formulation: five projected AO
    channels, occupied MO columns only, every ordered electron pair."""
    ne, nat = mol.nelec, mol.natom
    nprim, ncomp, nmu = info["nprim"], info["ncomp"], info["nmo_used"]
    npair = ne * (ne - 1) // 2
    ao = ne * (nat * 8 + nprim * 12 + ncomp * 12)          # r^2; exp + radial sums; harmonic products
    mo = ne * ncomp * nmu * 5 * 2                         # 5 channels contracted on the fly
    bkin = ne * nmu * 10
    jast = 2 * npair * 45 + ne * nat * 6                  # ordered pairs (each pair visited twice) + V_en
    jast += 2 * npair * nat * een_nterm * 40              # Boys-Handy: (ordered pair, atom, term)
    n = max(mol.nup, mol.ndown)
    nun = info["nuniq_up"] + info["nuniq_down"]
    slater = nun * (12 if n <= 2 else 4 * n ** 3)
    return ao + mo + bkin + jast + slater + 4 * wf.nci + 10


def executed_flops(mol, wf, info, een_nterm=0):
    """Same conventions, for the formulas the E_L kernels EXECUTE: two projected
    channels (folded kinetic channel) instead of five, lap R from two radial sums, every electron pair
    visited once."""
    ne, nat = mol.nelec, mol.natom
    nprim, ncomp, nmu, nshell = info["nprim"], info["ncomp"], info["nmo_used"], info["nshell"]
    npair = ne * (ne - 1) // 2
    ao = ne * (nat * 14 + nprim * 8 + nshell * 6 + ncomp * 4)
    mo = ne * ncomp * nmu * 2 * 2
    jast = npair * 60 + ne * nat * 2 + npair * nat * een_nterm * 50
    n = max(mol.nup, mol.ndown)
    nun = info["nuniq_up"] + info["nuniq_down"]
    slater = nun * (12 if n <= 2 else 4 * n ** 3)
    return ao + mo + ne * nmu + jast + slater + 4 * wf.nci + 10


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the reference's own implementation (unmodified QMCTorch from baseline/_ref or /root/reference,
# imported behind the stub modules of oracle/ref_shim.py) when it is present, else the oracle port
# ------------------------------------------------------------------------------------------------
def _reference_wf(spec):
    """(kind, mol, wf, local_energy(pos), psi(pos), sample(nw, nstep)) on the CPU."""
    import torch
    from qmctorch_b200.molecules import fixture_molecule
    torch.set_num_threads(os.cpu_count() or 1)
    mol = fixture_molecule(spec["fixture"])
    import ref_shim
    if ref_shim.available():
        ref_shim.load_reference()
        from qmctorch.wavefunction import SlaterJastrow
        from qmctorch.sampler import Metropolis
        j = "default"
        if spec["jastrow"] == "ee+een":
            from qmctorch.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
            from qmctorch.wavefunction.jastrows.elec_elec_nuclei import (JastrowFactor as JEEN,
                                                                         BoysHandyJastrowKernel as BH)
            j = [JEE(mol, PEE), JEEN(mol, BH)]
        wf = SlaterJastrow(mol, configs=spec["configs"], jastrow=j, include_all_mo=True)
        autograd = spec["jastrow"] == "ee+een"      # the reference differentiates the three-body term by autograd

        def eloc(pos):
            if autograd:
                return wf.local_energy(pos.detach().clone().requires_grad_()).detach()
            with torch.no_grad():
                return wf.local_energy(pos)

        def sample(nw, nstep):
            torch.manual_seed(0)
            s = Metropolis(nwalkers=nw, nstep=nstep, step_size=spec["step"], nelec=wf.nelec, ndim=3,
                           init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"})
            return s(wf.pdf, with_tqdm=False).detach()
        return "reference", mol, wf, eloc, sample
    import sj_oracle as orc
    from qmctorch_b200.wavefunction.pooling import OrbitalConfigurations
    cfg = OrbitalConfigurations(mol).get_configs(spec["configs"])
    P = orc.make_params(mol, cfg, jastrow_weight=1.0)
    if spec["jastrow"] == "ee+een":
        g = torch.Generator().manual_seed(5)
        P.een = dict(num=0.05 + 0.3 * torch.rand(1, 2, 5, generator=g, dtype=torch.float64),
                     denom=0.5 + torch.rand(1, 2, 5, generator=g, dtype=torch.float64),
                     fc=torch.rand(1, 5, generator=g, dtype=torch.float64) - 0.5)

    def eloc(pos):
        with torch.no_grad():
            return orc.local_energy(P, pos)

    def sample(nw, nstep):
        g = torch.Generator().manual_seed(0)
        mean = torch.as_tensor(mol.domain("normal")["mean"])
        sig = torch.as_tensor(mol.domain("normal")["sigma"]).diagonal().sqrt()
        pos = (mean + sig * torch.randn(nw, mol.nelec, 3, generator=g, dtype=torch.float64)).view(nw, -1)
        with torch.no_grad():
            fx = (orc.psi(P, pos) ** 2).reshape(-1)
            for _ in range(nstep):
                d = torch.randn(pos.shape, generator=g, dtype=torch.float64) * (orc.proposal_sigma(spec["step"]) ** 0.5)
                tau = torch.rand(nw, generator=g, dtype=torch.float64)
                pos, fx, _, _ = orc.metropolis_step(P, pos, fx, d, tau)
        return pos
    return "port", mol, P, eloc, sample


def cpu_leg(name, steps, warmup, budget_s=20.0):
    """Times the CPU implementation of workload `name` on a bounded sample.  Every step is timed on its
    own; the value is sample / median step time (robust against the +-40 % jitter of single runs)."""
    import torch
    spec = WORKLOADS[name]
    kind, mol, wf, eloc, sample = _reference_wf(spec)
    nw = spec["cpu_walkers"]
    pos = sample(nw, 10)
    cores = torch.get_num_threads()
    if spec.get("kind") == "opt":
        return _cpu_opt_leg(kind, spec, mol, wf, pos, steps, warmup, cores, budget_s)
    times, t_all = [], time.perf_counter()
    e = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        e = eloc(pos)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_all > budget_s and len(times) >= 3:
            break
    med = statistics.median(times)
    what = ("unmodified QMCTorch wf.local_energy" if kind == "reference" else "oracle.local_energy (port)")
    return {"value": nw / med, "unit": "evals/s", "cores": cores, "kind": kind,
            "sample": "%d walkers x %d steps of %s, torch CPU FP64, median step" % (nw, len(times), what),
            "ms_per_step": med * 1e3, "steps_timed": len(times), "energy": float(e.mean())}


def _cpu_opt_leg(kind, spec, mol, wf, pos, steps, warmup, cores, budget_s):
    """BASELINE config 3 on the CPU: one optimisation step = Solver.evaluate_grad_manual + opt.step."""
    import torch
    nw = pos.shape[0]
    if kind == "reference":
        from qmctorch.solver import Solver
        from qmctorch.sampler import Metropolis
        s = Metropolis(nwalkers=nw, nstep=10, step_size=spec["step"], nelec=wf.nelec, ndim=3,
                       init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"})
        opt = torch.optim.Adam(wf.parameters(), lr=1e-3)
        solver = Solver(wf=wf, sampler=s, optimizer=opt)
        solver.configure(track=["local_energy"], freeze=["ci", "ao"], loss="energy", grad="manual")

        def step():
            opt.zero_grad()
            solver.evaluate_grad_manual(pos)
            opt.step()
    else:
        import sj_oracle as orc

        def step():
            orc.param_grads(wf, pos, names=("jastrow_weight", "mo_modifier"))
    times, t_all = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_all > budget_s and len(times) >= 3:
            break
    med = statistics.median(times)
    return {"value": nw / med, "unit": "walker-gradient evals/s", "cores": cores, "kind": kind,
            "sample": "%d walkers x %d optimisation steps (E_L + psi forward/backward + Adam), torch CPU FP64, "
                      "median step" % (nw, len(times)), "ms_per_step": med * 1e3, "steps_timed": len(times)}


def config_dict(name, wpg, world):
    """The `config` object: identical in the b200 and the reference arm (the driver compares them)."""
    spec = WORKLOADS[name]
    what = ("optimisation step: E_L + psi + parameter-gradient backward + all-reduces + Adam + table update"
            if spec.get("kind") == "opt" else "VMC local energy (Jacobi kinetic) + energy statistics")
    return {"workload": "%s %s, %d walkers/GPU, %s" % (spec["label"], spec["configs"], wpg, what),
            "baseline_config": BASELINE_CONFIGS[spec["baseline"]], "walkers_per_gpu": wpg,
            "jastrow": {"ee": "pade e-e", "ee+een": "pade e-e x boys-handy e-e-n"}[spec["jastrow"]],
            "parallelism": "walker shards x%d" % world,
            "l2": "%d ensembles cycled (inputs > L2 between timed iterations)" % spec["nbuf"],
            "cpu_sample_walkers": spec["cpu_walkers"]}


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def build_wf(spec, fixture=None):
    import torch
    from qmctorch_b200.molecules import fixture_molecule
    from qmctorch_b200.wavefunction import SlaterJastrow
    mol = fixture_molecule(fixture or spec["fixture"])
    j = "default"
    if spec["jastrow"] == "ee+een":
        from qmctorch_b200.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
        from qmctorch_b200.wavefunction.jastrows.elec_elec_nuclei import (JastrowFactor as JEEN,
                                                                         BoysHandyJastrowKernel as BH)
        torch.manual_seed(3)
        j = [JEE(mol, PEE, cuda=True), JEEN(mol, BH, cuda=True)]
    wf = SlaterJastrow(mol, configs=spec["configs"], jastrow=j, cuda=True)
    if wf.nci > 1:
        with torch.no_grad():
            g = torch.Generator().manual_seed(11)
            wf.fc.weight.add_((0.05 * torch.rand(wf.fc.weight.shape, generator=g, dtype=torch.float64)).to(wf.fc.weight.device))
    return mol, wf


def plan_info(wf):
    names = ["nshell", "nprim", "ncomp", "nmo_used", "nuniq_up", "nuniq_down", "tw_eloc", "threads_eloc",
             "smem_eloc", "tw_psi"]
    return {n: wf._handle.info(i) for i, n in enumerate(names)}


def thermalised(C, mol, wf, spec, W, nbuf, therm, seed0):
    """nbuf independent ensembles: reference 'normal' start (drawn on the device), thermalised by the
    fused Metropolis kernel.  (QMCB_BENCH_THERM=<n> shortens the thermalisation for profiler runs.)"""
    import torch
    from qmctorch_b200.sampler import Metropolis
    therm = int(os.environ.get("QMCB_BENCH_THERM", therm))
    ens = []
    for b in range(nbuf):
        torch.manual_seed(1234 + 17 * C.rank + b)
        s = Metropolis(nwalkers=W, nstep=therm, step_size=spec["step"], nelec=wf.nelec, ndim=3,
                       init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True,
                       seed=seed0 + 1000 * C.rank + b, keep_on_device=True, init_rng="philox")
        ens.append(s(wf.pdf, with_tqdm=False).detach().contiguous())
    return ens


def time_eloc(C, wf, ens, W, steps, warmup, allreduce=True):
    """K local-energy + statistics steps over the cycled ensembles.  Returns (ms over the K steps = max
    over ranks, mean ms of the E_L call's own event pair, [sum, sum2, n, nbad], launches per step)."""
    import torch
    import torch.distributed as dist
    from qmctorch_b200 import _lib
    L = _lib.lib()
    dev = C.dev
    world = C.world if allreduce else 1
    eloc = torch.empty(W, 1, dtype=torch.float64, device=dev)
    out4 = torch.zeros(4, dtype=torch.float64, device=dev)
    # N > 1: the 4-double all-reduce of step i is issued asynchronously (its own buffer) and only waited
    # for when the buffer comes round again, so it overlaps the kernel of step i+1
    ring = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(8)]
    ws = torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8, device=dev)
    plan = wf._handle.plan()
    stream = torch.cuda.current_stream(dev)
    sp = _lib.stream_ptr(dev)
    nb = len(ens)

    def call(x, o4):
        _lib.check(L.qmcb_local_energy_stats(plan, _lib.ptr(x), W, _lib.ptr(eloc), None, None, _lib.ptr(o4),
                                             _lib.ptr(ws), sp), "local_energy_stats")
    for i in range(warmup):
        call(ens[i % nb], out4)
        if world > 1:
            dist.all_reduce(out4)
    C.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    C.barrier()
    ev0.record(stream)
    pending = []
    for i in range(steps):
        o4 = ring[i % len(ring)] if world > 1 else out4
        if world > 1 and len(pending) >= len(ring):
            pending.pop(0).wait()          # the buffer about to be reused must have been reduced
        kev[i][0].record(stream)
        call(ens[i % nb], o4)
        kev[i][1].record(stream)
        if world > 1:
            pending.append(dist.all_reduce(o4, async_op=True))
    for h in pending:
        h.wait()
    if world > 1:
        out4.copy_(ring[(steps - 1) % len(ring)])
    ev1.record(stream)
    C.barrier()
    elapsed_ms = C.max_over_ranks(ev0.elapsed_time(ev1))
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / steps
    spec_on = wf._handle.info(13) == 1
    two_stage = os.environ.get("QMCB_STATS_2STAGE", "0") not in ("", "0")
    launches = (2 if two_stage else 1) if spec_on else 3
    return elapsed_ms, kern_ms, out4.tolist(), launches, spec_on


def roofline_obj(C, name, mol, wf, info, W, kern_ms, spec_on, een_nterm=0):
    F = algorithmic_flops(mol, wf, info, een_nterm)
    Fx = executed_flops(mol, wf, info, een_nterm)
    ach = W * F / (kern_ms * 1e-3) / 1e12
    achx = W * Fx / (kern_ms * 1e-3) / 1e12
    bpe = 24 * mol.nelec + 8
    ncu = NCU.get((name, W)) or next((v for (k, w), v in NCU.items() if k == name), None)
    return {"bound": "fp64", "kernel": C.kernel_name(wf), "achieved": ach, "peak": C.peak_tf,
            "unit": "TFLOP/s", "frac": ach / C.peak_tf if C.peak_tf else None,
            "frac_executed": achx / C.peak_tf if C.peak_tf else None,
            "peak_source": "own DFMA probe (qmcb_fp64_probe) on this GPU in this run; MEASURED_PEAKS.json has "
                           "no FP64 entry",
            "flops_per_eval": F, "flops_convention": "SURVEY 8(d): reference formulation (five projected AO "
            "channels, occupied MO columns, ordered pairs); add/mul 1, fma 2, div/sqrt/exp 1",
            "flops_executed_per_eval": Fx, "kernel_ms": kern_ms,
            "pipe_fp64_ncu_pct": ncu[0] if ncu else None, "ncu_file": ncu[2] if ncu else None,
            "traffic": ncu[1] if ncu and (name, W) in NCU else None,
            "hbm": {"achieved_gbs": W * bpe / (kern_ms * 1e-3) / 1e9, "peak_gbs": C.hbm_peak,
                    "bytes_per_eval": bpe, "peak_source": C.hbm_src}}


def run_eloc_config(C, name, steps, warmup, with_cpu, W=None, fixture=None, sweep=True):
    """One E_L workload: thermalise, time K steps, roofline, optional sweep points and CPU leg."""
    import torch
    spec = WORKLOADS[name]
    W = W or spec["walkers"]
    mol, wf = build_wf(spec, fixture)
    info = plan_info(wf)
    nt = 5 if spec["jastrow"] == "ee+een" else 0
    ens = thermalised(C, mol, wf, spec, W, spec["nbuf"], spec["therm"], 77)
    k = max(2, min(steps, 5)) if W * mol.nelec >= 20_000_000 else steps
    ms, kern_ms, st, launches, spec_on = time_eloc(C, wf, ens, W, k, min(warmup, 3) if W * mol.nelec >= 20_000_000 else warmup)
    ent = {"name": name, "baseline_config": BASELINE_CONFIGS[spec["baseline"]],
           "workload": config_dict(name, W, C.world)["workload"],
           "evals_per_s": C.world * W * k / (ms * 1e-3), "steps": k, "ms_per_step": ms / k, "kernel_ms": kern_ms,
           "walkers_per_gpu": W, "energy_hartree": st[0] / st[2] if st[2] else None, "n_nonfinite": st[3],
           "specialised_kernel": bool(spec_on), "launches_per_step": launches,
           "tile_walkers": info["tw_eloc"], "threads_per_cta": info["threads_eloc"], "smem_bytes": info["smem_eloc"],
           "roofline": roofline_obj(C, name if nt == 0 or name != "h2o" else "h2o", mol, wf, info, W, kern_ms, spec_on, nt)}
    if sweep and spec.get("sweep"):
        pts = []
        for fx, w in spec["sweep"]:
            if fx == spec["fixture"]:
                m2, wf2, e2 = mol, wf, [e[:w].contiguous() for e in ens]
            else:
                m2, wf2 = build_wf(spec, fx)
                e2 = thermalised(C, m2, wf2, spec, w, 1, spec["therm"], 99)
            i2 = plan_info(wf2)
            ms2, kms2, st2, _, on2 = time_eloc(C, wf2, e2, w, max(3, min(steps, 10)), 3)
            k2 = max(3, min(steps, 10))
            pts.append({"fixture": fx, "walkers_per_gpu": w, "nao": int(wf2.ao.norb), "nmo_used": i2["nmo_used"],
                        "evals_per_s": C.world * w * k2 / (ms2 * 1e-3), "kernel_ms": kms2,
                        "specialised_kernel": bool(on2),
                        "frac": w * algorithmic_flops(m2, wf2, i2) / (kms2 * 1e-3) / 1e12 / C.peak_tf,
                        "frac_executed": w * executed_flops(m2, wf2, i2) / (kms2 * 1e-3) / 1e12 / C.peak_tf})
            del e2
        ent["sweep"] = pts
    del ens
    torch.cuda.empty_cache()
    if with_cpu:
        try:
            ent["cpu_baseline"] = cpu_leg(name, 3, 1, budget_s=15.0)
        except Exception as ex:   # a CPU leg must never take the GPU numbers down with it
            ent["cpu_baseline"] = {"error": repr(ex)[:200]}
    return ent


def run_h2_single_point(C):
    """BASELINE config 1 literally: H2 STO-3G, Metropolis 1000 walkers x 2000 steps (step 0.5, as in the
    reference's tests/solver/test_h2_pyscf_metropolis.py:41-53), Solver.single_point; wall clock."""
    import torch
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    spec = WORKLOADS["h2"]
    mol, wf = build_wf(spec)
    out = {}
    for rep in range(2):       # first pass warms NVRTC / allocator
        torch.manual_seed(0)
        s = Metropolis(nwalkers=1000, nstep=2000, step_size=0.5, ntherm=-1, ndecor=1, nelec=wf.nelec, ndim=3,
                       init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=5)
        solver = Solver(wf=wf, sampler=s, optimizer=None)
        torch.cuda.synchronize(C.dev)
        t0 = time.perf_counter()
        obs = solver.single_point(with_tqdm=False)
        torch.cuda.synchronize(C.dev)
        out = {"walkers": 1000, "metropolis_steps": 2000, "wall_s": time.perf_counter() - t0,
               "psi_evals_per_s": 1000 * 2000 / (time.perf_counter() - t0), "energy": float(obs.energy),
               "error": float(obs.error), "acceptance": s.acceptance_rate}
    return out


def run_opt_config(C, steps, warmup, with_cpu):
    """BASELINE config 3: one optimisation step of LiH (Jastrow weight + MO coefficients trainable: 122
    doubles) = ONE E_L launch (E_L and psi) + 2-double all-reduce of (sum E_L, n) + qmcb_psi_backward
    (+ its reduce kernel) + ONE flat all-reduce of the gradients + Adam + table update
    (qmcb_plan_update, no recompilation).  Resampling between steps is NOT in the step (it is the
    Metropolis kernel timed elsewhere)."""
    import torch
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    from qmctorch_b200.solver import distributed as D
    spec = WORKLOADS["lih-opt"]
    W = spec["walkers"]
    mol, wf = build_wf(spec)
    pos = thermalised(C, mol, wf, spec, W, 1, spec["therm"], 55)[0]
    s = Metropolis(nwalkers=W, nstep=25, step_size=spec["step"], nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                   move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=9, keep_on_device=True)
    opt = torch.optim.Adam(wf.parameters(), lr=1e-3)
    solver = Solver(wf=wf, sampler=s, optimizer=opt, rank=C.rank)
    solver.configure(track=["local_energy"], freeze=["ci", "ao"], loss="energy", grad="manual")
    nparam = sum(p.numel() for p in solver._trainable())
    stream = torch.cuda.current_stream(C.dev)

    def step():
        opt.zero_grad()
        mean, _ = solver.evaluate_grad_manual(pos, allreduce=True)
        opt.step()
        wf._handle.plan()          # device tables follow the new parameters
        return mean
    for _ in range(max(warmup, 3)):
        step()
    C.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        mean = step()
    e1.record(stream)
    C.barrier()
    ms = C.max_over_ranks(e0.elapsed_time(e1))
    # parameters must be bit-identical on every rank after the all-reduced updates
    flat = torch.cat([p.detach().reshape(-1) for p in wf.parameters()])
    same = True
    if C.world > 1:
        import torch.distributed as dist
        ref = flat.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([float(torch.equal(ref, flat))], device=C.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    # the backward kernel alone
    from qmctorch_b200 import _lib
    wgt = torch.rand(W, dtype=torch.float64, device=C.dev)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wf._psi_backward(pos, wgt, {"mo_modifier", "jee_w"})
    b0.record(stream)
    for _ in range(5):
        wf._psi_backward(pos, wgt, {"mo_modifier", "jee_w"})
    b1.record(stream)
    torch.cuda.synchronize(C.dev)
    bwd_ms = b0.elapsed_time(b1) / 5
    wf._psi_backward(pos, wgt, None)
    b0.record(stream)
    for _ in range(3):
        wf._psi_backward(pos, wgt, None)
    b1.record(stream)
    torch.cuda.synchronize(C.dev)
    bwd_all_ms = b0.elapsed_time(b1) / 3
    ent = {"name": "lih-opt", "baseline_config": BASELINE_CONFIGS[2],
           "workload": config_dict("lih-opt", W, C.world)["workload"],
           "evals_per_s": C.world * W * steps / (ms * 1e-3), "unit": "walker-gradient evals/s", "steps": steps,
           "ms_per_step": ms / steps, "walkers_per_gpu": W, "trainable_doubles": nparam,
           "allreduce_bytes_per_step": 16 + 8 * nparam, "collectives_per_step": 2 if C.world > 1 else 0,
           "parameters_identical_on_all_ranks": same, "energy_hartree": float(mean),
           "backward_kernel_ms": {"jastrow+mo": bwd_ms, "all_parameters": bwd_all_ms},
           "specialised_kernel": wf._handle.info(13) == 1}
    if with_cpu:
        try:
            ent["cpu_baseline"] = cpu_leg("lih-opt", 3, 1, budget_s=15.0)
        except Exception as ex:
            ent["cpu_baseline"] = {"error": repr(ex)[:200]}
    return ent


def run_adjoint(C, with_cpu):
    """grad="auto" / forces: the adjoint of the local energy (qmcb_local_energy_backward) on the LiH ensemble -
    every wave-function parameter, the Jastrow + MO set of BASELINE config 3, and the atom coordinates - next
    to the reference back-propagating through its own wf.local_energy on the host cores (bounded sample)."""
    import torch
    spec = WORKLOADS["lih"]
    W = spec["walkers"]
    mol, wf = build_wf(spec)
    pos = thermalised(C, mol, wf, spec, W, 1, spec["therm"], 77)[0]
    wgt = torch.randn(W, dtype=torch.float64, device=C.dev)
    stream = torch.cuda.current_stream(C.dev)
    sets = {"all_parameters": {"bas_exp", "bas_coeffs", "mo_modifier", "ci", "jee_w"},
            "jastrow+mo": {"mo_modifier", "jee_w"}, "atom_coordinates": {"atom_coords"}}
    out = {"workload": "LiH 6-31G ground_state, %d walkers/GPU: sum_w w d E_L / d theta (backward of "
                       "wf.local_energy: Solver grad='auto', compute_forces)" % W, "walkers_per_gpu": W, "ms": {}}
    for name, want in sets.items():
        wf._eloc_backward(pos, wgt, None, want)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            wf._eloc_backward(pos, wgt, None, want)
        e1.record(stream)
        torch.cuda.synchronize(C.dev)
        out["ms"][name] = C.max_over_ranks(e0.elapsed_time(e1) / 3)
    out["walkers_per_s_all_parameters"] = C.world * W / (out["ms"]["all_parameters"] * 1e-3)
    if with_cpu:
        try:
            kind, rmol, rwf, _, sample = _reference_wf(spec)
            nw = 5000
            rpos = sample(nw, 10)
            cores = torch.get_num_threads()
            times = []
            if kind == "reference":
                leaves = [p_ for p_ in rwf.parameters() if p_.requires_grad]

                def step():
                    e = rwf.local_energy(rpos)
                    torch.autograd.grad(e.sum(), leaves, allow_unused=True)
            else:
                import sj_oracle as orc

                def step():
                    orc.local_energy_adjoint(rwf, rpos, w_eloc=torch.ones(nw, dtype=torch.float64))
            t_all = time.perf_counter()
            for i in range(4):
                t0 = time.perf_counter()
                step()
                if i:
                    times.append(time.perf_counter() - t0)
                if time.perf_counter() - t_all > 15.0 and times:
                    break
            med = statistics.median(times)
            out["cpu_baseline"] = {"value": nw / med, "unit": "walkers/s", "cores": cores, "kind": kind,
                                   "sample": "%d walkers x %d backward passes through wf.local_energy (autograd), "
                                             "torch CPU FP64, median" % (nw, len(times)), "ms_per_step": med * 1e3}
        except Exception as ex:
            out["cpu_baseline"] = {"error": repr(ex)[:200]}
    return out


def run_strong(C, steps, warmup):
    """BASELINE config 2 as written: 1e6 LiH walkers in TOTAL, contiguous shards over the N GPUs, the
    4-double all-reduce every step.  The step (kernel + all-reduce) is replayed from a CUDA graph of
    GRAPH steps so that 18 us kernels are not launch-bound; eager stepping is the fallback."""
    import torch
    import torch.distributed as dist
    from qmctorch_b200 import _lib
    from qmctorch_b200.solver import distributed as D
    spec = WORKLOADS["lih"]
    total = 1_000_000
    first, W = D.shard_walkers(total, C.rank, C.world)
    mol, wf = build_wf(spec)
    ens = thermalised(C, mol, wf, spec, W, 4, 50, 31)
    L = _lib.lib()
    dev = C.dev
    eloc = torch.empty(W, 1, dtype=torch.float64, device=dev)
    out4 = torch.zeros(4, dtype=torch.float64, device=dev)
    ws = torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8, device=dev)
    plan = wf._handle.plan()
    GRAPH = 10
    nrep = max(1, (steps + GRAPH - 1) // GRAPH)

    def body(i):
        _lib.check(L.qmcb_local_energy_stats(plan, _lib.ptr(ens[i % 4]), W, _lib.ptr(eloc), None, None,
                                             _lib.ptr(out4), _lib.ptr(ws), _lib.stream_ptr(dev)), "eloc")
        if C.world > 1:
            dist.all_reduce(out4)
    for i in range(max(warmup, 3)):
        body(i)
    torch.cuda.synchronize(dev)
    graph, used_graph = None, False
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(3):
                body(i)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(g, stream=side):
            for i in range(GRAPH):
                body(i)
        graph, used_graph = g, True
        graph.replay()
        torch.cuda.synchronize(dev)
    except Exception as ex:
        C.note("strong: CUDA graph capture failed (%s); eager stepping" % repr(ex)[:120])
        graph = None
        try:
            torch.cuda.synchronize(dev)
        except Exception:
            pass
    ok = torch.tensor([1.0 if graph is not None else 0.0], device=dev)
    if C.world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not bool(ok.item()):
        graph, used_graph = None, False
    stream = torch.cuda.current_stream(dev)
    C.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if graph is not None:
        for _ in range(nrep):
            graph.replay()
        nsteps = nrep * GRAPH
    else:
        for i in range(steps):
            body(i)
        nsteps = steps
    e1.record(stream)
    C.barrier()
    ms = C.max_over_ranks(e0.elapsed_time(e1))
    st = out4.tolist()
    return {"workload": "LiH 6-31G ground_state, %d walkers in TOTAL over %d GPU(s) (%d on this rank), E_L + "
                        "statistics + 4-double all-reduce every step" % (total, C.world, W),
            "scaling": "strong", "walkers_total": total, "evals_per_s": total * nsteps / (ms * 1e-3),
            "steps": nsteps, "ms_per_step": ms / nsteps, "cuda_graph": used_graph, "graph_steps": GRAPH if used_graph else 0,
            "energy_hartree": st[0] / st[2] if st[2] else None, "walkers_reduced": st[2]}


def run_e2e(C, wf, ens, W, steps):
    """The public call with HOST buffers: SlaterJastrow.local_energy(pinned host tensor) streams the walkers
    to the device in chunks (H2D of chunk k+1 overlaps the kernel on chunk k), then the statistics, the
    4-double all-reduce (N > 1) and the D2H of the four sums - every step.  Results land in a ring of
    pinned buffers; a slot is waited for before it is reused and all of them before the clock stops."""
    import torch
    import torch.distributed as dist
    from qmctorch_b200 import _lib
    L = _lib.lib()
    dev = C.dev
    host = [e.cpu().pin_memory() for e in ens[:2]]
    h2d = host[0].numel() * 8
    stream = torch.cuda.current_stream(dev)
    sp = _lib.stream_ptr(dev)
    ws = torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8, device=dev)
    NR = 4
    res_host = [torch.empty(4, dtype=torch.float64).pin_memory() for _ in range(NR)]
    res_dev = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(NR)]
    done = [torch.cuda.Event() for _ in range(NR)]

    def e2e_step(i):
        k = i % NR
        if i >= NR:
            done[k].synchronize()                 # the host has consumed slot k of NR steps ago
        e = wf.local_energy(host[i % 2])
        _lib.check(L.qmcb_energy_stats(_lib.ptr(e), W, _lib.ptr(res_dev[k]), _lib.ptr(ws), sp), "stats")
        if C.world > 1:
            dist.all_reduce(res_dev[k])
        res_host[k].copy_(res_dev[k], non_blocking=True)
        done[k].record(stream)
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    C.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(3, min(steps, 10))
    e0.record(stream)
    for i in range(n):
        e2e_step(i)
    for ev in done:
        ev.synchronize()
    e1.record(stream)
    C.barrier()
    ms = C.max_over_ranks(e0.elapsed_time(e1))
    energy = float(res_host[(n - 1) % NR][0] / res_host[(n - 1) % NR][2])
    # raw copy ceiling: the same pinned buffer -> device, nothing else
    xbuf = torch.empty_like(ens[0])
    for _ in range(2):
        xbuf.copy_(host[0], non_blocking=True)
    torch.cuda.synchronize(dev)
    C.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    for i in range(5):
        xbuf.copy_(host[i % 2], non_blocking=True)
    c1.record(stream)
    C.barrier()
    cms = C.max_over_ranks(c0.elapsed_time(c1)) / 5
    val = C.world * W * n / (ms * 1e-3)
    ceil = C.world * W / (cms * 1e-3)
    return {"value": val, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32, "steps": n,
            "api": "SlaterJastrow.local_energy(pinned host tensor) + qmcb_energy_stats + D2H of the 4 sums",
            "energy_hartree": energy,
            "h2d_copy_ceiling": {"evals_per_s": ceil, "gbs_per_gpu": h2d / (cms * 1e-3) / 1e9,
                                 "gbs_aggregate": C.world * h2d / (cms * 1e-3) / 1e9,
                                 "what": "cudaMemcpyAsync of the same pinned walkers alone, all ranks at once"},
            "frac_of_copy_ceiling": val / ceil}


def run_single_point_e2e(C):
    """Second end-to-end figure: Solver.single_point through the public API, LiH, 1e6 walkers x 500
    Metropolis steps + E_L + statistics, wall clock; ensemble drawn on the device (init_rng='philox')."""
    import torch
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import Solver
    spec = WORKLOADS["lih"]
    mol, wf = build_wf(spec)
    out = None
    for rep in range(2):
        torch.manual_seed(3)
        s = Metropolis(nwalkers=1_000_000, nstep=500, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                       move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=2, init_rng="philox")
        solver = Solver(wf=wf, sampler=s, optimizer=None, rank=C.rank)
        torch.cuda.synchronize(C.dev)
        C.barrier()
        t0 = time.perf_counter()
        obs = solver.single_point(with_tqdm=False)
        torch.cuda.synchronize(C.dev)
        dt = time.perf_counter() - t0
        out = {"walkers_per_gpu": 1_000_000, "metropolis_steps": 500, "wall_s": dt,
               "walker_steps_per_s": C.world * 1_000_000 * 500 / dt, "energy": float(obs.energy),
               "error": float(obs.error), "init_rng": "philox"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lih", choices=sorted(WORKLOADS))
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: workload's)")
    ap.add_argument("--configs", default="all", help="all | none | comma list of h2,lih,lih-opt,h2o,c4h6")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = WORKLOADS[args.workload]
    wpg = args.walkers or spec["walkers"]
    cfg = config_dict(args.workload, wpg, world)
    metric = "local_energy_evals_per_s"

    if args.impl == "reference":
        if rank != 0:
            return
        leg = cpu_leg(args.workload, args.steps, args.warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": metric, "value": leg["value"], "unit": "evals/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg,
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample", "steps_timed")},
                "e2e": {"value": leg["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from qmctorch_b200 import build as _build
    if not os.path.isfile(_build.LIB):
        if local_rank == 0:
            _build.build()
        else:
            while not os.path.isfile(_build.LIB):
                time.sleep(1.0)
    from qmctorch_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner goes to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    C = Ctx()
    C.rank, C.world, C.dev, C.notes = rank, world, dev, []
    C.note = lambda s: C.notes.append(s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    C.barrier, C.max_over_ranks = barrier, max_over_ranks
    C.kernel_name = lambda wf: ("spec kernel (NVRTC, structure-specialised E_L)" if wf._handle.info(13) == 1
                                else "fused_kernel<MODE_ELOC>")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    C.hbm_peak = peaks.get("hbm_gbs", 6650.0)
    C.hbm_src = "MEASURED_PEAKS.json" if peaks else "fallback"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- FP64 pipe peak (own probe: dependent-chain-free DFMA loop over the whole GPU)
    import ctypes
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    fl = ctypes.c_double(0.0)
    stream = torch.cuda.current_stream(dev)
    sp = _lib.stream_ptr(dev)
    C.peak_tf = None
    for rep in range(3):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        L.qmcb_fp64_probe(0, 20000, _lib.ptr(sink), ctypes.byref(fl), sp)
        p1.record(stream)
        torch.cuda.synchronize(dev)
        tf = fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12
        C.peak_tf = tf if C.peak_tf is None else max(C.peak_tf, tf)

    # ---- headline workload
    with_cpu = (not args.no_cpu_baseline) and world == 1
    if spec.get("kind") == "opt":
        ent = run_opt_config(C, args.steps, args.warmup, with_cpu)
        value, ms_per_step, roof, launches, e2e = ent["evals_per_s"], ent["ms_per_step"], None, None, None
        main_ent = ent
    else:
        mol, wf = build_wf(spec)
        info = plan_info(wf)
        nt = 5 if spec["jastrow"] == "ee+een" else 0
        ens = thermalised(C, mol, wf, spec, wpg, spec["nbuf"], spec["therm"], 0)
        elapsed_ms, kern_ms, st, lps, spec_on = time_eloc(C, wf, ens, wpg, args.steps, args.warmup)
        value = world * wpg * args.steps / (elapsed_ms * 1e-3)
        ms_per_step = elapsed_ms / args.steps
        roof = roofline_obj(C, args.workload, mol, wf, info, wpg, kern_ms, spec_on, nt)
        launches = lps * args.steps
        e2e = run_e2e(C, wf, ens, wpg, args.steps) if wpg * mol.nelec < 20_000_000 else None
        main_ent = {"name": args.workload, "baseline_config": BASELINE_CONFIGS[spec["baseline"]],
                    "workload": cfg["workload"], "evals_per_s": value, "steps": args.steps,
                    "ms_per_step": ms_per_step, "kernel_ms": kern_ms, "walkers_per_gpu": wpg,
                    "energy_hartree": st[0] / st[2] if st[2] else None, "n_nonfinite": st[3],
                    "specialised_kernel": bool(spec_on), "launches_per_step": lps,
                    "tile_walkers": info["tw_eloc"], "threads_per_cta": info["threads_eloc"],
                    "smem_bytes": info["smem_eloc"], "roofline": roof}
        del ens
        torch.cuda.empty_cache()
    clocks = None
    if rank == 0:
        clocks = sampler.summary()      # clocks of the headline region (sampling goes on for the rest)
    if with_cpu and "cpu_baseline" not in main_ent:
        try:
            main_ent["cpu_baseline"] = cpu_leg(args.workload, 3, 1, budget_s=20.0)
        except Exception as ex:
            main_ent["cpu_baseline"] = {"error": repr(ex)[:200]}

    # ---- the other BASELINE configurations
    want = []
    if args.configs == "all":
        want = ["h2", "lih", "lih-opt", "h2o", "c4h6"]
    elif args.configs != "none":
        want = [w for w in args.configs.split(",") if w in WORKLOADS]
    entries = []
    for name in want:
        try:
            if name == args.workload:
                ent = dict(main_ent)
            elif WORKLOADS[name].get("kind") == "opt":
                ent = run_opt_config(C, max(3, min(args.steps, 20)), 3, with_cpu)
            else:
                ent = run_eloc_config(C, name, max(3, min(args.steps, 20)), 3, with_cpu)
            if name == "h2":
                ent["single_point_1000x2000"] = run_h2_single_point(C)
        except Exception as ex:
            ent = {"name": name, "baseline_config": BASELINE_CONFIGS[WORKLOADS[name]["baseline"]],
                   "error": repr(ex)[:300]}
            torch.cuda.synchronize(dev)
        entries.append(ent)
    strong = None
    if not args.no_strong:
        try:
            strong = run_strong(C, max(args.steps, 20), 3)
        except Exception as ex:
            strong = {"error": repr(ex)[:300]}
    adjoint = None
    if args.configs == "all":
        try:
            adjoint = run_adjoint(C, with_cpu)
        except Exception as ex:
            adjoint = {"error": repr(ex)[:300]}
            torch.cuda.synchronize(dev)
    sp_e2e = None
    if e2e is not None and args.configs != "none":
        try:
            sp_e2e = run_single_point_e2e(C)
        except Exception as ex:
            sp_e2e = {"error": repr(ex)[:300]}
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": metric, "value": value, "unit": "evals/s" if spec.get("kind") != "opt" else "walker-gradient evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "result": {k: main_ent.get(k) for k in ("energy_hartree", "n_nonfinite", "specialised_kernel",
                                                               "tile_walkers", "threads_per_cta", "smem_bytes")},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": e2e,
        "configs": entries, "strong": strong, "single_point_e2e": sp_e2e, "adjoint": adjoint, "notes": C.notes,
    }
    if with_cpu:
        line["cpu_baseline"] = main_ent.get("cpu_baseline")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
