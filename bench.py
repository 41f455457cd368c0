"""Benchmark of the hot path: local-energy evaluations per second (walkers x steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lih|h2|h2o|c4h6] [--walkers M]
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on host cores

A "step" is ONE local-energy evaluation of every walker of this rank (BASELINE config 2: LiH
6-31G single determinant + Pade Jastrow, 1e6 walkers per GPU) followed by the energy statistics
(deterministic two-stage reduction; one 4-double all-reduce when N > 1).  Walkers are
thermalised by the fused Metropolis kernel beforehand and are resident in HBM when the timed
region starts; four independent ensembles (4 x 96 MB > 126 MB L2) are cycled so that no step
re-reads an ensemble that is still in L2.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # key: (fixture, configs, walkers/GPU, Metropolis step size, cpu sample walkers)
    "lih": ("lih", "ground_state", 1_000_000, 0.3, 100_000),
    "h2": ("h2", "single(2,2)", 1_000_000, 0.5, 100_000),
    "h2o": ("h2o", "cas(4,4)", 250_000, 0.15, 10_000),
    "c4h6": ("c4h6", "ground_state", 100_000, 0.05, 2_000),
}


def algorithmic_flops(mol, wf, info):
    """Flops per local-energy evaluation of the formulas this repo evaluates:
    add/mul = 1, fma = 2, div/sqrt/exp = 1 each; only the MO columns some configuration occupies."""
    ne, nat = mol.nelec, mol.natom
    nprim, ncomp, nmu = info["nprim"], info["ncomp"], info["nmo_used"]
    npair = ne * (ne - 1) // 2
    ao = ne * (nat * 8 + nprim * 12 + ncomp * 12)          # r^2; exp + radial sums; harmonic products
    mo = ne * ncomp * nmu * 5 * 2                         # 5 channels contracted on the fly
    bkin = ne * nmu * 10
    jast = 2 * npair * 45 + ne * nat * 6                  # ordered pairs (each pair visited twice) + V_en
    n = max(mol.nup, mol.ndown)
    nun = info["nuniq_up"] + info["nuniq_down"]
    slater = nun * (12 if n <= 2 else 4 * n ** 3)
    return ao + mo + bkin + jast + slater + 4 * wf.nci + 10


def executed_flops(mol, wf, info):
    """Same conventions, for the formulas the E_L kernels execute since the kinetic channel is folded
    per AO: two projected channels instead of five, lap R from two radial sums,
    every electron pair visited once in one-walker-per-thread kernels."""
    ne, nat = mol.nelec, mol.natom
    nprim, ncomp, nmu, nshell = info["nprim"], info["ncomp"], info["nmo_used"], info["nshell"]
    npair = ne * (ne - 1) // 2
    ao = ne * (nat * 14 + nprim * 8 + nshell * 6 + ncomp * 4)
    mo = ne * ncomp * nmu * 2 * 2
    jast = npair * 60 + ne * nat * 2
    n = max(mol.nup, mol.ndown)
    nun = info["nuniq_up"] + info["nuniq_down"]
    slater = nun * (12 if n <= 2 else 4 * n ** 3)
    return ao + mo + ne * nmu + jast + slater + 4 * wf.nci + 10


# dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the dominant
# kernel, bytes per launch, keyed by (kernel, workload, walkers): profiles/r1_spec_ncu_raw.csv
# (structure-specialised kernel, final capture spec_r1i: profiles/r1_fold_ncu_raw.csv) and
# profiles/r1_fused_ncu_raw.csv (generic kernel)
NCU_TRAFFIC = {("spec_eloc", "lih", 1_000_000): 96.036608e6 + 4.964608e6,
               ("fused_kernel<MODE_ELOC>", "lih", 1_000_000): 96.070912e6 + 5.570048e6}


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


def cpu_reference_run(key, nsample, steps, warmup):
    """The reference algorithm (oracle port, torch CPU, all host threads) on a bounded sample."""
    import torch
    import sj_oracle as orc
    from qmctorch_b200.molecules import fixture_molecule
    from qmctorch_b200.wavefunction.pooling import OrbitalConfigurations
    fixture, configs, _, step_size, _ = WORKLOADS[key]
    torch.set_num_threads(os.cpu_count() or 1)
    mol = fixture_molecule(fixture)
    cfg = OrbitalConfigurations(mol).get_configs(configs)
    P = orc.make_params(mol, cfg, jastrow_weight=1.0)
    g = torch.Generator().manual_seed(0)
    mean = torch.as_tensor(mol.domain("normal")["mean"])
    sig = torch.as_tensor(mol.domain("normal")["sigma"]).diagonal().sqrt()
    pos = (mean + sig * torch.randn(nsample, mol.nelec, 3, generator=g, dtype=torch.float64)).view(nsample, -1)
    with torch.no_grad():
        fx = (orc.psi(P, pos) ** 2).reshape(-1)
        for _ in range(10):       # a few Metropolis moves so that |psi|^2 is roughly sampled
            d = torch.randn(pos.shape, generator=g, dtype=torch.float64) * (orc.proposal_sigma(step_size) ** 0.5)
            tau = torch.rand(nsample, generator=g, dtype=torch.float64)
            pos, fx, _, _ = orc.metropolis_step(P, pos, fx, d, tau)
        for _ in range(warmup):
            orc.local_energy(P, pos)
        t0 = time.perf_counter()
        for _ in range(steps):
            e = orc.local_energy(P, pos)
        dt = time.perf_counter() - t0
    return nsample * steps / dt, dt / steps, float(e.mean()), torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lih", choices=sorted(WORKLOADS))
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--therm", type=int, default=100, help="Metropolis thermalisation moves per ensemble")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fixture, configs, wpg, step_size, ncpu = WORKLOADS[args.workload]
    if args.walkers:
        wpg = args.walkers
    cfg_common = {"workload": "%s %s, %d walkers/GPU, VMC local energy (Jacobi kinetic) + energy statistics"
                              % ({"lih": "LiH 6-31G", "h2": "H2 STO-3G", "h2o": "H2O cc-pVDZ",
                                  "c4h6": "C4H6 DZP"}[args.workload], configs, wpg),
                  "walkers_per_gpu": wpg, "jastrow": "pade e-e", "parallelism": "walker shards x%d" % world,
                  "l2": "4 ensembles cycled (4x input > L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        val, per, emean, thr = cpu_reference_run(args.workload, ncpu, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": "local_energy_evals_per_s", "value": val, "unit": "evals/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg_common,
                "cpu_baseline": {"value": val, "unit": "evals/s", "cores": thr, "kind": "port",
                                 "sample": "%d walkers x %d steps of oracle.local_energy (torch CPU FP64, "
                                           "reference algorithm)" % (ncpu, steps)},
                "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    from qmctorch_b200.molecules import fixture_molecule
    from qmctorch_b200.sampler import Metropolis
    from qmctorch_b200.solver import distributed as D
    from qmctorch_b200.wavefunction import SlaterJastrow

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed to stdout
        # with NCCL_DEBUG=VERSION, which some images export) goes to stderr instead
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    mol = fixture_molecule(fixture)
    wf = SlaterJastrow(mol, configs=configs, cuda=True)
    info = {n: wf._handle.info(i) for i, n in enumerate(
        ["nshell", "nprim", "ncomp", "nmo_used", "nuniq_up", "nuniq_down", "tw_eloc", "threads_eloc",
         "smem_eloc", "tw_psi"])}
    L = _lib.lib()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()           # samples through thermalisation, the timed region and the e2e leg
    # ---- synthetic ensembles: reference 'normal' init, thermalised on the device
    NBUF = 4
    torch.manual_seed(1234 + rank)
    ens = []
    for b in range(NBUF):
        s = Metropolis(nwalkers=wpg, nstep=args.therm, step_size=step_size, nelec=wf.nelec, ndim=3,
                       init=mol.domain("normal"), move={"type": "all-elec", "proba": "normal"}, cuda=True,
                       seed=1000 * rank + b, keep_on_device=True)
        ens.append(s(wf.pdf, with_tqdm=False).detach().contiguous())
    W = wpg
    eloc = torch.empty(W, 1, dtype=torch.float64, device=dev)
    out4 = torch.zeros(4, dtype=torch.float64, device=dev)
    # N > 1: the 4-double all-reduce of step i is issued asynchronously (its own buffer) and only
    # waited for at the end of the timed region, so it overlaps the kernel of step i+1
    ring = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(8)]
    ws = torch.empty(int(L.qmcb_stats_workspace_bytes(W)), dtype=torch.uint8, device=dev)
    plan = wf._handle.plan()
    stream = torch.cuda.current_stream(dev)
    sp = _lib.stream_ptr(dev)

    def step(i):
        x = ens[i % NBUF]
        _lib.check(L.qmcb_local_energy_stats(plan, _lib.ptr(x), W, _lib.ptr(eloc), None, None, _lib.ptr(out4),
                                             _lib.ptr(ws), sp), "local_energy_stats")
        if world > 1:
            dist.all_reduce(out4)
    # one call = ONE kernel when structure-specialised (E_L with both statistics stages fused: the last
    # CTA adds the partials); generic kernels: E_L + two statistics kernels
    two_stage = os.environ.get("QMCB_STATS_2STAGE", "0") not in ("", "0")
    launches_per_step = (2 if two_stage else 1) if wf._handle.info(13) == 1 else 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    # timed region: K steps; the dominant kernel is also timed alone with its own event pairs
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    pending = []
    for i in range(args.steps):
        x = ens[i % NBUF]
        o4 = ring[i % len(ring)] if world > 1 else out4
        if world > 1 and len(pending) >= len(ring):
            pending.pop(0).wait()          # the buffer about to be reused must have been reduced
        kev[i][0].record(stream)
        _lib.check(L.qmcb_local_energy_stats(plan, _lib.ptr(x), W, _lib.ptr(eloc), None, None, _lib.ptr(o4),
                                             _lib.ptr(ws), sp), "local_energy_stats")
        kev[i][1].record(stream)
        if world > 1:
            pending.append(dist.all_reduce(o4, async_op=True))
    for h in pending:
        h.wait()
    if world > 1:
        out4.copy_(ring[(args.steps - 1) % len(ring)])
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    # event pairs around each qmcb_local_energy_stats call: the E_L kernel with its fused statistics
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t)
    stats = out4.tolist()
    energy = stats[0] / stats[2] if stats[2] else float("nan")

    # ---- e2e: public API with HOST buffers: H2D of the walkers, E_L, statistics, D2H of the result
    host = [e.cpu().pin_memory() for e in ens[:2]]
    h2d = host[0].numel() * 8
    res_host = torch.empty(4, dtype=torch.float64).pin_memory()

    def e2e_step(i):
        # the public call with a HOST tensor: SlaterJastrow.local_energy streams it to the device in
        # chunks on two side streams, the copy of chunk k+1 overlapping the kernel on chunk k
        e = wf.local_energy(host[i % 2])
        _lib.check(L.qmcb_energy_stats(_lib.ptr(e), W, _lib.ptr(out4), _lib.ptr(ws), sp), "stats")
        if world > 1:
            dist.all_reduce(out4)
        res_host.copy_(out4, non_blocking=True)
        stream.synchronize()
        return res_host[0].item()
    for i in range(2):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nsteps_e2e = max(3, min(args.steps, 10))
    e0.record(stream)
    for i in range(nsteps_e2e):
        e2e_step(i)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * W * nsteps_e2e / (float(t) * 1e-3)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    # last collective is behind us: leave the process group together, rank 0 finishes alone
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    # ---- FP64 pipe peak (own probe: dependent-chain-free DFMA loop over the whole GPU)
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    import ctypes
    fl = ctypes.c_double(0.0)
    peak_tf = None
    for rep in range(3):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        L.qmcb_fp64_probe(0, 20000, _lib.ptr(sink), ctypes.byref(fl), sp)
        p1.record(stream)
        torch.cuda.synchronize(dev)
        tf = fl.value / (p0.elapsed_time(p1) * 1e-3) / 1e12
        peak_tf = tf if peak_tf is None else max(peak_tf, tf)

    if rank != 0:
        return
    value = world * W * args.steps / (elapsed_ms * 1e-3)
    F = algorithmic_flops(mol, wf, info)
    # the local-energy call runs the NVRTC structure-specialised kernel when the plan has one
    kernel_name = "spec_eloc" if wf._handle.info(13) == 1 else "fused_kernel<MODE_ELOC>"
    achieved_tf = W * F / (kern_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_per_eval = 24 * mol.nelec + 8
    line = {
        "metric": "local_energy_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(cfg_common, energy_hartree=energy, tile_walkers=info["tw_eloc"],
                       threads_per_cta=info["threads_eloc"], smem_bytes=info["smem_eloc"]),
        "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                "steps": nsteps_e2e, "api": "SlaterJastrow.local_energy(pinned host tensor) + qmcb_energy_stats"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(),
        "roofline": {"bound": "fp64", "kernel": kernel_name, "achieved": achieved_tf,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                     "peak_source": "own DFMA probe (qmcb_fp64_probe) on this GPU; MEASURED_PEAKS.json has no FP64 entry",
                     "flops_per_eval": F, "flops_convention": "SURVEY 8(d): reference formulation, five projected "
                     "AO channels, occupied MO columns only", "flops_executed_per_eval": executed_flops(mol, wf, info),
                     "kernel_ms": kern_ms, "traffic": NCU_TRAFFIC.get((kernel_name, args.workload, W)),
                     "hbm": {"achieved_gbs": W * bytes_per_eval / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "bytes_per_eval": bytes_per_eval,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
    }
    if not args.no_cpu_baseline:
        val, per, emean, thr = cpu_reference_run(args.workload, ncpu, 3, 1)
        line["cpu_baseline"] = {"value": val, "unit": "evals/s", "cores": thr, "kind": "port",
                                "sample": "%d walkers x 3 steps of oracle.local_energy (torch CPU FP64, "
                                          "reference algorithm)" % ncpu}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
