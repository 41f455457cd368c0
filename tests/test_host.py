"""CPU: host-side logic - the C-ABI library loads and exports what include/qmcb.h declares,
plan grouping, configurations, the generic sampler path, loud failure without CUDA, and the
world_size-2 (gloo) collectives."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import _cases as C
from qmctorch_b200 import _lib
from qmctorch_b200.molecules import fixture_molecule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qmcb.h")).read()
    declared = set(re.findall(r"\b(qmcb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "libqmcb.so does not export %s" % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().qmcb_abi_version() == 1


def test_plan_grouping_shares_exponentials():
    from qmctorch_b200.wavefunction import SlaterJastrow
    info = SlaterJastrow(fixture_molecule("lih"))._handle.host_plan_info()
    # 26 flat primitives -> 17 exponentials: p components share one exp per primitive, and
    # the diffuse Li s/p shells with the same exponent share theirs
    assert info["nprim"] == 17 and info["ncomp"] == 11 and info["nmo_used"] == 2
    assert info["nuniq_up"] == 1 and info["nuniq_down"] == 1
    info = SlaterJastrow(fixture_molecule("h2o"), configs="cas(4,4)")._handle.host_plan_info()
    assert info["ncomp"] == 25 and info["nmo_used"] == 7
    assert info["nuniq_up"] == 6 and info["nuniq_down"] == 6      # 36 configurations, 6+6 determinants
    info = SlaterJastrow(fixture_molecule("c4h6"))._handle.host_plan_info()
    assert info["ncomp"] == 94 and info["nmo_used"] == 15 and info["smem_eloc"] <= 227 * 1024


def test_plan_rejects_bad_configuration():
    from qmctorch_b200.wavefunction import SlaterJastrow
    mol = fixture_molecule("lih")
    bad = (torch.tensor([[0, 99]]), torch.tensor([[0, 1]]))
    wf = SlaterJastrow(mol, configs=bad)
    with pytest.raises(RuntimeError, match="outside"):
        wf._handle.host_plan_info()


def test_no_cpu_fallback():
    from qmctorch_b200.wavefunction import SlaterJastrow
    wf = SlaterJastrow(fixture_molecule("h2"), cuda=False)
    pos = torch.zeros(4, 6, dtype=torch.float64)
    for call in (lambda: wf(pos), lambda: wf.local_energy(pos), lambda: wf.pdf(pos),
                 lambda: wf.ao(pos), lambda: wf.gradients_jacobi(pos)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_constructor_errors_match_reference():
    from qmctorch_b200.wavefunction import SlaterJastrow
    from qmctorch_b200.sampler import Metropolis
    mol = fixture_molecule("lih")
    with pytest.raises(ValueError):
        SlaterJastrow(mol, configs="cas(2,2)", include_all_mo=False)
    with pytest.raises(ValueError):
        SlaterJastrow(mol, configs="triple(2,2)")
    with pytest.raises(ValueError):
        Metropolis(move={"type": "two-elec", "proba": "normal"})
    s = Metropolis(nwalkers=4, nstep=5, ntherm=7, nelec=2, init=mol.domain("normal"))
    with pytest.raises(ValueError, match="Thermalisation"):
        s(lambda x: (x ** 2).sum(1))


def test_state_dict_names_match_reference():
    from qmctorch_b200.wavefunction import SlaterJastrow
    wf = SlaterJastrow(fixture_molecule("lih"), configs="single_double(2,2)")
    assert list(wf.state_dict()) == ["ao.atom_coords", "ao.bas_exp", "mo.mo_modifier", "fc.weight",
                                     "jastrow.jastrow_kernel.weight"]
    assert wf.fc.weight.shape == (1, wf.nci) and float(wf.fc.weight[0, 0]) == 1.0
    assert wf.get_number_parameters() > 0


def test_generic_sampler_reproduces_reference_algorithm(double_default):
    """Arbitrary pdf callable -> the reference's torch loop; same generator calls, so the
    same seed gives the same trajectory as a direct restatement."""
    from qmctorch_b200.sampler import Metropolis
    mol = fixture_molecule("h2")
    pdf = lambda x: torch.exp(-(x ** 2).sum(1))   # noqa: E731
    torch.manual_seed(3)
    s = Metropolis(nwalkers=50, nstep=20, step_size=0.5, nelec=2, ndim=3, init=mol.domain("normal"),
                   move={"type": "all-elec", "proba": "normal"})
    out = s(pdf, with_tqdm=False)
    # restatement with the same draws
    torch.manual_seed(3)
    from torch.distributions import MultivariateNormal
    d = mol.domain("normal")
    pos = MultivariateNormal(torch.as_tensor(d["mean"]), torch.as_tensor(d["sigma"])).sample((50, 2))
    pos = pos.type(torch.float64).view(50, 6)
    fx = pdf(pos)
    mv = s.multiVariate
    for _ in range(20):
        xn = pos + mv.sample((50, 2)).view(50, 6)
        fxn = pdf(xn)
        df = fxn / fx
        df[df > 1] = 1.0
        acc = (df - torch.rand_like(df) >= 0)
        pos[acc] = xn[acc]
        fx[acc] = fxn[acc]
    assert torch.equal(out.detach(), pos)
    assert out.requires_grad and s.get_sampling_size() == 50


def test_walker_initialisation_matches_reference_fixture(double_default):
    """Walkers.initialize (sampler/walkers.py:41-150): same generator calls in the same order as the
    reference for every Molecule.domain method -> bit-identical start ensembles."""
    from qmctorch_b200.sampler.ensemble import Walkers
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "walkers_init.npz"))
    for key in ("lih", "h2o"):
        mol = fixture_molecule(key)
        for method in ("center", "uniform", "normal", "atomic"):
            torch.manual_seed(5)
            np.random.seed(5)
            w = Walkers(nwalkers=9, nelec=mol.nelec, ndim=3, init=mol.domain(method))
            w.initialize()
            assert w.pos.dtype == torch.float64
            assert np.array_equal(w.pos.numpy(), gold["%s_%s" % (key, method)]), (key, method)
    with pytest.raises(ValueError):
        Walkers(nwalkers=2, nelec=2, init={"bogus": 1}).initialize()
    w = Walkers(nwalkers=3, nelec=2, init=mol.domain("center"))
    w.initialize(pos=torch.arange(30.0).reshape(5, 6))
    assert torch.equal(w.pos, torch.arange(30.0).reshape(5, 6)[-3:])


def test_specialised_kernel_source_compiles_without_a_gpu():
    """spec.cu generates the straight-line program of a wave function and NVRTC compiles it for sm_100a
    on a host-only plan (no driver needed): one walker per thread for small systems (kind 1), warp
    tiles for mid-size / large ones (kind 2: H2O CAS with the three-body Jastrow); structures that
    fit neither (more than 32 electrons) stay on the generic kernels."""
    from qmctorch_b200.wavefunction import SlaterJastrow
    L = _lib.lib()

    def status(key, configs):
        wf = SlaterJastrow(fixture_molecule(key), configs=configs, cuda=False)
        arrays = wf._handle._system()
        p = ctypes.c_void_p()
        _lib.check(L.qmcb_plan_create(ctypes.byref(arrays.struct), -1, ctypes.byref(p)), "qmcb_plan_create")
        out = (L.qmcb_plan_info(p, 14), L.qmcb_plan_info(p, 13), L.qmcb_last_error().decode())
        L.qmcb_plan_destroy(p)
        return out
    elig, on, why = status("lih", "single_double(2,2)")
    assert elig == 1
    if not on and "libnvrtc not found" in why:
        pytest.skip("NVRTC is not installed here")
    assert on == 1, why
    elig, on, why = status("h2o", "cas(4,4)")
    assert (elig, on) == (2, 1), why
    # 34 electrons: a walker does not fit one warp
    from qmctorch_b200.molecules import Molecule, _seeded_mos, build_basis, _parse_atoms
    atoms = "C 0 0 0; C 0 0 1.4; C 0 1.3 2.1; C 0 2.5 1.4; C 0 2.5 0; H 0 -0.9 -0.5; H 0 -0.9 1.9; H 0 3.4 1.9; H 0 3.4 -0.5"
    names, coords = _parse_atoms(atoms, "angs")
    mol = Molecule(atoms, basis="dz", unit="angs", name="C5H4", mos=_seeded_mos(build_basis(names, coords, "dz").nao, 3))
    wf = SlaterJastrow(mol, configs="ground_state", cuda=False)
    assert mol.nelec == 34
    arrays = wf._handle._system()
    p = ctypes.c_void_p()
    _lib.check(L.qmcb_plan_create(ctypes.byref(arrays.struct), -1, ctypes.byref(p)), "qmcb_plan_create")
    assert (L.qmcb_plan_info(p, 14), L.qmcb_plan_info(p, 13)) == (0, 0)
    assert "not eligible" in L.qmcb_last_error().decode()
    L.qmcb_plan_destroy(p)


def test_generated_program_shares_sp_exponentials_and_launch_shapes(tmp_path, monkeypatch):
    """Host logic of the second half of round 1: (1) the generated LiH 6-31G program evaluates ONE
    exponential per distinct exponent of an atom (the s and p functions of the SP shells share theirs:
    14 for 17 grouped primitives) and as many as there are primitives once an update has separated the
    exponents; (2) the 16-column kernels of large systems run as 192-thread CTAs (C4H6: 6 walkers)."""
    from qmctorch_b200.wavefunction import SlaterJastrow
    L = _lib.lib()

    def generated(wf, tag):
        monkeypatch.setenv("QMCB_JIT_DUMP", str(tmp_path / tag))
        arrays = wf._handle._system()
        p = ctypes.c_void_p()
        _lib.check(L.qmcb_plan_create(ctypes.byref(arrays.struct), -1, ctypes.byref(p)), "qmcb_plan_create")
        on, why = L.qmcb_plan_info(p, 13), L.qmcb_last_error().decode()
        nprim = L.qmcb_plan_info(p, 1)
        L.qmcb_plan_destroy(p)
        if not on and "libnvrtc not found" in why:
            pytest.skip("NVRTC is not installed here")
        assert on == 1, why
        src = open(str(tmp_path / tag) + ".cu").read()
        body = src[src.index("void spec_aos("):src.index("void spec_dets(")]
        return len(re.findall(r"= spec_exp<MODE, \d+>", body)), nprim

    wf = SlaterJastrow(fixture_molecule("lih"), configs="ground_state", cuda=False)
    nexp, nprim = generated(wf, "shared")
    assert (nexp, nprim) == (14, 17)
    with torch.no_grad():
        torch.manual_seed(0)
        wf.ao.bas_exp.mul_(1.0 + 0.05 * torch.rand_like(wf.ao.bas_exp))
    nexp, nprim = generated(wf, "separate")
    assert nexp == nprim == 26          # every flat primitive is its own shell now
    info = SlaterJastrow(fixture_molecule("c4h6"))._handle.host_plan_info()
    assert info["threads_eloc"] == 192 and info["tw_eloc"] == 6 and info["tw_psi"] == 6


def test_bench_flop_counts_and_reference_arm_line():
    """bench.py host logic: the SURVEY 8(d) unit of work for LiH 6-31G is 2994 flop/eval, the
    kernels' executed count is smaller (folded kinetic channel), and the reference arm prints one
    JSON line with the contract's keys."""
    import json
    import bench
    from qmctorch_b200.wavefunction import SlaterJastrow
    mol = fixture_molecule("lih")
    wf = SlaterJastrow(mol, configs="ground_state", cuda=False)
    info = wf._handle.host_plan_info()
    assert bench.algorithmic_flops(mol, wf, info) == 2994
    assert 0.4 * 2994 < bench.executed_flops(mol, wf, info) < 2994
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "local_energy_evals_per_s" and line["value"] > 0
    # the real reference wherever it is present (/root/reference here, baseline/_ref on the GPU box), else the port
    import ref_shim
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["steps"] == 1
    assert "LiH 6-31G" in line["config"]["workload"] and line["config"]["cpu_sample_walkers"] == 50000
    # the b200 arm and the reference arm must print the same config object
    assert line["config"] == bench.config_dict("lih", 1_000_000, 1)


def test_molecule_load_json_dump_of_adf_hdf5():
    """Molecule(load=...) (scf/molecule.py:96-100,394-402): the committed JSON dumps of the reference's
    tests/hdf5/*_adf_*.hdf5 carry the ADF basis verbatim; MOs are not renormalised."""
    mol = fixture_molecule("lih_adf")
    b = mol.basis
    assert (mol.nelec, mol.nup, mol.ndown, mol.natom, mol.name) == (4, 2, 2, 2, "LiH")
    assert (b.nao, b.nmo, b.radial_type, b.harmonics_type) == (9, 9, "sto", "cart")
    assert list(mol.atoms) == ["Li", "H"] and mol.atom_coords[1] == [0.0, 0.0, 3.015]
    assert b.TotalEnergy == -7.976594691003629 and b.bas_exp[0] == 4.24 and int(b.bas_kr[2]) == 1
    assert b.mos.shape == (9, 9) and abs(b.mos[0, 0] + 0.211348386) < 1e-9
    assert abs(float((b.mos[:, 0] ** 2).sum()) - 1.0) > 1e-3      # non-orthogonal AO basis: taken verbatim
    assert mol.hdf5file.endswith("LiH_adf_dz.json")
    co2 = fixture_molecule("co2_adf")
    assert (co2.nelec, co2.basis.nao, co2.basis.nmo) == (22, 48, 45)
    assert mol.domain("atomic")["atom_num"] == [3, 1]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/hdf5"), reason="reference HDF5 files not on this box")
@pytest.mark.parametrize("stem", ["LiH_adf_dz", "H2_adf_dzp", "CO2_adf_dzp"])
def test_hdf5_reader_against_reference_files(stem):
    """utils/hdf5_min.py (pure-Python HDF5 subset) reads the reference's own molecule files and gives
    exactly what the committed JSON dump holds (tools/hdf5_to_fixture.py)."""
    from qmctorch_b200.molecules import Molecule
    from qmctorch_b200.utils.hdf5_min import read_hdf5
    path = "/root/reference/tests/hdf5/%s.hdf5" % stem
    tree = read_hdf5(path)["molecule"]
    assert tree["basis"]["harmonics_type"] == "cart" and tree["basis"]["radial_type"] == "sto"
    a = Molecule(load=path)
    b = Molecule(load=os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", stem + ".json"))
    for n in ["bas_exp", "bas_coeffs", "mos", "bas_kx", "bas_ky", "bas_kz", "bas_kr", "index_ctr", "nctr_per_ao",
              "nshells", "nao_per_atom", "atom_coords_internal", "bas_n"]:
        assert np.array_equal(np.asarray(getattr(a.basis, n)), np.asarray(getattr(b.basis, n))), n
    assert a.atom_coords == b.atom_coords and list(a.atoms) == list(b.atoms)
    assert (a.nelec, a.nup, a.ndown, a.atomic_number) == (b.nelec, b.nup, b.ndown, b.atomic_number)


@pytest.mark.parametrize("key", ["lih", "h2"])
def test_gto2sto_reproduces_reference_fit(key, double_default):
    """SlaterJastrow.gto2sto (slater_jastrow.py:649-733) is host set-up code: the fitted single-zeta Slater
    basis equals the reference's (golden made by oracle/make_golden.py gto2sto)."""
    from qmctorch_b200.wavefunction import SlaterJastrow
    g = np.load(os.path.join(C.GOLDEN, "gto2sto.npz"))
    wf = SlaterJastrow(fixture_molecule(key), configs="ground_state", cuda=False).gto2sto()
    b = wf.mol.basis
    assert b.radial_type == "sto_pure" and wf.ao.radial_type == "sto_pure"
    np.testing.assert_allclose(b.bas_exp, g[key + "_bas_exp"], rtol=1e-12)
    np.testing.assert_allclose(b.bas_norm, g[key + "_bas_norm"], rtol=1e-10)
    assert np.array_equal(np.stack([b.bas_kx, b.bas_ky, b.bas_kz, b.bas_kr]).astype(np.int64), g[key + "_bas_k"])
    assert np.array_equal(np.asarray(b.nshells), g[key + "_nshells"])
    np.testing.assert_allclose(wf.ao.norm_cst.numpy(), g[key + "_norm_cst"], rtol=1e-12)
    assert wf.ao.nbas == b.nao and not wf.ao.contract
    with pytest.raises(AssertionError):
        wf.gto2sto()                                   # already a Slater basis


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/hdf5"), reason="reference HDF5 files not on this box")
def test_wf_load_reads_reference_checkpoint_group():
    """WaveFunction.load (wf_base.py:257-277) reads <group>/models/<model> through the pure-Python reader and
    hands it to load_state_dict.  The reference's own tests/hdf5/LiH_adf_dz_QMCTorch.hdf5 predates the current
    parameter names (mo.weight, jastrow.weight), so - exactly like the reference - strict loading reports them."""
    from qmctorch_b200.utils.hdf5_min import read_hdf5
    from qmctorch_b200.wavefunction import SlaterJastrow
    path = "/root/reference/tests/hdf5/LiH_adf_dz_QMCTorch.hdf5"
    best = read_hdf5(path)["wf_opt"]["models"]["best"]
    assert best["ao.bas_exp"].shape == (9,) and best["fc.weight"].shape == (1, 4)
    wf = SlaterJastrow(fixture_molecule("lih_adf"), configs="single_double(2,2)")
    assert np.array_equal(best["ao.atom_coords"], wf.ao.atom_coords.detach().numpy())
    with pytest.raises(RuntimeError, match="mo.mo_modifier"):
        wf.load(path)


def test_hdf5_reader_rejects_other_files(tmp_path):
    from qmctorch_b200.utils.hdf5_min import read_hdf5
    p = tmp_path / "x.hdf5"
    p.write_bytes(b"not an hdf5 file at all, just bytes" * 4)
    with pytest.raises(ValueError):
        read_hdf5(str(p))


def test_shard_walkers_partitions_everything():
    from qmctorch_b200.solver.distributed import shard_walkers
    for n, w in ((10, 3), (1000000, 8), (7, 8), (0, 2)):
        spans = [shard_walkers(n, r, w) for r in range(w)]
        assert sum(c for _, c in spans) == n
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from qmctorch_b200.solver import distributed as D
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = D.world()
torch.manual_seed(0)
eloc = torch.randn(1001, dtype=torch.float64) - 7.9          # same on every rank
first, count = D.shard_walkers(len(eloc), rank, world)
mine = eloc[first:first + count]
mean, var, err, n, bad = D.global_stats(mine.sum(), (mine ** 2).sum(), float(count))
assert n == len(eloc) and bad == 0
assert abs(mean - float(eloc.mean())) < 1e-12 and abs(var - float(eloc.var())) < 1e-10
# gradient all-reduce: each rank holds a partial sum
p = torch.nn.Parameter(torch.zeros(5, dtype=torch.float64)); q = torch.nn.Parameter(torch.zeros(2, 3, dtype=torch.float64))
p.grad = torch.full((5,), float(rank + 1), dtype=torch.float64); q.grad = torch.full((2, 3), 10.0 * (rank + 1), dtype=torch.float64)
D.allreduce_gradients([p, q])
tot = sum(range(1, world + 1))
assert torch.equal(p.grad, torch.full((5,), float(tot), dtype=torch.float64)) and torch.equal(q.grad, torch.full((2, 3), 10.0 * tot, dtype=torch.float64))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_distributed_collectives_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out
        assert "ok" in out
