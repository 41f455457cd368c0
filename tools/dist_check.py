"""Multi-GPU check of the Solver path (torchrun, one rank per GPU, NCCL): walkers are sharded, the
energy statistics and the parameter gradients are all-reduced; after optimisation steps every rank
must hold bit-identical parameters, and the energy must agree with a single-rank run of the same
total ensemble within the statistical error.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qmctorch_b200 import set_torch_double_precision
from qmctorch_b200.scf import fixture_molecule
from qmctorch_b200.sampler import Metropolis
from qmctorch_b200.solver import Solver
from qmctorch_b200.solver import distributed as D
from qmctorch_b200.wavefunction import SlaterJastrow

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
set_torch_double_precision()
mol = fixture_molecule("lih")
total = 400_000
first, count = D.shard_walkers(total, rank, world)
wf = SlaterJastrow(mol, configs="single_double(2,2)", cuda=True)
sampler = Metropolis(nwalkers=count, nstep=150, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                     move={"type": "all-elec", "proba": "normal"}, cuda=True, seed=100 + rank)
torch.manual_seed(7 + rank)
solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.Adam(wf.parameters(), lr=5e-3), rank=rank)
solver.configure(track=["local_energy"], freeze=["ao"], loss="energy", grad="manual",
                 resampling={"mode": "update", "resample_every": 1, "nstep_update": 20, "ntherm_update": -1})
obs = solver.single_point(with_tqdm=False)
e0, err0 = float(obs.energy), float(obs.error)
solver.run(3)
flat = torch.cat([p.detach().reshape(-1) for p in wf.parameters()])
same = True
if world > 1:
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, flat))
    flags = torch.tensor([float(same)], device=flat.device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    same = bool(flags.item())
obs = solver.single_point(with_tqdm=False)
if rank == 0:
    print("world=%d  walkers/rank=%d  E0 = %.5f +- %.5f  after 3 epochs E = %.5f +- %.5f  parameters identical on all ranks: %s  "
          "specialised=%d" % (world, count, e0, err0, float(obs.energy), float(obs.error), same, wf._handle.info(13)), flush=True)
    assert same
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
