"""Worker of tests/test_gpu_dist.py (one process per GPU under torchrun, NCCL).

Checks, for BASELINE config 3 (LiH optimisation, Jastrow + MO coefficients + CI):
  1. the all-reduced manual energy gradient of the sharded ensemble equals the oracle's gradient of
     the WHOLE ensemble (1e-10 relative, north_star bar);
  2. mini-batches: gradients accumulate locally and are summed over ranks once per epoch
     (two batches give the same epoch gradient as the oracle's sum over the batches);
  3. after optimisation epochs every rank holds bit-identical parameters;
  4. shards drawn from the same torch.manual_seed are different (rank-folded seeds).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from qmctorch_b200 import set_torch_double_precision  # noqa: E402
from qmctorch_b200.molecules import fixture_molecule  # noqa: E402
from qmctorch_b200.sampler import Metropolis  # noqa: E402
from qmctorch_b200.solver import Solver  # noqa: E402
from qmctorch_b200.wavefunction import SlaterJastrow  # noqa: E402
import sj_oracle as orc  # noqa: E402  (checker only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
set_torch_double_precision()
mol = fixture_molecule("lih")
W = 768
wf = SlaterJastrow(mol, configs="single_double(2,2)", cuda=True)
with torch.no_grad():
    g = torch.Generator().manual_seed(3)
    wf.fc.weight.add_((0.1 * torch.rand(wf.fc.weight.shape, generator=g, dtype=torch.float64)).to(dev))
torch.manual_seed(11)                      # the SAME seed on every rank, on purpose (check 4)
sampler = Metropolis(nwalkers=W, nstep=60, step_size=0.3, nelec=wf.nelec, ndim=3, init=mol.domain("normal"),
                     move={"type": "all-elec", "proba": "normal"}, cuda=True)
solver = Solver(wf=wf, sampler=sampler, optimizer=torch.optim.SGD(wf.parameters(), lr=1e-3), rank=rank)
solver.configure(track=["local_energy"], freeze=["ao"], loss="energy", grad="manual",
                 resampling={"mode": "update", "resample_every": 1, "nstep_update": 10, "ntherm_update": -1})
pos = sampler(wf.pdf, with_tqdm=False).detach().to(dev)
allpos = [torch.empty_like(pos) for _ in range(world)]
dist.all_gather(allpos, pos)
assert not torch.equal(allpos[0], allpos[1]), "shards are identical: seeds are not folded with the rank"
whole = torch.cat(allpos).cpu()


def oracle_grads(batches):
    cfg = (wf.configs[0].cpu(), wf.configs[1].cpu())
    P = orc.make_params(mol, cfg, jastrow_weight=float(wf.jastrow.jastrow_kernel.weight))
    P.ci = wf.fc.weight.detach().cpu().clone()
    P.mo_modifier = wf.mo.mo_modifier.detach().cpu().clone()
    tot = None
    for b in batches:
        gr, _ = orc.param_grads(P, b, names=("jastrow_weight", "mo_modifier", "ci"))
        tot = gr if tot is None else {k: tot[k] + gr[k] for k in gr}
    return tot


def mine():
    return {"jastrow_weight": wf.jastrow.jastrow_kernel.weight.grad.detach().cpu(),
            "mo_modifier": wf.mo.mo_modifier.grad.detach().cpu(), "ci": wf.fc.weight.grad.detach().cpu()}


def check(tag, got, ref):
    for k in ref:
        err = float((got[k] - ref[k]).abs().max() / ref[k].abs().max().clamp(min=1e-300))
        assert err < 1e-10, "%s: gradient %s differs from the oracle by %.2e" % (tag, k, err)


# 1. one batch, direct call: all-reduced inside
wf.zero_grad()
solver.evaluate_grad_manual(pos)
check("full batch", mine(), oracle_grads([whole]))
# 2. two mini-batches per rank through run_epochs' accumulation rule (one all-reduce per epoch).
#    Each (rank, batch) pair uses the GLOBAL mean of its batch round, as solver.py:418-421 does on
#    the concatenated batch: the oracle sees round b = the b-th halves of every shard.
from qmctorch_b200.solver import distributed as D  # noqa: E402
wf.zero_grad()
half = W // 2
for b in range(2):
    solver.evaluate_grad_manual(pos[b * half:(b + 1) * half], allreduce=False)
D.allreduce_gradients(solver._trainable())
rounds = [torch.cat([p[b * half:(b + 1) * half] for p in allpos]).cpu() for b in range(2)]
check("two batches", mine(), oracle_grads(rounds))
# 3. epochs with batches -> identical parameters everywhere
solver.run(2, batchsize=half)
flat = torch.cat([p.detach().reshape(-1) for p in wf.parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
flag = torch.tensor([float(torch.equal(ref, flat))], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
assert bool(flag.item()), "parameters differ between ranks after optimisation"
obs = solver.single_point(with_tqdm=False)
if rank == 0:
    print("DIST_OK world=%d E=%.5f +- %.5f specialised=%d" % (world, float(obs.energy), float(obs.error),
                                                            wf._handle.info(13)), flush=True)
dist.barrier()
dist.destroy_process_group()
