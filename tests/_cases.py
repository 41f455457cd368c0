"""Shared helpers: golden cases -> oracle parameters and qmctorch_b200 wave functions."""
import os

import numpy as np
import torch

import sj_oracle as orc
from qmctorch_b200.molecules import fixture_molecule

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["h2_single22", "h2_ground", "lih_ground", "lih_nojastrow", "lih_sd22", "lih_cas24", "lih_een",
         "h2o_ground", "h2o_cas44", "c4h6_ground", "lih_sd22_een3", "h2o_cas44_een",
         "lih_sto", "lih_sto_pure", "lih_gto_kr",
         # real ADF SCF results (Slater basis, MOs of the SCF) read from the reference's HDF5 files
         "lih_adf_sd22", "co2_adf_ground"]


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    meta = [str(x) for x in g["meta"]]
    g["name"], g["key"], g["configs"], g["jastrow"], g["step"] = meta[0], meta[1], meta[2], meta[3], float(meta[4])
    return g


def oracle_params(g):
    mol = fixture_molecule(g["key"])
    jw = None if g["jastrow"] == "None" else float(g["jw"][0])
    enw = float(g["enw"][0]) if "+en" in g["jastrow"].replace("+een", "") else None
    P = orc.make_params(mol, (g["cfg_up"], g["cfg_down"]), jastrow_weight=jw, en_weight=enw)
    P.mo_modifier = torch.tensor(g["mo_modifier"])
    P.ci = torch.tensor(g["ci"])
    if g["jastrow"].endswith("een"):
        P.een = dict(num=torch.tensor(g["bh_num"]), denom=torch.tensor(g["bh_denom"]), fc=torch.tensor(g["bh_fc"]))
    return mol, P


def build_wf(g, cuda=True):
    """qmctorch_b200 SlaterJastrow with the golden case's parameters."""
    from qmctorch_b200.wavefunction import SlaterJastrow
    from qmctorch_b200.wavefunction.jastrows.elec_elec import JastrowFactor as JEE, PadeJastrowKernel as PEE
    from qmctorch_b200.wavefunction.jastrows.elec_nuclei import JastrowFactor as JEN, PadeJastrowKernel as PEN
    from qmctorch_b200.wavefunction.jastrows.elec_elec_nuclei import JastrowFactor as JEEN, BoysHandyJastrowKernel
    mol = fixture_molecule(g["key"])
    jt = g["jastrow"]
    has_en = "+en" in jt.replace("+een", "")
    has_een = jt.endswith("een")
    if jt == "None":
        j = None
    elif jt == "ee":
        j = "default"
    else:
        j = [JEE(mol, PEE, cuda=cuda)]
        if has_en:
            j.append(JEN(mol, PEN, cuda=cuda))
        if has_een:
            j.append(JEEN(mol, BoysHandyJastrowKernel, cuda=cuda))
    wf = SlaterJastrow(mol, configs=g["configs"], jastrow=j, cuda=cuda)
    with torch.no_grad():
        wf.mo.mo_modifier.copy_(torch.tensor(g["mo_modifier"]))
        wf.fc.weight.copy_(torch.tensor(g["ci"]))
        if g["jastrow"] == "ee":
            wf.jastrow.jastrow_kernel.weight.fill_(float(g["jw"][0]))
        elif jt != "None":
            wf.jastrow.jastrow_terms[0].jastrow_kernel.weight.fill_(float(g["jw"][0]))
            if has_en:
                wf.jastrow.jastrow_terms[1].jastrow_kernel.weight.fill_(float(g["enw"][0]))
            if has_een:
                bh = wf.jastrow.jastrow_terms[-1].jastrow_kernel
                bh.weight_num.copy_(torch.tensor(g["bh_num"]))
                bh.weight_denom.copy_(torch.tensor(g["bh_denom"]))
                bh.fc.weight.copy_(torch.tensor(g["bh_fc"]))
    return mol, wf


def rel_err(a, b):
    """max_w |a-b| / |b| elementwise (per-walker relative error)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float(((a - b).abs() / b.abs().clamp(min=1e-300)).max())


def scaled_err(a, b):
    """max |a-b| / max |b| (for arrays with structural zeros)."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-300))
