"""Product of Jastrow factors (qmctorch/wavefunction/jastrows/combine_jastrow.py:8-195).
The fused kernel forms the product rule from the per-term log-derivatives
(grad ln J = sum of terms, lap J / J = sum h_t + |sum g_t|^2)."""
from torch import nn

from ._base import jastrow_forward
from .elec_elec import JastrowFactorElectronElectron
from .elec_nuclei import JastrowFactorElectronNuclei
from .elec_elec_nuclei import JastrowFactorElectronElectronNuclei


class CombineJastrow(nn.Module):
    def __init__(self, jastrow):
        super().__init__()
        self.jastrow_terms = nn.ModuleList()
        for j in jastrow:
            self.jastrow_terms.append(j)
        self.requires_autograd = True
        self.nterms = len(self.jastrow_terms)
        ee = [j for j in jastrow if isinstance(j, JastrowFactorElectronElectron)]
        en = [j for j in jastrow if isinstance(j, JastrowFactorElectronNuclei)]
        een = [j for j in jastrow if isinstance(j, JastrowFactorElectronElectronNuclei)]
        if len(ee) > 1 or len(en) > 1 or len(een) > 1 or len(ee) + len(en) + len(een) != len(jastrow):
            raise NotImplementedError(
                "the CUDA path combines at most one e-e Pade, one e-n Pade and one e-e-n Boys-Handy factor")
        self.__dict__["ee"] = ee[0] if ee else None      # aliases, not extra sub-modules
        self.__dict__["en"] = en[0] if en else None
        self.__dict__["een"] = een[0] if een else None
        self.nelec = jastrow[0].nelec
        self._handle = None

    def __repr__(self):
        return " + ".join(t.__repr__() for t in self.jastrow_terms)

    def _own_handle(self):
        if self._handle is None:
            from .._standalone import standalone_handle
            first = self.jastrow_terms[0]
            self._handle = standalone_handle(first._mol, self, jee=self.ee, jen=self.en, jeen=self.een)
        return self._handle

    def forward(self, pos, derivative=0, sum_grad=True):
        return jastrow_forward(self, self._own_handle(), 0, pos, derivative, sum_grad)
