"""CI configuration tables (host side).  Emits the same index tensors, in the same order,
as qmctorch/wavefunction/pooling/orbital_configurations.py:14-214 so that ``fc.weight``
checkpoints interchange."""
import re
from itertools import combinations, product

import torch


class OrbitalConfigurations:
    def __init__(self, mol):
        self.nup = mol.nup
        self.ndown = mol.ndown
        self.nelec = self.nup + self.ndown
        self.spin = mol.spin
        self.norb = mol.basis.nmo

    def get_configs(self, configs):
        if isinstance(configs, tuple):
            assert len(configs) == 2
            assert configs[0].shape == configs[1].shape
            assert len(configs[0][0]) == self.nup
            assert len(configs[0][0]) == self.ndown
            return configs
        if not isinstance(configs, str):
            raise ValueError("Config error")
        spec = configs.lower().replace(" ", "")
        if spec == "ground_state":
            return self._tensors([self._gs()])
        m = re.fullmatch(r"(cas|single|single_double)\((\d+),(\d+)\)", spec)
        if m is None:
            print(configs, " not recognized as valid configuration")
            print("Options are : ground_state, single(nelec,norb), single_double(nelec,norb), "
                  "cas(nelec,norb), tuple(tensor,tensor)")
            raise ValueError("Config error")
        kind, nelec, norb = m.group(1), int(m.group(2)), int(m.group(3))
        if nelec > self.nelec:
            raise ValueError("required number of electron in config too large")
        if norb > self.norb:
            raise ValueError("required number of orbitals in config too large")
        nocc = (nelec // 2 + nelec % 2, nelec // 2)
        nvirt = (norb - nocc[0], norb - nocc[1])
        if kind == "cas":
            return self._tensors(self._cas(nocc, nvirt, nelec))
        pairs = self._singles(nocc, nvirt)
        if kind == "single_double":
            pairs += self._doubles(nocc, nvirt)
        return self._tensors(pairs)

    # -- helpers: a "pair" is (occupied up list, occupied down list)
    def _gs(self):
        return list(range(self.nup)), list(range(self.ndown))

    @staticmethod
    def _tensors(pairs):
        return (torch.LongTensor([p[0] for p in pairs]), torch.LongTensor([p[1] for p in pairs]))

    @staticmethod
    def _excite(conf, iocc, ivirt):
        conf = list(conf)
        conf[iocc] = ivirt          # replace in place, no re-ordering (:303-313)
        return conf

    def _active(self, nocc, nvirt):
        occ_up = list(range(self.nup - 1, self.nup - 1 - nocc[0], -1))
        vrt_up = list(range(self.nup, self.nup + nvirt[0]))
        occ_dn = list(range(self.ndown - 1, self.ndown - 1 - nocc[1], -1))
        vrt_dn = list(range(self.ndown, self.ndown + nvirt[1]))
        return occ_up, vrt_up, occ_dn, vrt_dn

    def _singles(self, nocc, nvirt):
        gu, gd = self._gs()
        occ_up, vrt_up, occ_dn, vrt_dn = self._active(nocc, nvirt)
        out = [(gu, gd)]
        out += [(self._excite(gu, o, v), gd) for o in occ_up for v in vrt_up]
        out += [(gu, self._excite(gd, o, v)) for o in occ_dn for v in vrt_dn]
        return out

    def _doubles(self, nocc, nvirt):
        gu, gd = self._gs()
        occ_up, vrt_up, occ_dn, vrt_dn = self._active(nocc, nvirt)
        out = []
        for ou in occ_up:
            for vu in vrt_up:
                for od in occ_dn:
                    for vd in vrt_dn:
                        out.append((self._excite(gu, ou, vu), self._excite(gd, od, vd)))
        for o1, o2 in combinations(occ_up, 2):
            for v1, v2 in combinations(vrt_up, 2):
                out.append((self._excite(self._excite(gu, o1, v2), o2, v1), gd))
        for o1, o2 in combinations(occ_dn, 2):
            for v1, v2 in combinations(vrt_dn, 2):
                out.append((gu, self._excite(self._excite(gd, o1, v2), o2, v1)))
        return out

    def _cas(self, nocc, nvirt, nelec):
        if self.spin != 0:
            raise ValueError("CAS active space not possible with spin polarized calculation")
        lo, hi = self.nup - nocc[0], self.nup + nvirt[0]
        frozen = list(range(lo))
        occs = [frozen + list(c) for c in combinations(range(lo, hi), nelec // 2)]
        return list(product(occs, occs))
