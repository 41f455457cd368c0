"""WaveFunction base: the method surface of qmctorch/wavefunction/wf_base.py:7-277."""
import torch


class WaveFunction(torch.nn.Module):
    def __init__(self, nelec, ndim, kinetic="auto", cuda=False):
        super().__init__()
        self.ndim = ndim
        self.nelec = nelec
        self.ndim_tot = self.nelec * self.ndim
        self.kinetic = kinetic
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")

    def forward(self, x):
        raise NotImplementedError()

    def local_energy(self, pos):
        raise NotImplementedError()

    def energy(self, pos):
        """wf_base.py:217-219."""
        return torch.mean(self.local_energy(pos))

    def variance(self, pos):
        """wf_base.py:221-223."""
        return torch.var(self.local_energy(pos))

    def sampling_error(self, eloc):
        """wf_base.py:225-229."""
        return torch.sqrt(eloc.var() / eloc.shape[0])

    def _energy_variance(self, pos):
        el = self.local_energy(pos)
        return torch.mean(el), torch.var(el)

    def _energy_variance_error(self, pos):
        el = self.local_energy(pos)
        return torch.mean(el), torch.var(el), self.sampling_error(el)

    def pdf(self, pos, return_grad=False):
        """wf_base.py:241-247."""
        if return_grad:
            return self.gradients(pos, pdf=True)
        return (self.forward(pos) ** 2).reshape(-1)

    def get_number_parameters(self):
        """wf_base.py:249-255."""
        return sum(p.data.numel() for p in self.parameters() if p.requires_grad)

    def load(self, filename, group="wf_opt", model="best"):
        """Load trained parameters from ``<group>/models/<model>`` of a QMCTorch HDF5 file
        (wf_base.py:257-277); read with the pure-Python reader ``utils/hdf5_min.py`` (no h5py).
        As in the reference, the stored names must be the ``state_dict`` names of this object."""
        from ..utils.hdf5_min import read_hdf5
        grp = read_hdf5(filename)[group]["models"][model]
        self.load_state_dict({name: torch.as_tensor(val) for name, val in grp.items()})
