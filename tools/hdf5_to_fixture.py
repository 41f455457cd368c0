"""Convert the `molecule` group of a QMCTorch HDF5 file (e.g. the reference's
tests/hdf5/LiH_adf_dz.hdf5, written by `Molecule(calculator='adf')`) into the JSON dump that
`qmctorch_b200.molecules.Molecule(load=...)` also accepts.  h5py is not needed: the file is read by
qmctorch_b200/utils/hdf5_min.py.

    python tools/hdf5_to_fixture.py /root/reference/tests/hdf5/LiH_adf_dz.hdf5 tests/data/LiH_adf_dz.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qmctorch_b200.molecules import _tree_to_json  # noqa: E402
from qmctorch_b200.utils.hdf5_min import read_hdf5  # noqa: E402

if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    tree = read_hdf5(src)["molecule"]
    tree.get("calculator", {}).pop("additional_basis_path", None)   # a path of the machine that ran ADF
    with open(dst, "w") as f:
        json.dump(_tree_to_json(tree), f, indent=0)
    print(dst, os.path.getsize(dst), "bytes")
