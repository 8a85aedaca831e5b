"""Golden vectors for the FAUST archive reader from the EXECUTED reference:
/root/reference/src/geoconv_examples/mpi_faust/pytorch/faust_data_set.py::faust_generator / FaustDataset.

    oracle/build_ref.sh && python tests/golden/make_golden_faust_reader.py

The reader imports `get_file_number` from preprocess_faust.py, which pulls in pyshot, trimesh, pygeodesic and
matplotlib at module level (never called by the reader): they are replaced by empty stub modules, and the reference's
C extension comes from oracle/_ref/ (see make_golden_wavefront.py).  The archive is a seeded synthetic stand-in with
the layout preprocess_faust.py:153-210 writes (100 meshes: SIGNAL_i.npy, BC_i.npy float64, GT_i.npy, COORD_i.npy),
small enough to commit.  Stored: what the reference yields for every set type (train split with seeded `random` /
`numpy.random`: order and the |N(0, 1e-5)| noise on the weight column), dtypes included.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from make_golden_wavefront import install_stub_modules  # noqa: E402

V, F, R, A = 9, 6, 2, 3


def write_archive(path):
    rng = np.random.default_rng(17)
    items = {}
    for i in range(100):
        items[f"SIGNAL_tr_reg_{i:03d}.npy"] = rng.standard_normal((V, F))
        bc = np.zeros((V, R, A, 3, 2))
        bc[..., 0] = rng.integers(0, V, (V, R, A, 3))
        bc[..., 1] = rng.random((V, R, A, 3))
        items[f"BC_tr_reg_{i:03d}.npy"] = bc
        items[f"GT_tr_reg_{i:03d}.npy"] = rng.permutation(V)
        items[f"COORD_tr_reg_{i:03d}.npy"] = rng.standard_normal((V, 3))
    order = list(items)
    random.Random(3).shuffle(order)                 # archive order must not matter: the reader sorts by file number
    np.savez(path, **{k: items[k] for k in order})


if __name__ == "__main__":
    install_stub_modules()
    for name in ("pyshot",):
        import types
        sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "_ref"))
    sys.path.insert(0, "/root/reference/src")
    from geoconv_examples.mpi_faust.pytorch.faust_data_set import FaustDataset, faust_generator
    archive = os.path.join(HERE, "faust_reader_archive.npz")
    write_archive(archive)
    out = {}
    for set_type in (0, 1, 2, 3):
        random.seed(5)
        np.random.seed(6)
        got = list(faust_generator(archive, set_type=set_type))
        out[f"n_{set_type}"] = np.array(len(got))
        out[f"signal_{set_type}"] = np.stack([s.numpy() for (s, _), _ in got])
        out[f"bc_{set_type}"] = np.stack([b.numpy() for (_, b), _ in got])
        out[f"gt_{set_type}"] = np.stack([g.numpy() for _, g in got])
        out[f"dtypes_{set_type}"] = np.array([str(got[0][0][0].dtype), str(got[0][0][1].dtype), str(got[0][1].dtype)])
    random.seed(8)
    np.random.seed(9)
    got = list(faust_generator(archive, set_type=0, set_indices=[4, 99, 0], return_coordinates=True))
    out["coord_sel"] = np.stack([c.numpy() for (_, _, c), _ in got])
    out["bc_sel"] = np.stack([b.numpy() for (_, b, _), _ in got])
    out["signal_only"] = np.stack([s.numpy() for s in faust_generator(archive, set_type=1, only_signal=True)])
    ds = FaustDataset(archive, set_type=2)
    first = next(iter(ds))
    ds.reset()
    again = next(iter(ds))
    out["dataset_first_gt"] = first[1].numpy()
    assert torch.equal(first[0][0], again[0][0])
    np.savez_compressed(os.path.join(HERE, "faust_reader_expected.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items() if k.startswith(("bc_0", "signal_0", "dtypes_0", "gt_0"))})
