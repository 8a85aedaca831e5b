"""reader of the preprocessed FAUST archive (host generator + device-resident cache)."""
import os
import zipfile

import numpy as np
import pytest
import torch

from geoconv_b200.data.faust_data_set import DeviceFaustCache, FaustDataset, faust_generator, get_file_number


@pytest.fixture(scope="module")
def archive(tmp_path_factory):
    """100 tiny meshes in the layout preprocess_faust.py:153-210 writes (one .npy per kind and mesh, zipped)."""
    root = tmp_path_factory.mktemp("faust")
    rng = np.random.default_rng(0)
    path = os.path.join(root, "faust.zip")
    with zipfile.ZipFile(path, "w") as zf:
        for m in reversed(range(100)):                      # written out of order: the reader sorts by number
            v = 12 + m % 5
            sig = rng.random((v, 6))                         # float64 on disk, like np.save of SHOT descriptors
            bc = np.zeros((v, 2, 4, 3, 2))
            bc[..., 0] = rng.integers(0, v, (v, 2, 4, 3))
            bc[..., 1] = rng.random((v, 2, 4, 3))
            gt = rng.permutation(v)
            for kind, arr in (("SIGNAL", sig), ("BC", bc), ("GT", gt), ("COORD", rng.random((v, 3)))):
                tmp = os.path.join(root, f"{kind}_tr_reg_{m:03d}.npy")
                np.save(tmp, arr)
                zf.write(tmp, arcname=f"faust/{kind}_tr_reg_{m:03d}.npy")
                os.remove(tmp)
    return path


def test_file_number():
    assert get_file_number("SIGNAL_tr_reg_042.npy") == 42
    with pytest.raises(RuntimeError):
        get_file_number("SIGNAL_tr_reg.npy")


def test_splits_order_and_dtypes(archive):
    val = list(faust_generator(archive, set_type=1))
    assert len(val) == 10
    (signal, bc), gt = val[0]
    assert signal.dtype == torch.float32 and bc.dtype == torch.float32 and gt.dtype == torch.int64
    assert signal.shape[0] == 12 + 70 % 5 and gt.dim() == 1            # mesh 70 comes first: sorted by file number
    assert len(list(faust_generator(archive, set_type=2))) == 20
    assert len(list(faust_generator(archive, set_type=3))) == 100
    assert len(list(faust_generator(archive, set_type=3, set_indices=[5, 1]))) == 2
    with pytest.raises(RuntimeError):
        next(faust_generator(archive, set_type=7))
    (s, b, c), g = next(faust_generator(archive, set_type=1, return_coordinates=True))
    assert c.shape == (s.shape[0], 3)
    assert torch.equal(next(faust_generator(archive, set_type=1, only_signal=True)), signal)


def test_training_noise_only_on_weights(archive):
    clean = {tuple(bc.shape): bc for (_, bc), _ in faust_generator(archive, set_type=3, set_indices=[3])}
    np.random.seed(0)
    (_, noisy), _ = next(faust_generator(archive, set_type=0, set_indices=[3]))
    ref = clean[tuple(noisy.shape)]
    assert torch.equal(noisy[..., 0].float(), ref[..., 0])              # indices untouched
    d = noisy[..., 1].double() - ref[..., 1].double()
    assert (d >= 0).all() and d.max() < 1e-3 and d.mean() > 1e-6        # |N(0, 1e-5)|
    ds = FaustDataset(archive, set_type=0)
    assert len(list(ds)) == 70
    ds.reset()
    assert len(list(ds)) == 70


@pytest.mark.gpu
def test_device_cache_feeds_the_layer(archive):
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
    host = list(faust_generator(archive, set_type=1))
    cache = DeviceFaustCache(archive, set_type=1)
    assert len(cache) == 10 and cache.nbytes() > 0
    layer = ConvDirac(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=4, template_radius=0.03).cuda()
    n = 0
    for ((signal, bc), gt), ((hs, hb), hg) in zip(cache, host):     # same order: sorted by file number
        assert signal.is_cuda and bc.is_cuda and gt.is_cuda
        assert torch.equal(signal.cpu(), hs) and torch.equal(bc.cpu(), hb) and torch.equal(gt.cpu(), hg)
        z = AngularMaxPooling()(layer([signal, bc]))
        assert z.shape == (signal.shape[0], 4)
        n += 1
    assert n == 10
    train = DeviceFaustCache(archive, set_type=0, set_indices=[0, 1], generator=torch.Generator(device="cuda").manual_seed(1))
    (s1, b1), _ = next(iter(train))
    (s2, b2), _ = next(iter(train))
    assert b1 is not b2                                                  # a new tensor per epoch -> re-packed by the layer
    clean = train.meshes[0][1] if b1.shape == train.meshes[0][1].shape else train.meshes[1][1]
    assert torch.equal(b1[..., 0], clean[..., 0])
    d = (b1[..., 1] - clean[..., 1])
    assert (d >= 0).all() and d.max() < 1e-3


@pytest.mark.gpu
def test_device_cache_packs_every_mesh_once(archive):
    """the cache hands out tensors that carry their int32 indices and CSR inversion (computed when
    the mesh was loaded); a training step through them equals a step through a plain tensor with the same values."""
    from geoconv_b200 import ops
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    gen = torch.Generator(device="cuda").manual_seed(3)
    train = DeviceFaustCache(archive, set_type=0, set_indices=[0, 1, 2], generator=gen)
    layer = ConvGeodesic(input_shape=[(None, 6), (None, 2, 4, 3, 2)], amt_templates=4, template_radius=0.03).cuda()
    packs_seen = set()
    for epoch in range(2):
        for (signal, bc), _gt in train:
            pack, version = bc._gcb_pack
            assert version == bc._version and pack.keeps_all_slots
            assert torch.equal(pack.idx.long().cpu(), bc[..., 0].long().cpu())
            assert torch.equal(pack.w, bc[..., 1])
            assert ops.GLOBAL_BARY_CACHE.get(bc, signal.shape[0]) is pack          # the layers take it as is
            packs_seen.add(pack.csr()[1].data_ptr())                               # inversion shared across epochs
            grads = []
            for tensor in (bc, bc.clone()):                                        # packed by the cache / by the layer
                xs = signal.clone().requires_grad_(True)
                layer.zero_grad()
                AngularMaxPooling()(layer([xs, tensor])).sum().backward()
                grads.append((xs.grad.clone(), layer._template_neighbor_weights.grad.clone()))
            assert torch.allclose(grads[0][0], grads[1][0], rtol=1e-5, atol=1e-7)
            assert torch.allclose(grads[0][1], grads[1][1], rtol=1e-5, atol=1e-7)
    assert len(packs_seen) == 3                                                    # one inversion per mesh, not per epoch
    plain = DeviceFaustCache(archive, set_type=1, pack=False)
    (_s, bc), _ = next(iter(plain))
    assert not hasattr(bc, "_gcb_pack")


def test_reader_matches_the_executed_reference():
    """tests/golden/faust_reader_expected.npz holds what the reference's own faust_generator / FaustDataset yield on
    tests/golden/faust_reader_archive.npz (make_golden_faust_reader.py): same order, values, noise and dtypes, bit for
    bit - the reader draws from `random` / `numpy.random` exactly like the reference."""
    import os
    import random

    import numpy as np
    import torch

    from geoconv_b200.data.faust_data_set import FaustDataset, faust_generator
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    archive = os.path.join(here, "faust_reader_archive.npz")
    want = np.load(os.path.join(here, "faust_reader_expected.npz"))
    for set_type in (0, 1, 2, 3):
        random.seed(5)
        np.random.seed(6)
        got = list(faust_generator(archive, set_type=set_type))
        assert len(got) == int(want[f"n_{set_type}"]) == {0: 70, 1: 10, 2: 20, 3: 100}[set_type]
        assert [str(got[0][0][0].dtype), str(got[0][0][1].dtype), str(got[0][1].dtype)] == list(want[f"dtypes_{set_type}"])
        assert np.array_equal(np.stack([s.numpy() for (s, _), _ in got]), want[f"signal_{set_type}"])
        assert np.array_equal(np.stack([b.numpy() for (_, b), _ in got]), want[f"bc_{set_type}"])
        assert np.array_equal(np.stack([g.numpy() for _, g in got]), want[f"gt_{set_type}"])
    random.seed(8)
    np.random.seed(9)
    got = list(faust_generator(archive, set_type=0, set_indices=[4, 99, 0], return_coordinates=True))
    assert np.array_equal(np.stack([c.numpy() for (_, _, c), _ in got]), want["coord_sel"])
    assert np.array_equal(np.stack([b.numpy() for (_, b, _), _ in got]), want["bc_sel"])
    only = np.stack([s.numpy() for s in faust_generator(archive, set_type=1, only_signal=True)])
    assert np.array_equal(only, want["signal_only"])
    ds = FaustDataset(archive, set_type=2)
    first = next(iter(ds))
    assert np.array_equal(first[1].numpy(), want["dataset_first_gt"])
    ds.reset()
    assert torch.equal(next(iter(ds))[0][0], first[0][0])
