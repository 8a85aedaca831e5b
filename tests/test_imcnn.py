"""The realistic caller of the path: the FAUST demo network on the drop-in layers."""
import pytest
import torch

from geoconv_b200 import synthetic as S


class _Signals:
    """Signal-only dataset with reset(), what Normalization consumes (faust_data_set.py only_signal=True)."""

    def __init__(self, signals):
        self.signals = signals
        self.reset()

    def reset(self):
        self._it = iter(self.signals)

    def __iter__(self):
        return self._it


def test_module_names_match_the_reference():
    from geoconv_b200.examples.imcnn import Imcnn
    sig = [torch.rand(20, 12) for _ in range(3)]
    net = Imcnn(signal_dim=12, kernel_size=(2, 4), template_radius=0.03, adapt_data=_Signals(sig),
                layer_conf=[(8, 1), (16, 2)], variant="geodesic", segmentation_classes=5)
    keys = list(net.state_dict().keys())
    assert keys[:2] == ["downsize_dense.weight", "downsize_dense.bias"]
    assert "isc_layers.0._template_neighbor_weights" in keys and "isc_layers.1._bias" in keys
    assert "bn_layers.1.running_var" in keys and keys[-2:] == ["output_dense.weight", "output_dense.bias"]
    assert net.isc_layers[1].rotation_delta == 2 and net.input_dims == [64, 8, 16]
    with pytest.raises(RuntimeError):
        Imcnn(12, (2, 4), 0.03, _Signals(sig), variant="nope")


@pytest.mark.gpu
def test_training_reduces_the_loss():
    from geoconv_b200.examples.imcnn import Imcnn
    dev = "cuda:0"
    torch.manual_seed(0)
    v, f, r, a, classes = 400, 32, 3, 6, 7
    meshes = []
    for m in range(4):
        signal = S.make_signal(v, f, seed=10 + m, kind="randn").to(dev)
        bary = S.make_bary(v, r, a, seed=10 + m).to(dev)
        gt = (signal[:, :classes].argmax(dim=1)).to(dev)        # a learnable per-vertex label
        meshes.append(((signal, bary), gt))
    net = Imcnn(signal_dim=f, kernel_size=(r, a), template_radius=0.03, adapt_data=_Signals([m[0][0] for m in meshes]),
                layer_conf=[(32, 1), (32, 1)], variant="geodesic", segmentation_classes=classes).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    loss_fn = torch.nn.CrossEntropyLoss()
    first = net.train_loop(meshes, loss_fn, opt, verbose=False)["epoch_loss"].item()
    for _ in range(14):
        last = net.train_loop(meshes, loss_fn, opt, verbose=False)
    assert last["epoch_loss"].item() < 0.6 * first
    val = net.validation_loop(meshes, loss_fn, verbose=False)
    assert val["val_epoch_accuracy"].item() > 1.5 / classes
