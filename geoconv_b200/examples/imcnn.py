"""Drop-in for the FAUST demo network ``geoconv_examples.mpi_faust.pytorch.model``
(/root/reference/src/geoconv_examples/mpi_faust/pytorch/model.py:13-198): the realistic caller of the hot path
 - Normalization -> Linear 544->64 + ReLU + BatchNorm -> [Dropout -> ISC -> AngularMaxPooling ->
BatchNorm] x 4 -> Linear.  Same constructor, same sub-module names (state_dict compatible), same loops; the ISC
and pooling layers are the geoconv_b200 ones, everything else is stock torch.  ``multiclass_accuracy`` replaces
the ``torcheval`` import of the reference with the same definition (micro average of argmax hits).

With ``dataset`` = ``geoconv_b200.data.faust_data_set.DeviceFaustCache`` a training epoch touches the host only
for the Python loop itself.
"""
import sys

import torch
from torch import nn

from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero


def multiclass_accuracy(pred, target):
    """torcheval.metrics.functional.multiclass_accuracy with its defaults (micro average)."""
    return (pred.argmax(dim=1) == target).float().mean()


def custom_exp_scheduler(opt, step, decay_rate=0.95, decay_steps=500):
    """Smooth exponential scheduler (model.py:13-15)."""
    opt.param_groups[0]["lr"] = opt.param_groups[0]["initial_lr"] * decay_rate ** (step / decay_steps)


def print_mem():
    mem = torch.cuda.memory_allocated()
    max_mem = torch.cuda.max_memory_allocated()
    return f"{mem / 1024 ** 2:.3f} MB / Max memory: {max_mem / 1024 ** 2:.3f} MB"


class Normalization(nn.Module):
    """Per-feature statistics of a signal-only dataset, applied as (x - mean) / var (sic, model.py:24-44).
    ``dataset`` iterates signal matrices and has ``reset()``."""

    def __init__(self, dataset):
        super().__init__()
        mean, var, n_samples = 0, 0, 0
        for s in dataset:
            n_samples += s.shape[0]
            mean = mean + torch.sum(s, dim=-2)
        mean = mean / n_samples
        dataset.reset()
        for s in dataset:
            var = var + torch.sum((s - mean) ** 2, dim=-2)
        var = var / (n_samples - 1)
        # plain attributes in the reference (not moved by .to()); buffers here so the module follows its device
        self.register_buffer("mean", torch.as_tensor(mean), persistent=False)
        self.register_buffer("var", torch.as_tensor(var), persistent=False)

    def forward(self, inputs):
        return (inputs - self.mean) / self.var


class Imcnn(nn.Module):
    """model.py:47-137."""

    def __init__(self, signal_dim, kernel_size, template_radius, adapt_data, layer_conf=None, variant="dirac",
                 segmentation_classes=-1, precision=None):
        super().__init__()
        self.signal_dim = signal_dim
        self.kernel_size = kernel_size
        self.template_radius = template_radius
        layer_types = {"dirac": ConvDirac, "geodesic": ConvGeodesic, "zero": ConvZero}
        if variant not in layer_types:
            raise RuntimeError("Select a layer type from: ['dirac', 'geodesic', 'zero']")
        self.layer_type = layer_types[variant]
        if layer_conf is None:
            self.output_dims = [96, 256, 384, 384]
            self.rotation_deltas = [1 for _ in range(len(self.output_dims))]
        else:
            self.output_dims, self.rotation_deltas = list(zip(*layer_conf))
        self.downsize_dim = 64
        self.input_dims = [self.downsize_dim]
        self.input_dims.extend(self.output_dims)

        self.normalize = Normalization(adapt_data)
        self.downsize_dense = nn.Linear(in_features=signal_dim, out_features=self.downsize_dim)
        self.downsize_activation = nn.ReLU()
        self.downsize_bn = nn.BatchNorm1d(num_features=self.downsize_dim)

        self.do_layers = nn.ModuleList()
        self.isc_layers = nn.ModuleList()
        self.bn_layers = nn.ModuleList()
        self.amp_layers = nn.ModuleList()
        for idx in range(len(self.output_dims)):
            self.do_layers.append(nn.Dropout(p=0.2))
            self.isc_layers.append(
                self.layer_type(
                    input_shape=[(None, self.input_dims[idx]), (None, kernel_size[0], kernel_size[1], 3, 2)],
                    amt_templates=self.output_dims[idx], template_radius=self.template_radius, activation="relu",
                    rotation_delta=self.rotation_deltas[idx], precision=precision))
            self.bn_layers.append(nn.BatchNorm1d(num_features=self.output_dims[idx]))
            self.amp_layers.append(AngularMaxPooling())

        if segmentation_classes:
            self.output_dense = nn.Linear(in_features=self.output_dims[-1], out_features=segmentation_classes)
        else:
            self.output_dense = nn.Linear(in_features=self.output_dims[-1], out_features=6890)

    def forward(self, inputs):
        signal, bc = inputs
        signal = self.normalize(signal)
        signal = self.downsize_dense(signal)
        signal = self.downsize_activation(signal)
        signal = self.downsize_bn(signal)
        for idx in range(len(self.output_dims)):
            signal = self.do_layers[idx](signal)
            signal = self.isc_layers[idx]([signal, bc])
            signal = self.amp_layers[idx](signal)
            signal = self.bn_layers[idx](signal)
        return self.output_dense(signal)

    def train_loop(self, dataset, loss_fn, opt, decay_rate=0.95, decay_steps=500, verbose=True, epoch=None,
                   prev_steps=None, use_lr_decay=False):
        """model.py:139-183."""
        self.train()
        epoch_accuracy, epoch_loss, mean_accuracy, mean_loss = 0., 0., 0., 0.
        for step, ((signal, bc), gt) in enumerate(dataset):
            pred = self([signal, bc])
            loss = loss_fn(pred, gt)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if use_lr_decay:
                custom_exp_scheduler(opt, prev_steps + step, decay_rate=decay_rate, decay_steps=decay_steps)
            epoch_accuracy = epoch_accuracy + multiclass_accuracy(pred, gt).detach()
            epoch_loss = epoch_loss + loss.detach()
            mean_accuracy = epoch_accuracy / (step + 1)
            mean_loss = epoch_loss / (step + 1)
            if verbose:
                sys.stdout.write(f"\rEpoch: {epoch} - Training step: {step} - Loss {mean_loss:.4f} - "
                                 f"Accuracy {mean_accuracy:.4f} - Memory: {print_mem()}")
        return {"epoch_loss": mean_loss, "epoch_accuracy": mean_accuracy}

    def validation_loop(self, dataset, loss_fn, verbose=True):
        """model.py:185-210."""
        self.eval()
        with torch.no_grad():
            val_loss, val_accuracy, step = 0., 0., 0
            for step, ((signal, bc), gt) in enumerate(dataset):
                pred = self([signal, bc])
                val_accuracy = val_accuracy + multiclass_accuracy(pred, gt).detach()
                val_loss = val_loss + loss_fn(pred, gt).detach()
            mean_accuracy = val_accuracy / (step + 1)
            mean_loss = val_loss / (step + 1)
            if verbose:
                sys.stdout.write(f" - Val.-Loss: {mean_loss:.4f} - Val.-Accuracy: {mean_accuracy:.4f}")
        return {"val_epoch_loss": mean_loss, "val_epoch_accuracy": mean_accuracy}
