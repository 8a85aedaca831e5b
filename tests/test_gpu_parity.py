"""GPU parity: the CUDA path (through the drop-in layers, i.e. through the C ABI) against the
golden vectors of the executed reference and against the CPU oracle on seeded inputs.

Tolerances (normalised max error): fp32-faithful modes 1e-5, TF32 mode 2e-3.
Integer outputs (gather indices, AMP argmax) must be bit exact."""
import pytest
import torch

from oracle import geoconv_oracle as O
from geoconv_b200 import _native, ops
from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
from geoconv_b200.pytorch.layers.conv_dirac import ConvDirac
from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
from geoconv_b200.pytorch.layers.conv_zero import ConvZero
from tests.helpers import TOL_FP32, TOL_TF32, check_gradients, golden_cases, load_golden, nerr

pytestmark = pytest.mark.gpu
LAYERS = {"geodesic": ConvGeodesic, "dirac": ConvDirac, "zero": ConvZero}
DEV = "cuda:0"


def layer_from_golden(g, precision):
    v, r, a = g["bary"].shape[:3]
    f = g["signal"].shape[1]
    t = g["w_self"].shape[0]
    radius = g["template"][-1, 0, 0].item()
    layer = LAYERS[g["kind"]](input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t,
                              template_radius=radius, activation=g["act"], rotation_delta=g["delta"],
                              precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": g["w_neighbor"], "_template_self_weights": g["w_self"],
                           "_bias": g["bias"]})
    return layer.to(DEV)


def tol_for(layer, g, precision):
    r, a = g["bary"].shape[1:3]
    n_rot = len(range(0, a, g["delta"]))
    code = ops.resolve_precision(precision, r, a, g["signal"].shape[1], g["w_self"].shape[0], n_rot)
    return TOL_TF32 if code == _native.PREC_TF32 else TOL_FP32


def argmax_ok(arg, ref_arg, y_ref, tol):
    """Exact, except where the reference's best and runner-up norms are closer than `tol`
    (relative): there a different-but-equally-valid rotation may win."""
    arg = arg.cpu().long()
    ref_arg = ref_arg.long()
    bad = (arg != ref_arg).nonzero().flatten()
    norms = torch.linalg.vector_norm(y_ref.double(), dim=-1)
    for k in bad.tolist():
        n_ref, n_got = norms[k, ref_arg[k]], norms[k, arg[k]]
        if (n_ref - n_got).abs() > tol * n_ref.abs().clamp_min(1e-30):
            return False
    return True


@pytest.mark.parametrize("precision", ["ffma", "fp32", "tf32"])
@pytest.mark.parametrize("name", golden_cases())
def test_golden_forward_backward(name, precision):
    g = load_golden(name)
    layer = layer_from_golden(g, precision)
    tol = tol_for(layer, g, precision)
    amp = AngularMaxPooling()
    signal = g["signal"].to(DEV).requires_grad_(True)
    bary = g["bary"].to(DEV)
    y = layer([signal, bary])
    z = amp(y)
    assert y.shape == g["y"].shape and z.shape == g["z"].shape
    assert nerr(y, g["y"]) < tol
    arg = y._gcb_pooled[1]
    if tol == TOL_FP32:
        assert torch.equal(arg.cpu(), g["argmax"]) or argmax_ok(arg, g["argmax"], g["y"], 1e-6)
        assert nerr(z, g["z"]) < tol
    else:
        assert argmax_ok(arg, g["argmax"], g["y"], 4 * tol)
    z.backward(g["grad_z"].to(DEV))
    # (a near-tie rotation flip or a ReLU bit at a pre-activation within rounding of zero sends the gradients down
    # another, equally valid branch: check_gradients then verifies them in closed form at this forward's decisions)
    prior = g["prior"] if g["include_prior"] else None
    check_gradients(layer, signal.grad, z, arg, g, g["signal"], g["bary"], g["w_neighbor"], g["w_self"], prior,
                    g["grad_z"], g["act"], range(0, g["bary"].shape[2], g["delta"]), tol)
    assert layer._template_neighbor_weights.grad.shape == g["w_neighbor"].shape
    assert layer._template_self_weights.grad.shape == g["w_self"].shape
    assert layer._bias.grad.shape == g["bias"].shape


@pytest.mark.parametrize("name", ["geodesic_r5a8", "dirac_selu", "geodesic_a6_delta4"])
def test_golden_orientations_argument(name):
    g = load_golden(name)
    layer = layer_from_golden(g, "ffma")
    y = layer([g["signal"].to(DEV), g["bary"].to(DEV)], orientations=g["orientations"])
    assert nerr(y, g["y_orient"]) < TOL_FP32


@pytest.mark.parametrize("name", ["geodesic_r5a8", "dirac_r5a8", "zero_r5a8", "geodesic_delta2"])
def test_unfused_sequence_and_dense_backward(name):
    """conv -> (tensor op) -> stand-alone AMP kernel, and a loss taken on Y itself: the dense
    backward (one rotation at a time) must equal autograd over the oracle."""
    g = load_golden(name)
    layer = layer_from_golden(g, "ffma")
    signal = g["signal"].to(DEV).requires_grad_(True)
    bary = g["bary"].to(DEV)
    y = layer([signal, bary])
    y2 = y * 1.0                       # drops the fused handshake
    assert not hasattr(y2, "_gcb_pooled")
    z = AngularMaxPooling()(y2)
    assert nerr(z, g["z"]) < TOL_FP32
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
    (y * gy.to(DEV)).sum().backward()
    # oracle: same loss through autograd on CPU
    xs = g["signal"].clone().requires_grad_(True)
    ps = [g[k].clone().requires_grad_(True) for k in ("w_neighbor", "w_self", "bias")]
    prior = g["prior"] if g["include_prior"] else None
    yo = O.conv_forward(xs, g["bary"], ps[0], ps[1], ps[2], prior, g["act"], g["delta"])
    grads = torch.autograd.grad((yo * gy).sum(), [xs] + ps, allow_unused=True)
    assert nerr(signal.grad, grads[0]) < TOL_FP32
    ref_dw = grads[1] if grads[1] is not None else torch.zeros_like(ps[0])
    assert nerr(layer._template_neighbor_weights.grad, ref_dw) < TOL_FP32
    assert nerr(layer._template_self_weights.grad, grads[2]) < TOL_FP32
    assert nerr(layer._bias.grad, grads[3]) < TOL_FP32


def test_unfused_amp_backward_matches_fused():
    g = load_golden("geodesic_r5a8")
    grads = []
    for fused in (True, False):
        layer = layer_from_golden(g, "ffma")
        signal = g["signal"].to(DEV).requires_grad_(True)
        y = layer([signal, g["bary"].to(DEV)])
        z = AngularMaxPooling()(y if fused else y.clone())
        z.backward(g["grad_z"].to(DEV))
        grads.append((signal.grad, layer._template_neighbor_weights.grad, layer._bias.grad))
    for a, b in zip(*grads):
        assert nerr(a, b) < 1e-6


def test_prep_bary_and_csr_bit_exact():
    v, r, a = 301, 4, 6
    for dtype in (torch.float32, torch.float64):
        bary = O.make_bary(v, r, a, seed=3, fallback_frac=0.1, dtype=dtype)
        pack = ops.prep_bary(bary.to(DEV), v)
        idx, w = O.split_bary(bary)
        assert torch.equal(pack.idx.cpu().long(), idx)
        assert torch.equal(pack.w.cpu(), w)
        row_ptr, entries, _stats = pack.csr()
        row_ptr, entries = row_ptr.cpu().long(), entries.cpu().long()
        flat_idx, flat_w = idx.reshape(-1), w.reshape(-1)
        keep = (flat_w != 0).nonzero().flatten()
        assert row_ptr[0] == 0 and row_ptr[-1] == keep.numel()
        # rows grouped by source vertex, inside a row by the radial ring of the slot, inside a ring by slot id
        ring = (keep // 3) % (r * a) // a
        order = torch.argsort(flat_idx[keep] * r + ring, stable=True)
        assert torch.equal(entries[: keep.numel()], keep[order])       # stable => unique answer
        counts = torch.bincount(flat_idx[keep], minlength=v)
        assert torch.equal(row_ptr[1:] - row_ptr[:-1], counts)


def test_out_of_range_index_raises():
    """Index validation without a host sync per tensor: gcb_prep_bary counts bad indices on the device (and
    neutralises them); the count reaches the host behind an event and the next poll raises."""
    bary = O.make_bary(10, 2, 3, seed=1)
    bary[3, 1, 2, 0, 0] = 10.0
    with pytest.raises(IndexError):
        ops.prep_bary(bary.to(DEV), 10, check_indices="sync")
    pack = ops.prep_bary(bary.to(DEV), 10)            # default: deferred report
    with pytest.raises(IndexError):
        pack.validate()
    assert pack.oob_count == 1
    assert int(pack.idx[3, 1, 2, 0]) == 0 and float(pack.w[3, 1, 2, 0]) == 0.0   # never read out of bounds
    # through the layer: the call that sees the bad tensor does not block; a later call reports it
    layer = ConvDirac(input_shape=[(None, 8), (None, 2, 3, 3, 2)], amt_templates=16, template_radius=0.1).to(DEV)
    signal = torch.randn(10, 8, device=DEV)
    layer([signal, bary.to(DEV)])
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        layer([signal, O.make_bary(10, 2, 3, seed=2).to(DEV)])
    layer([signal, O.make_bary(10, 2, 3, seed=3).to(DEV)])   # reported once; good tensors keep working


def test_orientations_outside_0_A_backward():
    """`orientations=` may hold any integer (torch.roll semantics, conv_intrinsic.py:163): negative shifts and
    shifts >= A run through the fused-pooling backward and match autograd over the oracle."""
    v, f, r, a, t = 257, 32, 3, 6, 32
    wn, ws, b = O.make_weights(t, r, a, f, seed=11)
    signal = O.make_signal(v, f, seed=12, kind="randn")
    bary = O.make_bary(v, r, a, seed=12, fallback_frac=0.05)
    gz = torch.randn((v, t), generator=torch.Generator().manual_seed(13))
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.05))).float()
    orient = [-1, a + 2, 3, -a - 4]
    xs = signal.clone().requires_grad_(True)
    ps = [p_.clone().requires_grad_(True) for p_ in (wn, ws, b)]
    y_ref = O.conv_forward(xs, bary, ps[0], ps[1], ps[2], prior, "tanh", 1, orientations=orient)
    z_ref, arg_ref = O.amp_forward(y_ref)
    gr = torch.autograd.grad(z_ref, [xs] + ps, grad_outputs=gz)
    ref = {"z": z_ref.detach(), "argmax": arg_ref, "d_signal": gr[0], "d_w_neighbor": gr[1], "d_w_self": gr[2],
           "d_bias": gr[3]}
    for precision in ("ffma", "fp32", "tf32x3"):
        layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.05,
                             activation="tanh", precision=precision)
        layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
        layer = layer.to(DEV)
        xd = signal.to(DEV).requires_grad_(True)
        y = layer([xd, bary.to(DEV)], orientations=torch.tensor(orient))
        assert nerr(y, y_ref) < TOL_FP32
        z = AngularMaxPooling()(y)
        z.backward(gz.to(DEV))
        check_gradients(layer, xd.grad, z, y._gcb_pooled[1], ref, signal, bary, wn, ws, prior, gz, "tanh",
                        [o % a for o in orient], TOL_FP32)


def test_fold_prior_matches_oracle():
    t, r, a, f = 7, 3, 6, 10
    w = torch.randn(t, r, a, f)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.1))).float()
    wd = w.to(DEV).requires_grad_(True)
    out = ops.FoldPrior.apply(wd, prior.to(DEV))
    assert nerr(out, O.fold_prior(w.double(), prior.double())) < 1e-6
    gout = torch.randn(t, r, a, f)
    out.backward(gout.to(DEV))
    assert nerr(wd.grad, O.unfold_prior_grad(gout.double(), prior.double())) < 1e-6


def test_determinism_bitwise():
    """No atomics anywhere on the path: two runs give bit-identical outputs and gradients."""
    g = load_golden("geodesic_pool")
    outs = []
    for _ in range(2):
        layer = layer_from_golden(g, "fp32")
        signal = g["signal"].to(DEV).requires_grad_(True)
        z = AngularMaxPooling()(layer([signal, g["bary"].to(DEV)]))
        z.backward(g["grad_z"].to(DEV))
        outs.append((z.detach().clone(), signal.grad.clone(), layer._template_neighbor_weights.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_rotation_equivariance_property():
    """Size-independent property: rolling the barycentric tensor by s angular steps permutes the
    rotation axis of Y (Dirac prior, delta=1):  Y'[k,i] = Y[k,(i+s)%A]."""
    v, f, r, a, t = 500, 32, 3, 8, 16
    layer = ConvDirac(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.1,
                      precision="ffma").to(DEV)
    signal = O.make_signal(v, f, kind="randn").to(DEV)
    bary = O.make_bary(v, r, a, seed=4).to(DEV)
    y0 = layer([signal, bary])
    for s in (1, 3):
        ys = layer([signal, torch.roll(bary, shifts=s, dims=2).contiguous()])
        assert nerr(ys, torch.roll(y0, shifts=-s, dims=1)) < 1e-6


@pytest.mark.parametrize("precision", ["ffma", "fp32", "tf32"])
def test_faust_shape_against_oracle(precision):
    """BASELINE config S1: V=6890, 544->96, R5xA8, delta 1, relu, SHOT-shaped signal."""
    v, f, r, a, t = 6890, 544, 5, 8, 96
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="shot")
    bary = O.make_bary(v, r, a, seed=1)
    grad_z = torch.randn((v, t), generator=torch.Generator().manual_seed(2))
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028))).float()
    ref = O.conv_amp_forward_backward(signal, bary, wn, ws, b, prior, grad_z, "relu", 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028,
                         precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    code = ops.resolve_precision(precision, r, a, f, t, a)
    tol = TOL_TF32 if code == _native.PREC_TF32 else TOL_FP32
    xs = signal.to(DEV).requires_grad_(True)
    y = layer([xs, bary.to(DEV)])
    z = AngularMaxPooling()(y)
    arg = y._gcb_pooled[1].cpu()
    assert nerr(y, ref["y"]) < tol
    flips = (arg != ref["argmax"]).sum().item()
    assert argmax_ok(arg, ref["argmax"], ref["y"], 2e-6 if tol == TOL_FP32 else 4 * tol)
    if tol == TOL_TF32:
        # gradients of a TF32 forward legitimately differ where a ReLU mask or a rotation flips;
        # check the TF32 backward arithmetic behind an exact forward instead
        layer.precision, layer.backward_precision = "ffma", precision
        for prm in layer.parameters():
            prm.grad = None
        xs = signal.to(DEV).requires_grad_(True)
        y = layer([xs, bary.to(DEV)])
        z = AngularMaxPooling()(y)
        flips = (y._gcb_pooled[1].cpu() != ref["argmax"]).sum().item()
    z.backward(grad_z.to(DEV))
    # A ReLU derivative bit may differ from the reference's where the pre-activation is within rounding of
    # zero (|z| and |z_ref| below 2e-6 of the largest output): admissible, like a rotation flip at a near tie.
    zc, zr = z.detach().cpu(), ref["z"]
    mask_diff = (zc > 0) != (zr > 0)
    if flips == 0:
        assert nerr(z, ref["z"]) < tol
        assert (torch.maximum(zc, zr)[mask_diff] <= 2e-6 * ref["y"].abs().max()).all()
    assert flips <= 8 and mask_diff.sum().item() <= 8
    check_gradients(layer, xs.grad, z, y._gcb_pooled[1], ref, signal, bary, wn, ws, prior, grad_z, "relu", range(a), tol)


@pytest.mark.gpu
def test_survives_torch_compile():
    """The reference demo wraps its model in torch.compile (training_demo.py:127): the drop-in layers are
    torch.library operators with fake kernels, so the model compiles as ONE graph (fullgraph=True) and gives the
    same result and gradients as the eager call."""
    dev = "cuda:0"
    v, f, r, a, t = 300, 32, 3, 6, 32

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.down = torch.nn.Linear(f, f)
            self.conv = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03)
            self.amp = AngularMaxPooling()
            self.head = torch.nn.Linear(t, 5)

        def forward(self, x, bary):
            return self.head(self.amp(self.conv([self.down(x), bary])))

    torch.manual_seed(0)
    net = Net().to(dev)
    x = O.make_signal(v, f, kind="randn").to(dev)
    bary = O.make_bary(v, r, a, seed=2).to(dev)
    ref = net(x, bary)
    ref.sum().backward()
    ref_grad = net.conv._template_neighbor_weights.grad.clone()
    net.zero_grad()
    compiled = torch.compile(net, backend="aot_eager", fullgraph=True)   # one graph: every step is a registered operator
    out = compiled(x, bary)
    out.sum().backward()
    assert nerr(out, ref) < 1e-6
    assert nerr(net.conv._template_neighbor_weights.grad, ref_grad) < 1e-6


@pytest.mark.gpu
def test_faust_shape_randn_signal_stays_inside_fp32_tolerance():
    """Parity stress of SURVEY.md §8d: a randn signal makes the contraction cancel heavily, which is where the
    truncating fp32 accumulation of the tensor core shows; the K-split count is chosen so that the fp32 mode
    stays inside 1e-5 against a float64 evaluation of the reference algorithm."""
    import os
    v, f, r, a, t = 6890, 544, 5, 8, 96
    wn, ws, b = O.make_weights(t, r, a, f, seed=0)
    signal = O.make_signal(v, f, seed=1, kind="randn")
    bary = O.make_bary(v, r, a, seed=1)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.028)))
    torch.set_num_threads(os.cpu_count() or 1)
    ref64 = O.conv_forward(signal.double(), bary, wn.double(), ws.double(), b.double(), prior.double(), "relu", 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028,
                         precision="fp32")
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to("cuda:0")
    with torch.no_grad():
        y = layer([signal.to("cuda:0"), bary.to("cuda:0")])
    assert nerr(y, ref64) < TOL_FP32


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["ffma", "fp32", "tf32"])
def test_row_block_with_row_offset(precision):
    """`row_offset` (extension): a block of rows from the middle of the mesh, indices addressing the whole signal,
    reproduces those rows of the whole-mesh result - centre term included - and of the oracle."""
    v, f, r, a, t = 700, 64, 3, 8, 32
    wn, ws, b = O.make_weights(t, r, a, f, seed=4)
    signal = O.make_signal(v, f, seed=4, kind="randn")
    bary = O.make_bary(v, r, a, seed=4, fallback_frac=0.05)
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(r, a, 0.03))).float()
    y_ref = O.conv_forward(signal, bary, wn, ws, b, prior, "relu", 1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.03,
                         precision=precision)
    layer.load_state_dict({"_template_neighbor_weights": wn, "_template_self_weights": ws, "_bias": b})
    layer = layer.to(DEV)
    tol = TOL_TF32 if precision == "tf32" else TOL_FP32
    xd, bd = signal.to(DEV), bary.to(DEV)
    with torch.no_grad():
        for lo, hi in ((0, 130), (130, 517), (517, 700)):
            y = layer([xd, bd[lo:hi].contiguous()], row_offset=lo)
            assert nerr(y, y_ref[lo:hi]) < tol
            z = AngularMaxPooling()(y)
            assert torch.equal(z, y[torch.arange(hi - lo, device=DEV), y._gcb_pooled[1].long()])
    with pytest.raises(ValueError):
        layer([xd, bd[600:].contiguous()], row_offset=650)
    y = layer([xd.clone().requires_grad_(True), bd[100:200].contiguous()], row_offset=100)
    with pytest.raises(NotImplementedError):
        AngularMaxPooling()(y).sum().backward()
