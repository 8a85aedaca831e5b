"""the BatchNorm1d -> Dropout chain between two convolutions (gcb_bn_stats / gcb_bn_dropout_fwd /
gcb_bn_dropout_bwd, geoconv_b200.pytorch.layers.batch_norm_dropout) against torch's own modules."""
import pytest
import torch
from torch import nn


def test_module_surface_matches_batchnorm1d():
    from geoconv_b200.pytorch.layers.batch_norm_dropout import BatchNormDropout
    ours, ref = BatchNormDropout(96), nn.BatchNorm1d(96)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    with pytest.raises(RuntimeError):
        ours(torch.zeros(8, 96))           # no CPU path
    with pytest.raises(ValueError):
        BatchNormDropout(10)


def _pair(t, p, seed=0):
    from geoconv_b200.pytorch.layers.batch_norm_dropout import BatchNormDropout
    torch.manual_seed(seed)
    ours, ref = BatchNormDropout(t, p=p).cuda(), nn.BatchNorm1d(t).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
    ours.load_state_dict(ref.state_dict())
    return ours, ref


@pytest.mark.gpu
@pytest.mark.parametrize("v,t", [(6890, 96), (1000, 384), (37, 64)])
def test_batchnorm_parity_without_dropout(v, t):
    ours, ref = _pair(t, 0.0)
    z = (torch.randn(v, t, device="cuda") * 3 + 1.5)
    za, zb = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    g = torch.randn(v, t, device="cuda")
    for _step in range(2):                          # running statistics accumulate over two steps
        ya, yb = ours(za), ref(zb)
    assert torch.allclose(ya, yb, atol=2e-5, rtol=1e-5)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(za.grad, zb.grad, atol=2e-5, rtol=1e-4)
    assert torch.allclose(ours.weight.grad, ref.weight.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(ours.bias.grad, ref.bias.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(ours.running_mean, ref.running_mean, atol=1e-6)
    assert torch.allclose(ours.running_var, ref.running_var, rtol=1e-5)
    assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked) == 2
    ours.eval(), ref.eval()
    za2, zb2 = z.clone().requires_grad_(True), z.clone().requires_grad_(True)
    ya, yb = ours(za2), ref(zb2)
    assert torch.allclose(ya, yb, atol=2e-5, rtol=1e-5)
    ya.backward(g)
    yb.backward(g)
    assert torch.allclose(za2.grad, zb2.grad, atol=1e-5, rtol=1e-5)


@pytest.mark.gpu
def test_dropout_mask_statistics_and_backward():
    v, t, p = 6890, 96, 0.2
    ours, ref = _pair(t, p)
    z = torch.randn(v, t, device="cuda") + 2.0
    za = z.clone().requires_grad_(True)
    ya = ours(za)
    kept = ya != 0
    assert abs(kept.float().mean().item() - (1 - p)) < 5e-3            # 661 k draws: sigma = 5e-4
    assert abs(kept.float().mean(0).min().item() - (1 - p)) < 0.03     # no dead / always-on column
    assert abs(kept.float().mean(1).min().item() - (1 - p)) < 0.2
    # the kept elements are BatchNorm's, scaled by 1 / (1 - p); the backward applies the SAME mask
    yb = ref(z)
    assert torch.allclose(ya[kept], yb[kept] / (1 - p), atol=3e-5, rtol=1e-5)
    g = torch.randn(v, t, device="cuda")
    ya.backward(g)
    zb = z.clone().requires_grad_(True)
    (ref(zb) * kept / (1 - p)).backward(g)
    assert torch.allclose(za.grad, zb.grad, atol=2e-5, rtol=1e-4)
    assert torch.allclose(ours.weight.grad, ref.weight.grad, atol=1e-3, rtol=1e-4)
    # a new call draws a new mask; the same (seed, counter) reproduces it bit for bit
    yc = ours(z)
    assert not torch.equal(yc != 0, kept)
    from geoconv_b200 import ops
    mean, invstd = ops.bn_stats_op(z, 1e-5, None, None, 0.0)
    one = ops.bn_dropout_op(z, mean, invstd, ours.weight, ours.bias, True, p, 1234, 7)
    two = ops.bn_dropout_op(z, mean, invstd, ours.weight, ours.bias, True, p, 1234, 7)
    other = ops.bn_dropout_op(z, mean, invstd, ours.weight, ours.bias, True, p, 1235, 7)
    assert torch.equal(one, two) and not torch.equal(one != 0, other != 0)


@pytest.mark.gpu
def test_block_chain_with_the_layers():
    """do -> isc -> amp -> bn as the reference's block runs it, with the bn -> do pair of consecutive blocks fused:
    gradients reach the first block's weights through two convolutions and the fused chain, bit-reproducibly."""
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.batch_norm_dropout import BatchNormDropout
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    from oracle import geoconv_oracle as O
    v, r, a = 1500, 5, 8
    torch.manual_seed(0)
    c1 = ConvGeodesic(input_shape=[(None, 64), (None, r, a, 3, 2)], amt_templates=96, template_radius=0.028).cuda()
    c2 = ConvGeodesic(input_shape=[(None, 96), (None, r, a, 3, 2)], amt_templates=128, template_radius=0.028).cuda()
    chain, amp = BatchNormDropout(96, p=0.2).cuda(), AngularMaxPooling()
    signal = O.make_signal(v, 64, seed=1).cuda()
    bary = O.make_bary(v, r, a, seed=1).cuda()

    def run():
        import geoconv_b200.pytorch.layers.batch_norm_dropout as M
        M._CALLS = 0
        for m in (c1, c2, chain):
            m.zero_grad()
        out = amp(c2([chain(amp(c1([signal, bary]))), bary]))
        out.square().mean().backward()
        return out.detach().clone(), c1._template_neighbor_weights.grad.clone(), chain.weight.grad.clone()
    o1, g1, b1 = run()
    o2, g2, b2 = run()
    assert torch.isfinite(o1).all() and g1.abs().max() > 0 and b1.abs().max() > 0
    assert torch.equal(o1, o2) and torch.equal(g1, g2) and torch.equal(b1, b2)
