"""geoconv_b200.data.staging: one pinned write-combined arena per step, one H2D copy."""
import pytest
import torch


@pytest.mark.gpu
def test_pinned_arena_round_trip():
    from geoconv_b200.data.staging import PinnedArena, device_arena
    like = [torch.randn(37, 5), torch.randn(11, 2, 3, 3, 2, dtype=torch.float64), torch.arange(7, dtype=torch.int32)]
    host = PinnedArena(like).fill(like)
    assert host.buffer.is_pinned() and host.nbytes % 256 == 0 and host.write_combined
    buf, views = device_arena(like, "cuda:0")
    assert buf.numel() == host.nbytes
    buf.copy_(host.buffer, non_blocking=True)
    torch.cuda.synchronize()
    for got, want in zip(views, like):
        assert got.dtype == want.dtype and got.shape == want.shape and got.data_ptr() % 256 == 0
        assert torch.equal(got.cpu(), want)
    plain = PinnedArena(like, write_combined=False).fill(like)
    assert plain.buffer.is_pinned() and torch.equal(plain.views[0], like[0])
    del host, plain
