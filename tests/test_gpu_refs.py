"""RefsFromFrames (per-frame control references -> per-substep, on the device) against the composed torch expression of
the reference's host-side linear interpolation (scipy interp1d(kind="linear"), dp_model.py:421-427)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def lerp_torch(frames, stride, T):
    t = torch.arange(T, device=frames.device)
    k0 = torch.clamp(t // stride, max=max(frames.shape[0] - 2, 0))
    a = ((t - k0 * stride).to(frames.dtype) / stride)[:, None]
    k1 = torch.clamp(k0 + 1, max=frames.shape[0] - 1)
    return frames[k0] * (1 - a) + frames[k1] * a


@pytest.mark.parametrize("F,stride,T,n", [(3, 32, 65, 7 * 18), (24, 33, 760, 10 * 18), (2, 16, 17, 5), (3, 32, 64, 11)])
def test_refs_from_frames_matches_torch_and_autograd(F, stride, T, n):
    from ppr_diffphys_b200 import RefsFromFrames
    g = torch.Generator().manual_seed(0)
    frames = torch.randn(F, n, generator=g).cuda().requires_grad_(True)
    refs = RefsFromFrames.apply(frames, stride, T)
    ref = lerp_torch(frames.detach().double(), stride, T)
    assert refs.shape == (T, n) and (refs.double() - ref).abs().max() < 1e-6
    # frame steps reproduce the frames exactly
    nf = (T - 1) // stride + 1
    assert torch.equal(refs[::stride][:nf], frames.detach()[:nf])
    w = torch.randn(T, n, generator=g).cuda()
    (refs * w).sum().backward()
    f2 = frames.detach().double().clone().requires_grad_(True)
    (lerp_torch(f2, stride, T) * w.double()).sum().backward()
    assert (frames.grad.double() - f2.grad).abs().max() < 1e-4 * max(1.0, float(f2.grad.abs().max()))
