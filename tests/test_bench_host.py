"""Host-side logic of bench.py / synth.py that needs no GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ppr_diffphys_b200.synth import lerp_frames  # noqa: E402


def test_workload_table_and_algorithmic_bytes():
    for name, w in bench.WORKLOADS.items():
        assert {"robot", "bs", "window", "stride", "clearance", "lin_vel"} <= set(w), name
        assert w["window"] % w["stride"] == 0 or name.startswith("laikago-trot")   # frames land on substeps
    assert bench.DEFAULT_WORKLOAD in bench.WORKLOADS
    # S = nb*13*4, R = nqd*4; fwd 2S+R, fwd+bwd 5S+3R
    assert bench.algorithmic_bytes(13, 18) == dict(fwd=1424, bwd=2172, fwdbwd=3596)        # laikago
    assert bench.algorithmic_bytes(19, 60) == dict(fwd=2216, bwd=3444, fwdbwd=5660)        # human
    assert bench.algorithmic_bytes(26, 81) == dict(fwd=3028, bwd=4704, fwdbwd=7732)        # quad


def test_lerp_frames_is_the_linear_interpolation_of_the_frames():
    g = torch.Generator().manual_seed(0)
    frames = torch.randn(4, 5, generator=g)
    for stride, T in ((32, 97), (33, 100), (8, 20)):
        r = lerp_frames(frames, stride, T)
        assert r.shape == (T, 5)
        for t in range(T):
            k = min(t // stride, frames.shape[0] - 2)
            a = (t - k * stride) / stride
            assert torch.allclose(r[t], (1 - a) * frames[k] + a * frames[k + 1], atol=1e-6)
        nf = (T - 1) // stride + 1
        assert torch.equal(r[::stride][:nf], frames[:nf])           # exact at the frame steps


def test_traffic_file_is_stamped_with_a_kernel_source_hash():
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    h = bench.kernel_source_hash()
    assert isinstance(tj.get("kernel_source_hash"), str) and len(tj["kernel_source_hash"]) == len(h) == 16
    for wl in ("laikago-scaling-65536x64", "human-65536x64-contact"):
        assert tj[wl]["rollout_forward_kernel"] > 0 and tj[wl]["rollout_backward_kernel"] > 0
    if tj["kernel_source_hash"] != h:    # not an error: bench.py then reports traffic = null with a note
        print("profiles/traffic.json is stale for the current kernel sources (%s vs %s)" % (tj["kernel_source_hash"], h))


def test_step_functions_equal_the_plain_torch_expressions():
    """bench.py's hand-written StepLoss / MassChain: same value and gradients as the composed expressions."""
    StepLoss, MassChain = bench.step_functions()
    g = torch.Generator().manual_seed(1)
    for F in (1, 3):
        pos = torch.randn(F, 26, 7, generator=g, dtype=torch.float64).requires_grad_(True)
        vel = torch.randn(F, 26, 6, generator=g, dtype=torch.float64).requires_grad_(True)
        ref = (pos[-1, :, :3] - pos[0, :, :3]).pow(2).mean() + 1e-3 * vel[-1].pow(2).mean()
        gr = torch.autograd.grad(ref * 1.7, (pos, vel))
        out = StepLoss.apply(pos, vel)
        go = torch.autograd.grad(out * 1.7, (pos, vel))
        assert torch.allclose(out, ref, rtol=1e-12, atol=1e-14)
        for a, b in zip(go, gr):
            assert torch.allclose(a, b, rtol=1e-12, atol=1e-14)
    m = (torch.rand(13, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    nI = torch.randn(13, 3, 3, generator=g, dtype=torch.float64)
    nI = nI @ nI.transpose(1, 2) + torch.eye(3, dtype=torch.float64)
    nI_inv = torch.linalg.inv(nI)
    w = [torch.randn(13, generator=g, dtype=torch.float64), torch.randn(13, 3, 3, generator=g, dtype=torch.float64),
         torch.randn(13, 3, 3, generator=g, dtype=torch.float64)]
    ref = (1.0 / m, nI * m[:, None, None], torch.linalg.inv(nI * m[:, None, None]))
    out = MassChain.apply(m, nI, nI_inv)
    for a, b in zip(out, ref):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-12)
    gr, = torch.autograd.grad(sum((a * b).sum() for a, b in zip(ref, w)), m)
    go, = torch.autograd.grad(sum((a * b).sum() for a, b in zip(out, w)), m)
    assert torch.allclose(go, gr, rtol=1e-9, atol=1e-12)
