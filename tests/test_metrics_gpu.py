"""Device-side evaluation metrics against the oracle and the real reference `PAEval` golden."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import poem_oracle as orc  # noqa: E402
from poem_v2_b200 import _native as nat  # noqa: E402
from poem_v2_b200.metrics import MeanEPE, PAEval, pa_distances  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_pa.npz")


def test_pa_eval_matches_reference_golden():
    z = np.load(GOLD)
    ev = PAEval(None, mesh_score=True)
    t = {k: torch.from_numpy(z[k]).cuda() for k in ("gt_j", "gt_v", "pr_j", "pr_v")}
    ev.feed(t["pr_j"], t["gt_j"], t["pr_v"], t["gt_v"])
    m = ev.get_measures()
    got = np.array([m[k] for k in ("pa_mpjpe", "mpjpe", "pa_mpvpe", "mpvpe")])
    print("PAEval", got, "reference", z["measures"])
    assert np.allclose(got, z["measures"], rtol=2e-5)
    _, aligned = pa_distances(t["gt_j"], t["pr_j"], return_aligned=True)
    assert np.allclose(aligned.cpu().numpy(), z["aligned_j"], atol=2e-6)
    assert "pa_mpjpe(mm)" in str(ev) and ev.get_result() == m["pa_mpjpe"]
    ev.feed(t["pr_j"], t["gt_j"], t["pr_v"], t["gt_v"])           # running average over two identical batches
    assert abs(ev.get_measures()["pa_mpvpe"] - m["pa_mpvpe"]) < 1e-9
    ev.reset()
    assert ev.get_measures()["mpjpe"] == 0


@pytest.mark.parametrize("B,N", [(1, 21), (7, 778), (64, 799), (3, 4)])
def test_pa_distances_match_oracle(B, N):
    g = torch.Generator().manual_seed(B * 1000 + N)
    gt = torch.randn(B, N, 3, generator=g) * 0.05 + torch.tensor([0.1, -0.2, 0.6])
    ang = torch.randn(B, 3, generator=g)
    K = torch.zeros(B, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ang[:, 2], ang[:, 1], ang[:, 2], -ang[:, 0], -ang[:, 1], ang[:, 0]
    R = torch.linalg.matrix_exp(K)
    pred = (gt - gt.mean(1, keepdim=True)) @ R.transpose(1, 2) * (0.7 + 0.6 * torch.rand(B, 1, 1, generator=g)) \
        + gt.mean(1, keepdim=True) + 0.05 * torch.randn(B, 1, 3, generator=g) + 0.003 * torch.randn(B, N, 3, generator=g)
    want = orc.pa_distances(gt.numpy(), pred.numpy())
    got = pa_distances(gt.cuda(), pred.cuda()).cpu().numpy()
    print(f"B={B} N={N}: max rel diff {np.abs(got - want).max() / np.abs(want).max():.2e}")
    assert np.allclose(got, want, rtol=5e-5, atol=1e-8)


def test_mean_epe_and_errors():
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(4, 21, 3, generator=g), torch.randn(4, 21, 3, generator=g)
    m = MeanEPE(None, "joints_3d")
    m.feed(a.cuda(), b.cuda())
    want = torch.norm(a - b, dim=2).mean(1).sum().item() / 4
    assert abs(m.get_result() - want) < 1e-6 and list(m.get_measures()) == ["joints_3d_mepe"]
    with pytest.raises(nat.PoemError):
        pa_distances(a, b)                                   # CPU tensors: no CPU implementation
