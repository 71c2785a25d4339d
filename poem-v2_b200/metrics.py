"""Device-side mirrors of the reference's evaluation metrics: `PAEval` (lib/metrics/pa_eval.py:15-124)
and `MeanEPE` (lib/metrics/mean_epe.py:11-45) with the same `feed / get_measures / get_result / reset / __str__`
interface.  The reference copies every batch to the host and loops over samples around
`scipy.linalg.orthogonal_procrustes`; here one kernel launch per point set computes the per-sample aligned and raw
mean distances, the running sums stay on the device and the host only synchronises when a measure is read.
No CPU fallback: inputs must be CUDA tensors."""
import torch

from . import _native as nat


def pa_distances(gt, pred, return_aligned=False):
    """gt, pred: (B, N, 3) CUDA tensors -> (B, 2) [Procrustes-aligned mean distance, raw mean distance] (+ aligned pred)."""
    if not (gt.is_cuda and pred.is_cuda):
        raise nat.PoemError("metrics inputs must be CUDA tensors: there is no CPU implementation")
    assert gt.shape == pred.shape and gt.dim() == 3 and gt.shape[-1] == 3, (tuple(gt.shape), tuple(pred.shape))
    g = gt.detach().contiguous().float()
    p = pred.detach().contiguous().float()
    out = torch.empty(g.shape[0], 2, device=g.device)
    aligned = torch.empty_like(p) if return_aligned else None
    nat.check(nat.load().poem_pa_metrics(g.data_ptr(), p.data_ptr(), g.shape[0], g.shape[1], out.data_ptr(),
                                         aligned.data_ptr() if return_aligned else None,
                                         torch.cuda.current_stream(g.device).cuda_stream))
    return (out, aligned) if return_aligned else out


class _DeviceMeter:
    """AverageMeter (lib/metrics/basic_metric.py:32-57) whose sum lives on the device until it is read."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._sum, self.count = None, 0

    def update(self, total, n):
        total = total.double()     # the reference's AverageMeter accumulates Python floats (float64)
        self._sum = total if self._sum is None else self._sum + total
        self.count += n

    @property
    def sum(self):
        return 0.0 if self._sum is None else float(self._sum.item())

    @property
    def avg(self):
        return 0 if self.count == 0 else self.sum / self.count


class PAEval:
    def __init__(self, cfg=None, mesh_score=False):
        self.mesh_score = mesh_score
        self.pa_mpjpe, self.mpjpe = _DeviceMeter(), _DeviceMeter()
        self.pa_mpvpe, self.mpvpe = _DeviceMeter(), _DeviceMeter()
        self.count, self.skip = 0, False

    def reset(self):
        for m in (self.pa_mpjpe, self.mpjpe, self.pa_mpvpe, self.mpvpe):
            m.reset()

    def feed(self, pred_joints_3d_abs, joints_3d_abs, pred_verts_3d_abs=None, verts_3d_abs=None, **kwargs):
        bs = pred_joints_3d_abs.shape[0]
        d = pa_distances(joints_3d_abs.to(pred_joints_3d_abs.device), pred_joints_3d_abs).sum(dim=0)
        self.pa_mpjpe.update(d[0], bs)
        self.mpjpe.update(d[1], bs)
        if self.mesh_score:
            d = pa_distances(verts_3d_abs.to(pred_verts_3d_abs.device), pred_verts_3d_abs).sum(dim=0)
            self.pa_mpvpe.update(d[0], bs)
            self.mpvpe.update(d[1], bs)

    def get_measures(self, **kwargs):
        m = {"pa_mpjpe": self.pa_mpjpe.avg, "mpjpe": self.mpjpe.avg}
        if self.mesh_score:
            m["pa_mpvpe"], m["mpvpe"] = self.pa_mpvpe.avg, self.mpvpe.avg
        return m

    def get_result(self):
        return self.pa_mpjpe.avg

    def __str__(self):
        s = f"pa_mpjpe(mm): {self.pa_mpjpe.avg * 1000.0 :6.4f} | mpjpe: {self.mpjpe.avg:6.4f}"
        if self.mesh_score:
            s += f" | pa_mpvpe(mm): {self.pa_mpvpe.avg * 1000.0:6.4f} | mpvpe: {self.mpvpe.avg:6.4f}"
        return s


class MeanEPE:
    def __init__(self, cfg=None, name=""):
        self.name = f"{name}_mepe"
        self.avg_meter = _DeviceMeter()
        self.count, self.skip = 0, False

    def reset(self):
        self.avg_meter.reset()

    def feed(self, pred_kp, gt_kp, kp_vis=None, **kwargs):
        assert pred_kp.dim() == 3, "pred shape should be (BATCH, NPOINTS, 1|2|3)"
        total = torch.norm(pred_kp - gt_kp.to(pred_kp.device), p="fro", dim=2).mean(dim=1).sum()
        self.avg_meter.update(total, pred_kp.shape[0])
        return total

    def get_measures(self, **kwargs):
        return {self.name: self.avg_meter.avg}

    def get_result(self):
        return self.avg_meter.avg

    def __str__(self):
        return f"{self.name}: {self.avg_meter.avg:6.4f}"
