"""Host-side collation of multi-view samples with a variable number of views per sample.

Mirror of the reference's `collation_random_n_views` (lib/utils/collation.py:7-25): a sample is a dict whose array
fields carry the views on axis 0; the batch concatenates them into flat `(sum V, ...)` tensors and records the view
counts in `cam_view_num` — the layout every entry point of this package consumes (`head.forward`,
`model.PtEmbedMultiviewStereoV2`).  Same name, argument and return structure, so it can be passed as the `collate_fn`
of the reference's DataLoader (`lib/datasets/__init__.py`).

One addition for the host-buffer entry points (`poem_head_forward_host`, include/poem_b200.h): `pin=True` places the
concatenated tensors in page-locked memory, so the host->device copies inside the C call are asynchronous.
"""
import numpy as np
import torch


def collation_random_n_views(batch, pin=False):
    if not isinstance(batch, list):
        batch = [batch]                      # a single sample (reference: "only 1 sample is provided")
    out = {}
    view_counts = [b["target_joints_3d"].shape[0] for b in batch]
    for key, first in batch[0].items():
        # the reference's rule: numeric ndarrays are concatenated along the view axis, everything else is listed
        if isinstance(first, np.ndarray) and not isinstance(first[0], str):
            t = torch.Tensor(np.concatenate([b[key] for b in batch], axis=0))
            out[key] = t.pin_memory() if pin and torch.cuda.is_available() else t
        else:
            out[key] = [b[key] for b in batch]
    out["cam_view_num"] = np.array(view_counts)
    return out
