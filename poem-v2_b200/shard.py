"""Sample sharding across the GPUs of one box.

Samples are independent units of the decoder path (all views of a sample stay together because the cross-view
merge mixes them), so N ranks each run the whole path on a contiguous slice of the batch; there is no
data-path collective.  The only optional exchange is an all_gather of the (NB, B/N, 799, 3) results.

When there are fewer samples than GPUs (serving: one sample, eight views, eight GPUs) the image half is the part that
still shards: the B*V images are split across the ranks (`image_bounds`), every rank runs backbone + feat_decode on its
images and ONE all_gather exchanges the (., 160, 16, 16) feature maps (`gather_features`, 82 KB per image in bf16, 164
KB in fp32) — the single real exchange step of the path; the decoder then runs on the sample-sharding
above (ranks without a sample idle through it).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(view_counts, world_size):
    """Contiguous split of the samples into `world_size` slices, balanced by number of views (the unit of work
    of the sampler / merge stage).  Returns [(s0, e0), ...] sample ranges; slices may be empty when B < world."""
    v = np.asarray(view_counts, dtype=np.int64)
    B = len(v)
    cum = np.concatenate([[0], np.cumsum(v)])
    total = cum[-1]
    bounds, start = [], 0
    for r in range(world_size):
        if r == world_size - 1:
            end = B
        else:
            target = total * (r + 1) / world_size
            end = int(np.searchsorted(cum, target, side="left"))
            end = min(max(end, start), B)
            # choose the closer of end-1 / end to the target
            if end > start and abs(cum[end - 1] - target) <= abs(cum[end] - target):
                end -= 1
            if B >= world_size:
                # never empty when there are enough samples: at least one for this rank, and at least one left for
                # every later rank (a rank with an empty slice would skip the head while the others wait in all_gather)
                end = min(max(end, start + 1), B - (world_size - 1 - r))
            else:
                end = min(max(end, start), B)
        bounds.append((start, end))
        start = end
    return bounds


def shard_inputs(mlvl_feat, img_metas, reference_joints, rank, world_size):
    """Slice one rank's samples (and their images) out of a full batch."""
    views = np.asarray(img_metas["cam_view_num"]).astype(np.int64)
    s, e = shard_bounds(views, world_size)[rank]
    img0, img1 = int(views[:s].sum()), int(views[:e].sum())
    metas = dict(img_metas)
    metas["cam_intr"] = img_metas["cam_intr"][img0:img1]
    metas["cam_extr"] = img_metas["cam_extr"][img0:img1]
    metas["cam_view_num"] = views[s:e]
    metas["master_id"] = list(img_metas["master_id"][s:e])
    return mlvl_feat[img0:img1], metas, reference_joints[s:e], (s, e)


def gather_outputs(local_coords, batch_total, bounds, group=None):
    """all_gather the per-rank (NB, b_r, Q, 3) predictions into (NB, B, Q, 3) on every rank."""
    world = dist.get_world_size(group)
    nb, _, q, _ = local_coords.shape
    width = max(e - s for s, e in bounds)
    pad = torch.zeros(nb, width, q, 3, dtype=local_coords.dtype, device=local_coords.device)
    pad[:, :local_coords.shape[1]] = local_coords
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty(nb, batch_total, q, 3, dtype=local_coords.dtype, device=local_coords.device)
    for (s, e), part in zip(bounds, parts):
        out[:, s:e] = part[:, :e - s]
    return out


def image_bounds(n_images, world_size):
    """Contiguous split of the flat image axis into `world_size` near-equal slices [(i0, i1), ...] (may be empty)."""
    base, extra = divmod(int(n_images), world_size)
    bounds, start = [], 0
    for r in range(world_size):
        end = start + base + (1 if r < extra else 0)
        bounds.append((start, end))
        start = end
    return bounds


def gather_features(local_feat, n_images, bounds, group=None):
    """all_gather of the per-rank feature maps (n_r, C, h, w) into (n_images, C, h, w) on every rank: the exchange step
    of the image-sharded mode.  Slices are padded to the widest one so a single fixed-size collective moves them."""
    world = dist.get_world_size(group)
    width = max(e - s for s, e in bounds)
    pad = torch.zeros((width,) + tuple(local_feat.shape[1:]), dtype=local_feat.dtype, device=local_feat.device)
    pad[:local_feat.shape[0]] = local_feat
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((int(n_images),) + tuple(local_feat.shape[1:]), dtype=local_feat.dtype, device=local_feat.device)
    for (s, e), part in zip(bounds, parts):
        out[s:e] = part[:e - s]
    return out
