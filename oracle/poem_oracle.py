"""TEST INFRASTRUCTURE — CPU/fp32 restatement of the reference decoder path. NOT product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file; the product (`poem-v2_b200/`) must never route through it.

Parity status: **pinned against outputs of the reference itself run in the build container** —
`oracle/make_golden.py` runs the unmodified reference `POEM_Generalized_Head` (under the stub layer
`oracle/ref_shim.py`) on seeded inputs/weights and stores its outputs under `tests/golden/`;
`tests/test_oracle.py` asserts this restatement reproduces them (≤1e-5 normalised units).
Unpinned sub-parts (third-party arithmetic absent offline): pytorch3d `knn_points` tie order and the
real MANO template (a seeded stand-in is used) — see DESIGN.md.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ a2
def sine_pos_3d(n_views, h, w, num_feats, normalize=True, temperature=10000.0, scale=2 * math.pi, eps=1e-6):  # CPU
    """`SinePositionalEncoding3D.forward` on an all-false mask (lib/models/layers/petr_transformer.py:434-469).
    Returns (n_views, 3*num_feats, h, w); channel order [view, y, x]."""
    ones = torch.ones(1, n_views, h, w, dtype=torch.float32)
    n_e, y_e, x_e = ones.cumsum(1), ones.cumsum(2), ones.cumsum(3)
    if normalize:
        n_e = n_e / (n_e[:, -1:] + eps) * scale
        y_e = y_e / (y_e[:, :, -1:] + eps) * scale
        x_e = x_e / (x_e[:, :, :, -1:] + eps) * scale
    i = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_feats)

    def enc(e):
        p = e[..., None] / dim_t
        # NB: stacking a 5-D tensor at dim=4 gives (...,2,F/2): all sines first, then all cosines
        # (NOT interleaved as in the 2-D DETR encoding) — petr_transformer.py:465-467
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=4).reshape(1, n_views, h, w, -1)
    pos = torch.cat((enc(n_e), enc(y_e), enc(x_e)), dim=4).permute(0, 1, 4, 2, 3)
    return pos[0].contiguous()


def feature_volume(sd, feat, view_counts, dims):
    """x = input_proj(feat) + adapt_pos3d(sine) (lib/models/heads/ptEmb_head.py:835-870)."""
    D = dims.embed_dims
    x = F.conv2d(feat, sd["input_proj.weight"], sd["input_proj.bias"])
    pos = []
    for n in view_counts:
        s = sine_pos_3d(int(n), x.shape[-2], x.shape[-1], dims.pos_feats, dims.pos_normalize).to(feat.device)
        pos.append(F.conv2d(s, sd["adapt_pos3d.weight"], sd["adapt_pos3d.bias"]))
    return x + torch.cat(pos, dim=0)


# ------------------------------------------------------------------------------------------ a3/a4
def project_bps(bps_world, cam_intr, cam_extr, view_counts, inp_res):
    """lib/utils/collation.py:48-65 + lib/utils/transform.py:898-930 + ptEmb_head.py:880-883.
    bps_world (B,P,3); returns the grid_sample grid (BV,P,1,2) in [-1,1] units."""
    out = []
    s = 0
    for b, n in enumerate(view_counts):
        n = int(n)
        T = torch.linalg.inv(cam_extr[s:s + n])                      # master -> camera
        p = bps_world[b][None].expand(n, -1, -1)                    # (n,P,3)
        pc = (T[:, :3, :3] @ p.transpose(1, 2)).transpose(1, 2) + T[:, :3, 3][:, None]
        q = (cam_intr[s:s + n] @ pc.transpose(1, 2)).transpose(1, 2)
        z = q[..., 2:].clone()
        z[z.abs() < 1e-7] = 1e-7
        out.append(q[..., :2] / z)
        s += n
    uv = torch.cat(out, dim=0)[:, :, None, :]                       # (BV,P,1,2)
    uv = uv * (1.0 / inp_res)
    return uv * 2 - 1


# ------------------------------------------------------------------------------------------ a6
def _mlp2(sd, prefix, x):
    x = F.relu(F.linear(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"]))
    return F.linear(x, sd[prefix + ".2.weight"], sd[prefix + ".2.bias"])


def merge_views(sd, sampled, view_counts):
    """Raw `.view(1,-1,N,D)` regroup + merge_features_mv / _sv (ptEmb_head.py:745-771,910-926)."""
    outs = []
    s = 0
    D = sampled.shape[1]
    for n in view_counts:
        n = int(n)
        q = sampled[s:s + n].contiguous().view(1, -1, n, D)         # memory reinterpretation, NOT a permute
        s += n
        if n == 1:
            q = q.squeeze(2)
            outs.append(q + _mlp2(sd, "merge_net_feature.1", _mlp2(sd, "merge_net_feature.0", q)))
            continue
        q1 = q[:, :, 0]
        m = _mlp2(sd, "merge_net_feature.0", q)                     # (1,P,n,D/2)
        master, other = m[:, :, 0], m[:, :, 1:]
        w = other @ master[..., None]                               # (1,P,n-1,1)
        agg = (other.transpose(2, 3) @ w).squeeze(-1)               # (1,P,D/2)
        outs.append(q1 + _mlp2(sd, "merge_net_feature.1", agg) / n)
    return torch.cat(outs, dim=0)


# Test hook (not in the reference): {site name: keep / (1 - p) mask} of the training-mode dropout layers, so that a device
# run with dropout and the oracle can be compared on the same masks (nn.Dropout(hidden_dropout_prob) at
# pt_metro_transformer.py:117,185-186 and inside HF BertSelfAttention / BertSelfOutput / BertOutput).  None = eval mode.
DROPOUT_MASKS = None


def _drop(name, x):
    if DROPOUT_MASKS is None:
        return x
    return x * DROPOUT_MASKS[name].reshape(x.shape)


# ------------------------------------------------------------------------------------------ a10
def bert_cross_attention(sd, prefix, hidden, enc, n_heads):
    """HF 4.x `BertAttention` with `encoder_hidden_states`: Q from `hidden`, K/V from `enc`, no mask,
    then BertSelfOutput (dense + residual + LayerNorm eps 1e-12). Call sites
    lib/models/bricks/pt_metro_transformer.py:57-74."""
    B, Lq, D = hidden.shape
    hd = D // n_heads

    def split(t):
        return t.view(B, -1, n_heads, hd).transpose(1, 2)
    q = split(F.linear(hidden, sd[prefix + ".self.query.weight"], sd[prefix + ".self.query.bias"]))
    k = split(F.linear(enc, sd[prefix + ".self.key.weight"], sd[prefix + ".self.key.bias"]))
    v = split(F.linear(enc, sd[prefix + ".self.value.weight"], sd[prefix + ".self.value.bias"]))
    p = _drop(prefix + ".probs", torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1))
    ctx = (p @ v).transpose(1, 2).reshape(B, Lq, D)
    o = _drop(prefix + ".hidden", F.linear(ctx, sd[prefix + ".output.dense.weight"], sd[prefix + ".output.dense.bias"]))
    return F.layer_norm(o + hidden, (D,), sd[prefix + ".output.LayerNorm.weight"],
                        sd[prefix + ".output.LayerNorm.bias"], eps=1e-12)


# ------------------------------------------------------------------------------------------ a14
def knn(query_xyz, ref_xyz, K):
    """pytorch3d `knn_points(p1,p2,K)` (third-party, v0.7.2, not vendored): squared L2, K smallest in
    ascending order; ties resolved towards the lower index. Returns idx (B,Lq,K) int64."""
    d = ((query_xyz[:, :, None, :] - ref_xyz[:, None, :, :]) ** 2).sum(-1)
    return torch.sort(d, dim=-1, stable=True).indices[..., :K]


def gather_rows(table, idx):
    """`index_points` (lib/utils/points_utils.py:9-20): table (B,N,C), idx (B,S,K) -> (B,S,K,C)."""
    B, S, K = idx.shape
    flat = idx.reshape(B, S * K)
    return torch.gather(table, 1, flat[..., None].expand(-1, -1, table.shape[-1])).reshape(B, S, K, -1)


# ------------------------------------------------------------------------------------------ a12/a13
def _vector_attention_core(sd, p, q, k, v, rel):
    pos = _mlp2(sd, p + "fc_delta", rel)
    a = _mlp2(sd, p + "fc_gamma", q[:, :, None] - k + pos)
    a = torch.softmax(a / math.sqrt(k.shape[-1]), dim=-2)
    return (a * (v + pos)).sum(dim=2)


def vec_self_attention(sd, p, xyz, feats, K, anchors=None, idx_forced=None):
    """`ptTransformerBlock._forward` (lib/models/bricks/point_transformers.py:70-96).
    `idx_forced` (test hook, not in the reference): use these neighbour indices instead of running `knn`."""
    B = xyz.shape[0]
    if anchors is not None:                                          # block 0: fixed anchors for every query
        a_xyz, a_idx = anchors
        idx = a_idx[None, None].expand(B, xyz.shape[1], -1)
        nbr_xyz = a_xyz[None, None].expand(B, xyz.shape[1], -1, -1)
    else:
        idx = knn(xyz, xyz, K) if idx_forced is None else idx_forced
        nbr_xyz = gather_rows(xyz, idx)
    x = F.linear(feats, sd[p + "fc1.weight"], sd[p + "fc1.bias"])
    q = F.linear(x, sd[p + "w_qs.weight"])
    k = gather_rows(F.linear(x, sd[p + "w_ks.weight"]), idx)
    v = gather_rows(F.linear(x, sd[p + "w_vs.weight"]), idx)
    res = _vector_attention_core(sd, p, q, k, v, xyz[:, :, None] - nbr_xyz)
    return F.linear(res, sd[p + "fc2.weight"], sd[p + "fc2.bias"]) + feats, idx


def vec_cross_attention(sd, p, pt_xyz, pt_feats, q_xyz, q_feats, K, anchors=None, idx_forced=None):
    """`ptTransformerBlock_CrossAttn._forward` (lib/models/bricks/point_transformers.py:125-156).
    `idx_forced`: see `vec_self_attention`."""
    B = q_xyz.shape[0]
    if anchors is not None:                                          # anchor idx gathers rows of the 4096-row table
        a_xyz, a_idx = anchors
        idx = a_idx[None, None].expand(B, q_xyz.shape[1], -1)
        nbr_xyz = a_xyz[None, None].expand(B, q_xyz.shape[1], -1, -1)
    else:
        idx = knn(q_xyz, pt_xyz, K) if idx_forced is None else idx_forced
        nbr_xyz = gather_rows(pt_xyz, idx)
    nf = gather_rows(pt_feats, idx)
    q = F.linear(q_feats, sd[p + "w_qs.weight"])
    x = F.linear(nf, sd[p + "fc1.weight"], sd[p + "fc1.bias"])
    k = F.linear(x, sd[p + "w_ks.weight"])
    v = F.linear(x, sd[p + "w_vs.weight"])
    res = _vector_attention_core(sd, p, q, k, v, q_xyz[:, :, None] - nbr_xyz)
    return F.linear(res, sd[p + "fc2.weight"], sd[p + "fc2.bias"]) + q_feats, idx


# ------------------------------------------------------------------------------------------ a9-a11
def metro_block(sd, i, dims, q_xyz, q_feats, pt_xyz, pt_feats, anchors, stages=None, neighbours=None):
    """`point_METRO_block.forward` + `point_METRO_layer.forward` + `pointer_layer.forward`
    (lib/models/bricks/pt_metro_transformer.py:153-200, 56-91, 34-40); eval mode (dropout = identity)."""
    p = f"transformer.pt_metro_encoder.{i}."
    D = dims.embed_dims
    qe = _drop(p + "qe", F.linear(q_feats, sd[p + "embedding.weight"], sd[p + "embedding.bias"]))
    ke = _drop(p + "ke", F.linear(pt_feats, sd[p + "embedding.weight"], sd[p + "embedding.bias"]))
    a1 = bert_cross_attention(sd, p + "encoder.attn", qe, ke, dims.n_heads)
    a2 = bert_cross_attention(sd, p + "encoder.cross_attn", a1, ke, dims.n_heads)
    anc = anchors if i == 0 else None
    nb_s, nb_c = (None, None) if (neighbours is None or i == 0) else neighbours
    f1, idx_s = vec_self_attention(sd, p + "encoder.vec_attn.query_self_attn.", q_xyz, a2, dims.n_neighbor, anc, nb_s)
    f2, idx_c = vec_cross_attention(sd, p + "encoder.vec_attn.query_cross_attn.", pt_xyz, ke, q_xyz, f1,
                                    dims.n_neighbor, anc, nb_c)
    xyz = _mlp2(sd, p + "encoder.vec_attn.reg_branch", f2) + q_xyz
    h = F.gelu(F.linear(f2, sd[p + "encoder.intermediate.dense.weight"], sd[p + "encoder.intermediate.dense.bias"]))
    o = _drop(p + "ffn", F.linear(h, sd[p + "encoder.output.dense.weight"], sd[p + "encoder.output.dense.bias"]))
    out = F.layer_norm(o + f2, (D,), sd[p + "encoder.output.LayerNorm.weight"],
                       sd[p + "encoder.output.LayerNorm.bias"], eps=1e-12)
    if stages is not None:
        stages[f"b{i}.a1"], stages[f"b{i}.a2"], stages[f"b{i}.f1"], stages[f"b{i}.f2"] = a1, a2, f1, f2
        stages[f"b{i}.xyz"], stages[f"b{i}.out"] = xyz, out
        stages[f"b{i}.idx_self"], stages[f"b{i}.idx_cross"] = idx_s, idx_c
    return out, xyz


# ------------------------------------------------------------------------------------------ a16
def rotation_6d_to_matrix(d6):
    """pytorch3d v0.7.2 `transforms.rotation_6d_to_matrix` (third-party, absent offline — published algorithm of
    Zhou et al. 2019): Gram-Schmidt on the two 3-vectors, rows (b1, b2, b1 x b2)."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def matrix_to_quaternion(m):
    """pytorch3d v0.7.2 `transforms.matrix_to_quaternion` (real part first): four candidate quaternions, the one
    with the largest denominator is taken (no sign standardisation in this version)."""
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(batch + (9,)), dim=-1)
    pos = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22],
                      dim=-1)
    q_abs = torch.where(pos > 0, torch.sqrt(pos.clamp_min(0)), torch.zeros_like(pos))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    pick = q_abs.argmax(dim=-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(batch + (1, 4))).squeeze(-2)


def quaternion_to_axis_angle(q):
    """pytorch3d v0.7.2 `transforms.quaternion_to_axis_angle`."""
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    k = torch.empty_like(ang)
    k[~small] = torch.sin(half[~small]) / ang[~small]
    k[small] = 0.5 - (ang[small] * ang[small]) / 48
    return q[..., 1:] / k


def rot6d_to_aa(r6):
    """lib/utils/transform.py:448-466: Compose([rotation_6d_to_matrix, matrix_to_quaternion,
    quaternion_to_axis_angle])."""
    return quaternion_to_axis_angle(matrix_to_quaternion(rotation_6d_to_matrix(r6)))


def axis_angle_to_rotmat(aa):
    """Rodrigues through a unit quaternion, as manopth/manotorch do (`batch_rodrigues`: norm of aa + 1e-8, half
    angle, quaternion -> matrix).  Third-party (manotorch v0.0.2), restated from the published algorithm."""
    ang = torch.norm(aa + 1e-8, p=2, dim=-1, keepdim=True)
    axis = aa / ang
    q = torch.cat([torch.cos(0.5 * ang), torch.sin(0.5 * ang) * axis], dim=-1)
    q = q / q.norm(p=2, dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=-1).reshape(aa.shape[:-1] + (3, 3))


# MANO kinematic tree (parent of each of the 16 joints), the fingertip vertices manotorch v0.0.2 appends for a right hand
# and its 16 joints + 5 tips -> 21-joint hand order (third-party constants, restated; the product keeps its own copy in
# poem_v2_b200/params.py and tests/test_oracle.py asserts the two agree)
MANO_PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)
MANO_TIP_VERTS = (745, 317, 444, 556, 673)
MANO_JOINT_ORDER = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)


def mano_forward(mano, pose_aa, betas, center_idx=None):
    """manotorch v0.0.2 `ManoLayer(rot_mode="axisang", use_pca=False, flat_hand_mean=True, center_idx=...)` —
    the MANO linear-blend-skinning forward (Romero et al. 2017; third-party, absent offline, restated):
    shape blend -> joints -> pose blend -> kinematic chain -> skinning -> 16 joints + 5 tip vertices reordered to
    the 21-joint convention -> optional root-centring.  `mano`: dict of `synth.synthetic_mano()` roles.
    Returns verts (B,778,3), joints (B,21,3)."""
    B = pose_aa.shape[0]
    R = axis_angle_to_rotmat(pose_aa.reshape(B, 16, 3))                         # (B,16,3,3)
    pose_map = (R[:, 1:] - torch.eye(3, dtype=R.dtype)).reshape(B, 135)
    v_shaped = mano["v_template"][None] + torch.einsum("vck,bk->bvc", mano["shapedirs"], betas)
    J = torch.einsum("jv,bvc->bjc", mano["J_regressor"], v_shaped)              # (B,16,3)
    v_posed = v_shaped + torch.einsum("vck,bk->bvc", mano["posedirs"], pose_map)
    G = [None] * 16
    for j in range(16):
        T = torch.zeros(B, 4, 4, dtype=R.dtype)
        T[:, :3, :3] = R[:, j]
        T[:, 3, 3] = 1.0
        par = MANO_PARENTS[j]
        if par < 0:
            T[:, :3, 3] = J[:, j]
            G[j] = T
        else:
            T[:, :3, 3] = J[:, j] - J[:, par]
            G[j] = G[par] @ T
    G = torch.stack(G, dim=1)                                                   # (B,16,4,4)
    A = G.clone()
    A[..., :3, 3] = G[..., :3, 3] - torch.einsum("bjrc,bjc->bjr", G[..., :3, :3], J)   # remove the rest pose
    Tv = torch.einsum("vj,bjrc->bvrc", mano["weights"], A)                      # (B,778,4,4)
    vh = torch.cat([v_posed, torch.ones(B, 778, 1, dtype=R.dtype)], dim=-1)
    verts = torch.einsum("bvrc,bvc->bvr", Tv, vh)[..., :3]
    jtr = torch.cat([G[..., :3, 3], verts[:, list(MANO_TIP_VERTS)]], dim=1)[:, list(MANO_JOINT_ORDER)]
    if center_idx is not None:
        c = jtr[:, center_idx:center_idx + 1]
        jtr, verts = jtr - c, verts - c
    return verts, jtr


def parametric_tail(sd, i, dims, feats, xyz, mano):
    """`point_METRO_block.get_parametric_output` (lib/models/bricks/pt_metro_transformer.py:139-151): the
    (B,799,D) features are RE-INTERPRETED as (B*D,799) rows (not transposed), Linear(799,1), Linear(D,106),
    16 x 6-D rotations -> axis-angle, MANO forward; joints and vertices of `xyz` are overwritten."""
    p = f"transformer.pt_metro_encoder.{i}."
    D = dims.embed_dims
    flat = F.linear(feats.reshape(-1, dims.n_query), sd[p + "flat_verts.weight"], sd[p + "flat_verts.bias"])
    par = F.linear(flat.reshape(-1, D), sd[p + "mano_linear.weight"], sd[p + "mano_linear.bias"])
    pose = rot6d_to_aa(par[:, :96].reshape(-1, 16, 6)).reshape(-1, 48)
    betas = par[:, 96:]
    verts, joints = mano_forward(mano, pose, betas, dims.center_idx)
    xyz = xyz.clone()
    xyz[:, 21:] = verts
    xyz[:, :21] = joints
    return xyz, pose, betas


# ------------------------------------------------------------------------------------------ a1
def head_forward(sd, dims, feat, img_metas, reference_joints, template, bps, anchor_xyz, anchor_idx, stages=None,
                 mano=None, neighbours=None):
    """`POEM_Generalized_Head.forward` (lib/models/heads/ptEmb_head.py:825-964).
    `neighbours` (test hook, not in the reference): (NB-1, 2, B, Q, K) int64 — the 32-NN index sets (self, cross) to use
    in blocks 1..NB-1 instead of running `knn`, so that a device run and the oracle can be compared on the same sets
    (32-NN selection is discontinuous in the coordinates).
    Returns all_coords_preds (NB,B,799,3) in metres; with `dims.parametric` (medium_MANO) the last block goes
    through the MANO tail and (coords, pred_pose (B,16,3), pred_shape (B,10)) is returned."""
    views = [int(v) for v in img_metas["cam_view_num"]]
    B = len(views)
    inp_w, inp_h = img_metas["inp_img_shape"]
    inp_res = torch.tensor([float(inp_w), float(inp_h)], device=feat.device)
    x = feature_volume(sd, feat, views, dims)
    centre = reference_joints[:, 9]                                 # (B,3)  ptEmb_head.py:873 (fixed joint 9)
    bps_world = bps[None] + centre[:, None]
    grid = project_bps(bps_world, img_metas["cam_intr"], img_metas["cam_extr"], views, inp_res)
    sampled = F.grid_sample(x, grid, align_corners=False).squeeze(-1)   # (BV,D,P)
    pt_feats = merge_views(sd, sampled, views)                       # (B,P,D)
    q_feats = sd["query_feat_embedding.weight"][None].expand(B, -1, -1)
    pt_xyz = (bps_world - centre[:, None]) / dims.radius
    q_xyz = ((centre[:, None] + template[None]) - centre[:, None]) / dims.radius
    if stages is not None:
        stages.update(x=x, grid=grid, sampled=sampled, pt_feats=pt_feats, pt_xyz=pt_xyz, q_xyz=q_xyz)
    anchors = (anchor_xyz, anchor_idx)
    xyz_all = []
    for i in range(dims.n_blocks):
        nb = None if (neighbours is None or i == 0) else (neighbours[i - 1][0], neighbours[i - 1][1])
        q_feats, q_xyz = metro_block(sd, i, dims, q_xyz, q_feats, pt_xyz, pt_feats, anchors, stages, nb)
        if dims.parametric and i == dims.n_blocks - 1:          # pt_metro_transformer.py:194-195
            if stages is not None:
                stages["tail.feats"], stages["tail.xyz_in"] = q_feats, q_xyz
            q_xyz, pose, betas = parametric_tail(sd, i, dims, q_feats, q_xyz, mano)
        xyz_all.append(q_xyz)
    coords = torch.nan_to_num(torch.stack(xyz_all))
    if stages is not None:
        stages["xyz_norm"] = coords
    if not dims.parametric:
        return coords * dims.radius + centre[None, :, None, :]
    # ptEmb_head.py:950-963: the MANO output of the last block is metric already, only the offset is added
    out = torch.cat([coords[:-1] * dims.radius, coords[-1:]]) + centre[None, :, None, :]
    return out, pose.reshape(-1, 16, 3), betas.reshape(-1, 10)


# ------------------------------------------------------------------------------------------ a17
def _conv_bn(sd, conv_key, bn_key, x, stride=1, relu=False):
    """Conv2d (+ bias when the checkpoint has one) + BatchNorm2d in eval mode when bn_key is given (+ ReLU)."""
    w = sd[conv_key + ".weight"]
    y = F.conv2d(x, w, sd.get(conv_key + ".bias"), stride=stride, padding=w.shape[-1] // 2)
    if bn_key is not None:
        y = F.batch_norm(y, sd[bn_key + ".running_mean"], sd[bn_key + ".running_var"], sd[bn_key + ".weight"],
                         sd[bn_key + ".bias"], training=False, eps=1e-5)
    return F.relu(y) if relu else y


def hrnet_stage4(sd, x_list, n_modules=3, prefix=""):
    """`HighResolutionNet.stage4` = n_modules x `HighResolutionModule.forward`
    (lib/models/backbones/hrnet.py:217-234; BasicBlock :38-67; fuse layers :177-207). sd keys relative to `stage4.`"""
    xs = list(x_list)
    nb = len(xs)
    for m in range(n_modules):
        m = f"{prefix}{m}"
        for b in range(nb):
            x = xs[b]
            for k in range(4):
                p = f"{m}.branches.{b}.{k}."
                t = _conv_bn(sd, p + "conv1", p + "bn1", x, relu=True)
                x = F.relu(_conv_bn(sd, p + "conv2", p + "bn2", t) + x)
            xs[b] = x
        fused = []
        for i in range(nb):
            y = None
            for j in range(nb):
                p = f"{m}.fuse_layers.{i}.{j}."
                if j == i:
                    t = xs[j]
                elif j > i:
                    t = F.interpolate(_conv_bn(sd, p + "0", p + "1", xs[j]), scale_factor=2 ** (j - i), mode="nearest")
                else:
                    t = xs[j]
                    for k in range(i - j):
                        t = _conv_bn(sd, p + f"{k}.0", p + f"{k}.1", t, stride=2, relu=(k != i - j - 1))
                y = t if y is None else y + t
            fused.append(F.relu(y))
        xs = fused
    return xs


def hrnet_forward(sd, img):
    """Whole `HighResolutionNet.forward` for the W40 yaml (lib/models/backbones/hrnet.py:385-420): stem :388-393,
    layer1 = 4 Bottlenecks :70-104, transitions :318-342, stages 2/3/4 with 1/4/3 modules. BatchNorm in eval mode."""
    x = _conv_bn(sd, "conv1", "bn1", img, stride=2, relu=True)
    x = _conv_bn(sd, "conv2", "bn2", x, stride=2, relu=True)
    for k in range(4):
        p = f"layer1.{k}."
        t = _conv_bn(sd, p + "conv1", p + "bn1", x, relu=True)
        t = _conv_bn(sd, p + "conv2", p + "bn2", t, relu=True)
        t = _conv_bn(sd, p + "conv3", p + "bn3", t)
        res = _conv_bn(sd, p + "downsample.0", p + "downsample.1", x) if (p + "downsample.0.weight") in sd else x
        x = F.relu(t + res)
    xs = [_conv_bn(sd, "transition1.0.0", "transition1.0.1", x, relu=True),
          _conv_bn(sd, "transition1.1.0.0", "transition1.1.0.1", x, stride=2, relu=True)]
    ys = hrnet_stage4(sd, xs, 1, prefix="stage2.")
    xs = ys + [_conv_bn(sd, "transition2.2.0.0", "transition2.2.0.1", ys[-1], stride=2, relu=True)]
    ys = hrnet_stage4(sd, xs, 4, prefix="stage3.")
    xs = ys + [_conv_bn(sd, "transition3.3.0.0", "transition3.3.0.1", ys[-1], stride=2, relu=True)]
    return hrnet_stage4(sd, xs, 3, prefix="stage4.")


def feat_decode(sd, feats):
    """`PtEmbedMultiviewStereoV2.feat_decode`, HRNet branch (lib/models/POEM.py:189-203): ConvBlocks are
    conv(bias) + BN + ReLU (lib/models/bricks/conv.py:4-44); `feat_in` is a bare 1x1 convolution."""
    x = feats[0]
    for i in range(3):
        p = f"feat_delayer.{i}."
        x = _conv_bn(sd, p + "conv", p + "norm", x, stride=2, relu=True) + feats[i + 1]
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return _conv_bn(sd, "feat_in.conv", None, x)


def image_features(sd, img):
    """images -> mlvl_feat: `extract_img_feat` + `feat_decode` (POEM.py:255-265); sd holds full-model keys."""
    feats = hrnet_forward({k[len("img_backbone."):]: v for k, v in sd.items() if k.startswith("img_backbone.")}, img)
    return feat_decode(sd, feats), feats


def uv_decode_heatmap(sd, feats, img_w=256, img_h=256):
    """`uv_decode` + `heatmap_stage`, HRNet branch (lib/models/POEM.py:205-229) with `integral_heatmap2d`
    (lib/models/integal_pose.py:196-220).  Returns (uv_px (BN,21,2), uv_hmap (BN,21,32,32))."""
    rev = list(reversed(feats))
    x = rev[0]
    for i in range(3):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = torch.cat((x, rev[i + 1]), dim=1)
        p = f"uv_delayer.{i}."
        x = _conv_bn(sd, p + "conv", p + "norm", x, relu=True)
    x = F.max_pool2d(x, kernel_size=2, stride=2)
    hmap = torch.sigmoid(_conv_bn(sd, "uv_out.conv", None, x))
    pdf = hmap.reshape(*hmap.shape[:2], -1)
    pdf = (pdf / (pdf.sum(dim=-1, keepdim=True) + 1e-6)).view_as(hmap)
    v_accu, u_accu = pdf.sum(dim=3), pdf.sum(dim=2)
    wv = torch.arange(v_accu.shape[-1], dtype=pdf.dtype, device=pdf.device) / v_accu.shape[-1]
    wu = torch.arange(u_accu.shape[-1], dtype=pdf.dtype, device=pdf.device) / u_accu.shape[-1]
    uv = torch.cat([(u_accu * wu).sum(-1, keepdim=True), (v_accu * wv).sum(-1, keepdim=True)], dim=-1)
    return uv * torch.tensor([float(img_w), float(img_h)], device=uv.device), hmap


def triangulate_dlt(uv_px, cam_intr, cam_extr, view_counts):
    """Reference joints by DLT, per sample over its own views (POEM.py:284-299 calling
    lib/utils/triangulation.py:5-45): M = K * inv(cam_extr)[:3]; A rows u*M[2]-M[0], v*M[2]-M[1]; X = last row of VT."""
    out = []
    T = torch.linalg.inv(cam_extr.view(-1, 4, 4))
    K = cam_intr.view(-1, 3, 3)
    start = 0
    for n in view_counts:
        M = torch.matmul(K[start:start + n], T[start:start + n, :3, :])          # (n,3,4)
        kp = uv_px[start:start + n].permute(1, 0, 2).unsqueeze(3)                 # (J,n,2,1)
        A = (kp * M[None, :, 2:3, :] - M[None, :, :2, :]).reshape(kp.shape[0], -1, 4)
        VT = torch.linalg.svd(A)[2]
        out.append(VT[:, -1, :3] / (VT[:, -1, 3:] + 1e-7))
        start += n
    return torch.stack(out)


def model_forward(sd, dims, batch, template, bps, anchor_xyz, anchor_idx, data_center_idx=0):
    """`PtEmbedMultiviewStereoV2._forward_impl(mode="test")` (lib/models/POEM.py:251-333) on full-model keys:
    backbone -> feat_decode -> heatmap_stage -> per-sample DLT (or the given joints when every sample is single-view)
    -> head.  Returns the reference's prediction dict (evaluation keys)."""
    img = batch["image"].reshape(-1, *batch["image"].shape[-3:])
    views = [int(v) for v in batch["cam_view_num"]]
    H, W = img.shape[-2:]
    mlvl_feat, feats = image_features(sd, img)
    uv, _ = uv_decode_heatmap(sd, feats, W, H)
    intr, extr = batch["target_cam_intr"].reshape(-1, 3, 3), batch["target_cam_extr"].reshape(-1, 4, 4)
    if img.shape[0] == len(views):
        ref_joints = batch["master_joints_3d"].reshape(-1, 21, 3)
    else:
        ref_joints = triangulate_dlt(uv, intr, extr, views)
    metas = {"inp_img_shape": (H, W), "cam_intr": intr, "cam_extr": extr, "master_id": batch["master_id"],
             "cam_view_num": views}
    head_sd = {k[len("ptEmb_head."):]: v for k, v in sd.items() if k.startswith("ptEmb_head.")}
    coords = head_forward(head_sd, dims, mlvl_feat, metas, ref_joints, template, bps, anchor_xyz, anchor_idx)
    pj, pv = coords[-1, :, :21], coords[-1, :, 21:]
    c = pj[:, data_center_idx].unsqueeze(1)                          # POEM.py:328 (DATA_PRESET.CENTER_IDX = 0)
    return {"all_coords_preds": coords, "pred_joints_3d": pj, "pred_verts_3d": pv, "pred_joints_3d_rel": pj - c,
            "pred_verts_3d_rel": pv - c, "pred_joints_uv": uv, "pred_ref_joints_3d": ref_joints}


def pa_align(gt, pred):
    """`PAEval.align_w_scale` (lib/metrics/pa_eval.py:103-124) for one sample, numpy fp32 like the reference; the
    orthogonal Procrustes step is scipy's published algorithm (scipy.linalg.orthogonal_procrustes: u, w, vt =
    svd(A^T B); R = u vt; scale = sum(w))."""
    import numpy as np
    mtx1, mtx2 = np.asarray(gt, dtype=np.float32), np.asarray(pred, dtype=np.float32)
    t1, t2 = mtx1.mean(0), mtx2.mean(0)
    a, b = mtx1 - t1, mtx2 - t2
    s1 = np.linalg.norm(a) + 1e-8
    a = a / s1
    s2 = np.linalg.norm(b) + 1e-8
    b = b / s2
    u, w, vt = np.linalg.svd(a.T.dot(b))
    R, s = u.dot(vt), w.sum()
    return np.dot(b, R.T) * s * s1 + t1


def pa_distances(gt, pred):
    """(B,N,3) x2 -> (B,2): [aligned mean distance, raw mean distance] as `PAEval.feed` / `get_dist` (pa_eval.py:41-66)."""
    import numpy as np
    gt, pred = np.asarray(gt, dtype=np.float32), np.asarray(pred, dtype=np.float32)
    al = np.stack([pa_align(g, p) for g, p in zip(gt, pred)])
    return np.stack([np.linalg.norm(al - gt, axis=2).mean(1), np.linalg.norm(pred - gt, axis=2).mean(1)], axis=1)


# ------------------------------------------------------------------------------------------ f3 (training loss; groundwork)
OPENPOSE_TIP_VERTS = (744, 320, 443, 555, 672)      # lib/utils/misc.py:76-82 (CONST.MANO_KPID_2_VERTICES)
OPENPOSE_ORDER = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)


def mano_to_openpose(j_regressor, verts):
    """lib/utils/transform.py:836-872: 16 regressed joints + the reference's own 5 tip vertices, re-ordered."""
    j = torch.matmul(j_regressor, verts)
    return torch.cat([j, verts[:, list(OPENPOSE_TIP_VERTS)]], dim=1)[:, list(OPENPOSE_ORDER)]


def _project_to_views(points, T_m2c, K, view_counts):
    """`batch_cam_extr_transf` + `batch_cam_intr_projection` (lib/utils/transform.py:898-930) of each sample's master-frame
    points into all of its views: (B,P,3) -> (sum V, P, 2)."""
    out, start = [], 0
    for b, n in enumerate(view_counts):
        T, Kb = T_m2c[start:start + n], K[start:start + n]
        pc = torch.einsum("nrc,pc->npr", T[:, :3, :3], points[b]) + T[:, None, :3, 3]
        q = torch.einsum("nrc,npc->npr", Kb, pc)
        z = q[..., 2:]
        z = torch.where(z.abs() < 1e-7, torch.full_like(z, 1e-7), z)
        out.append(q[..., :2] / z)
        start += n
    return torch.cat(out)


def compute_loss(preds, gt, weights, j_regressor, parametric=False, transformer_center_idx=9, num_joints=21):
    """`PtEmbedMultiviewStereoV2.compute_loss` (lib/models/POEM.py:363-466) with the release loss types (joints l2,
    vertices l1, parameters l2; config/release/train_*.yaml LOSS).  `weights`: dict with the cfg.LOSS weights.
    Groundwork for SURVEY §8f row f3 (training path): pinned against the reference method in tests/golden/loss_*.npz;
    no product kernel consumes it yet."""
    coords = preds["all_coords_preds"]
    views = [int(v) for v in gt["cam_view_num"]]
    H, W = gt["image"].shape[-2:]
    img_scale = math.sqrt(float(W ** 2 + H ** 2))
    jg, vg = gt["master_joints_3d"].reshape(-1, 21, 3), gt["master_verts_3d"].reshape(-1, 778, 3)
    out = {}
    d = (preds["pred_joints_uv"] - gt["target_joints_2d"]) / img_scale
    out["loss_heatmap_joints"] = (d ** 2).sum(2).mean()
    loss = weights["HEATMAP_JOINTS_WEIGHT"] * out["loss_heatmap_joints"]
    T = torch.linalg.inv(gt["target_cam_extr"].reshape(-1, 4, 4))
    K = gt["target_cam_intr"].reshape(-1, 3, 3)
    pj, pv = coords[-1, :, :num_joints], coords[-1, :, num_joints:]
    out["loss_3d_joints_from_mesh"] = F.mse_loss(mano_to_openpose(j_regressor, pv), mano_to_openpose(j_regressor, vg))
    out["loss_3d_joints"] = F.mse_loss(pj, jg)
    recon = weights["JOINTS_LOSS_WEIGHT"] * (out["loss_3d_joints"] + out["loss_3d_joints_from_mesh"])
    if parametric:
        c = jg[:, transformer_center_idx:transformer_center_idx + 1]
        out["loss_3d_verts"] = F.l1_loss(pv - c, vg - c)
    else:
        out["loss_3d_verts"] = F.l1_loss(pv, vg)
    recon = recon + weights["VERTICES_LOSS_WEIGHT"] * out["loss_3d_verts"]

    def proj_loss(points, target_2d):                      # loss_proj_to_multicam, POEM.py:335-361
        off = torch.clamp(_project_to_views(points, T, K, views) - target_2d, min=-0.5 * img_scale,
                          max=0.5 * img_scale) / img_scale
        return (off ** 2).sum(2).mean()
    if weights.get("JOINTS_2D_LOSS_WEIGHT", 0.0) != 0:
        out["loss_2d_joints"] = proj_loss(pj, gt["target_joints_2d"])
        recon = recon + weights["JOINTS_2D_LOSS_WEIGHT"] * out["loss_2d_joints"]
    if weights.get("VERTICES_2D_LOSS_WEIGHT", 0.0) != 0:
        out["loss_2d_verts"] = proj_loss(pv, _project_to_views(vg, T, K, views))
        recon = recon + weights["VERTICES_2D_LOSS_WEIGHT"] * out["loss_2d_verts"]
    if parametric:
        first = [sum(views[:j]) for j in range(len(views))]
        out["loss_pose"] = F.mse_loss(preds["pred_pose"], gt["mano_pose"][first])
        out["loss_shape"] = F.mse_loss(preds["pred_shape"], gt["mano_shape"][first])
        recon = recon + weights.get("POSE_LOSS_WEIGHT", 0.001) * out["loss_pose"] \
            + weights.get("SHAPE_LOSS_WEIGHT", 0.0005) * out["loss_shape"]
    out["loss_recon"] = recon
    out["loss"] = loss + recon
    return out
