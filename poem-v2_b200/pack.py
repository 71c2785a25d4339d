"""Weight packing: reference state dict -> kernel layout (`PoemWeights` of include/poem_b200.h).

Host-side, one-off (re-run only when parameters change):
  * matrices -> op16 (fp16) [out, in] row-major, biases / LayerNorm / tiny matrices -> fp32
  * algebraic folds done in fp64 before rounding to fp16:
      - BPS-token projections: `embedding` composed into K/V of both BERT attentions and (through
        query_cross_attn.fc1) into k'/v' of the vector cross-attention, so the 4096 tokens are projected once
        per point instead of once per (query, neighbour) pair  (reference point_transformers.py:136-141,
        pt_metro_transformer.py:180-181)
      - query_self_attn.fc1 composed into w_qs / w_ks / w_vs (point_transformers.py:86-87)
      - fc_gamma.0 distributed over (q - k + pos): W_g1 is composed into the q and k projections and with fc_delta.2
        (point_transformers.py:90-91,144-147), so pos and the gamma hidden layer are independent GEMMs of the same input
  * positional table: adapt_pos3d(SinePositionalEncoding3D(N views)) + bias for N = 1..max_views — it depends
    only on the weights and the view count, not on the input (ptEmb_head.py:842-860,
    layers/petr_transformer.py:434-469)
"""
import math

import torch

from . import _native as nat
from .config import HeadDims


def sine_pos_3d(n_views, h, w, num_feats, normalize=True, temperature=10000.0, scale=2 * math.pi, eps=1e-6):
    """SinePositionalEncoding3D on an all-false mask -> (n_views, 3*num_feats, h, w) fp32.
    Channel order [view | y | x]; inside each third: all sines (even dims) then all cosines (odd dims)."""
    def axis(n):
        e = torch.arange(1, n + 1, dtype=torch.float32)
        return e / (e[-1] + eps) * scale if normalize else e
    i = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_feats)

    def enc(e):                                   # (L,) -> (L, num_feats)
        p = e[:, None] / dim_t
        return torch.cat((p[:, 0::2].sin(), p[:, 1::2].cos()), dim=1)
    en, ey, ex = enc(axis(n_views)), enc(axis(h)), enc(axis(w))
    out = torch.empty(n_views, 3 * num_feats, h, w)
    out[:, :num_feats] = en[:, :, None, None].expand(-1, -1, h, w)
    out[:, num_feats:2 * num_feats] = ey.t()[None, :, :, None].expand(n_views, -1, -1, w)
    out[:, 2 * num_feats:] = ex.t()[None, :, None, :].expand(n_views, -1, h, -1)
    return out


def bps_spatial_chunks(bps, radius, grow=1e-4):
    """k-d partition of the basis points into compact 32-point chunks: returns the permutation (chunk-major order,
    int32 [P]) and each chunk's bounding box (fp32 [P/32, 6] = min xyz, max xyz; normalised units, grown by `grow` to
    absorb the per-sample rounding of ((bps + c) - c) / r).  P must be 32 * 2^k."""
    x = bps.detach().cpu().double() / radius
    order = torch.arange(x.shape[0])

    def split(idx):
        if idx.numel() <= 32:
            return [idx]
        pts = x[idx]
        axis = int((pts.max(dim=0).values - pts.min(dim=0).values).argmax())
        srt = idx[torch.argsort(pts[:, axis], stable=True)]
        half = srt.numel() // 2
        return split(srt[:half]) + split(srt[half:])
    leaves = split(order)
    assert all(l.numel() == 32 for l in leaves), "basis point count must be 32 * 2^k"
    perm = torch.cat(leaves)
    pts = x[perm].reshape(-1, 32, 3)
    boxes = torch.cat([pts.min(dim=1).values - grow, pts.max(dim=1).values + grow], dim=1)
    return perm.to(torch.int32), boxes.float()


class PackedWeights:
    """Owns the device tensors referenced by the ctypes `PoemWeights` struct."""

    def __init__(self, sd, dims: HeadDims, device, template_xyz, bps, anchor_xyz, anchor_idx, max_views=10):
        self.dims = dims
        self.device = torch.device(device)
        self.max_views = max_views
        self._keep = []
        self.struct = nat.PoemWeights()
        D, C = dims.embed_dims, dims.in_channels
        f64 = {k: v.detach().to("cpu", torch.float64) for k, v in sd.items()}
        W = self.struct

        W.input_proj = self._linear(f64["input_proj.weight"].reshape(D, C), f64["input_proj.bias"])
        W.pos_table = self._f32(self._pos_table(f64, dims, max_views))
        W.merge0a = self._linear(f64["merge_net_feature.0.0.weight"], f64["merge_net_feature.0.0.bias"])
        W.merge0b = self._linear(f64["merge_net_feature.0.2.weight"], f64["merge_net_feature.0.2.bias"])
        W.merge1a = self._linear(f64["merge_net_feature.1.0.weight"], f64["merge_net_feature.1.0.bias"])
        W.merge1b = self._linear(f64["merge_net_feature.1.2.weight"], f64["merge_net_feature.1.2.bias"])
        W.query_embed = self._f32(f64["query_feat_embedding.weight"])
        W.bps = self._f32(bps.reshape(-1, 3))
        W.anchor_xyz = self._f32(anchor_xyz.reshape(-1, 3))
        W.anchor_idx = self._i32(anchor_idx.reshape(-1))
        assert template_xyz.shape == (dims.n_query, 3)
        W.template_xyz = self._f32(template_xyz)
        perm, boxes = bps_spatial_chunks(bps.reshape(-1, 3), dims.radius)
        W.bps_perm = self._i32(perm)
        W.bps_chunk_box = self._f32(boxes)
        for i in range(dims.n_blocks):
            self._pack_block(W.blocks[i], f64, f"transformer.pt_metro_encoder.{i}.")

    # ------------------------------------------------------------------ helpers
    def _dev(self, t, dtype):
        t = t.to(dtype).contiguous().to(self.device)
        self._keep.append(t)
        return t.data_ptr()

    def _f32(self, t):
        return self._dev(t, torch.float32)

    def _i32(self, t):
        return self._dev(t, torch.int32)

    def _op16(self, t):
        t = nat.to_op16(t).contiguous().to(self.device)
        self._keep.append(t)
        return t.data_ptr()

    def _linear(self, w, b=None):
        return nat.PoemLinear(self._op16(w), None if b is None else self._f32(b))

    @staticmethod
    def _pos_table(f64, dims, max_views):
        D = dims.embed_dims
        wa = f64["adapt_pos3d.weight"].reshape(D, -1)
        ba = f64["adapt_pos3d.bias"]
        rows = []
        for n in range(1, max_views + 1):
            s = sine_pos_3d(n, dims.feat_hw, dims.feat_hw, dims.pos_feats, dims.pos_normalize).double()
            t = torch.einsum("dc,nchw->nhwd", wa, s) + ba          # (n, h, w, D)
            rows.append(t.reshape(n, dims.feat_hw * dims.feat_hw, D))
        return torch.cat(rows, dim=0)

    def _vec_attn(self, dst, f64, p):
        dst.wd1 = self._f32(f64[p + "fc_delta.0.weight"])
        dst.bd1 = self._f32(f64[p + "fc_delta.0.bias"])
        dst.delta2 = self._linear(f64[p + "fc_delta.2.weight"], f64[p + "fc_delta.2.bias"])
        dst.gamma1_delta2 = self._linear(f64[p + "fc_gamma.0.weight"] @ f64[p + "fc_delta.2.weight"], None)
        dst.gamma2 = self._linear(f64[p + "fc_gamma.2.weight"], f64[p + "fc_gamma.2.bias"])
        dst.fc2 = self._linear(f64[p + "fc2.weight"], f64[p + "fc2.bias"])

    def _pack_block(self, blk, f64, p):
        We, be = f64[p + "embedding.weight"], f64[p + "embedding.bias"]
        blk.embedding = self._linear(We, be)
        e = p + "encoder."

        def through_embedding(w, b):          # y = w (We x + be) + b
            return w @ We, w @ be + (b if b is not None else 0.0)
        c = e + "vec_attn.query_cross_attn."
        W1c, b1c = f64[c + "fc1.weight"], f64[c + "fc1.bias"]
        # fc_gamma.0 distributed over (q - k + pos): kt = W_g1 k, qt = W_g1 q + W_g1 b_d2 + b_g1 (include/poem_b200.h)
        Wg1c = f64[c + "fc_gamma.0.weight"]
        qt_bias_c = Wg1c @ f64[c + "fc_delta.2.bias"] + f64[c + "fc_gamma.0.bias"]
        parts = [
            through_embedding(f64[e + "attn.self.key.weight"], f64[e + "attn.self.key.bias"]),
            through_embedding(f64[e + "cross_attn.self.key.weight"], f64[e + "cross_attn.self.key.bias"]),
            through_embedding(Wg1c @ f64[c + "w_ks.weight"] @ W1c, Wg1c @ f64[c + "w_ks.weight"] @ b1c),
            through_embedding(f64[c + "w_vs.weight"] @ W1c, f64[c + "w_vs.weight"] @ b1c),
            through_embedding(f64[e + "attn.self.value.weight"], f64[e + "attn.self.value.bias"]),
            through_embedding(f64[e + "cross_attn.self.value.weight"], f64[e + "cross_attn.self.value.bias"]),
        ]
        blk.pt_proj = self._linear(torch.cat([w for w, _ in parts]), torch.cat([b for _, b in parts]))
        blk.q1 = self._linear(f64[e + "attn.self.query.weight"], f64[e + "attn.self.query.bias"])
        blk.o1 = self._linear(f64[e + "attn.output.dense.weight"], f64[e + "attn.output.dense.bias"])
        blk.ln1_g = self._f32(f64[e + "attn.output.LayerNorm.weight"])
        blk.ln1_b = self._f32(f64[e + "attn.output.LayerNorm.bias"])
        blk.q2 = self._linear(f64[e + "cross_attn.self.query.weight"], f64[e + "cross_attn.self.query.bias"])
        blk.o2 = self._linear(f64[e + "cross_attn.output.dense.weight"], f64[e + "cross_attn.output.dense.bias"])
        blk.ln2_g = self._f32(f64[e + "cross_attn.output.LayerNorm.weight"])
        blk.ln2_b = self._f32(f64[e + "cross_attn.output.LayerNorm.bias"])
        s = e + "vec_attn.query_self_attn."
        W1s, b1s = f64[s + "fc1.weight"], f64[s + "fc1.bias"]
        Wg1s = f64[s + "fc_gamma.0.weight"]
        qt_bias_s = Wg1s @ f64[s + "fc_delta.2.bias"] + f64[s + "fc_gamma.0.bias"]
        pre = {"w_qs": Wg1s, "w_ks": Wg1s, "w_vs": torch.eye(Wg1s.shape[0], dtype=torch.float64)}
        qkv_w = torch.cat([pre[n] @ f64[s + n + ".weight"] @ W1s for n in ("w_qs", "w_ks", "w_vs")])
        qkv_b = torch.cat([pre[n] @ f64[s + n + ".weight"] @ b1s + (qt_bias_s if n == "w_qs" else 0.0)
                           for n in ("w_qs", "w_ks", "w_vs")])
        blk.self_qkv = self._linear(qkv_w, qkv_b)
        self._vec_attn(blk.self_attn, f64, s)
        blk.cross_q = self._linear(Wg1c @ f64[c + "w_qs.weight"], qt_bias_c)
        self._vec_attn(blk.cross_attn, f64, c)
        blk.reg1 = self._linear(f64[e + "vec_attn.reg_branch.0.weight"], f64[e + "vec_attn.reg_branch.0.bias"])
        blk.reg2_w = self._f32(f64[e + "vec_attn.reg_branch.2.weight"])
        blk.reg2_b = self._f32(f64[e + "vec_attn.reg_branch.2.bias"])
        blk.ffn1 = self._linear(f64[e + "intermediate.dense.weight"], f64[e + "intermediate.dense.bias"])
        blk.ffn2 = self._linear(f64[e + "output.dense.weight"], f64[e + "output.dense.bias"])
        blk.ln3_g = self._f32(f64[e + "output.LayerNorm.weight"])
        blk.ln3_b = self._f32(f64[e + "output.LayerNorm.bias"])


MANO_KEYS = ("v_template", "shapedirs", "posedirs", "J_regressor", "weights")


def mano_zero_pose_template(mano, center_idx=9):
    """(799,3) joints ‖ vertices of the MANO layer at zero pose / zero shape, centred on joint `center_idx` — what the
    reference head recomputes on every forward (ptEmb_head.py:885-891).  At zero pose the skinning is the identity:
    vertices = v_template, the 16 joints = J_regressor · v_template, plus the 5 fingertip vertices, re-ordered."""
    from .params import MANO_JOINT_ORDER, MANO_TIP_VERTS
    v = mano["v_template"].detach().to("cpu", torch.float64).reshape(778, 3)
    j16 = mano["J_regressor"].detach().to("cpu", torch.float64).reshape(16, 778) @ v
    j21 = torch.cat([j16, v[list(MANO_TIP_VERTS)]])[list(MANO_JOINT_ORDER)]
    c = j21[center_idx:center_idx + 1]
    return torch.cat([j21 - c, v - c]).float()


class PackedManoTail:
    """Device tensors behind the ctypes `PoemManoTail` struct: the last block's `flat_verts` / `mano_linear` (fp32, as
    in the reference) and the MANO model parameters with the blend-coefficient axis first."""

    def __init__(self, sd, dims: HeadDims, mano, device):
        self._keep = []
        p = f"transformer.pt_metro_encoder.{dims.n_blocks - 1}."
        dev = torch.device(device)

        def f32(t):
            t = t.detach().to(torch.float32).contiguous().to(dev)
            self._keep.append(t)
            return t.data_ptr()
        m = nat.PoemManoTail()
        m.flat_w = f32(sd[p + "flat_verts.weight"].reshape(-1))
        m.flat_b = f32(sd[p + "flat_verts.bias"].reshape(-1))
        m.lin_w = f32(sd[p + "mano_linear.weight"].reshape(106, dims.embed_dims))
        m.lin_b = f32(sd[p + "mano_linear.bias"].reshape(-1))
        m.v_template = f32(mano["v_template"].reshape(778, 3))
        m.shapedirs = f32(mano["shapedirs"].reshape(778 * 3, 10).t())
        m.posedirs = f32(mano["posedirs"].reshape(778 * 3, 135).t())
        m.j_regressor = f32(mano["J_regressor"].reshape(16, 778))
        m.skin_weights = f32(mano["weights"].reshape(778, 16))
        self.struct = m
