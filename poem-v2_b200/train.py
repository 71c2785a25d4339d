"""Training path of the decoder head: forward with saved activations and the hand-written
backward of `POEM_Generalized_Head.forward` (reference lib/models/heads/ptEmb_head.py:825-964,
lib/models/bricks/pt_metro_transformer.py:34-200, lib/models/bricks/point_transformers.py:70-156), i.e. what
`loss.backward()` does in the reference's `scripts/train_ddp.py:103-104`.

All arithmetic runs in libpoem_train.so (include/poem_train.h): fp32 activations in HBM, TF32 tcgen05 GEMMs for every
Linear / 1x1 conv / attention product (forward, dgrad and wgrad read the same row-major tensors through K-major or
MN-major operand descriptors, no transposes), SIMT kernels for the rest; the 32-NN search is the inference library's
`poem_knn32`.  This module is the schedule only — which kernel runs on which buffer, forward and in reverse — the part
that is Python (torch.autograd) in the reference as well.  torch is used for device memory (`torch.empty`, `clone`,
`zero_`) and the stream.  Everything is kept: the BERT attention probabilities (B*h*799*4096 fp32 per layer) and the
per-edge tensors of the vector attention (B*799*32*D per layer) stay in HBM between forward and backward
(~40 GB at POEM-medium, batch 32: sized for the 180 GB of a B200) instead of the reference's recompute
(`cp.checkpoint`, point_transformers.py:63,119).

Dropout (TRANSFORMER.DROPOUT, 0.1 in the release configs: on both embedding outputs, after the two attention output
projections and the FFN output projection, and on the attention probabilities) is counter-based — masks are regenerated
in the backward from a device seed, nothing is stored.  The parametric MANO tail of medium_MANO (flat_verts, mano_linear, 6-D rotations ->
axis-angle, MANO skinning) has its backward too (csrc/mano_bwd.cuh).  Widths: D = 128 / 256 / 512 (small, medium, large);
D = 1024 is refused like in the inference path.
"""
import math

import numpy as np
import torch

from . import _native as nat
from . import _train_native as tn
from . import params as _params
from .config import HeadDims
from .pack import sine_pos_3d

NBR = 32


def _on_own_device(fn):
    """The primitives launch on torch's CURRENT device / stream: run the method with the trainer's device current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = self.dev if hasattr(self, "dev") else self.tr.dev
        if dev.type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped
# operand rounding of the gradient GEMMs (poem_tr_gemm round_ops).  3 (default): both operands rounded to nearest, like
# the forward GEMMs (those always round: the 1e-3 bound on the coordinates needs it).  POEM_TR_GRAD_ROUND=0 lets the
# tensor core truncate the gradient GEMMs' operands (a 2^-11 relative shrink of every product term) and drops their
# shared-memory rounding pass: backward 82.8 -> 73.5 ms at batch 32, gradient error vs the oracle 3e-3 -> 4.9e-3 median,
# 2.6e-2 -> 5.4e-2 on the conv biases.
GRAD_ROUND = int(__import__("os").environ.get("POEM_TR_GRAD_ROUND", "3"))


class HeadTrainer:
    """Holds fp32 master parameters `p[name]` and gradient buffers `g[name]` (reference state-dict names) on one GPU.

        tr = HeadTrainer(dims, state_dict, template)
        coords = tr.forward(mlvl_feat, img_metas, reference_joints)      # (NB, B, 799, 3), metres
        dfeat = tr.backward(dcoords)                                     # d loss / d mlvl_feat ; tr.g[name] += d loss / d p
    """

    def __init__(self, dims: HeadDims, state_dict, template, assets=None, device="cuda", dropout=None, mano=None):
        """dropout: probability of the BERT layers' hidden / attention-probability dropout (None: dims.dropout, i.e.
        TRANSFORMER.DROPOUT of the config; 0 = the eval-mode arithmetic the gradient goldens are pinned on).
        mano: MANO model parameters (v_template, shapedirs, posedirs, J_regressor, weights) for a PARAMETRIC_OUTPUT head."""
        if dims.parametric and mano is None:
            raise ValueError("parametric (medium_MANO) head: pass the MANO model parameters, mano={v_template, shapedirs, ...}")
        if dims.embed_dims > 512 or dims.n_neighbor != NBR:
            raise ValueError("training path: embed_dims <= 512 and 32 neighbours")
        self.dims, self.dev = dims, torch.device(device)
        tn.load()
        self.lib = nat.load()
        # parameters and gradients are views into ONE flat buffer each (tensor starts padded to 16 bytes for TMA): one
        # fill zeroes the gradients, one launch clips / steps every tensor, a block's gradients are one contiguous
        # all-reduce bucket (order of live_param_shapes: head stage, query embedding, block 0, 1, ...)
        live = _params.live_param_shapes(dims)
        offs, total = {}, 0
        for k, shp in live.items():
            offs[k] = total
            total += (int(np.prod(shp)) + 3) // 4 * 4
        self.p_flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.g_flat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.pr_flat = torch.zeros(total, dtype=torch.float32, device=self.dev)   # TF32-rounded operand copy of the weights
        self.p, self.g, self.pr = {}, {}, {}
        for k, shp in live.items():
            n = int(np.prod(shp))
            self.p[k] = self.p_flat[offs[k]:offs[k] + n].view(*shp)
            self.g[k] = self.g_flat[offs[k]:offs[k] + n].view(*shp)
            self.pr[k] = self.pr_flat[offs[k]:offs[k] + n].view(*shp)
            self.p[k].copy_(state_dict[k].detach().to(self.dev, torch.float32).reshape(shp))
        keys = list(live)
        self.seg_off = torch.tensor([offs[k] for k in keys], dtype=torch.int64, device=self.dev)
        self.seg_len = torch.tensor([int(np.prod(live[k])) for k in keys], dtype=torch.int64, device=self.dev)
        first = {}                                   # bucket name -> (start, end) in the flat buffers
        for k in keys:
            b = k.split(".")[2] if k.startswith("transformer.pt_metro_encoder.") else "head"
            lo, hi = first.get(b, (offs[k], offs[k]))
            first[b] = (min(lo, offs[k]), max(hi, offs[k] + (int(np.prod(live[k])) + 3) // 4 * 4))
        self.buckets = first
        self._const = {}
        bps, a_xyz, a_idx = assets if assets is not None else _params.load_assets()
        self.bps = bps.to(self.dev, torch.float32).contiguous()
        self.anchor_xyz = a_xyz.to(self.dev, torch.float32).contiguous()
        self.anchor_idx = a_idx.to(self.dev, torch.int32).contiguous()
        self.template = torch.as_tensor(template, dtype=torch.float32).reshape(dims.n_query, 3).to(self.dev).contiguous()
        self.mano = None
        if dims.parametric:                       # kernel layout: blend axis first (as pack.PackedManoTail)
            f32 = lambda t: t.detach().to(self.dev, torch.float32).contiguous()  # noqa: E731
            self.mano = dict(v_template=f32(mano["v_template"].reshape(778 * 3)),
                             shapedirs=f32(mano["shapedirs"].reshape(778 * 3, 10).t()),
                             posedirs=f32(mano["posedirs"].reshape(778 * 3, 135).t()),
                             j_regressor=f32(mano["J_regressor"].reshape(16, 778)),
                             skin_weights=f32(mano["weights"].reshape(778, 16)))
        self.pred_pose = self.pred_shape = None
        self.tape = None
        self.last_neighbours = None
        self.p_drop = float(dims.dropout if dropout is None else dropout)
        if not 0.0 <= self.p_drop < 1.0:
            raise ValueError("dropout probability must be in [0, 1)")
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)    # device scalar: bumped every forward
        self.drop_sites = {}                                                   # name -> (site id, shape) of the last forward
        self._site = 0

    def manual_seed(self, seed):
        self.seed_dev.fill_(int(seed))

    def dropout_(self, x, name=None, site=None):
        """In place x <- keep ? x / (1 - p) : 0.  Forward: a new site (recorded under `name`); backward: pass the site."""
        if self.p_drop == 0.0:
            return None
        if site is None:
            site = self._site
            self._site += 1
            if name is not None:
                self.drop_sites[name] = (site, tuple(x.shape))
        tn.call("poem_tr_dropout", x, x, x.numel(), self.p_drop, self.seed_dev, site)
        return site

    def dropout_mask(self, name):
        """keep / (1 - p) of a site of the last forward, regenerated (tests: the oracle is run on the same masks)."""
        site, shape = self.drop_sites[name]
        m = torch.ones(*shape, dtype=torch.float32, device=self.dev)
        tn.call("poem_tr_dropout", m, m, m.numel(), self.p_drop, self.seed_dev, site)
        return m

    # ------------------------------------------------------------------------------------------ small helpers
    def new(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.dev)

    def zero_grad(self):
        self.g_flat.zero_()

    def lin(self, x, w, b=None, out=None, acc=False, relu=False, x_clean=False, round_out=False):
        """y (+)= x W^T + b   (relu: y = max(., 0) in the GEMM epilogue).  x_clean: x was stored TF32-rounded by its
        producer (no rounding pass for it); round_out: y is only ever a GEMM operand again, store it rounded."""
        M, K = x.shape
        W = self.pr[w]                                # pre-rounded copy of the weight
        N = W.shape[0]
        y = out if out is not None else self.new(M, N)
        tn.gemm(x, W, y, M, N, K, bias=self.p[b] if b else None, accumulate=acc, relu=relu, round_ops=0 if x_clean else 1,
                round_out=round_out)
        return y

    def lin_bwd(self, dy, x, w, b=None, need_dx=True, out=None, acc=False, relu_in=False, dy_clean=False, x_clean=False,
                round_out=False):
        """g[w] += dy^T x ; g[b] += colsum(dy) ; returns dx (+)= dy W.  relu_in: x is a ReLU output and the gradient
        w.r.t. the ReLU's input is wanted (dx = 0 where x <= 0, applied in the GEMM epilogue).  dy_clean / x_clean /
        round_out as in `lin`."""
        M, N = dy.shape
        K = x.shape[1]
        if b:
            tn.call("poem_tr_colsum", dy, N, M, N, self.g[b])
        ra, rb = (0 if dy_clean else 1), (0 if x_clean else 2)
        tn.gemm(dy, x, self.g[w], N, K, M, a_mn=True, b_mn=True, accumulate=True, round_ops=GRAD_ROUND & (ra | rb))
        if not need_dx:
            return None
        dx = out if out is not None else self.new(M, K)
        tn.gemm(dy, self.pr[w], dx, M, K, N, b_mn=True, accumulate=acc, relu_mask=x if relu_in else None,
                round_ops=GRAD_ROUND & ra, round_out=round_out)
        return dx

    def ln(self, x, res, pre):
        M, D = x.shape
        y, xhat, rstd = self.new(M, D), self.new(M, D), self.new(M)
        tn.call("poem_tr_layernorm", x, res, self.p[pre + ".weight"], self.p[pre + ".bias"], 1e-12, y, xhat, rstd, M, D)
        return y, xhat, rstd

    def ln_bwd(self, dy, xhat, rstd, pre):
        M, D = dy.shape
        dx = self.new(M, D)
        tn.call("poem_tr_layernorm_bwd", dy, xhat, rstd, self.p[pre + ".weight"], dx, self.g[pre + ".weight"],
                self.g[pre + ".bias"], M, D)
        return dx

    # ------------------------------------------------------------------------------------------ BERT cross-attention
    def _attn_strides(self, B, Lq, Lk, D, H):
        hd = D // H
        return dict(batch=(H, B)), (hd, Lq * D), (hd, Lk * D), (Lq * Lk, H * Lq * Lk)

    def bert_fwd(self, hid, enc, pre, B, Lq, Lk, enc_clean=True):
        D, H = hid.shape[1], self.dims.n_heads
        hd = D // H
        kw, sq, sk, sp = self._attn_strides(B, Lq, Lk, D, H)
        # Q, K, V, P, ctx are only ever GEMM operands: stored TF32-rounded by their producers, no rounding pass downstream
        Q = self.lin(hid, pre + ".self.query.weight", pre + ".self.query.bias", round_out=True)
        K = self.lin(enc, pre + ".self.key.weight", pre + ".self.key.bias", x_clean=enc_clean, round_out=True)     # enc = ke
        V = self.lin(enc, pre + ".self.value.weight", pre + ".self.value.bias", x_clean=enc_clean, round_out=True)
        P = self.new(B, H, Lq, Lk)
        tn.gemm(Q, K, P, Lq, Lk, hd, lda=D, ldb=D, ldc=Lk, a_strides=sq, b_strides=sk, c_strides=sp, round_ops=0, **kw)
        Pd, site_p = None, 0
        if self.p_drop > 0.0:                       # attention-probability dropout: P (kept for the backward) and its dropped copy
            Pd, site_p = self.new(B, H, Lq, Lk), self._site
            self._site += 1
            self.drop_sites[pre + ".probs"] = (site_p, (B, H, Lq, Lk))
        tn.call("poem_tr_softmax_rows", P, B * H * Lq, Lk, 1.0 / math.sqrt(hd), Pd, self.p_drop, self.seed_dev, site_p)
        ctx = self.new(B * Lq, D)
        tn.gemm(P if Pd is None else Pd, V, ctx, Lq, hd, Lk, b_mn=True, lda=Lk, ldb=D, ldc=D, a_strides=sp, b_strides=sk,
                c_strides=sq, round_ops=0, round_out=True, **kw)
        o = self.lin(ctx, pre + ".output.dense.weight", pre + ".output.dense.bias", x_clean=True)
        site_h = self.dropout_(o, pre + ".hidden")                  # BertSelfOutput: dense -> dropout -> + residual -> LayerNorm
        y, xhat, rstd = self.ln(o, hid, pre + ".output.LayerNorm")
        return y, dict(hid=hid, enc=enc, Q=Q, K=K, V=V, P=P, Pd=Pd, ctx=ctx, xhat=xhat, rstd=rstd, B=B, Lq=Lq, Lk=Lk,
                       site_p=site_p, site_h=site_h, enc_clean=enc_clean)

    def bert_bwd(self, dy, t, pre, denc):
        """returns d hid; accumulates into denc"""
        B, Lq, Lk = t["B"], t["Lq"], t["Lk"]
        D, H = dy.shape[1], self.dims.n_heads
        hd = D // H
        kw, sq, sk, sp = self._attn_strides(B, Lq, Lk, D, H)
        ds = self.ln_bwd(dy, t["xhat"], t["rstd"], pre + ".output.LayerNorm")        # grad of (dropout(o) + hid)
        do = ds
        if t["site_h"] is not None:
            do = ds.clone()
            self.dropout_(do, site=t["site_h"])
        dctx = self.lin_bwd(do, t["ctx"], pre + ".output.dense.weight", pre + ".output.dense.bias", x_clean=True, round_out=True)
        P = t["P"]
        Pv = P if t["Pd"] is None else t["Pd"]                      # what multiplied V in the forward
        drop_seed = None if t["Pd"] is None else self.seed_dev
        # every operand below was stored rounded by its producer (P, dS by the softmax kernels; Q, K, V, dctx, dQ, dK, dV by
        # GEMM epilogues): no rounding pass in these GEMMs
        dV = self.new(B * Lk, D)
        tn.gemm(Pv, dctx, dV, Lk, hd, Lq, a_mn=True, b_mn=True, lda=Lk, ldb=D, ldc=D, a_strides=sp, b_strides=sq, c_strides=sk,
                round_ops=0, round_out=True, **kw)
        dP = self.new(B, H, Lq, Lk)
        tn.gemm(dctx, t["V"], dP, Lq, Lk, hd, lda=D, ldb=D, ldc=Lk, a_strides=sq, b_strides=sk, c_strides=sp, round_ops=0, **kw)
        tn.call("poem_tr_softmax_rows_bwd", P, dP, B * H * Lq, Lk, 1.0 / math.sqrt(hd), self.p_drop, drop_seed, t["site_p"])   # dP <- dS
        dQ = self.new(B * Lq, D)
        tn.gemm(dP, t["K"], dQ, Lq, hd, Lk, b_mn=True, lda=Lk, ldb=D, ldc=D, a_strides=sp, b_strides=sk, c_strides=sq,
                round_ops=0, round_out=True, **kw)
        dK = self.new(B * Lk, D)
        tn.gemm(dP, t["Q"], dK, Lk, hd, Lq, a_mn=True, b_mn=True, lda=Lk, ldb=D, ldc=D, a_strides=sp, b_strides=sq, c_strides=sk,
                round_ops=0, round_out=True, **kw)
        del dP
        self.lin_bwd(dQ, t["hid"], pre + ".self.query.weight", pre + ".self.query.bias", out=ds, acc=True, dy_clean=True)   # ds -> d hid
        self.lin_bwd(dK, t["enc"], pre + ".self.key.weight", pre + ".self.key.bias", out=denc, acc=True, dy_clean=True, x_clean=t["enc_clean"])
        self.lin_bwd(dV, t["enc"], pre + ".self.value.weight", pre + ".self.value.bias", out=denc, acc=True, dy_clean=True, x_clean=t["enc_clean"])
        return ds

    # ------------------------------------------------------------------------------------------ vector attention core
    def va_core_fwd(self, q, ktab, vtab, gidx, rel, pre):
        E, D = rel.shape[0], q.shape[1]
        hd = self.new(E, D)
        tn.call("poem_tr_lin3_relu", rel, self.p[pre + "fc_delta.0.weight"], self.p[pre + "fc_delta.0.bias"], hd, E, D)
        # hd, t, hg (and da, dhg, dpos in the backward) are only ever GEMM operands: their producers store them TF32-rounded
        pos = self.lin(hd, pre + "fc_delta.2.weight", pre + "fc_delta.2.bias", x_clean=True)
        t = self.new(E, D)
        tn.call("poem_tr_va_gather_t", q, ktab, gidx, pos, t, E, D)
        hg = self.lin(t, pre + "fc_gamma.0.weight", pre + "fc_gamma.0.bias", relu=True, x_clean=True, round_out=True)
        w = self.lin(hg, pre + "fc_gamma.2.weight", pre + "fc_gamma.2.bias", x_clean=True)
        res = self.new(q.shape[0], D)
        tn.call("poem_tr_va_softmax_agg", w, vtab, pos, gidx, 1.0 / math.sqrt(D), res, q.shape[0], D)   # w <- softmax weights
        return res, dict(q=q, ktab=ktab, vtab=vtab, gidx=gidx, rel=rel, hd=hd, pos=pos, t=t, hg=hg, w=w)

    def va_core_bwd(self, dres, c, pre, dq, dktab, dvtab, dxyz_q=None, dxyz_ref=None):
        """dq, dktab, dvtab (+=); coordinates: dxyz_q (+=, query side), dxyz_ref (+=, neighbour side) when given"""
        NQ, D = dres.shape
        E = NQ * NBR
        gidx = c["gidx"]
        dvp = self.new(E, D)
        da = c["w"]                                                       # overwritten: the tape entry is dead afterwards
        tn.call("poem_tr_va_softmax_agg_bwd", dres, da, c["vtab"], c["pos"], gidx, 1.0 / math.sqrt(D), dvp, NQ, D)
        dhg = self.lin_bwd(da, c["hg"], pre + "fc_gamma.2.weight", pre + "fc_gamma.2.bias", relu_in=True, dy_clean=True,
                           x_clean=True, round_out=True)
        dt = self.lin_bwd(dhg, c["t"], pre + "fc_gamma.0.weight", pre + "fc_gamma.0.bias", dy_clean=True, x_clean=True)
        del dhg
        tn.call("poem_tr_va_scatter", dt, dvp, gidx, dq, dktab, dvtab, NQ, D)           # dt <- dpos
        del dvp
        dhd = self.lin_bwd(dt, c["hd"], pre + "fc_delta.2.weight", pre + "fc_delta.2.bias", relu_in=True, dy_clean=True, x_clean=True)
        drel = self.new(E, 3) if dxyz_q is not None else None
        tn.call("poem_tr_lin3_bwd", dhd, c["rel"], self.p[pre + "fc_delta.0.weight"], self.g[pre + "fc_delta.0.weight"],
                self.g[pre + "fc_delta.0.bias"], drel, E, D)
        if drel is not None:
            tn.call("poem_tr_va_drel_scatter", drel, gidx, dxyz_q, dxyz_ref, NQ)

    def _neighbours(self, q_xyz, ref_xyz, B, Q, R, forced):
        """global row indices (B*Q*32) of the 32 nearest reference points of every query"""
        if forced is not None:
            local = forced.to(self.dev, torch.int32).contiguous()
        else:
            local = self.new(B, Q, NBR, dtype=torch.int32)
            nat.check(self.lib.poem_knn32(q_xyz.data_ptr(), ref_xyz.data_ptr(), local.data_ptr(), B, Q, R,
                                          torch.cuda.current_stream().cuda_stream))
        gidx = self.new(B * Q * NBR, dtype=torch.int32)
        tn.call("poem_tr_va_make_idx", local, None, B, Q, R, gidx)
        return gidx, local

    def _anchor_idx(self, B, Q, R):
        gidx = self.new(B * Q * NBR, dtype=torch.int32)
        tn.call("poem_tr_va_make_idx", None, self.anchor_idx, B, Q, R, gidx)
        return gidx

    # ------------------------------------------------------------------------------------------ one point_METRO_block
    def block_fwd(self, i, q_feats, q_xyz, pt_feats, pt_xyz, B, forced):
        d = self.dims
        Q, P, D = d.n_query, d.n_sample, d.embed_dims
        p = f"transformer.pt_metro_encoder.{i}."
        ps, pc = p + "encoder.vec_attn.query_self_attn.", p + "encoder.vec_attn.query_cross_attn."
        t = dict(q_feats=q_feats, pt_feats=pt_feats)
        qe = self.lin(q_feats, p + "embedding.weight", p + "embedding.bias")
        ke = self.lin(pt_feats, p + "embedding.weight", p + "embedding.bias", round_out=True)    # ke, xc: GEMM operands only
        t["site_qe"] = self.dropout_(qe, p + "qe")                    # pt_metro_transformer.py:185-186
        t["site_ke"] = self.dropout_(ke, p + "ke")
        ke_clean = self.p_drop == 0.0                                  # x / (1 - p) is no longer TF32-representable
        a1, t["attn1"] = self.bert_fwd(qe, ke, p + "encoder.attn", B, Q, P, ke_clean)
        a2, t["attn2"] = self.bert_fwd(a1, ke, p + "encoder.cross_attn", B, Q, P, ke_clean)
        E = B * Q * NBR
        # --- vector self-attention over the queries (point_transformers.py:70-96)
        if i == 0:
            gs, gc = self._anchor_idx(B, Q, Q), self._anchor_idx(B, Q, P)
            nb = None
        else:
            gs, ls = self._neighbours(q_xyz, q_xyz, B, Q, Q, None if forced is None else forced[0])
            gc, lc = self._neighbours(q_xyz, pt_xyz, B, Q, P, None if forced is None else forced[1])
            nb = torch.stack([ls, lc])
        rel = self.new(E, 3)
        tn.call("poem_tr_va_rel", q_xyz, q_xyz, self.anchor_xyz if i == 0 else None, gs, E, rel)
        xs = self.lin(a2, ps + "fc1.weight", ps + "fc1.bias")
        qs, ks, vs = self.lin(xs, ps + "w_qs.weight"), self.lin(xs, ps + "w_ks.weight"), self.lin(xs, ps + "w_vs.weight")
        res_s, t["core_s"] = self.va_core_fwd(qs, ks, vs, gs, rel, ps)
        f1 = a2.clone()
        self.lin(res_s, ps + "fc2.weight", ps + "fc2.bias", out=f1, acc=True)
        # --- vector cross-attention queries -> basis points (point_transformers.py:125-156)
        rel_c = self.new(E, 3)
        tn.call("poem_tr_va_rel", q_xyz, pt_xyz, self.anchor_xyz if i == 0 else None, gc, E, rel_c)
        qc = self.lin(f1, pc + "w_qs.weight")
        xc = self.lin(ke, pc + "fc1.weight", pc + "fc1.bias", x_clean=ke_clean, round_out=True)
        kc, vc = self.lin(xc, pc + "w_ks.weight", x_clean=True), self.lin(xc, pc + "w_vs.weight", x_clean=True)
        res_c, t["core_c"] = self.va_core_fwd(qc, kc, vc, gc, rel_c, pc)
        f2 = f1.clone()
        self.lin(res_c, pc + "fc2.weight", pc + "fc2.bias", out=f2, acc=True)
        # --- regression branch + FFN
        r = self.lin(f2, p + "encoder.vec_attn.reg_branch.0.weight", p + "encoder.vec_attn.reg_branch.0.bias", relu=True)
        xyz = self.new(B * Q, 3)
        tn.call("poem_tr_lin_n3", r, self.p[p + "encoder.vec_attn.reg_branch.2.weight"],
                self.p[p + "encoder.vec_attn.reg_branch.2.bias"], q_xyz, xyz, B * Q, D)
        hpre = self.lin(f2, p + "encoder.intermediate.dense.weight", p + "encoder.intermediate.dense.bias")
        h = self.new(*hpre.shape)
        tn.call("poem_tr_gelu", hpre, h, h.numel())
        o = self.lin(h, p + "encoder.output.dense.weight", p + "encoder.output.dense.bias")
        t["site_ffn"] = self.dropout_(o, p + "ffn")                    # BertOutput: dense -> dropout -> + residual -> LayerNorm
        t["ke_clean"] = ke_clean
        out, xhat, rstd = self.ln(o, f2, p + "encoder.output.LayerNorm")
        t.update(qe=qe, ke=ke, a1=a1, a2=a2, xs=xs, res_s=res_s, f1=f1, xc=xc, res_c=res_c, f2=f2, r=r, hpre=hpre, h=h,
                 xhat=xhat, rstd=rstd)
        return out, xyz, t, nb

    def block_bwd(self, i, t, dout, dxyz_out, dpt_feats, B):
        """dout: grad of the block's feature output (None for the last block), dxyz_out: grad of its coordinates.
        returns (d q_feats, d q_xyz or None); accumulates d pt_feats"""
        d = self.dims
        Q, P, D = d.n_query, d.n_sample, d.embed_dims
        p = f"transformer.pt_metro_encoder.{i}."
        ps, pc = p + "encoder.vec_attn.query_self_attn.", p + "encoder.vec_attn.query_cross_attn."
        f2 = t["f2"]
        if dout is not None:
            df2 = self.ln_bwd(dout, t["xhat"], t["rstd"], p + "encoder.output.LayerNorm")       # grad of (dropout(o) + f2)
            do = df2
            if t["site_ffn"] is not None:
                do = df2.clone()
                self.dropout_(do, site=t["site_ffn"])
            dh = self.lin_bwd(do, t["h"], p + "encoder.output.dense.weight", p + "encoder.output.dense.bias")
            tn.call("poem_tr_gelu_bwd", dh, t["hpre"], dh.numel())
            self.lin_bwd(dh, f2, p + "encoder.intermediate.dense.weight", p + "encoder.intermediate.dense.bias", out=df2, acc=True)
            del dh
        else:
            df2 = self.zeros(B * Q, D)
        dr = self.new(B * Q, D)
        tn.call("poem_tr_lin_n3_bwd", dxyz_out, t["r"], self.p[p + "encoder.vec_attn.reg_branch.2.weight"], dr,
                self.g[p + "encoder.vec_attn.reg_branch.2.weight"], self.g[p + "encoder.vec_attn.reg_branch.2.bias"], B * Q, D, 1)
        self.lin_bwd(dr, f2, p + "encoder.vec_attn.reg_branch.0.weight", p + "encoder.vec_attn.reg_branch.0.bias", out=df2, acc=True)
        dq_xyz = dxyz_out.clone() if i > 0 else None            # block 0 starts from the constant template
        # --- vector cross-attention: f2 = fc2(res_c) + f1
        df1 = df2.clone()
        dres = self.lin_bwd(df2, t["res_c"], pc + "fc2.weight", pc + "fc2.bias")
        dqc, dkc, dvc = self.zeros(B * Q, D), self.zeros(B * P, D), self.zeros(B * P, D)
        self.va_core_bwd(dres, t["core_c"], pc, dqc, dkc, dvc, dq_xyz, None)             # basis points are constants
        self.lin_bwd(dqc, t["f1"], pc + "w_qs.weight", out=df1, acc=True)
        dxc = self.lin_bwd(dkc, t["xc"], pc + "w_ks.weight", x_clean=True)
        self.lin_bwd(dvc, t["xc"], pc + "w_vs.weight", out=dxc, acc=True, x_clean=True)
        dke = self.lin_bwd(dxc, t["ke"], pc + "fc1.weight", pc + "fc1.bias", x_clean=t["ke_clean"])
        del dqc, dkc, dvc, dxc
        # --- vector self-attention: f1 = fc2(res_s) + a2
        da2 = df1.clone()
        dres = self.lin_bwd(df1, t["res_s"], ps + "fc2.weight", ps + "fc2.bias")
        dqs, dks, dvs = self.zeros(B * Q, D), self.zeros(B * Q, D), self.zeros(B * Q, D)
        self.va_core_bwd(dres, t["core_s"], ps, dqs, dks, dvs, dq_xyz, dq_xyz)           # both ends are query coordinates
        dxs = self.lin_bwd(dqs, t["xs"], ps + "w_qs.weight")
        self.lin_bwd(dks, t["xs"], ps + "w_ks.weight", out=dxs, acc=True)
        self.lin_bwd(dvs, t["xs"], ps + "w_vs.weight", out=dxs, acc=True)
        self.lin_bwd(dxs, t["a2"], ps + "fc1.weight", ps + "fc1.bias", out=da2, acc=True)
        # --- the two BERT cross-attention layers
        da1 = self.bert_bwd(da2, t["attn2"], p + "encoder.cross_attn", dke)
        dqe = self.bert_bwd(da1, t["attn1"], p + "encoder.attn", dke)
        if t["site_qe"] is not None:
            self.dropout_(dqe, site=t["site_qe"])
            self.dropout_(dke, site=t["site_ke"])
        dq_feats = self.lin_bwd(dqe, t["q_feats"], p + "embedding.weight", p + "embedding.bias")
        self.lin_bwd(dke, t["pt_feats"], p + "embedding.weight", p + "embedding.bias", out=dpt_feats, acc=True)
        return dq_feats, dq_xyz

    # ------------------------------------------------------------------------------------------ image features -> point features
    def head_stage_fwd(self, feat, views, intr, extr, centre, inp_w, inp_h):
        d = self.dims
        D, C, P, hw = d.embed_dims, d.in_channels, d.n_sample, d.feat_hw
        HW = hw * hw
        NV, B = int(feat.shape[0]), len(views)
        planes = self.new(NV, D, HW)
        tn.gemm(self.pr["input_proj.weight"], feat, planes, D, HW, C, b_mn=True, ldb=HW, ldc=HW, batch=(NV, 1),
                b_strides=(C * HW, 0), c_strides=(D * HW, 0), bias=self.p["input_proj.bias"], bias_on_m=True, round_ops=2)
        F3 = 3 * d.pos_feats
        key = tuple(int(n) for n in views)
        if key not in self._const:                  # constants of the graph for this view layout (host -> device once)
            sine = torch.cat([sine_pos_3d(int(n), hw, hw, d.pos_feats, d.pos_normalize) for n in views]).reshape(NV, F3, HW)
            self._const[key] = dict(
                sine=sine.to(self.dev).contiguous(),                                  # petr_transformer.py:434-469
                img_sample=torch.tensor([b for b, n in enumerate(views) for _ in range(int(n))], dtype=torch.int32, device=self.dev),
                row0=torch.tensor(np.concatenate([[0], np.cumsum(views)[:-1]]) * P, dtype=torch.int32, device=self.dev),
                nv=torch.tensor(np.asarray(views), dtype=torch.int32, device=self.dev))
        cst = self._const[key]
        sine, img_sample, row0, nv = cst["sine"], cst["img_sample"], cst["row0"], cst["nv"]
        tn.gemm(self.pr["adapt_pos3d.weight"], sine, planes, D, HW, F3, b_mn=True, ldb=HW, ldc=HW, batch=(NV, 1),
                b_strides=(F3 * HW, 0), c_strides=(D * HW, 0), bias=self.p["adapt_pos3d.bias"], bias_on_m=True, accumulate=True)
        grid = self.new(NV, P, 2)
        tn.call("poem_tr_project", self.bps, centre, intr, extr, img_sample, NV, P, float(inp_w), float(inp_h), grid)
        S = self.new(NV, D, P)
        tn.call("poem_tr_sample", planes, grid, S, NV, D, P, hw)
        X = S.view(NV * P, D)                                   # the reference's raw `.view(1, -1, n, D)` regroup, per sample
        h0 = self.lin(X, "merge_net_feature.0.0.weight", "merge_net_feature.0.0.bias", relu=True)
        m = self.lin(h0, "merge_net_feature.0.2.weight", "merge_net_feature.0.2.bias")
        Dm = m.shape[1]
        agg = self.new(B * P, Dm)
        tn.call("poem_tr_merge_agg", m, row0, nv, B, P, Dm, agg)
        h1 = self.lin(agg, "merge_net_feature.1.0.weight", "merge_net_feature.1.0.bias", relu=True)
        y = self.lin(h1, "merge_net_feature.1.2.weight", "merge_net_feature.1.2.bias")
        pt = self.new(B * P, D)
        tn.call("poem_tr_merge_out", X, y, row0, nv, B, P, D, pt)
        return pt, dict(feat=feat, sine=sine, grid=grid, X=X, row0=row0, nv=nv, h0=h0, m=m, agg=agg, h1=h1, NV=NV, B=B)

    def head_stage_bwd(self, dpt, t):
        d = self.dims
        D, C, P, hw = d.embed_dims, d.in_channels, d.n_sample, d.feat_hw
        HW, NV, B = hw * hw, t["NV"], t["B"]
        F3 = 3 * d.pos_feats
        dX = self.zeros(NV * P, D)
        dy = self.new(B * P, D)
        tn.call("poem_tr_merge_out_bwd", dpt, t["row0"], t["nv"], B, P, D, dX, dy)
        dh1 = self.lin_bwd(dy, t["h1"], "merge_net_feature.1.2.weight", "merge_net_feature.1.2.bias", relu_in=True)
        dagg = self.lin_bwd(dh1, t["agg"], "merge_net_feature.1.0.weight", "merge_net_feature.1.0.bias")
        Dm = dagg.shape[1]
        dm = self.zeros(NV * P, Dm)
        tn.call("poem_tr_merge_agg_bwd", dagg, t["m"], t["row0"], t["nv"], B, P, Dm, dm)
        dh0 = self.lin_bwd(dm, t["h0"], "merge_net_feature.0.2.weight", "merge_net_feature.0.2.bias", relu_in=True)
        self.lin_bwd(dh0, t["X"], "merge_net_feature.0.0.weight", "merge_net_feature.0.0.bias", out=dX, acc=True)
        dplanes = self.zeros(NV, D, HW)
        tn.call("poem_tr_sample_bwd", dX, t["grid"], dplanes, NV, D, P, hw)           # dX memory == dS (NV, D, P)
        for b in ("input_proj.bias", "adapt_pos3d.bias"):
            tn.call("poem_tr_rowsum_groups", dplanes, NV * D, HW, D, self.g[b])
        tn.gemm(dplanes, t["feat"], self.g["input_proj.weight"], D, C, HW, lda=HW, ldb=HW, ldc=C, batch=(NV, 1),
                a_strides=(D * HW, 0), b_strides=(C * HW, 0), c_strides=(0, 0), accumulate=True, round_ops=GRAD_ROUND)
        tn.gemm(dplanes, t["sine"], self.g["adapt_pos3d.weight"], D, F3, HW, lda=HW, ldb=HW, ldc=F3, batch=(NV, 1),
                a_strides=(D * HW, 0), b_strides=(F3 * HW, 0), c_strides=(0, 0), accumulate=True, round_ops=GRAD_ROUND)
        dfeat = self.new(NV, C, hw, hw)
        tn.gemm(self.pr["input_proj.weight"], dplanes, dfeat, C, HW, D, a_mn=True, b_mn=True, lda=C, ldb=HW, ldc=HW,
                batch=(NV, 1), b_strides=(D * HW, 0), c_strides=(C * HW, 0), round_ops=GRAD_ROUND)
        return dfeat

    # ------------------------------------------------------------------------------------------ whole head
    @_on_own_device
    def forward(self, mlvl_feat, img_metas, reference_joints, neighbours=None):
        """all_coords_preds (NB, B, 799, 3) in metres; keeps the activations for `backward`.
        `neighbours` (NB-1, 2, B, 799, 32): test hook, use these 32-NN sets instead of searching."""
        d = self.dims
        if self.dev.type != "cuda":
            raise nat.PoemError("the training path has no CPU implementation: build the trainer on a CUDA device")
        tn.call("poem_tr_round_tf32", self.p_flat, self.pr_flat, self.p_flat.numel())      # this step's operand copy of the weights
        if self.p_drop > 0.0:
            self.seed_dev.add_(1)                                   # fresh masks every step (also under CUDA-graph replay)
        self._site, self.drop_sites = 0, {}
        views = [int(v) for v in np.asarray(img_metas["cam_view_num"]).reshape(-1)]
        B = len(views)
        Q, P, D = d.n_query, d.n_sample, d.embed_dims
        feat = mlvl_feat.detach().to(self.dev, torch.float32).contiguous()
        assert feat.shape[0] == sum(views) and feat.shape[1] == d.in_channels and tuple(feat.shape[-2:]) == (d.feat_hw, d.feat_hw)
        intr = img_metas["cam_intr"].to(self.dev, torch.float32).contiguous()
        extr = img_metas["cam_extr"].to(self.dev, torch.float32).contiguous()
        refj = reference_joints.to(self.dev, torch.float32).contiguous()
        centre = refj[:, 9].contiguous()                                            # ptEmb_head.py:873 (fixed joint 9)
        inp_w, inp_h = img_metas["inp_img_shape"]
        pt_feats, t_head = self.head_stage_fwd(feat.view(feat.shape[0], d.in_channels, -1), views, intr, extr, centre, inp_w, inp_h)
        q_feats = self.new(B * Q, D)
        tn.call("poem_tr_bcast_batch", self.p["query_feat_embedding.weight"], B, Q * D, q_feats)
        # normalised coordinates ((x + c) - c) / r as the reference forms them (ptEmb_head.py:896-899)
        pt_xyz, q_xyz = self.new(B * P, 3), self.new(B * Q, 3)
        tmp = self.new(B * P, 3)
        neg = (-centre / d.radius).contiguous()
        tn.call("poem_tr_bcast_batch", self.bps, B, P * 3, pt_xyz)
        tn.call("poem_tr_affine_rows", pt_xyz, centre, 1.0, tmp, B * P, P, B, 3)
        tn.call("poem_tr_affine_rows", tmp, neg, 1.0 / d.radius, pt_xyz, B * P, P, B, 3)
        tn.call("poem_tr_bcast_batch", self.template, B, Q * 3, q_xyz)
        tn.call("poem_tr_affine_rows", q_xyz, centre, 1.0, tmp, B * Q, Q, B, 3)
        tn.call("poem_tr_affine_rows", tmp, neg, 1.0 / d.radius, q_xyz, B * Q, Q, B, 3)
        coords = self.new(d.n_blocks, B, Q, 3)
        blocks, nbs = [], []
        for i in range(d.n_blocks):
            forced = None if (neighbours is None or i == 0) else neighbours[i - 1]
            q_feats, q_xyz, tb, nb = self.block_fwd(i, q_feats, q_xyz, pt_feats, pt_xyz, B, forced)
            blocks.append(tb)
            if nb is not None:
                nbs.append(nb)
            if not (d.parametric and i == d.n_blocks - 1):
                tn.call("poem_tr_affine_rows", q_xyz, centre, d.radius, coords[i], B * Q, Q, B, 3)
        self.last_neighbours = torch.stack(nbs) if nbs else None
        self.tape = dict(head=t_head, blocks=blocks, B=B)
        if d.parametric:        # pt_metro_transformer.py:139-151,194-195: the last block's mesh comes from the MANO layer
            pl = f"transformer.pt_metro_encoder.{d.n_blocks - 1}."
            m = self.mano
            flat, self.pred_pose, self.pred_shape = self.new(B * D), self.new(B, 48), self.new(B, 10)
            tn.call("poem_tr_mano_tail", q_feats, self.p[pl + "flat_verts.weight"], self.p[pl + "flat_verts.bias"],
                    self.p[pl + "mano_linear.weight"], self.p[pl + "mano_linear.bias"], m["v_template"], m["shapedirs"],
                    m["posedirs"], m["j_regressor"], m["skin_weights"], refj, d.center_idx, B, Q, D, flat,
                    coords[d.n_blocks - 1], self.pred_pose, self.pred_shape)
            self.tape["tail"] = dict(feats=q_feats, flat=flat)
        return coords

    @_on_own_device
    def backward(self, dcoords, on_bucket_done=None, dpose=None, dshape=None):
        """dcoords (NB, B, 799, 3): d loss / d all_coords_preds.  Accumulates into `g`, returns d loss / d mlvl_feat.
        `on_bucket_done(name)`: called when every gradient of bucket `name` ("2", "1", "0", then "head") is final, so a
        data-parallel caller can start that bucket's all-reduce while the rest of the backward runs."""
        if self.tape is None:
            raise RuntimeError("backward() without a forward()")
        d = self.dims
        Q, P, D = d.n_query, d.n_sample, d.embed_dims
        B = self.tape["B"]
        dc = dcoords.detach().to(self.dev, torch.float32).contiguous()
        dpt = self.zeros(B * P, D)
        dfe, dxyz_next = None, None
        if d.parametric:        # MANO tail: d coords[-1] (and d pose / d shape of the parameter losses) -> d features of the last block
            pl = f"transformer.pt_metro_encoder.{d.n_blocks - 1}."
            m, tt = self.mano, self.tape["tail"]
            f32 = lambda t: None if t is None else t.detach().to(self.dev, torch.float32).contiguous()  # noqa: E731
            dfe, dflat = self.new(B * Q, D), self.new(B * D)
            tn.call("poem_tr_mano_tail_bwd", tt["feats"], self.p[pl + "flat_verts.weight"], self.p[pl + "mano_linear.weight"],
                    self.p[pl + "mano_linear.bias"], m["v_template"], m["shapedirs"], m["posedirs"], m["j_regressor"],
                    m["skin_weights"], d.center_idx, B, Q, D, tt["flat"], dc[d.n_blocks - 1].reshape(B * Q, 3), f32(dpose),
                    f32(dshape), dflat, dfe, self.g[pl + "flat_verts.weight"], self.g[pl + "flat_verts.bias"],
                    self.g[pl + "mano_linear.weight"], self.g[pl + "mano_linear.bias"])
        for i in reversed(range(d.n_blocks)):
            dxyz = self.new(B * Q, 3)
            if d.parametric and i == d.n_blocks - 1:
                dxyz.zero_()                                   # the last block's own coordinates are overwritten by the MANO mesh
            else:
                tn.call("poem_tr_affine_rows", dc[i].reshape(B * Q, 3), None, d.radius, dxyz, B * Q, Q, B, 3)
            if dxyz_next is not None:
                tn.call("poem_tr_axpy", dxyz, dxyz_next, 1.0, dxyz.numel())
            dfe, dxyz_next = self.block_bwd(i, self.tape["blocks"][i], dfe, dxyz, dpt, B)
            self.tape["blocks"][i] = None                                           # free the block's activations
            if on_bucket_done is not None:
                on_bucket_done(str(i))
        tn.call("poem_tr_sum_batch", dfe, B, Q * D, self.g["query_feat_embedding.weight"])
        dfeat = self.head_stage_bwd(dpt, self.tape["head"])
        self.tape = None
        if on_bucket_done is not None:
            on_bucket_done("head")
        return dfeat

    # ------------------------------------------------------------------------------------------ after backward
    @_on_own_device
    def clip_grad_norm_per_tensor(self, max_norm):
        """lib/utils/net_utils.py:122-132: clip_grad_norm_(param, max_norm, 2) on every parameter tensor by itself.
        Two launches over the flat gradient buffer; returns the squared norms (device, one per tensor)."""
        n = int(self.seg_off.numel())
        ss = self.new(n)
        tn.call("poem_tr_seg_sumsq", self.g_flat, self.seg_off, self.seg_len, n, ss)
        tn.call("poem_tr_seg_clip", self.g_flat, self.seg_off, self.seg_len, n, ss, float(max_norm))
        return ss


class TrainStep:
    """One optimisation step of the head as `scripts/train_ddp.py:96-116` runs it — zero_grad, forward, loss, backward,
    (DDP gradient average,) clip_gradient, Adam — with every piece in device kernels:

        step = TrainStep(trainer, lr=1e-4, max_norm=1.0)            # config/release/train_*.yaml TRAIN
        loss = step(mlvl_feat, img_metas, reference_joints, gt_joints, gt_verts)

    Data parallel (`torch.distributed` initialised, backend nccl): each rank steps on its shard of the batch; the
    gradient buckets (block 2, 1, 0, head stage) are averaged with NCCL all-reduce as soon as the backward has
    finished them, overlapping the rest of the backward — what DDP's reducer does for the reference.
    `graph=True` captures forward + loss + backward into one CUDA graph per (batch, view layout) and replays it
    (the eager schedule is ~520 launches from Python: launch-bound below batch ~16)."""

    def __init__(self, trainer, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0, loss_weights=(1.0, 1.0),
                 group=None, graph=False, j_regressor=None, loss_cfg=None):
        """loss_weights = (joints, vertices) of the built-in 3-D terms.  With `j_regressor` (MANO, (16, 778)) the step uses
        the head's terms of the reference's compute_loss instead (poem_tr_compute_loss: + joints-from-mesh, the 2-D
        projection terms when `target_joints_2d` is passed to the call, pose / shape for parametric heads); `loss_cfg`
        holds the cfg.LOSS weights (JOINTS_LOSS_WEIGHT, VERTICES_LOSS_WEIGHT, JOINTS_2D_LOSS_WEIGHT, VERTICES_2D_LOSS_WEIGHT,
        POSE_LOSS_WEIGHT, SHAPE_LOSS_WEIGHT; defaults of config/release)."""
        import torch.distributed as dist
        self.tr, self.lr, self.betas, self.eps, self.wd, self.max_norm = trainer, lr, betas, eps, weight_decay, max_norm
        self.wj, self.wv = loss_weights
        self.jreg = None if j_regressor is None else j_regressor.detach().to(trainer.dev, torch.float32).reshape(16, 778).contiguous()
        lc = dict(JOINTS_LOSS_WEIGHT=1.0, VERTICES_LOSS_WEIGHT=1.0, JOINTS_2D_LOSS_WEIGHT=1.0, VERTICES_2D_LOSS_WEIGHT=0.0,
                  POSE_LOSS_WEIGHT=0.001, SHAPE_LOSS_WEIGHT=0.0005)
        lc.update(loss_cfg or {})
        self.loss_cfg = lc
        self.losses = None                                   # the 8 terms of the last step (device), reference-loss mode
        self.m = torch.zeros_like(trainer.p_flat)
        self.v = torch.zeros_like(trainer.p_flat)
        self.t = 0
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.allreduce_bytes = 0
        self.use_graph = graph
        self._graphs = {}
        self._pending = []

    # the part that is identical every step for fixed shapes
    def _fwd_loss_bwd(self, feat, metas, refj, gt_j, gt_v, loss, on_bucket_done, extra=None):
        tr = self.tr
        d = tr.dims
        tr.zero_grad()
        loss.zero_()
        coords = tr.forward(feat, metas, refj)
        B = coords.shape[1]
        dpose = dshape = None
        if self.jreg is None:
            dco = tr.new(*coords.shape)
            tn.call("poem_tr_coord_loss", coords, gt_j, gt_v, d.n_blocks, B, 21, d.n_query - 21, float(self.wj), float(self.wv), loss, dco)
        else:                                                  # the reference's compute_loss (head terms) on the last block
            ex, lc = extra or {}, self.loss_cfg
            dco = tr.zeros(*coords.shape)
            views = tuple(int(v) for v in np.asarray(metas["cam_view_num"]).reshape(-1))
            cst = tr._const[views]
            t2d = ex.get("target_joints_2d")
            w2j = float(lc["JOINTS_2D_LOSS_WEIGHT"]) if t2d is not None else 0.0
            w2v = float(lc["VERTICES_2D_LOSS_WEIGHT"]) if t2d is not None else 0.0
            H, W = metas["inp_img_shape"]
            par = d.parametric and ex.get("gt_pose") is not None
            if par:
                dpose, dshape = tr.new(B, 48), tr.new(B, 10)
            losses = tr.new(8)
            tn.call("poem_tr_compute_loss", coords[d.n_blocks - 1], gt_j, gt_v, self.jreg, metas["cam_intr"], metas["cam_extr"],
                    cst["img_sample"], t2d, B, len(cst["img_sample"]), math.sqrt(float(W) ** 2 + float(H) ** 2),
                    float(lc["JOINTS_LOSS_WEIGHT"]), float(lc["VERTICES_LOSS_WEIGHT"]), w2j, w2v,
                    tr.pred_pose if par else None, ex.get("gt_pose") if par else None, tr.pred_shape if par else None,
                    ex.get("gt_shape") if par else None, float(lc["POSE_LOSS_WEIGHT"]), float(lc["SHAPE_LOSS_WEIGHT"]), losses,
                    dco[d.n_blocks - 1], dpose, dshape)
            self.losses = losses
            loss.copy_(losses[7:8])
        tr.backward(dco, on_bucket_done, dpose=dpose, dshape=dshape)
        return coords

    def _bucket_hook(self, name):
        lo, hi = self.tr.buckets[name]
        seg = self.tr.g_flat[lo:hi]
        self.allreduce_bytes += seg.numel() * 4
        if self.dist.get_backend(self.group) == "nccl":
            self._pending.append(self.dist.all_reduce(seg, op=self.dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                          # gloo (CPU tests of this logic): no AVG reduction
            seg.div_(self.world)
            self._pending.append(self.dist.all_reduce(seg, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    @_on_own_device
    def __call__(self, mlvl_feat, img_metas, reference_joints, gt_joints, gt_verts, target_joints_2d=None, gt_pose=None,
                 gt_shape=None):
        """One step; returns the loss (device scalar).  target_joints_2d (NV, 21, 2), gt_pose (B, 48) / gt_shape (B, 10) (the
        first view's MANO parameters): inputs of the reference-loss mode (eager schedule only)."""
        tr = self.tr
        dev = tr.dev
        extra = None
        if self.jreg is not None:
            if self.use_graph:
                raise NotImplementedError("TrainStep: the reference-loss mode runs the eager schedule (graph=False)")
            f32 = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
            extra = dict(target_joints_2d=f32(target_joints_2d), gt_pose=f32(gt_pose), gt_shape=f32(gt_shape))
            img_metas = dict(img_metas)
            img_metas["cam_intr"] = f32(img_metas["cam_intr"])
            img_metas["cam_extr"] = f32(img_metas["cam_extr"])
        feat = mlvl_feat.detach().to(dev, torch.float32).contiguous()
        refj = reference_joints.to(dev, torch.float32).contiguous()
        gt_j = gt_joints.to(dev, torch.float32).contiguous()
        gt_v = gt_verts.to(dev, torch.float32).contiguous()
        hook = self._bucket_hook if self.world > 1 else None
        self.allreduce_bytes = 0
        if not self.use_graph:
            loss = tr.zeros(1)
            self.coords = self._fwd_loss_bwd(feat, img_metas, refj, gt_j, gt_v, loss, hook, extra)
        else:
            loss = self._replay(feat, img_metas, refj, gt_j, gt_v)
            if self.world > 1:                         # graph replay: the buckets are final when the graph is
                for name in list(tr.buckets):
                    self._bucket_hook(name)
        for h in self._pending:
            h.wait()
        self._pending = []
        if self.max_norm is not None:
            tr.clip_grad_norm_per_tensor(self.max_norm)
        self.t += 1
        tn.call("poem_tr_adam", tr.p_flat, tr.g_flat, self.m, self.v, tr.p_flat.numel(), float(self.lr), float(self.betas[0]),
                float(self.betas[1]), float(self.eps), float(self.wd), self.t)
        return loss

    def _replay(self, feat, metas, refj, gt_j, gt_v):
        views = tuple(int(v) for v in np.asarray(metas["cam_view_num"]).reshape(-1))
        key = (views, tuple(metas["inp_img_shape"]))
        if key not in self._graphs:
            tr = self.tr
            st = dict(feat=feat.clone(), refj=refj.clone(), gt_j=gt_j.clone(), gt_v=gt_v.clone(),
                      intr=metas["cam_intr"].to(tr.dev, torch.float32).contiguous().clone(),
                      extr=metas["cam_extr"].to(tr.dev, torch.float32).contiguous().clone(), loss=tr.zeros(1))
            m = dict(metas)
            m["cam_intr"], m["cam_extr"] = st["intr"], st["extr"]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                              # warm-up: allocator pools, kernel attributes, constants
                self._fwd_loss_bwd(st["feat"], m, st["refj"], st["gt_j"], st["gt_v"], st["loss"], None)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["coords"] = self._fwd_loss_bwd(st["feat"], m, st["refj"], st["gt_j"], st["gt_v"], st["loss"], None)
            self._graphs[key] = (g, st)
        g, st = self._graphs[key]
        st["feat"].copy_(feat)
        st["refj"].copy_(refj)
        st["gt_j"].copy_(gt_j)
        st["gt_v"].copy_(gt_v)
        st["intr"].copy_(metas["cam_intr"])
        st["extr"].copy_(metas["cam_extr"])
        g.replay()
        self.coords = st["coords"]
        return st["loss"].clone()              # the graph overwrites its loss buffer on the next replay


class HeadFunction(torch.autograd.Function):
    """torch.autograd bridge: `coords = HeadFunction.apply(trainer, mlvl_feat, img_metas, reference_joints, *params)` lets a
    torch loss (the reference's `compute_loss`) and a torch backbone sit on either side of the hand-written head; the
    parameter gradients come back through `trainer.g` in the order of `trainer.p`."""

    @staticmethod
    def forward(ctx, trainer, mlvl_feat, img_metas, reference_joints, *params):
        ctx.trainer = trainer
        coords = trainer.forward(mlvl_feat, img_metas, reference_joints)
        if trainer.dims.parametric:          # medium_MANO: pred_pose / pred_shape carry the parameter losses' gradients
            return coords, trainer.pred_pose, trainer.pred_shape
        return coords

    @staticmethod
    def backward(ctx, dcoords, dpose=None, dshape=None):
        tr = ctx.trainer
        tr.zero_grad()
        dfeat = tr.backward(dcoords, dpose=dpose, dshape=dshape)
        return (None, dfeat, None, None) + tuple(tr.g[k].clone() for k in tr.p)
