"""The 2-layer GAT forward as GALA generates it (reference src/codegen/common.h:622-675,
735-810, 835-927), on top of the fused kernel.

  layer 1: res = fc0(X); aL = efc0(res); aR = efc1(res)
           res = relu( softmax_row(LeakyReLU_0.2(aL[row] + aR[col])) @ res )
  layer 2 (FFN-recompute rewrite, src/middle-end/middle-end.h:324-375): the attention
           logits come from fc1(res) but the aggregation runs at the narrower hidden
           width and fc1 is applied after it:
           t = fc1(res); aL = efc2(t); aR = efc3(t); res = fc1( attention @ res )

The dense transforms are torch.nn.functional.linear (cuBLAS through libtorch, fp32),
exactly what the generated program uses (common.h:1185-1281); they are adjacent to,
not part of, the sparse hot path.  Random-init weights (nn.Linear default init under a
fixed seed), since no checkpoint exists.
"""
import math

import torch
import torch.nn.functional as F

from . import ops


def _linear_init(gen, out_f, in_f, device):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, device=device) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, device=device) * 2 - 1) * bound
    return w, b


class GAT2:
    def __init__(self, in_feats, hidden, classes, device, seed=0):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.fc0 = _linear_init(gen, hidden, in_feats, device)
        self.efc0 = _linear_init(gen, 1, hidden, device)
        self.efc1 = _linear_init(gen, 1, hidden, device)
        self.fc1 = _linear_init(gen, classes, hidden, device)
        self.efc2 = _linear_init(gen, 1, classes, device)
        self.efc3 = _linear_init(gen, 1, classes, device)
        self.slope = 0.2
        self.fold()

    def fold(self):
        """Attention projections as linear functions of the AGGREGATED features, so that the
        fused kernel can recompute aR[col] from the row it gathers (ops.gat_forward_dot):
          layer 1: aR = efc1(res)                      -> wR1 = efc1.w,        bR1 = efc1.b
          layer 2: aL/aR = efc2/efc3(fc1(res))         -> w = efc.w @ fc1.W,   b = efc.w @ fc1.b + efc.b
        Scalars are read back once here, never inside a step."""
        W1, b1 = self.fc1
        self.wR1 = self.efc1[0].reshape(-1).contiguous()
        self.bR1 = float(self.efc1[1])
        self.wL2 = (self.efc2[0] @ W1).reshape(-1).contiguous()
        self.bL2 = float(self.efc2[0] @ b1 + self.efc2[1])
        self.wR2 = (self.efc3[0] @ W1).reshape(-1).contiguous()
        self.bR2 = float(self.efc3[0] @ b1 + self.efc3[1])
        # both projections of a layer as one [2, K] weight
        self.W_att1 = torch.cat([self.efc0[0], self.efc1[0]], 0).contiguous()
        self.b_att1 = torch.cat([self.efc0[1], self.efc1[1]], 0).contiguous()
        self.b_att1_host = [float(self.efc0[1]), float(self.efc1[1])]
        self.W_att2 = torch.stack([self.wL2, self.wR2], 0).contiguous()
        self.b_att2 = torch.tensor([self.bL2, self.bR2], device=W1.device)
        self.b_att2_host = [self.bL2, self.bR2]
        self.fc1_wT = W1.t().contiguous()
        self._reflected = None

    def fold_reflected(self):
        """mode="reflected": every aggregation gathers its rows in a basis reflected so that the LAST feature column is
        the right-hand attention term (gala_gat_forward_col_f32): H = I - 2 v v^T with H e_last = -+ wR/|wR|.  The
        reflection of layer 1 folds into the layer-1 transform (W' = H1 W, b' = H1 b) and into the left-hand
        projection (w' = H1 w); layer 1's kernel reflects its finished rows back, applies the ReLU and reflects them
        into layer 2's basis (H2 from the folded projection wR2); layer 2's kernel reflects back before the
        classifier.  Same math as "folded" up to fp32 rounding (H is orthogonal); one gather per edge instead of two.
        Host arithmetic on [32] / [32, F] weights, done once per weight update."""
        if self._reflected is None:
            v1, s1 = ops.reflection(self.wR1)
            v2, s2 = ops.reflection(self.wR2)
            W0, b0 = self.fc0
            self._reflected = dict(
                v1=v1, s1=s1, v2=v2, s2=s2,
                W0=ops.reflect(W0, v1, dim=0), b0=ops.reflect(b0, v1),
                W_att1=ops.reflect(self.W_att1, v1), W_att2=ops.reflect(self.W_att2, v2))
        return self._reflected

    def attention_inputs(self, t, wl, wr):
        return F.linear(t, *wl).reshape(-1), F.linear(t, *wr).reshape(-1)

    def forward(self, g, X, hook=None, mode="folded", dense="tcgen05"):
        """g: TiledGraph (rows = output nodes, cols index X's rows).  `hook(name, fn)` lets the
        benchmark time the sparse kernels individually.
          mode="literal": the op sequence of the generated program, one Linear per projection
          mode="folded" : same math with the attention projections folded (one [K,2] matmul
                          per layer; layer 2 never materialises fc1(res) for the logits)
          mode="fused"  : folded projections AND the dense ops that follow each aggregation computed in
                          the kernels' epilogues (gala_linear_f32 + 2 x gala_gat_forward_ex_f32)
          mode="dot"    : aR recomputed inside the kernel (gala_gat_forward_dot_f32)
          mode="folded_dot": "folded" (own kernels for every dense op) + aR recomputed inside the kernel
          mode="reflected_fused": "fused" in that basis (gala_gat_forward_col_f32 with its dense epilogue)
          mode="reflected": "folded" with the aggregated rows stored in a reflected basis whose last column IS the
                          right-hand attention term (fold_reflected; gala_gat_forward_col_f32): one gather per edge
        dense="tcgen05" runs the layer-1 transform (and, in folded mode, its two attention
        projections, fused in the epilogue) on the tensor cores (gala_linear_f32); "torch" = cuBLAS."""
        run = hook if hook is not None else (lambda name, fn: fn())
        hidden, classes = self.fc0[0].shape[0], self.fc1[0].shape[0]
        if dense == "tcgen05" and (hidden > ops.LINEAR_SMALL_MAX or classes > ops.LINEAR_SMALL_MAX):
            # widths outside the hand-written transforms (gala_linear_f32: N <= LINEAR_MAX_N; gala_linear_small_f32:
            # K, N <= 64): same op sequence with the dense parts on cuBLAS (what the generated program uses)
            dense = "torch"
            mode = "folded" if mode == "fused" else mode
        if mode == "fused":
            # three launches, no library kernel: tcgen05 transform (+ layer-1 projections), fused GAT layer
            # (+ layer-2 projections of its own output rows), fused GAT layer (+ classifier on its rows)
            res, a = run("linear1", lambda: ops.linear(X, self.fc0[0], self.fc0[1], att_w=self.W_att1, att_b=self.b_att1_host))
            res, a2, _ = run("gat_layer1", lambda: ops.gat_forward_ex(g, a[0], a[1], res, self.slope, relu=True,
                                                                      att_w=self.W_att2, att_b=self.b_att2_host))
            _, _, out = run("gat_layer2", lambda: ops.gat_forward_ex(g, a2[0], a2[1], res, self.slope, relu=False,
                                                                     cls_wT=self.fc1_wT, cls_b=self.fc1[1], want_y=False))
            return out
        if mode == "reflected" and hidden not in (4, 8, 16, 32):
            mode = "folded"             # the column mode holds a row in hidden/4 lanes of one warp pass
        if mode == "reflected_fused" and (hidden not in (4, 8, 16, 32) or dense != "tcgen05"):
            mode = "fused" if dense == "tcgen05" else "folded"
        if mode == "reflected_fused":
            # "fused" in the reflected basis: three launches, the left-hand projection of layer 2 and the classifier in
            # the aggregation kernels' epilogues
            r = self.fold_reflected()
            res, a = run("linear1", lambda: ops.linear(X, r["W0"], r["b0"], att_w=r["W_att1"], att_b=self.b_att1_host))
            res, a2, _ = run("gat_layer1", lambda: ops.gat_forward_col_ex(
                g, a[0], r["s1"], self.bR1, res, self.slope, relu=True, reflect_in=r["v1"], reflect_out=r["v2"],
                att_w=r["W_att2"], att_b=self.b_att2_host))
            _, _, out = run("gat_layer2", lambda: ops.gat_forward_col_ex(
                g, a2[0], r["s2"], self.bR2, res, self.slope, relu=False, reflect_in=r["v2"],
                cls_wT=self.fc1_wT, cls_b=self.fc1[1], want_y=False))
            return out
        if mode == "reflected":
            r = self.fold_reflected()
            if dense == "tcgen05":
                res, a = run("linear1", lambda: ops.linear(X, r["W0"], r["b0"], att_w=r["W_att1"], att_b=self.b_att1_host))
                aL = a[0]
            else:
                res = F.linear(X, r["W0"], r["b0"])
                aL = F.linear(res, r["W_att1"][:1], self.b_att1[:1]).reshape(-1)
            # layer 2's left-hand projection of every finished (twice reflected) row comes from the kernel's epilogue
            res, a, _ = run("gat_layer1", lambda: ops.gat_forward_col_ex(
                g, aL, r["s1"], self.bR1, res, self.slope, relu=True, reflect_in=r["v1"], reflect_out=r["v2"],
                att_w=r["W_att2"], att_b=self.b_att2_host))
            aL = a[0]
            agg = run("gat_layer2", lambda: ops.gat_forward_col(g, aL, r["s2"], self.bR2, res, self.slope, relu=False,
                                                                reflect_in=r["v2"]))
            if dense == "tcgen05":
                return run("classifier", lambda: ops.linear_small(agg, self.fc1[0], self.fc1[1]))
            return F.linear(agg, *self.fc1)
        if dense == "tcgen05" and mode == "folded_dot":
            # as "folded", with the right-hand attention term recomputed inside the aggregation kernel from the row
            # it gathers (gala_gat_forward_dot_f32: one gather per edge instead of two); aR is never read
            res, a = run("linear1", lambda: ops.linear(X, self.fc0[0], self.fc0[1], att_w=self.W_att1, att_b=self.b_att1_host))
            res = run("gat_layer1", lambda: ops.gat_forward_dot(g, a[0], self.wR1, self.bR1, res, self.slope, relu=True))
            a = run("att2", lambda: ops.linear_small(res, self.W_att2, self.b_att2, transpose_out=True))
            agg = run("gat_layer2", lambda: ops.gat_forward_dot(g, a[0], self.wR2, self.bR2, res, self.slope, relu=False))
            return run("classifier", lambda: ops.linear_small(agg, self.fc1[0], self.fc1[1]))
        if dense == "tcgen05" and mode == "folded":
            # five launches, all this repository's kernels (no library call in the step)
            res, a = run("linear1", lambda: ops.linear(X, self.fc0[0], self.fc0[1], att_w=self.W_att1, att_b=self.b_att1_host))
            res = run("gat_layer1", lambda: ops.gat_forward(g, a[0], a[1], res, self.slope, relu=True))
            a = run("att2", lambda: ops.linear_small(res, self.W_att2, self.b_att2, transpose_out=True))
            agg = run("gat_layer2", lambda: ops.gat_forward(g, a[0], a[1], res, self.slope, relu=False))
            return run("classifier", lambda: ops.linear_small(agg, self.fc1[0], self.fc1[1]))
        res = ops.linear(X, *self.fc0) if dense == "tcgen05" else F.linear(X, *self.fc0)
        if mode == "dot":
            aL = F.linear(res, *self.efc0).reshape(-1)
            res = run("gat_layer1", lambda: ops.gat_forward_dot(g, aL, self.wR1, self.bR1, res, self.slope, relu=True))
            aL = torch.addmv(torch.full((res.shape[0],), self.bL2, device=res.device), res, self.wL2)
            agg = run("gat_layer2", lambda: ops.gat_forward_dot(g, aL, self.wR2, self.bR2, res, self.slope, relu=False))
            return F.linear(agg, *self.fc1)
        if mode == "folded":
            a = F.linear(res, self.W_att1, self.b_att1).t().contiguous()      # [2, N]: aL, aR
            res = run("gat_layer1", lambda: ops.gat_forward(g, a[0], a[1], res, self.slope, relu=True))
            a = F.linear(res, self.W_att2, self.b_att2).t().contiguous()
            agg = run("gat_layer2", lambda: ops.gat_forward(g, a[0], a[1], res, self.slope, relu=False))
            return F.linear(agg, *self.fc1)
        aL, aR = self.attention_inputs(res, self.efc0, self.efc1)
        res = run("gat_layer1", lambda: ops.gat_forward(g, aL, aR, res, self.slope, relu=True))
        t = F.linear(res, *self.fc1)
        aL, aR = self.attention_inputs(t, self.efc2, self.efc3)
        agg = run("gat_layer2", lambda: ops.gat_forward(g, aL, aR, res, self.slope, relu=False))
        return F.linear(agg, *self.fc1)


    def forward_host(self, g, X_host, out_host=None, chunks=8, mode="folded", stage=None):
        """The same forward for features that live in (pinned) HOST memory: the [N, F] matrix is uploaded in `chunks`
        row blocks on a copy stream while the row-tiled layer-1 transform already consumes the blocks that have
        landed (it reads X front to back exactly once), so the PCIe transfer and the transform overlap; the logits
        are copied back into out_host (pinned) if given.  Returns the device logits (asynchronously: synchronise the
        stream before reading out_host).  Consecutive calls pipeline: the next call's upload starts as soon as this
        call's transform has consumed the staging buffer, i.e. under this call's aggregation layers and download."""
        n, dev = X_host.shape[0], g.device
        if stage is None:
            stage = torch.empty(X_host.shape, dtype=torch.float32, device=dev)
        hidden = self.fc0[0].shape[0]
        res = torch.empty((n, hidden), dtype=torch.float32, device=dev)
        if mode == "reflected" and hidden in (4, 8, 16, 32):
            r = self.fold_reflected()
            W0, b0, W_att1 = r["W0"], r["b0"], r["W_att1"]
        else:
            mode = "folded" if mode == "reflected" else mode
            W0, b0, W_att1 = self.fc0[0], self.fc0[1], self.W_att1
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs, main = self._copy_stream, torch.cuda.current_stream()
        # The staging buffer may still be read by the previous call's transform -- and by nothing later in that call:
        # waiting for that point only (not for the whole stream) lets this call's upload run under the previous call's
        # aggregation layers and under its logits download (PCIe is full duplex), so that back-to-back calls are paced by
        # the upload alone.  All compute stays in order on the caller's stream.
        free = getattr(self, "_stage_free", None)
        if free is not None and free[0] == stage.data_ptr():
            cs.wait_event(free[1])
        else:
            cs.wait_stream(main)
        step = (n + chunks - 1) // chunks
        step = (step + 127) // 128 * 128          # whole 128-row tiles per block
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            with torch.cuda.stream(cs):
                stage[lo:hi].copy_(X_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            main.wait_event(ev)
            ops.linear(stage[lo:hi], W0, b0, out=res[lo:hi])
        ev = torch.cuda.Event()
        ev.record(main)
        self._stage_free = (stage.data_ptr(), ev)
        a = ops.linear_small(res, W_att1, self.b_att1, transpose_out=True)
        if mode == "reflected":
            res, a, _ = ops.gat_forward_col_ex(g, a[0], r["s1"], self.bR1, res, self.slope, relu=True, reflect_in=r["v1"],
                                               reflect_out=r["v2"], att_w=r["W_att2"], att_b=self.b_att2_host)
            agg = ops.gat_forward_col(g, a[0], r["s2"], self.bR2, res, self.slope, relu=False, reflect_in=r["v2"])
        elif mode == "folded_dot":
            res = ops.gat_forward_dot(g, a[0], self.wR1, self.bR1, res, self.slope, relu=True)
            a = ops.linear_small(res, self.W_att2, self.b_att2, transpose_out=True)
            agg = ops.gat_forward_dot(g, a[0], self.wR2, self.bR2, res, self.slope, relu=False)
        else:
            res = ops.gat_forward(g, a[0], a[1], res, self.slope, relu=True)
            a = ops.linear_small(res, self.W_att2, self.b_att2, transpose_out=True)
            agg = ops.gat_forward(g, a[0], a[1], res, self.slope, relu=False)
        out = ops.linear_small(agg, self.fc1[0], self.fc1[1])
        if out_host is not None:
            out_host.copy_(out, non_blocking=True)
        return out


class GATN:
    """L-layer GAT in the shape GALA emits for deeper programs (BASELINE.json configs[4]: 3 layers on the
    Papers shape): hidden layers are  t = fc_i(res); res = relu(attention_i(t) @ t)  and the last layer uses
    the FFN-recompute rewrite (middle-end.h:324-375) as GAT2 does: logits from fc_L(res), aggregation at the
    hidden width, fc_L applied after it.  dims = [feats, hidden, ..., hidden, classes]."""

    def __init__(self, dims, device, seed=0):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.dims = list(dims)
        self.L = len(dims) - 1
        self.fc = [_linear_init(gen, dims[i + 1], dims[i], device) for i in range(self.L)]
        self.efcL = [_linear_init(gen, 1, dims[i + 1], device) for i in range(self.L)]
        self.efcR = [_linear_init(gen, 1, dims[i + 1], device) for i in range(self.L)]
        self.slope = 0.2
        # [2, width] attention projections per layer; the last one folded through fc_L
        self.W_att, self.b_att = [], []
        for i in range(self.L - 1):
            self.W_att.append(torch.cat([self.efcL[i][0], self.efcR[i][0]], 0).contiguous())
            self.b_att.append(torch.cat([self.efcL[i][1], self.efcR[i][1]], 0).contiguous())
        W, b = self.fc[-1]
        wl, wr = self.efcL[-1][0] @ W, self.efcR[-1][0] @ W
        self.W_att.append(torch.cat([wl, wr], 0).contiguous())
        self.b_att.append(torch.cat([self.efcL[-1][0] @ b + self.efcL[-1][1],
                                     self.efcR[-1][0] @ b + self.efcR[-1][1]], 0).contiguous())

    def reflected_ok(self):
        return all(d in (4, 8, 16, 32) for d in self.dims[1:-1])

    def fold_reflected(self):
        """The reflected basis of gala_gat_forward_col_f32 for every aggregation (see GAT2.fold_reflected): layer i
        gathers rows whose last column is its right-hand attention term.  Hidden layer i < L-1 gathers t_i = fc_i(res):
        H_i (from efcR_i) folds into fc_i and into the left-hand projection; its kernel reflects the finished rows back
        (they feed the next transform in the original basis) -- except the last hidden layer, whose rows the final
        layer gathers directly: those leave in H_{L-1}'s basis (from the folded right-hand projection).
        Returns dict(v, s, bR: per aggregation; W, b: per transform; W_att: left-hand rows, reflected)."""
        if getattr(self, "_reflected", None) is None:
            if not hasattr(self, "_bh"):
                self.host_biases()
            L = self.L
            v, s = [], []
            for i in range(L):
                vi, si = ops.reflection(self.W_att[i][1])
                v.append(vi)
                s.append(si)
            self._reflected = dict(
                v=v, s=s, bR=[self._bh[i][1] for i in range(L)],
                W=[ops.reflect(self.fc[i][0], v[i], dim=0) for i in range(L - 1)],
                b=[ops.reflect(self.fc[i][1], v[i]) for i in range(L - 1)],
                W_att=[ops.reflect(self.W_att[i], v[i]) for i in range(L)])
        return self._reflected

    def forward_literal(self, g, X):
        """Op by op (one Linear per projection, logits of the last layer from fc_L(res))."""
        res = X
        for i in range(self.L - 1):
            t = F.linear(res, *self.fc[i])
            aL = F.linear(t, *self.efcL[i]).reshape(-1)
            aR = F.linear(t, *self.efcR[i]).reshape(-1)
            res = ops.gat_forward(g, aL, aR, t, self.slope, relu=True)
        t = F.linear(res, *self.fc[-1])
        aL = F.linear(t, *self.efcL[-1]).reshape(-1)
        aR = F.linear(t, *self.efcR[-1]).reshape(-1)
        agg = ops.gat_forward(g, aL, aR, res, self.slope, relu=False)
        return F.linear(agg, *self.fc[-1])

    def forward(self, g, X, hook=None, logits_chunk=None, mode="folded"):
        """Own kernels for the transforms (tcgen05 + folded projections in its epilogue).
        mode="reflected": every aggregation in the reflected basis (fold_reflected; hidden widths in {4, 8, 16, 32}).
        logits_chunk = (rows, buffer): the classifier runs in row chunks into a re-used buffer (for
        graphs whose [N, classes] logits do not fit next to the features on one GPU); returns None."""
        run = hook if hook is not None else (lambda name, fn: fn())
        if not hasattr(self, "_bh"):
            self.host_biases()
        res, a_last = X, None
        if mode == "reflected" and self.reflected_ok():
            r, L = self.fold_reflected(), self.L
            for i in range(L - 1):
                t, a = run(f"linear{i + 1}", lambda: ops.linear(res, r["W"][i], r["b"][i], att_w=r["W_att"][i],
                                                                att_b=self._bh[i]))
                if i == L - 2:
                    res, a_last, _ = run(f"gat_layer{i + 1}", lambda: ops.gat_forward_col_ex(
                        g, a[0], r["s"][i], r["bR"][i], t, self.slope, relu=True, reflect_in=r["v"][i],
                        reflect_out=r["v"][L - 1], att_w=r["W_att"][L - 1], att_b=self._bh[L - 1]))
                else:
                    res = run(f"gat_layer{i + 1}", lambda: ops.gat_forward_col(
                        g, a[0], r["s"][i], r["bR"][i], t, self.slope, relu=True, reflect_in=r["v"][i]))
            agg = run(f"gat_layer{L}", lambda: ops.gat_forward_col(
                g, a_last[0], r["s"][L - 1], r["bR"][L - 1], res, self.slope, relu=False, reflect_in=r["v"][L - 1]))
            return self._classify(run, agg, logits_chunk)
        for i in range(self.L - 1):
            t, a = run(f"linear{i + 1}", lambda: ops.linear(res, self.fc[i][0], self.fc[i][1], att_w=self.W_att[i],
                                                            att_b=self._bh[i]))
            if i == self.L - 2:     # the last hidden layer also projects its rows for the final layer's logits
                res, a_last, _ = run(f"gat_layer{i + 1}", lambda: ops.gat_forward_ex(
                    g, a[0], a[1], t, self.slope, relu=True, att_w=self.W_att[-1], att_b=self._bh[-1]))
            else:
                res = run(f"gat_layer{i + 1}", lambda: ops.gat_forward(g, a[0], a[1], t, self.slope, relu=True))
        agg = run(f"gat_layer{self.L}", lambda: ops.gat_forward(g, a_last[0], a_last[1], res, self.slope, relu=False))
        return self._classify(run, agg, logits_chunk)

    def _classify(self, run, agg, logits_chunk):
        if logits_chunk is not None:
            rows, buf = logits_chunk
            for lo in range(0, agg.shape[0], rows):
                hi = min(agg.shape[0], lo + rows)
                run("classifier", lambda: ops.dense(agg[lo:hi], self.fc[-1][0], self.fc[-1][1], out=buf[:hi - lo]))
            return None
        return run("classifier", lambda: ops.dense(agg, *self.fc[-1]))

    def host_biases(self):
        """Reads the folded biases back once (outside any timed step)."""
        self._bh = [[float(v) for v in b] for b in self.b_att]
        return self
