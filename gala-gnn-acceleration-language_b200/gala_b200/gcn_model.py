"""The 2-layer GCN forward as GALA generates it (reference codegen/gala.cu:422-459, emitted by
src/codegen/common.h:1128-1180; after complexityOperatorReordering the transform runs first):

    deg = A @ 1; norm = deg^-0.5                       (training-invariant, computed once)
    layer 1: res = fc0(X); res = norm*res; res = A@res; res = norm*res; res = relu(res)
    layer 2: res = norm*res; res = A@res; res = norm*res; out = fc1(res)

Here every `norm*res` pass is folded into a kernel epilogue: the transform scales its rows
(gala_linear_f32 row_scale), layer 1's aggregation applies norm^2 before the ReLU
(norm * relu(norm * x) == relu(norm^2 * x) for norm > 0, which also pre-scales layer 2's input),
layer 2's aggregation applies norm.  Four launches instead of eleven."""
import torch
import torch.nn.functional as F

from . import ops
from .gat_model import _linear_init


class GCN2:
    def __init__(self, in_feats, hidden, classes, device, seed=0):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.fc0 = _linear_init(gen, hidden, in_feats, device)
        self.fc1 = _linear_init(gen, classes, hidden, device)
        self.norm = None

    def prepare(self, g):
        """Invariant code: degrees through the aggregation kernel itself (A @ ones, K = 1)."""
        ones = torch.ones(g.ncols, 1, device=g.device)
        deg = ops.spmm(g, ones).reshape(-1)
        self.norm = torch.pow(deg, -0.5).contiguous()
        self.norm2 = (self.norm * self.norm).contiguous()
        return self

    def forward_literal(self, g, X):
        """Op-by-op, as emitted (ATen elementwise passes between the kernels)."""
        n = self.norm[:, None]
        res = F.linear(X, *self.fc0)
        res = n * res
        res = ops.spmm(g, res)
        res = torch.relu(n * res)
        res = n * res
        res = ops.spmm(g, res)
        res = n * res
        return F.linear(res, *self.fc1)

    def forward(self, g, X, hook=None, dense="tcgen05"):
        run = hook if hook is not None else (lambda name, fn: fn())
        if dense == "tcgen05":
            res = run("linear1", lambda: ops.linear(X, self.fc0[0], self.fc0[1], row_scale=self.norm))
        else:
            res = self.norm[:, None] * F.linear(X, *self.fc0)
        res = run("gcn_aggregate1", lambda: ops.spmm(g, res, row_scale=self.norm2, relu=True))
        res = run("gcn_aggregate2", lambda: ops.spmm(g, res, row_scale=self.norm))
        return F.linear(res, *self.fc1)


class GCNN:
    """L-layer GCN in the shape GALA emits (BASELINE.json configs[4]: 3 layers on the Papers shape).
    Hidden layers run transform-first (complexityOperatorReordering picks the narrower side):
        h_i = relu(norm * (A @ (norm * fc_i(h_{i-1}))))
    the last layer aggregates at the hidden width and applies fc_L afterwards, as GCN2's second layer:
        out = fc_L(norm * (A @ (norm * h_{L-2})))
    Every `norm * res` pass lives in an epilogue: the transform scales its rows, the aggregation applies norm
    (norm^2 before the ReLU on the last hidden layer, which pre-scales the final aggregation's input).
    dims = [feats, hidden, ..., hidden, classes]."""

    def __init__(self, dims, device, seed=0):
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.dims = list(dims)
        self.L = len(dims) - 1
        self.fc = [_linear_init(gen, dims[i + 1], dims[i], device) for i in range(self.L)]
        self.norm = self.norm2 = None

    def prepare(self, g):
        ones = torch.ones(g.ncols, 1, device=g.device)
        self.norm = torch.pow(ops.spmm(g, ones).reshape(-1), -0.5).contiguous()
        self.norm2 = (self.norm * self.norm).contiguous()
        return self

    def forward_literal(self, g, X):
        n = self.norm[:, None]
        res = X
        for i in range(self.L - 1):
            res = torch.relu(n * ops.spmm(g, n * F.linear(res, *self.fc[i])))
        return F.linear(n * ops.spmm(g, n * res), *self.fc[-1])

    def forward(self, g, X, hook=None):
        run = hook if hook is not None else (lambda name, fn: fn())
        res = X
        for i in range(self.L - 1):
            t = run(f"linear{i + 1}", lambda: ops.linear(res, self.fc[i][0], self.fc[i][1], row_scale=self.norm))
            last_hidden = i == self.L - 2
            res = run(f"gcn_aggregate{i + 1}", lambda: ops.spmm(g, t, row_scale=self.norm2 if last_hidden else self.norm,
                                                                 relu=True))
        agg = run(f"gcn_aggregate{self.L}", lambda: ops.spmm(g, res, row_scale=self.norm))
        return run("classifier", lambda: ops.dense(agg, *self.fc[-1]))
