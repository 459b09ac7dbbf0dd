"""GPU parity: the tcgen05 (3xTF32) dense transform against an fp64 reference and against the
fp32 path the reference uses (torch.nn.functional.linear with TF32 off = cuBLAS fp32)."""
import pytest
import torch
import torch.nn.functional as F

from gala_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (128, 64, 32), (1000, 602, 32), (5000, 100, 32), (777, 32, 41),
                                   (300, 1433, 32), (4096, 602, 16), (130, 7, 8), (1, 5, 64), (233000, 602, 32),
                                   # wide outputs (64 < N <= 256): the 172-class classifier of the Papers shape & co.
                                   (5000, 32, 172), (1000, 128, 172), (300, 602, 100), (777, 32, 96), (2000, 64, 128),
                                   (129, 32, 173), (4000, 100, 256), (64, 8, 65),
                                   # wide output + small K + enough tiles for the persistent classifier kernel
                                   (40000, 32, 172), (50000, 64, 128), (38000, 20, 200), (40001, 48, 173),
                                   (300000, 32, 172), (45000, 64, 256)])
def test_linear_matches_fp64_and_torch_fp32(M, K, N):
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device=DEV)
    gen.manual_seed(M + K + N)
    X = torch.rand(M, K, generator=gen, device=DEV) - 0.5
    W = (torch.rand(N, K, generator=gen, device=DEV) - 0.5) * 0.2
    b = torch.rand(N, generator=gen, device=DEV) - 0.5
    want64 = X.double() @ W.double().t() + b.double()
    got = ops.linear(X, W, b)
    ref32 = F.linear(X, W, b)
    err = float((got.double() - want64).norm() / want64.norm())
    err_ref = float((ref32.double() - want64).norm() / want64.norm())
    assert err < 1e-5                       # north_star fp32 bound
    assert err <= max(4 * err_ref, 2e-6)    # same accuracy class as the fp32 SIMT path it replaces
    got_relu = ops.linear(X, W, b, relu=True)
    assert torch.equal(got_relu, torch.relu(got))
    got_nb = ops.linear(X, W)
    assert float((got_nb.double() - (want64 - b.double())).norm() / want64.norm()) < 1e-5


def test_linear_width_limits():
    """N > 256 and row epilogues on wide outputs are refused with GALA_ERR_UNSUPPORTED (no silent fallback)."""
    from gala_b200 import lib as _l
    X = torch.rand(256, 32, device=DEV)
    with pytest.raises(_l.GalaError):
        ops.linear(X, torch.rand(257, 32, device=DEV))
    with pytest.raises(_l.GalaError):
        ops.linear(X, torch.rand(100, 32, device=DEV), att_w=torch.rand(2, 100, device=DEV), att_b=[0.0, 0.0])


def test_linear_fused_attention_projections():
    M, K, N = 3000, 602, 32
    gen = torch.Generator(device=DEV)
    gen.manual_seed(0)
    X = torch.rand(M, K, generator=gen, device=DEV) - 0.5
    W = (torch.rand(N, K, generator=gen, device=DEV) - 0.5) * 0.2
    b = torch.rand(N, generator=gen, device=DEV) - 0.5
    aw = torch.rand(2, N, generator=gen, device=DEV) - 0.5
    ab = [0.25, -0.5]
    y, att = ops.linear(X, W, b, relu=False, att_w=aw, att_b=ab)
    y64 = X.double() @ W.double().t() + b.double()
    want = y64 @ aw.double().t() + torch.tensor(ab, device=DEV, dtype=torch.float64)
    assert float((att.double().t() - want).norm() / want.norm()) < 1e-5
    assert float((y.double() - y64).norm() / y64.norm()) < 1e-5


def test_linear_row_scale_epilogue():
    M, K, N = 2000, 100, 32
    X = torch.rand(M, K, device=DEV) - 0.5
    W = torch.rand(N, K, device=DEV) - 0.5
    b = torch.rand(N, device=DEV)
    rs = torch.rand(M, device=DEV) + 0.1
    got = ops.linear(X, W, b, row_scale=rs, relu=True)
    want = torch.relu(rs[:, None].double() * (X.double() @ W.double().t() + b.double()))
    assert float((got.double() - want).norm() / want.norm()) < 1e-5


@pytest.mark.parametrize("M,K,N", [(1000, 32, 41), (233, 32, 2), (5000, 64, 64), (77, 7, 3), (4096, 47, 33), (1, 1, 1),
                                   (100003, 32, 41), (1000, 32, 47), (65, 33, 64), (31, 64, 5), (32, 32, 32), (4097, 16, 7)])
def test_linear_small_matches_fp64(M, K, N):
    X = torch.rand(M, K, device=DEV) - 0.5
    W = torch.rand(N, K, device=DEV) - 0.5
    b = torch.rand(N, device=DEV)
    want = X.double() @ W.double().t() + b.double()
    got = ops.linear_small(X, W, b)
    assert float((got.double() - want).norm() / want.norm()) < 1e-6
    got_t = ops.linear_small(X, W, b, transpose_out=True)
    assert torch.equal(got_t.t().contiguous(), got)
    assert torch.equal(ops.linear_small(X, W, b, relu=True), torch.relu(got))


@pytest.mark.parametrize("M,K,N", [(1000, 32, 32), (4099, 32, 8), (777, 64, 64), (513, 20, 41)])
def test_linear_small_ex_row_epilogues(M, K, N):
    """gala_linear_small_ex_f32 (the light transform of the row-block pipeline): bias, row scale before the ReLU, the
    two attention projections of the pre-activation row -- against fp64 torch, <= 1e-6 like the tcgen05 kernel."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    X = torch.rand(M, K, device="cuda", generator=g) - 0.5
    W = torch.rand(N, K, device="cuda", generator=g) - 0.5
    b = torch.rand(N, device="cuda", generator=g) - 0.5
    rs = torch.rand(M, device="cuda", generator=g) + 0.5
    aw = torch.rand(2, N, device="cuda", generator=g) - 0.5
    ab = [0.25, -0.5]
    y, att = ops.linear_small_ex(X, W, b, relu=True, row_scale=rs, att_w=aw, att_b=ab, max_ctas=7)
    pre = X.double() @ W.double().t() + b.double()
    want = torch.relu(pre * rs.double()[:, None])
    want_att = (pre @ aw.double().t() + torch.tensor(ab, device="cuda", dtype=torch.float64)).t()
    assert float((y.double() - want).norm() / want.norm()) < 1e-6
    assert float((att.double() - want_att).norm() / want_att.norm()) < 1e-6
    y2 = ops.linear_small_ex(X, W, b)
    assert float((y2.double() - pre).norm() / pre.norm()) < 1e-6
