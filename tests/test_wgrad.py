"""nrh_wgrad_f16 (csrc/wgrad_tc.cu): the weight-gradient reductions of a training step on tcgen05 with TMA tensor-map operand loads,
against fp32 torch matmuls of the same fp16 operands.  Ragged point counts (TMA zero fill), column windows, all N widths, several
jobs per launch (split-K ranges crossing job boundaries), accumulation into a non-zero output, row / column limits, device scale,
and the skinny (m <= 8) path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, scale):
    return (a.float().t() @ b.float()) * scale


@pytest.mark.parametrize("P", [64, 1000, 70001])
def test_wgrad_matches_matmul(P):
    from nrhints_b200.train_ops import WgradBatch
    g = torch.Generator().manual_seed(P)
    dev = "cuda"
    A1 = (torch.randn(P, 256, generator=g) * 0.5).half().to(dev)
    A2 = (torch.randn(P, 512, generator=g) * 0.5).half().to(dev)          # column window [256, 512) is used
    B256 = (torch.randn(P, 256, generator=g) * 0.5).half().to(dev)
    B64 = (torch.randn(P, 64, generator=g) * 0.5).half().to(dev)
    B384 = (torch.randn(P, 384, generator=g) * 0.5).half().to(dev)
    dscale = torch.tensor([0.25], device=dev)
    o1 = torch.full((256, 256), 3.0, device=dev)                           # accumulated into
    o2 = torch.zeros(256, 64, device=dev)
    o3 = torch.zeros(256, 384, device=dev)
    o4 = torch.zeros(256, 256, device=dev)
    o5 = torch.zeros(256, 39, device=dev)                                  # ld_out = 39: unaligned rows, scalar atomics
    wb = WgradBatch()
    wb.add(A1, B256, o1, scale=2.0)
    wb.add(A2, B64, o2, scale=1.0, dev_scale=dscale, a_col0=256, m=256)
    wb.add(A1, B384, o3, n=256)                                            # N = 384 as 256 + 128
    wb.add(A1, B384, o3[:, 256:], b_col0=256, n=128)
    wb.add(A1, B256, o4, rows_valid=217)
    wb.add(A2, B256, o4, a_col0=256, m=256, rows_valid=217)                # two products into one output (as every SDF layer has)
    wb.add(A1, B64, o5, cols_valid=39)
    wb.run()
    torch.cuda.synchronize()
    tol = 2e-3 * (P ** 0.5) * 0.25 + 1e-3                                   # fp32 accumulation of P products of ~0.25 magnitude
    assert float((o1 - (3.0 + _ref(A1, B256, 2.0))).abs().max()) < tol
    assert float((o2 - _ref(A2[:, 256:], B64, 0.25)).abs().max()) < tol
    assert float((o3 - _ref(A1, B384, 1.0)).abs().max()) < tol
    want4 = _ref(A1, B256, 1.0) + _ref(A2[:, 256:], B256, 1.0)
    assert float((o4[:217] - want4[:217]).abs().max()) < 2 * tol and float(o4[217:].abs().max()) == 0.0
    assert float((o5 - _ref(A1, B64, 1.0)[:, :39]).abs().max()) < tol


def test_wgrad_skinny_rows():
    from nrhints_b200.train_ops import WgradBatch
    g = torch.Generator().manual_seed(1)
    P = 5003
    A = torch.zeros(P, 8, dtype=torch.float16, device="cuda")
    A[:, :3] = (torch.randn(P, 3, generator=g) * 0.3).half().cuda()
    B = (torch.randn(P, 256, generator=g) * 0.5).half().cuda()
    out = torch.zeros(3, 256, device="cuda")
    WgradBatch().add(A, B, out, scale=0.5, m=3).run()
    torch.cuda.synchronize()
    assert float((out - _ref(A[:, :3], B, 0.5)).abs().max()) < 2e-2
    # one result row (the sdf head's d_sdf row), a column window of a wider matrix, and the scalar-load path (odd column count)
    W = (torch.randn(P, 384, generator=g) * 0.5).half().cuda()
    o1 = torch.zeros(1, 256, device="cuda")
    WgradBatch().add(A, W, o1, m=1, n=256, b_col0=128).run()
    assert float((o1 - _ref(A[:, :1], W[:, 128:384], 1.0)).abs().max()) < 2e-2
    o2 = torch.zeros(2, 256, device="cuda")
    WgradBatch().add(A, B, o2, m=2, n=100, cols_valid=100).run()
    assert float((o2[:, :100] - _ref(A[:, :2], B[:, :100], 1.0)).abs().max()) < 2e-2 and float(o2[:, 100:].abs().max()) == 0.0
    o3 = torch.zeros(3, 256, device="cuda")
    WgradBatch().add(A, B, o3, m=3, n=99, cols_valid=99).run()
    assert float((o3[:, :99] - _ref(A[:, :3], B[:, :99], 1.0)).abs().max()) < 2e-2 and float(o3[:, 99:].abs().max()) == 0.0


def test_wgrad_many_jobs_one_launch():
    """40 jobs of different lengths: the contiguous split-K ranges cross job boundaries inside CTAs."""
    from nrhints_b200.train_ops import WgradBatch
    g = torch.Generator().manual_seed(2)
    wb, want, outs = WgradBatch(), [], []
    for i in range(40):
        P = 64 * (3 + 7 * i) + (i % 5)
        a = (torch.randn(P, 256, generator=g) * 0.5).half().cuda()
        b = (torch.randn(P, 128 if i % 3 else 256, generator=g) * 0.5).half().cuda()
        o = torch.zeros(256, b.shape[1], device="cuda")
        wb.add(a, b, o, scale=1.0 / P)
        want.append(_ref(a, b, 1.0 / P)); outs.append(o)
    wb.run()
    torch.cuda.synchronize()
    for i, (o, w) in enumerate(zip(outs, want)):
        assert float((o - w).abs().max()) < 2e-3, i
