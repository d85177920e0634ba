"""The reflectance network's training kernels on tcgen05 (csrc/color_train_tc.inc: nrh_color_train_forward / _backward) against a
torch evaluation of the same network with the kernels' operand roundings (fp16 operands, fp32 accumulation): hidden activations,
pre-sigmoid output, the adjoint chain dz_3..dz_0 and the input adjoint; ragged point counts."""
import pytest
import torch

import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200.train_ops import color_train_backward, color_train_forward

pytestmark = pytest.mark.gpu


def _q(t):
    return t.half().float()


@pytest.mark.parametrize("P", [128, 1000, 20000])
def test_color_train_forward_and_backward(P):
    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg, mlp_impl="tcgen05"); m.load_state_dict(T.make_state("sharp", cfg)); m.cuda()
    cn = m.color_network
    Ws = [getattr(cn, f"lin{l}").effective_weight().detach() for l in range(5)]
    bs = [getattr(cn, f"lin{l}").bias.detach() for l in range(5)]
    g = torch.Generator().manual_seed(P)
    x = torch.zeros(P, 384)
    x[:, :361] = torch.randn(P, 361, generator=g) * 0.7
    x16 = x.half().cuda()
    acts, y = color_train_forward(m, x16)
    torch.cuda.synchronize()
    # reference with the kernel's roundings: weights x64 rounded to fp16, activations x16 rounded to fp16
    h = x16.float()
    want_acts = []
    for l in range(4):
        w16 = _q(Ws[l] * 64.0) / 64.0
        if l == 0:
            w16 = torch.nn.functional.pad(w16, (0, 384 - 361))
        z = h @ w16.t() + bs[l]
        a = _q(torch.relu(z) * 16.0)                 # stored x16
        want_acts.append(a)
        h = a / 16.0
    want_y = h @ Ws[4].t() + bs[4]
    for l in range(4):
        d = (acts[l].float() - want_acts[l]).abs().max() / want_acts[l].abs().max()
        assert float(d) < 2e-3, (l, float(d))
    assert float((y[:, :3] - want_y).abs().max()) < 2e-3 * float(want_y.abs().max()) + 1e-4
    # ---- backward ----
    dy = (torch.randn(P, 3, generator=g) * 1e-3).cuda()
    S = torch.tensor([2.0 ** 16], device="cuda")
    dz, dy16, dx = color_train_backward(m, dy, S, acts)
    torch.cuda.synchronize()
    assert float((dy16[:, :3].float() - _q(dy * S)).abs().max()) == 0.0 and float(dy16[:, 3:].abs().max()) == 0.0
    a_saved = [a.float() for a in acts]
    dzl = _q((dy * S) @ Ws[4]) * (a_saved[3] > 0)            # dz_3
    dzl = _q(dzl)
    want_dz = {3: dzl}
    for l in (3, 2, 1):
        w16 = _q(Ws[l] * 64.0) / 64.0
        dzl = _q((dzl @ w16) * (a_saved[l - 1] > 0))
        want_dz[l - 1] = dzl
    w0 = torch.nn.functional.pad(_q(Ws[0] * 64.0) / 64.0, (0, 384 - 361))
    want_dx = want_dz[0] @ w0
    for l in range(4):
        d = (dz[l].float() - want_dz[l]).abs().max() / want_dz[l].abs().max()
        assert float(d) < 5e-3, ("dz", l, float(d))
    assert float((dx.float() - want_dx).abs().max() / want_dx.abs().max()) < 5e-3
    assert float(dx[:, 361:].abs().max()) == 0.0
