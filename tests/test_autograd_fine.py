"""The interim autograd backend (nrhints_b200/autograd_fine.py) against the oracle on CPU: values and gradients of the
differentiable fine pass given the same sample positions / hints, including the second-order path through the
analytic normals (eikonal loss) and the near/far gradient of the coarse samples (camera optimisation)."""
import torch

import nrh_testlib as T
import nrhints_b200 as nb
from nrhints_b200 import autograd_fine
from oracle import nrh_oracle as orc


def _setup(R=6):
    case = dict(T.CASES["cfg1_64x32"]); case["R"] = R
    cfg = T.make_config(case)
    torch.manual_seed(3407)
    m = nb.NeuSHintRenderer(cfg)
    sd = T.make_state("sharp", cfg)
    m.load_state_dict(sd)
    rays, bg = T.case_inputs(case)
    return case, cfg, m, sd, rays, bg


def test_fine_pass_values_and_parameter_gradients_match_oracle():
    case, cfg, m, sd, rays, bg = _setup()
    ocfg = orc.OracleConfig.from_model_config(cfg)
    # oracle, autograd through its explicit reverse sweep
    sd_req = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out_o = orc.render_forward(sd_req, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"], rays["fars"],
                               background_rgb=bg)
    gt = torch.rand(case["R"], 3, generator=torch.Generator().manual_seed(1))
    loss_o = orc.training_loss(out_o, gt)
    keys = sorted(sd_req)
    grads_o = dict(zip(keys, torch.autograd.grad(loss_o, [sd_req[k] for k in keys])))
    # product-side torch fine pass fed with the oracle's sample positions / hints
    r = cfg.renderer
    inv_s = torch.exp(m.deviation_network.variance * 10.0).clip(1e-6, 1e6)
    fine = autograd_fine.render_fine(m._autograd_weights(), rays["origins"], rays["directions"], rays["pl_positions"],
                                     out_o["z_vals"].detach(), 2.0 / r.n_samples, out_o["visibilities"].detach(),
                                     out_o["specular_cue"][:, 0, :].detach(), bg, 1.0, inv_s, True)
    assert torch.allclose(fine["rgb"], out_o["rgb"], atol=2e-6)
    assert torch.allclose(fine["weights"], out_o["weights"], atol=2e-5)
    assert torch.allclose(fine["analytic_normals"], out_o["analytic_normals"], atol=2e-4)
    mine = {"rgb": fine["rgb"], "analytic_normals": fine["analytic_normals"], "relax_inside_sphere": out_o["relax_inside_sphere"]}
    loss_m = orc.training_loss(mine, gt)
    assert abs(float(loss_m) - float(loss_o)) < 1e-6
    loss_m.backward()
    for name, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        go = grads_o[name]
        denom = go.abs().max().clamp_min(1e-8)
        # the variance gradient is ONE scalar summed over all samples with heavy cancellation (|grad| ~ 6e-5 here): two fp32
        # summation orders differ by ~1.2e-7 absolute = 1.0e-3 .. 2.2e-3 relative depending on the host's thread count
        tol = 6e-3 if name == "deviation_network.variance" else 2e-3
        assert (p.grad - go).abs().max() / denom < tol, (name, float((p.grad - go).abs().max() / denom))


def test_coarse_sample_gradient_path():
    """z_vals = sort(cat(coarse(near, far), importance.detach())): the coarse entries keep their near/far gradient."""
    R, n = 3, 8
    near = torch.tensor([[1.0], [2.0], [3.0]], requires_grad=True)
    far = near.detach() + 2.0
    zc = near + (far - near) * torch.linspace(0.0, 1.0, n)[None, :]
    extra = zc.detach()[:, :4] + 0.1
    z_final, _ = torch.sort(torch.cat([zc.detach(), extra], -1), -1)
    z = autograd_fine.attach_coarse_gradient(z_final, zc)
    assert torch.equal(z.detach(), z_final)
    (z * torch.arange(1, z.shape[1] + 1)[None, :]).sum().backward()
    # reference semantics: d/d near of the sorted concat
    near2 = near.detach().clone().requires_grad_(True)
    zc2 = near2 + (far - near2) * torch.linspace(0.0, 1.0, n)[None, :]
    zs2, _ = torch.sort(torch.cat([zc2, extra], -1), -1)
    (zs2 * torch.arange(1, z.shape[1] + 1)[None, :]).sum().backward()
    assert torch.allclose(near.grad, near2.grad, atol=1e-5)


def test_reflectance_node_chain_rule(monkeypatch):
    """autograd_fine._ReflectanceF16 (block inputs, padded fp16 operand, one loss scale, fp16 adjoints between layers) against
    torch.autograd through the same network with the same fp16 operand roundings.  The library calls it makes on the GPU
    (fp16 GEMMs with fp32 accumulation, bias+ReLU epilogue, nrh_colsum_f16) are replaced by fp32 stand-ins with identical
    rounding points, so the chain rule itself is checked without a GPU."""
    import nrhints_b200.train_ops as train_ops
    orig_mm = torch.mm

    def mm(a, b, out_dtype=None):
        r = orig_mm(a.float(), b.float())
        return r if out_dtype == torch.float32 else r.half()
    monkeypatch.setattr(torch, "mm", mm)
    monkeypatch.setattr(torch, "_addmm_activation", lambda b, x, w: torch.relu(orig_mm(x.float(), w.float()) + b.float()).half())
    monkeypatch.setattr(train_ops, "colsum_f16", lambda m, scale=1.0: m.float().sum(0) * scale)

    class HostWgrad:                    # nrh_wgrad_f16 (one tcgen05 launch on the GPU): out += scale * dev_scale * A^T B, fp32 accumulation
        def __init__(self):
            self.jobs = []

        def add(self, a, b, out, scale=1.0, dev_scale=None, m=None, n=None, a_col0=0, b_col0=0, rows_valid=0, cols_valid=0):
            m = m if m is not None else a.shape[1] - a_col0
            n = n if n is not None else b.shape[1] - b_col0
            self.jobs.append((a[:, a_col0:a_col0 + m], b[:, b_col0:b_col0 + n], out, scale * (float(dev_scale) if dev_scale is not None else 1.0), m, n))
            return self

        def run(self):
            for a, b, out, sc, m, n in self.jobs:
                out[:m, :n] += orig_mm(a.float().t(), b.float()) * sc
    monkeypatch.setattr(train_ops, "WgradBatch", HostWgrad)
    g = torch.Generator().manual_seed(0)
    R, S, widths = 8, 20, [3, 27, 3, 27, 256, 9, 36]                   # the 361-wide input of the nr-hints preset
    P = R * S
    ray_block = [False, True, False, True, False, True, True]        # view / light / hint encodings are per ray
    ws = [torch.randn(256, 361, generator=g) * 0.05] + [torch.randn(256, 256, generator=g) * 0.08 for _ in range(3)] + \
         [torch.randn(3, 256, generator=g) * 0.1]
    bs = [torch.randn(256, generator=g) * 0.1 for _ in range(4)] + [torch.randn(3, generator=g) * 0.1]
    ws, bs = [w.requires_grad_(True) for w in ws], [b.requires_grad_(True) for b in bs]
    for sample_major in (False, True):
        parts = [torch.randn(R if rb else P, k, generator=g) for k, rb in zip(widths, ray_block)]
        for i in (0, 1, 2, 3, 4):                                      # position, view, normal, light, feature carry gradients
            parts[i].requires_grad_(True)
        ray_dims = (S, R, 1) if sample_major else (R, S, 0)
        y = autograd_fine._ReflectanceF16.apply(None, len(parts), ray_dims, *parts, *ws, *bs)
        cot = torch.randn(P, 3, generator=g) * 1e-3
        leaves = parts[:5] + ws + bs
        got = torch.autograd.grad((y * cot).sum(), leaves)

        def q(t):                                                      # value rounded to fp16, gradient passed straight through
            return t + (t.detach().half().float() - t.detach())

        def per_point(t, rb):                                          # expand a per-ray block to the evaluation order
            if not rb:
                return t
            return (t[None, :, :].expand(S, R, -1) if sample_major else t[:, None, :].expand(R, S, -1)).reshape(P, -1)
        h = torch.cat([per_point(t, rb) for t, rb in zip(parts, ray_block)], -1)
        for l in range(5):
            h = torch.nn.functional.linear(q(h), q(ws[l]), q(bs[l]) if l < 4 else bs[l])
            if l < 4:
                h = torch.relu(h)
        want = torch.autograd.grad((h * cot).sum(), leaves)
        # identical rounding points, but torch.mm (node) and F.linear (check) may use different host GEMM kernels: where an fp32
        # pre-activation lands on an fp16 rounding boundary one summation order flips it by one fp16 ulp (~5e-4 at |h| ~ 1), which
        # reaches the output of a few points.  Hence: nearly all points agree to fp32 noise, none is off by more than an ulp or two.
        dy = (y - h).abs()
        assert float(dy.median()) < 1e-6 and float(dy.max()) < 2e-3, (float(dy.median()), float(dy.max()))
        for a, b in zip(got, want):
            # per-ray gradients are sums of fp16-rounded per-point adjoints: a slightly wider band
            assert float((a - b).abs().max()) < 4e-3 * float(b.abs().max()), (sample_major, tuple(a.shape), float((a - b).abs().max() / b.abs().max()))
