"""Ray generation: oracle vs the reference fixture, the kernels' math compiled for the host vs the
oracle (forward and hand-derived backward), and -- on the GPU -- the CUDA operators behind the RayGenerator mirror."""
import ctypes as C
import subprocess
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import nrh_testlib as T
from oracle import raygen_oracle as rgo

HERE = Path(__file__).resolve().parent
GOLD = np.load(HERE / "golden" / "raygen.npz")
FIELDS = ("origins", "directions", "pl_positions", "nears", "fars")
MODES = {"off": 0, "SO3xR3": 1, "SE3": 2}


def oracle_with_grads(case, inp, dtype):
    kw = T.raygen_oracle_kwargs(case, inp, dtype)
    leaves = {}
    for k in ("cam_pose_adjustment", "pl_adjustment"):
        if kw[k] is not None:
            kw[k] = kw[k].clone().requires_grad_(True)
            leaves[k] = kw[k]
    out = rgo.raygen_forward(**kw)
    cot = T.raygen_cotangents(case)
    grads = {k: torch.zeros_like(v) for k, v in leaves.items()}
    loss = sum((out[k] * cot[k].to(dtype)).sum() for k in FIELDS)
    if leaves and loss.requires_grad:
        for k, g in zip(leaves, torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)):
            if g is not None:
                grads[k] = g
    return {k: v.detach() for k, v in out.items()}, grads


@pytest.mark.parametrize("name", list(T.RAYGEN_CASES))
def test_oracle_matches_reference_fixture(name):
    case = T.RAYGEN_CASES[name]
    inp = T.raygen_inputs(case)
    out, grads = oracle_with_grads(case, inp, torch.float32)
    for k in FIELDS:
        np.testing.assert_allclose(out[k].numpy(), GOLD[f"{name}.out_{k}"], rtol=0, atol=2e-6, err_msg=k)
    for k, g in grads.items():
        want = GOLD[f"{name}.grad_{k}"]
        np.testing.assert_allclose(g.numpy(), want, rtol=1e-4, atol=1e-4 * max(1.0, float(np.abs(want).max())), err_msg=k)


@pytest.fixture(scope="module")
def hlib():
    out = HERE / "_build" / "libnrh_hostcheck.so"
    out.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", str(out), str(HERE / "host_harness.cpp")], check=True)
    return C.CDLL(str(out))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _np(t):
    return np.ascontiguousarray(t.numpy(), dtype=np.float32) if t is not None else None


@pytest.mark.parametrize("name", list(T.RAYGEN_CASES))
def test_kernel_math_on_host_matches_oracle(hlib, name):
    """raygen_math.cuh (what k_raygen_forward / k_raygen_backward execute per ray) against the fp64 oracle + its autograd."""
    case = T.RAYGEN_CASES[name]
    inp = T.raygen_inputs(case)
    want, want_g = oracle_with_grads(case, inp, torch.float64)
    R, cam = case["R"], inp["camera"]
    camv = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["zn"], cam["zf"]], np.float32)
    img = inp["img_indices"].numpy().astype(np.int64) if inp["img_indices"] is not None else None
    imgp = img.ctypes.data_as(C.POINTER(C.c_long)) if img is not None else None
    w, h, poses, pls = _np(inp["w_indices"]), _np(inp["h_indices"]), _np(inp["poses"]), _np(inp["pls"])
    noise = _np(inp["cam_pose_noise"]) if case["noise"] else None
    plnoise = _np(inp["pl_noise"]) if case["noise"] else None
    adj = _np(inp["cam_pose_adjustment"]) if case["cam_opt_mode"] != "off" else None
    pladj = _np(inp["pl_adjustment"]) if case["pl_opt"] else None
    o, d, pl = (np.zeros((R, 3), np.float32) for _ in range(3))
    near, far = np.zeros(R, np.float32), np.zeros(R, np.float32)
    hlib.h_raygen_forward(_fp(camv), MODES[case["cam_opt_mode"]], int(case["override_near_far"]), C.c_long(R), _fp(w), _fp(h), imgp,
                          _fp(poses), _fp(pls), _fp(noise), _fp(plnoise), _fp(adj), _fp(pladj), _fp(o), _fp(d), _fp(pl), _fp(near), _fp(far))
    got = dict(origins=o, directions=d, pl_positions=pl, nears=near[:, None], fars=far[:, None])
    for k in FIELDS:
        np.testing.assert_allclose(got[k], want[k].numpy(), rtol=0, atol=3e-6, err_msg=k)
        np.testing.assert_allclose(got[k], GOLD[f"{name}.out_{k}"], rtol=0, atol=3e-6, err_msg=k + " (reference fixture)")
    if adj is None and pladj is None:
        return
    cot = {k: _np(v) for k, v in T.raygen_cotangents(case).items()}
    d_adj = np.zeros_like(adj) if adj is not None else None
    d_pl = np.zeros_like(pladj) if pladj is not None else None
    hlib.h_raygen_backward(_fp(camv), MODES[case["cam_opt_mode"]], int(case["override_near_far"]), C.c_long(R), _fp(w), _fp(h), imgp,
                           _fp(poses), _fp(noise), _fp(adj), _fp(cot["origins"]), _fp(cot["directions"]), _fp(cot["pl_positions"]),
                           _fp(cot["nears"]), _fp(cot["fars"]), _fp(d_adj), _fp(d_pl))
    for k, g in (("cam_pose_adjustment", d_adj), ("pl_adjustment", d_pl)):
        if g is None:
            continue
        wg = want_g[k].numpy()
        np.testing.assert_allclose(g, wg, rtol=2e-4, atol=2e-4 * max(1.0, float(np.abs(wg).max())), err_msg=k)
        ref = GOLD[f"{name}.grad_{k}"]
        np.testing.assert_allclose(g, ref, rtol=3e-4, atol=3e-4 * max(1.0, float(np.abs(ref).max())), err_msg=k + " (reference fixture)")


def _bundle(inp, dev):
    img = inp["img_indices"]
    return SimpleNamespace(img_indices=img[:, None].to(dev) if img is not None else None, h_indices=inp["h_indices"][:, None].to(dev),
                           w_indices=inp["w_indices"][:, None].to(dev), poses=inp["poses"].to(dev), pls=inp["pls"].to(dev))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(T.RAYGEN_CASES))
def test_cuda_raygen_matches_oracle_and_fixture(name):
    import nrhints_b200 as nb
    case = T.RAYGEN_CASES[name]
    inp = T.raygen_inputs(case)
    cam = inp["camera"]
    cfg = nb.RayGeneratorConfig(override_near_far_from_sphere=case["override_near_far"], cam_opt_mode=case["cam_opt_mode"],
                                pl_opt=case["pl_opt"], cam_position_noise_std=0.01 if case["noise"] else 0.0,
                                cam_orientation_noise_std=0.01 if case["noise"] else 0.0,
                                pl_position_noise_std=0.01 if case["noise"] else 0.0)
    gen = nb.RayGenerator(nb.CameraModel(**cam), inp["n_cameras"], cfg).cuda()
    with torch.no_grad():
        if case["cam_opt_mode"] != "off":
            gen.cam_pose_adjustment.copy_(inp["cam_pose_adjustment"])
        if case["pl_opt"]:
            gen.pl_adjustment.copy_(inp["pl_adjustment"])
        if case["noise"]:
            gen.cam_pose_noise.copy_(inp["cam_pose_noise"]); gen.pl_noise.copy_(inp["pl_noise"])
    out = gen(_bundle(inp, "cuda"))
    want, want_g = oracle_with_grads(case, inp, torch.float64)
    for k in FIELDS:
        got = getattr(out, k).detach().cpu().numpy()
        np.testing.assert_allclose(got, want[k].numpy(), rtol=0, atol=3e-6, err_msg=k)
        np.testing.assert_allclose(got, GOLD[f"{name}.out_{k}"], rtol=0, atol=3e-6, err_msg=k + " (reference fixture)")
    params = dict(gen.named_parameters())
    if not params:
        assert not out.origins.requires_grad
        return
    cot = T.raygen_cotangents(case)
    loss = sum((getattr(out, k) * cot[k].cuda()).sum() for k in FIELDS)
    if loss.requires_grad:
        loss.backward()
    for k, p in params.items():
        g = p.grad.cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        ref = GOLD[f"{name}.grad_{k}"]
        np.testing.assert_allclose(g, want_g[k].numpy(), rtol=3e-4, atol=3e-4 * max(1.0, float(np.abs(ref).max())), err_msg=k)
        np.testing.assert_allclose(g, ref, rtol=3e-4, atol=3e-4 * max(1.0, float(np.abs(ref).max())), err_msg=k + " (reference fixture)")


@pytest.mark.gpu
def test_cuda_raygen_feeds_the_renderer_and_reaches_the_pose_parameters():
    """ray generator -> renderer -> loss.backward(): the camera-registration path of pipelines/base_pipeline.py:71-91 end to end on
    the CUDA operators; gradient of the image loss w.r.t. cam_pose_adjustment against the all-oracle computation."""
    import nrhints_b200 as nb
    from oracle import nrh_oracle as orc
    case = dict(R=24, cam_opt_mode="SO3xR3", pl_opt=True, noise=False, override_near_far=True, same_image=True, seed=11)
    inp = T.raygen_inputs(case)
    inp["h_indices"] = torch.randint(300, 500, (case["R"],), generator=torch.Generator().manual_seed(5)).float()
    inp["w_indices"] = torch.randint(300, 500, (case["R"],), generator=torch.Generator().manual_seed(6)).float()
    mcfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_samples=16, n_importance_samples=16, n_shadow_samples=16,
                                                             n_shadow_importance_samples=16))
    torch.manual_seed(3407)
    m = nb.NeuSHintRenderer(mcfg, mlp_impl="fp32").cuda()
    for p in m.parameters():
        p.requires_grad_(False)
    gen = nb.RayGenerator(nb.CameraModel(**inp["camera"]), inp["n_cameras"],
                          nb.RayGeneratorConfig(cam_opt_mode="SO3xR3", pl_opt=True)).cuda()
    with torch.no_grad():
        gen.cam_pose_adjustment.copy_(inp["cam_pose_adjustment"]); gen.pl_adjustment.copy_(inp["pl_adjustment"])
    gt = torch.rand(case["R"], 3, generator=torch.Generator().manual_seed(7)).cuda()
    bg = torch.ones(1, 3).cuda()
    out = m(gen(_bundle(inp, "cuda")), is_training=False, background_rgb=bg)
    loss = torch.nn.functional.l1_loss(out.rgb, gt, reduction="sum") / (out.rgb.size(0) + 1e-5)
    loss.backward()
    # oracle: same thing in torch on the CPU
    kw = T.raygen_oracle_kwargs(case, inp, torch.float32)
    adj = kw["cam_pose_adjustment"].clone().requires_grad_(True); pla = kw["pl_adjustment"].clone().requires_grad_(True)
    kw["cam_pose_adjustment"], kw["pl_adjustment"] = adj, pla
    rays = rgo.raygen_forward(**kw)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    o_out = orc.render_forward(sd, orc.OracleConfig.from_model_config(mcfg), rays["origins"], rays["directions"], rays["pl_positions"],
                               rays["nears"], rays["fars"], background_rgb=torch.ones(1, 3))
    o_loss = torch.nn.functional.l1_loss(o_out["rgb"], gt.cpu(), reduction="sum") / (case["R"] + 1e-5)
    g_adj, g_pl = torch.autograd.grad(o_loss, [adj, pla], allow_unused=True)
    assert abs(float(loss) - float(o_loss)) < 1e-4
    for got, want in ((gen.cam_pose_adjustment.grad, g_adj), (gen.pl_adjustment.grad, g_pl)):
        want = want if want is not None else torch.zeros_like(got.cpu())
        scale = max(float(want.abs().max()), 1e-6)
        assert float((got.cpu() - want).abs().max()) < 3e-2 * scale + 1e-6, (got.cpu(), want)


def test_constructor_matches_the_reference_state_dict_and_rng_draws():
    """Same state_dict keys / shapes as the reference RayGenerator and, for the same torch seed, bit-identical noise buffers
    (the constructor draws torch.normal in the reference's order, camera/ray_generator.py:62-73)."""
    import nrhints_b200 as nb
    torch.manual_seed(123)
    gen = nb.RayGenerator(nb.CameraModel(H=8, W=8, cx=4.0, cy=4.0, fx=10.0, fy=10.0, zn=0.1, zf=10.0), 5,
                          nb.RayGeneratorConfig(cam_opt_mode="SE3", pl_opt=True, cam_position_noise_std=0.02,
                                                cam_orientation_noise_std=0.01, pl_position_noise_std=0.03))
    sd = gen.state_dict()
    assert list(sd.keys()) == [str(k) for k in GOLD["ctor_seed123.keys"]]
    for k, v in sd.items():
        ref = GOLD[f"ctor_seed123.{k}"]
        assert tuple(v.shape) == ref.shape, k
        if k == "pl_noise":
            np.testing.assert_array_equal(v.numpy(), ref)                      # a plain torch.normal draw
        else:
            np.testing.assert_allclose(v.numpy(), ref, rtol=0, atol=1e-7, err_msg=k)   # exp_map_SE3 of the draw


@pytest.mark.gpu
def test_negative_image_indices_wrap_like_torch_indexing():
    """The reference indexes its per-image tables with torch indexing (camera/ray_generator.py:92-98,:121-126): index -k means
    image n - k.  (Indices outside [-n, n) raise an IndexError there; here RayGenerator.validate_indices asserts the range on the
    device -- a device-side assert poisons the CUDA context, so that branch is not exercised in this process.)"""
    import nrhints_b200 as nb
    name = next(n for n, c in T.RAYGEN_CASES.items() if c["cam_opt_mode"] != "off" and c["noise"])
    case = T.RAYGEN_CASES[name]
    inp = T.raygen_inputs(case)
    cfg = nb.RayGeneratorConfig(override_near_far_from_sphere=case["override_near_far"], cam_opt_mode=case["cam_opt_mode"],
                                pl_opt=case["pl_opt"], cam_position_noise_std=0.01, cam_orientation_noise_std=0.01,
                                pl_position_noise_std=0.01)
    gen = nb.RayGenerator(nb.CameraModel(**inp["camera"]), inp["n_cameras"], cfg).cuda()
    with torch.no_grad():
        gen.cam_pose_adjustment.copy_(inp["cam_pose_adjustment"])
        if case["pl_opt"]:
            gen.pl_adjustment.copy_(inp["pl_adjustment"])
        gen.cam_pose_noise.copy_(inp["cam_pose_noise"]); gen.pl_noise.copy_(inp["pl_noise"])
        a = gen(_bundle(inp, "cuda"))
        neg = dict(inp); neg["img_indices"] = inp["img_indices"] - inp["n_cameras"]
        b = gen(_bundle(neg, "cuda"))
    for k in FIELDS:
        assert torch.equal(getattr(a, k), getattr(b, k)), k
