"""Generate the golden fixtures tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
For every case in tests/nrh_testlib.CASES this imports the reference NeuSHintRenderer
(/root/reference/models/neus_hint_model.py), loads the seeded state_dict, runs its forward on the
seeded rays and stores inputs + RenderOutput fields + a digest of the weights.  It also runs the
oracle on the same case and asserts agreement, which is what pins oracle/nrh_oracle.py to the reference.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import nrh_testlib as T  # noqa: E402

REF = Path("/root/reference")


def import_reference():
    sys.path.insert(0, str(REF))
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))     # only extract_geometry uses it
    import models.neus_hint_model as M
    from camera.ray_utils import RayBundle
    return M, RayBundle


def ref_config(M, case):
    kw = dict(case.get("renderer", {}))
    if "normal_type" in kw:
        kw["normal_type"] = M.NormalComputationType(kw["normal_type"].value)
    if "depth_type" in kw:
        kw["depth_type"] = M.DepthComputationType(kw["depth_type"].value)
    return M.NeuSModelConfig(renderer=M.NeuSRendererConfig(**kw))


def main():
    M, RayBundle = import_reference()
    torch.set_num_threads(8)
    only = set(sys.argv[1:])                       # optional: regenerate just the named cases
    for name, case in T.CASES.items():
        if only and name not in only:
            continue
        cfg = T.make_config(case)
        sd = T.make_state(case["weights"], cfg)
        rays, bg = T.case_inputs(case)
        ref = M.NeuSHintRenderer(ref_config(M, case))
        ref.load_state_dict(sd, strict=True)
        bundle = RayBundle(origins=rays["origins"], directions=rays["directions"], pl_positions=rays["pl_positions"],
                           nears=rays["nears"], fars=rays["fars"])
        training = bool(case.get("training"))
        if training:
            torch.manual_seed(case["rng_seed"])        # the reference draws its own jitters from this stream
        # record the final sample positions the reference hands to render_core (instance-level wrapper;
        # the reference source is untouched)
        seen = {}
        inner = ref.render_core

        def spy(rays_o, rays_d, rays_pl, z_vals, *a, **k):
            seen["z_vals"] = z_vals.detach().clone()
            return inner(rays_o, rays_d, rays_pl, z_vals, *a, **k)
        ref.render_core = spy
        out = ref.forward(bundle, is_training=training, background_rgb=bg, global_step=case.get("global_step", 0))
        ref_np = T.to_np(out)
        ref_np["z_vals"] = seen["z_vals"].numpy()
        orc_np = T.to_np(T.run_oracle(case))
        stats = T.compare_outputs(orc_np, ref_np,
                                  label=f"oracle-vs-reference[{name}]", **T.TOL_ORACLE_VS_REF[case["weights"]])
        print(name, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in stats.items()})
        np.savez_compressed(HERE / f"{name}.npz", digest=np.array(T.state_digest(sd)),
                            **{"in_" + k: v.numpy() for k, v in rays.items()}, in_bg=bg.numpy(),
                            **{"out_" + k: v for k, v in ref_np.items()})
    # ---- gradients of the training loss (pipelines/base_pipeline.py:57-62) from the reference's autograd: all 46 parameter tensors
    #      and the ray inputs (camera / light optimisation, camera/ray_generator.py:105-126) --------------------------------
    for name, gt_seed in T.GRAD_CASES.items():
        if only and name + "_grads" not in only:
            continue
        case = T.CASES[name]
        cfg = T.make_config(case)
        sd = T.make_state(case["weights"], cfg)
        rays, bg = T.case_inputs(case)
        ref = M.NeuSHintRenderer(ref_config(M, case))
        ref.load_state_dict(sd, strict=True)
        leaves = {k: rays[k].clone().requires_grad_(True) for k in ("origins", "directions", "pl_positions")}
        bundle = RayBundle(origins=leaves["origins"], directions=leaves["directions"], pl_positions=leaves["pl_positions"],
                           nears=rays["nears"], fars=rays["fars"])
        torch.manual_seed(case["rng_seed"])
        out = ref.forward(bundle, is_training=True, background_rgb=bg, global_step=case["global_step"])
        gt = torch.rand(case["R"], 3, generator=torch.Generator().manual_seed(gt_seed))
        rgb_loss = torch.nn.functional.l1_loss(out.rgb, gt, reduction="sum") / (out.rgb.size(0) + 1e-5)
        gerr = (torch.linalg.norm(out.analytic_normals, ord=2, dim=-1) - 1.0) ** 2
        eik = (out.relax_inside_sphere * gerr).sum() / (out.relax_inside_sphere.sum() + 1e-5)
        loss = rgb_loss + eik * 0.1
        loss.backward()
        gfix = {"loss": np.array(float(loss)), "gt": gt.numpy()}
        for k, p in ref.named_parameters():
            g = p.grad.detach()
            gfix["norm::" + k] = np.array(float(g.norm()))
            if g.numel() <= 512:
                gfix["full::" + k] = g.numpy()
        for k, t in leaves.items():
            gfix["ray::" + k] = t.grad.detach().numpy()
        np.savez_compressed(HERE / f"{name}_grads.npz", **gfix)
        print("gradient fixture", name, ":", len([k for k in gfix if k.startswith("norm::")]), "parameters, loss", float(loss),
              "max |d origins|", float(leaves["origins"].grad.abs().max()))
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
