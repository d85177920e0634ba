"""The N>1 path of bench.py on CPU with 2 gloo ranks: ray sharding bookkeeping (each rank renders its own slice, no
data-path collective), max-over-ranks timing reduction and whole-job throughput aggregation.  The compute itself is
replaced by the oracle on a handful of rays (the CUDA library needs a GPU); what is tested is the host logic."""
import os
import socket
import sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nrh_testlib as T


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(T.ROOT))
    from nrhints_b200.workload import synthetic_rays, shard_rays
    from oracle import nrh_oracle as orc
    import nrhints_b200 as nb
    torch.set_num_threads(2)
    R_total = 8
    rays = synthetic_rays(R_total, seed=11, crop=300)
    mine = shard_rays(rays, rank, world)                      # contiguous slice, trainer/trainer.py:118 semantics
    cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(n_samples=8, n_importance_samples=8, n_shadow_samples=8,
                                                            n_shadow_importance_samples=8))
    sd = T.make_state("init", cfg)
    ocfg = orc.OracleConfig.from_model_config(cfg)
    with torch.no_grad():
        out = orc.render_forward(sd, ocfg, mine["origins"], mine["directions"], mine["pl_positions"], mine["nears"], mine["fars"],
                                 background_rgb=torch.ones(1, 3))
    # gather the slices on rank 0 and compare with the unsharded render
    parts = [torch.zeros_like(out["rgb"]) for _ in range(world)]
    dist.all_gather(parts, out["rgb"])
    t = torch.tensor([1.0 + rank], dtype=torch.float64)       # fake per-rank elapsed ms: the slowest rank defines the step
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with torch.no_grad():
            full = orc.render_forward(sd, ocfg, rays["origins"], rays["directions"], rays["pl_positions"], rays["nears"], rays["fars"],
                                      background_rgb=torch.ones(1, 3))
        q.put((torch.allclose(torch.cat(parts), full["rgb"], atol=1e-6), float(t.item()), world * mine["origins"].shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ray_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax, total = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "sharded render != unsharded render"
    assert tmax == 2.0 and total == 8


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(T.ROOT))
    from nrhints_b200.workload import synthetic_rays, shard_rays
    from nrhints_b200 import autograd_fine
    from nrhints_b200.grad_sync import allreduce_gradients
    import nrhints_b200 as nb
    torch.set_num_threads(2)
    calls = []
    real_all_reduce = dist.all_reduce
    dist.all_reduce = lambda *a, **k: (calls.append(a[0].numel()), real_all_reduce(*a, **k))[1]

    cfg = nb.NeuSModelConfig()
    m = nb.NeuSHintRenderer(cfg)                                  # parameters on the CPU; the kernels are not touched
    m.load_state_dict(T.make_state("init", cfg))
    R, S = 8, 12
    rays = synthetic_rays(R, seed=5, crop=300)

    def loss_of(batch):
        n = batch["origins"].shape[0]
        z = batch["nears"] + (batch["fars"] - batch["nears"]) * torch.linspace(0, 1, S)[None, :]
        out = autograd_fine.render_fine(m._autograd_weights(), batch["origins"], batch["directions"], batch["pl_positions"], z,
                                        2.0 / S, torch.full((n, 1), 0.5), torch.full((n, 4), 0.1), torch.ones(1, 3), 1.0,
                                        torch.exp(m.deviation_network.variance * 10.0), True)
        return out["rgb"].square().mean() + 0.1 * (out["analytic_normals"].norm(dim=-1) - 1.0).square().mean()

    loss_of(shard_rays(rays, rank, world)).backward()              # my slice of the batch
    flat = allreduce_gradients(m)                                  # mean over ranks, one collective
    got = [p.grad.clone() for p in m.parameters()]
    if rank == 0:
        m.zero_grad(set_to_none=True)
        loss_of(rays).backward()                                   # the unsharded step
        want = [p.grad for p in m.parameters()]
        err = max(float((g - w).abs().max() / (w.abs().max() + 1e-12)) for g, w in zip(got, want))
        q.put((err, calls, flat.numel(), sum(p.numel() for p in m.parameters())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_is_one_flat_collective():
    """the data-parallel step exchanges exactly one flat gradient buffer; the averaged sharded gradient
    equals the gradient of the unsharded batch (equal slices, mean losses)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, calls, nflat, nparams = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert calls == [nparams] and nflat == nparams, (calls, nflat, nparams)
    assert err < 1e-4, err


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(T.ROOT))
    from nrhints_b200.grad_sync import allreduce_flat
    g0 = torch.arange(10, dtype=torch.float32) * (rank + 1)          # group 0 flat gradient buffer of this rank
    g1 = torch.full((4,), float(rank))                               # group 1 (ray-generator parameters)
    views = [g0[2:6].view(2, 2), g1[1:3]]                            # what `p.grad` are: views into the flat buffers
    scale = allreduce_flat([g0, g1])
    if rank == 0:
        q.put((scale, g0.tolist(), g1.tolist(), views[0].tolist(), views[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_buffer_allreduce_in_place():
    """FlatAdam's gradient buffers are reduced in place (the parameters' .grad are views of them: nothing is packed or
    unpacked) and the mean is returned as a scale for the optimiser launch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scale, g0, g1, v0, v1 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale == 0.5
    assert g0 == [3.0 * i for i in range(10)] and g1 == [1.0] * 4
    assert v0 == [[6.0, 9.0], [12.0, 15.0]] and v1 == [1.0, 1.0]
