"""CPU tests of the data-parallel host logic (world_size 2, gloo): flat layout / bucket / shard geometry, bucketed
gradient reduction, global-norm clipping, sharded Raven update == single-process update of the summed gradient, and
the gather / scatter of sharded optimizer state in the reference's checkpoint format.  The device math is injected from
the CPU oracle here (the product backend is the CUDA extension and has no CPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aozora_sdxl_training_b200.parallel import ALIGN, FlatLayout


def test_flat_layout_geometry():
    numels = [5, 1000, 17, 4096, 3, 70001]
    for world in (1, 2, 4, 8):
        L = FlatLayout(numels, world, bucket_elems=2048)
        assert all(o % ALIGN == 0 for o in L.offsets)
        assert all((e - s) % (world * ALIGN) == 0 for s, e in L.buckets)
        assert L.buckets[0][0] == 0 and all(a[1] == b[0] for a, b in zip(L.buckets, L.buckets[1:]))
        assert L.total >= L.used and L.total == L.buckets[-1][1]
        covered = [torch.zeros(n, dtype=torch.int32) for n in numels]
        for r in range(world):
            total = 0
            for (i, poff, foff, n, soff) in L.segments(r):
                assert foff == L.offsets[i] + poff
                covered[i][poff:poff + n] += 1
                total += n
            assert total <= L.shard_elems()
        assert all(bool((c == 1).all()) for c in covered)       # every element owned by exactly one rank
        for i in range(len(numels)):
            ks = L.bucket_of_param(i)
            assert ks == list(range(ks[0], ks[-1] + 1)) and len(ks) >= 1


class _OracleBackend:
    """CPU stand-in for the CUDA kernels, built from the oracle (tests only)."""

    @staticmethod
    def sumsq(seg_g, numels, plan, out3, gdt):
        out3[2] = sum(float((g.double() ** 2).sum()) for g in seg_g)

    @staticmethod
    def clip_coef(sumsq, max_norm, out2):
        nrm = float(sumsq[0]) ** 0.5
        out2[0] = nrm
        out2[1] = min(1.0, max_norm / (nrm + 1e-6))

    @staticmethod
    def raven(seg_p, seg_g, seg_m, seg_v, plan, hyper, clip_coef, pdt, gdt, mdt):
        from oracle import host_ref
        c = 1.0 if clip_coef is None else float(clip_coef[0])
        for p, g, m, v in zip(seg_p, seg_g, seg_m, seg_v):
            host_ref.raven_update_(p, g * c, m, v, **_OracleBackend.hp, step=_OracleBackend.step)


HP = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, debias_strength=0.3)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aozora_sdxl_training_b200.parallel import DataParallel
        from oracle import host_ref
        torch.manual_seed(0)
        shapes = [(5,), (40, 25), (17,), (3, 3, 8, 8), (3,), (701,)]
        model = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s)) for s in shapes])
        ref_p = [p.detach().clone() for p in model]
        ref_m = [torch.zeros_like(p) for p in ref_p]
        ref_v = [torch.zeros_like(p) for p in ref_p]
        dp = DataParallel(model, momentum_dtype=torch.float32, backend=_OracleBackend, flat_dtype=torch.float32, bucket_elems=256)
        opt = dp.make_optimizer(**HP)
        _OracleBackend.hp = HP
        assert len(dp.layout.buckets) >= 4
        for step in (1, 2, 3):
            _OracleBackend.step = step
            grads_all = [[torch.randn(s, generator=torch.Generator().manual_seed(1000 * step + 10 * r + i)) for i, s in enumerate(shapes)]
                         for r in range(world)]
            # reverse order, as the sweep produces them
            versions = [p._version for p in dp.params]
            for i in reversed(range(len(shapes))):
                if i % 2 == 0:
                    dp.grad_ready(dp.params[i], grads_all[rank][i])            # staged by copy
                else:
                    dp.grad_view(dp.params[i]).copy_(grads_all[rank][i])       # produced in place (what the GEMM kernels do)
                    dp.grad_ready(dp.params[i], dp.grad_view(dp.params[i]))
            out = dp.reduce_clip_step(opt, 0.5)
            # parameters changed behind autograd's back: version counters must move (the UNet's packed conv weights key on them)
            assert all(p._version > v for p, v in zip(dp.params, versions))
            summed = [sum(grads_all[r][i] for r in range(world)) for i in range(len(shapes))]
            norm = torch.sqrt(sum((g.double() ** 2).sum() for g in summed)).item()
            coef = min(1.0, 0.5 / (norm + 1e-6))
            assert abs(float(out[0]) - norm) <= 1e-5 * norm
            for i in range(len(shapes)):
                host_ref.raven_update_(ref_p[i], summed[i] * coef, ref_m[i], ref_v[i], **HP, step=step)
            for i, p in enumerate(dp.params):
                assert torch.allclose(p.detach(), ref_p[i], rtol=1e-6, atol=1e-8), (step, i)
        # checkpoint format: gathered state equals the single-process moments; scatter restores the shards
        st = opt.save_cpu_state()
        assert set(st) == {"_momentum_dtype"} | set(range(len(shapes)))
        for i in range(len(shapes)):
            assert st[i]["step"] == 3 and torch.allclose(st[i]["exp_avg_cpu"], ref_m[i], rtol=1e-6, atol=1e-9)
            assert st[i]["exp_avg_sq_cpu"].shape == ref_v[i].shape
        m_before = opt.m_shard.clone()
        opt.m_shard.zero_()
        opt.load_cpu_state(st)
        assert torch.equal(opt.m_shard, m_before) and opt.step_count == 3
        q.put((rank, "ok"))
    except Exception as e:                                   # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker_deferred(rank, world, port, q):
    """Deferred all-gather: after an update only the owned slices of flat_p are current; begin_forward + gates (in forward-use
    order, here: reverse registration order) restore every rank's copy to what the immediate all-gather produces."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aozora_sdxl_training_b200.parallel import DataParallel
        shapes = [(5,), (40, 25), (17,), (3, 3, 8, 8), (3,), (701,)]

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                torch.manual_seed(0)
                self.units = torch.nn.ModuleList([torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s))]) for s in shapes])

            def _forward_units(self):                # "forward" touches the units in reverse registration order
                return reversed(list(self.units))

        results = {}
        for defer in (False, True):
            net = Net()
            dp = DataParallel(net, momentum_dtype=torch.float32, backend=_OracleBackend, flat_dtype=torch.float32, bucket_elems=256,
                              defer_all_gather=defer)
            opt = dp.make_optimizer(**HP)
            _OracleBackend.hp = HP
            if defer:
                nb = len(dp.layout.buckets)
                assert sorted(dp._ag_order) == list(range(nb)) and dp._ag_order[0] == dp.param_buckets[-1][0]     # the last parameter's buckets go first
                for u in net.units:                                                              # a unit's gate covers its own buckets
                    own = {k for p in u.parameters() for k in dp.param_buckets[dp.index[p]]}
                    assert own <= set(dp._ag_order[:dp._gate_rank[u]])
            for step in (1, 2):
                _OracleBackend.step = step
                if defer:
                    dp.begin_forward()
                    for u in net._forward_units():
                        dp.gate(u)
                    dp.gate(None)
                    assert not dp._params_stale
                for i in reversed(range(len(shapes))):
                    g = torch.randn(shapes[i], generator=torch.Generator().manual_seed(1000 * step + 10 * rank + i))
                    dp.grad_ready(dp.params[i], g)
                dp.reduce_clip_step(opt, 0.5)
            if defer:
                assert dp._params_stale
                stale = dp.flat_p.clone()
                dp.gather_params()
                assert not dp._params_stale and not torch.equal(stale, dp.flat_p)
                dp.gather_params()                                                                  # idempotent
            results[defer] = dp.flat_p.clone()
        assert torch.equal(results[True], results[False])
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_deferred_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_deferred, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


@pytest.mark.timeout(180)
def test_sharded_raven_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
