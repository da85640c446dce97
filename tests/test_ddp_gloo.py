"""CPU, world_size 2, gloo: the host-side logic of the data-parallel step (regda_b200/parallel.py, trainer.ParamArena):
image sharding, parameter broadcast, gradient-arena all-reduce == full-batch gradient, prototype-sum all-reduce, and
bench.py's rank handling of the reference arm."""
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from regda_b200 import parallel
        from regda_b200.trainer import ParamArena
        assert parallel.world_info() == (rank, world)
        # --- sharding: contiguous, disjoint, covering
        bounds = [parallel.shard_bounds(13, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 13 and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        # --- every rank builds its own (differently seeded) model; broadcast makes them rank 0's
        model = _tiny_model(100 + rank)
        arena = ParamArena(model)
        parallel.broadcast_parameters(arena.param)
        arena.sync_shadow()
        ref = _tiny_model(100)
        for p, r in zip(model.parameters(), ref.parameters()):
            assert torch.equal(p.detach(), r.detach())
        assert torch.equal(arena.param_bf16.float(), arena.param.bfloat16().float())
        # --- gradient of the GLOBAL mean loss == all-reduced sum of per-rank mean-loss gradients * 1/world
        torch.manual_seed(7)
        x = torch.randn(8, 3, 6, 6)
        y = torch.randn(8, 4, 6, 6)
        lo, hi = parallel.shard_bounds(8, rank, world)
        arena.zero_grad()
        ((model(x[lo:hi]) - y[lo:hi]) ** 2).mean().backward()
        assert all(p.grad.data_ptr() >= arena.grad.data_ptr() for p in model.parameters())      # grads live in the arena
        # the bucketed, backward-overlapped form the trainer uses: cut at the second conv, report parameters in backward order
        local = arena.grad.clone()
        bk = parallel.GradBuckets(arena.grad, arena.bucket_boundaries(("2.",)))
        assert bk.edges == [0, arena.offset_of["2.weight"], arena.numel]
        bk.begin()
        for name in ("2.bias", "2.weight"):
            bk.reached(arena.offset_of[name])
        assert len(bk.works) == 0                              # still inside the last bucket: nothing launched yet
        bk.reached(arena.offset_of["0.bias"])
        assert len(bk.works) == 1                              # the bucket behind the boundary went out
        bk.reached(arena.offset_of["0.weight"])
        bk.finish()
        assert bk.works == [] and not bk.active
        # every element reduced exactly once
        check = local.clone()
        dist.all_reduce(check)
        assert torch.equal(arena.grad, check)
        # random cuts / arrival patterns: still exactly one reduction per element
        g = torch.Generator().manual_seed(11)
        for trial in range(5):
            buf = local.clone()
            cuts = sorted(set(int(v) for v in torch.randint(1, arena.numel, (3,), generator=g)))
            b2 = parallel.GradBuckets(buf, cuts)
            b2.begin()
            for off in sorted((int(v) for v in torch.randint(0, arena.numel, (6,), generator=g)), reverse=True):
                b2.reached(off)
            b2.finish()
            assert torch.equal(buf, check), trial
        got = arena.grad * parallel.grad_scale()
        ((ref(x) - y) ** 2).mean().backward()
        want = torch.cat([torch.cat([p.grad.reshape(-1), torch.zeros((-p.numel()) % 8)]) for p in ref.parameters()])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), float((got - want).abs().max())
        # --- prototype class sums: every rank ends with the same totals
        sums = torch.full((6, 16), float(rank + 1))
        cnt = torch.full((6,), float(10 * (rank + 1)))
        parallel.allreduce_sum_(sums, cnt)
        assert torch.equal(sums, torch.full((6, 16), float(sum(range(1, world + 1))))) and float(cnt[0]) == 10 * sum(range(1, world + 1))
        assert parallel.max_over_ranks(float(rank)) == world - 1
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_shard_bounds_properties():
    sys.path.insert(0, ROOT)
    from regda_b200 import parallel
    for n in (0, 1, 7, 8, 64, 129):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_reference_arm_under_torchrun_prints_once():
    """bench.py --impl reference under a 2-rank launch: rank 0 alone runs and prints one JSON line, rank 1 exits 0."""
    import json
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "lrh", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--regions", "50"]
    env = dict(os.environ, OMP_NUM_THREADS="2", REGDA_REF_LRH_TILES="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["e2e"]["h2d_bytes_per_step"] == 0 and lines[0]["cpu_baseline"]["kind"] == "port"
