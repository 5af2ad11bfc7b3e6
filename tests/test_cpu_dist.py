"""N>1 host logic on CPU (gloo, world_size 2): keyframe pack/unpack and the broadcast protocol."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_pack_unpack_roundtrip():
    from nerf_slam_b200 import dist as nd
    g = torch.Generator().manual_seed(0)
    n, H, W = 3, 8, 12
    idx = torch.tensor([4, 9, 2])
    c2w = torch.randn(n, 7, generator=g)                    # cam_T_world [t, q]
    img = torch.randint(0, 255, (n, 3, H, W), dtype=torch.uint8, generator=g)
    idep = torch.rand(n, H, W, generator=g); cov = torch.rand(n, H, W, generator=g)
    buf = nd.pack_keyframes(idx, c2w, img, idep, cov)
    assert buf.numel() == n * nd.kf_bytes(H, W)
    i2, c2, im2, d2, v2 = nd.unpack_keyframes(buf, n, H, W)
    assert torch.equal(i2.long(), idx) and torch.equal(c2, c2w) and torch.equal(im2, img)
    assert torch.equal(d2, idep) and torch.equal(v2, cov)


def _worker(rank, world, port, q):
    """rank 0 = sender (SLAM side), ranks 1.. = free-running trainers with a stand-in fusion object"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types
    from nerf_slam_b200 import dist as nd
    H, W = 6, 10
    h = nd.Handoff(torch.device("cpu"), 2, H, W)                 # capacity 2: the 3-keyframe tick needs two messages
    grp = dist.new_group(list(range(1, world)))
    g = torch.Generator().manual_seed(7)
    mk = lambda n: (torch.arange(n) + 1, torch.randn(n, 7, generator=g), torch.randint(0, 255, (n, 3, H, W), dtype=torch.uint8, generator=g),
                    torch.rand(n, H, W, generator=g), torch.rand(n, H, W, generator=g))
    ticks = [mk(3), mk(1), mk(2)]                                 # same on every rank (same generator seed)
    ok = True
    if rank == 0:
        h.send(*ticks[0])
        h.send(torch.zeros(0, dtype=torch.long), None, torch.zeros(0, 3, H, W, dtype=torch.uint8), None, None)   # nothing dirty: no message
        h.send(*ticks[1])
        nd.send_sync(h)
        dist.barrier()
        h.send(*ticks[2])
        nd.send_sync(h, is_last=True)
        dist.barrier()
        h.flush()
        ok &= h.seq == 6                                          # 2 + 1 + sync, 1 + sync
    else:
        got = []
        training = types.SimpleNamespace(n_images_for_training=0)
        steps = [0]

        def fit():
            steps[0] += 1
        nf = types.SimpleNamespace(ngp=types.SimpleNamespace(nerf=types.SimpleNamespace(training=training)), fit_volume_once=fit)

        def ingest(idx, tq, img, idep, cov):
            got.append((idx.clone(), tq.clone(), img.clone(), idep.clone(), cov.clone()))
            training.n_images_for_training += len(idx)
        loop = nd.TrainerLoop(h, nf, ingest, grp, world - 1, agree_every=4, device=torch.device("cpu"))
        it1 = loop.run_until_sync()
        dist.barrier()
        ok &= not loop.last and training.n_images_for_training == 4 and len(got) == 3       # 3 keyframes arrived as 2 + 1
        it2 = loop.run_until_sync()
        dist.barrier()
        ok &= loop.last and training.n_images_for_training == 6
        cat = lambda k: torch.cat([m[k] for m in got])
        for k in range(5):
            ok &= torch.equal(cat(k).to(torch.cat([t[k] for t in ticks]).dtype), torch.cat([t[k] for t in ticks]))
        # every trainer left each phase after the SAME number of steps (the gradient all-reduces would need that)
        t = torch.tensor([it1, it2, steps[0]])
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=grp); dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=grp)
        ok &= torch.equal(lo, hi) and steps[0] >= 1
    # gradient averaging across "trainers"
    class TB: pass
    tb = TB(); tb.grid_grad = torch.full((5,), float(rank + 1)); tb.mlp_grad = torch.full((3,), float(rank)); tb.cam_grad = torch.full((2,), 3.0 * rank)
    nd.allreduce_grads(tb, None, world)
    m = (world + 1) / 2.0
    ok &= torch.allclose(tb.grid_grad, torch.full((5,), m)) and torch.allclose(tb.mlp_grad, torch.full((3,), m - 1)) and torch.allclose(tb.cam_grad, torch.full((2,), 3.0 * (m - 1)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_async_handoff_and_free_running_trainers_gloo_world3():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
