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
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_slam_b200 import dist as nd
    H, W = 6, 10
    h = nd.Handoff(torch.device("cpu"), 4, H, W)
    g = torch.Generator().manual_seed(7)
    ref = (torch.tensor([1, 3]), torch.randn(2, 7, generator=g), torch.randint(0, 255, (2, 3, H, W), dtype=torch.uint8, generator=g),
           torch.rand(2, H, W, generator=g), torch.rand(2, H, W, generator=g))
    ok = True
    if rank == 0:
        h.send(*ref)
        h.send(torch.zeros(0, dtype=torch.long), None, torch.zeros(0, 3, H, W, dtype=torch.uint8), None, None, is_last=True)
    else:
        n, last, data = h.recv()
        ok &= n == 2 and not last and all(torch.equal(a.to(b.dtype), b) for a, b in zip(data, ref))
        n, last, data = h.recv()
        ok &= n == 0 and last and data is None
    # gradient averaging across "trainers"
    class TB: pass
    tb = TB(); tb.grid_grad = torch.full((5,), float(rank + 1)); tb.mlp_grad = torch.full((3,), float(rank))
    nd.allreduce_grads(tb, None, world)
    ok &= torch.allclose(tb.grid_grad, torch.full((5,), 1.5)) and torch.allclose(tb.mlp_grad, torch.full((3,), 0.5))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_handoff_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
