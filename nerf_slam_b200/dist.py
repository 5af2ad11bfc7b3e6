"""Multi-GPU plumbing: one process per GPU, torch.distributed.

Topology for N ranks: rank 0 = SLAM front-end (sequential over frames), ranks 1..N-1 = NeRF trainers (data-parallel
over rays, gradient all-reduce in their own sub-group).  With N == 1 both run in one process on two CUDA streams.

The keyframe hand-off replaces the reference's `.to("cpu")` + torch.multiprocessing.Queue
(slam/visual_frontends/visual_frontend.py:1355-1360, "super slow"): the dirty keyframes of a SLAM tick are packed into
ONE fixed-size device message and broadcast rank 0 -> trainers with NCCL over NVLink (device to device, no host
staging).  Like the reference's queue, the hand-off is ASYNCHRONOUS on both sides:
  * the SLAM rank posts the broadcast with async_op and goes on with the next frame (two message buffers; a buffer is
    re-packed only after its previous broadcast has completed — a stream-side wait, the host never blocks);
  * every trainer keeps ONE receive posted and trains freely (fusion/fusion_module.py:30-33: "don't block fusion
    waiting for input"); between training steps it polls the pending receive (`Work.is_completed()`, no
    synchronisation), ingests the keyframes when a message has landed and posts the next receive.
Message = 64-byte header (int32: n_keyframes, H, W, flags, sequence number) + `capacity` keyframe records; more dirty
keyframes than `capacity` travel as several messages (flag MORE).  Record per keyframe: idx (int32) | pose cam_T_world
[t, q_xyzw] fp32 (28 B) + 20 B reserved | image u8 3xHxW | idepth_up fp32 HxW | depth_cov_up fp32 HxW.  The pose travels
as the front end stores it; the receiving trainer's ingest kernel turns it into its world_T_cam record (no host copy).
"""
import numpy as np
import torch
import torch.distributed as dist

HDR_BYTES = 64
FLAG_LAST, FLAG_MORE, FLAG_SYNC = 1, 2, 4     # SYNC: phase boundary — trainers stop consuming and meet the sender at a barrier


def kf_bytes(H, W):
    return 4 + 48 + 3 * H * W + 4 * H * W + 4 * H * W


def pack_keyframes(idx, poses_tq, images_u8, idepths_up, depths_cov_up, out=None):
    """tensors on one device -> uint8 buffer [n * kf_bytes]"""
    n, _, H, W = images_u8.shape
    kb = kf_bytes(H, W)
    dev = images_u8.device
    if out is None:
        out = torch.empty(n * kb, dtype=torch.uint8, device=dev)
    v = out[:n * kb].view(n, kb)
    v[:, 0:4] = idx.to(torch.int32).contiguous().view(torch.uint8).view(n, 4)
    v[:, 4:32] = poses_tq.to(torch.float32).contiguous().view(n, 7).view(torch.uint8).view(n, 28)
    o = 52
    v[:, o:o + 3 * H * W] = images_u8.reshape(n, -1); o += 3 * H * W
    v[:, o:o + 4 * H * W] = idepths_up.to(torch.float32).contiguous().view(n, H * W).view(torch.uint8).view(n, -1); o += 4 * H * W
    v[:, o:o + 4 * H * W] = depths_cov_up.to(torch.float32).contiguous().view(n, H * W).view(torch.uint8).view(n, -1)
    return out[:n * kb]


def unpack_keyframes(buf, n, H, W):
    kb = kf_bytes(H, W)
    v = buf[:n * kb].view(n, kb)
    idx = v[:, 0:4].contiguous().view(torch.int32).view(n)
    tq = v[:, 4:32].contiguous().view(torch.float32).view(n, 7)           # cam_T_world [t, q_xyzw]
    o = 52
    img = v[:, o:o + 3 * H * W].reshape(n, 3, H, W); o += 3 * H * W
    idep = v[:, o:o + 4 * H * W].contiguous().view(torch.float32).view(n, H, W); o += 4 * H * W
    cov = v[:, o:o + 4 * H * W].contiguous().view(torch.float32).view(n, H, W)
    return idx, tq, img, idep, cov


class Handoff:
    """asynchronous rank 0 -> all ranks hand-off of dirty keyframes (see the module docstring)"""

    def __init__(self, device, capacity, H, W, group=None, n_buffers=2):
        self.device, self.H, self.W, self.group, self.capacity = device, H, W, group, int(capacity)
        self.msg_bytes = HDR_BYTES + self.capacity * kf_bytes(H, W)
        self.bufs = [torch.zeros(self.msg_bytes, dtype=torch.uint8, device=device) for _ in range(n_buffers)]
        self.works = [None] * n_buffers
        self.seq = 0
        self.bytes_sent = 0
        self._rx = None            # (slot, Work) of the posted receive

    # ---------------------------------------------------------------- sender (rank 0)
    def _post(self, n, flags, fill):
        slot = self.seq % len(self.bufs)
        if self.works[slot] is not None:
            self.works[slot].wait()                     # stream-side: the buffer's previous broadcast has been sent
        buf = self.bufs[slot]
        hdr = torch.tensor([n, self.H, self.W, flags, self.seq], dtype=torch.int32)
        if buf.is_cuda:
            hdr = hdr.pin_memory()
        buf[:20].copy_(hdr.view(torch.uint8), non_blocking=True)
        if n:
            fill(buf[HDR_BYTES:])
        self.works[slot] = dist.broadcast(buf, src=0, group=self.group, async_op=True)
        self.seq += 1
        self.bytes_sent += self.msg_bytes

    def send(self, idx, poses_tq, images_u8, idepths_up, depths_cov_up, is_last=False):
        """post the dirty keyframes (possibly none: only `is_last` matters then); returns immediately"""
        n = int(idx.shape[0])
        if n == 0:
            if is_last:
                self._post(0, FLAG_LAST, None)
            return
        for a in range(0, n, self.capacity):
            b = min(a + self.capacity, n)
            flags = (FLAG_MORE if b < n else 0) | (FLAG_LAST if (is_last and b == n) else 0)
            self._post(b - a, flags, lambda dst, a=a, b=b: pack_keyframes(idx[a:b], poses_tq[a:b], images_u8[a:b],
                                                                         idepths_up[a:b], depths_cov_up[a:b], dst))

    def flush(self):
        """sender: all posted messages have left (call before tearing the process group down)"""
        for w in self.works:
            if w is not None:
                w.wait()

    # ---------------------------------------------------------------- receivers (ranks > 0)
    def post_recv(self):
        if self._rx is None:
            slot = self.seq % len(self.bufs)
            self._rx = (slot, dist.broadcast(self.bufs[slot], src=0, group=self.group, async_op=True))
            self.seq += 1

    def poll(self, block=False):
        """-> None when no message has landed yet (a receive stays posted), else (n, flags, unpacked or None).
        Never synchronises the device unless `block`."""
        self.post_recv()
        slot, work = self._rx
        if not block and not work.is_completed():
            return None
        work.wait()
        self._rx = None
        buf = self.bufs[slot]
        n, H, W, flags, _ = [int(x) for x in buf[:20].cpu().view(torch.int32)]
        data = unpack_keyframes(buf[HDR_BYTES:], n, H, W) if n else None
        return n, flags, data


def allreduce_grads(tb, group, world):
    """data-parallel NeRF: sum gradients across trainer ranks (NCCL ring/tree over NVSwitch), then
    every rank applies the same Adam step.  Gradients are averaged (each rank's loss is a mean
    over its own rays)."""
    dist.all_reduce(tb.grid_grad, group=group)
    dist.all_reduce(tb.mlp_grad, group=group)
    cg = getattr(tb, "cam_grad", None)
    if cg is not None:
        dist.all_reduce(cg, group=group)
    if world > 1:
        tb.grid_grad.mul_(1.0 / world)
        tb.mlp_grad.mul_(1.0 / world)
        if cg is not None:
            cg.mul_(1.0 / world)



def send_sync(handoff, is_last=False):
    """rank 0: phase marker (no keyframes).  The caller follows it with dist.barrier(); so do the trainers."""
    handoff._post(0, FLAG_SYNC | (FLAG_LAST if is_last else 0), None)


class TrainerLoop:
    """ranks > 0: free-running NeRF training next to an asynchronous keyframe feed (fusion/fusion_module.py:30-33,
    fusion/nerf_fusion.py:291-307: "fit whenever no packet is pending").

    run_until_sync(): train; between steps poll the posted receive and ingest what has landed; return at the next
    phase marker.  With several trainers the exit is agreed through a 1-element all-reduce every `agree_every` steps
    (each trainer sees the same messages, but not at the same iteration; the gradient all-reduces need the same number
    of steps on every rank)."""

    def __init__(self, handoff, fusion, ingest, group=None, n_trainers=1, agree_every=8, device=None):
        self.h, self.nf, self.ingest, self.group, self.n_trainers, self.agree_every = handoff, fusion, ingest, group, n_trainers, agree_every
        self.flag = torch.zeros(1, dtype=torch.int32, device=device if device is not None else handoff.device)
        self.messages = 0
        self.last = False

    def _consume(self, block=False):
        m = self.h.poll(block=block)
        if m is None:
            return None
        n, flags, data = m
        self.messages += 1
        self.last |= bool(flags & FLAG_LAST)
        if n:
            self.ingest(*data)
        return bool(flags & FLAG_SYNC)

    def run_until_sync(self):
        seen, it = False, 0
        has_data = lambda: self.nf.ngp.nerf.training.n_images_for_training > 0
        while True:
            if not seen:
                got = self._consume(block=not has_data())     # nothing to train on yet: wait for the first keyframes
                if got is not None:
                    seen = got
                    if not seen:
                        continue                               # drain what is already there before the next step
            if self.n_trainers == 1:
                if seen:
                    return it
            elif it % self.agree_every == 0:
                self.flag.fill_(1 if seen else 0)
                dist.all_reduce(self.flag, op=dist.ReduceOp.MIN, group=self.group)
                if int(self.flag.item()) == 1:
                    return it
            if has_data():
                self.nf.fit_volume_once()
            it += 1
