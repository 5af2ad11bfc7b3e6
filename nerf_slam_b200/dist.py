"""Multi-GPU plumbing: one process per GPU, torch.distributed.

Topology for N ranks: rank 0 = SLAM front-end (sequential over frames), ranks 1..N-1 = NeRF
trainers (data-parallel over rays, gradient all-reduce in their own sub-group).  With N == 1 both
run in one process on two CUDA streams.

The keyframe hand-off replaces the reference's `.to("cpu")` + torch.multiprocessing.Queue
(slam/visual_frontends/visual_frontend.py:1355-1360, "super slow"): dirty keyframes are packed into
ONE device buffer and broadcast rank 0 -> trainers with NCCL over NVLink (device to device, no host
staging); a 4-int header (n_keyframes, H, W, is_last) precedes the payload.
Packet layout per keyframe: idx(int32 as 4 bytes) | pose cam_T_world [t, q_xyzw] fp32 (28 B) + 20 B reserved | image u8
3xHxW | idepth_up fp32 HxW | depth_cov_up fp32 HxW.  The pose travels as the front end stores it; the receiving trainer's
ingest kernel turns it into its world_T_cam record (no host copy on either side).
"""
import numpy as np
import torch
import torch.distributed as dist


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
    c2w = v[:, 4:32].contiguous().view(torch.float32).view(n, 7)           # cam_T_world [t, q_xyzw]
    o = 52
    img = v[:, o:o + 3 * H * W].reshape(n, 3, H, W); o += 3 * H * W
    idep = v[:, o:o + 4 * H * W].contiguous().view(torch.float32).view(n, H, W); o += 4 * H * W
    cov = v[:, o:o + 4 * H * W].contiguous().view(torch.float32).view(n, H, W)
    return idx, c2w, img, idep, cov


class Handoff:
    """rank 0 -> all ranks broadcast of the dirty keyframes of one SLAM tick"""

    def __init__(self, device, max_kf, H, W, group=None):
        self.device, self.H, self.W, self.group = device, H, W, group
        self.buf = torch.empty(max_kf * kf_bytes(H, W), dtype=torch.uint8, device=device)
        self.hdr = torch.zeros(4, dtype=torch.int32, device=device)
        self.max_kf = max_kf
        self.bytes_sent = 0

    def send(self, idx, poses_tq, images_u8, idepths_up, depths_cov_up, is_last=False):
        n = int(idx.shape[0])
        assert n <= self.max_kf
        self.hdr.copy_(torch.tensor([n, self.H, self.W, int(is_last)], dtype=torch.int32))
        dist.broadcast(self.hdr, src=0, group=self.group)
        if n:
            pack_keyframes(idx, poses_tq, images_u8, idepths_up, depths_cov_up, self.buf)
            dist.broadcast(self.buf[:n * kf_bytes(self.H, self.W)], src=0, group=self.group)
            self.bytes_sent += n * kf_bytes(self.H, self.W)

    def recv(self):
        """-> (n, is_last, unpacked or None); blocks this rank's host on the header"""
        dist.broadcast(self.hdr, src=0, group=self.group)
        n, H, W, last = [int(x) for x in self.hdr.cpu()]
        if n == 0:
            return 0, bool(last), None
        dist.broadcast(self.buf[:n * kf_bytes(H, W)], src=0, group=self.group)
        return n, bool(last), unpack_keyframes(self.buf, n, H, W)


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
