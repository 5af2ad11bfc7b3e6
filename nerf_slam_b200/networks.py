"""Host-side mirror of the reference's network modules — same architecture and state_dict keys,
so the reference's `droid.pth` loads unchanged:

  BasicEncoder   networks/modules/extractor.py:118-198 (ResidualBlock :6-55)            [A1]
  ConvGRU        networks/modules/gru.py:5-32                                           [A5]
  GraphAgg       networks/droid_net.py:44-75                                            [A5]
  UpdateModule   networks/droid_net.py:78-150                                           [A5]

Inference-only (the demo runs with autograd disabled, examples/slam_demo.py:198).  Parameters are
plain tensors; `conv2d` is the single place where a library convolution is executed (encoders, and
the per-layer reference form of the update operator that the fused tensor-core operator in conv.py is
tested against).  Activations are channels-last fp16; torch_scatter.scatter_mean of the reference is
an index_add here.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

def conv2d(x, w, b, stride=1, padding=0, act=None):
    """x [N,C,H,W] (channels_last fp16/fp32), w [O,C,kh,kw], b [O]; act in {None,'relu','sigmoid','tanh'}"""
    y = F.conv2d(x, w, b, stride=stride, padding=padding)
    if act == "relu":
        y = F.relu(y, inplace=True)
    elif act == "sigmoid":
        y = torch.sigmoid(y)
    elif act == "tanh":
        y = torch.tanh(y)
    return y


class _Params:
    """tiny parameter container with nn.Module-compatible state_dict keys"""

    def __init__(self):
        self._p = OrderedDict()

    def add_conv(self, name, cin, cout, k, gen, device="cpu"):
        fan_out = cout * k * k
        w = torch.randn(cout, cin, k, k, generator=gen) * (2.0 / fan_out) ** 0.5   # kaiming_normal fan_out
        bound = 1.0 / (cin * k * k) ** 0.5
        b = (torch.rand(cout, generator=gen) * 2 - 1) * bound
        self._p[name + ".weight"] = w.to(device)
        self._p[name + ".bias"] = b.to(device)

    def w(self, name):
        return self._p[name + ".weight"], self._p[name + ".bias"]

    def state_dict(self):
        return OrderedDict(self._p)

    def load_state_dict(self, sd, prefix="", strict=True):
        missing = []
        for k in self._p:
            if prefix + k in sd:
                v = sd[prefix + k]
                if tuple(v.shape) != tuple(self._p[k].shape):
                    raise RuntimeError(f"shape mismatch for {prefix + k}: {tuple(v.shape)} vs {tuple(self._p[k].shape)}")
                self._p[k] = v.to(self._p[k].device, self._p[k].dtype)
            else:
                missing.append(prefix + k)
        if strict and missing:
            raise RuntimeError(f"missing keys: {missing}")
        return missing

    def to(self, device=None, dtype=None, channels_last=True):
        for k, v in self._p.items():
            v = v.to(device=device, dtype=dtype)
            if channels_last and v.dim() == 4:
                v = v.contiguous(memory_format=torch.channels_last)
            self._p[k] = v
        return self


def _inorm(x):
    return F.instance_norm(x)


def _cl(y):
    return y if y.is_contiguous(memory_format=torch.channels_last) else y.contiguous(memory_format=torch.channels_last)


def _inorm_stats(y, out=None):
    """per-(image, channel) sum / sum of squares of a channels-last fp16 tensor (csrc/inorm.cu).
    out: already-zeroed [B,C,2] fp32 view (one arena zeroed once per forward) or None to allocate + zero here"""
    from . import _lib
    B, C, H, W = y.shape
    st = out if out is not None else torch.empty(B, C, 2, dtype=torch.float32, device=y.device)
    _lib.check(_lib.load().nslam_inorm_stats(_lib.ptr(y), _lib.ptr(st), B, H * W, C, 0 if out is not None else 1,
                                             _lib.stream_ptr()), "inorm_stats")
    return st


def _inorm_apply(y, st, relu=True, res=None, res_st=None):
    """in place: y <- relu?(IN(y)); with res: y <- relu(res' + y), res' = IN(res) if res_st is given"""
    from . import _lib
    B, C, H, W = y.shape
    _lib.check(_lib.load().nslam_inorm_apply(_lib.ptr(y), _lib.ptr(st), _lib.ptr(res), _lib.ptr(res_st), _lib.ptr(y),
                                             B, H * W, C, 1e-5, int(relu), _lib.stream_ptr()), "inorm_apply")
    return y


class BasicEncoder(_Params):
    """7x7/2 conv -> 3 stages x 2 residual blocks (32, 64/2, 128/2) -> 1x1 conv.  [A1]"""
    DIM = 32

    def __init__(self, output_dim=128, norm_fn="instance", gen=None):
        super().__init__()
        gen = gen or torch.Generator().manual_seed(0)
        self.norm = _inorm if norm_fn == "instance" else (lambda x: x)
        assert norm_fn in ("instance", "none")
        D = self.DIM
        self.add_conv("conv1", 3, D, 7, gen)
        self.blocks = []
        cin = D
        for li, (dim, stride) in enumerate([(D, 1), (2 * D, 2), (4 * D, 2)], start=1):
            for bi, (ci, st) in enumerate([(cin, stride), (dim, 1)]):
                name = f"layer{li}.{bi}"
                self.add_conv(name + ".conv1", ci, dim, 3, gen)
                self.add_conv(name + ".conv2", dim, dim, 3, gen)
                if st != 1:
                    self.add_conv(name + ".downsample.0", ci, dim, 1, gen)
                self.blocks.append((name, st))
            cin = dim
        self.add_conv("conv2", 4 * D, output_dim, 1, gen)

    def __call__(self, x):
        """x [b, n, 3, H, W] normalised image -> [b, n, C, H/8, W/8]"""
        b, n, c, h, w = x.shape
        x = x.reshape(b * n, c, h, w).contiguous(memory_format=torch.channels_last)
        wt, bs = self.w("conv1")
        x = x.to(wt.dtype)
        if self.norm is _inorm and x.is_cuda and wt.dtype == torch.float16:
            return self._forward_fused_norm(x).view(b, n, -1, h // 8, w // 8)
        x = F.relu(self.norm(conv2d(x, wt, bs, stride=2, padding=3)), inplace=True)
        for name, st in self.blocks:
            y = F.relu(self.norm(conv2d(x, *self.w(name + ".conv1"), stride=st, padding=1)), inplace=True)
            y = F.relu(self.norm(conv2d(y, *self.w(name + ".conv2"), stride=1, padding=1)), inplace=True)
            if st != 1:
                x = self.norm(conv2d(x, *self.w(name + ".downsample.0"), stride=st, padding=0))
            x = F.relu(x + y, inplace=True)
        x = conv2d(x, *self.w("conv2"))
        return x.view(b, n, x.shape[1], x.shape[2], x.shape[3])


    def _forward_fused_norm(self, x):
        """same network with the instance norms, ReLUs and residual adds on the fused NHWC kernels
        (csrc/inorm.cu): per conv one statistics pass + one apply pass instead of ~8 library launches.
        A per-channel bias in front of an instance norm cancels exactly (the norm subtracts the channel mean),
        so those convolutions run without their bias (saves one elementwise launch each)."""
        B = x.shape[0]
        arena = torch.zeros(16, B, 128, 2, dtype=torch.float32, device=x.device)      # one memset for all norm layers
        slot = [0]

        def stats(t):
            v = arena[slot[0]].view(-1)[:B * t.shape[1] * 2].view(B, t.shape[1], 2)
            slot[0] += 1
            return _inorm_stats(t, out=v)
        wt, _ = self.w("conv1")
        x = _cl(conv2d(x, wt, None, stride=2, padding=3))
        _inorm_apply(x, stats(x))
        for name, st in self.blocks:
            y = _cl(conv2d(x, self.w(name + ".conv1")[0], None, stride=st, padding=1))
            _inorm_apply(y, stats(y))
            z = _cl(conv2d(y, self.w(name + ".conv2")[0], None, stride=1, padding=1))
            if st != 1:
                d = _cl(conv2d(x, self.w(name + ".downsample.0")[0], None, stride=st, padding=0))
                _inorm_apply(z, stats(z), res=d, res_st=stats(d))
            else:
                _inorm_apply(z, stats(z), res=x)
            x = z
        return conv2d(x, *self.w("conv2"))


class UpdateModule(_Params):
    """corr/flow encoders + ConvGRU(128, 320) + delta/weight heads + GraphAgg.  [A5]"""

    def __init__(self, gen=None):
        super().__init__()
        gen = gen or torch.Generator().manual_seed(1)
        cor_planes = 4 * (2 * 3 + 1) ** 2
        self.add_conv("corr_encoder.0", cor_planes, 128, 1, gen)
        self.add_conv("corr_encoder.2", 128, 128, 3, gen)
        self.add_conv("flow_encoder.0", 4, 128, 7, gen)
        self.add_conv("flow_encoder.2", 128, 64, 3, gen)
        self.add_conv("weight.0", 128, 128, 3, gen)
        self.add_conv("weight.2", 128, 2, 3, gen)
        self.add_conv("delta.0", 128, 128, 3, gen)
        self.add_conv("delta.2", 128, 2, 3, gen)
        for nme in ("convz", "convr", "convq"):
            self.add_conv("gru." + nme, 128 + 320, 128, 3, gen)
        for nme in ("w", "convz_glo", "convr_glo", "convq_glo"):
            self.add_conv("gru." + nme, 128, 128, 1, gen)
        self.add_conv("agg.conv1", 128, 128, 3, gen)
        self.add_conv("agg.conv2", 128, 128, 3, gen)
        self.add_conv("agg.eta.0", 128, 1, 3, gen)
        self.add_conv("agg.upmask.0", 128, 8 * 8 * 9, 1, gen)

    def gru(self, net, inp_parts):
        """ConvGRU.forward (networks/modules/gru.py:19-32)"""
        inp = torch.cat(inp_parts, dim=1)
        net_inp = torch.cat([net, inp], dim=1)
        b, c, h, w = net.shape
        glo = conv2d(net, *self.w("gru.w"), act="sigmoid") * net
        glo = glo.float().mean(dim=(2, 3), keepdim=True).to(net.dtype)
        z = torch.sigmoid(conv2d(net_inp, *self.w("gru.convz"), padding=1) + conv2d(glo, *self.w("gru.convz_glo")))
        r = torch.sigmoid(conv2d(net_inp, *self.w("gru.convr"), padding=1) + conv2d(glo, *self.w("gru.convr_glo")))
        q = torch.tanh(conv2d(torch.cat([r * net, inp], dim=1), *self.w("gru.convq"), padding=1) +
                       conv2d(glo, *self.w("gru.convq_glo")))
        return (1 - z) * net + z * q

    def agg(self, net, ii):
        """GraphAgg.forward (networks/droid_net.py:59-75); net [num,128,ht,wd]"""
        _, ix = torch.unique(ii, return_inverse=True)
        K = int(ix.max().item()) + 1 if ix.numel() else 0
        x = conv2d(net, *self.w("agg.conv1"), padding=1, act="relu")
        s = torch.zeros(K, *x.shape[1:], dtype=torch.float32, device=x.device)
        s.index_add_(0, ix, x.float())
        cnt = torch.zeros(K, dtype=torch.float32, device=x.device).index_add_(0, ix, torch.ones_like(ix, dtype=torch.float32))
        x = (s / cnt.view(-1, 1, 1, 1)).to(net.dtype).contiguous(memory_format=torch.channels_last)
        x = conv2d(x, *self.w("agg.conv2"), padding=1, act="relu")
        eta = F.softplus(conv2d(x, *self.w("agg.eta.0"), padding=1).float())
        upmask = conv2d(x, *self.w("agg.upmask.0"))
        return 0.01 * eta[:, 0], upmask

    def __call__(self, net, inp, corr, flow=None, ii=None, jj=None):
        """UpdateModule.forward (networks/droid_net.py:118-150).
        net, inp [1,num,128,ht,wd]; corr [1,num,196,ht,wd]; flow [1,num,4,ht,wd] or None"""
        batch, num, ch, ht, wd = net.shape
        dt = self._p["gru.w.weight"].dtype
        cl = dict(memory_format=torch.channels_last)
        if flow is None:
            flow = torch.zeros(batch, num, 4, ht, wd, device=net.device, dtype=dt)
        net = net.reshape(batch * num, -1, ht, wd).to(dt).contiguous(**cl)
        inp = inp.reshape(batch * num, -1, ht, wd).to(dt).contiguous(**cl)
        corr = corr.reshape(batch * num, -1, ht, wd).to(dt).contiguous(**cl)
        flow = flow.reshape(batch * num, -1, ht, wd).to(dt).contiguous(**cl)
        corr = conv2d(corr, *self.w("corr_encoder.0"), act="relu")
        corr = conv2d(corr, *self.w("corr_encoder.2"), padding=1, act="relu")
        flow = conv2d(flow, *self.w("flow_encoder.0"), padding=3, act="relu")
        flow = conv2d(flow, *self.w("flow_encoder.2"), padding=1, act="relu")
        net = self.gru(net, [inp, corr, flow])
        d = conv2d(net, *self.w("delta.0"), padding=1, act="relu")
        delta = conv2d(d, *self.w("delta.2"), padding=1)
        wgt = conv2d(net, *self.w("weight.0"), padding=1, act="relu")
        weight = conv2d(wgt, *self.w("weight.2"), padding=1, act="sigmoid")
        delta = delta.permute(0, 2, 3, 1)[..., :2].contiguous().view(batch, num, ht, wd, 2)
        weight = weight.permute(0, 2, 3, 1)[..., :2].contiguous().view(batch, num, ht, wd, 2)
        out_net = net.view(batch, num, -1, ht, wd)
        if ii is not None:
            eta, upmask = self.agg(net, ii.to(net.device))
            return out_net, delta, weight, eta.view(batch, -1, ht, wd), upmask.view(batch, -1, 8 * 8 * 9, ht, wd)
        return out_net, delta, weight


def load_droid_weights(path):
    """key remap + 3->2 channel slice of RaftVisualFrontend.load_weights
    (slam/visual_frontends/visual_frontend.py:1051-1068)"""
    sd = torch.load(path, map_location="cpu")
    out = OrderedDict()
    for k, v in sd.items():
        k = k.replace("module.", "").replace("fnet.", "feature_net.").replace("cnet.", "context_net.") \
             .replace("update.", "update_net.")
        out[k] = v
    for k in ("update_net.weight.2.weight", "update_net.weight.2.bias", "update_net.delta.2.weight",
              "update_net.delta.2.bias"):
        out[k] = out[k][:2]
    return out
