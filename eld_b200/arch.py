"""Drop-in for the reference's netG factory seam: `arch.unet(in_channels, out_channels) -> nn.Module`
(reference models/arch/__init__.py:6-7, models/arch/Unet.py:6-91).

The module keeps the reference's parameter names / shapes / default init (state_dict keys
conv{1..9}_{1,2}.{weight,bias}, upv{6..9}.*, conv10_1.*; Conv2d OIHW, ConvTranspose2d IOHW) so
released checkpoints load, but all parameters are views into ONE flat fp32 buffer and the forward
runs the tcgen05 engine behind the C ABI (csrc/unet_engine.cu) on NHWC bf16 activations.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib

_SPEC = [('conv1_1', 'c', 4, 32), ('conv1_2', 'c', 32, 32), ('conv2_1', 'c', 32, 64), ('conv2_2', 'c', 64, 64),
         ('conv3_1', 'c', 64, 128), ('conv3_2', 'c', 128, 128), ('conv4_1', 'c', 128, 256), ('conv4_2', 'c', 256, 256),
         ('conv5_1', 'c', 256, 512), ('conv5_2', 'c', 512, 512), ('upv6', 'd', 512, 256), ('conv6_1', 'c', 512, 256),
         ('conv6_2', 'c', 256, 256), ('upv7', 'd', 256, 128), ('conv7_1', 'c', 256, 128), ('conv7_2', 'c', 128, 128),
         ('upv8', 'd', 128, 64), ('conv8_1', 'c', 128, 64), ('conv8_2', 'c', 64, 64), ('upv9', 'd', 64, 32),
         ('conv9_1', 'c', 64, 32), ('conv9_2', 'c', 32, 32), ('conv10_1', 'o', 32, 4)]


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _EngineFunction(torch.autograd.Function):
    """The netG seam as an autograd node: forward = the training engine's forward (activations stay in its
    workspace), backward = eld_unet_backward on the incoming d(loss)/d(out) - so the reference's own
    `loss.backward(); optimizer.step()` (ELD_model.py:411-420,469-475) runs against this module unchanged, with any loss.
    Parameters enter as inputs only so that autograd routes their gradients; the math reads the flat buffer."""

    @staticmethod
    def forward(ctx, net, x, *params):
        n, _, h, w = x.shape
        eng = net._engine(n, h, w, True)
        out = torch.empty((n, net.out_channels, h, w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().eld_unet_forward(eng, net._flat.data_ptr(), x.data_ptr(), out.data_ptr(), _st()), 'eld_unet_forward')
        ctx.net, ctx.eng = net, eng
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        net = ctx.net
        (x,) = ctx.saved_tensors
        g = torch.empty_like(net._flat)
        _lib.check(_lib.load().eld_unet_backward(ctx.eng, net._flat.data_ptr(), x.data_ptr(), dout.contiguous().data_ptr(),
                                                g.data_ptr(), _st()), 'eld_unet_backward')
        grads, off = [], 0
        for p in net.parameters():
            k = p.numel()
            grads.append(g[off:off + k].view(p.shape))
            off += k
        return (None, None) + tuple(grads)       # no gradient wrt the input frame (the reference never asks for one)


class UNetSeeInDark(nn.Module):
    """B200-native UNetSeeInDark(in_channels, out_channels) for 4-channel packed raw and 3-channel sRGB frames on either
    side (ELD_model.py:377-389: --stage_in / --stage_out raw | srgb with --channels 4)."""

    def __init__(self, in_channels=4, out_channels=4):
        super().__init__()
        if in_channels not in (3, 4) or out_channels not in (3, 4):
            raise NotImplementedError('the B200 engine takes 4-channel (packed Bayer) or 3-channel (sRGB) frames; X-Trans '
                                      '(--channels 9) is out of scope (SURVEY 8a row a-X)')
        self.in_channels, self.out_channels = in_channels, out_channels
        # real torch layers, constructed in the reference ORDER, only to reproduce the default init
        # and the state_dict keys; they are never called.
        for name, kind, cin, cout in _SPEC:
            cin = in_channels if name == 'conv1_1' else cin
            cout = out_channels if name == 'conv10_1' else cout
            if kind == 'c':
                m = nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)
            elif kind == 'd':
                m = nn.ConvTranspose2d(cin, cout, 2, stride=2)
            else:
                m = nn.Conv2d(cin, cout, kernel_size=1, stride=1)
            setattr(self, name, m)
        self._flat = None
        self._flat_grad = None
        self._engines = {}
        self._ddp_ready = None
        self._flatten()

    # ---- flat parameter storage --------------------------------------------------------------------
    def _flatten(self):
        params = list(self.parameters())
        dev = params[0].device
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1).float())
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grad[off:off + n].view(p.shape)
            off += n
        self._flat, self._flat_grad = flat, grad
        self._drop_engines()

    _MAX_ENGINES = 4      # (n, h, w, train) launch plans kept alive, least recently used evicted (each owns a workspace)

    def _drop_engines(self, keep=0):
        eng = getattr(self, '_engines', None) or {}
        while len(eng) > keep:
            key = next(iter(eng))
            handle, _ws = eng.pop(key)
            try:
                _lib.load().eld_unet_destroy(handle)
            except Exception:
                pass
        self._engines = eng
        self._ddp_ready = None

    def __del__(self):
        try:
            self._drop_engines()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flatten()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)   # copy_ into the views keeps the flat buffer
        return out

    @property
    def flat_params(self):
        return self._flat

    @property
    def flat_grads(self):
        return self._flat_grad

    # ---- engine ------------------------------------------------------------------------------------
    def _engine(self, n, h, w, train):
        key = (n, h, w, bool(train))
        if key in self._engines:
            self._engines[key] = self._engines.pop(key)          # most recently used last
        else:
            self._drop_engines(keep=self._MAX_ENGINES - 1)
            lib = _lib.load()
            dev = self._flat.device
            assert dev.type == 'cuda', 'the B200 engine has no CPU path'
            assert lib.eld_unet_param_count_io(self.in_channels, self.out_channels) == self._flat.numel()
            nbytes = lib.eld_unet_workspace_bytes(n, h, w, int(train))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            handle = ctypes.c_void_p()
            _lib.check(lib.eld_unet_create_io(_lib.ctx(dev.index or 0), n, h, w, int(train), ws.data_ptr(), nbytes,
                                              self.in_channels, self.out_channels, ctypes.byref(handle)), 'eld_unet_create_io')
            self._engines[key] = (handle, ws)
        return self._engines[key][0]

    def forward(self, x):
        """x: cuda float32 NCHW [n,4,h,w] -> float32 NCHW [n,4,h,w].  Under torch.enable_grad() in training mode the call
        is an autograd node (`_EngineFunction`: shapes the training tiles accept, H % 128 == 0 and W % 256 == 0);
        otherwise plain inference."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == self.in_channels
        x = x.contiguous()
        n, _, h, w = x.shape
        if self.training and torch.is_grad_enabled() and h % 128 == 0 and w % 256 == 0:
            return _EngineFunction.apply(self, x, *self.parameters())
        out = torch.empty((n, self.out_channels, h, w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().eld_unet_forward(self._engine(n, h, w, False), self._flat.data_ptr(), x.data_ptr(),
                                               out.data_ptr(), _st()), 'eld_unet_forward')
        return out

    loss_kind = 'l1'      # 'l1' (nn.L1Loss, the reference default) or 'l2' (nn.MSELoss) - models/losses.py:29-36

    def train_step(self, x, target, loss_out=None):
        """forward + pixel loss + backward in one launch sequence.  Fills self.flat_grads (== every
        parameter's .grad) and returns (out, loss) with loss a 0-dim cuda tensor (no host sync)."""
        n, _, h, w = x.shape
        assert x.is_cuda and x.dtype == torch.float32 and x.shape[1] == self.in_channels
        assert target.shape == (n, self.out_channels, h, w) and target.dtype == torch.float32
        x, target = x.contiguous(), target.contiguous()
        out = torch.empty_like(target)
        loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().eld_unet_set_loss(self._engine(n, h, w, True), 1 if self.loss_kind == 'l2' else 0), 'eld_unet_set_loss')
        _lib.check(_lib.load().eld_unet_train_step(self._engine(n, h, w, True), self._flat.data_ptr(), x.data_ptr(),
                                                   target.data_ptr(), out.data_ptr(), self._flat_grad.data_ptr(),
                                                   loss.data_ptr(), _st()), 'eld_unet_train_step')
        return out, loss


    # ---- data parallel: bucketed all-reduce overlapped with backward ----------------------
    def grad_buckets(self):
        """[(offset, count)] of the flat gradient in backward-completion order (decoder, bottleneck, encoder, first layer)."""
        import ctypes as c
        arr = (c.c_size_t * 16)()
        k = _lib.load().eld_unet_grad_buckets_io(self.in_channels, self.out_channels, arr, 16)
        return [(int(arr[2 * i]), int(arr[2 * i + 1])) for i in range(k)]

    def train_step_ddp(self, x, target, loss_out=None, group=None, timeline=None):
        """train_step + SUM all-reduce of the flat gradient, bucket by bucket on a side stream: bucket k's NCCL kernel
        waits only for the event the engine records when that bucket is final, so the decoder and bottleneck
        gradients travel while the encoder's backward still runs; the calling stream waits for all buckets at the end
        (Adam follows).  No host synchronisation."""
        import torch.distributed as dist
        n, _, h, w = x.shape
        eng = self._engine(n, h, w, True)
        lib = _lib.load()
        if not getattr(self, '_ddp_ready', None) == eng.value:
            _lib.check(lib.eld_unet_bucket_events(eng, 1), 'eld_unet_bucket_events')
            self._ddp_ready = eng.value
            self._ddp_stream = torch.cuda.Stream(device=x.device)
            self._ddp_buckets = self.grad_buckets()
            if group is None and getattr(self, '_ddp_group', None) is None:
                # A communicator of its own for the gradient exchange, limited to a few CTAs: the all-reduce kernels run
                # UNDER the persistent one-CTA-per-SM tiles, and every SM they occupy delays a tile kernel's slowest CTA.
                # The exchange is hidden behind backward, so its own speed is worth less than the SMs it leaves alone.
                import os
                ctas = int(os.environ.get('ELD_NCCL_MAX_CTAS', '0'))
                self._ddp_group = False
                if ctas > 0:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = ctas
                    opts.config.min_ctas = min(ctas, 4)
                    self._ddp_group = dist.new_group(pg_options=opts)
        if group is None and getattr(self, '_ddp_group', None):
            group = self._ddp_group
        ev = (lambda: torch.cuda.Event(enable_timing=True)) if timeline is not None else None
        if ev:
            timeline['step_start'] = ev(); timeline['step_start'].record()
        out, loss = self.train_step(x, target, loss_out=loss_out)
        if ev:
            timeline['backward_end'] = ev(); timeline['backward_end'].record()
            timeline['buckets'] = []
        works = []
        with torch.cuda.stream(self._ddp_stream):
            for k, (off, cnt) in enumerate(self._ddp_buckets):
                _lib.check(lib.eld_unet_wait_bucket(eng, k, ctypes.c_void_p(self._ddp_stream.cuda_stream)), 'eld_unet_wait_bucket')
                if ev:
                    e0 = ev(); e0.record(self._ddp_stream)
                works.append(dist.all_reduce(self._flat_grad[off:off + cnt], group=group, async_op=True))
                if ev:
                    works[-1].wait()      # (timeline mode only) the side stream waits for this bucket so its end can be stamped
                    e1 = ev(); e1.record(self._ddp_stream)
                    timeline['buckets'].append((e0, e1, cnt * 4))
        if ev:
            for wk in works:
                wk.wait()
            timeline['allreduce_joined'] = ev(); timeline['allreduce_joined'].record()
            works = []
        # FusedAdam.step joins them bucket by bucket: Adam on the first buckets runs while the last, tiny one (conv1_*:
        # final only when backward ends, its all-reduce pure latency) is still in flight
        self._pending_allreduce = [(off, cnt, wk) for (off, cnt), wk in zip(self._ddp_buckets, works)]
        return out, loss

    def join_allreduce(self):
        """make the current stream wait for every outstanding bucket (callers that read .grad right after train_step_ddp)"""
        for _, _, wk in getattr(self, '_pending_allreduce', []) or []:
            wk.wait()
        self._pending_allreduce = []

    def _profile(self, eng, run, steps):
        import numpy as np
        lib = _lib.load()
        run()
        acc = None
        for _ in range(steps):
            lib.eld_unet_profile(eng, 1)
            run()
            cap = 512
            names = ctypes.create_string_buffer(32 * cap)
            ms = np.zeros(cap, np.float32)
            fl = np.zeros(cap, np.float64)
            by = np.zeros(cap, np.float64)
            cnt = ctypes.c_int(0)
            _lib.check(lib.eld_unet_profile_read(eng, cap, names, ms.ctypes.data, fl.ctypes.data, by.ctypes.data,
                                                 ctypes.byref(cnt)), 'eld_unet_profile_read')
            k = cnt.value
            recs = [dict(name=names.raw[32 * i:32 * i + 32].split(b'\0')[0].decode(), ms=float(ms[i]),
                         flops=float(fl[i]), bytes=float(by[i])) for i in range(k)]
            if acc is None:
                acc = recs
            else:
                for a, r in zip(acc, recs):
                    a['ms'] += r['ms']
        lib.eld_unet_profile(eng, 0)
        for a in acc:
            a['ms'] /= steps
        return acc

    def profile(self, x, target, steps=3):
        """Per-launch CUDA-event timings of `steps` training steps: list of dicts
        {name, ms (mean), flops, bytes} in launch order (see eld_unet_profile in the C ABI)."""
        n, _, h, w = x.shape
        return self._profile(self._engine(n, h, w, True), lambda: self.train_step(x, target), steps)

    def profile_forward(self, x, steps=3):
        """Same for the inference launch sequence."""
        n, _, h, w = x.shape
        return self._profile(self._engine(n, h, w, False), lambda: self.forward(x), steps)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (ELD_model.py:400-401) as ONE kernel over the flat buffers.
    Keeps `param_groups` so Engine.set_learning_rate / util.set_opt_param keep working."""

    def __init__(self, net, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(list(net.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.net = net
        self.m = torch.zeros_like(net.flat_params)
        self.v = torch.zeros_like(net.flat_params)
        self.t = 0

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        g = self.param_groups[0]
        if self.m.data_ptr() == 0 or self.m.device != self.net.flat_params.device:
            self.m = torch.zeros_like(self.net.flat_params)
            self.v = torch.zeros_like(self.net.flat_params)
        self.t += 1
        p = self.net.flat_params

        def adam(off, cnt):
            _lib.check(_lib.load().eld_adam_step(_lib.ctx(p.device.index or 0), p.data_ptr() + 4 * off,
                                                 self.net.flat_grads.data_ptr() + 4 * off, self.m.data_ptr() + 4 * off,
                                                 self.v.data_ptr() + 4 * off, cnt, float(g['lr']),
                                                 float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                                                 float(g['weight_decay']), self.t, float(grad_scale), _st()), 'eld_adam_step')
        pend = getattr(self.net, '_pending_allreduce', None)
        if pend:
            # data parallel: buckets arrive in backward-completion order and tile the buffer from its end towards its start;
            # everything but the last bucket in ONE launch, then the last (tiny) bucket when its all-reduce has landed
            self.net._pending_allreduce = []
            for _, _, wk in pend[:-1]:
                wk.wait()
            lo = min(off for off, _, _ in pend[:-1]) if len(pend) > 1 else p.numel()
            if lo < p.numel():
                adam(lo, p.numel() - lo)
            pend[-1][2].wait()
            if lo > 0:
                adam(0, lo)
        else:
            adam(0, p.numel())

    def zero_grad(self, set_to_none=False):
        self.net.flat_grads.zero_()

    # checkpoint format of torch.optim.Adam ('opt_g' in ELD_model.py:516-523)
    def state_dict(self):
        state, off = {}, 0
        for i, p in enumerate(self.net.parameters()):
            n = p.numel()
            state[i] = {'step': torch.tensor(float(self.t)), 'exp_avg': self.m[off:off + n].view(p.shape).clone(),
                        'exp_avg_sq': self.v[off:off + n].view(p.shape).clone()}
            off += n
        groups = [dict((k, v) for k, v in self.param_groups[0].items() if k != 'params')]
        groups[0]['params'] = list(range(len(state)))
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        off = 0
        for i, p in enumerate(self.net.parameters()):
            n = p.numel()
            st = sd['state'].get(i)
            if st is not None:
                self.m[off:off + n].copy_(st['exp_avg'].reshape(-1))
                self.v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
                self.t = int(float(st['step']))
            off += n
        for k, v in sd['param_groups'][0].items():
            if k != 'params':
                self.param_groups[0][k] = v


def unet(in_channels, out_channels, **kwargs):
    """models/arch/__init__.py:6-7"""
    return UNetSeeInDark(in_channels, out_channels)
