"""Parameter update of the reference's training loop, fused over one flat buffer.

The reference does, per step (main/train_vlp_ddp.py:63-68 = main/train_mr.py:61-66, optimizer from main/config.py:350):

    optimizer.zero_grad(); losses.backward()
    nn.utils.clip_grad_norm_(model.parameters(), opt.grad_clip)      # grad_clip default 0.1
    optimizer.step()                                                  # torch.optim.AdamW(lr, weight_decay=wd)

`FlatAdamW` is that sequence for a `univtg_b200.plugin.Model`: the model's parameters are re-seated as views of one flat fp32
buffer laid out like the flat gradient buffer the backward kernels write (and the DDP hook all-reduces), so clipping and AdamW
are two kernel launches (univtg_adamw_step) instead of ~90 per-parameter launches.  CUDA only - there is no CPU path.
"""
import ctypes

import torch

from . import _lib


class FlatAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=0.1,
                 write_clipped_grads=False, dynamic_loss_scale=True, growth_interval=2000, max_loss_scale=65536.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm is not None else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        self.step_count = 0
        # dynamic loss scale of the fp16 backward (model.grad_scale): a step whose gradients are not finite is skipped ON THE
        # DEVICE (univtg_adamw_step); the flag is read back one step later (no host synchronisation in the step), the scale is
        # halved and the step counter corrected; after `growth_interval` good steps in a row the scale doubles again.
        self.dynamic_loss_scale = bool(dynamic_loss_scale) and model.operand_format == 0
        self.growth_interval, self.max_loss_scale = int(growth_interval), float(max_loss_scale)
        self.skipped_steps, self._good_streak = 0, 0
        self._flag_host, self._flag_event = None, None
        self._flat_p = None
        self._views = None
        self._m = self._v = self._scratch = None
        model.direct_grad = True  # gradients stay in the flat buffer; param.grad are views of it
        self._flatten()

    # -- layout ---------------------------------------------------------------------------------------------------------
    def _flatten(self):
        """(Re)seat every parameter as a view of one flat buffer that mirrors Model._grad_buffer()'s offsets."""
        model = self.model
        params = model._abi_params()
        for p in params:
            if not p.requires_grad:
                raise RuntimeError("FlatAdamW: frozen parameters are not supported (the reference trains all of them)")
        dev = model._device()
        if dev.type != "cuda":
            raise RuntimeError("univtg_b200: FlatAdamW needs the model on a CUDA device (no CPU path)")
        flat_g, _ = model._grad_buffer()
        flat_p = torch.zeros_like(flat_g)
        views, off = [], 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                v = flat_p[off:off + n].view_as(p)
                v.copy_(p.data)
                p.data = v
                views.append(v)
                off += (n + 3) // 4 * 4
        keep = self._m is not None and self._m.numel() == flat_p.numel() and self._m.device == flat_p.device
        self._flat_p, self._views = flat_p, views
        if not keep:
            self._m = torch.zeros_like(flat_p)
            self._v = torch.zeros_like(flat_p)
            self._scratch = torch.zeros(2048, dtype=torch.float32, device=dev)  # UNIVTG_ADAMW_SCRATCH_FLOATS

    def _seated(self):
        ps = self.model._abi_params()  # model.to() / load_state_dict(assign=True) move every parameter: two sentinels suffice per step
        return ps[0].data_ptr() == self._views[0].data_ptr() and ps[-1].data_ptr() == self._views[-1].data_ptr()

    # -- torch.optim-like surface ---------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """The next backward writes the flat gradient buffer from scratch (it zero-fills it itself); without a zero_grad() in
        between, further backwards accumulate into it like torch's .grad (univtg_b200/autograd.py)."""
        self.model.__dict__["_flat_grad_dirty"] = False
        return None

    @torch.no_grad()
    def step(self):
        """clip_grad_norm_(max_grad_norm) + AdamW over the flat buffers; returns the total gradient norm (device scalar)."""
        model = self.model
        if not self._seated():  # e.g. model.to(...) or load_state_dict(assign=True) replaced parameter storage
            self._flatten()
        flat_g, _ = model._grad_buffer()
        if flat_g.numel() != self._flat_p.numel():
            raise RuntimeError("FlatAdamW: gradient / parameter buffer size mismatch")
        lib = _lib.load_library()
        self._consume_overflow_flag()
        self.step_count += 1
        with torch.cuda.device(flat_g.device):
            # The update kernel also refreshes the 16-bit GEMM operand copies of the weight matrices it has just computed (the
            # packed buffer of the training format), so the next forward needs no re-packing pass over the 43 M fp32 weights;
            # only the fp32 vectors (LayerNorm terms, biases, token-type rows) are re-copied by one small launch.
            model._ensure_packed(training=True)  # allocates / fully packs once; a no-op afterwards (see below)
            fmt = model._fmt(True)
            cfg = model._cfgs[fmt]
            packed = model._packed[fmt]
            _lib.check(lib.univtg_adamw_step(_lib.ptr(self._flat_p), _lib.ptr(flat_g), _lib.ptr(self._m), _lib.ptr(self._v),
                                             flat_g.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                             self.step_count, self.max_grad_norm, int(self.write_clipped_grads),
                                             _lib.ptr(self._scratch), ctypes.byref(cfg), _lib.ptr(packed), _lib.stream_ptr()),
                       "univtg_adamw_step")
            arr = self.__dict__.get("_ptr_array")
            if arr is None or self.__dict__.get("_ptr_array_base") != self._flat_p.data_ptr():
                params = model._abi_params()
                arr = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
                self.__dict__["_ptr_array"], self.__dict__["_ptr_array_base"] = arr, self._flat_p.data_ptr()
            _lib.check(lib.univtg_pack_vectors(ctypes.byref(cfg), arr, len(arr), _lib.ptr(packed), _lib.stream_ptr()),
                       "univtg_pack_vectors")
            # the raw-pointer update does not bump autograd's version counters, so the packed-buffer key of this format is still
            # current; a second operand format (if ever packed) is stale
            for other in list(model._packed_key):
                if other != fmt:
                    model._packed_key.pop(other)
            if self.dynamic_loss_scale:
                if self._flag_host is None:
                    self._flag_host = torch.zeros(1, dtype=torch.float32).pin_memory()
                    self._flag_event = torch.cuda.Event()
                self._flag_host.copy_(self._scratch[2:3], non_blocking=True)
                self._flag_event.record()
        return self._scratch[1]

    def _consume_overflow_flag(self):
        """Outcome of the PREVIOUS step (its flag copy finished long ago: no stall): back off / grow the loss scale."""
        if not self.dynamic_loss_scale or self._flag_event is None:
            return
        self._flag_event.synchronize()
        if float(self._flag_host[0]) != 0.0:
            self.step_count -= 1  # that update never happened
            self.skipped_steps += 1
            self._good_streak = 0
            self.model.grad_scale = max(1.0, self.model.grad_scale * 0.5)
        else:
            self._good_streak += 1
            if self._good_streak >= self.growth_interval and self.model.grad_scale < self.max_loss_scale:
                self.model.grad_scale *= 2.0
                self._good_streak = 0

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self._m, "exp_avg_sq": self._v,
                "hyper": {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                          "max_grad_norm": self.max_grad_norm}}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self._m.copy_(sd["exp_avg"])
        self._v.copy_(sd["exp_avg_sq"])
        h = sd.get("hyper", {})
        self.lr = float(h.get("lr", self.lr))
        self.betas = tuple(h.get("betas", self.betas))
        self.eps = float(h.get("eps", self.eps))
        self.weight_decay = float(h.get("weight_decay", self.weight_decay))
        self.max_grad_norm = float(h.get("max_grad_norm", self.max_grad_norm))
