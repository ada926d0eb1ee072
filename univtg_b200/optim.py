"""Parameter update of the reference's training loop, fused over one flat buffer.

The reference does, per step (main/train_vlp_ddp.py:63-68 = main/train_mr.py:61-66, optimizer from main/config.py:350):

    optimizer.zero_grad(); losses.backward()
    nn.utils.clip_grad_norm_(model.parameters(), opt.grad_clip)      # grad_clip default 0.1
    optimizer.step()                                                  # torch.optim.AdamW(lr, weight_decay=wd)

`FlatAdamW` is that sequence for a `univtg_b200.plugin.Model`: the model's parameters are re-seated as views of one flat fp32
buffer laid out like the flat gradient buffer the backward kernels write (and the DDP hook all-reduces), so clipping and AdamW
are two kernel launches (univtg_adamw_step) instead of ~90 per-parameter launches.  CUDA only - there is no CPU path.
"""
import torch

from . import _lib


class FlatAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_grad_norm=0.1,
                 write_clipped_grads=False):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm is not None else 0.0
        self.write_clipped_grads = bool(write_clipped_grads)
        self.step_count = 0
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
            self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)

    def _seated(self):
        return all(p.data_ptr() == v.data_ptr() for p, v in zip(self.model._abi_params(), self._views))

    # -- torch.optim-like surface ---------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        # the backward zero-fills the flat gradient buffer itself; nothing to do per parameter
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
        self.step_count += 1
        with torch.cuda.device(flat_g.device):
            _lib.check(lib.univtg_adamw_step(_lib.ptr(self._flat_p), _lib.ptr(flat_g), _lib.ptr(self._m), _lib.ptr(self._v),
                                             flat_g.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                             self.step_count, self.max_grad_norm, int(self.write_clipped_grads),
                                             _lib.ptr(self._scratch), _lib.stream_ptr()), "univtg_adamw_step")
        model._packed_key = {}  # parameters changed behind autograd's version counters: repack the 16-bit operands
        return self._scratch[1]

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
