"""Adam over flat arenas (the reference's optimiser is torch.optim.Adam over model.parameters(),
trainers/train_CXR14.py:139; trainers/train_datasetDM.py:46).

`FusedAdam` keeps the reference's calling convention (`opt.zero_grad(); loss.backward(); opt.step()`,
`state_dict()` with `exp_avg` / `exp_avg_sq` / `step` per parameter) but stores parameters, moments
and -- when the backward of tedm_b200.engine produced them -- gradients as slices of single flat
fp32 buffers, so one step is ONE kernel launch and a data-parallel step is ONE all-reduce.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import native as N


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        params = [p for p in params]
        if not params:
            raise ValueError("FusedAdam: no parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam keeps one flat arena: a single parameter group")
        self._params = [p for p in self.param_groups[0]["params"]]
        dev = self._params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam runs on CUDA (sm_100a) only; there is no CPU fallback")
        if any(p.dtype != torch.float32 or p.device != dev for p in self._params):
            raise TypeError("FusedAdam: fp32 parameters on one device")
        n = sum(p.numel() for p in self._params)
        self._n = n
        pad = (n + 3) // 4 * 4
        self.flat_param = torch.zeros(pad, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(pad, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(pad, device=dev, dtype=torch.float32)
        self._grad_buf: Optional[torch.Tensor] = None
        self._offsets = []
        off = 0
        with torch.no_grad():
            for p in self._params:            # re-home every parameter into the arena (same values, same names)
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                self._offsets.append(off)
                off += k
        self._step = 0
        self.step_counter = torch.zeros(1, device=dev, dtype=torch.int32)   # device copy: the count a CUDA graph replays
        for p, o in zip(self._params, self._offsets):
            k = p.numel()
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.exp_avg[o:o + k].view(p.shape),
                             "exp_avg_sq": self.exp_avg_sq[o:o + k].view(p.shape)}

    def flat_grad(self) -> torch.Tensor:
        """The gradients as one flat buffer: zero-copy when they already are slices of one arena in parameter
        order (what tedm_b200.engine.backward produces), otherwise gathered."""
        first = self._params[0].grad
        if first is not None:
            base = first.data_ptr()
            contiguous = True
            for p, o in zip(self._params, self._offsets):
                g = p.grad
                if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * o:
                    contiguous = False
                    break
            if contiguous:
                storage_elems = first.untyped_storage().nbytes() // 4 - first.storage_offset()
                pad = (self._n + 3) // 4 * 4
                if storage_elems >= pad:
                    return first.as_strided((pad,), (1,), first.storage_offset())
        if self._grad_buf is None:
            self._grad_buf = torch.zeros_like(self.flat_param)
        for p, o in zip(self._params, self._offsets):
            k = p.numel()
            if p.grad is None:
                self._grad_buf[o:o + k].zero_()
            else:
                self._grad_buf[o:o + k].copy_(p.grad.reshape(-1))
        return self._grad_buf

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, flat_grad: Optional[torch.Tensor] = None):
        """`flat_grad`: the buffer a caller already obtained from `flat_grad()` (and e.g. all-reduced)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grp = self.param_groups[0]
        self._step += 1
        # torch.optim.Adam skips parameters whose grad is None (no moment decay, no weight decay, no movement); the one
        # flat launch below would treat them as zero-gradient, so their slices are put back afterwards
        skipped = [(o, p.numel()) for p, o in zip(self._params, self._offsets) if p.grad is None]
        if len(skipped) == len(self._params):
            skipped = [] if flat_grad is not None else skipped
        saved = [(o, k, [a[o:o + k].clone() for a in (self.flat_param, self.exp_avg, self.exp_avg_sq)]) for o, k in skipped]
        N.adam_step(self.flat_param, self.flat_grad() if flat_grad is None else flat_grad, self.exp_avg, self.exp_avg_sq, float(grp["lr"]),
                    float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]),
                    self._step, grad_scale, step_counter=self.step_counter)
        for o, k, (sp, sm, sv) in saved:
            self.flat_param[o:o + k].copy_(sp)
            self.exp_avg[o:o + k].copy_(sm)
            self.exp_avg_sq[o:o + k].copy_(sv)
        step_t = torch.tensor(float(self._step))
        for p in self._params:
            self.state[p]["step"] = step_t
        self.bump_versions()
        return loss

    def state_dict(self):
        """torch.optim.Adam's format.  Every parameter gets its OWN step tensor: internally they share one, and an aliased
        tensor would be incremented once per parameter by torch's foreach Adam after `load_state_dict`."""
        sd = super().state_dict()
        for st in sd["state"].values():
            if "step" in st:
                st["step"] = torch.tensor(float(st["step"]))
        return sd

    def load_state_dict(self, state_dict) -> None:
        """Accepts what torch.optim.Adam.state_dict() produces (the reference's checkpoints, train_CXR14.py:96-114)
        as well as its own: the moments are copied into the flat arenas and the per-parameter views rebound."""
        super().load_state_dict(state_dict)
        step = 0
        with torch.no_grad():
            for p, o in zip(self._params, self._offsets):
                k = p.numel()
                st = self.state.get(p, {})
                for name, arena in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                    view = arena[o:o + k].view(p.shape)
                    if name in st and st[name].data_ptr() != view.data_ptr():
                        view.copy_(st[name].to(view.dtype))
                    elif name not in st:
                        view.zero_()
                    st[name] = view
                step = max(step, int(float(st.get("step", 0.0))))
                self.state[p] = st
        self._step = step
        self.step_counter.fill_(step)
        step_t = torch.tensor(float(step))
        for p in self._params:
            self.state[p]["step"] = step_t

    def bump_versions(self) -> None:
        """The update went through the flat arena, not through each parameter tensor: bump every parameter's
        version counter so derived-weight caches keyed on (data_ptr, _version) notice the change."""
        torch._C._autograd._unsafe_set_version_counter(tuple(self._params), tuple(p._version + 1 for p in self._params))
