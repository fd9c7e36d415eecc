"""FusedAdam: multi-tensor Adam/AdamW with built-in global-norm clipping, for the SQuAD fine-tune step
(SURVEY.md 8f rank 1).  Stands in for `apex.optimizers.FusedAdam(params, lr, bias_correction=False)` plus the
`GradientClipper(max_grad_norm=1.0)` of PyTorch/LanguageModeling/BERT/run_squad.py:703-724,969-975,1092-1099: one
grad-norm/found_inf pass and one fused apply pass over a device-resident tensor table (csrc/lamb.cu, dle_adam_step).

apex is not vendored in the reference tree (version unpinned), so the arithmetic is the published Adam / AdamW update
    m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2 ;  p <- p - lr ( m^ / (sqrt(v^) + eps) + wd p )      (adam_w_mode)
with m^, v^ bias-corrected only if bias_correction; parity is pinned against torch.optim.AdamW and the C oracle.
16-bit (bf16) parameters get fp32 masters, as amp O2 does.
"""
import ctypes

import torch

from . import _lib as L
from .lamb import FusedLAMBAMP


class FusedAdam(FusedLAMBAMP):
    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True, weight_decay=0.0,
                 amsgrad=False, set_grad_none=True, max_grad_norm=0.0, clip_eps=1e-6):
        super().__init__(params, lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                         amsgrad=amsgrad, adam_w_mode=adam_w_mode, grad_averaging=True, set_grad_none=set_grad_none,
                         max_grad_norm=max_grad_norm)
        self.clip_eps = clip_eps

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        loss = closure() if closure is not None else None
        self._ensure_plan()
        if self._plan is None:
            return loss
        device = self.param_groups[0]["params"][0].device
        for gi, group in enumerate(self.param_groups):
            lr = group['lr']
            if isinstance(lr, torch.Tensor):
                self._lr_dev[gi].copy_(lr.reshape(()), non_blocking=True)
            else:
                self._lr_dev[gi].fill_(float(lr))
        scale = grad_scaler._get_scale_async() if (grad_scaler is not None and grad_scaler.is_enabled()) else None
        L.check(L.load().dle_adam_step(self._plan, ctypes.c_void_p(0 if scale is None else scale.data_ptr()),
                                       float(self.defaults['max_grad_norm']), float(self.clip_eps), self.adam_w_mode,
                                       ctypes.c_void_p(self._found_inf.data_ptr()), ctypes.c_void_p(self._global_grad_norm.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "dle_adam_step")
        L.launch_count["n"] += 2
        from . import ops
        ops.weight_epoch["n"] += 1
        if grad_scaler is not None and grad_scaler.is_enabled():
            grad_scaler._per_optimizer_states[id(self)]["found_inf_per_device"] = {device: self._found_inf}
        return loss
