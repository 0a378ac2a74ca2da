"""Adam + GradScaler for the NeRF step in two launches (row f-4 of SURVEY.md section 8: "Optimizer + AMP glue on the
12.2 M-param table").

The reference trains with `torch.optim.Adam(lr 1e-2, betas (0.9, 0.99), eps 1e-15)` under `torch.cuda.amp.GradScaler`
(main_nerf.py:223, nerf/utils.py:1474-1484) and casts the whole hash table to fp16 on every forward
(gridencoder/grid.py:43-44).  Per step that is ~610 MB of HBM traffic in six or more launches around the 49 MB table.
`AmpAdam` keeps

  * a persistent fp16 shadow of every parameter tensor (what the kernels read under autocast), rewritten by the update,
  * a persistent fp16 gradient buffer per tensor that the backward kernels accumulate into (cleared by the update),

and runs `lnrf_grad_nonfinite_check` + `lnrf_adam_step` + `lnrf_amp_update` (csrc/optim.cu): GradScaler's inf check,
unscale, skip-on-inf, scale growth/backoff and torch's Adam arithmetic, with no host synchronisation, so the whole
training step stays capturable in a CUDA graph.  `state_dict()` uses torch.optim.Adam's layout so checkpoints
interchange with the reference's optimizer.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N


class AmpAdam:
    def __init__(self, model, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.0, fp16=True, init_scale=2.0 ** 16,
                 growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.fp16 = bool(fp16)
        self.growth_factor, self.backoff_factor, self.growth_interval = float(growth_factor), float(backoff_factor), int(growth_interval)
        # (owner module, parameter) in torch.optim order of NeRFNetwork.get_params (network_ff.py:139-153)
        self.owners = [(model.encoder, model.encoder.embeddings), (model.sigma_net, model.sigma_net.weights),
                       (model.color_net, model.color_net.weights)]
        dev = model.encoder.embeddings.device
        if dev.type != "cuda":
            raise RuntimeError("AmpAdam: the model must live on a CUDA device (there is no CPU path)")
        self.state = []
        for owner, p in self.owners:
            st = {"exp_avg": torch.zeros_like(p.data), "exp_avg_sq": torch.zeros_like(p.data)}
            if self.fp16:
                owner._shadow_f16 = p.data.half()
                owner._grad_f16 = torch.zeros_like(owner._shadow_f16)
            else:
                owner._shadow_f16 = owner._grad_f16 = None
            self.state.append(st)
        self.step_count = torch.ones(1, dtype=torch.float32, device=dev)       # 1-based number of the NEXT update
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scale = torch.full((1,), float(init_scale), dtype=torch.float32, device=dev) if self.fp16 else None
        self._growth_tracker = torch.zeros(1, dtype=torch.int32, device=dev) if self.fp16 else None
        self.lr_scale = torch.ones(1, dtype=torch.float32, device=dev)  # schedule factor (LambdaLR), read on the device

    # ---- GradScaler surface -------------------------------------------------------------------------------------
    def scale(self, loss):
        return loss * self._scale if self.fp16 else loss

    def get_scale(self) -> float:
        return float(self._scale.item()) if self.fp16 else 1.0

    def grads(self):
        """The gradient tensors the next step() will consume (the all-reduce of ray-sharded training acts on these)."""
        out = []
        for owner, p in self.owners:
            g = owner._grad_f16 if self.fp16 else p.grad
            if g is not None:
                out.append(g)
        return out

    def zero_grad(self, set_to_none=True):
        if not self.fp16:
            for _, p in self.owners:
                p.grad = None  # fp32 mode: autograd allocates; the update kernel has consumed them

    def _descriptors(self):
        arr = (N.OptTensor * len(self.owners))()
        keep = []
        for i, ((owner, p), st) in enumerate(zip(self.owners, self.state)):
            g = owner._grad_f16 if self.fp16 else p.grad
            if g is None:
                raise RuntimeError("AmpAdam.step(): a parameter has no gradient (call backward first)")
            if self.fp16 and p.grad is not None:
                raise RuntimeError("AmpAdam.step(): a parameter received an autograd .grad -- the fp16 mode needs the fused network "
                                   "path (NeRFNetwork.fused = True under fp16 autocast) so that gradients land in its fp16 buffers")
            keep.append(g)
            arr[i].params, arr[i].exp_avg, arr[i].exp_avg_sq = p.data.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            arr[i].grad = g.data_ptr()
            arr[i].params_f16 = owner._shadow_f16.data_ptr() if self.fp16 else None
            arr[i].n = p.numel()
            arr[i].grad_dtype = N.F16 if g.dtype == torch.float16 else N.F32
        return arr, keep

    @torch.no_grad()
    def step(self):
        lib = N.lib()
        arr, keep = self._descriptors()
        n = len(self.owners)
        st = N.stream()
        if self.fp16:
            N.check(lib.lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), n, N.ptr(self.found_inf), st))
        N.check(lib.lnrf_adam_step(C.cast(arr, C.c_void_p), n, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                   N.ptr(self._scale), N.ptr(self.found_inf), N.ptr(self.step_count), N.ptr(self.lr_scale), st))
        N.check(lib.lnrf_amp_update(N.ptr(self._scale), N.ptr(self._growth_tracker), N.ptr(self.found_inf), N.ptr(self.step_count),
                                    self.growth_factor, self.backoff_factor, self.growth_interval, st))
        del keep

    def sync_shadows(self):
        """Re-derive the fp16 shadows after the fp32 parameters were changed from outside (load_state_dict, manual edits)."""
        if self.fp16:
            for owner, p in self.owners:
                owner._shadow_f16.copy_(p.data)

    def detach(self):
        """Give the modules back to the plain torch path (drops the shadows and persistent gradient buffers)."""
        for owner, _ in self.owners:
            owner._shadow_f16 = owner._grad_f16 = None

    # ---- torch.optim.Adam-compatible checkpoint layout ------------------------------------------------------------
    def state_dict(self):
        step = float(self.step_count.item()) - 1.0
        state = {i: {"step": torch.tensor(step), "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]} for i, st in enumerate(self.state)}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "params": list(range(len(self.state)))}
        scaler = {"scale": self.get_scale(), "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                  "growth_interval": self.growth_interval, "_growth_tracker": int(self._growth_tracker.item()) if self.fp16 else 0}
        return {"state": state, "param_groups": [group], "scaler": scaler}

    def load_state_dict(self, sd):
        for i, st in enumerate(self.state):
            src = sd["state"][i]
            st["exp_avg"].copy_(src["exp_avg"])
            st["exp_avg_sq"].copy_(src["exp_avg_sq"])
        steps = [float(sd["state"][i]["step"]) for i in range(len(self.state))]
        self.step_count.fill_(steps[0] + 1.0)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), tuple(map(float, g["betas"])), float(g["eps"]), float(g["weight_decay"])
        if self.fp16 and "scaler" in sd:
            self._scale.fill_(float(sd["scaler"]["scale"]))
            self._growth_tracker.fill_(int(sd["scaler"]["_growth_tracker"]))
        self.sync_shadows()
