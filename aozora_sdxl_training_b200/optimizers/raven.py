"""RavenAdamW for B200: same interface and math as the reference, moments resident in HBM, ONE kernel per step.

Mirrors ``training_utils/optimizers/raven.py`` of the reference (ctor kwargs and validation raven.py:25-42, ``step``
raven.py:89-149, ``state_dict`` / ``save_cpu_state`` / ``load_cpu_state`` / ``load_state_dict`` raven.py:151-222).
What changes is *where* the state lives and how the update runs: the reference keeps ``exp_avg`` / ``exp_avg_sq`` in
CPU RAM and streams them through a ``3 x max_numel`` fp32 GPU scratch, ~15 ATen launches and 5 copies per parameter
tensor; here they are CUDA tensors in ``momentum_dtype`` and the whole parameter list is updated by one multi-tensor
sm_100a kernel (``aoz_raven_step_mt``) at HBM bandwidth.  ``save_cpu_state()`` still returns the reference's
CPU-tensor format so ``.pt`` training states stay interchangeable.

There is no CPU fallback: parameters must be CUDA tensors.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.optim import Optimizer

from .. import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
VALID_MOMENTUM_DTYPES = [torch.float32, torch.float16, torch.bfloat16]


def raven_host_scalars(lr, betas, eps, weight_decay, debias_strength, step):
    """The float64 scalars of raven.py:101-137, packed as the 8 fp32 values the kernel consumes."""
    beta1, beta2 = betas
    wd_factor = 1.0 - lr * weight_decay if weight_decay != 0 else 1.0
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    if debias_strength < 1.0:
        bc1 = 1.0 - (1.0 - bc1) * debias_strength
        bc2 = 1.0 - (1.0 - bc2) * debias_strength
    sqrt_bc2 = math.sqrt(bc2)
    step_size = lr / bc1
    # torch divides a CUDA tensor by a python scalar as a * (1/b) with 1/b formed in fp32
    inv_sqrt_bc2 = float(np.float32(1.0) / np.float32(sqrt_bc2))
    return (beta1, 1.0 - beta1, beta2, 1.0 - beta2, eps, step_size, inv_sqrt_bc2, wd_factor)


class MultiTensorPlan:
    """Device tables for the multi-tensor kernels (chunk map is rebuilt only when the tensor shapes change)."""

    def __init__(self):
        self.key = None
        self.n_tensors = 0
        self.n_chunks = 0
        self.numel = self.chunk_start = self.chunk_tensor = None
        self.partial = None

    def ensure(self, numels, device):
        key = (tuple(numels), str(device))
        if key == self.key:
            return
        chunk = _lib.query("aoz_mt_chunk_elems")
        counts = [(n + chunk - 1) // chunk for n in numels]
        starts = np.zeros(len(numels) + 1, dtype=np.int32)
        np.cumsum(counts, out=starts[1:])
        chunk_tensor = np.repeat(np.arange(len(numels), dtype=np.int32), counts)
        self.n_tensors = len(numels)
        self.n_chunks = int(starts[-1])
        self.numel = torch.from_numpy(np.asarray(numels, dtype=np.int64)).to(device)
        self.chunk_start = torch.from_numpy(starts).to(device)
        self.chunk_tensor = torch.from_numpy(chunk_tensor).to(device)
        self.partial = torch.empty(max(self.n_chunks, 1), dtype=torch.float32, device=device)
        self.key = key


def _ptr_table(tensors, device):
    arr = np.fromiter((t.data_ptr() for t in tensors), dtype=np.uint64, count=len(tensors))
    return torch.from_numpy(arr.view(np.int64)).to(device, non_blocking=True)


class _Staging:
    """Pinned-host -> device staging of the per-step tables (pointer tables, hyper-parameters).

    Eager mode: a ring of pinned buffers per table, each guarded by a CUDA event, so the host can run ahead of the GPU
    without overwriting a buffer whose copy has not executed yet.  Under CUDA-graph capture: the copy is captured from a
    dedicated pinned buffer that is never touched again (replays re-read it), into the same device buffer."""

    RING = 4

    def __init__(self):
        self.dev = {}
        self.rings = {}
        self.keep = []

    def device_buffer(self, name):
        return self.dev[name]

    def put(self, name, array: np.ndarray, device):
        flat = np.ascontiguousarray(array).reshape(-1)
        tdtype = torch.from_numpy(flat[:0].copy()).dtype
        dev = self.dev.get(name)
        if dev is None or dev.numel() != flat.size or dev.dtype != tdtype or dev.device != device:
            dev = self.dev[name] = torch.empty(flat.size, dtype=tdtype, device=device)
            self.rings.pop(name, None)
        if torch.cuda.is_current_stream_capturing():
            host = torch.empty(flat.size, dtype=tdtype).pin_memory()
            host.numpy()[:] = flat
            self.keep.append(host)
            dev.copy_(host, non_blocking=True)
            return dev
        ring = self.rings.get(name)
        if ring is None:
            ring = self.rings[name] = dict(i=0, slots=[(torch.empty(flat.size, dtype=tdtype).pin_memory(), torch.cuda.Event())
                                                        for _ in range(self.RING)])
        host, ev = ring["slots"][ring["i"]]
        ring["i"] = (ring["i"] + 1) % self.RING
        ev.synchronize()
        host.numpy()[:] = flat
        dev.copy_(host, non_blocking=True)
        ev.record()
        return dev


class RavenAdamW(Optimizer):
    """AdamW with partial bias correction (``debias_strength``); FP32 update math, moments in ``momentum_dtype``."""

    _name = "RavenAdamW"

    def __init__(self, params, lr: float = 1e-4, betas: tuple[float, float] = (0.9, 0.98), weight_decay: float = 0.06,
                 eps: float = 1e-8, debias_strength: float = 0.9, momentum_dtype: torch.dtype = torch.bfloat16):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum_dtype not in VALID_MOMENTUM_DTYPES:
            raise ValueError(f"momentum_dtype must be one of {VALID_MOMENTUM_DTYPES}, got {momentum_dtype}")
        defaults = dict(lr=lr, betas=betas, weight_decay=weight_decay, eps=eps, debias_strength=debias_strength,
                        momentum_dtype=momentum_dtype)
        super().__init__(params, defaults)
        self.max_numel = 0
        self.param_device = None
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad:
                    if self.param_device is None:
                        self.param_device = p.device
                    self.max_numel = max(self.max_numel, p.numel())
        self._momentum_dtype = momentum_dtype
        self._plan = MultiTensorPlan()
        self._norm_out = None          # device [3]: total norm, clip coefficient, sum of squares
        self._stage = _Staging()
        self._last_items = None
        self.last_launches = 0
        self._grads_override = None

    # -- state helpers ------------------------------------------------------------------------------------
    def _new_state_tensor(self, p, dtype):
        return torch.zeros_like(p, dtype=dtype, memory_format=torch.contiguous_format)

    def _restore_state_tensor(self, tensor, p):
        return tensor.to(device=p.device, dtype=self._momentum_dtype).contiguous()

    def _grad_of(self, p):
        if self._grads_override is not None:
            return self._grads_override.get(p)
        return p.grad

    def _collect(self):
        """[(group, p, grad)] for every parameter that has a gradient, in param_groups order."""
        out = []
        for group in self.param_groups:
            for p in group["params"]:
                g = self._grad_of(p)
                if g is None:
                    continue
                if not p.is_cuda:
                    raise _lib.AozoraError(f"{self._name}: parameters must be CUDA tensors (no CPU fallback)")
                if not p.is_contiguous():
                    raise _lib.AozoraError(f"{self._name}: parameters must be contiguous")
                if not g.is_contiguous():
                    g = g.contiguous()
                out.append((group, p, g))
        return out

    # -- gradient norm / clip (train.py:2772-2781) ---------------------------------------------------------
    def _gradnorm(self, items, max_norm, emulate_bf16):
        dev = items[0][1].device
        grads = [g for _, _, g in items]
        gdt = grads[0].dtype
        if any(g.dtype != gdt for g in grads):
            raise _lib.AozoraError(f"{self._name}: gradients must share one dtype")
        self._plan.ensure([p.numel() for _, p, _ in items], dev)
        if self._norm_out is None or self._norm_out.device != dev:
            self._norm_out = torch.zeros(4, dtype=torch.float32, device=dev)
        gp = self._stage.put("norm_g", np.fromiter((g.data_ptr() for g in grads), dtype=np.uint64, count=len(grads)).view(np.int64), dev)
        pl = self._plan
        _lib.call("aoz_gradnorm_mt", pl.n_tensors, pl.n_chunks, gp.data_ptr(), pl.numel.data_ptr(), pl.chunk_start.data_ptr(),
                  pl.chunk_tensor.data_ptr(), pl.partial.data_ptr(), float(max_norm), int(emulate_bf16),
                  self._norm_out.data_ptr(), _DT[gdt], torch.cuda.current_stream().cuda_stream)
        self.last_launches += 2
        return self._norm_out

    @torch.no_grad()
    def grad_norm(self, max_norm=float("inf"), emulate_torch_dtype=True):
        """Global L2 norm of the gradients as a device tensor [3] = (norm, clip coefficient, sum of squares).
        With ``emulate_torch_dtype`` and bf16 gradients the values carry torch's bf16 rounding."""
        items = self._collect()
        if not items:
            return torch.zeros(3, device=self.param_device or "cuda")
        emulate = emulate_torch_dtype and items[0][2].dtype == torch.bfloat16
        mx = max_norm if math.isfinite(max_norm) else 3.0e38
        return self._gradnorm(items, mx, emulate)[:3]

    # -- the update -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, clip_coef: torch.Tensor | None = None):
        """One Raven update (raven.py:89-149).  ``clip_coef``: optional device scalar multiplied into every gradient
        inside the kernel (fused ``clip_grad_norm_``); the reference calls ``step()`` without it."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.last_launches = 0 if clip_coef is None else self.last_launches
        items = self._collect()                    # raises for CPU parameters before any CUDA call is made
        if not items:
            return loss
        capturing = torch.cuda.is_current_stream_capturing()
        dev = items[0][1].device
        pdt, gdt = items[0][1].dtype, items[0][2].dtype
        hyper = np.empty((len(items), 8), dtype=np.float32)
        ms, vs = [], []
        mdt = None
        for i, (group, p, g) in enumerate(items):
            if p.dtype != pdt or g.dtype != gdt:
                raise _lib.AozoraError(f"{self._name}: all parameters (and all gradients) must share one dtype")
            momentum_dtype = group.get("momentum_dtype", self._momentum_dtype)
            if mdt is None:
                mdt = momentum_dtype
            elif momentum_dtype != mdt:
                raise _lib.AozoraError(f"{self._name}: one momentum_dtype per step() call")
            state = self.state[p]
            if "step" not in state:
                if capturing:
                    raise _lib.AozoraError(f"{self._name}: run at least one eager step before CUDA-graph capture")
                state["step"] = 0
                state["exp_avg"] = self._new_state_tensor(p, momentum_dtype)
                state["exp_avg_sq"] = self._new_state_tensor(p, momentum_dtype)
            if not capturing:                      # a captured step is replayed later; advance_host_state() counts those
                state["step"] += 1
            hyper[i] = raven_host_scalars(group["lr"], group["betas"], group["eps"], group["weight_decay"],
                                          group["debias_strength"], state["step"])
            m, v = state["exp_avg"], state["exp_avg_sq"]
            if m.dtype != momentum_dtype or not m.is_cuda:
                m = state["exp_avg"] = self._restore_state_tensor(m, p).to(momentum_dtype)
                v = state["exp_avg_sq"] = self._restore_state_tensor(v, p).to(momentum_dtype)
            ms.append(m)
            vs.append(v)
        pl = self._plan
        pl.ensure([p.numel() for _, p, _ in items], dev)

        def ptrs(ts):
            return np.fromiter((t.data_ptr() for t in ts), dtype=np.uint64, count=len(ts)).view(np.int64)

        pp = self._stage.put("p", ptrs([p for _, p, _ in items]), dev)
        gp = self._stage.put("g", ptrs([g for _, _, g in items]), dev)
        mp = self._stage.put("m", ptrs(ms), dev)
        vp = self._stage.put("v", ptrs(vs), dev)
        # hyper-parameters change every step: under capture they are NOT baked into the graph but uploaded eagerly
        # before each replay by advance_host_state()
        hy = self._stage.device_buffer("hyper") if capturing else self._stage.put("hyper", hyper, dev)
        self._last_items = [(group, p) for group, p, _ in items]
        _lib.call("aoz_raven_step_mt", pl.n_tensors, pl.n_chunks, pp.data_ptr(), gp.data_ptr(), mp.data_ptr(), vp.data_ptr(),
                  pl.numel.data_ptr(), pl.chunk_start.data_ptr(), pl.chunk_tensor.data_ptr(), hy.data_ptr(),
                  0 if clip_coef is None else clip_coef.data_ptr(), _DT[pdt], _DT[gdt], _DT[mdt],
                  torch.cuda.current_stream().cuda_stream)
        self.last_launches += 1
        if not capturing:
            # the kernel wrote the parameters through raw pointers: tell torch (and the UNet's packed-weight cache)
            torch.autograd.graph.increment_version([p for _, p, _ in items])
        return loss

    @torch.no_grad()
    def clip_and_step(self, max_norm: float, emulate_torch_dtype: bool = True, grads=None):
        """Fused replacement of ``clip_grad_norm_(params, max_norm)`` + ``step()`` (train.py:2772-2783) with no host
        synchronisation: norm and coefficient stay on the device and the coefficient is applied inside the update
        kernel.  ``max_norm <= 0`` means "norm only" as in train.py:2778.  Returns the device tensor
        ``[norm, coef, sumsq]``."""
        self.last_launches = 0
        self._grads_override = grads
        try:
            items = self._collect()
            if not items:
                return None
            emulate = emulate_torch_dtype and items[0][2].dtype == torch.bfloat16
            clip = max_norm is not None and max_norm > 0
            out = self._gradnorm(items, max_norm if clip else 3.0e38, emulate)
            self.step(clip_coef=out[1:2] if clip else None)
        finally:
            self._grads_override = None
        return out[:3]

    def advance_host_state(self):
        """CUDA-graph replay support: a captured ``clip_and_step`` replays only the device work (table copies + kernels).
        This performs the HOST side of one more step -- bump the per-parameter step counters and upload the
        hyper-parameter table (lr, bias corrections) that the replayed kernel reads."""
        if not self._last_items:
            raise _lib.AozoraError(f"{self._name}: advance_host_state() before any step()")
        hyper = np.empty((len(self._last_items), 8), dtype=np.float32)
        for i, (group, p) in enumerate(self._last_items):
            state = self.state[p]
            state["step"] += 1
            hyper[i] = raven_host_scalars(group["lr"], group["betas"], group["eps"], group["weight_decay"],
                                          group["debias_strength"], state["step"])
        self._stage.put("hyper", hyper, self._last_items[0][1].device)      # eager, stream-ordered before the replay

    # -- state I/O ------------------------------------------------------------------------------------------
    # File contract (what the reference's resume path reads and writes, raven.py:151-222): ``state_dict()`` is torch's plus a
    # ``_momentum_dtype`` entry; ``save_cpu_state()`` is ``{"_momentum_dtype": dtype, k: {"step", "exp_avg_cpu",
    # "exp_avg_sq_cpu"}}`` where k counts the trainable parameters in ``param_groups`` order (entries only for parameters
    # that have state); loading assigns by k without a shape check and also accepts the older ``exp_avg`` / ``exp_avg_sq`` keys.
    def _trainable_by_index(self):
        k = 0
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad:
                    yield k, p
                    k += 1

    @staticmethod
    def _plain_step(value):
        return int(value.item()) if torch.is_tensor(value) else value

    def _note_dtype_change(self, saved_dtype, verb):
        if saved_dtype != self._momentum_dtype:
            print(f"[{self._name}] {verb} moments stored as {saved_dtype}; this optimizer keeps {self._momentum_dtype} (converted).")

    def state_dict(self):
        out = super().state_dict()
        out["_momentum_dtype"] = self._momentum_dtype
        return out

    def save_cpu_state(self):
        def host(t):
            return None if t is None else t.detach().to("cpu")

        out = {"_momentum_dtype": self._momentum_dtype}
        for k, p in self._trainable_by_index():
            st = self.state.get(p)
            if st is not None:
                out[k] = dict(step=st.get("step", 0), exp_avg_cpu=host(st.get("exp_avg")), exp_avg_sq_cpu=host(st.get("exp_avg_sq")))
        return out

    def load_cpu_state(self, cpu_state):
        def device(entry, new_key, old_key, p):
            t = entry.get(old_key, entry.get(new_key))
            return None if t is None else self._restore_state_tensor(t, p)

        for k, p in self._trainable_by_index():
            entry = cpu_state.get(k)
            if entry is None:
                continue
            self.state[p] = dict(step=self._plain_step(entry.get("step", 0)),
                                 exp_avg=device(entry, "exp_avg_cpu", "exp_avg", p),
                                 exp_avg_sq=device(entry, "exp_avg_sq_cpu", "exp_avg_sq", p))
        self._note_dtype_change(cpu_state.get("_momentum_dtype", self._momentum_dtype), "loaded")

    def load_state_dict(self, state_dict):
        rest = {k: v for k, v in state_dict.items() if k != "_momentum_dtype"}
        self._note_dtype_change(state_dict.get("_momentum_dtype", torch.float32), "loading")
        super().load_state_dict(rest)
        for _, p in self._trainable_by_index():
            st = self.state.get(p)
            if not st:
                continue
            for key in ("exp_avg", "exp_avg_sq"):
                if key in st:
                    st[key] = st[key].to(self._momentum_dtype)
            if "step" in st:
                st["step"] = self._plain_step(st["step"])
