"""Shared machinery of the model classes: checkpoint-compatible parameter
containers + the handle of the CUDA implementation behind forward()."""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..synthetic import MODEL_SPECS, _ALIASES


class _Leaf(nn.Module):
    """Holds the tensors of one conv / linear / batch-norm of the checkpoint format."""

    def __init__(self, kind, shape, bias):
        super().__init__()
        if kind == "bn":
            c = shape[0]
            self.weight = nn.Parameter(torch.ones(c))
            self.bias = nn.Parameter(torch.zeros(c))
            self.register_buffer("running_mean", torch.zeros(c))
            self.register_buffer("running_var", torch.ones(c))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            self.weight = nn.Parameter(torch.empty(*shape).uniform_(-bound, bound))
            if bias:
                self.bias = nn.Parameter(torch.empty(shape[0]).uniform_(-bound, bound))


def _attach(root, path, leaf):
    parts = path.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    mod.add_module(parts[-1], leaf)


class IQModule(nn.Module):
    """nn.Module whose state_dict() has the reference's keys and shapes and whose forward
    runs in libiq_b200.so.  Inference only (eval mode, no autograd)."""

    KIND = None

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.output_channels = 40 if getattr(args, "dataset", None) == "modelnet40" else 10
        leaves = {}
        alias = _ALIASES if self.KIND in ("dgcnn", "gcnn") else {}
        for prefix, kind, shape, bias in MODEL_SPECS[self.KIND]():
            if kind == "lin" and prefix in ("linear3", "fc3"):
                shape = (self.output_channels, shape[1])
            src = alias.get(prefix)
            leaf = leaves[src] if src is not None else _Leaf(kind, shape, bias)
            leaves[prefix] = leaf
            _attach(self, prefix, leaf)
        self._handle = None
        self._handle_sig = None
        self._ws = None

    # ---------------------------------------------------------------- CUDA handle
    def _signature(self):
        return tuple((id(t), t._version, t.device) for t in self.state_dict(keep_vars=True).values())

    def _knn_k(self):
        return int(getattr(self.args, "k", 20))

    def _get_handle(self):
        sig = self._signature()
        if self._handle is not None and sig == self._handle_sig:
            return self._handle
        self._free_handle()
        lib = _lib.load()
        sd = {k: v for k, v in self.state_dict().items() if v.dtype == torch.float32}
        names = list(sd.keys())
        arrays = [np.ascontiguousarray(sd[k].detach().cpu().numpy()) for k in names]
        c_names = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
        c_ptrs = (ctypes.c_void_p * len(names))(*[a.ctypes.data for a in arrays])
        c_numel = (ctypes.c_int64 * len(names))(*[a.size for a in arrays])
        h = lib.iq_model_create(self.KIND.encode(), len(names), c_names, c_ptrs, c_numel, self._knn_k(),
                                self.output_channels)
        if not h:
            raise _lib.IQError(_lib.last_error())
        self._handle, self._handle_sig = h, sig
        return h

    def _free_handle(self):
        if self._handle is not None:
            _lib.load().iq_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._free_handle()
        except Exception:
            pass

    def set_engine(self, engine):
        """GEMM engine of the forward pass: "3xtf32" (tcgen05, default) or "fp32" (exact SIMT)."""
        code = {"3xtf32": 1, "tc": 1, "fp32": 0, "simt": 0}[engine]
        _lib.check(_lib.load().iq_model_set_engine(self._get_handle(), code))
        self._ws = None

    def set_chunk(self, chunk):
        """Clouds per internal pass (tuning knob; results do not depend on it)."""
        _lib.check(_lib.load().iq_model_set_chunk(self._get_handle(), int(chunk)))
        self._ws = None

    def set_lanes(self, lanes):
        """Chunks in flight, 1..4 (tuning knob; results do not depend on it): chunks alternate between the caller's
        stream and internal side streams so that one chunk's kernel tails overlap the next chunk's work."""
        _lib.check(_lib.load().iq_model_set_lanes(self._get_handle(), int(lanes)))
        self._ws = None

    def get_lanes(self):
        return int(_lib.load().iq_model_get_lanes(self._get_handle()))

    def _workspace(self, B, N, device):
        lib = _lib.load()
        need = lib.iq_model_workspace_bytes(self._get_handle(), B, N)
        if need < 0:
            raise _lib.IQError(_lib.last_error())
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=device)
        return self._ws

    def _run(self, x, point_major, want_aux=False, out=None):
        if self.training:
            raise RuntimeError("%s is inference-only: call .eval() first (the reference's load_model does, "
                               "tools/final_util.py:261)" % type(self).__name__)
        if not x.is_cuda:
            raise _lib.IQError("input must live on a CUDA device (iq_b200 has no CPU path)")
        if x.dtype != torch.float32:
            raise TypeError("input must be float32")
        if x.dim() != 3 or (x.shape[2] if point_major else x.shape[1]) != 3:
            raise ValueError("input must be (B,3,N)")
        x = x.contiguous()
        B = x.shape[0]
        N = x.shape[1] if point_major else x.shape[2]
        h = self._get_handle()
        ws = self._workspace(B, N, x.device)
        if out is None:
            out = torch.empty((B, self.output_channels), dtype=torch.float32, device=x.device)
        tf = crt = None
        if want_aux:
            tf = torch.empty((B, 64, 64), dtype=torch.float32, device=x.device)
            crt = torch.empty((B, 1024), dtype=torch.int64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().iq_model_forward(h, x.data_ptr(), 1 if point_major else 0, B, N, out.data_ptr(),
                                                    ws.data_ptr(), ws.numel(), tf.data_ptr() if want_aux else 0,
                                                    crt.data_ptr() if want_aux else 0,
                                                    torch.cuda.current_stream().cuda_stream))
        return (out, tf, crt) if want_aux else out

    def forward(self, x):
        """x (B,3,N) float32 CUDA -> logits (B,num_classes)."""
        return self._run(x, point_major=False)

    def forward_point_major(self, x, out=None):
        """x (B,N,3): skips the permute of cal_reward (tools/final_common.py:35)."""
        return self._run(x, point_major=True, out=out)
