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
        self._handle_device = None
        self._knobs = {}
        self._ws = None

    # ---------------------------------------------------------------- CUDA handle
    def _signature(self):
        return tuple((id(t), t._version, t.device) for t in self.state_dict(keep_vars=True).values())

    def _knn_k(self):
        return int(getattr(self.args, "k", 20))

    def _home_device(self):
        """The CUDA device the library's copy of the weights lives on: the module's own device when it is a CUDA
        device, else the current one."""
        for t in self.parameters():
            if t.is_cuda:
                return t.device
            break
        return torch.device("cuda", torch.cuda.current_device())

    def _get_handle(self, device=None):
        device = torch.device(device) if device is not None else (self._handle_device or self._home_device())
        sig = self._signature()
        if self._handle is not None and sig == self._handle_sig and device == self._handle_device:
            return self._handle
        self._free_handle()
        lib = _lib.load()
        sd = {k: v for k, v in self.state_dict().items() if v.dtype == torch.float32}
        names = list(sd.keys())
        arrays = [np.ascontiguousarray(sd[k].detach().cpu().numpy()) for k in names]
        c_names = (ctypes.c_char_p * len(names))(*[n.encode() for n in names])
        c_ptrs = (ctypes.c_void_p * len(names))(*[a.ctypes.data for a in arrays])
        c_numel = (ctypes.c_int64 * len(names))(*[a.size for a in arrays])
        with torch.cuda.device(device):                       # iq_model_create uploads to the current device
            h = lib.iq_model_create(self.KIND.encode(), len(names), c_names, c_ptrs, c_numel, self._knn_k(),
                                    self.output_channels)
        if not h:
            raise _lib.IQError(_lib.last_error())
        self._handle, self._handle_sig, self._handle_device = h, sig, device
        self._ws = None
        # tuning knobs survive a re-creation of the handle (new weights, another device)
        for name, fn in (("engine", lib.iq_model_set_engine), ("chunk", lib.iq_model_set_chunk),
                         ("lanes", lib.iq_model_set_lanes)):
            if name in self._knobs:
                _lib.check(fn(h, self._knobs[name]))
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

    def _set_knob(self, name, value, fn):
        self._knobs[name] = value
        if self._handle is not None:
            _lib.check(fn(self._handle, value))
        self._ws = None

    def set_engine(self, engine):
        """GEMM engine of the forward pass: "3xtf32" / "tc" (tcgen05 on two-term operand splits -- fp16 pairs for the DGCNN /
        GCNN tensor products, tf32 pairs elsewhere; default) or "fp32" (exact SIMT)."""
        self._set_knob("engine", {"3xtf32": 1, "tc": 1, "fp32": 0, "simt": 0}[engine], _lib.load().iq_model_set_engine)

    def set_chunk(self, chunk):
        """Clouds per internal pass (tuning knob; results do not depend on it)."""
        if int(chunk) < 1:
            raise ValueError("chunk must be >= 1")
        self._set_knob("chunk", int(chunk), _lib.load().iq_model_set_chunk)

    def set_lanes(self, lanes):
        """Chunks in flight, 1..4 (tuning knob; results do not depend on it): chunks alternate between the caller's
        stream and internal side streams so that one chunk's kernel tails overlap the next chunk's work."""
        if not 1 <= int(lanes) <= 4:
            raise ValueError("lanes must be 1..4")
        self._set_knob("lanes", int(lanes), _lib.load().iq_model_set_lanes)

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

    def _run(self, x, point_major, want_aux=False, out=None, masked_to=None):
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
        lib = _lib.load()
        # everything the library allocates or launches for this call belongs to the input's device: the folded
        # weights (created on first use), the workspace, the kernels and the stream
        with torch.cuda.device(x.device):
            h = self._get_handle(x.device)
            ws = self._workspace(B, N, x.device)
            if out is None:
                out = torch.empty((B, self.output_channels), dtype=torch.float32, device=x.device)
            elif (not isinstance(out, torch.Tensor) or out.dtype != torch.float32 or out.device != x.device
                  or tuple(out.shape) != (B, self.output_channels) or not out.is_contiguous()):
                raise ValueError("out must be a contiguous float32 (%d, %d) tensor on %s" % (B, self.output_channels, x.device))
            stream = torch.cuda.current_stream(x.device).cuda_stream
            if masked_to is not None:
                if want_aux:
                    raise ValueError("masked_to is not available together with the PointNet aux outputs")
                if (not isinstance(masked_to, torch.Tensor) or masked_to.dtype != torch.float32 or masked_to.numel() != 3
                        or masked_to.device != x.device or not masked_to.is_contiguous()):
                    raise ValueError("masked_to must be a contiguous float32 tensor of 3 values on %s" % x.device)
                _lib.check(lib.iq_model_forward_coalitions(h, x.data_ptr(), 1 if point_major else 0, B, N, masked_to.data_ptr(),
                                                           out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
                return out
            tf = crt = None
            if want_aux:
                tf = torch.empty((B, 64, 64), dtype=torch.float32, device=x.device)
                crt = torch.empty((B, 1024), dtype=torch.int64, device=x.device)
            _lib.check(lib.iq_model_forward(h, x.data_ptr(), 1 if point_major else 0, B, N, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), tf.data_ptr() if want_aux else 0,
                                            crt.data_ptr() if want_aux else 0, stream))
        return (out, tf, crt) if want_aux else out

    def last_row_fraction(self):
        """Rows evaluated / rows of the batch in the last forward with masked_to (1.0 when nothing collapsed)."""
        return float(_lib.load().iq_model_last_row_fraction(self._get_handle()))

    def last_buckets(self):
        """{points per evaluated cloud: clouds} of the last forward (a plain forward: {N: B})."""
        buf = (ctypes.c_int64 * 64)()
        n = _lib.load().iq_model_last_buckets(self._get_handle(), buf, 64)
        return {128 * (t + 1): int(buf[t]) for t in range(min(n, 64)) if buf[t]}

    def forward(self, x):
        """x (B,3,N) float32 CUDA -> logits (B,num_classes)."""
        return self._run(x, point_major=False)

    def forward_point_major(self, x, out=None, masked_to=None):
        """x (B,N,3): skips the permute of cal_reward (tools/final_common.py:35).  masked_to (3,) float32 CUDA: the
        location the coalition masks moved the absent regions to (`center`); the library then evaluates every cloud on
        its kept points plus a few copies of it (iq_model_forward_coalitions) -- same logits, fewer rows."""
        return self._run(x, point_major=True, out=out, masked_to=masked_to)

    def forward_coalitions(self, x, masked_to, out=None):
        """x (B,3,N) like forward(), for clouds built by the reference's masking rule around `masked_to`."""
        return self._run(x, point_major=False, out=out, masked_to=masked_to)
