"""Shared plumbing of the nn.Module drop-ins: turn a module's state_dict into an a2m_model handle
(BatchNorm folded, weights packed to bf16 on the device) and keep it in sync with the parameters."""
import ctypes

import torch
import torch.nn as nn

from . import _cabi


class NativeHandle:
    """Owns one a2m_model*; destroyed with the Python object.  `block`: (kind, in_channels, out_channels, leaky) makes
    it a stand-alone building block (a2m_block_create) instead of a generator section (a2m_model_create)."""

    def __init__(self, state, device, block=None):
        descs = (_cabi.TensorDesc * len(state))()
        keep = []                                   # tensors (possibly converted copies) alive during create
        for i, (name, t) in enumerate(state.items()):
            if t.dtype == torch.int64:
                dtype = 1
            else:
                dtype = 0
                if t.dtype != torch.float32:
                    t = t.to(torch.float32)
            t = t.detach().to(device).contiguous()
            keep.append(t)
            descs[i].name = name.encode()
            descs[i].data = t.data_ptr()
            descs[i].dtype = dtype
            descs[i].ndim = t.dim()
            for k, s in enumerate(t.shape):
                descs[i].shape[k] = s
        out = ctypes.c_void_p()
        with torch.cuda.device(device):
            if block is None:
                _cabi.check(_cabi.lib().a2m_model_create(descs, len(state), device.index, ctypes.byref(out)))
            else:
                kind, cin, cout, leaky = block
                _cabi.check(_cabi.lib().a2m_block_create(kind, descs, len(state), cin, cout, int(leaky), device.index,
                                                         ctypes.byref(out)))
        self.ptr = out
        self.device = device

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _cabi.lib().a2m_model_destroy(ptr)
            except Exception:
                pass


class NativeModule(nn.Module):
    """Base of the generator-level drop-ins.  Inference only (eval() semantics --
    BatchNorm running statistics, dropout off); forward raises while the module is in training mode."""

    _state_prefix = ""              # prefix that maps this module's keys onto SelfAttention_G's names
    _block = None                   # stand-alone building blocks: (kind, in_channels, out_channels, leaky), set per instance

    def _native_state(self):
        sd = self.state_dict()
        skip = "num_batches_tracked"
        return {self._state_prefix + k: v for k, v in sd.items() if not k.endswith(skip)}

    def _tracked_tensors(self):
        """The parameter / buffer tensors behind the packed handle, listed once (walking state_dict() on every forward
        costs hundreds of microseconds on a 1.7 ms step).  The list is rebuilt whenever nn.Module machinery may have
        replaced tensor objects (_apply: .to()/.cuda()/.float(); load_state_dict with assign=True)."""
        cached = self.__dict__.get("_tracked")
        if cached is None:
            cached = self.__dict__["_tracked"] = list(self.state_dict(keep_vars=True).values())
        return cached

    def _fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in self._tracked_tensors())

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_tracked", None)
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_tracked", None)
        out = super().load_state_dict(*args, **kwargs)
        self.__dict__.pop("_tracked", None)
        return out

    def native(self, lane=0):
        """The packed device handle of `lane`, rebuilt whenever a parameter or buffer changed.  Every lane (the stream
        lanes of AudioToPosePipeline) owns a handle -- packed weights, activation arena, launch plans -- built from THIS
        module's parameters, so a later load_state_dict / in-place weight edit / set_output_denorm reaches all lanes."""
        _cabi.require_cuda(type(self).__name__)
        fp = self._fingerprint()
        handles = self.__dict__.setdefault("_handles", {})
        entry = handles.get(lane)
        if entry is None or entry[1] != fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("%s: parameters are on %s; move the module to a CUDA device (.cuda()) -- "
                                   "there is no CPU fallback" % (type(self).__name__, dev))
            entry = handles[lane] = (NativeHandle(self._native_state(), dev, self._block), fp)
        return entry[0]

    def repack(self):
        """Force re-folding / re-packing of the weights on the next forward (all lanes)."""
        self.__dict__.pop("_handles", None)
        self.__dict__.pop("_tracked", None)

    def check_device_status(self):
        """Synchronise and raise if a kernel's bounded barrier wait expired (any lane's handle)."""
        for h, _ in self.__dict__.get("_handles", {}).values():
            _cabi.check(_cabi.lib().a2m_model_status(h.ptr))

    def _require_eval(self):
        if self.training:
            raise RuntimeError("%s implements inference (eval-mode) semantics only: call .eval() first "
                               "(SURVEY.md D3; BatchNorm batch statistics and dropout are not implemented)"
                               % type(self).__name__)

    def __getstate__(self):
        d = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        d = dict(d)
        d.pop("_handles", None)
        d.pop("_tracked", None)
        return d


def as_input(x, device, shape_msg, keep_strides=False):
    """fp32 tensor on `device`; contiguous unless keep_strides and the tensor is a row-strided view
    (unit inner stride, rows ordered) that the native kernels can read in place."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dim() != 3:
        raise ValueError(shape_msg % (tuple(x.shape),))
    x = x.to(device=device, dtype=torch.float32)
    if keep_strides and x.stride(2) == 1 and x.stride(1) >= x.shape[2] and \
            (x.shape[0] == 1 or x.stride(0) >= x.stride(1)):
        return x
    return x.contiguous()
