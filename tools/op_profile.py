"""Per-launch CUDA-event profile of one SelfAttention_G forward (a2m_model_profile_ops).
    python tools/op_profile.py [B] [iters]
"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
a2m = importlib.import_module("audio-to-motion-generation_b200")
cabi = importlib.import_module("audio-to-motion-generation_b200._cabi")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
lib = a2m.load_library()
torch.manual_seed(0)
model = rmm.SelfAttention_G()
with torch.no_grad():
    for n, p in model.state_dict().items():
        if n.endswith("gamma"):
            p.fill_(0.5)
model = model.cuda().eval()
x = (-1.5 + 1.5 * torch.randn(B, 64, 64, device="cuda")).contiguous()
model(x)
torch.cuda.synchronize()
h = model.native()
cap = 256
ms = (ctypes.c_float * cap)()
n = ctypes.c_int()
dev = x.device
for _ in range(2):
    cabi.check(lib.a2m_model_profile_ops(h.ptr, cabi.ptr(x), x.stride(0), x.stride(1), B, 64, 64, ITERS, ms, cap,
                                         ctypes.byref(n), cabi.stream_ptr(dev)))
tot = sum(ms[i] for i in range(n.value))
print("B=%d  %d launches  %.3f ms per forward (sum of per-launch event times)" % (B, n.value, tot))
groups = {}
for i in range(n.value):
    fl = ctypes.c_int64()
    name = lib.a2m_model_op_name(h.ptr, B, 64, 64, i, ctypes.byref(fl)).decode()
    tf = fl.value / (ms[i] * 1e-3) / 1e12 if fl.value else 0.0
    print("%3d %-28s %8.1f us %5.1f%%  %9.1f MFLOP %7.1f TF/s" % (i, name, ms[i] * 1e3, 100 * ms[i] / tot, fl.value / 1e6, tf))
    g = groups.setdefault(name.split(".")[0] + (".gemm" if fl.value else ".other"), [0.0, 0])
    g[0] += ms[i]; g[1] += fl.value
for k, (t, f) in sorted(groups.items()):
    print("%-16s %8.1f us  %6.1f TF/s" % (k, t * 1e3, f / (t * 1e-3) / 1e12 if f else 0.0))
