"""CUDA-event timing of the fused graph stack alone.   python tools/gnn_time.py [part] [n_graphs] [iters]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
part = sys.argv[1] if len(sys.argv) > 1 else "hand"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
a2m = importlib.import_module("audio-to-motion-generation_b200")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
torch.manual_seed(0)
m = rmm.SelfAttention_G().cuda().eval()
J = 10 if part == "body" else 42
x = torch.randn(n, J, 64, device="cuda")
for _ in range(3):
    y = m.graph_stack(part, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = m.graph_stack(part, x)
e1.record()
torch.cuda.synchronize()
m.check_device_status()
print("%s n=%d: %.1f us per call (includes the fp32<->bf16 conversions of graph_stack), mean|y| %.4f" % (
    part, n, 1e3 * e0.elapsed_time(e1) / iters, float(y.abs().mean())))
