"""Phase clocks of the fused graph stack (probe build only: gnn_fused.cu compiled with -DA2M_GNN_TRACE, see
tools/probes/gnn_trace.sh).  Prints, for CTA 0's second tile, the clocks between the stamps of every layer.
    python tools/probes/gnn_trace.py [part] [n_graphs]"""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
part = sys.argv[1] if len(sys.argv) > 1 else "hand"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
a2m = importlib.import_module("audio-to-motion-generation_b200")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
lib = a2m.load_library()
torch.manual_seed(0)
m = rmm.SelfAttention_G().cuda().eval()
J = 10 if part == "body" else 42
x = torch.randn(n, J, 64, device="cuda")
for _ in range(3):
    y = m.graph_stack(part, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = m.graph_stack(part, x)
e1.record()
torch.cuda.synchronize()
print("graph_stack %s %d graphs: %.1f us per call (includes layout kernels)" % (part, n, e0.elapsed_time(e1) * 200))
buf = (ctypes.c_longlong * 128)()
assert lib.a2m_gnn_trace_read(buf) == 0
names = {0: "start", 1: "S mma done", 2: "S read+sync", 3: "P r1+sync", 4: "Z mma done", 5: "P r2/convert+sync",
         6: "Z2 mma done", 7: "convert+sync", 8: "OUT mma done", 9: "epilogue+sync"}
t00 = buf[0]
for layer in range(5):
    st = [(k, buf[layer * 12 + k]) for k in range(10) if buf[layer * 12 + k]]
    line = []
    for (k0, a), (k1, b) in zip(st, st[1:]):
        line.append("%s %d" % (names[k1], b - a))
    print("layer %d (%s) total %d clk: %s" % (layer, "GAT" if layer % 2 == 0 else "GC", st[-1][1] - st[0][1], " | ".join(line)))
print("tile total %d clk" % (buf[4 * 12 + 9] - t00))
print("per tile of CTA 0: wait for the node tile | layers | (clk since the first stamp at tile end)")
for it in range(20):
    a, b, c = buf[64 + it * 3], buf[64 + it * 3 + 1], buf[64 + it * 3 + 2]
    if c:
        print("  tile %2d: wait %6d  layers %6d  end at %8d" % (it, b - a, c - b, c - buf[64]))
