"""Event timeline of the two-lane pipeline (no nsys in this environment): every op of every forward of both lanes gets
an end-of-op event (a2m_model_timeline_*); prints, for the middle steps, which op classes were running when, how much of
the wall time had a GEMM / GNN op in flight, and the per-step critical path.
    python tools/timeline.py [steps] [lanes] > profiles/r1_timeline.txt
Event recording adds a few microseconds per op; compare shares, not absolutes."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 256
a2m = importlib.import_module("audio-to-motion-generation_b200")
cabi = importlib.import_module("audio-to-motion-generation_b200._cabi")
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
lib = a2m.load_library()
torch.manual_seed(0)
model = rmm.SelfAttention_G().cuda().eval()
pipe = pipeline.AudioToPosePipeline(model, lanes=lanes)
wav = [0.1 * torch.randn(B, 68267, device="cuda") for _ in range(3)]
gt = [50 * torch.randn(B, 64, 104, device="cuda") for _ in range(3)]
for i in range(2 * lanes + 2):
    pipe.step(wav[i % 3], gt[i % 3])
pipe.finish()
handles = [pipe.model.native(i) for i in range(pipe._n_lanes)]
per_lane = (steps + lanes - 1) // lanes
for h in handles:
    cabi.check(lib.a2m_model_timeline_begin(h.ptr, B, 64, 64, per_lane))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(per_lane * lanes):
    pipe.step(wav[i % 3], gt[i % 3])
pipe.sync_lanes()
e1.record()
torch.cuda.synchronize()
total_ms = e0.elapsed_time(e1)
ops = []          # (lane, step, op index, name, stream id, end ms)
names = None
for lane, h in enumerate(handles):
    n_steps, n_ops, unet_end, body_end = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    buf = (ctypes.c_float * (per_lane * 80))()
    cabi.check(lib.a2m_model_timeline_read(h.ptr, buf, len(buf), ctypes.byref(n_steps), ctypes.byref(n_ops),
                                           ctypes.byref(unet_end), ctypes.byref(body_end), B, 64, 64))
    n = n_ops.value
    if names is None:
        names = [lib.a2m_model_op_name(h.ptr, B, 64, 64, i, None).decode() for i in range(n)]
    for s in range(n_steps.value):
        row = [buf[s * (n + 1) + j] for j in range(n + 1)]
        ops.append((lane, s, -1, "start", 0, row[0]))
        for i in range(n):
            stream = 1 if unet_end.value <= i < body_end.value else 0
            ops.append((lane, s, i, names[i], stream, row[1 + i]))
t0 = min(o[5] for o in ops)
# per (lane, step, stream): durations = end - previous end on that stream (the body stream starts at the end of the UNet)
rows = []
for lane in range(lanes):
    for s in range(per_lane):
        mine = [o for o in ops if o[0] == lane and o[1] == s]
        start = [o for o in mine if o[2] == -1][0][5]
        prev = {0: start, 1: None}
        unet_end_t = None
        for o in sorted((o for o in mine if o[2] >= 0), key=lambda o: o[2]):
            st = o[4]
            if st == 1 and prev[1] is None:
                prev[1] = unet_end_t
            begin = prev[st]
            rows.append((lane, s, o[2], o[3], st, begin - t0, o[5] - t0))
            prev[st] = o[5]
            if o[3].startswith("unet"):
                unet_end_t = o[5]
print("# tools/timeline.py: %d steps over %d lanes, B = %d; wall %.3f ms per step (with event overhead)" % (
    per_lane * lanes, lanes, B, total_ms / (per_lane * lanes)))


def klass(name):
    if name.endswith(".gnn"):
        return "gnn"
    if name.endswith(".gemm"):
        return "gemm"
    return "other"


# occupancy of the wall time by class, sampled on a 1 us grid over the middle steps
lo = sorted(r[5] for r in rows if r[2] == 0)[lanes]              # skip the first step of each lane
hi = sorted(r[6] for r in rows)[-1 - 4 * lanes]
grid = int((hi - lo) * 1000)
busy = {"gemm": [0] * grid, "gnn": [0] * grid, "other": [0] * grid}
for r in rows:
    k = klass(r[3])
    a, b = int((r[5] - lo) * 1000), int((r[6] - lo) * 1000)
    for t in range(max(a, 0), min(b, grid)):
        busy[k][t] += 1
any_busy = sum(1 for t in range(grid) if busy["gemm"][t] or busy["gnn"][t] or busy["other"][t])
print("# window %.3f ms: some op in flight %.1f %%; a GEMM op in flight %.1f %%; a GNN op in flight %.1f %%; GEMM and GNN "
      "together %.1f %%; two or more GNN ops %.1f %%; only 'other' ops %.1f %%" % (
          hi - lo, 100 * any_busy / grid, 100 * sum(1 for t in range(grid) if busy["gemm"][t]) / grid,
          100 * sum(1 for t in range(grid) if busy["gnn"][t]) / grid,
          100 * sum(1 for t in range(grid) if busy["gemm"][t] and busy["gnn"][t]) / grid,
          100 * sum(1 for t in range(grid) if busy["gnn"][t] >= 2) / grid,
          100 * sum(1 for t in range(grid) if busy["other"][t] and not busy["gemm"][t] and not busy["gnn"][t]) / grid))
print("# (an op is 'in flight' from the end of its predecessor on the same stream to its own end: queueing behind other "
      "streams' kernels is included)")
mid = per_lane // 2
for lane in range(lanes):
    print("# lane %d, step %d: op, stream, begin us, end us, in-flight us" % (lane, mid))
    for r in rows:
        if r[0] == lane and r[1] == mid:
            print("%d %-28s s%d %9.1f %9.1f %8.1f" % (lane, r[3], r[4], 1e3 * r[5], 1e3 * r[6], 1e3 * (r[6] - r[5])))
