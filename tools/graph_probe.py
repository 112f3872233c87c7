"""Probe: capture one device-resident step (mel -> generator -> eval) into a CUDA graph and time replays."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
a2m = importlib.import_module("audio-to-motion-generation_b200")
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
me = importlib.import_module("audio-to-motion-generation_b200.motion_evaluation")
torch.manual_seed(0)
model = rmm.SelfAttention_G().cuda().eval()
B = 256
wav = 0.1 * torch.randn(B, 68267, device="cuda")
gt = 50 * torch.randn(B, 64, 104, device="cuda")
pipe = pipeline.AudioToPosePipeline(model, lanes=1)
for _ in range(3):
    pipe.step(wav, gt)
pipe.finish(); pipe.reset()
torch.cuda.synchronize()
def timeit(fn, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def eager():
    pose = pipe.generate(wav)
    me.evaluate_poses(pose, gt, 0.2, accum=pipe.accum)
print("eager ms/step", timeit(eager))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    eager()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        eager()
    torch.cuda.synchronize()
    print("graph ms/step", timeit(g.replay))
    pipe.reset(); g.replay(); r1 = pipe.finish()
    pipe.reset(); eager(); r2 = pipe.finish()
    print("graph vs eager pck", r1["pck"], r2["pck"], r1["pck_hits"] == r2["pck_hits"])
except Exception as e:
    print("capture failed:", repr(e)[:500])
