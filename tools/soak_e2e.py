"""Soak of the end-to-end path (pinned host batches -> copy stream -> two lanes -> metrics): N steps, both input formats.
    python tools/soak_e2e.py [steps]"""
import faulthandler, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
readback = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
torch.manual_seed(0)
model = rmm.SelfAttention_G().cuda().eval()
pipe = pipeline.AudioToPosePipeline(model, lanes=2)
B = 256
wav = [(0.1 * torch.randn(B, 68267)).pin_memory() for _ in range(4)]
pcm = [(w * 65534).round().clamp(-32768, 32767).to(torch.int16).pin_memory() for w in wav]
gt = [(50 * torch.randn(B, 64, 104)).pin_memory() for _ in range(4)]
order = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "int16"]
for name in order:
    src = wav if name == "fp32" else pcm
    faulthandler.dump_traceback_later(45, exit=True)      # a stalled phase shows where the host is blocked
    pipe.reset()
    t0 = time.perf_counter()
    n = pipe.run_host_batches(((src[i % 4], gt[i % 4]) for i in range(steps)), per_step_readback=readback)
    res = pipe.finish()
    dt = time.perf_counter() - t0
    faulthandler.cancel_dump_traceback_later()
    print("readback %s " % readback, end="")
    print("%s: %d clips in %.2f s = %.0f clips/s, pck %.4f" % (name, n, dt, n / dt, res["pck"]), flush=True)
