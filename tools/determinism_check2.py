import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth, weights
from oracle.make_golden import model_input
a2m = importlib.import_module("audio-to-motion-generation_b200")
mods = a2m.install_dropin()
model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
model.load_state_dict(weights.make_state_dict(0, "stress"))
x = model_input(3, 8, 64, 64).cuda()
p8, _ = model(x)
torch.cuda.synchronize()
for sync in (True, False):
    outs = []
    for i in range(0, 8, 2):
        o, _ = model(x[i:i + 2])
        if sync:
            torch.cuda.synchronize()
        outs.append(o)
    p2 = torch.cat(outs)
    d = (p8 - p2).abs().amax(dim=(1, 2))
    print("sync" if sync else "nosync", "per-clip max|d|", [float("%.2e" % v) for v in d.tolist()])
    cols = (p8 - p2).abs().amax(dim=(0, 1))
    print("   body cols max", cols[:20].max().item(), " hand cols max", cols[20:].max().item())
p4a, _ = model(x[0:4]); p4b, _ = model(x[4:8])
d = (p8 - torch.cat([p4a, p4b])).abs().amax(dim=(1, 2))
print("B4 per-clip max|d|", [float("%.2e" % v) for v in d.tolist()])
p1 = torch.cat([model(x[i:i + 1])[0] for i in range(8)])
d = (p8 - p1).abs().amax(dim=(1, 2))
print("B1 per-clip max|d|", [float("%.2e" % v) for v in d.tolist()])
