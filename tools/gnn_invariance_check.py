import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import weights
a2m = importlib.import_module("audio-to-motion-generation_b200")
mods = a2m.install_dropin()
model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
model.load_state_dict(weights.make_state_dict(0, "stress"))
torch.manual_seed(1)
for part, J, gpc in (("hand", 42, 3), ("body", 10, 12)):
    x = torch.randn(600, J, 64, device="cuda")
    full = model.graph_stack(part, x)           # 600 = multiple of 3 and 12: only full tiles
    for n in (1, 2, 3, 4, 5, 8, 11, 127, 128, 130, 599):
        got = model.graph_stack(part, x[:n])
        d = (got - full[:n]).abs().amax(dim=(1, 2))
        bad = torch.nonzero(d > 0).flatten().tolist()
        print(part, "n=%d" % n, "max|d| %.3e" % d.max().item(), "bad graphs", bad[:10])
