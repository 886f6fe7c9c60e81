"""Short K1 / K3 runs for ncu captures (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ultrasoundaugmentation_b200 import ops, synth
B, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 256, 256)))
dev = torch.device("cuda", 0)
img, lab = synth.make_batch(B, H, W, device=dev)
d = ((torch.rand(B) * 0.2 - 0.05) * H).to(dev); tgc = (torch.rand(B, 8) * 0.8 + 0.6).to(dev)
taps = ops.gaussian_taps(0.02 * W).to(dev)
for _ in range(3):
    oi, ol = ops.fused_tgc_deform(img, lab, d, tgc, taps)
cm = torch.rand(B, H, W, device=dev)
a = (torch.rand(B) + 0.5).to(dev); rho = (torch.rand(B) * 0.4 + 0.3).to(dev); n = torch.randint(1, 3, (B,), dtype=torch.int32).to(dev)
for _ in range(3):
    o = ops.snr_reverb(oi, cm, ol, a, rho, n, ops.gaussian_taps(2.0).to(dev))
torch.cuda.synchronize()
print("ok")
