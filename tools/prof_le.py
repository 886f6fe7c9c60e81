"""Short local-energy run for ncu captures (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ultrasoundaugmentation_b200 import ops, synth
B, H, W = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 512, 512)))
dev = torch.device("cuda", 0)
img, _ = synth.make_batch(B, H, W, device=dev)
for _ in range(3):
    e, s = ops.local_energy(img)
torch.cuda.synchronize()
print("ok", float(s.mean()))
