"""Development aid: max|C_gpu - C_oracle| as a function of the PCG rtol (feeds DESIGN.md's tolerance)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, torch
from oracle import confmap as ocm, synth as osynth
from ultrasoundaugmentation_b200 import ops

dev = torch.device("cuda", 0)
for (B, H, W) in [(3, 256, 256), (2, 512, 512)]:
    img, _ = osynth.make_batch(B, H, W)
    ref = ocm.confidence_map_batch(img)
    for rtol in (1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9):
        cm, it, rr = ops.confidence_map(torch.from_numpy(img).to(dev), rtol=rtol, return_info=True)
        torch.cuda.synchronize()
        err = np.abs(cm.cpu().numpy() - ref).max(axis=(1, 2))
        print(f"{H}x{W} rtol={rtol:.0e} iters={it.tolist()} relres={[f'{v:.1e}' for v in rr.tolist()]} "
              f"max|dC|={[f'{e:.2e}' for e in err]} K={[f'{e / rtol:.1e}' for e in err]}", flush=True)
