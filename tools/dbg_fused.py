import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import synth as osynth
from ultrasoundaugmentation_b200 import ops
dev = torch.device("cuda", 0)
for (B, H, W) in [(1, 64, 48), (1, 96, 48), (1, 64, 80), (1, 96, 80), (2, 96, 80), (1, 64, 128), (1, 64, 132), (1, 64, 256), (1, 40, 512), (1, 200, 64), (1,500,64)]:
    img, _ = osynth.make_batch(B, H, W)
    t = torch.from_numpy(img).to(dev)
    a, ita, rra = ops.confidence_map(t, rtol=1e-8, maxiter=3000, return_info=True)
    b, itb, rrb = ops.confidence_map(t, rtol=1e-8, maxiter=3000, return_info=True, legacy_kernel=True)
    torch.cuda.synchronize()
    print(B, H, W, "fused iters", ita.tolist(), "relres", [f"{v:.1e}" for v in rra.tolist()], "| legacy iters", itb.tolist(),
          "maxdiff", float((a - b).abs().max()), flush=True)
