import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from oracle import synth as osynth
from ultrasoundaugmentation_b200 import ops
dev = torch.device("cuda", 0)
for (B, H, W) in [(1, 64, 128), (1, 64, 256), (2, 64, 256)]:
    img, _ = osynth.make_batch(B, H, W)
    t = torch.from_numpy(img).to(dev)
    for mi in (0, 1, 2, 3, 20):
        a, ita, rra = ops.confidence_map(t, rtol=1e-12, maxiter=mi, return_info=True)
        b, itb, rrb = ops.confidence_map(t, rtol=1e-12, maxiter=mi, return_info=True, legacy_kernel=True)
        torch.cuda.synchronize()
        d = (a - b).abs()
        bad = torch.nonzero(~torch.isfinite(a))
        print(B, H, W, "maxiter", mi, "iters", ita.tolist(), itb.tolist(), "relres", [f"{v:.3e}" for v in rra.tolist()], [f"{v:.3e}" for v in rrb.tolist()],
              "maxdiff", float(d.nan_to_num(9).max()), "first nonfinite", bad[:2].tolist(), "rows with diff>1e-5", torch.nonzero(d.nan_to_num(9).amax(dim=(0,2)) > 1e-5).flatten()[:6].tolist(), flush=True)
