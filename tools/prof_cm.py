"""Short fixed-iteration confidence-map solve for ncu captures (development aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ultrasoundaugmentation_b200 import ops, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)
img, _ = synth.make_batch(B, 512, 512, device=dev)
import os
cm, it, rr = ops.confidence_map(img, rtol=1e-14, maxiter=iters, return_info=True, debug_timing=bool(os.environ.get('PCG_DEBUG')))
torch.cuda.synchronize()
print("iters", it[:4].tolist(), "relres", rr[:2].tolist())
