"""Small shapes through every kernel, for compute-sanitizer (memcheck / racecheck) runs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, torch
import ultrasoundaugmentation_b200 as us
from oracle import synth as osynth
from ultrasoundaugmentation_b200 import ops

dev = torch.device("cuda", 0)
for (B, H, W) in [(3, 40, 136), (2, 37, 50), (1, 24, 260)]:
    img, lab = osynth.make_batch(B, H, W)
    ti, tl = torch.from_numpy(img).to(dev), torch.from_numpy(lab).to(dev)
    chain = us.PhysicsChain.full(seed=3, rtol=1e-6, maxiter=300)
    oi, ol = chain(ti, tl)
    cm = ops.confidence_map(ti, rtol=1e-6, maxiter=200, legacy_kernel=True)
    cm = ops.confidence_map(ti, rtol=1e-6, maxiter=200, fp64_vectors=True)
    box = ops.bone_box(tl)
    torch.cuda.synchronize()
    print(B, H, W, "ok", float(oi.mean()), chain.snr.last_info[0].tolist(), box[0].tolist(), flush=True)
