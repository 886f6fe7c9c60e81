"""Development aid: which solver variant wins at which batch size (512x512 images, deformed like the chain's).

    python tools/cm_variants.py "96,128,192,256,384"
"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import ops, synth
dev = torch.device("cuda", 0)
H = W = 512
for B in (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "128,256").split(",")):
    img, lab = synth.make_batch(B, H, W, device=dev)
    chain = us.PhysicsChain.full(seed=2021)
    p = chain.get_params(B, image_size=(H, W))
    img, lab = ops.fused_tgc_deform(img, lab, p["d"], p["tgc"], chain.deform.taps(W))
    for name, kw in (("auto", {}), ("planes+twist+std", dict(up_variant="planes", twist=True, coarse="standard")),
                     ("planes+twist+fine", dict(up_variant="planes", twist=True, coarse="fine")),
                     ("recomp+twist+std", dict(up_variant="recompute", twist=True, coarse="standard")),
                     ("recomp+notwist+std", dict(up_variant="recompute", twist=False, coarse="standard")),
                     ("recomp+notwist+fine", dict(up_variant="recompute", twist=False, coarse="fine")),
                     ("recomp+twist+fine", dict(up_variant="recompute", twist=True, coarse="fine"))):
        for rep in range(2):
            tm = []
            cm, it, rr = ops.confidence_map(img, rtol=1e-8, return_info=True, timing=tm, **kw)
            torch.cuda.synchronize()
        kms = sum(a.elapsed_time(b) for a, b, _, _ in tm)
        print(f"B={B:4d} {name:22s} kernel {kms:7.1f} ms  iters mean {it.float().mean().item():5.0f} max {it.max().item():4d}  {B / kms * 1e3:7.1f} img/s", flush=True)
