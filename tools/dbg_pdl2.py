"""Development aid: is the PDL race tied to host-to-device parameter copies that follow the PDL-launched kernel?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import synth
dev = torch.device("cuda", 0)
B, H, W = 32, 1024, 1024
img, lab = synth.make_batch(B, H, W, seed=5, device=dev)
chain = us.PhysicsChain.full(seed=9, rtol=1e-8)
p = chain.get_params(B, image_size=(H, W))
pg = {k: v.to(dev) for k, v in p.items()}
for name, prm in (("host params", p), ("device params", pg), ("host params", p)):
    ok = []
    for rep in range(3):
        a = chain(img, lab, prm); b = chain(img, lab, prm)
        torch.cuda.synchronize()
        ok.append(bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])))
    print(name, ok, flush=True)
