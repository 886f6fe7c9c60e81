"""Development aid: which part of the chain can be captured into a CUDA graph."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import ops, synth
dev = torch.device("cuda", 0)
B, H, W = 3, 96, 128
img, lab = synth.make_batch(B, H, W, device=dev)
d = torch.zeros(B, device=dev) + 3.0; tgc = torch.ones(B, 8, device=dev); taps = ops.gaussian_taps(2.0).to(dev)
a = torch.ones(B, device=dev) * 1.2; rho = torch.ones(B, device=dev) * 0.5; n = torch.ones(B, dtype=torch.int32, device=dev)
cm = torch.rand(B, H, W, device=dev)
def attempt(name, fn):
    try:
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "captured OK", flush=True)
    except Exception as e:
        print(name, "FAILED:", str(e).splitlines()[0], flush=True)
        torch.cuda.synchronize()
attempt("K1", lambda: ops.fused_tgc_deform(img, lab, d, tgc, taps))
attempt("K3", lambda: ops.snr_reverb(img, cm, lab, a, rho, n, taps))
attempt("CM generic (legacy kernel)", lambda: ops.confidence_map(img, rtol=1e-6, legacy_kernel=True))
attempt("CM fused", lambda: ops.confidence_map(img, rtol=1e-6))
st = ops.ParamStager(dev, depth=1)
attempt("ParamStager", lambda: st.stage({"d": torch.zeros(B), "n": torch.ones(B, dtype=torch.int32)}))
