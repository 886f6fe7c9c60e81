"""Development aid: confidence-map solver timing over batch sizes / coarse-grid knobs (no oracle; parity is in tests/).

    python tools/cm_sweep.py "64x512x512,512x512x512" [coarse=auto|fine|standard|off] [deformed]
Environment knobs of the library (development): USAUG_CM_COARSE_SHIFT, USAUG_CM_COARSE_BXSHIFT.
"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import ops, synth

dev = torch.device("cuda", 0)
PEAK = 6549.1
shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1] if len(sys.argv) > 1 else "64x512x512").split(",")]
coarse = {"auto": True, "off": False}.get(sys.argv[2], sys.argv[2]) if len(sys.argv) > 2 else True
deformed = len(sys.argv) > 3 and sys.argv[3] == "deformed"
debug = len(sys.argv) > 4 and sys.argv[4] == "debug"
for B, H, W in shapes:
    img, lab = synth.make_batch(B, H, W, device=dev)
    if deformed:
        chain = us.PhysicsChain.full(seed=2021)
        p = chain.get_params(B, image_size=(H, W))
        img, lab = ops.fused_tgc_deform(img, lab, p["d"], p["tgc"], chain.deform.taps(W))
    for rep in range(2):
        tm = []
        torch.cuda.synchronize(); t0 = time.time()
        cm, it, rr = ops.confidence_map(img, rtol=1e-8, return_info=True, coarse=coarse, timing=tm, debug_timing=debug and rep == 1)
        torch.cuda.synchronize(); t = time.time() - t0
    kms = sum(a.elapsed_time(b) for a, b, _, _ in tm)
    itm = it.float().mean().item()
    eff = itm * 64 * (H - 2) * W * B / (kms / 1e3) / 1e9
    print(f"CM B={B} {H}x{W} coarse={coarse} deformed={deformed}: wall {t*1e3:.1f} ms kernel {kms:.1f} ms  iters mean {itm:.0f} max {it.max().item()} "
          f"min {it.min().item()} relres max {rr.max().item():.2e}  {B/t:.1f} img/s  64B-model {eff:.0f} GB/s ({eff/PEAK:.2%})  "
          f"us/iter(max) {kms/it.max().item()*1e3:.1f}", flush=True)
