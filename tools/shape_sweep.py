"""Development aid: confidence-map throughput for other image shapes (sanity sweep, not the bench)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ultrasoundaugmentation_b200 import ops, synth
dev = torch.device("cuda", 0)
for (B, H, W) in [(2048, 256, 256), (512, 384, 512), (128, 768, 768), (512, 512, 384), (256, 600, 800), (1024, 300, 400), (64, 1024, 512)]:
    img, _ = synth.make_batch(B, H, W, device=dev)
    ops.confidence_map(img[: max(B // 8, 1)], rtol=1e-8); torch.cuda.synchronize()
    t0 = time.time()
    cm, it, rr = ops.confidence_map(img, rtol=1e-8, return_info=True)
    torch.cuda.synchronize(); t = time.time() - t0
    itm = it.float().mean().item()
    eff = itm * 64 * (H - 2) * W * B / t / 1e9
    print(f"CM B={B} {H}x{W}: {t*1e3:.1f} ms iters mean {itm:.0f} max {it.max().item()} relres max {rr.max().item():.2e} "
          f"{B/t:.1f} img/s  {B*H*W/t/1e6:.0f} Mpx/s  eff(64B model) {eff:.0f} GB/s ({eff/6549.1:.2%})", flush=True)
