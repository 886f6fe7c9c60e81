"""Scratch timing of the three kernel groups on one GPU (not the bench; development aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import ops, synth

dev = torch.device("cuda", 0)
PEAK = 6549.1

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    which = sys.argv[1:] or ["k1", "k3", "cm"]
    if "k1" in which:
        for B, H, W in [(256, 256, 256), (1024, 512, 512)]:
            img, lab = synth.make_batch(B, H, W, device=dev)
            d = (torch.rand(B) * 0.2 - 0.05) * H; tgc = torch.rand(B, 8) * 0.8 + 0.6
            taps = ops.gaussian_taps(0.02 * W)
            ms = timeit(lambda: ops.fused_tgc_deform(img, lab, d, tgc, taps), reps=20)
            gb = 10 * B * H * W / 1e9
            print(f"K1 B={B} {H}x{W}: {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s ({gb/ms*1e3/PEAK:.2%})  {B/ms*1e3:.0f} img/s", flush=True)
    if "k3" in which:
        B, H, W = 1024, 512, 512
        img, lab = synth.make_batch(B, H, W, device=dev)
        cm = torch.rand(B, H, W, device=dev)
        a = torch.rand(B) + 0.5; rho = torch.rand(B) * 0.4 + 0.3; n = torch.randint(1, 3, (B,), dtype=torch.int32)
        taps = ops.gaussian_taps(2.0)
        ms = timeit(lambda: ops.snr_reverb(img, cm, lab, a, rho, n, taps), reps=20)
        gb = 13 * B * H * W / 1e9
        print(f"K3 B={B} {H}x{W}: {ms*1e3:.1f} us  {gb/ms*1e3:.0f} GB/s ({gb/ms*1e3/PEAK:.2%})", flush=True)
    if "cm" in which:
        for (B, H, W, rtol, fp64) in [(8, 512, 512, 1e-8, False), (64, 512, 512, 1e-8, False), (256, 512, 512, 1e-8, False),
                                     (512, 512, 512, 1e-8, False), (32, 1024, 1024, 1e-8, False)]:
            img, lab = synth.make_batch(B, H, W, device=dev)
            torch.cuda.synchronize(); t0 = time.time()
            cm, it, rr = ops.confidence_map(img, rtol=rtol, fp64_vectors=fp64, return_info=True)
            torch.cuda.synchronize(); t = time.time() - t0
            t0 = time.time()
            cm, it, rr = ops.confidence_map(img, rtol=rtol, fp64_vectors=fp64, return_info=True)
            torch.cuda.synchronize(); t = time.time() - t0
            itm = it.float().mean().item()
            eff = itm * 64 * (H - 2) * W * B / t / 1e9
            print(f"CM B={B} {H}x{W} rtol={rtol} fp64={fp64}: {t*1e3:.1f} ms  iters mean {itm:.0f} max {it.max().item()} "
                  f"relres max {rr.max().item():.2e}  {B/t:.1f} img/s  eff(64B model) {eff:.0f} GB/s ({eff/PEAK:.2%})  "
                  f"us/iter {t/it.max().item()*1e6:.1f}", flush=True)

main()
