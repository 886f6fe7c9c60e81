"""Development aid: K1 stage split -- bone-box (= surface scan + tiny reduce) timed alone vs the whole K1 call."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from ultrasoundaugmentation_b200 import _ffi, ops, synth
dev = torch.device("cuda", 0)
lib = _ffi.load()
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for B, H, W in [(256, 256, 256), (1024, 512, 512)]:
    img, lab = synth.make_batch(B, H, W, device=dev)
    d = ((torch.rand(B) * 0.2 - 0.05) * H).to(dev); tgc = (torch.rand(B, 8) * 0.8 + 0.6).to(dev)
    taps = ops.gaussian_taps(0.02 * W).to(dev); R = (taps.numel() - 1) // 2
    oi, ol = torch.empty_like(img), torch.empty_like(lab)
    ws = torch.empty(lib.usaug_warp_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    box = torch.empty(B, 4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    t_all = timed(lambda: lib.usaug_fused_tgc_deform(img.data_ptr(), lab.data_ptr(), oi.data_ptr(), ol.data_ptr(), B, H, W,
                                                     d.data_ptr(), tgc.data_ptr(), 8, taps.data_ptr(), R, ws.data_ptr(), ws.numel(), st))
    t_box = timed(lambda: lib.usaug_bone_box(lab.data_ptr(), B, H, W, box.data_ptr(), ws.data_ptr(), ws.numel(), st))
    t_cp = timed(lambda: (oi.copy_(img), ol.copy_(lab)))
    print(f"B={B} {H}x{W}: K1 {t_all:.1f} us | scan+box {t_box:.1f} us | plain copy of image+label (10 B/px) {t_cp:.1f} us", flush=True)
