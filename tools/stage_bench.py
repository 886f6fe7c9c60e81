"""Per-stage roofline numbers measured through the C ABI with
preallocated device buffers: CUDA events around back-to-back calls, L2 flushed between reps for the
small config.  Prints one JSON line per stage; results are copied into profiles/."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import ultrasoundaugmentation_b200 as us
from ultrasoundaugmentation_b200 import _ffi, ops, synth

dev = torch.device("cuda", 0)
PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0
lib = _ffi.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps, flush_l2):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    if flush_l2:
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def k1(B, H, W, flush_l2):
    img, lab = synth.make_batch(B, H, W, device=dev)
    d = ((torch.rand(B) * 0.2 - 0.05) * H).to(dev); tgc = (torch.rand(B, 8) * 0.8 + 0.6).to(dev)
    taps = ops.gaussian_taps(0.02 * W).to(dev); R = (taps.numel() - 1) // 2
    oi, ol = torch.empty_like(img), torch.empty_like(lab)
    ws = torch.empty(lib.usaug_warp_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fn():
        rc = lib.usaug_fused_tgc_deform(img.data_ptr(), lab.data_ptr(), oi.data_ptr(), ol.data_ptr(), B, H, W,
                                        d.data_ptr(), tgc.data_ptr(), 8, taps.data_ptr(), R, ws.data_ptr(), ws.numel(), st)
        assert rc == 0
    ms = timed(fn, 20, flush_l2)
    gb = 10 * B * H * W / 1e9
    return dict(stage="K1 fused TGC+deform+remap (scan+profile+warp)", B=B, H=H, W=W, us=ms * 1e3, bytes_per_px=10,
                achieved_GBps=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK, img_per_s=B / ms * 1e3, l2_flushed=flush_l2)


def k3(B, H, W):
    img, lab = synth.make_batch(B, H, W, device=dev)
    cm = torch.rand(B, H, W, device=dev)
    a = (torch.rand(B) + 0.5).to(dev); rho = (torch.rand(B) * 0.4 + 0.3).to(dev)
    n = torch.randint(1, 3, (B,), dtype=torch.int32).to(dev)
    taps = ops.gaussian_taps(2.0).to(dev); R = (taps.numel() - 1) // 2
    out = torch.empty_like(img)
    ws = torch.empty(lib.usaug_snr_reverb_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fn():
        rc = lib.usaug_snr_reverb(img.data_ptr(), cm.data_ptr(), lab.data_ptr(), out.data_ptr(), B, H, W, a.data_ptr(),
                                  rho.data_ptr(), n.data_ptr(), taps.data_ptr(), R, ws.data_ptr(), ws.numel(), st)
        assert rc == 0
    ms = timed(fn, 20, False)
    gb = 13 * B * H * W / 1e9
    return dict(stage="K3 SNR+reverb (scan+box+fused)", B=B, H=H, W=W, us=ms * 1e3, bytes_per_px=13,
                achieved_GBps=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK, img_per_s=B / ms * 1e3)


def le(B, H, W):
    img, _ = synth.make_batch(B, H, W, device=dev)
    energy = torch.empty_like(img); scale = torch.empty(B, device=dev)
    ws = torch.empty(lib.usaug_local_energy_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fn():
        rc = lib.usaug_local_energy(img.data_ptr(), energy.data_ptr(), scale.data_ptr(), B, H, W, 8.0, 2.1, 3, 0.55,
                                    ws.data_ptr(), ws.numel(), st)
        assert rc == 0
    ms = timed(fn, 10, False)
    gb = 56 * B * H * W / 1e9
    return dict(stage="local energy (rows fwd + cols + rows inv + scale)", B=B, H=H, W=W, us=ms * 1e3, bytes_per_px=56,
                achieved_GBps=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK, img_per_s=B / ms * 1e3)


def sc(B, H, W):
    img, lab = synth.make_batch(B, H, W, device=dev)
    g = us.ConvexGeometry.default(H, W, theta_max=0.4)
    ax, ay, thm, rmin, dr, inv_dr, inv_dth = (float(v) for v in g.scalars())
    s_, c_ = g.tables()
    sin_t, cos_t = torch.from_numpy(s_).to(dev), torch.from_numpy(c_).to(dev)
    pi, pl = torch.empty_like(img), torch.empty_like(lab)
    ci, cl = torch.empty_like(img), torch.empty_like(lab)
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for direction, (a, b, c, d_) in ((0, (img, lab, pi, pl)), (1, (pi, pl, ci, cl))):
        def fn():
            rc = lib.usaug_scan_convert(a.data_ptr(), b.data_ptr(), c.data_ptr(), d_.data_ptr(), B, H, W, H, W, direction,
                                        ax, ay, thm, rmin, dr, inv_dr, inv_dth, sin_t.data_ptr(), cos_t.data_ptr(), st)
            assert rc == 0
        ms = timed(fn, 20, False)
        gb = 10 * B * H * W / 1e9
        out.append(dict(stage="scan conversion " + ("Cartesian -> polar" if direction == 0 else "polar -> Cartesian"), B=B, H=H, W=W,
                        us=ms * 1e3, bytes_per_px=10, achieved_GBps=gb / ms * 1e3, frac=gb / ms * 1e3 / PEAK, img_per_s=B / ms * 1e3))
    return out


def cm(B, H, W, rtol=1e-8):
    img, _ = synth.make_batch(B, H, W, device=dev)
    ops.confidence_map(img, rtol=rtol)
    torch.cuda.synchronize()
    tm = []
    t0 = time.perf_counter()
    c, it, rr = ops.confidence_map(img, rtol=rtol, return_info=True, timing=tm)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kms = sum(a.elapsed_time(b) for a, b, _, _ in tm)
    alg = float(it.double().sum()) * 64 * (H - 2) * W
    return dict(stage="K2 confidence-map PCG", B=B, H=H, W=W, rtol=rtol, ms=wall * 1e3, pcg_kernel_ms=kms,
                iters_mean=float(it.float().mean()), iters_max=int(it.max()), relres_max=float(rr.max()),
                achieved_GBps=alg / kms / 1e6, frac=alg / kms / 1e6 / PEAK, img_per_s=B / wall,
                model="64 B/unknown/iteration")


def chain(B, H, W, rtol=1e-8):
    img, lab = synth.make_batch(B, H, W, device=dev)
    ch = us.PhysicsChain.full(seed=1, rtol=rtol)
    p = {k: v.to(dev) for k, v in ch.get_params(B, image_size=(H, W)).items()}
    ch(img, lab, p); torch.cuda.synchronize()
    t0 = time.perf_counter(); ch(img, lab, p); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    it = ch.snr.last_info[0]
    return dict(stage="full chain", B=B, H=H, W=W, rtol=rtol, ms=wall * 1e3, img_per_s=B / wall,
                iters_mean=float(it.float().mean()), iters_max=int(it.max()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["k1", "k3", "le", "sc", "cm", "chain"]
    out = []
    if "k1" in which:
        out += [k1(256, 256, 256, True), k1(256, 256, 256, False), k1(1024, 512, 512, False)]
    if "k3" in which:
        out += [k3(1024, 512, 512)]
    if "le" in which:
        out += [le(512, 512, 512), le(64, 1024, 1024), le(256, 256, 256)]
    if "sc" in which:
        out += sc(1024, 512, 512)
    if "cm" in which:
        out += [cm(64, 512, 512)]
    if "chain" in which:
        out += [chain(32, 1024, 1024)]
    for o in out:
        print(json.dumps(o), flush=True)
