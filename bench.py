#!/usr/bin/env python
"""Headline benchmark: augmented images/sec at 512x512 (BASELINE.json `metric`).

Workload = BASELINE.json configs[4], the configuration the metric is quoted on: the FULL chain
(probe-pressure deformation + TGC + remap -> confidence map PCG -> SNR -> reverb) with per-sample
random parameters on synthetic 512x512 B-mode images, sharded over N GPUs with no collective on
the data path.  Every GPU owns 512 images per step (weak scaling), so N = 8 is exactly configs[4]'s
batch of 4096 and N = 1 is its per-GPU share.

    python bench.py --gpus N --steps K --warmup W            # N>1: launched by torch.distributed.run
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --gpus N --strong                        # the literal configs[4] batch: 4096 images over N GPUs

Besides the headline the JSON line carries (N = 1 only, outside the headline's timed region): `stages` -- the
other BASELINE.json configs (configs[1] fused K1, K3, configs[2] confidence-map batch 64, configs[3] full chain
32 x 1024^2) timed with CUDA events and set against the HBM roofline --, `config0_cpu` (configs[0], one core),
`cpu_baseline` (oracle, direct solve, all cores), `cpu_baseline_cg` (SciPy cg + Jacobi to the same rtol) and
`parity_check` (two images of the timed batch against the oracle).

One "step" = one pass of the chain over the whole (sharded) batch.  Timed with CUDA events,
barrier + synchronize on both sides, max over ranks.  Rank 0 prints ONE JSON line.
The reference mount holds no code (README.md only), so the reference arm times the self-written
CPU oracle (`oracle/`, kind "port") on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "augmented images/sec at 512x512 (full chain)"
UNIT = "images/s"
H = W = 512
PER_GPU_BATCH = 512             # configs[4]: 4096 images over 8 GPUs
WAVE = 512                      # images per PCG launch (workspace ~7.7 GiB of the 180 GB)
PCG_BYTES_PER_UNKNOWN_ITER = 64  # SURVEY.md 8(d) algorithmic-bytes model


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step")
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--strong", action="store_true", help="strong scaling: configs[4]'s 4096 images split over the GPUs")
    ap.add_argument("--no-stages", action="store_true", help="skip the per-config stage timings (N = 1 only)")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-cpu-cg", action="store_true", help="skip the SciPy cg + Jacobi CPU baseline (~80 s on one core)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (one image per task, BLAS threads pinned to 1)
# ------------------------------------------------------------------------------------------------
def _cpu_warm(_):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    import numpy as np
    from oracle import chain as ochain, deform as odeform, synth as osynth
    img, lab = osynth.make_bmode(24, 24, 1)
    ochain.full_chain(img, lab, 1.0, np.ones(2, np.float32), 1.1, 0.5, 1, odeform.gaussian_taps(1.0),
                      odeform.gaussian_taps(1.0))
    return os.getpid()


def _cpu_one(args):
    idx, p = args
    from oracle import chain as ochain, deform as odeform, synth as osynth
    img, lab = osynth.make_bmode(H, W, 1234 + idx)
    t0 = time.perf_counter()
    ochain.full_chain(img, lab, p["d"], p["tgc"], p["a_snr"], p["rho"], p["n_ghost"],
                      odeform.gaussian_taps(0.02 * W), odeform.gaussian_taps(2.0))
    return time.perf_counter() - t0


def cpu_oracle_rate(n_images: int, cores: int, first: int = 0):
    """images/s of the oracle full chain (`oracle/`, the checker) on `cores` host processes."""
    import multiprocessing as mp

    import ultrasoundaugmentation_b200 as us
    ch = us.PhysicsChain.full(seed=2021)
    idx = [first + i for i in range(n_images)]
    prm = {k: v.numpy() for k, v in ch.get_params(n_images, sample_index=idx, image_size=(H, W)).items()}
    tasks = [(idx[i], {k: v[i] for k, v in prm.items()}) for i in range(n_images)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_warm, range(cores), chunksize=1)     # imports + first-touch outside the timed region
        t0 = time.perf_counter()
        per = pool.map(_cpu_one, tasks, chunksize=1)
        wall = time.perf_counter() - t0
    return n_images / wall, wall, sum(per) / len(per)


def _cpu_chain_ref(args):
    """oracle full chain on one explicit image (parity spot-check of the timed batch)"""
    img, lab, p = args
    from oracle import chain as ochain, deform as odeform
    return ochain.full_chain(img, lab, p["d"], p["tgc"], p["a_snr"], p["rho"], p["n_ghost"],
                             odeform.gaussian_taps(0.02 * img.shape[1]), odeform.gaussian_taps(2.0))


def _cpu_cg_one(conn, rtol):
    """SciPy cg + Jacobi on the confidence-map system of one synthetic 512x512 image, to the GPU run's rtol."""
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    import numpy as np
    import scipy.sparse.linalg as spla
    from oracle import confmap as ocm, synth as osynth
    img, _ = osynth.make_bmode(H, W, 1234)
    t0 = time.perf_counter()
    A, b = ocm.system(*ocm.edge_weights(img))
    A = A.tocsr()
    d = A.diagonal()
    n = [0]

    def cb(_x):
        n[0] += 1
    x0 = np.repeat(1.0 - np.arange(1, H - 1) / (H - 1.0), W)        # the GPU solver's initial guess
    x, info = spla.cg(A, b, x0=x0, rtol=rtol, atol=0.0, maxiter=200000,
                      M=spla.LinearOperator(A.shape, matvec=lambda r: r / d), callback=cb)
    sec = time.perf_counter() - t0
    conn.send({"seconds": sec, "iters": n[0], "info": int(info),
               "relres": float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))})
    conn.close()


def config0_cpu():
    """configs[0]: TGC + probe-pressure deformation on ONE synthetic 256x256 image, oracle, one core, wall time."""
    import numpy as np
    from oracle import deform as odeform, synth as osynth
    img, lab = osynth.make_bmode(256, 256, 1234)
    gains = np.linspace(0.7, 1.3, 8).astype(np.float32)
    taps = odeform.gaussian_taps(0.02 * 256)
    odeform.fused_tgc_deform(img, lab, np.float32(12.0), gains, taps)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        odeform.fused_tgc_deform(img, lab, np.float32(12.0), gains, taps)
    ms = (time.perf_counter() - t0) / reps * 1e3
    return {"config": "configs[0]: CPU NumPy reference, TGC + deformation, one 256x256 image", "ms": ms, "cores": 1,
            "kind": "port", "images_per_s": 1e3 / ms}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = a.cpu_sample or cores
    rates = []
    for s in range(a.warmup + a.steps):
        # warm-up steps import numpy/scipy and page in the workers; keep them short
        r, wall, per = cpu_oracle_rate(n if s >= a.warmup else min(n, cores), cores, first=s * n)
        if s >= a.warmup:
            rates.append((r, wall))
    total_wall = sum(w for _, w in rates)
    value = a.steps * n / total_wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_wall / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[4]: full chain 512x512 (bounded sample per step)",
                   "sample_images_per_step": n, "solver": "scipy.sparse.linalg.splu (direct, fp64)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} images/step of the same synthetic 512x512 workload, oracle full chain, "
                                   f"multiprocessing.Pool({cores}), 1 BLAS thread each"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# per-config stage timings (BASELINE.json configs[1]..[3] + K3), CUDA events, through the C ABI
# ------------------------------------------------------------------------------------------------
def run_stages(a, dev, peak):
    import torch

    import ultrasoundaugmentation_b200 as us
    from ultrasoundaugmentation_b200 import _ffi, ops, synth
    lib = _ffi.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)

    def timed(fn, reps, flush_l2):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        if flush_l2:
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
                tot += e0.elapsed_time(e1)
            return tot / reps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def entry(config, kernel, ms, alg_bytes, n_img, **extra):
        gbs = alg_bytes / (ms / 1e3) / 1e9
        d = {"config": config, "kernel": kernel, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_GBps": gbs,
             "frac": gbs / peak, "images_per_s": n_img / (ms / 1e3)}
        d.update(extra)
        return d

    out = []
    # configs[1]: fused TGC + deformation + remap, batch 256 of 256x256 (10 B/px), L2 flushed between repetitions
    B, Hs, Ws = 256, 256, 256
    img, lab = synth.make_batch(B, Hs, Ws, device=dev)
    g = torch.Generator().manual_seed(1)
    d = ((torch.rand(B, generator=g) * 0.2 - 0.05) * Hs).to(dev)
    tgc = (torch.rand(B, 8, generator=g) * 0.8 + 0.6).to(dev)
    taps = ops.gaussian_taps(0.02 * Ws).to(dev)
    R = (taps.numel() - 1) // 2
    oi, ol = torch.empty_like(img), torch.empty_like(lab)
    ws = torch.empty(lib.usaug_warp_workspace_bytes(B, Hs, Ws), dtype=torch.uint8, device=dev)

    def k1():
        _ffi.check(lib.usaug_fused_tgc_deform(img.data_ptr(), lab.data_ptr(), oi.data_ptr(), ol.data_ptr(), B, Hs, Ws, d.data_ptr(),
                                              tgc.data_ptr(), 8, taps.data_ptr(), R, ws.data_ptr(), ws.numel(), st), "K1")
    out.append(entry("configs[1]: fused TGC+deformation+remap, batch 256 of 256x256", "surface_scan + profile + fused_warp_kernel",
                     timed(k1, 20, True), 10.0 * B * Hs * Ws, B, reps=20, l2="flushed between repetitions", bytes_model="10 B/px"))
    del img, lab, oi, ol
    # K3 at a size that is HBM-bound by itself: SNR + reverb, batch 1024 of 512x512 (13 B/px)
    B, Hs, Ws = 1024, 512, 512
    img, lab = synth.make_batch(B, Hs, Ws, device=dev)
    cm = torch.rand(B, Hs, Ws, device=dev)
    av = (torch.rand(B, generator=g) + 0.5).to(dev)
    rho = (torch.rand(B, generator=g) * 0.4 + 0.3).to(dev)
    ng = torch.randint(1, 3, (B,), generator=g, dtype=torch.int32).to(dev)
    ftaps = ops.gaussian_taps(2.0).to(dev)
    Rf = (ftaps.numel() - 1) // 2
    outb = torch.empty_like(img)
    ws3 = torch.empty(lib.usaug_snr_reverb_workspace_bytes(B, Hs, Ws), dtype=torch.uint8, device=dev)
    # ... and K1 at the same size (at configs[1]'s 168 MB the three launches' fixed costs dominate)
    dL = ((torch.rand(B, generator=g) * 0.2 - 0.05) * Hs).to(dev)
    tgcL = (torch.rand(B, 8, generator=g) * 0.8 + 0.6).to(dev)
    tapsL = ops.gaussian_taps(0.02 * Ws).to(dev)
    RL = (tapsL.numel() - 1) // 2
    olL = torch.empty_like(lab)
    wsL = torch.empty(lib.usaug_warp_workspace_bytes(B, Hs, Ws), dtype=torch.uint8, device=dev)

    def k1_large():
        _ffi.check(lib.usaug_fused_tgc_deform(img.data_ptr(), lab.data_ptr(), outb.data_ptr(), olL.data_ptr(), B, Hs, Ws, dL.data_ptr(),
                                              tgcL.data_ptr(), 8, tapsL.data_ptr(), RL, wsL.data_ptr(), wsL.numel(), st), "K1")
    out.append(entry("K1: fused TGC+deformation+remap, batch 1024 of 512x512", "surface_scan + profile + fused_warp_kernel",
                     timed(k1_large, 20, False), 10.0 * B * Hs * Ws, B, reps=20, l2="working set 2.7 GB >> L2", bytes_model="10 B/px"))
    del olL, wsL

    def k3():
        _ffi.check(lib.usaug_snr_reverb(img.data_ptr(), cm.data_ptr(), lab.data_ptr(), outb.data_ptr(), B, Hs, Ws, av.data_ptr(),
                                        rho.data_ptr(), ng.data_ptr(), ftaps.data_ptr(), Rf, ws3.data_ptr(), ws3.numel(), st), "K3")
    out.append(entry("K3: SNR + reverb, batch 1024 of 512x512", "surface_scan + bone_box + snr_reverb_kernel",
                     timed(k3, 20, False), 13.0 * B * Hs * Ws, B, reps=20, l2="working set 3.5 GB >> L2", bytes_model="13 B/px"))
    # SURVEY 8(f) rows at the same size: scan conversion both ways (10 B/px) ...
    geo = us.ConvexGeometry.default(Hs, Ws, theta_max=0.4)
    gax, gay, gth, grmin, gdr, ginv_dr, ginv_dth = (float(v) for v in geo.scalars())
    s_, c_ = geo.tables()
    sin_t, cos_t = torch.from_numpy(s_).to(dev), torch.from_numpy(c_).to(dev)
    pl, cl = torch.empty_like(lab), torch.empty_like(lab)
    for direction, name, (si, sl, di, dl) in ((0, "Cartesian -> polar", (img, lab, outb, pl)), (1, "polar -> Cartesian", (outb, pl, cm, cl))):
        def scv():
            _ffi.check(lib.usaug_scan_convert(si.data_ptr(), sl.data_ptr(), di.data_ptr(), dl.data_ptr(), B, Hs, Ws, Hs, Ws, direction,
                                              gax, gay, gth, grmin, gdr, ginv_dr, ginv_dth, sin_t.data_ptr(), cos_t.data_ptr(), st),
                       "scan conversion")
        out.append(entry(f"8(f): scan conversion {name}, batch 1024 of 512x512", "scan_convert_kernel", timed(scv, 20, False),
                         10.0 * B * Hs * Ws, B, reps=20, l2="working set 2.7 GB >> L2", bytes_model="10 B/px"))
    del pl, cl, cm, outb, lab
    # ... and local energy (monogenic signal, separable DFTs: 56 B/px), batch 512 of 512x512
    B = 512
    img = img[:B].contiguous()
    energy, scale = torch.empty_like(img), torch.empty(B, device=dev)
    wsE = torch.empty(lib.usaug_local_energy_workspace_bytes(B, Hs, Ws), dtype=torch.uint8, device=dev)

    def le():
        _ffi.check(lib.usaug_local_energy(img.data_ptr(), energy.data_ptr(), scale.data_ptr(), B, Hs, Ws, 8.0, 2.1, 3, 0.55,
                                          wsE.data_ptr(), wsE.numel(), st), "local energy")
    out.append(entry("8(f): local energy, batch 512 of 512x512", "le_rows_fwd + le_cols + le_rows_inv (Stockham radix-8 FFT)", timed(le, 10, False),
                     56.0 * B * Hs * Ws, B, reps=10, l2="working set >> L2", bytes_model="56 B/px"))
    del img, energy, wsE
    torch.cuda.empty_cache()
    # configs[2]: confidence-map PCG, batch 64 of 512x512 (64 B x unknowns x iterations)
    B, Hs, Ws = 64, 512, 512
    img, _ = synth.make_batch(B, Hs, Ws, device=dev)
    info = {}

    def cmf():
        info["it"] = ops.confidence_map(img, rtol=a.rtol, return_info=True)[1]
    ms = timed(cmf, 20, False)
    it = info["it"].double()
    out.append(entry("configs[2]: confidence-map PCG, batch 64 of 512x512", "setup kernels + pcg_fused_kernel", ms,
                     float(it.sum()) * 64.0 * (Hs - 2) * Ws + 24.0 * B * Hs * Ws, B, reps=20, rtol=a.rtol,
                     iters_mean=float(it.mean()), iters_max=float(it.max()), l2="PCG working set 1.1 GB >> L2",
                     bytes_model="64 B x unknowns x per-image iterations + 24 B/px setup"))
    del img
    # configs[3]: full chain, per-sample random parameters, batch 32 of 1024x1024
    B, Hs, Ws = 32, 1024, 1024
    img, lab = synth.make_batch(B, Hs, Ws, device=dev)
    chain = us.PhysicsChain.full(seed=2021, rtol=a.rtol)
    prm = chain.get_params(B, image_size=(Hs, Ws))

    def full():
        chain(img, lab, prm)
    ms = timed(full, 20, False)
    it = chain.snr.last_info[0].double()
    out.append(entry("configs[3]: full chain, per-sample parameters, batch 32 of 1024x1024", "K1 + confidence-map PCG + K3", ms,
                     float(it.sum()) * 64.0 * (Hs - 2) * Ws + (10.0 + 24.0 + 13.0) * B * Hs * Ws, B, reps=20, rtol=a.rtol,
                     iters_mean=float(it.mean()), iters_max=float(it.max()), l2="PCG working set 2.3 GB >> L2",
                     bytes_model="K1 10 B/px + PCG (64 B x unknowns x iterations + 24 B/px setup) + K3 13 B/px"))
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import ultrasoundaugmentation_b200 as us
    from ultrasoundaugmentation_b200 import _ffi, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and a.gpus > 1:
        # convenience: re-launch under torch.distributed.run exactly as the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (there is no CPU fallback); use --impl reference for the CPU arm")
    _ffi.load()                                       # fail loudly if libusaug.so is missing
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # SciPy cg + Jacobi CPU baseline: ~80 s on one core, so it runs in its own process NEXT TO the GPU phases
    # (one busy host core out of >= 16 does not touch device-timed numbers) and is collected at the end
    cg_proc = cg_conn = None
    if rank == 0 and world == 1 and not (a.no_cpu_baseline or a.no_cpu_cg):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        cg_conn, child = ctx.Pipe(duplex=False)
        cg_proc = ctx.Process(target=_cpu_cg_one, args=(child, a.rtol), daemon=True)
        cg_proc.start()
        child.close()

    if a.strong:
        a.batch = 4096                            # configs[4] literally: 4096 images split over the GPUs
        a.per_gpu_batch = (a.batch + world - 1) // world
    else:
        a.batch = a.per_gpu_batch * world         # total images per step over all GPUs
    lo, hi = sharding.shard_bounds(a.batch, rank, world)
    nloc = hi - lo
    img, lab = synth.make_batch(nloc, H, W, seed=1234 + rank, device=dev)
    chain = us.PhysicsChain.full(seed=2021, rtol=a.rtol, max_wave=WAVE)
    idx = torch.arange(lo, hi).numpy()

    def params_for(step):
        return chain.get_params(nloc, sample_index=idx + step * a.batch, image_size=(H, W))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------------
    for s in range(a.warmup):
        chain(img, lab, params_for(s))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    chain.snr.timing = []
    iters_log = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out_last = params_last = None
    for s in range(a.steps):
        params_last = params_for(a.warmup + s)
        out_last = chain(img, lab, params_last)
        iters_log.append(chain.snr.last_info[0])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    timing, chain.snr.timing = chain.snr.timing, None
    value = a.batch * a.steps / (ms / 1e3)
    # two images of the last timed step, copied out now for the parity spot-check further down
    pick = [0, nloc - 1]
    spot = [(img[j].cpu().numpy(), lab[j].cpu().numpy(), {k: v[j].numpy() for k, v in params_last.items()},
             out_last[0][j].cpu().numpy(), out_last[1][j].cpu().numpy()) for j in pick]
    out_last = None

    # ---- roofline of the dominant kernel (pcg_persistent_kernel), measured live -------------------
    iters_all = torch.cat(iters_log).to(torch.float64)          # [steps * nloc]
    unknowns = (H - 2) * W
    launches = len(timing)
    pcg_ms = sum(t0.elapsed_time(t1) for t0, t1, _, _ in timing)
    alg_bytes = float(iters_all.sum().item()) * PCG_BYTES_PER_UNKNOWN_ITER * unknowns
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (pcg_ms / 1e3) / 1e9 if pcg_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:   # ncu dram__bytes per unknown per iteration (one --set full capture), scaled to this run's launches
        tj = json.loads((ROOT / "profiles" / "pcg_traffic.json").read_text())
        traffic = tj["dram_bytes_per_unknown_iteration"] * float(iters_all.sum().item()) * unknowns / max(launches, 1)
        traffic_src = tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": "pcg_fused_kernel (persistent cooperative PCG)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "launches": launches, "avg_launch_ms": pcg_ms / max(launches, 1),
                "algorithmic_bytes_per_launch": alg_bytes / max(launches, 1),
                "model": "64 B/unknown/iteration x 261120 unknowns x sum of per-image iterations (SURVEY.md 8(d))",
                "share_of_step": pcg_ms / (e0.elapsed_time(e1))}

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    # PhysicsChain.augment_host_stream: every step copies its inputs from pinned host memory and its results back;
    # the copies of neighbouring steps overlap the kernels (three streams), as a dataloader would run it.
    h_img = img.cpu().pin_memory()
    h_lab = lab.cpu().pin_memory()

    def host_run(chain_, hi_, steps, first_step):
        feed = ((hi_, h_lab, idx + (first_step + s) * a.batch) for s in range(steps))
        n = 0
        for oi, ol in chain_.augment_host_stream(feed, device=dev, depth=2):
            n += oi.shape[0]
        return n

    def host_rate(chain_, hi_, px_bytes_in, px_bytes_out):
        host_run(chain_, hi_, 4, 900)                       # warm-up: all depth + 2 = 4 pinned output pairs of the pipeline's ring
                                                            # (3.2 GB of cudaHostAlloc; with one warm-up step three of them fell into the timed region)
        barrier()
        t0 = time.perf_counter()
        host_run(chain_, hi_, a.e2e_steps, 1000)
        torch.cuda.synchronize(dev)
        wall = max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3
        return {"value": a.batch * a.e2e_steps / wall, "unit": UNIT, "h2d_bytes_per_step": nloc * H * W * px_bytes_in * world,
                "d2h_bytes_per_step": nloc * H * W * px_bytes_out * world, "steps": a.e2e_steps}

    e2e = e2e_u8 = None
    if a.e2e_steps > 0:
        e2e = host_rate(chain, h_img, 5, 5)
        e2e["note"] = ("PhysicsChain.augment_host_stream on pinned host tensors (f32 image + u8 label in and out); "
                       "H2D + D2H of every step inside the timed region, overlapped with the kernels of the neighbouring steps")
        # the same chain on 8-bit B-mode input and output: 2 B/px over PCIe each way
        chain8 = us.PhysicsChain.full(seed=2021, rtol=a.rtol, max_wave=WAVE, output_dtype=torch.uint8)
        h_img8 = (h_img * 255.0 + 0.5).floor().to(torch.uint8).pin_memory()
        e2e_u8 = host_rate(chain8, h_img8, 2, 2)
        e2e_u8["pcg_iters_mean"] = float(chain8.snr.last_info[0].to(torch.float64).mean().item())
        e2e_u8["note"] = ("as e2e with uint8 image in (ingested as v/255 in K1) and uint8 image out (quantised in K3); the "
                          "8-bit pixels are a different input, so the PCG iteration count differs from the f32 run's")

    # ---- the other BASELINE.json configs, each against its roofline (N = 1 only) -----------------------
    stages = None
    if rank == 0 and world == 1 and not a.no_stages:
        h_img = h_lab = None
        torch.cuda.empty_cache()
        stages = run_stages(a, dev, peak)

    # ---- parity spot-check: two images of the timed batch against the oracle (rank 0) ---------------
    parity = None
    if rank == 0 and not a.no_parity_check:
        import multiprocessing as mp
        import numpy as np
        with mp.get_context("spawn").Pool(2) as pool:
            refs = pool.map(_cpu_chain_ref, [t[:3] for t in spot], chunksize=1)
        tol_cm = 5e4 * a.rtol + 1e-6
        worst, labels_ok, ok, ndiff, changed = 0.0, True, True, 0, 0.0
        for t, (ri, rl) in zip(spot, refs):
            gi = t[3].astype(np.float64)
            labels_ok &= bool(np.array_equal(t[4], rl))
            err = np.abs(gi - ri.astype(np.float64))
            worst = max(worst, float(err.max()))
            ndiff += int((err > 0).sum())
            changed = max(changed, float(np.abs(ri.astype(np.float64) - t[0]).max()))
            ok &= bool(np.isfinite(gi).all() and (err <= 0.4 * tol_cm + 1e-5 * np.maximum(np.abs(ri), 1e-3)).all())
        parity = {"images": [int(lo + j) for j in pick], "step": a.warmup + a.steps - 1, "labels_bit_exact": labels_ok,
                  "image_max_abs_err": worst, "pixels_differing": ndiff, "max_change_by_the_chain": changed, "image_tol": f"0.4 * (5e4 * rtol + 1e-6) + 1e-5 relative = {0.4 * tol_cm:.2e} + 1e-5 rel",
                  "pass": bool(labels_ok and ok),
                  "what": "chain output of the last timed step vs oracle.chain.full_chain (NumPy + SciPy splu) on the same inputs and parameters"}

    # ---- CPU baselines (rank 0, N=1 only) -----------------------------------------------------------
    cpu = cpu_cg = cfg0 = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or cores
        r, wall, per = cpu_oracle_rate(n, cores)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} images of the same synthetic 512x512 full-chain workload, oracle (NumPy + SciPy splu), "
                         f"Pool({cores}), wall {wall:.1f} s, {per:.2f} s/image/core"}
        cfg0 = config0_cpu()
        if cg_proc is not None:
            if cg_conn.poll(240):
                g = cg_conn.recv()
                cpu_cg = {"value": cores / g["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                          "seconds_per_image_per_core": g["seconds"], "iters": g["iters"], "relres": g["relres"], "rtol": a.rtol,
                          "sample": "confidence-map solve of ONE synthetic 512x512 image with scipy.sparse.linalg.cg + Jacobi to the GPU "
                                    "run's rtol on one core (the solve is > 99 % of the CPU chain); value = cores / seconds, i.e. "
                                    "extrapolated to one image per core"}
            else:
                cpu_cg = {"value": None, "note": "SciPy cg did not finish within the time budget"}
            cg_proc.join(5)

    if rank == 0:
        waves = (nloc + WAVE - 1) // WAVE
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if a.strong else "weak", "vs_baseline": None,
            "dtype": "f32 (fp64 dot products + fp64 residual refinement; the preconditioner's line factors and h plane are stored as bf16)", "data": "synthetic",
            "config": {"workload": "configs[4]: full chain (deform+TGC+remap, confidence-map PCG, SNR, reverb), "
                                   f"{nloc} images of {H}x{W} per GPU per step, {world} GPU(s) = batch {a.batch} "
                                   + ("(the literal configs[4] batch, strong scaling)" if a.strong else "(8 GPUs = configs[4]'s 4096)"),
                       "per_gpu_batch": nloc, "pcg_rtol": a.rtol, "pcg_precond": "column-line",
                       "pcg_iters_mean": float(iters_all.mean().item()), "pcg_iters_max": float(iters_all.max().item()),
                       "pcg_wave": WAVE, "l2_policy": f"inputs larger than L2 ({nloc * H * W * 5 / 2**20:.0f} MiB per GPU)"},
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_cg": cpu_cg, "config0_cpu": cfg0, "stages": stages,
            "parity_check": parity, "e2e": e2e, "e2e_u8": e2e_u8, "clocks": clocks,
            "gpu_launches": a.steps * (3 + 3 + 11 * waves),   # K1: 3, K3: 3, confidence map: 10 setup kernels + the persistent PCG kernel per wave
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
