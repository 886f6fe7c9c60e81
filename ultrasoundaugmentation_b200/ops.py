"""Functional layer: torch CUDA tensors in, torch CUDA tensors out, all compute in libusaug.so.

Every function enqueues on the CURRENT torch stream and never synchronises.  CPU tensors are
rejected (there is no CPU fallback); batches are [B,H,W] here -- the transform classes handle
the [H,W] / [1,H,W] / [B,1,H,W] conveniences.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _ffi


def gaussian_taps(sigma: float, radius: int | None = None) -> torch.Tensor:
    """exp(-k^2/2 sigma^2) normalised in fp64, rounded to fp32."""
    if radius is None:
        radius = int(math.ceil(3.0 * float(sigma)))
    radius = max(int(radius), 0)
    k = np.arange(-radius, radius + 1, dtype=np.float64)
    t = (k == 0).astype(np.float64) if sigma < 1e-6 else np.exp(-(k * k) / (2.0 * float(sigma) ** 2))
    t /= t.sum()
    return torch.from_numpy(t.astype(np.float32))


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, name: str, dtype) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (no CPU fallback exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _dev_param(v, B: int, device, dtype, name: str, width: int | None = None) -> torch.Tensor:
    t = torch.as_tensor(v, dtype=dtype)
    if t.dim() == 0:
        t = t.expand(B) if width is None else t.expand(B, width)
    want = (B,) if width is None else (B, width)
    if tuple(t.shape) != want:
        raise ValueError(f"{name} must have shape {want}, got {tuple(t.shape)}")
    return t.contiguous().to(device, non_blocking=True)


class ParamStager:
    """Per-call parameters in ONE pinned block and ONE host-to-device copy.

    Every per-sample parameter vector and tap table of a chain call (d, TGC gains, a_snr, rho, n_ghost, both
    Gaussian tap tables) is packed into one pinned host buffer and uploaded with a single ``cudaMemcpyAsync`` on
    the current stream; the kernels read 256-byte aligned views of the device copy.  No pageable staging copies
    (which are neither stream-ordered with respect to PDL launches nor capturable in a CUDA graph).  A small ring of
    pinned buffers, each guarded by an event, lets several calls be in flight.
    """

    def __init__(self, device, depth: int = 4):
        self.device = torch.device(device)
        self.depth = max(int(depth), 1)
        self._ring = [None] * self.depth      # [pinned float32 tensor, event]
        self._next = 0

    @staticmethod
    def _words(t: torch.Tensor) -> int:
        return (t.numel() + 63) // 64 * 64          # 256-byte granules of 4-byte words

    @staticmethod
    def _host(items: dict) -> dict:
        host = {}
        for k, v in items.items():
            t = torch.as_tensor(v)
            if t.is_cuda:
                raise ValueError("ParamStager takes host tensors")
            if t.dtype not in (torch.float32, torch.int32):
                t = t.to(torch.int32 if not t.dtype.is_floating_point else torch.float32)
            host[k] = t.contiguous()
        return host

    def _pack(self, pinned: torch.Tensor, host: dict) -> dict:
        off, views = 0, {}
        for k, t in host.items():
            n = t.numel()
            dst = pinned[off:off + n]
            (dst.view(torch.int32) if t.dtype == torch.int32 else dst).copy_(t.reshape(-1))
            views[k] = (off, n, t.dtype, tuple(t.shape))
            off += self._words(t)
        return views

    def rewrite(self, items: dict) -> None:
        """Host side only: new values into the pinned block of slot 0, same names and shapes as the last ``stage``
        (a captured CUDA graph re-reads that block on every replay).  The caller orders it after the previous replay."""
        host = self._host(items)
        entry = self._ring[0]
        if entry is None or self._layout != {k: (tuple(t.shape), t.dtype) for k, t in host.items()}:
            raise ValueError("ParamStager.rewrite: parameter names / shapes differ from the captured call")
        self._pack(entry[0], host)

    def stage(self, items: dict) -> dict:
        """items: name -> host tensor (float32 or int32).  Returns name -> device tensor of the same shape / dtype."""
        host = self._host(items)
        self._layout = {k: (tuple(t.shape), t.dtype) for k, t in host.items()}
        total = max(sum(self._words(t) for t in host.values()), 64)
        slot = self._next
        self._next = (slot + 1) % self.depth
        entry = self._ring[slot]
        if entry is None or entry[0].numel() < total:
            entry = [torch.empty(max(total, 1024), dtype=torch.float32).pin_memory(), None]
            self._ring[slot] = entry
        pinned, ev = entry
        capturing = torch.cuda.is_current_stream_capturing()
        if ev is not None and not capturing:         # (event queries are not allowed while a graph is being captured)
            ev.synchronize()                         # the copy that last read this buffer has completed (normally long ago)
        views = self._pack(pinned, host)
        with torch.cuda.device(self.device):
            dev = torch.empty(total, dtype=torch.float32, device=self.device)
            # one cudaMemcpyAsync on the current stream, issued by libusaug itself: torch's pinned-copy path also
            # records host-allocator events, which is not capturable in a CUDA graph
            _ffi.check(_ffi.load().usaug_upload(pinned.data_ptr(), dev.data_ptr(), total * 4, _stream_ptr(self.device)),
                       "usaug_upload")
            if not capturing:
                ev = torch.cuda.Event()
                ev.record()
                entry[1] = ev
        out = {}
        for k, (o, n, dt, shp) in views.items():
            v = dev[o:o + n]
            out[k] = (v.view(torch.int32) if dt == torch.int32 else v).reshape(shp)
        return out


def _ws(nbytes: int, device) -> torch.Tensor:
    """Caller-owned workspace of a C-ABI call.  Deliberately a fresh ``torch.empty`` per call: torch's caching allocator IS the
    workspace cache (a repeated size is a free-list hit, no cudaMalloc), and unlike a buffer cached on the chain object it is
    stream-safe -- ``augment_host_stream`` keeps three calls in flight, and a CUDA graph gets its private pool."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _check_pair(image, label, allow_u8_image=False):
    if allow_u8_image and isinstance(image, torch.Tensor) and image.dtype == torch.uint8:
        image = _need_cuda(image, "image", torch.uint8)
    else:
        image = _need_cuda(image, "image", torch.float32)
    label = _need_cuda(label, "label", torch.uint8)
    if image.dim() != 3 or image.shape != label.shape or image.device != label.device:
        raise ValueError(f"image/label must both be [B,H,W] on one device, got {tuple(image.shape)} / {tuple(label.shape)}")
    return image, label


def fused_tgc_deform(image, label, d, tgc, taps):
    """K1 (A.1 + A.2).  image f32 [B,H,W] in [0,1] -- or uint8 (8-bit B-mode ingest, read as v/255) --,
    label u8 [B,H,W], d [B], tgc [B,K], taps [2R+1].  Returns (image f32, label u8)."""
    lib = _ffi.load()
    image, label = _check_pair(image, label, allow_u8_image=True)
    B, H, W = image.shape
    dev = image.device
    tgc_t = torch.as_tensor(tgc, dtype=torch.float32)
    if tgc_t.dim() == 1:
        tgc_t = tgc_t.unsqueeze(0).expand(B, -1)
    K = int(tgc_t.shape[1])
    with torch.cuda.device(dev):
        d_t = _dev_param(d, B, dev, torch.float32, "d")
        tgc_t = _dev_param(tgc_t, B, dev, torch.float32, "tgc", K)
        taps_t = torch.as_tensor(taps, dtype=torch.float32).contiguous().to(dev, non_blocking=True)
        if taps_t.numel() % 2 != 1:
            raise ValueError("taps must have odd length 2R+1")
        R = (taps_t.numel() - 1) // 2
        out_img = torch.empty(image.shape, dtype=torch.float32, device=dev)
        out_lab = torch.empty_like(label)
        nbytes = lib.usaug_warp_workspace_bytes(B, H, W)
        ws = _ws(nbytes, dev)
        entry = lib.usaug_fused_tgc_deform_u8 if image.dtype == torch.uint8 else lib.usaug_fused_tgc_deform
        rc = entry(image.data_ptr(), label.data_ptr(), out_img.data_ptr(), out_lab.data_ptr(), B, H, W,
                   d_t.data_ptr(), tgc_t.data_ptr(), K, taps_t.data_ptr(), R, ws.data_ptr(), ws.numel(),
                   _stream_ptr(dev))
        _ffi.check(rc, "usaug_fused_tgc_deform")
    return out_img, out_lab


def confidence_map(image, alpha=2.0, beta=90.0, gamma=0.05, rtol=1e-8, maxiter=20000,
                   precond="line", fp64_vectors=False, zero_guess=False, max_wave=None,
                   return_info=False, timing: list | None = None, debug_timing=False, legacy_kernel=False, coarse=True,
                   twist=None, up_variant=None, level=True, sink="bottom_row", sink_mask=None):
    """K2 (A.3).  image f32 [B,H,W] -> cm f32 [B,H,W] (+ iters int32 [B], true relres f64 [B]).

    The batch is solved in waves of at most ``max_wave`` images (workspace ~56 B/pixel/image
    with fp32 vectors); default keeps the workspace under ~16 GiB.  ``timing`` (measurement
    only): a list that receives one (start_event, stop_event, lo, n) tuple per wave, recorded on
    the launching stream right around the persistent PCG kernel.

    Solver variants (all solve the same system to the same ``rtol``; they differ in rounding only).  By default the
    library picks them from the size of a launch -- pin them for results that do not depend on the batch size:
    ``coarse``: True (automatic), False, "fine" (2-row aggregates) or "standard" (4-row);
    ``twist``: None (automatic), True / False -- twisted column-line factorisation, two half sweeps per strip;
    ``up_variant``: None (automatic), "recompute" or "planes" -- edge weights recomputed or read in the UP phase;
    ``level``: False switches the level preconditioner (line solve on 1 x 4 pixel aggregates) off (A/B testing).

    ``sink``: where the random walks are absorbed besides the bottom row -- "bottom_row" (A.3),
    "bottom_and_sides" (also the first and last column) or "mask" (also every pixel of rows 1..H-2 where ``sink_mask``,
    uint8 [B,H,W], is non-zero).  The confidence map is exactly 0 on sink pixels.
    """
    lib = _ffi.load()
    image = _need_cuda(image, "image", torch.float32)
    if image.dim() != 3:
        raise ValueError(f"image must be [B,H,W], got {tuple(image.shape)}")
    B, H, W = image.shape
    if H < 3:
        raise ValueError("confidence map needs H >= 3")
    dev = image.device
    pc = {"jacobi": _ffi.PRECOND_JACOBI, "line": _ffi.PRECOND_LINE}[precond]
    flags = (_ffi.CM_FP64_VECTORS if fp64_vectors else 0) | (_ffi.CM_ZERO_GUESS if zero_guess else 0)
    flags |= 4 if debug_timing else 0
    flags |= 8 if legacy_kernel else 0     # A/B testing: two-slot kernel instead of the fused fp32 kernel
    if coarse not in (True, False, "fine", "standard"):
        raise ValueError(f"coarse must be True (automatic), False, 'fine' or 'standard', got {coarse!r}")
    flags |= {True: 0, False: 32, "fine": 64, "standard": 128}[coarse]   # coarse-grid correction: auto / off / 2-row / 4-row aggregates
    if twist not in (None, True, False):
        raise ValueError(f"twist must be None (automatic), True or False, got {twist!r}")
    flags |= {None: 0, True: _ffi.CM_TWIST_ON, False: _ffi.CM_TWIST_OFF}[twist]
    if up_variant not in (None, "recompute", "planes"):
        raise ValueError(f"up_variant must be None (automatic), 'recompute' or 'planes', got {up_variant!r}")
    flags |= {None: 0, "recompute": _ffi.CM_FORCE_RECOMPUTE, "planes": _ffi.CM_FORCE_PLANES}[up_variant]
    flags |= 0 if level else _ffi.CM_NO_LEVEL
    if sink not in _ffi.SINK_MODES:
        raise ValueError(f"sink must be one of {sorted(_ffi.SINK_MODES)}, got {sink!r}")
    sink_mode = _ffi.SINK_MODES[sink]
    if sink == "mask":
        if sink_mask is None:
            raise ValueError("sink='mask' needs sink_mask (uint8 [B,H,W])")
        sink_mask = _need_cuda(sink_mask, "sink_mask", torch.uint8)
        if sink_mask.shape != image.shape or sink_mask.device != dev:
            raise ValueError("sink_mask must have the image's shape and device")
    elif sink_mask is not None:
        raise ValueError("sink_mask is only read with sink='mask'")
    if sink_mode:
        flags |= _ffi.CM_SINKS          # workspace query: room for the sink plane
    with torch.cuda.device(dev):
        per_img = lib.usaug_confmap_workspace_bytes(1, H, W, pc, flags)
        if max_wave is None:
            max_wave = max(1, min(B, (16 << 30) // max(per_img, 1)))
        max_wave = max(1, min(int(max_wave), B))
        cm = torch.empty_like(image)
        iters = torch.full((B,), -1, dtype=torch.int32, device=dev)
        relres = torch.full((B,), float("nan"), dtype=torch.float64, device=dev)
        ws = _ws(lib.usaug_confmap_workspace_bytes(max_wave, H, W, pc, flags), dev)
        for lo in range(0, B, max_wave):
            n = min(max_wave, B - lo)
            if timing is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record(); ev[1].record()        # materialise the underlying cudaEvent_t handles
                lib.usaug_confmap_set_timing_events(ev[0].cuda_event, ev[1].cuda_event)
                timing.append((ev[0], ev[1], lo, n))
            rc = lib.usaug_confmap_pcg_sinks(image[lo:lo + n].data_ptr(), cm[lo:lo + n].data_ptr(), n, H, W,
                                             float(alpha), float(beta), float(gamma), float(rtol),
                                             int(maxiter), pc, flags, sink_mode,
                                             sink_mask[lo:lo + n].data_ptr() if sink_mode == 2 else None,
                                             iters[lo:lo + n].data_ptr(), relres[lo:lo + n].data_ptr(), ws.data_ptr(),
                                             ws.numel(), _stream_ptr(dev))
            if timing is not None:
                lib.usaug_confmap_set_timing_events(None, None)
            _ffi.check(rc, "usaug_confmap_pcg")
    if return_info:
        return cm, iters, relres
    return cm


def local_energy(image, lambda_min=8.0, mult=2.1, scales=3, sigma_f=0.55):
    """Monogenic-signal local energy.  image f32 [B,H,W], 4 <= H, W <= 2048 (sizes that
    are not powers of two are mirror-extended to the next one for the transform).  Returns (energy f32 [B,H,W], scale f32 [B]) with scale = 1 / max(energy) per image (0 for an
    image without band-pass content): energy * scale[:, None, None] is the normalised map in [0,1]."""
    lib = _ffi.load()
    image = _need_cuda(image, "image", torch.float32)
    if image.dim() != 3:
        raise ValueError(f"image must be [B,H,W], got {tuple(image.shape)}")
    B, H, W = image.shape
    dev = image.device
    for n, name in ((H, "H"), (W, "W")):
        if n < 4 or n > 2048:
            raise ValueError(f"local energy needs image sizes in [4, 2048], got {name}={n}")
    with torch.cuda.device(dev):
        energy = torch.empty_like(image)
        scale = torch.empty((B,), dtype=torch.float32, device=dev)
        ws = _ws(lib.usaug_local_energy_workspace_bytes(B, H, W), dev)
        rc = lib.usaug_local_energy(image.data_ptr(), energy.data_ptr(), scale.data_ptr(), B, H, W, float(lambda_min),
                                    float(mult), int(scales), float(sigma_f), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _ffi.check(rc, "usaug_local_energy")
    return energy, scale


def snr_reverb(image, cm, label, a_snr, rho, n_ghost, feather_taps, out_dtype=torch.float32, signal=None,
               signal_scale=None):
    """K3 (A.4 + A.5).  cm may be None (reverb only); n_ghost = 0 disables the ghosts.
    out_dtype=torch.uint8 quantises the result on the way out: floor(255 v + 0.5) (8-bit egress).
    signal / signal_scale (from ``local_energy``): SNR term (a-1) * C * signal * scale instead of (a-1) * C * I."""
    if out_dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"out_dtype must be torch.float32 or torch.uint8, got {out_dtype}")
    lib = _ffi.load()
    image, label = _check_pair(image, label)
    B, H, W = image.shape
    dev = image.device
    if cm is not None:
        cm = _need_cuda(cm, "cm", torch.float32)
        if cm.shape != image.shape:
            raise ValueError("cm must have the image's shape")
    with torch.cuda.device(dev):
        a_t = _dev_param(a_snr, B, dev, torch.float32, "a_snr")
        r_t = _dev_param(rho, B, dev, torch.float32, "rho")
        n_t = _dev_param(n_ghost, B, dev, torch.int32, "n_ghost")
        taps_t = torch.as_tensor(feather_taps, dtype=torch.float32).contiguous().to(dev, non_blocking=True)
        if taps_t.numel() % 2 != 1:
            raise ValueError("feather_taps must have odd length 2R+1")
        R = (taps_t.numel() - 1) // 2
        out = torch.empty(image.shape, dtype=out_dtype, device=dev)
        ws = _ws(lib.usaug_snr_reverb_workspace_bytes(B, H, W), dev)
        if signal is not None:
            signal = _need_cuda(signal, "signal", torch.float32)
            signal_scale = _need_cuda(signal_scale, "signal_scale", torch.float32)
            if signal.shape != image.shape or tuple(signal_scale.shape) != (B,):
                raise ValueError("signal must have the image's shape and signal_scale shape [B]")
            rc = lib.usaug_snr_reverb_signal(image.data_ptr(), 0 if cm is None else cm.data_ptr(), signal.data_ptr(),
                                             signal_scale.data_ptr(), label.data_ptr(), out.data_ptr(),
                                             1 if out_dtype == torch.uint8 else 0, B, H, W, a_t.data_ptr(), r_t.data_ptr(),
                                             n_t.data_ptr(), taps_t.data_ptr(), R, ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        else:
            entry = lib.usaug_snr_reverb_u8 if out_dtype == torch.uint8 else lib.usaug_snr_reverb
            rc = entry(image.data_ptr(), 0 if cm is None else cm.data_ptr(), label.data_ptr(), out.data_ptr(), B, H, W,
                       a_t.data_ptr(), r_t.data_ptr(), n_t.data_ptr(), taps_t.data_ptr(), R, ws.data_ptr(), ws.numel(),
                       _stream_ptr(dev))
        _ffi.check(rc, "usaug_snr_reverb")
    return out


def _scan_convert(image, label, geom, direction, H, W):
    lib = _ffi.load()
    image, label = _check_pair(image, label)
    B = image.shape[0]
    dev = image.device
    ns, nr = int(geom.n_samples), int(geom.n_rays)
    want = (B, H, W) if direction == 0 else (B, ns, nr)
    if tuple(image.shape) != want:
        raise ValueError(f"expected input of shape {want}, got {tuple(image.shape)}")
    out_shape = (B, ns, nr) if direction == 0 else (B, H, W)
    ax, ay, thm, rmin, dr, inv_dr, inv_dth = (float(v) for v in geom.scalars())
    with torch.cuda.device(dev):
        out_img = torch.empty(out_shape, dtype=torch.float32, device=dev)
        out_lab = torch.empty(out_shape, dtype=torch.uint8, device=dev)
        sin_t = cos_t = None
        if direction == 0:
            s, c = geom.tables()
            sin_t = torch.from_numpy(s).to(dev, non_blocking=True)
            cos_t = torch.from_numpy(c).to(dev, non_blocking=True)
        rc = lib.usaug_scan_convert(image.data_ptr(), label.data_ptr(), out_img.data_ptr(), out_lab.data_ptr(), B, H, W,
                                    ns, nr, direction, ax, ay, thm, rmin, dr, inv_dr, inv_dth,
                                    0 if sin_t is None else sin_t.data_ptr(), 0 if cos_t is None else cos_t.data_ptr(),
                                    _stream_ptr(dev))
        _ffi.check(rc, "usaug_scan_convert")
    return out_img, out_lab


def cart_to_polar(image, label, geom):
    """Cartesian [B,H,W] -> polar [B,n_samples,n_rays] of a ``geometry.ConvexGeometry`` (image bilinear, label nearest)."""
    return _scan_convert(image, label, geom, 0, int(image.shape[1]), int(image.shape[2]))


def polar_to_cart(image_p, label_p, geom, H, W):
    """polar [B,n_samples,n_rays] -> Cartesian [B,H,W]; pixels outside the sector are 0."""
    return _scan_convert(image_p, label_p, geom, 1, int(H), int(W))


def bone_box(label):
    """[B,4] int32 {y_top, y_bottom, x_left, x_right}; y_top = -1 for an image without bone."""
    lib = _ffi.load()
    label = _need_cuda(label, "label", torch.uint8)
    if label.dim() != 3:
        raise ValueError("label must be [B,H,W]")
    B, H, W = label.shape
    dev = label.device
    with torch.cuda.device(dev):
        box = torch.empty((B, 4), dtype=torch.int32, device=dev)
        ws = _ws(lib.usaug_bone_box_workspace_bytes(B, H, W), dev)
        rc = lib.usaug_bone_box(label.data_ptr(), B, H, W, box.data_ptr(), ws.data_ptr(), ws.numel(),
                                _stream_ptr(dev))
        _ffi.check(rc, "usaug_bone_box")
    return box
