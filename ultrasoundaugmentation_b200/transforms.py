"""Per-transform callables with a parameter-sampling API (the dataloader-facing boundary).

The reference mount (/root/reference/README.md:1-2) defines no API, so this follows the
de-facto PyTorch transform protocol that SURVEY.md section 8(b) pins:

    t = ProbePressureDeformation(...)            # constructed once, picklable (config only)
    params = t.get_params(batch_size, generator=None, sample_index=None, image_size=(H, W))
    image_out, label_out = t(image, label, params=None)

image: float32 in [0,1], label: uint8 (bone = non-zero); shapes [H,W], [1,H,W], [B,H,W] or
[B,1,H,W]; CUDA tensors only (TypeError / ValueError otherwise -- there is no CPU fallback).
Outputs are new tensors on the same device, enqueued on the current stream, no host sync.
Physics: SURVEY.md Appendix A.1-A.5; kernels: csrc/*.cu through the C ABI in include/usaug.h.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .params import uniforms

__all__ = ["ProbePressureDeformation", "TimeGainCompensation", "ConfidenceSNR", "Reverberation",
           "PhysicsChain", "GraphedChain"]


def _to_bhw(image: torch.Tensor, label: torch.Tensor):
    if not isinstance(image, torch.Tensor) or not isinstance(label, torch.Tensor):
        raise TypeError("image and label must be torch tensors")
    if image.shape != label.shape:
        raise ValueError(f"image {tuple(image.shape)} and label {tuple(label.shape)} differ in shape")
    shp = tuple(image.shape)
    if image.dim() == 2:
        v = (1,) + shp
    elif image.dim() == 3:
        v = shp
    elif image.dim() == 4 and shp[1] == 1:
        v = (shp[0], shp[2], shp[3])
    else:
        raise ValueError(f"expected [H,W], [1,H,W], [B,H,W] or [B,1,H,W], got {shp}")
    return image.reshape(v), label.reshape(v), shp


def _as_tensor_dict(d: dict) -> dict:
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)))
            for k, v in d.items()}


class _Transform:
    """Common machinery: application probability p, seed, RNG stream id."""

    _stream = 0

    def __init__(self, p: float = 1.0, seed: int = 0):
        if not 0.0 <= p <= 1.0:
            raise ValueError("p must be in [0,1]")
        self.p = float(p)
        self.seed = int(seed)

    def _draw(self, B, n, generator, sample_index):
        # column 0 decides application; columns 1.. are the transform's own draws
        u = uniforms(B, n + 1, seed=self.seed, stream=self._stream, generator=generator,
                     sample_index=sample_index)
        return u[:, 0] < self.p, u[:, 1:]

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None) -> dict:
        raise NotImplementedError

    def __call__(self, image, label, params: Optional[dict] = None, sample_index=None):
        chain = PhysicsChain([self])
        return chain(image, label, params=params, sample_index=sample_index)


class ProbePressureDeformation(_Transform):
    """A.1: axial displacement field along scanlines, scaled by depth and the bone surface.

    d ~ U(d_range) * H pixels (positive = more pressure; clamped on device to half the
    shallowest smoothed bone depth).  sigma_rel*W is the Gaussian sigma of the surface profile.
    """

    _stream = 1

    def __init__(self, d_range: Sequence[float] = (-0.05, 0.15), sigma_rel: float = 0.02,
                 p: float = 1.0, seed: int = 0):
        super().__init__(p, seed)
        self.d_range = (float(d_range[0]), float(d_range[1]))
        self.sigma_rel = float(sigma_rel)

    def taps(self, W: int) -> torch.Tensor:
        return ops.gaussian_taps(self.sigma_rel * W)

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None):
        if image_size is None:
            raise ValueError("ProbePressureDeformation.get_params needs image_size=(H, W)")
        H = int(image_size[0])
        apply, u = self._draw(batch_size, 1, generator, sample_index)
        lo, hi = self.d_range
        d = (lo + (hi - lo) * u[:, 0]) * H
        d = np.where(apply, d, 0.0).astype(np.float32)
        return _as_tensor_dict({"d": d})


class TimeGainCompensation(_Transform):
    """A.2 TGC: K control gains ~ U(gain_range), linear in output depth."""

    _stream = 2

    def __init__(self, num_gains: int = 8, gain_range: Sequence[float] = (0.6, 1.4), p: float = 1.0,
                 seed: int = 0):
        super().__init__(p, seed)
        if num_gains < 1:
            raise ValueError("num_gains must be >= 1")
        self.num_gains = int(num_gains)
        self.gain_range = (float(gain_range[0]), float(gain_range[1]))

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None):
        apply, u = self._draw(batch_size, self.num_gains, generator, sample_index)
        lo, hi = self.gain_range
        g = lo + (hi - lo) * u
        g = np.where(apply[:, None], g, 1.0).astype(np.float32)
        return _as_tensor_dict({"tgc": g})


class ConfidenceSNR(_Transform):
    """A.3 + A.4: confidence map of the (deformed) image, then I + (a-1) * C * S, clipped.

    ``signal="intensity"`` (default, the hot path): S = I.  ``signal="local_energy"``:
    S = the image's monogenic-signal local energy, normalised to [0,1] -- sum of ``scales`` log-Gabor band-pass
    filters with wavelengths ``lambda_min * mult**s`` pixels and their Riesz transforms (image sizes 4...2048;
    the hand-written FFT runs on the next power of two, mirror-extended).

    ``sink``: where the random walks of the confidence map are absorbed besides the bottom row:
    "bottom_row" (A.3), "bottom_and_sides" (also the first and the last scanline) or "label" (also every bone pixel of
    the label plane the transform is given -- in a chain, the deformed label)."""

    _stream = 3

    def __init__(self, a_range: Sequence[float] = (0.7, 1.4), alpha: float = 2.0, beta: float = 90.0,
                 gamma: float = 0.05, rtol: float = 1e-8, maxiter: int = 20000, precond: str = "line",
                 fp64_vectors: bool = False, max_wave: Optional[int] = None, p: float = 1.0,
                 seed: int = 0, signal: str = "intensity", lambda_min: float = 8.0, mult: float = 2.1,
                 scales: int = 3, sigma_f: float = 0.55, sink: str = "bottom_row"):
        super().__init__(p, seed)
        if sink not in ("bottom_row", "bottom_and_sides", "label"):
            raise ValueError("sink must be 'bottom_row', 'bottom_and_sides' or 'label'")
        self.sink = sink
        if precond not in ("line", "jacobi"):
            raise ValueError("precond must be 'line' or 'jacobi'")
        if signal not in ("intensity", "local_energy"):
            raise ValueError("signal must be 'intensity' or 'local_energy'")
        if not (1 <= int(scales) <= 4 and lambda_min >= 2.0 and mult > 0.0 and 0.0 < sigma_f < 1.0):
            raise ValueError("local-energy filter: 1 <= scales <= 4, lambda_min >= 2, mult > 0, 0 < sigma_f < 1")
        self.signal = signal
        self.lambda_min, self.mult, self.scales, self.sigma_f = float(lambda_min), float(mult), int(scales), float(sigma_f)
        self.a_range = (float(a_range[0]), float(a_range[1]))
        self.alpha, self.beta, self.gamma = float(alpha), float(beta), float(gamma)
        self.rtol, self.maxiter, self.precond = float(rtol), int(maxiter), precond
        self.fp64_vectors = bool(fp64_vectors)
        self.max_wave = max_wave
        self.last_info = None      # (iters, relres) device tensors of the most recent call
        self.timing = None         # measurement only: list receiving per-wave PCG kernel events

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None):
        apply, u = self._draw(batch_size, 1, generator, sample_index)
        lo, hi = self.a_range
        a = np.where(apply, lo + (hi - lo) * u[:, 0], 1.0).astype(np.float32)
        return _as_tensor_dict({"a_snr": a})

    def confidence_map(self, image_bhw: torch.Tensor, label_bhw: Optional[torch.Tensor] = None) -> torch.Tensor:
        sink, mask = self.sink, None
        if sink == "label":
            if label_bhw is None:
                raise ValueError("sink='label' needs the label plane")
            sink, mask = "mask", label_bhw
        cm, it, rr = ops.confidence_map(image_bhw, self.alpha, self.beta, self.gamma, self.rtol,
                                        self.maxiter, self.precond, self.fp64_vectors,
                                        max_wave=self.max_wave, return_info=True, timing=self.timing,
                                        sink=sink, sink_mask=mask)
        self.last_info = (it, rr)
        return cm

    def signal_map(self, image_bhw: torch.Tensor):
        """(signal plane, per-image scale) for ops.snr_reverb, or (None, None) for signal = intensity."""
        if self.signal == "intensity":
            return None, None
        return ops.local_energy(image_bhw, self.lambda_min, self.mult, self.scales, self.sigma_f)

    def __getstate__(self):
        s = dict(self.__dict__)
        s["last_info"] = None
        s["timing"] = None
        return s


class Reverberation(_Transform):
    """A.5: ghosts of the bone region at multiples of the bone-surface depth, attenuated by rho^k."""

    _stream = 4

    def __init__(self, rho_range: Sequence[float] = (0.3, 0.7), max_ghosts: int = 2,
                 sigma_feather: float = 2.0, p: float = 1.0, seed: int = 0):
        super().__init__(p, seed)
        if max_ghosts < 1:
            raise ValueError("max_ghosts must be >= 1")
        if math.ceil(3.0 * sigma_feather) > 32:
            raise ValueError("sigma_feather too large (tap radius must be <= 32)")
        self.rho_range = (float(rho_range[0]), float(rho_range[1]))
        self.max_ghosts = int(max_ghosts)
        self.sigma_feather = float(sigma_feather)

    def taps(self) -> torch.Tensor:
        return ops.gaussian_taps(self.sigma_feather)

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None):
        apply, u = self._draw(batch_size, 2, generator, sample_index)
        lo, hi = self.rho_range
        rho = (lo + (hi - lo) * u[:, 0]).astype(np.float32)
        n = np.minimum(1 + np.floor(u[:, 1] * self.max_ghosts), self.max_ghosts).astype(np.int32)
        n = np.where(apply, n, 0).astype(np.int32)
        return _as_tensor_dict({"rho": rho, "n_ghost": n})


class PhysicsChain:
    """deformation (+TGC, one fused launch) -> confidence-map SNR -> reverberation (one fused launch).

    Accepts any subset of the four transforms (at most one of each); the physical order is fixed
    (SURVEY.md Appendix A "Chain order").

    8-bit pipelines: a uint8 image is ingested as v/255 inside the first kernel, and
    ``output_dtype=torch.uint8`` quantises the result inside the last one (floor(255 v + 0.5)) -- 2 instead of
    5 bytes per pixel over PCIe in each direction.  The arithmetic in between is fp32 either way.
    """

    def __init__(self, transforms: Sequence[_Transform], output_dtype=torch.float32, geometry=None):
        if output_dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"output_dtype must be torch.float32 or torch.uint8, got {output_dtype}")
        if geometry is not None and output_dtype != torch.float32:
            raise ValueError("a probe geometry needs float32 images in and out")
        self.output_dtype = output_dtype
        # geometry.ConvexGeometry: the chain runs on the polar grid of rays x samples;
        # None = linear probe, scanlines are the image columns
        self.geometry = geometry
        self.deform = self.tgc = self.snr = self.reverb = None
        for t in transforms:
            if isinstance(t, ProbePressureDeformation):
                slot = "deform"
            elif isinstance(t, TimeGainCompensation):
                slot = "tgc"
            elif isinstance(t, ConfidenceSNR):
                slot = "snr"
            elif isinstance(t, Reverberation):
                slot = "reverb"
            else:
                raise TypeError(f"unsupported transform {type(t).__name__}")
            if getattr(self, slot) is not None:
                raise ValueError(f"duplicate {slot} transform")
            setattr(self, slot, t)
        self.transforms = [t for t in (self.deform, self.tgc, self.snr, self.reverb) if t is not None]
        if not self.transforms:
            raise ValueError("empty chain")
        self._stagers = {}         # device -> ops.ParamStager (pinned parameter blocks; never pickled)
        self._host_ring = {}       # (device, depth, dtype) -> pinned output buffers of augment_host_stream (never pickled)

    def __getstate__(self):
        s = dict(self.__dict__)
        s["_stagers"] = {}
        s["_host_ring"] = {}
        return s

    def _stager(self, device) -> "ops.ParamStager":
        st = self._stagers.get(device)
        if st is None:
            st = self._stagers[device] = ops.ParamStager(device)
        return st

    @classmethod
    def full(cls, seed: int = 0, output_dtype=torch.float32, geometry=None, **snr_kwargs) -> "PhysicsChain":
        return cls([ProbePressureDeformation(seed=seed), TimeGainCompensation(seed=seed),
                    ConfidenceSNR(seed=seed, **snr_kwargs), Reverberation(seed=seed)], output_dtype=output_dtype,
                   geometry=geometry)

    def get_params(self, batch_size, generator=None, sample_index=None, image_size=None) -> dict:
        out = {}
        for t in self.transforms:
            out.update(t.get_params(batch_size, generator, sample_index, image_size))
        return out

    def _scanline_grid(self, H, W):
        """(samples per scanline, scanlines) the transforms act on: the image itself, or the polar grid of a convex probe."""
        return (H, W) if self.geometry is None else (int(self.geometry.n_samples), int(self.geometry.n_rays))

    # ---- device path --------------------------------------------------------------------------
    def __call__(self, image, label, params: Optional[dict] = None, sample_index=None):
        img, lab, shp = _to_bhw(image, label)
        if not img.is_cuda or not lab.is_cuda:
            raise ValueError("PhysicsChain needs CUDA tensors (no CPU fallback exists); "
                             "use augment_host() for host buffers")
        if img.dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"image must be float32 (or uint8 for 8-bit ingest), got {img.dtype}")
        if lab.dtype != torch.uint8:
            raise TypeError(f"label must be uint8, got {lab.dtype}")
        B, H, W = img.shape
        if params is None:
            params = self.get_params(B, None, sample_index, self._scanline_grid(H, W))
        out_img, out_lab = self._run(img.contiguous(), lab.contiguous(), params)
        return out_img.reshape(shp), out_lab.reshape(shp)

    def _run(self, img, lab, params):
        if self.geometry is not None:
            if img.dtype != torch.float32:
                raise TypeError("a probe geometry needs a float32 image")
            Hc, Wc = int(img.shape[1]), int(img.shape[2])
            pi, pl = ops.cart_to_polar(img, lab, self.geometry)
            pi, pl = self._run_scanlines(pi, pl, params)
            return ops.polar_to_cart(pi, pl, self.geometry, Hc, Wc)
        return self._run_scanlines(img, lab, params)

    def _host_params(self, B, W, params, image_dtype):
        """(k1, k3, host dict): which fused kernels run and every parameter they need, as host tensors."""
        u8_out = self.output_dtype == torch.uint8
        k1 = self.deform is not None or self.tgc is not None or image_dtype == torch.uint8
        k3 = self.snr is not None or self.reverb is not None or u8_out
        host = {}
        if k1:
            host["d"] = params["d"] if self.deform is not None else torch.zeros(B)
            host["tgc"] = params["tgc"] if self.tgc is not None else torch.ones(B, 1)
            host["taps"] = self.deform.taps(W) if self.deform is not None else ops.gaussian_taps(0.0, 0)
        if k3:
            host["a_snr"] = params["a_snr"] if self.snr is not None else torch.ones(B)
            if self.reverb is not None:
                host["rho"], host["n_ghost"], host["ftaps"] = params["rho"], params["n_ghost"], self.reverb.taps()
            else:
                host["rho"], host["n_ghost"], host["ftaps"] = torch.zeros(B), torch.zeros(B, dtype=torch.int32), ops.gaussian_taps(0.0, 0)
        for k, v in host.items():
            want = torch.int32 if k == "n_ghost" else torch.float32
            host[k] = torch.as_tensor(v).to(want)
            if k in ("d", "a_snr", "rho", "n_ghost") and tuple(host[k].shape) != (B,):
                raise ValueError(f"{k} must have shape ({B},), got {tuple(host[k].shape)}")
        return k1, k3, host

    def _run_scanlines(self, img, lab, params):
        B, H, W = img.shape
        # every parameter of the call in one pinned block, one host-to-device copy (ops.ParamStager)
        k1, k3, host = self._host_params(B, W, params, img.dtype)
        if any(v.is_cuda for v in host.values()):
            dv = {k: v.to(img.device) for k, v in host.items()}     # caller staged the parameters itself
        else:
            dv = self._stager(img.device).stage(host)
        # a uint8 image always passes the first kernel (identity warp if need be): that is where it is ingested
        if k1:
            img, lab = ops.fused_tgc_deform(img, lab, dv["d"], dv["tgc"], dv["taps"])
        cm = sig = sig_scale = None
        if self.snr is not None:
            cm = self.snr.confidence_map(img, lab)
            sig, sig_scale = self.snr.signal_map(img)
        # ... and a uint8 result always leaves through the last kernel (pass-through if need be)
        if k3:
            img = ops.snr_reverb(img, cm, lab, dv["a_snr"], dv["rho"], dv["n_ghost"], dv["ftaps"], out_dtype=self.output_dtype,
                                 signal=sig, signal_scale=sig_scale)
        return img, lab

    # ---- CUDA graph: the whole chain as ONE submission -----------------------------------------
    def capture(self, image, label) -> "GraphedChain":
        """Capture the chain for this batch shape into a CUDA graph (one submission per call instead of ~20 launches).
        ``image`` / ``label`` give shape, dtype and device; see GraphedChain."""
        return GraphedChain(self, image, label)

    # ---- host path: pinned host buffers in, pinned host buffers out ---------------------------
    def augment_host(self, image: torch.Tensor, label: torch.Tensor, params: Optional[dict] = None,
                     sample_index=None, device=None, out: Optional[tuple] = None):
        """Host tensors -> H2D -> chain on the GPU -> D2H.  Not a CPU fallback: compute is CUDA.

        Everything is enqueued on the current stream of `device`; the call returns after the
        stream has finished so the host buffers are valid.
        """
        if image.is_cuda or label.is_cuda:
            raise ValueError("augment_host takes host tensors; call the chain directly for CUDA tensors")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        img, lab, shp = _to_bhw(image, label)
        B, H, W = img.shape
        if params is None:
            params = self.get_params(B, None, sample_index, self._scanline_grid(H, W))
        with torch.cuda.device(device):
            d_img = img.contiguous().to(device, non_blocking=True)
            d_lab = lab.contiguous().to(device, non_blocking=True)
            o_img, o_lab = self._run(d_img, d_lab, params)
            if out is None:
                h_img = torch.empty((B, H, W), dtype=self.output_dtype, pin_memory=True)
                h_lab = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
            else:
                h_img, h_lab = out[0].reshape(B, H, W), out[1].reshape(B, H, W)
            h_img.copy_(o_img, non_blocking=True)
            h_lab.copy_(o_lab, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
        return h_img.reshape(shp), h_lab.reshape(shp)

    def augment_host_stream(self, batches, device=None, depth: int = 2):
        """Double-buffered host pipeline: generator over augmented batches.

        ``batches`` yields ``(image, label)`` or ``(image, label, sample_index)`` host tensors (pinned
        memory for true overlap).  The host-to-device copy of batch i+1 and the device-to-host copy of
        batch i-1 run on their own streams while batch i computes, so PCIe traffic hides behind the
        kernels.  Results come back in order as pinned host tensors from a ring of ``depth + 2`` buffer
        pairs: a yielded pair stays valid while the consumer fetches ONE further batch -- copy it if it has
        to live longer.  Input tensors must stay unmodified until their batch has been yielded.  The ring is kept on the
        chain and reused by later calls (run ONE such pipeline at a time per chain object).
        Bitwise identical to ``augment_host`` batch by batch.
        """
        if depth < 1:
            raise ValueError("depth must be >= 1")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(device):
            s_in, s_run, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device), torch.cuda.Stream(device)
            # the pinned output ring lives on the chain: cudaHostAlloc of its depth + 2 buffer pairs (3.2 GB for 512 images of
            # 512 x 512) costs about a second, once -- not once per epoch
            ring, pending = self._host_ring.setdefault((device, depth, self.output_dtype), {}), []

            def drain_one():
                ev, h_img, h_lab, shp, keep = pending.pop(0)
                ev.synchronize()
                del keep
                return h_img.reshape(shp), h_lab.reshape(shp)

            for i, item in enumerate(batches):
                image, label = item[0], item[1]
                sample_index = item[2] if len(item) > 2 else None
                if image.is_cuda or label.is_cuda:
                    raise ValueError("augment_host_stream takes host tensors")
                img, lab, shp = _to_bhw(image, label)
                B, H, W = img.shape
                params = self.get_params(B, None, sample_index, self._scanline_grid(H, W))
                with torch.cuda.stream(s_in):
                    d_img = img.contiguous().to(device, non_blocking=True)
                    d_lab = lab.contiguous().to(device, non_blocking=True)
                    ev_in = torch.cuda.Event(); ev_in.record(s_in)
                with torch.cuda.stream(s_run):
                    s_run.wait_event(ev_in)
                    d_img.record_stream(s_run); d_lab.record_stream(s_run)
                    o_img, o_lab = self._run(d_img, d_lab, params)
                    ev_run = torch.cuda.Event(); ev_run.record(s_run)
                key = (i % (depth + 2), B, H, W)
                if key not in ring:
                    ring[key] = (torch.empty((B, H, W), dtype=self.output_dtype, pin_memory=True),
                                 torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True))
                h_img, h_lab = ring[key]
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_run)
                    o_img.record_stream(s_out); o_lab.record_stream(s_out)
                    h_img.copy_(o_img, non_blocking=True)
                    h_lab.copy_(o_lab, non_blocking=True)
                    ev_out = torch.cuda.Event(); ev_out.record(s_out)
                pending.append((ev_out, h_img, h_lab, shp, (d_img, d_lab, o_img, o_lab)))
                if len(pending) > depth:
                    yield drain_one()
            while pending:
                yield drain_one()


class GraphedChain:
    """A PhysicsChain captured into a CUDA graph for one batch shape (SURVEY.md 7.2 hard part 5: launch overhead).

    Everything a call enqueues -- the single parameter upload from a pinned block, the fused K1 launch group, the
    confidence-map setup kernels and its persistent cooperative PCG kernel, the fused K3 launch group -- becomes one
    graph launch.  ``g(image, label, params=None, sample_index=None)`` copies the inputs into the graph's static
    buffers, rewrites the pinned parameter block and replays; it returns the graph's static output tensors, which
    the NEXT call overwrites (clone them to keep them).  Results are bitwise those of the eager chain.
    """

    def __init__(self, chain: PhysicsChain, image: torch.Tensor, label: torch.Tensor):
        if chain.geometry is not None:
            raise ValueError("a chain with a probe geometry cannot be captured (its table uploads are not graph-safe)")
        img, lab, shp = _to_bhw(image, label)
        if not img.is_cuda:
            raise ValueError("capture needs CUDA tensors")
        B, H, W = img.shape
        dev = img.device
        self.chain, self.shape, self.device = chain, shp, dev
        self.image, self.label = img.contiguous().clone(), lab.contiguous().clone()
        self._stager = ops.ParamStager(dev, depth=1)
        self._done = None
        params = chain.get_params(B, None, None, chain._scanline_grid(H, W))
        saved = chain._stagers.get(dev)
        chain._stagers[dev] = self._stager
        try:
            with torch.cuda.device(dev):
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(2):                     # warm-up outside the capture: lazy initialisation, allocator
                        chain._run(self.image, self.label, params)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out_image, self.out_label = chain._run(self.image, self.label, params)
                self.info = chain.snr.last_info if chain.snr is not None else None
        finally:
            if saved is None:
                chain._stagers.pop(dev, None)
            else:
                chain._stagers[dev] = saved

    def __call__(self, image, label, params: Optional[dict] = None, sample_index=None):
        img, lab, shp = _to_bhw(image, label)
        if tuple(img.shape) != tuple(self.image.shape) or img.dtype != self.image.dtype or img.device != self.device:
            raise ValueError("GraphedChain was captured for another shape, dtype or device")
        B, H, W = img.shape
        if params is None:
            params = self.chain.get_params(B, None, sample_index, self.chain._scanline_grid(H, W))
        _, _, host = self.chain._host_params(B, W, params, img.dtype)
        if self._done is not None:
            self._done.synchronize()                       # the previous replay has read the pinned block
        self._stager.rewrite(host)
        with torch.cuda.device(self.device):
            self.image.copy_(img, non_blocking=True)
            self.label.copy_(lab, non_blocking=True)
            self.graph.replay()
            self._done = torch.cuda.Event()
            self._done.record()
        return self.out_image.reshape(shp), self.out_label.reshape(shp)
