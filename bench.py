#!/usr/bin/env python
"""Benchmark of the hot path: sorted-Gaussian rasterizer forward + backward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3b|C2|C3a|C4|C5|C1]
                    [--sharding auto|bands|views] [--no-extras]

A "step" is one fwd+bwd pass of the rasterizer over one view of the synthetic cloud of SURVEY 8(d).
Default workload = the configuration north_star's target is quoted on (BASELINE.json configs[2], "C3b"):
4M Gaussians, 1920x1080, StopThePop preset (HIER sort, per-tile depth, tile / 4x4 culling), fwd+bwd.
N>1 (default sharding of C3a / C3b / C4): ONE frame, tile-row bands sharded across the ranks (north_star: "sharding
screen-space tile ranges ... >= 6x aggregate at 8 GPUs tile-sharded"): every rank holds the same Gaussians,
preprocesses all of them, bins / sorts / renders only its band (bands balanced by the instance counts of a warm-up
frame); exchange = ONE NCCL all-gather of the image bands (forward) and ONE all-reduce of the 36 B/Gaussian packed
screen-space gradient accumulator (backward), after which every rank finishes the per-Gaussian backward and holds the
full gradients.  Total work is fixed => "scaling": "strong".
--sharding views (default of C2 / C5): one camera per rank (yaw = rank * 0.08 rad), ONE logical all-reduce of the
parameter-gradient slab, overlapped with the preprocess-backward stage; per-GPU work fixed => "weak".

Printed JSON line (rank 0):
  value     Mpixels/s, whole job, Gaussians + camera + upstream gradient resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public autograd API (GaussianRasterizer) with the step's inputs (camera
            matrices + upstream-gradient image) copied from pinned HOST memory and the rendered image read back
            to the host inside the timed region.  The Gaussian parameters are the model state and stay resident.
            `e2e_full_upload` additionally uploads every Gaussian parameter and downloads every gradient.
  roofline  dominant kernel: algorithmic bytes (SURVEY 8d formula at the measured P,V,R) / CUDA-event time of
            that kernel averaged over the timed steps (stage events recorded inside the library, resolved lazily);
            `traffic` = DRAM bytes of that kernel from the committed ncu capture named in `traffic_source`
  extra     short measurements of the other BASELINE.json configurations in the same run (N=1: C2, C4; N>1: C4 tile
            bands), each with its own value / e2e / stage times -- not the headline
  cpu_baseline  the plain-C oracle port (OpenMP, all host cores) on a bounded sample, N=1 only
--impl reference: the reference's own implementation.  The reference ships NO CPU path, so this
arm times the UNMODIFIED reference CUDA build (oracle/_ref, compiled from /root/reference for sm_100) on the same
GPU through its own _C API, same workload, same copies; if oracle/_ref is absent it falls back to the C oracle port.
The reference cannot shard a frame: for tile-band workloads at N>1 rank 0 alone runs the whole frame on one GPU
("for C4 at 2/4/8 GPUs the baseline stays the 1-GPU reference number") and the other ranks exit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "stopthepop-rasterization_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import stp_scenes as S  # noqa: E402

WORKLOADS = {
    # name: (scene config, settings overrides, description)
    "C1": ("C1", dict(), "C1: 1k Gaussians, 256x256, GLOBAL"),
    "C2": ("C2", dict(), "C2: 1M Gaussians, 1920x1080, GLOBAL tile sort, fwd+bwd (BASELINE.json configs[1])"),
    "C3a": ("C3", dict(sort_mode=3), "C3a: 4M Gaussians, 1920x1080, HIER 64/8/4, fwd+bwd"),
    "C3b": ("C3", dict(S.STOPTHEPOP_PRESET), "C3b: 4M Gaussians, 1920x1080, StopThePop preset, fwd+bwd"),
    # BASELINE.json configs[3]: one 4K view, full per-pixel sort, screen tiles sharded across the ranks (tile-row bands)
    "C4": ("C4", dict(sort_mode=1), "C4: 4M Gaussians, 3840x2160, PPX_FULL, fwd+bwd, tile-row bands sharded across ranks"),
    # BASELINE.json configs[4]: 10M Gaussians, one 1080p view per rank (8 views at 8 GPUs), StopThePop preset
    "C5": ("C5", dict(S.STOPTHEPOP_PRESET), "C5: 10M Gaussians, one 1920x1080 view per rank, StopThePop preset, fwd+bwd"),
}
# default multi-GPU split: tile-row bands of ONE frame (strong scaling) / one view per rank (weak scaling)
BAND_DEFAULT = {"C3a", "C3b", "C4"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  Uses NVML
    in-process (nvidia_ml_py) from a background thread: a query costs tens of microseconds, whereas an `nvidia-smi -lms`
    child process was measured to stall kernel launches for tens of milliseconds per poll (visible as a 30-45 % longer
    step on the 1.8 ms C2 step).  Falls back to nvidia-smi if NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first, self.stop_flag, self.nvml = index, [], None, 0, False, None

    def mark(self):
        self.first = len(self.rows)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi takes a few hundred ms to attach to the driver and stalls kernel launches while it does:
            # wait for its first sample so that this start-up never overlaps the warm-up or the timed region
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10.0 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append([str(sm), str(self.max_sm), "0"] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            time.sleep(0.03)
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def algorithmic_bytes(P, V, R, N, T, M, s_flag, c_flag):
    """SURVEY.md 8(d) per-stage algorithmic bytes (fixed formula, measured P,V,R)."""
    passes = 6
    return {
        "Preprocess": 12 * P + V * (12 + 16 + 4 + 12 * M) + 8 * P + V * (4 + 8 + 8 + 24 + 16 + 12 + 3 + 48 * c_flag) + 8 * P,
        "Duplicate": 8 * P + 20 * V + 12 * R,
        "Sort": passes * 24 * R + 8 * R + 8 * R + 16 * T,
        "Render": 8 * T + R * (28 + 48 * s_flag) + 12 * V + 20 * N,
        "RenderBackward": 8 * T + R * (40 + 48 * s_flag) + N * (20 + 12 * s_flag) + 72 * V,
        "PreprocessBackward": 4 * P + 88 * V + 4 * P + V * (103 + 12 * M) + V * (40 + 12 * M),
    }


KERNEL_OF_STAGE = {"Preprocess": "preprocess_kernel", "Duplicate": "duplicate_kernel", "Sort": "tile_sort_small_kernel + tile_sort_large_kernel",
                   "Render": "render_*_fwd_kernel", "RenderBackward": "render_*_bwd_kernel",
                   "PreprocessBackward": "preprocess_bwd_kernel"}


def kernel_name(stage, settings):
    """the kernel(s) behind a stage for these settings (names as they appear in the ncu captures under profiles/)"""
    ss = settings["sort_settings"]
    mode, q = ss["sort_mode"], ss["queue_sizes"]
    cull = int(bool(settings["culling_settings"]["hierarchical_4x4_culling"]))
    if stage == "Render":
        return {0: "render_global_fwd_kernel", 1: "render_full_fast_kernel (+ render_full_kernel for lists > 1024)",
                2: f"render_kbuffer_kernel<{q['per_pixel']},fwd>",
                3: f"render_hier_kernel<{q['per_pixel']},{q['tile_2x2']},{cull},0>"}[mode]
    if stage == "RenderBackward":
        return {0: "render_global_bwd_kernel", 1: "blend_replay_bwd_kernel<2,0>",
                2: f"render_kbuffer_kernel<{q['per_pixel']},bwd>",
                3: "blend_replay_bwd_kernel<1,0> (+ render_hier_kernel<..,1> for pixels whose log overflowed)"}[mode]
    return KERNEL_OF_STAGE[stage]


TRACE_STEPS = False
MIN_WARM_S = 0.4  # the W warm-up steps are extended to at least this long (SM clocks ramp up from idle)
STEP_TRACE = []  # --trace-steps: per-step event times of every timed region (diagnostics, adds one event per step)


def warm_up(fn, warmup, world, min_warm_s):
    """W warm-up steps, extended to at least min_warm_s seconds (the SM clocks ramp up from idle).  The number of extra
    steps is derived from the all-reduced (max) time of the first W, so every rank runs the SAME number of steps -- a
    per-rank time-based loop would desynchronise the collectives inside fn."""
    t0 = time.perf_counter()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    per_step = max(elapsed / max(warmup, 1), 1e-4)
    extra = 0 if elapsed >= min_warm_s else min(20000, int((min_warm_s - elapsed) / per_step) + 1)
    for i in range(extra):
        fn()
        if (i + 1) % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()


def event_time_ms(fn, steps, warmup, world, min_warm_s=0.0):
    """W untimed warm-ups (see warm_up), then exactly K steps bracketed by barrier + synchronize; max over ranks."""
    if warmup > 0 or min_warm_s > 0:
        warm_up(fn, warmup, world, min_warm_s)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = []
    for _ in range(steps):
        fn()
        if TRACE_STEPS:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
    e1.record()
    torch.cuda.synchronize()
    if TRACE_STEPS:
        prev, row = e0, []
        for ev in marks:
            row.append(round(prev.elapsed_time(ev), 3))
            prev = ev
        STEP_TRACE.append(row)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def cpu_port_step_seconds(sc_c, cam_c, settings, dL_c, repeats):
    from oracle import cpu_oracle as co
    args = [t.numpy() for t in (sc_c.means3D, sc_c.scales, sc_c.rotations, sc_c.opacities, sc_c.shs)]
    cargs = [t.numpy() for t in (cam_c.viewmatrix, cam_c.projmatrix, cam_c.inv_viewprojmatrix, cam_c.campos, cam_c.bg)]
    co.lib()
    t0 = time.perf_counter()
    for _ in range(repeats):
        o = co.Oracle(settings, *args, sc_c.sh_degree, *cargs, cam_c.tanfovx, cam_c.tanfovy, cam_c.image_width,
                      cam_c.image_height)
        o.backward(dL_c.numpy())
        o.close()
    return (time.perf_counter() - t0) / repeats


def own_sort_bytes(R, T, slab):
    """what the tile-bucket sort of this library moves per launch: 8 B record in, 12 B (key + id) out and, for the
    depth-resorting modes, the slab records (means2D 8 + conic/opacity 16 + inverse covariance 48 + colour 12 B gathered,
    64 B geometry + 16 B {r,g,b,id} written) -- next to SURVEY 8d's radix-sort formula"""
    return R * (8 + 12 + (84 + 80 if slab else 0)) + 16 * T


def bands_mode_of(workload, sharding, world):
    if world <= 1:
        return False
    if sharding == "auto":
        return workload in BAND_DEFAULT
    return sharding == "bands"


def measure(a, workload, world, rank, dev, steps, warmup, full):
    """one workload on the current process group -> dict (rank 0) / None.  full: e2e_full_upload, roofline details,
    cpu_baseline; extras run with full=False."""
    scene_name, overrides, desc = WORKLOADS[workload]
    settings = S.default_settings_dict(**overrides)
    cid, P, W, H = S.CONFIGS[scene_name]
    if a.points and full:
        P = a.points
        desc += f" [--points {P}]"
    pixels = W * H
    bands_mode = bands_mode_of(workload, a.sharding, world)
    ref_single = a.impl == "reference" and bands_mode  # the reference cannot shard a frame: one GPU renders all of it
    eff_world = 1 if ref_single else world
    use_dist = eff_world > 1

    ref = None
    if a.impl == "reference":
        from oracle import ref_api as ref

    # ---- workload: same Gaussians on every rank; one camera per rank (views) or one camera for all (bands) ----
    sc_c, cam0_c = S.make_config(scene_name, P=P)
    per_rank_view = use_dist and not bands_mode
    cam_c, _, _ = S.make_camera(W, H, yaw=0.08 * rank if per_rank_view else 0.0)
    dL_c = S.make_upstream_grad(W, H, 2000 + cid + (rank if per_rank_view else 0))
    import stp_sharding as SH
    grid_y = (H + 15) // 16
    grid_x = (W + 15) // 16
    bands = SH.equal_bands(grid_y, world) if bands_mode and not ref_single else None
    sc, cam = S.to_device(sc_c, dev), S.to_device(cam_c, dev)
    dL = dL_c.to(dev)
    e = torch.empty(0, device=dev)
    M = sc.shs.shape[1]
    band_box = {"band": bands[rank] if bands else None}

    sync_group = dist.group.WORLD if (use_dist and a.impl == "ours") else None
    if a.impl == "ours":
        from diff_gaussian_rasterization import (ExtendedSettings, GaussianRasterizationSettings, GaussianRasterizer, _C)

        def fwd(c, dbg=2):
            return _C.rasterize_gaussians(c.bg, sc.means3D, e, sc.opacities, sc.scales, sc.rotations, 1.0, e,
                                          c.viewmatrix, c.projmatrix, c.inv_viewprojmatrix, c.tanfovx, c.tanfovy, H, W,
                                          sc.shs, sc.sh_degree, c.campos, False, settings, False, dbg,
                                          tile_band=band_box["band"], async_forward=a.async_forward)

        def bwd(c, out, g, dbg=2):
            return _C.rasterize_gaussians_backward(c.bg, sc.means3D, out[2], sc.opacities, e, sc.scales, sc.rotations,
                                                   1.0, e, c.viewmatrix, c.projmatrix, c.inv_viewprojmatrix, c.tanfovx,
                                                   c.tanfovy, out[1], g, sc.shs, sc.sh_degree, c.campos, out[3], out[0],
                                                   out[4], out[5], settings, dbg, want_param_slab=True,
                                                   tile_band=band_box["band"], sync_group=sync_group)
    else:
        def fwd(c, dbg=False):
            return ref.forward(sc, c, settings)

        def bwd(c, out, g, dbg=False):
            grads = ref.backward(sc, c, settings, out, g)
            return grads, grads  # the reference returns 8 separate tensors

    state = {}
    full_sort_ref = a.impl == "reference" and settings["sort_settings"]["sort_mode"] == 1  # the reference has no backward

    if bands is not None and a.bands == "balanced":
        # bands of (almost) equal instance count instead of equal height: per-row instance counts of one warm-up frame
        # rendered with equal bands, summed over the ranks ("balanced by instance count ... from the
        # previous frame")
        out0 = fwd(cam, dbg=0)
        rows = SH.row_weights_from_ranges(_C.view_image(out0[5], W, H)["ranges"], grid_x, grid_y).to(torch.float32)
        dist.all_reduce(rows)
        bands = SH.balanced_bands(rows.cpu().tolist(), world)
        band_box["band"] = bands[rank]
        del out0

    gather_stream = torch.cuda.Stream(device=dev) if bands is not None else None

    def gather_bands(img):
        """the ONE forward exchange of tile sharding, on a side stream: it overlaps the render-backward stage (nothing
        on this rank's backward path needs the other ranks' bands); the step waits for it at its end"""
        main = torch.cuda.current_stream(dev)
        gather_stream.wait_stream(main)
        with torch.cuda.stream(gather_stream):
            full = SH.gather_image_bands(img, bands)
        img.record_stream(gather_stream)
        return full

    def step_resident():
        out = fwd(cam)
        if bands is not None:
            state["image"] = gather_bands(out[1])
        if full_sort_ref:
            state["out"] = out
            return
        grads, slab = bwd(cam, out, dL)
        if use_dist and a.impl != "ours":
            for t in (slab[3], slab[5], slab[2], slab[6], slab[7]):
                dist.all_reduce(t)
        if bands is not None:
            torch.cuda.current_stream(dev).wait_stream(gather_stream)
        state["out"], state["grads"] = out, grads

    # ---- e2e: public API, per-step host inputs -------------------------------------------------------------
    pin = lambda t: t.clone().pin_memory()  # noqa: E731
    h_cam = [pin(t) for t in (cam_c.viewmatrix, cam_c.projmatrix, cam_c.inv_viewprojmatrix, cam_c.campos, cam_c.bg)]
    h_dL = pin(dL_c)
    h_img = torch.empty(3, H, W, dtype=torch.float32).pin_memory()
    # tile bands: a rank only consumes the upstream gradient of ITS band and only produces its band of the image, so
    # that is what it moves over PCIe (pixel rows [r0, r1)); the byte counts are totals over all ranks
    if bands is not None:
        prow = (min(bands[rank][0] * 16, H), min(bands[rank][1] * 16, H))
        h2d = sum(t.numel() * 4 for t in h_cam) * world + h_dL.numel() * 4
        d2h = h_img.numel() * 4
    else:
        prow = (0, H)
        h2d = (sum(t.numel() * 4 for t in h_cam) + h_dL.numel() * 4) * (eff_world if per_rank_view else 1)
        d2h = h_img.numel() * 4 * (eff_world if per_rank_view else 1)
    leaves = [t.clone().requires_grad_(True) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations)]
    means2D = torch.zeros_like(sc.means3D, requires_grad=True)

    # copies overlap compute on a side stream (both arms use the same schedule): the upstream-gradient upload runs
    # under the forward pass, the image download under the backward pass; the step ends when both streams are done
    copy_stream = torch.cuda.Stream(device=dev)
    g_dev = torch.empty_like(dL)
    if bands is not None:  # strided host<->device copies would be staged through pageable memory by torch
        h_dL_band = h_dL[:, prow[0]:prow[1]].contiguous().pin_memory()
        h_img_band = torch.empty(3, prow[1] - prow[0], W, dtype=torch.float32).pin_memory()
        g_band = torch.empty(3, prow[1] - prow[0], W, dtype=torch.float32, device=dev)
    ev_g, ev_f = torch.cuda.Event(), torch.cuda.Event()
    ext_settings = ExtendedSettings.from_dict(settings) if a.impl == "ours" else None

    def step_e2e():
        main = torch.cuda.current_stream(dev)
        vm, pm, iv, cp, bg = [t.to(dev, non_blocking=True) for t in h_cam]
        copy_stream.wait_stream(main)  # g_dev / h_img of the previous step are no longer in use
        with torch.cuda.stream(copy_stream):
            if bands is None:
                g_dev.copy_(h_dL, non_blocking=True)
            else:  # contiguous pinned band -> contiguous device staging -> the band's rows of the gradient image
                g_band.copy_(h_dL_band, non_blocking=True)
                g_dev[:, prow[0]:prow[1]].copy_(g_band)
            ev_g.record(copy_stream)
        for t in leaves + [means2D]:
            t.grad = None
        m3, op, sh, scl, rot = leaves
        color_full = None
        if a.impl == "ours":
            rs = GaussianRasterizationSettings(H, W, cam_c.tanfovx, cam_c.tanfovy, bg, 1.0, vm, pm, iv, sc.sh_degree, cp,
                                               False, ext_settings, False, False)
            color, radii = GaussianRasterizer(rs, tile_band=band_box["band"], sync_group=sync_group,
                                              async_forward=a.async_forward)(
                m3, means2D, op, shs=sh, scales=scl, rotations=rot)
            if bands is not None:
                color_full = gather_bands(color.detach())
        else:
            c = cam._replace(viewmatrix=vm, projmatrix=pm, inv_viewprojmatrix=iv, campos=cp, bg=bg)
            out = ref.forward(sc, c, settings)
            color = out[1]
        img_out = color.detach()  # (bands: this rank's rows; the gathered frame stays on the GPUs)
        ev_f.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_f)
            if bands is None:
                h_img.copy_(img_out, non_blocking=True)
            else:
                h_img_band.copy_(img_out[:, prow[0]:prow[1]].contiguous(), non_blocking=True)
        img_out.record_stream(copy_stream)
        main.wait_event(ev_g)
        if full_sort_ref:
            main.wait_stream(copy_stream)
            return
        if a.impl == "ours":
            color.backward(g_dev)  # gradients come back already summed over the ranks (sync_group)
        else:
            grads = ref.backward(sc, c, settings, out, g_dev)
            if use_dist:
                for t in (grads[3], grads[5], grads[2], grads[6], grads[7]):
                    dist.all_reduce(t)
        main.wait_stream(copy_stream)
        if bands is not None:
            main.wait_stream(gather_stream)
            state["image_e2e"] = color_full

    # full-upload variant: every Gaussian parameter host->device, every parameter gradient device->host
    h_params = [pin(t) for t in (sc_c.means3D, sc_c.opacities, sc_c.shs, sc_c.scales, sc_c.rotations)] if full else []
    h_grads = [torch.empty_like(t).pin_memory() for t in h_params]
    h2d_full = h2d + sum(t.numel() * 4 for t in h_params)
    d2h_full = d2h + sum(t.numel() * 4 for t in h_grads)

    def step_e2e_full():
        nonlocal leaves
        leaves = [t.to(dev, non_blocking=True).requires_grad_(True) for t in h_params]
        step_e2e()
        if a.impl == "ours":
            for hg, t in zip(h_grads, leaves):
                hg.copy_(t.grad, non_blocking=True)

    # ---- timed regions --------------------------------------------------------------------------------------
    tw = eff_world
    sampler = ClockSampler(dev.index or 0)
    if rank == 0 and full:
        sampler.start()  # before the warm-up: the sampler's start-up cost stays outside the timed region
    launches0 = 0
    if a.impl == "ours":
        _C.timing_reset()
        warm_up(step_resident, warmup, tw, MIN_WARM_S)
        _C.timing_reset()
        launches0 = _C.kernel_launches()
    if rank == 0 and full:
        sampler.mark()  # only samples taken from here on (= during the timed region) are reported
    ms_total = event_time_ms(step_resident, steps, warmup if a.impl != "ours" else 0, tw,
                             MIN_WARM_S if a.impl != "ours" else 0.0)
    launches, stages = 0, {}
    if a.impl == "ours":
        launches = _C.kernel_launches() - launches0
        stages = _C.timing_summary()
    clocks = sampler.stop() if (rank == 0 and full) else None
    ms_step = ms_total / steps
    views = eff_world if per_rank_view else 1  # tile sharding: all ranks render ONE frame (strong scaling)
    value = views * pixels / (ms_step * 1e-3) / 1e6

    ms_e2e = event_time_ms(step_e2e, steps, warmup, tw) / steps
    e2e_value = views * pixels / (ms_e2e * 1e-3) / 1e6
    e2e_full = None
    if a.impl == "ours" and world == 1 and full:
        ms_full = event_time_ms(step_e2e_full, max(3, steps // 4), 3, tw) / max(3, steps // 4)
        e2e_full = {"value": pixels / (ms_full * 1e-3) / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d_full,
                    "d2h_bytes_per_step": d2h_full}
    if rank != 0:
        return None

    out = state["out"]
    R = int(out[0])
    V = int((out[2] > 0).sum().item())
    if bands is None and world == 1:
        sharding = "single GPU"
    elif ref_single:
        sharding = ("reference arm: ONE GPU renders the whole frame (the reference cannot shard a frame; SURVEY 8d keeps "
                    "the 1-GPU reference number as the baseline of the tile-sharded configurations)")
    elif bands is not None:
        sharding = (f"tile-row bands of one view ({a.bands}: {bands}); replicated Gaussians, NCCL all-gather of the image "
                    "bands + all-reduce of the 36 B/Gaussian screen-space gradient accumulator between the two backward "
                    "stages")
    else:
        sharding = ("views (one camera per rank, replicated Gaussians, NCCL all-reduce of parameter grads, overlapped "
                    "with the preprocess-backward stage)")
    res = {
        "value": value, "ms_per_step": ms_step, "steps": steps,
        "scaling": "strong" if (bands_mode or (world == 1 and workload in BAND_DEFAULT and a.sharding != "views")) else "weak",
        "config": {"workload": desc, "P": P, "W": W, "H": H, "visible": V, "num_rendered": R, "sharding": sharding,
                   "l2_policy": "inputs larger than L2 (236 B/Gaussian x P + instance lists >> 126 MB)" if P >= 10**6
                   else "small parity config; L2-resident"},
        "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e,
                "what": "GaussianRasterizer autograd API; camera + upstream-gradient image from pinned host memory, "
                        "rendered image read back (copies on a side stream, overlapping fwd / bwd; step ends when both "
                        "streams are done); Gaussian parameters (model state) resident"},
        "clocks": clocks,
    }
    if a.impl == "ours":
        res["config"]["forward"] = ("asynchronous (async_forward=True: num_rendered resolved lazily, no host<->device "
                                    "synchronisation in the forward call)" if a.async_forward else "synchronous")
    if bands is not None:
        res["config"]["num_rendered_is"] = "rank 0's band only"
    if full_sort_ref:
        res["config"]["note"] = ("reference arm is FORWARD ONLY: the reference has no PPX_FULL backward "
                                 "(backward.cu:733-736); ours is forward + backward")
    if a.impl == "ours":
        N, T = pixels, grid_x * grid_y
        s_flag = 0 if settings["sort_settings"]["sort_mode"] == 0 else 1
        c_flag = 1 if (s_flag or settings["sort_settings"]["sort_order"] in (2, 3)) else 0
        ab = algorithmic_bytes(P, V, R, N, T, M, s_flag, c_flag)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        measured = prof.get(workload, {}) if world == 1 else {}
        per_stage = {}
        for k, v in stages.items():
            if k not in ab:
                continue
            row = {"ms": v[0], "GBps": ab[k] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None, "bytes": ab[k]}
            if k == "Sort":
                # SURVEY 8d charges the reference's six radix passes; the tile-bucket sort moves far fewer bytes, so the
                # formula alone would show a "fraction" above 1 on bytes that are never moved: both are printed
                row["bytes_is"] = "SURVEY 8d formula (reference radix sort: 6 passes x 24 B/instance)"
                row["own_bytes"] = own_sort_bytes(R, T, bool(s_flag))
                row["own_GBps"] = row["own_bytes"] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None
            if k in measured:
                row["dram_bytes_ncu"] = measured[k]
            per_stage[k] = row
        dom = max(per_stage, key=lambda k: per_stage[k]["ms"])
        res["stages_ms"] = {k: round(v["ms"], 4) for k, v in per_stage.items()}
        res["gpu_launches"] = int(launches)
        if full:
            res["roofline"] = {
                "bound": "hbm", "kernel": kernel_name(dom, settings), "stage": dom,
                "achieved": per_stage[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": per_stage[dom]["GBps"] / peak, "traffic": measured.get(dom),
                "traffic_source": (prof.get("_source", "committed ncu capture under profiles/") if measured.get(dom)
                                   else "none for this workload / GPU count"),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "algorithmic_bytes": ab[dom], "kernel_ms": per_stage[dom]["ms"],
                # what actually bounds the dominant kernel (ncu smsp__issue_active of the committed capture,
                # profiles/*_ncu_full_summary_*.csv): the render kernels are instruction-issue bound, not HBM bound
                "issue_active_pct_ncu": prof.get(workload + "_issue_active_pct", {}).get(dom),
                "stages": per_stage,
                "whole_step": {"bytes": sum(ab.values()), "GBps": sum(ab.values()) / (ms_step * 1e-3) / 1e9,
                               "frac": sum(ab.values()) / (ms_step * 1e-3) / 1e9 / peak}}
        if e2e_full:
            res["e2e_full_upload"] = e2e_full
        if world == 1 and full and not a.no_cpu_baseline and settings["sort_settings"]["sort_mode"] != 1:
            reps = 2 if P <= 10**6 else 1
            sc_b, cam_b, dL_b, sample = sc_c, cam_c, dL_c, f"{reps} full step(s) of the workload"
            if P > 10**6:  # bound the CPU work: same camera, first 1M Gaussians of the cloud
                sc_b = S.Scene(*[t[:10**6].contiguous() if isinstance(t, torch.Tensor) else t for t in sc_c])
                sample = "1 step over the first 1M Gaussians of the cloud at full resolution"
            sec = cpu_port_step_seconds(sc_b, cam_b, settings, dL_b, reps)
            res["cpu_baseline"] = {"value": pixels / sec / 1e6, "unit": "Mpixels/s", "cores": os.cpu_count(),
                                   "kind": "port", "sample": sample + " (oracle/stp_oracle.c, OpenMP)",
                                   "ms_per_step": sec * 1e3}
    else:
        res["cpu_baseline"] = {"value": e2e_value, "unit": "Mpixels/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "unmodified reference CUDA build (oracle/_ref, sm_100) on the same GPU, full "
                                         "workload -- the reference ships no CPU path; host cores only launch kernels"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3b", choices=sorted(WORKLOADS))
    ap.add_argument("--sharding", default="auto", choices=["auto", "bands", "views"],
                    help="multi-GPU split: tile-row bands of one frame / one view per rank (auto: per workload)")
    ap.add_argument("--bands", default="balanced", choices=["balanced", "equal"],
                    help="tile-row bands of equal instance count (from a warm-up frame) or of equal height")
    ap.add_argument("--points", type=int, default=0, help="override the number of Gaussians of the workload's scene")
    ap.add_argument("--async-forward", dest="async_forward", action="store_true",
                    help="ours: GaussianRasterizer(async_forward=True) -- no host<->device synchronisation inside the forward "
                         "call, num_rendered resolved lazily (measured on B200: no difference, the synchronous call's "
                         "bubble is already hidden by the speculative arena request; profiles/r02_async_forward.txt)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--trace-steps", action="store_true", help="diagnostics: per-step times of every timed region")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    global TRACE_STEPS
    TRACE_STEPS = a.trace_steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    scene_name, overrides, desc = WORKLOADS[a.workload]
    settings = S.default_settings_dict(**overrides)
    cid, P, W, H = S.CONFIGS[scene_name]

    use_ref_gpu = False
    if a.impl == "reference":  # the only arm that touches oracle/ (besides the cpu_baseline leg)
        from oracle import ref_api as ref
        use_ref_gpu = ref.available() and torch.cuda.is_available()

    if a.impl == "reference" and not use_ref_gpu:
        # no reference CUDA build travelled with the repo: the C oracle port on the host cores
        if rank != 0:
            return
        sc_c, cam_c = S.make_config(scene_name, P=a.points or P)
        dL_c = S.make_upstream_grad(W, H, 2000 + cid)
        for _ in range(1):
            cpu_port_step_seconds(sc_c, cam_c, settings, dL_c, 1)
        sec = cpu_port_step_seconds(sc_c, cam_c, settings, dL_c, max(1, min(a.steps, 3)))
        v = W * H / sec / 1e6
        print(json.dumps({"impl": "reference", "metric": "Mpixels/s fwd+bwd", "value": v, "unit": "Mpixels/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": desc},
                          "cpu_baseline": {"value": v, "unit": "Mpixels/s", "cores": os.cpu_count(), "kind": "port",
                                           "sample": "full workload, oracle/stp_oracle.c with OpenMP"},
                          "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    main_bands = bands_mode_of(a.workload, a.sharding, world)
    # extras: short runs of the other BASELINE.json configurations, reported under "extra" (never the headline)
    extras = []
    if not a.no_extras and not a.points:
        extras = [w for w in (("C2", "C4") if world == 1 else ("C4",)) if w != a.workload]
        if world > 1:
            extras = [w for w in extras if bands_mode_of(w, a.sharding, world) == main_bands]
    ref_rank0_only = a.impl == "reference" and main_bands
    if ref_rank0_only and rank != 0:
        return  # the reference renders the whole frame on one GPU; nothing to do for the other ranks
    if world > 1 and not ref_rank0_only:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    res = measure(a, a.workload, world, rank, dev, a.steps, a.warmup, True)
    extra = {}
    for w in extras:
        torch.cuda.empty_cache()
        ws = max(3, min(a.steps, 5 if w == "C4" else 20))
        if a.impl == "reference" and w == "C4":
            ws = 3  # the reference's full per-pixel sort takes seconds per frame
        r = measure(a, w, world, rank, dev, ws, 3, False)
        if r is not None:
            r.pop("clocks", None)
            extra[w] = r
    if rank != 0:
        if dist.is_initialized():
            dist.destroy_process_group()
        return

    line = {
        "metric": "Mpixels/s fwd+bwd", "value": res["value"], "unit": "Mpixels/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": res["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": res["config"], "e2e": res["e2e"], "clocks": res["clocks"],
    }
    for k in ("roofline", "gpu_launches", "e2e_full_upload", "cpu_baseline"):
        if k in res:
            line[k] = res[k]
    if a.impl != "ours":
        line["impl"] = "reference"
    if extra:
        line["extra"] = extra
    if TRACE_STEPS:
        line["step_trace_ms"] = STEP_TRACE
    print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
