#!/usr/bin/env python
"""bench.py -- stabilized frames/s of the per-frame hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one video frame through the hot path: the PWC-Net custom ops (Correlation + Warp at every
decoder level, both flow directions), the flow upsample, and the full stabilization of the frame
(fused warp x5 + adaptive combination + consistency weight, pyramid, screened-Poisson solve, 8-bit output).
PWC-Net's dense convolutions stay in the reference's ORT graph and are out of scope (north_star); their
activations are represented by synthetic device-resident feature / flow tensors of the right shapes.

Workloads (BASELINE.json configs):
  1080p-light  configs[1]: 1080p frames, pwcnet-light level shapes at FLOWDOWNSCALE=2, flow 960x540 upsampled  (default)
  4k-dense     configs[2]: 4K frames, pwcnet dense level shapes at full resolution
  4k-stab      configs[3]: precomputed-flow stabilization only at 4K

value  = frames/s with every input already resident in HBM (CUDA events on the launching stream).
e2e    = frames/s through the public pipeline object with HOST frame buffers: per step two RGBA8 frames are
         uploaded from pinned memory and the stabilized RGBA8 frame is read back, inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".
Multi-GPU: independent streams, one per GPU (replicas; the recurrence is sequential per stream): value is the
aggregate over ranks divided by the slowest rank's time; no data-path collective exists.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

LIGHT_CORR = [(196, 9, 15), (128, 18, 30), (96, 36, 60), (64, 72, 120)]
LIGHT_WARP = [(128, 18, 30), (96, 36, 60), (64, 72, 120)]
DENSE_CORR = [(196, 34, 60), (128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]
DENSE_WARP = [(128, 68, 120), (96, 136, 240), (64, 272, 480), (32, 544, 960)]

WORKLOADS = {
    "1080p-light": dict(W=1920, H=1080, flowW=960, flowH=540, corr=LIGHT_CORR, warp=LIGHT_WARP,
                        desc="1080p frames, pwcnet-light custom-op level shapes at FLOWDOWNSCALE=2 (both directions), "
                             "flow 960x540 upsampled, full stabilization (BASELINE configs[1])"),
    "4k-dense": dict(W=3840, H=2160, flowW=3840, flowH=2160, corr=DENSE_CORR, warp=DENSE_WARP,
                     desc="4K frames, pwcnet dense custom-op level shapes at full resolution (both directions), "
                          "full stabilization (BASELINE configs[2])"),
    "4k-stab": dict(W=3840, H=2160, flowW=3840, flowH=2160, corr=[], warp=[],
                    desc="precomputed-flow stabilization only at 4K (BASELINE configs[3])"),
    # file mode of configs[3] (-f <flowdir>): the e2e arm also reads the two .flo files of every frame from disk
    # (page cache) through vsc_stabilizer_step_flow_files + prefetch of the next frame's pair
    "4k-stab-files": dict(W=3840, H=2160, flowW=3840, flowH=2160, corr=[], warp=[], files=True,
                          desc="precomputed-flow stabilization at 4K, flows ingested from .flo files in the e2e arm "
                               "(BASELINE configs[3], file mode)"),
    "1080p-stab-files": dict(W=1920, H=1080, flowW=1920, flowH=1080, corr=[], warp=[], files=True,
                             desc="precomputed-flow stabilization at 1080p, flows ingested from .flo files in the "
                                  "e2e arm"),
}
NFRAMES = 8  # distinct synthetic frame pairs, cycled


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.rows = []
        self.proc = None
        self.gpu_id = gpu_id

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.gpu_id], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ inputs
def make_host_frames(W, H, pin):
    import synth
    import torch

    o8, p8 = synth.frames(W, H, NFRAMES, seed=1234)
    ho = [torch.from_numpy(np.ascontiguousarray(x)) for x in o8]
    hp = [torch.from_numpy(np.ascontiguousarray(x)) for x in p8]
    if pin:
        ho = [x.pin_memory() for x in ho]
        hp = [x.pin_memory() for x in hp]
    return ho, hp


def make_op_tensors(wl, dev):
    """two independent sets (one per flow direction) of feature / flow tensors per decoder level"""
    import synth
    import torch

    sets = []
    for d in range(2):
        corr = [(torch.from_numpy(synth.features(1, C, h, w, 10 * d + i)).to(dev),
                 torch.from_numpy(synth.features(1, C, h, w, 10 * d + i + 5)).to(dev),
                 torch.empty((1, 9, 9, h, w), device=dev)) for i, (C, h, w) in enumerate(wl["corr"])]
        warp = [(torch.from_numpy(synth.features(1, C, h, w, 20 * d + i)).to(dev),
                 torch.from_numpy(synth.op_flow_smooth(1, h, w, 30 * d + i)).to(dev),
                 torch.empty((1, C, h, w), device=dev)) for i, (C, h, w) in enumerate(wl["warp"])]
        sets.append((corr, warp))
    return sets


def op_bytes(wl):
    b = 0
    for C, h, w in wl["corr"]:
        b += 4 * h * w * (2 * C + 81)
    for C, h, w in wl["warp"]:
        b += 4 * h * w * (2 * C + 2)
    return 2 * b  # both directions


def run_ops(V, sets):
    for corr, warp in sets:
        for a, b, o in corr:
            V.correlation(a, b, out=o)
        for x, f, o in warp:
            V.warp(x, f, out=o)


# ------------------------------------------------------------------------------------------------ ranks
def reduce_over_ranks(times_ms, counts, world, device):
    """replica scaling: every rank runs an independent stream; the job time is the MAX over ranks, counters add.
    (torch.distributed is used for this bookkeeping and the barriers only -- the data path has no collective.)"""
    if world <= 1:
        return list(times_ms), list(counts)
    import torch
    import torch.distributed as dist

    tt = torch.tensor(list(times_ms), device=device, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    cc = torch.tensor(list(counts), device=device, dtype=torch.int64)
    dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    return [float(x) for x in tt], [int(x) for x in cc]


def aggregate_fps(world, steps, ms):
    """whole-job throughput: all ranks' frames divided by the slowest rank's time"""
    return world * steps / (ms * 1e-3)


def stream_frame_index(t):
    """window (prev, cur, next) of step t over the cycled synthetic frames"""
    return (t - 1) % NFRAMES, t % NFRAMES, (t + 1) % NFRAMES


# ------------------------------------------------------------------------------------------------ config
def job_config(workload, world):
    """the `config` object of the JSON line -- built by BOTH arms from the same arguments, so that the two lines
    describe the same job key for key"""
    wl = WORKLOADS[workload]
    W, H, fw, fh = wl["W"], wl["H"], wl["flowW"], wl["flowH"]
    return {"workload": workload + ": " + wl["desc"], "resolution": f"{W}x{H}", "flow_resolution": f"{fw}x{fh}",
            "numIter": 150, "pyramidLevels": 2, "streams": world,
            "parallelism": f"replicas x{world} (independent video streams, no collective)",
            "l2": (f"inputs larger than L2: {NFRAMES} frame pairs cycled, per-step working set "
                   f"{(12 * 4 + 6 * 4) * W * H / 1e6:.0f} MB of solver state > 126 MB L2" if W * H * 72 > 126e6
                   else f"{NFRAMES} frame pairs cycled; solver state {72 * W * H / 1e6:.0f} MB per sweep"),
            "custom_op_bytes_per_step": op_bytes(wl),
            "out_of_scope": "PWC-Net convolutions (ORT graph): synthetic device-resident activations"}


# ------------------------------------------------------------------------------------------------ host placement
def bind_near_gpu(local_rank):
    """Pin this process to the CPU cores of the GPU's NUMA node BEFORE any pinned buffer is allocated.

    torchrun starts every rank with the same affinity mask; pinned staging memory is then first-touched on whatever
    node the rank happens to run on, and the per-frame 25 MB (1080p) of H2D + D2H of ranks whose GPU hangs off the
    other socket crosses the inter-socket link.  Binding by GPU locality keeps every stream's host traffic local.
    Returns a short description for the JSON line; never fails (a restricted cpuset just keeps the old mask)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            tok = vis.split(",")[local_rank].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(tok))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        if not want:
            return {"numa_node": node, "bound": False, "why": "GPU-local cores are outside this process's cpuset",
                    "allowed": len(allowed)}
        os.sched_setaffinity(0, want)
        return {"numa_node": node, "bound": True, "cores": len(want)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------------ our arm
_frames_cache = {}


def host_frames_cached(W, H):
    if (W, H) not in _frames_cache:
        _frames_cache.clear()            # one resolution at a time (pinned 4K sets are 0.5 GB)
        _frames_cache[(W, H)] = make_host_frames(W, H, pin=True)
    return _frames_cache[(W, H)]


def run_workload(args, name, K, Wm, rank, world, dev, sustained_s=0.0, sample_clocks=True):
    """one workload on this rank's GPU: device-resident arm, end-to-end arm, kernel rooflines.
    -> dict of per-rank raw numbers (times not yet reduced over ranks)"""
    import ctypes as C

    import torch
    import torch.distributed as dist

    import synth
    import vsc_b200 as V

    local = dev.index
    wl = WORKLOADS[name]
    W, H, fw, fh = wl["W"], wl["H"], wl["flowW"], wl["flowH"]
    peaks, peak_src = measured_peaks()
    ho, hp = host_frames_cached(W, H)
    files = bool(wl.get("files"))
    flow_c = 2 if files else 3    # .flo files hold (u, v) pairs
    flf, flb = synth.flows(fw, fh, flow_c)
    d_flf, d_flb = torch.from_numpy(flf).to(dev), torch.from_numpy(flb).to(dev)
    sets = make_op_tensors(wl, dev)
    hpar = V.HyperParams()
    V.check(V.lib().vsc_set_solver_mode(args.solver_mode))
    V.check(V.lib().vsc_set_stage_a_mode(args.stage_a_mode))

    # ---------------- device-resident arm: every input already in HBM ----------------
    d_o = [V.image_to_gpu(x.to(dev)) for x in ho]
    d_p = [V.image_to_gpu(x.to(dev)) for x in hp]
    last = d_p[2].clone()
    cons = torch.empty_like(last)
    ws = torch.empty(int(V.lib().vsc_frame_stabilize_workspace_bytes(W, H, hpar.pyramidLevels)), device=dev,
                     dtype=torch.uint8)
    lowres = (fw, fh) != (W, H)
    upf = torch.empty((H, W, 3), device=dev) if lowres else d_flf
    upb = torch.empty((H, W, 3), device=dev) if lowres else d_flb
    L = V.lib()

    def dptr(t):
        return C.c_void_p(t.data_ptr())

    def stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    out8 = torch.empty((H, W, 4), device=dev, dtype=torch.uint8)

    # Two streams (default; --no-overlap-ops = one stream): the custom ops of a frame (its flow network) run on a second
    # stream.  The flow of frame t does not depend on the stabilized output of frame t-1, so ops(t+1) may run while frame
    # t is being stabilized (the solver passes leave issue slots and, at level 1, half of every SM free): stabilization(t)
    # waits for ops(t), and ops(t+1) waits for stabilization(t-1) -- one frame of look-ahead, no more.  Every op and
    # every stabilization call goes through the same C-ABI entry points as before; only the stream argument differs.
    overlap = bool(wl["corr"]) and not args.no_overlap_ops
    ops_stream = torch.cuda.Stream(device=dev) if overlap else None
    stab_done = []

    def step_resident(t, two_streams=True):
        nonlocal last, cons
        if ops_stream is not None and two_streams:
            cur = torch.cuda.current_stream()
            if len(stab_done) >= 2:
                ops_stream.wait_event(stab_done[-2])
            with torch.cuda.stream(ops_stream):
                run_ops(V, sets)
                ev = torch.cuda.Event()
                ev.record(ops_stream)
            cur.wait_event(ev)
        else:
            run_ops(V, sets)
        if lowres:
            V.check(L.vsc_bilinear(dptr(d_flf), fw, fh, 3, dptr(upf), W, H, 3, stream()))
            V.check(L.vsc_bilinear(dptr(d_flb), fw, fh, 3, dptr(upb), W, H, 3, stream()))
        i0, i1, i2 = stream_frame_index(t)
        V.frame_stabilize(d_o[i0], d_o[i1], d_o[i2], d_p[i0], d_p[i1], d_p[i2], last, upf, upb, hpar, out=cons,
                          workspace=ws)   # flow channel count is taken from the flow tensors
        V.check(L.vsc_f32x3_to_rgba8(dptr(cons), dptr(out8), W, H, stream()))
        last, cons = cons, last
        if ops_stream is not None and two_streams:
            ev2 = torch.cuda.Event()
            ev2.record(torch.cuda.current_stream())
            stab_done.append(ev2)
            del stab_done[:-2]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(Wm):
        step_resident(1 + t)
    barrier()
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        gpu_id = str(local)
    clocks = ClockSampler(gpu_id) if sample_clocks else None
    if clocks:
        clocks.start()
        time.sleep(0.25)
    n0 = V.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for t in range(K):
        step_resident(1 + Wm + t)
    e1.record()
    barrier()
    launches = V.launch_count() - n0
    ms_res = e0.elapsed_time(e1)
    # the same K steps with everything on ONE stream, for the record (`one_stream` in the JSON line)
    ms_one = None
    if ops_stream is not None:
        torch.cuda.current_stream().wait_stream(ops_stream)
        step_resident(1 + Wm + K, two_streams=False)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for t in range(K):
            step_resident(2 + Wm + K + t, two_streams=False)
        f1.record()
        barrier()
        ms_one = f0.elapsed_time(f1)

    # ---------------- end-to-end arm: host frame buffers through the pipeline object ----------------
    st = V.Stabilizer(W, H, flow_c)
    outs = [V.pinned_empty((H, W, 4)) for _ in range(2)]
    ext = torch.cuda.ExternalStream(st.compute_stream, device=dev)
    flow_dir, NFLO = None, 3
    if files:
        import shutil
        import tempfile

        # frames 1..NFLO cycle: forward file of frame i is frame_(i+1).flo, backward file frame_(i)_bwd.flo
        flow_dir = tempfile.mkdtemp(prefix=f"vsc_flo_r{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        hdr = b"PIEH" + np.array([fw, fh], np.int32).tobytes()
        for i in range(1, NFLO + 1):
            for path, arr in ((V.flo_frame_path(flow_dir, i + 1), flf), (V.flo_frame_path(flow_dir, i, True), flb)):
                with open(path, "wb") as f:
                    f.write(hdr)
                    f.write(np.roll(arr, i, axis=1).tobytes())

    # flow network input (FlowModel::run's frame path, on the device): the three window frames at the network's
    # input size (frame size / FLOWDOWNSCALE rounded up to a multiple of 64, convert_onnx.py:17-20)
    net = None
    if wl["corr"]:
        net = ((fw + 63) // 64 * 64, (fh + 63) // 64 * 64)
        net_in = [torch.empty((net[1], net[0], 4), device=dev, dtype=torch.uint8) for _ in range(3)]

    e2e_done = []

    def step_e2e(t):
        if net:
            for i in range(3):
                st.flow_input(i, net[0], net[1], net_in[i])
        if files:
            st.step_flow_files(flow_dir, 1 + t % NFLO, outs[t & 1])
            st.prefetch_flow_files(flow_dir, 1 + (t + 1) % NFLO)
        else:
            if ops_stream is not None:   # ops(t) on the second stream; see step_resident
                if len(e2e_done) >= 2:
                    ops_stream.wait_event(e2e_done[-2])
                with torch.cuda.stream(ops_stream):
                    run_ops(V, sets)
                    ev = torch.cuda.Event()
                    ev.record(ops_stream)
                ext.wait_event(ev)
            else:
                with torch.cuda.stream(ext):  # custom ops share the pipeline's compute stream
                    run_ops(V, sets)
            st.step(d_flf, d_flb, outs[t & 1])
            if ops_stream is not None:
                ev2 = torch.cuda.Event()
                ev2.record(ext)
                e2e_done.append(ev2)
                del e2e_done[:-2]
        st.push_frame(ho[(t + 2) % NFRAMES], hp[(t + 2) % NFRAMES])

    for t in range(3):
        st.push_frame(ho[t], hp[t])
    for t in range(Wm):
        step_e2e(1 + t)
    st.sync()
    barrier()
    t0 = time.perf_counter()
    for t in range(K):
        step_e2e(1 + Wm + t)
    st.sync()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms_e2e = (t1 - t0) * 1e3
    clk = clocks.stop(t_host0, t1) if clocks else None

    # pinned-copy bandwidth of this rank's host<->GPU path, alone (explains e2e when ranks share a root complex)
    probe = torch.empty((H, W, 4), device=dev, dtype=torch.uint8)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe.copy_(ho[0], non_blocking=True)
    torch.cuda.synchronize()
    ea.record()
    for i in range(4):
        probe.copy_(ho[i % NFRAMES], non_blocking=True)
    eb.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * W * H * 4 / (ea.elapsed_time(eb) * 1e-3) / 1e9
    del probe

    # ---------------- sustained: seconds of back-to-back frames (SM clock settles below boost) ----------------
    sustained = None
    if sustained_s > 0:
        sclk = ClockSampler(gpu_id)
        sclk.start()
        time.sleep(0.2)
        chunk = max(8, int(0.25 / max(ms_res / K * 1e-3, 1e-5)))
        evs = [torch.cuda.Event(enable_timing=True)]
        ts0 = time.perf_counter()
        evs[0].record()
        frames_done, t = 0, 1 + Wm + K
        while True:
            for _ in range(chunk):
                step_resident(t)
                t += 1
            frames_done += chunk
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            evs.append(ev)
            ev.synchronize()
            if evs[0].elapsed_time(ev) >= sustained_s * 1e3:
                break
        ts1 = time.perf_counter()
        total_ms = evs[0].elapsed_time(evs[-1])
        last_ms = evs[-2].elapsed_time(evs[-1])
        c = sclk.stop(ts0, ts1)
        sustained = {"seconds": total_ms * 1e-3, "frames": frames_done, "value": frames_done / (total_ms * 1e-3),
                     "unit": "frames/s", "last_chunk_value": chunk / (last_ms * 1e-3),
                     "sm_mhz_median": c["sm_mhz"], "sm_max_mhz": c["sm_max_mhz"], "clock_samples": c.get("samples"),
                     "reasons": c["reasons"],
                     "note": "device-resident arm run back to back for this long on rank 0's GPU; CUDA events"}
    st.close()
    if flow_dir:
        shutil.rmtree(flow_dir, ignore_errors=True)

    # ---------------- roofline of the dominant kernel: the level-0 solver ----------------
    # marginal time of n sweeps = t(2n) - t(n), CUDA events on the launching stream.  In the default mode the
    # sweeps run as temporally blocked passes (solver_rolled_kernel<T>, T sweeps per launch); the unblocked
    # sweep kernel is timed beside it (mode 1).  Algorithmic bytes: 72 B/pixel/sweep.
    def time_solve(iters):
        x = d_p[1].clone()
        wsb = torch.empty(int(L.vsc_consist_solve_workspace_bytes(W, H)), device=dev, dtype=torch.uint8)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        V.check(L.vsc_consist_solve(dptr(d_p[1]), dptr(d_p[2]), dptr(d_o[1]), iters, C.c_float(0.15), C.c_float(0.15),
                                    dptr(x), W, H, dptr(wsb), C.c_size_t(wsb.numel()), stream()))
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    # the main pass: 8 sweeps per launch at every size since the quad-gather exchange ring (stab_solver.cu, plan_sweeps:
    # 150 sweeps = 18 x 8 + 6)
    T_main = 8
    force = args.solver_mode & ~0x3000 | (0x2000 if T_main == 10 else 0x1000)
    n = 16 * T_main
    L.vsc_set_solver_mode(force)
    time_solve(2 * T_main)
    pass_ms = (time_solve(2 * n) - time_solve(n)) / 16
    L.vsc_set_solver_mode(1)
    time_solve(8)
    sweep_ms = (time_solve(2 * n) - time_solve(n)) / n
    L.vsc_set_solver_mode(args.solver_mode)
    alg_bytes = T_main * 72.0 * W * H
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
    unblocked = 72.0 * W * H / (sweep_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get(name, {}).get(f"solver_stream{T_main}_dram_bytes_per_launch")
        traffic_src = tj.get("_source")
    except Exception:
        pass

    # the fused stage-A kernel (5 warps + adaptive combination + consistency weight in one pass) timed alone through
    # vsc_stage_a_fused, CUDA events over 16 launches cycling the frame pairs (each launch reads 7 images + 2 flows:
    # more than L2 at both resolutions).  Algorithmic bytes 132 B/pixel.
    f_full = [d_flf, d_flb] if (fw, fh) == (W, H) else [V.get_bilinear(d_flf, W, H), V.get_bilinear(d_flb, W, H)]
    sa_out = [torch.empty_like(d_p[0]) for _ in range(2)]

    def stage_a_once(i):
        a, b, c = i % NFRAMES, (i + 1) % NFRAMES, (i + 2) % NFRAMES
        V.check(L.vsc_stage_a_fused(dptr(d_o[a]), dptr(d_o[b]), dptr(d_o[c]), dptr(d_p[a]), dptr(d_p[b]), dptr(d_p[c]),
                                    dptr(last), dptr(f_full[0]), dptr(f_full[1]), flow_c, C.c_float(hpar.alpha),
                                    C.c_float(hpar.beta), C.c_float(hpar.gamma), None, dptr(sa_out[0]), dptr(sa_out[1]),
                                    W, H, stream()))
    for i in range(3):
        stage_a_once(i)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for i in range(16):
        stage_a_once(i)
    eb.record()
    torch.cuda.synchronize()
    stage_a_ms = ea.elapsed_time(eb) / 16
    stage_a_bytes = (84.0 + 24.0 + 4.0 * 2 * flow_c) * W * H   # 7 images read, 2 written, 2 flows of flow_c channels
    stage_a_gbs = stage_a_bytes / (stage_a_ms * 1e-3) / 1e9

    # the custom ops at the two largest level shapes of this workload, each timed alone (CUDA events, 20 launches
    # over two independent tensor sets): Warp 4(2C+2)HW bytes, Correlation 4(2C+81)HW bytes
    ops_roof = []
    if wl["corr"]:
        def time_op(fn, reps=20):
            for _ in range(3):
                fn(0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(reps):
                fn(i & 1)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        for li in sorted(range(len(wl["warp"])), key=lambda i: -wl["warp"][i][1] * wl["warp"][i][2])[:2]:
            Cc, h, w = wl["warp"][li]
            ms = time_op(lambda d: V.warp(sets[d][1][li][0], sets[d][1][li][1], out=sets[d][1][li][2]))
            nb = 4.0 * h * w * (2 * Cc + 2)
            ops_roof.append({"kernel": "custom::Warp", "shape": [Cc, h, w], "us_per_launch": ms * 1e3,
                             "achieved": nb / (ms * 1e-3) / 1e9, "frac": nb / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "algorithmic_bytes_per_launch": nb})
        for li in sorted(range(len(wl["corr"])), key=lambda i: -wl["corr"][i][1] * wl["corr"][i][2])[:2]:
            Cc, h, w = wl["corr"][li]
            ms = time_op(lambda d: V.correlation(sets[d][0][li][0], sets[d][0][li][1], out=sets[d][0][li][2]))
            nb = 4.0 * h * w * (2 * Cc + 81)
            ops_roof.append({"kernel": "custom::Correlation", "shape": [Cc, h, w], "us_per_launch": ms * 1e3,
                             "achieved": nb / (ms * 1e-3) / 1e9, "frac": nb / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "algorithmic_bytes_per_launch": nb,
                             "gflops": 2.0 * 81 * Cc * h * w / (ms * 1e-3) / 1e9})

    dram_frac = (traffic / (pass_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic else None
    # (the 4-step-loop form of the pass everywhere except the 10-sweep passes of 4K-class images: stab_solver_rolled.cu)
    pass_kernel = "solver_stream_kernel" if (T_main == 10 and W * H >= 4000000) else "solver_rolled_kernel"
    roofline = {"kernel": f"{pass_kernel}<{T_main}> (level 0, {T_main} Jacobi sweeps per launch, on-chip)",
                # what limits the kernel per ncu (profiles/): instruction issue + shared-memory wavefronts; the HBM
                # roofline in ALGORITHMIC bytes is kept as the contract's yardstick, dram_frac is the real DRAM load
                "bound": "issue/smem", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "dram_frac": dram_frac,
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": pass_ms * 1e3,
                "hbm_floor_us": (traffic / (peaks["hbm_gbs"] * 1e9) * 1e6) if traffic else None,
                "peak_source": peak_src,
                "note": f"algorithmic bytes = {T_main} sweeps x 72 B/pixel; temporal blocking keeps "
                        f"{T_main - 1} of {T_main} sweeps on chip, so `frac` may exceed 1; `dram_frac` = ncu DRAM "
                        "bytes per launch / launch time / HBM peak is the fraction of HBM bandwidth actually used",
                "unblocked_sweep": {"kernel": "solver_sweep_vec_kernel", "bound": "hbm", "achieved": unblocked,
                                    "frac": unblocked / peaks["hbm_gbs"], "us_per_launch": sweep_ms * 1e3},
                "fused_stage_a": {"kernel": "stage_a_rows_kernel (warp x5 -> adaptive blend -> weight, one pass)",
                                  "bound": "hbm", "achieved": stage_a_gbs, "frac": stage_a_gbs / peaks["hbm_gbs"],
                                  "us_per_launch": stage_a_ms * 1e3, "algorithmic_bytes_per_launch": stage_a_bytes},
                "custom_ops": ops_roof}
    e2e = {"h2d_bytes_per_step": 2 * W * H * 4 + (2 * fw * fh * 8 if files else 0),
           "file_bytes_per_step": 2 * (12 + fw * fh * 8) if files else 0, "d2h_bytes_per_step": W * H * 4,
           "timing": "host clock between full device synchronisations"}
    del d_o, d_p, last, cons, ws, sets, f_full, sa_out, out8, outs
    torch.cuda.empty_cache()
    return dict(ms_res=ms_res, ms_e2e=ms_e2e, launches=launches, clocks=clk, roofline=roofline, e2e=e2e,
                sustained=sustained, h2d_gbs=h2d_gbs, ms_one=ms_one)


def bench_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    placement = bind_near_gpu(local) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, args.warmup
    r = run_workload(args, args.workload, K, Wm, rank, world, dev,
                     sustained_s=args.sustained if (rank == 0 and world == 1) else 0.0)

    # ---------------- aggregate over ranks (max time) ----------------
    (ms_res, ms_e2e, neg_h2d), (launches, nbound) = reduce_over_ranks(
        [r["ms_res"], r["ms_e2e"], -r["h2d_gbs"]], [r["launches"], int(bool(placement.get("bound")))], world, dev)

    result = None
    if rank == 0:
        wl = WORKLOADS[args.workload]
        cpu = cpu_baseline(wl, args) if world == 1 and not args.no_cpu_baseline else None
        result = {
            "metric": "stabilized_frames_per_sec", "value": aggregate_fps(world, K, ms_res), "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_res / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": job_config(args.workload, world),
            "e2e": dict(value=aggregate_fps(world, K, ms_e2e), unit="frames/s", ms_per_step=ms_e2e / K,
                        pinned_h2d_gbs_min_over_ranks=-neg_h2d, host_placement=placement,
                        ranks_bound_to_gpu_numa_node=nbound, **r["e2e"]),
            "gpu_launches": int(launches),
            "clocks": r["clocks"],
            "roofline": r["roofline"],
            "pipeline": ("two streams: the custom ops of frame t+1 run on a second stream while frame t is stabilized "
                         "(stabilization(t) waits for ops(t), ops(t+1) for stabilization(t-1)); both arms of this line"
                         if (WORKLOADS[args.workload]["corr"] and not args.no_overlap_ops)
                         else "one stream: custom ops, then stabilization"),
        }
        if r.get("ms_one") is not None and world == 1:
            result["one_stream"] = {"value": aggregate_fps(1, K, r["ms_one"]), "unit": "frames/s",
                                    "ms_per_step": r["ms_one"] / K,
                                    "note": "the same steps with custom ops and stabilization on ONE stream (--no-overlap-ops)"}
        if r["sustained"]:
            result["sustained"] = r["sustained"]
        if cpu:
            result["cpu_baseline"] = cpu
    # ---------------- the other single-GPU BASELINE configs, short runs (N = 1 only) ----------------
    if world == 1 and not args.no_extras and args.workload == "1080p-light":
        extras = []
        for name in ("4k-stab", "4k-dense"):
            Ke = max(6, min(K, 12))
            x = run_workload(args, name, Ke, 3, rank, world, dev, sample_clocks=True)
            wlx = WORKLOADS[name]
            rec = {"workload": name, "config": job_config(name, 1), "steps": Ke, "warmup": 3,
                   "value": aggregate_fps(1, Ke, x["ms_res"]), "unit": "frames/s", "ms_per_step": x["ms_res"] / Ke,
                   "e2e": dict(value=aggregate_fps(1, Ke, x["ms_e2e"]), unit="frames/s", **x["e2e"]),
                   "gpu_launches": int(x["launches"]), "clocks": x["clocks"],
                   **({"one_stream": {"value": aggregate_fps(1, Ke, x["ms_one"]), "unit": "frames/s"}}
                      if x.get("ms_one") is not None else {}),
                   "roofline": {k: x["roofline"][k] for k in ("kernel", "bound", "achieved", "frac", "traffic", "dram_frac",
                                                              "us_per_launch", "fused_stage_a", "unblocked_sweep",
                                                              "custom_ops")}}
            if not args.no_cpu_baseline:
                rec["cpu_baseline"] = cpu_baseline(wlx, args, seconds=6.0, band_h=wlx["H"] // 8)
            extras.append(rec)
        result["extra"] = extras
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_frame(O, wl, band_h, state):
    """one frame (or a horizontal band of it) of the hot path on the host: reference CPU custom ops
    (oracle/_ref, the reference's own kernels, 1 thread as built) + the oracle port of the stabilization
    (the reference has no CPU stabilization code), OpenMP over all cores."""
    # the 14 op calls of a frame are independent of each other: run them on a thread pool (ctypes releases the GIL),
    # so that the reference's single-threaded CPU kernels still use all host cores
    jobs = []
    for d in range(2):
        jobs += [(state["corr_fn"], a, b) for a, b in state["corr"][d]]
        jobs += [(state["warp_fn"], x, f) for x, f in state["warp"][d]]
    if jobs:
        list(state["pool"].map(lambda j: j[0](j[1], j[2]), jobs))
    of, pf, ff, fb = state["of"], state["pf"], state["ff"], state["fb"]
    co, _ = O.do_one_step(of[0], of[1], of[2], pf[0], pf[1], pf[2], state["last"], ff, fb)
    state["last"] = co


def cpu_state(O, wl, band_h):
    import synth

    O.use_all_cores()
    W, H = wl["W"], band_h
    o8, p8 = synth.frames(W, H, 3, seed=1234)
    ffl, fbl = synth.flows(W, H, 3)
    st = {"of": [O.rgba8_to_f32x3(x) for x in o8], "pf": [O.rgba8_to_f32x3(x) for x in p8], "ff": ffl, "fb": fbl}
    st["last"] = st["pf"][2]
    use_ref = O.ref_cpu_available()
    st["corr_fn"] = O.ref_cpu_correlation if use_ref else O.correlation
    st["warp_fn"] = O.ref_cpu_warp if use_ref else O.warp_nchw
    st["ops_kind"] = ("reference (oracle/_ref; its kernels are single-threaded as built, the independent op calls of a "
                      "frame run concurrently on a thread pool)") if use_ref else "oracle port"
    from concurrent.futures import ThreadPoolExecutor

    st["pool"] = ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1))
    frac = band_h / wl["H"]
    st["corr"], st["warp"] = [], []
    for d in range(2):
        st["corr"].append([(synth.features(1, C, max(1, round(h * frac)), w, d + i),
                            synth.features(1, C, max(1, round(h * frac)), w, d + i + 7))
                           for i, (C, h, w) in enumerate(wl["corr"])])
        st["warp"].append([(synth.features(1, C, max(1, round(h * frac)), w, d + i),
                            synth.op_flow(1, max(1, round(h * frac)), w, d + i + 9))
                           for i, (C, h, w) in enumerate(wl["warp"])])
    return st


def cpu_baseline(wl, args, seconds=10.0, band_h=None):
    """the CPU arm on a bounded sample: whole frames (or a horizontal band of the frame above 1080p) repeated
    until about `seconds` of CPU work have been timed"""
    from oracle import oracle as O

    H = wl["H"]
    band_h = band_h or (H if wl["W"] * H <= 1920 * 1080 else H // 4)
    band_h -= band_h % 2
    st = cpu_state(O, wl, band_h)
    cpu_frame(O, wl, band_h, st)  # warm-up
    n = 0
    t0 = time.perf_counter()
    while True:
        cpu_frame(O, wl, band_h, st)
        n += 1
        dt = time.perf_counter() - t0
        if (dt >= seconds and n >= 3) or n >= 400:
            break
    fps = n / dt * (band_h / H)
    return {"value": fps, "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{n} frames of a {wl['W']}x{band_h} band ({band_h}/{H} of the frame, scaled linearly); "
                      f"stabilization = oracle port (OpenMP, {O.num_threads()} threads; the reference has no CPU "
                      f"stabilization code); custom ops = {st['ops_kind']}",
            "seconds": dt}


def bench_reference(args, rank, world):
    """The reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return None
    from oracle import oracle as O

    wl = WORKLOADS[args.workload]
    K, Wm = args.steps, args.warmup
    H = wl["H"]
    # bounded sample: a full 1080p frame costs ~2 s on 8 cores; keep the whole run within a few minutes
    budget_frames = 40.0 * (1920 * 1080) / (wl["W"] * H)
    band_h = H if (K + Wm) <= budget_frames else max(32, int(H * budget_frames / (K + Wm)))
    band_h -= band_h % 2
    st = cpu_state(O, wl, band_h)
    for _ in range(Wm):
        cpu_frame(O, wl, band_h, st)
    t0 = time.perf_counter()
    for _ in range(K):
        cpu_frame(O, wl, band_h, st)
    dt = time.perf_counter() - t0
    fps = K / dt * (band_h / H)
    sample = (f"each step = one {wl['W']}x{band_h} band ({band_h}/{H} of a frame, throughput scaled linearly); "
              f"stabilization = oracle port (OpenMP {O.num_threads()} threads; no reference CPU code exists), "
              f"custom ops = {st['ops_kind']}")
    return {
        "impl": "reference", "metric": "stabilized_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": dt / K * 1e3 / (band_h / H),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": job_config(args.workload, world),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p-light", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short 4k-stab / 4k-dense runs (N = 1 only)")
    ap.add_argument("--no-numa-bind", action="store_true", help="keep the launcher's CPU affinity")
    ap.add_argument("--no-overlap-ops", action="store_true",
                    help="custom ops and stabilization on ONE stream (default: the ops of frame t+1 run on a second stream "
                         "while frame t is stabilized)")
    ap.add_argument("--sustained", type=float, default=3.0,
                    help="seconds of back-to-back device-resident frames for the `sustained` record (N = 1; 0 = off)")
    ap.add_argument("--solver-mode", type=lambda x: int(x, 0), default=0, help="vsc_set_solver_mode value (A/B runs)")
    ap.add_argument("--stage-a-mode", type=lambda x: int(x, 0), default=0, help="vsc_set_stage_a_mode value (A/B runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line, the JSON result: anything a library prints while the benchmark runs (NCCL's
    # version banner at communicator creation, compiler chatter) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            res = bench_reference(args, rank, world)
        else:
            res = bench_ours(args, rank, world)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if rank == 0 and res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
