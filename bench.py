#!/usr/bin/env python
"""bench.py - LBG codebook-training throughput on B200 (the metric BASELINE.json names).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c1|c5] [--weak] [--impl reference]

Default workload: BASELINE config 3 (16384 x 16384, 2x2 blocks, K = 4096, 67 M vectors) - north_star's scaling target -
STRONG scaling: the one image is split into `--gpus` bands of block rows (N = 1 trains the whole image on one GPU).

A STEP is one complete LBG codebook train (LBGQuantizer::quantize, /root/reference/src/Quantizer.cpp:
121-143: mean -> nbits x {split, assign, accumulate, fix}) over one synthetic image.  Per step the
brute-force-equivalent work is N*(2K-2) distance evaluations (one assignment pass per split level,
SURVEY.md D2); `value` = that / device time, in Gdist-evals/s, whole job over all ranks.

  value      image already resident in HBM (borrowed device pointer); codebook comes back to the host
             every level because the next level's KD tree is built there (it is part of the path)
  e2e        the same train through the reference-facing call sequence with HOST buffers: pinned host
             RGB bytes -> qb200_set_image_band (H2D) -> qb200_train -> qb200_get_assign (D2H indices)
  roofline   the assignment kernel of the LAST level (K = 2^nbits), timed live by CUDA events on the
             library's stream: algorithmic flops = N*K*3*dim (sub, mul, add per dimension, SURVEY.md
             8d) against the FP32 FMA peak measured in the same run (MEASURED_PEAKS.json has no FP32
             figure); `hbm` = the accumulate kernel of that level against the measured copy bandwidth
  cpu_baseline / --impl reference
             the UNMODIFIED reference (oracle/_ref/libquantref_release.so: its sources + its Release
             flags, driven through CompressedImage::compress) on the box's host cores, on a bounded
             band of the same image; else the plain-C oracle port (1 thread)

Multi-GPU (torchrun, one rank per GPU): every rank owns one band of block rows (strong scaling, default: the named
image split `world` ways; --weak: one workload-sized band per rank of a `world` times taller image); the only
exchange is one sum all-reduce of K*(dim+2) 64-bit integers per split level.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (band xSize per rank, ySize, w, h, nbits, description)
WORKLOADS = {
    "c1": (768, 512, 2, 2, 10, "synthetic noise 768x512 PPM, 2x2 block, 1024 codevectors (Kodak shape)"),
    "c2": (4096, 4096, 2, 2, 10, "synthetic noise 4096x4096 PPM, 2x2 block (12-dim), 1024 codevectors, 4.2M vectors"),
    "c3": (16384, 16384, 2, 2, 12, "synthetic noise 16384x16384 PPM, 2x2 block, 4096 codevectors, 67M vectors"),
    "c4": (8192, 8192, 4, 4, 8, "synthetic noise 8192x8192 PPM, 4x4 block (48-dim), 256 codevectors"),
}
METRIC = "LBG Gdist-evals/s (N*K summed over split levels = N*(2K-2) per codebook train)"
UNIT = "Gdist-evals/s"
EPS = float(np.float32(1e-6))


def noise_band(xs, ys, seed):
    """`xs` pixel lines of `ys` pixels (the reference addresses pixel (x, y) at x*ySize + y)."""
    return np.random.default_rng(seed).integers(0, 256, (xs, ys, 3), dtype=np.uint8)


_KODAK = None


def natural_band(xs, ys, first_line):
    """`xs` pixel lines of `ys` pixels cut from an endless tiling of kodim01 (768 x 512, tests/golden): natural
    statistics - flat areas, duplicated blocks, dead cells - instead of noise."""
    global _KODAK
    if _KODAK is None:
        z = np.load(os.path.join(ROOT, "tests", "golden", "kodim01_full_2x2_n10.npz"))
        kx, ky = int(z["params"][0]), int(z["params"][1])
        _KODAK = np.ascontiguousarray(z["rgb"], np.uint8).reshape(ky, kx, 3)
    ky, kx = _KODAK.shape[:2]
    lines = (first_line + np.arange(xs)) % ky
    cols = np.arange(ys) % kx
    return np.ascontiguousarray(_KODAK[lines][:, cols])


def make_band(kind, xs, ys, seed, first_line=0, total_lines=None):
    """Lines [first_line, first_line + xs).  total_lines (strong scaling): the band is a slice of the ONE
    total_lines-high workload image, so that every rank count trains the same image."""
    if kind == "natural":
        return natural_band(xs, ys, first_line)
    if total_lines is not None and total_lines != xs:
        return np.ascontiguousarray(noise_band(total_lines, ys, seed)[first_line:first_line + xs])
    return noise_band(xs, ys, seed)


def desc_of(desc, data):
    return desc if data == "noise" else desc.replace("synthetic noise", "synthetic (kodim01 tiled: natural statistics)")


def evals_per_train(n_vectors, nbits):
    return float(n_vectors) * (2.0 * (1 << nbits) - 2.0)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx.append(float(f[1]))
                if t0 <= t <= t1 + 0.1:
                    sm.append(float(f[0]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if t0 <= t <= t1 + 0.1 and v.lower().startswith("active"):
                    reasons.add(name)
        if not mx:
            return None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------
def cpu_engine():
    """('reference', RefLib release) when oracle/_ref travelled here, else ('port', PortLib)."""
    from oracle.pyoracle import PortLib, RefLib, have_ref
    if have_ref("release"):
        try:
            return "reference", RefLib("release")
        except OSError:
            pass
    so = os.path.join(ROOT, "oracle", "liblbg_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return "port", PortLib()


def cpu_train_seconds(kind, eng, band, xs, ys, w, h, nbits):
    """One train on the CPU, timed over the reference's own scope (block extraction + quantize,
    src/Compressor.cpp:118-123)."""
    if kind == "reference":
        return eng.compress(band, xs, ys, w, h, nbits, want_outputs=False)["seconds"]
    t = time.perf_counter()
    X = eng.blocks(band, xs, ys, w, h)
    eng.quantize(X, nbits)
    return time.perf_counter() - t


def cpu_sample(kind, eng, wl, budget_s, steps, warmup, data="noise"):
    """Times `steps` trains (after `warmup`) on a band of the workload image sized to fit budget_s."""
    bx, ys, w, h, nbits, _ = WORKLOADS[wl]
    cores = 1
    if kind == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for every host core explicitly
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        eng.set_threads(cores)
        cores = eng.max_threads()
    # calibrate t(lines) = a + b*lines on two small bands (the first call also warms the OpenMP pool)
    p1 = max(w * 8, min(bx, 64 if kind == "reference" else 16))
    p2 = min(bx, 4 * p1)
    cpu_train_seconds(kind, eng, make_band(data, p1, ys, 1234), p1, ys, w, h, nbits)
    t1 = cpu_train_seconds(kind, eng, make_band(data, p1, ys, 1234), p1, ys, w, h, nbits)
    t2 = cpu_train_seconds(kind, eng, make_band(data, p2, ys, 1234), p2, ys, w, h, nbits) if p2 > p1 else t1
    b = max((t2 - t1) / max(p2 - p1, 1), 1e-9)
    a = max(t1 - b * p1, 0.0)
    xs = int((budget_s / (steps + warmup) - a) / b)
    xs = max(p1, min(bx, (xs // (8 * w)) * 8 * w))
    band = make_band(data, xs, ys, 1234)
    n_vec = ((xs + w - 1) // w) * ((ys + h - 1) // h)
    for _ in range(warmup):
        cpu_train_seconds(kind, eng, band, xs, ys, w, h, nbits)
    times = [cpu_train_seconds(kind, eng, band, xs, ys, w, h, nbits) for _ in range(steps)]
    sec = float(np.mean(times))
    val = evals_per_train(n_vec, nbits) / sec / 1e9
    sample = (f"first {xs} of {bx} pixel lines of the {wl} image ({n_vec} of "
              f"{(bx // w) * (ys // h)} vectors), full K={1 << nbits} train, {steps} timed run(s), mean")
    return dict(value=val, unit=UNIT, cores=cores, kind=kind, sample=sample), sec, xs, n_vec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, eng = cpu_engine()
    wl = args.workload
    bx, ys, w, h, nbits, desc = WORKLOADS[wl]
    base, sec, xs, n_vec = cpu_sample(kind, eng, wl, args.cpu_budget, args.steps, max(args.warmup, 1), args.data)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc_of(desc, args.data), "bounded_sample": base["sample"], "block": [w, h], "nbits": nbits,
                   "colorspace": "SCALED", "cpu_threads": base["cores"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.data == "noise" and not args.no_natural:          # the same shape on natural data, bounded the same way
        nat, nsec, _, _ = cpu_sample(kind, eng, wl, 0.4 * args.cpu_budget, 1, 1, "natural")
        line["natural"] = {"data": "kodim01 (tests/golden) tiled to the workload's shape", "value": nat["value"], "unit": UNIT,
                           "ms_per_step": nsec * 1e3, "bounded_sample": nat["sample"], "cpu_threads": nat["cores"]}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def cpp_compress_leg(band, xs, ys, w, h, nbits, steps, n_total, cb_expect):
    import ctypes as C
    so = os.path.join(ROOT, "quant_b200", "host", "libquantsrc.so")
    if not os.path.exists(so):
        return {"unavailable": "quant_b200/host/libquantsrc.so not built"}
    L = C.CDLL(so)
    L.quantsrc_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float),
                                    C.POINTER(C.c_double)]
    L.quantsrc_last_error.restype = C.c_char_p
    K, dim = 1 << nbits, 3 * w * h
    pix = np.array(band, np.uint8, copy=True)                 # plain pageable memory
    cbb = np.empty((K, dim), np.uint8)
    a = np.empty(n_total, np.uint64)
    d, bpp, sec = C.c_double(), C.c_float(), C.c_double()
    times = []
    for i in range(steps + 1):                                # first call: context creation + allocations (warm-up)
        rc = L.quantsrc_compress(pix.ctypes.data, xs, ys, 1, w, h, EPS, nbits, cbb.ctypes.data, a.ctypes.data,
                                 C.byref(d), C.byref(bpp), C.byref(sec))
        if rc < 0:
            return {"unavailable": (L.quantsrc_last_error() or b"").decode()}
        if i:
            times.append(sec.value)
    import quant_b200 as qb
    same = bool(np.array_equal(cbb, qb.codebook_to_bytes(cb_expect, qb.CS_SCALED)))
    t = float(np.mean(times))
    return {"value": evals_per_train(n_total, nbits) / t / 1e9, "unit": UNIT, "seconds_per_train": t, "steps": steps,
            "through": "libquantsrc.so CompressedImage::compress (C++ drop-in API): pageable std::vector<RGB> in, "
                       "std::vector<size_t> indices out; the report's own 'Compression time' (host clock)",
            "h2d_bytes_per_step": int(pix.size), "d2h_bytes_per_step": int(n_total) * 4,
            "codebook_bytes_equal_c_abi_run": same, "report_mse": d.value, "bits_per_pixel": bpp.value}


def bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this rank's host thread - and so the first touch of its pinned buffers - on the NUMA node
    its GPU hangs off (sysfs numa_node of the device's PCI address).  Returns the node, or None when it is unknown or
    the binding is not possible; never fatal."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) < 2:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def run_gpu_arm(args):
    import torch
    import quant_b200 as qb
    from quant_b200.distributed import join_peer_group, make_allreduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    # opt-in (QB200_BENCH_NUMA=1): not measured on a multi-GPU box this round, so the default run stays as validated
    numa_node = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("QB200_BENCH_NUMA", "0") == "1" else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    bx, ys, w, h, nbits, desc = WORKLOADS[wl]
    K, dim = 1 << nbits, 3 * w * h
    if args.strong:                            # strong scaling (default): the named image is split into `world` bands
        if bx % (w * world):
            raise SystemExit("strong scaling needs the image's pixel lines to divide by block width x ranks")
        bx = bx // world
    xs_total = bx * world                      # --weak: one workload-sized band per rank
    wB_band = bx // w
    row_begin, row_end = rank * wB_band, (rank + 1) * wB_band
    n_local = wB_band * (ys // h)
    n_total = n_local * world

    stream = torch.cuda.Stream()
    ctx = qb.Context(local)
    ctx.set_stream(stream.cuda_stream)
    # the per-level sum all-reduce: the library's own peer-memory exchange (qb200_comm.cu; torch.distributed only
    # carries the CUDA IPC handles once), or - with --nccl - NCCL through a torch.distributed callback per level
    allreduce = None
    if world > 1 and args.nccl:
        allreduce = make_allreduce()
    elif world > 1:
        join_peer_group(ctx)
    exact = args.exact or os.environ.get("QB200_EXACT_CENTROIDS") == "1"
    ctx.set_exact_centroids(1 if exact else ("auto" if args.centroids == "auto" else 0))
    if world > 1:
        ctx.set_rank(rank, world)

    if args.strong:   # the same image whatever the rank count
        band_np = make_band(args.data, bx, ys, 1234, rank * bx, total_lines=xs_total).reshape(-1)
    else:
        band_np = make_band(args.data, bx, ys, 1234 + rank, rank * bx).reshape(-1)
    host_band = torch.empty(band_np.size, dtype=torch.uint8, pin_memory=True)
    host_band.numpy()[:] = band_np
    host_assign = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    host_assign_np = host_assign.numpy().view(np.uint32)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, mode, collect=None):
        """mode 'resident': device image borrowed; 'e2e': host bytes in, indices out. Returns device ms."""
        total = 0.0
        for _ in range(n):
            with torch.cuda.stream(stream):
                flush.zero_()                                  # evict the image and codebook from L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if mode == "resident":
                    cb, d, rep = ctx.train(nbits, EPS, n_total=n_total, allreduce=allreduce)
                else:
                    ctx.set_image_band(host_band.numpy(), xs_total, ys, w, h, qb.CS_SCALED, row_begin, row_end)
                    cb, d, rep = ctx.train(nbits, EPS, n_total=n_total, allreduce=allreduce)
                    ctx.get_assign(out=host_assign_np)
                e1.record(stream)
                e1.synchronize()
                total += e0.elapsed_time(e1)
                if collect is not None:
                    collect.append(rep)
        return total, cb, d

    def timed(mode, steps, warmup):
        if mode == "resident":
            with torch.cuda.stream(stream):
                dev_band = host_band.to("cuda", non_blocking=True)
            stream.synchronize()
            ctx.set_image_band(None, xs_total, ys, w, h, qb.CS_SCALED, row_begin, row_end,
                               device_ptr=dev_band.data_ptr(), nbytes=dev_band.numel(), keep=dev_band)
        run_steps(warmup, mode)
        reps = []
        barrier()
        qb.launch_count(reset=True)
        t0 = time.time()
        ms, cb, d = run_steps(steps, mode, reps)
        barrier()
        t1 = time.time()
        launches = qb.launch_count()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, reps, launches, (t0, t1), cb, d

    fp32_peak = ctx.measure_fp32_peak() if rank == 0 else None
    sampler = ClockSampler(local) if rank == 0 else None
    ms_res, reps, launches, (t0, t1), cb_res, d_res = timed("resident", args.steps, args.warmup)
    took_exact_res = bool(ctx.last_train_exact)
    ms_e2e, _, _, (t2, t3), cb_e2e, d_e2e = timed("e2e", args.steps, args.warmup)
    # clocks over both timed regions (a train lasts milliseconds: the e2e warm-up in between is under load as well)
    clocks = sampler.stop(t0, t3) if sampler else None
    assert np.array_equal(cb_res, cb_e2e) and d_res == d_e2e, "resident and e2e trains disagree"

    # When the default (auto) mode had to repeat the train with the compensated sums, also time the integer-sum mode
    # alone (mode 0): what the same workload costs without the end-to-end identity guarantee (round 1's default).
    integer_only = None
    if took_exact_res and not exact:
        ctx.set_exact_centroids(0)
        ms_int, _, _, _, _, _ = timed("resident", min(args.steps, 3), 2)
        ctx.set_exact_centroids("auto" if args.centroids == "auto" else 0)
        integer_only = {"ms_per_step": ms_int / min(args.steps, 3),
                        "value": evals_per_train(n_total, nbits) / (ms_int / min(args.steps, 3) * 1e-3) / 1e9, "unit": UNIT,
                        "what": "qb200_set_exact_centroids(ctx, 0): integer-sum centroids only; every assignment pass is still "
                                "bit-identical given the same codebook, the final indices only on inputs without tie-sensitive decisions"}

    # The same shape on NATURAL data (kodim01 tiled): flat areas, duplicated blocks and dead cells push far more
    # queries through the exact FP64 resolver and its tree walk than noise does (and the reference's KD search is
    # faster there), so the noise headline alone would flatter the comparison.  Reported next to it.
    natural = None
    # (one GPU by default; `--natural` asks for it at N > 1 too - there it takes the sharded exact repeat, whose chains
    #  run rank after rank, and was not timed under torchrun this round)
    if args.data == "noise" and not args.no_natural and (world == 1 or args.natural):
        host_band.numpy()[:] = natural_band(bx, ys, rank * bx).reshape(-1)
        n_steps = min(args.steps, 3)
        ms_nr, reps_n, _, _, _, d_nat = timed("resident", n_steps, 3)
        nat_exact = bool(ctx.last_train_exact)
        ms_ne, _, _, _, _, _ = timed("e2e", n_steps, 3)
        if rank == 0:
            ev = evals_per_train(n_total, nbits)
            natural = {"data": "kodim01 (tests/golden) tiled to the workload's shape",
                       "value": ev * n_steps / (ms_nr * 1e-3) / 1e9, "ms_per_step": ms_nr / n_steps,
                       "e2e": {"value": ev * n_steps / (ms_ne * 1e-3) / 1e9, "ms_per_step": ms_ne / n_steps}, "unit": UNIT,
                       "steps": n_steps, "distortion": d_nat,
                       "flagged_per_level": {str(r["K"]): int(r["flagged"]) for r in reps_n[-1]},
                       "tie_sensitive_decisions": int(sum(r["sensitive"] for r in reps_n[-1])),
                       "repeated_with_compensated_sums": nat_exact,
                       "kd_walk_ties_per_level": {str(r["K"]): int(r["ties"]) for r in reps_n[-1]},
                       "dead_cells_last_level": int(reps_n[-1][-1]["dead_cells"]),
                       "ms_resolve_per_train": float(np.mean([sum(r["ms_resolve"] for r in rep) for rep in reps_n]))}
        host_band.numpy()[:] = band_np

    # Second end-to-end leg: the reference-facing C++ call itself - CompressedImage::compress of libquantsrc.so
    # (quant_b200/host), PAGEABLE std::vector<RGB> pixels in, std::vector<size_t> indices out, timed by the C++
    # layer's own host clock over the reference's scope (src/Compressor.cpp:118-123).  One process, one GPU.
    e2e_cpp = None
    if rank == 0 and world == 1 and not args.no_cpp:
        e2e_cpp = cpp_compress_leg(band_np, xs_total, ys, w, h, nbits, min(args.steps, 5), n_total, cb_res)

    if rank == 0:
        evals = evals_per_train(n_total, nbits)
        value = evals * args.steps / (ms_res * 1e-3) / 1e9
        e2e = evals * args.steps / (ms_e2e * 1e-3) / 1e9
        last = [r[-1] for r in reps]
        ms_assign = float(np.mean([r["ms_assign"] for r in last]))
        ms_acc = float(np.mean([r["ms_accumulate"] for r in last]))
        per_level = {str(r["K"]): {k: round(float(np.mean([x[i][k] for x in reps])), 4)
                                   for k in ("ms_assign", "ms_resolve", "ms_accumulate")}
                     for i, r in enumerate(reps[0])}
        sensitive_total = int(sum(r["sensitive"] for r in reps[-1]))
        took_exact = took_exact_res
        flagged_last = int(last[-1]["flagged"])
        ties_last = int(last[-1]["ties"])
        flops = float(n_local) * K * 3.0 * dim            # algorithmic: sub, mul, add per dimension and evaluation
        alg_tflops = flops / (ms_assign * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        tc_peak = float(peaks.get("bf16_tflops", 1590.0))
        tc_src = "MEASURED_PEAKS.json bf16_tflops, burst (of measured)" if "bf16_tflops" in peaks else "1590 TFLOP/s (of fallback)"
        acc_gbs = float(n_local) * (dim + 4) / (ms_acc * 1e-3) / 1e9 if ms_acc > 0 else None
        uses_tc = K >= max(int(os.environ.get("QB200_TC_MIN_K", "128") or 128), 64) and os.environ.get("QB200_DISABLE_TC", "0") != "1"
        kb = (dim + 1 + 15) // 16
        k_pad = ((K + 255) // 256) * 256
        # executed tensor flops of the filter: 128-query tiles x padded codebook x (3 bf16 limbs x 16*kb) x 2
        tc_flops = float(((n_local + 127) // 128) * 128) * k_pad * (3 * 16 * kb) * 2.0
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's launch, from the committed ncu capture
        # of this workload (profiles/traffic.json, written by tools/ncu_summary.py); null when there is none
        traffic = traffic_note = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{wl}/n{world}/{'tc' if uses_tc else 'cc'}"
            traffic = tj.get(key, {}).get("bytes")
            traffic_note = tj.get(key, {}).get("note") or None
        except (OSError, ValueError):
            pass
        if uses_tc:
            roof = {"bound": "tensor", "kernel": f"assign_tc_kernel<{dim}> (+ its finalise kernel) at K={K}, last split level",
                    "achieved": alg_tflops, "peak": tc_peak, "unit": "TFLOP/s", "frac": alg_tflops / tc_peak, "peak_source": tc_src,
                    "algorithmic": "SURVEY 8d: 3*dim flop per distance evaluation x N*K evaluations of the launch",
                    "executed": {"achieved": tc_flops / (ms_assign * 1e-3) / 1e12, "frac": tc_flops / (ms_assign * 1e-3) / 1e12 / tc_peak,
                                 "what": "bf16 MMA flops really issued: ceil(N/128)*128 x K padded to 256 x 3 limbs x 16*ceil((dim+1)/16) x 2"},
                    "note": "the contraction runs on the tensor pipe (3 bf16 limbs = 24 mantissa bits); the kernel's own limiter is the "
                            "min-selection over the accumulator columns on the alu pipe (1.125 min/max ops per evaluation, ncu: profiles/)",
                    "ms_per_launch": ms_assign, "traffic": traffic, "traffic_algorithmic": float(n_local) * dim}
        else:
            roof = {"bound": "fp32", "kernel": f"assign_kernel<{dim}> at K={K}, last split level", "achieved": alg_tflops,
                    "peak": fp32_peak, "unit": "TFLOP/s", "frac": alg_tflops / fp32_peak,
                    "peak_source": "FP32 FFMA probe in this run (FMA = 2 flop)", "ms_per_launch": ms_assign, "traffic": traffic}
        if traffic_note:
            roof["traffic_note"] = traffic_note   # which launch of the step the capture is (profiles/traffic.json)
        roof["gdist_evals_per_s"] = float(n_local) * K / (ms_assign * 1e-3) / 1e9
        roof["fp32_equivalent"] = {"achieved": alg_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": alg_tflops / fp32_peak,
                                   "algorithmic": "3*dim flop per distance evaluation x N*K evaluations (the reference's sub, mul, add)",
                                   "peak_source": "FP32 FFMA probe in this run (FMA = 2 flop); MEASURED_PEAKS.json has no FP32 figure"}
        if acc_gbs:
            roof["hbm"] = {"kernel": "per-cell integer statistics pass, same level", "achieved": acc_gbs, "peak": hbm_peak,
                           "unit": "GB/s", "frac": acc_gbs / hbm_peak, "peak_source": hbm_src, "ms_per_launch": ms_acc,
                           "algorithmic": "(dim + 4) bytes per vector"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
            "dtype": "bf16x3 tensor-core / f32 filter + f64 exact re-check; i64 sums",
            "data": "synthetic",
            "config": {"workload": desc_of(desc, args.data) + ((f"; strong scaling: split into {world} bands" if args.strong else
                                            f"; weak scaling: {world} such bands, one per rank") if world > 1 else ""),
                       "block": [w, h], "nbits": nbits, "colorspace": "SCALED", "vectors_per_rank": n_local,
                       "schedule": "reference HEAD: one assignment pass per split level, no empty-cell repair",
                       "centroids": ("exact: the reference's compensated FP64 member sums (parallel bit-exact evaluation)" if exact else
                                     ("auto (library default): integer per-cell sums; the timed trains had "
                                      f"{sensitive_total} tie-sensitive decisions and "
                                      + ("were repeated with the reference's compensated sums" if took_exact else
                                         "needed no repeat: index-identical to the reference's by construction"))
                                     if args.centroids == "auto" else "from integer per-cell sums (<= 4e-16 relative of the reference's)"),
                       "allreduce": (None if world == 1 else "NCCL via torch.distributed callback" if args.nccl else
                                     "libqb200 peer-memory all-reduce (qb200_comm.cu), CUDA IPC between the rank processes"),
                       "host_numa_node_rank0": numa_node,   # multi-rank: host thread + pinned buffers bound to the GPU's node
                       "l2": "512 MiB flush (memset) before every timed step, outside the timed events"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(host_band.numel()) * world,
                    "d2h_bytes_per_step": (n_local * 4) * world,
                    "seconds_per_train": ms_e2e / args.steps / 1e3},
            "e2e_cpp": e2e_cpp, "natural": natural, "integer_sum_mode": integer_only,
            "gpu_launches": launches,
            "roofline": roof,
            "per_level_ms": per_level, "sensitive_per_level": {str(r["K"]): int(r["sensitive"]) for r in reps[-1]},
            "sensitive_diag": {str(r["K"]): int(r["reserved"]) for r in reps[-1] if r["reserved"]},
            "flagged_last_level": flagged_last, "kd_walk_ties_last_level": ties_last,
            "distortion": d_res, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            kind, eng = cpu_engine()
            base, _, _, _ = cpu_sample(kind, eng, wl, args.cpu_budget, 1, 1, args.data)
            line["cpu_baseline"] = base
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner from C) write to
    file descriptor 1 as well, so fd 1 is pointed at stderr for the whole run and the line goes out through
    a private duplicate of the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


_JSON_FD = None


def run_encode_arm(args):
    """BASELINE config 5: encode-only assignment of a batch of Kodak-shaped images against a fixed trained
    1024-entry codebook (the body of Solution::assignCodeVectors = KDTree + nearestNeighbour per vector).
    A step is one assignment pass over the batch; value = N*K / time.  No collective: ranks take batches."""
    import torch
    import quant_b200 as qb
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xs, ys, w, h, nbits, n_img = 768, 512, 2, 2, 10, args.images
    K, dim = 1 << nbits, 12
    n_local = n_img * (xs // w) * (ys // h)
    stream = torch.cuda.Stream()
    ctx = qb.Context(local)
    ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1234 + rank)
    host = torch.empty(n_img * xs * ys * 3, dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = rng.integers(0, 256, host.numel(), dtype=np.uint8)
    host_assign = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
    with torch.cuda.stream(stream):
        ctx.set_image(host.numpy()[: xs * ys * 3], xs, ys, w, h, qb.CS_SCALED)
        codebook, _, _ = ctx.train(nbits)                       # the fixed codebook: trained on the first image
        dev = host.to("cuda", non_blocking=True)
    stream.synchronize()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def run(n, mode):
        total, last = 0.0, None
        for _ in range(n):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if mode == "resident":
                    last = ctx.assign_only(codebook)
                else:
                    ctx.set_image(host.numpy(), xs, ys, w, h, qb.CS_SCALED, n_images=n_img)
                    last = ctx.assign_only(codebook)
                    ctx.get_assign(out=host_assign.numpy().view(np.uint32))
                e1.record(stream)
                e1.synchronize()
                total += e0.elapsed_time(e1)
        return total, last

    def timed(mode):
        if mode == "resident":
            ctx.set_image_device(dev.data_ptr(), xs, ys, w, h, qb.CS_SCALED, n_images=n_img, keep=dev)
        run(args.warmup, mode)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        qb.launch_count(reset=True)
        ms, last = run(args.steps, mode)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        launches = qb.launch_count()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last, launches

    ms_res, last, launches = timed("resident")
    ms_e2e, _, _ = timed("e2e")
    if rank == 0:
        evals = float(n_local) * world * K
        line = {"metric": "encode-only Gdist-evals/s (N*K per assignment pass)", "value": evals * args.steps / (ms_res * 1e-3) / 1e9,
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16x3 tensor-core filter + f64 exact re-check", "data": "synthetic",
                "config": {"workload": f"encode-only: {n_img} synthetic noise 768x512 images per rank, 2x2 block, against a fixed "
                                       f"trained 1024-entry codebook ({n_local} vectors per rank)",
                           "l2": "512 MiB flush before every timed step"},
                "e2e": {"value": evals * args.steps / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": int(host.numel()) * world, "d2h_bytes_per_step": n_local * 4 * world},
                "gpu_launches": launches, "flagged": last["flagged"],
                "ms_filter": last["ms_assign"], "ms_resolver": last["ms_resolve"]}
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--images", type=int, default=1024, help="images per rank of the encode-only workload c5")
    ap.add_argument("--cpu-budget", type=float, default=None, help="seconds of CPU work for the CPU arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nccl", action="store_true",
                    help="multi-GPU: reduce through NCCL (torch.distributed callback per level) instead of the library's "
                         "peer-memory all-reduce")
    ap.add_argument("--no-cpp", action="store_true", help="skip the C++ CompressedImage::compress end-to-end leg")
    ap.add_argument("--no-natural", action="store_true", help="skip the natural-image leg")
    ap.add_argument("--natural", action="store_true", help="run the natural-image leg under torchrun (N > 1) as well")
    ap.add_argument("--centroids", default="auto", choices=["auto", "integer"],
                    help="auto (the library's default): integer-sum centroids, repeated with the reference's compensated sums "
                         "when the train had tie-sensitive decisions; integer: integer sums only")
    ap.add_argument("--exact", action="store_true",
                    help="bit-exact centroid mode (qb200_set_exact_centroids): slower, identical to the reference on any input")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: one workload-sized band per rank (default: strong, the workload image is split across the ranks)")
    ap.add_argument("--strong", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--data", default="noise", choices=["noise", "natural"],
                    help="noise: uniform random bytes (worst case for the reference's KD tree, few ties); natural: the Kodak "
                         "image of tests/golden tiled to the workload's shape (flat areas, duplicates, dead cells)")
    args = ap.parse_args()
    args.strong = not args.weak
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.cpu_budget is None:
        args.cpu_budget = 120.0 if args.impl == "reference" else 20.0
    if args.workload == "c5":
        if args.impl == "reference":
            raise SystemExit("--workload c5 has no reference arm (the reference has no encode-only entry point)")
        sys.exit(run_encode_arm(args))
    sys.exit(run_reference_arm(args) if args.impl == "reference" else run_gpu_arm(args))


if __name__ == "__main__":
    main()
