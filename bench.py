#!/usr/bin/env python
"""bench.py — msa2eds -l 10 on the synthetic alignment of BASELINE.json (config 2 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (eds_msa_transform_device) over one resident alignment window.
N = 1: config 2 (100 sequences x 10 Mbp, 1 % variable columns, wrap 80, l = 10).
N > 1 (torchrun, one rank per GPU): weak scaling — the alignment is 100 x (N * 10 Mbp), rank g holds
the column window [g * 10 Mbp, (g + 1) * 10 Mbp) plus a halo, transforms it, and the ranks all-gather
their (eds, seds) byte counts over NCCL to get the file offsets of their slices (the path's only
exchange step). value = total cells / max-over-ranks device time.

The JSON line carries: value (device-resident), e2e (host buffers through eds_msa_transform_host, H2D
and D2H inside the timed region), roofline (k_scan, the dominant kernel: algorithmic bytes R * C per
launch / CUDA-event time, against MEASURED_PEAKS.json), cpu_baseline (the UNMODIFIED reference library
oracle/_ref/ref_driver, or the oracle port, on a bounded sample of the same alignment).

--impl reference: the reference's own CPU implementation (oracle/_ref/ref_driver msa2eds, single
threaded by construction: msa2eds has no thread option) on a bounded sample per step.
"""
import argparse
import ctypes
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R = 100
C_PER_GPU = 10_000_000
WRAP = 80
L = 10
SEED = 1
PPM = 10_000
HALO = 4096
SAMPLE_COLS = 2_000_000  # CPU legs: 100 x 2 Mbp = 2e8 cells, a few seconds of reference time per pass
METRIC = "MSA cells/s (msa2eds -l 10)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, rows, cols):
    """dram read+write bytes per launch of `kernel` on a rows x cols window, from the committed ncu capture of that
    very shape (profiles/ncu_traffic.json: {"<kernel>": {"<rows>x<cols>": bytes}}); None for any other shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel, {}).get(f"{rows}x{cols}")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def write_sample_file(cols):
    from edsparser_b200 import synth

    text = synth.fasta_window(R, cols, WRAP, seed=SEED, variable_ppm=PPM)
    fd, path = tempfile.mkstemp(suffix=".msa", prefix="edsb_sample_")
    with os.fdopen(fd, "wb") as f:
        f.write(text)
    return path, len(text)


def reference_pass(path, workdir):
    """One msa2eds -l 10 pass of the reference's own library over `path`. Returns (seconds, kind)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    eds, seds = os.path.join(workdir, "o.leds"), os.path.join(workdir, "o.seds")
    if os.path.exists(ref):
        out = subprocess.run([ref, "msa2eds", path, str(L), eds, seds], check=True, capture_output=True, text=True).stdout
        return float(re.search(r"seconds=([0-9.eE+-]+)", out).group(1)), "reference"
    port = os.path.join(ROOT, "oracle", "eds_oracle")
    if not os.path.exists(port):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "restatement"], stdout=subprocess.DEVNULL)
    t0 = time.perf_counter()
    subprocess.run([port, "msa2eds", path, str(L), eds, seds], check=True)
    return time.perf_counter() - t0, "port"


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded sample per step, sized so that warmup + steps passes stay near two minutes of reference time
    # (the reference sustains ~1.3e8 cells/s on one host core)
    cols = int(max(100_000, min(SAMPLE_COLS, 1.5e10 / (max(1, args.steps + args.warmup) * R))))
    path, nbytes = write_sample_file(cols)
    cells = R * cols
    with tempfile.TemporaryDirectory() as wd:
        kind = "reference"
        for _ in range(args.warmup):
            _, kind = reference_pass(path, wd)
        times = []
        for _ in range(args.steps):
            t, kind = reference_pass(path, wd)
            times.append(t)
    os.unlink(path)
    total = sum(times)
    value = cells * args.steps / total
    sample = f"{R} seq x {cols} columns of the config-2 generator (seed {SEED}), msa2eds -l {L}, in-memory library call"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": dict(workload_config(args.gpus), workload=workload_config(args.gpus)["workload"] +
                       f" [this arm times a bounded sample per step: {R} seq x {cols} columns of the same generator; "
                       "cells/s is size-independent for the reference, which is linear in the column count]"),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, config=2):
    if config == 4:
        return {"workload": f"config 4: synthetic MSA {R} seq x {C_PER_GPU * n} columns, 1% variable columns, wrap {WRAP}, "
                            f"msa2eds -l {L}, column-sharded over {n} GPU(s), halo {HALO}",
                "rows": R, "cols_per_gpu": C_PER_GPU, "context_length": L, "seed": SEED,
                "l2": "inputs larger than L2, no explicit flush", "parallelism": f"column-sharded x{n}"}
    return {"workload": f"config 2 per GPU: synthetic MSA {R} seq x {C_PER_GPU} columns, 1% variable columns, "
                        f"wrap {WRAP}, msa2eds -l {L}" + ("" if n == 1 else f"; {n} column shards of a {R} x {n * C_PER_GPU} alignment, halo {HALO}"),
            "rows": R, "cols_per_gpu": C_PER_GPU, "context_length": L, "seed": SEED,
            "l2": "inputs larger than L2 (1.01 GB window per GPU vs 126 MB L2), no explicit flush",
            "parallelism": f"column-sharded x{n}" if n > 1 else "single GPU"}


def pin_to_gpu_numa(index):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    afterwards (first touch) and its copy submissions stay on that socket: with 8 ranks streaming 1 GB each per step
    through the host, remote-node memory is what makes the end-to-end leg stop scaling. Best effort; returns a note."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out  # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: no allowed cpu"
        os.sched_setaffinity(0, cpus)
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception as err:
        return f"not pinned ({type(err).__name__})"


def kernel_profile(ctx, view, steps):
    """Per-kernel CUDA-event times of `steps` transforms (profiling adds an event pair per launch)."""
    ctx.set_profiling(True)
    acc = {}
    for _ in range(steps):
        ctx.msa_transform_device(view, L)
        for name, t in ctx.kernel_times():
            acc[name] = acc.get(name, 0.0) + t
    ctx.set_profiling(False)
    return {k: v / steps for k, v in acc.items()}


NO_EXCHANGE = os.environ.get("EDSB_BENCH_NO_EXCHANGE") == "1"  # diagnosis only: a line measured this way is not a bench value


def measure_device(lib, ctx, dist, dev, rank, world, rows, cols_per_gpu, steps, warmup, exchange):
    """Device-resident leg: window [rank] of a rows x (world * cols_per_gpu) alignment generated in HBM, `steps`
    transforms timed with CUDA events (max over ranks), then the per-kernel profile. Returns a dict (rank-local)."""
    import torch

    import edsparser_b200 as E
    from edsparser_b200 import shard

    total_cols = cols_per_gpu * world
    halo = HALO
    while True:
        lo, hi, wb, we = shard.plan(total_cols, world, rank, halo)
        view = ctx.msa_synth(rows, total_cols, WRAP, col_begin=wb, col_count=we - wb, seed=SEED, variable_ppm=PPM)
        view.own_begin, view.own_end = lo, hi
        try:
            ctx.msa_transform_device(view, L)
            break
        except E.EdsError as err:  # EDS_ERR_HALO: a symbol does not close inside the window -> widen and retry
            if err.status != E.EDS_ERR_HALO or halo > (1 << 24):
                raise
            halo *= 4

    c_e, c_s, c_st = E.Buffer(), E.Buffer(), E.capi.MsaStats()
    c_args = (ctx.handle, ctypes.byref(view), L, 1, ctypes.byref(c_e), ctypes.byref(c_s), ctypes.byref(c_st))
    c_call = lib.L.eds_msa_transform_device

    def step():
        # the bare C call with preallocated argument structs: the transform returns after a device synchronisation, so
        # every microsecond of Python between two calls is device idle time inside the timed region
        rc = c_call(*c_args)
        if rc:
            lib.check(rc)
        # file offsets of this rank's slices: all-gather of the byte counts, issued behind the transform; the host
        # reads them once, before it writes (exchange.offsets() after the loop)
        if not NO_EXCHANGE:
            exchange.post(c_e.bytes, c_s.bytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    launches_per_step = int(c_st.gpu_launches)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    exchange.flush()  # the last exchanges are inside the timed region
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    out_bytes = int(c_e.bytes) + int(c_s.bytes)
    kern = kernel_profile(ctx, view, min(steps, 5))
    row_bytes = (we - 1) + (we - 1) // WRAP - (wb + wb // WRAP) + 1
    return {"view": view, "ms_total": ms_total, "out_bytes": out_bytes, "kern": kern, "lo": lo, "hi": hi, "wb": wb, "we": we,
            "row_bytes": row_bytes, "launches_per_step": launches_per_step, "halo": halo}


def roofline_of(m, rows, steps, world, serial_kern=None):
    """roofline object of one measured leg: dominant kernel = the scan (fused with the column gather when it runs)."""
    peak, peak_src = peaks()
    kern = m["kern"]
    name = "k_scan_l2" if "k_scan_l2" in kern else ("k_scan_fused" if "k_scan_fused" in kern else "k_scan")
    scan_ms = kern.get(name, 0.0)
    scan_bytes = rows * m["row_bytes"]  # every cell of the window read once (newlines ride along)
    achieved = scan_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
    # whole step, summed over ranks: cells of the owned ranges + output bytes of this rank x ranks (ranks are alike)
    step_alg_bytes = (rows * (m["hi"] - m["lo"]) + m["out_bytes"]) * world
    step_gbs = step_alg_bytes * steps / (m["ms_total"] / 1e3) / 1e9
    out = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": ncu_traffic(name, rows, m["we"] - m["wb"]), "peak_source": peak_src, "kernel_ms": scan_ms,
           "algorithmic_bytes_per_launch": scan_bytes,
           "step": {"algorithmic_bytes": step_alg_bytes, "achieved_gbs": step_gbs, "frac": step_gbs / (peak * world),
                    "peak_aggregate": peak * world},
           "kernels_ms": {k: round(v, 4) for k, v in kern.items()},
           "kernels_ms_note": "CUDA events around each launch while independent kernels overlap on side streams "
                              "(overlapped ones read long); kernels_ms_serial = the same with everything on one stream"}
    if serial_kern:
        out["kernels_ms_serial"] = {k: round(v, 4) for k, v in serial_kern.items()}
    return out


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    import edsparser_b200 as E

    local = int(os.environ.get("LOCAL_RANK", 0))
    numa_note = pin_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = E.load()  # raises if the CUDA library is missing: there is no CPU fallback
    stream = torch.cuda.current_stream().cuda_stream
    ctx = lib.context(local, stream)

    # the path's one exchange — an all-gather of the ranks' output byte counts — is posted by the LIBRARY (NCCL called
    # from C++, include/edsparser_b200.h eds_comm_*); torch.distributed only ships the 128-byte NCCL id and the barriers
    uid = [lib.nccl_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    exchange = E.Comm(ctx, uid[0], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def serial_profile(view):
        """the same transform on a context whose kernels all run on one stream: per-kernel times without overlap"""
        os.environ["EDSB_DEBUG_SERIAL"] = "1"
        try:
            c2 = lib.context(local, stream)
        finally:
            del os.environ["EDSB_DEBUG_SERIAL"]
        try:
            c2.msa_transform_device(view, L)
            return kernel_profile(c2, view, 3)
        finally:
            c2.close()

    warmup = max(args.warmup, 3)
    sampler = ClockSampler(local) if rank == 0 else None
    m = measure_device(lib, ctx, dist, dev, rank, world, R, C_PER_GPU, args.steps, warmup, exchange)
    my_offsets = (0, 0, 0, 0) if NO_EXCHANGE else exchange.offsets()  # (eds offset, seds offset, eds total, seds total) of this rank's slices
    clocks = sampler.stop() if sampler else None
    view, ms_total, out_bytes = m["view"], m["ms_total"], m["out_bytes"]
    cells_step = R * C_PER_GPU * world
    value = cells_step * args.steps / (ms_total / 1e3)
    roofline = roofline_of(m, R, args.steps, world, serial_profile(view))
    launches_per_step = m["launches_per_step"]

    # ---- end to end: .msa bytes in pinned host memory -> eds_msa_transform_host -> host strings
    text = ctx.download(E.Buffer(view.text, view.text_bytes)) if world == 1 and args.config == 2 else None
    e2e = None
    if args.config == 4:
        pass  # device-resident scaling figure only (30 GB of pinned host text is not a bench default)
    elif world == 1:
        pinned = torch.empty(len(text), dtype=torch.uint8).pin_memory()
        pinned.copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
        ctx.msa_synth_free()
        e2e_steps = max(1, min(args.steps, 5))
        he, hs, _ = ctx.msa_transform_host(pinned, L)  # warm-up, and the bytes for the size fields
        ctx.msa_transform_host_view_raw(pinned.data_ptr(), len(text), L)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            # H2D + index + kernels + D2H; the results land in pinned host memory kept by the context
            ctx.msa_transform_host_view_raw(pinned.data_ptr(), len(text), L)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": cells_step * e2e_steps / dt, "unit": "cells/s", "h2d_bytes_per_step": len(text),
               "d2h_bytes_per_step": len(he) + len(hs), "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
               "api": "eds_msa_transform_host_view (index + H2D + kernels + D2H into pinned host memory), pinned input"}
        del pinned
    else:
        # per-rank shard from pinned host memory, same call; slowest rank decides
        shard_text = ctx.download(E.Buffer(view.text, view.text_bytes))
        pinned = torch.empty(len(shard_text), dtype=torch.uint8).pin_memory()
        pinned.copy_(torch.frombuffer(bytearray(shard_text), dtype=torch.uint8))
        dtext = torch.empty(len(shard_text) + 64, dtype=torch.uint8, device=dev)
        hview = E.MsaView()
        ctypes.memmove(ctypes.byref(hview), ctypes.byref(view), ctypes.sizeof(view))
        hview.text = dtext.data_ptr()
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            dtext[: len(shard_text)].copy_(pinned, non_blocking=True)
            ee, ss, _ = ctx.msa_transform_device(hview, L)
            exchange.post(int(ee.bytes), int(ss.bytes))
            a, b = ctx.download_view(0, ee), ctx.download_view(1, ss)  # D2H into pinned host memory kept by the context
            exchange.offsets()  # where this rank's slices go in the one file pair
            return int(a.bytes) + int(b.bytes)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nout = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": cells_step * e2e_steps / float(dt.item()), "unit": "cells/s",
               "h2d_bytes_per_step": len(shard_text) * world, "d2h_bytes_per_step": nout * world,
               "ms_per_step": 1e3 * float(dt.item()) / e2e_steps, "steps": e2e_steps, "host_placement": numa_note,
               "api": "per rank: pinned H2D + eds_msa_transform_device + D2H into pinned host memory + eds_comm_post / eds_comm_offsets (NCCL all-gather of the byte counts, issued by the library)"}
        del pinned, dtext
    ctx.msa_synth_free()

    # ---- BASELINE config 4 (1000 x 30 Mbp, strong-scaled over the ranks) rides in the same line: the configuration the
    # north-star fraction is quoted on. Device-resident only.
    config4 = None
    if args.config == 2 and not args.no_config4:
        try:
            rows4, cols4 = 1000, 30_000_000 // world
            steps4 = max(3, min(args.steps, 10))
            m4 = measure_device(lib, ctx, dist, dev, rank, world, rows4, cols4, steps4, 3, exchange)
            exchange.offsets()
            r4 = roofline_of(m4, rows4, steps4, world, serial_profile(m4["view"]) if world == 1 else None)
            ctx.msa_synth_free()
            config4 = {"workload": f"config 4: synthetic MSA {rows4} seq x {cols4 * world} columns, 1% variable columns, wrap {WRAP}, "
                                   f"msa2eds -l {L}, column-sharded over {world} GPU(s), halo {m4['halo']}",
                       "scaling": "strong", "steps": steps4, "ms_per_step": m4["ms_total"] / steps4,
                       "value": rows4 * cols4 * world * steps4 / (m4["ms_total"] / 1e3), "unit": "cells/s",
                       "output_bytes_per_rank": m4["out_bytes"], "roofline": r4}
        except Exception as err:  # the headline line must still be printed
            config4 = {"error": str(err)[:300]}

    # ---- CPU baseline: the reference library on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and args.config == 2:
        path, _ = write_sample_file(SAMPLE_COLS)
        with tempfile.TemporaryDirectory() as wd:
            reference_pass(path, wd)
            ts = []
            kind = "reference"
            t_begin = time.perf_counter()
            while len(ts) < 3 and time.perf_counter() - t_begin < 25:
                t, kind = reference_pass(path, wd)
                ts.append(t)
        os.unlink(path)
        cpu = {"value": R * SAMPLE_COLS / min(ts), "unit": "cells/s", "cores": 1, "kind": kind,
               "sample": f"{R} seq x {SAMPLE_COLS} columns of the same generator (seed {SEED}), msa2eds -l {L}, "
                         f"best of {len(ts)} in-memory library calls; msa2eds is single-threaded by construction",
               "host_cores": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.config == 2 else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world, args.config), "gb_per_s": value / 1e9, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "output_bytes_per_step": out_bytes, "library": lib.version(), "config4": config4,
        }
        print(json.dumps(line), flush=True)
    barrier()  # every rank has finished its device work before any rank tears its context down
    exchange.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def use_config4(world):
    """BASELINE config 4: 1000 sequences x 30 Mbp in total, column-sharded over the ranks (strong scaling)."""
    global R, C_PER_GPU, SAMPLE_COLS
    R = 1000
    C_PER_GPU = 30_000_000 // world
    SAMPLE_COLS = 200_000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 (1000 x 30 Mbp) leg of the default run")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="2 (default, the metric's configuration): 100 x 10 Mbp per GPU, weak scaling; 4: BASELINE config 4, "
                         "1000 x 30 Mbp column-sharded over the N GPUs (strong scaling, no e2e / cpu legs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.config == 4:
        use_config4(max(world, 1))
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
