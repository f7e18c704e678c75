#!/usr/bin/env python
"""bench.py -- seed-extension throughput (GCUPS on cells_band, tasks/s alongside) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--tasks T] [--workload NAME]

A step = one pass of the hot path (bsw level-1 batch = sw_pe_array_sw_extend per task) over one batch of synthetic
extension tasks of BASELINE.json configs[1] (1M x 150 bp, defaults a=1 b=4 o=6 e=1 w=100 zdrop=100 end_bonus=5).
  value     whole-job GCUPS with the packed batch already resident in HBM (bsw_resident_run), device-timed
  e2e       the same metric through the public C-ABI call with HOST buffers (pack + H2D + kernels + D2H inside)
  roofline  INT-ALU pipe: 13 integer ops per DP cell against the IADD3/VIMNMX rate measured live on
            this GPU (MEASURED_PEAKS.json has no integer figure); roofline_hbm gives the streaming side
  cpu_baseline  the oracle (C restatement of the reference recurrence) on this box's host cores, bounded sample
N > 1: one process per GPU (torchrun), tasks are independent so every rank takes its own shard (weak scaling), no
data-path collective; the barrier / max-over-ranks plumbing uses torch.distributed (nccl).
--impl reference: the reference's own CPU implementation of the path cannot be built (it is Verilog RTL; its software
twin bwa-0.7.8 is not mounted), so the arm times the oracle port on all host cores (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# libbsw.so's host pipeline drives ~100 streams: ask the driver for 32 hardware queues before any CUDA context exists
# (the bsw_b200 package sets the same default on import; torch would otherwise create the context first)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: NCCL's version banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

METRIC = "seed-extension GCUPS (cells_band; ksw_extend2 recurrence, bit-exact), tasks/s alongside"
OPS_PER_CELL = 13            # 1 add, 4 sub, 7 max, 1 select
FALLBACK_HBM_GBS = 6650.0    # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum per bench step (1 M x 150 bp) come from profiles/dram_traffic.json, written
# from an ncu capture of this command (tools/ncu_traffic.py) together with the commit it was taken at; absent -> null


def ncu_traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    except Exception:
        return None


def workload_config(args):
    """The `config` object: identical in both arms (the driver compares it)."""
    return {"workload": f"{args.workload}: {args.tasks} synthetic extension tasks per GPU per step (BASELINE.json configs[1] generator, "
                        "tools/../csrc/synth.cpp, seed 1)",
            "scoring": "a=1 b=4 o=6 e=1 w=100 zdrop=100 end_bonus=5", "variant": "V1 (RTL / BWA 0.7.8 recurrence)"}


def native_oracle():
    """The CPU baseline as BASELINE.md states it: the oracle compiled -O3 -march=native ON THIS BOX (the portable
    libbswref.so that travels with the repo is x86-64-v2).  Falls back to the portable build if gcc is missing."""
    import oracle as O
    try:
        out_dir = os.path.join(ROOT, "oracle", "_native")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libbswref.so")
        subprocess.check_call(["gcc", "-O3", "-march=native", "-pthread", "-fPIC", "-std=c11", "-shared", "-o", so,
                               os.path.join(ROOT, "oracle", "ksw_extend_ref.c"), os.path.join(ROOT, "oracle", "rtl_width_model.c")],
                              stderr=subprocess.DEVNULL)
        O.use_library(so)
        return "gcc -O3 -march=native"
    except Exception:
        return "gcc -O3 -march=x86-64-v2 (portable build; native rebuild failed)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="bsw", choices=["bsw", "reference"])
    ap.add_argument("--workload", default="cfg2_150bp")
    ap.add_argument("--tasks", type=int, default=1_000_000, help="tasks per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="tasks of the CPU baseline sample (default: the whole batch, ~18 core-seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_run(workload: str, n: int, first: int, threads: int):
    """Time the oracle on `n` tasks of the workload with `threads` host threads.  Returns (seconds, cells)."""
    import bsw_b200 as B
    import oracle as O
    t = B.synth_tasks(workload, n, first=first)
    po = O.make_params()
    O.extend_batch(po, t["qbuf"][:64], t["qoff"][:2], t["tbuf"][:512], t["toff"][:2], t["h0"][:1], t["w"][:1])   # load the .so
    t0 = time.perf_counter()
    _, cells = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=1, nthreads=threads)
    dt = time.perf_counter() - t0
    return dt, int(cells.sum())


def run_reference(args, rank: int, world: int):
    """--impl reference: the oracle port on all host cores, bounded sample per step (rank 0 only)."""
    if rank != 0:
        return
    import oracle as O
    flags = native_oracle()
    threads = O.max_threads()
    n = min(args.tasks, args.cpu_sample)
    times, cells = [], 0
    for s in range(args.warmup + args.steps):
        dt, cells = cpu_oracle_run(args.workload, n, 0, threads)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    gcups = cells * len(times) / total * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 (CPU port)", "data": "synthetic",
        "tasks_per_s": n * len(times) / total,
        "config": workload_config(args),
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port", "compiler": flags,
                         "sample": f"{n} tasks of {args.workload} per step, {len(times)} steps, oracle/ksw_extend_ref.c with {threads} pthreads "
                                   "(the reference is Verilog RTL: it is pinned through a translated cycle model, oracle/_ref, which "
                                   "runs ~1 M clocks/s and is no baseline)"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "bsw":
        # launched without torchrun: re-launch as one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import bsw_b200 as B

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout must carry the one JSON
        # line only, so file descriptor 1 points at stderr until the communicator exists.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ncores = os.cpu_count() or 1
    host_threads = max(1, ncores // world)
    ctx = B.Context(devices=[local_rank], streams_per_device=2, host_threads=host_threads)
    n = args.tasks
    t = B.synth_tasks(args.workload, n, first=rank * n)            # every rank owns its own shard (weak scaling)
    p = B.make_params()
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    alg_bytes = int((t["qoff"][-1] + t["toff"][-1] + 1) // 2 + n * (16 + 32))   # 4-bit bases + slot scalars + result record
    cells_rect = float((np.diff(t["qoff"]).astype(np.float64) * np.diff(t["toff"]).astype(np.float64)).sum())   # sum of qlen*tlen

    peak = ctx.measure_int_peak(0) if rank == 0 else None
    res = ctx.resident(p, *flat)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    # ---------------- value: inputs resident in HBM ----------------
    for _ in range(max(args.warmup, 3)):
        res.run()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_begin = time.time()
    kernel_ms, cells, launches = [], 0, 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        ms, cells, nl = res.run()                                  # CUDA events on the library's launch stream
        kernel_ms.append(ms)
        launches += nl
    barrier()
    dev_ms = float(sum(kernel_ms))
    res.free()

    # ---------------- e2e: public C-ABI call with host buffers ----------------
    out_buf = np.zeros(n, dtype=B.RESULT_DTYPE)                     # caller-owned result array, as a C caller would hold
    # the caller's base buffers are page-locked in place, as a host that reuses its batch buffers would hold them; the
    # library then chooses per call between staging with host threads and copying the bases in place ("raw_inputs" auto)
    ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"]); ctx.register_host(out_buf)
    for _ in range(max(args.warmup, 3)):
        ctx.sw_extend_batch(p, *flat, want_cells=False, out=out_buf)
    ctx.reset_stats()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.sw_extend_batch(p, *flat, want_cells=False, out=out_buf)
    barrier()
    e2e_s = time.perf_counter() - e0
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None       # clocks over both timed regions (value + e2e)
    st = ctx.stats()

    # ---------------- max over ranks ----------------
    agg = torch.tensor([dev_ms, e2e_s, float(cells), float(launches), cells_rect], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = float(mx[0]), float(mx[1])
        cells_all, launches_all, rect_all = float(sm[2]), int(sm[3]), float(sm[4])
    else:
        cells_all, launches_all, rect_all = float(cells), launches, cells_rect
    if rank == 0:
        steps = args.steps
        default_wl = args.workload == "cfg2_150bp" and n == 1_000_000
        traffic = ncu_traffic()
        gcups = cells_all * steps / (dev_ms * 1e-3) * 1e-9
        e2e_gcups = cells_all * steps / e2e_s * 1e-9
        int_peak = peak["vimnmx_tops"]
        achieved = OPS_PER_CELL * float(cells) * steps / (dev_ms * 1e-3) * 1e-12          # this GPU, Tops/s
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
        hbm_ach = alg_bytes * steps / (dev_ms * 1e-3) * 1e-9
        line = {
            "metric": METRIC, "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2", "data": "synthetic",
            "tasks_per_s": n * world * steps / (dev_ms * 1e-3),
            "gcups_rect": rect_all * steps / (dev_ms * 1e-3) * 1e-9,    # qlen*tlen cells, for comparison with the literature
            "config": workload_config(args),
            "measurement": {"l2": "256 MB flush write between timed steps", "cells_per_step_per_gpu": int(cells),
                            "timing": "value = sum of CUDA-event durations of the kernel launches on the library stream; e2e = wall clock around the blocking C-ABI call"},
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "gcups_rect": rect_all * steps / e2e_s * 1e-9,
                    "tasks_per_s": n * world * steps / e2e_s, "ms_per_step": e2e_s / steps * 1e3,
                    "h2d_bytes_per_step": int(st["h2d_bytes"] // steps), "d2h_bytes_per_step": int(st["d2h_bytes"] // steps),
                    "host_pack_ms_per_step": st["pack_ms"] / steps, "host_threads": host_threads,
                    "inputs": "numpy base and result arrays registered with bsw_host_register (page-locked in place); raw_inputs=auto, device_plan=1"},
            "gpu_launches": launches_all,
            "clocks": clocks,
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": int_peak, "unit": "Tops/s", "frac": achieved / int_peak,
                         "frac_dpx": achieved / peak["dpx_tops"],
                         "traffic": (traffic or {}).get("k1") if default_wl else None,
                         "traffic_unit": "bytes of DRAM per step, K1 launches (ncu)", "traffic_capture": (traffic or {}).get("capture"),
                         "ops_per_cell": OPS_PER_CELL,
                         "peak_source": "bsw_measure_int_peak on this GPU: dependency-free VIMNMX stream (ALU pipe, where every max of the "
                                        "recurrence must issue); IADD3 %.2f, fused VIADDMNMX %.2f (2 ops/instr), IADD3+IMAD both pipes %.2f Tops/s at %.0f MHz"
                                        % (peak["iadd_tops"], peak["dpx_tops"], peak["dual_tops"], peak["sm_clock_mhz"])},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                             "traffic": ((traffic or {}).get("k0", 0) + (traffic or {}).get("k1", 0)) if (default_wl and traffic) else None,
                             "algorithmic_bytes": alg_bytes, "traffic_unit": "bytes of DRAM per step, K0 + K1 launches (ncu)",
                             "traffic_capture": (traffic or {}).get("capture"),
                             "peak_source": hbm_src,
                             "note": "streaming evidence only: the path is compute-bound at ~1e3 int-ops per byte"},
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle as O
            flags = native_oracle()
            threads = O.max_threads()
            ns = min(n, args.cpu_sample)
            best = None
            for _ in range(2):
                dt, c = cpu_oracle_run(args.workload, ns, 0, threads)
                best = dt if best is None else min(best, dt)
            dt1, c1 = cpu_oracle_run(args.workload, min(ns, 20000), 0, 1)
            line["cpu_baseline"] = {"value": c / best * 1e-9, "unit": "GCUPS", "cores": threads, "kind": "port", "compiler": flags,
                                    "tasks_per_s": ns / best, "single_thread_gcups": c1 / dt1 * 1e-9,
                                    "sample": f"first {ns} tasks of the same workload, best of 2, oracle/ksw_extend_ref.c with {threads} pthreads"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
