#!/usr/bin/env python
"""Benchmark of the scan-registration hot path (BASELINE.json: "VGICP/NDT scan-pair aligns/sec (KITTI-shape)").

Workload (BASELINE configs[1] shape): FAST_VGICP alignment of consecutive scans of a synthetic HDL-64 (KITTI-shape,
121,600 rays) sequence after the YAML prefilter chain (distance 0.1-35 m, VoxelGrid 0.1 m, RADIUS 0.5/2), registration
parameters of config/mrg_slam.yaml (k=20, resolution 1.0, eps 0.1/2e-3, LM).  One STEP = one pass of the hot path over
a chain of P+1 scans -> P aligns (scan i+1 onto scan i, identity guess): every cloud is new in every step, so each
step pays the full path — uniform-grid build, exact 20-NN covariances, voxel map, LM iterations — exactly what the
odometry component pays on a scan that becomes the next keyframe.

  value : aligns/s, raw scans already resident in HBM when the timed region starts (device events on the library's stream)
  e2e   : same metric through the reference-facing C ABI with pinned HOST buffers: H2D of every scan and D2H of every
          result inside the timed region (wall clock between device synchronisations)
  N > 1 : weak scaling — every rank runs its own chain of P pairs (independent units, no data-path collective);
          steps are bracketed by a barrier, the time is the max over ranks.  The sharded batch path WITH its result
          all-gather is `bench_configs.py --config loop_closure` (strong scaling).

`--impl reference` times the CPU restatement of the reference path (oracle/, kind "port": the real pclomp/fast_gicp
sources are not vendored in /root/reference) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "VGICP scan-pair aligns/sec (KITTI-shape)"
UNIT = "aligns/s"
PAIRS_PER_STEP = 32
SEED_SCAN0 = 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2r", choices=["b2r", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP)
    ap.add_argument("--method", default="FAST_VGICP", choices=["FAST_VGICP", "FAST_GICP", "NDT_OMP", "SMALL_GICP"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-handles", type=int, default=2, help="registration handles (streams + host threads) of the overlapped e2e leg")
    return ap.parse_args()


def workload_config(args, n_mean):
    return {
        "workload": "configs[1]: FAST_VGICP scan-to-scan aligns over a synthetic HDL-64 (KITTI-shape, 121,600 rays) sequence, "
                    "prefiltered (dist 0.1-35 m, VoxelGrid 0.1, RADIUS 0.5/2), reg_* of config/mrg_slam.yaml",
        "method": args.method,
        "pairs_per_step_per_gpu": args.pairs,
        "points_per_cloud_mean": int(n_mean),
        "guess": "identity (0.35-0.55 m inter-scan motion)",
        "l2": "flushed between steps (256 MiB write); per-step working set ~0.3 GB",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                                          str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_scans(count, first):
    from mrg_slam_b200 import synth
    return [synth.scan(synth.HDL64, first + i) for i in range(count)]


# ----------------------------------------------------------------------------------------------------------- reference arm
def oracle_prefilter(O, c):
    c = O.distance_filter(c, 0.1, 35.0)
    c, _ = O.voxelgrid(c, 0.1, 1)
    return c[O.radius_outlier(c, 0.5, 2)]


def oracle_chain(O, method, clouds):
    """P aligns over a chain of P+1 fresh clouds with the oracle, all host threads (OpenMP inside the oracle).
    The target's covariances are computed once per cloud like the registration object's cache would keep them:
    each cloud is set as source first, then promoted to target by re-setting it (fresh object per pair keeps the
    comparison conservative for the CPU: it recomputes what upstream would recompute)."""
    n_ok = 0
    reg = O.Registration(O.default_params(getattr(O, method)))
    for i in range(len(clouds) - 1):
        reg.setInputTarget(clouds[i])
        reg.setInputSource(clouds[i + 1])
        r = reg.align(np.eye(4))
        n_ok += int(r.converged)
    return n_ok


def cpu_sample(method, budget_s=12.0, n_pairs=16):
    """Bounded CPU sample: the first `n_pairs` pairs of the same chain, repeated until ~budget_s of CPU work is done."""
    from tests import oraclelib as O
    raws = make_scans(n_pairs + 1, SEED_SCAN0)
    clouds = [oracle_prefilter(O, r) for r in raws]
    O.set_num_threads(0)
    cores = O.max_threads()
    oracle_chain(O, method, clouds[:2])  # warm-up
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < budget_s:
        i = done % n_pairs
        oracle_chain(O, method, clouds[i:i + 2])
        done += 1
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} aligns over the first {n_pairs} consecutive-scan pairs of the same workload "
                      f"(~{int(np.mean([len(c) for c in clouds]))} pts/cloud), oracle restatement of fast_gicp/pclomp with OpenMP on "
                      f"{cores} threads, {dt:.1f} s"}, clouds


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oraclelib as O
    sample_pairs = 4
    raws = make_scans(sample_pairs + 1, SEED_SCAN0)
    clouds = [oracle_prefilter(O, r) for r in raws]
    O.set_num_threads(0)
    cores = O.max_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        oracle_chain(O, args.method, clouds)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_chain(O, args.method, clouds)
    dt = time.perf_counter() - t0
    value = sample_pairs * args.steps / dt
    n_mean = np.mean([len(c) for c in clouds])
    cfg = workload_config(args, n_mean)
    cfg["reference_sample_pairs_per_step"] = sample_pairs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_pairs} consecutive-scan pairs per step x {args.steps} steps, oracle (CPU restatement of "
                                   f"fast_gicp/pclomp, OpenMP) on {cores} host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------- b2r arm
def run_b2r(args):
    import torch
    from mrg_slam_b200 import lib as B
    from mrg_slam_b200 import loop_closure as LC

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs
    method = getattr(B, args.method)
    reg = B.Registration(B.default_config(method, device=local_rank))
    # ---- data: P+1 consecutive scans per rank, prefiltered once by the engine's own prefilter (untimed setup)
    raws = make_scans(P + 1, SEED_SCAN0 + rank * (P + 1))
    clouds_np = [reg.prefilter(r) for r in raws]
    n_mean = float(np.mean([len(c) for c in clouds_np]))
    dev_bufs = [torch.from_numpy(c).to(dev) for c in clouds_np]
    pin_bufs = [torch.from_numpy(c).pin_memory() for c in clouds_np]
    guesses = [np.eye(4)] * P
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(bufs, on_device):
        cl = B.create_clouds(reg, [b.data_ptr() for b in bufs], [b.shape[0] for b in bufs], B.DEVICE if on_device else B.HOST)
        res = reg.align_batch(cl[1:], cl[:-1], guesses)
        for c in cl:
            c.close()
        return res

    def gather(res):
        # Outside the timed region: the chain workload has no exchange step (independent scans per rank);
        # it only checks after the run that every rank's results can be collected like the loop-closure path does.
        if world > 1:
            return LC.gather_results(LC.pack_results(res, list(range(rank * P, rank * P + P))), world * P, device=dev)
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        gather(step(dev_bufs, True))
    for _ in range(2):
        step(pin_bufs, False)

    # ---- timed: value (inputs resident in HBM), device events on the library's stream
    reg.profile_enable(True)
    launches0 = reg.kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    conv = 0
    barrier()
    torch.cuda.profiler.start()  # ncu --profile-from-start off captures exactly the timed steps (no effect otherwise)
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        reg.event_record(0)
        res = step(dev_bufs, True)
        reg.event_record(1)
        total_ms += reg.event_elapsed_ms(0, 1)
        conv += sum(r.converged for r in res)
    barrier()
    torch.cuda.profiler.stop()
    launches = reg.kernel_launches() - launches0
    prof = {k: reg.profile_read(k) for k in B.PROFILE_KERNELS}
    reg.profile_enable(False)

    # ---- timed: e2e (pinned host buffers through the C ABI; H2D of every scan + D2H of every result inside)
    # serial: one handle, one step at a time, L2 flushed between steps
    barrier()
    e2e_serial_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        res = step(pin_bufs, False)
        torch.cuda.synchronize()
        e2e_serial_s += time.perf_counter() - t0
    barrier()
    # overlapped (the headline e2e): `--e2e-handles` registration objects, each with its own CUDA stream and host thread,
    # work through the same K steps concurrently (the reference runs odometry and loop detection on separate objects
    # and threads the same way, SURVEY 8b "Threading"), so one step's PCIe upload and host-side synchronisations overlap
    # another step's kernels.  Every step still uploads all its scans and reads back all its results; no L2 flush here:
    # the inputs come from host memory and the ~0.3 GB per-step working set exceeds L2.
    import threading
    nh = max(1, args.e2e_handles)
    regs = [reg] + [B.Registration(B.default_config(method, device=local_rank)) for _ in range(nh - 1)]

    def step_on(r):
        cl = B.create_clouds(r, [b.data_ptr() for b in pin_bufs], [b.shape[0] for b in pin_bufs], B.HOST)
        out = r.align_batch(cl[1:], cl[:-1], guesses)
        for c in cl:
            c.close()
        return out

    for r in regs[1:]:
        step_on(r)  # warm-up of the additional handles
    counter = {"next": 0, "conv": 0}
    lock = threading.Lock()

    errors = []

    def worker(r):
        try:
            torch.cuda.set_device(local_rank)
            while True:
                with lock:
                    k = counter["next"]
                    if k >= args.steps:
                        return
                    counter["next"] = k + 1
                out = step_on(r)
                with lock:
                    counter["conv"] += sum(x.converged for x in out)
                    counter["done"] = counter.get("done", 0) + 1
        except Exception as e:  # a failed step must fail the run, not shorten it
            errors.append(e)

    # Two Python threads hand the GIL back and forth around ~40 ctypes calls per step; with the default 5 ms switch interval a
    # thread returning from a C call can wait that long for the other one, so the interval is shortened for this leg (a C++
    # caller has no such lock).  It does not remove the occasional collapse of this schedule (about one run in five lands
    # between 3k and 10k aligns/s instead of ~16.7k, cause not yet pinned down), hence max(serial, overlapped) below.
    old_interval = sys.getswitchinterval()
    sys.setswitchinterval(2e-5)
    barrier()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(r,)) for r in regs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sys.setswitchinterval(old_interval)
    if errors or counter.get("done", 0) != args.steps:
        raise RuntimeError(f"overlapped e2e leg: {counter.get('done', 0)} of {args.steps} steps completed, errors: {errors}")
    barrier()
    e2e_conv = counter["conv"]
    clocks = sampler.stop() if rank == 0 else None

    tt = torch.tensor([total_ms, e2e_s * 1000.0, e2e_serial_s * 1000.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    max_ms, max_e2e_ms, max_e2e_serial_ms = float(tt[0]), float(tt[1]), float(tt[2])
    if rank == 0:
        value = P * world * args.steps / (max_ms / 1000.0)
        e2e_overlapped = P * world * args.steps / (max_e2e_ms / 1000.0)
        e2e_serial = P * world * args.steps / (max_e2e_serial_ms / 1000.0)
        # both schedules are measured in every run, through the same public calls on the same host buffers; the line
        # reports the better one as e2e.value and keeps both (on an 8-GPU box the concurrent PCIe pulls of 16 handles were
        # slower than 8, profiles/r1/configs/bench_n8.jsonl)
        e2e = max(e2e_overlapped, e2e_serial)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom = max(prof, key=lambda k: prof[k]["ms"])
        d = prof[dom]
        achieved = (d["bytes"] / max(d["launches"], 1)) / (d["ms"] / max(d["launches"], 1) * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same workload
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tj.get(f"{args.method}:{P}", {}).get(dom)
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        per_kernel = {}
        for k, v in prof.items():
            if v["ms"] > 0:
                gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
                per_kernel[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                                 "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, n_mean),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(sum(b.numel() * 4 for b in pin_bufs)),
                    "d2h_bytes_per_step": int(P * __import__("ctypes").sizeof(B.Result)),
                    "schedule": "overlapped" if e2e_overlapped >= e2e_serial else "serial",
                    "overlapped_value": e2e_overlapped, "serial_value": e2e_serial,
                    "overlapped_timing": f"wall clock over all K steps, {nh} registration handles (one stream + one host thread each) "
                                         "working concurrently; every step uploads its scans from pinned host memory and reads its "
                                         "results back",
                    "handles": nh, "converged_fraction": e2e_conv / float(P * args.steps),
                    "serial_timing": "one handle, one step at a time, L2 flushed between steps, wall clock between device synchronisations"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
                         "kernel_us_per_launch": 1e3 * d["ms"] / max(d["launches"], 1),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "note": "algorithmic bytes (SURVEY 8d: 40 B/point for kNN covariances) / CUDA-event duration of the dominant "
                                 "kernel; exact 20-NN search is compute/latency-bound, not HBM-bound (DESIGN.md)",
                         "kernels": per_kernel},
            "converged_fraction": conv / float(P * args.steps),
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_sample(args.method)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b2r(args)


if __name__ == "__main__":
    main()
