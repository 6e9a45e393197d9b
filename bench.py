#!/usr/bin/env python
"""bench.py — candidate x trace reward evals/sec (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference arm: CPU oracle, all host threads

A "step" is one pass of the hot path over the resident workload:
K1 reward9 (Form D, 36 B/eval) -> K2 detect6 over the corpus (+ fused segmented sum and
radix top-K at one rank; ncclAllReduce + K3 at N > 1) -> result block to the host.
Workload at N=1: BASELINE configs[2] = 256 candidates x 10 M records (92.16 GB resident) + the
10 M-record corpus; weak scaling: every rank holds 256 x 10 M, the global record axis is
N x 10 M, sharded with no data-path collective except the one allreduce of the partials.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate_x_trace_reward_evals_per_sec"
UNIT = "evals/s"
SEED = 0x5EED0003


def read_traffic(C, T):
    """DRAM bytes of one K1 launch from the committed ncu capture (profiles/traffic.json); None when the
    benchmarked configuration is not the profiled one."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)["k_reward9"]
        if (t["C"], t["T"]) == (C, T):
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def read_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local: int) -> str:
    """Pin this rank to the CPUs of its GPU's NUMA node before it allocates pinned host buffers
    (first touch places them next to the GPU's PCIe root: matters for the 8-rank e2e leg)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node} ({len(cpus)} cpus)"
    except Exception as ex:
        return f"numa binding skipped: {ex}"
    return "numa binding skipped"


def cpu_baseline(threads: int, target_s: float = 12.0) -> dict:
    """The oracle (C port of the reference TypeScript) timed on this box's host cores on a
    bounded sample of the same workload: `C` candidates x `T` records of the same generator."""
    import oracle
    oracle.build()
    Cn, T = 32, 200_000
    dims = oracle.gen_dims(SEED, 0, Cn, 0, T, 300, threads)
    t0 = time.perf_counter()
    oracle.score_dims(dims, nthreads=threads)
    dt = time.perf_counter() - t0
    reps = max(1, min(200, int(target_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        s, n = oracle.score_dims(dims, nthreads=threads)
    dt = (time.perf_counter() - t0) / reps
    # the parity oracle itself: one thread, exact reference order (SURVEY 8d "CPU baseline" (i)); ~3 s
    t0 = time.perf_counter()
    one = 0
    while time.perf_counter() - t0 < 3.0:
        oracle.score_dims(dims[:8], nthreads=1)
        one += 1
    dt1 = (time.perf_counter() - t0) / one
    return {"value": Cn * T / dt, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": 8 * T / dt1,
            "sample": f"{Cn} candidates x {T} records of the configs[2] generator (seed {SEED:#x}), {reps} passes, "
                      f"oracle/apo_oracle.c orc_score_dims_mt; the reference TypeScript cannot run here (no JS runtime)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port
    (the reference is TypeScript and no JS runtime exists here), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = len(os.sched_getaffinity(0))
    Cn, T = 32, 200_000                       # bounded sample of configs[2] (256 x 10 M)
    dims = oracle.gen_dims(SEED, 0, Cn, 0, T, 300, threads)
    recs = oracle.gen_records(SEED, oracle.STREAM_CORPUS, 0, 1, 0, T, 300, threads).reshape(-1)

    def step():
        s, n = oracle.score_dims(dims, nthreads=threads)
        oracle.topk(s, Cn // 4)
        oracle.report(recs)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = Cn * T / dt
    # the same step when the caller holds trace records (the reference's own representation) instead of dims:
    # TCS:668-763 per evaluation + the mean — the CPU side of the e2e_records16 leg of the GPU arm
    roll = oracle.gen_records(SEED, oracle.STREAM_ROLLOUT, 0, Cn, 0, T, 300, threads)
    oracle.score_records(roll, nthreads=threads)
    t0 = time.perf_counter()
    nrec = max(2, args.steps // 2)
    for _ in range(nrec):
        oracle.score_records(roll, nthreads=threads)
    v_rec = Cn * T / ((time.perf_counter() - t0) / nrec)
    sample = f"{Cn} candidates x {T} records + {T}-record corpus per step (bounded sample of 256 x 10M), C oracle port, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2] 256-beam x 10M-span finalReward + detect6 + top-K (bounded CPU sample)", "C": Cn, "T": T},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "from_records": {"value": v_rec, "unit": UNIT, "note": "same sample as per-(candidate, record) trace records: dims derived per evaluation on the CPU"},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--candidates", type=int, default=256)
    ap.add_argument("--records", type=int, default=10_000_000, help="records per GPU")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--recip", type=int, default=0)
    ap.add_argument("--e2e-candidates", type=int, default=64)
    ap.add_argument("--e2e-records", type=int, default=1_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "single rank: no binding"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("senweaver-ide_b200")
    eng = pkg.Engine(local)
    if world > 1:
        box = [pkg.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(world, rank, box[0])

    C, T, K = args.candidates, args.records, max(1, args.candidates // 4)
    free, total = torch.cuda.mem_get_info()
    need = C * ((T + 31) // 32 * 32) * 36 + T * 32 + (3 << 30)
    note = ""
    if need > free:
        T = int((free - (4 << 30)) // (C * 36 + 32)) // 1024 * 1024
        note = f"records per GPU reduced to {T} to fit {free >> 30} GiB free"
    t0 = rank * T
    eng.dims_generate(SEED, 0, C, t0, T, 300)
    eng.corpus_generate(SEED, t0, T, 300)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.score(C, K, corpus=True, variant=args.variant, recip=bool(args.recip))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms, k2_ms, ar_ms, launches = [], [], [], 0
    w0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        r = eng.score(C, K, corpus=True, variant=args.variant, recip=bool(args.recip))
        k1_ms.append(r.timing.reward_ms); k2_ms.append(r.timing.corpus_ms); ar_ms.append(r.timing.allreduce_ms + r.timing.finalize_ms)
        launches += r.timing.launches
    ev1.record(stream)
    barrier()
    w1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1) / args.steps
    tms = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    clocks = sampler.stop(w0, w1) if rank == 0 else None

    # ---- the same workload held in the compact resident layout (Form Q, 14 B/eval, lossless; csrc/apo_compact.cu):
    # transcoded once at load, scored by K1q.  Reported beside the fp32 numbers, not instead of them.
    compact = None
    try:
        engq = pkg.Engine(local)
        if world > 1:
            box = [pkg.Engine.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            engq.comm_init(world, rank, box[0])
        engq.dims_generate_compact(SEED, 0, C, t0, T, 300)
        engq.corpus_generate(SEED, t0, T, 300)
        engq.set_stream(stream.cuda_stream)
        for _ in range(args.warmup):
            rq = engq.score(C, K, corpus=True)
        same = bool(np.array_equal(rq.scores, r.scores) and np.array_equal(rq.topk, r.topk))
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kq_ms = []
        q0.record(stream)
        for _ in range(args.steps):
            rq = engq.score(C, K, corpus=True)
            kq_ms.append(rq.timing.reward_ms)
        q1.record(stream)
        barrier()
        qms = torch.tensor([q0.elapsed_time(q1) / args.steps], device="cuda")
        if world > 1:
            dist.all_reduce(qms, op=dist.ReduceOp.MAX)
        qms = float(qms.item())
        kq = float(np.mean(kq_ms))
        compact = {"value": C * T * world / (qms * 1e-3), "unit": UNIT, "ms_per_step": qms, "bytes_per_eval": 14,
                   "k1q_ms": kq, "k1q_GBps": 14.0 * C * T / (kq * 1e-3) / 1e9, "identical_to_fp32_layout": same,
                   "layout": "Form Q: 8 one-byte codebook indices + fp32 tool_success_rate + 2-byte presence index per evaluation, "
                             "lossless recoding of the same Form D tensor done once at load (apo_dims_generate_compact)"}
        engq.close()
    except Exception as ex:                         # the compact path is optional: never take the primary numbers down with it
        compact = {"error": str(ex)}

    # ---- end to end through the C ABI with HOST buffers (pinned), H2D inside the timed region
    Ce, Te = min(args.e2e_candidates, C), min(args.e2e_records, T)
    eng.set_stream(0)
    eng2 = pkg.Engine(local)                                # separate handle: keeps the resident workload intact
    eng2.dims_generate(SEED, 0, Ce, t0, Te, 300)            # device generator == oracle generator, bit for bit
    host = torch.empty((Ce, Te, 9), dtype=torch.float32, pin_memory=True)
    hnp = host.numpy()
    for c in range(Ce):
        hnp[c] = eng2.dims_download(c, 0, Te)
    eng2.corpus_generate(SEED, t0, Te, 300)
    hrec_t = torch.empty((Te * 32,), dtype=torch.uint8, pin_memory=True)
    hrec = hrec_t.numpy().view(pkg.RECORD_DTYPE)
    hrec[:] = eng2.corpus_download(0, Te)
    Ke = max(1, Ce // 4)

    def e2e_step():
        eng2.corpus_upload(hrec, idx_base=t0)
        return eng2.score_host(hnp, Ke, corpus=True, variant=args.variant, recip=bool(args.recip))

    for _ in range(2):
        e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        re = e2e_step()
    torch.cuda.synchronize()
    e_ms = (time.perf_counter() - e0) * 1e3 / args.e2e_steps
    ems = torch.tensor([e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e_ms = float(ems.item())
    d2h = 16 * Ce + 4 * Ke + 1024

    # ---- the same call with the evaluations as packed trace records (Form R16, 16 B/eval): dims are
    # derived on the device (TCS:668-763), 2.25x fewer bytes cross PCIe.  Reported beside the Form D number.
    eng2.rollouts16_generate(SEED, 0, Ce, t0, Te, 300)
    host16_t = torch.empty((Ce * Te * 16,), dtype=torch.uint8, pin_memory=True)
    host16 = host16_t.numpy().view(pkg.RECORD16_DTYPE).reshape(Ce, Te)
    for c in range(Ce):
        host16[c] = eng2.rollouts16_download(c, 0, Te)

    def e2e16_step():
        eng2.corpus_upload(hrec, idx_base=t0)
        return eng2.score_host_records(host16, Ke, corpus=True, variant=args.variant, recip=bool(args.recip))

    for _ in range(2):
        e2e16_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e16_step()
    torch.cuda.synchronize()
    e16_ms = (time.perf_counter() - e0) * 1e3 / args.e2e_steps
    ems = torch.tensor([e16_ms], device="cuda")
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e16_ms = float(ems.item())
    eng2.close()

    if rank == 0:
        peak, peak_src = read_peak()
        k1 = float(np.mean(k1_ms))
        fused = launches == args.steps * (1 if world == 1 else 3)            # the corpus scan rode inside the scoring launch
        alg_bytes = 36.0 * C * T + (32.0 * T if fused else 0.0)                # 36*C*T + 32*T = 92.48 GB at configs[2]
        achieved = alg_bytes / (k1 * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": C * T * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2]: {C}-beam x {T}-record Form D (36 B/eval) resident per GPU + {T}-record corpus; "
                                   f"K1 reward9 -> K2 detect6 + segmented sum + radix top-K (K={K})",
                       "C": C, "T_per_gpu": T, "T_global": T * world, "K": K, "parallelism": f"record-axis shards x{world}",
                       "l2": "inputs (>= 2.3 GB per step) exceed the 126 MB L2; no flush needed", "variant": args.variant,
                       "recip": args.recip, "note": note},
            "roofline": {"bound": "hbm", "kernel": "k_reward9 (K1, with the K2 corpus scan on its extra warp)" if fused else "k_reward9 (K1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": read_traffic(C, T), "peak_source": peak_src,
                         "alg_bytes_per_launch": alg_bytes, "k1_ms": k1, "k2_ms": float(np.mean(k2_ms)),
                         "join_ms": float(np.mean(ar_ms))},
            "e2e": {"value": Ce * Te * world / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": Ce * Te * 36 + Te * 32,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e_ms,
                    "workload": f"{Ce} x {Te} Form D + {Te}-record corpus from pinned host memory per rank via apo_corpus_upload + apo_score_host"},
            "e2e_records16": {"value": Ce * Te * world / (e16_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": Ce * Te * 16 + Te * 32,
                              "d2h_bytes_per_step": d2h, "ms_per_step": e16_ms,
                              "workload": f"{Ce} x {Te} packed trace records (Form R16, 16 B/eval) + corpus from pinned host memory per rank via apo_score_host_records"},
            "compact_layout": compact,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        out["e2e"]["host_placement"] = numa_note
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)
            out["cpu_baseline"] = cpu_baseline(len(os.sched_getaffinity(0)))
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
