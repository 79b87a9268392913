#!/usr/bin/env python3
"""bench.py -- headline benchmark of the decaf377 batch engine on B200.

Metric (BASELINE.json): decaf377 Pippenger MSM throughput in Mpoints/s.
Default workload: `vartime_multiscalar_mul` over 2^24 (scalar, Element) pairs
per GPU (BASELINE.json configs[3]/[4]); with --gpus N every rank owns its own
2^24-pair slice (weak scaling, 2^24 .. 2^27 points in total) and the 128-byte
partial sums are combined with one NCCL all-gather + N-1 point additions.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU port of the reference path, same metric

One JSON line on stdout (rank 0).  Other workloads (--workload encode |
fixed_base | pipeline | decompress | compress) time the remaining BASELINE
configs with the same harness.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FQ_OPS = {  # exact Fq multiplications + squarings per element (tools/count_ops.py);
    # products with a small curve constant (fq_mul_small, 16 wide multiplies) are not counted
    "isqrt": 308, "decompress": 320, "compress": 315, "encode": 323,
    "encode_compress": 338,   # fused: the encoding is read off the Jacobi-quartic pair (one isqrt)
    "hash_compress": 670,     # fused: two maps, the sum on the quartic, encoding from the sum
    # fixed base on the Jacobi quartic: 12 x (9 M + 2 S) + encoding tail 16 M + 4 S;
    # the Edwards path it replaces is 16 x 7 + 315 = 427
    "fixed_base_jq": 152,
    "scalar_mul": 3133, "pipeline": 320 + 3133 + 315,
}
IMAD_PER_FQ_OP = 128
# IMAD.WIDE.U32 actually issued per element (tools/count_ops.py): 120 per multiplication,
# 92 per squaring, 56 per from-Montgomery reduction.  `achieved` counts the reference's
# 128 per Fq-op; `issued_frac` is this count against the same peak, i.e. the
# share of the multiply pipe's issue slots the kernel really fills.
WIDE_ISSUED = {"decompress": 31612, "compress": 31140, "encode_compress": 33916, "hash_compress": 66932,
               "fixed_base_jq": 17760,
               # decompress + compress + the kernel's signed 4-bit ladder: 192 x (4S + 3M) + 64 x
               # (4S + 4M) doublings, 64 x 7M + 8M cached additions, 71 M for the table of 8 multiples
               "pipeline": 31612 + 31140 + 192 * 728 + 64 * 848 + 64 * 840 + 960 + 8520}
WIDE_PER_MUL = 120
# DRAM bytes (read + write) per launch from `ncu --set full` captures of the same
# configuration (profiles/); None where no capture of that configuration is committed.
NCU_TRAFFIC = {
    # k_msm_accumulate<1>, 2^24 pairs, c = 18: 32.54 GB read + 0.46 GB written over all
    # launches of one MSM (profiles/r1_ncu_full_summary.csv, capture msm24_one_group; the
    # three group launches of msm24_pipelined add up to the same figure)
    ("msm", 24, True): 33.00e9,
    # codec kernels: per-element DRAM bytes of the 2^20 captures (codec20) x n
    ("compress", 22): 4 * 157.9e6, ("decompress", 22): 4 * 116.3e6, ("encode", 22): 4 * 34.8e6,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msm",
                    choices=["msm", "encode", "hash", "fixed_base", "pipeline", "decompress", "compress"])
    ap.add_argument("--logn", type=int, default=None, help="log2 of units per GPU")
    ap.add_argument("--ref-logn", type=int, default=None, help="log2 of the CPU sample size")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


DEFAULT_LOGN = {"msm": 24, "encode": 22, "hash": 22, "fixed_base": 24, "pipeline": 16, "decompress": 22,
                "compress": 22}
DEFAULT_REF_LOGN = {"msm": 18, "encode": 17, "hash": 16, "fixed_base": 14, "pipeline": 14, "decompress": 17,
                    "compress": 17}
UNIT = {"msm": "Mpoints/s"}
METRIC = {
    "msm": "decaf377 MSM throughput (vartime_multiscalar_mul, Pippenger)",
    "encode": "decaf377 batch encode_to_curve + vartime_compress throughput",
    "hash": "decaf377 batch hash_to_curve + vartime_compress throughput",
    "fixed_base": "decaf377 fixed-base (generator) scalar mul + compress throughput",
    "pipeline": "decaf377 vartime_decompress -> scalar mul -> vartime_compress throughput",
    "decompress": "decaf377 batch vartime_decompress throughput",
    "compress": "decaf377 batch vartime_compress throughput",
}
WORKLOAD_NAME = {
    "msm": "Pippenger vartime_multiscalar_mul, 2^{logn} (Fr, Element) pairs per GPU",
    "encode": "batch Elligator encode_to_curve + compress of 2^{logn} Fq elements per GPU",
    "hash": "batch hash_to_curve (two Elligator maps + add) + compress of 2^{logn} Fq pairs per GPU",
    "fixed_base": "fixed-base generator mul + compress of 2^{logn} Fr scalars per GPU",
    "pipeline": "2^{logn} encodings: vartime_decompress -> scalar mul -> vartime_compress",
    "decompress": "batch vartime_decompress of 2^{logn} encodings per GPU",
    "compress": "batch vartime_compress of 2^{logn} elements per GPU",
}


# ---------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "_source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------
# CPU port of the reference path (oracle) -- cpu_baseline and --impl reference
# ---------------------------------------------------------------------------
def cpu_inputs(workload: str, n: int):
    import numpy as np
    from oracle import c_oracle as co
    from oracle import decaf377_ref as o
    threads = os.cpu_count() or 1
    raw = np.frombuffer(o.xof_bytes("bench_fq", n), np.uint8).reshape(n, 32).copy()
    sc = np.frombuffer(o.xof_bytes("bench_sc", n), np.uint8).reshape(n, 32).copy()
    sc[:, 31] &= 0x03
    if workload == "msm":
        return (sc, co.encode_to_curve(raw, threads=threads))
    if workload == "encode":
        return (raw,)
    if workload == "hash":
        raw2 = np.frombuffer(o.xof_bytes("bench_fq2", n), np.uint8).reshape(n, 32).copy()
        return (raw, raw2)
    if workload == "fixed_base":
        return (sc,)
    el = co.encode_to_curve(raw, threads=threads)
    if workload == "compress":
        return (el,)
    enc = co.compress(el, threads=threads)
    if workload == "decompress":
        return (enc,)
    return (enc, sc)


def cpu_step(workload: str, inputs, threads: int):
    from oracle import c_oracle as co
    if workload == "msm":
        return co.msm_pippenger(inputs[0], inputs[1], threads=threads)
    if workload == "encode":
        return co.encode_to_curve(inputs[0], out_enc=True, threads=threads)
    if workload == "hash":
        return co.hash_to_curve(inputs[0], inputs[1], out_enc=True, threads=threads)
    if workload == "fixed_base":
        return co.fixed_base(inputs[0], out_enc=True, threads=threads)
    if workload == "compress":
        return co.compress(inputs[0], threads=threads)
    if workload == "decompress":
        return co.decompress(inputs[0], threads=threads)
    return co.pipeline(inputs[0], inputs[1], threads=threads)


CPU_KIND_NOTE = ("C port of the reference algorithms (oracle/d377_oracle.c: 4x64 Montgomery, "
                 "Sarkar sqrt, ark-ec style Pippenger, pthreads); the Rust crate cannot be built here")


def run_cpu(workload: str, logn: int, steps: int, warmup: int):
    threads = os.cpu_count() or 1
    n = 1 << logn
    inputs = cpu_inputs(workload, n)
    for _ in range(min(warmup, 1)):
        cpu_step(workload, inputs, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(workload, inputs, threads)
    dt = time.perf_counter() - t0
    return n * steps / dt / 1e6, dt / steps * 1e3, threads, n


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    logn = args.ref_logn or DEFAULT_REF_LOGN[wl]
    steps = max(1, min(args.steps, 5))
    value, ms, threads, n = run_cpu(wl, logn, steps, args.warmup)
    unit = UNIT.get(wl, "Melem/s")
    sample = "%d steps of 2^%d units on %d host threads; %s" % (steps, logn, threads, CPU_KIND_NOTE)
    line = {
        "impl": "reference", "metric": METRIC[wl], "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME[wl].format(logn=args.logn or DEFAULT_LOGN[wl]),
                   "sample": "2^%d units per step" % logn},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import decaf377_b200 as d
    from decaf377_b200 import device as dev
    from decaf377_b200 import dist as ddist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: decaf377_b200 has no CPU fallback (use --impl reference "
                         "for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    d.init(local_rank)
    st = dev.engine_stream()
    wl = args.workload
    logn = args.logn or DEFAULT_LOGN[wl]
    n = 1 << logn
    cuda = torch.device("cuda", local_rank)

    # ---- synthetic inputs, generated on the device (seed differs per rank) ----
    g = torch.Generator(device=cuda).manual_seed(377 + rank)
    raw = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=cuda, generator=g)
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=cuda, generator=g)
    sc[:, 31] &= 0x03           # < 2^250 < r: canonical Fr
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    if wl == "msm":
        pts = dev.encode_to_curve(raw, d.OUT_ELEMENT)   # Element wire format, 128 B
        d.sync()
        del raw
        h2d, d2h = n * (32 + 128), 160

        def step_dev():
            if world == 1:
                return dev.msm(sc, pts, d.PT_ELEMENT, want_encoding=True)
            return ddist.msm_sharded(sc, pts, d.PT_ELEMENT)

        host_in = None

        def make_host():
            return (d.pinned_copy(sc.cpu().numpy()), d.pinned_copy(pts.cpu().numpy()))

        make_out = lambda: ()

        def step_e2e(h, o):
            res = d.vartime_multiscalar_mul(h[0], h[1], d.PT_ELEMENT)
            if world == 1:
                return res
            part = torch.from_numpy(res[0]).to(cuda)
            gathered = ddist.gather_partials(part)
            oe, oc = dev.element_sum(gathered)
            d.sync()
            return oe.cpu(), oc.cpu()
    else:
        el = dev.encode_to_curve(raw, d.OUT_ELEMENT) if wl in ("compress", "decompress", "pipeline") else None
        enc = dev.compress(el) if wl in ("decompress", "pipeline") else None
        d.sync()
        # e2e: pinned inputs AND pinned result buffers (d377_host_alloc), so the chunked
        # host API overlaps upload, kernel and download
        if wl == "encode":
            ins, h2d, d2h = (raw,), n * 32, n * 32
            step_dev = lambda: dev.encode_to_curve(raw, d.OUT_ENCODING)
            make_out = lambda: (d.pinned_empty((n, 32)),)
            step_e2e = lambda h, o: d.batch_encode_to_curve(h[0], d.OUT_ENCODING, out=o[0])
        elif wl == "hash":
            raw2 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=cuda, generator=g)
            ins, h2d, d2h = (raw, raw2), n * 64, n * 32
            step_dev = lambda: dev.hash_to_curve(raw, raw2, d.OUT_ENCODING)
            make_out = lambda: (d.pinned_empty((n, 32)),)
            step_e2e = lambda h, o: d.batch_hash_to_curve(h[0], h[1], d.OUT_ENCODING, out=o[0])
        elif wl == "fixed_base":
            ins, h2d, d2h = (sc,), n * 32, n * 32
            step_dev = lambda: dev.fixed_base_mul(sc, d.OUT_ENCODING)
            make_out = lambda: (d.pinned_empty((n, 32)),)
            step_e2e = lambda h, o: d.fixed_base_mul(h[0], d.OUT_ENCODING, out=o[0])
        elif wl == "compress":
            ins, h2d, d2h = (el,), n * 128, n * 32
            step_dev = lambda: dev.compress(el)
            make_out = lambda: (d.pinned_empty((n, 32)),)
            step_e2e = lambda h, o: d.batch_compress(h[0], out=o[0])
        elif wl == "decompress":
            ins, h2d, d2h = (enc,), n * 32, n * 129
            step_dev = lambda: dev.decompress(enc)
            make_out = lambda: (d.pinned_empty((n, 128)), d.pinned_empty((n,)))
            step_e2e = lambda h, o: d.batch_decompress(h[0], out=o[0], ok=o[1])
        else:
            ins, h2d, d2h = (enc, sc), n * 64, n * 33
            step_dev = lambda: dev.scalar_mul(enc, sc, d.PT_ENCODING, d.OUT_ENCODING)
            make_out = lambda: (d.pinned_empty((n, 32)), d.pinned_empty((n,)))
            step_e2e = lambda h, o: d.batch_scalar_mul(h[0], h[1], d.PT_ENCODING, d.OUT_ENCODING,
                                                       return_ok=True, out=o[0], ok=o[1])
        make_host = lambda: tuple(d.pinned_copy(t.cpu().numpy()) for t in ins)

    # ---- device-resident timing ---------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    d.sync()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = d.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0.record()
    for _ in range(args.steps):
        step_dev()
    with torch.cuda.stream(st):
        if world > 1:
            st.wait_stream(torch.cuda.current_stream())
        e1.record()
    d.sync()
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = d.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=cuda, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6

    # ---- strong scaling beside it (N > 1, MSM): ONE 2^24-pair MSM cut into N slices ----
    # BASELINE.json's configs[4] / north star: "MSM at 2^24 points sharded across 8 GPUs".
    # `value` above is weak scaling (2^24 pairs per GPU); this is the same call on the first
    # 2^24 / N pairs of every rank, timed the same way (CUDA events, max over ranks).
    strong = None
    if wl == "msm" and world > 1 and logn == 24:
        ns = n // world
        sc_s, pts_s = sc[:ns], pts[:ns]
        for _ in range(3):
            ddist.msm_sharded(sc_s, pts_s, d.PT_ELEMENT)
        d.sync()
        torch.cuda.synchronize()
        barrier()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            s0.record()
        for _ in range(args.steps):
            ddist.msm_sharded(sc_s, pts_s, d.PT_ELEMENT)
        with torch.cuda.stream(st):
            st.wait_stream(torch.cuda.current_stream())
            s1.record()
        d.sync()
        torch.cuda.synchronize()
        barrier()
        t = torch.tensor([s0.elapsed_time(s1)], device=cuda, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_s = float(t.item()) / args.steps
        sinfo = d.msm_stage_info()
        s_adds = ns * sinfo["W"] * (7 if sinfo["mixed"] else 8) * IMAD_PER_FQ_OP
        strong = {"scaling": "strong", "total_units": n, "units_per_gpu": ns, "ms_per_step": ms_s,
                  "value": n / (ms_s * 1e-3) / 1e6, "unit": "Mpoints/s", "window_c": sinfo["c"],
                  "accumulate_ms_rank0": sinfo["ms"]["accumulate"],
                  "accumulate_imad_frac_rank0":
                      s_adds / (sinfo["ms"]["accumulate"] * 1e-3) / 1e9 / d.imad_peak()}
        # the weak-scaling stage info below must describe the full-size call again
        ddist.msm_sharded(sc, pts, d.PT_ELEMENT)
        d.sync()

    # ---- roofline of the dominant kernel, measured live ---------------------------
    imad_peak = d.imad_peak()            # G IMAD.WIDE.U32 / s on this GPU, just measured
    peaks = measured_peaks()
    roofline, roofline_hbm, stages = None, None, None
    if wl == "msm":
        info = d.msm_stage_info()
        stages = {k: round(v, 4) for k, v in info["ms"].items()}
        acc_ms = info["ms"]["accumulate"]
        adds = n * info["W"]                      # one bucket addition per non-zero digit
        # Fq multiplications per bucket addition: 7 for a mixed addition against an affine
        # point (SURVEY 8d "A = 7 (mixed)"), 8 against a cached projective point
        per_add = 7 if info["mixed"] else 8
        rec = 128                                  # one cache line per gathered operand
        imads = adds * per_add * IMAD_PER_FQ_OP
        ach = imads / (acc_ms * 1e-3) / 1e9
        issued = adds * per_add * WIDE_PER_MUL / (acc_ms * 1e-3) / 1e9
        bytes_alg = adds * (rec + 4) + (n * info["W"] / 32) * 128
        ach_bw = bytes_alg / (acc_ms * 1e-3) / 1e9
        traffic = NCU_TRAFFIC.get(("msm", logn, info["mixed"]))
        roofline = {"bound": "imad", "kernel": "k_msm_accumulate", "achieved": ach, "peak": imad_peak,
                    "unit": "GIMAD/s (32x32->64 multiply-adds)", "frac": ach / imad_peak,
                    "issued_frac": issued / imad_peak, "fq_mults_per_bucket_addition": per_add,
                    "traffic": traffic, "algorithmic_bytes": bytes_alg,
                    "launch_ms": acc_ms, "window_c": info["c"], "windows": info["W"],
                    "peak_source": "IMAD.WIDE.U32 issue-rate microbenchmark run in this process"}
        roofline_hbm = {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": ach_bw,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_bw / peaks["hbm_gbs"],
                        "traffic": traffic, "peak_source": peaks["_source"]}
    else:
        ops = {"encode": FQ_OPS["encode_compress"], "hash": FQ_OPS["hash_compress"],
               "fixed_base": FQ_OPS["fixed_base_jq"],
               "compress": FQ_OPS["compress"], "decompress": FQ_OPS["decompress"],
               "pipeline": FQ_OPS["pipeline"]}[wl]
        ach = n * ops * IMAD_PER_FQ_OP / (ms_step * 1e-3) / 1e9
        ach_bw = (h2d + d2h) / (ms_step * 1e-3) / 1e9
        wide = {"encode": WIDE_ISSUED["encode_compress"], "hash": WIDE_ISSUED["hash_compress"],
                "fixed_base": WIDE_ISSUED["fixed_base_jq"],
                "compress": WIDE_ISSUED["compress"], "decompress": WIDE_ISSUED["decompress"],
                "pipeline": WIDE_ISSUED["pipeline"]}.get(wl)
        issued_frac = (n * wide / (ms_step * 1e-3) / 1e9 / imad_peak) if wide else None
        roofline = {"bound": "imad", "kernel": wl, "achieved": ach, "peak": imad_peak,
                    "unit": "GIMAD/s (32x32->64 multiply-adds)", "frac": ach / imad_peak,
                    "issued_frac": issued_frac,
                    "traffic": NCU_TRAFFIC.get((wl, logn)), "launch_ms": ms_step,
                    "peak_source": "IMAD.WIDE.U32 issue-rate microbenchmark run in this process"}
        roofline_hbm = {"bound": "hbm", "kernel": wl, "achieved": ach_bw, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": ach_bw / peaks["hbm_gbs"], "traffic": None,
                        "peak_source": peaks["_source"]}

    # ---- end to end through the host-buffer C ABI ------------------------------------
    # Every step copies that step's inputs from pinned host memory to the device and reads
    # the result back.  `e2e` is the pipelined form a throughput-oriented caller uses
    # (d377_msm_submit / d377_msm_wait, two slots: the upload of step i+1 overlaps the
    # MSM of step i); `e2e_sync` is the plain blocking call.
    e2e, e2e_sync, e2e_affine, e2e_element, e2e_bases = None, None, None, None, None
    if not args.no_e2e:
        host = make_host()
        outb = make_out()
        e2e_steps = max(2, min(args.steps, 10))

        def finish_step(res):
            if world == 1 or wl != "msm":
                return res
            part = torch.from_numpy(res[0]).to(cuda)
            gathered = ddist.gather_partials(part)
            oe, oc = dev.element_sum(gathered)
            d.sync()
            return oe.cpu(), oc.cpu()

        def timed(fn):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=cuda, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        def run_sync():
            for _ in range(e2e_steps):
                step_e2e(host, outb)

        def run_pipelined():
            d.msm_submit(host[0], host[1], d.PT_ELEMENT, slot=0)
            for i in range(1, e2e_steps):
                d.msm_submit(host[0], host[1], d.PT_ELEMENT, slot=i & 1)
                finish_step(d.msm_wait((i - 1) & 1))
            finish_step(d.msm_wait((e2e_steps - 1) & 1))

        step_e2e(host, outb)           # warm the staging buffers
        dt = timed(run_sync)
        e2e_sync = {"value": world * n * e2e_steps / dt / 1e6, "unit": UNIT.get(wl, "Melem/s"),
                    "ms_per_step": dt / e2e_steps * 1e3}
        api = "host-buffer C ABI, blocking call per step, pinned in/out buffers, chunk-pipelined"
        if wl == "msm":
            run_pipelined()            # warm both slots
            dt = timed(run_pipelined)
            api = "d377_msm_submit/d377_msm_wait, 2 slots, pinned host buffers"
        e2e = {"value": world * n * e2e_steps / dt / 1e6, "unit": UNIT.get(wl, "Melem/s"),
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3, "api": api}
        if wl == "msm":
            # The Element wire image above is X||Y||Z||T (128 B).  The Rust shim copies the
            # coordinates of every Element one by one (Projective is not repr(C)), and T = XY/Z
            # is redundant, so the shim's natural wire image is X||Y||Z (D377_PT_XYZ, 96 B):
            # 128 instead of 160 bytes per pair over a link that is the bound of this call.
            # That is the end-to-end path a caller of vartime_multiscalar_mul gets; the 128-byte
            # image is kept beside it as `e2e_element`.
            e2e_element = e2e
            xyz = pts.cpu().numpy()[:, :96].copy()
            host_x = (host[0], d.pinned_copy(xyz))
            del xyz

            def run_pipelined_xyz():
                d.msm_submit(host_x[0], host_x[1], d.PT_XYZ, slot=0)
                for i in range(1, e2e_steps):
                    d.msm_submit(host_x[0], host_x[1], d.PT_XYZ, slot=i & 1)
                    finish_step(d.msm_wait((i - 1) & 1))
                return finish_step(d.msm_wait((e2e_steps - 1) & 1))

            res_x = run_pipelined_xyz()
            dt = timed(run_pipelined_xyz)
            same = None
            if world == 1:
                same = bytes(res_x[1].tobytes()) == dev.msm(sc, pts)[1].cpu().numpy().tobytes()
            e2e = {"value": world * n * e2e_steps / dt / 1e6, "unit": "Mpoints/s",
                   "h2d_bytes_per_step": n * 128, "d2h_bytes_per_step": 160, "steps": e2e_steps,
                   "ms_per_step": dt / e2e_steps * 1e3,
                   "api": "d377_msm_submit/d377_msm_wait, 2 slots, pinned host buffers, "
                          "Element wire image X||Y||Z (D377_PT_XYZ, 96 B)",
                   "same_result_as_element_input": same}
            del host_x
        del host, outb
        if wl == "msm" and world == 1:
            # Same MSM fed with AffinePoint bases (64 B), the input type of the reference's
            # VariableBaseMSM::msm (ark_curve/element.rs:22-37): 96 B per pair over PCIe
            # instead of 160 B, which moves the e2e bound from the link to the kernels.
            aff = dev.normalize(pts)
            d.sync()
            host_a = (d.pinned_copy(sc.cpu().numpy()), d.pinned_copy(aff.cpu().numpy()))
            del aff

            def run_pipelined_affine():
                d.msm_submit(host_a[0], host_a[1], d.PT_AFFINE, slot=0)
                for i in range(1, e2e_steps):
                    d.msm_submit(host_a[0], host_a[1], d.PT_AFFINE, slot=i & 1)
                    d.msm_wait((i - 1) & 1)
                return d.msm_wait((e2e_steps - 1) & 1)

            res_a = run_pipelined_affine()
            dt = timed(run_pipelined_affine)
            e2e_affine = {"value": n * e2e_steps / dt / 1e6, "unit": "Mpoints/s",
                          "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": 160,
                          "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                          "api": "d377_msm_submit/d377_msm_wait, D377_PT_AFFINE bases",
                          "same_result_as_element_input":
                              res_a[1].tobytes() == dev.msm(sc, pts)[1].cpu().numpy().tobytes()}
            del host_a
            # VariableBaseMSM::msm with LONG-LIVED bases (batch_convert_to_mul_base once,
            # ark_curve/element.rs:27-37): d377_msm_bases_create uploads and normalises the
            # bases once, untimed; every step then moves only the 32-byte scalars.
            bases = d.MsmBases(device_ptr=pts.data_ptr(), n=n, point_format=d.PT_ELEMENT)
            host_s = d.pinned_copy(sc.cpu().numpy())

            def run_pipelined_bases():
                d.msm_submit(host_s, bases, slot=0)
                for i in range(1, e2e_steps):
                    d.msm_submit(host_s, bases, slot=i & 1)
                    d.msm_wait((i - 1) & 1)
                return d.msm_wait((e2e_steps - 1) & 1)

            res_b = run_pipelined_bases()
            dt = timed(run_pipelined_bases)
            for _ in range(2):
                dev.msm(sc, bases)
            d.sync()
            b0 = torch.cuda.Event(enable_timing=True)
            b1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                b0.record()
            for _ in range(args.steps):
                dev.msm(sc, bases)
            with torch.cuda.stream(st):
                b1.record()
            d.sync()
            b1.synchronize()
            ms_b = b0.elapsed_time(b1) / args.steps
            e2e_bases = {"value": n * e2e_steps / dt / 1e6, "unit": "Mpoints/s",
                         "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 160,
                         "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
                         "device_resident_value": n / (ms_b * 1e-3) / 1e6,
                         "device_resident_ms_per_step": ms_b,
                         "api": "d377_msm_bases_create once (untimed), then d377_msm_submit/"
                                "d377_msm_wait with D377_PT_BASES",
                         "same_result_as_element_input":
                             res_b[1].tobytes() == dev.msm(sc, pts)[1].cpu().numpy().tobytes()}
            bases.close()
            del host_s

    # ---- correctness spot check against the oracle (untimed) ---------------------------
    verified = None
    if rank == 0:
        try:
            from oracle import c_oracle as co
            m = 2048
            if wl == "msm":
                s_h, p_h = sc[:m].cpu().numpy(), pts[:m].cpu().numpy()
                got = d.vartime_multiscalar_mul(s_h, p_h)[1].tobytes()
                verified = got == co.msm_pippenger(s_h, p_h, threads=os.cpu_count() or 1)[1].tobytes()
            elif wl == "encode":
                r_h = raw[:m].cpu().numpy()
                verified = bool((d.batch_encode_to_curve(r_h, d.OUT_ENCODING)
                                 == co.encode_to_curve(r_h, out_enc=True, threads=8)).all())
            elif wl == "hash":
                r_h, r2_h = raw[:m].cpu().numpy(), raw2[:m].cpu().numpy()
                verified = bool((d.batch_hash_to_curve(r_h, r2_h, d.OUT_ENCODING)
                                 == co.hash_to_curve(r_h, r2_h, out_enc=True, threads=8)).all())
            elif wl == "fixed_base":
                s_h = sc[:256].cpu().numpy()
                verified = bool((d.fixed_base_mul(s_h, d.OUT_ENCODING)
                                 == co.fixed_base(s_h, threads=8)).all())
            elif wl == "compress":
                e_h = el[:m].cpu().numpy()
                verified = bool((d.batch_compress(e_h) == co.compress(e_h, threads=8)).all())
            elif wl == "decompress":
                e_h = enc[:m].cpu().numpy()
                a, b = d.batch_decompress(e_h), co.decompress(e_h, threads=8)
                verified = bool((a[1] == b[1]).all() and (d.batch_compress(a[0]) == e_h).all())
            else:
                e_h, s_h = enc[:256].cpu().numpy(), sc[:256].cpu().numpy()
                verified = bool((d.batch_scalar_mul(e_h, s_h, d.PT_ENCODING, d.OUT_ENCODING)
                                 == co.pipeline(e_h, s_h, threads=8)[0]).all())
        except Exception as ex:  # the check is informative; never fail the measurement
            verified = "check failed to run: %r" % (ex,)

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_logn = args.ref_logn or DEFAULT_REF_LOGN[wl]
        v, ms, threads, nn = run_cpu(wl, ref_logn, 3, 1)
        cpu = {"value": v, "unit": UNIT.get(wl, "Melem/s"), "cores": threads, "kind": "port",
               "sample": "3 steps of 2^%d units on %d host threads (%.0f ms/step); %s"
                         % (ref_logn, threads, ms, CPU_KIND_NOTE)}

    if rank == 0:
        line = {
            "metric": METRIC[wl], "value": value, "unit": UNIT.get(wl, "Melem/s"),
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (8x32-bit limb Montgomery, IMAD.WIDE)",
            "data": "synthetic (torch CUDA RNG bytes; points = encode_to_curve of random Fq)",
            "config": {"workload": WORKLOAD_NAME[wl].format(logn=logn), "units_per_gpu": n,
                       "total_units": world * n, "point_format": "Element X||Y||Z||T 128 B" if wl == "msm" else None,
                       "parallelism": "point-slice sharding x%d, 128 B all-gather" % world if world > 1 else "single GPU",
                       "l2": "inputs (%.0f MiB per GPU) exceed the 126 MB L2" % ((h2d) / 2**20)},
            "roofline": roofline, "roofline_hbm": roofline_hbm, "e2e_element": e2e_element,
            "cpu_baseline": cpu, "e2e": e2e, "e2e_sync": e2e_sync, "e2e_affine": e2e_affine,
            "e2e_prepared_bases": e2e_bases,
            "gpu_launches": int(launches),
            "clocks": clocks, "verified_vs_oracle": verified,
        }
        if stages:
            line["msm_stage_ms"] = stages
        if strong:
            line["strong_scaling_2p24"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
