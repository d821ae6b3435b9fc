#!/usr/bin/env python3
"""bench.py -- sfft_exec throughput on B200 (BASELINE.json metric), one JSON line.

A "step" is one sparse-FFT transform of one synthetic k-sparse signal per GPU (one
pass of the hot path: Comb pre-filter, permuted windowed gather, bucket FFTs,
top-2k selection, voting, median estimation).  Default workload = BASELINE.json
configs[1]: sFFT v2, n = 2^24, k = 1000, exact k-sparse, 1 B200.

  value : Gsamples/s = (signals * n) / time, inputs resident in HBM, result left as a
          sparse (loc, val) list in HBM; CUDA events on the launching stream.
  e2e   : same metric through the legacy C-ABI call sfft_exec(plan, in, out) with
          HOST buffers: H2D of the signal + transform + dense out + D2H inside the
          timed region.
  roofline / cpu_baseline : see DESIGN.md "Measurement".

N > 1 (torchrun): every rank transforms its own signals (sfft_exec_many-style
partition, no data-path collective), weak scaling, max-over-ranks timing.

`--impl reference` times the reference's own CPU implementation (oracle/_ref: the
unmodified reference sources compiled over the FFTW shim) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (version, n, k, noisy_snr_db or None, description)
    "C1": (1, 1 << 22, 50, None, "sFFT v1 exact k-sparse n=2^22 k=50"),
    "C2": (2, 1 << 24, 1000, None, "sFFT v2 (Comb) exact k-sparse n=2^24 k=1000"),
    # the nearest v2 point the as-shipped (asserts on) reference can run
    "C2b": (2, 1 << 24, 50, None, "sFFT v2 (Comb) exact k-sparse n=2^24 k=50"),
    "C3": (3, 1 << 26, 2000, None, "sFFT v3 exact-sparse n=2^26 k=2000"),
    "C4": (1, 1 << 27, 500, 20.0, "sFFT v1 noisy 20 dB n=2^27 k=500"),
    "C5": (1, 1 << 20, 100, None, "sFFT v1 n=2^20 k=100, sfft_exec_many batch of 256 signals per step"),
}
# signals transformed per step and per GPU (sfft_exec_many); BASELINE config 5 is a batch of
# 4096 = 64 GiB of input, benchmarked here 256 signals (4 GiB) at a time
BATCH = {"C5": 256}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------- #
# synthetic inputs: the SAME bits for this engine and for the reference arm
# --------------------------------------------------------------------------- #
def host_signal(n, k, seed, snr_db=None):
    """k unit spikes at random locations -> x = unnormalised inverse DFT (the reference's
    generator, src/simulation.cc:104-111, with numpy's RNG/FFT; synthesis is outside every
    timed region).  Noisy: complex AWGN of the reference's model (src/utils.cc:263-273),
    std = sqrt(k / (2 * 10^(SNR/10))).  Deterministic in (n, k, seed): both arms of the
    bench, in different processes, transform identical input bits."""
    import numpy as np
    rng = np.random.default_rng(seed)
    loc = rng.integers(0, n, k)
    xf = np.zeros(n, dtype=np.complex128)
    xf[loc] = 1.0
    x = np.fft.ifft(xf) * n
    if snr_db is not None:
        std = (k / (2.0 * 10 ** (snr_db / 10.0))) ** 0.5
        u = np.maximum(rng.random(n), 1e-300)
        v = rng.random(n)
        x = x + std * np.sqrt(-2 * np.log(u)) * np.exp(2j * np.pi * v)
    return np.ascontiguousarray(x)


SIGNAL_SEED = 1000          # signal i of a workload uses seed SIGNAL_SEED + i (rank r: + 100000 r)


# --------------------------------------------------------------------------- #
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------- #
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
# reference arm: the reference's CPU path on the host cores
# --------------------------------------------------------------------------- #
def cpu_reference_run(workload, nsig, steps, warmup, budget_s=150.0):
    """Times sfft_exec / sfft_exec_many of the compiled reference (oracle/_ref fast
    build = the reference's own optimisation flags).  Returns a dict."""
    import numpy as np
    from oracle import ref

    version, n, k, snr_db, desc = WORKLOADS[workload]
    # timing build: reference flags over MKL DFTI when this image's libtorch_cpu.so provides it
    # (validated against numpy.fft by tests/test_oracle_vs_ref.py), else over the oracle's own FFT
    kind = "mkl" if ref.available("mkl") and ref.mkl_provider() else ("fast" if ref.available("fast") else "parity")
    if not ref.available(kind):
        return {"unavailable": "oracle/_ref not built (needs the reference sources at build time)"}
    fft_backend = {"mkl": "MKL DFTI from libtorch_cpu.so (not FFTW: FFTW is not installable here)",
                   "fast": "the oracle's radix-2/Bluestein FFT shim (FFTW is not installable here)",
                   "parity": "the oracle's radix-2/Bluestein FFT shim, IEEE flags"}[kind]
    cores = min(nsig, os.cpu_count() or 1)
    L = ref.lib(kind)
    t0 = time.time()
    plan = ref.RefPlan(n, k, version, kind=kind, threads=cores)
    plan_s = time.time() - t0
    # the very signals the GPU arm transforms (rank 0's first `nsig`)
    xs = [host_signal(n, k, SIGNAL_SEED + i, snr_db) for i in range(nsig)]

    def one():
        plan.seed(17, 12345)        # the GPU arm (rank 0) reseeds with the same pair before every step
        t = time.perf_counter()
        if nsig == 1:
            plan.exec(xs[0])
        else:
            plan.exec_many(xs)
        return time.perf_counter() - t

    first = one()                       # also serves as the only warm-up we can afford
    per_step = first
    fit = int(max(1, min(steps, (budget_s - first) // max(per_step, 1e-9))))
    times = []
    if warmup <= 0 or first * (fit + 1) > budget_s:
        times.append(first)
        fit -= 1
    for _ in range(max(0, fit)):
        times.append(one())
    total = sum(times)
    value = nsig * n * len(times) / total / 1e9
    return {
        "value": value, "unit": "Gsamples/s", "cores": cores, "kind": "reference",
        "sample": (f"{len(times)} full sfft_exec{'_many' if nsig > 1 else ''} call(s) of {desc}, "
                   f"{nsig} signal(s), same input bits as the GPU arm; reference sources built -O3 -ffast-math "
                   f"-march=x86-64-v3 (not -march=native: the binary is built on another host) -fopenmp "
                   f"-DNDEBUG over {fft_backend}; plan build {plan_s:.1f}s excluded"),
        "steps_timed": len(times), "ms_per_step": 1e3 * total / len(times),
        "cpu_model": _cpu_model(),
    }


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline_leg(workload, nsig):
    """The cpu_baseline object of the GPU arm's line, taken in a CHILD process: the reference's v3
    path writes one element past a heap buffer (computefourier-3.0.cc:235) and can abort the
    process in free(); that must not sink the bench line."""
    try:
        env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="32768")
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                              "--steps", "1", "--warmup", "0", "--cpu-leg-signals", str(nsig), "--cpu-leg-budget", "90"],
                             capture_output=True, text=True, timeout=600, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                return d.get("cpu_baseline", d)
        return {"unavailable": "the reference process died (rc %d): %s" % (out.returncode, out.stderr.strip()[-200:])}
    except Exception as e:     # noqa: BLE001
        return {"unavailable": repr(e)}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    nsig = args.cpu_leg_signals or args.gpus * min(BATCH.get(args.workload, 1), os.cpu_count() or 1)
    r = cpu_reference_run(args.workload, nsig, args.steps, args.warmup, budget_s=args.cpu_leg_budget)
    version, n, k, snr_db, desc = WORKLOADS[args.workload]
    if "unavailable" in r:
        print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}))
        return
    line = {
        "metric": "sfft_exec throughput", "value": r["value"], "unit": "Gsamples/s",
        "n_gpus": args.gpus, "steps": r["steps_timed"], "steps_requested": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": f"{args.workload}: {desc}", "signals_per_step": args.gpus,
                   "n": n, "k": k, "version": version},
        "cpu_baseline": {"value": r["value"], "unit": "Gsamples/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"], "cpu_model": r["cpu_model"],
                         "ms_per_step": r["ms_per_step"], "steps_timed": r["steps_timed"]},
        "e2e": {"value": r["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
def stage_bytes(info, stage, count):
    """Algorithmic bytes of one stage per transform (DESIGN.md 'Measurement'): what the stage
    must read and write once, 16 B per complex double, 4 B per index."""
    n, num = info["n"], 2 * info["k"]
    xs = info["x_samp_size"]
    if info["version"] == 3:
        W, B1, B2 = info["W_Man"], info["B_g1"], info["B_g2"]
        slots = 2 * (W + B1 + B2)
        if stage == "bucketise":      # samples + taps read, bucket arrays written
            return 16 * info["gather_samples"] + info["gather_tap_bytes"] + 16 * slots
        if stage == "bucket_fft":
            return 2 * 16 * slots
        if stage == "peel":           # every bucket read once + the recovered list written
            return 16 * slots + 20 * count
        if stage == "stage_draws":
            return 32
        return 0
    loops = info["loops_loc"] + info["loops_est"]
    if stage == "gather":
        return 16 * (info["gather_samples"] - info["Comb_loops"] * info["W_Comb"] * (info["version"] == 2)) \
            + info["gather_tap_bytes"] + 16 * xs
    if stage == "estimate":
        # read the bucket spectra once, write (loc:4 B, val:16 B) per recovered coefficient
        return 16 * xs + 20 * count
    if stage == "bucket_fft":
        return 2 * 16 * xs
    if stage == "select":             # location rows read, J + bitmap written
        return 16 * info["loops_loc"] * info["B_loc"] + info["loops_loc"] * (4 * num + info["B_loc"] // 8)
    if stage == "vote":               # J + bitmaps read, voted list written (upper bound: result count)
        return info["loops_loc"] * (4 * num + info["B_loc"] // 8) + 4 * count
    if stage == "comb":               # W samples read + spectrum written/read + approved list
        W = info["W_Comb"] * info["Comb_loops"]
        return 16 * W * 3 + 4 * num
    if stage == "stage_draws":
        return 4 * (2 * loops + info["Comb_loops"])
    if stage == "exchange":
        return 16 * xs
    return 0


# --------------------------------------------------------------------------- #
# extras: the other BASELINE configs, at whatever N this run has
# --------------------------------------------------------------------------- #
def guarded(fn, *a):
    """An extra must not sink the bench line."""
    try:
        return fn(*a)
    except Exception as e:       # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"error": repr(e)}


def timed_ms(ctx, fn, reps, warm=3):
    """ms per call of fn(): CUDA events on the engine's stream, barrier + synchronize on both
    sides, max over ranks."""
    torch, dist = ctx["torch"], ctx["dist"]
    for i in range(warm):
        fn(i)
    ctx["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    ctx["barrier"]()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def device_signal(torch, n, k, seed, snr_db, dev):
    """Device-side synthesis for the extras (never compared with the CPU arm)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    loc = torch.randint(0, n, (k,), generator=g)
    xf = torch.zeros(n, dtype=torch.complex128, device=dev)
    xf[loc.to(dev)] = 1.0
    x = torch.fft.ifft(xf) * n
    del xf
    if snr_db is not None:
        std = (k / (2.0 * 10 ** (snr_db / 10.0))) ** 0.5
        gd = torch.Generator(device=dev).manual_seed(seed * 7919 + 1)
        u = torch.rand(n, generator=gd, device=dev, dtype=torch.float64).clamp_min(1e-300)
        v = torch.rand(n, generator=gd, device=dev, dtype=torch.float64)
        x = x + std * torch.sqrt(-2 * torch.log(u)) * torch.exp(2j * torch.pi * v)
    return x.contiguous()


def pcie_probe(ctx):
    """What the host link gives N ranks AT THE SAME TIME: every rank copies 256 MiB host->device
    and 256 MiB device->host concurrently (pinned memory, two streams), 5 times; max over
    ranks.  This is the ceiling of the legacy host-buffer path (`e2e`), which moves 16 n bytes
    each way per transform."""
    torch, dist, dev, world = ctx["torch"], ctx["dist"], ctx["dev"], ctx["world"]
    nbytes, reps = 256 << 20, 5
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once(both):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)

    out = {}
    for name, both in (("h2d_only", False), ("duplex", True)):
        once(both)
        ctx["barrier"]()
        t0 = time.perf_counter()
        for _ in range(reps):
            once(both)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        moved = reps * nbytes * (2 if both else 1)
        out[name + "_gbs_per_rank"] = moved / float(t.item()) / 1e9
        out[name + "_gbs_aggregate"] = world * moved / float(t.item()) / 1e9
        ctx["barrier"]()
    try:
        out["cpu_affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return out


def extra_loop_sharded(ctx, plan, x, what):
    """ONE v1/v2 signal, loops block-partitioned over the ranks, bucket spectra completed by the
    library's NVLink peer exchange (NCCL all-reduce timed beside it), every rank drawing the
    same permutations from identically seeded libc.  Reports the single-GPU time of the same
    transform on the same plan, and whether the sharded result equals it bit for bit."""
    torch, dist, world = ctx["torch"], ctx["dist"], ctx["world"]
    from sfft_b200 import dist as sd
    reps = max(5, min(ctx["steps"], 20))
    x_same = x
    if world > 1:
        x_same = x.clone()
        dist.broadcast(x_same, src=0)               # every rank holds the same signal
    single_ms = timed_ms(ctx, lambda i: plan.execute_device(x_same, None, sync=False), reps)
    out = {"signals": 1, "what": what, "single_gpu_ms": single_ms, "n_gpus": world}
    if world == 1:
        out["ms_per_transform"] = single_ms
        return out
    for exchange in ("peer", "nccl"):
        st = sd.ShardedTransform(plan, exchange=exchange)
        st.seed(17, 4711)
        parity, compared = sd.sharded_matches_single(plan, st, x_same, plan.draw())
        parity2, _ = sd.sharded_matches_single(plan, st, x_same, plan.draw())      # graph replay
        st.seed(17, 4712)
        ms = timed_ms(ctx, lambda i: st.execute(x_same, None, sync=False), reps)
        key = "" if exchange == "peer" else "nccl_"
        out[key + "ms_per_transform"] = ms
        out[key + "sharded_parity"] = bool(parity and parity2)
        if exchange == "peer":
            out["exchange"] = st.exchange
            out["entries_compared_rank0"] = compared
            out["flag_wait_timeouts"] = st.status()[1]
            if st.peer_error:
                out["peer_error"] = st.peer_error
        st.close()
    out["speedup_vs_single_gpu"] = single_ms / out["ms_per_transform"]
    out["note"] = ("one signal resident on every rank; rank r gathers + FFTs its block of loops, stores its rows into "
                   "every peer's spectra buffer over NVLink (CUDA IPC mappings) and flags; selection/voting replicated; "
                   "v2 estimation sliced; whole transform incl. exchange replayed from one CUDA graph; max over ranks")
    return out


def extra_c4(ctx):
    """BASELINE configs[3]: v1, n = 2^27, k = 500, 20 dB AWGN, loops sharded over all ranks."""
    torch = ctx["torch"]
    version, n, k, snr_db, desc = WORKLOADS["C4"]
    plan = ctx["sfft_mod"].sfft(n, k, version, strict_parameters=False)
    plan.set_stream(ctx["stream"].cuda_stream)
    x = device_signal(torch, n, k, 4242, snr_db, ctx["dev"])
    out = extra_loop_sharded(ctx, plan, x, desc)
    out["gsamples_per_s"] = n / (out["ms_per_transform"] * 1e-3) / 1e9
    plan.close()
    del x
    torch.cuda.empty_cache()
    return out


def extra_c5(ctx):
    """BASELINE configs[4]: sfft_exec_many batch of 4096 signals (64 GiB), n = 2^20, k = 100,
    block-partitioned over the ranks (strong scaling of the whole batch); no data-path collective."""
    torch, world, rank, dev = ctx["torch"], ctx["world"], ctx["rank"], ctx["dev"]
    version, n, k, _, _ = WORKLOADS["C5"]
    total, chunk = 4096, 256
    from sfft_b200 import dist as sd
    b, e = sd.partition(total, rank, world)
    plan = ctx["sfft_mod"].sfft(n, k, version, strict_parameters=False)
    plan.set_stream(ctx["stream"].cuda_stream)
    # this rank's signals, all resident in HBM (16 MiB each)
    sig = torch.empty((e - b, n), dtype=torch.complex128, device=dev)
    for i in range(e - b):
        sig[i] = device_signal(torch, n, k, 50000 + b + i, None, dev)

    def one_pass(_):
        for c0 in range(0, e - b, chunk):
            plan.execute_many_device(sig[c0:min(c0 + chunk, e - b)], None, sync=False)

    ms = timed_ms(ctx, one_pass, 3, warm=1)
    plan.close()
    del sig
    torch.cuda.empty_cache()
    return {"signals_total": total, "signals_per_gpu": e - b, "n_gpus": world, "ms_per_batch": ms,
            "gsamples_per_s": total * n / (ms * 1e-3) / 1e9, "scaling": "strong",
            "note": "whole 4096-signal batch resident in HBM, partitioned over ranks, %d signals per launch; "
                    "max over ranks" % chunk}


def extra_c3(ctx):
    """BASELINE configs[2]: v3, n = 2^26, k = 2000.  v3 does not shard (two adjacent time
    shifts + sequential peeling): N independent replicas, each its own signal."""
    torch, world, rank, dev = ctx["torch"], ctx["world"], ctx["rank"], ctx["dev"]
    version, n, k, _, _ = WORKLOADS["C3"]
    plan = ctx["sfft_mod"].sfft(n, k, version, strict_parameters=False)
    plan.set_stream(ctx["stream"].cuda_stream)
    x = device_signal(torch, n, k, 777 + rank, None, dev)
    reps = max(5, min(ctx["steps"], 20))
    dist = ctx["dist"]
    # identical libc state on every rank, so that replicas differ by their signal only
    libc = C.CDLL(None)

    def one(_):
        libc.srand(17)
        libc.srand48(4321)
        plan.execute_device(x, None, sync=False)

    ms = timed_ms(ctx, one, reps)
    # this rank's own time and peeling rounds: v3's round count depends on the signal and the draw
    # (the reference loops until its occupied-bucket counts repeat, computefourier-3.0.cc:1047-1071)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one(0)
    e1.record()
    torch.cuda.synchronize()
    mine = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    libc.srand(17)
    libc.srand48(4321)
    cnt = plan.execute_device(x, None, sync=True)
    import numpy as np
    rounds = torch.tensor([int(plan.debug_fetch("rounds", np.int32, 1)[0])], dtype=torch.float64, device=dev)
    per_rank_ms, per_rank_rounds = [mine.clone() for _ in range(world)], [rounds.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_ms, mine)
        dist.all_gather(per_rank_rounds, rounds)
    plan.close()
    del x
    torch.cuda.empty_cache()
    return {"replicas": world, "ms_per_transform": ms, "gsamples_per_s": world * n / (ms * 1e-3) / 1e9,
            "ms_per_rank": [round(float(t.item()), 4) for t in per_rank_ms],
            "last_transform_rounds_per_rank": [int(t.item()) for t in per_rank_rounds],
            "recovered_coefficients_rank0": int(cnt), "scaling": "weak (replicas only: v3 does not shard)",
            "note": "max over ranks; every replica transforms its own signal and v3's number of peeling rounds "
                    "depends on the signal and the draw"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import sfft_b200.sfft as sfft_mod
    from sfft_b200 import _lib
    L = _lib.load()

    version, n, k, snr_db, desc = WORKLOADS[args.workload]
    t_plan = time.perf_counter()
    plan = sfft_mod.sfft(n, k, version, strict_parameters=False)
    torch.cuda.synchronize()
    plan_ms = 1e3 * (time.perf_counter() - t_plan)
    # every kernel of the engine and every timing event go to ONE explicit stream
    # (torch's default stream has handle 0, which the C ABI reads as "plan's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    plan.set_stream(stream.cuda_stream)
    info = plan.info()

    # rotate over several distinct signals; each is >= L2-sized at the default workload,
    # smaller workloads additionally get an explicit L2 flush between steps
    batch = BATCH.get(args.workload, 1)
    nsig_rot = max(2, min(4, (1 << 28) // n)) if n <= (1 << 26) else 1
    seed0 = SIGNAL_SEED + 100000 * rank
    if batch > 1:
        nsig_rot = 1
        signals = [torch.stack([torch.from_numpy(host_signal(n, k, seed0 + i, snr_db)).to(dev)
                                for i in range(batch)])]
    else:
        signals = [torch.from_numpy(host_signal(n, k, seed0 + i, snr_db)).to(dev) for i in range(nsig_rot)]
    flush = None
    if batch * n * 16 < 256 * 1024 * 1024:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    libc = C.CDLL(None)

    def step(i):
        # "reseeded identically each rep" -- every step draws the same permutations, as
        # the reference arm does, so the work per step does not depend on where in libc's stream it
        # falls (v3 in particular: on a few (signal, draw) pairs the reference ALGORITHM needs a
        # thousand peeling rounds instead of a dozen)
        libc.srand(17 + rank)
        libc.srand48(12345 + rank)
        if batch > 1:
            plan.execute_many_device(signals[0], None, sync=False)
        else:
            plan.execute_device(signals[i % nsig_rot], None, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
        if flush is not None:
            flush.zero_()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.sfftb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
        if flush is not None:
            flush.zero_()          # outside the per-step events
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.sfftb_launch_count() - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(ms_steps)
    clocks = sampler.stop() if rank == 0 else None
    if batch > 1:
        count = int(sum(plan.execute_many_device(signals[0], None, sync=True)) / batch)
    else:
        count = plan.execute_device(signals[0], None, sync=True)

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * args.steps * batch * n / (total_ms_max * 1e-3) / 1e9

    # ---- per-stage device times (separate pass, events between stages) ----
    plan.stage_timing(True)
    acc = {}
    reps = max(3, min(10, args.steps))
    for i in range(reps):
        step(i)
        for nm, ms in plan.stage_times().items():
            acc.setdefault(nm, []).append(ms)
        if flush is not None:
            flush.zero_()
    plan.stage_timing(False)
    stages = {nm: sum(v) / len(v) for nm, v in acc.items()}

    # ---- e2e through the legacy host API (rank-local, all ranks concurrently) ----
    e2e_batch = min(batch, 64)
    h_in = [L.sfft_malloc(16 * n) for _ in range(e2e_batch)]
    h_out = [L.sfft_malloc(16 * n) for _ in range(e2e_batch)]
    e2e_steps = max(1, min(args.steps, 10))
    for i in range(e2e_batch):
        src = signals[0][i] if batch > 1 else signals[0]
        host_sig = src.cpu().numpy()
        C.memmove(h_in[i], host_sig.ctypes.data, 16 * n)
    in_arr = (C.c_void_p * e2e_batch)(*h_in)
    out_arr = (C.c_void_p * e2e_batch)(*h_out)

    def e2e_call():
        if e2e_batch > 1:
            L.sfft_exec_many(plan.sfft_plan, e2e_batch, in_arr, out_arr)
        else:
            L.sfft_exec(plan.sfft_plan, h_in[0], h_out[0])

    for _ in range(min(2, args.warmup)):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * e2e_batch * n / float(te.item()) / 1e9
    for ptr in h_in + h_out:
        L.sfft_free(ptr)

    # ---- what BASELINE.json names beyond configs[1] (every N; --no-extras skips) ----
    extras = {}
    ctx = dict(torch=torch, dist=dist, dev=dev, rank=rank, world=world, stream=stream, barrier=barrier,
               sfft_mod=sfft_mod, steps=args.steps)
    if not args.no_extras:
        extras["pcie_probe"] = guarded(pcie_probe, ctx)
        if version in (1, 2) and batch == 1:
            # one signal of THIS workload with its loops sharded over all ranks
            extras["loop_sharded_single_signal"] = guarded(
                extra_loop_sharded, ctx, plan, signals[0], "this workload's signal")
        if args.workload == "C2":
            extras["c4_loop_sharded"] = guarded(extra_c4, ctx)
            extras["c5_partitioned"] = guarded(extra_c5, ctx)
            extras["c3_replicas"] = guarded(extra_c3, ctx)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"

        # DRAM bytes per launch of each stage's kernels: ncu capture of THIS source tree
        # (tools/make_traffic.py over `ncu --set full` exports, committed per round)
        traffic, traffic_file = {}, None
        for cand in ("r02_traffic.json",):
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", cand))).get(args.workload, {}) \
                    .get("per_stage_dram_bytes", {})
                traffic_file = "profiles/" + cand
                break
            except Exception:
                pass
        notes = {
            "estimate": ("v2_regroup_kernel + v2_fused_kernel: DRAM traffic equals the algorithmic bytes (spectra in, "
                         "result list out); bound by instruction issue: per coefficient ~410 ALU-pipe instructions (two "
                         "medians of 20 selected on 32-bit keys: 258 FMNMX + recovery of the low word) and 320 FP64 ones "
                         "(40 exact divisions), which the 16 warps of a tile execute in two separate phases -- ncu: ALU "
                         "pipe 44 %, FP64 pipe 35 %, shared-memory wavefronts 40 % busy, no eligible warp in 49 % of the "
                         "cycles -- DESIGN.md K7', profiles/r02_v2_fused_keymedian_ncu_details.txt, "
                         "profiles/r02_pipe_overlap.jsonl") if version == 2 else None,
            "gather": ("bound by the random-REQUEST rate of HBM, not its bandwidth: 16-byte samples at a random odd "
                       "stride, one per 32-byte sector (sector efficiency 50 % on the signal stream, the ceiling); "
                       "profiles/r02_gather_ab.md"),
            "peel": "one thread-block cluster per signal; a chain of ~40 dependent steps (global round trips, "
                    "double-precision libm latency), not a bandwidth-bound kernel -- DESIGN.md 5",
        }
        # random 16-byte reads per second the memory system sustains with nothing else in the kernel
        # (tools/microbench/ld_variants under ncu, profiles/r02_ld_variants_*: 72.6 G/s over 256 MiB, 39 G/s over 2 GiB)
        req_ceiling = 72.6e9 if n * 16 <= (256 << 20) else 39e9

        def roof(stage):
            ms = stages.get(stage)
            if not ms:
                return None
            b = batch * stage_bytes(info, stage, count)
            ach = b / (ms * 1e-3) / 1e9
            r = {"kernel": stage, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                 "frac": ach / hbm_peak, "traffic": traffic[stage] if stage in traffic else None,
                 "ms": ms, "algorithmic_bytes": b, "peak_source": peak_src,
                 "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, " + traffic_file)
                 if stage in traffic else None}
            if notes.get(stage):
                r["note"] = notes[stage]
            if stage == "estimate" and version == 2:
                # floor set by the instruction mix on the measured pipe rates (DESIGN.md K7'): per
                # recovered coefficient ~410 ALU + 320 FP64 instructions, 2.06 / 2.45 cycles each alone,
                # 1.5 cycles per instruction when mixed (tools/microbench/pipe_overlap): 4330 cycles per
                # 512-coefficient tile (its 16 warps sit four to a sub-partition), one tile at a time per SM
                floor_ms = batch * count / 512.0 / 148 * 4330 / 1.965e9 * 1e3
                r["instruction_floor_ms"] = floor_ms
                r["frac_of_instruction_floor"] = floor_ms / ms
            if stage == "gather" and batch == 1:
                samples = info["gather_samples"] - (info["Comb_loops"] * info["W_Comb"] if version == 2 else 0)
                rate = samples / (ms * 1e-3)
                r["random_requests_per_s"] = rate
                r["random_request_ceiling_per_s"] = req_ceiling
                r["frac_of_request_ceiling"] = rate / req_ceiling
                r["sector_efficiency_signal_stream"] = 0.5
            return r

        dominant = max(stages, key=stages.get) if stages else None
        line = {
            "metric": "sfft_exec throughput", "value": value, "unit": "Gsamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "n": n, "k": k, "version": version,
                       "signals_per_gpu_per_step": batch,
                       "cache": ("input %d MiB > L2, %d signals rotated" % (16 * n * batch >> 20, nsig_rot))
                       + ("" if flush is None else ", 256 MiB L2 flush between steps"),
                       "recovered_coefficients": int(count),
                       "plan": {kk: info[kk] for kk in ("B_loc", "B_est", "loops_loc", "loops_est", "w_loc",
                                                       "w_est", "W_Comb", "Comb_loops", "x_samp_size")}},
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": 16 * n * e2e_batch,
                    "d2h_bytes_per_step": 16 * n * e2e_batch, "steps": e2e_steps,
                    "api": ("sfft_exec_many(plan, %d, host_in[], host_out[])" % e2e_batch if e2e_batch > 1
                            else "sfft_exec(plan, host_in, host_out)") + ", pinned buffers from sfft_malloc"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof(dominant) if dominant else None,
            "roofline_gather": roof("gather"),
            "stage_ms": stages,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "plan_ms": plan_ms,
        }
        line.update(extras)
        probe = extras.get("pcie_probe") or {}
        if "duplex_gbs_aggregate" in probe:
            # the legacy path moves 16 n bytes in and 16 n bytes out per transform
            e2e_gbs = e2e_value * 1e9 * 32 / 1e9
            line["e2e"]["host_link_gbs_used"] = e2e_gbs
            line["e2e"]["host_link_gbs_one_way_ceiling"] = probe["h2d_only_gbs_aggregate"]
            line["e2e"]["host_link_gbs_duplex_ceiling"] = probe["duplex_gbs_aggregate"]
            line["e2e"]["note"] = ("bytes moved over the host link per second by the e2e run vs what the N ranks reach "
                                   "copying concurrently on this box (pcie_probe). v2's result is dense, so its input "
                                   "and output copies cannot overlap: the one-way figure is its ceiling; v1/v3 stream "
                                   "the zero fill out while the input streams in (duplex)")
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args.workload, min(batch, os.cpu_count() or 1))
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-leg-signals", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-leg-budget", type=float, default=150.0, help=argparse.SUPPRESS)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra BASELINE configs (C4 loop-sharded, C5 partitioned, C3 replicas)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if WORKLOADS[args.workload][0] == 3 and "MALLOC_MMAP_THRESHOLD_" not in os.environ:
            # the reference's v3 path overruns a heap buffer by one element (computefourier-3.0.cc:235 vs
            # sfft.cc:497-498): give big allocations their own mappings so the overrun lands in page slack
            os.execve(sys.executable, [sys.executable] + sys.argv, dict(os.environ, MALLOC_MMAP_THRESHOLD_="32768"))
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
