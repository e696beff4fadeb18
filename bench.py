#!/usr/bin/env python
"""bench.py -- throughput of the FFT worker hot path on B200.

Metric (BASELINE.json): complex IQ Msamples/s (bins x spectra / s) at 4096 bins,
and the achieved fraction of the HBM-read roofline (2 algorithmic bytes per
complex sample).

One "step" = one whole integration of BASELINE config 2 on each GPU: 4096 bins,
1e6 spectra, a device-resident synthetic u8 IQ stream of 8.192 GB (>> the 126 MB
L2, so every step streams from HBM): rpf_reset -> rpf_push_device ->
rpf_finish_device.  With N > 1 (one process per GPU, torchrun) every rank's
context joins one NCCL communicator THROUGH THE C ABI (rpf_comm_unique_id on
rank 0, the 128-byte id shipped with torch.distributed, rpf_comm_init_rank), so
the one ncclAllReduce of pwr[4096] (fp64) per step is issued by the library
inside rpf_finish_device.  Weak scaling: every rank integrates its own
1e6-spectra shard.

  value      : whole-job Msamples/s, input resident in HBM, CUDA-event timed,
               max over ranks.
  e2e        : the same metric through the reference-facing C ABI with HOST
               buffers (pinned), host->device copies and the pwr[] read-back
               inside the timed region, on the full 1e6 spectra per GPU;
               e2e.roofline compares it with the aggregate pinned H2D rate
               measured on the same box in the same run (no kernels).
  roofline   : dominant kernel (pipe4096_kernel): 2 B x samples per launch /
               its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  parity     : exact Parseval of the timed result, allreduce == sum of the
               gathered per-rank vectors, rank 0's prefix against the oracle.
  configs    : the other single-GPU BASELINE configs at their stated sizes
               (1: 512 x 100 through rpf_push; 3: 4096 Hann x 1e7; 4: 65536 x
               1e5), each with Msamples/s, roofline fraction and parity; at
               N > 1 config 5 as written (1e7 spectra per GPU).
  cpu_baseline / --impl reference: the reference's own worker
               (src/datastore.cxx compiled from /root/reference into
               oracle/_ref, FFT = stand-in Stockham since libfftw3f is absent)
               on the box's host cores.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_BINS = 4096
SPECTRA_PER_STEP = 1000000            # config 2 (per GPU)
E2E_RING_BYTES = 256 << 20            # pinned host ring per rank, replayed to cover the whole integration
E2E_CHUNK = 32 << 20
SEED = 0x5EED + 2
METRIC = "IQ Msamples/s (bins x spectra/s) at 4096 bins; achieved HBM GB/s vs roofline"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw = [], [], []
        reasons = set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------
# CPU reference arm (oracle/_ref: the reference's own Datastore::fftThread).
# Nothing here touches librpf_b200.so: the byte stream comes from numpy.
# ---------------------------------------------------------------------------
def numpy_stream(nbytes, seed=SEED):
    """Clipped-Gaussian u8 IQ bytes (mean 127.4, sigma 30), the distribution of the GPU arm's generator mode 1."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.normal(127.4, 30.0, nbytes)), 0, 255).astype(np.uint8)


def cpu_reference_run(workers, repeats_each, ring):
    import oracle_lib as ol
    pwr = np.zeros(N_BINS)
    done = C.c_int64(0)
    t0 = time.perf_counter()
    rc = ol.ref().rpf_ref_run_parallel(workers, N_BINS, repeats_each, None, 5, 1638400,
                                       ring.ctypes.data_as(C.POINTER(C.c_uint8)), ring.size,
                                       pwr.ctypes.data_as(C.POINTER(C.c_double)), C.byref(done))
    dt = time.perf_counter() - t0
    if rc:
        raise RuntimeError("rpf_ref_run_parallel rc=%d" % rc)
    return done.value * N_BINS / dt / 1e6, dt, done.value


def calibrate_reps(workers, ring, target_s):
    """Spectra per worker so that one rpf_ref_run_parallel call lasts ~target_s."""
    reps = 500
    while True:
        _, dt, _ = cpu_reference_run(workers, reps, ring)
        if dt >= 0.4 or reps >= (1 << 24):
            return max(500, int(reps * target_s / dt))
        reps *= 4


def fft_only_rates(seconds=1.0):
    """FFT-only rate of the stand-in behind oracle/_ref's fftw3.h shim next to scipy.fft (pocketfft, complex64,
    one worker) on the same core: validates the denominator of the CPU arm (libfftw3f itself is absent)."""
    import oracle_lib as ol
    out = {}
    lib = ol.ref()
    lib.fftwf_alloc_complex.restype = C.c_void_p
    lib.fftwf_alloc_complex.argtypes = [C.c_ulong]
    lib.fftwf_plan_dft_1d.restype = C.c_void_p
    lib.fftwf_plan_dft_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    lib.fftwf_execute.argtypes = [C.c_void_p]
    a, b = lib.fftwf_alloc_complex(N_BINS), lib.fftwf_alloc_complex(N_BINS)
    x = (np.random.default_rng(1).normal(size=2 * N_BINS) * 30).astype(np.float32)
    C.memmove(a, x.ctypes.data, x.nbytes)
    plan = lib.fftwf_plan_dft_1d(N_BINS, a, b, -1, 0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(200):
            lib.fftwf_execute(plan)
        n += 200
    out["standin_fft_only_msamples_s"] = n * N_BINS / (time.perf_counter() - t0) / 1e6
    try:
        import scipy.fft as sf
        z = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
        batch = np.tile(z, (256, 1))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            sf.fft(batch, axis=1, workers=1)
            n += 256
        out["scipy_pocketfft_c64_1worker_msamples_s"] = n * N_BINS / (time.perf_counter() - t0) / 1e6
    except Exception as e:                                   # scipy missing: report, do not fail the arm
        out["scipy_pocketfft_c64_1worker_msamples_s"] = None
        out["scipy_error"] = str(e)[:80]
    return out


def cpu_baseline(target_s=10.0):
    """Bounded sample on the host cores; returns the cpu_baseline JSON object."""
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    ring = numpy_stream(1638400 * 8)
    if not ol.have_ref():
        # fall back to the C restatement (single thread)
        t0 = time.perf_counter()
        reps = 2000
        ol.oracle_run_stream(N_BINS, reps, ring, buf_length=1638400, fft_kind=1, wrap=True)
        dt = time.perf_counter() - t0
        return {"value": reps * N_BINS / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                "sample": "%d spectra of 4096 bins, oracle/rpf_oracle.c float32 flavour" % reps}
    reps = calibrate_reps(cores, ring, target_s)
    rate, dt, done = cpu_reference_run(cores, reps, ring)
    r1, dt1, done1 = cpu_reference_run(1, max(400, reps // 4), ring)
    out = {"value": rate, "unit": "Msamples/s", "cores": cores, "kind": "reference",
           "single_worker_value": r1,
           "sample": ("%d independent reference workers (src/datastore.cxx fftThread, stand-in FFT for the absent "
                      "libfftw3f) x %d spectra of 4096 bins each, %.1f s; single worker (the reference's real "
                      "configuration): %d spectra in %.1f s" % (cores, reps, dt, done1, dt1))}
    out.update(fft_only_rates(0.7))
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    ring = numpy_stream(1638400 * 8)
    if not ol.have_ref():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference sources absent)"}))
        return
    reps = calibrate_reps(cores, ring, 1.5)      # a step = ~1.5 s of wall clock
    for _ in range(args.warmup):
        cpu_reference_run(cores, reps, ring)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        _, _, done = cpu_reference_run(cores, reps, ring)
        total += done
    dt = time.perf_counter() - t0
    value = total * N_BINS / dt / 1e6
    sample = ("each step: %d reference workers x %d spectra of 4096 bins (bounded sample of config 2), "
              "FFT = stand-in Stockham (libfftw3f absent); bytes from numpy" % (cores, reps))
    cb = {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample}
    cb.update(fft_only_rates(0.7))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "BASELINE config 2 per GPU: 4096-bin FFT, 1e6 spectra averaged, synthetic u8 IQ stream "
                        "(clipped Gaussian, sigma 30), no window",
            "bins": N_BINS, "spectra_per_step_per_gpu": SPECTRA_PER_STEP,
            "input_bytes_per_step_per_gpu": 2 * N_BINS * SPECTRA_PER_STEP,
            "l2": "input 8.19 GB per step per GPU >> 126 MB L2 (no flush needed)",
            "parallelism": ("spectra sharded over %d GPU(s), one process per GPU; one NCCL allreduce of pwr[4096] fp64 "
                            "per step, issued by librpf_b200.so (rpf_comm_init_rank)" % n_gpus)
            if n_gpus > 1 else "single GPU"}


# ---------------------------------------------------------------------------
# parity helpers (torch on the device; the oracle only as the checker)
# ---------------------------------------------------------------------------
def parseval_rhs(torch, stream_dev, n_bins, window=None, chunk_spectra=None):
    """N * sum_n w_n^2 |x_n - 127(1+j)|^2 over the whole device-resident stream, exact integer sums per chunk
    (float64 once the window weights enter)."""
    nspec = stream_dev.numel() // (2 * n_bins)
    if chunk_spectra is None:
        chunk_spectra = max(1, (256 << 20) // (2 * n_bins))
    w2 = None
    if window is not None:
        w2 = torch.from_numpy(np.asarray(window, np.float32).astype(np.float64) ** 2).to(stream_dev.device)
    total = 0.0
    for s0 in range(0, nspec, chunk_spectra):
        s1 = min(nspec, s0 + chunk_spectra)
        x = stream_dev[2 * n_bins * s0:2 * n_bins * s1].view(s1 - s0, n_bins, 2).to(torch.int32) - 127
        e = (x * x).sum(dim=2)                                   # |x|^2 per sample, exact (< 2^16)
        if w2 is None:
            total += float(e.sum(dtype=torch.int64).item())
        else:
            total += float((e.sum(dim=0, dtype=torch.int64).to(torch.float64) * w2).sum().item())
        del x, e
    return n_bins * total


def oracle_prefix_err(pkg, torch, stream_dev, n_bins, nspec, window, device):
    """max per-bin relative error of the CUDA path against the float64 oracle on the first `nspec` spectra."""
    import oracle_lib as ol
    host = stream_dev[:2 * n_bins * nspec].cpu().numpy()
    po, _ = ol.oracle_run_stream(n_bins, nspec, host, window=window, buf_length=host.size)
    with pkg.Context(n_bins, window=window, device=device) as c:
        c.reset(nspec)
        c.push_device(stream_dev.data_ptr(), 2 * n_bins * nspec)
        pg, done = c.finish()
    assert done == nspec
    return float(np.max(np.abs(pg - po) / po)), float(np.max(np.abs(pg - po)) / np.max(po))


def bench_config_device(pkg, torch, dev, local, name, n_bins, nspec, window, seed, steps, peak, kernel):
    """One device-resident BASELINE config at its stated size: Msamples/s, roofline fraction, parity."""
    nbytes = 2 * n_bins * nspec
    stream_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream(dev)
    pkg.synth_fill_device(stream_dev.data_ptr(), nbytes, 0, seed=seed, mode=1, device=local, stream=cur.cuda_stream)
    pwr_dev = torch.zeros(n_bins, dtype=torch.float64, device=dev)
    ctx = pkg.Context(n_bins, window=window, device=local)
    ctx.set_stream(cur.cuda_stream)
    if isinstance(kernel, dict):
        kernel = kernel[ctx.plan_name]
    for _ in range(2):
        ctx.reset(nspec); ctx.push_device(stream_dev.data_ptr(), nbytes); ctx.finish_device(pwr_dev.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.kernel_launches
    e0.record()
    for _ in range(steps):
        ctx.reset(nspec); ctx.push_device(stream_dev.data_ptr(), nbytes); ctx.finish_device(pwr_dev.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (ctx.kernel_launches - l0) // steps
    ctx.close()
    pw = pwr_dev.cpu().numpy()
    rhs = parseval_rhs(torch, stream_dev, n_bins, window=window)
    perbin, relmax = oracle_prefix_err(pkg, torch, stream_dev, n_bins, 64 if n_bins <= 4096 else 8, window, local)
    pv = abs(pw.sum() - rhs) / rhs
    rate = nspec * n_bins / (ms * 1e-3)
    ok = bool(np.all(np.isfinite(pw)) and pv < (2e-6 if window is not None else 1e-6) and relmax < 1e-5 and perbin < 1e-4)
    del stream_dev
    torch.cuda.empty_cache()
    return {"name": name, "bins": n_bins, "spectra": nspec, "window": "hann" if window is not None else None,
            "input": "device-resident %.2f GB" % (nbytes / 1e9), "steps": steps, "ms_per_step": ms,
            "Msamples_per_s": rate / 1e6, "kernel": kernel, "launches_per_step": int(launches),
            "roofline": {"bound": "hbm", "achieved": 2 * rate / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": 2 * rate / 1e9 / peak, "note": "2 B per complex sample over the whole step"},
            "parity": {"parseval_rel": pv, "oracle_prefix_perbin": perbin, "oracle_prefix_relmax": relmax, "ok": ok}}


def bench_config1(pkg, local, reps=300):
    """BASELINE config 1 as the reference runs it: 512 bins, 100 spectra, ONE 114,688-byte host buffer through
    rpf_reset -> rpf_push -> rpf_finish (host wall clock per integration)."""
    import oracle_lib as ol
    N, R, nbytes = 512, 100, 114688
    ptr = C.c_void_p()
    assert pkg.lib().rpf_host_alloc(C.byref(ptr), nbytes) == 0
    pkg.lib().rpf_synth_fill_host(ptr, nbytes, 0, 0x5EED + 1, 1)
    host = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy()
    with pkg.Context(N, device=local, buf_length=nbytes) as c:
        for _ in range(20):
            c.reset(R); c.push_ptr(ptr.value, nbytes); pw, done = c.finish()
        t0 = time.perf_counter()
        for _ in range(reps):
            c.reset(R); c.push_ptr(ptr.value, nbytes); pw, done = c.finish()
        dt = (time.perf_counter() - t0) / reps
        launches = c.kernel_launches
        plan = c.plan_name
    pkg.lib().rpf_host_free(ptr)
    po, done_o = ol.oracle_run_stream(N, R, host, buf_length=nbytes)
    perbin = float(np.max(np.abs(pw - po) / po))
    return {"name": "config1", "bins": N, "spectra": R, "window": None,
            "input": "one 114,688-byte pinned host buffer through rpf_push (12 surplus spectra discarded)",
            "us_per_integration": dt * 1e6, "Msamples_per_s": R * N / dt / 1e6, "kernel": "%s_kernel<512> + reduce_rows_kernel" % plan,
            "timing": "host wall clock around rpf_reset + rpf_push + rpf_finish, %d integrations" % reps,
            "parity": {"oracle_perbin": perbin, "repeats_done": int(done), "ok": bool(done == done_o == R and perbin < 1e-5)},
            "launches_total": int(launches)}


def check_parity(pkg, torch, dist, stream_dev, pwr_dev, pw, nspec, nbytes, cur, dev, local, rank, world, allsum):
    """Exact Parseval of the timed result over every rank's stream, allreduce == sum of the gathered per-rank
    vectors (rank order, float64), rank 0's 64-spectrum prefix against the float64 oracle."""
    rhs = allsum(parseval_rhs(torch, stream_dev, N_BINS))            # N * sum |x - 127|^2 over every rank's stream
    parity = {"parseval_rel": abs(pw.sum() - rhs) / rhs}
    with pkg.Context(N_BINS, device=local) as cl:                    # no communicator: this rank's own vector
        cl.set_stream(cur.cuda_stream)
        cl.reset(nspec)
        cl.push_device(stream_dev.data_ptr(), nbytes)
        local_dev = torch.zeros(N_BINS, dtype=torch.float64, device=dev)
        cl.finish_device(local_dev.data_ptr())
        torch.cuda.synchronize()
    if world > 1:
        gathered = [torch.zeros_like(local_dev) for _ in range(world)]
        dist.all_gather(gathered, local_dev)
        ssum = torch.zeros_like(local_dev)
        for g in gathered:                                           # rank order, float64
            ssum += g
        parity["allreduce_vs_gathered_sum_rel"] = float(((pwr_dev - ssum).abs() / ssum).max().item())
    else:
        parity["allreduce_vs_gathered_sum_rel"] = float(((pwr_dev - local_dev).abs() / local_dev).max().item())
    if rank == 0:
        perbin, relmax = oracle_prefix_err(pkg, torch, stream_dev, N_BINS, 64, None, local)
        parity["oracle_prefix_perbin_rank0"] = perbin
    parity["ok"] = bool(np.all(np.isfinite(pw)) and parity["parseval_rel"] < 1e-6 and
                        parity["allreduce_vs_gathered_sum_rel"] < 1e-12 and
                        parity.get("oracle_prefix_perbin_rank0", 0.0) < 1e-5)
    assert parity["ok"], parity

    return parity


_OUT = None


def emit(text):
    """The one JSON line, on the process's real stdout."""
    out = _OUT if _OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spectra", type=int, default=SPECTRA_PER_STEP, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[] legs (1, 3, 4 / 5)")
    ap.add_argument("--no-parity", action="store_true", help=argparse.SUPPRESS)   # ncu launch lists: only the step's kernels
    ap.add_argument("--flags", type=int, default=0, help=argparse.SUPPRESS)   # rpf_config.flags of the timed context
    args = ap.parse_args()
    # The JSON line must be the ONLY thing on stdout: libraries chat there (NCCL prints a version banner when torch
    # creates its communicator under NCCL_DEBUG).  Keep the real stdout aside and point fd 1 at stderr for the run.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("rtl-power-fftw_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = load_peaks()
    nspec = args.spectra
    nbytes = 2 * N_BINS * nspec
    stream_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    # a real (non-NULL) stream: rpf_set_stream(NULL) means "the context's own stream", and
    # torch.cuda.Event only sees torch's current stream -- so make them the same one.
    side = torch.cuda.Stream(dev)
    torch.cuda.set_stream(side)
    cur = torch.cuda.current_stream(dev)
    assert cur.cuda_stream != 0
    # each rank owns a different slice of the global stream (rank-dependent offset)
    pkg.synth_fill_device(stream_dev.data_ptr(), nbytes, rank * nbytes, seed=SEED, mode=1, device=local,
                          stream=cur.cuda_stream)
    pwr_dev = torch.zeros(N_BINS, dtype=torch.float64, device=dev)
    ctx = pkg.Context(N_BINS, device=local, buf_length=E2E_CHUNK, flags=args.flags)
    ctx.set_stream(cur.cuda_stream)
    if world > 1:
        # the communicator lives in the library: id from rank 0 through the C ABI, shipped by torch.distributed
        box = [pkg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init_rank(box[0], rank, world)
    torch.cuda.synchronize()

    def step():
        ctx.reset(nspec)
        ctx.push_device(stream_dev.data_ptr(), nbytes)
        ctx.finish_device(pwr_dev.data_ptr())          # reduce_rows + (N > 1) ncclAllReduce inside the library

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()

    # ---- timed region: device-resident input --------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    barrier()
    ev[0].record()
    for i in range(args.steps):
        ctx.reset(nspec)
        kev[i][0].record()
        ctx.push_device(stream_dev.data_ptr(), nbytes)       # exactly one pipe4096_kernel launch
        kev[i][1].record()
        ctx.finish_device(pwr_dev.data_ptr())
    ev[1].record()
    torch.cuda.synchronize()
    barrier()
    launches = ctx.kernel_launches - l0
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ms_total = allmax(ev[0].elapsed_time(ev[1]))
    value = world * nspec * N_BINS * args.steps / (ms_total * 1e-3) / 1e6

    # ---- parity of the timed result (no skipped work, nothing lost in the reduction) ----------
    pw = pwr_dev.cpu().numpy()
    parity = None
    if not args.no_parity:
        parity = check_parity(pkg, torch, dist, stream_dev, pwr_dev, pw, nspec, nbytes, cur, dev, local, rank, world, allsum)
    # ---- e2e: host buffers through the C ABI, full integration per step ----------------------
    e2e = None
    if not args.no_e2e:
        ring_ptr = C.c_void_p()
        assert pkg.lib().rpf_host_alloc(C.byref(ring_ptr), E2E_RING_BYTES) == 0
        pkg.lib().rpf_synth_fill_host(ring_ptr, E2E_RING_BYTES, rank * E2E_RING_BYTES, SEED, 1)
        host_pwr = np.empty(N_BINS)
        total_bytes = 2 * N_BINS * nspec
        ectx = pkg.Context(N_BINS, device=local, buf_length=E2E_CHUNK, flags=pkg.capi.RPF_FLAG_PIPELINED_PUSH)
        if world > 1:
            box = [pkg.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ectx.comm_init_rank(box[0], rank, world)

        def e2e_step():
            ectx.reset(nspec)
            sent = 0
            while sent < total_bytes:
                off = sent % E2E_RING_BYTES
                n = min(E2E_CHUNK, total_bytes - sent, E2E_RING_BYTES - off)
                ectx.push_ptr(ring_ptr.value + off, n)                 # rpf_push: H2D inside, 2 chunks in flight
                sent += n
            p, d = ectx.finish()                                       # reduce (+ allreduce) + D2H of pwr[]
            assert d == nspec
            host_pwr[:] = p

        # aggregate pinned host->device rate of this box, same buffers, no kernels: the e2e roofline
        h2d_dst = torch.empty(E2E_CHUNK, dtype=torch.uint8, device=dev)
        ring_t = torch.frombuffer((C.c_uint8 * E2E_RING_BYTES).from_address(ring_ptr.value), dtype=torch.uint8)
        cs = [torch.cuda.Stream(dev) for _ in range(2)]
        def h2d_pass(passes):
            k = 0
            for _ in range(passes):
                for off in range(0, E2E_RING_BYTES, E2E_CHUNK):
                    with torch.cuda.stream(cs[k & 1]):
                        h2d_dst.copy_(ring_t[off:off + E2E_CHUNK], non_blocking=True)
                    k += 1
            torch.cuda.synchronize()
        h2d_pass(1)
        barrier()
        t0 = time.perf_counter()
        h2d_pass(4)
        h2d_dt = allmax(time.perf_counter() - t0)
        h2d_peak = world * 4 * E2E_RING_BYTES / h2d_dt / 1e9
        del h2d_dst

        e2e_steps = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = allmax(time.perf_counter() - t0)
        e2e_launches = ectx.kernel_launches
        ectx.close()
        # the e2e result is a real spectrum: exact Parseval against the ring it was fed from
        ring_dev = ring_t.to(dev)
        ring_rhs = parseval_rhs(torch, ring_dev, N_BINS)
        full, rem = divmod(total_bytes, E2E_RING_BYTES)
        e2e_rhs = full * ring_rhs + (parseval_rhs(torch, ring_dev[:rem], N_BINS) if rem else 0.0)
        e2e_rhs = allsum(e2e_rhs)
        e2e_pv = abs(host_pwr.sum() - e2e_rhs) / e2e_rhs
        assert np.all(np.isfinite(host_pwr)) and e2e_pv < 1e-6, e2e_pv
        rate = world * nspec * N_BINS * e2e_steps / dt / 1e6
        e2e = {"value": rate, "unit": "Msamples/s",
               "h2d_bytes_per_step": world * total_bytes, "d2h_bytes_per_step": world * N_BINS * 8,
               "steps": e2e_steps, "parseval_rel": e2e_pv, "kernel_launches": int(e2e_launches),
               "roofline": {"bound": "pcie_h2d", "achieved": 2 * rate / 1e3, "peak": h2d_peak, "unit": "GB/s",
                            "frac": 2 * rate / 1e3 / h2d_peak,
                            "peak_source": "measured in this run: %d rank(s) x pinned cudaMemcpyAsync of the same ring, "
                                           "two streams, no kernels" % world},
               "sample": "the full %d spectra per GPU per step, replayed from a %d MiB pinned host ring, rpf_push in "
                         "%d MiB chunks with RPF_FLAG_PIPELINED_PUSH (two chunks in flight), rpf_finish reads pwr[] "
                         "back; host wall clock (max over ranks)" % (nspec, E2E_RING_BYTES >> 20, E2E_CHUNK >> 20)}
        del ring_dev, ring_t
        pkg.lib().rpf_host_free(ring_ptr)
    clocks = sampler.stop() if sampler else None      # covers the device-timed and the e2e regions
    ctx.close()
    del stream_dev
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs at their stated sizes -------------------------------------
    configs = []
    if not args.no_configs:
        hann = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N_BINS) / N_BINS)).astype(np.float32)
        if world == 1:
            configs.append(bench_config1(pkg, local))
            configs.append(bench_config_device(pkg, torch, dev, local, "config3", N_BINS, 10000000, hann, 0x5EED + 3, 3,
                                               peak, "pipe4096_kernel<window>"))
            configs.append(bench_config_device(pkg, torch, dev, local, "config4", 65536, 100000, None, 0x5EED + 4, 3,
                                               peak, {"solo65536": "solo65536_kernel (one launch, one CTA per spectrum)",
                                                      "fourstep": "fourstep_outer_tma_kernel<16> + pipe4096_kernel<c32>"}))
        else:
            # config 5 as written: 8e7 spectra over 8 GPUs = 1e7 per GPU, one allreduce at the end
            n5 = 10000000
            b5 = 2 * N_BINS * n5
            s5 = torch.empty(b5, dtype=torch.uint8, device=dev)
            pkg.synth_fill_device(s5.data_ptr(), b5, rank * b5, seed=0x5EED + 5, mode=1, device=local, stream=cur.cuda_stream)
            c5 = pkg.Context(N_BINS, device=local)
            c5.set_stream(cur.cuda_stream)
            box = [pkg.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            c5.comm_init_rank(box[0], rank, world)
            p5 = torch.zeros(N_BINS, dtype=torch.float64, device=dev)
            for _ in range(2):
                c5.reset(n5); c5.push_device(s5.data_ptr(), b5); c5.finish_device(p5.data_ptr())
            torch.cuda.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                c5.reset(n5); c5.push_device(s5.data_ptr(), b5); c5.finish_device(p5.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms5 = allmax(e0.elapsed_time(e1)) / 3
            c5.close()
            rhs5 = allsum(parseval_rhs(torch, s5, N_BINS))
            pv5 = abs(float(p5.sum().item()) - rhs5) / rhs5
            rate5 = world * n5 * N_BINS / (ms5 * 1e-3)
            configs.append({"name": "config5", "bins": N_BINS, "spectra": world * n5, "spectra_per_gpu": n5, "window": None,
                            "input": "device-resident %.1f GB per GPU" % (b5 / 1e9), "steps": 3, "ms_per_step": ms5,
                            "Msamples_per_s": rate5 / 1e6, "kernel": "pipe4096_kernel + one ncclAllReduce(pwr, 4096 x f64)",
                            "roofline": {"bound": "hbm", "achieved": 2 * rate5 / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                                         "frac": 2 * rate5 / 1e9 / world / peak},
                            "parity": {"parseval_rel": pv5, "ok": bool(pv5 < 1e-6)}})
            del s5
            torch.cuda.empty_cache()

    if rank == 0:
        achieved = 2.0 * nspec * N_BINS / (kern_ms * 1e-3) / 1e9          # algorithmic GB/s of one launch
        tr = load_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "kernel": "pipe4096_tc_kernel" if args.flags & 1 else "pipe4096_kernel",
                         "kernel_ms_per_launch": kern_ms,
                         "algorithmic_bytes_per_launch": 2 * nspec * N_BINS,
                         "note": "2 B per complex sample; FP32-issue and shared-memory bounds in DESIGN.md"},
            "parity": parity, "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
            "configs": configs,
        }
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        emit(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
