#!/usr/bin/env python
"""bench.py -- throughput of the FFT worker hot path on B200.

Metric (BASELINE.json): complex IQ Msamples/s (bins x spectra / s) at 4096 bins,
and the achieved fraction of the HBM-read roofline (2 algorithmic bytes per
complex sample).

One "step" = one whole integration of BASELINE config 2 on each GPU: 4096 bins,
1e6 spectra, a device-resident synthetic u8 IQ stream of 8.192 GB (>> the 126 MB
L2, so every step streams from HBM), rpf_reset -> rpf_push_device ->
rpf_finish_device, plus -- when N > 1 -- one NCCL allreduce of pwr[4096] (fp64).
Weak scaling: every rank integrates its own 1e6-spectra shard (config 5 shape).

  value      : whole-job Msamples/s, input resident in HBM, CUDA-event timed,
               max over ranks.
  e2e        : the same metric through the reference-facing C ABI with HOST
               buffers (pinned), host->device copies and the pwr[] read-back
               inside the timed region.
  roofline   : dominant kernel (pipe4096_kernel): 2 B x samples per launch /
               its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
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
SPECTRA_PER_STEP = 1000000            # config 2 (per GPU; config 5 = 8 x 1e7 is the same shape, weak)
E2E_RING_BYTES = 256 << 20            # pinned host ring per rank
E2E_PASSES = 4                        # e2e step = 4 passes over the ring = 1 GiB = 131072 spectra
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
# CPU reference arm (oracle/_ref: the reference's own Datastore::fftThread)
# ---------------------------------------------------------------------------
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


def cpu_baseline(target_s=10.0):
    """Bounded sample on the host cores; returns the cpu_baseline JSON object."""
    import oracle_lib as ol
    pkg = importlib.import_module("rtl-power-fftw_b200")
    cores = os.cpu_count() or 1
    ring = pkg.synth_fill_host(1638400 * 8, 0, seed=SEED, mode=1)
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
    return {"value": rate, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "single_worker_value": r1,
            "sample": ("%d independent reference workers (src/datastore.cxx fftThread, stand-in FFT for the absent "
                       "libfftw3f) x %d spectra of 4096 bins each, %.1f s; single worker (the reference's real "
                       "configuration): %d spectra in %.1f s" % (cores, reps, dt, done1, dt1))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as ol
    pkg = importlib.import_module("rtl-power-fftw_b200")
    cores = os.cpu_count() or 1
    ring = pkg.synth_fill_host(1638400 * 8, 0, seed=SEED, mode=1)
    if not ol.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (reference sources absent)"}))
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
              "FFT = stand-in Stockham (libfftw3f absent)" % (cores, reps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                             "sample": sample},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "BASELINE config 2 per GPU: 4096-bin FFT, 1e6 spectra averaged, synthetic u8 IQ stream "
                        "(clipped Gaussian, sigma 30), no window",
            "bins": N_BINS, "spectra_per_step_per_gpu": SPECTRA_PER_STEP,
            "input_bytes_per_step_per_gpu": 2 * N_BINS * SPECTRA_PER_STEP,
            "l2": "input 8.19 GB per step per GPU >> 126 MB L2 (no flush needed)",
            "parallelism": "spectra sharded over %d GPU(s); one NCCL allreduce of pwr[4096] fp64 per step" % n_gpus
            if n_gpus > 1 else "single GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spectra", type=int, default=SPECTRA_PER_STEP, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
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
    ctx = pkg.Context(N_BINS, device=local, buf_length=E2E_CHUNK)
    ctx.set_stream(cur.cuda_stream)
    torch.cuda.synchronize()

    def step():
        ctx.reset(nspec)
        ctx.push_device(stream_dev.data_ptr(), nbytes)
        ctx.finish_device(pwr_dev.data_ptr())
        if world > 1:
            dist.all_reduce(pwr_dev)          # NCCL, same stream

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
        ctx.push_device(stream_dev.data_ptr(), nbytes)       # exactly one fused3_kernel launch
        kev[i][1].record()
        ctx.finish_device(pwr_dev.data_ptr())
        if world > 1:
            dist.all_reduce(pwr_dev)
    ev[1].record()
    torch.cuda.synchronize()
    barrier()
    launches = ctx.kernel_launches - l0
    ms_total = ev[0].elapsed_time(ev[1])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * nspec * N_BINS * args.steps / (ms_total * 1e-3) / 1e6

    # sanity: result of the last step must be a real spectrum (no skipped work)
    pw = pwr_dev.cpu().numpy()
    expect = world * nspec * N_BINS * (2 * 30.0 ** 2)        # Parseval: ~ N * sum |x|^2 per bin
    assert np.all(np.isfinite(pw)) and 0.5 * expect < pw.mean() < 1.5 * expect, (pw.mean(), expect)

    # ---- e2e: host buffers through the C ABI --------------------------------
    e2e = None
    if not args.no_e2e:
        ring_ptr = C.c_void_p()
        assert pkg.lib().rpf_host_alloc(C.byref(ring_ptr), E2E_RING_BYTES) == 0
        pkg.lib().rpf_synth_fill_host(ring_ptr, E2E_RING_BYTES, rank * E2E_RING_BYTES, SEED, 1)
        e2e_spec = E2E_PASSES * E2E_RING_BYTES // (2 * N_BINS)
        host_pwr = np.empty(N_BINS)

        def e2e_step():
            ctx.reset(e2e_spec)
            for _ in range(E2E_PASSES):
                for off in range(0, E2E_RING_BYTES, E2E_CHUNK):
                    ctx.push_ptr(ring_ptr.value + off, E2E_CHUNK)         # rpf_push: H2D inside
            if world > 1:
                ctx.finish_device(pwr_dev.data_ptr())
                dist.all_reduce(pwr_dev)
                host_pwr[:] = pwr_dev.cpu().numpy()                        # D2H of the result
            else:
                p, _ = ctx.finish()                                        # D2H of the result
                host_pwr[:] = p

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # the e2e result must be a real spectrum too (Parseval: mean bin power ~ spectra * N * 2 sigma^2)
        e2e_expect = world * e2e_spec * N_BINS * (2 * 30.0 ** 2)
        assert np.all(np.isfinite(host_pwr)) and 0.5 * e2e_expect < host_pwr.mean() < 1.5 * e2e_expect, (host_pwr.mean(), e2e_expect)
        e2e = {"value": world * e2e_spec * N_BINS * e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": world * E2E_PASSES * E2E_RING_BYTES, "d2h_bytes_per_step": world * N_BINS * 8,
               "steps": e2e_steps,
               "sample": "%d spectra per GPU per step from a %d MiB pinned host ring, rpf_push in %d MiB chunks, "
                         "host wall clock (max over ranks)" % (e2e_spec, E2E_RING_BYTES >> 20, E2E_CHUNK >> 20)}
        pkg.lib().rpf_host_free(ring_ptr)
    clocks = sampler.stop() if sampler else None      # covers the device-timed and the e2e regions
    ctx.close()

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = 2.0 * nspec * N_BINS / (kern_ms * 1e-3) / 1e9          # algorithmic GB/s of one launch
        tr = load_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "kernel": "pipe4096_kernel", "kernel_ms_per_launch": kern_ms,
                         "algorithmic_bytes_per_launch": 2 * nspec * N_BINS,
                         "note": "2 B per complex sample; FP32-issue and shared-memory bounds in DESIGN.md"},
            "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        elif not args.no_cpu:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
