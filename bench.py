#!/usr/bin/env python
"""Benchmark of the audio->pose hot path (mel + SelfAttention_G forward + L1/PCK eval).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic PATS-shaped clips (BASELINE config 2:
256 clips of 68 267 samples @16 kHz -> 425x64 log-mel -> [256,64,64] -> poses [256,64,104] -> L1/PCK
against synthetic ground truth).  Weak scaling: every rank runs its own batch each step (clips shard
naturally, weights replicated); the only collective is the 64-byte metric all-reduce at the end of the
timed region.  Prints ONE JSON line (rank 0).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clips/sec audio->pose (mel+fwd+eval)"
UNIT = "clips/s"
WORKLOAD = ("config2: PATS-shaped batch %d (68267 samples -> 425x64 log-mel -> 64x64 -> 64x104 poses), "
            "mel + SelfAttention_G forward + L1/PCK")
CLIP_SAMPLES = 68267
POOL = 6                     # distinct input batches cycled through: 6 x 77 MB > 126 MB of L2


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel class, summed over its
    launches in one step, from the committed ncu capture (profiles/r1_traffic.json; ncu serialises launches and
    starts each one cold, so this is an upper bound for the pipelined run).  None if the file is absent."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples taken before this call are warm-up samples; they are only used if the timed region was too short
        for nvidia-smi to report during it."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.first:] if len(self.rows) - self.first >= 2 else self.rows[max(0, self.first - 3):]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(n_clips, threads):
    """The reference's CPU path through the oracle port: per-clip fp64 log-mel (the reference has no batch
    dimension), adapter D2, fp32 generator forward in eval mode, compute_pck + L1.  Returns seconds."""
    import numpy as np
    import torch
    from oracle import mel_oracle, model_oracle, eval_oracle, synth, weights
    torch.set_num_threads(threads)
    sd = cpu_reference_run.sd if hasattr(cpu_reference_run, "sd") else weights.make_state_dict(0, "stress")
    cpu_reference_run.sd = sd
    wav = synth.wav_batch(0, n_clips)
    gt = synth.gt_pose_batch(0, n_clips)
    t0 = time.perf_counter()
    mel = np.stack([mel_oracle.log_mel_audio_repr(w) for w in wav])
    x = torch.from_numpy(synth.adapter(mel).astype(np.float32))
    pose, _ = model_oracle.generator_forward(sd, x)
    eval_oracle.finalize(eval_oracle.metric_partials(pose.numpy(), gt))
    return time.perf_counter() - t0


def run_reference(args, rank, world, out=sys.stdout):
    """--impl reference: the reference algorithm on the host cores (oracle port; the reference itself is
    pure Python and not importable as shipped)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(4, min(args.ref_clips, 32))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(sample, threads)
    t = [cpu_reference_run(sample, threads) for _ in range(args.steps)]
    total = sum(t)
    value = sample * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 mel / f32 model", "data": "synthetic",
            "config": {"workload": WORKLOAD % args.batch, "clips_per_step": sample, "clip_samples": CLIP_SAMPLES,
                       "note": "the reference's CPU path (oracle port) on a bounded sample of the same workload per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d clips per step x %d steps (oracle port of the reference CPU path)" % (sample, args.steps)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def bind_to_gpu_cpus(index):
    """One process per GPU: run (and first-touch the pinned staging memory) on the CPU cores NVML reports as local to
    this GPU, so that eight ranks do not pull their host batches across the socket interconnect.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cores local to GPU %d" % (len(cpus), index)
    except Exception as exc:                      # no NVML, restricted container, ...: keep the default placement
        return "unbound (%s)" % type(exc).__name__
    return "unbound"


def _claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1: point fd 1 at stderr for the run and return a file on
    the real stdout, which then carries exactly one line -- the JSON result."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ref-clips", type=int, default=16)
    ap.add_argument("--lanes", type=int, default=2, help="CUDA-stream lanes consecutive batches alternate over")
    ap.add_argument("--graphs", type=int, default=0, help="1: replay each lane's step as a CUDA graph (measured: no gain "
                    "over two eager stream lanes, which already hide the launch gaps)")
    ap.add_argument("--adapter-frames-only", type=int, default=0, help="1: compute only the 64 log-mel frames the adapter "
                    "feeds to the generator (an end-to-end shortcut; NOT the headline configuration, which produces all 425)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    out = _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world, out)
    args.warmup = max(args.warmup, 3, args.lanes)      # every lane builds its launch program on its first step

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None      # pinned host batches on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    a2m = importlib.import_module("audio-to-motion-generation_b200")
    cabi = importlib.import_module("audio-to-motion-generation_b200._cabi")
    pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
    rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
    lib = a2m.load_library()

    # ---- model: random-init weights of the reference architecture (no checkpoints offline) ----------
    torch.manual_seed(0)
    model = rmm.SelfAttention_G()
    with torch.no_grad():                       # make every attention / BatchNorm path non-trivial
        g = torch.Generator().manual_seed(1)
        for name, p in model.state_dict().items():
            if name.endswith("gamma"):
                p.fill_(0.5)
            elif name.endswith("running_var"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif name.endswith("running_mean"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    model = model.to(device).eval()
    comm = pipeline.Communicator(rank, world, device) if world > 1 else None
    pipe = pipeline.AudioToPosePipeline(model, comm=comm, lanes=args.lanes, graphs=bool(args.graphs),
                                        adapter_frames_only=bool(args.adapter_frames_only))

    # ---- synthetic inputs: per-clip seeds make any sharding reproduce the same clips -----------------
    B = args.batch
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    wav_dev = [0.1 * torch.randn(B, CLIP_SAMPLES, device=device, generator=gen) for _ in range(POOL)]
    gt_dev = [50.0 * torch.randn(B, 64, 104, device=device, generator=gen) for _ in range(POOL)]
    wav_host = [w.cpu().pin_memory() for w in wav_dev]
    gt_host = [g_.cpu().pin_memory() for g_ in gt_dev]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ---- device-resident arm -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                          # nvidia-smi needs ~0.1 s to produce its first line: start it before the warm-up
    for i in range(args.warmup):
        pipe.step(wav_dev[i % POOL], gt_dev[i % POOL])
    pipe.finish()
    model.check_device_status()
    pipe.reset()
    barrier()
    if rank == 0:
        sampler.mark()                           # samples from here on are "under load"
    lib.a2m_launch_count_reset()
    pipe.replayed_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe.step(wav_dev[i % POOL], gt_dev[i % POOL])
    result = pipe.finish()                       # all-reduce + 64-byte D2H inside the timed region
    e1.record()
    barrier()
    launches = int(lib.a2m_launch_count()) + int(pipe.replayed_launches)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end-to-end arm: pinned host inputs, H2D inside the timed region, result read back ------------
    pipe.reset()
    pipe.run_host_batches([(wav_host[i % POOL], gt_host[i % POOL]) for i in range(2)])
    pipe.finish(); pipe.reset()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    pipe.run_host_batches((wav_host[i % POOL], gt_host[i % POOL]) for i in range(args.steps))
    result_e2e = pipe.finish()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    wall_e2e = time.perf_counter() - t0
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = wav_host[0].numel() * 4 + gt_host[0].numel() * 4

    # ---- roofline of the dominant kernel (tcgen05 conv GEMM) and of the two HBM-bound kernels ----------
    pk = peaks()
    roof = None
    extra = {}
    if rank == 0:
        import ctypes
        logmel = pipeline.audio_repr.log_mel_spectograms(wav_dev[0])
        x = pipeline.adapter(logmel)
        out_ms = (ctypes.c_float * 3)()
        n_gemm = ctypes.c_int()
        h = model.native()
        cabi.check(lib.a2m_model_profile(h.ptr, cabi.ptr(x), x.stride(0), x.stride(1), B, 64, 64, max(3, min(args.steps, 10)),
                                         out_ms, ctypes.byref(n_gemm), cabi.stream_ptr(device)))
        flops = int(lib.a2m_model_gemm_flops(h.ptr, B, 64, 64))
        achieved = flops / (out_ms[1] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, %d launches/step)" % n_gemm.value,
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                "peak_kind": pk["src"] + " sustained (kernel timed inside a long step)",
                "traffic": (ncu_traffic() or {}).get("conv_gemm_bytes_per_step"),
                "traffic_note": (ncu_traffic() or {}).get("note"),
                "gemm_ms_per_step": out_ms[1], "other_ms_per_step": out_ms[2], "algorithmic_gflop_per_step": flops / 1e9}

        def ev_time(fn, n=10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(0); torch.cuda.synchronize()
            a.record()
            for i in range(n):
                fn(i)
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        mel_ms = ev_time(lambda i: pipeline.audio_repr.log_mel_spectograms(wav_dev[i % POOL]))
        mel_bytes = B * (CLIP_SAMPLES * 4 + 425 * 64 * 4)
        acc = pipe.accum
        ev_ms = ev_time(lambda i: pipeline.motion_evaluation.evaluate_poses(gt_dev[i % POOL], gt_dev[(i + 1) % POOL], accum=acc))
        ev_bytes = B * 64 * 104 * 4 * 2
        extra = {"roofline_mel": {"bound": "hbm", "achieved": mel_bytes / mel_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": mel_bytes / mel_ms / 1e6 / pk["hbm"], "ms": mel_ms,
                                  "note": "HBM-bound by contract, fp32-issue / shared-memory bound in practice (DESIGN.md section 4): "
                                          "the FFT butterflies alone need more issue slots than 50 % of the HBM roofline leaves"},
                 "roofline_eval": {"bound": "hbm", "achieved": ev_bytes / ev_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": ev_bytes / ev_ms / 1e6 / pk["hbm"], "ms": ev_ms}}
        pipe.reset()

    # ---- CPU baseline (oracle port of the reference path) on the host cores, rank 0, N = 1 -----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_run(2, threads)
        n, spent, runs = args.ref_clips, 0.0, 0
        while spent < 12.0 and runs < 200:
            spent += cpu_reference_run(n, threads)
            runs += 1
        cpu = {"value": n * runs / spent, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d runs x %d clips of the same workload (fp64 mel per clip + fp32 generator + PCK/L1), %.1f s" % (runs, n, spent)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 (fp32 accumulate; mel and eval fp32)", "data": "synthetic",
                "config": {"workload": WORKLOAD % B,
                           "clips_per_step_per_gpu": B, "parallelism": "clip-sharded x%d" % world,
                           "stream_lanes": args.lanes, "cuda_graphs": bool(args.graphs), "host_binding": numa,
                           "mel_frames": "64 adapter frames only (shortcut)" if args.adapter_frames_only else "all 425 per clip",
                           "l2": "inputs cycle through %d distinct batches (%.0f MB each) > L2" % (POOL, h2d / 1e6)},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                        "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps},
                "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cpu,
                "result": {k: result[k] for k in ("pck", "l1_pose", "l1_motion", "n_frames")},
                "result_e2e": {k: result_e2e[k] for k in ("pck", "n_frames")}}
        line.update(extra)
        print(json.dumps(line), file=out, flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
