#!/usr/bin/env python
"""Benchmark of the audio->pose hot path (mel + SelfAttention_G forward + L1/PCK eval).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 256] [--config 2|3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic PATS-shaped clips (BASELINE config 2:
256 clips of 68 267 samples @16 kHz -> 425x64 log-mel -> [256,64,64] -> poses [256,64,104] -> L1/PCK
against synthetic ground truth).  Weak scaling: every rank runs its own batch each step (clips shard
naturally, weights replicated); the only collective is the 64-byte metric all-reduce at the end of the
timed region.  Prints ONE JSON line (rank 0).

Every synthetic clip is a function of its GLOBAL index only (torch.Generator seeded with 1234 + index for the waveform,
4321 + index for the ground truth), so any sharding sees the same clips.  The line therefore also
carries `sharded_check`: a fixed evaluation set (clips 0 .. 2047) split over the N ranks, whose integer hit count is the
same for every N.  `--config 3` runs BASELINE config 3 instead: the 100 000-clip evaluation set, STRONG scaling (rank r
takes clips [r * ceil(n / N), (r + 1) * ceil(n / N))), one all-reduce at the end.
"""
import argparse
import faulthandler
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


def synth_clips(lo, hi, device, dtype=None):
    """Clips [lo, hi) of the synthetic set, generated ON THE DEVICE from per-clip seeds: wav 0.1 * N(0, 1) (seed
    1234 + index), ground truth 50 * N(0, 1) (seed 4321 + index).  dtype torch.int16: the waveform as 16-bit PCM
    (round(wav * 32767 / 0.5), i.e. +-0.5 full scale -- 5 sigma)."""
    import torch
    n = hi - lo
    wav = torch.empty(n, CLIP_SAMPLES, device=device)
    gt = torch.empty(n, 64, 104, device=device)
    g = torch.Generator(device=device)
    for i in range(n):
        g.manual_seed(1234 + lo + i)
        wav[i].normal_(0.0, 0.1, generator=g)
        g.manual_seed(4321 + lo + i)
        gt[i].normal_(0.0, 50.0, generator=g)
    if dtype is not None and dtype == torch.int16:
        wav = (wav * (32767.0 / 0.5)).round_().clamp_(-32768, 32767).to(torch.int16)
    return wav, gt


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel class, summed over its
    launches in one step, from the committed ncu capture (profiles/r1_traffic.json; ncu serialises launches and
    starts each one cold, so this is an upper bound for the pipelined run).  None if the file is absent."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
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


def cpu_reference_run(n_clips, threads, first_clip=0):
    """The reference's CPU path through the oracle port: per-clip fp64 log-mel (the reference has no batch
    dimension; the clips are spread over the host threads -- numpy's FFT releases the GIL), adapter D2, fp32 generator
    forward in eval mode on all threads, compute_pck + L1.  Returns seconds."""
    import concurrent.futures
    import numpy as np
    import torch
    from oracle import mel_oracle, model_oracle, eval_oracle, synth, weights
    torch.set_num_threads(threads)
    sd = cpu_reference_run.sd if hasattr(cpu_reference_run, "sd") else weights.make_state_dict(0, "stress")
    cpu_reference_run.sd = sd
    wav = synth.wav_batch(first_clip, n_clips)
    gt = synth.gt_pose_batch(first_clip, n_clips)
    t0 = time.perf_counter()
    if n_clips >= 4 and threads > 1:
        with concurrent.futures.ThreadPoolExecutor(max_workers=threads) as ex:
            mel = np.stack(list(ex.map(mel_oracle.log_mel_audio_repr, wav)))
    else:
        mel = np.stack([mel_oracle.log_mel_audio_repr(w) for w in wav])
    x = torch.from_numpy(synth.adapter(mel).astype(np.float32))
    pose, _ = model_oracle.generator_forward(sd, x)
    eval_oracle.finalize(eval_oracle.metric_partials(pose.numpy(), gt))
    return time.perf_counter() - t0


def run_reference(args, rank, world, out=sys.stdout):
    """--impl reference: the reference algorithm on the host cores (oracle port; the reference itself is pure Python
    and not importable as shipped).  BASELINE.md section 4: one step = ONE batch of config 2
    (256 clips, the same configuration as the B200 arm); the config-1 latency (batch 1) is reported beside it.
    If a step takes so long that K steps would not end within a few minutes, the remaining steps run on a
    quarter batch and the line says so."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.batch
    budget_s = 240.0
    cpu_reference_run(4, threads)                                            # imports, weights, thread pools
    warm = min(args.warmup, 1)
    times, clips = [], []
    t_first = cpu_reference_run(B, threads, 0) if warm else None              # the warm-up step doubles as the probe
    per_step = t_first if t_first is not None else cpu_reference_run(B, threads, 0)
    n_step = B if per_step * args.steps <= budget_s else max(16, B // 4)
    for i in range(args.steps):
        times.append(cpu_reference_run(n_step, threads, (i + 1) * B))
        clips.append(n_step)
    total = sum(times)
    value = sum(clips) / total
    b1 = sorted(cpu_reference_run(1, threads, 7) for _ in range(5))[2]
    sample = "%d steps x %d clips (oracle port of the reference CPU path, %d threads)" % (args.steps, n_step, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 mel / f32 model", "data": "synthetic",
            "config": {"workload": WORKLOAD % B, "clips_per_step": n_step, "clip_samples": CLIP_SAMPLES,
                       "note": "the reference's CPU path (oracle port); one step = one batch of the workload"
                               + ("" if n_step == B else " -- reduced to %d clips per step to bound the run" % n_step)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "config1_batch1": {"ms_per_clip": 1e3 * b1, "clips_per_s": 1.0 / b1,
                               "note": "BASELINE config 1: one 4.27 s clip, log-mel + forward + eval, median of 5"},
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
        if len(cpus) >= 8:                        # never squeeze a rank (main thread + NCCL proxy threads) onto a few cores
            os.sched_setaffinity(0, cpus)
            return "%d cores local to GPU %d" % (len(cpus), index)
        return "unbound (%d local cores allowed)" % len(cpus)
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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3], help="2: the headline (weak scaling, batches of 256 "
                    "per rank and step); 3: BASELINE config 3, the 100 000-clip evaluation set strong-scaled over the ranks")
    ap.add_argument("--eval-clips", type=int, default=100000, help="size of the config-3 evaluation set")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the sustained leg (0: skip)")
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
    faulthandler.dump_traceback_later(300, exit=True)          # watchdog (re-armed per phase below): never hang the node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        print("[bench rank %d] process group up" % rank, file=sys.stderr, flush=True)

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
    if world > 1:
        print("[bench rank %d] communicator up" % rank, file=sys.stderr, flush=True)
    pipe = pipeline.AudioToPosePipeline(model, comm=comm, lanes=args.lanes, graphs=bool(args.graphs),
                                        adapter_frames_only=bool(args.adapter_frames_only))

    B = args.batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.perf_counter()

    def phase(name):
        """progress marker on stderr (stdout carries only the JSON line) + watchdog: a phase that makes no progress for
        five minutes dumps every thread's stack and ends the process instead of hanging the node"""
        print("[bench rank %d +%.1fs] %s" % (rank, time.perf_counter() - t_start, name), file=sys.stderr, flush=True)
        faulthandler.dump_traceback_later(300, exit=True)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    def sharded_eval(n_clips):
        """Strong-scaled evaluation of clips [0, n_clips): rank r owns [r * ceil(n / W), (r + 1) * ceil(n / W)), batches
        of B, one all-reduce.  Inputs are generated into HBM before the timed region.  Returns (ms, result)."""
        lo, hi = pipeline.shard_range(n_clips, rank, world)
        chunks = [synth_clips(c, min(hi, c + B), device) for c in range(lo, hi, B)]
        pipe.reset()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for w, g_ in chunks:
            pipe.step(w, g_)
        res = pipe.finish()
        t1.record()
        barrier()
        return max_over_ranks(t0.elapsed_time(t1)), res

    if args.config == 3:
        # ---- BASELINE config 3: the full synthetic evaluation set, clip-sharded, strong scaling -------------
        for i in range(max(args.warmup, args.lanes)):
            w, g_ = synth_clips(i * B, (i + 1) * B, device)
            pipe.step(w, g_)
        pipe.finish()
        lib.a2m_launch_count_reset()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms, res = sharded_eval(args.eval_clips)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            line = {"metric": METRIC, "value": args.eval_clips / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                    "steps": -(-args.eval_clips // (B * world)), "warmup": max(args.warmup, args.lanes),
                    "ms_per_step": ms / max(1, -(-args.eval_clips // (B * world))), "ms_total": ms, "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "bf16 (fp32 accumulate; mel and eval fp32)",
                    "data": "synthetic",
                    "config": {"workload": "config3: full synthetic eval set, %d clips with L1/PCK, clip-sharded x%d, "
                                           "batches of %d, per-clip seeds" % (args.eval_clips, world, B),
                               "parallelism": "clip-sharded x%d" % world, "stream_lanes": args.lanes,
                               "l2": "each rank's shard (%.1f GB of waveforms) is read once" %
                                     (-(-args.eval_clips // world) * CLIP_SAMPLES * 4 / 1e9)},
                    "clocks": clocks, "gpu_launches": int(lib.a2m_launch_count()),
                    "result": {k: res[k] for k in ("pck_hits", "n_keypoints", "pck", "l1_pose", "l1_motion", "n_frames")}}
            print(json.dumps(line), file=out, flush=True)
        faulthandler.cancel_dump_traceback_later()
        if comm is not None:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- synthetic inputs: per-clip seeds make any sharding reproduce the same clips; batch j of rank r holds the
    # global clips [(j * world + r) * B, +B) ---------------------------------------------------------------------
    pairs = [synth_clips((j * world + rank) * B, (j * world + rank + 1) * B, device) for j in range(POOL)]
    wav_dev, gt_dev = [p_[0] for p_ in pairs], [p_[1] for p_ in pairs]
    wav_host = [w.cpu().pin_memory() for w in wav_dev]
    gt_host = [g_.cpu().pin_memory() for g_ in gt_dev]
    pcm_host = [(w * (32767.0 / 0.5)).round().clamp(-32768, 32767).to(torch.int16).cpu().pin_memory() for w in wav_dev]

    def timed_steps(n_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            pipe.step(wav_dev[i % POOL], gt_dev[i % POOL])
        res = pipe.finish()                      # all-reduce + 64-byte D2H inside the timed region
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), res

    # ---- device-resident arm -------------------------------------------------------------------------
    phase("inputs ready")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                          # nvidia-smi needs ~0.1 s to produce its first line: start it before the warm-up
    for i in range(args.warmup):
        pipe.step(wav_dev[i % POOL], gt_dev[i % POOL])
    pipe.finish()
    pipe.reset()
    barrier()
    if rank == 0:
        sampler.mark()                           # samples from here on are "under load"
    lib.a2m_launch_count_reset()
    pipe.replayed_launches = 0
    phase("warm-up done")
    ms_total, result = timed_steps(args.steps)
    phase("timed steps done")
    launches = int(lib.a2m_launch_count()) + int(pipe.replayed_launches)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- sustained leg: the same loop for >= --sustain-seconds (clocks settle under the power cap) ------------
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds / (ms_total / args.steps / 1e3)) + 1)
        pipe.reset()
        barrier()
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
            time.sleep(0.2)
            sampler2.mark()
        ms_sus, _ = timed_steps(n_sus)
        phase("sustained leg done (%d steps)" % n_sus)
        clocks2 = sampler2.stop() if rank == 0 else None
        sustained = {"value": world * B * n_sus / (ms_sus / 1e3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus / 1e3,
                     "ms_per_step": ms_sus / n_sus, "clocks": clocks2}

    # ---- end-to-end arm: pinned host inputs, H2D inside the timed region, result read back ------------
    def e2e_run(wavs):
        pipe.reset()
        pipe.run_host_batches([(wavs[i % POOL], gt_host[i % POOL]) for i in range(2)])
        pipe.finish(); pipe.reset()
        barrier()
        t0 = time.perf_counter()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pipe.run_host_batches((wavs[i % POOL], gt_host[i % POOL]) for i in range(args.steps))
        res = pipe.finish()
        b_.record()
        barrier()
        ms = max_over_ranks(max(a.elapsed_time(b_), 0.0))
        return ms, time.perf_counter() - t0, res

    ms_e2e, wall_e2e, result_e2e = e2e_run(wav_host)
    phase("e2e fp32 done")
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = wav_host[0].numel() * 4 + gt_host[0].numel() * 4
    ms_e2e16, wall_e2e16, result_e2e16 = e2e_run(pcm_host)
    phase("e2e int16 done")
    h2d16 = pcm_host[0].numel() * 2 + gt_host[0].numel() * 4

    # ---- the same fixed evaluation set for every N (per-clip seeds): integer hit counts must not depend on N ----
    check_clips = 2048
    _, check = sharded_eval(check_clips)
    phase("sharded check done")

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
        # a2m_model_profile times every launch between its own pair of events on an otherwise idle GPU: each kernel is
        # timed ALONE, so the burst peak is the denominator; the sustained one is given beside it
        roof = {"bound": "tensor", "kernel": "conv_gemm_pair_kernel / conv_gemm_pair_single_kernel / conv_gemm_kernel (tcgen05 implicit GEMM; cta_group::2 tiles where N %% 256 == 0; %d launches/step)" % n_gemm.value,
                "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"],
                "peak_kind": pk["src"] + " burst (every launch timed alone between its own events)",
                "frac_of_sustained_peak": achieved / pk["tf_sustained"], "sustained_peak": pk["tf_sustained"],
                "traffic": (ncu_traffic() or {}).get("conv_gemm_bytes_per_step"),
                "traffic_note": (ncu_traffic() or {}).get("note"),
                "gemm_ms_per_step": out_ms[1], "other_ms_per_step": out_ms[2], "algorithmic_gflop_per_step": flops / 1e9,
                "whole_forward": {"algorithmic_gflop_per_step": 3.434 * B, "achieved_tflops_in_step": 3.434 * B / (ms_total / args.steps) ,
                                  "frac_of_burst": 3.434 * B / (ms_total / args.steps) / pk["tf_burst"]}}

        def ev_time(fn, n=10):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(0); torch.cuda.synchronize()
            a.record()
            for i in range(n):
                fn(i)
            b_.record(); torch.cuda.synchronize()
            return a.elapsed_time(b_) / n
        lm = pipeline.audio_repr.log_mel_spectograms
        mel_ms = ev_time(lambda i: lm(wav_dev[i % POOL]))
        mel_bytes = B * (CLIP_SAMPLES * 4 + 425 * 64 * 4)
        pcm_dev = [p_.to(device) for p_ in pcm_host[:3]]
        mel16_ms = ev_time(lambda i: lm(pcm_dev[i % 3]))
        mel16_bytes = B * (CLIP_SAMPLES * 2 + 425 * 64 * 4)
        sweep = []
        for nb in (1, 16, 4096):                 # BASELINE config 5 (bandwidth sweep); 256 is the line above
            wv = torch.empty(nb, CLIP_SAMPLES, device=device).normal_(0.0, 0.1)
            ms = ev_time(lambda i: lm(wv), n=5)
            sweep.append({"clips": nb, "ms": ms, "gbs": nb * (CLIP_SAMPLES * 4 + 425 * 64 * 4) / ms / 1e6,
                          "frac": nb * (CLIP_SAMPLES * 4 + 425 * 64 * 4) / ms / 1e6 / pk["hbm"]})
            del wv
        acc = pipe.accum
        ev_ms = ev_time(lambda i: pipeline.motion_evaluation.evaluate_poses(gt_dev[i % POOL], gt_dev[(i + 1) % POOL], accum=acc))
        ev_bytes = B * 64 * 104 * 4 * 2
        extra = {"roofline_mel": {"bound": "hbm", "achieved": mel_bytes / mel_ms / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": mel_bytes / mel_ms / 1e6 / pk["hbm"], "ms": mel_ms,
                                  "int16_pcm": {"achieved": mel16_bytes / mel16_ms / 1e6, "frac": mel16_bytes / mel16_ms / 1e6 / pk["hbm"],
                                                "ms": mel16_ms, "bytes_per_clip": CLIP_SAMPLES * 2 + 425 * 64 * 4},
                                  "sweep": sweep,
                                  "note": "HBM-bound by contract; in practice shared-memory-crossbar / issue bound (DESIGN.md "
                                          "section 4): ncu shows 68 % of the shared-memory wavefront peak, 61 % issue, 42 % fp32 pipe"},
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
                           "clip_seeds": "per clip: 1234 + global index (wav), 4321 + global index (ground truth)",
                           "l2": "inputs cycle through %d distinct batches (%.0f MB each) > L2" % (POOL, h2d / 1e6)},
                "clocks": clocks,
                "value_sustained": sustained,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                        "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": 1e3 * wall_e2e / args.steps,
                        "input": "fp32 waveforms from pinned host memory"},
                "e2e_int16": {"value": world * B * args.steps / (ms_e2e16 / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d16,
                              "d2h_bytes_per_step": 64, "ms_per_step": ms_e2e16 / args.steps,
                              "input": "16-bit PCM waveforms (the source format; a2m_logmel_i16) from pinned host memory",
                              "result": {k: result_e2e16[k] for k in ("pck", "n_frames")}},
                "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cpu,
                "cpu_baseline_note": None if cpu is not None else "measured on rank 0 at N = 1 only (see the N = 1 line)",
                "result": {k: result[k] for k in ("pck", "l1_pose", "l1_motion", "n_frames")},
                "result_e2e": {k: result_e2e[k] for k in ("pck", "n_frames")},
                "sharded_check": {"clips": check_clips, "note": "clips 0 .. %d split over the %d rank(s), per-clip seeds: "
                                  "identical integers for every N" % (check_clips - 1, world),
                                  "pck_hits": check["pck_hits"], "n_keypoints": check["n_keypoints"], "pck": check["pck"],
                                  "l1_pose": check["l1_pose"]}}
        line.update(extra)
        print(json.dumps(line), file=out, flush=True)
    faulthandler.cancel_dump_traceback_later()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
