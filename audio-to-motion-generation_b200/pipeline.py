"""The composed hot path: waveform -> log-mel -> (adapter D2) -> SelfAttention_G -> L1 / PCK partials,
clip-sharded over the GPUs of one box with a single 64-byte all-reduce at the end.

Nothing in the reference connects these stages; this module is the composition the
build defines, written against the drop-in modules of this package:

    mel    = audio_repr.log_mel_spectograms(wav)            # [B, 425, 64]
    x      = mel[:, 0:384:6, :]                             # D2: the reference's strided-slice feed, in place
    pose,_ = SelfAttention_G(x)                             # [B, 64, 104]
    motion_evaluation.evaluate_poses(pose, gt, accum=...)   # PCK hits, |d pose|, |d motion| partial sums

One process per GPU (torchrun); ranks own contiguous clip ranges; weights are replicated; the only
collective is ``allreduce_metrics`` (NCCL over NVLink through liba2m_b200's run-time binding).
"""
import ctypes

import torch

from . import _cabi, motion_evaluation, pats_audio
from .pose_video import audio_repr

ADAPTER_STRIDE = 6          # dataUtils.py:654 strided slice, fs_ratio = 6 (audio.py:177)
POSE_FRAMES = 64
CLIP_SAMPLES = 68267        # 64 / 15 s at 16 kHz -> 425 mel frames


def shard_range(n_items, rank, world):
    """Contiguous clip range [lo, hi) of `rank`: rank r takes [r*ceil(n/W), (r+1)*ceil(n/W))."""
    per = -(-n_items // world)
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


WINDOW_HOP = 5              # version5_model_train.py:205 (window_hop): consecutive windows start 5 * fs_ratio frames apart


def window_starts(n_frames, frames=POSE_FRAMES, stride=ADAPTER_STRIDE, window_hop=WINDOW_HOP):
    """First log-mel frame of every sliding window, the reference's index arithmetic (MiniData.update_idx_list,
    dataUtils.py:585-620): window = frames * stride feature rows, starts = range(0, len - window, window_hop * stride)."""
    return list(range(0, n_frames - frames * stride, window_hop * stride))


def sliding_windows(logmel_clip, frames=POSE_FRAMES, stride=ADAPTER_STRIDE, window_hop=WINDOW_HOP):
    """One clip's log-mel [n_frames, F] -> all its model inputs [n_windows, frames, F] as a strided VIEW
    (window w, step t = row w * window_hop * stride + t * stride): nothing is gathered or copied, the first
    encoder kernel reads the overlapping windows in place."""
    n, f = logmel_clip.shape
    n_win = len(window_starts(n, frames, stride, window_hop))
    if n_win == 0:
        return logmel_clip.new_empty((0, frames, f))
    s0 = logmel_clip.stride(0)
    return logmel_clip.as_strided((n_win, frames, f), (window_hop * stride * s0, stride * s0, 1),
                                  logmel_clip.storage_offset())


def adapter(logmel, frames=POSE_FRAMES, stride=ADAPTER_STRIDE):
    """D2: [B, >=frames*stride, F] -> strided view [B, frames, F] (no copy)."""
    span = frames * stride
    if logmel.shape[1] < span - stride + 1:
        raise ValueError("log-mel has %d frames, the adapter needs at least %d" % (logmel.shape[1], span - stride + 1))
    return logmel[:, 0:span:stride, :]


class Communicator:
    """NCCL communicator owned by liba2m_b200 (a2m_comm_*); the unique id travels over torch.distributed."""

    def __init__(self, rank, world, device):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        ident = (ctypes.c_char * 128)()
        if rank == 0:
            _cabi.check(_cabi.lib().a2m_comm_unique_id(ctypes.cast(ident, ctypes.c_void_p)))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        buf = (ctypes.c_char * 128).from_buffer_copy(box[0])
        out = ctypes.c_void_p()
        _cabi.check(_cabi.lib().a2m_comm_init(ctypes.cast(buf, ctypes.c_void_p), rank, world, device.index,
                                              ctypes.byref(out)))
        self.ptr, self.device = out, device

    def allreduce(self, accum):
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().a2m_allreduce_metrics(self.ptr, _cabi.ptr(accum), _cabi.stream_ptr(self.device)))

    def allreduce_smoothness(self, accum):
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().a2m_allreduce_smoothness(self.ptr, _cabi.ptr(accum), _cabi.stream_ptr(self.device)))

    def close(self):
        if self.ptr:
            _cabi.lib().a2m_comm_destroy(self.ptr)
            self.ptr = None


def allreduce_metrics(accum, comm=None):
    """Sum the 64-byte partials over the ranks.  CUDA accumulators go through NCCL (a2m_allreduce_metrics);
    host accumulators (the CPU/gloo tests of the sharding logic) through torch.distributed."""
    if comm is not None and accum.is_cuda:
        comm.allreduce(accum)
        return accum
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if accum.is_cuda:
            raise RuntimeError("CUDA accumulator without a Communicator: create one with Communicator(rank, world, device)")
        counts = accum[:5].clone()
        sums = accum[5:7].view(torch.float64).clone()
        dist.all_reduce(counts)
        dist.all_reduce(sums)
        accum[:5] = counts
        accum[5:7] = sums.view(torch.int64)
    return accum


def allreduce_smoothness(accum, comm=None):
    """The same for the 32-byte smoothness / jerk partials (two fp64 sums, two int64 counts)."""
    if comm is not None and accum.is_cuda:
        comm.allreduce_smoothness(accum)
        return accum
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if accum.is_cuda:
            raise RuntimeError("CUDA accumulator without a Communicator: create one with Communicator(rank, world, device)")
        sums = accum[0:2].view(torch.float64).clone()
        counts = accum[2:4].clone()
        dist.all_reduce(sums)
        dist.all_reduce(counts)
        accum[0:2] = sums.view(torch.int64)
        accum[2:4] = counts
    return accum


class AudioToPosePipeline:
    """mel -> generator -> evaluation on one GPU.  `model` is a SelfAttention_G drop-in in eval mode on `device`."""

    FRONT_ENDS = ("vggish", "log_mel_400", "log_mel_512")

    def __init__(self, model, alpha=0.2, comm=None, lanes=2, graphs=False, smoothness=False, front_end="vggish",
                 sample_rate=None, adapter_frames_only=False):
        """`lanes` > 1 runs consecutive batches on alternating CUDA streams, each lane with its own native handle
        (packed weights and activation arena) built from the SAME module, so the latency-bound tail of one batch
        (graph decoders, small GEMMs) overlaps the head of the next, and a later load_state_dict / weight edit /
        set_output_denorm on `model` reaches every lane.  `graphs=True` captures each lane's whole step (about 60 launches, the
        two-stream decoder fork and the programmatic-dependent-launch edges included) into a CUDA graph per input
        shape and replays it; inputs are copied into the graph's static buffers.  Results depend on neither.
        `front_end` selects the audio features: "vggish" (pose_video/audio_repr.py, 64 bands -- the path BASELINE's
        configs name), or the PATS-native "log_mel_400" (64 bands) / "log_mel_512" (128 bands at `sample_rate`, the
        representation the shipped training configuration reads; pats/data_loading/audio.py).  All are fed to the
        generator through the same stride-6 adapter.
        `adapter_frames_only=True` ("vggish" front end) computes only the log-mel frames the stride-6 adapter feeds to the
        generator (hop 60 ms = 6 x 10 ms: frame t of that transform is frame 6 t of the full one, bit for bit) instead of
        all 425 -- an end-to-end shortcut; off by default, and never used by bench.py's headline.
        `smoothness=True` also accumulates the validation loop's temporal-smoothness and jerk metrics of the generated
        poses (version5_model_train.py:456-459), one more small kernel per step."""
        if front_end not in self.FRONT_ENDS:
            raise ValueError("front_end must be one of %s, got %r" % (self.FRONT_ENDS, front_end))
        self.front_end = front_end
        if adapter_frames_only and front_end != "vggish":
            raise ValueError("adapter_frames_only is implemented for the 'vggish' front end")
        self.adapter_frames_only = bool(adapter_frames_only)
        self.sample_rate = sample_rate if sample_rate is not None else (44100 if front_end == "log_mel_512" else 16000)
        self.model = model
        self.alpha = alpha
        self.comm = comm
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("AudioToPosePipeline needs the model on a CUDA device; there is no CPU fallback")
        self.accum = motion_evaluation.new_metrics(self.device)
        self.smooth = motion_evaluation.new_smoothness(self.device) if smoothness else None
        self._copy_stream = torch.cuda.Stream(self.device)
        self._staging = {}                      # (wav shape, gt shape, depth) -> ring of [wav_dev, gt_dev, last-use event]
        self._readback = None                   # pinned ring for the per-step read-back of the running metrics
        self._n_lanes = max(1, int(lanes))
        self._lane_streams = [torch.cuda.Stream(self.device) for _ in range(self._n_lanes)]
        self._turn = 0
        self._use_graphs = bool(graphs)
        self._graphs = {}                       # (lane, wav shape, gt shape) -> (graph, static wav, gt, pose, kernels per replay)
        self.replayed_launches = 0              # kernels launched through graph replays (a2m_launch_count only sees host launches)

    def reset(self):
        self.sync_lanes()
        self.accum.zero_()
        if self.smooth is not None:
            self.smooth.zero_()

    def sync_lanes(self):
        """Make the caller's current stream wait for everything the lanes have been given so far."""
        cur = torch.cuda.current_stream(self.device)
        for st in self._lane_streams:
            cur.wait_stream(st)

    def features(self, wav):
        """wav [B, N] -> the front end's log-mel [B, frames, 64 | 128]."""
        if self.front_end == "log_mel_512":
            return pats_audio.log_mel_512(wav, self.sample_rate)
        if self.front_end == "log_mel_400":
            return pats_audio.log_mel_400(wav, self.sample_rate)
        return audio_repr.log_mel_spectograms(wav, audio_sample_rate=self.sample_rate)

    def generate(self, wav, lane=0):
        """wav [B, N] fp32 (or int16 PCM) CUDA tensor -> pose [B, 64, 104] fp32 (on the current stream)."""
        if self.adapter_frames_only:
            # frames 0, 6, 12, ... only: same window, hop = ADAPTER_STRIDE x 10 ms
            logmel = audio_repr.log_mel_spectograms(wav, audio_sample_rate=self.sample_rate,
                                                    hop_length_secs=0.010 * ADAPTER_STRIDE)
            if logmel.shape[1] < POSE_FRAMES:
                raise ValueError("the audio gives %d adapter frames, the generator needs %d" % (logmel.shape[1], POSE_FRAMES))
            x = logmel[:, :POSE_FRAMES, :]
        else:
            x = adapter(self.features(wav))
        pose, _ = self.model(x, lane=lane)
        return pose

    def generate_long(self, wav, window_hop=WINDOW_HOP, lane=0):
        """Long-form audio (BASELINE config 4): wav [B, N] (e.g. 60 s = 960 000 samples) -> poses
        [B, n_windows, 64, 104].  The log-mel of every stream is computed once; each clip's overlapping windows
        (384-frame span, stride 6, hop window_hop * 6 frames -- the reference's window arithmetic) are fed to the
        generator in place by ONE launch program for all streams (SelfAttention_G.forward_windows), so no window is ever
        materialised and nothing loops over clips."""
        logmel = self.features(wav)                                        # [B, frames, 64 | 128]
        n_win = len(window_starts(logmel.shape[1], window_hop=window_hop))
        if n_win == 0:
            raise ValueError("the audio is shorter than one window (%d log-mel frames)" % logmel.shape[1])
        per_call = max(1, 65535 // n_win)                                  # streams per launch program
        out = [self.model.forward_windows(logmel[b:b + per_call], n_win, POSE_FRAMES, ADAPTER_STRIDE, window_hop, lane=lane)
               for b in range(0, logmel.shape[0], per_call)]
        return out[0] if len(out) == 1 else torch.cat(out)

    def step(self, wav, gt_pose, done_event=None, readback=None):
        """One batch, inputs already on the device: enqueues mel -> generator -> evaluation on the next lane and
        accumulates the metric partials.  Returns the poses; they (and the metrics) are complete once
        ``sync_lanes()`` / ``finish()`` has been called on the consuming stream."""
        lane = self._turn % self._n_lanes
        self._turn += 1
        st = self._lane_streams[lane]
        st.wait_stream(torch.cuda.current_stream(self.device))      # inputs were produced on the caller's stream
        with torch.cuda.stream(st):
            if self._use_graphs:
                pose = self._replay(lane, st, wav, gt_pose)
            else:
                pose = self.generate(wav, lane)
                motion_evaluation.evaluate_poses(pose, gt_pose, self.alpha, accum=self.accum)
            if self.smooth is not None:
                motion_evaluation.evaluate_smoothness(pose, accum=self.smooth, from_pose=True)
            if readback is not None:                                # running metric partials of this step, 64 bytes D2H
                readback.copy_(self.accum, non_blocking=True)
            if done_event is not None:
                done_event.record(st)                               # the lane no longer reads wav / gt_pose after this
        wav.record_stream(st)
        gt_pose.record_stream(st)
        return pose

    def _replay(self, lane, st, wav, gt_pose):
        """Run one step of `lane` through its CUDA graph (captured on first use for this input shape).  Called with
        `st` current.  The returned poses live in the graph's static output buffer: they are overwritten by the
        lane's next step."""
        key = (lane, tuple(wav.shape), wav.dtype, tuple(gt_pose.shape))
        entry = self._graphs.get(key)
        handle = self.model.native(lane)
        # a captured graph has the arena pointers of its launch plan baked in: if the handle was rebuilt (weights
        # changed) or released a cached plan since the capture, capture again
        if entry is not None and (entry[5] is not handle or
                                  entry[6] != int(_cabi.lib().a2m_model_plan_generation(handle.ptr))):
            entry = None
        if entry is None:
            s_wav = torch.empty(wav.shape, dtype=wav.dtype, device=self.device)
            s_gt = torch.empty(gt_pose.shape, dtype=torch.float32, device=self.device)
            s_wav.copy_(wav)
            s_gt.copy_(gt_pose)
            scratch = motion_evaluation.new_metrics(self.device)
            for _ in range(2):                  # warm up outside the capture: plans, attributes, allocator pools
                motion_evaluation.evaluate_poses(self.generate(s_wav, lane), s_gt, self.alpha, accum=scratch)
            st.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = _cabi.lib().a2m_launch_count()
            with torch.cuda.graph(graph, stream=st):
                s_pose = self.generate(s_wav, lane)
                motion_evaluation.evaluate_poses(s_pose, s_gt, self.alpha, accum=self.accum)
            handle = self.model.native(lane)
            entry = (graph, s_wav, s_gt, s_pose, int(_cabi.lib().a2m_launch_count() - before), handle,
                     int(_cabi.lib().a2m_model_plan_generation(handle.ptr)))
            self._graphs[key] = entry
        graph, s_wav, s_gt, s_pose, n_kernels = entry[:5]
        s_wav.copy_(wav, non_blocking=True)
        s_gt.copy_(gt_pose, non_blocking=True)
        graph.replay()
        self.replayed_launches += n_kernels
        return s_pose

    def run_host_batches(self, batches, depth=3, per_step_readback=True):
        """End-to-end over HOST batches [(wav_pinned [B,N], gt_pinned [B,64,104]), ...]: batch i+1 is copied to the
        device on a side stream while the kernels of batch i run.  The copies land in a ring of `depth` preallocated
        device buffers per input shape (no allocation inside the loop); a slot is rewritten only after the lane that
        consumed it has finished with it; after every step the running 64-byte metric accumulator is read back into a
        pinned host ring (asynchronously: `finish()` is what waits).  Returns the number of clips processed."""
        main = torch.cuda.current_stream(self.device)
        clips = 0
        it = iter(batches)
        turn = 0

        def stage(pair):
            nonlocal turn
            key = (tuple(pair[0].shape), pair[0].dtype, tuple(pair[1].shape), depth)
            ring = self._staging.get(key)
            if ring is None:
                ring = self._staging[key] = [                       # int16 PCM stays int16: half the bytes over PCIe
                    [torch.empty(pair[0].shape, dtype=pair[0].dtype if pair[0].dtype == torch.int16 else torch.float32,
                                 device=self.device),
                     torch.empty(pair[1].shape, dtype=torch.float32, device=self.device), None] for _ in range(depth)]
            slot = ring[turn % depth]
            turn += 1
            with torch.cuda.stream(self._copy_stream):
                if slot[2] is not None:
                    self._copy_stream.wait_event(slot[2])           # the step that last read this slot is done
                slot[0].copy_(pair[0], non_blocking=True)
                slot[1].copy_(pair[1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return slot, ev

        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        while staged is not None:
            slot, ev = staged
            nxt = next(it, None)
            staged = stage(nxt) if nxt is not None else None
            main.wait_event(ev)
            done = torch.cuda.Event()
            rb = None
            if per_step_readback:
                if self._readback is None:
                    self._readback = [torch.empty(self.accum.shape, dtype=self.accum.dtype).pin_memory() for _ in range(4)]
                rb = self._readback[turn % 4]
            self.step(slot[0], slot[1], done_event=done, readback=rb)
            slot[2] = done
            clips += slot[0].shape[0]
        return clips

    def finish(self):
        """All-reduce (if sharded) and read the 64-byte result: {'pck', 'l1_pose', 'l1_motion', counts...}."""
        self.sync_lanes()
        allreduce_metrics(self.accum, self.comm)
        m = motion_evaluation.read_metrics(self.accum)          # D2H copy: every lane's kernels have completed
        self.model.check_device_status()                        # a kernel's bounded barrier wait expired -> A2MError
        m.update(motion_evaluation.finalize_metrics(m))
        if self.smooth is not None:
            allreduce_smoothness(self.smooth, self.comm)
            m.update(motion_evaluation.read_smoothness(self.smooth))
        return m
