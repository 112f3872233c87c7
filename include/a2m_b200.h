/* liba2m_b200.so -- C ABI of the B200-native audio->pose hot path.
 *
 * The reference (Xukai-UoA/Audio-to-Motion-Generation) is pure Python with no FFI; its boundary is
 * the Python call surface.  The Python drop-in modules of this repo bind the
 * entry points below with ctypes; each entry point names the reference interface it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - return 0 = OK, <0 = argument / state error, >0 = cudaError_t (or 1000 + ncclResult_t);
 *     a2m_last_error() returns a thread-local message for the last non-zero return.
 *   - all data pointers are DEVICE pointers unless the name says host; the caller owns every buffer.
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*); no hidden device syncs
 *     except in the *_create functions.
 *   - there is no CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef A2M_B200_H
#define A2M_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2M_VERSION 100            /* 0.1.0 */
#define A2M_OK 0
#define A2M_ERR_ARGUMENT (-1)
#define A2M_ERR_STATE (-2)
#define A2M_ERR_UNSUPPORTED (-3)
#define A2M_ERR_PIPELINE (-4)      /* a kernel's bounded barrier wait expired (bug guard, never a hang) */

int a2m_version(void);
const char* a2m_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (bench.py's
 * gpu_launches claim is read from here). */
int64_t a2m_launch_count(void);
void a2m_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * log-mel front end: replaces pose_video/mel_features.py:192-223 log_mel_spectrogram()
 * (frame :21-45, periodic_hann :48-68, stft_magnitude :71-92, the np.dot with
 * spectrogram_to_mel_matrix :114-189 and the log :223) for a batch of clips in one launch.
 *
 * The plan carries the constants the reference recomputes on every call: the periodic Hann window
 * and the mel matrix, both computed by the caller on the host in fp64 with the reference formulas
 * (the Python drop-in does that) and rounded to fp32 here.
 *   nfft: a power of two from 64 to 4096 (mel_features.py:212-214 picks 2^ceil(log2(window)): 512 for the hot path's
 *   25 ms window at 16 kHz -- the tuned kernel --, 256 at the function's 8 kHz default, 1024 / 2048 at 22.05 / 44.1 kHz;
 *   anything else returns A2M_ERR_UNSUPPORTED),
 *   1 <= window <= nfft, hop >= 1, 1 <= n_mel <= 128,
 *   mel_weights: host fp64 [nfft/2+1, n_mel] row-major (each band is evaluated over the run of bins from its first to
 *   its last non-zero weight).
 * ---------------------------------------------------------------------------------------------- */
typedef struct a2m_mel_plan a2m_mel_plan;
int a2m_mel_plan_create(int window, int hop, int nfft, int n_mel, const double* hann_host,
                        const double* mel_weights_host, double log_offset, int device, a2m_mel_plan** out);
/* The same with the log variant chosen: A2M_LOG_ADD_OFFSET log(x + log_offset) (mel_features.py:223), or
 * A2M_LOG_FLOOR_ZEROS log(x == 0 ? log_offset : x), the masking of the PATS front ends
 * (pats/data_loading/audio.py:117-119 log_mel_400, :70-79 log_mel_512).  The window table may carry leading zeros
 * (librosa centres a short window inside the fft frame: window = nfft, pats/data_loading/audio.py:98-104). */
#define A2M_LOG_ADD_OFFSET 0
#define A2M_LOG_FLOOR_ZEROS 1
int a2m_mel_plan_create_ex(int window, int hop, int nfft, int n_mel, const double* hann_host,
                           const double* mel_weights_host, double log_offset, int log_mode, int device,
                           a2m_mel_plan** out);
void a2m_mel_plan_destroy(a2m_mel_plan* plan);
/* 1 + floor((n_samples - window) / hop), mel_features.py:41-42; negative when the reference would raise */
int64_t a2m_mel_num_frames(const a2m_mel_plan* plan, int64_t n_samples);
/* wav: [n_clips] rows of n_samples fp32, row stride wav_stride elements;
 * out: [n_clips, num_frames, n_mel] fp32 contiguous. */
int a2m_logmel_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                   int64_t wav_stride, float* out, void* stream);
/* The same for 16-bit PCM (the reference accepts any real dtype, mel_features.py:192-197; the samples are converted
 * exactly and multiplied by the fp32 window): half the bytes per sample on PCIe and in HBM.  wav_stride in elements. */
int a2m_logmel_i16(const a2m_mel_plan* plan, const int16_t* wav, int64_t n_clips, int64_t n_samples,
                   int64_t wav_stride, float* out, void* stream);
/* Host-only diagnostic (no GPU needed): the plan-time schedule of the 512-point kernel's mel sum for a [257, n_mel]
 * filterbank -- per step and lane the bin read (272 + b = none: a zero entry in bank b) and its weights (u into band g, v into band g - 1, both
 * carrying the factor 0.5 of the kernel's |2 X|), for segment g = 16 * round + lane.  Returns the number of steps
 * (rows of 16), or A2M_ERR_UNSUPPORTED when the matrix is not a triangular filterbank (then the general kernel runs).
 * bin_out: int[capacity_steps * 16]; uv_out: float[capacity_steps * 32]; round_steps_out: int[12]. */
int a2m_mel_schedule_host(const double* mel_weights_host, int n_mel, int capacity_steps, int* bin_out, float* uv_out,
                          int* round_steps_out, int* n_rounds_out, int* dist_last_out);
/* |STFT| only (mel_features.py:71-92 stft_magnitude): out [n_clips, num_frames, nfft/2+1] fp32 */
int a2m_stft_magnitude_f32(const a2m_mel_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                           int64_t wav_stride, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * long-FFT mel spectrogram: replaces pats/data_loading/audio.py:58-79 log_mel_512 =
 * log(floor-zeros(librosa.feature.melspectrogram(y, sr, n_fft=2048, hop_length=512))) transposed to [frames, n_mel].
 *   nfft == 2048; window_host: fp64 [nfft] (librosa pads a shorter window to nfft on both sides);
 *   power: 1 = |X|, 2 = |X|^2;  pad_mode: A2M_PAD_NONE (center=False, 1 + (n - nfft) / hop frames), A2M_PAD_REFLECT
 *   (center=True with np.pad 'reflect', librosa < 0.10 default) or A2M_PAD_ZEROS (librosa >= 0.10 default), both
 *   1 + n / hop frames;  mel_weights_host: fp64 [nfft/2+1, n_mel] row-major, every column one contiguous run of bins.
 * wav: [n_clips] rows of n_samples fp32, row stride wav_stride; out: [n_clips, frames, n_mel] fp32 contiguous.
 * ---------------------------------------------------------------------------------------------- */
#define A2M_PAD_NONE 0
#define A2M_PAD_REFLECT 1
#define A2M_PAD_ZEROS 2
typedef struct a2m_melspec_plan a2m_melspec_plan;
int a2m_melspec_plan_create(int nfft, int hop, int n_mel, int power, int pad_mode, const double* window_host,
                            const double* mel_weights_host, double log_offset, int log_mode, int device,
                            a2m_melspec_plan** out);
void a2m_melspec_plan_destroy(a2m_melspec_plan* plan);
int64_t a2m_melspec_num_frames(const a2m_melspec_plan* plan, int64_t n_samples);   /* negative: too short */
int a2m_melspec_f32(const a2m_melspec_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                    int64_t wav_stride, float* out, void* stream);


/* ------------------------------------------------------------------------------------------------
 * resampling: replaces the first step of pats/data_loading/audio.py:86-120 log_mel_400,
 * librosa.core.resample(y, orig_sr=sr, target_sr=16000) -- band-limited sinc interpolation with resampy's
 * 'kaiser_best' table (third-party, absent and unpinned in the reference: restated from its published algorithm).
 * wav: [n_clips] rows of n_samples fp32 (row stride wav_stride) -> out [n_clips, a2m_resample_out_length()] fp32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct a2m_resample_plan a2m_resample_plan;
int a2m_resample_plan_create(double orig_sr, double target_sr, int device, a2m_resample_plan** out);
void a2m_resample_plan_destroy(a2m_resample_plan* plan);
int64_t a2m_resample_out_length(const a2m_resample_plan* plan, int64_t n_samples);     /* ceil(n * target / orig) */
int a2m_resample_f32(const a2m_resample_plan* plan, const float* wav, int64_t n_clips, int64_t n_samples,
                     int64_t wav_stride, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * evaluation: replaces motion_evaluation.py:4-23 compute_pck()/compute_pck_radius() and the
 * nn.L1Loss() metric of version5_model_train.py:264,367,467 (on poses and on pos_to_motion :208-213).
 *
 * pred, gt: [n_clips, frames_per_clip, 104] fp32, each frame = 52 x then 52 y (version5_model_train.py:301).
 * Partial sums are ACCUMULATED into *accum (zero it first); they are what the ranks all-reduce.
 * PCK arithmetic is numpy's fp32 order, bit for bit: radius = fl(max(w,h) * fl32(alpha)),
 * hit = fl(sqrt(fl(fl(dx*dx) + fl(dy*dy)))) <= radius, no FMA contraction.
 * ---------------------------------------------------------------------------------------------- */
typedef struct a2m_metrics {
    int64_t pck_hits;      /* keypoints within the radius */
    int64_t n_keypoints;   /* 52 * n_frames */
    int64_t n_frames;
    int64_t n_pose;        /* elements behind abs_pose   */
    int64_t n_motion;      /* elements behind abs_motion */
    double abs_pose;       /* sum |pred - gt|                         (fp32 difference, fp64 sum) */
    double abs_motion;     /* sum |diff_t(pred) - diff_t(gt)|                                      */
    int64_t reserved;
} a2m_metrics;             /* 64 bytes */

int a2m_eval_l1_pck_f32(const float* pred, const float* gt, int64_t n_clips, int frames_per_clip, float alpha,
                        double* pck_per_frame /* nullable [n_clips*frames_per_clip] */,
                        float* radius_per_frame /* nullable [n_clips*frames_per_clip] */,
                        a2m_metrics* accum /* device */, void* stream);
/* The same for float64 poses: the reference computes in the dtype of its inputs (numpy promotion in
 * motion_evaluation.py:11-23: fp64 arrays -> fp64 bounding box, radius = side * alpha with alpha the Python float,
 * fp64 distances), so hit counts on fp64 inputs can differ from the fp32 ones near the radius.  Same partial sums. */
int a2m_eval_l1_pck_f64(const double* pred, const double* gt, int64_t n_clips, int frames_per_clip, double alpha,
                        double* pck_per_frame /* nullable */, double* radius_per_frame /* nullable */,
                        a2m_metrics* accum /* device */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * temporal smoothness / jerk metrics of the validation loop: replaces
 * version5_model_train.py:216-230 compute_temporal_smoothness_loss() and :233-248 compute_jerk_loss():
 *   accel = m[:, 1:] - m[:, :-1];  smoothness = mean over (clip, t) of ||accel||_2 over the features
 *   jerk  = accel[:, 1:] - accel[:, :-1];  jerk loss = mean of ||jerk||_2
 * seq: [n_clips, frames_per_clip, n_features] fp32 (n_features <= 128).  from_pose = 0: seq is the motion
 * (velocities) the reference functions take; from_pose = 1: seq holds poses and the motion is their first
 * difference (pos_to_motion, :208-213), taken on the fly.  Differences are single fp32 subtractions in the
 * reference's order; norms are fp32, the sums over frames fp64.  Partial sums are ADDED to *accum (zero it
 * first): mean = sum / n; n_accel = n_clips * (M - 1), n_jerk = n_clips * (M - 2), M velocities per clip.
 * ---------------------------------------------------------------------------------------------- */
typedef struct a2m_smooth_metrics {
    double sum_accel_norm;
    double sum_jerk_norm;
    int64_t n_accel;
    int64_t n_jerk;
} a2m_smooth_metrics;      /* 32 bytes */

int a2m_motion_smoothness_f32(const float* seq, int64_t n_clips, int frames_per_clip, int n_features, int from_pose,
                              a2m_smooth_metrics* accum /* device */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * pose (de)normalisation on either side of the generator: replaces the element-wise
 * steps of version5_model_train.py:300-307 / generate_motion_video.py:247-255 (view [.., 2, 52], subtract the
 * neck = joint 0 of the x block and of the y block, then (x - mean) / std), generate_motion_video.py:259-260
 * (x * std + mean) and the accumulation of normalization_tools.py:24-45 get_mean_std_necksub.
 * pose, out: [n_frames, 104] fp32; mean, std: [104] fp32 (device).  Single IEEE operations in the reference's
 * order, so results equal torch's CPU results bit for bit.
 * a2m_pose_stats_f64 ADDS to accum (device double[209]): [0,104) sum, [104,208) sum of fp32 squares of the
 * neck-subtracted poses, [208] frame count.
 * ---------------------------------------------------------------------------------------------- */
int a2m_pose_normalize_f32(const float* pose, const float* mean, const float* std, int64_t n_frames, float* out, void* stream);
int a2m_pose_denormalize_f32(const float* pose, const float* mean, const float* std, int64_t n_frames, float* out,
                             void* stream);
int a2m_pose_stats_f64(const float* pose, int64_t n_frames, double* accum, void* stream);
/* neck_sub = 0: statistics of the poses as they are (normalization_tools.py:5-20 get_mean_std); 1: as above. */
int a2m_pose_stats_ex_f64(const float* pose, int64_t n_frames, int neck_sub, double* accum, void* stream);

/* ------------------------------------------------------------------------------------------------
 * the one collective of the path: sum the 64-byte partials over the ranks.
 * NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by torch); the unique id is
 * distributed by the caller (torch.distributed store).
 * ---------------------------------------------------------------------------------------------- */
typedef struct a2m_comm a2m_comm;
int a2m_comm_unique_id(void* out128_host);
int a2m_comm_init(const void* id128_host, int rank, int world, int device, a2m_comm** out);
int a2m_allreduce_metrics(a2m_comm* comm, a2m_metrics* inout /* device */, void* stream);
int a2m_allreduce_smoothness(a2m_comm* comm, a2m_smooth_metrics* inout /* device */, void* stream);
void a2m_comm_destroy(a2m_comm* comm);

/* ------------------------------------------------------------------------------------------------
 * the dense-contraction operator of the generator: tap-offset implicit GEMM on tcgen05 tensor cores
 * (bf16 operands, fp32 TMEM accumulation).  Replaces what ConvNormRelu (model_layers.py:94-118),
 * ConvTranspose1D (:200-215), the 1x1 convs of SelfAttention (:126-128) / final_conv (:334) and
 * nn.Linear (real_motion_model.py:76,86,102,112) dispatch to cuDNN/cuBLAS in the reference:
 *   D[m, n] = act( sum_taps sum_c A_src[coords(m) + tap_off, c] * W[n, tap, c] * scale[n] + bias[n] )
 * A sources are channels-last bf16 tensors described as <= 5-D strided views (dim 0 = channels,
 * multiple of 64); rows of one 128-row M tile are the box[] extents along dims 1..4; out-of-range
 * coordinates read zeros (convolution halo, batch tail).  W is the reference's fp32 weight, addressed
 * as w[n * w_stride_n + tap_w_off + c * w_stride_c]; it is packed to bf16 on the device.
 * This entry point packs, launches and synchronises (it is the unit-test / diagnostic surface); the
 * model handle below keeps packed weights and launch plans resident.
 * ---------------------------------------------------------------------------------------------- */
#define A2M_MAX_TAPS 24
#define A2M_ACT_NONE 0
#define A2M_ACT_LEAKY 1      /* LeakyReLU(0.2) */
#define A2M_ACT_RELU 2
#define A2M_OUT_BF16 0
#define A2M_OUT_F32 1
typedef struct a2m_gemm_desc {
    int32_t n_src;                       /* 1 or 2 activation sources */
    int32_t a_rank[2];
    const void* a_ptr[2];                /* bf16, device */
    int64_t a_dims[2][5];
    int64_t a_strides[2][5];             /* elements; [0] == 1 */
    int32_t box[4];                      /* product == 128 */
    int32_t m_extent[4];                 /* output positions along dims 1..4 */
    int32_t n_taps;
    int32_t tap_src[A2M_MAX_TAPS];
    int32_t tap_off[A2M_MAX_TAPS][4];
    int32_t tap_channels[A2M_MAX_TAPS];
    int64_t tap_w_off[A2M_MAX_TAPS];
    int32_t N;
    int32_t act;
    int32_t out_type;
    int32_t tile_hint;                   /* 0: the planner's default tile; 256: 128 x 256 tiles; 512: 256 x 256 tiles on CTA pairs
                                            (cta_group::2) -- both only where the layer allows it (N % 256 == 0, bf16 output) */
    int64_t out_stride[4];               /* elements */
    int64_t out_base;
} a2m_gemm_desc;
int a2m_gemm_taps(const a2m_gemm_desc* desc, const float* w_src, int64_t w_stride_n, int64_t w_stride_c,
                  const float* scale /* nullable [N] */, const float* bias /* nullable [N] */, void* out,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * the generator: replaces real_motion_model.py:16-278 SelfAttention_G (eval-mode forward) together
 * with the model_layers.py classes it instantiates (AudioEncoder :219-280, UNet1D :283-374 with
 * decision D1, ResBlock/ConvNormRelu/SelfAttention/ChannelAttention/ConvTranspose1D) and the
 * torch_geometric GATConv/GraphConv layers it calls (real_motion_model.py:78-82,104-108).
 *
 * a2m_model_create takes the state_dict (reference key names) as device
 * tensors: fp32 parameters / buffers and the two int64 edge templates; it folds BatchNorm running
 * statistics into the conv weights, packs every matrix to bf16 in the tcgen05 K-major layout and
 * copies the small fp32 vectors, so the caller's tensors are not referenced afterwards.
 * ---------------------------------------------------------------------------------------------- */
#define A2M_DTYPE_F32 0
#define A2M_DTYPE_I64 1
typedef struct a2m_tensor_desc {
    const char* name;
    const void* data;        /* device */
    int32_t dtype;
    int32_t ndim;            /* <= 4 */
    int64_t shape[4];
} a2m_tensor_desc;
typedef struct a2m_model a2m_model;
int a2m_model_create(const a2m_tensor_desc* tensors, int n_tensors, int device, a2m_model** out);
void a2m_model_destroy(a2m_model* model);
/* mel [B, T, F] fp32, element strides (mel_stride_b, mel_stride_t, 1) so a strided slice of a longer
 * log-mel (the D2 adapter logmel[:, 0:384:6, :]) is consumed in place
 * -> pose [B, T, 104] fp32 contiguous (columns 0..19 body, 20..103 hand);
 * losses (nullable, device float[2]): [0] = 0.7*hand + 0.3*body angle penalty (:359-461),
 * [1] = bone-length MSE against real_pose (:307-347) when real_pose != NULL, else 0.
 * T multiple of 8, T <= 64; F multiple of 16 with F/8 even; 1 <= B <= 65535. */
int a2m_model_forward(a2m_model* model, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B, int T,
                      int F, float* pose, float* losses, const float* real_pose /* nullable [B, T, 104] */,
                      void* stream);
/* Sliding-window generation over long streams (BASELINE config 4; window arithmetic of pats/data_loading/dataUtils.py:
 * 585-620,648-654): clip (s, w) = rows w * (stride_window / row) + t * (stride_t / row) of stream s, read in place --
 * mel element (s, w, t, f) at mel[s * stride_stream + w * stride_window + t * stride_t + f]; one launch program for all
 * n_streams * n_windows (<= 65535) clips.  pose: [n_streams * n_windows, T, 104]. */
int a2m_model_forward_windows(a2m_model* model, const float* mel, int64_t stride_stream, int64_t stride_window,
                              int64_t stride_t, int64_t n_streams, int64_t n_windows, int T, int F, float* pose,
                              float* losses, void* stream);
/* AudioEncoder.forward (model_layers.py:267-280): mel [B, T, F] -> [B, 256, T] fp32 (reference NCW layout) */
/* Optional fused output de-normalisation (generate_motion_video.py:259-260): when
 * set, a2m_model_forward writes pose * std + mean (single fp32 multiply then add, equal to
 * a2m_pose_denormalize_f32 bit for bit) in the pass that copies the poses out of the arena; the internal
 * losses are still computed on the network's (normalised) output as in the reference.  mean, std: device
 * float[104], copied into the handle on `stream`; NULL, NULL switches it off. */
int a2m_model_set_output_denorm(a2m_model* model, const float* mean, const float* std, void* stream);
/* Diagnostic timeline (there is no nsys here): record an event after every launch-program op of the next `steps`
 * forwards of shape (B, T, F); read them back as milliseconds since a per-device reference event that all handles
 * share (so the stream lanes of a pipeline line up).  out_ms: [steps][n_ops + 1], column 0 = start of the forward;
 * ops [0, unet_end) and [body_end, n_ops) run on the caller's stream, [unet_end, body_end) on the side stream.
 * a2m_model_timeline_read synchronises the device.  Names: a2m_model_op_name. */
int a2m_model_timeline_begin(a2m_model* model, int64_t B, int T, int F, int steps);
int a2m_model_timeline_read(a2m_model* model, float* out_ms_host, int capacity, int* steps_host, int* n_ops_host,
                            int* unet_end_host, int* body_end_host, int64_t B, int T, int F);
int a2m_model_encoder_forward(a2m_model* model, const float* mel, int64_t B, int T, int F, float* out_nct, void* stream);
/* The same with the reference's optional argument: AudioEncoder.forward(x, time_steps) resizes the encoder output to
 * time_steps steps (bilinear, model_layers.py:267-279) -> [B, 256, time_steps]; time_steps <= 0 means T. */
int a2m_model_encoder_forward_ex(a2m_model* model, const float* mel, int64_t B, int T, int F, int time_steps,
                                 float* out_nct, void* stream);
/* Counts the cached launch plans (activation arenas) this handle has released: a caller that captured a CUDA graph
 * over a forward must re-capture when the number changes. */
int64_t a2m_model_plan_generation(a2m_model* model);
/* UNet1D.forward (model_layers.py:341-374, D1): [B, 256, T] fp32 -> [B, 256, T] fp32 */
int a2m_model_unet_forward(a2m_model* model, const float* x_nct, int64_t B, int T, float* out_nct, void* stream);
/* The fused five-layer graph stack of one decoder branch (real_motion_model.py:172-201 body / :224-253 hand:
 * GATConv, GraphConv, GATConv, GraphConv, GATConv, each + LayerNorm(64) -> LeakyReLU(0.2) -> + residual) on its own:
 * x, out fp32 [n_graphs, J, 64] (J = 10 body / 42 hand), converted to/from the kernel's bf16 node tiles.
 * Unit-test / diagnostic surface: allocates scratch and synchronises. part: 0 = body, 1 = hand. */
int a2m_model_gnn_forward(a2m_model* model, int part, const float* x, int64_t n_graphs, float* out, void* stream);
/* Synchronises the device and reports whether any kernel's bounded barrier wait expired. */
int a2m_model_status(a2m_model* model);
/* Algorithmic FLOPs (2*M*N*K over valid rows) of the tensor-core GEMMs one forward of this shape launches. */
int64_t a2m_model_gemm_flops(a2m_model* model, int64_t B, int T, int F);
/* Measurement aid: runs the forward `iters` times with CUDA events around every launch (synchronising)
 * and returns host floats out_ms[3] = per-forward milliseconds in {all launches, tcgen05 GEMM launches,
 * other launches}; n_gemm (nullable) = GEMM launches per forward. */
int a2m_model_profile(a2m_model* model, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B, int T,
                      int F, int iters, float* out_ms_host, int* n_gemm_host, void* stream);
/* Same, per launch: out_ms_host[min(capacity, n_ops)] = average ms of each launch of one forward, in launch
 * order; a2m_model_op_name(i) labels launch i (e.g. "unet.up1.gemm") and reports its algorithmic FLOPs
 * (0 for the non-GEMM kernels).  The returned string lives as long as the handle's plan for that shape. */
int a2m_model_profile_ops(a2m_model* model, const float* mel, int64_t mel_stride_b, int64_t mel_stride_t, int64_t B,
                          int T, int F, int iters, float* out_ms_host, int capacity, int* n_ops_host, void* stream);
const char* a2m_model_op_name(a2m_model* model, int64_t B, int T, int F, int index, int64_t* flops_host /* nullable */);

/* ------------------------------------------------------------------------------------------------
 * stand-alone building blocks: the layer classes of model_layers.py with their own forward, on the kernels the
 * generator uses (eval semantics: BatchNorm running statistics, dropout off).  The state_dict of the module is passed
 * with the prefix "blk." (e.g. "blk.conv.weight"); x, out: fp32 [B, C, T] (the reference's NCW layout).
 *   A2M_BLOCK_CONV_K3 / _K4S2    ConvNormRelu 1-D, k3 s1 p1 / k4 s2 p1 (downsample), :94-118; leaky: LeakyReLU(0.2) or ReLU
 *   A2M_BLOCK_CONV_TRANSPOSE     ConvTranspose1D k3 s2 p1 op1 + BatchNorm + ReLU, :200-215          T -> 2 T
 *   A2M_BLOCK_SELF_ATTENTION     SelfAttention, :133-146          A2M_BLOCK_CHANNEL_ATTENTION  ChannelAttention, :167-174
 *   A2M_BLOCK_RESBLOCK           ResBlock, :185-190
 * in_channels: a multiple of 64; handles are destroyed with a2m_model_destroy.
 * ---------------------------------------------------------------------------------------------- */
#define A2M_BLOCK_CONV_K3 1
#define A2M_BLOCK_CONV_K4S2 2
#define A2M_BLOCK_CONV_TRANSPOSE 3
#define A2M_BLOCK_SELF_ATTENTION 4
#define A2M_BLOCK_CHANNEL_ATTENTION 5
#define A2M_BLOCK_RESBLOCK 6
int a2m_block_create(int kind, const a2m_tensor_desc* tensors, int n_tensors, int in_channels, int out_channels, int leaky,
                     int device, a2m_model** out);
int a2m_block_forward(a2m_model* block, const float* x_nct, int64_t B, int T, float* out_nct, void* stream);

/* ------------------------------------------------------------------------------------------------
 * the discriminator: replaces real_motion_model.py:464-642 SelfAttention_D, eval-mode
 * forward(x) with audio = None and aux_labels = None (neither optional argument can work in the reference as shipped:
 * the concat with audio features has 6144 channels where `logits` takes 4096, and the auxiliary classifier is handed
 * a [B] tensor).  State_dict with the reference's key names (conv1.0.weight ... logits.bias; groups = 1,
 * in_channels 104, out_channels 64); pose [B, T, 104] fp32 -> scores [B, a2m_disc_out_length(T)] fp32.
 * Every Conv1d + BatchNorm1d + LeakyReLU is one tcgen05 implicit GEMM; the two single-layer GAT branches run in fp32.
 * Handles are destroyed with a2m_model_destroy.
 * ---------------------------------------------------------------------------------------------- */
int a2m_disc_create(const a2m_tensor_desc* tensors, int n_tensors, int n_downsampling, int device, a2m_model** out);
int a2m_disc_out_length(int T, int n_downsampling);      /* < 1: the sequence is too short */
int a2m_disc_forward(a2m_model* disc, const float* pose, int64_t B, int T, float* scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* A2M_B200_H */
