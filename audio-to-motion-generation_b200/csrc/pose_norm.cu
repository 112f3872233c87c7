// Pose (de)normalisation around the generator: the element-wise steps the
// reference applies on either side of the hot path.
//   normalise   (version5_model_train.py:300-307, generate_motion_video.py:247-255):
//       view [.., 2, 52]; subtract the neck (joint 0 of the x block and of the y block); (x - mean) / std
//   denormalise (generate_motion_video.py:259-260):  x * std + mean
//   statistics  (normalization_tools.py:24-45 get_mean_std_necksub): per-feature sum and sum of squares of the
//       neck-subtracted poses, accumulated in fp64; the host finishes mean / std (std[0] = std[52] = 1)
// All arithmetic is single IEEE operations in the reference's order (no FMA contraction), so results equal
// torch's CPU results bit for bit.
#include "a2m_common.cuh"

void a2m_count_launch();

namespace {

constexpr int kJoints = 52, kFeat = 104;

__global__ void __launch_bounds__(256)
pose_normalize_kernel(const float* __restrict__ pose, const float* __restrict__ mean, const float* __restrict__ stdv,
                      long long n_frames, float* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n_frames * kFeat) return;
    const long long f = idx / kFeat;
    const int c = static_cast<int>(idx - f * kFeat);
    const float neck = __ldg(pose + f * kFeat + (c < kJoints ? 0 : kJoints));
    const float v = __fsub_rn(__fsub_rn(__ldg(pose + idx), neck), __ldg(mean + c));
    out[idx] = __fdiv_rn(v, __ldg(stdv + c));
}

__global__ void __launch_bounds__(256)
pose_denormalize_kernel(const float* __restrict__ pose, const float* __restrict__ mean, const float* __restrict__ stdv,
                        long long n_frames, float* __restrict__ out) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n_frames * kFeat) return;
    const int c = static_cast<int>(idx % kFeat);
    out[idx] = __fadd_rn(__fmul_rn(__ldg(pose + idx), __ldg(stdv + c)), __ldg(mean + c));
}

// accum: double[2 * 104 + 1] = sum, sum of squares, frame count (as double)
__global__ void __launch_bounds__(128)
pose_stats_kernel(const float* __restrict__ pose, long long n_frames, int neck_sub, double* __restrict__ accum) {
    const int c = threadIdx.x;                      // feature; 104 of 128 threads active
    if (c >= kFeat) return;
    double s = 0.0, q = 0.0;
    const int neck_col = c < kJoints ? 0 : kJoints;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float* row = pose + f * kFeat;
        const float v = neck_sub ? __fsub_rn(__ldg(row + c), __ldg(row + neck_col)) : __ldg(row + c);
        s += static_cast<double>(v);
        q += static_cast<double>(__fmul_rn(v, v));   // torch: mean(pose ** 2) squares in fp32
    }
    atomicAdd(accum + c, s);
    atomicAdd(accum + kFeat + c, q);
    if (c == 0 && blockIdx.x == 0) atomicAdd(accum + 2 * kFeat, static_cast<double>(n_frames));
}

}  // namespace

extern "C" int a2m_pose_normalize_f32(const float* pose, const float* mean, const float* stdv, int64_t n_frames, float* out,
                                      void* stream) {
    A2M_ARG_CHECK(n_frames >= 0, "a2m_pose_normalize_f32: negative size");
    if (n_frames == 0) return A2M_OK;
    A2M_ARG_CHECK(pose && mean && stdv && out, "a2m_pose_normalize_f32: NULL argument");
    const long long total = n_frames * kFeat;
    pose_normalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pose, mean, stdv, n_frames, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

extern "C" int a2m_pose_denormalize_f32(const float* pose, const float* mean, const float* stdv, int64_t n_frames, float* out,
                                        void* stream) {
    A2M_ARG_CHECK(n_frames >= 0, "a2m_pose_denormalize_f32: negative size");
    if (n_frames == 0) return A2M_OK;
    A2M_ARG_CHECK(pose && mean && stdv && out, "a2m_pose_denormalize_f32: NULL argument");
    const long long total = n_frames * kFeat;
    pose_denormalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pose, mean, stdv, n_frames, out);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

extern "C" int a2m_pose_stats_ex_f64(const float* pose, int64_t n_frames, int neck_sub, double* accum, void* stream) {
    A2M_ARG_CHECK(n_frames >= 0 && accum != nullptr, "a2m_pose_stats_f64: bad argument");
    if (n_frames == 0) return A2M_OK;
    A2M_ARG_CHECK(pose != nullptr, "a2m_pose_stats_f64: NULL pose");
    long long blocks = n_frames < 4LL * a2m_num_sms() ? n_frames : 4LL * a2m_num_sms();
    pose_stats_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(pose, n_frames,
                                                                                                   neck_sub ? 1 : 0, accum);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

extern "C" int a2m_pose_stats_f64(const float* pose, int64_t n_frames, double* accum, void* stream) {
    return a2m_pose_stats_ex_f64(pose, n_frames, 1, accum, stream);
}
