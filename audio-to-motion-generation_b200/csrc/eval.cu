// L1 / PCK evaluation and the one collective of the path (K13).
//
// Replaces motion_evaluation.py:4-23 and the nn.L1Loss metric of version5_model_train.py:264
// (see include/a2m_b200.h).  One warp walks the frames of one clip: lane l owns keypoints l and
// l + 32; bounding box by warp-shuffle min/max, hit count by ballot, |a-b| sums in fp64.
// PCK arithmetic reproduces numpy's fp32 operation order exactly (no FMA contraction):
//   radius = fl(max(|maxx-minx|, |maxy-miny|) * fl32(alpha));  hit = fl(sqrt(fl(dx*dx)+fl(dy*dy))) <= radius
#include <dlfcn.h>
#include <cstring>
#include "a2m_common.cuh"

void a2m_count_launch();

namespace {

constexpr int kJoints = 52;
constexpr int kFeat = 2 * kJoints;
constexpr int kWarpsPerBlock = 8;

// IEEE single operations without FMA contraction, in the input's own width (numpy computes fp32 inputs in fp32 and
// fp64 inputs in fp64, motion_evaluation.py:4-23)
template <typename T> struct Ieee;
template <> struct Ieee<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Ieee<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};
template <typename T>
__device__ __forceinline__ T warp_min_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename T>
struct FrameRegs {   // this lane's share of one frame: keypoints l (a) and l+32 (b)
    T gxa, gya, pxa, pya, gxb, gyb, pxb, pyb;
};

template <typename T>
__device__ __forceinline__ FrameRegs<T> load_frame(const T* __restrict__ p, const T* __restrict__ g, int lane, bool has_b) {
    FrameRegs<T> r;
    r.gxa = __ldcs(g + lane);
    r.gya = __ldcs(g + kJoints + lane);
    r.pxa = __ldcs(p + lane);
    r.pya = __ldcs(p + kJoints + lane);
    r.gxb = r.gyb = r.pxb = r.pyb = T(0);
    if (has_b) {
        r.gxb = __ldcs(g + 32 + lane);
        r.gyb = __ldcs(g + kJoints + 32 + lane);
        r.pxb = __ldcs(p + 32 + lane);
        r.pyb = __ldcs(p + kJoints + 32 + lane);
    }
    return r;
}

template <typename T>
__device__ __forceinline__ bool pck_hit(T gx, T gy, T px, T py, T radius) {
    const T dx = Ieee<T>::sub(gx, px), dy = Ieee<T>::sub(gy, py);
    const T d2 = Ieee<T>::add(Ieee<T>::mul(dx, dx), Ieee<T>::mul(dy, dy));
    return Ieee<T>::sqrt(d2) <= radius;
}

// T = float: the hot path (fp32 poses).  T = double: the reference called on float64 arrays; alpha arrives as the
// Python float it was, the radius and the distances are fp64, the L1 differences are fp64.
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
eval_l1_pck_kernel(const T* __restrict__ pred, const T* __restrict__ gt, long long n_clips, int T_frames, int seg_frames,
                   int segs, T alpha, double* __restrict__ pck_per_frame, T* __restrict__ radius_per_frame,
                   a2m_metrics* __restrict__ accum) {
    __shared__ double s_pose[kWarpsPerBlock], s_motion[kWarpsPerBlock];
    __shared__ unsigned long long s_hits[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool has_b = lane < kJoints - 32;
    double abs_pose = 0.0, abs_motion = 0.0;
    unsigned long long hits = 0;

    // work item = (clip, time segment): small batches still fill the machine; a segment that does not start the clip
    // loads the frame before it for the motion difference
    const long long n_items = n_clips * segs;
    for (long long item = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; item < n_items;
         item += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
        const long long clip = item / segs;
        const int t_begin = static_cast<int>(item - clip * segs) * seg_frames;
        const int t_end = min(T_frames, t_begin + seg_frames);
        const T* p = pred + clip * T_frames * kFeat;
        const T* g = gt + clip * T_frames * kFeat;
        FrameRegs<T> cur = load_frame(p + t_begin * kFeat, g + t_begin * kFeat, lane, has_b), prev = cur;
        if (t_begin > 0) prev = load_frame(p + (t_begin - 1) * kFeat, g + (t_begin - 1) * kFeat, lane, has_b);
        for (int t = t_begin; t < t_end; ++t) {
            FrameRegs<T> nxt = cur;
            if (t + 1 < t_end) nxt = load_frame(p + (t + 1) * kFeat, g + (t + 1) * kFeat, lane, has_b);   // prefetch
            // bounding box of the ground truth (lanes without a second keypoint contribute neutral values)
            T mnx = has_b ? min(cur.gxa, cur.gxb) : cur.gxa, mxx = has_b ? max(cur.gxa, cur.gxb) : cur.gxa;
            T mny = has_b ? min(cur.gya, cur.gyb) : cur.gya, mxy = has_b ? max(cur.gya, cur.gyb) : cur.gya;
            mnx = warp_min_t(mnx); mxx = warp_max_t(mxx);
            mny = warp_min_t(mny); mxy = warp_max_t(mxy);
            const T side = max(abs(Ieee<T>::sub(mxx, mnx)), abs(Ieee<T>::sub(mxy, mny)));
            const T radius = Ieee<T>::mul(side, alpha);
            const bool ha = pck_hit(cur.gxa, cur.gya, cur.pxa, cur.pya, radius);
            const bool hb = has_b && pck_hit(cur.gxb, cur.gyb, cur.pxb, cur.pyb, radius);
            const int frame_hits = __popc(__ballot_sync(0xffffffffu, ha)) + __popc(__ballot_sync(0xffffffffu, hb));
            if (lane == 0) {
                hits += frame_hits;
                const long long f = clip * T_frames + t;
                if (pck_per_frame) pck_per_frame[f] = static_cast<double>(frame_hits) / static_cast<double>(kJoints);
                if (radius_per_frame) radius_per_frame[f] = radius;
            }
            // L1 on poses: |a-b| in the input's width (torch L1Loss element op), fp64 accumulation
            abs_pose += abs(Ieee<T>::sub(cur.pxa, cur.gxa));
            abs_pose += abs(Ieee<T>::sub(cur.pya, cur.gya));
            if (has_b) {
                abs_pose += abs(Ieee<T>::sub(cur.pxb, cur.gxb));
                abs_pose += abs(Ieee<T>::sub(cur.pyb, cur.gyb));
            }
            // L1 on motion: first differences along time inside the clip (pos_to_motion)
            if (t > 0) {
                abs_motion += abs(Ieee<T>::sub(Ieee<T>::sub(cur.pxa, prev.pxa), Ieee<T>::sub(cur.gxa, prev.gxa)));
                abs_motion += abs(Ieee<T>::sub(Ieee<T>::sub(cur.pya, prev.pya), Ieee<T>::sub(cur.gya, prev.gya)));
                if (has_b) {
                    abs_motion += abs(Ieee<T>::sub(Ieee<T>::sub(cur.pxb, prev.pxb), Ieee<T>::sub(cur.gxb, prev.gxb)));
                    abs_motion += abs(Ieee<T>::sub(Ieee<T>::sub(cur.pyb, prev.pyb), Ieee<T>::sub(cur.gyb, prev.gyb)));
                }
            }
            prev = cur;
            cur = nxt;
        }
    }
    abs_pose = a2m::warp_sum(abs_pose);
    abs_motion = a2m::warp_sum(abs_motion);
    if (lane == 0) { s_pose[warp] = abs_pose; s_motion[warp] = abs_motion; s_hits[warp] = hits; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sp = 0.0, sm = 0.0;
        unsigned long long sh = 0;
        for (int w = 0; w < kWarpsPerBlock; ++w) { sp += s_pose[w]; sm += s_motion[w]; sh += s_hits[w]; }
        atomicAdd(reinterpret_cast<unsigned long long*>(&accum->pck_hits), sh);
        atomicAdd(&accum->abs_pose, sp);
        atomicAdd(&accum->abs_motion, sm);
        if (blockIdx.x == 0) {      // the counts are pure functions of the shape
            const unsigned long long frames = static_cast<unsigned long long>(n_clips) * T_frames;
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_frames), frames);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_keypoints), frames * kJoints);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_pose), frames * kFeat);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_motion),
                      static_cast<unsigned long long>(n_clips) * (T_frames > 0 ? T_frames - 1 : 0) * kFeat);
        }
    }
}

template <typename T>
int launch_eval(const T* pred, const T* gt, int64_t n_clips, int frames_per_clip, T alpha, double* pck_per_frame,
                T* radius_per_frame, a2m_metrics* accum, void* stream, const char* who) {
    A2M_ARG_CHECK(n_clips >= 0 && frames_per_clip >= 0, "%s: negative size", who);
    A2M_ARG_CHECK(accum != nullptr, "%s: accum is NULL", who);
    if (n_clips == 0 || frames_per_clip == 0) return A2M_OK;
    A2M_ARG_CHECK(pred != nullptr && gt != nullptr, "%s: NULL pose buffer", who);
    // split clips into time segments until there are ~8 warps of work per SM sub-partition (never below 4 frames)
    const long long want = 32LL * a2m_num_sms();
    int segs = static_cast<int>((want + n_clips - 1) / n_clips);
    segs = segs < 1 ? 1 : segs;
    int seg_frames = (frames_per_clip + segs - 1) / segs;
    if (seg_frames < 4) seg_frames = frames_per_clip < 4 ? frames_per_clip : 4;
    segs = (frames_per_clip + seg_frames - 1) / seg_frames;
    const long long n_items = static_cast<long long>(n_clips) * segs;
    long long blocks = (n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long cap = 32LL * a2m_num_sms();
    if (blocks > cap) blocks = cap;
    eval_l1_pck_kernel<T><<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        pred, gt, n_clips, frames_per_clip, seg_frames, segs, alpha, pck_per_frame, radius_per_frame, accum);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

}  // namespace

extern "C" int a2m_eval_l1_pck_f32(const float* pred, const float* gt, int64_t n_clips, int frames_per_clip, float alpha,
                                   double* pck_per_frame, float* radius_per_frame, a2m_metrics* accum, void* stream) {
    return launch_eval<float>(pred, gt, n_clips, frames_per_clip, alpha, pck_per_frame, radius_per_frame, accum, stream,
                              "a2m_eval_l1_pck_f32");
}

extern "C" int a2m_eval_l1_pck_f64(const double* pred, const double* gt, int64_t n_clips, int frames_per_clip, double alpha,
                                   double* pck_per_frame, double* radius_per_frame, a2m_metrics* accum, void* stream) {
    return launch_eval<double>(pred, gt, n_clips, frames_per_clip, alpha, pck_per_frame, radius_per_frame, accum, stream,
                               "a2m_eval_l1_pck_f64");
}

// ---------------------------------------------------------------------------------------------
// temporal smoothness / jerk of a motion sequence (version5_model_train.py:216-248): mean over (clip, t) of the L2 norm
// over the features of the second / third difference.  One warp per (clip, time segment); lane l owns features
// l, l + 32, l + 64, l + 96; differences are single fp32 subtractions in the reference's order, the squared norm is an
// fp32 lane sum + warp shuffle tree, the sum over frames is fp64.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kSmoothMaxPerLane = 4;       // features <= 128

struct FeatRegs { float v[kSmoothMaxPerLane]; };

__device__ __forceinline__ FeatRegs load_feats(const float* __restrict__ row, int lane, int feat) {
    FeatRegs r;
#pragma unroll
    for (int i = 0; i < kSmoothMaxPerLane; ++i) r.v[i] = lane + 32 * i < feat ? __ldg(row + lane + 32 * i) : 0.f;
    return r;
}
__device__ __forceinline__ FeatRegs feat_sub(const FeatRegs& a, const FeatRegs& b) {
    FeatRegs r;
#pragma unroll
    for (int i = 0; i < kSmoothMaxPerLane; ++i) r.v[i] = __fsub_rn(a.v[i], b.v[i]);
    return r;
}
__device__ __forceinline__ float feat_norm(const FeatRegs& a) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kSmoothMaxPerLane; ++i) s = fmaf(a.v[i], a.v[i], s);
    return sqrtf(a2m::warp_sum(s));
}

// seq [n_clips, L, feat]; from_pose: the motion is the first difference of seq (L poses -> L - 1 velocities)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
motion_smoothness_kernel(const float* __restrict__ seq, long long n_clips, int L, int feat, int from_pose, int seg_frames,
                         int segs, a2m_smooth_metrics* __restrict__ accum) {
    __shared__ double s_acc[kWarpsPerBlock], s_jerk[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = from_pose ? L - 1 : L;          // velocities per clip
    const int n_acc = M - 1, n_jerk = M - 2;      // accelerations a[t] = m[t+1] - m[t], jerks j[t] = a[t+1] - a[t]
    double sum_acc = 0.0, sum_jerk = 0.0;
    const long long n_items = n_clips * segs;
    for (long long item = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp; item < n_items;
         item += static_cast<long long>(gridDim.x) * kWarpsPerBlock) {
        const long long clip = item / segs;
        const int t0 = static_cast<int>(item - clip * segs) * seg_frames;
        const int t1 = min(n_acc, t0 + seg_frames);
        const float* base = seq + clip * L * feat;
        auto velocity = [&](int i) {               // m[i]
            if (from_pose) return feat_sub(load_feats(base + static_cast<long long>(i + 1) * feat, lane, feat),
                                           load_feats(base + static_cast<long long>(i) * feat, lane, feat));
            return load_feats(base + static_cast<long long>(i) * feat, lane, feat);
        };
        if (t0 >= t1) continue;
        FeatRegs m1 = velocity(t0 + 1);
        FeatRegs a0 = feat_sub(m1, velocity(t0));  // a[t0]
        for (int t = t0; t < t1; ++t) {
            sum_acc += feat_norm(a0);
            if (t < n_jerk) {
                const FeatRegs m2 = velocity(t + 2);
                const FeatRegs a1 = feat_sub(m2, m1);
                sum_jerk += feat_norm(feat_sub(a1, a0));
                a0 = a1;
                m1 = m2;
            }
        }
    }
    if (lane == 0) { s_acc[warp] = sum_acc; s_jerk[warp] = sum_jerk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sj = 0.0;
        for (int w = 0; w < kWarpsPerBlock; ++w) { sa += s_acc[w]; sj += s_jerk[w]; }
        atomicAdd(&accum->sum_accel_norm, sa);
        atomicAdd(&accum->sum_jerk_norm, sj);
        if (blockIdx.x == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_accel),
                      static_cast<unsigned long long>(n_clips) * (n_acc > 0 ? n_acc : 0));
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_jerk),
                      static_cast<unsigned long long>(n_clips) * (n_jerk > 0 ? n_jerk : 0));
        }
    }
}

}  // namespace

extern "C" int a2m_motion_smoothness_f32(const float* seq, int64_t n_clips, int frames_per_clip, int n_features, int from_pose,
                                         a2m_smooth_metrics* accum, void* stream) {
    A2M_ARG_CHECK(n_clips >= 0 && frames_per_clip >= 0, "a2m_motion_smoothness_f32: negative size");
    A2M_ARG_CHECK(n_features >= 1 && n_features <= 32 * kSmoothMaxPerLane, "a2m_motion_smoothness_f32: %d features (1..%d)",
                  n_features, 32 * kSmoothMaxPerLane);
    A2M_ARG_CHECK(accum != nullptr, "a2m_motion_smoothness_f32: accum is NULL");
    const int M = from_pose ? frames_per_clip - 1 : frames_per_clip;
    if (n_clips == 0 || M < 2) return A2M_OK;             // no acceleration: the reference's mean of an empty tensor is NaN
    A2M_ARG_CHECK(seq != nullptr, "a2m_motion_smoothness_f32: NULL sequence");
    const int n_acc = M - 1;
    const long long want = 32LL * a2m_num_sms();
    int segs = static_cast<int>((want + n_clips - 1) / n_clips);
    segs = segs < 1 ? 1 : segs;
    int seg_frames = (n_acc + segs - 1) / segs;
    if (seg_frames < 8) seg_frames = n_acc < 8 ? n_acc : 8;
    segs = (n_acc + seg_frames - 1) / seg_frames;
    const long long n_items = static_cast<long long>(n_clips) * segs;
    long long blocks = (n_items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long cap = 32LL * a2m_num_sms();
    if (blocks > cap) blocks = cap;
    motion_smoothness_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        seq, n_clips, frames_per_clip, n_features, from_pose ? 1 : 0, seg_frames, segs, accum);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time so the library has no link-time dependency on a particular libnccl
// ---------------------------------------------------------------------------------------------
namespace {
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*);
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int);
    int (*CommDestroy)(nccl_comm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    bool ok;
};
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0;

NcclApi* nccl_api() {
    static NcclApi api = {};
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { a2m_set_error("cannot load libnccl.so.2: %s", dlerror()); return nullptr; }
#define A2M_SYM(field, name)                                                      \
    *reinterpret_cast<void**>(&api.field) = dlsym(h, name);                       \
    if (!api.field) { a2m_set_error("libnccl: missing symbol %s", name); return nullptr; }
    A2M_SYM(GetUniqueId, "ncclGetUniqueId")
    A2M_SYM(CommInitRank, "ncclCommInitRank")
    A2M_SYM(CommDestroy, "ncclCommDestroy")
    A2M_SYM(AllReduce, "ncclAllReduce")
    A2M_SYM(GroupStart, "ncclGroupStart")
    A2M_SYM(GroupEnd, "ncclGroupEnd")
    A2M_SYM(GetErrorString, "ncclGetErrorString")
#undef A2M_SYM
    api.ok = true;
    return &api;
}
#define A2M_NCCL_CHECK(api, expr)                                                            \
    do {                                                                                     \
        int _r = (expr);                                                                     \
        if (_r != 0) { a2m_set_error("%s failed: %s", #expr, (api)->GetErrorString(_r)); return 1000 + _r; } \
    } while (0)
}  // namespace

struct a2m_comm {
    nccl_comm_t comm;
    int rank, world, device;
};

extern "C" int a2m_comm_unique_id(void* out128_host) {
    A2M_ARG_CHECK(out128_host != nullptr, "a2m_comm_unique_id: NULL");
    NcclApi* api = nccl_api();
    if (!api) return A2M_ERR_STATE;
    A2M_NCCL_CHECK(api, api->GetUniqueId(reinterpret_cast<nccl_unique_id*>(out128_host)));
    return A2M_OK;
}

extern "C" int a2m_comm_init(const void* id128_host, int rank, int world, int device, a2m_comm** out) {
    A2M_ARG_CHECK(out != nullptr && id128_host != nullptr, "a2m_comm_init: NULL");
    A2M_ARG_CHECK(world >= 1 && rank >= 0 && rank < world, "a2m_comm_init: bad rank %d / world %d", rank, world);
    *out = nullptr;
    NcclApi* api = nccl_api();
    if (!api) return A2M_ERR_STATE;
    A2M_CUDA_CHECK(cudaSetDevice(device));
    nccl_unique_id id;
    memcpy(&id, id128_host, sizeof(id));
    nccl_comm_t c = nullptr;
    A2M_NCCL_CHECK(api, api->CommInitRank(&c, world, id, rank));
    *out = new a2m_comm{c, rank, world, device};
    return A2M_OK;
}

extern "C" int a2m_allreduce_metrics(a2m_comm* comm, a2m_metrics* inout, void* stream) {
    A2M_ARG_CHECK(comm != nullptr && inout != nullptr, "a2m_allreduce_metrics: NULL");
    NcclApi* api = nccl_api();
    if (!api) return A2M_ERR_STATE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // 5 x int64 counters, then 2 x fp64 sums: integer hits make the N-GPU PCK bit-identical to 1 GPU
    A2M_NCCL_CHECK(api, api->GroupStart());
    A2M_NCCL_CHECK(api, api->AllReduce(&inout->pck_hits, &inout->pck_hits, 5, kNcclInt64, kNcclSum, comm->comm, s));
    A2M_NCCL_CHECK(api, api->AllReduce(&inout->abs_pose, &inout->abs_pose, 2, kNcclFloat64, kNcclSum, comm->comm, s));
    A2M_NCCL_CHECK(api, api->GroupEnd());
    return A2M_OK;
}

extern "C" int a2m_allreduce_smoothness(a2m_comm* comm, a2m_smooth_metrics* inout, void* stream) {
    A2M_ARG_CHECK(comm != nullptr && inout != nullptr, "a2m_allreduce_smoothness: NULL");
    NcclApi* api = nccl_api();
    if (!api) return A2M_ERR_STATE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    A2M_NCCL_CHECK(api, api->GroupStart());
    A2M_NCCL_CHECK(api, api->AllReduce(&inout->sum_accel_norm, &inout->sum_accel_norm, 2, kNcclFloat64, kNcclSum, comm->comm, s));
    A2M_NCCL_CHECK(api, api->AllReduce(&inout->n_accel, &inout->n_accel, 2, kNcclInt64, kNcclSum, comm->comm, s));
    A2M_NCCL_CHECK(api, api->GroupEnd());
    return A2M_OK;
}

extern "C" void a2m_comm_destroy(a2m_comm* comm) {
    if (!comm) return;
    NcclApi* api = nccl_api();
    if (api && comm->comm) api->CommDestroy(comm->comm);
    delete comm;
}
