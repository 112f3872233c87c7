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

struct FrameRegs {   // this lane's share of one frame: keypoints l (a) and l+32 (b)
    float gxa, gya, pxa, pya, gxb, gyb, pxb, pyb;
};

__device__ __forceinline__ FrameRegs load_frame(const float* __restrict__ p, const float* __restrict__ g, int lane,
                                                bool has_b) {
    FrameRegs r;
    r.gxa = __ldcs(g + lane);
    r.gya = __ldcs(g + kJoints + lane);
    r.pxa = __ldcs(p + lane);
    r.pya = __ldcs(p + kJoints + lane);
    r.gxb = r.gyb = r.pxb = r.pyb = 0.f;
    if (has_b) {
        r.gxb = __ldcs(g + 32 + lane);
        r.gyb = __ldcs(g + kJoints + 32 + lane);
        r.pxb = __ldcs(p + 32 + lane);
        r.pyb = __ldcs(p + kJoints + 32 + lane);
    }
    return r;
}

__device__ __forceinline__ bool pck_hit(float gx, float gy, float px, float py, float radius) {
    const float dx = __fsub_rn(gx, px), dy = __fsub_rn(gy, py);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    return __fsqrt_rn(d2) <= radius;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
eval_l1_pck_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n_clips, int T, int seg_frames,
                   int segs, float alpha, double* __restrict__ pck_per_frame, float* __restrict__ radius_per_frame,
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
        const int t_end = min(T, t_begin + seg_frames);
        const float* p = pred + clip * T * kFeat;
        const float* g = gt + clip * T * kFeat;
        FrameRegs cur = load_frame(p + t_begin * kFeat, g + t_begin * kFeat, lane, has_b), prev = cur;
        if (t_begin > 0) prev = load_frame(p + (t_begin - 1) * kFeat, g + (t_begin - 1) * kFeat, lane, has_b);
        for (int t = t_begin; t < t_end; ++t) {
            FrameRegs nxt = cur;
            if (t + 1 < t_end) nxt = load_frame(p + (t + 1) * kFeat, g + (t + 1) * kFeat, lane, has_b);   // prefetch
            // bounding box of the ground truth (lanes without a second keypoint contribute neutral values)
            float mnx = has_b ? fminf(cur.gxa, cur.gxb) : cur.gxa, mxx = has_b ? fmaxf(cur.gxa, cur.gxb) : cur.gxa;
            float mny = has_b ? fminf(cur.gya, cur.gyb) : cur.gya, mxy = has_b ? fmaxf(cur.gya, cur.gyb) : cur.gya;
            mnx = a2m::warp_min(mnx); mxx = a2m::warp_max(mxx);
            mny = a2m::warp_min(mny); mxy = a2m::warp_max(mxy);
            const float side = fmaxf(fabsf(__fsub_rn(mxx, mnx)), fabsf(__fsub_rn(mxy, mny)));
            const float radius = __fmul_rn(side, alpha);
            const bool ha = pck_hit(cur.gxa, cur.gya, cur.pxa, cur.pya, radius);
            const bool hb = has_b && pck_hit(cur.gxb, cur.gyb, cur.pxb, cur.pyb, radius);
            const int frame_hits = __popc(__ballot_sync(0xffffffffu, ha)) + __popc(__ballot_sync(0xffffffffu, hb));
            if (lane == 0) {
                hits += frame_hits;
                const long long f = clip * T + t;
                if (pck_per_frame) pck_per_frame[f] = static_cast<double>(frame_hits) / static_cast<double>(kJoints);
                if (radius_per_frame) radius_per_frame[f] = radius;
            }
            // L1 on poses: fp32 |a-b| (torch L1Loss element op), fp64 accumulation
            float e = fabsf(__fsub_rn(cur.pxa, cur.gxa)) ;
            abs_pose += e;
            abs_pose += fabsf(__fsub_rn(cur.pya, cur.gya));
            if (has_b) {
                abs_pose += fabsf(__fsub_rn(cur.pxb, cur.gxb));
                abs_pose += fabsf(__fsub_rn(cur.pyb, cur.gyb));
            }
            // L1 on motion: first differences along time inside the clip (pos_to_motion)
            if (t > 0) {
                abs_motion += fabsf(__fsub_rn(__fsub_rn(cur.pxa, prev.pxa), __fsub_rn(cur.gxa, prev.gxa)));
                abs_motion += fabsf(__fsub_rn(__fsub_rn(cur.pya, prev.pya), __fsub_rn(cur.gya, prev.gya)));
                if (has_b) {
                    abs_motion += fabsf(__fsub_rn(__fsub_rn(cur.pxb, prev.pxb), __fsub_rn(cur.gxb, prev.gxb)));
                    abs_motion += fabsf(__fsub_rn(__fsub_rn(cur.pyb, prev.pyb), __fsub_rn(cur.gyb, prev.gyb)));
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
            const unsigned long long frames = static_cast<unsigned long long>(n_clips) * T;
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_frames), frames);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_keypoints), frames * kJoints);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_pose), frames * kFeat);
            atomicAdd(reinterpret_cast<unsigned long long*>(&accum->n_motion),
                      static_cast<unsigned long long>(n_clips) * (T > 0 ? T - 1 : 0) * kFeat);
        }
    }
}

}  // namespace

extern "C" int a2m_eval_l1_pck_f32(const float* pred, const float* gt, int64_t n_clips, int frames_per_clip, float alpha,
                                   double* pck_per_frame, float* radius_per_frame, a2m_metrics* accum, void* stream) {
    A2M_ARG_CHECK(n_clips >= 0 && frames_per_clip >= 0, "a2m_eval_l1_pck_f32: negative size");
    A2M_ARG_CHECK(accum != nullptr, "a2m_eval_l1_pck_f32: accum is NULL");
    if (n_clips == 0 || frames_per_clip == 0) return A2M_OK;
    A2M_ARG_CHECK(pred != nullptr && gt != nullptr, "a2m_eval_l1_pck_f32: NULL pose buffer");
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
    eval_l1_pck_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        pred, gt, n_clips, frames_per_clip, seg_frames, segs, alpha, pck_per_frame, radius_per_frame, accum);
    a2m_count_launch();
    A2M_LAUNCH_CHECK();
    return A2M_OK;
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
