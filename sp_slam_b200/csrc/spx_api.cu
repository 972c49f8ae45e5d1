// spx_api.cu -- the C ABI of include/spx.h: context, device memory layout, kernel schedule, result transfer.
//
// Replaces the two Frame member functions of the reference (/root/reference/src/Frame.cc:186 and :194, bodies
// :854-936 and :938-999) for a batch of independent frames.  Compiled for sm_100a with -fmad=false: every product
// and sum of the reference's fp32 arithmetic is rounded separately, as in the PCL path.
//
// There is no CPU fallback: without a usable CUDA device spx_create fails with SPX_ERR_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "spx_internal.h"
#include "spx_lines.cuh"
#include "spx_normals.cuh"
#include "spx_normals_strip.cuh"
#include "spx_normals_cov.cuh"
#include "spx_refine.cuh"
#include "spx_segment.cuh"

using namespace spx;

namespace {

thread_local std::string g_create_error;   // spx_last_error(NULL): the calling thread's last failed spx_create / spx_host_register

struct DevArena {
    char *base = nullptr;
    size_t size = 0, used = 0;
    template <typename T>
    T *take(size_t n) {
        const size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
};
template <typename T>
size_t padded(size_t n) { return (n * sizeof(T) + 255) & ~size_t(255); }

// ---- upload mode 3: the organized cloud's samples are picked out of the caller's image by host threads ---------------
// The sampling kernels read every Cloud.Dis-th row AND column of the depth image (src/Frame.cc:857-872): 1 / Dis^2 of its
// bytes.  The copy engine can skip rows (a pitch) but not columns, so the sparse upload still moves whole sampled rows (410 MB per
// 1000 frames at 640x480, 8 ms of PCIe).  Here a small pool of host threads copies the h x w samples of every frame into a
// page-locked staging buffer (49 MB per 1000 frames), frame group by frame group, while the device works on the groups
// that are complete; the kernels read that buffer with a column step of one (Params::samp_cstep).
struct GatherPool {
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    unsigned long long epoch = 0;       // bumped per job
    bool stop = false;
    int busy = 0;                       // workers still inside the current job
    // the job
    const char *src = nullptr; size_t pitch = 0, fstride = 0;
    float *dst = nullptr;
    int n_frames = 0, h = 0, w = 0, sw = 0, dis = 1;
    std::atomic<int> next{0};
    std::vector<int> group_of;          // frame -> group
    std::vector<int> left;              // frames of group g not yet gathered (guarded by m)

    template <int kDis>
    static void gather_rows(const char *s0, size_t pitch, float *d0, int h, int w, int sw, int dis_rt) {
        const int dis = kDis ? kDis : dis_rt;
        for (int r = 0; r < h; ++r) {
            const float *s = reinterpret_cast<const float *>(s0 + size_t(r) * size_t(dis) * pitch);
            float *d = d0 + size_t(r) * size_t(sw);
            // every sampled row is a fresh 40-line stream the hardware prefetcher has to find again: the lines of the sampled row
            // two below are requested while this one is copied (a prefetch never faults; +30 % per thread measured)
            const char *ahead = reinterpret_cast<const char *>(s) + 2 * size_t(dis) * pitch;
            int c = 0;
            for (; c + 16 <= w; c += 16) {
                const char *pf = ahead + size_t(c) * size_t(dis) * sizeof(float);
                for (int k = 0; k < 16 * dis * int(sizeof(float)); k += 64) __builtin_prefetch(pf + k, 0, 3);
                for (int k = 0; k < 16; ++k) d[c + k] = s[size_t(c + k) * size_t(dis)];
            }
            for (; c < w; ++c) d[c] = s[size_t(c) * size_t(dis)];
            for (c = w; c < sw; ++c) d[c] = 0.0f;
        }
    }
    void run_worker(unsigned long long seen) {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_work.wait(lk, [&] { return stop || epoch != seen; });
                if (stop) return;
                seen = epoch;
            }
            for (;;) {
                const int f = next.fetch_add(1, std::memory_order_relaxed);
                if (f >= n_frames) break;
                const char *s0 = src + size_t(f) * fstride;
                float *d0 = dst + size_t(f) * size_t(h) * size_t(sw);
                if (dis == 5) gather_rows<5>(s0, pitch, d0, h, w, sw, 5);
                else gather_rows<0>(s0, pitch, d0, h, w, sw, dis);
                std::lock_guard<std::mutex> lk(m);
                if (--left[size_t(group_of[size_t(f)])] == 0) cv_done.notify_all();
            }
            std::lock_guard<std::mutex> lk(m);
            if (--busy == 0) cv_done.notify_all();
        }
    }
    void start(int n_threads) {      // (between jobs only)
        unsigned long long now;
        { std::lock_guard<std::mutex> lk(m); now = epoch; }
        while (int(workers.size()) < n_threads) workers.emplace_back([this, now] { run_worker(now); });
    }
    // frames [bounds[g], bounds[g + 1]) form group g; returns at once, wait_group(g) blocks until group g is in `dst`
    void post(const char *src_, size_t pitch_, size_t fstride_, float *dst_, int h_, int w_, int sw_, int dis_, const std::vector<int> &bounds) {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return busy == 0; });     // (no worker may still be leaving the previous job)
        src = src_; pitch = pitch_; fstride = fstride_; dst = dst_; h = h_; w = w_; sw = sw_; dis = dis_;
        const int G = int(bounds.size()) - 1;
        n_frames = bounds[size_t(G)];
        group_of.assign(size_t(n_frames), 0);
        left.assign(size_t(G), 0);
        for (int g = 0; g < G; ++g) {
            left[size_t(g)] = bounds[size_t(g) + 1] - bounds[size_t(g)];
            for (int f = bounds[size_t(g)]; f < bounds[size_t(g) + 1]; ++f) group_of[size_t(f)] = g;
        }
        next.store(0);
        busy = int(workers.size());
        ++epoch;
        cv_work.notify_all();
    }
    void wait_group(int g) {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return left[size_t(g)] == 0; });
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return busy == 0; });
    }
    ~GatherPool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_work.notify_all();
        for (std::thread &t : workers) t.join();
    }
};

}  // namespace

struct spx_ctx {
    spx_config cfg;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // host path with several frame groups: all uploads go through one stream in group order and all downloads through
    // another, so a group's copies never queue behind another group's kernels (the device has few hardware queues:
    // streams beyond CUDA_DEVICE_MAX_CONNECTIONS share one and serialise)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaStream_t up2_stream = nullptr;    // the gathered groups' uploads: ready at their own pace, they do not queue behind the sampled rows
    bool up2 = true;                      // (tuning knob SPX_UP2)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // frame groups: internal streams + per group events (start, end of the plane section, end of the supposed-plane section)
    int n_streams = 1, min_group = 32, last_groups = 1;
    int refine_fast_max = 700;  // batches of at most this many frames use k_refine2
    // K3: 2 = strip kernel with TMA-staged depth chunks (production), 1 = strip kernel with plain loads (also what a depth
    // pointer / pitch that TMA cannot describe gets), 0 = the 32x16 tile kernel of round 1 (test knob SPX_NORMALS)
    int normals_mode = 2;
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeTiledFn encode_tiled = nullptr;      // cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (libcuda is not linked)
    CUtensorMap tmap;                          // the depth batch of the current call as a (columns, sampled rows, frames) tensor
    bool tmap_ok = false;
    int list_grid = 148 * 2;
    int sm_count = 148;
    bool strip_always = false;                 // test knob SPX_STRIP_ALWAYS: the strip kernel for small launches too
    int strip_occ = 4;                         // CTAs of k_normals_strip per SM the register budget is set for (tuning knob SPX_STRIP_OCC: 3 or 4)
    bool ccl_four = true;      // k_ccl_merge4 (N % 4 == 0) instead of the one-pixel-per-thread k_ccl_merge
    int ccl_frame = 1;         // k_ccl_frame (forest of a frame in shared memory, one CTA per frame): 1 = for launches of >= 512 frames
                               // (0.40 vs 0.46 ms per 1000 frames; at 1 / 64 / 128 frames the many-CTA kernels win: 0.036 / 0.056 / 0.082 vs
                               // 0.090 / 0.112 / 0.118 ms), 0 = never, 2 = whenever it fits (test knob SPX_CCL_FRAME)
    bool flatten_runs = true;  // k_ccl_flatten_runs (one pointer chase per row run) instead of k_ccl_flatten (one per pixel)
    bool refine_per_group = true;
    bool refine_dev_group = false;   // tuning knob SPX_REFINE_DEV_GROUP: the k_refine / k_refine2 choice per frame group on the resident path too
    std::vector<double> group_weights;   // tuning knob SPX_GROUP_WEIGHTS="w0,w1,...": relative group sizes on the host path
    double edge_weight = 0.5;   // host path: size of the first and the last frame group relative to the others
    int group_prio = 1;      // 1: group g's stream gets a priority that falls with g (earlier groups finish first)
    int border_grid_cap = 0;
    int upload_mode = 0;     // host input: 0 auto (pinned images: gathered samples for float batches >= gather_min_frames, else the sparse
                             // upload for batches >= sparse_min_frames), 1 whole image, 2 sparse whenever possible, 3 gathered whenever possible
    int sparse_min_frames = 1;
    int gather_min_frames = 64;   // a few frames are latency bound: one strided copy of the sampled rows is the shorter path
    size_t gather_min_bytes = size_t(200) << 20;   // ... and so is a batch whose sampled rows cross the bus in a few milliseconds: the call is
                                  // then bound by the device (120 frames 1280x720, 147 MB: 7.1 ms with the gathered route, 6.7 without)
    int gather_threads = 0;       // 0 = half the host's hardware threads, at most 16 (spx_set_gather_threads, SPX_GATHER_THREADS)
    // automatic mode: both routes at once.  The first frame groups take the sampled-rows copy (the copy engine starts at once and
    // needs no host thread), the last `gather_share` of the batch is gathered meanwhile; mode 3 gathers every group.
    double gather_share = -1.0;   // < 0: from the thread count, t / (t + 4) -- the copy engine moves a frame's sampled rows in 8.1 us, one host
                                  // thread gathers them in ~39 us (measured on the GPU box), so both routes end together near that share
                                  // (8 threads: 0.67; measured optimum 0.6 .. 0.7).  spx_set_gather_share, SPX_GATHER_SHARE
    int gather_first = 0;         // first gathered group of the current call
    size_t g_rstep = 0, g_fstride = 0;   // layout of the gathered buffer (Params::samp_* of the gathered groups)
    CUtensorMap tmap_g;           // ... and its tensor map
    bool tmap_g_ok = false;
    GatherPool *pool = nullptr;
    float *h_samp = nullptr;      // page-locked staging of the gathered samples: frames x h x samp_w floats
    size_t h_samp_cap = 0;
    float *d_samp = nullptr;      // ... and its device copy (what the sampling kernels read in upload mode 3)
    bool use_prio = false;   // back stream with higher priority: measured slower (co-running kernels slow each other), kept as a knob
    std::vector<cudaStream_t> g_streams;      // front: upload, chamfer, normals (low priority)
    std::vector<cudaStream_t> g_hstreams;     // host path: group g's stream, priority falling with g (the groups start one upload
                                              // apart; the earlier one should finish first so that its download can start)
    std::vector<cudaStream_t> g_back;         // back: everything after the normals (high priority: it is what frees a group)
    std::vector<cudaEvent_t> g_link;          // fork of the group's side stream
    std::vector<cudaEvent_t> g_side;          // side stream done (real clouds packed)
    std::vector<cudaEvent_t> g_ev;
    std::vector<float> g_host_ms;             // host clock, 2 per group: enqueue began / totals seen, ms after the call began
    std::chrono::steady_clock::time_point t_call;
    std::vector<cudaEvent_t> g_xev;           // host path: 2 per group, upload done / results on the host (spx_get_group_timeline)
    size_t work_stride = 0, work2_stride = 0;
    // host path: every group compacts its own results (at the device offset of its first frame) and ships them itself
    bool group_pack = false;
    std::vector<cudaEvent_t> g_tot_ev;
    std::vector<cudaEvent_t> g_real_ev;       // host path: the group's real-plane clouds are packed and their sizes are on the host
    struct GroupOut { int f0 = 0, f1 = 0; long long dev_pl = 0, dev_pt = 0, dev_bd = 0, dev_ix = 0; };
    std::vector<GroupOut> g_out;
    Params P;           // geometry of the last call (capacities fixed at create)
    Buffers B;
    DevArena arena;
    double *d_cov_sat = nullptr;       // COVARIANCE_MATRIX method only: 9-channel integral images, (w+1)(h+1) nodes per frame
    unsigned *d_cov_cnt = nullptr;
    float *d_curv = nullptr;           // ... and the curvature tap
    float *d_depth = nullptr;          // staging for host-side depth (tight pitch)
    uint16_t *d_depth16 = nullptr;     // staging for CV_16U host depth
    int capN = 0, cap_w = 0, cap_h = 0;
    int n_grid = 0;
    int border_grid = 148 * 8;
    int fetch_grid = 148 * 4;
    int lines_grid = 148 * 3;          // persistent CTAs of k_lines: SM count x resident CTAs per SM
    // host results (pinned, grown on demand)
    spx_frame_header *h_frames = nullptr;
    spx_plane *h_planes = nullptr;
    spx_point *h_pts = nullptr, *h_bnd = nullptr;
    size_t h_planes_cap = 0, h_pts_cap = 0, h_bnd_cap = 0;
    long long *h_totals = nullptr, *h_totals_dev = nullptr;   // page-locked; _dev = the address the device stores through
    // compact results (include/spx.h): the real planes' clouds travel as inlier index lists
    int result_mode = 0;               // what spx_extract_batch_device packs (spx_set_result_mode)
    bool compact = false;              // mode of the last run
    unsigned char *h_pidx = nullptr;   // page-locked, index_width bytes per entry
    size_t h_pidx_cap = 0;             // bytes
    spx_group_fn group_fn = nullptr;   // streaming delivery of the host-input compact calls
    void *group_user = nullptr;
    spx_compact_result view;           // what the callback sees
    std::vector<void *> graveyard;     // page-locked buffers outgrown during a call: a callback's view may still point into
                                       // them, so they are released at the start of the next call
    // state of the last extract
    bool have_run = false;
    bool debug = false;
    int last_frames = 0;
    const float *last_depth_dev = nullptr;
    int launches = 0;
    // bytes the last host-input extract moved: explicit uploads, in-place reads of the caller's pinned image by the border
    // tests (sparse upload), results copied back
    unsigned long long xfer_h2d = 0, xfer_inplace = 0, xfer_d2h = 0;
    size_t inplace_esz = 0;
    // optional per-kernel timing (spx_set_profile): event k is recorded before launch k, one more after the last
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<const char *> prof_names;
    int prof_n = 0;
    std::string err;
};

namespace {

int fail(spx_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define SPX_CK(c, call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail((c), SPX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));   \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#define SPX_DEVICE(c) DeviceGuard dev_guard_((c)->device); SPX_CK((c), dev_guard_.err)

// organized-cloud size (src/Frame.cc:873-874): ceil(cols / dis) x ceil(rows / dis), evaluated in float as there
void cloud_dims(int rows, int cols, int dis, int *w, int *h) {
    *h = int(std::ceil(rows / float(dis)));
    *w = int(std::ceil(cols / float(dis)));
}

void mt19937_seeded_state(uint32_t seed, uint32_t s[624]) {
    s[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + i;
}

void mt19937_twist(uint32_t s[624]) {
    for (int i = 0; i < 624; ++i) {
        const uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
}
uint32_t mt19937_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// the values visited by `for(float i = -0.25; i < 0.25;) { ...; i = i + 0.01; }` (src/Frame.cc:1095-1106)
int grid_values(float out[64]) {
    int n = 0;
    for (float i = -0.25; i < 0.25;) {
        if (n < 64) out[n] = i;
        ++n;
        i = i + 0.01;
    }
    return n;
}

int set_geometry(spx_ctx *c, int n_frames, int rows, int cols, size_t pitch, size_t frame_stride) {
    if (n_frames < 1 || n_frames > c->cfg.max_frames) return fail(c, SPX_ERR_ARG, "n_frames %d outside [1, %d]", n_frames, c->cfg.max_frames);
    if (rows < 1 || cols < 1 || rows > c->cfg.max_rows || cols > c->cfg.max_cols)
        return fail(c, SPX_ERR_ARG, "image %dx%d exceeds the context capacity %dx%d", rows, cols, c->cfg.max_rows, c->cfg.max_cols);
    if (pitch < size_t(cols) * sizeof(float) || pitch % sizeof(float)) return fail(c, SPX_ERR_ARG, "bad pitch");
    if (n_frames > 1 && frame_stride < pitch * size_t(rows)) return fail(c, SPX_ERR_ARG, "bad frame stride");
    Params &P = c->P;
    P.rows = rows; P.cols = cols; P.pitch = pitch; P.frame_stride = frame_stride; P.n_frames = n_frames;
    P.samp_rstep = size_t(P.dis) * pitch; P.samp_fstride = frame_stride; P.samp_cstep = P.dis; P.samp_cols = cols; P.full_alpha = 1.0f;
    cloud_dims(rows, cols, P.dis, &P.w, &P.h);
    P.N = P.w * P.h;
    // smallest non-zero |n - cx|, |m - cy| over the sampled columns / rows (the exactness bound of k_models)
    P.min_axf = P.min_ayf = 3.0e38f;
    for (int x = 0; x < P.w; ++x) { const float v = std::fabs(float(x * P.dis) - P.cx); if (v > 0.f && v < P.min_axf) P.min_axf = v; }
    for (int y = 0; y < P.h; ++y) { const float v = std::fabs(float(y * P.dis) - P.cy); if (v > 0.f && v < P.min_ayf) P.min_ayf = v; }
    if (P.N > c->capN) return fail(c, SPX_ERR_ARG, "organized cloud exceeds the context capacity");
    return SPX_OK;
}

// ---- the kernel schedule -----------------------------------------------------------------------------------
// Frames are independent, so a batch is cut into groups of consecutive frames and every group runs the whole schedule
// on its own internal stream: the per-frame serial kernels (one warp or one CTA per frame: chamfer bands, CCL rank,
// moment chains, refine, contour trace, RANSAC) of one group overlap with the bandwidth/issue bound kernels of the
// others, a group's working set stays L2 sized, and -- for host input -- group g+1's upload overlaps group g's
// kernels.  The groups join on the caller-visible stream, where the offsets of every frame's results are scanned and
// the clouds are packed into the contiguous output buffers.
struct HostSrc {           // host depth of the batch (null: the depth is already on the device)
    const void *depth = nullptr;
    size_t pitch = 0, frame_stride = 0;
    bool u16 = false;      // CV_16U source: uploaded as is and converted on the device
    float alpha = 1.0f;    // mDepthMapFactor
    // sparse upload: only the rows the organized cloud samples (every Cloud.Dis-th) are copied to the device; the border
    // tests of GeneratePlanesFromBoundries, the one consumer of full-resolution depth, read their 21x21 windows in place
    // from the caller's pinned image (`mapped` = its device-visible address)
    bool sparse = false;
    const void *mapped = nullptr;
    // upload mode 3 (implies sparse): host threads gather the organized cloud's samples, only those are uploaded
    bool gather = false;
};

int n_groups_for(const spx_ctx *c, int n_frames) {
    int g = c->n_streams;
    const int by_size = (n_frames + c->min_group - 1) / c->min_group;
    if (g > by_size) g = by_size;
    return g < 1 ? 1 : g;
}

int prof_slot(spx_ctx *c, const char *name, cudaStream_t st) {
    const int k = c->prof_n;
    while (int(c->prof_ev.size()) < 2 * (k + 1)) {
        cudaEvent_t e;
        SPX_CK(c, cudaEventCreate(&e));
        c->prof_ev.push_back(e);
    }
    if (int(c->prof_names.size()) <= k) c->prof_names.resize(k + 1);
    c->prof_names[k] = name;
    SPX_CK(c, cudaEventRecord(c->prof_ev[2 * k], st));
    ++c->prof_n;
    return SPX_OK;
}

int run_group(spx_ctx *c, const float *depth_dev, const void *depth_full, bool normals_given, int g, int f0, int ng, cudaStream_t st,
              cudaStream_t st_back, const HostSrc &src) {
    Params P = c->P;
    P.frame0 = f0; P.n_frames = ng;
    // k_refine2 (a CTA per frame) has the shorter critical path but executes ~1.8x the instructions of k_refine (a warp per
    // frame): it wins while the whole batch fits the machine in about one wave (measured cross-over ~700 frames)
    // On the host path the groups start one upload apart instead of together, so the decision is made per group there.
    const int refine_load = ((c->group_pack && c->refine_per_group) || c->refine_dev_group) ? ng : c->P.n_frames;
    P.refine_fast = (refine_load <= c->refine_fast_max && P.h <= kRefMaxH) ? 1 : 0;   // small launches are latency bound: parallelism inside the frame
    // (resident batches run their groups side by side, a throughput regime: the batch size decides there -- 3.87 against 4.03 ms per 1000 frames)
    P.forest_in_smem = (ccl_frame_fits(P.N) && (c->ccl_frame == 2 || (c->ccl_frame == 1 && (c->group_pack ? ng : c->P.n_frames) >= 512))) ? 1 : 0;
    const bool gathered = src.gather && g >= c->gather_first;
    if (gathered) {               // this group's samples come from the gathered buffer
        depth_dev = c->d_samp;
        P.samp_cstep = 1; P.samp_cols = P.w; P.samp_rstep = c->g_rstep; P.samp_fstride = c->g_fstride;
        P.fetch_skip_sampled = 0;   // the device image holds nothing of this group yet: the windows fetch every row they touch
    }
    const bool tmap_ok = gathered ? c->tmap_g_ok : c->tmap_ok;
    Buffers B = c->B;
    B.work = c->B.work + size_t(g) * c->work_stride;
    B.work2 = c->B.work2 + size_t(g) * c->work2_stride;
    B.nf_list = c->B.nf_list + size_t(g) * size_t(1 + c->cfg.max_frames);
    const int F = ng, N = P.N;
    const dim3 gpix(cdiv(N, 256), F);
    int &L = c->launches;
    // LAUNCH(kernel, grid, block, smem, args...): counts the launch and, when profiling, brackets it with events
#define LAUNCH(kern, grid, block, smem, ...)                                                        \
    do {                                                                                            \
        if (c->profile) { int rc_ = prof_slot(c, #kern, st); if (rc_ != SPX_OK) return rc_; }       \
        kern<<<grid, block, smem, st>>>(__VA_ARGS__);                                               \
        if (c->profile) SPX_CK(c, cudaEventRecord(c->prof_ev[2 * (c->prof_n - 1) + 1], st));        \
        ++L;                                                                                        \
    } while (0)

    cudaStream_t compute_st = st;
    if (src.depth && c->group_pack && c->last_groups > 1) st = (gathered && c->up2) ? c->up2_stream : c->up_stream;   // copies: the upload streams
    if (src.depth && gathered) {     // host depth of this group: its gathered samples (c->pool filled the staging buffer; run_pipeline waited)
        const size_t per = size_t(P.h) * (P.samp_rstep / sizeof(float));
        c->xfer_h2d += per * sizeof(float) * size_t(ng);
        SPX_CK(c, cudaMemcpyAsync(c->d_samp + per * size_t(f0), c->h_samp + per * size_t(f0), per * sizeof(float) * size_t(ng), cudaMemcpyHostToDevice, st));
    } else if (src.depth && src.sparse) {   // host depth of this group: the sampled rows only, at their place in the device image
        const size_t esz = src.u16 ? sizeof(uint16_t) : sizeof(float);
        const size_t tight = size_t(P.cols) * esz;
        c->xfer_h2d += tight * size_t(P.h) * size_t(ng);
        char *dst = (src.u16 ? reinterpret_cast<char *>(c->d_depth16) : reinterpret_cast<char *>(c->d_depth)) + tight * P.rows * size_t(f0);
        const char *hp = reinterpret_cast<const char *>(src.depth) + src.frame_stride * size_t(f0);
        if (ng == 1 || (P.rows % P.dis == 0 && src.frame_stride == src.pitch * size_t(P.rows))) {
            // the sampled rows of consecutive frames are equally spaced: one strided copy for the whole group
            SPX_CK(c, cudaMemcpy2DAsync(dst, tight * size_t(P.dis), hp, src.pitch * size_t(P.dis), tight, size_t(P.h) * size_t(ng), cudaMemcpyHostToDevice, st));
        } else {
            for (int f = 0; f < ng; ++f)
                SPX_CK(c, cudaMemcpy2DAsync(dst + tight * P.rows * size_t(f), tight * size_t(P.dis), hp + src.frame_stride * size_t(f),
                                            src.pitch * size_t(P.dis), tight, size_t(P.h), cudaMemcpyHostToDevice, st));
        }
        if (src.u16) {
            if (st != compute_st) { SPX_CK(c, cudaEventRecord(c->g_xev[2 * g + 0], st)); SPX_CK(c, cudaStreamWaitEvent(compute_st, c->g_xev[2 * g + 0], 0)); st = compute_st; }
            const size_t per = size_t(P.rows) * P.cols;   // cols % 4 == 0 on this path
            LAUNCH(k_convert_u16_rows, dim3(P.h * ng, cdiv(P.cols / 4, 128)), 128, 0, c->d_depth16 + per * size_t(f0), c->d_depth + per * size_t(f0),
                   P.cols, P.rows, P.h, P.dis, src.alpha);
        }
    } else if (src.depth) {   // host depth of this group -> staging buffer (tight pitch)
        const size_t esz = src.u16 ? sizeof(uint16_t) : sizeof(float);
        const size_t tight = size_t(P.cols) * esz;
        c->xfer_h2d += tight * size_t(P.rows) * size_t(ng);
        char *dst = (src.u16 ? reinterpret_cast<char *>(c->d_depth16) : reinterpret_cast<char *>(c->d_depth)) + tight * P.rows * size_t(f0);
        const char *hp = reinterpret_cast<const char *>(src.depth) + src.frame_stride * size_t(f0);
        if (src.pitch == tight && (ng == 1 || src.frame_stride == tight * P.rows)) {
            SPX_CK(c, cudaMemcpyAsync(dst, hp, tight * P.rows * size_t(ng), cudaMemcpyHostToDevice, st));
        } else {
            for (int f = 0; f < ng; ++f)
                SPX_CK(c, cudaMemcpy2DAsync(dst + tight * P.rows * size_t(f), tight, hp + src.frame_stride * size_t(f), src.pitch, tight,
                                            size_t(P.rows), cudaMemcpyHostToDevice, st));
        }
        if (src.u16) {
            const size_t n = size_t(P.rows) * P.cols * size_t(ng);   // a multiple of 4 is not guaranteed: pad handled by capacity
            const size_t n4 = (n + 3) / 4;
            if (st != compute_st) { SPX_CK(c, cudaEventRecord(c->g_xev[2 * g + 0], st)); SPX_CK(c, cudaStreamWaitEvent(compute_st, c->g_xev[2 * g + 0], 0)); st = compute_st; }
            LAUNCH(k_convert_u16, unsigned((n4 + 255) / 256), 256, 0, c->d_depth16 + size_t(P.rows) * P.cols * size_t(f0),
                   c->d_depth + size_t(P.rows) * P.cols * size_t(f0), n4, src.alpha);
        }
    }
    if (src.depth && st != compute_st) {   // (float input) hand over from the upload stream
        SPX_CK(c, cudaEventRecord(c->g_xev[2 * g + 0], st));
        SPX_CK(c, cudaStreamWaitEvent(compute_st, c->g_xev[2 * g + 0], 0));
        st = compute_st;
    } else if (src.depth && !(src.u16 && c->group_pack && c->last_groups > 1)) {
        SPX_CK(c, cudaEventRecord(c->g_xev[2 * g + 0], st));
    }
    SPX_CK(c, cudaMemsetAsync(B.ctl + f0, 0, sizeof(FrameCtl) * size_t(F), st));
    SPX_CK(c, cudaMemsetAsync(B.work, 0, 2 * sizeof(int), st));
    SPX_CK(c, cudaMemsetAsync(B.work2, 0, 2 * sizeof(int), st));
    SPX_CK(c, cudaMemsetAsync(B.nf_list, 0, sizeof(int), st));
    if (src.sparse && P.enable_supposed) {
        const size_t words = size_t(P.rows) * size_t(((P.cols + 7) / 8 + 31) / 32);
        SPX_CK(c, cudaMemsetAsync(B.fetch_bits + words * size_t(f0), 0, words * sizeof(unsigned) * size_t(F), st));
    }
    if (c->group_pack) SPX_CK(c, cudaMemsetAsync(B.out_totals + 8 * (g + 1) + 3, 0, sizeof(long long), st));
    const bool cov_method = !normals_given && c->cfg.normal_method == 1;
    if (cov_method) LAUNCH(k_backproject, gpix, 256, 0, depth_dev, P, B);      // (before the chamfer: it clears the distance tap)
    if (!normals_given) {
        const int nb = cdiv(P.h, kBandRows), nch = cdiv(P.w, 32);
        const int dbg = c->debug ? 1 : 0;
        const size_t csm = size_t(kChamferWarps) * (kBandSpan * (nch <= 7 ? 7 : (nch <= 14 ? 14 : 16))) * sizeof(unsigned);
        if (nch <= 7) LAUNCH(k_edge_chamfer<7>, cdiv(F * nb, kChamferWarps), kChamferWarps * 32, csm, depth_dev, P, B, dbg);
        else if (nch <= 14) LAUNCH(k_edge_chamfer<14>, cdiv(F * nb, kChamferWarps), kChamferWarps * 32, csm, depth_dev, P, B, dbg);
        else LAUNCH(k_edge_chamfer<16>, cdiv(F * nb, kChamferWarps), kChamferWarps * 32, csm, depth_dev, P, B, dbg);
        // a strip walks the frame's rows in sequence (22 batches at 480p): a throughput design.  A launch with fewer strips than SMs
        // (the tracking loop's single frame) is latency bound instead and takes the 32x16 tile kernel, whose 70 CTAs per frame
        // run side by side (0.05 ms against 0.12 ms for one frame).
        const bool few = F * cdiv(P.w, kStW) < c->sm_count;
        if (cov_method) {
            // COVARIANCE_MATRIX: whole-image 9-channel integral images in PCL's recurrence order, then one thread per pixel
            LAUNCH(k_cov_sat, F, kCovThreads, 0, P, B, c->d_cov_sat, c->d_cov_cnt);
            LAUNCH(k_cov_normals, gpix, 256, 0, P, B, c->d_cov_sat, c->d_cov_cnt, c->d_curv, 0);
            LAUNCH(k_plane_d, gpix, 256, 0, P, B);
            LAUNCH(k_ccl_link, dim3(cdiv(P.w, 32), cdiv(P.h, 8), F), dim3(32, 8), 0, P, B);
        } else if (c->normals_mode == 0 || (few && !c->strip_always)) {
            LAUNCH(k_normals_link, dim3(cdiv(P.w, kTW), cdiv(P.h, kTH), F), kNormThreads, kNormalsSmem, depth_dev, P, B, dbg);
        } else {
            const dim3 sgrid(cdiv(P.w, kStW), F);
            void (*k_normals_strip_fn)(const CUtensorMap, const float *, Params, Buffers, int) =
                tmap_ok ? (c->strip_occ == 3 ? k_normals_strip<true, 3> : k_normals_strip<true, 4>)
                        : (c->strip_occ == 3 ? k_normals_strip<false, 3> : k_normals_strip<false, 4>);
            LAUNCH(k_normals_strip_fn, sgrid, kSThreads, strip_smem_bytes(P.samp_cstep, tmap_ok), gathered ? c->tmap_g : c->tmap, depth_dev, P, B, dbg);
            // frames with NaN / Inf depth (queued by k_edge_chamfer; none on sensor data: the CTAs leave at once)
            LAUNCH(k_normals_link_list, std::min(c->list_grid, F * cdiv(P.w, kTW) * cdiv(P.h, kTH)), kNormThreads, kNormalsSmem, depth_dev, P, B, dbg);
        }
    } else {
        LAUNCH(k_backproject, gpix, 256, 0, depth_dev, P, B);
        LAUNCH(k_plane_d, gpix, 256, 0, P, B);
        LAUNCH(k_ccl_link, dim3(cdiv(P.w, 32), cdiv(P.h, 8), F), dim3(32, 8), 0, P, B);
    }
    if (st_back != st) {   // the rest of the group's chain runs on its high-priority stream
        SPX_CK(c, cudaEventRecord(c->g_link[g], st));
        SPX_CK(c, cudaStreamWaitEvent(st_back, c->g_link[g], 0));
        st = st_back;
    }
    if (P.forest_in_smem) {
        // merge + flatten of a frame's forest in shared memory (one CTA per frame)
        LAUNCH(k_ccl_frame, F, kCFThreads, size_t(N) * sizeof(unsigned short), P, B);
    } else {
        if (c->ccl_four && N % 4 == 0) LAUNCH(k_ccl_merge4, dim3(cdiv(N / 4, 256), F), 256, 0, P, B);   // four pixels per thread
        else LAUNCH(k_ccl_merge, gpix, 256, 0, P, B, 1);
        if (c->flatten_runs) LAUNCH(k_ccl_flatten_runs, dim3(cdiv(P.w, 32), cdiv(P.h, 8 * kFlatRows), F), dim3(32, 8), 0, P, B);
        else LAUNCH(k_ccl_flatten, gpix, 256, 0, P, B);
    }
    LAUNCH(k_ccl_rank, F, kRankThreads, 0, P, B);
    if (c->debug) LAUNCH(k_ccl_label, gpix, 256, 0, P, B);
    LAUNCH(k_moments_fit, dim3(kMomCands, F), 96, 0, P, B);
    LAUNCH(k_models, cdiv(F, 128), 128, 0, P, B);
    if (c->ccl_four && N % 4 == 0) LAUNCH(k_pid_init4, dim3(cdiv(N / 4, 256), F), 256, 0, P, B);
    else LAUNCH(k_pid_init, gpix, 256, 0, P, B);
    {
        if (P.refine_fast) {
            if (P.w <= 224) LAUNCH(k_refine2<7>, F, 7 * 32, refine2_smem_bytes(P.h, 7), P, B);
            else if (P.w <= 448) LAUNCH(k_refine2<14>, F, 14 * 32, refine2_smem_bytes(P.h, 14), P, B);
            else LAUNCH(k_refine2<16>, F, 16 * 32, refine2_smem_bytes(P.h, 16), P, B);
        }
    }
    // (a three-warp variant of this kernel -- loader / propagate / emit warps around a shared-memory ring -- was measured
    // slower, 0.94 vs 0.69 ms per 1000 frames: the per-row CTA barrier costs more than the shorter critical warp saves)
    if (P.w <= 224) LAUNCH(k_refine<7>, cdiv(F, kRefWarps), kRefWarps * 32, 0, P, B);
    else if (P.w <= 448) LAUNCH(k_refine<14>, cdiv(F, kRefWarps), kRefWarps * 32, 0, P, B);
    else LAUNCH(k_refine<16>, cdiv(F, kRefWarps), kRefWarps * 32, 0, P, B);
    LAUNCH(k_contour, F, kContourThreads, size_t(P.w + 2) * (P.h + 2), P, B);
    LAUNCH(k_postfilter, cdiv(F, 128), 128, 0, P, B);
    SPX_CK(c, cudaEventRecord(c->g_ev[3 * g + 1], st));
    // The clouds of the real planes are final now.  When this group compacts its own results (host path, or a batch
    // that is one group) their packing runs on a side stream while the line fits / border tests (latency bound, the
    // GPU is mostly idle) continue here; the supposed planes' clouds follow once k_supposed has run.
    const bool own_pack = c->group_pack || c->last_groups == 1;
    long long *tot = B.out_totals + 8 * (c->group_pack ? g + 1 : 0);
    long long base_pl = 0, base_pt = 0, base_bd = 0, base_ix = 0;
    if (own_pack) {
        if (c->group_pack) {
            spx_ctx::GroupOut &go = c->g_out[g];
            go.f0 = f0; go.f1 = f0 + ng;
            go.dev_pl = (long long)f0 * SPX_MAX_PLANES; go.dev_pt = (long long)f0 * P.pts_cap; go.dev_bd = (long long)f0 * P.bnd_cap;
            go.dev_ix = (long long)f0 * P.N;
            base_pl = go.dev_pl; base_pt = go.dev_pt; base_bd = go.dev_bd; base_ix = go.dev_ix;
        }
        long long *host_tot = c->group_pack ? c->h_totals_dev + 8 * (g + 1) : nullptr;
        // (compact results: the real planes' points are entries of the index arena, N per frame)
        LAUNCH(k_scan_frames, 1, 1024, 0, P, B, 0, base_pl, P.compact ? base_ix : base_pt, base_bd, tot, host_tot);
        const bool use_side = c->last_groups == 1;   // with several groups in flight the other groups fill the gaps already
        cudaStream_t side = use_side ? c->g_back[g] : st, keep = st;
        if (use_side) {
            SPX_CK(c, cudaEventRecord(c->g_link[g], st));
            SPX_CK(c, cudaStreamWaitEvent(side, c->g_link[g], 0));
        }
        st = side;
        LAUNCH(k_pack_points, gpix, 256, 0, P, B);
        LAUNCH(k_pack_contours, dim3(SPX_MAX_MODELS, F), 128, 0, P, B);
        if (use_side) SPX_CK(c, cudaEventRecord(c->g_side[g], side));
        st = keep;
        if (c->group_pack && c->last_groups > 1) {
            // the real planes' clouds (95 % of the result bytes) can leave for the host while the line fits still run
            SPX_CK(c, cudaEventRecord(c->g_real_ev[g], st));   // (their sizes were stored to the host by the stage-0 scan)
        }
    }
    if (P.enable_supposed) {
        const int lg = std::min(c->lines_grid, F * SPX_MAX_MODELS), bg = std::min(c->border_grid, F * SPX_MAX_MODELS * SPX_MAX_LINES);
        LAUNCH(k_lines, lg, kLineThreads, kLinesSmem, depth_dev, P, B);
        if (src.sparse) {
            // bring in the window sectors the border tests will read (counted in out_totals slot [3] of the group)
            unsigned long long *n_sectors = reinterpret_cast<unsigned long long *>(tot + 3);
            const int fg = std::min(c->fetch_grid, F * SPX_MAX_MODELS * SPX_MAX_LINES);
            if (src.u16) LAUNCH(k_border_fetch<uint16_t>, fg, 256, 0, static_cast<const uint16_t *>(src.mapped), c->d_depth, P, B, n_sectors);
            else LAUNCH(k_border_fetch<float>, fg, 256, 0, static_cast<const float *>(src.mapped), c->d_depth, P, B, n_sectors);
        }
        LAUNCH(k_border, bg, kBorderWarps * 32, 0, static_cast<const float *>(depth_full), P, B);
        LAUNCH(k_supposed, cdiv(F, 128), 128, 0, P, B);
    }
    if (own_pack) {
        LAUNCH(k_scan_frames, 1, 1024, 0, P, B, 1, base_pl, base_pt, base_bd, tot, c->group_pack ? c->h_totals_dev + 8 * (g + 1) : nullptr);
        LAUNCH(k_emit_records, F, 128, 0, P, B);
        if (c->last_groups == 1) SPX_CK(c, cudaStreamWaitEvent(st, c->g_side[g], 0));   // the fallback boundaries read the packed real clouds
        if (P.enable_supposed) LAUNCH(k_pack_supposed, dim3(SPX_MAX_PLANES, F), 128, 0, P, B);
        if (c->group_pack) SPX_CK(c, cudaEventRecord(c->g_tot_ev[g], st));   // (the totals are on the host already: stage-1 scan)
    }
    SPX_CK(c, cudaEventRecord(c->g_ev[3 * g + 2], st));
    return SPX_OK;
}

// depth_dev: what the sampling kernels read (layout P.samp_*); depth_full: the full-resolution image of the border tests
// (layout P.pitch / P.frame_stride; device memory, or the caller's mapped host image on the sparse-upload path)
int run_pipeline(spx_ctx *c, const float *depth_dev, const void *depth_full, bool normals_given, const HostSrc &src, bool group_pack,
                 bool compact = false) {
    cudaStream_t main_st = c->stream;
    c->group_pack = group_pack;
    c->compact = compact;
    c->P.compact = compact ? 1 : 0;
    c->P.idx16 = c->P.N <= 65536 ? 1 : 0;
    for (void *p : c->graveyard) cudaFreeHost(p);
    c->graveyard.clear();
    // the depth batch as a 3-d tensor for the strip kernel's TMA loads: (image columns, SAMPLED rows, frames) -- the row stride
    // is Cloud.Dis image rows, so only the rows the organized cloud samples are ever touched
    c->tmap_ok = false;
    if (!normals_given && c->normals_mode == 2 && c->encode_tiled && strip_tma_ok(c->P.samp_cstep) && reinterpret_cast<uintptr_t>(depth_dev) % 16 == 0 &&
        c->P.samp_rstep % 16 == 0 && (c->P.n_frames == 1 || c->P.samp_fstride % 16 == 0)) {
        const Params &Q = c->P;
        const cuuint64_t gdim[3] = {cuuint64_t(Q.samp_cols), cuuint64_t(Q.h), cuuint64_t(Q.n_frames)};
        const cuuint64_t gstr[2] = {cuuint64_t(Q.samp_rstep), cuuint64_t(Q.n_frames == 1 ? Q.samp_rstep * size_t(Q.h) : Q.samp_fstride)};
        const cuuint32_t box[3] = {cuuint32_t(strip_box_w(Q.samp_cstep)), cuuint32_t(kSB), 1u};
        const cuuint32_t es[3] = {1u, 1u, 1u};
        const CUresult r = c->encode_tiled(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(depth_dev), gdim, gstr, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->tmap_ok = r == CUDA_SUCCESS;
    }
    c->tmap_g_ok = false;
    if (src.gather && c->normals_mode == 2 && c->encode_tiled) {      // the gathered buffer: rows of samp_w floats, a column step of one
        const Params &Q = c->P;
        const cuuint64_t gdim[3] = {cuuint64_t(Q.w), cuuint64_t(Q.h), cuuint64_t(Q.n_frames)};
        const cuuint64_t gstr[2] = {cuuint64_t(c->g_rstep), cuuint64_t(Q.n_frames == 1 ? c->g_rstep * size_t(Q.h) : c->g_fstride)};
        const cuuint32_t box[3] = {cuuint32_t(strip_box_w(1)), cuuint32_t(kSB), 1u};
        const cuuint32_t es[3] = {1u, 1u, 1u};
        const CUresult r = c->encode_tiled(&c->tmap_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->d_samp, gdim, gstr, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        c->tmap_g_ok = r == CUDA_SUCCESS;
    }
    const int F = c->P.n_frames;
    c->P.frame0 = 0;
    c->launches = 0;
    c->prof_n = 0;
    c->xfer_h2d = c->xfer_inplace = c->xfer_d2h = 0;
    c->inplace_esz = src.sparse ? (src.u16 ? sizeof(uint16_t) : sizeof(float)) : 0;
    const int G = normals_given ? 1 : n_groups_for(c, F);
    c->last_groups = G;
    SPX_CK(c, cudaEventRecord(c->ev[0], main_st));
    c->t_call = std::chrono::steady_clock::now();
    c->g_host_ms.assign(size_t(2 * G), 0.f);
    if (group_pack && G > 1) { SPX_CK(c, cudaStreamWaitEvent(c->up_stream, c->ev[0], 0)); SPX_CK(c, cudaStreamWaitEvent(c->up2_stream, c->ev[0], 0)); }
    // Host path: the first group is what the device waits for and the last is what the host waits for, so both are
    // smaller than the ones in between (weights edge_w : 1 ... 1 : edge_w).
    std::vector<int> bounds(size_t(G) + 1, 0);
    {
        std::vector<double> wts(size_t(G), 1.0);
        if (int(c->group_weights.size()) == G) wts = c->group_weights;
        else if (group_pack && G >= 4) { wts[0] = c->edge_weight; wts[size_t(G) - 1] = c->edge_weight; }
        double tw = 0.0, acc = 0.0;
        for (double v : wts) tw += v;
        for (int g = 0; g < G; ++g) {
            acc += wts[size_t(g)];
            bounds[size_t(g) + 1] = G == 1 ? F : std::min(F, std::max(bounds[size_t(g)] + 1, int(F * acc / tw + 0.5)));
        }
        bounds[size_t(G)] = F;
    }
    c->gather_first = 0;
    if (src.gather) {
        // mode 3: every group is gathered.  Automatic mode: the first groups take the sampled-rows copy, the tail of the batch
        // (about gather_share of the frames) is gathered by the host threads meanwhile.
        if (c->upload_mode != 3) {
            const int nt = int(c->pool->workers.size());
            const double share = c->gather_share >= 0.0 ? c->gather_share : double(nt) / double(nt + 4);
            const int f_split = int(double(F) * (1.0 - share) + 0.5);
            while (c->gather_first < G && bounds[size_t(c->gather_first)] < f_split) ++c->gather_first;
        }
        if (c->gather_first < G) {
            std::vector<int> gb(bounds.begin() + c->gather_first, bounds.end());      // frames [gb[0], F): groups gather_first .. G - 1
            const int fa = gb[0];
            for (int &v : gb) v -= fa;
            const size_t per = size_t(c->P.h) * (c->g_rstep / sizeof(float));
            c->pool->post(static_cast<const char *>(src.depth) + src.frame_stride * size_t(fa), src.pitch, src.frame_stride, c->h_samp + per * size_t(fa),
                          c->P.h, c->P.w, int(c->g_rstep / sizeof(float)), c->P.dis, gb);
        }
    }
    for (int g = 0; g < G; ++g) {
        const int f0 = bounds[size_t(g)], f1 = bounds[size_t(g) + 1];
        if (src.gather && g >= c->gather_first) c->pool->wait_group(g - c->gather_first);   // (the workers carry on with the later groups)
        cudaStream_t st = (G == 1) ? main_st : (group_pack ? c->g_hstreams[g] : c->g_streams[g]);
        cudaStream_t st_back = (G == 1) ? main_st : (c->use_prio ? c->g_back[g] : st);
        if (G > 1) SPX_CK(c, cudaStreamWaitEvent(st, c->ev[0], 0));
        SPX_CK(c, cudaEventRecord(c->g_ev[3 * g + 0], st));
        c->g_host_ms[2 * g] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - c->t_call).count();
        int rc = run_group(c, depth_dev, depth_full, normals_given, g, f0, f1 - f0, st, st_back, src);
        if (rc != SPX_OK) { if (src.gather) c->pool->wait_idle(); return rc; }   // (the job reads the caller's image)
        if (G > 1 && group_pack) SPX_CK(c, cudaStreamWaitEvent(main_st, c->g_ev[3 * g + 2], 0));
    }
    if (!group_pack && G > 1) {
        // several groups, one contiguous device result: the real planes' clouds are packed on a side stream as soon as
        // every group has passed its post-filter; the rest follows when the groups are done
        const Params &P = c->P;
        const Buffers &B = c->B;
        const dim3 gpix(cdiv(P.N, 256), F);
        int &L = c->launches;
        cudaStream_t side = c->g_back[0];
        cudaStream_t st = side;
        for (int g = 0; g < G; ++g) SPX_CK(c, cudaStreamWaitEvent(side, c->g_ev[3 * g + 1], 0));
        LAUNCH(k_scan_frames, 1, 1024, 0, P, B, 0, 0ll, 0ll, 0ll, B.out_totals, static_cast<long long *>(nullptr));
        LAUNCH(k_pack_points, gpix, 256, 0, P, B);
        LAUNCH(k_pack_contours, dim3(SPX_MAX_MODELS, F), 128, 0, P, B);
        SPX_CK(c, cudaEventRecord(c->g_side[0], side));
        st = main_st;
        for (int g = 0; g < G; ++g) SPX_CK(c, cudaStreamWaitEvent(main_st, c->g_ev[3 * g + 2], 0));
        SPX_CK(c, cudaStreamWaitEvent(main_st, c->g_side[0], 0));
        SPX_CK(c, cudaEventRecord(c->ev[1], main_st));
        LAUNCH(k_scan_frames, 1, 1024, 0, P, B, 1, 0ll, 0ll, 0ll, B.out_totals, static_cast<long long *>(nullptr));
        LAUNCH(k_emit_records, F, 128, 0, P, B);
        if (P.enable_supposed) LAUNCH(k_pack_supposed, dim3(SPX_MAX_PLANES, F), 128, 0, P, B);
    } else {
        SPX_CK(c, cudaEventRecord(c->ev[1], main_st));
    }
#undef LAUNCH
    SPX_CK(c, cudaEventRecord(c->ev[2], main_st));
    SPX_CK(c, cudaGetLastError());
    c->have_run = true;
    c->last_frames = F;
    c->last_depth_dev = depth_dev;
    return SPX_OK;
}

// Host input: may only the sampled rows be uploaded?  Yes when nothing reads full-resolution depth (supposed planes off),
// or when the caller's image is page-locked (cudaHostAlloc / cudaHostRegister / spx_host_register): then the border tests
// read their windows in place through the image's device-visible address.  A pageable image is uploaded whole.
bool sparse_upload(spx_ctx *c, const void *host, size_t extent_bytes, int n_frames, HostSrc *src) {
    src->sparse = false; src->mapped = nullptr;
    if (c->upload_mode == 1) return false;
    if (c->upload_mode == 0 && n_frames < c->sparse_min_frames) return false;
    if (!c->P.enable_supposed) { src->sparse = true; return true; }
    cudaPointerAttributes at, at_end;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    // the whole batch must be page-locked, not just its first byte (k_border_fetch reads anywhere inside it)
    const char *last = static_cast<const char *>(host) + (extent_bytes ? extent_bytes - 1 : 0);
    if (cudaPointerGetAttributes(&at_end, last) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at_end.type != cudaMemoryTypeHost || !at_end.devicePointer ||
        static_cast<const char *>(at_end.devicePointer) - static_cast<const char *>(at.devicePointer) != last - static_cast<const char *>(host)) return false;
    src->sparse = true; src->mapped = at.devicePointer;
    return true;
}

template <typename T>
int grow_pinned(spx_ctx *c, T **p, size_t *cap, size_t need) {
    if (need <= *cap) return SPX_OK;
    size_t ncap = *cap ? *cap : 1024;
    while (ncap < need) ncap *= 2;
    if (*p) SPX_CK(c, cudaFreeHost(*p));
    *p = nullptr; *cap = 0;
    SPX_CK(c, cudaHostAlloc(reinterpret_cast<void **>(p), ncap * sizeof(T), cudaHostAllocDefault));
    *cap = ncap;
    return SPX_OK;
}

void fill_view(const spx_ctx *c, spx_compact_result *v, int n_frames, long long n_pl, long long n_ix, long long n_pt, long long n_bd) {
    v->n_frames = n_frames;
    v->n_planes_total = int(n_pl);
    v->index_width = c->compact ? (c->P.idx16 ? 2 : 4) : 0;     // 0: every plane's cloud is in `points` (the 16-byte-cloud calls)
    v->cloud_width = c->P.w; v->cloud_height = c->P.h; v->cloud_dis = c->P.dis;
    v->n_index_total = n_ix; v->n_points_total = n_pt; v->n_boundary_total = n_bd;
    v->frames = c->h_frames; v->planes = c->h_planes;
    v->point_index = c->compact ? c->h_pidx : nullptr; v->points = c->h_pts; v->boundary = c->h_bnd;
}

// results of spx_extract_batch_device -> host.  cout: the compact layout (the run must have packed it), else `out`.
int fetch(spx_ctx *c, spx_batch_result *out, bool with_clouds, spx_compact_result *cout = nullptr) {
    if (!c->have_run) return fail(c, SPX_ERR_STATE, "no extract call has been made on this context");
    if (c->group_pack) return fail(c, SPX_ERR_STATE, "the last extract already delivered its results to the host; fetch follows spx_extract_batch_device");
    if (c->compact != (cout != nullptr))
        return fail(c, SPX_ERR_STATE, c->compact ? "the last extract packed compact results: use spx_fetch_compact" : "the last extract packed point clouds: use spx_fetch_results (spx_set_result_mode selects what is packed)");
    cudaStream_t st = c->stream;
    const int F = c->last_frames;
    SPX_CK(c, cudaMemcpyAsync(c->h_totals, c->B.out_totals, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    SPX_CK(c, cudaMemcpyAsync(c->h_frames, c->B.out_frames, sizeof(spx_frame_header) * size_t(F), cudaMemcpyDeviceToHost, st));
    SPX_CK(c, cudaStreamSynchronize(st));
    const long long n_pl = c->h_totals[0], n_pt = c->h_totals[1], n_bd = c->h_totals[2], n_ix = cout ? c->h_totals[5] : 0;
    const size_t iw = c->P.idx16 ? 2 : 4;
    int rc;
    if ((rc = grow_pinned(c, &c->h_planes, &c->h_planes_cap, size_t(n_pl))) != SPX_OK) return rc;
    if (n_pl) SPX_CK(c, cudaMemcpyAsync(c->h_planes, c->B.out_planes, sizeof(spx_plane) * size_t(n_pl), cudaMemcpyDeviceToHost, st));
    if (with_clouds) {
        if ((rc = grow_pinned(c, &c->h_pts, &c->h_pts_cap, size_t(n_pt))) != SPX_OK) return rc;
        if ((rc = grow_pinned(c, &c->h_bnd, &c->h_bnd_cap, size_t(n_bd))) != SPX_OK) return rc;
        if ((rc = grow_pinned(c, &c->h_pidx, &c->h_pidx_cap, size_t(n_ix) * iw)) != SPX_OK) return rc;
        if (n_pt) SPX_CK(c, cudaMemcpyAsync(c->h_pts, c->B.out_pts, sizeof(spx_point) * size_t(n_pt), cudaMemcpyDeviceToHost, st));
        if (n_bd) SPX_CK(c, cudaMemcpyAsync(c->h_bnd, c->B.out_bnd, sizeof(spx_point) * size_t(n_bd), cudaMemcpyDeviceToHost, st));
        if (n_ix) SPX_CK(c, cudaMemcpyAsync(c->h_pidx, c->B.out_pidx, size_t(n_ix) * iw, cudaMemcpyDeviceToHost, st));
    }
    SPX_CK(c, cudaStreamSynchronize(st));
    if (cout) {
        fill_view(c, cout, F, n_pl, n_ix, n_pt, n_bd);
    } else if (out) {
        out->n_frames = F;
        out->n_planes_total = int(n_pl);
        out->n_points_total = with_clouds ? n_pt : 0;
        out->n_boundary_total = with_clouds ? n_bd : 0;
        out->frames = c->h_frames;
        out->planes = c->h_planes;
        out->points = with_clouds ? c->h_pts : nullptr;
        out->boundary = with_clouds ? c->h_bnd : nullptr;
    }
    return SPX_OK;
}

// pinned buffer growth that keeps the first `used` elements (other groups' copies may already have landed there).  The old
// buffer is not released before the next extract call: a group callback's view may still point into it.
template <typename T>
int grow_pinned_keep(spx_ctx *c, T **p, size_t *cap, size_t need, size_t used) {
    if (need <= *cap) return SPX_OK;
    SPX_CK(c, cudaStreamSynchronize(c->last_groups == 1 ? c->stream : c->down_stream));
    size_t ncap = *cap ? *cap : 1024;
    while (ncap < need) ncap *= 2;
    T *np = nullptr;
    SPX_CK(c, cudaHostAlloc(reinterpret_cast<void **>(&np), ncap * sizeof(T), cudaHostAllocDefault));
    if (*p && used) std::memcpy(np, *p, used * sizeof(T));
    if (*p) c->graveyard.push_back(*p);
    *p = np; *cap = ncap;
    return SPX_OK;
}

// host path: as soon as a group's totals are known its frame headers, plane records and clouds are copied to where
// they belong in the contiguous host arrays (on the download stream, overlapping the other groups' work).  When a
// group's copies have landed, the offsets inside its records are rebased from the device layout to the host layout and
// (compact calls) the group is handed to the caller's callback while the later groups are still on the device.
int fetch_groups(spx_ctx *c, spx_batch_result *out, spx_compact_result *cout) {
    const int F = c->last_frames, G = c->last_groups;
    const bool cp = c->compact;
    const size_t iw = c->P.idx16 ? 2 : 4;
    long long run_pl = 0, run_pt = 0, run_bd = 0, run_ix = 0;
    std::vector<long long> host_pl(G), host_pt(G), host_bd(G), host_ix(G), n_pls(G), end_ix(G), end_pt(G), end_bd(G);
    int rc;
    auto finalize = [&](int g) -> int {
        if (G > 1) SPX_CK(c, cudaEventSynchronize(c->g_xev[2 * g + 1])); else SPX_CK(c, cudaStreamSynchronize(c->stream));
        const spx_ctx::GroupOut &go = c->g_out[g];
        const long long d_pl = host_pl[g] - go.dev_pl, d_pt = host_pt[g] - go.dev_pt, d_bd = host_bd[g] - go.dev_bd, d_ix = host_ix[g] - go.dev_ix;
        for (int f = go.f0; f < go.f1; ++f) c->h_frames[f].first_plane += int(d_pl);
        for (long long k = 0; k < n_pls[g]; ++k) {
            spx_plane &pl = c->h_planes[host_pl[g] + k];
            pl.points_off += (cp && !pl.is_supposed) ? d_ix : d_pt;
            pl.boundary_off += d_bd;
        }
        if (c->group_fn) {
            fill_view(c, &c->view, go.f1, host_pl[g] + n_pls[g], end_ix[g], end_pt[g], end_bd[g]);
            c->group_fn(c->group_user, go.f0, go.f1, &c->view);
        }
        return SPX_OK;
    };
    for (int g = 0; g < G; ++g) {
        cudaStream_t st = (G == 1) ? c->stream : c->down_stream;   // the group's kernels are done once its totals have arrived
        const spx_ctx::GroupOut &go = c->g_out[g];
        long long early = 0;   // entries of the real planes' clouds (points, or indices) already on their way
        if (G > 1) {   // first the clouds of the real planes, while the group's line fits / border tests still run
            SPX_CK(c, cudaEventSynchronize(c->g_real_ev[g]));
            const long long *t0 = c->h_totals + 8 * (g + 1);
            early = t0[5];   // (the boundary arena waits: k_pack_supposed may still write a real plane's every-20th-inlier fallback boundary)
            if (cp) {
                if ((rc = grow_pinned_keep(c, &c->h_pidx, &c->h_pidx_cap, size_t(run_ix + early) * iw, size_t(run_ix) * iw)) != SPX_OK) return rc;
                if (early) SPX_CK(c, cudaMemcpyAsync(c->h_pidx + size_t(run_ix) * iw, static_cast<const unsigned char *>(c->B.out_pidx) + size_t(go.dev_ix) * iw,
                                                     size_t(early) * iw, cudaMemcpyDeviceToHost, st));
            } else {
                if ((rc = grow_pinned_keep(c, &c->h_pts, &c->h_pts_cap, size_t(run_pt + early), size_t(run_pt))) != SPX_OK) return rc;
                if (early) SPX_CK(c, cudaMemcpyAsync(c->h_pts + run_pt, c->B.out_pts + go.dev_pt, sizeof(spx_point) * size_t(early), cudaMemcpyDeviceToHost, st));
            }
        }
        SPX_CK(c, cudaEventSynchronize(c->g_tot_ev[g]));
        c->g_host_ms[2 * g + 1] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - c->t_call).count();
        const long long *t = c->h_totals + 8 * (g + 1);
        const long long n_pl = t[0], n_pt = t[1], n_bd = t[2], n_ix = cp ? t[5] : 0;
        c->xfer_inplace += (unsigned long long)t[3] * 8ull * c->inplace_esz;   // sectors of 8 pixels
        c->xfer_d2h += sizeof(spx_frame_header) * size_t(go.f1 - go.f0) + sizeof(spx_plane) * size_t(n_pl) + sizeof(spx_point) * size_t(n_pt + n_bd) +
                       iw * size_t(n_ix) + 4 * sizeof(long long);
        const long long early_pt = cp ? 0 : early, early_ix = cp ? early : 0;
        if ((rc = grow_pinned_keep(c, &c->h_planes, &c->h_planes_cap, size_t(run_pl + n_pl), size_t(run_pl))) != SPX_OK) return rc;
        if ((rc = grow_pinned_keep(c, &c->h_pts, &c->h_pts_cap, size_t(run_pt + n_pt), size_t(run_pt + early_pt))) != SPX_OK) return rc;
        if ((rc = grow_pinned_keep(c, &c->h_bnd, &c->h_bnd_cap, size_t(run_bd + n_bd), size_t(run_bd))) != SPX_OK) return rc;
        if ((rc = grow_pinned_keep(c, &c->h_pidx, &c->h_pidx_cap, size_t(run_ix + n_ix) * iw, size_t(run_ix + early_ix) * iw)) != SPX_OK) return rc;
        SPX_CK(c, cudaMemcpyAsync(c->h_frames + go.f0, c->B.out_frames + go.f0, sizeof(spx_frame_header) * size_t(go.f1 - go.f0), cudaMemcpyDeviceToHost, st));
        if (n_pl) SPX_CK(c, cudaMemcpyAsync(c->h_planes + run_pl, c->B.out_planes + go.dev_pl, sizeof(spx_plane) * size_t(n_pl), cudaMemcpyDeviceToHost, st));
        if (n_ix > early_ix) SPX_CK(c, cudaMemcpyAsync(c->h_pidx + size_t(run_ix + early_ix) * iw, static_cast<const unsigned char *>(c->B.out_pidx) + size_t(go.dev_ix + early_ix) * iw,
                                                       size_t(n_ix - early_ix) * iw, cudaMemcpyDeviceToHost, st));
        if (n_pt > early_pt) SPX_CK(c, cudaMemcpyAsync(c->h_pts + run_pt + early_pt, c->B.out_pts + go.dev_pt + early_pt, sizeof(spx_point) * size_t(n_pt - early_pt), cudaMemcpyDeviceToHost, st));
        if (n_bd) SPX_CK(c, cudaMemcpyAsync(c->h_bnd + run_bd, c->B.out_bnd + go.dev_bd, sizeof(spx_point) * size_t(n_bd), cudaMemcpyDeviceToHost, st));
        SPX_CK(c, cudaEventRecord(c->g_xev[2 * g + 1], st));
        host_pl[g] = run_pl; host_pt[g] = run_pt; host_bd[g] = run_bd; host_ix[g] = run_ix; n_pls[g] = n_pl;
        run_pl += n_pl; run_pt += n_pt; run_bd += n_bd; run_ix += n_ix;
        end_ix[g] = run_ix; end_pt[g] = run_pt; end_bd[g] = run_bd;
        if (g > 0 && (rc = finalize(g - 1)) != SPX_OK) return rc;   // (its copies were enqueued one group ago: landed by now)
    }
    if ((rc = finalize(G - 1)) != SPX_OK) return rc;
    if (G > 1) SPX_CK(c, cudaStreamSynchronize(c->down_stream));
    SPX_CK(c, cudaStreamSynchronize(c->stream));
    if (cout) {
        fill_view(c, cout, F, run_pl, run_ix, run_pt, run_bd);
    } else {
        out->n_frames = F;
        out->n_planes_total = int(run_pl);
        out->n_points_total = run_pt;
        out->n_boundary_total = run_bd;
        out->frames = c->h_frames;
        out->planes = c->h_planes;
        out->points = c->h_pts;
        out->boundary = c->h_bnd;
    }
    return SPX_OK;
}

int check_frame(spx_ctx *c, int frame) {
    if (!c) return SPX_ERR_ARG;
    if (!c->have_run) return fail(c, SPX_ERR_STATE, "no extract call has been made on this context");
    if (!c->debug) return fail(c, SPX_ERR_STATE, "debug taps are off: call spx_set_debug(ctx, 1) before the extract");
    if (frame < 0 || frame >= c->last_frames) return fail(c, SPX_ERR_ARG, "frame %d outside the last batch", frame);
    SPX_DEVICE(c);
    SPX_CK(c, cudaStreamSynchronize(c->stream));
    return SPX_OK;
}

template <typename T>
int d2h(spx_ctx *c, T *dst, const T *src, size_t n) {
    if (!dst) return SPX_OK;
    SPX_CK(c, cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost));
    return SPX_OK;
}

int get_ctl(spx_ctx *c, int frame, FrameCtl *out) {
    SPX_CK(c, cudaMemcpy(out, c->B.ctl + frame, sizeof(FrameCtl), cudaMemcpyDeviceToHost));
    return SPX_OK;
}

}  // namespace

int spx_internal_fail(spx_ctx *c, int code, const char *what, const char *msg) { return fail(c, code, "%s: %s", what, msg); }
void *spx_internal_stream(spx_ctx *c) { return c->stream; }
int spx_internal_device(spx_ctx *c) { return c->device; }
int spx_internal_last_results(spx_ctx *c, const spx_frame_header **frames, const spx_plane **planes, const spx_point **boundary) {
    if (!c->have_run) return 0;
    *frames = c->B.out_frames; *planes = c->B.out_planes; *boundary = c->B.out_bnd;
    return c->last_frames;
}

extern "C" {

void spx_default_config(spx_config *cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->cloud_dis = 3; cfg->min_size = 500; cfg->angle_thr_deg = 3.0f; cfg->dist_thr = 0.05f;
    cfg->line_ratio = 0.2; cfg->line_dist_thr = 0.01f;
    cfg->fx = 517.306408f; cfg->fy = 516.469215f; cfg->cx = 318.643040f; cfg->cy = 255.313989f;
    cfg->min_x = 0.0f; cfg->max_x = 640.0f; cfg->min_y = 0.0f; cfg->max_y = 480.0f;
    cfg->max_depth_change_factor = 0.05f; cfg->normal_smoothing_size = 10.0f;
    cfg->ransac_max_iter = 1000; cfg->enable_supposed = 1;
    cfg->max_frames = 1; cfg->max_rows = 480; cfg->max_cols = 640; cfg->device = 0;
}

int spx_create(const spx_config *cfg, spx_ctx **out) {
    if (!cfg || !out) return fail(nullptr, SPX_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->cloud_dis < 1 || cfg->max_frames < 1 || cfg->max_frames > 65535 || cfg->max_rows < 1 || cfg->max_cols < 1)
        return fail(nullptr, SPX_ERR_ARG, "bad capacity / Cloud.Dis (1 <= max_frames <= 65535)");
    if (cfg->normal_smoothing_size != 10.0f) return fail(nullptr, SPX_ERR_ARG, "only normal_smoothing_size = 10 is supported");
    if (cfg->min_size < 0 || cfg->ransac_max_iter < 1) return fail(nullptr, SPX_ERR_ARG, "bad Plane.MinSize / RANSAC iteration count");
    if (cfg->normal_method != 0 && cfg->normal_method != 1) return fail(nullptr, SPX_ERR_ARG, "normal_method must be 0 (AVERAGE_3D_GRADIENT) or 1 (COVARIANCE_MATRIX)");
    int w, h;
    cloud_dims(cfg->max_rows, cfg->max_cols, cfg->cloud_dis, &w, &h);
    if (w > kMaxW) return fail(nullptr, SPX_ERR_ARG, "organized cloud wider than %d columns", kMaxW);
    if (size_t(w) * h >= (1u << 20)) return fail(nullptr, SPX_ERR_ARG, "organized cloud larger than 2^20 points");
    if (size_t(w + 2) * (h + 2) > 200u * 1024u) return fail(nullptr, SPX_ERR_ARG, "organized cloud too large for the shared-memory plane-id map of the contour trace");

    // A batch runs as several frame groups on their own streams plus an upload and a download stream.  The driver maps
    // streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams that share a queue serialise, which
    // costs the host path ~20 %.  The variable is read when the process creates its CUDA context, so it is the HOST's to
    // set (the Python binding and bench.py export 32 before CUDA starts): this library does not touch the
    // process environment.
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, SPX_ERR_CUDA, "no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, SPX_ERR_ARG, "device %d out of range (%d devices)", cfg->device, n_dev);
    cudaDeviceProp prop;
    DeviceGuard dev_guard_(cfg->device);
    if ((e = dev_guard_.err) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
        return fail(nullptr, SPX_ERR_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SPX_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

    spx_ctx *c = new (std::nothrow) spx_ctx();
    if (!c) return fail(nullptr, SPX_ERR_ARG, "out of host memory");
    c->cfg = *cfg;
    c->device = cfg->device;
    c->cap_w = w; c->cap_h = h; c->capN = w * h;
    float grid[64];
    c->n_grid = grid_values(grid);

    Params &P = c->P;
    std::memset(&P, 0, sizeof(P));
    P.dis = cfg->cloud_dis;
    P.fx = cfg->fx; P.fy = cfg->fy; P.cx = cfg->cx; P.cy = cfg->cy;
    P.min_x = cfg->min_x; P.max_x = cfg->max_x; P.min_y = cfg->min_y; P.max_y = cfg->max_y;
    P.mdcf = cfg->max_depth_change_factor;
    P.min_size = cfg->min_size;
    P.ang_cos = cosf(float(0.017453 * cfg->angle_thr_deg));   // src/Frame.cc:900 -> setAngularThreshold stores cosf()
    P.dist_thr = cfg->dist_thr;
    P.line_ratio = cfg->line_ratio;
    P.line_thr = double(cfg->line_dist_thr);
    P.ransac_max_iter = cfg->ransac_max_iter;
    P.enable_supposed = cfg->enable_supposed ? 1 : 0;
    P.n_grid = c->n_grid;
    const size_t N = size_t(c->capN), F = size_t(cfg->max_frames);
    P.contour_cap = int(2 * N + 16 * SPX_MAX_MODELS);
    P.pts_cap = int(N + size_t(P.contour_cap) + size_t(32) * c->n_grid * c->n_grid);
    P.bnd_cap = 2 * P.contour_cap;

#define SPX_CK_CREATE(call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            fail(nullptr, SPX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                             \
            spx_destroy(c);                                                                                   \
            return SPX_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)

    SPX_CK_CREATE(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    SPX_CK_CREATE(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    SPX_CK_CREATE(cudaStreamCreateWithFlags(&c->up2_stream, cudaStreamNonBlocking));
    if (const char *e = std::getenv("SPX_UP2")) c->up2 = std::atoi(e) != 0;   // tuning knob
    SPX_CK_CREATE(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) SPX_CK_CREATE(cudaEventCreate(&c->ev[i]));
    c->n_streams = cfg->n_streams > 0 ? (cfg->n_streams > 32 ? 32 : cfg->n_streams) : 8;
    c->min_group = 32;
    if (const char *e = std::getenv("SPX_REFINE_FAST_MAX")) c->refine_fast_max = std::atoi(e);   // tuning knob
    if (const char *e = std::getenv("SPX_LINES_GLOBAL")) c->P.lines_in_global = std::atoi(e) != 0 ? 1 : 0;   // test knob
    if (const char *e = std::getenv("SPX_UPLOAD")) c->upload_mode = std::atoi(e);   // test / tuning knob (see spx_ctx::upload_mode)
    if (const char *e = std::getenv("SPX_SPARSE_MIN_FRAMES")) c->sparse_min_frames = std::atoi(e);   // tuning knob
    if (const char *e = std::getenv("SPX_GATHER_MIN_FRAMES")) c->gather_min_frames = std::atoi(e);   // tuning knob
    if (const char *e = std::getenv("SPX_GATHER_MIN_MB")) c->gather_min_bytes = size_t(std::atoi(e)) << 20;   // tuning / test knob
    if (const char *e = std::getenv("SPX_GATHER_THREADS")) c->gather_threads = std::atoi(e);         // tuning knob (spx_set_gather_threads)
    if (const char *e = std::getenv("SPX_GATHER_SHARE")) { const double v = std::atof(e); if (v <= 1.0) c->gather_share = v; }   // tuning knob
    if (const char *e = std::getenv("SPX_STRIP_OCC")) { const int v = std::atoi(e); if (v == 3 || v == 4) c->strip_occ = v; }   // tuning knob
    if (const char *e = std::getenv("SPX_NORMALS")) { const int v = std::atoi(e); if (v >= 0 && v <= 2) c->normals_mode = v; }   // test knob
    if (const char *e = std::getenv("SPX_CCL_FOUR")) c->ccl_four = std::atoi(e) != 0;   // test knob: the one-pixel-per-thread kernel
    if (const char *e = std::getenv("SPX_FLATTEN_RUNS")) c->flatten_runs = std::atoi(e) != 0;   // test knob
    if (const char *e = std::getenv("SPX_CCL_FRAME")) c->ccl_frame = std::atoi(e);   // test knob
    if (const char *e = std::getenv("SPX_REFINE_PER_GROUP")) c->refine_per_group = std::atoi(e) != 0;   // tuning knob
    if (const char *e = std::getenv("SPX_REFINE_DEV_GROUP")) c->refine_dev_group = std::atoi(e) != 0;   // tuning knob
    if (const char *e = std::getenv("SPX_EDGE_WEIGHT")) { const double v = std::atof(e); if (v > 0.05 && v <= 1.0) c->edge_weight = v; }   // tuning knob
    if (const char *e = std::getenv("SPX_GROUP_WEIGHTS")) {   // tuning knob
        const char *q = e;
        while (*q) { char *end = nullptr; const double v = std::strtod(q, &end); if (end == q) break; if (v > 0) c->group_weights.push_back(v); q = (*end == ',') ? end + 1 : end; }
    }
    if (const char *e = std::getenv("SPX_GROUP_PRIO")) c->group_prio = std::atoi(e);   // tuning knob
    if (const char *e = std::getenv("SPX_BORDER_GRID")) c->border_grid_cap = std::atoi(e);   // tuning knob
    if (const char *e = std::getenv("SPX_PRIO")) c->use_prio = std::atoi(e) != 0;   // tuning knob
    if (const char *e = std::getenv("SPX_MIN_GROUP")) { const int v = std::atoi(e); if (v > 0) c->min_group = v; }   // tuning knob
    int prio_lo = 0, prio_hi = 0;
    SPX_CK_CREATE(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));   // numerically lower = higher priority
    for (int g = 0; g < c->n_streams; ++g) {
        cudaStream_t gs;
        SPX_CK_CREATE(cudaStreamCreateWithPriority(&gs, cudaStreamNonBlocking, prio_lo));
        c->g_streams.push_back(gs);
        int pr = prio_lo;
        if (c->group_prio) {   // prio_hi (numerically lowest) for group 0, falling to prio_lo for the last group
            const int levels = prio_lo - prio_hi + 1;
            pr = prio_hi + std::min(levels - 1, g * levels / c->n_streams);
        }
        SPX_CK_CREATE(cudaStreamCreateWithPriority(&gs, cudaStreamNonBlocking, pr));
        c->g_hstreams.push_back(gs);
        SPX_CK_CREATE(cudaStreamCreateWithPriority(&gs, cudaStreamNonBlocking, prio_hi));
        c->g_back.push_back(gs);
        cudaEvent_t le;
        SPX_CK_CREATE(cudaEventCreateWithFlags(&le, cudaEventDisableTiming));
        c->g_link.push_back(le);
        SPX_CK_CREATE(cudaEventCreateWithFlags(&le, cudaEventDisableTiming));
        c->g_side.push_back(le);
        for (int k = 0; k < 3; ++k) {
            cudaEvent_t e;
            SPX_CK_CREATE(cudaEventCreate(&e));
            c->g_ev.push_back(e);
        }
        for (int k = 0; k < 2; ++k) {
            cudaEvent_t e;
            SPX_CK_CREATE(cudaEventCreate(&e));
            c->g_xev.push_back(e);
        }
        cudaEvent_t te;
        SPX_CK_CREATE(cudaEventCreateWithFlags(&te, cudaEventDisableTiming));
        c->g_tot_ev.push_back(te);
        SPX_CK_CREATE(cudaEventCreateWithFlags(&te, cudaEventDisableTiming));
        c->g_real_ev.push_back(te);
        c->g_out.push_back(spx_ctx::GroupOut());
    }

    const size_t FN = F * N, FC = F * size_t(P.contour_cap);
    size_t total = 0;
    total += 8 * padded<float>(FN);                        // px py pz dist nx ny nz pd
    total += 2 * padded<uint8_t>(FN) + 5 * padded<int>(FN);    // conn kwin | parent cnt lab pos cand_idx
    const size_t n_cham = F * size_t(cdiv(h, kBandRows)) * size_t(kBandRows + kBandHalo) * size_t(w);
    total += padded<float>(n_cham);
    total += padded<int16_t>(FN) + 2 * padded<int8_t>(FN + 16 * F);     // root_model pid pid_bak
    total += 3 * padded<int>(FC) + padded<float4>(FC) + padded<spx_point>(FC) + padded<int>(c->n_streams * (2 + F * SPX_MAX_MODELS)) + padded<int>(c->n_streams * (2 + F * SPX_MAX_MODELS * SPX_MAX_LINES));
    total += padded<FrameCtl>(F);
    total += padded<int>(size_t(c->n_streams) * (1 + F));   // nf_list
    const size_t fetch_words = F * size_t(cfg->max_rows) * size_t(((cfg->max_cols + 7) / 8 + 31) / 32);
    total += padded<unsigned>(fetch_words);
    total += padded<spx_frame_header>(F) + padded<spx_plane>(F * SPX_MAX_PLANES);
    total += padded<spx_point>(F * size_t(P.pts_cap)) + padded<spx_point>(F * size_t(P.bnd_cap));
    total += padded<long long>(8 * size_t(c->n_streams + 1)) + padded<long long>(5 * F);
    total += padded<uint32_t>(FN);                         // out_pidx
    const size_t cov_nodes = cfg->normal_method == 1 ? F * size_t(w + 1) * size_t(h + 1) : 0;
    total += padded<double>(cov_nodes * kCovCh) + padded<unsigned>(cov_nodes) + padded<float>(cfg->normal_method == 1 ? FN : 0);
    total += padded<float>(F * size_t(cfg->max_rows) * size_t(cfg->max_cols) + 4) + padded<uint16_t>(F * size_t(cfg->max_rows) * size_t(cfg->max_cols) + 4);
    const size_t samp_floats = F * size_t(h) * ((size_t(w) + 3) & ~size_t(3)) + 4;
    total += padded<float>(samp_floats);
    SPX_CK_CREATE(cudaMalloc(reinterpret_cast<void **>(&c->arena.base), total));
    c->arena.size = total;
    DevArena &A = c->arena;
    Buffers &B = c->B;
    B.px = A.take<float>(FN); B.py = A.take<float>(FN); B.pz = A.take<float>(FN); B.dist = A.take<float>(FN);
    B.nx = A.take<float>(FN); B.ny = A.take<float>(FN); B.nz = A.take<float>(FN); B.pd = A.take<float>(FN);
    B.conn = A.take<uint8_t>(FN); B.kwin = A.take<uint8_t>(FN); B.cham_tmp = A.take<float>(n_cham);
    B.parent = A.take<int>(FN); B.cnt = A.take<int>(FN); B.lab = A.take<int>(FN); B.pos = A.take<int>(FN); B.cand_idx = A.take<int>(FN);
    B.root_model = A.take<int16_t>(FN); B.pid = A.take<int8_t>(FN + 16 * F); B.pid_bak = A.take<int8_t>(FN + 16 * F);
    B.contour_idx = A.take<int>(FC); B.line_sh = A.take<int>(FC); B.line_inl = A.take<int>(FC);
    B.line_a = A.take<float4>(FC); B.line_pts = A.take<spx_point>(FC); c->work_stride = 2 + F * SPX_MAX_MODELS; c->work2_stride = 2 + F * SPX_MAX_MODELS * SPX_MAX_LINES;
    B.work = A.take<int>(c->n_streams * c->work_stride); B.work2 = A.take<int>(c->n_streams * c->work2_stride);
    B.ctl = A.take<FrameCtl>(F);
    B.nf_list = A.take<int>(size_t(c->n_streams) * (1 + F));
    B.fetch_bits = A.take<unsigned>(fetch_words);
    B.out_frames = A.take<spx_frame_header>(F); B.out_planes = A.take<spx_plane>(F * SPX_MAX_PLANES);
    B.out_pts = A.take<spx_point>(F * size_t(P.pts_cap)); B.out_bnd = A.take<spx_point>(F * size_t(P.bnd_cap));
    B.out_totals = A.take<long long>(8 * size_t(c->n_streams + 1)); B.frame_offs = A.take<long long>(5 * F);
    B.out_pidx = A.take<uint32_t>(FN);
    if (cfg->normal_method == 1) { c->d_cov_sat = A.take<double>(cov_nodes * kCovCh); c->d_cov_cnt = A.take<unsigned>(cov_nodes); c->d_curv = A.take<float>(FN); }
    c->d_depth = A.take<float>(F * size_t(cfg->max_rows) * size_t(cfg->max_cols) + 4);
    c->d_depth16 = A.take<uint16_t>(F * size_t(cfg->max_rows) * size_t(cfg->max_cols) + 4);
    c->d_samp = A.take<float>(samp_floats);
    if (A.used > A.size) { fail(nullptr, SPX_ERR_ARG, "internal: arena accounting"); spx_destroy(c); return SPX_ERR_ARG; }

    uint32_t mt[624], mt_out[624];
    mt19937_seeded_state(12345u, mt);   // boost::mt19937 rng_alg_ seeded in SampleConsensusModel's ctor (random = false)
    mt19937_twist(mt);
    for (int i = 0; i < 624; ++i) mt_out[i] = mt19937_temper(mt[i]);
    SPX_CK_CREATE(cudaMemcpyToSymbol(c_mt_state1, mt, sizeof(mt)));
    SPX_CK_CREATE(cudaMemcpyToSymbol(c_mt_out0, mt_out, sizeof(mt_out)));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kLinesSmem)));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_refine2<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(refine2_smem_bytes(kRefMaxH, 7))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_refine2<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(refine2_smem_bytes(kRefMaxH, 14))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_refine2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(refine2_smem_bytes(kRefMaxH, 16))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_contour, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(w + 2) * (h + 2))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_ccl_frame, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(std::min(size_t(c->capN) * sizeof(unsigned short), size_t(72) * 1024))));
    {
        int per_sm = 0;
        SPX_CK_CREATE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lines, kLineThreads, kLinesSmem));
        c->lines_grid = prop.multiProcessorCount * (per_sm > 0 ? per_sm : 1);
        SPX_CK_CREATE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_border, kBorderWarps * 32, 0));
        c->border_grid = prop.multiProcessorCount * (per_sm > 0 ? per_sm : 1);
        if (c->border_grid_cap > 0) c->border_grid = std::min(c->border_grid, prop.multiProcessorCount * c->border_grid_cap);
    }
    SPX_CK_CREATE(cudaMemcpyToSymbol(c_grid, grid, sizeof(grid)));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_link, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kNormalsSmem)));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_link_list, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kNormalsSmem)));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_strip<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(strip_smem_bytes(P.dis, strip_tma_ok(P.dis)))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_strip<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(strip_smem_bytes(P.dis, false))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_strip<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(strip_smem_bytes(P.dis, strip_tma_ok(P.dis)))));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_normals_strip<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(strip_smem_bytes(P.dis, false))));
    c->list_grid = prop.multiProcessorCount * 2;
    c->sm_count = prop.multiProcessorCount;
    if (const char *e = std::getenv("SPX_STRIP_ALWAYS")) c->strip_always = std::atoi(e) != 0;   // test knob
    {
        // cuTensorMapEncodeTiled without linking libcuda
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            c->encode_tiled = reinterpret_cast<spx_ctx::EncodeTiledFn>(fn);
        else
            cudaGetLastError();
        // the strip kernel divides by the constants fx, fy with a reciprocal and one fused correction: verified here against
        // the IEEE quotient for every significand, per constant; a constant that fails keeps the true division
        P.rfx = 1.0f / P.fx; P.rfy = 1.0f / P.fy; P.fast_div = 0;
        int *d_mis = nullptr, h_mis[2] = {1, 1};
        SPX_CK_CREATE(cudaMalloc(reinterpret_cast<void **>(&d_mis), 2 * sizeof(int)));
        cudaMemset(d_mis, 0, 2 * sizeof(int));
        if (std::isfinite(P.rfx) && P.fx != 0.0f) k_check_div<<<(1u << 23) / 256, 256>>>(P.fx, P.rfx, d_mis); else h_mis[0] = -1;
        if (std::isfinite(P.rfy) && P.fy != 0.0f) k_check_div<<<(1u << 23) / 256, 256>>>(P.fy, P.rfy, d_mis + 1); else h_mis[1] = -1;
        int got[2] = {1, 1};
        const cudaError_t ce = cudaMemcpy(got, d_mis, sizeof(got), cudaMemcpyDeviceToHost);
        cudaFree(d_mis);
        if (ce != cudaSuccess) { fail(nullptr, SPX_ERR_CUDA, "division check: %s", cudaGetErrorString(ce)); spx_destroy(c); return SPX_ERR_CUDA; }
        if (h_mis[0] >= 0 && got[0] == 0) P.fast_div |= 1;
        if (h_mis[1] >= 0 && got[1] == 0) P.fast_div |= 2;
    }
    SPX_CK_CREATE(cudaFuncSetAttribute(k_edge_chamfer<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    SPX_CK_CREATE(cudaFuncSetAttribute(k_edge_chamfer<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    SPX_CK_CREATE(cudaHostAlloc(reinterpret_cast<void **>(&c->h_totals), 8 * size_t(c->n_streams + 1) * sizeof(long long), cudaHostAllocMapped));
    SPX_CK_CREATE(cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->h_totals_dev), c->h_totals, 0));
    SPX_CK_CREATE(cudaHostAlloc(reinterpret_cast<void **>(&c->h_frames), F * sizeof(spx_frame_header), cudaHostAllocDefault));
#undef SPX_CK_CREATE
    *out = c;
    return SPX_OK;
}

void spx_destroy(spx_ctx *c) {
    if (!c) return;
    DeviceGuard dev_guard_(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->h_totals) cudaFreeHost(c->h_totals);
    if (c->h_frames) cudaFreeHost(c->h_frames);
    if (c->h_planes) cudaFreeHost(c->h_planes);
    if (c->h_pts) cudaFreeHost(c->h_pts);
    if (c->h_bnd) cudaFreeHost(c->h_bnd);
    if (c->h_pidx) cudaFreeHost(c->h_pidx);
    delete c->pool;                      // (joins the gather threads)
    if (c->h_samp) cudaFreeHost(c->h_samp);
    for (void *p : c->graveyard) cudaFreeHost(p);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->g_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->g_tot_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->g_real_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->g_xev) cudaEventDestroy(e);
    for (cudaStream_t gs : c->g_streams) { cudaStreamSynchronize(gs); cudaStreamDestroy(gs); }
    for (cudaStream_t gs : c->g_hstreams) { cudaStreamSynchronize(gs); cudaStreamDestroy(gs); }
    for (cudaStream_t gs : c->g_back) { cudaStreamSynchronize(gs); cudaStreamDestroy(gs); }
    for (cudaEvent_t e : c->g_link) cudaEventDestroy(e);
    for (cudaEvent_t e : c->g_side) cudaEventDestroy(e);
    if (c->up_stream) { cudaStreamSynchronize(c->up_stream); cudaStreamDestroy(c->up_stream); }
    if (c->up2_stream) { cudaStreamSynchronize(c->up2_stream); cudaStreamDestroy(c->up2_stream); }
    if (c->down_stream) { cudaStreamSynchronize(c->down_stream); cudaStreamDestroy(c->down_stream); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char *spx_last_error(const spx_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int spx_set_stream(spx_ctx *c, void *cuda_stream) {
    if (!c) return SPX_ERR_ARG;
    SPX_DEVICE(c);
    SPX_CK(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return SPX_OK;
}

int spx_set_debug(spx_ctx *c, int on) {
    if (!c) return SPX_ERR_ARG;
    c->debug = on != 0;
    return SPX_OK;
}

int spx_extract_batch_device(spx_ctx *c, const float *depth_dev, int n_frames, int rows, int cols, size_t pitch_bytes,
                             size_t frame_stride_bytes) {
    if (!c) return SPX_ERR_ARG;
    if (!depth_dev) return fail(c, SPX_ERR_ARG, "null depth pointer");
    SPX_DEVICE(c);
    int rc = set_geometry(c, n_frames, rows, cols, pitch_bytes, frame_stride_bytes);
    if (rc != SPX_OK) return rc;
    return run_pipeline(c, depth_dev, depth_dev, false, HostSrc(), false, c->result_mode == 1);
}

int spx_fetch_results(spx_ctx *c, spx_batch_result *out) {
    if (!c || !out) return SPX_ERR_ARG;
    SPX_DEVICE(c);
    return fetch(c, out, true);
}

int spx_fetch_planes(spx_ctx *c, spx_batch_result *out) {
    if (!c || !out) return SPX_ERR_ARG;
    SPX_DEVICE(c);
    return fetch(c, out, false);
}

int spx_get_device_results(spx_ctx *c, spx_device_result *out) {
    if (!c || !out) return SPX_ERR_ARG;
    if (!c->have_run || c->group_pack) return fail(c, SPX_ERR_STATE, "device results follow spx_extract_batch_device");
    if (c->compact) return fail(c, SPX_ERR_STATE, "the last extract packed compact results (spx_set_result_mode): the device view describes point clouds");
    out->n_frames = c->last_frames;
    out->frames = c->B.out_frames; out->planes = c->B.out_planes; out->points = c->B.out_pts; out->boundary = c->B.out_bnd;
    out->totals = c->B.out_totals;
    out->planes_capacity = int64_t(c->cfg.max_frames) * SPX_MAX_PLANES;
    return SPX_OK;
}

// host input, float or 16-bit, results as point clouds (`out`) or compact (`cout`)
static int extract_host(spx_ctx *c, const void *depth, bool u16, float depth_map_factor, int n_frames, int rows, int cols, size_t pitch_bytes,
                        size_t frame_stride_bytes, spx_batch_result *out, spx_compact_result *cout) {
    if (!c) return SPX_ERR_ARG;
    if (!depth || (!out && !cout)) return fail(c, SPX_ERR_ARG, "null argument");
    SPX_DEVICE(c);
    const size_t tight = size_t(cols) * sizeof(float);
    const size_t esz = u16 ? sizeof(uint16_t) : sizeof(float);
    int rc;
    if (u16) {
        // geometry checks are stated for the float image; the 16-bit pitch / stride are validated here
        if (pitch_bytes < size_t(cols) * sizeof(uint16_t) || pitch_bytes % sizeof(uint16_t)) return fail(c, SPX_ERR_ARG, "bad pitch");
        if (n_frames > 1 && frame_stride_bytes < pitch_bytes * size_t(rows > 0 ? rows : 0)) return fail(c, SPX_ERR_ARG, "bad frame stride");
        if ((size_t(rows) * size_t(cols)) % 4 != 0 && n_frames > 1) return fail(c, SPX_ERR_ARG, "16-bit batches need rows * cols to be a multiple of 4");
        rc = set_geometry(c, n_frames, rows, cols, tight, tight * size_t(rows));
    } else {
        rc = set_geometry(c, n_frames, rows, cols, pitch_bytes, frame_stride_bytes);
    }
    if (rc != SPX_OK) return rc;
    HostSrc src;
    src.depth = depth; src.pitch = pitch_bytes; src.frame_stride = frame_stride_bytes; src.u16 = u16; src.alpha = u16 ? depth_map_factor : 1.0f;
    const void *full = c->d_depth;
    c->P.pitch = tight; c->P.frame_stride = tight * size_t(rows);   // layout of the device image the kernels read
    c->P.samp_rstep = tight * size_t(c->P.dis); c->P.samp_fstride = c->P.frame_stride;
    if ((!u16 || cols % 4 == 0) &&
        sparse_upload(c, depth, frame_stride_bytes * size_t(n_frames - 1) + pitch_bytes * size_t(rows - 1) + size_t(cols) * esz, n_frames, &src)) {
        // only the sampled rows are uploaded (to their place in the device image); k_border_fetch adds the window sectors
        c->P.host_pitch = pitch_bytes; c->P.host_fstride = frame_stride_bytes; c->P.fetch_skip_sampled = 1;
        if (u16) c->P.full_alpha = depth_map_factor;
        c->P.fetch_vec = (reinterpret_cast<uintptr_t>(src.mapped) % 16 == 0 && pitch_bytes % 16 == 0 && frame_stride_bytes % 16 == 0 && cols % 8 == 0) ? 1 : 0;
        // float images: the samples themselves instead of the sampled rows (gathered by host threads, see GatherPool)
        // (automatic mode: for the compact call only -- with 16-byte clouds coming back the call is bound by the downloads, and the
        // gather threads' memory traffic then costs more than the smaller upload saves: 14.5 against 12.5 ms per 1000 frames)
        if (!u16 && (c->upload_mode == 3 || (c->upload_mode == 0 && cout != nullptr && n_frames >= c->gather_min_frames &&
                                             size_t(n_frames) * size_t(c->P.h) * size_t(cols) * sizeof(float) >= c->gather_min_bytes))) {
            const size_t sw = (size_t(c->P.w) + 3) & ~size_t(3);             // rows of the staging buffer start on 16 bytes (TMA)
            if ((rc = grow_pinned(c, &c->h_samp, &c->h_samp_cap, size_t(n_frames) * size_t(c->P.h) * sw)) != SPX_OK) return rc;
            if (!c->pool) c->pool = new (std::nothrow) GatherPool();
            if (!c->pool) return fail(c, SPX_ERR_ARG, "out of host memory");
            int nt = c->gather_threads;
            if (nt <= 0) { nt = int(std::thread::hardware_concurrency()) / 2; if (nt > 16) nt = 16; }   // (the caller's own threads need cores too)
            c->pool->start(nt < 1 ? 1 : nt);
            src.gather = true;                                                // (which groups: run_pipeline; their layout: run_group)
            c->g_rstep = sw * sizeof(float); c->g_fstride = sw * sizeof(float) * size_t(c->P.h);
        }
    } else {
        src.sparse = false;
    }
    rc = run_pipeline(c, c->d_depth, full, false, src, true, cout != nullptr);
    if (src.gather) c->pool->wait_idle();     // (whatever happened: the host threads read the caller's image)
    if (rc != SPX_OK) return rc;
    return fetch_groups(c, out, cout);
}

int spx_extract_batch(spx_ctx *c, const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                      size_t frame_stride_bytes, spx_batch_result *out) {
    if (c && !out) return fail(c, SPX_ERR_ARG, "null argument");
    return extract_host(c, depth, false, 1.0f, n_frames, rows, cols, pitch_bytes, frame_stride_bytes, out, nullptr);
}

int spx_extract_batch_u16(spx_ctx *c, const uint16_t *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                          size_t frame_stride_bytes, float depth_map_factor, spx_batch_result *out) {
    if (c && !out) return fail(c, SPX_ERR_ARG, "null argument");
    return extract_host(c, depth, true, depth_map_factor, n_frames, rows, cols, pitch_bytes, frame_stride_bytes, out, nullptr);
}

int spx_extract_batch_compact(spx_ctx *c, const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                              size_t frame_stride_bytes, spx_compact_result *out) {
    if (c && !out) return fail(c, SPX_ERR_ARG, "null argument");
    return extract_host(c, depth, false, 1.0f, n_frames, rows, cols, pitch_bytes, frame_stride_bytes, nullptr, out);
}

int spx_extract_batch_u16_compact(spx_ctx *c, const uint16_t *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                                  size_t frame_stride_bytes, float depth_map_factor, spx_compact_result *out) {
    if (c && !out) return fail(c, SPX_ERR_ARG, "null argument");
    return extract_host(c, depth, true, depth_map_factor, n_frames, rows, cols, pitch_bytes, frame_stride_bytes, nullptr, out);
}

int spx_set_result_mode(spx_ctx *c, int mode) {
    if (!c || mode < 0 || mode > 1) return SPX_ERR_ARG;
    c->result_mode = mode;
    return SPX_OK;
}

int spx_fetch_compact(spx_ctx *c, spx_compact_result *out) {
    if (!c || !out) return SPX_ERR_ARG;
    SPX_DEVICE(c);
    return fetch(c, nullptr, true, out);
}

int spx_set_group_callback(spx_ctx *c, spx_group_fn fn, void *user) {
    if (!c) return SPX_ERR_ARG;
    c->group_fn = fn; c->group_user = user;
    return SPX_OK;
}

int spx_extract(spx_ctx *c, const float *depth, int rows, int cols, size_t pitch_bytes, spx_batch_result *out) {
    return spx_extract_batch(c, depth, 1, rows, cols, pitch_bytes, pitch_bytes * size_t(rows > 0 ? rows : 0), out);
}

int spx_segment_from_normals(spx_ctx *c, const float *depth, int rows, int cols, size_t pitch_bytes, const float *normals,
                             spx_batch_result *out) {
    if (!c) return SPX_ERR_ARG;
    if (!depth || !normals || !out) return fail(c, SPX_ERR_ARG, "null argument");
    SPX_DEVICE(c);
    const size_t tight = size_t(cols) * sizeof(float);
    int rc = set_geometry(c, 1, rows, cols, pitch_bytes, pitch_bytes * size_t(rows));
    if (rc != SPX_OK) return rc;
    HostSrc src;
    src.depth = depth; src.pitch = pitch_bytes; src.frame_stride = pitch_bytes * size_t(rows);
    c->P.pitch = tight; c->P.frame_stride = tight * size_t(rows);
    c->P.samp_rstep = tight * size_t(c->P.dis); c->P.samp_fstride = c->P.frame_stride;
    const size_t N = size_t(c->P.N);
    SPX_CK(c, cudaMemcpyAsync(c->B.nx, normals, N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SPX_CK(c, cudaMemcpyAsync(c->B.ny, normals + N, N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SPX_CK(c, cudaMemcpyAsync(c->B.nz, normals + 2 * N, N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if ((rc = run_pipeline(c, c->d_depth, c->d_depth, true, src, false)) != SPX_OK) return rc;
    return fetch(c, out, true);
}

int spx_cloud_dims(const spx_ctx *c, int rows, int cols, int *width, int *height) {
    if (!c || !width || !height || rows < 1 || cols < 1) return SPX_ERR_ARG;
    cloud_dims(rows, cols, c->P.dis, width, height);
    return SPX_OK;
}

int spx_get_times(spx_ctx *c, double *t_plane, double *t_splane) {
    if (!c) return SPX_ERR_ARG;
    if (!c->have_run) return fail(c, SPX_ERR_STATE, "no extract call has been made on this context");
    SPX_DEVICE(c);
    SPX_CK(c, cudaEventSynchronize(c->ev[2]));
    float total = 0;
    SPX_CK(c, cudaEventElapsedTime(&total, c->ev[0], c->ev[2]));
    // frame groups overlap on the device: the supposed-plane share of the groups' own time is applied to the wall time
    double sum_all = 0, sum_sp = 0;
    for (int g = 0; g < c->last_groups; ++g) {
        float a = 0, b = 0;
        SPX_CK(c, cudaEventElapsedTime(&a, c->g_ev[3 * g + 0], c->g_ev[3 * g + 2]));
        SPX_CK(c, cudaEventElapsedTime(&b, c->g_ev[3 * g + 1], c->g_ev[3 * g + 2]));
        sum_all += a; sum_sp += b;
    }
    const double sp = sum_all > 0 ? double(total) * sum_sp / sum_all : 0.0;
    if (t_plane) *t_plane = (double(total) - sp) * 1e-3;   // segmentation + packing of the clouds
    if (t_splane) *t_splane = sp * 1e-3;
    return SPX_OK;
}

int spx_get_group_timeline(spx_ctx *c, float *t_ms, int cap_groups, int *n_groups) {
    if (!c || !n_groups) return SPX_ERR_ARG;
    if (!c->have_run || !c->group_pack) return fail(c, SPX_ERR_STATE, "the group timeline follows a host-input extract");
    SPX_DEVICE(c);
    *n_groups = c->last_groups;
    for (int g = 0; g < c->last_groups && g < cap_groups && t_ms; ++g) {
        cudaEvent_t evs[5] = {c->g_ev[3 * g + 0], c->g_xev[2 * g + 0], c->g_ev[3 * g + 1], c->g_ev[3 * g + 2], c->g_xev[2 * g + 1]};
        for (int k = 0; k < 5; ++k) SPX_CK(c, cudaEventElapsedTime(&t_ms[7 * g + k], c->ev[0], evs[k]));
        t_ms[7 * g + 5] = c->g_host_ms[2 * g]; t_ms[7 * g + 6] = c->g_host_ms[2 * g + 1];
    }
    return SPX_OK;
}

int spx_last_launch_count(const spx_ctx *c) { return c ? c->launches : 0; }

int spx_get_transfer_bytes(const spx_ctx *c, unsigned long long *h2d_copied, unsigned long long *h2d_in_place, unsigned long long *d2h) {
    if (!c) return SPX_ERR_ARG;
    if (h2d_copied) *h2d_copied = c->xfer_h2d;
    if (h2d_in_place) *h2d_in_place = c->xfer_inplace;
    if (d2h) *d2h = c->xfer_d2h;
    return SPX_OK;
}

int spx_set_upload_mode(spx_ctx *c, int mode) {
    if (!c || mode < 0 || mode > 3) return SPX_ERR_ARG;
    c->upload_mode = mode;
    return SPX_OK;
}

int spx_set_gather_threads(spx_ctx *c, int n_threads) {
    if (!c || n_threads < 0 || n_threads > 256) return SPX_ERR_ARG;
    if (c->pool && n_threads > 0 && int(c->pool->workers.size()) > n_threads) {   // shrink: the pool is rebuilt at the next call
        delete c->pool;
        c->pool = nullptr;
    }
    c->gather_threads = n_threads;
    return SPX_OK;
}

int spx_set_gather_share(spx_ctx *c, double share) {
    if (!c || !(share <= 1.0)) return SPX_ERR_ARG;      // (negative: chosen from the thread count)
    c->gather_share = share;
    return SPX_OK;
}

int spx_host_gather_samples(const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes, size_t frame_stride_bytes, int cloud_dis,
                            int n_groups, int n_threads, float *out, size_t out_row_floats) {
    if (!depth || !out || n_frames < 1 || rows < 1 || cols < 1 || cloud_dis < 1 || n_groups < 1 || n_threads < 1 || n_threads > 256) return SPX_ERR_ARG;
    int w = 0, h = 0;
    cloud_dims(rows, cols, cloud_dis, &w, &h);
    if (out_row_floats < size_t(w) || pitch_bytes < size_t(cols) * sizeof(float)) return SPX_ERR_ARG;
    GatherPool pool;
    pool.start(n_threads);
    std::vector<int> bounds(size_t(n_groups) + 1, 0);
    for (int g = 1; g <= n_groups; ++g) bounds[size_t(g)] = std::max(bounds[size_t(g) - 1], int((long long)n_frames * g / n_groups));
    pool.post(reinterpret_cast<const char *>(depth), pitch_bytes, frame_stride_bytes, out, h, w, int(out_row_floats), cloud_dis, bounds);
    for (int g = 0; g < n_groups; ++g) pool.wait_group(g);      // in group order, as the upload path does
    pool.wait_idle();
    return SPX_OK;
}

int spx_host_register(void *ptr, size_t bytes) {
    if (!ptr || !bytes) return SPX_ERR_ARG;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); fail(nullptr, SPX_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); return SPX_ERR_CUDA; }
    return SPX_OK;
}

int spx_host_unregister(void *ptr) {
    if (!ptr) return SPX_ERR_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); fail(nullptr, SPX_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); return SPX_ERR_CUDA; }
    return SPX_OK;
}

int spx_set_profile(spx_ctx *c, int on) {
    if (!c) return SPX_ERR_ARG;
    c->profile = on != 0;
    return SPX_OK;
}

int spx_get_kernel_times(spx_ctx *c, const char **names, float *ms, int cap, int *n) {
    if (!c || !n) return SPX_ERR_ARG;
    if (!c->have_run || !c->profile || c->prof_n == 0) return fail(c, SPX_ERR_STATE, "no profiled extract call (spx_set_profile) on this context");
    SPX_DEVICE(c);
    SPX_CK(c, cudaEventSynchronize(c->ev[2]));
    *n = c->prof_n;
    for (int k = 0; k < c->prof_n && k < cap; ++k) {
        if (names) names[k] = c->prof_names[k];
        if (ms) SPX_CK(c, cudaEventElapsedTime(&ms[k], c->prof_ev[2 * k], c->prof_ev[2 * k + 1]));
    }
    return SPX_OK;
}

int spx_get_kernel_timeline(spx_ctx *c, const char **names, float *start_ms, float *end_ms, int cap, int *n) {
    if (!c || !n) return SPX_ERR_ARG;
    if (!c->have_run || !c->profile || c->prof_n == 0) return fail(c, SPX_ERR_STATE, "no profiled extract call (spx_set_profile) on this context");
    SPX_DEVICE(c);
    SPX_CK(c, cudaEventSynchronize(c->ev[2]));
    *n = c->prof_n;
    for (int k = 0; k < c->prof_n && k < cap; ++k) {
        if (names) names[k] = c->prof_names[k];
        if (start_ms) SPX_CK(c, cudaEventElapsedTime(&start_ms[k], c->ev[0], c->prof_ev[2 * k]));
        if (end_ms) SPX_CK(c, cudaEventElapsedTime(&end_ms[k], c->ev[0], c->prof_ev[2 * k + 1]));
    }
    return SPX_OK;
}

// ---- debug taps ----------------------------------------------------------------------------------------------
int spx_get_cloud(spx_ctx *c, int frame, float *x, float *y, float *z) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    const size_t N = size_t(c->P.N), o = N * size_t(frame);
    if ((rc = d2h(c, x, c->B.px + o, N)) != SPX_OK) return rc;
    if ((rc = d2h(c, y, c->B.py + o, N)) != SPX_OK) return rc;
    return d2h(c, z, c->B.pz + o, N);
}

int spx_get_distance_map(spx_ctx *c, int frame, float *dist) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    return d2h(c, dist, c->B.dist + size_t(c->P.N) * size_t(frame), size_t(c->P.N));
}

int spx_get_normals(spx_ctx *c, int frame, float *nx, float *ny, float *nz, float *plane_d) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    const size_t N = size_t(c->P.N), o = N * size_t(frame);
    if ((rc = d2h(c, nx, c->B.nx + o, N)) != SPX_OK) return rc;
    if ((rc = d2h(c, ny, c->B.ny + o, N)) != SPX_OK) return rc;
    if ((rc = d2h(c, nz, c->B.nz + o, N)) != SPX_OK) return rc;
    return d2h(c, plane_d, c->B.pd + o, N);
}

int spx_get_curvature(spx_ctx *c, int frame, float *curvature) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    if (!c->d_curv) return fail(c, SPX_ERR_STATE, "curvature exists under normal_method = 1 (COVARIANCE_MATRIX) only");
    return d2h(c, curvature, c->d_curv + size_t(c->P.N) * size_t(frame), size_t(c->P.N));
}

int spx_get_labels_raw(spx_ctx *c, int frame, uint32_t *labels, int *n_label_lists) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    if ((rc = d2h(c, reinterpret_cast<int *>(labels), c->B.lab + size_t(c->P.N) * size_t(frame), size_t(c->P.N))) != SPX_OK) return rc;
    if (n_label_lists) {
        int n = 0;
        SPX_CK(c, cudaMemcpy(&n, &c->B.ctl[frame].n_labels, sizeof(int), cudaMemcpyDeviceToHost));
        *n_label_lists = n;
    }
    return SPX_OK;
}

int spx_get_plane_ids(spx_ctx *c, int frame, int8_t *ids) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    return d2h(c, ids, c->B.pid + size_t(c->P.N) * size_t(frame), size_t(c->P.N));
}

int spx_get_models(spx_ctx *c, int frame, spx_model_info *models, int *n_models) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    std::vector<FrameCtl> ctl(1);
    if ((rc = get_ctl(c, frame, ctl.data())) != SPX_OK) return rc;
    const FrameCtl &K = ctl[0];
    if (n_models) *n_models = K.n_models;
    if (models)
        for (int i = 0; i < K.n_models; ++i) {
            const Model &M = K.models[i];
            spx_model_info &o = models[i];
            std::memcpy(o.coef, M.coef, sizeof(o.coef));
            std::memcpy(o.centroid, M.centroid, sizeof(o.centroid));
            std::memcpy(o.cov, M.cov, sizeof(o.cov));
            o.curvature = M.curvature; o.label = uint32_t(M.label);
            o.n_segment = M.n0; o.n_inliers = M.n0 + M.n1 + M.n2; o.n_contour = M.n_contour;
        }
    return SPX_OK;
}

int spx_get_model_inliers(spx_ctx *c, int frame, int model, int32_t *idx) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    if (model < 0 || model >= SPX_MAX_MODELS || !idx) return fail(c, SPX_ERR_ARG, "bad model index");
    const size_t N = size_t(c->P.N), o = N * size_t(frame);
    std::vector<int8_t> pid(N);
    std::vector<int> pos(N);
    if ((rc = d2h(c, pid.data(), c->B.pid + o, N)) != SPX_OK) return rc;
    if ((rc = d2h(c, pos.data(), c->B.pos + o, N)) != SPX_OK) return rc;
    for (size_t q = 0; q < N; ++q)
        if (pid[q] == model) idx[pos[q]] = int32_t(q);   // inlier_indices[model].indices[pos] = q
    return SPX_OK;
}

int spx_get_model_contour(spx_ctx *c, int frame, int model, int32_t *idx) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    if (model < 0 || model >= SPX_MAX_MODELS || !idx) return fail(c, SPX_ERR_ARG, "bad model index");
    Model M;
    SPX_CK(c, cudaMemcpy(&M, &c->B.ctl[frame].models[model], sizeof(Model), cudaMemcpyDeviceToHost));
    return d2h(c, idx, c->B.contour_idx + size_t(frame) * size_t(c->P.contour_cap) + M.contour_off, size_t(M.n_contour));
}

int spx_get_lines(spx_ctx *c, int frame, spx_line_info *lines, int *n_lines) {
    int rc = check_frame(c, frame);
    if (rc != SPX_OK) return rc;
    std::vector<FrameCtl> ctl(1);
    if ((rc = get_ctl(c, frame, ctl.data())) != SPX_OK) return rc;
    const FrameCtl &K = ctl[0];
    int n = 0;
    if (c->P.enable_supposed)
        for (int i = K.n_real - 1; i >= 0; --i) {   // the reference visits the real planes last to first
            const int m = K.planes[i].src;
            const Model &M = K.models[m];
            if (M.n_contour < 50) continue;
            for (int j = 0; j < M.n_rounds; ++j) {
                const Line &Ln = K.lines[m * SPX_MAX_LINES + j];
                if (lines) {
                    spx_line_info &o = lines[n];
                    o.plane = i; o.round = j; o.n_points = Ln.n_points; o.iterations = Ln.iterations; o.n_inliers = Ln.n_inliers;
                    o.in_range = Ln.in_range; o.is_border = Ln.is_border; o.emitted = Ln.emitted;
                    std::memcpy(o.coef, Ln.coef, sizeof(o.coef));
                }
                ++n;
            }
        }
    if (n_lines) *n_lines = n;
    return SPX_OK;
}

}  // extern "C"
