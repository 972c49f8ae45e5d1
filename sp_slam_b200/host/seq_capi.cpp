// seq_capi.cpp -- plain-C handle around spx_host::SequencePlanes (libspx_host.so) so that bench.py and the tests can drive
// the C++ host adapter through ctypes: the `e2e_adapter` leg of bench.py ends when every Frame field of every frame of
// the batch is filled (mvPlanePoints / mvBoundaryPoints as 32-byte pcl::PointXYZRGB-layout points, mvPlaneCoefficients).
#include <chrono>
#include <cstring>

#include "SequencePlanes.h"

using spx_host::PlaneFields;
using spx_host::SequencePlanes;

namespace {
thread_local std::string g_err;
uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
}  // namespace

extern "C" {

const char *spx_seq_last_error() { return g_err.c_str(); }

void *spx_seq_create(const spx_config *cfg, int n_threads) {
    try { return new SequencePlanes(*cfg, n_threads); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void spx_seq_destroy(void *h) { delete static_cast<SequencePlanes *>(h); }
int spx_seq_threads(void *h) { return static_cast<SequencePlanes *>(h)->threads(); }
void *spx_seq_context(void *h) { return static_cast<SequencePlanes *>(h)->context(); }

// one batch; returns the wall time of the call in ms (< 0: error)
double spx_seq_process(void *h, const float *depth, int n_frames, int rows, int cols, size_t step, size_t frame_stride) {
    try {
        const auto t0 = std::chrono::steady_clock::now();
        static_cast<SequencePlanes *>(h)->Process(depth, n_frames, rows, cols, step, frame_stride);
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    } catch (const std::exception &e) { g_err = e.what(); return -1.0; }
}
// the same with the real planes' clouds transferred as 16-byte points (spx_extract_batch) instead of index lists
double spx_seq_process_clouds(void *h, const float *depth, int n_frames, int rows, int cols, size_t step, size_t frame_stride) {
    try {
        const auto t0 = std::chrono::steady_clock::now();
        static_cast<SequencePlanes *>(h)->Process(depth, n_frames, rows, cols, step, frame_stride, false);
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    } catch (const std::exception &e) { g_err = e.what(); return -1.0; }
}
double spx_seq_process_u16(void *h, const uint16_t *depth, int n_frames, int rows, int cols, size_t step, size_t frame_stride, float factor) {
    try {
        const auto t0 = std::chrono::steady_clock::now();
        static_cast<SequencePlanes *>(h)->ProcessU16(depth, n_frames, rows, cols, step, frame_stride, factor);
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    } catch (const std::exception &e) { g_err = e.what(); return -1.0; }
}

// totals of the filled fields: planes, points of mvPlanePoints, points of mvBoundaryPoints, bytes they occupy
void spx_seq_summary(void *h, long long out[4]) {
    const SequencePlanes *s = static_cast<SequencePlanes *>(h);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (const PlaneFields &f : s->frames) {
        out[0] += f.mnPlaneNum;
        for (const auto &c : f.mvPlanePoints) out[1] += (long long)c.points.size();
        for (const auto &c : f.mvBoundaryPoints) out[2] += (long long)c.points.size();
    }
    out[3] = (out[1] + out[2]) * (long long)sizeof(spx_host::PointT) + out[0] * 16;
}

// frame f: mnRealPlaneNum, mnPlaneNum, flags
int spx_seq_frame(void *h, int f, int out[3]) {
    const SequencePlanes *s = static_cast<SequencePlanes *>(h);
    if (f < 0 || f >= int(s->frames.size())) return 1;
    const PlaneFields &F = s->frames[size_t(f)];
    out[0] = F.mnRealPlaneNum; out[1] = F.mnPlaneNum; out[2] = int(F.flags);
    return 0;
}
// plane i of frame f: coefficients, sizes, width / height of both clouds, and the address of the 32-byte points
int spx_seq_plane(void *h, int f, int i, float coef[4], int sizes[6], const void **points, const void **boundary) {
    const SequencePlanes *s = static_cast<SequencePlanes *>(h);
    if (f < 0 || f >= int(s->frames.size())) return 1;
    const PlaneFields &F = s->frames[size_t(f)];
    if (i < 0 || i >= F.mnPlaneNum) return 1;
    for (int k = 0; k < 4; ++k) coef[k] = F.mvPlaneCoefficients[size_t(i)].at<float>(k);
    const auto &pc = F.mvPlanePoints[size_t(i)], &bc = F.mvBoundaryPoints[size_t(i)];
    sizes[0] = int(pc.points.size()); sizes[1] = int(pc.width); sizes[2] = int(pc.height);
    sizes[3] = int(bc.points.size()); sizes[4] = int(bc.width); sizes[5] = int(bc.height);
    *points = pc.points.data(); *boundary = bc.points.data();
    return 0;
}
// FNV-1a over (x, y, z, data[3], rgba) of every point of every cloud, in order: one number that pins all fields of a batch
unsigned long long spx_seq_hash(void *h) {
    const SequencePlanes *s = static_cast<SequencePlanes *>(h);
    uint64_t v = 1469598103934665603ull;
    for (const PlaneFields &F : s->frames) {
        v = fnv(v, &F.mnRealPlaneNum, 4); v = fnv(v, &F.mnPlaneNum, 4);
        for (int i = 0; i < F.mnPlaneNum; ++i) {
            for (int k = 0; k < 4; ++k) { const float c = F.mvPlaneCoefficients[size_t(i)].at<float>(k); v = fnv(v, &c, 4); }
            for (const auto *cl : {&F.mvPlanePoints[size_t(i)], &F.mvBoundaryPoints[size_t(i)]})
                for (const auto &p : cl->points) { v = fnv(v, &p, 20); }
        }
    }
    return v;
}

}  // extern "C"
