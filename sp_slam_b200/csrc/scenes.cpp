// scenes.cpp -- synthetic indoor depth renderer (input synthesis for tests and bench.py; not on the hot path).
// Renders exact ray/rectangle depth images (z-depth in metres, 0 = no geometry) for a pinhole camera, the
// workloads of SURVEY.md section 8(d): a box room with two interior boxes (640x480) and the same room plus a
// clutter field of small tilted patches (1280x720).  Frames are rendered in parallel on host threads.
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace {
struct Rect { double o[3], eu[3], ev[3], n[3], iu, iv; };

static inline double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

void render_one(const std::vector<Rect> &rects, const double *pose, double fx, double fy, double cx, double cy,
                int width, int height, float *out) {
    const double *R = pose;       // camera -> world rotation, row major
    const double *t = pose + 9;   // camera centre in world
    for (int v = 0; v < height; ++v) {
        for (int u = 0; u < width; ++u) {
            const double dc[3] = {(u - cx) / fx, (v - cy) / fy, 1.0};
            const double dw[3] = {R[0] * dc[0] + R[1] * dc[1] + R[2] * dc[2],
                                  R[3] * dc[0] + R[4] * dc[1] + R[5] * dc[2],
                                  R[6] * dc[0] + R[7] * dc[1] + R[8] * dc[2]};
            double best = 1e30;
            for (const Rect &q : rects) {
                const double denom = dot(q.n, dw);
                if (std::fabs(denom) < 1e-12) continue;
                const double oc[3] = {q.o[0] - t[0], q.o[1] - t[1], q.o[2] - t[2]};
                const double s = dot(q.n, oc) / denom;   // z-depth because dc.z == 1
                if (s <= 1e-6 || s >= best) continue;
                const double p[3] = {s * dw[0] - oc[0], s * dw[1] - oc[1], s * dw[2] - oc[2]};
                const double a = dot(p, q.eu) * q.iu, b = dot(p, q.ev) * q.iv;
                if (a < 0.0 || a > 1.0 || b < 0.0 || b > 1.0) continue;
                best = s;
            }
            out[size_t(v) * width + u] = best < 1e29 ? float(best) : 0.0f;
        }
    }
}
}  // namespace

extern "C" int spx_scene_render_batch(const double *rects9, int n_rects, const double *poses12, int n_frames,
                                      double fx, double fy, double cx, double cy, int width, int height,
                                      float *depth_out, int n_threads) {
    std::vector<Rect> rects(n_rects);
    for (int i = 0; i < n_rects; ++i) {
        Rect &q = rects[i];
        for (int k = 0; k < 3; ++k) { q.o[k] = rects9[i * 9 + k]; q.eu[k] = rects9[i * 9 + 3 + k]; q.ev[k] = rects9[i * 9 + 6 + k]; }
        q.n[0] = q.eu[1] * q.ev[2] - q.eu[2] * q.ev[1];
        q.n[1] = q.eu[2] * q.ev[0] - q.eu[0] * q.ev[2];
        q.n[2] = q.eu[0] * q.ev[1] - q.eu[1] * q.ev[0];
        q.iu = 1.0 / dot(q.eu, q.eu);
        q.iv = 1.0 / dot(q.ev, q.ev);
    }
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    for (int tid = 0; tid < n_threads; ++tid)
        pool.emplace_back([&, tid]() {
            for (int f = tid; f < n_frames; f += n_threads)
                render_one(rects, poses12 + size_t(f) * 12, fx, fy, cx, cy, width, height,
                           depth_out + size_t(f) * width * height);
        });
    for (auto &th : pool) th.join();
    return 0;
}
