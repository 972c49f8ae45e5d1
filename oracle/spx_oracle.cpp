/*
 * spx_oracle.cpp -- dependency-free CPU restatement of SP-SLAM's per-frame plane extraction.
 *
 * TEST INFRASTRUCTURE ONLY (see spx_oracle.h).  PARITY UNPINNED: PCL 1.8.0 is not available in this
 * environment and the reference has no tests for this path, so the oracle cannot be checked against the
 * real thing here.  It follows
 *   - /root/reference/src/Frame.cc:854-1144 for SP-SLAM's own logic (cited per function below), and
 *   - the published PCL 1.8.0 sources (pinned by /root/reference/build.sh:4) for the library calls:
 *       features/impl/integral_image_normal.hpp, features/impl/integral_image2D.hpp,
 *       segmentation/impl/organized_multi_plane_segmentation.hpp,
 *       segmentation/impl/organized_connected_component_segmentation.hpp,
 *       segmentation/plane_coefficient_comparator.h, segmentation/plane_refinement_comparator.h,
 *       common/impl/centroid.hpp, common/impl/eigen.hpp, features/normal_3d.h,
 *       sample_consensus/impl/{ransac,sac_model_line}.hpp, sample_consensus/sac_model.h,
 *       segmentation/impl/sac_segmentation.hpp, filters/impl/extract_indices.hpp.
 *
 * Under-determined choices (the reference's binary depends on its Eigen / Boost / compiler flags; each is a
 * last-ulp effect, far below the 1e-4 rad / 1e-4 m parity tolerances).  The CUDA path makes the SAME choices:
 *   E1  Eigen >= 3.3 semantics: `v /= s` is a true division (3.2 multiplied by 1/s).
 *   E2  fixed-size reductions are evaluated in Eigen 3.3's unrolled order:
 *         Vector3f dot / squaredNorm / sum : a0 + (a1 + a2)      (redux_novec_unroller)
 *         Vector3d squaredNorm            : (a0 + a1) + a2      (one Packet2d + scalar tail)
 *         Vector4f dot / squaredNorm      : (a0 + a2) + (a1 + a3)  (SSE predux)
 *   E3  no FMA contraction anywhere (build with -ffp-contract=off).
 *   E4  boost::uniform_int<>(0,INT_MAX) over mt19937 returns mt() >> 1.
 *   E5  pcl::PointXYZRGB default alpha = 255; computeMeanAndCovarianceMatrix sets centroid[3] = 1.
 *   E6  SampleConsensusModelLine::isSampleGood uses && (PCL 1.8.0) -- a pair is good only if x, y AND z differ.
 *   E7  Frame::IsBorderPoint reads imDepth.ptr<float>(j)[i] with no bounds check (src/Frame.cc:1039-1050).
 *       We define it by flat addressing on the continuous image, data[j*cols+i]; indices outside the buffer
 *       count as invalid samples, and a non-finite projection (PcZ == 0) makes the point "not border".
 *   E8  computeRoots' atan2 / cos / sin on floats resolve to the platform libm's atan2f / cosf / sinf, whose last
 *       bit is implementation defined; we take the correctly rounded value (evaluate in double, round to float).
 */
#include "spx_oracle.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

namespace {

typedef orc_point Pt;

static inline uint32_t pack_rgba(uint8_t r, uint8_t g, uint8_t b, uint8_t a = 255) {
    return (uint32_t(a) << 24) | (uint32_t(r) << 16) | (uint32_t(g) << 8) | uint32_t(b);
}
static inline float dot3f(float a0, float a1, float a2, float b0, float b1, float b2) {  // E2
    return a0 * b0 + (a1 * b1 + a2 * b2);
}
static inline float dot4f(const float *a, const float *b) {  // E2
    return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]);
}

// ------------------------------------------------------------------------------------------------
// pcl::computeRoots2 / computeRoots / eigen33 / computeCorrespondingEigenVector (common/impl/eigen.hpp)
// ------------------------------------------------------------------------------------------------
static void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    // Scalar (b * b - 4.0 * c): b*b is a float product, promoted to double for the subtraction
    float d = float(double(b * b) - 4.0 * double(c));
    if (d < 0.0f) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

static void compute_roots(const float m[9], float roots[3]) {
    // m row-major symmetric; m(i,j) = m[3*i+j]
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (std::fabs(c0) < FLT_EPSILON) {
        compute_roots2(c2, c1, roots);
    } else {
        const float s_inv3 = float(1.0 / 3.0);
        const float s_sqrt3 = std::sqrt(3.0f);
        float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.0f) a_over_3 = 0.0f;
        float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.0f) q = 0.0f;
        float rho = std::sqrt(-a_over_3);
        // E8: the three transcendental calls are evaluated in double and rounded to float
        float theta = float(std::atan2(double(std::sqrt(-q)), double(half_b))) * s_inv3;
        float cos_theta = float(std::cos(double(theta)));
        float sin_theta = float(std::sin(double(theta)));
        roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
        roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        if (roots[1] >= roots[2]) {
            std::swap(roots[1], roots[2]);
            if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        }
        if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
    }
}

static inline void cross3f(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static float scale_of(const float mat[9]) {
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    return scale;
}

// eigenvector of (S - lambda*I) by the largest row cross product; S is already scaled
static void eigvec_from_shifted(float S[9], float shift, float vec[3]) {
    S[0] -= shift; S[4] -= shift; S[8] -= shift;
    float v1[3], v2[3], v3[3];
    cross3f(&S[0], &S[3], v1);
    cross3f(&S[0], &S[6], v2);
    cross3f(&S[3], &S[6], v3);
    float l1 = dot3f(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
    float l2 = dot3f(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]);
    float l3 = dot3f(v3[0], v3[1], v3[2], v3[0], v3[1], v3[2]);
    const float *v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
    else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
    else { v = v3; l = l3; }
    float s = std::sqrt(l);
    vec[0] = v[0] / s; vec[1] = v[1] / s; vec[2] = v[2] / s;
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair
static void eigen33_smallest(const float mat[9], float &eigenvalue, float vec[3]) {
    float scale = scale_of(mat);
    float S[9];
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    float roots[3];
    compute_roots(S, roots);
    eigenvalue = roots[0] * scale;
    eigvec_from_shifted(S, roots[0], vec);
}

// pcl::eigen33(mat, evals)
static void eigen33_values(const float mat[9], float evals[3]) {
    float scale = scale_of(mat);
    float S[9];
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    compute_roots(S, evals);
    for (int i = 0; i < 3; ++i) evals[i] *= scale;
}

// pcl::computeCorrespondingEigenVector(mat, eigenvalue, eigenvector)
static void corresponding_eigvec(const float mat[9], float eigenvalue, float vec[3]) {
    float scale = scale_of(mat);
    float S[9];
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    eigvec_from_shifted(S, eigenvalue / scale, vec);
}

// ------------------------------------------------------------------------------------------------
// chamfer distance map of IntegralImageNormalEstimation::computeFeature (integral_image_normal.hpp)
// ------------------------------------------------------------------------------------------------
static void chamfer(const uint8_t *mask, int w, int h, float *dist, bool wrap = true) {
    const float far_ = 3.0e38f;   // (no-wrap alternative: the aliased neighbour does not exist)
    const int N = w * h;
    for (int i = 0; i < N; ++i) dist[i] = mask[i] == 0 ? 0.0f : float(w + h);
    // first pass; at ci = w-1, previous_row[ci+1] aliases current_row[0] (in-bounds row wrap)
    for (int ri = 1; ri < h; ++ri) {
        float *prev = dist + size_t(ri - 1) * w;
        float *cur = prev + w;
        for (int ci = 1; ci < w; ++ci) {
            const float upLeft = prev[ci - 1] + 1.4f;
            const float up = prev[ci] + 1.0f;
            const float upRight = (wrap || ci + 1 < w) ? prev[ci + 1] + 1.4f : far_;
            const float left = cur[ci - 1] + 1.0f;
            const float center = cur[ci];
            const float minValue = std::min(std::min(upLeft, up), std::min(left, upRight));
            if (minValue < center) cur[ci] = minValue;
        }
    }
    // second pass; at ci = 0, next_row[ci-1] aliases current_row[w-1]
    for (int ri = h - 2; ri >= 0; --ri) {
        float *cur = dist + size_t(ri) * w;
        float *next = cur + w;
        for (int ci = w - 2; ci >= 0; --ci) {
            const float lowerLeft = (wrap || ci >= 1) ? next[ci - 1] + 1.4f : far_;
            const float lower = next[ci] + 1.0f;
            const float lowerRight = next[ci + 1] + 1.4f;
            const float right = cur[ci + 1] + 1.0f;
            const float center = cur[ci];
            const float minValue = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
            if (minValue < center) cur[ci] = minValue;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// IntegralImage2D<float,3> first-order only, double accumulators (integral_image2D.hpp)
// ------------------------------------------------------------------------------------------------
struct Sat3 {
    int w = 0, h = 0;
    std::vector<double> I;        // (w+1)*(h+1)*3
    std::vector<unsigned> cnt;    // (w+1)*(h+1)
    bool exact = true;

    static inline double add_chk(double a, double b, bool &exact) {
        double s = a + b;
        double bb = s - a;
        double err = (a - (s - bb)) + (b - bb);
        if (err != 0.0) exact = false;
        return s;
    }
    void build(const float *data /* stride 4 */, int w_, int h_) {
        w = w_; h = h_;
        const int W1 = w + 1;
        I.assign(size_t(W1) * (h + 1) * 3, 0.0);
        cnt.assign(size_t(W1) * (h + 1), 0u);
        for (int r = 0; r < h; ++r) {
            const double *prev = &I[size_t(r) * W1 * 3];
            double *cur = &I[size_t(r + 1) * W1 * 3];
            const unsigned *cprev = &cnt[size_t(r) * W1];
            unsigned *ccur = &cnt[size_t(r + 1) * W1];
            for (int c = 0; c < w; ++c) {
                const float *v = data + (size_t(r) * w + c) * 4;
                ccur[c + 1] = cprev[c + 1] + ccur[c] - cprev[c];
                const bool fin = std::isfinite(v[0] + (v[1] + v[2]));
                for (int k = 0; k < 3; ++k) {
                    double t = add_chk(prev[(c + 1) * 3 + k], cur[c * 3 + k], exact);
                    t = add_chk(t, -prev[c * 3 + k], exact);
                    if (fin) t = add_chk(t, double(v[k]), exact);
                    cur[(c + 1) * 3 + k] = t;
                }
                if (fin) ++ccur[c + 1];
            }
        }
    }
    unsigned count(int x0, int y0, int kw, int kh) const {
        const int W1 = w + 1;
        const size_t ul = size_t(y0) * W1 + x0, ur = ul + kw, ll = size_t(y0 + kh) * W1 + x0, lr = ll + kw;
        return cnt[lr] + cnt[ul] - cnt[ur] - cnt[ll];
    }
    void sum(int x0, int y0, int kw, int kh, double out[3], bool &ex) const {
        const int W1 = w + 1;
        const size_t ul = size_t(y0) * W1 + x0, ur = ul + kw, ll = size_t(y0 + kh) * W1 + x0, lr = ll + kw;
        for (int k = 0; k < 3; ++k) {
            double t = add_chk(I[lr * 3 + k], I[ul * 3 + k], ex);
            t = add_chk(t, -I[ur * 3 + k], ex);
            t = add_chk(t, -I[ll * 3 + k], ex);
            out[k] = t;
        }
    }
};

// IntegralImage2D<float,3> with second-order computation (COVARIANCE_MATRIX method): x y z and their six products
// xx xy xz yy yz zz in double, one finite-element count; the recurrence and its operand order are PCL's
// (computeIntegralImages): cur[c+1] = prev[c+1] + cur[c] - prev[c], then += element when the point is finite.  These
// sums DO round, so the order is part of the result.
struct Sat9 {
    int w = 0, h = 0;
    std::vector<double> F, S;     // (w+1)*(h+1) * 3 / * 6
    std::vector<unsigned> cnt;
    void build(const Pt *pts, int w_, int h_, bool products_in_double) {
        w = w_; h = h_;
        const int W1 = w + 1;
        F.assign(size_t(W1) * (h + 1) * 3, 0.0); S.assign(size_t(W1) * (h + 1) * 6, 0.0);
        cnt.assign(size_t(W1) * (h + 1), 0u);
        for (int r = 0; r < h; ++r) {
            const size_t p0 = size_t(r) * W1, c0 = size_t(r + 1) * W1;
            for (int c = 0; c < w; ++c) {
                const Pt &q = pts[size_t(r) * w + c];
                const float e[3] = {q.x, q.y, q.z};
                for (int k = 0; k < 3; ++k) F[(c0 + c + 1) * 3 + k] = (F[(p0 + c + 1) * 3 + k] + F[(c0 + c) * 3 + k]) - F[(p0 + c) * 3 + k];
                for (int k = 0; k < 6; ++k) S[(c0 + c + 1) * 6 + k] = (S[(p0 + c + 1) * 6 + k] + S[(c0 + c) * 6 + k]) - S[(p0 + c) * 6 + k];
                cnt[c0 + c + 1] = cnt[p0 + c + 1] + cnt[c0 + c] - cnt[p0 + c];
                if (std::isfinite(e[0] + (e[1] + e[2]))) {
                    for (int k = 0; k < 3; ++k) F[(c0 + c + 1) * 3 + k] += double(e[k]);
                    int el = 0;
                    for (int a = 0; a < 3; ++a)
                        for (int b = a; b < 3; ++b, ++el)
                            S[(c0 + c + 1) * 6 + el] += products_in_double ? double(e[a]) * double(e[b]) : double(e[a] * e[b]);
                    ++cnt[c0 + c + 1];
                }
            }
        }
    }
    unsigned count(int x0, int y0, int kw, int kh) const {
        const int W1 = w + 1;
        const size_t ul = size_t(y0) * W1 + x0, ur = ul + kw, ll = size_t(y0 + kh) * W1 + x0, lr = ll + kw;
        return cnt[lr] + cnt[ul] - cnt[ur] - cnt[ll];
    }
    void sums(int x0, int y0, int kw, int kh, double fo[3], double so[6]) const {
        const int W1 = w + 1;
        const size_t ul = size_t(y0) * W1 + x0, ur = ul + kw, ll = size_t(y0 + kh) * W1 + x0, lr = ll + kw;
        for (int k = 0; k < 3; ++k) fo[k] = ((F[lr * 3 + k] + F[ul * 3 + k]) - F[ur * 3 + k]) - F[ll * 3 + k];
        for (int k = 0; k < 6; ++k) so[k] = ((S[lr * 6 + k] + S[ul * 6 + k]) - S[ur * 6 + k]) - S[ll * 6 + k];
    }
};

// ------------------------------------------------------------------------------------------------
// SACSegmentation<PointT> with SACMODEL_LINE / RANSAC (sac_segmentation.hpp, ransac.hpp, sac_model.h,
// sac_model_line.hpp); indices_ is the identity over the input cloud.
// ------------------------------------------------------------------------------------------------
struct SacLine {
    const Pt *pts; int n;
    std::vector<int> shuffled;
    std::mt19937 rng;          // boost::mt19937 is the same generator
    unsigned alt = 0;
    explicit SacLine(const Pt *p, int n_, unsigned alt_ = 0) : pts(p), n(n_), rng(12345u), alt(alt_) {
        shuffled.resize(n);
        for (int i = 0; i < n; ++i) shuffled[i] = i;
    }
    int rnd() { return (alt & ORC_ALT_RNG_MASK) ? int(rng() & 0x7fffffffu) : int(rng() >> 1); }   // E4
    void draw(int s[2]) {
        for (unsigned i = 0; i < 2; ++i)
            std::swap(shuffled[i], shuffled[i + (size_t(rnd()) % size_t(n - i))]);
        s[0] = shuffled[0]; s[1] = shuffled[1];
    }
    bool sample_good(const int s[2]) const {   // E6
        if (alt & ORC_ALT_SAMPLE_GOOD_OR) return pts[s[0]].x != pts[s[1]].x || pts[s[0]].y != pts[s[1]].y || pts[s[0]].z != pts[s[1]].z;
        return pts[s[0]].x != pts[s[1]].x && pts[s[0]].y != pts[s[1]].y && pts[s[0]].z != pts[s[1]].z;
    }
    // getSamples: false = "no samples could be selected"
    bool get_samples(int s[2]) {
        if (n < 2) return false;
        for (unsigned it = 0; it < 1000; ++it) {
            draw(s);
            if (sample_good(s)) return true;
        }
        return false;
    }
    static void normalize3(float *v) {   // Eigen normalize(): z = squaredNorm(); if (z>0) v /= sqrt(z)
        float z = dot3f(v[0], v[1], v[2], v[0], v[1], v[2]);
        if (z > 0.0f) { float s = std::sqrt(z); v[0] /= s; v[1] /= s; v[2] /= s; }
    }
    bool model_from(const int s[2], float c[6]) const {
        const Pt &a = pts[s[0]], &b = pts[s[1]];
        if (std::fabs(a.x - b.x) <= FLT_EPSILON && std::fabs(a.y - b.y) <= FLT_EPSILON &&
            std::fabs(a.z - b.z) <= FLT_EPSILON)
            return false;
        c[0] = a.x; c[1] = a.y; c[2] = a.z;
        c[3] = b.x - c[0]; c[4] = b.y - c[1]; c[5] = b.z - c[2];
        normalize3(c + 3);
        return true;
    }
    // the per-point test of countWithinDistance / selectWithinDistance
    static inline void prep_dir(const float c[6], float dir[4]) {
        dir[0] = c[3]; dir[1] = c[4]; dir[2] = c[5]; dir[3] = 0.0f;
        float z = dot4f(dir, dir);   // Vector4f::normalize()
        if (z > 0.0f) { float s = std::sqrt(z); dir[0] /= s; dir[1] /= s; dir[2] /= s; dir[3] /= s; }
    }
    static inline double sqr_dist(const float c[6], const float dir[4], const Pt &p) {
        // (line_pt - p).cross3(line_dir).squaredNorm(), Vector4f, w of the cross product is 0
        float a0 = c[0] - p.x, a1 = c[1] - p.y, a2 = c[2] - p.z;
        float x = a1 * dir[2] - a2 * dir[1];
        float y = a2 * dir[0] - a0 * dir[2];
        float z = a0 * dir[1] - a1 * dir[0];
        float sq = (x * x + z * z) + (y * y + 0.0f);   // E2
        return double(sq);
    }
    int count_within(const float c[6], double thr) const {
        const double sqr_thr = thr * thr;
        float dir[4]; prep_dir(c, dir);
        int nr = 0;
        for (int i = 0; i < n; ++i) if (sqr_dist(c, dir, pts[i]) < sqr_thr) ++nr;
        return nr;
    }
    void select_within(const float c[6], double thr, std::vector<int> &inl) const {
        const double sqr_thr = thr * thr;
        float dir[4]; prep_dir(c, dir);
        inl.clear();
        for (int i = 0; i < n; ++i) if (sqr_dist(c, dir, pts[i]) < sqr_thr) inl.push_back(i);
    }
    // optimizeModelCoefficients: centroid + eigenvector of the largest eigenvalue
    void optimize(const std::vector<int> &inl, const float c[6], float out[6]) const {
        if (inl.size() <= 2) { std::memcpy(out, c, 6 * sizeof(float)); return; }
        float cen[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < inl.size(); ++i) { cen[0] += pts[inl[i]].x; cen[1] += pts[inl[i]].y; cen[2] += pts[inl[i]].z; }
        const float cnt = float(inl.size());
        cen[0] /= cnt; cen[1] /= cnt; cen[2] /= cnt;   // E1
        float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < inl.size(); ++i) {
            float px = pts[inl[i]].x - cen[0], py = pts[inl[i]].y - cen[1], pz = pts[inl[i]].z - cen[2];
            cov[4] += py * py;
            cov[5] += py * pz;
            cov[8] += pz * pz;
            float qx = px * px, qy = py * px, qz = pz * px;   // pt *= pt.x()
            cov[0] += qx; cov[1] += qy; cov[2] += qz;
        }
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        out[0] = cen[0]; out[1] = cen[1]; out[2] = cen[2];
        float evals[3];
        eigen33_values(cov, evals);
        corresponding_eigvec(cov, evals[2], out + 3);
    }
    // returns false when segment() clears its outputs ("no solution found")
    bool segment(double threshold, int max_iter, float coef[6], std::vector<int> &inliers, int &iterations) {
        iterations = 0;
        int n_best = -INT_MAX;
        double k = 1.0;
        const double log_probability = std::log(1.0 - 0.99);
        const double one_over_indices = 1.0 / double(n);
        unsigned skipped = 0;
        const unsigned max_skip = unsigned(max_iter) * 10u;
        bool have_model = false;
        float best[6] = {0, 0, 0, 0, 0, 0};
        int sel[2];
        while (iterations < k && skipped < max_skip) {
            if (!get_samples(sel)) break;
            float mc[6];
            if (!model_from(sel, mc)) { ++skipped; continue; }
            int cnt = count_within(mc, threshold);
            if (cnt > n_best) {
                n_best = cnt;
                have_model = true;
                std::memcpy(best, mc, sizeof(best));
                double w = double(n_best) * one_over_indices;
                double p_no_outliers = 1.0 - std::pow(w, 2.0);
                p_no_outliers = std::max(DBL_EPSILON, p_no_outliers);
                p_no_outliers = std::min(1.0 - DBL_EPSILON, p_no_outliers);
                k = log_probability / std::log(p_no_outliers);
            }
            ++iterations;
            if (iterations > max_iter) break;
        }
        if (!have_model) { inliers.clear(); return false; }
        select_within(best, threshold, inliers);
        float refined[6];
        optimize(inliers, best, refined);
        std::memcpy(coef, refined, sizeof(refined));
        select_within(refined, threshold, inliers);
        return true;
    }
};

struct Model {
    float coef[4];
    float centroid[4];
    float cov[9];
    float curvature;
    uint32_t label;
    int n_segment;
    std::vector<int> inliers;
    std::vector<int> contour;
};

struct Plane {
    float coef[4];
    std::vector<Pt> points;
    std::vector<Pt> boundary;
    int src;
};

}  // namespace

struct orc_ctx {
    orc_config cfg;
    int rows = 0, cols = 0, w = 0, h = 0;
    const float *depth = nullptr;
    std::vector<Pt> cloud;
    std::vector<float> dist, nx, ny, nz, plane_d, curv;
    std::vector<uint32_t> labels_raw, labels;
    int n_label_lists = 0;
    bool sat_exact = true;
    std::vector<Model> models;
    std::vector<Plane> planes;   // mvPlanePoints / mvBoundaryPoints / mvPlaneCoefficients
    int n_real = 0, n_all = 0;
    std::vector<orc_line_rec> line_recs;
    double t_plane = 0, t_splane = 0;

    void back_project();
    void estimate_normals();
    void segment_and_refine();
    void post_filter();
    void generate_supposed();
    bool plane_not_seen(const float coef[4]) const;
    bool is_border_point(float X, float Y, float Z) const;
    bool line_in_range(float X, float Y, float Z) const;
    bool calculate_planes(int plane_i, const float line[6]);
};

// src/Frame.cc:855-874
void orc_ctx::back_project() {
    const int dis = cfg.cloud_dis;
    h = int(std::ceil(rows / float(dis)));
    w = int(std::ceil(cols / float(dis)));
    cloud.clear();
    cloud.reserve(size_t(w) * h);
    for (int m = 0; m < rows; m += dis) {
        for (int n = 0; n < cols; n += dis) {
            float d = depth[size_t(m) * cols + n];
            Pt p;
            p.z = d;
            p.x = (n - cfg.cx) * p.z / cfg.fx;
            p.y = (m - cfg.cy) * p.z / cfg.fy;
            p.rgba = pack_rgba(0, 0, 250);
            cloud.push_back(p);
        }
    }
}

// src/Frame.cc:878-885 -> IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT, BORDER_POLICY_IGNORE,
// no depth dependent smoothing, viewpoint 0): initAverage3DGradientMethod + computeFeature(Full)
void orc_ctx::estimate_normals() {
    const int N = w * h;
    const float qnan = std::numeric_limits<float>::quiet_NaN();
    nx.assign(N, qnan); ny.assign(N, qnan); nz.assign(N, qnan); curv.assign(N, qnan);
    dist.assign(N, 0.0f);
    if (w < 3 || h < 3) return;
    const bool cov_method = cfg.normal_method == 1;

    // initAverage3DGradientMethod
    std::vector<float> dfx(size_t(N) * 4, 0.0f), dfy(size_t(N) * 4, 0.0f);
    for (int r = 1; r < h - 1; ++r)
        for (int c = 1; c < w - 1; ++c) {
            const size_t i = size_t(r) * w + c;
            const Pt &R = cloud[i + 1], &L = cloud[i - 1], &D = cloud[i + w], &U = cloud[i - w];
            dfx[i * 4 + 0] = R.x - L.x; dfx[i * 4 + 1] = R.y - L.y; dfx[i * 4 + 2] = R.z - L.z;
            dfy[i * 4 + 0] = D.x - U.x; dfy[i * 4 + 1] = D.y - U.y; dfy[i * 4 + 2] = D.z - U.z;
        }
    Sat3 DX, DY;
    Sat9 XYZ;
    if (!cov_method) {
        DX.build(dfx.data(), w, h);
        DY.build(dfy.data(), w, h);
        sat_exact = DX.exact && DY.exact;
    } else {
        XYZ.build(cloud.data(), w, h, (cfg.alt & ORC_ALT_SO_DOUBLE) != 0);
        sat_exact = false;      // sums of products round: PCL's result depends on its summation order, which Sat9 follows
    }

    // computeFeature: depth change map
    std::vector<uint8_t> mask(N, 255);
    for (int ri = 0; ri < h - 1; ++ri)
        for (int ci = 0; ci < w - 1; ++ci) {
            const int index = ri * w + ci;
            const float depth_ = cloud[index].z, depthR = cloud[index + 1].z, depthD = cloud[index + w].z;
            const float thr = (cfg.max_depth_change_factor * (fabsf(depth_) + 1.0f) * 2.0f);
            if (std::fabs(depth_ - depthR) > thr || !std::isfinite(depth_) || !std::isfinite(depthR)) {
                mask[index] = 0; mask[index + 1] = 0;
            }
            if (std::fabs(depth_ - depthD) > thr || !std::isfinite(depth_) || !std::isfinite(depthD)) {
                mask[index] = 0; mask[index + w] = 0;
            }
        }
    chamfer(mask.data(), w, h, dist.data(), !(cfg.alt & ORC_ALT_CHAMFER_NO_WRAP));

    // computeFeatureFull, BORDER_POLICY_IGNORE
    const int border = int(cfg.normal_smoothing_size);
    const float smoothing_constant = cfg.normal_smoothing_size;
    for (int ri = border; ri < h - border; ++ri)
        for (int ci = border; ci < w - border; ++ci) {
            const int index = ri * w + ci;
            if (!std::isfinite(cloud[index].z)) continue;
            float smoothing = std::min(dist[index], smoothing_constant);
            if (!(smoothing > 2.0f)) continue;
            const int k = int(smoothing), half = k / 2;
            const int x0 = ci - half, y0 = ri - half;
            if (cov_method) {
                // computePointNormal, COVARIANCE_MATRIX
                const unsigned count = XYZ.count(x0, y0, k, k);
                if (count == 0) continue;
                double fo[3], so[6];
                XYZ.sums(x0, y0, k, k, fo, so);
                const float center[3] = {float(fo[0]), float(fo[1]), float(fo[2])};
                float C[9];
                C[0] = float(so[0]); C[1] = C[3] = float(so[1]); C[2] = C[6] = float(so[2]);
                C[4] = float(so[3]); C[5] = C[7] = float(so[4]); C[8] = float(so[5]);
                const float cntf = float(count);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) C[a * 3 + b] -= (center[a] * center[b]) / cntf;   // covariance_matrix -= (center * center^T) / count
                float eigen_value, ev[3];
                eigen33_smallest(C, eigen_value, ev);
                float fx_ = ev[0], fy_ = ev[1], fz_ = ev[2];
                float vpx = 0.0f - cloud[index].x, vpy = 0.0f - cloud[index].y, vpz = 0.0f - cloud[index].z;
                float cos_theta = (vpx * fx_ + vpy * fy_ + vpz * fz_);
                if (cos_theta < 0) { fx_ *= -1; fy_ *= -1; fz_ *= -1; }
                nx[index] = fx_; ny[index] = fy_; nz[index] = fz_;
                const float den = (cfg.alt & ORC_ALT_COV_TRACE) ? (C[0] + C[4] + C[8]) : (C[0] + C[2] + C[4]);
                curv[index] = eigen_value > 0.0f ? fabsf(eigen_value / den) : 0.0f;
                continue;
            }
            // computePointNormal, AVERAGE_3D_GRADIENT
            if (DX.count(x0, y0, k, k) == 0 || DY.count(x0, y0, k, k) == 0) continue;
            double gx[3], gy[3];
            DX.sum(x0, y0, k, k, gx, sat_exact);
            DY.sum(x0, y0, k, k, gy, sat_exact);
            double n[3] = { gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0] };
            double len = (n[0] * n[0] + n[1] * n[1]) + n[2] * n[2];   // E2
            if (len == 0.0) continue;
            double s = std::sqrt(len);
            float fx_ = float(n[0] / s), fy_ = float(n[1] / s), fz_ = float(n[2] / s);   // E1
            // flipNormalTowardsViewpoint(point, 0, 0, 0, nx, ny, nz)  (features/normal_3d.h)
            float vpx = 0.0f - cloud[index].x, vpy = 0.0f - cloud[index].y, vpz = 0.0f - cloud[index].z;
            float cos_theta = (vpx * fx_ + vpy * fy_ + vpz * fz_);
            if (cos_theta < 0) { fx_ *= -1; fy_ *= -1; fz_ *= -1; }
            nx[index] = fx_; ny[index] = fy_; nz[index] = fz_;
        }
}

// src/Frame.cc:898-905 -> OrganizedMultiPlaneSegmentation::segmentAndRefine
void orc_ctx::segment_and_refine() {
    const int N = w * h;
    models.clear();
    plane_d.assign(N, 0.0f);
    for (int i = 0; i < N; ++i)
        plane_d[i] = dot3f(cloud[i].x, cloud[i].y, cloud[i].z, nx[i], ny[i], nz[i]);

    // PlaneCoefficientComparator::compare (depth dependent distance threshold)
    const float ang_thr = cosf(float(0.017453 * cfg.angle_thr_deg));   // src/Frame.cc:900
    const float dist_thr = float(double(cfg.dist_thr));               // src/Frame.cc:901
    auto compare = [&](int i1, int i2) -> bool {
        float threshold = dist_thr;
        float z = cloud[i1].x * 0.0f + (cloud[i1].y * 0.0f + cloud[i1].z * 1.0f);   // vec.dot(z_axis_)
        threshold *= z * z;
        return (std::fabs(plane_d[i1] - plane_d[i2]) < threshold) &&
               (dot3f(nx[i1], ny[i1], nz[i1], nx[i2], ny[i2], nz[i2]) > ang_thr);
    };

    // OrganizedConnectedComponentSegmentation::segment
    const uint32_t invalid = 0xFFFFFFFFu;
    labels.assign(N, invalid);
    std::vector<unsigned> run_ids;
    auto find_root = [&](unsigned idx) { while (run_ids[idx] != idx) idx = run_ids[idx]; return idx; };
    unsigned clust_id = 0;
    if (N > 0 && std::isfinite(cloud[0].x)) { labels[0] = clust_id++; run_ids.push_back(labels[0]); }
    for (int c = 1; c < w; ++c) {
        if (!std::isfinite(cloud[c].x)) continue;
        else if (compare(c, c - 1)) labels[c] = labels[c - 1];
        else { labels[c] = clust_id++; run_ids.push_back(labels[c]); }
    }
    for (int r = 1; r < h; ++r) {
        const int cur = r * w, prev = cur - w;
        if (std::isfinite(cloud[cur].x)) {
            if (compare(cur, prev)) labels[cur] = labels[prev];
            else { labels[cur] = clust_id++; run_ids.push_back(labels[cur]); }
        }
        for (int c = 1; c < w; ++c) {
            if (!std::isfinite(cloud[cur + c].x)) continue;
            if (compare(cur + c, cur + c - 1)) labels[cur + c] = labels[cur + c - 1];
            if (compare(cur + c, prev + c)) {
                if (labels[cur + c] == invalid) labels[cur + c] = labels[prev + c];
                else if (labels[prev + c] != invalid) {
                    unsigned root1 = find_root(labels[cur + c]);
                    unsigned root2 = find_root(labels[prev + c]);
                    if (root1 < root2) run_ids[root2] = root1; else run_ids[root1] = root2;
                }
            }
            if (labels[cur + c] == invalid) { labels[cur + c] = clust_id++; run_ids.push_back(labels[cur + c]); }
        }
    }
    std::vector<unsigned> map(clust_id);
    unsigned max_id = 0;
    for (unsigned run = 0; run < run_ids.size(); ++run) {
        if (run_ids[run] == run) map[run] = max_id++;
        else map[run] = map[find_root(run)];
    }
    std::vector<std::vector<int>> label_indices(max_id + 1);
    for (int i = 0; i < N; ++i)
        if (labels[i] != invalid) {
            labels[i] = map[labels[i]];
            label_indices[labels[i]].push_back(i);
        }
    n_label_lists = int(label_indices.size());
    labels_raw = labels;

    // segment(): plane fit per label
    float vp[4] = {0, 0, 0, 0};
    for (size_t l = 0; l < label_indices.size(); ++l) {
        const std::vector<int> &idx = label_indices[l];
        if (!(unsigned(idx.size()) > unsigned(cfg.min_size))) continue;
        // computeMeanAndCovarianceMatrix (dense)
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i : idx) {
            const Pt &p = cloud[i];
            accu[0] += p.x * p.x; accu[1] += p.x * p.y; accu[2] += p.x * p.z;
            accu[3] += p.y * p.y; accu[4] += p.y * p.z; accu[5] += p.z * p.z;
            accu[6] += p.x; accu[7] += p.y; accu[8] += p.z;
        }
        const float cnt = float(idx.size());
        for (int k = 0; k < 9; ++k) accu[k] /= cnt;   // E1
        Model m;
        m.centroid[0] = accu[6]; m.centroid[1] = accu[7]; m.centroid[2] = accu[8]; m.centroid[3] = 1.0f;   // E5
        float *cov = m.cov;
        cov[0] = accu[0] - accu[6] * accu[6];
        cov[1] = accu[1] - accu[6] * accu[7];
        cov[2] = accu[2] - accu[6] * accu[8];
        cov[4] = accu[3] - accu[7] * accu[7];
        cov[5] = accu[4] - accu[7] * accu[8];
        cov[8] = accu[5] - accu[8] * accu[8];
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        float eigen_value, ev[3];
        eigen33_smallest(cov, eigen_value, ev);
        float pp[4] = {ev[0], ev[1], ev[2], 0.0f};
        pp[3] = -1 * dot4f(pp, m.centroid);
        if (cfg.alt & ORC_ALT_VP_RESET) vp[0] = vp[1] = vp[2] = vp[3] = 0.0f;
        for (int k = 0; k < 4; ++k) vp[k] -= m.centroid[k];   // vp is never reset between clusters
        float cos_theta = dot4f(vp, pp);
        if (cos_theta < 0) {
            for (int k = 0; k < 4; ++k) pp[k] *= -1;
            pp[3] = 0;
            pp[3] = -1 * dot4f(pp, m.centroid);
        }
        float curvature;
        float eig_sum = cov[0] + cov[4] + cov[8];
        if (eig_sum != 0) curvature = fabsf(eigen_value / eig_sum);
        else curvature = 0;
        if (curvature < 0.001 /* maximum_curvature_ is a double */) {
            std::memcpy(m.coef, pp, sizeof(pp));
            m.curvature = curvature;
            m.label = uint32_t(l);
            m.n_segment = int(idx.size());
            m.inliers = idx;
            models.push_back(std::move(m));
        }
    }

    // refine(): PlaneRefinementComparator, distance_threshold_ = 0.02f, not depth dependent
    std::vector<char> grow(label_indices.size(), 0);
    std::vector<int> label_to_model(label_indices.size(), 0);
    for (size_t i = 0; i < models.size(); ++i) {
        int model_label = int(labels[models[i].inliers[0]]);
        label_to_model[model_label] = int(i);
        grow[model_label] = 1;
    }
    const float refine_thr = 0.02f;
    auto rcompare = [&](int i1, int i2) -> bool {
        int cl = int(labels[i1]), nl = int(labels[i2]);
        if (!(grow[cl] && !grow[nl])) return false;
        const float *mc = models[label_to_model[cl]].coef;
        const Pt &pt = cloud[i2];
        double ptp = std::fabs(double(mc[0] * pt.x + mc[1] * pt.y + mc[2] * pt.z + mc[3]));
        return ptp < double(refine_thr);
    };
    auto claim = [&](int cl, int q) {
        labels[q] = uint32_t(cl);
        models[label_to_model[cl]].inliers.push_back(q);
    };
    if (!models.empty()) {
        for (int r = 0; r < h - 1; ++r) {
            const int cur = r * w, next = cur + w;
            for (int c = 0; c < w - 1; ++c) {
                int current_label = int(labels[cur + c]);
                int right_label = int(labels[cur + c + 1]);
                if (current_label < 0 || right_label < 0) continue;
                if (rcompare(cur + c, cur + c + 1)) claim(current_label, cur + c + 1);
                int lower_label = int(labels[next + c]);
                if (lower_label < 0) continue;
                if (rcompare(cur + c, next + c)) claim(current_label, next + c);
            }
        }
        for (int r = h - 1; r >= 1; --r) {
            const int cur = r * w, prev = cur - w;
            for (int c = w - 1; c >= 0; --c) {
                int current_label = int(labels[cur + c]);
                int left_label = int(labels[cur + c - 1]);   // c == 0: last pixel of the previous row
                if (current_label < 0 || left_label < 0) continue;
                if ((c > 0 || !(cfg.alt & ORC_ALT_REFINE_NO_WRAP)) && rcompare(cur + c, cur + c - 1)) claim(current_label, cur + c - 1);
                int upper_label = int(labels[prev + c]);
                if (upper_label < 0) continue;
                if (rcompare(cur + c, prev + c)) claim(current_label, prev + c);
            }
        }
    }

    // segmentAndRefine tail: findLabeledRegionBoundary from the LAST inlier index
    const int ddx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int ddy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    for (Model &m : models) {
        m.contour.clear();
        const int start = m.inliers.back();
        int curr_idx = start, curr_x = start % w, curr_y = start / w;
        const uint32_t label = labels[start];
        int direction = -1;
        for (int d = 0; d < 8; ++d) {
            int x = curr_x + ddx[d], y = curr_y + ddy[d];
            if (x >= 0 && x < w && y >= 0 && y < h && labels[y * w + x] != label) { direction = d; break; }
        }
        if (direction == -1) continue;
        m.contour.push_back(start);
        const size_t guard = size_t(8) * N + 8;   // the trace is finite; guard only against restatement bugs
        do {
            int nIdx = direction;
            for (int d = 1; d <= 8; ++d) {
                nIdx = (direction + d) & 7;
                int x = curr_x + ddx[nIdx], y = curr_y + ddy[nIdx];
                if (x >= 0 && x < w && y >= 0 && y < h && labels[y * w + x] == label) break;
            }
            direction = (nIdx + 4) & 7;
            curr_x += ddx[nIdx]; curr_y += ddy[nIdx];
            curr_idx = curr_y * w + curr_x;
            m.contour.push_back(curr_idx);
        } while (curr_idx != start && m.contour.size() < guard);
    }
}

// src/Frame.cc:1116-1144 (mvNotSeenPlaneCoefficients is never filled)
bool orc_ctx::plane_not_seen(const float coef[4]) const {
    for (const Plane &pl : planes) {
        const float *pM = pl.coef;
        float d = pM[3] - coef[3];
        float angle = pM[0] * coef[0] + pM[1] * coef[1] + pM[2] * coef[2];
        if (d > 0.2 || d < -0.2) continue;
        if (angle < 0.9397 && angle > -0.9397) continue;
        return false;
    }
    return true;
}

// src/Frame.cc:912-934
void orc_ctx::post_filter() {
    planes.clear();
    for (size_t i = 0; i < models.size(); ++i) {
        float coef[4] = {models[i].coef[0], models[i].coef[1], models[i].coef[2], models[i].coef[3]};
        if (coef[3] < 0) for (int k = 0; k < 4; ++k) coef[k] = -coef[k];
        if (!plane_not_seen(coef)) continue;
        Plane pl;
        std::memcpy(pl.coef, coef, sizeof(coef));
        pl.points.reserve(models[i].inliers.size());
        for (int idx : models[i].inliers) pl.points.push_back(cloud[idx]);     // ExtractIndices, list order
        pl.boundary.reserve(models[i].contour.size());
        for (int idx : models[i].contour) pl.boundary.push_back(cloud[idx]);   // regions[i].getContour()
        pl.src = int(i);
        planes.push_back(std::move(pl));
    }
    n_real = int(planes.size());
}

// src/Frame.cc:1058-1074
bool orc_ctx::line_in_range(float PcX, float PcY, float PcZ) const {
    if (PcZ < 0.0f) return false;
    const float invz = 1.0f / PcZ;
    const float u = cfg.fx * PcX * invz + cfg.cx;
    const float v = cfg.fy * PcY * invz + cfg.cy;
    if (u < (cfg.min_x + 50) || u > (cfg.max_x - 50)) return false;
    if (v < (cfg.min_y + 50) || v > (cfg.max_y - 50)) return false;
    return true;
}

// src/Frame.cc:1026-1056, with E7
bool orc_ctx::is_border_point(float PcX, float PcY, float PcZ) const {
    if (PcZ < 0.0f) return false;
    const float invz = 1.0f / PcZ;
    const float u = cfg.fx * PcX * invz + cfg.cx;
    const float v = cfg.fy * PcY * invz + cfg.cy;
    // not finite (a point at depth 0: u = fx * (+-0) * inf + cx = NaN): with v NaN or -inf the loop over j below never runs, with
    // u NaN or -inf the loop over i never does; res / num = NaN, the final test is false and the reference returns true.  Only
    // +inf, where the reference walks out of the image buffer, is answered false (E7)
    const float pinf = std::numeric_limits<float>::infinity();
    if (!std::isfinite(v)) return v != pinf;
    if (!std::isfinite(u)) return u != pinf;
    int num = 0, nan = 0;
    float res = 0;
    int b = 10;
    const long total = long(rows) * cols;
    for (int j = int(v - b); j < v + b; ++j) {
        for (int i = int(u - b); i < u + b; ++i) {
            const long f = long(j) * cols + i;
            const bool inside = f >= 0 && f < total;
            if (inside && depth[f] > 0.05) {
                res += depth[f];
                num++;
            } else {
                nan++;
                if (nan > b * b) return false;
            }
        }
    }
    if (PcZ - res / num > 0.1) return false;
    return true;
}

// src/Frame.cc:1077-1114
bool orc_ctx::calculate_planes(int plane_i, const float l[6]) {
    const float *p = planes[plane_i].coef;
    float a, b, c, d;
    a = p[1] * l[5] - p[2] * l[4];
    b = p[2] * l[3] - p[0] * l[5];
    c = p[0] * l[4] - p[1] * l[3];
    d = a * l[0] + b * l[1] + c * l[2];
    float v = std::sqrt(a * a + b * b + c * c);
    float coef[4] = {a / v, b / v, c / v, -d / v};
    if (coef[3] < 0) for (int k = 0; k < 4; ++k) coef[k] = -coef[k];
    if (plane_not_seen(coef)) {
        Plane pl;
        std::memcpy(pl.coef, coef, sizeof(coef));
        Pt q;
        q.rgba = pack_rgba(0, 255, 0);
        for (float i = -0.25; i < 0.25;) {
            for (float j = -0.25; j < 0.25;) {
                q.x = l[0] + i * l[3] + j * p[0];
                q.y = l[1] + i * l[4] + j * p[1];
                q.z = (coef[0] * q.x + coef[1] * q.y + coef[3]) / (-coef[2]);
                pl.points.push_back(q);
                j = j + 0.01;
            }
            i = i + 0.01;
        }
        pl.src = plane_i;
        planes.push_back(std::move(pl));
        return true;
    }
    return false;
}

// src/Frame.cc:938-1011
void orc_ctx::generate_supposed() {
    line_recs.clear();
    const double lineRatio = cfg.line_ratio;
    const double thr = double(cfg.line_dist_thr);
    std::vector<Pt> bound, tmp, linePts;
    std::vector<int> inl;
    const int iend = int(planes.size()) - 1;
    for (int i = iend; i >= 0; --i) {
        bound = planes[i].boundary;
        const int boundSize = int(bound.size());
        if (boundSize < 50) {
            if (boundSize == 0) {
                // GenerateBoundaryPoints: every 20th inlier, default-constructed colour
                for (size_t j = 0; j < planes[i].points.size(); j += 20) {
                    Pt q = planes[i].points[j];
                    q.rgba = pack_rgba(0, 0, 0);
                    planes[i].boundary.push_back(q);
                }
            }
            continue;
        }
        for (int j = 0; j < 4; ++j) {
            orc_line_rec rec;
            std::memset(&rec, 0, sizeof(rec));
            rec.plane = i; rec.round = j; rec.n_points = int(bound.size());
            float coef[6] = {0, 0, 0, 0, 0, 0};
            bool ok = false;
            int iters = 0;
            inl.clear();
            if (!bound.empty()) {   // PCLBase::initCompute fails on an empty cloud -> outputs cleared
                SacLine sac(bound.data(), int(bound.size()), cfg.alt);
                ok = sac.segment(thr, cfg.ransac_max_iter, coef, inl, iters);
            }
            if (!ok) inl.clear();
            rec.iterations = iters; rec.n_inliers = int(inl.size());
            std::memcpy(rec.coef, coef, sizeof(coef));
            if (double(inl.size()) < lineRatio * boundSize) { line_recs.push_back(rec); break; }
            linePts.clear();
            for (int k : inl) linePts.push_back(bound[k]);
            rec.in_range = line_in_range(coef[0], coef[1], coef[2]) ? 1 : 0;
            bool border = false;
            if (rec.in_range) {
                // IsBorderLine (src/Frame.cc:1013-1024)
                int s = int(linePts.size()), res = 0;
                border = true;
                for (const Pt &q : linePts) {
                    if (!is_border_point(q.x, q.y, q.z)) res++;
                    if (res > s / 4) { border = false; break; }
                }
                rec.is_border = border ? 1 : 0;
            }
            if (rec.in_range && border) {
                if (calculate_planes(i, coef)) {
                    rec.emitted = 1;
                    for (Pt &q : linePts) q.rgba = pack_rgba(255, 0, 0);
                    Plane &sp = planes.back();
                    sp.points.insert(sp.points.end(), linePts.begin(), linePts.end());
                    sp.boundary = linePts;   // mvBoundaryPoints.push_back(*linePoints)
                }
            }
            line_recs.push_back(rec);
            // extract.setNegative(true): sorted set difference (inl is ascending)
            tmp.clear();
            size_t k = 0;
            for (int q = 0; q < int(bound.size()); ++q) {
                if (k < inl.size() && inl[k] == q) { ++k; continue; }
                tmp.push_back(bound[q]);
            }
            bound.swap(tmp);
        }
    }
}

// ------------------------------------------------------------------------------------------------
extern "C" {

void orc_default_config(orc_config *c) {
    c->cloud_dis = 3; c->min_size = 500; c->angle_thr_deg = 3.0f; c->dist_thr = 0.05f;
    c->line_ratio = 0.2; c->line_dist_thr = 0.01f;
    c->fx = 517.306408f; c->fy = 516.469215f; c->cx = 318.643040f; c->cy = 255.313989f;
    c->min_x = 0.0f; c->max_x = 640.0f; c->min_y = 0.0f; c->max_y = 480.0f;
    c->max_depth_change_factor = 0.05f; c->normal_smoothing_size = 10.0f;
    c->ransac_max_iter = 1000; c->enable_supposed = 1;
    c->alt = 0;
    c->normal_method = 0;
}

orc_ctx *orc_create(const orc_config *cfg) { orc_ctx *c = new orc_ctx(); c->cfg = *cfg; return c; }
void orc_destroy(orc_ctx *c) { delete c; }

int orc_run(orc_ctx *c, const float *depth, int rows, int cols, const float *normals_in) {
    using clk = std::chrono::steady_clock;
    c->depth = depth; c->rows = rows; c->cols = cols;
    auto t1 = clk::now();
    c->back_project();
    const int N = c->w * c->h;
    if (normals_in) {
        c->dist.assign(N, 0.0f);
        c->nx.assign(normals_in, normals_in + N);
        c->ny.assign(normals_in + N, normals_in + 2 * size_t(N));
        c->nz.assign(normals_in + 2 * size_t(N), normals_in + 3 * size_t(N));
        c->sat_exact = true;
    } else {
        c->estimate_normals();
    }
    c->segment_and_refine();
    c->post_filter();
    auto t2 = clk::now();
    c->line_recs.clear();
    if (c->cfg.enable_supposed) c->generate_supposed();
    auto t3 = clk::now();
    c->n_all = int(c->planes.size());
    c->t_plane = std::chrono::duration<double>(t2 - t1).count();
    c->t_splane = std::chrono::duration<double>(t3 - t2).count();
    c->depth = nullptr;
    return 0;
}

void orc_dims(const orc_ctx *c, int *w, int *h) { *w = c->w; *h = c->h; }
void orc_get_cloud(const orc_ctx *c, float *x, float *y, float *z) {
    for (size_t i = 0; i < c->cloud.size(); ++i) { x[i] = c->cloud[i].x; y[i] = c->cloud[i].y; z[i] = c->cloud[i].z; }
}
void orc_get_distance_map(const orc_ctx *c, float *d) { std::memcpy(d, c->dist.data(), c->dist.size() * sizeof(float)); }
void orc_get_normals(const orc_ctx *c, float *x, float *y, float *z) {
    std::memcpy(x, c->nx.data(), c->nx.size() * 4); std::memcpy(y, c->ny.data(), c->ny.size() * 4);
    std::memcpy(z, c->nz.data(), c->nz.size() * 4);
}
void orc_get_plane_d(const orc_ctx *c, float *d) { std::memcpy(d, c->plane_d.data(), c->plane_d.size() * 4); }
void orc_get_curvature(const orc_ctx *c, float *d) { std::memcpy(d, c->curv.data(), c->curv.size() * 4); }
int orc_get_labels_raw(const orc_ctx *c, uint32_t *l) {
    std::memcpy(l, c->labels_raw.data(), c->labels_raw.size() * 4); return c->n_label_lists;
}
void orc_get_labels_refined(const orc_ctx *c, uint32_t *l) { std::memcpy(l, c->labels.data(), c->labels.size() * 4); }
int orc_num_models(const orc_ctx *c) { return int(c->models.size()); }
void orc_get_model(const orc_ctx *c, int i, float coef[4], float centroid[3], float cov[9], float *curv,
                   uint32_t *label, int *n_seg, int *n_ref, int *n_contour) {
    const Model &m = c->models[i];
    std::memcpy(coef, m.coef, 16); std::memcpy(centroid, m.centroid, 12); std::memcpy(cov, m.cov, 36);
    *curv = m.curvature; *label = m.label; *n_seg = m.n_segment; *n_ref = int(m.inliers.size());
    *n_contour = int(m.contour.size());
}
void orc_get_model_inliers(const orc_ctx *c, int i, int32_t *idx) {
    std::memcpy(idx, c->models[i].inliers.data(), c->models[i].inliers.size() * 4);
}
void orc_get_model_contour(const orc_ctx *c, int i, int32_t *idx) {
    std::memcpy(idx, c->models[i].contour.data(), c->models[i].contour.size() * 4);
}
int orc_sat_exact(const orc_ctx *c) { return c->sat_exact ? 1 : 0; }
int orc_num_real_planes(const orc_ctx *c) { return c->n_real; }
int orc_num_planes(const orc_ctx *c) { return c->n_all; }
void orc_get_plane(const orc_ctx *c, int i, float coef[4], int *np, int *nb, int *src) {
    const Plane &p = c->planes[i];
    std::memcpy(coef, p.coef, 16); *np = int(p.points.size()); *nb = int(p.boundary.size()); *src = p.src;
}
void orc_get_plane_points(const orc_ctx *c, int i, orc_point *pts) {
    std::memcpy(pts, c->planes[i].points.data(), c->planes[i].points.size() * sizeof(orc_point));
}
void orc_get_plane_boundary(const orc_ctx *c, int i, orc_point *pts) {
    std::memcpy(pts, c->planes[i].boundary.data(), c->planes[i].boundary.size() * sizeof(orc_point));
}
int orc_num_line_recs(const orc_ctx *c) { return int(c->line_recs.size()); }
void orc_get_line_recs(const orc_ctx *c, orc_line_rec *out) {
    std::memcpy(out, c->line_recs.data(), c->line_recs.size() * sizeof(orc_line_rec));
}
void orc_get_times(const orc_ctx *c, double *a, double *b) { *a = c->t_plane; *b = c->t_splane; }

int orc_run_batch(const orc_config *cfg, const float *depth, int n_frames, int rows, int cols, int n_threads,
                  int32_t *n_real, int32_t *n_planes, double *t_plane_sum, double *t_splane_sum) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
    std::mutex mu;
    double tp = 0, ts = 0;
    std::vector<std::thread> pool;
    for (int tid = 0; tid < n_threads; ++tid)
        pool.emplace_back([&, tid]() {
            orc_ctx *c = orc_create(cfg);
            double a = 0, b = 0;
            for (int f = tid; f < n_frames; f += n_threads) {
                orc_run(c, depth + size_t(f) * rows * cols, rows, cols, nullptr);
                if (n_real) n_real[f] = c->n_real;
                if (n_planes) n_planes[f] = c->n_all;
                a += c->t_plane; b += c->t_splane;
            }
            orc_destroy(c);
            std::lock_guard<std::mutex> lk(mu);
            tp += a; ts += b;
        });
    for (auto &t : pool) t.join();
    if (t_plane_sum) *t_plane_sum = tp;
    if (t_splane_sum) *t_splane_sum = ts;
    return 0;
}

// exactness of the integral-image arithmetic for a batch of frames (cloud + normal estimation only), frame-parallel
int orc_sat_exact_batch(const orc_config *cfg, const float *depth, int n_frames, int rows, int cols, int n_threads, uint8_t *exact) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
    std::vector<std::thread> pool;
    for (int tid = 0; tid < n_threads; ++tid)
        pool.emplace_back([&, tid]() {
            orc_ctx *c = orc_create(cfg);
            for (int f = tid; f < n_frames; f += n_threads) {
                c->depth = depth + size_t(f) * rows * cols; c->rows = rows; c->cols = cols;
                c->back_project();
                c->estimate_normals();
                exact[f] = c->sat_exact ? 1 : 0;
            }
            orc_destroy(c);
        });
    for (auto &t : pool) t.join();
    return 0;
}

void orc_chamfer(const uint8_t *mask, int w, int h, float *dist) { chamfer(mask, w, h, dist); }
void orc_chamfer_alt(const uint8_t *mask, int w, int h, float *dist, uint32_t alt) { chamfer(mask, w, h, dist, !(alt & ORC_ALT_CHAMFER_NO_WRAP)); }
int orc_is_border_point(const orc_config *cfg, const float *depth, int rows, int cols, float x, float y, float z) {
    orc_ctx c;
    c.cfg = *cfg; c.depth = depth; c.rows = rows; c.cols = cols;
    return c.is_border_point(x, y, z) ? 1 : 0;
}
void orc_eigen33_smallest(const float cov[9], float *ev, float vec[3]) { eigen33_smallest(cov, *ev, vec); }
void orc_eigen33_largest(const float cov[9], float evals[3], float vec[3]) {
    eigen33_values(cov, evals); corresponding_eigvec(cov, evals[2], vec);
}
int orc_sac_line(const orc_point *pts, int n, double thr, int max_iter, float coef[6], int32_t *inliers, int *iterations) {
    return orc_sac_line_alt(pts, n, thr, max_iter, coef, inliers, iterations, 0u);
}
int orc_sac_line_alt(const orc_point *pts, int n, double thr, int max_iter, float coef[6], int32_t *inliers, int *iterations, uint32_t alt) {
    std::vector<int> inl; int it = 0;
    for (int k = 0; k < 6; ++k) coef[k] = 0;
    bool ok = false;
    if (n > 0) { SacLine sac(pts, n, alt); ok = sac.segment(thr, max_iter, coef, inl, it); }
    if (!ok) inl.clear();
    if (iterations) *iterations = it;
    if (inliers) std::memcpy(inliers, inl.data(), inl.size() * 4);
    return int(inl.size());
}
void orc_ransac_draws(int n, int n_draws, int32_t *pairs) {
    std::vector<Pt> dummy(n);
    SacLine sac(dummy.data(), n);
    for (int i = 0; i < n_draws; ++i) { int s[2]; sac.draw(s); pairs[2 * i] = s[0]; pairs[2 * i + 1] = s[1]; }
}

// ---------------------------------------------------------------------------------------------------------------
// N4: pcl::VoxelGrid<pcl::PointXYZRGB>::applyFilter, PCL 1.8.0 filters/impl/voxel_grid.hpp, as SP-SLAM
// configures it: setLeafSize(l, l, l), everything else default (downsample_all_data_ = true, min_points_per_voxel_ = 0,
// no filter field).  Call sites: src/MapDrawer.cc:91-92,115-116, src/PointCloudMapping.cc:117-118,172-173 and the dead
// path src/Frame.cc:810-814.
//   1. getMinMax3D over the finite points; 2. if the voxel index range overflows an int the filter warns and returns
//   the input unchanged; 3. min_b = int(floor(min * inverse_leaf)), div_b = max_b - min_b + 1; 4. every finite point gets
//   idx = ijk . (1, div_b.x, div_b.x * div_b.y) with ijk = int(floor(p * inverse_leaf) - float(min_b));
//   5. std::sort of (idx, point) by idx -- NOT stable: the order of points inside a voxel is libstdc++'s introsort order;
//   6. one output point per run of equal idx, in ascending idx order: CentroidPoint<PointXYZRGB> = fp32 running sums of
//   x, y, z and of r, g, b, a (as floats) in that order, divided by float(n); colours truncated to uint32.
// Returns the number of output points (out must hold n).
// ---------------------------------------------------------------------------------------------------------------
int orc_voxel_grid(const orc_point *pts, int n, const float leaf[3], orc_point *out, int32_t *out_idx) {
    if (n <= 0) return 0;
    const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; ++i) {
        const float v[3] = {pts[i].x, pts[i].y, pts[i].z};
        if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], v[k]); mx[k] = std::max(mx[k], v[k]); }
    }
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv[0]) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv[1]) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv[2]) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
        for (int i = 0; i < n; ++i) { out[i] = pts[i]; if (out_idx) out_idx[i] = -1; }   // "Leaf size is too small": output = *input_
        return n;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int k = 0; k < 3; ++k) {
        min_b[k] = static_cast<int>(std::floor(mn[k] * inv[k]));
        max_b[k] = static_cast<int>(std::floor(mx[k] * inv[k]));
        div_b[k] = max_b[k] - min_b[k] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct Item { unsigned idx; unsigned pt; bool operator<(const Item &o) const { return idx < o.idx; } };   // cloud_point_index_idx
    std::vector<Item> iv;
    iv.reserve(size_t(n));
    for (int i = 0; i < n; ++i) {
        const float v[3] = {pts[i].x, pts[i].y, pts[i].z};
        if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
        const int i0 = static_cast<int>(std::floor(v[0] * inv[0]) - static_cast<float>(min_b[0]));
        const int i1 = static_cast<int>(std::floor(v[1] * inv[1]) - static_cast<float>(min_b[1]));
        const int i2 = static_cast<int>(std::floor(v[2] * inv[2]) - static_cast<float>(min_b[2]));
        iv.push_back({static_cast<unsigned>(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), static_cast<unsigned>(i)});
    }
    std::sort(iv.begin(), iv.end(), std::less<Item>());
    int n_out = 0;
    for (size_t a = 0; a < iv.size();) {
        size_t b = a + 1;
        while (b < iv.size() && iv[b].idx == iv[a].idx) ++b;
        float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
        for (size_t k = a; k < b; ++k) {
            const orc_point &p = pts[iv[k].pt];
            sx += p.x; sy += p.y; sz += p.z;
            sr += float((p.rgba >> 16) & 255u); sg += float((p.rgba >> 8) & 255u); sb += float(p.rgba & 255u); sa += float(p.rgba >> 24);
        }
        const float fn = float(b - a);
        orc_point o;
        o.x = sx / fn; o.y = sy / fn; o.z = sz / fn;
        o.rgba = (static_cast<uint32_t>(sa / fn) << 24) | (static_cast<uint32_t>(sr / fn) << 16) | (static_cast<uint32_t>(sg / fn) << 8) |
                 static_cast<uint32_t>(sb / fn);
        if (out_idx) out_idx[n_out] = int32_t(iv[a].idx);
        out[n_out++] = o;
        a = b;
    }
    return n_out;
}

// ---------------------------------------------------------------------------------------------------------------
// N1: Map::AssociatePlanesByBoundary for the planes of one frame (src/Map.cc:196-283) and
// Map::PointDistanceFromPlane (src/Map.cc:345-361).  plane_w: world-frame coefficients of the frame's planes
// (Frame::ComputePlaneWorldCoeff, src/Frame.cc:1146-1150, is the caller's); map planes are visited in the order given
// (the reference iterates a std::set<MapPlane*>, i.e. in pointer order), first `n_seen` = mspMapPlanes, the rest =
// mspNotSeenMapPlanes.  Outputs per frame plane: index of the associated / vertical / parallel map plane or -1.
// ---------------------------------------------------------------------------------------------------------------
static double point_distance_from_plane(const float pl[4], const orc_point *b, int n) {
    double res = 100;
    for (int k = 0; k < n; ++k) {
        const float e = pl[0] * b[k].x + pl[1] * b[k].y + pl[2] * b[k].z + pl[3];
        const double dis = std::abs(e);
        if (dis < res) res = dis;
    }
    return res;
}

void orc_associate_planes(const float *plane_w, int n_planes, const float *map_w, const orc_point *map_bnd, const int64_t *map_off,
                          int n_seen, int n_map, float dis_th, float ang_th, float ver_th, float par_th,
                          int32_t *assoc, int32_t *vertical, int32_t *parallel, float *assoc_dist) {
    for (int i = 0; i < n_planes; ++i) {
        const float *pM = plane_w + 4 * i;
        assoc[i] = vertical[i] = parallel[i] = -1;
        float ldTh = dis_th, lverTh = ver_th, lparTh = par_th;
        for (int j = 0; j < n_seen; ++j) {
            const float *pW = map_w + 4 * j;
            const float angle = pM[0] * pW[0] + pM[1] * pW[1] + pM[2] * pW[2];
            if (angle > ang_th || angle < -ang_th) {
                const double dis = point_distance_from_plane(pM, map_bnd + map_off[j], int(map_off[j + 1] - map_off[j]));
                if (dis < ldTh) { ldTh = float(dis); assoc[i] = j; continue; }
            }
            if (angle < lverTh && angle > -lverTh) { lverTh = std::abs(angle); vertical[i] = j; continue; }
            if (angle > lparTh || angle < -lparTh) { lparTh = std::abs(angle); parallel[i] = j; }
        }
        if (ldTh == dis_th) {
            for (int j = n_seen; j < n_map; ++j) {
                const float *pW = map_w + 4 * j;
                const float angle = pM[0] * pW[0] + pM[1] * pW[1] + pM[2] * pW[2];
                if (angle > ang_th || angle < -ang_th) {
                    const double dis = point_distance_from_plane(pM, map_bnd + map_off[j], int(map_off[j + 1] - map_off[j]));
                    if (dis < ldTh) { ldTh = float(dis); assoc[i] = j; }
                }
            }
        }
        if (assoc_dist) assoc_dist[i] = ldTh;
    }
}

// MapPlane::MapPlane / MapPlane::UpdateBoundary (src/MapPlane.cc:25-31,144-147): pcl::transformPointCloud(cloud, out,
// T.inverse().matrix()) with a Matrix4d -- PCL 1.8.0 common/impl/transforms.hpp, dense branch, Scalar = double:
//   out.x = static_cast<float>(t(0,0) * x + t(0,1) * y + t(0,2) * z + t(0,3)), rows 1 and 2 alike; every other field copied.
// m: row-major 4x4.
void orc_transform_cloud(const orc_point *pts, int n, const double m[16], orc_point *out) {
    for (int i = 0; i < n; ++i) {
        const double x = pts[i].x, y = pts[i].y, z = pts[i].z;
        orc_point o = pts[i];
        o.x = static_cast<float>(m[0] * x + m[1] * y + m[2] * z + m[3]);
        o.y = static_cast<float>(m[4] * x + m[5] * y + m[6] * z + m[7]);
        o.z = static_cast<float>(m[8] * x + m[9] * y + m[10] * z + m[11]);
        out[i] = o;
    }
}

}  // extern "C"
