// PlanePoseOptimizer.h -- SURVEY 8(f) row N3: the plane part of ORB_SLAM2::Optimizer::PoseOptimization
// (/root/reference/src/Optimizer.cc:519-1160): a pose-only Levenberg-Marquardt over the frame's SE3 pose with the unary
// plane edges of g2oAddition --
//   EdgePlane          (g2oAddition/EdgePlane.h:25-36)          e = (Tcw * plane_w).ominus(measurement)       3 residuals
//   EdgeParallelPlane  (g2oAddition/EdgeParallelPlane.h:24-36)  e = (Tcw * plane_w).ominus_par(measurement)   2 residuals
//   EdgeVerticalPlane  (g2oAddition/EdgeVerticalPlane.h:24-36)  e = (Tcw * plane_w).ominus_ver(measurement)   2 residuals
// with Plane3D's parametrisation (azimuth, elevation, distance; g2oAddition/Plane3D.h:46-125), Huber kernels, the outlier
// rounds of PoseOptimization (4 x optimize(10), chi2 test against Plane.Chi / Plane.VPChi, kernels dropped after the
// third round, `if (optimizer.edges().size() < 10) break`), g2o's numeric Jacobians (BaseBinaryEdge::linearizeOplus,
// central differences, delta 1e-9) and g2o's OptimizationAlgorithmLevenberg (tau 1e-5, the rho / scale acceptance test,
// lambda *= max(1/3, min(2/3, 1 - (2 rho - 1)^3)) on success, lambda *= ni, ni *= 2 on failure, 10 trials).
//
// This is HOST code, as in the reference (a 6x6 dense system a few times per frame: nothing data-parallel); it closes
// BASELINE configs[4] -- extraction (GPU) -> association (GPU) -> pose optimisation -> boundary update (GPU) -- for
// the plane landmarks.  The ORB point edges of the same function (EdgeSE3ProjectXYZOnlyPose, EdgeStereoSE3ProjectXYZ-
// OnlyPose) are outside this repository's scope; a caller that has them adds their
// Hessian blocks through `extra` below.
//
// PARITY UNPINNED: g2o is vendored in the reference as sources only and needs Eigen, which is not available here, so the
// reference optimiser cannot be run; the restatement is double precision throughout and is tested against an
// independent numpy restatement of the residuals and a scipy least-squares solve (tolerances in tests/test_pose_planes.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace spx_host {

struct Vec3 { double x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(Vec3 a) { return std::sqrt(dot(a, a)); }

struct Mat3 {
    double m[3][3];
    static Mat3 identity() { Mat3 r{}; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0; return r; }
};
inline Mat3 operator*(const Mat3 &a, const Mat3 &b) {
    Mat3 r{};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
inline Vec3 operator*(const Mat3 &a, Vec3 v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline Mat3 transpose(const Mat3 &a) { Mat3 r{}; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
inline Mat3 add(const Mat3 &a, const Mat3 &b, double sb = 1.0) { Mat3 r{}; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + sb * b.m[i][j]; return r; }
inline Mat3 skew(Vec3 w) { Mat3 r{}; r.m[0][1] = -w.z; r.m[0][2] = w.y; r.m[1][0] = w.z; r.m[1][2] = -w.x; r.m[2][0] = -w.y; r.m[2][1] = w.x; return r; }
// Eigen::AngleAxisd(angle, unit axis).toRotationMatrix()
inline Mat3 angle_axis(double angle, Vec3 a) {
    const double c = std::cos(angle), s = std::sin(angle), t = 1.0 - c;
    Mat3 r{};
    r.m[0][0] = t * a.x * a.x + c;       r.m[0][1] = t * a.x * a.y - s * a.z; r.m[0][2] = t * a.x * a.z + s * a.y;
    r.m[1][0] = t * a.x * a.y + s * a.z; r.m[1][1] = t * a.y * a.y + c;       r.m[1][2] = t * a.y * a.z - s * a.x;
    r.m[2][0] = t * a.x * a.z - s * a.y; r.m[2][1] = t * a.y * a.z + s * a.x; r.m[2][2] = t * a.z * a.z + c;
    return r;
}

// ---- g2o::Plane3D (g2oAddition/Plane3D.h) ----
struct Plane3D {
    double c[4];                                     // unit normal + d, d >= 0
    static void normalize(double v[4]) {             // Plane3D::normalize (:117-122)
        const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int k = 0; k < 4; ++k) v[k] = v[k] * (1.0 / n);
        if (v[3] < 0.0) for (int k = 0; k < 4; ++k) v[k] = -v[k];
    }
    static Plane3D from(const double v[4]) { Plane3D p; std::memcpy(p.c, v, sizeof(p.c)); normalize(p.c); return p; }
    // Converter::toPlane3D (src/Converter.cc:151-160)
    static Plane3D from_coefficients(const float coe[4]) {
        double v[4] = {coe[0], coe[1], coe[2], coe[3]};
        if (coe[3] < 0.0f) for (int k = 0; k < 4; ++k) v[k] = -v[k];
        return from(v);
    }
    Vec3 normal() const { return {c[0], c[1], c[2]}; }
    double distance() const { return -c[3]; }
    static double azimuth(Vec3 v) { return std::atan2(v.y, v.x); }
    static double elevation(Vec3 v) { return std::atan2(v.z, std::sqrt(v.x * v.x + v.y * v.y)); }
    static Mat3 rotation(Vec3 v) {                   // (:61-67) azimuth about Z, then -elevation about Y
        return angle_axis(azimuth(v), {0, 0, 1}) * angle_axis(-elevation(v), {0, 1, 0});
    }
    void ominus(const Plane3D &plane, double e[3]) const {          // (:85-91)
        const Vec3 n = transpose(rotation(normal())) * plane.normal();
        e[0] = azimuth(n); e[1] = elevation(n); e[2] = distance() - plane.distance();
    }
    void ominus_ver(const Plane3D &plane, double e[2]) const {      // (:93-102)
        const Vec3 v = cross(normal(), plane.normal());
        const Vec3 b = angle_axis(M_PI / 2, (1.0 / norm(v)) * v) * normal();
        const Vec3 n = transpose(rotation(b)) * plane.normal();
        e[0] = azimuth(n); e[1] = elevation(n);
    }
    void ominus_par(const Plane3D &plane, double e[2]) const {      // (:104-113)
        Vec3 nor = normal();
        if (dot(plane.normal(), nor) < 0) nor = -1.0 * nor;
        const Vec3 n = transpose(rotation(nor)) * plane.normal();
        e[0] = azimuth(n); e[1] = elevation(n);
    }
};

// ---- g2o::SE3Quat as far as VertexSE3Expmap needs it (Thirdparty/g2o/g2o/types/se3quat.h) ----
struct Pose {
    double q[4];   // x y z w, unit
    Vec3 t;
    Mat3 R() const {
        const double x = q[0], y = q[1], z = q[2], w = q[3];
        Mat3 r{};
        r.m[0][0] = 1 - 2 * (y * y + z * z); r.m[0][1] = 2 * (x * y - z * w);     r.m[0][2] = 2 * (x * z + y * w);
        r.m[1][0] = 2 * (x * y + z * w);     r.m[1][1] = 1 - 2 * (x * x + z * z); r.m[1][2] = 2 * (y * z - x * w);
        r.m[2][0] = 2 * (x * z - y * w);     r.m[2][1] = 2 * (y * z + x * w);     r.m[2][2] = 1 - 2 * (x * x + y * y);
        return r;
    }
    static void quat_from(const Mat3 &m, double q[4]) {   // Eigen::Quaterniond(Matrix3d)
        const double tr = m.m[0][0] + m.m[1][1] + m.m[2][2];
        if (tr > 0) {
            double t = std::sqrt(tr + 1.0);
            q[3] = 0.5 * t; t = 0.5 / t;
            q[0] = (m.m[2][1] - m.m[1][2]) * t; q[1] = (m.m[0][2] - m.m[2][0]) * t; q[2] = (m.m[1][0] - m.m[0][1]) * t;
        } else {
            int i = 0;
            if (m.m[1][1] > m.m[0][0]) i = 1;
            if (m.m[2][2] > m.m[i][i]) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            double t = std::sqrt(m.m[i][i] - m.m[j][j] - m.m[k][k] + 1.0);
            q[i] = 0.5 * t; t = 0.5 / t;
            q[3] = (m.m[k][j] - m.m[j][k]) * t; q[j] = (m.m[j][i] + m.m[i][j]) * t; q[k] = (m.m[k][i] + m.m[i][k]) * t;
        }
    }
    void normalize_rotation() {                       // SE3Quat::normalizeRotation
        if (q[3] < 0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
        const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        for (int k = 0; k < 4; ++k) q[k] /= n;
    }
    static Pose from_matrix(const double T[16]) {     // Converter::toSE3Quat (src/Converter.cc:37-47): SE3Quat(R, t)
        Mat3 r{};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = T[4 * i + j];
        Pose p;
        quat_from(r, p.q);
        p.normalize_rotation();
        p.t = {T[3], T[7], T[11]};
        return p;
    }
    void to_matrix(double T[16]) const {
        const Mat3 r = R();
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[4 * i + j] = r.m[i][j];
        T[3] = t.x; T[7] = t.y; T[11] = t.z;
        T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
    }
    // SE3Quat::exp(update): update = (omega, upsilon)
    static Pose exp(const double u[6]) {
        const Vec3 omega{u[0], u[1], u[2]}, upsilon{u[3], u[4], u[5]};
        const double theta = norm(omega);
        const Mat3 Om = skew(omega);
        Mat3 R, V;
        if (theta < 0.00001) {
            R = add(add(Mat3::identity(), Om), Om * Om);
            V = R;
        } else {
            const Mat3 Om2 = Om * Om;
            R = add(add(Mat3::identity(), Om, std::sin(theta) / theta), Om2, (1 - std::cos(theta)) / (theta * theta));
            V = add(add(Mat3::identity(), Om, (1 - std::cos(theta)) / (theta * theta)), Om2, (theta - std::sin(theta)) / std::pow(theta, 3));
        }
        Pose p;
        quat_from(R, p.q);
        p.t = V * upsilon;
        p.normalize_rotation();
        return p;
    }
    // SE3Quat operator*: r = r1 * r2, t = r1 * t2 + t1, then normalizeRotation
    Pose operator*(const Pose &b) const {
        Pose r;
        const double *a = q;
        r.q[3] = a[3] * b.q[3] - a[0] * b.q[0] - a[1] * b.q[1] - a[2] * b.q[2];
        r.q[0] = a[3] * b.q[0] + a[0] * b.q[3] + a[1] * b.q[2] - a[2] * b.q[1];
        r.q[1] = a[3] * b.q[1] + a[1] * b.q[3] + a[2] * b.q[0] - a[0] * b.q[2];
        r.q[2] = a[3] * b.q[2] + a[2] * b.q[3] + a[0] * b.q[1] - a[1] * b.q[0];
        r.t = R() * b.t + t;
        r.normalize_rotation();
        return r;
    }
    // Plane3D operator*(Isometry3D, Plane3D) (g2oAddition/Plane3D.h:127-137)
    Plane3D apply(const Plane3D &pl) const {
        const Vec3 n = R() * pl.normal();
        double v[4] = {n.x, n.y, n.z, pl.c[3] - dot(t, n)};
        if (v[3] < 0.0) for (int k = 0; k < 4; ++k) v[k] = -v[k];
        return Plane3D::from(v);
    }
};

enum PlaneEdgeKind { kEdgePlane = 0, kEdgeParallelPlane = 1, kEdgeVerticalPlane = 2 };

struct PlaneEdge {
    int kind = kEdgePlane;
    Plane3D world;            // VertexPlane estimate (fixed): Converter::toPlane3D(pMP->GetWorldPos())
    Plane3D measurement;      // Converter::toPlane3D(pFrame->mvPlaneCoefficients[i])
    double info[3] = {1, 1, 1};   // diagonal of the information matrix (2 entries used by the 2-d edges)
    double huber_delta = 0;   // rk->setDelta(...)
    double chi2_max = 0;      // planeChi or VPplaneChi: the outlier test after every round
    // state
    bool robust = true, outlier = false;
    int level = 0;
    double error[3] = {0, 0, 0};
    int dim() const { return kind == kEdgePlane ? 3 : 2; }
    void compute_error(const Pose &pose) {
        const Plane3D local = pose.apply(world);
        if (kind == kEdgePlane) local.ominus(measurement, error);
        else if (kind == kEdgeParallelPlane) local.ominus_par(measurement, error);
        else local.ominus_ver(measurement, error);
    }
    double chi2() const { double s = 0; for (int k = 0; k < dim(); ++k) s += error[k] * info[k] * error[k]; return s; }
    // RobustKernelHuber::robustify: rho[0] value, rho[1] first derivative
    void robustify(double e2, double rho[2]) const {
        const double dsqr = huber_delta * huber_delta;
        if (!robust || e2 <= dsqr) { rho[0] = e2; rho[1] = 1.0; }
        else { const double sq = std::sqrt(e2); rho[0] = 2 * sq * huber_delta - dsqr; rho[1] = huber_delta / sq; }
    }
};

// Hessian / gradient contributions of edges this file does not know (the ORB point edges): called with the trial pose,
// adds to H (6x6 row-major, omega first) and b, returns the robustified chi2 of those edges.
typedef double (*ExtraTerms)(const double Tcw[16], double H[36], double b[6], bool linearize, void *user);

class PlanePoseOptimizer {
public:
    std::vector<PlaneEdge> edges;
    ExtraTerms extra = nullptr;
    void *extra_user = nullptr;
    int extra_edges = 0;      // how many graph edges the extra-terms hook stands for (the caller's ORB point edges)
    int iterations_run = 0;
    double final_chi2 = 0;

    // Optimizer::PoseOptimization: rounds x optimize(its) from the initial pose, outlier classification in between;
    // returns the number of outlier edges (nBad of the last round), Tcw is updated in place (row-major 4x4)
    int PoseOptimization(double Tcw[16], int rounds = 4, int its = 10) {
        const Pose initial = Pose::from_matrix(Tcw);
        Pose pose = initial;
        for (PlaneEdge &e : edges) { e.outlier = false; e.level = 0; e.robust = true; }
        int nBad = 0;
        iterations_run = 0;
        for (int it = 0; it < rounds; ++it) {
            pose = initial;                                  // vSE3->setEstimate(Converter::toSE3Quat(pFrame->mTcw))
            optimize(pose, its);
            nBad = 0;
            for (PlaneEdge &e : edges) {
                if (e.outlier) e.compute_error(pose);
                const float chi2 = float(e.chi2());          // `const float chi2 = e->chi2();` (src/Optimizer.cc:1003): truncated before the test
                if (double(chi2) > e.chi2_max) { e.outlier = true; e.level = 1; ++nBad; }
                else { e.outlier = false; e.level = 0; }
                if (it == 2) e.robust = false;
            }
            if (edges.size() + size_t(extra_edges) < 10) break;   // `if(optimizer.edges().size()<10) break;` counts ALL edges of the graph
        }
        pose.to_matrix(Tcw);
        return nBad;
    }

private:
    double active_chi2(const Pose &pose, bool recompute) {
        double s = 0;
        for (PlaneEdge &e : edges) {
            if (e.level != 0) continue;
            if (recompute) e.compute_error(pose);
            double rho[2];
            e.robustify(e.chi2(), rho);
            s += rho[0];
        }
        if (extra) { double T[16]; pose.to_matrix(T); s += extra(T, nullptr, nullptr, false, extra_user); }
        return s;
    }

    void build_system(const Pose &pose, double H[36], double b[6]) {
        std::memset(H, 0, 36 * sizeof(double));
        std::memset(b, 0, 6 * sizeof(double));
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        for (PlaneEdge &e : edges) {
            if (e.level != 0) continue;
            const int D = e.dim();
            double err0[3] = {e.error[0], e.error[1], e.error[2]};
            double J[3][6];
            for (int d = 0; d < 6; ++d) {                    // BaseBinaryEdge::linearizeOplus, numeric
                double u[6] = {0, 0, 0, 0, 0, 0};
                u[d] = delta;
                e.compute_error(Pose::exp(u) * pose);
                double ep[3] = {e.error[0], e.error[1], e.error[2]};
                u[d] = -delta;
                e.compute_error(Pose::exp(u) * pose);
                for (int k = 0; k < D; ++k) J[k][d] = scalar * (ep[k] - e.error[k]);
            }
            for (int k = 0; k < 3; ++k) e.error[k] = err0[k];
            double rho[2];
            e.robustify(e.chi2(), rho);
            for (int k = 0; k < D; ++k) {
                const double w = rho[1] * e.info[k];
                for (int i = 0; i < 6; ++i) {
                    b[i] -= J[k][i] * w * e.error[k];
                    for (int j = 0; j < 6; ++j) H[6 * i + j] += J[k][i] * w * J[k][j];
                }
            }
        }
        if (extra) { double T[16]; pose.to_matrix(T); extra(T, H, b, true, extra_user); }
    }

    static bool solve6(const double H[36], double lambda, const double b[6], double x[6]) {   // Cholesky of H + lambda I
        double L[6][6] = {};
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = H[6 * i + j] + (i == j ? lambda : 0.0);
                for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
                if (i == j) { if (!(s > 0.0)) return false; L[i][i] = std::sqrt(s); }
                else L[i][j] = s / L[j][j];
            }
        double y[6];
        for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
        for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
        return true;
    }

    // SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
    void optimize(Pose &pose, int its) {
        double lambda = 0, ni = 2;
        bool any = extra != nullptr;
        for (const PlaneEdge &e : edges) any = any || e.level == 0;
        if (!any) return;
        for (int iteration = 0; iteration < its; ++iteration) {
            ++iterations_run;
            double currentChi = active_chi2(pose, true);
            double H[36], b[6];
            build_system(pose, H, b);
            if (iteration == 0) {                            // computeLambdaInit: tau * max diagonal
                double mx = 0;
                for (int i = 0; i < 6; ++i) mx = std::fmax(std::fabs(H[7 * i]), mx);
                lambda = 1e-5 * mx;
                ni = 2;
            }
            double rho = 0;
            int qmax = 0;
            do {
                const Pose backup = pose;
                double x[6] = {0, 0, 0, 0, 0, 0};
                const bool ok = solve6(H, lambda, b, x);
                if (ok) pose = Pose::exp(x) * pose;          // VertexSE3Expmap::oplusImpl
                double tempChi = active_chi2(pose, true);
                if (!ok) tempChi = 1e300;
                rho = currentChi - tempChi;
                double scale = 1e-3;
                for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
                rho /= scale;
                if (rho > 0 && std::isfinite(tempChi)) {
                    double alpha = 1.0 - std::pow(2 * rho - 1, 3);
                    alpha = std::fmin(alpha, 2.0 / 3.0);
                    lambda *= std::fmax(1.0 / 3.0, alpha);
                    ni = 2;
                    currentChi = tempChi;
                } else {
                    lambda *= ni;
                    ni *= 2;
                    pose = backup;                           // (the edges keep the errors of the rejected trial, as in g2o)
                    if (!std::isfinite(lambda)) break;
                }
                ++qmax;
            } while (rho < 0 && qmax < 10);
            final_chi2 = currentChi;
            if (qmax == 10 || rho == 0) break;               // OptimizationAlgorithm::Terminate
        }
    }
};

}  // namespace spx_host
