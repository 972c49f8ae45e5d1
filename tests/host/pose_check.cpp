// Stand-alone use of sp_slam_b200/host/PlanePoseOptimizer.h (no CUDA, no libspx): the plane edges of a box room plus the
// `extra` hook through which a caller adds the Hessian blocks of edges this repository does not implement (the ORB point
// edges of Optimizer::PoseOptimization).  Prints the recovered pose error; tests/test_pose_planes.py checks the numbers.
#include <cstdio>
#include <cmath>

#include "../../sp_slam_b200/host/PlanePoseOptimizer.h"

using namespace spx_host;

// a quadratic prior that pulls the translation towards (tx, ty, tz): chi2 = w * |t - t0|^2, linearised on the se3 update
struct Prior { double t0[3]; double w; };
static double prior_terms(const double T[16], double H[36], double b[6], bool linearize, void *user) {
    const Prior *p = static_cast<const Prior *>(user);
    const double r[3] = {T[3] - p->t0[0], T[7] - p->t0[1], T[11] - p->t0[2]};
    if (linearize) {
        // t' = exp(update) * T: dt/d(upsilon) = I, dt/d(omega) = -[t]x
        const double t[3] = {T[3], T[7], T[11]};
        double J[3][6] = {{0, t[2], -t[1], 1, 0, 0}, {-t[2], 0, t[0], 0, 1, 0}, {t[1], -t[0], 0, 0, 0, 1}};
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 6; ++i) {
                b[i] -= J[k][i] * p->w * r[k];
                for (int j = 0; j < 6; ++j) H[6 * i + j] += J[k][i] * p->w * J[k][j];
            }
    }
    return p->w * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
}

int main() {
    const double room[6][4] = {{0, 1, 0, 1.4}, {0, -1, 0, 1.6}, {1, 0, 0, 3.0}, {-1, 0, 0, 3.0}, {0, 0, 1, 2.5}, {0, 0, -1, 2.5}};
    const double u[6] = {0.04, -0.08, 0.03, 0.2, -0.1, 0.3};
    const Pose gt = Pose::exp(u);
    PlanePoseOptimizer opt;
    for (int k = 0; k < 6; ++k) {
        PlaneEdge e;
        e.kind = kEdgePlane;
        e.world = Plane3D::from(room[k]);
        e.measurement = gt.apply(e.world);
        e.info[0] = e.info[1] = 3282.8; e.info[2] = 1e4;
        e.huber_delta = std::sqrt(300.0); e.chi2_max = 300.0;
        opt.edges.push_back(e);
    }
    // start: the previous frame's pose, a few centimetres / a degree away.  (Not the identity: with the camera axes exactly
    // on the plane normals Plane3D's azimuth is atan2(0, 0) and g2o's numeric Jacobian -- restated here -- is meaningless.)
    const double u0[6] = {0.02, -0.05, 0.01, 0.15, -0.05, 0.2};
    double T[16];
    Pose::exp(u0).to_matrix(T);
    int bad = opt.PoseOptimization(T);
    double Tg[16];
    gt.to_matrix(Tg);
    double err = 0;
    for (int k = 0; k < 12; ++k) err = std::fmax(err, std::fabs(T[k] - Tg[k]));
    std::printf("planes_only bad %d err %.3e iterations %d\n", bad, err, opt.iterations_run);

    // only the floor and the ceiling: the translation in x and z is unobservable; the prior fixes it
    PlanePoseOptimizer two;
    two.edges.assign(opt.edges.begin(), opt.edges.begin() + 2);
    for (PlaneEdge &e : two.edges) { e.outlier = false; e.level = 0; e.robust = true; }
    Prior pr{{Tg[3], Tg[7], Tg[11]}, 50.0};
    two.extra = prior_terms; two.extra_user = &pr;
    double T2[16];
    Pose::exp(u0).to_matrix(T2);
    bad = two.PoseOptimization(T2);
    const double terr = std::fmax(std::fabs(T2[3] - Tg[3]), std::fmax(std::fabs(T2[7] - Tg[7]), std::fabs(T2[11] - Tg[11])));
    std::printf("with_prior bad %d translation_err %.3e\n", bad, terr);
    return 0;
}
