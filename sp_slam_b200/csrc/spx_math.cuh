// spx_math.cuh -- scalar arithmetic shared by the kernels: the closed-form 3x3 eigen solver and the fixed
// evaluation orders of the reference's Eigen expressions.  The translation unit is compiled with -fmad=false so no
// multiply-add is ever contracted (the reference's PCL path rounds every product and every sum).
//
// Follows PCL 1.8.0 common/impl/eigen.hpp (computeRoots2, computeRoots, eigen33, computeCorrespondingEigenVector)
// as called from OrganizedMultiPlaneSegmentation::segment (reference call site /root/reference/src/Frame.cc:905)
// and SampleConsensusModelLine::optimizeModelCoefficients (call site src/Frame.cc:964).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#define SPX_HD __host__ __device__ __forceinline__
#define SPX_FULL 0xffffffffu

namespace spx {

// Eigen 3.3 fixed-size reduction orders (see DESIGN.md "arithmetic conventions")
SPX_HD float dot3f(float a0, float a1, float a2, float b0, float b1, float b2) { return a0 * b0 + (a1 * b1 + a2 * b2); }
SPX_HD float dot4f(const float *a, const float *b) { return (a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3]); }

SPX_HD void cross3f(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

SPX_HD void swapf(float &a, float &b) { float t = a; a = b; b = t; }

SPX_HD void compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = float(double(b * b) - 4.0 * double(c));
    if (d < 0.0f) d = 0.0f;
    float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}

// The three transcendental calls are evaluated in double and rounded to float: the closest thing to the correctly
// rounded float result that both glibc's atan2f/cosf/sinf and CUDA's approximate.
SPX_HD void compute_roots(const float m[9], float roots[3]) {
    const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
    float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
    float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
    float c2 = m00 + m11 + m22;
    if (fabsf(c0) < FLT_EPSILON) {
        compute_roots2(c2, c1, roots);
    } else {
        const float s_inv3 = float(1.0 / 3.0);
        const float s_sqrt3 = sqrtf(3.0f);
        float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.0f) a_over_3 = 0.0f;
        float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.0f) q = 0.0f;
        float rho = sqrtf(-a_over_3);
        float theta = float(atan2(double(sqrtf(-q)), double(half_b))) * s_inv3;
        float cos_theta = float(cos(double(theta)));
        float sin_theta = float(sin(double(theta)));
        roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
        roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
        if (roots[1] >= roots[2]) {
            swapf(roots[1], roots[2]);
            if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
        }
        if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
    }
}

SPX_HD float scale_of(const float mat[9]) {
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(mat[i]));
    if (scale <= FLT_MIN) scale = 1.0f;
    return scale;
}

// eigenvector of (S - shift*I): the largest of the three row cross products, normalised
SPX_HD void eigvec_from_shifted(float S[9], float shift, float vec[3]) {
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
    float s = sqrtf(l);
    vec[0] = v[0] / s; vec[1] = v[1] / s; vec[2] = v[2] / s;
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair of a symmetric 3x3 (row major)
SPX_HD void eigen33_smallest(const float mat[9], float &eigenvalue, float vec[3]) {
    float scale = scale_of(mat);
    float S[9];
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    float roots[3];
    compute_roots(S, roots);
    eigenvalue = roots[0] * scale;
    eigvec_from_shifted(S, roots[0], vec);
}

// pcl::eigen33(mat, evals) followed by computeCorrespondingEigenVector(mat, evals[2], vec)
SPX_HD void eigen33_largest_vec(const float mat[9], float vec[3]) {
    float scale = scale_of(mat);
    float S[9];
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    float evals[3];
    compute_roots(S, evals);
    for (int i = 0; i < 3; ++i) evals[i] *= scale;
    for (int i = 0; i < 9; ++i) S[i] = mat[i] / scale;
    eigvec_from_shifted(S, evals[2] / scale, vec);
}

#ifdef __CUDACC__
// (n0, n1, n2) / sqrt(len) rounded to float -- the bits of `float(n_i / sqrt(len))` evaluated in IEEE double (Eigen's
// `normal_vector /= sqrt(length)` followed by the cast in computePointNormal) -- without the double square root and the
// three double divisions.  r ~ len^-1/2 comes from the single-precision rsqrt refined by two Newton steps in double
// (relative error below 2^-50 whatever the seed's last bits are); q_i = RN(n_i * r) and the reference
// t_i = RN(n_i / RN(sqrt(len))) then differ by less than 2^-49 relative, i.e. fewer than 16 units in the last place of a
// double.  A float keeps 23 of the 52 fraction bits: q_i and t_i round to the same float unless the 29 dropped bits of q_i
// lie within 32 units of the tie pattern 0x10000000 (probability 2^-22 per component), in which case -- like for results a
// float would hold as a subnormal, and for a len outside the range the float seed covers -- the exact sequence is evaluated.
__device__ __forceinline__ void normalize_to_float(double n0, double n1, double n2, double len, float &x, float &y, float &z) {
    bool exact_path = !(len > 1.0e-30 && len < 1.0e30);
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    if (!exact_path) {
        double r = double(rsqrtf(float(len)));          // ~2^-22
        const double hl = 0.5 * len;
        r = r * (1.5 - hl * r * r);                     // ~2^-43
        r = r * (1.5 - hl * r * r);                     // rounding only
        q0 = n0 * r; q1 = n1 * r; q2 = n2 * r;
        auto risky = [](double q) -> bool {
            const unsigned lo = unsigned(__double2loint(q)) & 0x1fffffffu;
            return (lo - 0x0fffffe0u) <= 64u || (fabs(q) < 1.0e-30 && q != 0.0);
        };
        exact_path = risky(q0) || risky(q1) || risky(q2);
    }
    if (exact_path) {
        const double sl = sqrt(len);
        q0 = n0 / sl; q1 = n1 / sl; q2 = n2 / sl;
    }
    x = float(q0); y = float(q1); z = float(q2);
}
#endif

SPX_HD uint32_t pack_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a = 255u) {
    return (a << 24) | (r << 16) | (g << 8) | b;
}

}  // namespace spx
