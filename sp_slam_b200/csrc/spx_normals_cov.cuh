// spx_normals_cov.cuh -- optional normal estimation method COVARIANCE_MATRIX of pcl::IntegralImageNormalEstimation
// (spx_config::normal_method = 1): the 9-channel integral image of x, y, z and their products, per-pixel 3x3 covariance,
// closed-form eigen solve, normal and curvature -- the method BASELINE.json's north_star words.  The reference itself selects
// AVERAGE_3D_GRADIENT (/root/reference/src/Frame.cc:880), which stays the default and the production path (spx_normals*.cuh).
//
// PCL 1.8.0 features/impl/integral_image2D.hpp (computeIntegralImages with second order), features/impl/
// integral_image_normal.hpp (computePointNormal, COVARIANCE_MATRIX branch), common/impl/eigen.hpp (eigen33).
//
// Unlike the gradient images, sums of coordinates and of their products are NOT exact in double, so PCL's result depends on
// the order of its recurrence  I[r+1][c+1] = (I[r][c+1] + I[r+1][c]) - I[r][c] (+ element if finite)  over the WHOLE image.
// k_cov_sat evaluates exactly that recurrence, element by element, on an anti-diagonal wavefront (one CTA per frame, one
// thread per row, a barrier per diagonal): every node is the same function of the same three predecessors as in PCL, so the
// integral images -- and everything computed from them -- carry PCL's own rounding.
#pragma once
#include "spx_math.cuh"
#include "spx_types.cuh"

namespace spx {

constexpr int kCovThreads = 256;
constexpr int kCovCh = 9;             // x y z | xx xy xz yy yz zz

__global__ void __launch_bounds__(kCovThreads) k_cov_sat(Params P, Buffers B, double *__restrict__ sat_all, unsigned *__restrict__ cnt_all) {
    __shared__ double ex[2][kCovThreads][kCovCh];     // the results of the previous diagonal, per row
    __shared__ unsigned cex[2][kCovThreads];
    const int f = P.frame0 + blockIdx.x, t = threadIdx.x;
    const int w = P.w, h = P.h, W1 = w + 1;
    const size_t fo = size_t(f) * P.N;
    const size_t nodes = size_t(W1) * (h + 1);
    double *sat = sat_all + size_t(f) * nodes * kCovCh;
    unsigned *cnt = cnt_all + size_t(f) * nodes;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    for (int i = t; i < W1; i += kCovThreads) {                   // node row 0
#pragma unroll
        for (int k = 0; k < kCovCh; ++k) sat[size_t(i) * kCovCh + k] = 0.0;
        cnt[i] = 0u;
    }
    for (int rb = 0; rb < h; rb += kCovThreads) {
        const int nrow = min(kCovThreads, h - rb);
        const int r = rb + t;
        double left[kCovCh], upl[kCovCh];
        unsigned cleft = 0u, cupl = 0u;
#pragma unroll
        for (int k = 0; k < kCovCh; ++k) { left[k] = 0.0; upl[k] = 0.0; }
        if (t < nrow) {                                           // node column 0 of the thread's row
            double *n0 = sat + size_t(r + 1) * W1 * kCovCh;
#pragma unroll
            for (int k = 0; k < kCovCh; ++k) n0[k] = 0.0;
            cnt[size_t(r + 1) * W1] = 0u;
        }
        __syncthreads();                                          // (the previous block's last row is in global memory)
        for (int s = 0; s < w + nrow - 1; ++s) {
            const int cur = s & 1, prv = cur ^ 1;
            const int c = s - t;
            if (t < nrow && c >= 0 && c < w) {
                double up[kCovCh];
                unsigned cup;
                if (t == 0) {
                    const double *g = sat + (size_t(rb) * W1 + c + 1) * kCovCh;      // node (rb, c + 1): zero row, or the block above
#pragma unroll
                    for (int k = 0; k < kCovCh; ++k) up[k] = __ldcg(g + k);
                    cup = __ldcg(cnt + size_t(rb) * W1 + c + 1);
                } else {
#pragma unroll
                    for (int k = 0; k < kCovCh; ++k) up[k] = ex[prv][t - 1][k];
                    cup = cex[prv][t - 1];
                }
                const int q = r * w + c;
                const float e0 = px[q], e1 = py[q], e2 = pz[q];
                const bool fin = isfinite(e0 + (e1 + e2));       // pcl_isfinite(element->sum())
                double v[kCovCh];
#pragma unroll
                for (int k = 0; k < kCovCh; ++k) v[k] = (up[k] + left[k]) - upl[k];
                unsigned cv = cup + cleft - cupl;
                if (fin) {
                    // so_element[el] = element[a] * element[b]: a float product, widened afterwards
                    v[0] += double(e0); v[1] += double(e1); v[2] += double(e2);
                    v[3] += double(e0 * e0); v[4] += double(e0 * e1); v[5] += double(e0 * e2);
                    v[6] += double(e1 * e1); v[7] += double(e1 * e2); v[8] += double(e2 * e2);
                    ++cv;
                }
                double *g = sat + (size_t(r + 1) * W1 + c + 1) * kCovCh;
#pragma unroll
                for (int k = 0; k < kCovCh; ++k) { g[k] = v[k]; ex[cur][t][k] = v[k]; left[k] = v[k]; upl[k] = up[k]; }
                cnt[size_t(r + 1) * W1 + c + 1] = cv;
                cex[cur][t] = cv; cleft = cv; cupl = cup;
            }
            __syncthreads();
        }
    }
}

// computePointNormal, COVARIANCE_MATRIX: one thread per organized pixel
__global__ void __launch_bounds__(256) k_cov_normals(Params P, Buffers B, const double *__restrict__ sat_all, const unsigned *__restrict__ cnt_all,
                                                     float *__restrict__ curv, int alt_trace) {
    const int f = P.frame0 + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const int w = P.w, h = P.h, W1 = w + 1;
    const size_t o = size_t(f) * P.N + i;
    const float qnan = __int_as_float(0x7fc00000);
    float nx = qnan, ny = qnan, nz = qnan, cv = qnan;
    const int r = i / w, c = i - r * w;
    constexpr int border = 10;
    const float X = B.px[o], Y = B.py[o], Zv = B.pz[o];
    if (r >= border && r < h - border && c >= border && c < w - border && isfinite(Zv)) {
        const int k = int(B.kwin[o]);
        if (k > 0) {
            const int half = k / 2, x0 = c - half, y0 = r - half;
            const size_t nodes = size_t(W1) * (h + 1);
            const size_t ul = size_t(y0) * W1 + x0, ur = ul + k, ll = size_t(y0 + k) * W1 + x0, lr = ll + k;
            const unsigned *cn = cnt_all + size_t(f) * nodes;
            const unsigned count = cn[lr] + cn[ul] - cn[ur] - cn[ll];
            if (count != 0u) {
                const double *I = sat_all + size_t(f) * nodes * kCovCh;
                double s[kCovCh];
#pragma unroll
                for (int ch = 0; ch < kCovCh; ++ch) s[ch] = ((I[lr * kCovCh + ch] + I[ul * kCovCh + ch]) - I[ur * kCovCh + ch]) - I[ll * kCovCh + ch];
                const float center[3] = {float(s[0]), float(s[1]), float(s[2])};
                float C[9];
                C[0] = float(s[3]); C[1] = C[3] = float(s[4]); C[2] = C[6] = float(s[5]);
                C[4] = float(s[6]); C[5] = C[7] = float(s[7]); C[8] = float(s[8]);
                const float cntf = float(count);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) C[a * 3 + b] -= (center[a] * center[b]) / cntf;
                float ev, vec[3];
                eigen33_smallest(C, ev, vec);
                nx = vec[0]; ny = vec[1]; nz = vec[2];
                const float vx = 0.0f - X, vy = 0.0f - Y, vz = 0.0f - Zv;
                const float cos_theta = (vx * nx + vy * ny + vz * nz);
                if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
                // PCL 1.8.0 divides by coeff(0) + coeff(2) + coeff(4) (not the trace); alt_trace selects the trace
                const float den = alt_trace ? (C[0] + C[4] + C[8]) : (C[0] + C[2] + C[4]);
                cv = ev > 0.0f ? fabsf(ev / den) : 0.0f;
            }
        }
    }
    B.nx[o] = nx; B.ny[o] = ny; B.nz[o] = nz;
    if (curv) curv[o] = cv;
}

}  // namespace spx
