// spx_normals.cuh -- K1..K3: back-projection + depth-change mask, exact capped chamfer distance, integral-image
// normals (AVERAGE_3D_GRADIENT) + plane_d.
//
// Reference: /root/reference/src/Frame.cc:855-885 (cloud construction, IntegralImageNormalEstimation settings) and
// PCL 1.8.0 features/impl/integral_image_normal.hpp (initAverage3DGradientMethod, computeFeature,
// computeFeatureFull, computePointNormal), features/impl/integral_image2D.hpp (IntegralImage2D<float,3>).
#pragma once
#include "spx_math.cuh"
#include "spx_types.cuh"

namespace spx {

constexpr float kDistCap = 10.0f;   // normal_smoothing_size_: the distance map is only consumed through min(d, 10)

// ---------------------------------------------------------------------------------------------------------------
// K1: one thread per organized pixel.  z = d; x = (n - cx) * z / fx; y = (m - cy) * z / fy  (src/Frame.cc:861-865),
// plus the depth-change mask of computeFeature evaluated from the pixel's four neighbours:
//   index pixel (r<=h-2, c<=w-2): |z - zR| > t(z) or |z - zD| > t(z) marks it; it is also marked as the right
//   neighbour of (r, c-1) and the lower neighbour of (r-1, c), with t taken at THAT pixel.
// dist is initialised to 0 (edge) or 10 (= min(width + height, cap)).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_backproject(const float *__restrict__ depth, Params P, Buffers B) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const int r = i / P.w, c = i - r * P.w;
    const char *img = reinterpret_cast<const char *>(depth) + size_t(f) * P.frame_stride;
    auto Z = [&](int rr, int cc) -> float {
        return *reinterpret_cast<const float *>(img + size_t(rr) * P.dis * P.pitch + size_t(cc) * P.dis * sizeof(float));
    };
    auto thr = [&](float z) -> float { return (P.mdcf * (fabsf(z) + 1.0f)) * 2.0f; };
    const float z = Z(r, c);
    const float x = (float(c * P.dis) - P.cx) * z / P.fx;
    const float y = (float(r * P.dis) - P.cy) * z / P.fy;
    bool edge = false;
    const bool zf = isfinite(z);
    if (r <= P.h - 2 && c <= P.w - 2) {
        const float zR = Z(r, c + 1), zD = Z(r + 1, c), t = thr(z);
        if (fabsf(z - zR) > t || !zf || !isfinite(zR)) edge = true;
        if (fabsf(z - zD) > t || !zf || !isfinite(zD)) edge = true;
    }
    if (c >= 1 && r <= P.h - 2) {
        const float zL = Z(r, c - 1);
        if (fabsf(zL - z) > thr(zL) || !zf || !isfinite(zL)) edge = true;
    }
    if (r >= 1 && c <= P.w - 2) {
        const float zU = Z(r - 1, c);
        if (fabsf(zU - z) > thr(zU) || !zf || !isfinite(zU)) edge = true;
    }
    const size_t o = size_t(f) * P.N + i;
    B.px[o] = x; B.py[o] = y; B.pz[o] = z;
    B.dist[o] = edge ? 0.0f : fminf(float(P.w + P.h), kDistCap);
}

// ---------------------------------------------------------------------------------------------------------------
// K2: PCL's two-pass chamfer (1.0 / 1.4f), bit-exact under the cap min(d, 10).  One warp per frame; rows are
// sequential (the recurrence needs the finished previous row), columns are parallel:
//   cur[c] = min(B[c], cur[c-1] (+) 1.0f),  B[c] = min(init, prev[c-1] (+) 1.4f, prev[c] (+) 1.0f, prev[c+1] (+) 1.4f)
// fp32 addition of a positive constant is monotone, so min commutes with it and
//   cur[c] = min_j (B[c-j] (+) 1.0f j times);  every step costs >= 1, so j <= 10 suffices under the cap.
// Each lane owns a contiguous chunk of columns and starts its sequential scan 10 columns early.
// The reference's row wrap-around (previous_row[w] aliases current_row[0]; next_row[-1] aliases current_row[w-1])
// is reproduced.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kChamferWarps = 4;

__global__ void __launch_bounds__(kChamferWarps * 32) k_chamfer(Params P, Buffers B) {
    extern __shared__ float sm_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kChamferWarps + warp;
    if (f >= P.n_frames) return;
    const int w = P.w, h = P.h;
    float *rowA = sm_f + size_t(warp) * 3 * w;
    float *rowB = rowA + w;
    float *Bv = rowB + w;
    float *d = B.dist + size_t(f) * P.N;
    const int CH = (w + 31) / 32;
    const int c0 = lane * CH;
    const int c1 = min(c0 + CH, w);

    float *prev = rowA, *cur = rowB;
    for (int c = lane; c < w; c += 32) prev[c] = d[c];
    __syncwarp();
    // forward pass: rows 1..h-1, columns 1..w-1
    for (int r = 1; r < h; ++r) {
        float *dr = d + size_t(r) * w;
        for (int c = lane; c < w; c += 32) cur[c] = dr[c];
        __syncwarp();
        const float cur0 = cur[0];
        for (int c = lane; c < w; c += 32) {
            float v = cur[c];
            if (c >= 1) {
                const float upLeft = prev[c - 1] + 1.4f;
                const float up = prev[c] + 1.0f;
                const float upRight = (c + 1 < w ? prev[c + 1] : cur0) + 1.4f;
                v = fminf(v, fminf(fminf(upLeft, up), upRight));
            }
            Bv[c] = fminf(v, kDistCap);
        }
        __syncwarp();
        if (c0 < w) {
            int start = max(c0 - 10, 0);
            float run = Bv[start];
            if (start >= c0) cur[start] = run;
            for (int c = start + 1; c < c1; ++c) {
                run = fminf(Bv[c], run + 1.0f);
                if (c >= c0) cur[c] = fminf(run, kDistCap);
            }
        }
        __syncwarp();
        for (int c = lane; c < w; c += 32) dr[c] = cur[c];
        float *t = prev; prev = cur; cur = t;
    }
    // backward pass: rows h-2..0, columns w-2..0; `prev` holds the finished row below
    for (int r = h - 2; r >= 0; --r) {
        float *dr = d + size_t(r) * w;
        for (int c = lane; c < w; c += 32) cur[c] = dr[c];
        __syncwarp();
        const float curLast = cur[w - 1];
        for (int c = lane; c < w; c += 32) {
            float v = cur[c];
            if (c <= w - 2) {
                const float lowerLeft = (c >= 1 ? prev[c - 1] : curLast) + 1.4f;
                const float lower = prev[c] + 1.0f;
                const float lowerRight = prev[c + 1] + 1.4f;
                v = fminf(v, fminf(fminf(lowerLeft, lower), lowerRight));
            }
            Bv[c] = fminf(v, kDistCap);
        }
        __syncwarp();
        if (c0 < w) {
            int start = min(c1 - 1 + 10, w - 1);
            float run = Bv[start];
            if (start < c1) cur[start] = run;
            for (int c = start - 1; c >= c0; --c) {
                run = fminf(Bv[c], run + 1.0f);
                if (c < c1) cur[c] = fminf(run, kDistCap);
            }
        }
        __syncwarp();
        for (int c = lane; c < w; c += 32) dr[c] = cur[c];
        float *t = prev; prev = cur; cur = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: normals.  One CTA per 32x16 tile of a frame.  The central differences of the tile plus a (5 left/up, 4
// right/down) halo are summed into tile-local first-order integral images in fp64 (6 channels + a finite count),
// then every pixel takes its k x k window sum (k = int(min(dist, 10)), rect [c - k/2, c - k/2 + k)) as
//   ((I[y1][x1] + I[y0][x0]) - I[y0][x1]) - I[y1][x0]
// and forms n = normalize(gy x gx) in fp64, casts to fp32 and flips it towards the origin.
// The fp64 sums of fp32 differences are exact for depth data (DESIGN.md "SAT exactness"), hence independent of the
// summation origin/order and bit-identical to PCL's whole-image double integral image.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 16;
constexpr int kHaloL = 5, kHaloR = 4;
constexpr int kRW = kTW + kHaloL + kHaloR;   // 41 region columns
constexpr int kRH = kTH + kHaloL + kHaloR;   // 25 region rows
constexpr int kSW = kRW + 1;                 // 42 integral columns
constexpr int kSH = kRH + 1;                 // 26 integral rows
constexpr int kSATStride = kSW + 1;          // 43 (odd: conflict-light column scans)
constexpr size_t kNormalsSmem = size_t(6) * kSH * kSATStride * sizeof(double) + size_t(kSH) * kSATStride * sizeof(int);

__global__ void __launch_bounds__(256) k_normals(Params P, Buffers B) {
    extern __shared__ double sm_d[];
    double *S = sm_d;                                                   // [6][kSH][kSATStride]
    int *Cn = reinterpret_cast<int *>(S + size_t(6) * kSH * kSATStride);  // [kSH][kSATStride]
    const int f = blockIdx.z;
    const int tc = blockIdx.x * kTW, tr = blockIdx.y * kTH;
    const int w = P.w, h = P.h;
    const size_t fo = size_t(f) * P.N;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    const int tid = threadIdx.x;

    // zero row 0 and column 0 of every integral image
    for (int i = tid; i < 7 * kSW; i += 256) {
        int ch = i / kSW, x = i - ch * kSW;
        if (ch < 6) S[(size_t(ch) * kSH) * kSATStride + x] = 0.0; else Cn[x] = 0;
    }
    for (int i = tid; i < 7 * kSH; i += 256) {
        int ch = i / kSH, y = i - ch * kSH;
        if (ch < 6) S[(size_t(ch) * kSH + y) * kSATStride] = 0.0; else Cn[y * kSATStride] = 0;
    }
    // central differences of the region (initAverage3DGradientMethod): zero on the image border
    for (int i = tid; i < kRW * kRH; i += 256) {
        const int ly = i / kRW, lx = i - ly * kRW;
        const int r = tr - kHaloL + ly, c = tc - kHaloL + lx;
        float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f, dy0 = 0.f, dy1 = 0.f, dy2 = 0.f;
        if (r >= 1 && r <= h - 2 && c >= 1 && c <= w - 2) {
            const int q = r * w + c;
            dx0 = px[q + 1] - px[q - 1]; dx1 = py[q + 1] - py[q - 1]; dx2 = pz[q + 1] - pz[q - 1];
            dy0 = px[q + w] - px[q - w]; dy1 = py[q + w] - py[q - w]; dy2 = pz[q + w] - pz[q - w];
        }
        const bool finx = isfinite(dx0 + (dx1 + dx2)), finy = isfinite(dy0 + (dy1 + dy2));
        const size_t o = size_t(ly + 1) * kSATStride + (lx + 1);
        const size_t cs = size_t(kSH) * kSATStride;
        S[0 * cs + o] = finx ? double(dx0) : 0.0; S[1 * cs + o] = finx ? double(dx1) : 0.0; S[2 * cs + o] = finx ? double(dx2) : 0.0;
        S[3 * cs + o] = finy ? double(dy0) : 0.0; S[4 * cs + o] = finy ? double(dy1) : 0.0; S[5 * cs + o] = finy ? double(dy2) : 0.0;
        // the two images share one count only when both are finite or both are not; keep them packed: lo16 = DX, hi16 = DY
        Cn[o] = (finx ? 1 : 0) | (finy ? 0x10000 : 0);
    }
    __syncthreads();
    // row prefix sums: one thread per (channel, row)
    for (int i = tid; i < 7 * kRH; i += 256) {
        const int ch = i / kRH, y = i - ch * kRH + 1;
        if (ch < 6) {
            double *row = S + (size_t(ch) * kSH + y) * kSATStride;
            double run = 0.0;
            for (int x = 1; x < kSW; ++x) { run += row[x]; row[x] = run; }
        } else {
            int *row = Cn + y * kSATStride;
            int run = 0;
            for (int x = 1; x < kSW; ++x) { run += row[x]; row[x] = run; }
        }
    }
    __syncthreads();
    // column prefix sums: one thread per (channel, column)
    for (int i = tid; i < 7 * kRW; i += 256) {
        const int ch = i / kRW, x = i - ch * kRW + 1;
        if (ch < 6) {
            double *col = S + size_t(ch) * kSH * kSATStride + x;
            double run = 0.0;
            for (int y = 1; y < kSH; ++y) { run += col[size_t(y) * kSATStride]; col[size_t(y) * kSATStride] = run; }
        } else {
            int *col = Cn + x;
            int run = 0;
            for (int y = 1; y < kSH; ++y) { run += col[y * kSATStride]; col[y * kSATStride] = run; }
        }
    }
    __syncthreads();

    const float qnan = __int_as_float(0x7fc00000);
    const int border = 10;
    for (int i = tid; i < kTW * kTH; i += 256) {
        const int ly = i / kTW, lx = i - ly * kTW;
        const int r = tr + ly, c = tc + lx;
        if (r >= h || c >= w) continue;
        const int q = r * w + c;
        float nx = qnan, ny = qnan, nz = qnan;
        const float X = px[q], Y = py[q], Zv = pz[q];
        if (r >= border && r < h - border && c >= border && c < w - border && isfinite(Zv)) {
            const float smoothing = fminf(B.dist[fo + q], kDistCap);
            if (smoothing > 2.0f) {
                const int k = int(smoothing), half = k / 2;
                const int x0 = lx + kHaloL - half, y0 = ly + kHaloL - half;   // integral-image coordinates
                const int x1 = x0 + k, y1 = y0 + k;
                const size_t ul = size_t(y0) * kSATStride + x0, ur = size_t(y0) * kSATStride + x1;
                const size_t ll = size_t(y1) * kSATStride + x0, lr = size_t(y1) * kSATStride + x1;
                const int cn = Cn[lr] + Cn[ul] - Cn[ur] - Cn[ll];
                if ((cn & 0xffff) != 0 && (cn >> 16) != 0) {
                    const size_t cs = size_t(kSH) * kSATStride;
                    double g[6];
#pragma unroll
                    for (int ch = 0; ch < 6; ++ch) {
                        const double *I = S + ch * cs;
                        g[ch] = ((I[lr] + I[ul]) - I[ur]) - I[ll];
                    }
                    // normal_vector = gradient_y.cross(gradient_x)
                    const double n0 = g[4] * g[2] - g[5] * g[1];
                    const double n1 = g[5] * g[0] - g[3] * g[2];
                    const double n2 = g[3] * g[1] - g[4] * g[0];
                    const double len = (n0 * n0 + n1 * n1) + n2 * n2;
                    if (len != 0.0) {
                        const double s = sqrt(len);
                        nx = float(n0 / s); ny = float(n1 / s); nz = float(n2 / s);
                        // flipNormalTowardsViewpoint(point, 0, 0, 0, nx, ny, nz)
                        const float vx = 0.0f - X, vy = 0.0f - Y, vz = 0.0f - Zv;
                        const float cos_theta = (vx * nx + vy * ny + vz * nz);
                        if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
                    }
                }
            }
        }
        B.nx[fo + q] = nx; B.ny[fo + q] = ny; B.nz[fo + q] = nz;
        B.pd[fo + q] = dot3f(X, Y, Zv, nx, ny, nz);
    }
}

// plane_d for caller-supplied normals ("feed the reference's normals")
__global__ void __launch_bounds__(256) k_plane_d(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const size_t o = size_t(f) * P.N + i;
    B.pd[o] = dot3f(B.px[o], B.py[o], B.pz[o], B.nx[o], B.ny[o], B.nz[o]);
}

}  // namespace spx
