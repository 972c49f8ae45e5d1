// spx_normals.cuh -- K1..K3: depth-change mask + exact capped chamfer distance (band parallel), and the fused tile
// kernel: back-projection, integral-image normals (AVERAGE_3D_GRADIENT), plane_d, comparator links.
//
// Reference: /root/reference/src/Frame.cc:855-885 (cloud construction, IntegralImageNormalEstimation settings) and
// PCL 1.8.0 features/impl/integral_image_normal.hpp (initAverage3DGradientMethod, computeFeature,
// computeFeatureFull, computePointNormal), features/impl/integral_image2D.hpp (IntegralImage2D<float,3>).
#pragma once
#include <climits>
#include "spx_math.cuh"
#include "spx_types.cuh"

namespace spx {

constexpr float kDistCap = 10.0f;   // normal_smoothing_size_: the distance map is only consumed through min(d, 10)

// ---------------------------------------------------------------------------------------------------------------
// K1 (feed-normals path only; the normal path back-projects inside k_normals_link): one thread per organized pixel.
// z = d; x = (n - cx) * z / fx; y = (m - cy) * z / fy  (src/Frame.cc:861-865).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_backproject(const float *__restrict__ depth, Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const int r = i / P.w, c = i - r * P.w;
    const char *img = reinterpret_cast<const char *>(depth) + size_t(f) * P.samp_fstride;
    const float z = *reinterpret_cast<const float *>(img + size_t(r) * P.samp_rstep + size_t(c) * P.samp_cstep * sizeof(float));
    const float x = (float(c * P.dis) - P.cx) * z / P.fx;
    const float y = (float(r * P.dis) - P.cy) * z / P.fy;
    const size_t o = size_t(f) * P.N + i;
    B.px[o] = x; B.py[o] = y; B.pz[o] = z;
    B.dist[o] = 0.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: depth-change mask + PCL's two-pass chamfer (1.0 / 1.4f), bit-exact under the cap min(d, 10), straight from the
// depth image.  One warp per BAND of 32 organized rows; columns are parallel, rows sequential:
//   cur[c] = min(B[c], cur[c-1] (+) 1.0f),  B[c] = min(init, prev[c-1] (+) 1.4f, prev[c] (+) 1.0f, prev[c+1] (+) 1.4f)
// fp32 addition of a positive constant is monotone, so min commutes with it and
//   cur[c] = min_j (B[c-j] (+) 1.0f j times);  every step costs >= 1, so j <= 10 suffices under the cap
// (each lane owns a contiguous chunk of columns and starts its sequential scan 10 columns early).  For the same
// reason a row only depends on the 10 rows before it in pass order: a band runs the forward pass over
// [r0-10, r1+10) and the backward pass over [r1+9 .. r0], starting both from the initial values, and is exact on its
// own rows [r0, r1) -- 5 independent warps per 160-row frame instead of one 320-step chain.
// The depth-change mask of computeFeature is evaluated from the pixel's four neighbours:
//   index pixel (r<=h-2, c<=w-2): |z - zR| > t(z) or |z - zD| > t(z) marks it; it is also marked as the right
//   neighbour of (r, c-1) and the lower neighbour of (r-1, c), with t taken at THAT pixel.
// The reference's row wrap-around (previous_row[w] aliases current_row[0]; next_row[-1] aliases current_row[w-1])
// is reproduced.  Output: kwin = int(min(dist, 10)) if that is > 2 else 0 (all the normal estimation consumes).
// ---------------------------------------------------------------------------------------------------------------
// One row of a chamfer pass for the CW contiguous columns [c0, c0 + CW) a lane owns (k_edge_chamfer).
// kFwd: left-to-right chain, `prev` = the row above; else right-to-left, `prev` = the row below.  `wrapv` is the
// reference's aliased neighbour (forward: current row's column 0; backward: its column w-1).
template <int CW, bool kFwd>
__device__ __forceinline__ void chamfer_row_step(const float (&prev)[CW], float (&cur)[CW], float wrapv, int c0, int w, int lane) {
    const float kInf = 3.0e38f;
    // neighbours of the own columns in the previous row
    float pl = __shfl_up_sync(SPX_FULL, prev[CW - 1], 1);     // prev[c0 - 1]
    float pr = __shfl_down_sync(SPX_FULL, prev[0], 1);        // prev[c0 + CW]
    if (lane == 0) pl = kInf;
    if (lane == 31) pr = kInf;
    float Bv[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) {
        const int c = c0 + k;
        float a = (k == 0) ? pl : prev[k == 0 ? 0 : k - 1];                  // prev[c - 1]
        float d = (k == CW - 1) ? pr : prev[k == CW - 1 ? k : k + 1];        // prev[c + 1]
        if (kFwd) { if (c == w - 1) d = wrapv; } else { if (c == 0) a = wrapv; }
        float v = cur[k];
        const bool upd = kFwd ? (c >= 1 && c < w) : (c <= w - 2);
        if (upd) v = fminf(v, fminf(fminf(a + 1.4f, prev[k] + 1.0f), d + 1.4f));
        Bv[k] = (c < w) ? fminf(v, kDistCap) : kInf;
    }
    // local chain in visiting order, then two carry rounds from the neighbouring lane
    // (column 0 in the forward pass and column w-1 in the backward pass are not updated by the row at all)
    if (kFwd) {
        float run = kInf;
#pragma unroll
        for (int k = 0; k < CW; ++k) { const int c = c0 + k; run = (c >= 1) ? fminf(Bv[k], run + 1.0f) : Bv[k]; cur[k] = fminf(run, kDistCap); }
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            float carry = __shfl_up_sync(SPX_FULL, cur[CW - 1], 1);
            if (lane == 0) carry = kInf;
#pragma unroll
            for (int k = 0; k < CW; ++k) { const int c = c0 + k; carry = carry + 1.0f; if (c >= 1 && c < w) cur[k] = fminf(cur[k], fminf(carry, kDistCap)); }
        }
    } else {
        float run = kInf;
#pragma unroll
        for (int k = CW - 1; k >= 0; --k) { const int c = c0 + k; run = (c <= w - 2) ? fminf(Bv[k], run + 1.0f) : Bv[k]; cur[k] = fminf(run, kDistCap); }
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            float carry = __shfl_down_sync(SPX_FULL, cur[0], 1);
            if (lane == 31) carry = kInf;
#pragma unroll
            for (int k = CW - 1; k >= 0; --k) { const int c = c0 + k; carry = carry + 1.0f; if (c <= w - 2) cur[k] = fminf(cur[k], fminf(carry, kDistCap)); }
        }
    }
}

constexpr int kChamferWarps = 4;
constexpr int kBandRows = 32;
constexpr int kBandHalo = 10;
constexpr int kBandSpan = kBandRows + 2 * kBandHalo;   // mask rows a band looks at

template <int NCH>
__global__ void __launch_bounds__(kChamferWarps * 32) k_edge_chamfer(const float *__restrict__ depth, Params P, Buffers B, int write_dist) {
    extern __shared__ float sm_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = P.w, h = P.h;
    const int nb = (h + kBandRows - 1) / kBandRows;
    const int g = blockIdx.x * kChamferWarps + warp;
    const int fl = g / nb, band = g - fl * nb;
    if (fl >= P.n_frames) return;
    const int f = P.frame0 + fl;
    unsigned *mbits = reinterpret_cast<unsigned *>(sm_f) + size_t(warp) * (kBandSpan * NCH);   // [kBandSpan][NCH]
    const size_t fo = size_t(f) * P.N;
    const int r0 = band * kBandRows, r1 = min(r0 + kBandRows, h);
    const int ra = max(r0 - kBandHalo, 0), rb = min(r1 + kBandHalo, h);
    const char *img = reinterpret_cast<const char *>(depth) + size_t(f) * P.samp_fstride;
    const float initv = fminf(float(w + h), kDistCap);

    // ---- phase 1: mask bits of rows [ra, rb) ----
    {
        // PCL marks BOTH pixels of a pair whose depth change exceeds the threshold of the pair's first pixel (right and lower
        // neighbour, pixels with r <= h-2 and c <= w-2).  Every pixel evaluates its own two tests once; the marks a pixel receives
        // from its left / upper neighbour are those neighbours' ballots, shifted by one lane / kept from the previous row.
        float zC[NCH], zD[NCH], zN[NCH];
        unsigned downPrev[NCH];                         // lower-neighbour tests of the previous row
        auto load_row = [&](int r, float (&dst)[NCH]) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = ch * 32 + lane;
                dst[ch] = (r >= 0 && r < h && c < w)
                              ? *reinterpret_cast<const float *>(img + size_t(r) * P.samp_rstep + size_t(c) * P.samp_cstep * sizeof(float))
                              : 0.0f;
            }
        };
        auto thr = [&](float z) -> float { return (P.mdcf * (fabsf(z) + 1.0f)) * 2.0f; };
        const int rs = ra >= 1 ? ra - 1 : ra;          // one row early: its lower-neighbour tests mark row ra
        load_row(rs, zC); load_row(rs + 1, zD);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) downPrev[ch] = 0u;
        bool nonfinite = false;
        for (int r = rs; r < rb; ++r) {
            load_row(r + 2, zN);
            unsigned rightPrev = 0u;                   // right-neighbour tests of the previous chunk of this row
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = ch * 32 + lane;
                const float z = zC[ch];
                const bool zf = isfinite(z);
                if (c < w && !zf) nonfinite = true;
                float zR = __shfl_down_sync(SPX_FULL, z, 1);
                if (ch + 1 < NCH) { const float t = __shfl_sync(SPX_FULL, zC[ch + 1 < NCH ? ch + 1 : ch], 0); if (lane == 31) zR = t; }
                const float zDn = zD[ch], t = thr(z);
                const bool dom = r <= h - 2 && c <= w - 2;
                const bool eR = dom && (fabsf(z - zR) > t || !zf || !isfinite(zR));
                const bool eD = dom && (fabsf(z - zDn) > t || !zf || !isfinite(zDn));
                const unsigned bR = __ballot_sync(SPX_FULL, eR), bD = __ballot_sync(SPX_FULL, eD);
                const unsigned bits = bR | bD | (bR << 1) | (rightPrev >> 31) | downPrev[ch];
                if (lane == 0 && r >= ra) mbits[(r - ra) * NCH + ch] = bits;
                rightPrev = bR; downPrev[ch] = bD;
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) { zC[ch] = zD[ch]; zD[ch] = zN[ch]; }
        }
        // a frame with NaN / Inf depth is flagged and queued for k_normals_link_list (finite-count integral images); the strip
        // kernel skips it
        if (__any_sync(SPX_FULL, nonfinite) && lane == 0) {
            const unsigned old = atomicOr(&B.ctl[f].flags, unsigned(SPX_FRAME_NONFINITE));
            if (!(old & unsigned(SPX_FRAME_NONFINITE))) B.nf_list[1 + atomicAdd(&B.nf_list[0], 1)] = f;
        }
    }
    __syncwarp();
    // no edge anywhere near the band: the distance is the initial value everywhere
    unsigned anyedge = 0u;
    for (int i = lane; i < (rb - ra) * NCH; i += 32) anyedge |= mbits[i];
    anyedge = __reduce_or_sync(SPX_FULL, anyedge);
    uint8_t *kwin = B.kwin + fo;
    float *dist = B.dist + fo;
    if (anyedge == 0u) {
        const uint8_t kk = initv > 2.0f ? uint8_t(int(initv)) : uint8_t(0);
        for (int r = r0; r < r1; ++r)
            for (int c = lane; c < w; c += 32) { kwin[r * w + c] = kk; if (write_dist) dist[r * w + c] = initv; }
        return;
    }
    // ---- phase 2: the two raster passes, rows in registers ----
    // lane l owns the CW contiguous columns [l*CW, l*CW + CW); neighbours across lanes come from shuffles.  Within a
    // row the recurrence cur[c] = min(B[c], cur[c-1] (+) 1.0f) only reaches 10 columns back under the cap (< 2 lanes),
    // so a local scan plus two rounds of carry from the neighbouring lane reproduce the sequential scan exactly.
    constexpr int CW = NCH;   // 32 * NCH columns over 32 lanes
    const int c0 = lane * CW;
    float *scratch = B.cham_tmp + (size_t(f) * nb + band) * size_t(kBandRows + kBandHalo) * w;   // forward rows [r0, rb)
    const float kInf = 3.0e38f;
    auto init_row = [&](int r, float (&v)[CW]) {
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            const int c = c0 + k;
            v[k] = (c < w) ? (((mbits[(r - ra) * NCH + (c >> 5)] >> (c & 31)) & 1u) ? 0.0f : initv) : kInf;
        }
    };
    auto store_row = [&](float *dst, const float (&v)[CW]) {
#pragma unroll
        for (int k = 0; k < CW; ++k) if (c0 + k < w) dst[c0 + k] = v[k];
    };
    auto load_row_f = [&](const float *src, float (&v)[CW]) {
#pragma unroll
        for (int k = 0; k < CW; ++k) v[k] = (c0 + k < w) ? src[c0 + k] : kInf;
    };
    auto emit = [&](int r, const float (&v)[CW]) {
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            const int c = c0 + k;
            if (c < w) {
                const float sv = fminf(v[k], kDistCap);
                kwin[r * w + c] = sv > 2.0f ? uint8_t(int(sv)) : uint8_t(0);
                if (write_dist) dist[r * w + c] = sv;
            }
        }
    };

    float prev[CW], cur[CW];
    init_row(ra, prev);
    if (ra >= r0) store_row(scratch + (ra - r0) * w, prev);
    // forward pass: rows ra+1 .. rb-1, columns 1..w-1; previous_row[w] aliases current_row[0]
    for (int r = ra + 1; r < rb; ++r) {
        init_row(r, cur);
        const float cur0 = __shfl_sync(SPX_FULL, cur[0], 0);
        chamfer_row_step<CW, true>(prev, cur, cur0, c0, w, lane);
        if (r >= r0) store_row(scratch + (r - r0) * w, cur);
#pragma unroll
        for (int k = 0; k < CW; ++k) prev[k] = cur[k];
    }
    // backward pass: rows rb-2 .. r0, columns w-2..0; `prev` holds the finished row below (row rb-1: forward values);
    // next_row[-1] aliases current_row[w-1]
    if (rb - 1 < r1) emit(rb - 1, prev);   // the frame's last row is never touched by the backward pass
    __syncwarp();
    const int lane_last = (w - 1) / CW, k_last = (w - 1) - lane_last * CW;
    for (int r = rb - 2; r >= r0; --r) {
        load_row_f(scratch + (r - r0) * w, cur);
        float lastv = 0.f;
#pragma unroll
        for (int k = 0; k < CW; ++k) if (k == k_last) lastv = cur[k];
        const float curLast = __shfl_sync(SPX_FULL, lastv, lane_last);
        chamfer_row_step<CW, false>(prev, cur, curLast, c0, w, lane);
        if (r < r1) emit(r, cur);
#pragma unroll
        for (int k = 0; k < CW; ++k) prev[k] = cur[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3: cloud + normals + plane_d + comparator links, one CTA per 32x16 tile of a frame.
//  1. back-projection of the tile plus a 7-pixel halo straight from the depth image into shared memory
//     (z = d; x = (n - cx) * z / fx; y = (m - cy) * z / fy, src/Frame.cc:861-865); the tile's own points go to HBM.
//  2. the central differences of initAverage3DGradientMethod (zero on the image border) are summed into tile-local
//     first-order integral images in fp64 (2 x 3 channels; a finite-count image only if the region holds a
//     non-finite depth), rows then columns.
//  3. every pixel of the tile AND of the column to its left / the row above takes its k x k window sum
//     (k = int(min(dist, 10)), rect [c - k/2, c - k/2 + k)) as ((I[y1][x1] + I[y0][x0]) - I[y0][x1]) - I[y1][x0],
//     forms n = normalize(gy x gx) in fp64, casts to fp32, flips it towards the origin, and plane_d = p . n.
//  4. PlaneCoefficientComparator::compare(current, left / upper): |d1 - d2| < DisTh * z1^2 && n1 . n2 > cos(AngTh)
//     -> 2 link bits per pixel + the row-run initialisation of the union-find forest.  Normals never reach HBM
//     (unless the parity taps are on).
// The fp64 sums of fp32 differences are exact for depth data (DESIGN.md "SAT exactness"), hence independent of the
// summation origin/order and bit-identical to PCL's whole-image double integral image.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 16;
constexpr int kNW = kTW + 1, kNH = kTH + 1;      // pixels that get a normal: the tile + one column left + one row up
constexpr int kHaloL = 5, kHaloR = 4;
constexpr int kRW = kNW + kHaloL + kHaloR;       // 42 columns of differences: image columns [tc-6, tc+35]
constexpr int kRH = kNH + kHaloL + kHaloR;       // 26 rows
constexpr int kCW = kRW + 2, kCHt = kRH + 2;     // 44 x 28 cloud points (differences reach +-1)
constexpr int kCStride = kCW + 1;
constexpr int kSW = kRW + 1, kSH = kRH + 1;      // 43 x 27 integral-image nodes
constexpr int kSATStride = kSW;                  // 43, odd
constexpr int kNStride = kNW + 1;
constexpr int kNormThreads = 288;                // 9 warps: the 33 x 17 = 561 normals take exactly two rounds
constexpr size_t kNormalsSmem = size_t(3) * kSH * kSATStride * sizeof(double) + size_t(3) * kCHt * kCStride * sizeof(float) +
                                size_t(kSH) * kSATStride * sizeof(int);

__device__ __forceinline__ void normals_tile(const float *__restrict__ depth, const Params &P, const Buffers &B, int write_normals, int f, int tc, int tr) {
    extern __shared__ double sm_d[];
    double *S = sm_d;                                                    // [3][kSH][kSATStride]: DX, then DY
    float *C = reinterpret_cast<float *>(S + 3 * kSH * kSATStride);      // [3][kCHt][kCStride]
    int *Cn = reinterpret_cast<int *>(C + 3 * kCHt * kCStride);          // [kSH][kSATStride], only with non-finite depth
    float *Nrm = reinterpret_cast<float *>(S);                           // [4][kNH][kNStride], aliases S after step 3
    constexpr int cs = kSH * kSATStride;          // channel stride of S
    constexpr int ccs = kCHt * kCStride;          // channel stride of C
    const int w = P.w, h = P.h;
    const size_t fo = size_t(f) * P.N;
    const int tid = threadIdx.x;
    const char *img = reinterpret_cast<const char *>(depth) + size_t(f) * P.samp_fstride;

    // ---- 1. cloud region: thread = one region column, 5 region rows per sweep (all loads of a thread in flight) ----
    int nonfinite = 0;
    unsigned zinv = 0u;
    if (tid < 5 * kCW) {
        const int lx = tid % kCW, ly0 = tid / kCW;
        const int c = tc - 7 + lx;
        const bool cin = c >= 0 && c < w;
        const float xfac = float(c * P.dis) - P.cx;
        const char *colp = img + size_t(cin ? c : 0) * P.samp_cstep * sizeof(float);
        const size_t rstep = P.samp_rstep;
        const bool cown = lx >= 7 && lx < 7 + kTW;
        float zz[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int ly = ly0 + 5 * i, r = tr - 7 + ly;
            zz[i] = 0.f;
            if (ly < kCHt && cin && r >= 0 && r < h) zz[i] = *reinterpret_cast<const float *>(colp + size_t(r) * rstep);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int ly = ly0 + 5 * i, r = tr - 7 + ly;
            if (ly < kCHt) {
                float x = 0.f, y = 0.f;
                const float z = zz[i];
                if (cin && r >= 0 && r < h) {
                    x = xfac * z / P.fx;
                    y = (float(r * P.dis) - P.cy) * z / P.fy;
                    if (!isfinite(z)) nonfinite = 1;
                    if (cown && ly >= 7 && ly < 7 + kTH) {
                        const size_t o = fo + size_t(r * w + c);
                        B.px[o] = x; B.py[o] = y; B.pz[o] = z;
                        const unsigned uz = __float_as_uint(z) & 0x7fffffffu;      // exactness bound (see k_normals_strip / k_models)
                        if (uz && uz < 0x7f800000u) zinv = max(zinv, 0x7f800000u - uz);
                    }
                }
                const int o = ly * kCStride + lx;
                C[o] = x; C[ccs + o] = y; C[2 * ccs + o] = z;
            }
        }
    }
    const int anynf = __syncthreads_or(nonfinite);
    zinv = __reduce_max_sync(SPX_FULL, zinv);
    if ((tid & 31) == 0 && zinv) atomicMax(&B.ctl[f].sat_zinv, zinv);

    // ---- per-pixel window geometry of the kNW x kNH pixels that get a normal (two per thread) ----
    const float qnan = __int_as_float(0x7fc00000);
    const uint8_t *kwin = B.kwin + fo;
    const int border = 10;
    constexpr int kPer = (kNW * kNH + kNormThreads - 1) / kNormThreads;
    int w_ul[kPer], w_k[kPer];          // upper-left integral-image node of the window, window size (0 = no window)
    bool inimg[kPer];
    float PX[kPer], PY[kPer], PZ[kPer];
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
        const int j = tid + it * kNormThreads;
        w_ul[it] = 0; w_k[it] = 0; inimg[it] = false; PX[it] = PY[it] = PZ[it] = 0.f;
        if (j < kNW * kNH) {
            const int ny_ = j / kNW, nx_ = j - ny_ * kNW;
            const int r = tr - 1 + ny_, c = tc - 1 + nx_;
            if (r >= 0 && c >= 0 && r < h && c < w) {
                inimg[it] = true;
                const int co = (ny_ + 6) * kCStride + nx_ + 6;
                PX[it] = C[co]; PY[it] = C[ccs + co]; PZ[it] = C[2 * ccs + co];
                if (r >= border && r < h - border && c >= border && c < w - border && isfinite(PZ[it])) {
                    const int k = int(kwin[r * w + c]);
                    if (k > 0) {
                        const int half = k / 2;
                        w_k[it] = k;
                        w_ul[it] = (ny_ + kHaloL - half) * kSATStride + (nx_ + kHaloL - half);   // integral-image coordinates
                    }
                }
            }
        }
    }

    // ---- 2. the two gradient images one after the other in the same 3-channel buffer: DX (pass 0), DY (pass 1).
    // difference (ly, lx) sits at image (tr-6+ly, tc-6+lx) = cloud index (ly+1, lx+1); differences are zero on the
    // image border: non-zero only for lx in [lo, hi) of rows with 1 <= r <= h-2
    double g[kPer][6];
    bool have[kPer];
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
        have[it] = w_k[it] > 0;
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) g[it][ch] = 0.0;
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        // zero row 0 and column 0 of the integral images
        for (int i = tid; i < 4 * kSW; i += kNormThreads) {
            const int ch = i / kSW, x = i - ch * kSW;
            if (ch < 3) S[ch * cs + x] = 0.0; else Cn[x] = 0;
        }
        for (int i = tid; i < 4 * kSH; i += kNormThreads) {
            const int ch = i / kSH, y = i - ch * kSH;
            if (ch < 3) S[ch * cs + y * kSATStride] = 0.0; else Cn[y * kSATStride] = 0;
        }
        // row prefix sums: one thread per (channel, row)
        if (!anynf) {
            if (tid < 3 * kRH) {
                const int ch = tid / kRH, ly = tid - ch * kRH;
                const float *Ck = C + ch * ccs;
                double *row = S + ch * cs + (ly + 1) * kSATStride;
                const int r = tr - 6 + ly;
                const bool rowok = r >= 1 && r <= h - 2;
                const int lo = rowok ? max(0, 7 - tc) : 0, hi = rowok ? min(kRW, w + 5 - tc) : 0;
                double run = 0.0;
                float sabs = 0.0f;                                     // sum of |differences| over the tile's own pixels
                const bool ownrow = ly >= 6 && ly < 6 + kTH;
                if (pass == 0) {
                    const float *p = Ck + (ly + 1) * kCStride;         // dx = P(r, c+1) - P(r, c-1)
                    float pa = p[0], pb = p[1];
#pragma unroll
                    for (int lx = 0; lx < kRW; ++lx) {
                        const float pc = p[lx + 2];
                        const float d = (lx >= lo && lx < hi) ? pc - pa : 0.0f;
                        pa = pb; pb = pc;
                        run += double(d);
                        if (lx >= 6 && lx < 6 + kTW) sabs += fabsf(d);
                        row[lx + 1] = run;
                    }
                } else {
                    const float *pu = Ck + ly * kCStride + 1, *pd = Ck + (ly + 2) * kCStride + 1;   // dy = P(r+1, c) - P(r-1, c)
#pragma unroll
                    for (int lx = 0; lx < kRW; ++lx) {
                        const float d = (lx >= lo && lx < hi) ? pd[lx] - pu[lx] : 0.0f;
                        run += double(d);
                        if (lx >= 6 && lx < 6 + kTW) sabs += fabsf(d);
                        row[lx + 1] = run;
                    }
                }
                if (ownrow && sabs > 0.0f) atomicAdd(&B.ctl[f].sat_sum[pass * 3 + ch], sabs);
            }
        } else {
            // generic path: a difference enters its image only if isfinite(d0 + (d1 + d2)); Cn counts the finite ones
            for (int i = tid; i < 4 * kRH; i += kNormThreads) {
                const int ch = i / kRH, ly = i - ch * kRH;
                const int r = tr - 6 + ly;
                double run = 0.0;
                int crun = 0;
                for (int lx = 0; lx < kRW; ++lx) {
                    const int c = tc - 6 + lx;
                    float d[3] = {0.f, 0.f, 0.f};
                    if (r >= 1 && r <= h - 2 && c >= 1 && c <= w - 2) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float *Ck = C + k * ccs;
                            d[k] = pass == 0 ? Ck[(ly + 1) * kCStride + lx + 2] - Ck[(ly + 1) * kCStride + lx]
                                             : Ck[(ly + 2) * kCStride + lx + 1] - Ck[ly * kCStride + lx + 1];
                        }
                    }
                    const bool fin = isfinite(d[0] + (d[1] + d[2]));
                    if (ch < 3) {
                        if (fin) run += double(d[ch]);
                        S[ch * cs + (ly + 1) * kSATStride + lx + 1] = run;
                    } else {
                        crun += fin ? 1 : 0;
                        Cn[(ly + 1) * kSATStride + lx + 1] = crun;
                    }
                }
            }
        }
        __syncthreads();
        // column prefix sums: one thread per (channel, column)
        for (int i = tid; i < (anynf ? 4 : 3) * kRW; i += kNormThreads) {
            const int ch = i / kRW, x = i - ch * kRW + 1;
            if (ch < 3) {
                double *col = S + ch * cs + x;
                double run = 0.0;
#pragma unroll
                for (int y = 1; y < kSH; ++y) { run += col[y * kSATStride]; col[y * kSATStride] = run; }
            } else {
                int *col = Cn + x;
                int run = 0;
                for (int y = 1; y < kSH; ++y) { run += col[y * kSATStride]; col[y * kSATStride] = run; }
            }
        }
        __syncthreads();
        // window sums of this gradient image: ((I[y1][x1] + I[y0][x0]) - I[y0][x1]) - I[y1][x0]
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            if (w_k[it] > 0) {
                const int k = w_k[it];
                const int ul = w_ul[it], ur = ul + k, ll = ul + k * kSATStride, lr = ll + k;
                if (anynf) { if (Cn[lr] + Cn[ul] - Cn[ur] - Cn[ll] == 0) have[it] = false; }
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const double *I = S + ch * cs;
                    g[it][pass * 3 + ch] = ((I[lr] + I[ul]) - I[ur]) - I[ll];
                }
            }
        }
        __syncthreads();   // the buffer is rebuilt by the next pass / re-used for the normals
    }

    // ---- 3. normals of the tile + left column + upper row (kNW x kNH pixels), kept in registers ----
    float rnx[kPer], rny[kPer], rnz[kPer], rpd[kPer];
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
        float nx = qnan, ny = qnan, nz = qnan, pd = qnan;
        if (inimg[it]) {
            const float X = PX[it], Y = PY[it], Zv = PZ[it];
            if (have[it]) {
                // normal_vector = gradient_y.cross(gradient_x)
                const double n0 = g[it][4] * g[it][2] - g[it][5] * g[it][1];
                const double n1 = g[it][5] * g[it][0] - g[it][3] * g[it][2];
                const double n2 = g[it][3] * g[it][1] - g[it][4] * g[it][0];
                const double len = (n0 * n0 + n1 * n1) + n2 * n2;
                if (len != 0.0) {
                    const double sl = sqrt(len);
                    nx = float(n0 / sl); ny = float(n1 / sl); nz = float(n2 / sl);
                    // flipNormalTowardsViewpoint(point, 0, 0, 0, nx, ny, nz)
                    const float vx = 0.0f - X, vy = 0.0f - Y, vz = 0.0f - Zv;
                    const float cos_theta = (vx * nx + vy * ny + vz * nz);
                    if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
                }
            }
            pd = dot3f(X, Y, Zv, nx, ny, nz);
        }
        rnx[it] = nx; rny[it] = ny; rnz[it] = nz; rpd[it] = pd;
    }
    // (the barrier closing the second pass: everybody is done with the integral images, their memory now holds the normals)
    constexpr int ncs = kNH * kNStride;
#pragma unroll
    for (int it = 0; it < kPer; ++it) {
        const int j = tid + it * kNormThreads;
        if (j < kNW * kNH) {
            const int ny_ = j / kNW, nx_ = j - ny_ * kNW;
            const int o = ny_ * kNStride + nx_;
            Nrm[o] = rnx[it]; Nrm[ncs + o] = rny[it]; Nrm[2 * ncs + o] = rnz[it]; Nrm[3 * ncs + o] = rpd[it];
        }
    }
    __syncthreads();

    // ---- 4. comparator links + row runs: warp = one tile row of 32 columns ----
    const int lane = tid & 31, wy = tid >> 5;
#pragma unroll
    for (int yy = 0; yy < kTH / 8; ++yy) {
        const int ty = wy + yy * 8;
        const int r = tr + ty, c = tc + lane;
        if (wy >= 8 || r >= h) continue;   // warp uniform (the ninth warp only helps with the normals)
        const bool valid = c < w;
        bool L = false, U = false;
        const int q = r * w + c;
        if (valid) {
            const int o = (ty + 1) * kNStride + lane + 1;
            const float d1 = Nrm[3 * ncs + o], n1x = Nrm[o], n1y = Nrm[ncs + o], n1z = Nrm[2 * ncs + o];
            const int co = (ty + 7) * kCStride + lane + 7;
            const float X = C[co], Y = C[ccs + co], Zv = C[2 * ccs + co];
            const float z = X * 0.0f + (Y * 0.0f + Zv * 1.0f);   // vec.dot(z_axis_)
            float threshold = P.dist_thr;
            threshold *= z * z;
            if (c >= 1) {
                const int ol = o - 1;
                L = (fabsf(d1 - Nrm[3 * ncs + ol]) < threshold) && (dot3f(n1x, n1y, n1z, Nrm[ol], Nrm[ncs + ol], Nrm[2 * ncs + ol]) > P.ang_cos);
            }
            if (r >= 1) {
                const int ou = o - kNStride;
                U = (fabsf(d1 - Nrm[3 * ncs + ou]) < threshold) && (dot3f(n1x, n1y, n1z, Nrm[ou], Nrm[ncs + ou], Nrm[2 * ncs + ou]) > P.ang_cos);
            }
            B.conn[fo + q] = uint8_t((L ? 1 : 0) | (U ? 2 : 0));
            B.cnt[fo + q] = isfinite(Zv) ? 0 : INT_MIN;   // a non-finite point: its (singleton) component size stays negative
            if (write_normals) { B.nx[fo + q] = n1x; B.ny[fo + q] = n1y; B.nz[fo + q] = n1z; B.pd[fo + q] = d1; }
        }
        const unsigned linked = __ballot_sync(SPX_FULL, valid && L);
        const unsigned starts = ~linked | 1u;                       // lane 0 always starts a run inside the segment
        const int s0 = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
        if (valid) B.parent[fo + q] = r * w + tc + s0;
    }
}

// one CTA per tile of every frame of the group (test knob SPX_NORMALS=0; the strip kernel is the production path)
__global__ void __launch_bounds__(kNormThreads) k_normals_link(const float *__restrict__ depth, Params P, Buffers B, int write_normals) {
    normals_tile(depth, P, B, write_normals, P.frame0 + blockIdx.z, blockIdx.x * kTW, blockIdx.y * kTH);
}

// the frames k_edge_chamfer queued because their depth holds NaN / Inf: a few persistent CTAs walk (frame, tile) items; with
// an empty queue -- every frame of a real sensor -- they leave at once
__global__ void __launch_bounds__(kNormThreads) k_normals_link_list(const float *__restrict__ depth, Params P, Buffers B, int write_normals) {
    const int n = B.nf_list[0];
    if (n == 0) return;
    const int tx = (P.w + kTW - 1) / kTW, ty = (P.h + kTH - 1) / kTH;
    for (int item = blockIdx.x; item < n * tx * ty; item += gridDim.x) {
        const int fi = item / (tx * ty), t = item - fi * (tx * ty);
        normals_tile(depth, P, B, write_normals, B.nf_list[1 + fi], (t % tx) * kTW, (t / tx) * kTH);
        __syncthreads();
    }
}

// cv::Mat::convertTo(CV_32F, alpha) of a CV_16U depth image (src/Tracking.cc:230-231): dst = float(src) * alpha, one
// rounding.  Four pixels per thread; the 16-bit image is tightly packed on the device.
__global__ void __launch_bounds__(256) k_convert_u16(const uint16_t *__restrict__ src, float *__restrict__ dst, size_t n4, float alpha) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const ushort4 v = reinterpret_cast<const ushort4 *>(src)[i];
    float4 o;
    o.x = float(v.x) * alpha; o.y = float(v.y) * alpha; o.z = float(v.z) * alpha; o.w = float(v.w) * alpha;
    reinterpret_cast<float4 *>(dst)[i] = o;
}

// the same for the sampled rows only (sparse upload): row m of frame f sits at its place in the full image, (f * rows + m * dis) * cols
__global__ void __launch_bounds__(128) k_convert_u16_rows(const uint16_t *__restrict__ src, float *__restrict__ dst, int cols, int rows, int h,
                                                          int dis, float alpha) {
    const int x4 = blockIdx.y * blockDim.x + threadIdx.x;   // rows on grid.x: frames * h exceeds the 65535 of grid.y
    if (x4 >= cols / 4) return;
    const int f = blockIdx.x / h, m = blockIdx.x - f * h;
    const size_t off = (size_t(f) * rows + size_t(m) * dis) * cols;
    const ushort4 v = reinterpret_cast<const ushort4 *>(src + off)[x4];
    float4 o;
    o.x = float(v.x) * alpha; o.y = float(v.y) * alpha; o.z = float(v.z) * alpha; o.w = float(v.w) * alpha;
    reinterpret_cast<float4 *>(dst + off)[x4] = o;
}

// plane_d for caller-supplied normals ("feed the reference's normals")
__global__ void __launch_bounds__(256) k_plane_d(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const size_t o = size_t(f) * P.N + i;
    B.pd[o] = dot3f(B.px[o], B.py[o], B.pz[o], B.nx[o], B.ny[o], B.nz[o]);
}

}  // namespace spx
