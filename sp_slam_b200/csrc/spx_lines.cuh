// spx_lines.cuh -- K8: supposed planes from plane edges.  RANSAC line fits on every kept plane's contour (seeded
// mt19937 draw sequence replayed exactly, hypotheses scored in parallel, adaptive stop replayed in order), PCA
// refinement, border tests on the full-resolution depth image, emission of perpendicular planes, and the offsets
// of every frame's results in the contiguous output buffers.
//
// Reference: /root/reference/src/Frame.cc:938-1114 (GeneratePlanesFromBoundries, IsBorderLine, IsBorderPoint,
// LineInRange, CaculatePlanes) and PCL 1.8.0 segmentation/impl/sac_segmentation.hpp (segment),
// sample_consensus/impl/ransac.hpp (computeModel), sample_consensus/sac_model.h (getSamples, drawIndexSample),
// sample_consensus/impl/sac_model_line.hpp, common/impl/centroid.hpp, filters/impl/extract_indices.hpp.
#pragma once
#include <climits>
#include "spx_math.cuh"
#include "spx_refine.cuh"   // plane_not_seen
#include "spx_types.cuh"

namespace spx {

constexpr int kLineThreads = 256;
constexpr int kLineCap = 2048;          // contours up to this many points live in shared memory

__constant__ uint32_t c_mt_state1[624]; // mt19937 state after seed(12345u) and the first twist
__constant__ uint32_t c_mt_out0[624];   // its first 624 tempered outputs
__constant__ float c_grid[64];          // the fp32 values visited by `for(float i=-0.25; i<0.25; i=i+0.01)`

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t m) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// in-place twist of the 624-word state by the whole CTA: words [0,227) depend on old words only, [227,454) on the
// new [0,227), [454,623) on the new [227,396), word 623 on the new words 0 and 396.
__device__ void mt_twist_cta(uint32_t *s) {
    const int tid = threadIdx.x;
    uint32_t v = 0;
    if (tid < 227) v = mt_mix(s[tid], s[tid + 1], s[tid + 397]);
    __syncthreads();
    if (tid < 227) s[tid] = v;
    __syncthreads();
    if (tid < 227) v = mt_mix(s[tid + 227], s[tid + 228], s[tid]);
    __syncthreads();
    if (tid < 227) s[tid + 227] = v;
    __syncthreads();
    if (tid < 169) v = mt_mix(s[tid + 454], s[tid + 455], s[tid + 227]);
    __syncthreads();
    if (tid < 169) s[tid + 454] = v;
    __syncthreads();
    if (tid == 0) s[623] = mt_mix(s[623], s[0], s[396]);
    __syncthreads();
}

struct Hyp {            // one RANSAC hypothesis
    int s0, s1;         // sample indices
    int state;          // 0 = valid model, 1 = no samples could be selected (the loop ends)
    unsigned skipped;   // skipped_count before this hypothesis' while-condition is evaluated
};

__device__ __forceinline__ void line_prep_dir(const float c[6], float dir[3]) {
    float d[4] = {c[3], c[4], c[5], 0.0f};
    const float z = dot4f(d, d);
    if (z > 0.0f) { const float s = sqrtf(z); d[0] /= s; d[1] /= s; d[2] /= s; }
    dir[0] = d[0]; dir[1] = d[1]; dir[2] = d[2];
}

// computeModelCoefficients from a sample pair (the caller has excluded the degenerate case)
__device__ __forceinline__ void line_model(const float4 a, const float4 b, float c[6]) {
    c[0] = a.x; c[1] = a.y; c[2] = a.z;
    float d0 = b.x - c[0], d1 = b.y - c[1], d2 = b.z - c[2];
    const float z = dot3f(d0, d1, d2, d0, d1, d2);
    if (z > 0.0f) { const float s = sqrtf(z); d0 /= s; d1 /= s; d2 /= s; }
    c[3] = d0; c[4] = d1; c[5] = d2;
}

// (line_pt - p).cross3(line_dir).squaredNorm() on Vector4f (w = 0), compared with the squared threshold in double.
// For a float sq and a double T, double(sq) < T  <=>  sq < Tf with Tf the smallest float >= T.
__device__ __forceinline__ bool line_within(const float c[6], const float dir[3], float X, float Y, float Z, float thr_f) {
    const float a0 = c[0] - X, a1 = c[1] - Y, a2 = c[2] - Z;
    const float x = a1 * dir[2] - a2 * dir[1];
    const float y = a2 * dir[0] - a0 * dir[2];
    const float z = a0 * dir[1] - a1 * dir[0];
    const float sq = (x * x + z * z) + y * y;   // Eigen's (a0 + a2) + (a1 + a3) with a3 = +0
    return sq < thr_f;
}

__device__ __forceinline__ float float_at_least(double t) {
    float f = float(t);
    if (double(f) < t) f = __int_as_float(__float_as_int(f) + 1);   // t > 0
    return f;
}

// ordered compaction: out[k] = i for the k-th i in [0, n) with pred(i); returns the count (all threads)
template <typename IdxT, typename Pred>
__device__ int block_select(int n, IdxT *out, int *s_warp, Pred pred) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = kLineThreads / 32;
    int total = 0;
    for (int base = 0; base < n; base += kLineThreads) {
        const int i = base + tid;
        const bool p = i < n && pred(i);
        const unsigned b = __ballot_sync(SPX_FULL, p);
        if (lane == 0) s_warp[wid] = __popc(b);
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int k = 0; k < NW; ++k) { const int t = s_warp[k]; if (k < wid) before += t; chunk += t; }
        if (p) out[total + before + __popc(b & ((1u << lane) - 1u))] = IdxT(i);
        total += chunk;
        __syncthreads();
    }
    return total;
}

// ---------------------------------------------------------------------------------------------------------------
// IsBorderLine / IsBorderPoint (src/Frame.cc:1013-1056) for every fitted line that passed LineInRange, after the
// fits (the border test does not feed back into the RANSAC rounds).  Out-of-buffer samples count as invalid,
// non-finite projections fail (arithmetic convention E7).
// A warp takes 32 points of a line, lane p owns point p.  The 20..21 x 20..21 depth window of a point is summed in
// the reference's raster order (fp32 running sum), so rows are walked in order: for window row t the warp loads row t
// of all 32 windows (lanes = window columns, one coalesced segment per point) into shared memory while every lane
// accumulates its own point's previous row -- the loads of 32 windows are in flight together instead of one
// dependent row at a time.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBorderWarps = 4;

// window of IsBorderPoint around a line point (src/Frame.cc:1030-1037): columns i0 .. while < u + 10, rows j0 .. while < v + 10
// A projection that is not finite (a line point at depth 0: u = fx * (+-0) * inf + cx = NaN) follows the reference's loops
// `for(int j = v - b; j < v + b; ++j) for(int i = u - b; i < u + b; ++i)`: with v NaN or -inf the outer loop never runs, with
// u NaN or -inf the inner one never does, res / num = 0 / 0 = NaN, `PcZ - NaN > 0.1` is false and IsBorderPoint returns TRUE
// (`empty`).  Only +inf, where the reference walks out of the image buffer, is answered "not a border point" (convention E7).
struct BorderWin { int i0, j0; float ue, ve; bool live, empty; };
__device__ __forceinline__ BorderWin border_window(const Params &P, const spx_point &p) {
    BorderWin W;
    W.i0 = 0; W.j0 = 0; W.ue = 0.f; W.ve = 0.f; W.live = false; W.empty = false;
    const int b = 10;
    const float PcZ = p.z;
    if (!(PcZ < 0.0f)) {
        const float invz = 1.0f / PcZ;
        const float u = P.fx * p.x * invz + P.cx;
        const float v = P.fy * p.y * invz + P.cy;
        const float pinf = __int_as_float(0x7f800000);
        if (isfinite(u) && isfinite(v)) {
            W.live = true;
            W.i0 = int(u - b); W.j0 = int(v - b);
            W.ue = u + b; W.ve = v + b;
        } else if (!isfinite(v)) {
            W.empty = v != pinf;
        } else {
            W.empty = u != pinf;
        }
    }
    return W;
}

// ---------------------------------------------------------------------------------------------------------------
// Sparse upload (host input, page-locked image): only the rows the organized cloud samples are copied to the device
// image (at their natural position).  This kernel brings in the rest of what the border tests will read: for every
// window row of every line point, the 8-pixel sectors it touches are claimed in a per-frame bitmap (atomicOr) and the
// claimer copies the sector from the caller's image (mapped host memory, read over PCIe) to the same place of the
// device image -- every sector crosses the bus once, all reads are independent (one round trip of latency), and
// k_border below then works on device memory.  DepthT = uint16_t: the raw CV_16U image, converted on the way
// (float(d) * mDepthMapFactor, as Tracking::GrabImageRGBD's convertTo).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fetch_convert8(const float *src, float *dst, int n, bool vec, float) {
    if (vec) {
        const float4 a = reinterpret_cast<const float4 *>(src)[0], b = reinterpret_cast<const float4 *>(src)[1];
        reinterpret_cast<float4 *>(dst)[0] = a; reinterpret_cast<float4 *>(dst)[1] = b;
    } else {
        for (int k = 0; k < n; ++k) dst[k] = src[k];
    }
}
__device__ __forceinline__ void fetch_convert8(const uint16_t *src, float *dst, int n, bool vec, float alpha) {
    if (vec) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src);
        float4 a, b;
        a.x = float(v.x & 0xffffu) * alpha; a.y = float(v.x >> 16) * alpha; a.z = float(v.y & 0xffffu) * alpha; a.w = float(v.y >> 16) * alpha;
        b.x = float(v.z & 0xffffu) * alpha; b.y = float(v.z >> 16) * alpha; b.z = float(v.w & 0xffffu) * alpha; b.w = float(v.w >> 16) * alpha;
        reinterpret_cast<float4 *>(dst)[0] = a; reinterpret_cast<float4 *>(dst)[1] = b;
    } else {
        for (int k = 0; k < n; ++k) dst[k] = float(src[k]) * alpha;
    }
}

template <typename DepthT>
__global__ void __launch_bounds__(256) k_border_fetch(const DepthT *__restrict__ host_img, float *__restrict__ dev_img, Params P, Buffers B,
                                                      unsigned long long *n_sectors) {
    const int tid = threadIdx.x;
    const int n_items = B.work2[0];
    const long long total = (long long)P.rows * P.cols;
    const int dev_pitch = int(P.pitch / sizeof(float));
    const int spr = (P.cols + 7) >> 3;                 // sectors per row
    const int wpr = (spr + 31) >> 5;                   // bitmap words per row
    const bool vec = P.fetch_vec != 0;
    unsigned mine_total = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int code = B.work2[2 + it];
        const int f = code / (SPX_MAX_MODELS * SPX_MAX_LINES), ls = code - f * (SPX_MAX_MODELS * SPX_MAX_LINES);
        const Line &L = B.ctl[f].lines[ls];
        const int n_pts = L.n_inliers;
        const spx_point *pts = B.line_pts + size_t(f) * P.contour_cap + L.pts_off;
        const DepthT *himg = reinterpret_cast<const DepthT *>(reinterpret_cast<const char *>(host_img) + size_t(f) * P.host_fstride);
        float *dimg = reinterpret_cast<float *>(reinterpret_cast<char *>(dev_img) + size_t(f) * P.frame_stride);
        unsigned *bits = B.fetch_bits + size_t(f) * size_t(P.rows) * wpr;
        // one (point, window row) per thread
        for (int idx = tid; idx < n_pts * 21; idx += blockDim.x) {
            const int pi = idx / 21, t = idx - pi * 21;
            const BorderWin W = border_window(P, pts[pi]);
            if (!W.live) continue;
            const int qj = W.j0 + t;
            if (!(qj < W.ve)) continue;
            int ncol = 0;
#pragma unroll
            for (int k = 0; k < 21; ++k) if (W.i0 + k < W.ue) ++ncol;
            // sectors [s0, s1] of image row r, columns [c_lo, c_hi): claim, then copy the claimed ones (loads first)
            auto fetch_row = [&](int r, int c_lo, int c_hi) {
                if (c_lo >= c_hi) return;
                if (P.fetch_skip_sampled && r % P.dis == 0) return;        // uploaded with the sampled rows
                const int s0 = c_lo >> 3, s1 = (c_hi - 1) >> 3;
                unsigned *brow = bits + size_t(r) * wpr;
                const DepthT *hrow = reinterpret_cast<const DepthT *>(reinterpret_cast<const char *>(himg) + size_t(r) * P.host_pitch);
                float *drow = dimg + size_t(r) * dev_pitch;
                for (int sb = s0; sb <= s1; sb += 4) {
                    bool mine[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int sct = sb + k;
                        mine[k] = false;
                        if (sct <= s1) {
                            const unsigned bit = 1u << (sct & 31);
                            mine[k] = (atomicOr(&brow[sct >> 5], bit) & bit) == 0u;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (mine[k]) {
                            const int c0 = (sb + k) << 3;
                            const int n = min(8, P.cols - c0);
                            fetch_convert8(hrow + c0, drow + c0, n, vec && n == 8, P.full_alpha);
                            ++mine_total;
                        }
                    }
                }
            };
            if (qj >= 0 && qj < P.rows) fetch_row(qj, max(W.i0, 0), min(W.i0 + ncol, P.cols));
            if (W.i0 < 0 || W.i0 + ncol > P.cols || qj < 0 || qj >= P.rows) {
                // samples outside the image are read through the flat index of the continuous cv::Mat (convention E7)
                for (int k = 0; k < ncol; ++k) {
                    const int i = W.i0 + k;
                    if (qj >= 0 && qj < P.rows && i >= 0 && i < P.cols) continue;
                    const long long fidx = (long long)qj * P.cols + i;
                    if (fidx >= 0 && fidx < total) {
                        const int rr = int(fidx / P.cols), cc = int(fidx - (long long)rr * P.cols);
                        fetch_row(rr, cc, cc + 1);
                    }
                }
            }
        }
    }
    if (n_sectors) {
        mine_total = __reduce_add_sync(SPX_FULL, mine_total);
        if ((tid & 31) == 0 && mine_total) atomicAdd(n_sectors, (unsigned long long)mine_total);
    }
}

// IsBorderPoint counts a window sample when `d > 0.05` (float against a double literal).  0.05 is not a float; 0.05f lies above it and
// the float below 0.05f lies below it, so the test is `d >= 0.05f` without the conversion.
constexpr float kMinDepthF = 0.05f;
static_assert(double(kMinDepthF) > 0.05 && double(0.049999997f) < 0.05, "0.05f must be the smallest float above the double 0.05");

__global__ void __launch_bounds__(kBorderWarps * 32) k_border(const float *__restrict__ depth, Params P, Buffers B) {
    __shared__ float s_win[kBorderWarps][2][32][21];
    __shared__ int4 s_geo[kBorderWarps][32];
    __shared__ int s_fail, s_item;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_items = B.work2[0];
    const int b = 10;
    const long long total = (long long)P.rows * P.cols;
    const int pitch_f = int(P.pitch / sizeof(float));
    while (true) {
        if (tid == 0) { s_item = atomicAdd(&B.work2[1], 1); s_fail = 0; }
        __syncthreads();
        const int it = s_item;
        if (it >= n_items) break;
        const int code = B.work2[2 + it];
        const int f = code / (SPX_MAX_MODELS * SPX_MAX_LINES), ls = code - f * (SPX_MAX_MODELS * SPX_MAX_LINES);
        Line &L = B.ctl[f].lines[ls];
        const int n_pts = L.n_inliers;
        const spx_point *pts = B.line_pts + size_t(f) * P.contour_cap + L.pts_off;
        const float *img = reinterpret_cast<const float *>(reinterpret_cast<const char *>(depth) + size_t(f) * P.frame_stride);
        int fails = 0;
        for (int base = wid * 32; base < n_pts; base += kBorderWarps * 32) {
            const int pi = base + lane;
            // per-point window geometry (lane = point)
            bool live = false;   // still needs its window walked
            bool result = false, forced = false;
            float PcZ = 0.f, ue = 0.f, ve = 0.f;
            int i0 = 0, j0 = 0;
            if (pi < n_pts) {
                const spx_point p = pts[pi];
                PcZ = p.z;
                const BorderWin W = border_window(P, p);
                live = W.live; forced = W.empty; i0 = W.i0; j0 = W.j0; ue = W.ue; ve = W.ve;
            }
            int num = 0, nan = 0;
            float res = 0.f;
            // window row t of point q: image row j0_q + t while that is < ve_q (at most 21 rows)
            // window geometry of the warp's 32 points, staged once: (i0, j0, columns, rows) of every window
            __syncwarp();
            {
                int ncol = 0, nrow = 0;
                if (live) {
                    for (int k = 0; k < 21; ++k) { if (i0 + k < ue) ++ncol; if (j0 + k < ve) ++nrow; }
                }
                // (bit 8 of the column count: the whole 21 x 21 window lies inside the image -- no bounds test per sample)
                const bool inside = live && i0 >= 0 && i0 + 20 < P.cols && j0 >= 0 && j0 + 20 < P.rows;
                s_geo[wid][lane] = make_int4(i0, j0, ncol | (inside ? 256 : 0), nrow);   // a dead point has an empty window
            }
            __syncwarp();
            auto load_row = [&](int t, int buf) {
                // 16 loads are issued before the first of them is stored, so their latencies overlap
#pragma unroll
                for (int q0 = 0; q0 < 32; q0 += 16) {
                    float d[16];
                    bool wr[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int4 gq = s_geo[wid][q0 + k];
                        const int qj = gq.y + t, i = gq.x + lane;
                        wr[k] = t < gq.w && lane < 21;
                        d[k] = 0.0f;   // invalid sample
                        if (wr[k] && lane < (gq.z & 255)) {
                            if ((gq.z & 256) || (qj >= 0 && qj < P.rows && i >= 0 && i < P.cols)) {
                                d[k] = img[qj * pitch_f + i];
                            } else {
                                const long long fidx = (long long)qj * P.cols + i;   // flat index on the continuous cv::Mat
                                if (fidx >= 0 && fidx < total) {
                                    const int rr = int(fidx / P.cols), cc = int(fidx - (long long)rr * P.cols);
                                    d[k] = img[size_t(rr) * pitch_f + cc];
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (wr[k]) s_win[wid][buf][q0 + k][lane] = d[k];
                }
            };
            load_row(0, 0);
            __syncwarp();
            for (int t = 0; t < 21; ++t) {
                if (t + 1 < 21) load_row(t + 1, (t + 1) & 1);
                if (live && (j0 + t < ve)) {
                    const float *row = s_win[wid][t & 1][lane];
#pragma unroll
                    for (int k = 0; k < 21; ++k) {
                        if (i0 + k < ue) {
                            const float d = row[k];
                            if (d >= kMinDepthF) { res += d; num++; }   // == (double(d) > 0.05): 0.05f is the smallest float above 0.05
                            else { nan++; }
                        }
                    }
                    if (nan > b * b) { live = false; result = false; }   // the reference returns false here
                }
                __syncwarp();
            }
            if (live) result = !(double(PcZ - res / num) > 0.1);
            if (forced) result = true;   // a window without samples: the reference's test on NaN is false, the point counts as border
            if (pi < n_pts && !result) ++fails;
        }
        fails = __reduce_add_sync(SPX_FULL, fails);
        if (lane == 0 && fails) atomicAdd(&s_fail, fails);
        __syncthreads();
        // IsBorderLine: fails as soon as more than s/4 points are not border points
        if (tid == 0) L.is_border = (s_fail > n_pts / 4) ? 0 : 1;
        __syncthreads();
    }
}

constexpr int kTripMax = 64;           // draw attempts per round trip (first trip: 32)

struct LineShared {
    uint32_t mt[624];
    uint16_t J[624];            // draw positions of the current output block (smem path: n <= kLineCap < 65536)
    int      Jg[624];           // the same for the global-memory path
    int      att0[kTripMax], att1[kTripMax];   // shuffled[0], shuffled[1] after each draw attempt of the trip
    Hyp      hyp[2][kTripMax];  // hypotheses of the trip being scored and of the next one (drawn meanwhile)
    int      counts[kTripMax];
    int      s_warp[kLineThreads / 32];
    float    stage[32 * 7];     // per-point terms of the centroid / covariance sums, staged for the ordered adds
    float    best[6], coef[6], dir[3], cen[3], cov[6];
    int      stop, iter, have, nhyp[2], natt, refill, jpos, item;
    int      n_best, iterations, best_s0, best_s1;
    unsigned skipped, run_bad;
    unsigned skipped_end[2];    // skipped_count after the attempts of the trip
    double   kk;
};

// One work item = one kept real plane with a contour of >= 50 points: the <= 4 segLine.segment() rounds on its contour
// (src/Frame.cc:953-997).  A = the working cloud (boundPoints), sh = RANSAC's shuffled index array, inl = inlier list.
template <typename IdxT, bool kSmem>
__device__ void lines_item(LineShared &S, float4 *A, IdxT *sh, IdxT *inl, const float *__restrict__ depth, const Params &P,
                           const Buffers &B, int f, int m) {
    FrameCtl &ctl = B.ctl[f];
    Model &M = ctl.models[m];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t fo = size_t(f) * P.N;
    const size_t co = size_t(f) * P.contour_cap + M.contour_off;
    spx_point *lpts = B.line_pts + co;
    const int boundSize = M.n_contour;
    const float thr_f = float_at_least(P.line_thr * P.line_thr);

    // boundPoints->points = mvBoundaryPoints[i].points
    {
        const int *cidx = B.contour_idx + co;
        for (int i = tid; i < boundSize; i += kLineThreads) {
            const int q = cidx[i];
            A[i] = make_float4(B.px[fo + q], B.py[fo + q], B.pz[fo + q], 0.0f);
        }
    }
    __syncthreads();

    int n = boundSize;
    int lp_off = 0;
    for (int round = 0; round < SPX_MAX_LINES; ++round) {
        Line &L = ctl.lines[m * SPX_MAX_LINES + round];
        // ---- RandomSampleConsensus::computeModel ----
        // a fresh model per segment(): mt19937 seeded with 12345, shuffled = identity
        for (int i = tid; i < 624; i += kLineThreads) {
            S.mt[i] = c_mt_state1[i];
            const unsigned i01 = unsigned(i & 1);
            const unsigned j = (n >= 2) ? i01 + (c_mt_out0[i] >> 1) % unsigned(n - int(i01)) : 0u;
            if (kSmem) S.J[i] = uint16_t(j); else S.Jg[i] = int(j);
        }
        for (int i = tid; i < n; i += kLineThreads) sh[i] = IdxT(i);
        if (tid == 0) { S.stop = 0; S.iter = 0; S.have = 0; S.jpos = 0; S.refill = 0; }
        __syncthreads();
        if (tid == 0) { S.n_best = -INT_MAX; S.iterations = 0; S.kk = 1.0; S.skipped = 0u; S.run_bad = 0u; }
        int v0 = 0, v1 = 1;   // shuffled[0], shuffled[1] live in registers of thread 0
        const unsigned max_skip = unsigned(P.ransac_max_iter) * 10u;
        const double log_probability = log(1.0 - 0.99);
        const double one_over_indices = 1.0 / double(n);
        int trip = 32;        // draw attempts of the next round trip (thread 0)
        __syncthreads();
        // One trip = a batch of draw attempts -> hypotheses -> scores -> replay of the loop condition.  Drawing is the only
        // inherently serial part (the swaps of drawIndexSample, thread 0) and does not depend on the scores, so warp 0 draws and
        // folds the NEXT trip while warps 1..7 score the current one; attempts drawn beyond the stop are simply not used (every
        // segment() round starts from a fresh shuffle and a fresh generator).
        auto draw_and_fold = [&](int buf) {          // warp 0
            Hyp *HB = S.hyp[buf];
            // (a) thread 0: the swaps of drawIndexSample for `trip` attempts -- the only inherently serial part
            if (tid == 0) {
                int na = 0, jp = S.jpos;
                if (n >= 2) {
                    while (na < trip && jp < 624) {
                    // two attempts at a time when their four positions are distinct and beyond the two sample slots (almost always):
                    // the four loads are independent of the four stores, so the swaps do not wait for one another
                    while (na + 1 < trip && jp + 3 < 624) {
                        const int a0 = kSmem ? int(S.J[jp]) : S.Jg[jp], a1 = kSmem ? int(S.J[jp + 1]) : S.Jg[jp + 1];
                        const int b0 = kSmem ? int(S.J[jp + 2]) : S.Jg[jp + 2], b1 = kSmem ? int(S.J[jp + 3]) : S.Jg[jp + 3];
                        if (a0 <= 1 || a1 <= 1 || b0 <= 1 || b1 <= 1 || a0 == a1 || a0 == b0 || a0 == b1 || a1 == b0 || a1 == b1 || b0 == b1) break;
                        const int ta0 = int(sh[a0]), ta1 = int(sh[a1]), tb0 = int(sh[b0]), tb1 = int(sh[b1]);
                        sh[a0] = IdxT(v0); sh[a1] = IdxT(v1);           // attempt 1: swap(shuffled[0], shuffled[a0]); swap(shuffled[1], shuffled[a1])
                        S.att0[na] = ta0; S.att1[na] = ta1;
                        sh[b0] = IdxT(ta0); sh[b1] = IdxT(ta1);         // attempt 2 swaps what attempt 1 left in slots 0 and 1
                        S.att0[na + 1] = tb0; S.att1[na + 1] = tb1;
                        v0 = tb0; v1 = tb1;
                        na += 2; jp += 4;
                    }
                    if (na < trip && jp < 624) {                       // one attempt the general way (then the pairs resume)
                        const int j0 = kSmem ? int(S.J[jp]) : S.Jg[jp];
                        const int j1 = kSmem ? int(S.J[jp + 1]) : S.Jg[jp + 1];
                        jp += 2;
                        // swap(shuffled[0], shuffled[j0]); swap(shuffled[1], shuffled[j1])
                        if (j0 == 1) { const int t = v0; v0 = v1; v1 = t; }
                        else if (j0 > 1) { const int t = int(sh[j0]); sh[j0] = IdxT(v0); v0 = t; }
                        if (j1 > 1) { const int t = int(sh[j1]); sh[j1] = IdxT(v1); v1 = t; }
                        S.att0[na] = v0; S.att1[na] = v1;
                        ++na;
                    }
                    }
                }
                S.natt = na; S.jpos = jp; S.refill = (n >= 2 && jp >= 624) ? 1 : 0;
            }
            __syncwarp();
            // (b) warp 0: isSampleGood / the degeneracy test of computeModelCoefficients for every attempt, then the
            // attempts are folded into hypotheses in order (getSamples gives up after 1000 bad draws in a row)
            if (wid == 0) {
                const int na = S.natt;
                unsigned good[2] = {0u, 0u}, degen[2] = {0u, 0u};
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int t = half * 32 + lane;
                    bool g = false, dg = false;
                    if (t < na) {
                        const float4 a = A[S.att0[t]], b = A[S.att1[t]];
                        g = (a.x != b.x) && (a.y != b.y) && (a.z != b.z);
                        dg = fabsf(a.x - b.x) <= FLT_EPSILON && fabsf(a.y - b.y) <= FLT_EPSILON && fabsf(a.z - b.z) <= FLT_EPSILON;
                    }
                    good[half] = __ballot_sync(SPX_FULL, g);
                    degen[half] = __ballot_sync(SPX_FULL, dg);
                }
                if (lane == 0) {
                    int nh = 0;
                    unsigned skipped = S.skipped, run_bad = S.run_bad;
                    bool ended = false;
                    if (n < 2) { HB[0].state = 1; HB[0].skipped = skipped; nh = 1; ended = true; }
                    for (int t = 0; t < na && !ended; ++t) {
                        const bool g = (good[t >> 5] >> (t & 31)) & 1u, dg = (degen[t >> 5] >> (t & 31)) & 1u;
                        if (!g) {
                            if (++run_bad >= 1000u) { HB[nh].state = 1; HB[nh].skipped = skipped; ++nh; ended = true; }
                            continue;
                        }
                        run_bad = 0u;
                        if (dg) {
                            ++skipped;
                            if (skipped >= max_skip) { HB[nh].state = 1; HB[nh].skipped = skipped; ++nh; ended = true; }
                            continue;
                        }
                        Hyp &H = HB[nh];
                        H.s0 = S.att0[t]; H.s1 = S.att1[t]; H.state = 0; H.skipped = skipped;
                        ++nh;
                    }
                    S.nhyp[buf] = nh; S.skipped = skipped; S.run_bad = run_bad; S.skipped_end[buf] = skipped;
                }
            }
            __syncwarp();
        };
        if (wid == 0) draw_and_fold(0);
        __syncthreads();
        for (int cur = 0;; cur ^= 1) {
            const int nh = S.nhyp[cur];
            const Hyp *HC = S.hyp[cur];
            if (wid == 0) {
                // the next trip (unless the generator block is used up: the refill below comes first)
                if (tid == 0) {
                    const double left = S.kk - double(S.iterations) - double(nh);
                    trip = left >= double(kTripMax) ? kTripMax : (left <= 8.0 ? 8 : int(left) + 1);
                }
                draw_and_fold(cur ^ 1);
            } else {
                // (c) score: warps 1..7 take hypotheses wid-1, wid-1+7, ...
                for (int hh = wid - 1; hh < nh; hh += kLineThreads / 32 - 1) {
                    if (HC[hh].state != 0) continue;
                    float c[6], dir[3];
                    line_model(A[HC[hh].s0], A[HC[hh].s1], c);
                    line_prep_dir(c, dir);
                    int cnt = 0;
                    for (int i = lane; i < n; i += 32) {
                        const float4 p = A[i];
                        cnt += line_within(c, dir, p.x, p.y, p.z, thr_f) ? 1 : 0;
                    }
                    cnt = __reduce_add_sync(SPX_FULL, cnt);
                    if (lane == 0) S.counts[hh] = cnt;
                }
            }
            __syncthreads();
            // (d) warp 0 replays `while (iterations_ < k && skipped_count < max_skip)` over the scored hypotheses:
            // k only depends on the running maximum of the counts, so prefix maxima give every hypothesis its k
            if (wid == 0) {
                int n_best = S.n_best, iterations = S.iterations;
                double kk = S.kk;
                int stop = 0, have_new = -1;
                for (int h0 = 0; h0 < nh && !stop; h0 += 32) {
                    const int hh = h0 + lane;
                    const bool valid = hh < nh;
                    const int st = valid ? HC[hh].state : 1;
                    const int cnt = (valid && st == 0) ? S.counts[hh] : -INT_MAX;
                    int run = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(SPX_FULL, run, o); if (lane >= o) run = max(run, t); }
                    run = max(run, n_best);                                   // n_best after this hypothesis
                    int prevmax = __shfl_up_sync(SPX_FULL, run, 1);
                    if (lane == 0) prevmax = n_best;
                    double k_after = kk;
                    if (run > n_best) {
                        const double wv = double(run) * one_over_indices;
                        double p_no_outliers = 1.0 - wv * wv;
                        p_no_outliers = fmax(DBL_EPSILON, p_no_outliers);
                        p_no_outliers = fmin(1.0 - DBL_EPSILON, p_no_outliers);
                        k_after = log_probability / log(p_no_outliers);
                    }
                    double k_before = __shfl_up_sync(SPX_FULL, k_after, 1);
                    if (lane == 0) k_before = kk;
                    const int it_before = iterations + lane;
                    const bool enter = valid && (double(it_before) < k_before) && (HC[valid ? hh : 0].skipped < max_skip) && st == 0;
                    const bool brk_after = enter && (it_before + 1 > P.ransac_max_iter);
                    const unsigned m_noenter = __ballot_sync(SPX_FULL, valid && !enter);
                    const unsigned m_brk = __ballot_sync(SPX_FULL, brk_after);
                    int n_exec = valid ? 32 : 0;
                    n_exec = __popc(__ballot_sync(SPX_FULL, valid));
                    if (m_noenter) { n_exec = min(n_exec, __ffs(m_noenter) - 1); }
                    if (m_brk) { n_exec = min(n_exec, __ffs(m_brk)); }
                    if (m_noenter || m_brk) stop = 1;
                    // the last executed hypothesis that raised the maximum carries the best model
                    const unsigned m_impr = __ballot_sync(SPX_FULL, valid && lane < n_exec && cnt > prevmax);
                    if (m_impr) have_new = h0 + (31 - __clz(m_impr));
                    if (n_exec > 0) {
                        n_best = __shfl_sync(SPX_FULL, run, n_exec - 1);
                        kk = __shfl_sync(SPX_FULL, k_after, n_exec - 1);
                        iterations += n_exec;
                    }
                }
                if (lane == 0) {
                    if (!stop && !(double(iterations) < kk && S.skipped_end[cur] < max_skip)) stop = 1;
                    S.n_best = n_best; S.iterations = iterations; S.kk = kk;
                    if (have_new >= 0) { S.have = 1; S.best_s0 = HC[have_new].s0; S.best_s1 = HC[have_new].s1; }
                    S.stop = stop; S.iter = iterations;
                }
            }
            __syncthreads();
            if (S.stop) break;
            if (S.refill) {   // the 624 outputs of the block are used up: twist and re-derive the draw positions
                mt_twist_cta(S.mt);
                for (int i = tid; i < 624; i += kLineThreads) {
                    const unsigned i01 = unsigned(i & 1);
                    const unsigned j = i01 + (mt_temper(S.mt[i]) >> 1) % unsigned(n - int(i01));
                    if (kSmem) S.J[i] = uint16_t(j); else S.Jg[i] = int(j);
                }
                if (tid == 0) { S.jpos = 0; S.refill = 0; }
                __syncthreads();
            }
        }
        if (tid == 0 && S.have) { float mc[6]; line_model(A[S.best_s0], A[S.best_s1], mc); for (int q = 0; q < 6; ++q) S.best[q] = mc[q]; }
        __syncthreads();
        const bool have = S.have != 0;
        int n_inl = 0;
        if (have) {
            // sac_->getInliers: selectWithinDistance(best)
            if (tid == 0) { for (int q = 0; q < 6; ++q) S.coef[q] = S.best[q]; line_prep_dir(S.best, S.dir); }
            __syncthreads();
            n_inl = block_select(n, inl, S.s_warp, [&](int i) { const float4 p = A[i]; return line_within(S.coef, S.dir, p.x, p.y, p.z, thr_f); });
            __syncthreads();
            // optimizeModelCoefficients: centroid + principal direction of the inliers.  compute3DCentroid and
            // computeCovarianceMatrix add the inliers in list order (fp32): warp 0 gathers 32 inliers at a time
            // (coalesced), stages their terms in shared memory, and lanes 0..2 / 0..5 carry the ordered sums.
            if (n_inl > 2) {
                if (wid == 0) {
                    float acc = 0.0f;
                    for (int base = 0; base < n_inl; base += 32) {
                        const int k = base + lane;
                        if (k < n_inl) { const float4 p = A[inl[k]]; S.stage[lane * 7 + 0] = p.x; S.stage[lane * 7 + 1] = p.y; S.stage[lane * 7 + 2] = p.z; }
                        __syncwarp();
                        const int cnt = min(32, n_inl - base);
                        if (lane < 3) for (int j = 0; j < cnt; ++j) acc += S.stage[j * 7 + lane];
                        __syncwarp();
                    }
                    if (lane < 3) S.cen[lane] = acc / float(n_inl);
                    __syncwarp();
                    const float c0 = S.cen[0], c1 = S.cen[1], c2 = S.cen[2];
                    acc = 0.0f;
                    for (int base = 0; base < n_inl; base += 32) {
                        const int k = base + lane;
                        if (k < n_inl) {
                            const float4 p = A[inl[k]];
                            const float ptx = p.x - c0, pty = p.y - c1, ptz = p.z - c2;
                            float *st = S.stage + lane * 7;
                            st[0] = ptx * ptx; st[1] = pty * ptx; st[2] = ptz * ptx;   // (0,0) (0,1) (0,2)
                            st[3] = pty * pty; st[4] = pty * ptz; st[5] = ptz * ptz;   // (1,1) (1,2) (2,2)
                        }
                        __syncwarp();
                        const int cnt = min(32, n_inl - base);
                        if (lane < 6) for (int j = 0; j < cnt; ++j) acc += S.stage[j * 7 + lane];
                        __syncwarp();
                    }
                    if (lane < 6) S.cov[lane] = acc;
                    __syncwarp();
                    if (lane == 0) {
                        const float cov[9] = {S.cov[0], S.cov[1], S.cov[2], S.cov[1], S.cov[3], S.cov[4], S.cov[2], S.cov[4], S.cov[5]};
                        float vec[3];
                        eigen33_largest_vec(cov, vec);
                        S.coef[0] = S.cen[0]; S.coef[1] = S.cen[1]; S.coef[2] = S.cen[2];
                        S.coef[3] = vec[0]; S.coef[4] = vec[1]; S.coef[5] = vec[2];
                        line_prep_dir(S.coef, S.dir);
                    }
                }
                __syncthreads();
            }
            // refine inliers with the optimised coefficients
            n_inl = block_select(n, inl, S.s_warp, [&](int i) { const float4 p = A[i]; return line_within(S.coef, S.dir, p.x, p.y, p.z, thr_f); });
            __syncthreads();
        }
        // record
        const bool stop_round = double(n_inl) < P.line_ratio * double(boundSize);
        int in_range = 0;
        if (!stop_round) {
            // LineInRange (src/Frame.cc:1058-1074) on the line point (= inlier centroid)
            {
                const float PcX = S.coef[0], PcY = S.coef[1], PcZ = S.coef[2];
                if (!(PcZ < 0.0f)) {
                    const float invz = 1.0f / PcZ;
                    const float u = P.fx * PcX * invz + P.cx;
                    const float v = P.fy * PcY * invz + P.cy;
                    in_range = 1;
                    if (u < (P.min_x + 50) || u > (P.max_x - 50)) in_range = 0;
                    if (v < (P.min_y + 50) || v > (P.max_y - 50)) in_range = 0;
                }
            }
            // linePoints (ExtractIndices positive) -> line arena, coloured red for a possible supposed plane
            for (int k = tid; k < n_inl; k += kLineThreads) {
                const float4 p = A[inl[k]];
                spx_point q; q.x = p.x; q.y = p.y; q.z = p.z; q.rgba = pack_rgba(255, 0, 0);
                lpts[lp_off + k] = q;
            }
        }
        if (tid == 0) {
            L.model = m; L.round = round; L.n_points = n; L.iterations = S.iterations; L.n_inliers = n_inl;
            L.in_range = in_range; L.is_border = 0; L.emitted = 0; L.pts_off = M.contour_off + lp_off;
            // the border test of an in-range line runs in k_border
            if (in_range) B.work2[2 + atomicAdd(&B.work2[0], 1)] = (f * SPX_MAX_MODELS + m) * SPX_MAX_LINES + round;
            for (int q = 0; q < 6; ++q) L.coef[q] = have ? S.coef[q] : 0.0f;
            M.n_rounds = round + 1;
        }
        if (stop_round) break;
        // ExtractIndices negative: the remaining cloud keeps its order (in-place, chunk by chunk: sources never
        // precede their destination)
        {
            const int n_rem = block_select(n, inl, S.s_warp, [&](int i) { const float4 p = A[i]; return !line_within(S.coef, S.dir, p.x, p.y, p.z, thr_f); });
            __syncthreads();
            for (int base = 0; base < n_rem; base += kLineThreads) {
                const int k = base + tid;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < n_rem) v = A[inl[k]];
                __syncthreads();
                if (k < n_rem) A[k] = v;
            }
            lp_off += n_inl;
            n = n_rem;
        }
        __syncthreads();
    }
    __syncthreads();
}

constexpr size_t kLinesSmem = sizeof(float4) * kLineCap + 2 * sizeof(uint16_t) * kLineCap;

// persistent CTAs pull (frame, model) items from the queue k_postfilter filled
__global__ void __launch_bounds__(kLineThreads) k_lines(const float *__restrict__ depth, Params P, Buffers B) {
    extern __shared__ float4 sm_pts[];
    __shared__ LineShared S;
    float4 *A = sm_pts;
    uint16_t *sh = reinterpret_cast<uint16_t *>(A + kLineCap);
    uint16_t *inl = sh + kLineCap;
    const int n_items = B.work[0];
    while (true) {
        if (threadIdx.x == 0) S.item = atomicAdd(&B.work[1], 1);
        __syncthreads();
        const int it = S.item;
        if (it >= n_items) break;
        const int code = B.work[2 + it];
        const int f = code / SPX_MAX_MODELS, m = code - f * SPX_MAX_MODELS;
        const Model &M = B.ctl[f].models[m];
        if (M.n_contour <= kLineCap && !P.lines_in_global) {
            lines_item<uint16_t, true>(S, A, sh, inl, depth, P, B, f, m);
        } else {
            const size_t co = size_t(f) * P.contour_cap + M.contour_off;
            lines_item<int, false>(S, B.line_a + co, B.line_sh + co, B.line_inl + co, depth, P, B, f, m);
        }
    }
}

// CaculatePlanes + the serial control flow of GeneratePlanesFromBoundries: one thread per frame
__global__ void __launch_bounds__(128) k_supposed(Params P, Buffers B) {
    const int fl = blockIdx.x * blockDim.x + threadIdx.x;
    if (fl >= P.n_frames) return;
    const int f = P.frame0 + fl;
    FrameCtl &ctl = B.ctl[f];
    // offsets of a supposed plane's clouds are relative to the frame's supposed-plane region of the arenas
    int np = ctl.n_real, poff = 0, boff = 0;
    const int preal = ctl.pts_used, breal = ctl.bnd_used;
    for (int i = ctl.n_real - 1; i >= 0; --i) {
        const int m = ctl.planes[i].src;
        const Model &M = ctl.models[m];
        const int boundSize = M.n_contour;
        if (boundSize < 50) continue;
        for (int j = 0; j < M.n_rounds; ++j) {
            Line &L = ctl.lines[m * SPX_MAX_LINES + j];
            if (double(L.n_inliers) < P.line_ratio * double(boundSize)) break;
            if (!(L.in_range && L.is_border)) continue;
            const float *p = ctl.planes[i].coef;
            const float *l = L.coef;
            float a, b, c, d;
            a = p[1] * l[5] - p[2] * l[4];
            b = p[2] * l[3] - p[0] * l[5];
            c = p[0] * l[4] - p[1] * l[3];
            d = a * l[0] + b * l[1] + c * l[2];
            const float v = sqrtf(a * a + b * b + c * c);
            float coef[4] = {a / v, b / v, c / v, -d / v};
            if (coef[3] < 0) { coef[0] = -coef[0]; coef[1] = -coef[1]; coef[2] = -coef[2]; coef[3] = -coef[3]; }
            if (!plane_not_seen(ctl, np, coef)) continue;
            const int npts = P.n_grid * P.n_grid + L.n_inliers;
            if (np >= SPX_MAX_PLANES || preal + poff + npts > P.pts_cap || breal + boff + L.n_inliers > P.bnd_cap) { ctl.flags |= unsigned(SPX_FRAME_OVERFLOW); continue; }
            PlaneRec &R = ctl.planes[np];
            R.coef[0] = coef[0]; R.coef[1] = coef[1]; R.coef[2] = coef[2]; R.coef[3] = coef[3];
            R.n_points = npts; R.n_boundary = L.n_inliers; R.points_off = poff; R.boundary_off = boff;
            R.src = i; R.is_supposed = 1; R.line = m * SPX_MAX_LINES + j; R.pad = 0;
            L.emitted = 1;
            poff += npts; boff += L.n_inliers;
            ++np;
        }
    }
    ctl.n_planes = np; ctl.pts_sup = poff; ctl.bnd_sup = boff;
}

// clouds of the supposed planes (src/Frame.cc:1092-1110, 983-989) and the every-20th-inlier boundary fallback
// (src/Frame.cc:1001-1011); one CTA per (plane, frame)
__global__ void __launch_bounds__(128) k_pack_supposed(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y, k = blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    if (k >= ctl.n_planes) return;
    const PlaneRec &R = ctl.planes[k];
    spx_point *pts = B.out_pts + B.frame_offs[size_t(f) * 5 + (R.is_supposed ? 3 : 1)] + R.points_off;
    spx_point *bnd = B.out_bnd + B.frame_offs[size_t(f) * 5 + (R.is_supposed ? 4 : 2)] + R.boundary_off;
    if (!R.is_supposed) {
        if (ctl.models[R.src].n_contour == 0 && P.enable_supposed) {
            const size_t fo = size_t(f) * P.N;
            const long long io = B.frame_offs[size_t(f) * 5 + 1] + R.points_off;   // compact results: the plane's index list
            for (int j = threadIdx.x; j < R.n_boundary; j += blockDim.x) {
                spx_point q;
                if (P.compact) {
                    const int ix = P.idx16 ? int(static_cast<const uint16_t *>(B.out_pidx)[io + j * 20]) : int(static_cast<const uint32_t *>(B.out_pidx)[io + j * 20]);
                    q.x = B.px[fo + ix]; q.y = B.py[fo + ix]; q.z = B.pz[fo + ix];
                } else {
                    q = pts[j * 20];
                }
                q.rgba = pack_rgba(0, 0, 0);
                bnd[j] = q;
            }
        }
        return;
    }
    const Line &L = ctl.lines[R.line];
    const float *l = L.coef;
    const float *p = ctl.planes[R.src].coef;
    const int ng = P.n_grid;
    for (int t = threadIdx.x; t < ng * ng; t += blockDim.x) {
        const float i = c_grid[t / ng], j = c_grid[t % ng];
        spx_point q;
        q.x = l[0] + i * l[3] + j * p[0];
        q.y = l[1] + i * l[4] + j * p[1];
        q.z = (R.coef[0] * q.x + R.coef[1] * q.y + R.coef[3]) / (-R.coef[2]);
        q.rgba = pack_rgba(0, 255, 0);
        pts[t] = q;
    }
    const spx_point *lp = B.line_pts + size_t(f) * P.contour_cap + L.pts_off;
    for (int t = threadIdx.x; t < L.n_inliers; t += blockDim.x) {
        const spx_point q = lp[t];
        pts[ng * ng + t] = q;
        bnd[t] = q;
    }
}

// exclusive scan of the per-frame totals of a frame range (one CTA): where each frame's plane records and clouds
// start in the output buffers (base_* = where the range's results start).  stage 0, right after k_postfilter: the
// real planes' points / boundary points (their packing then overlaps the line fits); stage 1, after k_supposed: the
// plane records and the supposed planes' clouds, which follow the real part of the range.
// tot: 8 values, [0..2] final totals (planes, points, boundary points), [5], [6] the real part of [1], [2].
// host_tot (optional): the same slot in page-locked host memory; the totals are stored there directly (over PCIe) so that
// no device-to-host copy -- which would queue behind the other groups' result downloads in the copy engine -- sits on
// the group's compute stream.
__global__ void __launch_bounds__(1024) k_scan_frames(Params P, Buffers B, int stage, long long base_pl, long long base_pt, long long base_bd,
                                                        long long *tot, long long *host_tot) {
    __shared__ long long s_run[3];
    __shared__ long long s_w[3][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // (compact results: the real planes' clouds live in the index arena, the point arena holds the supposed planes only)
    const long long real_pt = (stage && !P.compact) ? tot[5] : 0, real_bd = stage ? tot[6] : 0;
    if (tid < 3) s_run[tid] = 0;
    __syncthreads();
    for (int base = 0; base < P.n_frames; base += 1024) {
        const int f = base + tid;
        long long v[3] = {0, 0, 0};
        if (f < P.n_frames) {
            const FrameCtl &K = B.ctl[P.frame0 + f];
            if (stage == 0) { v[1] = K.pts_used; v[2] = K.bnd_used; }
            else { v[0] = K.n_planes; v[1] = K.pts_sup; v[2] = K.bnd_sup; }
        }
        long long inc[3];
        for (int k = 0; k < 3; ++k) {
            long long x = v[k];
            for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(SPX_FULL, x, o); if (lane >= o) x += t; }
            inc[k] = x;
            if (lane == 31) s_w[k][wid] = x;
        }
        __syncthreads();
        if (wid == 0) {
            for (int k = 0; k < 3; ++k) {
                const long long t = s_w[k][lane];
                long long x = t;
                for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(SPX_FULL, x, o); if (lane >= o) x += u; }
                s_w[k][lane] = x - t;
            }
        }
        __syncthreads();
        long long totl[3];
        for (int k = 0; k < 3; ++k) {
            const long long excl = s_run[k] + s_w[k][wid] + inc[k] - v[k];
            if (f < P.n_frames) {
                long long *fo = B.frame_offs + size_t(P.frame0 + f) * 5;
                if (stage == 0) { if (k == 1) fo[1] = base_pt + excl; if (k == 2) fo[2] = base_bd + excl; }
                else { if (k == 0) fo[0] = base_pl + excl; if (k == 1) fo[3] = base_pt + real_pt + excl; if (k == 2) fo[4] = base_bd + real_bd + excl; }
            }
            totl[k] = excl + v[k];
        }
        __syncthreads();
        if (tid == 1023) for (int k = 0; k < 3; ++k) s_run[k] = totl[k];
        __syncthreads();
    }
    if (tid == 0) {
        if (stage == 0) {
            tot[5] = s_run[1]; tot[6] = s_run[2];
            if (host_tot) { host_tot[5] = s_run[1]; host_tot[6] = s_run[2]; __threadfence_system(); }
        } else {
            tot[0] = s_run[0]; tot[1] = real_pt + s_run[1]; tot[2] = real_bd + s_run[2];
            if (host_tot) { host_tot[0] = tot[0]; host_tot[1] = tot[1]; host_tot[2] = tot[2]; host_tot[3] = tot[3]; __threadfence_system(); }
        }
    }
}

// frame headers and plane records with batch-global offsets; one CTA per frame
__global__ void __launch_bounds__(128) k_emit_records(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    const long long *fo5 = B.frame_offs + size_t(f) * 5;
    const long long o_pl = fo5[0];
    if (threadIdx.x == 0) {
        spx_frame_header h;
        h.n_real = ctl.n_real; h.n_planes = ctl.n_planes; h.first_plane = int(o_pl); h.flags = ctl.flags;
        B.out_frames[f] = h;
    }
    for (int k = threadIdx.x; k < ctl.n_planes; k += blockDim.x) {
        const PlaneRec &R = ctl.planes[k];
        spx_plane o;
        o.coef[0] = R.coef[0]; o.coef[1] = R.coef[1]; o.coef[2] = R.coef[2]; o.coef[3] = R.coef[3];
        o.n_points = R.n_points; o.n_boundary = R.n_boundary;
        o.points_off = fo5[R.is_supposed ? 3 : 1] + R.points_off; o.boundary_off = fo5[R.is_supposed ? 4 : 2] + R.boundary_off;
        o.src = R.src; o.is_supposed = R.is_supposed;
        B.out_planes[o_pl + k] = o;
    }
}

}  // namespace spx
