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
constexpr int kHyp = 32;                // hypotheses scored per round trip
constexpr int kShuffleSmem = 6144;      // contours up to this many points keep the shuffled index array in shared memory

__constant__ uint32_t c_mt_init[624];   // mt19937 state after seed(12345u)
__constant__ float c_grid[64];          // the fp32 values visited by `for(float i=-0.25; i<0.25; i=i+0.01)`

struct Mt {
    uint32_t s[624];
    int idx;
};

__device__ __forceinline__ uint32_t mt_next(Mt &g) {
    if (g.idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (g.s[i] & 0x80000000u) | (g.s[(i + 1) % 624] & 0x7fffffffu);
            g.s[i] = g.s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g.idx = 0;
    }
    uint32_t y = g.s[g.idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

struct Hyp {            // one RANSAC hypothesis: point + direction after countWithinDistance's second normalisation
    float c[6];         // model_coefficients as computeModelCoefficients leaves them
    float dir[3];       // line_dir.normalize() of the Vector4f (w = 0)
    int   state;        // 0 = valid model, 1 = no samples could be selected (loop ends)
    unsigned skipped;   // skipped_count before this hypothesis' while-condition is evaluated
};

__device__ __forceinline__ void line_prep_dir(const float c[6], float dir[3]) {
    float d[4] = {c[3], c[4], c[5], 0.0f};
    const float z = dot4f(d, d);
    if (z > 0.0f) { const float s = sqrtf(z); d[0] /= s; d[1] /= s; d[2] /= s; }
    dir[0] = d[0]; dir[1] = d[1]; dir[2] = d[2];
}

// (line_pt - p).cross3(line_dir).squaredNorm() on Vector4f, compared in double
__device__ __forceinline__ bool line_within(const float c[6], const float dir[3], float X, float Y, float Z, double sqr_thr) {
    const float a0 = c[0] - X, a1 = c[1] - Y, a2 = c[2] - Z;
    const float x = a1 * dir[2] - a2 * dir[1];
    const float y = a2 * dir[0] - a0 * dir[2];
    const float z = a0 * dir[1] - a1 * dir[0];
    const float sq = (x * x + z * z) + (y * y + 0.0f);
    return double(sq) < sqr_thr;
}

// ordered compaction: out[k] = i for the k-th i in [0, n) with pred(i); returns the count (all threads)
template <typename Pred>
__device__ int block_select(int n, int *out, int *s_warp, Pred pred) {
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
        if (p) out[total + before + __popc(b & ((1u << lane) - 1u))] = i;
        total += chunk;
        __syncthreads();
    }
    return total;
}

// IsBorderPoint (src/Frame.cc:1026-1056); out-of-buffer samples count as invalid, non-finite projections fail
__device__ bool is_border_point(const Params &P, const float *img, float PcX, float PcY, float PcZ) {
    if (PcZ < 0.0f) return false;
    const float invz = 1.0f / PcZ;
    const float u = P.fx * PcX * invz + P.cx;
    const float v = P.fy * PcY * invz + P.cy;
    if (!isfinite(u) || !isfinite(v)) return false;
    int num = 0, nan = 0;
    float res = 0;
    const int b = 10;
    const long long total = (long long)P.rows * P.cols;
    const int pitch_f = int(P.pitch / sizeof(float));
    bool bail = false;
    for (int j = int(v - b); j < v + b && !bail; ++j) {
        for (int i = int(u - b); i < u + b; ++i) {
            const long long fidx = (long long)j * P.cols + i;   // flat index on the continuous cv::Mat
            float d = 0.0f;
            const bool inside = fidx >= 0 && fidx < total;
            if (inside) { const int rr = int(fidx / P.cols), cc = int(fidx - (long long)rr * P.cols); d = img[size_t(rr) * pitch_f + cc]; }
            if (inside && double(d) > 0.05) {
                res += d;
                num++;
            } else {
                nan++;
                if (nan > b * b) { bail = true; break; }
            }
        }
    }
    if (bail) return false;
    if (double(PcZ - res / num) > 0.1) return false;
    return true;
}

// one CTA per (model, frame): the <= 4 segLine.segment() rounds on the model's contour (src/Frame.cc:953-997)
__global__ void __launch_bounds__(kLineThreads) k_lines(const float *__restrict__ depth, Params P, Buffers B) {
    __shared__ Mt mt;
    __shared__ Hyp hyp[kHyp];
    __shared__ int counts[kHyp];
    __shared__ int s_warp[kLineThreads / 32];
    __shared__ int sh_smem[kShuffleSmem];
    __shared__ float s_best[6];
    __shared__ float s_coef[6];
    __shared__ float s_dir[3];
    __shared__ float s_cen[3];
    __shared__ float s_cov[6];
    __shared__ int s_stop, s_iter, s_have, s_nhyp, s_fail;

    const int f = blockIdx.y, m = blockIdx.x;
    FrameCtl &ctl = B.ctl[f];
    if (m >= ctl.n_models) return;
    Model &M = ctl.models[m];
    if (threadIdx.x == 0) M.n_rounds = 0;
    if (M.plane < 0 || M.n_contour < 50) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t fo = size_t(f) * P.N;
    const size_t co = size_t(f) * P.contour_cap + M.contour_off;
    float4 *A = B.line_a + co, *Bb = B.line_b + co;
    int *inl = B.line_inl + co;
    int *sh = (M.n_contour <= kShuffleSmem) ? sh_smem : (B.line_sh + co);
    spx_point *lpts = B.line_pts + co;
    const float *img = reinterpret_cast<const float *>(reinterpret_cast<const char *>(depth) + size_t(f) * P.frame_stride);
    const int boundSize = M.n_contour;
    const double sqr_thr = P.line_thr * P.line_thr;

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
        for (int i = tid; i < 624; i += kLineThreads) mt.s[i] = c_mt_init[i];
        for (int i = tid; i < n; i += kLineThreads) sh[i] = i;
        if (tid == 0) { mt.idx = 624; s_stop = 0; s_iter = 0; s_have = 0; }
        __syncthreads();
        // thread-0 serial state of the RANSAC loop
        int iterations = 0, n_best = -INT_MAX;
        double kk = 1.0;
        unsigned skipped = 0;
        const unsigned max_skip = unsigned(P.ransac_max_iter) * 10u;
        const double log_probability = log(1.0 - 0.99);
        const double one_over_indices = 1.0 / double(n);
        while (true) {
            if (tid == 0) {
                // draw the next kHyp hypotheses exactly as getSamples / drawIndexSample / isSampleGood /
                // computeModelCoefficients would, in sequence
                int nh = 0;
                while (nh < kHyp) {
                    Hyp &H = hyp[nh];
                    H.skipped = skipped;
                    bool got = false;
                    int s0 = 0, s1 = 0;
                    if (n >= 2) {
                        for (unsigned it = 0; it < 1000u && !got; ++it) {
                            for (unsigned i = 0; i < 2u; ++i) {
                                const unsigned rnd = mt_next(mt) >> 1;
                                const unsigned j = i + rnd % unsigned(n - int(i));
                                const int t = sh[i]; sh[i] = sh[j]; sh[j] = t;
                            }
                            s0 = sh[0]; s1 = sh[1];
                            const float4 a = A[s0], b = A[s1];
                            got = (a.x != b.x) && (a.y != b.y) && (a.z != b.z);
                        }
                    }
                    if (!got) { H.state = 1; ++nh; break; }
                    const float4 a = A[s0], b = A[s1];
                    if (fabsf(a.x - b.x) <= FLT_EPSILON && fabsf(a.y - b.y) <= FLT_EPSILON && fabsf(a.z - b.z) <= FLT_EPSILON) {
                        ++skipped;
                        if (skipped >= max_skip) { H.state = 1; ++nh; break; }
                        continue;
                    }
                    H.c[0] = a.x; H.c[1] = a.y; H.c[2] = a.z;
                    float d0 = b.x - H.c[0], d1 = b.y - H.c[1], d2 = b.z - H.c[2];
                    const float z = dot3f(d0, d1, d2, d0, d1, d2);
                    if (z > 0.0f) { const float s = sqrtf(z); d0 /= s; d1 /= s; d2 /= s; }
                    H.c[3] = d0; H.c[4] = d1; H.c[5] = d2;
                    line_prep_dir(H.c, H.dir);
                    H.state = 0;
                    ++nh;
                }
                s_nhyp = nh;
            }
            __syncthreads();
            const int nh = s_nhyp;
            // score: warp `wid` takes hypotheses wid, wid+8, ...
            for (int hh = wid; hh < nh; hh += kLineThreads / 32) {
                if (hyp[hh].state != 0) continue;
                int cnt = 0;
                for (int i = lane; i < n; i += 32) {
                    const float4 p = A[i];
                    cnt += line_within(hyp[hh].c, hyp[hh].dir, p.x, p.y, p.z, sqr_thr) ? 1 : 0;
                }
                cnt = __reduce_add_sync(SPX_FULL, cnt);
                if (lane == 0) counts[hh] = cnt;
            }
            __syncthreads();
            if (tid == 0) {
                // replay `while (iterations_ < k && skipped_count < max_skip)` over the scored hypotheses
                int stop = 0;
                for (int hh = 0; hh < nh && !stop; ++hh) {
                    if (!(iterations < kk && hyp[hh].skipped < max_skip)) { stop = 1; break; }
                    if (hyp[hh].state != 0) { stop = 1; break; }
                    const int c = counts[hh];
                    if (c > n_best) {
                        n_best = c;
                        s_have = 1;
                        for (int q = 0; q < 6; ++q) s_best[q] = hyp[hh].c[q];
                        const double wv = double(n_best) * one_over_indices;
                        double p_no_outliers = 1.0 - wv * wv;
                        p_no_outliers = fmax(DBL_EPSILON, p_no_outliers);
                        p_no_outliers = fmin(1.0 - DBL_EPSILON, p_no_outliers);
                        kk = log_probability / log(p_no_outliers);
                    }
                    ++iterations;
                    if (iterations > P.ransac_max_iter) { stop = 1; break; }
                }
                if (!stop && !(iterations < kk && skipped < max_skip)) stop = 1;
                s_stop = stop; s_iter = iterations;
            }
            __syncthreads();
            if (s_stop) break;
        }
        const bool have = s_have != 0;
        int n_inl = 0;
        if (have) {
            // sac_->getInliers: selectWithinDistance(best)
            if (tid == 0) { for (int q = 0; q < 6; ++q) s_coef[q] = s_best[q]; line_prep_dir(s_best, s_dir); }
            __syncthreads();
            n_inl = block_select(n, inl, s_warp, [&](int i) { const float4 p = A[i]; return line_within(s_coef, s_dir, p.x, p.y, p.z, sqr_thr); });
            __syncthreads();
            // optimizeModelCoefficients: centroid + principal direction of the inliers (fp32, sequential sums)
            if (n_inl > 2) {
                if (tid < 3) {
                    float s = 0.0f;
                    for (int k = 0; k < n_inl; ++k) { const float4 p = A[inl[k]]; s += (tid == 0 ? p.x : (tid == 1 ? p.y : p.z)); }
                    s_cen[tid] = s / float(n_inl);
                }
                __syncthreads();
                if (tid < 6) {
                    const float c0 = s_cen[0], c1 = s_cen[1], c2 = s_cen[2];
                    float s = 0.0f;
                    for (int k = 0; k < n_inl; ++k) {
                        const float4 p = A[inl[k]];
                        const float ptx = p.x - c0, pty = p.y - c1, ptz = p.z - c2;
                        float t;
                        switch (tid) {
                            case 0: t = ptx * ptx; break;   // (0,0)
                            case 1: t = pty * ptx; break;   // (0,1)
                            case 2: t = ptz * ptx; break;   // (0,2)
                            case 3: t = pty * pty; break;   // (1,1)
                            case 4: t = pty * ptz; break;   // (1,2)
                            default: t = ptz * ptz; break;  // (2,2)
                        }
                        s += t;
                    }
                    s_cov[tid] = s;
                }
                __syncthreads();
                if (tid == 0) {
                    const float cov[9] = {s_cov[0], s_cov[1], s_cov[2], s_cov[1], s_cov[3], s_cov[4], s_cov[2], s_cov[4], s_cov[5]};
                    float vec[3];
                    eigen33_largest_vec(cov, vec);
                    s_coef[0] = s_cen[0]; s_coef[1] = s_cen[1]; s_coef[2] = s_cen[2];
                    s_coef[3] = vec[0]; s_coef[4] = vec[1]; s_coef[5] = vec[2];
                    line_prep_dir(s_coef, s_dir);
                }
                __syncthreads();
            }
            // refine inliers with the optimised coefficients
            n_inl = block_select(n, inl, s_warp, [&](int i) { const float4 p = A[i]; return line_within(s_coef, s_dir, p.x, p.y, p.z, sqr_thr); });
            __syncthreads();
        }
        // record
        const bool stop_round = double(n_inl) < P.line_ratio * double(boundSize);
        int in_range = 0, is_border = 0;
        if (!stop_round) {
            // LineInRange (src/Frame.cc:1058-1074) on the line point (= inlier centroid)
            {
                const float PcX = s_coef[0], PcY = s_coef[1], PcZ = s_coef[2];
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
            if (in_range) {
                // IsBorderLine (src/Frame.cc:1013-1024): fails as soon as more than s/4 points are not border points
                if (tid == 0) s_fail = 0;
                __syncthreads();
                int fails = 0;
                for (int k = tid; k < n_inl; k += kLineThreads) {
                    const float4 p = A[inl[k]];
                    if (!is_border_point(P, img, p.x, p.y, p.z)) ++fails;
                }
                if (fails) atomicAdd(&s_fail, fails);
                __syncthreads();
                is_border = (s_fail > n_inl / 4) ? 0 : 1;
            }
        }
        if (tid == 0) {
            L.model = m; L.round = round; L.n_points = n; L.iterations = s_iter; L.n_inliers = n_inl;
            L.in_range = in_range; L.is_border = is_border; L.emitted = 0; L.pts_off = M.contour_off + lp_off;
            for (int q = 0; q < 6; ++q) L.coef[q] = have ? s_coef[q] : 0.0f;
            M.n_rounds = round + 1;
        }
        if (stop_round) break;
        // ExtractIndices negative: the remaining cloud keeps its order
        {
            const int n_rem = block_select(n, inl, s_warp, [&](int i) { const float4 p = A[i]; return !line_within(s_coef, s_dir, p.x, p.y, p.z, sqr_thr); });
            __syncthreads();
            for (int k = tid; k < n_rem; k += kLineThreads) Bb[k] = A[inl[k]];
            __syncthreads();
            float4 *t = A; A = Bb; Bb = t;
            lp_off += n_inl;
            n = n_rem;
        }
        __syncthreads();
    }
}

// CaculatePlanes + the serial control flow of GeneratePlanesFromBoundries: one thread per frame
__global__ void __launch_bounds__(128) k_supposed(Params P, Buffers B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.n_frames) return;
    FrameCtl &ctl = B.ctl[f];
    int np = ctl.n_real, poff = ctl.pts_used, boff = ctl.bnd_used;
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
            if (np >= SPX_MAX_PLANES || poff + npts > P.pts_cap || boff + L.n_inliers > P.bnd_cap) { ctl.flags |= unsigned(SPX_FRAME_OVERFLOW); continue; }
            PlaneRec &R = ctl.planes[np];
            R.coef[0] = coef[0]; R.coef[1] = coef[1]; R.coef[2] = coef[2]; R.coef[3] = coef[3];
            R.n_points = npts; R.n_boundary = L.n_inliers; R.points_off = poff; R.boundary_off = boff;
            R.src = i; R.is_supposed = 1; R.line = m * SPX_MAX_LINES + j; R.pad = 0;
            L.emitted = 1;
            poff += npts; boff += L.n_inliers;
            ++np;
        }
    }
    ctl.n_planes = np; ctl.pts_used = poff; ctl.bnd_used = boff;
}

// clouds of the supposed planes (src/Frame.cc:1092-1110, 983-989) and the every-20th-inlier boundary fallback
// (src/Frame.cc:1001-1011); one CTA per (plane, frame)
__global__ void __launch_bounds__(128) k_pack_supposed(Params P, Buffers B) {
    const int f = blockIdx.y, k = blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    if (k >= ctl.n_planes) return;
    const PlaneRec &R = ctl.planes[k];
    spx_point *pts = B.out_pts + B.frame_offs[size_t(f) * 3 + 1] + R.points_off;
    spx_point *bnd = B.out_bnd + B.frame_offs[size_t(f) * 3 + 2] + R.boundary_off;
    if (!R.is_supposed) {
        if (ctl.models[R.src].n_contour == 0 && P.enable_supposed) {
            for (int j = threadIdx.x; j < R.n_boundary; j += blockDim.x) {
                spx_point q = pts[j * 20];
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

// exclusive scan of the per-frame totals (one CTA): where each frame's planes / points / boundary points start in the
// contiguous output buffers the pack kernels write to
__global__ void __launch_bounds__(1024) k_scan_frames(Params P, Buffers B) {
    __shared__ long long s_run[3];
    __shared__ long long s_w[3][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 3) s_run[tid] = 0;
    __syncthreads();
    for (int base = 0; base < P.n_frames; base += 1024) {
        const int f = base + tid;
        long long v[3] = {0, 0, 0};
        if (f < P.n_frames) { v[0] = B.ctl[f].n_planes; v[1] = B.ctl[f].pts_used; v[2] = B.ctl[f].bnd_used; }
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
        long long tot[3];
        for (int k = 0; k < 3; ++k) {
            const long long excl = s_run[k] + s_w[k][wid] + inc[k] - v[k];
            if (f < P.n_frames) B.frame_offs[size_t(f) * 3 + k] = excl;
            tot[k] = excl + v[k];
        }
        __syncthreads();
        if (tid == 1023) for (int k = 0; k < 3; ++k) s_run[k] = tot[k];
        __syncthreads();
    }
    if (tid < 3) B.out_totals[tid] = s_run[tid];
}

// frame headers and plane records with batch-global offsets; one CTA per frame
__global__ void __launch_bounds__(128) k_emit_records(Params P, Buffers B) {
    const int f = blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    const long long o_pl = B.frame_offs[size_t(f) * 3 + 0], o_pt = B.frame_offs[size_t(f) * 3 + 1], o_bd = B.frame_offs[size_t(f) * 3 + 2];
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
        o.points_off = o_pt + R.points_off; o.boundary_off = o_bd + R.boundary_off;
        o.src = R.src; o.is_supposed = R.is_supposed;
        B.out_planes[o_pl + k] = o;
    }
}

}  // namespace spx
