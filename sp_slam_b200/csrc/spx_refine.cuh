// spx_refine.cuh -- K6..K7 and the real-plane tail: refine() as two row-sequential sweeps with in-row bit-parallel
// chains, the serial Moore contour trace, SP-SLAM's sign fix + PlaneNotSeen de-duplication, and packing of the
// inlier clouds / contours.
//
// Reference: PCL 1.8.0 segmentation/impl/organized_multi_plane_segmentation.hpp (refine, segmentAndRefine),
// segmentation/plane_refinement_comparator.h (compare: |n.p + d| < 0.02f, not depth dependent),
// segmentation/impl/organized_connected_component_segmentation.hpp (findLabeledRegionBoundary);
// /root/reference/src/Frame.cc:912-934 (post filter), :1116-1144 (PlaneNotSeen), :1001-1011 (GenerateBoundaryPoints).
#pragma once
#include "spx_math.cuh"
#include "spx_types.cuh"

namespace spx {

constexpr int kRefWarps = 2;
constexpr int kMaxW = 512;           // widest organized cloud the row buffers hold (1280/3 = 427)
constexpr float kRefineThr = 0.02f;  // PlaneCoefficientComparator's default distance_threshold_

struct RefineSmem {
    float  coef[SPX_MAX_MODELS][4];
    int    n0[SPX_MAX_MODELS];
    int    cnt1[SPX_MAX_MODELS], cnt2[SPX_MAX_MODELS];
    int    last1[SPX_MAX_MODELS], last2[SPX_MAX_MODELS];
    int8_t rowA[kMaxW], rowB[kMaxW];   // plane ids of the previously finished row / the row being processed
    int8_t cmA[kMaxW];                 // model claimed sideways by claimer c  (pass 1: right, pass 2: left)
    int8_t cmB[kMaxW];                 // model claimed vertically by claimer c (pass 1: down, pass 2: up)
};

// point-to-plane test of PlaneRefinementComparator::compare (fp32 products and sums, no contraction)
__device__ __forceinline__ bool refine_dist_ok(const float *cf, float x, float y, float z) {
    const float v = cf[0] * x + cf[1] * y + cf[2] * z + cf[3];
    return fabsf(v) < kRefineThr;
}

// Ordered emission of the claims made by one row of claimers.  In visiting order claimer k issues its sideways claim
// (cmA) and then its vertical claim (cmB); `lane` enumerates claimers in visiting order inside a 32-wide segment.
// Every claimed pixel receives its position in inlier_indices[model] (= base[model] + running count).
template <bool kReverse>
__device__ __forceinline__ void refine_emit(RefineSmem &S, int w, int claimer_row, int *cnt, int *last, const int *pos_base,
                                            int *pos, int lane) {
    for (int base = 0; base < w; base += 32) {
        const int k = base + lane;
        const int c = kReverse ? (w - 1 - k) : k;
        const bool valid = k < w;
        const int mS = valid ? int(S.cmA[c]) : -1;
        const int mV = valid ? int(S.cmB[c]) : -1;
        unsigned todoS = __ballot_sync(SPX_FULL, mS >= 0), todoV = __ballot_sync(SPX_FULL, mV >= 0);
        while (todoS | todoV) {
            const int src = __ffs(todoS | todoV) - 1;
            const int mine = mS >= 0 ? mS : mV;
            const int mm = __shfl_sync(SPX_FULL, mine, src);
            const unsigned bS = __ballot_sync(SPX_FULL, mS == mm), bV = __ballot_sync(SPX_FULL, mV == mm);
            const unsigned lt = (1u << lane) - 1u;
            const int before = __popc(bS & lt) + __popc(bV & lt);
            const int b0 = cnt[mm];
            int qS, qV;
            if (kReverse) { qS = claimer_row * w + c - 1; qV = (claimer_row - 1) * w + c; }   // left (wraps at c == 0), up
            else          { qS = claimer_row * w + c + 1; qV = (claimer_row + 1) * w + c; }   // right, down
            if (mS == mm) pos[qS] = pos_base[mm] + b0 + before;
            if (mV == mm) pos[qV] = pos_base[mm] + b0 + before + (mS == mm ? 1 : 0);
            const int hl = 31 - __clz(bS | bV);
            __syncwarp();
            if (lane == hl) { last[mm] = (mV == mm) ? qV : qS; cnt[mm] = b0 + __popc(bS) + __popc(bV); }
            __syncwarp();
            todoS &= ~bS; todoV &= ~bV;
        }
    }
}

// In-row chain: every free pixel whose neighbour on the claimer side carries model m and that lies within 0.02 m of
// plane m is claimed, and then claims onward.  For one model this is a carry ripple: F = free & near-plane pixels,
// seeds = (sources shifted one step | carry-in) & F, claimed = F & ~(F + seeds).  Chains of different models never
// overlap (a pixel's fate is decided by its single claimer-side neighbour), so each is resolved independently.
// `lane` enumerates pixels in visiting order (pass 1: left to right; pass 2: right to left).
template <bool kReverse>
__device__ __forceinline__ void refine_chain(RefineSmem &S, int8_t *rowCur, int w, int r, const float *px, const float *py,
                                             const float *pz, int lane) {
    int carry = -1;
    for (int base = 0; base < w; base += 32) {
        const int k = base + lane;
        const int c = kReverse ? (w - 1 - k) : k;
        const bool valid = k < w;
        const int cur = valid ? int(rowCur[c]) : -2;
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid && cur == -1) { const int q = r * w + c; x = px[q]; y = py[q]; z = pz[q]; }
        int newcur = cur;
        unsigned todo = __ballot_sync(SPX_FULL, cur >= 0);
        bool carry_pending = carry >= 0;
        while (carry_pending || todo) {
            int mm;
            if (carry_pending) { mm = carry; carry_pending = false; }
            else { mm = __shfl_sync(SPX_FULL, cur, __ffs(todo) - 1); }
            const unsigned src = __ballot_sync(SPX_FULL, cur == mm);
            todo &= ~src;
            const unsigned F = __ballot_sync(SPX_FULL, cur == -1 && refine_dist_ok(S.coef[mm], x, y, z));
            const unsigned seeds = ((src << 1) | (carry == mm ? 1u : 0u)) & F;
            if (seeds == 0u) continue;
            const unsigned claimed = F & ~(F + seeds);
            if ((claimed >> lane) & 1u) {
                newcur = mm;
                S.cmA[kReverse ? c + 1 : c - 1] = int8_t(mm);   // the claimer is the previously visited pixel
            }
        }
        if (valid) rowCur[c] = int8_t(newcur);
        carry = __shfl_sync(SPX_FULL, newcur, 31);
        if (carry < 0) carry = -1;
    }
}

// K6: one warp per frame.
// Pass 1 (PCL: rows 0..h-2, cols 0..w-2, right neighbour then lower neighbour, labels read live):
//   target row r:  A) pixels claimed from above by the finished row r-1 (claimers c <= w-2), then
//                  B) the left-to-right chain inside row r (only rows <= h-2 have claimers).
// Pass 2 (PCL: rows h-1..1, cols w-1..0, left neighbour -- which at c == 0 is the last pixel of the row above --
//   then upper neighbour):
//   target row r:  A') claimed from below by the finished row r+1 (all columns), W) the wrap claim of (r+1, 0) on
//                  (r, w-1), then B') the right-to-left chain inside row r (rows >= 1).
// Claim order (= order of the appended inlier indices) is the visiting order of the CLAIMERS, so the claims of
// claimer row k are emitted once both its sideways and its vertical claims are known.
__global__ void __launch_bounds__(kRefWarps * 32) k_refine(Params P, Buffers B) {
    __shared__ RefineSmem smem[kRefWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kRefWarps + warp;
    if (f >= P.n_frames) return;
    RefineSmem &S = smem[warp];
    FrameCtl &ctl = B.ctl[f];
    const int nm = ctl.n_models;
    if (nm == 0) return;
    const int w = P.w, h = P.h;
    const size_t fo = size_t(f) * P.N;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    int8_t *pid = B.pid + fo;
    int *pos = B.pos + fo;

    for (int m = lane; m < nm; m += 32) {
        const Model &M = ctl.models[m];
        S.coef[m][0] = M.coef[0]; S.coef[m][1] = M.coef[1]; S.coef[m][2] = M.coef[2]; S.coef[m][3] = M.coef[3];
        S.n0[m] = M.n0; S.cnt1[m] = 0; S.cnt2[m] = 0; S.last1[m] = -1; S.last2[m] = -1;
    }
    for (int c = lane; c < w; c += 32) { S.cmA[c] = -1; S.cmB[c] = -1; }
    __syncwarp();

    int8_t *rowPrev = S.rowA, *rowCur = S.rowB;
    // ---------------- pass 1 ----------------
    for (int r = 0; r < h; ++r) {
        for (int c = lane; c < w; c += 32) { rowCur[c] = pid[r * w + c]; S.cmB[c] = -1; }
        __syncwarp();
        if (r >= 1) {
            for (int c = lane; c < w - 1; c += 32) {
                if (rowCur[c] < 0) {
                    const int m = rowPrev[c];
                    if (m >= 0) {
                        const int q = r * w + c;
                        if (refine_dist_ok(S.coef[m], px[q], py[q], pz[q])) { rowCur[c] = int8_t(m); S.cmB[c] = int8_t(m); }
                    }
                }
            }
            __syncwarp();
            refine_emit<false>(S, w, r - 1, S.cnt1, S.last1, S.n0, pos, lane);
        }
        for (int c = lane; c < w; c += 32) S.cmA[c] = -1;
        __syncwarp();
        if (r <= h - 2) refine_chain<false>(S, rowCur, w, r, px, py, pz, lane);
        __syncwarp();
        for (int c = lane; c < w; c += 32) pid[r * w + c] = rowCur[c];
        int8_t *t = rowPrev; rowPrev = rowCur; rowCur = t;
    }
    __threadfence_block();
    __syncwarp();
    // positions of pass-2 claims start after the originals and the pass-1 claims
    for (int m = lane; m < nm; m += 32) S.n0[m] += S.cnt1[m];
    for (int c = lane; c < w; c += 32) { S.cmA[c] = -1; S.cmB[c] = -1; }
    __syncwarp();
    // ---------------- pass 2 ----------------
    for (int r = h - 1; r >= 0; --r) {
        for (int c = lane; c < w; c += 32) { rowCur[c] = pid[r * w + c]; S.cmB[c] = -1; }
        __syncwarp();
        if (r <= h - 2) {
            for (int c = lane; c < w; c += 32) {
                if (rowCur[c] < 0) {
                    const int m = rowPrev[c];
                    if (m >= 0) {
                        const int q = r * w + c;
                        if (refine_dist_ok(S.coef[m], px[q], py[q], pz[q])) { rowCur[c] = int8_t(m); S.cmB[c] = int8_t(m); }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {   // wrap claim of (r+1, 0) on (r, w-1)
                const int m = rowPrev[0];
                if (m >= 0 && rowCur[w - 1] < 0) {
                    const int q = r * w + w - 1;
                    if (refine_dist_ok(S.coef[m], px[q], py[q], pz[q])) { rowCur[w - 1] = int8_t(m); S.cmA[0] = int8_t(m); }
                }
            }
            __syncwarp();
            refine_emit<true>(S, w, r + 1, S.cnt2, S.last2, S.n0, pos, lane);
        }
        for (int c = lane; c < w; c += 32) S.cmA[c] = -1;
        __syncwarp();
        if (r >= 1) refine_chain<true>(S, rowCur, w, r, px, py, pz, lane);
        __syncwarp();
        for (int c = lane; c < w; c += 32) pid[r * w + c] = rowCur[c];
        int8_t *t = rowPrev; rowPrev = rowCur; rowCur = t;
    }
    __syncwarp();
    for (int m = lane; m < nm; m += 32) {
        Model &M = ctl.models[m];
        M.n1 = S.cnt1[m]; M.n2 = S.cnt2[m];
        if (S.cnt2[m] > 0) M.last_inlier = S.last2[m];
        else if (S.cnt1[m] > 0) M.last_inlier = S.last1[m];
    }
}

// K7: findLabeledRegionBoundary from inlier_indices[i].back(); one thread per model (the walk is serial).
__constant__ int c_ddx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
__constant__ int c_ddy[8] = {0, -1, -1, -1, 0, 1, 1, 1};

__global__ void __launch_bounds__(32) k_contour(Params P, Buffers B) {
    const int f = blockIdx.x, lane = threadIdx.x;
    FrameCtl &ctl = B.ctl[f];
    const int nm = ctl.n_models;
    const int w = P.w, h = P.h;
    const size_t fo = size_t(f) * P.N;
    const int8_t *pid = B.pid + fo;
    int *arena = B.contour_idx + size_t(f) * P.contour_cap;
    for (int m = lane; m < nm; m += 32) {
        Model &M = ctl.models[m];
        int off = 0;
        for (int k = 0; k < m; ++k) { const Model &K = ctl.models[k]; off += 2 * (K.n0 + K.n1 + K.n2) + 16; }
        const int cap = 2 * (M.n0 + M.n1 + M.n2) + 16;
        M.contour_off = off;
        int *out = arena + off;
        const int start = M.last_inlier;
        int cx = start % w, cy = start / w, cidx = start;
        int direction = -1;
        for (int d = 0; d < 8; ++d) {
            const int x = cx + c_ddx[d], y = cy + c_ddy[d];
            if (x >= 0 && x < w && y >= 0 && y < h && pid[y * w + x] != m) { direction = d; break; }
        }
        int n = 0;
        if (direction != -1) {
            out[n++] = start;
            const long long guard = 8ll * P.N + 8;
            long long steps = 0;
            bool overflow = false;
            do {
                int nIdx = direction;
                for (int d = 1; d <= 8; ++d) {
                    nIdx = (direction + d) & 7;
                    const int x = cx + c_ddx[nIdx], y = cy + c_ddy[nIdx];
                    if (x >= 0 && x < w && y >= 0 && y < h && pid[y * w + x] == m) break;
                }
                direction = (nIdx + 4) & 7;
                cx += c_ddx[nIdx]; cy += c_ddy[nIdx];
                cidx = cy * w + cx;
                if (n < cap) out[n++] = cidx; else overflow = true;
                if (++steps > guard || cx < 0 || cx >= w || cy < 0 || cy >= h) { overflow = true; break; }
            } while (cidx != start);
            if (overflow) atomicOr(&ctl.flags, unsigned(SPX_FRAME_OVERFLOW));
        }
        M.n_contour = n;
    }
}

// SP-SLAM's PlaneNotSeen (src/Frame.cc:1116-1144) against the planes kept so far
__device__ __forceinline__ bool plane_not_seen(const FrameCtl &ctl, int n_planes, const float coef[4]) {
    for (int j = 0; j < n_planes; ++j) {
        const float *pM = ctl.planes[j].coef;
        const float d = pM[3] - coef[3];
        const float angle = pM[0] * coef[0] + pM[1] * coef[1] + pM[2] * coef[2];
        if (double(d) > 0.2 || double(d) < -0.2) continue;
        if (double(angle) < 0.9397 && double(angle) > -0.9397) continue;
        return false;
    }
    return true;
}

// real-plane tail of ComputePlanesFromOrganizedPointCloud (src/Frame.cc:912-934): one thread per frame
__global__ void __launch_bounds__(128) k_postfilter(Params P, Buffers B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.n_frames) return;
    FrameCtl &ctl = B.ctl[f];
    int np = 0, poff = 0, boff = 0;
    for (int i = 0; i < ctl.n_models; ++i) {
        Model &M = ctl.models[i];
        float coef[4] = {M.coef[0], M.coef[1], M.coef[2], M.coef[3]};
        if (coef[3] < 0) { coef[0] = -coef[0]; coef[1] = -coef[1]; coef[2] = -coef[2]; coef[3] = -coef[3]; }
        M.plane = -1;
        if (!plane_not_seen(ctl, np, coef)) continue;
        const int npts = M.n0 + M.n1 + M.n2;
        // an empty contour is replaced by every 20th inlier inside GeneratePlanesFromBoundries (src/Frame.cc:958-959)
        const int nb = (M.n_contour == 0 && P.enable_supposed) ? (npts + 19) / 20 : M.n_contour;
        if (np >= SPX_MAX_PLANES || poff + npts > P.pts_cap || boff + nb > P.bnd_cap) { ctl.flags |= unsigned(SPX_FRAME_OVERFLOW); continue; }
        PlaneRec &R = ctl.planes[np];
        R.coef[0] = coef[0]; R.coef[1] = coef[1]; R.coef[2] = coef[2]; R.coef[3] = coef[3];
        R.n_points = npts; R.n_boundary = nb; R.points_off = poff; R.boundary_off = boff;
        R.src = i; R.is_supposed = 0; R.line = -1; R.pad = 0;
        M.plane = np;
        poff += npts; boff += nb;
        ++np;
    }
    ctl.n_real = np; ctl.n_planes = np; ctl.pts_used = poff; ctl.bnd_used = boff; ctl.n_lines = 0;
}

// ExtractIndices(negative = false) of the kept planes: inlier points in inlier_indices order (src/Frame.cc:925-928)
__global__ void __launch_bounds__(256) k_pack_points(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    const int m = B.pid[fo + q];
    if (m < 0) return;
    const FrameCtl &ctl = B.ctl[f];
    const int k = ctl.models[m].plane;
    if (k < 0) return;
    spx_point pt;
    pt.x = B.px[fo + q]; pt.y = B.py[fo + q]; pt.z = B.pz[fo + q]; pt.rgba = pack_rgba(0, 0, 250);
    B.out_pts[B.frame_offs[size_t(f) * 3 + 1] + ctl.planes[k].points_off + B.pos[fo + q]] = pt;
}

// regions[i].getContour() of the kept planes (src/Frame.cc:930-932); one CTA per (model, frame)
__global__ void __launch_bounds__(128) k_pack_contours(Params P, Buffers B) {
    const int f = blockIdx.y, m = blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    if (m >= ctl.n_models) return;
    const Model &M = ctl.models[m];
    if (M.plane < 0 || M.n_contour == 0) return;
    const size_t fo = size_t(f) * P.N;
    const int *src = B.contour_idx + size_t(f) * P.contour_cap + M.contour_off;
    spx_point *dst = B.out_bnd + B.frame_offs[size_t(f) * 3 + 2] + ctl.planes[M.plane].boundary_off;
    for (int j = threadIdx.x; j < M.n_contour; j += blockDim.x) {
        const int q = src[j];
        spx_point pt;
        pt.x = B.px[fo + q]; pt.y = B.py[fo + q]; pt.z = B.pz[fo + q]; pt.rgba = pack_rgba(0, 0, 250);
        dst[j] = pt;
    }
}

}  // namespace spx
