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
// frames handled by the multi-warp kernel k_refine2 (its per-(row, model) count table must fit); the rest: k_refine
constexpr int kRefMaxH = 512;
constexpr int kRefTableCap = 4096;   // entries of the per-(row, model) count table

__host__ __device__ __forceinline__ bool refine_fast_ok(int h, int nm) { return h <= kRefMaxH && h * nm <= kRefTableCap; }


struct RefineSmem {
    float  coef[SPX_MAX_MODELS][4];
    int    n0[SPX_MAX_MODELS];
    int    cnt1[SPX_MAX_MODELS], cnt2[SPX_MAX_MODELS];
    int    last1[SPX_MAX_MODELS], last2[SPX_MAX_MODELS];
};

__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

// point-to-plane test of PlaneRefinementComparator::compare (fp32 products and sums, no contraction)
__device__ __forceinline__ bool refine_dist_ok(const float *cf, float x, float y, float z) {
    const float v = cf[0] * x + cf[1] * y + cf[2] * z + cf[3];
    return fabsf(v) < kRefineThr;
}

// One raster pass of refine() over a frame, one warp.  kReverse = false: PCL's first pass (rows 0..h-2, cols 0..w-2,
// right neighbour then lower neighbour, labels read live); kReverse = true: the second pass (rows h-1..1, cols
// w-1..0, left neighbour -- which at c == 0 is the last pixel of the row above -- then upper neighbour).
//
// Rows are visited in pass order; lane l of chunk ch owns visiting index k = 32 ch + l (column c = k, or w-1-k in the
// reverse pass).  For target row r:
//   A) pixels claimed vertically by the finished previous row (the claimer sits in the same column),
//   W) reverse pass only: the wrap claim of (r+1, 0) on (r, w-1),
//   E) ordered emission of the claims made by the previous row's claimers: in visiting order claimer k issues its
//      sideways claim and then its vertical claim; every claimed pixel gets its position in inlier_indices[model],
//   B) the chain inside row r: a free pixel is claimed by its already-final neighbour on the claimer side.  With
//      src = nearest labelled pixel on that side (or the carry from the previous chunk), a free pixel is claimed iff
//      every pixel between src and itself is free and it and they lie within 0.02 m of src's plane (a point PCL left
//      unlabelled stops the chain) -- ballots and bit masks only.
// Row r+1's plane ids and the xyz of its free pixels are prefetched into registers while row r is processed (plane ids
// two rows ahead), so no global-load latency sits on the row-to-row dependency chain.
// All of it is steered by warp-uniform lane masks per chunk (plane pixels / free pixels of the previous, this and the next
// row; the claims of the previous row's claimers): where a mask is zero the chunk costs a scalar test.
template <int NCH, bool kReverse>
__device__ __forceinline__ void refine_pass(RefineSmem &S, const Params &P, const float *__restrict__ px, const float *__restrict__ py,
                                            const float *__restrict__ pz, int8_t *pid, int *pos, int *cnt, int *last,
                                            const int *pos_base, int lane, bool has_invalid) {
    const int w = P.w, h = P.h;
    constexpr int kD = kReverse ? -32 : 32;      // column step from chunk to chunk
    const int rstep = kReverse ? -w : w;         // flat-index step from visited row to visited row
    const int c0 = kReverse ? (w - 1 - lane) : lane;   // this lane's column in chunk 0
    const unsigned lt = (1u << lane) - 1u;
    const int ch_last = (w - 1) >> 5, lane_last = (w - 1) & 31;   // where the last visited pixel of a row lives
    int a[NCH], prev[NCH], pn[NCH], pn2[NCH];
    float x[NCH], y[NCH], z[NCH], xn[NCH], yn[NCH], zn[NCH];
    // warp-uniform lane masks per chunk.  Everything below is steered by them: a chunk of a row in which no free pixel can meet a
    // plane costs a few scalar tests, not a round of votes.
    unsigned lab[NCH], fre[NCH];      // row being processed: pixels of a plane / free pixels (labelled by PCL, part of no plane)
    unsigned labn[NCH], fren[NCH];    // the next row
    unsigned labp[NCH];               // previous row, final
    unsigned smaskp[NCH];             // claimers of the previous row that claimed sideways
    unsigned vmask[NCH];              // pixels of this row claimed vertically
    // running flat indices / pointers: row being processed, one row ahead (xyz), two rows ahead (plane ids)
    int q0 = (kReverse ? (h - 1) * w : 0) + c0;  // (r, c0)
    const float *px1 = px + q0, *py1 = py + q0, *pz1 = pz + q0;   // advanced to row t + 1 before use
    const int8_t *pid2 = pid + q0 + rstep;                        // advanced to row t + 2 before use
    // prologue: plane ids of rows t = 0 and 1, xyz of the free pixels of row 0
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const bool valid = ch * 32 + lane < w;
        pn[ch] = valid ? int(pid[q0 + ch * kD]) : -2;
        pn2[ch] = (valid && h > 1) ? int(pid2[ch * kD]) : -2;
        prev[ch] = -2;
        labp[ch] = 0u; smaskp[ch] = 0u; vmask[ch] = 0u;
        xn[ch] = yn[ch] = zn[ch] = 0.f;
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        labn[ch] = __ballot_sync(SPX_FULL, pn[ch] >= 0);
        fren[ch] = __ballot_sync(SPX_FULL, pn[ch] == -1);
        if (pn[ch] == -1) { xn[ch] = px1[ch * kD]; yn[ch] = py1[ch * kD]; zn[ch] = pz1[ch * kD]; }
    }
    int wrapm = -1;         // (warp uniform) reverse pass: final plane of the last visited pixel of the previous row

    for (int t = 0; t < h; ++t, q0 += rstep) {
        // rotate the prefetch registers
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            a[ch] = pn[ch]; x[ch] = xn[ch]; y[ch] = yn[ch]; z[ch] = zn[ch]; pn[ch] = pn2[ch];
            lab[ch] = labn[ch]; fre[ch] = fren[ch];
        }
        px1 += rstep; py1 += rstep; pz1 += rstep; pid2 += rstep;
        // the next row: its masks, and the xyz of those of its free pixels that a plane can reach.  A plane grows by one row per
        // row step and along a row only in visiting direction, so a chunk of the next row can be tested only if a plane pixel
        // exists in the same or an earlier-visited chunk of the previous, this or the next row (reverse pass: the wrap claim
        // also leads from the last visited chunk to the first).
        if (t + 1 < h) {
            unsigned M = 0u, Fn = 0u;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                labn[ch] = __ballot_sync(SPX_FULL, pn[ch] >= 0);
                fren[ch] = __ballot_sync(SPX_FULL, pn[ch] == -1);
                if (labp[ch] | lab[ch] | labn[ch]) M |= 1u << ch;
                if (fren[ch]) Fn |= 1u << ch;
            }
            unsigned reach = M;
            reach |= reach << 1; reach |= reach << 2; reach |= reach << 4; reach |= reach << 8;
            if (kReverse && M) reach = ~0u;   // (a chain that reaches the end of a row wraps into the first chunk of the next)
            const unsigned need = Fn & reach;
            if (need) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
                    if ((need >> ch) & 1u) {
                        if (pn[ch] == -1) { xn[ch] = px1[ch * kD]; yn[ch] = py1[ch * kD]; zn[ch] = pz1[ch * kD]; }
                    }
            }
        }
        if (t + 2 < h) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) pn2[ch] = (ch * 32 + lane < w) ? int(pid2[ch * kD]) : -2;
        }
        unsigned changed = 0u;   // chunks whose labels changed in this row
        if (t >= 1) {
            // A) vertical claims by the previous row: a free pixel under (above) a plane pixel
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                unsigned cm = fre[ch] & labp[ch];
                if (!kReverse && ch == ch_last) cm &= ~(1u << lane_last);   // PCL's first pass visits columns 0 .. w-2
                vmask[ch] = 0u;
                if (cm == 0u) continue;
                const int m = prev[ch];
                bool cand = (cm >> lane) & 1u;
                // PCL: `if (current_label < 0 || right_label < 0) continue;` also drops the claimer's vertical claim when
                // its sideways neighbour (flat index +-1) is an unlabelled point; only frames with non-finite depth have those
                if (has_invalid && cand && pid[q0 - rstep + ch * kD + (kReverse ? -1 : 1)] == -2) cand = false;
                const bool take = cand && refine_dist_ok(S.coef[cand ? m : 0], x[ch], y[ch], z[ch]);
                if (take) a[ch] = m;
                const unsigned vm = __ballot_sync(SPX_FULL, take);
                vmask[ch] = vm;
                if (vm) { lab[ch] |= vm; fre[ch] &= ~vm; changed |= 1u << ch; }
            }
            // W) wrap claim of (r+1, 0) on (r, w-1): visiting index w-1 of the previous row claims visiting index 0
            if (kReverse && wrapm >= 0 && (fre[0] & 1u)) {
                const bool take = lane == 0 && refine_dist_ok(S.coef[wrapm], x[0], y[0], z[0]);
                if (take) a[0] = wrapm;
                if (__any_sync(SPX_FULL, take)) {
                    lab[0] |= 1u; fre[0] &= ~1u; changed |= 1u;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) if (ch == ch_last) smaskp[ch] |= 1u << lane_last;
                }
            }
            // E) emission for the claimers of the previous row: in visiting order claimer k issues its sideways claim and then
            // its vertical claim; every claimed pixel gets its position in inlier_indices[model]
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                unsigned todoS = smaskp[ch], todoV = vmask[ch];
                if ((todoS | todoV) == 0u) continue;
                const int mS = ((todoS >> lane) & 1u) ? prev[ch] : -1;      // the claimer's own plane
                const int mv = ((todoV >> lane) & 1u) ? a[ch] : -1;
                const int qV = q0 + ch * kD;                                  // (r, c): below / above the claimer
                const int qS = qV - rstep + (kReverse ? -1 : 1);              // the claimer's sideways neighbour (wraps at c == 0)
                while (todoS | todoV) {
                    const int src = __ffs(todoS | todoV) - 1;
                    const int mine = mS >= 0 ? mS : mv;
                    const int mm = __shfl_sync(SPX_FULL, mine, src);
                    const unsigned bS = __ballot_sync(SPX_FULL, mS == mm), bV = __ballot_sync(SPX_FULL, mv == mm);
                    const int before = __popc(bS & lt) + __popc(bV & lt);
                    const int b0 = cnt[mm];
                    if (mS == mm) pos[qS] = pos_base[mm] + b0 + before;
                    if (mv == mm) pos[qV] = pos_base[mm] + b0 + before + (mS == mm ? 1 : 0);
                    const int hl = 31 - __clz(bS | bV);
                    __syncwarp();
                    if (lane == hl) { last[mm] = (mv == mm) ? qV : qS; cnt[mm] = b0 + __popc(bS) + __popc(bV); }
                    __syncwarp();
                    todoS &= ~bS; todoV &= ~bV;
                }
            }
        }
        // B) the chain inside row r (claimers: rows <= h-2 in the forward pass, rows >= 1 in the reverse pass): a free pixel is
        // claimed by its already-final neighbour on the claimer side.  Only chunks in which a free pixel follows a plane pixel
        // (or the carry) can claim anything.
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) smaskp[ch] = 0u;
        if (t <= h - 2) {
            int carry = -1;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                unsigned clm = 0u;
                if (fre[ch] & ((lab[ch] << 1) | (carry >= 0 ? 1u : 0u))) {
                    const int cur = a[ch];
                    const unsigned below = lab[ch] & lt;
                    const int src_lane = below ? (31 - __clz(below)) : -1;
                    const int m_lane = __shfl_sync(SPX_FULL, cur, src_lane < 0 ? 0 : src_lane);
                    const int m_src = src_lane < 0 ? carry : m_lane;
                    const bool ok = cur == -1 && m_src >= 0 && refine_dist_ok(S.coef[m_src < 0 ? 0 : m_src], x[ch], y[ch], z[ch]);
                    // what stops a chain: a free pixel that fails the test, and a point PCL left unlabelled (non-finite; lanes
                    // beyond the row count as such)
                    const unsigned blocked = ~(lab[ch] | __ballot_sync(SPX_FULL, ok));
                    // the pixels visited between src and this one: bits src_lane+1 .. lane-1
                    const unsigned le_src = src_lane < 0 ? 0u : ((2u << src_lane) - 1u);
                    const bool claimed = ok && (blocked & lt & ~le_src) == 0u;
                    clm = __ballot_sync(SPX_FULL, claimed);
                    if (claimed) a[ch] = m_src;
                    if (clm) {
                        lab[ch] |= clm; changed |= 1u << ch;
                        // the claimer is the previously visited pixel
                        smaskp[ch] |= clm >> 1;
                        if (ch > 0 && (clm & 1u)) smaskp[ch - 1] |= 1u << 31;
                    }
                }
                carry = -1;
                if (ch + 1 < NCH && (fre[ch + 1 < NCH ? ch + 1 : ch] & 1u) && ((lab[ch] >> 31) & 1u)) carry = __shfl_sync(SPX_FULL, a[ch], 31);
            }
        }
        // final labels of row r
        if (kReverse) {
            int lastv = a[NCH - 1];
            if (ch_last != NCH - 1) {
#pragma unroll
                for (int ch = 0; ch < NCH - 1; ++ch) if (ch == ch_last) lastv = a[ch];
            }
            wrapm = __shfl_sync(SPX_FULL, lastv, lane_last);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if ((changed >> ch) & 1u) { if (ch * 32 + lane < w) pid[q0 + ch * kD] = int8_t(a[ch]); }
            prev[ch] = a[ch];
            labp[ch] = lab[ch];
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(kRefWarps * 32) k_refine(Params P, Buffers B) {
    __shared__ RefineSmem smem[kRefWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fl = blockIdx.x * kRefWarps + warp;
    if (fl >= P.n_frames) return;
    const int f = P.frame0 + fl;
    RefineSmem &S = smem[warp];
    FrameCtl &ctl = B.ctl[f];
    const int nm = ctl.n_models;
    if (nm == 0 || (P.refine_fast && refine_fast_ok(P.h, nm))) return;   // k_refine2 takes the frames whose count table fits
    const size_t fo = size_t(f) * P.N;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    int8_t *pid = B.pid + fo;
    int *pos = B.pos + fo;

    for (int m = lane; m < nm; m += 32) {
        const Model &M = ctl.models[m];
        S.coef[m][0] = M.coef[0]; S.coef[m][1] = M.coef[1]; S.coef[m][2] = M.coef[2]; S.coef[m][3] = M.coef[3];
        S.n0[m] = M.n0; S.cnt1[m] = 0; S.cnt2[m] = 0; S.last1[m] = -1; S.last2[m] = -1;
    }
    __syncwarp();
    const bool has_invalid = (ctl.flags & unsigned(SPX_FRAME_NONFINITE)) != 0u;
    refine_pass<NCH, false>(S, P, px, py, pz, pid, pos, S.cnt1, S.last1, S.n0, lane, has_invalid);
    __threadfence_block();
    __syncwarp();
    // positions of pass-2 claims start after the originals and the pass-1 claims
    for (int m = lane; m < nm; m += 32) S.n0[m] += S.cnt1[m];
    __syncwarp();
    refine_pass<NCH, true>(S, P, px, py, pz, pid, pos, S.cnt2, S.last2, S.n0, lane, has_invalid);
    __syncwarp();
    for (int m = lane; m < nm; m += 32) {
        Model &M = ctl.models[m];
        M.n1 = S.cnt1[m]; M.n2 = S.cnt2[m];
        if (S.cnt2[m] > 0) M.last_inlier = S.last2[m];
        else if (S.cnt1[m] > 0) M.last_inlier = S.last1[m];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K6 (fast path): the same two raster passes with the parallelism INSIDE a frame.  One CTA per frame, one warp per
// 32-column chunk.  A pass is split in two:
//  (1) label propagation on a skewed (systolic) schedule: warp k handles chunk k of the row at step T(row) + k, so
//      the vertical claim (same chunk, previous row) and the chain carry (previous chunk, same row) it depends on
//      were produced one step earlier.  Which pixels were claimed sideways / vertically is kept as bit rows.
//      In the reverse pass the "left neighbour" of (r, 0) is the last pixel of the row above, i.e. the chain runs on
//      through the flattened image: a row whose first visited pixel could take such a wrap claim (it is free and
//      within 0.02 m of some model plane) starts only after the row below is complete (T advances by NW for it).
//  (2) emission, parallel over rows: per (claimer row, model) claim counts, an exclusive prefix over the rows, then
//      every claimed pixel gets its position in inlier_indices[model] -- claimers in visiting order, the sideways
//      claim of a claimer before its vertical claim, exactly the order in which the reference appends them.
// Frames whose (rows x models) table does not fit are left to the one-warp kernel k_refine.  The multi-warp kernel
// executes more instructions per frame (every warp pays the per-step overhead), so it is the choice for small launches
// (latency: the tracking loop), while large batches keep one warp per frame (throughput) -- Params::refine_fast.
// ---------------------------------------------------------------------------------------------------------------
struct Refine2Smem {
    float    coef[SPX_MAX_MODELS][4];
    int      n0[SPX_MAX_MODELS];
    int      cnt[2][SPX_MAX_MODELS];
    unsigned long long lastkey[2][SPX_MAX_MODELS];
    uint8_t  wrapflag[kRefMaxH];      // rows that wait for the row below (wrap claim possible)
    uint8_t  wrapcand[kRefMaxH];      // rows whose first visited pixel is free and close to some model plane
};
// hand-over between the warps of a frame (dynamic shared memory, behind the count table): carryq[i * NW + k] = label the chain
// carries out of chunk k of row i (pass order), rowlastq[i] = final label of the last visited pixel of row i; kRefPending
// until the owning warp has posted it
constexpr int kRefPending = -128;
__host__ __device__ __forceinline__ size_t refine2_smem_bytes(int h, int nw) {
    return (size_t(3) * h * nw + kRefTableCap) * sizeof(unsigned) + ((size_t(h) * (nw + 1) + 15) & ~size_t(15));   // (3 h nw >= 2 h nw + h: rowmask)
}
__device__ __forceinline__ int refine2_wait(const volatile signed char *q) {
    int v = *q;
    while (v == kRefPending) { __nanosleep(20); v = *q; }
    return v;
}

template <int NW, bool kReverse>
__device__ __forceinline__ void refine2_propagate(Refine2Smem &S, unsigned *Hb, unsigned *Vb, unsigned *rowmask, volatile signed char *carryq,
                                                  volatile signed char *rowlastq, const Params &P, const float *__restrict__ px,
                                                  const float *__restrict__ py, const float *__restrict__ pz, int8_t *pid, bool has_invalid) {
    const int w = P.w, h = P.h;
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = k * 32 + lane;                  // visiting index inside a row
    const bool valid = v < w;
    const int c = kReverse ? (w - 1 - v) : v;
    auto row_of = [&](int i) { return kReverse ? (h - 1 - i) : i; };
    const int k_last = (w - 1) >> 5, lane_last = (w - 1) & 31;   // where the last visited pixel of a row lives
    // statically rotated prefetch registers: plane ids 8 rows ahead, xyz of the free pixels 4 rows ahead -- a load has
    // at least four rows to land and no register holding a pending load is ever moved
    int pr[8];
    float xr[4], yr[4], zr[4];
#pragma unroll
    for (int u = 0; u < 8; ++u) pr[u] = (valid && u < h) ? int(pid[row_of(u) * w + c]) : -2;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        xr[u] = yr[u] = zr[u] = 0.f;
        if (u < h && pr[u] == -1) { const int q = row_of(u) * w + c; xr[u] = px[q]; yr[u] = py[q]; zr[u] = pz[q]; }
    }
    int prev = -2;
    // Data flow instead of a lock-step schedule: warp k works down the rows of its chunk on its own and waits only where it
    // needs something another warp produces -- the label the chain carries out of chunk k-1 of the same row (and, in the
    // reverse pass, the wrap claim of the row below into the first chunk of a waiting row).  A chunk without a free pixel has
    // nothing to claim and nothing to wait for: it posts the label of its last pixel and moves on.
    for (int ib = 0; ib < h; ib += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = ib + u;
            if (i >= h) break;
            const int r = row_of(i);
            int a = pr[u];
            const bool chain_on = kReverse ? (r >= 1) : (r <= h - 2);
            unsigned hb = 0u, vb = 0u;
            if (__any_sync(SPX_FULL, valid && a == -1)) {
                const float x = xr[u & 3], y = yr[u & 3], z = zr[u & 3];
                // A) vertical claim by the previous row (same column)
                bool cV = false;
                {
                    bool cand = i >= 1 && a == -1 && prev >= 0 && (kReverse || c <= w - 2);
                    if (has_invalid && cand) {   // (see k_refine: an unlabelled sideways neighbour of the claimer cancels its vertical claim)
                        const int cr = kReverse ? r + 1 : r - 1;
                        if (pid[cr * w + c + (kReverse ? -1 : 1)] == -2) cand = false;
                    }
                    if (__any_sync(SPX_FULL, cand)) {
                        if (cand && refine_dist_ok(S.coef[prev], x, y, z)) { a = prev; cV = true; }
                    }
                }
                // B) chain along the row: claimers are rows <= h-2 (forward) / rows >= 1 (reverse); in the reverse pass the
                // carry into the first chunk is the wrap claim of the row below
                bool cH = false;
                const unsigned freem = __ballot_sync(SPX_FULL, valid && a == -1);
                if (freem != 0u && (chain_on || k == 0)) {
                    int carry = -1;
                    if (k > 0) { if (chain_on) carry = refine2_wait(carryq + i * NW + k - 1); }
                    else if (kReverse && i >= 1 && S.wrapflag[r]) carry = refine2_wait(rowlastq + i - 1);
                    const unsigned labelled = __ballot_sync(SPX_FULL, a >= 0);
                    if (chain_on ? (labelled != 0u || carry >= 0) : (k == 0 && carry >= 0)) {
                        const unsigned below = chain_on ? (labelled & ((1u << lane) - 1u)) : 0u;
                        const int src_lane = below ? (31 - __clz(below)) : -1;
                        const int m_lane = __shfl_sync(SPX_FULL, a, src_lane < 0 ? 0 : src_lane);
                        const int m_src = src_lane < 0 ? carry : m_lane;
                        const bool isfree = valid && a == -1;
                        const bool ok = isfree && m_src >= 0 && refine_dist_ok(S.coef[m_src], x, y, z);
                        const unsigned blocked = __ballot_sync(SPX_FULL, !valid || a == -2 || (a == -1 && !ok));   // (-2: a point PCL left unlabelled stops a chain)
                        const unsigned le_src = src_lane < 0 ? 0u : ((2u << src_lane) - 1u);
                        const unsigned mask = ((1u << lane) - 1u) & ~le_src;
                        bool claimed = ok && (blocked & mask) == 0u;
                        if (!chain_on) claimed = claimed && lane == 0;   // row 0 of the reverse pass: only the wrap claim itself
                        if (claimed) { a = m_src; cH = true; }
                    }
                }
                hb = __ballot_sync(SPX_FULL, cH); vb = __ballot_sync(SPX_FULL, cV);
                if (valid && (cH || cV)) pid[r * w + c] = int8_t(a);
            }
            const int last_lbl = __shfl_sync(SPX_FULL, a, (k == k_last) ? lane_last : 31);
            if (lane == 0) {
                carryq[i * NW + k] = (signed char)(chain_on ? (last_lbl < 0 ? -1 : last_lbl) : -1);
                if (k == k_last) rowlastq[i] = (signed char)(last_lbl < 0 ? -1 : last_lbl);
                Hb[i * NW + k] = hb; Vb[i * NW + k] = vb;
                if (hb | vb) atomicOr(&rowmask[i], 1u << k);      // chunks of the row that hold claimed pixels (what the emission visits)
            }
            prev = a;
            // refill the slots just consumed: plane ids of row i+8, xyz of row i+4 (its plane ids arrived long ago)
            pr[u] = (valid && i + 8 < h) ? int(pid[row_of(i + 8) * w + c]) : -2;
            if (i + 4 < h && pr[(u + 4) & 7] == -1) { const int q = row_of(i + 4) * w + c; xr[u & 3] = px[q]; yr[u & 3] = py[q]; zr[u & 3] = pz[q]; }
        }
    }
}

// emission of one pass; kCount: accumulate the per-(row, model) counts, else assign positions from the prefixed table
template <int NW, bool kReverse, bool kCount>
__device__ __forceinline__ void refine2_emit(Refine2Smem &S, const unsigned *Hb, const unsigned *Vb, const unsigned *rowmask, unsigned *table, int nm, const Params &P,
                                             const int8_t *pid, int *pos, const int *pos_base, int pass) {
    const int w = P.w, h = P.h;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_last = (w - 1) >> 5, lane_last = (w - 1) & 31;
    for (int i = wid; i < h; i += NW) {           // claimer row i (pass order)
        const int r = kReverse ? (h - 1 - i) : i;
        // chunks that can hold a claimer of row i with a claim: a sideways target in chunk k or k + 1 of this row (the wrap target
        // is bit 0 of the next row), a vertical target in chunk k of the next row
        unsigned cm = rowmask[i];
        cm |= cm >> 1;
        if (i + 1 < h) { const unsigned m1 = rowmask[i + 1]; cm |= m1; if (kReverse && (m1 & 1u)) cm |= 1u << k_last; }
        while (cm) {
            const int k = __ffs(cm) - 1;
            cm &= cm - 1u;
            // sideways targets of the claimers of chunk k: visiting index v + 1 (the wrap target is bit 0 of the next row)
            unsigned sm = Hb[i * NW + k] >> 1;
            if (k + 1 < NW) sm |= Hb[i * NW + k + 1] << 31;
            if (kReverse && k == k_last && i + 1 < h) sm |= (Hb[(i + 1) * NW] & 1u) << lane_last;
            const unsigned vm = (i + 1 < h) ? Vb[(i + 1) * NW + k] : 0u;
            if ((sm | vm) == 0u) continue;        // warp uniform
            const int v = k * 32 + lane;
            const int c = kReverse ? (w - 1 - v) : v;
            const int qS = kReverse ? (r * w + c - 1) : (r * w + c + 1);      // (wraps at c == 0 in the reverse pass)
            const int qV = kReverse ? ((r - 1) * w + c) : ((r + 1) * w + c);
            const int mS = ((sm >> lane) & 1u) ? int(pid[qS]) : -1;
            const int mv = ((vm >> lane) & 1u) ? int(pid[qV]) : -1;
            unsigned todoS = sm, todoV = vm;
            while (todoS | todoV) {
                const int src = __ffs(todoS | todoV) - 1;
                const int mine = mS >= 0 ? mS : mv;
                const int mm = __shfl_sync(SPX_FULL, mine, src);
                const unsigned bS = __ballot_sync(SPX_FULL, mS == mm), bV = __ballot_sync(SPX_FULL, mv == mm);
                const int n_here = __popc(bS) + __popc(bV);
                if (kCount) {
                    if (lane == 0) table[i * nm + mm] += unsigned(n_here);
                } else {
                    const unsigned lt = (1u << lane) - 1u;
                    const int before = __popc(bS & lt) + __popc(bV & lt);
                    const int b0 = int(table[i * nm + mm]);      // claims of this model before this chunk (whole pass)
                    if (mS == mm) pos[qS] = pos_base[mm] + b0 + before;
                    if (mv == mm) pos[qV] = pos_base[mm] + b0 + before + (mS == mm ? 1 : 0);
                    const int hl = 31 - __clz(bS | bV);
                    __syncwarp();
                    if (lane == hl) {
                        table[i * nm + mm] = unsigned(b0 + n_here);
                        // the last claim of the pass: largest (row, visiting index, vertical-after-sideways) key
                        const unsigned key = unsigned(i) * unsigned(2 * w + 2) + unsigned(2 * v + (mv == mm ? 1 : 0));
                        const int tgt = (mv == mm) ? qV : qS;
                        atomicMax(&S.lastkey[pass][mm], (static_cast<unsigned long long>(key) << 32) | unsigned(tgt));
                    }
                    __syncwarp();
                }
                todoS &= ~bS; todoV &= ~bV;
            }
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) k_refine2(Params P, Buffers B) {
    extern __shared__ unsigned sm_ref[];
    __shared__ Refine2Smem S;
    const int f = P.frame0 + blockIdx.x;
    FrameCtl &ctl = B.ctl[f];
    const int nm = ctl.n_models;
    const int w = P.w, h = P.h;
    if (nm == 0 || !P.refine_fast || !refine_fast_ok(h, nm)) return;
    unsigned *Hb = sm_ref, *Vb = Hb + h * NW;
    const bool has_invalid = (ctl.flags & unsigned(SPX_FRAME_NONFINITE)) != 0u;
    unsigned *table = Vb + h * NW;
    const int tid = threadIdx.x;
    unsigned *rowmask = table + kRefTableCap;         // [h]: chunks of a row with claimed pixels
    volatile signed char *carryq = reinterpret_cast<volatile signed char *>(rowmask + h), *rowlastq = carryq + h * NW;
    auto reset_handover = [&]() {
        for (int i = tid; i < h * (NW + 1); i += NW * 32) carryq[i] = (signed char)kRefPending;
        for (int i = tid; i < h; i += NW * 32) rowmask[i] = 0u;
    };
    const size_t fo = size_t(f) * P.N;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    int8_t *pid = B.pid + fo;
    int *pos = B.pos + fo;
    for (int m = tid; m < nm; m += NW * 32) {
        const Model &M = ctl.models[m];
        S.coef[m][0] = M.coef[0]; S.coef[m][1] = M.coef[1]; S.coef[m][2] = M.coef[2]; S.coef[m][3] = M.coef[3];
        S.n0[m] = M.n0; S.cnt[0][m] = 0; S.cnt[1][m] = 0; S.lastkey[0][m] = 0ull; S.lastkey[1][m] = 0ull;
    }
    for (int i = tid; i < h * nm; i += NW * 32) table[i] = 0;
    reset_handover();
    __syncthreads();

    // ---------------- pass 1 ----------------
    refine2_propagate<NW, false>(S, Hb, Vb, rowmask, carryq, rowlastq, P, px, py, pz, pid, has_invalid);
    __threadfence_block();
    __syncthreads();
    refine2_emit<NW, false, true>(S, Hb, Vb, rowmask, table, nm, P, pid, pos, S.n0, 0);
    __syncthreads();
    for (int m = tid; m < nm; m += NW * 32) {     // exclusive prefix over the rows
        int run = 0;
        for (int i = 0; i < h; ++i) { const int t = int(table[i * nm + m]); table[i * nm + m] = unsigned(run); run += t; }
        S.cnt[0][m] = run;
    }
    __syncthreads();
    refine2_emit<NW, false, false>(S, Hb, Vb, rowmask, table, nm, P, pid, pos, S.n0, 0);
    __syncthreads();

    // ---------------- pass 2 ----------------
    // positions of pass-2 claims start after the originals and the pass-1 claims
    for (int m = tid; m < nm; m += NW * 32) S.n0[m] += S.cnt[0][m];
    for (int i = tid; i < h * nm; i += NW * 32) table[i] = 0;
    // rows whose first visited pixel (column w-1) could take the wrap claim of the row below: it is free and within
    // 0.02 m of some model plane.  Such rows are common (a free pixel on the right image border that is coplanar with
    // any model), real wrap claims are rare (the model must also end up at column 0 of the row below), so the pass
    // runs SPECULATIVELY on the plain skewed schedule; afterwards every candidate row is checked with the final
    // labels, and only if a wrap claim would have fired the labels are restored and the pass is redone with the
    // candidate rows waiting for the row below to complete.
    int8_t *bak = B.pid_bak + fo;
    int n_cand_local = 0;
    for (int r = tid; r < h; r += NW * 32) {
        bool fl = false;
        const int q = r * w + w - 1;
        if (r <= h - 2 && pid[q] == -1) {
            const float x = px[q], y = py[q], z = pz[q];
            for (int m = 0; m < nm && !fl; ++m) fl = refine_dist_ok(S.coef[m], x, y, z);
        }
        S.wrapcand[r] = fl ? 1 : 0;
        S.wrapflag[r] = 0;
        n_cand_local += fl ? 1 : 0;
    }
    const int any_cand = __syncthreads_or(n_cand_local);
    if (any_cand) {
        if ((P.N & 15) == 0) { for (int i = tid; i < P.N / 16; i += NW * 32) reinterpret_cast<uint4 *>(bak)[i] = reinterpret_cast<const uint4 *>(pid)[i]; }
        else { for (int i = tid; i < P.N; i += NW * 32) bak[i] = pid[i]; }
    }
    reset_handover();
    __syncthreads();
    refine2_propagate<NW, true>(S, Hb, Vb, rowmask, carryq, rowlastq, P, px, py, pz, pid, has_invalid);
    __threadfence_block();
    __syncthreads();
    if (any_cand) {
        // a wrap claim fires at row r iff (r, w-1) stayed free and the final label m of (r+1, 0) is a model with
        // |n.p + d| < 0.02 at (r, w-1).  Rows below the first such row are correct as computed, so the test is exact.
        int viol = 0;
        for (int r = tid; r < h - 1; r += NW * 32) {
            if (S.wrapcand[r]) {
                const int q = r * w + w - 1;
                const int m0 = int(pid[(r + 1) * w]);
                if (pid[q] == -1 && m0 >= 0 && refine_dist_ok(S.coef[m0], px[q], py[q], pz[q])) viol = 1;
            }
        }
        if (__syncthreads_or(viol)) {
            if ((P.N & 15) == 0) { for (int i = tid; i < P.N / 16; i += NW * 32) reinterpret_cast<uint4 *>(pid)[i] = reinterpret_cast<const uint4 *>(bak)[i]; }
            else { for (int i = tid; i < P.N; i += NW * 32) pid[i] = bak[i]; }
            for (int r = tid; r < h; r += NW * 32) S.wrapflag[r] = S.wrapcand[r];
            reset_handover();
            __threadfence_block();
            __syncthreads();
            refine2_propagate<NW, true>(S, Hb, Vb, rowmask, carryq, rowlastq, P, px, py, pz, pid, has_invalid);
            __threadfence_block();
            __syncthreads();
        }
    }
    refine2_emit<NW, true, true>(S, Hb, Vb, rowmask, table, nm, P, pid, pos, S.n0, 1);
    __syncthreads();
    for (int m = tid; m < nm; m += NW * 32) {
        int run = 0;
        for (int i = 0; i < h; ++i) { const int t = int(table[i * nm + m]); table[i * nm + m] = unsigned(run); run += t; }
        S.cnt[1][m] = run;
    }
    __syncthreads();
    refine2_emit<NW, true, false>(S, Hb, Vb, rowmask, table, nm, P, pid, pos, S.n0, 1);
    __syncthreads();
    for (int m = tid; m < nm; m += NW * 32) {
        Model &M = ctl.models[m];
        M.n1 = S.cnt[0][m]; M.n2 = S.cnt[1][m];
        if (S.cnt[1][m] > 0) M.last_inlier = int(S.lastkey[1][m] & 0xffffffffull);
        else if (S.cnt[0][m] > 0) M.last_inlier = int(S.lastkey[0][m] & 0xffffffffull);
    }
}

// K7: findLabeledRegionBoundary from inlier_indices[i].back().  The Moore walk is serial per model, so it is made
// short instead: one CTA per frame stages the plane-id map in shared memory with a one-pixel sentinel frame (no bounds
// tests), thread m walks model m, and every step reads the 8 neighbours at once (independent loads) and picks the
// next direction from an 8-bit "same label" mask with a rotate + find-first-set.
__constant__ int c_ddx[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
__constant__ int c_ddy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
constexpr int kContourThreads = 256;
constexpr int kPidOutside = 0x7f;   // sentinel of the frame around the image: neither "same" nor an in-image "different"

__global__ void __launch_bounds__(kContourThreads) k_contour(Params P, Buffers B) {
    extern __shared__ int8_t sm_pid[];   // (h + 2) x (w + 2)
    const int f = P.frame0 + blockIdx.x, tid = threadIdx.x;
    FrameCtl &ctl = B.ctl[f];
    const int nm = ctl.n_models;
    if (nm == 0) return;
    const int w = P.w, h = P.h, sw = w + 2;
    const size_t fo = size_t(f) * P.N;
    const int8_t *pid = B.pid + fo;
    // a warp per row (no division per element); the sentinel frame separately
    for (int y = tid >> 5; y < h; y += kContourThreads / 32) {
        const int8_t *src = pid + y * w;
        int8_t *dst = sm_pid + (y + 1) * sw + 1;
        for (int x = tid & 31; x < w; x += 32) dst[x] = src[x];
    }
    for (int i = tid; i < sw; i += kContourThreads) { sm_pid[i] = int8_t(kPidOutside); sm_pid[(h + 1) * sw + i] = int8_t(kPidOutside); }
    for (int y = tid; y < h; y += kContourThreads) { sm_pid[(y + 1) * sw] = int8_t(kPidOutside); sm_pid[(y + 1) * sw + w + 1] = int8_t(kPidOutside); }
    __syncthreads();
    int *arena = B.contour_idx + size_t(f) * P.contour_cap;
    // offsets of the neighbours in the padded map, in the reference's direction order
    int noff[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) noff[d] = c_ddy[d] * sw + c_ddx[d];
    for (int m = tid; m < nm; m += kContourThreads) {
        Model &M = ctl.models[m];
        int off = 0;
        for (int k = 0; k < m; ++k) { const Model &K = ctl.models[k]; off += 2 * (K.n0 + K.n1 + K.n2) + 16; }
        const int cap = 2 * (M.n0 + M.n1 + M.n2) + 16;
        M.contour_off = off;
        int *out = arena + off;
        const int start = M.last_inlier;
        int cx = start % w, cy = start / w;
        int sp = (cy + 1) * sw + cx + 1;           // position in the padded map
        const int sp_start = sp;
        auto masks = [&](int p, unsigned &same, unsigned &diff) {
            same = 0u; diff = 0u;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int v = sm_pid[p + noff[d]];
                same |= (v == m ? 1u : 0u) << d;
                diff |= ((v != m && v != kPidOutside) ? 1u : 0u) << d;
            }
        };
        unsigned same, diff;
        masks(sp, same, diff);
        int n = 0;
        if (diff != 0u) {                          // first direction whose in-image neighbour carries another label
            int direction = __ffs(diff) - 1;
            out[n++] = start;
            const long long guard = 8ll * P.N + 8;
            long long steps = 0;
            bool overflow = false;
            do {
                // for d = 1..8: nIdx = (direction + d) & 7, stop at the first neighbour with the same label
                const unsigned rot = ((same | (same << 8)) >> ((direction + 1) & 7)) & 0xffu;
                const int nIdx = rot ? ((direction + 1 + (__ffs(rot) - 1)) & 7) : direction;
                direction = (nIdx + 4) & 7;
                sp += noff[nIdx];
                cx += c_ddx[nIdx]; cy += c_ddy[nIdx];
                if (++steps > guard || cx < 0 || cx >= w || cy < 0 || cy >= h) { overflow = true; break; }
                if (n < cap) out[n++] = cy * w + cx; else overflow = true;
                masks(sp, same, diff);
            } while (sp != sp_start);
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
    const int fl = blockIdx.x * blockDim.x + threadIdx.x;
    if (fl >= P.n_frames) return;
    const int f = P.frame0 + fl;
    FrameCtl &ctl = B.ctl[f];
    int np = 0, poff = 0, boff = 0;
    for (int i = 0; i < ctl.n_models; ++i) {
        Model &M = ctl.models[i];
        float coef[4] = {M.coef[0], M.coef[1], M.coef[2], M.coef[3]};
        if (coef[3] < 0) { coef[0] = -coef[0]; coef[1] = -coef[1]; coef[2] = -coef[2]; coef[3] = -coef[3]; }
        M.plane = -1; M.n_rounds = 0;
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
        if (P.enable_supposed && M.n_contour >= 50) B.work[2 + atomicAdd(&B.work[0], 1)] = f * SPX_MAX_MODELS + i;   // a k_lines item
        poff += npts; boff += nb;
        ++np;
    }
    ctl.n_real = np; ctl.n_planes = np; ctl.pts_used = poff; ctl.bnd_used = boff; ctl.pts_sup = 0; ctl.bnd_sup = 0; ctl.n_lines = 0;
}

// ExtractIndices(negative = false) of the kept planes: inlier points in inlier_indices order (src/Frame.cc:925-928)
__global__ void __launch_bounds__(256) k_pack_points(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    const int m = B.pid[fo + q];
    if (m < 0) return;
    const FrameCtl &ctl = B.ctl[f];
    const int k = ctl.models[m].plane;
    if (k < 0) return;
    const long long o = B.frame_offs[size_t(f) * 5 + 1] + ctl.planes[k].points_off + B.pos[fo + q];
    if (P.compact) {
        // compact results: the cloud is a pure function of (index, depth), so only inlier_indices[k] leaves the device
        if (P.idx16) static_cast<uint16_t *>(B.out_pidx)[o] = uint16_t(q); else static_cast<uint32_t *>(B.out_pidx)[o] = uint32_t(q);
        return;
    }
    spx_point pt;
    pt.x = B.px[fo + q]; pt.y = B.py[fo + q]; pt.z = B.pz[fo + q]; pt.rgba = pack_rgba(0, 0, 250);
    B.out_pts[o] = pt;
}

// regions[i].getContour() of the kept planes (src/Frame.cc:930-932); one CTA per (model, frame)
__global__ void __launch_bounds__(128) k_pack_contours(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y, m = blockIdx.x;
    const FrameCtl &ctl = B.ctl[f];
    if (m >= ctl.n_models) return;
    const Model &M = ctl.models[m];
    if (M.plane < 0 || M.n_contour == 0) return;
    const size_t fo = size_t(f) * P.N;
    const int *src = B.contour_idx + size_t(f) * P.contour_cap + M.contour_off;
    spx_point *dst = B.out_bnd + B.frame_offs[size_t(f) * 5 + 2] + ctl.planes[M.plane].boundary_off;
    for (int j = threadIdx.x; j < M.n_contour; j += blockDim.x) {
        const int q = src[j];
        spx_point pt;
        pt.x = B.px[fo + q]; pt.y = B.py[fo + q]; pt.z = B.pz[fo + q]; pt.rgba = pack_rgba(0, 0, 250);
        dst[j] = pt;
    }
}

}  // namespace spx
