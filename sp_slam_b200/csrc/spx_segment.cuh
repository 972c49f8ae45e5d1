// spx_segment.cuh -- K4..K5: comparator connected components (union-find with atomics), PCL label ranks,
// per-component fp32 moments in the reference's accumulation order, closed-form plane fit, curvature filter.
//
// Reference: /root/reference/src/Frame.cc:898-905 (OrganizedMultiPlaneSegmentation settings) and PCL 1.8.0
// segmentation/impl/organized_multi_plane_segmentation.hpp (segment), segmentation/plane_coefficient_comparator.h
// (compare), segmentation/impl/organized_connected_component_segmentation.hpp (segment),
// common/impl/centroid.hpp (computeMeanAndCovarianceMatrix), common/impl/eigen.hpp (eigen33).
#pragma once
#include <climits>
#include "spx_math.cuh"
#include "spx_types.cuh"

namespace spx {

// ---------------------------------------------------------------------------------------------------------------
// union-find on pixel indices; a root is only ever attached below a smaller index, so the root of a finished
// component is its minimum pixel index = its first pixel in raster order (PCL: smaller run id wins).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int *parent, int x) {
    while (true) {
        const int p = __ldcg(parent + x);
        if (p == x) return x;
        x = p;
    }
}
// find with path halving: every visited node is re-pointed at its grandparent.  Only non-roots are written here and
// only roots are written by uf_unite's atomicMin (a non-root never becomes a root again), so the plain stores cannot
// undo a union; they only ever move a pointer further up its own tree.
__device__ __forceinline__ int uf_find_halving(int *parent, int x) {
    while (true) {
        const int p = __ldcg(parent + x);
        if (p == x) return x;
        const int gp = __ldcg(parent + p);
        if (gp == p) return p;
        __stcg(parent + x, gp);
        x = gp;
    }
}
__device__ __forceinline__ void uf_unite(int *parent, int a, int b) {
    while (true) {
        a = uf_find_halving(parent, a);
        b = uf_find_halving(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;
    }
}

// K4a: comparator edges + row-run initialisation.  A warp covers 32 consecutive columns of one row.
// PlaneCoefficientComparator::compare(idx1 = current pixel, idx2 = left / upper pixel):
//   |d1 - d2| < DisTh * z1^2  &&  n1 . n2 > cos(AngTh)          (depth dependent threshold on the CURRENT pixel)
__global__ void __launch_bounds__(256) k_ccl_link(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.z, lane = threadIdx.x;
    const int r = blockIdx.y * 8 + threadIdx.y;
    const int c = blockIdx.x * 32 + lane;
    const int w = P.w;
    if (r >= P.h) return;
    const bool valid = c < w;
    const size_t fo = size_t(f) * P.N;
    bool L = false, U = false;
    const int q = r * w + c;
    if (valid) {
        const float d1 = B.pd[fo + q], n1x = B.nx[fo + q], n1y = B.ny[fo + q], n1z = B.nz[fo + q];
        const float X = B.px[fo + q], Y = B.py[fo + q], Zv = B.pz[fo + q];
        const float z = X * 0.0f + (Y * 0.0f + Zv * 1.0f);   // vec.dot(z_axis_)
        float threshold = P.dist_thr;
        threshold *= z * z;
        if (c >= 1) {
            const int o = q - 1;
            L = (fabsf(d1 - B.pd[fo + o]) < threshold) &&
                (dot3f(n1x, n1y, n1z, B.nx[fo + o], B.ny[fo + o], B.nz[fo + o]) > P.ang_cos);
        }
        if (r >= 1) {
            const int o = q - w;
            U = (fabsf(d1 - B.pd[fo + o]) < threshold) &&
                (dot3f(n1x, n1y, n1z, B.nx[fo + o], B.ny[fo + o], B.nz[fo + o]) > P.ang_cos);
        }
        B.conn[fo + q] = uint8_t((L ? 1 : 0) | (U ? 2 : 0));
        B.cnt[fo + q] = isfinite(X) ? 0 : INT_MIN;
    }
    const unsigned linked = __ballot_sync(SPX_FULL, valid && L);
    const unsigned starts = ~linked | 1u;                       // lane 0 always starts a run inside the segment
    const int s = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
    if (valid) B.parent[fo + q] = r * w + blockIdx.x * 32 + s;
}

// K4b: joins the tile-local forests of k_normals_link across tile borders: left links on the first column of a
// tile, upper links on its first row (tile_w x tile_h = 32 x 16).  With tile_h = 1 (the feed-normals path, where
// k_ccl_link only forms row runs) every upper link is a border.  A vertical union is skipped when the two pixels are
// already connected through their left neighbours (L(q) & L(up) & U(left)).
__global__ void __launch_bounds__(256) k_ccl_merge(Params P, Buffers B, int tile_h) {
    const int f = P.frame0 + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const int w = P.w;
    const size_t fo = size_t(f) * P.N;
    const int r = q / w, c = q - r * w;
    const unsigned cb = B.conn[fo + q];
    int *parent = B.parent + fo;
    if ((cb & 1u) && (c & 31) == 0) {
        // redundant when the pair is already joined through the row above: U(q) & U(left) & L(up)
        const bool skip = r > 0 && (cb & 2u) && (B.conn[fo + q - 1] & 2u) && (B.conn[fo + q - w] & 1u);
        if (!skip) uf_unite(parent, q, q - 1);
    }
    if ((cb & 2u) && (r % tile_h) == 0) {
        const bool skip = (c & 31) != 0 && (cb & 1u) && (B.conn[fo + q - w] & 1u) && (B.conn[fo + q - 1] & 2u);
        if (!skip) uf_unite(parent, q, q - w);
    }
}

// K4c: path compression to the root + component sizes (warp-aggregated atomics on the root's counter)
__global__ void __launch_bounds__(256) k_ccl_flatten(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < P.N;
    const size_t fo = size_t(f) * P.N;
    int root = -1 - int(threadIdx.x & 31);
    if (valid) {
        root = uf_find(B.parent + fo, q);
        B.parent[fo + q] = root;
    }
    const unsigned peers = __match_any_sync(SPX_FULL, root);
    if (valid && (__ffs(peers) - 1) == int(threadIdx.x & 31)) atomicAdd(B.cnt + fo + root, __popc(peers));
}

// K4b with four consecutive pixels per thread (organized clouds with N % 4 == 0): the link bytes of the four pixels and
// of the four pixels above them arrive as one 32-bit word each (the upper row is not word aligned: two aligned words
// and a funnel shift), and a thread whose four pixels have no links leaves at once.  Same unions, same skip rules, same
// result as k_ccl_merge with tile_h = 1 (0.18 vs 0.32 ms per 1000 frames).  The same treatment of the flatten pass was
// slower (0.41 vs 0.33 ms: four warp matches per thread instead of one, fewer threads to hide the pointer chases).
__global__ void __launch_bounds__(256) k_ccl_merge4(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q0 >= P.N) return;
    const int w = P.w;
    const size_t fo = size_t(f) * P.N;
    const uint8_t *conn = B.conn + fo;
    int *parent = B.parent + fo;
    const unsigned cb4 = *reinterpret_cast<const unsigned *>(conn + q0);
    if (!(cb4 & 0x03030303u)) return;                       // no links at all in these four pixels
    unsigned up4 = 0;
    if (q0 >= w) {                                          // link bytes of q0 - w .. q0 - w + 3 (inside the frame)
        const int a = q0 - w, al = a & ~3;
        const unsigned lo = *reinterpret_cast<const unsigned *>(conn + al);
        const unsigned hi = (a & 3) ? *reinterpret_cast<const unsigned *>(conn + al + 4) : 0u;
        up4 = __funnelshift_r(lo, hi, 8 * (a & 3));
    } else if (q0 + 3 >= w) {                               // the four pixels straddle rows 0 and 1
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q0 + k >= w) up4 |= unsigned(conn[q0 + k - w]) << (8 * k);
    }
    unsigned left = q0 > 0 ? conn[q0 - 1] : 0u;
    int r = q0 / w, c = q0 - r * w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q = q0 + k;
        const unsigned cb = (cb4 >> (8 * k)) & 0xffu, up = (up4 >> (8 * k)) & 0xffu;
        if ((cb & 1u) && (c & 31) == 0) {
            const bool skip = r > 0 && (cb & 2u) && (left & 2u) && (up & 1u);
            if (!skip) uf_unite(parent, q, q - 1);
        }
        if (cb & 2u) {
            const bool skip = (c & 31) != 0 && (cb & 1u) && (up & 1u) && (left & 2u);
            if (!skip) uf_unite(parent, q, q - w);
        }
        left = cb;
        if (++c == w) { c = 0; ++r; }
    }
}

// K4c by row runs.  The forest was initialised with one tree per row run (maximal chain of left links inside a 32-column
// segment: k_normals_link / k_ccl_link) and unions only ever merge trees, so all pixels of a run share their root: a warp
// takes one 32-column segment of a row, rebuilds the runs from the link bits with a ballot, only the first pixel of a run
// chases the pointers and adds the run's length to the root's counter, and the root reaches the run's lanes by shuffle.
constexpr int kFlatRows = 4;        // row segments per warp: their pointer chases are in flight together
__global__ void __launch_bounds__(256) k_ccl_flatten_runs(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.z, lane = threadIdx.x;
    const int r0 = (blockIdx.y * 8 + threadIdx.y) * kFlatRows;
    const int c = blockIdx.x * 32 + lane;
    const int w = P.w;
    if (r0 >= P.h) return;                                  // warp uniform
    const size_t fo = size_t(f) * P.N;
    int *parent = B.parent + fo;
    const int n_valid = min(32, w - blockIdx.x * 32);
    const bool cvalid = c < w;
    unsigned cb[kFlatRows];
#pragma unroll
    for (int k = 0; k < kFlatRows; ++k) cb[k] = (cvalid && r0 + k < P.h) ? B.conn[fo + (r0 + k) * w + c] : 0u;
    int s0[kFlatRows], len[kFlatRows], x[kFlatRows];
    bool start[kFlatRows], done[kFlatRows];
#pragma unroll
    for (int k = 0; k < kFlatRows; ++k) {
        const bool valid = cvalid && r0 + k < P.h;
        const unsigned linked = __ballot_sync(SPX_FULL, valid && (cb[k] & 1u));
        const unsigned starts = ~linked | 1u;               // lane 0 always starts a run inside the segment
        s0[k] = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
        const unsigned above = lane == 31 ? 0u : (starts & (SPX_FULL << (lane + 1)));
        const int next = above ? __ffs(above) - 1 : 32;
        start[k] = valid && s0[k] == lane;
        len[k] = min(next, n_valid) - lane;
        x[k] = (r0 + k) * w + c;
        done[k] = !start[k];
    }
    // the run starts chase their pointers, all rows of the warp together
    while (true) {
        bool all = true;
        int nx[kFlatRows];
#pragma unroll
        for (int k = 0; k < kFlatRows; ++k) nx[k] = done[k] ? x[k] : __ldcg(parent + x[k]);
#pragma unroll
        for (int k = 0; k < kFlatRows; ++k) {
            if (!done[k]) { done[k] = nx[k] == x[k]; x[k] = nx[k]; }
            all = all && done[k];
        }
        if (__all_sync(SPX_FULL, all)) break;
    }
#pragma unroll
    for (int k = 0; k < kFlatRows; ++k) {
        const int root = __shfl_sync(SPX_FULL, x[k], s0[k]);
        if (cvalid && r0 + k < P.h) parent[(r0 + k) * w + c] = root;
        if (start[k]) atomicAdd(B.cnt + fo + root, len[k]);
    }
}

// K4b + K4c in shared memory: one CTA per frame whose forest fits (16-bit entries: N <= 65535 and 2 N bytes of shared memory).
// The unions of k_ccl_merge4 and the pointer chases of k_ccl_flatten_runs are chains of dependent loads; in global memory
// every hop is an L2 round trip, here it is a shared-memory access.  Same forest, same result: a union points the larger
// root at the smaller, so every root is its component's first raster pixel whatever the order of the unions; the output
// is the flattened forest (parent[q] = root) and the component sizes added to cnt[root], exactly what the two kernels leave.
//  (1) a warp takes a 32-column segment of a block of rows and rebuilds the row runs from the left-link bits (the forest
//      k_normals_strip / k_ccl_link initialised in global memory is not read);
//  (2) vertical unions and the left links across segment borders, with k_ccl_merge's skip rules;
//  (3) per run: the first pixel finds the root, the run's lanes get it by shuffle, the length goes to the root's counter.
constexpr int kCFThreads = 448;      // 14 warps: 7 segments x 2 row blocks at 214 x 160
constexpr int kCFRows = 40;          // rows of an item (segment x row block)

__device__ __forceinline__ unsigned sm_find(volatile unsigned short *lab, unsigned x) {   // path halving (see uf_find_halving)
    while (true) {
        const unsigned p = lab[x];
        if (p == x) return x;
        const unsigned gp = lab[p];
        if (gp == p) return p;
        lab[x] = static_cast<unsigned short>(gp);
        x = gp;
    }
}
__device__ __noinline__ void sm_unite(unsigned short *lab, unsigned a, unsigned b) {
    while (true) {
        a = sm_find(lab, a);
        b = sm_find(lab, b);
        if (a == b) return;
        if (a < b) { const unsigned t = a; a = b; b = t; }
        // a is a root as far as this thread knows: point it at the smaller root unless somebody else already did
        if (atomicCAS(lab + a, static_cast<unsigned short>(a), static_cast<unsigned short>(b)) == a) return;
    }
}

__host__ __device__ __forceinline__ bool ccl_frame_fits(int N) { return N <= 65535 && size_t(N) * 2 <= size_t(72) * 1024; }

__global__ void __launch_bounds__(kCFThreads) k_ccl_frame(Params P, Buffers B) {
    extern __shared__ unsigned short sm_lab[];   // N entries
    const int f = P.frame0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int w = P.w, h = P.h;
    const size_t fo = size_t(f) * P.N;
    const uint8_t *conn = B.conn + fo;
    const int segs = (w + 31) >> 5, rblocks = (h + kCFRows - 1) / kCFRows, items = segs * rblocks;
    constexpr int kWarps = kCFThreads / 32;
    // The link bytes are read eight rows at a time (independent loads in flight together), the chains run in shared memory.
    constexpr int kU = 8;
    // (1) row runs
    for (int it = wid; it < items; it += kWarps) {
        const int seg = it % segs, rb = it / segs;
        const int c = seg * 32 + lane;
        const bool valid = c < w;
        const int r1 = min(h, (rb + 1) * kCFRows);
        for (int rr = rb * kCFRows; rr < r1; rr += kU) {
            unsigned long long pack = 0ull;                         // the eight link bytes, one per row
#pragma unroll
            for (int u = 0; u < kU; ++u) pack |= static_cast<unsigned long long>((valid && rr + u < r1) ? conn[(rr + u) * w + c] : 0u) << (8 * u);
#pragma unroll 1
            for (int u = 0; u < kU; ++u, pack >>= 8) {
                if (rr + u >= r1) break;                            // (warp uniform)
                const int q = (rr + u) * w + c;
                const unsigned linked = __ballot_sync(SPX_FULL, valid && (unsigned(pack) & 1u));
                const unsigned starts = ~linked | 1u;               // lane 0 always starts a run inside the segment
                const int s0 = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
                if (valid) sm_lab[q] = static_cast<unsigned short>(q - lane + s0);
            }
        }
    }
    __syncthreads();
    // (2) unions (k_ccl_merge's rules: a link whose union is implied by the three other links of its 2 x 2 cell is skipped)
    for (int it = wid; it < items; it += kWarps) {
        const int seg = it % segs, rb = it / segs;
        const int c = seg * 32 + lane;
        const bool valid = c < w;
        const int rbeg = rb * kCFRows, r1 = min(h, (rb + 1) * kCFRows);
        unsigned up = (valid && rbeg > 0) ? conn[(rbeg - 1) * w + c] : 0u;
        for (int rr = rbeg; rr < r1; rr += kU) {
            unsigned long long pack = 0ull;                         // the eight link bytes, one per row
#pragma unroll
            for (int u = 0; u < kU; ++u) pack |= static_cast<unsigned long long>((valid && rr + u < r1) ? conn[(rr + u) * w + c] : 0u) << (8 * u);
#pragma unroll 1
            for (int u = 0; u < kU; ++u, pack >>= 8) {
                if (rr + u >= r1) break;
                const int r = rr + u, q = r * w + c;
                const unsigned cb = unsigned(pack) & 0xffu;
                unsigned left = __shfl_up_sync(SPX_FULL, cb, 1);
                if (lane == 0) {
                    left = 0u;
                    if ((cb & 1u) && r > 0 && (cb & 2u)) left = conn[q - 1];     // only the skip test of the border link reads it
                    if (cb & 1u) {                                                // (c > 0: column 0 has no left link)
                        const bool skip = r > 0 && (cb & 2u) && (left & 2u) && (up & 1u);
                        if (!skip) sm_unite(sm_lab, unsigned(q), unsigned(q - 1));
                    }
                }
                if (cb & 2u) {
                    const bool skip = lane != 0 && (cb & 1u) && (up & 1u) && (left & 2u);
                    if (!skip) sm_unite(sm_lab, unsigned(q), unsigned(q - w));
                }
                up = cb;
            }
        }
    }
    __syncthreads();
    // (3) flatten by runs, sizes to the roots
    int *parent = B.parent + fo;
    int *cnt = B.cnt + fo;
    for (int it = wid; it < items; it += kWarps) {
        const int seg = it % segs, rb = it / segs;
        const int c = seg * 32 + lane;
        const bool valid = c < w;
        const int n_valid = min(32, w - seg * 32);
        const int r1 = min(h, (rb + 1) * kCFRows);
        for (int rr = rb * kCFRows; rr < r1; rr += kU) {
            unsigned long long pack = 0ull;                         // the eight link bytes, one per row
#pragma unroll
            for (int u = 0; u < kU; ++u) pack |= static_cast<unsigned long long>((valid && rr + u < r1) ? conn[(rr + u) * w + c] : 0u) << (8 * u);
#pragma unroll 1
            for (int u = 0; u < kU; ++u, pack >>= 8) {
                if (rr + u >= r1) break;
                const int q = (rr + u) * w + c;
                const unsigned linked = __ballot_sync(SPX_FULL, valid && (unsigned(pack) & 1u));
                const unsigned starts = ~linked | 1u;
                const int s0 = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
                const unsigned above = lane == 31 ? 0u : (starts & (SPX_FULL << (lane + 1)));
                const int next = above ? __ffs(above) - 1 : 32;
                const bool start = valid && s0 == lane;
                int root = 0;
                if (start) root = int(sm_find(sm_lab, unsigned(q)));
                root = __shfl_sync(SPX_FULL, root, s0);
                if (valid) parent[q] = root;
                if (start) atomicAdd(cnt + root, min(next, n_valid) - lane);
            }
        }
    }
}

// K4d: one CTA per frame, one pass over the frame in 2048-pixel chunks (raster order).
//  (1) exclusive prefix count of roots = PCL's dense label of each component; components with size > Plane.MinSize
//      become plane candidates, in label order, and get the offset of their index list;
//  (2) order-preserving multi-way compaction: every pixel of a candidate component is appended to that candidate's
//      raster-ordered index list (label_indices[l] of PCL) and learns its position in it.  A component's root is its
//      first raster pixel, so the candidate is always known before (or in the same chunk as) its members.
// Every thread owns 4 consecutive pixels of a 2048-pixel chunk, so one block scan and one set of barriers serve four
// pixels, and the ranking of a candidate's members inside a warp is a small-count scan per candidate present.
constexpr int kRankThreads = 512;
constexpr int kRankPer = 4;
constexpr int kRankChunk = kRankThreads * kRankPer;
constexpr int kRankWarps = kRankThreads / 32;

__global__ void __launch_bounds__(kRankThreads) k_ccl_rank(Params P, Buffers B) {
    __shared__ unsigned warp_tot[kRankWarps];
    __shared__ unsigned running_s, block_tot;
    __shared__ int s_off[SPX_MAX_CAND];          // start of the candidate's index list
    __shared__ int s_cnt[SPX_MAX_CAND];          // members emitted so far
    __shared__ int s_size[SPX_MAX_CAND];
    __shared__ int s_wcnt[SPX_MAX_CAND][kRankWarps];   // members of candidate c found by warp w in this chunk -> their base
    __shared__ int s_ncand, s_noff;
    const int f = P.frame0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t fo = size_t(f) * P.N;
    FrameCtl &ctl = B.ctl[f];
    const int *parent = B.parent + fo;
    const int *cntp = B.cnt + fo;
    int16_t *root_cand = B.root_model + fo;      // at roots: candidate index, -1, or -2 for an unlabelled point (re-used for model ids by k_models)
    int *cand_idx = B.cand_idx + fo;
    int *pos = B.pos + fo;
    if (tid == 0) { running_s = 0; s_ncand = 0; s_noff = 0; }
    __syncthreads();
    // software pipeline: the next chunk's forest entries and sizes are loaded while this chunk is processed, and the
    // candidate id of a pixel whose root lies in an EARLIER chunk (final by then) is fetched ahead as well
    int root_n[kRankPer], cnt_n[kRankPer], cand_n[kRankPer];
#pragma unroll
    for (int j = 0; j < kRankPer; ++j) {
        const int q = tid * kRankPer + j;
        root_n[j] = q < P.N ? parent[q] : -1;
        cnt_n[j] = q < P.N ? cntp[q] : 0;
        cand_n[j] = -1;
    }
    for (int base = 0; base < P.N; base += kRankChunk) {
        const int q0 = base + tid * kRankPer;
        int root[kRankPer], sz[kRankPer], cand_pre[kRankPer];
#pragma unroll
        for (int j = 0; j < kRankPer; ++j) {
            root[j] = root_n[j]; sz[j] = cnt_n[j]; cand_pre[j] = cand_n[j];
            const int qn = q0 + kRankChunk + j;
            root_n[j] = qn < P.N ? parent[qn] : -1;
            cnt_n[j] = qn < P.N ? cntp[qn] : 0;
        }
        // roots are counted in bits 0..19 (N < 2^20), candidates in bits 20..31
        unsigned v[kRankPer], vt = 0u;
        bool nolabel[kRankPer];
#pragma unroll
        for (int j = 0; j < kRankPer; ++j) {
            const int q = q0 + j;
            bool isroot = q < P.N && root[j] == q;
            // PCL skips points with a non-finite x: they get no label at all.  They are singletons of the forest (every
            // comparison with a NaN fails) whose size counter was preset to INT_MIN by the link kernel
            nolabel[j] = isroot && sz[j] < 0;
            if (nolabel[j]) isroot = false;
            const bool iscand = isroot && unsigned(sz[j]) > unsigned(P.min_size);
            v[j] = (isroot ? 1u : 0u) | (iscand ? (1u << 20) : 0u);
            vt += v[j];
        }
        unsigned incl = vt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(SPX_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned t = lane < kRankWarps ? warp_tot[lane] : 0u;
            unsigned sc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(SPX_FULL, sc, o);
                if (lane >= o) sc += u;
            }
            if (lane < kRankWarps) warp_tot[lane] = sc - t;   // exclusive prefix of the warp totals
            if (lane == 31) block_tot = sc;
        }
        __syncthreads();
        const unsigned run_before = running_s;   // thread 0 advances it after the next barrier
        const unsigned chunk_tot = block_tot;
        unsigned excl = run_before + warp_tot[wid] + incl - vt;
#pragma unroll
        for (int j = 0; j < kRankPer; ++j) {
            const int q = q0 + j;
            if (nolabel[j]) { B.lab[fo + q] = -1; root_cand[q] = -2; atomicOr(&ctl.flags, unsigned(SPX_FRAME_NONFINITE)); }
            if (v[j] & 1u) {
                B.lab[fo + q] = int(excl & 0xFFFFFu);
                int16_t rc = -1;
                if (v[j] >> 20) {
                    const unsigned k = excl >> 20;
                    if (k < SPX_MAX_CAND) {
                        ctl.cand[k].root = q; ctl.cand[k].label = int(excl & 0xFFFFFu); ctl.cand[k].size = sz[j];
                        s_size[k] = sz[j];
                        rc = int16_t(k);
                    } else {
                        atomicOr(&ctl.flags, unsigned(SPX_FRAME_OVERFLOW));
                    }
                }
                root_cand[q] = rc;
            }
            excl += v[j];
        }
        __syncthreads();
        int nc_now = int((run_before + chunk_tot) >> 20);
        if (nc_now > SPX_MAX_CAND) nc_now = SPX_MAX_CAND;
        if (tid == 0) {
            int off = s_noff;
            for (int k = s_ncand; k < nc_now; ++k) { s_off[k] = off; s_cnt[k] = 0; ctl.cand[k].idx_off = off; off += s_size[k]; }
            s_noff = off; s_ncand = nc_now;
            running_s = run_before + chunk_tot;
        }
        for (int i = tid; i < nc_now * kRankWarps; i += kRankThreads) s_wcnt[i / kRankWarps][i % kRankWarps] = 0;
        __syncthreads();
        // every root below base + kRankChunk has its candidate id in memory now
#pragma unroll
        for (int j = 0; j < kRankPer; ++j)
            cand_n[j] = (root_n[j] >= 0 && root_n[j] < base + kRankChunk) ? int(root_cand[root_n[j]]) : -1;
        if (nc_now > 0) {   // uniform over the CTA
            // roots of earlier chunks were ranked before their id was prefetched; same-chunk roots are read now
            int c[kRankPer], rk[kRankPer];
#pragma unroll
            for (int j = 0; j < kRankPer; ++j) {
                c[j] = (q0 + j < P.N) ? ((root[j] < base) ? cand_pre[j] : int(root_cand[root[j]])) : -1;
                rk[j] = 0;
            }
            // rank inside the warp, one candidate value at a time (ascending): lane counts of 0..4, prefix over the lanes
            int donebelow = -1;   // candidates <= donebelow are ranked
            while (true) {
                int cmin = INT_MAX;
#pragma unroll
                for (int j = 0; j < kRankPer; ++j) if (c[j] > donebelow && c[j] < cmin) cmin = c[j];
                const int wmin = __reduce_min_sync(SPX_FULL, cmin);
                if (wmin == INT_MAX) break;
                int n_me = 0;
#pragma unroll
                for (int j = 0; j < kRankPer; ++j) { if (c[j] == wmin) { rk[j] = n_me; ++n_me; } }
                int sc = n_me;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(SPX_FULL, sc, o);
                    if (lane >= o) sc += u;
                }
#pragma unroll
                for (int j = 0; j < kRankPer; ++j) if (c[j] == wmin) rk[j] += sc - n_me;
                if (lane == 31) s_wcnt[wmin][wid] = sc;
                donebelow = wmin;
            }
            __syncthreads();
            for (int cc = wid; cc < nc_now; cc += kRankWarps) {
                const int t = lane < kRankWarps ? s_wcnt[cc][lane] : 0;
                int sc = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(SPX_FULL, sc, o);
                    if (lane >= o) sc += u;
                }
                const int b0 = s_cnt[cc];
                if (lane < kRankWarps) s_wcnt[cc][lane] = b0 + sc - t;
                __syncwarp();
                if (lane == 31) s_cnt[cc] = b0 + sc;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kRankPer; ++j) {
                if (c[j] >= 0) {
                    const int k = s_wcnt[c[j]][wid] + rk[j];
                    cand_idx[s_off[c[j]] + k] = q0 + j;
                    pos[q0 + j] = k;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        const unsigned tot = running_s;
        ctl.n_labels = int(tot & 0xFFFFFu) + 1;
        int nc = int(tot >> 20);
        if (nc > SPX_MAX_CAND) nc = SPX_MAX_CAND;
        ctl.n_cand = nc;
    }
}

// K4e: PCL label per pixel (labels.points[i].label before refine); only needed by the parity taps
__global__ void __launch_bounds__(256) k_ccl_label(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    const int root = B.parent[fo + q];
    if (root != q) B.lab[fo + q] = B.lab[fo + root];
}

// ---------------------------------------------------------------------------------------------------------------
// K5a: one warp per candidate.  computeMeanAndCovarianceMatrix accumulates nine fp32 moments over the component's
// pixels IN RASTER ORDER, one rounding per add; a tree reduction would round differently, so the chain is kept: the
// warp walks the candidate's dense index list 32 members at a time (coalesced index loads, gathered xyz, next batch
// prefetched while the current one is consumed), every lane forms the nine products of its member into shared
// memory, and lanes 0..8 each carry one accumulator through the 32 staged products in order.  Lane 0 then solves
// the 3x3 eigenproblem.  The time of the kernel is the longest chain: size x one dependent fp32 add.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMomBatch = 128;   // members per batch (4 per producer lane)
constexpr int kMomCands = 8;     // CTAs per frame; CTA c takes candidates c, c + kMomCands, ...

// three warps per candidate: warps 1 and 2 (producers, even / odd batches) gather 128 members at a time -- coalesced
// index loads, gathered xyz, each producer two of its own batches ahead -- and stage their nine products in shared
// memory; warp 0 (consumer) does nothing but the ordered chain (one LDS + one dependent FADD per member on lanes
// 0..8), so the chain runs at the latency of the add.
__global__ void __launch_bounds__(96) k_moments_fit(Params P, Buffers B) {
    __shared__ float s_prod[2][kMomBatch * 9];
    const int f = P.frame0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FrameCtl &ctl = B.ctl[f];
    const int n_cand = ctl.n_cand;
    const size_t fo = size_t(f) * P.N;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    for (int ci = blockIdx.x; ci < n_cand; ci += kMomCands) {
        Cand &cd = ctl.cand[ci];
        const int size = cd.size;
        const int *idx = B.cand_idx + fo + cd.idx_off;
        const int n_batches = (size + kMomBatch - 1) / kMomBatch;
        float accu = 0.0f;
        // producer state: two register stages (A: even batches, B: odd batches), indices two batches further ahead
        float xa[4], ya[4], za[4], xb[4], yb[4], zb[4];
        int ia[4], ib[4];
        const int par = warp - 1;   // producer: parity of its batches
        if (warp >= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m0 = par * kMomBatch + j * 32 + lane;        // its 1st batch
                const int m1 = m0 + 2 * kMomBatch;                     // its 2nd batch
                xa[j] = ya[j] = za[j] = xb[j] = yb[j] = zb[j] = 0.f;
                if (m0 < size) { const int p = idx[m0]; xa[j] = px[p]; ya[j] = py[p]; za[j] = pz[p]; }
                if (m1 < size) { const int p = idx[m1]; xb[j] = px[p]; yb[j] = py[p]; zb[j] = pz[p]; }
                ia[j] = (m0 + 4 * kMomBatch < size) ? idx[m0 + 4 * kMomBatch] : -1;
                ib[j] = (m1 + 4 * kMomBatch < size) ? idx[m1 + 4 * kMomBatch] : -1;
            }
        }
        auto produce = [&](int bt, float (&x)[4], float (&y)[4], float (&z)[4], int (&inext)[4]) {
            float *pr = s_prod[bt & 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float *q = pr + (j * 32 + lane) * 9;
                q[0] = x[j] * x[j]; q[1] = x[j] * y[j]; q[2] = x[j] * z[j]; q[3] = y[j] * y[j]; q[4] = y[j] * z[j]; q[5] = z[j] * z[j];
                q[6] = x[j]; q[7] = y[j]; q[8] = z[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // reload this stage for batch bt + 4 (this producer's batch after next); indices of bt + 8
                const int p = inext[j];
                x[j] = y[j] = z[j] = 0.f;
                if (p >= 0) { x[j] = px[p]; y[j] = py[p]; z[j] = pz[p]; }
                const int m8 = (bt + 8) * kMomBatch + j * 32 + lane;
                inext[j] = m8 < size ? idx[m8] : -1;
            }
        };
        // batch bt is produced while batch bt - 1 is consumed
        if (warp == 1) produce(0, xa, ya, za, ia);
        __syncthreads();
        for (int bt = 0; bt < n_batches; ++bt) {
            if (warp >= 1) {
                // batch bt + 1 belongs to the producer of its parity; that producer alternates its stages A, B
                const int nb = bt + 1;
                if (nb < n_batches && (nb & 1) == par) { if ((nb >> 1) & 1) produce(nb, xb, yb, zb, ib); else produce(nb, xa, ya, za, ia); }
            } else if (lane < 9) {
                const float *src = s_prod[bt & 1] + lane;
                const int cnt = min(kMomBatch, size - bt * kMomBatch);
                if (cnt == kMomBatch) {
#pragma unroll 32
                    for (int j = 0; j < kMomBatch; ++j) accu += src[j * 9];
                } else {
                    for (int j = 0; j < cnt; ++j) accu += src[j * 9];
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
            float acc[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = __shfl_sync(SPX_FULL, accu, k);
            if (lane == 0) {
                const float cnt = float(size);
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[k] /= cnt;
                float cov[9];
                cov[0] = acc[0] - acc[6] * acc[6];
                cov[1] = acc[1] - acc[6] * acc[7];
                cov[2] = acc[2] - acc[6] * acc[8];
                cov[4] = acc[3] - acc[7] * acc[7];
                cov[5] = acc[4] - acc[7] * acc[8];
                cov[8] = acc[5] - acc[8] * acc[8];
                cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
                float ev, vec[3];
                eigen33_smallest(cov, ev, vec);
                const float eig_sum = cov[0] + cov[4] + cov[8];
                float curvature;
                if (eig_sum != 0) curvature = fabsf(ev / eig_sum); else curvature = 0;
                cd.vec[0] = vec[0]; cd.vec[1] = vec[1]; cd.vec[2] = vec[2];
                cd.eigenvalue = ev;
                cd.centroid[0] = acc[6]; cd.centroid[1] = acc[7]; cd.centroid[2] = acc[8];
                cd.curvature = curvature;
#pragma unroll
                for (int k = 0; k < 9; ++k) cd.cov[k] = cov[k];
            }
        }
        __syncthreads();
    }
}

// K5b: one thread per frame replays segment()'s serial tail over the candidates in label order: the viewpoint
// vector `vp` that is never reset between clusters, the orientation flip, and the curvature acceptance.
__global__ void __launch_bounds__(128) k_models(Params P, Buffers B) {
    const int fl = blockIdx.x * blockDim.x + threadIdx.x;
    if (fl >= P.n_frames) return;
    const int f = P.frame0 + fl;
    FrameCtl &ctl = B.ctl[f];
    const size_t fo = size_t(f) * P.N;
    float vp[4] = {0.f, 0.f, 0.f, 0.f};
    int nm = 0;
    for (int ci = 0; ci < ctl.n_cand; ++ci) {
        const Cand &cd = ctl.cand[ci];
        const float cen[4] = {cd.centroid[0], cd.centroid[1], cd.centroid[2], 1.0f};
        float pp[4] = {cd.vec[0], cd.vec[1], cd.vec[2], 0.0f};
        pp[3] = -1 * dot4f(pp, cen);
#pragma unroll
        for (int k = 0; k < 4; ++k) vp[k] -= cen[k];
        const float cos_theta = dot4f(vp, pp);
        if (cos_theta < 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pp[k] *= -1;
            pp[3] = 0;
            pp[3] = -1 * dot4f(pp, cen);
        }
        B.root_model[fo + cd.root] = -1;   // held the candidate index until here (k_ccl_rank)
        if (double(cd.curvature) < 0.001) {
            if (nm < SPX_MAX_MODELS) {
                Model &m = ctl.models[nm];
#pragma unroll
                for (int k = 0; k < 4; ++k) m.coef[k] = pp[k];
                m.centroid[0] = cen[0]; m.centroid[1] = cen[1]; m.centroid[2] = cen[2];
                m.curvature = cd.curvature;
#pragma unroll
                for (int k = 0; k < 9; ++k) m.cov[k] = cd.cov[k];
                m.label = cd.label; m.root = cd.root; m.n0 = cd.size; m.n1 = 0; m.n2 = 0;
                m.cand_off = cd.idx_off;
                m.last_inlier = B.cand_idx[fo + cd.idx_off + cd.size - 1];
                m.contour_off = 0; m.n_contour = 0; m.plane = -1;
                B.root_model[fo + cd.root] = int16_t(nm);
                ++nm;
            } else {
                ctl.flags |= unsigned(SPX_FRAME_OVERFLOW);
            }
        }
    }
    ctl.n_models = nm;
    // were PCL's double-precision integral images provably free of rounding on this frame?  Every partial sum of channel ch is
    // an integer multiple of the finest unit in the last place u of that axis' coordinates and smaller than sat_sum[ch] < 2^e,
    // i.e. an integer below 2^e / u: a double holds it exactly when that is <= 2^53.  |x| >= min|n - cx| zmin / |fx| and
    // |y| >= min|m - cy| zmin / |fy| for every non-zero coordinate (two roundings: the factor 0.999).
    if (ctl.sat_zinv != 0u) {
        const float zmin = __uint_as_float(0x7f800000u - ctl.sat_zinv);
        const float lb[3] = {P.min_axf * zmin / fabsf(P.fx) * 0.999f, P.min_ayf * zmin / fabsf(P.fy) * 0.999f, zmin};
        for (int ch = 0; ch < 6; ++ch) {
            const float s = ctl.sat_sum[ch];
            if (s > 0.0f) {
                int e = 0, eu = 0;
                frexpf(s * 1.001f, &e);                       // s < 2^e
                frexpf(fmaxf(lb[ch % 3], 1.0e-45f), &eu);     // lb >= 2^(eu - 1): its unit in the last place is >= 2^(eu - 24)
                if (!(lb[ch % 3] > 0.0f) || e - (eu - 24) > 53) ctl.flags |= unsigned(SPX_FRAME_SAT_UNPROVEN);
            }
        }
    }
}

// K5c: plane id per pixel = model of its component (or -1)
__global__ void __launch_bounds__(256) k_pid_init(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    // -1: labelled but not part of a plane (refine may claim it); -2: a point PCL left unlabelled (non-finite)
    B.pid[fo + q] = int8_t(B.root_model[fo + B.parent[fo + q]]);   // k_ccl_rank left -2 at the roots of unlabelled points
}

// the same, four pixels per thread (N % 4 == 0): one 16-byte load of the forest, four gathers, one 4-byte store
__global__ void __launch_bounds__(256) k_pid_init4(Params P, Buffers B) {
    const int f = P.frame0 + blockIdx.y;
    const int q0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q0 >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    const int4 p4 = *reinterpret_cast<const int4 *>(B.parent + fo + q0);
    const int16_t *rm = B.root_model + fo;
    const unsigned a = uint8_t(int8_t(rm[p4.x])), b = uint8_t(int8_t(rm[p4.y])), c = uint8_t(int8_t(rm[p4.z])), d = uint8_t(int8_t(rm[p4.w]));
    *reinterpret_cast<unsigned *>(B.pid + fo + q0) = a | (b << 8) | (c << 16) | (d << 24);
}

}  // namespace spx
