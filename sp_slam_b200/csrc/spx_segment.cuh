// spx_segment.cuh -- K4..K5: comparator connected components (union-find with atomics), PCL label ranks,
// per-component fp32 moments in the reference's accumulation order, closed-form plane fit, curvature filter.
//
// Reference: /root/reference/src/Frame.cc:898-905 (OrganizedMultiPlaneSegmentation settings) and PCL 1.8.0
// segmentation/impl/organized_multi_plane_segmentation.hpp (segment), segmentation/plane_coefficient_comparator.h
// (compare), segmentation/impl/organized_connected_component_segmentation.hpp (segment),
// common/impl/centroid.hpp (computeMeanAndCovarianceMatrix), common/impl/eigen.hpp (eigen33).
#pragma once
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
__device__ __forceinline__ void uf_unite(int *parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
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
    const int f = blockIdx.z, lane = threadIdx.x;
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
        B.cnt[fo + q] = 0;
    }
    const unsigned linked = __ballot_sync(SPX_FULL, valid && L);
    const unsigned starts = ~linked | 1u;                       // lane 0 always starts a run inside the segment
    const int s = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
    if (valid) B.parent[fo + q] = r * w + blockIdx.x * 32 + s;
}

// K4b: merges across the 32-column segment seams and along vertical edges.  A vertical union is skipped when the
// two pixels are already connected through their left neighbours (L(q) & L(up) & U(left)).
__global__ void __launch_bounds__(256) k_ccl_merge(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const int w = P.w;
    const size_t fo = size_t(f) * P.N;
    const int r = q / w, c = q - r * w;
    const unsigned cb = B.conn[fo + q];
    int *parent = B.parent + fo;
    if ((cb & 1u) && (c & 31) == 0) uf_unite(parent, q, q - 1);
    if (cb & 2u) {
        const bool skip = c > 0 && (cb & 1u) && (B.conn[fo + q - w] & 1u) && (B.conn[fo + q - 1] & 2u);
        if (!skip) uf_unite(parent, q, q - w);
    }
}

// K4c: path compression to the root + component sizes (warp-aggregated atomics on the root's counter)
__global__ void __launch_bounds__(256) k_ccl_flatten(Params P, Buffers B) {
    const int f = blockIdx.y;
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

// K4d: one CTA per frame.  Exclusive prefix count of roots in raster order = PCL's dense label of each component;
// components with size > Plane.MinSize become plane candidates, in label order.
__global__ void __launch_bounds__(1024) k_ccl_rank(Params P, Buffers B) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned running_s, block_tot;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t fo = size_t(f) * P.N;
    FrameCtl &ctl = B.ctl[f];
    if (tid == 0) running_s = 0;
    __syncthreads();
    for (int base = 0; base < P.N; base += 1024) {
        const int q = base + tid;
        bool isroot = false, iscand = false;
        int sz = 0;
        if (q < P.N) {
            isroot = B.parent[fo + q] == q;
            if (isroot) { sz = B.cnt[fo + q]; iscand = unsigned(sz) > unsigned(P.min_size); }
        }
        // roots are counted in bits 0..19 (N < 2^20), candidates in bits 20..31
        const unsigned v = (isroot ? 1u : 0u) | (iscand ? (1u << 20) : 0u);
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(SPX_FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned t = warp_tot[lane];
            unsigned s = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(SPX_FULL, s, o);
                if (lane >= o) s += u;
            }
            warp_tot[lane] = s - t;   // exclusive prefix of the warp totals
            if (lane == 31) block_tot = s;
        }
        __syncthreads();
        const unsigned excl = running_s + warp_tot[wid] + incl - v;
        if (isroot) {
            B.lab[fo + q] = int(excl & 0xFFFFFu);
            B.root_model[fo + q] = -1;
        }
        if (iscand) {
            const unsigned k = excl >> 20;
            if (k < SPX_MAX_CAND) {
                ctl.cand[k].root = q; ctl.cand[k].label = int(excl & 0xFFFFFu); ctl.cand[k].size = sz;
            } else {
                atomicOr(&ctl.flags, unsigned(SPX_FRAME_OVERFLOW));
            }
        }
        __syncthreads();
        if (tid == 0) running_s += block_tot;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned tot = running_s;
        ctl.n_labels = int(tot & 0xFFFFFu) + 1;
        int nc = int(tot >> 20);
        if (nc > SPX_MAX_CAND) nc = SPX_MAX_CAND;
        ctl.n_cand = nc;
        int off = 0;
        for (int k = 0; k < nc; ++k) { ctl.cand[k].idx_off = off; off += ctl.cand[k].size; }
    }
}

// K4e: PCL label per pixel (labels.points[i].label before refine); only needed by the parity taps
__global__ void __launch_bounds__(256) k_ccl_label(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    const int root = B.parent[fo + q];
    if (root != q) B.lab[fo + q] = B.lab[fo + root];
}

// ---------------------------------------------------------------------------------------------------------------
// K5a: one warp per candidate.  computeMeanAndCovarianceMatrix accumulates nine fp32 moments over the component's
// pixels IN RASTER ORDER, one rounding per add; a tree reduction would round differently, so the order is kept:
// the warp walks the frame in 32-pixel chunks, ballots the members, and lanes 0..8 each carry one accumulator
// through the members in order (values broadcast by shuffle).  The raster-ordered index list (label_indices[l])
// and each pixel's position in it fall out of the same walk.  Lane 0 then solves the 3x3 eigenproblem.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_moments_fit(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * 4 + warp;
    FrameCtl &ctl = B.ctl[f];
    if (ci >= ctl.n_cand) return;
    Cand &cd = ctl.cand[ci];
    const int root = cd.root, size = cd.size, off = cd.idx_off;
    const size_t fo = size_t(f) * P.N;
    const int *parent = B.parent + fo;
    const float *px = B.px + fo, *py = B.py + fo, *pz = B.pz + fo;
    int *cand_idx = B.cand_idx + fo;
    int *pos = B.pos + fo;
    // factor selectors of accu[lane]: xx xy xz yy yz zz x y z
    const int sa = (lane <= 2 || lane == 6) ? 0 : ((lane == 3 || lane == 4 || lane == 7) ? 1 : 2);
    const int sb = (lane == 0) ? 0 : ((lane == 1 || lane == 3) ? 1 : ((lane == 2 || lane == 4 || lane == 5) ? 2 : 3));
    float accu = 0.0f;
    int found = 0;
    for (int base = (root >> 5) << 5; found < size && base < P.N; base += 32) {
        const int p = base + lane;
        const bool m = p < P.N && parent[p] == root;
        unsigned bits = __ballot_sync(SPX_FULL, m);
        if (bits == 0u) continue;
        float x = 0.f, y = 0.f, z = 0.f;
        if (m) {
            x = px[p]; y = py[p]; z = pz[p];
            const int k = found + __popc(bits & ((1u << lane) - 1u));
            cand_idx[off + k] = p;
            pos[p] = k;
        }
        found += __popc(bits);
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1u;
            const float xj = __shfl_sync(SPX_FULL, x, j), yj = __shfl_sync(SPX_FULL, y, j), zj = __shfl_sync(SPX_FULL, z, j);
            const float a = sa == 0 ? xj : (sa == 1 ? yj : zj);
            const float b = sb == 0 ? xj : (sb == 1 ? yj : (sb == 2 ? zj : 1.0f));
            accu += a * b;
        }
    }
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

// K5b: one thread per frame replays segment()'s serial tail over the candidates in label order: the viewpoint
// vector `vp` that is never reset between clusters, the orientation flip, and the curvature acceptance.
__global__ void __launch_bounds__(128) k_models(Params P, Buffers B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.n_frames) return;
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
}

// K5c: plane id per pixel = model of its component (or -1)
__global__ void __launch_bounds__(256) k_pid_init(Params P, Buffers B) {
    const int f = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.N) return;
    const size_t fo = size_t(f) * P.N;
    B.pid[fo + q] = int8_t(B.root_model[fo + B.parent[fo + q]]);
}

}  // namespace spx
