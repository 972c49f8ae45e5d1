// spx_next.cu -- the callers / consumers either side of the plane-extraction path:
//   N4  pcl::VoxelGrid<PointXYZRGB> downsampling of plane clouds / contours (spx_voxel_grid, spx_voxel_downsample_results)
//   N1  Map::AssociatePlanesByBoundary + Map::PointDistanceFromPlane against a device-resident copy of the map planes'
//       boundary clouds (spx_map_*)
// Same build as spx_api.cu: sm_100a, -fmad=false (every product and sum rounded separately, as the reference's fp32 code).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/spx.h"
#include "../host/PlanePoseOptimizer.h"
#include "spx_internal.h"

namespace {

#define NX_CK(c, call)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) return spx_internal_fail((c), SPX_ERR_CUDA, #call, cudaGetErrorString(e_));      \
    } while (0)

inline int cdiv(long long a, int b) { return int((a + b - 1) / b); }

// ---- growable device scratch of the voxel-grid calls: one set per device, guarded by a mutex (the calls are synchronous) ----
constexpr int kMaxDevices = 64;
struct Scratch {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t ncap = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(&p, ncap);
        if (e == cudaSuccess) cap = ncap;
        return e;
    }
    ~Scratch() { /* freed with the process: contexts may already be gone */ }
};
struct Carver {
    char *base; size_t used = 0;
    explicit Carver(void *b) : base(static_cast<char *>(b)) {}
    template <typename T> T *take(size_t n) { T *q = reinterpret_cast<T *>(base + used); used += (n * sizeof(T) + 255) & ~size_t(255); return q; }
    template <typename T> static size_t bytes(size_t n) { return (n * sizeof(T) + 255) & ~size_t(255); }
};

// =================================================================================================================
// N4: pcl::VoxelGrid<pcl::PointXYZRGB>::applyFilter (PCL 1.8.0 filters/impl/voxel_grid.hpp) with SP-SLAM's settings
// (setLeafSize(l, l, l), defaults otherwise; src/MapDrawer.cc:91-92,115-116, src/PointCloudMapping.cc:117-118,172-173,
// dead path src/Frame.cc:810-814), for many independent clouds ("segments") in one pass.
//   1. per segment: min / max of the finite points, min_b = int(floor(min * inv_leaf)), div_b, the int-overflow test
//      (a segment that fails it is returned unchanged, as PCL does after its warning);
//   2. per point: idx = ijk . (1, div_b.x, div_b.x * div_b.y); 64-bit key = segment << 32 | idx;
//   3. stable radix sort of (key, point) -- PCL's output order IS ascending voxel index, so a hash table would still
//      need this sort over the occupied voxels, and at the reference's leaf (1 cm) against a point spacing of >= 1 cm
//      nearly every point is its own voxel;
//   4. heads of equal-key runs -> output slots (prefix sum); one thread per voxel adds its points in sorted order
//      (fp32 running sums of x, y, z and of r, g, b, a as floats, divided by float(n): CentroidPoint<PointXYZRGB>).
// PCL sorts with std::sort, which is not stable: inside a voxel its summation order is libstdc++'s introsort order.
// Here the order is the input order.  Voxels of <= 2 points are therefore bit-identical to PCL, larger ones agree to
// the last ulps of an n-term fp32 sum (tests state the bound).
// =================================================================================================================
struct VoxSeg {            // per segment
    long long start;       // first point in the source array
    int len;
    int gpos;              // first position in the gathered index space
    int min_b[3];
    int mul[3];
    int copy;              // leaf too small for int indices: output = input
};

__device__ __forceinline__ bool finite3(const spx_point &p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__global__ void __launch_bounds__(256) k_vox_bounds(const spx_point *__restrict__ pts, VoxSeg *segs, int n_seg, float ix, float iy, float iz) {
    __shared__ float s_mn[3][8], s_mx[3][8];
    const int s = blockIdx.x;
    if (s >= n_seg) return;
    VoxSeg &S = segs[s];
    const spx_point *p = pts + S.start;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < S.len; i += blockDim.x) {
        const spx_point q = p[i];
        if (!finite3(q)) continue;
        mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
        mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) { s_mn[k][wid] = mn[k]; s_mx[k][wid] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float inv[3] = {ix, iy, iz};
        long long d[3];
        int div_b[3];
        for (int k = 0; k < 3; ++k) {
            float a = s_mn[k][0], b = s_mx[k][0];
            for (int w = 1; w < int(blockDim.x >> 5); ++w) { a = fminf(a, s_mn[k][w]); b = fmaxf(b, s_mx[k][w]); }
            d[k] = (long long)((b - a) * inv[k]) + 1;
            S.min_b[k] = int(floorf(a * inv[k]));
            div_b[k] = int(floorf(b * inv[k])) - S.min_b[k] + 1;
        }
        S.copy = (S.len > 0 && d[0] * d[1] * d[2] > (long long)INT_MAX) ? 1 : 0;
        S.mul[0] = 1; S.mul[1] = div_b[0]; S.mul[2] = div_b[0] * div_b[1];
    }
}

__device__ __forceinline__ int seg_of(const VoxSeg *segs, int n_seg, int q) {
    int lo = 0, hi = n_seg - 1;          // last segment with gpos <= q (empty segments share a gpos: take the last)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].gpos <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// non-finite points get the key of a voxel past every real one of their segment and are dropped at the centroid stage
__global__ void __launch_bounds__(256) k_vox_keys(const spx_point *__restrict__ pts, const VoxSeg *__restrict__ segs, int n_seg, int total,
                                                  float ix, float iy, float iz, unsigned long long *keys, int *vals) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const int s = seg_of(segs, n_seg, q);
    const VoxSeg S = segs[s];
    const int j = q - S.gpos;
    const spx_point p = pts[S.start + j];
    unsigned idx;
    if (S.copy) idx = unsigned(j);
    else if (!finite3(p)) idx = 0xffffffffu;
    else {
        const int i0 = int(floorf(p.x * ix) - float(S.min_b[0]));
        const int i1 = int(floorf(p.y * iy) - float(S.min_b[1]));
        const int i2 = int(floorf(p.z * iz) - float(S.min_b[2]));
        idx = unsigned(i0 * S.mul[0] + i1 * S.mul[1] + i2 * S.mul[2]);
    }
    keys[q] = ((unsigned long long)unsigned(s) << 32) | idx;
    vals[q] = q;
}

__global__ void __launch_bounds__(256) k_vox_heads(const unsigned long long *__restrict__ keys, int total, int *head) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const unsigned long long k = keys[q];
    const bool dropped = unsigned(k) == 0xffffffffu;
    head[q] = (!dropped && (q == 0 || keys[q - 1] != k)) ? 1 : 0;
}

__global__ void __launch_bounds__(128) k_vox_centroids(const spx_point *__restrict__ pts, const VoxSeg *__restrict__ segs,
                                                       const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                                                       const int *__restrict__ head, const int *__restrict__ slot, int total,
                                                       spx_point *out, int *seg_count, int *out_idx) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total || !head[q]) return;
    const unsigned long long k = keys[q];
    const int s = int(k >> 32);
    const VoxSeg S = segs[s];
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
    int n = 0;
    for (int t = q; t < total && keys[t] == k; ++t) {
        const spx_point p = pts[S.start + (vals[t] - S.gpos)];
        sx += p.x; sy += p.y; sz += p.z;
        sr += float((p.rgba >> 16) & 255u); sg += float((p.rgba >> 8) & 255u); sb += float(p.rgba & 255u); sa += float(p.rgba >> 24);
        ++n;
    }
    spx_point o;
    if (S.copy) {
        o = pts[S.start + (vals[q] - S.gpos)];
    } else {
        const float fn = float(n);
        o.x = sx / fn; o.y = sy / fn; o.z = sz / fn;
        o.rgba = (uint32_t(sa / fn) << 24) | (uint32_t(sr / fn) << 16) | (uint32_t(sg / fn) << 8) | uint32_t(sb / fn);
    }
    const int dst = slot[q] - 1;       // inclusive prefix sum of the heads
    out[dst] = o;
    if (out_idx) out_idx[dst] = S.copy ? -1 : int(unsigned(k));
    atomicAdd(&seg_count[s], 1);
}

struct VoxPlan {           // device buffers of one run, carved from the scratch
    VoxSeg *segs; unsigned long long *keys, *keys2; int *vals, *vals2, *head, *slot, *seg_count, *seg_off; int *out_idx;
    spx_point *out; void *cub_tmp; size_t cub_bytes;
};

Scratch g_vox_scratch[kMaxDevices], g_vox_in[kMaxDevices];
std::mutex g_vox_mutex[kMaxDevices];

// core: `d_pts` device points, segments described on the host; leaves out / seg_count / seg_off(exclusive) on the device
int voxel_core(spx_ctx *c, cudaStream_t st, const spx_point *d_pts, std::vector<VoxSeg> &segs, int total, const float leaf[3], VoxPlan *plan,
               long long *n_out_total) {
    const int n_seg = int(segs.size());
    const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
    int seg_bits = 1;
    while ((1ll << seg_bits) < n_seg) ++seg_bits;
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int *)nullptr, (int *)nullptr,
                                    total, 0, 32 + seg_bits, st);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int *)nullptr, (int *)nullptr, total > n_seg + 1 ? total : n_seg + 1, st);
    const size_t cub_bytes = (sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 256;
    const size_t T = size_t(total > 0 ? total : 1), S = size_t(n_seg) + 1;
    size_t need = Carver::bytes<VoxSeg>(S) + 2 * Carver::bytes<unsigned long long>(T) + 5 * Carver::bytes<int>(T) + 2 * Carver::bytes<int>(S) +
                  Carver::bytes<spx_point>(T) + cub_bytes + 4096;
    Scratch &scratch = g_vox_scratch[spx_internal_device(c) % kMaxDevices];
    NX_CK(c, scratch.need(need));
    Carver A(scratch.p);
    VoxPlan &P = *plan;
    P.segs = A.take<VoxSeg>(S); P.keys = A.take<unsigned long long>(T); P.keys2 = A.take<unsigned long long>(T);
    P.vals = A.take<int>(T); P.vals2 = A.take<int>(T); P.head = A.take<int>(T); P.slot = A.take<int>(T); P.out_idx = A.take<int>(T);
    P.seg_count = A.take<int>(S); P.seg_off = A.take<int>(S); P.out = A.take<spx_point>(T);
    P.cub_tmp = A.base + A.used; P.cub_bytes = cub_bytes;
    *n_out_total = 0;
    if (total == 0 || n_seg == 0) {
        NX_CK(c, cudaMemsetAsync(P.seg_off, 0, S * sizeof(int), st));
        NX_CK(c, cudaMemsetAsync(P.seg_count, 0, S * sizeof(int), st));
        return SPX_OK;
    }
    NX_CK(c, cudaMemcpyAsync(P.segs, segs.data(), n_seg * sizeof(VoxSeg), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaMemsetAsync(P.seg_count, 0, S * sizeof(int), st));
    k_vox_bounds<<<n_seg, 256, 0, st>>>(d_pts, P.segs, n_seg, inv[0], inv[1], inv[2]);
    k_vox_keys<<<cdiv(total, 256), 256, 0, st>>>(d_pts, P.segs, n_seg, total, inv[0], inv[1], inv[2], P.keys, P.vals);
    size_t tb = P.cub_bytes;
    NX_CK(c, cub::DeviceRadixSort::SortPairs(P.cub_tmp, tb, P.keys, P.keys2, P.vals, P.vals2, total, 0, 32 + seg_bits, st));
    k_vox_heads<<<cdiv(total, 256), 256, 0, st>>>(P.keys2, total, P.head);
    tb = P.cub_bytes;
    NX_CK(c, cub::DeviceScan::InclusiveSum(P.cub_tmp, tb, P.head, P.slot, total, st));
    k_vox_centroids<<<cdiv(total, 128), 128, 0, st>>>(d_pts, P.segs, P.keys2, P.vals2, P.head, P.slot, total, P.out, P.seg_count, P.out_idx);
    tb = P.cub_bytes;
    NX_CK(c, cub::DeviceScan::ExclusiveSum(P.cub_tmp, tb, P.seg_count, P.seg_off, n_seg + 1, st));
    NX_CK(c, cudaGetLastError());
    int last = 0;
    NX_CK(c, cudaMemcpyAsync(&last, P.slot + (total - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    *n_out_total = last;
    return SPX_OK;
}

// plane records of the last extract -> segments (which = 0: mvPlanePoints, 1: mvBoundaryPoints), then the records are
// rewritten for the downsampled clouds, which replace the originals at the front of the arena
__global__ void __launch_bounds__(256) k_vox_patch_records(spx_plane *planes, int n_planes, const int *seg_count, const int *seg_off, int which) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_planes) return;
    if (which == 0) { planes[k].n_points = seg_count[k]; planes[k].points_off = seg_off[k]; }
    else { planes[k].n_boundary = seg_count[k]; planes[k].boundary_off = seg_off[k]; }
}

// =================================================================================================================
// N1: Map::AssociatePlanesByBoundary (src/Map.cc:196-283) for the planes of one frame.
// =================================================================================================================
__global__ void __launch_bounds__(256) k_assoc_dist(const float *__restrict__ plane_w, int n_planes, const float *__restrict__ map_w,
                                                    const spx_point *__restrict__ bnd, const long long *__restrict__ start,
                                                    const int *__restrict__ count, int n_map, float ang_th, float *angle_out, float *dist_out) {
    // one CTA per map plane: its boundary cloud against every frame plane that passes the angle test
    // (PointDistanceFromPlane, src/Map.cc:345-361: min over the cloud of |a x + b y + c z + d|, fp32, starting from 100)
    __shared__ float s_min[8];
    const int j = blockIdx.x;
    const float w0 = map_w[4 * j], w1 = map_w[4 * j + 1], w2 = map_w[4 * j + 2];
    const spx_point *b = bnd + start[j];
    const int n = count[j];
    for (int i = 0; i < n_planes; ++i) {
        const float a0 = plane_w[4 * i], a1 = plane_w[4 * i + 1], a2 = plane_w[4 * i + 2], a3 = plane_w[4 * i + 3];
        const float angle = a0 * w0 + a1 * w1 + a2 * w2;
        if (threadIdx.x == 0) angle_out[size_t(i) * n_map + j] = angle;
        if (!(angle > ang_th || angle < -ang_th)) continue;     // CTA uniform
        float m = 100.0f;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const spx_point p = b[k];
            const float e = a0 * p.x + a1 * p.y + a2 * p.z + a3;
            m = fminf(m, fabsf(e));                              // `dis < res` never takes a NaN, nor does fminf
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float r = s_min[0];
            for (int w = 1; w < int(blockDim.x >> 5); ++w) r = fminf(r, s_min[w]);
            dist_out[size_t(i) * n_map + j] = r;
        }
        __syncthreads();
    }
}

// MapPlane::MapPlane / MapPlane::UpdateBoundary (src/MapPlane.cc:25-31,144-147): pcl::transformPointCloud(cloud, *mvBoundaryPoints,
// T.inverse().matrix()), PCL 1.8.0 common/impl/transforms.hpp, dense branch with Scalar = double:
// x' = float(m00 x + m01 y + m02 z + m03) evaluated in double, left to right (no FMA: -fmad=false); other fields copied.
// The source is either a cloud the caller uploaded or, `from_result`, the boundary cloud of a plane of the last extract
// read where it lies in the device result arena (no host round trip).
struct MapXform { double m[12]; };
__global__ void __launch_bounds__(256) k_map_transform(const spx_point *__restrict__ src, const spx_frame_header *frames, const spx_plane *planes,
                                                       int frame, int plane, int n, MapXform X, spx_point *__restrict__ dst, int *err) {
    if (frames) {
        const spx_plane &R = planes[frames[frame].first_plane + plane];
        if (plane >= frames[frame].n_planes || R.n_boundary != n) { if (threadIdx.x == 0 && blockIdx.x == 0) *err = 1; return; }
        src += R.boundary_off;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const spx_point p = src[i];
    const double x = p.x, y = p.y, z = p.z;
    spx_point o = p;
    o.x = float(X.m[0] * x + X.m[1] * y + X.m[2] * z + X.m[3]);
    o.y = float(X.m[4] * x + X.m[5] * y + X.m[6] * z + X.m[7]);
    o.z = float(X.m[8] * x + X.m[9] * y + X.m[10] * z + X.m[11]);
    dst[i] = o;
}

// the reference's visiting loop over the map planes (order-dependent thresholds ldTh / lverTh / lparTh), one thread per frame plane
__global__ void __launch_bounds__(128) k_assoc_select(const float *__restrict__ angle_m, const float *__restrict__ dist_m, int n_planes, int n_seen,
                                                      int n_map, float dis_th, float ang_th, float ver_th, float par_th, int *assoc, int *vertical,
                                                      int *parallel, float *assoc_dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_planes) return;
    int a = -1, v = -1, p = -1;
    float ldTh = dis_th, lverTh = ver_th, lparTh = par_th;
    for (int j = 0; j < n_seen; ++j) {
        const float angle = angle_m[size_t(i) * n_map + j];
        if (angle > ang_th || angle < -ang_th) {
            const float dis = dist_m[size_t(i) * n_map + j];
            if (dis < ldTh) { ldTh = dis; a = j; continue; }
        }
        if (angle < lverTh && angle > -lverTh) { lverTh = fabsf(angle); v = j; continue; }
        if (angle > lparTh || angle < -lparTh) { lparTh = fabsf(angle); p = j; }
    }
    if (ldTh == dis_th) {
        for (int j = n_seen; j < n_map; ++j) {
            const float angle = angle_m[size_t(i) * n_map + j];
            if (angle > ang_th || angle < -ang_th) {
                const float dis = dist_m[size_t(i) * n_map + j];
                if (dis < ldTh) { ldTh = dis; a = j; }
            }
        }
    }
    assoc[i] = a; vertical[i] = v; parallel[i] = p; assoc_dist[i] = ldTh;
}

}  // namespace

struct spx_map {
    spx_ctx *ctx = nullptr;
    // map plane j: world coefficients d_w[4j..], boundary cloud d_bnd[start[j] .. start[j] + count[j]) with room for cap[j]
    // points (a cloud that outgrows its slot moves to the end of the arena; the arena is rebuilt when it runs out)
    float *d_w = nullptr; spx_point *d_bnd = nullptr; long long *d_start = nullptr; int *d_count = nullptr;
    std::vector<long long> start; std::vector<int> count, cap;
    size_t cap_w = 0, cap_bnd = 0, used_bnd = 0;
    int n_map = 0, n_seen = 0;
    spx_point *d_stage = nullptr; size_t cap_stage = 0;      // upload staging of spx_map_update_boundary
    int *d_err = nullptr;
    // per-call scratch (frame planes <= SPX_MAX_PLANES)
    float *d_plane_w = nullptr, *d_angle = nullptr, *d_dist = nullptr, *d_res_f = nullptr;
    int *d_res_i = nullptr;
    size_t cap_mat = 0;
    float *h_plane_w = nullptr, *h_res_f = nullptr; int *h_res_i = nullptr;   // pinned
};

extern "C" {

int spx_voxel_grid(spx_ctx *c, const spx_point *points, const int64_t *cloud_off, int n_clouds, const float leaf[3], spx_point *out,
                   int64_t *out_off) {
    if (!c) return SPX_ERR_ARG;
    if (!cloud_off || !leaf || !out_off || n_clouds < 0 || (n_clouds > 0 && !points && cloud_off[n_clouds] > 0))
        return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_grid", "null argument");
    if (!(leaf[0] > 0.f) || !(leaf[1] > 0.f) || !(leaf[2] > 0.f)) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_grid", "leaf size must be positive");
    const long long total = n_clouds ? cloud_off[n_clouds] - cloud_off[0] : 0;
    if (total > INT_MAX || n_clouds >= (1 << 24)) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_grid", "more than 2^31 points or 2^24 clouds in one call");
    if (total > 0 && !out) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_grid", "null output");
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    std::lock_guard<std::mutex> lock(g_vox_mutex[spx_internal_device(c) % kMaxDevices]);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    std::vector<VoxSeg> segs(static_cast<size_t>(n_clouds));
    for (int s = 0; s < n_clouds; ++s) {
        if (cloud_off[s + 1] < cloud_off[s]) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_grid", "offsets must not decrease");
        std::memset(&segs[s], 0, sizeof(VoxSeg));
        segs[s].start = cloud_off[s] - cloud_off[0]; segs[s].len = int(cloud_off[s + 1] - cloud_off[s]); segs[s].gpos = int(segs[s].start);
    }
    Scratch &in_scratch = g_vox_in[spx_internal_device(c) % kMaxDevices];
    NX_CK(c, in_scratch.need(size_t(total > 0 ? total : 1) * sizeof(spx_point)));
    spx_point *d_in = static_cast<spx_point *>(in_scratch.p);
    if (total) NX_CK(c, cudaMemcpyAsync(d_in, points + cloud_off[0], size_t(total) * sizeof(spx_point), cudaMemcpyHostToDevice, st));
    VoxPlan plan;
    long long n_out = 0;
    int rc = voxel_core(c, st, d_in, segs, int(total), leaf, &plan, &n_out);
    if (rc != SPX_OK) return rc;
    std::vector<int> off(static_cast<size_t>(n_clouds) + 1, 0);
    if (n_clouds) NX_CK(c, cudaMemcpyAsync(off.data(), plan.seg_off, (size_t(n_clouds) + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_out) NX_CK(c, cudaMemcpyAsync(out, plan.out, size_t(n_out) * sizeof(spx_point), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    for (int s = 0; s <= n_clouds; ++s) out_off[s] = off[size_t(s)];
    return SPX_OK;
}

int spx_voxel_downsample_results(spx_ctx *c, float leaf, int which) {
    if (!c || (which != 0 && which != 1)) return SPX_ERR_ARG;
    if (!(leaf > 0.f)) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_downsample_results", "leaf size must be positive");
    spx_device_result R;
    int rc = spx_get_device_results(c, &R);     // fails unless the last call was spx_extract_batch_device
    if (rc != SPX_OK) return rc;
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    std::lock_guard<std::mutex> lock(g_vox_mutex[spx_internal_device(c) % kMaxDevices]);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    long long tot[3];
    NX_CK(c, cudaMemcpyAsync(tot, R.totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    const int n_planes = int(tot[0]);
    const long long total = which == 0 ? tot[1] : tot[2];
    if (total > INT_MAX) return spx_internal_fail(c, SPX_ERR_ARG, "spx_voxel_downsample_results", "more than 2^31 points in the batch");
    if (n_planes == 0) return SPX_OK;
    std::vector<spx_plane> planes(static_cast<size_t>(n_planes));
    NX_CK(c, cudaMemcpyAsync(planes.data(), R.planes, planes.size() * sizeof(spx_plane), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    std::vector<VoxSeg> segs(planes.size());
    long long run = 0;
    for (size_t k = 0; k < planes.size(); ++k) {
        std::memset(&segs[k], 0, sizeof(VoxSeg));
        segs[k].start = which == 0 ? planes[k].points_off : planes[k].boundary_off;
        segs[k].len = which == 0 ? planes[k].n_points : planes[k].n_boundary;
        segs[k].gpos = int(run);
        run += segs[k].len;
    }
    spx_point *arena = const_cast<spx_point *>(which == 0 ? R.points : R.boundary);
    VoxPlan plan;
    long long n_out = 0;
    const float l3[3] = {leaf, leaf, leaf};
    rc = voxel_core(c, st, arena, segs, int(run), l3, &plan, &n_out);
    if (rc != SPX_OK) return rc;
    if (n_out) NX_CK(c, cudaMemcpyAsync(arena, plan.out, size_t(n_out) * sizeof(spx_point), cudaMemcpyDeviceToDevice, st));
    k_vox_patch_records<<<cdiv(n_planes, 256), 256, 0, st>>>(const_cast<spx_plane *>(R.planes), n_planes, plan.seg_count, plan.seg_off, which);
    NX_CK(c, cudaMemcpyAsync(const_cast<long long *>(R.totals) + (which == 0 ? 1 : 2), &n_out, sizeof(long long), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaStreamSynchronize(st));
    return SPX_OK;
}

// ---- N1 ----
int spx_map_create(spx_ctx *c, spx_map **out) {
    if (!c || !out) return SPX_ERR_ARG;
    *out = nullptr;
    spx_map *m = new (std::nothrow) spx_map();
    if (!m) return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_create", "out of host memory");
    m->ctx = c;
    DeviceGuard dev_guard_(spx_internal_device(c));
    cudaError_t e = dev_guard_.err;
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->h_plane_w), SPX_MAX_PLANES * 4 * sizeof(float), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->h_res_f), SPX_MAX_PLANES * sizeof(float), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&m->h_res_i), (SPX_MAX_PLANES * 3 + 1) * sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_plane_w), SPX_MAX_PLANES * 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_res_f), SPX_MAX_PLANES * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_res_i), SPX_MAX_PLANES * 3 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_err), sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(m->d_err, 0, sizeof(int));
    if (e != cudaSuccess) { spx_map_destroy(m); return spx_internal_fail(c, SPX_ERR_CUDA, "spx_map_create", cudaGetErrorString(e)); }
    *out = m;
    return SPX_OK;
}

void spx_map_destroy(spx_map *m) {
    if (!m) return;
    DeviceGuard dev_guard_(spx_internal_device(m->ctx));
    cudaFree(m->d_w); cudaFree(m->d_bnd); cudaFree(m->d_start); cudaFree(m->d_count); cudaFree(m->d_plane_w); cudaFree(m->d_angle);
    cudaFree(m->d_dist); cudaFree(m->d_res_f); cudaFree(m->d_res_i); cudaFree(m->d_stage); cudaFree(m->d_err);
    cudaFreeHost(m->h_plane_w); cudaFreeHost(m->h_res_f); cudaFreeHost(m->h_res_i);
    delete m;
}

}  // extern "C"

namespace {

// room for `need` more boundary points at the end of the arena: grow it (keeping the live clouds) when it is full
int map_reserve(spx_map *m, cudaStream_t st, size_t need) {
    spx_ctx *c = m->ctx;
    if (m->used_bnd + need <= m->cap_bnd) return SPX_OK;
    size_t live = 0;
    for (int j = 0; j < m->n_map; ++j) live += size_t(m->cap[size_t(j)]);
    const size_t ncap = (live + need) * 2 + 4096;
    spx_point *nb = nullptr;
    NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&nb), ncap * sizeof(spx_point)));
    size_t at = 0;
    for (int j = 0; j < m->n_map; ++j) {          // compaction: every plane keeps its slot size
        if (m->count[size_t(j)])
            NX_CK(c, cudaMemcpyAsync(nb + at, m->d_bnd + m->start[size_t(j)], size_t(m->count[size_t(j)]) * sizeof(spx_point), cudaMemcpyDeviceToDevice, st));
        m->start[size_t(j)] = (long long)at;
        at += size_t(m->cap[size_t(j)]);
    }
    NX_CK(c, cudaStreamSynchronize(st));
    cudaFree(m->d_bnd);
    m->d_bnd = nb; m->cap_bnd = ncap; m->used_bnd = at;
    if (m->n_map) NX_CK(c, cudaMemcpyAsync(m->d_start, m->start.data(), size_t(m->n_map) * sizeof(long long), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaStreamSynchronize(st));
    return SPX_OK;
}

// map plane j <- transform applied to a cloud (src on the device, or frames/planes/arena of the last extract)
int map_update(spx_map *m, int j, const double transform[16], const spx_point *d_src, const spx_frame_header *frames, const spx_plane *planes,
               int frame, int plane, int n) {
    spx_ctx *c = m->ctx;
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    const int old_count = m->count[size_t(j)];
    if (n > m->cap[size_t(j)]) {                  // the cloud outgrew its slot: a new one (1.5x) at the end of the arena
        const size_t slot = size_t(n) + size_t(n) / 2 + 64;
        int rc = map_reserve(m, st, slot);
        if (rc != SPX_OK) return rc;
        if (frames && old_count) {                // a refused update must leave the old cloud in place: it moves along
            NX_CK(c, cudaMemcpyAsync(m->d_bnd + m->used_bnd, m->d_bnd + m->start[size_t(j)], size_t(old_count) * sizeof(spx_point), cudaMemcpyDeviceToDevice, st));
        }
        m->start[size_t(j)] = (long long)m->used_bnd; m->cap[size_t(j)] = int(slot); m->used_bnd += slot;
    }
    m->count[size_t(j)] = n;
    MapXform X;
    for (int k = 0; k < 12; ++k) X.m[k] = transform[k];
    if (n) k_map_transform<<<cdiv(n, 256), 256, 0, st>>>(d_src, frames, planes, frame, plane, n, X, m->d_bnd + m->start[size_t(j)], m->d_err);
    NX_CK(c, cudaMemcpyAsync(m->d_start + j, &m->start[size_t(j)], sizeof(long long), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaMemcpyAsync(m->d_count + j, &m->count[size_t(j)], sizeof(int), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaGetLastError());
    if (frames) {                                 // the device checked the caller's n against the record
        int err = 0;
        NX_CK(c, cudaMemcpyAsync(&err, m->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
        NX_CK(c, cudaStreamSynchronize(st));
        if (err) {
            cudaMemsetAsync(m->d_err, 0, sizeof(int), st);
            m->count[size_t(j)] = old_count;      // nothing was written: the plane keeps its cloud
            cudaMemcpyAsync(m->d_count + j, &m->count[size_t(j)], sizeof(int), cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_update_boundary_from_result", "plane index / boundary size do not match the last extract");
        }
    } else {
        NX_CK(c, cudaStreamSynchronize(st));      // (the host arrays handed to the async copies above must stay put)
    }
    return SPX_OK;
}

}  // namespace

extern "C" {

int spx_map_upload(spx_map *m, const float *map_w, const spx_point *boundary, const int64_t *boundary_off, int n_seen, int n_map) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    if (n_map < 0 || n_seen < 0 || n_seen > n_map || (n_map > 0 && (!map_w || !boundary_off)))
        return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_upload", "bad argument");
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    const long long n_pts = n_map ? boundary_off[n_map] - boundary_off[0] : 0;
    if (n_pts > 0 && !boundary) return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_upload", "null boundary cloud");
    for (int j = 0; j < n_map; ++j)
        if (boundary_off[j + 1] < boundary_off[j] || boundary_off[j + 1] - boundary_off[j] > INT_MAX)
            return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_upload", "offsets must not decrease");
    NX_CK(c, cudaStreamSynchronize(st));
    if (size_t(n_map) > m->cap_w) {
        cudaFree(m->d_w); cudaFree(m->d_start); cudaFree(m->d_count); m->d_w = nullptr; m->d_start = nullptr; m->d_count = nullptr; m->cap_w = 0;
        const size_t cap = size_t(n_map) * 2 + 16;
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_w), cap * 4 * sizeof(float)));
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_start), cap * sizeof(long long)));
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_count), cap * sizeof(int)));
        m->cap_w = cap;
    }
    if (size_t(n_map) * SPX_MAX_PLANES > m->cap_mat) {
        cudaFree(m->d_angle); cudaFree(m->d_dist); m->d_angle = m->d_dist = nullptr; m->cap_mat = 0;
        const size_t cap = (size_t(n_map) * 2 + 16) * SPX_MAX_PLANES;
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_angle), cap * sizeof(float)));
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_dist), cap * sizeof(float)));
        m->cap_mat = cap;
    }
    // every plane gets a slot of 1.5x its cloud (+64), so that a boundary that grows a little is updated in place
    m->start.assign(size_t(n_map), 0); m->count.assign(size_t(n_map), 0); m->cap.assign(size_t(n_map), 0);
    size_t at = 0;
    for (int j = 0; j < n_map; ++j) {
        const int n = int(boundary_off[j + 1] - boundary_off[j]);
        m->start[size_t(j)] = (long long)at; m->count[size_t(j)] = n; m->cap[size_t(j)] = n + n / 2 + 64;
        at += size_t(m->cap[size_t(j)]);
    }
    if (at > m->cap_bnd) {
        cudaFree(m->d_bnd); m->d_bnd = nullptr; m->cap_bnd = 0;
        const size_t cap = at * 2 + 4096;
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_bnd), cap * sizeof(spx_point)));
        m->cap_bnd = cap;
    }
    m->used_bnd = at;
    m->n_map = n_map; m->n_seen = n_seen;
    if (n_map) {
        NX_CK(c, cudaMemcpyAsync(m->d_w, map_w, size_t(n_map) * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
        NX_CK(c, cudaMemcpyAsync(m->d_start, m->start.data(), size_t(n_map) * sizeof(long long), cudaMemcpyHostToDevice, st));
        NX_CK(c, cudaMemcpyAsync(m->d_count, m->count.data(), size_t(n_map) * sizeof(int), cudaMemcpyHostToDevice, st));
        for (int j = 0; j < n_map; ++j)
            if (m->count[size_t(j)])
                NX_CK(c, cudaMemcpyAsync(m->d_bnd + m->start[size_t(j)], boundary + boundary_off[j], size_t(m->count[size_t(j)]) * sizeof(spx_point),
                                         cudaMemcpyHostToDevice, st));
        NX_CK(c, cudaStreamSynchronize(st));
    }
    return SPX_OK;
}

int spx_map_set_world_pos(spx_map *m, int j, const float coef_w[4]) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    if (j < 0 || j >= m->n_map || !coef_w) return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_set_world_pos", "bad argument");
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    NX_CK(c, cudaMemcpyAsync(m->d_w + 4 * j, coef_w, 4 * sizeof(float), cudaMemcpyHostToDevice, st));
    NX_CK(c, cudaStreamSynchronize(st));
    return SPX_OK;
}

int spx_map_update_boundary(spx_map *m, int j, const double transform[16], const spx_point *cloud, int n) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    if (j < 0 || j >= m->n_map || !transform || n < 0 || (n > 0 && !cloud)) return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_update_boundary", "bad argument");
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    if (size_t(n) > m->cap_stage) {
        NX_CK(c, cudaStreamSynchronize(st));
        cudaFree(m->d_stage); m->d_stage = nullptr; m->cap_stage = 0;
        const size_t cap = size_t(n) * 2 + 1024;
        NX_CK(c, cudaMalloc(reinterpret_cast<void **>(&m->d_stage), cap * sizeof(spx_point)));
        m->cap_stage = cap;
    }
    if (n) NX_CK(c, cudaMemcpyAsync(m->d_stage, cloud, size_t(n) * sizeof(spx_point), cudaMemcpyHostToDevice, st));
    return map_update(m, j, transform, m->d_stage, nullptr, nullptr, 0, 0, n);
}

int spx_map_update_boundary_from_result(spx_map *m, int j, const double transform[16], int frame, int plane, int n_boundary) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    const spx_frame_header *frames = nullptr; const spx_plane *planes = nullptr; const spx_point *bnd = nullptr;
    const int n_frames = spx_internal_last_results(c, &frames, &planes, &bnd);
    if (j < 0 || j >= m->n_map || !transform || n_boundary < 0 || plane < 0 || frame < 0 || frame >= n_frames)
        return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_update_boundary_from_result", "bad argument (or no extract yet)");
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    return map_update(m, j, transform, bnd, frames, planes, frame, plane, n_boundary);
}

int spx_map_get_boundary(spx_map *m, int j, spx_point *out, int cap, int *n) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    if (j < 0 || j >= m->n_map || !n) return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_get_boundary", "bad argument");
    *n = m->count[size_t(j)];
    if (!out) return SPX_OK;
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    const int k = *n < cap ? *n : cap;
    if (k > 0) NX_CK(c, cudaMemcpyAsync(out, m->d_bnd + m->start[size_t(j)], size_t(k) * sizeof(spx_point), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    return SPX_OK;
}

int spx_map_associate(spx_map *m, const float *plane_w, int n_planes, float dis_th, float ang_th, float ver_th, float par_th, int32_t *assoc,
                      int32_t *vertical, int32_t *parallel, float *assoc_dist) {
    if (!m) return SPX_ERR_ARG;
    spx_ctx *c = m->ctx;
    if (n_planes < 0 || n_planes > SPX_MAX_PLANES || (n_planes > 0 && (!plane_w || !assoc || !vertical || !parallel)))
        return spx_internal_fail(c, SPX_ERR_ARG, "spx_map_associate", "bad argument");
    if (n_planes == 0) return SPX_OK;
    if (m->n_map == 0) {
        for (int i = 0; i < n_planes; ++i) { assoc[i] = vertical[i] = parallel[i] = -1; if (assoc_dist) assoc_dist[i] = dis_th; }
        return SPX_OK;
    }
    DeviceGuard dev_guard_(spx_internal_device(c)); NX_CK(c, dev_guard_.err);
    cudaStream_t st = static_cast<cudaStream_t>(spx_internal_stream(c));
    std::memcpy(m->h_plane_w, plane_w, size_t(n_planes) * 4 * sizeof(float));
    NX_CK(c, cudaMemcpyAsync(m->d_plane_w, m->h_plane_w, size_t(n_planes) * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
    k_assoc_dist<<<m->n_map, 256, 0, st>>>(m->d_plane_w, n_planes, m->d_w, m->d_bnd, m->d_start, m->d_count, m->n_map, ang_th, m->d_angle, m->d_dist);
    k_assoc_select<<<cdiv(n_planes, 128), 128, 0, st>>>(m->d_angle, m->d_dist, n_planes, m->n_seen, m->n_map, dis_th, ang_th, ver_th, par_th,
                                                       m->d_res_i, m->d_res_i + SPX_MAX_PLANES, m->d_res_i + 2 * SPX_MAX_PLANES, m->d_res_f);
    NX_CK(c, cudaMemcpyAsync(m->h_res_i, m->d_res_i, 3 * SPX_MAX_PLANES * sizeof(int), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaMemcpyAsync(m->h_res_f, m->d_res_f, SPX_MAX_PLANES * sizeof(float), cudaMemcpyDeviceToHost, st));
    NX_CK(c, cudaStreamSynchronize(st));
    NX_CK(c, cudaGetLastError());
    for (int i = 0; i < n_planes; ++i) {
        assoc[i] = m->h_res_i[i]; vertical[i] = m->h_res_i[SPX_MAX_PLANES + i]; parallel[i] = m->h_res_i[2 * SPX_MAX_PLANES + i];
        if (assoc_dist) assoc_dist[i] = m->h_res_f[i];
    }
    return SPX_OK;
}

// ---- N3: host-side pose-only optimisation with plane edges (sp_slam_b200/host/PlanePoseOptimizer.h) ----
int spx_pose_optimize_planes(double Tcw[16], const spx_plane_edge *edges, int n_edges, int rounds, int iterations, uint8_t *outlier,
                             double *chi2, int *n_bad) {
    if (!Tcw || n_edges < 0 || (n_edges > 0 && !edges) || rounds < 1 || iterations < 1) return SPX_ERR_ARG;
    for (int k = 0; k < 16; ++k) if (!std::isfinite(Tcw[k])) return SPX_ERR_ARG;
    spx_host::PlanePoseOptimizer opt;
    opt.edges.resize(size_t(n_edges));
    for (int i = 0; i < n_edges; ++i) {
        if (edges[i].kind < 0 || edges[i].kind > 2) return SPX_ERR_ARG;
        for (const float *v : {edges[i].plane_w, edges[i].measurement}) {     // a plane needs a finite, non-zero normal
            const float n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            if (!std::isfinite(n2) || !std::isfinite(v[3]) || n2 == 0.0f) return SPX_ERR_ARG;
        }
        spx_host::PlaneEdge &e = opt.edges[size_t(i)];
        e.kind = edges[i].kind;
        e.world = spx_host::Plane3D::from_coefficients(edges[i].plane_w);
        e.measurement = spx_host::Plane3D::from_coefficients(edges[i].measurement);
        for (int k = 0; k < 3; ++k) e.info[k] = edges[i].info[k];
        e.huber_delta = edges[i].huber_delta;
        e.chi2_max = edges[i].chi2_max;
    }
    const int bad = opt.PoseOptimization(Tcw, rounds, iterations);
    for (int i = 0; i < n_edges; ++i) {
        if (outlier) outlier[i] = opt.edges[size_t(i)].outlier ? 1 : 0;
        if (chi2) chi2[i] = opt.edges[size_t(i)].chi2();
    }
    if (n_bad) *n_bad = bad;
    return SPX_OK;
}

int spx_plane_edge_errors(const double Tcw[16], const spx_plane_edge *edges, int n_edges, double *errors) {
    if (!Tcw || n_edges < 0 || (n_edges > 0 && (!edges || !errors))) return SPX_ERR_ARG;
    const spx_host::Pose pose = spx_host::Pose::from_matrix(Tcw);
    for (int i = 0; i < n_edges; ++i) {
        if (edges[i].kind < 0 || edges[i].kind > 2) return SPX_ERR_ARG;
        spx_host::PlaneEdge e;
        e.kind = edges[i].kind;
        e.world = spx_host::Plane3D::from_coefficients(edges[i].plane_w);
        e.measurement = spx_host::Plane3D::from_coefficients(edges[i].measurement);
        e.compute_error(pose);
        for (int k = 0; k < 3; ++k) errors[3 * i + k] = k < e.dim() ? e.error[k] : 0.0;
    }
    return SPX_OK;
}

}  // extern "C"
