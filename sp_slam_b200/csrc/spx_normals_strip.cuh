// spx_normals_strip.cuh -- K3 as a strip kernel: cloud + integral-image normals + plane_d + comparator links for a
// 32-column strip of a frame, streamed top to bottom, the depth rows staged by TMA.
//
// Same reference stages and the same outputs as k_normals_link (spx_normals.cuh): /root/reference/src/Frame.cc:855-885 and
// PCL 1.8.0 features/impl/integral_image_normal.hpp (initAverage3DGradientMethod, computeFeatureFull, computePointNormal),
// features/impl/integral_image2D.hpp, segmentation/plane_coefficient_comparator.h (compare).
//
// Why a strip: a 32x16 tile drags a 7-pixel halo on all four sides (44 x 28 cloud points for 32 x 16 outputs, 2.4x); a strip
// that runs the full height of the frame has no vertical halo at all (44 columns for 32, 1.375x).  The fp64 integral images
// of the six gradient channels are built incrementally: every (column, channel) thread keeps the running COLUMN sum of its
// differences in a register as the rows stream by, and a node row of the integral image is the prefix sum of those column
// sums along the row; only the 18 node rows the k x k windows (k <= 10) of the current batch of 8 rows can touch live in
// shared memory, in a ring.  The fp64 sums of fp32 differences are exact for depth data (DESIGN.md "Exactness devices";
// the per-frame bound is evaluated here and a frame that cannot be proven exact is flagged SPX_FRAME_SAT_UNPROVEN), so the
// summation order does not matter and the window sums equal PCL's whole-image ones bit for bit.
//
// Per batch of 8 rows (r0 = 8k), four phases separated by CTA barriers:
//   A  links of batch k-1 (8 warps: one per row) || back-projection of cloud rows [r0+5, r0+13) from the staged depth chunk
//   B  column sums: differences of rows [r0+4, r0+12) -> node rows [r0+5, r0+13) hold the running column sums
//   C  row prefix of those node rows (one thread per (row, channel): 48 serial chains packed into two warps)
//   D  window sums + normal + flip + plane_d of rows [r0, r0+8) (one thread per pixel of the 8 x 33 block)
// Depth chunks (8 sampled rows x the 44 * Cloud.Dis image columns the strip's samples span) arrive through
// cp.async.bulk.tensor -- the sampled rows are the tensor's dimension 1, so only they are touched; the hardware zero-fills
// outside the image -- one chunk ahead of their use, signalled on an mbarrier.  Frames with NaN / Inf depth (flagged by k_edge_chamfer) are left to k_normals_link.
#pragma once
#include <cuda.h>

#include "spx_normals.cuh"

namespace spx {

constexpr int kStW = 32;                  // output columns of a strip
constexpr int kSB = 8;                   // rows per batch
constexpr int kSCW = kStW + 12;           // 44 cloud columns: image columns [tc-7, tc+36]
constexpr int kSDW = kStW + 10;           // 42 difference columns: [tc-6, tc+35]
constexpr int kSNW = kStW + 1;            // 33 columns that get a normal: [tc-1, tc+31]
constexpr int kSSW = kSDW + 1;           // 43 integral-image nodes per row
constexpr int kSRing = 18;               // node rows in flight: [r0-5, r0+12]
constexpr int kSCRing = 14;              // cloud rows in flight: 13 needed
constexpr int kSNRing = 9;               // rows of normals the link step looks at: the batch and the row above it
constexpr int kSThreads = 288;           // 9 warps: 8 x 33 = 264 normals in one round, 6 x 42 = 252 column sums

struct StripSmem {
    double sat[kSRing][6][kSSW];         // 37 152 B
    float cloud[kSCRing][3][kSCW];       //  7 392 B   (sat + cloud = 348 x 128 B)
    float nrm[kSNRing][5][kSNW];         //  5 940 B: nx ny nz plane_d z
    unsigned long long bar;
    float red_sum[6];
    int red_exp[3];
};
// the TMA destination follows the fixed part, 128-byte aligned: one depth chunk of 8 sampled rows x (44 * Cloud.Dis) columns.
// TMA has no element stride in dimension 0 (tools/tma_probe.cu: the instruction faults), so the row segments arrive whole and
// the back-projection reads every Cloud.Dis-th float (a stride of 3 words is conflict free).  Four CTAs fit an SM.
constexpr size_t kStripStageOff = (sizeof(StripSmem) + 127) / 128 * 128;
inline size_t strip_smem_bytes(int dis, bool tma) { return kStripStageOff + (tma ? size_t(kSB) * kSCW * dis * 4 : 0); }
inline bool strip_tma_ok(int dis) { return kSCW * dis <= 256 && (kSCW * dis) % 4 == 0; }

__device__ __forceinline__ unsigned strip_smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }

// a / b for the constant b (fx, fy): q = RN(a * RN(1/b)), one residual correction with fused multiply-adds.  Whether this
// equals the IEEE quotient for EVERY significand of a is verified exhaustively for the context's two constants when the
// context is created (k_check_div below); otherwise, and outside the range where no intermediate can be subnormal or
// overflow, the true division is used.
__device__ __forceinline__ float div_const(float a, float b, float rb, int fast) {
    const float aa = fabsf(a);
    if (fast && aa > 1.0e-30f && aa < 1.0e30f) {
        const float q = __fmul_rn(a, rb);
        const float r = __fmaf_rn(-q, b, a);
        return __fmaf_rn(r, rb, q);
    }
    return a / b;
}

// exhaustive check of div_const for one constant: all 2^23 significands of a (the quotient of scaled operands scales exactly)
__global__ void __launch_bounds__(256) k_check_div(float b, float rb, int *mismatch) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << 23)) return;
    const float a = __uint_as_float(0x3f800000u | i);
    if (div_const(a, b, rb, 1) != a / b || div_const(-a, b, rb, 1) != -a / b) atomicAdd(mismatch, 1);
}

template <bool kTMA>
__global__ void __launch_bounds__(kSThreads, 4)
k_normals_strip(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ depth, Params P, Buffers B, int write_normals) {
    extern __shared__ __align__(128) unsigned char strip_raw[];
    StripSmem &S = *reinterpret_cast<StripSmem *>(strip_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int f = P.frame0 + blockIdx.y;
    const int tc = blockIdx.x * kStW;
    const int w = P.w, h = P.h;
    FrameCtl &ctl = B.ctl[f];
    if (ctl.flags & unsigned(SPX_FRAME_NONFINITE)) return;      // NaN / Inf depth: the frame goes through k_normals_link (finite-count images)
    const size_t fo = size_t(f) * P.N;
    const char *img = reinterpret_cast<const char *>(depth) + size_t(f) * P.samp_fstride;
    const int K = (h + kSB - 1) / kSB;                          // batches

    float *const stage = reinterpret_cast<float *>(strip_raw + kStripStageOff);
    const int bw = kSCW * P.dis;                                // floats per staged row
    if (kTMA) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(strip_smem_u32(&S.bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    // node row 0 of the integral images is zero
    for (int i = tid; i < 6 * kSSW; i += kSThreads) S.sat[0][i / kSSW][i % kSSW] = 0.0;
    __syncthreads();
    // chunk m = sampled rows [8m - 3, 8m + 5), image columns [(tc - 7) dis, (tc + 37) dis); rows / columns outside the image
    // arrive as zeros.  One stage: chunk m + 1 is requested as soon as phase A has consumed chunk m and lands during phases B..D.
    auto issue_chunk = [&](int m) {
        const unsigned bar = strip_smem_u32(&S.bar), dst = strip_smem_u32(stage);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kSB * bw * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&tmap), "r"((tc - 7) * P.dis), "r"(kSB * m - 3), "r"(f), "r"(bar) : "memory");
    };
    if (kTMA && tid == 0) issue_chunk(0);

    // ---- per-thread constants ----
    // back-projection: items tid and tid + 288 of the 8 x 44 chunk
    int bp_rr[2], bp_lx[2];
    float bp_xf[2];
    bool bp_cin[2], bp_own[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * kSThreads;
        bp_rr[j] = i / kSCW; bp_lx[j] = i - bp_rr[j] * kSCW;
        const int c = tc - 7 + bp_lx[j];
        bp_cin[j] = i < kSB * kSCW && c >= 0 && c < w;
        bp_own[j] = bp_cin[j] && bp_lx[j] >= 7 && bp_lx[j] < 7 + kStW;
        bp_xf[j] = float(c * P.dis) - P.cx;
    }
    // column sums: thread = (channel, difference column); channels 0..2 = d/dx of x y z, 3..5 = d/dy
    const int cs_ch = tid / kSDW, cs_dxl = tid - cs_ch * kSDW;
    const bool cs_on = tid < 6 * kSDW;
    const int cs_c = tc - 6 + cs_dxl;
    const bool cs_colok = cs_c >= 1 && cs_c <= w - 2;
    double colsum = 0.0;
    // exactness bound: every partial sum of channel ch is a multiple of the finest unit in the last place of that axis'
    // coordinates and smaller than the channel's sum of |differences| (own columns only, so the strips of a frame add up)
    const bool cs_owncol = cs_dxl >= 6 && cs_dxl < 6 + kStW;
    float sabs = 0.0f;
    int negx = 0, negy = 0, negz = 0;
    // window pixel of phase D: (row i, column nx_) of the 8 x 33 block
    const int wn_i = tid / kSNW, wn_x = tid - wn_i * kSNW;
    const bool wn_on = tid < kSB * kSNW;
    const int wn_c = tc - 1 + wn_x;
    const float qnan = __int_as_float(0x7fc00000);
    const uint8_t *kwin = B.kwin + fo;
    constexpr int border = 10;
    auto win_k = [&](int r) -> int {     // window size of the thread's pixel in row r (0: none); the point's z is tested later
        if (wn_on && r < h && r >= border && r < h - border && wn_c >= border && wn_c < w - border) return int(kwin[r * w + wn_c]);
        return 0;
    };
    int k_next = win_k(wn_i);            // prefetched one batch ahead

    for (int k = -1; k <= K; ++k) {
        const int r0 = kSB * k;
        // ================= phase A: links of batch k - 1, back-projection of rows [r0 + 5, r0 + 13) =================
        if (k >= 1 && wid < 8) {
            const int r = r0 - kSB + wid, c = tc + lane;
            if (r < h) {                                                  // warp uniform
                const bool valid = c < w;
                bool L = false, U = false;
                const int q = r * w + c;
                if (valid) {
                    const int sr = r % kSNRing;
                    const float n1x = S.nrm[sr][0][lane + 1], n1y = S.nrm[sr][1][lane + 1], n1z = S.nrm[sr][2][lane + 1];
                    const float d1 = S.nrm[sr][3][lane + 1], Zv = S.nrm[sr][4][lane + 1];
                    const float z = Zv;                                   // vec.dot(z_axis_): x * 0 + (y * 0 + z * 1) = z for finite x, y
                    float threshold = P.dist_thr;
                    threshold *= z * z;
                    if (c >= 1)
                        L = (fabsf(d1 - S.nrm[sr][3][lane]) < threshold) &&
                            (dot3f(n1x, n1y, n1z, S.nrm[sr][0][lane], S.nrm[sr][1][lane], S.nrm[sr][2][lane]) > P.ang_cos);
                    if (r >= 1) {
                        const int su = (r - 1) % kSNRing;
                        U = (fabsf(d1 - S.nrm[su][3][lane + 1]) < threshold) &&
                            (dot3f(n1x, n1y, n1z, S.nrm[su][0][lane + 1], S.nrm[su][1][lane + 1], S.nrm[su][2][lane + 1]) > P.ang_cos);
                    }
                    B.conn[fo + q] = uint8_t((L ? 1 : 0) | (U ? 2 : 0));
                    B.cnt[fo + q] = 0;                                    // (finite frame: every point gets a label)
                    if (write_normals) { B.nx[fo + q] = n1x; B.ny[fo + q] = n1y; B.nz[fo + q] = n1z; B.pd[fo + q] = d1; }
                }
                const unsigned linked = __ballot_sync(SPX_FULL, valid && L);
                const unsigned starts = ~linked | 1u;                     // lane 0 always starts a run inside the segment
                const int s0 = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
                if (valid) B.parent[fo + q] = r * w + tc + s0;
            }
        }
        if (k < K) {
            const int m = k + 1;
            const int rbase = r0 + 5;
            if (kTMA && rbase < h) {
                const unsigned bar = strip_smem_u32(&S.bar), parity = unsigned(m) & 1u;
                unsigned done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rbase + bp_rr[j];
                if ((j == 0 || tid < kSB * kSCW - kSThreads) && r >= 0 && r < h) {
                    float x = 0.f, y = 0.f, z = 0.f;
                    if (bp_cin[j]) {
                        if (kTMA) z = stage[bp_rr[j] * bw + bp_lx[j] * P.dis];
                        else z = *reinterpret_cast<const float *>(img + size_t(r) * P.samp_rstep + size_t(tc - 7 + bp_lx[j]) * P.dis * sizeof(float));
                        x = div_const(bp_xf[j] * z, P.fx, P.rfx, P.fast_div & 1);
                        y = div_const((float(r * P.dis) - P.cy) * z, P.fy, P.rfy, P.fast_div & 2);
                        if (bp_own[j]) {
                            const size_t o = fo + size_t(r * w + tc + bp_lx[j] - 7);
                            B.px[o] = x; B.py[o] = y; B.pz[o] = z;
                        }
                        // exactness bound: the unit in the last place of every non-zero coordinate
                        const unsigned ex = (__float_as_uint(x) >> 23) & 255u, ey = (__float_as_uint(y) >> 23) & 255u, ez = (__float_as_uint(z) >> 23) & 255u;
                        if (x != 0.f) negx = max(negx, 150 - int(ex ? ex : 1u));
                        if (y != 0.f) negy = max(negy, 150 - int(ey ? ey : 1u));
                        if (z != 0.f) negz = max(negz, 150 - int(ez ? ez : 1u));
                    }
                    const int sr = r % kSCRing;
                    S.cloud[sr][0][bp_lx[j]] = x; S.cloud[sr][1][bp_lx[j]] = y; S.cloud[sr][2][bp_lx[j]] = z;
                }
            }
        }
        __syncthreads();
        if (kTMA && tid == 0 && k + 1 < K && kSB * (k + 2) - 3 < h) issue_chunk(k + 2);   // into the stage phase A has just read
        if (k == K) break;

        // ================= phase B: column sums of difference rows [r0 + 4, r0 + 12) =================
        if (cs_on) {
            const int chn = cs_ch < 3 ? cs_ch : cs_ch - 3;
            const int lx = cs_dxl + 1;
#pragma unroll
            for (int i = 0; i < kSB; ++i) {
                const int y = r0 + 4 + i;
                if (y >= 0 && y < h) {
                    float d = 0.0f;
                    if (cs_colok && y >= 1 && y <= h - 2) {
                        if (cs_ch < 3) { const float *p = S.cloud[y % kSCRing][chn]; d = p[lx + 1] - p[lx - 1]; }
                        else d = S.cloud[(y + 1) % kSCRing][chn][lx] - S.cloud[(y - 1) % kSCRing][chn][lx];
                    }
                    colsum += double(d);
                    if (cs_owncol) sabs += fabsf(d);
                    double *row = S.sat[(y + 1) % kSRing][cs_ch];
                    row[cs_dxl + 1] = colsum;
                    if (cs_dxl == 0) row[0] = 0.0;
                }
            }
        }
        __syncthreads();
        // ================= phase C: row prefix of node rows [r0 + 5, r0 + 13) =================
        if (tid < kSB * 6) {
            const int i = tid / 6, ch = tid - i * 6;
            const int y = r0 + 4 + i;
            if (y >= 0 && y < h) {
                double *row = S.sat[(y + 1) % kSRing][ch];
                double run = 0.0;
#pragma unroll
                for (int j = 1; j < kSSW; ++j) { run += row[j]; row[j] = run; }
            }
        }
        __syncthreads();
        // ================= phase D: normals of rows [r0, r0 + 8) =================
        if (k >= 0 && wn_on) {
            const int r = r0 + wn_i;
            const int kk = k_next;
            k_next = win_k(r + kSB);
            if (r < h) {
                float nx = qnan, ny = qnan, nz = qnan, pd = qnan, Zs = qnan;
                if (wn_c >= 0 && wn_c < w) {
                    const int sc = r % kSCRing;
                    const float X = S.cloud[sc][0][wn_x + 6], Y = S.cloud[sc][1][wn_x + 6], Zv = S.cloud[sc][2][wn_x + 6];
                    Zs = Zv;
                    if (kk > 0 && isfinite(Zv)) {
                        const int half = kk / 2;
                        const int j0 = wn_x + 5 - half, j1 = j0 + kk;
                        const int s0 = (r - half + kSRing) % kSRing, s1 = (r - half + kk) % kSRing;
                        double g[6];
#pragma unroll
                        for (int ch = 0; ch < 6; ++ch)
                            g[ch] = ((S.sat[s1][ch][j1] + S.sat[s0][ch][j0]) - S.sat[s0][ch][j1]) - S.sat[s1][ch][j0];
                        // normal_vector = gradient_y.cross(gradient_x)
                        const double n0 = g[4] * g[2] - g[5] * g[1];
                        const double n1 = g[5] * g[0] - g[3] * g[2];
                        const double n2 = g[3] * g[1] - g[4] * g[0];
                        const double len = (n0 * n0 + n1 * n1) + n2 * n2;
                        if (len != 0.0) {
                            normalize_to_float(n0, n1, n2, len, nx, ny, nz);
                            // flipNormalTowardsViewpoint(point, 0, 0, 0, nx, ny, nz)
                            const float vx = 0.0f - X, vy = 0.0f - Y, vz = 0.0f - Zv;
                            const float cos_theta = (vx * nx + vy * ny + vz * nz);
                            if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
                        }
                    }
                    pd = dot3f(X, Y, Zv, nx, ny, nz);
                }
                const int sn = r % kSNRing;
                S.nrm[sn][0][wn_x] = nx; S.nrm[sn][1][wn_x] = ny; S.nrm[sn][2][wn_x] = nz; S.nrm[sn][3][wn_x] = pd; S.nrm[sn][4][wn_x] = Zs;
            }
        }
        __syncthreads();
    }

    // ---- the frame's exactness bound ----
    if (tid < 6) S.red_sum[tid] = 0.f;
    if (tid < 3) S.red_exp[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        negx = max(negx, __shfl_xor_sync(SPX_FULL, negx, o));
        negy = max(negy, __shfl_xor_sync(SPX_FULL, negy, o));
        negz = max(negz, __shfl_xor_sync(SPX_FULL, negz, o));
    }
    if (lane == 0) { atomicMax(&S.red_exp[0], negx); atomicMax(&S.red_exp[1], negy); atomicMax(&S.red_exp[2], negz); }
    if (cs_on) atomicAdd(&S.red_sum[cs_ch], sabs);
    __syncthreads();
    if (tid < 6) atomicAdd(&ctl.sat_sum[tid], S.red_sum[tid]);
    if (tid < 3) atomicMax(&ctl.sat_negexp[tid], S.red_exp[tid]);
}

}  // namespace spx
