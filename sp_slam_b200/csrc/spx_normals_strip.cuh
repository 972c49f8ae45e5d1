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
constexpr int kSCRing = 16;              // cloud rows in flight: 13 needed (a power of two: the slot is row & 15)
constexpr int kSNRing = 9;               // rows of normals the link step looks at: the batch and the row above it
constexpr int kSThreads = 288;           // 9 warps: 8 x 33 = 264 normals in one round, 6 x 42 = 252 column sums

struct StripSmem {
    double sat[kSRing][6][kSSW];         // 37 152 B
    float cloud[kSCRing][3][kSCW];       //  8 448 B
    float nrm[kSNRing][5][kSNW];         //  5 940 B: nx ny nz plane_d z
    unsigned long long bar;
    float red_sum[6];
    unsigned red_z;
};
// The TMA destination follows the fixed part, 128-byte aligned: one depth chunk of 8 sampled rows.  TMA has no element
// stride in dimension 0 (cuda.h: "the first element of this array is ignored"), so the row segments arrive whole and the
// back-projection reads every Cloud.Dis-th float (a stride of 3 words is conflict free); and the box must START on a
// 16-byte boundary of the row (tools/tma_probe3.cu: any other start coordinate raises "illegal instruction"), so the
// segment begins at the multiple of four columns at or below the strip's first sample and is up to three floats longer.
constexpr size_t kStripStageOff = (sizeof(StripSmem) + 127) / 128 * 128;
inline int strip_box_w(int cstep) { return ((kSCW - 1) * cstep + 1 + 3 + 3) & ~3; }      // cstep = Params::samp_cstep
inline size_t strip_smem_bytes(int dis, bool tma) { return kStripStageOff + (tma ? size_t(kSB) * strip_box_w(dis) * 4 : 0); }
inline bool strip_tma_ok(int dis) { return strip_box_w(dis) <= 256; }

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
    // a zero (the depth of an invalid pixel is 0): +-0 / b = +-0 with the product's sign -- no division on the sensor's dropouts
    if (fast && aa == 0.0f) return __fmul_rn(a, rb);
    return a / b;
}

// exhaustive check of div_const for one constant: all 2^23 significands of a (the quotient of scaled operands scales exactly)
__global__ void __launch_bounds__(256) k_check_div(float b, float rb, int *mismatch) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << 23)) return;
    const float a = __uint_as_float(0x3f800000u | i);
    if (div_const(a, b, rb, 1) != a / b || div_const(-a, b, rb, 1) != -a / b) atomicAdd(mismatch, 1);
}

// compile-time switches handed to the phase lambdas
template <bool V> struct StripBool { static constexpr bool value = V; };
template <int V> struct StripInt { static constexpr int value = V; };

template <bool kTMA, int kOcc>
__global__ void __launch_bounds__(kSThreads, kOcc)
k_normals_strip(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ depth, Params P, Buffers B, int write_normals) {
    extern __shared__ __align__(128) unsigned char strip_raw[];
    StripSmem &S = *reinterpret_cast<StripSmem *>(strip_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int f = P.frame0 + blockIdx.y;
    const int tc = blockIdx.x * kStW;
    const int w = P.w, h = P.h;
    FrameCtl &ctl = B.ctl[f];
    if (ctl.flags & unsigned(SPX_FRAME_NONFINITE)) return;      // NaN / Inf depth: the frame goes through k_normals_link_list (finite-count images)
    const size_t fo = size_t(f) * P.N;
    const int K = (h + kSB - 1) / kSB;                          // batches
    float *const cl_base = &S.cloud[0][0][0];
    double *const nd_base = &S.sat[0][0][0];
    float *const nr_base = &S.nrm[0][0][0];
    constexpr int kClRow = 3 * kSCW, kNdRow = 6 * kSSW, kNrRow = 5 * kSNW;     // elements per ring row

    float *const stage = reinterpret_cast<float *>(strip_raw + kStripStageOff);
    const int x_first = (tc - 7) * P.samp_cstep;                // column of the strip's first sample in the sampling buffer
    const int x_box = x_first & ~3;                             // where the staged row segment starts (16-byte aligned)
    const int bw = ((kSCW - 1) * P.samp_cstep + 7) & ~3;        // floats per staged row (strip_box_w)
    if (kTMA && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(strip_smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // node row 0 of the integral images is zero, and so is node column 0 of every row (never written again)
    for (int i = tid; i < 6 * kSSW; i += kSThreads) nd_base[i] = 0.0;
    for (int i = tid; i < kSRing * 6; i += kSThreads) nd_base[i * kSSW] = 0.0;
    __syncthreads();
    // chunk m = sampled rows [8m - 3, 8m + 5); rows / columns outside the image arrive as zeros.  One stage: chunk m + 1 is
    // requested as soon as phase A has consumed chunk m and lands during phases B..D.
    auto issue_chunk = [&](int m) {
        const unsigned bar = strip_smem_u32(&S.bar), dst = strip_smem_u32(stage);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kSB * bw * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&tmap), "r"(x_box), "r"(kSB * m - 3), "r"(f), "r"(bar) : "memory");
    };
    if (kTMA && tid == 0) issue_chunk(0);

    // ---------------- per-thread constants and running state (whatever depends on the row advances by additions) ----------------
    // back-projection: thread = (cloud column lx, chunk row rr in 0..5); threads 0..87 also take chunk rows 6, 7
    const int bp_rr = tid / kSCW, bp_lx = tid - bp_rr * kSCW;
    const int bp_c = tc - 7 + bp_lx;
    const bool bp_on = tid < 6 * kSCW, bp_two = tid < 2 * kSCW;
    const bool bp_cin = bp_on && bp_c >= 0 && bp_c < w;
    const bool bp_own = bp_cin && bp_lx >= 7 && bp_lx < 7 + kStW;
    const float bp_xf = float(bp_c * P.dis) - P.cx;
    const float *const img = reinterpret_cast<const float *>(reinterpret_cast<const char *>(depth) + size_t(f) * P.samp_fstride);   // (CTA uniform)
    const int rstep_f = int(P.samp_rstep / sizeof(float));
    float *const gx = B.px + fo, *const gy = B.py + fo, *const gz = B.pz + fo;
    int bp_go = (5 - kSB + bp_rr) * w + bp_c;                   // index of the thread's first point of the chunk in the cloud arrays
    int bp_so = kTMA ? bp_rr * bw + (x_first - x_box) + bp_lx * P.samp_cstep      // its sample in the staged chunk ...
                     : (5 - kSB + bp_rr) * rstep_f + bp_c * P.samp_cstep;         // ... or in the frame's image (plain loads)
    float bp_z0 = 0.f, bp_z1 = 0.f;                             // plain loads: the samples of the next chunk, fetched one phase round ahead
    unsigned zmin = 0x7f800000u;                                // smallest non-zero |z| the thread has seen (as bits)
    // column sums: warps 0..3 = d/dx (channels 0..2 of x y z), warps 4..7 = d/dy (channels 3..5); 42 difference columns each
    const int cs_i = tid & 127;
    const bool cs_dx = tid < 128;
    const bool cs_on = tid < 256 && cs_i < 3 * kSDW;
    const int cs_chn = cs_on ? cs_i / kSDW : 0, cs_dxl = cs_i - cs_chn * kSDW;
    const int cs_ch = cs_chn + (cs_dx ? 0 : 3);
    const int cs_c = tc - 6 + cs_dxl;
    const bool cs_colok = cs_on && cs_c >= 1 && cs_c <= w - 2;
    const bool cs_owncol = cs_on && cs_dxl >= 6 && cs_dxl < 6 + kStW;
    const int cs_cl = cs_chn * kSCW + cs_dxl + 1;               // the thread's column in a cloud-ring row
    const int cs_nd = cs_ch * kSSW + cs_dxl + 1;                // ... and in a node-ring row
    double colsum = 0.0;
    float sabs = 0.0f;                                          // exactness bound: sum of |differences| of the strip's own columns
    float dy_m = 0.f, dy_0 = 0.f;                               // d/dy threads: the cloud values of rows y - 1 and y, carried along
    int nb = 5 - kSB;                                           // node-ring slot of node row r0 + 5 (the first this iteration writes)
    // window pixel of phase D: (row i, column nx_) of the 8 x 33 block
    const int wn_i = tid / kSNW, wn_x = tid - wn_i * kSNW;
    const bool wn_on = tid < kSB * kSNW;
    const int wn_c = tc - 1 + wn_x;
    const bool wn_cin = wn_on && wn_c >= 0 && wn_c < w;
    constexpr int border = 10;
    const bool wn_cwin = wn_on && wn_c >= border && wn_c < w - border;
    const float qnan = __int_as_float(0x7fc00000);
    const uint8_t *const kwin = B.kwin + fo;
    int kw_o = wn_c + wn_i * w;                                 // the pixel's window size, advanced by 8 rows per batch
    auto win_k = [&](int r) -> int { return (wn_cwin && r >= border && r < h - border) ? int(kwin[kw_o]) : 0; };
    int k_next = win_k(wn_i);                                   // prefetched one batch ahead
    int wn_s18 = wn_i;                                          // node-ring slot of the pixel's row (r mod 18)
    int wn_s9 = wn_i;                                           // normal-ring slot of the pixel's row (r mod 9): + 8 = - 1 per batch
    const int wn_cl = wn_i * kClRow + wn_x + 6;
    // links: warp = row of the batch, lane = column
    const int lk_c = tc + lane;
    const bool lk_valid = lk_c < w;
    int lk_q = (wid - 2 * kSB) * w + lk_c;                      // pixel index of the warp's row of batch k - 1 (at k = -1, where the loop starts)
    int lk_s9 = (wid + 2) % kSNRing;                            // its normal-ring slot: (wid - 16) mod 9

    // ================= phase A, part 1: links of batch k - 1 (rows r0 - 8 .. r0 - 1) =================
    auto links = [&](auto edge_c, int r0) {
        constexpr bool kEdge = decltype(edge_c)::value;
        if (wid >= 8) return;
        const int r = r0 - kSB + wid;
        if (kEdge && (r < 0 || r >= h)) return;                           // warp uniform
        bool L = false, U = false;
        if (lk_valid) {
            const float *nr = nr_base + lk_s9 * kNrRow + lane;
            const float n1x = nr[1], n1y = nr[kSNW + 1], n1z = nr[2 * kSNW + 1], d1 = nr[3 * kSNW + 1], z = nr[4 * kSNW + 1];
            float threshold = P.dist_thr;
            threshold *= z * z;                                           // vec.dot(z_axis_) = z
            if (lk_c >= 1)
                L = (fabsf(d1 - nr[3 * kSNW]) < threshold) && (dot3f(n1x, n1y, n1z, nr[0], nr[kSNW], nr[2 * kSNW]) > P.ang_cos);
            if (!kEdge || r >= 1) {
                const float *nu = nr_base + (lk_s9 == 0 ? kSNRing - 1 : lk_s9 - 1) * kNrRow + lane + 1;
                U = (fabsf(d1 - nu[3 * kSNW]) < threshold) && (dot3f(n1x, n1y, n1z, nu[0], nu[kSNW], nu[2 * kSNW]) > P.ang_cos);
            }
            B.conn[fo + lk_q] = uint8_t((L ? 1 : 0) | (U ? 2 : 0));
            B.cnt[fo + lk_q] = 0;                                         // (finite frame: every point gets a label)
            if (write_normals) { B.nx[fo + lk_q] = n1x; B.ny[fo + lk_q] = n1y; B.nz[fo + lk_q] = n1z; B.pd[fo + lk_q] = d1; }
        }
        if (!P.forest_in_smem) {                                          // (k_ccl_frame rebuilds the row runs from the link bits)
            const unsigned linked = __ballot_sync(SPX_FULL, lk_valid && L);
            const unsigned starts = ~linked | 1u;                         // lane 0 always starts a run inside the segment
            const int s0 = 31 - __clz(starts & (SPX_FULL >> (31 - lane)));
            if (lk_valid) B.parent[fo + lk_q] = lk_q - lane + s0;
        }
    };
    // ================= phase A, part 2: back-projection of cloud rows [r0 + 5, r0 + 13) =================
    auto backproject = [&](auto edge_c, auto par_c, int r0) {
        constexpr bool kEdge = decltype(edge_c)::value;
        constexpr int kPar = decltype(par_c)::value;
        if (kTMA && (!kEdge || r0 + 5 < h)) {
            const unsigned bar = strip_smem_u32(&S.bar), parity = unsigned(kPar ^ 1);     // chunk k + 1
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        if (bp_on) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j == 1 && !bp_two) break;
                const int rr = bp_rr + 6 * j, r = r0 + 5 + rr;
                if (kEdge && (r < 0 || r >= h)) continue;
                float x = 0.f, y = 0.f, z = 0.f;
                if (bp_cin) {
                    z = kTMA ? stage[bp_so + 6 * j * bw] : (j ? bp_z1 : bp_z0);
                    x = div_const(bp_xf * z, P.fx, P.rfx, P.fast_div & 1);
                    y = div_const((float(r * P.dis) - P.cy) * z, P.fy, P.rfy, P.fast_div & 2);
                    if (bp_own) { const int o = bp_go + 6 * j * w; gx[o] = x; gy[o] = y; gz[o] = z; }
                    const unsigned uz = __float_as_uint(z) & 0x7fffffffu;
                    if (uz) zmin = min(zmin, uz);
                }
                float *cl = cl_base + ((8 * kPar + 5 + rr) & (kSCRing - 1)) * kClRow + bp_lx;
                cl[0] = x; cl[kSCW] = y; cl[2 * kSCW] = z;
            }
            bp_go += kSB * w;
            if (!kTMA) {                                                  // the next chunk's samples: in flight during phases B..D
                bp_so += kSB * rstep_f;
                const int r = r0 + 5 + kSB + bp_rr;
                if (bp_cin && r >= 0 && r < h) bp_z0 = img[bp_so];
                if (bp_two && bp_cin && r + 6 >= 0 && r + 6 < h) bp_z1 = img[bp_so + 6 * rstep_f];
            }
        }
    };
    // ================= phase B: column sums of difference rows [r0 + 4, r0 + 12) -> node rows [r0 + 5, r0 + 13) =================
    auto colsums = [&](auto edge_c, auto par_c, int r0) {
        constexpr bool kEdge = decltype(edge_c)::value;
        constexpr int kPar = decltype(par_c)::value;
        if (!cs_on) return;
        const int wrap_at = kSRing - nb;                                  // row i == wrap_at of this batch wraps around the node ring
        double *nd = nd_base + cs_nd + nb * kNdRow;
        const float *const cl0 = cl_base + cs_cl;
        if (!kEdge) {
            // steady batch: every row is an interior row of the frame.  Branch-free per row: the column test becomes a select (the
            // ring holds zeros outside the frame, so the loads are always in bounds), the d/dx - d/dy split is taken once, the ring
            // wrap is a choice between two base pointers with compile-time row offsets.
            double *const ndw = nd - kSRing * kNdRow;
            if (cs_dx) {
#pragma unroll
                for (int i = 0; i < kSB; ++i) {
                    const float *p = cl0 + ((8 * kPar + 4 + i) & (kSCRing - 1)) * kClRow;
                    const float dd = p[1] - p[-1];
                    const float d = cs_colok ? dd : 0.0f;
                    colsum += double(d);
                    sabs += cs_owncol ? fabsf(d) : 0.0f;
                    (i >= wrap_at ? ndw : nd)[i * kNdRow] = colsum;
                }
            } else {
#pragma unroll
                for (int i = 0; i < kSB; ++i) {
                    const float dy_p = cl0[((8 * kPar + 5 + i) & (kSCRing - 1)) * kClRow];
                    const float dd = dy_p - dy_m;
                    const float d = cs_colok ? dd : 0.0f;
                    dy_m = dy_0; dy_0 = dy_p;
                    colsum += double(d);
                    sabs += cs_owncol ? fabsf(d) : 0.0f;
                    (i >= wrap_at ? ndw : nd)[i * kNdRow] = colsum;
                }
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < kSB; ++i) {
            const int y = r0 + 4 + i;                                     // (row tests are uniform over the CTA)
            if (i == wrap_at) nd -= kSRing * kNdRow;                      // (uniform)
            if (kEdge && (y < 0 || y >= h)) { nd += kNdRow; continue; }
            const bool inner = !kEdge || (y >= 1 && y <= h - 2);
            float d = 0.0f;
            if (cs_dx) {
                const float *p = cl0 + ((8 * kPar + 4 + i) & (kSCRing - 1)) * kClRow;
                if (inner && cs_colok) d = p[1] - p[-1];
            } else {
                if (kEdge && y == 0) dy_0 = cl0[((8 * kPar + 4 + i) & (kSCRing - 1)) * kClRow];      // the chain starts: row y itself
                const float dy_p = (!kEdge || y + 1 < h) ? cl0[((8 * kPar + 5 + i) & (kSCRing - 1)) * kClRow] : 0.0f;
                if (inner && cs_colok) d = dy_p - dy_m;
                dy_m = dy_0; dy_0 = dy_p;
            }
            colsum += double(d);
            if (cs_owncol) sabs += fabsf(d);
            *nd = colsum;
            nd += kNdRow;
        }
    };
    // ================= phase C: row prefix of node rows [r0 + 5, r0 + 13) =================
    auto prefix = [&](auto edge_c, int r0) {
        constexpr bool kEdge = decltype(edge_c)::value;
        if (tid >= kSB * 6) return;
        const int i = tid / 6, ch = tid - i * 6;
        const int y = r0 + 4 + i;
        if (kEdge && (y < 0 || y >= h)) return;
        int slot = nb + i; if (slot >= kSRing) slot -= kSRing;
        double *row = nd_base + slot * kNdRow + ch * kSSW;
        double run = 0.0;
#pragma unroll
        for (int j = 1; j < kSSW; ++j) { run += row[j]; row[j] = run; }
    };
    // ================= phase D: normals of rows [r0, r0 + 8) =================
    auto normals = [&](auto edge_c, auto par_c, int r0) {
        constexpr bool kEdge = decltype(edge_c)::value;
        constexpr int kPar = decltype(par_c)::value;
        if (!wn_on) return;
        const int r = r0 + wn_i;
        const int kk = k_next;
        kw_o += kSB * w;
        k_next = win_k(r + kSB);
        if (!kEdge || r < h) {
            float nx = qnan, ny = qnan, nz = qnan, pd = qnan, Zs = qnan;
            if (wn_cin) {
                const float *cl = cl_base + wn_cl + 8 * kPar * kClRow;
                const float X = cl[0], Y = cl[kSCW], Zv = cl[2 * kSCW];
                Zs = Zv;
                if (kk > 0 && isfinite(Zv)) {
                    const int half = kk >> 1;
                    int s0 = wn_s18 - half; if (s0 < 0) s0 += kSRing;
                    int s1 = s0 + kk; if (s1 >= kSRing) s1 -= kSRing;
                    const double *a0 = nd_base + s0 * kNdRow + wn_x + 5 - half, *a1 = nd_base + s1 * kNdRow + wn_x + 5 - half;
                    double g[6];
#pragma unroll
                    for (int ch = 0; ch < 6; ++ch)
                        g[ch] = ((a1[ch * kSSW + kk] + a0[ch * kSSW]) - a0[ch * kSSW + kk]) - a1[ch * kSSW];
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
            float *nr = nr_base + wn_s9 * kNrRow + wn_x;
            nr[0] = nx; nr[kSNW] = ny; nr[2 * kSNW] = nz; nr[3 * kSNW] = pd; nr[4 * kSNW] = Zs;
        }
        wn_s18 += kSB; if (wn_s18 >= kSRing) wn_s18 -= kSRing;
        wn_s9 = wn_s9 == 0 ? kSNRing - 1 : wn_s9 - 1;
    };

    // one batch: phase A (links of batch k - 1, cloud rows of chunk k + 1), B, C, D; kEdge = some row the phases touch may lie
    // outside the frame or on its first / last row (the first two and the last one or two batches), kPar = k & 1
    auto iteration = [&](auto edge_c, auto par_c, int k) {
        constexpr bool kEdge = decltype(edge_c)::value;
        const int r0 = kSB * k;
        if (!kEdge || k >= 1) links(edge_c, r0);
        lk_q += kSB * w; lk_s9 = lk_s9 == 0 ? kSNRing - 1 : lk_s9 - 1;
        if (!kEdge || k < K) backproject(edge_c, par_c, r0);
        __syncthreads();
        if (kTMA && tid == 0 && k + 1 < K && kSB * (k + 2) - 3 < h) issue_chunk(k + 2);   // into the stage phase A has just read
        if (kEdge && k == K) return;
        colsums(edge_c, par_c, r0);
        __syncthreads();
        prefix(edge_c, r0);
        __syncthreads();
        if (!kEdge || k >= 0) normals(edge_c, par_c, r0);
        nb += kSB; if (nb >= kSRing) nb -= kSRing;
        __syncthreads();
    };
    if (!kTMA && bp_on && bp_cin) {                                       // plain loads: the first chunk's samples (rows -3 .. 4)
        const int r = 5 - kSB + bp_rr;
        if (r >= 0 && r < h) bp_z0 = img[bp_so];
        if (bp_two && r + 6 >= 0 && r + 6 < h) bp_z1 = img[bp_so + 6 * rstep_f];
    }
    nb = (nb + kSRing) % kSRing;
    for (int k = -1; k <= K; ++k) {
        const bool steady = k >= 2 && kSB * k + 12 <= h - 1;
        if (steady) { if (k & 1) iteration(StripBool<false>(), StripInt<1>(), k); else iteration(StripBool<false>(), StripInt<0>(), k); }
        else        { if (k & 1) iteration(StripBool<true>(), StripInt<1>(), k); else iteration(StripBool<true>(), StripInt<0>(), k); }
    }

    // ---- the frame's exactness bound: per channel the sum of |differences|, and the smallest non-zero depth ----
    if (tid < 6) S.red_sum[tid] = 0.f;
    if (tid == 0) S.red_z = 0u;
    __syncthreads();
    unsigned zinv = 0x7f800000u - zmin;                                   // (larger = smaller depth; 0 = none seen)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zinv = max(zinv, __shfl_xor_sync(SPX_FULL, zinv, o));
    if (lane == 0) atomicMax(&S.red_z, zinv);
    if (cs_on) atomicAdd(&S.red_sum[cs_ch], sabs);
    __syncthreads();
    if (tid < 6) atomicAdd(&ctl.sat_sum[tid], S.red_sum[tid]);
    if (tid == 0) atomicMax(&ctl.sat_zinv, S.red_z);
}

}  // namespace spx
