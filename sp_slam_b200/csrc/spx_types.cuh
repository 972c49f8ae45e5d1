// spx_types.cuh -- device-side records shared by the kernels and the host API.
#pragma once
#include <cstdint>
#include "../../include/spx.h"

namespace spx {

// geometry + thresholds, passed by value to every kernel
struct Params {
    // depth image
    int rows, cols;
    size_t pitch;          // bytes
    size_t frame_stride;   // bytes
    // where the kernels that only need the organized cloud's samples (every Cloud.Dis-th row and column) read them:
    // bytes between consecutive SAMPLED rows / between frames.  The full image: dis * pitch, frame_stride; a staging
    // buffer that holds only the sampled rows (host input, sparse upload): its row pitch, h * pitch.
    size_t samp_rstep, samp_fstride;
    // ... and floats between consecutive SAMPLED columns / floats per row of that buffer.  The full image and the sampled-rows
    // buffer: dis, cols; the buffer of gathered samples (host input, upload mode 3: host threads pick the h x w samples out
    // of the caller's image and only those cross the bus): 1, w.
    int samp_cstep, samp_cols;
    // sparse upload (k_border_fetch): the caller's page-locked image the missing window sectors are fetched from
    size_t host_pitch, host_fstride;   // bytes
    float  full_alpha;     // 16-bit host image: depth = float(d) * full_alpha (mDepthMapFactor)
    int    fetch_vec;      // 16-byte aligned host rows and cols % 8 == 0: whole sectors move as vectors
    int    fetch_skip_sampled;   // the rows at multiples of Cloud.Dis are on the device already
    int n_frames;          // frames of this launch group
    int lines_in_global;   // test knob: run every k_lines item on the global-memory path (contours beyond the smem buffer use it)
    int refine_fast;       // 1: k_refine2 (a CTA per frame) where its table fits; 0: k_refine (a warp per frame) for all
    int compact;           // 1: the real planes' clouds leave as ordered inlier INDEX lists (out_pidx) instead of 16-byte points
    int idx16;             // compact: 16-bit indices (organized cloud of at most 65536 points), else 32-bit
    int frame0;            // first frame of the group (frames are independent: groups run on separate streams)
    // organized cloud (src/Frame.cc:856-874)
    int dis, w, h, N;
    float fx, fy, cx, cy;
    float rfx, rfy;        // RN(1 / fx), RN(1 / fy) for the strip kernel's division by a constant
    float min_axf, min_ayf; // smallest non-zero |n - cx| / |m - cy| over the sampled columns / rows of the current image
    int   forest_in_smem;  // k_ccl_frame rebuilds the row runs itself: the normals kernel need not initialise the forest in global memory
    int   fast_div;        // bit 0 / 1: that division was verified against the IEEE quotient for fx / fy (k_check_div)
    float min_x, max_x, min_y, max_y;
    float mdcf;            // max depth change factor
    int   min_size;
    float ang_cos;         // cosf(float(0.017453 * AngTh))   src/Frame.cc:900
    float dist_thr;        // Plane.DistanceThreshold
    double line_ratio;
    double line_thr;       // double(float Line.DistanceThreshold)
    int   ransac_max_iter;
    int   enable_supposed;
    int   n_grid;          // iterations of the `for(float i=-0.25;i<0.25;i=i+0.01)` loop (src/Frame.cc:1095)
    // per-frame arena capacities (elements)
    int contour_cap;       // contour index arena (2N + 16*SPX_MAX_MODELS)
    int pts_cap;           // points a frame may emit (N + line inliers + 32 supposed-plane grids)
    int bnd_cap;           // boundary points a frame may emit
};

struct Cand {              // connected component with size > Plane.MinSize
    int   root;            // minimum pixel index of the component
    int   label;           // PCL label = rank of the component by first raster pixel
    int   size;
    int   idx_off;         // offset of its raster-ordered index list in the frame's cand_idx arena
    float vec[3];          // smallest eigenvector (sign as eigen33 leaves it)
    float eigenvalue;
    float centroid[3];
    float curvature;
    float cov[9];
};

struct Model {             // accepted by the curvature test (model_coefficients[i] of segment())
    float coef[4];
    float centroid[3];
    float curvature;
    float cov[9];
    int   label;
    int   root;
    int   n0;              // inliers after segment()
    int   n1, n2;          // claimed in refine pass 1 / pass 2
    int   cand_off;        // raster-ordered originals in cand_idx
    int   last_inlier;     // inlier_indices[i].indices.back() after refine
    int   contour_off;     // slice of the frame's contour arena (capacity 2*(n0+n1+n2)+16)
    int   n_contour;
    int   plane;           // index into planes[] after the post filter, -1 = dropped by PlaneNotSeen
    int   n_rounds;        // SACSegmentation::segment calls made on this model's contour (0..4)
};

struct Line {              // one SACSegmentation::segment call of GeneratePlanesFromBoundries
    int   model;           // segmentation model whose contour is being fitted
    int   round;
    int   n_points, iterations, n_inliers, in_range, is_border, emitted;
    int   pts_off;         // line inlier points in the frame's line arena
    float coef[6];
};

struct PlaneRec {          // device-side spx_plane (frame-local offsets)
    float coef[4];
    int   n_points, n_boundary;
    int   points_off, boundary_off;
    int   src;
    int   is_supposed;
    int   line;            // supposed: index into lines[]; real: -1
    int   pad;
};

struct FrameCtl {
    int n_labels;          // label_indices.size() = number of components + 1
    int n_cand;
    int n_models;
    int n_real, n_planes;
    unsigned flags;
    int pts_used, bnd_used;   // points / boundary points of the REAL planes (set by k_postfilter)
    int pts_sup, bnd_sup;     // ... of the supposed planes (set by k_supposed)
    int n_lines;
    // exactness bound of the integral images (k_normals_strip; evaluated by k_models): the smallest non-zero depth of the frame
    // (with the smallest |n - cx|, |m - cy| of the image it bounds the finest unit in the last place of the cloud's coordinates)
    // and per gradient channel the sum of |central differences|
    unsigned sat_zinv;     // 0x7f800000 - bits of the smallest non-zero |z| of the cloud (0: none)
    float sat_sum[6];
    int pad[4];
    Cand     cand[SPX_MAX_CAND];
    Model    models[SPX_MAX_MODELS];
    PlaneRec planes[SPX_MAX_PLANES];
    Line     lines[SPX_MAX_MODELS * SPX_MAX_LINES];
};

// all device buffers of a context; per-frame arrays are frame-major (frame f at f * stride)
struct Buffers {
    float *px, *py, *pz;          // organized cloud, N per frame
    float *dist;                  // chamfer distance, clamped at 10 (parity tap only)
    uint8_t *kwin;                // smoothing window size int(min(dist, 10)), 0 where it is <= 2
    float *cham_tmp;              // forward-pass rows of the chamfer bands
    float *nx, *ny, *nz, *pd;     // normals (NaN = invalid) and plane_d
    uint8_t *conn;                // bit0: comparator edge to the left pixel, bit1: to the upper pixel
    int   *parent;                // union-find forest, then flattened roots
    int   *cnt;                   // component size at its root
    int   *lab;                   // PCL labels (rank of the component)
    int16_t *root_model;          // model index at component roots, -1 otherwise
    int8_t *pid;                  // plane (model) id per pixel, -1 = none
    int8_t *pid_bak;              // plane ids after the first refine pass (restored if the speculative second pass fails)
    int   *pos;                   // position of the pixel in its model's inlier list
    int   *cand_idx;              // raster-ordered index lists of the candidates
    int   *contour_idx;           // contour arena, contour_cap per frame
    float4 *line_a;               // RANSAC working cloud of contours too long for shared memory (contour_cap per frame)
    int   *work2;                 // k_border queue: [0] items, [1] pull counter, [2..] (frame * SPX_MAX_MODELS + model) * SPX_MAX_LINES + round
    int   *work;                  // k_lines queue: [0] items, [1] pull counter, [2..] frame * SPX_MAX_MODELS + model
    int   *line_sh;               // shuffled indices (contour_cap per frame)
    int   *line_inl;              // inlier index scratch (contour_cap per frame)
    spx_point *line_pts;          // accepted line inlier points (contour_cap per frame)
    FrameCtl *ctl;
    int *nf_list;                 // frames of a group whose depth holds NaN / Inf ([0] = count, then frame indices): they take k_normals_link_list
    unsigned *fetch_bits;         // sparse upload: claimed 8-pixel sectors, rows x ceil(ceil(cols/8)/32) words per frame
    // compacted outputs.  The point / boundary arenas hold the clouds of all REAL planes first and those of all supposed
    // planes behind them, so the real part can be packed while the line fits still run.  frame_offs[5f + k]: k = 0 plane
    // records, 1 / 2 points / boundary points of the frame's real planes, 3 / 4 of its supposed planes (all absolute).
    spx_frame_header *out_frames;
    spx_plane *out_planes;
    spx_point *out_pts;
    spx_point *out_bnd;
    void *out_pidx;               // compact results: inlier_indices of the real planes (uint16_t or uint32_t, N per frame)
    long long *out_totals;        // slots of 8: [0] planes, [1] points, [2] boundary points; [5], [6] the real-plane part of [1], [2]
    long long *frame_offs;        // 5 per frame
};

}  // namespace spx
