/*
 * spx_oracle.h -- C API of the CPU oracle for SP-SLAM's plane-extraction front end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in PCL 1.8.0 (pinned by the reference's
 * build.sh:4, README.md:25, CMakeLists.txt:42), which is neither vendored in /root/reference nor
 * installed here, and the reference ships no tests or golden vectors for it.  This oracle restates
 * the reference's own code (src/Frame.cc:854-1144) and the published PCL 1.8.0 algorithms it
 * calls (see spx_oracle.cpp for per-function citations and the list of under-determined choices).
 */
#ifndef SPX_ORACLE_H
#define SPX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_config {
    /* YAML keys read through Config::Get (Examples/RGB-D/TUM1.yaml:73-78,99-100) */
    int32_t cloud_dis;          /* Cloud.Dis                src/Frame.cc:856 */
    int32_t min_size;           /* Plane.MinSize            src/Frame.cc:888 */
    float   angle_thr_deg;      /* Plane.AngleThreshold     src/Frame.cc:889 */
    float   dist_thr;           /* Plane.DistanceThreshold  src/Frame.cc:890 */
    double  line_ratio;         /* Line.Ratio               src/Frame.cc:941 */
    float   line_dist_thr;      /* Line.DistanceThreshold   src/Frame.cc:942 */
    /* Frame statics (src/Frame.cc:169-172) and image bounds (src/Frame.cc:536-564) */
    float fx, fy, cx, cy;
    float min_x, max_x, min_y, max_y;
    /* hard-coded in src/Frame.cc:881-882,945 */
    float   max_depth_change_factor;   /* 0.05f */
    float   normal_smoothing_size;     /* 10.0f */
    int32_t ransac_max_iter;           /* 1000  */
    int32_t enable_supposed;           /* run GeneratePlanesFromBoundries (src/Frame.cc:194) */
    /* Alternative readings of PCL 1.8.0 where the restatement rests on recollection (SURVEY.md Appendix A, items marked
     * with a warning sign).  0 = the reading this oracle and the CUDA path implement.  A future run of tools/pcl_pin against
     * real PCL that disagrees with the oracle can be bisected by flipping these one at a time. */
    uint32_t alt;
    /* normal estimation method of IntegralImageNormalEstimation: 0 = AVERAGE_3D_GRADIENT (what the reference selects,
     * src/Frame.cc:880), 1 = COVARIANCE_MATRIX (PCL's 9-channel integral image of x y z and their products, per-pixel covariance,
     * eigen33, curvature: the method BASELINE.json's north_star words; not used by the reference) */
    int32_t normal_method;
} orc_config;
#define ORC_ALT_VP_RESET         1u   /* A.4: segment()'s viewpoint vector is reset for every cluster (default: it accumulates -centroid) */
#define ORC_ALT_CHAMFER_NO_WRAP  2u   /* A.2: no row wrap-around in the distance-map passes (default: previous_row[w] aliases current_row[0]) */
#define ORC_ALT_REFINE_NO_WRAP   4u   /* A.5: refine()'s second pass makes no left claim at column 0 (default: it claims the previous row's last pixel) */
#define ORC_ALT_SAMPLE_GOOD_OR   8u   /* A.8: isSampleGood accepts a pair when ANY coordinate differs (default: x, y and z must all differ) */
#define ORC_ALT_RNG_MASK        16u   /* A.8: rnd() = mt() & INT_MAX (default: boost::uniform_int(0, INT_MAX) = mt() >> 1) */
#define ORC_ALT_SO_DOUBLE       32u   /* A.2(5): the second-order products x*x, x*y, ... are formed in double (default: float product, then widened) */
#define ORC_ALT_COV_TRACE       64u   /* A.2(5): curvature = lambda / (C00 + C11 + C22) (default: PCL's coeff(0) + coeff(2) + coeff(4)) */

/* 16-byte packed point: xyz + packed rgba (a<<24|r<<16|g<<8|b), the payload of pcl::PointXYZRGB */
typedef struct orc_point { float x, y, z; uint32_t rgba; } orc_point;

typedef struct orc_ctx orc_ctx;

void      orc_default_config(orc_config *cfg);
orc_ctx * orc_create(const orc_config *cfg);
void      orc_destroy(orc_ctx *);

/* Run the whole path on one depth image (CV_32F metres, row-major, pitch in floats = cols).
 * normals_in: NULL, or 3*N floats (nx[N],ny[N],nz[N] SoA) to bypass normal estimation
 * ("feed the reference's normals" mode).  Returns 0. */
int orc_run(orc_ctx *, const float *depth, int rows, int cols, const float *normals_in);

/* ---- intermediates of the last orc_run (sizes: N = width*height of the organized cloud) ---- */
void   orc_dims(const orc_ctx *, int *width, int *height);
void   orc_get_cloud(const orc_ctx *, float *x, float *y, float *z);          /* N each */
void   orc_get_distance_map(const orc_ctx *, float *dist);                    /* N, unclamped */
void   orc_get_normals(const orc_ctx *, float *nx, float *ny, float *nz);     /* N each, NaN = invalid */
void   orc_get_plane_d(const orc_ctx *, float *d);                            /* N */
void   orc_get_curvature(const orc_ctx *, float *curv);                       /* N: Normal::curvature (NaN under AVERAGE_3D_GRADIENT) */
int    orc_get_labels_raw(const orc_ctx *, uint32_t *labels);                 /* N; returns label_indices.size() */
void   orc_get_labels_refined(const orc_ctx *, uint32_t *labels);             /* N */
/* models accepted by OrganizedMultiPlaneSegmentation::segment (before SP-SLAM's post filter) */
int    orc_num_models(const orc_ctx *);
void   orc_get_model(const orc_ctx *, int i, float coef[4], float centroid[3], float cov[9],
                     float *curvature, uint32_t *label, int *n_inliers_segment,
                     int *n_inliers_refined, int *n_contour);
void   orc_get_model_inliers(const orc_ctx *, int i, int32_t *idx);           /* n_inliers_refined */
void   orc_get_model_contour(const orc_ctx *, int i, int32_t *idx);           /* n_contour */
/* was the SAT arithmetic exact (order independent)?  1 = every fp64 partial sum was exact */
int    orc_sat_exact(const orc_ctx *);

/* ---- final Frame fields (src/Frame.cc:187,199; include/Frame.h:223-244) ---- */
int    orc_num_real_planes(const orc_ctx *);     /* mnRealPlaneNum */
int    orc_num_planes(const orc_ctx *);          /* mnPlaneNum     */
void   orc_get_plane(const orc_ctx *, int i, float coef[4], int *n_points, int *n_boundary,
                     int *src_model /* model index for real planes, parent plane for supposed */);
void   orc_get_plane_points(const orc_ctx *, int i, orc_point *pts);
void   orc_get_plane_boundary(const orc_ctx *, int i, orc_point *pts);

/* ---- RANSAC line log of GeneratePlanesFromBoundries: one record per segLine.segment call ---- */
typedef struct orc_line_rec {
    int32_t plane;          /* index into mvBoundaryPoints being processed */
    int32_t round;          /* j in 0..3 */
    int32_t n_points;       /* size of boundPoints at the call */
    int32_t iterations;     /* RANSAC iterations_ at exit */
    int32_t n_inliers;      /* refined inlier count returned by segment() */
    int32_t in_range;       /* LineInRange */
    int32_t is_border;      /* IsBorderLine (only evaluated if in_range) */
    int32_t emitted;        /* CaculatePlanes returned true */
    float   coef[6];        /* optimised line: centroid + direction */
} orc_line_rec;
int    orc_num_line_recs(const orc_ctx *);
void   orc_get_line_recs(const orc_ctx *, orc_line_rec *out);

/* seconds spent in the two Timer sections of the last run (src/Frame.cc:184-197) */
void   orc_get_times(const orc_ctx *, double *t_plane, double *t_splane);

/* Frame-parallel run of the whole path on host threads (one orc_ctx per thread; frames are independent): the CPU
 * baseline of bench.py.  depth: n_frames contiguous images.  Fills per-frame mnRealPlaneNum / mnPlaneNum and the summed
 * Timer sections (seconds of CPU time over all frames).  Returns 0. */
int    orc_run_batch(const orc_config *cfg, const float *depth, int n_frames, int rows, int cols, int n_threads,
                     int32_t *n_real, int32_t *n_planes, double *t_plane_sum, double *t_splane_sum);

/* exact[f] = 1 when every fp64 partial sum of frame f's integral images (and every window sum) was exact: the device's
 * tile-local integral images then give bit-identical window sums to PCL's whole-image ones.  Runs cloud + normals only. */
int    orc_sat_exact_batch(const orc_config *cfg, const float *depth, int n_frames, int rows, int cols, int n_threads, uint8_t *exact);

/* ---- stage-level entry points used by known-answer tests ---- */
/* two-pass chamfer of PCL's computeFeature on a caller-supplied mask (0 = edge) */
void   orc_chamfer(const uint8_t *mask, int width, int height, float *dist);
void   orc_chamfer_alt(const uint8_t *mask, int width, int height, float *dist, uint32_t alt);
/* Frame::IsBorderPoint (src/Frame.cc:1026-1056) of the camera-frame point (x, y, z) against a depth image */
int    orc_is_border_point(const orc_config *cfg, const float *depth, int rows, int cols, float x, float y, float z);
/* pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair of a symmetric 3x3 (row-major) */
void   orc_eigen33_smallest(const float cov[9], float *eigenvalue, float eigenvector[3]);
/* pcl::eigen33(mat, evals) + computeCorrespondingEigenVector(mat, evals[2]) */
void   orc_eigen33_largest(const float cov[9], float evals[3], float eigenvector[3]);
/* SACSegmentation<LINE>::segment on a point list; returns n_inliers, fills coef[6], inlier idx, iterations */
int    orc_sac_line(const orc_point *pts, int n, double threshold, int max_iter,
                    float coef[6], int32_t *inliers, int *iterations);
int    orc_sac_line_alt(const orc_point *pts, int n, double threshold, int max_iter,
                        float coef[6], int32_t *inliers, int *iterations, uint32_t alt);
/* the index pairs RANSAC would draw for a cloud of n points, ignoring isSampleGood rejections */
void   orc_ransac_draws(int n, int n_draws, int32_t *pairs /* 2*n_draws */);

/* ---- SURVEY 8(f) rows ---- */
/* N4: pcl::VoxelGrid<PointXYZRGB> (leaf per axis, defaults otherwise) on one cloud; returns the number of output points;
 * out_idx (optional): the voxel index of every output point (-1 for the "leaf size too small" copy path) */
int    orc_voxel_grid(const orc_point *pts, int n, const float leaf[3], orc_point *out, int32_t *out_idx);
/* N1: Map::AssociatePlanesByBoundary for one frame's planes (world coefficients) against map planes given in visiting
 * order (first n_seen seen, then not-seen), boundary clouds concatenated with offsets map_off[n_map + 1] */
void   orc_associate_planes(const float *plane_w, int n_planes, const float *map_w, const orc_point *map_bnd, const int64_t *map_off,
                            int n_seen, int n_map, float dis_th, float ang_th, float ver_th, float par_th,
                            int32_t *assoc, int32_t *vertical, int32_t *parallel, float *assoc_dist);

/* N1: pcl::transformPointCloud(cloud, out, Matrix4d) as MapPlane's constructor and MapPlane::UpdateBoundary call it (m row-major) */
void   orc_transform_cloud(const orc_point *pts, int n, const double m[16], orc_point *out);

#ifdef __cplusplus
}
#endif
#endif
