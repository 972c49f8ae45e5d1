/*
 * spx.h -- C ABI of the B200-native plane-extraction front end (drop-in for SP-SLAM's PCL path).
 *
 * The reference has no FFI: the path is two Frame member functions that fill public Frame fields
 * (/root/reference/src/Frame.cc:186,194; fields at /root/reference/include/Frame.h:223-244).  Each entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; the shared library
 * (sp_slam_b200/libspx.so) is sm_100a CUDA with no CPU fallback: every call fails with SPX_ERR_CUDA when no
 * usable device is present.
 */
#ifndef SPX_H
#define SPX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_VERSION 200

/* per-frame capacities (a frame exceeding one sets SPX_FRAME_OVERFLOW in spx_frame_header.flags) */
#define SPX_MAX_CAND    96   /* connected components larger than Plane.MinSize                    */
#define SPX_MAX_MODELS  64   /* planes accepted by the curvature test                             */
#define SPX_MAX_PLANES  128  /* real + supposed planes (mvPlaneCoefficients.size())               */
#define SPX_MAX_LINES   4    /* line fits per real plane (src/Frame.cc:962)                       */

enum {
    SPX_OK = 0,
    SPX_ERR_ARG = 1,       /* bad argument (null pointer, size beyond the context's capacity, ...) */
    SPX_ERR_CUDA = 2,      /* CUDA runtime error or no device; see spx_last_error()                */
    SPX_ERR_STATE = 3      /* call order violated (e.g. results fetched before an extract)         */
};

enum {
    SPX_FRAME_OVERFLOW = 1u,    /* a per-frame capacity above was exceeded; the frame's plane list is truncated */
    SPX_FRAME_NONFINITE = 2u,   /* the depth image held NaN / Inf samples (such points stay unlabelled, as in PCL) */
    SPX_FRAME_SAT_UNPROVEN = 4u /* PCL's double-precision integral images could not be PROVEN free of rounding for this frame
                                   (depth values spanning more than ~2^28 in magnitude): the normals may then differ from PCL's
                                   in the last bit, because PCL's own result depends on its summation order.  Never set on the
                                   3 200 frames of the bench workloads (profiles/r2_sat_exact_sweep.json). */
};

/* Thresholds the reference reads from YAML through Config::Get (Examples/RGB-D/TUM1.yaml:73-78,99-100), the Frame
 * statics (src/Frame.cc:169-177,536-564) and the constants hard-coded in src/Frame.cc:881-882,945. */
typedef struct spx_config {
    int32_t cloud_dis;                 /* Cloud.Dis                 src/Frame.cc:856 */
    int32_t min_size;                  /* Plane.MinSize             src/Frame.cc:888 */
    float   angle_thr_deg;             /* Plane.AngleThreshold      src/Frame.cc:889 */
    float   dist_thr;                  /* Plane.DistanceThreshold   src/Frame.cc:890 */
    double  line_ratio;                /* Line.Ratio                src/Frame.cc:941 */
    float   line_dist_thr;             /* Line.DistanceThreshold    src/Frame.cc:942 */
    float   fx, fy, cx, cy;            /* Frame::fx..cy             src/Frame.cc:169-172 */
    float   min_x, max_x, min_y, max_y;/* Frame::mnMinX..mnMaxY     src/Frame.cc:536-564 */
    float   max_depth_change_factor;   /* 0.05f                     src/Frame.cc:881 */
    float   normal_smoothing_size;     /* 10.0f (only 10 is supported)  src/Frame.cc:882 */
    int32_t ransac_max_iter;           /* 1000                      src/Frame.cc:945 */
    int32_t enable_supposed;           /* 0 skips GeneratePlanesFromBoundries (src/Frame.cc:194) */
    /* capacity of the context */
    int32_t max_frames;                /* frames per batch */
    int32_t max_rows, max_cols;        /* depth image size */
    int32_t device;                    /* CUDA device ordinal */
    int32_t n_streams;                 /* internal streams a batch's frame groups run on; 0 = default (8) */
    int32_t normal_method;             /* ne.setNormalEstimationMethod (src/Frame.cc:880): 0 = AVERAGE_3D_GRADIENT (the reference's choice,
                                          default), 1 = COVARIANCE_MATRIX (9-channel integral image, per-pixel covariance + eigen33 +
                                          curvature; slower, for callers that want PCL's other method; spx_get_curvature taps it) */
} spx_config;

/* payload of pcl::PointXYZRGB: xyz + rgba packed as (a<<24 | r<<16 | g<<8 | b) */
typedef struct spx_point { float x, y, z; uint32_t rgba; } spx_point;

/* one entry of mvPlaneCoefficients / mvPlanePoints / mvBoundaryPoints (include/Frame.h:223-230) */
typedef struct spx_plane {
    float    coef[4];        /* unit normal + d, d >= 0 (src/Frame.cc:918-919,1087-1089) */
    int32_t  n_points;       /* mvPlanePoints[i].points.size()    */
    int32_t  n_boundary;     /* mvBoundaryPoints[i].points.size() */
    int64_t  points_off;     /* offset into spx_batch_result.points   */
    int64_t  boundary_off;   /* offset into spx_batch_result.boundary */
    int32_t  src;            /* real plane: index of the segmentation model; supposed plane: parent plane index */
    int32_t  is_supposed;
} spx_plane;

typedef struct spx_frame_header {
    int32_t  n_real;         /* mnRealPlaneNum (src/Frame.cc:187) */
    int32_t  n_planes;       /* mnPlaneNum     (src/Frame.cc:199) */
    int32_t  first_plane;    /* index of this frame's first entry in spx_batch_result.planes */
    uint32_t flags;
} spx_frame_header;

/* Results of the last extract call; all pointers are host memory owned by the context, valid until the next
 * extract / fetch / destroy on that context. */
typedef struct spx_batch_result {
    int32_t  n_frames;
    int32_t  n_planes_total;
    int64_t  n_points_total;
    int64_t  n_boundary_total;
    const spx_frame_header *frames;
    const spx_plane        *planes;
    const spx_point        *points;
    const spx_point        *boundary;
} spx_batch_result;

typedef struct spx_ctx spx_ctx;

/* fills the reference's defaults (TUM1.yaml) and capacity 1 frame of 480x640 on device 0 */
void spx_default_config(spx_config *cfg);

/* replaces: Config::SetParameterFile + the Frame statics set on the first frame (src/Frame.cc:162-177) */
int  spx_create(const spx_config *cfg, spx_ctx **out);
void spx_destroy(spx_ctx *ctx);
const char *spx_last_error(const spx_ctx *ctx);   /* ctx may be NULL: error of the last failed spx_create */

/* Use a caller-owned CUDA stream (cudaStream_t) for all work of this context; NULL restores the context's own. */
int spx_set_stream(spx_ctx *ctx, void *cuda_stream);

/* replaces: Frame::ComputePlanesFromOrganizedPointCloud + Frame::GeneratePlanesFromBoundries for one frame
 * (src/Frame.cc:186-201).  depth: HOST memory, CV_32F metres, `rows` x `cols`, row pitch in bytes.  Host->device
 * copy, kernels and device->host copy of the results all happen inside the call. */
int spx_extract(spx_ctx *ctx, const float *depth, int rows, int cols, size_t pitch_bytes, spx_batch_result *out);

/* the same for `n_frames` frames of one sequence (offline processing, frames are independent);
 * frame f starts at (const char*)depth + f * frame_stride_bytes */
int spx_extract_batch(spx_ctx *ctx, const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                      size_t frame_stride_bytes, spx_batch_result *out);

/* Host input, how the depth gets to the device.  The organized cloud samples every Cloud.Dis-th row and column
 * (src/Frame.cc:857-872); full-resolution depth is read only by IsBorderPoint's 21x21 windows around line points
 * (src/Frame.cc:1038-1052).  When the caller's image is page-locked (cudaHostAlloc, cudaHostRegister or
 * spx_host_register below) the library uploads only the sampled rows (1/Cloud.Dis of the bytes, one strided copy) and the
 * window sectors the border tests will read are fetched from the image over PCIe by a kernel, each once; a pageable
 * image is uploaded whole.  Large float batches (64 frames or more and at least 200 MB of sampled rows, i.e. calls that
 * are bound by the upload) use a second route beside it: a pool of host threads
 * inside the library copies the h x w samples themselves (1/Cloud.Dis^2 of the image; the copy engine can skip rows but
 * not columns) of the LAST frame groups of the batch into a page-locked staging buffer while the copy engine moves the
 * sampled rows of the first groups, and only those samples are uploaded for the gathered groups (138 instead of 410 KB
 * per frame at 640x480, Cloud.Dis 3).  Results are identical.
 * mode 0 = automatic (default), 1 = always upload the whole image, 2 = sampled rows whenever the image is page-locked,
 * 3 = every group gathered whenever the image is page-locked and float. */
int spx_set_upload_mode(spx_ctx *ctx, int mode);
/* host threads of the gathered route: 0 (default) = half the hardware threads, at most 16; also SPX_GATHER_THREADS */
int spx_set_gather_threads(spx_ctx *ctx, int n_threads);
/* automatic mode: the share of a batch's frames (its tail) that takes the gathered route, 0 .. 1; negative (default) =
 * chosen from the thread count t as t / (t + 4), where the copy engine and the host threads finish together on the
 * machines measured; also SPX_GATHER_SHARE */
int spx_set_gather_share(spx_ctx *ctx, double share);
/* the gather step on its own (host code, no device): the organized cloud's samples depth[f][m * dis][n * dis]
 * (src/Frame.cc:857-872) of `n_frames` frames, as the upload path of mode 3 stages them -- out[f][m][n] with rows of
 * `out_row_floats` floats (>= the cloud width; the tail is zeroed), the frames cut into `n_groups` groups and handed to
 * `n_threads` threads.  For tests and for hosts that want to stage the samples themselves. */
int spx_host_gather_samples(const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes, size_t frame_stride_bytes,
                            int cloud_dis, int n_groups, int n_threads, float *out, size_t out_row_floats);
/* page-lock / release a caller-owned host buffer (e.g. the cv::Mat data of the depth images a loader recycles) */
int spx_host_register(void *ptr, size_t bytes);
int spx_host_unregister(void *ptr);
/* bytes moved by the last host-input extract: uploaded by copies, fetched from the page-locked image by k_border_fetch
 * (the border tests' window sectors), results copied back */
int spx_get_transfer_bytes(const spx_ctx *ctx, unsigned long long *h2d_copied, unsigned long long *h2d_in_place,
                           unsigned long long *d2h);

/* The same from the raw 16-bit depth image (CV_16U, e.g. a TUM PNG): replaces, in addition, the conversion
 * imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) of Tracking::GrabImageRGBD (src/Tracking.cc:230-231; the factor
 * is 1.0f / DepthMapFactor, src/Tracking.cc:142-146).  Half the upload; depth = float(d) * depth_map_factor on the device. */
int spx_extract_batch_u16(spx_ctx *ctx, const uint16_t *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                          size_t frame_stride_bytes, float depth_map_factor, spx_batch_result *out);

/* device-resident variant: `depth_dev` is DEVICE memory; only kernels run (asynchronously on the context's stream),
 * results stay on the device until spx_fetch_results (which synchronises and copies them to the host). */
int spx_extract_batch_device(spx_ctx *ctx, const float *depth_dev, int n_frames, int rows, int cols,
                             size_t pitch_bytes, size_t frame_stride_bytes);
int spx_fetch_results(spx_ctx *ctx, spx_batch_result *out);
/* only the frame headers and plane records (coefficients and counts), not the clouds */
int spx_fetch_planes(spx_ctx *ctx, spx_batch_result *out);

/* ---- compact results: ordered inlier index lists instead of the real planes' point clouds ----
 * mvPlanePoints[i] of a REAL plane is ExtractIndices(inputCloud, inliers[i]) (src/Frame.cc:907-928): point k is the
 * back-projection (src/Frame.cc:861-868) of organized pixel inliers[i].indices[k], colour (0, 0, 250) -- a pure function of
 * (index, depth image, intrinsics).  95 % of the result bytes of a frame are these clouds, so the compact calls return
 * inlier_indices itself (2 bytes per point when the organized cloud has at most 65536 points, else 4) and the host adapter
 * (sp_slam_b200/host/FramePlanes.h) rebuilds pcl::PointXYZRGB from the caller's depth image with the reference's fp32
 * expression; bit-identical to the 16-byte path (tests/test_gpu_compact.py).  Boundary clouds (contours) and the supposed
 * planes' clouds (50x50 grid + line inliers) come back as 16-byte points, as in spx_batch_result.
 *   real plane:     points_off indexes point_index (n_points entries, inlier_indices order)
 *   supposed plane: points_off indexes points
 *   any plane:      boundary_off indexes boundary
 * organized pixel q = r * cloud_width + c samples depth(r * cloud_dis, c * cloud_dis). */
typedef struct spx_compact_result {
    int32_t  n_frames;
    int32_t  n_planes_total;
    int32_t  index_width;            /* bytes per entry of point_index: 2 or 4 */
    int32_t  cloud_width, cloud_height, cloud_dis;
    int64_t  n_index_total;          /* entries of point_index (real planes' inliers) */
    int64_t  n_points_total;         /* entries of points (supposed planes only)      */
    int64_t  n_boundary_total;
    const spx_frame_header *frames;
    const spx_plane        *planes;
    const void             *point_index;
    const spx_point        *points;
    const spx_point        *boundary;
} spx_compact_result;
/* same contract as spx_extract_batch / spx_extract_batch_u16 (host depth in, results in host memory owned by the context) */
int spx_extract_batch_compact(spx_ctx *ctx, const float *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                              size_t frame_stride_bytes, spx_compact_result *out);
int spx_extract_batch_u16_compact(spx_ctx *ctx, const uint16_t *depth, int n_frames, int rows, int cols, size_t pitch_bytes,
                                  size_t frame_stride_bytes, float depth_map_factor, spx_compact_result *out);
/* what spx_extract_batch_device packs: 0 = point clouds (default; spx_fetch_results), 1 = compact (spx_fetch_compact) */
int spx_set_result_mode(spx_ctx *ctx, int mode);
int spx_fetch_compact(spx_ctx *ctx, spx_compact_result *out);
/* Streaming delivery for the host-input calls: a batch runs as frame groups; `fn` is called on the calling thread, inside
 * spx_extract_batch* (compact or not), as soon as the results of frames [frame0, frame1) are final in host memory (offsets
 * already in host layout), while later groups are still on the device -- the adapter starts filling Frame fields then.
 * Under the 16-byte-cloud calls the view has index_width = 0, point_index = NULL and EVERY plane's points_off indexes
 * `points`.  `view` and the arrays it points to stay valid until the next extract call on the context.  fn = NULL turns it off. */
typedef void (*spx_group_fn)(void *user, int frame0, int frame1, const spx_compact_result *view);
int spx_set_group_callback(spx_ctx *ctx, spx_group_fn fn, void *user);

/* Device-side view of the last extract's results (valid until the next extract on the context; read them on the
 * context's stream or after synchronising it).  This is what a multi-GPU caller hands to NCCL for the end-of-sequence
 * gather of plane lists without a host round trip. */
typedef struct spx_device_result {
    int32_t  n_frames;
    const spx_frame_header *frames;    /* n_frames */
    const spx_plane        *planes;    /* totals[0] records */
    const spx_point        *points;    /* totals[1] */
    const spx_point        *boundary;  /* totals[2] */
    const long long        *totals;    /* 3 values: planes, points, boundary points of the batch */
    int64_t  planes_capacity;          /* records the planes buffer can hold */
} spx_device_result;
int spx_get_device_results(spx_ctx *ctx, spx_device_result *out);

/* "feed the reference's normals": like spx_extract for one frame, but integral-image normal estimation is replaced by
 * the caller's normals (HOST memory, 3*N floats: nx[N], ny[N], nz[N], N = organized cloud size, NaN = invalid). */
int spx_segment_from_normals(spx_ctx *ctx, const float *depth, int rows, int cols, size_t pitch_bytes,
                             const float *normals, spx_batch_result *out);

/* organized-cloud size for an image: width = ceil(cols / Cloud.Dis), height = ceil(rows / Cloud.Dis) (src/Frame.cc:873-874) */
int spx_cloud_dims(const spx_ctx *ctx, int rows, int cols, int *width, int *height);

/* replaces: Timer::SetTPlane / SetTSPlane (src/Frame.cc:184-197): device time in seconds of the two sections of the
 * last extract call (whole batch), measured with CUDA events on the context's stream. */
int spx_get_times(spx_ctx *ctx, double *t_plane, double *t_splane);
/* number of kernel launches issued by the last extract call */
int spx_last_launch_count(const spx_ctx *ctx);
/* per-kernel device times of the last extract call: with profiling on, every launch is bracketed by CUDA events on
 * the stream it is issued to (a batch is cut into frame groups that run on internal streams, so one kernel appears
 * once per group and the launches of different groups overlap).  names[k] (static strings) / ms[k] for launch k,
 * k < min(*n, cap). */
int spx_set_profile(spx_ctx *ctx, int on);
int spx_get_kernel_times(spx_ctx *ctx, const char **names, float *ms, int cap, int *n);
/* the same launches as a timeline: start / end of launch k in ms after the beginning of the extract call */
int spx_get_kernel_timeline(spx_ctx *ctx, const char **names, float *start_ms, float *end_ms, int cap, int *n);

/* host-input calls (spx_extract_batch / _u16): when each frame group of the last call reached five points, in ms after
 * the call began: [7g+0] group started, [7g+1] its depth is on the device, [7g+2] real planes final (post-filter),
 * [7g+3] its last kernel ended, [7g+4] its results are on the host (device clock); [7g+5] the host began to enqueue the
 * group, [7g+6] the host saw the group's totals (host clock) */
int spx_get_group_timeline(spx_ctx *ctx, float *t_ms, int cap_groups, int *n_groups);

/* ---- debug taps for the parity tests: intermediates of frame `frame` of the last extract call, copied to host.
 * They need spx_set_debug(ctx, 1) BEFORE the extract call (it adds the per-pixel label kernel to the schedule). ---- */
int spx_set_debug(spx_ctx *ctx, int on);
int spx_get_cloud(spx_ctx *ctx, int frame, float *x, float *y, float *z);                 /* N each */
int spx_get_distance_map(spx_ctx *ctx, int frame, float *dist);                           /* N, min(PCL distance map, 10) */
int spx_get_normals(spx_ctx *ctx, int frame, float *nx, float *ny, float *nz, float *plane_d);
int spx_get_curvature(spx_ctx *ctx, int frame, float *curvature);    /* N, pcl::Normal::curvature (COVARIANCE_MATRIX method only) */
int spx_get_labels_raw(spx_ctx *ctx, int frame, uint32_t *labels, int *n_label_lists);    /* CCL labels before refine */
int spx_get_plane_ids(spx_ctx *ctx, int frame, int8_t *ids);     /* after refine: model index per pixel, -1 = none */

typedef struct spx_model_info {
    float    coef[4];          /* OrganizedMultiPlaneSegmentation model_coefficients (sign as PCL leaves it) */
    float    centroid[3];
    float    cov[9];
    float    curvature;
    uint32_t label;            /* CCL label of the component */
    int32_t  n_segment;        /* inliers after segment() */
    int32_t  n_inliers;        /* inliers after refine()  */
    int32_t  n_contour;
} spx_model_info;
int spx_get_models(spx_ctx *ctx, int frame, spx_model_info *models /* SPX_MAX_MODELS */, int *n_models);
int spx_get_model_inliers(spx_ctx *ctx, int frame, int model, int32_t *idx /* n_inliers */);
int spx_get_model_contour(spx_ctx *ctx, int frame, int model, int32_t *idx /* n_contour */);

typedef struct spx_line_info {
    int32_t plane, round, n_points, iterations, n_inliers, in_range, is_border, emitted;
    float   coef[6];
} spx_line_info;
/* one record per SACSegmentation::segment call of GeneratePlanesFromBoundries, in the reference's call order */
int spx_get_lines(spx_ctx *ctx, int frame, spx_line_info *lines /* SPX_MAX_MODELS*SPX_MAX_LINES */, int *n_lines);

/* ================= the steps either side of the path ================= */

/* N4.  replaces: pcl::VoxelGrid<pcl::PointXYZRGB> voxel; voxel.setLeafSize(lx, ly, lz); voxel.setInputCloud(c); voxel.filter(out)
 * (src/MapDrawer.cc:91-92,115-116, src/PointCloudMapping.cc:117-118,172-173; dead path src/Frame.cc:810-814) for
 * `n_clouds` independent clouds in one call.  Cloud s = points[cloud_off[s] .. cloud_off[s+1]) (HOST memory); its
 * downsampled points land in out[out_off[s] .. out_off[s+1]) in PCL's order (ascending voxel index); `out` must hold
 * as many points as the input, `out_off` n_clouds + 1 entries.  A cloud whose voxel indices would overflow an int comes
 * back unchanged, as PCL does after its "leaf size is too small" warning. */
int spx_voxel_grid(spx_ctx *ctx, const spx_point *points, const int64_t *cloud_off, int n_clouds, const float leaf[3],
                   spx_point *out, int64_t *out_off);
/* the same on the device for the clouds of the last spx_extract_batch_device, before spx_fetch_results:
 * which = 0 every plane's mvPlanePoints, which = 1 every plane's mvBoundaryPoints (the contours); the plane records'
 * counts / offsets and the totals are rewritten.  Off by default: the reference stores both raw (src/Frame.cc:925-932). */
int spx_voxel_downsample_results(spx_ctx *ctx, float leaf, int which);

/* N1.  replaces: Map::AssociatePlanesByBoundary + Map::PointDistanceFromPlane (src/Map.cc:196-283,345-361) for the planes
 * of one frame against a device-resident copy of the map planes.  spx_map_upload mirrors the map whenever it changes
 * (new MapPlane, MapPlane::UpdateBoundary, src/Tracking.cc:434-441,1288-1291): world coefficients (4 per plane) and
 * world-frame boundary clouds, map planes in the VISITING order of the reference's loops -- the first n_seen are
 * mspMapPlanes, the rest mspNotSeenMapPlanes (the reference iterates std::set<MapPlane*>, i.e. in pointer order).
 * spx_map_associate takes the frame planes' world coefficients (Frame::ComputePlaneWorldCoeff, src/Frame.cc:1146-1150)
 * and the thresholds Map reads from YAML (Plane.AssociationDisRef, AssociationAngRef, VerticalThreshold,
 * ParallelThreshold; src/Map.cc:30-37) and returns, per frame plane, the index of the associated / vertical / parallel
 * map plane in that order (or -1: mvpMapPlanes[i] / mvpVerticalPlanes[i] / mvpParallelPlanes[i] stay null) and the
 * final ldTh. */
/* A map belongs to the context it was created on (same device, same stream, same single-thread rule) and must be
 * destroyed before that context. */
typedef struct spx_map spx_map;
int  spx_map_create(spx_ctx *ctx, spx_map **out);
void spx_map_destroy(spx_map *map);
int  spx_map_upload(spx_map *map, const float *map_w, const spx_point *boundary, const int64_t *boundary_off, int n_seen, int n_map);
int  spx_map_associate(spx_map *map, const float *plane_w, int n_planes, float dis_th, float ang_th, float ver_th, float par_th,
                       int32_t *assoc, int32_t *vertical, int32_t *parallel, float *assoc_dist);
/* replaces: MapPlane::UpdateBoundary(pF, id) and the MapPlane constructor's boundary cloud (src/MapPlane.cc:25-31,144-147):
 * pcl::transformPointCloud(cloud, *mvBoundaryPoints, T.inverse().matrix()) -- map plane `j`'s boundary cloud BECOMES the
 * transformed cloud (the reference overwrites, it does not append).  transform: the caller's T.inverse().matrix(),
 * row-major 4x4 double; x' = float(m00 x + m01 y + m02 z + m03) evaluated in double as PCL 1.8.0 does.  `cloud`: HOST points. */
int  spx_map_update_boundary(spx_map *map, int j, const double transform[16], const spx_point *cloud, int n);
/* the same with the cloud read on the device: the boundary of plane `plane` of frame `frame` of the last extract call on the
 * map's context (either path); n_boundary = mvBoundaryPoints[plane].size(), checked against the device record.  This is the
 * per-frame call of Tracking (src/Tracking.cc:434-441): no cloud crosses PCIe. */
int  spx_map_update_boundary_from_result(spx_map *map, int j, const double transform[16], int frame, int plane, int n_boundary);
/* MapPlane::SetWorldPos: new world coefficients of map plane j */
int  spx_map_set_world_pos(spx_map *map, int j, const float coef_w[4]);
/* map plane j's boundary cloud as stored on the device (parity tap / MapPlane::mvBoundaryPoints for the drawers) */
int  spx_map_get_boundary(spx_map *map, int j, spx_point *out, int cap, int *n);

/* N3.  replaces: the plane part of Optimizer::PoseOptimization (src/Optimizer.cc:519-1160): pose-only Levenberg-Marquardt
 * with the unary plane edges of g2oAddition (EdgePlane, EdgeParallelPlane, EdgeVerticalPlane; Plane3D's azimuth /
 * elevation / distance parametrisation), Huber kernels and the four outlier rounds.  HOST code (a 6x6 system a few
 * times per frame), double precision; no CUDA context involved.  Tcw: row-major 4x4, in = Frame::mTcw, out = the
 * optimised pose.  One record per edge, as PoseOptimization builds them (src/Optimizer.cc:695-907):
 *   kind 0 EdgePlane          info = (angleInfo, angleInfo, disInfo) [x2 for a not-seen map plane], delta = sqrt(Plane.Chi),   chi2_max = Plane.Chi
 *   kind 1 EdgeParallelPlane  info = (parInfo, parInfo),                                           delta = sqrt(Plane.VPChi), chi2_max = Plane.VPChi
 *   kind 2 EdgeVerticalPlane  info = (verInfo, verInfo),                                           delta = sqrt(Plane.VPChi), chi2_max = Plane.VPChi
 * with angleInfo = 3282.8 / Plane.AngleInfo^2, disInfo = Plane.DistanceInfo^2, parInfo / verInfo = 3282.8 / Plane.ParallelInfo^2 / VerticalInfo^2.
 * outlier[i] (optional) = mvbPlaneOutlier / mvbParPlaneOutlier / mvbVerPlaneOutlier of edge i; chi2[i] (optional) = its last chi2;
 * *n_bad = nBad of the last round. */
typedef struct spx_plane_edge {
    int32_t kind;
    int32_t reserved;
    float   plane_w[4];        /* MapPlane::GetWorldPos() of the associated / parallel / vertical map plane */
    float   measurement[4];    /* mvPlaneCoefficients[i] */
    double  info[3];
    double  huber_delta;
    double  chi2_max;
} spx_plane_edge;
int spx_pose_optimize_planes(double Tcw[16], const spx_plane_edge *edges, int n_edges, int rounds, int iterations,
                             uint8_t *outlier, double *chi2, int *n_bad);
/* parity tap: the residuals of the edges at the pose Tcw (computeError of each edge; 3 doubles per edge, the third is 0 for the 2-d edges) */
int spx_plane_edge_errors(const double Tcw[16], const spx_plane_edge *edges, int n_edges, double *errors);

#ifdef __cplusplus
}
#endif
#endif
