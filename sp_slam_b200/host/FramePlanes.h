// FramePlanes.h -- C++ host adapter above the C ABI (include/spx.h): the plane members of ORB_SLAM2::Frame and the two
// member functions that fill them, with the reference's names, so that Frame.cc only swaps the bodies.
//
// Reference interface mirrored here (all in /root/reference):
//   include/Frame.h:118-131   ComputePlanesFromOrganizedPointCloud / GeneratePlanesFromBoundries declarations
//   include/Frame.h:223-244   mvPlanePoints, mvBoundaryPoints, mvPlaneCoefficients, mnPlaneNum, mnRealPlaneNum, ...
//   src/Frame.cc:184-201      call sites in the RGB-D constructor + the Timer hooks around them
//
// PCL and OpenCV headers are not available in the build container, so by default the adapter fills layout-compatible
// stand-ins (32-byte PointXYZRGB, a 4x1 float "Mat").  With -DSPX_WITH_PCL it uses the real pcl:: / cv:: types
// (that variant cannot be compiled here).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/spx.h"

#if defined(__SSE2__) && !defined(SPX_HOST_NO_SSE2)
#include <emmintrin.h>
#include <xmmintrin.h>
#define SPX_HOST_SSE2 1
#endif

#ifdef SPX_WITH_PCL
#include <opencv2/core/core.hpp>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#endif

namespace spx_host {

#ifdef SPX_WITH_PCL
typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> PointCloud;
typedef cv::Mat CoefMat;
inline CoefMat make_coef(const float c[4]) { return (cv::Mat_<float>(4, 1) << c[0], c[1], c[2], c[3]); }
#else
// same layout as pcl::PointXYZRGB (PCL 1.8 point_types.hpp): xyz + pad | rgba + pad x3, 32 bytes, 16-byte aligned
struct alignas(16) PointT {
    float x, y, z, data_w;
    union { struct { uint8_t b, g, r, a; }; uint32_t rgba; float rgb; };
    uint32_t pad_[3];
};
struct PointCloud {   // the members of pcl::PointCloud<PointT> the reference touches
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
};
struct CoefMat {      // stand-in for the 4x1 CV_32F cv::Mat of mvPlaneCoefficients
    float v[4];
    template <typename T> T &at(int i) { return v[i]; }
    template <typename T> const T &at(int i) const { return v[i]; }
};
inline CoefMat make_coef(const float c[4]) { CoefMat m; std::memcpy(m.v, c, sizeof(m.v)); return m; }
#endif
static_assert(sizeof(PointT) == 32, "pcl::PointXYZRGB is 32 bytes: data[4] | rgba, pad x3");

// ---- rebuilding mvPlanePoints of the real planes from the compact results (include/spx.h, spx_compact_result) ----
// ExtractIndices(inputCloud, inliers[i]) copies inputCloud.points[idx] in list order (src/Frame.cc:907-928), and
// inputCloud.points[r * w + c] is (src/Frame.cc:857-870, m = r * Cloud.Dis, n = c * Cloud.Dis):
//     p.z = d;  p.x = (n - cx) * p.z / fx;  p.y = (m - cy) * p.z / fy;  p.r = 0; p.g = 0; p.b = 250;
// in fp32, one rounding per operation (no multiply-add exists in these expressions, so the result does not depend on
// -ffp-contract).  The library returns idx; this is the loop that turns it back into pcl::PointXYZRGB, reading d from the
// caller's own depth image -- the same bits the 16-byte path returns (tests/test_gpu_compact.py, tests/host/adapter_check.cpp).
struct DepthSource {
    const void *data = nullptr;     // first frame
    size_t pitch = 0;               // bytes per row
    size_t frame_stride = 0;        // bytes per frame
    bool u16 = false;               // CV_16U image, depth = float(d) * factor (Tracking::GrabImageRGBD, src/Tracking.cc:230-231)
    float factor = 1.0f;
};

class CloudExpander {
public:
    void Configure(float fx, float fy, float cx, float cy, int w, int h, int dis) {
        if (w == w_ && h == h_ && dis == dis_ && fx == fx_ && fy == fy_ && cx == cx_ && cy == cy_) return;
        fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy; w_ = w; h_ = h; dis_ = dis;
        xf_.resize(size_t(w)); yf_.resize(size_t(h)); rc_.resize(size_t(w) * size_t(h));
        for (int c = 0; c < w; ++c) xf_[size_t(c)] = float(c * dis) - cx;       // (n - cx): int - float
        for (int r = 0; r < h; ++r) yf_[size_t(r)] = float(r * dis) - cy;
        for (int r = 0; r < h; ++r) for (int c = 0; c < w; ++c) rc_[size_t(r) * size_t(w) + size_t(c)] = (uint32_t(r) << 16) | uint32_t(c);
    }
    // dst[k] = back-projection of organized pixel idx[k] of the frame whose image starts at `frame`.
    // SSE2 build: four points at a time (mulps / divps round exactly like mulss / divss), written with streaming stores --
    // a cloud is produced once and read much later, and write-combining halves the DRAM traffic of the 32-byte points.
    template <typename IDX>
    void Expand(PointT *dst, const IDX *idx, int n, const DepthSource &src, const char *frame) const {
        const size_t rstep = src.pitch * size_t(dis_), cstep = (src.u16 ? sizeof(uint16_t) : sizeof(float)) * size_t(dis_);
        const float fx = fx_, fy = fy_;
        const uint32_t *rc = rc_.data();
        const float *xf = xf_.data(), *yf = yf_.data();
        auto depth_at = [&](uint32_t r, uint32_t c) -> float {
            const char *px = frame + size_t(r) * rstep + size_t(c) * cstep;
            return src.u16 ? float(*reinterpret_cast<const uint16_t *>(px)) * src.factor : *reinterpret_cast<const float *>(px);
        };
        int k = 0;
#ifdef SPX_HOST_SSE2
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            const __m128 vfx = _mm_set1_ps(fx), vfy = _mm_set1_ps(fy), one = _mm_set1_ps(1.0f);
            const __m128 tail = _mm_castsi128_ps(_mm_set_epi32(0, 0, 0, int(0xff0000fau)));   // rgba | pad pad pad
            for (; k + 4 <= n; k += 4) {
                float z[4], a[4], b[4];
                const uint32_t i0 = idx[k];
                const uint32_t q0 = rc[i0];
                const uint32_t r0i = q0 >> 16, c0 = q0 & 0xffffu;
                __m128 va, vb;
                if (uint32_t(idx[k + 1]) == i0 + 1 && uint32_t(idx[k + 2]) == i0 + 2 && uint32_t(idx[k + 3]) == i0 + 3 && int(c0) + 3 < w_) {
                    // four neighbours of one row (inlier lists are mostly raster runs): one table look-up, contiguous factors
                    const char *px = frame + size_t(r0i) * rstep + size_t(c0) * cstep;
                    for (int j = 0; j < 4; ++j, px += cstep)
                        z[j] = src.u16 ? float(*reinterpret_cast<const uint16_t *>(px)) * src.factor : *reinterpret_cast<const float *>(px);
                    va = _mm_loadu_ps(xf + c0);
                    vb = _mm_set1_ps(yf[r0i]);
                } else {
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t q = rc[idx[k + j]];
                        const uint32_t r = q >> 16, c = q & 0xffffu;
                        z[j] = depth_at(r, c); a[j] = xf[c]; b[j] = yf[r];
                    }
                    va = _mm_loadu_ps(a); vb = _mm_loadu_ps(b);
                }
                const __m128 vz = _mm_loadu_ps(z);
                __m128 r0 = _mm_div_ps(_mm_mul_ps(va, vz), vfx);
                __m128 r1 = _mm_div_ps(_mm_mul_ps(vb, vz), vfy);
                __m128 r2 = vz, r3 = one;
                _MM_TRANSPOSE4_PS(r0, r1, r2, r3);                                           // rows: (x, y, z, 1) of the four points
                float *o = reinterpret_cast<float *>(dst + k);
                _mm_stream_ps(o, r0);      _mm_stream_ps(o + 4, tail);
                _mm_stream_ps(o + 8, r1);  _mm_stream_ps(o + 12, tail);
                _mm_stream_ps(o + 16, r2); _mm_stream_ps(o + 20, tail);
                _mm_stream_ps(o + 24, r3); _mm_stream_ps(o + 28, tail);
            }
        }
#endif
        for (; k < n; ++k) {
            const uint32_t q = rc[idx[k]];
            const uint32_t r = q >> 16, c = q & 0xffffu;
            const float z = depth_at(r, c);
            store(dst[k], xf[c] * z / fx, yf[r] * z / fy, z, 0xff0000fau);        // a = 255, (r, g, b) = (0, 0, 250)
        }
    }
    static void store(PointT &q, float x, float y, float z, uint32_t rgba) {
        float *o = reinterpret_cast<float *>(&q);
        o[0] = x; o[1] = y; o[2] = z; o[3] = 1.0f;               // data[3] = 1 (PointXYZRGB's constructor)
        uint32_t *u = reinterpret_cast<uint32_t *>(o + 4);
        u[0] = rgba; u[1] = u[2] = u[3] = 0u;
    }
    // 16-byte points of the C ABI -> pcl::PointXYZRGB
    static void copy(PointT *dst, const spx_point *src, int n) {
        int i = 0;
#ifdef SPX_HOST_SSE2
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            const __m128 keep = _mm_castsi128_ps(_mm_set_epi32(0, -1, -1, -1)), w1 = _mm_set_ps(1.0f, 0.f, 0.f, 0.f);
            for (; i < n; ++i) {
                const __m128 v = _mm_loadu_ps(reinterpret_cast<const float *>(src + i));
                float *o = reinterpret_cast<float *>(dst + i);
                _mm_stream_ps(o, _mm_or_ps(_mm_and_ps(v, keep), w1));
                _mm_stream_ps(o + 4, _mm_castsi128_ps(_mm_srli_si128(_mm_castps_si128(v), 12)));
            }
        }
#endif
        for (; i < n; ++i) store(dst[i], src[i].x, src[i].y, src[i].z, src[i].rgba);
    }
    // streaming stores are weakly ordered: call before handing the clouds to another thread
    static void fence() {
#ifdef SPX_HOST_SSE2
        _mm_sfence();
#endif
    }

private:
    float fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0;
    int w_ = 0, h_ = 0, dis_ = 0;
    std::vector<float> xf_, yf_;
    std::vector<uint32_t> rc_;     // organized pixel -> (row << 16 | column); rows, columns < 65536
};

// The plane fields of one Frame (include/Frame.h:223-244) and how a compact result fills them.  The clouds keep their storage
// between frames (a vector that shrinks hands its clouds to `spare_`), so a steady sequence allocates nothing.
struct PlaneFields {
    std::vector<PointCloud> mvPlanePoints;
    std::vector<PointCloud> mvBoundaryPoints;
    std::vector<CoefMat> mvPlaneCoefficients;
    int mnPlaneNum = 0, mnRealPlaneNum = 0;
    uint32_t flags = 0;

    // planes [lo, hi) of frame `f` of `res` are appended (lo = the frame's first plane starts over)
    void Fill(const spx_compact_result &res, int f, int lo, int hi, const CloudExpander &ex, const DepthSource &src) {
        const spx_frame_header &h = res.frames[f];
        const char *frame = static_cast<const char *>(src.data) + src.frame_stride * size_t(f);
        if (lo == 0) { mvPlaneCoefficients.clear(); }
        set_count(mvPlanePoints, spare_, size_t(hi));
        set_count(mvBoundaryPoints, spare_, size_t(hi));
        for (int k = lo; k < hi; ++k) {
            const spx_plane &p = res.planes[h.first_plane + k];
            mvPlaneCoefficients.push_back(make_coef(p.coef));
            PointCloud &pc = mvPlanePoints[size_t(k)], &bc = mvBoundaryPoints[size_t(k)];
            pc.points.resize(size_t(p.n_points));
            bc.points.resize(size_t(p.n_boundary));
            if (!p.is_supposed && res.index_width == 0) {         // 16-byte clouds (spx_extract_batch): a widening copy
                CloudExpander::copy(pc.points.data(), res.points + p.points_off, p.n_points);
                bc.width = 0; bc.height = 0;
            } else if (!p.is_supposed) {
                if (res.index_width == 2) ex.Expand(pc.points.data(), static_cast<const uint16_t *>(res.point_index) + p.points_off, p.n_points, src, frame);
                else ex.Expand(pc.points.data(), static_cast<const uint32_t *>(res.point_index) + p.points_off, p.n_points, src, frame);
                // `boundaryPoints->points = regions[i].getContour()` (src/Frame.cc:930-932) and GenerateBoundaryPoints' push_backs
                // (src/Frame.cc:1001-1011) leave width = height = 0
                bc.width = 0; bc.height = 0;
            } else {
                CloudExpander::copy(pc.points.data(), res.points + p.points_off, p.n_points);
                bc.width = uint32_t(p.n_boundary); bc.height = 1;     // a copy of linePoints, the output of ExtractIndices (src/Frame.cc:989)
            }
            pc.width = uint32_t(p.n_points); pc.height = 1;           // ExtractIndices::filter / PointCloud::operator+= (src/Frame.cc:988)
            pc.is_dense = true; bc.is_dense = true;
            CloudExpander::copy(bc.points.data(), res.boundary + p.boundary_off, p.n_boundary);
        }
        CloudExpander::fence();
        mnPlaneNum = hi;
        if (hi <= h.n_real || lo == 0) mnRealPlaneNum = hi < h.n_real ? hi : h.n_real;
        flags = h.flags;
    }

private:
    static void set_count(std::vector<PointCloud> &v, std::vector<PointCloud> &spare, size_t n) {
        while (v.size() > n) { spare.push_back(std::move(v.back())); v.pop_back(); }
        while (v.size() < n) {
            if (!spare.empty()) { v.push_back(std::move(spare.back())); spare.pop_back(); }
            else v.emplace_back();
        }
    }
    std::vector<PointCloud> spare_;
};

class FramePlanes : public PlaneFields {
public:
    // Timer::SetTPlane / SetTSPlane / AddPlane / AddSPlane arguments (src/Frame.cc:187-201)
    double tPlane = 0.0, tSPlane = 0.0;

    explicit FramePlanes(const spx_config &cfg) : cfg_(cfg) {
        if (spx_create(&cfg, &ctx_) != SPX_OK) throw std::runtime_error(std::string("spx_create: ") + spx_last_error(nullptr));
    }
    ~FramePlanes() { spx_destroy(ctx_); }
    FramePlanes(const FramePlanes &) = delete;
    FramePlanes &operator=(const FramePlanes &) = delete;

    // src/Frame.cc:186 -- imDepth: CV_32F metres, `step` = cv::Mat::step (bytes per row).
    // Runs the whole CUDA path once (compact results: the clouds of the real planes are rebuilt here from imDepth); the
    // supposed planes are kept back until GeneratePlanesFromBoundries.
    void ComputePlanesFromOrganizedPointCloud(const float *imDepth, int rows, int cols, size_t step) {
        mnPlaneNum = mnRealPlaneNum = 0;
        if (spx_extract_batch_compact(ctx_, imDepth, 1, rows, cols, step, step * size_t(rows), &res_) != SPX_OK)
            throw std::runtime_error(std::string("spx_extract_batch_compact: ") + spx_last_error(ctx_));
        spx_get_times(ctx_, &tPlane, &tSPlane);
        src_.data = imDepth; src_.pitch = step; src_.frame_stride = step * size_t(rows); src_.u16 = false; src_.factor = 1.0f;
        ex_.Configure(cfg_.fx, cfg_.fy, cfg_.cx, cfg_.cy, res_.cloud_width, res_.cloud_height, res_.cloud_dis);
        const spx_frame_header &h = res_.frames[0];
        Fill(res_, 0, 0, h.n_real, ex_, src_);     // Timer::AddPlane(mvPlaneCoefficients.size())  src/Frame.cc:187-188
        pending_ = true;
    }

    // src/Frame.cc:194 -- appends the supposed perpendicular planes of the frame processed by the previous call
    void GeneratePlanesFromBoundries(const float * /*imDepth*/ = nullptr) {
        if (!pending_) return;
        const spx_frame_header &h = res_.frames[0];
        Fill(res_, 0, h.n_real, h.n_planes, ex_, src_);   // mnPlaneNum = mvPlaneCoefficients.size()  src/Frame.cc:199
        pending_ = false;
    }

    spx_ctx *context() { return ctx_; }

    // Page-lock the buffers the depth images live in (e.g. the cv::Mat data a loader recycles): uploads then run at PCIe
    // rate and only the rows the organized cloud samples are copied (include/spx.h, "Host input").  Optional.
    static bool PinHostBuffer(void *ptr, size_t bytes) { return spx_host_register(ptr, bytes) == SPX_OK; }
    static void UnpinHostBuffer(void *ptr) { spx_host_unregister(ptr); }

private:
    spx_config cfg_;
    spx_ctx *ctx_ = nullptr;
    spx_compact_result res_{};
    DepthSource src_;
    CloudExpander ex_;
    bool pending_ = false;
};

inline std::vector<spx_point> pack_cloud(const PointCloud &c) {
    std::vector<spx_point> v(c.points.size());
    for (size_t i = 0; i < v.size(); ++i) { v[i].x = c.points[i].x; v[i].y = c.points[i].y; v[i].z = c.points[i].z; v[i].rgba = c.points[i].rgba; }
    return v;
}

// ---- N1: the plane-association part of ORB_SLAM2::Map (src/Map.cc:196-283,345-361) against a device copy of the map ----
// Map calls Upload() whenever a MapPlane is added or its boundary grows (src/Tracking.cc:434-441,1288-1291) with the
// planes in the order its loops visit them (mspMapPlanes first, then mspNotSeenMapPlanes), and Associate() where it
// called AssociatePlanesByBoundary; the returned indices address that order (-1 = pointer stays null).
class PlaneAssociator {
public:
    float mfDisTh = 0.2f, mfAngleTh = 0.8f, mfVerTh = 0.08716f, mfParTh = 0.9962f;   // Plane.Association* / Vertical / Parallel, src/Map.cc:30-37
    explicit PlaneAssociator(spx_ctx *ctx) {
        if (spx_map_create(ctx, &map_) != SPX_OK) throw std::runtime_error(std::string("spx_map_create: ") + spx_last_error(ctx));
        ctx_ = ctx;
    }
    ~PlaneAssociator() { spx_map_destroy(map_); }
    PlaneAssociator(const PlaneAssociator &) = delete;
    PlaneAssociator &operator=(const PlaneAssociator &) = delete;

    // worldPos[j]: MapPlane::GetWorldPos() (4x1), boundary[j]: MapPlane::mvBoundaryPoints (world frame)
    void Upload(const std::vector<CoefMat> &worldPos, const std::vector<const PointCloud *> &boundary, int nSeen) {
        std::vector<float> w(worldPos.size() * 4);
        std::vector<int64_t> off(worldPos.size() + 1, 0);
        std::vector<spx_point> pts;
        for (size_t j = 0; j < worldPos.size(); ++j) {
            for (int k = 0; k < 4; ++k) w[4 * j + size_t(k)] = worldPos[j].template at<float>(k);
            const std::vector<spx_point> b = pack_cloud(*boundary[j]);
            pts.insert(pts.end(), b.begin(), b.end());
            off[j + 1] = int64_t(pts.size());
        }
        if (spx_map_upload(map_, w.data(), pts.data(), off.data(), nSeen, int(worldPos.size())) != SPX_OK)
            throw std::runtime_error(std::string("spx_map_upload: ") + spx_last_error(ctx_));
    }

    // planeWorld[i]: Frame::ComputePlaneWorldCoeff(i) (src/Frame.cc:1146-1150).  Fills the indices of mvpMapPlanes[i],
    // mvpVerticalPlanes[i], mvpParallelPlanes[i]; returns pF.mbNewPlane (some plane found no association).
    bool AssociatePlanesByBoundary(const std::vector<CoefMat> &planeWorld, std::vector<int> &mapPlanes, std::vector<int> &verticalPlanes,
                                   std::vector<int> &parallelPlanes) {
        const int n = int(planeWorld.size());
        std::vector<float> w(size_t(n) * 4);
        for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) w[size_t(4 * i + k)] = planeWorld[size_t(i)].template at<float>(k);
        mapPlanes.assign(size_t(n), -1); verticalPlanes.assign(size_t(n), -1); parallelPlanes.assign(size_t(n), -1);
        if (n && spx_map_associate(map_, w.data(), n, mfDisTh, mfAngleTh, mfVerTh, mfParTh, mapPlanes.data(), verticalPlanes.data(),
                                   parallelPlanes.data(), nullptr) != SPX_OK)
            throw std::runtime_error(std::string("spx_map_associate: ") + spx_last_error(ctx_));
        bool newPlane = false;
        for (int v : mapPlanes) if (v < 0) newPlane = true;
        return newPlane;
    }

    // MapPlane::UpdateBoundary(pF, id) (src/MapPlane.cc:144-147) for map plane j (index in the uploaded order): its boundary
    // cloud becomes TwcMatrix (= T.inverse().matrix(), row-major 4x4 double) applied to mvBoundaryPoints[id] of the frame the
    // context processed last -- read where it lies on the device, nothing crosses PCIe.
    void UpdateBoundary(int j, const double TwcMatrix[16], int id, size_t boundarySize, int frameInBatch = 0) {
        if (spx_map_update_boundary_from_result(map_, j, TwcMatrix, frameInBatch, id, int(boundarySize)) != SPX_OK)
            throw std::runtime_error(std::string("spx_map_update_boundary_from_result: ") + spx_last_error(ctx_));
    }
    // the same from a host cloud (the MapPlane constructor with a KeyFrame's cloud, src/MapPlane.cc:25-31)
    void UpdateBoundary(int j, const double TwcMatrix[16], const PointCloud &cloud) {
        const std::vector<spx_point> pts = pack_cloud(cloud);
        if (spx_map_update_boundary(map_, j, TwcMatrix, pts.data(), int(pts.size())) != SPX_OK)
            throw std::runtime_error(std::string("spx_map_update_boundary: ") + spx_last_error(ctx_));
    }
    void SetWorldPos(int j, const CoefMat &Pos) {
        float w[4];
        for (int k = 0; k < 4; ++k) w[k] = Pos.template at<float>(k);
        if (spx_map_set_world_pos(map_, j, w) != SPX_OK) throw std::runtime_error(std::string("spx_map_set_world_pos: ") + spx_last_error(ctx_));
    }

private:
    spx_ctx *ctx_ = nullptr;
    spx_map *map_ = nullptr;
};

// ---- N4: the pcl::VoxelGrid<PointT> calls of the drawers (src/MapDrawer.cc:91-92,115-116; src/PointCloudMapping.cc:117-118) ----
class VoxelGrid {
public:
    explicit VoxelGrid(spx_ctx *ctx) : ctx_(ctx) {}
    void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
    void setInputCloud(const PointCloud *cloud) { in_ = cloud; }
    void filter(PointCloud &output) {
        const std::vector<spx_point> src = pack_cloud(*in_);
        std::vector<spx_point> dst(src.size() ? src.size() : 1);
        const int64_t off[2] = {0, int64_t(src.size())};
        int64_t out_off[2] = {0, 0};
        if (spx_voxel_grid(ctx_, src.data(), off, 1, leaf_, dst.data(), out_off) != SPX_OK)
            throw std::runtime_error(std::string("spx_voxel_grid: ") + spx_last_error(ctx_));
        output.points.resize(size_t(out_off[1]));
        for (size_t i = 0; i < output.points.size(); ++i) {
            PointT &q = output.points[i];
            q.x = dst[i].x; q.y = dst[i].y; q.z = dst[i].z; q.rgba = dst[i].rgba;
#ifndef SPX_WITH_PCL
            q.data_w = 1.0f;
#endif
        }
        output.width = uint32_t(output.points.size()); output.height = 1; output.is_dense = true;
    }

private:
    spx_ctx *ctx_;
    const PointCloud *in_ = nullptr;
    float leaf_[3] = {0.01f, 0.01f, 0.01f};
};

}  // namespace spx_host
