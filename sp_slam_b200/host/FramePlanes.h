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
static_assert(sizeof(PointT) == 32, "pcl::PointXYZRGB is 32 bytes");
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

class FramePlanes {
public:
    // ---- the Frame members (include/Frame.h:223-244) ----
    std::vector<PointCloud> mvPlanePoints;
    std::vector<PointCloud> mvBoundaryPoints;
    std::vector<CoefMat> mvPlaneCoefficients;
    int mnPlaneNum = 0, mnRealPlaneNum = 0;
    // Timer::SetTPlane / SetTSPlane / AddPlane / AddSPlane arguments (src/Frame.cc:187-201)
    double tPlane = 0.0, tSPlane = 0.0;
    uint32_t flags = 0;

    explicit FramePlanes(const spx_config &cfg) {
        if (spx_create(&cfg, &ctx_) != SPX_OK) throw std::runtime_error(std::string("spx_create: ") + spx_last_error(nullptr));
    }
    ~FramePlanes() { spx_destroy(ctx_); }
    FramePlanes(const FramePlanes &) = delete;
    FramePlanes &operator=(const FramePlanes &) = delete;

    // src/Frame.cc:186 -- imDepth: CV_32F metres, `step` = cv::Mat::step (bytes per row).
    // Runs the whole CUDA path once; the supposed planes are kept back until GeneratePlanesFromBoundries.
    void ComputePlanesFromOrganizedPointCloud(const float *imDepth, int rows, int cols, size_t step) {
        mvPlanePoints.clear(); mvBoundaryPoints.clear(); mvPlaneCoefficients.clear();
        mnPlaneNum = mnRealPlaneNum = 0;
        if (spx_extract(ctx_, imDepth, rows, cols, step, &res_) != SPX_OK)
            throw std::runtime_error(std::string("spx_extract: ") + spx_last_error(ctx_));
        spx_get_times(ctx_, &tPlane, &tSPlane);
        const spx_frame_header &h = res_.frames[0];
        flags = h.flags;
        append(h.first_plane, h.first_plane + h.n_real);
        mnRealPlaneNum = h.n_real;          // Timer::AddPlane(mvPlaneCoefficients.size())  src/Frame.cc:187-188
        mnPlaneNum = h.n_real;
        pending_ = true;
    }

    // src/Frame.cc:194 -- appends the supposed perpendicular planes of the frame processed by the previous call
    void GeneratePlanesFromBoundries(const float * /*imDepth*/ = nullptr) {
        if (!pending_) return;
        const spx_frame_header &h = res_.frames[0];
        append(h.first_plane + h.n_real, h.first_plane + h.n_planes);
        mnPlaneNum = h.n_planes;            // src/Frame.cc:199
        pending_ = false;
    }

    spx_ctx *context() { return ctx_; }

    // Page-lock the buffers the depth images live in (e.g. the cv::Mat data a loader recycles): uploads then run at PCIe
    // rate and only the rows the organized cloud samples are copied (include/spx.h, "Host input").  Optional.
    static bool PinHostBuffer(void *ptr, size_t bytes) { return spx_host_register(ptr, bytes) == SPX_OK; }
    static void UnpinHostBuffer(void *ptr) { spx_host_unregister(ptr); }

private:
    void append(int lo, int hi) {
        for (int k = lo; k < hi; ++k) {
            const spx_plane &p = res_.planes[k];
            mvPlaneCoefficients.push_back(make_coef(p.coef));
            mvPlanePoints.push_back(cloud_of(res_.points + p.points_off, p.n_points));
            mvBoundaryPoints.push_back(cloud_of(res_.boundary + p.boundary_off, p.n_boundary));
        }
    }
    static PointCloud cloud_of(const spx_point *src, int n) {
        PointCloud c;
        c.points.resize(size_t(n));
        for (int i = 0; i < n; ++i) {
            PointT &q = c.points[size_t(i)];
            q.x = src[i].x; q.y = src[i].y; q.z = src[i].z;
#ifndef SPX_WITH_PCL
            q.data_w = 1.0f;
#endif
            q.rgba = src[i].rgba;
        }
        c.width = uint32_t(n); c.height = 1; c.is_dense = true;
        return c;
    }
    spx_ctx *ctx_ = nullptr;
    spx_batch_result res_{};
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
