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

}  // namespace spx_host
