// Runs the C++ host adapter on a raw float32 depth file and prints the Frame fields, so the Python test can compare
// them with the C-ABI result:  adapter_check depth.bin rows cols
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sp_slam_b200/host/FramePlanes.h"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    std::vector<float> depth(size_t(rows) * cols);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(depth.data(), sizeof(float), depth.size(), f) != depth.size()) return 3;
    std::fclose(f);
    spx_config cfg;
    spx_default_config(&cfg);
    cfg.max_rows = rows; cfg.max_cols = cols;
    spx_host::FramePlanes fp(cfg);
    fp.ComputePlanesFromOrganizedPointCloud(depth.data(), rows, cols, size_t(cols) * sizeof(float));
    std::printf("real %d\n", fp.mnRealPlaneNum);
    fp.GeneratePlanesFromBoundries(depth.data());
    std::printf("all %d\n", fp.mnPlaneNum);
    for (int i = 0; i < fp.mnPlaneNum; ++i) {
        const auto &c = fp.mvPlaneCoefficients[size_t(i)];
        double sx = 0;
        for (const auto &p : fp.mvPlanePoints[size_t(i)].points) sx += double(p.x) + 2.0 * double(p.y) + 3.0 * double(p.z);
        std::printf("plane %d %.9g %.9g %.9g %.9g %zu %zu %.12g\n", i, c.at<float>(0), c.at<float>(1), c.at<float>(2), c.at<float>(3),
                    fp.mvPlanePoints[size_t(i)].points.size(), fp.mvBoundaryPoints[size_t(i)].points.size(), sx);
    }
    // N1 / N4 adapters: the frame's own planes as the map (identity pose), then the voxel grid of the first cloud
    if (fp.mnPlaneNum > 0) {
        spx_host::PlaneAssociator assoc(fp.context());
        std::vector<const spx_host::PointCloud *> bnd;
        for (const auto &b : fp.mvBoundaryPoints) bnd.push_back(&b);
        assoc.Upload(fp.mvPlaneCoefficients, bnd, fp.mnPlaneNum);
        std::vector<int> a, v, p;
        const bool newPlane = assoc.AssociatePlanesByBoundary(fp.mvPlaneCoefficients, a, v, p);
        std::printf("assoc new %d", newPlane ? 1 : 0);
        for (int i = 0; i < fp.mnPlaneNum; ++i) std::printf(" %d/%d/%d", a[size_t(i)], v[size_t(i)], p[size_t(i)]);
        std::printf("\n");
        spx_host::VoxelGrid voxel(fp.context());
        voxel.setLeafSize(0.05f, 0.05f, 0.05f);
        voxel.setInputCloud(&fp.mvPlanePoints[0]);
        spx_host::PointCloud tmp;
        voxel.filter(tmp);
        std::printf("voxel %zu -> %zu\n", fp.mvPlanePoints[0].points.size(), tmp.points.size());
    }
    return 0;
}
