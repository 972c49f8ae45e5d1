// pcl_pin.cpp -- runs the REAL PCL (the reference pins 1.8.0: /root/reference/build.sh:4, README.md:25, CMakeLists.txt:42) on a
// raw float32 depth image with exactly the calls and parameters of SP-SLAM's plane extraction
// (/root/reference/src/Frame.cc:855-905 and :939-995) and dumps every intermediate as .npy files, so that the repository's
// CPU oracle (oracle/spx_oracle.cpp, a restatement of those PCL algorithms written without access to PCL) can be pinned:
//
//     cmake -S tools/pcl_pin -B /tmp/pcl_pin && cmake --build /tmp/pcl_pin
//     python tools/pcl_pin/make_inputs.py /tmp/pcl_in                        # the fixture frames as raw float32
//     for f in /tmp/pcl_in/*.bin; do /tmp/pcl_pin/pcl_pin $f 480 640 /tmp/pcl_out/$(basename $f .bin); done
//     python tools/pcl_pin/to_npz.py /tmp/pcl_in /tmp/pcl_out tests/golden    # -> tests/golden/pcl_<frame>.npz
//     python -m pytest tests/test_pcl_pin.py                                  # oracle vs PCL, stage by stage
//
// This program cannot be built in the repository's own container (no PCL / Eigen / Boost there); it is the recipe a person
// WITH PCL runs once.  Only PCL's outputs are dumped; SP-SLAM's own post-processing (PlaneNotSeen, LineInRange,
// IsBorderLine, CaculatePlanes) is plain code in the reference tree and needs no pinning.
//
// usage: pcl_pin depth.bin rows cols out_dir [fx fy cx cy] [Cloud.Dis Plane.MinSize Plane.AngleThreshold Plane.DistanceThreshold Line.Ratio Line.DistanceThreshold]
#include <pcl/ModelCoefficients.h>
#include <pcl/PointIndices.h>
#include <pcl/features/integral_image_normal.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/segmentation/organized_multi_plane_segmentation.h>
#include <pcl/segmentation/sac_segmentation.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

typedef pcl::PointXYZRGB PointT;
typedef pcl::PointCloud<PointT> PointCloud;

static void save_npy(const std::string &path, const char *descr, const std::vector<size_t> &shape, const void *data, size_t bytes) {
    std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < shape.size(); ++i) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
    dict += "), }";
    while ((10 + dict.size() + 1) % 64) dict += ' ';
    dict += '\n';
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { std::perror(path.c_str()); std::exit(4); }
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const uint16_t hl = uint16_t(dict.size());
    std::fwrite(magic, 1, 8, f); std::fwrite(&hl, 2, 1, f); std::fwrite(dict.data(), 1, dict.size(), f);
    if (bytes) std::fwrite(data, 1, bytes, f);
    std::fclose(f);
}
static void save_f32(const std::string &p, const std::vector<float> &v, std::vector<size_t> shape) { save_npy(p, "<f4", shape, v.data(), v.size() * 4); }
static void save_i32(const std::string &p, const std::vector<int32_t> &v, std::vector<size_t> shape) { save_npy(p, "<i4", shape, v.data(), v.size() * 4); }

// flattened list of index lists: values + offsets
static void save_lists(const std::string &stem, const std::vector<pcl::PointIndices> &lists) {
    std::vector<int32_t> val, off(1, 0);
    for (const pcl::PointIndices &l : lists) { val.insert(val.end(), l.indices.begin(), l.indices.end()); off.push_back(int32_t(val.size())); }
    save_i32(stem + "_val.npy", val, {val.size()});
    save_i32(stem + "_off.npy", off, {off.size()});
}
static void save_models(const std::string &stem, const std::vector<pcl::ModelCoefficients> &m) {
    std::vector<float> v;
    for (const pcl::ModelCoefficients &c : m) v.insert(v.end(), c.values.begin(), c.values.end());
    save_f32(stem + ".npy", v, {m.size(), size_t(4)});
}

int main(int argc, char **argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: pcl_pin depth.bin rows cols out_dir [fx fy cx cy] [dis minsize ang dist ratio linethr]\n"); return 2; }
    const int rows = std::atoi(argv[2]), cols = std::atoi(argv[3]);
    const std::string out = std::string(argv[4]) + "/";
    float fx = 517.306408f, fy = 516.469215f, cx = 318.643040f, cy = 255.313989f;     // Examples/RGB-D/TUM1.yaml:8-11
    if (argc >= 9) { fx = float(std::atof(argv[5])); fy = float(std::atof(argv[6])); cx = float(std::atof(argv[7])); cy = float(std::atof(argv[8])); }
    int cloudDis = 3, min_plane = 500; float AngTh = 3.0f, DisTh = 0.05f; double lineRatio = 0.2; float disTh = 0.01f;   // TUM1.yaml:73-76,99-100
    if (argc >= 15) { cloudDis = std::atoi(argv[9]); min_plane = std::atoi(argv[10]); AngTh = float(std::atof(argv[11])); DisTh = float(std::atof(argv[12]));
                      lineRatio = std::atof(argv[13]); disTh = float(std::atof(argv[14])); }
    std::vector<float> depth(size_t(rows) * cols);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(depth.data(), 4, depth.size(), f) != depth.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    std::fclose(f);

    // ---- src/Frame.cc:855-874: the organized cloud ----
    PointCloud::Ptr inputCloud(new PointCloud());
    for (int m = 0; m < rows; m += cloudDis)
        for (int n = 0; n < cols; n += cloudDis) {
            PointT p;
            p.z = depth[size_t(m) * cols + n];
            p.x = (n - cx) * p.z / fx;
            p.y = (m - cy) * p.z / fy;
            p.r = 0; p.g = 0; p.b = 250;
            inputCloud->points.push_back(p);
        }
    inputCloud->height = uint32_t(std::ceil(rows / float(cloudDis)));
    inputCloud->width = uint32_t(std::ceil(cols / float(cloudDis)));
    const size_t N = inputCloud->points.size();
    {
        std::vector<float> xyz(3 * N);
        std::vector<int32_t> rgba(N);
        for (size_t i = 0; i < N; ++i) { xyz[i] = inputCloud->points[i].x; xyz[N + i] = inputCloud->points[i].y; xyz[2 * N + i] = inputCloud->points[i].z; rgba[i] = int32_t(inputCloud->points[i].rgba); }
        save_f32(out + "cloud.npy", xyz, {size_t(3), N});
        save_i32(out + "cloud_rgba.npy", rgba, {N});
        save_i32(out + "dims.npy", {int32_t(inputCloud->width), int32_t(inputCloud->height)}, {size_t(2)});
    }

    // ---- src/Frame.cc:878-885: IntegralImageNormalEstimation ----
    pcl::IntegralImageNormalEstimation<PointT, pcl::Normal> ne;
    pcl::PointCloud<pcl::Normal>::Ptr cloud_normals(new pcl::PointCloud<pcl::Normal>);
    ne.setNormalEstimationMethod(ne.AVERAGE_3D_GRADIENT);
    ne.setMaxDepthChangeFactor(0.05f);
    ne.setNormalSmoothingSize(10.0f);
    ne.setInputCloud(inputCloud);
    ne.compute(*cloud_normals);
    {
        std::vector<float> nrm(4 * N);
        for (size_t i = 0; i < N; ++i) {
            const pcl::Normal &q = cloud_normals->points[i];
            nrm[i] = q.normal_x; nrm[N + i] = q.normal_y; nrm[2 * N + i] = q.normal_z; nrm[3 * N + i] = q.curvature;
        }
        save_f32(out + "normals.npy", nrm, {size_t(4), N});
    }

    // ---- src/Frame.cc:887-905: OrganizedMultiPlaneSegmentation ----
    auto configure = [&](pcl::OrganizedMultiPlaneSegmentation<PointT, pcl::Normal, pcl::Label> &mps) {
        mps.setMinInliers(min_plane);
        mps.setAngularThreshold(0.017453 * AngTh);
        mps.setDistanceThreshold(DisTh);
        mps.setInputNormals(cloud_normals);
        mps.setInputCloud(inputCloud);
    };
    {   // segment() alone: the state before refine() (raw connected-component labels, models, centroids, covariances)
        pcl::OrganizedMultiPlaneSegmentation<PointT, pcl::Normal, pcl::Label> mps;
        configure(mps);
        std::vector<pcl::ModelCoefficients> coefficients;
        std::vector<pcl::PointIndices> inliers, label_indices;
        std::vector<Eigen::Vector4f, Eigen::aligned_allocator<Eigen::Vector4f> > centroids;
        std::vector<Eigen::Matrix3f, Eigen::aligned_allocator<Eigen::Matrix3f> > covariances;
        pcl::PointCloud<pcl::Label> labels;
        mps.segment(coefficients, inliers, centroids, covariances, labels, label_indices);
        std::vector<int32_t> lab(N);
        for (size_t i = 0; i < N; ++i) lab[i] = int32_t(labels.points[i].label);
        save_i32(out + "seg_labels.npy", lab, {N});
        save_i32(out + "seg_n_label_lists.npy", {int32_t(label_indices.size())}, {size_t(1)});
        save_models(out + "seg_coef", coefficients);
        save_lists(out + "seg_inliers", inliers);
        std::vector<float> cen, cov;
        for (size_t i = 0; i < centroids.size(); ++i) {
            for (int k = 0; k < 4; ++k) cen.push_back(centroids[i][k]);
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov.push_back(covariances[i](r, c));
        }
        save_f32(out + "seg_centroids.npy", cen, {centroids.size(), size_t(4)});
        save_f32(out + "seg_cov.npy", cov, {covariances.size(), size_t(9)});
    }
    std::vector<pcl::PlanarRegion<PointT>, Eigen::aligned_allocator<pcl::PlanarRegion<PointT> > > regions;
    std::vector<pcl::ModelCoefficients> coefficients;
    std::vector<pcl::PointIndices> inliers;
    {
        pcl::PointCloud<pcl::Label>::Ptr labels(new pcl::PointCloud<pcl::Label>);
        std::vector<pcl::PointIndices> label_indices, boundary;
        pcl::OrganizedMultiPlaneSegmentation<PointT, pcl::Normal, pcl::Label> mps;
        configure(mps);
        mps.segmentAndRefine(regions, coefficients, inliers, labels, label_indices, boundary);
        std::vector<int32_t> lab(N);
        for (size_t i = 0; i < N; ++i) lab[i] = int32_t(labels->points[i].label);
        save_i32(out + "ref_labels.npy", lab, {N});
        save_models(out + "ref_coef", coefficients);
        save_lists(out + "ref_inliers", inliers);
        save_lists(out + "ref_boundary_idx", boundary);
        std::vector<float> cpts;
        std::vector<int32_t> coff(1, 0);
        for (size_t i = 0; i < regions.size(); ++i) {
            for (const PointT &p : regions[i].getContour()) { cpts.push_back(p.x); cpts.push_back(p.y); cpts.push_back(p.z); }
            coff.push_back(int32_t(cpts.size() / 3));
        }
        save_f32(out + "contour_pts.npy", cpts, {cpts.size() / 3, size_t(3)});
        save_i32(out + "contour_off.npy", coff, {coff.size()});
    }

    // ---- src/Frame.cc:939-995: SACSegmentation<LINE> on every contour, at most 4 rounds, inliers removed in between ----
    // (the reference only does this for planes that survive PlaneNotSeen and have >= 50 contour points; every region is run
    //  here and the test picks the ones it needs)
    {
        pcl::SACSegmentation<PointT> segLine;
        pcl::ExtractIndices<PointT> extract;
        segLine.setOptimizeCoefficients(true);
        segLine.setModelType(pcl::SACMODEL_LINE);
        segLine.setMaxIterations(1000);
        segLine.setDistanceThreshold(disTh);
        std::vector<float> lcoef;
        std::vector<int32_t> lrec;      // model, round, n_points, n_inliers
        std::vector<pcl::PointIndices> linl;
        for (size_t i = 0; i < regions.size(); ++i) {
            PointCloud::Ptr boundPoints(new PointCloud), tempPoints(new PointCloud);
            boundPoints->points = regions[i].getContour();
            const int boundSize = int(boundPoints->points.size());
            if (boundSize < 50) continue;
            for (int j = 0; j < 4; ++j) {
                pcl::PointIndices::Ptr lineins(new pcl::PointIndices());
                pcl::ModelCoefficients::Ptr coeffline(new pcl::ModelCoefficients());
                segLine.setInputCloud(boundPoints);
                segLine.segment(*lineins, *coeffline);
                lrec.push_back(int32_t(i)); lrec.push_back(j); lrec.push_back(int32_t(boundPoints->points.size())); lrec.push_back(int32_t(lineins->indices.size()));
                for (int k = 0; k < 6; ++k) lcoef.push_back(coeffline->values.size() == 6 ? coeffline->values[k] : 0.0f);
                linl.push_back(*lineins);
                if (lineins->indices.size() < lineRatio * boundSize) break;
                extract.setInputCloud(boundPoints);
                extract.setIndices(lineins);
                extract.setNegative(true);
                extract.filter(*tempPoints);
                boundPoints.swap(tempPoints);
            }
        }
        save_i32(out + "line_rec.npy", lrec, {lrec.size() / 4, size_t(4)});
        save_f32(out + "line_coef.npy", lcoef, {lcoef.size() / 6, size_t(6)});
        save_lists(out + "line_inliers", linl);
    }
    std::printf("%s: %zu points, %zu models\n", argv[1], N, coefficients.size());
    return 0;
}
