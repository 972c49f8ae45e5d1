// SequencePlanes.h -- offline sequence processing above the C ABI: the plane fields of EVERY frame of a depth sequence
// (BASELINE configs[1]: "1000-frame sequence, batched plane extraction"), filled as Frame's constructor would fill them
// one frame at a time (/root/reference/src/Frame.cc:186-201, fields include/Frame.h:223-244).
//
// One spx_extract_batch_compact call per batch.  The library cuts the batch into frame groups and calls back as soon as
// a group's results are in host memory (spx_set_group_callback); a pool of host threads then rebuilds the groups' clouds
// (CloudExpander, FramePlanes.h) while the later groups are still on the device, so the host work overlaps the device
// work and the transfers.  The clouds keep their storage between batches (PlaneFields): a steady sequence allocates nothing.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "FramePlanes.h"

namespace spx_host {

class SequencePlanes {
public:
    std::vector<PlaneFields> frames;    // frames[f]: the fields of frame f of the last Process call

    explicit SequencePlanes(const spx_config &cfg, int n_threads = 0) : cfg_(cfg) {
        if (spx_create(&cfg, &ctx_) != SPX_OK) throw std::runtime_error(std::string("spx_create: ") + spx_last_error(nullptr));
        spx_set_group_callback(ctx_, &SequencePlanes::on_group, this);
        if (n_threads <= 0) n_threads = int(std::thread::hardware_concurrency()) - 1;   // the calling thread drives the device
        if (n_threads < 1) n_threads = 1;
        // the library's gathered upload route would run on host threads of its own; here the workers already load the host's cores and
        // memory system (they write ~0.9 GB of clouds per 1000 frames), so the depth goes up as sampled rows through the copy engine
        spx_set_upload_mode(ctx_, 2);
        for (int t = 0; t < n_threads; ++t) workers_.emplace_back([this] { work(); });
    }
    ~SequencePlanes() {
        { std::lock_guard<std::mutex> lk(mu_); quit_ = true; }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
        spx_destroy(ctx_);
    }
    SequencePlanes(const SequencePlanes &) = delete;
    SequencePlanes &operator=(const SequencePlanes &) = delete;

    // depth: CV_32F metres, frame f at depth + f * frame_stride bytes, `step` bytes per row.
    // compact = true: the real planes' clouds cross PCIe as index lists and are rebuilt here (fewer bytes, more host work);
    // false: they arrive as 16-byte points and are only widened to pcl::PointXYZRGB (what pays on a host with few cores per GPU
    // is measured by bench.py: e2e_adapter / e2e_adapter_clouds).
    void Process(const float *depth, int n_frames, int rows, int cols, size_t step, size_t frame_stride, bool compact = true) {
        begin(depth, n_frames, step, frame_stride, false, 1.0f);
        int rc;
        if (compact) { spx_compact_result res; rc = spx_extract_batch_compact(ctx_, depth, n_frames, rows, cols, step, frame_stride, &res); }
        else { spx_batch_result res; rc = spx_extract_batch(ctx_, depth, n_frames, rows, cols, step, frame_stride, &res); }
        finish(rc);
    }
    // raw CV_16U images + mDepthMapFactor (Tracking::GrabImageRGBD's convertTo, src/Tracking.cc:230-231)
    void ProcessU16(const uint16_t *depth, int n_frames, int rows, int cols, size_t step, size_t frame_stride, float factor) {
        begin(depth, n_frames, step, frame_stride, true, factor);
        spx_compact_result res;
        const int rc = spx_extract_batch_u16_compact(ctx_, depth, n_frames, rows, cols, step, frame_stride, factor, &res);
        finish(rc);
    }

    spx_ctx *context() { return ctx_; }
    int threads() const { return int(workers_.size()); }

private:
    struct Task { int f0, f1; spx_compact_result view; };

    void begin(const void *depth, int n_frames, size_t step, size_t frame_stride, bool u16, float factor) {
        if (int(frames.size()) != n_frames) frames.resize(size_t(n_frames));
        src_.data = depth; src_.pitch = step; src_.frame_stride = frame_stride; src_.u16 = u16; src_.factor = factor;
        configured_ = false;
    }
    void finish(int rc) {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        if (rc != SPX_OK) throw std::runtime_error(std::string("spx_extract_batch_compact: ") + spx_last_error(ctx_));
    }
    static void on_group(void *user, int f0, int f1, const spx_compact_result *view) {
        SequencePlanes *self = static_cast<SequencePlanes *>(user);
        if (!self->configured_) {    // (no worker is running yet: the first group of a call)
            self->ex_.Configure(self->cfg_.fx, self->cfg_.fy, self->cfg_.cx, self->cfg_.cy, view->cloud_width, view->cloud_height, view->cloud_dis);
            self->configured_ = true;
        }
        const int chunk = 4;
        {
            std::lock_guard<std::mutex> lk(self->mu_);
            for (int f = f0; f < f1; f += chunk) {
                self->queue_.push_back(Task{f, f + chunk < f1 ? f + chunk : f1, *view});
                ++self->pending_;
            }
        }
        self->cv_.notify_all();
    }
    void work() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return quit_ || !queue_.empty(); });
                if (queue_.empty()) return;
                t = queue_.front();
                queue_.pop_front();
            }
            for (int f = t.f0; f < t.f1; ++f) frames[size_t(f)].Fill(t.view, f, 0, t.view.frames[f].n_planes, ex_, src_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }

    spx_config cfg_;
    spx_ctx *ctx_ = nullptr;
    DepthSource src_;
    CloudExpander ex_;
    bool configured_ = false;
    std::vector<std::thread> workers_;
    std::deque<Task> queue_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    int pending_ = 0;
    bool quit_ = false;
};

}  // namespace spx_host
