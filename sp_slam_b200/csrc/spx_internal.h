// spx_internal.h -- what the other translation units of libspx.so need from the context (C++ linkage: not exported C symbols)
#pragma once
#include <cuda_runtime.h>
struct spx_ctx;
int   spx_internal_fail(spx_ctx *c, int code, const char *what, const char *msg);   // records "what: msg" as the context's last error
void *spx_internal_stream(spx_ctx *c);                                                // cudaStream_t the context's work runs on
int   spx_internal_device(spx_ctx *c);
struct spx_frame_header; struct spx_plane; struct spx_point;
// device-side results of the last extract (either path): frame headers, plane records (offsets into the device arenas) and the
// boundary arena; returns the number of frames (0: nothing yet)
int   spx_internal_last_results(spx_ctx *c, const spx_frame_header **frames, const spx_plane **planes, const spx_point **boundary);

// Every entry point makes the context's device current for its own duration and restores the caller's on the way out: the
// library lives inside a host with its own threads (tracking, mapping, viewer) and possibly other CUDA users.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
