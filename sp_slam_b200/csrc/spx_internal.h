// spx_internal.h -- what the other translation units of libspx.so need from the context (C++ linkage: not exported C symbols)
#pragma once
struct spx_ctx;
int   spx_internal_fail(spx_ctx *c, int code, const char *what, const char *msg);   // records "what: msg" as the context's last error
void *spx_internal_stream(spx_ctx *c);                                                // cudaStream_t the context's work runs on
int   spx_internal_device(spx_ctx *c);
struct spx_frame_header; struct spx_plane; struct spx_point;
// device-side results of the last extract (either path): frame headers, plane records (offsets into the device arenas) and the
// boundary arena; returns the number of frames (0: nothing yet)
int   spx_internal_last_results(spx_ctx *c, const spx_frame_header **frames, const spx_plane **planes, const spx_point **boundary);
