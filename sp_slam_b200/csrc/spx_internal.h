// spx_internal.h -- what the other translation units of libspx.so need from the context (C++ linkage: not exported C symbols)
#pragma once
struct spx_ctx;
int   spx_internal_fail(spx_ctx *c, int code, const char *what, const char *msg);   // records "what: msg" as the context's last error
void *spx_internal_stream(spx_ctx *c);                                                // cudaStream_t the context's work runs on
int   spx_internal_device(spx_ctx *c);
