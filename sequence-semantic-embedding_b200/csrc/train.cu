// Training step (placeholder until the BPTT kernels land in this file).
#include "sse_common.cuh"
#include "sse_handle.cuh"
using namespace sse;
extern "C" {
int sse_pair_score(sse_handle* h, const int32_t*, const int32_t*, int, float*, void*) { (void)h; set_error("sse_pair_score: not built yet"); return SSE_ESTATE; }
int sse_train_step(sse_handle* h, const int32_t*, const int32_t*, const float*, int, float*, float*, float*, void*) { (void)h; set_error("sse_train_step: not built yet"); return SSE_ESTATE; }
int sse_train_grads(sse_handle* h, const int32_t*, const int32_t*, const float*, int, int, float*, float*, void*) { (void)h; set_error("sse_train_grads: not built yet"); return SSE_ESTATE; }
int sse_grad_arena(sse_handle* h, float**, int64_t*) { (void)h; set_error("sse_grad_arena: not built yet"); return SSE_ESTATE; }
int sse_train_apply(sse_handle* h, float*, void*) { (void)h; set_error("sse_train_apply: not built yet"); return SSE_ESTATE; }
}
