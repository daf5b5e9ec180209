// capi_todo.cpp — entry points not implemented yet (return QWGPU_EUNSUPPORTED).
#include "common.h"
#define TODO(name) qw::set_last_error(#name ": not implemented yet"); return QWGPU_EUNSUPPORTED;
extern "C" {
int qwgpu_leaf_search(qwgpu_ctx*, const uint8_t*, size_t, uint8_t**, size_t*) { TODO(qwgpu_leaf_search) }
int qwgpu_invoke_leaf_search(qwgpu_ctx*, const uint8_t*, size_t, uint8_t**, size_t*) { TODO(qwgpu_invoke_leaf_search) }
int qwgpu_compile_plan(const uint8_t*, uint64_t, const char*, const uint8_t*, size_t, const char*, uint8_t**, size_t*) { TODO(qwgpu_compile_plan) }
int qwgpu_merge_leaf_responses(const uint8_t*, size_t, uint32_t, const uint8_t* const*, const size_t*, uint8_t**, size_t*) { TODO(qwgpu_merge_leaf_responses) }
int qwgpu_finalize_aggregation(const char*, const uint8_t*, size_t, char**) { TODO(qwgpu_finalize_aggregation) }
int qwgpu_partial_size(const uint8_t*, size_t, uint64_t*) { TODO(qwgpu_partial_size) }
int qwgpu_response_to_partial(const uint8_t*, size_t, const uint8_t*, size_t, uint8_t*, uint64_t) { TODO(qwgpu_response_to_partial) }
int qwgpu_merge_partials(const uint8_t*, size_t, uint32_t, const uint8_t*, uint64_t, uint8_t**, size_t*) { TODO(qwgpu_merge_partials) }
}
