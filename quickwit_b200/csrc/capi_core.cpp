// capi_core.cpp — error plumbing and trivial entry points of the C ABI (include/qwgpu.h).
#include "common.h"

namespace qw {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace qw

extern "C" {
const char* qwgpu_last_error(void) { return qw::g_last_error.c_str(); }
const char* qwgpu_version(void) { return "qwgpu 0.1.0 (sm_100a)"; }
void qwgpu_buf_free(void* buf) { free(buf); }
}
