#include "errors.h"

#include "../../include/upb200.h"

namespace upb {
namespace {
thread_local std::string g_last_error;
}
int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
const char* last_error_text() { return g_last_error.c_str(); }
}  // namespace upb

extern "C" const char* upb_last_error(void) { return upb::last_error_text(); }
extern "C" int upb_abi_version(void) { return UPB_ABI_VERSION; }
