// Thread-local last-error text behind upb_last_error(); every C-ABI entry point reports through set_error.
#pragma once
#include <string>

namespace upb {
int set_error(int code, const std::string& msg);   // stores msg, returns code
const char* last_error_text();
}  // namespace upb
