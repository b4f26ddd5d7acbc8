// api_internal.hpp -- what the translation units behind the C ABI share (not exported).
#pragma once

#include <string>

namespace sshash_b200 {

// records `msg` as this thread's sshash_gpu_last_error() and returns `status`
int set_last_error(int status, const std::string& msg);

}  // namespace sshash_b200
