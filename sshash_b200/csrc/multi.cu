// multi.cu -- multi-GPU handle behind the C ABI (filled in below)
#include "api_internal.hpp"
