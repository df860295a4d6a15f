// C-ABI housekeeping entry points (include/feddat_b200.h).
#include "feddat_b200.h"
#include "host_common.h"

extern "C" const char* feddat_last_error(void) { return fd::last_error_buf(); }
extern "C" int feddat_abi_version(void) { return 5; }
