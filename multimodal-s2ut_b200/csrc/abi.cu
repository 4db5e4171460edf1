// Library self-description for the C ABI.
#include "host.cuh"
#include "../../include/mms2ut_b200.h"

extern "C" int mm_abi_version(void) { return 5; }
extern "C" const char* mm_last_error(void) { return mm::g_last_error; }
