// Miscellaneous C-ABI entry points.
#include "mmb200_internal.h"
extern "C" int mmb_version(void) { return 100; }
