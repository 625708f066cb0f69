// triangle.h — forwards to the single-header mirror of the reference API (see realtrace_api.h).
#include "realtrace_api.h"
