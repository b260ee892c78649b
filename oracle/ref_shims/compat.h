// Build shim (NOT reference code): glog's CHECK_EQ macro is no longer exported by torch >= 2.
#pragma once
#include <c10/util/Exception.h>
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b))
#endif
