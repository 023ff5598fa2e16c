/* ORACLE shim — stands in for the OptiX SDK's <optix.h>, which is not installed here.
 * The reference headers we compile on the host only need the handle typedef. */
#pragma once
#include <cstdint>
typedef unsigned long long OptixTraversableHandle;
