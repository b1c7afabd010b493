// isx_host_expand.h — see isx_host_expand.cpp
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace isx {
// rec  : n compact records of 32 floats (31 obs features + alive flag != 0)
// hits : n x R lidar hit indices (u8, dense stride R)
// dst  : n obs rows of 127 floats (fully written, zero tail included)
void expand_obs_rows(const float* rec, const uint8_t* hits, int R, float* dst, size_t n);
const float* expand_lidar_lut();   // 256 floats: value of obs[31 + i] for hit index k
}  // namespace isx
