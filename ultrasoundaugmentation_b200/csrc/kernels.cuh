// Internal launch helpers shared between translation units of libusaug.so.
#pragma once
#include "common.cuh"

namespace usaug {

// "no bone seen" marker produced by cudaMemsetAsync(.., 0x7f, ..)
constexpr int kNone = 0x7f7f7f7f;

// surf[B][W][2]: {first bone row, H-1-last bone row}, both kNone when the column has no bone.
// Enqueues memset + scan on `stream`.
int launch_surface_scan(const uint8_t* lab, int B, int H, int W, int* surf, cudaStream_t stream);
size_t surface_scan_bytes(int B, int W);

}  // namespace usaug
