// Internal launch helpers shared between translation units of libusaug.so.
#pragma once
#include "common.cuh"

namespace usaug {

// "no bone seen" marker for first-row values
constexpr int kNone = 0x7f7f7f7f;

// surf[B][scan_tiles(H)][W][2]: per 64-row tile {first bone row or kNone, last bone row or -1}.
// One scan kernel on `stream` (no memset, no atomics); consumers reduce over the tiles.
int launch_surface_scan(const uint8_t* lab, int B, int H, int W, int* surf, cudaStream_t stream);
int scan_tiles(int H);
// surface scan + per-image bone box {y_top, y_bottom, x_left, x_right}; see warp_tgc.cu
int launch_bone_box(const uint8_t* lab, int B, int H, int W, int* box, int* surf, cudaStream_t stream, bool followed_by_kernel);
size_t surface_scan_bytes(int B, int H, int W);

}  // namespace usaug
