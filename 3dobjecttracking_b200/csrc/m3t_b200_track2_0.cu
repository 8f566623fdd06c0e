// k_track2 instantiation 0 of 4 (own translation unit: nvcc -t builds them in parallel)
#define M3TB_TRACK_TU 1
#include "m3t_b200_track2.cuh"

namespace m3tb {
template __global__ void k_track2<1024, true>(const __grid_constant__ TrackArgs);
}  // namespace m3tb
