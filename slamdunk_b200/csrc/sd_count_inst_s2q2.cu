// one column form of the count kernel: SEQ form 2, quality planes 2 (0 = bytes); see sd_count_kernel.cuh
#include "sd_count_lean.cuh"

template int launch_count<4, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
template int launch_count<6, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
template int launch_count<8, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
template int launch_count<10, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
template int launch_count<12, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
template int launch_count<16, 2, 2>(sd_ctx *, CountArgs &, cudaStream_t, bool);
