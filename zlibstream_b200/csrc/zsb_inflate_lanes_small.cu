// zsb_inflate_lanes_small.cu -- second instance of the lane-per-segment inflate kernel with smaller per-lane tables:
// 7-bit literal/length and 5-bit distance root tables, 32 long-code symbols in shared memory (496 bytes per lane instead of
// 936 => 12 warps per SM instead of 7; 14 would fit, 12 leave the L1 its 60 KiB -- see ZSB_LN_WARPS_FULL).  More long-code searches per symbol, so it loses while every segment has a lane
// anyway (2 GiB of 64 KiB segments: 98 against 105 GB/s), and wins once segments queue for lanes (16 GiB: 135 against 105 GB/s,
// profiles/r01n_inflate_sweep.txt).  zsb_inflate_segments picks it from the segment count.
#define ZSB_LN_NS lanes_small
#define ZSB_LN_ROOT_L 7
#define ZSB_LN_ROOT_D 5
#ifndef ZSB_LN_LONG_L
#define ZSB_LN_LONG_L 32
#endif
#ifndef ZSB_LN_WARPS_FULL
#define ZSB_LN_WARPS_FULL 12
#endif
#define ZSB_LN_KERNEL inflate_lanes_small_kernel
#define ZSB_LN_LAUNCH launch_inflate_lanes_small
#define ZSB_LN_SECONDARY 1
#include "zsb_inflate_lanes.cu"
