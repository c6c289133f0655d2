// hp_dispatch.cuh - pick the thread-group shape for a map of HW floats.
//
// <TPM threads per map, NV float4 per thread, walk mode, MPB maps per block>
//   16x16   ( 1 KB): one warp per map, 2 x 128-bit loads per lane, 4 maps per block
//   32x32   ( 4 KB): one warp per map, 8 loads per lane (no block barrier at all)
//   64x64   (16 KB): LIGHT 128 threads x 8 loads | HEAVY (two inputs in registers) 256 x 4
//   128x128 (64 KB): LIGHT 256 x 16             | HEAVY 512 x 8
//   anything else  : 256-thread guarded tiles (float4 when H*W % 4 == 0 and the base is 16-byte
//                    aligned, scalar otherwise), online statistics across tiles.
#pragma once
#include "hp_common.cuh"

namespace hp {

template <bool HEAVY = false, class Launcher>
inline void dispatch_map_walk(int HW, bool base_aligned16, const Launcher& l) {
    if (base_aligned16 && (HW % 4) == 0) {
        if (HW == 256) return l.template run<32, 2, WALK_EXACT, 4>();
        if (HW == 1024) return l.template run<32, 8, WALK_EXACT, 4>();
        if (HW == 4096) {
            if (HEAVY) return l.template run<256, 4, WALK_EXACT, 1>();
            return l.template run<128, 8, WALK_EXACT, 1>();
        }
        if (HW == 16384) {
            if (HEAVY) return l.template run<512, 8, WALK_EXACT, 1>();
            return l.template run<256, 16, WALK_EXACT, 1>();
        }
        return l.template run<256, 4, WALK_VEC, 1>();
    }
    return l.template run<256, 4, WALK_SCALAR, 1>();
}

}  // namespace hp
