// hp_internal.cuh - launch helpers shared between translation units of libhp_b200.so.
#pragma once
#include "hp_common.cuh"

namespace hp {

// decode n_maps maps; any of preds/maxvals/idx/centres may be null.
// centres[map] = (int(px) >> shift, int(py) >> shift)  -- regda_7.py:3033 / :3195 `(preds / d).astype(int)`
int launch_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals, int32_t* idx,
                  int32_t* centres, int shift, cudaStream_t stream);

}  // namespace hp
