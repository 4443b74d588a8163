// scores.cuh -- the dense float32 score track shared by aggregate.cu (reductions) and scores.cu (BinnedArray-style
// construction and reads).
#pragma once
#include "common.cuh"

struct bxg_scores {
    float *v = nullptr;      // device; v[i] = score of position origin + i
    int64_t n = 0;           // logical length
    int64_t cap = 0;         // allocated floats (>= n)
    int32_t origin = 0;
    float fill = 0.0f;       // value of never-written cells (BinnedArray.default; bit pattern kept)
};
