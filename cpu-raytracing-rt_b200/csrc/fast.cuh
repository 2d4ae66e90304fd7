// fast.cuh — FAST mode geometry queries (fp32 conservative culling + exact f64 confirmation).
#pragma once
#include "shade.cuh"

namespace rt {

struct FastPolicy : ExactPolicy {};     // placeholder until the fp32 culling hierarchy lands

}  // namespace rt
