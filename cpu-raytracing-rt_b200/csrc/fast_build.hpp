// fast_build.hpp — host: derives the fp32 culling hierarchy of FAST mode from a built BVH.
#pragma once
#include <vector>

#include "device_scene.cuh"
#include "scene_host.hpp"

namespace rt {

struct FastHierarchy {
    std::vector<FNode> nodes;
    std::vector<FTri> tris;
    int32_t root = kFastEmpty;
    uint32_t depth = 0;
    uint32_t top = 0;                   // nodes [0, top) are the top levels in breadth-first order
};
void build_fast_hierarchy(const HostBvh<HostTri> &bvh, FastHierarchy &out);

}  // namespace rt
