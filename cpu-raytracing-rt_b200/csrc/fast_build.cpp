#include "fast_build.hpp"

namespace rt {

void build_fast_hierarchy(const HostBvh<HostTri> &, FastHierarchy &out) { out = FastHierarchy{}; }

}  // namespace rt
