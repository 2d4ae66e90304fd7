#pragma once
namespace rt {
constexpr int kExactStackHost = 96;     // must equal kExactStack in exact.cuh
}
