#pragma once
namespace rt {
constexpr int kExactStackHost = 96;     // must equal kExactStack in exact.cuh
constexpr int kFastStackHost = 64;      // must equal kFastStack in fast.cuh
}
