#pragma once
namespace rt {
constexpr int kExactStackHost = 96;     // must equal kExactStack in exact.cuh
constexpr int kFastStackHost = 64;      // must equal kFastStack in fast.cuh
constexpr float kFastRange = 1e15f;     // scene bound / |ray origin| up to which FAST mode uses its fp32 culling arithmetic (fast.cuh)
}
