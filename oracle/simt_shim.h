// SIMT emulation shim: lets the CUDA kernels emitted by the REFERENCE's code generator be compiled
// with g++ and executed on the CPU, block by block and thread by thread (oracle/build_ref.py).
// TEST INFRASTRUCTURE ONLY.  The generated kernels use no shared memory, no barriers and no warp
// intrinsics, so a serial sweep over (blockIdx, threadIdx) is exact;
// float atomics become ordered read-modify-writes.
#pragma once
#include <algorithm>
#include <cmath>

struct Dim3Emu { int x = 0, y = 0, z = 0; };
static thread_local Dim3Emu blockIdx, threadIdx, blockDim, gridDim;

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline float atomicMax(float* p, float v) { float o = *p; *p = std::max(o, v); return o; }
static inline float atomicMin(float* p, float v) { float o = *p; *p = std::min(o, v); return o; }
using std::exp;
using std::max;
using std::min;
