// oracle/shim/cuda_runtime.h -- stands in for the CUDA runtime header when the
// reference's header-only scene classes are compiled for the HOST by g++
// (oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY.
//
// The reference marks everything __host__ __device__ / __device__ and includes
// "cuda_runtime.h" (e.g. reference Vec3.h:7, Interval.h:5); with the qualifiers
// defined away the same classes are ordinary C++.  nvcc also puts float
// overloads of the math functions in the global namespace, which decides what
// `log(curand_uniform(..))` means (reference ConstantMedium.h:79 -> logf); the
// overload below keeps that resolution under g++.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline

struct RtShimDim3 {
    unsigned x = 0, y = 0, z = 0;
};
static const RtShimDim3 threadIdx, blockIdx;

#ifndef RTSHIM_NO_FLOAT_LOG
inline float log(float x) { return ::logf(x); }
#endif
