// Minimal stand-in so the reference's HOST sources (input_parser.cpp, utils.cpp) compile with g++
// without ROCm.  Only types/enums named by utils.h outside USE_CUDA are provided.  Test infrastructure.
#pragma once
typedef int hipError_t;
#define hipSuccess 0
inline const char *hipGetErrorString(hipError_t) { return "hip-stub"; }
