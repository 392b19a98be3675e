// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory): the few half-precision types and conversions the device sources use.
#pragma once
#include "cuda_runtime.h"
struct __half { _Float16 v; };
struct __half2 { _Float16 x, y; };
typedef __half half;
typedef __half2 half2;
inline __half2 __floats2half2_rn(float a, float b) { return {(_Float16)a, (_Float16)b}; }
inline float2 __half22float2(__half2 h) { return {(float)h.x, (float)h.y}; }
inline __half __float2half_rn(float a) { return {(_Float16)a}; }
inline __half __float2half(float a) { return {(_Float16)a}; }
inline float __half2float(__half h) { return (float)h.v; }
