// TEST INFRASTRUCTURE ONLY: stand-in for onepiece_b200/csrc/opb_common.cuh when the kd-tree device code runs on the host emulator
#pragma once
#include "cuda_emu.h"
namespace opb
{
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float fdiv(float a, float b) { return a / b; }
} // namespace opb
