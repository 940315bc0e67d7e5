// FP32, MUFU sin/cos (UQOC_FLAG_FAST_SINCOS)
#include "uqoc_su2_launch.cuh"
namespace uqoc { UQOC_INSTANTIATE_SU2(float, SC_MUFU) }
