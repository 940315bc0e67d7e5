// FP64
#include "uqoc_su2_launch.cuh"
namespace uqoc { UQOC_INSTANTIATE_SU2(double, SC_LIBM) }
