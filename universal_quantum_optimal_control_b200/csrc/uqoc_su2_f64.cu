// FP64: table sin/cos (default) and libm sincos (UQOC_FLAG_NO_TABLE)
#include "uqoc_su2_launch.cuh"
namespace uqoc { UQOC_INSTANTIATE_SU2(double, SC_TABLE) UQOC_INSTANTIATE_SU2(double, SC_LIBM) }
