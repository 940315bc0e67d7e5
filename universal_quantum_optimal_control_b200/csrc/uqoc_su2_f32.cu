// FP32, polynomial sin/cos (default accuracy path)
#include "uqoc_su2_launch.cuh"
namespace uqoc { UQOC_INSTANTIATE_SU2(float, SC_POLY) }
