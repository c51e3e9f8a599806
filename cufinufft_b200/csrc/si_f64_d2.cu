// si_f64_d2.cu -- instantiates the 2-D double spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(double, 2) }
