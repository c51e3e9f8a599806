// si_f64_d1.cu -- instantiates the 1-D double spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(double, 1) }
