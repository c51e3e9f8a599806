// si_f64_d3.cu -- instantiates the 3-D double spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(double, 3) }
