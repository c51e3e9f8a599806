// si_f32_d3.cu -- instantiates the 3-D float spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(float, 3) }
