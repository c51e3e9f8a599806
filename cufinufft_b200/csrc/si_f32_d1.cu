// si_f32_d1.cu -- instantiates the 1-D float spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(float, 1) }
