// si_f32_d2.cu -- instantiates the 2-D float spread / interp kernels (see spreadinterp.cuh).
#include "spreadinterp_launch.cuh"
namespace cfb { CFB_INSTANTIATE_SI(float, 2) }
