// kernels_f32.cu — RTC_F32 (production) instantiation of the wavefront kernels. Built with the default -fmad=true.
#include "rtc_device.cuh"

namespace rtc {
template struct Kernels<float>;
}
