#include "fading_inst.cuh"
namespace hb {
HB_INSTANTIATE_FADING(1)
}
