#include "fading_tma_inst.cuh"
namespace hb {
HB_INSTANTIATE_FADING_TMA(1)
}
