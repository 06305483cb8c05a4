#include "cdl_umma_inst.cuh"
namespace hb {
template int launch_cdl_umma_io<float2>(int, int, const CdlArgs&, const CdlTable&, size_t, cudaStream_t);
}
