#include "cdl_umma_inst.cuh"
namespace hb {
template int launch_cdl_umma_io<double2>(int, int, const CdlArgs&, const CdlTable&, size_t, cudaStream_t);
}
