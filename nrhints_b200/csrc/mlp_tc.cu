// tcgen05 tensor-core MLP engine -- placeholder until the split-fp16 kernel lands.
#include "mlp_tc.cuh"

namespace nrh {

bool tc_available() { return false; }
size_t tc_packed_bytes(const NrhConfig&) { return 0; }
size_t tc_scratch_bytes(int) { return 0; }
int tc_pack(const NrhConfig&, const PackedLayout&, void*, cudaStream_t) { return NRH_OK; }
int sdf_mlp_tc(const void*, const PackedLayout&, Strided3, int64_t, float*, float*, float*, float*, int64_t, float*,
               float*, size_t, int, cudaStream_t) {
    set_error("tcgen05 engine not built");
    return NRH_ERR_UNSUPPORTED;
}
int color_mlp_tc(const void*, const PackedLayout&, Strided3, Strided3, const float*, const float*, int64_t, int64_t,
                 float*, float*, float*, float*, size_t, int, cudaStream_t) {
    set_error("tcgen05 engine not built");
    return NRH_ERR_UNSUPPORTED;
}

}  // namespace nrh
