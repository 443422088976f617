// placeholder — replaced by the tcgen05 implementation
#include "common.cuh"
namespace spn {
size_t mlp_tc_packed_bytes() { return 16; }
size_t mlp_tc_stash_bytes(int64_t) { return 16; }
size_t mlp_tc_bwd_ws_bytes(int64_t) { return 16; }
int mlp_tc_pack(const float*, void*, cudaStream_t) { set_error("tcgen05 path not built"); return SPN_E_ARG; }
int mlp_tc_fwd(const void*, const SampleSource&, int64_t, float*, void*, cudaStream_t) { set_error("tcgen05 path not built"); return SPN_E_ARG; }
int mlp_tc_bwd(const void*, const void*, const float*, int64_t, float*, void*, cudaStream_t) { set_error("tcgen05 path not built"); return SPN_E_ARG; }
}
