// mrefsr_b200/csrc/dcn_tc.cu -- tcgen05 (TF32) DCNv2 forward.  Placeholder until the kernel lands: reports
// "not eligible" so MREFSR_DCN_AUTO resolves to the exact-fp32 CUDA-core kernel and MREFSR_DCN_TF32 errors out.
#include "dcn_common.cuh"

namespace mrefsr {
bool dcn_tc_eligible(const DcnShape&) { return false; }
size_t dcn_tc_workspace_bytes(const DcnShape&, int) { return 0; }
int dcn_forward_tc(const float*, const float*, const float*, const float*, const float*, float*, const DcnShape&,
                   void*, size_t, cudaStream_t) {
    set_error("dcn forward: tcgen05 path not built");
    return ERR_UNSUPPORTED;
}
}  // namespace mrefsr
