// render_hier.cu -- HIER sort mode (placeholder until the hierarchical kernels land).
#include "stp_kernels.cuh"
namespace stp {
cudaError_t launch_render_hier_fwd(const Frame&, const Settings&, const RenderArgs&, cudaStream_t) {
    return cudaErrorNotSupported;
}
cudaError_t launch_render_hier_bwd(const Frame&, const Settings&, const RenderBwdArgs&, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace stp
