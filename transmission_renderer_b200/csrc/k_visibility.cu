// k_visibility.cu — K3 (placeholder until the rasteriser lands in this round)
#include "tr_internal.h"
namespace tr {
int32_t launch_visibility(tr_ctx* c, const tr_push_constants& pc) {
    (void)c; (void)pc;
    return fail(TR_ERR_UNSUPPORTED, "tr_visibility: not built yet");
}
}  // namespace tr
