// log_prior_tc.cu -- tcgen05/TMEM log-prior (placeholder until the tensor-core kernel lands).
#include "mas_host.h"

namespace masb200 {

int launch_log_prior_tc(const float *, const float *, int, int, int, int, float *, cudaStream_t) {
    return MAS_B200_ERR_UNSUPPORTED;
}

}  // namespace masb200
