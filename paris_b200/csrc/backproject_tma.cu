// backproject_tma.cu -- placeholder until the TMA-staged kernel lands: reports "not handled".
#include "common.cuh"
#include "backproject.cuh"

namespace pb
{
    int launch_bp_tma(paris_b200_ctx*, const float*, size_t, uint32_t, const bp_geometry&, const bp_angles&, float*,
                      bool required, bool* handled)
    {
        *handled = false;
        if(required)
        {
            set_error("TMA backprojection kernel not available");
            return PARIS_B200_ESTATE;
        }
        return PARIS_B200_OK;
    }
}
