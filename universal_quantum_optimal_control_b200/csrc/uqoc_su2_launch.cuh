// Template dispatch (ST, LPS, BWD) -> kernel launch.  Included by one translation unit per
// (scalar type, sin/cos policy) so the 16 instantiations of each compile in parallel.
#pragma once
#include <type_traits>
#include "uqoc_su2_kernels.cuh"
#include "uqoc_su2_x2.cuh"

namespace uqoc {

template <typename T, int SC>
int su2_launch(const Su2Params<T>& p, const Su2Plan& plan, bool bwd, cudaStream_t stream);

template <typename T, int ST, int LPS, int SC, bool BWD>
static int su2_launch_one(const Su2Params<T>& p, const Su2Plan& plan, cudaStream_t stream) {
    auto kern = su2_kernel<T, ST, LPS, SC, BWD>;
    if (plan.smem > 48 * 1024) {
        static std::atomic<int> smem_set[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(kern, plan.smem, smem_set);
        if (e != cudaSuccess) {
            set_error("su2 kernel needs %zu bytes of shared memory (L too large): %s", plan.smem, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return UQOC_E_UNSUPPORTED;
        }
    }
    const unsigned grid = (unsigned)((long long)p.B * plan.cps);
    // programmatic dependent launch: the blocks may be scheduled while the previous kernel of the stream drains; they
    // touch no input before griddepcontrol.wait (su2_kernel)
    launch_dependent(kern, grid, kThreads, plan.smem, stream, p);
    return launch_status("su2_kernel");
}

template <int NP, int SC, bool BWD, int WPS, int VB>
static int su2_launch_x2w(const Su2Params<float>& p, const Su2Plan& plan, cudaStream_t stream) {
    auto kern = su2_kernel_x2<NP, SC, BWD, WPS, VB>;
    if (plan.smem > 48 * 1024) {
        static std::atomic<int> smem_set[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(kern, plan.smem, smem_set);
        if (e != cudaSuccess) {
            set_error("su2 kernel needs %zu bytes of shared memory (L too large): %s", plan.smem, cudaGetErrorString(e));
            (void)cudaGetLastError();
            return UQOC_E_UNSUPPORTED;
        }
    }
    const unsigned grid = (unsigned)((long long)p.B * plan.cps);
    // programmatic dependent launch: block scheduling and the staging of the (immutable) sin/cos tables overlap the tail
    // of the previous kernel of the stream; nothing else is read before griddepcontrol.wait (su2_kernel_x2)
    launch_dependent(kern, grid, kThreads * VB, plan.smem, stream, p);
    return launch_status("su2_kernel_x2");
}
template <int NP, int SC, bool BWD>
static int su2_launch_x2(const Su2Params<float>& p, const Su2Plan& plan, cudaStream_t stream) {
    if constexpr (NP == 1 && SC == SC_TABLE && BWD) {
        if (plan.wps == 4 && plan.vb == kX2FatVB) return su2_launch_x2w<NP, SC, BWD, 4, kX2FatVB>(p, plan, stream);
    }
    if (plan.vb != 1) {
        set_error("internal: fat-block variant not built for this shape");
        return UQOC_E_UNSUPPORTED;
    }
    if (plan.wps == 4) return su2_launch_x2w<NP, SC, BWD, 4, 1>(p, plan, stream);
    return su2_launch_x2w<NP, SC, BWD, 1, 1>(p, plan, stream);
}

template <typename T, int SC, bool BWD>
static int su2_launch_bwd(const Su2Params<T>& p, const Su2Plan& plan, cudaStream_t stream) {
    const int key = plan.st * 100 + plan.lps;
    if constexpr (std::is_same<T, float>::value && SC != SC_LIBM) {
        if constexpr (SC == SC_POLY) {
            if (plan.packed && plan.table && key == 201) return su2_launch_x2<1, SC_TABLE, BWD>(p, plan, stream);
            if (plan.packed && plan.table && key == 401) return su2_launch_x2<2, SC_TABLE, BWD>(p, plan, stream);
        }
        if (plan.packed && key == 201) return su2_launch_x2<1, SC, BWD>(p, plan, stream);
        if (plan.packed && key == 401) return su2_launch_x2<2, SC, BWD>(p, plan, stream);
    }
    switch (key) {
        case 101: return su2_launch_one<T, 1, 1, SC, BWD>(p, plan, stream);
        case 201: return su2_launch_one<T, 2, 1, SC, BWD>(p, plan, stream);
        case 401: return su2_launch_one<T, 4, 1, SC, BWD>(p, plan, stream);
        case 102: return su2_launch_one<T, 1, 2, SC, BWD>(p, plan, stream);
        case 104: return su2_launch_one<T, 1, 4, SC, BWD>(p, plan, stream);
        case 108: return su2_launch_one<T, 1, 8, SC, BWD>(p, plan, stream);
        case 116: return su2_launch_one<T, 1, 16, SC, BWD>(p, plan, stream);
        case 132: return su2_launch_one<T, 1, 32, SC, BWD>(p, plan, stream);
        default:
            set_error("unsupported launch shape ST=%d LPS=%d", plan.st, plan.lps);
            return UQOC_E_UNSUPPORTED;
    }
}

#define UQOC_INSTANTIATE_SU2(T, SC)                                                                      \
    template <>                                                                                          \
    int su2_launch<T, SC>(const Su2Params<T>& p, const Su2Plan& plan, bool bwd, cudaStream_t stream) {   \
        return bwd ? su2_launch_bwd<T, SC, true>(p, plan, stream) : su2_launch_bwd<T, SC, false>(p, plan, stream); \
    }

}  // namespace uqoc
