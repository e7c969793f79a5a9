// spa_qc_spec.cu -- launches of the compile-time specialised resident kernels (qc_kernel.cuh)
// for the base matrices listed in qc_registry.json.  A graph whose (z, mb, nb, shift table)
// matches an entry exactly runs the specialised kernel; any other quasi-cyclic graph falls
// back to the table-driven kernel in spa_qc_resident.cu.
#include "ldpc_common.cuh"
#include "qc_codes_gen.cuh"

#include <algorithm>

namespace ldpc {

namespace {

template <class C, bool EARLY, int THREADS, int MINB>
int launch_spec(C code, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                const qc::Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kp = qc::k_qc_spec<THREADS, MINB, EARLY, C>;
    const size_t smem = sizeof(float) * 4 * (size_t)C::N;      // channel + 2 posteriors + TMA stage
    static thread_local bool configured = false;
    static thread_local int per_sm = 0;
    if (!configured) {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kp, THREADS, smem));
        configured = true;
    }
    if (per_sm < 1) { set_error("specialised resident kernel does not fit on an SM"); return LDPC_ERR_UNSUPPORTED; }
    const int grid = (int)std::min<int64_t>(frames, (int64_t)per_sm * di.sm_count);
    unsigned long long* counter = nullptr;
    if (EARLY) {
        if (!ws) { set_error("early termination needs a workspace (work counter)"); return LDPC_ERR_WORKSPACE; }
        counter = (unsigned long long*)ws;
        LDPC_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    }
    (void)code;
    kp<<<grid, THREADS, smem, stream>>>(llr, out, (long long)frames, max_iter,
                                        (flags & LDPC_FLAG_FIX_ODD_SIGN) ? 1 : 0, mc, counter);
    LDPC_LAUNCH_CHECK();
    return LDPC_OK;
}

struct Args {
    const ldpc_graph* g; int64_t frames; int max_iter; unsigned flags; const float* llr; qc::Outputs out;
    const McParams* mc; void* ws; cudaStream_t stream;
};

template <class C>
int launch_code(C code, const Args& a)
{
    const bool early = (a.flags & LDPC_FLAG_EARLY_TERM) != 0;
    constexpr int T = qc::LaunchShape<C>::THREADS;
    constexpr int B = qc::LaunchShape<C>::MINB;
    return early ? launch_spec<C, true, T, B>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream)
                 : launch_spec<C, false, T, B>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream);
}

}  // namespace

int qc_spec_find(const ldpc_graph* g)
{
    if (!g || !g->is_qc) return -1;
    for (int i = 0; i < qc::kRegistrySize; ++i) {
        const qc::RegistryEntry& e = qc::kRegistry[i];
        if (e.z != g->qc.z || e.mb != g->qc.mb || e.nb != g->qc.nb) continue;
        if (std::equal(g->qc.shift.begin(), g->qc.shift.end(), e.shift)) return i;
    }
    return -1;
}

int qc_spec_decode(int idx, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                   const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                   uint8_t* ok_dev, float* post_dev, const McParams& mc, void* ws, cudaStream_t stream)
{
    Args a{g, frames, max_iter, flags, llr_dev, qc::Outputs{z_dev, zbits_dev, conv_dev, ok_dev, post_dev}, &mc, ws, stream};
    auto go = [&](auto code) { return launch_code(code, a); };
    LDPC_QC_SPEC_DISPATCH(idx, go)
    set_error("no specialised kernel with index %d", idx);
    return LDPC_ERR_UNSUPPORTED;
}

}  // namespace ldpc
