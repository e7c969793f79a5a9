// spa_qc_spec.cu -- launches of the compile-time specialised resident kernels (qc_kernel.cuh)
// for the base matrices listed in qc_registry.json.  A graph whose (z, mb, nb, shift table)
// matches an entry exactly runs the specialised kernel; any other quasi-cyclic graph falls
// back to the table-driven kernel in spa_qc_resident.cu.
#include "ldpc_common.cuh"
#include "qc_codes_gen.cuh"
#include "qc_kernel_pair.cuh"
#include "qc_kernel_gather.cuh"
#include "qc_kernel_gather1.cuh"

#include <algorithm>
#include <cstdlib>

namespace ldpc {

namespace {

// Per-device launch configuration of one kernel instantiation (function attributes are per device: a process
// that touches a second GPU must set them again there).
struct KernelConfig {
    bool configured[64] = {};
    int per_sm[64] = {};
};

// CTAs of kernel kp that fit on one SM when the SM's whole shared memory is available to it (the kernels ask
// for the maximum carve-out; cudaOccupancyMaxActiveBlocksPerMultiprocessor answers for the DEFAULT carve-out
// and reports one CTA of 78 KB where two fit).
template <class KP>
int ctas_per_sm(KP kp, int threads, size_t dyn_smem, int* out)
{
    cudaFuncAttributes fa;
    LDPC_CUDA_TRY(cudaFuncGetAttributes(&fa, kp));
    int dev = 0, regs_sm = 0, smem_sm = 0, warps_sm = 0, reserved = 0;
    LDPC_CUDA_TRY(cudaGetDevice(&dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&warps_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev));
    const int warps = (threads + 31) / 32;
    const int regs_warp = ((fa.numRegs + 7) / 8 * 8) * 32;                  // allocation unit: 8 registers per thread
    const size_t smem_cta = (dyn_smem + fa.sharedSizeBytes + (size_t)reserved + 127) / 128 * 128;
    int n = regs_sm / (regs_warp * warps);
    n = std::min<int>(n, (int)((size_t)smem_sm / smem_cta));
    n = std::min<int>(n, warps_sm / 32 / warps);
    *out = std::min(n, 32);
    return LDPC_OK;
}

template <class KP>
int configure(KP kp, KernelConfig& kc, int device, int threads, size_t smem, int* per_sm)
{
    if (device < 0 || device >= 64) { set_error("device ordinal %d out of range", device); return LDPC_ERR_INVALID; }
    if (!kc.configured[device]) {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int api = 0, own = 0;
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&api, kp, threads, smem));
        if (int rc = ctas_per_sm(kp, threads, smem, &own)) return rc;
        kc.per_sm[device] = std::max(api, own);
        kc.configured[device] = true;
    }
    *per_sm = kc.per_sm[device];
    return LDPC_OK;
}

template <class C, bool EARLY, int THREADS, int MINB>
int launch_spec(C code, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                const qc::Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kp = qc::k_qc_spec<THREADS, MINB, EARLY, C>;
    const size_t smem = sizeof(float) * 4 * (size_t)C::N;      // channel + 2 posteriors + TMA stage
    static KernelConfig kc;
    int per_sm = 0;
    if ((rc = configure(kp, kc, di.device, THREADS, smem, &per_sm)) != LDPC_OK) return rc;
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

// Two frames per thread (qc_kernel_pair.cuh): float2 posteriors and messages, packed fp32 arithmetic.
// TM: the messages live in tensor memory between passes; the CTAs per SM are then capped by the TMEM columns
// (a CTA that cannot allocate would wait for ever), which is enforced through the shared-memory request.
template <class C, bool EARLY, bool TM, int THREADS, int MINB>
int launch_pair(C code, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                const qc::Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kp = [] { if constexpr (TM) return qc::k_qc_pair_tmem<THREADS, MINB, EARLY, C>; else return qc::k_qc_pair<THREADS, MINB, EARLY, C>; }();
    size_t smem = sizeof(float2) * 4 * (size_t)C::N;           // float2: channel + 2 posteriors; stage = two LLR rows
    if (TM) {
        // at most MAX_CTAS CTAs may share an SM: ask for more than 1/(MAX_CTAS+1) of the SM's shared memory
        const size_t floor_bytes = (size_t)di.max_smem_optin / (qc::TmemShape<C>::MAX_CTAS + 1) + 1024;
        smem = std::max(smem, floor_bytes);
    }
    static KernelConfig kc;
    int per_sm = 0;
    if ((rc = configure(kp, kc, di.device, THREADS, smem, &per_sm)) != LDPC_OK) return rc;
    if (per_sm < 1) { set_error("pair resident kernel does not fit on an SM"); return LDPC_ERR_UNSUPPORTED; }
    if (TM && per_sm > qc::TmemShape<C>::MAX_CTAS) { set_error("internal: TMEM occupancy cap not effective"); return LDPC_ERR_UNSUPPORTED; }
    const int64_t pairs = (frames + 1) / 2;
    if (const char* force = getenv("LDPC_PAIR_PER_SM")) per_sm = atoi(force);      // tuning experiments only
    const int grid = (int)std::min<int64_t>(pairs, (int64_t)per_sm * di.sm_count);
    if (getenv("LDPC_TRACE_LAUNCH"))
        fprintf(stderr, "[ldpc] pair kernel tm=%d early=%d threads=%d smem=%zu per_sm=%d grid=%d frames=%lld\n", (int)TM, (int)EARLY,
                THREADS, smem, per_sm, grid, (long long)frames);
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

// Two frames per thread, barrier-free check-node phase + gather variable-node phase (qc_kernel_gather.cuh).
template <class C, bool EARLY, int THREADS, int MINB>
int launch_gather(C code, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                  const qc::Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    using SH = qc::GatherShape<C>;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kp = qc::k_qc_gather<THREADS, MINB, EARLY, C>;
    // at most MAX_CTAS CTAs may share an SM (TMEM columns): ask for more than 1/(MAX_CTAS+1) of its shared memory
    const size_t smem = std::max(SH::SMEM, (size_t)di.max_smem_optin / (SH::MAX_CTAS + 1) + 1024);
    if (smem > (size_t)di.max_smem_optin) { set_error("gather kernel: %zu bytes of shared memory do not fit", smem); return LDPC_ERR_UNSUPPORTED; }
    static KernelConfig kc;
    int per_sm = 0;
    if ((rc = configure(kp, kc, di.device, THREADS, smem, &per_sm)) != LDPC_OK) return rc;
    if (per_sm < 1) { set_error("gather resident kernel does not fit on an SM"); return LDPC_ERR_UNSUPPORTED; }
    if (per_sm > SH::MAX_CTAS) { set_error("internal: TMEM occupancy cap not effective"); return LDPC_ERR_UNSUPPORTED; }
    const int64_t pairs = (frames + 1) / 2;
    if (const char* force = getenv("LDPC_PAIR_PER_SM")) per_sm = atoi(force);      // tuning experiments only
    const int grid = (int)std::min<int64_t>(pairs, (int64_t)per_sm * di.sm_count);
    if (getenv("LDPC_TRACE_LAUNCH"))
        fprintf(stderr, "[ldpc] gather kernel early=%d threads=%d smem=%zu per_sm=%d grid=%d frames=%lld\n", (int)EARLY,
                THREADS, smem, per_sm, grid, (long long)frames);
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

// One frame per thread, gather structure, tensor-memory messages (qc_kernel_gather1.cuh): up to four CTAs per SM.
template <class C, bool EARLY, int THREADS, int MINB>
int launch_gather1(C code, const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags, const float* llr,
                   const qc::Outputs& out, const McParams& mc, void* ws, cudaStream_t stream)
{
    using SH = qc::Gather1Shape<C>;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    auto kp = qc::k_qc_gather1<THREADS, MINB, EARLY, C>;
    const size_t smem = std::max(SH::SMEM, (size_t)di.max_smem_optin / (SH::MAX_CTAS + 1) + 1024);
    if (smem > (size_t)di.max_smem_optin) { set_error("gather1 kernel: %zu bytes of shared memory do not fit", smem); return LDPC_ERR_UNSUPPORTED; }
    static KernelConfig kc;
    int per_sm = 0;
    if ((rc = configure(kp, kc, di.device, THREADS, smem, &per_sm)) != LDPC_OK) return rc;
    if (per_sm < 1) { set_error("gather1 resident kernel does not fit on an SM"); return LDPC_ERR_UNSUPPORTED; }
    if (per_sm > SH::MAX_CTAS) { set_error("internal: TMEM occupancy cap not effective"); return LDPC_ERR_UNSUPPORTED; }
    if (const char* force = getenv("LDPC_PAIR_PER_SM")) per_sm = atoi(force);      // tuning experiments only
    const int grid = (int)std::min<int64_t>(frames, (int64_t)per_sm * di.sm_count);
    if (getenv("LDPC_TRACE_LAUNCH"))
        fprintf(stderr, "[ldpc] gather1 kernel early=%d threads=%d smem=%zu per_sm=%d grid=%d frames=%lld\n", (int)EARLY,
                THREADS, smem, per_sm, grid, (long long)frames);
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
    // The pair kernel needs enough frames to fill the machine with CTAs of two frames each; with fewer (or on
    // request) the one-frame kernel spreads the batch over more SMs.
    DeviceInfo di;
    if (int rc = get_device_info(&di)) return rc;
    // Kernel choice.  Fixed iteration count (the throughput configuration): two frames per thread, barrier-free
    // check-node phase + gather (qc_kernel_gather.cuh), 8 % faster than the one-frame kernel.  Early termination: the
    // one-frame kernel, because a pair iterates until BOTH of its frames are done (E[max] of two iteration counts
    // costs more than the pair kernel gains).  The LDPC_FLAG_PAIR_* flags force a variant (tests, A/B timing).
    // Early termination on a batch that fills the machine: the one-frame GATHER kernel (24 warps per SM) is 4-10 % faster
    // than the one-frame scatter kernel there (tools/mc_et_probe.py); LDPC_FLAG_ONE_FRAME keeps the latter.
    const bool et_gather1 = (a.flags & LDPC_FLAG_EARLY_TERM) && a.frames >= 4 * (int64_t)di.sm_count &&
                            !(a.flags & (LDPC_FLAG_ONE_FRAME | LDPC_FLAG_PAIR_REGS | LDPC_FLAG_PAIR_SCATTER | LDPC_FLAG_PAIR_GATHER));
    if constexpr (qc::Gather1Shape<C>::FITS) if ((a.flags & LDPC_FLAG_ONE_GATHER) || et_gather1) {
        constexpr int GB = qc::Gather1Shape<C>::MINB;
        return early ? launch_gather1<C, true, T, GB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream)
                     : launch_gather1<C, false, T, GB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream);
    }
    const bool enough = a.frames >= 4 * (int64_t)di.sm_count;
    const unsigned forced = a.flags & (LDPC_FLAG_PAIR_REGS | LDPC_FLAG_PAIR_SCATTER | LDPC_FLAG_PAIR_GATHER);
    const bool pair_ok = !(a.flags & LDPC_FLAG_ONE_FRAME) && enough && (!early || forced);
    if constexpr (qc::GatherShape<C>::FITS) if (pair_ok && !(forced & (LDPC_FLAG_PAIR_REGS | LDPC_FLAG_PAIR_SCATTER))) {
        constexpr int GB = qc::GatherShape<C>::MINB;
        return early ? launch_gather<C, true, T, GB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream)
                     : launch_gather<C, false, T, GB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream);
    }
    if constexpr (qc::PairShape<C>::TM_ENABLED) if (pair_ok && !(forced & LDPC_FLAG_PAIR_REGS)) {
        constexpr int PB = qc::PairShape<C>::TM_MINB;
        return early ? launch_pair<C, true, true, T, PB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream)
                     : launch_pair<C, false, true, T, PB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream);
    }
    if constexpr (qc::PairShape<C>::ENABLED) if (pair_ok) {
        constexpr int PB = qc::PairShape<C>::MINB;
        return early ? launch_pair<C, true, false, T, PB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream)
                     : launch_pair<C, false, false, T, PB>(code, a.g, a.frames, a.max_iter, a.flags, a.llr, a.out, *a.mc, a.ws, a.stream);
    }
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
