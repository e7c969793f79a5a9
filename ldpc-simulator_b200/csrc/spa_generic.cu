// spa_generic.cu -- flooding sum-product decoder for ANY parity-check graph.
//
// Replaces SPA_Decoder.decode (python_ldpc_app/spa_decoder.py:63-280), one
// call per batch of frames instead of one call per frame.  The arithmetic of
// the reference is kept literally (that is the point of the fp64 parity mode):
//   check node   t = tanh(M/2) with the +-17.5 branch clip      (:133-146)
//                P = prod t ; r = P / t_j  (|t_j| > 1e-10)       (:151-161)
//                     or the explicit product of the others      (:162-164)
//                E = 2 atanh(clip(r, +-0.99999999999999878))     (:167-168)
//   posterior    L = Lch + sum_i E_ij ; z = (L < 0)              (:173-188)
//   syndrome     H (z xor 1) mod 2 == 0                          (:191-204)
//   var node     M_ij = L_j - E_ij                               (:260-268)
//
// Data layout in HBM (one chunk of Fc frames, Fc a multiple of 32):
//   lch  [n][Fc]       channel LLRs, frame-minor
//   E    [2][nnz][Fc]  check->variable messages of the previous / the current pass,
//                      edge-major (CSR edge order)
//   post [n][Fc]       posteriors of the last executed pass
//   zb   [n][Fc]       hard decisions (uint8)
// so that a warp = 32 consecutive frames of one node and every access is a
// coalesced 128/256-byte row segment.  Frames that have converged are dropped
// from the active list (LDPC_FLAG_COMPACT) or masked (default).
//
// The variable->check messages are never stored: the check-node pass forms
// M_ij = L_j - E_ij (:260-268) from the posterior and the message of the previous
// pass -- the same subtraction on the same operands, so results are unchanged --
// which removes one message sweep in each direction.  Work is ordered frame block
// major (consecutive warps = consecutive nodes of the SAME 32 frames): the
// posterior words a check gathers were written / are re-read by warps that run at
// about the same time and come from L2, not HBM.
//
// The kernels are HBM-streaming by construction (per pass and frame: read E,
// write E, read E, plus O(n) vectors); DESIGN.md gives the roofline.
#include "ldpc_common.cuh"
#include "f64_math.cuh"

namespace ldpc {

namespace {

constexpr int kThreads = 256;

template <typename T> struct Num;
template <> struct Num<double> {
    // spa_decoder.py:140-146,167: clip constants of the reference.
    static __device__ __forceinline__ double tanh_arg_limit() { return 17.5; }
    static __device__ __forceinline__ double unit_clip() { return 0.99999999999999878; }
    static __device__ __forceinline__ double small_tanh() { return 1e-10; }
    // f64_math.cuh: tanh / atanh / division restricted to the arguments the check node can produce (3.3 x fewer
    // instructions than libm's; <= 3.5 ulp, correctly rounded tanh near saturation).  -DLDPC_F64_LIBM: CUDA's libm.
#ifdef LDPC_F64_LIBM
    static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
    static __device__ __forceinline__ double atanh_(double x) { return atanh(x); }
#else
    static __device__ __forceinline__ double tanh_(double x) { return f64::tanh_half(x + x); }
    static __device__ __forceinline__ double atanh_(double x) { return 0.5 * f64::two_atanh(x); }
#endif
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
};
template <> struct Num<float> {
    // fp32 restatement: tanhf saturates to 1 near 9.01, so the clip value is the
    // largest float below 1; |E| <= 2*atanhf(1-2^-24) = 17.33 instead of 35.03.
    static __device__ __forceinline__ float tanh_arg_limit() { return 17.5f; }
    static __device__ __forceinline__ float unit_clip() { return 0.99999994f; }
    static __device__ __forceinline__ float small_tanh() { return 1e-10f; }
    static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
    static __device__ __forceinline__ float atanh_(float x) { return atanhf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
};

template <typename T, bool TAB = true>
__device__ __forceinline__ T tanh_half_clipped(T msg)
{
#ifndef LDPC_F64_LIBM
    if constexpr (sizeof(T) == 8) return f64::tanh_half<TAB>(msg);    // clip included
#endif
    const T h = msg / T(2);
    if (h > Num<T>::tanh_arg_limit()) return Num<T>::unit_clip();
    if (h < -Num<T>::tanh_arg_limit()) return -Num<T>::unit_clip();
    T t = Num<T>::tanh_(h);
    // fp32 only: tanhf may round to exactly +-1, which the reference's fp64 never does
    // inside the clip window; keep |t| <= unit_clip so that P/t stays finite.
    if (sizeof(T) == 4) {
        if (t > Num<T>::unit_clip()) t = Num<T>::unit_clip();
        if (t < -Num<T>::unit_clip()) t = -Num<T>::unit_clip();
    }
    return t;
}

template <typename T>
__device__ __forceinline__ T clip_unit(T r)   // np.clip: NaN passes through
{
    if (r < -Num<T>::unit_clip()) return -Num<T>::unit_clip();
    if (r > Num<T>::unit_clip()) return Num<T>::unit_clip();
    return r;
}

// ---- LDPC_F32_FAST on the generic kernels: the same tanh-domain formulas with MUFU approximations ----
//   t = tanh(M/2) = sign(M) (1 - x) / (1 + x),  x = e^-|M| = 2^(-|M| log2 e)      (EX2 + RCP)
//   r = P / t                                                                          (RCP)
//   E = 2 atanh(r) = sign(r) ln2 lg2((1 + |r|) / (1 - |r|))                            (RCP + LG2)
// about 20 instead of 130 instructions per edge; absolute errors ~1e-7, saturation at |E| <= 17.3 like
// the accurate fp32 path (unit_clip = 1 - 2^-24).
__device__ __forceinline__ float mufu_ex2(float v) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float mufu_lg2(float v) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }
__device__ __forceinline__ float mufu_rcp(float v) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v)); return y; }

// TAB: fp64 coefficients from constant memory (unrolled rows) or as literals (rolled loops), see f64_math.cuh
template <bool FAST, typename T, bool TAB = true>
__device__ __forceinline__ T tanh_half(T msg)
{
    if constexpr (FAST) {
        const float x = mufu_ex2(-fminf(fabsf(msg), 35.0f) * 1.4426950408889634f);
        const float t = fminf((1.0f - x) * mufu_rcp(1.0f + x), Num<float>::unit_clip());
        return copysignf(t, msg);
    } else {
        return tanh_half_clipped<T, TAB>(msg);
    }
}

template <bool FAST, typename T>
__device__ __forceinline__ T quotient(T total, T tq)
{
    if constexpr (FAST) return total * mufu_rcp(tq);
#ifndef LDPC_F64_LIBM
    else if constexpr (sizeof(T) == 8) return f64::divide(total, tq);
#endif
    else return total / tq;
}

template <bool FAST, typename T>
__device__ __forceinline__ T two_atanh(T r)          // r already clipped to +-unit_clip
{
    if constexpr (FAST) {
        const float a = fabsf(r);
        return copysignf(0.6931471805599453f * mufu_lg2((1.0f + a) * mufu_rcp(1.0f - a)), r);
    } else {
        return T(2) * Num<T>::atanh_(r);
    }
}

template <bool FAST, typename T, bool TAB = true>
__device__ __forceinline__ T clipped_two_atanh(T r)  // E = 2 atanh(clip(r)), :167-168
{
#ifndef LDPC_F64_LIBM
    if constexpr (!FAST && sizeof(T) == 8) return f64::two_atanh<TAB>(r);      // clip included
#endif
    return two_atanh<FAST, T>(clip_unit<T>(r));
}

// Per-chunk bookkeeping living in the workspace.
struct ChunkState {
    int32_t* active[2];   // frame slots still decoding (double buffered)
    int32_t* count;       // [3] number of active slots, rotating: pass p reads slot p%3,
                          //     fills slot (p+1)%3 and clears slot (p+2)%3
    uint8_t* failed;      // [Fc] some check unsatisfied in this pass
    uint8_t* done;        // [Fc] frame finished (converged)
    uint8_t* keep;        // [Fc] slot stays in the active list (compaction only)
    int32_t* removed;     // [2] finished entries of the active list, by pass parity (compaction only)
    int32_t* norm_cnt;    // [Fc] sign changes of the metric in this pass
};

// llr [F][n] row-major  ->  lch [n][Fc] frame-minor (32x32 tiles through shared memory).
template <typename T>
__global__ void k_load_llr(const T* __restrict__ llr, int64_t f0, int64_t F, int n, int Fc,
                           T* __restrict__ lch)
{
    __shared__ T tile[32][33];
    const int tiles_n = (n + 31) / 32;
    const int64_t tiles = (int64_t)tiles_n * (Fc / 32);
    for (int64_t tix = blockIdx.x; tix < tiles; tix += gridDim.x) {
        const int tn = (int)(tix % tiles_n);
        const int tf = (int)(tix / tiles_n);
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int64_t f = f0 + (int64_t)tf * 32 + r;
            const int j = tn * 32 + threadIdx.x;
            tile[r][threadIdx.x] = (f < F && j < n) ? llr[f * n + j] : T(1);
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int j = tn * 32 + r;
            if (j < n) lch[(size_t)j * Fc + tf * 32 + threadIdx.x] = tile[threadIdx.x][r];
        }
        __syncthreads();
    }
}

__global__ void k_init_chunk(ChunkState st, int Fc, int64_t valid)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Fc; t += gridDim.x * blockDim.x) {
        st.active[0][t] = t;
        st.failed[t] = 0;
        st.done[t] = 0;
        st.norm_cnt[t] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st.count[0] = (int)valid;
        st.count[1] = 0;
        st.count[2] = 0;
        st.removed[0] = 0;
        st.removed[1] = 0;
    }
}

// Frame-block-major walk over (node, 32-frame block) pairs: warp w of the grid starts at pair w and
// advances by the number of warps in the grid, without a 64-bit division per item.
struct PairWalk {
    int node, tb, nodes, nblk, dq, dr;
    __device__ __forceinline__ PairWalk(int nodes_, int cpad)
    {
        nodes = nodes_;
        nblk = cpad >> 5;
        const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const unsigned stride = (gridDim.x * blockDim.x) >> 5;
        tb = (int)(warp / (unsigned)nodes);
        node = (int)(warp - (unsigned)tb * (unsigned)nodes);
        dq = (int)(stride / (unsigned)nodes);
        dr = (int)(stride - (unsigned)dq * (unsigned)nodes);
    }
    __device__ __forceinline__ bool valid() const { return tb < nblk; }
    __device__ __forceinline__ void next()
    {
        tb += dq;
        node += dr;
        if (node >= nodes) { node -= nodes; ++tb; }
    }
};

// Variable->check message of edge e (column j) for frame slot f: the channel value in the first pass
// (:88-96), afterwards posterior minus the check's own previous message (:260-268).
template <typename T>
__device__ __forceinline__ T v2c_message(const T* __restrict__ lch, const T* __restrict__ post,
                                         const T* __restrict__ Eold, int j, int e, int Fc, int f, int first_pass)
{
    if (first_pass) return lch[(size_t)j * Fc + f];
    return post[(size_t)j * Fc + f] - Eold[(size_t)e * Fc + f];
}

// Check-node pass.  One thread = (check i, frame slot t); a warp covers 32
// consecutive slots of one check, so the degree loop is warp-uniform.
// MAXD > 0: tanh values are kept in registers (degree <= MAXD);
// MAXD == 0: two sweeps over the row, tanh values parked in the output slots in between.
template <typename T, int MAXD, bool FAST>
__global__ void __launch_bounds__(kThreads)
k_check_nodes(int m, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
              const T* __restrict__ lch, const T* __restrict__ post, const T* __restrict__ Eold,
              T* __restrict__ E, int Fc,
              const int32_t* __restrict__ active, const int32_t* __restrict__ count_ptr,
              const uint8_t* __restrict__ done, int first_pass, int fix_odd)
{
    const int count = *count_ptr;
    if (count <= 0) return;
    const int cpad = (count + 31) & ~31;
    const int lane = threadIdx.x & 31;
    for (PairWalk pw(m, cpad); pw.valid(); pw.next()) {          // frame block major: see the header
        const int i = pw.node;
        const int t = pw.tb * 32 + lane;
        if (t >= count) continue;
        const int f = active[t];
        if (done[f]) continue;
        const int a = row_ptr[i], d = row_ptr[i + 1] - a;
        if (d == 0) continue;                                   // spa_decoder.py:115-122
        T total = T(1);
        if (MAXD > 0) {
            T tv[MAXD > 0 ? MAXD : 1];
#pragma unroll
            for (int q = 0; q < MAXD; ++q) {
                if (q < d) {
                    tv[q] = tanh_half<FAST, T>(v2c_message<T>(lch, post, Eold, col_idx[a + q], a + q, Fc, f, first_pass));
                    total *= tv[q];
                }
            }
#pragma unroll
            for (int q = 0; q < MAXD; ++q) {
                if (q < d) {
                    T r;
                    if (Num<T>::abs_(tv[q]) > Num<T>::small_tanh()) {
                        r = quotient<FAST, T>(total, tv[q]);
                    } else {
                        r = T(1);
#pragma unroll
                        for (int u = 0; u < MAXD; ++u)
                            if (u < d && u != q) r *= tv[u];
                    }
                    T e = clipped_two_atanh<FAST, T>(r);
                    if (fix_odd && (d & 1)) e = -e;
                    E[(size_t)(a + q) * Fc + f] = e;
                }
            }
        } else {
            // first sweep: park tanh(M/2) in this pass' message slot (E is double buffered, so the
            // previous messages stay intact); second sweep: read it back, divide, atanh, overwrite.
            for (int q = 0; q < d; ++q) {
                const T tq = tanh_half<FAST, T, false>(v2c_message<T>(lch, post, Eold, col_idx[a + q], a + q, Fc, f, first_pass));
                E[(size_t)(a + q) * Fc + f] = tq;
                total *= tq;
            }
            for (int q = 0; q < d; ++q) {
                const T tq = E[(size_t)(a + q) * Fc + f];
                T r;
                if (Num<T>::abs_(tq) > Num<T>::small_tanh()) {
                    r = quotient<FAST, T>(total, tq);
                } else {
                    r = T(1);
                    for (int u = 0; u < d; ++u) {
                        if (u == q) continue;
                        r *= tanh_half<FAST, T, false>(v2c_message<T>(lch, post, Eold, col_idx[a + u], a + u, Fc, f, first_pass));
                    }
                }
                T e = clipped_two_atanh<FAST, T, false>(r);
                if (fix_odd && (d & 1)) e = -e;
                E[(size_t)(a + q) * Fc + f] = e;
            }
        }
    }
}

// Check-node pass for SMALL batches (the per-frame SPA_Decoder.decode call): one warp = (check i,
// frame slot t) with the LANES spread over the edges of the row, instead of one thread walking the
// whole row for one of 32 frames.  tanh / division / atanh run lane-parallel; the product is still
// taken sequentially in edge order (by one lane, from the tanh values parked in shared memory), so the
// result is bit-identical to k_check_nodes.  Dynamic shared memory: warps per CTA x max degree values.
constexpr int kSmallWarps = 4;
template <typename T, bool FAST>
__global__ void __launch_bounds__(kSmallWarps * 32)
k_check_rows_small(int m, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                   const T* __restrict__ lch, const T* __restrict__ post, const T* __restrict__ Eold,
                   T* __restrict__ E, int Fc, const int32_t* __restrict__ active,
                   const int32_t* __restrict__ count_ptr, const uint8_t* __restrict__ done,
                   int first_pass, int fix_odd, int max_deg)
{
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int count = *count_ptr;
    if (count <= 0) return;
    const int lane = threadIdx.x & 31;
    T* tv = reinterpret_cast<T*>(small_smem) + (size_t)(threadIdx.x >> 5) * max_deg;
    const int64_t pairs = (int64_t)m * count;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < pairs; w += stride) {
        const int t = (int)(w / m);
        const int i = (int)(w - (int64_t)t * m);
        const int f = active[t];
        if (done[f]) continue;                                   // warp-uniform
        const int a = row_ptr[i], d = row_ptr[i + 1] - a;
        if (d == 0) continue;
        for (int q = lane; q < d; q += 32)
            tv[q] = tanh_half<FAST, T, false>(v2c_message<T>(lch, post, Eold, col_idx[a + q], a + q, Fc, f, first_pass));
        __syncwarp();
        T total = T(1);
        for (int q = 0; q < d; ++q) total *= tv[q];             // every lane: same order, same value
        for (int q = lane; q < d; q += 32) {
            const T tq = tv[q];
            T r;
            if (Num<T>::abs_(tq) > Num<T>::small_tanh()) {
                r = quotient<FAST, T>(total, tq);
            } else {
                r = T(1);
                for (int u = 0; u < d; ++u)
                    if (u != q) r *= tv[u];
            }
            T e = clipped_two_atanh<FAST, T, false>(r);
            if (fix_odd && (d & 1)) e = -e;
            E[(size_t)(a + q) * Fc + f] = e;
        }
        __syncwarp();                                            // tv is reused by the next pair
    }
}

// Posterior + hard decision (+ optional "normalized LLR" sign-change count over the
// first k_info bits, spa_decoder.py:210-228).  The variable->check messages of
// :260-268 are formed by the next check-node pass (v2c_message).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_var_nodes(int n, const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_edge,
            const T* __restrict__ lch, const T* __restrict__ E,
            T* __restrict__ post, uint8_t* __restrict__ zb, int Fc,
            const int32_t* __restrict__ active, const int32_t* __restrict__ count_ptr,
            const uint8_t* __restrict__ done, int first_pass, int k_norm, int32_t* __restrict__ norm_cnt)
{
    const int count = *count_ptr;
    if (count <= 0) return;
    const int cpad = (count + 31) & ~31;
    const int lane = threadIdx.x & 31;
    for (PairWalk pw(n, cpad); pw.valid(); pw.next()) {
        const int j = pw.node;
        const int t = pw.tb * 32 + lane;
        if (t >= count) continue;
        const int f = active[t];
        if (done[f]) continue;
        const int a = col_ptr[j], b = col_ptr[j + 1];
        T s = T(0);                                               // :177-182, ascending check order
        for (int q = a; q < b; ++q) s += E[(size_t)csc_edge[q] * Fc + f];
        const T ch = lch[(size_t)j * Fc + f];
        const T L = ch + s;                                       // :185
        if (k_norm > 0 && j < k_norm) {                           // :210-228
            const T prior = first_pass ? ch : post[(size_t)j * Fc + f];
            if (!(Num<T>::abs_(L) > T(7)) && prior * L < T(0)) atomicAdd(&norm_cnt[f], 1);
        }
        post[(size_t)j * Fc + f] = L;
        zb[(size_t)j * Fc + f] = (uint8_t)(L < T(0));             // :188
    }
}

// Small-batch twins of k_var_nodes / k_syndrome: one warp = (node, frame slot), lanes over the edges.
// The posterior sum keeps the ascending check order of :177-182 (values parked in shared memory, one
// lane adds them up), so it is bit-identical to k_var_nodes; the syndrome parity is order free.
template <typename T>
__global__ void __launch_bounds__(kSmallWarps * 32)
k_var_cols_small(int n, const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_edge,
                 const T* __restrict__ lch, const T* __restrict__ E, T* __restrict__ post,
                 uint8_t* __restrict__ zb, int Fc, const int32_t* __restrict__ active,
                 const int32_t* __restrict__ count_ptr, const uint8_t* __restrict__ done, int first_pass,
                 int k_norm, int32_t* __restrict__ norm_cnt, int max_deg)
{
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int count = *count_ptr;
    if (count <= 0) return;
    const int lane = threadIdx.x & 31;
    T* ev = reinterpret_cast<T*>(small_smem) + (size_t)(threadIdx.x >> 5) * max_deg;
    const int64_t pairs = (int64_t)n * count;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < pairs; w += stride) {
        const int t = (int)(w / n);
        const int j = (int)(w - (int64_t)t * n);
        const int f = active[t];
        if (done[f]) continue;
        const int a = col_ptr[j], d = col_ptr[j + 1] - a;
        for (int q = lane; q < d; q += 32) ev[q] = E[(size_t)csc_edge[a + q] * Fc + f];
        __syncwarp();
        if (lane == 0) {
            T s = T(0);
            for (int q = 0; q < d; ++q) s += ev[q];
            const T ch = lch[(size_t)j * Fc + f];
            const T L = ch + s;
            if (k_norm > 0 && j < k_norm) {
                const T prior = first_pass ? ch : post[(size_t)j * Fc + f];
                if (!(Num<T>::abs_(L) > T(7)) && prior * L < T(0)) atomicAdd(&norm_cnt[f], 1);
            }
            post[(size_t)j * Fc + f] = L;
            zb[(size_t)j * Fc + f] = (uint8_t)(L < T(0));
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kSmallWarps * 32)
k_syndrome_small(int m, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                 const uint8_t* __restrict__ zb, int Fc, const int32_t* __restrict__ active,
                 const int32_t* __restrict__ count_ptr, const uint8_t* __restrict__ done,
                 uint8_t* __restrict__ failed)
{
    const int count = *count_ptr;
    if (count <= 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t pairs = (int64_t)m * count;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < pairs; w += stride) {
        const int t = (int)(w / m);
        const int i = (int)(w - (int64_t)t * m);
        const int f = active[t];
        if (done[f]) continue;
        unsigned par = 0;
        for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32)
            par ^= (unsigned)(zb[(size_t)col_idx[e] * Fc + f] ^ 1u);
        const unsigned odd = __ballot_sync(0xffffffffu, par & 1u);
        if (lane == 0 && (__popc(odd) & 1)) failed[f] = 1;
    }
}

// Syndrome of the complemented decisions; a thread = (check, frame slot).
__global__ void __launch_bounds__(kThreads)
k_syndrome(int m, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
           const uint8_t* __restrict__ zb, int Fc, const int32_t* __restrict__ active,
           const int32_t* __restrict__ count_ptr, const uint8_t* __restrict__ done,
           uint8_t* __restrict__ failed)
{
    const int count = *count_ptr;
    if (count <= 0) return;
    const int cpad = (count + 31) & ~31;
    const int lane = threadIdx.x & 31;
    for (PairWalk pw(m, cpad); pw.valid(); pw.next()) {
        const int i = pw.node;
        const int t = pw.tb * 32 + lane;
        if (t >= count) continue;
        const int f = active[t];
        if (done[f]) continue;
        unsigned par = 0;
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e)
            par ^= (unsigned)(zb[(size_t)col_idx[e] * Fc + f] ^ 1u);   // :191-195
        if (par & 1u) failed[f] = 1;                                    // same value from every writer
    }
}

// End of a pass: record converged frames, rebuild the active list (warp ballot +
// popc, one atomic per warp), latch the metric of the exit pass.
__global__ void __launch_bounds__(kThreads)
k_finish_pass(ChunkState st, int pass, int last_pass, int early_term, int compact,
              int32_t* __restrict__ conv_it, uint8_t* __restrict__ ok, float* __restrict__ norm_out,
              int k_norm)
{
    const int count = st.count[pass % 3];
    const int32_t* cur = st.active[pass & 1];
    int32_t* nxt = st.active[(pass & 1) ^ 1];
    int32_t* next_count = st.count + (pass + 1) % 3;
    if (blockIdx.x == 0 && threadIdx.x == 0) st.count[(pass + 2) % 3] = 0;
    const unsigned lane = threadIdx.x & 31;
    const int cpad = (count + 31) & ~31;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cpad; t += gridDim.x * blockDim.x) {
        bool keep = false;
        if (t < count) {
            const int f = cur[t];
            if (!st.done[f]) {
                const bool conv = !st.failed[f];
                st.failed[f] = 0;
                const bool exit_now = (conv && (early_term || last_pass)) || last_pass;
                if (exit_now) {
                    if (conv) { conv_it[f] = pass; ok[f] = 1; }       // :231-241
                    if (norm_out) norm_out[f] = k_norm > 0 ? (float)st.norm_cnt[f] / (float)k_norm : 0.f;
                }
                if (conv && early_term) st.done[f] = 1;
                st.norm_cnt[f] = 0;
                keep = !(conv && early_term);
            }
            if (!compact) keep = true;
        }
        if (!compact) {
            // every slot keeps its place: appending warp segments in atomic order would shuffle them, and with a
            // frame count that is not a multiple of 32 the one short segment would shift all others off their
            // 128-byte lines (measured: 5x slower passes for 29 127 instead of 29 120 frames)
            if (t < count) nxt[t] = cur[t];
            if (t == 0) *next_count = count;
            continue;
        }
        // compaction: only mark here; k_compact_list rebuilds the list -- and leaves it untouched when no frame left
        // in this pass, because every rebuild appends the warps' survivors in atomic order and so moves frames
        // off the 128-byte lines of their neighbours
        if (t < count) st.keep[t] = keep ? 1 : 0;
        const unsigned gone = __ballot_sync(0xffffffffu, t < count && !keep);
        if (lane == 0 && gone) atomicAdd(st.removed + (pass & 1), __popc(gone));
    }
}

__global__ void __launch_bounds__(kThreads)
k_compact_list(ChunkState st, int pass)
{
    const int count = st.count[pass % 3];
    const int32_t* cur = st.active[pass & 1];
    int32_t* nxt = st.active[(pass & 1) ^ 1];
    int32_t* next_count = st.count + (pass + 1) % 3;
    // entries of the current list that are finished (this pass or, still masked in the list, earlier ones)
    const int removed = st.removed[pass & 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) st.removed[(pass & 1) ^ 1] = 0;     // last read in the previous pass
    // Rebuild only when at most a quarter of the list is still decoding.  Every rebuild scatters the survivors
    // (the lanes of a warp then read different 128-byte lines: up to 8x the sectors), which pays off only
    // against a matching cut in work; in between finished frames stay in the list and are masked.  Measured
    // (tools/compaction_probe.py): rebuilding in every pass is 2x slower than masking when 20-60 % of the
    // frames converge, 2x faster when nearly all do.
    const bool rebuild = removed > 0 && (int64_t)removed * 4 >= (int64_t)count * 3;
    const unsigned lane = threadIdx.x & 31;
    const int cpad = (count + 31) & ~31;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < cpad; t += gridDim.x * blockDim.x) {
        if (!rebuild) {
            if (t < count) nxt[t] = cur[t];
            if (t == 0) *next_count = count;
            continue;
        }
        const bool keep = t < count && st.keep[t];
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        int base = 0;
        if (lane == 0 && mask) base = atomicAdd(next_count, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) nxt[base + __popc(mask & ((1u << lane) - 1u))] = cur[t];
    }
}


// zb/post [n][Fc] frame-minor -> z_out/post_out [F][n] row-major.
template <typename T>
__global__ void k_store_out(const uint8_t* __restrict__ zb, const T* __restrict__ post, int n, int Fc,
                            int64_t f0, int64_t F, uint8_t* __restrict__ z_out, T* __restrict__ post_out)
{
    __shared__ T tp[32][33];
    __shared__ uint8_t tz[32][33];
    const int tiles_n = (n + 31) / 32;
    const int64_t tiles = (int64_t)tiles_n * (Fc / 32);
    for (int64_t tix = blockIdx.x; tix < tiles; tix += gridDim.x) {
        const int tn = (int)(tix % tiles_n);
        const int tf = (int)(tix / tiles_n);
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int j = tn * 32 + r;
            if (j < n) {
                tz[r][threadIdx.x] = zb[(size_t)j * Fc + tf * 32 + threadIdx.x];
                if (post_out) tp[r][threadIdx.x] = post[(size_t)j * Fc + tf * 32 + threadIdx.x];
            }
        }
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) {
            const int64_t f = f0 + (int64_t)tf * 32 + r;
            const int j = tn * 32 + threadIdx.x;
            if (f < F && j < n) {
                z_out[f * n + j] = tz[threadIdx.x][r];
                if (post_out) post_out[f * n + j] = tp[threadIdx.x][r];
            }
        }
        __syncthreads();
    }
}

__global__ void k_init_outputs(int64_t F, int32_t* conv_it, uint8_t* ok, float* norm)
{
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < F; f += (int64_t)gridDim.x * blockDim.x) {
        conv_it[f] = -1;      // spa_decoder.py:65
        ok[f] = 0;
        if (norm) norm[f] = 0.f;
    }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
size_t bytes_per_chunk(const ldpc_graph* g, int64_t Fc)
{
    size_t b = 0;
    b += align_up(sizeof(T) * (size_t)g->n * Fc, 256);        // lch
    b += align_up(sizeof(T) * (size_t)g->nnz * Fc, 256) * 2;  // E of the previous / current pass
    b += align_up(sizeof(T) * (size_t)g->n * Fc, 256);        // post
    b += align_up((size_t)g->n * Fc, 256);                    // zb
    b += align_up(sizeof(int32_t) * (size_t)Fc, 256) * 3;     // active x2, norm_cnt
    b += align_up((size_t)Fc, 256) * 3;                       // failed, done, keep
    b += 256;                                                 // counts
    return b;
}

template <typename T, bool FAST>
int decode_typed(const ldpc_graph* g, int64_t F, int max_iter, unsigned flags, const T* llr,
                 uint8_t* z_out, int32_t* conv_out, uint8_t* ok_out, T* post_out,
                 float* norm_out, int k_info, void* ws, size_t ws_bytes, cudaStream_t stream)
{
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    // chunk size: as many frames as fit, multiple of 32
    int64_t Fc = (F + 31) / 32 * 32;
    while (Fc > 32 && bytes_per_chunk<T>(g, Fc) > ws_bytes) {
        int64_t half = (Fc / 2 + 31) / 32 * 32;
        Fc = half < Fc ? half : Fc - 32;
    }
    if (bytes_per_chunk<T>(g, Fc) > ws_bytes || !ws) {
        set_error("workspace of %zu bytes cannot hold one 32-frame chunk (%zu needed)", ws_bytes,
                  bytes_per_chunk<T>(g, 32));
        return LDPC_ERR_WORKSPACE;
    }
    if (Fc > (int64_t)1 << 30) { set_error("chunk too large"); return LDPC_ERR_INVALID; }
    char* p = (char*)ws;
    auto take = [&](size_t bytes) { char* q = p; p += align_up(bytes, 256); return (void*)q; };
    T* lch = (T*)take(sizeof(T) * (size_t)g->n * Fc);
    T* Ebuf[2];                                   // messages of the previous / the current pass
    Ebuf[0] = (T*)take(sizeof(T) * (size_t)g->nnz * Fc);
    Ebuf[1] = (T*)take(sizeof(T) * (size_t)g->nnz * Fc);
    T* post = (T*)take(sizeof(T) * (size_t)g->n * Fc);
    uint8_t* zb = (uint8_t*)take((size_t)g->n * Fc);
    ChunkState st;
    st.active[0] = (int32_t*)take(sizeof(int32_t) * (size_t)Fc);
    st.active[1] = (int32_t*)take(sizeof(int32_t) * (size_t)Fc);
    st.norm_cnt = (int32_t*)take(sizeof(int32_t) * (size_t)Fc);
    st.failed = (uint8_t*)take((size_t)Fc);
    st.done = (uint8_t*)take((size_t)Fc);
    st.keep = (uint8_t*)take((size_t)Fc);
    st.count = (int32_t*)take(256);
    st.removed = st.count + 8;

    const int early = (flags & LDPC_FLAG_EARLY_TERM) ? 1 : 0;
    const bool want_compact = (flags & LDPC_FLAG_COMPACT) != 0;
    const int fix_odd = (flags & LDPC_FLAG_FIX_ODD_SIGN) ? 1 : 0;
    const int k_norm = norm_out ? k_info : 0;
    const int grid_cap = di.sm_count * 8;

    k_init_outputs<<<(int)std::min<int64_t>((F + 255) / 256, grid_cap), 256, 0, stream>>>(F, conv_out, ok_out, norm_out);
    LDPC_LAUNCH_CHECK();

    for (int64_t f0 = 0; f0 < F; f0 += Fc) {
        const int64_t valid = std::min<int64_t>(Fc, F - f0);
        const int Fci = (int)Fc;
        const int compact = want_compact && valid > 256 ? 1 : 0;      // two more launches per pass: not for a few frames
        {
            const int64_t tiles = (int64_t)((g->n + 31) / 32) * (Fc / 32);
            k_load_llr<T><<<(int)std::min<int64_t>(tiles, grid_cap * 4), dim3(32, 8), 0, stream>>>(llr, f0, F, g->n, Fci, lch);
            LDPC_LAUNCH_CHECK();
        }
        k_init_chunk<<<std::min((Fci + 255) / 256, grid_cap), 256, 0, stream>>>(st, Fci, valid);
        LDPC_LAUNCH_CHECK();
        // Few frames (the per-frame decode call): lanes across the row instead of across frames, when a row
        // has more edges than the chunk has frames to fill a warp with.
        const size_t small_smem_bytes = sizeof(T) * (size_t)kSmallWarps * g->max_cdeg;
        const int64_t avg_cdeg = g->nnz / std::max(1, g->m);
        const size_t small_vsmem_bytes = sizeof(T) * (size_t)kSmallWarps * std::max(1, g->max_vdeg);
        // measured on the dense H_std of WiMAX-576 (average check degree 143): lanes-across-edges wins up to ~300
        // frames (100 frames: 3.8 instead of 10.2 ms per Monte-Carlo interval), loses from ~500 on
        const bool small_rows = ((valid <= 32 && valid * 4 <= std::max<int64_t>(4, avg_cdeg)) || (avg_cdeg >= 64 && valid <= 256)) &&
                                small_smem_bytes <= (size_t)di.max_smem_optin && small_vsmem_bytes <= (size_t)di.max_smem_optin;
        const int small_vgrid = (int)std::min<int64_t>(((int64_t)g->n * valid + kSmallWarps - 1) / kSmallWarps, (int64_t)grid_cap * 4);
        if (small_rows && small_vsmem_bytes > 48 * 1024)
            LDPC_CUDA_TRY(cudaFuncSetAttribute(k_var_cols_small<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_vsmem_bytes));
        const int small_grid = (int)std::min<int64_t>(((int64_t)g->m * valid + kSmallWarps - 1) / kSmallWarps, (int64_t)grid_cap * 4);
        if (small_rows && small_smem_bytes > 48 * 1024)
            LDPC_CUDA_TRY(cudaFuncSetAttribute(k_check_rows_small<T, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes));
        const int64_t cn_items = (int64_t)g->m * Fc, vn_items = (int64_t)g->n * Fc;
        const int cn_grid = (int)std::min<int64_t>((cn_items + kThreads - 1) / kThreads, grid_cap);
        const int vn_grid = (int)std::min<int64_t>((vn_items + kThreads - 1) / kThreads, grid_cap);
        for (int it = 0; it < max_iter; ++it) {
            const int par = it & 1;
            const int32_t* cnt = st.count + it % 3;
            const int first = it == 0;
            const int last = it == max_iter - 1;
#define LDPC_CN(MAXD)                                                                               \
    k_check_nodes<T, MAXD, FAST><<<cn_grid, kThreads, 0, stream>>>(g->m, g->d_row_ptr, g->d_col_idx, lch, \
        post, Ebuf[par ^ 1], Ebuf[par], Fci, st.active[par], cnt, st.done, first, fix_odd)
            if (small_rows) {
                k_check_rows_small<T, FAST><<<small_grid, kSmallWarps * 32, small_smem_bytes, stream>>>(
                    g->m, g->d_row_ptr, g->d_col_idx, lch, post, Ebuf[par ^ 1], Ebuf[par], Fci, st.active[par], cnt,
                    st.done, first, fix_odd, g->max_cdeg);
            }
            else if (g->max_cdeg <= 8) LDPC_CN(8);
            else if (g->max_cdeg <= 24) LDPC_CN(24);
            else LDPC_CN(0);
#undef LDPC_CN
            LDPC_LAUNCH_CHECK();
            if (small_rows)
                k_var_cols_small<T><<<small_vgrid, kSmallWarps * 32, small_vsmem_bytes, stream>>>(
                    g->n, g->d_col_ptr, g->d_csc_edge, lch, Ebuf[par], post, zb, Fci, st.active[par], cnt, st.done, first,
                    k_norm, st.norm_cnt, g->max_vdeg);
            else
                k_var_nodes<T><<<vn_grid, kThreads, 0, stream>>>(g->n, g->d_col_ptr, g->d_csc_edge, lch, Ebuf[par], post,
                    zb, Fci, st.active[par], cnt, st.done, first, k_norm, st.norm_cnt);
            LDPC_LAUNCH_CHECK();
            // without early termination only the last pass needs a syndrome
            if (early || last) {
                if (small_rows)
                    k_syndrome_small<<<small_grid, kSmallWarps * 32, 0, stream>>>(g->m, g->d_row_ptr, g->d_col_idx, zb, Fci,
                        st.active[par], cnt, st.done, st.failed);
                else
                    k_syndrome<<<cn_grid, kThreads, 0, stream>>>(g->m, g->d_row_ptr, g->d_col_idx, zb, Fci,
                        st.active[par], cnt, st.done, st.failed);
                LDPC_LAUNCH_CHECK();
            }
            k_finish_pass<<<std::min((Fci + kThreads - 1) / kThreads, grid_cap), kThreads, 0, stream>>>(
                st, it, last, early, compact, conv_out + f0, ok_out + f0, norm_out ? norm_out + f0 : nullptr, k_norm);
            LDPC_LAUNCH_CHECK();
            if (compact) {
                k_compact_list<<<std::min((Fci + kThreads - 1) / kThreads, grid_cap), kThreads, 0, stream>>>(st, it);
                LDPC_LAUNCH_CHECK();
            }
        }
        const int64_t tiles = (int64_t)((g->n + 31) / 32) * (Fc / 32);
        k_store_out<T><<<(int)std::min<int64_t>(tiles, grid_cap * 4), dim3(32, 8), 0, stream>>>(
            zb, post, g->n, Fci, f0, F, z_out, post_out);
        LDPC_LAUNCH_CHECK();
    }
    return LDPC_OK;
}

}  // namespace

size_t generic_workspace_bytes(const ldpc_graph* g, int64_t frames, int dtype)
{
    const int64_t Fc = (frames + 31) / 32 * 32;
    return dtype == LDPC_F64 ? bytes_per_chunk<double>(g, Fc) : bytes_per_chunk<float>(g, Fc);
}

int generic_decode(const ldpc_graph* g, int dtype, int64_t frames, int max_iter, unsigned flags,
                   const void* llr_dev, uint8_t* z_dev, int32_t* conv_dev,
                   uint8_t* ok_dev, void* post_dev, float* norm_dev, int k_info, void* ws,
                   size_t ws_bytes, cudaStream_t stream)
{
    if (dtype == LDPC_F64)
        return decode_typed<double, false>(g, frames, max_iter, flags, (const double*)llr_dev, z_dev,
                                           conv_dev, ok_dev, (double*)post_dev, norm_dev, k_info, ws, ws_bytes, stream);
    // LDPC_F32_FAST: same kernels, MUFU arithmetic in the check node -- for sparse rows only.  On the dense rows
    // of an H_std graph (degree 100+) the approximation errors of the long product P and of P / t_j add up
    // (98.5 % instead of > 99.9 % of the decisions equal to the fp64 oracle), so those keep the accurate formulas.
    if (dtype == LDPC_F32_FAST && g->max_cdeg <= 24)
        return decode_typed<float, true>(g, frames, max_iter, flags, (const float*)llr_dev, z_dev,
                                         conv_dev, ok_dev, (float*)post_dev, norm_dev, k_info, ws, ws_bytes, stream);
    return decode_typed<float, false>(g, frames, max_iter, flags, (const float*)llr_dev, z_dev,
                                      conv_dev, ok_dev, (float*)post_dev, norm_dev, k_info, ws, ws_bytes, stream);
}

}  // namespace ldpc
