// graph.cu -- host-side graph analysis and the ldpc_graph handle.
//
// Replaces, for the SPA decode path of omkuprin7/ldpc-simulator:
//   * SPA_Decoder.__init__/_init_neighbor_structures (spa_decoder.py:16-61) and
//     EncoderDecoderData._init_decoder_structures (encoder_decoder_data.py:718-747):
//     per-decoder Python dicts become one immutable device edge index;
//   * gaussian_elimination + create_standart_parity_check_matrix
//     (encoder_decoder_data.py:13-183, 269-317): dict-of-rows GF(2) elimination
//     becomes a bit-packed Gauss-Jordan with the same pivot rule;
//   * (new) quasi-cyclic structure detection for the resident kernel.
#include "ldpc_common.cuh"

#include <algorithm>
#include <mutex>
#include <numeric>

namespace ldpc {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int get_device_info(DeviceInfo* out)
{
    static DeviceInfo cached;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device available: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return LDPC_ERR_CUDA;
    }
    if (!cached.ok || cached.device != dev) {
        cudaDeviceProp p;
        e = cudaGetDeviceProperties(&p, dev);
        if (e != cudaSuccess) {
            set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
            return LDPC_ERR_CUDA;
        }
        cached.device = dev;
        cached.sm_count = p.multiProcessorCount;
        cached.max_smem_optin = (int)p.sharedMemPerBlockOptin;
        cached.ok = true;
    }
    *out = cached;
    return LDPC_OK;
}

static int validate_csr(int m, int n, const int32_t* rp, const int32_t* ci)
{
    if (m <= 0 || n <= 0 || !rp || !ci) {
        set_error("empty parity-check matrix (m=%d, n=%d)", m, n);
        return LDPC_ERR_INVALID;
    }
    if (rp[0] != 0) { set_error("row_ptr[0] must be 0"); return LDPC_ERR_INVALID; }
    for (int i = 0; i < m; ++i) {
        if (rp[i + 1] < rp[i]) { set_error("row_ptr not monotone at row %d", i); return LDPC_ERR_INVALID; }
        for (int32_t e = rp[i]; e < rp[i + 1]; ++e) {
            if (ci[e] < 0 || ci[e] >= n) { set_error("column %d out of range in row %d", ci[e], i); return LDPC_ERR_INVALID; }
            if (e > rp[i] && ci[e] <= ci[e - 1]) {
                set_error("columns of row %d are not strictly ascending", i);
                return LDPC_ERR_INVALID;
            }
        }
    }
    return LDPC_OK;
}

static void edge_index(int m, int n, const int32_t* rp, const int32_t* ci, int32_t* cp, int32_t* ce,
                       int32_t* er)
{
    const int64_t nnz = rp[m];
    std::fill(cp, cp + n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) cp[ci[e] + 1]++;
    for (int j = 0; j < n; ++j) cp[j + 1] += cp[j];
    std::vector<int32_t> cursor(cp, cp + n);
    for (int i = 0; i < m; ++i)
        for (int32_t e = rp[i]; e < rp[i + 1]; ++e) {
            ce[cursor[ci[e]]++] = e;      // rows visited in ascending order => ascending row per column
            if (er) er[e] = i;
        }
}

// Largest z > 1 for which H splits into z x z blocks that are zero or a single
// cyclic shift of the identity.  O(nnz) per candidate divisor of gcd(m, n).
static bool detect_qc(int m, int n, const int32_t* rp, const int32_t* ci, QcInfo* out)
{
    int g = std::gcd(m, n);
    std::vector<int> cand;
    for (int z = g; z >= 2; --z)
        if (g % z == 0) cand.push_back(z);
    for (int z : cand) {
        const int mb = m / z, nb = n / z;
        if ((int64_t)mb * nb > (int64_t)1 << 24) continue;
        std::vector<int16_t> sh((size_t)mb * nb, -1);
        std::vector<int32_t> cnt((size_t)mb * nb, 0);
        bool ok = z <= 32767;
        for (int r = 0; r < m && ok; ++r) {
            const int br = r / z, lr = r % z;
            for (int32_t e = rp[r]; e < rp[r + 1]; ++e) {
                const int bc = ci[e] / z, lc = ci[e] % z;
                const int s = (lc - lr + z) % z;
                const size_t b = (size_t)br * nb + bc;
                if (cnt[b] == 0) sh[b] = (int16_t)s;
                else if (sh[b] != s) { ok = false; break; }
                cnt[b]++;
            }
        }
        for (size_t b = 0; ok && b < cnt.size(); ++b)
            if (cnt[b] != 0 && cnt[b] != z) ok = false;
        if (ok) {
            out->z = z; out->mb = mb; out->nb = nb; out->shift.swap(sh);
            return true;
        }
    }
    return false;
}

}  // namespace ldpc

using namespace ldpc;

// ---------------------------------------------------------------------------
extern "C" const char* ldpc_last_error(void) { return t_err; }
extern "C" int ldpc_abi_version(void) { return LDPC_B200_ABI_VERSION; }
extern "C" uint64_t ldpc_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int ldpc_host_edge_index(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                                    int32_t* col_ptr, int32_t* csc_edge, int32_t* edge_row)
{
    int rc = validate_csr(m, n, row_ptr, col_idx);
    if (rc) return rc;
    if (!col_ptr || !csc_edge) { set_error("null output"); return LDPC_ERR_INVALID; }
    edge_index(m, n, row_ptr, col_idx, col_ptr, csc_edge, edge_row);
    return LDPC_OK;
}

extern "C" int ldpc_host_detect_qc(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                                   int* z, int* mb, int* nb, int16_t* shift, int64_t shift_cap)
{
    int rc = validate_csr(m, n, row_ptr, col_idx);
    if (rc) return rc;
    QcInfo q;
    if (!detect_qc(m, n, row_ptr, col_idx, &q)) return 0;
    if (z) *z = q.z;
    if (mb) *mb = q.mb;
    if (nb) *nb = q.nb;
    if (shift) {
        if (shift_cap < (int64_t)q.shift.size()) { set_error("shift table needs %zu entries", q.shift.size()); return LDPC_ERR_INVALID; }
        std::copy(q.shift.begin(), q.shift.end(), shift);
    }
    return 1;
}

// Bit-packed GF(2) Gauss-Jordan.  The reduced row-echelon form is unique, and
// with the pivot rule "columns left to right, first row at or below the cursor"
// (encoder_decoder_data.py:37-54) so is the pivot list, hence H_std and the
// permutation equal the reference's bit for bit.
extern "C" int ldpc_host_standard_form(int m, int n, const int32_t* row_ptr, const int32_t* col_idx,
                                       uint64_t* h_std_bits, int32_t* perm, int32_t* rank)
{
    int rc = validate_csr(m, n, row_ptr, col_idx);
    if (rc) return rc;
    if (!h_std_bits || !perm || !rank) { set_error("null output"); return LDPC_ERR_INVALID; }
    const int W = (n + 63) / 64;
    std::vector<uint64_t> a((size_t)m * W, 0);
    for (int i = 0; i < m; ++i)
        for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e)
            a[(size_t)i * W + (col_idx[e] >> 6)] ^= (uint64_t)1 << (col_idx[e] & 63);
    std::vector<int32_t> piv;
    piv.reserve(m);
    int cur = 0;
    for (int c = 0; c < n && cur < m; ++c) {
        const int w = c >> 6;
        const uint64_t bit = (uint64_t)1 << (c & 63);
        int p = -1;
        for (int r = cur; r < m; ++r)
            if (a[(size_t)r * W + w] & bit) { p = r; break; }
        if (p < 0) continue;
        if (p != cur) std::swap_ranges(a.begin() + (size_t)p * W, a.begin() + (size_t)(p + 1) * W, a.begin() + (size_t)cur * W);
        const uint64_t* src = &a[(size_t)cur * W];
        for (int r = 0; r < m; ++r) {
            if (r == cur) continue;
            uint64_t* dst = &a[(size_t)r * W];
            if (dst[w] & bit)
                for (int x = 0; x < W; ++x) dst[x] ^= src[x];
        }
        piv.push_back(c);
        ++cur;
    }
    const int r = (int)piv.size();
    std::vector<uint8_t> is_piv(n, 0);
    for (int c : piv) is_piv[c] = 1;
    int q = 0;
    for (int c = 0; c < n; ++c)
        if (!is_piv[c]) perm[q++] = c;             // encoder_decoder_data.py:307-311
    for (int c : piv) perm[q++] = c;               // :313
    std::fill(h_std_bits, h_std_bits + (size_t)m * W, 0);
    for (int i = 0; i < r; ++i) {
        const uint64_t* src = &a[(size_t)i * W];
        uint64_t* dst = h_std_bits + (size_t)i * W;
        for (int c = 0; c < n; ++c) {
            const int o = perm[c];
            if ((src[o >> 6] >> (o & 63)) & 1) dst[c >> 6] |= (uint64_t)1 << (c & 63);
        }
    }
    *rank = r;
    return LDPC_OK;
}

// ---------------------------------------------------------------------------
static int upload(const std::vector<int32_t>& h, int32_t** d)
{
    size_t bytes = sizeof(int32_t) * std::max<size_t>(h.size(), 1);
    LDPC_CUDA_TRY(cudaMalloc((void**)d, bytes));
    if (!h.empty()) LDPC_CUDA_TRY(cudaMemcpy(*d, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
    return LDPC_OK;
}

extern "C" int ldpc_graph_create_csr(int m, int n, int64_t nnz, const int32_t* row_ptr,
                                     const int32_t* col_idx, ldpc_graph** out)
{
    if (!out) { set_error("null out"); return LDPC_ERR_INVALID; }
    *out = nullptr;
    int rc = validate_csr(m, n, row_ptr, col_idx);
    if (rc) return rc;
    if (row_ptr[m] != nnz) { set_error("nnz (%lld) != row_ptr[m] (%d)", (long long)nnz, row_ptr[m]); return LDPC_ERR_INVALID; }
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc) return rc;
    static std::atomic<uint64_t> next_serial{1};
    ldpc_graph* g = new (std::nothrow) ldpc_graph();
    if (g) g->serial = next_serial.fetch_add(1);
    if (!g) { set_error("out of host memory"); return LDPC_ERR_NOMEM; }
    g->m = m; g->n = n; g->nnz = nnz; g->device = di.device;
    g->row_ptr.assign(row_ptr, row_ptr + m + 1);
    g->col_idx.assign(col_idx, col_idx + nnz);
    g->col_ptr.resize(n + 1);
    g->csc_edge.resize(nnz);
    g->edge_row.resize(nnz);
    edge_index(m, n, row_ptr, col_idx, g->col_ptr.data(), g->csc_edge.data(), g->edge_row.data());
    for (int i = 0; i < m; ++i) g->max_cdeg = std::max(g->max_cdeg, row_ptr[i + 1] - row_ptr[i]);
    for (int j = 0; j < n; ++j) g->max_vdeg = std::max(g->max_vdeg, g->col_ptr[j + 1] - g->col_ptr[j]);
    g->is_qc = detect_qc(m, n, row_ptr, col_idx, &g->qc);
    rc = upload(g->row_ptr, &g->d_row_ptr);
    if (!rc) rc = upload(g->col_idx, &g->d_col_idx);
    if (!rc) rc = upload(g->col_ptr, &g->d_col_ptr);
    if (!rc) rc = upload(g->csc_edge, &g->d_csc_edge);
    if (rc) { ldpc_graph_destroy(g); return rc; }
    *out = g;
    return LDPC_OK;
}

extern "C" int ldpc_graph_create_qc(int z, int mb, int nb, const int16_t* shift, ldpc_graph** out)
{
    if (!out) { set_error("null out"); return LDPC_ERR_INVALID; }
    *out = nullptr;
    if (z <= 0 || mb <= 0 || nb <= 0 || !shift || (int64_t)z * nb > INT32_MAX || (int64_t)z * mb > INT32_MAX) {
        set_error("bad quasi-cyclic description (z=%d, mb=%d, nb=%d)", z, mb, nb);
        return LDPC_ERR_INVALID;
    }
    const int m = z * mb, n = z * nb;
    std::vector<int32_t> rp(m + 1, 0), ci;
    std::vector<int32_t> cols;
    for (int br = 0; br < mb; ++br)
        for (int lr = 0; lr < z; ++lr) {
            cols.clear();
            for (int bc = 0; bc < nb; ++bc) {
                int s = shift[(size_t)br * nb + bc];
                if (s < 0) continue;
                if (s >= z) { set_error("shift %d >= z=%d", s, z); return LDPC_ERR_INVALID; }
                cols.push_back(bc * z + (lr + s) % z);
            }
            std::sort(cols.begin(), cols.end());
            ci.insert(ci.end(), cols.begin(), cols.end());
            rp[br * z + lr + 1] = (int32_t)ci.size();
        }
    return ldpc_graph_create_csr(m, n, (int64_t)ci.size(), rp.data(), ci.data(), out);
}

extern "C" int ldpc_graph_info(const ldpc_graph* g, int* m, int* n, int64_t* nnz, int* max_check_degree,
                               int* max_var_degree, int* qc_z, int* qc_mb, int* qc_nb)
{
    if (!g) { set_error("null graph"); return LDPC_ERR_INVALID; }
    if (m) *m = g->m;
    if (n) *n = g->n;
    if (nnz) *nnz = g->nnz;
    if (max_check_degree) *max_check_degree = g->max_cdeg;
    if (max_var_degree) *max_var_degree = g->max_vdeg;
    if (qc_z) *qc_z = g->is_qc ? g->qc.z : 0;
    if (qc_mb) *qc_mb = g->is_qc ? g->qc.mb : 0;
    if (qc_nb) *qc_nb = g->is_qc ? g->qc.nb : 0;
    return LDPC_OK;
}

extern "C" int ldpc_graph_qc_shifts(const ldpc_graph* g, int16_t* shift, int64_t shift_cap)
{
    if (!g) { set_error("null graph"); return LDPC_ERR_INVALID; }
    if (!g->is_qc) return 0;
    if (!shift || shift_cap < (int64_t)g->qc.shift.size()) { set_error("shift table needs %zu entries", g->qc.shift.size()); return LDPC_ERR_INVALID; }
    std::copy(g->qc.shift.begin(), g->qc.shift.end(), shift);
    return 1;
}

extern "C" void ldpc_graph_destroy(ldpc_graph* g)
{
    if (!g) return;
    qc_resident_release(g);
    cudaFree(g->d_row_ptr);
    cudaFree(g->d_col_idx);
    cudaFree(g->d_col_ptr);
    cudaFree(g->d_csc_edge);
    delete g;
}
