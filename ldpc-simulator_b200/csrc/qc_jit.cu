// qc_jit.cu -- run-time specialisation of the resident quasi-cyclic kernel.
//
// qc_kernel.cuh is specialised at COMPILE time on the base matrix of a code; the static registry
// (qc_registry.json -> qc_codes_gen.cuh) covers the codes the benchmarks name.  Every other
// quasi-cyclic graph (the reference's database holds 119 matrices, most of them QC) gets the same
// kernel here: the block rows are scheduled with the algorithm build_native.py uses, the
// Code<...> type is written out as text, and NVRTC compiles qc_kernel.cuh for sm_100a into a cubin
// that is loaded through the driver API.  Compiled kernels are cached per process (keyed by the
// base matrix) and on disk ($LDPC_JIT_CACHE, default ~/.cache/ldpc_b200), so a code costs one
// compilation (about 2 s) per machine.
//
// libnvrtc is opened with dlopen and the driver entry points come from cudaGetDriverEntryPoint:
// the library has no link-time dependency on either, loads on a machine without a GPU, and a
// machine without NVRTC still runs every registered code and the table-driven kernel.
#include "ldpc_common.cuh"
#include "qc_jit_src_gen.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>

namespace ldpc {

namespace {

// ---- schedule of the block rows (same algorithm as build_native.py: schedule_rows) -------------
struct Schedule {
    int teams = 1;
    int max_slot = 0;                       // most messages a thread keeps in registers
    int min_deg = 0, max_deg = 0;
    int edges = 0;                          // circulants of the base matrix
    std::vector<std::vector<int>> groups;   // rows of each group by team, -1 = none
    std::vector<char> first;                // [mb*nb] first row (in schedule order) touching the column block
};

Schedule schedule_rows(const QcInfo& qc)
{
    const int mb = qc.mb, nb = qc.nb;
    auto present = [&](int a, int c) { return qc.shift[(size_t)a * nb + c] >= 0; };
    std::vector<int> deg(mb, 0);
    for (int a = 0; a < mb; ++a)
        for (int c = 0; c < nb; ++c) deg[a] += present(a, c);
    std::vector<std::vector<char>> conflict(mb, std::vector<char>(mb, 0));
    std::vector<int> nconf(mb, 0);
    for (int a = 0; a < mb; ++a)
        for (int b = 0; b < mb; ++b) {
            if (a == b) continue;
            for (int c = 0; c < nb; ++c)
                if (present(a, c) && present(b, c)) { conflict[a][b] = 1; break; }
            nconf[a] += conflict[a][b];
        }
    // greedy colouring, most constrained row first
    std::vector<int> order(mb);
    for (int a = 0; a < mb; ++a) order[a] = a;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nconf[a] > nconf[b]; });
    std::vector<int> colour(mb, -1);
    int ncolours = 0;
    for (int a : order) {
        std::set<int> used;
        for (int b = 0; b < mb; ++b)
            if (conflict[a][b] && colour[b] >= 0) used.insert(colour[b]);
        int c = 0;
        while (used.count(c)) ++c;
        colour[a] = c;
        ncolours = std::max(ncolours, c + 1);
    }
    std::vector<std::vector<int>> classes(ncolours);
    for (int a = 0; a < mb; ++a) classes[colour[a]].push_back(a);
    Schedule s;
    size_t widest = 1;
    for (auto& v : classes) widest = std::max(widest, v.size());
    s.teams = (int)std::min<size_t>(4, widest);
    std::vector<int> load(s.teams, 0);
    for (auto& cls : classes) {
        std::vector<int> rows = cls;
        std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) { return deg[a] > deg[b]; });
        for (size_t i = 0; i < rows.size(); i += s.teams) {          // a colour wider than the CTA is split
            std::vector<int> placed(s.teams, -1);
            for (size_t j = i; j < std::min(rows.size(), i + s.teams); ++j) {
                int best = -1;                                       // heaviest row to the least loaded free team
                for (int t = 0; t < s.teams; ++t)
                    if (placed[t] < 0 && (best < 0 || load[t] < load[best])) best = t;
                placed[best] = rows[j];
                load[best] += deg[rows[j]];
            }
            s.groups.push_back(placed);
        }
    }
    s.first.assign((size_t)mb * nb, 0);
    std::vector<char> seen(nb, 0);
    for (auto& g : s.groups) {
        for (int a : g)
            if (a >= 0)
                for (int c = 0; c < nb; ++c)
                    if (present(a, c)) s.first[(size_t)a * nb + c] = !seen[c];
        for (int a : g)
            if (a >= 0)
                for (int c = 0; c < nb; ++c)
                    if (present(a, c)) seen[c] = 1;
    }
    s.max_slot = *std::max_element(load.begin(), load.end());
    s.min_deg = *std::min_element(deg.begin(), deg.end());
    s.max_deg = *std::max_element(deg.begin(), deg.end());
    for (int d : deg) s.edges += d;
    return s;
}

std::string code_type_text(const QcInfo& qc, const Schedule& s)
{
    std::string t = "Code<" + std::to_string(qc.z) + ", " + std::to_string(qc.z * qc.nb);
    for (auto& g : s.groups) {
        int last = (int)g.size() - 1;
        while (last >= 0 && g[last] < 0) --last;                     // trailing empty rows are dropped
        t += ",\n    Group<";
        for (int k = 0; k <= last; ++k) {
            if (k) t += ",\n          ";
            t += "Row<";
            if (g[k] >= 0) {
                bool any = false;
                for (int c = 0; c < qc.nb; ++c) {
                    const int sh = qc.shift[(size_t)g[k] * qc.nb + c];
                    if (sh < 0) continue;
                    if (any) t += ", ";
                    any = true;
                    t += "Slot<" + std::to_string(c) + ", " + std::to_string(sh) + ", " +
                         (s.first[(size_t)g[k] * qc.nb + c] ? "true" : "false") + ">";
                }
            }
            t += ">";
        }
        t += ">";
    }
    return t + ">";
}

// Host-side twin of qc::GatherShape (qc_kernel_gather.cuh): does the two-frames-per-thread gather kernel exist for this
// base matrix, and with which launch shape.
struct GatherHost {
    bool fits = false;
    int threads = 0, cols = 32, max_ctas = 1, minb = 1;
    size_t smem = 0;
};

// NVRTC's front end computes shared-memory addresses in 64 bits (nvcc's uses 32-bit pointers for the shared window):
// compiled from the plain source the gather kernel carries 0.2 address instructions per edge more than the nvcc build,
// its pass body (38 KB) no longer fits the instruction cache, and it is SLOWER than the one-frame kernel of the same
// module (z = 92: 6.57 against 7.37 Gbit/s).  kJitOptions therefore defines LDPC_SMEM_ASM: the two-frame kernels address
// the shared window explicitly (qc_kernel.cuh: sh_ld2 / sh_st2), which gives 2004 instead of 2428 instructions per pass
// and 8.41 Gbit/s (z = 48: 7.49 against 5.73-6.36; profiles/r2_jit_gather.jsonl).  LDPC_JIT_GATHER=0 keeps the module to
// the one-frame kernels (shorter compilation: 4 instead of 17 s per code, once per machine).
bool jit_gather_enabled()
{
    const char* e = getenv("LDPC_JIT_GATHER");
    return !(e && *e == '0');
}

GatherHost gather_shape(const QcInfo& qc, const Schedule& s)
{
    GatherHost h;
    const int tz = (qc.z + 31) / 32 * 32;
    h.threads = tz * s.teams;
    const int warps = h.threads / 32;
    const int own = (qc.nb + s.teams - 1) / s.teams;
    const int need = (warps + 3) / 4 * 32;
    h.cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    h.max_ctas = 512 / h.cols;
    h.smem = sizeof(float) * 2 * ((size_t)qc.z * qc.nb * 2 + (size_t)s.edges * qc.z);
    const int b0 = 65536 / (h.threads * 168);
    h.minb = std::min(std::max(b0, 1), h.max_ctas);
    h.fits = need <= 512 && s.max_deg <= 8 && 2 * own <= 32 && h.threads <= 1024 && h.smem + 2048 <= 227 * 1024 &&
             jit_gather_enabled();
    return h;
}

std::string jit_source(const std::string& code_type, bool with_gather)
{
    std::string src = with_gather ? "#include \"qc_kernel_gather.cuh\"\n" : "#include \"qc_kernel.cuh\"\n";
    src += "namespace ldpc { namespace qc {\nusing JitCode = " + code_type + ";\nusing JitShape = LaunchShape<JitCode>;\n} }\n";
    src += "#define LDPC_JIT_ARGS const float* __restrict__ llr, ldpc::qc::Outputs out, long long frames, int max_iter, \\\n"
           "    int fix_odd, ldpc::McParams mc, unsigned long long* __restrict__ work_counter\n"
           "extern \"C\" __global__ void __launch_bounds__(ldpc::qc::JitShape::THREADS, ldpc::qc::JitShape::MINB)\n"
           "ldpc_jit_fixed(LDPC_JIT_ARGS)\n"
           "{ ldpc::qc::decode_frames<ldpc::qc::JitShape::THREADS, false>(ldpc::qc::JitCode(), llr, out, frames, max_iter, fix_odd, mc, work_counter); }\n"
           "extern \"C\" __global__ void __launch_bounds__(ldpc::qc::JitShape::THREADS, ldpc::qc::JitShape::MINB)\n"
           "ldpc_jit_early(LDPC_JIT_ARGS)\n"
           "{ ldpc::qc::decode_frames<ldpc::qc::JitShape::THREADS, true>(ldpc::qc::JitCode(), llr, out, frames, max_iter, fix_odd, mc, work_counter); }\n";
    if (with_gather)      // two frames per thread, gather structure (qc_kernel_gather.cuh), fixed iteration count
        src += "using JitGather = ldpc::qc::GatherShape<ldpc::qc::JitCode>;\n"
               "extern \"C\" __global__ void __launch_bounds__(JitGather::THREADS, JitGather::MINB)\n"
               "ldpc_jit_gather(LDPC_JIT_ARGS)\n"
               "{ ldpc::qc::decode_gather<JitGather::THREADS, false>(ldpc::qc::JitCode(), llr, out, frames, max_iter, fix_odd, mc, work_counter); }\n";
    return src;
}

// ---- NVRTC through dlopen ----------------------------------------------------------------------
struct Nvrtc {
    void* handle = nullptr;
    std::string path, why;
    decltype(&nvrtcCreateProgram) createProgram = nullptr;
    decltype(&nvrtcCompileProgram) compileProgram = nullptr;
    decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) getCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) getCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) getProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) getProgramLog = nullptr;
    decltype(&nvrtcGetErrorString) getErrorString = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    int major = 0, minor = 0;
};

const Nvrtc& nvrtc()
{
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cand;
        if (const char* e = getenv("LDPC_NVRTC_LIB")) cand.push_back(e);
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                                 "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.13", "/usr/local/cuda/lib64/libnvrtc.so.13"})
            cand.push_back(name);
        for (auto& c : cand) {
            n.handle = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.handle) { n.path = c; break; }
        }
        if (!n.handle) { n.why = "libnvrtc not found (set LDPC_NVRTC_LIB)"; return; }
        bool all = true;
#define LDPC_SYM(field, sym) \
        n.field = (decltype(n.field))dlsym(n.handle, #sym); \
        all = all && n.field != nullptr
        LDPC_SYM(createProgram, nvrtcCreateProgram);
        LDPC_SYM(compileProgram, nvrtcCompileProgram);
        LDPC_SYM(destroyProgram, nvrtcDestroyProgram);
        LDPC_SYM(getCUBINSize, nvrtcGetCUBINSize);
        LDPC_SYM(getCUBIN, nvrtcGetCUBIN);
        LDPC_SYM(getProgramLogSize, nvrtcGetProgramLogSize);
        LDPC_SYM(getProgramLog, nvrtcGetProgramLog);
        LDPC_SYM(getErrorString, nvrtcGetErrorString);
        LDPC_SYM(version, nvrtcVersion);
#undef LDPC_SYM
        if (!all) { n.why = "libnvrtc lacks a required entry point"; dlclose(n.handle); n.handle = nullptr; return; }
        n.version(&n.major, &n.minor);
        if (n.major < 12 || (n.major == 12 && n.minor < 8)) {
            n.why = "NVRTC " + std::to_string(n.major) + "." + std::to_string(n.minor) + " cannot target sm_100a (needs 12.8+)";
            dlclose(n.handle); n.handle = nullptr;
        }
    });
    return n;
}

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull)
{
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

std::string cache_dir()
{
    if (const char* e = getenv("LDPC_JIT_CACHE")) return *e ? std::string(e) : std::string();   // empty = no disk cache
    if (const char* h = getenv("HOME")) return std::string(h) + "/.cache/ldpc_b200";
    return std::string();
}

bool read_file(const std::string& path, std::string* out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    out->clear();
    char buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, k);
    fclose(f);
    return !out->empty();
}

void write_file_atomic(const std::string& dir, const std::string& path, const std::string& data)
{
    std::string partial;
    for (size_t i = 1; i <= dir.size(); ++i)                     // mkdir -p
        if (i == dir.size() || dir[i] == '/') { partial = dir.substr(0, i); mkdir(partial.c_str(), 0755); }
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str()); else unlink(tmp.c_str());
}

const char* const kJitOptions[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DLDPC_SMEM_ASM"};

// source -> cubin (through the disk cache).  Host only: works without a GPU.
int compile_cubin(const std::string& src, std::string* cubin)
{
    const Nvrtc& n = nvrtc();
    if (!n.handle) { set_error("run-time specialisation unavailable: %s", n.why.c_str()); return LDPC_ERR_UNSUPPORTED; }
    std::string keytext = src + "\n//nvrtc " + std::to_string(n.major) + "." + std::to_string(n.minor);
    for (const char* o : kJitOptions) keytext += std::string(" ") + o;
    for (int i = 0; i < qc::kJitHeaderCount; ++i) keytext += qc::kJitHeaderText[i];
    char name[64];
    snprintf(name, sizeof name, "qc_%016llx.cubin", (unsigned long long)fnv1a(keytext));
    const std::string dir = cache_dir();
    const std::string path = dir.empty() ? std::string() : dir + "/" + name;
    if (!path.empty() && read_file(path, cubin)) return LDPC_OK;

    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.createProgram(&prog, src.c_str(), "ldpc_qc_jit.cu", qc::kJitHeaderCount, qc::kJitHeaderText, qc::kJitHeaderName);
    if (r != NVRTC_SUCCESS) { set_error("nvrtcCreateProgram: %s", n.getErrorString(r)); return LDPC_ERR_CUDA; }
    r = n.compileProgram(prog, (int)(sizeof kJitOptions / sizeof kJitOptions[0]), kJitOptions);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        n.getProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) n.getProgramLog(prog, &log[0]);
        if (log.size() > 1500) log = log.substr(0, 1500) + " ...";
        set_error("NVRTC compilation failed (%s): %s", n.getErrorString(r), log.c_str());
        n.destroyProgram(&prog);
        return LDPC_ERR_CUDA;
    }
    size_t sz = 0;
    r = n.getCUBINSize(prog, &sz);
    if (r == NVRTC_SUCCESS && sz) { cubin->assign(sz, '\0'); r = n.getCUBIN(prog, &(*cubin)[0]); }
    n.destroyProgram(&prog);
    if (r != NVRTC_SUCCESS || !sz) { set_error("NVRTC produced no cubin: %s", n.getErrorString(r)); return LDPC_ERR_CUDA; }
    if (!path.empty()) write_file_atomic(dir, path, *cubin);
    return LDPC_OK;
}

// ---- driver API through the runtime --------------------------------------------------------------
struct Driver {
    bool ok = false;
    decltype(&cuModuleLoadData) moduleLoadData = nullptr;
    decltype(&cuModuleGetFunction) moduleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) funcSetAttribute = nullptr;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) occupancy = nullptr;
    decltype(&cuLaunchKernel) launchKernel = nullptr;
    decltype(&cuGetErrorString) getErrorString = nullptr;
};

const Driver& driver()
{
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        bool all = cudaFree(nullptr) == cudaSuccess;       // primary context
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) all = false;
        };
        get("cuModuleLoadData", (void**)&d.moduleLoadData);
        get("cuModuleGetFunction", (void**)&d.moduleGetFunction);
        get("cuFuncSetAttribute", (void**)&d.funcSetAttribute);
        get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&d.occupancy);
        get("cuLaunchKernel", (void**)&d.launchKernel);
        get("cuGetErrorString", (void**)&d.getErrorString);
        cudaGetLastError();
        d.ok = all;
    });
    return d;
}

struct JitKernel {
    bool ok = false;
    std::string error;
    CUmodule module = nullptr;
    CUfunction fn[2] = {nullptr, nullptr};     // fixed iterations, early termination
    int per_sm[2] = {0, 0};
    int threads = 0;
    size_t smem = 0;
    // two frames per thread, gather structure (fixed iterations only), when the code fits it
    CUfunction gather = nullptr;
    int gather_per_sm = 0, gather_threads = 0;
    size_t gather_smem = 0;
};

std::mutex g_jit_mu;
std::map<std::string, std::shared_ptr<JitKernel>> g_jit_cache;

std::string cache_key(const QcInfo& qc)
{
    std::string k = std::to_string(qc.z) + "/" + std::to_string(qc.mb) + "/" + std::to_string(qc.nb) + ":";
    k.append((const char*)qc.shift.data(), qc.shift.size() * sizeof(int16_t));
    return k;
}

#define LDPC_DRV_TRY(k, expr)                                                  \
    do {                                                                       \
        CUresult _r = (expr);                                                  \
        if (_r != CUDA_SUCCESS) {                                              \
            const char* _s = nullptr;                                          \
            drv.getErrorString(_r, &_s);                                       \
            (k)->error = std::string(#expr " failed: ") + (_s ? _s : "?");     \
            return k;                                                          \
        }                                                                      \
    } while (0)

std::shared_ptr<JitKernel> build_kernel(const ldpc_graph* g)
{
    auto k = std::make_shared<JitKernel>();
    const Schedule s = schedule_rows(g->qc);
    std::string cubin;
    const GatherHost gh = gather_shape(g->qc, s);
    if (compile_cubin(jit_source(code_type_text(g->qc, s), gh.fits), &cubin) != LDPC_OK) { k->error = ldpc_last_error(); return k; }
    const Driver& drv = driver();
    if (!drv.ok) { k->error = "CUDA driver entry points unavailable"; return k; }
    k->threads = (g->qc.z + 31) / 32 * 32 * s.teams;
    k->smem = sizeof(float) * 4 * (size_t)g->n;
    LDPC_DRV_TRY(k, drv.moduleLoadData(&k->module, cubin.data()));
    LDPC_DRV_TRY(k, drv.moduleGetFunction(&k->fn[0], k->module, "ldpc_jit_fixed"));
    LDPC_DRV_TRY(k, drv.moduleGetFunction(&k->fn[1], k->module, "ldpc_jit_early"));
    for (int i = 0; i < 2; ++i) {
        LDPC_DRV_TRY(k, drv.funcSetAttribute(k->fn[i], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k->smem));
        LDPC_DRV_TRY(k, drv.funcSetAttribute(k->fn[i], CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100));
        LDPC_DRV_TRY(k, drv.occupancy(&k->per_sm[i], k->fn[i], k->threads, k->smem));
        if (k->per_sm[i] < 1) { k->error = "run-time specialised kernel does not fit on an SM"; return k; }
    }
    if (gh.fits) {
        DeviceInfo di;
        if (get_device_info(&di) != LDPC_OK) { k->error = ldpc_last_error(); return k; }
        // at most max_ctas CTAs may share an SM (tensor-memory columns): ask for more than 1/(max_ctas+1) of its shared
        // memory, as the static launcher does (spa_qc_spec.cu: launch_gather)
        k->gather_smem = std::max(gh.smem, (size_t)di.max_smem_optin / (gh.max_ctas + 1) + 1024);
        k->gather_threads = gh.threads;
        if (k->gather_smem <= (size_t)di.max_smem_optin) {
            LDPC_DRV_TRY(k, drv.moduleGetFunction(&k->gather, k->module, "ldpc_jit_gather"));
            LDPC_DRV_TRY(k, drv.funcSetAttribute(k->gather, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k->gather_smem));
            LDPC_DRV_TRY(k, drv.funcSetAttribute(k->gather, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100));
            int api = 0;
            LDPC_DRV_TRY(k, drv.occupancy(&api, k->gather, k->gather_threads, k->gather_smem));
            // the occupancy query answers for the default carve-out; the kernel is compiled for minb CTAs per SM
            // (launch bounds), and the whole shared memory of the SM is available to it
            const int by_smem = (int)((size_t)228 * 1024 / (k->gather_smem + 1024 + 256));
            k->gather_per_sm = std::min(gh.max_ctas, std::max(api, std::min(gh.minb, by_smem)));
            if (k->gather_per_sm < 1) k->gather = nullptr;
        }
    }
    k->ok = true;
    return k;
}

std::shared_ptr<JitKernel> get_kernel(const ldpc_graph* g)
{
    std::lock_guard<std::mutex> lk(g_jit_mu);
    // a loaded module belongs to one device's context: the handle's device is part of the key
    const std::string key = "dev" + std::to_string(g->device) + ":" + cache_key(g->qc);
    auto it = g_jit_cache.find(key);
    if (it != g_jit_cache.end()) return it->second;
    auto k = build_kernel(g);
    g_jit_cache[key] = k;            // failures are remembered too: one attempt per code and process
    return k;
}

}  // namespace

// Shape limits of the specialised kernel (qc_kernel.cuh).
const char* jit_shape_problem(const QcInfo& qc, const Schedule& s)
{
    if (qc.z < 1 || qc.z > 256) return "circulant size outside 1..256";
    if ((size_t)qc.z * qc.nb * 16 > 227 * 1024) return "four n-vectors do not fit the shared memory of an SM";
    std::vector<char> seen(qc.nb, 0);
    for (int b = 0; b < qc.mb; ++b)
        for (int c = 0; c < qc.nb; ++c)
            if (qc.shift[(size_t)b * qc.nb + c] >= 0) seen[c] = 1;
    for (char v : seen) if (!v) return "a column block has no check";
    if (s.min_deg < 2) return "a check of degree < 2";
    if (s.max_slot > 96) return "more than 96 messages per thread";       // registers: 2 per message + working set
    if ((qc.z + 31) / 32 * 32 * s.teams > 1024) return "CTA larger than 1024 threads";
    return nullptr;
}

// ... plus a usable NVRTC.
bool qc_jit_supported(const ldpc_graph* g)
{
    if (!g || !g->is_qc) return false;
    if (jit_shape_problem(g->qc, schedule_rows(g->qc))) return false;
    return nvrtc().handle != nullptr;
}

int qc_jit_prepare(const ldpc_graph* g)
{
    auto k = get_kernel(g);
    if (!k->ok) { set_error("%s", k->error.c_str()); return LDPC_ERR_UNSUPPORTED; }
    return LDPC_OK;
}

int qc_jit_decode(const ldpc_graph* g, int64_t frames, int max_iter, unsigned flags,
                  const float* llr_dev, uint8_t* z_dev, uint32_t* zbits_dev, int32_t* conv_dev,
                  uint8_t* ok_dev, float* post_dev, const McParams& mc_in, void* ws, cudaStream_t stream)
{
    auto k = get_kernel(g);
    if (!k->ok) { set_error("%s", k->error.c_str()); return LDPC_ERR_UNSUPPORTED; }
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    const int early = (flags & LDPC_FLAG_EARLY_TERM) ? 1 : 0;
    unsigned long long* counter = nullptr;
    if (early) {
        if (!ws) { set_error("early termination needs a workspace (work counter)"); return LDPC_ERR_WORKSPACE; }
        counter = (unsigned long long*)ws;
        LDPC_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
    }
    // same choice as the static launcher (spa_qc_spec.cu: launch_code): fixed iteration count and enough frames to fill
    // the machine with CTAs of two frames -> the gather kernel; early termination or a small batch -> one frame per CTA
    const bool use_gather = k->gather && !early && !(flags & LDPC_FLAG_ONE_FRAME) && frames >= 4 * (int64_t)di.sm_count;
    const int grid = use_gather ? (int)std::min<int64_t>((frames + 1) / 2, (int64_t)k->gather_per_sm * di.sm_count)
                                : (int)std::min<int64_t>(frames, (int64_t)k->per_sm[early] * di.sm_count);
    qc::Outputs out{z_dev, zbits_dev, conv_dev, ok_dev, post_dev};
    long long nframes = frames;
    int fix_odd = (flags & LDPC_FLAG_FIX_ODD_SIGN) ? 1 : 0;
    McParams mc = mc_in;
    void* args[] = {&llr_dev, &out, &nframes, &max_iter, &fix_odd, &mc, &counter};
    const Driver& drv = driver();
    CUresult r = use_gather
        ? drv.launchKernel(k->gather, grid, 1, 1, k->gather_threads, 1, 1, (unsigned)k->gather_smem, (CUstream)stream, args, nullptr)
        : drv.launchKernel(k->fn[early], grid, 1, 1, k->threads, 1, 1, (unsigned)k->smem, (CUstream)stream, args, nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (r != CUDA_SUCCESS) {
        const char* s = nullptr;
        drv.getErrorString(r, &s);
        set_error("launch of the run-time specialised kernel failed: %s", s ? s : "?");
        return LDPC_ERR_CUDA;
    }
    return LDPC_OK;
}

}  // namespace ldpc

// Host-only view of the specialisation (no GPU needed): the Code<...> type the scheduler derives
// from a base matrix and, optionally, the size of the cubin NVRTC builds from it.
extern "C" int ldpc_host_jit_compile(int z, int mb, int nb, const int16_t* shift, char* type_text, size_t type_cap,
                                     size_t* cubin_bytes)
{
    using namespace ldpc;
    if (z < 1 || mb < 1 || nb < 1 || !shift) { set_error("bad base matrix"); return LDPC_ERR_INVALID; }
    QcInfo qc;
    qc.z = z; qc.mb = mb; qc.nb = nb;
    qc.shift.assign(shift, shift + (size_t)mb * nb);
    for (int16_t s : qc.shift)
        if (s < -1 || s >= z) { set_error("shift %d out of range for z=%d", (int)s, z); return LDPC_ERR_INVALID; }
    const Schedule s = schedule_rows(qc);
    const std::string text = code_type_text(qc, s);
    if (type_text && type_cap) {
        if (text.size() + 1 > type_cap) { set_error("type text needs %zu bytes", text.size() + 1); return LDPC_ERR_INVALID; }
        memcpy(type_text, text.c_str(), text.size() + 1);
    }
    if (cubin_bytes) {
        if (const char* why = jit_shape_problem(qc, s)) { set_error("no specialised kernel for this base matrix: %s", why); return LDPC_ERR_UNSUPPORTED; }
        std::string cubin;
        int rc = compile_cubin(jit_source(text, gather_shape(qc, s).fits), &cubin);
        if (rc) return rc;
        *cubin_bytes = cubin.size();
    }
    return LDPC_OK;
}
