// C-ABI layer of libopensc2_b200.so (see include/opensc2_b200.h).
// Host buffers in the reference's per-conductor layout go through one staging
// buffer and a tiled transpose into the batch-innermost device layout.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "step_launch.h"
#include <cmath>
#include "props_kernels.cuh"
#include "tcs_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(OSC2_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

struct EventPair { cudaEvent_t a, b; int kid; };

}  // namespace

struct osc2_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    osc2_topology topo;
    osc2_topology *topo_dev = nullptr;
    int B = 0, N = 0, E = 0, M = 0, S = 0, d = 0, NS = 0;
    size_t total = 0;
    StepDev p;                       // device pointers
    std::vector<void *> allocs;
    size_t bytes = 0;
    double *staging = nullptr;       // [max array] for layout changes
    size_t staging_doubles = 0;
    double *nodal = nullptr;         // [OSC2_N_NODAL][M][N][B], allocated by the first osc2_eval_nodal_properties
    int *cb_nodal = nullptr;         // [M][B] array-wide Colebrook iteration count of the nodal pass
    double *sysvar_a = nullptr, *sysvar_b = nullptr;
    bool have_state = false, have_inputs = false;
    bool have_geometry = false;      // dz, fl_bc, contact perimeters, qsource: never computed on the device
    bool have_peri_ss_nodal = false;
    bool capture = false, profiling = false;
    bool fast = false;               // register-resident kernels (step_kernels_fast.cuh)
    int max_smem = 0, sm_count = 0;
    std::vector<EventPair> pending, pool;
    double kernel_ms[OSC2_K_COUNT] = {0};
    long long kernel_launches[OSC2_K_COUNT] = {0};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int *status_host = nullptr;
    std::string launch_error;
    // device-side operating conditions (props_kernels.cuh)
    osc2_model *model_dev = nullptr;
    PropsDev q;
    bool have_model = false, have_sweep = false;
    bool need_peri_ss_nodal = false; // some |ss_choice| == 3
    double *tcs_nodal = nullptr;     // [S][N][B] (osc2_eval_tcs)
    double *margin_out = nullptr;    // [S][B] decoded temperature margins (osc2_download_margin)
    bool have_table[OSC2_MAX_FLUIDS] = {false, false, false, false};
    int n_fluids_used = 0;

    // ---- Launcher policy for osc2_run_step --------------------------------
    size_t get_max_smem() const { return (size_t)max_smem; }
};

namespace {

struct CudaLauncher {
    osc2_context *h;
    size_t max_smem() const { return (size_t)h->max_smem; }
    int sm_count() const { return h->sm_count; }

    template <class K, class... A>
    void launch(int kid, K kernel, unsigned grid, unsigned block, size_t smem, A... args)
    {
        if (!h->launch_error.empty()) return;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { h->launch_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return; }
        }
        if (kid == OSC2_K_FORWARD && block == 64)   // split K_F: 7 CTAs of ~26 KB per SM need a >= 180 KB carve-out
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 80);
        EventPair ev{};
        if (h->profiling) {
            if (!h->pool.empty()) { ev = h->pool.back(); h->pool.pop_back(); }
            else { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
            ev.kid = kid;
            cudaEventRecord(ev.a, h->stream);
        }
        kernel<<<grid, block, smem, h->stream>>>(args...);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) h->launch_error = std::string("kernel launch: ") + cudaGetErrorString(e);
        if (h->profiling) { cudaEventRecord(ev.b, h->stream); h->pending.push_back(ev); }
        h->kernel_launches[kid]++;
    }
};

int dev_alloc(osc2_context *h, double **ptr, size_t n_doubles)
{
    if (n_doubles == 0) n_doubles = 1;
    void *q = nullptr;
    CK(cudaMalloc(&q, n_doubles * sizeof(double)));
    h->allocs.push_back(q);
    h->bytes += n_doubles * sizeof(double);
    *ptr = (double *)q;
    return OSC2_OK;
}

// host [B][K] -> device [K][B]
int upload_transposed(osc2_context *h, const double *host, double *dev, size_t K)
{
    const size_t n = (size_t)h->B * K;
    if (n == 0) return OSC2_OK;
    if (n > h->staging_doubles) return fail(OSC2_ERR_STATE, "staging buffer too small");
    CK(cudaMemcpyAsync(h->staging, host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CudaLauncher L{h};
    const long long tiles = (((long long)h->B + 31) / 32) * (((long long)K + 31) / 32);
    L.launch(OSC2_K_TRANSPOSE, osc2_transpose_kernel, (unsigned)tiles, 256u, (size_t)0,
             (const double *)h->staging, dev, (long long)h->B, (long long)K);
    if (!h->launch_error.empty()) return fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
    return OSC2_OK;
}

// device [K][B] -> host [B][K]
int download_transposed(osc2_context *h, const double *dev, double *host, size_t K)
{
    const size_t n = (size_t)h->B * K;
    if (n == 0) return OSC2_OK;
    if (n > h->staging_doubles) return fail(OSC2_ERR_STATE, "staging buffer too small");
    CudaLauncher L{h};
    const long long tiles = (((long long)K + 31) / 32) * (((long long)h->B + 31) / 32);
    L.launch(OSC2_K_TRANSPOSE, osc2_transpose_kernel, (unsigned)tiles, 256u, (size_t)0,
             dev, h->staging, (long long)K, (long long)h->B);
    if (!h->launch_error.empty()) return fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
    CK(cudaMemcpyAsync(host, h->staging, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OSC2_OK;
}

int drain_events(osc2_context *h)
{
    for (auto &ev : h->pending) {
        CK(cudaEventSynchronize(ev.b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
        h->kernel_ms[ev.kid] += ms;
        h->pool.push_back(ev);
    }
    h->pending.clear();
    return OSC2_OK;
}

int validate_topology(const osc2_topology *t)
{
    if (!t) return fail(OSC2_ERR_INVALID, "topology is NULL");
    if (t->n_ch < 0 || t->n_ch > OSC2_MAX_CHANNELS || t->n_sol < 0 || t->n_sol > OSC2_MAX_SOLIDS)
        return fail(OSC2_ERR_INVALID, "component counts out of range (%d channels, %d solids)", t->n_ch, t->n_sol);
    if (3 * t->n_ch + t->n_sol < 1 || 3 * t->n_ch + t->n_sol > 64)
        return fail(OSC2_ERR_INVALID, "NODOFS = %d outside 1..64", 3 * t->n_ch + t->n_sol);
    const int counts[4] = {t->nff, t->nfs, t->nss, t->nes};
    for (int c : counts)
        if (c < 0 || c > OSC2_MAX_INTERFACES) return fail(OSC2_ERR_INVALID, "interface count out of range");
    for (int j = 0; j < t->n_ch; ++j)
        if (t->ch_intial[j] < 1 || t->ch_intial[j] > 3)
            return fail(OSC2_ERR_INVALID, "abs(INTIAL) of channel %d must be 1, 2 or 3", j);
    for (int k = 0; k < t->nff; ++k)
        if (t->ff[k][0] < 0 || t->ff[k][0] >= t->n_ch || t->ff[k][1] < 0 || t->ff[k][1] >= t->n_ch)
            return fail(OSC2_ERR_INVALID, "fluid-fluid interface %d names a missing channel", k);
    for (int k = 0; k < t->nfs; ++k)
        if (t->fs[k][0] < 0 || t->fs[k][0] >= t->n_ch || t->fs[k][1] < 0 || t->fs[k][1] >= t->n_sol)
            return fail(OSC2_ERR_INVALID, "fluid-solid interface %d names a missing component", k);
    for (int k = 0; k < t->nss; ++k)
        if (t->ss[k][0] < 0 || t->ss[k][0] >= t->n_sol || t->ss[k][1] < 0 || t->ss[k][1] >= t->n_sol)
            return fail(OSC2_ERR_INVALID, "solid-solid interface %d names a missing solid", k);
    for (int k = 0; k < t->nes; ++k)
        if (t->es[k] < 0 || t->es[k] >= t->n_sol)
            return fail(OSC2_ERR_INVALID, "environment interface %d names a missing solid", k);
    if (t->method != 0 && t->method != 1)
        return fail(OSC2_ERR_INVALID, "METHOD must be BE (0) or CN (1); AM4 is not supported");
    return OSC2_OK;
}

// ---- self test of the shared-divisor division (osc2_div.cuh) ---------------
__device__ inline unsigned long long splitmix64(unsigned long long &s)
{
    unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ inline double random_double(unsigned long long &s, int exp_span)
{
    const unsigned long long m = splitmix64(s);
    const unsigned long long mant = m & 0xFFFFFFFFFFFFFull;
    const int e = 1023 + (int)((m >> 52) % (unsigned)(2 * exp_span + 1)) - exp_span;
    const unsigned long long sign = (m >> 63) << 63;
    return __longlong_as_double((long long)(sign | ((unsigned long long)e << 52) | mant));
}

__global__ void osc2_div_selftest_kernel(unsigned long long seed, long long per_thread, unsigned long long *mismatch)
{
    unsigned long long s = seed + 0x1234567ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
    unsigned long long bad = 0;
    for (long long i = 0; i < per_thread; ++i) {
        const int mode = (int)(splitmix64(s) & 15);
        double a, b;
        if (mode < 8) { a = random_double(s, 40); b = random_double(s, 40); }
        else if (mode < 11) { a = random_double(s, 1000); b = random_double(s, 1000); }
        else if (mode == 11) { a = (splitmix64(s) & 1) ? 0.0 : -0.0; b = random_double(s, 300); }
        else if (mode == 12) { a = random_double(s, 20) * 1e-300; b = random_double(s, 20); }
        else if (mode == 13) { a = random_double(s, 20); b = random_double(s, 3) * 1e305; }
        else if (mode == 14) { a = random_double(s, 3); b = 1.0 + (double)(splitmix64(s) & 0xFFFF) * 2.220446049250313e-16; }
        else { a = random_double(s, 600); b = (double)(1 + (splitmix64(s) & 1023)); }
        const Osc2Rcp r = osc2_rcp_prepare(b);
        const double q = osc2_div(a, r);
        const double ref = a / b;
        if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) ++bad;
        // the forward sweep's forms: with the divisor in [2^-100, 2^101) a quotient that is not
        // poisoned must be a / b (a zero may come back with the other sign, see osc2_div.cuh);
        // a numerator at or above 2^900 is left to the non-finite propagation described there
        const bool div_ok = osc2_abs_hi(b) - OSC2_HI_DIV_LO < OSC2_HI_DIV_SPAN;
        const double qp = osc2_div_p(a, r);
        if (div_ok && qp == qp && osc2_abs_hi(a) < OSC2_HI_2P900 &&
            __double_as_longlong(qp) != __double_as_longlong(ref) && !(a == 0.0 && qp == 0.0)) ++bad;
        if (div_ok && osc2_num_key(a) >= OSC2_KEY_MIN && osc2_abs_hi(a) < OSC2_HI_2P900) {
            const double qf = osc2_div_fast(a, r);
            if (__double_as_longlong(qf) != __double_as_longlong(ref) && !(a == 0.0 && qf == 0.0)) ++bad;
        }
    }
    if (bad) atomicAdd(mismatch, bad);
}

}  // namespace

extern "C" {

int osc2_selftest_division(int device, long long n, unsigned long long seed, unsigned long long *mismatches)
{
    if (!mismatches) return fail(OSC2_ERR_INVALID, "NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(OSC2_ERR_NO_DEVICE, "no CUDA device visible"); }
    CK(cudaSetDevice(device));
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **)&d, sizeof(unsigned long long)));
    CK(cudaMemset(d, 0, sizeof(unsigned long long)));
    const int threads = 148 * 8 * 128;
    const long long per = (n + threads - 1) / threads;
    osc2_div_selftest_kernel<<<148 * 8, 128>>>(seed, per, d);
    CK(cudaGetLastError());
    CK(cudaMemcpy(mismatches, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    return OSC2_OK;
}

const char *osc2_last_error(void) { return g_err.c_str(); }
int osc2_version(void) { return 100; }

int osc2_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int osc2_upload_math_tables();

int osc2_create(const osc2_topology *topo, int batch, int n_nodes, int device, osc2_handle *out)
{
    if (!out) return fail(OSC2_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    int rc = validate_topology(topo);
    if (rc) return rc;
    if (batch < 1) return fail(OSC2_ERR_INVALID, "batch must be >= 1");
    if (n_nodes < 4)
        return fail(OSC2_ERR_INVALID, "n_nodes = %d: the reference needs at least 4 nodes "
                    "(build_known_therm_vector indexes out of bounds below that)", n_nodes);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(OSC2_ERR_NO_DEVICE, "no CUDA device visible; opensc2_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(OSC2_ERR_INVALID, "device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    osc2_context *h = new osc2_context();
    struct Guard { osc2_context *h; ~Guard() { if (h) osc2_destroy(h); } } guard{h};   // frees everything on any early return
    h->device = device;
    h->topo = *topo;
    h->B = batch; h->N = n_nodes; h->E = n_nodes - 1;
    h->M = topo->n_ch; h->S = topo->n_sol; h->d = 3 * h->M + h->S;
    h->total = (size_t)h->d * h->N;
    {
        const char *force = getenv("OSC2_FORCE_GENERIC");
        h->fast = osc2_has_fast_path(h->d) && !(force && force[0] == '1');
    }
    h->NS = osc2_gm_slots(h->d, h->M, h->S, h->fast);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    h->sm_count = prop.multiProcessorCount;
    { const int trc = osc2_upload_math_tables(); if (trc) return trc; }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->t0));
    CK(cudaEventCreate(&h->t1));
    CK(cudaMalloc((void **)&h->topo_dev, sizeof(osc2_topology)));
    CK(cudaMemcpy(h->topo_dev, topo, sizeof(osc2_topology), cudaMemcpyHostToDevice));
    CK(cudaMallocHost((void **)&h->status_host, sizeof(int) * (size_t)batch));

    const size_t B = batch, E = h->E, N = h->N, M = h->M, S = h->S, d = h->d;
    StepDev &p = h->p;
    memset(&p, 0, sizeof p);
    p.B = batch; p.N = h->N; p.E = h->E; p.M = h->M; p.S = h->S; p.d = h->d; p.NS = h->NS;
    p.cpw = h->fast ? osc2_fast_cpw(h->d) : 1;
    p.NT = h->fast ? (int)osc2_fast_tiles(h->d, batch) : batch;
    p.nff = topo->nff; p.nfs = topo->nfs; p.nss = topo->nss; p.nes = topo->nes;
    struct { const double **dst; size_t k; } ins[] = {
        {&p.dz, E}, {&p.fl_gauss, 9 * M * E}, {&p.fl_node, 4 * M}, {&p.fl_bc, 6 * M},
        {&p.sol_gauss, 3 * S * E}, {&p.sol_q, 2 * S * E * 2},
        {&p.peri_ff, 2 * (size_t)p.nff * E}, {&p.htc_ff, 2 * (size_t)p.nff * E},
        {&p.peri_fs, (size_t)p.nfs * E}, {&p.htc_fs, (size_t)p.nfs * E},
        {&p.peri_ss, (size_t)p.nss * E}, {&p.htc_ss, (size_t)p.nss * E},
        {&p.peri_es, (size_t)p.nes * E}, {&p.htc_es, (size_t)p.nes * 3 * E}, {&p.qsource, N * S},
        {&p.peri_ss_nodal, (size_t)p.nss * N},
    };
    size_t kmax = h->total * 2;   // syslod
    for (auto &in : ins) {
        double *q = nullptr;
        rc = dev_alloc(h, &q, in.k * B);
        if (rc) return rc;
        *in.dst = q;
        if (in.k > kmax) kmax = in.k;
    }
    const size_t k123 = 3 * (size_t)p.nff * E;
    if (k123 > kmax) kmax = k123;
    rc = dev_alloc(h, &h->sysvar_a, h->total * B);
    if (!rc) rc = dev_alloc(h, &h->sysvar_b, h->total * B);
    if (!rc) rc = dev_alloc(h, &p.syslod, h->total * 2 * B);
    if (!rc) rc = dev_alloc(h, &p.gm, osc2_gm_doubles(h->d, h->M, h->S, h->fast, E, B));
    if (!rc) rc = dev_alloc(h, &p.k123, k123 * B);
    if (!rc) rc = dev_alloc(h, &p.spill, osc2_spill_doubles(h->d, h->fast, N, B));
    if (!rc) rc = dev_alloc(h, &p.norm_solution, d * B);
    if (!rc) rc = dev_alloc(h, &p.norm_change, d * B);
    if (!rc) rc = dev_alloc(h, &p.eqteig, d * B);
    if (!rc) { double *q = nullptr; rc = dev_alloc(h, &q, (B + 1) / 2 + 1); p.status = (int *)q; }
    if (!rc) {   // sparsity pattern of SMAT (K_G stages only the entries the topology can touch)
        std::vector<unsigned char> map;
        int slots = osc2_build_smap(*topo, map);
        double *qm = nullptr;
        // large systems (d > 15) can exceed a byte; they use the generic kernels with dense staging
        if (slots > 0) rc = dev_alloc(h, &qm, map.size() / 8 + 1);
        if (!rc && slots > 0) {
            CK(cudaMemcpy(qm, map.data(), map.size(), cudaMemcpyHostToDevice));
            p.smap = (const unsigned char *)qm;
            p.s_slots = slots;
            if (h->fast) {   // inverse map for K_G's final pass; the constant zeros of `gm` are written once, here
                std::vector<unsigned short> inv((size_t)slots, 0);
                for (size_t i = 0; i < map.size(); ++i)
                    if (map[i]) inv[map[i]] = (unsigned short)((i % h->d) * 32 + (i / h->d) * p.cpw);
                double *qi = nullptr;
                rc = dev_alloc(h, &qi, inv.size() / 4 + 1);
                if (!rc) {
                    CK(cudaMemcpy(qi, inv.data(), inv.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
                    p.sinv = (const unsigned short *)qi;
                    CK(cudaMemset(p.gm, 0, osc2_gm_doubles(h->d, h->M, h->S, true, E, B) * sizeof(double)));
                    p.gm_prezeroed = 1;
                }
            }
        }
    }
    if (!rc) {   // numpy pairwise-sum tree of the N nodal values of one equation (norms)
        std::vector<int> leaves, prog;
        osc2_build_norm_plan(0, h->N, leaves, prog);
        p.n_leaves = (int)(leaves.size() / 2);
        p.n_prog = (int)prog.size();
        double *ql = nullptr, *qp = nullptr;
        rc = dev_alloc(h, &ql, leaves.size() / 2 + 1);
        if (!rc) rc = dev_alloc(h, &qp, prog.size() / 2 + 1);
        if (!rc) rc = dev_alloc(h, &p.norm_part, (size_t)p.n_leaves * 3 * d * B);
        if (!rc) {
            CK(cudaMemcpy(ql, leaves.data(), leaves.size() * sizeof(int), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(qp, prog.data(), prog.size() * sizeof(int), cudaMemcpyHostToDevice));
            p.norm_leaf = (const int *)ql;
            p.norm_prog = (const int *)qp;
        }
    }
    h->staging_doubles = kmax * B;
    if (!rc) rc = dev_alloc(h, &h->staging, h->staging_doubles);
    if (rc) return rc;
    p.sysvar = h->sysvar_a;
    p.sysvar_new = h->sysvar_b;
    cudaMemsetAsync(p.status, 0, sizeof(int) * B, h->stream);
    cudaMemsetAsync(p.syslod, 0, h->total * 2 * B * sizeof(double), h->stream);
    CK(cudaStreamSynchronize(h->stream));
    guard.h = nullptr;
    *out = h;
    return OSC2_OK;
}

// tables of the device exp / log (props_kernels.cuh), rounded from long double; written to the current device
static int osc2_upload_math_tables()
{
    double lt[128][2], et[64];
    for (int i = 0; i < 128; ++i) {
        const long double c = 1.0L + (i + 0.5L) / 128.0L;
        const double rc = (double)(1.0L / c);
        lt[i][0] = rc;
        lt[i][1] = (double)(-logl((long double)rc));      // log of the value m is really divided by
    }
    for (int k = 0; k < 64; ++k) et[k] = (double)powl(2.0L, k / 64.0L);
    CK(cudaMemcpyToSymbol(osc2p::osc2_log_table, lt, sizeof lt));
    CK(cudaMemcpyToSymbol(osc2p::osc2_exp_table, et, sizeof et));
    return OSC2_OK;
}

int osc2_destroy(osc2_handle h)
{
    if (!h) return OSC2_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *q : h->allocs) cudaFree(q);
    if (h->topo_dev) cudaFree(h->topo_dev);
    if (h->model_dev) cudaFree(h->model_dev);
    if (h->status_host) cudaFreeHost(h->status_host);
    for (auto &ev : h->pending) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto &ev : h->pool) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    if (h->t0) cudaEventDestroy(h->t0);
    if (h->t1) cudaEventDestroy(h->t1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return OSC2_OK;
}

int osc2_upload_state(osc2_handle h, const double *sysvar, const double *syslod)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc = OSC2_OK;
    if (sysvar) {
        rc = upload_transposed(h, sysvar, (double *)h->p.sysvar, h->total);
        if (rc) return rc;
        h->have_state = true;
        CK(cudaMemsetAsync(h->p.status, 0, sizeof(int) * (size_t)h->B, h->stream));   // a new state: the sticky status starts over
    }
    if (syslod) { rc = upload_transposed(h, syslod, h->p.syslod, h->total * 2); if (rc) return rc; }
    return OSC2_OK;
}

int osc2_upload_step_inputs(osc2_handle h, const osc2_step_inputs *in)
{
    if (!h || !in) return fail(OSC2_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(h->device));
    const size_t E = h->E, N = h->N, M = h->M, S = h->S;
    StepDev &p = h->p;
    struct { const double *src; const double *dst; size_t k; const char *name; } a[] = {
        {in->dz, p.dz, E, "dz"}, {in->fl_gauss, p.fl_gauss, 9 * M * E, "fl_gauss"},
        {in->fl_node, p.fl_node, 4 * M, "fl_node"}, {in->fl_bc, p.fl_bc, 6 * M, "fl_bc"},
        {in->sol_gauss, p.sol_gauss, 3 * S * E, "sol_gauss"}, {in->sol_q, p.sol_q, 2 * S * E * 2, "sol_q"},
        {in->peri_ff, p.peri_ff, 2 * (size_t)p.nff * E, "peri_ff"}, {in->htc_ff, p.htc_ff, 2 * (size_t)p.nff * E, "htc_ff"},
        {in->peri_fs, p.peri_fs, (size_t)p.nfs * E, "peri_fs"}, {in->htc_fs, p.htc_fs, (size_t)p.nfs * E, "htc_fs"},
        {in->peri_ss, p.peri_ss, (size_t)p.nss * E, "peri_ss"}, {in->htc_ss, p.htc_ss, (size_t)p.nss * E, "htc_ss"},
        {in->peri_es, p.peri_es, (size_t)p.nes * E, "peri_es"}, {in->htc_es, p.htc_es, (size_t)p.nes * 3 * E, "htc_es"},
        {in->qsource, p.qsource, N * S, "qsource"},
    };
    if (in->peri_ss_nodal && p.nss > 0) {
        int rc = upload_transposed(h, in->peri_ss_nodal, (double *)p.peri_ss_nodal, (size_t)p.nss * N);
        if (rc) return rc;
        h->have_peri_ss_nodal = true;
    }
    // what the device never computes: a first upload must bring all of it (later calls may refresh single arrays,
    // e.g. the boundary values every step)
    const void *geometry[] = {in->dz, in->fl_bc, in->peri_ff, in->peri_fs, in->peri_ss, in->peri_es, in->qsource};
    const size_t geometry_k[] = {E, 6 * M, 2 * (size_t)p.nff * E, (size_t)p.nfs * E, (size_t)p.nss * E, (size_t)p.nes * E, N * S};
    bool all_geometry = true;
    for (int i = 0; i < 7; ++i) if (geometry_k[i] && !geometry[i]) all_geometry = false;
    bool all_present = true;
    for (auto &x : a) {
        if (x.k == 0) continue;
        if (!x.src) { all_present = false; continue; }
        int rc = upload_transposed(h, x.src, (double *)x.dst, x.k);
        if (rc) return rc;
    }
    if (!all_present && !h->have_model)
        return fail(OSC2_ERR_INVALID, "a step input is NULL and no model is set to compute it (osc2_set_model)");
    p.t_env = in->t_env;
    if (all_geometry) h->have_geometry = true;
    if (all_present) h->have_inputs = true;
    return OSC2_OK;
}

int osc2_set_model(osc2_handle h, const osc2_model *model)
{
    if (!h || !model) return fail(OSC2_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(h->device));
    int nf = 0;
    for (int j = 0; j < h->M; ++j) {
        if (model->ch_fluid[j] < 0 || model->ch_fluid[j] >= OSC2_MAX_FLUIDS)
            return fail(OSC2_ERR_INVALID, "channel %d: fluid table index %d", j, model->ch_fluid[j]);
        if (model->ch_fluid[j] + 1 > nf) nf = model->ch_fluid[j] + 1;
        {
            const int ifr = model->ch_ifriction[j];
            static const int known[] = {-99, 100, 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111, 112, 113, 114, 115, 118, 119, 120, 121, 122, 124,
                                        200, 201, 202, 203, 204, 205, 206, 207, 209, 210, 211};
            bool ok = false;
            for (int v : known) ok = ok || (ifr == v);
            if (!ok)
                return fail(OSC2_ERR_INVALID, "channel %d: IFRICTION %d is not implemented on the device (-99, 100-115, 118-122, 124, 200-207, 209-211; "
                            "116 / 117 / 123 fail and 208 raises in the reference itself)", j, ifr);
        }
        {
            const int hc = model->ch_htc_corr[j];
            if (hc != 1 && hc != 212 && hc != 119 && hc != 120 && hc != 121 && hc != 211)
                return fail(OSC2_ERR_INVALID, "channel %d: Flag_htc_steady_corr %d is not implemented on the device (1, 119, 120, 121, 211, 212)", j, hc);
        }
    }
    for (int l = 0; l < h->S; ++l) {
        if (model->sol_nmat[l] < 1 || model->sol_nmat[l] > OSC2_MAX_MATERIALS)
            return fail(OSC2_ERR_INVALID, "solid %d: %d materials", l, model->sol_nmat[l]);
        if (model->sol_kind[l] < OSC2_SOLID_STRAND_MIXED || model->sol_kind[l] > OSC2_SOLID_STACK)
            return fail(OSC2_ERR_INVALID, "solid %d: unknown kind %d", l, model->sol_kind[l]);
        for (int i = 0; i < model->sol_nmat[l]; ++i)
            if (model->sol_mat[l][i] < 0 || model->sol_mat[l][i] >= OSC2_MAT_COUNT)
                return fail(OSC2_ERR_INVALID, "solid %d: material id %d is not implemented on the device", l, model->sol_mat[l][i]);
    }
    for (int k = 0; k < h->p.nfs; ++k)
        if (model->fs_choice[k] != 2 && model->fs_choice[k] != -2)
            return fail(OSC2_ERR_INVALID, "channel-solid interface %d: HTC_choice %d is not implemented on the device (2, -2)", k, model->fs_choice[k]);
    for (int k = 0; k < h->p.nff; ++k)
        if (model->ff_choice[k] != 2 && model->ff_choice[k] != -2)
            return fail(OSC2_ERR_INVALID, "channel-channel interface %d: HTC_choice %d is not implemented on the device (2, -2)", k, model->ff_choice[k]);
    for (int k = 0; k < h->p.nss; ++k)
        if (model->ss_choice[k] != 1 && model->ss_choice[k] != -1 && model->ss_choice[k] != 3 && model->ss_choice[k] != -3)
            return fail(OSC2_ERR_INVALID, "solid-solid interface %d: HTC_choice %d is not implemented on the device (1, -1, 3, -3)", k, model->ss_choice[k]);
    if (!h->model_dev) CK(cudaMalloc((void **)&h->model_dev, sizeof(osc2_model)));
    CK(cudaStreamSynchronize(h->stream));    // kernels in flight on the (non-blocking) stream still read the old model
    CK(cudaMemcpy(h->model_dev, model, sizeof(osc2_model), cudaMemcpyHostToDevice));
    h->need_peri_ss_nodal = false;
    for (int k = 0; k < h->p.nss; ++k)
        if (model->ss_choice[k] == 3 || model->ss_choice[k] == -3) h->need_peri_ss_nodal = true;
    if (!h->have_model) {
        memset(&h->q, 0, sizeof h->q);
        double *t = nullptr;
        int rc = dev_alloc(h, &t, (size_t)h->S * h->B);
        if (rc) return rc;
        h->q.tmax = t;
        t = nullptr;
        rc = dev_alloc(h, &t, (size_t)h->S * h->B);
        if (rc) return rc;
        h->q.tmin = t;
        t = nullptr;
        rc = dev_alloc(h, &t, ((size_t)(h->M > 0 ? h->M : 1) * h->B + 1) / 2 + 1);   // ints: max(M, 1) * B of them
        if (rc) return rc;
        h->q.cb_iters = (int *)t;
        double *c = nullptr;
        rc = dev_alloc(h, &c, 2 * (size_t)h->S);
        if (rc) return rc;
        h->q.sol_const = c;
        t = nullptr;
        rc = dev_alloc(h, &t, 2 * (size_t)(h->M > 0 ? h->M : 1) * h->E * h->B);
        if (rc) return rc;
        h->q.ch_scratch = t;
        t = nullptr;
        rc = dev_alloc(h, &t, (size_t)h->S * h->B);
        if (rc) return rc;
        h->q.margin_key = (unsigned long long *)t;
        rc = dev_alloc(h, &h->margin_out, (size_t)h->S * h->B);
        if (rc) return rc;
        CK(cudaMemsetAsync(t, 0xff, sizeof(double) * (size_t)h->S * h->B, h->stream));
    }
    if (h->S > 0) {
        osc2_model_consts_kernel<<<(h->S + 63) / 64, 64, 0, h->stream>>>(h->model_dev, h->S, h->q.sol_const);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->stream));
    }
    h->q.model = h->model_dev;
    h->n_fluids_used = nf;
    h->have_model = true;
    return OSC2_OK;
}

int osc2_set_fluid_table(osc2_handle h, int fluid, int n_p, int n_t, double p_min, double p_step,
                         double t_min, double t_step, const double *props)
{
    if (!h || !props) return fail(OSC2_ERR_INVALID, "NULL argument");
    if (!h->have_model) return fail(OSC2_ERR_STATE, "osc2_set_model first");
    if (fluid < 0 || fluid >= OSC2_MAX_FLUIDS) return fail(OSC2_ERR_INVALID, "fluid index %d", fluid);
    if (n_p < 2 || n_t < 2 || !(p_step > 0.0) || !(t_step > 0.0)) return fail(OSC2_ERR_INVALID, "table needs >= 2 x 2 points and positive steps");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)OSC2_N_FLUID_PROPS * n_p * n_t;
    FluidTableDev &old = h->q.fluid[fluid];
    double *dev = (h->have_table[fluid] && old.n_p == n_p && old.n_t == n_t) ? const_cast<double *>(old.props) : nullptr;
    if (!dev) {
        int rc = dev_alloc(h, &dev, n);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));    // do not overwrite a table kernels in flight are reading
    CK(cudaMemcpy(dev, props, n * sizeof(double), cudaMemcpyHostToDevice));
    FluidTableDev &t = h->q.fluid[fluid];
    t.n_p = n_p; t.n_t = n_t; t.p_min = p_min; t.p_step = p_step; t.t_min = t_min; t.t_step = t_step; t.props = dev;
    h->have_table[fluid] = true;
    return OSC2_OK;
}

// host [B][K] int-free doubles -> device [K][B]; allocates on first use
static int upload_sweep_array(osc2_context *h, const double *host, const double **dev, size_t K)
{
    if (K == 0) return OSC2_OK;
    if (!*dev) { double *t = nullptr; int rc = dev_alloc(h, &t, K * h->B); if (rc) return rc; *dev = t; }
    if (K * h->B > h->staging_doubles) return fail(OSC2_ERR_STATE, "staging buffer too small");
    return upload_transposed(h, host, (double *)*dev, K);
}

int osc2_upload_sweep(osc2_handle h, const osc2_sweep *sw)
{
    if (!h || !sw) return fail(OSC2_ERR_INVALID, "NULL argument");
    if (!h->have_model) return fail(OSC2_ERR_STATE, "osc2_set_model first");
    if (h->S > 0 && (!sw->b_field || !sw->eps || !sw->q0 || !sw->q_window))
        return fail(OSC2_ERR_INVALID, "b_field, eps, q0 and q_window are required");
    CK(cudaSetDevice(h->device));
    const size_t S = h->S, E = h->E;
    int rc = upload_sweep_array(h, sw->b_field, &h->q.b_field, S * 2);
    if (!rc) rc = upload_sweep_array(h, sw->eps, &h->q.eps, S);
    if (!rc) rc = upload_sweep_array(h, sw->q0, &h->q.q0, S);
    if (!rc) rc = upload_sweep_array(h, sw->q_window, &h->q.q_window, S * 4);
    if (!rc && sw->tcs) rc = upload_sweep_array(h, sw->tcs, &h->q.tcs, S * E);
    if (!rc && sw->jop) rc = upload_sweep_array(h, sw->jop, &h->q.jop, S);
    if (rc) return rc;
    h->have_sweep = true;
    return OSC2_OK;
}

static int launch_tmax(osc2_context *h);

int osc2_eval_operating_conditions(osc2_handle h, double time)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    if (!h->have_model || !h->have_sweep || !h->have_state)
        return fail(OSC2_ERR_STATE, "osc2_eval_operating_conditions needs osc2_set_model, osc2_upload_sweep and osc2_upload_state first");
    for (int f = 0; f < h->n_fluids_used; ++f)
        if (!h->have_table[f]) return fail(OSC2_ERR_STATE, "fluid table %d was not set", f);
    if (!h->have_geometry)
        return fail(OSC2_ERR_STATE, "dz, fl_bc, the contact perimeters and qsource are never computed on the device: "
                    "upload them once with osc2_upload_step_inputs before osc2_eval_operating_conditions / osc2_advance");
    if (h->need_peri_ss_nodal && !h->have_peri_ss_nodal)
        return fail(OSC2_ERR_STATE, "a radiative solid-solid interface (HTC_choice +-3) needs peri_ss_nodal");
    CK(cudaSetDevice(h->device));
    h->q.time = time;
    { int rc = launch_tmax(h); if (rc) return rc; }
    CudaLauncher L{h};
    const long long work = (long long)h->E * h->B;
    if (h->M > 0)
        L.launch(OSC2_K_PROPS, osc2_opcond_coolant_kernel, (unsigned)((work * h->M + 127) / 128), 128u, (size_t)0,
                 (const osc2_topology *)h->topo_dev, h->p, h->q);
    if (h->S > 0)
        L.launch(OSC2_K_PROPS, osc2_opcond_solid_kernel, (unsigned)((work * h->S + 127) / 128), 128u, (size_t)0, h->p, h->q);
    L.launch(OSC2_K_PROPS, osc2_opcond_interface_kernel, (unsigned)((work + 127) / 128), 128u, (size_t)0,
             (const osc2_topology *)h->topo_dev, h->p, h->q);
    if (!h->launch_error.empty()) {
        const int rc = fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
        h->launch_error.clear();
        return rc;
    }
    h->have_inputs = true;
    return OSC2_OK;
}

int osc2_download_step_inputs(osc2_handle h, double *fl_gauss, double *fl_node, double *sol_gauss, double *sol_q,
                              double *htc_ff, double *htc_fs, double *htc_ss, double *htc_es)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    const size_t E = h->E, M = h->M, S = h->S;
    const StepDev &p = h->p;
    struct { double *dst; const double *src; size_t k; } a[] = {
        {fl_gauss, p.fl_gauss, 9 * M * E}, {fl_node, p.fl_node, 4 * M}, {sol_gauss, p.sol_gauss, 3 * S * E},
        {sol_q, p.sol_q, 2 * S * E * 2}, {htc_ff, p.htc_ff, 2 * (size_t)p.nff * E}, {htc_fs, p.htc_fs, (size_t)p.nfs * E},
        {htc_ss, p.htc_ss, (size_t)p.nss * E}, {htc_es, p.htc_es, (size_t)p.nes * 3 * E},
    };
    for (auto &x : a) {
        if (!x.dst || x.k == 0) continue;
        int rc = download_transposed(h, x.src, x.dst, x.k);
        if (rc) return rc;
    }
    return OSC2_OK;
}

static int run_step(osc2_handle h, double dt, int num_step, bool records_ready);

int osc2_step(osc2_handle h, double dt, int num_step)
{
    return run_step(h, dt, num_step, false);
}

static int launch_tmax(osc2_context *h)
{
    if (h->S == 0 && h->M == 0) return OSC2_OK;
    CudaLauncher L{h};
    CK(cudaMemsetAsync(h->q.tmax, 0, sizeof(double) * (size_t)h->S * h->B, h->stream));
    CK(cudaMemsetAsync(h->q.tmin, 0x7f, sizeof(double) * (size_t)h->S * h->B, h->stream));   // a huge positive double
    if (h->M > 0) CK(cudaMemsetAsync(h->q.cb_iters, 0, sizeof(int) * (size_t)h->M * h->B, h->stream));
    if (h->q.margin_key) CK(cudaMemsetAsync(h->q.margin_key, 0xff, sizeof(unsigned long long) * (size_t)h->S * h->B, h->stream));
    const long long w = (long long)((h->E + 63) / 64) * h->B;
    L.launch(OSC2_K_PROPS, osc2_solid_tmax_kernel, (unsigned)((w + 127) / 128), 128u, (size_t)0,
             (const osc2_topology *)h->topo_dev, h->p, h->q);
    return OSC2_OK;
}

int osc2_advance(osc2_handle h, double time, double dt, int num_step)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    int rc = osc2_eval_operating_conditions(h, time);
    if (rc) return rc;
    return run_step(h, dt, num_step, false);
}

int osc2_eval_tcs(osc2_handle h)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    if (!h->have_model || !h->have_sweep) return fail(OSC2_ERR_STATE, "osc2_eval_tcs needs osc2_set_model and osc2_upload_sweep first");
    if (!h->q.jop) return fail(OSC2_ERR_STATE, "osc2_eval_tcs needs the operating current density (osc2_sweep.jop)");
    CK(cudaSetDevice(h->device));
    const size_t S = h->S, E = h->E, N = h->N, B = h->B;
    if (!h->q.tcs) { double *t = nullptr; int rc = dev_alloc(h, &t, S * E * B); if (rc) return rc; h->q.tcs = t; }
    if (!h->tcs_nodal) { int rc = dev_alloc(h, &h->tcs_nodal, S * N * B); if (rc) return rc; }
    CudaLauncher L{h};
    const long long work = (long long)(N + E) * B;
    L.launch(OSC2_K_PROPS, osc2_tcs_kernel, (unsigned)((work + 127) / 128), 128u, (size_t)0, h->p, h->q,
             const_cast<double *>(h->q.tcs), h->tcs_nodal);
    if (!h->launch_error.empty()) {
        const int rc = fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
        h->launch_error.clear();
        return rc;
    }
    return OSC2_OK;
}

int osc2_eval_nodal_properties(osc2_handle h)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    if (!h->have_model || !h->have_state) return fail(OSC2_ERR_STATE, "osc2_eval_nodal_properties needs osc2_set_model and a state first");
    for (int f = 0; f < h->n_fluids_used; ++f)
        if (!h->have_table[f]) return fail(OSC2_ERR_STATE, "fluid table %d was not set", f);
    if (h->M == 0) return OSC2_OK;
    CK(cudaSetDevice(h->device));
    const size_t M = h->M, N = h->N, B = h->B;
    if (!h->nodal) { int rc = dev_alloc(h, &h->nodal, (size_t)OSC2_N_NODAL * M * N * B); if (rc) return rc; }
    if (!h->cb_nodal) { double *t = nullptr; int rc = dev_alloc(h, &t, (M * B + 1) / 2 + 1); if (rc) return rc; h->cb_nodal = (int *)t; }
    CK(cudaMemsetAsync(h->cb_nodal, 0, sizeof(int) * M * B, h->stream));
    CudaLauncher L{h};
    const long long work = (long long)(M * N * B);
    for (int pass = 0; pass < 2; ++pass)
        L.launch(OSC2_K_PROPS, osc2_nodal_coolant_kernel, (unsigned)((work + 127) / 128), 128u, (size_t)0,
                 (const osc2_topology *)h->topo_dev, h->p, h->q, h->nodal, h->cb_nodal, pass);
    if (!h->launch_error.empty()) {
        const int rc = fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
        h->launch_error.clear();
        return rc;
    }
    return OSC2_OK;
}

int osc2_download_nodal_properties(osc2_handle h, double *chan)
{
    if (!h || !chan) return fail(OSC2_ERR_INVALID, "NULL argument");
    if (!h->nodal) return fail(OSC2_ERR_STATE, "osc2_eval_nodal_properties first");
    CK(cudaSetDevice(h->device));
    const size_t K = (size_t)h->M * h->N;
    for (int q = 0; q < OSC2_N_NODAL; ++q) {
        int rc = download_transposed(h, h->nodal + (size_t)q * K * h->B, chan + (size_t)q * K * h->B, K);
        if (rc) return rc;
    }
    return OSC2_OK;
}

int osc2_download_tcs(osc2_handle h, double *tcs_gauss, double *tcs_nodal)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc;
    if (tcs_gauss) {
        if (!h->q.tcs) return fail(OSC2_ERR_STATE, "no current sharing temperature on the device (osc2_upload_sweep with tcs, or osc2_eval_tcs)");
        rc = download_transposed(h, h->q.tcs, tcs_gauss, (size_t)h->S * h->E);
        if (rc) return rc;
    }
    if (tcs_nodal) {
        if (!h->tcs_nodal) return fail(OSC2_ERR_STATE, "nodal current sharing temperature needs osc2_eval_tcs");
        rc = download_transposed(h, h->tcs_nodal, tcs_nodal, (size_t)h->S * h->N);
        if (rc) return rc;
    }
    return OSC2_OK;
}

__global__ void osc2_margin_decode_kernel(const unsigned long long *key, double *out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    out[i] = (k == ~0ull) ? __longlong_as_double(0x7ff0000000000000ll) : __longlong_as_double((long long)u);
}

int osc2_download_margin(osc2_handle h, double *margin)
{
    if (!h || !margin) return fail(OSC2_ERR_INVALID, "NULL argument");
    if (!h->q.margin_key) return fail(OSC2_ERR_STATE, "osc2_download_margin needs osc2_set_model first");
    CK(cudaSetDevice(h->device));
    const long long n = (long long)h->S * h->B;
    if ((size_t)n > h->staging_doubles) return fail(OSC2_ERR_STATE, "staging buffer too small");
    osc2_margin_decode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->q.margin_key, h->margin_out, n);
    CK(cudaGetLastError());
    return download_transposed(h, h->margin_out, margin, (size_t)h->S);
}

static int run_step(osc2_handle h, double dt, int num_step, bool records_ready)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    if (!h->have_state || !h->have_inputs)
        return fail(OSC2_ERR_STATE, "osc2_step needs osc2_upload_state and osc2_upload_step_inputs first");
    if (!(dt > 0.0)) return fail(OSC2_ERR_INVALID, "time step must be positive");
    if (num_step < 1) return fail(OSC2_ERR_INVALID, "num_step must be >= 1");
    CK(cudaSetDevice(h->device));
    h->p.dt = dt;
    h->p.num_step = num_step;
    if (h->capture) {
        const size_t full = 4 * (size_t)h->d - 1;
        CK(cudaMemsetAsync(h->p.cap_sysmat, 0, full * h->total * h->B * sizeof(double), h->stream));
    }
    CudaLauncher L{h};
    StepDev pp = h->p;
    if (!h->capture) { pp.cap_sysmat = nullptr; pp.cap_known = nullptr; }
    osc2_run_step(L, h->topo_dev, pp, h->fast, h->topo.theta, records_ready);
    if (!h->launch_error.empty()) {
        const int rc = fail(OSC2_ERR_CUDA, "%s", h->launch_error.c_str());
        h->launch_error.clear();
        return rc;
    }
    // the new solution becomes the state of the next step
    const double *old = h->p.sysvar;
    h->p.sysvar = h->p.sysvar_new;
    h->p.sysvar_new = (double *)old;
    return OSC2_OK;
}

int osc2_synchronize(osc2_handle h)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->status_host, h->p.status, sizeof(int) * (size_t)h->B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < h->B; ++b)
        if (h->status_host[b] == OSC2_STATUS_TCS)
            return fail(OSC2_ERR_TCS, "conductor %d of the batch: f(a) and f(b) must have different signs -- the operating "
                        "current density is not below Jc(3.5 K) (current_sharing_temperature_nb3sn)", b);
    for (int b = 0; b < h->B; ++b)
        if (h->status_host[b] == OSC2_STATUS_RANGE)
            return fail(OSC2_ERR_RANGE, "conductor %d of the batch: a value left the range in which the fast kernels' "
                        "division is verified exact (0 < |x| < 2^-900 or a divisor outside 2^+-100); "
                        "set OSC2_FORCE_GENERIC=1 to use the generic kernels", b);
    for (int b = 0; b < h->B; ++b)
        if (h->status_host[b] != 0)
            return fail(OSC2_ERR_SINGULAR, "ERROR! Matrix is singular at line %d (conductor %d of the batch)",
                        -h->status_host[b] - 1, b);
    return OSC2_OK;
}

int osc2_download_state(osc2_handle h, double *sysvar, double *syslod, int *status)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc;
    if (sysvar) { rc = download_transposed(h, h->p.sysvar, sysvar, h->total); if (rc) return rc; }
    if (syslod) { rc = download_transposed(h, h->p.syslod, syslod, h->total * 2); if (rc) return rc; }
    if (status) {
        CK(cudaMemcpyAsync(status, h->p.status, sizeof(int) * (size_t)h->B, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return OSC2_OK;
}

int osc2_download_diagnostics(osc2_handle h, double *norm_solution, double *norm_change, double *eqteig, double *k123)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc;
    if (norm_solution) { rc = download_transposed(h, h->p.norm_solution, norm_solution, h->d); if (rc) return rc; }
    if (norm_change) { rc = download_transposed(h, h->p.norm_change, norm_change, h->d); if (rc) return rc; }
    if (eqteig) { rc = download_transposed(h, h->p.eqteig, eqteig, h->d); if (rc) return rc; }
    if (k123) { rc = download_transposed(h, h->p.k123, k123, 3 * (size_t)h->p.nff * h->E); if (rc) return rc; }
    return OSC2_OK;
}

int osc2_set_capture(osc2_handle h, int on)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    if (on && !h->p.cap_sysmat) {
        const size_t full = 4 * (size_t)h->d - 1;
        int rc = dev_alloc(h, &h->p.cap_sysmat, full * h->total * h->B);
        if (!rc) rc = dev_alloc(h, &h->p.cap_known, h->total * h->B);
        if (rc) return rc;
        if (full * h->total * h->B > h->staging_doubles) {
            double *q = nullptr;
            rc = dev_alloc(h, &q, full * h->total * h->B);
            if (rc) return rc;
            h->staging = q;
            h->staging_doubles = full * h->total * h->B;
        }
    }
    h->capture = on != 0;   // buffers are kept; osc2_step hides the pointers when off
    return OSC2_OK;
}

int osc2_download_system(osc2_handle h, double *sysmat, double *known)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    if (!h->capture || !h->p.cap_sysmat) return fail(OSC2_ERR_STATE, "enable osc2_set_capture before the step");
    CK(cudaSetDevice(h->device));
    const size_t full = 4 * (size_t)h->d - 1;
    int rc;
    if (sysmat) { rc = download_transposed(h, h->p.cap_sysmat, sysmat, full * h->total); if (rc) return rc; }
    if (known) { rc = download_transposed(h, h->p.cap_known, known, h->total); if (rc) return rc; }
    return OSC2_OK;
}

int osc2_set_profiling(osc2_handle h, int on)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    h->profiling = on != 0;
    return OSC2_OK;
}

int osc2_get_kernel_ms(osc2_handle h, double ms[OSC2_K_COUNT], long long launches[OSC2_K_COUNT])
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc = drain_events(h);
    if (rc) return rc;
    for (int k = 0; k < OSC2_K_COUNT; ++k) {
        if (ms) ms[k] = h->kernel_ms[k];
        if (launches) launches[k] = h->kernel_launches[k];
    }
    return OSC2_OK;
}

int osc2_reset_kernel_ms(osc2_handle h)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    int rc = drain_events(h);
    if (rc) return rc;
    for (int k = 0; k < OSC2_K_COUNT; ++k) { h->kernel_ms[k] = 0.0; h->kernel_launches[k] = 0; }
    return OSC2_OK;
}

int osc2_timer_start(osc2_handle h)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->t0, h->stream));
    return OSC2_OK;
}

int osc2_timer_stop(osc2_handle h, double *elapsed_ms)
{
    if (!h) return fail(OSC2_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->t1, h->stream));
    CK(cudaEventSynchronize(h->t1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->t0, h->t1));
    if (elapsed_ms) *elapsed_ms = ms;
    return OSC2_OK;
}

size_t osc2_device_bytes(osc2_handle h) { return h ? h->bytes : 0; }

}  // extern "C"
