// tsfx_api.cu -- the C ABI of libtsfx.so (include/tsfx.h): context, plan, extraction entry points.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "tsfx_common.cuh"
#include "tsfx_kernels.h"
#include "tsfx_csr.h"
#include "tsfx_impute.h"

using namespace tsfx;

static std::string g_create_error;

namespace tsfx {
int grid_waves(int dflt) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TSFX_GRID_WAVES"); v = e ? atoi(e) : 0; }
    return v > 0 ? v : dflt;
}
int global_above() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TSFX_GLOBAL_ABOVE"); v = e ? atoi(e) : 0; if (v > 227 * 1024 || v < 0) v = 0; }
    return v;
}
int global_ctas_env() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TSFX_GLOBAL_CTAS"); v = e ? atoi(e) : 0; if (v < 1 || v > 16) v = 0; }
    return v;
}
}  // namespace tsfx

static int env_streams() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("TSFX_STREAMS"); v = e ? atoi(e) : 1; if (v < 1) v = 1; if (v > 4) v = 4; }
    return v;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

static const char* kGroupNames[G_EVENTS] = {"basic", "sorted", "spectral", "la", "entropy", "seq", "peaks", "assemble"};

struct tsfx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    DevBuf values, begin, len, out, misc, stage;
    double* d_dec = nullptr;
    double2* d_tw = nullptr;
    int tw_n = 0;
    cudaEvent_t ev[G_EVENTS][2];
    bool ev_used[G_EVENTS];
    float ms[G_EVENTS];
    int launches = 0;
    CsrWorkspace csr;
    ImputeWorkspace imp;
    int64_t held_series = -1;    // CSR kept on the device by the last stage-(a) call (-1: none)
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined host path
    cudaStream_t s_side[3] = {nullptr, nullptr, nullptr};   // optional side streams so kernel groups can overlap
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
};

struct tsfx_plan {
    tsfx_ctx* ctx = nullptr;
    std::vector<Desc> host[G_COUNT];
    Desc* dev[G_COUNT] = {nullptr};
    int32_t* d_final_col = nullptr;   // final column of every staged column, groups concatenated
    int basic_nfin = 0;               // leading "finisher" descriptors of the BASIC group
    int sorted_nfin = 0;              // same for the SORTED group
    int spectral_nfft = 0;            // leading fft_coefficient descriptors of the SPECTRAL group
    int cum[G_COUNT + 1] = {0};
    int ncols = 0;
    int lag_needed = 0, pacf_want = -1;
    int basic_bins = 0, fourier_bins = 0;
    int need_fft = 0, need_welch = 0;
    int max_ar_k = 0, need_adf = 0;
    int max_lz_bins = 0, max_perm_dim = 0, max_cwt_peaks_n = 0, n_lz = 0;
    int friedrich_r = 0;
    double* d_tables = nullptr;
    int64_t* d_toff = nullptr;
    int32_t* d_thalf = nullptr;
    int n_tables = 0;
    std::vector<int64_t> toff;
    std::vector<int32_t> thalf;
};

static int fail(tsfx_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(ctx, TSFX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
    } while (0)

static int group_of(int calc) {
    switch (calc) {
        case TSFX_SYMMETRY_LOOKING: case TSFX_HAS_DUPLICATE: case TSFX_MEDIAN:
        case TSFX_PERCENTAGE_OF_REOCCURRING_VALUES_TO_ALL_VALUES:
        case TSFX_PERCENTAGE_OF_REOCCURRING_DATAPOINTS_TO_ALL_DATAPOINTS:
        case TSFX_SUM_OF_REOCCURRING_VALUES: case TSFX_SUM_OF_REOCCURRING_DATA_POINTS:
        case TSFX_RATIO_VALUE_NUMBER_TO_TIME_SERIES_LENGTH: case TSFX_QUANTILE:
        case TSFX_MEAN_N_ABSOLUTE_MAX: case TSFX_CHANGE_QUANTILES: case TSFX_FRIEDRICH_COEFFICIENTS:
        case TSFX_MAX_LANGEVIN_FIXED_POINT:
            return G_SORTED;
        case TSFX_FFT_COEFFICIENT: case TSFX_FFT_AGGREGATED: case TSFX_SPKT_WELCH_DENSITY:
        case TSFX_FOURIER_ENTROPY: case TSFX_CWT_COEFFICIENTS:
            return G_SPECTRAL;
        case TSFX_AR_COEFFICIENT: case TSFX_AUGMENTED_DICKEY_FULLER:
            return G_LA;
        case TSFX_SAMPLE_ENTROPY: case TSFX_APPROXIMATE_ENTROPY:
            return G_ENTROPY;
        case TSFX_LEMPEL_ZIV_COMPLEXITY: case TSFX_PERMUTATION_ENTROPY:
            return G_SEQ;
        case TSFX_NUMBER_CWT_PEAKS:
            return G_PEAKS;
        default:
            return G_BASIC;
    }
}

extern "C" int tsfx_version(void) { return TSFX_VERSION; }

extern "C" const char* tsfx_last_error(const tsfx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int tsfx_ctx_create(int device, void* cuda_stream, tsfx_ctx** out) {
    if (!out) return fail(nullptr, TSFX_E_INVALID, "out is NULL");
    *out = nullptr;
    tsfx_ctx* ctx = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TSFX_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, TSFX_E_INVALID, "device index out of range");
    ctx = new (std::nothrow) tsfx_ctx();
    if (!ctx) return fail(nullptr, TSFX_E_NOMEM, "out of host memory");
    ctx->device = device;
    for (int g = 0; g < G_EVENTS; ++g) { ctx->ev_used[g] = false; ctx->ms[g] = 0.f; ctx->ev[g][0] = ctx->ev[g][1] = nullptr; }
#define CKC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            std::string m = std::string(#call) + ": " + cudaGetErrorString(e__);                     \
            delete ctx;                                                                               \
            return fail(nullptr, TSFX_E_CUDA, m);                                                     \
        }                                                                                             \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    for (int g = 0; g < G_EVENTS; ++g) { CKC(cudaEventCreate(&ctx->ev[g][0])); CKC(cudaEventCreate(&ctx->ev[g][1])); }
    CKC(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 3; ++i) { CKC(cudaStreamCreateWithFlags(&ctx->s_side[i], cudaStreamNonBlocking)); CKC(cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming)); }
    for (int i = 0; i < 2; ++i) { CKC(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming)); }
    // decimal threshold table d * 10^k (correctly rounded literals via strtod)
    {
        std::vector<double> dec((TSFX_DEC_MAX - TSFX_DEC_MIN + 1) * 9);
        for (int k = TSFX_DEC_MIN; k <= TSFX_DEC_MAX; ++k)
            for (int d = 1; d <= 9; ++d) {
                char buf[32];
                snprintf(buf, sizeof buf, "%de%d", d, k);
                dec[(k - TSFX_DEC_MIN) * 9 + (d - 1)] = strtod(buf, nullptr);
            }
        CKC(cudaMalloc(&ctx->d_dec, dec.size() * sizeof(double)));
        CKC(cudaMemcpy(ctx->d_dec, dec.data(), dec.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
#undef CKC
    *out = ctx;
    return TSFX_OK;
}

extern "C" void tsfx_ctx_destroy(tsfx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->values.release(); ctx->begin.release(); ctx->len.release(); ctx->out.release(); ctx->misc.release(); ctx->stage.release();
    ctx->csr.release();
    ctx->imp.release();
    if (ctx->d_dec) cudaFree(ctx->d_dec);
    if (ctx->d_tw) cudaFree(ctx->d_tw);
    for (int g = 0; g < G_EVENTS; ++g) { if (ctx->ev[g][0]) cudaEventDestroy(ctx->ev[g][0]); if (ctx->ev[g][1]) cudaEventDestroy(ctx->ev[g][1]); }
    for (int i = 0; i < 3; ++i) { if (ctx->s_side[i]) cudaStreamDestroy(ctx->s_side[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (int i = 0; i < 2; ++i) { if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]); if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int tsfx_sync(tsfx_ctx* ctx) {
    if (!ctx) return TSFX_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSFX_OK;
}

// ------------------------------------------------------------------------------------------ plan
extern "C" int tsfx_plan_create(tsfx_ctx* ctx, const tsfx_feature_desc* descs, int32_t n_descs, int32_t n_cols,
                                const double* tables, const int64_t* table_off, const int32_t* table_half,
                                int32_t n_tables, tsfx_plan** out) {
    if (!ctx) return TSFX_E_INVALID;
    if (!out || (!descs && n_descs > 0) || n_descs < 0 || n_cols < 0)
        return fail(ctx, TSFX_E_INVALID, "tsfx_plan_create: bad arguments");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    tsfx_plan* P = new (std::nothrow) tsfx_plan();
    if (!P) return fail(ctx, TSFX_E_NOMEM, "out of host memory");
    P->ctx = ctx;
    P->ncols = n_cols;
    for (int i = 0; i < n_descs; ++i) {
        const Desc& d = descs[i];
        if (d.calc < 0 || d.calc >= TSFX_N_CALCS || d.col < 0 || d.col >= n_cols) {
            delete P;
            return fail(ctx, TSFX_E_INVALID, "tsfx_plan_create: descriptor " + std::to_string(i) + " out of range");
        }
        P->host[group_of(d.calc)].push_back(d);
        switch (d.calc) {
            case TSFX_AUTOCORRELATION: P->lag_needed = std::max(P->lag_needed, d.i0); break;
            case TSFX_AGG_AUTOCORRELATION: P->lag_needed = std::max(P->lag_needed, d.i0); break;
            case TSFX_PARTIAL_AUTOCORRELATION:
                P->lag_needed = std::max(P->lag_needed, d.i1);
                P->pacf_want = std::max(P->pacf_want, d.i1);
                break;
            case TSFX_BINNED_ENTROPY: P->basic_bins = std::max(P->basic_bins, d.i0); break;
            case TSFX_FOURIER_ENTROPY: P->fourier_bins = std::max(P->fourier_bins, d.i0); P->need_welch = 1; break;
            case TSFX_SPKT_WELCH_DENSITY: P->need_welch = 1; break;
            case TSFX_FFT_COEFFICIENT: case TSFX_FFT_AGGREGATED: P->need_fft = 1; break;
            case TSFX_CWT_COEFFICIENTS:
                if (d.i1 < 0 || d.i1 >= n_tables) { delete P; return fail(ctx, TSFX_E_INVALID, "cwt table index out of range"); }
                break;
            case TSFX_AR_COEFFICIENT:
                if (d.i1 < 1 || d.i1 > 32) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "ar_coefficient: k must be in 1..32"); }
                P->max_ar_k = std::max(P->max_ar_k, d.i1);
                break;
            case TSFX_AUGMENTED_DICKEY_FULLER: P->need_adf = 1; break;
            case TSFX_APPROXIMATE_ENTROPY:
                if (d.i0 != 2) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "approximate_entropy: only m=2"); }
                break;
            case TSFX_LEMPEL_ZIV_COMPLEXITY: P->max_lz_bins = std::max(P->max_lz_bins, d.i0); P->n_lz += 1; break;
            case TSFX_PERMUTATION_ENTROPY:
                if (d.i1 < 2 || d.i1 > 8 || d.i0 < 1) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "permutation_entropy: dimension 2..8, tau >= 1"); }
                P->max_perm_dim = std::max(P->max_perm_dim, d.i1);
                break;
            case TSFX_NUMBER_CWT_PEAKS:
                if (d.i0 < 1 || d.i0 > 16) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "number_cwt_peaks: n must be in 1..16"); }
                P->max_cwt_peaks_n = std::max(P->max_cwt_peaks_n, d.i0);
                break;
            case TSFX_FRIEDRICH_COEFFICIENTS: case TSFX_MAX_LANGEVIN_FIXED_POINT:
                if (d.i1 != 3 || d.i2 < 1 || d.i2 > 256) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "friedrich: only m=3, r in 1..256"); }
                P->friedrich_r = std::max(P->friedrich_r, d.i2);
                break;
            default: break;
        }
    }
    if (P->lag_needed > 4096) { delete P; return fail(ctx, TSFX_E_UNSUPPORTED, "lag > 4096"); }
    std::vector<int32_t> final_col;
    for (int g = 0; g < G_COUNT; ++g) {
        std::stable_sort(P->host[g].begin(), P->host[g].end(), [g](const Desc& a, const Desc& b) {
            if (g == G_BASIC) {               // O(1) finishers first (evaluated lane-parallel)
                const bool fa = basic_finisher_calc(a.calc), fb = basic_finisher_calc(b.calc);
                if (fa != fb) return fa;
            }
            if (g == G_SORTED) {
                const bool fa = sorted_finisher_calc(a.calc), fb = sorted_finisher_calc(b.calc);
                if (fa != fb) return fa;
            }
            if (g == G_SPECTRAL) {            // fft_coefficient first, grouped by attribute
                const bool fa = a.calc == TSFX_FFT_COEFFICIENT, fb = b.calc == TSFX_FFT_COEFFICIENT;
                if (fa != fb) return fa;
                if (fa && a.attr != b.attr) return a.attr < b.attr;
            }
            if (a.calc != b.calc) return a.calc < b.calc;
            if (a.i1 != b.i1) return a.i1 < b.i1;
            if (a.i2 != b.i2) return a.i2 < b.i2;
            if (a.p0 != b.p0) return a.p0 < b.p0;
            if (a.p1 != b.p1) return a.p1 < b.p1;
            if (a.i0 != b.i0) return a.i0 < b.i0;
            return a.col < b.col;
        });
        if (g == G_BASIC)
            for (const Desc& d : P->host[g]) P->basic_nfin += basic_finisher_calc(d.calc) ? 1 : 0;
        if (g == G_SORTED)
            for (const Desc& d : P->host[g]) P->sorted_nfin += sorted_finisher_calc(d.calc) ? 1 : 0;
        if (g == G_SPECTRAL)
            for (const Desc& d : P->host[g]) P->spectral_nfft += (d.calc == TSFX_FFT_COEFFICIENT) ? 1 : 0;
        P->cum[g + 1] = P->cum[g] + (int)P->host[g].size();
        for (size_t j = 0; j < P->host[g].size(); ++j) {      // col becomes the index inside the group's staging row
            final_col.push_back(P->host[g][j].col);
            P->host[g][j].col = (int32_t)j;
        }
        if (!P->host[g].empty()) {
            size_t bytes = P->host[g].size() * sizeof(Desc);
            cudaError_t e = cudaMalloc(&P->dev[g], bytes);
            if (e == cudaSuccess) e = cudaMemcpy(P->dev[g], P->host[g].data(), bytes, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { tsfx_plan_destroy(P); return fail(ctx, TSFX_E_CUDA, cudaGetErrorString(e)); }
        }
    }
    if (!final_col.empty()) {
        cudaError_t e = cudaMalloc(&P->d_final_col, final_col.size() * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(P->d_final_col, final_col.data(), final_col.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { tsfx_plan_destroy(P); return fail(ctx, TSFX_E_CUDA, cudaGetErrorString(e)); }
    }
    if (n_tables > 0) {
        if (!tables || !table_off || !table_half) { tsfx_plan_destroy(P); return fail(ctx, TSFX_E_INVALID, "cwt tables missing"); }
        P->n_tables = n_tables;
        P->toff.assign(table_off, table_off + n_tables + 1);
        P->thalf.assign(table_half, table_half + n_tables);
        size_t tb = (size_t)table_off[n_tables] * sizeof(double);
        cudaError_t e = cudaMalloc(&P->d_tables, std::max<size_t>(tb, 8));
        if (e == cudaSuccess) e = cudaMemcpy(P->d_tables, tables, tb, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&P->d_toff, (n_tables + 1) * sizeof(int64_t));
        if (e == cudaSuccess) e = cudaMemcpy(P->d_toff, table_off, (n_tables + 1) * sizeof(int64_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&P->d_thalf, n_tables * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(P->d_thalf, table_half, n_tables * sizeof(int32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { tsfx_plan_destroy(P); return fail(ctx, TSFX_E_CUDA, cudaGetErrorString(e)); }
    }
    *out = P;
    return TSFX_OK;
}

extern "C" void tsfx_plan_destroy(tsfx_plan* P) {
    if (!P) return;
    if (P->ctx) cudaSetDevice(P->ctx->device);
    for (int g = 0; g < G_COUNT; ++g) if (P->dev[g]) cudaFree(P->dev[g]);
    if (P->d_final_col) cudaFree(P->d_final_col);
    if (P->d_tables) cudaFree(P->d_tables);
    if (P->d_toff) cudaFree(P->d_toff);
    if (P->d_thalf) cudaFree(P->d_thalf);
    delete P;
}

// ------------------------------------------------------------------------------------------ launch all groups
static int even(int v) { return (v + 1) & ~1; }

static int ensure_twiddle(tsfx_ctx* ctx, int n_pow2) {
    if (n_pow2 <= ctx->tw_n) return TSFX_OK;
    if (ctx->d_tw) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_tw); ctx->d_tw = nullptr; ctx->tw_n = 0; }
    CK(cudaMalloc(&ctx->d_tw, (size_t)(n_pow2 / 2 + 1) * sizeof(double2)));
    CK(launch_fill_twiddle(ctx->d_tw, n_pow2, ctx->stream));
    ctx->tw_n = n_pow2;
    return TSFX_OK;
}

static int run_groups(tsfx_ctx* ctx, const tsfx_plan* P, const SeriesRef& R, int max_len, double* d_out, uint32_t flags) {
    const bool timing = (flags & TSFX_FLAG_TIMING) != 0;
    for (int g = 0; g < G_EVENTS; ++g) ctx->ev_used[g] = false;
    ctx->launches = 0;
    if (R.n_series == 0) return TSFX_OK;
    const int staged = P->cum[G_COUNT];
    if (staged == 0) return TSFX_OK;
    CK(ctx->stage.reserve((size_t)R.n_series * staged * sizeof(double)));
    CK(ctx->misc.reserve(max_len > 1024 ? ((size_t)1 << 30) : ((size_t)256 << 20)));      // global working regions for series too long for shared memory
    double* const d_final = d_out;
    (void)d_final;
    if (!P->host[G_SPECTRAL].empty()) {      // FFT twiddle table (filled once, on the main stream, before any fork)
        int p2 = 1;
        while (p2 < max_len) p2 <<= 1;
        if (p2 > max_len) p2 >>= 1;            // largest power of two <= max_len
        p2 = std::max(p2, 256);
        int rc = ensure_twiddle(ctx, p2);
        if (rc) return rc;
    }
    // kernel groups are independent (own staging matrix): optionally spread them over side streams
    const int nstreams = timing ? 1 : env_streams();
    if (nstreams > 1) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        for (int i = 0; i < nstreams - 1; ++i) CK(cudaStreamWaitEvent(ctx->s_side[i], ctx->ev_fork, 0));
    }
    int launched = 0;
    if (max_len < 1) return fail(ctx, TSFX_E_INVALID, "series of length < 1");
    auto too_long = [&](const char* g) {
        return fail(ctx, TSFX_E_TOO_LONG, std::string("series length ") + std::to_string(max_len) +
                                              " exceeds the shared-memory staging of kernel group " + g);
    };
    for (int g = 0; g < G_COUNT; ++g) {
        if (P->host[g].empty()) continue;
        if (timing) { CK(cudaEventRecord(ctx->ev[g][0], ctx->stream)); }
        cudaError_t e = cudaSuccess;
        double* d_out = (double*)ctx->stage.p + (size_t)R.n_series * P->cum[g];      // this group's staging matrix
        const int sidx = launched++ % nstreams;
        cudaStream_t gs = (sidx == 0) ? ctx->stream : ctx->s_side[sidx - 1];
        const int g_ncols = (int)P->host[g].size();
        switch (g) {
            case G_BASIC: {
                BasicArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.lag_needed = P->lag_needed;
                A.nfin = P->basic_nfin;
                int pac = P->pacf_want >= 0 ? 4 * (P->pacf_want + 1) : 0;
                A.pacf_off = P->lag_needed + 1;
                A.nlag = even(P->lag_needed + 1 + pac);
                A.nscr = even(std::max(std::max(max_len, 64), (P->basic_bins + 1) / 2));
                A.nalt = 0;
                {
                    int prev = -1;
                    for (const Desc& q : P->host[g])
                        if (q.calc == TSFX_AGG_LINEAR_TREND) {
                            const int key = (q.i0 << 4) | q.i1;
                            if (key != prev) { ++A.nalt; prev = key; }
                        }
                }
                A.dec = ctx->d_dec;
                e = launch_basic(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_SORTED: {
                SortedArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.nscr = even(4 * (P->friedrich_r + 2) + 16);
                A.nfin = P->sorted_nfin;
                A.ncq = 0;
                {
                    double pl = -1.0, ph = -1.0;
                    for (const Desc& q : P->host[g])
                        if (q.calc == TSFX_CHANGE_QUANTILES && !(q.p0 == pl && q.p1 == ph)) { ++A.ncq; pl = q.p0; ph = q.p1; }
                }
                e = launch_sorted(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_SPECTRAL: {
                SpectralArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.twiddle = ctx->d_tw; A.tw_n = ctx->tw_n;
                A.tables = P->d_tables; A.table_off = P->d_toff; A.table_half = P->d_thalf;
                A.need_fft = P->need_fft; A.need_welch = P->need_welch;
                A.max_hist = P->fourier_bins;
                A.nfft = P->spectral_nfft;
                e = launch_spectral(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_LA: {
                LaArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.nscr = P->max_ar_k;
                e = launch_la(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_ENTROPY: {
                EntropyArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                e = launch_entropy(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_SEQ: {
                SeqArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.nscr = (P->max_lz_bins > 0 ? 1 : 0) | (P->max_perm_dim > 0 ? 2 : 0) | (P->max_cwt_peaks_n << 8) |
                         (std::min(P->n_lz, 255) << 16);
                e = launch_seq(A, max_len, gs, ctx->sm_count);
                break;
            }
            case G_PEAKS: {
                SeqArgs A;
                A.R = R; A.gscratch = (unsigned char*)ctx->misc.p; A.gscratch_bytes = ctx->misc.cap; A.descs = P->dev[g]; A.nd = (int)P->host[g].size(); A.out = d_out; A.ncols = g_ncols;
                A.nscr = (P->max_cwt_peaks_n << 8);
                e = launch_peaks(A, max_len, gs, ctx->sm_count);
                break;
            }
        }
        if (e == cudaErrorInvalidConfiguration) return too_long(kGroupNames[g]);
        if (e != cudaSuccess) return fail(ctx, TSFX_E_CUDA, std::string("launch ") + kGroupNames[g] + ": " + cudaGetErrorString(e));
        ctx->launches += 1;
        if (timing) { CK(cudaEventRecord(ctx->ev[g][1], ctx->stream)); ctx->ev_used[g] = true; }
    }
    if (nstreams > 1)
        for (int i = 0; i < nstreams - 1; ++i) {
            CK(cudaEventRecord(ctx->ev_join[i], ctx->s_side[i]));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
        }
    {   // scatter the staging matrices into the caller's [n_series x ncols] matrix
        if (timing) { CK(cudaEventRecord(ctx->ev[G_COUNT][0], ctx->stream)); }
        AssembleArgs A;
        A.stage = (const double*)ctx->stage.p; A.out = d_final; A.n_series = R.n_series; A.ncols = P->ncols;
        A.n_groups = G_COUNT;
        for (int g = 0; g <= G_COUNT; ++g) A.cum[g] = P->cum[g];
        A.final_col = P->d_final_col;
        cudaError_t e = launch_assemble(A, ctx->stream, ctx->sm_count);
        if (e != cudaSuccess) return fail(ctx, TSFX_E_CUDA, std::string("launch assemble: ") + cudaGetErrorString(e));
        ctx->launches += 1;
        if (timing) { CK(cudaEventRecord(ctx->ev[G_COUNT][1], ctx->stream)); ctx->ev_used[G_COUNT] = true; }
    }
    return TSFX_OK;
}

static int impute_after_extract(tsfx_ctx* ctx, double* d_out, int64_t rows, int cols);

static int check_args(tsfx_ctx* ctx, const tsfx_plan* plan, const void* values, const void* out, int64_t n_series) {
    if (!ctx) return TSFX_E_INVALID;
    if (!plan || plan->ctx != ctx) return fail(ctx, TSFX_E_INVALID, "plan does not belong to this context");
    if (n_series < 0) return fail(ctx, TSFX_E_INVALID, "n_series < 0");
    if (n_series > 0 && (!values || !out)) return fail(ctx, TSFX_E_INVALID, "NULL values/out");
    return TSFX_OK;
}

extern "C" int tsfx_extract_csr(tsfx_ctx* ctx, const tsfx_plan* plan, const float* values, int64_t n_values,
                                const int64_t* begin, const int32_t* len, int64_t n_series, double* out,
                                uint32_t flags) {
    int rc = check_args(ctx, plan, values, out, n_series);
    if (rc) return rc;
    if (n_series == 0) return TSFX_OK;
    if (!begin || !len) return fail(ctx, TSFX_E_INVALID, "NULL begin/len");
    CK(cudaSetDevice(ctx->device));
    SeriesRef R;
    R.dense_len = 0;
    R.n_series = n_series;
    int max_len = 0;
    if (flags & TSFX_FLAG_DEVICE_PTRS) {
        R.values = values; R.begin = begin; R.len = len;
        int rc2 = csr_max_len(ctx->csr, len, n_series, ctx->stream, &max_len);
        if (rc2) return fail(ctx, TSFX_E_CUDA, "max-length reduction failed");
        rc = run_groups(ctx, plan, R, max_len, out, flags);
        if (!rc && (flags & TSFX_FLAG_IMPUTE)) rc = impute_after_extract(ctx, out, n_series, plan->ncols);
        return rc;
    }
    for (int64_t s = 0; s < n_series; ++s) {
        if (len[s] < 1 || begin[s] < 0 || begin[s] + len[s] > n_values)
            return fail(ctx, TSFX_E_INVALID, "series " + std::to_string(s) + " has an invalid (begin, len)");
        max_len = std::max(max_len, (int)len[s]);
    }
    size_t ob = (size_t)n_series * plan->ncols * sizeof(double);
    CK(ctx->values.reserve((size_t)n_values * sizeof(float) + 16));
    CK(ctx->begin.reserve((size_t)n_series * sizeof(int64_t)));
    CK(ctx->len.reserve((size_t)n_series * sizeof(int32_t)));
    CK(ctx->out.reserve(std::max<size_t>(ob, 8)));
    CK(cudaMemcpyAsync(ctx->values.p, values, (size_t)n_values * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->begin.p, begin, (size_t)n_series * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->len.p, len, (size_t)n_series * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    R.values = (const float*)ctx->values.p; R.begin = (const int64_t*)ctx->begin.p; R.len = (const int32_t*)ctx->len.p;
    rc = run_groups(ctx, plan, R, max_len, (double*)ctx->out.p, flags);
    if (rc) return rc;
    if (flags & TSFX_FLAG_IMPUTE) { rc = impute_after_extract(ctx, (double*)ctx->out.p, n_series, plan->ncols); if (rc) return rc; }
    CK(cudaMemcpyAsync(out, ctx->out.p, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSFX_OK;
}

extern "C" int tsfx_extract_dense(tsfx_ctx* ctx, const tsfx_plan* plan, const float* values, int64_t n_series,
                                  int32_t len, double* out, uint32_t flags) {
    int rc = check_args(ctx, plan, values, out, n_series);
    if (rc) return rc;
    if (n_series == 0) return TSFX_OK;
    if (len < 1) return fail(ctx, TSFX_E_INVALID, "len < 1");
    CK(cudaSetDevice(ctx->device));
    SeriesRef R;
    R.begin = nullptr; R.len = nullptr; R.dense_len = len; R.n_series = n_series;
    if (flags & TSFX_FLAG_DEVICE_PTRS) {
        R.values = values;
        rc = run_groups(ctx, plan, R, len, out, flags);
        if (!rc && (flags & TSFX_FLAG_IMPUTE)) rc = impute_after_extract(ctx, out, n_series, plan->ncols);
        return rc;
    }
    // host path: pipelined over row blocks -- the H2D copy of block b+1 and the D2H copy of block b-1 run on
    // their own streams while the kernels of block b execute (pinned host buffers make the copies truly async)
    const size_t ncols = (size_t)plan->ncols;
    const size_t vb = (size_t)n_series * len * sizeof(float), ob = (size_t)n_series * ncols * sizeof(double);
    CK(ctx->values.reserve(vb + 16));
    CK(ctx->out.reserve(std::max<size_t>(ob, 8)));
    int64_t block = std::max<int64_t>(16384, (n_series + 15) / 16);
    if (flags & TSFX_FLAG_TIMING) block = n_series;            // per-group events describe one whole pass
    float* dv = (float*)ctx->values.p;
    double* dout = (double*)ctx->out.p;
    int nb = 0;
    for (int64_t lo = 0; lo < n_series; lo += block, ++nb) {
        const int64_t cnt = std::min<int64_t>(block, n_series - lo);
        const int slot = nb & 1;
        CK(cudaMemcpyAsync(dv + (size_t)lo * len, values + (size_t)lo * len, (size_t)cnt * len * sizeof(float),
                           cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0));
        R.values = dv + (size_t)lo * len;
        R.n_series = cnt;
        rc = run_groups(ctx, plan, R, len, dout + (size_t)lo * ncols, flags);
        if (rc) return rc;
        if (flags & TSFX_FLAG_IMPUTE) continue;         // column statistics need every row: one copy at the end
        CK(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[slot], 0));
        CK(cudaMemcpyAsync(out + (size_t)lo * ncols, dout + (size_t)lo * ncols, (size_t)cnt * ncols * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->s_out));
    }
    if (flags & TSFX_FLAG_IMPUTE) {
        rc = impute_after_extract(ctx, dout, n_series, plan->ncols);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out, dout, ob, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSFX_OK;
}

// impute the device matrix of the extract call that just ran (TSFX_FLAG_IMPUTE)
static int impute_after_extract(tsfx_ctx* ctx, double* d_out, int64_t rows, int cols) {
    int n = 0;
    cudaError_t e = impute_device(ctx->imp, d_out, rows, cols, TSFX_IMPUTE_RANGE, false, nullptr, ctx->sm_count, ctx->stream, &n);
    if (e != cudaSuccess) return fail(ctx, TSFX_E_CUDA, std::string("impute: ") + cudaGetErrorString(e));
    ctx->launches += n;
    return TSFX_OK;
}

extern "C" int tsfx_impute(tsfx_ctx* ctx, double* matrix, int64_t n_rows, int32_t n_cols, int32_t mode, double* col_stats,
                           uint32_t flags) {
    if (!ctx) return TSFX_E_INVALID;
    if (n_rows < 0 || n_cols < 0 || mode < TSFX_IMPUTE_RANGE || mode > TSFX_IMPUTE_STATS)
        return fail(ctx, TSFX_E_INVALID, "tsfx_impute: bad arguments");
    if (n_rows == 0 || n_cols == 0) return TSFX_OK;
    if (!matrix) return fail(ctx, TSFX_E_INVALID, "tsfx_impute: NULL matrix");
    if ((mode == TSFX_IMPUTE_GIVEN || mode == TSFX_IMPUTE_STATS) && !col_stats)
        return fail(ctx, TSFX_E_INVALID, "tsfx_impute: col_stats is required for this mode");
    if (mode == TSFX_IMPUTE_GIVEN)
        for (int64_t i = 0; i < (int64_t)3 * n_cols; ++i)
            if (!std::isfinite(col_stats[i]))       // dataframe_functions.py:147-156 raises ValueError
                return fail(ctx, TSFX_E_INVALID, "tsfx_impute: non-finite replacement value");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n_rows * n_cols * sizeof(double);
    double* d_m = matrix;
    const bool host = !(flags & TSFX_FLAG_DEVICE_PTRS);
    if (host) {
        CK(ctx->out.reserve(bytes));
        d_m = (double*)ctx->out.p;
        CK(cudaMemcpyAsync(d_m, matrix, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    int n = 0;
    cudaError_t e = impute_device(ctx->imp, d_m, n_rows, n_cols, mode, (flags & TSFX_FLAG_ALL_MEDIANS) != 0 || mode == TSFX_IMPUTE_STATS,
                                  col_stats, ctx->sm_count, ctx->stream, &n);
    if (e != cudaSuccess) return fail(ctx, TSFX_E_CUDA, std::string("impute: ") + cudaGetErrorString(e));
    ctx->launches = n;
    if (host) {
        if (mode != TSFX_IMPUTE_STATS) CK(cudaMemcpyAsync(matrix, d_m, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return TSFX_OK;
}

extern "C" int tsfx_get_timings(tsfx_ctx* ctx, float* ms_out, const char** names_out, int32_t cap) {
    if (!ctx) return TSFX_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    int k = 0;
    for (int g = 0; g < G_EVENTS && k < cap; ++g) {
        if (!ctx->ev_used[g]) continue;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev[g][0], ctx->ev[g][1]));
        if (ms_out) ms_out[k] = ms;
        if (names_out) names_out[k] = kGroupNames[g];
        ++k;
    }
    return k;
}

extern "C" int tsfx_last_launch_count(const tsfx_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------ stage (a)
extern "C" int tsfx_build_csr(tsfx_ctx* ctx, const int64_t* ids, const void* sort_keys, int32_t sort_key_is_f64,
                              const float* values, int64_t n_rows, int64_t* out_ids, int64_t* out_begin,
                              int32_t* out_len, float* sorted_values, int64_t out_capacity, int64_t* n_series_out) {
    if (!ctx) return TSFX_E_INVALID;
    if (n_rows < 0 || !n_series_out || (n_rows > 0 && (!ids || !values)))
        return fail(ctx, TSFX_E_INVALID, "tsfx_build_csr: bad arguments");
    CK(cudaSetDevice(ctx->device));
    *n_series_out = 0;
    if (n_rows == 0) return TSFX_OK;
    std::string msg;
    int64_t ns = 0;
    ctx->held_series = -1;
    int rc = csr_build_from_host(ctx->csr, ids, sort_keys, sort_key_is_f64, values, n_rows, ctx->stream, &ns, &msg);
    if (rc) return fail(ctx, rc, msg);
    ctx->held_series = ns;
    *n_series_out = ns;
    if (!out_ids && !out_begin && !out_len && !sorted_values) return TSFX_OK;      // count + keep on device
    if (ns > out_capacity) return fail(ctx, TSFX_E_INVALID, "out_capacity too small: " + std::to_string(ns) + " series");
    if (out_ids) CK(cudaMemcpyAsync(out_ids, ctx->csr.d_uid, ns * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_begin) CK(cudaMemcpyAsync(out_begin, ctx->csr.d_begin, ns * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_len) CK(cudaMemcpyAsync(out_len, ctx->csr.d_len, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (sorted_values) CK(cudaMemcpyAsync(sorted_values, ctx->csr.d_values, n_rows * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSFX_OK;
}

extern "C" int tsfx_extract_long(tsfx_ctx* ctx, const tsfx_plan* plan, const int64_t* ids, const void* sort_keys,
                                 int32_t sort_key_is_f64, const float* values, int64_t n_rows, int64_t* out_ids,
                                 double* out, int64_t out_capacity, int64_t* n_series_out, uint32_t flags) {
    if (!ctx) return TSFX_E_INVALID;
    if (!plan || plan->ctx != ctx) return fail(ctx, TSFX_E_INVALID, "plan does not belong to this context");
    if (flags & TSFX_FLAG_DEVICE_PTRS) return fail(ctx, TSFX_E_INVALID, "tsfx_extract_long takes host pointers");
    const bool reuse = (ids == nullptr && values == nullptr);     // run on the CSR held from tsfx_build_csr
    if (n_rows < 0 || !n_series_out || (!reuse && n_rows > 0 && (!ids || !values)) || !out)
        return fail(ctx, TSFX_E_INVALID, "tsfx_extract_long: bad arguments");
    CK(cudaSetDevice(ctx->device));
    *n_series_out = 0;
    std::string msg;
    int64_t ns = 0;
    int rc;
    if (reuse) {
        if (ctx->held_series < 0) return fail(ctx, TSFX_E_INVALID, "tsfx_extract_long: no CSR is held by this context");
        ns = ctx->held_series;
    } else {
        if (n_rows == 0) return TSFX_OK;
        ctx->held_series = -1;
        rc = csr_build_from_host(ctx->csr, ids, sort_keys, sort_key_is_f64, values, n_rows, ctx->stream, &ns, &msg);
        if (rc) return fail(ctx, rc, msg);
        ctx->held_series = ns;
    }
    *n_series_out = ns;
    if (ns == 0) return TSFX_OK;
    if (ns > out_capacity) return fail(ctx, TSFX_E_INVALID, "out_capacity too small: " + std::to_string(ns) + " series");
    int max_len = 0;
    if (csr_max_len(ctx->csr, ctx->csr.d_len, ns, ctx->stream, &max_len)) return fail(ctx, TSFX_E_CUDA, "max-length reduction failed");
    size_t ob = (size_t)ns * plan->ncols * sizeof(double);
    CK(ctx->out.reserve(std::max<size_t>(ob, 8)));
    SeriesRef R;
    R.values = ctx->csr.d_values; R.begin = ctx->csr.d_begin; R.len = ctx->csr.d_len; R.dense_len = 0; R.n_series = ns;
    rc = run_groups(ctx, plan, R, max_len, (double*)ctx->out.p, flags);
    if (rc) return rc;
    if (flags & TSFX_FLAG_IMPUTE) { rc = impute_after_extract(ctx, (double*)ctx->out.p, ns, plan->ncols); if (rc) return rc; }
    if (out_ids) CK(cudaMemcpyAsync(out_ids, ctx->csr.d_uid, ns * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->out.p, ob, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSFX_OK;
}

// ------------------------------------------------------------------------------------------ roll_time_series views
extern "C" int64_t tsfx_roll_windows(const int64_t* begin, const int32_t* len, int64_t n_series,
                                     int32_t rolling_direction, int32_t max_timeshift, int32_t min_timeshift,
                                     int64_t* win_begin, int32_t* win_len, int64_t* win_parent,
                                     int32_t* win_end_index, int64_t capacity) {
    // dataframe_functions.py:340-373, 548-562.  rolling_direction > 0: the shifts are
    // reversed(range(Lmax, 0, -rolling_direction)) where Lmax is the LONGEST series of the frame (:555-560), so
    // window ends are anchored to Lmax for every series; shift t applies to a series of length L when t <= L,
    // the window is rows [max(t-max_timeshift-1, 0), t), kept when it has at least min_timeshift+1 rows, and
    // its id is (parent id, time of row t-1).
    // rolling_direction < 0 (:351-356, 365-366): shifts range(1, Lmax+1, |rolling_direction|), the window is rows
    // [t-1, min(t+max_timeshift, L)), same minimum length, id = (parent id, time of row t-1) -- the window's FIRST row.
    // win_end_index is therefore "the row whose sort value names the window": last row (positive) / first row (negative).
    if (!begin || !len || n_series < 0 || rolling_direction == 0 || max_timeshift < 0 || min_timeshift < 0)
        return TSFX_E_INVALID;
    int32_t Lmax = 0;
    for (int64_t s = 0; s < n_series; ++s) { if (len[s] < 1) return TSFX_E_INVALID; Lmax = std::max(Lmax, len[s]); }
    const int32_t amount = rolling_direction > 0 ? rolling_direction : -rolling_direction;
    const int32_t first = rolling_direction > 0 ? (Lmax > 0 ? Lmax - ((Lmax - 1) / amount) * amount : 1) : 1;   // smallest shift
    int64_t k = 0;
    for (int64_t s = 0; s < n_series; ++s) {
        const int32_t L = len[s];
        for (int32_t t = first; t <= L; t += amount) {
            int32_t lo, wl;
            if (rolling_direction > 0) {
                lo = t - max_timeshift - 1;
                if (lo < 0) lo = 0;
                wl = t - lo;
            } else {
                lo = t - 1;
                const int64_t hi = std::min<int64_t>((int64_t)lo + max_timeshift + 1, L);
                wl = (int32_t)(hi - lo);
            }
            if (wl < min_timeshift + 1) continue;
            if (win_begin) {
                if (k >= capacity) return TSFX_E_INVALID;
                win_begin[k] = begin[s] + lo;
                win_len[k] = wl;
                if (win_parent) win_parent[k] = s;
                if (win_end_index) win_end_index[k] = t - 1;
            }
            ++k;
        }
    }
    return k;
}
