// Level-1 ABI: the Fortran-style symbols the reference's Julia wrapper ccalls in
// MUMPS/src/MUMPSfuncs.jl (factor :32-35/:49-52, solve :105-107/:128-130, sparse rhs :115-118/:139-143,
// destroy :155-156/:170-171).  The factorisation runs on the B200 band kernel; matrices must be
// structurally symmetric (sym = 1 or 2; the hot path passes sym = 1 for complex-symmetric Aii).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/hmcmt_b200.h"
#include "band_factor.cuh"
#include "band_solve.cuh"

namespace hmcmt {
int round_T(int b);
int launch_factor(cudaStream_t st, int T, const BandSys* sys, int nsys, const BandDom& dom, const SolveJob* fwdJobs, int* nLaunches);
size_t big_work_bytes(int T);
int launch_solve(cudaStream_t st, int T, const SolveJob* jobs, int njobs, const BandDom& dom);
int max_band_T();
}  // namespace hmcmt
using namespace hmcmt;

namespace {

struct Factor {
    int n = 0, b = 0, T = 0, S = 0;
    bool isReal = false;
    double* panels = nullptr;
    cplx* ainvz = nullptr;
};
std::mutex g_mu;
std::unordered_map<int64_t, Factor*> g_factors;
int64_t g_next = 1;

void free_factor(Factor* f) {
    if (!f) return;
    if (f->panels) cudaFree(f->panels);
    if (f->ainvz) cudaFree(f->ainvz);
    delete f;
}

// common factor path; vals are complex (re,im) or real depending on isReal
int64_t factor_common(int64_t n, int64_t sym, const double* nzval, const int64_t* rowval, const int64_t* colptr, bool isReal,
                      int64_t* status) {
    auto fail = [&](int code) { if (status) *status = code; return (int64_t)0; };
    if (status) *status = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[hmcmt_b200] no CUDA device: this library has no CPU fallback\n");
        return fail(kErrNoDevice);
    }
    if (n < 1 || !nzval || !rowval || !colptr) return fail(kErrArg);
    if (sym != 1 && sym != 2) {
        fprintf(stderr, "[hmcmt_b200] factor_mumps: only symmetric matrices (sym=1,2) are supported\n");
        return fail(kErrArg);
    }
    // half-bandwidth from the pattern (1-based CSC, MUMPSfuncs.jl:32-35)
    int64_t b = 0;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) b = std::max<int64_t>(b, std::llabs(rowval[k] - 1 - j));
    if (b < 1) b = 1;
    int T = round_T((int)b);
    if (T > max_band_T()) {
        fprintf(stderr, "[hmcmt_b200] factor_mumps: half-bandwidth %lld exceeds the supported maximum %d\n",
                (long long)b, 8 * max_band_T() - 32);
        return fail(kErrArg);
    }
    // lower band image: band[g*(b+1)+d] = A[g][g-d]   (the lower triangle is what LDL^T reads)
    std::vector<cplx> band((size_t)n * (b + 1), mk(0.0, 0.0));
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) {
            int64_t i = rowval[k] - 1;
            if (i < j) continue;
            cplx v = isReal ? mk(nzval[k], 0.0) : mk(nzval[2 * k], nzval[2 * k + 1]);
            band[(size_t)i * (b + 1) + (i - j)] = v;
        }
    Factor* f = new Factor();
    f->n = (int)n; f->b = (int)b; f->T = T; f->S = (int)((n + TS - 1) / TS); f->isReal = isReal;
    cplx* dband = nullptr;
    cplx* dbig = nullptr;
    BandSys* dsys = nullptr;
    int* dstatus = nullptr;
    auto cleanup = [&]() { if (dband) cudaFree(dband); if (dbig) cudaFree(dbig); if (dsys) cudaFree(dsys); if (dstatus) cudaFree(dstatus); };
    if (cudaMalloc(&dband, band.size() * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc(&f->panels, (size_t)f->S * panel_doubles(T) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&f->ainvz, (size_t)f->S * AZ * sizeof(cplx)) != cudaSuccess || cudaMalloc(&dsys, sizeof(BandSys)) != cudaSuccess ||
        cudaMalloc(&dstatus, sizeof(int)) != cudaSuccess || (big_work_bytes(T) && cudaMalloc(&dbig, big_work_bytes(T)) != cudaSuccess)) {
        cleanup(); free_factor(f);
        return fail(kErrAlloc);
    }
    cudaMemcpy(dband, band.data(), band.size() * sizeof(cplx), cudaMemcpyHostToDevice);
    cudaMemset(dstatus, 0, sizeof(int));
    BandSys s{};
    s.band = dband; s.omega = 0.0; s.rhs = nullptr; s.panels[0] = f->panels; s.panels[1] = nullptr; s.ainvz[0] = f->ainvz; s.ainvz[1] = nullptr; s.wexp = nullptr; s.big = dbig; s.x = nullptr; s.status = dstatus;
    cudaMemcpy(dsys, &s, sizeof(BandSys), cudaMemcpyHostToDevice);
    BandDom dom{(int)n, (int)b, (int)b, 0, 0, 0};
    int rc = launch_factor(nullptr, T, dsys, 1, dom, nullptr, nullptr);
    int hst = 0;
    if (rc == kOk && cudaDeviceSynchronize() != cudaSuccess) rc = kErrCuda;
    if (rc == kOk) cudaMemcpy(&hst, dstatus, sizeof(int), cudaMemcpyDeviceToHost);
    cleanup();
    if (rc != kOk || hst != 0) { free_factor(f); return fail(rc != kOk ? rc : hst); }
    std::lock_guard<std::mutex> lk(g_mu);
    int64_t h = g_next++;
    g_factors[h] = f;
    return h;
}

Factor* lookup(int64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_factors.find(h);
    return it == g_factors.end() ? nullptr : it->second;
}

// dense rhs (complex interleaved or real), n x nrhs column-major
int64_t solve_common(int64_t h, int64_t nrhs, const double* rhs, double* x, bool isRealIO) {
    Factor* f = lookup(h);
    if (!f || nrhs < 1 || !rhs || !x) return kErrArg;
    const size_t n = f->n;
    std::vector<cplx> hb(n * nrhs);
    for (size_t i = 0; i < n * (size_t)nrhs; ++i) hb[i] = isRealIO ? mk(rhs[i], 0.0) : mk(rhs[2 * i], rhs[2 * i + 1]);
    cplx *dx = nullptr, *dz = nullptr;
    SolveJob* djobs = nullptr;
    auto cleanup = [&]() { if (dx) cudaFree(dx); if (dz) cudaFree(dz); if (djobs) cudaFree(djobs); };
    if (cudaMalloc(&dx, hb.size() * sizeof(cplx)) != cudaSuccess || cudaMalloc(&dz, (size_t)nrhs * f->S * 8 * sizeof(cplx)) != cudaSuccess ||
        cudaMalloc(&djobs, sizeof(SolveJob) * nrhs) != cudaSuccess) {
        cleanup();
        return kErrAlloc;
    }
    cudaMemcpy(dx, hb.data(), hb.size() * sizeof(cplx), cudaMemcpyHostToDevice);
    std::vector<SolveJob> jobs(nrhs);
    for (int64_t r = 0; r < nrhs; ++r) {
        jobs[r].panels[0] = f->panels; jobs[r].panels[1] = nullptr; jobs[r].ainvz[0] = f->ainvz; jobs[r].ainvz[1] = nullptr;
        jobs[r].rhs = dx + r * n; jobs[r].x = dx + r * n; jobs[r].zbuf[0] = dz + (size_t)r * f->S * 8; jobs[r].zbuf[1] = nullptr;
        jobs[r].wexp = nullptr;
    }
    cudaMemcpy(djobs, jobs.data(), sizeof(SolveJob) * nrhs, cudaMemcpyHostToDevice);
    BandDom dom{f->n, f->b, f->b, 0, 0, 0};
    int rc = launch_solve(nullptr, f->T, djobs, (int)nrhs, dom);
    if (rc == kOk && cudaDeviceSynchronize() != cudaSuccess) rc = kErrCuda;
    if (rc == kOk) {
        cudaMemcpy(hb.data(), dx, hb.size() * sizeof(cplx), cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < hb.size(); ++i) {
            if (isRealIO) x[i] = hb[i].x;
            else { x[2 * i] = hb[i].x; x[2 * i + 1] = hb[i].y; }
        }
    }
    cleanup();
    return rc;
}

void solve_sparse_common(int64_t h, int64_t nrhs, const double* nzval, const int64_t* rowval, const int64_t* colptr, double* x,
                         bool isRealIO) {
    Factor* f = lookup(h);
    if (!f || nrhs < 1) return;
    const size_t n = f->n, w = isRealIO ? 1 : 2;
    std::vector<double> dense(n * nrhs * w, 0.0);
    for (int64_t j = 0; j < nrhs; ++j)
        for (int64_t k = colptr[j] - 1; k < colptr[j + 1] - 1; ++k) {
            size_t i = (size_t)(rowval[k] - 1) + (size_t)j * n;
            if (isRealIO) dense[i] = nzval[k];
            else { dense[2 * i] = nzval[2 * k]; dense[2 * i + 1] = nzval[2 * k + 1]; }
        }
    solve_common(h, nrhs, dense.data(), x, isRealIO);
}

int64_t destroy_common(int64_t h) {
    Factor* f = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_factors.find(h);
        if (it == g_factors.end()) return kErrArg;
        f = it->second;
        g_factors.erase(it);
    }
    free_factor(f);
    return kOk;
}

}  // namespace

extern "C" {

int64_t factor_mumps_cmplx_(const int64_t* n, const int64_t* sym, const int64_t* ooc, const double* nzval, const int64_t* rowval,
                            const int64_t* colptr, int64_t* status) {
    (void)ooc;
    if (!n || !sym) { if (status) *status = kErrArg; return 0; }
    return factor_common(*n, *sym, nzval, rowval, colptr, false, status);
}
int64_t factor_mumps_(const int64_t* n, const int64_t* sym, const int64_t* ooc, const double* nzval, const int64_t* rowval,
                      const int64_t* colptr, int64_t* status) {
    (void)ooc;
    if (!n || !sym) { if (status) *status = kErrArg; return 0; }
    return factor_common(*n, *sym, nzval, rowval, colptr, true, status);
}
int64_t solve_mumps_cmplx_(const int64_t* handle, const int64_t* nrhs, const double* rhs, double* x, const int64_t* transpose) {
    (void)transpose;      // A is symmetric: A^T x = b is the same system (compJacTMatVec.jl:221-224)
    if (!handle || !nrhs) return kErrArg;
    return solve_common(*handle, *nrhs, rhs, x, false);
}
int64_t solve_mumps_(const int64_t* handle, const int64_t* nrhs, const double* rhs, double* x, const int64_t* transpose) {
    (void)transpose;
    if (!handle || !nrhs) return kErrArg;
    return solve_common(*handle, *nrhs, rhs, x, true);
}
void solve_mumps_sparse_rhs_(const int64_t* handle, const int64_t* nzrhs, const int64_t* nrhs, const double* nzval,
                             const int64_t* rowval, const int64_t* colptr, double* x, const int64_t* transpose) {
    (void)nzrhs; (void)transpose;
    if (!handle || !nrhs) return;
    solve_sparse_common(*handle, *nrhs, nzval, rowval, colptr, x, true);
}
void solve_mumps_cmplx_sparse_rhs_(const int64_t* handle, const int64_t* nzrhs, const int64_t* nrhs, const double* nzval,
                                   const int64_t* rowval, const int64_t* colptr, double* x, const int64_t* transpose) {
    (void)nzrhs; (void)transpose;
    if (!handle || !nrhs) return;
    solve_sparse_common(*handle, *nrhs, nzval, rowval, colptr, x, false);
}
// debug: copy the raw factor (panel images, pivot-block inverses) back to the host
int64_t hmcmt_debug_get_factor(int64_t handle, double* panels, double* ainv, int64_t* dims) {
    Factor* f = lookup(handle);
    if (!f) return kErrArg;
    if (dims) { dims[0] = f->n; dims[1] = f->b; dims[2] = f->T; dims[3] = f->S; }
    if (panels) cudaMemcpy(panels, f->panels, (size_t)f->S * panel_doubles(f->T) * sizeof(double), cudaMemcpyDeviceToHost);
    if (ainv) cudaMemcpy(ainv, f->ainvz, (size_t)f->S * AZ * sizeof(cplx), cudaMemcpyDeviceToHost);
    return kOk;
}
int64_t destroy_mumps_(const int64_t* handle) { return handle ? destroy_common(*handle) : kErrArg; }
int64_t destroy_mumps_cmplx_(const int64_t* handle) { return handle ? destroy_common(*handle) : kErrArg; }

}  // extern "C"
