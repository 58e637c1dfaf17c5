// vboc_cuda.cu -- kernels and the C-ABI of libvboc_b200.so (see include/vboc_b200.h).
//
// solve_kernel: persistent CTAs of WARPS_PER_CTA independent warps; every warp pulls the next
// problem index from a global counter (iteration counts vary 100x between problems, so static
// assignment would leave most of the machine idle at the tail) and runs the whole SQP solve of
// that problem (ocp_warp.h) on its private workspace slot in HBM / L2 and its private block of
// shared memory.  No inter-warp synchronisation anywhere.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vboc_b200.h"
#include "mlp_forward.cuh"
#include "mlp_tc.cuh"
#include "mlp_pipe.cuh"
#include "pool_select.cuh"
#include "ocp_lane.h"
#include "ocp_warp.h"
#include "datagen_warp.h"

using namespace vboc;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_OK(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(VBOC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    } while (0)

#ifndef VB_WARPS_PER_CTA
#define VB_WARPS_PER_CTA 4
#endif
// largest horizon: both workspaces (warp and lane kernels) hold N_max + 1 stages
constexpr int VBOC_N_MAX = 128;
static_assert(VBOC_N_MAX + 1 <= (int)Work<3>::SMAX && VBOC_N_MAX + 1 <= (int)LaneLayout<3>::SMAX, "workspace stages");
constexpr int WARPS_PER_CTA = VB_WARPS_PER_CTA;  // tuning builds may change the CTA shape
// The batch solve kernel runs the same 20 warps per SM as 2-warp CTAs (10 per SM): a persistent CTA frees its registers
// and shared memory only when ALL its warps have run out of work, and a warp that holds a problem of several hundred SQP
// iterations keeps its CTA's idle warps' slots from the next launch for seconds.  Measured on the bench (4 steps on 4
// streams): 4 / 2 / 1 warps per CTA = 3.51 / 3.59 / 3.42 k OCP/s (steady state 3.61 / 3.61 / 3.40 M IPM iterations/s).
#ifndef VB_SOLVE_WARPS
#define VB_SOLVE_WARPS (VB_WARPS_PER_CTA < 2 ? VB_WARPS_PER_CTA : 2)
#endif
constexpr int SOLVE_WARPS = VB_SOLVE_WARPS;
static_assert(WARPS_PER_CTA % SOLVE_WARPS == 0, "CTA shapes");
#ifdef VB_TUNE_MINB
#define VB_LB_MINB(m) VB_TUNE_MINB
#else
#define VB_LB_MINB(m) m
#endif

// device-resident problem batch, reference-shaped
struct Batch {
    int batch, Nmax, nxr;
    const int *N;
    const double *xg, *ug, *p, *lbx0, *ubx0, *lbx, *ubx, *lbxN, *ubxN, *lbu, *ubu, *dir, *h;
    double *x, *u;
    vboc_stats *st;
    double *pi_out = nullptr, *lam_out = nullptr;  // optional multiplier export (vboc_download_multipliers)
    // MPC family: tracking weights (shared by the batch), per-problem references, the margin network, the
    // bounds of the terminal constraint and the export of its multipliers
    const double *Wz = nullptr, *WzN = nullptr, *yref = nullptr, *yrefN = nullptr;
    NnNet nn;
    double lh = 0.0, uh = 0.0;
    double *lamg_out = nullptr;
    int rows_soft = 0;               // vboc_set_mpc_rows: the margin row at every stage, softened
    const double *rowZ = nullptr;    // [batch][Nmax + 1][4] penalties (Zl, Zu, zl, zu)
    double *rowm_out = nullptr;      // [batch][Nmax + 1][6] row multipliers and slacks
    double cart_xc = 0.0, cart_yc = 0.0;  // vboc_set_cartesian (VBOC family, n = 2): circle centre; lh / uh above
    // AL family: guess network evaluated in the kernel (vboc_set_guess_network) and the export of the computed guesses
    const GuessNet *gnn = nullptr;
    double *xg_out = nullptr;
    double *work;  // slots * work_doubles
    size_t work_doubles;
    unsigned int *counter;
    int mode;
    vboc_opts opts;
    // streaming launches (vboc_stream_*): the i-th problem of the launch lives in slot index[i] of the
    // arrays above; done[slot] is raised (system scope) when its results are written; the CTA takes its
    // workspace from a per-SM pool (ws_mask[sm]: one bit per workspace of that SM) because any number of
    // launches may be resident at once.
    const int *index;
    int *done;
    unsigned int *ws_mask;
    int ws_per_sm;
};

template <int NQ, int FAM, int MINB, bool STREAM = false>
__global__ void __launch_bounds__((STREAM ? WARPS_PER_CTA : SOLVE_WARPS) * 32,
                                  VB_LB_MINB(MINB) * (STREAM ? 1 : WARPS_PER_CTA / SOLVE_WARPS)) solve_kernel(const Batch B) {
    constexpr int W = STREAM ? WARPS_PER_CTA : SOLVE_WARPS;  // warps of this CTA (MINB counts 4-warp CTAs' worth of registers)
    __shared__ Smem<NQ> smem[W];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int slot = blockIdx.x * W + warp;
    unsigned int *ws_word = nullptr;
    unsigned int ws_bit = 0;
    if constexpr (STREAM) {
        // Workspaces belong to RESIDENT CTAs, not to launches: at most MINB CTAs of this kernel fit on an SM,
        // so a pool of ws_per_sm >= MINB workspaces per SM serves every launch that is in flight.
        __shared__ int cta_ws;
        if (threadIdx.x == 0) {
            unsigned int sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            ws_word = B.ws_mask + sm;
            const unsigned int all = (1u << B.ws_per_sm) - 1u;
            for (;;) {
                unsigned int cur = *(volatile unsigned int *)ws_word, fr = ~cur & all;
                if (fr) {
                    ws_bit = fr & (0u - fr);
                    if (atomicCAS(ws_word, cur, cur | ws_bit) == cur) break;
                } else {
                    __nanosleep(500);  // a CTA that migrated here after a preemption holds a bit: wait for a release
                }
            }
            cta_ws = (int)sm * B.ws_per_sm + (__ffs(ws_bit) - 1);
        }
        __syncthreads();
        slot = cta_ws * WARPS_PER_CTA + warp;
    }
    Work<NQ> w;
    w.carve(B.work + (size_t)slot * B.work_doubles, B.Nmax);
    WarpSolver<NQ, FAM> sol(smem[warp], w, B.opts);
    const int nu = NQ;
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(B.counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)B.batch) break;
        if constexpr (STREAM) b = (unsigned)B.index[b];
        Prob pb;
        pb.N = B.N[b], pb.nxr = B.nxr, pb.h = B.h[b];
        pb.p = B.p ? B.p + (size_t)b * (NQ + 1) : nullptr;
        pb.wt = B.p ? pb.p[NQ] : 0.0;
        pb.xg = B.xg + (size_t)b * (B.Nmax + 1) * B.nxr, pb.ug = B.ug + (size_t)b * B.Nmax * nu;
        pb.lbx0 = B.lbx0 + (size_t)b * B.nxr, pb.ubx0 = B.ubx0 + (size_t)b * B.nxr;
        pb.lbx = B.lbx + (size_t)b * B.nxr, pb.ubx = B.ubx + (size_t)b * B.nxr;
        pb.lbxN = B.lbxN + (size_t)b * B.nxr, pb.ubxN = B.ubxN + (size_t)b * B.nxr;
        pb.lbu = B.lbu + (size_t)b * nu, pb.ubu = B.ubu + (size_t)b * nu;
        pb.dir = B.dir ? B.dir + (size_t)b * NQ : nullptr;
        if constexpr (STREAM) pb.dir = (B.dir && B.dir[(size_t)b * NQ] == B.dir[(size_t)b * NQ]) ? pb.dir : nullptr;
        pb.x = B.x + (size_t)b * (B.Nmax + 1) * B.nxr, pb.u = B.u + (size_t)b * B.Nmax * nu;
        pb.st = B.st + b;
        if (B.pi_out) {
            pb.pi_out = B.pi_out + (size_t)b * B.Nmax * 2 * NQ;
            pb.lam_out = B.lam_out + (size_t)b * (B.Nmax + 1) * 6 * NQ;
        }
        if constexpr (FAM == VBOC_FAMILY_AL) {
            pb.gnn = B.gnn;
            pb.xg_out = B.xg_out ? B.xg_out + (size_t)b * (B.Nmax + 1) * B.nxr : nullptr;
        }
        sol.solve(pb, B.mode);
        __syncwarp();
        if constexpr (STREAM) {
            // results (mapped host memory) before the flag, at system scope: the host polls done[]
            __threadfence_system();
            __syncwarp();
            if (lane == 0) *(volatile int *)(B.done + b) = 1;
        }
    }
    if constexpr (STREAM) {
        __syncthreads();
        if (threadIdx.x == 0) atomicAnd(ws_word, ~ws_bit);
    }
}

// solve_rows_kernel: solve_kernel for the families that carry general rows (RowF records behind the workspace, the extra
// SmemMpc block per warp): FAM = VBOC_FAMILY_MPC (tracking cost + the learned margin as a hard terminal
// row or as soft rows at every stage) and FAM = VBOC_FAMILY_CART (the VBOC OCP + the Cartesian path constraint).
template <int NQ, int FAM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4) solve_rows_kernel(const Batch B) {
    __shared__ Smem<NQ> smem[WARPS_PER_CTA];
    __shared__ SmemMpc<NQ> mpc[WARPS_PER_CTA];
    __shared__ NnNet net;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WARPS_PER_CTA + warp;
    if constexpr (FAM == VBOC_FAMILY_MPC) {
        if (threadIdx.x == 0) net = B.nn;
        __syncthreads();
    }
    Work<NQ> w;
    w.carve(B.work + (size_t)slot * B.work_doubles, B.Nmax);
    WarpSolver<NQ, FAM> sol(smem[warp], w, B.opts, &mpc[warp]);
    const int nu = NQ, nz = 3 * NQ;
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(B.counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)B.batch) break;
        Prob pb;
        pb.N = B.N[b], pb.nxr = B.nxr, pb.h = B.h[b], pb.p = nullptr, pb.wt = 0.0, pb.dir = nullptr;
        pb.xg = B.xg + (size_t)b * (B.Nmax + 1) * B.nxr, pb.ug = B.ug + (size_t)b * B.Nmax * nu;
        pb.lbx0 = B.lbx0 + (size_t)b * B.nxr, pb.ubx0 = B.ubx0 + (size_t)b * B.nxr;
        pb.lbx = B.lbx + (size_t)b * B.nxr, pb.ubx = B.ubx + (size_t)b * B.nxr;
        pb.lbxN = B.lbxN + (size_t)b * B.nxr, pb.ubxN = B.ubxN + (size_t)b * B.nxr;
        pb.lbu = B.lbu + (size_t)b * nu, pb.ubu = B.ubu + (size_t)b * nu;
        pb.x = B.x + (size_t)b * (B.Nmax + 1) * B.nxr, pb.u = B.u + (size_t)b * B.Nmax * nu;
        pb.st = B.st + b;
        pb.lh = B.lh, pb.uh = B.uh;
        if constexpr (FAM == VBOC_FAMILY_MPC) {
            pb.Wz = B.Wz, pb.WzN = B.WzN, pb.yref = B.yref + (size_t)b * nz, pb.yrefN = B.yrefN + (size_t)b * B.nxr;
            pb.nn = &net;
            pb.lamg_out = B.lamg_out ? B.lamg_out + 2 * (size_t)b : nullptr;
            pb.rows_soft = B.rows_soft;
            pb.rowZ = B.rowZ ? B.rowZ + (size_t)b * (B.Nmax + 1) * 4 : nullptr;
        } else {
            pb.p = B.p + (size_t)b * (NQ + 1), pb.wt = pb.p[NQ];
            pb.dir = B.dir ? B.dir + (size_t)b * NQ : nullptr;
            pb.cart_xc = B.cart_xc, pb.cart_yc = B.cart_yc;
        }
        pb.rowm_out = B.rowm_out ? B.rowm_out + (size_t)b * (B.Nmax + 1) * 6 : nullptr;
        if (B.pi_out) {
            pb.pi_out = B.pi_out + (size_t)b * B.Nmax * 2 * NQ;
            pb.lam_out = B.lam_out + (size_t)b * (B.Nmax + 1) * 6 * NQ;
        }
        sol.solve(pb, B.mode);
        __syncwarp();
    }
}

// solve_lane_kernel: one LANE per OCP (ocp_lane.h).  The 32 lanes of a warp advance 32 problems in
// lockstep, one SQP iteration per trip of the loop; a lane whose problem finished pulls the next index
// at the top of the loop, so the warp stays full until the queue is empty.
constexpr int LANE_THREADS = 128;
template <int NQ, int FAM, int DTS>
__global__ void __launch_bounds__(LANE_THREADS, 2) solve_lane_kernel(const Batch B, double *work) {
    // `work` is a direct pointer parameter (not a member of B) so that the compiler knows it is global memory
    // and emits LDG/STG instead of generic loads / stores
    const int lane = threadIdx.x & 31;
    const size_t wslot = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double *base = work + wslot * LaneLayout<NQ, DTS>::TOTAL * 32;
    // per-thread scratch [element][thread] (conflict free): the symmetric stage Hessian and its gradient
    extern __shared__ double lane_scratch[];  // SM_TOTAL * LANE_THREADS doubles (dynamic: > 48 KB for n = 3)
    double *scratch = lane_scratch + threadIdx.x;
    // address-space facts for the optimiser: typed loads/stores (LDG/STG, LDS/STS) instead of generic ones,
    // and no aliasing between the workspace and the scratch
    LaneSolver<NQ, FAM, 32, DTS> sol(base, scratch, lane, B.opts);
    LaneState ls;
    ls.have = 0, ls.it = 0;
    Prob pb;
    bool exhausted = false;
    const int nu = NQ;
    for (;;) {
        if (!ls.have && !exhausted) {
            unsigned int b = atomicAdd(B.counter, 1u);
            if (b < (unsigned)B.batch) {
                pb.N = B.N[b], pb.nxr = B.nxr, pb.h = B.h[b];
                pb.p = B.p ? B.p + (size_t)b * (NQ + 1) : nullptr;
                pb.wt = B.p ? pb.p[NQ] : 0.0;
                pb.xg = B.xg + (size_t)b * (B.Nmax + 1) * B.nxr, pb.ug = B.ug + (size_t)b * B.Nmax * nu;
                pb.lbx0 = B.lbx0 + (size_t)b * B.nxr, pb.ubx0 = B.ubx0 + (size_t)b * B.nxr;
                pb.lbx = B.lbx + (size_t)b * B.nxr, pb.ubx = B.ubx + (size_t)b * B.nxr;
                pb.lbxN = B.lbxN + (size_t)b * B.nxr, pb.ubxN = B.ubxN + (size_t)b * B.nxr;
                pb.lbu = B.lbu + (size_t)b * nu, pb.ubu = B.ubu + (size_t)b * nu;
                pb.dir = B.dir ? B.dir + (size_t)b * NQ : nullptr;
                pb.x = B.x + (size_t)b * (B.Nmax + 1) * B.nxr, pb.u = B.u + (size_t)b * B.Nmax * nu;
                pb.st = B.st + b;
                sol.begin(ls, pb);
            } else {
                exhausted = true;
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, ls.have)) break;
        if (ls.have && sol.sqp_iteration(ls, B.mode)) sol.finish(ls, pb);
        __syncwarp();
    }
}

// one RK4 step of the unscaled model per thread (the reference's AcadosSimSolver)
template <int NQ>
__global__ void sim_kernel(int batch, const double *x, const double *u, double T, double *xn) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double xi[2 * NQ], ui[NQ], xo[2 * NQ];
#pragma unroll
    for (int i = 0; i < 2 * NQ; ++i) xi[i] = x[(size_t)b * 2 * NQ + i];
#pragma unroll
    for (int i = 0; i < NQ; ++i) ui[i] = u[(size_t)b * NQ + i];
    rk4_step<NQ, double>(xi, ui, T, xo);
#pragma unroll
    for (int i = 0; i < 2 * NQ; ++i) xn[(size_t)b * 2 * NQ + i] = xo[i];
}

// FP64 FMA peak probe: 8 independent DFMA chains per thread, enough warps to fill every SM
__global__ void dfma_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c), a1 = fma(a1, m, c), a2 = fma(a2, m, c), a3 = fma(a3, m, c);
        a4 = fma(a4, m, c), a5 = fma(a5, m, c), a6 = fma(a6, m, c), a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// number of SM identifiers (%smid < %nsmid; may exceed the SM count of the device)
__global__ void nsmid_kernel(unsigned int *out) {
    unsigned int v;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
    *out = v;
}

// datagen_kernel: one warp = one data_generation(v) problem, start to finish (datagen_warp.h).  Persistent CTAs,
// problems handed out by a global counter like the solves of solve_kernel.
template <int NQ>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 5) datagen_kernel(const DgIO<NQ> io, int count, const DgParams P,
                                                                      const vboc_opts opts, double *work,
                                                                      size_t work_doubles, unsigned int *counter,
                                                                      unsigned long long *t_start) {
    __shared__ Smem<NQ> smem[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WARPS_PER_CTA + warp;
    // kernel start on the device clock: the first warp to arrive sets it
    unsigned long long t0 = 0;
    if (lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const unsigned long long prev = atomicCAS(t_start, 0ull, now);
        t0 = prev ? prev : now;
    }
    double *base = work + (size_t)slot * work_doubles;
    Work<NQ> w;
    w.carve(base, DG_N_CAP);
    DgWork<NQ> g;
    g.carve(base + Work<NQ>::TOTAL);
    WarpSolver<NQ, VBOC_FAMILY_VBOC> sol(smem[warp], w, opts);
    DataGen<NQ> dg(sol, g, P);
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)count) break;
        dg.run(io, (int)b);
        __syncwarp();
        if (lane == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            dg.c.t_done_us = (int)((now - t0) / 1000ull);
            io.cnt[b] = dg.c;
        }
        __syncwarp();
    }
}

// testing_kernel: datagen_kernel for the test-data worker `testing(v)` (DataGen::run_testing)
template <int NQ>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 5) testing_kernel(const TestIO<NQ> io, int count, const DgParams P,
                                                                      const vboc_opts opts, double *work,
                                                                      size_t work_doubles, unsigned int *counter,
                                                                      unsigned long long *t_start) {
    __shared__ Smem<NQ> smem[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * WARPS_PER_CTA + warp;
    unsigned long long t0 = 0;
    if (lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        const unsigned long long prev = atomicCAS(t_start, 0ull, now);
        t0 = prev ? prev : now;
    }
    double *base = work + (size_t)slot * work_doubles;
    Work<NQ> w;
    w.carve(base, DG_N_CAP);
    DgWork<NQ> g;
    g.carve(base + Work<NQ>::TOTAL);
    WarpSolver<NQ, VBOC_FAMILY_VBOC> sol(smem[warp], w, opts);
    DataGen<NQ> dg(sol, g, P);
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)count) break;
        dg.run_testing(io, (int)b);
        __syncwarp();
        if (lane == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            dg.c.t_done_us = (int)((now - t0) / 1000ull);
            io.cnt[b] = dg.c;
        }
        __syncwarp();
    }
}

// gathers the saved rows of all problems into one contiguous array (problem order)
__global__ void dg_compact_kernel(const double *rows, const DgCounters *cnt, const long long *off, double *out, int nx) {
    const int b = blockIdx.x;
    const size_t n = (size_t)cnt[b].n_rows * nx;
    const double *src = rows + (size_t)b * DG_ROWS_MAX * nx;
    double *dst = out + (size_t)off[b] * nx;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

struct vboc_solver {
    int n, family, cap, Nmax, device, nxr, nu;
    int slots, grid, ctas_per_sm, lane_kernel;
    int free_dt, grid_free_dt;  // 1-DOF VBOC with a free dt state: lane kernel with the dt state kept
    double *dwork_free_dt;
    cudaStream_t stream;
    vboc_opts opts;
    // device buffers
    int *dN;
    double *dxg, *dug, *dp, *dlbx0, *dubx0, *dlbx, *dubx, *dlbxN, *dubxN, *dlbu, *dubu, *ddir, *dh;
    double *dx, *du, *dwork;
    double *dpi, *dlam;  // multiplier export, allocated by vboc_export_multipliers
    // MPC family (vboc_set_mpc / vboc_set_mpc_reference)
    double *dnn, *dWz, *dWzN, *dyref, *dyrefN, *dlamg;
    double *drowZ, *drowm;  // soft rows (vboc_set_mpc_rows)
    int rows_soft, rows_batch;
    int cart_on;            // vboc_set_cartesian (VBOC family, n = 2)
    double cart_xc, cart_yc;
    NnNet nn;
    double mpc_lh, mpc_uh;
    int mpc_set, mpc_ref_batch;
    // AL family: guess network (vboc_set_guess_network)
    double *dgn, *dxg_out;
    GuessNet *dgnn;   // device copy of the descriptor
    int gn_on, gn_out;
    vboc_stats *dst;
    unsigned int *dcounter;
    size_t work_doubles;
    cudaEvent_t ev0, ev1;
    int batch;  // resident batch, 0 if none
    int has_dir, has_p;
    double last_ms;
    // pinned staging for the host <-> device copies
    char *stage;
    size_t stage_bytes;
};

// VB_TUNE_BUILD: kernel-tuning builds (tools/build_variant.sh) compile only the benchmark's instantiation
// (3-DOF VBOC, 5 CTAs / SM) -- seconds instead of minutes; everything else is refused at run time.
#ifdef VB_TUNE_BUILD
#define VB_ALL_SYSTEMS(GO) GO(3, 0)
#else
#define VB_ALL_SYSTEMS(GO) GO(1, 0) GO(2, 0) GO(3, 0) GO(1, 1) GO(2, 1) GO(3, 1)
#endif

template <int NQ, int FAM>
static cudaError_t launch(vboc_solver *s, const Batch &B) {
#ifndef VB_TUNE_BUILD
    if (s->lane_kernel) {
        const int smem = LaneSolver<NQ, FAM, 32, 0>::SM_TOTAL * LANE_THREADS * (int)sizeof(double);
        cudaFuncSetAttribute(solve_lane_kernel<NQ, FAM, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        solve_lane_kernel<NQ, FAM, 0><<<s->grid, LANE_THREADS, smem, s->stream>>>(B, B.work);
        return cudaGetLastError();
    }
    if (s->ctas_per_sm >= 6)
        solve_kernel<NQ, FAM, 6><<<s->grid, SOLVE_WARPS * 32, 0, s->stream>>>(B);
    else if (s->ctas_per_sm == 4)
        solve_kernel<NQ, FAM, 4><<<s->grid, SOLVE_WARPS * 32, 0, s->stream>>>(B);
    else
#endif
    {
#ifdef VB_TUNE_CARVEOUT  // tuning builds: shared-memory carve-out preference in percent of the SM's 228 KB
        cudaFuncSetAttribute(solve_kernel<NQ, FAM, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, VB_TUNE_CARVEOUT);
#endif
        solve_kernel<NQ, FAM, 5><<<s->grid, SOLVE_WARPS * 32, 0, s->stream>>>(B);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Streaming engine: problems enter and leave one at a time (tickets), launches never wait for each other.
// The drivers' per-problem state machines (data_generation / testing: chains of dependent solves whose
// lengths differ 100x between problems) keep the GPU full this way; with whole-batch calls every round
// waits for its slowest solve.
//
//  * problem slots (inputs, results, done flags) live in MAPPED PINNED host memory: the kernel reads a
//    problem's ~9 KB once and writes its ~8 KB result over the host link, the host polls done[slot];
//  * every submit is one launch of solve_kernel<.., STREAM> on a non-blocking stream of its own;
//  * workspaces are bound to resident CTAs (per-SM pool), so any number of launches can be in flight.
struct vboc_stream {
    int n, family, cap, Nmax, device, nxr, nu, num_sms, nsmid;
    vboc_opts opts;
    // mapped pinned slot arrays
    int *N, *done;
    double *xg, *ug, *p, *lbx0, *ubx0, *lbx, *ubx, *lbxN, *ubxN, *lbu, *ubu, *dir, *h, *x, *u;
    vboc_stats *st;
    // device
    double *dwork;
    size_t work_doubles;
    unsigned int *ws_mask, *dcounters;
    int ws_per_sm;
    // launch records
    static constexpr int NREC = 96;
    struct Rec {
        cudaStream_t stream;
        cudaEvent_t ev;
        int *index;  // mapped pinned, cap entries
        bool busy;
    } rec[NREC];
    std::vector<int> free_slots, inflight;
    // batched RK4 steps on their own stream
    cudaStream_t sim_stream;
    double *sx, *su, *sxn;  // mapped pinned, sim_cap rows
    int sim_cap;
    long long launches, solved, idle_polls;
};

template <int NQ, int FAM>
static cudaError_t launch_stream(int grid, cudaStream_t st, const Batch &B) {
    solve_kernel<NQ, FAM, 5, true><<<grid, WARPS_PER_CTA * 32, 0, st>>>(B);
    return cudaGetLastError();
}

extern "C" {

const char *vboc_last_error(void) { return g_err.c_str(); }
const char *vboc_version(void) { return "vboc_b200 0.1 (sm_100a)"; }

void vboc_default_opts(int family, vboc_opts *o) {
    memset(o, 0, sizeof(*o));
    // acados / HPIPM (BALANCE) defaults
    o->tol_eq = o->tol_ineq = o->tol_comp = 1e-6;
    o->alpha_min = 0.05, o->alpha_reduction = 0.7;
    o->qp_tol_stat = 1e-6, o->qp_tol_eq = o->qp_tol_ineq = o->qp_tol_comp = 1e-8;
    o->qp_mu0 = 1e1, o->qp_alpha_min = 1e-12, o->qp_reg_prim = 1e-15;
    o->qp_lam_min = 1e-16, o->qp_t_min = 1e-16, o->qp_tau_min = 1e-16;
    if (family == VBOC_FAMILY_VBOC) {
        // VBOC/triplependulum_class_vboc.py:129-141
        o->tol_stat = 1e-3, o->qp_tol_stat = 1e-3;
        o->qp_iter_max = 100, o->max_iter = 1000;
        o->globalization = 1, o->alpha_reduction = 0.3, o->alpha_min = 1e-2;
        o->levenberg_marquardt = 1e-5;
    } else if (family == VBOC_FAMILY_MPC) {
        // VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:183-190: `tol = 1e-2` sets the four
        // NLP tolerances and the QP solver's [acados_template: the `tol` setter], iter limits 1000 / 100, merit
        // back-tracking 0.3 / 1e-2, Levenberg-Marquardt 1
        o->tol_stat = o->tol_eq = o->tol_ineq = o->tol_comp = 1e-2;
        o->qp_tol_stat = o->qp_tol_eq = o->qp_tol_ineq = o->qp_tol_comp = 1e-2;
        o->qp_iter_max = 100, o->max_iter = 1000;
        o->globalization = 1, o->alpha_reduction = 0.3, o->alpha_min = 1e-2;
        o->levenberg_marquardt = 1.0;
    } else {
        // AL classes: SQP_RTI with acados defaults (AL/pendulum_class_al.py:124)
        o->tol_stat = 1e-6;
        o->qp_iter_max = 50, o->max_iter = 100;
        o->globalization = 0, o->levenberg_marquardt = 0.0;
    }
}

static int solver_create_impl(vboc_solver *s, int n_dof, int family, int batch_capacity, int N_max, int device);
static int stream_create_impl(vboc_stream *s, int n_dof, int family, int capacity, int N_max, int device);

static size_t work_doubles_for(int n, int Nmax) {
    return n == 1 ? Work<1>::doubles(Nmax) : (n == 2 ? Work<2>::doubles(Nmax) : Work<3>::doubles(Nmax));
}

int vboc_create(int n_dof, int family, int batch_capacity, int N_max, int device, vboc_solver **out) {
    if (!out || n_dof < 1 || n_dof > 3 ||
        (family != VBOC_FAMILY_VBOC && family != VBOC_FAMILY_AL && family != VBOC_FAMILY_MPC) || batch_capacity < 1 ||
        N_max < 1 || N_max > VBOC_N_MAX || (family == VBOC_FAMILY_MPC && n_dof < 2))
        return fail(VBOC_ERR_ARG, "vboc_create: bad argument");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VBOC_ERR_CUDA, "vboc_create: no such CUDA device");
    CUDA_OK(cudaSetDevice(device));
    vboc_solver *s = new vboc_solver();
    memset(s, 0, sizeof(*s));
    // single cleanup path: a failure below frees whatever the half-built handle already owns
    int rc = solver_create_impl(s, n_dof, family, batch_capacity, N_max, device);
    if (rc) {
        const std::string keep = g_err;
        vboc_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return 0;
}

static int solver_create_impl(vboc_solver *s, int n_dof, int family, int batch_capacity, int N_max, int device) {
    s->n = n_dof, s->family = family, s->cap = batch_capacity, s->Nmax = N_max, s->device = device;
    s->nxr = 2 * n_dof + (family == VBOC_FAMILY_VBOC), s->nu = n_dof;
    s->last_ms = -1.0;
    vboc_default_opts(family, &s->opts);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    // resident CTAs per SM: 4 (128 registers / thread), 5 (96, measured best: 2.40 M vs 2.25 M IPM iterations/s)
    // or 6 (80, spills); VBOC_CTAS_PER_SM overrides for tuning
    int ctas_per_sm = 5;
    if (const char *e = getenv("VBOC_CTAS_PER_SM")) ctas_per_sm = atoi(e) >= 6 ? 6 : (atoi(e) == 5 ? 5 : 4);
#ifdef VB_TUNE_CTAS
    ctas_per_sm = VB_TUNE_CTAS;
#endif
    s->ctas_per_sm = ctas_per_sm;
    // warp slots resident at once (ctas_per_sm counts 4-warp CTAs), a multiple of WARPS_PER_CTA; the batch solve kernel
    // runs them as CTAs of SOLVE_WARPS warps, the row-family kernels as CTAs of WARPS_PER_CTA
    int max_slots = prop.multiProcessorCount * ctas_per_sm * WARPS_PER_CTA;
    int need = (batch_capacity + WARPS_PER_CTA - 1) / WARPS_PER_CTA * WARPS_PER_CTA;
    s->slots = need < max_slots ? need : max_slots;
    s->grid = s->slots / SOLVE_WARPS;
    s->work_doubles = work_doubles_for(n_dof, N_max);
    if (family == VBOC_FAMILY_MPC)  // + the row records of the margin constraint
        s->work_doubles = n_dof == 2 ? Work<2>::doubles_rows(N_max) : Work<3>::doubles_rows(N_max);
    // kernel mapping: VBOC_KERNEL=lane selects one lane per OCP (ocp_lane.h), default one warp per OCP
    s->lane_kernel = 0;
    if (const char *e = getenv("VBOC_KERNEL")) s->lane_kernel = strcmp(e, "lane") == 0 && family != VBOC_FAMILY_MPC;
    if (s->lane_kernel) {
        size_t per_lane = n_dof == 1 ? LaneLayout<1>::TOTAL : (n_dof == 2 ? LaneLayout<2>::TOTAL : LaneLayout<3>::TOTAL);
        int warps_per_cta = LANE_THREADS / 32;
        int max_g = prop.multiProcessorCount * 2;
        int need_g = (batch_capacity + LANE_THREADS - 1) / LANE_THREADS;
        s->grid = need_g < max_g ? need_g : max_g;
        s->slots = s->grid * warps_per_cta;  // warp slots
        s->work_doubles = per_lane * 32;     // per warp slot
    }
    size_t B = (size_t)batch_capacity, nxr = s->nxr, nu = s->nu;
#define DEV_ALLOC(ptr, count) CUDA_OK(cudaMalloc((void **)&s->ptr, (size_t)(count) * sizeof(*s->ptr)))
    DEV_ALLOC(dN, B);
    DEV_ALLOC(dxg, B * (N_max + 1) * nxr);
    DEV_ALLOC(dug, B * N_max * nu);
    DEV_ALLOC(dp, B * (n_dof + 1));
    DEV_ALLOC(dlbx0, B * nxr);
    DEV_ALLOC(dubx0, B * nxr);
    DEV_ALLOC(dlbx, B * nxr);
    DEV_ALLOC(dubx, B * nxr);
    DEV_ALLOC(dlbxN, B * nxr);
    DEV_ALLOC(dubxN, B * nxr);
    DEV_ALLOC(dlbu, B * nu);
    DEV_ALLOC(dubu, B * nu);
    DEV_ALLOC(ddir, B * n_dof);
    DEV_ALLOC(dh, B);
    DEV_ALLOC(dx, B * (N_max + 1) * nxr);
    DEV_ALLOC(du, B * N_max * nu);
    DEV_ALLOC(dst, B);
    DEV_ALLOC(dcounter, 1);
    DEV_ALLOC(dwork, (size_t)s->slots * s->work_doubles);
#undef DEV_ALLOC
    CUDA_OK(cudaEventCreate(&s->ev0));
    CUDA_OK(cudaEventCreate(&s->ev1));
    s->stage_bytes = B * (N_max + 1) * nxr * sizeof(double);
    CUDA_OK(cudaMallocHost((void **)&s->stage, s->stage_bytes));
    return 0;
}

void vboc_destroy(vboc_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    void *ptrs[] = {s->dN,    s->dxg,   s->dug,  s->dp,   s->dlbx0, s->dubx0, s->dlbx,
                    s->dubx,  s->dlbxN, s->dubxN, s->dlbu, s->dubu,  s->ddir,  s->dh,
                    s->dx,    s->du,    s->dst,  s->dcounter, s->dwork, s->dpi, s->dlam,
                    s->dnn,   s->dWz,   s->dWzN, s->dyref, s->dyrefN, s->dlamg, s->drowZ, s->drowm, s->dgn,  s->dxg_out, s->dgnn};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (s->dwork_free_dt) cudaFree(s->dwork_free_dt);
    if (s->stage) cudaFreeHost(s->stage);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

int vboc_set_opts(vboc_solver *s, const vboc_opts *o) {
    if (!s || !o) return fail(VBOC_ERR_ARG, "vboc_set_opts: null argument");
    s->opts = *o;
    return 0;
}

int vboc_set_stream(vboc_solver *s, void *cuda_stream) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_set_stream: null handle");
    s->stream = (cudaStream_t)cuda_stream;
    return 0;
}

// host -> device through the pinned staging buffer (host arrays may be pageable)
static int h2d(vboc_solver *s, void *dst, const void *src, size_t bytes) {
    const char *p = (const char *)src;
    char *d = (char *)dst;
    while (bytes) {
        size_t c = bytes < s->stage_bytes ? bytes : s->stage_bytes;
        memcpy(s->stage, p, c);
        CUDA_OK(cudaMemcpyAsync(d, s->stage, c, cudaMemcpyHostToDevice, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        p += c, d += c, bytes -= c;
    }
    return 0;
}
static int d2h(vboc_solver *s, void *dst, const void *src, size_t bytes) {
    char *p = (char *)dst;
    const char *d = (const char *)src;
    while (bytes) {
        size_t c = bytes < s->stage_bytes ? bytes : s->stage_bytes;
        CUDA_OK(cudaMemcpyAsync(s->stage, d, c, cudaMemcpyDeviceToHost, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        memcpy(p, s->stage, c);
        p += c, d += c, bytes -= c;
    }
    return 0;
}

// Validation and host-side preparation of one problem (what the shim's OCP_solve would otherwise do stage by
// stage through ocp_solver.set / constraints_set): the RK4 step h, the unit direction d of the stage-0
// projector, and the refusals listed in include/vboc_b200.h.  *pinned = false reports a free dt state
// (VBOC family; the caller decides whether that is supported), h is then 1.
static int prepare_problem(int n, int nxr, int Nmax, int family, int Nb, const double *l0, const double *u0,
                           const double *l1, const double *u1, const double *lN, const double *uN,
                           const double *xg, const double *C, double Tf, double *h, double *dir, bool *pinned_out) {
    if (Nb < 1 || Nb > Nmax) return fail(VBOC_ERR_ARG, "vboc_upload: horizon out of range");
    *pinned_out = true;
    if (family == VBOC_FAMILY_VBOC) {
        // the dt state must be pinned (VBOC/triplependulum_vboc.py:98-103)
        double dt = l0[2 * n];
        bool pinned = u0[2 * n] == dt && l1[2 * n] == dt && u1[2 * n] == dt && lN[2 * n] == dt && uN[2 * n] == dt;
        for (int k = 0; k <= Nb && pinned; ++k) pinned = xg[(size_t)k * nxr + 2 * n] == dt;
        if (!pinned) {
            *pinned_out = false;
            *h = 1.0;
            return 0;
        }
        if (!(dt > 0.0)) return fail(VBOC_ERR_UNSUPPORTED, "vboc_upload: dt <= 0");
        *h = dt;
    } else {
        *h = Tf / Nb;
    }
    int nf = 0, nfv = 0;
    for (int i = 0; i < 2 * n; ++i)
        if (lN[i] == uN[i]) {
            ++nf;
            nfv += i >= n;
        }
    if (!(nf == 0 || (nf == n && nfv == n)))
        return fail(VBOC_ERR_UNSUPPORTED, "vboc_upload: terminal equalities must be on none or exactly all velocities");
    if (C) {
        // only C0 = [0 | I - d d' | 0] (VBOC/triplependulum_class_vboc.py:174-178)
        double Md[3][3], d[3];
        int jm = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) Md[i][j] = (i == j) - C[i * nxr + n + j];
        for (int i = 1; i < n; ++i)
            if (Md[i][i] > Md[jm][jm]) jm = i;
        if (!(Md[jm][jm] > 0.0)) return fail(VBOC_ERR_UNSUPPORTED, "vboc_upload: C0 is not a projector I - d d'");
        double dj = sqrt(Md[jm][jm]), nrm = 0.0;
        for (int i = 0; i < n; ++i) d[i] = Md[i][jm] / dj, nrm += d[i] * d[i];
        nrm = sqrt(nrm);
        for (int i = 0; i < n; ++i) d[i] /= nrm;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < nxr; ++j) {
                double want = (j >= n && j < 2 * n) ? (i == j - n) - d[i] * d[j - n] : 0.0;
                if (fabs(C[i * nxr + j] - want) > 1e-9)
                    return fail(VBOC_ERR_UNSUPPORTED, "vboc_upload: C0 is not of the form [0 | I - d d' | 0]");
            }
        for (int i = 0; i < n; ++i)
            if (l0[n + i] == u0[n + i])
                return fail(VBOC_ERR_UNSUPPORTED,
                            "vboc_upload: fixed initial velocity together with a direction constraint");
        for (int i = 0; i < n; ++i) dir[i] = d[i];
    }
    return 0;
}

int vboc_upload(vboc_solver *s, int batch, const int *N, const double *x_guess,
                const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                const double *lbu, const double *ubu, const double *C0, double Tf) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_upload: null handle");
    if (batch < 1 || batch > s->cap) return fail(VBOC_ERR_ARG, "vboc_upload: batch exceeds capacity");
    if (!N || !x_guess || !u_guess || !lbx0 || !ubx0 || !lbx || !ubx || !lbxN || !ubxN || !lbu || !ubu)
        return fail(VBOC_ERR_ARG, "vboc_upload: null array");
    const int n = s->n, nxr = s->nxr, Nmax = s->Nmax;
    if (s->family == VBOC_FAMILY_VBOC && !p) return fail(VBOC_ERR_ARG, "vboc_upload: p required");
    CUDA_OK(cudaSetDevice(s->device));
    // ---- validation and host-side preparation (what the shim's OCP_solve would otherwise do
    // stage by stage through ocp_solver.set / constraints_set)
    std::vector<double> h(batch), dir;
    bool free_dt = false;
    if (C0) dir.resize((size_t)batch * n);
    for (int b = 0; b < batch; ++b) {
        bool pinned = true;
        int rc = prepare_problem(n, nxr, Nmax, s->family, N[b], lbx0 + (size_t)b * nxr, ubx0 + (size_t)b * nxr,
                                 lbx + (size_t)b * nxr, ubx + (size_t)b * nxr, lbxN + (size_t)b * nxr,
                                 ubxN + (size_t)b * nxr, x_guess + (size_t)b * (Nmax + 1) * nxr,
                                 C0 ? C0 + (size_t)b * n * nxr : nullptr, Tf, &h[b],
                                 C0 ? &dir[(size_t)b * n] : nullptr, &pinned);
        if (rc) return rc;
        if (!pinned) {
            // free dt (VBOC/pendulum_vboc.py:69-70): 1-DOF only, whole batch
            if (n != 1 || (b > 0 && !free_dt))
                return fail(VBOC_ERR_UNSUPPORTED,
                            "vboc_upload: the dt state must be pinned to one positive value at every "
                            "stage and in the guess (a free dt is supported for the 1-DOF model only, "
                            "and not mixed with pinned problems in one batch)");
            free_dt = true;
        } else if (free_dt) {
            return fail(VBOC_ERR_UNSUPPORTED, "vboc_upload: pinned and free dt mixed in one batch");
        }
    }
    if (s->family == VBOC_FAMILY_AL && s->gn_on)
        for (int b = 0; b < batch; ++b)
            if (N[b] * 2 * n != s->gn_out)
                return fail(VBOC_ERR_ARG, "vboc_upload: the guess network predicts N * 2n values for another horizon N");
    size_t B = batch;
    int rc = 0;
#define UP(dst, src, count) \
    if ((rc = h2d(s, s->dst, src, (size_t)(count) * sizeof(*s->dst)))) return rc
    UP(dN, N, B);
    UP(dxg, x_guess, B * (Nmax + 1) * nxr);
    UP(dug, u_guess, B * Nmax * s->nu);
    if (p) UP(dp, p, B * (n + 1));
    UP(dlbx0, lbx0, B * nxr);
    UP(dubx0, ubx0, B * nxr);
    UP(dlbx, lbx, B * nxr);
    UP(dubx, ubx, B * nxr);
    UP(dlbxN, lbxN, B * nxr);
    UP(dubxN, ubxN, B * nxr);
    UP(dlbu, lbu, B * s->nu);
    UP(dubu, ubu, B * s->nu);
    UP(dh, h.data(), B);
    if (C0) UP(ddir, dir.data(), B * n);
#undef UP
    s->batch = batch, s->has_dir = C0 != nullptr, s->has_p = p != nullptr;
    s->free_dt = free_dt;
    if (free_dt && !s->dwork_free_dt) {
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, s->device));
        int max_g = prop.multiProcessorCount * 2, need_g = (s->cap + LANE_THREADS - 1) / LANE_THREADS;
        s->grid_free_dt = need_g < max_g ? need_g : max_g;
        size_t doubles = (size_t)s->grid_free_dt * (LANE_THREADS / 32) * 32 * LaneLayout<1, 1>::TOTAL;
        CUDA_OK(cudaMalloc((void **)&s->dwork_free_dt, doubles * sizeof(double)));
    }
    return 0;
}

int vboc_solve_resident_async(vboc_solver *s, int mode) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_solve_resident: null handle");
    if (!s->batch) return fail(VBOC_ERR_ARG, "vboc_solve_resident: nothing uploaded");
    if (mode != VBOC_MODE_SQP && mode != VBOC_MODE_RTI) return fail(VBOC_ERR_ARG, "bad mode");
    CUDA_OK(cudaSetDevice(s->device));
    Batch B;
    B.batch = s->batch, B.Nmax = s->Nmax, B.nxr = s->nxr;
    B.N = s->dN, B.xg = s->dxg, B.ug = s->dug, B.p = s->has_p ? s->dp : nullptr;
    B.lbx0 = s->dlbx0, B.ubx0 = s->dubx0, B.lbx = s->dlbx, B.ubx = s->dubx;
    B.lbxN = s->dlbxN, B.ubxN = s->dubxN, B.lbu = s->dlbu, B.ubu = s->dubu;
    B.dir = s->has_dir ? s->ddir : nullptr, B.h = s->dh;
    B.x = s->dx, B.u = s->du, B.st = s->dst;
    B.pi_out = s->dpi, B.lam_out = s->dpi ? s->dlam : nullptr;
    if (s->family == VBOC_FAMILY_AL && s->gn_on) B.gnn = s->dgnn, B.xg_out = s->dxg_out;
    B.work = s->dwork, B.work_doubles = s->work_doubles, B.counter = s->dcounter;
    B.mode = mode, B.opts = s->opts;
    CUDA_OK(cudaMemsetAsync(s->dcounter, 0, sizeof(unsigned int), s->stream));
    CUDA_OK(cudaEventRecord(s->ev0, s->stream));
    cudaError_t e = cudaErrorInvalidValue;
    if (s->family == VBOC_FAMILY_MPC) {
#ifndef VB_TUNE_BUILD
        if (!s->mpc_set) return fail(VBOC_ERR_ARG, "vboc_solve_resident: call vboc_set_mpc first (MPC family)");
        if (s->mpc_ref_batch < s->batch)
            return fail(VBOC_ERR_ARG, "vboc_solve_resident: call vboc_set_mpc_reference for this batch (MPC family)");
        B.Wz = s->dWz, B.WzN = s->dWzN, B.yref = s->dyref, B.yrefN = s->dyrefN, B.nn = s->nn;
        B.lh = s->mpc_lh, B.uh = s->mpc_uh, B.lamg_out = s->dlamg;
        if (s->rows_soft && s->rows_batch < s->batch)
            return fail(VBOC_ERR_ARG, "vboc_solve_resident: call vboc_set_mpc_rows for this batch (soft rows are on)");
        B.rows_soft = s->rows_soft, B.rowZ = s->rows_soft ? s->drowZ : nullptr, B.rowm_out = s->drowm;
        const int g4 = (s->batch + WARPS_PER_CTA - 1) / WARPS_PER_CTA, gmax = s->slots / WARPS_PER_CTA;
        const int grid = g4 < gmax ? g4 : gmax;
        if (s->n == 2) solve_rows_kernel<2, VBOC_FAMILY_MPC><<<grid, WARPS_PER_CTA * 32, 0, s->stream>>>(B);
        else solve_rows_kernel<3, VBOC_FAMILY_MPC><<<grid, WARPS_PER_CTA * 32, 0, s->stream>>>(B);
        e = cudaGetLastError();
#endif
    } else if (s->cart_on) {
#ifndef VB_TUNE_BUILD
        if (s->free_dt || s->lane_kernel) return fail(VBOC_ERR_UNSUPPORTED, "Cartesian constraint: warp kernel, pinned dt only");
        B.lh = s->mpc_lh, B.uh = s->mpc_uh, B.cart_xc = s->cart_xc, B.cart_yc = s->cart_yc, B.rowm_out = s->drowm;
        const int g4 = (s->batch + WARPS_PER_CTA - 1) / WARPS_PER_CTA, gmax = s->slots / WARPS_PER_CTA;
        const int grid = g4 < gmax ? g4 : gmax;
        solve_rows_kernel<2, VBOC_FAMILY_CART><<<grid, WARPS_PER_CTA * 32, 0, s->stream>>>(B);
        e = cudaGetLastError();
#endif
    } else if (s->free_dt) {
#ifndef VB_TUNE_BUILD
        B.work = s->dwork_free_dt;
        const int smem = LaneSolver<1, VBOC_FAMILY_VBOC, 32, 1>::SM_TOTAL * LANE_THREADS * (int)sizeof(double);
        solve_lane_kernel<1, VBOC_FAMILY_VBOC, 1><<<s->grid_free_dt, LANE_THREADS, smem, s->stream>>>(B, B.work);
        e = cudaGetLastError();
#endif
    } else {
#define GO(NQ, FAM) \
    if (s->n == NQ && s->family == FAM) e = launch<NQ, FAM>(s, B);
        VB_ALL_SYSTEMS(GO)
#undef GO
    }
    if (e != cudaSuccess) return fail(VBOC_ERR_CUDA, std::string("solve_kernel launch: ") + cudaGetErrorString(e));
    CUDA_OK(cudaEventRecord(s->ev1, s->stream));
    s->last_ms = -1.0;
    return 0;
}

int vboc_sync(vboc_solver *s) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_sync: null handle");
    CUDA_OK(cudaSetDevice(s->device));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    if (s->last_ms < 0.0 && s->batch) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->last_ms = ms;
        cudaGetLastError();
    }
    return 0;
}

int vboc_solve_resident(vboc_solver *s, int mode) {
    int rc = vboc_solve_resident_async(s, mode);
    return rc ? rc : vboc_sync(s);
}

double vboc_last_kernel_ms(vboc_solver *s) { return s ? s->last_ms : -1.0; }

int vboc_download(vboc_solver *s, double *x, double *u, vboc_stats *stats) {
    if (!s || !s->batch) return fail(VBOC_ERR_ARG, "vboc_download: nothing resident");
    CUDA_OK(cudaSetDevice(s->device));
    size_t B = s->batch;
    int rc;
    if (x && (rc = d2h(s, x, s->dx, B * (s->Nmax + 1) * s->nxr * sizeof(double)))) return rc;
    if (u && (rc = d2h(s, u, s->du, B * s->Nmax * s->nu * sizeof(double)))) return rc;
    if (stats && (rc = d2h(s, stats, s->dst, B * sizeof(vboc_stats)))) return rc;
    return 0;
}

int vboc_set_mpc(vboc_solver *s, int hidden, const float *W1, const float *b1, const float *W2, const float *b2,
                 const float *W3, const float *b3, double mean, double stdv, double safety_margin, double lh, double uh,
                 const double *W, const double *W_e) {
    if (!s || s->family != VBOC_FAMILY_MPC) return fail(VBOC_ERR_ARG, "vboc_set_mpc: not an MPC-family solver");
    if (hidden < 1 || hidden > NN_HMAX || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W || !W_e || !(stdv > 0.0))
        return fail(VBOC_ERR_ARG, "vboc_set_mpc: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    const int n = s->n, nx = 2 * n, nz = 3 * n, H = hidden;
    // FP64 copies of the FP32-trained weights (CasADi evaluates the network in double on float-valued parameters);
    // W2 in both layouts so that the forward and the reverse pass read coalesced
    const size_t o_b1 = (size_t)H * nx, o_W2 = o_b1 + H, o_W2T = o_W2 + (size_t)H * H, o_b2 = o_W2T + (size_t)H * H,
                 o_W3 = o_b2 + H, total = o_W3 + H;
    std::vector<double> h(total);
    for (size_t i = 0; i < (size_t)H * nx; ++i) h[i] = W1[i];
    for (int i = 0; i < H; ++i) h[o_b1 + i] = b1[i], h[o_b2 + i] = b2[i], h[o_W3 + i] = W3[i];
    for (int j = 0; j < H; ++j)
        for (int k = 0; k < H; ++k) h[o_W2 + (size_t)j * H + k] = W2[(size_t)j * H + k], h[o_W2T + (size_t)k * H + j] = W2[(size_t)j * H + k];
    if (s->dnn) cudaFree(s->dnn);
    s->dnn = nullptr;
    CUDA_OK(cudaMalloc((void **)&s->dnn, total * sizeof(double)));
    CUDA_OK(cudaMemcpy(s->dnn, h.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    s->nn.n_in = nx, s->nn.hidden = H, s->nn.W1 = s->dnn, s->nn.b1 = s->dnn + o_b1, s->nn.W2 = s->dnn + o_W2;
    s->nn.W2T = s->dnn + o_W2T, s->nn.b2 = s->dnn + o_b2, s->nn.W3 = s->dnn + o_W3, s->nn.b3 = b3[0];
    s->nn.mean = mean, s->nn.stdv = stdv, s->nn.scale = (100.0 - safety_margin) / 100.0;
    s->nn.vstart = s->n;  // vel_norm over the velocities (vboc_set_mpc_velnorm_start changes it)
    s->mpc_lh = lh, s->mpc_uh = uh;
    // weights arrive in acados' y = [x; u] order (cost.W = blkdiag(Q, R)); the engine orders z = [u; x]
    std::vector<double> wz(nz), wzN(nx);
    for (int i = 0; i < nx; ++i) wz[n + i] = W[i], wzN[i] = W_e[i];
    for (int i = 0; i < n; ++i) wz[i] = W[nx + i];
    if (!s->dWz) {
        CUDA_OK(cudaMalloc((void **)&s->dWz, nz * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->dWzN, nx * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->dyref, (size_t)s->cap * nz * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->dyrefN, (size_t)s->cap * nx * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->dlamg, (size_t)s->cap * 2 * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->drowm, (size_t)s->cap * (s->Nmax + 1) * 6 * sizeof(double)));
    }
    CUDA_OK(cudaMemcpy(s->dWz, wz.data(), nz * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(s->dWzN, wzN.data(), nx * sizeof(double), cudaMemcpyHostToDevice));
    s->mpc_set = 1;
    return 0;
}

int vboc_set_mpc_reference(vboc_solver *s, int batch, const double *yref, const double *yref_e) {
    if (!s || s->family != VBOC_FAMILY_MPC || !s->mpc_set) return fail(VBOC_ERR_ARG, "vboc_set_mpc_reference: call vboc_set_mpc first");
    if (batch < 1 || batch > s->cap || !yref || !yref_e) return fail(VBOC_ERR_ARG, "vboc_set_mpc_reference: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    const int n = s->n, nx = 2 * n, nz = 3 * n;
    std::vector<double> yz((size_t)batch * nz);
    for (int b = 0; b < batch; ++b) {
        for (int i = 0; i < nx; ++i) yz[(size_t)b * nz + n + i] = yref[(size_t)b * nz + i];
        for (int i = 0; i < n; ++i) yz[(size_t)b * nz + i] = yref[(size_t)b * nz + nx + i];
    }
    int rc;
    if ((rc = h2d(s, s->dyref, yz.data(), yz.size() * sizeof(double)))) return rc;
    if ((rc = h2d(s, s->dyrefN, yref_e, (size_t)batch * nx * sizeof(double)))) return rc;
    s->mpc_ref_batch = batch;
    return 0;
}

int vboc_set_mpc_velnorm_start(vboc_solver *s, int vstart) {
    if (!s || s->family != VBOC_FAMILY_MPC || !s->mpc_set) return fail(VBOC_ERR_ARG, "vboc_set_mpc_velnorm_start: call vboc_set_mpc first");
    if (vstart < 0 || vstart > s->n) return fail(VBOC_ERR_ARG, "vboc_set_mpc_velnorm_start: 0 <= vstart <= n_dof");
    s->nn.vstart = vstart;
    return 0;
}

int vboc_set_mpc_rows(vboc_solver *s, int batch, const double *Z) {
    if (!s || s->family != VBOC_FAMILY_MPC || !s->mpc_set) return fail(VBOC_ERR_ARG, "vboc_set_mpc_rows: call vboc_set_mpc first");
    if (batch == 0) {  // back to the hard terminal row
        s->rows_soft = 0;
        return 0;
    }
    if (batch < 1 || batch > s->cap || !Z) return fail(VBOC_ERR_ARG, "vboc_set_mpc_rows: bad argument");
    const size_t cnt = (size_t)batch * (s->Nmax + 1) * 4;
    for (size_t i = 0; i < cnt; ++i)
        if ((i & 3) < 2 && !(Z[i] >= 0.0)) return fail(VBOC_ERR_ARG, "vboc_set_mpc_rows: Zl, Zu must be >= 0");
    CUDA_OK(cudaSetDevice(s->device));
    if (!s->drowZ) CUDA_OK(cudaMalloc((void **)&s->drowZ, (size_t)s->cap * (s->Nmax + 1) * 4 * sizeof(double)));
    int rc;
    if ((rc = h2d(s, s->drowZ, Z, cnt * sizeof(double)))) return rc;
    s->rows_soft = 1, s->rows_batch = batch;
    return 0;
}

int vboc_set_cartesian(vboc_solver *s, int on, double xc, double yc, double lh, double uh) {
    if (!s || s->family != VBOC_FAMILY_VBOC || s->n != 2)
        return fail(VBOC_ERR_ARG, "vboc_set_cartesian: a VBOC-family handle of the double pendulum is required");
    if (s->lane_kernel) return fail(VBOC_ERR_UNSUPPORTED, "vboc_set_cartesian: served by the warp kernel only");
    CUDA_OK(cudaSetDevice(s->device));
    if (!on) {
        s->cart_on = 0;
        return 0;
    }
    if (!(lh <= uh)) return fail(VBOC_ERR_ARG, "vboc_set_cartesian: lh > uh");
    const size_t need = Work<2>::doubles_rows(s->Nmax);
    if (s->work_doubles < need) {  // the row records live behind every warp slot's workspace
        CUDA_OK(cudaStreamSynchronize(s->stream));
        cudaFree(s->dwork);
        s->dwork = nullptr;
        s->work_doubles = need;
        CUDA_OK(cudaMalloc((void **)&s->dwork, (size_t)s->slots * s->work_doubles * sizeof(double)));
    }
    if (!s->drowm) CUDA_OK(cudaMalloc((void **)&s->drowm, (size_t)s->cap * (s->Nmax + 1) * 6 * sizeof(double)));
    s->cart_on = 1, s->cart_xc = xc, s->cart_yc = yc, s->mpc_lh = lh, s->mpc_uh = uh;
    return 0;
}

int vboc_download_mpc_rows(vboc_solver *s, double *rows) {
    if (!s || !s->batch || (s->family != VBOC_FAMILY_MPC && !s->cart_on) || !rows || !s->drowm)
        return fail(VBOC_ERR_ARG, "vboc_download_mpc_rows: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    return d2h(s, rows, s->drowm, (size_t)s->batch * (s->Nmax + 1) * 6 * sizeof(double));
}

int vboc_download_mpc_multipliers(vboc_solver *s, double *lamg) {
    if (!s || !s->batch || s->family != VBOC_FAMILY_MPC || !lamg) return fail(VBOC_ERR_ARG, "vboc_download_mpc_multipliers: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    return d2h(s, lamg, s->dlamg, (size_t)s->batch * 2 * sizeof(double));
}

int vboc_set_guess_network(vboc_solver *s, int hidden, int n_out, const float *W1, const float *b1, const float *W2,
                           const float *b2, const float *W3, const float *b3, double mean, double stdv) {
    if (!s || s->family != VBOC_FAMILY_AL) return fail(VBOC_ERR_ARG, "vboc_set_guess_network: not an AL-family solver");
    if (s->lane_kernel) return fail(VBOC_ERR_UNSUPPORTED, "vboc_set_guess_network: served by the warp kernel only");
    CUDA_OK(cudaSetDevice(s->device));
    if (hidden == 0) {  // switch it off
        s->gn_on = 0;
        return 0;
    }
    const int nx = 2 * s->n, H = hidden;
    if (H < 1 || H > NN_HMAX || n_out < nx || n_out % nx || n_out / nx > s->Nmax || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 ||
        !(stdv > 0.0))
        return fail(VBOC_ERR_ARG, "vboc_set_guess_network: bad argument (n_out must be N * 2n with N <= N_max)");
    const size_t o_b1 = (size_t)H * nx, o_W2T = o_b1 + H, o_b2 = o_W2T + (size_t)H * H, o_W3T = o_b2 + H,
                 o_b3 = o_W3T + (size_t)H * n_out, total = o_b3 + n_out;
    std::vector<double> h(total);
    for (size_t i = 0; i < (size_t)H * nx; ++i) h[i] = W1[i];
    for (int i = 0; i < H; ++i) h[o_b1 + i] = b1[i], h[o_b2 + i] = b2[i];
    for (int j = 0; j < H; ++j)
        for (int k = 0; k < H; ++k) h[o_W2T + (size_t)k * H + j] = W2[(size_t)j * H + k];
    for (int j = 0; j < n_out; ++j) {
        h[o_b3 + j] = b3[j];
        for (int k = 0; k < H; ++k) h[o_W3T + (size_t)k * n_out + j] = W3[(size_t)j * H + k];
    }
    if (s->dgn) cudaFree(s->dgn);
    s->dgn = nullptr;
    CUDA_OK(cudaMalloc((void **)&s->dgn, total * sizeof(double)));
    CUDA_OK(cudaMemcpy(s->dgn, h.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    GuessNet g;
    g.hidden = H, g.n_out = n_out, g.W1 = s->dgn, g.b1 = s->dgn + o_b1, g.W2T = s->dgn + o_W2T, g.b2 = s->dgn + o_b2;
    g.W3T = s->dgn + o_W3T, g.b3 = s->dgn + o_b3, g.mean = mean, g.stdv = stdv;
    if (!s->dgnn) CUDA_OK(cudaMalloc((void **)&s->dgnn, sizeof(GuessNet)));
    CUDA_OK(cudaMemcpy(s->dgnn, &g, sizeof(GuessNet), cudaMemcpyHostToDevice));
    if (!s->dxg_out) CUDA_OK(cudaMalloc((void **)&s->dxg_out, (size_t)s->cap * (s->Nmax + 1) * s->nxr * sizeof(double)));
    s->gn_on = 1, s->gn_out = n_out;
    return 0;
}

int vboc_download_guess(vboc_solver *s, double *x_guess) {
    if (!s || !s->batch || !s->gn_on || !x_guess) return fail(VBOC_ERR_ARG, "vboc_download_guess: no guess network active");
    CUDA_OK(cudaSetDevice(s->device));
    return d2h(s, x_guess, s->dxg_out, (size_t)s->batch * (s->Nmax + 1) * s->nxr * sizeof(double));
}

int vboc_export_multipliers(vboc_solver *s, int on) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_export_multipliers: null handle");
    CUDA_OK(cudaSetDevice(s->device));
    if (on && (s->lane_kernel || (s->n == 1 && s->family == VBOC_FAMILY_VBOC && s->free_dt)))
        return fail(VBOC_ERR_UNSUPPORTED, "vboc_export_multipliers: served by the warp kernel only");
    if (on && !s->dpi) {
        const size_t B = s->cap;
        CUDA_OK(cudaMalloc((void **)&s->dpi, B * s->Nmax * 2 * s->n * sizeof(double)));
        CUDA_OK(cudaMalloc((void **)&s->dlam, B * (s->Nmax + 1) * 6 * s->n * sizeof(double)));
    } else if (!on && s->dpi) {
        cudaFree(s->dpi), cudaFree(s->dlam);
        s->dpi = s->dlam = nullptr;
    }
    return 0;
}

int vboc_download_multipliers(vboc_solver *s, double *pi, double *lam) {
    if (!s || !s->batch) return fail(VBOC_ERR_ARG, "vboc_download_multipliers: nothing resident");
    if (!s->dpi) return fail(VBOC_ERR_ARG, "vboc_download_multipliers: call vboc_export_multipliers(s, 1) before the solve");
    if (s->free_dt) return fail(VBOC_ERR_UNSUPPORTED, "vboc_download_multipliers: not available for the free-dt kernel");
    CUDA_OK(cudaSetDevice(s->device));
    const size_t B = s->batch;
    int rc;
    if (pi && (rc = d2h(s, pi, s->dpi, B * s->Nmax * 2 * s->n * sizeof(double)))) return rc;
    if (lam && (rc = d2h(s, lam, s->dlam, B * (s->Nmax + 1) * 6 * s->n * sizeof(double)))) return rc;
    return 0;
}

int vboc_solve_batch(vboc_solver *s, int mode, int batch, const int *N, const double *x_guess,
                     const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                     const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                     const double *lbu, const double *ubu, const double *C0, double Tf, double *x,
                     double *u, vboc_stats *stats) {
    int rc = vboc_upload(s, batch, N, x_guess, u_guess, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu,
                         C0, Tf);
    if (rc) return rc;
    if ((rc = vboc_solve_resident(s, mode))) return rc;
    return vboc_download(s, x, u, stats);
}

struct vboc_mlp {
    int device, n_in, hidden, n_out, final_relu;
    float *W1, *b1, *W2T, *b2, *W3, *b3;  // CUDA-core kernel (W2 transposed)
    int Hp;                               // hidden size padded to a multiple of 32 (tensor-core kernels)
    float *W1p, *b1p, *W2p, *b2p, *W3p;   // zero padded
    unsigned char *W2img;                 // hi / lo split of W2p as shared-memory chunk images (mlp_pipe.cuh)
    int num_sms;
    // persistent, grow-on-demand I/O buffers of vboc_mlp_forward (host arrays in, host arrays out)
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    float *dx, *dout, *daux;
    int *dlab;
    long long io_rows;
    char *stage;                          // pinned staging for the copies
    size_t stage_bytes;
    double last_ms;                       // device time of the last forward kernel (CUDA events); < 0 before the first
};

static int mlp_create_impl(vboc_mlp *m, int device, int n_in, int hidden, int n_out, int final_relu, const float *W1,
                           const float *b1, const float *W2, const float *b2, const float *W3, const float *b3) {
    m->device = device, m->n_in = n_in, m->hidden = hidden, m->n_out = n_out, m->final_relu = final_relu;
    m->last_ms = -1.0;
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    m->num_sms = prop.multiProcessorCount;
    CUDA_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&m->ev0));
    CUDA_OK(cudaEventCreate(&m->ev1));
    m->stage_bytes = (size_t)16 << 20;
    CUDA_OK(cudaMallocHost((void **)&m->stage, m->stage_bytes));
    std::vector<float> w2t((size_t)hidden * hidden);
    for (int j = 0; j < hidden; ++j)
        for (int k = 0; k < hidden; ++k) w2t[(size_t)k * hidden + j] = W2[(size_t)j * hidden + k];
#define UPF(ptr, src, count)                                                         \
    CUDA_OK(cudaMalloc((void **)&m->ptr, (size_t)(count) * sizeof(float)));         \
    CUDA_OK(cudaMemcpy(m->ptr, src, (size_t)(count) * sizeof(float), cudaMemcpyHostToDevice))
    UPF(W1, W1, hidden * n_in);
    UPF(b1, b1, hidden);
    UPF(W2T, w2t.data(), (size_t)hidden * hidden);
    UPF(b2, b2, hidden);
    UPF(W3, W3, n_out * hidden);
    UPF(b3, b3, n_out);
    {   // zero-padded copies for the tcgen05 kernels
        const int Hp = (hidden + 31) / 32 * 32;
        m->Hp = Hp;
        std::vector<float> w1p((size_t)Hp * n_in, 0.f), b1p(Hp, 0.f), w2p((size_t)Hp * Hp, 0.f), b2p(Hp, 0.f),
            w3p((size_t)n_out * Hp, 0.f);
        for (int j = 0; j < hidden; ++j) {
            for (int i = 0; i < n_in; ++i) w1p[(size_t)j * n_in + i] = W1[(size_t)j * n_in + i];
            b1p[j] = b1[j], b2p[j] = b2[j];
            for (int k = 0; k < hidden; ++k) w2p[(size_t)j * Hp + k] = W2[(size_t)j * hidden + k];
            for (int o = 0; o < n_out; ++o) w3p[(size_t)o * Hp + j] = W3[(size_t)o * hidden + j];
        }
        UPF(W1p, w1p.data(), w1p.size());
        UPF(b1p, b1p.data(), b1p.size());
        UPF(W2p, w2p.data(), w2p.size());
        UPF(b2p, b2p.data(), b2p.size());
        UPF(W3p, w3p.data(), w3p.size());
        // W2 split into hi / lo ONCE, laid out as the shared-memory image of every K-chunk of the B operand:
        // chunk ch (K = 8 ch .. 8 ch + 7), part (hi, lo), 16-byte K-group c (2), output unit n (Hp + 1 padded), 4 floats
        if (Hp <= 512) {
            const TpLayout lay(Hp, n_in, n_out);
            const size_t nch = Hp / TP_KC, img = (size_t)lay.b_bytes / 4;  // floats per (hi or lo) image
            std::vector<float> blob(nch * 2 * img, 0.f);
            for (size_t ch = 0; ch < nch; ++ch)
                for (int c = 0; c < TP_KC / 4; ++c)
                    for (int nn = 0; nn < Hp; ++nn)
                        for (int e = 0; e < 4; ++e) {
                            const float wv = w2p[(size_t)nn * Hp + ch * TP_KC + 4 * c + e];
                            uint32_t bits;
                            memcpy(&bits, &wv, 4);
                            bits &= 0xFFFFE000u;
                            float hi;
                            memcpy(&hi, &bits, 4);
                            const size_t o = (size_t)c * (lay.b_lbo / 4) + (size_t)nn * 4 + e;
                            blob[(ch * 2 + 0) * img + o] = hi;
                            blob[(ch * 2 + 1) * img + o] = wv - hi;
                        }
            CUDA_OK(cudaMalloc((void **)&m->W2img, blob.size() * sizeof(float)));
            CUDA_OK(cudaMemcpy(m->W2img, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
#undef UPF
    return 0;
}

int vboc_mlp_create(int device, int n_in, int hidden, int n_out, int final_relu, const float *W1,
                    const float *b1, const float *W2, const float *b2, const float *W3, const float *b3,
                    vboc_mlp **out) {
    if (!out || n_in < 2 || n_in > MLP_MAX_IN || hidden < 1 || hidden > 1024 || n_out < 1 || n_out > MLP_MAX_OUT ||
        !W1 || !b1 || !W2 || !b2 || !W3 || !b3)
        return fail(VBOC_ERR_ARG, "vboc_mlp_create: bad argument");
    CUDA_OK(cudaSetDevice(device));
    vboc_mlp *m = new vboc_mlp();
    memset(m, 0, sizeof(*m));
    int rc = mlp_create_impl(m, device, n_in, hidden, n_out, final_relu, W1, b1, W2, b2, W3, b3);
    if (rc) {
        const std::string keep = g_err;
        vboc_mlp_destroy(m);
        g_err = keep;
        return rc;
    }
    *out = m;
    return 0;
}

void vboc_mlp_destroy(vboc_mlp *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    void *ptrs[] = {m->W1, m->b1, m->W2T, m->b2, m->W3, m->b3, m->W1p, m->b1p, m->W2p, m->b2p, m->W3p, m->W2img,
                    m->dx, m->dout, m->daux, m->dlab};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (m->stage) cudaFreeHost(m->stage);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    cudaGetLastError();
    delete m;
}

// Launch of the fused MLP on DEVICE arrays (x [batch][n_in] -> out [batch][n_out] / aux / label, any of the outputs may
// be null) on `stream`; events ev0 / ev1 of the handle bracket the kernel.  Kernel choice: the pipelined tcgen05
// kernel (mlp_pipe.cuh) for hidden sizes the 512-column TMEM holds; VBOC_MLP_SERIAL=1 selects the serial tcgen05
// kernel of round 1, VBOC_MLP_CUDA_CORES=1 the plain FP32 kernel (both kept as cross-checks).
static int mlp_launch(vboc_mlp *m, long long batch, const float *dx, int mode, double mean, double stdv,
                      double safety_margin, float *dout, float *daux, int *dlab, cudaStream_t stream) {
    MlpParams P;
    P.batch = (int)(batch > 0x7fffffff ? 0x7fffffff : batch), P.n_in = m->n_in, P.hidden = m->hidden, P.n_out = m->n_out;
    P.mode = mode, P.final_relu = m->final_relu;
    P.mean = (float)mean, P.stdv = (float)stdv, P.margin_scale = (float)((100.0 - safety_margin) / 100.0);
    P.W1 = m->W1, P.b1 = m->b1, P.W2T = m->W2T, P.b2 = m->b2, P.W3 = m->W3, P.b3 = m->b3;
    P.x = dx, P.out = dout, P.aux = daux, P.label = dlab;
    const char *force = getenv("VBOC_MLP_CUDA_CORES"), *serial = getenv("VBOC_MLP_SERIAL");
    CUDA_OK(cudaEventRecord(m->ev0, stream));
    if (m->Hp <= 512 && !(force && atoi(force))) {
        P.W1 = m->W1p, P.b1 = m->b1p, P.W2T = m->W2p, P.b2 = m->b2p, P.W3 = m->W3p;
        int cols = 32;
        while (cols < m->Hp) cols *= 2;
        const long long tiles = (batch + TC_ROWS - 1) / TC_ROWS;
        if (serial && atoi(serial)) {
            if (!dout) return fail(VBOC_ERR_ARG, "mlp: the serial kernel needs an output array");
            TcLayout lay(m->Hp);
            CUDA_OK(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total()));
            mlp_tc_kernel<<<(unsigned)tiles, TC_THREADS, lay.total(), stream>>>(P, m->Hp, cols);
        } else {
            TpLayout lay(m->Hp, m->n_in, m->n_out);
            const unsigned grid = (unsigned)(tiles < m->num_sms ? tiles : m->num_sms);  // persistent: one CTA per SM
            bool launched = false;
#define GO(NIN, NOUT)                                                                                                  \
    if (m->n_in == NIN && m->n_out == NOUT) {                                                                          \
        CUDA_OK(cudaFuncSetAttribute(mlp_pipe_kernel<NIN, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                     (int)lay.total()));                                                               \
        mlp_pipe_kernel<NIN, NOUT><<<grid, TC_THREADS, lay.total(), stream>>>(P, m->Hp, cols, m->W2img, batch);        \
        launched = true;                                                                                               \
    }
            GO(2, 1) GO(2, 2) GO(4, 1) GO(4, 2) GO(6, 1) GO(6, 2)
#undef GO
            if (!launched) {  // a network shape the reference does not have: the serial kernel handles any n_in / n_out
                if (!dout) return fail(VBOC_ERR_ARG, "mlp: this network shape needs an output array");
                TcLayout lay1(m->Hp);
                CUDA_OK(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay1.total()));
                mlp_tc_kernel<<<(unsigned)tiles, TC_THREADS, lay1.total(), stream>>>(P, m->Hp, cols);
            }
        }
    } else {
        if (!dout) return fail(VBOC_ERR_ARG, "mlp: the CUDA-core kernel needs an output array");
        size_t smem = mlp_smem_bytes(m->hidden);
        CUDA_OK(cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)((batch + MLP_ROWS - 1) / MLP_ROWS);
        mlp_forward_kernel<<<grid, MLP_THREADS, smem, stream>>>(P);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(m->ev1, stream));
    return 0;
}

static int mlp_finish(vboc_mlp *m, cudaStream_t stream) {
    CUDA_OK(cudaStreamSynchronize(stream));
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_ms = ms;
    return 0;
}

// pinned-staged copy on the handle's stream (host arrays may be pageable)
static int mlp_copy(vboc_mlp *m, void *dst, const void *src, size_t bytes, bool to_device) {
    const char *ps = (const char *)src;
    char *pd = (char *)dst;
    while (bytes) {
        size_t c = bytes < m->stage_bytes ? bytes : m->stage_bytes;
        if (to_device) {
            memcpy(m->stage, ps, c);
            CUDA_OK(cudaMemcpyAsync(pd, m->stage, c, cudaMemcpyHostToDevice, m->stream));
            CUDA_OK(cudaStreamSynchronize(m->stream));
        } else {
            CUDA_OK(cudaMemcpyAsync(m->stage, ps, c, cudaMemcpyDeviceToHost, m->stream));
            CUDA_OK(cudaStreamSynchronize(m->stream));
            memcpy(pd, m->stage, c);
        }
        ps += c, pd += c, bytes -= c;
    }
    return 0;
}

int vboc_mlp_forward(vboc_mlp *m, int batch, const float *x, int mode, double mean, double stdv,
                     double safety_margin, float *out, float *aux, int *label) {
    if (!m || batch < 1 || !x || !out || mode < 0 || mode > 2) return fail(VBOC_ERR_ARG, "vboc_mlp_forward: bad argument");
    CUDA_OK(cudaSetDevice(m->device));
    const size_t B = batch;
    if ((long long)B > m->io_rows) {  // grow-on-demand persistent buffers: no allocation on the steady-state path
        void *old[] = {m->dx, m->dout, m->daux, m->dlab};
        for (void *q : old)
            if (q) cudaFree(q);
        m->dx = m->dout = m->daux = nullptr, m->dlab = nullptr, m->io_rows = 0;
        const size_t rows = B + B / 4;
        CUDA_OK(cudaMalloc((void **)&m->dx, rows * m->n_in * sizeof(float)));
        CUDA_OK(cudaMalloc((void **)&m->dout, rows * m->n_out * sizeof(float)));
        CUDA_OK(cudaMalloc((void **)&m->daux, rows * sizeof(float)));
        CUDA_OK(cudaMalloc((void **)&m->dlab, rows * sizeof(int)));
        m->io_rows = (long long)rows;
    }
    int rc;
    if ((rc = mlp_copy(m, m->dx, x, B * m->n_in * sizeof(float), true))) return rc;
    if ((rc = mlp_launch(m, batch, m->dx, mode, mean, stdv, safety_margin, m->dout, m->daux, m->dlab, m->stream))) return rc;
    if ((rc = mlp_finish(m, m->stream))) return rc;
    if ((rc = mlp_copy(m, out, m->dout, B * m->n_out * sizeof(float), false))) return rc;
    if (aux && (rc = mlp_copy(m, aux, m->daux, B * sizeof(float), false))) return rc;
    if (label && (rc = mlp_copy(m, label, m->dlab, B * sizeof(int), false))) return rc;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Resident unlabeled pool of the AL drivers (include/vboc_b200.h: vboc_pool_*): the pool lives in HBM across
// rounds; scoring, the top-B query and the removal of the queried rows run on the device (pool_select.cuh).
struct vboc_pool {
    int device, n_in;
    long long cap, size;
    float *x[2];        // double buffer: the stable compaction writes the surviving rows to the other one
    int cur;
    float *score;
    unsigned int *hist;     // 2048 bins
    long long *sel;         // indices of the last selection (device), sel_count of them
    int *flags;             // 1 = selected (to be removed)
    unsigned int *counters; // [0] selected above the threshold, [1] ties taken
    long long *blk;         // per-block kept counts / offsets
    int sel_count;
    cudaStream_t stream;
    char *stage;
    size_t stage_bytes;
    double last_score_ms;
};

static int pool_copy(vboc_pool *s, void *dst, const void *src, size_t bytes, bool to_device) {
    const char *ps = (const char *)src;
    char *pd = (char *)dst;
    while (bytes) {
        size_t c = bytes < s->stage_bytes ? bytes : s->stage_bytes;
        if (to_device) {
            memcpy(s->stage, ps, c);
            CUDA_OK(cudaMemcpyAsync(pd, s->stage, c, cudaMemcpyHostToDevice, s->stream));
            CUDA_OK(cudaStreamSynchronize(s->stream));
        } else {
            CUDA_OK(cudaMemcpyAsync(s->stage, ps, c, cudaMemcpyDeviceToHost, s->stream));
            CUDA_OK(cudaStreamSynchronize(s->stream));
            memcpy(pd, s->stage, c);
        }
        ps += c, pd += c, bytes -= c;
    }
    return 0;
}

static int pool_create_impl(vboc_pool *s, int device, int n_in, long long capacity) {
    s->device = device, s->n_in = n_in, s->cap = capacity, s->size = 0, s->cur = 0, s->sel_count = 0, s->last_score_ms = -1.0;
    CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CUDA_OK(cudaMalloc((void **)&s->x[i], (size_t)capacity * n_in * sizeof(float)));
    CUDA_OK(cudaMalloc((void **)&s->score, (size_t)capacity * sizeof(float)));
    CUDA_OK(cudaMalloc((void **)&s->hist, 2048 * sizeof(unsigned int)));
    CUDA_OK(cudaMalloc((void **)&s->sel, (size_t)capacity * sizeof(long long)));
    CUDA_OK(cudaMalloc((void **)&s->flags, (size_t)capacity * sizeof(int)));
    CUDA_OK(cudaMalloc((void **)&s->counters, 4 * sizeof(unsigned int)));
    CUDA_OK(cudaMalloc((void **)&s->blk, (size_t)((capacity + POOL_BLOCK - 1) / POOL_BLOCK + 1) * sizeof(long long)));
    s->stage_bytes = (size_t)16 << 20;
    CUDA_OK(cudaMallocHost((void **)&s->stage, s->stage_bytes));
    return 0;
}

int vboc_pool_create(int device, int n_in, long long capacity, vboc_pool **out) {
    if (!out || n_in < 2 || n_in > MLP_MAX_IN || capacity < 1) return fail(VBOC_ERR_ARG, "vboc_pool_create: bad argument");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VBOC_ERR_CUDA, "vboc_pool_create: no such CUDA device");
    CUDA_OK(cudaSetDevice(device));
    vboc_pool *s = new vboc_pool();
    memset(s, 0, sizeof(*s));
    int rc = pool_create_impl(s, device, n_in, capacity);
    if (rc) {
        const std::string keep = g_err;
        vboc_pool_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return 0;
}

void vboc_pool_destroy(vboc_pool *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    void *ptrs[] = {s->x[0], s->x[1], s->score, s->hist, s->sel, s->flags, s->counters, s->blk};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (s->stage) cudaFreeHost(s->stage);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
}

long long vboc_pool_size(vboc_pool *s) { return s ? s->size : 0; }

int vboc_pool_upload(vboc_pool *s, long long count, const float *x) {
    if (!s || !x || count < 0 || count > s->cap) return fail(VBOC_ERR_ARG, "vboc_pool_upload: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    int rc = pool_copy(s, s->x[s->cur], x, (size_t)count * s->n_in * sizeof(float), true);
    if (rc) return rc;
    s->size = count, s->sel_count = 0;
    return 0;
}

int vboc_pool_score(vboc_pool *s, vboc_mlp *m, double mean, double stdv) {
    if (!s || !m || m->n_in != s->n_in) return fail(VBOC_ERR_ARG, "vboc_pool_score: bad argument");
    if (m->device != s->device) return fail(VBOC_ERR_ARG, "vboc_pool_score: pool and network live on different devices");
    if (s->size == 0) return 0;
    CUDA_OK(cudaSetDevice(s->device));
    int rc = mlp_launch(m, s->size, s->x[s->cur], 2, mean, stdv, 0.0, nullptr, s->score, nullptr, s->stream);
    if (rc) return rc;
    if ((rc = mlp_finish(m, s->stream))) return rc;
    s->last_score_ms = m->last_ms;
    return 0;
}

int vboc_pool_download_scores(vboc_pool *s, float *score) {
    if (!s || !score) return fail(VBOC_ERR_ARG, "vboc_pool_download_scores: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    return pool_copy(s, score, s->score, (size_t)s->size * sizeof(float), false);
}

int vboc_pool_select(vboc_pool *s, int k, long long *idx, float *x, float *score) {
    if (!s || k < 0 || !idx) return fail(VBOC_ERR_ARG, "vboc_pool_select: bad argument");
    if (k > s->size) return fail(VBOC_ERR_ARG, "vboc_pool_select: k exceeds the pool size");
    s->sel_count = 0;
    if (k == 0) return 0;
    CUDA_OK(cudaSetDevice(s->device));
    const long long P = s->size;
    const unsigned grid = (unsigned)((P + POOL_BLOCK - 1) / POOL_BLOCK);
    // radix select of the k-th largest score on the (order preserving) bit pattern: 11 + 11 + 10 bits
    unsigned int prefix = 0, prefix_mask = 0, hist[2048];
    long long need = k;  // how many of the elements matching the prefix are still to be taken from the top
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        CUDA_OK(cudaMemsetAsync(s->hist, 0, 2048 * sizeof(unsigned int), s->stream));
        pool_hist_kernel<<<grid, POOL_THREADS, 0, s->stream>>>(s->score, P, prefix, prefix_mask, shifts[pass], bits[pass], s->hist);
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpyAsync(hist, s->hist, 2048 * sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_OK(cudaStreamSynchronize(s->stream));
        int b = (1 << bits[pass]) - 1;
        for (; b > 0 && (long long)hist[b] < need; --b) need -= hist[b];
        prefix |= (unsigned)b << shifts[pass];
        prefix_mask |= (unsigned)((1 << bits[pass]) - 1) << shifts[pass];
    }
    // prefix = key of the k-th largest score; `need` of the elements with exactly that key are taken
    const long long above = k - need;
    CUDA_OK(cudaMemsetAsync(s->counters, 0, 4 * sizeof(unsigned int), s->stream));
    CUDA_OK(cudaMemsetAsync(s->flags, 0, (size_t)P * sizeof(int), s->stream));
    pool_pick_kernel<<<grid, POOL_THREADS, 0, s->stream>>>(s->score, P, prefix, (unsigned)above, (unsigned)need, s->sel,
                                                           s->flags, s->counters);
    CUDA_OK(cudaGetLastError());
    std::vector<long long> h(k);
    CUDA_OK(cudaMemcpyAsync(h.data(), s->sel, (size_t)k * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    std::sort(h.begin(), h.end(), [](long long a, long long b) { return a > b; });  // largest index first (the drivers' order)
    memcpy(idx, h.data(), (size_t)k * sizeof(long long));
    s->sel_count = k;
    if (x || score) {
        // the selected rows / scores in that order: gathered on the device, one copy back
        CUDA_OK(cudaMemcpyAsync(s->sel, h.data(), (size_t)k * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
        float *gx = s->x[s->cur ^ 1];  // the other buffer is scratch between compactions
        pool_gather_kernel<<<(k + 127) / 128, 128, 0, s->stream>>>(s->x[s->cur], s->score, s->sel, k, s->n_in, gx,
                                                                  gx + (size_t)k * s->n_in);
        CUDA_OK(cudaGetLastError());
        int rc;
        if (x && (rc = pool_copy(s, x, gx, (size_t)k * s->n_in * sizeof(float), false))) return rc;
        if (score && (rc = pool_copy(s, score, gx + (size_t)k * s->n_in, (size_t)k * sizeof(float), false))) return rc;
    }
    return 0;
}

int vboc_pool_remove_selected(vboc_pool *s) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_pool_remove_selected: null handle");
    if (s->sel_count == 0) return 0;
    CUDA_OK(cudaSetDevice(s->device));
    const long long P = s->size;
    const unsigned grid = (unsigned)((P + POOL_BLOCK - 1) / POOL_BLOCK);
    // stable compaction (np.delete keeps the order of the survivors): per-block kept counts, offsets on the host
    pool_count_kernel<<<grid, POOL_THREADS, 0, s->stream>>>(s->flags, P, s->blk);
    CUDA_OK(cudaGetLastError());
    std::vector<long long> cnt(grid + 1);
    CUDA_OK(cudaMemcpyAsync(cnt.data(), s->blk, grid * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    long long run = 0;
    for (unsigned b = 0; b < grid; ++b) {
        long long c = cnt[b];
        cnt[b] = run;
        run += c;
    }
    CUDA_OK(cudaMemcpyAsync(s->blk, cnt.data(), grid * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
    pool_compact_kernel<<<grid, POOL_THREADS, 0, s->stream>>>(s->x[s->cur], s->flags, P, s->n_in, s->blk, s->x[s->cur ^ 1]);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s->stream));
    if (run != P - s->sel_count) return fail(VBOC_ERR_CUDA, "vboc_pool_remove_selected: compaction count mismatch");
    s->cur ^= 1, s->size = run, s->sel_count = 0;
    return 0;
}

int vboc_pool_download(vboc_pool *s, float *x) {
    if (!s || !x) return fail(VBOC_ERR_ARG, "vboc_pool_download: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    return pool_copy(s, x, s->x[s->cur], (size_t)s->size * s->n_in * sizeof(float), false);
}

double vboc_pool_last_score_ms(vboc_pool *s) { return s ? s->last_score_ms : -1.0; }

double vboc_mlp_last_kernel_ms(vboc_mlp *m) { return m ? m->last_ms : -1.0; }

int vboc_fp64_peak(int device, double *tflops) {
    if (!tflops) return fail(VBOC_ERR_ARG, "vboc_fp64_peak: null argument");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *out = nullptr;
    CUDA_OK(cudaMalloc((void **)&out, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_OK(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(out, iters);
        CUDA_OK(cudaEventRecord(e1));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(out);
    *tflops = best;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Device-resident data generation (include/vboc_b200.h: vboc_datagen_*)
struct vboc_datagen {
    int n, cap, device, grid;
    vboc_opts opts;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int *djs;
    double *dp, *dlb0, *dub0, *dretry, *drows, *dwork;
    DgCounters *dcnt;
    unsigned int *dcounter;  // [0] problem counter, [2..3] kernel start time (64 bit)
    size_t work_doubles;
    char *stage;  // pinned staging
    size_t stage_bytes;
    double last_ms;
};
static_assert(sizeof(DgCounters) == sizeof(vboc_dg_stats), "vboc_dg_stats mirrors DgCounters");

static int datagen_create_impl(vboc_datagen *s, int n_dof, int capacity, int device) {
    s->n = n_dof, s->cap = capacity, s->device = device, s->last_ms = -1.0;
    vboc_default_opts(VBOC_FAMILY_VBOC, &s->opts);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    int max_grid = prop.multiProcessorCount * 5, need = (capacity + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    s->grid = need < max_grid ? need : max_grid;
    s->work_doubles = n_dof == 2 ? Work<2>::TOTAL + DgWork<2>::TOTAL : Work<3>::TOTAL + DgWork<3>::TOTAL;
    const size_t B = capacity, nxr = 2 * n_dof + 1, nx = 2 * n_dof;
    CUDA_OK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&s->ev0));
    CUDA_OK(cudaEventCreate(&s->ev1));
    CUDA_OK(cudaMalloc((void **)&s->djs, B * sizeof(int)));
    CUDA_OK(cudaMalloc((void **)&s->dp, B * (n_dof + 1) * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&s->dlb0, B * nxr * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&s->dub0, B * nxr * sizeof(double)));
    {   // one buffer serves both state machines' restart tables: 10 x (n + 1) (data generation), 60 x 2n (test data)
        const size_t per = (size_t)DG_RETRIES * (n_dof + 1) > (size_t)TG_MAX_SOLVES * 2 * n_dof ? (size_t)DG_RETRIES * (n_dof + 1)
                                                                                              : (size_t)TG_MAX_SOLVES * 2 * n_dof;
        CUDA_OK(cudaMalloc((void **)&s->dretry, B * per * sizeof(double)));
    }
    CUDA_OK(cudaMalloc((void **)&s->drows, B * DG_ROWS_MAX * nx * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&s->dcnt, B * sizeof(DgCounters)));
    CUDA_OK(cudaMalloc((void **)&s->dcounter, 4 * sizeof(unsigned int)));
    CUDA_OK(cudaMalloc((void **)&s->dwork, (size_t)s->grid * WARPS_PER_CTA * s->work_doubles * sizeof(double)));
    s->stage_bytes = (size_t)8 << 20;
    CUDA_OK(cudaMallocHost((void **)&s->stage, s->stage_bytes));
    return 0;
}

int vboc_datagen_create(int n_dof, int capacity, int device, vboc_datagen **out) {
    if (!out || (n_dof != 2 && n_dof != 3) || capacity < 1) return fail(VBOC_ERR_ARG, "vboc_datagen_create: bad argument");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VBOC_ERR_CUDA, "vboc_datagen_create: no such CUDA device");
    CUDA_OK(cudaSetDevice(device));
    vboc_datagen *s = new vboc_datagen();
    memset(s, 0, sizeof(*s));
    int rc = datagen_create_impl(s, n_dof, capacity, device);
    if (rc) {
        const std::string keep = g_err;
        vboc_datagen_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return 0;
}

void vboc_datagen_destroy(vboc_datagen *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    void *ptrs[] = {s->djs, s->dp, s->dlb0, s->dub0, s->dretry, s->drows, s->dcnt, s->dcounter, s->dwork};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (s->stage) cudaFreeHost(s->stage);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
}

int vboc_datagen_set_opts(vboc_datagen *s, const vboc_opts *o) {
    if (!s || !o) return fail(VBOC_ERR_ARG, "vboc_datagen_set_opts: null argument");
    s->opts = *o;
    return 0;
}

double vboc_datagen_last_kernel_ms(vboc_datagen *s) { return s ? s->last_ms : -1.0; }

// pinned-staged copies on the handle's stream
static int dg_copy(vboc_datagen *s, void *dst, const void *src, size_t bytes, bool to_device) {
    const char *ps = (const char *)src;
    char *pd = (char *)dst;
    while (bytes) {
        size_t c = bytes < s->stage_bytes ? bytes : s->stage_bytes;
        if (to_device) {
            memcpy(s->stage, ps, c);
            CUDA_OK(cudaMemcpyAsync(pd, s->stage, c, cudaMemcpyHostToDevice, s->stream));
            CUDA_OK(cudaStreamSynchronize(s->stream));
        } else {
            CUDA_OK(cudaMemcpyAsync(s->stage, ps, c, cudaMemcpyDeviceToHost, s->stream));
            CUDA_OK(cudaStreamSynchronize(s->stream));
            memcpy(pd, s->stage, c);
        }
        ps += c, pd += c, bytes -= c;
    }
    return 0;
}

int vboc_datagen_run(vboc_datagen *s, int count, int N0, double dt, double tol, const int *joint_sel, const double *p,
                     const double *lb0, const double *ub0, const double *retry, double *rows, long long rows_capacity,
                     long long *total_rows, vboc_dg_stats *stats) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_datagen_run: null handle");
    if (count < 1 || count > s->cap) return fail(VBOC_ERR_ARG, "vboc_datagen_run: count exceeds capacity");
    if (!joint_sel || !p || !lb0 || !ub0 || !retry || !rows || !stats || !total_rows)
        return fail(VBOC_ERR_ARG, "vboc_datagen_run: null array");
    if (N0 < 2 || N0 > DG_N_CAP || !(dt > 0.0) || !(tol > 0.0)) return fail(VBOC_ERR_ARG, "vboc_datagen_run: bad N0 / dt / tol");
    const int n = s->n;
    const size_t B = count, nxr = 2 * n + 1, nx = 2 * n;
    for (int b = 0; b < count; ++b) {
        if (joint_sel[b] < 0 || joint_sel[b] >= n) return fail(VBOC_ERR_ARG, "vboc_datagen_run: joint_sel out of range");
        if (lb0[b * nxr + 2 * n] != dt || ub0[b * nxr + 2 * n] != dt)
            return fail(VBOC_ERR_UNSUPPORTED, "vboc_datagen_run: the dt state must be pinned to dt");
    }
    CUDA_OK(cudaSetDevice(s->device));
    int rc;
    if ((rc = dg_copy(s, s->djs, joint_sel, B * sizeof(int), true))) return rc;
    if ((rc = dg_copy(s, s->dp, p, B * (n + 1) * sizeof(double), true))) return rc;
    if ((rc = dg_copy(s, s->dlb0, lb0, B * nxr * sizeof(double), true))) return rc;
    if ((rc = dg_copy(s, s->dub0, ub0, B * nxr * sizeof(double), true))) return rc;
    if ((rc = dg_copy(s, s->dretry, retry, B * DG_RETRIES * (n + 1) * sizeof(double), true))) return rc;
    DgParams P;
    P.N0 = N0, P.dt = dt, P.tol = tol;
    // limits of the reference models (VBOC/triplependulum_class_vboc.py:90-93, VBOC/doublependulum_class_vboc.py:114-117)
    P.q_min = M_PI - M_PI / 4, P.q_max = M_PI + M_PI / 4, P.v_max = 10.0, P.u_max = 10.0;
    CUDA_OK(cudaMemsetAsync(s->dcounter, 0, 4 * sizeof(unsigned int), s->stream));
    CUDA_OK(cudaEventRecord(s->ev0, s->stream));
    unsigned long long *t_start = reinterpret_cast<unsigned long long *>(s->dcounter + 2);
    if (n == 2) {
        DgIO<2> io{s->djs, s->dp, s->dlb0, s->dub0, s->dretry, s->drows, s->dcnt};
        datagen_kernel<2><<<s->grid, WARPS_PER_CTA * 32, 0, s->stream>>>(io, count, P, s->opts, s->dwork, s->work_doubles,
                                                                       s->dcounter, t_start);
    } else {
        DgIO<3> io{s->djs, s->dp, s->dlb0, s->dub0, s->dretry, s->drows, s->dcnt};
        datagen_kernel<3><<<s->grid, WARPS_PER_CTA * 32, 0, s->stream>>>(io, count, P, s->opts, s->dwork, s->work_doubles,
                                                                       s->dcounter, t_start);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(VBOC_ERR_CUDA, std::string("datagen_kernel launch: ") + cudaGetErrorString(e));
    CUDA_OK(cudaEventRecord(s->ev1, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->last_ms = ms;
    if ((rc = dg_copy(s, stats, s->dcnt, B * sizeof(DgCounters), false))) return rc;
    // rows: compacted on the device (problem order), one copy back
    std::vector<long long> off(B + 1, 0);
    for (size_t b = 0; b < B; ++b) off[b + 1] = off[b] + stats[b].n_rows;
    *total_rows = off[B];
    if (off[B] > rows_capacity) return fail(VBOC_ERR_ARG, "vboc_datagen_run: rows_capacity too small (see *total_rows)");
    if (off[B] > 0) {
        // the per-slot workspaces are idle now: their head serves as the offset table and the compact buffer
        long long *doff = reinterpret_cast<long long *>(s->dwork);
        double *dcompact = s->dwork + ((B + 2) & ~(size_t)1);
        const size_t avail = (size_t)s->grid * WARPS_PER_CTA * s->work_doubles;
        if (((B + 2) & ~(size_t)1) + (size_t)off[B] * nx > avail) return fail(VBOC_ERR_CUDA, "vboc_datagen_run: compaction buffer");
        if ((rc = dg_copy(s, doff, off.data(), (B + 1) * sizeof(long long), true))) return rc;
        dg_compact_kernel<<<(unsigned)B, 128, 0, s->stream>>>(s->drows, s->dcnt, doff, dcompact, (int)nx);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(VBOC_ERR_CUDA, std::string("dg_compact_kernel launch: ") + cudaGetErrorString(e));
        if ((rc = dg_copy(s, rows, dcompact, (size_t)off[B] * nx * sizeof(double), false))) return rc;
    }
    return 0;
}

int vboc_testdata_run(vboc_datagen *s, int count, int N0, double dt, int max_solves, const double *ran,
                      const double *q_init, const double *retry, double *rows, vboc_dg_stats *stats) {
    if (!s) return fail(VBOC_ERR_ARG, "vboc_testdata_run: null handle");
    if (count < 1 || count > s->cap) return fail(VBOC_ERR_ARG, "vboc_testdata_run: count exceeds capacity");
    if (!ran || !q_init || !retry || !rows || !stats) return fail(VBOC_ERR_ARG, "vboc_testdata_run: null array");
    if (N0 < 2 || N0 > DG_N_CAP || !(dt > 0.0) || max_solves < 1 || max_solves > TG_MAX_SOLVES)
        return fail(VBOC_ERR_ARG, "vboc_testdata_run: bad N0 / dt / max_solves");
    const int n = s->n;
    const size_t B = count, nx = 2 * n;
    CUDA_OK(cudaSetDevice(s->device));
    int rc;
    // the data-generation input buffers are reused: dp holds ran, dlb0 holds q_init
    if ((rc = dg_copy(s, s->dp, ran, B * n * sizeof(double), true))) return rc;
    if ((rc = dg_copy(s, s->dlb0, q_init, B * n * sizeof(double), true))) return rc;
    if ((rc = dg_copy(s, s->dretry, retry, B * TG_MAX_SOLVES * 2 * n * sizeof(double), true))) return rc;
    DgParams P;
    P.N0 = N0, P.dt = dt, P.tol = 1e-3;
    P.q_min = M_PI - M_PI / 4, P.q_max = M_PI + M_PI / 4, P.v_max = 10.0, P.u_max = 10.0;
    CUDA_OK(cudaMemsetAsync(s->dcounter, 0, 4 * sizeof(unsigned int), s->stream));
    CUDA_OK(cudaEventRecord(s->ev0, s->stream));
    unsigned long long *t_start = reinterpret_cast<unsigned long long *>(s->dcounter + 2);
    // triplependulum_testdata.py:82 compares with the cost rounded to 3 decimals minus 1e-3, doublependulum_testdata.py:80
    // with 4 decimals minus 1e-4
    if (n == 2) {
        TestIO<2> io{s->dp, s->dlb0, s->dretry, s->drows, s->dcnt, max_solves, 4, 1e-4};
        testing_kernel<2><<<s->grid, WARPS_PER_CTA * 32, 0, s->stream>>>(io, count, P, s->opts, s->dwork, s->work_doubles,
                                                                       s->dcounter, t_start);
    } else {
        TestIO<3> io{s->dp, s->dlb0, s->dretry, s->drows, s->dcnt, max_solves, 3, 1e-3};
        testing_kernel<3><<<s->grid, WARPS_PER_CTA * 32, 0, s->stream>>>(io, count, P, s->opts, s->dwork, s->work_doubles,
                                                                       s->dcounter, t_start);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(VBOC_ERR_CUDA, std::string("testing_kernel launch: ") + cudaGetErrorString(e));
    CUDA_OK(cudaEventRecord(s->ev1, s->stream));
    CUDA_OK(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->last_ms = ms;
    if ((rc = dg_copy(s, stats, s->dcnt, B * sizeof(DgCounters), false))) return rc;
    return dg_copy(s, rows, s->drows, B * nx * sizeof(double), false);
}

// vboc_sim_step keeps one grow-on-demand set of device / pinned buffers and a non-blocking stream per device
// (the drop-in simulator shim calls it once per RK4 step: no cudaMalloc, no legacy-stream serialisation per call)
namespace {
struct SimCache {
    double *d = nullptr, *h = nullptr;  // device / pinned host: x | u | x_next
    size_t rows = 0;
    cudaStream_t stream = nullptr;
};
SimCache g_sim[64];
std::mutex g_sim_mu;
}  // namespace

int vboc_sim_step(int n_dof, int device, int batch, const double *x, const double *u, double T,
                  double *x_next) {
    if (n_dof < 1 || n_dof > 3 || batch < 1 || !x || !u || !x_next || device < 0 || device >= 64)
        return fail(VBOC_ERR_ARG, "vboc_sim_step: bad argument");
    CUDA_OK(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_sim_mu);
    SimCache &c = g_sim[device];
    const size_t B = batch, nx = 2 * n_dof, per_row = 5 * 3;  // doubles per row, sized for n = 3
    if (!c.stream) CUDA_OK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    if (B > c.rows) {
        if (c.d) cudaFree(c.d);
        if (c.h) cudaFreeHost(c.h);
        c.d = c.h = nullptr, c.rows = 0;
        size_t rows = B < 1024 ? 1024 : B + B / 2;
        CUDA_OK(cudaMalloc((void **)&c.d, rows * per_row * sizeof(double)));
        CUDA_OK(cudaMallocHost((void **)&c.h, rows * per_row * sizeof(double)));
        c.rows = rows;
    }
    double *dx = c.d, *du = c.d + B * nx, *dxn = du + B * n_dof;
    memcpy(c.h, x, B * nx * sizeof(double));
    memcpy(c.h + B * nx, u, B * n_dof * sizeof(double));
    CUDA_OK(cudaMemcpyAsync(dx, c.h, B * (nx + n_dof) * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    int th = 128, bl = (batch + th - 1) / th;
    if (n_dof == 1) sim_kernel<1><<<bl, th, 0, c.stream>>>(batch, dx, du, T, dxn);
    if (n_dof == 2) sim_kernel<2><<<bl, th, 0, c.stream>>>(batch, dx, du, T, dxn);
    if (n_dof == 3) sim_kernel<3><<<bl, th, 0, c.stream>>>(batch, dx, du, T, dxn);
    CUDA_OK(cudaGetLastError());
    double *hxn = c.h + B * (nx + n_dof);
    CUDA_OK(cudaMemcpyAsync(hxn, dxn, B * nx * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_OK(cudaStreamSynchronize(c.stream));
    memcpy(x_next, hxn, B * nx * sizeof(double));
    return 0;
}

int vboc_stream_create(int n_dof, int family, int capacity, int N_max, int device, vboc_stream **out) {
    if (!out || n_dof < 1 || n_dof > 3 || (family != VBOC_FAMILY_VBOC && family != VBOC_FAMILY_AL) || capacity < 1 ||
        N_max < 1 || N_max > VBOC_N_MAX)
        return fail(VBOC_ERR_ARG, "vboc_stream_create: bad argument");
    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VBOC_ERR_CUDA, "vboc_stream_create: no such CUDA device");
    CUDA_OK(cudaSetDevice(device));
    vboc_stream *s = new vboc_stream();
    int rc = stream_create_impl(s, n_dof, family, capacity, N_max, device);
    if (rc) {
        const std::string keep = g_err;
        vboc_stream_destroy(s);
        g_err = keep;
        return rc;
    }
    *out = s;
    return 0;
}

static int stream_create_impl(vboc_stream *s, int n_dof, int family, int capacity, int N_max, int device) {
    s->n = n_dof, s->family = family, s->cap = capacity, s->Nmax = N_max, s->device = device;
    s->nxr = 2 * n_dof + (family == VBOC_FAMILY_VBOC), s->nu = n_dof;
    s->launches = s->solved = s->idle_polls = 0;
    vboc_default_opts(family, &s->opts);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    s->num_sms = prop.multiProcessorCount;
    size_t C = (size_t)capacity, nxr = s->nxr, nu = s->nu;
#define HOST_ALLOC(ptr, count) \
    CUDA_OK(cudaHostAlloc((void **)&s->ptr, (size_t)(count) * sizeof(*s->ptr), cudaHostAllocMapped | cudaHostAllocPortable))
    HOST_ALLOC(N, C);
    HOST_ALLOC(done, C);
    HOST_ALLOC(xg, C * (N_max + 1) * nxr);
    HOST_ALLOC(ug, C * N_max * nu);
    HOST_ALLOC(p, C * (n_dof + 1));
    HOST_ALLOC(lbx0, C * nxr);
    HOST_ALLOC(ubx0, C * nxr);
    HOST_ALLOC(lbx, C * nxr);
    HOST_ALLOC(ubx, C * nxr);
    HOST_ALLOC(lbxN, C * nxr);
    HOST_ALLOC(ubxN, C * nxr);
    HOST_ALLOC(lbu, C * nu);
    HOST_ALLOC(ubu, C * nu);
    HOST_ALLOC(dir, C * n_dof);
    HOST_ALLOC(h, C);
    HOST_ALLOC(x, C * (N_max + 1) * nxr);
    HOST_ALLOC(u, C * N_max * nu);
    HOST_ALLOC(st, C);
    s->sim_cap = capacity;
    HOST_ALLOC(sx, C * 2 * n_dof);
    HOST_ALLOC(su, C * n_dof);
    HOST_ALLOC(sxn, C * 2 * n_dof);
    for (int r = 0; r < vboc_stream::NREC; ++r) {
        HOST_ALLOC(rec[r].index, C);
        CUDA_OK(cudaStreamCreateWithFlags(&s->rec[r].stream, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&s->rec[r].ev, cudaEventDisableTiming));
        s->rec[r].busy = false;
    }
#undef HOST_ALLOC
    memset(s->done, 0, C * sizeof(int));
    CUDA_OK(cudaStreamCreateWithFlags(&s->sim_stream, cudaStreamNonBlocking));
    // per-SM workspace pool: 5 CTAs of solve_kernel<.., 5, STREAM> are resident per SM at most; one spare
    unsigned int *dn = nullptr, nsmid = 0;
    CUDA_OK(cudaMalloc((void **)&dn, sizeof(unsigned int)));
    nsmid_kernel<<<1, 1>>>(dn);
    CUDA_OK(cudaMemcpy(&nsmid, dn, sizeof(nsmid), cudaMemcpyDeviceToHost));
    cudaFree(dn);
    s->nsmid = (int)nsmid > s->num_sms ? (int)nsmid : s->num_sms;
    s->ws_per_sm = 6;
    s->work_doubles = work_doubles_for(n_dof, N_max);
    CUDA_OK(cudaMalloc((void **)&s->dwork,
                       (size_t)s->nsmid * s->ws_per_sm * WARPS_PER_CTA * s->work_doubles * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&s->ws_mask, (size_t)s->nsmid * sizeof(unsigned int)));
    CUDA_OK(cudaMemset(s->ws_mask, 0, (size_t)s->nsmid * sizeof(unsigned int)));
    CUDA_OK(cudaMalloc((void **)&s->dcounters, vboc_stream::NREC * sizeof(unsigned int)));
    s->free_slots.reserve(C);
    for (int i = capacity - 1; i >= 0; --i) s->free_slots.push_back(i);
    return 0;
}

void vboc_stream_destroy(vboc_stream *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    void *hp[] = {s->N,    s->done,  s->xg,  s->ug,  s->p,   s->lbx0, s->ubx0, s->lbx, s->ubx, s->lbxN, s->ubxN,
                  s->lbu,  s->ubu,   s->dir, s->h,   s->x,   s->u,    s->st,   s->sx,  s->su,  s->sxn};
    for (void *q : hp)
        if (q) cudaFreeHost(q);
    for (int r = 0; r < vboc_stream::NREC; ++r) {
        if (s->rec[r].index) cudaFreeHost(s->rec[r].index);
        if (s->rec[r].stream) cudaStreamDestroy(s->rec[r].stream);
        if (s->rec[r].ev) cudaEventDestroy(s->rec[r].ev);
    }
    if (s->sim_stream) cudaStreamDestroy(s->sim_stream);
    if (s->dwork) cudaFree(s->dwork);
    if (s->ws_mask) cudaFree(s->ws_mask);
    if (s->dcounters) cudaFree(s->dcounters);
    cudaGetLastError();
    delete s;
}

int vboc_stream_set_opts(vboc_stream *s, const vboc_opts *o) {
    if (!s || !o) return fail(VBOC_ERR_ARG, "vboc_stream_set_opts: null argument");
    s->opts = *o;
    return 0;
}

int vboc_stream_free_slots(vboc_stream *s) { return s ? (int)s->free_slots.size() : 0; }
int vboc_stream_pending(vboc_stream *s) { return s ? (int)s->inflight.size() : 0; }

int vboc_stream_submit(vboc_stream *s, int mode, int count, const int *N, const double *x_guess,
                       const double *u_guess, const double *p, const double *lbx0, const double *ubx0,
                       const double *lbx, const double *ubx, const double *lbxN, const double *ubxN,
                       const double *lbu, const double *ubu, const double *C0, double Tf, int *tickets) {
    if (!s || !tickets) return fail(VBOC_ERR_ARG, "vboc_stream_submit: null argument");
    if (mode != VBOC_MODE_SQP && mode != VBOC_MODE_RTI) return fail(VBOC_ERR_ARG, "vboc_stream_submit: bad mode");
    if (count < 1 || count > (int)s->free_slots.size())
        return fail(VBOC_ERR_ARG, "vboc_stream_submit: count exceeds the free slots (vboc_stream_free_slots)");
    if (!N || !x_guess || !u_guess || !lbx0 || !ubx0 || !lbx || !ubx || !lbxN || !ubxN || !lbu || !ubu)
        return fail(VBOC_ERR_ARG, "vboc_stream_submit: null array");
    if (s->family == VBOC_FAMILY_VBOC && !p) return fail(VBOC_ERR_ARG, "vboc_stream_submit: p required");
    CUDA_OK(cudaSetDevice(s->device));
    const int n = s->n, nxr = s->nxr, nu = s->nu, Nmax = s->Nmax;
    // a free launch record (its previous kernel has finished)
    int r = -1;
    for (int spin = 0; r < 0; ++spin) {
        for (int i = 0; i < vboc_stream::NREC && r < 0; ++i) {
            if (s->rec[i].busy && cudaEventQuery(s->rec[i].ev) == cudaSuccess) s->rec[i].busy = false;
            if (!s->rec[i].busy) r = i;
        }
        cudaGetLastError();  // cudaErrorNotReady of the queries is not an error
        if (r < 0) std::this_thread::yield();  // every record holds a running launch: wait for one to finish
    }
    // validate everything before touching a slot
    std::vector<double> h(count), dir((size_t)count * n);
    for (int b = 0; b < count; ++b) {
        bool pinned = true;
        int rc = prepare_problem(n, nxr, Nmax, s->family, N[b], lbx0 + (size_t)b * nxr, ubx0 + (size_t)b * nxr,
                                 lbx + (size_t)b * nxr, ubx + (size_t)b * nxr, lbxN + (size_t)b * nxr,
                                 ubxN + (size_t)b * nxr, x_guess + (size_t)b * (Nmax + 1) * nxr,
                                 C0 ? C0 + (size_t)b * n * nxr : nullptr, Tf, &h[b], &dir[(size_t)b * n], &pinned);
        if (rc) return rc;
        if (!pinned)
            return fail(VBOC_ERR_UNSUPPORTED, "vboc_stream_submit: a free dt state is served by vboc_solve_batch only");
        if (!C0) dir[(size_t)b * n] = NAN;  // no direction constraint for this problem
    }
    vboc_stream::Rec &R = s->rec[r];
    for (int b = 0; b < count; ++b) {
        const int t = s->free_slots.back();
        s->free_slots.pop_back();
        tickets[b] = t, R.index[b] = t;
        s->N[t] = N[b], s->h[t] = h[b], s->done[t] = 0;
        const size_t xs = (size_t)(Nmax + 1) * nxr, us = (size_t)Nmax * nu;
        memcpy(s->xg + t * xs, x_guess + b * xs, (size_t)(N[b] + 1) * nxr * sizeof(double));
        memcpy(s->ug + t * us, u_guess + b * us, (size_t)N[b] * nu * sizeof(double));
        if (p) memcpy(s->p + (size_t)t * (n + 1), p + (size_t)b * (n + 1), (n + 1) * sizeof(double));
#define ROW(dst, src, w) memcpy(s->dst + (size_t)t * (w), src + (size_t)b * (w), (w) * sizeof(double))
        ROW(lbx0, lbx0, nxr), ROW(ubx0, ubx0, nxr), ROW(lbx, lbx, nxr), ROW(ubx, ubx, nxr);
        ROW(lbxN, lbxN, nxr), ROW(ubxN, ubxN, nxr), ROW(lbu, lbu, nu), ROW(ubu, ubu, nu);
#undef ROW
        memcpy(s->dir + (size_t)t * n, &dir[(size_t)b * n], n * sizeof(double));
        s->inflight.push_back(t);
    }
    Batch B;
    B.batch = count, B.Nmax = Nmax, B.nxr = nxr;
    B.N = s->N, B.xg = s->xg, B.ug = s->ug, B.p = p ? s->p : nullptr;
    B.lbx0 = s->lbx0, B.ubx0 = s->ubx0, B.lbx = s->lbx, B.ubx = s->ubx, B.lbxN = s->lbxN, B.ubxN = s->ubxN;
    B.lbu = s->lbu, B.ubu = s->ubu, B.dir = s->dir, B.h = s->h, B.x = s->x, B.u = s->u, B.st = s->st;
    B.work = s->dwork, B.work_doubles = s->work_doubles, B.counter = s->dcounters + r;
    B.mode = mode, B.opts = s->opts;
    B.index = R.index, B.done = s->done, B.ws_mask = s->ws_mask, B.ws_per_sm = s->ws_per_sm;
    // on any failure from here on the slots go back to the free list and the tickets leave `inflight`: nothing
    // was launched for them, so a poll loop must not wait for them
    auto rollback = [&]() {
        for (int b = 0; b < count; ++b) {
            s->free_slots.push_back(tickets[b]);
            s->inflight.pop_back();
        }
    };
    cudaError_t e = cudaMemsetAsync(s->dcounters + r, 0, sizeof(unsigned int), R.stream);
    if (e != cudaSuccess) {
        rollback();
        return fail(VBOC_ERR_CUDA, std::string("vboc_stream_submit: counter reset: ") + cudaGetErrorString(e));
    }
    int grid = (count + WARPS_PER_CTA - 1) / WARPS_PER_CTA, maxg = s->num_sms * 5;
    if (grid > maxg) grid = maxg;
    e = cudaErrorInvalidValue;
#define GO(NQ, FAM) \
    if (n == NQ && s->family == FAM) e = launch_stream<NQ, FAM>(grid, R.stream, B);
    VB_ALL_SYSTEMS(GO)
#undef GO
    if (e != cudaSuccess) {
        rollback();
        return fail(VBOC_ERR_CUDA, std::string("solve_kernel (stream) launch: ") + cudaGetErrorString(e));
    }
    // the kernel is queued: if the event cannot be recorded the record is simply treated as busy until the stream
    // drains (cudaEventQuery on an unrecorded event reports success)
    if ((e = cudaEventRecord(R.ev, R.stream)) != cudaSuccess) cudaStreamSynchronize(R.stream);
    R.busy = true;
    ++s->launches;
    return 0;
}

int vboc_stream_poll(vboc_stream *s, int max, int *tickets) {
    if (!s || !tickets || max < 0) return fail(VBOC_ERR_ARG, "vboc_stream_poll: bad argument");
    int nout = 0;
    size_t keep = 0;
    for (size_t i = 0; i < s->inflight.size(); ++i) {
        const int t = s->inflight[i];
        if (nout < max && __atomic_load_n(&s->done[t], __ATOMIC_ACQUIRE) == 1) {
            s->done[t] = 2;  // finished, waiting for vboc_stream_fetch
            tickets[nout++] = t;
        } else {
            s->inflight[keep++] = t;
        }
    }
    s->inflight.resize(keep);
    if (nout == 0 && keep && (++s->idle_polls & 255) == 0) {
        // now and then: surface a failed launch / a faulted kernel instead of polling forever
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) return fail(VBOC_ERR_CUDA, std::string("vboc_stream_poll: ") + cudaGetErrorString(e));
        for (int i = 0; i < vboc_stream::NREC; ++i)
            if (s->rec[i].busy) {
                e = cudaEventQuery(s->rec[i].ev);
                if (e == cudaSuccess) s->rec[i].busy = false;
                else if (e != cudaErrorNotReady)
                    return fail(VBOC_ERR_CUDA, std::string("vboc_stream_poll: ") + cudaGetErrorString(e));
            }
        cudaGetLastError();
    }
    return nout;
}

int vboc_stream_fetch(vboc_stream *s, int ticket, double *x, double *u, vboc_stats *stats) {
    if (!s || ticket < 0 || ticket >= s->cap || s->done[ticket] != 2)
        return fail(VBOC_ERR_ARG, "vboc_stream_fetch: not a finished ticket");
    const int t = ticket, Nb = s->N[t];
    if (x) memcpy(x, s->x + (size_t)t * (s->Nmax + 1) * s->nxr, (size_t)(Nb + 1) * s->nxr * sizeof(double));
    if (u) memcpy(u, s->u + (size_t)t * s->Nmax * s->nu, (size_t)Nb * s->nu * sizeof(double));
    if (stats) *stats = s->st[t];
    s->done[t] = 0;
    s->free_slots.push_back(t);
    ++s->solved;
    return 0;
}

int vboc_stream_sim_step(vboc_stream *s, int count, const double *x, const double *u, double T, double *x_next) {
    if (!s || count < 1 || count > s->sim_cap || !x || !u || !x_next)
        return fail(VBOC_ERR_ARG, "vboc_stream_sim_step: bad argument");
    CUDA_OK(cudaSetDevice(s->device));
    const int n = s->n;
    memcpy(s->sx, x, (size_t)count * 2 * n * sizeof(double));
    memcpy(s->su, u, (size_t)count * n * sizeof(double));
    // 32-thread CTAs: they fit beside the resident solve CTAs (registers), so a step never waits for a solve
    const int th = 32, bl = (count + th - 1) / th;
    if (n == 1) sim_kernel<1><<<bl, th, 0, s->sim_stream>>>(count, s->sx, s->su, T, s->sxn);
    if (n == 2) sim_kernel<2><<<bl, th, 0, s->sim_stream>>>(count, s->sx, s->su, T, s->sxn);
    if (n == 3) sim_kernel<3><<<bl, th, 0, s->sim_stream>>>(count, s->sx, s->su, T, s->sxn);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s->sim_stream));
    memcpy(x_next, s->sxn, (size_t)count * 2 * n * sizeof(double));
    return 0;
}

}  // extern "C"
