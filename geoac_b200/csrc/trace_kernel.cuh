// trace_kernel.cuh -- the persistent one-thread-per-ray RK4 kernel, generic over the equation set.
//
// Replaces, for a whole batch of launch angles, the body of the reference's `-prop` loops
// (Code/GeoAc3D_main.cpp:226-304 and siblings) including GeoAc_Propagate_RK4 (Code/GeoAc/GeoAc.Solver.cpp:12-72).
//
// Design (B200-first, not a translation):
//  * ONE resident CTA per SM; the spline table is pulled into shared memory by a single TMA bulk copy
//    (cp.async.bulk + mbarrier) and stays there for the life of the kernel.
//  * a lane owns one ray at a time and advances it by exactly ONE RK4 step (+ its travel-time / absorption
//    segment) per trip round a flat loop, so all lanes of a warp execute the same instruction stream no matter how
//    far along their rays are.  Bounces / arrivals / breaks are handled in a rare divergent tail of the trip.
//  * finished lanes are refilled at the top of the trip: a warp ballot counts the idle lanes, one lane claims that
//    many ray indices from a global counter with a single atomicAdd, the indices are handed out by lane rank.
//  * only three states are ever needed (k-2, k-1, k): y_k and y_{k+1} live in registers, y_{k-1} is parked in a
//    per-thread shared-memory column (conflict-free [eq][thread] layout).  The reference stores 500 000 x EqCnt.
//  * RK4 combination in the reference's association order  y + k1/6 + k2/3 + k3/3 + k4/6  (Solver.cpp:54), built
//    on the fly so no k_i is kept.
#pragma once
#include <type_traits>
#include "core.cuh"
#include "mspline.cuh"

namespace geoac {

struct TraceArgs {
    Grid3D grid;                // range-dependent variants: node tables in global memory (L1/L2 resident working set)
    const double* table;        // global copy of the table (TAB_NARR * n_pad doubles, 16-byte aligned)
    int table_n, table_npad;
    double table_xmin, table_xmax;
    const LaunchConsts* consts; // device
    const double* theta;        // [n_rays]
    const double* phi;
    int64_t n_rays;
    int n_rec;                  // bounces + 1
    double* rec;                // [GEOAC_NFIELDS][n_rays*n_rec]
    int32_t* status;            // [n_rays*n_rec]
    int32_t* n_steps;
    unsigned long long* counter;      // next unclaimed ray
    unsigned long long* total_steps;  // RK4 steps taken (all rays)
};

// Per-launch invariants of the stratified variants (source / ground / reference-level atmosphere samples and the
// Sutherland-Bass reference state), evaluated once with the same spline routines the rays use.
GEOAC_HD void fill_launch_consts_1d(LaunchConsts& L, const Table1D& T, int variant) {
    const bool glob = (variant == GEOAC_GLOBAL);
    int cur = 0;
    double c, u, v, rho, dc, du, dv;
    auto sample = [&](double x) {
        const SegPos sp = seg_locate(T, clampd(x, T.xmin, T.xmax), cur);
        double Tv, dT, ddT, d2;
        spl_f2(T.arr(TAB_T), T.arr(TAB_ST), sp, Tv, dT, ddT);
        spl_f2(T.arr(TAB_U), T.arr(TAB_SU), sp, u, du, d2);
        spl_f2(T.arr(TAB_V), T.arr(TAB_SV), sp, v, dv, d2);
        rho = spl_f(T.arr(TAB_RHO), T.arr(TAB_SRHO), sp);
        c = sqrt(kGamR * Tv);
        dc = kGamR / (2.0 * c) * dT;
    };
    L.ground = glob ? kREarth + L.z_grnd : L.z_grnd;
    sample(glob ? L.src[0] + kREarth : L.src[2]);
    L.c_src = c; L.u_src = u; L.v_src = v; L.rho_src = rho;
    sample(0.0);
    L.c_000 = c;                                   // c(0,0,0) of the 3-D travel time (App. A-4)
    sample(L.ground);
    L.c_gnd = c; L.rho_gnd = rho; L.dc_gnd = dc; L.du_gnd = du; L.dv_gnd = dv;
    // Sutherland-Bass reference state: c, rho at (0,0,z_grnd) (Cartesian, Absorption.cpp:33-34) or at r = z_grnd, i.e.
    // clamped to the lowest level (Global, Absorption.Global.cpp:31-32)
    sample(L.z_grnd);
    suthbass_setup(L, c, rho);
}

struct RecOut { double* rec; int32_t* status; int32_t* n_steps; int64_t n_slots; int n_rec; };

// ---------------------------------------------------------------------------------------------------------------
// One lane = one ray in flight.  advance() performs exactly one RK4 step, the travel-time/absorption bookkeeping of
// that step's segment, and -- rarely -- the end-of-bounce tail (arrival record, reflection, break).  It is the whole
// per-ray state machine of SURVEY 3.4; the kernel below only adds lane refill around it.
// ---------------------------------------------------------------------------------------------------------------
template <class EQ>
struct Lane {
    static constexpr int NEQ = EQ::NEQ;
    double y[NEQ];
    typename EQ::RayC rc;
    typename EQ::Cursor cur;
    int bounce, ksteps;
    int64_t ray;
    double tt_total, att_total, tt_b, att_b, zmax;

    GEOAC_HD void start(const LaunchConsts& L, const typename EQ::Atmo& T, int64_t idx, double theta, double phi) {
        ray = idx; bounce = 0; ksteps = 0; cur = typename EQ::Cursor{};
        tt_total = att_total = tt_b = att_b = zmax = 0.0;
        EQ::init(L, T, theta, phi, rc, y, cur);
    }

    // prev[i * pstride] holds y_{k-1}[i]; returns false when the ray has ended
    GEOAC_HD bool advance(const LaunchConsts& L, const typename EQ::Atmo& T, double* prev, int pstride, const RecOut& o) {
        zmax = fmax(zmax, EQ::altitude(y));                    // running turning height over m < k (App. A-3)
        const double ds = EQ::step_size(L, y);
        double acc[NEQ], p[NEQ], f[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; i++) { acc[i] = y[i]; p[i] = y[i]; }
#pragma unroll 1
        for (int s = 0; s < 4; s++) {
            EQ::rhs(L, T, rc, p, f, cur);
            const double wa = (s == 2) ? 1.0 : 0.5;
            const double wb = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
#pragma unroll
            for (int i = 0; i < NEQ; i++) {
                const double k = ds * f[i];
                acc[i] += k * wb;                              // y + k1/6 + k2/3 + k3/3 + k4/6, left to right
                p[i] = y[i] + k * wa;
            }
        }
        ksteps++;
        double dtt, datt;
        EQ::segment(L, T, rc, y, acc, cur, dtt, datt);
        const bool brk = EQ::left_region(L, rc, acc);            // BreakCheck first (Solver.cpp:57-64)
        const bool gnd = !brk && EQ::below_ground(L, acc);
        const bool lim = !brk && !gnd && (ksteps >= L.step_limit - 1);
        if (L.seg_mode) { if (!(brk || gnd || lim)) { tt_total += dtt; att_total += datt; } }
        else            { tt_b += dtt; att_b += datt; }

        if (!(brk || gnd || lim)) {
#pragma unroll
            for (int i = 0; i < NEQ; i++) { prev[i * pstride] = y[i]; y[i] = acc[i]; }
            return true;
        }
        // ---------------- rare tail: end of a bounce segment ----------------
        const int64_t slot = ray * o.n_rec + bounce;
        if (!gnd) {
            o.status[slot] = brk ? GEOAC_ST_BREAK : GEOAC_ST_LIMIT;
            o.n_steps[slot] = brk ? ksteps : L.step_limit;
            return false;
        }
        if (!L.seg_mode) { tt_total += tt_b; att_total += att_b; tt_b = 0.0; att_b = 0.0; }
        double amp, incl, baz, aux, margin;
        EQ::arrival(L, T, rc, y, acc, tt_total, cur, amp, incl, baz, aux, margin);
#pragma unroll
        for (int i = 0; i < NEQ; i++) o.rec[(int64_t)i * o.n_slots + slot] = acc[i];
        o.rec[(int64_t)GEOAC_F_TRAVELTIME * o.n_slots + slot] = tt_total;
        o.rec[(int64_t)GEOAC_F_ATTEN * o.n_slots + slot] = att_total;
        o.rec[(int64_t)GEOAC_F_TURNHEIGHT * o.n_slots + slot] = zmax;
        o.rec[(int64_t)GEOAC_F_AMPLITUDE * o.n_slots + slot] = amp;
        o.rec[(int64_t)GEOAC_F_INCLINATION * o.n_slots + slot] = incl;
        o.rec[(int64_t)GEOAC_F_BACKAZ * o.n_slots + slot] = baz;
        o.rec[(int64_t)GEOAC_F_AUX * o.n_slots + slot] = aux;
        o.rec[(int64_t)GEOAC_F_MARGIN * o.n_slots + slot] = margin;
        o.status[slot] = GEOAC_ST_ARRIVAL;
        o.n_steps[slot] = ksteps;
        if (bounce >= L.bounces) return false;
        double ym2[NEQ], y0[NEQ];
#pragma unroll
        for (int i = 0; i < NEQ; i++) ym2[i] = prev[i * pstride];
        EQ::reflect(L, T, rc, ym2, y, acc, y0, cur);
#pragma unroll
        for (int i = 0; i < NEQ; i++) y[i] = y0[i];
        bounce++; ksteps = 0;
        if (L.per_bounce_zmax) zmax = 0.0;
        return true;
    }
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// stage `bytes` (multiple of 16) from global to shared with TMA bulk copies issued by thread 0
__device__ __forceinline__ void tma_stage_table(double* dst, const double* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        uint32_t off = 0;
        while (off < bytes) {
            uint32_t chunk = min(bytes - off, 65536u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32((const char*)dst + off)), "l"((const char*)src + off), "r"(chunk), "r"(bar_a)
                         : "memory");
            off += chunk;
        }
    }
    uint32_t done = 0;                                   // everyone waits for phase 0 to complete
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
}

template <class EQ, int BLOCK, bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(BLOCK, 1) trace_kernel(const __grid_constant__ TraceArgs a) {
    constexpr int NEQ = EQ::NEQ;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [LaunchConsts][mbarrier][prev: NEQ*BLOCK doubles][table]
    LaunchConsts* Ls = reinterpret_cast<LaunchConsts*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(LaunchConsts) + 15) / 16) * 16);
    double* prev = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(bar) + 16);
    double* tab_s = prev + NEQ * BLOCK;

    for (int i = threadIdx.x; i < (int)(sizeof(LaunchConsts) / 8); i += BLOCK)
        reinterpret_cast<double*>(Ls)[i] = reinterpret_cast<const double*>(a.consts)[i];
    typename EQ::Atmo T;
    if constexpr (std::is_same<typename EQ::Atmo, Grid3D>::value) {
        T = a.grid;
    } else {
        T.n = a.table_n; T.n_pad = a.table_npad; T.xmin = a.table_xmin; T.xmax = a.table_xmax;
        if (TABLE_IN_SMEM) {
            tma_stage_table(tab_s, a.table, (uint32_t)(TAB_NARR * a.table_npad * sizeof(double)), bar);
            T.base = tab_s;
        } else {
            T.base = a.table;
        }
    }
    __syncthreads();
    const LaunchConsts& L = *Ls;

    const unsigned lane = threadIdx.x & 31;
    RecOut o; o.rec = a.rec; o.status = a.status; o.n_steps = a.n_steps; o.n_rec = a.n_rec; o.n_slots = a.n_rays * a.n_rec;
    Lane<EQ> ln;
    bool have_ray = false, exhausted = false;
    unsigned long long my_steps = 0;

    while (true) {
        // ---------------- refill idle lanes (warp-aggregated claim) ----------------
        const bool want = !have_ray && !exhausted;
        const unsigned wmask = __ballot_sync(0xffffffffu, want);
        if (wmask) {
            unsigned long long base = 0;
            const int leader = __ffs(wmask) - 1;
            if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(wmask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const int64_t idx = (int64_t)(base + __popc(wmask & ((1u << lane) - 1u)));
                if (idx < a.n_rays) { ln.start(L, T, idx, a.theta[idx], a.phi[idx]); have_ray = true; }
                else exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, have_ray)) break;
        if (have_ray) {
            have_ray = ln.advance(L, T, prev + threadIdx.x, BLOCK, o);
            my_steps++;
        }
    }
    for (int off = 16; off > 0; off >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, off);
    if (lane == 0 && my_steps) atomicAdd(a.total_steps, my_steps);      // one atomic per warp
}

#endif  // __CUDACC__
}  // namespace geoac
