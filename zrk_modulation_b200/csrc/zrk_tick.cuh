// zrk_tick.cuh — the batched tick kernel (sm_100a).
//
// One TEAM of threads per scenario, one lane per AirEnv slot (AirEnv.__objects position,
// AirEnv.py:24).  A launch runs n_ticks passes of the loop body of Manager.run_simulation
// (Manager.py:116-140) with the scenario's state resident in shared memory; it is read from /
// written back to HBM once per launch.
//
// Arithmetic is fp64 and written so that every result that reaches state or a decision is
// the same IEEE operation sequence the reference (numpy) performs: no FMA contraction (the
// file is compiled with -fmad=false), explicit fma() only where OpenBLAS ddot fuses
// (np.dot / np.linalg.norm on 3-vectors), sqrt() and `/` correctly rounded.  The two fp32 /
// squared-distance pre-filters below only ever decide cases that are provably far from a
// threshold and defer everything else to the exact path, so decisions are unchanged.
//
// The file also compiles under tests/emu (a fiber SIMT emulator, -DZRK_EMU) so that the
// host-side logic tests can run the identical source without a GPU.  The product is the nvcc
// build only.
#ifndef ZRK_TICK_CUH
#define ZRK_TICK_CUH

#include "zrk_layout.h"

#ifdef ZRK_EMU
#include "cuda_emu.h"
#define ZRK_SMEM_DECL unsigned char *zrk_smem = zrk_emu::dyn_smem()
#define ZRK_LD_EARLY(p) (*(p))
#define ZRK_ST_GLOBAL(p, v) (*(p) = (v))
#define ZRK_ST_MASK(p, v) (*(p) = (v))
#define ZRK_NOINLINE __attribute__((noinline))
#else
#include <cuda_runtime.h>
#define ZRK_SMEM_DECL extern __shared__ __align__(16) unsigned char zrk_smem[]
// a read-only global load the compiler must issue where it is written (not sunk to its first use),
// so that its latency overlaps the work in between
__device__ __forceinline__ double zrk_ld_early(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
#define ZRK_LD_EARLY(p) zrk_ld_early(p)
// a store the compiler knows to be to global memory (the pointer came out of a structure: it would be a generic one)
__device__ __forceinline__ void zrk_st_global(unsigned *p, unsigned v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#define ZRK_ST_GLOBAL(p, v) zrk_st_global(p, v)
// the same for a word nothing in the storing code reads back (a mask row): values the code holds in registers stay there
__device__ __forceinline__ void zrk_st_mask(unsigned *p, unsigned v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v)); }
#define ZRK_ST_MASK(p, v) zrk_st_mask(p, v)
#define ZRK_NOINLINE __noinline__
#endif

#include <math.h>

// the base of the dynamic shared memory (for code that is handed an offset instead of a pointer, so that its accesses
// are shared-memory instructions rather than generic ones)
__device__ __forceinline__ unsigned char *zrk_smem_base() {
    ZRK_SMEM_DECL;
    return zrk_smem;
}

// branch hints: the rare side is moved out of the straight-line code (the hot loop has to stay
// within the instruction cache)
#if defined(__GNUC__) || defined(__CUDACC__)
#define ZRK_RARE(x) __builtin_expect(!!(x), 0)
#define ZRK_OFTEN(x) __builtin_expect(!!(x), 1)
#else
#define ZRK_RARE(x) (x)
#define ZRK_OFTEN(x) (x)
#endif

namespace zrk {

#define ZRK_FULL 0xffffffffu
#define ZRK_IMAX 0x7fffffff

// ------------------------------------------------------------------ arithmetic helpers

// np.dot on 3-vectors: OpenBLAS ddot scalar tail, an FMA chain (oracle/zrk_oracle.c:dot3).
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return fma(az, bz, fma(ay, by, ax * bx));
}
// Python float % for a positive modulus.
__device__ __forceinline__ double pymod(double a, double m) {
    // (0 < a < 2 m, what a sweeping sector's angle + step always is: fmod is a or a - m, both exact)
    if (a > 0.0 && a < m) return a;
    if (a >= m && a < m + m) return a - m;
    double r = fmod(a, m);
    if (r != 0.0) {
        if (r < 0.0) r += m;
    } else {
        r = 0.0;
    }
    return r;
}
__device__ __forceinline__ double to_seconds(long long ms) { return (double)ms / 1000.0; }
// to_seconds(tick * dt) from the table the host fills with exactly that expression
__device__ ZRK_NOINLINE double tick_seconds_beyond(long long tick, int dt_ms) { return to_seconds(tick * dt_ms); }
__device__ __forceinline__ double tick_seconds(const Params &P, int tick) {
    return ZRK_OFTEN(tick >= 0 && tick < P.tsec_n) ? ZRK_LD_EARLY(&P.tsec[tick]) : tick_seconds_beyond(tick, P.dt_ms);
}

#define ZRK_RAD2DEG 57.29577951308232 /* 180.0 / pi, as np.degrees multiplies */

// fp32 square root for the conservative pre-filters only (never for a value that reaches state)
__device__ __forceinline__ float sqrt_fast(float x) {  // relative error <= 2^-23 (sqrt.approx.f32)
#ifdef ZRK_EMU
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// SectorRadar.find_visible_objects for one object (Radar.py:53-71), exact path.
__device__ ZRK_NOINLINE bool radar_sees_exact(double dx, double dy, double dz, double max_dist, double az_lo,
                                              double az_hi, double el_lo, double el_hi) {
    double distance = sqrt(dot3(dx, dy, dz, dx, dy, dz));
    if (distance > max_dist) return false;
    double az = pymod(atan2(dy, dx) * ZRK_RAD2DEG, 360.0);
    double el = pymod(asin(dz / distance) * ZRK_RAD2DEG, 180.0);
    return az_lo <= az && az <= az_hi && el_lo <= el && el <= el_hi;
}

// Visibility with an fp32 pre-filter.  The fp32 angles carry an absolute error far below
// ANG_MARGIN degrees; a case is decided in fp32 only if every comparison (including the
// 0/360 and 0/180 wrap points of the `%`) is at least ANG_MARGIN away from equality and the
// squared range is outside the bracket around max_distance^2.  Everything else (and every
// NaN) goes to radar_sees_exact.
// atan2 in degrees, |error| < 1e-3 degrees (Abramowitz-Stegun 4.4.49 polynomial, fp32).
__device__ __forceinline__ float atan2_deg_approx(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = __fdividef(mn, mx);  // 0/0 -> NaN -> every comparison false -> exact path
    const float q = a * a;
    // (explicit fmaf: the file is compiled without contraction for the sake of the fp64 path)
    float r = fmaf(fmaf(fmaf(fmaf(0.0208351f, q, -0.085133f), q, 0.180141f), q, -0.3302995f), q, 0.999866f) * a;
    if (ay > ax) r = 1.57079637f - r;
    if (x < 0.f) r = 3.14159274f - r;
    if (y < 0.f) r = -r;
    return r * 57.29577951f;
}

// The pre-filter itself: 1 = surely visible, 0 = surely not, -1 = not decided here.  p8 / p9: squared-range bracket
// around max_distance^2 (radar_par[8], [9]).
__device__ __forceinline__ int radar_fast(double dx, double dy, double dz, double p8, double p9, double az_lo, double az_hi,
                                          double el_lo, double el_hi) {
    const float M = 1e-2f;  // decision margin in degrees, 10x the approximation error
    const double d2 = dot3(dx, dy, dz, dx, dy, dz);
    if (d2 > p9) return 0;  // certainly farther than max_distance
    bool sure_in = d2 > 0.0 && d2 <= p8, sure_out = false, ok = true;  // (an object AT the radar: arcsin(0/0) is NaN, exact path)
    const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
    // azimuth % 360 lies in [0, 360]: a sector covering all of it needs no angle at all
    if (!(az_lo <= 0.0 && az_hi >= 360.0)) {
        float az = atan2_deg_approx(fy, fx);
        if (az < 0.f) az += 360.f;
        const float a0 = (float)az_lo, a1 = (float)az_hi;
        ok = ok && az > M && az < 360.f - M && fabsf(a0) < 1e4f && fabsf(a1) < 1e4f;  // away from the % wrap
        sure_in = sure_in && az > a0 + M && az < a1 - M;
        sure_out = sure_out || az < a0 - M || az > a1 + M;
    }
    if (el_lo == 0.0 && el_hi == 90.0) {
        // the upper hemisphere: degrees(arcsin(dz / d)) % 180 lies in [0, 90] iff dz >= 0 (a negative angle maps
        // beyond 90).  Decided here unless dz is within 1e-3 of zero relative to the horizontal range
        // (elevation within 0.06 deg of the horizon) or straight overhead (within 0.06 deg of 90).
        const float hz = sqrt_fast(fmaf(fy, fy, fx * fx));
        ok = ok && fabsf(fz) > 1e-3f * hz && hz > 1e-3f * fabsf(fz);
        sure_in = sure_in && fz > 0.f;
        sure_out = sure_out || fz < 0.f;
    } else if (!(el_lo <= 0.0 && el_hi >= 180.0)) {
        float el = atan2_deg_approx(fz, sqrt_fast(fmaf(fy, fy, fx * fx)));  // == asin(dz / distance), well conditioned
        if (el < 0.f) el += 180.f;
        const float e0 = (float)el_lo, e1 = (float)el_hi;
        ok = ok && el > M && el < 180.f - M && fabsf(e0) < 1e4f && fabsf(e1) < 1e4f;
        sure_in = sure_in && el > e0 + M && el < e1 - M;
        sure_out = sure_out || el < e0 - M || el > e1 + M;
    }
    if (ZRK_OFTEN(ok)) {
        if (sure_out) return 0;
        if (ZRK_OFTEN(sure_in)) return 1;
    }
    return -1;
}
template <bool FAST>
__device__ __forceinline__ bool radar_sees(double dx, double dy, double dz, const double *par, double az_lo,
                                           double az_hi, double el_lo, double el_hi) {
    if (FAST) {
        const int v = radar_fast(dx, dy, dz, par[8], par[9], az_lo, az_hi, el_lo, el_hi);
        if (ZRK_OFTEN(v >= 0)) return v != 0;
    }
    return radar_sees_exact(dx, dy, dz, par[2], az_lo, az_hi, el_lo, el_hi);
}

// SectorRadar.move_to_next_sector_circular (Radar.py:96-117); q = radar_par of the radar.
__device__ __forceinline__ void radar_move(const double *q, double &caz, double &cel) {
    const int mode = (int)q[7];
    if (mode == ZRK_SCAN_HORIZONTAL) {
        if (caz + q[3] < 360)
            caz = pymod(caz + q[5], 360.0);
        else
            caz = q[1];  // sic: elevation_start (Radar.py:105)
        if (caz < q[5]) {
            if (cel + q[6] < 90)
                cel = pymod(cel + q[6], 90.0);
            else
                cel = q[1];
        }
    } else if (mode == ZRK_SCAN_VERTICAL) {
        cel = pymod(cel + q[6], 90.0);
        if (cel < q[6]) caz = pymod(caz + q[5], 360.0);
    }
}

// ------------------------------------------------------------------ bulk (TMA) staging of the state
//
// The scenario's state blob moves between HBM and shared memory as ONE bulk asynchronous copy
// per direction (cp.async.bulk, the non-tensor TMA path: UBLKCP in SASS), issued by one lane and
// tracked by an mbarrier, instead of a loop of per-thread 16-byte loads and stores.
#ifdef ZRK_EMU
__device__ __forceinline__ void stage_in(unsigned char *sm, const unsigned char *g_dyn, int dyn_bytes,
                                         const unsigned char *g_cst, int cst_bytes, void *, bool leader) {
    if (leader) {
        memcpy(sm, g_dyn, (size_t)dyn_bytes);
        memcpy(sm + dyn_bytes, g_cst, (size_t)cst_bytes);
    }
}
__device__ __forceinline__ void stage_wait(void *) {}
__device__ __forceinline__ void stage_fence() {}
__device__ __forceinline__ void stage_out(unsigned char *g_dyn, const unsigned char *sm, int dyn_bytes, bool leader) {
    if (leader) memcpy(g_dyn, sm, (size_t)dyn_bytes);
}
#else
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// The team's leader arms the mbarrier with the byte count and issues the two bulk loads.
__device__ __forceinline__ void stage_in(unsigned char *sm, const unsigned char *g_dyn, int dyn_bytes,
                                         const unsigned char *g_cst, int cst_bytes, void *mbar, bool leader) {
    if (leader) {
        const unsigned bar = smem_addr(mbar);
        asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(dyn_bytes + cst_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sm)),
                     "l"(g_dyn), "r"(dyn_bytes), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(sm + dyn_bytes)),
                     "l"(g_cst), "r"(cst_bytes), "r"(bar)
                     : "memory");
    }
}
// Every lane waits for phase 0 of the barrier; bounded, so that a mistake traps instead of hanging.
__device__ __forceinline__ void stage_wait(void *mbar) {
    const unsigned bar = smem_addr(mbar);
    for (int spin = 0; spin < (1 << 24); spin++) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(0)
                     : "memory");
        if (done) return;
    }
    __trap();
}
// every lane that wrote the blob through the generic proxy fences before the team's last barrier ...
__device__ __forceinline__ void stage_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// ... and the leader then issues the one bulk store
__device__ __forceinline__ void stage_out(unsigned char *g_dyn, const unsigned char *sm, int dyn_bytes, bool leader) {
    if (leader) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_dyn), "r"(smem_addr(sm)), "r"(dyn_bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
#endif

// ------------------------------------------------------------------ team primitives

// A TEAM of 32 is a warp (several per CTA, warp-synchronous); larger teams are exactly one CTA
// and combine per-warp partial results through two alternating rows of shared scratch, so that a
// collective costs ONE __syncthreads (the row written by call k+2 was last read before the
// barrier of call k+1).
template <int TEAM, int TPC>
struct Team {
    static_assert(TEAM == 32 || TPC == 1, "teams larger than a warp own their CTA");
    static constexpr int WARPS = TEAM / 32;
    int lane;   // 0..TEAM-1 within the scenario's team
    int tic;    // team index inside the CTA
    int *wscr;  // 64 ints of shared scratch (TEAM > 32 only)
    int flip;   // which scratch row the next collective uses (uniform across the team)

    __device__ __forceinline__ void sync() const {
        if (TEAM == 32) {
            __syncwarp();
        } else {
            __syncthreads();
        }
    }
    __device__ __forceinline__ int *row() {
        int *r = wscr + (flip ? 32 : 0);
        flip ^= 1;
        return r;
    }
    // per-warp values v (one per warp, held by every lane of that warp) -> exclusive prefix over
    // warps for this thread's warp, and the total
    __device__ __forceinline__ int across_warps(int v, int &total) {
        const int l32 = lane & 31, w = lane >> 5;
        int *r = row();
        if (l32 == 0) r[w] = v;
        __syncthreads();
        const int x = l32 < WARPS ? r[l32] : 0;
        int inc = x;
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const int n = __shfl_up_sync(ZRK_FULL, inc, d);
            if (l32 >= d) inc += n;
        }
        total = __shfl_sync(ZRK_FULL, inc, WARPS - 1);
        return __shfl_sync(ZRK_FULL, inc - x, w);
    }
    // exclusive prefix count of a predicate in lane order; total returned to every lane
    __device__ __forceinline__ int rank(bool p, int &total) {
        const unsigned bal = __ballot_sync(ZRK_FULL, p);
        const int below = __popc(bal & ((1u << (lane & 31)) - 1u));
        if (TEAM == 32) {
            total = __popc(bal);
            return below;
        }
        return across_warps(__popc(bal), total) + below;
    }
    // Several ranks for the price of one barrier: per-warp counts packed into one word (field k at
    // bit SHk; the caller guarantees that every team-wide total fits its field), combined like
    // across_warps.  `orin` is OR-reduced over the team by the same barrier (__syncthreads_or).
    __device__ __forceinline__ unsigned across_warps_or(unsigned v, unsigned &total, bool orin, bool &orout) {
        const int l32 = lane & 31, w = lane >> 5;
        int *r = row();
        if (l32 == 0) r[w] = (int)v;
        orout = __syncthreads_or(orin) != 0;
        const unsigned x = l32 < WARPS ? (unsigned)r[l32] : 0u;
        unsigned inc = x;
#pragma unroll
        for (int d = 1; d < WARPS; d <<= 1) {
            const unsigned n = __shfl_up_sync(ZRK_FULL, inc, d);
            if (l32 >= d) inc += n;
        }
        total = __shfl_sync(ZRK_FULL, inc, WARPS - 1);
        return __shfl_sync(ZRK_FULL, inc - x, w);
    }
    // rank of p and the team-wide OR of `orin`
    __device__ __forceinline__ int rank_or(bool p, int &total, bool orin, bool &orout) {
        const unsigned bal = __ballot_sync(ZRK_FULL, p);
        const int below = __popc(bal & ((1u << (lane & 31)) - 1u));
        if (TEAM == 32) {
            total = __popc(bal);
            orout = __ballot_sync(ZRK_FULL, orin) != 0;
            return below;
        }
        unsigned tot;
        const unsigned ex = across_warps_or((unsigned)__popc(bal), tot, orin, orout);
        total = (int)tot;
        return (int)ex + below;
    }
    // ranks of p0 and p1 (totals < 65536 each) and the team-wide OR of `orin`
    __device__ __forceinline__ void rank2(bool p0, bool p1, int &r0, int &r1, int &t0, int &t1, bool orin, bool &orout) {
        const unsigned b0 = __ballot_sync(ZRK_FULL, p0), b1 = __ballot_sync(ZRK_FULL, p1);
        const unsigned lt = (1u << (lane & 31)) - 1u;
        r0 = __popc(b0 & lt);
        r1 = __popc(b1 & lt);
        if (TEAM == 32) {
            t0 = __popc(b0);
            t1 = __popc(b1);
            orout = __ballot_sync(ZRK_FULL, orin) != 0;
            return;
        }
        unsigned tot;
        const unsigned ex = across_warps_or((unsigned)__popc(b0) | ((unsigned)__popc(b1) << 16), tot, orin, orout);
        r0 += (int)(ex & 0xffffu);
        r1 += (int)(ex >> 16);
        t0 = (int)(tot & 0xffffu);
        t1 = (int)(tot >> 16);
    }
    // ranks of p0 (total < 1024), p1 and p2 (totals <= 1024 each) and the team-wide OR of `orin`
    __device__ __forceinline__ void rank3(bool p0, bool p1, bool p2, int &r0, int &r1, int &r2, int &t0, int &t1, int &t2,
                                          bool orin, bool &orout) {
        const unsigned b0 = __ballot_sync(ZRK_FULL, p0), b1 = __ballot_sync(ZRK_FULL, p1), b2 = __ballot_sync(ZRK_FULL, p2);
        const unsigned lt = (1u << (lane & 31)) - 1u;
        r0 = __popc(b0 & lt);
        r1 = __popc(b1 & lt);
        r2 = __popc(b2 & lt);
        if (TEAM == 32) {
            t0 = __popc(b0);
            t1 = __popc(b1);
            t2 = __popc(b2);
            orout = __ballot_sync(ZRK_FULL, orin) != 0;
            return;
        }
        unsigned tot;
        const unsigned ex =
            across_warps_or((unsigned)__popc(b0) | ((unsigned)__popc(b1) << 10) | ((unsigned)__popc(b2) << 21), tot, orin, orout);
        r0 += (int)(ex & 0x3ffu);
        r1 += (int)((ex >> 10) & 0x7ffu);
        r2 += (int)(ex >> 21);
        t0 = (int)(tot & 0x3ffu);
        t1 = (int)((tot >> 10) & 0x7ffu);
        t2 = (int)(tot >> 21);
    }
    __device__ __forceinline__ int min(int v) {
        v = __reduce_min_sync(ZRK_FULL, v);
        if (TEAM == 32) return v;
        const int l32 = lane & 31;
        int *r = row();
        if (l32 == 0) r[lane >> 5] = v;
        __syncthreads();
        return __reduce_min_sync(ZRK_FULL, l32 < WARPS ? r[l32] : ZRK_IMAX);
    }
    __device__ __forceinline__ bool any(bool p) const {
        if (TEAM == 32) return __ballot_sync(ZRK_FULL, p) != 0;
        return __syncthreads_or(p) != 0;
    }
};

// ------------------------------------------------------------------ record log

__device__ __forceinline__ void put_rec(const Params &P, zrk_msg *base, int idx, int tick, int type, int flag,
                                        int sender, int receiver, int a, int b, double x, double y, double z) {
    if (idx < P.log_cap) {  // one 48-byte record = three 16-byte stores
        uint4 *r = (uint4 *)(base + idx);
        const unsigned w1 = (unsigned)type | ((type == ZRK_MSG_MISSILE_COUNT_REQUEST ? 3u : 1u) << 8) | ((unsigned)(flag & 0xff) << 16);
        const unsigned long long xb = (unsigned long long)__double_as_longlong(x), yb = (unsigned long long)__double_as_longlong(y),
                                 zb = (unsigned long long)__double_as_longlong(z);
        r[0] = make_uint4((unsigned)tick, w1, (unsigned)sender, (unsigned)receiver);
        r[1] = make_uint4((unsigned)a, (unsigned)b, (unsigned)xb, (unsigned)(xb >> 32));
        r[2] = make_uint4((unsigned)yb, (unsigned)(yb >> 32), (unsigned)zb, (unsigned)(zb >> 32));
    }
}
__device__ __forceinline__ bool logged(const Params &P, int type) { return (P.log_types >> type) & 1u; }

// ------------------------------------------------------------------ Missile._calculate_trajectory_params

// Missile.py:35-102.  Returns -1 on success (V filled) or a zrk_cancel_reason.
__device__ ZRK_NOINLINE int missile_solve(const double *mp, const double *tp, const double *vt, double v0,
                                          double v0sq, double period, double *V) {
    double d0 = tp[0] - mp[0], d1 = tp[1] - mp[1], d2 = tp[2] - mp[2];
    double a = dot3(vt[0], vt[1], vt[2], vt[0], vt[1], vt[2]) - v0sq;
    double b = 2.0 * dot3(d0, d1, d2, vt[0], vt[1], vt[2]);
    double c = dot3(d0, d1, d2, d0, d1, d2);
    double t;
    if (fabs(a) < 1e-6) {
        if (fabs(b) < 1e-6) return ZRK_CANCEL_PARALLEL;
        t = -c / b;
        if (t <= 0) return ZRK_CANCEL_PAST;
    } else {
        double disc = b * b - 4.0 * a * c;
        if (disc < 0) return ZRK_CANCEL_NO_REAL_ROOT;
        double sq = sqrt(disc);
        double t1 = (-b + sq) / (2.0 * a);
        double t2 = (-b - sq) / (2.0 * a);
        bool p1 = t1 > 0, p2 = t2 > 0;
        if (!p1 && !p2) return ZRK_CANCEL_NOT_POSITIVE;
        if (p1 && p2)
            t = (t2 < t1) ? t2 : t1;
        else
            t = p1 ? t1 : t2;
    }
    if (t > period) return ZRK_CANCEL_TOO_FAR;
    double V0 = d0 / t + vt[0], V1 = d1 / t + vt[1], V2 = d2 / t + vt[2];
    double n = sqrt(dot3(V0, V1, V2, V0, V1, V2));
    V[0] = V0 / n * v0;
    V[1] = V1 / n * v0;
    V[2] = V2 / n * v0;
    return -1;
}

// ------------------------------------------------------------------ shared-memory views

// Hot arrays at compile-time offsets from the team's shared-memory base.
template <int TEAM>
struct Hot {
    static constexpr HotOff O = hot_off(TEAM);
    unsigned char *sm;
    __device__ __forceinline__ int *hdr() const { return (int *)(sm + O.hdr); }
    __device__ __forceinline__ unsigned *counts() const { return (unsigned *)(sm + O.counts); }
    __device__ __forceinline__ double2 *pxy() const { return (double2 *)(sm + O.pxy); }
    __device__ __forceinline__ double2 *pzs() const { return (double2 *)(sm + O.pzs); }
    __device__ __forceinline__ double2 *qxy() const { return (double2 *)(sm + O.qxy); }
    __device__ __forceinline__ double2 *qzw() const { return (double2 *)(sm + O.qzw); }
    __device__ __forceinline__ double2 *txy() const { return (double2 *)(sm + O.txy); }
    __device__ __forceinline__ double2 *tzt() const { return (double2 *)(sm + O.tzt); }
    __device__ __forceinline__ double2 *vxy() const { return (double2 *)(sm + O.vxy); }
    __device__ __forceinline__ int *sid() const { return (int *)(sm + O.sid); }
    __device__ __forceinline__ int *sflags() const { return (int *)(sm + O.sflags); }
    __device__ __forceinline__ int *srf() const { return (int *)(sm + O.srf); }
    __device__ __forceinline__ int *own() const { return (int *)(sm + O.own); }
    __device__ __forceinline__ int *te_key() const { return (int *)(sm + O.te_key); }
    __device__ __forceinline__ int *te_ref() const { return (int *)(sm + O.te_ref); }
    __device__ __forceinline__ int *te_upd() const { return (int *)(sm + O.te_upd); }
    __device__ __forceinline__ int *te_follow() const { return (int *)(sm + O.te_follow); }
    __device__ __forceinline__ int *te_mid() const { return (int *)(sm + O.te_mid); }
    __device__ __forceinline__ double *te_upds() const { return (double *)(sm + O.te_upds); }
    __device__ __forceinline__ int *c_in() const { return (int *)(sm + O.c_in); }
    __device__ __forceinline__ int *c_out() const { return (int *)(sm + O.c_out); }
    __device__ __forceinline__ int *c_refc() const { return (int *)(sm + O.c_refc); }
    __device__ __forceinline__ int *c_flag() const { return (int *)(sm + O.counts) + CW_FLAG; }
    __device__ __forceinline__ int *c_removed() const { return (int *)(sm + O.counts) + CW_REMOVED; }
};

// Cold arrays (run-time offsets), resolved on use.
#define ZRK_COLD_I(name) ((int *)(sm + o.name))
#define ZRK_COLD_D(name) ((double *)(sm + o.name))
#define ZRK_CST_I(name) ((const int *)(sm + o.dyn_bytes + o.name))
#define ZRK_CST_D(name) ((const double *)(sm + o.dyn_bytes + o.name))

// ------------------------------------------------------------------ CCP.link_object

// calc_range (CCP.py:176-188): d_t = cur_time - upd_time has been computed per entry.
__device__ __forceinline__ void calc_range(double speed, double d_t, double range_step, double &lo, double &hi) {
    lo = speed * (d_t - range_step);
    hi = speed * (d_t + range_step);
    lo = lo > 0 ? lo : 0;  // max(0, x)
    hi = hi > 0 ? hi : 0;
}

#define ZRK_FAR 1e300 /* link position of an entry link_object skips (upd_time == now) */

// CombatControlPoint.link_object (CCP.py:171-219) for one detection, exact: sqrt per pair.
// Entries are addressed by a unified index: [0, n_te) = _target_dict, TEAM + [0, n_me) =
// _missile_dict; lxy/lzd hold, per entry, the position the reference measures to
// (prev_pos, or pos for a missile whose prev_pos is None) and d_t.
template <int TEAM>
__device__ ZRK_NOINLINE void link_exact(const double2 *lxy, const double2 *lzd, int n_te, int n_me, double range_step,
                                        double px, double py, double pz, double speed, int &cls, int &cand) {
    double curr = INFINITY;
    cls = 0;
    cand = -1;
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass ? n_me : n_te, base = pass ? TEAM : 0;
        for (int e = 0; e < n; e++) {
            const double2 a = lxy[base + e], c = lzd[base + e];
            if (a.x >= ZRK_FAR) continue;
            double ex = a.x - px, ey = a.y - py, ez = c.x - pz;
            double diff = sqrt(dot3(ex, ey, ez, ex, ey, ez));
            double lo, hi;
            calc_range(speed, c.y, range_step, lo, hi);
            if (diff < curr && lo <= diff && diff <= hi) {
                curr = diff;
                cls = pass + 1;
                cand = e;
            }
        }
    }
}

// State of the squared-domain scan of one detection.
struct LinkAcc {
    double best, best_lo;  // smallest accepted d2 and best * (1 - EPS)
    int ue;                // unified index of the accepted entry, -1 = none
    bool unsure;           // a comparison was too close to call: redo with link_exact
};

// Full evaluation of one candidate entry in the squared-distance domain.  sqrt is monotone, so
// "diff < curr" and "lo <= diff <= hi" are decided from d2 unless d2 is within a few ulps of
// best / lo^2 / hi^2 (or NaN), which sets `unsure`.  Exact ties (d2 == best) go to the entry
// that comes first in the reference's scan order.
__device__ __forceinline__ void link_eval(LinkAcc &A, int ue, double d2, double d_t, double speed, double range_step) {
    const double EPS = 4e-15;
    double lo, hi;
    calc_range(speed, d_t, range_step, lo, hi);
    const double l2 = lo * lo, h2 = hi * hi;
    // certainly outside [lo, hi]; also drops the entries link_object skips (upd_time == now),
    // whose link position ZRK_FAR makes d2 infinite
    if (d2 > h2 * (1.0 + EPS) || d2 < l2 * (1.0 - EPS)) return;
    const bool in_sure = d2 >= l2 * (1.0 + EPS) && d2 <= h2 * (1.0 - EPS);
    if (d2 > A.best) return;
    if (d2 == A.best) {
        if (A.ue >= 0 && ue > A.ue) return;  // the earlier entry keeps an exact tie
        if (in_sure) {
            A.ue = ue;
            return;
        }
    } else if (in_sure && d2 < A.best_lo) {
        A.best = d2;
        A.best_lo = d2 * (1.0 - EPS);
        A.ue = ue;
        return;
    }
    A.unsure = true;  // includes NaN: every comparison above was false
}

__device__ __forceinline__ double link_d2(const double2 *lxy, const double2 *lzd, int ue, double px, double py,
                                          double pz) {
    const double2 a = lxy[ue];
    const double ex = a.x - px, ey = a.y - py, ez = lzd[ue].x - pz;
    return dot3(ex, ey, ez, ex, ey, ez);
}

// fp32 pre-filter of the scan.  Link positions and the detection are rounded to fp32.  For
// |coordinate| <= C = 2^20 m each component of the fp32 difference vector is off by at most
// 2 C 2^-24 (two roundings to fp32) plus a relative 2^-24, so the fp32 distance differs from the
// fp64 one by less than sqrt(3) * 2^-3 = 0.22 m plus 2e-7 relative: a pair whose fp32 squared
// distance exceeds (sqrt(best) + 0.5 m)^2 * (1 + 1e-6) cannot beat or tie the best (the margin
// also absorbs the fp32 evaluation of that bound itself).  A detection
// or an entry outside that range sends the detection (the whole team, for an entry) to link_exact.
#define ZRK_F32_COORD_MAX 1048576.0
#define ZRK_F32_MARGIN 0.5
#define ZRK_F32_ALL 3.0e38f /* threshold "everything real": pads and skipped entries measure inf */
__device__ __forceinline__ float link_threshold(double best) {
    if (!(best < 1e30)) return ZRK_F32_ALL;
    // an upper bound of (sqrt(best) + margin)^2 in fp32: every rounding below is covered by the factors
    const float r = sqrt_fast((float)best) * 1.000001f + (float)ZRK_F32_MARGIN;
    return r * r * 1.000001f;
}

// The fp32 images live in blocks of four entries, x[4] y[4] z[4] (48 bytes), so that one block is
// three 16-byte loads and the arithmetic runs on the packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2):
// two pairs per instruction, the same rounding as the scalar ex*ex -> fma -> fma chain.
// Teams larger than a warp keep one float4 per entry instead (x, y, z, next entry of the hash-grid
// chain): their scans mostly walk the grid, one 16-byte load per visited entry.
template <int TEAM>
__device__ __forceinline__ float &lf_x(float *lf, int ue) {
    return TEAM > 32 ? lf[4 * ue] : lf[(ue >> 2) * 12 + (ue & 3)];
}
template <int TEAM>
__device__ __forceinline__ void lf_store(float *lf, int ue, float x, float y, float z) {
    if (TEAM > 32) {
        reinterpret_cast<float4 *>(lf)[ue] = make_float4(x, y, z, 0.f);
    } else {
        float *q = &lf_x<TEAM>(lf, ue);
        q[0] = x;
        q[4] = y;
        q[8] = z;
    }
}
__device__ __forceinline__ float link_d2f(float4 a, float fx, float fy, float fz) {
    const float ex = a.x - fx, ey = a.y - fy, ez = a.z - fz;
    return fmaf(ez, ez, fmaf(ey, ey, ex * ex));
}

#ifdef ZRK_EMU
struct F2 {
    float lo, hi;
};
static inline F2 f2_dup(float a) { return F2{a, a}; }
static inline F2 f2_add(F2 a, F2 b) { return F2{a.lo + b.lo, a.hi + b.hi}; }
static inline F2 f2_mul(F2 a, F2 b) { return F2{a.lo * b.lo, a.hi * b.hi}; }
static inline F2 f2_fma(F2 a, F2 b, F2 c) { return F2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
static inline void f2_load4(const float *p, F2 &a, F2 &b) {
    a = F2{p[0], p[1]};
    b = F2{p[2], p[3]};
}
static inline void f2_split(F2 a, float &lo, float &hi) {
    lo = a.lo;
    hi = a.hi;
}
#else
typedef unsigned long long F2;
__device__ __forceinline__ F2 f2_dup(float a) {
    F2 r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void f2_load4(const float *p, F2 &a, F2 &b) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void f2_split(F2 a, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
#endif

// fp32 squared distances of the four entries of block `blk` to the detection (nf* = -coordinate, twice)
__device__ __forceinline__ void link_block(const float *lf, int blk, F2 nfx, F2 nfy, F2 nfz, float d[4]) {
    const float *q = lf + blk * 12;
    F2 x0, x1, y0, y1, z0, z1;
    f2_load4(q, x0, x1);
    f2_load4(q + 4, y0, y1);
    f2_load4(q + 8, z0, z1);
    x0 = f2_add(x0, nfx);
    x1 = f2_add(x1, nfx);
    y0 = f2_add(y0, nfy);
    y1 = f2_add(y1, nfy);
    z0 = f2_add(z0, nfz);
    z1 = f2_add(z1, nfz);
    const F2 s0 = f2_fma(z0, z0, f2_fma(y0, y0, f2_mul(x0, x0)));
    const F2 s1 = f2_fma(z1, z1, f2_fma(y1, y1, f2_mul(x1, x1)));
    f2_split(s0, d[0], d[1]);
    f2_split(s1, d[2], d[3]);
}

// Hash grid over the entries' link positions (x, y of the fp32 image; cells of ZRK_CELL metres,
// TEAM buckets, chained through the w component of the images).  A detection whose acceptance
// radius sqrt(thr) is at most one cell wide only has to look at the buckets of the <= 2 x 2 cells
// its radius touches: every entry within that radius is in one of them (hash collisions only add
// harmless extra tests).
#define ZRK_CELL 512.0f
__device__ __forceinline__ int grid_cell(float v) { return (int)floorf(v * (1.0f / ZRK_CELL)); }
template <int TEAM>
__device__ __forceinline__ int grid_bucket(int ix, int iy) {
    return (int)(((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u)) & (TEAM - 1);
}

// Squared-domain scan of one detection.  The caller has checked that the detection and every
// entry image are inside the fp32-safe range (otherwise it runs link_exact).
template <int TEAM>
__device__ __forceinline__ void link_fast(const int *ghead, bool grid_ok, const float *lf, const double2 *lxy, const double2 *lzd, int n_te, int n_me,
                                          int seed, double range_step, double px, double py, double pz, double speed,
                                          int &cls, int &cand, bool &unsure) {
    LinkAcc A;
    A.best = INFINITY;
    A.best_lo = INFINITY;
    A.ue = -1;
    A.unsure = false;
    // Seed with the entry that tracked this slot last tick: it is almost always the nearest, and
    // with best = its d2 nearly every other pair fails the single fp32 compare below.
    if (seed >= 0) link_eval(A, seed, link_d2(lxy, lzd, seed, px, py, pz), lzd[seed].y, speed, range_step);
    const float fx = (float)px, fy = (float)py, fz = (float)pz;
    float thr = link_threshold(A.best);
    if (TEAM > 32) {
        const float4 *lf4 = (const float4 *)lf;
        if (grid_ok && thr <= ZRK_CELL * ZRK_CELL) {
            // the common case: a seeded detection only visits the few entries near it
            const float r = sqrtf(thr);
            const int ix0 = grid_cell(fx - r), ix1 = grid_cell(fx + r), iy0 = grid_cell(fy - r), iy1 = grid_cell(fy + r);
            for (int ix = ix0; ix <= ix1; ix++)
                for (int iy = iy0; iy <= iy1; iy++)
                    for (int u = ghead[grid_bucket<TEAM>(ix, iy)]; u >= 0;) {
                        const float4 a = lf4[u];
                        if (u != seed && !(link_d2f(a, fx, fy, fz) > thr)) {
                            const double dk = link_d2(lxy, lzd, u, px, py, pz);
                            if (!(dk > A.best)) {
                                link_eval(A, u, dk, lzd[u].y, speed, range_step);
                                thr = link_threshold(A.best);
                            }
                        }
                        u = __float_as_int(a.w);
                    }
        } else {
            for (int pass = 0; pass < 2; pass++) {
                const int n = pass ? n_me : n_te, base = pass ? TEAM : 0;
                int e = 0;
                // four pairs at a time; the common case is that none of them can beat the best
                for (; e + 4 <= n; e += 4) {
                    const int u = base + e;
                    const float d0 = link_d2f(lf4[u], fx, fy, fz), d1 = link_d2f(lf4[u + 1], fx, fy, fz);
                    const float d2 = link_d2f(lf4[u + 2], fx, fy, fz), d3 = link_d2f(lf4[u + 3], fx, fy, fz);
                    // bitwise, not short-circuit: all four chains are issued before the one branch
                    const bool h = (!(d0 > thr) & (u != seed)) | (!(d1 > thr) & (u + 1 != seed)) |
                                   (!(d2 > thr) & (u + 2 != seed)) | (!(d3 > thr) & (u + 3 != seed));
                    if (h) {
#pragma unroll 1
                        for (int k = 0; k < 4; k++) {
                            if (!(link_d2f(lf4[u + k], fx, fy, fz) > thr) && u + k != seed) {
                                const double dk = link_d2(lxy, lzd, u + k, px, py, pz);
                                if (!(dk > A.best)) {
                                    link_eval(A, u + k, dk, lzd[u + k].y, speed, range_step);
                                    thr = link_threshold(A.best);
                                }
                            }
                        }
                    }
                }
#pragma unroll 1
                for (; e < n; e++) {
                    const int u = base + e;
                    if (!(link_d2f(lf4[u], fx, fy, fz) > thr) && u != seed) {
                        const double dk = link_d2(lxy, lzd, u, px, py, pz);
                        if (!(dk > A.best)) {
                            link_eval(A, u, dk, lzd[u].y, speed, range_step);
                            thr = link_threshold(A.best);
                        }
                    }
                }
            }
        }
    } else {
        const F2 nfx = f2_dup(-fx), nfy = f2_dup(-fy), nfz = f2_dup(-fz);
        // fp64 evaluation of the entries of step `blk` flagged in `hit`
        auto visit = [&](int blk, unsigned hit) {
            const int lim = blk < TEAM / 4 ? n_te : TEAM + n_me;
            while (hit) {
                const int u = blk * 4 + __ffs(hit) - 1;
                hit &= hit - 1;
                if (u < lim) {
                    const double dk = link_d2(lxy, lzd, u, px, py, pz);
                    if (!(dk > A.best)) {
                        link_eval(A, u, dk, lzd[u].y, speed, range_step);
                        thr = link_threshold(A.best);
                    }
                }
            }
        };
        // Eight entries (two blocks) per step; both regions are padded with far images to a multiple
        // of eight.  The step that holds the seed always has a pair within the threshold (the seed
        // itself) and the seeds of the lanes are spread over all steps, so the sweep leaves that step
        // out and all lanes visit theirs together afterwards, with the seed masked.
        const int sblk = seed >= 0 ? (seed >> 3) * 2 : -2;
        const int te_end = ((n_te + 7) >> 3) * 2, me_end = TEAM / 4 + ((n_me + 7) >> 3) * 2;
        for (int blk = 0;; blk += 2) {
            if (blk == te_end) blk = TEAM / 4;  // from the target entries to the missile entries
            if (blk >= me_end) break;
            float d[8];
            link_block(lf, blk, nfx, nfy, nfz, d);
            link_block(lf, blk + 1, nfx, nfy, nfz, d + 4);
            const float m = fminf(fminf(fminf(d[0], d[1]), fminf(d[2], d[3])), fminf(fminf(d[4], d[5]), fminf(d[6], d[7])));
            if (!(m > thr) && blk != sblk) {  // rare
                unsigned hit = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) hit |= (unsigned)!(d[k] > thr) << k;
                visit(blk, hit);
            }
        }
        if (seed >= 0) {
            float d[8];
            link_block(lf, sblk, nfx, nfy, nfz, d);
            link_block(lf, sblk + 1, nfx, nfy, nfz, d + 4);
            const int ks = seed & 7;
#pragma unroll
            for (int k = 0; k < 8; k++) d[k] = k == ks ? INFINITY : d[k];
            const float m = fminf(fminf(fminf(d[0], d[1]), fminf(d[2], d[3])), fminf(fminf(d[4], d[5]), fminf(d[6], d[7])));
            if (!(m > thr)) {  // rare
                unsigned hit = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) hit |= (unsigned)!(d[k] > thr) << k;
                visit(sblk, hit);
            }
        }
    }
    unsure = A.unsure;
    cls = A.ue < 0 ? 0 : (A.ue >= TEAM ? 2 : 1);
    cand = A.ue < 0 ? -1 : (A.ue >= TEAM ? A.ue - TEAM : A.ue);
}

// Is a link position (or a detection) inside the range where the fp32 image is trusted?  Tested on
// the image itself, branch-free: |x| + |y| + |z| <= C implies every |coordinate| <= C (+ a quarter
// metre of rounding, nothing against the margins above); NaN and inf fail the compare.
__device__ __forceinline__ bool link_f32_ok(float x, float y, float z) {
    return fabsf(x) + fabsf(y) + fabsf(z) <= (float)ZRK_F32_COORD_MAX;
}

// ------------------------------------------------------------------ association certificate
//
// Nearly every tick of nearly every scenario is "quiet": each detection re-acquires the track that
// followed its own slot last tick (CCP.link_object returns that entry), nothing is appended and
// nothing can be launched.  link_object still scans every track to find that out — O(detections x
// tracks) distance tests, the dominant cost of the tick.  The certificate proves the outcome of the
// scan without running it.
//
// Every air object flies a straight line from the tick it is created (Trajectory.get_pos,
// AirObject.py:23-25; mid-course retargeting is dead code), a radar moves it by at
// most `nb` metres per tick (sum over radars of the largest |noise vector| of the table), and the
// position link_object measures a track to is the tracked object's previous position
// (CCP.py:181-183).  So for the detection of slot i at tick t, with c(.) the noise-free positions:
//     distance to its own track     = |p_i(t) - p_i(t-1)|               <= |v_i| dt + 2 nb
//     distance to a track on slot j >= |c_j(t-1) - c_i(t)| - 2 nb           (j moving: it measures to p_j(t-1) = c_j(t-1) + noise)
//                                   >= |F_j - c_i(t)| - nb                   (j removed or not yet born: frozen at F_j)
// So the own track is strictly the nearest
// whenever |c_j(t-1) - c_i(t)| > R_i = |v_i| dt + 6 nb + slack for every other slot j (2 nb for the own track, 2 nb for
// the pair, 2 nb because the quadratic below starts from the current noisy positions); for a slot that is gone,
// c_j(t-1) is the point F_j it is frozen at.  cert_recompute solves |dp + dv tau| <= R_i with dp one tick of v_j behind
// the slots' current distance, for every j in fp64, and keeps, per slot,
// the first window of ticks [cin, cout] in which some pair is that close: the slot is "certified"
// while tick < cin and looks again once tick > cout.  Anything that changes the picture — a launch,
// a birth, a removal, a track changing hands or being created — clears all windows (x_cflag); ticks
// on which that happens and detections that are not certified take the ordinary scan, so the
// certificate can only ever skip work, never change a result.
#define ZRK_CERT_NEVER 0x3fffffff
template <int TEAM>
__device__ __forceinline__ void team_barrier() {
    if (TEAM == 32) {
        __syncwarp();
    } else {
        __syncthreads();
    }
}
// The fp64 solve of one pair of the certificate: b = p_j - p_i and u = v_j - v_i at `tick`, sp = |v_i| + |v_j|.
// Merges the window of ticks (from tick_now on) in which the two are within R of each other into [best_in, best_out].
__device__ __forceinline__ void cert_solve(double bx, double by, double bz, double ux, double uy, double uz, double sp, int tick,
                                           int tick_now, double dt_s, double nb, int &best_in, int &best_out) {
    const double R = (sp * dt_s * 1.000000001 + 6.0 * nb + 1.0) * 1.000001;
    const double a = dot3(ux, uy, uz, ux, uy, uz), bq = dot3(bx, by, bz, ux, uy, uz);
    const double c = dot3(bx, by, bz, bx, by, bz) - R * R;
    const double disc = bq * bq - a * c;
    if (c > 0 && (bq >= 0 || disc < 0)) return;  // never within R of each other from now on
    int kin = 0, kout = ZRK_CERT_NEVER;          // (also what every NaN ends up with)
    if (a > 0 && disc >= 0) {
        // within R for tau in [tin, tout]; tick + k is tau = k dt away (the clock table deviates from
        // that by rounding only, which the 1e-9 relative + 1e-6 absolute allowance swallows)
        const double sq = sqrt(disc), ia = 1.0 / (a * dt_s);
        const double qo = (-bq + sq) * ia * 1.000000001 + 1e-6;
        if (qo < 1e9) kout = tick + (int)qo;  // the last tick inside
        if (c > 0) {                          // approaching: tin is ahead; ticks k < tin / dt are outside
            const double qi = (-bq - sq) * ia * 0.999999999 - 1e-6;
            kin = ZRK_CERT_NEVER;
            if (qi < 1e9) kin = tick + (qi > 0 ? (int)qi + 1 : 0);
            if (!(qi >= -1.0)) kin = 0;  // (NaN)
        }
    }
    if (kin > ZRK_CERT_NEVER) kin = ZRK_CERT_NEVER;
    if (kout > ZRK_CERT_NEVER) kout = ZRK_CERT_NEVER;
    if (kout < tick_now) return;  // (positions of an older tick: this window has already passed)
    if (kin < best_in) {
        best_in = kin;
        best_out = kout;
    } else if (kin == best_in && kout > best_out) {
        best_out = kout;
    }
}
// One pair of the certificate: slot `lane` (exact position / velocity pi, zi, vi, viz; fp32 images fpi, fvi)
// against slot j, whose tracks measure to the fp32 image pj and move with vj; the positions are those of
// `tick`, the answer is wanted from tick_now on.  Merges the window in
// which the two are within R of each other into [best_in, best_out].  Nearly all pairs never come near
// each other: they are dismissed by an fp32 evaluation of the quadratic whose margins (1e-4 relative,
// 3 m) are far wider than its rounding errors; the others are solved in fp64.
// Returns false if slot j cannot be modelled (the caller then gives up on this lane).
template <int TEAM>
__device__ __forceinline__ bool cert_pair(const Hot<TEAM> &H, int j, float4 pj, float4 vj, float4 fpi, float4 fvi, double2 pi,
                                          double2 zi, double2 vi, double viz, int tick, int tick_now, double dt_s, double nb,
                                          float fdt, float fnb, int &best_in, int &best_out) {
    {   // fp32: never within R of each other from now on?
        const float bx = fmaf(-vj.x, fdt, pj.x) - fpi.x, by = fmaf(-vj.y, fdt, pj.y) - fpi.y, bz = fmaf(-vj.z, fdt, pj.z) - fpi.z;
        const float ux = vj.x - fvi.x, uy = vj.y - fvi.y, uz = vj.z - fvi.z;
        const float R = fmaf(fpi.w, fdt, fnb) * 1.0001f;
        const float a = fmaf(uz, uz, fmaf(uy, uy, ux * ux)), bq = fmaf(bz, uz, fmaf(by, uy, bx * ux));
        const float p2 = fmaf(bz, bz, fmaf(by, by, bx * bx)), c = fmaf(-R, R, p2);
        const float b2 = bq * bq;
        if (c > 1e-4f * p2 && (bq >= 0.f || fmaf(-a, c, b2) < -1e-4f * fmaf(a, p2, b2))) return true;
    }
    const int flj = H.sflags()[j];
    const double spi = fabs(zi.y);
    double bx, by, bz, ux, uy, uz, spj;
    if (flj & F_ALIVE) {
        if (!(flj & F_HASPREV)) return false;  // (its first tick: its tracks still measure to an older position)
        const double2 pj2 = H.pxy()[j], zj = H.pzs()[j], vj2 = H.vxy()[j];
        const double vjz = H.qzw()[j].y;
        bx = (pj2.x - vj2.x * dt_s) - pi.x, by = (pj2.y - vj2.y * dt_s) - pi.y, bz = (zj.x - vjz * dt_s) - zi.x;  // (one tick behind)
        ux = vj2.x - vi.x, uy = vj2.y - vi.y, uz = vjz - viz;
        spj = 0.0;
    } else {
        double fx, fy, fz;
        if (flj & F_PENDING) {  // on the launcher until NEW_MISSILE is consumed (its track measures to pos)
            fx = H.pxy()[j].x, fy = H.pxy()[j].y, fz = H.pzs()[j].x;
        } else if (flj & F_HASPREV) {  // removed: prev_pos stays what it was
            fx = H.qxy()[j].x, fy = H.qxy()[j].y, fz = H.qzw()[j].x;
        } else {
            return false;
        }
        bx = fx - pi.x, by = fy - pi.y, bz = fz - zi.x;
        ux = -vi.x, uy = -vi.y, uz = -viz;
        spj = 0.0;
    }
    cert_solve(bx, by, bz, ux, uy, uz, spi + spj, tick, tick_now, dt_s, nb, best_in, best_out);
    return true;
}
// this slot as the others see it: where its tracks measure to (+ its speed), and how that moves
template <int TEAM>
__device__ __forceinline__ void cert_image(const Hot<TEAM> &H, int lane, int fl, float4 &p4, float4 &v4) {
    const float qn = __int_as_float(0x7fc00000);
    p4 = make_float4(qn, qn, qn, 0.f);
    v4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((fl & F_ALIVE) && (fl & F_HASPREV)) {
        const double2 a = H.pxy()[lane], z = H.pzs()[lane], v = H.vxy()[lane];
        p4 = make_float4((float)a.x, (float)a.y, (float)z.x, fabsf((float)z.y) * 1.000001f);
        v4 = make_float4((float)v.x, (float)v.y, (float)H.qzw()[lane].y, 0.f);
    } else if (!(fl & F_ALIVE) && (fl & F_PENDING)) {
        p4 = make_float4((float)H.pxy()[lane].x, (float)H.pxy()[lane].y, (float)H.pzs()[lane].x, 0.f);
    } else if (!(fl & F_ALIVE) && (fl & F_HASPREV)) {
        p4 = make_float4((float)H.qxy()[lane].x, (float)H.qxy()[lane].y, (float)H.qzw()[lane].x, 0.f);
    }
}
// How many tracks refer to each slot: a slot followed by two tracks measures the same distance to
// both, an untracked one has no own track — neither is certified.  The windows do not depend on the
// tracks (every slot counts as an obstacle, tracked or not), so when tracks change hands or are
// created this recount is all that is needed.  All lanes of the team call it together.
template <int TEAM>
__device__ __forceinline__ void cert_recount(unsigned char *sm, const int *me_ref, int lane, int n_te, int n_me, int *refc) {
    Hot<TEAM> H;
    H.sm = sm;
    refc[lane] = 0;
    team_barrier<TEAM>();
    if (lane < n_te) atomicAdd(&refc[H.te_ref()[lane]], 1);
    if (lane < n_me) atomicAdd(&refc[me_ref[lane]], 1);
    team_barrier<TEAM>();
}
// Rebuilds every window (and the reference counts).  All lanes of the team call this together.
// fp / fv: scratch for the fp32 images.
template <int TEAM>
__device__ ZRK_NOINLINE void cert_recompute(unsigned char *sm, const int *me_ref, int lane, int n_slots, int tick, int tick_now,
                                            int n_te, int n_me, double dt_s, double nb, int *cin, int *cout, int *refc, float4 *fp,
                                            float4 *fv, bool stale_only = false) {
    Hot<TEAM> H;
    H.sm = sm;
    const int fl = lane < n_slots ? H.sflags()[lane] : 0;
    float4 fpi, fvi;
    cert_image<TEAM>(H, lane, fl, fpi, fvi);
    if (lane < n_slots) {
        fp[lane] = fpi;
        fv[lane] = fvi;
    }
    cert_recount<TEAM>(sm, me_ref, lane, n_te, n_me, refc);  // (its barriers also publish the images)
    if (!(fl & F_ALIVE)) {
        cin[lane] = ZRK_CERT_NEVER;
        cout[lane] = ZRK_CERT_NEVER;
        return;
    }
    // stale_only (nothing has changed since the windows were built): a window that has not passed stays
    if (stale_only && !(cout[lane] < tick_now)) return;
    int best_in = ZRK_CERT_NEVER, best_out = ZRK_CERT_NEVER;
    if (!(fl & F_HASPREV)) {
        best_in = 0;
        best_out = tick;
    } else {
        const double2 pi = H.pxy()[lane], zi = H.pzs()[lane], vi = H.vxy()[lane];
        const double viz = H.qzw()[lane].y;
        const float fdt = (float)dt_s * 1.000001f, fnb = 6.0f * (float)nb * 1.000001f + 3.0f;
        for (int j = 0; j < n_slots; j++) {
            if (j == lane) continue;
            if (!cert_pair<TEAM>(H, j, fp[j], fv[j], fpi, fvi, pi, zi, vi, viz, tick, tick_now, dt_s, nb, fdt, fnb, best_in, best_out)) {
                best_in = 0;
                best_out = tick;
                break;
            }
        }
    }
    cin[lane] = best_in;
    cout[lane] = best_out;
}
// A removal (AirEnv.py:33-40) only turns moving objects into fixed ones: the windows of the other
// slots stay valid (wider than needed where they were due to a removed object) once the pairs with
// the removed slots — the bits of `removed`, now frozen at their prev_pos — are merged in.  Warp teams.
template <int TEAM>
__device__ ZRK_NOINLINE void cert_removed(unsigned char *sm, int lane, int fl, unsigned removed, int tick, double dt_s, double nb,
                                          int *cin, int *cout) {
    Hot<TEAM> H;
    H.sm = sm;
    if (!(fl & F_ALIVE) || !(fl & F_HASPREV)) return;
    int best_in = cin[lane], best_out = cout[lane];
    float4 fpi, fvi;
    cert_image<TEAM>(H, lane, fl, fpi, fvi);
    const double2 pi = H.pxy()[lane], zi = H.pzs()[lane], vi = H.vxy()[lane];
    const double viz = H.qzw()[lane].y;
    const float fdt = (float)dt_s * 1.000001f, fnb = 6.0f * (float)nb * 1.000001f + 3.0f;
    for (unsigned m = removed; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        float4 pj, vj;
        cert_image<TEAM>(H, j, H.sflags()[j], pj, vj);
        if (!cert_pair<TEAM>(H, j, pj, vj, fpi, fvi, pi, zi, vi, viz, tick, tick + 1, dt_s, nb, fdt, fnb, best_in, best_out)) {
            best_in = 0;
            best_out = tick;
            break;
        }
    }
    cin[lane] = best_in;
    cout[lane] = best_out;
}

// ------------------------------------------------------------------ quiet streak
//
// A run of consecutive ticks on which nothing happens but flying and tracking: no launcher or CCP
// bookkeeping, no removal, no missile about to be born or to detonate, every visibility decision made
// by the fp32 pre-filter or the exact path as usual, and every detection certified to re-acquire the
// track of its own slot (see "association certificate").  On such a tick the lanes exchange nothing
// but votes (and, with missiles in flight, the targets' positions by shuffle), so the slot's state —
// position, previous position, trajectory, the own track's update time, a missile's remaining
// detonate_period — stays in registers from tick to tick, nothing is written to shared memory, and
// the message counters are settled when the streak ends.  Every value is computed by the same fp64
// operations as in the general tick below; a tick is committed only if no lane objects, otherwise the
// streak ends BEFORE that tick with nothing of it applied and the general tick runs it from the
// state written back here.  Returns the number of ticks committed.
// The few detections of a streak tick that are not certified (bits of dmask) are checked against every
// track, all lanes at once (lane k: entry k of _target_dict and of _missile_dict): is any other track
// as near as the own one?  If not, link_object returns the own track; if so (returns true), the general
// tick decides.  What a slot's tracks measure to is its prev_pos — the position still in p while the
// tick is not committed (use_p) — or, for a slot that is gone, what was frozen in q.  Rare: out of line.
template <int TEAM>
__device__ ZRK_NOINLINE bool streak_rivals(unsigned dmask, int lane, bool use_p, double px, double py, double pz, double qx,
                                           double qy, double qz, int r_te, int r_me, int n_te, int n_me, double nx, double ny,
                                           double nz, double d2, int seedu, int key) {
    const double ax = use_p ? px : qx, ay = use_p ? py : qy, az = use_p ? pz : qz;
    const double tx = __shfl_sync(ZRK_FULL, ax, r_te), ty = __shfl_sync(ZRK_FULL, ay, r_te), tz = __shfl_sync(ZRK_FULL, az, r_te);
    const double mx = __shfl_sync(ZRK_FULL, ax, r_me), my = __shfl_sync(ZRK_FULL, ay, r_me), mz = __shfl_sync(ZRK_FULL, az, r_me);
    // link_object passes over the tracks already updated this tick (upd_time == now, CCP.py:195 / :208), and detections
    // are taken in the order of `key` (first radar that saw the slot, then slot; -1 = not seen): a track that is the
    // own track of a slot seen earlier than d has been re-acquired by that slot's detection when d's turn comes — by
    // induction over the order, every earlier detection being certified or passing this very check — and is no
    // candidate for d.
    const int kt = __shfl_sync(ZRK_FULL, key, r_te), km = __shfl_sync(ZRK_FULL, key, r_me);
    const int ot = __shfl_sync(ZRK_FULL, seedu, r_te), om = __shfl_sync(ZRK_FULL, seedu, r_me);
    const int gone_te = (kt >= 0 && ot == lane) ? kt : ZRK_IMAX, gone_me = (km >= 0 && om == TEAM + lane) ? km : ZRK_IMAX;
    bool rival = false;
    for (unsigned mm = dmask; mm; mm &= mm - 1) {
        const int d = __ffs(mm) - 1;
        const double cx = __shfl_sync(ZRK_FULL, nx, d), cy = __shfl_sync(ZRK_FULL, ny, d), cz = __shfl_sync(ZRK_FULL, nz, d);
        const double thr = __shfl_sync(ZRK_FULL, d2, d) * (1.0 + 1e-12);
        const int sd = __shfl_sync(ZRK_FULL, seedu, d), kd = __shfl_sync(ZRK_FULL, key, d);
        if (lane < n_te && lane != sd && !(gone_te < kd)) {
            const double ex = tx - cx, ey = ty - cy, ez = tz - cz;
            rival = rival || !(dot3(ex, ey, ez, ex, ey, ez) > thr);
        }
        if (lane < n_me && TEAM + lane != sd && !(gone_me < kd)) {
            const double ex = mx - cx, ey = my - cy, ez = mz - cz;
            rival = rival || !(dot3(ex, ey, ez, ex, ey, ez) > thr);
        }
    }
    return __ballot_sync(ZRK_FULL, rival) != 0;
}
// The sectors the radars will look at next tick (Radar.py:96-117), computed by one lane while the others
// finish this tick: written to the OTHER of two scratch copies, which everybody switches to after the vote
// that commits the tick — a sweeping radar then costs a streak tick no barrier of its own.  Out of line.
__device__ ZRK_NOINLINE void streak_sweep_next(const double *cur, double *nxt, const double *c_rad_par, int nR) {
    for (int r = 0; r < nR; r++) {
        const double *q = c_rad_par + RAD_PAR_STRIDE * r;
        double caz = cur[2 * r], cel = cur[2 * r + 1];
        if (q[10] == 0.0) radar_move(q, caz, cel);
        nxt[2 * r] = caz;
        nxt[2 * r + 1] = cel;
    }
}
// move_to_next_sector_circular for the radars of a streak tick that sweep (Radar.py:96-117): out of line
template <int TEAM>
__device__ ZRK_NOINLINE void streak_sweep(double *rc, const double *c_rad_par, int nR, int lane) {
    team_barrier<TEAM>();
    if (lane == 0) {
        for (int r = 0; r < nR; r++) {
            const double *q = c_rad_par + RAD_PAR_STRIDE * r;
            if (q[10] != 0.0) continue;
            double caz = rc[2 * r], cel = rc[2 * r + 1];
            radar_move(q, caz, cel);
            rc[2 * r] = caz;
            rc[2 * r + 1] = cel;
        }
    }
    team_barrier<TEAM>();
}
// A lane's share of the streak's state.  The careful tick of quiet_streak works on it where it lies; lean_run (warp
// teams) pulls what a plain tick needs into registers for as many ticks as it can run.
struct StreakLane {
    double sx, sy, sz, vx, vy, vz, t0;  // Trajectory (a slot that is gone: s = p, v = 0)
    double px, py, pz, qx, qy, qz;      // pos, prev_pos
    double upds, speed, period, dm2, r2hi, now_s;
    const double *noise;                // this slot's column of the noise table, row 0
    int tg, updt, tick, n_q, n_upd, noise_w, mt;
    int cinl, cout, n_look, r_te, r_me, seedu;  // lean_run's rival check: the window (cinl: 0 unless exactly one track follows the slot), the tracks this lane holds
};
// what lean_run needs besides, the same for all lanes of the warp: in shared memory behind the radar records
struct LeanArgs {
    double dt_s, range_step;
    const double *tsec;
    unsigned *mbase;       // the scenario's words in mask row 0, or null
    long long mstride;     // words from one tick's mask row to the next
    int mask_ticks, mask_words, noise_row, n_obj3, noise_period, nR, n_te, n_me;
};
#define ZRK_RREC 10        /* doubles per radar record of the lean loop */
#define ZRK_RREC_RADARS 8  /* radars the lean loop has room for */
#define ZRK_RREC_DOUBLES (ZRK_RREC * ZRK_RREC_RADARS + 16)
#define ZRK_LEAN_ALIVE 1u
#define ZRK_LEAN_MSL 2u
#define ZRK_LEAN_TRACK_OK 4u
#define ZRK_LEAN_UPDL 8u
#define ZRK_LEAN_HASPREV 16u
// The plain ticks of a streak (warp teams, radars that do not sweep): no detonation or removal under way, every
// visibility decided by the pre-filter, every detection re-acquiring its own track within the bounds the certificate
// relies on — certified, or (the uncertified ones) checked against every track by streak_rivals.  A function of its own
// so that its loop holds nothing in registers but what such a tick needs.  Runs ticks L.tick .. lean_end - 1 and stops
// before the first one that is not plain, leaving nothing of it behind (the careful tick of quiet_streak redoes it).
// The arithmetic is that of the careful tick, operation for operation.
template <int TEAM>
__device__ ZRK_NOINLINE bool streak_rivals(unsigned dmask, int lane, bool use_p, double px, double py, double pz, double qx,
                                           double qy, double qz, int r_te, int r_me, int n_te, int n_me, double nx, double ny,
                                           double nz, double d2, int seedu, int key);
__device__ ZRK_NOINLINE void lean_run(StreakLane *Lp, int rrec_off, int lane, unsigned flags, unsigned alive_mask, unsigned msl_mask,
                                      int lean_end) {
    ZRK_SMEM_DECL;
    const double *const rrec = (const double *)(zrk_smem + rrec_off);
    const LeanArgs &A = *(const LeanArgs *)(rrec + ZRK_RREC * ZRK_RREC_RADARS);
    const bool alive = flags & ZRK_LEAN_ALIVE, msl = flags & ZRK_LEAN_MSL, track_ok = flags & ZRK_LEAN_TRACK_OK, updl = flags & ZRK_LEAN_UPDL;
    const double sx = Lp->sx, sy = Lp->sy, sz = Lp->sz, vx = Lp->vx, vy = Lp->vy, vz = Lp->vz, t0 = Lp->t0;
    double px = Lp->px, py = Lp->py, pz = Lp->pz, qx = Lp->qx, qy = Lp->qy, qz = Lp->qz;
    double upds = Lp->upds, period = Lp->period, now_s = Lp->now_s;
    const double dm2 = Lp->dm2, r2hi = Lp->r2hi;
    const double *const noise = Lp->noise;
    const int tg = Lp->tg, nR = A.nR, cinl = Lp->cinl;
    int updt = Lp->updt, tick = Lp->tick, n_upd = 0, noise_w = Lp->noise_w, mt = Lp->mt;
    const int tick_in = tick;
    const double dt_m = msl ? A.dt_s : 0.0, range_step = A.range_step;
    const double *tsec = A.tsec + tick + 1;
    // (the arguments are copied out of shared memory: the stores below could alias it as far as the compiler knows)
    unsigned *const mbase = A.mbase;
    const long long mstride = A.mstride;
    const int mask_ticks = A.mask_ticks, mask_words = A.mask_words, noise_row = A.noise_row, n_obj3 = A.n_obj3, noise_period = A.noise_period;
    const bool has_masks = mbase != nullptr, writer = has_masks && lane == 0;
    unsigned *mrow = mbase + (long long)mt * mstride;
    const double *nq = noise + (size_t)noise_w * noise_row;
#pragma unroll 1
    do {
        const double now_nx = ZRK_LD_EARLY(tsec);
        double z0 = ZRK_LD_EARLY(nq), z1 = ZRK_LD_EARLY(nq + 1), z2 = ZRK_LD_EARLY(nq + 2);  // the first radar's noise, in flight early
        const double dtm = now_s - t0;
        double nx = sx + vx * dtm, ny = sy + vy * dtm, nz = sz + vz * dtm;  // AirObject.step (a lane that is gone: s = p, v = 0)
        bool odd = false;
        double per = period;
        if (msl_mask) {  // Missile.step: nothing but the clock of a missile that stays outside its radius
            const double tx = __shfl_sync(ZRK_FULL, nx, tg), ty = __shfl_sync(ZRK_FULL, ny, tg), tz = __shfl_sync(ZRK_FULL, nz, tg);
            const double ex = tx - nx, ey = ty - ny, ez = tz - nz;
            per = period - dt_m;
            odd = msl && (!(dot3(ex, ey, ez, ex, ey, ez) > r2hi) || !(per > 0));
        }
        if (writer) ZRK_ST_GLOBAL(mrow, alive_mask);
        int rf = -1;
        for (int r = 0; r < nR; r++) {  // SectorRadar.step
            const double *w = rrec + ZRK_RREC * r;
            const double dx = nx - w[0], dy = ny - w[1], dz = nz - w[2];
            int v;
            if (w[9] != 0.0) {
                // all of the azimuth, the upper hemisphere: radar_fast for that sector, without its detours
                const double d2 = dot3(dx, dy, dz, dx, dy, dz);
                const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
                const float hz = sqrt_fast(fmaf(fy, fy, fx * fx));
                const bool ok = fabsf(fz) > 1e-3f * hz && hz > 1e-3f * fabsf(fz);
                v = (d2 > w[4]) ? 0 : (!ok ? -1 : (fz < 0.f ? 0 : ((d2 > 0.0 && d2 <= w[3] && fz > 0.f) ? 1 : -1)));
            } else {
                v = radar_fast(dx, dy, dz, w[3], w[4], w[5], w[6], w[7], w[8]);
            }
            odd = odd || (alive && v < 0);
            const bool vis = alive && v > 0;
            if (vis) {  // smooth_objects
                if (r) {
                    const double *q = nq + r * n_obj3;
                    z0 = ZRK_LD_EARLY(q), z1 = ZRK_LD_EARLY(q + 1), z2 = ZRK_LD_EARLY(q + 2);
                }
                nx += z0;
                ny += z1;
                nz += z2;
                if (rf < 0) rf = r;
            }
            if (has_masks) {
                const unsigned bal = __ballot_sync(ZRK_FULL, vis);
                if (writer) ZRK_ST_GLOBAL(mrow + (1 + r) * mask_words, bal);
            }
        }
        const bool seen = rf >= 0;
        // CombatControlPoint.step: the own track as link_d2 measures it.  track_ok includes that its hi bound lies beyond
        // dm whatever d_t is, and with d_t < range_step its lo bound is 0: d2 <= dm2 then puts d2 surely inside.
        const double ex = px - nx, ey = py - ny, ez = pz - nz;
        const double d2 = dot3(ex, ey, ez, ex, ey, ez);
        const bool fine = track_ok && d2 <= dm2 && now_s - upds < range_step;
        odd = odd || (seen && !fine);
        if (__ballot_sync(ZRK_FULL, odd) != 0) break;  // (nothing of this tick is kept)
        const unsigned dmask = __ballot_sync(ZRK_FULL, seen && !(tick < cinl));
        if (ZRK_RARE(dmask)) {
            // detections that are not certified: no other track as near as the own one?  (A window that has passed
            // is the careful tick's business: it has them rebuilt.)
            if (tick >= Lp->n_look && __ballot_sync(ZRK_FULL, seen && !(tick < cinl) && Lp->cout < tick) != 0) break;
            if (streak_rivals<32>(dmask, lane, alive || !(flags & ZRK_LEAN_HASPREV), px, py, pz, qx, qy, qz, Lp->r_te, Lp->r_me, A.n_te, A.n_me,
                                  nx, ny, nz, d2, Lp->seedu, seen ? (rf << 8) | lane : -1))
                break;
        }
        if (alive) {
            qx = px, qy = py, qz = pz;
            px = nx, py = ny, pz = nz;
        }
        if (seen) {
            upds = now_s;
            updt = tick;
        }
        period = per;
        n_upd += __popc(__ballot_sync(ZRK_FULL, seen && updl));
        tick++;
        tsec++;
        now_s = now_nx;
        nq += noise_row;
        if (++noise_w == noise_period) noise_w = 0, nq = noise;
        mrow += mstride;
        if (++mt == mask_ticks) mt = 0, mrow = mbase;
    } while (tick < lean_end);
    Lp->px = px, Lp->py = py, Lp->pz = pz, Lp->qx = qx, Lp->qy = qy, Lp->qz = qz;
    Lp->upds = upds, Lp->period = period, Lp->now_s = now_s;
    Lp->updt = updt, Lp->n_q += tick - tick_in, Lp->tick = tick, Lp->n_upd += n_upd, Lp->noise_w = noise_w, Lp->mt = mt;
}
// The lean loop for the commonest shape — ONE radar that looks at all of the azimuth and the upper hemisphere (rrec[9]) —
// two ticks per pass.  A tick's own arithmetic depends on the tick before only through the position its track measures to
// (prev_pos), the track's update time and the missile clock, so the two ticks of a pass are computed side by side in
// straight-line, branch-free code (the scheduler interleaves their dependent fp64 chains) and settled by ONE vote
// (a 4-bit OR reduction: either tick odd, either tick holding an uncertified detection); nothing is stored — mask rows
// included — before the tick it belongs to is committed.  The arithmetic is that of lean_run / the careful tick,
// operation for operation; the pre-filter's horizon / zenith guard compares squares instead of taking a root (it only
// ever decides what is far from both).  Stops before the first tick that is not plain, or with fewer than two ticks to go.
__device__ __forceinline__ void lean_run1(StreakLane *Lp, int rrec_off, int lane, unsigned flags, unsigned alive_mask, unsigned msl_mask, int lean_end) {
    ZRK_SMEM_DECL;
    const double *const rrec = (const double *)(zrk_smem + rrec_off);
    const LeanArgs &A = *(const LeanArgs *)(rrec + ZRK_RREC * ZRK_RREC_RADARS);
    const bool alive = flags & ZRK_LEAN_ALIVE, msl = flags & ZRK_LEAN_MSL, track_ok = flags & ZRK_LEAN_TRACK_OK, updl = flags & ZRK_LEAN_UPDL;
    const double sx = Lp->sx, sy = Lp->sy, sz = Lp->sz, vx = Lp->vx, vy = Lp->vy, vz = Lp->vz, t0 = Lp->t0;
    double px = Lp->px, py = Lp->py, pz = Lp->pz, qx = Lp->qx, qy = Lp->qy, qz = Lp->qz;
    double upds = Lp->upds, period = Lp->period, now_s = Lp->now_s;
    const double dm2 = Lp->dm2, r2hi = Lp->r2hi;
    const double rx = rrec[0], ry = rrec[1], rz = rrec[2], r2in = rrec[3], r2out = rrec[4];
    const double *const noise = Lp->noise;
    const int tg = Lp->tg, cinl = Lp->cinl;
    int updt = Lp->updt, tick = Lp->tick, cnt = 0, noise_w = Lp->noise_w, mt = Lp->mt;
    const int tick_in = tick;
    const double dt_m = msl ? A.dt_s : 0.0, range_step = A.range_step;
    const double *tsec = A.tsec + tick + 1;
    unsigned *const mbase = A.mbase;
    const long long mstride = A.mstride;
    const int mask_ticks = A.mask_ticks, mask_words = A.mask_words, noise_row = A.noise_row, noise_period = A.noise_period;
    const bool has_masks = mbase != nullptr, writer = has_masks && lane == 0;
#pragma unroll 1
    while (tick + 2 <= lean_end) {
        const int wb = noise_w + 1 == noise_period ? 0 : noise_w + 1, mb = mt + 1 == mask_ticks ? 0 : mt + 1;
        const double *const na = noise + (long long)noise_w * noise_row, *const nb = noise + (long long)wb * noise_row;
        const double now_b = ZRK_LD_EARLY(tsec), now_c = ZRK_LD_EARLY(tsec + 1);
        const double a0 = ZRK_LD_EARLY(na), a1 = ZRK_LD_EARLY(na + 1), a2 = ZRK_LD_EARLY(na + 2);
        const double b0 = ZRK_LD_EARLY(nb), b1 = ZRK_LD_EARLY(nb + 1), b2 = ZRK_LD_EARLY(nb + 2);
        // AirObject.step of both ticks (a lane that is gone: s = p, v = 0)
        const double dta = now_s - t0, dtb = now_b - t0;
        double ax = sx + vx * dta, ay = sy + vy * dta, az = sz + vz * dta;
        double bx = sx + vx * dtb, by = sy + vy * dtb, bz = sz + vz * dtb;
        // Missile.step: nothing but the clock of a missile that stays outside its radius
        const double per_a = period - dt_m, per_b = per_a - dt_m;
        bool odd_a = false, odd_b = false;
        if (msl_mask) {
            const double tax = __shfl_sync(ZRK_FULL, ax, tg), tay = __shfl_sync(ZRK_FULL, ay, tg), taz = __shfl_sync(ZRK_FULL, az, tg);
            const double tbx = __shfl_sync(ZRK_FULL, bx, tg), tby = __shfl_sync(ZRK_FULL, by, tg), tbz = __shfl_sync(ZRK_FULL, bz, tg);
            const double eax = tax - ax, eay = tay - ay, eaz = taz - az, ebx = tbx - bx, eby = tby - by, ebz = tbz - bz;
            odd_a = msl & (!(dot3(eax, eay, eaz, eax, eay, eaz) > r2hi) | !(per_a > 0));
            odd_b = msl & (!(dot3(ebx, eby, ebz, ebx, eby, ebz) > r2hi) | !(per_b > 0));
        }
        // SectorRadar.step: radar_fast for that sector.  sure: decided here (beyond the outer bracket, or clear of the
        // horizon and the zenith); vis: inside the inner bracket and above the horizon
        bool vis_a, vis_b;
        {
            const double dx = ax - rx, dy = ay - ry, dz = az - rz;
            const double d2 = dot3(dx, dy, dz, dx, dy, dz);
            const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
            const float h2 = fmaf(fy, fy, fx * fx), z2 = fz * fz;
            const bool ok = (z2 > 1e-6f * h2) & (h2 > 1e-6f * z2), far = d2 > r2out;
            vis_a = alive & !far & ok & (fz > 0.f) & (d2 > 0.0) & (d2 <= r2in);
            odd_a = odd_a | (alive & !far & !(ok & ((fz < 0.f) | vis_a)));
        }
        {
            const double dx = bx - rx, dy = by - ry, dz = bz - rz;
            const double d2 = dot3(dx, dy, dz, dx, dy, dz);
            const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
            const float h2 = fmaf(fy, fy, fx * fx), z2 = fz * fz;
            const bool ok = (z2 > 1e-6f * h2) & (h2 > 1e-6f * z2), far = d2 > r2out;
            vis_b = alive & !far & ok & (fz > 0.f) & (d2 > 0.0) & (d2 <= r2in);
            odd_b = odd_b | (alive & !far & !(ok & ((fz < 0.f) | vis_b)));
        }
        // smooth_objects
        ax += vis_a ? a0 : 0.0, ay += vis_a ? a1 : 0.0, az += vis_a ? a2 : 0.0;
        bx += vis_b ? b0 : 0.0, by += vis_b ? b1 : 0.0, bz += vis_b ? b2 : 0.0;
        // CombatControlPoint.step: the own track as link_d2 measures it (see lean_run); tick b measures to what tick a leaves
        const double eax = px - ax, eay = py - ay, eaz = pz - az;
        const double d2a = dot3(eax, eay, eaz, eax, eay, eaz);
        const double mx = alive ? ax : px, my = alive ? ay : py, mz = alive ? az : pz;
        const double ebx = mx - bx, eby = my - by, ebz = mz - bz;
        const double d2b = dot3(ebx, eby, ebz, ebx, eby, ebz);
        const double upd_b = vis_a ? now_s : upds;
        const bool fine_a = track_ok & (d2a <= dm2) & (now_s - upds < range_step);
        const bool fine_b = track_ok & (d2b <= dm2) & (now_b - upd_b < range_step);
        odd_a = odd_a | (vis_a & !fine_a);
        odd_b = odd_b | (vis_b & !fine_b);
        const bool unc_a = vis_a & !(tick < cinl), unc_b = vis_b & !(tick + 1 < cinl);
        const unsigned r = __reduce_or_sync(ZRK_FULL, (odd_a ? 1u : 0u) | (odd_b ? 2u : 0u) | (unc_a ? 4u : 0u) | (unc_b ? 8u : 0u));
        if (ZRK_RARE(r != 0)) {
            // not two plain ticks: tick a on its own; the next pass looks at tick b again, as its first tick
            if (r & 1u) break;  // (nothing of tick a is kept)
            if (r & 4u) {
                // detections that are not certified: no other track as near as the own one?  (A window that has passed
                // is the careful tick's business: it has them rebuilt.)
                if (tick >= Lp->n_look && __ballot_sync(ZRK_FULL, unc_a && Lp->cout < tick) != 0) break;
                if (streak_rivals<32>(__ballot_sync(ZRK_FULL, unc_a), lane, alive || !(flags & ZRK_LEAN_HASPREV), px, py, pz, qx, qy, qz, Lp->r_te,
                                      Lp->r_me, A.n_te, A.n_me, ax, ay, az, d2a, Lp->seedu, vis_a ? lane : -1))
                    break;
            }
            if (has_masks) {
                const unsigned bal = __ballot_sync(ZRK_FULL, vis_a);
                if (writer) {
                    unsigned *mrow = mbase + (long long)mt * mstride;
                    ZRK_ST_MASK(mrow, alive_mask);
                    ZRK_ST_MASK(mrow + mask_words, bal);
                }
            }
            if (alive) qx = px, qy = py, qz = pz, px = ax, py = ay, pz = az;
            if (vis_a) upds = now_s, updt = tick;
            period = per_a;
            cnt += vis_a && updl;
            tick++, tsec++, now_s = now_b, noise_w = wb, mt = mb;
            continue;
        }
        // both ticks are plain: commit them
        if (has_masks) {
            const unsigned bal_a = __ballot_sync(ZRK_FULL, vis_a), bal_b = __ballot_sync(ZRK_FULL, vis_b);
            if (writer) {
                unsigned *mra = mbase + (long long)mt * mstride, *mrb = mbase + (long long)mb * mstride;
                ZRK_ST_MASK(mra, alive_mask);
                ZRK_ST_MASK(mra + mask_words, bal_a);
                ZRK_ST_MASK(mrb, alive_mask);
                ZRK_ST_MASK(mrb + mask_words, bal_b);
            }
        }
        if (alive) qx = ax, qy = ay, qz = az, px = bx, py = by, pz = bz;
        if (vis_a) upds = now_s, updt = tick;
        if (vis_b) upds = now_b, updt = tick + 1;
        period = per_b;
        cnt += (vis_a && updl) + (vis_b && updl);
        tick += 2, tsec += 2, now_s = now_c;
        noise_w = wb + 1 == noise_period ? 0 : wb + 1;
        mt = mb + 1 == mask_ticks ? 0 : mb + 1;
    }
    const int n_upd = __reduce_add_sync(ZRK_FULL, cnt);
    Lp->px = px, Lp->py = py, Lp->pz = pz, Lp->qx = qx, Lp->qy = qy, Lp->qz = qz;
    Lp->upds = upds, Lp->period = period, Lp->now_s = now_s;
    Lp->updt = updt, Lp->n_q += tick - tick_in, Lp->tick = tick, Lp->n_upd += n_upd, Lp->noise_w = noise_w, Lp->mt = mt;
}
// AirEnv.py:33-40 inside a streak (warp teams): the slots of `rm` were removed at the start of this tick and stay frozen
// where their tracks measure to (their prev_pos, in q).  Merges the pairs with them into this lane's window, as
// cert_removed does on the state a general tick leaves; the lane's own position is that of tick_pos.  Rare: out of line.
__device__ ZRK_NOINLINE void streak_frozen(unsigned rm, bool alive, double px, double py, double pz, double qx, double qy, double qz,
                                           double vx, double vy, double vz, double speed, int tick_pos, double dt_s, double nb, int &cin,
                                           int &cout) {
    for (unsigned mm = rm; mm; mm &= mm - 1) {
        const int j = __ffs(mm) - 1;
        const double fx = __shfl_sync(ZRK_FULL, qx, j), fy = __shfl_sync(ZRK_FULL, qy, j), fz = __shfl_sync(ZRK_FULL, qz, j);
        if (alive)
            cert_solve(fx - px, fy - py, fz - pz, -vx, -vy, -vz, fabs(speed) + 0.0, tick_pos, tick_pos + 1, dt_s, nb, cin, cout);
    }
}
// What a streak tick of a warp team leaves behind when missiles detonate on it (Missile._detonate, Missile.py:135-160:
// the MISSILE_DETONATE record, the list AirEnv.step and the radars read next tick) and when they detonated on the
// tick before (Radar.py:193-203: one DESTROYED_MISSILE per radar and detonation; the list is consumed).  det: how this
// lane's missile detonates now (0 no, 1 at its target, 2 end of life), det_mask: the lanes with det != 0; was / prev:
// the same for the tick before.  Called once the tick is committed.  Rare: out of line.
template <int TEAM, int MPB>
__device__ ZRK_NOINLINE void streak_events(const Params *Pg, unsigned char *sm, const unsigned char *cs, int b, int lane, int tick, int det,
                                           unsigned det_mask, int was, unsigned prev, int tg, int m, int nR) {
    const Params &P = *Pg;
    const Layout &o = P.lo;
    Hot<TEAM> H;
    H.sm = sm;
    int *const hdr = H.hdr();
    const int MP = MPB ? MPB : o.MP, par = tick & 1, prv = par ^ 1;
    zrk_msg *rec = P.rec + (size_t)b * P.log_cap;
    int nrec = hdr[H_NREC];
    const unsigned lt = (1u << lane) - 1u;
    const int n_det = __popc(det_mask), n_prev = __popc(prev);
    __syncwarp();  // everybody has read the cursor
    if (det) {
        const int r_det = __popc(det_mask & lt), my_id = H.sid()[lane], det_tid = det == 1 ? H.sid()[tg] : ZRK_NONE;
        if (logged(P, ZRK_MSG_MISSILE_DETONATE))
            put_rec(P, rec, nrec + r_det, tick, ZRK_MSG_MISSILE_DETONATE, det == 2, my_id, ZRK_NONE, my_id, det_tid, 0, 0, 0);
        ZRK_COLD_I(det_mid)[par * MP + r_det] = my_id;
        ZRK_COLD_I(det_tid)[par * MP + r_det] = det_tid;
        ZRK_COLD_I(det_flag)[par * MP + r_det] = det == 2;
        ZRK_COLD_I(m_status)[m] = ST_DETONATED;
    }
    if (logged(P, ZRK_MSG_MISSILE_DETONATE)) nrec += n_det;
    if (nR && logged(P, ZRK_MSG_DESTROYED_MISSILE)) {
        if ((prev >> lane) & 1u) {
            const int k = __popc(prev & lt), my_id = H.sid()[lane];
            const int *c_rad_id = (const int *)(cs + o.c_rad_id);
            for (int r = 0; r < nR; r++)
                put_rec(P, rec, nrec + r * n_prev + k, tick, ZRK_MSG_DESTROYED_MISSILE, was == 2, c_rad_id[r], ZRK_CCP_ID, my_id, ZRK_NONE,
                        0, 0, 0);
        }
        nrec += nR * n_prev;
    }
    if (lane == 0) {
        unsigned *cnt = H.counts();
        hdr[H_NREC] = nrec;
        hdr[H_DETN0 + par] = n_det;
        hdr[H_DETN0 + prv] = 0;
        cnt[ZRK_MSG_MISSILE_DETONATE] += (unsigned)n_det;
        if (nR) cnt[ZRK_MSG_DESTROYED_MISSILE] += (unsigned)(nR * n_prev);
    }
}
template <int TEAM, int MPB>
__device__ __forceinline__ int quiet_streak(const Params &P, unsigned char *sm, const unsigned char *cs, int lane, int b,
                                            int max_ticks, int noise_w0, float4 *fp, float4 *fv, double *dscr, double *rrec = nullptr) {
    // dscr: 32 doubles of scratch (two copies of the sweep state of up to 8 radars)
    // rrec: ZRK_RREC_DOUBLES of shared scratch for the lean loop's radar records (warp teams), or null: no lean loop
    // sm: the scenario's state blob, cs: its parameter blob — in shared memory (the tick kernel) or in
    // global memory (the streak kernel: a streak touches each word once at its beginning and its end).
    // Inlined into its kernel: P is then the kernel parameter, read from the constant bank.
    const Layout &o = P.lo;
    constexpr MeOff XM = me_off(TEAM, MPB ? MPB : 2);
#define ZRK_QME_I(name) ((int *)(sm + (MPB ? XM.name : o.name)))
#define ZRK_QME_D(name) ((double *)(sm + (MPB ? XM.name : o.name)))
    Hot<TEAM> H;
    H.sm = sm;
    int *const hdr = H.hdr();
    int *const c_in = H.c_in(), *const c_out = H.c_out(), *const c_refc = H.c_refc();
    const int *chdr = (const int *)(cs + o.c_hdr);
    // Everything a lane reads of its own slot is asked for at once, before any of it is looked at (every array of the
    // hot layout has TEAM entries: a lane beyond the scenario's slots reads padding, and drops it below) — a streak that
    // begins in HBM (the streak kernel) then waits for one round trip to memory instead of a chain of them.
    const int fl_raw = H.sflags()[lane], own_raw = H.own()[lane];
    const double2 ld_a = H.pxy()[lane], ld_c = H.pzs()[lane], ld_q = H.qxy()[lane], ld_w = H.qzw()[lane];
    const double2 ld_s0 = H.txy()[lane], ld_s1 = H.tzt()[lane], ld_v = H.vxy()[lane];
    const int te_ref_raw = H.te_ref()[lane];
    const int cin_raw = c_in[lane], cout_raw = c_out[lane], refc_raw = c_refc[lane];
    const int nR = chdr[C_NR], T = o.T, tick_limit = chdr[C_TICK_LIMIT], n_slots = hdr[H_NSLOTS];
    // (the lane's state in plain variables: the record lean_run works on is filled in around its call only, so that
    // the careful tick below does not go through local memory for every value)
    double px, py, pz, qx, qy, qz, sx, sy, sz, t0, vx, vy, vz, speed, upds, period, r2hi, now_s;
    int tick, tg, updt, n_q, n_upd, noise_w, mt;
    tick = hdr[H_TICK];
    noise_w = noise_w0;
    mt = P.masks ? tick % P.mask_ticks : 0;
    const bool nolaunch = hdr[H_FREECAP] <= 0;
    int fl = lane < n_slots ? fl_raw : 0;
    bool alive = (fl & F_ALIVE) != 0, msl = alive && (fl & F_MISSILE);
    // lanes the streak cannot carry at all
    bool bail0 = (fl & F_PENDING) != 0 || (alive && !(fl & F_HASPREV)) || (TEAM > 32 && msl);
    px = 0, py = 0, pz = 0, qx = 0, qy = 0, qz = 0, sx = 0, sy = 0, sz = 0, t0 = 0, vx = 0, vy = 0, vz = 0, speed = 0;
    if (lane < n_slots) {
        px = ld_a.x, py = ld_a.y, pz = ld_c.x, speed = ld_c.y;
        qx = ld_q.x, qy = ld_q.y, qz = ld_w.x, vz = ld_w.y;
    }
    if (alive) {
        sx = ld_s0.x, sy = ld_s0.y, sz = ld_s1.x, t0 = ld_s1.y;
        vx = ld_v.x, vy = ld_v.y;
    }
    if (TEAM == 32 && !alive) sx = px, sy = py, sz = pz, vx = 0.0, vy = 0.0, vz = 0.0;  // as a trajectory: it stays where it is
    // tracks number `lane` of the two dicts: the slots they follow (fixed during a streak)
    const int n_te = hdr[H_NTE], n_me = hdr[H_NME];
    const int r_te = lane < n_te ? te_ref_raw : 0, r_me = lane < n_me ? ZRK_QME_I(me_ref)[lane] : 0;
    // the own track
    bool seed_ok = false, is_me = false, follow = false;
    int e = 0, seedu = -1;
    updt = -1;
    upds = 0;
    if (alive) {
        const int seed = own_raw;
        seedu = seed;
        if (seed >= 0) {
            is_me = seed >= TEAM;
            e = is_me ? seed - TEAM : seed;
            seed_ok = (is_me ? ZRK_QME_I(me_ref)[e] : H.te_ref()[e]) == lane;
            upds = is_me ? ZRK_QME_D(me_upds)[e] : H.te_upds()[e];
            follow = is_me || H.te_follow()[e] != 0;
        }
    }
    int cin = cin_raw, cout = cout_raw;
    bool refc1 = refc_raw == 1;
    const double dm = fabs(speed) * P.dt_s * 1.000000001 + 2.0 * P.noise_bound + 1e-3, dm2 = dm * dm;
    // a missile in flight: no detonation as long as it stays surely outside its radius of its target
    int m = 0;
    tg = 0;
    period = 0, r2hi = 0;
    if (msl) {
        m = (fl >> F_STATIC_SHIFT) - T;
        tg = ZRK_COLD_I(m_target)[m];
        period = ZRK_COLD_D(m_period)[m];
        const double r = ((const double *)(cs + o.c_m_radius))[m];
        r2hi = r * r * (1.0 + 1e-15);
        if (!(r >= 0.0) || tg < 0 || tg >= n_slots) bail0 = true;
    }
    unsigned alive_mask = __ballot_sync(ZRK_FULL, alive), msl_mask = __ballot_sync(ZRK_FULL, msl);
    const double *noise = P.noise ? P.noise + (size_t)b * P.noise_stride + (fl >> F_STATIC_SHIFT) * 3 : nullptr;
    const int noise_row = P.noise_radars * P.noise_objects * 3;
    const double *rad_cur = ZRK_QME_D(rad_cur);
    const double *const c_rad_pos = (const double *)(cs + o.c_rad_pos), *const c_rad_par = (const double *)(cs + o.c_rad_par);
    bool sweeping = false;
    for (int r = 0; r < nR; r++) sweeping = sweeping || c_rad_par[RAD_PAR_STRIDE * r + 10] == 0.0;
    // sweeping radars: the sweep state lives in two scratch copies during the streak (see streak_sweep_next)
    const bool sweep_db = sweeping && nR <= 8;
    int rcp = 0;
    if (sweep_db) {
        if (lane < 2 * nR) dscr[lane] = rad_cur[lane];
        team_barrier<TEAM>();
        rad_cur = dscr;
    }
    // The lean loop (warp teams, radars that do not sweep): the plain ticks of a streak — no detonation or removal under
    // way, every detection certified, every visibility decided by the pre-filter — run in a loop of their own that holds
    // nothing but what such a tick needs; a tick it cannot commit is left untouched for the careful tick below it.
    // Per radar: position, the squared-range bracket, the sector (ZRK_RREC doubles, in shared memory).
    const bool lean = TEAM == 32 && rrec != nullptr && !sweeping && nR <= ZRK_RREC_RADARS && noise != nullptr;
    if (TEAM == 32 && lean) {
        if (lane < nR) {
            const double *q = c_rad_par + RAD_PAR_STRIDE * lane;
            double *w = rrec + ZRK_RREC * lane;
            const double caz = rad_cur[2 * lane], cel = rad_cur[2 * lane + 1];
            w[0] = c_rad_pos[3 * lane], w[1] = c_rad_pos[3 * lane + 1], w[2] = c_rad_pos[3 * lane + 2];
            w[3] = q[8], w[4] = q[9];
            w[5] = caz, w[6] = caz + q[3], w[7] = cel, w[8] = cel + q[4];
            w[9] = (w[5] <= 0.0 && w[6] >= 360.0 && w[7] == 0.0 && w[8] == 90.0) ? 1.0 : 0.0;  // all of the azimuth, the upper hemisphere
        }
        if (lane == 0) {
            LeanArgs &A = *(LeanArgs *)(rrec + ZRK_RREC * ZRK_RREC_RADARS);
            const long long roww = (long long)(1 + o.RP) * P.mask_words;
            A.dt_s = P.dt_s, A.range_step = P.range_step, A.tsec = P.tsec;
            A.mbase = P.masks ? P.masks + (long long)b * roww : nullptr;
            A.mstride = (long long)o.B * roww;
            A.mask_ticks = P.masks ? P.mask_ticks : 1, A.mask_words = P.mask_words, A.noise_row = noise_row, A.n_obj3 = P.noise_objects * 3;
            A.noise_period = P.noise_period, A.nR = nR, A.n_te = n_te, A.n_me = n_me;
        }
        __syncwarp();

    }
    // what never changes for a lane during a streak, folded: its own track is where it should be and nothing can be
    // launched at it; its re-acquisitions are reported (CCP_UPDATE_TARGET)
    const bool track_ok = seed_ok && (follow || nolaunch) && !(msl && !is_me), updl = !is_me && follow;
    // (lean_run: the upper end of calc_range, speed (d_t + range_step) with d_t >= one tick, lies beyond dm by a wide margin)
    const bool hi_far = speed >= 0.0 && speed * P.range_step > 2.0 * P.noise_bound + 1.0 + speed * P.dt_s * 1e-6;
    int lean_end = 0;  // the lean loop runs the ticks before this one (0: find out)
    // The streak runs in segments: the blob holds the state of the tick before the segment's first one; a segment ends
    // when the streak does, or (again) when windows that have passed are to be rebuilt — from the blob, which therefore
    // has to be brought up to date first.
    int n_done = 0;            // ticks committed by the segments before this one
    int n_look = 0;            // no window rebuild before this tick (teams of a warp)
    bool looked = false;       // (larger teams: the windows have been rebuilt once during this streak)
    bool again;
    // detonations of the last committed tick (warp teams): the lanes of the missiles, how this lane's missile went
    // (1 at its target, 2 end of life), and the slots AirEnv.step removes because of them at the start of the next tick
    unsigned prev_det = 0, prev_rm = 0;
    int was = 0;
    now_s = tick_seconds(P, tick);
    do {
        again = false;
        n_q = 0, n_upd = 0;
        int acc_alive = 0, acc_msl = 0;  // alive objects / missiles summed over the segment's ticks, beyond n_q x the final counts
        const bool alive0 = alive, msl0 = msl;
        const int fl0 = fl;
        bool cw_dirty = false;
        unsigned rm = 0;  // the slots removed at the start of the tick under way (not committed yet)
        updt = -1;
        while (n_done + n_q < max_ticks) {
            if (tick_limit >= 0 && tick >= tick_limit) break;
            if (TEAM == 32 && lean && !prev_det) {
                if (lean_end == 0) {
                    // the end of the streak; the end of the clock table
                    lean_end = tick + (max_ticks - n_done - n_q);
                    if (tick_limit >= 0 && lean_end > tick_limit) lean_end = tick_limit;
                    if (lean_end > P.tsec_n - 1) lean_end = P.tsec_n - 1;
                    if (__ballot_sync(ZRK_FULL, bail0) != 0) lean_end = -1;
                }
                if (tick < lean_end) {
                    const unsigned lflags = (alive ? ZRK_LEAN_ALIVE : 0u) | (msl ? ZRK_LEAN_MSL : 0u) | (track_ok && hi_far ? ZRK_LEAN_TRACK_OK : 0u) |
                                            (updl ? ZRK_LEAN_UPDL : 0u) | ((fl & F_HASPREV) ? ZRK_LEAN_HASPREV : 0u);
                    const int rrec_off = (int)((unsigned char *)rrec - zrk_smem_base());
                    StreakLane L;
                    L.sx = sx, L.sy = sy, L.sz = sz, L.vx = vx, L.vy = vy, L.vz = vz, L.t0 = t0;
                    L.px = px, L.py = py, L.pz = pz, L.qx = qx, L.qy = qy, L.qz = qz;
                    L.upds = upds, L.speed = speed, L.period = period, L.dm2 = dm2, L.r2hi = r2hi, L.now_s = now_s;
                    L.noise = noise;
                    L.tg = tg, L.updt = updt, L.tick = tick, L.n_q = n_q, L.n_upd = n_upd, L.noise_w = noise_w, L.mt = mt;
                    L.cinl = refc1 ? cin : 0, L.cout = cout, L.n_look = n_look, L.r_te = r_te, L.r_me = r_me, L.seedu = seedu;
                    if (nR == 1 && rrec[9] != 0.0 && tick + 2 <= lean_end) {  // one radar, the upper hemisphere: two ticks per pass
                        lean_run1(&L, rrec_off, lane, lflags, alive_mask, msl_mask, lean_end);
                    } else {
                        lean_run(&L, rrec_off, lane, lflags, alive_mask, msl_mask, lean_end);
                    }
                    px = L.px, py = L.py, pz = L.pz, qx = L.qx, qy = L.qy, qz = L.qz;
                    upds = L.upds, period = L.period, now_s = L.now_s;
                    updt = L.updt, n_q = L.n_q, tick = L.tick, n_upd = L.n_upd, noise_w = L.noise_w, mt = L.mt;
                    if (tick >= lean_end) {
                        lean_end = 0;
                        continue;  // (the loop's own conditions decide what comes next)
                    }
                }
            }
            const double now_next = tick_seconds(P, tick + 1);
            bool bail = bail0;
            if (TEAM == 32 && ZRK_RARE(prev_det)) {
                // AirEnv.step (AirEnv.py:33-40): the objects last tick's MISSILE_DETONATE messages name are gone
                rm = prev_rm & alive_mask;
                alive = alive && !((rm >> lane) & 1u);
                msl = msl && alive;
                alive_mask &= ~rm;
                msl_mask = __ballot_sync(ZRK_FULL, msl);
                if (rm) {
                    streak_frozen(rm, alive, px, py, pz, qx, qy, qz, vx, vy, vz, speed, tick - 1, P.dt_s, P.noise_bound, cin, cout);
                    cw_dirty = true;
                    lean_end = 0;
                }
            }
            // AirObject.step (AirObject.py:39-42): the new position; the old one becomes prev_pos at commit
            double nx = px, ny = py, nz = pz;
            if (alive) {
                const double dtm = now_s - t0;
                nx = sx + vx * dtm;
                ny = sy + vy * dtm;
                nz = sz + vz * dtm;
            }
            // Missile.step (Missile.py:162-193): target's position of this tick, before any radar touches it
            double per = period;
            int det = 0;
            unsigned det_mask = 0;
            if (TEAM == 32 && msl_mask) {
                const double tx = __shfl_sync(ZRK_FULL, nx, tg), ty = __shfl_sync(ZRK_FULL, ny, tg), tz = __shfl_sync(ZRK_FULL, nz, tg);
                if (msl) {
                    const double ex = tx - nx, ey = ty - ny, ez = tz - nz;
                    const double d2m = dot3(ex, ey, ez, ex, ey, ez);
                    per = period - P.dt_s;
                    if (ZRK_RARE(!(d2m > r2hi))) {  // (may be) within the radius: the reference's own test decides
                        if (sqrt(d2m) <= ((const double *)(cs + o.c_m_radius))[m]) {
                            det = 1;
                            per = period;
                        }
                    }
                    if (ZRK_RARE(!(per > 0)) && !det) {  // end of life (a NaN period: the general tick's business)
                        if (per <= 0)
                            det = 2;
                        else
                            bail = true;
                    }
                }
                det_mask = __ballot_sync(ZRK_FULL, det != 0);
            }
            if (TEAM == 32 && ZRK_RARE((det_mask | prev_det) != 0)) {
                // the records of this tick have to fit (otherwise the general tick is the one to report the overflow)
                const int need = (logged(P, ZRK_MSG_MISSILE_DETONATE) ? __popc(det_mask) : 0) +
                                 (logged(P, ZRK_MSG_DESTROYED_MISSILE) ? nR * __popc(prev_det) : 0);
                if (need && hdr[H_NREC] + need > P.log_cap) bail = true;
            }
            // SectorRadar.step (Radar.py:144-205)
            int rf = -1;
            unsigned *mrow = nullptr;
            if (P.masks) {
                mrow = P.masks + ((size_t)mt * o.B + b) * (size_t)((1 + o.RP) * P.mask_words);  // (mt = tick % P.mask_ticks)
                if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[lane >> 5] = alive_mask;  // (this warp's 32 slots)
            }
            for (int r = 0; r < nR; r++) {
                const double *par_r = c_rad_par + RAD_PAR_STRIDE * r;
                bool vis = false;
                if (alive) {
                    if (TEAM == 32 && lean) {
                        // (the radar records of the lean loop, in shared memory: the same values, no trip to the parameter blob)
                        const double *w = rrec + ZRK_RREC * r;
                        const double dx = nx - w[0], dy = ny - w[1], dz = nz - w[2];
                        const int v = radar_fast(dx, dy, dz, w[3], w[4], w[5], w[6], w[7], w[8]);
                        vis = ZRK_OFTEN(v >= 0) ? v != 0 : radar_sees_exact(dx, dy, dz, par_r[2], w[5], w[6], w[7], w[8]);
                    } else {
                        const double caz = rad_cur[2 * r], cel = rad_cur[2 * r + 1];
                        const double dx = nx - c_rad_pos[3 * r], dy = ny - c_rad_pos[3 * r + 1], dz = nz - c_rad_pos[3 * r + 2];
                        vis = radar_sees<true>(dx, dy, dz, par_r, caz, caz + par_r[3], cel, cel + par_r[4]);
                    }
                    if (vis) {  // smooth_objects (Radar.py:138-142)
                        if (noise) {
                            const double *q = noise + (size_t)noise_w * noise_row + (size_t)r * P.noise_objects * 3;
                            nx += q[0];
                            ny += q[1];
                            nz += q[2];
                        }
                        if (rf < 0) rf = r;
                    }
                }
                if (mrow) {
                    const unsigned bal = __ballot_sync(ZRK_FULL, vis);
                    if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[(1 + r) * P.mask_words + (lane >> 5)] = bal;
                }
            }
            if (sweep_db && lane == 0) streak_sweep_next(rad_cur, dscr + 16 * (rcp ^ 1), c_rad_par, nR);
            // CombatControlPoint.step: the own track, exactly as link_d2 / link_eval measure it
            const bool det_ = alive && rf >= 0;
            bool upd_msg = false, doubt = false;
            double d2 = 0;
            if (det_) {
                const double ex = px - nx, ey = py - ny, ez = pz - nz;
                d2 = dot3(ex, ey, ez, ex, ey, ez);
                double lo, hi;
                calc_range(speed, now_s - upds, P.range_step, lo, hi);
                const double EPS = 4e-15;
                const double l2 = lo * lo, h2 = hi * hi;
                const bool in_sure = d2 >= l2 * (1.0 + EPS) && d2 <= h2 * (1.0 - EPS);
                // (a Missile followed by a target track is the reference's AttributeError, CCP.py:259)
                const bool hard = !(seed_ok && in_sure && (follow || nolaunch)) || (msl && !is_me);
                doubt = !(tick < cin && refc1 && d2 <= dm2);
                bail = bail || hard;
                upd_msg = !is_me && follow;
            }
            if (TEAM == 32) {
                if (__ballot_sync(ZRK_FULL, bail) != 0) break;
                const unsigned dmask = __ballot_sync(ZRK_FULL, doubt);
                if (ZRK_RARE(dmask)) {
                    if (streak_rivals<TEAM>(dmask, lane, alive || !(fl & F_HASPREV), px, py, pz, qx, qy, qz, r_te, r_me, n_te, n_me, nx, ny,
                                            nz, d2, seedu, det_ ? (rf << 8) | lane : -1))
                        break;
                    // a window that has passed keeps its slot uncertified: end the segment, rebuild the windows and look
                    // at this tick again
                    if (tick >= n_look && __ballot_sync(ZRK_FULL, doubt && cout < tick) != 0) {
                        again = true;
                        break;
                    }
                }
            } else if (__syncthreads_or(bail || doubt) != 0) {
                if (((tick & 15) == 0 || (n_done + n_q == 0 && !looked)) && __syncthreads_or(!bail && doubt && cout < tick) != 0) {  // (as above)
                    again = true;
                    break;
                }
                // Teams larger than a warp check their uncertified detections like a warp does (streak_rivals),
                // through shared scratch instead of shuffles: every slot publishes what its tracks measure to, the
                // uncertified detections line up in a short list, every lane tests the two tracks it holds.
                {
                    double *const Lx = (double *)fp, *const Ly = Lx + TEAM, *const Lz = Ly + TEAM;  // (fp: 48 x TEAM bytes, see the caller)
                    double *const dl = Lz + TEAM;                                                  // [16][4] x, y, z, threshold
                    int *const dsd = (int *)(dl + 64), *const dcnt = dsd + 16;                     // [16] own track, list length
                    if (lane == 0) *dcnt = 0;  // (the scratch is shared with the window rebuild: emptied before every use)
                    if (__syncthreads_or(bail) != 0) break;
                    const bool use_p = alive || !(fl & F_HASPREV);
                    if (lane < n_slots) Lx[lane] = use_p ? px : qx, Ly[lane] = use_p ? py : qy, Lz[lane] = use_p ? pz : qz;
                    if (doubt) {
                        const int k = atomicAdd(dcnt, 1);
                        if (k < 16) {
                            dl[4 * k] = nx, dl[4 * k + 1] = ny, dl[4 * k + 2] = nz, dl[4 * k + 3] = d2 * (1.0 + 1e-12);
                            dsd[k] = seedu;
                        }
                    }
                    __syncthreads();
                    const int nd = *dcnt;
                    bool rival = nd > 16;  // (too many: let the general tick do it)
                    for (int k = 0; k < nd && k < 16; k++) {
                        const double cx = dl[4 * k], cy = dl[4 * k + 1], cz = dl[4 * k + 2], thr = dl[4 * k + 3];
                        const int sd = dsd[k];
                        if (lane < n_te && lane != sd) {
                            const double ex = Lx[r_te] - cx, ey = Ly[r_te] - cy, ez = Lz[r_te] - cz;
                            rival = rival || !(dot3(ex, ey, ez, ex, ey, ez) > thr);
                        }
                        if (lane < n_me && TEAM + lane != sd) {
                            const double ex = Lx[r_me] - cx, ey = Ly[r_me] - cy, ez = Lz[r_me] - cz;
                            rival = rival || !(dot3(ex, ey, ez, ex, ey, ez) > thr);
                        }
                    }
                    if (__syncthreads_or(rival) != 0) break;
                }
            }
            // ---- commit
            if (TEAM == 32 && ZRK_RARE((det_mask | prev_det) != 0)) {
                streak_events<TEAM, MPB>(P.self, sm, cs, b, lane, tick, det, det_mask, was, prev_det, tg, m, nR);
                if (rm) {  // what left at the start of this tick had been counted as present on the n_q ticks before it
                    acc_alive += n_q * __popc(rm);
                    acc_msl += n_q * __popc(__ballot_sync(ZRK_FULL, ((rm >> lane) & 1u) && (fl & F_MISSILE)));
                    if ((rm >> lane) & 1u) {
                        fl &= ~F_ALIVE;
                        sx = px, sy = py, sz = pz, vx = 0.0, vy = 0.0, vz = 0.0;
                    }
                    rm = 0;
                }
                // the slots AirEnv.step will remove next tick: the missiles and what they hit
                prev_rm = det_mask;
                for (unsigned mm = det_mask; mm; mm &= mm - 1) {
                    const int d = __ffs(mm) - 1;
                    const int dk = __shfl_sync(ZRK_FULL, det, d), dt_ = __shfl_sync(ZRK_FULL, tg, d);
                    if (dk == 1) prev_rm |= 1u << dt_;
                }
                prev_det = det_mask;
                was = det;
            }
            if (alive) {
                qx = px, qy = py, qz = pz;
                px = nx, py = ny, pz = nz;
            }
            if (det_) {
                upds = now_s;
                updt = tick;
            }
            period = per;
            n_upd += __popc(__ballot_sync(ZRK_FULL, upd_msg));
            if (sweep_db) {
                if (TEAM == 32) __syncwarp();  // (a larger team's vote was a barrier already)
                rcp ^= 1;
                rad_cur = dscr + 16 * rcp;
            } else if (ZRK_RARE(sweeping)) {
                streak_sweep<TEAM>(ZRK_QME_D(rad_cur), c_rad_par, nR, lane);
            }
            tick++;
            n_q++;
            now_s = now_next;
            if (++noise_w == P.noise_period) noise_w = 0;
            if (P.masks && ++mt == P.mask_ticks) mt = 0;
        }
        if (TEAM == 32 && ZRK_RARE(rm)) {  // the tick under way was not committed: neither is its removal
            lean_end = 0;
            alive = alive || ((rm >> lane) & 1u);
            msl = alive && (fl & F_MISSILE);
            alive_mask |= rm;
            msl_mask = __ballot_sync(ZRK_FULL, msl);
        }
        if (n_q) {  // the state the general tick expects, and the messages of the committed ticks
            if (sweep_db && lane < 2 * nR) ZRK_QME_D(rad_cur)[lane] = rad_cur[lane];
            if (alive0) {
                H.qxy()[lane] = make_double2(qx, qy);
                H.qzw()[lane].x = qz;
                H.pxy()[lane] = make_double2(px, py);
                H.pzs()[lane].x = pz;
            }
            if (TEAM == 32 && fl != fl0) H.sflags()[lane] = fl;
            if (TEAM == 32 && cw_dirty) {
                c_in[lane] = cin;
                c_out[lane] = cout;
            }
            if (updt >= 0) {
                if (is_me) {
                    ZRK_QME_I(me_upd)[e] = updt;
                    ZRK_QME_D(me_upds)[e] = upds;
                } else {
                    H.te_upd()[e] = updt;
                    H.te_upds()[e] = upds;
                }
            }
            if (msl0) ZRK_COLD_D(m_period)[m] = period;
            if (TEAM == 32) {
                if (lane == 0) {
                    const int na = n_q * __popc(alive_mask) + acc_alive;
                    unsigned *cnt = H.counts();
                    hdr[H_TICK] = tick;
                    hdr[H_FAST] += n_q;
                    hdr[H_ALIVE] += na;
                    cnt[ZRK_MSG_ACTIVE_OBJECTS] += (unsigned)n_q;
                    cnt[ZRK_MSG_MISSILE_POS] += (unsigned)(n_q * __popc(msl_mask) + acc_msl);
                    cnt[ZRK_MSG_CCP_UPDATE_TARGET] += (unsigned)n_upd;
                    if (nR) {
                        cnt[ZRK_MSG_ALL_OBJECTS] += (unsigned)(n_q * nR);
                        cnt[ZRK_MSG_FOUND_OBJECTS] += (unsigned)(n_q * nR);
                        cnt[ZRK_MSG_DRAW_OBJECTS] += (unsigned)na;
                    }
                }
            } else {
                if ((lane & 31) == 0) {
                    const int na = __popc(alive_mask);
                    if (na) atomicAdd(&hdr[H_ALIVE], n_q * na);
                    if (n_upd) atomicAdd((int *)&H.counts()[ZRK_MSG_CCP_UPDATE_TARGET], n_upd);
                    if (nR && na) atomicAdd((int *)&H.counts()[ZRK_MSG_DRAW_OBJECTS], n_q * na);
                }
                if (lane == 0) {
                    unsigned *cnt = H.counts();
                    hdr[H_TICK] = tick;
                    hdr[H_FAST] += n_q;
                    cnt[ZRK_MSG_ACTIVE_OBJECTS] += (unsigned)n_q;
                    if (nR) {
                        cnt[ZRK_MSG_ALL_OBJECTS] += (unsigned)(n_q * nR);
                        cnt[ZRK_MSG_FOUND_OBJECTS] += (unsigned)(n_q * nR);
                    }
                }
            }
        }
        n_done += n_q;
        team_barrier<TEAM>();
        if (again) {
            // the windows that have passed, rebuilt from the state just written (that of tick - 1)
            looked = true;
            n_look = tick + 4;
            cert_recompute<TEAM>(sm, ZRK_QME_I(me_ref), lane, n_slots, tick - 1, tick, n_te, n_me, P.dt_s, P.noise_bound, c_in, c_out, c_refc,
                                 fp, fv, true);
            cin = c_in[lane], cout = c_out[lane], refc1 = c_refc[lane] == 1;
            lean_end = 0;
        }
    } while (again);
    return n_done;
#undef ZRK_QME_I
#undef ZRK_QME_D
}

// ------------------------------------------------------------------ lane 0 alone: launchers + CCP bookkeeping

// MissileLauncher.step for every launcher (MissileLauncher.py:82-138) and the CCP's first-tick
// capacity requests / LAUNCHED_MISSILE intake (CCP.py:122-169).  Rare, serial, kept out of line.
template <int TEAM>
__device__ ZRK_NOINLINE int serial_section(const Params *Pg, unsigned char *sm, zrk_msg *rec, int tick,
                                           double now_s, int mb_prev, int nrec) {
    const Params &P = *Pg;  // the copy in global memory: a by-reference kernel parameter would be spilled
    const Layout &o = P.lo;
    const int MP = o.MP, Q = o.Q, T = o.T;
    Hot<TEAM> H;
    H.sm = sm;
    int *const hdr = H.hdr();
    const int *chdr = ZRK_CST_I(c_hdr);
    const int nL = chdr[C_NL], nC = chdr[C_NC], ccp_id = chdr[C_CCP_ID];
    const bool has_ccp = ccp_id != ZRK_NONE;
    const int par = tick & 1, prv = par ^ 1;
    int nr = nrec;
    {
        int *mb_type = ZRK_COLD_I(mb_type), *mb_l = ZRK_COLD_I(mb_l), *mb_a = ZRK_COLD_I(mb_a);
        int *st_n = ZRK_COLD_I(st_n), *st = ZRK_COLD_I(st);
        const int *c_m_id = ZRK_CST_I(c_m_id), *c_l_id = ZRK_CST_I(c_l_id);
        const double *c_l_pos = ZRK_CST_D(c_l_pos);
        int *m_target = ZRK_COLD_I(m_target);
        unsigned *counts = H.counts();
        // MissileLauncher.step (MissileLauncher.py:82-138)
        for (int l = 0; l < nL && mb_prev; l++) {
            const int l_id = c_l_id[l];
            for (int i = 0; i < mb_prev; i++) {
                if (mb_l[prv * Q + i] != l) continue;
                const int ty = mb_type[prv * Q + i], arg = mb_a[prv * Q + i];
                if (ty == MB_COMMAND) {  // launch_missile (:58-80)
                    if (st_n[l] == 0) continue;
                    const int m = st[l * MP + (--st_n[l])];
                    const int m_id = c_m_id[m], tg = arg;
                    counts[ZRK_MSG_LAUNCH_MISSILE]++;
                    if (logged(P, ZRK_MSG_LAUNCH_MISSILE))
                        put_rec(P, rec, nr++, tick, ZRK_MSG_LAUNCH_MISSILE, 0, l_id, m_id, H.sid()[tg], ZRK_NONE, 0, 0, 0);
                    // Missile._launch (Missile.py:104-133)
                    double tp[3] = {H.pxy()[tg].x, H.pxy()[tg].y, H.pzs()[tg].x};
                    double vt[3];
                    if (H.sflags()[tg] & F_MISSILE) {
                        vt[0] = vt[1] = vt[2] = NAN;  // Missile.velocity is 0/0 (AirObject.py:35)
                    } else {
                        double wx = H.vxy()[tg].x, wy = H.vxy()[tg].y, wz = H.qzw()[tg].y;
                        double sp = H.pzs()[tg].y;
                        vt[0] = wx / sp * sp;  // target.velocity * target.speed_mod (Missile.py:59)
                        vt[1] = wy / sp * sp;
                        vt[2] = wz / sp * sp;
                    }
                    double V[3];
                    const double v0 = ZRK_CST_D(c_m_v0)[m];
                    int why = missile_solve(&c_l_pos[3 * l], tp, vt, v0, ZRK_CST_D(c_m_v0sq)[m], ZRK_COLD_D(m_period)[m], V);
                    const int q = hdr[H_MBN0 + par]++;
                    mb_l[par * Q + q] = l;
                    mb_a[par * Q + q] = m;
                    if (why < 0) {
                        const int k = hdr[H_NSLOTS]++;
                        H.txy()[k] = make_double2(c_l_pos[3 * l], c_l_pos[3 * l + 1]);  // new Trajectory (Missile.py:110-114)
                        H.tzt()[k] = make_double2(c_l_pos[3 * l + 2], now_s);
                        H.vxy()[k] = make_double2(V[0], V[1]);
                        H.qzw()[k] = make_double2(0.0, V[2]);
                        H.pxy()[k] = make_double2(c_l_pos[3 * l], c_l_pos[3 * l + 1]);
                        H.pzs()[k] = make_double2(c_l_pos[3 * l + 2], v0);
                        H.sid()[k] = m_id;
                        H.sflags()[k] = F_PENDING | F_MISSILE | ((T + m) << F_STATIC_SHIFT);
                        H.srf()[k] = -1;
                        H.own()[k] = -1;
                        ZRK_COLD_I(m_status)[m] = ST_ACTIVE;
                        m_target[m] = tg;
                        ZRK_COLD_I(m_slot)[m] = k;
                        ZRK_COLD_I(m_birth)[m] = tick + 2;
                        mb_type[par * Q + q] = MB_SUCCESS;
                        counts[ZRK_MSG_LAUNCH_SUCCESSFUL]++;
                        if (logged(P, ZRK_MSG_LAUNCH_SUCCESSFUL))
                            put_rec(P, rec, nr++, tick, ZRK_MSG_LAUNCH_SUCCESSFUL, 0, m_id, l_id, m_id, H.sid()[tg], V[0], V[1], V[2]);
                    } else {
                        mb_type[par * Q + q] = MB_CANCEL;
                        counts[ZRK_MSG_LAUNCH_CANCELLED]++;
                        if (logged(P, ZRK_MSG_LAUNCH_CANCELLED))
                            put_rec(P, rec, nr++, tick, ZRK_MSG_LAUNCH_CANCELLED, 0, m_id, l_id, m_id, why, 0, 0, 0);
                    }
                } else if (ty == MB_SUCCESS) {  // :103-124
                    const int m = arg, m_id = c_m_id[m], tid = H.sid()[m_target[m]];
                    counts[ZRK_MSG_LAUNCHED_MISSILE]++;
                    counts[ZRK_MSG_NEW_MISSILE]++;
                    if (logged(P, ZRK_MSG_LAUNCHED_MISSILE))
                        put_rec(P, rec, nr++, tick, ZRK_MSG_LAUNCHED_MISSILE, 0, l_id, ZRK_CCP_ID, m_id, tid, 0, 0, 0);
                    if (logged(P, ZRK_MSG_NEW_MISSILE))
                        put_rec(P, rec, nr++, tick, ZRK_MSG_NEW_MISSILE, 0, l_id, ZRK_NONE, m_id, ZRK_NONE, 0, 0, 0);
                    ZRK_COLD_I(ln_m)[hdr[H_LN]++] = m;
                } else if (ty == MB_CANCEL) {  // :125-128
                    st[l * MP + (st_n[l]++)] = arg;
                } else {  // MissileCountRequestMessage (:129-138)
                    counts[ZRK_MSG_MISSILE_COUNT_RESPONSE]++;
                    if (logged(P, ZRK_MSG_MISSILE_COUNT_RESPONSE))
                        put_rec(P, rec, nr++, tick, ZRK_MSG_MISSILE_COUNT_RESPONSE, 0, l_id, ccp_id, ZRK_NONE, st_n[l], 0, 0, 0);
                    // CCP.get_current_missile_launcher_capacity (CCP.py:139-147), same tick
                    for (int c = 0; c < nC; c++)
                        if (ZRK_CST_I(c_cl_launcher)[c] == l) ZRK_COLD_I(cap)[c] = st_n[l];
                    int fc = 0;
                    for (int c = 0; c < nC; c++) {
                        int d = ZRK_COLD_I(cap)[c] - ZRK_COLD_I(launched)[c];
                        fc += d > 0 ? d : 0;
                    }
                    hdr[H_FREECAP] = fc;
                }
            }
        }
        if (has_ccp) {
            if (!hdr[H_CCP_INIT]) {  // send_request_msg_to_ml_capacity (CCP.py:122-137)
                for (int c = 0; c < nC; c++) {
                    ZRK_COLD_I(cap)[c] = 0;
                    ZRK_COLD_I(launched)[c] = 0;
                    hdr[H_FREECAP] = 0;
                    counts[ZRK_MSG_MISSILE_COUNT_REQUEST]++;
                    if (logged(P, ZRK_MSG_MISSILE_COUNT_REQUEST))
                        put_rec(P, rec, nr++, tick, ZRK_MSG_MISSILE_COUNT_REQUEST, 0, ccp_id, ZRK_CST_I(c_cl_id)[c], ZRK_NONE, ZRK_NONE, 0, 0, 0);
                    const int q = hdr[H_MBN0 + par]++;
                    mb_type[par * Q + q] = MB_COUNTREQ;
                    mb_l[par * Q + q] = ZRK_CST_I(c_cl_launcher)[c];
                    mb_a[par * Q + q] = 0;
                }
                hdr[H_CCP_INIT] = 1;
            }
            // check_if_missiles_launched + add_missile (CCP.py:161-169, :102-109)
            const int n_ln = hdr[H_LN];
            int *me_key = ZRK_COLD_I(me_key);
            for (int i = 0; i < n_ln; i++) {
                const int m = ZRK_COLD_I(ln_m)[i], m_id = c_m_id[m], tid = H.sid()[m_target[m]];
                int e, n_me = hdr[H_NME];
                for (e = 0; e < n_me; e++)
                    if (me_key[e] == m_id) break;
                if (e == n_me) hdr[H_NME] = n_me + 1;
                me_key[e] = m_id;
                ZRK_COLD_I(me_ref)[e] = ZRK_COLD_I(m_slot)[m];
                ZRK_COLD_I(me_upd)[e] = tick;
                ZRK_COLD_D(me_upds)[e] = now_s;
                ZRK_COLD_I(me_tid)[e] = tid;
                H.own()[ZRK_COLD_I(m_slot)[m]] = TEAM + e;
                int t, n_te = hdr[H_NTE];
                for (t = 0; t < n_te; t++)
                    if (H.te_key()[t] == tid) break;
                if (t == n_te) {
                    hdr[H_FAULT] |= ZRK_FAULT_KEY_ERROR_CCP107;
                    hdr[H_FAULT_TICK] = tick;
                    break;
                }
                H.te_mid()[t] = m_id;
            }
        }
    }
    return nr;
}

// ------------------------------------------------------------------ the kernel

// resident CTAs per SM the register allocation is bounded for
#ifndef ZRK_MINB_SMALL
#define ZRK_MINB_SMALL 4
#endif
#ifndef ZRK_MINB_256
#define ZRK_MINB_256 2
#endif
#ifndef ZRK_MINB_TINY
#define ZRK_MINB_TINY 8
#endif
constexpr int min_blocks(int threads) { return threads <= 32 ? ZRK_MINB_TINY : threads <= 64 ? 8 : (threads <= 128 ? ZRK_MINB_SMALL : (threads == 256 ? ZRK_MINB_256 : 1)); }

#ifdef ZRK_MAXNREG
#define ZRK_KERNEL_ATTR __maxnreg__(ZRK_MAXNREG)
#else
#define ZRK_KERNEL_ATTR __launch_bounds__(TEAM *TPC, min_blocks(TEAM *TPC))
#endif
// MPB: missile capacity bucket of a warp team (Layout::MP as a compile-time constant, so that the
// scratch and track-entry offsets below are immediates); 0 = offsets from the Layout at run time.
// target_tick < 0: run n_ticks ticks of every scenario (quiet streaks included, in this kernel).
// target_tick >= 0 and pending != null (warp teams, "two-kernel stepping"): this launch runs only the ticks the streak kernel
// (zrk_streak_kernel below) cannot — of the scenarios flagged CF_GENERAL, until each of them is fit for a
// streak again or has reached target_tick — and counts in *pending the scenarios still short of target_tick.
// work (two-kernel stepping, may be null): the scenarios to look at, n = b_end - b_first of them, instead of
// [b_first, b_end) — work_n (may be null): where the device keeps the length of that list, when the host only knows an
// upper bound of it; work_out (may be null): the scenarios still pending are appended here, at the index
// the pending counter hands out.
template <int TEAM, int TPC, bool FAST, int MPB>
__global__ void ZRK_KERNEL_ATTR zrk_tick_kernel(const Params P, int n_ticks, int b_first, int b_end, int target_tick, int *pending,
                                                const int *work, int *work_out, const int *work_n) {
    ZRK_SMEM_DECL;
    const Layout &o = P.lo;
    constexpr ScratchOff XS = scratch_off(TEAM, MPB ? MPB : 2);
    constexpr MeOff XM = me_off(TEAM, MPB ? MPB : 2);
#define ZRK_X(name) (MPB ? XS.name - XS.bytes : o.name)   /* scratch offset from the dyn base */
#define ZRK_ME_I(name) ((int *)(sm + (MPB ? XM.name : o.name)))
#define ZRK_ME_D(name) ((double *)(sm + (MPB ? XM.name : o.name)))
    const int MP = MPB ? MPB : o.MP, Q = o.Q, MP8 = (MP + 7) & ~7;
    Team<TEAM, TPC> tm;
    // (a warp team's index is the same in all its lanes: say so, the addresses built from it can then be uniform)
    tm.tic = TEAM == 32 ? __shfl_sync(ZRK_FULL, (int)(threadIdx.x / TEAM), 0) : (int)(threadIdx.x / TEAM);
    tm.lane = threadIdx.x % TEAM;
    const int lane = tm.lane;
    int b = b_first + blockIdx.x * TPC + tm.tic;  // this launch covers scenarios [b_first, b_end)
    if (b >= b_end) return;  // whole team leaves together
    const bool two = TEAM == 32 && target_tick >= 0 && pending != nullptr;
    // (target_tick >= 0 without a pending counter: "finish" — run the listed scenarios up to target_tick
    // in this one launch, quiet streaks included; the host's way to end the tail of a step)
    const bool finish = TEAM == 32 && target_tick >= 0 && pending == nullptr;
    if ((two || finish) && work) {
        if (work_n && b - b_first >= *work_n) return;  // (the list is shorter than the grid that was sized for it)
        b = work[b - b_first];
    }
    if (two) {  // is there anything for this kernel to do?  (three words of the scenario's header, before any staging)
        const unsigned char *g = P.dyn + (size_t)b * o.dyn_bytes;
        const int *gh = (const int *)(g + Hot<TEAM>::O.hdr);
        const int lim = ((const int *)(P.cst + (size_t)b * o.cst_bytes + o.c_hdr))[C_TICK_LIMIT];
        if (gh[H_TICK] >= target_tick || (gh[H_FAULT] & 63) || (lim >= 0 && gh[H_TICK] >= lim)) return;
        if (!(((const int *)(g + Hot<TEAM>::O.counts))[CW_FLAG] & CF_GENERAL)) {  // the streak kernel's turn
            if (lane == 0) {
                const int k = atomicAdd(pending, 1);
                if (work_out) work_out[k] = b;
            }
            return;
        }
    }
    unsigned char *sm = zrk_smem + (size_t)tm.tic * o.team_smem_bytes + (MPB ? XS.bytes : o.scratch_bytes);  // the dyn blob; scratch in front of it
    Hot<TEAM> H;
    H.sm = sm;
    tm.wscr = (int *)(sm + ZRK_X(x_wscr));
    tm.flip = 0;
    double2 *lxy = (double2 *)(sm + ZRK_X(x_lxy)), *lzd = (double2 *)(sm + ZRK_X(x_lzd));
    float *lf = (float *)(sm + ZRK_X(x_lf));
    int *ghead = (int *)(sm + ZRK_X(x_ghead));
    // association certificate (teams up to 256 slots: the pair loop is O(slots) per lane)
    constexpr bool CERT = FAST && TEAM <= 256;
    int *const c_in = H.c_in(), *const c_out = H.c_out(), *const c_flag = H.c_flag(), *const c_refc = H.c_refc();
    const bool use_cert = CERT && P.cert && (two || finish || n_ticks >= 4);  // (a short launch would only pay for the upkeep)

    // ---- prologue: HBM -> shared memory, one bulk copy per blob
    {
        void *mbar = sm + ZRK_X(x_mbar);
        stage_in(sm, P.dyn + (size_t)b * o.dyn_bytes, o.dyn_bytes, P.cst + (size_t)b * o.cst_bytes, o.cst_bytes, mbar,
                 lane == 0);
        tm.sync();  // the barrier is initialised before anyone polls it
        stage_wait(mbar);
    }
    tm.sync();

    // (a launch that does not keep the certificate up leaves word that the windows are void)
    if (CERT && !use_cert && lane == 0) *c_flag |= CF_WINDOWS;
    const int *chdr = ZRK_CST_I(c_hdr);
    const int nR = chdr[C_NR], nC = chdr[C_NC], T = o.T;
    const int ccp_id = chdr[C_CCP_ID];
    const int tick_limit = chdr[C_TICK_LIMIT];
    const bool has_ccp = ccp_id != ZRK_NONE;
    zrk_msg *rec = P.rec + (size_t)b * P.log_cap;
    const double *noise = P.noise ? P.noise + (size_t)b * P.noise_stride : nullptr;
    int *const hdr = H.hdr();

    double now_next = tick_seconds(P, hdr[H_TICK]);
    // row of the noise table for this tick: (tick % period), advanced incrementally
    int noise_w = noise ? hdr[H_TICK] % P.noise_period : 0;
    const int noise_row = P.noise_radars * P.noise_objects * 3;  // doubles per tick row
    int streak_wait = 0;  // ticks to go before the next attempt at a quiet streak (after one that ended at once)
    if (two || finish) n_ticks = target_tick - hdr[H_TICK];
    int *const c_removed = H.c_removed();  // slots the last general tick removed (bit mask, warp teams)
    for (int it = 0; it < n_ticks; it++) {
        // Is the next tick one for a quiet streak (in the streak kernel under two-kernel stepping, in place otherwise)?
        bool to_streak = false;
        if (CERT && use_cert && has_ccp) {
            const int prv0 = (hdr[H_TICK] & 1) ^ 1;
            bool calm = hdr[H_DETN0 + prv0] == 0 && hdr[H_MBN0 + prv0] == 0 && hdr[H_CCP_INIT] && !(hdr[H_FAULT] & 63) &&
                        !(P.log_types & ZRK_LOG_BULK);
            if (calm) {  // ... and no object is waiting to be born or was born on the last tick (a streak would end at once)
                const int f = lane < hdr[H_NSLOTS] ? H.sflags()[lane] : 0;
                calm = !tm.any((f & F_PENDING) != 0 || ((f & F_ALIVE) && !(f & F_HASPREV)));
            }
            if (two)
                to_streak = calm && it > ((*c_flag >> CF_HOLD_SHIFT) & 127);
            else
                to_streak = --streak_wait < 0 && calm;
            // Certificate upkeep, on the state the last tick left (positions of tick H_TICK - 1): after an
            // event every window is rebuilt; after tracks changed hands the reference counts; after a removal
            // the pairs with the removed slots are merged in.  Only a streak looks at the windows, so the upkeep
            // waits until one is about to begin (a removal that is not merged in on the spot costs a full rebuild then).
            const int cf = *c_flag & (CF_WINDOWS | CF_TRACKS);
            const unsigned removed_carry = (unsigned)*c_removed;
            if (ZRK_RARE(cf != 0 || removed_carry != 0)) {
                if (to_streak) {
                    const int tprev = hdr[H_TICK] - 1, ns = hdr[H_NSLOTS];
                    if (cf & CF_WINDOWS) {
                        cert_recompute<TEAM>(sm, ZRK_ME_I(me_ref), lane, ns, tprev, tprev + 1, hdr[H_NTE], hdr[H_NME], P.dt_s,
                                             P.noise_bound, c_in, c_out, c_refc, (float4 *)lxy, (float4 *)lzd);
                    } else {
                        if (cf) cert_recount<TEAM>(sm, ZRK_ME_I(me_ref), lane, hdr[H_NTE], hdr[H_NME], c_refc);
                        if (removed_carry)
                            cert_removed<TEAM>(sm, lane, lane < ns ? H.sflags()[lane] : 0, removed_carry, tprev, P.dt_s, P.noise_bound,
                                               c_in, c_out);
                    }
                    tm.sync();  // everybody has read the flags
                    if (lane == 0) {
                        *c_flag &= ~(CF_WINDOWS | CF_TRACKS);
                        *c_removed = 0;
                    }
                } else if (removed_carry) {
                    tm.sync();
                    if (lane == 0) {
                        *c_flag |= CF_WINDOWS;
                        *c_removed = 0;
                    }
                }
            }
        }
        if (two) {  // hand the scenario back to the streak kernel as soon as its next tick is one for it
            if (to_streak) {
                if (lane == 0) *c_flag &= ~CF_GENERAL;
                break;
            }
        } else if (to_streak) {
            // (the lean loop's records of a warp team go where the fp32 images, the claims and the detection list of a
            // general tick are: ZRK_RREC_DOUBLES x 8 bytes, free between ticks)
            static_assert(XS.x_wscr - XS.x_lf >= ZRK_RREC_DOUBLES * 8 || TEAM != 32 || !MPB, "no room for the lean loop's records");
            const int nq = quiet_streak<TEAM, MPB>(P, sm, sm + o.dyn_bytes, lane, b, n_ticks - it, noise_w, (float4 *)lxy, (float4 *)lzd,
                                                   (double *)tm.wscr, (TEAM == 32 && MPB) ? (double *)lf : nullptr);
            if (nq) {
                it += nq;
                if (it >= n_ticks) break;
                if (noise) noise_w = (noise_w + nq) % P.noise_period;
                now_next = tick_seconds(P, hdr[H_TICK]);
            } else {
                streak_wait = 3;
            }
        }
        const int tick = hdr[H_TICK];
        if (hdr[H_FAULT] & 63) break;  // the reference run aborted here
        if (tick_limit >= 0 && tick >= tick_limit) break;
        const int par = tick & 1, prv = par ^ 1;
        const double now_s = now_next;          // to_seconds(time.get_time())
        now_next = tick_seconds(P, tick + 1);  // loaded a tick ahead: its latency is never waited for
        const int n_slots = hdr[H_NSLOTS];
        const int det_prev = hdr[H_DETN0 + prv];
        const int mb_prev = hdr[H_MBN0 + prv];
        int nrec = hdr[H_NREC];  // records written so far; kept uniform across the team
        // (the per-tick lists of this parity were emptied at the end of the previous tick; nobody
        // writes a header word read above before the barrier that ends AirEnv.step)

        // ============ AirEnv.step (AirEnv.py:26-53)
        int fl = 0, my_id = 0;
        double px = 0, py = 0, pz = 0;
        double nz0 = 0, nz1 = 0, nz2 = 0;  // radar 0's noise for this slot, fetched early
        bool odd = false;  // a slot whose tracks do not follow the certificate's model this tick (not yet born / first tick)
        bool was_alive = false;
        if (lane < n_slots) {
            fl = H.sflags()[lane];
            odd = (fl & F_PENDING) != 0;
            my_id = H.sid()[lane];
            if (noise && nR) {
                const double *nz = noise + (size_t)noise_w * noise_row + (fl >> F_STATIC_SHIFT) * 3;
                nz0 = ZRK_LD_EARLY(nz);
                nz1 = ZRK_LD_EARLY(nz + 1);
                nz2 = ZRK_LD_EARLY(nz + 2);
            }
            was_alive = (fl & F_ALIVE) != 0;
            if (ZRK_RARE((fl & F_ALIVE) && det_prev)) {  // :33-40 removal by id
                const int *dm = ZRK_COLD_I(det_mid) + prv * MP, *dt = ZRK_COLD_I(det_tid) + prv * MP;
                for (int i = 0; i < det_prev; i++) {
                    int tid = dt[i];
                    if (my_id == dm[i] || (tid != ZRK_NONE && my_id == tid)) fl &= ~F_ALIVE;
                }
            }
            if (ZRK_RARE(fl & F_PENDING)) {  // :42-43 NEW_MISSILE of the previous tick
                int m = (fl >> F_STATIC_SHIFT) - T;
                if (ZRK_COLD_I(m_birth)[m] == tick) {
                    fl = (fl & ~F_PENDING) | F_ALIVE;
                }
            }
            if (fl & F_ALIVE) {  // AirObject.step (AirObject.py:39-42)
                const double2 oxy = H.pxy()[lane], ozs = H.pzs()[lane], szt = H.tzt()[lane];
                const double t0 = szt.y;
                if (t0 != now_s) {
                    H.qxy()[lane] = oxy;
                    H.qzw()[lane].x = ozs.x;
                    fl |= F_HASPREV;
                } else {
                    fl &= ~F_HASPREV;
                }
                const double dtm = now_s - t0;
                const double2 sxy = H.txy()[lane], wxy = H.vxy()[lane];
                px = sxy.x + wxy.x * dtm;  // Trajectory.get_pos (AirObject.py:23-25)
                py = sxy.y + wxy.y * dtm;
                pz = szt.x + H.qzw()[lane].y * dtm;
                H.pxy()[lane] = make_double2(px, py);
                H.pzs()[lane] = make_double2(pz, ozs.y);
            }
            H.sflags()[lane] = fl;
        }
        const bool alive = (fl & F_ALIVE) != 0;
        odd = odd || (alive && !(fl & F_HASPREV));
        {  // activity counter: alive slots summed over the ticks started (the "active-slot fraction" of a run)
            const unsigned live = __ballot_sync(ZRK_FULL, alive);
            if (TEAM == 32) {
                if (lane == 0) hdr[H_ALIVE] += __popc(live);
            } else if ((lane & 31) == 0 && live) {
                atomicAdd(&hdr[H_ALIVE], __popc(live));
            }
        }
        // A tick on which the certificate's model does not hold: launcher / CCP bookkeeping, a missile
        // between launch and its first step, an object on its first tick.  (uniform)  A removal is not one:
        // the removed objects' tracks stay where they were from this very tick on, so the windows are rebuilt
        // now and the tick goes on as usual.
        bool event = true;
        if (use_cert) {
            event = mb_prev != 0 || (has_ccp && !hdr[H_CCP_INIT]);
            event = tm.any(odd) || event;  // (a team larger than a warp: this is the barrier that ends AirEnv.step)
            if (TEAM == 32) tm.sync();
        } else {
            tm.sync();
        }

        // Missile.step, status 'active' (Missile.py:162-193), in slot order
        {
            const bool msl = alive && (fl & F_MISSILE);
            if ((TEAM == 32 || chdr[C_NM] > 0) && tm.any(msl)) {  // (uniform: a large team without missiles skips the barrier)
                int det = 0, det_tid = ZRK_NONE;
                if (msl) {
                    const int m = (fl >> F_STATIC_SHIFT) - T;
                    const int tg = ZRK_COLD_I(m_target)[m];
                    const double2 txy = H.pxy()[tg], tzs = H.pzs()[tg];
                    double ex = txy.x - px, ey = txy.y - py, ez = tzs.x - pz;
                    double dist = sqrt(dot3(ex, ey, ez, ex, ey, ez));
                    if (dist <= ZRK_CST_D(c_m_radius)[m]) {
                        det = 1;
                        det_tid = H.sid()[tg];
                    } else {
                        double per = ZRK_COLD_D(m_period)[m] - P.dt_s;
                        ZRK_COLD_D(m_period)[m] = per;
                        if (per <= 0) det = 2;
                    }
                    if (det) ZRK_COLD_I(m_status)[m] = ST_DETONATED;
                }
                const bool lp = logged(P, ZRK_MSG_MISSILE_POS), ld = logged(P, ZRK_MSG_MISSILE_DETONATE);
                int n_msl, n_det;
                const int r_msl = tm.rank(msl, n_msl);
                const int r_det = tm.rank(det != 0, n_det);
                int ri = nrec + (lp ? r_msl : 0) + (ld ? r_det : 0);
                if (msl && lp) put_rec(P, rec, ri++, tick, ZRK_MSG_MISSILE_POS, 0, my_id, ZRK_NONE, my_id, ZRK_NONE, 0, 0, 0);
                if (det) {
                    if (ld)
                        put_rec(P, rec, ri, tick, ZRK_MSG_MISSILE_DETONATE, det == 2, my_id, ZRK_NONE, my_id, det_tid, 0, 0, 0);
                    ZRK_COLD_I(det_mid)[par * MP + r_det] = my_id;
                    ZRK_COLD_I(det_tid)[par * MP + r_det] = det_tid;
                    ZRK_COLD_I(det_flag)[par * MP + r_det] = det == 2;
                }
                nrec += (lp ? n_msl : 0) + (ld ? n_det : 0);
                if (lane == 0) {
                    hdr[H_DETN0 + par] = n_det;
                    H.counts()[ZRK_MSG_MISSILE_POS] += (unsigned)n_msl;
                    H.counts()[ZRK_MSG_MISSILE_DETONATE] += (unsigned)n_det;
                }
            }
            if (lane == 0) H.counts()[ZRK_MSG_ACTIVE_OBJECTS] += 1;
        }
        unsigned *mrow = nullptr;
        if (P.masks) {
            mrow = P.masks + ((size_t)(tick % P.mask_ticks) * o.B + b) * (size_t)((1 + o.RP) * P.mask_words);
            unsigned bal = __ballot_sync(ZRK_FULL, alive);
            if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[lane >> 5] = bal;
        }

        // ============ SectorRadar.step for every radar (Radar.py:144-205)
        int rf = -1;
        {
            const double *rad_cur = ZRK_ME_D(rad_cur), *c_rad_pos = ZRK_CST_D(c_rad_pos), *c_rad_par = ZRK_CST_D(c_rad_par);
            for (int r = 0; r < nR; r++) {
                const double *par_r = c_rad_par + RAD_PAR_STRIDE * r;
                bool vis = false;
                if (alive) {
                    double caz = rad_cur[2 * r], cel = rad_cur[2 * r + 1];
                    double dx = px - c_rad_pos[3 * r], dy = py - c_rad_pos[3 * r + 1], dz = pz - c_rad_pos[3 * r + 2];
                    vis = radar_sees<FAST>(dx, dy, dz, par_r, caz, caz + par_r[3], cel, cel + par_r[4]);
                    if (vis) {  // smooth_objects (Radar.py:138-142)
                        if (noise) {
                            if (r) {
                                const double *nz = noise + (size_t)noise_w * noise_row +
                                                   ((size_t)r * P.noise_objects + (fl >> F_STATIC_SHIFT)) * 3;
                                nz0 = nz[0];
                                nz1 = nz[1];
                                nz2 = nz[2];
                            }
                            px += nz0;
                            py += nz1;
                            pz += nz2;
                        }
                        if (rf < 0) rf = r;
                    }
                }
                if (mrow) {
                    unsigned bal = __ballot_sync(ZRK_FULL, vis);
                    if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[(1 + r) * P.mask_words + (lane >> 5)] = bal;
                }
                // :193-203 one DESTROYED_MISSILE per MISSILE_DETONATE of the previous tick
                if (ZRK_RARE(det_prev && logged(P, ZRK_MSG_DESTROYED_MISSILE) && lane < det_prev))
                    put_rec(P, rec, nrec + r * det_prev + lane, tick, ZRK_MSG_DESTROYED_MISSILE,
                            ZRK_COLD_I(det_flag)[prv * MP + lane], ZRK_CST_I(c_rad_id)[r], ZRK_CCP_ID,
                            ZRK_COLD_I(det_mid)[prv * MP + lane], ZRK_NONE, 0, 0, 0);
            }
            if (nR && logged(P, ZRK_MSG_DESTROYED_MISSILE)) nrec += nR * det_prev;
        }
        H.srf()[lane] = rf;
        if (alive && rf >= 0) {
            H.pxy()[lane] = make_double2(px, py);
            H.pzs()[lane].x = pz;
        }
        // anything for lane 0 to do alone this tick?  (uniform)
        const bool serial = mb_prev != 0 || (has_ccp && !hdr[H_CCP_INIT]);
        if (lane == 0) {
            hdr[H_NREC] = nrec;
            if (nR) {
                H.counts()[ZRK_MSG_ALL_OBJECTS] += nR;
                H.counts()[ZRK_MSG_FOUND_OBJECTS] += nR;
                H.counts()[ZRK_MSG_DESTROYED_MISSILE] += nR * det_prev;
            }
        }
        // What this tick means for the certificate (the upkeep itself runs before the next tick, see the top of
        // the loop): an event tick voids every window; a removal freezes the removed slots where they are.
        if (CERT && use_cert) {
            if (TEAM == 32) {
                if (ZRK_RARE(det_prev)) {
                    const unsigned rm = __ballot_sync(ZRK_FULL, was_alive && !alive);
                    if (lane == 0) *c_removed = (int)rm;
                }
                if (event && lane == 0) *c_flag |= CF_WINDOWS;
            } else if ((event || det_prev) && lane == 0) {
                *c_flag |= CF_WINDOWS;
            }
        }
        tm.sync();

        // ============ lane 0: sector sweep, MissileLauncher.step, CCP bookkeeping
        if (lane == 0) {
            double *rad_cur = ZRK_ME_D(rad_cur);
            const double *c_rad_par = ZRK_CST_D(c_rad_par);
            for (int r = 0; r < nR; r++) {  // move_to_next_sector_circular (Radar.py:96-117)
                const double *q = c_rad_par + RAD_PAR_STRIDE * r;
                if (q[10] != 0.0) continue;  // the start sector is a fixed point of the sweep
                double caz = rad_cur[2 * r], cel = rad_cur[2 * r + 1];
                radar_move(q, caz, cel);
                rad_cur[2 * r] = caz;
                rad_cur[2 * r + 1] = cel;
            }
        }
        if (ZRK_RARE(serial)) {
            if (lane == 0) {
                nrec = serial_section<TEAM>(P.self, sm, rec, tick, now_s, mb_prev, nrec);
                hdr[H_NREC] = nrec;
            }
            tm.sync();
            nrec = hdr[H_NREC];
        }

        // ============ CombatControlPoint.step, detections (CCP.py:406-429)
        if (has_ccp && !(hdr[H_FAULT] & 63)) {
            int n_te = hdr[H_NTE];
            const int n_me = hdr[H_NME];
            {
            bool structural = false;  // this tick's rounds went beyond re-stamping own tracks
            // With the quiet streak carrying the ordinary ticks inside this kernel, a warp team comes here on
            // the rare eventful tick only, and what it costs then is the code it has to fetch, not the
            // arithmetic: it scans with link_exact (a small loop) rather than the unrolled fp32 sweep.  (Under
            // two-kernel stepping every warp of this kernel runs general ticks: the sweep's code stays cached.)
            const bool scan_fast = FAST && !(TEAM == 32 && use_cert && !two && !finish);  // (finish: the scenarios left are the ones with general tick after general tick)
            // what link_object measures to, per entry: prev_pos of the tracked object (or pos
            // for a missile whose prev_pos is None) and d_t = now - upd_time
            // the hash grid pays off once a scan is long; warp-sized scenarios scan their few entries directly
            const bool want_grid = scan_fast && TEAM > 32 && n_te + n_me >= 48;
            bool bad_image = false;  // an entry outside the fp32-safe range: the team scans exactly this tick
            if (want_grid) ghead[lane] = -1;
            if (lane < n_te) {
                const int ref = H.te_ref()[lane], upd = H.te_upd()[lane];
                const double2 a = H.qxy()[ref], c = H.qzw()[ref];
                const bool skip = upd == tick;
                lxy[lane] = make_double2(skip ? ZRK_FAR : a.x, a.y);
                lzd[lane] = make_double2(c.x, now_s - H.te_upds()[lane]);
                if (scan_fast) {
                    lf_store<TEAM>(lf, lane, skip ? 1e30f : (float)a.x, (float)a.y, (float)c.x);
                    bad_image = !skip & !link_f32_ok((float)a.x, (float)a.y, (float)c.x);
                }
            } else if (scan_fast && TEAM == 32) {
                lf_store<TEAM>(lf, lane, 1e30f, 0.f, 0.f);  // pad: measures inf to everything
            }
            if (lane < n_me) {
                const int ref = ZRK_ME_I(me_ref)[lane], upd = ZRK_ME_I(me_upd)[lane];
                const bool hp = (H.sflags()[ref] & F_HASPREV) != 0;
                const double2 a = hp ? H.qxy()[ref] : H.pxy()[ref];
                const double cz = hp ? H.qzw()[ref].x : H.pzs()[ref].x;
                const bool skip = upd == tick;
                lxy[TEAM + lane] = make_double2(skip ? ZRK_FAR : a.x, a.y);
                lzd[TEAM + lane] = make_double2(cz, now_s - ZRK_ME_D(me_upds)[lane]);
                if (scan_fast) {
                    lf_store<TEAM>(lf, TEAM + lane, skip ? 1e30f : (float)a.x, (float)a.y, (float)cz);
                    bad_image = bad_image | (!skip & !link_f32_ok((float)a.x, (float)a.y, (float)cz));
                }
            } else if (scan_fast && TEAM == 32 && lane < MP8) {
                lf_store<TEAM>(lf, TEAM + lane, 1e30f, 0.f, 0.f);
            }
            bool grid_ok = false, any_bad = false;
            // processing order: FOUND_OBJECTS messages in radar order, objects in list order,
            // each id once -> (first radar that saw it, slot)
            int n_det = 0, myrank = -1;
            bool any_seen = false;  // OR of bad_image over the team, taken along with the first ranks
            if (TEAM == 32) {
                for (int r = 0; r < nR; r++) {
                    int tot;
                    const int ex = tm.rank(rf == r, tot);
                    if (rf == r) myrank = n_det + ex;
                    n_det += tot;
                }
                any_seen = tm.any(bad_image);
            } else {
                for (int r = 0; r < nR; r += 2) {  // two radars per barrier
                    int e0, e1, t0, t1;
                    bool orv;
                    tm.rank2(rf == r, rf == r + 1, e0, e1, t0, t1, r == 0 && bad_image, orv);
                    if (r == 0) any_seen = orv;
                    if (rf == r) myrank = n_det + e0;
                    if (rf == r + 1) myrank = n_det + t0 + e1;
                    n_det += t0 + t1;
                }
            }
            int *detlist = (int *)(sm + ZRK_X(x_detlist));
            if (myrank >= 0) detlist[myrank] = lane;
            if (scan_fast) any_bad = nR > 0 ? any_seen : false;  // (no radar: no detection, nothing scans)
            if (want_grid && nR > 0) {
                // the barrier of the first ranks also published the images and the reset bucket heads
                grid_ok = !any_bad;
                if (grid_ok) {  // entries link_object skips (image x = 1e30) stay out of the grid
                    if (lane < n_te && lf[4 * lane] < 1e29f) {
                        float *q = lf + 4 * lane;
                        q[3] = __int_as_float(atomicExch(&ghead[grid_bucket<TEAM>(grid_cell(q[0]), grid_cell(q[1]))], lane));
                    }
                    if (lane < n_me && lf[4 * (TEAM + lane)] < 1e29f) {
                        float *q = lf + 4 * (TEAM + lane);
                        q[3] = __int_as_float(atomicExch(&ghead[grid_bucket<TEAM>(grid_cell(q[0]), grid_cell(q[1]))], TEAM + lane));
                    }
                }
            }
            tm.sync();
            const int i = lane;
            int slot = 0, seed = -1;
            // id of the detected object and of the radar that reported it: only messages and new tracks need them
            auto det_id = [&]() { return H.sid()[slot]; };
            auto det_radar = [&]() { return ZRK_CST_I(c_rad_id)[H.srf()[slot]]; };
            bool seed_own = false;
            double dpx = 0, dpy = 0, dpz = 0, dspeed = 0;
            bool unresolved = i < n_det;
            if (unresolved) {
                slot = detlist[i];
                const double2 a = H.pxy()[slot], c = H.pzs()[slot];
                dpx = a.x;
                dpy = a.y;
                dpz = c.x;
                dspeed = c.y;
                seed = H.own()[slot];
                // (entries are never removed — CCP's dicts only grow, CCP.py:88-109 — so a stored seed stays in range)
                // does the seed entry still refer to this very slot?  (then distinct slots have distinct seeds)
                if (seed >= 0) seed_own = (seed >= TEAM ? ZRK_ME_I(me_ref)[seed - TEAM] : H.te_ref()[seed]) == slot;
            }
            const bool coords_ok = scan_fast && !any_bad && link_f32_ok((float)dpx, (float)dpy, (float)dpz);
            int lo_rank = 0;
            for (;;) {
                int cls = 0, cand = -1, claim = -1;
                if (unresolved) {
                    if (scan_fast) {
                        bool unsure = !coords_ok;
                        if (coords_ok)
                            link_fast<TEAM>(ghead, grid_ok, lf, lxy, lzd, n_te, n_me, seed, P.range_step, dpx, dpy, dpz, dspeed, cls, cand, unsure);
                        if (unsure) link_exact<TEAM>(lxy, lzd, n_te, n_me, P.range_step, dpx, dpy, dpz, dspeed, cls, cand);
                    } else {
                        link_exact<TEAM>(lxy, lzd, n_te, n_me, P.range_step, dpx, dpy, dpz, dspeed, cls, cand);
                    }
                    if (cls == 1)
                        claim = cand;
                    else if (cls == 2)
                        claim = TEAM + cand;
                    else {  // add_target overwrites an existing key in place (CCP.py:92)
                        const int *te_key = H.te_key();
                        const int did_ = det_id();
                        for (int e = 0; e < n_te; e++)
                            if (te_key[e] == did_) {
                                claim = e;
                                break;
                            }
                    }
                }
                // Quiet tick (the common case): every detection re-acquired the track that followed its
                // own slot, so all claims are distinct, nothing is appended, and with no launcher
                // capacity left nothing can be launched: the whole round reduces to stamping the entries.
                if (lo_rank == 0 && hdr[H_FREECAP] <= 0) {
                    const bool quiet = !unresolved || (cls != 0 && claim == seed && seed_own);
                    bool upd_msg = false, noisy;
                    int n_upd = 0, r_upd = 0;
                    if (TEAM == 32) {
                        noisy = tm.any(!quiet);
                        if (!noisy) {
                            upd_msg = unresolved && cls == 1 && H.te_follow()[cand] != 0;
                            r_upd = tm.rank(upd_msg, n_upd);
                        }
                    } else {  // one barrier: the vote rides on the rank's
                        upd_msg = quiet && unresolved && cls == 1 && H.te_follow()[cand] != 0;
                        r_upd = tm.rank_or(upd_msg, n_upd, !quiet, noisy);
                    }
                    if (!noisy) {
                        if (unresolved) {
                            if (cls == 1) {  // old_target, already followed or no missile to send (CCP.py:332-359)
                                if (upd_msg && logged(P, ZRK_MSG_CCP_UPDATE_TARGET))
                                    put_rec(P, rec, nrec + r_upd, tick, ZRK_MSG_CCP_UPDATE_TARGET, 0, ccp_id, det_radar(), det_id(),
                                            H.te_mid()[cand], 0, 0, 0);
                                H.te_upd()[cand] = tick;
                                H.te_upds()[cand] = now_s;
                            } else {  // old_rocket (CCP.py:361-366)
                                ZRK_ME_I(me_upd)[cand] = tick;
                                ZRK_ME_D(me_upds)[cand] = now_s;
                            }
                        }
                        if (logged(P, ZRK_MSG_CCP_UPDATE_TARGET)) nrec += n_upd;
                        if (lane == 0) H.counts()[ZRK_MSG_CCP_UPDATE_TARGET] += (unsigned)n_upd;
                        break;
                    }
                }
                structural = true;
                // a detection is final once no earlier unresolved detection touches the same entry
                bool conflict;
                if (TEAM == 32) {
                    // lanes with my claim: claims are entry indices < 2 * TEAM = 64, compared bit by bit
                    const bool has = unresolved && claim >= 0;
                    unsigned same = __ballot_sync(ZRK_FULL, has);
#pragma unroll
                    for (int bit = 0; bit < 6; bit++) {
                        const unsigned bb = __ballot_sync(ZRK_FULL, (claim >> bit) & 1);
                        same &= ((claim >> bit) & 1) ? bb : ~bb;
                    }
                    conflict = has && (same & ((1u << lane) - 1u)) != 0;
                } else {
                    int *x_claim = (int *)(sm + ZRK_X(x_claim));
                    for (int x = lane; x < TEAM + MP; x += TEAM) x_claim[x] = ZRK_IMAX;
                    tm.sync();
                    if (unresolved && claim >= 0) atomicMin(&x_claim[claim], i);
                    tm.sync();
                    conflict = unresolved && claim >= 0 && x_claim[claim] != i;
                }
                const int istar = tm.min(conflict ? i : ZRK_IMAX);
                const bool fin = unresolved && i < istar;

                const int following = (fin && cls == 1) ? H.te_follow()[cand] : 0;
                const int old_mid = (fin && cls == 1) ? H.te_mid()[cand] : ZRK_NONE;
                const bool need = fin && (cls == 0 || (cls == 1 && !following));
                int ok = 0;
                // hdr[H_FREECAP] = sum over launchers of max(0, capacity - launched), kept by lane 0
                if (hdr[H_FREECAP] > 0 && tm.any(need)) {  // try_to_launch_missile (CCP.py:287-320), in order
                    int *x_need = (int *)(sm + ZRK_X(x_need));
                    x_need[i] = need;
                    tm.sync();
                    if (lane == 0) {
                        const int hi_rank = istar < n_det ? istar : n_det;
                        int *launched = ZRK_COLD_I(launched);
                        const int *cap = ZRK_COLD_I(cap);
                        const double *c_cl_pos = ZRK_CST_D(c_cl_pos);
                        for (int j = lo_rank; j < hi_rank; j++) {
                            if (!x_need[j]) continue;
                            const int sl = detlist[j];
                            double ox = H.pxy()[sl].x, oy = H.pxy()[sl].y, oz = H.pzs()[sl].x;
                            double min_dist = INFINITY;
                            int best = -1;
                            for (int c = 0; c < nC; c++) {
                                if (launched[c] < cap[c]) {
                                    double q0 = c_cl_pos[3 * c] - ox, q1 = c_cl_pos[3 * c + 1] - oy, q2 = c_cl_pos[3 * c + 2] - oz;
                                    double dist = sqrt((q0 * q0 + q1 * q1) + q2 * q2);
                                    if (dist < min_dist) {
                                        best = c;
                                        min_dist = dist;
                                    }
                                }
                            }
                            x_need[j] = best + 1;  // overwritten with the result: 0 = no launcher, c + 1
                            if (best >= 0) {
                                launched[best]++;
                                hdr[H_FREECAP]--;
                                const int q = hdr[H_MBN0 + par]++;
                                ZRK_COLD_I(mb_type)[par * Q + q] = MB_COMMAND;
                                ZRK_COLD_I(mb_l)[par * Q + q] = ZRK_CST_I(c_cl_launcher)[best];
                                ZRK_COLD_I(mb_a)[par * Q + q] = sl;
                            }
                        }
                    }
                    tm.sync();
                    if (need) ok = x_need[i];
                }
                // one message per finalised detection at most
                int mtype = -1;
                if (fin) {
                    if (ok)
                        mtype = ZRK_MSG_LAUNCH_COMMAND;
                    else if (cls == 1 && following)
                        mtype = ZRK_MSG_CCP_UPDATE_TARGET;
                }
                const bool wr = mtype >= 0 && logged(P, mtype);
                const bool app = fin && cls == 0 && claim < 0;
                int n_wr = 0, n_app, n_cmd, n_upd;
                int r_wr = 0;
                const int r_app = tm.rank(app, n_app);
                (void)tm.rank(mtype == ZRK_MSG_LAUNCH_COMMAND, n_cmd);
                (void)tm.rank(mtype == ZRK_MSG_CCP_UPDATE_TARGET, n_upd);
                if (P.log_types & (ZRK_TYPE_BIT(ZRK_MSG_LAUNCH_COMMAND) | ZRK_TYPE_BIT(ZRK_MSG_CCP_UPDATE_TARGET)))
                    r_wr = tm.rank(wr, n_wr);
                if (wr) {
                    if (mtype == ZRK_MSG_LAUNCH_COMMAND)
                        put_rec(P, rec, nrec + r_wr, tick, mtype, 0, ccp_id, ZRK_CST_I(c_cl_id)[ok - 1], det_id(), det_radar(), dpx, dpy, dpz);
                    else
                        put_rec(P, rec, nrec + r_wr, tick, mtype, 0, ccp_id, det_radar(), det_id(), old_mid, 0, 0, 0);
                }
                // Track journal (ZRK_LOG_TRACKS): every association that is not "the detection re-acquired the track
                // of its own slot" — from these and the event records the host replays the CCP's two dicts and rebuilds
                // the per-tick CCP_UPDATE_TARGET / DRAW_OBJECTS messages without their ever being stored (zrk_host.cu).
                if (logged(P, ZRK_REC_TRACK)) {
                    const bool trk = fin && !(cls != 0 && claim == seed && seed_own);
                    int n_trk;
                    const int r_trk = tm.rank(trk, n_trk);
                    if (trk)
                        put_rec(P, rec, nrec + n_wr + r_trk, tick, ZRK_REC_TRACK, cls, slot, cls == 0 ? (claim >= 0 ? claim : n_te + r_app) : cand,
                                det_id(), ok != 0, 0, 0, 0);
                    n_wr += n_trk;
                }
                if (fin) {
                    if (cls == 1) {  // old_target (CCP.py:332-359)
                        H.te_ref()[cand] = slot;
                        H.te_upd()[cand] = tick;
                        H.te_upds()[cand] = now_s;
                        if (!following) H.te_follow()[cand] = ok != 0;
                        H.own()[slot] = cand;
                        lxy[cand].x = ZRK_FAR;
                        if (scan_fast) lf_x<TEAM>(lf, cand) = 1e30f;
                    } else if (cls == 2) {  // old_rocket (CCP.py:361-366)
                        ZRK_ME_I(me_ref)[cand] = slot;
                        ZRK_ME_I(me_upd)[cand] = tick;
                        ZRK_ME_D(me_upds)[cand] = now_s;
                        H.own()[slot] = TEAM + cand;
                        lxy[TEAM + cand].x = ZRK_FAR;
                        if (scan_fast) lf_x<TEAM>(lf, TEAM + cand) = 1e30f;
                    } else {  // new_target + add_target (CCP.py:322-330, :88-93)
                        const int e = claim >= 0 ? claim : n_te + r_app;
                        H.te_key()[e] = det_id();
                        H.te_ref()[e] = slot;
                        H.te_upd()[e] = tick;
                        H.te_upds()[e] = now_s;
                        H.te_follow()[e] = ok != 0;
                        H.te_mid()[e] = ZRK_NONE;
                        H.own()[slot] = e;
                        lxy[e].x = ZRK_FAR;
                        if (scan_fast) lf_x<TEAM>(lf, e) = 1e30f;
                    }
                    unresolved = false;
                }
                nrec += n_wr;
                n_te += n_app;
                if (lane == 0) {
                    H.counts()[ZRK_MSG_LAUNCH_COMMAND] += (unsigned)n_cmd;
                    H.counts()[ZRK_MSG_CCP_UPDATE_TARGET] += (unsigned)n_upd;
                }
                if (istar == ZRK_IMAX) break;
                lo_rank = istar;
                tm.sync();  // entry updates of this round are visible to the next
            }
            if (CERT && structural && lane == 0) *c_flag |= CF_TRACKS;  // tracks changed hands or were created
            }
            tm.sync();

            // ============ send_objects_to_GUI (CCP.py:237-284)
            {
                const bool ldr = logged(P, ZRK_MSG_DRAW_OBJECTS);
                int mref = -1, tref = -1;
                bool dm = false, dt = false, bad = false;
                if (lane < n_me) {
                    mref = ZRK_ME_I(me_ref)[lane];
                    dm = H.srf()[mref] >= 0;
                }
                if (lane < n_te) {
                    tref = H.te_ref()[lane];
                    dt = H.srf()[tref] >= 0;
                    bad = dt && (H.sflags()[tref] & F_MISSILE);  // .type on a Missile (CCP.py:259)
                }
                int tot_m, tot_t, tot_o, ex_m, ex_t, ex_o, efault = ZRK_IMAX;
                bool dother;
                if (TEAM == 32) {
                    efault = tm.any(bad) ? tm.min(bad ? lane : ZRK_IMAX) : ZRK_IMAX;
                    dt = dt && lane < efault;
                    dother = efault == ZRK_IMAX && alive && rf < 0 && nR > 0;
                    ex_m = tm.rank(dm, tot_m);
                    ex_t = tm.rank(dt, tot_t);
                    ex_o = tm.rank(dother, tot_o);
                } else {  // one barrier for the three ranks and the fault vote
                    bool any_fault;
                    dother = alive && rf < 0 && nR > 0;
                    tm.rank3(dm, dt, dother, ex_m, ex_t, ex_o, tot_m, tot_t, tot_o, bad, any_fault);
                    if (any_fault) {  // rare: the entries before the faulting one are still drawn, nothing after it
                        efault = tm.min(bad ? lane : ZRK_IMAX);
                        dt = dt && lane < efault;
                        dother = false;
                        ex_t = tm.rank(dt, tot_t);
                        tot_o = 0;
                    }
                }
                if (ldr) {
                    if (dm)
                        put_rec(P, rec, nrec + ex_m, tick, ZRK_MSG_DRAW_OBJECTS, 1, ccp_id, ZRK_MANAGER_ID, H.sid()[mref], ZRK_ZUR,
                                H.pxy()[mref].x, H.pxy()[mref].y, H.pzs()[mref].x);
                    if (dt)
                        put_rec(P, rec, nrec + tot_m + ex_t, tick, ZRK_MSG_DRAW_OBJECTS, 1, ccp_id, ZRK_MANAGER_ID, H.sid()[tref],
                                (H.sflags()[tref] >> F_TTYPE_SHIFT) & 3, H.pxy()[tref].x, H.pxy()[tref].y, H.pzs()[tref].x);
                    if (dother)
                        put_rec(P, rec, nrec + tot_m + tot_t + ex_o, tick, ZRK_MSG_DRAW_OBJECTS, 0, ccp_id, ZRK_MANAGER_ID, my_id,
                                (fl & F_MISSILE) ? ZRK_ZUR : ((fl >> F_TTYPE_SHIFT) & 3), px, py, pz);
                    nrec += tot_m + tot_t + tot_o;
                }
                if (lane == 0) {
                    hdr[H_NTE] = n_te;
                    H.counts()[ZRK_MSG_DRAW_OBJECTS] += (unsigned)(tot_m + tot_t + tot_o);
                    if (efault != ZRK_IMAX) {
                        hdr[H_FAULT] |= ZRK_FAULT_ATTRIBUTE_ERROR_CCP259;
                        hdr[H_FAULT_TICK] = tick;
                    }
                }
            }
        }
        if (lane == 0) {
            hdr[H_NREC] = nrec;
            if (nrec > P.log_cap) hdr[H_FAULT] |= ZRK_FAULT_LOG_OVERFLOW;
            if (!(hdr[H_FAULT] & 63)) {
                hdr[H_TICK] = tick + 1;
                // the lists consumed during this tick become the next tick's output lists
                hdr[H_MBN0 + prv] = 0;
                hdr[H_DETN0 + prv] = 0;
                hdr[H_LN] = 0;
            }
        }
        if (++noise_w == P.noise_period) noise_w = 0;
        tm.sync();
    }

    if (two && lane == 0) {  // still short of the target: there will be another round
        const int t = hdr[H_TICK];
        if (t < target_tick && !(hdr[H_FAULT] & 63) && !(tick_limit >= 0 && t >= tick_limit)) {
            const int k = atomicAdd(pending, 1);
            if (work_out) work_out[k] = b;
        }
    }
    // ---- epilogue: shared memory -> HBM, one bulk copy
    stage_fence();
    tm.sync();
    stage_out(P.dyn + (size_t)b * o.dyn_bytes, sm, o.dyn_bytes, lane == 0);
}

#undef ZRK_X
#undef ZRK_ME_I
#undef ZRK_ME_D

// ------------------------------------------------------------------ the streak kernel (two-kernel stepping)
//
// Warp teams only.  One warp per scenario runs quiet ticks straight from the scenario's blob in global
// memory: a streak reads each word of the state once when it begins (coalesced: the blob is a structure
// of arrays, one element per lane) and writes back what changed when it ends, so the kernel streams the
// batch through HBM once per launch and otherwise lives in registers.  Its code is a few KB — the reason
// for a kernel of its own: with the streak inlined in the tick kernel, the two code paths evict each
// other from the SM's instruction cache.  A scenario whose next tick is not a quiet one is flagged
// CF_GENERAL for zrk_tick_kernel; the host alternates the two kernels until no scenario is pending.
#ifndef ZRK_STREAK_WARPS
#define ZRK_STREAK_WARPS 4
#endif
#ifndef ZRK_STREAK_MINB
#define ZRK_STREAK_MINB 4 /* resident CTAs per SM the register allocation is bounded for: 128 registers (the lean loop spills at 96: 36.3 vs 35.7 ms per c5 step) */
#endif
#define ZRK_STREAK_SMEM (ZRK_STREAK_WARPS * (2 * 32 * 16 + 32 * 8 + ZRK_RREC_DOUBLES * 8))
// One scenario's turn in the streak kernel (warp w of the CTA, scratch: the CTA's shared memory).
template <int MPB>
__device__ __forceinline__ void streak_scenario(const Params &P, int b, int target_tick, int w, int lane, float4 *scratch) {
    const Layout &o = P.lo;
    unsigned char *sm = P.dyn + (size_t)b * o.dyn_bytes;
    const unsigned char *cs = P.cst + (size_t)b * o.cst_bytes;
    Hot<32> H;
    H.sm = sm;
    int *const hdr = H.hdr();
    const int tick = hdr[H_TICK], flag = *H.c_flag();
    const int *chdr = (const int *)(cs + o.c_hdr);
    const int lim = chdr[C_TICK_LIMIT];
    if (tick >= target_tick || (hdr[H_FAULT] & 63) || (lim >= 0 && tick >= lim) || (flag & CF_GENERAL)) return;
    const int prv0 = (tick & 1) ^ 1;
    const bool fit = P.cert && chdr[C_CCP_ID] != ZRK_NONE && !(flag & (CF_WINDOWS | CF_TRACKS)) && *H.c_removed() == 0 &&
                     hdr[H_DETN0 + prv0] == 0 && hdr[H_MBN0 + prv0] == 0 && hdr[H_CCP_INIT] && !(P.log_types & ZRK_LOG_BULK);
    int done = 0;
    if (fit) {
        const int noise_w = P.noise ? tick % P.noise_period : 0;
        done = quiet_streak<32, MPB>(P, sm, cs, lane, b, target_tick - tick, noise_w, scratch + (2 * w) * 32, scratch + (2 * w + 1) * 32,
                                     (double *)(scratch + 2 * ZRK_STREAK_WARPS * 32) + 32 * w,
                                     (double *)(scratch + 2 * ZRK_STREAK_WARPS * 32) + 32 * ZRK_STREAK_WARPS + ZRK_RREC_DOUBLES * w);
    }
    if (tick + done < target_tick && lane == 0) {
        // the tick kernel's turn.  A streak that ended before its first tick doubles the number of ticks the
        // scenario then stays there (bits CF_HOLD_SHIFT.. of the flag word; up to 64), one that got going resets it.
        int f = *H.c_flag();
        int hold = (f >> CF_HOLD_SHIFT) & 127;
        hold = fit && done == 0 ? (hold ? (hold * 2 > 64 ? 64 : hold * 2) : 1) : (done >= 4 ? 0 : hold);
        *H.c_flag() = (f & ((1 << CF_HOLD_SHIFT) - 1)) | (hold << CF_HOLD_SHIFT) | CF_GENERAL;
    }
}
// The kernel is persistent: the grid is sized for the SMs, and every warp draws scenarios from a queue (*queue: a
// counter the host has zeroed) until it is empty — streaks differ in length by two orders of magnitude, and a warp that
// is done early would otherwise leave its slot idle until the slowest warp of its CTA has finished.
// The scenarios: [b_first, b_end), or — work given — the entries [0, b_end - b_first) of that list (work_n, may be null:
// where the device keeps the length of the list, when the host only knows an upper bound of it).
template <int MPB>
__global__ void __launch_bounds__(32 * ZRK_STREAK_WARPS, ZRK_STREAK_MINB) zrk_streak_kernel(const Params P, int b_first, int b_end, int target_tick,
                                                                                           const int *work, const int *work_n, int *queue) {
    ZRK_SMEM_DECL;  // ZRK_STREAK_SMEM bytes: per warp, two rows of fp32 images for a window rebuild inside a streak
    float4 *const scratch = (float4 *)zrk_smem;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n = b_end - b_first;
    if (work && work_n && *work_n < n) n = *work_n;
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(queue, 1);
        i = __shfl_sync(ZRK_FULL, i, 0);
        if (i >= n) break;
        streak_scenario<MPB>(P, work ? work[i] : b_first + i, target_tick, w, lane, scratch);
        __syncwarp();  // (the warp's scratch is the next scenario's)
    }
}

}  // namespace zrk
#endif
