// zrk_tick.cuh — the batched tick kernel (sm_100a).
//
// One TEAM of threads per scenario, one lane per AirEnv slot (AirEnv.__objects position,
// AirEnv.py:24).  A launch runs n_ticks passes of the loop body of Manager.run_simulation
// (Manager.py:116-140) with the scenario's mutable state resident in shared memory and the
// slot trajectories in registers; the state is read from / written back to HBM once per launch.
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
#define ZRK_NAMED_BARRIER(id, n) zrk_emu::named_barrier((id), (n))
#define ZRK_SMEM_DECL unsigned char *zrk_smem = zrk_emu::dyn_smem()
#define ZRK_LDCG(p) (*(p))
#define ZRK_NOINLINE __attribute__((noinline))
#else
#include <cuda_runtime.h>
#define ZRK_NAMED_BARRIER(id, n) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory")
#define ZRK_SMEM_DECL extern __shared__ __align__(16) unsigned char zrk_smem[]
#define ZRK_LDCG(p) __ldcg(p)
#define ZRK_NOINLINE __noinline__
#endif

#include <math.h>

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
    double r = fmod(a, m);
    if (r != 0.0) {
        if (r < 0.0) r += m;
    } else {
        r = 0.0;
    }
    return r;
}
__device__ __forceinline__ double to_seconds(long long ms) { return (double)ms / 1000.0; }

#define ZRK_RAD2DEG 57.29577951308232 /* 180.0 / pi, as np.degrees multiplies */

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
template <bool FAST>
__device__ __forceinline__ bool radar_sees(double dx, double dy, double dz, const double *par, double az_lo,
                                           double az_hi, double el_lo, double el_hi) {
    if (FAST) {
        const float ANG_MARGIN = 4e-3f;
        double d2 = dot3(dx, dy, dz, dx, dy, dz);
        if (d2 > par[9]) return false;  // certainly farther than max_distance
        bool range_sure = d2 <= par[8];
        float fx = (float)dx, fy = (float)dy, fz = (float)dz;
        float az = atan2f(fy, fx) * 57.29577951f;
        if (az < 0.f) az += 360.f;
        float rho = sqrtf(fx * fx + fy * fy);
        float el = atan2f(fz, rho) * 57.29577951f;  // == asin(dz / distance), well conditioned
        if (el < 0.f) el += 180.f;
        float a0 = (float)az_lo, a1 = (float)az_hi, e0 = (float)el_lo, e1 = (float)el_hi;
        bool wrap_far = az > ANG_MARGIN && az < 360.f - ANG_MARGIN && el > ANG_MARGIN && el < 180.f - ANG_MARGIN;
        bool in_sure = az > a0 + ANG_MARGIN && az < a1 - ANG_MARGIN && el > e0 + ANG_MARGIN && el < e1 - ANG_MARGIN;
        bool out_sure = az < a0 - ANG_MARGIN || az > a1 + ANG_MARGIN || el < e0 - ANG_MARGIN || el > e1 + ANG_MARGIN;
        // thresholds themselves must be representable well enough in fp32: |x| <= 1e4 degrees
        bool thr_ok = fabsf(a0) < 1e4f && fabsf(a1) < 1e4f && fabsf(e0) < 1e4f && fabsf(e1) < 1e4f;
        if (wrap_far && thr_ok) {
            if (out_sure) return false;
            if (in_sure && range_sure) return true;
        }
    }
    return radar_sees_exact(dx, dy, dz, par[2], az_lo, az_hi, el_lo, el_hi);
}

// ------------------------------------------------------------------ team primitives

template <int TEAM, int TPC>
struct Team {
    int lane;      // 0..TEAM-1 within the scenario's team
    int tic;       // team index inside the CTA
    int *wscr;     // >= 34 ints of shared scratch

    __device__ __forceinline__ void sync() const {
        if (TEAM == 32) {
            __syncwarp();
        } else if (TPC == 1) {
            __syncthreads();
        } else {
            ZRK_NAMED_BARRIER(tic + 1, TEAM);
        }
    }
    // exclusive prefix sum in lane order; total returned to every lane
    __device__ __forceinline__ int scan(int v, int &total) const {
        int l32 = lane & 31;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int n = __shfl_up_sync(ZRK_FULL, inc, d);
            if (l32 >= d) inc += n;
        }
        if (TEAM == 32) {
            total = __shfl_sync(ZRK_FULL, inc, 31);
            return inc - v;
        }
        int w = lane >> 5;
        if (l32 == 31) wscr[w] = inc;
        sync();
        int base = 0, tot = 0;
        for (int i = 0; i < TEAM / 32; i++) {
            int x = wscr[i];
            if (i < w) base += x;
            tot += x;
        }
        sync();
        total = tot;
        return base + inc - v;
    }
    __device__ __forceinline__ int min(int v) const {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            int n = __shfl_xor_sync(ZRK_FULL, v, d);
            v = n < v ? n : v;
        }
        if (TEAM == 32) return v;
        int w = lane >> 5;
        if ((lane & 31) == 0) wscr[w] = v;
        sync();
        int m = ZRK_IMAX;
        for (int i = 0; i < TEAM / 32; i++) {
            int x = wscr[i];
            m = x < m ? x : m;
        }
        sync();
        return m;
    }
    __device__ __forceinline__ bool any(bool p) const { return min(p ? 0 : 1) == 0; }
};

// ------------------------------------------------------------------ record log

__device__ __forceinline__ void put_rec(const Params &P, zrk_msg *base, int idx, int tick, int type, int flag,
                                        int sender, int receiver, int a, int b, double x, double y, double z) {
    if (idx < P.log_cap) {
        zrk_msg *r = base + idx;
        r->tick = tick;
        r->type = (unsigned char)type;
        r->relevance = (unsigned char)(type == ZRK_MSG_MISSILE_COUNT_REQUEST ? 3 : 1);
        r->flag = (unsigned char)flag;
        r->pad = 0;
        r->sender = sender;
        r->receiver = receiver;
        r->a = a;
        r->b = b;
        r->xyz[0] = x;
        r->xyz[1] = y;
        r->xyz[2] = z;
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

// ------------------------------------------------------------------ CCP.link_object

// Shared-memory views of one scenario.
struct Scen {
    int *hdr;
    unsigned *counts;
    double *pos;  // [6][NSP]: pos xyz, prev xyz
    double *speed;
    int *sid, *sflags, *srf;
    int *te_key, *te_ref, *te_upd, *te_follow, *te_mid;
    int *me_key, *me_ref, *me_upd, *me_tid;
    double *rad_cur;
    int *st_n, *st, *cap, *launched;
    int *m_status, *m_target, *m_slot, *m_birth;
    double *m_period;
    int *det_mid, *det_tid, *det_flag, *mb_type, *mb_l, *mb_a, *ln_m;
    const int *c_hdr, *c_rad_id, *c_l_id, *c_m_id, *c_m_launcher, *c_cl_id, *c_cl_launcher;
    const double *c_rad_pos, *c_rad_par, *c_l_pos, *c_m_v0, *c_m_v0sq, *c_m_radius, *c_cl_pos;
    int *x_claim, *x_detlist, *x_wscr, *x_need, *x_ok;
};

__device__ __forceinline__ void bind(Scen &s, unsigned char *sm, const Layout &o) {
    unsigned char *c = sm + o.dyn_bytes;
    s.hdr = (int *)(sm + o.hdr);
    s.counts = (unsigned *)(sm + o.counts);
    s.pos = (double *)(sm + o.pos);
    s.speed = (double *)(sm + o.speed);
    s.sid = (int *)(sm + o.sid);
    s.sflags = (int *)(sm + o.sflags);
    s.srf = (int *)(sm + o.srf);
    s.te_key = (int *)(sm + o.te_key);
    s.te_ref = (int *)(sm + o.te_ref);
    s.te_upd = (int *)(sm + o.te_upd);
    s.te_follow = (int *)(sm + o.te_follow);
    s.te_mid = (int *)(sm + o.te_mid);
    s.me_key = (int *)(sm + o.me_key);
    s.me_ref = (int *)(sm + o.me_ref);
    s.me_upd = (int *)(sm + o.me_upd);
    s.me_tid = (int *)(sm + o.me_tid);
    s.rad_cur = (double *)(sm + o.rad_cur);
    s.st_n = (int *)(sm + o.st_n);
    s.st = (int *)(sm + o.st);
    s.cap = (int *)(sm + o.cap);
    s.launched = (int *)(sm + o.launched);
    s.m_status = (int *)(sm + o.m_status);
    s.m_target = (int *)(sm + o.m_target);
    s.m_slot = (int *)(sm + o.m_slot);
    s.m_birth = (int *)(sm + o.m_birth);
    s.m_period = (double *)(sm + o.m_period);
    s.det_mid = (int *)(sm + o.det_mid);
    s.det_tid = (int *)(sm + o.det_tid);
    s.det_flag = (int *)(sm + o.det_flag);
    s.mb_type = (int *)(sm + o.mb_type);
    s.mb_l = (int *)(sm + o.mb_l);
    s.mb_a = (int *)(sm + o.mb_a);
    s.ln_m = (int *)(sm + o.ln_m);
    s.c_hdr = (const int *)(c + o.c_hdr);
    s.c_rad_id = (const int *)(c + o.c_rad_id);
    s.c_rad_pos = (const double *)(c + o.c_rad_pos);
    s.c_rad_par = (const double *)(c + o.c_rad_par);
    s.c_l_id = (const int *)(c + o.c_l_id);
    s.c_l_pos = (const double *)(c + o.c_l_pos);
    s.c_m_id = (const int *)(c + o.c_m_id);
    s.c_m_launcher = (const int *)(c + o.c_m_launcher);
    s.c_m_v0 = (const double *)(c + o.c_m_v0);
    s.c_m_v0sq = (const double *)(c + o.c_m_v0sq);
    s.c_m_radius = (const double *)(c + o.c_m_radius);
    s.c_cl_id = (const int *)(c + o.c_cl_id);
    s.c_cl_launcher = (const int *)(c + o.c_cl_launcher);
    s.c_cl_pos = (const double *)(c + o.c_cl_pos);
    s.x_claim = (int *)(sm + o.x_claim);
    s.x_detlist = (int *)(sm + o.x_detlist);
    s.x_wscr = (int *)(sm + o.x_wscr);
    s.x_need = (int *)(sm + o.x_need);
    s.x_ok = (int *)(sm + o.x_ok);
}

// calc_range (CCP.py:176-188) for one (detection, entry) pair, exact.
__device__ __forceinline__ void calc_range(double speed, double cur_time, int upd, const Params &P, double &lo,
                                           double &hi) {
    double d_t = cur_time - to_seconds((long long)upd * P.dt_ms);
    lo = speed * (d_t - P.range_step);
    hi = speed * (d_t + P.range_step);
    lo = lo > 0 ? lo : 0;  // max(0, x)
    hi = hi > 0 ? hi : 0;
}

// CombatControlPoint.link_object (CCP.py:171-219): exact scan, one detection per lane.
// cls: 0 NEW_TARGET, 1 OLD_TARGET (cand indexes te_*), 2 OLD_ROCKET (cand indexes me_*).
__device__ ZRK_NOINLINE void link_exact(const Scen &s, const Params &P, int NSP, int tick, double cur_time,
                                        double px, double py, double pz, double speed, int &cls, int &cand) {
    double curr = INFINITY;
    cls = 0;
    cand = -1;
    int n_te = s.hdr[H_NTE], n_me = s.hdr[H_NME];
    for (int e = 0; e < n_te; e++) {
        int upd = s.te_upd[e];
        if (upd == tick) continue;
        int ref = s.te_ref[e];
        double ex = s.pos[3 * NSP + ref] - px, ey = s.pos[4 * NSP + ref] - py, ez = s.pos[5 * NSP + ref] - pz;
        double diff = sqrt(dot3(ex, ey, ez, ex, ey, ez));
        double lo, hi;
        calc_range(speed, cur_time, upd, P, lo, hi);
        if (diff < curr && lo <= diff && diff <= hi) {
            curr = diff;
            cls = 1;
            cand = e;
        }
    }
    for (int e = 0; e < n_me; e++) {
        int upd = s.me_upd[e];
        if (upd == tick) continue;
        int ref = s.me_ref[e];
        int o = (s.sflags[ref] & F_HASPREV) ? 3 : 0;  // prev_pos, or pos when prev_pos is None
        double ex = s.pos[(o + 0) * NSP + ref] - px, ey = s.pos[(o + 1) * NSP + ref] - py,
               ez = s.pos[(o + 2) * NSP + ref] - pz;
        double diff = sqrt(dot3(ex, ey, ez, ex, ey, ez));
        double lo, hi;
        calc_range(speed, cur_time, upd, P, lo, hi);
        if (diff < curr && lo <= diff && diff <= hi) {
            curr = diff;
            cls = 2;
            cand = e;
        }
    }
}

// The same scan in the squared-distance domain.  sqrt is monotone, so "diff < curr" and the
// range test can be decided from d2 whenever d2 is not within a few ulps of best_d2 / lo^2 /
// hi^2; any such borderline (or NaN) sets `unsure` and the caller redoes this detection with
// link_exact.  Range bounds are cached per distinct upd tick.
__device__ __forceinline__ void link_fast(const Scen &s, const Params &P, int NSP, int tick, double cur_time,
                                          double px, double py, double pz, double speed, int &cls, int &cand,
                                          bool &unsure) {
    const double EPS = 4e-15;
    double best = INFINITY, best_lo = INFINITY;  // best_lo = best * (1 - EPS)
    cls = 0;
    cand = -1;
    unsure = false;
    int c_upd = -0x40000000;
    double lo2a = 0, lo2b = 0, hi2a = 0, hi2b = 0;  // lo^2*(1-EPS), lo^2*(1+EPS), hi^2*(1-EPS), hi^2*(1+EPS)
    int n_te = s.hdr[H_NTE], n_me = s.hdr[H_NME];
    for (int pass = 0; pass < 2; pass++) {
        int n = pass ? n_me : n_te;
        const int *a_upd = pass ? s.me_upd : s.te_upd;
        const int *a_ref = pass ? s.me_ref : s.te_ref;
        for (int e = 0; e < n; e++) {
            int upd = a_upd[e];
            if (upd == tick) continue;
            int ref = a_ref[e];
            int o = (!pass || (s.sflags[ref] & F_HASPREV)) ? 3 : 0;
            double ex = s.pos[(o + 0) * NSP + ref] - px, ey = s.pos[(o + 1) * NSP + ref] - py,
                   ez = s.pos[(o + 2) * NSP + ref] - pz;
            double d2 = dot3(ex, ey, ez, ex, ey, ez);
            if (upd != c_upd) {
                double lo, hi;
                calc_range(speed, cur_time, upd, P, lo, hi);
                double l2 = lo * lo, h2 = hi * hi;
                lo2a = l2 * (1.0 - EPS);
                lo2b = l2 * (1.0 + EPS);
                hi2a = h2 * (1.0 - EPS);
                hi2b = h2 * (1.0 + EPS);
                c_upd = upd;
            }
            if (d2 > hi2b || d2 < lo2a) continue;     // certainly outside [lo, hi]
            if (d2 > best) continue;                  // sqrt(d2) >= curr
            if (d2 == best) continue;                 // equal diff is not strictly smaller
            bool in_sure = d2 >= lo2b && d2 <= hi2a;
            bool less_sure = d2 < best_lo;
            if (in_sure && less_sure) {
                best = d2;
                best_lo = d2 * (1.0 - EPS);
                cls = pass + 1;
                cand = e;
            } else {
                unsure = true;  // includes NaN: every comparison above was false
            }
        }
    }
}

// ------------------------------------------------------------------ the kernel

template <int TEAM, int TPC, bool FAST>
__global__ void __launch_bounds__(TEAM *TPC) zrk_tick_kernel(const Params P, int n_ticks) {
    ZRK_SMEM_DECL;
    const Layout &o = P.lo;
    const int NSP = o.NSP, MP = o.MP, Q = o.Q;
    Team<TEAM, TPC> tm;
    tm.tic = threadIdx.x / TEAM;
    tm.lane = threadIdx.x % TEAM;
    const int lane = tm.lane;
    const int b = blockIdx.x * TPC + tm.tic;
    if (b >= o.B) return;  // whole team leaves together
    unsigned char *sm = zrk_smem + (size_t)tm.tic * o.team_smem_bytes;
    Scen s;
    bind(s, sm, o);
    tm.wscr = s.x_wscr;

    // ---- prologue: HBM -> shared memory / registers
    {
        const uint4 *gd = (const uint4 *)(P.dyn + (size_t)b * o.dyn_bytes);
        uint4 *sd = (uint4 *)sm;
        for (int i = lane; i < o.dyn_bytes / 16; i += TEAM) sd[i] = gd[i];
        const uint4 *gc = (const uint4 *)(P.cst + (size_t)b * o.cst_bytes);
        uint4 *sc = (uint4 *)(sm + o.dyn_bytes);
        for (int i = lane; i < o.cst_bytes / 16; i += TEAM) sc[i] = gc[i];
    }
    double *gtraj = P.traj + (size_t)b * 8 * NSP;  // [8][NSP]: p0x p0y p0z t0 vx vy vz speed
    double p0x = 0, p0y = 0, p0z = 0, t0 = 0, vx = 0, vy = 0, vz = 0;
    if (lane < o.NS) {
        p0x = gtraj[0 * NSP + lane];
        p0y = gtraj[1 * NSP + lane];
        p0z = gtraj[2 * NSP + lane];
        t0 = gtraj[3 * NSP + lane];
        vx = gtraj[4 * NSP + lane];
        vy = gtraj[5 * NSP + lane];
        vz = gtraj[6 * NSP + lane];
    }
    tm.sync();

    const int nR = s.c_hdr[C_NR], nL = s.c_hdr[C_NL], nC = s.c_hdr[C_NC], T = o.T;
    const int ccp_id = s.c_hdr[C_CCP_ID], airenv_id = s.c_hdr[C_AIRENV_ID];
    const int tick_limit = s.c_hdr[C_TICK_LIMIT];
    const bool has_ccp = ccp_id != ZRK_NONE;
    zrk_msg *rec = P.rec + (size_t)b * P.log_cap;
    const double *noise = P.noise ? P.noise + (size_t)b * P.noise_stride : nullptr;
    (void)airenv_id;

    for (int it = 0; it < n_ticks; it++) {
        const int tick = s.hdr[H_TICK];
        if (s.hdr[H_FAULT] & 63) break;             // the reference run aborted here
        if (tick_limit >= 0 && tick >= tick_limit) break;
        const int par = tick & 1, prv = par ^ 1;
        const double now_s = to_seconds((long long)tick * P.dt_ms);
        const int n_slots = s.hdr[H_NSLOTS];
        const int det_prev = s.hdr[H_DETN0 + prv];
        tm.sync();  // everyone has read the header before lane 0 resets the per-tick lists
        if (lane == 0) {
            s.hdr[H_MBN0 + par] = 0;
            s.hdr[H_DETN0 + par] = 0;
            s.hdr[H_LN] = 0;
        }

        // ============ AirEnv.step (AirEnv.py:26-53)
        int fl = 0, my_id = 0;
        double px = 0, py = 0, pz = 0;
        if (lane < n_slots) {
            fl = s.sflags[lane];
            my_id = s.sid[lane];
            if (fl & F_ALIVE) {  // :33-40 removal by id
                for (int i = 0; i < det_prev; i++) {
                    int tid = s.det_tid[prv * MP + i];
                    if (my_id == s.det_mid[prv * MP + i] || (tid != ZRK_NONE && my_id == tid)) fl &= ~F_ALIVE;
                }
            }
            if (fl & F_PENDING) {  // :42-43 NEW_MISSILE of the previous tick
                int m = (fl >> F_STATIC_SHIFT) - T;
                if (s.m_birth[m] == tick) {
                    fl = (fl & ~F_PENDING) | F_ALIVE;
                    p0x = ZRK_LDCG(&gtraj[0 * NSP + lane]);
                    p0y = ZRK_LDCG(&gtraj[1 * NSP + lane]);
                    p0z = ZRK_LDCG(&gtraj[2 * NSP + lane]);
                    t0 = ZRK_LDCG(&gtraj[3 * NSP + lane]);
                    vx = ZRK_LDCG(&gtraj[4 * NSP + lane]);
                    vy = ZRK_LDCG(&gtraj[5 * NSP + lane]);
                    vz = ZRK_LDCG(&gtraj[6 * NSP + lane]);
                }
            }
            if (fl & F_ALIVE) {  // AirObject.step (AirObject.py:39-42)
                if (t0 != now_s) {
                    s.pos[3 * NSP + lane] = s.pos[0 * NSP + lane];
                    s.pos[4 * NSP + lane] = s.pos[1 * NSP + lane];
                    s.pos[5 * NSP + lane] = s.pos[2 * NSP + lane];
                    fl |= F_HASPREV;
                } else {
                    fl &= ~F_HASPREV;
                }
                double dtm = now_s - t0;
                px = p0x + vx * dtm;
                py = p0y + vy * dtm;
                pz = p0z + vz * dtm;
                s.pos[0 * NSP + lane] = px;
                s.pos[1 * NSP + lane] = py;
                s.pos[2 * NSP + lane] = pz;
            }
            s.sflags[lane] = fl;
        }
        const bool alive = (fl & F_ALIVE) != 0;
        tm.sync();

        // Missile.step, status 'active' (Missile.py:162-193), in slot order
        int nrec_a;  // records written so far this tick (uniform across the team)
        {
            int det = 0, det_tid = ZRK_NONE;
            const bool msl = alive && (fl & F_MISSILE);
            if (msl) {
                int m = (fl >> F_STATIC_SHIFT) - T;
                int tg = s.m_target[m];
                double ex = s.pos[0 * NSP + tg] - px, ey = s.pos[1 * NSP + tg] - py, ez = s.pos[2 * NSP + tg] - pz;
                double dist = sqrt(dot3(ex, ey, ez, ex, ey, ez));
                if (dist <= s.c_m_radius[m]) {
                    det = 1;
                    det_tid = s.sid[tg];
                } else {
                    double per = s.m_period[m] - P.dt_s;
                    s.m_period[m] = per;
                    if (per <= 0) det = 2;
                }
                if (det) s.m_status[m] = ST_DETONATED;
            }
            const bool lp = logged(P, ZRK_MSG_MISSILE_POS), ld = logged(P, ZRK_MSG_MISSILE_DETONATE);
            const int nrec = (msl && lp ? 1 : 0) + (det && ld ? 1 : 0);
            int ptot, n_msl;
            const int pex = tm.scan(nrec | ((det ? 1 : 0) << 16), ptot);
            (void)tm.scan(msl ? 1 : 0, n_msl);
            const int rbase = s.hdr[H_NREC];
            int ri = rbase + (pex & 0xffff);
            const int di = pex >> 16;
            if (msl && lp) put_rec(P, rec, ri++, tick, ZRK_MSG_MISSILE_POS, 0, my_id, ZRK_NONE, my_id, ZRK_NONE, 0, 0, 0);
            if (det) {
                if (ld)
                    put_rec(P, rec, ri, tick, ZRK_MSG_MISSILE_DETONATE, det == 2, my_id, ZRK_NONE, my_id, det_tid, 0, 0, 0);
                s.det_mid[par * MP + di] = my_id;
                s.det_tid[par * MP + di] = det_tid;
                s.det_flag[par * MP + di] = det == 2;
            }
            nrec_a = rbase + (ptot & 0xffff);
            if (lane == 0) {
                s.hdr[H_DETN0 + par] = ptot >> 16;
                s.counts[ZRK_MSG_MISSILE_POS] += (unsigned)n_msl;
                s.counts[ZRK_MSG_MISSILE_DETONATE] += (unsigned)(ptot >> 16);
                s.counts[ZRK_MSG_ACTIVE_OBJECTS] += 1;
            }
        }
        unsigned *mrow = nullptr;
        if (P.masks) {
            mrow = P.masks + ((size_t)(tick % P.mask_ticks) * o.B + b) * (size_t)((1 + o.RP) * P.mask_words);
            unsigned bal = __ballot_sync(ZRK_FULL, alive);
            if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[lane >> 5] = bal;
        }

        // ============ SectorRadar.step for every radar (Radar.py:144-205)
        int rf = -1;
        for (int r = 0; r < nR; r++) {
            const double *par_r = s.c_rad_par + RAD_PAR_STRIDE * r;
            bool vis = false;
            if (alive) {
                double caz = s.rad_cur[2 * r], cel = s.rad_cur[2 * r + 1];
                double dx = px - s.c_rad_pos[3 * r], dy = py - s.c_rad_pos[3 * r + 1], dz = pz - s.c_rad_pos[3 * r + 2];
                vis = radar_sees<FAST>(dx, dy, dz, par_r, caz, caz + par_r[3], cel, cel + par_r[4]);
                if (vis) {  // smooth_objects (Radar.py:138-142)
                    if (noise) {
                        const double *nz = noise + (((size_t)(tick % P.noise_period) * P.noise_radars + r) * P.noise_objects +
                                                    (fl >> F_STATIC_SHIFT)) * 3;
                        px += nz[0];
                        py += nz[1];
                        pz += nz[2];
                    }
                    if (rf < 0) rf = r;
                }
            }
            if (mrow) {
                unsigned bal = __ballot_sync(ZRK_FULL, vis);
                if ((lane & 31) == 0 && (lane >> 5) < P.mask_words) mrow[(1 + r) * P.mask_words + (lane >> 5)] = bal;
            }
            // :193-203 one DESTROYED_MISSILE per MISSILE_DETONATE of the previous tick
            if (det_prev && logged(P, ZRK_MSG_DESTROYED_MISSILE) && lane < det_prev)
                put_rec(P, rec, nrec_a + r * det_prev + lane, tick, ZRK_MSG_DESTROYED_MISSILE,
                        s.det_flag[prv * MP + lane], s.c_rad_id[r], ZRK_CCP_ID, s.det_mid[prv * MP + lane], ZRK_NONE,
                        0, 0, 0);
        }
        if (lane < NSP) s.srf[lane] = rf;
        if (alive) {
            s.pos[0 * NSP + lane] = px;
            s.pos[1 * NSP + lane] = py;
            s.pos[2 * NSP + lane] = pz;
        }
        tm.sync();

        // ============ lane 0: radar bookkeeping, MissileLauncher.step, CCP bookkeeping
        if (lane == 0) {
            s.hdr[H_NREC] = nrec_a;
            if (nR) {
                if (logged(P, ZRK_MSG_DESTROYED_MISSILE)) s.hdr[H_NREC] += nR * det_prev;
                s.counts[ZRK_MSG_ALL_OBJECTS] += nR;
                s.counts[ZRK_MSG_FOUND_OBJECTS] += nR;
                s.counts[ZRK_MSG_DESTROYED_MISSILE] += nR * det_prev;
            }
            for (int r = 0; r < nR; r++) {  // move_to_next_sector_circular (Radar.py:96-117)
                const double *q = s.c_rad_par + RAD_PAR_STRIDE * r;
                double caz = s.rad_cur[2 * r], cel = s.rad_cur[2 * r + 1];
                int mode = (int)q[7];
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
                s.rad_cur[2 * r] = caz;
                s.rad_cur[2 * r + 1] = cel;
            }
            // MissileLauncher.step (MissileLauncher.py:82-138)
            const int n_prev = s.hdr[H_MBN0 + prv];
            if (n_prev) {
                int nrec = s.hdr[H_NREC];
                for (int l = 0; l < nL; l++) {
                    const int l_id = s.c_l_id[l];
                    for (int i = 0; i < n_prev; i++) {
                        if (s.mb_l[prv * Q + i] != l) continue;
                        const int ty = s.mb_type[prv * Q + i], arg = s.mb_a[prv * Q + i];
                        if (ty == MB_COMMAND) {  // launch_missile (:58-80)
                            if (s.st_n[l] == 0) continue;
                            const int m = s.st[l * MP + (--s.st_n[l])];
                            const int m_id = s.c_m_id[m], tg = arg;
                            s.counts[ZRK_MSG_LAUNCH_MISSILE]++;
                            if (logged(P, ZRK_MSG_LAUNCH_MISSILE))
                                put_rec(P, rec, nrec++, tick, ZRK_MSG_LAUNCH_MISSILE, 0, l_id, m_id, s.sid[tg], ZRK_NONE, 0, 0, 0);
                            // Missile._launch (Missile.py:104-133)
                            double tp[3] = {s.pos[0 * NSP + tg], s.pos[1 * NSP + tg], s.pos[2 * NSP + tg]};
                            double vt[3];
                            if (s.sflags[tg] & F_MISSILE) {
                                vt[0] = vt[1] = vt[2] = NAN;  // Missile.velocity is 0/0 (AirObject.py:35)
                            } else {
                                double wx = ZRK_LDCG(&gtraj[4 * NSP + tg]), wy = ZRK_LDCG(&gtraj[5 * NSP + tg]),
                                       wz = ZRK_LDCG(&gtraj[6 * NSP + tg]);
                                double sp = s.speed[tg];
                                vt[0] = wx / sp * sp;  // target.velocity * target.speed_mod (Missile.py:59)
                                vt[1] = wy / sp * sp;
                                vt[2] = wz / sp * sp;
                            }
                            double V[3];
                            int why = missile_solve(&s.c_l_pos[3 * l], tp, vt, s.c_m_v0[m], s.c_m_v0sq[m], s.m_period[m], V);
                            const int q = s.hdr[H_MBN0 + par]++;
                            s.mb_l[par * Q + q] = l;
                            s.mb_a[par * Q + q] = m;
                            if (why < 0) {
                                const int k = s.hdr[H_NSLOTS]++;
                                gtraj[0 * NSP + k] = s.c_l_pos[3 * l];
                                gtraj[1 * NSP + k] = s.c_l_pos[3 * l + 1];
                                gtraj[2 * NSP + k] = s.c_l_pos[3 * l + 2];
                                gtraj[3 * NSP + k] = now_s;
                                gtraj[4 * NSP + k] = V[0];
                                gtraj[5 * NSP + k] = V[1];
                                gtraj[6 * NSP + k] = V[2];
                                gtraj[7 * NSP + k] = s.c_m_v0[m];
                                s.pos[0 * NSP + k] = s.c_l_pos[3 * l];
                                s.pos[1 * NSP + k] = s.c_l_pos[3 * l + 1];
                                s.pos[2 * NSP + k] = s.c_l_pos[3 * l + 2];
                                s.speed[k] = s.c_m_v0[m];
                                s.sid[k] = m_id;
                                s.sflags[k] = F_PENDING | F_MISSILE | ((T + m) << F_STATIC_SHIFT);
                                s.m_status[m] = ST_ACTIVE;
                                s.m_target[m] = tg;
                                s.m_slot[m] = k;
                                s.m_birth[m] = tick + 2;
                                s.mb_type[par * Q + q] = MB_SUCCESS;
                                s.counts[ZRK_MSG_LAUNCH_SUCCESSFUL]++;
                                if (logged(P, ZRK_MSG_LAUNCH_SUCCESSFUL))
                                    put_rec(P, rec, nrec++, tick, ZRK_MSG_LAUNCH_SUCCESSFUL, 0, m_id, l_id, m_id, s.sid[tg], V[0], V[1], V[2]);
                            } else {
                                s.mb_type[par * Q + q] = MB_CANCEL;
                                s.counts[ZRK_MSG_LAUNCH_CANCELLED]++;
                                if (logged(P, ZRK_MSG_LAUNCH_CANCELLED))
                                    put_rec(P, rec, nrec++, tick, ZRK_MSG_LAUNCH_CANCELLED, 0, m_id, l_id, m_id, why, 0, 0, 0);
                            }
                        } else if (ty == MB_SUCCESS) {  // :103-124
                            const int m = arg, m_id = s.c_m_id[m], tid = s.sid[s.m_target[m]];
                            s.counts[ZRK_MSG_LAUNCHED_MISSILE]++;
                            s.counts[ZRK_MSG_NEW_MISSILE]++;
                            if (logged(P, ZRK_MSG_LAUNCHED_MISSILE))
                                put_rec(P, rec, nrec++, tick, ZRK_MSG_LAUNCHED_MISSILE, 0, l_id, ZRK_CCP_ID, m_id, tid, 0, 0, 0);
                            if (logged(P, ZRK_MSG_NEW_MISSILE))
                                put_rec(P, rec, nrec++, tick, ZRK_MSG_NEW_MISSILE, 0, l_id, ZRK_NONE, m_id, ZRK_NONE, 0, 0, 0);
                            s.ln_m[s.hdr[H_LN]++] = m;
                        } else if (ty == MB_CANCEL) {  // :125-128
                            s.st[l * MP + (s.st_n[l]++)] = arg;
                        } else {  // MissileCountRequestMessage (:129-138)
                            s.counts[ZRK_MSG_MISSILE_COUNT_RESPONSE]++;
                            if (logged(P, ZRK_MSG_MISSILE_COUNT_RESPONSE))
                                put_rec(P, rec, nrec++, tick, ZRK_MSG_MISSILE_COUNT_RESPONSE, 0, l_id, ccp_id, ZRK_NONE, s.st_n[l], 0, 0, 0);
                            // CCP.get_current_missile_launcher_capacity (CCP.py:139-147), same tick
                            for (int c = 0; c < nC; c++)
                                if (s.c_cl_launcher[c] == l) s.cap[c] = s.st_n[l];
                        }
                    }
                }
                s.hdr[H_NREC] = nrec;
            }
            if (has_ccp) {
                int nrec = s.hdr[H_NREC];
                if (!s.hdr[H_CCP_INIT]) {  // send_request_msg_to_ml_capacity (CCP.py:122-137)
                    for (int c = 0; c < nC; c++) {
                        s.cap[c] = 0;
                        s.launched[c] = 0;
                        s.counts[ZRK_MSG_MISSILE_COUNT_REQUEST]++;
                        if (logged(P, ZRK_MSG_MISSILE_COUNT_REQUEST))
                            put_rec(P, rec, nrec++, tick, ZRK_MSG_MISSILE_COUNT_REQUEST, 0, ccp_id, s.c_cl_id[c], ZRK_NONE, ZRK_NONE, 0, 0, 0);
                        const int q = s.hdr[H_MBN0 + par]++;
                        s.mb_type[par * Q + q] = MB_COUNTREQ;
                        s.mb_l[par * Q + q] = s.c_cl_launcher[c];
                        s.mb_a[par * Q + q] = 0;
                    }
                    s.hdr[H_CCP_INIT] = 1;
                }
                s.hdr[H_NREC] = nrec;
                // check_if_missiles_launched + add_missile (CCP.py:161-169, :102-109)
                const int n_ln = s.hdr[H_LN];
                for (int i = 0; i < n_ln; i++) {
                    const int m = s.ln_m[i], m_id = s.c_m_id[m], tid = s.sid[s.m_target[m]];
                    int e, n_me = s.hdr[H_NME];
                    for (e = 0; e < n_me; e++)
                        if (s.me_key[e] == m_id) break;
                    if (e == n_me) s.hdr[H_NME] = n_me + 1;
                    s.me_key[e] = m_id;
                    s.me_ref[e] = s.m_slot[m];
                    s.me_upd[e] = tick;
                    s.me_tid[e] = tid;
                    int t, n_te = s.hdr[H_NTE];
                    for (t = 0; t < n_te; t++)
                        if (s.te_key[t] == tid) break;
                    if (t == n_te) {
                        s.hdr[H_FAULT] |= ZRK_FAULT_KEY_ERROR_CCP107;
                        s.hdr[H_FAULT_TICK] = tick;
                        break;
                    }
                    s.te_mid[t] = m_id;
                }
            }
        }
        tm.sync();

        // ============ CombatControlPoint.step, detections (CCP.py:406-429)
        if (has_ccp && !(s.hdr[H_FAULT] & 63)) {
            // processing order: FOUND_OBJECTS messages in radar order, objects in list order,
            // each id once -> (first radar that saw it, slot)
            int n_det = 0, myrank = -1;
            for (int r = 0; r < nR; r++) {
                int tot;
                int ex = tm.scan(rf == r ? 1 : 0, tot);
                if (rf == r) myrank = n_det + ex;
                n_det += tot;
            }
            if (myrank >= 0) s.x_detlist[myrank] = lane;
            tm.sync();
            const int i = lane;
            int slot = 0, did = 0, radar_id = 0;
            double dpx = 0, dpy = 0, dpz = 0, dspeed = 0;
            bool unresolved = i < n_det;
            if (unresolved) {
                slot = s.x_detlist[i];
                did = s.sid[slot];
                radar_id = s.c_rad_id[s.srf[slot]];
                dpx = s.pos[0 * NSP + slot];
                dpy = s.pos[1 * NSP + slot];
                dpz = s.pos[2 * NSP + slot];
                dspeed = s.speed[slot];
            }
            int lo_rank = 0;
            for (;;) {
                int cls = 0, cand = -1, claim = -1;
                if (unresolved) {
                    if (FAST) {
                        bool unsure;
                        link_fast(s, P, NSP, tick, now_s, dpx, dpy, dpz, dspeed, cls, cand, unsure);
                        if (unsure) link_exact(s, P, NSP, tick, now_s, dpx, dpy, dpz, dspeed, cls, cand);
                    } else {
                        link_exact(s, P, NSP, tick, now_s, dpx, dpy, dpz, dspeed, cls, cand);
                    }
                    if (cls == 1)
                        claim = cand;
                    else if (cls == 2)
                        claim = NSP + cand;
                    else {  // add_target overwrites an existing key in place (CCP.py:92)
                        const int n_te = s.hdr[H_NTE];
                        for (int e = 0; e < n_te; e++)
                            if (s.te_key[e] == did) {
                                claim = e;
                                break;
                            }
                    }
                }
                // a detection is final once no earlier unresolved detection touches the same entry
                for (int x = lane; x < NSP + MP; x += TEAM) s.x_claim[x] = ZRK_IMAX;
                tm.sync();
                if (unresolved && claim >= 0) atomicMin(&s.x_claim[claim], i);
                tm.sync();
                const bool conflict = unresolved && claim >= 0 && s.x_claim[claim] != i;
                const int istar = tm.min(conflict ? i : ZRK_IMAX);
                const bool fin = unresolved && i < istar;

                const int following = (fin && cls == 1) ? s.te_follow[cand] : 0;
                const int old_mid = (fin && cls == 1) ? s.te_mid[cand] : ZRK_NONE;
                const bool need = fin && (cls == 0 || (cls == 1 && !following));
                int ok = 0;
                bool anycap = false;
                for (int c = 0; c < nC; c++) anycap = anycap || (s.launched[c] < s.cap[c]);
                if (anycap && tm.any(need)) {  // try_to_launch_missile (CCP.py:287-320), in order
                    if (i < NSP) s.x_need[i] = need;
                    tm.sync();
                    if (lane == 0) {
                        const int hi_rank = istar < n_det ? istar : n_det;
                        for (int j = lo_rank; j < hi_rank; j++) {
                            if (!s.x_need[j]) continue;
                            const int sl = s.x_detlist[j];
                            double ox = s.pos[0 * NSP + sl], oy = s.pos[1 * NSP + sl], oz = s.pos[2 * NSP + sl];
                            double min_dist = INFINITY;
                            int best = -1;
                            for (int c = 0; c < nC; c++) {
                                if (s.launched[c] < s.cap[c]) {
                                    double q0 = s.c_cl_pos[3 * c] - ox, q1 = s.c_cl_pos[3 * c + 1] - oy, q2 = s.c_cl_pos[3 * c + 2] - oz;
                                    double dist = sqrt((q0 * q0 + q1 * q1) + q2 * q2);
                                    if (dist < min_dist) {
                                        best = c;
                                        min_dist = dist;
                                    }
                                }
                            }
                            s.x_ok[j] = best + 1;
                            if (best >= 0) {
                                s.launched[best]++;
                                const int q = s.hdr[H_MBN0 + par]++;
                                s.mb_type[par * Q + q] = MB_COMMAND;
                                s.mb_l[par * Q + q] = s.c_cl_launcher[best];
                                s.mb_a[par * Q + q] = sl;
                            }
                        }
                    }
                    tm.sync();
                    if (need) ok = s.x_ok[i];
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
                int ptot, ctot;
                const int pex = tm.scan((wr ? 1 : 0) | ((app ? 1 : 0) << 16), ptot);
                (void)tm.scan((mtype == ZRK_MSG_CCP_UPDATE_TARGET ? 1 : 0) | ((mtype == ZRK_MSG_LAUNCH_COMMAND ? 1 : 0) << 16), ctot);
                const int rbase = s.hdr[H_NREC], n_te0 = s.hdr[H_NTE];
                if (wr) {
                    if (mtype == ZRK_MSG_LAUNCH_COMMAND)
                        put_rec(P, rec, rbase + (pex & 0xffff), tick, mtype, 0, ccp_id, s.c_cl_id[ok - 1], did, radar_id, dpx, dpy, dpz);
                    else
                        put_rec(P, rec, rbase + (pex & 0xffff), tick, mtype, 0, ccp_id, radar_id, did, old_mid, 0, 0, 0);
                }
                if (fin) {
                    if (cls == 1) {  // old_target (CCP.py:332-359)
                        s.te_ref[cand] = slot;
                        s.te_upd[cand] = tick;
                        if (!following) s.te_follow[cand] = ok != 0;
                    } else if (cls == 2) {  // old_rocket (CCP.py:361-366)
                        s.me_ref[cand] = slot;
                        s.me_upd[cand] = tick;
                    } else {  // new_target + add_target (CCP.py:322-330, :88-93)
                        const int e = claim >= 0 ? claim : n_te0 + (pex >> 16);
                        s.te_key[e] = did;
                        s.te_ref[e] = slot;
                        s.te_upd[e] = tick;
                        s.te_follow[e] = ok != 0;
                        s.te_mid[e] = ZRK_NONE;
                    }
                    unresolved = false;
                }
                tm.sync();
                if (lane == 0) {
                    s.hdr[H_NREC] = rbase + (ptot & 0xffff);
                    s.hdr[H_NTE] = n_te0 + (ptot >> 16);
                    s.counts[ZRK_MSG_LAUNCH_COMMAND] += (unsigned)(ctot >> 16);
                    s.counts[ZRK_MSG_CCP_UPDATE_TARGET] += (unsigned)(ctot & 0xffff);
                }
                tm.sync();
                if (istar == ZRK_IMAX) break;
                lo_rank = istar;
            }

            // ============ send_objects_to_GUI (CCP.py:237-284)
            {
                const int n_me = s.hdr[H_NME], n_te = s.hdr[H_NTE];
                const bool ldr = logged(P, ZRK_MSG_DRAW_OBJECTS);
                int mref = -1, tref = -1;
                bool dm = false, dt = false, bad = false;
                if (lane < n_me) {
                    mref = s.me_ref[lane];
                    dm = s.srf[mref] >= 0;
                }
                if (lane < n_te) {
                    tref = s.te_ref[lane];
                    dt = s.srf[tref] >= 0;
                    bad = dt && (s.sflags[tref] & F_MISSILE);  // .type on a Missile (CCP.py:259)
                }
                const int efault = tm.min(bad ? lane : ZRK_IMAX);
                dt = dt && lane < efault;
                const bool dother = efault == ZRK_IMAX && alive && rf < 0 && nR > 0;
                int tot_m, tot_t, tot_o;
                int ex_m = tm.scan(dm ? 1 : 0, tot_m);
                int ex_t = tm.scan(dt ? 1 : 0, tot_t);
                int ex_o = tm.scan(dother ? 1 : 0, tot_o);
                const int rbase = s.hdr[H_NREC];
                if (ldr) {
                    if (dm)
                        put_rec(P, rec, rbase + ex_m, tick, ZRK_MSG_DRAW_OBJECTS, 1, ccp_id, ZRK_MANAGER_ID, s.sid[mref], ZRK_ZUR,
                                s.pos[0 * NSP + mref], s.pos[1 * NSP + mref], s.pos[2 * NSP + mref]);
                    if (dt)
                        put_rec(P, rec, rbase + tot_m + ex_t, tick, ZRK_MSG_DRAW_OBJECTS, 1, ccp_id, ZRK_MANAGER_ID, s.sid[tref],
                                (s.sflags[tref] >> F_TTYPE_SHIFT) & 3, s.pos[0 * NSP + tref], s.pos[1 * NSP + tref],
                                s.pos[2 * NSP + tref]);
                    if (dother)
                        put_rec(P, rec, rbase + tot_m + tot_t + ex_o, tick, ZRK_MSG_DRAW_OBJECTS, 0, ccp_id, ZRK_MANAGER_ID, my_id,
                                (fl & F_MISSILE) ? ZRK_ZUR : ((fl >> F_TTYPE_SHIFT) & 3), px, py, pz);
                }
                tm.sync();
                if (lane == 0) {
                    if (ldr) s.hdr[H_NREC] = rbase + tot_m + tot_t + tot_o;
                    s.counts[ZRK_MSG_DRAW_OBJECTS] += (unsigned)(tot_m + tot_t + tot_o);
                    if (efault != ZRK_IMAX) {
                        s.hdr[H_FAULT] |= ZRK_FAULT_ATTRIBUTE_ERROR_CCP259;
                        s.hdr[H_FAULT_TICK] = tick;
                    }
                }
            }
        }
        tm.sync();
        if (lane == 0) {
            if (s.hdr[H_NREC] > P.log_cap) s.hdr[H_FAULT] |= ZRK_FAULT_LOG_OVERFLOW;
            if (!(s.hdr[H_FAULT] & 63)) s.hdr[H_TICK] = tick + 1;
        }
        tm.sync();
    }

    // ---- epilogue: shared memory -> HBM
    tm.sync();
    {
        uint4 *gd = (uint4 *)(P.dyn + (size_t)b * o.dyn_bytes);
        const uint4 *sd = (const uint4 *)sm;
        for (int i = lane; i < o.dyn_bytes / 16; i += TEAM) gd[i] = sd[i];
    }
}

}  // namespace zrk
#endif
