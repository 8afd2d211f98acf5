// zrk_layout.h — device data layout shared by the host library and the tick kernel.
//
// Per scenario the device keeps two things:
//   dyn    [B][dyn_bytes]        the scenario's state (Trajectory + speed_mod per AirEnv slot,
//                                 AirObject.py:15-36; positions; CCP tracks; ...) as one blob whose image in
//                                 shared memory is byte-identical (prologue/epilogue are plain
//                                 coalesced 16-byte copies).  It starts with the "hot" part,
//                                 whose offsets depend only on TEAM (compile-time constants in
//                                 the kernel), followed by the small launcher/missile/radar
//                                 arrays at run-time offsets.
//   cst    [B][cst_bytes]        read-only scenario parameters (radars, launchers, missiles, CCP).
#ifndef ZRK_LAYOUT_H
#define ZRK_LAYOUT_H

#include "../../include/zrk_b200.h"

namespace zrk {

// slot flag bits (dyn.sflags); static index in bits 8..
enum { F_ALIVE = 1, F_PENDING = 2, F_HASPREV = 4, F_MISSILE = 8, F_TTYPE_SHIFT = 4, F_STATIC_SHIFT = 8 };
// header words (ints) at dyn.hdr
enum {
    H_TICK = 0, H_NSLOTS, H_NTE, H_NME, H_CCP_INIT, H_FAULT, H_FAULT_TICK, H_NREC,
    H_DETN0, H_DETN1, H_MBN0, H_MBN1, H_LN, H_FREECAP, H_ALIVE /* alive slot-ticks so far (mod 2^32) */, H_FAST /* ticks whose CCP step took the certified path */, H_WORDS = 16
};
enum { MB_COMMAND = 0, MB_SUCCESS = 1, MB_CANCEL = 2, MB_COUNTREQ = 3 };
enum { ST_READY = 0, ST_ACTIVE = 1, ST_DETONATED = 2 };
enum { RAD_PAR_STRIDE = 12 };  // radar_par[8] + max_distance^2 bracket (lo, hi) + [10] 1.0 if the sector never moves
// cst header words
enum { C_NT = 0, C_NR, C_NL, C_NM, C_NC, C_AIRENV_ID, C_CCP_ID, C_TICK_LIMIT, C_WORDS = 8 };

#if defined(__CUDACC__)
#define ZRK_HD __host__ __device__
#else
#define ZRK_HD
#endif

// Hot part of the dyn blob: byte offsets as a function of TEAM only.
struct HotOff {
    int hdr, counts;
    int pxy, pzs;  // double2[TEAM]: (pos.x, pos.y), (pos.z, speed_mod)
    int qxy, qzw;  // double2[TEAM]: (prev.x, prev.y), (prev.z, trajectory velocity.z)
    int txy, tzt;  // double2[TEAM]: trajectory (start_pos.x, start_pos.y), (start_pos.z, start_time)
    int vxy;       // double2[TEAM]: trajectory velocity (x, y)
    int sid, sflags, srf, own;                        // int[TEAM]
    int te_key, te_ref, te_upd, te_follow, te_mid;    // int[TEAM]  CCP._target_dict
    int te_upds;                                      // double[TEAM] TargetCCP.upd_time (seconds)
    int c_in, c_out, c_refc;                          // int[TEAM] (teams up to 256 slots) association certificate, zrk_tick.cuh:
                                                      // per slot the window of ticks it is not certified in, tracks per slot
    int end;
};
// words 18 and 19 of `counts` (after the 18 message counters) carry the certificate's flags
enum { CW_FLAG = 18 /* CF_* */, CW_REMOVED = 19 /* slots the last general tick removed (bit mask, warp teams) */ };
enum {
    CF_WINDOWS = 1, /* the windows are void: rebuild them */
    CF_TRACKS = 2,  /* tracks changed hands or were created: recount the tracks per slot */
    CF_GENERAL = 4, /* (two-kernel stepping) the next tick of this scenario is not one the streak kernel can run */
    CF_HOLD_SHIFT = 8 /* bits 8..14: general ticks to run before the streak kernel is tried again */
};
ZRK_HD constexpr HotOff hot_off(int team) {
    HotOff h{};
    int p = 0;
    h.hdr = p; p += 4 * H_WORDS;
    h.counts = p; p += 80;  // 18 counters, padded to 16 bytes
    h.pxy = p; p += 16 * team;
    h.pzs = p; p += 16 * team;
    h.qxy = p; p += 16 * team;
    h.qzw = p; p += 16 * team;
    h.txy = p; p += 16 * team;
    h.tzt = p; p += 16 * team;
    h.vxy = p; p += 16 * team;
    h.sid = p; p += 4 * team;
    h.sflags = p; p += 4 * team;
    h.srf = p; p += 4 * team;
    h.own = p; p += 4 * team;
    h.te_key = p; p += 4 * team;
    h.te_ref = p; p += 4 * team;
    h.te_upd = p; p += 4 * team;
    h.te_follow = p; p += 4 * team;
    h.te_mid = p; p += 4 * team;
    h.te_upds = p; p += 8 * team;
    const int ct = team <= 256 ? team : 0;  // (larger teams run without the certificate)
    h.c_in = p; p += 4 * ct;
    h.c_out = p; p += 4 * ct;
    h.c_refc = p; p += 4 * ct;
    h.end = p;
    return h;
}

// threads per CTA for teams smaller than a CTA (TEAM = 32, 64)
#ifndef ZRK_CTA_THREADS
#define ZRK_CTA_THREADS 128
#endif

// Scratch of one team (shared memory only).  It sits in front of the dyn blob, so that its
// offsets from the blob base depend on the team size and the missile capacity only — compile-time
// constants for warp teams, whose kernel is instantiated per capacity bucket.
struct ScratchOff {
    int x_lxy, x_lzd, x_lf, x_claim, x_detlist, x_wscr, x_need, x_ghead, x_mbar;
    int bytes;
};
ZRK_HD constexpr int zrk_align_c(int x, int a) { return (x + a - 1) / a * a; }
ZRK_HD constexpr ScratchOff scratch_off(int team, int mp) {
    ScratchOff x{};
    int p = 0;
    x.x_lxy = p; p += 16 * (team + mp);
    x.x_lzd = p; p += 16 * (team + mp);
    x.x_lf = p; p += team > 32 ? 16 * (team + mp) : 12 * (team + zrk_align_c(mp, 8));
    x.x_claim = p; p += zrk_align_c(4 * (team + mp), 8);
    x.x_detlist = p; p += 4 * team;
    x.x_wscr = p; p += 4 * 64;
    x.x_need = p; p += 4 * team;
    x.x_ghead = p; p += 4 * team;
    x.x_mbar = p; p += 8;
    x.bytes = zrk_align_c(p, 16);
    return x;
}
// Byte offsets of the CCP missile entries, the first arrays of the cold part: hot_off(team).end + ...
struct MeOff {
    int me_key, me_ref, me_upd, me_tid, me_upds, rad_cur /* the array after them */, end;
};
ZRK_HD constexpr MeOff me_off(int team, int mp) {
    MeOff m{};
    int p = hot_off(team).end;
    m.me_key = p; p += zrk_align_c(4 * mp, 8);
    m.me_ref = p; p += zrk_align_c(4 * mp, 8);
    m.me_upd = p; p += zrk_align_c(4 * mp, 8);
    m.me_tid = p; p += zrk_align_c(4 * mp, 8);
    m.me_upds = p; p += 8 * mp;
    m.rad_cur = p;
    m.end = p;
    return m;
}
// missile capacity of a scenario's arrays: warp teams use one of two buckets
ZRK_HD constexpr int zrk_mp_for(int team, int M) {
    return team == 32 ? (M <= 8 ? 8 : 32) : zrk_align_c(M > 0 ? M : 1, 2);
}

struct Layout {
    int B, T, M, NS, R, L, C;
    int MP, RP, LP, CP, Q;
    int team, teams_per_cta;
    HotOff hot;
    // cold part of the dyn blob (byte offsets)
    int me_key, me_ref, me_upd, me_tid, me_upds;
    int rad_cur, st_n, st, cap, launched;
    int m_status, m_target, m_slot, m_birth, m_period;
    int det_mid, det_tid, det_flag, mb_type, mb_l, mb_a, ln_m;
    int dyn_bytes;
    // cst blob
    int c_hdr, c_rad_id, c_rad_pos, c_rad_par, c_l_id, c_l_pos;
    int c_m_id, c_m_launcher, c_m_v0, c_m_v0sq, c_m_radius;
    int c_cl_id, c_cl_launcher, c_cl_pos;
    int cst_bytes;
    int scratch_bytes;  // the team's region is [scratch][dyn blob][cst blob]; `sm` in the kernel is the dyn base
    // scratch (shared memory only): offsets from the dyn base, all negative (scratch_off(...) - scratch_bytes)
    int x_lxy, x_lzd;  // double2[TEAM + MP]: per CCP entry, the position link_object measures to
    int x_lf;          // the same positions rounded to fp32: warp teams float[12][(TEAM + MP8) / 4], blocks of
                       // x[4] y[4] z[4]; larger teams float4[TEAM + MP] = x, y, z, next entry in the grid chain
    int x_claim, x_detlist, x_wscr, x_need;  // x_need doubles as the launch result (x_ok)
    int x_ghead;       // int[TEAM]: bucket heads of the hash grid over the entries' link positions
    int x_mbar;            // one 8-byte mbarrier: completion of the bulk (TMA) load of the state
    int team_smem_bytes;
};

struct Params {
    Layout lo;
    int dt_ms;
    double dt_s;        // to_seconds(dt)                      (AirObject.py:5-7)
    double range_step;  // POSSIBLE_TARGET_RADIUS * to_seconds(dt) (CCP.py:186-187)
    unsigned log_types;
    int log_cap;
    int mask_ticks, mask_words;  // words per mask row; a tick's row set is (1 + R) rows
    int noise_period, noise_radars, noise_objects;
    long long noise_stride;
    double noise_bound;  // upper bound of the displacement the radars' noise can give an object in one tick (sum over radars of the
                         // largest |noise vector| of the table); < 0 = unknown: the association certificate is off
    int cert;            // association certificate enabled (zrk_tick.cuh)
    unsigned char *dyn;
    const unsigned char *cst;
    zrk_msg *rec;
    unsigned *masks;
    const double *noise;
    const double *tsec;  // tsec[t] = to_seconds(t * dt_ms) for t < tsec_n
    int tsec_n;
    const Params *self;  // this struct in device global memory (for out-of-line device functions)
};

static inline int zrk_align(int x, int a) { return (x + a - 1) / a * a; }

// Smallest supported team (threads per scenario) covering ns slots; 0 if too large.
static inline int zrk_team_for(int ns) {
    for (int t = 32; t <= 1024; t *= 2)
        if (ns <= t) return t;
    return 0;
}

// ns_used / m_used: the largest number of air-object slots (targets + missiles) and of missiles any
// scenario of the batch really has (the batch maxima T, M may be padding): they size the team and
// the per-missile arrays.  T and M stay the strides of the caller's arrays and of the static index.
static inline Layout make_layout(int B, int T, int M, int R, int L, int C, int ns_used = -1, int m_used = -1) {
    Layout o{};
    o.B = B; o.T = T; o.M = M; o.NS = T + M; o.R = R; o.L = L; o.C = C;
    if (ns_used < 0 || ns_used > o.NS) ns_used = o.NS;
    if (m_used < 0 || m_used > M) m_used = M;
    o.team = zrk_team_for(ns_used > 0 ? ns_used : 1);
    const int team = o.team ? o.team : 32;
    o.teams_per_cta = team > 32 ? 1 : ZRK_CTA_THREADS / 32;  // teams larger than a warp own their CTA
    o.MP = zrk_mp_for(team, m_used);
    o.RP = R > 0 ? R : 1;
    o.LP = zrk_align(L > 0 ? L : 1, 2);
    o.CP = zrk_align(C > 0 ? C : 1, 2);
    o.Q = zrk_align(2 * o.MP + o.CP, 2);
    o.hot = hot_off(team);
    int p = o.hot.end;
    auto take = [&p](int bytes) { int at = p; p += zrk_align(bytes, 8); return at; };
    {
        const MeOff m = me_off(team, o.MP);
        o.me_key = m.me_key; o.me_ref = m.me_ref; o.me_upd = m.me_upd; o.me_tid = m.me_tid; o.me_upds = m.me_upds;
        p = m.end;
    }
    o.rad_cur = take(8 * 2 * o.RP);
    o.st_n = take(4 * o.LP);
    o.st = take(4 * o.LP * o.MP);
    o.cap = take(4 * o.CP);
    o.launched = take(4 * o.CP);
    o.m_status = take(4 * o.MP); o.m_target = take(4 * o.MP); o.m_slot = take(4 * o.MP);
    o.m_birth = take(4 * o.MP);
    o.m_period = take(8 * o.MP);
    o.det_mid = take(4 * 2 * o.MP); o.det_tid = take(4 * 2 * o.MP); o.det_flag = take(4 * 2 * o.MP);
    o.mb_type = take(4 * 2 * o.Q); o.mb_l = take(4 * 2 * o.Q); o.mb_a = take(4 * 2 * o.Q);
    o.ln_m = take(4 * o.MP);
    o.dyn_bytes = zrk_align(p, 16);
    p = 0;
    o.c_hdr = take(4 * C_WORDS);
    o.c_rad_id = take(4 * o.RP);
    o.c_rad_pos = take(8 * 3 * o.RP);
    o.c_rad_par = take(8 * RAD_PAR_STRIDE * o.RP);
    o.c_l_id = take(4 * o.LP);
    o.c_l_pos = take(8 * 3 * o.LP);
    o.c_m_id = take(4 * o.MP); o.c_m_launcher = take(4 * o.MP);
    o.c_m_v0 = take(8 * o.MP); o.c_m_v0sq = take(8 * o.MP); o.c_m_radius = take(8 * o.MP);
    o.c_cl_id = take(4 * o.CP); o.c_cl_launcher = take(4 * o.CP);
    o.c_cl_pos = take(8 * 3 * o.CP);
    o.cst_bytes = zrk_align(p, 16);
    {
        const ScratchOff x = scratch_off(team, o.MP);
        o.scratch_bytes = x.bytes;
        o.x_lxy = x.x_lxy - x.bytes; o.x_lzd = x.x_lzd - x.bytes; o.x_lf = x.x_lf - x.bytes;
        o.x_claim = x.x_claim - x.bytes; o.x_detlist = x.x_detlist - x.bytes; o.x_wscr = x.x_wscr - x.bytes;
        o.x_need = x.x_need - x.bytes; o.x_ghead = x.x_ghead - x.bytes; o.x_mbar = x.x_mbar - x.bytes;
        o.team_smem_bytes = zrk_align(x.bytes + o.dyn_bytes + o.cst_bytes, 16);
    }
    return o;
}

}  // namespace zrk
#endif
