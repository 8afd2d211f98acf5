/*
 * zrk_oracle.c — plain-C, serial, fp64 restatement of the reference's per-tick loop.
 * TEST INFRASTRUCTURE ONLY (see zrk_oracle.h).  Compile with -ffp-contract=off: every
 * multiply-add below that must be fused is written as fma() explicitly.
 *
 * Arithmetic notes (measured in the build container, numpy 2.3.5 + scipy-openblas 0.3.30):
 *   np.dot / np.linalg.norm on 3-vectors go through OpenBLAS ddot, whose scalar tail is an FMA
 *   chain: dot = fma(x2,y2, fma(x1,y1, x0*y0)); norm = sqrt(dot(x,x)).
 *   np.degrees(x) = x * (180/pi);  float `%` = fmod + sign fix-up (result takes the sign of
 *   the modulus).
 */
#include "zrk_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- small helpers */

static double dot3(const double *x, const double *y) {
    return fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0]));
}
static double norm3(const double *x) { return sqrt(dot3(x, x)); }
static void sub3(double *o, const double *a, const double *b) {
    o[0] = a[0] - b[0];
    o[1] = a[1] - b[1];
    o[2] = a[2] - b[2];
}
/* Python float % for a positive modulus. */
static double pymod(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0) {
        if (r < 0.0) r += m;
    } else {
        r = copysign(0.0, m);
    }
    return r;
}
/* AirObject.to_seconds (AirObject.py:5-7): int ms / 1000 */
static double to_seconds(int64_t ms) { return (double)ms / 1000.0; }

/* ---------------------------------------------------------------- model state */

enum { ST_READY = 0, ST_ACTIVE = 1, ST_DETONATED = 2 };

typedef struct {
    int id, is_missile, ttype;
    double pos[3], prev[3];
    int has_prev;
    double p0[3], v[3], t0; /* Trajectory (AirObject.py:15-25) */
    double unit[3], speed;  /* AirObject.velocity / speed_mod (AirObject.py:35-36) */
    /* Missile (Missile.py:18-33) */
    int status, target, launcher;
    double period, radius, v0, v0sq;
} obj_t;

typedef struct {
    int id;
    double pos[3];
    double az_start, el_start, max_dist, az_range, el_range, az_speed, el_speed;
    int mode;
    double cur_az, cur_el;
} radar_t;

typedef struct {
    int id;
    double pos[3];
    int *stack; /* static indices of missiles, top = last (MissileLauncher.py:70) */
    int n;
} launcher_t;

typedef struct {
    int key, ref, upd, following, missile_id;
} tentry_t; /* TargetCCP (CCP.py:15-41) */
typedef struct {
    int key, ref, upd, target_id;
} mentry_t; /* MissileCCP (CCP.py:44-65) */

/* internal message: canonical fields + object references */
typedef struct {
    int type, sender, receiver, relevance, a, b, flag;
    double xyz[3];
    int ref;              /* static index of the object the message carries (target / missile) */
    int list_off, list_n; /* into sim_t.lists (static indices) */
} m_t;

typedef struct {
    m_t *m;
    int n, cap;
} bucket_t;

typedef struct {
    const zrk_batch_desc *d;
    int sc, flags;
    int dt_ms, tick;
    int airenv_id, ccp_id, has_ccp;
    int nT, nR, nL, nM, nC, NS;
    obj_t *obj;   /* [NS] by static index */
    int *slots;   /* AirEnv.__objects: static index or -1 (None) */
    int *slot_hist; /* who was appended k-th (survives removal) */
    int n_slots;
    radar_t *radar;
    launcher_t *launcher;
    /* CCP */
    int ccp_initialized;
    int *cl_id;
    double *cl_pos;
    int *cl_capacity, *cl_launched;
    tentry_t *te;
    int n_te;
    mentry_t *me;
    int n_me;
    /* message buckets of tick t and t-dt */
    bucket_t cur, prev;
    int *lists;
    int n_lists, cap_lists;
    int *prev_lists;
    int n_prev_lists, cap_prev_lists;
    int fault_code, fault_tick;
    zo_run *out;
    int64_t counts[ZRK_MSG_NTYPES];
} sim_t;

static void *xrealloc(void *p, size_t n) {
    void *q = realloc(p, n ? n : 1);
    if (!q) abort();
    return q;
}

/* Manager.add_message (Manager.py:41-58): append, then stable sort by -relevance. */
static m_t *add_msg(sim_t *s, int type, int sender, int receiver, int relevance) {
    bucket_t *b = &s->cur;
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 64;
        b->m = xrealloc(b->m, sizeof(m_t) * b->cap);
    }
    int at = b->n;
    while (at > 0 && b->m[at - 1].relevance < relevance) at--;
    memmove(&b->m[at + 1], &b->m[at], sizeof(m_t) * (b->n - at));
    b->n++;
    m_t *m = &b->m[at];
    memset(m, 0, sizeof *m);
    m->type = type;
    m->sender = sender;
    m->receiver = receiver;
    m->relevance = relevance;
    m->a = ZRK_NONE;
    m->b = ZRK_NONE;
    m->ref = -1;
    s->counts[type]++;
    return m;
}

static int add_list(sim_t *s, const int *idx, int n) {
    if (s->n_lists + n > s->cap_lists) {
        s->cap_lists = 2 * (s->n_lists + n) + 64;
        s->lists = xrealloc(s->lists, sizeof(int) * s->cap_lists);
    }
    int off = s->n_lists;
    memcpy(s->lists + off, idx, sizeof(int) * n);
    s->n_lists += n;
    return off;
}

/* ---------------------------------------------------------------- AirObject / Missile */

/* AirObject.step (AirObject.py:39-42) */
static void air_object_step(sim_t *s, obj_t *o) {
    double now = to_seconds((int64_t)s->tick * s->dt_ms);
    if (o->t0 != now) {
        memcpy(o->prev, o->pos, sizeof o->prev);
        o->has_prev = 1;
    } else {
        o->has_prev = 0;
    }
    double dtm = now - o->t0;
    for (int k = 0; k < 3; k++) o->pos[k] = o->p0[k] + o->v[k] * dtm; /* Trajectory.get_pos */
}

/* Missile._detonate (Missile.py:138-146) */
static void missile_detonate(sim_t *s, obj_t *m, int target_id, int self_det) {
    m_t *msg = add_msg(s, ZRK_MSG_MISSILE_DETONATE, m->id, ZRK_NONE, 1);
    msg->a = m->id;
    msg->b = target_id;
    msg->flag = self_det;
    m->status = ST_DETONATED;
}

/* Missile._calculate_trajectory_params (Missile.py:35-102); returns -1 or a cancel reason */
static int missile_solve(sim_t *s, const obj_t *m, const obj_t *g, double *V) {
    double d[3], vt[3];
    sub3(d, g->pos, m->pos);
    for (int k = 0; k < 3; k++) vt[k] = g->unit[k] * g->speed;
    double v0 = m->v0;
    double a = dot3(vt, vt) - m->v0sq;
    double b = 2.0 * dot3(d, vt);
    double c = dot3(d, d);
    double t;
    if (fabs(a) < 1e-6) {
        if (fabs(b) < 1e-6) return ZRK_CANCEL_PARALLEL;
        t = -c / b;
        if (t <= 0) return ZRK_CANCEL_PAST;
    } else {
        double bsq = (s->flags & ZO_NUMPY_POW) ? pow(b, 2.0) : b * b;
        double disc = bsq - 4.0 * a * c;
        if (disc < 0) return ZRK_CANCEL_NO_REAL_ROOT;
        double sq = sqrt(disc);
        double t1 = (-b + sq) / (2.0 * a);
        double t2 = (-b - sq) / (2.0 * a);
        int p1 = t1 > 0, p2 = t2 > 0;
        if (!p1 && !p2) return ZRK_CANCEL_NOT_POSITIVE;
        if (p1 && p2)
            t = (t2 < t1) ? t2 : t1; /* min([t1, t2]) keeps the first on ties */
        else
            t = p1 ? t1 : t2;
    }
    if (t > m->period) return ZRK_CANCEL_TOO_FAR;
    for (int k = 0; k < 3; k++) V[k] = d[k] / t + vt[k];
    double n = norm3(V);
    for (int k = 0; k < 3; k++) V[k] = V[k] / n * v0;
    return -1;
}

/* Missile._launch (Missile.py:104-133) */
static void missile_launch(sim_t *s, obj_t *m, int target, int launcher_id) {
    double V[3];
    int why = missile_solve(s, m, &s->obj[target], V);
    if (why < 0) {
        m->target = target;
        memcpy(m->v, V, sizeof V);
        memcpy(m->p0, m->pos, sizeof m->p0);
        m->t0 = to_seconds((int64_t)s->tick * s->dt_ms);
        m_t *msg = add_msg(s, ZRK_MSG_LAUNCH_SUCCESSFUL, m->id, launcher_id, 1);
        msg->a = m->id;
        msg->b = s->obj[target].id;
        memcpy(msg->xyz, V, sizeof V);
        msg->ref = (int)(m - s->obj);
        m->status = ST_ACTIVE;
    } else {
        m_t *msg = add_msg(s, ZRK_MSG_LAUNCH_CANCELLED, m->id, launcher_id, 1);
        msg->a = m->id;
        msg->b = why;
        msg->ref = (int)(m - s->obj);
    }
}

/* Missile.step, status 'active' (Missile.py:162-193).  The UPDATE_TARGET loop (:163-179)
 * never runs: nothing emits that type. */
static void missile_step_active(sim_t *s, obj_t *m) {
    air_object_step(s, m);
    m_t *msg = add_msg(s, ZRK_MSG_MISSILE_POS, m->id, ZRK_NONE, 1);
    msg->a = m->id;
    double d[3];
    sub3(d, s->obj[m->target].pos, m->pos);
    if (norm3(d) <= m->radius) {
        missile_detonate(s, m, s->obj[m->target].id, 0);
        return;
    }
    m->period -= to_seconds(s->dt_ms);
    if (m->period <= 0) missile_detonate(s, m, ZRK_NONE, 1);
}

/* ---------------------------------------------------------------- AirEnv */

/* AirEnv.step (AirEnv.py:26-53) */
static void airenv_step(sim_t *s) {
    /* :33-40 ids named by last tick's MISSILE_DETONATE become None */
    for (int i = 0; i < s->prev.n; i++) {
        const m_t *m = &s->prev.m[i];
        if (m->type != ZRK_MSG_MISSILE_DETONATE) continue;
        for (int k = 0; k < s->n_slots; k++) {
            int o = s->slots[k];
            if (o < 0) continue;
            if (s->obj[o].id == m->a || (m->b != ZRK_NONE && s->obj[o].id == m->b)) s->slots[k] = -1;
        }
    }
    /* :42-43 */
    for (int i = 0; i < s->prev.n; i++)
        if (s->prev.m[i].type == ZRK_MSG_NEW_MISSILE) {
            s->slot_hist[s->n_slots] = s->prev.m[i].ref;
            s->slots[s->n_slots++] = s->prev.m[i].ref;
        }
    /* :45-48 */
    for (int k = 0; k < s->n_slots; k++) {
        int o = s->slots[k];
        if (o < 0) continue;
        obj_t *ob = &s->obj[o];
        if (!ob->is_missile)
            air_object_step(s, ob);
        else if (ob->status == ST_ACTIVE)
            missile_step_active(s, ob);
        /* 'ready' cannot be in the list; 'detonated' does nothing (Missile.py:195-196) */
    }
    /* :50-53 */
    int *live = malloc(sizeof(int) * (s->n_slots + 1));
    int n = 0;
    for (int k = 0; k < s->n_slots; k++)
        if (s->slots[k] >= 0) live[n++] = s->slots[k];
    m_t *msg = add_msg(s, ZRK_MSG_ACTIVE_OBJECTS, s->airenv_id, ZRK_NONE, 1);
    msg->list_off = add_list(s, live, n);
    msg->list_n = n;
    free(live);
}

/* ---------------------------------------------------------------- SectorRadar */

/* SectorRadar.find_visible_objects, one object (Radar.py:53-71) */
static int radar_sees(const radar_t *r, const obj_t *o) {
    double delta[3];
    sub3(delta, o->pos, r->pos);
    double distance = norm3(delta);
    if (distance > r->max_dist) return 0;
    double az = pymod(atan2(delta[1], delta[0]) * (180.0 / M_PI), 360.0);
    double el = pymod(asin(delta[2] / distance) * (180.0 / M_PI), 180.0);
    return r->cur_az <= az && az <= r->cur_az + r->az_range && r->cur_el <= el &&
           el <= r->cur_el + r->el_range;
}

/* SectorRadar.move_to_next_sector_circular (Radar.py:96-117) */
static void radar_move(radar_t *r) {
    if (r->mode == ZRK_SCAN_HORIZONTAL) {
        if (r->cur_az + r->az_range < 360)
            r->cur_az = pymod(r->cur_az + r->az_speed, 360.0);
        else
            r->cur_az = r->el_start; /* sic, Radar.py:105 */
        if (r->cur_az < r->az_speed) {
            if (r->cur_el + r->el_speed < 90)
                r->cur_el = pymod(r->cur_el + r->el_speed, 90.0);
            else
                r->cur_el = r->el_start;
        }
    } else if (r->mode == ZRK_SCAN_VERTICAL) {
        r->cur_el = pymod(r->cur_el + r->el_speed, 90.0);
        if (r->cur_el < r->el_speed) r->cur_az = pymod(r->cur_az + r->az_speed, 360.0);
    }
}

/* SectorRadar.step (Radar.py:144-205) */
static void radar_step(sim_t *s, int ri) {
    radar_t *r = &s->radar[ri];
    /* :151 first ACTIVE_OBJECTS message of this tick */
    const m_t *act = NULL;
    for (int i = 0; i < s->cur.n; i++)
        if (s->cur.m[i].type == ZRK_MSG_ACTIVE_OBJECTS) {
            act = &s->cur.m[i];
            break;
        }
    int n = act->list_n, off = act->list_off;
    m_t *all = add_msg(s, ZRK_MSG_ALL_OBJECTS, r->id, ZRK_CCP_ID, 1);
    all->list_off = off; /* the same list object */
    all->list_n = n;
    int *vis = malloc(sizeof(int) * (n + 1));
    int nv = 0;
    for (int i = 0; i < n; i++) {
        int o = s->lists[off + i];
        if (radar_sees(r, &s->obj[o])) vis[nv++] = o;
    }
    /* smooth_objects (Radar.py:138-142) with the injected noise table */
    const zrk_batch_desc *d = s->d;
    if (d->noise) {
        int w = s->tick % d->noise_period;
        for (int i = 0; i < nv; i++) {
            obj_t *o = &s->obj[vis[i]];
            const double *nz = d->noise + (size_t)s->sc * d->noise_scenario_stride +
                               (((size_t)w * d->noise_radars + ri) * d->noise_objects + vis[i]) * 3;
            for (int k = 0; k < 3; k++) o->pos[k] += nz[k];
        }
    }
    m_t *found = add_msg(s, ZRK_MSG_FOUND_OBJECTS, r->id, ZRK_CCP_ID, 1);
    found->list_off = add_list(s, vis, nv);
    found->list_n = nv;
    free(vis);
    /* :179-190 UPDATE_TARGET forwarding: no producer exists */
    /* :193-203 */
    for (int i = 0; i < s->prev.n; i++) {
        const m_t *m = &s->prev.m[i];
        if (m->type != ZRK_MSG_MISSILE_DETONATE) continue;
        m_t *dm = add_msg(s, ZRK_MSG_DESTROYED_MISSILE, r->id, ZRK_CCP_ID, 1);
        dm->a = m->a;
        dm->flag = m->flag;
    }
    radar_move(r);
}

/* ---------------------------------------------------------------- MissileLauncher */

/* MissileLauncher.launch_missile (MissileLauncher.py:58-80) */
static void launcher_launch(sim_t *s, launcher_t *l, int target) {
    if (l->n == 0) return;
    int mi = l->stack[--l->n];
    obj_t *m = &s->obj[mi];
    m_t *msg = add_msg(s, ZRK_MSG_LAUNCH_MISSILE, l->id, m->id, 1);
    msg->a = s->obj[target].id;
    msg->ref = target;
    /* missile.step(), status 'ready' (Missile.py:153-160) */
    if (m->status == ST_READY) missile_launch(s, m, target, l->id);
    else if (m->status == ST_ACTIVE) missile_step_active(s, m); /* unreachable in practice */
}

/* MissileLauncher.step (MissileLauncher.py:82-138) */
static void launcher_step(sim_t *s, int li) {
    launcher_t *l = &s->launcher[li];
    int n_prev = s->prev.n;
    for (int i = 0; i < n_prev; i++) {
        const m_t m = s->prev.m[i];
        if (m.receiver != l->id) continue;
        switch (m.type) {
        case ZRK_MSG_LAUNCH_COMMAND:
            launcher_launch(s, l, m.ref);
            break;
        case ZRK_MSG_LAUNCH_SUCCESSFUL: {
            m_t *a = add_msg(s, ZRK_MSG_LAUNCHED_MISSILE, l->id, ZRK_CCP_ID, 1);
            a->a = m.a;
            a->b = m.b;
            a->ref = m.ref;
            m_t *b = add_msg(s, ZRK_MSG_NEW_MISSILE, l->id, ZRK_NONE, 1);
            b->a = m.a;
            b->ref = m.ref;
            break;
        }
        case ZRK_MSG_LAUNCH_CANCELLED:
            l->stack[l->n++] = m.ref;
            break;
        case ZRK_MSG_MISSILE_COUNT_REQUEST: {
            m_t *a = add_msg(s, ZRK_MSG_MISSILE_COUNT_RESPONSE, l->id, m.sender, 1);
            a->b = l->n;
            break;
        }
        default:
            break;
        }
    }
}

/* ---------------------------------------------------------------- CombatControlPoint */

/* CombatControlPoint.link_object (CCP.py:171-219): 0 NEW_TARGET, 1 OLD_TARGET, 2 OLD_ROCKET */
static int ccp_link(sim_t *s, const obj_t *det, int *matched) {
    double cur_time = to_seconds((int64_t)s->tick * s->dt_ms);
    double sim_step = to_seconds(s->dt_ms);
    double curr = INFINITY;
    int cls = 0;
    *matched = -1;
    for (int pass = 0; pass < 2; pass++) {
        int n = pass ? s->n_me : s->n_te;
        for (int e = 0; e < n; e++) {
            int upd = pass ? s->me[e].upd : s->te[e].upd;
            if (upd == s->tick) continue;
            const obj_t *o = &s->obj[pass ? s->me[e].ref : s->te[e].ref];
            const double *lp = (o->has_prev || !pass) ? o->prev : o->pos;
            double dd[3];
            sub3(dd, lp, det->pos);
            double diff = norm3(dd);
            double d_t = cur_time - to_seconds((int64_t)upd * s->dt_ms);
            double lo = det->speed * (d_t - ZRK_POSSIBLE_TARGET_RADIUS * sim_step);
            double hi = det->speed * (d_t + ZRK_POSSIBLE_TARGET_RADIUS * sim_step);
            lo = lo > 0 ? lo : 0; /* max(0, x) */
            hi = hi > 0 ? hi : 0;
            if (diff < curr && lo <= diff && diff <= hi) {
                curr = diff;
                cls = pass ? 2 : 1;
                *matched = e;
            }
        }
    }
    return cls;
}

/* CombatControlPoint.try_to_launch_missile (CCP.py:287-320) */
static int ccp_try_launch(sim_t *s, int oi, int radar_id) {
    const obj_t *o = &s->obj[oi];
    double min_dist = INFINITY;
    int best = -1;
    for (int c = 0; c < s->nC; c++) {
        if (s->cl_launched[c] < s->cl_capacity[c]) {
            double q[3];
            sub3(q, &s->cl_pos[3 * c], o->pos);
            double ss = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]; /* np.sum(x ** 2) */
            double dist = (s->flags & ZO_NUMPY_POW) ? pow(ss, 0.5) : sqrt(ss);
            if (dist < min_dist) {
                best = c;
                min_dist = dist;
            }
        }
    }
    if (best < 0) return 0;
    s->cl_launched[best]++;
    m_t *m = add_msg(s, ZRK_MSG_LAUNCH_COMMAND, s->ccp_id, s->cl_id[best], 1);
    m->a = o->id;
    m->b = radar_id;
    memcpy(m->xyz, o->pos, sizeof m->xyz);
    m->ref = oi;
    return 1;
}

/* CombatControlPoint.add_target (CCP.py:88-93): dict assignment keeps an existing key's place */
static void ccp_add_target(sim_t *s, int oi, int following) {
    int key = s->obj[oi].id, e;
    for (e = 0; e < s->n_te; e++)
        if (s->te[e].key == key) break;
    if (e == s->n_te) s->n_te++;
    s->te[e].key = key;
    s->te[e].ref = oi;
    s->te[e].upd = s->tick;
    s->te[e].following = following;
    s->te[e].missile_id = ZRK_NONE;
}

static int is_in(const int *v, int n, int x) {
    for (int i = 0; i < n; i++)
        if (v[i] == x) return 1;
    return 0;
}

/* CombatControlPoint.step (CCP.py:368-431) */
static void ccp_step(sim_t *s) {
    /* :382-384 send_request_msg_to_ml_capacity (CCP.py:122-137) */
    if (!s->ccp_initialized) {
        for (int c = 0; c < s->nC; c++) {
            s->cl_capacity[c] = 0;
            s->cl_launched[c] = 0;
            add_msg(s, ZRK_MSG_MISSILE_COUNT_REQUEST, s->ccp_id, s->cl_id[c], 3);
        }
        s->ccp_initialized = 1;
    }
    /* :386 get_current_missile_launcher_capacity (CCP.py:139-147): dict[msg.sender_id] = count;
       a sender that is not yet a key would add one, but every launcher that answers was asked,
       i.e. already is a key. */
    for (int i = 0; i < s->cur.n; i++) {
        const m_t *m = &s->cur.m[i];
        if (m->type != ZRK_MSG_MISSILE_COUNT_RESPONSE) continue;
        for (int c = 0; c < s->nC; c++)
            if (s->cl_id[c] == m->sender) s->cl_capacity[c] = m->b;
    }
    /* :387 check_if_missile_get_hit (CCP.py:149-159): `(id,) in dict` is never true. */
    /* :388 check_if_missiles_launched (CCP.py:161-169) + add_missile (:102-109) */
    int n_cur = s->cur.n;
    for (int i = 0; i < n_cur; i++) {
        const m_t m = s->cur.m[i];
        if (m.type != ZRK_MSG_LAUNCHED_MISSILE) continue;
        const obj_t *mo = &s->obj[m.ref];
        int e;
        for (e = 0; e < s->n_me; e++)
            if (s->me[e].key == mo->id) break;
        if (e == s->n_me) s->n_me++;
        s->me[e].key = mo->id;
        s->me[e].ref = m.ref;
        s->me[e].upd = s->tick;
        s->me[e].target_id = s->obj[mo->target].id;
        int tid = s->obj[mo->target].id, t;
        for (t = 0; t < s->n_te; t++)
            if (s->te[t].key == tid) break;
        if (t == s->n_te) {
            s->fault_code = ZRK_FAULT_KEY_ERROR_CCP107;
            return;
        }
        s->te[t].missile_id = mo->id;
    }
    /* :390-404 to_visualize: union of ALL_OBJECTS lists, first-seen order, by id */
    int cap = 0;
    for (int i = 0; i < s->cur.n; i++)
        if (s->cur.m[i].type == ZRK_MSG_ALL_OBJECTS) cap += s->cur.m[i].list_n;
    int *viz = malloc(sizeof(int) * (cap + 1));
    int *viz_id = malloc(sizeof(int) * (cap + 1));
    int n_viz = 0;
    for (int i = 0; i < s->cur.n; i++) {
        const m_t *m = &s->cur.m[i];
        if (m->type != ZRK_MSG_ALL_OBJECTS) continue;
        for (int k = 0; k < m->list_n; k++) {
            int o = s->lists[m->list_off + k];
            if (!is_in(viz_id, n_viz, s->obj[o].id)) {
                viz[n_viz] = o;
                viz_id[n_viz++] = s->obj[o].id;
            }
        }
    }
    /* :409-429 */
    int *proc = malloc(sizeof(int) * (cap + 1));
    int n_proc = 0;
    n_cur = s->cur.n;
    for (int i = 0; i < n_cur; i++) {
        if (s->cur.m[i].type != ZRK_MSG_FOUND_OBJECTS) continue;
        int radar_id = s->cur.m[i].sender;
        int off = s->cur.m[i].list_off, n = s->cur.m[i].list_n;
        for (int k = 0; k < n; k++) {
            int oi = s->lists[off + k];
            const obj_t *o = &s->obj[oi];
            if (is_in(proc, n_proc, o->id)) continue;
            proc[n_proc++] = o->id;
            int e, cls = ccp_link(s, o, &e);
            if (cls == 0) { /* new_target (CCP.py:322-330) */
                ccp_add_target(s, oi, ccp_try_launch(s, oi, radar_id));
            } else if (cls == 1) { /* old_target (CCP.py:332-359) */
                tentry_t *t = &s->te[e];
                if (!t->following) {
                    int ok = ccp_try_launch(s, oi, radar_id);
                    t->ref = oi;
                    t->upd = s->tick;
                    t->following = ok;
                } else {
                    t->ref = oi;
                    t->upd = s->tick;
                    m_t *m = add_msg(s, ZRK_MSG_CCP_UPDATE_TARGET, s->ccp_id, radar_id, 1);
                    m->a = o->id;
                    m->b = t->missile_id;
                    m->ref = oi;
                }
            } else { /* old_rocket (CCP.py:361-366) */
                s->me[e].ref = oi;
                s->me[e].upd = s->tick;
            }
        }
    }
    /* :431 send_objects_to_GUI (CCP.py:237-284) */
    for (int e = 0; e < s->n_me; e++) {
        const obj_t *o = &s->obj[s->me[e].ref];
        if (!is_in(proc, n_proc, o->id)) continue;
        m_t *m = add_msg(s, ZRK_MSG_DRAW_OBJECTS, s->ccp_id, ZRK_MANAGER_ID, 1);
        m->a = o->id;
        m->b = ZRK_ZUR;
        m->flag = 1;
        memcpy(m->xyz, o->pos, sizeof m->xyz);
    }
    for (int e = 0; e < s->n_te && !s->fault_code; e++) {
        const obj_t *o = &s->obj[s->te[e].ref];
        if (!is_in(proc, n_proc, o->id)) continue;
        if (o->is_missile) { /* .type on a Missile (CCP.py:259) */
            s->fault_code = ZRK_FAULT_ATTRIBUTE_ERROR_CCP259;
            break;
        }
        m_t *m = add_msg(s, ZRK_MSG_DRAW_OBJECTS, s->ccp_id, ZRK_MANAGER_ID, 1);
        m->a = o->id;
        m->b = o->ttype;
        m->flag = 1;
        memcpy(m->xyz, o->pos, sizeof m->xyz);
    }
    for (int i = 0; i < n_viz && !s->fault_code; i++) {
        const obj_t *o = &s->obj[viz[i]];
        if (is_in(proc, n_proc, o->id)) continue;
        m_t *m = add_msg(s, ZRK_MSG_DRAW_OBJECTS, s->ccp_id, ZRK_MANAGER_ID, 1);
        m->a = o->id;
        m->b = o->is_missile ? ZRK_ZUR : o->ttype;
        m->flag = 0;
        memcpy(m->xyz, o->pos, sizeof m->xyz);
    }
    free(viz);
    free(viz_id);
    free(proc);
}

/* ---------------------------------------------------------------- Manager */

static void flush_bucket(sim_t *s) {
    zo_run *r = s->out;
    if (!r || !(s->flags & ZO_KEEP_MSGS)) return;
    for (int i = 0; i < s->cur.n; i++) {
        const m_t *m = &s->cur.m[i];
        if (r->n_msgs == r->cap_msgs) {
            r->cap_msgs = r->cap_msgs ? 2 * r->cap_msgs : 1024;
            r->msgs = xrealloc(r->msgs, sizeof(zrk_msg) * r->cap_msgs);
        }
        zrk_msg *o = &r->msgs[r->n_msgs++];
        memset(o, 0, sizeof *o);
        o->tick = s->tick;
        o->type = (uint8_t)m->type;
        o->relevance = (uint8_t)m->relevance;
        o->flag = (uint8_t)m->flag;
        o->sender = m->sender;
        o->receiver = m->receiver;
        o->a = m->a;
        o->b = m->b;
        memcpy(o->xyz, m->xyz, sizeof o->xyz);
        if (m->type == ZRK_MSG_ACTIVE_OBJECTS || m->type == ZRK_MSG_ALL_OBJECTS ||
            m->type == ZRK_MSG_FOUND_OBJECTS) {
            if (r->n_ids + m->list_n > r->cap_ids) {
                r->cap_ids = 2 * (r->n_ids + m->list_n) + 1024;
                r->ids = xrealloc(r->ids, sizeof(int32_t) * r->cap_ids);
            }
            o->a = (int32_t)r->n_ids;
            o->b = m->list_n;
            for (int k = 0; k < m->list_n; k++)
                r->ids[r->n_ids++] = s->obj[s->lists[m->list_off + k]].id;
        }
    }
}

/* One pass of the loop body of Manager.run_simulation (Manager.py:116-140). */
static void manager_tick(sim_t *s) {
    airenv_step(s);
    for (int r = 0; r < s->nR; r++) radar_step(s, r);
    for (int l = 0; l < s->nL; l++) launcher_step(s, l);
    if (s->has_ccp) ccp_step(s);
    if (s->fault_code) s->fault_tick = s->tick;
    flush_bucket(s);
    /* rotate buckets: t becomes t-dt */
    bucket_t tb = s->prev;
    s->prev = s->cur;
    s->cur = tb;
    s->cur.n = 0;
    int *tl = s->prev_lists;
    int tc = s->cap_prev_lists;
    s->prev_lists = s->lists;
    s->cap_prev_lists = s->cap_lists;
    s->n_prev_lists = s->n_lists;
    s->lists = tl;
    s->cap_lists = tc;
    s->n_lists = 0;
    s->tick++;
}

#define CNT(field, maxv) (d->field ? d->field[sc] : d->maxv)

static void sim_init(sim_t *s, const zrk_batch_desc *d, int sc, int flags, zo_run *out) {
    memset(s, 0, sizeof *s);
    s->d = d;
    s->sc = sc;
    s->flags = flags;
    s->out = out;
    s->dt_ms = d->dt_ms;
    s->nT = CNT(n_targets, max_targets);
    s->nR = CNT(n_radars, max_radars);
    s->nL = CNT(n_launchers, max_launchers);
    s->nM = CNT(n_missiles, max_missiles);
    s->nC = CNT(n_ccp_launchers, max_ccp_launchers);
    s->NS = d->max_targets + d->max_missiles;
    s->airenv_id = d->airenv_id ? d->airenv_id[sc] : 1;
    s->ccp_id = d->ccp_id ? d->ccp_id[sc] : 0;
    s->has_ccp = s->ccp_id != ZRK_NONE;
    s->obj = calloc(s->NS + 1, sizeof(obj_t));
    s->slots = malloc(sizeof(int) * (s->NS + 1));
    s->slot_hist = malloc(sizeof(int) * (s->NS + 1));
    s->radar = calloc(s->nR + 1, sizeof(radar_t));
    s->launcher = calloc(s->nL + 1, sizeof(launcher_t));
    s->cl_id = malloc(sizeof(int) * (s->nC + 1));
    s->cl_pos = malloc(sizeof(double) * 3 * (s->nC + 1));
    s->cl_capacity = calloc(s->nC + 1, sizeof(int));
    s->cl_launched = calloc(s->nC + 1, sizeof(int));
    s->te = calloc(s->NS + 1, sizeof(tentry_t));
    s->me = calloc(s->NS + 1, sizeof(mentry_t));
    for (int i = 0; i < s->NS; i++) s->obj[i].target = -1;
    for (int t = 0; t < s->nT; t++) { /* main.py:128-147 */
        size_t g = (size_t)sc * d->max_targets + t;
        obj_t *o = &s->obj[t];
        o->id = d->target_id[g];
        o->ttype = d->target_type ? d->target_type[g] : ZRK_ANOTHER;
        for (int k = 0; k < 3; k++) {
            o->p0[k] = o->pos[k] = d->target_pos[3 * g + k];
            o->v[k] = d->target_vel[3 * g + k];
        }
        o->t0 = 0.0;
        o->speed = norm3(o->v);
        for (int k = 0; k < 3; k++) o->unit[k] = o->v[k] / o->speed;
        s->slot_hist[s->n_slots] = t;
        s->slots[s->n_slots++] = t; /* AirEnv.add_target (AirEnv.py:55-61) */
    }
    for (int r = 0; r < s->nR; r++) { /* main.py:56-71 */
        size_t g = (size_t)sc * d->max_radars + r;
        radar_t *q = &s->radar[r];
        const double *p = d->radar_par + 8 * g;
        q->id = d->radar_id[g];
        memcpy(q->pos, d->radar_pos + 3 * g, sizeof q->pos);
        q->az_start = p[0];
        q->el_start = p[1];
        q->max_dist = p[2];
        q->az_range = p[3];
        q->el_range = p[4];
        q->az_speed = p[5];
        q->el_speed = p[6];
        q->mode = (int)p[7];
        q->cur_az = q->az_start;
        q->cur_el = q->el_start;
    }
    for (int l = 0; l < s->nL; l++) { /* main.py:74-101 */
        size_t g = (size_t)sc * d->max_launchers + l;
        s->launcher[l].id = d->launcher_id[g];
        memcpy(s->launcher[l].pos, d->launcher_pos + 3 * g, sizeof(double) * 3);
        s->launcher[l].stack = malloc(sizeof(int) * (s->nM + 1));
    }
    for (int m = 0; m < s->nM; m++) { /* Missile.__init__ (Missile.py:18-33) */
        size_t g = (size_t)sc * d->max_missiles + m;
        obj_t *o = &s->obj[d->max_targets + m];
        launcher_t *l = &s->launcher[d->missile_launcher[g]];
        o->id = d->missile_id[g];
        o->is_missile = 1;
        o->launcher = d->missile_launcher[g];
        memcpy(o->pos, l->pos, sizeof o->pos);
        memcpy(o->p0, l->pos, sizeof o->p0);
        o->v0 = d->missile_v0[g];
        o->v0sq = d->missile_v0sq ? d->missile_v0sq[g] : o->v0 * o->v0;
        o->radius = d->missile_radius[g];
        o->period = d->missile_life[g];
        o->speed = o->v0; /* Missile.py:28 */
        o->unit[0] = o->unit[1] = o->unit[2] = NAN; /* 0/0 at AirObject.py:35 */
        o->status = ST_READY;
        l->stack[l->n++] = d->max_targets + m;
    }
    for (int c = 0; c < s->nC; c++) {
        size_t g = (size_t)sc * d->max_ccp_launchers + c;
        s->cl_id[c] = d->ccp_launcher_id[g];
        memcpy(&s->cl_pos[3 * c], d->ccp_launcher_pos + 3 * g, sizeof(double) * 3);
    }
}

static void sim_free(sim_t *s) {
    for (int l = 0; l < s->nL; l++) free(s->launcher[l].stack);
    free(s->obj);
    free(s->slots);
    free(s->slot_hist);
    free(s->radar);
    free(s->launcher);
    free(s->cl_id);
    free(s->cl_pos);
    free(s->cl_capacity);
    free(s->cl_launched);
    free(s->te);
    free(s->me);
    free(s->cur.m);
    free(s->prev.m);
    free(s->lists);
    free(s->prev_lists);
}

static void sim_run(sim_t *s, int n_ticks) {
    const zrk_batch_desc *d = s->d;
    int limit = d->tick_limit ? d->tick_limit[s->sc] : n_ticks;
    if (limit > n_ticks) limit = n_ticks;
    while (s->tick < limit && !s->fault_code) manager_tick(s);
}

static void export_state(const sim_t *s, const zrk_state_view *v) {
    const zrk_batch_desc *d = s->d;
    int NS = s->NS;
    for (int k = 0; k < NS; k++) {
        if (v->slot_obj) v->slot_obj[k] = -1;
        if (v->slot_alive) v->slot_alive[k] = 0;
    }
    /* slot k holds the k-th object ever appended; a slot set to None keeps its place */
    for (int k = 0; k < s->n_slots; k++) {
        if (v->slot_obj) v->slot_obj[k] = s->slot_hist[k];
        if (v->slot_alive) v->slot_alive[k] = s->slots[k] >= 0;
    }
    for (int i = 0; i < NS; i++) {
        const obj_t *o = &s->obj[i];
        for (int k = 0; k < 3; k++) {
            if (v->pos) v->pos[3 * i + k] = o->pos[k];
            if (v->prev_pos) v->prev_pos[3 * i + k] = o->has_prev ? o->prev[k] : 0.0;
        }
        if (v->has_prev) v->has_prev[i] = (uint8_t)o->has_prev;
    }
    for (int r = 0; r < d->max_radars; r++) {
        if (v->radar_az) v->radar_az[r] = r < s->nR ? s->radar[r].cur_az : 0.0;
        if (v->radar_el) v->radar_el[r] = r < s->nR ? s->radar[r].cur_el : 0.0;
    }
    for (int m = 0; m < d->max_missiles; m++) {
        const obj_t *o = &s->obj[d->max_targets + m];
        int live = m < s->nM;
        if (v->missile_status) v->missile_status[m] = live ? (uint8_t)o->status : 0;
        if (v->missile_period) v->missile_period[m] = live ? o->period : 0.0;
        if (v->missile_t0) v->missile_t0[m] = live ? o->t0 : 0.0;
        if (v->missile_target) v->missile_target[m] = live ? o->target : -1;
        for (int k = 0; k < 3; k++)
            if (v->missile_vel) v->missile_vel[3 * m + k] = live ? o->v[k] : 0.0;
    }
    for (int l = 0; l < d->max_launchers; l++)
        if (v->launcher_count) v->launcher_count[l] = l < s->nL ? s->launcher[l].n : 0;
    if (v->ccp_n_targets) v->ccp_n_targets[0] = s->n_te;
    if (v->ccp_n_missiles) v->ccp_n_missiles[0] = s->n_me;
    for (int e = 0; e < NS; e++) {
        int lt = e < s->n_te, lm = e < s->n_me;
        if (v->ccp_t_key) v->ccp_t_key[e] = lt ? s->te[e].key : 0;
        if (v->ccp_t_ref) v->ccp_t_ref[e] = lt ? s->te[e].ref : -1;
        if (v->ccp_t_upd) v->ccp_t_upd[e] = lt ? s->te[e].upd : 0;
        if (v->ccp_t_missile) v->ccp_t_missile[e] = lt ? s->te[e].missile_id : ZRK_NONE;
        if (v->ccp_t_following) v->ccp_t_following[e] = lt ? (uint8_t)s->te[e].following : 0;
        if (v->ccp_m_key) v->ccp_m_key[e] = lm ? s->me[e].key : 0;
        if (v->ccp_m_ref) v->ccp_m_ref[e] = lm ? s->me[e].ref : -1;
        if (v->ccp_m_upd) v->ccp_m_upd[e] = lm ? s->me[e].upd : 0;
        if (v->ccp_m_target) v->ccp_m_target[e] = lm ? s->me[e].target_id : ZRK_NONE;
    }
    for (int c = 0; c < d->max_ccp_launchers; c++) {
        if (v->ccp_capacity) v->ccp_capacity[c] = c < s->nC ? s->cl_capacity[c] : 0;
        if (v->ccp_launched) v->ccp_launched[c] = c < s->nC ? s->cl_launched[c] : 0;
    }
}

int zrk_oracle_run(const zrk_batch_desc *d, int scenario, int n_ticks, int flags, zo_run *out,
                   const zrk_state_view *state) {
    if (!d || scenario < 0 || scenario >= d->n_scenarios || n_ticks < 0) return ZRK_ERR_ARG;
    sim_t s;
    if (out) memset(out, 0, sizeof *out);
    sim_init(&s, d, scenario, flags, out);
    sim_run(&s, n_ticks);
    if (out) {
        out->fault_code = s.fault_code;
        out->fault_tick = s.fault_tick;
        memcpy(out->counts, s.counts, sizeof s.counts);
    }
    if (state) export_state(&s, state);
    sim_free(&s);
    return ZRK_OK;
}

void zrk_oracle_free(zo_run *r) {
    if (!r) return;
    free(r->msgs);
    free(r->ids);
    memset(r, 0, sizeof *r);
}

typedef struct {
    const zrk_batch_desc *d;
    int s0, s1, n_ticks, flags;
    int64_t *counts;
    int32_t *fault_code, *fault_tick;
    int *next;
    pthread_mutex_t *mu;
} job_t;

static void *worker(void *arg) {
    job_t *j = arg;
    for (;;) {
        pthread_mutex_lock(j->mu);
        int lo = *j->next;
        *j->next = lo + 16;
        pthread_mutex_unlock(j->mu);
        if (lo >= j->s1) break;
        int hi = lo + 16 < j->s1 ? lo + 16 : j->s1;
        for (int sc = lo; sc < hi; sc++) {
            sim_t s;
            sim_init(&s, j->d, sc, j->flags & ~ZO_KEEP_MSGS, NULL);
            sim_run(&s, j->n_ticks);
            if (j->counts)
                memcpy(j->counts + (size_t)(sc - j->s0) * ZRK_MSG_NTYPES, s.counts, sizeof s.counts);
            if (j->fault_code) j->fault_code[sc - j->s0] = s.fault_code;
            if (j->fault_tick) j->fault_tick[sc - j->s0] = s.fault_tick;
            sim_free(&s);
        }
    }
    return NULL;
}

int zrk_oracle_run_batch(const zrk_batch_desc *d, int s0, int s1, int n_ticks, int flags,
                         int n_threads, int64_t *counts, int32_t *fault_code, int32_t *fault_tick) {
    if (!d || s0 < 0 || s1 > d->n_scenarios || s0 > s1 || n_threads < 1) return ZRK_ERR_ARG;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = s0;
    job_t j = {d, s0, s1, n_ticks, flags, counts, fault_code, fault_tick, &next, &mu};
    pthread_t *th = malloc(sizeof(pthread_t) * n_threads);
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, worker, &j);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
    free(th);
    return ZRK_OK;
}
