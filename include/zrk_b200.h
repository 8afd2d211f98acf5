/*
 * zrk_b200.h — C ABI of the B200-native batched tick of Ollegorii/ZRK_modulation.
 *
 * The reference has no FFI: its seam is the Python class interface
 * `BaseModel.step()` (modules/BaseModel.py:13-44) + `Manager`
 * (modules/Manager.py:9-140).  This header is the boundary a host (ctypes, cgo,
 * a PyTorch custom op ...) binds to replace the per-tick loop
 * `Manager.run_simulation` (modules/Manager.py:111-140) and every `step()` it
 * calls, for B independent scenarios at once.  Plain pointers and sizes only.
 *
 * The same structs are consumed by the CPU oracle under oracle/ (test
 * infrastructure only) so that parity tests compare like with like.
 *
 * Entry point                      replaces (reference file:line)
 * -------------------------------  -------------------------------------------
 * zrk_create                       main.create_objects_from_config  main.py:35-149
 * zrk_step                         Manager.run_simulation loop body Manager.py:116-140
 *                                  (AirEnv.py:26-53, Radar.py:144-205,
 *                                   MissileLauncher.py:82-138, CCP.py:368-431,
 *                                   Missile.py:148-196, AirObject.py:39-42)
 * zrk_sync                         (run_simulation returning)       Manager.py:111
 * zrk_get_time                     Timer.get_time                   Timer.py:20-22
 * zrk_read_messages                Manager.give_messages            Manager.py:60-70
 * zrk_read_counts, zrk_read_totals sum(len(messages))               main.py:171
 * zrk_read_activity                sum(len(ActiveObjectsMessage.active_objects))  AirEnv.py:50-53
 * zrk_read_certified               (diagnostic: ticks on which CCP.link_object's scan, CCP.py:171-219, was proven away)
 * zrk_read_state                   model attributes (pos, prev_pos, current_azimuth,
 *                                  _target_dict ...)                BaseModel.py:28-36, CCP.py:80-86
 * zrk_read_status                  exception propagation            (CCP.py:259, CCP.py:107)
 * zrk_destroy, zrk_release_cached_memory   Python GC
 */
#ifndef ZRK_B200_H
#define ZRK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZRK_ABI_VERSION 1

/* Python `None` for id-valued fields (receiver_id, target_id, missile_id). */
#define ZRK_NONE INT32_MIN

/* modules/constants.py:36-38 */
#define ZRK_CCP_ID 0
#define ZRK_DRAWER_ID 1
#define ZRK_MANAGER_ID 2
/* modules/constants.py:31 POSSIBLE_TARGET_RADIUS */
#define ZRK_POSSIBLE_TARGET_RADIUS 100

/* MessageType, in the declaration order of modules/constants.py:3-21. */
enum zrk_msg_type {
    ZRK_MSG_LAUNCH_MISSILE = 0,
    ZRK_MSG_LAUNCHED_MISSILE = 1,
    ZRK_MSG_LAUNCH_COMMAND = 2,
    ZRK_MSG_MISSILE_COUNT_REQUEST = 3,
    ZRK_MSG_MISSILE_COUNT_RESPONSE = 4,
    ZRK_MSG_ALL_OBJECTS = 5,
    ZRK_MSG_FOUND_OBJECTS = 6,
    ZRK_MSG_ACTIVE_OBJECTS = 7,
    ZRK_MSG_CCP_UPDATE_TARGET = 8,
    ZRK_MSG_UPDATE_TARGET = 9,      /* never emitted */
    ZRK_MSG_MISSILE_GET_HIT = 10,   /* no class, no emitter */
    ZRK_MSG_DESTROYED_MISSILE = 11,
    ZRK_MSG_DRAW_OBJECTS = 12,
    ZRK_MSG_MISSILE_POS = 13,
    ZRK_MSG_MISSILE_DETONATE = 14,
    ZRK_MSG_NEW_MISSILE = 15,
    ZRK_MSG_LAUNCH_CANCELLED = 16,
    ZRK_MSG_LAUNCH_SUCCESSFUL = 17,
    ZRK_MSG_NTYPES = 18
};

/* TargetType (modules/utils.py:6-9) + the 'ЗУР' string CCP uses for missiles (CCP.py:247,400). */
enum zrk_obj_type { ZRK_AIR_PLANE = 0, ZRK_HELICOPTER = 1, ZRK_ANOTHER = 2, ZRK_ZUR = 3 };

/* scan_mode (Radar.py:26); anything else never moves the sector (Radar.py:100-117). */
enum zrk_scan_mode { ZRK_SCAN_HORIZONTAL = 0, ZRK_SCAN_VERTICAL = 1, ZRK_SCAN_OTHER = 2 };

/* Reason strings of the ValueErrors in Missile._calculate_trajectory_params (Missile.py:72-94). */
enum zrk_cancel_reason {
    ZRK_CANCEL_PARALLEL = 0,      /* Missile.py:74  */
    ZRK_CANCEL_PAST = 1,          /* Missile.py:79  */
    ZRK_CANCEL_NO_REAL_ROOT = 2,  /* Missile.py:83  */
    ZRK_CANCEL_NOT_POSITIVE = 3,  /* Missile.py:90  */
    ZRK_CANCEL_TOO_FAR = 4        /* Missile.py:94  */
};

/* Per-scenario status: the reference aborts the whole run on these exceptions; the batched
 * tick freezes just that scenario at the faulting point and reports it. */
enum zrk_fault {
    ZRK_FAULT_NONE = 0,
    ZRK_FAULT_ATTRIBUTE_ERROR_CCP259 = 1, /* 'Missile' object has no attribute 'type' (CCP.py:259) */
    ZRK_FAULT_KEY_ERROR_CCP107 = 2,       /* _target_dict[missile.target.id] (CCP.py:107)          */
    ZRK_FAULT_LOG_OVERFLOW = 64           /* not a reference fault: this scenario's log region is full */
};

/* Return codes (no exceptions cross the ABI). */
enum zrk_rc {
    ZRK_OK = 0,
    ZRK_ERR_ARG = -1,
    ZRK_ERR_CUDA = -2,
    ZRK_ERR_NO_DEVICE = -3,
    ZRK_ERR_NOMEM = -4,
    ZRK_ERR_UNSUPPORTED = -5,
    ZRK_ERR_CAPACITY = -6
};

/* Which message types are written as explicit records to the per-scenario log.
 * ACTIVE/ALL/FOUND carry object lists and are never stored as records: they are rebuilt
 * from the per-tick alive / visibility masks (ZRK_LOG_MASKS). */
#define ZRK_TYPE_BIT(t) (1u << (t))
#define ZRK_LOG_EVENTS                                                                          \
    (ZRK_TYPE_BIT(ZRK_MSG_LAUNCH_MISSILE) | ZRK_TYPE_BIT(ZRK_MSG_LAUNCHED_MISSILE) |              \
     ZRK_TYPE_BIT(ZRK_MSG_LAUNCH_COMMAND) | ZRK_TYPE_BIT(ZRK_MSG_MISSILE_COUNT_REQUEST) |         \
     ZRK_TYPE_BIT(ZRK_MSG_MISSILE_COUNT_RESPONSE) | ZRK_TYPE_BIT(ZRK_MSG_DESTROYED_MISSILE) |     \
     ZRK_TYPE_BIT(ZRK_MSG_MISSILE_DETONATE) | ZRK_TYPE_BIT(ZRK_MSG_NEW_MISSILE) |                 \
     ZRK_TYPE_BIT(ZRK_MSG_LAUNCH_CANCELLED) | ZRK_TYPE_BIT(ZRK_MSG_LAUNCH_SUCCESSFUL))
#define ZRK_LOG_BULK                                                                            \
    (ZRK_TYPE_BIT(ZRK_MSG_CCP_UPDATE_TARGET) | ZRK_TYPE_BIT(ZRK_MSG_DRAW_OBJECTS) |               \
     ZRK_TYPE_BIT(ZRK_MSG_MISSILE_POS))
#define ZRK_LOG_MASKS 0x80000000u /* keep per-tick alive + per-radar visibility masks */
/* Private record type 30, the "track journal": one record per association of a detection with a track other
 * than the one that followed its slot (flag = 0 new track / 1 _target_dict / 2 _missile_dict, sender = AirEnv
 * slot, receiver = entry index, a = object id, b = launch commanded).  Switched on together with ZRK_LOG_MASKS;
 * zrk_read_messages replays it to rebuild CCP_UPDATE_TARGET / DRAW_OBJECTS / MISSILE_POS when those are not
 * stored as records; zrk_read_records never returns it. */
#define ZRK_REC_TRACK 30
#define ZRK_LOG_TRACKS 0x40000000u
#define ZRK_LOG_FULL (ZRK_LOG_EVENTS | ZRK_LOG_BULK | ZRK_LOG_MASKS)

/*
 * One message of the reference's log (modules/BaseMessage.py:12-17 + modules/Messages.py),
 * flattened.  `tick` is send_time / time_step.  relevance is 3 for MISSILE_COUNT_REQUEST
 * (Messages.py:63) and 1 otherwise.  Payload by type:
 *   MISSILE_POS            a=missile_id
 *   MISSILE_DETONATE       a=missile_id  b=target_id|NONE        flag=self_detonation
 *   ACTIVE/ALL/FOUND       a=offset into the id pool, b=number of ids
 *   DESTROYED_MISSILE      a=missile_id (the reference stores the 1-tuple (id,), Messages.py:160)  flag=self_detonation
 *   LAUNCH_MISSILE         a=target.id
 *   LAUNCH_SUCCESSFUL      a=missile.id  b=target_id  xyz=new trajectory velocity (launch_time = tick*dt/1000)
 *   LAUNCH_CANCELLED       a=missile.id  b=zrk_cancel_reason
 *   LAUNCHED_MISSILE       a=missile.id  b=target_id
 *   NEW_MISSILE            a=missile.id
 *   MISSILE_COUNT_RESPONSE b=count
 *   LAUNCH_COMMAND         a=target.id   b=radar_id  xyz=target_position
 *   CCP_UPDATE_TARGET      a=target.id   b=missile_id|NONE
 *   DRAW_OBJECTS           a=obj_id      b=zrk_obj_type  flag=is_visible_by_radar  xyz=coordinates
 */
typedef struct zrk_msg {
    int32_t tick;
    uint8_t type;      /* zrk_msg_type */
    uint8_t relevance;
    uint8_t flag;
    uint8_t pad;
    int32_t sender;
    int32_t receiver;  /* ZRK_NONE for None */
    int32_t a;
    int32_t b;
    double xyz[3];
} zrk_msg; /* 48 bytes */

/*
 * A batch of B scenarios in the schema of main.create_objects_from_config (main.py:35-149),
 * as host structure-of-arrays padded to the batch maxima.  All pointers are host memory owned
 * by the caller; zrk_create copies what it needs.
 *
 * Object identity: "static index" s = targets in config order (0..n_targets-1) followed by
 * missiles in launcher-major config order (max_targets + m).  AirEnv "slots" are positions in
 * AirEnv.__objects (AirEnv.py:24): targets first, missiles in NEW_MISSILE arrival order.
 * Restrictions checked by zrk_create (ZRK_ERR_UNSUPPORTED): air-object ids unique within a
 * scenario; launcher ids unique within a scenario.
 */
typedef struct zrk_batch_desc {
    int32_t abi_version; /* ZRK_ABI_VERSION */
    int32_t n_scenarios;
    int32_t max_targets;
    int32_t max_radars;
    int32_t max_launchers;
    int32_t max_missiles;       /* missile objects per scenario, all launchers together */
    int32_t max_ccp_launchers;  /* entries of CCP.missile_launcher_coords (main.py:107-111) */
    int32_t dt_ms;              /* simulation.time_step (main.py:42) */

    /* per scenario [B]; NULL means "the max for every scenario" */
    const int32_t *n_targets, *n_radars, *n_launchers, *n_missiles, *n_ccp_launchers;
    const int32_t *airenv_id;  /* [B] air_environment.id */
    const int32_t *ccp_id;     /* [B] combat_control_point.id, ZRK_NONE = no CCP module (main.py:104-105) */
    const int32_t *tick_limit; /* [B] optional: scenario stops after this many ticks; NULL = never */

    /* targets [B][max_targets] (main.py:128-147) */
    const int32_t *target_id;
    const int32_t *target_type;  /* zrk_obj_type */
    const double *target_pos;    /* [..][3] */
    const double *target_vel;    /* [..][3] */

    /* radars [B][max_radars] (main.py:56-71) */
    const int32_t *radar_id;
    const double *radar_pos;     /* [..][3] */
    const double *radar_par;     /* [..][8] azimuth_start, elevation_start, max_distance, azimuth_range,
                                    elevation_range, azimuth_speed, elevation_speed, zrk_scan_mode */

    /* launchers [B][max_launchers] (main.py:74-101) */
    const int32_t *launcher_id;
    const double *launcher_pos;  /* [..][3] */

    /* missiles [B][max_missiles], launcher-major, only those MissileLauncher.add_missile accepted
       (MissileLauncher.py:43-48) */
    const int32_t *missile_id;
    const int32_t *missile_launcher; /* index into this scenario's launchers */
    const double *missile_v0;        /* velocity_module */
    const double *missile_v0sq;      /* Python `v0 ** 2` (Missile.py:65); NULL = v0*v0 */
    const double *missile_radius;    /* detonate_radius */
    const double *missile_life;      /* detonate_period, seconds */

    /* CCP launcher dict, in dict order [B][max_ccp_launchers] (main.py:107-111) */
    const int32_t *ccp_launcher_id;
    const double *ccp_launcher_pos;  /* [..][3] */

    /* Radar measurement noise replacing np.random.normal in SectorRadar.smooth_objects
       (Radar.py:138-142): noise[(tick % noise_period)][radar index][static index][3], added in
       place to the object's position for every (radar, visible object).  NULL = no noise.
       noise_scenario_stride (doubles) = 0 shares one table between all scenarios. */
    const double *noise;
    int32_t noise_period;
    int32_t noise_radars;
    int32_t noise_objects;
    int64_t noise_scenario_stride;
} zrk_batch_desc;

typedef struct zrk_options {
    int32_t device;            /* CUDA device ordinal */
    uint32_t log_types;        /* ZRK_LOG_* */
    int32_t log_capacity;      /* records per scenario; 0 = bound for ZRK_LOG_EVENTS computed from the batch */
    int32_t mask_ticks;        /* ticks of mask history to keep when ZRK_LOG_MASKS (ring if smaller than the run) */
    int32_t ticks_per_launch;  /* ticks fused into one kernel launch; 0 = library default */
    uint32_t flags;            /* ZRK_OPT_* */
    int32_t waves;             /* > 1: split the batch into that many contiguous scenario ranges ("waves"), each uploaded
                                  and initialised on its own CUDA stream, where the first kernels of the first zrk_step
                                  are queued as well, so that the host-to-device copy of a wave overlaps the kernels of
                                  the waves before it (later launches run over the whole batch).  zrk_create then returns without
                                  waiting: the arrays of the zrk_batch_desc must stay valid and unchanged until the first
                                  call that synchronises (zrk_sync or any zrk_read_*), and a per-scenario validation
                                  error is reported by that call (and by every call after it) instead of zrk_create.
                                  Page-locked host arrays are needed for the copies to overlap.  0 or 1 = off. */
    int32_t wave_rounds;       /* with waves > 1: how many rounds of the first zrk_step are queued wave by wave (1..4; 0 = 1).
                                  One round is ~1/5 of a c5 step's kernel time, enough to hide the uploads of one GPU; when
                                  several GPUs share the host's PCIe fabric the uploads take longer, and two rounds hide
                                  them at the price of smaller launches (c5, 8 GPUs: 35.5 -> 34.5 ms; 1 GPU: 30.9 -> 32.4 ms). */
} zrk_options;

#define ZRK_OPT_EXACT_ONLY 1u  /* run the tick without the fp32 / squared-distance pre-filters (same results, slower) */
#define ZRK_OPT_NO_CERTIFICATE 2u /* always run CCP.link_object's scan, even on ticks whose outcome the association
                                     certificate proves in advance (same results, slower; for A/B measurements) */

/* Caller-provided host buffers for zrk_read_state; any pointer may be NULL (skipped).
 * NS = max_targets + max_missiles slots / static indices, E = NS track entries per dict. */
typedef struct zrk_state_view {
    /* AirEnv slots [B][NS] */
    int32_t *slot_obj;    /* static index of the object in this slot, -1 = never filled */
    uint8_t *slot_alive;  /* 1 = in AirEnv.__objects (not None) */
    /* objects by static index [B][NS] */
    double *pos;          /* [..][3] BaseModel.pos */
    double *prev_pos;     /* [..][3] AirObject.prev_pos */
    uint8_t *has_prev;    /* 0 = prev_pos is None */
    /* radars [B][max_radars] */
    double *radar_az, *radar_el;
    /* missiles [B][max_missiles] */
    uint8_t *missile_status;   /* 0 ready, 1 active, 2 detonated */
    double *missile_period;    /* remaining detonate_period */
    double *missile_vel;       /* [..][3] trajectory.velocity */
    double *missile_t0;        /* trajectory.start_time, seconds */
    int32_t *missile_target;   /* static index of Missile.target, -1 = None */
    /* launchers [B][max_launchers] */
    int32_t *launcher_count;   /* len(missiles) */
    /* CCP [B] and [B][E] */
    int32_t *ccp_n_targets, *ccp_n_missiles;
    int32_t *ccp_t_key, *ccp_t_ref, *ccp_t_upd, *ccp_t_missile; /* ref = static index; upd = tick */
    uint8_t *ccp_t_following;
    int32_t *ccp_m_key, *ccp_m_ref, *ccp_m_upd, *ccp_m_target;
    int32_t *ccp_capacity, *ccp_launched;  /* [B][max_ccp_launchers] */
} zrk_state_view;

typedef struct zrk_handle zrk_handle;

/* Build the device state for a batch.  Returns ZRK_ERR_NO_DEVICE when no CUDA device is
 * usable: there is no CPU fallback. */
int zrk_create(const zrk_batch_desc *desc, const zrk_options *opt, zrk_handle **out);
/* Advance every scenario n_ticks ticks.  Asynchronous until zrk_sync or a read. */
int zrk_step(zrk_handle *h, int n_ticks);
int zrk_sync(zrk_handle *h);
/* Current tick index and Timer.get_time() = tick * dt_ms. */
int zrk_get_time(zrk_handle *h, int64_t *tick, int64_t *time_ms);
/* Per scenario: fault code (zrk_fault), tick of the fault, explicit records logged. */
int zrk_read_status(zrk_handle *h, int32_t *fault_code, int32_t *fault_tick, int32_t *n_records);
/* Messages emitted so far per scenario and type: counts[B][ZRK_MSG_NTYPES]. */
int zrk_read_counts(zrk_handle *h, int64_t *counts);
/* Batch totals, reduced on the device (a read-back of ZRK_MSG_NTYPES + 4 words instead of B x ZRK_MSG_NTYPES):
 * totals[0 .. ZRK_MSG_NTYPES) messages per type over all scenarios (main.py:171 summed over the batch), then the
 * number of scenarios that faulted, with ZRK_FAULT_ATTRIBUTE_ERROR_CCP259, with ZRK_FAULT_KEY_ERROR_CCP107, in all. */
int zrk_read_totals(zrk_handle *h, int64_t *totals);
/* Per scenario: alive air objects summed over the ticks started so far, i.e. the total length of
 * the active_objects lists AirEnv.step published (AirEnv.py:50-53); divided by slots x ticks it is
 * the active-slot fraction of the run.  Counted modulo 2^32. */
int zrk_read_activity(zrk_handle *h, int64_t *alive_slot_ticks);
/* Per scenario: ticks whose CombatControlPoint.step (CCP.py:368-431) took the certified path, i.e.
 * skipped link_object's scan because its outcome was proven in advance.  A diagnostic: results do not
 * depend on it.  Counted modulo 2^32. */
int zrk_read_certified(zrk_handle *h, int64_t *certified_ticks);
/* Copy the explicit records of scenarios [s0, s1) to host: recs[(s - s0) * cap + i] for
 * i < n[s - s0], cap = zrk_log_capacity().  Records of one scenario are in emission order. */
int zrk_read_records(zrk_handle *h, int s0, int s1, zrk_msg *recs, int32_t *n);
int zrk_log_capacity(zrk_handle *h);
/* The full message list of one scenario for ticks [tick0, tick1), in the reference's order
 * (Manager.py:41-58).  Needs ZRK_LOG_MASKS and a mask history (zrk_options.mask_ticks) covering the run so far.
 * ACTIVE/ALL/FOUND are always rebuilt from the masks; with ZRK_LOG_FULL the other bulk types come from their
 * records, with ZRK_LOG_EVENTS | ZRK_LOG_MASKS (the mode that scales: 8 + 4 R bytes per scenario-tick) MISSILE_POS,
 * CCP_UPDATE_TARGET and DRAW_OBJECTS are rebuilt on the host from the masks, the closed-form trajectories
 * (AirObject.py:23-25), the noise table and a replay of the event records + track journal — the same messages, bit for bit.
 * Returns ZRK_ERR_CAPACITY (and the needed sizes) if a buffer is too small. */
int zrk_read_messages(zrk_handle *h, int scenario, int tick0, int tick1, zrk_msg *msgs, size_t cap,
                      size_t *n, int32_t *ids, size_t ids_cap, size_t *n_ids);
int zrk_read_state(zrk_handle *h, const zrk_state_view *view);
/* Device pointers for zero-copy consumers (torch): borrowed, valid until zrk_destroy. */
int zrk_device_ptr(zrk_handle *h, const char *name, void **ptr, size_t *bytes); /* "dyn", "cst", "rec", "masks", "summary" */
/* Checkpoint / resume (absent in the reference): the complete device state of the
 * batch (scenario state, record log, mask history, clock) as one opaque host blob.
 * zrk_state_bytes gives the size; zrk_load_state needs a handle created from the same
 * zrk_batch_desc and zrk_options. */
int zrk_state_bytes(zrk_handle *h, size_t *bytes);
int zrk_save_state(zrk_handle *h, void *buf, size_t cap);
int zrk_load_state(zrk_handle *h, const void *buf, size_t bytes);
/* Kernel launches issued by this handle and device time spent in them (CUDA events), ms.
 * With zrk_options.waves > 1 the launches of different waves overlap: the figure is then the sum
 * of the per-launch durations, not the elapsed time. */
int zrk_launch_stats(zrk_handle *h, int64_t *n_launches, double *kernel_ms);
/* Device memory of destroyed handles is kept in a memory pool of the library's own for the next handle (the
 * device's default pool, the host application's, is never touched); this hands it back to the driver. */
int zrk_release_cached_memory(void);
const char *zrk_last_error(zrk_handle *h); /* h may be NULL: error of the last failed zrk_create */
void zrk_destroy(zrk_handle *h);
int zrk_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZRK_B200_H */
