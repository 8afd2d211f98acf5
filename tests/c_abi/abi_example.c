/* abi_example.c — the C ABI used from plain C (no Python, no torch): build the shipped
 * simulation_config.yaml scenario by hand, run its 200 ticks, print the per-type message counts.
 *   gcc -I include tests/c_abi/abi_example.c -o abi_example -ldl && ./abi_example path/to/libzrk_b200.so
 * tests/test_gpu_parity.py::test_c_abi_from_plain_c runs it on the GPU box and compares the output
 * with the reference's known answers (685 messages, detonations at ticks 11 and 13). */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zrk_b200.h"

#define LOAD(name) name##_t name##_f = (name##_t)dlsym(lib, #name); if (!name##_f) { fprintf(stderr, "missing %s\n", #name); return 2; }
typedef int (*zrk_create_t)(const zrk_batch_desc *, const zrk_options *, zrk_handle **);
typedef int (*zrk_step_t)(zrk_handle *, int);
typedef int (*zrk_read_counts_t)(zrk_handle *, int64_t *);
typedef int (*zrk_read_records_t)(zrk_handle *, int, int, zrk_msg *, int32_t *);
typedef int (*zrk_log_capacity_t)(zrk_handle *);
typedef const char *(*zrk_last_error_t)(zrk_handle *);
typedef void (*zrk_destroy_t)(zrk_handle *);

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    LOAD(zrk_create) LOAD(zrk_step) LOAD(zrk_read_counts) LOAD(zrk_read_records) LOAD(zrk_log_capacity)
    LOAD(zrk_last_error) LOAD(zrk_destroy)

    /* simulation_config.yaml: 2 targets, 1 radar, 2 launchers x 5 missiles, CCP 0 */
    int32_t target_id[2] = {39, 40}, target_type[2] = {ZRK_AIR_PLANE, ZRK_AIR_PLANE};
    double target_pos[6] = {-57, 1152, 1000, 919, 505, 1000}, target_vel[6] = {100, 0, 0, 100, 0, 0};
    int32_t radar_id[1] = {5};
    double radar_pos[3] = {2, 1, 0};
    double radar_par[8] = {0, 0, 20000, 180, 90, 180, 0, ZRK_SCAN_HORIZONTAL};
    int32_t launcher_id[2] = {3, 4};
    double launcher_pos[6] = {2238, -17, 0, -1071, 0, 0};
    int32_t missile_id[10], missile_launcher[10];
    double v0[10], v0sq[10], radius[10], life[10];
    for (int m = 0; m < 10; m++) {
        missile_launcher[m] = m / 5;
        missile_id[m] = (m < 5 ? 3001 : 4001) + m % 5;
        v0[m] = 1000; v0sq[m] = 1e6; radius[m] = 150; life[m] = 60;
    }
    int32_t airenv_id[1] = {999}, ccp_id[1] = {0};

    zrk_batch_desc d;
    memset(&d, 0, sizeof d);
    d.abi_version = ZRK_ABI_VERSION;
    d.n_scenarios = 1; d.max_targets = 2; d.max_radars = 1; d.max_launchers = 2; d.max_missiles = 10;
    d.max_ccp_launchers = 2; d.dt_ms = 200;
    d.airenv_id = airenv_id; d.ccp_id = ccp_id;
    d.target_id = target_id; d.target_type = target_type; d.target_pos = target_pos; d.target_vel = target_vel;
    d.radar_id = radar_id; d.radar_pos = radar_pos; d.radar_par = radar_par;
    d.launcher_id = launcher_id; d.launcher_pos = launcher_pos;
    d.missile_id = missile_id; d.missile_launcher = missile_launcher; d.missile_v0 = v0; d.missile_v0sq = v0sq;
    d.missile_radius = radius; d.missile_life = life;
    d.ccp_launcher_id = launcher_id; d.ccp_launcher_pos = launcher_pos;

    zrk_handle *h = NULL;
    int rc = zrk_create_f(&d, NULL, &h);
    if (rc != ZRK_OK) { fprintf(stderr, "zrk_create: %d %s\n", rc, zrk_last_error_f(NULL)); return 1; }
    if ((rc = zrk_step_f(h, 200)) != ZRK_OK) { fprintf(stderr, "zrk_step: %d %s\n", rc, zrk_last_error_f(h)); return 1; }
    int64_t counts[ZRK_MSG_NTYPES];
    if ((rc = zrk_read_counts_f(h, counts)) != ZRK_OK) return 1;
    int64_t total = 0;
    for (int t = 0; t < ZRK_MSG_NTYPES; t++) { printf("%lld ", (long long)counts[t]); total += counts[t]; }
    printf("\ntotal %lld\n", (long long)total);
    int cap = zrk_log_capacity_f(h);
    zrk_msg *recs = malloc(sizeof(zrk_msg) * cap);
    int32_t n = 0;
    if ((rc = zrk_read_records_f(h, 0, 1, recs, &n)) != ZRK_OK) return 1;
    for (int i = 0; i < n; i++)
        if (recs[i].type == ZRK_MSG_MISSILE_DETONATE)
            printf("detonate tick %d missile %d target %d\n", recs[i].tick, recs[i].a, recs[i].b);
    free(recs);
    zrk_destroy_f(h);
    return 0;
}
