// zrk_host.cu — the C ABI of include/zrk_b200.h on top of the sm_100a tick kernel.
//
// zrk_create uploads the caller's structure-of-arrays scenario description as it is and builds
// the device state with an init kernel (no per-scenario host loop); zrk_step launches the tick
// kernel on the handle's stream; reads copy back and, for zrk_read_messages, merge the explicit
// record log with the ACTIVE/ALL/FOUND lists rebuilt from the per-tick masks into the
// reference's message order (Manager.py:41-58).
#ifdef ZRK_EMU
#include "cuda_runtime_shim.h"
#define ZRK_LAUNCH(kern, grid, block, smem, stream, ...) \
    zrk_emu::launch(dim3(grid), dim3(block), (smem), [&]() { kern(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define ZRK_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#if defined(__linux__)
#include <sched.h>
#endif
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "zrk_tick.cuh"

using namespace zrk;

// ------------------------------------------------------------------ init kernel

struct DevDesc {  // zrk_batch_desc with device pointers
    int B, T, R, L, M, C;
    const int *n_targets, *n_radars, *n_launchers, *n_missiles, *n_ccp_launchers, *airenv_id, *ccp_id, *tick_limit;
    const int *target_id, *target_type;
    const double *target_pos, *target_vel;
    const int *radar_id;
    const double *radar_pos, *radar_par;
    const int *launcher_id;
    const double *launcher_pos;
    const int *missile_id, *missile_launcher;
    const double *missile_v0, *missile_v0sq, *missile_radius, *missile_life;
    const int *ccp_launcher_id;
    const double *ccp_launcher_pos;
};

// One warp per scenario (lanes stride over the slots, so the blob is written coalesced):
// main.create_objects_from_config (main.py:35-149) into the device layout.
#define ZRK_INIT_BLOCK 128
__global__ void zrk_init_kernel(const Params P, const DevDesc d, int *err, int b0, int b1) {
    const Layout &o = P.lo;
    const int lane = threadIdx.x & 31;
    const int b = b0 + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
    if (b >= b1) return;  // whole warp
    const int NSP = o.team, MP = o.MP;
    const HotOff &ho = o.hot;
    unsigned char *sm = P.dyn + (size_t)b * o.dyn_bytes;
    unsigned char *cs = const_cast<unsigned char *>(P.cst) + (size_t)b * o.cst_bytes;
    const int nT = d.n_targets ? d.n_targets[b] : d.T, nR = d.n_radars ? d.n_radars[b] : d.R;
    const int nL = d.n_launchers ? d.n_launchers[b] : d.L, nM = d.n_missiles ? d.n_missiles[b] : d.M;
    const int nC = d.n_ccp_launchers ? d.n_ccp_launchers[b] : d.C;
    if (nT < 0 || nT > d.T || nR < 0 || nR > d.R || nL < 0 || nL > d.L || nM < 0 || nM > d.M || nC < 0 || nC > d.C) {
        if (lane == 0) atomicOr(err, 1);
        return;
    }
    int *hdr = (int *)(sm + ho.hdr);
    double2 *pxy = (double2 *)(sm + ho.pxy), *pzs = (double2 *)(sm + ho.pzs);
    double2 *txy = (double2 *)(sm + ho.txy), *tzt = (double2 *)(sm + ho.tzt), *vxy = (double2 *)(sm + ho.vxy),
            *qzw = (double2 *)(sm + ho.qzw);
    int *sid = (int *)(sm + ho.sid), *sflags = (int *)(sm + ho.sflags), *srf = (int *)(sm + ho.srf),
        *own = (int *)(sm + ho.own);
    for (int t = lane; t < nT; t += 32) {
        size_t g = (size_t)b * d.T + t;
        double vx = d.target_vel[3 * g], vy = d.target_vel[3 * g + 1], vz = d.target_vel[3 * g + 2];
        double sp = sqrt(dot3(vx, vy, vz, vx, vy, vz));  // np.linalg.norm (AirObject.py:36)
        txy[t] = make_double2(d.target_pos[3 * g], d.target_pos[3 * g + 1]);  // Trajectory (main.py:134-138)
        tzt[t] = make_double2(d.target_pos[3 * g + 2], 0.0);
        vxy[t] = make_double2(vx, vy);
        qzw[t] = make_double2(0.0, vz);
        pxy[t] = make_double2(d.target_pos[3 * g], d.target_pos[3 * g + 1]);
        pzs[t] = make_double2(d.target_pos[3 * g + 2], sp);
        sid[t] = d.target_id[g];
        int tt = d.target_type ? d.target_type[g] : ZRK_ANOTHER;
        sflags[t] = F_ALIVE | ((tt & 3) << F_TTYPE_SHIFT) | (t << F_STATIC_SHIFT);
    }
    int *te_mid = (int *)(sm + ho.te_mid);
    for (int k = lane; k < NSP; k += 32) {
        srf[k] = -1;
        own[k] = -1;
        te_mid[k] = ZRK_NONE;
    }
    if (lane == 0) {
        ((int *)(sm + ho.counts))[CW_FLAG] = CF_WINDOWS | CF_GENERAL;  // no certificate yet; tick 0 is a general tick
        hdr[H_NSLOTS] = nT;
        int *chdr = (int *)(cs + o.c_hdr);
        chdr[C_NT] = nT; chdr[C_NR] = nR; chdr[C_NL] = nL; chdr[C_NM] = nM; chdr[C_NC] = nC;
        chdr[C_AIRENV_ID] = d.airenv_id ? d.airenv_id[b] : 1;
        chdr[C_CCP_ID] = d.ccp_id ? d.ccp_id[b] : 0;
        chdr[C_TICK_LIMIT] = d.tick_limit ? d.tick_limit[b] : -1;
    }
    double *rad_cur = (double *)(sm + o.rad_cur);
    int *c_rad_id = (int *)(cs + o.c_rad_id);
    double *c_rad_pos = (double *)(cs + o.c_rad_pos), *c_rad_par = (double *)(cs + o.c_rad_par);
    for (int r = lane; r < nR; r += 32) {
        size_t g = (size_t)b * d.R + r;
        c_rad_id[r] = d.radar_id[g];
        for (int k = 0; k < 3; k++) c_rad_pos[3 * r + k] = d.radar_pos[3 * g + k];
        for (int k = 0; k < 8; k++) c_rad_par[RAD_PAR_STRIDE * r + k] = d.radar_par[8 * g + k];
        double md = d.radar_par[8 * g + 2], md2 = md * md;
        c_rad_par[RAD_PAR_STRIDE * r + 8] = md >= 0 ? md2 * (1.0 - 1e-14) : -1.0;
        c_rad_par[RAD_PAR_STRIDE * r + 9] = md >= 0 ? md2 * (1.0 + 1e-14) : -1.0;
        const double az0 = d.radar_par[8 * g + 0], el0 = d.radar_par[8 * g + 1];
        rad_cur[2 * r] = az0;
        rad_cur[2 * r + 1] = el0;
        // a start sector that the sweep maps onto itself stays there forever: skip the sweep
        double caz = az0, cel = el0;
        radar_move(c_rad_par + RAD_PAR_STRIDE * r, caz, cel);
        const bool fixed = caz == az0 && cel == el0 && !(caz == 0.0 && signbit(caz) != signbit(az0)) &&
                           !(cel == 0.0 && signbit(cel) != signbit(el0));
        c_rad_par[RAD_PAR_STRIDE * r + 10] = fixed ? 1.0 : 0.0;
        c_rad_par[RAD_PAR_STRIDE * r + 11] = 0.0;
    }
    int *c_l_id = (int *)(cs + o.c_l_id);
    double *c_l_pos = (double *)(cs + o.c_l_pos);
    for (int l = lane; l < nL; l += 32) {
        size_t g = (size_t)b * d.L + l;
        c_l_id[l] = d.launcher_id[g];
        for (int k = 0; k < 3; k++) c_l_pos[3 * l + k] = d.launcher_pos[3 * g + k];
    }
    int *st_n = (int *)(sm + o.st_n), *st = (int *)(sm + o.st);
    int *m_status = (int *)(sm + o.m_status), *m_target = (int *)(sm + o.m_target), *m_slot = (int *)(sm + o.m_slot),
        *m_birth = (int *)(sm + o.m_birth);
    double *m_period = (double *)(sm + o.m_period);
    int *c_m_id = (int *)(cs + o.c_m_id), *c_m_launcher = (int *)(cs + o.c_m_launcher);
    double *c_m_v0 = (double *)(cs + o.c_m_v0), *c_m_v0sq = (double *)(cs + o.c_m_v0sq),
           *c_m_radius = (double *)(cs + o.c_m_radius);
    bool bad = false;
    for (int m = lane; m < nM; m += 32) {
        size_t g = (size_t)b * d.M + m;
        int l = d.missile_launcher[g];
        if (l < 0 || l >= nL) {
            bad = true;
            break;
        }
        c_m_id[m] = d.missile_id[g];
        c_m_launcher[m] = l;
        c_m_v0[m] = d.missile_v0[g];
        c_m_v0sq[m] = d.missile_v0sq ? d.missile_v0sq[g] : d.missile_v0[g] * d.missile_v0[g];
        c_m_radius[m] = d.missile_radius[g];
        m_period[m] = d.missile_life[g];
        m_status[m] = ST_READY;
        m_target[m] = -1;
        m_slot[m] = -1;
        m_birth[m] = -1;
    }
    if (__ballot_sync(0xffffffffu, bad)) {
        if (lane == 0) atomicOr(err, 1);
        return;
    }
    __syncwarp();  // ids, launchers and missiles are written: the checks below read other lanes' entries
    if (lane == 0)  // MissileLauncher.add_missile (MissileLauncher.py:36-48): the stacks fill in list order
        for (int m = 0; m < nM; m++) {
            const int l = c_m_launcher[m];
            st[l * MP + st_n[l]++] = m;
        }
    for (int l = lane; l < nL; l += 32)
        for (int j = 0; j < l; j++)
            if (c_l_id[j] == c_l_id[l]) atomicOr(err, 2);
    // air-object ids must be unique (the tick resolves object identity by slot)
    for (int i = lane; i < nT + nM; i += 32) {
        int a = i < nT ? sid[i] : c_m_id[i - nT];
        for (int j = 0; j < i; j++) {
            int c = j < nT ? sid[j] : c_m_id[j - nT];
            if (a == c) atomicOr(err, 2);
        }
    }
    int *c_cl_id = (int *)(cs + o.c_cl_id), *c_cl_launcher = (int *)(cs + o.c_cl_launcher);
    double *c_cl_pos = (double *)(cs + o.c_cl_pos);
    for (int c = lane; c < nC; c += 32) {
        size_t g = (size_t)b * d.C + c;
        c_cl_id[c] = d.ccp_launcher_id[g];
        for (int k = 0; k < 3; k++) c_cl_pos[3 * c + k] = d.ccp_launcher_pos[3 * g + k];
        int li = -1;
        for (int l = 0; l < nL; l++)
            if (c_l_id[l] == c_cl_id[c]) li = l;
        c_cl_launcher[c] = li;
    }
}

// Per-scenario header + message counters packed [B][H_WORDS + NTYPES] for cheap read-back.
#define ZRK_SUMMARY_WORDS (H_WORDS + ZRK_MSG_NTYPES)
__global__ void zrk_summary_kernel(const Params P, int *out) {
    const Layout &o = P.lo;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)o.B * ZRK_SUMMARY_WORDS) return;
    const int b = (int)(i / ZRK_SUMMARY_WORDS), w = (int)(i % ZRK_SUMMARY_WORDS);
    const unsigned char *sm = P.dyn + (size_t)b * o.dyn_bytes;
    out[i] = w < H_WORDS ? ((const int *)(sm + o.hot.hdr))[w] : ((const int *)(sm + o.hot.counts))[w - H_WORDS];
}

// Batch totals for a caller that does not need per-scenario counters: message totals per type, then
// scenarios faulted / with AttributeError / with KeyError / in all (the statistics vector the ranks of a
// multi-GPU job all-reduce).  One thread per scenario, block partial sums in shared memory.
#define ZRK_TOTALS_WORDS (ZRK_MSG_NTYPES + 4)
__global__ void zrk_totals_kernel(const Params P, unsigned long long *out) {
    ZRK_SMEM_DECL;
    unsigned long long *part = (unsigned long long *)zrk_smem;  // [ZRK_TOTALS_WORDS]
    const Layout &o = P.lo;
    for (int i = threadIdx.x; i < ZRK_TOTALS_WORDS; i += blockDim.x) part[i] = 0;
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < o.B) {
        const unsigned char *sm = P.dyn + (size_t)b * o.dyn_bytes;
        const unsigned *cnt = (const unsigned *)(sm + o.hot.counts);
        const int fault = ((const int *)(sm + o.hot.hdr))[H_FAULT] & 63;
        for (int t = 0; t < ZRK_MSG_NTYPES; t++)
            if (cnt[t]) atomicAdd(&part[t], (unsigned long long)cnt[t]);
        if (fault) atomicAdd(&part[ZRK_MSG_NTYPES], 1ull);
        if (fault == ZRK_FAULT_ATTRIBUTE_ERROR_CCP259) atomicAdd(&part[ZRK_MSG_NTYPES + 1], 1ull);
        if (fault == ZRK_FAULT_KEY_ERROR_CCP107) atomicAdd(&part[ZRK_MSG_NTYPES + 2], 1ull);
        atomicAdd(&part[ZRK_MSG_NTYPES + 3], 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ZRK_TOTALS_WORDS; i += blockDim.x)
        if (part[i]) atomicAdd(&out[i], part[i]);
}

// ------------------------------------------------------------------ handle

struct zrk_handle {
    Params P{};
    zrk_batch_desc hd{};  // scalar fields only are meaningful after create
    int device = 0;
    cudaStream_t stream{};
    bool has_stream = false;  // zrk_create got as far as creating the stream
    std::vector<void *> allocs;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;  // per-launch event pairs
    std::vector<cudaEvent_t> free_events;
    long long n_launches = 0;
    double kernel_ms = 0;
    int ticks_per_launch = 0;
    long long tick = 0;  // ticks requested so far
    bool fast = true;
    std::string err;
    size_t smem_bytes = 0;
    int grid = 0, block = 0;
    // host copies kept for reconstruction
    std::vector<int> h_radar_id, h_n_radars, h_airenv_id;
    unsigned long long *d_totals = nullptr, *h_totals = nullptr;  // zrk_read_totals
    int *d_summary = nullptr;
    int *h_summary = nullptr;  // page-locked [B][ZRK_SUMMARY_WORDS], allocated while the first step runs
    double *d_tsec = nullptr;
    Params *d_self = nullptr;
    unsigned long long cst_hash = 0;  // identity of the scenario parameters (checked by zrk_load_state)
    // Scenario waves (zrk_options.waves > 1): contiguous scenario ranges, each uploaded,
    // initialised and stepped on its own stream, so that the host-to-device copy of one wave
    // overlaps the kernels of the waves before it.
    std::vector<cudaStream_t> wave_stream;  // empty = everything on `stream`
    std::vector<int> wave_b;                // wave w covers scenarios [wave_b[w], wave_b[w + 1])
    std::vector<cudaEvent_t> wave_done;
    bool defer_zero = false;                         // zrk_create: dev_alloc(zero) only notes the buffer ...
    std::vector<std::pair<void *, size_t>> zero_list;  // ... and the memsets are queued after the last host synchronisation
    cudaEvent_t alloc_ev = nullptr;
    int wave_rounds = 1;   // zrk_options.wave_rounds
    void *bounce = nullptr;  // page-locked staging of a pageable description (zrk_create), kept until the handle goes
    size_t bounce_bytes = 0;
    bool stepped = false;  // zrk_step has run (the waves' streams matter to the first step only)
    cudaEvent_t fork_ev{};
    int *pin_err = nullptr;   // pinned: the init kernel's verdict when zrk_create did not wait for it
    bool deferred = false;    // that verdict has not been looked at yet
    int sticky_rc = ZRK_OK;   // a deferred create error: every later call returns it
    // two-kernel stepping (warp teams with the certificate): rounds of streak kernel + tick kernel
    bool two_kernel = false;
    int *d_queue = nullptr;    // [ZRK_ROUND_SLOTS + 64] the work-queue counters of the streak kernel's launches (one per round; + one per wave)
    int n_sms = 0;
    int *d_pending = nullptr;  // [waves][ZRK_ROUND_SLOTS] scenarios still short of the target after a round
    int *h_pending = nullptr;  // pinned [waves]
    int *d_work = nullptr;     // [2][B] worklists: the scenarios a batch of rounds found pending (per wave: its own range)
    int round_no = 0;
    long long n_rounds = 0;
};
#define ZRK_ROUND_SLOTS 1024
#define ZRK_FORK_ROUNDS_MAX 4  /* zrk_options.wave_rounds: rounds of a handle's first step that can be issued wave by wave */
#define ZRK_ROUND_BATCH 6  /* rounds queued between two looks at the pending counters */

static thread_local std::string g_create_err;

static unsigned long long fnv1a(const void *p, size_t n, unsigned long long h = 1469598103934665603ull) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) h = (h ^ c[i]) * 1099511628211ull;
    return h;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return ZRK_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

// Page-locked host buffers are expensive to create (the pages are pinned one by one, under a
// lock that the ranks of a multi-GPU job share), so freed ones are kept for the next handle of
// the process: a handle created per step, as bench.py's end-to-end leg does, pins nothing after
// the first one.  At most ZRK_PINNED_KEEP blocks are kept.
#include <mutex>
#define ZRK_PINNED_KEEP 8
struct PinnedBlock {
    void *p;
    size_t bytes;
};
static std::mutex g_pinned_mu;
static std::vector<PinnedBlock> g_pinned_free;
static size_t pinned_class(size_t bytes) {  // powers of two up to 1 MiB, whole MiB above
    if (bytes > (1u << 20)) return (bytes + (1u << 20) - 1) >> 20 << 20;
    size_t c = 64;
    while (c < bytes) c <<= 1;
    return c;
}
static cudaError_t pinned_get(void **p, size_t bytes) {
    const size_t c = pinned_class(bytes);
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        for (size_t i = 0; i < g_pinned_free.size(); i++)
            if (g_pinned_free[i].bytes == c) {
                *p = g_pinned_free[i].p;
                g_pinned_free.erase(g_pinned_free.begin() + i);
                return cudaSuccess;
            }
    }
    return cudaHostAlloc(p, c, cudaHostAllocDefault);
}
static void pinned_put(void *p, size_t bytes) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        if (g_pinned_free.size() < ZRK_PINNED_KEEP) {
            g_pinned_free.push_back({p, pinned_class(bytes)});
            return;
        }
    }
    cudaFreeHost(p);
}

// Streams and events are kept for the next handle of the process as well (a handle created per step creates nine
// streams and some forty events otherwise): idle ones, per device; a stream is only put back once it has drained.
#define ZRK_STREAMS_KEEP 80
#define ZRK_EVENTS_KEEP 512
#define ZRK_CACHE_DEVICES 64
static std::vector<cudaStream_t> g_streams_free[ZRK_CACHE_DEVICES];
static std::vector<cudaEvent_t> g_events_free[ZRK_CACHE_DEVICES][2];  // [device][timing?]
static cudaError_t stream_get(int dev, cudaStream_t *st) {
    if (dev >= 0 && dev < ZRK_CACHE_DEVICES) {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        auto &v = g_streams_free[dev];
        if (!v.empty()) {
            *st = v.back();
            v.pop_back();
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
}
static void stream_put(int dev, cudaStream_t st) {  // (the caller has synchronised it)
    if (dev >= 0 && dev < ZRK_CACHE_DEVICES) {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        if (g_streams_free[dev].size() < ZRK_STREAMS_KEEP) {
            g_streams_free[dev].push_back(st);
            return;
        }
    }
    cudaStreamDestroy(st);
}
static cudaError_t event_get(int dev, bool timing, cudaEvent_t *e) {
    if (dev >= 0 && dev < ZRK_CACHE_DEVICES) {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        auto &v = g_events_free[dev][timing ? 1 : 0];
        if (!v.empty()) {
            *e = v.back();
            v.pop_back();
            return cudaSuccess;
        }
    }
    return timing ? cudaEventCreate(e) : cudaEventCreateWithFlags(e, cudaEventDisableTiming);
}
static void event_put(int dev, bool timing, cudaEvent_t e) {  // (whatever it was recorded on has completed)
    if (!e) return;
    if (dev >= 0 && dev < ZRK_CACHE_DEVICES) {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        auto &v = g_events_free[dev][timing ? 1 : 0];
        if (v.size() < ZRK_EVENTS_KEEP) {
            v.push_back(e);
            return;
        }
    }
    cudaEventDestroy(e);
}

// Stream-ordered allocations from a memory pool of the library's own (one per device, created on first
// use), which keeps freed memory: creating and destroying handles in a loop then costs no cudaMalloc /
// cudaFree, and the device's default pool — the host application's — is left as it was.
// zrk_release_cached_memory() hands the kept memory back.
#define ZRK_MAX_DEVICES 64
static std::mutex g_pool_mu;
static cudaMemPool_t g_pool[ZRK_MAX_DEVICES];
static bool g_pool_ok[ZRK_MAX_DEVICES];
static int pool_alloc(zrk_handle *h, void **p, size_t bytes) {
    if (h->device < 0 || h->device >= ZRK_MAX_DEVICES) return ZRK_ERR_ARG;
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        if (!g_pool_ok[h->device]) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = h->device;
            CK(cudaMemPoolCreate(&g_pool[h->device], &props));
            unsigned long long keep = ~0ull;
            CK(cudaMemPoolSetAttribute(g_pool[h->device], cudaMemPoolAttrReleaseThreshold, &keep));
            g_pool_ok[h->device] = true;
        }
    }
    CK(cudaMallocFromPoolAsync(p, bytes ? bytes : 1, g_pool[h->device], h->stream));
    h->allocs.push_back(*p);
    return ZRK_OK;
}
extern "C" int zrk_release_cached_memory(void) {
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        for (int d = 0; d < ZRK_CACHE_DEVICES; d++) {
            for (auto st : g_streams_free[d]) cudaStreamDestroy(st);
            g_streams_free[d].clear();
            for (int k = 0; k < 2; k++) {
                for (auto e : g_events_free[d][k]) cudaEventDestroy(e);
                g_events_free[d][k].clear();
            }
        }
    }
    std::lock_guard<std::mutex> g(g_pool_mu);
    for (int d = 0; d < ZRK_MAX_DEVICES; d++)
        if (g_pool_ok[d] && cudaMemPoolTrimTo(g_pool[d], 0) != cudaSuccess) return ZRK_ERR_CUDA;
    return ZRK_OK;
}

// Every entry point runs on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
static int dev_upload(zrk_handle *h, const T *src, size_t n, const T **dst) {
    *dst = nullptr;
    if (!src || !n) return ZRK_OK;
    void *p = nullptr;
    int rc = pool_alloc(h, &p, n * sizeof(T));
    if (rc) return rc;
    CK(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *dst = (const T *)p;
    return ZRK_OK;
}
template <typename T>
static int dev_alloc(zrk_handle *h, size_t n, T **dst, bool zero) {
    void *p = nullptr;
    int rc = pool_alloc(h, &p, (n ? n : 1) * sizeof(T));
    if (rc) return rc;
    if (zero) {
        if (h->defer_zero)
            h->zero_list.push_back({p, (n ? n : 1) * sizeof(T)});  // (zrk_create: queued once nothing has to wait for them)
        else
            CK(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), h->stream));
    }
    *dst = (T *)p;
    return ZRK_OK;
}

typedef void (*tick_fn)(const Params, int, int, int, int, int *, const int *, int *, const int *);
typedef void (*streak_fn)(const Params, int, int, int, const int *, const int *, int *);

template <int TEAM, int TPC, int MPB>
static tick_fn pick_fast(bool fast) {
    return fast ? (tick_fn)zrk_tick_kernel<TEAM, TPC, true, MPB> : (tick_fn)zrk_tick_kernel<TEAM, TPC, false, MPB>;
}
static tick_fn pick_kernel(const Layout &lo, bool fast) {
    switch (lo.team) {
    case 32: return lo.MP == 8 ? pick_fast<32, ZRK_CTA_THREADS / 32, 8>(fast) : pick_fast<32, ZRK_CTA_THREADS / 32, 32>(fast);
    case 64: return pick_fast<64, 1, 0>(fast);
    case 128: return pick_fast<128, 1, 0>(fast);
    case 256: return pick_fast<256, 1, 0>(fast);
    case 512: return pick_fast<512, 1, 0>(fast);
    case 1024: return pick_fast<1024, 1, 0>(fast);
    }
    return nullptr;
}

static streak_fn pick_streak(const Layout &lo) {
    return lo.MP == 8 ? (streak_fn)zrk_streak_kernel<8> : (streak_fn)zrk_streak_kernel<32>;
}

extern "C" int zrk_abi_version(void) { return ZRK_ABI_VERSION; }

extern "C" const char *zrk_last_error(zrk_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" void zrk_destroy(zrk_handle *h) {
    if (!h) return;
    if (!h->has_stream) {  // zrk_create failed before touching the device
        delete h;
        return;
    }
    DeviceGuard dev_guard(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto ws : h->wave_stream) cudaStreamSynchronize(ws);
    for (auto &pr : h->pending) {
        event_put(h->device, true, pr.first);
        event_put(h->device, true, pr.second);
    }
    for (auto e : h->free_events) event_put(h->device, true, e);
    for (auto ws : h->wave_stream) stream_put(h->device, ws);
    for (auto e : h->wave_done) event_put(h->device, false, e);
    if (!h->wave_stream.empty()) event_put(h->device, false, h->fork_ev);
    event_put(h->device, false, h->alloc_ev);
    pinned_put(h->pin_err, sizeof(int));
    pinned_put(h->bounce, h->bounce_bytes);
    pinned_put(h->h_pending, sizeof(int) * 64);
    pinned_put(h->h_totals, sizeof(unsigned long long) * ZRK_TOTALS_WORDS);
    pinned_put(h->h_summary, (size_t)h->P.lo.B * ZRK_SUMMARY_WORDS * sizeof(int));
    for (void *p : h->allocs) cudaFreeAsync(p, h->stream);
    cudaStreamSynchronize(h->stream);
    stream_put(h->device, h->stream);
    delete h;
}

// keep the device copy of Params (Params::self) current
static int push_params(zrk_handle *h) {
    if (!h->d_self) {
        void *q = nullptr;
        int rc = pool_alloc(h, &q, sizeof(Params));
        if (rc) return rc;
        h->d_self = (Params *)q;
        h->P.self = h->d_self;
    }
    CK(cudaMemcpyAsync(h->d_self, &h->P, sizeof(Params), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return ZRK_OK;
}

// tsec[t] = to_seconds(t * dt): the same int / 1000 the reference evaluates (AirObject.py:5-7)
static int ensure_tsec(zrk_handle *h, long long upto) {
    if (upto <= h->P.tsec_n) return ZRK_OK;
    long long n = h->P.tsec_n ? h->P.tsec_n : 1024;
    while (n < upto) n *= 2;
    if (n > (1 << 26)) return ZRK_OK;  // beyond the table the kernel divides
    std::vector<double> t((size_t)n);
    for (long long i = 0; i < n; i++) t[(size_t)i] = (double)(i * h->P.dt_ms) / 1000.0;
    void *q = nullptr;
    int rc = pool_alloc(h, &q, (size_t)n * sizeof(double));  // earlier, shorter tables stay alive until destroy
    if (rc) return rc;
    double *p = (double *)q;
    CK(cudaMemcpyAsync(p, t.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->P.tsec = p;
    h->P.tsec_n = (int)n;
    return push_params(h);
}

static int init_verdict(zrk_handle *h, int herr) {
    if (herr & 1) {
        h->err = "per-scenario count exceeds the batch maximum, or missile_launcher out of range";
        return ZRK_ERR_ARG;
    }
    if (herr & 2) {
        h->err = "unsupported scenario: air-object ids and launcher ids must be unique within a scenario";
        return ZRK_ERR_UNSUPPORTED;
    }
    return ZRK_OK;
}

// ZRK_TRACE_HOST=1: where the host thread spends zrk_create (microseconds since its start, to stderr)
struct HostTrace {
    bool on = getenv("ZRK_TRACE_HOST") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) const {
        if (on) fprintf(stderr, "zrk_create %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
};

static int create_impl(zrk_handle *h, const zrk_batch_desc *d, const zrk_options *opt) {
    HostTrace trace;
    if (d->abi_version != ZRK_ABI_VERSION) {
        h->err = "zrk_batch_desc.abi_version mismatch";
        return ZRK_ERR_ARG;
    }
    const int B = d->n_scenarios, T = d->max_targets, R = d->max_radars, L = d->max_launchers, M = d->max_missiles,
              C = d->max_ccp_launchers;
    if (B <= 0 || T < 0 || R < 0 || L < 0 || M < 0 || C < 0 || d->dt_ms <= 0) {
        h->err = "bad batch dimensions";
        return ZRK_ERR_ARG;
    }
    if ((T && (!d->target_id || !d->target_pos || !d->target_vel)) || (R && (!d->radar_id || !d->radar_pos || !d->radar_par)) ||
        (L && (!d->launcher_id || !d->launcher_pos)) ||
        (M && (!d->missile_id || !d->missile_launcher || !d->missile_v0 || !d->missile_radius || !d->missile_life)) ||
        (C && (!d->ccp_launcher_id || !d->ccp_launcher_pos))) {
        h->err = "missing array in zrk_batch_desc";
        return ZRK_ERR_ARG;
    }
    int ns_used = 0, m_used = 0;  // what the scenarios really have (the maxima may be padding)
    for (int b = 0; b < B; b++) {
        const int nt = d->n_targets ? d->n_targets[b] : T, nm = d->n_missiles ? d->n_missiles[b] : M;
        if (nt >= 0 && nt <= T && nm >= 0 && nm <= M) {  // (anything else is reported by the init kernel)
            ns_used = std::max(ns_used, nt + nm);
            m_used = std::max(m_used, nm);
        } else {
            ns_used = T + M;
            m_used = M;
        }
    }
    Layout lo = make_layout(B, T, M, R, L, C, ns_used, m_used);
    if (!lo.team || T + M > 1024 || M > 1023) {
        h->err = "unsupported: more than 1024 air-object slots per scenario";
        return ZRK_ERR_UNSUPPORTED;
    }
    if (d->noise && (d->noise_period <= 0 || d->noise_radars < R || d->noise_objects < T + M)) {
        h->err = "noise table does not cover max_radars x (max_targets + max_missiles)";
        return ZRK_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        h->err = "no CUDA device: zrk_b200 has no CPU fallback";
        return ZRK_ERR_NO_DEVICE;
    }
    h->device = opt ? opt->device : 0;
    if (h->device < 0 || h->device >= ndev) {
        h->err = "bad device ordinal";
        return ZRK_ERR_ARG;
    }
    CK(cudaSetDevice(h->device));
    CK(stream_get(h->device, &h->stream));
    h->has_stream = true;
    h->hd = *d;
    Params &P = h->P;
    P.lo = lo;
    P.dt_ms = d->dt_ms;
    P.dt_s = (double)d->dt_ms / 1000.0;
    P.range_step = ZRK_POSSIBLE_TARGET_RADIUS * P.dt_s;
    P.log_types = opt ? opt->log_types : (unsigned)ZRK_LOG_EVENTS;
    if (P.log_types & ZRK_LOG_MASKS) P.log_types |= ZRK_LOG_TRACKS;  // the track journal goes with the masks
    int cap = opt ? opt->log_capacity : 0;
    if (cap <= 0) {
        cap = 2 * C + M * (6 + R) + 4;  // every ZRK_LOG_EVENTS record a scenario can ever emit
        if (P.log_types & ZRK_LOG_TRACKS) cap += 8 * (T + M) + 64;  // + room for the track journal (an overflow is reported per scenario)
    }
    P.log_cap = cap;
    P.mask_words = (lo.NS + 31) / 32;
    if (P.mask_words < 1) P.mask_words = 1;
    P.mask_ticks = (P.log_types & ZRK_LOG_MASKS) ? ((opt && opt->mask_ticks > 0) ? opt->mask_ticks : 1024) : 0;
    h->ticks_per_launch = (opt && opt->ticks_per_launch > 0) ? opt->ticks_per_launch : 256;
    h->fast = !(opt && (opt->flags & ZRK_OPT_EXACT_ONLY));

    // scenario waves: whole CTAs and whole init blocks per wave
    h->wave_rounds = opt && opt->wave_rounds > 0 ? std::min((int)opt->wave_rounds, ZRK_FORK_ROUNDS_MAX) : 1;
    int n_waves = opt ? opt->waves : 0;
    if (n_waves < 1) n_waves = 1;
    if (n_waves > 64) n_waves = 64;
    {
        const int quantum = 128;
        const int per = ((B + n_waves - 1) / n_waves + quantum - 1) / quantum * quantum;
        h->wave_b.clear();
        for (int b0 = 0; b0 < B; b0 += per) h->wave_b.push_back(b0);
        h->wave_b.push_back(B);
        n_waves = (int)h->wave_b.size() - 1;
    }

    trace.mark("layout, stream");
    DevDesc dd{};
    dd.B = B; dd.T = T; dd.R = R; dd.L = L; dd.M = M; dd.C = C;
    int rc;
    // device copies of the description: allocated whole, filled wave by wave
    struct Field {
        const char *src;
        char *dst;
        size_t per_scenario;  // bytes
    };
    std::vector<Field> fields;
    h->defer_zero = true;
#define UP(field, n)                                                                                   \
    if (d->field && (size_t)(n) != 0) {                                                                             \
        void *q_ = nullptr;                                                                            \
        if ((rc = pool_alloc(h, &q_, (size_t)B * (n) * sizeof(*d->field))) != ZRK_OK) return rc;       \
        dd.field = (decltype(dd.field))q_;                                                             \
        fields.push_back({(const char *)d->field, (char *)q_, (size_t)(n) * sizeof(*d->field)});       \
    }
    UP(n_targets, 1); UP(n_radars, 1); UP(n_launchers, 1); UP(n_missiles, 1); UP(n_ccp_launchers, 1);
    UP(airenv_id, 1); UP(ccp_id, 1); UP(tick_limit, 1);
    UP(target_id, T); UP(target_type, T);
    UP(target_pos, (size_t)T * 3); UP(target_vel, (size_t)T * 3);
    UP(radar_id, R); UP(radar_pos, (size_t)R * 3); UP(radar_par, (size_t)R * 8);
    UP(launcher_id, L); UP(launcher_pos, (size_t)L * 3);
    UP(missile_id, M); UP(missile_launcher, M); UP(missile_v0, M);
    UP(missile_v0sq, M); UP(missile_radius, M); UP(missile_life, M);
    UP(ccp_launcher_id, C); UP(ccp_launcher_pos, (size_t)C * 3);
#undef UP
    const size_t n_desc_allocs = h->allocs.size();
    trace.mark("description allocations");
    if (d->noise) {
        P.noise_period = d->noise_period;
        P.noise_radars = d->noise_radars;
        P.noise_objects = d->noise_objects;
        P.noise_stride = d->noise_scenario_stride;
        size_t one = (size_t)d->noise_period * d->noise_radars * d->noise_objects * 3;
        size_t n = d->noise_scenario_stride ? (size_t)d->noise_scenario_stride * (B - 1) + one : one;
        if ((rc = dev_upload(h, d->noise, n, &P.noise)) != ZRK_OK) return rc;
        // the association certificate's bound on what the radars can move an object by in one tick:
        // sum over radars of the largest |noise vector| anywhere in the table (NaN / inf: certificate off)
        std::vector<double> mx2((size_t)d->noise_radars, 0.0);
        bool finite = true;
        const size_t per_w = (size_t)d->noise_radars * d->noise_objects;
        const int nsc = d->noise_scenario_stride ? B : 1;
        for (int sc = 0; sc < nsc; sc++) {
            const double *t = d->noise + (size_t)sc * d->noise_scenario_stride;
            for (size_t i = 0; i < (size_t)d->noise_period * per_w; i++) {
                const double x = t[3 * i], y = t[3 * i + 1], z = t[3 * i + 2], n2 = x * x + y * y + z * z;
                if (!(n2 < 1e300)) finite = false;
                double &m = mx2[(i % per_w) / (size_t)d->noise_objects];
                if (n2 > m) m = n2;
            }
        }
        double nb = 0.0;
        for (double m : mx2) nb += sqrt(m) * (1.0 + 1e-12);
        P.noise_bound = finite ? nb : -1.0;
    }
    P.cert = (opt && (opt->flags & (ZRK_OPT_EXACT_ONLY | ZRK_OPT_NO_CERTIFICATE))) || P.noise_bound < 0 ? 0 : 1;
    if ((rc = dev_alloc(h, (size_t)B * lo.dyn_bytes, &P.dyn, true)) != ZRK_OK) return rc;
    unsigned char *cst = nullptr;
    if ((rc = dev_alloc(h, (size_t)B * lo.cst_bytes, &cst, true)) != ZRK_OK) return rc;
    P.cst = cst;
    if ((rc = dev_alloc(h, (size_t)B * P.log_cap, &P.rec, false)) != ZRK_OK) return rc;
    if (P.mask_ticks) {
        size_t n = (size_t)P.mask_ticks * B * (1 + lo.RP) * P.mask_words;
        if ((rc = dev_alloc(h, n, &P.masks, true)) != ZRK_OK) return rc;
    }
    trace.mark("noise upload + bound, state allocations");
    int *derr = nullptr;
    if ((rc = dev_alloc(h, 1, &derr, true)) != ZRK_OK) return rc;
    if ((rc = dev_alloc(h, (size_t)B * ZRK_SUMMARY_WORDS, &h->d_summary, false)) != ZRK_OK) return rc;

    // launch configuration, clock table and the device copy of Params: everything that has to
    // synchronise does so before the first wave is queued
    h->block = lo.team * lo.teams_per_cta;
    h->grid = (B + lo.teams_per_cta - 1) / lo.teams_per_cta;
    h->smem_bytes = (size_t)lo.team_smem_bytes * lo.teams_per_cta;
    if (h->smem_bytes > 227 * 1024) {
        h->err = "unsupported: scenario state does not fit in shared memory";
        return ZRK_ERR_UNSUPPORTED;
    }
    tick_fn fn = pick_kernel(lo, h->fast);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    {
        // shared-memory carve-out: what the resident CTAs need (registers bound their number), so
        // that the rest of the 256 KB stays L1 for the noise table, the clock table and spills
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int blocks = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, fn, h->block, h->smem_bytes));
        const char *env = getenv("ZRK_CARVEOUT");
        int pct = env ? atoi(env) : (int)((100.0 * blocks * (h->smem_bytes + 1024)) / (228.0 * 1024) + 4);
        if (pct > 100) pct = 100;
        if (pct < 8) pct = 8;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    trace.mark("kernel attributes");
    if ((rc = ensure_tsec(h, 1024)) != ZRK_OK) return rc;
    trace.mark("clock table + params (2 syncs)");
    {
        const char *env = getenv("ZRK_TWO_KERNEL");
        h->two_kernel = lo.team == 32 && P.cert && h->fast && !(env && atoi(env) == 0);
        if (h->two_kernel) {
            if ((rc = dev_alloc(h, (size_t)n_waves * ZRK_ROUND_SLOTS, &h->d_pending, true)) != ZRK_OK) return rc;
            if ((rc = dev_alloc(h, (size_t)ZRK_ROUND_SLOTS + 64 * ZRK_FORK_ROUNDS_MAX, &h->d_queue, true)) != ZRK_OK) return rc;
            CK(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, h->device));
            CK(pinned_get((void **)&h->h_pending, sizeof(int) * 64));
            if ((rc = dev_alloc(h, (size_t)2 * B, &h->d_work, false)) != ZRK_OK) return rc;
        }
    }

    trace.mark("round buffers");
    // upload + init, wave by wave.  The buffers the init kernel fills are zeroed now, after the last host
    // synchronisation above (nothing waited for the memsets); a wave's copies only need the allocations, its init
    // kernel the memsets as well.
    if (n_waves > 1) {
        CK(pinned_get((void **)&h->pin_err, sizeof(int)));
        *h->pin_err = 0;
        CK(event_get(h->device, false, &h->alloc_ev));
        CK(cudaEventRecord(h->alloc_ev, h->stream));  // the allocations above
    }
    h->defer_zero = false;
    for (auto &z : h->zero_list) CK(cudaMemsetAsync(z.first, 0, z.second, h->stream));
    h->zero_list.clear();
    if (n_waves > 1) {
        CK(event_get(h->device, false, &h->fork_ev));
        CK(cudaEventRecord(h->fork_ev, h->stream));  // ... and the memsets
        for (int w = 0; w < n_waves; w++) {
            cudaStream_t ws;
            cudaEvent_t ev;
            CK(stream_get(h->device, &ws));
            h->wave_stream.push_back(ws);
            CK(event_get(h->device, false, &ev));
            h->wave_done.push_back(ev);
            CK(cudaStreamWaitEvent(ws, h->alloc_ev, 0));
        }
    }
    // A description in pageable host memory (what a caller holding numpy arrays passes) would be copied by the driver
    // through its own staging buffer, one call at a time, at a fraction of the link's rate and with the host thread
    // blocked.  It is staged here instead: helper threads move it, wave by wave, into one page-locked buffer (kept for
    // the next handle of the process), and every wave's copies are queued from there as soon as the wave is in.
    size_t desc_bytes = 0;
    for (const Field &f : fields) desc_bytes += (size_t)B * f.per_scenario;
    bool staged = false;
#ifndef ZRK_EMU
    if (!fields.empty() && desc_bytes >= ((size_t)4 << 20) && !getenv("ZRK_NO_STAGING")) {
        cudaPointerAttributes at{};
        size_t big = 0;
        for (size_t i = 1; i < fields.size(); i++)
            if (fields[i].per_scenario > fields[big].per_scenario) big = i;
        if (cudaPointerGetAttributes(&at, fields[big].src) != cudaSuccess) cudaGetLastError();
        else staged = at.type == cudaMemoryTypeUnregistered;
    }
#endif
    struct Seg {
        char *dst;
        const char *src;
        size_t n;
        int wave;
    };
    std::vector<Seg> segs;
    std::vector<std::vector<std::pair<size_t, size_t>>> wave_off((size_t)n_waves);  // per wave, per field: offset in the staging buffer
    std::vector<std::thread> helpers;
    std::atomic<size_t> seg_next{0};
    std::vector<std::atomic<int>> seg_left((size_t)n_waves);
    if (staged) {
        CK(pinned_get(&h->bounce, desc_bytes));
        h->bounce_bytes = desc_bytes;
        size_t off = 0;
        for (int w = 0; w < n_waves; w++) {
            const int b0 = h->wave_b[w], b1 = h->wave_b[w + 1];
            int n_seg = 0;
            for (const Field &f : fields) {
                const size_t n = (size_t)(b1 - b0) * f.per_scenario;
                wave_off[(size_t)w].push_back({off, n});
                for (size_t o2 = 0; o2 < n; o2 += (size_t)1 << 20, n_seg++)
                    segs.push_back({(char *)h->bounce + off + o2, f.src + (size_t)b0 * f.per_scenario + o2, std::min(n - o2, (size_t)1 << 20), w});
                off += n;
            }
            seg_left[(size_t)w].store(n_seg);
        }
        auto work = [&]() {
            for (size_t i; (i = seg_next.fetch_add(1)) < segs.size();) {
                memcpy(segs[i].dst, segs[i].src, segs[i].n);
                seg_left[(size_t)segs[i].wave].fetch_sub(1, std::memory_order_release);
            }
        };
        unsigned hc = std::thread::hardware_concurrency();
#if defined(__linux__)
        {  // (the cores this thread may run on: a rank bound to a few cores of its GPU's node gets that many helpers)
            cpu_set_t cs;
            CPU_ZERO(&cs);
            if (sched_getaffinity(0, sizeof cs, &cs) == 0 && CPU_COUNT(&cs) > 0) hc = (unsigned)CPU_COUNT(&cs) + 1;
        }
#endif
        const unsigned want = getenv("ZRK_STAGING_THREADS") ? (unsigned)std::max(1, atoi(getenv("ZRK_STAGING_THREADS"))) : 8u;
        const int n_help = (int)std::min<size_t>(std::min<unsigned>(hc > 1 ? hc - 1 : 1, want), segs.size());
        for (int t = 0; t < n_help; t++) helpers.emplace_back(work);
    }
    struct JoinHelpers {
        std::vector<std::thread> &v;
        ~JoinHelpers() {
            for (auto &t : v)
                if (t.joinable()) t.join();
        }
    } join_helpers{helpers};
    for (int w = 0; w < n_waves; w++) {
        const int b0 = h->wave_b[w], b1 = h->wave_b[w + 1];
        cudaStream_t ws = n_waves > 1 ? h->wave_stream[w] : h->stream;
        if (staged) {
            // (this thread helps with the wave it is waiting for, then queues its copies)
            for (size_t i; seg_left[(size_t)w].load(std::memory_order_acquire) > 0;) {
                if ((i = seg_next.fetch_add(1)) < segs.size()) {
                    memcpy(segs[i].dst, segs[i].src, segs[i].n);
                    seg_left[(size_t)segs[i].wave].fetch_sub(1, std::memory_order_release);
                } else {
                    std::this_thread::yield();
                }
            }
            size_t k = 0;
            for (const Field &f : fields) {
                const auto &po = wave_off[(size_t)w][k++];
                CK(cudaMemcpyAsync(f.dst + (size_t)b0 * f.per_scenario, (const char *)h->bounce + po.first, po.second, cudaMemcpyHostToDevice, ws));
            }
        } else {
            for (const Field &f : fields)
                CK(cudaMemcpyAsync(f.dst + (size_t)b0 * f.per_scenario, f.src + (size_t)b0 * f.per_scenario,
                                   (size_t)(b1 - b0) * f.per_scenario, cudaMemcpyHostToDevice, ws));
        }
        if (n_waves > 1) CK(cudaStreamWaitEvent(ws, h->fork_ev, 0));
        ZRK_LAUNCH(zrk_init_kernel, (b1 - b0 + ZRK_INIT_BLOCK / 32 - 1) / (ZRK_INIT_BLOCK / 32), ZRK_INIT_BLOCK, 0, ws, P, dd, derr, b0, b1);
        CK(cudaGetLastError());
        if (n_waves > 1) {
            CK(cudaEventRecord(h->wave_done[w], ws));
            CK(cudaStreamWaitEvent(h->stream, h->wave_done[w], 0));
        }
    }
    trace.mark("streams, copies, init queued");
    // small host copies used by zrk_read_messages
    h->h_radar_id.assign(d->radar_id ? d->radar_id : nullptr, d->radar_id ? d->radar_id + (size_t)B * R : nullptr);
    h->h_n_radars.resize(B);
    h->h_airenv_id.resize(B);
    for (int b = 0; b < B; b++) {
        h->h_n_radars[b] = d->n_radars ? d->n_radars[b] : R;
        h->h_airenv_id[b] = d->airenv_id ? d->airenv_id[b] : 1;
    }
    trace.mark("host copies of ids");
    if (n_waves > 1) {
        // do not wait: the verdict of the init kernel is looked at by the first call that synchronises
        CK(cudaMemcpyAsync(h->pin_err, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        for (size_t i = 0; i < n_desc_allocs; i++) cudaFreeAsync(h->allocs[i], h->stream);
        h->allocs.erase(h->allocs.begin(), h->allocs.begin() + n_desc_allocs);
        h->deferred = true;
        return ZRK_OK;
    }
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // the uploaded description is no longer needed on the device
    for (size_t i = 0; i < n_desc_allocs; i++) cudaFreeAsync(h->allocs[i], h->stream);
    h->allocs.erase(h->allocs.begin(), h->allocs.begin() + n_desc_allocs);
    return init_verdict(h, herr);
}

extern "C" int zrk_create(const zrk_batch_desc *desc, const zrk_options *opt, zrk_handle **out) {
    if (!desc || !out) {
        g_create_err = "null argument";
        return ZRK_ERR_ARG;
    }
    *out = nullptr;
    zrk_handle *h = new zrk_handle();
    int caller_dev = -1;
    if (cudaGetDevice(&caller_dev) != cudaSuccess) caller_dev = -1;
    int rc = create_impl(h, desc, opt);
    if (caller_dev >= 0) cudaSetDevice(caller_dev);
    if (rc != ZRK_OK) {
        g_create_err = h->err;
        zrk_destroy(h);
        return rc;
    }
    *out = h;
    return ZRK_OK;
}

static int resolve_events(zrk_handle *h) {
    for (auto &pr : h->pending) {
        float ms = 0;
        CK(cudaEventSynchronize(pr.second));
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        h->kernel_ms += ms;
        h->free_events.push_back(pr.first);
        h->free_events.push_back(pr.second);
    }
    h->pending.clear();
    return ZRK_OK;
}
static int get_event(zrk_handle *h, cudaEvent_t *e) {
    if (!h->free_events.empty()) {
        *e = h->free_events.back();
        h->free_events.pop_back();
        return ZRK_OK;
    }
    CK(event_get(h->device, true, e));
    return ZRK_OK;
}

extern "C" int zrk_step(zrk_handle *h, int n_ticks) {
    if (!h || n_ticks < 0) return ZRK_ERR_ARG;
    if (h->sticky_rc) return h->sticky_rc;
    DeviceGuard dev_guard(h->device);
    CK(cudaSetDevice(h->device));
    tick_fn fn = pick_kernel(h->P.lo, h->fast);
    if (h->pending.size() > 4096) {
        int rc = resolve_events(h);
        if (rc) return rc;
    }
    {
        int rc = ensure_tsec(h, h->tick + n_ticks + 1);  // (synchronises only when the table has to grow)
        if (rc) return rc;
    }
    const int n_waves = h->wave_stream.empty() ? 1 : (int)h->wave_stream.size();
    const int tpc = h->P.lo.teams_per_cta;
    if (h->two_kernel && n_ticks >= 8) {
        // Two-kernel stepping: every round runs the streak kernel (quiet ticks, straight from HBM) and then the
        // tick kernel (whatever else) over the batch; a scenario advances in whichever of the two its next
        // tick belongs to, until all have reached the target.  The number of rounds depends on the data, so
        // the host looks at the pending counters after every few rounds.
        const int target = (int)(h->tick + n_ticks);
        streak_fn sf = pick_streak(h->P.lo);
        const int B = h->P.lo.B;
        cudaStream_t ms = h->stream;
        static const int round_batch = getenv("ZRK_ROUND_BATCH") ? std::max(1, atoi(getenv("ZRK_ROUND_BATCH"))) : ZRK_ROUND_BATCH;
        const int fork_rounds = std::min(h->wave_rounds, round_batch);
        static const int tail_div = getenv("ZRK_TAIL_DIV") ? std::max(1, atoi(getenv("ZRK_TAIL_DIV"))) : 64;
        // Scenario waves: the first round after zrk_create runs wave by wave on the waves' streams, where their uploads
        // and init kernels are queued (the copy of one wave overlaps the kernels of the waves before it); all waves
        // append to ONE worklist, and every later round runs over the whole batch on the main stream.
        bool fork_first = n_waves > 1 && !h->stepped;
        h->stepped = true;
        // how many scenarios the last batch of rounds left pending (-1: the whole batch, no worklist yet), and which of
        // the two worklist buffers holds them
        int n_work = -1, cur = 0;
        for (;;) {
            if (h->round_no + round_batch > ZRK_ROUND_SLOTS) {
                CK(cudaMemsetAsync(h->d_pending, 0, sizeof(int) * ZRK_ROUND_SLOTS, ms));
                CK(cudaMemsetAsync(h->d_queue, 0, sizeof(int) * (ZRK_ROUND_SLOTS + 64 * ZRK_FORK_ROUNDS_MAX), ms));
                h->round_no = 0;
                fork_first = false;  // (the waves' streams are not ordered behind this memset)
            }
            // Every round leaves the list of the scenarios still pending (device-side compaction through the round's
            // pending counter) in the worklist buffer the next round reads; the host only learns its length after
            // the batch, so the grids of a batch are sized for the last length it knows and the kernels look the
            // real one up (work_n).
            for (int r = 0; r < round_batch; r++) {
                // (forked: the first zrk_options.wave_rounds rounds after zrk_create, wave by wave over the waves' ranges;
                // the last of them writes the worklist)
                const bool forked = fork_first && n_work < 0 && r < fork_rounds;
                const bool listed = !forked && (n_work >= 0 || r > 0);  // (the first round of a step looks at the whole range)
                const int *wl = listed ? h->d_work + (size_t)cur * B : nullptr;
                int *wl_out = forked && r < fork_rounds - 1 ? nullptr : h->d_work + (size_t)(cur ^ 1) * B;
                int *pend = h->d_pending + h->round_no + r;
                const int *wn = listed && r > 0 ? pend - 1 : nullptr;  // (r = 0: the host knows the length)
                const int g1 = n_work < 0 ? B : n_work;                // index range of the launches: [0, g1)
                const int n_parts = forked ? n_waves : 1;
                for (int w = 0; w < n_parts; w++) {
                    const int p0 = n_parts > 1 ? h->wave_b[w] : 0, p1 = n_parts > 1 ? h->wave_b[w + 1] : g1;
                    cudaStream_t ws = n_parts > 1 ? h->wave_stream[w] : ms;
                    cudaEvent_t e0, e1;
                    int rc;
                    if ((rc = get_event(h, &e0)) || (rc = get_event(h, &e1))) return rc;
                    CK(cudaEventRecord(e0, ws));
                    // (persistent: at most the CTAs the SMs hold at once; every warp draws scenarios from the launch's queue)
                    const int s_ctas = std::min((p1 - p0 + ZRK_STREAK_WARPS - 1) / ZRK_STREAK_WARPS, std::max(h->n_sms, 1) * ZRK_STREAK_MINB);
                    ZRK_LAUNCH(sf, s_ctas, 32 * ZRK_STREAK_WARPS, ZRK_STREAK_SMEM, ws, h->P, p0, p1, target, wl, wn,
                               h->d_queue + (n_parts > 1 ? ZRK_ROUND_SLOTS + 64 * r + w : h->round_no + r));
                    ZRK_LAUNCH(fn, (p1 - p0 + tpc - 1) / tpc, h->block, h->smem_bytes, ws, h->P, 0, p0, p1, target, pend, wl, wl_out, wn);
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(e1, ws));
                    h->pending.push_back({e0, e1});
                    h->n_launches += 2;
                    if (n_parts > 1 && r == fork_rounds - 1) {  // the main stream goes on after the waves' last forked round
                        CK(cudaEventRecord(h->wave_done[w], ws));
                        CK(cudaStreamWaitEvent(ms, h->wave_done[w], 0));
                    }
                }
                cur ^= 1;
            }
            fork_first = false;
            CK(cudaMemcpyAsync(h->h_pending, h->d_pending + h->round_no + round_batch - 1, sizeof(int), cudaMemcpyDeviceToHost, ms));
            h->round_no += round_batch;
            h->n_rounds += round_batch;
            CK(cudaStreamSynchronize(ms));
            n_work = h->h_pending[0];
            if (n_work > 0 && n_work <= B / tail_div) {
                // the tail: a few scenarios with many rounds still to go (one tick per round, in the worst
                // case) are finished by one launch of the tick kernel on its own
                cudaEvent_t e0, e1;
                int rc;
                if ((rc = get_event(h, &e0)) || (rc = get_event(h, &e1))) return rc;
                CK(cudaEventRecord(e0, ms));
                ZRK_LAUNCH(fn, (n_work + tpc - 1) / tpc, h->block, h->smem_bytes, ms, h->P, 0, 0, n_work, target, (int *)nullptr,
                           (const int *)(h->d_work + (size_t)cur * B), (int *)nullptr, (const int *)nullptr);
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1, ms));
                h->pending.push_back({e0, e1});
                h->n_launches++;
                n_work = 0;
            }
            if (h->pending.size() > 2048) {
                int rc = resolve_events(h);
                if (rc) return rc;
            }
            if (n_work == 0) break;
        }
        h->tick += n_ticks;
        if (!h->h_summary)
            CK(pinned_get((void **)&h->h_summary, (size_t)h->P.lo.B * ZRK_SUMMARY_WORDS * sizeof(int)));
        return ZRK_OK;
    }
    // Every call that queues work on the main stream also waits for it (reads, table growth), so a
    // wave never has to wait for the main stream here — and must not: right after zrk_create the
    // main stream is queued behind the uploads of ALL waves, which is exactly what a wave's first
    // kernels are meant to overlap.
    for (int w = 0; w < n_waves; w++) {
        const int b0 = n_waves > 1 ? h->wave_b[w] : 0, b1 = n_waves > 1 ? h->wave_b[w + 1] : h->P.lo.B;
        cudaStream_t ws = n_waves > 1 ? h->wave_stream[w] : h->stream;
        int left = n_ticks;
        while (left > 0) {
            int k = left < h->ticks_per_launch ? left : h->ticks_per_launch;
            cudaEvent_t e0, e1;
            int rc;
            if ((rc = get_event(h, &e0)) || (rc = get_event(h, &e1))) return rc;
            CK(cudaEventRecord(e0, ws));
            ZRK_LAUNCH(fn, (b1 - b0 + tpc - 1) / tpc, h->block, h->smem_bytes, ws, h->P, k, b0, b1, -1, (int *)nullptr,
                       (const int *)nullptr, (int *)nullptr, (const int *)nullptr);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1, ws));
            h->pending.push_back({e0, e1});
            h->n_launches++;
            left -= k;
        }
        if (n_waves > 1) {  // and the main stream continues after all of them
            CK(cudaEventRecord(h->wave_done[w], ws));
            CK(cudaStreamWaitEvent(h->stream, h->wave_done[w], 0));
        }
    }
    h->tick += n_ticks;
    if (!h->h_summary && n_ticks > 0)  // page-lock the read-back buffer while the GPU is busy
        CK(pinned_get((void **)&h->h_summary, (size_t)h->P.lo.B * ZRK_SUMMARY_WORDS * sizeof(int)));
    return ZRK_OK;
}

extern "C" int zrk_sync(zrk_handle *h) {
    if (!h) return ZRK_ERR_ARG;
    if (h->sticky_rc) return h->sticky_rc;
    DeviceGuard dev_guard(h->device);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->deferred) {  // zrk_create with waves did not wait for the init kernel's verdict
        h->deferred = false;
        int rc = init_verdict(h, *h->pin_err);
        if (rc) {
            h->sticky_rc = rc;
            return rc;
        }
    }
    return resolve_events(h);
}

extern "C" int zrk_get_time(zrk_handle *h, int64_t *tick, int64_t *time_ms) {
    if (!h) return ZRK_ERR_ARG;
    if (tick) *tick = h->tick;
    if (time_ms) *time_ms = h->tick * h->P.dt_ms;
    return ZRK_OK;
}

extern "C" int zrk_launch_stats(zrk_handle *h, int64_t *n_launches, double *kernel_ms) {
    if (!h) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = zrk_sync(h);
    if (rc) return rc;
    if (n_launches) *n_launches = h->n_launches;
    if (kernel_ms) *kernel_ms = h->kernel_ms;
    return ZRK_OK;
}

extern "C" int zrk_log_capacity(zrk_handle *h) { return h ? h->P.log_cap : ZRK_ERR_ARG; }

// copy the dyn blobs of scenarios [s0, s1) to host
static int fetch_dyn(zrk_handle *h, int s0, int s1, std::vector<unsigned char> &out) {
    int rc = zrk_sync(h);
    if (rc) return rc;
    const Layout &o = h->P.lo;
    out.resize((size_t)(s1 - s0) * o.dyn_bytes);
    if (s1 > s0)
        CK(cudaMemcpy(out.data(), h->P.dyn + (size_t)s0 * o.dyn_bytes, out.size(), cudaMemcpyDeviceToHost));
    return ZRK_OK;
}

static int fetch_summary(zrk_handle *h) {
    const Layout &o = h->P.lo;
    CK(cudaSetDevice(h->device));
    const long long n = (long long)o.B * ZRK_SUMMARY_WORDS;
    ZRK_LAUNCH(zrk_summary_kernel, (unsigned)((n + 255) / 256), 256, 0, h->stream, h->P, h->d_summary);
    CK(cudaGetLastError());
    if (!h->h_summary) CK(pinned_get((void **)&h->h_summary, (size_t)n * sizeof(int)));
    CK(cudaMemcpyAsync(h->h_summary, h->d_summary, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return zrk_sync(h);
}

extern "C" int zrk_read_status(zrk_handle *h, int32_t *fault_code, int32_t *fault_tick, int32_t *n_records) {
    if (!h) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = fetch_summary(h);
    if (rc) return rc;
    for (int s = 0; s < h->P.lo.B; s++) {
        const int *hdr = h->h_summary + (size_t)s * ZRK_SUMMARY_WORDS;
        if (fault_code) fault_code[s] = hdr[H_FAULT];
        if (fault_tick) fault_tick[s] = hdr[H_FAULT_TICK];
        if (n_records) n_records[s] = hdr[H_NREC];
    }
    return ZRK_OK;
}

extern "C" int zrk_read_counts(zrk_handle *h, int64_t *counts) {
    if (!h || !counts) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = fetch_summary(h);
    if (rc) return rc;
    for (int s = 0; s < h->P.lo.B; s++) {
        const unsigned *c = (const unsigned *)(h->h_summary + (size_t)s * ZRK_SUMMARY_WORDS + H_WORDS);
        for (int t = 0; t < ZRK_MSG_NTYPES; t++) counts[(size_t)s * ZRK_MSG_NTYPES + t] = c[t];
    }
    return ZRK_OK;
}

extern "C" int zrk_read_totals(zrk_handle *h, int64_t *totals) {
    if (!h || !totals) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    if (!h->d_totals) {
        int rc = dev_alloc(h, (size_t)ZRK_TOTALS_WORDS, &h->d_totals, false);
        if (rc) return rc;
        CK(pinned_get((void **)&h->h_totals, sizeof(unsigned long long) * ZRK_TOTALS_WORDS));
    }
    CK(cudaMemsetAsync(h->d_totals, 0, sizeof(unsigned long long) * ZRK_TOTALS_WORDS, h->stream));
    ZRK_LAUNCH(zrk_totals_kernel, (h->P.lo.B + 255) / 256, 256, sizeof(unsigned long long) * ZRK_TOTALS_WORDS, h->stream, h->P,
               h->d_totals);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_totals, h->d_totals, sizeof(unsigned long long) * ZRK_TOTALS_WORDS, cudaMemcpyDeviceToHost, h->stream));
    int rc = zrk_sync(h);
    if (rc) return rc;
    for (int i = 0; i < ZRK_TOTALS_WORDS; i++) totals[i] = (int64_t)h->h_totals[i];
    return ZRK_OK;
}

extern "C" int zrk_read_activity(zrk_handle *h, int64_t *alive_slot_ticks) {
    if (!h || !alive_slot_ticks) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = fetch_summary(h);
    if (rc) return rc;
    for (int s = 0; s < h->P.lo.B; s++)
        alive_slot_ticks[s] = (unsigned)h->h_summary[(size_t)s * ZRK_SUMMARY_WORDS + H_ALIVE];
    return ZRK_OK;
}

extern "C" int zrk_read_certified(zrk_handle *h, int64_t *certified_ticks) {
    if (!h || !certified_ticks) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = fetch_summary(h);
    if (rc) return rc;
    for (int s = 0; s < h->P.lo.B; s++)
        certified_ticks[s] = (unsigned)h->h_summary[(size_t)s * ZRK_SUMMARY_WORDS + H_FAST];
    return ZRK_OK;
}

extern "C" int zrk_read_records(zrk_handle *h, int s0, int s1, zrk_msg *recs, int32_t *n) {
    if (!h || s0 < 0 || s1 > h->P.lo.B || s0 > s1) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    int rc = fetch_summary(h);
    if (rc) return rc;
    for (int s = s0; s < s1; s++) {
        const int *hdr = h->h_summary + (size_t)s * ZRK_SUMMARY_WORDS;
        if (n) n[s - s0] = std::min(hdr[H_NREC], h->P.log_cap);
    }
    if (recs && s1 > s0) {
        // The log leaves the device in chunks of whole scenarios through two page-locked buffers (a copy straight into
        // the caller's pageable array runs at a tenth of the link's rate); while one chunk is in flight the host moves
        // the one before it into place — each scenario's records only, the private track journal squeezed out.
        const size_t cap = (size_t)h->P.log_cap, per = cap * sizeof(zrk_msg);
        const int chunk_sc = (int)std::max<size_t>(1, ((size_t)8 << 20) / per);
        const bool squeeze = (h->P.log_types & ZRK_LOG_TRACKS) != 0;
        void *pin[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        for (int k = 0; k < 2; k++) {
            CK(pinned_get(&pin[k], (size_t)chunk_sc * per));
            CK(event_get(h->device, false, &ev[k]));
        }
        auto place_part = [&](int c0, int c1, const zrk_msg *src, int c_base) {
            for (int s = c0; s < c1; s++) {
                const zrk_msg *r = src + (size_t)(s - c_base) * cap;
                zrk_msg *dst = recs + (size_t)(s - s0) * cap;
                const int nr = std::min(h->h_summary[(size_t)s * ZRK_SUMMARY_WORDS + H_NREC], h->P.log_cap);
                if (!squeeze) {
                    memcpy(dst, r, (size_t)nr * sizeof(zrk_msg));
                } else {
                    int k = 0;
                    for (int i = 0; i < nr; i++)
                        if (r[i].type != ZRK_REC_TRACK) dst[k++] = r[i];
                    if (n) n[s - s0] = k;
                }
            }
        };
        // (the caller's array is often fresh memory: its first touch is what takes the time, and goes in parallel)
        auto place = [&](int c0, int c1, const zrk_msg *src) {
            const int n_thr = (c1 - c0) >= 256 ? 4 : 1, per_thr = (c1 - c0 + n_thr - 1) / n_thr;
            std::vector<std::thread> th;
            for (int t = 1; t < n_thr; t++)
                th.emplace_back(place_part, std::min(c1, c0 + t * per_thr), std::min(c1, c0 + (t + 1) * per_thr), src, c0);
            place_part(c0, std::min(c1, c0 + per_thr), src, c0);
            for (auto &t : th) t.join();
        };
        int rc_cuda = 0, prev0 = -1, prev1 = -1, slot = 0;
        for (int c0 = s0; c0 < s1 && !rc_cuda; c0 += chunk_sc, slot ^= 1) {
            const int c1 = std::min(s1, c0 + chunk_sc);
            if (cudaMemcpyAsync(pin[slot], h->P.rec + (size_t)c0 * cap, (size_t)(c1 - c0) * per, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                cudaEventRecord(ev[slot], h->stream) != cudaSuccess)
                rc_cuda = 1;
            if (prev0 >= 0 && !rc_cuda) {
                if (cudaEventSynchronize(ev[slot ^ 1]) != cudaSuccess) rc_cuda = 1;
                else place(prev0, prev1, (const zrk_msg *)pin[slot ^ 1]);
            }
            prev0 = c0, prev1 = c1;
        }
        if (prev0 >= 0 && !rc_cuda) {
            if (cudaEventSynchronize(ev[slot ^ 1]) != cudaSuccess) rc_cuda = 1;
            else place(prev0, prev1, (const zrk_msg *)pin[slot ^ 1]);
        }
        cudaStreamSynchronize(h->stream);
        for (int k = 0; k < 2; k++) {
            pinned_put(pin[k], (size_t)chunk_sc * per);
            event_put(h->device, false, ev[k]);
        }
        if (rc_cuda) {
            h->err = "zrk_read_records: device-to-host copy failed";
            return ZRK_ERR_CUDA;
        }
    }
    return ZRK_OK;
}

// ------------------------------------------------------------------ messages rebuilt on the host
//
// The log that scales: per scenario-tick the device keeps 1 + R mask words (who is in AirEnv's list, whom
// each radar saw) and, per scenario, the rare event records plus the track journal.  Everything else the
// reference stores in Manager.messages (Manager.py:41-58) follows from those:
//   MISSILE_POS        one per active missile per tick, in slot order (Missile.py:183-184)
//   ACTIVE/ALL/FOUND   the lists behind the masks (AirEnv.py:50-53, Radar.py:155-174)
//   CCP_UPDATE_TARGET  one per detection whose track is followed (CCP.py:350-357)
//   DRAW_OBJECTS       tracked missiles seen, tracked targets seen, everything else (CCP.py:237-284), with the
//                      object's position of that tick: Trajectory.get_pos (AirObject.py:23-25) plus the noise of
//                      every radar that saw it, in radar order (Radar.py:138-142) — the same fp64 operations
// The CCP's dicts are replayed from tick 0: LAUNCHED_MISSILE records (add_missile, CCP.py:102-109),
// LAUNCH_COMMAND records (following) and the track journal (every association other than "own track").
static int rebuild_messages(zrk_handle *h, int sc, int tick0, int tick1, std::vector<zrk_msg> &out, std::vector<int32_t> &pool) {
    const Params &P = h->P;
    const Layout &o = P.lo;
    const HotOff &ho = o.hot;
    std::vector<unsigned char> dyn, cst((size_t)o.cst_bytes);
    int rc = fetch_dyn(h, sc, sc + 1, dyn);
    if (rc) return rc;
    CK(cudaMemcpy(cst.data(), P.cst + (size_t)sc * o.cst_bytes, cst.size(), cudaMemcpyDeviceToHost));
    const int *hdr = (const int *)(dyn.data() + ho.hdr), *sid = (const int *)(dyn.data() + ho.sid),
              *sflags = (const int *)(dyn.data() + ho.sflags);
    const double *txy = (const double *)(dyn.data() + ho.txy), *tzt = (const double *)(dyn.data() + ho.tzt),
                 *vxy = (const double *)(dyn.data() + ho.vxy), *qzw = (const double *)(dyn.data() + ho.qzw);
    const int *chdr = (const int *)(cst.data() + o.c_hdr), *c_rad_id = (const int *)(cst.data() + o.c_rad_id);
    const int nR = chdr[C_NR], ccp_id = chdr[C_CCP_ID], airenv_id = chdr[C_AIRENV_ID];
    const bool has_ccp = ccp_id != ZRK_NONE;
    const int n_slots = hdr[H_NSLOTS], TEAM = o.team;
    const int fault = hdr[H_FAULT] & 63, fault_tick = hdr[H_FAULT_TICK];
    const int done = hdr[H_TICK] + (fault ? 1 : 0);  // the faulting tick's messages exist too
    if (hdr[H_FAULT] & ZRK_FAULT_LOG_OVERFLOW) {
        h->err = "record log overflow: raise zrk_options.log_capacity";
        return ZRK_ERR_CAPACITY;
    }
    if (tick1 > done) tick1 = done;
    if (tick1 <= tick0) return ZRK_OK;
    if (done > P.mask_ticks) {  // the dicts are replayed from tick 0
        h->err = "mask history ring is shorter than the run: raise zrk_options.mask_ticks";
        return ZRK_ERR_CAPACITY;
    }
    const int nrec = std::min(hdr[H_NREC], P.log_cap);
    std::vector<zrk_msg> rec(nrec ? nrec : 1);
    if (nrec) CK(cudaMemcpy(rec.data(), P.rec + (size_t)sc * P.log_cap, (size_t)nrec * sizeof(zrk_msg), cudaMemcpyDeviceToHost));
    // mask rows of this scenario, ticks [0, tick1): one strided copy
    const int W = P.mask_words, rows = (1 + o.RP) * W;
    std::vector<unsigned> masks((size_t)tick1 * rows);
    CK(cudaMemcpy2D(masks.data(), rows * sizeof(unsigned), P.masks + (size_t)sc * rows, (size_t)o.B * rows * sizeof(unsigned),
                    rows * sizeof(unsigned), tick1, cudaMemcpyDeviceToHost));
    // the scenario's noise table
    std::vector<double> noise;
    const size_t nrow = (size_t)P.noise_radars * P.noise_objects * 3;
    if (P.noise) {
        noise.resize((size_t)P.noise_period * nrow);
        CK(cudaMemcpy(noise.data(), P.noise + (size_t)sc * P.noise_stride, noise.size() * sizeof(double), cudaMemcpyDeviceToHost));
    }
    auto slot_of_id = [&](int id) {
        for (int k = 0; k < n_slots; k++)
            if (sid[k] == id) return k;
        return -1;
    };
    // the CCP's dicts (CCP.py:80-81), as the device keeps them: key, slot the entry's object sits in, ...
    std::vector<int> te_key, te_ref, te_follow, te_mid, me_key, me_ref, own(n_slots > 0 ? n_slots : 1, -1);
    std::vector<int> rf(n_slots > 0 ? n_slots : 1);
    int ri = 0, det_prev = 0;
    for (int t = 0; t < tick1; t++) {
        const bool emit = t >= tick0;
        const size_t first = out.size();
        const unsigned *m = masks.data() + (size_t)t * rows;
        auto bit = [&](const unsigned *row, int k) { return (row[k >> 5] >> (k & 31)) & 1u; };
        auto push = [&](const zrk_msg &x) { if (emit) out.push_back(x); };
        auto list_msg = [&](int type, int sender, int receiver, const unsigned *bits) {
            if (!emit) return;
            zrk_msg x{};
            x.tick = t; x.type = (uint8_t)type; x.relevance = 1; x.sender = sender; x.receiver = receiver;
            x.a = (int32_t)pool.size();
            int cnt = 0;
            for (int k = 0; k < n_slots; k++)
                if (bit(bits, k)) { pool.push_back(sid[k]); cnt++; }
            x.b = cnt;
            out.push_back(x);
        };
        // this tick's records, by phase
        const int r0 = ri;
        while (ri < nrec && rec[ri].tick == t) ri++;
        const int r1 = ri;
        const double now_s = (double)((long long)t * P.dt_ms) / 1000.0;
        // AirEnv.step: per active missile (slot order) MISSILE_POS, then its MISSILE_DETONATE if it has one
        int det_now = 0;
        for (int k = 0; k < n_slots; k++) {
            if (!bit(m, k) || !(sflags[k] & F_MISSILE)) continue;
            zrk_msg x{};
            x.tick = t; x.type = ZRK_MSG_MISSILE_POS; x.relevance = 1; x.sender = sid[k]; x.receiver = ZRK_NONE; x.a = sid[k]; x.b = ZRK_NONE;
            push(x);
            for (int i = r0; i < r1; i++)
                if (rec[i].type == ZRK_MSG_MISSILE_DETONATE && rec[i].a == sid[k]) { push(rec[i]); det_now++; }
        }
        list_msg(ZRK_MSG_ACTIVE_OBJECTS, airenv_id, ZRK_NONE, m);
        // radars: ALL_OBJECTS, FOUND_OBJECTS, one DESTROYED_MISSILE per detonation of the previous tick
        int di = r0;
        for (int r = 0; r < nR; r++) {
            list_msg(ZRK_MSG_ALL_OBJECTS, c_rad_id[r], ZRK_CCP_ID, m);
            list_msg(ZRK_MSG_FOUND_OBJECTS, c_rad_id[r], ZRK_CCP_ID, m + (1 + r) * W);
            for (int k = 0; k < det_prev; k++) {
                while (di < r1 && rec[di].type != ZRK_MSG_DESTROYED_MISSILE) di++;
                if (di < r1) push(rec[di++]);
            }
        }
        // launchers and the CCP's bookkeeping, in emission order; LAUNCHED_MISSILE feeds add_missile (CCP.py:102-109)
        bool key_fault = false;
        for (int i = r0; i < r1; i++) {
            const int ty = rec[i].type;
            if (ty == ZRK_MSG_MISSILE_DETONATE || ty == ZRK_MSG_DESTROYED_MISSILE || ty == ZRK_MSG_LAUNCH_COMMAND ||
                ty == ZRK_REC_TRACK || ty == ZRK_MSG_CCP_UPDATE_TARGET || ty == ZRK_MSG_DRAW_OBJECTS || ty == ZRK_MSG_MISSILE_POS)
                continue;
            push(rec[i]);
        }
        if (has_ccp)
            for (int i = r0; i < r1 && !key_fault; i++) {
                if (rec[i].type != ZRK_MSG_LAUNCHED_MISSILE) continue;
                const int m_id = rec[i].a, tid = rec[i].b, ms = slot_of_id(m_id);
                size_t e = 0;
                while (e < me_key.size() && me_key[e] != m_id) e++;
                if (e == me_key.size()) { me_key.push_back(m_id); me_ref.push_back(ms); } else { me_ref[e] = ms; }
                if (ms >= 0) own[ms] = TEAM + (int)e;
                size_t q = 0;
                while (q < te_key.size() && te_key[q] != tid) q++;
                if (q == te_key.size()) { key_fault = true; break; }  // KeyError (CCP.py:107): the run ends here
                te_mid[q] = m_id;
            }
        if (has_ccp && !key_fault && !(fault == ZRK_FAULT_KEY_ERROR_CCP107 && t == fault_tick)) {
            // positions of this tick: Trajectory.get_pos + the noise of every radar that saw the object, in radar order
            auto pos_of = [&](int k, double *p) {
                const double dtm = now_s - tzt[2 * k + 1];
                p[0] = txy[2 * k] + vxy[2 * k] * dtm;
                p[1] = txy[2 * k + 1] + vxy[2 * k + 1] * dtm;
                p[2] = tzt[2 * k] + qzw[2 * k + 1] * dtm;
                if (!noise.empty()) {
                    const int st = sflags[k] >> F_STATIC_SHIFT;
                    for (int r = 0; r < nR; r++)
                        if (bit(m + (1 + r) * W, k)) {
                            const double *q = noise.data() + (size_t)(t % P.noise_period) * nrow + ((size_t)r * P.noise_objects + st) * 3;
                            p[0] += q[0]; p[1] += q[1]; p[2] += q[2];
                        }
                }
            };
            for (int k = 0; k < n_slots; k++) {
                rf[k] = -1;
                for (int r = 0; r < nR && rf[k] < 0; r++)
                    if (bit(m + (1 + r) * W, k)) rf[k] = r;
            }
            // detections in processing order: FOUND_OBJECTS messages in radar order, each object once (CCP.py:406-429)
            for (int r = 0; r < nR; r++)
                for (int k = 0; k < n_slots; k++) {
                    if (rf[k] != r) continue;
                    const zrk_msg *trk = nullptr, *cmd = nullptr;
                    for (int i = r0; i < r1; i++) {
                        if (rec[i].type == ZRK_REC_TRACK && rec[i].sender == k) trk = &rec[i];
                        if (rec[i].type == ZRK_MSG_LAUNCH_COMMAND && rec[i].a == sid[k]) cmd = &rec[i];
                    }
                    int cls, e;
                    if (trk) {
                        cls = trk->flag;
                        e = trk->receiver;
                    } else {  // the detection re-acquired the track of its own slot
                        cls = own[k] >= TEAM ? 2 : 1;
                        e = own[k] >= TEAM ? own[k] - TEAM : own[k];
                        if (own[k] < 0) { h->err = "rebuild_messages: a detection without a track and without a journal entry"; return ZRK_ERR_ARG; }
                    }
                    if (cls == 1) {
                        const bool following = te_follow[e] != 0;
                        if (cmd) {
                            push(*cmd);
                            te_follow[e] = 1;
                        } else if (following) {
                            zrk_msg x{};
                            x.tick = t; x.type = ZRK_MSG_CCP_UPDATE_TARGET; x.relevance = 1; x.sender = ccp_id; x.receiver = c_rad_id[r];
                            x.a = sid[k]; x.b = te_mid[e];
                            push(x);
                        }
                        te_ref[e] = k;
                        own[k] = e;
                    } else if (cls == 2) {
                        me_ref[e] = k;
                        own[k] = TEAM + e;
                    } else {  // new_target + add_target (CCP.py:322-330, :88-93)
                        if (cmd) push(*cmd);
                        if ((size_t)e >= te_key.size()) { te_key.resize(e + 1, 0); te_ref.resize(e + 1, -1); te_follow.resize(e + 1, 0); te_mid.resize(e + 1, ZRK_NONE); }
                        te_key[e] = sid[k]; te_ref[e] = k; te_follow[e] = cmd ? 1 : 0; te_mid[e] = ZRK_NONE;
                        own[k] = e;
                    }
                }
            // send_objects_to_GUI (CCP.py:237-284)
            auto draw = [&](int k, int type, int vis) {
                zrk_msg x{};
                x.tick = t; x.type = ZRK_MSG_DRAW_OBJECTS; x.relevance = 1; x.flag = (uint8_t)vis; x.sender = ccp_id; x.receiver = ZRK_MANAGER_ID;
                x.a = sid[k]; x.b = type;
                pos_of(k, x.xyz);
                push(x);
            };
            bool attr_fault = false;
            for (size_t e = 0; e < me_key.size(); e++)
                if (me_ref[e] >= 0 && bit(m, me_ref[e]) && rf[me_ref[e]] >= 0) draw(me_ref[e], ZRK_ZUR, 1);
            for (size_t e = 0; e < te_key.size() && !attr_fault; e++) {
                const int k = te_ref[e];
                if (k < 0 || !bit(m, k) || rf[k] < 0) continue;
                if (sflags[k] & F_MISSILE) { attr_fault = true; break; }  // 'Missile' object has no attribute 'type' (CCP.py:259)
                draw(k, (sflags[k] >> F_TTYPE_SHIFT) & 3, 1);
            }
            if (!attr_fault && nR > 0)
                for (int k = 0; k < n_slots; k++)
                    if (bit(m, k) && rf[k] < 0) draw(k, (sflags[k] & F_MISSILE) ? ZRK_ZUR : ((sflags[k] >> F_TTYPE_SHIFT) & 3), 0);
        }
        // Manager.add_message keeps each tick stably sorted by -relevance (Manager.py:56-58)
        if (emit)
            std::stable_sort(out.begin() + first, out.end(), [](const zrk_msg &a, const zrk_msg &b) { return a.relevance > b.relevance; });
        det_prev = det_now;
    }
    return ZRK_OK;
}

extern "C" int zrk_read_messages(zrk_handle *h, int sc, int tick0, int tick1, zrk_msg *msgs, size_t cap, size_t *n_out,
                                 int32_t *ids, size_t ids_cap, size_t *n_ids_out) {
    if (!h || sc < 0 || sc >= h->P.lo.B || tick0 < 0 || tick1 < tick0) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    const Params &P = h->P;
    const Layout &o = P.lo;
    if (!(P.log_types & ZRK_LOG_MASKS) || (P.log_types & ZRK_LOG_EVENTS) != (unsigned)ZRK_LOG_EVENTS) {
        h->err = "zrk_read_messages needs a handle created with ZRK_LOG_EVENTS | ZRK_LOG_MASKS (or ZRK_LOG_FULL)";
        return ZRK_ERR_UNSUPPORTED;
    }
    if ((P.log_types & ZRK_LOG_BULK) != (unsigned)ZRK_LOG_BULK) {  // bulk types are not stored: rebuild them
        std::vector<zrk_msg> out;
        std::vector<int32_t> pool;
        int rc = rebuild_messages(h, sc, tick0, tick1, out, pool);
        if (rc) return rc;
        if (n_out) *n_out = out.size();
        if (n_ids_out) *n_ids_out = pool.size();
        if (out.size() > cap || pool.size() > ids_cap) return ZRK_ERR_CAPACITY;
        if (msgs && !out.empty()) memcpy(msgs, out.data(), out.size() * sizeof(zrk_msg));
        if (ids && !pool.empty()) memcpy(ids, pool.data(), pool.size() * sizeof(int32_t));
        return ZRK_OK;
    }
    std::vector<unsigned char> dyn;
    int rc = fetch_dyn(h, sc, sc + 1, dyn);
    if (rc) return rc;
    const int *hdr = (const int *)(dyn.data() + o.hot.hdr);
    const int *sid = (const int *)(dyn.data() + o.hot.sid);
    const int done = hdr[H_TICK] + ((hdr[H_FAULT] & 63) ? 1 : 0);  // the faulting tick's messages exist too
    if (hdr[H_FAULT] & ZRK_FAULT_LOG_OVERFLOW) {
        h->err = "record log overflow: raise zrk_options.log_capacity";
        return ZRK_ERR_CAPACITY;
    }
    if (tick1 > done) tick1 = done;
    if (done - tick0 > P.mask_ticks && tick0 < tick1) {
        h->err = "mask history ring is shorter than the requested range: raise zrk_options.mask_ticks";
        return ZRK_ERR_CAPACITY;
    }
    const int nrec = std::min(hdr[H_NREC], P.log_cap);
    std::vector<zrk_msg> rec(nrec ? nrec : 1);
    if (nrec) CK(cudaMemcpy(rec.data(), P.rec + (size_t)sc * P.log_cap, (size_t)nrec * sizeof(zrk_msg), cudaMemcpyDeviceToHost));
    const int nR = h->h_n_radars[sc], W = P.mask_words, rows = (1 + o.RP) * W;
    std::vector<unsigned> mask(rows);
    std::vector<zrk_msg> out;
    std::vector<int32_t> pool;
    int ri = 0;
    while (ri < nrec && rec[ri].tick < tick0) ri++;
    int det_prev = 0;
    if (tick0 > 0) {  // DESTROYED_MISSILE of tick0 answers the detonations of tick0 - 1
        for (int j = 0; j < nrec && rec[j].tick < tick0; j++)
            if (rec[j].tick == tick0 - 1 && rec[j].type == ZRK_MSG_MISSILE_DETONATE) det_prev++;
    }
    auto list_msg = [&](int tick, int type, int sender, int receiver, const unsigned *bits) {
        zrk_msg m{};
        m.tick = tick;
        m.type = (uint8_t)type;
        m.relevance = 1;
        m.sender = sender;
        m.receiver = receiver;
        m.a = (int32_t)pool.size();
        int cnt = 0;
        for (int k = 0; k < o.NS; k++)
            if (bits[k >> 5] >> (k & 31) & 1u) {
                pool.push_back(sid[k]);
                cnt++;
            }
        m.b = cnt;
        out.push_back(m);
    };
    for (int t = tick0; t < tick1; t++) {
        size_t first = out.size();
        CK(cudaMemcpy(mask.data(), P.masks + ((size_t)(t % P.mask_ticks) * o.B + sc) * rows, rows * sizeof(unsigned),
                      cudaMemcpyDeviceToHost));
        int det_now = 0;
        // AirEnv.step: MISSILE_POS / MISSILE_DETONATE in slot order, then ACTIVE_OBJECTS
        while (ri < nrec && rec[ri].tick == t &&
               (rec[ri].type == ZRK_MSG_MISSILE_POS || rec[ri].type == ZRK_MSG_MISSILE_DETONATE)) {
            det_now += rec[ri].type == ZRK_MSG_MISSILE_DETONATE;
            out.push_back(rec[ri++]);
        }
        list_msg(t, ZRK_MSG_ACTIVE_OBJECTS, h->h_airenv_id[sc], ZRK_NONE, mask.data());
        // every radar: ALL_OBJECTS, FOUND_OBJECTS, then its DESTROYED_MISSILE records
        for (int r = 0; r < nR; r++) {
            int rid = h->h_radar_id[(size_t)sc * o.R + r];
            list_msg(t, ZRK_MSG_ALL_OBJECTS, rid, ZRK_CCP_ID, mask.data());
            list_msg(t, ZRK_MSG_FOUND_OBJECTS, rid, ZRK_CCP_ID, mask.data() + (1 + r) * W);
            for (int k = 0; k < det_prev && ri < nrec && rec[ri].tick == t && rec[ri].type == ZRK_MSG_DESTROYED_MISSILE; k++)
                out.push_back(rec[ri++]);
        }
        // launchers and CCP, in emission order (the private track journal is not a message)
        for (; ri < nrec && rec[ri].tick == t; ri++)
            if (rec[ri].type != ZRK_REC_TRACK) out.push_back(rec[ri]);
        // Manager.add_message keeps each tick stably sorted by -relevance (Manager.py:56-58)
        std::stable_sort(out.begin() + first, out.end(),
                         [](const zrk_msg &a, const zrk_msg &b) { return a.relevance > b.relevance; });
        det_prev = det_now;
    }
    if (n_out) *n_out = out.size();
    if (n_ids_out) *n_ids_out = pool.size();
    if (out.size() > cap || pool.size() > ids_cap) return ZRK_ERR_CAPACITY;
    if (msgs && !out.empty()) memcpy(msgs, out.data(), out.size() * sizeof(zrk_msg));
    if (ids && !pool.empty()) memcpy(ids, pool.data(), pool.size() * sizeof(int32_t));
    return ZRK_OK;
}

extern "C" int zrk_read_state(zrk_handle *h, const zrk_state_view *v) {
    if (!h || !v) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    const Params &P = h->P;
    const Layout &o = P.lo;
    const int NS = o.NS, T = o.T, M = o.M, R = o.R, L = o.L, C = o.C;
    const HotOff &ho = o.hot;
    const int CH = 2048;
    std::vector<unsigned char> dyn, cst;
    for (int s0 = 0; s0 < o.B; s0 += CH) {
        int s1 = std::min(o.B, s0 + CH), n = s1 - s0;
        int rc = fetch_dyn(h, s0, s1, dyn);
        if (rc) return rc;
        cst.resize((size_t)n * o.cst_bytes);
        CK(cudaMemcpy(cst.data(), P.cst + (size_t)s0 * o.cst_bytes, cst.size(), cudaMemcpyDeviceToHost));
        for (int s = s0; s < s1; s++) {
            const unsigned char *sm = dyn.data() + (size_t)(s - s0) * o.dyn_bytes;
            const unsigned char *cs = cst.data() + (size_t)(s - s0) * o.cst_bytes;
            const double *tzt = (const double *)(sm + ho.tzt), *vxy = (const double *)(sm + ho.vxy);
            const int *hdr = (const int *)(sm + ho.hdr), *sflags = (const int *)(sm + ho.sflags);
            const double *pxy = (const double *)(sm + ho.pxy), *pzs = (const double *)(sm + ho.pzs);
            const double *qxy = (const double *)(sm + ho.qxy), *qzw = (const double *)(sm + ho.qzw);
            const int *chdr = (const int *)(cs + o.c_hdr);
            const int nM = chdr[C_NM], nR = chdr[C_NR], nL = chdr[C_NL], nC = chdr[C_NC];
            const int n_slots = hdr[H_NSLOTS];
            auto stat = [&](int slot) { return slot >= 0 && slot < n_slots ? (sflags[slot] >> F_STATIC_SHIFT) : -1; };
            // objects without a slot yet: missiles still on a launcher sit at the launcher position
            for (int i = 0; i < NS; i++) {
                size_t g = (size_t)s * NS + i;
                if (v->slot_obj) v->slot_obj[g] = -1;
                if (v->slot_alive) v->slot_alive[g] = 0;
                if (v->has_prev) v->has_prev[g] = 0;
                for (int k = 0; k < 3; k++) {
                    if (v->pos) v->pos[3 * g + k] = 0.0;
                    if (v->prev_pos) v->prev_pos[3 * g + k] = 0.0;
                }
            }
            const int *c_m_launcher = (const int *)(cs + o.c_m_launcher);
            const double *c_l_pos = (const double *)(cs + o.c_l_pos);
            for (int m = 0; m < nM; m++)
                for (int k = 0; k < 3; k++)
                    if (v->pos) v->pos[3 * ((size_t)s * NS + T + m) + k] = c_l_pos[3 * c_m_launcher[m] + k];
            for (int k = 0; k < n_slots; k++) {
                int si = sflags[k] >> F_STATIC_SHIFT;
                size_t gs = (size_t)s * NS + k, go = (size_t)s * NS + si;
                // a slot is part of AirEnv.__objects once its missile was appended (not while pending)
                bool listed = !(sflags[k] & F_PENDING);
                if (v->slot_obj) v->slot_obj[gs] = listed ? si : -1;
                if (v->slot_alive) v->slot_alive[gs] = (sflags[k] & F_ALIVE) ? 1 : 0;
                if (v->has_prev) v->has_prev[go] = (sflags[k] & F_HASPREV) ? 1 : 0;
                const double pk[3] = {pxy[2 * k], pxy[2 * k + 1], pzs[2 * k]};
                const double qk[3] = {qxy[2 * k], qxy[2 * k + 1], qzw[2 * k]};
                for (int c = 0; c < 3; c++) {
                    if (v->pos) v->pos[3 * go + c] = pk[c];
                    if (v->prev_pos) v->prev_pos[3 * go + c] = (sflags[k] & F_HASPREV) ? qk[c] : 0.0;
                }
            }
            const double *rad_cur = (const double *)(sm + o.rad_cur);
            for (int r = 0; r < R; r++) {
                if (v->radar_az) v->radar_az[(size_t)s * R + r] = r < nR ? rad_cur[2 * r] : 0.0;
                if (v->radar_el) v->radar_el[(size_t)s * R + r] = r < nR ? rad_cur[2 * r + 1] : 0.0;
            }
            const int *m_status = (const int *)(sm + o.m_status), *m_target = (const int *)(sm + o.m_target),
                      *m_slot = (const int *)(sm + o.m_slot);
            const double *m_period = (const double *)(sm + o.m_period);
            for (int m = 0; m < M; m++) {
                size_t g = (size_t)s * M + m;
                bool live = m < nM;
                int k = live ? m_slot[m] : -1;
                if (v->missile_status) v->missile_status[g] = live ? (uint8_t)m_status[m] : 0;
                if (v->missile_period) v->missile_period[g] = live ? m_period[m] : 0.0;
                if (v->missile_t0) v->missile_t0[g] = k >= 0 ? tzt[2 * k + 1] : 0.0;
                if (v->missile_target) v->missile_target[g] = live && m_target[m] >= 0 ? stat(m_target[m]) : -1;
                const double vk[3] = {k >= 0 ? vxy[2 * k] : 0.0, k >= 0 ? vxy[2 * k + 1] : 0.0, k >= 0 ? qzw[2 * k + 1] : 0.0};
                for (int c = 0; c < 3; c++)
                    if (v->missile_vel) v->missile_vel[3 * g + c] = vk[c];
            }
            const int *st_n = (const int *)(sm + o.st_n);
            for (int l = 0; l < L; l++)
                if (v->launcher_count) v->launcher_count[(size_t)s * L + l] = l < nL ? st_n[l] : 0;
            const int n_te = hdr[H_NTE], n_me = hdr[H_NME];
            if (v->ccp_n_targets) v->ccp_n_targets[s] = n_te;
            if (v->ccp_n_missiles) v->ccp_n_missiles[s] = n_me;
            const int *te_key = (const int *)(sm + ho.te_key), *te_ref = (const int *)(sm + ho.te_ref),
                      *te_upd = (const int *)(sm + ho.te_upd), *te_follow = (const int *)(sm + ho.te_follow),
                      *te_mid = (const int *)(sm + ho.te_mid);
            const int *me_key = (const int *)(sm + o.me_key), *me_ref = (const int *)(sm + o.me_ref),
                      *me_upd = (const int *)(sm + o.me_upd), *me_tid = (const int *)(sm + o.me_tid);
            for (int e = 0; e < NS; e++) {
                size_t g = (size_t)s * NS + e;
                bool lt = e < n_te, lm = e < n_me;
                if (v->ccp_t_key) v->ccp_t_key[g] = lt ? te_key[e] : 0;
                if (v->ccp_t_ref) v->ccp_t_ref[g] = lt ? stat(te_ref[e]) : -1;
                if (v->ccp_t_upd) v->ccp_t_upd[g] = lt ? te_upd[e] : 0;
                if (v->ccp_t_missile) v->ccp_t_missile[g] = lt ? te_mid[e] : ZRK_NONE;
                if (v->ccp_t_following) v->ccp_t_following[g] = lt ? (uint8_t)te_follow[e] : 0;
                if (v->ccp_m_key) v->ccp_m_key[g] = lm ? me_key[e] : 0;
                if (v->ccp_m_ref) v->ccp_m_ref[g] = lm ? stat(me_ref[e]) : -1;
                if (v->ccp_m_upd) v->ccp_m_upd[g] = lm ? me_upd[e] : 0;
                if (v->ccp_m_target) v->ccp_m_target[g] = lm ? me_tid[e] : ZRK_NONE;
            }
            const int *capv = (const int *)(sm + o.cap), *lnc = (const int *)(sm + o.launched);
            for (int c = 0; c < C; c++) {
                if (v->ccp_capacity) v->ccp_capacity[(size_t)s * C + c] = c < nC ? capv[c] : 0;
                if (v->ccp_launched) v->ccp_launched[(size_t)s * C + c] = c < nC ? lnc[c] : 0;
            }
        }
    }
    return ZRK_OK;
}

// ------------------------------------------------------------------ checkpoint / resume

struct SaveHeader {
    unsigned long long magic;
    long long tick;
    unsigned long long dyn_bytes, rec_bytes, mask_bytes;
    int B, team, log_cap, mask_ticks;
    int abi_version, reserved;
    unsigned long long layout_hash;  // FNV-1a of the Layout struct: every offset of the state blob
    unsigned long long cst_hash;     // FNV-1a of the parameter blobs (radars, launchers, missiles, CCP) + dt + noise shape
};

static const unsigned long long ZRK_SAVE_MAGIC = 0x5a524b5f53415645ull;  // "ZRK_SAVE"

static void save_sizes(const zrk_handle *h, SaveHeader &sh) {
    const Params &P = h->P;
    const Layout &o = P.lo;
    sh.magic = ZRK_SAVE_MAGIC;
    sh.tick = h->tick;
    sh.dyn_bytes = (unsigned long long)o.B * o.dyn_bytes;
    sh.rec_bytes = (unsigned long long)o.B * P.log_cap * sizeof(zrk_msg);
    sh.mask_bytes = P.masks ? (unsigned long long)P.mask_ticks * o.B * (1 + o.RP) * P.mask_words * sizeof(unsigned) : 0;
    sh.B = o.B;
    sh.team = o.team;
    sh.log_cap = P.log_cap;
    sh.mask_ticks = P.mask_ticks;
    sh.abi_version = ZRK_ABI_VERSION;
    sh.reserved = 0;
    sh.layout_hash = fnv1a(&o, sizeof(Layout));
    sh.cst_hash = h->cst_hash;
}
// identity of the batch for checkpoints: the parameter blobs (radars, launchers, missiles, CCP), the time step and the
// shape of the noise table.  Computed from the device copy the first time a checkpoint is written or read.
static int ensure_cst_hash(zrk_handle *h) {
    if (h->cst_hash) return ZRK_OK;
    int rc = zrk_sync(h);
    if (rc) return rc;
    const Layout &o = h->P.lo;
    std::vector<unsigned char> cst((size_t)o.B * o.cst_bytes);
    CK(cudaMemcpy(cst.data(), h->P.cst, cst.size(), cudaMemcpyDeviceToHost));
    unsigned long long hh = fnv1a(&h->P.dt_ms, sizeof h->P.dt_ms);
    const int nshape[4] = {h->P.noise ? h->P.noise_period : 0, h->P.noise_radars, h->P.noise_objects, (int)(h->P.noise_stride != 0)};
    hh = fnv1a(nshape, sizeof nshape, hh);
    hh = fnv1a(cst.data(), cst.size(), hh);
    h->cst_hash = hh ? hh : 1;
    return ZRK_OK;
}

extern "C" int zrk_state_bytes(zrk_handle *h, size_t *bytes) {
    if (!h || !bytes) return ZRK_ERR_ARG;
    SaveHeader sh;
    save_sizes(h, sh);
    *bytes = sizeof(SaveHeader) + sh.dyn_bytes + sh.rec_bytes + sh.mask_bytes;
    return ZRK_OK;
}

extern "C" int zrk_save_state(zrk_handle *h, void *buf, size_t cap) {
    if (!h || !buf) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    {
        int rc0 = ensure_cst_hash(h);
        if (rc0) return rc0;
    }
    SaveHeader sh;
    save_sizes(h, sh);
    if (cap < sizeof(SaveHeader) + sh.dyn_bytes + sh.rec_bytes + sh.mask_bytes) return ZRK_ERR_CAPACITY;
    int rc = zrk_sync(h);
    if (rc) return rc;
    unsigned char *p = (unsigned char *)buf;
    memcpy(p, &sh, sizeof sh);
    p += sizeof sh;
    CK(cudaMemcpy(p, h->P.dyn, sh.dyn_bytes, cudaMemcpyDeviceToHost));
    p += sh.dyn_bytes;
    CK(cudaMemcpy(p, h->P.rec, sh.rec_bytes, cudaMemcpyDeviceToHost));
    p += sh.rec_bytes;
    if (sh.mask_bytes) CK(cudaMemcpy(p, h->P.masks, sh.mask_bytes, cudaMemcpyDeviceToHost));
    return ZRK_OK;
}

extern "C" int zrk_load_state(zrk_handle *h, const void *buf, size_t bytes) {
    if (!h || !buf || bytes < sizeof(SaveHeader)) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    {
        int rc0 = ensure_cst_hash(h);
        if (rc0) return rc0;
    }
    SaveHeader sh, mine;
    memcpy(&sh, buf, sizeof sh);
    save_sizes(h, mine);
    if (sh.magic == ZRK_SAVE_MAGIC && (sh.abi_version != mine.abi_version || sh.layout_hash != mine.layout_hash)) {
        h->err = "checkpoint was written by a different build of the library (ABI version / state layout)";
        return ZRK_ERR_ARG;
    }
    if (sh.magic == ZRK_SAVE_MAGIC && sh.cst_hash != mine.cst_hash) {
        h->err = "checkpoint belongs to a batch with different scenario parameters";
        return ZRK_ERR_ARG;
    }
    if (sh.magic != ZRK_SAVE_MAGIC || sh.B != mine.B || sh.team != mine.team || sh.log_cap != mine.log_cap ||
        sh.mask_ticks != mine.mask_ticks || sh.dyn_bytes != mine.dyn_bytes || sh.rec_bytes != mine.rec_bytes ||
        sh.mask_bytes != mine.mask_bytes || bytes < sizeof(SaveHeader) + sh.dyn_bytes + sh.rec_bytes + sh.mask_bytes) {
        h->err = "checkpoint does not match this handle (different batch shape or options)";
        return ZRK_ERR_ARG;
    }
    int rc = zrk_sync(h);
    if (rc) return rc;
    const unsigned char *p = (const unsigned char *)buf + sizeof sh;
    CK(cudaMemcpy(h->P.dyn, p, sh.dyn_bytes, cudaMemcpyHostToDevice));
    p += sh.dyn_bytes;
    CK(cudaMemcpy(h->P.rec, p, sh.rec_bytes, cudaMemcpyHostToDevice));
    p += sh.rec_bytes;
    if (sh.mask_bytes) CK(cudaMemcpy(h->P.masks, p, sh.mask_bytes, cudaMemcpyHostToDevice));
    h->tick = sh.tick;
    h->round_no = 0;
    if (h->d_pending) CK(cudaMemset(h->d_pending, 0, sizeof(int) * ZRK_ROUND_SLOTS * (h->wave_stream.empty() ? 1 : h->wave_stream.size())));
    if (h->d_queue) CK(cudaMemset(h->d_queue, 0, sizeof(int) * (ZRK_ROUND_SLOTS + 64 * ZRK_FORK_ROUNDS_MAX)));
    return ensure_tsec(h, h->tick + 2);
}

extern "C" int zrk_device_ptr(zrk_handle *h, const char *name, void **ptr, size_t *bytes) {
    if (!h || !name || !ptr) return ZRK_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    const Params &P = h->P;
    const Layout &o = P.lo;
    std::string n(name);
    size_t sz = 0;
    void *p = nullptr;
    if (n == "dyn") { p = P.dyn; sz = (size_t)o.B * o.dyn_bytes; }
    else if (n == "cst") { p = (void *)P.cst; sz = (size_t)o.B * o.cst_bytes; }
    else if (n == "rec") { p = P.rec; sz = (size_t)o.B * P.log_cap * sizeof(zrk_msg); }
    else if (n == "summary") {  // [B][H_WORDS + NTYPES] int32: header + message counters, refreshed here
        int rc = fetch_summary(h);
        if (rc) return rc;
        p = h->d_summary;
        sz = (size_t)o.B * ZRK_SUMMARY_WORDS * sizeof(int);
    }
    else if (n == "masks") { p = P.masks; sz = (size_t)P.mask_ticks * o.B * (1 + o.RP) * P.mask_words * sizeof(unsigned); }
    else return ZRK_ERR_ARG;
    *ptr = p;
    if (bytes) *bytes = sz;
    return ZRK_OK;
}
