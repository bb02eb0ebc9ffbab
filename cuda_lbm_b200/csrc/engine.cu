// engine.cu — host side of the B200-native D2Q9 engine and its C ABI (include/lbm_b200.h).
//
// Replaces the reference's LBM<2> host methods (src/core/lbm.cuh:33-382) and IBMManager<2>
// (src/IBM/IBMManager.cuh:31-253).  Data layout in HBM (per slab of ny_local rows):
//   populations  9 planes x (ny_local+2) rows x nx fp32   (SoA, rows 0 / ny_local+1 = slab ghost rows)
//   rest plane   +1 plane when LBM_QK_D1_STALE_F0 (the reference's two interleaved f0 histories)
//   flags        1 byte / node (only when a scenario has non-FLUID nodes or bodies)
//   edge ring    2 x (2nx+2ny) x 9 fp32: post-collision values of domain-edge nodes (undelivered slots)
//   macroscopics rho (1 plane) + u (AoS float2), allocated on first request
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/sort.h>
#include <thrust/unique.h>
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/sequence.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include <unistd.h>
#include <sched.h>
#include <cctype>

#include "../../include/lbm_b200.h"
#include "kernels.cuh"

using namespace lbm;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(LBM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

struct lbm_handle {
    lbm_config cfg{};
    int y0 = 0, nyl = 0;
    long long nloc = 0;
    size_t plane = 0;               // floats per slot plane
    cudaStream_t stream = nullptr, own_stream = nullptr;
    // the general-path work of a step (IBM pre-pass, boundary / body segments) runs beside the vectorised kernel on a second
    // stream: under the AA pattern a cell reads exactly the slots it overwrites, so the two kernels touch disjoint memory
    cudaStream_t side_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool overlap = true;
    // small grids are bound by the host's launch rate, not by the GPU: lbm_step(h, n) replays a captured CUDA graph of
    // 2*GRAPH_PAIRS steps (an odd/even pair repeats identically: only the parity of t reaches the kernels)
    struct StepGraph { cudaGraphExec_t exec = nullptr; std::string key; long long launches = 0; int d_avg = 0, d_pre = 0, d_nbrg = 0; } graph[2];
    int graph_mode = -1;            // LBM_B200_GRAPH: 0 never, 1 whenever possible, unset = slabs of up to 2^22 cells
    cudaEvent_t ev_bridge[2] = {nullptr, nullptr};
    // lbm_set_lookahead: lbm_step(h, n) on a peer-mapped slab keeps at most 3 x 4 steps enqueued ahead of the device (events recorded
    // every 4 steps, the host waits for the one of 12 steps ago).  For callers that drive every slab from its own thread: when several
    // slabs share ONE device (more slabs than GPUs) their blocked launches otherwise exhaust the context's launch queue while the slab
    // they wait for cannot enqueue any more (seen with 4 slabs x 60 steps on one B200).  Off by default: a caller that enqueues the
    // slabs one after the other from a single thread needs the unbounded queue.
    cudaEvent_t ev_throttle[3] = {nullptr, nullptr, nullptr};
    bool bounded_lookahead = false;
    // lbm_run_from_host: copy streams and per-band events of the time-skewed pipeline (engine_pipeline.inc)
    cudaStream_t copy_in = nullptr, copy_out = nullptr; cudaEvent_t ev_pipe = nullptr; std::vector<cudaEvent_t> ev_band;
    unsigned long long pipe_epoch = 0;           // number of pipelined calls on several slabs (tags the level counters)      // legacy / per-thread user stream <-> own stream around graph replays
    float* pop = nullptr;           // 9 (+1) planes
    int nplanes = 9;
    uint8_t* flags = nullptr;
    float2* force_plane = nullptr;
    float* ring = nullptr; int perim = 0;
    long long* nbr_nodes = nullptr; long long* nbr_src = nullptr; float* nbr_g = nullptr; int nbr_count = 0;
    // IBM
    std::vector<float> h_pts;
    std::vector<float> h_vel; bool has_vel = false;        // IBMBody::velocities, parallel to h_pts (lbm_set_body_velocities)
    float2* d_utarget = nullptr;
    float* d_pts = nullptr; long long* ibm_nodes = nullptr; int* sten_idx = nullptr; float* sten_w = nullptr;
    int* csr_row = nullptr; int* csr_k = nullptr; float* csr_w = nullptr;
    float* ibm_rho = nullptr; float2* ibm_uprev = nullptr; float2* ibm_lagF = nullptr; float2* ibm_force = nullptr;
    int ibm_one_block_max = 8192;                // more markers / stencil nodes than this: seven many-block launches instead of one block (LBM_B200_IBM_ONE_BLOCK_MAX)
    int np = 0, ibm_count = 0, ibm_ss = 4;       // markers / stencil nodes this slab works on (world > 1: the bodies it owns a node of)
    // bodies across slab faces: every slab knows all bodies; node states travel through a mailbox indexed by the global node list
    std::vector<int> body_start;                 // first marker of each body in h_pts
    int np_total = 0;                            // markers of all bodies
    long long nall = 0;                          // stencil nodes of all bodies (mailbox slots in use)
    int mail_nodes = 0;                          // mailbox capacity in nodes (world > 1)
    float* ibm_mail = nullptr;                   // inside the `pop` allocation, so the neighbours reach it through the same mapping
    int* ibm_mail_idx = nullptr;
    int ibm_mail_for_ts = -1;                    // halo-API coupling: the mailbox holds the all-reduced node states of this step
    int nbrg_for_ts = -1;
    std::vector<int> ibm_rows;                   // distinct rows of this slab's active stencil nodes (peer-coverage check)
    // general-path segments (kernels.cuh): mask per 128-cell segment + compact list, rebuilt lazily
    uint8_t* segmask = nullptr; int* gen_list = nullptr; int gen_count = 0; int nsx = 0; bool segs_dirty = true;
    long long* gen_cells = nullptr; long long gen_cell_count = 0, gen_cell_cap = 0;      // general cells of the mixed segments (local node ids)
    uint8_t colclass[COLCLASS_MAX] = {};      // Params::colclass, valid for the rows pure[0] .. pure[1]
    int pure[4] = {0, 0, 0, 0};     // rows [0],[1]) x segments [2],[3]): the largest rectangle of all-vector segments (Params::pure_*)
    // adapter
    float* partials = nullptr; long long n_partials = 0; double* stage = nullptr; double* sums = nullptr; float* avg = nullptr; int avg_for_ts = -1; int pre_for_ts = -1;
    // macroscopics
    float* rho_out = nullptr; float2* u_out = nullptr; int macros_ts = -1;
    double* mass_acc = nullptr;
    double* val_stage = nullptr; long long val_stage_n = 0;      // validation reductions (lbm_*_error_sums, lbm_row_mean_velocity)
    bool odd_interleaved = true;    // odd phase with step_odd_kernel where rows are whole segments (LBM_B200_ODD=0: the shuffle-based kernel everywhere)
    int timestep = 0;
    long long launches = 0;
    long long bytes = 0;
    float omega = 1.0f;
    // peer-mapped slab coupling (lbm_peer_export / lbm_peer_attach)
    unsigned long long* sync_flags = nullptr;      // 2 step counters written by the neighbours + padding, at the end of `pop`
    int* sync_timeout = nullptr;
    struct Peer { float* base = nullptr; long long plane = 0, off = 0; unsigned long long* flag = nullptr; bool attached = false; int rank = -1; } peer[2];
    bool direct() const { return peer[0].attached || peer[1].attached; }
    // every slab of the decomposition this handle has mapped (lbm_peer_attach_all maps all of them, lbm_peer_attach its neighbour):
    // the device-side all-reduce of the adapter sums and the IBM node states of bodies across slab faces reach all of these
    struct Mapped { char* base = nullptr; void* ipc_base = nullptr; long long flags_off = 0, mail_off = -1; int y0 = 0, nyl = 0; } mapped[MAX_WORLD];
    SlabNet* d_net = nullptr;                      // device copy of the table the kernels use
    bool attached_all = false;                     // lbm_peer_attach_all was called (by contract on EVERY slab): adapter sums are all-reduced on the device
    int adp_published_ts = -1;                     // lagged OptimalAdapter on several slabs: sums for this step are on their way to every slab's mailbox
    unsigned long long ibm_need_mask = 0;          // ranks that own stencil nodes of the bodies this slab works on
};

// what one slab tells its neighbours (lbm_peer_export); opaque to callers, LBM_PEER_DESC_BYTES long
struct PeerDesc {
    cudaIpcMemHandle_t ipc;         // of the population allocation
    long long pid;
    unsigned long long raw;         // device pointer, valid inside the exporting process
    long long plane;                // floats per slot plane
    long long flags_off;            // byte offset of the step / IBM stage counters inside the allocation
    long long mail_off;             // byte offset of the IBM mailbox inside the allocation (the IBM stage counters sit at flags_off + TAIL_STAGE_OFF,
                                    // the adapter mailbox at flags_off + TAIL_ADP_OFF)
    int nx, nyl, device, rank, y0, mail_nodes;
};
static_assert(sizeof(PeerDesc) <= LBM_PEER_DESC_BYTES, "PeerDesc must fit the ABI buffer");
// tail of the population allocation (reached by the other slabs through the same mapping): 32 step / stage / level counters,
// the per-source-rank IBM stage counters, the adapter mailbox, then the IBM node mailbox
constexpr size_t TAIL_STAGE_OFF = 256, TAIL_ADP_OFF = TAIL_STAGE_OFF + MAX_WORLD * 8, TAIL_FIXED_BYTES = TAIL_ADP_OFF + 2 * MAX_WORLD * sizeof(AdpSlot);

template <typename T>
static cudaError_t dmalloc(lbm_handle* h, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e == cudaSuccess) h->bytes += (long long)(count * sizeof(T));
    return e;
}

static Params make_params(lbm_handle* h, int t) {
    Params p{};
    for (int q = 0; q < Q; q++) p.A[q] = h->pop + (size_t)q * h->plane;
    p.A0[0] = p.A[0];
    p.A0[1] = (h->nplanes == 10) ? h->pop + (size_t)9 * h->plane : p.A[0];
    p.nx = h->cfg.nx; p.ny = h->cfg.ny; p.y0 = h->y0; p.nyl = h->nyl;
    p.px = h->cfg.periodic_x; p.py = h->cfg.periodic_y;
    p.wrap_y = (h->cfg.world == 1 && h->cfg.periodic_y) ? 1 : 0;
    p.t = t; p.quirks = h->cfg.quirks; p.coll = h->cfg.collision;
    p.flags = h->flags;
    p.omega = h->omega;
    for (int i = 0; i < Q; i++) p.S[i] = h->cfg.S[i];
    p.u_max = h->cfg.u_max; p.fx = h->cfg.force_x; p.fy = h->cfg.force_y;
    p.force_plane = h->force_plane;
    p.ring = h->ring; p.perim = h->perim;
    p.nbr_nodes = h->nbr_nodes; p.nbr_g = h->nbr_g; p.nbr_count = h->nbr_count;
    p.ibm_nodes = h->ibm_nodes; p.ibm_force = h->ibm_force; p.ibm_count = h->ibm_count;
    p.avg = h->avg; p.partials = nullptr; p.rho_out = nullptr; p.u_out = nullptr;
    p.pure_y0 = h->pure[0]; p.pure_y1 = h->pure[1]; p.pure_s0 = h->pure[2]; p.pure_s1 = h->pure[3];
    memcpy(p.colclass, h->colclass, sizeof(p.colclass));
    p.segmask = nullptr; p.nsx = h->nsx; p.gen_list = nullptr; p.gen_cells = nullptr; p.gen_cell_count = 0; p.plane = (long long)h->plane;
    for (int sd = 0; sd < 2; sd++) { p.peer[sd] = h->peer[sd].attached ? h->peer[sd].base : nullptr; p.peer_plane[sd] = h->peer[sd].plane; p.peer_off[sd] = h->peer[sd].off; }
    return p;
}

// CUDA loads kernels lazily, at their first launch, and that load waits for the device to drain.  A first launch issued while
// a handshake kernel of another handle in this process spins would therefore dead-lock until the handshake timeout
// (seen on a B200: two peer-mapped slabs in one process, 7 steps enqueued at once).  Every kernel a step can launch is
// loaded up front instead, once per device, before any handle exists.
template <typename K> static void preload(K kernel) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, kernel); }
template <int COLL> static void preload_coll() {
    preload(step_vec_kernel<COLL, false>); preload(step_vec_kernel<COLL, true>); preload(step_odd_kernel<COLL>);
    preload(step_kernel<COLL, false, false>); preload(step_kernel<COLL, false, true>);
    preload(step_kernel<COLL, true, false>); preload(step_kernel<COLL, true, true>);
}
static void preload_kernels(int device) {
    static bool done[64] = {};
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    if (device < 0 || device >= 64 || done[device]) return;
    done[device] = true;
    preload_coll<C_BGK>(); preload_coll<C_MRT>(); preload_coll<C_CM>(); preload_coll<C_CMOPT>();
    preload(moments_kernel<false>); preload(moments_kernel<true>); preload(moments_vec_kernel<false>); preload(moments_vec_kernel<true>);
    preload(reduce_kernel); preload(sums_to_avg_kernel); preload(adapter_collect_kernel);
    preload(recover_macros_kernel<false>); preload(recover_macros_kernel<true>);
    preload(nbr_gather_kernel<false>); preload(nbr_gather_kernel<true>);
    preload(ibm_kernel<false>); preload(ibm_kernel<true>); preload(ibm_state_kernel<false>); preload(ibm_state_kernel<true>);
    preload(ibm_markers_kernel); preload(ibm_nodes_kernel); preload(ibm_gather_kernel<false>); preload(ibm_gather_kernel<true>); preload(ibm_solve_kernel);
    preload(wait_neighbours_kernel); preload(signal_neighbours_kernel); preload(build_segmask_kernel);
    preload(init_fields_kernel); preload(init_taylor_green_kernel);
    cudaGetLastError();
}

static dim3 grid_of(const lbm_handle* h) { return dim3((h->cfg.nx + BX - 1) / BX, h->nyl); }

extern "C" const char* lbm_last_error(void) { return g_err.c_str(); }

extern "C" int lbm_default_config(lbm_config* c) {
    if (!c) return fail(LBM_ERR_INVALID, "cfg is NULL");
    memset(c, 0, sizeof(*c));
    c->nx = 128; c->ny = 128; c->collision = LBM_BGK;
    c->viscosity = 1.0f / 6.0f;                      // ScenarioTrait::viscosity, scenario.cuh:35
    float om = 1.0f / (3 * c->viscosity + 0.5f);
    float S[9] = {0.f, om, om, 0.f, om, 0.f, om, om, om};   // scenario.cuh:47-57
    memcpy(c->S, S, sizeof(S));
    c->u_max = 0.1f;                                 // scenario.cuh:59
    c->quirks = LBM_QK_REFERENCE; c->adapter_mode = LBM_ADAPTER_EXACT;
    c->world = 1;
    return LBM_OK;
}

extern "C" int lbm_destroy(lbm_handle* h) {
    if (!h) return LBM_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& m : h->mapped) if (m.ipc_base) cudaIpcCloseMemHandle(m.ipc_base);
    void* ptrs[] = {h->d_net, h->sync_timeout, h->pop, h->flags, h->force_plane, h->ring, h->nbr_nodes, h->nbr_src, h->nbr_g, h->d_pts, h->ibm_nodes,
                    h->sten_idx, h->sten_w, h->csr_row, h->csr_k, h->csr_w, h->ibm_rho, h->ibm_uprev, h->ibm_lagF, h->ibm_force, h->ibm_mail_idx, h->d_utarget,
                    h->partials, h->stage, h->sums, h->avg, h->rho_out, h->u_out, h->mass_acc, h->segmask, h->gen_list, h->gen_cells, h->val_stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (auto& g : h->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& e : h->ev_bridge) if (e) cudaEventDestroy(e);
    for (auto& e : h->ev_throttle) if (e) cudaEventDestroy(e);
    for (auto& e : h->ev_band) cudaEventDestroy(e);
    if (h->ev_pipe) cudaEventDestroy(h->ev_pipe);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return LBM_OK;
}

extern "C" int lbm_create(const lbm_config* cfg, lbm_handle** out) {
    if (!cfg || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->nx < 3 || cfg->ny < 3) return fail(LBM_ERR_INVALID, "grid must be at least 3x3");
    if (cfg->collision < LBM_BGK || cfg->collision > LBM_CM_OPTIMAL) return fail(LBM_ERR_INVALID, "unknown collision operator");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return fail(LBM_ERR_INVALID, "bad rank/world");
    if (cfg->world > MAX_WORLD) return fail(LBM_ERR_INVALID, "at most 64 slabs");
    if (cfg->ny / cfg->world < 2) return fail(LBM_ERR_INVALID, "each slab needs at least 2 rows");
    // host asserts of LBM::init (src/core/init/init.cuh:62-64) become error returns
    if (!(cfg->viscosity > 0.0f)) return fail(LBM_ERR_INVALID, "Negative Viscosity");
    if (!(3 * cfg->viscosity + 0.5f > 0.5f)) return fail(LBM_ERR_INVALID, "Instability warning: tau < 0.5");
    if (!(cfg->u_max < 0.5f)) return fail(LBM_ERR_INVALID, "Instability warning: u_max > 0.5");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(LBM_ERR_INVALID, "no such CUDA device");
    CU(cudaSetDevice(cfg->device));
    preload_kernels(cfg->device);
    lbm_handle* h = new lbm_handle();
    h->cfg = *cfg;
    const float tau = 3 * cfg->viscosity + 0.5f;     // viscosity_to_tau, lbm_constants.cuh:365-367
    h->omega = 1.0f / tau;
    // slab rows: the first (ny % world) slabs get one extra row
    int base = cfg->ny / cfg->world, rem = cfg->ny % cfg->world;
    h->nyl = base + (cfg->rank < rem ? 1 : 0);
    h->y0 = cfg->rank * base + std::min(cfg->rank, rem);
    h->nloc = (long long)h->nyl * cfg->nx;
    h->plane = (size_t)(h->nyl + 2) * cfg->nx;
    h->nplanes = (cfg->quirks & LBM_QK_D1_STALE_F0) ? 10 : 9;
    h->nsx = (cfg->nx + SEG - 1) / SEG;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(LBM_ERR_CUDA, cudaGetErrorString(e)); }
    h->stream = h->own_stream;
    {   // LBM_B200_OVERLAP=0 serialises the general-path kernels behind the vectorised one (diagnosis / A-B timing)
        const char* ov = getenv("LBM_B200_OVERLAP");
        h->overlap = !(ov && ov[0] == '0');
        if (h->overlap && (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
                           cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                           cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess)) {
            std::string m = cudaGetErrorString(cudaGetLastError()); lbm_destroy(h); return fail(LBM_ERR_CUDA, "stream / event creation failed: " + m);
        }
    }
    if (const char* gm = getenv("LBM_B200_GRAPH")) h->graph_mode = gm[0] == '0' ? 0 : 1;
    if (const char* v = getenv("LBM_B200_IBM_ONE_BLOCK_MAX")) h->ibm_one_block_max = atoi(v);
    h->perim = 2 * cfg->nx + 2 * cfg->ny;
    const size_t pop_floats = (h->plane * h->nplanes + 7) / 8 * 8;     // the tail holds 8-byte counters: keep it 32-byte aligned
    if (cfg->ibm_mailbox_nodes < 0) { lbm_destroy(h); return fail(LBM_ERR_INVALID, "ibm_mailbox_nodes < 0"); }
    h->mail_nodes = cfg->world > 1 ? (cfg->ibm_mailbox_nodes > 0 ? cfg->ibm_mailbox_nodes : 65536) : 0;
    const size_t tail_floats = TAIL_FIXED_BYTES / 4 + (size_t)IBM_MAIL * h->mail_nodes;      // counters, adapter mailbox, IBM mailbox (exported with the same IPC handle)
    bool ok = dmalloc(h, &h->pop, pop_floats + tail_floats) == cudaSuccess &&
              dmalloc(h, &h->sync_timeout, 1) == cudaSuccess && dmalloc(h, &h->d_net, 1) == cudaSuccess &&
              dmalloc(h, &h->ring, (size_t)2 * h->perim * Q) == cudaSuccess &&
              dmalloc(h, &h->sums, 3) == cudaSuccess && dmalloc(h, &h->avg, 3) == cudaSuccess &&
              dmalloc(h, &h->mass_acc, 1) == cudaSuccess;
    if (ok && cfg->collision == LBM_CM_OPTIMAL) {     // fp64 triples of the reduction blocks + its ticket
        ok = dmalloc(h, &h->stage, (size_t)3 * RED_BLOCKS + 1) == cudaSuccess;
        if (ok) cudaMemsetAsync(h->stage, 0, ((size_t)3 * RED_BLOCKS + 1) * sizeof(double), h->stream);
    }
    if (!ok) { std::string m = cudaGetErrorString(cudaGetLastError()); lbm_destroy(h); return fail(LBM_ERR_CUDA, "device allocation failed: " + m); }
    cudaMemsetAsync(h->pop, 0, (pop_floats + tail_floats) * sizeof(float), h->stream);
    cudaMemsetAsync(h->sync_timeout, 0, sizeof(int), h->stream);
    h->sync_flags = reinterpret_cast<unsigned long long*>(h->pop + pop_floats);     // [0..1] step counters, [2..3] IBM stage counters
    h->ibm_mail = h->mail_nodes ? h->pop + pop_floats + TAIL_FIXED_BYTES / 4 : nullptr;
    cudaMemsetAsync(h->d_net, 0, sizeof(SlabNet), h->stream);
    cudaMemsetAsync(h->ring, 0, (size_t)2 * h->perim * Q * sizeof(float), h->stream);
    if (const char* v = getenv("LBM_B200_ODD")) h->odd_interleaved = v[0] != '0';
    float one[3] = {1.f, 1.f, 1.f};
    cudaMemcpyAsync(h->avg, one, sizeof(one), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    *out = h;
    return LBM_OK;
}

extern "C" int lbm_set_stream(lbm_handle* h, void* s) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    cudaSetDevice(h->cfg.device);
    CU(cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return LBM_OK;
}

static int ensure_flags(lbm_handle* h) {
    if (h->flags) return LBM_OK;
    CU(dmalloc(h, &h->flags, (size_t)h->nloc));
    CU(cudaMemsetAsync(h->flags, 0, (size_t)h->nloc, h->stream));
    return LBM_OK;
}

static int rebuild_ibm(lbm_handle* h);
// Everything a step builds lazily (segment classification and lists via thrust, partial-sum buffers: cudaMalloc and stream
// synchronisation inside) is settled by the calls that change the configuration and by the init calls, never inside lbm_step: with
// peer-mapped slabs stepping from several host threads, a cudaMalloc (a device-wide synchronisation) issued by one slab's first step
// while another slab's handshake kernel already spins for it would dead-lock until the handshake timeout.
static int prepare_resources(lbm_handle* h, bool want_macros);
static int settle(lbm_handle* h) { return h->direct() ? prepare_resources(h, false) : LBM_OK; }

extern "C" int lbm_set_flags(lbm_handle* h, const int32_t* flags) {
    if (!h || !flags) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    const int nx = h->cfg.nx, ny = h->cfg.ny;
    std::vector<uint8_t> loc((size_t)h->nloc);
    std::vector<std::pair<long long, long long>> nbr;     // (node, source node)
    bool any = false;
    for (int yl = 0; yl < h->nyl; yl++)
        for (int x = 0; x < nx; x++) {
            int y = h->y0 + yl;
            long long node = (long long)y * nx + x;
            int f = flags[node];
            if (f < 0 || f > 31) return fail(LBM_ERR_INVALID, "flag out of range");
            loc[(size_t)yl * nx + x] = (uint8_t)f;
            any |= (f != 0);
            long long src = -1;
            if (f == LBM_ZG_OUTFLOW) {                  // zeroGradientOutflow.cuh:16-42
                int ix = x, iy = y;
                if (x == 0) ix = 1; else if (x == nx - 1) ix = nx - 2; else if (y == 0) iy = 1; else if (y == ny - 1) iy = ny - 2; else continue;
                src = (long long)iy * nx + ix;
            } else if (f == LBM_PRESSURE_OUTLET) {      // pressureOutlet.cuh:10-13
                if (x == 0) return fail(LBM_ERR_INVALID, "PRESSURE_OUTLET at x=0 has no x-1 neighbour");
                src = (long long)y * nx + (x - 1);
            } else if (f == LBM_REGULARIZED_BOUNCE_BACK_CORNER) {   // regularizedBounceBack.cuh:113-145
                bool l = x == 0, r = x == nx - 1, b = y == 0, t = y == ny - 1;
                if (!((l || r) && (b || t))) continue;
                int dx = l ? x + 1 : x - 1, dy = b ? y + 1 : y - 1;
                dx = std::max(1, std::min(dx, nx - 2)); dy = std::max(1, std::min(dy, ny - 2));
                src = (long long)dy * nx + dx;
            }
            if (src >= 0) {
                int sy = (int)(src / nx);
                if (sy < h->y0 || sy >= h->y0 + h->nyl) return fail(LBM_ERR_INVALID, "boundary node reads a neighbour owned by another slab");
                nbr.emplace_back(node, src);
            }
        }
    // The interior neighbour a boundary node reads belongs to the general kernel too (a plain FLUID cell there: same arithmetic): the
    // gather of its post-stream populations then only has to precede the GENERAL kernels and runs beside the vectorised one, off the
    // step's critical path (it used to sit in front of the fork: ~4 us of every step of the cavity).
    for (const auto& ns : nbr) loc[(size_t)(ns.second - (long long)h->y0 * nx)] |= FLAG_OWNED;
    CU(cudaStreamSynchronize(h->stream));           // steps enqueued by an asynchronous lbm_step may still read the old flags / nbr_* arrays
    if (h->nbr_nodes) { cudaFree(h->nbr_nodes); cudaFree(h->nbr_src); cudaFree(h->nbr_g); h->nbr_nodes = h->nbr_src = nullptr; h->nbr_g = nullptr; }
    h->nbr_count = (int)nbr.size(); h->nbrg_for_ts = -1; h->pre_for_ts = -1;
    if (any || h->flags) {
        int rc = ensure_flags(h); if (rc) return rc;
        CU(cudaMemcpyAsync(h->flags, loc.data(), loc.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (h->nbr_count) {
        std::sort(nbr.begin(), nbr.end());
        std::vector<long long> a(nbr.size()), b(nbr.size());
        for (size_t i = 0; i < nbr.size(); i++) { a[i] = nbr[i].first; b[i] = nbr[i].second; }
        CU(dmalloc(h, &h->nbr_nodes, a.size())); CU(dmalloc(h, &h->nbr_src, a.size())); CU(dmalloc(h, &h->nbr_g, a.size() * Q));
        CU(cudaMemcpyAsync(h->nbr_nodes, a.data(), a.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->nbr_src, b.data(), b.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    h->segs_dirty = true;
    if (!h->h_pts.empty()) return rebuild_ibm(h);       // re-mark the IBM bit
    return settle(h);
}

extern "C" int lbm_set_body_force(lbm_handle* h, float fx, float fy) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    h->cfg.force_x = fx; h->cfg.force_y = fy;
    return LBM_OK;
}

extern "C" int lbm_set_force_field(lbm_handle* h, const float* force_aos) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    // h->stream is a non-blocking stream: order the update after the steps an asynchronous lbm_step has enqueued on it
    CU(cudaStreamSynchronize(h->stream));
    if (!force_aos) {
        if (h->force_plane) { cudaFree(h->force_plane); h->bytes -= (long long)h->nloc * (long long)sizeof(float2); h->force_plane = nullptr; h->segs_dirty = true; }
        return settle(h);
    }
    if (!h->force_plane) { CU(dmalloc(h, &h->force_plane, (size_t)h->nloc)); h->segs_dirty = true; }
    CU(cudaMemcpyAsync(h->force_plane, force_aos + (size_t)2 * h->y0 * h->cfg.nx, (size_t)h->nloc * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return settle(h);
}

extern "C" int lbm_set_force_field_device(lbm_handle* h, const float* d_force) {
    if (!h || !d_force) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    if (!h->force_plane) { CU(cudaStreamSynchronize(h->stream)); CU(dmalloc(h, &h->force_plane, (size_t)h->nloc)); h->segs_dirty = true; }
    // stream-ordered behind the steps already enqueued; the caller's buffer may be reused once the call returns
    CU(cudaMemcpyAsync(h->force_plane, d_force, (size_t)h->nloc * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return settle(h);
}


// The engine is ONE translation unit (every kernel template is instantiated once, every launch site sees the same
// function pointers that lbm_create preloads); its host code is kept in five parts by subject:
#include "engine_ibm.inc"    // immersed bodies: structure build on the GPU, groups, mailbox slots
#include "engine_step.inc"   // init, the time step, graph replay, read-back
#include "engine_pipeline.inc" // lbm_run_from_host: a driver segment with the PCIe copies hidden behind the kernels
#include "engine_io.inc"     // validation reductions, checkpoint / restart
#include "engine_slab.inc"   // halo rows, peer-mapped neighbours

// ------------------------------------------------------------------ diagnostics
extern "C" int lbm_moment_avg(lbm_handle* h, float out[3]) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(out, h->avg, 12, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

extern "C" int lbm_get_moment_sums(lbm_handle* h, double s[3]) {
    if (!h || !s) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(s, h->sums, 24, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

extern "C" int lbm_set_moment_sums(lbm_handle* h, const double s[3]) {
    if (!h || !s) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->sums, s, 24, cudaMemcpyHostToDevice, h->stream));
    sums_to_avg_kernel<<<1, 32, 0, h->stream>>>(h->sums, h->avg, 1.0 / ((double)h->cfg.nx * (double)h->cfg.ny));
    h->launches++;
    h->avg_for_ts = h->timestep + 1;
    return LBM_OK;
}

extern "C" int lbm_set_lookahead(lbm_handle* h, int32_t bounded) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    h->bounded_lookahead = bounded != 0;
    return LBM_OK;
}

extern "C" int lbm_adapter_sums_pending(lbm_handle* h) {
    if (!h || h->cfg.collision != LBM_CM_OPTIMAL) return 0;
    return h->avg_for_ts != h->timestep + 1 ? 1 : 0;
}

extern "C" int lbm_info(lbm_handle* h, lbm_info_t* o) {
    if (!h || !o) return fail(LBM_ERR_INVALID, "NULL argument");
    memset(o, 0, sizeof(*o));
    o->nx = h->cfg.nx; o->ny = h->cfg.ny; o->y0 = h->y0; o->ny_local = h->nyl; o->rank = h->cfg.rank; o->world = h->cfg.world;
    o->timestep = h->timestep; o->num_markers = h->np_total; o->num_ibm_nodes = h->ibm_count; o->num_neighbour_bc_nodes = h->nbr_count;
    o->device_bytes = h->bytes; o->bytes_per_cell = (double)h->bytes / (double)h->nloc; o->kernel_launches = h->launches;
    return LBM_OK;
}

