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
#include <string>
#include <vector>
#include <algorithm>
#include <unistd.h>

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
    cudaEvent_t ev_bridge[2] = {nullptr, nullptr};      // legacy / per-thread user stream <-> own stream around graph replays
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
    // adapter
    float* partials = nullptr; long long n_partials = 0; double* stage = nullptr; double* sums = nullptr; float* avg = nullptr; int avg_for_ts = -1; int pre_for_ts = -1;
    // macroscopics
    float* rho_out = nullptr; float2* u_out = nullptr; int macros_ts = -1;
    double* mass_acc = nullptr;
    double* val_stage = nullptr; long long val_stage_n = 0;      // validation reductions (lbm_*_error_sums, lbm_row_mean_velocity)
    int timestep = 0;
    long long launches = 0;
    long long bytes = 0;
    float omega = 1.0f;
    // peer-mapped slab coupling (lbm_peer_export / lbm_peer_attach)
    unsigned long long* sync_flags = nullptr;      // 2 step counters written by the neighbours + padding, at the end of `pop`
    int* sync_timeout = nullptr;
    struct Peer { float* base = nullptr; void* ipc_base = nullptr; long long plane = 0, off = 0; unsigned long long* flag = nullptr; bool attached = false;
                  float* mail = nullptr; int y0 = 0, nyl = 0; } peer[2];
    cudaIpcMemHandle_t peer_ipc[2]{};
    bool direct() const { return peer[0].attached || peer[1].attached; }
};

// what one slab tells its neighbours (lbm_peer_export); opaque to callers, LBM_PEER_DESC_BYTES long
struct PeerDesc {
    cudaIpcMemHandle_t ipc;         // of the population allocation
    long long pid;
    unsigned long long raw;         // device pointer, valid inside the exporting process
    long long plane;                // floats per slot plane
    long long flags_off;            // byte offset of the step / IBM stage counters inside the allocation
    long long mail_off;             // byte offset of the IBM mailbox inside the allocation
    int nx, nyl, device, rank, y0, mail_nodes;
};
static_assert(sizeof(PeerDesc) <= LBM_PEER_DESC_BYTES, "PeerDesc must fit the ABI buffer");

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
    p.segmask = nullptr; p.nsx = h->nsx; p.gen_list = nullptr; p.plane = (long long)h->plane;
    for (int sd = 0; sd < 2; sd++) { p.peer[sd] = h->peer[sd].attached ? h->peer[sd].base : nullptr; p.peer_plane[sd] = h->peer[sd].plane; p.peer_off[sd] = h->peer[sd].off; }
    return p;
}

// CUDA loads kernels lazily, at their first launch, and that load waits for the device to drain.  A first launch issued while
// a handshake kernel of another handle in this process spins would therefore dead-lock until the handshake timeout
// (seen on a B200: two peer-mapped slabs in one process, 7 steps enqueued at once).  Every kernel a step can launch is
// loaded up front instead, once per device, before any handle exists.
template <typename K> static void preload(K kernel) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, kernel); }
template <int COLL> static void preload_coll() {
    preload(step_vec_kernel<COLL, false>); preload(step_vec_kernel<COLL, true>);
    preload(step_kernel<COLL, false, false>); preload(step_kernel<COLL, false, true>);
    preload(step_kernel<COLL, true, false>); preload(step_kernel<COLL, true, true>);
}
static void preload_kernels(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    done[device] = true;
    preload_coll<C_BGK>(); preload_coll<C_MRT>(); preload_coll<C_CM>(); preload_coll<C_CMOPT>();
    preload(moments_kernel<false>); preload(moments_kernel<true>); preload(moments_vec_kernel<false>); preload(moments_vec_kernel<true>);
    preload(reduce_stage1_kernel); preload(reduce_stage2_kernel); preload(sums_to_avg_kernel);
    preload(nbr_gather_kernel<false>); preload(nbr_gather_kernel<true>);
    preload(ibm_kernel<false>); preload(ibm_kernel<true>); preload(ibm_gather_kernel<false>); preload(ibm_gather_kernel<true>); preload(ibm_solve_kernel);
    preload(wait_neighbours_kernel); preload(signal_neighbours_kernel); preload(build_segmask_kernel);
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
    for (int sd = 0; sd < 2; sd++)
        if (h->peer[sd].ipc_base && !(sd == 1 && h->peer[0].ipc_base == h->peer[1].ipc_base)) cudaIpcCloseMemHandle(h->peer[sd].ipc_base);
    void* ptrs[] = {h->sync_timeout, h->pop, h->flags, h->force_plane, h->ring, h->nbr_nodes, h->nbr_src, h->nbr_g, h->d_pts, h->ibm_nodes,
                    h->sten_idx, h->sten_w, h->csr_row, h->csr_k, h->csr_w, h->ibm_rho, h->ibm_uprev, h->ibm_lagF, h->ibm_force, h->ibm_mail_idx, h->d_utarget,
                    h->partials, h->stage, h->sums, h->avg, h->rho_out, h->u_out, h->mass_acc, h->segmask, h->gen_list, h->val_stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (auto& g : h->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& e : h->ev_bridge) if (e) cudaEventDestroy(e);
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
    h->perim = 2 * cfg->nx + 2 * cfg->ny;
    const size_t pop_floats = h->plane * h->nplanes;
    if (cfg->ibm_mailbox_nodes < 0) { cudaStreamDestroy(h->own_stream); delete h; return fail(LBM_ERR_INVALID, "ibm_mailbox_nodes < 0"); }
    h->mail_nodes = cfg->world > 1 ? (cfg->ibm_mailbox_nodes > 0 ? cfg->ibm_mailbox_nodes : 65536) : 0;
    const size_t tail_floats = 64 + (size_t)IBM_MAIL * h->mail_nodes;         // 256 B of neighbour counters + the IBM mailbox (exported with the same IPC handle)
    bool ok = dmalloc(h, &h->pop, pop_floats + tail_floats) == cudaSuccess &&
              dmalloc(h, &h->sync_timeout, 1) == cudaSuccess &&
              dmalloc(h, &h->ring, (size_t)2 * h->perim * Q) == cudaSuccess &&
              dmalloc(h, &h->sums, 3) == cudaSuccess && dmalloc(h, &h->avg, 3) == cudaSuccess &&
              dmalloc(h, &h->mass_acc, 1) == cudaSuccess;
    if (ok && cfg->collision == LBM_CM_OPTIMAL) ok = dmalloc(h, &h->stage, (size_t)3 * RED_BLOCKS) == cudaSuccess;
    if (!ok) { std::string m = cudaGetErrorString(cudaGetLastError()); lbm_destroy(h); return fail(LBM_ERR_CUDA, "device allocation failed: " + m); }
    cudaMemsetAsync(h->pop, 0, (pop_floats + tail_floats) * sizeof(float), h->stream);
    cudaMemsetAsync(h->sync_timeout, 0, sizeof(int), h->stream);
    h->sync_flags = reinterpret_cast<unsigned long long*>(h->pop + pop_floats);     // [0..1] step counters, [2..3] IBM stage counters
    h->ibm_mail = h->mail_nodes ? h->pop + pop_floats + 64 : nullptr;
    cudaMemsetAsync(h->ring, 0, (size_t)2 * h->perim * Q * sizeof(float), h->stream);
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
    if (h->nbr_nodes) { cudaFree(h->nbr_nodes); cudaFree(h->nbr_src); cudaFree(h->nbr_g); h->nbr_nodes = h->nbr_src = nullptr; h->nbr_g = nullptr; }
    h->nbr_count = (int)nbr.size();
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
        CU(cudaMemcpy(h->nbr_nodes, a.data(), a.size() * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->nbr_src, b.data(), b.size() * 8, cudaMemcpyHostToDevice));
    }
    h->segs_dirty = true;
    if (!h->h_pts.empty()) return rebuild_ibm(h);       // re-mark the IBM bit
    return LBM_OK;
}

extern "C" int lbm_set_body_force(lbm_handle* h, float fx, float fy) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    h->cfg.force_x = fx; h->cfg.force_y = fy;
    return LBM_OK;
}

extern "C" int lbm_set_force_field(lbm_handle* h, const float* force_aos) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    h->segs_dirty = true;
    if (!force_aos) { if (h->force_plane) { cudaFree(h->force_plane); h->force_plane = nullptr; } return LBM_OK; }
    if (!h->force_plane) CU(dmalloc(h, &h->force_plane, (size_t)h->nloc));
    CU(cudaMemcpy(h->force_plane, force_aos + (size_t)2 * h->y0 * h->cfg.nx, (size_t)h->nloc * sizeof(float2), cudaMemcpyHostToDevice));
    return LBM_OK;
}

extern "C" int lbm_set_force_field_device(lbm_handle* h, const float* d_force) {
    if (!h || !d_force) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    h->segs_dirty = true;
    if (!h->force_plane) CU(dmalloc(h, &h->force_plane, (size_t)h->nloc));
    CU(cudaMemcpyAsync(h->force_plane, d_force, (size_t)h->nloc * sizeof(float2), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

// ------------------------------------------------------------------ IBM structure, built on the GPU
__global__ void mark_ibm_kernel(uint8_t* flags, const long long* nodes, int n, long long node0, long long nloc, int set) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long ln = nodes[i] - node0;
    if (ln < 0 || ln >= nloc) return;               // a stencil node another slab owns
    if (set) flags[ln] |= FLAG_IBM; else flags[ln] &= (uint8_t)~FLAG_IBM;
}
__global__ void csr_fill_kernel(const int* order, const int* sten_idx_flat, const float* sten_w, int ss, int n, int* csr_k, float* csr_w) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int s = order[e];
    csr_k[e] = s / ss; csr_w[e] = sten_w[s];
}
struct not_neg_i { __host__ __device__ bool operator()(int v) const { return v >= 0; } };
struct fix_neg_slots { const long long* n; int* i; __device__ void operator()(int s) const { if (n[s] < 0) i[s] = -1; } };
struct gather_idx { const int* idx; const int* ord; int* out; __device__ void operator()(int e) const { out[e] = idx[ord[e]]; } };

// Several slabs: which bodies does this slab work on, and where do their nodes sit in the mailbox?  Host-side, O(markers).
// Bodies whose stencils share a lattice node are coupled through the spreading sum and form one group; a slab works on
// every group it owns a node of, with ALL markers of the group, so that each slab of a group computes the same bits.
// The stencil node ids follow ibm_stencil_kernel exactly (same float arithmetic).
static void select_bodies(lbm_handle* h, std::vector<float>& act_pts, std::vector<float>& act_vel, std::vector<long long>& all_nodes) {
    const bool two = (h->cfg.quirks & LBM_QK_D8_IBM_2X2) != 0;
    const int w = two ? 2 : 4, lo = two ? 0 : -1, ss = w * w, nx = h->cfg.nx, ny = h->cfg.ny;
    const int np = (int)(h->h_pts.size() / 2), nb = (int)h->body_start.size();
    std::vector<long long> node((size_t)np * ss, -1);
    for (int k = 0; k < np; k++) {
        const float px = h->h_pts[2 * k], py = h->h_pts[2 * k + 1];
        const float gx = floorf(px), gy = floorf(py);
        for (int i = 0; i < w; i++)
            for (int j = 0; j < w; j++) {
                const int nxx = (int)(gx + (i + lo)), nyy = (int)(gy + (j + lo));
                if (nxx >= nx || nxx < 0 || nyy >= ny || nyy < 0) continue;
                node[(size_t)k * ss + i * w + j] = (long long)nyy * nx + nxx;
            }
    }
    all_nodes.clear();
    for (long long v : node) if (v >= 0) all_nodes.push_back(v);
    std::sort(all_nodes.begin(), all_nodes.end());
    all_nodes.erase(std::unique(all_nodes.begin(), all_nodes.end()), all_nodes.end());
    // union-find over bodies through shared nodes
    std::vector<int> parent(nb), first_body(all_nodes.size(), -1);
    for (int b = 0; b < nb; b++) parent[b] = b;
    auto find = [&](int b) { while (parent[b] != b) { parent[b] = parent[parent[b]]; b = parent[b]; } return b; };
    auto body_end = [&](int b) { return b + 1 < nb ? h->body_start[b + 1] : np; };
    for (int b = 0; b < nb; b++)
        for (int k = h->body_start[b]; k < body_end(b); k++)
            for (int sidx = 0; sidx < ss; sidx++) {
                const long long v = node[(size_t)k * ss + sidx];
                if (v < 0) continue;
                const size_t i = std::lower_bound(all_nodes.begin(), all_nodes.end(), v) - all_nodes.begin();
                if (first_body[i] < 0) first_body[i] = b; else parent[find(b)] = find(first_body[i]);
            }
    std::vector<char> owns(nb, 0);
    for (size_t i = 0; i < all_nodes.size(); i++) {
        const int y = (int)(all_nodes[i] / nx);
        if (y >= h->y0 && y < h->y0 + h->nyl) owns[find(first_body[i])] = 1;
    }
    act_pts.clear(); act_vel.clear();
    for (int b = 0; b < nb; b++)
        if (owns[find(b)]) {
            act_pts.insert(act_pts.end(), h->h_pts.begin() + 2 * (size_t)h->body_start[b], h->h_pts.begin() + 2 * (size_t)body_end(b));
            act_vel.insert(act_vel.end(), h->h_vel.begin() + 2 * (size_t)h->body_start[b], h->h_vel.begin() + 2 * (size_t)body_end(b));
        }
}

static int rebuild_ibm(lbm_handle* h) {
    CU(cudaSetDevice(h->cfg.device));
    h->segs_dirty = true;
    auto pol = thrust::cuda::par.on(h->stream);
    if (h->ibm_nodes && h->flags && h->ibm_count) {
        mark_ibm_kernel<<<(h->ibm_count + 255) / 256, 256, 0, h->stream>>>(h->flags, h->ibm_nodes, h->ibm_count, (long long)h->y0 * h->cfg.nx, h->nloc, 0);
        h->launches++;
    }
    void* old[] = {h->d_pts, h->ibm_nodes, h->sten_idx, h->sten_w, h->csr_row, h->csr_k, h->csr_w, h->ibm_rho, h->ibm_uprev, h->ibm_lagF, h->ibm_force, h->ibm_mail_idx, h->d_utarget};
    CU(cudaStreamSynchronize(h->stream));
    for (void* p : old) if (p) cudaFree(p);
    h->d_pts = nullptr; h->ibm_nodes = nullptr; h->sten_idx = nullptr; h->sten_w = nullptr; h->csr_row = nullptr; h->csr_k = nullptr; h->csr_w = nullptr;
    h->ibm_rho = nullptr; h->ibm_uprev = nullptr; h->ibm_lagF = nullptr; h->ibm_force = nullptr; h->ibm_mail_idx = nullptr; h->d_utarget = nullptr;
    h->np_total = (int)(h->h_pts.size() / 2);
    h->np = 0; h->ibm_count = 0; h->nall = 0; h->ibm_rows.clear(); h->pre_for_ts = -1; h->ibm_mail_for_ts = -1;
    if (h->np_total == 0) return LBM_OK;
    // the markers this slab works on: all of them on a single slab, the bodies it owns a node of otherwise
    std::vector<float> act_pts, act_vel;
    std::vector<long long> all_nodes;
    const bool multi = h->cfg.world > 1;
    if (multi) {
        select_bodies(h, act_pts, act_vel, all_nodes);
        h->nall = (long long)all_nodes.size();
        if (h->nall > h->mail_nodes)
            return fail(LBM_ERR_INVALID, "the bodies touch " + std::to_string(h->nall) + " lattice nodes, more than the IBM mailbox holds (lbm_config.ibm_mailbox_nodes = " + std::to_string(h->mail_nodes) + ")");
    }
    const std::vector<float>& pts = multi ? act_pts : h->h_pts;
    const std::vector<float>& vel = multi ? act_vel : h->h_vel;
    h->np = (int)(pts.size() / 2);
    if (h->np == 0) return LBM_OK;
    const int np = h->np;
    if (h->has_vel && !(h->cfg.quirks & LBM_QK_D9_IBM_ZERO_TARGET)) {
        CU(dmalloc(h, &h->d_utarget, (size_t)np));
        CU(cudaMemcpyAsync(h->d_utarget, vel.data(), (size_t)2 * np * 4, cudaMemcpyHostToDevice, h->stream));
    }
    const bool two = (h->cfg.quirks & LBM_QK_D8_IBM_2X2) != 0;
    const int w = two ? 2 : 4, lo = two ? 0 : -1, ss = w * w;
    h->ibm_ss = ss;
    const int nslots = np * ss;
    long long* sten_node = nullptr; long long* keys = nullptr; int* order = nullptr; int* node_of = nullptr;
    CU(dmalloc(h, &h->d_pts, (size_t)2 * np));
    CU(cudaMemcpyAsync(h->d_pts, pts.data(), (size_t)2 * np * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMalloc(&sten_node, (size_t)nslots * 8)); CU(cudaMalloc(&keys, (size_t)nslots * 8));
    CU(cudaMalloc(&order, (size_t)nslots * 4)); CU(cudaMalloc(&node_of, (size_t)nslots * 4));
    CU(dmalloc(h, &h->sten_w, (size_t)nslots)); CU(dmalloc(h, &h->sten_idx, (size_t)nslots));
    ibm_stencil_kernel<<<(np + 127) / 128, 128, 0, h->stream>>>(h->d_pts, np, h->cfg.nx, h->cfg.ny, lo, w, sten_node, h->sten_w);
    h->launches++;
    // unique sorted node list
    CU(cudaMemcpyAsync(keys, sten_node, (size_t)nslots * 8, cudaMemcpyDeviceToDevice, h->stream));
    thrust::device_ptr<long long> kp(keys);
    thrust::sort(pol, kp, kp + nslots);
    auto kend = thrust::unique(pol, kp, kp + nslots);
    int nuniq = (int)(kend - kp);
    long long first = 0;
    CU(cudaMemcpyAsync(&first, keys, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int skip = first < 0 ? 1 : 0;           // the -1 marker of out-of-domain slots sorts first
    h->ibm_count = nuniq - skip;
    if (h->ibm_count <= 0) { h->ibm_count = 0; cudaFree(sten_node); cudaFree(keys); cudaFree(order); cudaFree(node_of); return LBM_OK; }
    CU(dmalloc(h, &h->ibm_nodes, (size_t)h->ibm_count));
    CU(cudaMemcpyAsync(h->ibm_nodes, keys + skip, (size_t)h->ibm_count * 8, cudaMemcpyDeviceToDevice, h->stream));
    // compact index of every stencil slot (-1 where the slot is outside the domain)
    thrust::device_ptr<long long> np_(h->ibm_nodes), sn(sten_node);
    thrust::device_ptr<int> si(h->sten_idx);
    thrust::lower_bound(pol, np_, np_ + h->ibm_count, sn, sn + nslots, si);
    // slots with node -1 get index 0 from lower_bound: overwrite with -1
    thrust::for_each(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(nslots), fix_neg_slots{sten_node, h->sten_idx});
    // CSR node <- (marker, weight): stable sort of the valid slots by compact node index keeps markers ascending
    thrust::device_ptr<int> op(order), nf(node_of);
    auto oend = thrust::copy_if(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(nslots), si, op, not_neg_i());
    int nvalid = (int)(oend - op);
    thrust::for_each(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(nvalid), gather_idx{h->sten_idx, order, node_of});
    thrust::stable_sort_by_key(pol, nf, nf + nvalid, op);
    CU(dmalloc(h, &h->csr_row, (size_t)h->ibm_count + 1)); CU(dmalloc(h, &h->csr_k, (size_t)nvalid)); CU(dmalloc(h, &h->csr_w, (size_t)nvalid));
    thrust::device_ptr<int> rp(h->csr_row);
    thrust::lower_bound(pol, nf, nf + nvalid, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(h->ibm_count + 1), rp);
    csr_fill_kernel<<<(nvalid + 255) / 256, 256, 0, h->stream>>>(order, h->sten_idx, h->sten_w, ss, nvalid, h->csr_k, h->csr_w);
    h->launches++;
    CU(dmalloc(h, &h->ibm_rho, (size_t)h->ibm_count)); CU(dmalloc(h, &h->ibm_uprev, (size_t)h->ibm_count));
    CU(dmalloc(h, &h->ibm_lagF, (size_t)np)); CU(dmalloc(h, &h->ibm_force, (size_t)h->ibm_count));
    if (multi) {
        // mailbox slot of every node of this slab's list = its rank in the global node list; rows for the peer-coverage check
        std::vector<long long> mine((size_t)h->ibm_count);
        CU(cudaMemcpyAsync(mine.data(), h->ibm_nodes, mine.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        std::vector<int> idx(mine.size());
        for (size_t i = 0; i < mine.size(); i++) {
            auto it = std::lower_bound(all_nodes.begin(), all_nodes.end(), mine[i]);
            if (it == all_nodes.end() || *it != mine[i]) { cudaFree(sten_node); cudaFree(keys); cudaFree(order); cudaFree(node_of); return fail(LBM_ERR_STATE, "IBM: host and device stencil nodes disagree"); }
            idx[i] = (int)(it - all_nodes.begin());
            const int y = (int)(mine[i] / h->cfg.nx);
            if (h->ibm_rows.empty() || h->ibm_rows.back() != y) h->ibm_rows.push_back(y);
        }
        CU(dmalloc(h, &h->ibm_mail_idx, idx.size()));
        CU(cudaMemcpyAsync(h->ibm_mail_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = ensure_flags(h); if (rc) return rc;
    mark_ibm_kernel<<<(h->ibm_count + 255) / 256, 256, 0, h->stream>>>(h->flags, h->ibm_nodes, h->ibm_count, (long long)h->y0 * h->cfg.nx, h->nloc, 1);
    h->launches++;
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(sten_node); cudaFree(keys); cudaFree(order); cudaFree(node_of);
    CU(cudaGetLastError());
    return LBM_OK;
}

extern "C" int lbm_add_body(lbm_handle* h, const float* pts, int32_t n) {
    if (!h || (!pts && n > 0) || n < 0) return fail(LBM_ERR_INVALID, "bad body");
    if (n == 0) return LBM_OK;
    h->body_start.push_back((int)(h->h_pts.size() / 2));
    h->h_pts.insert(h->h_pts.end(), pts, pts + (size_t)2 * n);
    h->h_vel.resize(h->h_pts.size(), 0.0f);
    int rc = rebuild_ibm(h);
    if (rc != LBM_OK) {         // leave the handle as it was before the call
        h->h_pts.resize((size_t)2 * h->body_start.back());
        h->h_vel.resize(h->h_pts.size());
        h->body_start.pop_back();
        const std::string keep = g_err;
        rebuild_ibm(h);
        g_err = keep;
    }
    return rc;
}

static int body_range(lbm_handle* h, int body, size_t& first, size_t& count) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    const int nb = (int)h->body_start.size();
    if (body < 0 || body >= nb) return fail(LBM_ERR_INVALID, "no such body");
    first = (size_t)h->body_start[body];
    count = (body + 1 < nb ? (size_t)h->body_start[body + 1] : h->h_pts.size() / 2) - first;
    return LBM_OK;
}

extern "C" int lbm_set_body_velocities(lbm_handle* h, int32_t body, const float* vel) {
    size_t first, count;
    int rc = body_range(h, body, first, count); if (rc) return rc;
    if (vel) { std::copy(vel, vel + 2 * count, h->h_vel.begin() + 2 * first); h->has_vel = true; }
    else std::fill(h->h_vel.begin() + 2 * first, h->h_vel.begin() + 2 * (first + count), 0.0f);
    return rebuild_ibm(h);
}

extern "C" int lbm_move_body(lbm_handle* h, int32_t body, const float* pts) {
    size_t first, count;
    int rc = body_range(h, body, first, count); if (rc) return rc;
    if (!pts) return fail(LBM_ERR_INVALID, "NULL argument");
    const std::vector<float> keep(h->h_pts.begin() + 2 * first, h->h_pts.begin() + 2 * (first + count));
    std::copy(pts, pts + 2 * count, h->h_pts.begin() + 2 * first);
    rc = rebuild_ibm(h);
    if (rc != LBM_OK) {         // e.g. the moved body no longer fits the mailbox: back to where it was
        std::copy(keep.begin(), keep.end(), h->h_pts.begin() + 2 * first);
        const std::string msg = g_err;
        rebuild_ibm(h);
        g_err = msg;
    }
    return rc;
}

// ------------------------------------------------------------------ init
static int ensure_macros(lbm_handle* h) {
    if (h->rho_out) return LBM_OK;
    CU(dmalloc(h, &h->rho_out, (size_t)h->nloc)); CU(dmalloc(h, &h->u_out, (size_t)h->nloc));
    return LBM_OK;
}

extern "C" int lbm_reserve_macroscopics(lbm_handle* h) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    return ensure_macros(h);
}

extern "C" int lbm_init_fields_device(lbm_handle* h, const float* d_rho, const float* d_u) {
    if (!h || !d_rho || !d_u) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    int rc = ensure_macros(h); if (rc) return rc;
    Params p = make_params(h, 0);
    p.rho_out = h->rho_out; p.u_out = h->u_out;
    init_fields_kernel<<<grid_of(h), BX, 0, h->stream>>>(p, d_rho, (const float2*)d_u);
    h->launches++;
    CU(cudaGetLastError());
    h->timestep = 0; h->macros_ts = 0; h->avg_for_ts = -1; h->pre_for_ts = -1;
    CU(cudaMemsetAsync(h->sync_flags, 0, 32, h->stream));
    h->ibm_mail_for_ts = -1; h->nbrg_for_ts = -1;
    return LBM_OK;
}

extern "C" int lbm_init_fields_local(lbm_handle* h, const float* rho, const float* u) {
    if (!h || !rho || !u) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    // staging in the macroscopic planes themselves: init_fields_kernel reads rho/u and writes the same values back
    int rc = ensure_macros(h); if (rc) return rc;
    CU(cudaMemcpyAsync(h->rho_out, rho, (size_t)h->nloc * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->u_out, u, (size_t)h->nloc * 8, cudaMemcpyHostToDevice, h->stream));
    return lbm_init_fields_device(h, h->rho_out, (const float*)h->u_out);
}

extern "C" int lbm_init_fields(lbm_handle* h, const float* rho, const float* u) {
    if (!h || !rho || !u) return fail(LBM_ERR_INVALID, "NULL argument");
    size_t off = (size_t)h->y0 * h->cfg.nx;
    return lbm_init_fields_local(h, rho + off, u + 2 * off);
}

extern "C" int lbm_init_taylor_green(lbm_handle* h, float nu, float u0) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    Params p = make_params(h, 0);
    if (h->rho_out) { p.rho_out = h->rho_out; p.u_out = h->u_out; }
    init_taylor_green_kernel<<<grid_of(h), BX, 0, h->stream>>>(p, nu, u0);
    h->launches++;
    CU(cudaGetLastError());
    h->timestep = 0; h->macros_ts = h->rho_out ? 0 : -1; h->avg_for_ts = -1; h->pre_for_ts = -1;
    h->ibm_mail_for_ts = -1; h->nbrg_for_ts = -1;
    CU(cudaMemsetAsync(h->sync_flags, 0, 32, h->stream));
    return LBM_OK;
}

extern "C" int lbm_set_populations(lbm_handle* h, const float* f, const float* fb) {
    if (!h || !f) return fail(LBM_ERR_INVALID, "NULL argument");
    if (h->timestep & 1) return fail(LBM_ERR_STATE, "lbm_set_populations needs an even timestep");
    CU(cudaSetDevice(h->cfg.device));
    if (!fb) fb = f;
    float *df = nullptr, *dfb = nullptr;
    size_t off = (size_t)h->y0 * h->cfg.nx * Q, cnt = (size_t)h->nloc * Q;
    CU(cudaMalloc(&df, cnt * 4)); CU(cudaMalloc(&dfb, cnt * 4));
    CU(cudaMemcpyAsync(df, f + off, cnt * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(dfb, fb + off, cnt * 4, cudaMemcpyHostToDevice, h->stream));
    Params p = make_params(h, h->timestep);
    set_populations_kernel<<<grid_of(h), BX, 0, h->stream>>>(p, df, dfb);
    h->launches++;
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(df); cudaFree(dfb);
    h->macros_ts = -1; h->avg_for_ts = -1; h->pre_for_ts = -1; h->ibm_mail_for_ts = -1; h->nbrg_for_ts = -1;
    return LBM_OK;
}

extern "C" int lbm_get_populations(lbm_handle* h, float* f) {
    if (!h || !f) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    float* df = nullptr;
    size_t cnt = (size_t)h->nloc * Q;
    CU(cudaMalloc(&df, cnt * 4));
    Params p = make_params(h, h->timestep);
    if (h->timestep & 1) get_populations_kernel<true><<<grid_of(h), BX, 0, h->stream>>>(p, df);
    else get_populations_kernel<false><<<grid_of(h), BX, 0, h->stream>>>(p, df);
    h->launches++;
    CU(cudaMemcpyAsync(f + (size_t)h->y0 * h->cfg.nx * Q, df, cnt * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(df);
    return LBM_OK;
}

// ------------------------------------------------------------------ the time step
static bool is_general(const lbm_handle* h) {
    return h->flags != nullptr || h->force_plane != nullptr || !h->cfg.periodic_x || !h->cfg.periodic_y;
}
static bool use_vec(const lbm_handle* h) { return (h->cfg.nx % 4) == 0; }
static dim3 vec_grid(const lbm_handle* h, int& threads) {
    const int nv = h->cfg.nx / 4;
    threads = std::min(BX, ((nv + 31) / 32) * 32);
    return dim3((nv + threads - 1) / threads, h->nyl);
}

struct is_set_u8 { __host__ __device__ bool operator()(uint8_t v) const { return v != 0; } };

// (re)build the general-segment mask and list after flags / bodies / force plane changed
static int ensure_segments(lbm_handle* h) {
    if (!h->segs_dirty) return LBM_OK;
    const long long nseg = (long long)h->nsx * h->nyl;
    if (!h->segmask) { CU(dmalloc(h, &h->segmask, (size_t)nseg)); CU(dmalloc(h, &h->gen_list, (size_t)nseg)); }
    Params p = make_params(h, 0);
    build_segmask_kernel<<<(unsigned)((nseg + 255) / 256), 256, 0, h->stream>>>(p, h->segmask);
    h->launches++;
    auto pol = thrust::cuda::par.on(h->stream);
    thrust::device_ptr<uint8_t> mp(h->segmask);
    thrust::device_ptr<int> lp(h->gen_list);
    auto end = thrust::copy_if(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>((int)nseg), mp, lp, is_set_u8());
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    h->gen_count = (int)(end - lp);
    h->segs_dirty = false;
    return LBM_OK;
}

static int ensure_partials(lbm_handle* h, long long n) {
    if (n <= h->n_partials) return LBM_OK;
    if (h->partials) { CU(cudaStreamSynchronize(h->stream)); cudaFree(h->partials); h->bytes -= 12 * h->n_partials; h->partials = nullptr; }
    CU(dmalloc(h, &h->partials, (size_t)3 * n));
    h->n_partials = n;
    return LBM_OK;
}

static void reduce_partials(lbm_handle* h, long long n) {
    const double inv_n = 1.0 / ((double)h->cfg.nx * (double)h->cfg.ny);
    const int nb = (int)std::min<long long>(RED_BLOCKS, (n + 255) / 256);
    reduce_stage1_kernel<<<nb, 256, 0, h->stream>>>(h->partials, n, h->stage);
    reduce_stage2_kernel<<<1, 256, 0, h->stream>>>(h->stage, nb, h->sums, h->avg, inv_n, h->cfg.world == 1);
    h->launches += 2;
}

template <int COLL, bool ODD>
static void launch_scalar(lbm_handle* h, const Params& p, bool general, dim3 grid) {
    if (general) step_kernel<COLL, ODD, true><<<grid, BX, 0, h->stream>>>(p);
    else step_kernel<COLL, ODD, false><<<grid, BX, 0, h->stream>>>(p);
}
template <int COLL, bool ODD>
static void launch_vec(lbm_handle* h, const Params& p) {
    int threads; dim3 g = vec_grid(h, threads);
    step_vec_kernel<COLL, ODD><<<g, threads, 0, h->stream>>>(p);
}
#define DISPATCH_COLL(ODDV, CALL)                                                     \
    switch (h->cfg.collision) {                                                      \
    case LBM_BGK: { constexpr int COLL = C_BGK; constexpr bool ODD = ODDV; CALL; } break;   \
    case LBM_MRT: { constexpr int COLL = C_MRT; constexpr bool ODD = ODDV; CALL; } break;   \
    case LBM_CM: { constexpr int COLL = C_CM; constexpr bool ODD = ODDV; CALL; } break;     \
    default: { constexpr int COLL = C_CMOPT; constexpr bool ODD = ODDV; CALL; } break;      \
    }

static IbmData ibm_data(lbm_handle* h) {
    IbmData d{};
    d.np = h->np; d.nnodes = h->ibm_count; d.ss = h->ibm_ss;
    d.nodes = h->ibm_nodes; d.sten_idx = h->sten_idx; d.sten_w = h->sten_w; d.row = h->csr_row; d.csr_k = h->csr_k; d.csr_w = h->csr_w;
    d.rho = h->ibm_rho; d.uprev = h->ibm_uprev; d.lagF = h->ibm_lagF; d.force = h->ibm_force;
    d.utarget = h->d_utarget;
    d.mail_idx = h->ibm_mail_idx; d.mail = h->ibm_mail;
    d.my_flags = h->sync_flags + 2; d.timed_out = h->sync_timeout;
    for (int sd = 0; sd < 2; sd++) {
        const bool on = h->peer[sd].attached && h->peer[sd].mail;
        d.peer_mail[sd] = on ? h->peer[sd].mail : nullptr;
        d.peer_flag[sd] = on ? h->peer[sd].flag + 2 : nullptr;
        d.need[sd] = on ? 1 : 0;
    }
    return d;
}

static void launch_nbr_gather(lbm_handle* h, const Params& p, int t) {
    if (h->nbrg_for_ts == t || !h->nbr_count) return;
    if (t & 1) nbr_gather_kernel<true><<<(h->nbr_count + 127) / 128, 128, 0, h->stream>>>(p, h->nbr_src, h->nbr_g, h->nbr_count);
    else nbr_gather_kernel<false><<<(h->nbr_count + 127) / 128, 128, 0, h->stream>>>(p, h->nbr_src, h->nbr_g, h->nbr_count);
    h->launches++;
    h->nbrg_for_ts = t;
}

// peer-mapped coupling: every stencil node this slab works on must belong to it or to an attached neighbour
static int check_ibm_coverage(const lbm_handle* h) {
    for (int y : h->ibm_rows) {
        bool ok = y >= h->y0 && y < h->y0 + h->nyl;
        for (int sd = 0; sd < 2 && !ok; sd++) ok = h->peer[sd].attached && y >= h->peer[sd].y0 && y < h->peer[sd].y0 + h->peer[sd].nyl;
        if (!ok) return fail(LBM_ERR_INVALID, "a body (or a group of overlapping bodies) spans more than this slab and its two peer-mapped neighbours: use the halo coupling (lbm_ibm_pack / all-reduce / lbm_ibm_unpack)");
    }
    return LBM_OK;
}

// nbr gather + IBM (+ moments pre-pass) for step t; idempotent per timestep
static int pre_passes(lbm_handle* h, int t, bool want_moments) {
    const bool odd = (t & 1) != 0;
    Params p = make_params(h, t);
    if (h->pre_for_ts != t) {
        launch_nbr_gather(h, p, t);
        if (h->cfg.world == 1) {
            if (h->ibm_count) {
                IbmData d = ibm_data(h);
                if (odd) ibm_kernel<true><<<1, 1024, 0, h->stream>>>(p, d); else ibm_kernel<false><<<1, 1024, 0, h->stream>>>(p, d);
                h->launches++;
            }
        } else if (h->np_total > 0) {
            IbmData d = ibm_data(h);
            if (h->direct()) {
                // every slab posts its nodes (possibly none) and the stage counter; slabs that own part of a body then solve it
                if (odd) ibm_gather_kernel<true><<<1, 1024, 0, h->stream>>>(p, d, h->ibm_mail, (unsigned long long)t);
                else ibm_gather_kernel<false><<<1, 1024, 0, h->stream>>>(p, d, h->ibm_mail, (unsigned long long)t);
                h->launches++;
            } else {
                if (h->ibm_count && h->ibm_mail_for_ts != t)
                    return fail(LBM_ERR_STATE, "bodies on several slabs without peer-mapped neighbours: lbm_ibm_pack, all-reduce (sum) the buffer over the slabs, lbm_ibm_unpack before every lbm_step");
                d.need[0] = d.need[1] = 0;
            }
            if (h->ibm_count) {
                ibm_solve_kernel<<<1, 1024, 0, h->stream>>>(p, d, (unsigned long long)t);
                h->launches++;
            }
        }
        h->pre_for_ts = t;
    }
    if (want_moments) {
        long long nparts;
        if (!use_vec(h)) {
            dim3 g = grid_of(h);
            nparts = (long long)g.x * g.y;
            int rc = ensure_partials(h, nparts); if (rc) return rc;
            Params pm = p; pm.partials = h->partials;
            if (odd) moments_kernel<true><<<g, BX, 0, h->stream>>>(pm); else moments_kernel<false><<<g, BX, 0, h->stream>>>(pm);
            h->launches++;
        } else {
            const bool general = is_general(h);
            if (general) { int rc = ensure_segments(h); if (rc) return rc; }
            const int ngen = general ? h->gen_count : 0;
            int threads; dim3 gv = vec_grid(h, threads);
            const long long nvb = (long long)gv.x * gv.y;
            nparts = nvb + ngen;
            int rc = ensure_partials(h, nparts); if (rc) return rc;
            Params pm = p; pm.partials = h->partials;
            if (general) pm.segmask = h->segmask;
            if (odd) moments_vec_kernel<true><<<gv, threads, 0, h->stream>>>(pm); else moments_vec_kernel<false><<<gv, threads, 0, h->stream>>>(pm);
            h->launches++;
            if (ngen > 0) {
                Params pg = p; pg.gen_list = h->gen_list; pg.partials = h->partials + 3 * nvb;
                if (odd) moments_kernel<true><<<dim3(ngen, 1), BX, 0, h->stream>>>(pg); else moments_kernel<false><<<dim3(ngen, 1), BX, 0, h->stream>>>(pg);
                h->launches++;
            }
        }
        reduce_partials(h, nparts);
        if (h->cfg.world == 1) h->avg_for_ts = t;
    }
    return LBM_OK;
}

// fork: work enqueued on h->stream from here on runs on the side stream, after everything enqueued on the main stream so far
static int fork_side(lbm_handle* h, cudaStream_t main) {
    CU(cudaEventRecord(h->ev_fork, main));
    CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    h->stream = h->side_stream;
    return LBM_OK;
}
// join: back to the main stream, which waits for the side stream's work
static int join_side(lbm_handle* h, cudaStream_t main) {
    h->stream = main;
    CU(cudaEventRecord(h->ev_join, h->side_stream));
    CU(cudaStreamWaitEvent(main, h->ev_join, 0));
    return LBM_OK;
}

static int one_step(lbm_handle* h, bool want_macros) {
    const int t = h->timestep + 1;
    const bool odd = (t & 1) != 0;
    const bool general = is_general(h);
    const bool opt = h->cfg.collision == LBM_CM_OPTIMAL;
    const bool need_moments = opt && h->avg_for_ts != t;
    if (need_moments && h->cfg.world > 1)
        return fail(LBM_ERR_STATE, "OptimalAdapter on several slabs: call lbm_adapter_prepass, all-reduce lbm_get_moment_sums, lbm_set_moment_sums before lbm_step");
    int rc;
    cudaStream_t main = h->stream;
    if (use_vec(h) && general) { rc = ensure_segments(h); if (rc) return rc; }
    const int ngen = (use_vec(h) && general) ? h->gen_count : 0;
    // two kernels side by side: the vectorised one on the main stream, pre-passes + general segments on the side stream
    const bool side = h->overlap && ngen > 0;
    bool forked = false;
    if (side && !need_moments) {
        // the neighbour-BC gather reads cells the vectorised kernel overwrites: it stays in front of the fork
        launch_nbr_gather(h, make_params(h, t), t);
        rc = fork_side(h, main); if (rc) return rc;
        forked = true;
    }
    rc = pre_passes(h, t, need_moments);
    if (rc) { h->stream = main; return rc; }
    Params p = make_params(h, t);
    const bool lagged = opt && h->cfg.adapter_mode == LBM_ADAPTER_LAGGED;
    if (want_macros) { rc = ensure_macros(h); if (rc) { h->stream = main; return rc; } p.rho_out = h->rho_out; p.u_out = h->u_out; }
    long long nparts = 0;
    if (!use_vec(h)) {
        // nx not a multiple of 4: the scalar kernel covers the whole slab
        dim3 g = grid_of(h);
        if (lagged) { nparts = (long long)g.x * g.y; rc = ensure_partials(h, nparts); if (rc) return rc; p.partials = h->partials; }
        if (odd) { DISPATCH_COLL(true, (launch_scalar<COLL, ODD>(h, p, general, g))) } else { DISPATCH_COLL(false, (launch_scalar<COLL, ODD>(h, p, general, g))) }
        h->launches++;
    } else {
        int threads; dim3 gv = vec_grid(h, threads);
        const long long nvb = (long long)gv.x * gv.y;
        if (lagged) { nparts = nvb + ngen; rc = ensure_partials(h, nparts); if (rc) { h->stream = main; return rc; } p.partials = h->partials; }
        if (ngen > 0) {
            if (side && !forked) { rc = fork_side(h, main); if (rc) return rc; forked = true; }      // the moments pre-pass came first, on the main stream
            Params pg = p; pg.segmask = nullptr; pg.gen_list = h->gen_list;
            if (lagged) pg.partials = h->partials + 3 * nvb;
            dim3 g(ngen, 1);
            if (odd) { DISPATCH_COLL(true, (launch_scalar<COLL, ODD>(h, pg, true, g))) } else { DISPATCH_COLL(false, (launch_scalar<COLL, ODD>(h, pg, true, g))) }
            h->launches++;
        }
        if (forked) h->stream = main;
        if (general) p.segmask = h->segmask;
        if (odd) { DISPATCH_COLL(true, (launch_vec<COLL, ODD>(h, p))) } else { DISPATCH_COLL(false, (launch_vec<COLL, ODD>(h, p))) }
        h->launches++;
        if (forked) { rc = join_side(h, main); if (rc) return rc; }
    }
    if (lagged) {
        reduce_partials(h, nparts);
        if (h->cfg.world == 1) h->avg_for_ts = t + 1;
    }
    h->timestep = t;
    if (want_macros) h->macros_ts = t;
    return LBM_OK;
}

static int prepare_resources(lbm_handle* h, bool want_macros);

// ---- bodies across slab faces with the halo coupling (the peer-mapped coupling needs none of these)
extern "C" int lbm_ibm_exchange_floats(lbm_handle* h, int64_t* out) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    *out = (h->cfg.world > 1 && !h->direct()) ? (int64_t)IBM_MAIL * h->nall : 0;
    return LBM_OK;
}
extern "C" int lbm_ibm_pack(lbm_handle* h, float* d_buf) {
    if (!h || !d_buf) return fail(LBM_ERR_INVALID, "NULL argument");
    if (h->cfg.world == 1 || h->nall == 0) return LBM_OK;
    if (h->direct()) return fail(LBM_ERR_STATE, "peer-mapped slabs exchange the IBM node states themselves");
    CU(cudaSetDevice(h->cfg.device));
    const int t = h->timestep + 1;
    Params p = make_params(h, t);
    CU(cudaMemsetAsync(d_buf, 0, (size_t)IBM_MAIL * h->nall * sizeof(float), h->stream));      // nodes other slabs own: x + 0 = x in the all-reduce
    launch_nbr_gather(h, p, t);
    if (h->ibm_count) {
        IbmData d = ibm_data(h);
        d.peer_mail[0] = d.peer_mail[1] = nullptr; d.peer_flag[0] = d.peer_flag[1] = nullptr;
        if (t & 1) ibm_gather_kernel<true><<<1, 1024, 0, h->stream>>>(p, d, d_buf, 0ull); else ibm_gather_kernel<false><<<1, 1024, 0, h->stream>>>(p, d, d_buf, 0ull);
        h->launches++;
    }
    CU(cudaGetLastError());
    return LBM_OK;
}
extern "C" int lbm_ibm_unpack(lbm_handle* h, const float* d_buf) {
    if (!h || !d_buf) return fail(LBM_ERR_INVALID, "NULL argument");
    if (h->cfg.world == 1 || h->nall == 0) return LBM_OK;
    if (h->direct()) return fail(LBM_ERR_STATE, "peer-mapped slabs exchange the IBM node states themselves");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->ibm_mail, d_buf, (size_t)IBM_MAIL * h->nall * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    h->ibm_mail_for_ts = h->timestep + 1;
    h->pre_for_ts = -1;
    return LBM_OK;
}

extern "C" int lbm_adapter_prepass(lbm_handle* h) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    if (h->cfg.collision != LBM_CM_OPTIMAL) return LBM_OK;
    CU(cudaSetDevice(h->cfg.device));
    { int rc0 = prepare_resources(h, false); if (rc0) return rc0; }
    if (h->direct() && h->ibm_count) { int rc0 = check_ibm_coverage(h); if (rc0) return rc0; }
    if (h->direct()) {      // the pre-pass already reads the neighbours' edge rows
        wait_neighbours_kernel<<<1, 1, 0, h->stream>>>(h->sync_flags, h->peer[0].attached, h->peer[1].attached, (unsigned long long)h->timestep, h->sync_timeout);
        h->launches++;
    }
    int rc = pre_passes(h, h->timestep + 1, true); if (rc) return rc;
    CU(cudaGetLastError());
    return LBM_OK;
}

// Everything a step may allocate lazily (cudaMalloc synchronises the whole device) is settled here, before the first
// handshake kernel of a call is enqueued: a wait kernel spinning on one handle's stream while another handle of the same
// process sits in cudaMalloc would dead-lock until the handshake timeout.
static int prepare_resources(lbm_handle* h, bool want_macros) {
    int rc;
    if (use_vec(h) && is_general(h)) { rc = ensure_segments(h); if (rc) return rc; }
    if (want_macros) { rc = ensure_macros(h); if (rc) return rc; }
    if (h->cfg.collision == LBM_CM_OPTIMAL) {
        dim3 g = grid_of(h);
        long long n = (long long)g.x * g.y;
        if (use_vec(h)) { int th; dim3 gv = vec_grid(h, th); n = std::max(n, (long long)gv.x * gv.y + (is_general(h) ? h->gen_count : 0)); }
        rc = ensure_partials(h, n); if (rc) return rc;
    }
    return LBM_OK;
}

// ------------------------------------------------------------------ CUDA-graph replay of step pairs (launch-bound grids)
constexpr int GRAPH_PAIRS = 8;

// everything a captured step bakes into its kernel arguments; a changed key re-captures
static std::string graph_key(lbm_handle* h, int parity) {
    std::string k;
    auto add = [&k](const auto& v) { k.append(reinterpret_cast<const char*>(&v), sizeof(v)); };
    const Params p = make_params(h, parity);
    for (int q = 0; q < Q; q++) { add(p.A[q]); add(p.S[q]); }
    add(p.A0[0]); add(p.A0[1]); add(p.nx); add(p.ny); add(p.y0); add(p.nyl); add(p.px); add(p.py); add(p.wrap_y); add(p.quirks); add(p.coll);
    add(p.flags); add(p.omega); add(p.u_max); add(p.fx); add(p.fy); add(p.force_plane); add(p.ring); add(p.perim);
    add(p.nbr_nodes); add(p.nbr_g); add(p.nbr_count); add(p.ibm_nodes); add(p.ibm_force); add(p.ibm_count); add(p.avg); add(p.nsx); add(p.plane);
    const IbmData d = ibm_data(h);
    add(d.np); add(d.ss); add(d.sten_idx); add(d.sten_w); add(d.row); add(d.csr_k); add(d.csr_w); add(d.rho); add(d.uprev); add(d.lagF); add(d.utarget);
    add(h->nbr_src); add(h->segmask); add(h->gen_list); add(h->gen_count); add(h->partials); add(h->n_partials); add(h->stage); add(h->sums);
    add(h->stream); add(h->side_stream); add(h->overlap); add(h->cfg.adapter_mode); add(parity);
    return k;
}

static bool graph_eligible(const lbm_handle* h) {
    if (h->cfg.world != 1) return false;
    if (h->graph_mode >= 0) return h->graph_mode == 1;
    return h->nloc <= (1ll << 22);
}

// captures 2*GRAPH_PAIRS steps starting at the current parity (nothing executes), leaves the host state untouched
static int capture_steps(lbm_handle* h, lbm_handle::StepGraph& g, const std::string& key) {
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    const int ts = h->timestep, avg = h->avg_for_ts, pre = h->pre_for_ts, nbrg = h->nbrg_for_ts, mts = h->macros_ts;
    const long long l0 = h->launches;
    cudaStream_t main = h->stream;
    CU(cudaStreamBeginCapture(main, cudaStreamCaptureModeThreadLocal));
    int rc = LBM_OK;
    for (int i = 0; i < 2 * GRAPH_PAIRS && rc == LBM_OK; i++) rc = one_step(h, false);
    cudaGraph_t graph = nullptr;
    h->stream = main;
    cudaError_t e = cudaStreamEndCapture(main, &graph);
    g.launches = h->launches - l0;
    g.d_avg = h->avg_for_ts - h->timestep; g.d_pre = h->pre_for_ts - h->timestep; g.d_nbrg = h->nbrg_for_ts - h->timestep;
    h->timestep = ts; h->avg_for_ts = avg; h->pre_for_ts = pre; h->nbrg_for_ts = nbrg; h->macros_ts = mts; h->launches = l0;
    if (rc != LBM_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_CUDA, std::string("stream capture failed: ") + cudaGetErrorString(e)); }
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { g.exec = nullptr; cudaGetLastError(); return fail(LBM_ERR_CUDA, std::string("graph instantiation failed: ") + cudaGetErrorString(e)); }
    g.key = key;
    return LBM_OK;
}

// runs as many whole graphs as fit into n steps; returns the number of steps done
static int run_graphs(lbm_handle* h, int n, int& done) {
    done = 0;
    if (n < 2 * GRAPH_PAIRS || !graph_eligible(h)) return LBM_OK;
    const int parity = (h->timestep + 1) & 1;
    // CM<2,OptimalAdapter>: the host decides per step whether the moments pre-pass runs (exact mode: unless the sums were handed
    // in; lagged mode: only while no previous step has produced them).  A graph is captured, and replayed, in the steady state only.
    if (h->cfg.collision == LBM_CM_OPTIMAL) {
        const bool have = h->avg_for_ts == h->timestep + 1;
        if (have != (h->cfg.adapter_mode == LBM_ADAPTER_LAGGED)) return LBM_OK;
    }
    // the legacy default stream (the header shim works on it, as the reference does) and the per-thread stream cannot be captured:
    // graphs are captured and replayed on the handle's own stream, ordered after / before the user's stream with two events
    cudaStream_t user = h->stream;
    const bool bridged = user == cudaStreamLegacy || user == cudaStreamPerThread;
    if (bridged) {
        for (auto& e : h->ev_bridge) if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->stream = h->own_stream;
    }
    lbm_handle::StepGraph& g = h->graph[parity];
    const std::string key = graph_key(h, parity);
    int rc = LBM_OK;
    if (!g.exec || g.key != key) rc = capture_steps(h, g, key);
    cudaError_t e = cudaSuccess;
    if (rc == LBM_OK && bridged) { e = cudaEventRecord(h->ev_bridge[0], user); if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_bridge[0], 0); }
    while (rc == LBM_OK && e == cudaSuccess && n - done >= 2 * GRAPH_PAIRS) {
        e = cudaGraphLaunch(g.exec, h->stream);
        if (e != cudaSuccess) break;
        h->timestep += 2 * GRAPH_PAIRS;
        h->launches += g.launches;
        h->avg_for_ts = h->timestep + g.d_avg; h->pre_for_ts = h->timestep + g.d_pre; h->nbrg_for_ts = h->timestep + g.d_nbrg;
        done += 2 * GRAPH_PAIRS;
    }
    if (rc == LBM_OK && e == cudaSuccess && bridged) { e = cudaEventRecord(h->ev_bridge[1], h->stream); if (e == cudaSuccess) e = cudaStreamWaitEvent(user, h->ev_bridge[1], 0); }
    h->stream = user;
    if (rc != LBM_OK) return rc;
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, std::string("graph replay failed: ") + cudaGetErrorString(e));
    return LBM_OK;
}

static int run_steps(lbm_handle* h, int n, bool macros_last) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    if (n < 0) return fail(LBM_ERR_INVALID, "nsteps < 0");
    // LBM_B200_NO_HANDSHAKE=1 (diagnosis only): skip the device-side step handshake; results are then racy
    static const bool no_handshake = getenv("LBM_B200_NO_HANDSHAKE") != nullptr;
    const bool direct = h->direct();
    const bool handshake = direct && !no_handshake;
    if (h->cfg.world > 1 && n > 1 && !direct) return fail(LBM_ERR_INVALID, "world > 1 without peer-mapped neighbours: step one at a time and exchange halos in between");
    if (h->cfg.world > 1 && n > 1 && h->cfg.collision == LBM_CM_OPTIMAL) return fail(LBM_ERR_INVALID, "OptimalAdapter on several slabs needs the all-reduce of the grid sums between steps: nsteps must be 1");
    CU(cudaSetDevice(h->cfg.device));
    if (n > 0) { int rc = prepare_resources(h, macros_last); if (rc) return rc; }
    if (n > 0 && direct && h->ibm_count) { int rc = check_ibm_coverage(h); if (rc) return rc; }
    int first = 0;
    if (!direct) { int rc = run_graphs(h, macros_last ? n - 1 : n, first); if (rc) return rc; }
    for (int i = first; i < n; i++) {
        if (handshake) {
            // every step touches cells the neighbours wrote (odd: their edge rows, even: what they stored into mine):
            // wait until both have completed step t-1, and tell them when step t is done
            wait_neighbours_kernel<<<1, 1, 0, h->stream>>>(h->sync_flags, h->peer[0].attached, h->peer[1].attached, (unsigned long long)h->timestep, h->sync_timeout);
            h->launches++;
        }
        int rc = one_step(h, macros_last && i == n - 1);
        if (rc) return rc;
        if (handshake) {
            signal_neighbours_kernel<<<1, 1, 0, h->stream>>>(h->peer[0].attached ? h->peer[0].flag : nullptr, h->peer[1].attached ? h->peer[1].flag : nullptr, (unsigned long long)h->timestep);
            h->launches++;
        }
    }
    CU(cudaGetLastError());
    return LBM_OK;
}

extern "C" int lbm_step(lbm_handle* h, int32_t n) { return run_steps(h, n, false); }
extern "C" int lbm_step_with_macroscopics(lbm_handle* h, int32_t n) { return run_steps(h, n, true); }

extern "C" int lbm_sync(lbm_handle* h) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->direct()) {
        int to = 0;
        CU(cudaMemcpy(&to, h->sync_timeout, sizeof(int), cudaMemcpyDeviceToHost));
        if (to) return fail(LBM_ERR_STATE, std::string(to == 2 ? "a neighbour slab did not post its IBM node states" : "a neighbour slab did not reach the previous time step") +
                            " within 10 s (peer-mapped handshake timed out at step " + std::to_string(h->timestep) + " of slab " + std::to_string(h->cfg.rank) + "); results are invalid");
    }
    return LBM_OK;
}

extern "C" int lbm_get_macroscopics_device(lbm_handle* h, const float** rho, const float** u) {
    if (!h || !rho || !u) return fail(LBM_ERR_INVALID, "NULL argument");
    if (h->macros_ts != h->timestep || !h->rho_out)
        return fail(LBM_ERR_STATE, "macroscopics of the current timestep were not produced: run the last step with lbm_step_with_macroscopics");
    *rho = h->rho_out; *u = (const float*)h->u_out;
    return LBM_OK;
}

extern "C" int lbm_get_macroscopics(lbm_handle* h, float* rho, float* u) {
    const float *dr, *du;
    int rc = lbm_get_macroscopics_device(h, &dr, &du);
    if (rc) return rc;
    if (!rho || !u) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(rho, dr, (size_t)h->nloc * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(u, du, (size_t)h->nloc * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

extern "C" int lbm_total_mass(lbm_handle* h, double* out) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemsetAsync(h->mass_acc, 0, 8, h->stream));
    Params p = make_params(h, h->timestep);
    if (h->timestep & 1) mass_kernel<true><<<grid_of(h), BX, 0, h->stream>>>(p, h->mass_acc);
    else mass_kernel<false><<<grid_of(h), BX, 0, h->stream>>>(p, h->mass_acc);
    h->launches++;
    CU(cudaMemcpyAsync(out, h->mass_acc, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

// ------------------------------------------------------------------ validation reductions on the device
static int need_current_macros(lbm_handle* h) {
    if (h->macros_ts != h->timestep || !h->rho_out)
        return fail(LBM_ERR_STATE, "macroscopics of the current timestep were not produced: run the last step with lbm_step_with_macroscopics");
    return LBM_OK;
}
static int ensure_val_stage(lbm_handle* h, long long n) {
    if (n <= h->val_stage_n) return LBM_OK;
    if (h->val_stage) { CU(cudaStreamSynchronize(h->stream)); cudaFree(h->val_stage); h->bytes -= 8 * h->val_stage_n; h->val_stage = nullptr; }
    CU(dmalloc(h, &h->val_stage, (size_t)n));
    h->val_stage_n = n;
    return LBM_OK;
}

static int error_sums(lbm_handle* h, const float2* d_ref, bool tg, float nu, float u0, float t, double out[2]) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    int rc = need_current_macros(h); if (rc) return rc;
    CU(cudaSetDevice(h->cfg.device));
    rc = ensure_val_stage(h, 2 * (RED_BLOCKS + 1)); if (rc) return rc;
    const int nb = (int)std::min<long long>(RED_BLOCKS, (h->nloc + 255) / 256);
    if (tg) error_sums_kernel<true><<<nb, 256, 0, h->stream>>>(h->u_out, nullptr, h->cfg.nx, h->cfg.ny, h->y0, h->nloc, nu, u0, t, h->val_stage);
    else error_sums_kernel<false><<<nb, 256, 0, h->stream>>>(h->u_out, d_ref, h->cfg.nx, h->cfg.ny, h->y0, h->nloc, 0.f, 0.f, 0.f, h->val_stage);
    error_sums_final_kernel<<<1, 256, 0, h->stream>>>(h->val_stage, nb, h->val_stage + 2 * RED_BLOCKS);
    h->launches += 2;
    CU(cudaMemcpyAsync(out, h->val_stage + 2 * RED_BLOCKS, 16, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return LBM_OK;
}

extern "C" int lbm_velocity_error_sums(lbm_handle* h, const float* d_u_ref, double out[2]) {
    if (!d_u_ref) return fail(LBM_ERR_INVALID, "NULL argument");
    return error_sums(h, (const float2*)d_u_ref, false, 0.f, 0.f, 0.f, out);
}
extern "C" int lbm_taylor_green_error_sums(lbm_handle* h, float nu, float u0, float t, double out[2]) {
    return error_sums(h, nullptr, true, nu, u0, t, out);
}

extern "C" int lbm_row_mean_velocity(lbm_handle* h, double* mean_ux, double* mean_uy) {
    if (!h || !mean_ux || !mean_uy) return fail(LBM_ERR_INVALID, "NULL argument");
    int rc = need_current_macros(h); if (rc) return rc;
    CU(cudaSetDevice(h->cfg.device));
    rc = ensure_val_stage(h, std::max<long long>(2 * (RED_BLOCKS + 1), 2ll * h->nyl)); if (rc) return rc;
    row_mean_kernel<<<h->nyl, 256, 0, h->stream>>>(h->u_out, h->cfg.nx, h->val_stage, h->val_stage + h->nyl);
    h->launches++;
    CU(cudaMemcpyAsync(mean_ux, h->val_stage, (size_t)h->nyl * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(mean_uy, h->val_stage + h->nyl, (size_t)h->nyl * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return LBM_OK;
}

// ------------------------------------------------------------------ checkpoint / restart
// File = CkptHeader + the population planes exactly as they sit in HBM (AA phase included, ghost rows included) + the edge
// ring.  Everything else a step reads is either configuration (re-created by the caller: flags, bodies, forces) or rebuilt
// from the populations at the start of the step (neighbour-BC gather, IBM, adapter sums in LBM_ADAPTER_EXACT).
struct CkptHeader {
    char magic[8];
    uint32_t version, header_bytes;
    int32_t nx, ny, rank, world, y0, nyl, nplanes, collision, quirks, periodic_x, periodic_y, adapter_mode;
    int32_t timestep, avg_for_ts, perim, pad;
    float avg[3]; float pad2;
    double sums[3];
    uint64_t pop_floats, ring_floats;
};
static const char kCkptMagic[8] = {'L', 'B', 'M', 'B', '2', '0', '0', 1};
constexpr size_t CKPT_CHUNK = (size_t)32 << 20;         // bytes per pinned staging buffer (two of them)

static CkptHeader ckpt_header(const lbm_handle* h) {
    CkptHeader k{};
    memcpy(k.magic, kCkptMagic, 8);
    k.version = 1; k.header_bytes = (uint32_t)sizeof(CkptHeader);
    k.nx = h->cfg.nx; k.ny = h->cfg.ny; k.rank = h->cfg.rank; k.world = h->cfg.world; k.y0 = h->y0; k.nyl = h->nyl;
    k.nplanes = h->nplanes; k.collision = h->cfg.collision; k.quirks = h->cfg.quirks;
    k.periodic_x = h->cfg.periodic_x; k.periodic_y = h->cfg.periodic_y; k.adapter_mode = h->cfg.adapter_mode;
    k.timestep = h->timestep; k.avg_for_ts = h->avg_for_ts; k.perim = h->perim;
    k.pop_floats = (uint64_t)h->plane * h->nplanes; k.ring_floats = (uint64_t)2 * h->perim * Q;
    return k;
}

extern "C" int lbm_checkpoint_bytes(lbm_handle* h, int64_t* out) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    CkptHeader k = ckpt_header(h);
    *out = (int64_t)(sizeof(CkptHeader) + 4 * (k.pop_floats + k.ring_floats));
    return LBM_OK;
}

// device <-> file through two pinned staging buffers: the copy of chunk i+1 runs while chunk i is written / read
static int stream_region(lbm_handle* h, FILE* fp, char* dev, size_t bytes, bool save, char* stage[2], cudaEvent_t ev[2]) {
    const size_t n = (bytes + CKPT_CHUNK - 1) / CKPT_CHUNK;
    auto len = [&](size_t i) { return std::min(CKPT_CHUNK, bytes - i * CKPT_CHUNK); };
    if (save) {
        for (size_t i = 0; i <= n; i++) {
            if (i < n) {
                CU(cudaMemcpyAsync(stage[i & 1], dev + i * CKPT_CHUNK, len(i), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaEventRecord(ev[i & 1], h->stream));
            }
            if (i > 0) {
                CU(cudaEventSynchronize(ev[(i - 1) & 1]));
                if (fwrite(stage[(i - 1) & 1], 1, len(i - 1), fp) != len(i - 1)) return fail(LBM_ERR_STATE, "checkpoint: short write");
            }
        }
    } else {
        for (size_t i = 0; i < n; i++) {
            if (i >= 2) CU(cudaEventSynchronize(ev[i & 1]));       // the copy that last used this buffer
            if (fread(stage[i & 1], 1, len(i), fp) != len(i)) return fail(LBM_ERR_STATE, "checkpoint: file is truncated");
            CU(cudaMemcpyAsync(dev + i * CKPT_CHUNK, stage[i & 1], len(i), cudaMemcpyHostToDevice, h->stream));
            CU(cudaEventRecord(ev[i & 1], h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
    }
    return LBM_OK;
}

static int checkpoint_io(lbm_handle* h, const char* path, bool save) {
    if (!h || !path) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    FILE* fp = fopen(path, save ? "wb" : "rb");
    if (!fp) return fail(LBM_ERR_INVALID, std::string("checkpoint: cannot open ") + path);
    CkptHeader k = ckpt_header(h);
    int rc = LBM_OK;
    if (save) {
        if (cudaMemcpy(k.avg, h->avg, 12, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(k.sums, h->sums, 24, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(LBM_ERR_CUDA, "checkpoint: reading the adapter means failed");
        else if (fwrite(&k, sizeof(k), 1, fp) != 1) rc = fail(LBM_ERR_STATE, "checkpoint: short write");
    } else {
        CkptHeader f{};
        if (fread(&f, sizeof(f), 1, fp) != 1 || memcmp(f.magic, kCkptMagic, 8) != 0 || f.version != 1 || f.header_bytes != sizeof(CkptHeader))
            rc = fail(LBM_ERR_INVALID, "checkpoint: not a checkpoint file of this engine version");
        else if (f.nx != k.nx || f.ny != k.ny || f.rank != k.rank || f.world != k.world || f.y0 != k.y0 || f.nyl != k.nyl || f.nplanes != k.nplanes ||
                 f.periodic_x != k.periodic_x || f.periodic_y != k.periodic_y || f.pop_floats != k.pop_floats || f.ring_floats != k.ring_floats)
            rc = fail(LBM_ERR_INVALID, "checkpoint: written for a different grid / slab decomposition / quirk set (nx, ny, rank, world, periodicity and LBM_QK_D1_STALE_F0 must match)");
        else k = f;
    }
    char* stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (rc == LBM_OK) {
        for (int i = 0; i < 2 && rc == LBM_OK; i++)
            if (cudaHostAlloc((void**)&stage[i], CKPT_CHUNK, cudaHostAllocDefault) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess)
                rc = fail(LBM_ERR_CUDA, "checkpoint: pinned staging allocation failed");
    }
    if (rc == LBM_OK) rc = stream_region(h, fp, (char*)h->pop, (size_t)k.pop_floats * 4, save, stage, ev);
    if (rc == LBM_OK) rc = stream_region(h, fp, (char*)h->ring, (size_t)k.ring_floats * 4, save, stage, ev);
    for (int i = 0; i < 2; i++) { if (stage[i]) cudaFreeHost(stage[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
    if (fclose(fp) != 0 && rc == LBM_OK && save) rc = fail(LBM_ERR_STATE, "checkpoint: close failed");
    if (rc != LBM_OK || save) return rc;
    // restart: the scalar state of the handle
    h->timestep = k.timestep; h->avg_for_ts = k.avg_for_ts; h->pre_for_ts = -1; h->macros_ts = -1;
    CU(cudaMemcpy(h->avg, k.avg, 12, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->sums, k.sums, 24, cudaMemcpyHostToDevice));
    // peer-mapped neighbours restart from the same step: all slabs must have loaded (host barrier) before any of them steps
    const unsigned long long ts = (unsigned long long)k.timestep;
    unsigned long long fl[4] = {ts, ts, ts, ts};
    CU(cudaMemcpy(h->sync_flags, fl, 32, cudaMemcpyHostToDevice));
    h->ibm_mail_for_ts = -1; h->nbrg_for_ts = -1;
    CU(cudaMemset(h->sync_timeout, 0, sizeof(int)));
    return LBM_OK;
}

extern "C" int lbm_checkpoint_write(lbm_handle* h, const char* path) { return checkpoint_io(h, path, true); }
extern "C" int lbm_checkpoint_read(lbm_handle* h, const char* path) { return checkpoint_io(h, path, false); }

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

extern "C" int lbm_info(lbm_handle* h, lbm_info_t* o) {
    if (!h || !o) return fail(LBM_ERR_INVALID, "NULL argument");
    memset(o, 0, sizeof(*o));
    o->nx = h->cfg.nx; o->ny = h->cfg.ny; o->y0 = h->y0; o->ny_local = h->nyl; o->rank = h->cfg.rank; o->world = h->cfg.world;
    o->timestep = h->timestep; o->num_markers = h->np_total; o->num_ibm_nodes = h->ibm_count; o->num_neighbour_bc_nodes = h->nbr_count;
    o->device_bytes = h->bytes; o->bytes_per_cell = (double)h->bytes / (double)h->nloc; o->kernel_launches = h->launches;
    return LBM_OK;
}

// ------------------------------------------------------------------ slab halos
// Odd steps read A[opp q] of the neighbour's edge row for the three q that enter this slab and write A[q]
// of it for the three q that leave.  side 0 (lower y): entering q = 2,5,6 -> slots 4,7,8; side 1: slots 2,5,6.
static const int kSlots[2][3] = {{4, 7, 8}, {2, 5, 6}};

extern "C" int lbm_next_step_needs_halo(lbm_handle* h) {
    if (!h) return 0;
    return (h->cfg.world > 1 && !h->direct() && ((h->timestep + 1) & 1)) ? 1 : 0;
}

static bool has_neighbour(const lbm_handle* h, int side) {
    if (h->cfg.world == 1) return false;
    if (h->cfg.periodic_y) return true;
    return side == 0 ? h->cfg.rank > 0 : h->cfg.rank < h->cfg.world - 1;
}

// row index (in plane rows, ghost offset included) of: own edge row / ghost row on a side
static int own_edge_row(const lbm_handle* h, int side) { return side == 0 ? 1 : h->nyl; }
static int ghost_row(const lbm_handle* h, int side) { return side == 0 ? 0 : h->nyl + 1; }

static int copy_rows(lbm_handle* h, int row, const int slots[3], float* buf, bool to_buf) {
    const size_t nx = h->cfg.nx;
    for (int i = 0; i < 3; i++) {
        float* src = h->pop + (size_t)slots[i] * h->plane + (size_t)row * nx;
        if (to_buf) CU(cudaMemcpyAsync(buf + i * nx, src, nx * 4, cudaMemcpyDeviceToDevice, h->stream));
        else CU(cudaMemcpyAsync(src, buf + i * nx, nx * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    return LBM_OK;
}

// pre: the OWNER of an edge row packs the slots its neighbour on `side` will read.
// The neighbour above me (side 1) reads my top row's slots 4,7,8 (it is its "side 0" data) and vice versa.
extern "C" int lbm_halo_pack_pre(lbm_handle* h, int side, float* buf) {
    if (!h || !buf || side < 0 || side > 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (!has_neighbour(h, side)) return LBM_OK;
    CU(cudaSetDevice(h->cfg.device));
    return copy_rows(h, own_edge_row(h, side), kSlots[1 - side], buf, true);
}
extern "C" int lbm_halo_unpack_pre(lbm_handle* h, int side, const float* buf) {
    if (!h || !buf || side < 0 || side > 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (!has_neighbour(h, side)) return LBM_OK;
    CU(cudaSetDevice(h->cfg.device));
    return copy_rows(h, ghost_row(h, side), kSlots[side], (float*)buf, false);
}
// post: what this slab wrote into its ghost row on `side` goes back into the neighbour's edge row.
extern "C" int lbm_halo_pack_post(lbm_handle* h, int side, float* buf) {
    if (!h || !buf || side < 0 || side > 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (!has_neighbour(h, side)) return LBM_OK;
    CU(cudaSetDevice(h->cfg.device));
    return copy_rows(h, ghost_row(h, side), kSlots[side], buf, true);
}
extern "C" int lbm_halo_unpack_post(lbm_handle* h, int side, const float* buf) {
    if (!h || !buf || side < 0 || side > 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (!has_neighbour(h, side)) return LBM_OK;
    CU(cudaSetDevice(h->cfg.device));
    return copy_rows(h, own_edge_row(h, side), kSlots[1 - side], (float*)buf, false);
}

// ------------------------------------------------------------------ peer-mapped neighbours (NVLink / same-device)
extern "C" int lbm_peer_export(lbm_handle* h, void* out) {
    if (!h || !out) return fail(LBM_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    PeerDesc d{};
    CU(cudaIpcGetMemHandle(&d.ipc, h->pop));
    d.pid = (long long)getpid(); d.raw = (unsigned long long)(uintptr_t)h->pop; d.plane = (long long)h->plane;
    d.flags_off = (long long)((char*)h->sync_flags - (char*)h->pop);
    d.mail_off = h->ibm_mail ? (long long)((char*)h->ibm_mail - (char*)h->pop) : -1;
    d.nx = h->cfg.nx; d.nyl = h->nyl; d.device = h->cfg.device; d.rank = h->cfg.rank; d.y0 = h->y0; d.mail_nodes = h->mail_nodes;
    memset(out, 0, LBM_PEER_DESC_BYTES);
    memcpy(out, &d, sizeof(d));
    return LBM_OK;
}

extern "C" int lbm_peer_attach(lbm_handle* h, int side, const void* desc) {
    if (!h || !desc || side < 0 || side > 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (!has_neighbour(h, side)) return fail(LBM_ERR_INVALID, "this slab has no neighbour on that side");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    PeerDesc d; memcpy(&d, desc, sizeof(d));
    if (d.nx != h->cfg.nx) return fail(LBM_ERR_INVALID, "neighbour slab has a different nx");
    if (d.mail_nodes != h->mail_nodes) return fail(LBM_ERR_INVALID, "neighbour slab has a different ibm_mailbox_nodes");
    lbm_handle::Peer& P = h->peer[side];
    if (P.attached) return fail(LBM_ERR_STATE, "side already attached");
    char* base = nullptr;
    if (d.pid == (long long)getpid()) {
        base = (char*)(uintptr_t)d.raw;                       // same process: the pointer is valid as it is
        if (d.device != h->cfg.device) {
            int can = 0; CU(cudaDeviceCanAccessPeer(&can, h->cfg.device, d.device));
            if (!can) return fail(LBM_ERR_INVALID, "devices cannot access each other's memory");
            cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(LBM_ERR_CUDA, cudaGetErrorString(e));
            cudaGetLastError();
        }
    } else {
        const lbm_handle::Peer& O = h->peer[1 - side];
        if (O.ipc_base && O.base == (float*)O.ipc_base && memcmp(&d.ipc, &h->peer_ipc[1 - side], sizeof(d.ipc)) == 0) base = (char*)O.ipc_base;   // world == 2, periodic: both faces touch the same slab
        else CU(cudaIpcOpenMemHandle((void**)&base, d.ipc, cudaIpcMemLazyEnablePeerAccess));
        P.ipc_base = base;
        h->peer_ipc[side] = d.ipc;
    }
    P.base = (float*)base; P.plane = d.plane;
    // side 0 (lower neighbour): its top edge row = local row nyl-1 -> plane row nyl; side 1: its local row 0 -> plane row 1
    P.off = (side == 0 ? (long long)d.nyl : 1ll) * d.nx;
    // I am the neighbour's upper peer when it is below me: it waits on its flags[1]
    P.flag = reinterpret_cast<unsigned long long*>(base + d.flags_off) + (side == 0 ? 1 : 0);     // its IBM stage counter: P.flag + 2
    P.mail = d.mail_off >= 0 ? reinterpret_cast<float*>(base + d.mail_off) : nullptr;
    P.y0 = d.y0; P.nyl = d.nyl;
    P.attached = true;
    h->pre_for_ts = -1;
    return prepare_resources(h, false);
}

extern "C" int lbm_peer_detach(lbm_handle* h) {
    if (!h) return fail(LBM_ERR_INVALID, "NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    for (int sd = 0; sd < 2; sd++) {
        if (h->peer[sd].ipc_base && !(sd == 1 && h->peer[0].ipc_base == h->peer[1].ipc_base)) cudaIpcCloseMemHandle(h->peer[sd].ipc_base);
        h->peer[sd] = lbm_handle::Peer();
    }
    return LBM_OK;
}

extern "C" int lbm_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes <= 0) return fail(LBM_ERR_INVALID, "bad argument");
    CU(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return LBM_OK;
}
extern "C" int lbm_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return LBM_OK;
}
