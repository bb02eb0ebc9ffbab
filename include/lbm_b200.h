/*
 * lbm_b200.h — C ABI of the B200-native D2Q9 lattice-Boltzmann engine.
 *
 * The reference (Carabalone/cuda-lbm) has no FFI: its boundary is the C++ template class
 * LBM<2> (src/core/lbm.cuh:33-382) driven by src/main.cu:72-153 and parameterised by a
 * ScenarioTrait (src/scenarios/scenario.cuh:22-78).  Each entry point below replaces one of
 * the reference's host methods; the reference interface it stands for is cited per function.
 * The C++ header shim in include/cuda-lbm/ (LBM<2>, ScenarioTrait, ...) is a thin layer over
 * exactly these calls.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every call returns LBM_OK (0) or a
 * negative error code and never exits the process (the reference calls exit(99) on any CUDA
 * error, src/util/utility.cu:4-12); lbm_last_error() returns the message of the last failure
 * on the calling thread.  Host arrays use the reference's layouts: rho[node], u[node*2+c]
 * (AoS), f[node*9+q] (AoS), node = y*NX + x (src/core/lbm_constants.cuh:344-350).
 * One handle = one y-slab on one GPU.  A handle is not thread-safe; different handles are
 * independent (no global mutable state, unlike the reference's __constant__ symbols).
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_OK 0
#define LBM_ERR_INVALID -1   /* bad argument / unsupported configuration */
#define LBM_ERR_CUDA -2      /* a CUDA runtime call failed */
#define LBM_ERR_STATE -3     /* call not valid in the current state */

/* collision operators: the template argument of LBM<dim>::collide<Op>() (src/core/collision/collision.cuh:47) */
#define LBM_BGK 0            /* BGK<2>                 src/core/collision/BGK/BGK.cuh:13-51   */
#define LBM_MRT 1            /* MRT<2>                 src/core/collision/MRT/MRT.cu:4-76     */
#define LBM_CM 2             /* CM<2,NoAdapter>        src/core/collision/CM/CM.cuh:27-139    */
#define LBM_CM_OPTIMAL 3     /* CM<2,OptimalAdapter>   src/core/collision/adapters.cuh:48-111 */

/* reference-compatibility switches: bit set = reproduce the reference's
 * behaviour, bit clear = repaired behaviour.  LBM_QK_REFERENCE is the parity configuration. */
#define LBM_QK_D1_STALE_F0 1   /* src/core/streaming/streaming.cu:9 — rest population lags by two steps */
#define LBM_QK_D2_MRT_ROWS 2   /* src/core/collision/MRT/MRT.cu:14-22 — force moments rows 4/5 swapped */
#define LBM_QK_D3_ZOUHE_RHO 4  /* src/functors/boundaryConditions/zouHeInflow.cuh:15-16 */
#define LBM_QK_D7_IBM_CLIP 8   /* src/IBM/IBM_impl.cuh:20-24,41-44 — positive-only clipping */
#define LBM_QK_D8_IBM_2X2 16   /* src/IBM/IBM_impl.cu:19-21,132-134 — 2x2 stencil */
#define LBM_QK_D11_BB_RAW 32   /* src/functors/boundaryConditions/bbDomainBoundary.cuh:35-36 */
#define LBM_QK_D9_IBM_ZERO_TARGET 64 /* src/IBM/IBM_impl.cuh:15 — direct forcing targets u = 0; IBMBody::velocities are dead data */
#define LBM_QK_REFERENCE 127
#define LBM_QK_FIXED 0

/* how CM<2,OptimalAdapter> obtains the grid means of rho, rho|u|, |Pi| (src/core/macroscopics/macroscopics.cuh:51-120,161-177) */
#define LBM_ADAPTER_EXACT 0    /* moments pre-pass over the current post-stream state (reference semantics, +36 B/cell) */
#define LBM_ADAPTER_LAGGED 1   /* means of the previous step, accumulated by the fused kernel itself (72 B/cell) */

/* BC_flag values accepted by lbm_set_flags — src/core/lbm_constants.cuh:377-397 */
#define LBM_FLUID 0
#define LBM_BOUNCE_BACK 1
#define LBM_ZOU_HE_TOP 2
#define LBM_ZOU_HE_LEFT 3
#define LBM_CYLINDER 6
#define LBM_ZG_OUTFLOW 7
#define LBM_PRESSURE_OUTLET 8
#define LBM_REGULARIZED_INLET_TOP 9
#define LBM_REGULARIZED_BOUNCE_BACK 11
#define LBM_REGULARIZED_BOUNCE_BACK_CORNER 12

typedef struct lbm_handle lbm_handle;

/* Everything the reference fixes at compile time (src/defines.hpp:4-61, streaming.cuh:8-11) or
 * uploads once in LBM::send_consts (src/core/lbm.cuh:63-82) is run-time data here. */
typedef struct lbm_config {
    int32_t nx, ny;             /* global grid: NX, NY (src/defines.hpp:20-66) */
    int32_t periodic_x;         /* PERIODIC_X (src/core/streaming/streaming.cuh:9) */
    int32_t periodic_y;         /* PERIODIC_Y (:10) */
    int32_t collision;          /* LBM_BGK .. LBM_CM_OPTIMAL: Scenario::CollisionOp */
    float viscosity;            /* Scenario::viscosity; tau = 3 nu + 1/2, omega = 1/tau (lbm_constants.cuh:365-367, scenario.cuh:35-37) */
    float S[9];                 /* Scenario::S — relaxation rates in the row order of the chosen operator */
    float u_max;                /* Scenario::u_max — lid / inlet speed of the BC functors (boundaries.cuh:40,44,70) */
    float force_x, force_y;     /* uniform body force: what Scenario::init().apply_forces writes each step (macroscopics.cuh:13-48) */
    int32_t quirks;             /* LBM_QK_* mask */
    int32_t adapter_mode;       /* LBM_ADAPTER_* (only read for LBM_CM_OPTIMAL) */
    int32_t device;             /* CUDA device ordinal */
    int32_t rank, world;        /* y-slab index and slab count (1 = whole domain on this GPU) */
    int32_t ibm_mailbox_nodes;  /* world > 1: capacity of the IBM node mailbox in lattice nodes (0 = 65536; 20 B each) */
    int32_t reserved[3];        /* must be zero */
} lbm_config;

typedef struct lbm_info_t {
    int32_t nx, ny;             /* global grid */
    int32_t y0, ny_local;       /* slab rows [y0, y0+ny_local) */
    int32_t rank, world;
    int32_t timestep;           /* LBM::timestep (src/core/lbm.cuh:90) */
    int32_t num_markers;        /* IBMManager::num_points */
    int32_t num_ibm_nodes;      /* lattice nodes touched by marker stencils */
    int32_t num_neighbour_bc_nodes; /* nodes whose BC reads an interior neighbour (ZG / pressure / corner) */
    int64_t device_bytes;       /* device memory owned by the handle */
    double bytes_per_cell;      /* device_bytes / (nx*ny_local) */
    int64_t kernel_launches;    /* kernels launched by this handle since creation */
} lbm_info_t;

/* lbm_config with the reference's defaults (ScenarioTrait base: nu=1/6, u_max=0.1, BGK, quirks=REFERENCE). */
int lbm_default_config(lbm_config* cfg);

/* LBM<2>::allocate<Scenario>() — src/core/lbm.cuh:92-125.  Single in-place SoA population buffer
 * (36 B/cell, +4 with LBM_QK_D1_STALE_F0, +1 flag byte) instead of the reference's 144 B/cell. */
int lbm_create(const lbm_config* cfg, lbm_handle** out);

/* LBM<2>::~LBM / free() — src/core/lbm.cuh:127-140,379-381 */
int lbm_destroy(lbm_handle* h);

/* Use an externally owned CUDA stream (cudaStream_t) for all work of this handle; NULL = the
 * handle's own stream.  (The reference uses the default stream and synchronises after every launch.) */
int lbm_set_stream(lbm_handle* h, void* cuda_stream);

/* LBM<2>::setup_boundary_flags — src/core/boundaries/boundaries.cuh:170-197.  flags = the Boundary functor
 * evaluated by the caller for every node of the GLOBAL grid (nx*ny int32, host memory). */
int lbm_set_flags(lbm_handle* h, const int32_t* flags);

/* Optional per-node body force (AoS [node*2+c], global grid, host memory) for scenarios whose
 * apply_forces is not uniform; NULL returns to the uniform cfg.force_x/force_y. */
int lbm_set_force_field(lbm_handle* h, const float* force_aos);
/* Same, from device memory holding only this slab's rows (ny_local*nx float2, AoS): what the header shim's
 * reset_forces<Scenario>() produces by running Init::apply_forces on the device (src/core/macroscopics/macroscopics.cuh:13-48). */
int lbm_set_force_field_device(lbm_handle* h, const float* d_force_aos_local);
int lbm_set_body_force(lbm_handle* h, float fx, float fy);

/* Scenario::add_bodies() + IBMManager<2>::init_and_dispatch — src/IBM/IBMManager.cuh:54-109.
 * points = IBMBody::points, AoS [i*2+c], host memory, global lattice coordinates.  The marker->node
 * stencil structure is built on the GPU.  The caller keeps ownership of `points`.  One call per body, any number of
 * bodies.  With several slabs EVERY slab is given EVERY body, in the same order: a slab works on the bodies whose
 * stencils reach into its rows (bodies that share lattice nodes count as one) and ignores the rest. */
int lbm_add_body(lbm_handle* h, const float* points_aos, int32_t num_points);
/* IBMBody::velocities (src/IBM/IBMBody.cuh:33-45), AoS [i*2+c], for body `body` (index in lbm_add_body order): the velocity its
 * markers force the fluid towards.  The reference uploads them and then targets a literal 0 (src/IBM/IBM_impl.cuh:15, SURVEY
 * A-D9), which LBM_QK_D9_IBM_ZERO_TARGET reproduces: they take effect only with that bit clear.  NULL = zero. */
int lbm_set_body_velocities(lbm_handle* h, int32_t body, const float* velocities_aos);
/* Move body `body`: new marker positions (same count as added).  The marker->node structure is rebuilt on the GPU
 * (a few small kernels and sorts over O(markers); synchronises).  No reference counterpart (its bodies are static). */
int lbm_move_body(lbm_handle* h, int32_t body, const float* points_aos);

/* LBM<2>::init<Scenario>() — src/core/init/init.cuh:45-86: rho,u = the Init functor evaluated for every
 * node of the GLOBAL grid (host memory); populations are set to f_eq(rho,u), timestep = 0. */
int lbm_init_fields(lbm_handle* h, const float* rho, const float* u_aos);
/* Same, from host memory holding only this slab's rows (ny_local*nx nodes); asynchronous when the memory is pinned. */
int lbm_init_fields_local(lbm_handle* h, const float* rho_local, const float* u_aos_local);
/* Same, from device memory holding only this slab's rows (ny_local*nx nodes). */
int lbm_init_fields_device(lbm_handle* h, const float* d_rho_local, const float* d_u_aos_local);
/* TaylorGreenInit evaluated on the device (src/scenarios/taylorGreen/taylorGreenFunctors.cuh:25-47);
 * u0 is the functor's u_max member (already divided by SCALE). */
int lbm_init_taylor_green(lbm_handle* h, float nu, float u0);

/* Load / read the post-collision populations in the reference's AoS layout (what d_f holds after
 * collide(), src/core/collision/collision.cuh:62).  GLOBAL grid, host memory.  f_back may be NULL (= f);
 * it only matters for the reference's stale slots (A-D1, undelivered edge slots).  Valid when timestep is even. */
int lbm_set_populations(lbm_handle* h, const float* f_aos, const float* f_back_aos);
int lbm_get_populations(lbm_handle* h, float* f_aos);

/* One call = nsteps iterations of the reference's time loop body, src/main.cu:96-114
 * (increase_ts, stream, swap_buffers, apply_boundaries, uncorrected_macroscopics, reset_forces,
 * ibm_step, correct_macroscopics, compute_equilibrium, collide) fused into one kernel per step.
 * Asynchronous: returns after enqueueing.  For world > 1 the caller exchanges halos between steps
 * (lbm_halo_*), so nsteps must be 1. */
int lbm_step(lbm_handle* h, int32_t nsteps);
/* Same, and the LAST of the nsteps also stores rho and u as the reference's d_rho/d_u hold them after that
 * step (uncorrected_macroscopics + correct_macroscopics, src/core/macroscopics/macroscopics.cu:5-38,99-110;
 * +12 B/cell on that step only).  Required before lbm_get_macroscopics. */
int lbm_step_with_macroscopics(lbm_handle* h, int32_t nsteps);

/* One driver segment of the reference in one call (src/main.cu:77-147: init<Scenario>(), n iterations of the time loop,
 * update_macroscopics()): populations from rho / u in host memory (this slab's rows, as lbm_init_fields_local), nsteps time
 * steps, rho / u of the last step back into host memory (as lbm_get_macroscopics).  Same results, bit for bit, as those three
 * calls.  When the slab is the whole periodic domain on the vectorised path, the slab is stepped in row bands in a skewed order
 * (band j covers the rows [r0 - t, r1 - t) at step t and runs all its steps as soon as it has arrived; needs ny_local >= 4 nsteps + 16,
 * otherwise the three calls run) so that the host->device copy, the kernels and the device->host copy of different bands overlap (pinned host memory,
 * lbm_host_alloc, for full effect).  The output arrays may be the input arrays.  Blocks until the output is complete.
 * Several slabs: supported with peer-mapped neighbours on both faces; every slab makes the call (one thread or process each), after
 * all of them have finished their previous work (lbm_sync + barrier, as for lbm_init_*); the slab faces are synchronised level by
 * level on the device. */
int lbm_run_from_host(lbm_handle* h, const float* rho_local, const float* u_aos_local, int32_t nsteps, float* rho_out_local, float* u_out_local);

/* cudaDeviceSynchronize() of the reference's methods, once. */
int lbm_sync(lbm_handle* h);

/* LBM<2>::update_macroscopics() — src/core/lbm.cuh:148-154: rho[node], u[node*2+c] as the reference defines
 * d_rho/d_u (moments of the post-stream, post-BC state of the last step, u including F/2rho).
 * Output covers this slab's rows only: rho[ny_local*nx], u[ny_local*nx*2], host memory. */
int lbm_get_macroscopics(lbm_handle* h, float* rho_local, float* u_aos_local);
/* Allocate the rho / u planes now (12 B/cell) so that device-side initialisers also record the initial fields. */
int lbm_reserve_macroscopics(lbm_handle* h);
/* Same, leaving the result in device memory owned by the handle (valid until the next step). */
int lbm_get_macroscopics_device(lbm_handle* h, const float** d_rho_local, const float** d_u_aos_local);

/* rho / u of the current step for callers that enqueued it with plain lbm_step: rebuilt from the post-collision populations
 * (collision conserves mass and adds a known momentum), i.e. the reference's d_rho / d_u — which are always current there,
 * src/core/lbm.cuh:148-154 — up to fp32 round-off.  No-op when the step already stored them. */
int lbm_recover_macroscopics(lbm_handle* h);

/* Sum of all populations of this slab in fp64 (mass diagnostic; no reference counterpart). */
int lbm_total_mass(lbm_handle* h, double* out);
/* d_moment_avg — src/core/lbm.cuh:25-31: grid means of rho, rho|u|, |Pi| used by the last step (this slab's sums / global N). */
int lbm_moment_avg(lbm_handle* h, float out[3]);
/* CM<2,OptimalAdapter> on several slabs: lbm_adapter_prepass computes this slab's sums of rho, rho|u|, |Pi| for
 * the NEXT step (LBM_ADAPTER_EXACT); the caller all-reduces lbm_get_moment_sums over the slabs and hands the
 * global sums back with lbm_set_moment_sums before lbm_step.  With LBM_ADAPTER_LAGGED the sums come out of
 * the step itself and only the get / all-reduce / set part is needed. */
int lbm_adapter_prepass(lbm_handle* h);
/* 1 when the next lbm_step of a CM<2,OptimalAdapter> slab still needs global sums handed in (several slabs, host-side all-reduce):
 * always in LBM_ADAPTER_EXACT mode, and in LBM_ADAPTER_LAGGED mode for the first step after init / restart / lbm_set_populations. */
int lbm_adapter_sums_pending(lbm_handle* h);
int lbm_set_moment_sums(lbm_handle* h, const double sums[3]);
int lbm_get_moment_sums(lbm_handle* h, double sums[3]);

int lbm_info(lbm_handle* h, lbm_info_t* out);

/* ---- validation on the device --------------------------------------------------
 * The sums the reference's Scenario::compute_error methods form on the host from h_u (src/core/lbm.cuh:163-171),
 * taken on the device over this slab's rows in fp64 with a fixed summation order, so that a 32768^2 run is validated
 * without moving 8.6 GB of velocities to the host.  All three need the macroscopics of the current step
 * (lbm_step_with_macroscopics).  With several slabs the caller adds the per-slab results.
 *   out[0] = sum |u - u_ref|^2, out[1] = sum |u_ref|^2; the reference's metric is 100*sqrt(out[0]/out[1])
 *   (src/scenarios/taylorGreen/taylorGreenScenario.cuh:59-88). */
int lbm_velocity_error_sums(lbm_handle* h, const float* d_u_ref_aos_local, double out[2]);
/* Same against TaylorGreenValidation evaluated in place (src/scenarios/taylorGreen/taylorGreenFunctors.cuh:66-81);
 * u0 = the functor's u_max/SCALE, t = Scenario::t. */
int lbm_taylor_green_error_sums(lbm_handle* h, float nu, float u0, float t, double out[2]);
/* Mean of u over x for every row of this slab (ny_local doubles each, host memory): the inner loop of
 * PoiseuilleScenario::compute_error (src/scenarios/poiseuille/poiseuilleScenario.cuh:63-70). */
int lbm_row_mean_velocity(lbm_handle* h, double* mean_ux_rows, double* mean_uy_rows);

/* u[node*2+c] at n listed GLOBAL nodes (host arrays): what a centre-line validation reads from h_u — LidDrivenScenario::compute_error
 * compares 2 x 17 samples with the tables of Ghia et al. (src/scenarios/lidDrivenCavity/lidDrivenCavityScenario.cuh:88-157) — without
 * moving the whole field to the host.  Entries whose node belongs to another slab keep the caller's values: call it on every slab. */
int lbm_sample_velocity(lbm_handle* h, const int64_t* nodes, int32_t n, float* u_out_aos);

/* ---- checkpoint / restart (no reference counterpart) ------------------------------
 * lbm_checkpoint_write stores the populations of this slab as they sit in HBM, the edge ring, the time step and the
 * adapter means in one file (lbm_checkpoint_bytes long), streamed through pinned staging buffers; lbm_checkpoint_read
 * restores them into a handle created with the same grid, slab decomposition and LBM_QK_D1_STALE_F0 setting.  Flags,
 * bodies and forces are configuration: the caller sets them as for a fresh run.  The continued run is bit-identical
 * to the uninterrupted one.  Both calls synchronise.  Peer-mapped slabs: every slab reads its file, then a host
 * barrier, then the first step. */
int lbm_checkpoint_bytes(lbm_handle* h, int64_t* out);
int lbm_checkpoint_write(lbm_handle* h, const char* path);
int lbm_checkpoint_read(lbm_handle* h, const char* path);

/* ---- y-slab halo exchange (no reference counterpart) ----------------------------
 * Only the odd ("neighbour") steps of the in-place AA pattern touch the neighbour slab.  Before such a
 * step each rank needs 3*nx floats from each neighbour (lbm_halo_pack_pre on the owner ->
 * transport -> lbm_halo_unpack_pre here); after it, the 3*nx floats it wrote for each neighbour travel
 * back (lbm_halo_pack_post here -> transport -> lbm_halo_unpack_post on the owner).  Buffers are device
 * memory of 3*nx floats; side 0 = lower-y neighbour, 1 = upper-y neighbour.  lbm_next_step_needs_halo
 * tells whether the next lbm_step(h,1) is such a step. */
int lbm_next_step_needs_halo(lbm_handle* h);
int lbm_halo_pack_pre(lbm_handle* h, int side, float* d_buf);
int lbm_halo_unpack_pre(lbm_handle* h, int side, const float* d_buf);
int lbm_halo_pack_post(lbm_handle* h, int side, float* d_buf);
int lbm_halo_unpack_post(lbm_handle* h, int side, const float* d_buf);

/* ---- bodies whose stencils cross a slab face (no reference counterpart) -------------------
 * Every slab that owns part of a body runs the three direct-forcing iterations for the WHOLE body, redundantly and
 * with identical bits, on the states (rho, u*, F) of all its stencil nodes; each state is computed by the slab that
 * owns the node.  With the halo coupling the states travel like the halo rows: lbm_ibm_pack writes this slab's node
 * states for the NEXT step into a device buffer of lbm_ibm_exchange_floats floats (zeros for nodes other slabs own),
 * the caller all-reduces (sum) the buffer over ALL slabs, lbm_ibm_unpack hands it back, then lbm_step.  (After the
 * pre-step halo exchange; before lbm_adapter_prepass.)  With peer-mapped neighbours none of this is needed: the owners
 * store the states into the mailboxes of the slabs they have mapped over NVLink inside the pre-pass kernel — a body may then span
 * a slab and its two neighbours (lbm_peer_attach) or any number of slabs (lbm_peer_attach_all).  lbm_ibm_exchange_floats is 0 when nothing has to be exchanged. */
int lbm_ibm_exchange_floats(lbm_handle* h, int64_t* out);
int lbm_ibm_pack(lbm_handle* h, float* d_buf);
int lbm_ibm_unpack(lbm_handle* h, const float* d_buf);

/* ---- peer-mapped neighbours: halo rows without copies (no reference counterpart; SURVEY.md §5, §8e option 1) --------
 * Each slab exports a descriptor of its population buffer (a CUDA IPC handle when the neighbour lives in another
 * process, the plain device pointer when it lives in this one); the caller moves the LBM_PEER_DESC_BYTES to the two
 * neighbours (torch.distributed all_gather in cuda_lbm_b200/slab.py) and attaches them: side 0 = lower-y neighbour,
 * 1 = upper-y neighbour.  From then on the odd (neighbour) steps load and store the three populations that cross a slab
 * face directly in the neighbour's edge row over NVLink, inside the fused kernel, and every step is bracketed by a
 * device-side handshake (a step counter each slab writes into its neighbours' memory), so lbm_step(h, n) runs n steps
 * without host involvement and lbm_halo_* are not needed.  All slabs must have finished lbm_init_* (lbm_sync + a host
 * barrier) before any of them steps.  A neighbour that never arrives makes lbm_sync fail after 10 s instead of hanging. */
#define LBM_PEER_DESC_BYTES 128
int lbm_peer_export(lbm_handle* h, void* desc);
int lbm_peer_attach(lbm_handle* h, int side, const void* desc);
/* Same for the whole decomposition in one call: descs = the descriptors of ALL slabs, indexed by rank (world x LBM_PEER_DESC_BYTES).
 * Maps every other slab and attaches the two y-neighbours.  With all slabs mapped, (1) CM<2,OptimalAdapter> needs no host
 * all-reduce any more: each slab's sums of rho, rho|u|, |Pi| are stored into every slab's mailbox over NVLink by the reduction
 * kernel and added there in rank order (identical bits on every slab), so lbm_step(h, n) runs n steps on its own; (2) a body may
 * span any number of slabs. */
int lbm_peer_attach_all(lbm_handle* h, const void* descs, int32_t count);
int lbm_peer_detach(lbm_handle* h);
/* bounded != 0: lbm_step(h, n) on a peer-mapped slab keeps at most 12 steps enqueued ahead of the device (the host waits inside the
 * call).  For callers that drive every slab from its own host thread and may run more slabs than GPUs (the header shim with
 * LBM_B200_GPUS): slabs that share a device would otherwise fill the context's launch queue with launches that wait for a slab which
 * can then no longer enqueue.  Leave it off (default) when one thread enqueues the slabs one after the other. */
int lbm_set_lookahead(lbm_handle* h, int32_t bounded);

/* Pinned host memory helpers for callers that want full-speed host<->device copies.  lbm_host_alloc places the pages on the NUMA node
 * of the CURRENT device (cudaSetDevice first) where the box reports one (sysfs numa_node), so that a slab's copies do not cross the
 * socket interconnect; LBM_B200_HOST_NUMA=0 leaves the placement to the caller's own CPU affinity. */
int lbm_host_alloc(void** out, int64_t bytes);
int lbm_host_free(void* p);

const char* lbm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
