// core/lbm.cuh — LBM<2>, the solver object of the reference (src/core/lbm.cuh:33-382), on top of the B200 engine's C ABI.
//
// Same public surface and driver protocol as the reference (src/main.cu:72-153):
//
//     LBM<dimensions> lbm;  lbm.allocate<Scenario>();  lbm.init<Scenario>();
//     per step:  increase_ts<S>(); stream(); swap_buffers(); apply_boundaries<S>(); uncorrected_macroscopics();
//                reset_forces<S>(); ibm_step(); correct_macroscopics(); compute_equilibrium(); collide<S::CollisionOp>();
//     now and then:  save_vtk(t) / save_macroscopics(t) / update_macroscopics() / compute_error<S>()
//
// What differs underneath: the nine per-step methods do not launch nine kernels over three population buffers.  They
// describe ONE time step, which the engine executes as one fused, register-resident kernel over a single in-place SoA
// buffer (cuda_lbm_b200/csrc/kernels.cuh).  collide<Op>() closes the description; the step is enqueued when the next one
// begins (increase_ts) or when its macroscopic fields are asked for — in which case that launch also stores rho and u
// exactly as the reference's d_rho / d_u hold them after correct_macroscopics().  Work runs on the legacy default stream,
// so cudaEventRecord(…, 0) pairs around a loop iteration time it as they do for the reference.
//
// The scenario's functors run in this translation unit: Init::operator() on the device over whole-grid AoS arrays
// (rho[node], u[2*node+c], force[2*node+c]) and Boundary::operator() once per node; the results go through
// lbm_init_fields_device / lbm_set_flags / lbm_set_body_force.  Errors are fatal, as in the reference
// (src/util/utility.cu:4-12): message to stderr, exit(99).
#ifndef LBM_H
#define LBM_H

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "defines.hpp"
#include "util/utility.cuh"
#include "core/lbm_constants.cuh"
#include "core/streaming/streaming.cuh"
#include "core/collision/collision.cuh"
#include "functors/includes.cuh"
#include "assert.cuh"
#include "IBM/IBMBody.cuh"
#include "../../lbm_b200.h"

namespace fs = std::filesystem;

// grid means used by CM<2,OptimalAdapter> in the last step (reference src/core/lbm.cuh:25-29)
struct MomentInfo {
    float rho_avg_norm;
    float j_avg_norm;
    float pi_avg_norm;
};

namespace lbm_b200_shim {

inline void fatal(const char* what) {
    std::fprintf(stderr, "[LBM] %s failed: %s\n", what, lbm_last_error());
    std::exit(99);
}
#define LBM_B200_CALL(expr) do { if ((expr) != LBM_OK) ::lbm_b200_shim::fatal(#expr); } while (0)

// periodic axes: a scenario's own `periodic_x` / `periodic_y` members win over the PERIODIC_X / PERIODIC_Y macros
template <typename S, typename = void> struct periodic_x_of { static constexpr bool value = lbm_b200_periodic_x_default; };
template <typename S> struct periodic_x_of<S, std::void_t<decltype(S::periodic_x)>> { static constexpr bool value = S::periodic_x; };
template <typename S, typename = void> struct periodic_y_of { static constexpr bool value = lbm_b200_periodic_y_default; };
template <typename S> struct periodic_y_of<S, std::void_t<decltype(S::periodic_y)>> { static constexpr bool value = S::periodic_y; };

// Init functor over a slab's rows, one node per thread: init_kernel of src/core/init/init.cuh:29-43 without the population part.
// The functor indexes whole-grid arrays with the global node; the base pointers are shifted so that node `first` is element 0.
template <typename Init>
__global__ void slab_init_functor_kernel(Init init, float* rho, float* u, float* force, long long first, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) init(rho, u, force, (int)(first + i));
}
// reset_forces_kernel of src/core/macroscopics/macroscopics.cuh:13-27
template <typename Init>
__global__ void slab_force_functor_kernel(Init init, float* rho, float* u, float* force, long long first, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) init.apply_forces(rho, u, force, (int)(first + i));
}
// Boundary functor over the whole grid (setup_boundary_flags, src/core/boundaries/boundaries.cuh:170-197, runs it on the host)
template <typename Boundary>
__global__ void boundary_functor_kernel(Boundary b, int* flags, int nx, int ny, int* any_non_fluid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny) return;
    const int v = b((int)(i % nx), (int)(i / nx));
    flags[i] = v;
    if (v != 0) *any_non_fluid = 1;
}
// is force[node] the same vector on every node?
static __global__ void force_uniform_kernel(const float2* force, int n, int* differs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 a = force[0], b = force[i];
    if (a.x != b.x || a.y != b.y) *differs = 1;
}

// Validation functor over a slab's rows: (x, y) -> analytic velocity (taylorGreenFunctors.cuh:66-81, poiseuilleFunctors.cuh:72-75)
template <typename Validation>
__global__ void validation_functor_kernel(Validation v, float2* u_ref, int nx, int y0, int nyl) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nx * nyl) return;
    float ux = 0.0f, uy = 0.0f;
    v((int)(i % nx), y0 + (int)(i / nx), ux, uy);
    u_ref[i] = make_float2(ux, uy);
}

// does the scenario's Validation functor map (x, y) to a velocity?  (void for scenarios without one; Ghia tables for the cavity)
template <typename V, typename = void> struct is_field_validation : std::false_type {};
template <typename V>
struct is_field_validation<V, std::void_t<decltype(std::declval<const V&>()(0, 0, std::declval<float&>(), std::declval<float&>()))>> : std::true_type {};

// does it carry the centre-line tables of the reference's LidDrivenValidation (lidDrivenCavityFunctors.cuh:58-221)?
template <typename V, typename = void> struct is_centerline_validation : std::false_type {};
template <typename V>
struct is_centerline_validation<V, std::void_t<decltype(V::ghia_u_count), decltype(std::declval<const V&>().get_closest_ref_data(0, true))>> : std::true_type {};

inline int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : fallback;
}

}  // namespace lbm_b200_shim

namespace lbm_b200_shim {

// One host thread per slab when the domain is split over several GPUs: every slab's calls are enqueued from its own thread,
// so no slab's launch queue can fill up (and block the host) while another slab — whose step counter it waits for on the
// device — has not been enqueued yet.  A single slab runs inline on the caller's thread.
class SlabThreads {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable go, done;
    std::function<void(int)> job;
    unsigned long generation = 0;
    int pending = 0;
    bool quit = false;

    void loop(int g, int device) {
        cudaSetDevice(device);
        unsigned long seen = 0;
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> lk(m);
                go.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
                f = job;
            }
            f(g);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) done.notify_all();
            }
        }
    }

public:
    void start(const std::vector<int>& devices) {
        stop();
        quit = false;
        if (devices.size() > 1)
            for (int g = 0; g < (int)devices.size(); g++) th.emplace_back([this, g, d = devices[g]] { loop(g, d); });
    }
    template <typename F>
    void run(int n, F&& f) {
        if (th.empty()) { for (int g = 0; g < n; g++) f(g); return; }
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::function<void(int)>(f);
            pending = (int)th.size();
            generation++;
        }
        go.notify_all();
        std::unique_lock<std::mutex> lk(m);
        done.wait(lk, [&] { return pending == 0; });
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        go.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
    ~SlabThreads() { stop(); }
};

// scenarios whose body force changes in time declare `static constexpr bool time_dependent_forces = true;` (detected, never required):
// reset_forces<Scenario>() then re-evaluates Init::apply_forces every step, as the reference does for every scenario
// (src/core/macroscopics/macroscopics.cuh:13-48); without the member the force is evaluated once, at init()
template <typename S, typename = void> struct has_time_dependent_forces : std::false_type {};
template <typename S> struct has_time_dependent_forces<S, std::void_t<decltype(S::time_dependent_forces)>> : std::bool_constant<S::time_dependent_forces> {};

}  // namespace lbm_b200_shim

template <int dim>
class LBM {
    static_assert(dim == 2, "the B200 engine covers the reference's D2Q9 path (LBM<2>)");

private:
    // The domain is one slab on the current device (the reference's configuration) or, with LBM_B200_GPUS=N in the environment
    // ("all" = every visible device), N y-slabs on N GPUs of the box: one engine handle, one host thread and one stream per slab,
    // all slabs peer-mapped (halo rows, IBM node states and the adapter sums travel over NVLink inside the kernels).
    struct Slab { lbm_handle* h = nullptr; int device = 0, y0 = 0, nyl = 0; float* d_force = nullptr; };
    std::vector<Slab> slabs;
    lbm_b200_shim::SlabThreads workers;
    lbm_handle* h = nullptr;             // slabs[0].h
    int home_device = 0;
    bool step_pending = false;           // a step has been described (collide) but not enqueued yet
    std::vector<IBMBody> bodies;         // owned from allocate() on, released in free()
    void (LBM::*force_refresh)() = nullptr;      // set by reset_forces<S>() of a scenario with time-dependent forces

    template <typename F>
    void each_slab(F&& f) {
        workers.run((int)slabs.size(), [&](int g) {
            if (slabs.size() > 1) cudaSetDevice(slabs[g].device);
            f(slabs[g], g);
        });
    }

    void flush(bool want_macroscopics) {
        if (!step_pending) return;
        if (force_refresh) { (this->*force_refresh)(); want_macroscopics = true; }      // the next evaluation reads this step's rho / u
        each_slab([&](Slab& s, int) {
            if (want_macroscopics) LBM_B200_CALL(lbm_step_with_macroscopics(s.h, 1));
            else LBM_B200_CALL(lbm_step(s.h, 1));
        });
        step_pending = false;
    }

    template <typename Scenario>
    void send_consts() {
        const float nu = Scenario::viscosity;
        for (const Slab& s : slabs) {
            checkCudaErrors(cudaSetDevice(s.device));
            checkCudaErrors(cudaMemcpyToSymbol(vis, &nu, sizeof(float)));
        }
        checkCudaErrors(cudaSetDevice(home_device));
    }

    template <typename BoundaryFunctor>
    void setup_boundary_flags(BoundaryFunctor boundary_func) {
        const long long n = (long long)NX * NY;
        int *d_flags = nullptr, *d_any = nullptr, any = 0;
        checkCudaErrors(cudaMalloc(&d_flags, n * sizeof(int)));
        checkCudaErrors(cudaMalloc(&d_any, sizeof(int)));
        checkCudaErrors(cudaMemset(d_any, 0, sizeof(int)));
        lbm_b200_shim::boundary_functor_kernel<<<(unsigned)((n + 255) / 256), 256>>>(boundary_func, d_flags, NX, NY, d_any);
        checkCudaErrors(cudaGetLastError());
        checkCudaErrors(cudaMemcpy(&any, d_any, sizeof(int), cudaMemcpyDeviceToHost));
        if (any) {      // an all-FLUID scenario (Taylor-Green) needs no flag plane at all
            std::vector<int32_t> flags((size_t)n);
            checkCudaErrors(cudaMemcpy(flags.data(), d_flags, n * sizeof(int), cudaMemcpyDeviceToHost));
            for (const Slab& s : slabs) LBM_B200_CALL(lbm_set_flags(s.h, flags.data()));       // every slab picks its rows
            checkCudaErrors(cudaSetDevice(home_device));
        }
        checkCudaErrors(cudaFree(d_flags));
        checkCudaErrors(cudaFree(d_any));
    }

    // The functors index whole-grid arrays with the GLOBAL node (rho[node], u[2*node+c] / u[get_vec_index(node, c)]): a
    // slab hands them base pointers shifted by its first node, so that the writes land in its own rows' storage.
    template <typename Init>
    void run_force_functor(Init init, const Slab& s, const float* d_rho_local, const float* d_u_local) {
        const long long n = (long long)s.nyl * NX, first = (long long)s.y0 * NX;
        lbm_b200_shim::slab_force_functor_kernel<<<(unsigned)((n + 255) / 256), 256>>>(init, const_cast<float*>(d_rho_local) - first,
                                                                                    const_cast<float*>(d_u_local) - 2 * first, s.d_force - 2 * first, first, n);
        checkCudaErrors(cudaGetLastError());
    }

    // Init::apply_forces for every node -> one uniform body force (two kernel constants) or a per-node force plane
    template <typename Init>
    void upload_forces(Init init, const std::vector<float*>& d_rho, const std::vector<float*>& d_u) {
        std::vector<int> differs(slabs.size(), 0);
        std::vector<float> f0(2 * slabs.size(), 0.0f);
        for (size_t g = 0; g < slabs.size(); g++) {
            const Slab& s = slabs[g];
            checkCudaErrors(cudaSetDevice(s.device));
            const int n = s.nyl * NX;
            run_force_functor(init, s, d_rho[g], d_u[g]);
            int* d_differs = nullptr;
            checkCudaErrors(cudaMalloc(&d_differs, sizeof(int)));
            checkCudaErrors(cudaMemset(d_differs, 0, sizeof(int)));
            lbm_b200_shim::force_uniform_kernel<<<(n + 255) / 256, 256>>>(reinterpret_cast<const float2*>(s.d_force), n, d_differs);
            checkCudaErrors(cudaMemcpy(&differs[g], d_differs, sizeof(int), cudaMemcpyDeviceToHost));
            checkCudaErrors(cudaMemcpy(&f0[2 * g], s.d_force, 2 * sizeof(float), cudaMemcpyDeviceToHost));
            checkCudaErrors(cudaFree(d_differs));
        }
        bool uniform = true;
        for (size_t g = 0; g < slabs.size(); g++) uniform = uniform && !differs[g] && f0[2 * g] == f0[0] && f0[2 * g + 1] == f0[1];
        for (const Slab& s : slabs) {
            checkCudaErrors(cudaSetDevice(s.device));
            if (uniform) {
                LBM_B200_CALL(lbm_set_force_field(s.h, nullptr));
                LBM_B200_CALL(lbm_set_body_force(s.h, f0[0], f0[1]));
            } else {
                LBM_B200_CALL(lbm_set_force_field_device(s.h, s.d_force));
            }
        }
        checkCudaErrors(cudaSetDevice(home_device));
    }

    // reset_forces<Scenario>() of a scenario with time-dependent forces: Init::apply_forces with Scenario::t of the step being
    // described and the macroscopic fields of the last completed step (the reference hands it the uncorrected moments of the
    // step in flight, which a fused step does not materialise; identical for forces that depend on position and time only)
    template <typename Scenario>
    void eval_forces_now() {
        auto init = Scenario::init();
        each_slab([&](Slab& s, int) {
            LBM_B200_CALL(lbm_recover_macroscopics(s.h));
            const float *r = nullptr, *u = nullptr;
            LBM_B200_CALL(lbm_get_macroscopics_device(s.h, &r, &u));
            run_force_functor(init, s, r, u);
            checkCudaErrors(cudaDeviceSynchronize());
            LBM_B200_CALL(lbm_set_force_field_device(s.h, s.d_force));
        });
    }

public:
    std::vector<float> h_rho;            // [node]
    std::vector<float> h_u;              // [node*2 + c]
    int timestep = 0, update_ts = 0;

    LBM() = default;
    LBM(const LBM&) = delete;
    LBM& operator=(const LBM&) = delete;

    int num_slabs() const { return (int)slabs.size(); }

    // LBM::allocate<Scenario>() — src/core/lbm.cuh:92-125
    template <typename Scenario>
    void allocate() {
        std::cout << "[LBM]: allocating\n";
        h_rho.resize((size_t)NX * NY);
        h_u.resize((size_t)NX * NY * dimensions);

        checkCudaErrors(cudaGetDevice(&home_device));
        int ndev = 1;
        checkCudaErrors(cudaGetDeviceCount(&ndev));
        int G = 1;
        if (const char* v = std::getenv("LBM_B200_GPUS")) G = (v[0] == 'a' || v[0] == 'A') ? ndev : std::max(1, std::atoi(v));
        if (G > NY / 2) G = std::max(1, NY / 2);
        slabs.assign((size_t)G, Slab());
        std::vector<int> devices;
        for (int g = 0; g < G; g++) {
            lbm_config cfg;
            LBM_B200_CALL(lbm_default_config(&cfg));
            cfg.nx = NX;
            cfg.ny = NY;
            cfg.periodic_x = lbm_b200_shim::periodic_x_of<Scenario>::value;
            cfg.periodic_y = lbm_b200_shim::periodic_y_of<Scenario>::value;
            cfg.collision = Scenario::CollisionOp::lbm_b200_op;
            cfg.viscosity = Scenario::viscosity;
            for (int i = 0; i < quadratures; i++) cfg.S[i] = Scenario::S[i];
            cfg.u_max = Scenario::u_max;
            // LBM_QK_REFERENCE reproduces the reference's arithmetic including its defects; 0 repairs them
            cfg.quirks = lbm_b200_shim::env_int("LBM_B200_QUIRKS", LBM_QK_REFERENCE);
            cfg.adapter_mode = lbm_b200_shim::env_int("LBM_B200_ADAPTER", LBM_ADAPTER_EXACT);
            cfg.device = G == 1 ? home_device : g % ndev;       // more slabs than devices: they share (diagnosis / tests on a one-GPU box)
            cfg.rank = g;
            cfg.world = G;
            Slab& s = slabs[(size_t)g];
            LBM_B200_CALL(lbm_create(&cfg, &s.h));
            lbm_info_t inf;
            LBM_B200_CALL(lbm_info(s.h, &inf));
            s.device = cfg.device; s.y0 = inf.y0; s.nyl = inf.ny_local;
            devices.push_back(s.device);
        }
        h = slabs[0].h;
        if (G == 1) {
            LBM_B200_CALL(lbm_set_stream(h, (void*)cudaStreamLegacy));      // cudaEventRecord(…, 0) pairs in a reference-style driver time the work
        } else {
            std::vector<unsigned char> descs((size_t)G * LBM_PEER_DESC_BYTES);
            for (int g = 0; g < G; g++) LBM_B200_CALL(lbm_peer_export(slabs[(size_t)g].h, descs.data() + (size_t)g * LBM_PEER_DESC_BYTES));
            for (int g = 0; g < G; g++) {
                LBM_B200_CALL(lbm_peer_attach_all(slabs[(size_t)g].h, descs.data(), G));
                LBM_B200_CALL(lbm_set_lookahead(slabs[(size_t)g].h, 1));       // every slab has its own host thread here, and slabs may share a device
            }
            std::cout << "[LBM]: " << G << " y-slabs on " << std::min(G, ndev) << " GPU(s), peer-mapped\n";
        }
        checkCudaErrors(cudaSetDevice(home_device));
        workers.start(devices);

        Scenario::add_bodies();
        for (const IBMBody& b : Scenario::IBM_bodies) {
            for (const Slab& s : slabs) {       // every slab is given every body; it works on those that reach into its rows
                LBM_B200_CALL(lbm_add_body(s.h, b.points, b.num_points));
                // IBMBody::velocities: dead data in the reference (IBM_impl.cuh:15, A-D9) and here under LBM_QK_D9_IBM_ZERO_TARGET
                // (part of LBM_QK_REFERENCE); with that bit cleared (LBM_B200_QUIRKS) the markers force the fluid towards them
                if (b.velocities && b.num_points > 0) LBM_B200_CALL(lbm_set_body_velocities(s.h, (int)bodies.size(), b.velocities));
            }
            bodies.push_back(b);
        }
        checkCudaErrors(cudaSetDevice(home_device));
    }

    void free() {
        if (slabs.empty()) return;
        std::cout << "[LBM]: Freeing\n";
        step_pending = false;
        workers.stop();
        for (Slab& s : slabs) {
            if (s.d_force) { cudaSetDevice(s.device); cudaFree(s.d_force); }
            LBM_B200_CALL(lbm_destroy(s.h));
        }
        cudaSetDevice(home_device);
        slabs.clear();
        h = nullptr;
        for (IBMBody& b : bodies) h_ibm_free(b);
        bodies.clear();
    }

    // LBM::init<Scenario>() — src/core/init/init.cuh:45-86
    template <typename Scenario>
    void init() {
        LBM_DEVICE_ASSERT(Scenario::viscosity > 0.0f, "Negative Viscosity");
        LBM_DEVICE_ASSERT(Scenario::tau > 0.5f, "Instability warning: tau < 0.5");
        LBM_DEVICE_ASSERT(Scenario::u_max < 0.5f, "Instability warning: u_max > 0.5");
        auto init = Scenario::init();
        auto boundary_func = Scenario::boundary();
        send_consts<Scenario>();

        // the Init functor on every slab's device over that slab's rows
        std::vector<float*> d_rho(slabs.size(), nullptr), d_u(slabs.size(), nullptr);
        for (size_t g = 0; g < slabs.size(); g++) {
            Slab& s = slabs[g];
            checkCudaErrors(cudaSetDevice(s.device));
            const long long n = (long long)s.nyl * NX, first = (long long)s.y0 * NX;
            checkCudaErrors(cudaMalloc(&d_rho[g], (size_t)n * sizeof(float)));
            checkCudaErrors(cudaMalloc(&d_u[g], (size_t)n * 2 * sizeof(float)));
            if (!s.d_force) checkCudaErrors(cudaMalloc(&s.d_force, (size_t)n * 2 * sizeof(float)));
            // entries a functor leaves unwritten are zero here (the reference reads uninitialised memory, Appendix A-D13)
            checkCudaErrors(cudaMemset(d_rho[g], 0, (size_t)n * sizeof(float)));
            checkCudaErrors(cudaMemset(d_u[g], 0, (size_t)n * 2 * sizeof(float)));
            checkCudaErrors(cudaMemset(s.d_force, 0, (size_t)n * 2 * sizeof(float)));
            lbm_b200_shim::slab_init_functor_kernel<<<(unsigned)((n + 255) / 256), 256>>>(init, d_rho[g] - first, d_u[g] - 2 * first, s.d_force - 2 * first, first, n);
            checkCudaErrors(cudaGetLastError());
            checkCudaErrors(cudaDeviceSynchronize());
        }
        checkCudaErrors(cudaSetDevice(home_device));

        setup_boundary_flags(boundary_func);
        upload_forces(init, d_rho, d_u);            // what reset_forces<Scenario>() computes every step in the reference
        for (size_t g = 0; g < slabs.size(); g++) {
            const Slab& s = slabs[g];
            checkCudaErrors(cudaSetDevice(s.device));
            LBM_B200_CALL(lbm_init_fields_device(s.h, d_rho[g], d_u[g]));
            LBM_B200_CALL(lbm_sync(s.h));
            checkCudaErrors(cudaFree(d_rho[g]));
            checkCudaErrors(cudaFree(d_u[g]));
            if (!lbm_b200_shim::has_time_dependent_forces<Scenario>::value) { checkCudaErrors(cudaFree(slabs[g].d_force)); slabs[g].d_force = nullptr; }
        }
        checkCudaErrors(cudaSetDevice(home_device));
        timestep = 0;
        update_ts = 0;
        step_pending = false;
        force_refresh = nullptr;
        printf("[init_kernel]: Threads executed: %d\n", NX * NY);
    }

    template <typename Scenario>
    void increase_ts() {
        flush(false);
        timestep++;
        Scenario::update_ts(timestep);
    }

    // ---- the reference's per-kernel host methods (src/core/lbm.cuh:345-377): one fused launch stands for all of them
    void stream() {}
    void swap_buffers() {}
    template <typename Scenario> void apply_boundaries() {}
    void uncorrected_macroscopics() {}
    // Init::apply_forces was evaluated in init().  A scenario that declares `static constexpr bool time_dependent_forces = true;`
    // gets the reference's behaviour — re-evaluated for every step (src/core/macroscopics/macroscopics.cuh:13-48) — at the price
    // of a force plane and the general (scalar) kernel path.
    template <typename Scenario>
    void reset_forces() {
        if constexpr (lbm_b200_shim::has_time_dependent_forces<Scenario>::value) force_refresh = &LBM::template eval_forces_now<Scenario>;
    }
    void ibm_step() {}
    void correct_macroscopics() {}
    void compute_equilibrium() {}
    void compute_forces() {}
    template <typename CollisionOp>
    void collide() {
        if (h == nullptr) { std::fprintf(stderr, "[LBM] collide() before allocate()\n"); std::exit(99); }
        flush(false);                    // a driver that never calls increase_ts still advances one step per collide()
        step_pending = true;
    }

    // extension: n whole time steps without per-step host calls (timestep and Scenario::t advance by n)
    template <typename Scenario>
    void run(int n) {
        if (n <= 0) return;
        flush(false);
        if constexpr (lbm_b200_shim::has_time_dependent_forces<Scenario>::value) {
            for (int i = 0; i < n; i++) {           // the force changes between steps: one at a time
                increase_ts<Scenario>();
                reset_forces<Scenario>();
                collide<typename Scenario::CollisionOp>();
            }
            return;
        }
        each_slab([&](Slab& s, int) { LBM_B200_CALL(lbm_step(s.h, n - 1)); });
        timestep += n;
        Scenario::update_ts(timestep);
        step_pending = true;
    }

    // extension: enqueue the step described so far now (instead of at the next increase_ts), keeping its rho / u for a
    // following update_macroscopics() — lets a driver bracket exactly the stepping work with CUDA events (one slab) or with
    // synchronize() and a wall clock (several slabs)
    void finish_step(bool keep_macroscopics = true) { flush(keep_macroscopics); }

    // extension: re-evaluate Init::apply_forces now (with the macroscopic fields of the last completed step)
    template <typename Scenario>
    void refresh_forces() {
        flush(true);
        for (Slab& s : slabs)
            if (!s.d_force) { checkCudaErrors(cudaSetDevice(s.device)); checkCudaErrors(cudaMalloc(&s.d_force, (size_t)s.nyl * NX * 2 * sizeof(float))); }
        checkCudaErrors(cudaSetDevice(home_device));
        std::vector<float*> d_rho, d_u;
        for (Slab& s : slabs) {
            LBM_B200_CALL(lbm_recover_macroscopics(s.h));
            const float *r = nullptr, *u = nullptr;
            LBM_B200_CALL(lbm_get_macroscopics_device(s.h, &r, &u));
            LBM_B200_CALL(lbm_sync(s.h));
            d_rho.push_back(const_cast<float*>(r)); d_u.push_back(const_cast<float*>(u));
        }
        upload_forces(Scenario::init(), d_rho, d_u);
    }

    // LBM::update_macroscopics() — src/core/lbm.cuh:148-154
    void update_macroscopics() {
        flush(true);        // no-op after finish_step(): the macroscopics of the current step are already on the device
        update_ts = timestep;
        each_slab([&](Slab& s, int) {
            // the step was closed without them (increase_ts() of the next step came first): rebuilt from its populations, so that —
            // as in the reference, whose d_rho / d_u are always current — this call is legal at any point of the driver loop
            LBM_B200_CALL(lbm_recover_macroscopics(s.h));
            LBM_B200_CALL(lbm_get_macroscopics(s.h, h_rho.data() + (size_t)s.y0 * NX, h_u.data() + (size_t)s.y0 * NX * 2));
        });
    }

    // device views of rho[node] and u[node*2+c], valid until the next step (d_rho / d_u of the reference); one slab only
    float* get_rho() { return const_cast<float*>(device_macroscopics(0)); }
    float* get_u() { return const_cast<float*>(device_macroscopics(1)); }

    // LBM::compute_error<Scenario>() — src/core/lbm.cuh:163-171
    template <typename Scenario>
    float compute_error() {
        if constexpr (Scenario::has_analytical_solution) {
            return Scenario::compute_error(*this);
        } else {
            printf("Scenario does not provide verification/validation.\n");
            return 0.0f;
        }
    }

    // extension: the relative L2 velocity error in percent against Scenario::validation() (the metric of
    // taylorGreenScenario.cuh:59-88) with the functor evaluated and both sums taken on the device (fp64, fixed order):
    // nothing but two doubles per slab crosses PCIe, where compute_error<S>() moves 12 B/node to the host first
    template <typename Scenario>
    float l2_error_device() {
        flush(true);
        std::vector<double> sums(2 * slabs.size(), 0.0);
        auto validation = Scenario::validation();
        each_slab([&](Slab& s, int g) {
            LBM_B200_CALL(lbm_recover_macroscopics(s.h));
            const long long n = (long long)s.nyl * NX;
            float2* d_ref = nullptr;
            checkCudaErrors(cudaMalloc(&d_ref, (size_t)n * sizeof(float2)));
            lbm_b200_shim::validation_functor_kernel<<<(unsigned)((n + 255) / 256), 256>>>(validation, d_ref, NX, s.y0, s.nyl);
            checkCudaErrors(cudaGetLastError());
            checkCudaErrors(cudaDeviceSynchronize());
            LBM_B200_CALL(lbm_velocity_error_sums(s.h, reinterpret_cast<const float*>(d_ref), &sums[2 * (size_t)g]));
            checkCudaErrors(cudaFree(d_ref));
        });
        double e = 0.0, r = 0.0;
        for (size_t g = 0; g < slabs.size(); g++) { e += sums[2 * g]; r += sums[2 * g + 1]; }
        return (float)(std::sqrt(e / r) * 100.0);
    }

    // extension: u[2*i + c] at the listed global nodes, gathered on the device (2 floats per node cross PCIe instead of the field)
    std::vector<float> sample_velocity(const std::vector<long long>& nodes) {
        flush(true);
        std::vector<float> out(2 * nodes.size(), 0.0f);
        static_assert(sizeof(long long) == sizeof(int64_t), "node ids are 64-bit");
        for (Slab& s : slabs) {
            if (slabs.size() > 1) checkCudaErrors(cudaSetDevice(s.device));
            LBM_B200_CALL(lbm_recover_macroscopics(s.h));
            LBM_B200_CALL(lbm_sample_velocity(s.h, reinterpret_cast<const int64_t*>(nodes.data()), (int32_t)nodes.size(), out.data()));
        }
        if (slabs.size() > 1) checkCudaErrors(cudaSetDevice(home_device));
        return out;
    }

    // extension: the centre-line metric of LidDrivenScenario::compute_error (src/scenarios/lidDrivenCavity/lidDrivenCavityScenario.cuh:88-157:
    // NRMSE of u_x on the vertical and u_y on the horizontal centre line against the tables of Ghia, Ghia & Shin 1982, mean of the two,
    // in percent) with the 2 x 17 samples gathered on the device.  Works with any Validation functor that has the reference's table
    // interface (ghia_u_count, ghia_x, ghia_y, get_closest_ref_data).
    template <typename Scenario>
    float centerline_error_device() {
        const auto validator = Scenario::validation();
        const int re = (int)compute_reynolds(Scenario::u_max, NY, Scenario::viscosity);
        const float* ux_ref = validator.get_closest_ref_data(re, true);
        const float* uy_ref = validator.get_closest_ref_data(re, false);
        const int n = validator.ghia_u_count;
        std::vector<long long> nodes((size_t)2 * n);
        for (int i = 0; i < n; i++) {
            int y = (int)std::round(validator.ghia_y[i] * (NY - 1)), x = (int)std::round(validator.ghia_x[i] * (NX - 1));
            y = std::max(0, std::min(y, NY - 1)); x = std::max(0, std::min(x, NX - 1));
            nodes[(size_t)i] = (long long)y * NX + NX / 2;
            nodes[(size_t)n + i] = (long long)(NY / 2) * NX + x;
        }
        const std::vector<float> u = sample_velocity(nodes);
        float ex = 0.0f, rx = 0.0f, ey = 0.0f, ry = 0.0f;
        for (int i = 0; i < n; i++) {
            const float dx = u[2 * (size_t)i] * (1.0f / Scenario::u_max) - ux_ref[i], dy = u[2 * ((size_t)n + i) + 1] * (1.0f / Scenario::u_max) - uy_ref[i];
            ex += dx * dx; rx += ux_ref[i] * ux_ref[i];
            ey += dy * dy; ry += uy_ref[i] * uy_ref[i];
        }
        const float rmse_x = rx > 0.0f ? std::sqrt(ex / rx) : 0.0f, rmse_y = ry > 0.0f ? std::sqrt(ey / ry) : 0.0f;
        return 100.0f * (rmse_x + rmse_y) / 2.0f;
    }

    // extension: checkpoint / restart of the population state (lbm_checkpoint_write / lbm_checkpoint_read).  load_checkpoint
    // is called after allocate<S>() and init<S>() (which set flags, forces and bodies) and continues bit-identically.
    // Several slabs: one file per slab, `path`.slab<g>.
    void save_checkpoint(const std::string& path) {
        flush(true);      // a pending step is closed WITH its rho / u: a following save_vtk() / update_macroscopics() finds them
        each_slab([&](Slab& s, int g) { LBM_B200_CALL(lbm_checkpoint_write(s.h, slab_path(path, g).c_str())); });
    }
    template <typename Scenario>
    void load_checkpoint(const std::string& path) {
        step_pending = false;
        each_slab([&](Slab& s, int g) { LBM_B200_CALL(lbm_checkpoint_read(s.h, slab_path(path, g).c_str())); });
        lbm_info_t inf;
        LBM_B200_CALL(lbm_info(h, &inf));
        timestep = inf.timestep;
        update_ts = -1;
        Scenario::update_ts(timestep);
    }

    // grid means of rho, rho|u|, |Pi| that the last CM<2,OptimalAdapter> step used (the reference's d_moment_avg)
    MomentInfo moment_avg() {
        flush(true);
        float a[3];
        LBM_B200_CALL(lbm_moment_avg(h, a));
        return MomentInfo{a[0], a[1], a[2]};
    }
    double total_mass() {
        flush(true);
        std::vector<double> m(slabs.size(), 0.0);
        each_slab([&](Slab& s, int g) { LBM_B200_CALL(lbm_total_mass(s.h, &m[(size_t)g])); });
        double t = 0.0;
        for (double v : m) t += v;
        return t;
    }
    void synchronize() {
        flush(true);
        each_slab([&](Slab& s, int) { LBM_B200_CALL(lbm_sync(s.h)); });
    }
    lbm_handle* handle(int slab = 0) { return slabs.at((size_t)slab).h; }

private:
    std::string slab_path(const std::string& path, int g) const { return slabs.size() == 1 ? path : path + ".slab" + std::to_string(g); }
    const float* device_macroscopics(int which) {
        if (slabs.size() != 1) { std::fprintf(stderr, "[LBM] get_rho() / get_u() are device views of ONE slab; with LBM_B200_GPUS > 1 use update_macroscopics() and h_rho / h_u\n"); std::exit(99); }
        flush(true);
        LBM_B200_CALL(lbm_recover_macroscopics(h));
        const float *r = nullptr, *u = nullptr;
        LBM_B200_CALL(lbm_get_macroscopics_device(h, &r, &u));
        return which == 0 ? r : u;
    }

public:
    // raw dumps read by src/graphics/*.py: output/density/density_<t>.bin (NX*NY floats) and
    // output/velocity/velocity_<t>.bin (NX*NY*2 floats, AoS) — src/core/lbm.cuh:173-204
    void save_macroscopics(int ts) {
        update_macroscopics();
        const struct { const char* dir; const char* stem; const std::vector<float>* data; } outs[2] = {
            {"output/density", "density_", &h_rho}, {"output/velocity", "velocity_", &h_u}};
        for (const auto& o : outs) {
            std::error_code ec;
            fs::create_directories(o.dir, ec);
            const std::string path = std::string(o.dir) + "/" + o.stem + std::to_string(ts) + ".bin";
            std::ofstream f(path, std::ios::binary);
            if (!f) { printf("Error: Could not open file %s for writing.\n", path.c_str()); return; }
            f.write(reinterpret_cast<const char*>(o.data->data()), (std::streamsize)(o.data->size() * sizeof(float)));
        }
    }

    // VTK ImageData with raw appended data, same arrays and names as the reference writes (src/core/lbm.cuh:262-343):
    // Float32 "Density", then Float32 x3 "Velocity" with a zero z component; UInt64 block headers.
    void save_vtk(int ts) {
        update_macroscopics();
        const std::string dir = "output/vtk";
        std::error_code ec;
        fs::create_directories(dir, ec);
        if (ec) { std::cerr << "Filesystem error creating directory " << dir << ": " << ec.message() << std::endl; return; }
        std::ostringstream name;
        name << dir << "/sim_data_" << std::setw(6) << std::setfill('0') << ts << ".vti";
        std::ofstream out(name.str(), std::ios::binary);
        if (!out) { std::cerr << "Error: Could not open file " << name.str() << " for writing." << std::endl; return; }
        std::cout << "[VTK Export] Saving data for timestep " << ts << " to " << name.str() << std::endl;
        const uint64_t rho_bytes = h_rho.size() * sizeof(float);
        const std::string extent = "0 " + std::to_string(NX - 1) + " 0 " + std::to_string(NY - 1) + " 0 " + std::to_string(NZ - 1);
        out << "<?xml version=\"1.0\"?>\n"
            << "<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
            << "  <ImageData WholeExtent=\"" << extent << "\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n"
            << "    <Piece Extent=\"" << extent << "\">\n"
            << "      <PointData Scalars=\"Density\" Vectors=\"Velocity\">\n"
            << "        <DataArray type=\"Float32\" Name=\"Density\" format=\"appended\" offset=\"0\"/>\n"
            << "        <DataArray type=\"Float32\" Name=\"Velocity\" NumberOfComponents=\"3\" format=\"appended\" offset=\"" << rho_bytes << "\"/>\n"
            << "      </PointData>\n"
            << "      <CellData>\n"
            << "      </CellData>\n"
            << "    </Piece>\n"
            << "  </ImageData>\n"
            << "  <AppendedData encoding=\"raw\">\n"
            << "   _";
        std::vector<float> u3((size_t)NX * NY * 3);
        for (size_t node = 0; node < (size_t)NX * NY; node++) {
            u3[3 * node] = h_u[2 * node];
            u3[3 * node + 1] = h_u[2 * node + 1];
            u3[3 * node + 2] = 0.0f;
        }
        const uint64_t u_bytes = u3.size() * sizeof(float);
        out.write(reinterpret_cast<const char*>(&rho_bytes), sizeof(uint64_t));
        out.write(reinterpret_cast<const char*>(h_rho.data()), (std::streamsize)rho_bytes);
        out.write(reinterpret_cast<const char*>(&u_bytes), sizeof(uint64_t));
        out.write(reinterpret_cast<const char*>(u3.data()), (std::streamsize)u_bytes);
        out << "\n  </AppendedData>\n</VTKFile>\n";
    }

    ~LBM() { free(); }
};

#endif  // LBM_H
