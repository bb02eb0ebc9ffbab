// kbench — times the fused step kernel through the C ABI, even and odd AA phases separately (CUDA events on the
// handle's stream).  Development tool; bench.py is the contract benchmark.
//   usage: kbench nx ny collision(0..3) steps [general(0/1/2)] [adapter_mode]
//   general = 2: BASELINE config 3, the lid-driven cavity (regularized walls / lid / corners, nu = 0.1 ny / 1000).  CM<OptimalAdapter> keeps
//   it physical for ~40 steps only, so the run is cut into cycles of init -> 16 warm-up steps -> 16 timed steps (one CUDA-graph replay)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../include/lbm_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed: %s\n", #x, lbm_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    int nx = argc > 1 ? atoi(argv[1]) : 8192, ny = argc > 2 ? atoi(argv[2]) : 8192;
    int coll = argc > 3 ? atoi(argv[3]) : 0, steps = argc > 4 ? atoi(argv[4]) : 20;
    int general = argc > 5 ? atoi(argv[5]) : 0, amode = argc > 6 ? atoi(argv[6]) : 1;
    lbm_config c; lbm_default_config(&c);
    c.nx = nx; c.ny = ny; c.collision = coll; c.periodic_x = 1; c.periodic_y = 1; c.u_max = 0.04f; c.adapter_mode = amode;
    if (coll >= 2) { float om = 1.0f; float S[9] = {0, 0, 0, 1, om, om, 1, 1, 1}; for (int i = 0; i < 9; i++) c.S[i] = S[i]; }
    if (general == 1) { c.periodic_y = 0; c.force_x = 1e-6f; }
    if (general == 2) {
        c.periodic_x = c.periodic_y = 0; c.u_max = 0.1f; c.viscosity = 0.1f * ny / 1000.0f;
        const float om = 1.0f / (3 * c.viscosity + 0.5f);
        float S[9] = {0, 0, 0, 1, om, om, 1, 1, 1};
        if (coll < 2) { float T[9] = {0, om, om, 0, om, 0, om, om, om}; for (int i = 0; i < 9; i++) S[i] = T[i]; }
        for (int i = 0; i < 9; i++) c.S[i] = S[i];
    }
    lbm_handle* h;
    CK(lbm_create(&c, &h));
    cudaStream_t s; cudaStreamCreate(&s);
    CK(lbm_set_stream(h, s));
    if (general == 1) {
        std::vector<int32_t> fl((size_t)nx * ny, 0);
        for (int x = 0; x < nx; x++) { fl[x] = LBM_BOUNCE_BACK; fl[(size_t)(ny - 1) * nx + x] = LBM_BOUNCE_BACK; }
        CK(lbm_set_flags(h, fl.data()));
    }
    if (general == 2) {
        std::vector<int32_t> fl((size_t)nx * ny, 0);
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                const bool ex = x == 0 || x == nx - 1, ey = y == 0 || y == ny - 1;
                fl[(size_t)y * nx + x] = (ex && ey) ? LBM_REGULARIZED_BOUNCE_BACK_CORNER : (y == ny - 1 ? LBM_REGULARIZED_INLET_TOP : ((ex || y == 0) ? LBM_REGULARIZED_BOUNCE_BACK : LBM_FLUID));
            }
        CK(lbm_set_flags(h, fl.data()));
        std::vector<float> rho((size_t)nx * ny, 1.0f), u((size_t)nx * ny * 2, 0.0f);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        double total = 0.0; int cycles = steps > 0 ? steps : 8;
        double mass = 0.0;
        for (int cy = 0; cy < cycles; cy++) {
            CK(lbm_init_fields(h, rho.data(), u.data()));
            CK(lbm_step(h, 16)); CK(lbm_sync(h));
            cudaEventRecord(e0, s);
            CK(lbm_step(h, 16));
            cudaEventRecord(e1, s); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (cy > 0) total += ms;                 // the first cycle captures the graph
            CK(lbm_total_mass(h, &mass));
        }
        const double msps = total / (16.0 * (cycles - 1)), n = (double)nx * ny;
        lbm_info_t inf; lbm_info(h, &inf);
        printf("KBENCH cavity nx %d ny %d coll %d adapter %d: %.4f ms/step  %.0f MLUPS  %.0f GB/s (72 B/cell)  mass/N %.6f  (%d cycles of 16 + 16 steps)\n", nx, ny, coll, amode, msps,
               n / msps / 1e3, 72.0 * n / msps / 1e6, mass / n, cycles);
        lbm_destroy(h);
        return 0;
    }
    CK(lbm_init_taylor_green(h, 1.0f / 6.0f, 0.04f / (nx / 128.0f)));
    CK(lbm_step(h, 4)); CK(lbm_sync(h));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> t[2];
    for (int i = 0; i < steps; i++) {
        int odd = lbm_next_step_needs_halo(h);     // 0 for world==1; use the info timestep instead
        lbm_info_t inf; lbm_info(h, &inf); odd = (inf.timestep + 1) & 1;
        cudaEventRecord(e0, s);
        CK(lbm_step(h, 1));
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        t[odd].push_back(ms);
    }
    cudaEventRecord(e0, s);
    CK(lbm_step(h, steps));
    cudaEventRecord(e1, s); cudaEventSynchronize(e1);
    float tot; cudaEventElapsedTime(&tot, e0, e1);
    double n = (double)nx * ny;
    for (int p = 0; p < 2; p++) {
        std::sort(t[p].begin(), t[p].end());
        float med = t[p][t[p].size() / 2];
        printf("%s step: median %.4f ms  %.0f MLUPS  %.0f GB/s (72 B/cell)\n", p ? "odd " : "even", med, n / med / 1e3, 72.0 * n / med / 1e6);
    }
    double ms = tot / steps;
    double mass; CK(lbm_total_mass(h, &mass));
    printf("KBENCH nx %d ny %d coll %d general %d: %.4f ms/step  %.0f MLUPS  %.0f GB/s  mass/N %.6f\n", nx, ny, coll, general, ms, n / ms / 1e3, 72.0 * n / ms / 1e6, mass / n);
    lbm_destroy(h);
    return 0;
}
