/* wl_c_host.c -- the C-ABI of include/wheeledlab_b200.h driven from plain C (no Python, no torch).
 *
 *   wl_c_host <config.bin> [steps]
 *
 * <config.bin> is a wl_config blob written by the Python task layer (python -m wheeledlab_b200.dump_config drift 4096
 * cfg.bin): task definitions live in one place, any host language loads the POD.  The program owns every device buffer
 * (cudaMalloc), creates the handle, runs `steps` env.steps with the library's synthetic actions and prints order-independent
 * checksums of everything it got back; tests/test_gpu_parity.py::test_c_host_example_matches_python_path compares them
 * with the same run through the Python host.  Build: see examples/c_host/Makefile (gcc + libcudart + libwheeledlab_b200). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "wheeledlab_b200.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        int rc_ = (call);                                                                          \
        if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, wl_last_error()); return 1; } \
    } while (0)
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } \
    } while (0)

static uint64_t bits_sum(const void* p, size_t n_words) {           /* sum of the 32-bit patterns: order-independent, exact */
    const uint32_t* w = (const uint32_t*)p;
    uint64_t s = 0;
    for (size_t i = 0; i < n_words; ++i) s += w[i];
    return s;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <config.bin> [steps]\n", argv[0]); return 2; }
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    wl_config cfg;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(&cfg, 1, sizeof cfg, f) != sizeof cfg || wl_config_sizeof() != sizeof cfg) {
        fprintf(stderr, "cannot read a %zu-byte wl_config from %s (library says %zu)\n", sizeof cfg, argv[1], wl_config_sizeof());
        return 2;
    }
    fclose(f);
    if (cfg.task != WL_TASK_DRIFT) { fprintf(stderr, "this example drives the Drift task (no auxiliary terrain data)\n"); return 2; }
    const int n = cfg.num_envs, od = WL_OBS_DIM_BLIND;
    const size_t sb = wl_state_bytes(n);
    void* d_state; float *d_act, *d_obs, *d_rew, *d_log; uint8_t *d_term, *d_trunc;
    CU(cudaMalloc(&d_state, sb)); CU(cudaMemset(d_state, 0, sb));
    CU(cudaMalloc((void**)&d_act, sizeof(float) * 2 * n)); CU(cudaMalloc((void**)&d_obs, sizeof(float) * od * n));
    CU(cudaMalloc((void**)&d_rew, sizeof(float) * n)); CU(cudaMalloc((void**)&d_log, sizeof(float) * WL_LOG_FLOATS));
    CU(cudaMalloc((void**)&d_term, n)); CU(cudaMalloc((void**)&d_trunc, n));
    float* h_obs = (float*)malloc(sizeof(float) * od * n); float* h_rew = (float*)malloc(sizeof(float) * n);
    uint8_t* h_term = (uint8_t*)malloc(n); uint8_t* h_trunc = (uint8_t*)malloc(n);
    cudaStream_t s; CU(cudaStreamCreate(&s));
    wl_sim* sim;
    CK(wl_create(&cfg, d_state, sb, NULL, &sim));
    CK(wl_startup(sim, s));
    CK(wl_reset(sim, NULL, 0, 0, s));
    uint64_t obs_sum = 0, rew_sum = 0; long n_term = 0, n_trunc = 0;
    for (int t = 0; t < steps; ++t) {
        CK(wl_synth_actions(sim, d_act, t, 0, s));
        CK(wl_step(sim, d_act, d_obs, d_rew, d_term, d_trunc, d_log, t, s));
        CU(cudaMemcpyAsync(h_obs, d_obs, sizeof(float) * od * n, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(h_rew, d_rew, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(h_term, d_term, n, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(h_trunc, d_trunc, n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        obs_sum += bits_sum(h_obs, (size_t)od * n); rew_sum += bits_sum(h_rew, n);
        for (int i = 0; i < n; ++i) { n_term += h_term[i]; n_trunc += h_trunc[i]; }
    }
    printf("{\"build\": \"%s\", \"envs\": %d, \"steps\": %d, \"obs_bits_sum\": %llu, \"rew_bits_sum\": %llu, \"terminated\": %ld, "
           "\"truncated\": %ld, \"launches\": %lld}\n", wl_build_info(), n, steps, (unsigned long long)obs_sum,
           (unsigned long long)rew_sum, n_term, n_trunc, (long long)wl_launch_count(sim));
    CK(wl_destroy(sim));
    cudaFree(d_state); cudaFree(d_act); cudaFree(d_obs); cudaFree(d_rew); cudaFree(d_log); cudaFree(d_term); cudaFree(d_trunc);
    free(h_obs); free(h_rew); free(h_term); free(h_trunc);
    return 0;
}
