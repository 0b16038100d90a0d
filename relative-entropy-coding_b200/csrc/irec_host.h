// irec_host.h -- host-side plumbing shared by the .cu translation units of libirec.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "../../include/irec.h"

#define IREC_RATIO_LEN 65536
#define IREC_T2_LEN 30020      // 3 * 10006 entries (index a' + c <= 30016, a' in {a, a + 10006}), padded to a multiple of 4

struct IrecDevice {
    bool ready;
    int device;
    int sm_count;
    int max_smem_optin;
    float* d_T;        // [10008] quantile table, entry 0 unused
    float* d_ratio;    // [IREC_RATIO_LEN] power-law auxiliary ratios
    int ratio_len;
    float* d_T2;       // [IREC_T2_LEN] exponent-indexed quantile table, stored three times: T2[e] = T[g^(e mod 10006) mod 10007]
    uint16_t* d_dl4;   // [10006] dl4[m] = 4 * dlog_g(m + 1): byte offset into T2 of r = m + 1
};

const IrecDevice& irec_device();                   // tables of the CURRENT device (after irec_init)
const float* irec_ratio_tab();                     // auxiliary variance ratios in force for the calling thread (device pointer)
int irec_ratio_len();
int irec_reserved_sms();                           // SMs the persistent batch kernels leave free (irec_set_thread_reserved_sms / IREC_RESERVE_SMS)
int irec_fail(int code, const char* msg);          // records the message, returns code
int irec_check_launch(const char* what);           // cudaGetLastError -> IREC_E_CUDA
void irec_count_launch();
bool irec_force_general();                          // env IREC_FORCE_GENERAL=1 (tests)

// irec_cluster.cu: one thread-block cluster per coder-block (few blocks per launch)
int irec_cluster_choice(int nb, int max_D, int S, int B);       // cluster size to use, 0 = use the persistent kernels
size_t irec_cluster_hist_bytes(int nb, int max_aux);
int irec_launch_cluster(int G, const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                        const int64_t* gidx, const int64_t* offs, int nb, int max_D, float omega, int S, int B, int64_t seed,
                        int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                        int2* hist, const int32_t* order, const void* plan, const void* tab, int tab_aux, const void* tab_priv,
                        cudaStream_t s);

// irec_tmem.cu: persistent kernel with the beams in tensor memory, two coder-blocks in flight per SM
struct TmemPlan { int bmax; int DPmax; int NC; int grid; size_t smem; };
bool irec_tmem_plan(int nb, int max_D, int S, int B, TmemPlan* out);      // false: sizes not covered (use resident2)
int irec_launch_tmem(const TmemPlan& p, const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gidx, const int64_t* offs, int nb, float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                     int2* hist, float* sched, int* work_counter, const int32_t* order, const void* plan, const void* tab, int tab_aux,
                     const void* tab_priv, cudaStream_t s);

// irec_wide.cu: 32 < n_beams <= IREC_WIDE_MAX_BEAMS (beams, scores and history in global memory, radix-select top-B)
#define IREC_WIDE_MAX_BEAMS 1024
bool irec_wide_supported(int nb, int max_D, int S, int B);
size_t irec_wide_workspace_bytes(int nb, int max_D, int S, int B, int max_aux);
int irec_launch_wide(const float* t_loc, const float* t_scale, const float* p_loc, const float* p_scale,
                     const int64_t* gidx, const int64_t* offs, int nb, int max_D, float omega, int S, int B, int64_t seed,
                     int32_t* out_indices, int max_aux, int32_t* out_n_aux, int32_t* out_status, float* out_sample,
                     void* workspace, cudaStream_t s);

#define IREC_ENSURE_INIT()                     \
    do {                                       \
        const int rc_init_ = irec_init();      \
        if (rc_init_ != IREC_OK) return rc_init_; \
    } while (0)
