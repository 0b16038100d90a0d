// irec_host.h -- host-side plumbing shared by the .cu translation units of libirec.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#include "../../include/irec.h"

#define IREC_RATIO_LEN 65536

struct IrecDevice {
    bool ready;
    int device;
    int sm_count;
    int max_smem_optin;
    float* d_T;        // [10008] quantile table, entry 0 unused
    float* d_ratio;    // [IREC_RATIO_LEN] power-law auxiliary ratios
    int ratio_len;
};

const IrecDevice& irec_device();                   // tables of the CURRENT device (after irec_init)
int irec_fail(int code, const char* msg);          // records the message, returns code
int irec_check_launch(const char* what);           // cudaGetLastError -> IREC_E_CUDA
void irec_count_launch();
bool irec_force_general();                          // env IREC_FORCE_GENERAL=1 (tests)

#define IREC_ENSURE_INIT()                     \
    do {                                       \
        const int rc_init_ = irec_init();      \
        if (rc_init_ != IREC_OK) return rc_init_; \
    } while (0)
