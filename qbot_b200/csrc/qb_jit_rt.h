// qbot_b200 -- run-time interface of the sweep specialiser (qb_jit.cu).
#pragma once
#include "qb_common.cuh"
#include "qb_jit.h"

struct QbJitKernel {
    void* fn = nullptr;          // CUfunction
    int threads = 0;
    int smem_bytes = 0;
    int npool = 0;
    bool pool_global = false;    // coefficients read from a device array instead of the parameter struct
    int ctas_per_sm = 2;
    int M = 12;
};

struct QbJitStats {
    uint64_t kernels_compiled = 0;
    uint64_t cache_hits = 0;
    double compile_ms = 0;
};

bool qb_jit_available(std::string* why);
std::string qb_jit_full_source(const uint8_t* program, QjSourceInfo* info, bool* pool_global, bool virtual_basis = false);
std::vector<char> qb_jit_compile(const std::string& src, std::string* log_out);
void qb_jit_precompile(const std::vector<const uint8_t*>& programs);
void qb_jit_compile_cached(const uint8_t* program, bool virtual_basis = false);
QbJitKernel qb_jit_get(const uint8_t* program, int device, bool virtual_basis = false);
void qb_jit_launch(const QbJitKernel& k, cudaStream_t stream, int sms, cplx* psi, uint64_t tile_begin, uint64_t tile_end, int prefetch,
                   const double* pool_host, const double* pool_dev, uint64_t virtual_index = 0);
QbJitStats qb_jit_stats();
// record one sighting of a specialised source (by hash): returns the number of sightings so far,
// or -1 when a kernel for it is already compiled
int qb_jit_note(uint64_t key);
