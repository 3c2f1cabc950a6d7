// qbot_b200 -- sweep specialiser: turns one fused-sweep program (qb_plan.h) into the source of
// a tile kernel in which everything structural is a compile-time constant -- tile bits, stage
// register / thread maps, op order, targets, predicates -- and only the gate coefficients (the
// program's pool) stay run-time kernel parameters.  A circuit with the same structure but other
// angles reuses the compiled kernel.
//
// The generated text is plain C++ over a handful of macros (QJ_DEV, QJ_C, QJ_LD, QJ_ST, QJ_P,
// QJ_POOL_PARAM, QJ_RESTRICT, QJ_WAR_SYNC, QJ_SYNC, QJ_PREFETCH) so that the very same text is
// compiled by NVRTC into the sm_100a kernel (qb_jit.cu supplies the CUDA prelude + kernel) and
// by g++ into the CPU emulation the tests run against the oracle (tests/jit_emu.py).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

struct QjSourceInfo {
    int M = 0;            // tile bits
    int R = 4;            // register bits per stage (2^R amplitudes per thread)
    int threads = 0;      // 2^(M-4)
    int npool = 0;        // doubles of run-time coefficients
    int nstages = 0;
    int tile_units = 0;   // 16-byte units of shared memory for the transposition buffer
};

// Source of the stage functions of `program` (QtHeader + stages + ops + pool).
std::string qj_generate(const uint8_t* program, QjSourceInfo* info);

// the run-time coefficients the generated code reads through QJ_P(i): the program's pool followed
// by the header scale (info->npool doubles)
std::vector<double> qj_pool(const uint8_t* program);

int qj_default_ctas(int M, int R);

// 64-bit FNV-1a of a source text (the key of the compiled-kernel cache)
uint64_t qj_hash(const std::string& src);
