// qbot_b200 -- ket-level gate record as queued by the C ABI (no CUDA dependency so that the
// planner can also be compiled into the CPU plan emulator used by the tests).
#pragma once
#include <stdint.h>
#include <vector>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
typedef double2 cplx;   // x = re, y = im : one amplitude == one 128-bit word
#else
struct cplx { double x, y; };
#endif

#define QB_BIG_MAXK   14      // largest dense gate (out-of-place fallback above QB_REG_MAXK)

enum QbGateType { QB_G_DENSE = 0, QB_G_DIAG = 1, QB_G_MONO = 2 };

struct QGate {
    int type;
    int k;
    int tb[QB_BIG_MAXK];          // target bits, tb[0] = most significant matrix index bit
    uint64_t cmask;
    std::vector<cplx> m;          // DENSE: 4^k row-major | DIAG: 2^k | MONO: 2^k coefficients
    std::vector<int> src;         // MONO: y[i] = m[i] * x[src[i]]
    uint64_t tmask() const { uint64_t t = 0; for (int i = 0; i < k; i++) t |= 1ull << tb[i]; return t; }
};

// Recognise structure (diagonal / monomial / dense) of a 2^k x 2^k row-major matrix.
QGate qb_classify(const cplx* m, int k, const int* tb, uint64_t cmask);
bool qb_is_identity(const QGate& g);
std::vector<cplx> qb_dense_of(const QGate& g);
