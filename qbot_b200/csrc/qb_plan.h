// qbot_b200 -- fused-sweep plan: data layout shared by the host planner, the CUDA tile kernel
// and the CPU plan emulator used by the tests.
//
// A SWEEP is one read + one write of the whole state.  The state is cut into TILES of
// 2^QT_M amplitudes: the QT_L lowest index bits (one contiguous 512-byte run) plus QT_H
// arbitrary higher "tile bits" chosen per sweep by the planner.  A CTA stages a tile in
// shared memory and runs the sweep's PROGRAM on it: a list of STAGES; in each stage every
// thread holds 2^R amplitudes (the R "register bits" of the stage) in registers and applies
// the stage's OPS to them; between stages the tile goes back through shared memory so that
// other bits can become register bits.  Controls and diagonal gates may sit on any bit
// (register / thread-local / outside the tile) -- they are predicates and scalars; only the
// non-diagonal targets of a gate have to be register bits.
#pragma once
#include <stdint.h>
#include <vector>

#define QT_M 12                 // tile bits
#define QT_L 5                  // contiguous low bits of a tile (2^5 * 16 B = 512 B runs)
#define QT_H (QT_M - QT_L)      // free tile bits
#define QT_RUNS (1 << QT_H)     // runs per tile
#define QT_MAXR 4               // register bits per stage (compile-time variants 3 and 4)
#define QT_MAX_PROGRAM_BYTES 12288

enum QtOpType : uint8_t {
    QT_OP_H = 1,        // s*[[1,1],[1,-1]] on register bit t0            pool: s
    QT_OP_X = 2,        // exchange the pair on register bit t0
    QT_OP_U2 = 3,       // general 2x2 on register bit t0                  pool: 4 complex
    QT_OP_U4 = 4,       // general 4x4 on register bits (t0 = msb, t1)     pool: 16 complex
    QT_OP_CDIAG = 5,    // diag(d0,d1) under a predicate; target by loc    pool: 2 complex
    QT_OP_PHASE = 6     // product of uncontrolled 1-qubit diagonals        pool: nent entries
};

enum QtLoc : uint8_t { QT_LOC_REG = 0, QT_LOC_LOCAL = 1, QT_LOC_GLOBAL = 2 };

struct QtOp {
    uint8_t type;
    uint8_t t0, t1;       // register-bit indices of the targets (CDIAG: t0 = position, t1 = QtLoc)
    uint8_t nent;         // PHASE: entries
    uint16_t regsel;      // register indices (bit i <-> a[i]) that satisfy the register-bit part of the predicate
    uint16_t lmask, lval; // predicate on the thread's tile-local index bits
    uint32_t pool;        // offset of the payload in the program's pool, in doubles
    uint32_t pad_;
    uint64_t gmask, gval; // predicate on index bits outside the tile (uniform per tile)
};

struct QtPhaseEntry {     // 5 doubles in the pool: {loc | pos<<8 as a double-encoded int, d0.re, d0.im, d1.re, d1.im}
    double code, d0re, d0im, d1re, d1im;
};

struct QtStage {
    uint8_t rb[QT_MAXR];       // tile-local positions of the register bits (a[i]: bit q of i <-> rb[q])
    uint8_t tpos[QT_M];        // tile-local position carried by thread-index bit q (QT_M - R entries used)
    uint16_t first_op, nops;
};

struct QtHeader {
    uint32_t total_bytes;
    uint16_t nstages, nops;
    uint16_t R;                // register bits per stage (3 or 4)
    uint16_t ngates;           // gates of the circuit executed by this sweep
    uint32_t stages_off, ops_off, pool_off;    // byte offsets from the start of the program
    uint8_t hb[QT_H];          // index-bit positions of the free tile bits, ascending
    uint8_t pad_[1];
};

// padded placement of a tile-local index in the shared-memory tile (in 16-byte units): every
// run of 32 amplitudes stays contiguous (bulk-copy friendly) while run k is shifted by
// k + k/8 + k/64 units so that tile bits 5..11 also select the bank
#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#else
#define QT_HD inline
#endif
QT_HD uint32_t qt_slot(uint32_t j) {
    uint32_t k = j >> QT_L;
    return j + k + (k >> 3) + (k >> 6);
}
#define QT_TILE_UNITS ((1 << QT_M) + 144)      // slots per buffer (16-byte units), >= qt_slot(4095)+1

// index of the first amplitude of tile t: t's bits deposited into the non-tile positions
QT_HD uint64_t qt_tile_base(uint64_t t, const uint8_t* hb) {
    uint64_t b = t << QT_L;
    for (int i = 0; i < QT_H; i++) {
        const int p = hb[i];
        b = ((b >> p) << (p + 1)) | (b & ((1ull << p) - 1ull));
    }
    return b;
}
// offset (in amplitudes) of run k inside a tile: k's bits deposited into the tile-bit positions
QT_HD uint64_t qt_run_offset(uint32_t k, const uint8_t* hb) {
    uint64_t o = 0;
    for (int i = 0; i < QT_H; i++) o |= (uint64_t)((k >> i) & 1u) << hb[i];
    return o;
}

// ---- host side --------------------------------------------------------------------------------
struct QGate;   // qb_common.cuh

struct QtPlanStep {
    bool fused;                         // true: run `program` with the tile kernel; false: gate_index unfused
    int gate_index;                     // for unfused steps
    std::vector<uint8_t> program;       // QtHeader + stages + ops + pool
    int ngates;
};

struct QtPlanOptions {
    int R = 4;
    bool merge_phases = true;
};

// Plan the execution of `gates` (in order) on a state with `nbits` index bits per branch.
// Every gate appears in exactly one step; the order of non-commuting gates is preserved.
std::vector<QtPlanStep> qt_plan(const std::vector<QGate>& gates, int nbits, const QtPlanOptions& opt);
