// qbot_b200 -- fused-sweep plan: data layout shared by the host planner, the CUDA tile kernel
// and the CPU plan emulator used by the tests.
//
// A SWEEP is one read + one write of the whole state.  The state is cut into TILES of 2^M
// amplitudes (M = 11 or 12, chosen per plan): the QT_L lowest index bits (one contiguous
// 512-byte run) plus H = M - QT_L arbitrary higher "tile bits" chosen per sweep by the
// planner.  A CTA of 2^(M-4) threads owns one tile at a time and runs the sweep's PROGRAM on
// it: a list of STAGES; in each stage every thread holds 16 amplitudes (the 4 "register bits"
// of the stage) in registers and applies the stage's OPS to them.  The first stage loads its
// amplitudes straight from HBM and the last one stores them straight back (their lanes are the
// low QT_L bits, so every warp access is one 512-byte run); between two stages the tile goes
// through shared memory once so that other bits can become register bits.  Controls and
// diagonal gates may sit on any bit (register / thread-local / outside the tile) -- they are
// predicates and scalars; only the non-diagonal targets of a gate have to be register bits.
#pragma once
#include <stdint.h>
#include <vector>

#define QT_L 5                  // contiguous low bits of a tile (2^5 * 16 B = 512 B runs)
#define QT_R 4                  // register bits per stage of programs the GENERIC kernel runs (16 amplitudes per thread)
#define QT_MAXR 5               // specialised kernels may hold 32 amplitudes per thread (QtHeader.R says which)
#define QT_MAXM 12              // largest tile (bits)
#define QT_MINM 11
#define QT_MAXH (QT_MAXM - QT_L)
#define QT_MAX_OPS 64           // ops per sweep (one bit each in the per-thread predicate mask)
#define QT_MAX_STAGES 8
#define QT_MAX_PROGRAM_BYTES 8192
// fixed layout of a program (lets the kernel address stages / ops / pool with constant offsets)
#define QT_STAGES_OFF 64
#define QT_OPS_OFF 256
#define QT_POOL_OFF (QT_OPS_OFF + QT_MAX_OPS * 40)

enum QtOpType : uint8_t {
    QT_OP_H = 1,        // UNSCALED butterfly [[1,1],[1,-1]] on register bit t0 (the 2^-1/2 factors
                        // are collected into a PHASE constant or the header scale)
    QT_OP_X = 2,        // exchange the pair on register bit t0
    QT_OP_U2 = 3,       // general 2x2 on register bit t0                  pool: 4 complex
    QT_OP_U4 = 4,       // general 4x4 on register bits (t0 = msb, t1)     pool: 16 complex
    QT_OP_CDIAG = 5,    // diag(d0,d1) under a predicate; target by loc    pool: 2 complex
    QT_OP_PHASE = 6     // product of uncontrolled 1-qubit diagonals       pool: nent entries
};

enum QtLoc : uint8_t { QT_LOC_REG = 0, QT_LOC_LOCAL = 1, QT_LOC_GLOBAL = 2, QT_LOC_CONST = 3 };

#define QT_FLAG_GLOBAL 1u       // the op has a predicate on bits outside the tile
#define QT_FLAG_ALLREG 2u       // regsel covers all 2^R registers

struct QtOp {             // 40 bytes
    uint8_t type;
    uint8_t t0, t1;       // register-bit indices of the targets (CDIAG: t0 = position, t1 = QtLoc)
    uint8_t nent;         // PHASE: entries
    uint32_t regsel;      // register indices (bit i <-> a[i]) that satisfy the register-bit part of the predicate
    uint16_t lmask, lval; // predicate on the thread's tile-local index bits
    uint16_t flags;
    uint16_t pad_;
    uint32_t pool;        // offset of the payload in the program's pool, in doubles
    uint32_t pad2_;
    uint64_t gmask, gval; // predicate on index bits outside the tile (uniform per tile)
};

struct QtPhaseEntry {     // 5 doubles in the pool; `code` holds loc | pos << 8 as an integer bit pattern
    int64_t code;
    double d0re, d0im, d1re, d1im;
};

struct QtStage {
    uint8_t rb[QT_MAXR];       // tile-local positions of the register bits (a[i]: bit q of i <-> rb[q]; R used)
    uint8_t tpos[QT_MAXM];     // tile-local position carried by thread-index bit q (M - R entries used)
    uint16_t first_op, nops;
};

struct QtHeader {
    uint32_t total_bytes;
    uint16_t nstages, nops;
    uint16_t M;                // tile bits of this program
    uint16_t ngates;           // gates of the circuit executed by this sweep
    uint32_t stages_off, ops_off, pool_off;    // byte offsets from the start of the program
    uint8_t hb[QT_MAXH];       // index-bit positions of the free tile bits (M - QT_L used): hb[i] sits at tile-local
                               // position QT_L + i.  Any order: the planner permutes it so that no stage's register bits
                               // exhaust a shared-memory bank class (see qt_slot)
    uint8_t R;                 // register bits per stage: 4 (16 amplitudes per thread) or 5 (32; specialised kernels only)
    double scale;              // every amplitude is multiplied by this before the store (1.0: skipped)
    uint8_t hbs[QT_MAXH];      // the same positions in ascending order (what qt_tile_base needs)
};

// padded placement of a tile-local index in the shared-memory tile (in 16-byte units): run k
// (32 amplitudes) is shifted by k + k/8 + k/64 units so that tile bits 5..11 also select the
// bank.  The map is additive over disjoint bit sets: qt_slot(a | b) = qt_slot(a) + qt_slot(b).
#if defined(__CUDACC__)
#define QT_HD __host__ __device__ __forceinline__
#else
#define QT_HD inline
#endif
QT_HD uint32_t qt_slot(uint32_t j) {
    uint32_t k = j >> QT_L;
    return j + k + (k >> 3) + (k >> 6);
}
#define QT_TILE_UNITS(M) ((1 << (M)) + (1 << ((M) - QT_L)) + (1 << ((M) - QT_L - 3)) + 8)

// index of the first amplitude of tile t: t's bits deposited into the non-tile positions
// (hb in ASCENDING order: QtHeader.hbs)
QT_HD uint64_t qt_tile_base(uint64_t t, const uint8_t* hb, int nh) {
    uint64_t b = t << QT_L;
    for (int i = 0; i < nh; i++) {
        const int p = hb[i];
        b = ((b >> p) << (p + 1)) | (b & ((1ull << p) - 1ull));
    }
    return b;
}
// offset (in amplitudes) of run k inside a tile: k's bits deposited into the tile-bit positions
QT_HD uint64_t qt_run_offset(uint32_t k, const uint8_t* hb, int nh) {
    uint64_t o = 0;
    for (int i = 0; i < nh; i++) o |= (uint64_t)((k >> i) & 1u) << hb[i];
    return o;
}

// ---- host side --------------------------------------------------------------------------------
struct QGate;   // qb_gate.h

struct QtPlanStep {
    bool fused;                         // true: run `program` with the tile kernel; false: gate_index unfused
    int gate_index;                     // for unfused steps
    std::vector<uint8_t> program;       // QtHeader + stages + ops + pool
    int ngates;
};

struct QtPlanOptions {
    int M = 12;
    int R = QT_R;               // register bits per stage (5: for the sweep specialiser only)
    bool merge_phases = true;
    int search_trials = 1;      // > 1: randomised search over the tile-bit choices, fewest steps wins
};

// Plan the execution of `gates` (in order) on a state with `nbits` index bits per branch.
// Every gate appears in exactly one step; the order of non-commuting gates is preserved.
std::vector<QtPlanStep> qt_plan(const std::vector<QGate>& gates, int nbits, const QtPlanOptions& opt);

// Unitary-preserving rewrite of a gate list into a cheaper one (see qb_plan.cpp); plan the result.
// level 0: none, 1: X..H -> H Z, 2: also H..X -> Z H
std::vector<QGate> qt_peephole(const std::vector<QGate>& gates, int level, int* rewritten);
// the cheapest plan over the rewrite levels; *planned = the gate list the returned steps refer to
std::vector<QtPlanStep> qt_plan_best(const std::vector<QGate>& gates, int nbits, const QtPlanOptions& opt, std::vector<QGate>* planned);
