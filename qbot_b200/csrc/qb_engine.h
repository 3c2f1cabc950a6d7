// qbot_b200 -- state handle layout and the fused-engine entry points.
#pragma once
#include "../../include/qbot_b200.h"
#include "qb_common.cuh"

struct qb_state {
    int kind = 0;
    int nq = 0;
    int nbits = 0;                 // index bits per branch
    int64_t nbranch = 1;
    int device = 0;
    int sms = 148;
    cplx* d = nullptr;
    cplx* scratch = nullptr;
    bool owns = true;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    bool fusion = true;
    int jit_mode = -1;             // -1: QBOT_B200_JIT (default 1 = specialise repeated plans), 0 never, 1 on repeat, 2 always
    std::vector<QGate> queue;
    qb_stats stats = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // staging ring for small per-gate tables
    static constexpr size_t STAGE_BYTES = 4u << 20;
    void* stage = nullptr;
    size_t stage_off = 0;
    // fused engine state (plan cache, device copies of the sweep programs)
    void* engine = nullptr;
    // The register IS the basis state |virt_index> and nothing has been written yet (qb_init_basis / a product of basis
    // kets on a large single-branch ket): the first fused sweep starts from the known state instead of loading it
    // (qj_kernel's virtual-basis variant); anything else that touches the amplitudes writes them first (materialize).
    bool virt = false;
    uint64_t virt_index = 0;
    void materialize();
    // a planned gate list whose steps the caller runs range by range (qb_plan_queue / qb_run_steps / qb_finish_queue)
    void* pending_plan = nullptr;
    int pending_jit = 0;
    std::vector<uint32_t> pending_done;       // per step: parts that have run
    std::vector<int> pending_parts_of;        // per step: number of parts it was run in

    ~qb_state();
    uint64_t per_branch() const { return 1ull << nbits; }
    uint64_t total() const { return (uint64_t)nbranch << nbits; }
    size_t bytes() const { return (size_t)total() * sizeof(cplx); }
    LaunchCtx ctx();
    cplx* get_scratch();
    cplx* upload_small(const void* host, size_t bytes);
    void enqueue(QGate&& g);
    void flush();
    void run_gate_unfused(const QGate& g);
};

// fused tile engine (qb_tile.cu)
bool qb_engine_available();
void qb_engine_run(qb_state* s, const std::vector<QGate>& gates);
void qb_engine_free(qb_state* s);
int qb_engine_plan_pending(qb_state* s, const std::vector<QGate>& gates, int park_bits, int* head, int* tail);
void qb_engine_run_pending(qb_state* s, int from, int to, int part, int nparts, int sms);
void qb_engine_finish_pending(qb_state* s);
