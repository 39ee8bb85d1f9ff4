// TEST INFRASTRUCTURE ONLY — the step kernels' CUDA source (fastace_b200/csrc/match_kernel.cuh, update_kernel of
// match_update_kernels.cuh) compiled for the CPU under tests/emu/warp_emu.h and driven exactly as
// launch_step (fastace_capi.cu) drives them on the device.  Lets `pytest -m "not gpu"` check the kernel logic
// against the oracle in a container without a GPU.  Built by tests/emu/build_emu.py with -ffp-contract=off
// (the device build uses --fmad=false).
#include "warp_emu.h"

#include "../../fastace_b200/csrc/match_update_kernels.cuh"
#include "../../fastace_b200/csrc/shuffle_kernel.cuh"
#include "../../fastace_b200/csrc/segmented_sort.cuh"
#include "../../fastace_b200/csrc/packed_kernel.cuh"

#include <vector>

using namespace fastace;

template <int G>
static void run_step(const StepParams& sp, std::vector<uint8_t>& pnh, std::vector<uint8_t>& pnb, uint32_t* err,
                     std::vector<uint32_t>& queue, uint32_t& queue_launches) {
    const uint32_t flags = sp.flags;
    const bool ph_p = (flags & FASTACE_STEP_PERSONS) != 0, only_f = (flags & FASTACE_STEP_FIRMS) != 0;
    const bool ph_t = (flags & FASTACE_STEP_PERSONS_TRADE) != 0, ph_c = (flags & FASTACE_STEP_PERSONS_CONSUME) != 0;
    const bool only_p = ph_p || ph_t || ph_c;
    MatchParams mp;
    mp.sp = sp; mp.scr_pnh = pnh.data(); mp.scr_pnb = pnb.data(); mp.dev_err = err;
    mp.lay = make_match_layout(sp.P, sp.F, G, sp.S);
    // as launch_step does: a full step goes through the completion queue
    const bool use_queue = !only_p && !only_f && sp.E > 0;
    mp.done_list = use_queue ? queue.data() : nullptr;
    mp.done_count = use_queue ? queue.data() + sp.E : nullptr;
    mp.ticket_base = queue_launches * (uint32_t)sp.E;
    mp.done_tag = (queue_launches % 255u) + 1u;
    if (!ph_c) {
        // as launch_step does: the specialised kernel when the call is what it was compiled for
        const bool special = sp.compact && !only_p && !only_f && !sp.out.p_job_ok && !sp.out.p_good_ok && !sp.out.f_good_ok &&
                             !sp.out.old_j_left && !sp.out.old_j_taken && !sp.out.old_m_left && !sp.out.old_m_taken;
        if (!special) emu::launch(match_kernel<G, kModeGeneric>, (unsigned)sp.E, 32u, (size_t)mp.lay.total, mp);
        else {
            const bool small = sp.F * (G + 1) + 2 <= 32, mod = (sp.flags & FASTACE_IDX_MODULO) != 0;
            if (small && mod) emu::launch(match_kernel<G, kModeCompact | kModeModulo | kModeSmall>, (unsigned)sp.E, 32u, (size_t)mp.lay.total, mp);
            else if (small) emu::launch(match_kernel<G, kModeCompact | kModeSmall>, (unsigned)sp.E, 32u, (size_t)mp.lay.total, mp);
            else if (mod) emu::launch(match_kernel<G, kModeCompact | kModeModulo>, (unsigned)sp.E, 32u, (size_t)mp.lay.total, mp);
            else emu::launch(match_kernel<G, kModeCompact>, (unsigned)sp.E, 32u, (size_t)mp.lay.total, mp);
        }
    }
    const bool ces = sp.util_kind == FASTACE_FN_CES && sp.prod_kind == FASTACE_FN_CES;   // as launch_step does
    UpdateParams up;
    up.sp = sp; up.scr_pnh = pnh.data(); up.scr_pnb = pnb.data();
    up.done_list = mp.done_list; up.done_tag = mp.done_tag; up.dev_err = err;
    if (use_queue) {
        const int qgroups = (sp.E + kQueueGroup - 1) / kQueueGroup;
        up.group_person_blocks = (kQueueGroup * sp.P + kUpdateThreads - 1) / kUpdateThreads;
        const int qblocks = qgroups * (kQueueGroup / (kUpdateThreads / 32) + up.group_person_blocks);
        if (ces) emu::launch(update_kernel<G, true, true>, (unsigned)qblocks, (unsigned)kUpdateThreads, 0, up);
        else emu::launch(update_kernel<G, false, true>, (unsigned)qblocks, (unsigned)kUpdateThreads, 0, up);
        queue_launches += 1;
        return;
    }
    const size_t persons = (size_t)sp.E * sp.P;
    const int person_blocks = only_f ? 0 : (int)((persons + kUpdateThreads - 1) / kUpdateThreads);
    up.firm_blocks = only_p ? 0 : (sp.E + kUpdateThreads / 32 - 1) / (kUpdateThreads / 32);
    if (person_blocks + up.firm_blocks > 0)
        if (ces) emu::launch(update_kernel<G, true, false>, (unsigned)(person_blocks + up.firm_blocks), (unsigned)kUpdateThreads, 0, up);
        else emu::launch(update_kernel<G, false, false>, (unsigned)(person_blocks + up.firm_blocks), (unsigned)kUpdateThreads, 0, up);
}

struct EmuEnv {
    std::vector<uint8_t> pnh, pnb;
    uint32_t err[4] = {0, 0, 0, 0};
    std::vector<uint32_t> queue;
    uint32_t queue_launches = 0;
};

extern "C" {

void* fastace_emu_create(const fastace_dims_t* d) {
    EmuEnv* e = new EmuEnv();
    e->pnh.assign((size_t)d->num_econ * d->num_persons + 1, 0xEE);
    e->pnb.assign((size_t)d->num_econ * d->num_persons * d->num_goods + 1, 0xEE);
    e->queue.assign((size_t)d->num_econ + 1, 0u);
    return e;
}
void fastace_emu_destroy(void* h) { delete static_cast<EmuEnv*>(h); }

// One call of launch_step's default path on HOST arrays.  Exactly one of actions / compact is non-null.
// Returns the error words the kernels raised (0 = converged).
int fastace_emu_step(void* h, const fastace_dims_t* d, fastace_state_t* state, const fastace_actions_t* actions,
                     const fastace_actions_compact_t* compact, const fastace_step_out_t* out, uint32_t flags,
                     uint32_t time_before, int util_kind, int prod_kind) {
    EmuEnv* env = static_cast<EmuEnv*>(h);
    StepParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.E = d->num_econ; sp.P = d->num_persons; sp.F = d->num_firms; sp.S = d->stack_size;
    sp.flags = flags; sp.time_before = time_before;
    sp.util_kind = util_kind; sp.prod_kind = prod_kind;
    sp.st = *state;
    if (compact) {
        sp.compact = 1;
        sp.cz = *compact;
        sp.ac.p_consume = compact->p_consume; sp.ac.f_prod = compact->f_prod; sp.ac.f_offer_amt = compact->f_offer_amt;
        sp.ac.f_offer_price = compact->f_offer_price; sp.ac.f_job_labor = compact->f_job_labor; sp.ac.f_job_wage = compact->f_job_wage;
    } else {
        sp.ac = *actions;
    }
    sp.out = *out;
    switch (d->num_goods) {
        case 1: run_step<1>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        case 2: run_step<2>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        case 3: run_step<3>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        case 4: run_step<4>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        case 5: run_step<5>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        case 8: run_step<8>(sp, env->pnh, env->pnb, env->err, env->queue, env->queue_launches); break;
        default: return -1;
    }
    return (int)(env->err[0] | (env->err[1] << 1));
}

// shuffle_orders_kernel on HOST arrays (the shared-memory variant when use_smem != 0)
void fastace_emu_shuffle(int E, int P, int F, uint32_t seed, int restart, int steps, uint64_t* rng, int32_t* state_person,
                         int32_t* state_firm, int32_t* out_person, int32_t* out_firm, uint16_t* out_person16, uint16_t* out_firm16,
                         int use_smem) {
    ShuffleParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.E = E; sp.P = P; sp.F = F; sp.seed = seed; sp.restart = restart; sp.steps = steps;
    sp.rng_state = rng; sp.state_person = state_person; sp.state_firm = state_firm;
    sp.out_person = out_person; sp.out_firm = out_firm; sp.out_person16 = out_person16; sp.out_firm16 = out_firm16;
    sp.use_smem = use_smem;
    const unsigned blocks = (unsigned)((E + kShuffleThreads - 1) / kShuffleThreads);
    emu::launch(shuffle_orders_kernel, blocks, (unsigned)kShuffleThreads,
                use_smem ? (size_t)(P + F) * kShuffleThreads * sizeof(uint16_t) : 0, sp);
}

// the three kernels of segmented_sort.cuh on HOST arrays; seg = exclusive prefix of the per-firm totals, seg[F] = #requests
void fastace_emu_segmented_sort(const uint16_t* key_in, const uint32_t* val_in, uint16_t* key_out, uint32_t* val_out,
                                const uint32_t* seg, uint32_t* hist, int n, int F, int chunk) {
    SortParams sp;
    sp.key_in = key_in; sp.val_in = val_in; sp.key_out = key_out; sp.val_out = val_out; sp.seg = seg; sp.hist = hist;
    sp.n = n; sp.F = F; sp.chunk = chunk; sp.chunks = (n + chunk - 1) / chunk;
    const size_t smem = (size_t)(F + 1) * sizeof(uint32_t);
    if (sp.chunks == 0) return;
    emu::launch(sort_chunk_hist, (unsigned)sp.chunks, 32u, smem, sp);
    emu::launch(sort_bin_scan, (unsigned)((F + 1 + kSortScanWarps - 1) / kSortScanWarps), (unsigned)(32 * kSortScanWarps), 0, sp);
    emu::launch(sort_chunk_scatter, (unsigned)sp.chunks, 32u, smem, sp);
}

// expand_packed_kernel on HOST arrays: one list
void fastace_emu_expand_packed(int agents, int S, int bits, int bytes, const uint8_t* packed, uint8_t* idx, uint16_t* take) {
    ExpandParams a = {agents, S, bits, bytes, packed, idx, take};
    ExpandParams none = a; none.agents = 0;
    if (agents > 0) emu::launch(expand_packed_kernel, (unsigned)((agents + 255) / 256), 256u, 0, a, none);
}

// pow_reward of common.cuh (the reward path's short-polynomial pow), for the accuracy test
double fastace_emu_pow_reward(double x, double y) { return pow_reward(x, y); }

// 0 ascending, 1 descending, 2 scrambled: the order in which the emulated blocks of every launch run
void fastace_emu_block_order(int order) { emu::block_order() = order; }

// event counters of the kernels since the last call (common.cuh: kStat*); resets them
void fastace_emu_stats(unsigned long long* out) {
    for (int i = 0; i < kStatCount; i++) { out[i] = emu::stats()[i]; emu::stats()[i] = 0; }
}

}  // extern "C"
