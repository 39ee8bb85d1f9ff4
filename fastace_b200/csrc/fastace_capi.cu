// C ABI of libfastace_b200.so: env lifetime, state I/O and the step launchers.
// Declarations and the reference interfaces each entry point replaces: include/fastace_b200.h.
// There is NO CPU fallback anywhere in this file: without a CUDA device every env call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/fastace_b200.h"
#include "fastace_internal.h"
#include "step_kernel.cuh"
#include "match_update_kernels.cuh"
#include "shuffle_kernel.cuh"
#include "stats_kernel.cuh"
#include "segmented_sort.cuh"
#include "packed_kernel.cuh"
#include "large_economy.cuh"
#include "mlp_stack.cuh"
#include "policy_kernels.cuh"
#include "layer_kernels.cuh"

namespace fastace {

struct FieldDesc {
    size_t offset;   // offset of the pointer member inside its struct
    size_t elem;     // element size in bytes
    size_t count;    // elements for the env's dims
};

static std::vector<FieldDesc> state_fields(const fastace_dims_t& d) {
    const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods;
    const size_t cap = F * G;
    return {
        {offsetof(fastace_state_t, p_money), 8, E * P},
        {offsetof(fastace_state_t, p_inv), 8, E * G * P},
        {offsetof(fastace_state_t, p_labor), 8, E * P},
        {offsetof(fastace_state_t, p_util_tfp), 8, E * P},
        {offsetof(fastace_state_t, p_util_share), 8, E * (G + 1) * P},
        {offsetof(fastace_state_t, p_util_rho), 8, E * P},
        {offsetof(fastace_state_t, f_money), 8, E * F},
        {offsetof(fastace_state_t, f_inv), 8, E * G * F},
        {offsetof(fastace_state_t, f_labor), 8, E * F},
        {offsetof(fastace_state_t, f_last_money), 8, E * F},
        {offsetof(fastace_state_t, f_prod_tfp), 8, E * G * F},
        {offsetof(fastace_state_t, f_prod_share), 8, E * G * (G + 1) * F},
        {offsetof(fastace_state_t, f_prod_rho), 8, E * G * F},
        {offsetof(fastace_state_t, m_count), 4, E},
        {offsetof(fastace_state_t, m_owner), 4, E * cap},
        {offsetof(fastace_state_t, m_good), 4, E * cap},
        {offsetof(fastace_state_t, m_left), 4, E * cap},
        {offsetof(fastace_state_t, m_taken), 4, E * cap},
        {offsetof(fastace_state_t, m_price), 8, E * cap},
        {offsetof(fastace_state_t, j_count), 4, E},
        {offsetof(fastace_state_t, j_owner), 4, E * F},
        {offsetof(fastace_state_t, j_left), 4, E * F},
        {offsetof(fastace_state_t, j_taken), 4, E * F},
        {offsetof(fastace_state_t, j_wage), 8, E * F},
        {offsetof(fastace_state_t, p_util_theta), 8, E * (G + 1) * P},
        {offsetof(fastace_state_t, f_prod_theta), 8, E * G * (G + 1) * F},
    };
}

static std::vector<FieldDesc> action_fields(const fastace_dims_t& d) {
    const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods, S = d.stack_size;
    return {
        {offsetof(fastace_actions_t, perm_person), 4, E * P},
        {offsetof(fastace_actions_t, perm_firm), 4, E * F},
        {offsetof(fastace_actions_t, p_job_idx), 4, E * S * P},
        {offsetof(fastace_actions_t, p_job_take), 1, E * S * P},
        {offsetof(fastace_actions_t, p_good_idx), 4, E * S * P},
        {offsetof(fastace_actions_t, p_good_take), 1, E * S * P},
        {offsetof(fastace_actions_t, p_consume), 4, E * G * P},
        {offsetof(fastace_actions_t, f_good_idx), 4, E * S * F},
        {offsetof(fastace_actions_t, f_good_take), 1, E * S * F},
        {offsetof(fastace_actions_t, f_prod), 4, E * G * F},
        {offsetof(fastace_actions_t, f_offer_amt), 4, E * G * F},
        {offsetof(fastace_actions_t, f_offer_price), 4, E * G * F},
        {offsetof(fastace_actions_t, f_job_labor), 4, E * F},
        {offsetof(fastace_actions_t, f_job_wage), 4, E * F},
    };
}

static int bits_for(int max_index_exclusive) {   // smallest width whose all-ones value is not a valid index
    int b = 1;
    while (((1 << b) - 1) < max_index_exclusive) b++;
    return b;
}
struct PackedLayout { int bits_job, bytes_job, bits_good, bytes_good; };
static PackedLayout packed_layout(const fastace_dims_t& d) {
    PackedLayout L;
    L.bits_job = bits_for(d.num_firms);
    L.bits_good = bits_for(d.num_firms * d.num_goods);
    L.bytes_job = (d.stack_size * L.bits_job + 7) / 8;
    L.bytes_good = (d.stack_size * L.bits_good + 7) / 8;
    return L;
}
static std::vector<FieldDesc> packed_fields(const fastace_dims_t& d) {
    const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods;
    const PackedLayout L = packed_layout(d);
    return {
        {offsetof(fastace_actions_packed_t, perm_person), 2, E * P},
        {offsetof(fastace_actions_packed_t, perm_firm), 2, E * F},
        {offsetof(fastace_actions_packed_t, p_job_idx), 1, E * P * (size_t)L.bytes_job},
        {offsetof(fastace_actions_packed_t, p_good_idx), 1, E * P * (size_t)L.bytes_good},
        {offsetof(fastace_actions_packed_t, p_consume), 4, E * G * P},
        {offsetof(fastace_actions_packed_t, f_good_idx), 1, E * F * (size_t)L.bytes_good},
        {offsetof(fastace_actions_packed_t, f_prod), 4, E * G * F},
        {offsetof(fastace_actions_packed_t, f_offer_amt), 4, E * G * F},
        {offsetof(fastace_actions_packed_t, f_offer_price), 4, E * G * F},
        {offsetof(fastace_actions_packed_t, f_job_labor), 4, E * F},
        {offsetof(fastace_actions_packed_t, f_job_wage), 4, E * F},
    };
}

static std::vector<FieldDesc> compact_fields(const fastace_dims_t& d) {
    const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods, S = d.stack_size;
    return {
        {offsetof(fastace_actions_compact_t, perm_person), 2, E * P},
        {offsetof(fastace_actions_compact_t, perm_firm), 2, E * F},
        {offsetof(fastace_actions_compact_t, p_job_idx), 1, E * S * P},
        {offsetof(fastace_actions_compact_t, p_job_take), 2, E * P},
        {offsetof(fastace_actions_compact_t, p_good_idx), 1, E * S * P},
        {offsetof(fastace_actions_compact_t, p_good_take), 2, E * P},
        {offsetof(fastace_actions_compact_t, p_consume), 4, E * G * P},
        {offsetof(fastace_actions_compact_t, f_good_idx), 1, E * S * F},
        {offsetof(fastace_actions_compact_t, f_good_take), 2, E * F},
        {offsetof(fastace_actions_compact_t, f_prod), 4, E * G * F},
        {offsetof(fastace_actions_compact_t, f_offer_amt), 4, E * G * F},
        {offsetof(fastace_actions_compact_t, f_offer_price), 4, E * G * F},
        {offsetof(fastace_actions_compact_t, f_job_labor), 4, E * F},
        {offsetof(fastace_actions_compact_t, f_job_wage), 4, E * F},
    };
}

static std::vector<FieldDesc> out_fields(const fastace_dims_t& d) {
    const size_t E = d.num_econ, P = d.num_persons, F = d.num_firms, G = d.num_goods, S = d.stack_size;
    const size_t cap = F * G;
    return {
        {offsetof(fastace_step_out_t, p_reward), 8, E * P},
        {offsetof(fastace_step_out_t, f_profit), 8, E * F},
        {offsetof(fastace_step_out_t, p_job_ok), 1, E * S * P},
        {offsetof(fastace_step_out_t, p_good_ok), 1, E * S * P},
        {offsetof(fastace_step_out_t, f_good_ok), 1, E * S * F},
        {offsetof(fastace_step_out_t, old_m_left), 4, E * cap},
        {offsetof(fastace_step_out_t, old_m_taken), 4, E * cap},
        {offsetof(fastace_step_out_t, old_j_left), 4, E * F},
        {offsetof(fastace_step_out_t, old_j_taken), 4, E * F},
    };
}

template <typename T>
static inline void*& member(T* s, size_t off) { return *reinterpret_cast<void**>(reinterpret_cast<char*>(s) + off); }
template <typename T>
static inline void* member(const T* s, size_t off) { return *reinterpret_cast<void* const*>(reinterpret_cast<const char*>(s) + off); }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace fastace

struct fastace_env {
    fastace_dims_t dims;
    int device;
    uint32_t time;
    int util_kind, prod_kind;
    uint64_t launches;
    void* state_block;      // one allocation holding every state array
    fastace_state_t dstate; // device pointers into state_block
    // double-buffered device staging for the host-pointer calls
    void* act_block[2];
    fastace_actions_t dact[2];
    void* cact_block[2];
    fastace_actions_compact_t dcact[2];
    void* pact_block[2];
    fastace_actions_packed_t dpact[2];
    void* out_block[2];
    fastace_step_out_t dout[2];
    cudaStream_t stream;        // kernels of the host-pointer calls
    cudaStream_t copy_in, copy_out;
    cudaEvent_t h2d_done[2], kern_done[2], d2h_done[2];
    bool buf_used[2];
    bool have_host_streams;
    uint64_t host_steps;
    size_t smem_bytes;        // serial fused kernel
    size_t match_smem_bytes;  // match_kernel
    uint8_t* scr_pnh;         // [E][P]    match_kernel -> update_kernel
    uint8_t* scr_pnb;         // [E][G][P]
    uint32_t* done_queue;     // completion queue between the two kernels of a full step: [E] slots + the ticket counter
    uint32_t queue_launches;  // full steps launched through the queue (tickets handed out = queue_launches * E, mod 2^32)
    // visiting orders generated on the device (fastace_env_shuffle_orders)
    uint64_t* ord_rng;        // [E]    one minstd_rand0 per economy
    int32_t* ord_person;      // [E][P] cumulative order of the persons
    int32_t* ord_firm;        // [E][F]
    bool ord_started;
    bool ord_restart_pending; // a steps = 0 (re)start: the next shuffle seeds from ord_seed
    uint32_t ord_seed;
    uint32_t* err_host;       // host-mapped error words the kernels raise (match_kernel.cuh: kDevErr*)
    uint32_t* err_dev;        // the same words as the device sees them
    cudaEvent_t ev[3];        // FASTACE_STEP_PROFILE
    bool have_ev;
    double prof_match_ms, prof_update_ms;
    uint64_t prof_steps;
    // large-economy path (large_economy.cuh)
    bool large_only;          // dims beyond the warp-per-economy kernels: every step takes the large path
    bool mid_large;           // the phase-wise step in progress runs on the large-economy path
    int mid_step;             // 0 between steps; 1 after PERSONS_TRADE (CONSUME is due); 2 after the person phase (FIRMS is due)
    bool have_large;
    int large_coop_blocks_p, large_coop_blocks_f;   // co-resident CTAs of the two cooperative iteration kernels
    void* large_block;
    fastace::LargeScratch large_sc;
    uint32_t* sort_hist;      // segmented sort: per-chunk histograms
    size_t sort_chunk;
};

#define FASTACE_CUDA_CHECK(expr)                                                              \
    do {                                                                                      \
        cudaError_t _err = (expr);                                                            \
        if (_err != cudaSuccess) {                                                            \
            fastace::set_error(std::string(#expr) + ": " + cudaGetErrorString(_err));         \
            return FASTACE_ERR_CUDA;                                                          \
        }                                                                                     \
    } while (0)

using namespace fastace;

// A kernel that could not finish its fixed-point iteration within the round cap leaves a word in host-mapped
// memory: such a step is an error, never a result.
static int check_device_errors(const fastace_env_t* env) {
    if (!env->err_host) return FASTACE_OK;
    volatile const uint32_t* w = env->err_host;
    if (w[kDevErrQueue]) {
        set_error("update_kernel: gave up waiting on the completion queue of match_kernel (state is not a valid step result)");
        return FASTACE_ERR_CUDA;
    }
    if (w[kDevErrRounds] | w[kDevErrLargeRounds]) {
        set_error(w[kDevErrRounds] ? "match_kernel: a window's fixed-point iteration hit the round cap (state is not a valid step result)"
                                   : "large-economy path: the iteration hit the round cap (state is not a valid step result)");
        return FASTACE_ERR_NOT_CONVERGED;
    }
    return FASTACE_OK;
}

template <typename StructT>
static int carve(const std::vector<FieldDesc>& fields, StructT* s, void** block, bool zero) {
    size_t total = 0;
    for (auto& f : fields) total += align_up(f.elem * (f.count ? f.count : 1), 256);
    FASTACE_CUDA_CHECK(cudaMalloc(block, total));
    if (zero) FASTACE_CUDA_CHECK(cudaMemset(*block, 0, total));
    size_t off = 0;
    for (auto& f : fields) {
        member(s, f.offset) = static_cast<char*>(*block) + off;
        off += align_up(f.elem * (f.count ? f.count : 1), 256);
    }
    return FASTACE_OK;
}

// Launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream is
// still draining (both step kernels call griddepcontrol.wait before they touch the predecessor's outputs), so launch
// latency and the ramp of one kernel hide under the tail of the other.
template <typename ParamsT>
static cudaError_t launch_dependent(const void* fn, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, ParamsT* params) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    void* args[] = {(void*)params};
    return cudaLaunchKernelExC(&cfg, fn, args);
}

typedef void (*serial_fn)(const StepParams);
typedef void (*match_fn)(const MatchParams);
typedef void (*update_fn)(const UpdateParams);
// match: the generic kernel; match_compact[modulo]: specialised for the compact encoding without per-person success flags
// update_ces: both function families CES (the reference's default)
// update[ces][queue]: whole-grid / completion-queue assignment, generic / CES-only function families
struct KernelSet { serial_fn serial; match_fn match; update_fn update[2][2]; match_fn match_compact[4]; };
template <int G>
static KernelSet kernels_of() {
    return {step_kernel<G>, match_kernel<G, kModeGeneric>,
            {{update_kernel<G, false, false>, update_kernel<G, false, true>}, {update_kernel<G, true, false>, update_kernel<G, true, true>}},
            {match_kernel<G, kModeCompact>, match_kernel<G, kModeCompact | kModeModulo>,
             match_kernel<G, kModeCompact | kModeSmall>, match_kernel<G, kModeCompact | kModeModulo | kModeSmall>}};
}
static KernelSet kernels_for_goods(int G) {
    switch (G) {
        case 1: return kernels_of<1>();
        case 2: return kernels_of<2>();
        case 3: return kernels_of<3>();
        case 4: return kernels_of<4>();
        case 5: return kernels_of<5>();
        case 6: return kernels_of<6>();
        case 7: return kernels_of<7>();
        case 8: return kernels_of<8>();
        default: return {nullptr, nullptr, {{nullptr, nullptr}, {nullptr, nullptr}}, {nullptr, nullptr, nullptr, nullptr}};
    }
}

extern "C" {

int fastace_device_count(void) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
        set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
        return FASTACE_ERR_NO_DEVICE;
    }
    return n;
}

int fastace_env_create(const fastace_dims_t* dims, int device, fastace_env_t** out_env) {
    if (!dims || !out_env) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *out_env = nullptr;
    const fastace_dims_t d = *dims;
    if (d.num_econ < 1 || d.num_persons < 0 || d.num_firms < 1 || d.num_goods < 1 ||
        d.num_goods > FASTACE_MAX_GOODS || d.stack_size < 0 || d.stack_size > FASTACE_MAX_STACK) {
        set_error("unsupported dims: need E>=1, F>=1, 1<=G<=8, 0<=S<=16");
        return FASTACE_ERR_INVALID;
    }
    if (d.num_firms > 56000 || d.num_persons > (1 << 20) || (size_t)2 * d.stack_size * d.num_persons >= (size_t)1 << 27) {
        set_error("unsupported dims: need F <= 56000, P <= 2^20");
        return FASTACE_ERR_INVALID;
    }
    bool large_only = d.num_firms * d.num_goods > 254 || d.num_persons > 65535;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        set_error("no CUDA device visible: fastace_b200 has no CPU path");
        return FASTACE_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { set_error("bad device index"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(device));

    SmemLayout L; std::memset(&L, 0, sizeof(L));
    MatchLayout ML; std::memset(&ML, 0, sizeof(ML));
    if (!large_only) {
        L = make_layout(d.num_persons, d.num_firms, d.num_goods, d.stack_size);
        ML = make_match_layout(d.num_persons, d.num_firms, d.num_goods, d.stack_size);
        int max_optin = 0;
        FASTACE_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        if (L.total > max_optin || ML.total > max_optin) large_only = true;   // books do not fit shared memory
    }
    if (!large_only) {
        const KernelSet ks = kernels_for_goods(d.num_goods);
        if (L.total > 48 * 1024)
            FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)ks.serial, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        if (ML.total > 48 * 1024) {
            FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)ks.match, cudaFuncAttributeMaxDynamicSharedMemorySize, ML.total));
            for (int m = 0; m < 4; m++)
                FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)ks.match_compact[m], cudaFuncAttributeMaxDynamicSharedMemorySize, ML.total));
        }
    }

    fastace_env* env = new (std::nothrow) fastace_env();
    if (!env) { set_error("out of host memory"); return FASTACE_ERR_ALLOC; }
    std::memset(env, 0, sizeof(*env));
    env->dims = d;
    env->device = device;
    env->smem_bytes = (size_t)L.total;
    env->match_smem_bytes = (size_t)ML.total;
    env->large_only = large_only;
    int rc = carve(state_fields(d), &env->dstate, &env->state_block, true);
    if (rc != FASTACE_OK) { delete env; return rc; }
    {
        const size_t np = (size_t)d.num_econ * d.num_persons;
        cudaError_t e1 = cudaMalloc((void**)&env->scr_pnh, np ? np : 1);
        cudaError_t e2 = cudaMalloc((void**)&env->scr_pnb, np ? np * d.num_goods : 1);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("cudaMalloc of matching scratch failed");
            cudaGetLastError();
            fastace_env_destroy(env); return FASTACE_ERR_ALLOC;
        }
        // FASTACE_NO_QUEUE=1 (diagnostics): update_kernel waits for the whole match grid, as the phase-wise calls do
        const char* noq = getenv("FASTACE_NO_QUEUE");
        if (!(noq && noq[0] == '1') && d.num_econ > 0 && (uint32_t)d.num_econ <= kQueueEconMask) {
            const size_t qbytes = ((size_t)d.num_econ + 1) * sizeof(uint32_t);
            if (cudaMalloc((void**)&env->done_queue, qbytes) != cudaSuccess || cudaMemset(env->done_queue, 0, qbytes) != cudaSuccess) {
                set_error("cudaMalloc of the completion queue failed");
                cudaGetLastError();
                fastace_env_destroy(env); return FASTACE_ERR_ALLOC;
            }
        }
    }
    {
        cudaError_t e1 = cudaHostAlloc((void**)&env->err_host, 4 * sizeof(uint32_t), cudaHostAllocMapped);
        cudaError_t e2 = e1 == cudaSuccess ? cudaHostGetDevicePointer((void**)&env->err_dev, env->err_host, 0) : e1;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("allocation of the device error words failed");
            cudaGetLastError();
            fastace_env_destroy(env); return FASTACE_ERR_ALLOC;
        }
        std::memset(env->err_host, 0, 4 * sizeof(uint32_t));
    }
    cudaError_t err = cudaStreamCreateWithFlags(&env->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) { set_error(cudaGetErrorString(err)); fastace_env_destroy(env); return FASTACE_ERR_CUDA; }
    *out_env = env;
    return FASTACE_OK;
}

int fastace_env_destroy(fastace_env_t* env) {
    if (!env) return FASTACE_OK;
    cudaSetDevice(env->device);
    if (env->stream) cudaStreamDestroy(env->stream);
    if (env->state_block) cudaFree(env->state_block);
    for (int b = 0; b < 2; b++) {
        if (env->act_block[b]) cudaFree(env->act_block[b]);
        if (env->cact_block[b]) cudaFree(env->cact_block[b]);
        if (env->pact_block[b]) cudaFree(env->pact_block[b]);
        if (env->out_block[b]) cudaFree(env->out_block[b]);
    }
    if (env->have_host_streams) {
        cudaStreamDestroy(env->copy_in); cudaStreamDestroy(env->copy_out);
        for (int b = 0; b < 2; b++) { cudaEventDestroy(env->h2d_done[b]); cudaEventDestroy(env->kern_done[b]); cudaEventDestroy(env->d2h_done[b]); }
    }
    if (env->have_ev) for (int i = 0; i < 3; i++) cudaEventDestroy(env->ev[i]);
    if (env->scr_pnh) cudaFree(env->scr_pnh);
    if (env->scr_pnb) cudaFree(env->scr_pnb);
    if (env->done_queue) cudaFree(env->done_queue);
    if (env->err_host) cudaFreeHost(env->err_host);
    if (env->ord_rng) cudaFree(env->ord_rng);
    if (env->ord_person) cudaFree(env->ord_person);
    if (env->ord_firm) cudaFree(env->ord_firm);
    if (env->large_block) cudaFree(env->large_block);
    if (env->sort_hist) cudaFree(env->sort_hist);
    delete env;
    return FASTACE_OK;
}

int fastace_env_set_function_kinds(fastace_env_t* env, int util_kind, int prod_kind) {
    if (!env) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    if (util_kind < 0 || util_kind > FASTACE_FN_LINEAR || prod_kind < 0 || prod_kind > FASTACE_FN_LINEAR) {
        set_error("unknown function kind"); return FASTACE_ERR_INVALID;
    }
    env->util_kind = util_kind; env->prod_kind = prod_kind;
    return FASTACE_OK;
}

int fastace_env_dims(const fastace_env_t* env, fastace_dims_t* out_dims) {
    if (!env || !out_dims) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *out_dims = env->dims;
    return FASTACE_OK;
}

int fastace_env_time(const fastace_env_t* env, uint32_t* out_time) {
    if (!env || !out_time) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *out_time = env->time;
    return FASTACE_OK;
}

int fastace_env_set_state(fastace_env_t* env, const fastace_state_t* host_state, uint32_t time) {
    if (!env || !host_state) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    // steps may still be running on non-blocking streams (the env's own, or the caller's): the state they read
    // must not be overwritten under them
    FASTACE_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& f : state_fields(env->dims)) {
        const void* src = member(host_state, f.offset);
        if (!src || f.count == 0) continue;
        FASTACE_CUDA_CHECK(cudaMemcpy(member(&env->dstate, f.offset), src, f.elem * f.count, cudaMemcpyHostToDevice));
    }
    env->time = time;
    env->mid_step = 0;
    if (env->err_host) std::memset(env->err_host, 0, 4 * sizeof(uint32_t));   // a fresh state: earlier errors are void
    return FASTACE_OK;
}

int fastace_env_get_state(const fastace_env_t* env, fastace_state_t* host_state) {
    if (!env || !host_state) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    FASTACE_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& f : state_fields(env->dims)) {
        void* dst = member(host_state, f.offset);
        if (!dst || f.count == 0) continue;
        FASTACE_CUDA_CHECK(cudaMemcpy(dst, member(&env->dstate, f.offset), f.elem * f.count, cudaMemcpyDeviceToHost));
    }
    return check_device_errors(env);
}

int fastace_env_device_state(const fastace_env_t* env, fastace_state_t* out_device_state) {
    if (!env || !out_device_state) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *out_device_state = env->dstate;
    return FASTACE_OK;
}

}  // extern "C"

// ---- large-economy path ------------------------------------------------------------------------------------
struct LargeKernels {
    void (*iterate_persons)(const LargeParams, int);
    void (*finalize_persons)(const LargeParams);
    void (*iterate_firms)(const LargeParams, int);
    void (*finalize_firms)(const LargeParams);
};
template <int G>
static LargeKernels large_kernels_of() {
    return {large_iterate_persons<G>, large_finalize_persons<G>, large_iterate_firms<G>, large_finalize_firms<G>};
}
static LargeKernels large_kernels_for_goods(int G) {
    switch (G) {
        case 1: return large_kernels_of<1>();
        case 2: return large_kernels_of<2>();
        case 3: return large_kernels_of<3>();
        case 4: return large_kernels_of<4>();
        case 5: return large_kernels_of<5>();
        case 6: return large_kernels_of<6>();
        case 7: return large_kernels_of<7>();
        default: return large_kernels_of<8>();
    }
}

// elements per chunk (= per warp) of the segmented sort: enough chunks to fill the GPU, chunks long enough to amortise
// the per-chunk bin table, and a histogram matrix of at most 96 MB
static size_t sort_chunk_for(size_t n, size_t F) {
    size_t chunk = 256;
    while (chunk < 4096 && n / chunk > 512) chunk *= 2;
    const size_t budget_words = (size_t)24 << 20;
    while ((n + chunk - 1) / chunk * (F + 1) > budget_words) chunk *= 2;
    return chunk;
}

// stable sort of n (key, value) pairs by firm id: the events of firm f become the segment [seg[f], seg[f+1])
static int segmented_sort(fastace_env_t* env, const uint16_t* key_in, const uint32_t* val_in, uint16_t* key_out, uint32_t* val_out,
                          const uint32_t* seg, size_t n, cudaStream_t stream) {
    if (n == 0) return FASTACE_OK;
    SortParams sp;
    sp.key_in = key_in; sp.val_in = val_in; sp.key_out = key_out; sp.val_out = val_out; sp.seg = seg; sp.hist = env->sort_hist;
    const size_t chunk = std::min(env->sort_chunk, sort_chunk_for(n, (size_t)env->dims.num_firms));   // never more chunks than allocated
    sp.n = (int)n; sp.F = env->dims.num_firms; sp.chunk = (int)chunk; sp.chunks = (int)((n + chunk - 1) / chunk);
    const size_t smem = (size_t)(sp.F + 1) * sizeof(uint32_t);
    sort_chunk_hist<<<sp.chunks, 32, smem, stream>>>(sp);
    sort_bin_scan<<<(sp.F + 1 + kSortScanWarps - 1) / kSortScanWarps, 32 * kSortScanWarps, 0, stream>>>(sp);
    sort_chunk_scatter<<<sp.chunks, 32, smem, stream>>>(sp);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    env->launches += 3;
    return FASTACE_OK;
}

static int ensure_large_scratch(fastace_env_t* env) {
    if (env->have_large) return FASTACE_OK;
    const size_t P = env->dims.num_persons, F = env->dims.num_firms, G = env->dims.num_goods, S = env->dims.stack_size;
    const size_t R = 2 * S * P, RF = S * F, cap = F * G;
    LargeScratch& sc = env->large_sc;
    struct Item { void** ptr; size_t bytes; };
    std::vector<Item> items = {
        {(void**)&sc.rank_f, 4 * F}, {(void**)&sc.own_offer, 4 * cap}, {(void**)&sc.own_job, 4 * F},
        {(void**)&sc.req_n, 4 * R}, {(void**)&sc.want, R}, {(void**)&sc.ok, R},
        {(void**)&sc.key_in, 2 * R}, {(void**)&sc.key_out, 2 * R}, {(void**)&sc.val_in, 4 * R}, {(void**)&sc.val_out, 4 * R},
        {(void**)&sc.hist, 4 * F}, {(void**)&sc.seg, 4 * (F + 1)},
        {(void**)&sc.pm_money, 8 * P}, {(void**)&sc.pm_hires, P}, {(void**)&sc.pm_bought, G * P},
        {(void**)&sc.fm_money, 8 * F}, {(void**)&sc.fm_labor, 8 * F}, {(void**)&sc.fm_inv, 8 * G * F},
        {(void**)&sc.fm_left, 4 * cap}, {(void**)&sc.fm_taken, 4 * cap}, {(void**)&sc.fm_jleft, 4 * F}, {(void**)&sc.fm_jtaken, 4 * F},
        {(void**)&sc.freq_n, 4 * RF}, {(void**)&sc.fwant, RF}, {(void**)&sc.fok, RF},
        {(void**)&sc.fkey_in, 2 * RF}, {(void**)&sc.fkey_out, 2 * RF}, {(void**)&sc.fval_in, 4 * RF}, {(void**)&sc.fval_out, 4 * RF},
        {(void**)&sc.fhist, 4 * F}, {(void**)&sc.fseg, 4 * (F + 1)},
        {(void**)&sc.ff_profit, 8 * F}, {(void**)&sc.ff_money, 8 * F}, {(void**)&sc.ff_last, 8 * F}, {(void**)&sc.ff_inv, 8 * G * F},
        {(void**)&sc.ff_left, 4 * cap}, {(void**)&sc.ff_taken, 4 * cap},
        {(void**)&sc.post_lots, 4 * cap}, {(void**)&sc.post_jlots, 4 * F}, {(void**)&sc.changed, 32},
        {(void**)&sc.req_firm, 2 * R}, {(void**)&sc.req_pos, 4 * R}, {(void**)&sc.want_sorted, R}, {(void**)&sc.ok_sorted, R}, {(void**)&sc.dirty_person, P}, {(void**)&sc.dirty_firm, F},
        {(void**)&sc.post_base_m, 4 * F}, {(void**)&sc.post_base_j, 4 * F},
    };
    size_t total = 0;
    for (auto& it : items) total += align_up(it.bytes ? it.bytes : 1, 256);
    FASTACE_CUDA_CHECK(cudaMalloc(&env->large_block, total));
    FASTACE_CUDA_CHECK(cudaMemset(env->large_block, 0, total));
    size_t off = 0;
    for (auto& it : items) { *it.ptr = static_cast<char*>(env->large_block) + off; off += align_up(it.bytes ? it.bytes : 1, 256); }
    {
        // per-chunk histograms of the segmented sort (segmented_sort.cuh): chunks x (F + 1) words
        const size_t n_max = R > RF ? R : RF;
        env->sort_chunk = sort_chunk_for(n_max, F);
        // room for the chunk count of ANY list length up to n_max (a shorter list takes shorter chunks, see sort_chunk_for)
        const size_t by_budget = ((size_t)24 << 20) / (F + 1) + 1;
        const size_t chunks = std::min(by_budget, std::max<size_t>(513, n_max / 4096 + 1));
        FASTACE_CUDA_CHECK(cudaMalloc((void**)&env->sort_hist, sizeof(uint32_t) * chunks * (F + 1)));
        const size_t smem = (F + 1) * sizeof(uint32_t);
        if (smem > 48 * 1024) {
            FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)sort_chunk_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)sort_chunk_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    {
        const LargeKernels lk = large_kernels_for_goods(env->dims.num_goods);
        int sms = 0, occ_p = 0, occ_f = 0;
        FASTACE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, env->device));
        FASTACE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, (const void*)lk.iterate_persons, kLargeThreads, 0));
        FASTACE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, (const void*)lk.iterate_firms, kLargeThreads, 0));
        if (occ_p < 1 || occ_f < 1) { set_error("large-economy iteration kernels do not fit an SM"); return FASTACE_ERR_CUDA; }
        env->large_coop_blocks_p = sms * occ_p;
        env->large_coop_blocks_f = sms * occ_f;
    }
    env->have_large = true;
    return FASTACE_OK;
}

// pointer members of a [E][...] struct moved to economy e
template <typename StructT>
static void offset_to_economy(const std::vector<FieldDesc>& fields, StructT* s, int e, int E) {
    for (auto& f : fields) {
        void*& m = member(s, f.offset);
        if (m) m = static_cast<char*>(m) + (size_t)e * (f.count / (size_t)E) * f.elem;
    }
}

static int launch_step_large(fastace_env_t* env, const fastace_actions_t* dact, const fastace_step_out_t* dout,
                             uint32_t flags, cudaStream_t stream) {
    int rc = ensure_large_scratch(env);
    if (rc != FASTACE_OK) return rc;
    const fastace_dims_t d = env->dims;
    const int P = d.num_persons, F = d.num_firms, G = d.num_goods, S = d.stack_size, cap = F * G;
    const LargeKernels lk = large_kernels_for_goods(G);
    const int T = kLargeThreads;
    auto blocks = [T](size_t n) { return (unsigned)((n + T - 1) / T); };
    const int kMaxRounds = 1 << 16;
    for (int e = 0; e < d.num_econ; e++) {
        LargeParams lp;
        std::memset(&lp, 0, sizeof(lp));
        lp.sp.E = 1; lp.sp.P = P; lp.sp.F = F; lp.sp.S = S;
        lp.sp.flags = flags; lp.sp.time_before = env->time;
        lp.sp.util_kind = env->util_kind; lp.sp.prod_kind = env->prod_kind;
        lp.sp.st = env->dstate; lp.sp.ac = *dact; lp.sp.out = *dout;
        offset_to_economy(state_fields(d), &lp.sp.st, e, d.num_econ);
        offset_to_economy(action_fields(d), const_cast<fastace_actions_t*>(&lp.sp.ac), e, d.num_econ);
        offset_to_economy(out_fields(d), &lp.sp.out, e, d.num_econ);
        lp.sc = env->large_sc;
        lp.G = G;
        lp.dev_err = env->err_dev;
        const LargeScratch& sc = lp.sc;
        const bool ph_p = (flags & FASTACE_STEP_PERSONS) != 0, ph_t = (flags & FASTACE_STEP_PERSONS_TRADE) != 0;
        const bool ph_c = (flags & FASTACE_STEP_PERSONS_CONSUME) != 0, ph_f = (flags & FASTACE_STEP_FIRMS) != 0;
        const bool run_trades = !ph_c && !ph_f;           // job search + purchases
        const bool run_consume = !ph_t && !ph_f;          // consumption (with the trades, or on its own after TRADE)
        const bool run_firms = !ph_p && !ph_t && !ph_c;
        const size_t R = (size_t)2 * S * P;
        if (run_trades) {
            const size_t n_index = std::max((size_t)(cap > F ? cap : F), (size_t)P);
            large_index_books<<<blocks(n_index), T, 0, stream>>>(lp);
            large_index_books2<<<blocks(cap > F ? cap : F), T, 0, stream>>>(lp);
            env->launches += 2;
            // ---- person phase
            FASTACE_CUDA_CHECK(cudaMemsetAsync(sc.changed, 0, 32, stream));
            if (R > 0) {
                large_prep_persons<<<blocks(R), T, 0, stream>>>(lp);
                env->launches += 1;
            }
            large_scan<<<1, 1024, 0, stream>>>(sc.hist, sc.seg, F);
            env->launches += 1;
            // the segmented sort of the event lists: requests, enumerated in event order, stably by firm
            if (int src = segmented_sort(env, sc.key_in, sc.val_in, sc.key_out, sc.val_out, sc.seg, R, stream)) return src;
            if (R > 0) {
                large_index_events<<<blocks(R), T, 0, stream>>>(lp);
                env->launches += 1;
            }
            {
                // requester pass / firm pass rounds until nothing changes: ONE cooperative launch, no host round trip
                int max_rounds = kMaxRounds;
                void* args[] = {(void*)&lp, (void*)&max_rounds};
                const size_t want_blocks = std::max<size_t>(blocks(P), blocks((size_t)F * 32));
                const int grid = (int)std::min<size_t>(std::max<size_t>(want_blocks, 1), (size_t)env->large_coop_blocks_p);
                FASTACE_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)lk.iterate_persons, dim3(grid), dim3(T), args, 0, stream));
                env->launches += 1;
            }
            large_old_jobs<<<blocks(F), T, 0, stream>>>(lp);
            env->launches += 1;
            if (S > 0 && P > 0 && (lp.sp.out.p_job_ok || lp.sp.out.p_good_ok)) {
                large_person_flags<<<blocks((size_t)S * P), T, 0, stream>>>(lp);
                env->launches += 1;
            }
            if (!run_firms) {   // phase-wise: the firms as they stand after the person phase become visible
                large_publish_firms<<<blocks(cap > F ? cap : F), T, 0, stream>>>(lp);
                env->launches += 1;
            }
        }
        if ((run_trades || run_consume) && P > 0) {
            lk.finalize_persons<<<blocks(P), T, 0, stream>>>(lp);   // trades and/or consumption, by the call's flags
            env->launches += 1;
        }
        if (run_firms) {
            // ---- firm phase
            const size_t RF = (size_t)S * F;
            large_index_firms<<<blocks(F), T, 0, stream>>>(lp);
            FASTACE_CUDA_CHECK(cudaMemsetAsync(sc.changed + 4, 0, 16, stream));
            if (RF > 0) {
                large_prep_firms<<<blocks(RF), T, 0, stream>>>(lp);
                env->launches += 1;
            }
            large_scan<<<1, 1024, 0, stream>>>(sc.fhist, sc.fseg, F);
            if (int src = segmented_sort(env, sc.fkey_in, sc.fval_in, sc.fkey_out, sc.fval_out, sc.fseg, RF, stream)) return src;
            {
                int max_rounds = kMaxRounds;
                void* args[] = {(void*)&lp, (void*)&max_rounds};
                const int grid = (int)std::min<size_t>(std::max<size_t>(blocks(F), 1), (size_t)env->large_coop_blocks_f);
                FASTACE_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)lk.iterate_firms, dim3(grid), dim3(T), args, 0, stream));
            }
            env->launches += 3;
            if (RF > 0 && lp.sp.out.f_good_ok)
                FASTACE_CUDA_CHECK(cudaMemcpyAsync(lp.sp.out.f_good_ok, sc.fok, RF, cudaMemcpyDeviceToDevice, stream));
            lk.finalize_firms<<<blocks(cap), T, 0, stream>>>(lp);
            large_post_scan<<<1, 1024, 0, stream>>>(lp);
            large_post_write<<<blocks(cap > F ? cap : F), T, 0, stream>>>(lp);
            env->launches += 3;
        }
        FASTACE_CUDA_CHECK(cudaGetLastError());
    }
    return FASTACE_OK;
}

extern "C" {

static int launch_step(fastace_env_t* env, const fastace_actions_t* dact, const fastace_actions_compact_t* dcz,
                       const fastace_step_out_t* dout, uint32_t flags, cudaStream_t stream) {
    const bool ph_p = (flags & FASTACE_STEP_PERSONS) != 0, only_f = (flags & FASTACE_STEP_FIRMS) != 0;
    const bool ph_t = (flags & FASTACE_STEP_PERSONS_TRADE) != 0, ph_c = (flags & FASTACE_STEP_PERSONS_CONSUME) != 0;
    const bool only_p = ph_p || ph_t || ph_c;          // some part of the person phase, no firm phase
    if ((int)ph_p + (int)ph_t + (int)ph_c + (int)only_f > 1) { set_error("the phase flags are separate calls"); return FASTACE_ERR_INVALID; }
    if (int erc = check_device_errors(env)) return erc;   // an earlier step did not converge
    if ((only_p || only_f) && (dcz || (flags & FASTACE_STEP_SERIAL))) {
        set_error("phase-wise stepping takes the int32 action encoding and is not implemented by the serial kernel");
        return FASTACE_ERR_INVALID;
    }
    {
        const int need = only_f ? 2 : ph_c ? 1 : 0;    // what must have run before this call
        if (env->mid_step != need) {
            set_error(env->mid_step == 1 ? "FASTACE_STEP_PERSONS_TRADE has run: the next call must be FASTACE_STEP_PERSONS_CONSUME"
                      : env->mid_step == 2 ? "the person phase has run: the next call must be FASTACE_STEP_FIRMS"
                      : only_f ? "FASTACE_STEP_FIRMS must follow the person phase" : "FASTACE_STEP_PERSONS_CONSUME must follow FASTACE_STEP_PERSONS_TRADE");
            return FASTACE_ERR_INVALID;
        }
    }
    if (dcz) {
        for (auto& f : compact_fields(env->dims))
            if (f.count && !member(dcz, f.offset)) { set_error("actions: every array is mandatory"); return FASTACE_ERR_INVALID; }
        if (flags & FASTACE_STEP_SERIAL) { set_error("the serial kernel takes the int32 action encoding"); return FASTACE_ERR_INVALID; }
    } else {
        for (auto& f : action_fields(env->dims)) {
            const bool consume_field = f.offset == offsetof(fastace_actions_t, p_consume);
            const bool trade_field = f.offset == offsetof(fastace_actions_t, perm_person) || f.offset == offsetof(fastace_actions_t, p_job_idx) ||
                                     f.offset == offsetof(fastace_actions_t, p_job_take) || f.offset == offsetof(fastace_actions_t, p_good_idx) ||
                                     f.offset == offsetof(fastace_actions_t, p_good_take);
            const bool firm_field = !consume_field && !trade_field;
            // only the arrays the call reads are mandatory
            const bool read = (!only_p && !only_f) || (ph_p && !firm_field) || (ph_t && trade_field) || (ph_c && consume_field) || (only_f && firm_field);
            if (read && f.count && !member(dact, f.offset)) { set_error("actions: every array the call reads is mandatory"); return FASTACE_ERR_INVALID; }
        }
    }
    const bool needs_reward = !only_f && !ph_t, needs_profit = !only_p;
    if ((needs_reward && env->dims.num_persons > 0 && !dout->p_reward) || (needs_profit && !dout->f_profit)) {
        set_error("out: p_reward (consumption) and f_profit (firm phase) are mandatory");
        return FASTACE_ERR_INVALID;
    }
    if (env->large_only || (flags & FASTACE_STEP_LARGE)) {
        if (dcz) { set_error("the large-economy path takes the int32 action encoding"); return FASTACE_ERR_INVALID; }
        if (env->mid_step != 0 && !env->mid_large) { set_error("a phase-wise step must be finished on the path it was started on"); return FASTACE_ERR_INVALID; }
        if ((only_p || only_f) && env->dims.num_econ != 1) {
            set_error("phase-wise stepping on the large-economy path needs num_econ == 1 (its scratch holds one economy between calls)");
            return FASTACE_ERR_INVALID;
        }
        const int rc = launch_step_large(env, dact, dout, flags, stream);
        if (rc != FASTACE_OK) return rc;
        if (ph_t) env->mid_step = 1;
        else if (only_p) env->mid_step = 2;
        else { env->mid_step = 0; env->time += 1; }
        env->mid_large = env->mid_step != 0;
        return FASTACE_OK;
    }
    if (env->mid_step != 0 && env->mid_large) { set_error("a phase-wise step must be finished on the path it was started on"); return FASTACE_ERR_INVALID; }
    StepParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.E = env->dims.num_econ; sp.P = env->dims.num_persons; sp.F = env->dims.num_firms; sp.S = env->dims.stack_size;
    sp.flags = flags;
    sp.time_before = env->time;
    sp.util_kind = env->util_kind; sp.prod_kind = env->prod_kind;
    sp.st = env->dstate;
    if (dcz) {
        sp.compact = 1;
        sp.cz = *dcz;
        sp.ac.p_consume = dcz->p_consume; sp.ac.f_prod = dcz->f_prod; sp.ac.f_offer_amt = dcz->f_offer_amt;
        sp.ac.f_offer_price = dcz->f_offer_price; sp.ac.f_job_labor = dcz->f_job_labor; sp.ac.f_job_wage = dcz->f_job_wage;
    } else {
        sp.ac = *dact;
    }
    sp.out = *dout;
    const KernelSet ks = kernels_for_goods(env->dims.num_goods);
    if (flags & FASTACE_STEP_SERIAL) {
        ks.serial<<<sp.E, 32, env->smem_bytes, stream>>>(sp);
        FASTACE_CUDA_CHECK(cudaGetLastError());
        env->launches += 1;
    } else {
        const bool prof = (flags & FASTACE_STEP_PROFILE) != 0;
        if (prof && !env->have_ev) {
            for (int i = 0; i < 3; i++) FASTACE_CUDA_CHECK(cudaEventCreate(&env->ev[i]));
            env->have_ev = true;
        }
        MatchParams mp;
        mp.sp = sp; mp.scr_pnh = env->scr_pnh; mp.scr_pnb = env->scr_pnb; mp.dev_err = env->err_dev;
        mp.lay = make_match_layout(sp.P, sp.F, env->dims.num_goods, sp.S);
        // a full step hands finished economies from match_kernel to update_kernel through the completion queue
        const bool queue = env->done_queue != nullptr && !only_p && !only_f;
        mp.done_list = queue ? env->done_queue : nullptr;
        mp.done_count = queue ? env->done_queue + sp.E : nullptr;
        mp.ticket_base = env->queue_launches * (uint32_t)sp.E;
        mp.done_tag = (env->queue_launches % 255u) + 1u;
        if (prof) FASTACE_CUDA_CHECK(cudaEventRecord(env->ev[0], stream));
        if (!ph_c) {   // a consume-only call has no matching to do
            // the specialised kernel when the call is what it was compiled for (match_kernel.cuh: MODE)
            const bool special = dcz != nullptr && !only_p && !only_f && !dout->p_job_ok && !dout->p_good_ok && !dout->f_good_ok &&
                                 !dout->old_j_left && !dout->old_j_taken && !dout->old_m_left && !dout->old_m_taken;
            const bool small = sp.F * (env->dims.num_goods + 1) + 2 <= 32;    // both books and the firms in one pass
            const match_fn fn = special ? ks.match_compact[((flags & FASTACE_IDX_MODULO) ? 1 : 0) + (small ? 2 : 0)] : ks.match;
            FASTACE_CUDA_CHECK(launch_dependent((const void*)fn, dim3((unsigned)sp.E), dim3(32), env->match_smem_bytes, stream, &mp));
            FASTACE_CUDA_CHECK(cudaGetLastError());
            env->launches += 1;
        }
        if (prof) FASTACE_CUDA_CHECK(cudaEventRecord(env->ev[1], stream));
        const int ces = (sp.util_kind == FASTACE_FN_CES && sp.prod_kind == FASTACE_FN_CES) ? 1 : 0;
        UpdateParams up;
        up.sp = sp; up.scr_pnh = env->scr_pnh; up.scr_pnb = env->scr_pnb;
        up.done_list = mp.done_list; up.done_tag = mp.done_tag; up.dev_err = env->err_dev;
        const size_t persons = (size_t)sp.E * sp.P;
        // update_kernel: blocks [0, firm_blocks) do the firms, the rest the persons; a phase call launches only its part
        const int person_blocks = only_f ? 0 : (int)((persons + kUpdateThreads - 1) / kUpdateThreads);
        up.firm_blocks = only_p ? 0 : (sp.E + kUpdateThreads / 32 - 1) / (kUpdateThreads / 32);
        if (queue) {
            const int qgroups = (sp.E + kQueueGroup - 1) / kQueueGroup;
            up.group_person_blocks = (kQueueGroup * sp.P + kUpdateThreads - 1) / kUpdateThreads;
            const int qblocks = qgroups * (kQueueGroup / (kUpdateThreads / 32) + up.group_person_blocks);
            FASTACE_CUDA_CHECK(launch_dependent((const void*)ks.update[ces][1], dim3((unsigned)qblocks), dim3(kUpdateThreads), 0, stream, &up));
            env->launches += 1;
            env->queue_launches += 1;
        } else if (person_blocks + up.firm_blocks > 0) {
            FASTACE_CUDA_CHECK(launch_dependent((const void*)ks.update[ces][0], dim3((unsigned)(person_blocks + up.firm_blocks)), dim3(kUpdateThreads), 0, stream, &up));
            env->launches += 1;
        }
        FASTACE_CUDA_CHECK(cudaGetLastError());
        if (prof) {
            FASTACE_CUDA_CHECK(cudaEventRecord(env->ev[2], stream));
            FASTACE_CUDA_CHECK(cudaEventSynchronize(env->ev[2]));
            float a = 0.f, b = 0.f;
            FASTACE_CUDA_CHECK(cudaEventElapsedTime(&a, env->ev[0], env->ev[1]));
            FASTACE_CUDA_CHECK(cudaEventElapsedTime(&b, env->ev[1], env->ev[2]));
            env->prof_match_ms += a; env->prof_update_ms += b; env->prof_steps += 1;
        }
    }
    if (ph_t) env->mid_step = 1;               // consumption is due
    else if (only_p) env->mid_step = 2;        // the step completes with the FASTACE_STEP_FIRMS call
    else { env->mid_step = 0; env->time += 1; }
    return FASTACE_OK;
}

int fastace_env_step_device(fastace_env_t* env, const fastace_actions_t* actions, const fastace_step_out_t* out,
                            uint32_t flags, void* cuda_stream) {
    if (!env || !actions || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    return launch_step(env, actions, nullptr, out, flags, static_cast<cudaStream_t>(cuda_stream));
}

int fastace_env_step_device_compact(fastace_env_t* env, const fastace_actions_compact_t* actions,
                                    const fastace_step_out_t* out, uint32_t flags, void* cuda_stream) {
    if (!env || !actions || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    return launch_step(env, nullptr, actions, out, flags, static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"

static int shuffle_one_step_16(fastace_env_t* env, uint16_t* perm_person16, uint16_t* perm_firm16, cudaStream_t stream) {
    if (env->dims.num_persons > 65535 || env->dims.num_firms > 65535) { set_error("16-bit orders need P, F <= 65535"); return FASTACE_ERR_INVALID; }
    return fastace_env_shuffle_orders(env, 0u, 0, 1, nullptr, nullptr, perm_person16, perm_firm16, stream);
}

// Host-pointer step.  Pipeline per call (buffer b = call parity):
//   copy_in : [wait kernels that last read buffer b] H2D of every action array -> h2d_done[b]
//   stream  : [wait h2d_done[b], d2h_done[b]] match + update kernels            -> kern_done[b]
//   copy_out: [wait kern_done[b]] D2H of the requested outputs                   -> d2h_done[b]
// so with FASTACE_STEP_ASYNC the copies of neighbouring steps overlap the kernels.
// forward: one shuffle step of the env's own orders into 16-bit device arrays (packed host calls without orders)
static int shuffle_one_step_16(fastace_env_t* env, uint16_t* perm_person16, uint16_t* perm_firm16, cudaStream_t stream);

// mode: 0 = int32 encoding, 1 = compact, 2 = packed (expanded on the device into the compact staging block)
template <typename ActT>
static int step_host_impl(fastace_env_t* env, const ActT* actions, const std::vector<FieldDesc>& fields,
                          ActT* dacts, void** blocks, int mode, const fastace_step_out_t* out, uint32_t flags) {
    const bool compact = mode == 1;
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    if (!env->have_host_streams) {
        FASTACE_CUDA_CHECK(cudaStreamCreateWithFlags(&env->copy_in, cudaStreamNonBlocking));
        FASTACE_CUDA_CHECK(cudaStreamCreateWithFlags(&env->copy_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            FASTACE_CUDA_CHECK(cudaEventCreateWithFlags(&env->h2d_done[b], cudaEventDisableTiming));
            FASTACE_CUDA_CHECK(cudaEventCreateWithFlags(&env->kern_done[b], cudaEventDisableTiming));
            FASTACE_CUDA_CHECK(cudaEventCreateWithFlags(&env->d2h_done[b], cudaEventDisableTiming));
        }
        env->have_host_streams = true;
    }
    const int b = (int)(env->host_steps & 1);
    if (!blocks[b]) {
        int rc = carve(fields, &dacts[b], &blocks[b], false);
        if (rc != FASTACE_OK) return rc;
    }
    if (!env->out_block[b]) {
        int rc = carve(out_fields(env->dims), &env->dout[b], &env->out_block[b], false);
        if (rc != FASTACE_OK) return rc;
    }
    if (mode == 2 && !env->cact_block[b]) {
        int rc = carve(compact_fields(env->dims), &env->dcact[b], &env->cact_block[b], false);
        if (rc != FASTACE_OK) return rc;
    }
    const bool own_orders = mode == 2 && !member(actions, fields[0].offset) && !member(actions, fields[1].offset);
    if (own_orders && !env->ord_started) { set_error("no visiting orders given and the env's own orders were never started (fastace_env_shuffle_orders)"); return FASTACE_ERR_INVALID; }
    for (size_t k = 0; k < fields.size(); k++) {
        const auto& f = fields[k];
        if (own_orders && k < 2) continue;
        if (f.count && !member(actions, f.offset)) { set_error("actions: every array is mandatory"); return FASTACE_ERR_INVALID; }
    }
    if (env->buf_used[b]) FASTACE_CUDA_CHECK(cudaStreamWaitEvent(env->copy_in, env->kern_done[b], 0));
    {
        // Host arrays that sit in one block with the staging buffer's own layout (fields in struct order, each
        // padded to 256 B — what fastace_b200._abi.alloc_host_block hands out) travel as ONE copy instead of 14
        // wherever consecutive arrays touch.
        const char* run_h = nullptr; char* run_d = nullptr; size_t run_n = 0;
        auto flush = [&]() -> cudaError_t {
            cudaError_t e = run_n ? cudaMemcpyAsync(run_d, run_h, run_n, cudaMemcpyHostToDevice, env->copy_in) : cudaSuccess;
            run_n = 0;
            return e;
        };
        for (auto& f : fields) {
            if (!f.count || !member(actions, f.offset)) continue;
            const char* h = static_cast<const char*>(member(actions, f.offset));
            char* d = static_cast<char*>(member(&dacts[b], f.offset));
            const size_t bytes = f.elem * f.count;
            // merged only when the arrays touch with no gap (sizes that are multiples of 256 B, e.g. any E multiple of
            // 256): bytes between two host arrays may belong to somebody else
            if (run_n && (run_n % 256) == 0 && h == run_h + run_n && d == run_d + run_n) {
                run_n += bytes;
            } else {
                FASTACE_CUDA_CHECK(flush());
                run_h = h; run_d = d; run_n = bytes;
            }
        }
        FASTACE_CUDA_CHECK(flush());
    }
    FASTACE_CUDA_CHECK(cudaEventRecord(env->h2d_done[b], env->copy_in));
    FASTACE_CUDA_CHECK(cudaStreamWaitEvent(env->stream, env->h2d_done[b], 0));
    if (env->buf_used[b]) FASTACE_CUDA_CHECK(cudaStreamWaitEvent(env->stream, env->d2h_done[b], 0));
    fastace_step_out_t dout;
    std::memset(&dout, 0, sizeof(dout));
    for (auto& f : out_fields(env->dims))
        if (member(out, f.offset)) member(&dout, f.offset) = member(&env->dout[b], f.offset);
    int rc;
    if (mode == 2) {
        // expand the packed block into the compact staging arrays, then step on those
        const fastace_actions_packed_t& pk = *reinterpret_cast<const fastace_actions_packed_t*>(&dacts[b]);
        fastace_actions_compact_t cz = env->dcact[b];
        const fastace_dims_t& dm = env->dims;
        const PackedLayout PL = packed_layout(dm);
        if (own_orders) {
            rc = shuffle_one_step_16(env, const_cast<uint16_t*>(cz.perm_person), const_cast<uint16_t*>(cz.perm_firm), env->stream);
            if (rc != FASTACE_OK) return rc;
        } else {
            cz.perm_person = pk.perm_person; cz.perm_firm = pk.perm_firm;
        }
        cz.p_consume = pk.p_consume; cz.f_prod = pk.f_prod; cz.f_offer_amt = pk.f_offer_amt; cz.f_offer_price = pk.f_offer_price;
        cz.f_job_labor = pk.f_job_labor; cz.f_job_wage = pk.f_job_wage;
        ExpandParams ej = {dm.num_econ * dm.num_persons, dm.stack_size, PL.bits_job, PL.bytes_job, pk.p_job_idx,
                           const_cast<uint8_t*>(cz.p_job_idx), const_cast<uint16_t*>(cz.p_job_take)};
        ExpandParams eg = {dm.num_econ * dm.num_persons, dm.stack_size, PL.bits_good, PL.bytes_good, pk.p_good_idx,
                           const_cast<uint8_t*>(cz.p_good_idx), const_cast<uint16_t*>(cz.p_good_take)};
        ExpandParams ef = {dm.num_econ * dm.num_firms, dm.stack_size, PL.bits_good, PL.bytes_good, pk.f_good_idx,
                           const_cast<uint8_t*>(cz.f_good_idx), const_cast<uint16_t*>(cz.f_good_take)};
        ExpandParams none = ef; none.agents = 0;
        if (ej.agents > 0) expand_packed_kernel<<<(ej.agents + 255) / 256, 256, 0, env->stream>>>(ej, eg);
        expand_packed_kernel<<<(ef.agents + 255) / 256, 256, 0, env->stream>>>(ef, none);
        FASTACE_CUDA_CHECK(cudaGetLastError());
        env->launches += 2;
        rc = launch_step(env, nullptr, &cz, &dout, flags & ~FASTACE_IDX_MODULO, env->stream);
    } else {
        rc = compact ? launch_step(env, nullptr, reinterpret_cast<const fastace_actions_compact_t*>(&dacts[b]), &dout, flags, env->stream)
                     : launch_step(env, reinterpret_cast<const fastace_actions_t*>(&dacts[b]), nullptr, &dout, flags, env->stream);
    }
    if (rc != FASTACE_OK) return rc;
    FASTACE_CUDA_CHECK(cudaEventRecord(env->kern_done[b], env->stream));
    FASTACE_CUDA_CHECK(cudaStreamWaitEvent(env->copy_out, env->kern_done[b], 0));
    {
        char* run_h = nullptr; const char* run_d = nullptr; size_t run_n = 0;
        auto flush = [&]() -> cudaError_t {
            cudaError_t e = run_n ? cudaMemcpyAsync(run_h, run_d, run_n, cudaMemcpyDeviceToHost, env->copy_out) : cudaSuccess;
            run_n = 0;
            return e;
        };
        for (auto& f : out_fields(env->dims)) {
            char* h = static_cast<char*>(member(out, f.offset));
            if (!h || !f.count) continue;
            const char* d = static_cast<const char*>(member(&env->dout[b], f.offset));
            const size_t bytes = f.elem * f.count;
            if (run_n && (run_n % 256) == 0 && h == run_h + run_n && d == run_d + run_n) {
                run_n += bytes;
            } else {
                FASTACE_CUDA_CHECK(flush());
                run_h = h; run_d = d; run_n = bytes;
            }
        }
        FASTACE_CUDA_CHECK(flush());
    }
    FASTACE_CUDA_CHECK(cudaEventRecord(env->d2h_done[b], env->copy_out));
    env->buf_used[b] = true;
    env->host_steps += 1;
    if (!(flags & FASTACE_STEP_ASYNC)) FASTACE_CUDA_CHECK(cudaStreamSynchronize(env->copy_out));
    return FASTACE_OK;
}

extern "C" {

int fastace_env_step_host(fastace_env_t* env, const fastace_actions_t* actions, const fastace_step_out_t* out,
                          uint32_t flags) {
    if (!env || !actions || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    return step_host_impl(env, actions, action_fields(env->dims), env->dact, env->act_block, 0, out, flags);
}

int fastace_packed_layout(const fastace_dims_t* dims, int* bits_job, int* bytes_job, int* bits_good, int* bytes_good) {
    if (!dims) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    const PackedLayout L = packed_layout(*dims);
    if (bits_job) *bits_job = L.bits_job;
    if (bytes_job) *bytes_job = L.bytes_job;
    if (bits_good) *bits_good = L.bits_good;
    if (bytes_good) *bytes_good = L.bytes_good;
    return FASTACE_OK;
}

int fastace_env_step_host_packed(fastace_env_t* env, const fastace_actions_packed_t* actions,
                                 const fastace_step_out_t* out, uint32_t flags) {
    if (!env || !actions || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    if (env->large_only || (flags & (FASTACE_STEP_LARGE | FASTACE_STEP_SERIAL))) {
        set_error("the packed encoding feeds the warp-per-economy kernels only"); return FASTACE_ERR_INVALID;
    }
    if ((actions->perm_person == nullptr) != (actions->perm_firm == nullptr)) {
        set_error("give both visiting orders or neither"); return FASTACE_ERR_INVALID;
    }
    return step_host_impl(env, actions, packed_fields(env->dims), env->dpact, env->pact_block, 2, out, flags);
}

int fastace_env_step_host_compact(fastace_env_t* env, const fastace_actions_compact_t* actions,
                                  const fastace_step_out_t* out, uint32_t flags) {
    if (!env || !actions || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    return step_host_impl(env, actions, compact_fields(env->dims), env->dcact, env->cact_block, 1, out, flags);
}

int fastace_env_market_stats(const fastace_env_t* env, const fastace_market_stats_t* out, void* cuda_stream) {
    if (!env || !out) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    StatsParams sp;
    sp.E = env->dims.num_econ; sp.F = env->dims.num_firms; sp.G = env->dims.num_goods;
    sp.st = env->dstate; sp.out = *out;
    market_stats_kernel<<<(sp.E + 127) / 128, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(sp);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

int fastace_env_shuffle_orders(fastace_env_t* env, uint32_t seed, int restart, int steps,
                               int32_t* perm_person, int32_t* perm_firm, uint16_t* perm_person16, uint16_t* perm_firm16,
                               void* cuda_stream) {
    if (!env || steps < 0 || (steps == 0 && !restart)) { set_error("bad argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    const int E = env->dims.num_econ, P = env->dims.num_persons, F = env->dims.num_firms;
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    if (!env->ord_rng) {
        FASTACE_CUDA_CHECK(cudaMalloc((void**)&env->ord_rng, sizeof(uint64_t) * (size_t)E));
        FASTACE_CUDA_CHECK(cudaMalloc((void**)&env->ord_person, sizeof(int32_t) * ((size_t)E * P + 1)));
        FASTACE_CUDA_CHECK(cudaMalloc((void**)&env->ord_firm, sizeof(int32_t) * ((size_t)E * F + 1)));
    }
    if (!restart && !env->ord_started) { set_error("the first call must (re)start the orders"); return FASTACE_ERR_INVALID; }
    if (steps == 0) {
        // (re)start only: the next shuffle seeds the engines and starts from the identity order
        env->ord_seed = seed; env->ord_restart_pending = true; env->ord_started = true;
        return FASTACE_OK;
    }
    if (env->ord_restart_pending && !restart) { restart = 1; seed = env->ord_seed; }
    env->ord_restart_pending = false;
    ShuffleParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.E = E; sp.P = P; sp.F = F; sp.seed = seed;
    sp.rng_state = env->ord_rng; sp.state_person = env->ord_person; sp.state_firm = env->ord_firm;
    const size_t smem = (size_t)(P + F) * kShuffleThreads * sizeof(uint16_t);
    int max_optin = 0;
    FASTACE_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, env->device));
    const bool use_smem = P <= 65535 && F <= 65535 && smem <= (size_t)max_optin;
    const unsigned blocks = (unsigned)((E + kShuffleThreads - 1) / kShuffleThreads);
    if (use_smem) {
        if (smem > 48 * 1024)
            FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)shuffle_orders_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sp.use_smem = 1; sp.restart = restart ? 1 : 0; sp.steps = steps;
        sp.out_person = perm_person; sp.out_firm = perm_firm; sp.out_person16 = perm_person16; sp.out_firm16 = perm_firm16;
        shuffle_orders_kernel<<<blocks, kShuffleThreads, smem, stream>>>(sp);
        FASTACE_CUDA_CHECK(cudaGetLastError());
        env->launches += 1;
    } else {
        // economies too large for shared memory: one launch per step, in place in the env's own arrays (32-bit ids only)
        if (perm_person16 || perm_firm16) { set_error("16-bit orders need P, F <= 65535"); return FASTACE_ERR_INVALID; }
        for (int t = 0; t < steps; t++) {
            sp.use_smem = 0; sp.restart = (restart && t == 0) ? 1 : 0; sp.steps = 1;
            shuffle_orders_kernel<<<blocks, kShuffleThreads, 0, stream>>>(sp);
            FASTACE_CUDA_CHECK(cudaGetLastError());
            env->launches += 1;
            if (perm_person) FASTACE_CUDA_CHECK(cudaMemcpyAsync(perm_person + (size_t)t * E * P, env->ord_person, sizeof(int32_t) * (size_t)E * P, cudaMemcpyDeviceToDevice, stream));
            if (perm_firm) FASTACE_CUDA_CHECK(cudaMemcpyAsync(perm_firm + (size_t)t * E * F, env->ord_firm, sizeof(int32_t) * (size_t)E * F, cudaMemcpyDeviceToDevice, stream));
        }
    }
    env->ord_started = true;
    return FASTACE_OK;
}

int fastace_env_sync(fastace_env_t* env) {
    if (!env) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
    if (env->have_host_streams) {
        FASTACE_CUDA_CHECK(cudaStreamSynchronize(env->copy_in));
        FASTACE_CUDA_CHECK(cudaStreamSynchronize(env->stream));
        FASTACE_CUDA_CHECK(cudaStreamSynchronize(env->copy_out));
    }
    return check_device_errors(env);
}

}  // extern "C"

// ---- policy side: fused residual tanh stack (mlp_stack.cuh) ------------------------------------------------
static int mlp_tiles_for(int hidden) {
    const int need = (hidden + 7) / 8;
    const int choices[] = {2, 4, 8, 13, 16};
    for (int c : choices) if (need <= c) return c;
    return 0;
}
template <int NT>
static int launch_mlp_stack(const MlpParams& mp, cudaStream_t stream) {
    using T = MlpTile<NT>;
    static size_t configured[64] = {};
    int dev = 0;
    FASTACE_CUDA_CHECK(cudaGetDevice(&dev));
    size_t smem = T::SMEM + T::LAST_BYTES;
    if (mp.x0) smem += (size_t)T::NP * mp.K0P * 2 + T::B_BYTES;
    if (dev < 64 && configured[dev] < smem) {
        FASTACE_CUDA_CHECK(cudaFuncSetAttribute((const void*)mlp_residual_stack_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    int sms = 0;
    FASTACE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long rows_per_block = (long long)kMlpWarps * 16;
    const long long nblocks = (mp.rows + rows_per_block - 1) / rows_per_block;
    const int grid = (int)std::min<long long>(nblocks, (long long)sms * kMlpBlocksPerSM);   // persistent
    mlp_residual_stack_kernel<NT><<<grid, kMlpThreads, smem, stream>>>(mp);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

static int dispatch_mlp(const MlpParams& mp, int nt, cudaStream_t stream) {
    switch (nt) {
        case 2: return launch_mlp_stack<2>(mp, stream);
        case 4: return launch_mlp_stack<4>(mp, stream);
        case 8: return launch_mlp_stack<8>(mp, stream);
        case 13: return launch_mlp_stack<13>(mp, stream);
        default: return launch_mlp_stack<16>(mp, stream);
    }
}

extern "C" {

int fastace_mlp_stack_layout(int hidden, int* padded_out, int* padded_in) {
    const int nt = mlp_tiles_for(hidden);
    if (hidden < 2 || (hidden & 1) || nt == 0) { set_error("fused stack needs an even hidden size in [2, 128]"); return FASTACE_ERR_INVALID; }
    if (padded_out) *padded_out = nt * 8;
    if (padded_in) *padded_in = ((nt + 1) / 2) * 16 + 8;
    return FASTACE_OK;
}

int fastace_mlp_residual_tanh_stack(const float* x, float* y, int64_t rows, int hidden, int layers,
                                    const uint16_t* w_bf16, const float* bias, void* cuda_stream) {
    if (!x || !y || !w_bf16 || !bias || rows < 0 || layers < 1) { set_error("bad argument"); return FASTACE_ERR_INVALID; }
    const int nt = mlp_tiles_for(hidden);
    if (hidden < 2 || (hidden & 1) || nt == 0) { set_error("fused stack needs an even hidden size in [2, 128]"); return FASTACE_ERR_INVALID; }
    if (rows == 0) return FASTACE_OK;
    MlpParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.x = x; mp.y = y; mp.rows = rows; mp.H = hidden; mp.L = layers; mp.w = w_bf16; mp.bias = bias;
    return dispatch_mlp(mp, nt, static_cast<cudaStream_t>(cuda_stream));
}

int fastace_mlp_forward(const fastace_mlp_desc_t* d, void* cuda_stream) {
    if (!d) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    const int nt = mlp_tiles_for(d->hidden);
    if (d->hidden < 2 || (d->hidden & 1) || nt == 0) { set_error("fused stack needs an even hidden size in [2, 128]"); return FASTACE_ERR_INVALID; }
    if (d->rows < 0 || d->layers < 0 || (d->layers > 0 && (!d->w_bf16 || !d->bias))) { set_error("bad stack description"); return FASTACE_ERR_INVALID; }
    if (!d->x0 && !d->x) { set_error("no input"); return FASTACE_ERR_INVALID; }
    if (d->x0 && (d->in_features < 1 || d->in_features > 128 || !d->w0_bf16 || !d->b0)) { set_error("first layer: 1..128 inputs"); return FASTACE_ERR_INVALID; }
    if (d->out && (d->out_features < 1 || d->out_features > 16 || !d->wl_bf16 || !d->bl || d->activation < 0 || d->activation > 2)) {
        set_error("last layer: 1..16 outputs, activation 0..2"); return FASTACE_ERR_INVALID;
    }
    if (!d->y && !d->out) { set_error("no output"); return FASTACE_ERR_INVALID; }
    if (d->rows == 0) return FASTACE_OK;
    MlpParams mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.x = d->x; mp.y = d->y; mp.rows = d->rows; mp.H = d->hidden; mp.L = d->layers; mp.w = d->w_bf16; mp.bias = d->bias;
    if (d->x0) { mp.x0 = d->x0; mp.K0 = d->in_features; mp.K0P = (d->in_features + 15) / 16 * 16 + 8; mp.w0 = d->w0_bf16; mp.b0 = d->b0; }
    if (d->out) { mp.out = d->out; mp.NOUT = d->out_features; mp.act = d->activation; mp.wl = d->wl_bf16; mp.bl = d->bl; }
    return dispatch_mlp(mp, nt, static_cast<cudaStream_t>(cuda_stream));
}

int fastace_layer_forward(const float* z, const float* bias, const float* x, float* y, float* t, int64_t rows, int hidden,
                          void* cuda_stream) {
    if (!z || !bias || !y || !t || rows < 0 || hidden < 1) { set_error("bad argument"); return FASTACE_ERR_INVALID; }
    const long long n = (long long)rows * hidden;
    if (n == 0) return FASTACE_OK;
    const unsigned blocks = (unsigned)((n / 4 + 1 + 255) / 256);
    layer_forward_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(z, bias, x, y, t, n, hidden);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

int fastace_layer_backward(const float* dy, const float* t, float* dz, int64_t n, void* cuda_stream) {
    if (!dy || !t || !dz || n < 0) { set_error("bad argument"); return FASTACE_ERR_INVALID; }
    if (n == 0) return FASTACE_OK;
    const unsigned blocks = (unsigned)((n / 4 + 1 + 255) / 256);
    layer_backward_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(dy, t, dz, (long long)n);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

int fastace_policy_bernoulli(const float* probas, const float* uniforms, const int64_t* idx, const uint8_t* valid,
                             int num_econ, int agents, int stack, int invalid_nan,
                             int32_t* out_idx, uint8_t* out_take, float* out_logp, void* cuda_stream) {
    if (!probas || !uniforms || !idx || !valid || !out_idx || !out_take || !out_logp || num_econ < 0 || agents < 0 || stack < 0) {
        set_error("bad argument"); return FASTACE_ERR_INVALID;
    }
    const long long n = (long long)num_econ * agents;
    if (n == 0) return FASTACE_OK;
    BernoulliParams q = {probas, uniforms, idx, valid, num_econ, agents, stack, invalid_nan, out_idx, out_take, out_logp};
    policy_bernoulli_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(q);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

int fastace_policy_normal(const float* params, int stride, int offset, const float* noise,
                          int num_econ, int agents, int components, int kind, int accumulate,
                          float* out_x, float* out_logp, void* cuda_stream) {
    if (!params || !noise || !out_x || !out_logp || num_econ < 0 || agents < 0 || components < 1 || stride < 2 || offset < 0 ||
        offset + 2 > stride || kind < 0 || kind > 1) {
        set_error("bad argument"); return FASTACE_ERR_INVALID;
    }
    const long long n = (long long)num_econ * agents;
    if (n == 0) return FASTACE_OK;
    NormalParams q;
    q.params = params; q.noise = noise; q.stride = stride; q.offset = offset; q.E = num_econ; q.A = agents; q.C = components;
    q.kind = kind; q.accumulate = accumulate; q.out_x = out_x; q.out_logp = out_logp;
    q.log_sqrt2pi_scale = (float)(2.0 / (1.1283791670955126 * 0.70710678118654752));   // neuralConstants.h:10
    policy_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(q);
    FASTACE_CUDA_CHECK(cudaGetLastError());
    return FASTACE_OK;
}

int fastace_env_large_stats(const fastace_env_t* env, uint32_t* person_rounds, uint32_t* firm_rounds) {
    if (!env) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    int flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (env->have_large) {   // the iteration kernels leave their round counts on the device
        FASTACE_CUDA_CHECK(cudaSetDevice(env->device));
        FASTACE_CUDA_CHECK(cudaDeviceSynchronize());
        FASTACE_CUDA_CHECK(cudaMemcpy(flags, env->large_sc.changed, sizeof(flags), cudaMemcpyDeviceToHost));
    }
    if (person_rounds) *person_rounds = (uint32_t)flags[3];
    if (firm_rounds) *firm_rounds = (uint32_t)flags[7];
    return FASTACE_OK;
}

int fastace_env_kernel_times(fastace_env_t* env, double* match_ms, double* update_ms, uint64_t* steps) {
    if (!env || !match_ms || !update_ms || !steps) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *match_ms = env->prof_match_ms; *update_ms = env->prof_update_ms; *steps = env->prof_steps;
    env->prof_match_ms = env->prof_update_ms = 0.0; env->prof_steps = 0;
    return FASTACE_OK;
}

int fastace_env_launch_count(const fastace_env_t* env, uint64_t* out_count) {
    if (!env || !out_count) { set_error("null argument"); return FASTACE_ERR_INVALID; }
    *out_count = env->launches;
    return FASTACE_OK;
}

#ifdef FASTACE_CTA_TIMING
// profiling build only: (start ns, end ns, SM) of every economy's warp in the last match_kernel launch
int fastace_debug_cta_times(unsigned long long* out, int economies) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, fastace::g_cta_times, sizeof(unsigned long long) * 3 * (size_t)economies) == cudaSuccess ? FASTACE_OK : FASTACE_ERR_CUDA;
}
#endif

}  // extern "C"
