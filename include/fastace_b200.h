/*
 * fastace_b200.h — C ABI of the B200-native batched environment for fastACE's
 * Economy::time_step hot path.
 *
 * Drop-in boundary.  The reference reaches this path in two ways:
 *   (1) C++ callers: Economy::time_step()            /root/reference/src/base/economy.cpp:95-139
 *       with behaviour supplied through the plugin interfaces
 *       PersonDecisionMaker                          src/persons/utilMaxer.h:11-28
 *       FirmDecisionMaker                            src/firms/profitMaxer.h:11-24
 *   (2) Python callers: ctypes over `extern "C"`     src/pybindings.h:8-28, py/main.py:10-126
 * The reference's FFI is plain `extern "C"` + ctypes, so this library is bound the same
 * way: plain pointers and sizes, no C++ or torch types in any signature.
 *
 * One `fastace_env_t` holds E independent economies as structure-of-arrays device
 * buffers.  One call to fastace_env_step_* advances every economy by exactly one
 * Economy::time_step().  The seven plugin decisions (utilMaxer.h:17-23,
 * profitMaxer.h:13-16) are supplied for all agents at once as action arrays
 * (fastace_actions_t) — this replaces the per-agent synchronous plugin callbacks; the
 * decode rules of the shipped plugins (src/neural/neuralFirmDecisionMaker.cpp:111-180,
 * neuralPersonDecisionMaker.cpp:93-111) are applied on the device.
 *
 * Array layout convention: the agent (or market slot) index is always the FASTEST
 * dimension inside an economy, the economy index the slowest:  [E][...][agent].
 * All pointers of one struct are either all host or all device pointers, as stated
 * by the function that takes them.
 *
 * Error convention: every function returns 0 (FASTACE_OK) or a negative
 * fastace_status_t; fastace_last_error() returns a thread-local message.  (The
 * reference has no error codes — `bool` "did not act" returns, economy.cpp:97-106 —
 * and agents that are always in sync here, so that condition cannot arise.)
 */
#ifndef FASTACE_B200_H
#define FASTACE_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASTACE_ABI_VERSION 3

typedef enum fastace_status {
    FASTACE_OK = 0,
    FASTACE_ERR_INVALID = -1,   /* bad argument / unsupported dimension */
    FASTACE_ERR_CUDA = -2,      /* CUDA runtime error (message in fastace_last_error) */
    FASTACE_ERR_NO_DEVICE = -3, /* no CUDA device: there is NO CPU fallback */
    FASTACE_ERR_ALLOC = -4,
    FASTACE_ERR_NOT_CONVERGED = -5 /* a matching kernel hit its iteration cap: the env's state is not a valid step
                                      result (cannot happen per the convergence proof; never silently committed).
                                      Returned by fastace_env_sync, fastace_env_get_state and the next step call;
                                      cleared by fastace_env_set_state */
} fastace_status_t;

/* Limits of every path; the warp-per-economy kernels (one economy = one warp, books in shared memory) additionally
 * need F*G <= 254 and P <= 65535, beyond which the env uses the large-economy path (F <= 56000, P <= 2^20). */
#define FASTACE_MAX_GOODS 8
#define FASTACE_MAX_STACK 16

/*
 * Function plugins (VecToScalar::f, src/functions/vecToScalar.h:13-106).  A person's utility is one
 * VecToScalar over (1 - labor, consumption); a firm's production is one VecToScalar per output good
 * over (laborHired, inputs) (SumOfVecToVec of VToVFromVToS, src/functions/vecToVec.h:27-73).
 * The kind is chosen per env; the parameter arrays of fastace_state_t are read as:
 *   CES           tfp, share (normalised), rho              tfp*(sum share_i (x_i+1e-8)^rho)^(1/rho)  vecToScalar.cpp:112-118
 *   COBB_DOUGLAS  tfp, share = elasticities                 tfp*prod x_i^e_i                          :45-47
 *                 (CobbDouglasCRS = elasticities normalised by the caller, :54-56)
 *   STONE_GEARY   tfp, share = elasticities, theta          tfp*prod (x_i-theta_i)^e_i                :67-69
 *   LEONTIEF      share = productivities                    min_i x_i*p_i                             :80-82
 *   LINEAR        share = productivities                    sum_i p_i*x_i                             :30-32
 * (`df`, ProfitFunc and the ifopt solvers are never called from Economy::time_step.)
 */
typedef enum fastace_function_kind {
    FASTACE_FN_CES = 0,
    FASTACE_FN_COBB_DOUGLAS = 1,
    FASTACE_FN_STONE_GEARY = 2,
    FASTACE_FN_LEONTIEF = 3,
    FASTACE_FN_LINEAR = 4
} fastace_function_kind_t;

typedef struct fastace_dims {
    int32_t num_econ;    /* E  independent economies in this env (this rank's shard)      */
    int32_t num_persons; /* P  persons per economy                                         */
    int32_t num_firms;   /* F  firms per economy                                           */
    int32_t num_goods;   /* G  goods (Economy::numGoods, base.h:124)                       */
    int32_t stack_size;  /* S  offers considered per decision (DEFAULT_STACK_SIZE,
                               src/neural/neuralConstants.h:13)                            */
} fastace_dims_t;

/*
 * Economy state.  Mirrors the reference's objects as SoA:
 *   Agent::money / inventory            base.h:168-171
 *   Person::laborSupplied               base.h:214
 *   Firm::laborHired                    base.h:257
 *   UtilMaxer::utilFunc  (CES)          utilMaxer.h:71, vecToScalar.h:79-93
 *   ProfitMaxer::prodFunc (one CES per output good, create_CES_VecToVec)
 *                                       profitMaxer.h:72, vecToVec.cpp:34-53
 *   NeuralFirmDecisionMaker::last_money src/neural/neuralEconomy.h:113
 *   Economy::market / jobMarket         base.h:125-126  (BaseOffer/Offer/JobOffer base.h:21-66)
 *
 * CES parameters are stored as the CES object holds them after construction
 * (vecToScalar.cpp:105-110): share parameters normalised to sum 1, and
 * rho = substitutionParam = 1/(1 - elasticityOfSubstitution).
 *
 * The goods book is kept in MARKET ORDER (the order of Economy::market after the
 * end-of-step flush, economy.cpp:125): entry n is what an action index n refers to.
 * Every Offer posted by the shipped plugin is for one unit (AMOUNT_PER_OFFER = 1.0) of
 * a single good, so `m_good` replaces Offer::quantities; every JobOffer is for
 * LABOR_AMOUNT_PER_OFFER = 0.5 labour (neuralFirmDecisionMaker.cpp:6-7).
 */
typedef struct fastace_state {
    /* persons */
    double*   p_money;      /* [E][P]                                   */
    double*   p_inv;        /* [E][G][P]                                */
    double*   p_labor;      /* [E][P]       laborSupplied               */
    double*   p_util_tfp;   /* [E][P]                                   */
    double*   p_util_share; /* [E][G+1][P]  input 0 = leisure (1-labor) */
    double*   p_util_rho;   /* [E][P]                                   */
    /* firms */
    double*   f_money;      /* [E][F]                                   */
    double*   f_inv;        /* [E][G][F]                                */
    double*   f_labor;      /* [E][F]       laborHired                  */
    double*   f_last_money; /* [E][F]       money at the firm's first decision of the previous step */
    double*   f_prod_tfp;   /* [E][G][F]        per output good         */
    double*   f_prod_share; /* [E][G][G+1][F]   [output][input], input 0 = labour */
    double*   f_prod_rho;   /* [E][G][F]                                */
    /* goods market, market order, capacity F*G entries per economy */
    int32_t*  m_count;      /* [E]                                      */
    int32_t*  m_owner;      /* [E][F*G]     firm id of the offerer      */
    int32_t*  m_good;       /* [E][F*G]                                 */
    uint32_t* m_left;       /* [E][F*G]     BaseOffer::amountLeft       */
    uint32_t* m_taken;      /* [E][F*G]     BaseOffer::amountTaken      */
    double*   m_price;      /* [E][F*G]                                 */
    /* job market, market order, capacity F entries per economy */
    int32_t*  j_count;      /* [E]                                      */
    int32_t*  j_owner;      /* [E][F]                                   */
    uint32_t* j_left;       /* [E][F]                                   */
    uint32_t* j_taken;      /* [E][F]                                   */
    double*   j_wage;       /* [E][F]       wage per job lot (= wage/0.5) */
    /* StoneGeary thresholds (only read for FASTACE_FN_STONE_GEARY; zero after env creation) */
    double*   p_util_theta; /* [E][G+1][P]                              */
    double*   f_prod_theta; /* [E][G][G+1][F]                           */
} fastace_state_t;

/*
 * One step's injected decisions and visiting orders.
 *   perm_*      : the order in which Economy::time_step visits agents after its two
 *                 std::shuffle calls (economy.cpp:110-111): perm[r] = id of the agent
 *                 visited r-th.  Must be a permutation of 0..n-1.
 *   *_idx/_take : the stack of S offer indices an agent drew for this step
 *                 (decisionNetHandler.cpp:327-365) and the Bernoulli "request it"
 *                 outcomes (decisionNetHandler.cpp:368-387, 446-466).  Index n refers to
 *                 entry n of the respective book as it stood when the step began
 *                 (snapshot rule, decisionNetHandler.cpp:236-275, 303-308).
 *   p_consume   : consumption proportions   (get_consumption_proportions, :496-517)
 *   f_prod      : production-input proportions (get_production_proportions, :519-539)
 *   f_offer_amt / f_offer_price : choose_offers  (:568-603)
 *   f_job_labor / f_job_wage    : choose_job_offers (:606-643); the 1e8 wage clip is
 *                 applied on the device.
 * Continuous actions are float32 exactly as the nets emit them and are widened to
 * double on the device (torchToEigen, decisionNetHandler.cpp:22-25).
 */
typedef struct fastace_actions {
    const int32_t* perm_person;   /* [E][P]    */
    const int32_t* perm_firm;     /* [E][F]    */
    const int32_t* p_job_idx;     /* [E][S][P] */
    const uint8_t* p_job_take;    /* [E][S][P] */
    const int32_t* p_good_idx;    /* [E][S][P] */
    const uint8_t* p_good_take;   /* [E][S][P] */
    const float*   p_consume;     /* [E][G][P] */
    const int32_t* f_good_idx;    /* [E][S][F] */
    const uint8_t* f_good_take;   /* [E][S][F] */
    const float*   f_prod;        /* [E][G][F] */
    const float*   f_offer_amt;   /* [E][G][F] */
    const float*   f_offer_price; /* [E][G][F] */
    const float*   f_job_labor;   /* [E][F]    */
    const float*   f_job_wage;    /* [E][F]    */
} fastace_actions_t;

/*
 * Compact encoding of exactly the same decisions (for callers that want less PCIe / HBM
 * traffic: 34 B instead of 116 B per person-step at S = 10, G = 2).  Offer indices are one byte
 * (the books of this kernel never exceed 254 entries), the Bernoulli outcomes of an agent's S
 * slots are one bit mask (bit i = slot i), visiting orders are 16-bit.  With FASTACE_IDX_MODULO
 * the byte is a raw draw mapped to draw % count; otherwise it is the index itself (>= count: no
 * request).  The index arrays are AGENT-MAJOR — the S bytes of one agent are contiguous, so that
 * the lane that owns a person fetches its whole request list with three 32-bit loads; device
 * arrays must be 4-byte aligned and readable up to the next multiple of 4 bytes.
 */
typedef struct fastace_actions_compact {
    const uint16_t* perm_person;   /* [E][P]    */
    const uint16_t* perm_firm;     /* [E][F]    */
    const uint8_t*  p_job_idx;     /* [E][P][S] */
    const uint16_t* p_job_take;    /* [E][P]    */
    const uint8_t*  p_good_idx;    /* [E][P][S] */
    const uint16_t* p_good_take;   /* [E][P]    */
    const float*    p_consume;     /* [E][G][P] */
    const uint8_t*  f_good_idx;    /* [E][F][S] */
    const uint16_t* f_good_take;   /* [E][F]    */
    const float*    f_prod;        /* [E][G][F] */
    const float*    f_offer_amt;   /* [E][G][F] */
    const float*    f_offer_price; /* [E][G][F] */
    const float*    f_job_labor;   /* [E][F]    */
    const float*    f_job_wage;    /* [E][F]    */
} fastace_actions_compact_t;

/*
 * Packed host encoding (fastace_env_step_host_packed only): the fewest bytes a step's decisions need on the host link.
 * An agent's S request slots are S bit fields of `bits` bits each (little-endian bit order inside the agent's
 * byte string), bits = smallest width whose all-ones value exceeds every valid index — jobs: 2^bits - 1 >= F, goods:
 * 2^bits - 1 >= F*G (fastace_packed_layout; 4 and 5 bits at 10 firms x 2 goods: 5 + 7 bytes instead of 20 + 4).  A
 * field holds the offer index itself; the all-ones value (or any value >= the book size) means "no request", so no
 * take masks travel.  perm_person / perm_firm may both be NULL: the env then uses its own device-generated visiting
 * orders (fastace_env_shuffle_orders with steps = 0 (re)starts them; every step advances them by one shuffle).
 * 20 B instead of 34 B per person-step at S = 10, G = 2.  The device expands the block into the compact encoding.
 */
typedef struct fastace_actions_packed {
    const uint16_t* perm_person;   /* [E][P] or NULL */
    const uint16_t* perm_firm;     /* [E][F] or NULL */
    const uint8_t*  p_job_idx;     /* [E][P][bytes_job]  */
    const uint8_t*  p_good_idx;    /* [E][P][bytes_good] */
    const float*    p_consume;     /* [E][G][P] */
    const uint8_t*  f_good_idx;    /* [E][F][bytes_good] */
    const float*    f_prod;        /* [E][G][F] */
    const float*    f_offer_amt;   /* [E][G][F] */
    const float*    f_offer_price; /* [E][G][F] */
    const float*    f_job_labor;   /* [E][F]    */
    const float*    f_job_wage;    /* [E][F]    */
} fastace_actions_packed_t;

/*
 * Per-step outputs.  p_reward and f_profit are mandatory; every other pointer may be
 * NULL (then it is not written).
 *   p_reward : utility of this step's consumption (neuralPersonDecisionMaker.cpp:107-108)
 *   f_profit : money at the firm's first decision of this step minus the same one step
 *              earlier (neuralFirmDecisionMaker.cpp:65-74); 0 on an env's first step,
 *              where the reference records nothing.
 *   *_ok     : 1 where the requested unit was actually transacted.
 *   old_*    : final counters of the PREVIOUS book's entries (the snapshot the step
 *              traded against), taken just before their owner withdrew them
 *              (profitMaxer.cpp:79-81, 93-95).  old_m_left is 0 for an entry that was
 *              already dead when its owner's turn came (flushed, agent.cpp:21).
 */
typedef struct fastace_step_out {
    double*   p_reward;   /* [E][P]    */
    double*   f_profit;   /* [E][F]    */
    uint8_t*  p_job_ok;   /* [E][S][P] optional */
    uint8_t*  p_good_ok;  /* [E][S][P] optional */
    uint8_t*  f_good_ok;  /* [E][S][F] optional */
    uint32_t* old_m_left; /* [E][F*G]  optional */
    uint32_t* old_m_taken;/* [E][F*G]  optional */
    uint32_t* old_j_left; /* [E][F]    optional */
    uint32_t* old_j_taken;/* [E][F]    optional */
} fastace_step_out_t;

/* step flags */
#define FASTACE_IDX_ABSOLUTE 0u /* index n used as is; n<0 or n>=count => no request (never happens with the reference's randint) */
#define FASTACE_IDX_MODULO   1u /* raw non-negative draw r mapped to r % count: stands in for torch::randint(0,count) (decisionNetHandler.cpp:332-334) */

/* Matching algorithm.  Default: lane-parallel fixed-point matching (kernel v2).  With
 * FASTACE_STEP_SERIAL persons are matched by a serial walk (kernel v1), which also keeps the
 * reference's fp64 operation order for FIRM money; outcomes are otherwise identical. */
#define FASTACE_STEP_SERIAL  2u
/* Diagnostic: bracket each kernel of the step with CUDA events on the launching stream and
 * synchronise after the step; totals are read with fastace_env_kernel_times. */
#define FASTACE_STEP_PROFILE 4u
/* Host-pointer calls only: return without waiting for the step.  Actions are double-buffered on
 * the device, so the copy-in of step t+1 overlaps the kernels of step t.  Host arrays passed to
 * an asynchronous call must stay valid, and outputs are complete, only after fastace_env_sync. */
#define FASTACE_STEP_ASYNC   8u

/* Take the large-economy path (csrc/large_economy.cuh: request/firm fixed-point iteration over stably sorted
 * per-firm event lists, books in global memory) even when the economy fits the warp-per-economy kernels.  Envs whose
 * dims exceed those kernels (F*G > 254, P > 65535, or books beyond shared memory — BASELINE config D) always take
 * it.  Every fp64 update is made in the reference's order (no rounding-order caveat); the iteration runs in
 * cooperative kernels, so the call stays asynchronous on its stream. */
#define FASTACE_STEP_LARGE   16u

/* One step in two calls, so that a caller can compute the firms' decisions from the state the firms actually see
 * (after every person has acted: money after sales and hires, inventories after sales, laborHired; the reference's
 * firms decide inside their own turn, economy.cpp:121-123).  FASTACE_STEP_PERSONS runs the person phase only (reads
 * the person members of `actions`, writes p_reward, p_*_ok, old_j_*); FASTACE_STEP_FIRMS, which must follow it, runs
 * the firm phase (reads perm_firm and the f_* members, writes f_profit, f_good_ok, old_m_*) and completes the step
 * (time advances here); give each call only its own output arrays (the others NULL).  The two calls together give
 * bit-identical results to one full call.  Not with FASTACE_STEP_SERIAL or the compact encoding; on the large-economy
 * path for single-economy envs (num_econ == 1). */
#define FASTACE_STEP_PERSONS 32u
#define FASTACE_STEP_FIRMS   64u
/* The person phase itself in two calls: FASTACE_STEP_PERSONS_TRADE runs job search and purchases (reads perm_person,
 * p_job_*, p_good_*; writes p_*_ok, old_j_*; afterwards p_money / p_inv / p_labor are the persons' money, inventory
 * and laborSupplied as they stand when they choose what to consume), FASTACE_STEP_PERSONS_CONSUME consumes (reads
 * p_consume, writes p_reward).  Consumption touches nobody but the person (utilMaxer.cpp:88-92), so deferring it past
 * the other persons' trades changes nothing: TRADE + CONSUME + FIRMS is bit-identical to one full call. */
#define FASTACE_STEP_PERSONS_TRADE   128u
#define FASTACE_STEP_PERSONS_CONSUME 256u

typedef struct fastace_env fastace_env_t;

/* ---- library ---------------------------------------------------------------------- */
int         fastace_abi_version(void);
const char* fastace_last_error(void);
/* number of visible CUDA devices, or a negative status */
int         fastace_device_count(void);

/* ---- env lifetime ------------------------------------------------------------------ */
/* Allocates all state buffers for E economies on CUDA device `device` (zero-filled,
 * empty markets, time 0).  Replaces `new Economy(goods)` + util::create<T> for every
 * agent (economy.cpp:3-27, util.h:41-46).  Fails with FASTACE_ERR_NO_DEVICE when no
 * GPU is present — there is no CPU path. */
int fastace_env_create(const fastace_dims_t* dims, int device, fastace_env_t** out_env);
int fastace_env_destroy(fastace_env_t* env);
int fastace_env_dims(const fastace_env_t* env, fastace_dims_t* out_dims);
/* Economy::get_time (base.h:91) — identical for all economies of an env */
int fastace_env_time(const fastace_env_t* env, uint32_t* out_time);

/* Utility / production function families of this env (default: CES / CES, the only pair the shipped
 * neural plugins support, neuralPersonDecisionMaker.cpp:34, neuralFirmDecisionMaker.cpp:40-45). */
int fastace_env_set_function_kinds(fastace_env_t* env, int util_kind, int prod_kind);

/* ---- state I/O --------------------------------------------------------------------- */
/* Host -> device.  NULL members are left untouched.  Also sets the env time. */
int fastace_env_set_state(fastace_env_t* env, const fastace_state_t* host_state, uint32_t time);
/* Device -> host.  NULL members are skipped.  Mirrors the reference's read API
 * (get_money/get_inventory/get_laborSupplied/get_laborHired/get_market/get_jobMarket,
 * base.h:90-106,147-150,205,245). */
int fastace_env_get_state(const fastace_env_t* env, fastace_state_t* host_state);
/* Zero-copy: device pointers of the env's own buffers (valid until destroy). */
int fastace_env_device_state(const fastace_env_t* env, fastace_state_t* out_device_state);

/* ---- the hot path ------------------------------------------------------------------ */
/* One Economy::time_step() for all E economies.  `actions`/`out` hold DEVICE pointers;
 * the kernel is enqueued on `cuda_stream` (a cudaStream_t, 0 = default stream) and the
 * call returns without synchronising. */
int fastace_env_step_device(fastace_env_t* env, const fastace_actions_t* actions,
                            const fastace_step_out_t* out, uint32_t flags, void* cuda_stream);
/* Same, with HOST pointers: copies the actions to the device, steps, copies the outputs
 * back and synchronises.  This is the end-to-end call a ctypes/C++ caller with host
 * arrays makes. */
int fastace_env_step_host(fastace_env_t* env, const fastace_actions_t* actions,
                          const fastace_step_out_t* out, uint32_t flags);
/* The same two calls for the compact action encoding. */
int fastace_env_step_device_compact(fastace_env_t* env, const fastace_actions_compact_t* actions,
                                    const fastace_step_out_t* out, uint32_t flags, void* cuda_stream);
int fastace_env_step_host_compact(fastace_env_t* env, const fastace_actions_compact_t* actions,
                                  const fastace_step_out_t* out, uint32_t flags);
/* Per-economy market statistics of the CURRENT books (what the reference's `run` prints after every step, print_info,
 * src/pybindings.cpp:20-75), reduced on the device: per good the sum of quantity/price over its offers, the number of
 * offers and the lots on offer; the sum of wage/labor over the job offers, their number and lots.  The sums are
 * accumulated in market order like the reference's loops (average = sum / count).  DEVICE pointers, any may be NULL;
 * enqueued on `cuda_stream`. */
typedef struct fastace_market_stats {
    double*   sum_quantity_per_price; /* [E][G] */
    uint32_t* offers;                 /* [E][G] */
    uint32_t* lots;                   /* [E][G] */
    double*   sum_wage_per_labor;     /* [E]    */
    uint32_t* job_offers;             /* [E]    */
    uint32_t* job_lots;               /* [E]    */
} fastace_market_stats_t;
int fastace_env_market_stats(const fastace_env_t* env, const fastace_market_stats_t* out, void* cuda_stream);

/* Field widths and byte-string lengths of the packed encoding for these dims. */
int fastace_packed_layout(const fastace_dims_t* dims, int* bits_job, int* bytes_job, int* bits_good, int* bytes_good);
/* The host-pointer step for the packed encoding (FASTACE_STEP_ASYNC as for the other host calls). */
int fastace_env_step_host_packed(fastace_env_t* env, const fastace_actions_packed_t* actions,
                                 const fastace_step_out_t* out, uint32_t flags);

/* Waits for every step enqueued by the host-pointer calls of this env. */
int fastace_env_sync(fastace_env_t* env);
/* number of kernel launches issued by this env's step calls so far */
int fastace_env_launch_count(const fastace_env_t* env, uint64_t* out_count);
/* Accumulated device time (milliseconds, CUDA events) of the steps run with FASTACE_STEP_PROFILE:
 * match_kernel, update_kernel, and the number of such steps.  Resets the accumulators. */
int fastace_env_kernel_times(fastace_env_t* env, double* match_ms, double* update_ms, uint64_t* steps);

/* Iteration rounds of the last large-economy step (person phase, firm phase). */
int fastace_env_large_stats(const fastace_env_t* env, uint32_t* person_rounds, uint32_t* firm_rounds);

/* ---- policy side: fused residual tanh stack ------------------------------------------------------------ */
/* x <- x + tanh(x W_l^T + b_l) for l = 0..layers-1 on a DEVICE [rows][hidden] fp32 matrix: the hidden stack of the
 * reference's decision networks (src/neural/decisionNets.cpp:66-70, 135-139, 186-190, 283-290, 368-372, 439-443) in
 * one kernel, activations resident in registers (bf16 tensor-core operands, fp32 accumulate and residual).
 * `w_bf16` is [layers][padded_out][padded_in] bf16 with W_l[out][in] zero-padded, `bias` [layers][padded_out] fp32
 * zero-padded; the padded sizes come from fastace_mlp_stack_layout.  y may alias x.  Rollout-only (no gradient). */
int fastace_mlp_stack_layout(int hidden, int* padded_out, int* padded_in);
int fastace_mlp_residual_tanh_stack(const float* x, float* y, int64_t rows, int hidden, int layers,
                                    const uint16_t* w_bf16, const float* bias, void* cuda_stream);
/* A whole decision-net body in the same kernel: optional first layer  h = tanh(x0 W0^T + b0)  (in_features <= 128),
 * `layers` residual tanh layers (may be 0), optional last layer  out = act(h Wl^T + bl)  (out_features <= 16,
 * activation 0 none / 1 sigmoid / 2 tanh).  Give x0 or x (the [rows][hidden] stack input); give y (the [rows][hidden]
 * stream after the stack) and/or out.  w0_bf16 is [padded_out][16*ceil(in_features/16)+8], wl_bf16 [16][padded_in],
 * b0 [padded_out], bl [16], all zero-padded. */
typedef struct fastace_mlp_desc {
    int64_t rows;
    int32_t hidden, layers;
    const float* x;  float* y;
    const uint16_t* w_bf16;  const float* bias;
    const float* x0;  int32_t in_features;  const uint16_t* w0_bf16;  const float* b0;
    float* out;  int32_t out_features, activation;  const uint16_t* wl_bf16;  const float* bl;
} fastace_mlp_desc_t;
int fastace_mlp_forward(const fastace_mlp_desc_t* desc, void* cuda_stream);

/* Training-time layer halves (fp32, DEVICE pointers): forward  y = [x +] tanh(z + bias), t = tanh(z + bias)  for the
 * [rows][hidden] GEMM output z (x NULL: no residual); backward  dz = dy * (1 - t^2)  over n elements.  One
 * element-wise pass each instead of eager autograd's three per layer and direction (src/neural/decisionNets.cpp). */
int fastace_layer_forward(const float* z, const float* bias, const float* x, float* y, float* t, int64_t rows, int hidden,
                          void* cuda_stream);
int fastace_layer_backward(const float* dy, const float* t, float* dz, int64_t n, void* cuda_stream);

/* Sampling rules of the decision nets for all agents at once (DEVICE pointers, fp32; rollout-only fast path — the
 * reference's per-agent versions: src/neural/decisionNetHandler.cpp:27-46 sample_normal / sample_logitNormal /
 * sample_logNormal, :368-387 the Bernoulli takes).  Inputs are agent-major rows ([E*A][...]), outputs are written in the
 * env's action layout.
 * bernoulli: take = u < p, log pi = sum_s log p or log(1-p); out_take[E][S][A] = take & valid[e], out_idx[E][S][A] =
 *   idx, out_logp[E][A] (an agent of an economy whose book is empty: NaN if invalid_nan else 0.0, :398-403, 476-480).
 * normal: per component c the pair (mu, log sigma) sits at params[(row*C + c)*stride + offset]; x = noise*sigma + mu,
 *   out_x[E][C][A] = sigmoid(x) (kind 0) or exp(x) (kind 1), out_logp[E][A] (+)= sum_c -0.5((x-mu)/sigma)^2 -
 *   log(sigma*sqrt(2 pi)) (no Jacobian term, as in the reference). */
int fastace_policy_bernoulli(const float* probas, const float* uniforms, const int64_t* idx, const uint8_t* valid,
                             int num_econ, int agents, int stack, int invalid_nan,
                             int32_t* out_idx, uint8_t* out_take, float* out_logp, void* cuda_stream);
int fastace_policy_normal(const float* params, int stride, int offset, const float* noise,
                          int num_econ, int agents, int components, int kind, int accumulate,
                          float* out_x, float* out_logp, void* cuda_stream);

/* ---- legacy entry points of libpybindings.so (src/pybindings.h:8-28) ------------------ */
/* Byte-identical layouts of neural::CustomScenarioParams (344 B) and
 * neural::TrainingParams (136 B), src/neural/neuralScenarios.h:49-186, py/main.py:12-85. */
typedef struct fastace_custom_scenario_params {
    uint32_t numPeople, numFirms;
    double money_mu, money_sigma, good1_mu, good1_sigma, good2_mu, good2_sigma;
    double labor_share_mu, labor_share_sigma, good1_share_mu, good1_share_sigma,
           good2_share_mu, good2_share_sigma;
    double discount_mu, discount_sigma, elasticity_mu, elasticity_sigma;
    double firm_money_mu, firm_money_sigma, firm_good1_mu, firm_good1_sigma,
           firm_good2_mu, firm_good2_sigma;
    double firm_tfp1_mu, firm_tfp1_sigma, firm_tfp2_mu, firm_tfp2_sigma;
    double firm_labor_share1_mu, firm_labor_share1_sigma, firm_good1_share1_mu,
           firm_good1_share1_sigma, firm_good2_share1_mu, firm_good2_share1_sigma;
    double firm_labor_share2_mu, firm_labor_share2_sigma, firm_good1_share2_mu,
           firm_good1_share2_sigma, firm_good2_share2_mu, firm_good2_share2_sigma;
    double firm_elasticity1_mu, firm_elasticity1_sigma, firm_elasticity2_mu,
           firm_elasticity2_sigma;
} fastace_custom_scenario_params_t;

typedef struct fastace_training_params {
    uint32_t numEpisodes, episodeLength, updateEveryNEpisodes, checkpointEveryNEpisodes;
    uint32_t stackSize, encodingSize, hiddenSize, nHidden, nHiddenSmall;
    double purchaseNetLR, firmPurchaseNetLR, laborSearchNetLR, consumptionNetLR,
           productionNetLR, offerNetLR, jobOfferNetLR, valueNetLR, firmValueNetLR;
    uint32_t episodeBatchSizeForLRDecay, patienceForLRDecay;
    double multiplierForLRDecay;
    uint32_t reverseAnnealingPeriod;
} fastace_training_params_t;

/* same names and by-value struct returns as src/pybindings.cpp:8-18 */
fastace_custom_scenario_params_t create_scenario_params(unsigned int numPeople, unsigned int numFirms);
fastace_training_params_t        create_training_params(void);

/* `run` and `train` of libpybindings.so, as py/main.py binds and calls them (py/main.py:96-126, 151-155;
 * src/pybindings.h:16-27, src/pybindings.cpp:78-114): structs BY VALUE for run; for train `output` receives
 * trainingParams->numEpisodes episode losses (caller-owned) and the learning rates the schedulers end on are written
 * back into *trainingParams; checkpoints under ../models/ relative to the working directory (neuralConstants.h:32).
 * The episode loop, the eleven decision networks and the actor-critic trainer live in fastace_b200/legacy.py; these
 * entry points call it through the CPython C API of the process they are loaded into (py/main.py's own interpreter),
 * or start an interpreter for a non-Python caller.  Batch size / device of the batched env: environment variables
 * FASTACE_NUM_ECONOMIES (default 64), FASTACE_DEVICE (default 0), FASTACE_SEED (default 0). */
void run(fastace_custom_scenario_params_t scenarioParams, fastace_training_params_t trainingParams);
void train(double* output, const fastace_custom_scenario_params_t* scenarioParams, fastace_training_params_t* trainingParams,
           bool fromPretrained, double perturbationSize);

/* ---- scenario initial state (host side) ---------------------------------------------- */
/* Fills a HOST fastace_state_t for E economies with the initial-state distributions of
 * CustomScenario::setup (src/neural/neuralScenarios.cpp:93-161; G must be 2): economy e
 * draws from std::minstd_rand0(seed + e) through std::normal_distribution, the same
 * libstdc++ generators the reference uses.  Markets start empty.  `p_discount` ([E][P],
 * may be NULL) receives the persons' discount rates, which the step itself never reads. */
int fastace_scenario_custom_init(const fastace_dims_t* dims,
                                 const fastace_custom_scenario_params_t* params,
                                 uint32_t seed, fastace_state_t* host_state, double* p_discount);
/* Fills perm_person/perm_firm ([E][P], [E][F], HOST) for `step_index` = 0,1,2,... exactly
 * as Economy::time_step would visit agents: economy e keeps a std::minstd_rand0(seed+e)
 * and applies std::shuffle to its person vector, then its firm vector, cumulatively
 * (economy.cpp:110-111).  `rng_state` ([E] uint64, caller-owned, zero-initialised before
 * step 0) and the perm arrays carry the stream and the cumulative order between calls. */
int fastace_shuffle_orders(const fastace_dims_t* dims, uint32_t seed, uint64_t* rng_state,
                           int32_t* perm_person, int32_t* perm_firm, int first_call);

/* The same on the DEVICE, for a rollout that never leaves it: advances the env's own engines and cumulative orders by
 * `steps` Economy::time_step shuffles (economy.cpp:110-111, bit-identical to libstdc++'s std::shuffle with
 * std::minstd_rand0 — one thread per economy replays the generator and the pairwise-swap algorithm) and writes step
 * t's orders to slice t of the DEVICE arrays perm_person [steps][E][P] / perm_firm [steps][E][F], in int32 and / or
 * 16-bit form (NULL = not wanted).  restart != 0: economy e seeds minstd_rand0(seed + e) and starts from the identity
 * order, exactly like first_call of fastace_shuffle_orders.  steps = 0 with restart != 0 only (re)starts the stream
 * (the orders the packed host calls then consume one step at a time).  Enqueued on `cuda_stream`; does not synchronise. */
int fastace_env_shuffle_orders(fastace_env_t* env, uint32_t seed, int restart, int steps,
                               int32_t* perm_person, int32_t* perm_firm, uint16_t* perm_person16, uint16_t* perm_firm16,
                               void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* FASTACE_B200_H */
