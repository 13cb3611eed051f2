/*
 * peuqse_b200.h -- C ABI of the B200-native PEUQSE posterior-evaluation engine.
 *
 * Drop-in boundary for ONE hot path of the reference (AdityaSavara/PEUQSE): the fp64 log-posterior
 * evaluation behind PEUQSE.parameter_estimation (PEUQSE/InverseProblem.py, "IP.py" below) and its three
 * data-parallel drivers.  Each entry point names the reference code it replaces.  The reference is pure
 * Python, so the "FFI" a maintainer adds is a ctypes binding; INTEGRATION.md shows it.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (e.g. a torch tensor's data_ptr());
 *     every pointer inside peuqse_problem_desc is HOST memory, copied at create time;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is stream-ordered
 *     and asynchronous unless stated; no entry point synchronises the device except peuqse_problem_create;
 *   - return value: 0 = PEUQSE_OK, otherwise a PEUQSE_ERR_* code; peuqse_last_error() gives the text;
 *   - numerical failure is a VALUE, never a trap: out-of-bounds parameters give logPrior = -inf
 *     (IP.py:2760-2762), NaN simulated responses give logLikelihood = -inf (IP.py:2880-2882);
 *   - one handle per device; calls on one handle must not run concurrently from several host threads.
 */
#ifndef PEUQSE_B200_H
#define PEUQSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEUQSE_MAX_K 16          /* parameters */
#define PEUQSE_MAX_RESP 8        /* response dimensions */

enum {
    PEUQSE_OK = 0,
    PEUQSE_ERR_INVALID = 1,      /* bad argument / unsupported combination */
    PEUQSE_ERR_CUDA = 2,         /* CUDA runtime error (text in peuqse_last_error) */
    PEUQSE_ERR_NOMEM = 3
};

/* built-in simulation models ("simulateByInputParametersOnlyFunction", UserInput.py:38) */
enum {
    PEUQSE_MODEL_EXTERNAL = 0,   /* simulated responses are supplied by the caller (torch callable path) */
    PEUQSE_MODEL_QUADRATIC = 1,  /* Examples/Example00/simulation_model_00.py:6-10: (x_j - th0)^2 + th1; const = x[n] */
    PEUQSE_MODEL_TPD2SITE = 2,   /* rate law of Examples/Example01/tprmodel.py:20, prescribed coverages;
                                    const = kBT[n], cov1[n], cov2[n]; theta = Ea1,Ea2,logA1,logA2,g1,g2 */
    PEUQSE_MODEL_LINEAR = 3,     /* sim_j = sum_d Phi[j][d]*th_d (d ascending); const = Phi[n][k] row-major */
    PEUQSE_MODEL_LANGMUIR_LOG = 4 /* Examples/Example13doeFunctionExample/simulation_functions_Langmuir_CO_H2O_four_parameters.py:23-52:
                                    log10(K pA/(1 + K pA)), K = exp(-(th0 - T_j th1)/(kB T_j)); const = T[n], kB*T[n], pA[n] */
};

/* likelihood evaluation mode of one response dimension (IP.py:2949-2979) */
enum {
    PEUQSE_RESP_MVN_DIAG = 0,    /* one logpdf, variances only (IP.py:2960, 1-D cov)            */
    PEUQSE_RESP_PER_VALUE = 1,   /* n > responses_observed_max_covmat_size: IP.py:2969-2978      */
    PEUQSE_RESP_MVN_FULL = 2     /* one logpdf with a full covariance matrix (IP.py:2960, 2-D)   */
};

/* prior mode (IP.py:2756-2794) */
enum {
    PEUQSE_PRIOR_NONE = 0,       /* every parameter uniform: logPrior = 0 inside the bounds (IP.py:2782-2783) */
    PEUQSE_PRIOR_DIAG = 1,       /* whitening is a (signed) permutation: diagonal scaled covariance */
    PEUQSE_PRIOR_DENSE = 2       /* covariance-matrix prior: dense whitening U (kp x kp)            */
};

/*
 * The constants parameter_estimation.__init__ derives (IP.py:146-350), with scipy's per-call
 * eigendecomposition (_PSD, scipy/stats/_multivariate.py:208-236) hoisted to set-up time by the host.
 * Arrays in "eigen order" are ordered like the columns of scipy's whitening matrix U = u*sqrt(1/s)
 * so that the kernels add the squared whitened deviations in the same order scipy does.
 */
typedef struct peuqse_problem_desc {
    int32_t k;                   /* number of parameters, <= PEUQSE_MAX_K */
    int32_t scaling_on;          /* 0: scaling_uncertainties_type == 'off' (IP.py:2763-2764) */
    int32_t has_lb, has_ub;      /* bounds lists non-empty (IP.py:2797, 2801) */
    const double* lb;            /* [k] lower bounds, -inf where None */
    const double* ub;            /* [k] upper bounds, +inf where None */
    const double* scale;         /* [k] scaling_uncertainties (IP.py:317-328) */
    const double* mu_scaled;     /* [k] mu_prior_scaled (IP.py:333) */
    int32_t prior_mode;          /* PEUQSE_PRIOR_* */
    int32_t kp;                  /* number of Gaussian (non-uniform) parameters */
    const int32_t* prior_gidx;   /* [kp] DIAG: parameter index feeding whitened column j (eigen order);
                                         DENSE: parameter index of Gaussian parameter i */
    const double* prior_w;       /* [kp] DIAG: |U[.,j]| in eigen order */
    const double* prior_U;       /* [kp*kp] DENSE: row-major U, whitened_j = sum_i dev_i*U[i][j] */
    double prior_c;              /* rank*log(2pi) + log_pdet */

    int32_t n_resp;              /* response dimensions, <= PEUQSE_MAX_RESP */
    int32_t n_total;             /* sum of resp_n */
    const int32_t* resp_n;       /* [n_resp] */
    const int32_t* resp_mode;    /* [n_resp] PEUQSE_RESP_* */
    const double* resp_c;        /* [n_resp] MVN modes: rank*log(2pi) + log_pdet */
    const double* y;             /* [n_total] observed values; MVN_DIAG responses in eigen order */
    const double* w;             /* [n_total] MVN_DIAG: |U[.,j]| eigen order; PER_VALUE: sqrt(1/V_j) */
    const double* cval;          /* [n_total] PER_VALUE: log(2pi) + ln V_j (else ignored) */
    const int32_t* src;          /* [n_total] position of this entry in the caller's flat (natural order)
                                              simulated-response vector (EXTERNAL model) */
    const double* const* resp_U; /* [n_resp] MVN_FULL: row-major n_r x n_r whitening, else NULL */

    int32_t model_id;            /* PEUQSE_MODEL_* */
    int32_t n_model_const;       /* doubles in model_const */
    const double* model_const;   /* per-observation constants, rows ordered like y (see enum) */
} peuqse_problem_desc;

typedef struct peuqse_problem peuqse_problem;   /* opaque; owns immutable device copies of the constants */

const char* peuqse_last_error(void);
int peuqse_version(void);
/* device properties the host layer sizes grids with: out[0]=SM count, out[1]=max smem/block (opt-in), out[2]=cc major*10+minor */
int peuqse_device_info(int32_t* out3);

/* fp64 pipe microbenchmark: *instr_per_s = device-wide fp64 instruction rate; mode 0 = DFMA chains, 1 = DMUL+DADD pairs;
 * modes 3 / 4: mma.sync.m8n8k4.f64 (DMMA) from registers, 8 / 4 accumulators per warp -- the result is then FLOP/s.
 * modes 5-8: DMMA in the contraction kernel's 4 x 4 fragment pattern at 8 / 16 / 32 / 64 warps per SM; modes 10-14: 1 ... 16 dependent
 * DFMA chains per thread, one warp per scheduler (latency); modes 20-25: 8 DFMA chains with 4 / 8 / 16 integer instructions, 2 / 4
 * shared loads or 5 + 2 of them per 8 DFMAs (issue mix; the result is DFMA instructions/s).
 * d_scratch: one double of device memory.  Synchronises the stream. */
int peuqse_bench_fp64(int32_t mode, double* instr_per_s, double* d_scratch, void* stream);

/* replaces: parameter_estimation.__init__ hand-off of constants (IP.py:38-438) */
int peuqse_problem_create(const peuqse_problem_desc* desc, peuqse_problem** out);
int peuqse_problem_destroy(peuqse_problem* p);

/*
 * replaces: getLogP / getLogPrior / getLogLikelihood called in a Python loop (IP.py:1647-1654, 2756-2794,
 * 2850-2980) for a built-in model.  d_theta[B][k] row-major.  Outputs (each nullable): d_logp[B],
 * d_logprior[B], d_loglik[B].  logP short-circuits to -inf where logPrior == -inf (IP.py:1649);
 * d_loglik is evaluated for every row regardless (what the MH loop does, IP.py:2636-2639).
 */
int peuqse_logp_batch(const peuqse_problem* p, const double* d_theta, int64_t B,
                      double* d_logp, double* d_logprior, double* d_loglik, void* stream);

/*
 * replaces: the same functions when the simulator is a user (torch) callable: the caller simulated
 * d_sim[B][n_total] (responses concatenated in natural order) and the kernel does prior + likelihood.
 * d_theta may be NULL when only d_loglik is wanted.  Full-covariance responses run the DMMA contraction.
 */
int peuqse_logp_from_sim(const peuqse_problem* p, const double* d_theta, const double* d_sim, int64_t B,
                         double* d_logp, double* d_logprior, double* d_loglik, void* stream);

/*
 * replaces: getLogLikelihood_byResponses (IP.py:2899-2980) when model['responses_simulation_uncertainties'] is a FUNCTION of the
 * parameters (IP.py:563-572, 2883-2896): the caller evaluated d_sim[B][n_total] and the simulation's standard deviations
 * d_sim_sd[B][n_total]; their squares are added to the observed variances d_var_obs[n_total] point by point (IP.py:2925-2927), so
 * the density has no hoistable part.  Responses with 1-D (variance) uncertainties only.  resp_n[n_resp]: values per response
 * (host); per_value[r] != 0: the response is longer than responses_observed_max_covmat_size and is a sum of one scalar density per
 * value, -inf replaced by -1E100 (IP.py:2961-2975).  A NaN among a row's simulated values gives -inf (IP.py:2866-2882).  d_loglik[B].
 */
int peuqse_loglik_simvar(const double* d_sim, const double* d_sim_sd, const double* d_y, const double* d_var_obs,
                         const int32_t* resp_n, const int32_t* per_value, int32_t n_resp, int64_t B, double* d_loglik, void* stream);

/*
 * replaces: the simulated side of transform_responses for responses['data_overcategory'] = 'transient_kinetics' (IP.py:609-658,
 * called per evaluation at IP.py:2890): in place on d_sim[B][n_total], response r is replaced passes[r] times (0, 1 or 2) by its
 * explicit-Euler integral over the abscissa, y[0] = 0, y[j+1] = y[j] + v[j] * d_dt[j] (littleEulerGivenArray, IP.py:3501-3513);
 * d_dt[n_total]: t[j+1] - t[j] per response (the last entry of a response is not used).  Sequential per row: bit-identical to the
 * reference's loop.  The observed side is transformed once on the host (peuqse_b200/problem.py:transform_transient).
 */
int peuqse_transform_transient(double* d_sim, const double* d_dt, const int32_t* resp_n, const int32_t* passes, int32_t n_resp,
                               int64_t B, void* stream);

/* replaces: getLogPrior alone (IP.py:2756-2794, bounds 2796-2805) -- for callers whose likelihood is their own
 * (UserInput.model['custom_logLikelihood'], IP.py:2854-2857).  d_logprior[B]. */
int peuqse_logprior_batch(const peuqse_problem* p, const double* d_theta, int64_t B, double* d_logprior, void* stream);

/* simulate only (getSimulatedResponses, IP.py:2822-2848) for a built-in model: d_sim[B][n_total], natural order */
int peuqse_simulate_batch(const peuqse_problem* p, const double* d_theta, int64_t B, double* d_sim, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multistart sweeps: doMultiStart / doListOfPermutationsSearch (IP.py:1077-1117, 772-1074) over
 * CombinationGeneratorModule.combinationGenerator's grid (CombinationGeneratorModule.py:40-81) or
 * generateInitialPoints('uniform') (IP.py:706, 749).
 * ------------------------------------------------------------------------------------------- */
typedef struct peuqse_sweep_spec {
    int32_t kind;                /* 0 = grid, 1 = uniform (Philox) */
    int32_t k;
    double center[PEUQSE_MAX_K];
    /* grid: point = center -/+ (interval*i) per axis, list order of combinationGenerator */
    double interval[PEUQSE_MAX_K];
    int64_t radii[PEUQSE_MAX_K];
    /* uniform: theta = spread*(-2+4u)*std + center, u = Philox(seed, point index) */
    double spread;
    double std_prior[PEUQSE_MAX_K];
    uint64_t seed;
} peuqse_sweep_spec;

/* per-launch result of pass 1 (lives in device memory, 16 bytes) */
typedef struct peuqse_best {
    double logp;                 /* highest logP in [first, first+count), -inf if none is > -inf */
    int64_t index;               /* lowest global index attaining it (strict '>' in list order, IP.py:933), -1 if none */
} peuqse_best;

/*
 * Pass 1: evaluate points first..first+count-1 of the sweep, keep the first strict maximum.
 * d_logp (nullable): logP of every point, d_logp[i] for global index first+i (the first column of
 * permutations_MAP_logP_and_parameters_values, IP.py:947).  d_best: one peuqse_best.
 * d_work: scratch of peuqse_sweep_work_bytes(p, spec, first, count) bytes, ZERO-INITIALISED before its first use
 * (calls leave it ready for the next call) and shared by passes 1-3 of the same range: pass 1 leaves per-tile
 * maxima in it that let passes 2-3 skip tiles without retained points.  (p = spec = NULL gives the base size,
 * enough for any range when logP is not stored.)
 */
int64_t peuqse_sweep_work_bytes(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count);
int peuqse_sweep_pass1(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                       double* d_logp, peuqse_best* d_best, void* d_work, void* stream);

/* theta[k] of the point d_best->index (NaN when index < 0), decoded on the device */
int peuqse_sweep_best_point(const peuqse_sweep_spec* spec, const peuqse_best* d_best, double* d_theta, void* stream);

/* decode points of a sweep without evaluating them: d_theta[count][k] (listOfPermutations, IP.py:776) */
int peuqse_sweep_points(const peuqse_sweep_spec* spec, int64_t first, int64_t count, double* d_theta, void* stream);

/*
 * Pass 2: convertPermutationsToSamples + calculatePostBurnInStatistics as weighted reductions
 * (IP.py:3463-3491, 1923-1931).  threshold = d_best->logp + log_rel_threshold, with
 * log_rel_threshold = log(10^-coef) (IP.py:1016, 3477); weight_g = (int)exp(logP_g - threshold) for
 * logP_g > threshold.  d_best: DEVICE struct written by pass 1 (or merged over ranks by the caller), so a
 * whole multistart needs no host round trip between passes.  Moments are taken about the MAP point
 * theta(d_best->index) (no cancellation in the variance).  d_logp: pass-1 logP (nullable: recompute).
 * d_acc: PEUQSE_ACC_DOUBLES(k) doubles, layout below; sums are ACCUMULATED into d_acc (zero it first; the
 * min/max slots start at +/-inf).  d_keep_index/d_keep_logp (nullable, capacity keep_capacity): compacted
 * list of retained points, unordered; d_keep_count: one int64 (running count; rows beyond the capacity
 * are dropped).  Does nothing when d_best->index < 0.
 */
#define PEUQSE_ACC_W 0                                  /* sum w                                   */
#define PEUQSE_ACC_WEXP 1                               /* sum w*exp(logP)                         */
#define PEUQSE_ACC_NKEPT 2                              /* number of retained points               */
#define PEUQSE_ACC_S1 3                                 /* [k]   sum w*(th-shift)                  */
#define PEUQSE_ACC_S2(k) (3 + (k))                      /* [k*k] sum w*(th-shift)(th-shift)^T      */
#define PEUQSE_ACC_MIN(k) (3 + (k) + (k) * (k))         /* [k]   min retained theta                */
#define PEUQSE_ACC_MAX(k) (3 + 2 * (k) + (k) * (k))     /* [k]   max retained theta                */
#define PEUQSE_ACC_DOUBLES(k) (3 + 3 * (k) + (k) * (k))
int peuqse_sweep_pass2(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                       const double* d_logp, const peuqse_best* d_best, double log_rel_threshold, double* d_acc,
                       int64_t* d_keep_index, double* d_keep_logp, int64_t keep_capacity, int64_t* d_keep_count,
                       void* d_work, void* stream);

/* bin edges np.linspace(min_d, max_d, nbins+1) from the pass-2 accumulator: d_edges[k][nbins+1] */
int peuqse_sweep_edges(const double* d_acc, int32_t k, int32_t nbins, double* d_edges, uint64_t* d_hist_to_zero /*nullable*/,
                       void* stream);

/*
 * Pass 3: weighted marginal histograms of the retained points (plotting_functions.py:135 -> ax.hist ->
 * np.histogram with equal bins over [min, max] of the retained values, last bin closed).
 * d_edges[k][nbins+1]: bin edges; d_hist[k][nbins]: integer weight sums, accumulated (zero it first).
 * reuse_pass2_sums != 0: peuqse_sweep_pass2 with the same (p, spec, first, count, d_logp, d_best, threshold, d_work) was
 * the previous sweep call on d_work; the per-row / per-column weight sums it left there are binned instead of
 * reading logP again.  d_work may be NULL (then logP is always re-read).
 */
int peuqse_sweep_hist(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                      const double* d_logp, const peuqse_best* d_best, double log_rel_threshold,
                      const double* d_edges, int32_t nbins, uint64_t* d_hist, void* d_work, int32_t reuse_pass2_sums,
                      void* stream);

/*
 * Multi-GPU communicator: NCCL calls issued by the library on the caller's stream (one rank per GPU, one process per
 * rank).  Rank 0 obtains a 128-byte id with peuqse_comm_unique_id and distributes it by any means (e.g. a
 * torch.distributed broadcast); every rank then calls peuqse_comm_create with its CUDA device current.
 */
typedef struct peuqse_comm peuqse_comm;
int peuqse_comm_unique_id(void* out128);
int peuqse_comm_create(const void* id128, int32_t rank, int32_t world, peuqse_comm** out);
int peuqse_comm_destroy(peuqse_comm* c);
int peuqse_comm_allgather(peuqse_comm* c, const void* d_send, void* d_recv /*[world][bytes]*/, int64_t bytes, void* stream);
int peuqse_comm_allreduce_sum_i64(peuqse_comm* c, void* d_buf, int64_t count, void* stream);

/*
 * Peer-memory exchange (ranks on one NVLink/NVSwitch node, world <= 16): the sweep's three tiny collectives as ONE
 * kernel each -- push this rank's payload into every peer's mailbox with remote stores, raise a sequence flag, wait for
 * the peers' flags, merge -- instead of an NCCL collective plus a merge kernel.  replaces: the pickle-file gather +
 * "keep only the best" / concatenate of consolidate_parallel_sampling_data (IP.py:1209-1319).
 * Setup: every rank calls peuqse_peer_alloc (a zeroed mailbox in its device memory + its 64-byte CUDA IPC handle),
 * the handles are all-gathered by any means, then peuqse_peer_create maps the peers' mailboxes.  Every rank must issue
 * the same exchanges in the same order.  The exchange counter lives in device memory: the calls can be captured in a
 * CUDA graph and replayed.  A peer that does not arrive within PEUQSE_PEER_TIMEOUT_S seconds (environment, default 60)
 * poisons the result (NaN / index -1) instead of hanging the device and raises a sticky status that peuqse_peer_status
 * reports (read it after synchronising the stream; the handle must not be used after a time-out).
 */
typedef struct peuqse_peer peuqse_peer;
int peuqse_peer_alloc(void** d_local, void* handle64);
int peuqse_peer_create(void* d_local, const void* handles /*[world][64]*/, int32_t rank, int32_t world, peuqse_peer** out);
int peuqse_peer_destroy(peuqse_peer* p);    /* also frees the local mailbox */
int64_t peuqse_peer_max_bytes(void);        /* largest payload of one exchange */
int peuqse_peer_status(const peuqse_peer* p); /* 0 = fine, 1 = an exchange timed out */
int peuqse_peer_merge_best(peuqse_peer* p, peuqse_best* d_best /*in: local, out: merged*/, void* stream);
int peuqse_peer_merge_acc(peuqse_peer* p, double* d_acc /*PEUQSE_ACC_DOUBLES(k), in/out*/, int32_t k, void* stream);
int peuqse_peer_sum_i64(peuqse_peer* p, void* d_buf /*count int64, in/out*/, int64_t count, void* stream);

/*
 * Multi-GPU merges (replace consolidate_parallel_sampling_data's "keep the best" / stacking, IP.py:1209-1319):
 * d_all holds the all-gathered per-rank records; the merged record is what a single process would have produced
 * (highest logP, lowest index among equals; sums added in rank order, min/max slots reduced accordingly).
 */
int peuqse_merge_best(const peuqse_best* d_all /*[n]*/, int32_t n, peuqse_best* d_out, void* stream);
int peuqse_merge_acc(const double* d_all /*[n][PEUQSE_ACC_DOUBLES(k)]*/, int32_t n, int32_t k, double* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Metropolis-Hastings: doMetropolisHastings (IP.py:2563-2754), C independent chains
 * (mcmc_parallel_sampling, IP.py:1282-1319) advanced by one persistent kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct peuqse_mh_spec {
    int64_t n_chains;
    int64_t n_steps;             /* rows produced per chain INCLUDING row 0 = start point (mcmc_length, IP.py:2606-2626) */
    int64_t burn_in;             /* rows [burn_in, n_steps) enter the statistics (IP.py:2718) */
    int64_t first_step;          /* draws for row i use Philox step index first_step + i (continue sampling) */
    int64_t chain_offset;        /* global index of chain 0 of this launch (multi-GPU sharding) */
    uint64_t seed;
    double modulate;             /* mcmc_modulate_accept_probability (0 = off, IP.py:2644-2649) */
    int32_t k;
    int32_t thin;                /* store every thin-th row (1 = all) */
    double step_chol[PEUQSE_MAX_K * PEUQSE_MAX_K]; /* row-major lower Cholesky factor of Q_covmat*step (IP.py:2630) */
    int32_t map_finding;         /* mcmc_mode 'MAP_finding' (IP.py:2631-2635, 2688-2713): the proposal covariance is multiplied by
                                    dyn * m_i, m_i = u + 0.5 a fresh uniform per step, dyn adapted every 100 steps by the sign of the
                                    slope of (proposal log-posterior against m) over the previous hundred proposals (x 1.05 below 10,
                                    x 0.95 above 0.1); 0 = 'unbiased' */
    int32_t reserved;
} peuqse_mh_spec;

/* per-chain statistics block written by the kernel: PEUQSE_MH_STAT_DOUBLES(k) doubles per chain */
#define PEUQSE_MH_STAT_NACC 0                           /* accepted proposals                      */
#define PEUQSE_MH_STAT_MAPLOGP 1                        /* highest logP among rows >= burn_in      */
#define PEUQSE_MH_STAT_MAPROW 2                         /* first row attaining it                  */
#define PEUQSE_MH_STAT_LASTLOGP 3
#define PEUQSE_MH_STAT_SUMEXP 4                         /* sum exp(logP) over rows >= burn_in      */
#define PEUQSE_MH_STAT_MAPTHETA 5                       /* [k]                                     */
#define PEUQSE_MH_STAT_LAST(k) (5 + (k))                /* [k] last state                          */
#define PEUQSE_MH_STAT_S1(k) (5 + 2 * (k))              /* [k] sum (th - th0) over rows >= burn_in */
#define PEUQSE_MH_STAT_S2(k) (5 + 3 * (k))              /* [k*k] sum (th-th0)(th-th0)^T            */
#define PEUQSE_MH_STAT_DYN(k) (5 + 3 * (k) + (k) * (k)) /* final step coefficient of 'MAP_finding' (1 when unbiased) */
#define PEUQSE_MH_STAT_DOUBLES(k) (6 + 3 * (k) + (k) * (k))

/*
 * d_theta0[C][k]: start points.  Outputs, all nullable except d_stats:
 *   d_samples[R][C][k], d_logp[R][C], d_logprior[R][C], d_loglik[R][C] with R = ceil(n_steps/thin) stored rows
 *   (layout = discrete_chains_post_burn_in_samples (nsteps, nchains, k), IP.py:2719);
 *   d_deltas[n_steps-1][C][k], d_us[n_steps-1][C]: the proposal increments and accept uniforms actually
 *   used (for feeding the reference the same draws); d_stats[C][PEUQSE_MH_STAT_DOUBLES(k)].
 *   With map_finding, d_us holds TWO values per step, d_us[n_steps-1][C][2] = (accept uniform, modulation uniform + 0.5).
 */
int peuqse_mh_run(const peuqse_problem* p, const peuqse_mh_spec* spec, const double* d_theta0,
                  double* d_samples, double* d_logp, double* d_logprior, double* d_loglik,
                  double* d_deltas, double* d_us, double* d_stats, void* stream);

/*
 * Lock-step form of the same sampler (csrc/mh_steps.cu) for what cannot run inside the persistent kernel: a simulator outside the
 * library (the reference's simulateByInputParametersOnlyFunction as a torch callable), custom_logPrior / custom_logLikelihood
 * (IP.py:2757-2759, 2854-2857) and responses with a full covariance matrix.  The caller evaluates log-prior and log-likelihood of
 * every chain's proposal between the two kernels of a step:
 *     peuqse_mhs_begin(spec, d_theta0, d_lp0, d_ll0, d_state, d_stats, outputs...);            row 0
 *     for (i = 1; i < n_steps; ++i) {
 *         peuqse_mhs_propose(spec, i, d_state, d_stats, d_prop, d_deltas, d_us);                d_prop[C][k] = proposals of step i
 *         <d_lp[c] = logPrior(d_prop[c]), d_ll[c] = logLikelihood(d_prop[c])>
 *         peuqse_mhs_accept(spec, i, d_prop, d_lp, d_ll, d_state, d_stats, outputs...); }
 *     peuqse_mhs_finish(spec, d_state, d_stats);
 * Draws, arithmetic, outputs and the statistics block are those of peuqse_mh_run (identical trajectories for a built-in model).
 * step < 0 in propose / accept: the step index is a per-chain counter kept in d_state (starts at 1, advanced by accept; calls beyond
 * n_steps - 1 do nothing), so that one captured propose -> evaluate -> accept sequence can be replayed as a CUDA graph.
 * d_state: C * peuqse_mhs_state_doubles(k) doubles of scratch; d_stats[C][PEUQSE_MH_STAT_DOUBLES(k)] carries the running state.
 */
int64_t peuqse_mhs_state_doubles(int32_t k);
int peuqse_mhs_begin(const peuqse_mh_spec* spec, const double* d_theta0, const double* d_logprior0, const double* d_loglik0,
                     double* d_state, double* d_stats, double* d_samples, double* d_logp, double* d_logprior, double* d_loglik,
                     void* stream);
int peuqse_mhs_propose(const peuqse_mh_spec* spec, int64_t step, double* d_state, const double* d_stats, double* d_prop,
                       double* d_deltas, double* d_us, void* stream);
int peuqse_mhs_accept(const peuqse_mh_spec* spec, int64_t step, const double* d_prop, const double* d_logprior_prop,
                      const double* d_loglik_prop, double* d_state, double* d_stats, double* d_samples, double* d_logp,
                      double* d_logprior, double* d_loglik, void* stream);
int peuqse_mhs_finish(const peuqse_mh_spec* spec, const double* d_state, double* d_stats, void* stream);

/* raw generator access (tests): d_out[count][4] uint32 Philox4x32-10 blocks for (seed, stream, idx=first+i, block) */
int peuqse_philox_blocks(uint64_t seed, uint32_t rng_stream, int64_t first, int64_t count, uint32_t block,
                         uint32_t* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Ensemble slice sampling: doEnsembleSliceSampling -> zeus.EnsembleSampler.run_mcmc (IP.py:2362-2547,
 * zeus call sites 2495-2518).  One ensemble iteration = two half-ensemble sweeps, each a sequence of
 * batched stepping-out / shrinking rounds whose log-posteriors are evaluated by one kernel launch.
 * ------------------------------------------------------------------------------------------- */
typedef struct peuqse_ess_spec {
    int64_t n_walkers;           /* even */
    int64_t n_iters;
    int64_t first_iter;          /* Philox iteration offset (continue sampling) */
    uint64_t seed;
    double mu0;                  /* initial scale factor (zeus default 1.0) */
    int32_t tune;                /* adapt mu during the run */
    int32_t max_rounds;          /* stepping-out / shrink rounds bound (mcmc_maxiter) */
    int32_t k;
    int32_t driver;              /* 0 = choose (per-walker kernels for diagonal responses, rounds for full covariance),
                                    1 = per-walker kernels (csrc/ess.cu), 2 = round-synchronous driver (csrc/ess_rounds.cu);
                                    both consume the same Philox draws, so the walkers do not depend on the choice */
} peuqse_ess_spec;

/*
 * d_x[W][k], d_logp_x[W]: in = start positions and their logP, out = final positions.
 * d_chain[n_iters][W][k], d_chain_logp[n_iters][W] (nullable).  d_info[5]: neval, final mu, n_expand, n_contract, error.
 * d_work: peuqse_ess_work_bytes(W,k) (per-walker kernels only; may be NULL for the round-synchronous driver).
 * Built-in models only.  The round-synchronous driver waits on the host once per round (it needs the size of the next
 * batch) and returns PEUQSE_ERR_INVALID when a stepping-out / shrink loop exceeds max_rounds.
 */
int64_t peuqse_ess_work_bytes(int64_t n_walkers, int32_t k);
int peuqse_ess_run(const peuqse_problem* p, const peuqse_ess_spec* spec, double* d_x, double* d_logp_x,
                   double* d_chain, double* d_chain_logp, double* d_info, void* d_work, void* stream);

/*
 * The round-synchronous driver step by step, for callers that evaluate the log-posterior themselves (torch-callable
 * simulators: simulate the block, then peuqse_logp_from_sim; multi-GPU: share the block's rows out over the ranks).
 * replaces: the per-walker Python calls of zeus.EnsembleSampler.run_mcmc (IP.py:2495-2498).  Per ensemble iteration `it`
 * and half-ensemble `half` in (0, 1):
 *     peuqse_essr_begin_half(h, it, half, d_x, d_logp_x, d_cand, &n, stream);
 *     while (n > 0) { <d_rows[i] = logP(d_cand[i][0..k)), i < n>;  peuqse_essr_round(h, d_rows, d_cand, &n, stream); }
 *     peuqse_essr_end_half(h, d_x, d_logp_x, stream);
 * then peuqse_essr_end_iteration (adapts mu).  d_cand[W][k] receives each round's block of candidate points (at most
 * W rows); begin_half and round wait (spinning on a word of mapped pinned memory that the device writes last, not through
 * a stream synchronisation) until the block is packed and return its row count in *n_rows.  spec->n_iters is
 * not used (the caller owns the loop); iteration numbers are offsets from spec->first_iter.
 */
typedef struct peuqse_essr peuqse_essr;
int peuqse_essr_create(const peuqse_ess_spec* spec, peuqse_essr** out);
int peuqse_essr_destroy(peuqse_essr* h);
int peuqse_essr_begin_half(peuqse_essr* h, int64_t it, int32_t half, const double* d_x, const double* d_logp_x,
                           double* d_cand, int64_t* n_rows, void* stream);
int peuqse_essr_round(peuqse_essr* h, const double* d_logp_rows, double* d_cand, int64_t* n_rows, void* stream);
int peuqse_essr_end_half(peuqse_essr* h, double* d_x, double* d_logp_x, void* stream);
int peuqse_essr_end_iteration(peuqse_essr* h, void* stream);
int peuqse_essr_info(peuqse_essr* h, double* d_info /*[5] as peuqse_ess_run*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Post-burn-in statistics of device-resident samples:
 * calculatePostBurnInStatistics (IP.py:1861-1937) incl. its tail filter (arrayThresholdFilter, IP.py:3635-3664)
 * and calculateInfoGain (IP.py:1770-1847; log-ratio form and the 100-bin KL-divergence form of
 * plt.hist([prior, posterior], bins=100, density=True)), as streaming reductions over
 * d_samples[n][k] (row-major), d_logp[n], d_logprior[n] -- the order the reference stacks the rows in
 * (chain after chain, IP.py:1282-1319).  mu_AP / stdap / evidence agree with numpy to ~1e-13 relative
 * (different summation order); the kept-row set, MAP row and histogram counts are exact.
 * ------------------------------------------------------------------------------------------- */
enum {
    PEUQSE_FILTER_NONE = 0,      /* mcmc_threshold_filter_samples = False */
    PEUQSE_FILTER_AUTO = 1,      /* benchmark 'auto': remove rows with log(-logP) > mean + coef*std (IP.py:1913-1918) */
    PEUQSE_FILTER_MAP = 2,       /* benchmark 'MAP': remove rows with logP < max(logP) - coef (IP.py:1907-1912) */
    PEUQSE_FILTER_MU_AP = 3      /* benchmark 'mu_AP': remove rows with logP < getLogP(mean of all samples) - coef (IP.py:1901-1906) */
};

typedef struct peuqse_stats_spec {
    int32_t k;
    int32_t filter;              /* PEUQSE_FILTER_* */
    double coef;                 /* mcmc_threshold_filter_coefficient ('auto' = 2.0, IP.py:1873-1875) */
    int32_t kl_bins;             /* bins of the KL-divergence histograms (reference: 100); 0 = skip */
    int32_t reserved;
} peuqse_stats_spec;

/* layout of d_out (doubles; index values are int64 bit patterns) */
#define PEUQSE_STATS_N_TOTAL 0
#define PEUQSE_STATS_N_KEPT 1
#define PEUQSE_STATS_CUT 2               /* a row is removed iff logP < cut (-inf: nothing removed) */
#define PEUQSE_STATS_THRESHOLD 3         /* the reference's filteringThreshold (on log(-logP) for 'auto', on logP otherwise) */
#define PEUQSE_STATS_MEAN_LOGNEG 4
#define PEUQSE_STATS_STD_LOGNEG 5
#define PEUQSE_STATS_MAX_LOGP_ALL 6      /* np.max over the unfiltered rows (NaN if any logP is NaN) */
#define PEUQSE_STATS_MAP_LOGP 7
#define PEUQSE_STATS_MAP_INDEX 8         /* int64: row of the first maximum among the KEPT rows (IP.py:1927) */
#define PEUQSE_STATS_MAP_INDEX_ALL 9     /* int64: the same row in the unfiltered numbering */
#define PEUQSE_STATS_EVIDENCE 10         /* mean(exp(logP)) / ||samples||_F (IP.py:1931) */
#define PEUQSE_STATS_INFO_GAIN_LOG_RATIO 11   /* NaN when d_logprior is NULL */
#define PEUQSE_STATS_INFO_GAIN_KL 12     /* NaN when no prior samples are given */
#define PEUQSE_STATS_SUM_EXP 13
#define PEUQSE_STATS_FROB2 14
#define PEUQSE_STATS_MEAN_ALL 16         /* [16] mean of the unfiltered samples (written by begin for PEUQSE_FILTER_MU_AP) */
#define PEUQSE_STATS_MU_AP 32            /* [16] */
#define PEUQSE_STATS_STDAP 48            /* [16] np.std, ddof 0 */
#define PEUQSE_STATS_MAP_THETA 64        /* [16] */
#define PEUQSE_STATS_MIN 80              /* [16] of the kept rows (NaN unless prior samples are given: only the KL range needs them) */
#define PEUQSE_STATS_MAX 96              /* [16] */
#define PEUQSE_STATS_KL_EACH 112         /* [16] info_gain_KL_each_parameter */
#define PEUQSE_STATS_JOINT_LO 128        /* [16] common histogram range of prior and posterior samples */
#define PEUQSE_STATS_JOINT_HI 144        /* [16] */
#define PEUQSE_STATS_DOUBLES 160

int64_t peuqse_stats_work_bytes(void);
/* unfiltered pass: fills N_TOTAL, MEAN_LOGNEG, STD_LOGNEG ('auto'), MAX_LOGP_ALL, MEAN_ALL ('mu_AP': the caller then
 * evaluates getLogP at MEAN_ALL, e.g. with peuqse_logp_batch, and hands the device scalar to finish) */
int peuqse_stats_begin(const peuqse_stats_spec* spec, const double* d_samples, const double* d_logp, int64_t n,
                       double* d_out /*PEUQSE_STATS_DOUBLES*/, void* d_work, void* stream);
/* d_logprior nullable (no log-ratio info gain); d_mu_ap_logp: device scalar, PEUQSE_FILTER_MU_AP only;
 * d_prior_samples[n_prior][k] nullable (samples_of_prior, IP.py:194; no KL info gain without it) */
int peuqse_stats_finish(const peuqse_stats_spec* spec, const double* d_samples, const double* d_logp, const double* d_logprior,
                        int64_t n, const double* d_mu_ap_logp, const double* d_prior_samples, int64_t n_prior,
                        double* d_out, void* d_work, void* stream);
/* after finish: the kept rows copied in order into d_out_*[N_KEPT] (what the reference leaves in post_burn_in_samples,
 * post_burn_in_log_posteriors_un_normed_vec, post_burn_in_log_priors_vec, IP.py:1919-1921) */
int peuqse_stats_compact(const peuqse_stats_spec* spec, const double* d_samples, const double* d_logp, const double* d_logprior,
                         int64_t n, const double* d_out, double* d_out_samples, double* d_out_logp,
                         double* d_out_logprior /*nullable*/, void* d_work, void* stream);

/*
 * Design-of-experiments scenario batches.  replaces: doeGetInfoGainMatrix / doeParameterModulationPermutationsScanner
 * (IP.py:1363-1572), which run one whole doMetropolisHastings per (parameter modulation, condition) pair on synthetic
 * observations (populateResponsesWithSyntheticData, IP.py:1326-1354).  A scenario = one row of model constants (the
 * independent variables the reference's populateIndependentVariablesFunction plugin pushes into the simulator,
 * UserInput.py model['populateIndependentVariablesFunction'], IP.py:1403) + one row of observed values; prior, observed
 * covariance and every other constant are the problem's (the reference keeps the first __init__'s covariance, IP.py:342-346).
 */
typedef struct peuqse_scenarios peuqse_scenarios;   /* opaque; owns d_model_const[S][..] and d_y[S][n_total] */

/* h_model_const[S][n_model_const]: HOST rows laid out like peuqse_problem_desc.model_const (rows ordered like y) */
int peuqse_scenarios_create(const peuqse_problem* p, int64_t n_scenarios, const double* h_model_const, peuqse_scenarios** out);
int peuqse_scenarios_destroy(peuqse_scenarios* sc);
/* observed values of scenario s := simulation at d_theta_true[s*theta_stride ..+k] under scenario s's constants
 * (theta_stride 0 = one "true" parameter vector for every scenario); d_y_natural[S][n_total] (nullable) receives them in
 * the caller's natural order (what the reference leaves in UserInput.responses['responses_observed'], IP.py:1351) */
int peuqse_scenarios_synthesize(const peuqse_problem* p, peuqse_scenarios* sc, const double* d_theta_true, int64_t theta_stride,
                                double* d_y_natural, void* stream);
/* explicit observed values, d_y_natural[S][n_total] in natural order */
int peuqse_scenarios_set_observed(const peuqse_problem* p, peuqse_scenarios* sc, const double* d_y_natural, void* stream);
/* getLogP / getLogPrior / getLogLikelihood of d_theta[S][B][k] under each scenario; outputs [S][B], each nullable */
int peuqse_scenarios_logp(const peuqse_problem* p, const peuqse_scenarios* sc, const double* d_theta, int64_t B,
                          double* d_logp, double* d_logprior, double* d_loglik, void* stream);
/* peuqse_mh_run with chain c sampling scenario c / chains_per_scenario; spec->n_chains = S * chains_per_scenario.  Chain j of
 * every scenario draws from Philox stream chain_offset + j: all scenarios see the same draws, as in the reference, which
 * reseeds numpy in every __init__ (IP.py:191-193).  Outputs as peuqse_mh_run. */
int peuqse_mh_run_scenarios(const peuqse_problem* p, const peuqse_scenarios* sc, const peuqse_mh_spec* spec,
                            int64_t chains_per_scenario, const double* d_theta0, double* d_samples, double* d_logp,
                            double* d_logprior, double* d_loglik, double* d_deltas, double* d_us, double* d_stats, void* stream);

/* per-scenario calculatePostBurnInStatistics (IP.py:1861-1937; filter NONE / AUTO / MAP) + log-ratio information gain
 * (IP.py:1793-1803) over the MH outputs d_samples[n_rows][C][k], d_logp / d_logprior[n_rows][C], C = S * chains_per_scenario,
 * rows [first_row, n_rows) of each chain, a scenario's chains stacked one after the other (IP.py:1282-1319).
 * d_out[S][PEUQSE_SCEN_STATS_DOUBLES(k)]; d_logprior nullable (no information gain). */
#define PEUQSE_SCEN_STATS_N_TOTAL 0
#define PEUQSE_SCEN_STATS_N_KEPT 1
#define PEUQSE_SCEN_STATS_THRESHOLD 2
#define PEUQSE_SCEN_STATS_MAP_LOGP 3
#define PEUQSE_SCEN_STATS_MAP_INDEX 4        /* as a double: row of the first maximum among the kept rows */
#define PEUQSE_SCEN_STATS_EVIDENCE 5
#define PEUQSE_SCEN_STATS_INFO_GAIN_LOG_RATIO 6
#define PEUQSE_SCEN_STATS_CUT 7              /* a row was removed iff logP < cut (-inf: nothing removed) */
#define PEUQSE_SCEN_STATS_MU_AP 8            /* [k] */
#define PEUQSE_SCEN_STATS_STDAP(k) (8 + (k))        /* [k] */
#define PEUQSE_SCEN_STATS_MAP_THETA(k) (8 + 2 * (k)) /* [k] */
#define PEUQSE_SCEN_STATS_DOUBLES(k) (8 + 3 * (k))   /* doubles per scenario: the whole read-back of a batch is S of these */
int peuqse_scenarios_stats(int32_t k, int32_t filter, double coef, int64_t n_scenarios, int64_t chains_per_scenario,
                           int64_t n_rows, int64_t first_row, const double* d_samples, const double* d_logp,
                           const double* d_logprior, double* d_out, void* stream);

/* per-scenario KL-divergence information gain (calculateInfoGain, IP.py:1772-1791) after peuqse_scenarios_stats: the kept
 * rows (logP >= d_stats[s][PEUQSE_SCEN_STATS_CUT]) and d_prior_samples[n_prior][k] (samples_of_prior, IP.py:194, shared by all
 * scenarios) are binned on each scenario's joint edges with np.histogram semantics; d_out[S][1 + k] = total, then each parameter */
int peuqse_scenarios_kl(int32_t k, int32_t bins, int64_t n_scenarios, int64_t chains_per_scenario, int64_t n_rows,
                        int64_t first_row, const double* d_samples, const double* d_logp, const double* d_prior_samples,
                        int64_t n_prior, const double* d_stats, double* d_out, void* stream);

/*
 * Geweke convergence diagnostic of many chains at once.  replaces: geweke_diagnostic (IP.py:4012-4062) called per window,
 * chain and parameter by calculateAndPlotConvergenceDiagnostics (IP.py:3960-3975).
 * d_samples[n_rows][n_cols]: discrete_chains_post_burn_in_samples (nsteps, nchains, k) viewed with n_cols = nchains*k.
 * positions[n_positions]: strictly increasing row numbers in [0, n_rows] at which prefix sums are recorded (host and
 * device copies).  d_slices[n_z][2][6]: for z-score i the first and the last slice [a, b) of IP.py:4053-4054, each as the
 * indices INTO positions[] of the rows a, a+1, a+2, b-2, b-1, b.  d_z[n_z][n_cols].  The spectral variance is
 * statsmodels' yule_walker(order 2, method 'adjusted', demeaned): sigma^2 / (1 - sum(beta))^2, NaN unless sigma^2 > 0.
 */
int64_t peuqse_geweke_work_bytes(int64_t n_rows, int64_t n_cols, int32_t n_positions);
int peuqse_geweke(const double* d_samples, int64_t n_rows, int64_t n_cols, const int64_t* h_positions,
                  const int64_t* d_positions, int32_t n_positions, const int32_t* d_slices, int32_t n_z, double* d_z,
                  void* d_work, void* stream);

/*
 * Integrated autocorrelation time, the layout step.  replaces: the y.reshape(-1) of the (walkers, steps) view inside
 * zeus.autocorr.AutoCorrTime(samples[:n]) (method 'mk'; called per growing window at IP.py:3916-3930; zeus is a third-party
 * dependency that the reference does not vendor).  d_samples[>= n_rows][n_chains][k] (row = step; only the first n_rows
 * rows are read), d_mean[k]: mean of each parameter over those rows and chains.  d_series[k][series_stride]:
 * series_d[c * n_rows + t] = samples[t][c][d] - mean[d], zero from n_rows * n_chains to series_stride (the padding of the
 * FFT that follows).
 */
int peuqse_act_concat(const double* d_samples, int64_t n_rows, int64_t n_chains, int32_t k, const double* d_mean,
                      double* d_series, int64_t series_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEUQSE_B200_H */
