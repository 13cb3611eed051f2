// Persistent Metropolis-Hastings (kernel K3 of SURVEY section 2): every chain lives in registers for the
// whole run; proposal draws come from Philox keyed on (seed, chain, step), so a chain's trajectory does not
// depend on how chains are mapped to threads or GPUs.
// Replaces doMetropolisHastings' Python loop (IP.py:2626-2675): proposal = theta + N(0, Q*step) (2630),
// log_accept = (ll_p + lp_p) - (ll_c + lp_c) (2641), optional 1/N flattening (2644-2649),
// accept iff log_accept > log(u) (2650-2652).  The reference re-simulates the current point each step
// (2638-2639); re-evaluating an unchanged point reproduces the same doubles, so the cached value is used.
// Three kernels, one trajectory per (seed, chain): k_mh (a thread per chain: small problems, many chains), k_mh_warp (a warp per
// chain: many observations, any built-in model; also the scenario batches of the design of experiments) and k_mh_tpd (the
// rate-law model with one response of many observations -- config C3: up to four chains per warp, exponents through a per-lane
// table; see its header for the issue-port arithmetic that shaped it).
#include <cstdlib>
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"
#include "philox.cuh"
#include "mh_common.cuh"


struct MHOut {
    double* samples;
    double* logp;
    double* lprior;
    double* llik;
    double* deltas;
    double* us;
    double* stats;
};

// per-chain running statistics (rows >= burn_in)
template <int KT>
struct ChainStats {
    double nacc, maplogp, maprow, sumexp;
    double maptheta[KT], s1[KT], s2[KT * (KT + 1) / 2];
    __device__ __forceinline__ void init() {
        nacc = 0.0; maplogp = pq_neg_inf(); maprow = -1.0; sumexp = 0.0;
#pragma unroll
        for (int d = 0; d < KT; ++d) { maptheta[d] = 0.0; s1[d] = 0.0; }
#pragma unroll
        for (int e = 0; e < KT * (KT + 1) / 2; ++e) s2[e] = 0.0;
    }
    // explp = exp(lp), computed by the caller when lp last changed (an unchanged chain state repeats the same value)
    __device__ __forceinline__ void add_row(int k, long long row, const double (&th)[KT], const double (&th0)[KT], double lp,
                                            double explp) {
        if (lp > maplogp) {                                    // first maximum (list.index of max, IP.py:1925-1927)
            maplogp = lp; maprow = (double)row;
#pragma unroll
            for (int d = 0; d < KT; ++d) maptheta[d] = th[d];
        }
        sumexp += explp;
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            const double dd = th[d] - th0[d];
            if (d < k) s1[d] += dd;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                if (d < k && d2 < k) s2[e] = fma(dd, th[d2] - th0[d2], s2[e]);
                ++e;
            }
        }
    }
    __device__ void store(int k, double* out, const double (&th)[KT], double lastlp) const {
        out[PEUQSE_MH_STAT_NACC] = nacc;
        out[PEUQSE_MH_STAT_MAPLOGP] = maplogp;
        out[PEUQSE_MH_STAT_MAPROW] = maprow;
        out[PEUQSE_MH_STAT_LASTLOGP] = lastlp;
        out[PEUQSE_MH_STAT_SUMEXP] = sumexp;
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            if (d < k) {
                out[PEUQSE_MH_STAT_MAPTHETA + d] = maptheta[d];
                out[PEUQSE_MH_STAT_LAST(k) + d] = th[d];
                out[PEUQSE_MH_STAT_S1(k) + d] = s1[d];
            }
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                if (d < k && d2 < k) {
                    out[PEUQSE_MH_STAT_S2(k) + d * k + d2] = s2[e];
                    out[PEUQSE_MH_STAT_S2(k) + d2 * k + d] = s2[e];
                }
                ++e;
            }
        }
    }
};
// ---------------------------------------------------------------------------------------------
// thread shape: one chain per thread (n_total <= PQ_THREAD_SHAPE_MAX_N)
// ---------------------------------------------------------------------------------------------
// SCEN: chain c samples the posterior of scenario c / SC.per (its own observed values and model constants, read from
// global memory) and draws from Philox stream chain_offset + c % SC.per, so every scenario sees the same draws -- what the
// reference's design-of-experiments loop does by reseeding numpy in each __init__ (IP.py:191-193, 1404-1407).
template <int MODEL, int KT, bool SCEN = false>
__global__ void __launch_bounds__(128) k_mh(const __grid_constant__ DevProblem P, const __grid_constant__ DevMH M,
                                            const double* __restrict__ theta0, const MHOut out, int use_smem,
                                            const DevScen SC) {
    extern __shared__ double smem[];
    ObsView O = stage_obs(P, smem, use_smem);
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.n_chains) return;
    const int k = P.k;
    const long long C = M.n_chains;
    unsigned long long gchain = (unsigned long long)(M.chain_offset + c);
    if constexpr (SCEN) {
        const long long s = c / SC.per;
        O.y = SC.y + s * SC.n_total;
        O.mc = SC.mc + s * SC.n_mconst;
        gchain = (unsigned long long)(M.chain_offset + (c - s * SC.per));
    }

    double th[KT], th0[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) { th[d] = (d < k) ? theta0[c * k + d] : 0.0; th0[d] = th[d]; }
    double cur_lp = eval_prior<KT>(P, th);
    double cur_ll = eval_loglik_thread<MODEL, KT>(P, O, th, nullptr);
    double cur_sum = cur_ll + cur_lp;                            // what the accept test uses (IP.py:2641)
    ChainStats<KT> st;
    st.init();
    // row 0: samples[0] = start, logP[0] = getLogP(start); log_priors[0] = log_likelihoods[0] = 0 (IP.py:2610-2625)
    {
        const double lp0 = combine_logp(cur_lp, cur_ll);
        if (out.samples) {
#pragma unroll
            for (int d = 0; d < KT; ++d) if (d < k) out.samples[c * k + d] = th[d];
        }
        if (out.logp) out.logp[c] = lp0;
        if (out.lprior) out.lprior[c] = 0.0;
        if (out.llik) out.llik[c] = 0.0;
        if (M.burn_in <= 0) st.add_row(k, 0, th, th0, lp0, exp(lp0));
    }
    double cur_exp = exp(cur_sum);                               // exp(logP) of the current state, for the evidence sum
    long long next_row = 1;                                      // rows are written every thin-th step
    int until_row = M.thin;
    MapFind mf;
    mf.init();
    for (long long i = 1; i < M.n_steps; ++i) {
        double z[KT], delta[KT], prop[KT], u, um;
        const unsigned long long step = (unsigned long long)(M.first_step + i);
        draws_thread<KT>(M, gchain, step, z, u, um);
        make_delta<KT>(M, z, delta);
        const double mod = um + 0.5;                              // IP.py:2633
        if (M.map_finding) {
            const double sc = mf.scale(mod);
#pragma unroll
            for (int d = 0; d < KT; ++d) delta[d] = sc * delta[d];
        }
#pragma unroll
        for (int d = 0; d < KT; ++d) prop[d] = th[d] + delta[d];  // IP.py:2630
        const double lp_p = eval_prior<KT>(P, prop);
        const double ll_p = eval_loglik_thread<MODEL, KT>(P, O, prop, nullptr);   // evaluated even if lp_p == -inf (IP.py:2637)
        const double sum_p = ll_p + lp_p;
        double la = sum_p - cur_sum;                              // IP.py:2641
        if (M.modulate != 0.0) la = (1.0 / M.modulate) * la;      // IP.py:2649
        const bool acc = la > log(u);                             // IP.py:2651-2652; NaN -> reject
        if (acc) {
#pragma unroll
            for (int d = 0; d < KT; ++d) th[d] = prop[d];
            cur_lp = lp_p; cur_ll = ll_p; cur_sum = sum_p;
            cur_exp = exp(cur_sum);
            st.nacc += 1.0;
        }
        if (out.deltas) {
#pragma unroll
            for (int d = 0; d < KT; ++d) if (d < k) out.deltas[((i - 1) * C + c) * k + d] = delta[d];
        }
        if (out.us) {
            if (M.map_finding) { out.us[((i - 1) * C + c) * 2] = u; out.us[((i - 1) * C + c) * 2 + 1] = mod; }
            else out.us[(i - 1) * C + c] = u;
        }
        if (M.map_finding) mf.after_step(i, mod, sum_p);
        if (--until_row == 0) {
            until_row = M.thin;
            const long long row = next_row++;
            if (out.samples) {
#pragma unroll
                for (int d = 0; d < KT; ++d) if (d < k) out.samples[(row * C + c) * k + d] = th[d];
            }
            if (out.logp) out.logp[row * C + c] = cur_sum;      // IP.py:2661, 2673
            if (out.lprior) out.lprior[row * C + c] = cur_lp;
            if (out.llik) out.llik[row * C + c] = cur_ll;
        }
        if (i >= M.burn_in) st.add_row(k, i, th, th0, cur_sum, cur_exp);
    }
    st.store(k, out.stats + c * PEUQSE_MH_STAT_DOUBLES(k), th, cur_sum);
    out.stats[c * PEUQSE_MH_STAT_DOUBLES(k) + PEUQSE_MH_STAT_DYN(k)] = mf.dyn;
}

// ---------------------------------------------------------------------------------------------
// warp shape: one chain per warp (n_total > PQ_THREAD_SHAPE_MAX_N); lanes stride over the observations.
// What is per chain rather than per observation is spread over the lanes instead of being repeated in all 32:
//   * the running moments: accumulator slot s (the KT(KT+1)/2 second moments in the order ChainStats
//     enumerates them, then the KT first moments) lives in lane s % 32, so the chain state costs ~20
//     registers instead of ~120 and four CTAs of four chains fit an SM;
//   * the draws: one Philox + Box-Muller (+ log u) pass fills the 8 blocks of FOUR consecutive steps
//     (lane = 8*substep + block) and the per-step values are broadcast reads of a 512-byte shared-memory row.
// The arithmetic on every value is that of the thread shape, so both shapes walk the same trajectory.
// ---------------------------------------------------------------------------------------------
#define MHW_WARPS 4
#define MH_FEW_CHAINS 8192   // below this the thread shape cannot fill the fp64 pipe (one warp step ~4000 cycles)
template <int KT>
struct MhwLayout {
    static constexpr int NPAIR = KT * (KT + 1) / 2;
    static constexpr int NSLOT = NPAIR + KT;
    static constexpr int NACC = (NSLOT + 31) / 32;
    static constexpr int KPAD = (KT + 1) & ~1;
    // cur[KPAD] | th0[KPAD] | proposal[KPAD] | draws[32][2] | modulation uniforms[4] | 'MAP_finding' state[6]
    static constexpr int WDOUBLES = 3 * KPAD + 64 + 4 + 6;
};

// the draws of one lane (k_mh_warp: lane = 8 * step offset + block, four steps a pass; k_mh_tpd: 32 / NBP steps a pass): a pair of
// normals (blocks 0..6) or the accept uniform, its logarithm and the modulation uniform of 'MAP_finding' (block 7).  Not inlined: the large-argument paths of sincos and log would otherwise set
// the register budget of the whole kernel although they run once every four steps.
__device__ __noinline__ double2 mhw_draw(unsigned long long seed, unsigned long long gchain, unsigned long long step, int blk, int k,
                                         double* um) {
    const Philox4 r = philox_block(seed, PQ_STREAM_MH, gchain, (uint32_t)(8ULL * step + (unsigned)blk));
    // ONE log for every lane: of 1 - u_a for a pair of normals (philox_normal_pair's radius), of u_a for the accept uniform
    const double ua = philox_u53(r.x, r.y), ub = philox_u53(r.z, r.w);
    const double lg = log(blk < 7 ? 1.0 - ua : ua);
    double a = 0.0, b2 = 0.0;
    if (blk < 7) {
        if (2 * blk < k) {                                         // Box-Muller, the operations of philox_normal_pair
            const double rad = sqrt(-2.0 * lg);
            double sn, cs;
            sincos(6.283185307179586 * ub, &sn, &cs);
            a = rad * cs;
            b2 = rad * sn;
        }
    } else {
        a = ua;                                                    // accept uniform ...
        b2 = lg;                                                   // ... and its logarithm (IP.py:2651)
        *um = ub;                                                  // 'MAP_finding' modulation uniform
    }
    return make_double2(a, b2);
}
__device__ __noinline__ double mhw_exp(double x) { return exp(x); }
template <int KT>
__device__ __noinline__ double mhw_prior(const DevProblem& P, const double (&th)[KT]) { return eval_prior<KT>(P, th); }

// TEVAL: every lane evaluates the whole likelihood in the thread shape's order (few observations, few chains):
// the values are bit-identical to k_mh's, and what the warp buys is latency -- draws for four steps per pass and
// lane-spread moments cut a chain step from ~4000 to a few hundred cycles when there are too few chains to fill
// the machine with one thread each.
template <int MODEL, int KT, int U, int SMEM, bool TEVAL = false, int MINB = 4, bool SCEN = false>
__global__ void __launch_bounds__(32 * MHW_WARPS, MINB) k_mh_warp(const __grid_constant__ DevProblem P,
                                                               const __grid_constant__ DevMH M,
                                                               const double* __restrict__ theta0, const MHOut out,
                                                               int use_smem, const DevScen SC) {
    using L = MhwLayout<KT>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* cur = smem + warp * L::WDOUBLES;
    double* th0s = cur + L::KPAD;
    double* prp = th0s + L::KPAD;
    double2* dr = reinterpret_cast<double2*>(prp + L::KPAD);
    double* dmod = reinterpret_cast<double*>(dr + 32);
    MapFind* mf = reinterpret_cast<MapFind*>(dmod + 4);
    ObsView O;
    O = stage_obs<SMEM>(P, smem + MHW_WARPS * L::WDOUBLES, use_smem);
    // log-likelihood of a parameter vector (every lane ends with the same value)
    auto loglik = [&](const double (&p)[KT]) -> double {
        if constexpr (TEVAL) {
            return eval_loglik_thread<MODEL, KT>(P, O, p, nullptr);
        } else {
            return eval_loglik_warp<MODEL, KT, U>(P, O, p, nullptr, lane);
        }
    };
    const long long c = (long long)blockIdx.x * MHW_WARPS + warp;
    if (c >= M.n_chains) return;
    const int k = P.k;
    const long long C = M.n_chains;
    unsigned long long gchain = (unsigned long long)(M.chain_offset + c);
    if constexpr (SCEN) {
        const long long s = c / SC.per;
        O.y = SC.y + s * SC.n_total;
        O.mc = SC.mc + s * SC.n_mconst;
        gchain = (unsigned long long)(M.chain_offset + (c - s * SC.per));
    }

    // the chain's point lives in shared memory (cur) and so does the proposal (prp): registers are for the observation loop
    if (lane < KT) {
        const double t0 = (lane < k) ? theta0[c * k + lane] : 0.0;
        cur[lane] = t0; th0s[lane] = t0; prp[lane] = t0;
    }
    if (lane == 0) mf->init();
    __syncwarp();
    // accumulator slots of this lane
    int sa[L::NACC], sb[L::NACC];
    double acc[L::NACC], va[L::NACC], vb[L::NACC];
#pragma unroll
    for (int q = 0; q < L::NACC; ++q) {
        const int s = lane + 32 * q;
        int a = -1, b = -1;
        if (s < L::NPAIR) {
            int e = 0;
            for (int d = 0; d < KT; ++d)
                for (int d2 = d; d2 < KT; ++d2) { if (e == s) { a = d; b = d2; } ++e; }
            if (a >= k || b >= k) a = -1;
        } else if (s < L::NSLOT) {
            a = s - L::NPAIR; b = -1;
            if (a >= k) a = -1;
        }
        sa[q] = a; sb[q] = b; acc[q] = 0.0; va[q] = 0.0; vb[q] = (a >= 0 && b < 0) ? 1.0 : 0.0;
    }
    double nacc = 0.0, maplogp = pq_neg_inf(), maprow = -1.0, sumexp = 0.0, mapth = 0.0;

    double cur_lp, cur_ll;
    {
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = cur[d];
        cur_lp = eval_prior<KT>(P, th);
        cur_ll = loglik(th);
    }
    double cur_sum = cur_ll + cur_lp;
    auto add_row = [&](long long row, double lp, double explp) {
        if (lp > maplogp) {                                       // first maximum (IP.py:1925-1927)
            maplogp = lp; maprow = (double)row;
            if (lane < KT) mapth = cur[lane];
        }
        sumexp += explp;
#pragma unroll
        for (int q = 0; q < L::NACC; ++q) acc[q] = fma(va[q], vb[q], acc[q]);   // first moments: fma(dd, 1, s1) == s1 + dd
    };
    {
        const double lp0 = combine_logp(cur_lp, cur_ll);
        if (out.samples && lane < k) out.samples[c * k + lane] = cur[lane];
        if (lane == 0) {
            if (out.logp) out.logp[c] = lp0;
            if (out.lprior) out.lprior[c] = 0.0;
            if (out.llik) out.llik[c] = 0.0;
        }
        if (M.burn_in <= 0) add_row(0, lp0, exp(lp0));
    }
    double cur_exp = exp(cur_sum);
    long long next_row = 1;
    int until_row = M.thin;
    const int sub = lane >> 3, blk = lane & 7;
    for (long long i = 1; i < M.n_steps; ++i) {
        const int g = (int)((i - 1) & 3);
        if (g == 0) {
            const unsigned long long step = (unsigned long long)(M.first_step + i + sub);
            double um = 0.0;
            const double2 d2 = mhw_draw(M.seed, gchain, step, blk, k, &um);
            __syncwarp();
            dr[lane] = d2;
            if (blk == 7) dmod[sub] = um;
            __syncwarp();
        }
        double z[KT], delta[KT], prop[KT];
#pragma unroll
        for (int j = 0; j < (KT + 1) / 2; ++j) {
            const double2 zz = dr[g * 8 + (j < 7 ? j : 0)];
            z[2 * j] = zz.x;
            if (2 * j + 1 < KT) z[2 * j + 1] = zz.y;
        }
        const double2 uu = dr[g * 8 + 7];                         // (accept uniform, its logarithm)
        make_delta<KT>(M, z, delta);
        const double mod = dmod[g] + 0.5;                         // IP.py:2633
        if (M.map_finding) {
            const double sc = mf->scale(mod);
#pragma unroll
            for (int d = 0; d < KT; ++d) delta[d] = sc * delta[d];
        }
#pragma unroll
        for (int d = 0; d < KT; ++d) prop[d] = cur[d] + delta[d]; // IP.py:2630
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < KT; ++d) prp[d] = prop[d];
            if (out.deltas) {
#pragma unroll
                for (int d = 0; d < KT; ++d) if (d < k) out.deltas[((i - 1) * C + c) * k + d] = delta[d];
            }
            if (out.us) {
                if (M.map_finding) { out.us[((i - 1) * C + c) * 2] = uu.x; out.us[((i - 1) * C + c) * 2 + 1] = mod; }
                else out.us[(i - 1) * C + c] = uu.x;
            }
        }
        const double lp_p = eval_prior<KT>(P, prop);
        const double ll_p = loglik(prop);
        const double sum_p = ll_p + lp_p;
        double la = sum_p - cur_sum;                              // IP.py:2641
        if (M.modulate != 0.0) la = (1.0 / M.modulate) * la;      // IP.py:2649
        const bool accept = la > dr[g * 8 + 7].y;                 // IP.py:2651-2652; NaN -> reject
        if (accept) {                                             // the same in every lane
            cur_lp = lp_p; cur_ll = ll_p; cur_sum = sum_p;
            cur_exp = mhw_exp(cur_sum);
            nacc += 1.0;
            __syncwarp();
            if (lane < KT) cur[lane] = prp[lane];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < L::NACC; ++q) {
                if (sa[q] >= 0) {
                    va[q] = cur[sa[q]] - th0s[sa[q]];
                    if (sb[q] >= 0) vb[q] = cur[sb[q]] - th0s[sb[q]];
                }
            }
        }
        if (M.map_finding) {
            if (lane == 0) mf->after_step(i, dmod[g] + 0.5, sum_p);
            __syncwarp();
        }
        if (--until_row == 0) {
            until_row = M.thin;
            const long long row = next_row++;
            if (out.samples && lane < k) out.samples[(row * C + c) * k + lane] = cur[lane];
            if (lane == 0) {
                if (out.logp) out.logp[row * C + c] = cur_sum;      // IP.py:2661, 2673
                if (out.lprior) out.lprior[row * C + c] = cur_lp;
                if (out.llik) out.llik[row * C + c] = cur_ll;
            }
        }
        if (i >= M.burn_in) add_row(i, cur_sum, cur_exp);
    }
    double* so = out.stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    if (lane == 0) {
        so[PEUQSE_MH_STAT_NACC] = nacc;
        so[PEUQSE_MH_STAT_MAPLOGP] = maplogp;
        so[PEUQSE_MH_STAT_MAPROW] = maprow;
        so[PEUQSE_MH_STAT_LASTLOGP] = cur_sum;
        so[PEUQSE_MH_STAT_SUMEXP] = sumexp;
        so[PEUQSE_MH_STAT_DYN(k)] = mf->dyn;
    }
    if (lane < k) {
        so[PEUQSE_MH_STAT_MAPTHETA + lane] = mapth;
        so[PEUQSE_MH_STAT_LAST(k) + lane] = cur[lane];
    }
#pragma unroll
    for (int q = 0; q < L::NACC; ++q) {
        if (sa[q] >= 0) {
            if (sb[q] >= 0) {
                so[PEUQSE_MH_STAT_S2(k) + sa[q] * k + sb[q]] = acc[q];
                so[PEUQSE_MH_STAT_S2(k) + sb[q] * k + sa[q]] = acc[q];
            } else {
                so[PEUQSE_MH_STAT_S1(k) + sa[q]] = acc[q];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// config C3 shape: the two-site rate-law model with one diagonal-covariance response, up to CPW chains per warp, ONE CTA per SM.
//
// What bounds this kernel is instruction ISSUE: an fp64 instruction holds a sub-partition's issue port for two cycles and every
// other instruction for one (profiles/fp64_latency.py: 8 DFMA + 12 integer instructions take 28 cycles, not 16), so the cost of a
// chain step is 2 x (fp64 instructions) + (all others), and the warps of a sub-partition share one port.  Hence:
//   * a lane evaluates its observations for ALL chains of its warp from one read of the constants (three 16-byte loads per
//     observation whatever the number of chains) with 2 x chains independent exp chains in flight, and what is per chain --
//     draws, proposal, prior, accept, moments, output rows -- belongs to a GROUP of 32/CPW lanes (group g = lane / GL owns chain
//     g of the warp; its state sits in a per-chain shared-memory block);
//   * the rate-law exponents are evaluated in units of ln2/256 from parameters scaled once per proposal, y = (A + (g cov - E)/kBT)
//     * 256/ln2 as two FMAs; k = rint(y) and f = y - k (exact) by the 1.5 * 2^52 shift; exp = 2^(k>>8) T[k & 255] p(f) with a
//     quartic p (truncation 4e-17); T comes from a table with one copy per lane (entry j of copy l at byte 256 j + 8 l of a
//     64 KB block at shared address 65536: never a bank conflict, and the address is ONE byte permute of k and the lane's base);
//     the entry's high word has (j << 12) subtracted so that adding (k << 12) applies 2^(k>>8) without isolating the exponent:
//     8 fp64 + 3 other instructions per exp (the library's exp: ~45 + 25);
//   * the residual uses cov w premultiplied: (y - sim) w = y w - (cov1 w) e1 - (cov2 w) e2, two FMAs;
//   => 23 fp64 + ~8 other instructions per chain-observation (k_mh_warp's general path: 52 + ~30);
//   * the chains are dealt to the SMs and, inside an SM, to the four sub-partitions as evenly as they go (4096 chains on 148 SMs:
//     28 per SM, 7 per sub-partition as warps of 4 + 3 chains).
// The per-chain sums of the lanes' partial sums are reduced with the butterfly of warp_sum, the first log2(CPW) levels
// exchanging the other groups' partials: a chain's value does not depend on CPW or on which warp it landed in.
// Values agree with the general path to a few ulp of the simulated response (log-posterior ~1e-14 relative).
// ---------------------------------------------------------------------------------------------
// 2^(j/256), j = 0..255, correctly rounded (gen_exp_tab256.py); k_mh_tpd stages 32 copies of it (mht_exp)
__device__ const double g_exp_tab256[256] = {
#include "exp_tab256.inc"
};

#define MHT_TAB_SA 65536u          // shared address of the exp table (64 KB: byte 1 of a lookup address is the table index)
#define MHT_OBS_SA 131072u         // shared address of the staged observations
template <int KT>
struct MhtLayout {
    using W = MhwLayout<KT>;
    static constexpr int KPAD = W::KPAD;
    // cur[KPAD] | th0[KPAD] | proposal[KPAD] | scaled proposal[KPAD] | draws[32][2] | modulation uniforms[16] | 'MAP_finding' state[6]
    static constexpr int CDOUBLES = 4 * KPAD + 64 + 16 + 6;
    // a step uses (KT + 1) / 2 blocks of normals and one of uniforms: NBP lanes per step, 32 / NBP steps per pass of draws
    static constexpr int NBP = (KT <= 2) ? 2 : (KT <= 6 ? 4 : 8), SPC = 32 / NBP;
    static constexpr int O_PRP = 2 * KPAD, O_PRS = 3 * KPAD, O_DR = 4 * KPAD;
};
static inline size_t mht_smem_bytes(int n_total) { return (size_t)MHT_OBS_SA + 48u * (size_t)((n_total + 31) & ~31); }

// exp of y * ln2/256 (header above); tab_lane = MHT_TAB_SA + 8 * lane
__device__ __forceinline__ double mht_exp(const DevProblem& P, unsigned int tab_lane, double y) {
    const double t = y + P.mathc3[4];
    const double fn = t - P.mathc3[4];
    const double f = y - fn;                                         // exact, |f| <= 1/2
    const int k = __double2loint(t);
    double Tb;
    asm("ld.shared.f64 %0, [%1];" : "=d"(Tb) : "r"(__byte_perm((unsigned int)k, tab_lane, 0x7604)));
    const double T = __hiloint2double(__double2hiint(Tb) + (int)((unsigned int)k << 12), __double2loint(Tb));
    double u = fma(f, P.mathc3[0], P.mathc3[1]);
    u = fma(f, u, P.mathc3[2]);
    u = fma(f, u, P.mathc3[3]);
    return fma(T, f * u, T);
}

template <int KT, int CPW, int NW, int UNR, int MAXL = CPW>
__global__ void __launch_bounds__(32 * NW, 1) k_mh_tpd(const __grid_constant__ DevProblem P, const __grid_constant__ DevMH M,
                                                       const double* __restrict__ theta0, const MHOut out, int cpc) {
    using L = MhtLayout<KT>;
    constexpr int GL = 32 / CPW;
    constexpr int NACC = (L::W::NSLOT + GL - 1) / GL;
    constexpr int MODEL = PEUQSE_MODEL_TPD2SITE;
    static_assert(KT >= 6 && KT <= GL && (CPW == 1 || CPW == 2 || CPW == 4), "a group of lanes holds one chain's parameters");
    extern __shared__ double smem[];
    __shared__ double s_ext[4];
    __shared__ double s_rs[KT];                                  // 1 / scale of the fast prior
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / GL, gl = lane % GL;
    const unsigned gmask = (CPW == 1) ? 0xffffffffu : (((1u << GL) - 1u) << (grp * GL));
    const unsigned int sa0 = (unsigned int)__cvta_generic_to_shared(smem);
    // shared-memory map: [sa0, 64 KB) the chains' blocks, [64 KB, 128 KB) the exp table, [128 KB, ...) the staged observations
    static_assert((size_t)NW * CPW * L::CDOUBLES * sizeof(double) + 4096 <= MHT_TAB_SA, "the chains' blocks must end below the exp table");
    if (sa0 + (unsigned int)(NW * CPW * L::CDOUBLES * sizeof(double)) > MHT_TAB_SA) __trap();
    double* wbase = smem + (size_t)warp * CPW * L::CDOUBLES;
    double* cur = wbase + grp * L::CDOUBLES;                     // this group's chain
    double* th0s = cur + L::KPAD;
    double* prp = cur + L::O_PRP;
    double* prs = cur + L::O_PRS;
    double2* dr = reinterpret_cast<double2*>(cur + L::O_DR);
    double* dmod = reinterpret_cast<double*>(dr + 32);
    MapFind* mf = reinterpret_cast<MapFind*>(dmod + 16);
    ObsView O;
    O.y = P.y; O.w = P.w; O.cval = P.cval; O.mc = P.mconst;                          // general path: constants in global memory
    const int n = P.n_total;
    const int ngrp = (n + 31) >> 5;
    {
        // staged observations, in groups of 32: 32 x {y w, 1/kBT} | 32 x {cov1, cov2} | 32 x {cov1 w, cov2 w} (1536 bytes a group;
        // a lane's three loads are 512 bytes apart and the lanes' 16 bytes are contiguous); zero rows pad the last group
        double2* ob = reinterpret_cast<double2*>(reinterpret_cast<char*>(smem) + (MHT_OBS_SA - sa0));
        double* tab = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + (MHT_TAB_SA - sa0));
        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
        for (int i = threadIdx.x; i < 32 * ngrp; i += blockDim.x) {
            double y = 0.0, w = 0.0, c1 = 0.0, c2 = 0.0, r = 0.0;
            if (i < n) { y = P.y[i]; w = P.w[i]; c1 = P.mconst[n + i]; c2 = P.mconst[2 * n + i]; r = P.mconst[3 * n + i]; }
            double2* g = ob + 96 * (i >> 5) + (i & 31);
            g[0] = make_double2(y * w, r);
            g[32] = make_double2(c1, c2);
            g[64] = make_double2(c1 * w, c2 * w);
            m0 = fmax(m0, fabs(r)); m1 = fmax(m1, fabs(c1)); m2 = fmax(m2, fabs(c2));
        }
        for (int e = threadIdx.x; e < 256 * 32; e += blockDim.x) {
            const int j = e >> 5;
            const double T = g_exp_tab256[j];                    // 2^(j/256)
            tab[e] = __hiloint2double(__double2hiint(T) - (j << 12), __double2loint(T));
        }
        if (threadIdx.x < 4) s_ext[threadIdx.x] = 0.0;
        if (threadIdx.x < KT) s_rs[threadIdx.x] = (threadIdx.x < P.k && P.scaling_on) ? 1.0 / P.scale[threadIdx.x] : 1.0;
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
            m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        }
        if (lane == 0) {                                         // non-negative doubles order like their bit patterns
            atomicMax(reinterpret_cast<unsigned long long*>(&s_ext[0]), (unsigned long long)__double_as_longlong(m0));
            atomicMax(reinterpret_cast<unsigned long long*>(&s_ext[1]), (unsigned long long)__double_as_longlong(m1));
            atomicMax(reinterpret_cast<unsigned long long*>(&s_ext[2]), (unsigned long long)__double_as_longlong(m2));
        }
        // the additive constant of the log-likelihood: the response's (one diagonal-covariance density), or -- a response longer
        // than responses_observed_max_covmat_size is a SUM of per-observation densities, sum_j -(c_j + t_j^2)/2 (IP.py:2969-2978) --
        // the sum of the per-observation constants, taken once in a fixed order; the value agrees with the term-by-term sum of the
        // general kernel to ~1e-15 relative
        if (warp == 0) {
            const bool per_value = P.resp_mode[0] == PEUQSE_RESP_PER_VALUE;
            double cs = 0.0;
            if (per_value) for (int j = lane; j < n; j += 32) cs += P.cval[j];
            cs = warp_sum(cs);
            if (lane == 0) s_ext[3] = per_value ? cs : P.resp_c[0];
        }
        __syncthreads();
    }
    // the CTA's cpc chains are dealt to its warps as evenly as they go (warp w sits on sub-partition w % 4: with cpc = 28 and
    // eight warps every sub-partition gets 4 + 3 chains); a warp evaluates only the chains it has
    const int wq = cpc / NW, wr = cpc % NW;
    const int nlw = wq + (warp < wr ? 1 : 0);
    const long long cbase = (long long)blockIdx.x * cpc + warp * wq + min(warp, wr);
    if (nlw == 0 || cbase >= M.n_chains) return;
    const long long C = M.n_chains;
    const bool live = grp < nlw && cbase + grp < C;
    const long long c = (cbase + grp < C) ? cbase + grp : C - 1;  // a group without a chain repeats one and stores nothing
    const int k = P.k;

    // log-likelihoods of the proposals in the chains' prs (scaled) / prp blocks; a lane returns its own group's.  `bad`: this
    // group's proposal is outside the bound of the fast form.
    auto loglik_all = [&](bool bad) -> double {
        double pr[MAXL][6];
#pragma unroll
        for (int q = 0; q < MAXL; ++q) {
            const double* pq = wbase + q * L::CDOUBLES + L::O_PRS;
#pragma unroll
            for (int d = 0; d < 6; ++d) pr[q][d] = pq[d];
        }
        const unsigned int tab_lane = MHT_TAB_SA + 8u * lane;
        double part[CPW];
#pragma unroll
        for (int q = 0; q < CPW; ++q) part[q] = 0.0;
        auto sweep = [&](auto nl) {
            constexpr int NL = decltype(nl)::value;
            const double2* po = reinterpret_cast<const double2*>(reinterpret_cast<const char*>(smem) + (MHT_OBS_SA - sa0)) + lane;
            const double2* const po_end = po + 96 * ngrp;
#pragma unroll UNR
            for (; po < po_end; po += 96) {
                const double2 yr = po[0], cv = po[32], cw = po[64];
#pragma unroll
                for (int q = 0; q < NL; ++q) {
                    const double e1 = mht_exp(P, tab_lane, fma(fma(pr[q][4], cv.x, pr[q][0]), yr.y, pr[q][2]));
                    const double e2 = mht_exp(P, tab_lane, fma(fma(pr[q][5], cv.y, pr[q][1]), yr.y, pr[q][3]));
                    const double t = fma(-cw.y, e2, fma(-cw.x, e1, yr.x));   // (y - sim) * w
                    part[q] = fma(t, t, part[q]);
                }
            }
        };
        // MAXL: the launcher deals at most MAXL chains to a warp (the wider sweeps are not compiled)
        if (MAXL == CPW && (CPW == 1 || nlw == CPW)) sweep(std::integral_constant<int, CPW>());
        else if (CPW > 2 && MAXL >= CPW - 1 && nlw == CPW - 1) sweep(std::integral_constant<int, (CPW > 2 ? CPW - 1 : 1)>());
        else if (CPW > 2 && MAXL >= CPW - 2 && nlw == CPW - 2) sweep(std::integral_constant<int, (CPW > 2 ? CPW - 2 : 1)>());
        else sweep(std::integral_constant<int, 1>());
        double v;
        if constexpr (CPW == 1) {
            v = part[0];
        } else if constexpr (CPW == 2) {
            const bool hi = (lane & 16) != 0;
            v = (hi ? part[1] : part[0]) + __shfl_xor_sync(0xffffffffu, hi ? part[0] : part[1], 16);
        } else {
            const bool hi = (lane & 16) != 0, h8 = (lane & 8) != 0;
            const double a0 = (hi ? part[2] : part[0]) + __shfl_xor_sync(0xffffffffu, hi ? part[0] : part[2], 16);
            const double a1 = (hi ? part[3] : part[1]) + __shfl_xor_sync(0xffffffffu, hi ? part[1] : part[3], 16);
            v = (h8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, h8 ? a0 : a1, 8);
        }
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const double total = -0.5 * (s_ext[3] + v);
        double ll = (total != total) ? pq_neg_inf() : total;
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
        if (badm) {                                              // practically never: the general, range-checked path, a chain at a time
#pragma unroll 1
            for (int q = 0; q < CPW; ++q) {
                if ((badm >> (q * GL)) & 1u) {
                    const double* pq = wbase + q * L::CDOUBLES + L::O_PRP;
                    double th[KT];
#pragma unroll
                    for (int d = 0; d < KT; ++d) th[d] = pq[d];
                    const double g = eval_loglik_warp_checked<MODEL, KT, 2>(P, O, th, nullptr, lane);
                    if (grp == q) ll = g;
                }
            }
        }
        return ll;
    };
    // stores the proposal and its scaled form {-E1, -E2, A1, A2, g1, g2} * 256/ln2 (group leader); returns `bad`
    auto put_proposal = [&](const double (&p)[KT]) -> bool {
        const double sc = P.mathc3[5];
        if (gl == 0) {
#pragma unroll
            for (int d = 0; d < KT; ++d) prp[d] = p[d];
            prs[0] = -p[0] * sc; prs[1] = -p[1] * sc; prs[2] = p[2] * sc; prs[3] = p[3] * sc; prs[4] = p[4] * sc; prs[5] = p[5] * sc;
        }
        const double rkm = s_ext[0];
        const double b1 = fabs(p[2]) + (fabs(p[4]) * s_ext[1] + fabs(p[0])) * rkm;
        const double b2 = fabs(p[3]) + (fabs(p[5]) * s_ext[2] + fabs(p[1])) * rkm;
        return !(fmax(b1, b2) < 700.0);
    };

    if (gl < KT) {
        const double t0 = (gl < k) ? theta0[c * k + gl] : 0.0;
        cur[gl] = t0; th0s[gl] = t0;
    }
    if (gl == 0) mf->init();
    for (int e = gl; e < 80; e += GL) reinterpret_cast<double*>(dr)[e] = 0.0;
    __syncwarp();
    // accumulator slots of this lane (slot s of the chain lives in lane s % GL of its group)
    int sa[NACC], sb[NACC];
    double acc[NACC], va[NACC], vb[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) {
        const int s = gl + GL * q;
        int a = -1, b = -1;
        if (s < L::W::NPAIR) {
            int e = 0;
            for (int d = 0; d < KT; ++d)
                for (int d2 = d; d2 < KT; ++d2) { if (e == s) { a = d; b = d2; } ++e; }
            if (a >= k || b >= k) a = -1;
        } else if (s < L::W::NSLOT) {
            a = s - L::W::NPAIR; b = -1;
            if (a >= k) a = -1;
        }
        sa[q] = a; sb[q] = b; acc[q] = 0.0; va[q] = 0.0; vb[q] = (a >= 0 && b < 0) ? 1.0 : 0.0;
    }
    double nacc = 0.0, maplogp = pq_neg_inf(), maprow = -1.0, sumexp = 0.0, mapth = 0.0;
    double cur_lp, cur_ll;
    {
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = cur[d];
        cur_lp = eval_prior<KT>(P, th);
        const bool bad = put_proposal(th);
        __syncwarp();
        cur_ll = loglik_all(bad);
    }
    double cur_sum = cur_ll + cur_lp;
    auto add_row = [&](long long row, double lp, double explp) {
        if (lp > maplogp) {                                       // first maximum (IP.py:1925-1927)
            maplogp = lp; maprow = (double)row;
            if (gl < KT) mapth = cur[gl];
        }
        sumexp += explp;
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = fma(va[q], vb[q], acc[q]);   // first moments: fma(dd, 1, s1) == s1 + dd
    };
    {
        const double lp0 = combine_logp(cur_lp, cur_ll);
        if (live) {
            if (out.samples && gl < k) out.samples[c * k + gl] = cur[gl];
            if (gl == 0) {
                if (out.logp) out.logp[c] = lp0;
                if (out.lprior) out.lprior[c] = 0.0;
                if (out.llik) out.llik[c] = 0.0;
            }
        }
        if (M.burn_in <= 0) add_row(0, lp0, exp(lp0));
    }
    double cur_exp = exp(cur_sum);
    long long next_row = 1;
    int until_row = M.thin;
    const bool fast_prior = P.prior_mode == PEUQSE_PRIOR_DIAG && P.prior_identity && P.kp == k;
    const int sub = lane / L::NBP, blk = ((lane % L::NBP) == L::NBP - 1) ? 7 : (lane % L::NBP);
    for (long long i = 1; i < M.n_steps; ++i) {
        const int g = (int)((i - 1) % L::SPC);
        if (g == 0) {                                             // the draws of the next SPC steps, a chain at a time
            const unsigned long long step = (unsigned long long)(M.first_step + i + sub);
            __syncwarp();
#pragma unroll 1
            for (int q = 0; q < nlw; ++q) {
                const long long cq = (cbase + q < C) ? cbase + q : C - 1;
                double um = 0.0;
                const double2 d2 = mhw_draw(M.seed, (unsigned long long)(M.chain_offset + cq), step, blk, k, &um);
                double* bq = wbase + q * L::CDOUBLES + L::O_DR;
                reinterpret_cast<double2*>(bq)[lane] = d2;
                if (blk == 7) bq[64 + sub] = um;
            }
            __syncwarp();
        }
        double lp_p;
        bool bad;
        {
            double z[KT], delta[KT], prop[KT];
#pragma unroll
            for (int j = 0; j < (KT + 1) / 2; ++j) {
                const double2 zz = dr[g * L::NBP + j];
                z[2 * j] = zz.x;
                if (2 * j + 1 < KT) z[2 * j + 1] = zz.y;
            }
            make_delta<KT>(M, z, delta);
            const double mod = dmod[g] + 0.5;                     // IP.py:2633
            if (M.map_finding) {
                const double sc = mf->scale(mod);
#pragma unroll
                for (int d = 0; d < KT; ++d) delta[d] = sc * delta[d];
            }
#pragma unroll
            for (int d = 0; d < KT; ++d) prop[d] = cur[d] + delta[d]; // IP.py:2630
            bad = put_proposal(prop);
            if (gl == 0) {
                if (live && out.deltas) {
#pragma unroll
                    for (int d = 0; d < KT; ++d) if (d < k) out.deltas[((i - 1) * C + c) * k + d] = delta[d];
                }
                if (live && out.us) {
                    const double u = dr[g * L::NBP + L::NBP - 1].x;
                    if (M.map_finding) { out.us[((i - 1) * C + c) * 2] = u; out.us[((i - 1) * C + c) * 2 + 1] = mod; }
                    else out.us[(i - 1) * C + c] = u;
                }
            }
            if (fast_prior) {
                // diagonal prior in natural order: theta / scale as a multiplication by the reciprocal (the log-prior moves in its
                // last digit; the fast form's budget is 1e-12)
                bool ok = true;
                if (P.has_ub | P.has_lb) {
#pragma unroll
                    for (int d = 0; d < KT; ++d) {
                        if (d < k) {
                            if (P.has_ub) ok = ok && (prop[d] <= P.ub[d]);
                            if (P.has_lb) ok = ok && (prop[d] >= P.lb[d]);
                        }
                    }
                }
                double q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int d = 0; d < KT; ++d) {
                    if (d < k) {
                        const double t = (prop[d] * s_rs[d] - P.mus[d]) * P.pw[d];
                        if (d & 1) q1 = fma(t, t, q1); else q0 = fma(t, t, q0);
                    }
                }
                lp_p = ok ? -0.5 * (P.prior_c + (q0 + q1)) : pq_neg_inf();
            } else {
                lp_p = mhw_prior<KT>(P, prop);
            }
        }
        __syncwarp();
        const double ll_p = loglik_all(bad);
        const double sum_p = ll_p + lp_p;
        double la = sum_p - cur_sum;                              // IP.py:2641
        if (M.modulate != 0.0) la = (1.0 / M.modulate) * la;      // IP.py:2649
        const bool accept = la > dr[g * L::NBP + L::NBP - 1].y;   // IP.py:2651-2652; NaN -> reject
        if (accept) {                                             // the same in every lane of the group
            cur_lp = lp_p; cur_ll = ll_p; cur_sum = sum_p;
            cur_exp = mhw_exp(cur_sum);
            nacc += 1.0;
            if (gl < KT) cur[gl] = prp[gl];
            __syncwarp(gmask);
#pragma unroll
            for (int q = 0; q < NACC; ++q) {
                if (sa[q] >= 0) {
                    va[q] = cur[sa[q]] - th0s[sa[q]];
                    if (sb[q] >= 0) vb[q] = cur[sb[q]] - th0s[sb[q]];
                }
            }
        }
        if (M.map_finding) {
            if (gl == 0) mf->after_step(i, dmod[g] + 0.5, sum_p);
        }
        __syncwarp();
        if (--until_row == 0) {
            until_row = M.thin;
            const long long row = next_row++;
            if (live) {
                if (out.samples && gl < k) out.samples[(row * C + c) * k + gl] = cur[gl];
                if (gl == 0) {
                    if (out.logp) out.logp[row * C + c] = cur_sum;      // IP.py:2661, 2673
                    if (out.lprior) out.lprior[row * C + c] = cur_lp;
                    if (out.llik) out.llik[row * C + c] = cur_ll;
                }
            }
        }
        if (i >= M.burn_in) add_row(i, cur_sum, cur_exp);
    }
    if (!live) return;
    double* so = out.stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    if (gl == 0) {
        so[PEUQSE_MH_STAT_NACC] = nacc;
        so[PEUQSE_MH_STAT_MAPLOGP] = maplogp;
        so[PEUQSE_MH_STAT_MAPROW] = maprow;
        so[PEUQSE_MH_STAT_LASTLOGP] = cur_sum;
        so[PEUQSE_MH_STAT_SUMEXP] = sumexp;
        so[PEUQSE_MH_STAT_DYN(k)] = mf->dyn;
    }
    if (gl < k) {
        so[PEUQSE_MH_STAT_MAPTHETA + gl] = mapth;
        so[PEUQSE_MH_STAT_LAST(k) + gl] = cur[gl];
    }
#pragma unroll
    for (int q = 0; q < NACC; ++q) {
        if (sa[q] >= 0) {
            if (sb[q] >= 0) {
                so[PEUQSE_MH_STAT_S2(k) + sa[q] * k + sb[q]] = acc[q];
                so[PEUQSE_MH_STAT_S2(k) + sb[q] * k + sa[q]] = acc[q];
            } else {
                so[PEUQSE_MH_STAT_S1(k) + sa[q]] = acc[q];
            }
        }
    }
}

namespace {
// tuning override for the shape crossover (profiles/mh_variants.py): PEUQSE_MH_FEW_CHAINS=<n>
long long mh_few_chains() {
    static long long v = -1;
    if (v < 0) {
        const char* e = getenv("PEUQSE_MH_FEW_CHAINS");
        v = e ? atoll(e) : MH_FEW_CHAINS;
    }
    return v;
}

template <int MODEL, int KT>
int launch_mh(const peuqse_problem* p, const DevMH& M, const double* th0, const MHOut& out, cudaStream_t st,
              const DevScen& SC = pq_no_scen()) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    // few chains with few observations: one warp per chain, thread-order evaluation (latency); many: one thread per chain
    const bool few = M.n_chains <= mh_few_chains();
    if (SC.y) {
        // scenario batch.  Many observations: one warp per chain, lanes stride over the scenario's own rows of constants in
        // global memory (L1/L2-resident; nothing is staged because every warp of a CTA may belong to another scenario).
        // Few observations: thread-order evaluation, one warp per chain when there are few chains, else one thread per chain.
        if (p->d.n_total > PQ_THREAD_SHAPE_MAX_N) {
            auto kern = k_mh_warp<MODEL, KT, 2, -1, false, 4, true>;
            const size_t wsmem = (size_t)MHW_WARPS * MhwLayout<KT>::WDOUBLES * sizeof(double);
            const long long blocks = (M.n_chains + MHW_WARPS - 1) / MHW_WARPS;
            kern<<<(unsigned)blocks, 32 * MHW_WARPS, wsmem, st>>>(p->d, M, th0, out, 0, SC);
        } else if (few) {
            auto kern = k_mh_warp<MODEL, KT, 2, -1, true, 4, true>;
            const size_t wsmem = smem + (size_t)MHW_WARPS * MhwLayout<KT>::WDOUBLES * sizeof(double);
            PQ_CUDA_OK(pq_allow_smem((const void*)kern, wsmem));
            const long long blocks = (M.n_chains + MHW_WARPS - 1) / MHW_WARPS;
            kern<<<(unsigned)blocks, 32 * MHW_WARPS, wsmem, st>>>(p->d, M, th0, out, use_smem, SC);
        } else {
            auto kern = k_mh<MODEL, KT, true>;
            PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
            const long long blocks = (M.n_chains + 127) / 128;
            kern<<<(unsigned)blocks, 128, smem, st>>>(p->d, M, th0, out, use_smem, SC);
        }
        PQ_CUDA_OK(cudaGetLastError());
        return PEUQSE_OK;
    }
    if (p->d.n_total > PQ_THREAD_SHAPE_MAX_N || few) {
        void (*kern)(const DevProblem, const DevMH, const double*, const MHOut, int, const DevScen) = k_mh_warp<MODEL, KT, 2, -1>;
        if (p->d.n_total <= PQ_THREAD_SHAPE_MAX_N) kern = k_mh_warp<MODEL, KT, 2, -1, true>;
        // 3 CTAs (12 chains) per SM at 168 registers: the 4-CTA shape spills inside the observation loop and is 8 % slower
        else if (use_smem) kern = k_mh_warp<MODEL, KT, 2, 1, false, 3>;
        if constexpr (MODEL == PEUQSE_MODEL_TPD2SITE && KT <= 8) {
            // the rate-law model with one diagonal-covariance response (config C3): k_mh_tpd, ONE CTA per SM, the chains dealt evenly to
            // the SMs and their sub-partitions.  Up to 12 chains per SM: a warp each (12 warps); more: up to four chains per warp
            // (8 warps) -- measured on 256 ... 16384 chains (profiles/c3_shapes_sweep.py).  PEUQSE_MH_SHAPE=4|5|7|8 forces a shape,
            // 1 the general kernel.
            const bool fast_ok = p->d.n_total > PQ_THREAD_SHAPE_MAX_N && p->d.n_resp == 1 &&
                                 (p->d.resp_mode[0] == PEUQSE_RESP_MVN_DIAG || p->d.resp_mode[0] == PEUQSE_RESP_PER_VALUE) &&
                                 mht_smem_bytes(p->d.n_total) + 1024 <= (size_t)p->max_smem_optin;   // + the kernel's static shared memory
            const char* fe = getenv("PEUQSE_MH_SHAPE");
            int shape = fe ? atoi(fe) : 0;
            if (shape == 0) shape = (M.n_chains <= 12LL * p->sm_count) ? 7 : 4;
            if (fast_ok && shape >= 4) {
                void (*tk)(const DevProblem, const DevMH, const double*, const MHOut, int) = k_mh_tpd<KT, 4, 8, 1>;
                int tnw = 8, maxl = 4;
                if (shape == 5) { tk = k_mh_tpd<KT, 2, 16, 1>; tnw = 16; maxl = 2; }
                if (shape == 7) { tk = k_mh_tpd<KT, 1, 12, 2>; tnw = 12; maxl = 1; }
                if (shape == 8) { tk = k_mh_tpd<KT, 4, 12, 1, 3>; tnw = 12; maxl = 3; }
                const long long cap = (long long)tnw * maxl;
                const long long slots = (long long)p->sm_count;
                const long long waves = (M.n_chains + cap * slots - 1) / (cap * slots);
                const long long ctas0 = waves * slots;
                const int cpc = (int)((M.n_chains + ctas0 - 1) / ctas0);
                const size_t tsmem = mht_smem_bytes(p->d.n_total);
                PQ_CUDA_OK(pq_allow_smem((const void*)tk, tsmem));
                tk<<<(unsigned)((M.n_chains + cpc - 1) / cpc), 32 * tnw, tsmem, st>>>(p->d, M, th0, out, cpc);
                PQ_CUDA_OK(cudaGetLastError());
                return PEUQSE_OK;
            }
        }
        const int nw = MHW_WARPS;
        const size_t wsmem = smem + (size_t)nw * MhwLayout<KT>::WDOUBLES * sizeof(double);
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, wsmem));
        const long long blocks = (M.n_chains + nw - 1) / nw;
        kern<<<(unsigned)blocks, 32 * nw, wsmem, st>>>(p->d, M, th0, out, use_smem, SC);
    } else {
        auto kern = k_mh<MODEL, KT>;
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
        const long long blocks = (M.n_chains + 127) / 128;
        kern<<<(unsigned)blocks, 128, smem, st>>>(p->d, M, th0, out, use_smem, SC);
    }
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
}  // namespace

static int mh_run_impl(const peuqse_problem* p, const peuqse_mh_spec* s, const DevScen& SC, const double* d_theta0,
                       double* d_samples, double* d_logp, double* d_logprior, double* d_loglik, double* d_deltas, double* d_us,
                       double* d_stats, void* stream) {
    if (!p || !s || !d_theta0 || !d_stats) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (s->k != p->d.k) return pq_set_error(PEUQSE_ERR_INVALID, "mh k != problem k");
    if (s->n_chains < 1 || s->n_steps < 1 || s->thin < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad mh sizes");
    if (s->k > 14) return pq_set_error(PEUQSE_ERR_INVALID, "mh supports k <= 14 (7 Philox blocks of normals per step)");
    if ((unsigned long long)(s->first_step + s->n_steps) * 8ULL > 0xffffffffULL)
        return pq_set_error(PEUQSE_ERR_INVALID, "first_step + n_steps exceeds the 2^29 steps a Philox stream addresses");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL) return pq_set_error(PEUQSE_ERR_INVALID, "mh needs a built-in model");
    if (p->d.any_full) return pq_set_error(PEUQSE_ERR_INVALID, "mh does not support full-covariance responses");
    DevMH M;
    M.n_chains = s->n_chains; M.n_steps = s->n_steps; M.burn_in = s->burn_in; M.first_step = s->first_step;
    M.chain_offset = s->chain_offset; M.seed = s->seed; M.modulate = s->modulate; M.k = s->k; M.thin = s->thin; M.map_finding = s->map_finding ? 1 : 0;
    M.chol_diag = 1;
    for (int i = 0; i < PQ_MAXK * PQ_MAXK; ++i) M.chol[i] = 0.0;
    for (int d = 0; d < s->k; ++d)
        for (int e = 0; e < s->k; ++e) {
            const double v = s->step_chol[d * s->k + e];
            M.chol[d * s->k + e] = v;
            if (d != e && v != 0.0) M.chol_diag = 0;
        }
    MHOut out{d_samples, d_logp, d_logprior, d_loglik, d_deltas, d_us, d_stats};
    cudaStream_t st = (cudaStream_t)stream;
    const int kb = pq_k_bucket(p->d.k);
    switch (p->d.model_id) {
        case PEUQSE_MODEL_QUADRATIC: return launch_mh<PEUQSE_MODEL_QUADRATIC, 2>(p, M, d_theta0, out, st, SC);
        case PEUQSE_MODEL_TPD2SITE: return launch_mh<PEUQSE_MODEL_TPD2SITE, 6>(p, M, d_theta0, out, st, SC);
        case PEUQSE_MODEL_LANGMUIR_LOG: return launch_mh<PEUQSE_MODEL_LANGMUIR_LOG, 2>(p, M, d_theta0, out, st, SC);
        case PEUQSE_MODEL_LINEAR:
#define MH_L(KT, ...) return launch_mh<PEUQSE_MODEL_LINEAR, KT>(p, M, d_theta0, out, st, SC)
            PQ_DISPATCH_K(kb, MH_L, 0)
#undef MH_L
            break;
        default: break;
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "unknown model");
}

extern "C" int peuqse_mh_run(const peuqse_problem* p, const peuqse_mh_spec* s, const double* d_theta0, double* d_samples,
                             double* d_logp, double* d_logprior, double* d_loglik, double* d_deltas, double* d_us,
                             double* d_stats, void* stream) {
    return mh_run_impl(p, s, pq_no_scen(), d_theta0, d_samples, d_logp, d_logprior, d_loglik, d_deltas, d_us, d_stats, stream);
}

extern "C" int peuqse_mh_run_scenarios(const peuqse_problem* p, const peuqse_scenarios* sc, const peuqse_mh_spec* s,
                                       int64_t chains_per_scenario, const double* d_theta0, double* d_samples, double* d_logp,
                                       double* d_logprior, double* d_loglik, double* d_deltas, double* d_us, double* d_stats,
                                       void* stream) {
    if (!p || !sc || !s) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (!sc->has_y) return pq_set_error(PEUQSE_ERR_INVALID, "scenarios have no observed values yet (peuqse_scenarios_synthesize / _set_observed)");
    if (sc->n_total != p->d.n_total || sc->n_mconst != p->d.n_mconst) return pq_set_error(PEUQSE_ERR_INVALID, "scenarios belong to another problem");
    if (chains_per_scenario < 1 || s->n_chains != sc->S * chains_per_scenario)
        return pq_set_error(PEUQSE_ERR_INVALID, "n_chains must equal n_scenarios * chains_per_scenario");
    DevScen SC{sc->d_y, sc->d_mc, (long long)chains_per_scenario, sc->n_total, sc->n_mconst};
    return mh_run_impl(p, s, SC, d_theta0, d_samples, d_logp, d_logprior, d_loglik, d_deltas, d_us, d_stats, stream);
}
