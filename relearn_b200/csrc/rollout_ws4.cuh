// rollout_ws4.cuh -- K2v: K2w (rollout_cartpole_ws_kernel, rollout.cu) with the two changes of the K2x / K2y experiments
// that do not lengthen the dynamics warp's instruction stream (profiles/r2_summary.md):
//  * an AUX warp (sixth warp) draws every Philox block -- the logit-space thresholds and the would-be reset states --
//    four steps per chunk into 16-step shared-memory rings, up to 11 steps ahead (progress counters in shared memory, no
//    barrier).  In K2w the policy warps drew them on the step's chain (~180 clk per step on average);
//  * shared memory is addressed through explicit 32-bit shared-window addresses: K2w's pointers made ptxas rematerialise
//    the window base (S2UR SR_CgaCtaId + ULEA) inside both loops.
// Protocol, operations and operands are K2w's: bit-identical trajectories.  Included by rollout.cu after rollout_ws3.cuh
// (yk_* shared-memory helpers, named barriers).
#pragma once

constexpr int VK_ENVS = 16, VK_THREADS = 192, VK_RING = 16, VK_CHUNK = 4, VK_AHEAD = 11, VK_SYNC = 160;

struct VkShared {
    float4 sw4[4 * GK_PAIRS];
    float tail[4 + GK_REM_TABLE_MAX];
    float rows[VK_ENVS][8];            // obs 0..4, flags (bit 0: this env takes the step, bit 1: some env of the CTA does)
    uint32_t act[VK_ENVS];
    float thr[VK_RING][VK_ENVS];       // logit-space thresholds
    double2 slot[VK_RING][VK_ENVS][2]; // would-be reset states (x, x') | (theta, theta')
    uint32_t prod, cons, done, pad;
};
#define VK_OFF(member) ((uint32_t)offsetof(VkShared, member))

#ifdef RL_WS_CLOCKS  // measurement build only (scripts/ws_clocks.sh): per-phase clocks of the dynamics / policy loops of two CTAs
__device__ __forceinline__ long long vk_clk(double dep, uint32_t dep2) {
    long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) : "d"(dep), "r"(dep2) : "memory");
    return c;
}
#define VK_CLK(var, dep, dep2) const long long var = vk_clk(dep, dep2)
#else
#define VK_CLK(var, dep, dep2)
#endif

__global__ void __launch_bounds__(VK_THREADS, 2) rollout_cartpole_ws4_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int LANES = 8, PPL = GK_PAIRS / LANES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char gk_smem[];
    VkShared &sh = *reinterpret_cast<VkShared *>(gk_smem);
    // (through a shuffle: ptxas otherwise rematerialises the window base at its uses in the loops)
    const uint32_t sb = __shfl_sync(FULL, (uint32_t)__cvta_generic_to_shared(gk_smem), 0);
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;
    stage_pair_weights(a.net, sh.sw4, sh.tail, p, rem_table ? (int)p.max_steps + 1 : 0);
    const uint32_t rem_addr = sb + VK_OFF(tail) + 8;
    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? yk_ldf(rem_addr + 4u * r) : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool later_cta = (int)blockIdx.x >= a.sm_count;
    const int dyn_warp = later_cta ? a.dyn_second : a.dyn_first, aux_warp = later_cta ? a.aux_second : a.aux_first;
    const bool is_dyn = hw_warp == dyn_warp, is_aux = hw_warp == aux_warp;
    const int warp = hw_warp - (hw_warp > dyn_warp ? 1 : 0) - (hw_warp > aux_warp ? 1 : 0);  // policy warp index 0..3
    const uint64_t e_base = (uint64_t)blockIdx.x * VK_ENVS;
    const uint32_t t0 = a.noise.step_counter;
    const uint64_t seed = a.noise.seed;
    const int F = a.F;
    const uint64_t FE = (uint64_t)F * a.E;
    const uint32_t cap = a.min_steps ? a.min_steps + a.slack : 0;  // no env takes more steps than this
    LaneStats st;
    st.init();
    bool contributes = false;

    // ---- aux: one chunk = VK_CHUNK steps x 16 envs; lane = (env, half) handles steps k0 + 2 half + {0, 1} ----
    auto aux_fill = [&](uint32_t k0) {
        const int el = lane & 15, half = lane >> 4;
        const uint64_t eg = e_base + el, lg = a.lane_offset + (eg < a.E ? eg : 0);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
            const uint32_t k = k0 + 2u * (uint32_t)half + (uint32_t)j;
            const uint32_t ring = k & (VK_RING - 1);
            uint32_t oa[4], o0[4], o1[4];
            // policies/actor.rs:42-55: the actor's uniform of step k as the logit-space threshold (rl_logit_threshold)
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ACTOR * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), oa);
            // cartpole.rs:103-115: four uniform draws in field order = blocks 0 (x, x') and 1 (theta, theta')
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o0);
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u + 1u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o1);
            yk_stf(sb + VK_OFF(thr) + 4u * (ring * VK_ENVS + el), rl_logit_threshold(rl_u32_to_f32(oa[0])));
            const double x = rl_u64_to_uniform((uint64_t)o0[0] | ((uint64_t)o0[1] << 32), p.reset_low, p.reset_scale);
            const double xd = rl_u64_to_uniform((uint64_t)o0[2] | ((uint64_t)o0[3] << 32), p.reset_low, p.reset_scale);
            const double th = rl_u64_to_uniform((uint64_t)o1[0] | ((uint64_t)o1[1] << 32), p.reset_low, p.reset_scale);
            const double thd = rl_u64_to_uniform((uint64_t)o1[2] | ((uint64_t)o1[3] << 32), p.reset_low, p.reset_scale);
            const uint32_t sa = sb + VK_OFF(slot) + 32u * (ring * VK_ENVS + el);
            yk_std2(sa, make_double2(x, xd));
            yk_std2(sa + 16, make_double2(th, thd));
        }
    };
    if (threadIdx.x == 0) { sh.prod = 0; sh.cons = 0; sh.done = 0; }
    if (is_aux) {
        aux_fill(0);
        aux_fill(VK_CHUNK);
    }
    __syncthreads();
    if (is_aux && lane == 0) yk_stu(sb + VK_OFF(prod), 2 * VK_CHUNK);

    if (is_aux) {
        // ------------------------------ aux warp ------------------------------
        uint32_t k0 = 2 * VK_CHUNK;
        while (k0 <= cap + 1) {
            // slots of steps k0 - 16 .. k0 - 13 are reused: their readers (thresholds at iteration k, reset states at
            // k - 1) are done once the dynamics warp is at iteration >= k0 - 12
            uint32_t c = yk_ldu(sb + VK_OFF(cons));
            bool over = false;
            while (k0 > c + VK_AHEAD) {
                if (yk_ldu(sb + VK_OFF(done))) { over = true; break; }
                __nanosleep(64);
                c = yk_ldu(sb + VK_OFF(cons));
            }
            if (over) break;
            aux_fill(k0);
            __threadfence_block();
            __syncwarp();
            k0 += VK_CHUNK;
            if (lane == 0) yk_stu(sb + VK_OFF(prod), k0);
        }
    } else if (is_dyn) {
        // ------------------------------ dynamics warp: lane = (env el, action act) ------------------------------
        const int el = lane & 15, act = lane >> 4;
        const uint64_t e = e_base + el;
        const bool valid = e < a.E;
        const uint64_t e_safe = valid ? e : 0;
        const float rem_full = remaining_feature(p.max_steps);
        const uint32_t slot0 = sb + VK_OFF(slot) + 32u * (uint32_t)el, row = sb + VK_OFF(rows) + 32u * (uint32_t)el;
        const uint32_t act_addr = sb + VK_OFF(act) + 4u * (uint32_t)el;
        // the reset state of noise step t0 + k, from the ring the aux warp keeps ahead
        auto fresh_state = [&](uint32_t k, EnvT::State &f) {
            const uint32_t sa = slot0 + 512u * (k & (VK_RING - 1));
            const double2 lo = yk_ldd2(sa), hi = yk_ldd2(sa + 16);
            f.x = lo.x; f.xd = lo.y; f.th = hi.x; f.thd = hi.y;
            f.meta = 0x80000000u | p.max_steps;
        };
        EnvT::State s;
        s.x = s.xd = s.th = s.thd = 0.0;
        s.meta = 0x80000000u | p.max_steps;
        uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
        uint32_t i = 0, cur_len = 0;
        int succ_last = RL_TERMINATE, succ_prev = RL_TERMINATE;
        double n_eps = 0.0, sum_el = 0.0, sum_el2 = 0.0;
        float cur_obs[5] = {0, 0, 0, 0, 0}, last_obs[5] = {0, 0, 0, 0, 0};
        {
            EnvT::State f;
            fresh_state(0, f);
            if (n > 0) {
                s = f;
                cur_obs[0] = (float)s.x; cur_obs[1] = (float)s.xd; cur_obs[2] = (float)s.th; cur_obs[3] = (float)s.thd;
                cur_obs[4] = rem_full;
            }
        }
        bool any = __any_sync(FULL, n > 0);
        if (act == 0) {
            yk_stf(row, cur_obs[0]); yk_stf(row + 4, cur_obs[1]); yk_stf(row + 8, cur_obs[2]);
        } else {
            yk_stf(row + 12, cur_obs[3]); yk_stf(row + 16, cur_obs[4]);
            yk_stu(row + 20, (n > 0 ? 1u : 0u) | (any ? 2u : 0u));
        }
        __syncwarp();
        named_bar_arrive(1, VK_SYNC);
        uint32_t it = 0;  // loop counter (= step index of the envs still active)
#ifdef RL_WS_CLOCKS
        long long ck[5] = {0, 0, 0, 0, 0};
#endif
        while (any) {
            VK_CLK(c0, s.x, it);
            const bool active = n > 0;
            if ((it & 3u) == 0u && lane == 0) yk_stu(sb + VK_OFF(cons), it);
            if ((it & 3u) == 3u) {  // the reset states of steps it + 1 .. it + 4
                while (yk_ldu(sb + VK_OFF(prod)) < it + 5u) { }
                asm volatile("fence.acq_rel.cta;" ::: "memory");
            }
            // before the action is known: this lane's candidate step, the would-be reset state, and the `remaining`
            // feature of the next observation
            EnvT::State cand = s;
            const int cand_sc = EnvT::step_fast(p, cand, (uint32_t)act);
            EnvT::State fresh;
            fresh_state(it + 1, fresh);
            const uint32_t r_now = s.meta & 0x7FFFFFFFu;
            const float rem_cont = remaining_feature(r_now > 0 ? r_now - 1 : 0);
            VK_CLK(c1, cand.x + cand.thd + fresh.x, (uint32_t)cand_sc + __float_as_uint(rem_cont));
            named_bar_sync(2, VK_SYNC);
            VK_CLK(c2, 0.0, 0u);
            const uint32_t action = yk_ldu(act_addr);
            const int src = el + 16 * (int)action;
            EnvT::State post;
            post.x = __shfl_sync(FULL, cand.x, src);
            post.xd = __shfl_sync(FULL, cand.xd, src);
            post.th = __shfl_sync(FULL, cand.th, src);
            post.thd = __shfl_sync(FULL, cand.thd, src);
            post.meta = __shfl_sync(FULL, cand.meta, src);
            const int sc = __shfl_sync(FULL, cand_sc, src);
            const bool ended = sc != RL_CONTINUE;  // steps.rs:116-124: the next call starts a new episode
            s.x = ended ? fresh.x : post.x; s.xd = ended ? fresh.xd : post.xd;
            s.th = ended ? fresh.th : post.th; s.thd = ended ? fresh.thd : post.thd;
            s.meta = ended ? fresh.meta : post.meta;
            uint32_t n_next = n;
            if (active) {
                n_next = n - 1;
                if (ended && n_next <= a.slack) n_next = 0;  // take_steps.rs:83-88
            }
            float nobs[5];
            nobs[0] = (float)s.x; nobs[1] = (float)s.xd; nobs[2] = (float)s.th; nobs[3] = (float)s.thd;
            nobs[4] = ended ? rem_full : rem_cont;
            if (act == 0) {
                yk_stf(row, nobs[0]); yk_stf(row + 4, nobs[1]); yk_stf(row + 8, nobs[2]);
            } else {
                yk_stf(row + 12, nobs[3]); yk_stf(row + 16, nobs[4]);
            }
            const bool any_next = __any_sync(FULL, n_next > 0);
            if (act == 1) yk_stu(row + 20, (n_next > 0 ? 1u : 0u) | (any_next ? 2u : 0u));
            __syncwarp();
            named_bar_arrive(1, VK_SYNC);
            VK_CLK(c3, 0.0, 0u);
            // ---- off the chain: the rest of the step record and the statistics ----
            if (active) {
                const uint64_t is = (uint64_t)i * a.E + e_safe;
                if (act == 1) a.reward[is] = 1.0f;  // cartpole.rs:140
                if (act == 0) a.succ[is] = (uint8_t)sc;
                if (sc == RL_INTERRUPT && act == 0) {  // rare: once per max_steps; the post-step observation (remaining == 0)
                    const float io4 = remaining_feature(post.meta & 0x7FFFFFFFu);
                    const uint64_t io = (uint64_t)i * FE + e_safe;
                    a.next_obs[io] = (float)post.x;
                    a.next_obs[io + a.E] = (float)post.xd;
                    a.next_obs[io + 2 * a.E] = (float)post.th;
                    a.next_obs[io + 3 * a.E] = (float)post.thd;
                    if (F > 4) a.next_obs[io + 4 * a.E] = io4;
                }
                cur_len += 1;
                if (ended) {
                    const double ld = (double)cur_len;
                    n_eps += 1.0;
                    sum_el += ld;
                    sum_el2 = fma(ld, ld, sum_el2);
                    cur_len = 0;
                }
#pragma unroll
                for (int f = 0; f < 5; ++f) last_obs[f] = cur_obs[f];
                succ_prev = succ_last;
                succ_last = sc;
                i += 1;
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) cur_obs[f] = nobs[f];
            n = n_next;
            any = any_next;
            it += 1;
#ifdef RL_WS_CLOCKS
            const long long c4 = vk_clk(sum_el2 + cur_obs[0], i + cur_len);
            ck[0] += c1 - c0; ck[1] += c2 - c1; ck[2] += c3 - c2; ck[3] += c4 - c3; ck[4] += 1;
#endif
        }
#ifdef RL_WS_CLOCKS
        if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
            printf("K2v dyn cta %d: %lld iterations; clk per iteration: candidate step %.1f, wait for the action %.1f, select + publish %.1f, "
                   "off-chain record %.1f\n", (int)blockIdx.x, ck[4], (double)ck[0] / ck[4], (double)ck[1] / ck[4], (double)ck[2] / ck[4],
                   (double)ck[3] / ck[4]);
#endif
        if (lane == 0) yk_stu(sb + VK_OFF(done), 1u);
        st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
        st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
        if (valid && act == 0) {
            // VecBuffer::end_experience -> finalize_last_episode (buffers/mod.rs:237-261); same thread as the in-loop
            // stores of succ, so program order applies
            uint32_t len = i, flags = 0;
            double eps = n_eps;
            if (i > 0 && succ_last == RL_CONTINUE) {
                len = i - 1;
                flags = 1;
                a.succ[(uint64_t)len * a.E + e] = RL_PAD;
                if (len > 0 && succ_prev == RL_CONTINUE) {
                    flags = 3;
                    a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
#pragma unroll
                    for (int f = 0; f < 5; ++f)
                        if (f < F) a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = last_obs[f];
                    eps += 1.0;
                }
            }
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
            st.v[ST_STORED_STEPS] = (double)len;
            st.v[ST_STORED_EPS] = eps;
            contributes = true;
        }
    } else {
        // ------------------------------ policy warps: 4 envs x 8 threads (K2c<8>) ------------------------------
        const int grp = lane >> 3, sub = lane & 7, el = 4 * warp + grp;
        const uint64_t e = e_base + el;
        const bool valid = e < a.E;
        const uint64_t e_safe = valid ? e : 0;
        const float b2d = sh.tail[0];
        float4 wA[PPL], wB[PPL], wC[PPL];
        float2 wD[PPL];
#pragma unroll
        for (int u = 0; u < PPL; ++u) {
            const int q = sub + LANES * u;
            wA[u] = sh.sw4[q]; wB[u] = sh.sw4[GK_PAIRS + q]; wC[u] = sh.sw4[2 * GK_PAIRS + q];
            wD[u] = make_float2(sh.sw4[3 * GK_PAIRS + q].x, sh.sw4[3 * GK_PAIRS + q].y);
        }
        // Thread `sub` stores column `sub` of the step record: observation feature sub (< F) or, sub == 5, the action.
        const bool stores_obs = valid && sub < 5 && sub < F, stores_action = valid && sub == 5;
        float *obs_ptr = a.obs + (uint64_t)(sub < 5 ? sub : 0) * a.E + e_safe;
        uint8_t *act_ptr = a.action + e_safe;
        const uint32_t row = sb + VK_OFF(rows) + 32u * (uint32_t)el, mine_addr = row + 4u * (uint32_t)(sub < 5 ? sub : 0);
        const uint32_t thr_addr = sb + VK_OFF(thr) + 4u * (uint32_t)el, act_addr = sb + VK_OFF(act) + 4u * (uint32_t)el;
#ifdef RL_WS_CLOCKS
        long long pk[4] = {0, 0, 0, 0};
#endif
        for (uint32_t i = 0;; ++i) {
            VK_CLK(q0, 0.0, i);
            if ((i & 3u) == 0u) {  // thresholds of steps i .. i + 3
                // (`done`: the dynamics warp has left its loop and the aux warp may have stopped; this iteration only breaks)
                while (yk_ldu(sb + VK_OFF(prod)) < i + 4u && !yk_ldu(sb + VK_OFF(done))) { }
                asm volatile("fence.acq_rel.cta;" ::: "memory");
            }
            const float theta = yk_ldf(thr_addr + 64u * (i & (VK_RING - 1)));
            named_bar_sync(1, VK_SYNC);
            VK_CLK(q1, 0.0, 0u);
            const float4 ov = yk_ld4(row);
            const float4 tv = yk_ld4(row + 16);
            const float mine = yk_ldf(mine_addr);
            const float ob4 = tv.x;
            const uint32_t flags = __float_as_uint(tv.y);
#ifdef RL_WS_CLOCKS
            if ((flags & 2u) == 0u) {
                if (warp == 0 && lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
                    printf("K2v policy cta %d: %lld iterations; clk per iteration: wait for the row %.1f, row -> action published %.1f, "
                           "record stores %.1f\n", (int)blockIdx.x, pk[3], (double)pk[0] / pk[3], (double)pk[1] / pk[3], (double)pk[2] / pk[3]);
            }
#endif
            if ((flags & 2u) == 0u) break;
            const bool active = (flags & 1u) != 0u;
            const float2 o0 = make_float2(ov.x, ov.x), o1 = make_float2(ov.y, ov.y), o2 = make_float2(ov.z, ov.z);
            const float2 o3 = make_float2(ov.w, ov.w), o4 = make_float2(ob4, ob4);
            float2 pre[PPL];
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].x, wA[u].y), o0, make_float2(wC[u].z, wC[u].w));
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wA[u].z, wA[u].w), o1, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].x, wB[u].y), o2, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wB[u].z, wB[u].w), o3, pre[u]);
#pragma unroll
            for (int u = 0; u < PPL; ++u) pre[u] = __ffma2_rn(make_float2(wC[u].x, wC[u].y), o4, pre[u]);
            float2 za = make_float2(0.0f, 0.0f), zc = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < PPL; ++u) {
                const float2 h = make_float2(fmaxf(pre[u].x, 0.0f), fmaxf(pre[u].y, 0.0f));
                if (u & 1) zc = __ffma2_rn(wD[u], h, zc);
                else za = __ffma2_rn(wD[u], h, za);
            }
            za = __fadd2_rn(za, zc);
            float d = za.x + za.y;
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
            d += b2d;
            const uint32_t action = d < theta ? 0u : 1u;  // policies/actor.rs:42-55 (rl_logit_threshold)
            if (sub == 0) yk_stu(act_addr, action);
            __syncwarp();
            named_bar_arrive(2, VK_SYNC);
            VK_CLK(q2, 0.0, action);
            // ---- off the chain: the observation and the action of the step record ----
            if (active && stores_obs) *obs_ptr = mine;
            if (active && stores_action) *act_ptr = (uint8_t)action;
            obs_ptr += FE;
            act_ptr += a.E;
#ifdef RL_WS_CLOCKS
            const long long q3 = vk_clk(0.0, (uint32_t)(uintptr_t)act_ptr);
            pk[0] += q1 - q0; pk[1] += q2 - q1; pk[2] += q3 - q2; pk[3] += 1;
#endif
        }
    }
    block_reduce_stats(st, contributes, a.partials);
}
