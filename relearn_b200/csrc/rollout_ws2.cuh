// rollout_ws2.cuh -- K2x: the warp-specialised CartPole rollout (K2w) with everything that does not depend on the
// sampled action taken off the per-step dependent chain.  Included by rollout.cu inside its anonymous namespace.
//
// What the ncu capture of K2w at E = 4096 showed (profiles/r1i_k2w_e4096.md, per-SASS stall samples): a step took
// ~1205 clk, made of  policy chain (bar.sync 1 -> hidden layer -> logit reduce -> bar.arrive 2, ~650 clk; the FFMA2
// sequence was issued as ONE dependent chain) + the dynamics warp's post-action section (shuffle-select, f64->f32
// conversions, mailbox stores, ~180 clk), with the Philox blocks the policy warps drew for thresholds / would-be reset
// states (~180 clk per step on average, 460 clk every fourth step) landing on that chain whenever they made the policy
// warps late, and the dynamics warp's own loop (f64 step ~580 clk + trajectory stores / statistics ~230 clk) barely
// hidden under it.  K2x changes, each bit-identical in its results:
//
//  * an AUX warp (sixth warp of the CTA) draws every Philox block: the logit-space thresholds and the would-be reset
//    states, FOUR steps per chunk into 16-step shared-memory rings, up to 11 steps ahead of the dynamics warp (progress
//    counters in shared memory, no barrier).  It also computes the reset states' f32 observations and their Head;
//  * the f64 step is software-pipelined (CartPoleEnv::Head, envs.cuh): theta_{t+1} = theta_t + dt * theta'_t does not
//    depend on the action, so sin/cos of the NEXT angle and the refined reciprocals of the next step's two
//    angular-acceleration denominators are computed during step t, beside its dependent chain instead of at the head
//    of step t + 1's (42 -> ~26 dependent f64 operations);
//  * each dynamics lane (env, a) prepares, BEFORE the action is known, the complete mailbox row it would publish if
//    action a is sampled (candidate or reset observation as f32, step-limit feature, activity flag, successor code);
//    once the action arrives the chosen lane stores its row and the warp arrives on the barrier -- the shuffles that
//    select the warp's own next state come after the hand-off, off the policy's path;
//  * the policy warps store the whole step record (observation, action, reward and the previous step's successor code,
//    one column per thread through one running pointer); the dynamics warp keeps only episode statistics in its loop
//    and reads the last two successor codes / the popped observation back from the trajectory for finalisation.
//
// Roles: warps {policy x 4, dynamics, aux}; which hardware warp carries the dynamics / aux role is a launch parameter
// (sub-partition placement, measured in launch_ws2).
#pragma once

constexpr int XK_ENVS = 16, XK_THREADS = 192, XK_RING = 16, XK_CHUNK = 4, XK_AHEAD = 11;

struct XkSlot {      // would-be reset state of one env at one noise step
    double x, xd, th, thd;
    double sn, cs, yp, ym;  // CartPoleEnv::Head of th
};

struct XkShared {
    float4 sw4[4 * GK_PAIRS];
    float tail[4 + GK_REM_TABLE_MAX];
    // mailbox [step parity][action][env]: (x, x', theta, theta') | (remaining, flags, successor of the previous step, -);
    // flags bit 0: this env takes the step, bit 1: some env of the CTA may
    float4 stage[2][2][XK_ENVS][2];
    uint32_t act[XK_ENVS];
    float thr[XK_RING][XK_ENVS];     // logit-space thresholds
    float4 fobs[XK_RING][XK_ENVS];   // reset observations as f32
    XkSlot slot[XK_RING][XK_ENVS];
    uint32_t prod, cons, done, pad;  // steps filled by the aux warp / iteration of the dynamics warp / loop over
};

// Named barriers, alternating by step parity so that a warp running ahead can never arrive twice in one phase:
// rows of step t are handed over on barrier 1 + 2 (t & 1), the action of step t on barrier 2 + 2 (t & 1).
__device__ __forceinline__ int xk_bar_rows(uint32_t t) { return 1 + 2 * (int)(t & 1u); }
__device__ __forceinline__ int xk_bar_act(uint32_t t) { return 2 + 2 * (int)(t & 1u); }

__device__ __forceinline__ uint32_t xk_ld_volatile(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void xk_st_volatile(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

__global__ void __launch_bounds__(XK_THREADS, 2) rollout_cartpole_ws2_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int LANES = 8, PPL = GK_PAIRS / LANES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char gk_smem[];
    XkShared &sh = *reinterpret_cast<XkShared *>(gk_smem);
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;
    stage_pair_weights(a.net, sh.sw4, sh.tail, p, rem_table ? (int)p.max_steps + 1 : 0);
    const float *rem = sh.tail + 2;
    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? rem[r] : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool later_cta = (int)blockIdx.x >= a.sm_count;
    const int dyn_warp = later_cta ? a.dyn_second : a.dyn_first, aux_warp = later_cta ? a.aux_second : a.aux_first;
    const bool is_dyn = hw_warp == dyn_warp, is_aux = hw_warp == aux_warp;
    const int warp = hw_warp - (hw_warp > dyn_warp ? 1 : 0) - (hw_warp > aux_warp ? 1 : 0);  // policy warp index 0..3
    const uint64_t e_base = (uint64_t)blockIdx.x * XK_ENVS;
    const uint32_t t0 = a.noise.step_counter;
    const uint64_t seed = a.noise.seed;
    const int F = a.F;
    const uint64_t FE = (uint64_t)F * a.E;
    const uint32_t cap = a.min_steps ? a.min_steps + a.slack : 0;  // no env takes more steps than this
    LaneStats st;
    st.init();
    bool contributes = false;

    // ---- aux: one chunk = XK_CHUNK steps x 16 envs; lane = (env, half) handles steps k0 + 2 half + {0, 1} ----
    auto aux_fill = [&](uint32_t k0) {
        const int el = lane & 15, half = lane >> 4;
        const uint64_t eg = e_base + el, lg = a.lane_offset + (eg < a.E ? eg : 0);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t k = k0 + 2u * (uint32_t)half + (uint32_t)j;
            const int ring = (int)(k & (XK_RING - 1));
            uint32_t oa[4], o0[4], o1[4];
            // policies/actor.rs:42-55: the actor's uniform of step k as the logit-space threshold (rl_logit_threshold)
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ACTOR * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), oa);
            // cartpole.rs:103-115: four uniform draws in field order = blocks 0 (x, x') and 1 (theta, theta')
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o0);
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u + 1u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o1);
            sh.thr[ring][el] = rl_logit_threshold(rl_u32_to_f32(oa[0]));
            XkSlot s;
            s.x = rl_u64_to_uniform((uint64_t)o0[0] | ((uint64_t)o0[1] << 32), p.reset_low, p.reset_scale);
            s.xd = rl_u64_to_uniform((uint64_t)o0[2] | ((uint64_t)o0[3] << 32), p.reset_low, p.reset_scale);
            s.th = rl_u64_to_uniform((uint64_t)o1[0] | ((uint64_t)o1[1] << 32), p.reset_low, p.reset_scale);
            s.thd = rl_u64_to_uniform((uint64_t)o1[2] | ((uint64_t)o1[3] << 32), p.reset_low, p.reset_scale);
            EnvT::Head h;
            EnvT::head_of(p, s.th, h);
            s.sn = h.sn; s.cs = h.cs; s.yp = h.yp; s.ym = h.ym;
            sh.slot[ring][el] = s;
            sh.fobs[ring][el] = make_float4((float)s.x, (float)s.xd, (float)s.th, (float)s.thd);
        }
    };
    if (threadIdx.x == 0) { sh.prod = 0; sh.cons = 0; sh.done = 0; }
    if (is_aux) {
        aux_fill(0);
        aux_fill(XK_CHUNK);
    }
    __syncthreads();
    if (is_aux && lane == 0) xk_st_volatile(&sh.prod, 2 * XK_CHUNK);
    // (the consumers' first checks of `prod` come at steps 3 / 4, after many barrier hand-offs with each other; the aux
    //  warp's store above is ordered before its next chunk by program order)

    if (is_aux) {
        // ------------------------------ aux warp ------------------------------
        uint32_t k0 = 2 * XK_CHUNK;
        while (k0 <= cap + 1) {
            // slots of steps k0 - 16 .. k0 - 13 are reused: their readers (thresholds at iteration k, reset states at k - 1)
            // are done once the dynamics warp is at iteration >= k0 - 12
            uint32_t c = xk_ld_volatile(&sh.cons);
            bool over = false;
            while (k0 > c + XK_AHEAD) {
                if (xk_ld_volatile(&sh.done)) { over = true; break; }
                __nanosleep(64);
                c = xk_ld_volatile(&sh.cons);
            }
            if (over) break;
            aux_fill(k0);
            __threadfence_block();
            __syncwarp();
            k0 += XK_CHUNK;
            if (lane == 0) xk_st_volatile(&sh.prod, k0);
        }
    } else if (is_dyn) {
        // ------------------------------ dynamics warp: lane = (env el, action act) ------------------------------
        const int el = lane & 15, act = lane >> 4;
        const uint64_t e = e_base + el;
        const bool valid = e < a.E;
        const uint64_t e_safe = valid ? e : 0;
        const float rem_full = remaining_feature(p.max_steps);
        const double y_ml = EnvT::rcp_refined(p.mass_length_pole);
        EnvT::State s;
        EnvT::Head h;
        uint32_t n = (valid && a.min_steps) ? a.min_steps + a.slack : 0;
        uint32_t i = 0, cur_len = 0;
        double n_eps = 0.0, sum_el = 0.0, sum_el2 = 0.0;
        {
            const XkSlot &f = sh.slot[0][el];
            s.x = f.x; s.xd = f.xd; s.th = f.th; s.thd = f.thd;
            s.meta = 0x80000000u | p.max_steps;
            h.sn = f.sn; h.cs = f.cs; h.yp = f.yp; h.ym = f.ym;
        }
        bool any = __any_sync(FULL, n > 0);
        sh.stage[0][act][el][0] = n > 0 ? sh.fobs[0][el] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        sh.stage[0][act][el][1] = make_float4(n > 0 ? rem_full : 0.0f, __uint_as_float((n > 0 ? 1u : 0u) | (any ? 2u : 0u)),
                                              __uint_as_float((uint32_t)RL_PAD), 0.0f);
        __syncwarp();
        named_bar_arrive(xk_bar_rows(0), XK_THREADS - 32);
        uint32_t it = 0;  // loop counter (= step index of the envs still active)
        while (any) {
            const bool active = n > 0;
            if ((it & 3u) == 0u && lane == 0) xk_st_volatile(&sh.cons, it);
            if ((it & 3u) == 3u) {  // the reset states of steps it + 1 .. it + 4
                while (xk_ld_volatile(&sh.prod) < it + 5u) { }
                __threadfence_block();
            }
            // ---- before the action is known: this lane's candidate step, the Head of the next angle, the would-be
            //      reset state and the complete mailbox row this lane publishes if its action is the sampled one ----
            EnvT::State cand = s;
            const int cand_sc = EnvT::step_with_head(p, cand, h, y_ml, (uint32_t)act);
            EnvT::Head hn;
            EnvT::head_of(p, cand.th, hn);  // cand.th = s.th + dt * s.thd: the same in both lanes of an env
            const int ring = (int)((it + 1u) & (XK_RING - 1));
            const XkSlot &fr = sh.slot[ring][el];
            const double2 f0 = *reinterpret_cast<const double2 *>(&fr.x), f1 = *reinterpret_cast<const double2 *>(&fr.th);
            const double2 f2 = *reinterpret_cast<const double2 *>(&fr.sn), f3 = *reinterpret_cast<const double2 *>(&fr.yp);
            const float4 ff = sh.fobs[ring][el];
            const uint32_t r_now = s.meta & 0x7FFFFFFFu;
            const float rem_cont = remaining_feature(r_now > 0 ? r_now - 1 : 0);
            const bool ended_a = cand_sc != RL_CONTINUE;  // steps.rs:116-124: the next call starts a new episode
            uint32_t n_a = n;
            if (active) {
                n_a = n - 1;
                if (ended_a && n_a <= a.slack) n_a = 0;  // take_steps.rs:83-88
            }
            float4 row0;
            row0.x = ended_a ? ff.x : (float)cand.x;
            row0.y = ended_a ? ff.y : (float)cand.xd;
            row0.z = ended_a ? ff.z : (float)cand.th;
            row0.w = ended_a ? ff.w : (float)cand.thd;
            const float rem_a = ended_a ? rem_full : rem_cont;
            const float sc_f = __uint_as_float(active ? (uint32_t)cand_sc : (uint32_t)RL_PAD);
            // Both rows go out BEFORE the action is known; the policy warps pick the row of the action they sample (they
            // know it first), so nothing this warp does after the action is on their path.  "Some env may take the step":
            // exact unless an episode end inside the slack stops the last envs (take_steps.rs:83-88) -- see the loop exit.
            const bool cons_any = __any_sync(FULL, n_a > 0);
            sh.stage[(it + 1u) & 1u][act][el][0] = row0;
            sh.stage[(it + 1u) & 1u][act][el][1] = make_float4(rem_a, __uint_as_float((n_a > 0 ? 1u : 0u) | (cons_any ? 2u : 0u)), sc_f, 0.0f);
            __syncwarp();
            named_bar_arrive(xk_bar_rows(it + 1u), XK_THREADS - 32);
            // The Head of the next angle must be COMPUTED before the barrier as well: the compiler otherwise sinks the whole
            // polynomial below it, next to its first use (seen in the first capture, profiles/r2_summary.md).  Empty
            // volatile asm statements keep their order relative to bar.sync.
            asm volatile("" : "+d"(hn.sn), "+d"(hn.cs), "+d"(hn.yp), "+d"(hn.ym));
            asm volatile("" : "+d"(cand.x), "+d"(cand.xd), "+d"(cand.thd));
            // ---- the action ----
            named_bar_sync(xk_bar_act(it), XK_THREADS - 32);
            const uint32_t action = sh.act[el];
            const bool chosen = action == (uint32_t)act;
            const bool any_next = __any_sync(FULL, chosen && n_a > 0);
            // ---- off the policy's path: this warp's own next state ----
            const int src = el + 16 * (int)action;
            const double px = __shfl_sync(FULL, cand.x, src), pxd = __shfl_sync(FULL, cand.xd, src);
            const double pthd = __shfl_sync(FULL, cand.thd, src);
            const uint32_t pmeta = __shfl_sync(FULL, cand.meta, src);
            const int sc = __shfl_sync(FULL, cand_sc, src);
            const bool ended = sc != RL_CONTINUE;
            if (active && sc == RL_INTERRUPT && act == 0) {  // rare: once per max_steps; the post-step observation (remaining == 0)
                const uint64_t io = (uint64_t)i * FE + e_safe;
                a.next_obs[io] = (float)px;
                a.next_obs[io + a.E] = (float)pxd;
                a.next_obs[io + 2 * a.E] = (float)cand.th;
                a.next_obs[io + 3 * a.E] = (float)pthd;
                if (F > 4) a.next_obs[io + 4 * a.E] = remaining_feature(pmeta & 0x7FFFFFFFu);
            }
            s.x = ended ? f0.x : px; s.xd = ended ? f0.y : pxd;
            s.th = ended ? f1.x : cand.th; s.thd = ended ? f1.y : pthd;
            s.meta = ended ? (0x80000000u | p.max_steps) : pmeta;
            h.sn = ended ? f2.x : hn.sn; h.cs = ended ? f2.y : hn.cs;
            h.yp = ended ? f3.x : hn.yp; h.ym = ended ? f3.y : hn.ym;
            if (active) {
                cur_len += 1;
                if (ended) {
                    const double ld = (double)cur_len;
                    n_eps += 1.0;
                    sum_el += ld;
                    sum_el2 = fma(ld, ld, sum_el2);
                    cur_len = 0;
                }
                i += 1;
                n = n - 1;
                if (ended && n <= a.slack) n = 0;
            }
            if (!any_next && cons_any) {
                // the rows of step it + 1 said "may take the step": the policy warps evaluate it once more (every env
                // inactive) and need a terminator to leave on
                sh.stage[it & 1u][act][el][1] = make_float4(0.0f, __uint_as_float(0u), __uint_as_float((uint32_t)RL_PAD), 0.0f);
                __syncwarp();
                named_bar_arrive(xk_bar_rows(it + 2u), XK_THREADS - 32);
                named_bar_sync(xk_bar_act(it + 1u), XK_THREADS - 32);
            }
            any = any_next;
            it += 1;
        }
        if (lane == 0) xk_st_volatile(&sh.done, 1u);
        st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
        st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
        st.cur_len = i;  // carried over the barrier below
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
        // Thread `sub` stores column `sub` of the step record through one running pointer: observation feature sub (< F),
        // the action (5), the reward (6) or -- one step late, when the dynamics warp has published it -- the successor (7).
        const bool col4 = valid && ((sub < 5 && sub < F) || sub == 6), col1 = valid && (sub == 5 || sub == 7);
        unsigned char *ptr;
        uint64_t stride;
        if (sub < 5) { ptr = reinterpret_cast<unsigned char *>(a.obs + (uint64_t)(sub < F ? sub : 0) * a.E + e_safe); stride = FE * 4; }
        else if (sub == 5) { ptr = a.action + e_safe; stride = a.E; }
        else if (sub == 6) { ptr = reinterpret_cast<unsigned char *>(a.reward + e_safe); stride = a.E * 4; }
        else { ptr = a.succ + e_safe - a.E; stride = a.E; }  // (never dereferenced before the first advance)
        bool prev_active = false;
        uint32_t a_prev = 0;  // (the rows of step 0 are the same for both actions)
        for (uint32_t i = 0;; ++i) {
            if ((i & 3u) == 0u) {  // thresholds of steps i .. i + 3
                // (`done`: the dynamics warp has left its loop and the aux warp may have stopped; this iteration only breaks)
                while (xk_ld_volatile(&sh.prod) < i + 4u && !xk_ld_volatile(&sh.done)) { }
                __threadfence_block();
            }
            const float theta = sh.thr[i & (XK_RING - 1)][el];
            named_bar_sync(xk_bar_rows(i), XK_THREADS - 32);
            const float4 *rp = &sh.stage[i & 1u][a_prev][el][0];  // the row the dynamics warp prepared for the sampled action
            const float4 ov = rp[0];
            const float4 tv = rp[1];
            const float mine = reinterpret_cast<const float *>(rp)[sub < 5 ? sub : 0];
            const uint32_t flags = __float_as_uint(tv.y);
            const float2 o0 = make_float2(ov.x, ov.x), o1 = make_float2(ov.y, ov.y), o2 = make_float2(ov.z, ov.z);
            const float2 o3 = make_float2(ov.w, ov.w), o4 = make_float2(tv.x, tv.x);
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
                const float2 hh = make_float2(fmaxf(pre[u].x, 0.0f), fmaxf(pre[u].y, 0.0f));
                if (u & 1) zc = __ffma2_rn(wD[u], hh, zc);
                else za = __ffma2_rn(wD[u], hh, za);
            }
            za = __fadd2_rn(za, zc);
            float d = za.x + za.y;
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
            d += b2d;
            const uint32_t action = d < theta ? 0u : 1u;  // policies/actor.rs:42-55 (rl_logit_threshold)
            const bool active = (flags & 1u) != 0u;
            const uint8_t sc_prev = (uint8_t)__float_as_uint(tv.z);
            if ((flags & 2u) == 0u) {  // every env of the CTA is done: the successor code of the last step is still owed
                if (sub == 7 && valid && prev_active) *ptr = sc_prev;
                break;
            }
            if (sub == 0) sh.act[el] = action;
            a_prev = action;
            __syncwarp();
            named_bar_arrive(xk_bar_act(i), XK_THREADS - 32);
            // ---- off the chain: the step record ----
            const float v4 = sub == 6 ? 1.0f : mine;  // cartpole.rs:140
            const uint8_t v1 = sub == 5 ? (uint8_t)action : sc_prev;
            const bool on = sub == 7 ? prev_active : active;
            if (col4 && on) *reinterpret_cast<float *>(ptr) = v4;
            if (col1 && on) *ptr = v1;
            ptr += stride;
            prev_active = active;
        }
    }
    __syncthreads();  // the step record (policy warps' stores) is complete and ordered before the fix-ups below
    if (is_dyn) {
        const int el = lane & 15, act = lane >> 4;
        const uint64_t e = e_base + el;
        if (e < a.E && act == 0) {
            // VecBuffer::end_experience -> finalize_last_episode (buffers/mod.rs:237-261)
            const uint32_t i = st.cur_len;
            uint32_t len = i, flags = 0;
            double eps = st.v[ST_EPS];
            const int succ_last = i > 0 ? (int)a.succ[(uint64_t)(i - 1) * a.E + e] : RL_TERMINATE;
            if (i > 0 && succ_last == RL_CONTINUE) {
                len = i - 1;
                flags = 1;
                a.succ[(uint64_t)len * a.E + e] = RL_PAD;
                const int succ_prev = len > 0 ? (int)a.succ[(uint64_t)(len - 1) * a.E + e] : RL_TERMINATE;
                if (len > 0 && succ_prev == RL_CONTINUE) {
                    flags = 3;
                    a.succ[(uint64_t)(len - 1) * a.E + e] = RL_INTERRUPT;
                    for (int f = 0; f < F; ++f)  // the popped step's observation
                        a.next_obs[((uint64_t)(len - 1) * F + f) * a.E + e] = a.obs[((uint64_t)len * F + f) * a.E + e];
                    eps += 1.0;
                }
            }
            a.lane_len[e] = len;
            a.lane_flags[e] = (uint8_t)flags;
            st.v[ST_STORED_STEPS] = (double)len;
            st.v[ST_STORED_EPS] = eps;
            contributes = true;
        }
    }
    st.cur_len = 0;
    block_reduce_stats(st, contributes, a.partials);
}
