// rollout_ws3.cuh -- K2y: the warp-specialised CartPole rollout with a six-role CTA.  Included by rollout.cu inside its
// anonymous namespace.  A MEASURED EXPERIMENT (profiles/r2_summary.md): bit-identical to K2w, 6 % slower at the bench
// size; selected by RL_WS_VARIANT=3 only.  Kept for its parts: the software-pipelined f64 step (CartPoleEnv::Head), the
// setmaxnreg register split by role and the explicit shared-memory addressing K2v (rollout_ws4.cuh) uses.
//
// K2x (round 2's first attempt, in the history) took the action-independent work off the per-step chain but left all of it on ONE dynamics
// warp: ~430 SASS instructions per step (124 of them f64, 64 moves of polynomial constants), which that warp issues in
// ~1050 clk -- its own instruction stream, not the chain, set the period (profiles/r2_summary.md, capture r2b).  K2y
// keeps K2x's protocol (both candidate rows published before the action is known, the policy warps pick theirs) and
//
//  * gives CartPoleEnv::Head of the next angle to a HEAD warp: lane (env, a) reads the (theta, theta') the dynamics
//    warp published for "the previous action was a" and computes sin / cos / the two refined reciprocals of the NEXT
//    angle theta + dt theta' while the policy warps evaluate the step; the dynamics warp picks the row of the action it
//    learnt one step earlier.  The dynamics warp is left with the 26-deep f64 step, the row conversion and bookkeeping;
//  * addresses shared memory through explicit 32-bit shared-window addresses (ld.shared / st.shared): the generic
//    pointers of K2x made ptxas rematerialise the shared-window base (S2R SR_CgaCtaId + LEA, ~40 clk) three times per
//    iteration in every role;
//  * splits the register file by role with setmaxnreg: the CTA is two warpgroups, {policy x 4} at 152 registers (112 of
//    them hold the hidden layer's weights) and {dynamics, head, aux, spare} at 104; 256 threads x 128 at launch, two
//    CTAs per SM as before.
//
// Same operations on the same operands as K2c<8> / K2w: bit-identical trajectories (tests/test_gpu_envs.py).
#pragma once

struct XkSlot {      // would-be reset state of one env at one noise step
    double x, xd, th, thd;
    double sn, cs, yp, ym;  // CartPoleEnv::Head of th
};
// Named barriers, alternating by step parity so that a warp running ahead can never arrive twice in one phase:
// rows of step t are handed over on barrier 1 + 2 (t & 1), the action of step t on barrier 2 + 2 (t & 1).
__device__ __forceinline__ int xk_bar_rows(uint32_t t) { return 1 + 2 * (int)(t & 1u); }
__device__ __forceinline__ int xk_bar_act(uint32_t t) { return 2 + 2 * (int)(t & 1u); }

constexpr int YK_ENVS = 16, YK_THREADS = 256, YK_RING = 16, YK_CHUNK = 4, YK_AHEAD = 11;
constexpr int YK_SYNC = 192;  // policy x 4 + dynamics + head

struct YkShared {
    float4 sw4[4 * GK_PAIRS];
    float tail[4 + GK_REM_TABLE_MAX];
    // mailbox [step parity][action][env]: (x, x', theta, theta') | (remaining, flags, successor of the previous step, -);
    // flags bit 0: this env takes the step, bit 1: some env of the CTA may
    float4 stage[2][2][YK_ENVS][2];
    double2 dynout[2][2][YK_ENVS];       // (theta, theta') of the state the row describes, f64, for the head warp
    double2 heads[2][2][YK_ENVS][2];     // Head of the following angle: (sn, cs) | (yp, ym)
    uint32_t act[YK_ENVS];
    float thr[YK_RING][YK_ENVS];         // logit-space thresholds
    float4 fobs[YK_RING][YK_ENVS];       // reset observations as f32
    XkSlot slot[YK_RING][YK_ENVS];       // reset states and their Head
    uint32_t prod, cons, done, pad;      // steps filled by the aux warp / iteration of the dynamics warp / loop over
};

__device__ __forceinline__ float4 yk_ld4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void yk_st4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double2 yk_ldd2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void yk_std2(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ float yk_ldf(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t yk_ldu(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void yk_stu(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void yk_stf(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

#define YK_OFF(member) ((uint32_t)offsetof(YkShared, member))

__global__ void __launch_bounds__(YK_THREADS, 2) rollout_cartpole_ws3_kernel(CartPoleEnv::Params p, RolloutArgs a) {
    using EnvT = CartPoleEnv;
    constexpr int LANES = 8, PPL = GK_PAIRS / LANES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char gk_smem[];
    YkShared &sh = *reinterpret_cast<YkShared *>(gk_smem);
    // (through a shuffle: ptxas otherwise rematerialises the window base -- S2R SR_CgaCtaId + LEA -- at its uses in the loops)
    const uint32_t sb = __shfl_sync(0xffffffffu, (uint32_t)__cvta_generic_to_shared(gk_smem), 0);
    const bool rem_table = p.max_steps != 0 && p.max_steps < GK_REM_TABLE_MAX;
    stage_pair_weights(a.net, sh.sw4, sh.tail, p, rem_table ? (int)p.max_steps + 1 : 0);
    const uint32_t rem_addr = sb + YK_OFF(tail) + 8;
    auto remaining_feature = [&](uint32_t r) {
        return p.max_steps == 0 ? 0.0f : rem_table ? yk_ldf(rem_addr + 4u * r) : (float)__ddiv_rn((double)r, (double)p.max_steps);
    };
    const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool later_cta = (int)blockIdx.x >= a.sm_count;
    // roles of warps 4..7 (one per sub-partition): launch parameters, measured placements in launch_ws3
    const int dyn_warp = later_cta ? a.dyn_second : a.dyn_first, aux_warp = later_cta ? a.aux_second : a.aux_first;
    const int head_warp = later_cta ? a.head_second : a.head_first;
    const bool is_policy = hw_warp < 4, is_dyn = hw_warp == dyn_warp, is_aux = hw_warp == aux_warp, is_head = hw_warp == head_warp;
    const uint64_t e_base = (uint64_t)blockIdx.x * YK_ENVS;
    const uint32_t t0 = a.noise.step_counter;
    const uint64_t seed = a.noise.seed;
    const int F = a.F;
    const uint64_t FE = (uint64_t)F * a.E;
    const uint32_t cap = a.min_steps ? a.min_steps + a.slack : 0;  // no env takes more steps than this
    LaneStats st;
    st.init();
    bool contributes = false;

    // ---- aux: one chunk = YK_CHUNK steps x 16 envs; lane = (env, half) handles steps k0 + 2 half + {0, 1} ----
    auto aux_fill = [&](uint32_t k0) {
        const int el = lane & 15, half = lane >> 4;
        const uint64_t eg = e_base + el, lg = a.lane_offset + (eg < a.E ? eg : 0);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
            const uint32_t k = k0 + 2u * (uint32_t)half + (uint32_t)j;
            const uint32_t ring = k & (YK_RING - 1);
            uint32_t oa[4], o0[4], o1[4];
            // policies/actor.rs:42-55: the actor's uniform of step k as the logit-space threshold (rl_logit_threshold)
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ACTOR * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), oa);
            // cartpole.rs:103-115: four uniform draws in field order = blocks 0 (x, x') and 1 (theta, theta')
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o0);
            rl_philox4x32_10((uint32_t)lg, (uint32_t)(lg >> 32), t0 + k, (uint32_t)RL_STREAM_ENV_RESET * 64u + 1u, (uint32_t)seed,
                             (uint32_t)(seed >> 32), o1);
            yk_stf(sb + YK_OFF(thr) + 4u * (ring * YK_ENVS + el), rl_logit_threshold(rl_u32_to_f32(oa[0])));
            const double x = rl_u64_to_uniform((uint64_t)o0[0] | ((uint64_t)o0[1] << 32), p.reset_low, p.reset_scale);
            const double xd = rl_u64_to_uniform((uint64_t)o0[2] | ((uint64_t)o0[3] << 32), p.reset_low, p.reset_scale);
            const double th = rl_u64_to_uniform((uint64_t)o1[0] | ((uint64_t)o1[1] << 32), p.reset_low, p.reset_scale);
            const double thd = rl_u64_to_uniform((uint64_t)o1[2] | ((uint64_t)o1[3] << 32), p.reset_low, p.reset_scale);
            EnvT::Head h;
            EnvT::head_of(p, th, h);
            const uint32_t sa = sb + YK_OFF(slot) + 64u * (ring * YK_ENVS + el);
            yk_std2(sa, make_double2(x, xd));
            yk_std2(sa + 16, make_double2(th, thd));
            yk_std2(sa + 32, make_double2(h.sn, h.cs));
            yk_std2(sa + 48, make_double2(h.yp, h.ym));
            yk_st4(sb + YK_OFF(fobs) + 16u * (ring * YK_ENVS + el), make_float4((float)x, (float)xd, (float)th, (float)thd));
        }
    };
    if (threadIdx.x == 0) { sh.prod = 0; sh.cons = 0; sh.done = 0; }
    if (is_aux) {
        aux_fill(0);
        aux_fill(YK_CHUNK);
    }
    __syncthreads();
    if (is_aux && lane == 0) yk_stu(sb + YK_OFF(prod), 2 * YK_CHUNK);
    // (the consumers' first checks of `prod` come at steps 3 / 4, after many barrier hand-offs with each other; the aux
    //  warp's store above is ordered before its next chunk by program order)

    if (is_policy) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        // ------------------------------ policy warps: 4 envs x 8 threads (K2c<8>) ------------------------------
        const int warp = hw_warp, grp = lane >> 3, sub = lane & 7, el = 4 * warp + grp;
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
        // row of (parity 0, action 0, env el); parity toggles 1024 B, the action 512 B
        uint32_t row = sb + YK_OFF(stage) + 32u * (uint32_t)el;  // (the rows of step 0 are the same for both actions)
        const uint32_t mine_off = 4u * (uint32_t)(sub < 5 ? sub : 0);
        const uint32_t thr_addr = sb + YK_OFF(thr) + 4u * (uint32_t)el, act_addr = sb + YK_OFF(act) + 4u * (uint32_t)el;
        const uint32_t row_base = sb + YK_OFF(stage) + 32u * (uint32_t)el;
        for (uint32_t i = 0;; ++i) {
            if ((i & 3u) == 0u) {  // thresholds of steps i .. i + 3
                // (`done`: the dynamics warp has left its loop and the aux warp may have stopped; this iteration only breaks)
                while (yk_ldu(sb + YK_OFF(prod)) < i + 4u && !yk_ldu(sb + YK_OFF(done))) { }
                asm volatile("fence.acq_rel.cta;" ::: "memory");
            }
            const float theta = yk_ldf(thr_addr + 64u * (i & (YK_RING - 1)));
            named_bar_sync(xk_bar_rows(i), YK_SYNC);
            const float4 ov = yk_ld4(row);
            const float4 tv = yk_ld4(row + 16);
            const float mine = yk_ldf(row + mine_off);
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
            if (sub == 0) yk_stu(act_addr, action);
            row = row_base + (((i + 1u) & 1u) << 10) + (action << 9);  // the row the dynamics warp prepared for this action
            __syncwarp();
            named_bar_arrive(xk_bar_act(i), YK_SYNC);
            // ---- off the chain: the step record ----
            const float v4 = sub == 6 ? 1.0f : mine;  // cartpole.rs:140
            const uint8_t v1 = sub == 5 ? (uint8_t)action : sc_prev;
            const bool on = sub == 7 ? prev_active : active;
            if (col4 && on) *reinterpret_cast<float *>(ptr) = v4;
            if (col1 && on) *ptr = v1;
            ptr += stride;
            prev_active = active;
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        if (is_aux) {
            // ------------------------------ aux warp ------------------------------
            uint32_t k0 = 2 * YK_CHUNK;
            while (k0 <= cap + 1) {
                // slots of steps k0 - 16 .. k0 - 13 are reused: their readers (thresholds at iteration k, reset states at
                // k - 1) are done once the dynamics warp is at iteration >= k0 - 12
                uint32_t c = yk_ldu(sb + YK_OFF(cons));
                bool over = false;
                while (k0 > c + YK_AHEAD) {
                    if (yk_ldu(sb + YK_OFF(done))) { over = true; break; }
                    __nanosleep(64);
                    c = yk_ldu(sb + YK_OFF(cons));
                }
                if (over) break;
                aux_fill(k0);
                __threadfence_block();
                __syncwarp();
                k0 += YK_CHUNK;
                if (lane == 0) yk_stu(sb + YK_OFF(prod), k0);
            }
        } else if (is_head) {
            // ------------------------------ head warp: lane = (env el, previous action act) ------------------------------
            const int el = lane & 15, act = lane >> 4;
            const uint32_t in0 = sb + YK_OFF(dynout) + 16u * (uint32_t)(act * YK_ENVS + el);           // parity toggles 512 B
            const uint32_t out0 = sb + YK_OFF(heads) + 32u * (uint32_t)(act * YK_ENVS + el);           // parity toggles 1024 B
            const uint32_t flag0 = sb + YK_OFF(stage) + 32u * (uint32_t)el + 16u;                      // parity toggles 1024 B
            for (uint32_t i = 0;; ++i) {
                named_bar_sync(xk_bar_rows(i), YK_SYNC);
                const double2 in = yk_ldd2(in0 + ((i & 1u) << 9));
                const uint32_t flags = __float_as_uint(yk_ld4(flag0 + ((i & 1u) << 10)).y);
                if ((flags & 2u) == 0u) break;
                // cartpole.rs:376: the next angle uses the OLD angular velocity
                const double th_next = __dadd_rn(in.x, __dmul_rn(p.time_step, in.y));
                EnvT::Head h;
                EnvT::head_of(p, th_next, h);
                const uint32_t out = out0 + ((i & 1u) << 10);
                yk_std2(out, make_double2(h.sn, h.cs));
                yk_std2(out + 16, make_double2(h.yp, h.ym));
                __syncwarp();
                named_bar_arrive(xk_bar_act(i), YK_SYNC);
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
            const uint32_t slot0 = sb + YK_OFF(slot) + 64u * (uint32_t)el, fobs0 = sb + YK_OFF(fobs) + 16u * (uint32_t)el;
            const uint32_t my_row = sb + YK_OFF(stage) + 32u * (uint32_t)(act * YK_ENVS + el);   // parity toggles 1024 B
            const uint32_t my_out = sb + YK_OFF(dynout) + 16u * (uint32_t)(act * YK_ENVS + el);  // parity toggles 512 B
            const uint32_t heads_el = sb + YK_OFF(heads) + 32u * (uint32_t)el;                   // parity 1024 B, action 512 B
            const uint32_t act_addr = sb + YK_OFF(act) + 4u * (uint32_t)el;
            {
                const double2 f0 = yk_ldd2(slot0), f1 = yk_ldd2(slot0 + 16), f2 = yk_ldd2(slot0 + 32), f3 = yk_ldd2(slot0 + 48);
                s.x = f0.x; s.xd = f0.y; s.th = f1.x; s.thd = f1.y;
                s.meta = 0x80000000u | p.max_steps;
                h.sn = f2.x; h.cs = f2.y; h.yp = f3.x; h.ym = f3.y;
            }
            bool any = __any_sync(FULL, n > 0);
            yk_st4(my_row, n > 0 ? yk_ld4(fobs0) : make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            yk_st4(my_row + 16, make_float4(n > 0 ? rem_full : 0.0f, __uint_as_float((n > 0 ? 1u : 0u) | (any ? 2u : 0u)),
                                            __uint_as_float((uint32_t)RL_PAD), 0.0f));
            yk_std2(my_out, make_double2(s.th, s.thd));
            __syncwarp();
            named_bar_arrive(xk_bar_rows(0), YK_SYNC);
            uint32_t it = 0;       // loop counter (= step index of the envs still active)
            uint32_t a_last = 0;   // the action of step it - 1 (the rows of step 0 are the same for both)
            while (any) {
                const bool active = n > 0;
                if ((it & 3u) == 0u && lane == 0) yk_stu(sb + YK_OFF(cons), it);
                if ((it & 3u) == 3u) {  // the reset states of steps it + 1 .. it + 4
                    while (yk_ldu(sb + YK_OFF(prod)) < it + 5u) { }
                    asm volatile("fence.acq_rel.cta;" ::: "memory");
                }
                // ---- before the action is known: this lane's candidate step, the would-be reset state and the complete
                //      mailbox row this lane publishes if its action is the sampled one ----
                EnvT::State cand = s;
                const int cand_sc = EnvT::step_with_head(p, cand, h, y_ml, (uint32_t)act);
                const uint32_t ring = (it + 1u) & (YK_RING - 1);
                const uint32_t sa = slot0 + 1024u * ring;
                const double2 f0 = yk_ldd2(sa), f1 = yk_ldd2(sa + 16), f2 = yk_ldd2(sa + 32), f3 = yk_ldd2(sa + 48);
                const float4 ff = yk_ld4(fobs0 + 256u * ring);
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
                const uint32_t par = ((it + 1u) & 1u);
                yk_st4(my_row + (par << 10), row0);
                yk_st4(my_row + (par << 10) + 16, make_float4(rem_a, __uint_as_float((n_a > 0 ? 1u : 0u) | (cons_any ? 2u : 0u)), sc_f, 0.0f));
                yk_std2(my_out + (par << 9), make_double2(ended_a ? f1.x : cand.th, ended_a ? f1.y : cand.thd));
                __syncwarp();
                named_bar_arrive(xk_bar_rows(it + 1u), YK_SYNC);
                // ---- the action, and the Head the head warp computed for the state this step started from ----
                const uint32_t head_addr = heads_el + ((it & 1u) << 10) + (a_last << 9);
                named_bar_sync(xk_bar_act(it), YK_SYNC);
                const uint32_t action = yk_ldu(act_addr);
                const double2 hn0 = yk_ldd2(head_addr), hn1 = yk_ldd2(head_addr + 16);
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
                h.sn = ended ? f2.x : hn0.x; h.cs = ended ? f2.y : hn0.y;
                h.yp = ended ? f3.x : hn1.x; h.ym = ended ? f3.y : hn1.y;
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
                    // the rows of step it + 1 said "may take the step": the policy and head warps evaluate it once more
                    // (every env inactive) and need a terminator to leave on
                    yk_st4(my_row + ((it & 1u) << 10) + 16, make_float4(0.0f, __uint_as_float(0u), __uint_as_float((uint32_t)RL_PAD), 0.0f));
                    __syncwarp();
                    named_bar_arrive(xk_bar_rows(it + 2u), YK_SYNC);
                    named_bar_sync(xk_bar_act(it + 1u), YK_SYNC);
                }
                a_last = action;
                any = any_next;
                it += 1;
            }
            if (lane == 0) yk_stu(sb + YK_OFF(done), 1u);
            st.v[ST_STEPS] = st.v[ST_R] = st.v[ST_R2] = (double)i;
            st.v[ST_EPS] = n_eps; st.v[ST_ER] = st.v[ST_EL] = sum_el; st.v[ST_ER2] = st.v[ST_EL2] = sum_el2;
            st.cur_len = i;  // carried over the barrier below
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
