// pass_tail.cuh -- the tail of a full-batch pass: the reduction of the CTAs' partial rows, the cross-GPU sum and the Adam
// step, run by the LAST CTAs to finish instead of by a second launch (80 reduce_rows_adam_kernel
// launches were 1.06 ms of the critic's 4.75 ms; with N > 1 GPUs the peer exchange was a second launch after every pass).
//
// Deterministic two-level "last block done" reduction (no floating-point atomics):
//   level 1  the CTAs are grouped by PT_GROUP consecutive block indices; the CTA that takes the last ticket of its group
//            sums the group's rows in row order into group_rows[g];
//   level 2  the CTA that takes the last ticket over the groups sums the group rows in group order  ->  this rank's sums.
// Then, in that one CTA: with N > 1 GPUs the sums are PUSHED into every peer's mailbox over NVLink and the peers' sums are
// added in rank order (the protocol of reduce_rows_x_kernel, update.cu: same slots, flags, parity by sequence number, so the
// two kernels can alternate within one update); with Adam requested libtorch's Adam::step is applied to the parameters.
// The order of every sum is fixed by indices, so the result is the same on every run and, with N > 1, on every rank.
// Tickets reset themselves (the last CTA zeroes them), so launches need no memset in between.
#pragma once

constexpr int PT_GROUP = 16;
constexpr int PT_MAX_GROUPS = 64;
enum { PT_NONE = 0, PT_REDUCE = 1, PT_ADAM = 2 };

struct PassTail {
    int mode;                  // PT_NONE: the kernel only writes its partial row (a later launch reduces)
    unsigned int *tickets;     // [1 + PT_MAX_GROUPS], zero between launches
    double *group_rows;        // [groups][W]
    double *sums;              // [W]
    // Adam (PT_ADAM)
    float *theta, *m, *v;
    AdamArgs c;
    unsigned long long step;
    double *loss_out;
    // data-parallel group (world > 1)
    int use_x;
    rl_xpeer x;
    unsigned long long seq;
};

__device__ __forceinline__ double pt_ld(const double *p) { return __ldcg(p); }

// Called by every thread of every CTA after the CTA's partial row is written.  `rows` = gridDim.x rows of W doubles.
// smem_flags: two ints of shared memory.  `skip`: this pass is skipped (device-side flag, identical on every rank): no
// sums and no Adam step are written, but with world > 1 the exchange still runs so that the parities stay aligned.
template <int NT>
__device__ __forceinline__ void pass_tail_run(const PassTail &t, const double *rows, int W, int P, int *smem_flags, double *smem_sc,
                                              bool skip) {
    if (t.mode == PT_NONE) return;
    const int tid = threadIdx.x, nblk = gridDim.x, g = blockIdx.x / PT_GROUP, ngroups = (nblk + PT_GROUP - 1) / PT_GROUP;
    const int r0 = g * PT_GROUP, r1 = min(r0 + PT_GROUP, nblk);
    __threadfence();
    __syncthreads();
    if (tid == 0) smem_flags[0] = atomicAdd(&t.tickets[1 + g], 1u) == (unsigned)(r1 - r0 - 1);
    __syncthreads();
    if (!smem_flags[0]) return;
    __threadfence();
    // (loads first, sums after: a thread keeps 2 x PT_GROUP independent L2 loads in flight instead of one dependent chain)
    for (int c0 = tid; c0 < W; c0 += 2 * NT) {
        const int c1 = c0 + NT;
        double v0[PT_GROUP], v1[PT_GROUP];
#pragma unroll
        for (int r = 0; r < PT_GROUP; ++r) {
            const bool ok = r0 + r < r1;
            v0[r] = ok ? pt_ld(rows + (size_t)(r0 + r) * W + c0) : 0.0;
            v1[r] = (ok && c1 < W) ? pt_ld(rows + (size_t)(r0 + r) * W + c1) : 0.0;
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < PT_GROUP; ++r) {  // row order; an absent row adds +0.0
            s0 += v0[r];
            s1 += v1[r];
        }
        t.group_rows[(size_t)g * W + c0] = s0;
        if (c1 < W) t.group_rows[(size_t)g * W + c1] = s1;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) smem_flags[1] = atomicAdd(&t.tickets[0], 1u) == (unsigned)(ngroups - 1);
    __syncthreads();
    if (!smem_flags[1]) return;
    __threadfence();
    // ---- the last CTA: this rank's sums ----
    constexpr int CPT = 16;  // columns per thread: W <= NT * CPT
    constexpr int GCH = 16;  // group rows loaded per round and column
    double loc[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) loc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < CPT; k += 2) {
        const int c0 = tid + k * NT, c1 = c0 + NT;
        if (c0 < W) {
            for (int g0 = 0; g0 < ngroups; g0 += GCH) {  // group order
                double v0[GCH], v1[GCH];
#pragma unroll
                for (int j = 0; j < GCH; ++j) {
                    const bool ok = g0 + j < ngroups;
                    v0[j] = ok ? pt_ld(t.group_rows + (size_t)(g0 + j) * W + c0) : 0.0;
                    v1[j] = (ok && c1 < W) ? pt_ld(t.group_rows + (size_t)(g0 + j) * W + c1) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < GCH; ++j) {
                    loc[k] += v0[j];
                    loc[k + 1] += v1[j];
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int col = tid + k * NT;
        if (col == P + SC_COUNT) smem_sc[0] = loc[k];
        if (col == P + SC_LOSS) smem_sc[1] = loc[k];
    }
    if (tid <= ngroups) t.tickets[tid] = 0u;  // (ngroups + 1 <= NT)
    __syncthreads();
    double N = smem_sc[0], lsum = smem_sc[1];
    if (t.use_x) {
        // ---- cross-GPU sum over the peer mailboxes (slots and flags of reduce_rows_x_kernel: block b = 32 columns) ----
        const rl_xpeer &x = t.x;
        const size_t par = (size_t)(t.seq & 1ull);
        const int nxb = (W + 31) / 32;
        const bool bad = *reinterpret_cast<const volatile int *>(x.error) != 0;  // sticky: a peer never arrived earlier
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = tid + k * NT;
            if (col < W) {
                const size_t slot_out = (par * x.world + x.rank) * RL_X_BLOCKS + (size_t)(col >> 5);
                for (int p = 0; p < x.world; ++p) {
                    double *dst = x.data[p] + slot_out * RL_X_SLOT;
                    dst[col & 31] = (skip || bad) ? 0.0 : loc[k];
                    if ((col & 31) == 0) {
                        dst[32] = N;
                        dst[33] = lsum;
                    }
                }
            }
        }
        __threadfence_system();
        __syncthreads();
        for (int i = tid; i < nxb * x.world; i += NT) {
            const int b = i / x.world, p = i - b * x.world;
            volatile unsigned long long *f = x.flag[p] + (par * x.world + x.rank) * RL_X_BLOCKS + b;
            *f = t.seq;
        }
        bool timed_out = false;
        for (int i = tid; i < nxb * x.world; i += NT) {
            const int b = i / x.world, src = i - b * x.world;
            const volatile unsigned long long *f = x.flag[x.rank] + (par * x.world + src) * RL_X_BLOCKS + b;
            const long long t0 = clock64();
            while (*f != t.seq) {
                if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer never arrived
                    atomicExch(x.error, 1);
                    timed_out = true;
                    break;
                }
            }
        }
        // a mailbox that never filled holds stale data: write neither the sums nor the Adam step (update.cu x_error_check)
        if (__syncthreads_or(timed_out ? 1 : 0) || bad) return;
        __threadfence_system();
        double Ntot = 0.0, ltot = 0.0;
        for (int src = 0; src < x.world; ++src) {
            const volatile double *msg = x.data[x.rank] + ((par * x.world + src) * RL_X_BLOCKS) * RL_X_SLOT;
            Ntot += msg[32];
            ltot += msg[33];
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = tid + k * NT;
            if (col < W) {
                double tot = 0.0;
                for (int src = 0; src < x.world; ++src) {
                    const volatile double *msg = x.data[x.rank] + ((par * x.world + src) * RL_X_BLOCKS + (size_t)(col >> 5)) * RL_X_SLOT;
                    tot += msg[col & 31];
                }
                loc[k] = tot;
            }
        }
        N = Ntot;
        lsum = ltot;
    }
    if (skip) return;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int col = tid + k * NT;
        if (col < W) t.sums[col] = loc[k];
    }
    if (t.mode == PT_ADAM) {
        const AdamArgs &c = t.c;
        const float beta1 = (float)c.beta1, beta2 = (float)c.beta2;
        const float omb1 = (float)(1.0 - c.beta1), omb2 = (float)(1.0 - c.beta2);
        const double bc1 = 1.0 - pow(c.beta1, (double)t.step), bc2 = 1.0 - pow(c.beta2, (double)t.step);
        const float step_size = (float)(c.lr / bc1), bc2_sqrt = (float)sqrt(bc2), eps = (float)c.eps;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int col = tid + k * NT;
            if (col < P) {
                float gr = (float)(loc[k] / N);
                const float th = t.theta[col];
                if (c.weight_decay != 0.0) gr = __fadd_rn(gr, __fmul_rn((float)c.weight_decay, th));
                // exp_avg.mul_(beta1).add_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
                const float mi = __fadd_rn(__fmul_rn(t.m[col], beta1), __fmul_rn(omb1, gr));
                const float vi = __fadd_rn(__fmul_rn(t.v[col], beta2), __fmul_rn(__fmul_rn(omb2, gr), gr));
                t.m[col] = mi;
                t.v[col] = vi;
                const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
                t.theta[col] = __fadd_rn(th, __fmul_rn(-step_size, __fdiv_rn(mi, denom)));
            }
        }
        if (tid == 0 && t.loss_out) *t.loss_out = lsum / N;
    }
}
