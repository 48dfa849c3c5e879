// gru_big_tc.cuh -- the per-step GEMMs of K10 (gru_big.cu) on the 5th-generation tensor cores, f32-accurate through the
// exact three-piece bf16 split of pass_tc.cuh.  Included by gru_big.cu inside its anonymous namespace; hidden == 128.
//
//   out[n][e] = sum_k W[n][k] In[k][e]      n = weight rows (N = 512 gate rows or 128), e = lanes, K <= 512
//
// A CTA owns 128 lanes: one lane per thread and per TMEM lane (M = 128), so the epilogue -- the gate nonlinearities, the
// forward tangent, or the carry update -- runs on the accumulators of its own lane and every global store is a coalesced
// 128-byte row segment.  Per 16 values of k:
//   A operand  the lanes' inputs [128 x 16], cut by the threads into hi + mid + lo bf16 pieces (exact) and stored in the
//              canonical no-swizzle K-major layout ([8-element chunk][row][16 B], one 16-byte store per thread and chunk);
//   B operand  the weight pieces, precomputed once per pass in the same layout by big_pieces_kernel and streamed with one
//              cp.async.bulk per step (48 KB for 512 rows) through a three-stage ring with mbarrier complete_tx;
//   MMAs       the six piece products that matter to f32 accuracy  hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi  (dropped:
//              <= 2^-24 relative each; bf16 products are exact in the f32 accumulator), each as tcgen05.mma 128 x 256 x 16
//              (two N halves for 512 rows) accumulating in TMEM: 512 columns = the whole tensor memory of the SM.
// The threads build the A pieces of the next step while the tensor core works on the current one.
#pragma once

#include "tcgen05.cuh"

namespace bt {

using namespace tc;

constexpr int BT_LANES = 128, BT_PRODUCERS = 256, BT_THREADS = 288;
// ring stages: 512 weight rows -> three stages of 60 KB; 128 rows -> eight stages of 24 KB (the weight pieces of a step are
// copied BT_STAGES - 1 steps ahead: one copy in flight at a time left every step waiting ~1800 clk for its 12 .. 48 KB)
__host__ __device__ constexpr int bt_stages(int N) { return N > 256 ? 3 : 8; }
constexpr int BT_A_STAGE = 3 * 2 * 2048;  // three pieces x two 8-element chunks x 128 rows x 16 B

__host__ __device__ constexpr int bt_b_stage(int N) { return 3 * 2 * N * 16; }
__host__ __device__ constexpr int bt_smem(int N) { return bt_stages(N) * (BT_A_STAGE + bt_b_stage(N)) + 512 + N * 4; }

enum { EPI_GATES = 0, EPI_TAN = 1, EPI_ACC = 2, EPI_CELL = 3 };  // EPI_CELL: gates, only h' is stored (rollouts)

struct BtArgs {
    const float *src0, *src1, *src2;  // A rows k < k0 from src0, k0 <= k < k1 from src1, k1 <= k < K from src2 (row stride E)
    int k0, k1, K, nsteps;            // nsteps = ceil(K / 16)
    const uint16_t *wp;               // weight pieces [nsteps][3][2][N][8] bf16
    uint64_t E;
    const float *bias;                // [N] or null
    const uint8_t *succ_t;
    // EPI_GATES: hprev = src1; outputs
    float *R, *U, *Nn, *HN, *HNEW, *hnext;
    // EPI_TAN: reads R, U, Nn, HN, hprev = src1, thp = src2 (in place), writes THNEW
    float *THNEW, *thp;
    // EPI_ACC: out[n][e] += acc
    float *out;
    const int *skip_flag;
    long long *dbg;  // optional (RL_SEQ_TC_DEBUG): clock64 deltas setup / steps / epilogue of block 0
};

// weight pieces of W [N x K] (row-major, lda), optionally followed by a second matrix W2 [N x K2] along k
__global__ void big_pieces_kernel(const float *__restrict__ W, int lda, int K, const float *__restrict__ W2, int lda2, int K2, int N,
                                  int nsteps, uint16_t *__restrict__ wp) {
    const size_t total = (size_t)nsteps * 3 * 2 * N * 8;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx & 7);
        size_t q = idx >> 3;
        const int n = (int)(q % N); q /= N;
        const int kc = (int)(q & 1); q >>= 1;
        const int p = (int)(q % 3);
        const int s = (int)(q / 3);
        const int k = 16 * s + 8 * kc + i;
        float v = 0.0f;
        if (k < K) v = W[(size_t)n * lda + k];
        else if (k < K + K2) v = W2[(size_t)n * lda2 + (k - K)];
        uint32_t hi, mid, lo;
        split3(v, hi, mid, lo);
        wp[idx] = (uint16_t)((p == 0 ? hi : p == 1 ? mid : lo) >> 16);
    }
}

__device__ __forceinline__ void bt_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bt_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// gate nonlinearities on the SFU (the forms of gru_tile.cuh): MUFU.EX2 + MUFU.RCP, ~2 ulp each; tanh as 2 sigmoid(2 x) - 1
// (absolute error ~2e-7: n enters h' additively)
__device__ __forceinline__ float bt_sigm(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float bt_tanh(float v) { return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * v)), -1.0f); }

template <int N, int EPI>
__global__ void __launch_bounds__(BT_THREADS, 1) big_gemm_tc_kernel(BtArgs a) {
    constexpr int B_STAGE = bt_b_stage(N), STAGE = BT_A_STAGE + B_STAGE, NH = N > 256 ? 2 : 1, NM = N > 256 ? 256 : N;
    // weight pieces travel BT_DIST steps ahead, into the stage step s + BT_DIST - BT_STAGES used: never wait for a commit younger than that
    constexpr int BT_STAGES = bt_stages(N), BT_DIST = N > 256 ? 2 : 4;
    constexpr uint32_t IDESC = make_idesc(128, NM, false, false);
    if (a.skip_flag && *a.skip_flag) return;
    extern __shared__ __align__(1024) unsigned char bsm[];
    // 256 producer threads (lane, half) + one issuer warp.  Staging: `half` is the 8-element chunk of the step a producer cuts
    // into pieces; epilogue: `half` selects the half of the units / columns.  The producers and the issuer meet only
    // through mbarriers (a_full: 256 arrivals per stage; b_full: the bulk copy's bytes; done: tcgen05.commit), so the
    // producers run up to three stages ahead of the tensor core.  (First version: 128 threads and a __syncthreads per step
    // -- every dependent instruction of the staging code stalled its whole sub-partition and thread 0's MMA issue sat on
    // everybody's path: ~3000 clk per step against ~1600 clk of MMAs.)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & (BT_LANES - 1), half = (tid >> 7) & 1;
    const bool is_issuer_warp = tid >= BT_PRODUCERS;
    const long long c_start = clock64();
    uint64_t *bars = reinterpret_cast<uint64_t *>(bsm + BT_STAGES * STAGE);  // b_full[3], done[3], a_full[3]
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bars + 3 * BT_STAGES);
    float *sbias = reinterpret_cast<float *>(bsm + BT_STAGES * STAGE + 512);  // [N] (epilogues with a bias)
    if (EPI != EPI_ACC)
        for (int i = tid; i < N; i += BT_THREADS) sbias[i] = a.bias[i];
    const uint32_t bar_full = smem_u32(bars), bar_done = smem_u32(bars + BT_STAGES), bar_afull = smem_u32(bars + 2 * BT_STAGES);
    if (warp == 0) {
        tmem_alloc(smem_u32(tptr), N < 32 ? 32 : N);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < BT_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_done + 8 * s, 1);
            mbar_init(bar_afull + 8 * s, BT_PRODUCERS / 32);  // one arrival per producer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tptr[0];
    const long long c_setup = clock64();
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;  // a warp reads the TMEM lanes of its position in its warpgroup
    const uint64_t e = (uint64_t)blockIdx.x * BT_LANES + lane, E = a.E;
    const bool in_range = e < E;

    // this thread's 8 inputs of a step (raw loads only: nothing here waits for the data).  Row k of the input lives at
    // base(k) + k E + e with base = src0, src1 - k0 E or src2 - k1 E: one running offset, two selects per row (the first
    // version recomputed a three-way 64-bit address per load: half of the kernel's instructions)
    const float *const b0 = a.src0, *const b1 = a.src1 - (uint64_t)a.k0 * E, *const b2 = a.src2 - (uint64_t)a.k1 * E;
    const int k0 = a.k0, k1 = a.k1, K = in_range ? a.K : 0;
    auto load_step = [&](int s, float *v) {
        const int kb = 16 * s + 8 * half;
        uint64_t off = (uint64_t)kb * E + e;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = kb + i;
            const float *base = k < k0 ? b0 : k < k1 ? b1 : b2;
            v[i] = k < K ? __ldg(base + off) : 0.0f;
            off += E;
        }
    };
    auto issue_b = [&](int s) {  // the weight pieces of step s into its stage (thread 0)
        const int st = s % BT_STAGES;
        bt_expect_tx(bar_full + 8 * st, (uint32_t)B_STAGE);
        bt_bulk_g2s(smem_u32(bsm + st * STAGE + BT_A_STAGE), a.wp + (size_t)s * (B_STAGE / 2), (uint32_t)B_STAGE, bar_full + 8 * st);
    };
    // The inputs are streamed from HBM: the loads run BT_AHEAD steps ahead of their use (Little's law).
    constexpr int BT_AHEAD = 4;
    if (is_issuer_warp) {
        // ------------------------------ issuer: weight-piece copies and MMAs (one thread) ------------------------------
        if (tid == BT_PRODUCERS) {
            // descriptors of every (stage, piece, N half) once: inside the loop an MMA costs its own issue only (building two
            // 64-bit descriptors per MMA in the issuing thread was ~150 clk per MMA, as long as the MMA itself runs)
            uint64_t dA[BT_STAGES][3], dB[BT_STAGES][3][NH];
#pragma unroll
            for (int st = 0; st < BT_STAGES; ++st) {
                const uint32_t aA = smem_u32(bsm + st * STAGE), aB = aA + BT_A_STAGE;
#pragma unroll
                for (int pc = 0; pc < 3; ++pc) {
                    dA[st][pc] = make_desc(aA + pc * 4096, 2048, 128);
#pragma unroll
                    for (int h = 0; h < NH; ++h) dB[st][pc][h] = make_desc(aB + pc * (2 * N * 16) + h * (256 * 16), N * 16, 128);
                }
            }
            for (int s = 0; s < BT_DIST && s < a.nsteps; ++s) issue_b(s);
            for (int s0 = 0; s0 < a.nsteps; s0 += BT_STAGES) {
#pragma unroll
                for (int st = 0; st < BT_STAGES; ++st) {
                    const int s = s0 + st;
                    if (s >= a.nsteps) break;
                    const uint32_t par = (uint32_t)((s / BT_STAGES) & 1);
                    mbar_wait(bar_afull + 8 * st, par);
                    mbar_wait(bar_full + 8 * st, par);
                    fence_after();
                    // piece pairs (A piece, B piece): hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi
                    constexpr int PA[6] = {0, 0, 1, 1, 0, 2}, PB[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int q = 0; q < 6; ++q)
                            umma_bf16(tmem_d + (uint32_t)(h * 256), dA[st][PA[q]], dB[st][PB[q]][h], IDESC, (s > 0 || q > 0) ? 1u : 0u);
                    umma_commit(bar_done + 8 * st);
                    // the weight pieces of step s + BT_DIST go into the stage step s - 1 used: free when its MMAs are done (the
                    // MMAs of step s are queued behind them, so the tensor core stays busy while this thread waits)
                    const int sn = s + BT_DIST, sold = sn - BT_STAGES;  // the step that used sn's stage before
                    if (sn < a.nsteps) {
                        if (sold >= 0) mbar_wait(bar_done + 8 * (sn % BT_STAGES), (uint32_t)((sold / BT_STAGES) & 1));
                        issue_b(sn);
                    }
                }
            }
        }
    } else {
        // ------------------------------ producers: this lane's chunk `half` of every step ------------------------------
        float vb[BT_AHEAD][8];
#pragma unroll
        for (int u = 0; u < BT_AHEAD; ++u)
            if (u < a.nsteps) load_step(u, vb[u]);
        for (int s0 = 0; s0 < a.nsteps; s0 += BT_AHEAD) {
#pragma unroll
            for (int u = 0; u < BT_AHEAD; ++u) {
                const int s = s0 + u;
                if (s >= a.nsteps) break;
                const int st = s % BT_STAGES;
                unsigned char *sA = bsm + st * STAGE;
                uint32_t hi[8], mid[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) split3(vb[u][i], hi[i], mid[i], lo[i]);
                if (s + BT_AHEAD < a.nsteps) load_step(s + BT_AHEAD, vb[u]);
                if (s >= BT_STAGES) mbar_wait(bar_done + 8 * st, (uint32_t)((s / BT_STAGES - 1) & 1));  // MMAs of step s - 3 have read the stage
                const int o = half * 2048 + lane * 16;
                *reinterpret_cast<uint4 *>(sA + o) = make_uint4(pack_hi16(hi[0], hi[1]), pack_hi16(hi[2], hi[3]), pack_hi16(hi[4], hi[5]), pack_hi16(hi[6], hi[7]));
                *reinterpret_cast<uint4 *>(sA + 4096 + o) = make_uint4(pack_hi16(mid[0], mid[1]), pack_hi16(mid[2], mid[3]), pack_hi16(mid[4], mid[5]), pack_hi16(mid[6], mid[7]));
                *reinterpret_cast<uint4 *>(sA + 8192 + o) = make_uint4(pack_hi16(lo[0], lo[1]), pack_hi16(lo[2], lo[3]), pack_hi16(lo[4], lo[5]), pack_hi16(lo[6], lo[7]));
                fence_async_smem();  // this thread's generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_afull + 8 * st) : "memory");
            }
        }
    }
    // ---- all MMAs done: the last commit covers every earlier one ----
    if (!is_issuer_warp) {
        const int s = a.nsteps - 1, st = s % BT_STAGES;
        mbar_wait(bar_done + 8 * st, (uint32_t)((s / BT_STAGES) & 1));
        fence_after();
    }
    const long long c_mma = clock64();
    constexpr int H = 128;
    if (is_issuer_warp) {
        // (the issuer warp only joins the closing barrier)
    } else
    // Epilogue.  Loads first, stores after, through __restrict__ pointers: written element by element the compiler kept every
    // load behind the previous element's store (possible aliasing) and the epilogue cost 60 k clk of serialised latency.
    if (EPI == EPI_ACC) {
        float *__restrict__ out = a.out;
        constexpr int CPH = N / 2;  // columns of this half
#pragma unroll 1
        for (int c = 0; c < CPH / 32; ++c) {
            const int col0 = half * CPH + c * 32;
            uint32_t r[32];
            tmem_ld32(tmem_d + lane_off + col0, r);
            if (in_range) {
                float old[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) old[j] = out[(uint64_t)(col0 + j) * E + e];
#pragma unroll
                for (int j = 0; j < 32; ++j) out[(uint64_t)(col0 + j) * E + e] = old[j] + __uint_as_float(r[j]);
            }
        }
    } else {
        const uint8_t sc = (in_range && a.succ_t) ? a.succ_t[e] : (uint8_t)RL_CONTINUE;
        const float *bias = sbias;
        const float *__restrict__ hprev = a.src1;
        float *__restrict__ R = a.R, *__restrict__ U = a.U, *__restrict__ Nn = a.Nn, *__restrict__ HN = a.HN;
        float *__restrict__ HNEW = a.HNEW, *__restrict__ hnext = a.hnext, *__restrict__ THNEW = a.THNEW, *__restrict__ thp = a.thp;
#pragma unroll 1
        for (int c = 0; c < H / 32; ++c) {  // this half's 64 units, 16 at a time
            const int j0 = half * (H / 2) + c * 16;
            uint32_t g0[16], g1[16], g2[16], g3[16];
            tmem_ld16(tmem_d + lane_off + j0, g0);
            tmem_ld16(tmem_d + lane_off + H + j0, g1);
            tmem_ld16(tmem_d + lane_off + 2 * H + j0, g2);
            tmem_ld16(tmem_d + lane_off + 3 * H + j0, g3);
            if (!in_range) continue;
            float hp[16], aux0[16], aux1[16], aux2[16], aux3[16], aux4[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const uint64_t i = (uint64_t)(j0 + jj) * E + e;
                hp[jj] = hprev[i];
                if (EPI == EPI_TAN) { aux0[jj] = R[i]; aux1[jj] = U[i]; aux2[jj] = Nn[i]; aux3[jj] = HN[i]; aux4[jj] = thp[i]; }
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = j0 + jj;
                const uint64_t i = (uint64_t)j * E + e;
                const float p0 = __uint_as_float(g0[jj]) + bias[j], p1 = __uint_as_float(g1[jj]) + bias[H + j];
                const float p2 = __uint_as_float(g2[jj]) + bias[2 * H + j], p3 = __uint_as_float(g3[jj]) + bias[3 * H + j];
                if (EPI == EPI_GATES || EPI == EPI_CELL) {
                    const float r = bt_sigm(p0), u = bt_sigm(p1), hn = p2, in = p3;
                    const float n = bt_tanh(__fadd_rn(in, __fmul_rn(hn, r)));
                    const float hnew = __fadd_rn(__fmul_rn(__fsub_rn(hp[jj], n), u), n);
                    HNEW[i] = hnew;
                    if (EPI == EPI_GATES) {
                        R[i] = r; U[i] = u; Nn[i] = n; HN[i] = hn;
                        if (hnext) hnext[i] = sc != RL_CONTINUE ? 0.0f : hnew;
                    }
                } else {  // EPI_TAN: p0 .. p3 = tangents of the r, u, hn, in pre-activations
                    const float r = aux0[jj], u = aux1[jj], n = aux2[jj], hn = aux3[jj];
                    const float rd = r * (1.0f - r) * p0;
                    const float ud = u * (1.0f - u) * p1;
                    const float nd = (1.0f - n * n) * (p3 + rd * hn + r * p2);
                    const float thnew = ud * (hp[jj] - n) + u * aux4[jj] + (1.0f - u) * nd;
                    THNEW[i] = thnew;
                    thp[i] = sc != RL_CONTINUE ? 0.0f : thnew;
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (a.dbg && blockIdx.x == 0 && tid == 0) {
        a.dbg[0] = c_setup - c_start; a.dbg[1] = c_mma - c_setup; a.dbg[2] = clock64() - c_mma;
    }
    if (warp == 0) {
        fence_after();
        tmem_dealloc(tmem_d, N < 32 ? 32 : N);
    }
}

}  // namespace bt

// ---------------------------------------------------------------------------------------------------------------------
// The weight gradients on the tensor cores:  dWc [512 x 144] = sum over all (t, e) of D_t [512 x E] . In_t [144 x E]^T with
// In_t = [x_t ; hprev_t ; 1 ; 0] -- the contraction runs over the LANES, which are contiguous in both operands (K-major).
// A CTA owns half of the rows (two 128-row M tiles, accumulators D [128 x 144] each in TMEM) and one split of the (t, e)
// range; a step = 16 lanes of one time step.  Both operands are activations: the producers cut 256 + 144 rows x 16 lanes
// into bf16 pieces per step (item = one row's 8 lanes: two LDG.128, eight split3, three STS.128).  The f32 accumulators
// are drained every NT_DRAIN steps into f32 slabs that splitk_reduce_f32_kernel adds in f64 in a fixed order.
// ---------------------------------------------------------------------------------------------------------------------
namespace bt {

constexpr int NT_ROWS_A = 256, NT_ROWS_B = 144, NT_STAGES = 4, NT_DRAIN = 512;
constexpr int NT_A_PIECE = 2 * NT_ROWS_A * 16, NT_B_PIECE = 2 * NT_ROWS_B * 16;  // bytes of one piece of one step
constexpr int NT_STAGE = 3 * (NT_A_PIECE + NT_B_PIECE);
__host__ __device__ constexpr int nt_smem() { return NT_STAGES * NT_STAGE + 512; }

struct NtArgs {
    const float *D;           // [T][512][E]
    const float *src0, *src1; // In rows: [0, rows0) from src0 [T][rows0][E], [rows0, rows0 + rows1) from src1, then a row of ones
    int rows0, rows1;
    uint64_t T, E;
    float *part;              // [splits][nd][2][256][144] f32 slabs (zero-initialised)
    int splits, nd;
    const int *skip_flag;
};

__global__ void __launch_bounds__(BT_THREADS, 1) big_nt_tc_kernel(NtArgs a) {
    constexpr uint32_t IDESC = make_idesc(128, NT_ROWS_B, false, false);
    constexpr int ITEMS = 2 * (NT_ROWS_A + NT_ROWS_B), IPT = (ITEMS + BT_PRODUCERS - 1) / BT_PRODUCERS;  // 800 items, 4 per thread
    if (a.skip_flag && *a.skip_flag) return;
    extern __shared__ __align__(1024) unsigned char bsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & (BT_LANES - 1), half = (tid >> 7) & 1;
    const bool is_issuer_warp = tid >= BT_PRODUCERS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bsm + NT_STAGES * NT_STAGE);  // a_full[4], done[4], acc_free
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bars + 2 * NT_STAGES + 1);
    const uint32_t bar_afull = smem_u32(bars), bar_done = smem_u32(bars + NT_STAGES), bar_free = smem_u32(bars + 2 * NT_STAGES);
    if (warp == 0) {
        tmem_alloc(smem_u32(tptr), 512);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < NT_STAGES; ++s) {
            mbar_init(bar_afull + 8 * s, BT_PRODUCERS / 32);
            mbar_init(bar_done + 8 * s, 1);
        }
        mbar_init(bar_free, BT_PRODUCERS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tptr[0];
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int mhalf = blockIdx.y, split = blockIdx.x;
    const uint64_t E = a.E, upt = (E + 15) / 16, units = a.T * upt;
    const uint64_t u_begin = units * (uint64_t)split / (uint64_t)a.splits, u_end = units * (uint64_t)(split + 1) / (uint64_t)a.splits;
    const int nsteps = (int)(u_end - u_begin);
    const int NIN = a.rows0 + a.rows1;  // index of the ones row

    if (is_issuer_warp) {
        if (tid == BT_PRODUCERS) {
            uint64_t dA[NT_STAGES][3][2], dB[NT_STAGES][3];
#pragma unroll
            for (int st = 0; st < NT_STAGES; ++st) {
                const uint32_t base = smem_u32(bsm + st * NT_STAGE);
#pragma unroll
                for (int pc = 0; pc < 3; ++pc) {
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) dA[st][pc][mt] = make_desc(base + pc * NT_A_PIECE + mt * (128 * 16), NT_ROWS_A * 16, 128);
                    dB[st][pc] = make_desc(base + 3 * NT_A_PIECE + pc * NT_B_PIECE, NT_ROWS_B * 16, 128);
                }
            }
            int since = 0, ndrain = 0;
            for (int s0 = 0; s0 < nsteps; s0 += NT_STAGES) {
#pragma unroll
                for (int st = 0; st < NT_STAGES; ++st) {
                    const int s = s0 + st;
                    if (s >= nsteps) break;
                    if (since == 0 && ndrain > 0) mbar_wait(bar_free, (uint32_t)((ndrain - 1) & 1));  // the producers have read the accumulators
                    mbar_wait(bar_afull + 8 * st, (uint32_t)((s / NT_STAGES) & 1));
                    fence_after();
                    constexpr int PA[6] = {0, 0, 1, 1, 0, 2}, PB[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int q = 0; q < 6; ++q)
                            umma_bf16(tmem_d + (uint32_t)(mt * 256), dA[st][PA[q]][mt], dB[st][PB[q]], IDESC, (since > 0 || q > 0) ? 1u : 0u);
                    umma_commit(bar_done + 8 * st);
                    since += 1;
                    if (since == NT_DRAIN || s == nsteps - 1) { since = 0; ndrain += 1; }
                }
            }
        }
    } else {
        // ------------------------------ producers ------------------------------
        const bool vec_ok = (E & 3) == 0;
        // item it of a step: row r = it >> 1 of the 400-row operand stack (256 rows of D, then 144 rows of In), chunk c = it & 1
        auto load_item = [&](uint64_t u, int it, float *v) {
            const uint64_t t = u / upt, e0 = (u - t * upt) * 16 + (uint64_t)(it & 1) * 8;
            const int r = it >> 1;
            const float *src = nullptr;
            float fill = 0.0f;
            if (r < NT_ROWS_A) src = a.D + (t * 512 + (uint64_t)(mhalf * NT_ROWS_A + r)) * E;
            else {
                const int n = r - NT_ROWS_A;
                if (n < a.rows0) src = a.src0 + (t * a.rows0 + n) * E;
                else if (n < NIN) src = a.src1 + (t * a.rows1 + (n - a.rows0)) * E;
                else if (n == NIN) fill = 1.0f;
            }
            if (src && vec_ok && e0 + 8 <= E) {
                const float4 x0 = __ldg(reinterpret_cast<const float4 *>(src + e0)), x1 = __ldg(reinterpret_cast<const float4 *>(src + e0 + 4));
                v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (e0 + i < E) ? (src ? __ldg(src + e0 + i) : fill) : 0.0f;
            }
        };
        auto store_item = [&](unsigned char *stage, int it, const float *v) {
            const int r = it >> 1, c = it & 1;
            uint32_t hi[8], mid[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split3(v[i], hi[i], mid[i], lo[i]);
            unsigned char *base;
            int piece_bytes, o;
            if (r < NT_ROWS_A) { base = stage; piece_bytes = NT_A_PIECE; o = c * (NT_ROWS_A * 16) + r * 16; }
            else { base = stage + 3 * NT_A_PIECE; piece_bytes = NT_B_PIECE; o = c * (NT_ROWS_B * 16) + (r - NT_ROWS_A) * 16; }
            *reinterpret_cast<uint4 *>(base + o) = make_uint4(pack_hi16(hi[0], hi[1]), pack_hi16(hi[2], hi[3]), pack_hi16(hi[4], hi[5]), pack_hi16(hi[6], hi[7]));
            *reinterpret_cast<uint4 *>(base + piece_bytes + o) = make_uint4(pack_hi16(mid[0], mid[1]), pack_hi16(mid[2], mid[3]), pack_hi16(mid[4], mid[5]), pack_hi16(mid[6], mid[7]));
            *reinterpret_cast<uint4 *>(base + 2 * piece_bytes + o) = make_uint4(pack_hi16(lo[0], lo[1]), pack_hi16(lo[2], lo[3]), pack_hi16(lo[4], lo[5]), pack_hi16(lo[6], lo[7]));
        };
        constexpr int AHEAD = 2;
        float vb[AHEAD][IPT][8];
#pragma unroll
        for (int d = 0; d < AHEAD; ++d)
            if (d < nsteps) {
#pragma unroll
                for (int q = 0; q < IPT; ++q)
                    if (tid + q * BT_PRODUCERS < ITEMS) load_item(u_begin + d, tid + q * BT_PRODUCERS, vb[d][q]);
            }
        int since = 0, ndrain = 0;
        for (int s0 = 0; s0 < nsteps; s0 += AHEAD) {
#pragma unroll
            for (int d = 0; d < AHEAD; ++d) {
                const int s = s0 + d;
                if (s >= nsteps) break;
                const int st = s % NT_STAGES;
                unsigned char *stage = bsm + st * NT_STAGE;
                if (s >= NT_STAGES) mbar_wait(bar_done + 8 * st, (uint32_t)((s / NT_STAGES - 1) & 1));  // MMAs of step s - 4 have read the stage
#pragma unroll
                for (int q = 0; q < IPT; ++q)
                    if (tid + q * BT_PRODUCERS < ITEMS) {
                        store_item(stage, tid + q * BT_PRODUCERS, vb[d][q]);
                        if (s + AHEAD < nsteps) load_item(u_begin + s + AHEAD, tid + q * BT_PRODUCERS, vb[d][q]);
                    }
                fence_async_smem();
                __syncwarp();
                if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_afull + 8 * st) : "memory");
                since += 1;
                if (since == NT_DRAIN || s == nsteps - 1) {
                    // ---- drain: wait for every MMA so far, move this thread's 2 x 72 accumulator columns into the slab ----
                    mbar_wait(bar_done + 8 * st, (uint32_t)((s / NT_STAGES) & 1));
                    fence_after();
                    float *slab = a.part + (((size_t)split * a.nd + ndrain) * 2 + mhalf) * (size_t)(NT_ROWS_A * NT_ROWS_B);
#pragma unroll 1
                    for (int mt = 0; mt < 2; ++mt) {
                        float *row = slab + (size_t)(mt * 128 + lane) * NT_ROWS_B + half * 72;
#pragma unroll 1
                        for (int c = 0; c < 9; ++c) {
                            uint32_t r8[8];
                            tmem_ld8(tmem_d + lane_off + (uint32_t)(mt * 256 + half * 72 + c * 8), r8);
                            *reinterpret_cast<float4 *>(row + c * 8) = make_float4(__uint_as_float(r8[0]), __uint_as_float(r8[1]), __uint_as_float(r8[2]), __uint_as_float(r8[3]));
                            *reinterpret_cast<float4 *>(row + c * 8 + 4) = make_float4(__uint_as_float(r8[4]), __uint_as_float(r8[5]), __uint_as_float(r8[6]), __uint_as_float(r8[7]));
                        }
                    }
                    fence_before();
                    __syncwarp();
                    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_free) : "memory");
                    since = 0;
                    ndrain += 1;
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        tmem_dealloc(tmem_d, 512);
    }
}

// sums[m][n] (f64, row stride NB) = sum over the slabs of part[slab][m][n] (f32, row stride 144) in slab order
__global__ void splitk_reduce_f32_kernel(const float *__restrict__ part, int nslabs, int NB, double *__restrict__ sums, const int *skip_flag) {
    if (skip_flag && *skip_flag) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 512 x 144
    if (i >= 512 * NT_ROWS_B) return;
    const int m = i / NT_ROWS_B, n = i - m * NT_ROWS_B;
    if (n >= NB) return;
    // slab layout: [slab][mhalf][256][144]
    const int mh = m >> 8, mr = m & 255;
    double s = 0.0;
    for (int k = 0; k < nslabs; ++k) s += (double)part[(((size_t)k * 2 + mh) * NT_ROWS_A + mr) * NT_ROWS_B + n];
    sums[(size_t)m * NB + n] = s;
}

}  // namespace bt
