// pass_tc.cuh -- the critic's full-batch pass (K6) on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same contract as mlp_pass_kernel<5, 1, 4, PASS_VALUE>: one f64 partial row [P + NSCALAR] per CTA holding the
// sums over this CTA's samples of the loss, the sample count and the gradient of
// mse_loss(V(obs), targets) for the 5 -> 128 -> 1 ReLU critic (ValuesOpt::update, critics/opt.rs:100-127).
//
// Why this maps to tensor cores although K = 5: a CTA owns tiles of 128 samples, one sample per thread and
// per TMEM lane, and both batch-sized contractions become MMAs whose operands are built in shared memory:
//
//   MMA1  pre[128 samples x 128 units] = X[128 x 48] . W1e[48 x 128]            (3 x tcgen05.mma K = 16)
//         Every f32 value is cut into three bf16 pieces hi + mid + lo (truncation: exact, 8 + 8 + 8 bits), and
//         the 48 K-slots hold the six piece products that matter to f32 accuracy for the 5 features and the
//         bias:  hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi   (dropped: <= 2^-24 relative each).  bf16
//         products are exact in the f32 accumulator, so pre is an f32-accurate W1 x + b1.
//   epilogue (thread = sample): tcgen05.ld the 128 pre-activations of the sample, h = relu, V = w2 . h + b2
//         thread-local (no shuffles), loss and dV = 2 (V - target); the 0/1 ReLU mask goes back to shared
//         memory as bf16 (exact), the six values y = dV * [x, 1] as 3 x bf16 pieces.
//   MMA2  G[128 units x 18] = Mask^T[128 x 128 samples] . Y[128 x 18]             (8 x tcgen05.mma K = 16)
//         exact 0/1 times bf16 pieces, f32 accumulation over the 128 samples of a tile, then f64 across tiles.
//
// The gradient follows from G alone:  dW1[j][f] = w2_j G[j][f],  db1[j] = w2_j G[j][5],
//   dW2[j] = sum_s dV_s relu(pre_sj) = sum_s dV_s mask_sj (b1_j + w1_j . x_s) = b1_j G[j][5] + sum_f w1_jf G[j][f]
// (ReLU is piecewise linear), so no second cross-sample contraction is needed.
//
// Shared-memory operand layout: the no-swizzle canonical UMMA layout, 8 x 16 B core matrices stored as
// [chunk of 8 elements along the thread-private dimension][row = thread][16 B], so every operand store is one
// conflict-free 16 B store per thread (a warp writes 512 contiguous bytes).
#pragma once

namespace tc {

constexpr int TC_THREADS = 128;
constexpr int TC_CHUNK = 2048;  // bytes of one 8-element chunk over 128 rows
constexpr int TC_A1 = 0, TC_B1 = 6 * TC_CHUNK, TC_A2 = 12 * TC_CHUNK, TC_B2 = 28 * TC_CHUNK, TC_W2 = 32 * TC_CHUNK;
constexpr int TC_RED = TC_W2 + 512, TC_BAR = TC_RED + 128, TC_TPTR = TC_BAR + 16;
constexpr int TC_SMEM = TC_TPTR + 16;
constexpr int TC_CTAS_PER_SM = 3;  // 66 KB of shared memory and 128 + 32 TMEM columns each

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (>> 4), version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor for kind::f16: D = f32, A = B = bf16, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
// 32 consecutive f32 columns of this thread's TMEM lane (load and wait in one statement: nothing may read r[] before the wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// v = hi + mid + lo exactly, each piece a bf16 (returned as the upper 16 bits of an f32 pattern); truncation keeps
// every remainder representable, so the two subtractions are exact.
__device__ __forceinline__ void split3(float v, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
    hi = __float_as_uint(v) & 0xFFFF0000u;
    const float r1 = __fsub_rn(v, __uint_as_float(hi));
    mid = __float_as_uint(r1) & 0xFFFF0000u;
    const float r2 = __fsub_rn(r1, __uint_as_float(mid));
    lo = __float_as_uint(r2) & 0xFFFF0000u;
}
// two upper halves -> one bf16x2 word (first element in the low half)
__device__ __forceinline__ uint32_t pack_hi16(uint32_t first, uint32_t second) { return __byte_perm(first, second, 0x7632); }

// Store e[0 .. 8 * NCHUNK) (upper-half bf16 patterns) as row `row` of an operand: chunk c at base + c * TC_CHUNK + row * 16.
template <int NCHUNK>
__device__ __forceinline__ void store_row(unsigned char *base, int row, const uint32_t *e) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
        *reinterpret_cast<uint4 *>(base + c * TC_CHUNK + row * 16) =
            make_uint4(pack_hi16(e[8 * c], e[8 * c + 1]), pack_hi16(e[8 * c + 2], e[8 * c + 3]),
                       pack_hi16(e[8 * c + 4], e[8 * c + 5]), pack_hi16(e[8 * c + 6], e[8 * c + 7]));
}

}  // namespace tc

// One launch = one full-batch pass; grid <= TC_CTAS_PER_SM * SMs, block = 128, dynamic smem = tc::TC_SMEM.
__global__ void __launch_bounds__(tc::TC_THREADS, tc::TC_CTAS_PER_SM) value_pass_tc_kernel(PassArgs a) {
    using namespace tc;
    constexpr int F = 5, H = 128, P = H * F + H + H + 1, W = P + NSCALAR;
    constexpr int NY = 3 * (F + 1);  // 18 meaningful columns of G
    if (a.skip_flag && *a.skip_flag) return;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA1 = smem + TC_A1, *sB1 = smem + TC_B1, *sA2 = smem + TC_A2, *sB2 = smem + TC_B2;
    float *w2s = reinterpret_cast<float *>(smem + TC_W2);
    double *red = reinterpret_cast<double *>(smem + TC_RED);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(smem + TC_TPTR);
    const uint32_t bar1 = smem_u32(smem + TC_BAR), bar2 = bar1 + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: this thread's hidden unit -> row `tid` of the B operand of MMA1 ----
    const float *tw1 = a.theta, *tb1 = tw1 + H * F, *tw2 = tb1 + H, *tb2 = tw2 + H;
    float wrow[F + 1];
#pragma unroll
    for (int f = 0; f < F; ++f) wrow[f] = tw1[tid * F + f];
    wrow[F] = tb1[tid];
    const float w2j = tw2[tid], b2 = tb2[0];
    w2s[tid] = w2j;
    {
        uint32_t hi[F + 1], mid[F + 1], lo[F + 1], e[48];
#pragma unroll
        for (int f = 0; f <= F; ++f) split3(wrow[f], hi[f], mid[f], lo[f]);
#pragma unroll
        for (int k = 0; k < 48; ++k) {
            const int g = k / (F + 1), f = k % (F + 1);  // piece pairing: x [hi hi mid mid hi lo] . w [hi mid hi mid lo hi]
            e[k] = k >= 6 * (F + 1) ? 0u : (g == 0 || g == 2 || g == 5) ? hi[f] : (g == 1 || g == 3) ? mid[f] : lo[f];
        }
        store_row<6>(sB1, tid, e);
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(tptr), 128);
        tmem_alloc(smem_u32(tptr + 1), 32);
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d1 = tptr[0], tmem_d2 = tptr[1];
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    constexpr uint32_t IDESC1 = make_idesc(128, 128, false, false);  // X (K-major) . W1e (K-major)
    constexpr uint32_t IDESC2 = make_idesc(128, 32, true, true);     // Mask^T (MN-major) . Y (MN-major)
    const uint32_t aA1 = smem_u32(sA1), aB1 = smem_u32(sB1), aA2 = smem_u32(sA2), aB2 = smem_u32(sB2);

    const uint64_t TE = a.T * a.E, ntiles = (TE + 127) / 128;
    double G[NY], loss_acc = 0.0, count_acc = 0.0, gb2_acc = 0.0;
#pragma unroll
    for (int n = 0; n < NY; ++n) G[n] = 0.0;

    struct Staged {
        float x[F], tgt;
        bool valid;
    };
    auto load_tile = [&](uint64_t tile, Staged &st) {
        const uint64_t n = tile * 128 + tid;
        const bool in_range = tile < ntiles && n < TE;
        const uint64_t t = in_range ? n / a.E : 0, e = in_range ? n - t * a.E : 0;
        const uint8_t code = in_range ? __ldg(a.succ + n) : (uint8_t)RL_PAD;
        float x[F];
#pragma unroll
        for (int f = 0; f < F; ++f) x[f] = in_range ? __ldg(a.obs + (t * F + f) * a.E + e) : 0.0f;
        const float tgt = in_range ? __ldg(a.target + n) : 0.0f;
        st.valid = code != RL_PAD;
#pragma unroll
        for (int f = 0; f < F; ++f) st.x[f] = st.valid ? x[f] : 0.0f;
        st.tgt = st.valid ? tgt : 0.0f;
    };
    auto drain_g = [&](uint32_t parity) {
        // G of the previous tile: wait for MMA2, read this unit's 18 columns, add in f64
        mbar_wait(bar2, parity);
        fence_after();
        uint32_t r[32];
        tmem_ld32(tmem_d2 + lane_off, r);
#pragma unroll
        for (int n = 0; n < NY; ++n) G[n] += (double)__uint_as_float(r[n]);
    };

    Staged nxt;
    load_tile(blockIdx.x, nxt);
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const Staged cur = nxt;
        if (it > 0) drain_g((it - 1) & 1u);

        // ---- A operand of MMA1: this sample's row of X (pieces of the 5 features and of the bias input 1) ----
        {
            uint32_t hi[F + 1], mid[F + 1], lo[F + 1], e[48];
#pragma unroll
            for (int f = 0; f < F; ++f) split3(cur.x[f], hi[f], mid[f], lo[f]);
            hi[F] = cur.valid ? 0x3F800000u : 0u;
            mid[F] = 0u;
            lo[F] = 0u;
#pragma unroll
            for (int k = 0; k < 48; ++k) {
                const int g = k / (F + 1), f = k % (F + 1);
                e[k] = k >= 6 * (F + 1) ? 0u : (g == 0 || g == 1 || g == 4) ? hi[f] : (g == 2 || g == 3) ? mid[f] : lo[f];
            }
            store_row<6>(sA1, tid, e);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int k = 0; k < 3; ++k)  // K = 16 per instruction = two 8-element chunks
                umma_bf16(tmem_d1, make_desc(aA1 + k * 2 * TC_CHUNK, TC_CHUNK, 128), make_desc(aB1 + k * 2 * TC_CHUNK, TC_CHUNK, 128),
                          IDESC1, k > 0);
            umma_commit(bar1);
        }
        load_tile(tile + gridDim.x, nxt);  // in flight during the MMA and the epilogue

        // ---- epilogue of MMA1: relu, V, mask ----
        mbar_wait(bar1, it & 1u);
        fence_after();
        float2 zacc = f2(0.0f, 0.0f);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_d1 + lane_off + c * 32, r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wa = *reinterpret_cast<const float4 *>(w2s + c * 32 + q * 8);
                const float4 wb = *reinterpret_cast<const float4 *>(w2s + c * 32 + q * 8 + 4);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                uint32_t m[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float p0 = __uint_as_float(r[q * 8 + 2 * i]), p1 = __uint_as_float(r[q * 8 + 2 * i + 1]);
                    zacc = __ffma2_rn(f2(wv[2 * i], wv[2 * i + 1]), f2(fmaxf(p0, 0.0f), fmaxf(p1, 0.0f)), zacc);
                    m[i] = (p0 > 0.0f ? 0x3F80u : 0u) | (p1 > 0.0f ? 0x3F800000u : 0u);  // bf16 1.0 / 0.0
                }
                // Mask^T, MN-major: chunk = 8 units, row = sample
                *reinterpret_cast<uint4 *>(sA2 + (c * 4 + q) * TC_CHUNK + tid * 16) = make_uint4(m[0], m[1], m[2], m[3]);
            }
        }
        // opt.rs:109-115: mse_loss(V(obs), targets, Mean)
        const float z = (zacc.x + zacc.y) + b2;
        const float diff = z - cur.tgt;
        float dz = 0.0f;
        if (cur.valid) {
            loss_acc += (double)(diff * diff);
            count_acc += 1.0;
            dz = 2.0f * diff;
            gb2_acc += (double)dz;
        }
        {
            uint32_t hi[F + 1], mid[F + 1], lo[F + 1], e[32];
#pragma unroll
            for (int f = 0; f < F; ++f) split3(dz * cur.x[f], hi[f], mid[f], lo[f]);
            split3(dz, hi[F], mid[F], lo[F]);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int g = k / (F + 1), f = k % (F + 1);
                e[k] = k >= NY ? 0u : g == 0 ? hi[f] : g == 1 ? mid[f] : lo[f];
            }
            store_row<4>(sB2, tid, e);
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)  // 16 samples per instruction = two 8-sample groups of 128 B
                umma_bf16(tmem_d2, make_desc(aA2 + k * 256, 128, TC_CHUNK), make_desc(aB2 + k * 256, 128, TC_CHUNK), IDESC2, k > 0);
            umma_commit(bar2);
        }
    }
    if (it > 0) drain_g((it - 1) & 1u);

    // ---- this CTA's partial row ----
    double *row = a.partials + (size_t)blockIdx.x * W;
    double Gf[F + 1];
#pragma unroll
    for (int f = 0; f <= F; ++f) Gf[f] = (G[f] + G[(F + 1) + f]) + G[2 * (F + 1) + f];
    double gw2 = (double)wrow[F] * Gf[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        row[tid * F + f] = (double)w2j * Gf[f];
        gw2 += (double)wrow[f] * Gf[f];
    }
    row[H * F + tid] = (double)w2j * Gf[F];
    row[H * F + H + tid] = gw2;
    const double s0 = warp_sum_f64(loss_acc), s1 = warp_sum_f64(count_acc), s2 = warp_sum_f64(gb2_acc);
    if (lane == 0) {
        red[warp * 4 + 0] = s0;
        red[warp * 4 + 1] = s1;
        red[warp * 4 + 2] = s2;
    }
    fence_before();
    __syncthreads();
    if (tid == 0) {
        double l = 0.0, n = 0.0, g = 0.0;
        for (int w = 0; w < TC_THREADS / 32; ++w) {
            l += red[w * 4 + 0];
            n += red[w * 4 + 1];
            g += red[w * 4 + 2];
        }
        row[P - 1] = g;
        row[P + SC_LOSS] = l;
        row[P + SC_KL] = 0.0;
        row[P + SC_ENTROPY] = 0.0;
        row[P + SC_COUNT] = n;
    }
    if (warp == 0) {
        fence_after();
        tmem_dealloc(tmem_d1, 128);
        tmem_dealloc(tmem_d2, 32);
    }
}
