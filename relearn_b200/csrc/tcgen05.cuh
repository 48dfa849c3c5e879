// tcgen05.cuh -- building blocks shared by the tensor-core kernels (the update passes in pass_tc.cuh and the
// tensor-core rollout in rollout.cu): tcgen05.mma / commit / TMEM alloc + load wrappers, mbarrier helpers, the
// no-swizzle canonical UMMA operand layout ([chunk of 8 K-elements][row][16 B]) and the exact three-piece bf16 split.
#pragma once

#include <cstdint>

namespace tc {

constexpr int TC_THREADS = 128;
constexpr int TC_CHUNK = 2048;  // bytes of one 8-element chunk over 128 rows
// operand regions, in chunks: X 5, W1e 5, one zero chunk (K-slots 40..47 of both), mask 16, Y 4 (Q-loss: 6), C 4 or 6
constexpr int TC_A1 = 0, TC_B1 = 5 * TC_CHUNK, TC_Z = 10 * TC_CHUNK, TC_A2 = 11 * TC_CHUNK, TC_B2 = 27 * TC_CHUNK;
__host__ __device__ constexpr int tc_n3(int blocks) { return (24 * blocks + 15) / 16 * 16; }  // MMA3 N: 32 / 48 (24 columns per block)
__host__ __device__ constexpr int tc_n2(int yblocks) { return (18 * yblocks + 15) / 16 * 16; }  // MMA2 N: 32 / 48
__host__ __device__ constexpr int tc_b3(int yblocks) { return TC_B2 + tc_n2(yblocks) / 8 * TC_CHUNK; }  // C follows Y
__host__ __device__ constexpr int tc_red(int blocks, int yblocks) { return tc_b3(yblocks) + tc_n3(blocks) / 8 * TC_CHUNK; }
__host__ __device__ constexpr int tc_smem(int blocks, int yblocks) { return tc_red(blocks, yblocks) + 512 + 32 + 16; }
constexpr int TC_CTAS_PER_SM = 3;  // 70.6 KB (critic) / 74.5 KB (policy) of shared memory and 128 + 32 TMEM columns each
constexpr int TC_DRAIN = 8;        // tiles accumulated in TMEM between f64 drains

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
// mbar_wait that the compiler may not hoist above the computation of the given values (they are inputs of the asm):
// keeps independent work -- issued to overlap an MMA -- ahead of the wait instead of behind it.
__device__ __forceinline__ void mbar_wait_after(uint32_t bar, uint32_t parity, double a, double b, double c, double d, uint32_t e) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity), "d"(a), "d"(b), "d"(c), "d"(d), "r"(e) : "memory");
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
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

// 64 consecutive f32 columns in one instruction (one wait for twice the data of tmem_ld32)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                 "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                 "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                   "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                   "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
                   "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// the first N meaningful columns of this thread's lane (padding columns are not fetched): 18 (G), 24 / 48 (Q)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t *r) {
    static_assert(N == 18 || N == 24 || N == 36 || N == 48, "column counts of this kernel");
    if (N == 18) {
        tmem_ld16(taddr, r);
        tmem_ld2(taddr + 16, r + 16);
    } else if (N == 36) {
        tmem_ld32(taddr, r);
        tmem_ld4(taddr + 32, r + 32);
    } else if (N == 24) {
        tmem_ld16(taddr, r);
        tmem_ld8(taddr + 16, r + 16);
    } else {
        tmem_ld32(taddr, r);
        tmem_ld16(taddr + 32, r + 32);
    }
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

// bf16 pair (1.0 where v > 0 else 0.0) from the sign bits of two f32 values that are never +0: PRMT in sign-replicate
// mode spreads bit 31 of each value over a half word, one LOP3 turns "negative" into 0 and the rest into 0x3F80.
__device__ __forceinline__ uint32_t relu_mask_bf16x2(uint32_t first, uint32_t second) {
    uint32_t neg;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(neg) : "r"(first), "r"(second));
    return ~neg & 0x3F803F80u;
}

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
