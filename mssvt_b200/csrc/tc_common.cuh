// tc_common.cuh -- thin inline-PTX layer over the 5th-generation tensor cores (tcgen05 + TMEM).
// Descriptor formats follow cute/arch/mma_sm100_desc.hpp of CUTLASS (read, not included).
#pragma once
#include <cstdio>
#include "block_common.cuh"

namespace mssvt {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, no swizzle.  LBO = byte stride between the two 16-byte K chunks of one MMA,
// SBO = byte stride between 8-row core-matrix groups (cute/arch/mma_sm100_desc.hpp semantics).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset 0, layout_type 0 (SWIZZLE_NONE)
}

// The same descriptor as (base, byte offset): the base is built once per operand region and kept in two registers,
// a tile / K chunk inside the region is one 32-bit add on the address field (shared-memory addresses >> 4 fit the
// 14-bit field with room to spare, so the add never carries out of it).  The issuing thread then spends one uniform
// add per MMA instead of rebuilding add / shift / mask / or chains in front of every tcgen05.mma.
struct UmmaDescBase { uint32_t lo, hi; };
__device__ __forceinline__ UmmaDescBase umma_desc_base(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    UmmaDescBase b;
    b.lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    b.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    asm volatile("" : "+r"(b.lo));   // (opaque: the compiler must not re-derive it from the address at every use)
    return b;
}
__device__ __forceinline__ uint64_t umma_desc_at(const UmmaDescBase &b, uint32_t byte_off) {
    return ((uint64_t)b.hi << 32) | (uint64_t)(b.lo + (byte_off >> 4));
}

// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// A operand read from TMEM (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// kind::f16 with bf16 operands, fp32 accumulate, both operands K-major (a_format = b_format = 1: BF16).  One
// instruction covers K = 16 (32 bytes = two 16-byte core-matrix chunks of 8 elements).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// A operand read from TMEM: lane = row, one 32-bit column per PAIR of K elements (element k in the low half)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// two fp32 -> packed bf16x2 (round to nearest even): `lo` in bits [0, 16), `hi` in bits [16, 32)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// "bf16x3": a = hi + mid with hi = bf16(a), mid = bf16(a - hi) -- 16 significant bits in two bf16 (and the same for the
// weights); A W^T ~= A_hi W_hi^T + A_mid W_hi^T + A_hi W_mid^T on kind::f16 with fp32 accumulation.  Relative error of a
// product ~2^-16: results within 1e-4 of an fp32 reference with a wide margin (measured ~2e-5), from operand tiles
// of the size of ONE TF32 tile -- the occupancy of the plain TF32 kernels instead of the split-TF32 ones.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t &hi, uint32_t &mid) {
    hi = pack_bf16x2(a, b);
    mid = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
                 : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}

// bounded wait: a mis-programmed MMA must not hang the GPU.  try_wait suspends the thread for a
// hardware-defined slice per call, so the bound is on wall time (%globaltimer), not on iterations:
// after 2 s the kernel reports where it is stuck and traps (-> cudaErrorLaunchFailure).
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    // try_wait suspends the thread for a hardware-defined slice per attempt.  Measured alternatives: a
    // suspend-time hint (every blocking wait of the 3xTF32 kernels then slept for the whole hint) and polling
    // with test_wait (no gain for the tile kernel, slower FFN: eight polling warps take issue slots from the
    // epilogue of the other CTA).
    unsigned long long t0 = 0;
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .u32 n;\n\t"
            "mov.u32 n, 256;\n\t"
            "MBAR_TRY:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "@p bra MBAR_DONE;\n\t"
            "sub.u32 n, n, 1;\n\t"
            "setp.ne.u32 p, n, 0;\n\t"
            "@p bra MBAR_TRY;\n\t"
            "setp.ne.u32 p, n, 0;\n\t"
            "MBAR_DONE:\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (done) break;
        // bounded: a mis-programmed MMA must not hang the GPU; after 2 s report and trap
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) {
            if ((threadIdx.x & 31) == 0)
                printf("mssvt_b200: mbarrier wait timed out (block %d, thread %d, parity %u)\n", blockIdx.x,
                       threadIdx.x, parity);
            __trap();
        }
    }
    // lanes leave the wait loop at different times; the tcgen05.ld / fence instructions that follow
    // are .sync.aligned (warp-collective), so reconverge explicitly
    __syncwarp();
}

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// store 32 consecutive TMEM columns of this thread's lane (completion: tmem_st_wait)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// store 16 consecutive TMEM columns of this thread's lane (completion: tmem_st_wait)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// read-only 16-byte load that does not allocate in L1: for rows that are used once per tile, so that
// they do not evict what the tile keeps in L1 (prefetched query rows, index lists)
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

// ---- warp-cooperative row transfer.  With one thread per row, a warp-wide 16-byte load touches 32
// different 128-byte lines and the L1 tag stage (one line per cycle) becomes the bound.  Here the 32
// lanes fetch 4 rows x 128 bytes per instruction (8 lanes per row), park them in a warp-private 4 KB
// staging area (16-byte chunks XOR-swizzled by the row: conflict-free both ways) and every lane then
// reads its own row back.  my_src / my_dst: the 128-byte segment of this lane's row (nullptr = no row).
template <bool STREAM = false>  // STREAM: do not allocate the rows in L1 (they are used once per tile)
__device__ __forceinline__ void warp_rows_load(char *stg, const float4 *my_src, float4 *dst /*[8]*/) {
    const int lane = threadIdx.x & 31, st_row = lane >> 3, st_ch = lane & 7;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 *p = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)my_src, 4 * i + st_row);
        v[i] = !p ? make_float4(0.f, 0.f, 0.f, 0.f) : STREAM ? ldg_stream(p + st_ch) : __ldg(p + st_ch);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + st_row;
        *(float4 *)(stg + rr * 128 + ((st_ch ^ (rr & 7)) << 4)) = v[i];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = *(const float4 *)(stg + lane * 128 + ((q ^ (lane & 7)) << 4));
    __syncwarp();
}
__device__ __forceinline__ void warp_rows_store(char *stg, float4 *my_dst, const float4 *src /*[8]*/) {
    const int lane = threadIdx.x & 31, st_row = lane >> 3, st_ch = lane & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q) *(float4 *)(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = src[q];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + st_row;
        float4 *p = (float4 *)__shfl_sync(0xffffffffu, (unsigned long long)my_dst, rr);
        if (p) p[st_ch] = *(const float4 *)(stg + rr * 128 + ((st_ch ^ (rr & 7)) << 4));
    }
    __syncwarp();
}

// 8 lanes per 128-byte slice: the warp copies the 128-byte slices (at `base`, row pitch 64 floats) of the
// feature rows of its 32 lanes (my_row: this lane's row, < 0 = none) asynchronously into its 4 KB staging area
// (16-byte chunks XOR-swizzled by the row).  Straight-line code: the eight row ids are shuffled first, the
// copies are predicated (a branch per copy serialised shuffle -> compare -> branch -> copy eight times).
__device__ __forceinline__ void warp_rows_copy_async(char *stg, const float *base, int my_row) {
    const int lane = threadIdx.x & 31, st_row = lane >> 3, st_ch = lane & 7;
    const uint32_t dst = smem_u32(stg);
    int rows[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rows[i] = __shfl_sync(0xffffffffu, my_row, 4 * i + st_row);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + st_row;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ge.s32 p, %2, 0;\n\t"
            "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t"
            "}\n" ::"r"(dst + (uint32_t)(rr * 128 + ((st_ch ^ (rr & 7)) << 4))),
            "l"(base + (size_t)max(rows[i], 0) * 64 + st_ch * 4), "r"(rows[i])
            : "memory");
    }
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- "3xTF32": a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi) (and the same for the weights);
// A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T with fp32 accumulation carries ~21 mantissa bits -- fp32-grade
// results (measured ~1e-6 relative) from the TF32 tensor pipe at three MMAs per K step.  Kernels are
// templated on TERMS (1: plain TF32, 3: split operands); operand tiles come in (hi, lo) pairs, packed weights
// as [hi | lo] (mssvt_pack_operand_tf32 with terms = 3).
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
    hi = to_tf32(v);
    lo = to_tf32(v - hi);
}
__device__ __forceinline__ void split_tf32(const float4 &v, float4 &hi, float4 &lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}
// one K = 8 step of D (+)= A B^T from shared-memory operands, TERMS MMAs; *_lo = byte offset of the lo tile
template <int TERMS>
__device__ __forceinline__ void umma_step(uint32_t d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_lo, uint32_t b_addr,
                                          uint32_t b_lbo, uint32_t b_lo, uint32_t idesc, bool first) {
    const uint64_t ah = umma_smem_desc(a_addr, a_lbo, 128), bh = umma_smem_desc(b_addr, b_lbo, 128);
    umma_tf32(d, ah, bh, idesc, first ? 0u : 1u);
    if (TERMS == 3) {
        umma_tf32(d, umma_smem_desc(a_addr + a_lo, a_lbo, 128), bh, idesc, 1u);
        umma_tf32(d, ah, umma_smem_desc(b_addr + b_lo, b_lbo, 128), idesc, 1u);
    }
}

// Weight operands are packed ONCE (mssvt_pack_operand_tf32: canonical K-major layout, TF32-rounded) and
// then only copied: asynchronous 16-byte copies global -> shared, no registers, no per-CTA conversion.
// Call stage_packed_wait() before the fence.proxy.async / barrier that precedes the first MMA.
__device__ __forceinline__ void stage_packed(const float *__restrict__ packed, int n_floats, char *dst) {
    const uint32_t d = smem_u32(dst);
    for (int o = threadIdx.x * 4; o < n_floats; o += blockDim.x * 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)o * 4u), "l"(packed + o)
                     : "memory");
}
__device__ __forceinline__ void stage_packed_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// cp.async groups: commit what this thread has issued so far as one group; wait until at most N of its most
// recent groups are still in flight
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The same for whole matrices, by the TMA engine: ONE thread posts the expected byte count on an mbarrier and
// issues one bulk copy global -> shared per matrix (cp.async.bulk, SASS UBLKCP); no thread touches the data,
// and the copy lands through the async proxy, the one tcgen05.mma reads shared memory through.  Every thread
// waits on the barrier (phase 0) before the first MMA that uses the weights.  src / dst 16-byte aligned, bytes a
// multiple of 16.
__device__ __forceinline__ void bulk_expect(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(const void *src, uint32_t bytes, char *dst, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}


// one full warp; `last`: this CTA allocates nothing more (the permit is given back, so that other CTAs of the SM
// do not queue behind it)
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols, bool last = true) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols)
                 : "memory");
    if (last) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {  // the same warp
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// One lane of a converged warp (elect.sync): the issuing thread of tcgen05.mma / commit.  Used under a warp-uniform
// condition (`warp_uniform() == 0 && elect_one()`), the compiler keeps the descriptors on the uniform datapath
// instead of looping over the "possibly different" values of a divergent `tid == 0` branch (R2UR under BRA.U.ANY).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int warp_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// `if (issuer_elected()) { tcgen05.mma ...; tcgen05.commit }`: one lane of warp 0.  (Evaluated at the site: one
// shuffle, instead of a flag that would have to live in a register across the whole tile loop.)
__device__ __forceinline__ bool issuer_elected() { return warp_uniform() == 0 && elect_one(); }

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

}  // namespace mssvt
