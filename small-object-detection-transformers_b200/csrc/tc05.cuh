// Thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with
// accumulators in TMEM, TMEM alloc / ld / st, mbarriers, proxy fences, cp.async.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace sodt {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"      // suspend-time hint: sleep until the phase completes
        "selp.u32 %0, 1, 0, p;\n\t}"                                            // instead of re-polling every few dozen cycles
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Busy poll (no suspend hint): for single-thread roles whose wake-up latency is on the critical path
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------- cp.async
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, N consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st16(taddr, r); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st32(taddr, r); }

// --------------------------------------------------------------------------------- tcgen05.mma
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" canonical layout of 8x16-byte
// core matrices).  lbo / sbo in bytes:
//   K-major  operand (rows = M or N index, 16-byte chunk = 8 consecutive K): sbo = stride between
//            8-row groups, lbo = stride between the two K chunks of one K=16 step;
//   MN-major operand (16-byte chunk = 8 consecutive M/N): lbo = stride between 8-K groups,
//            sbo = stride between 8-element M/N groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulation.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // D format: f32
           | (1u << 7)                     // A format: bf16
           | (1u << 10)                    // B format: bf16
           | ((a_mn_major ? 1u : 0u) << 15)
           | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17)
           | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with IEEE fp16 A/B (format code 0) and fp32 accumulation
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// Same, with a 128-bit "disable output lane" mask: TMEM lanes (accumulator rows) whose bit is set keep their old value.
__device__ __forceinline__ void mma_ss_masked(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate,
                                              uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
__device__ __forceinline__ void mma_ts_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate,
                                              uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
        : "memory");
}
// Arrives on the mbarrier when all tcgen05.mma issued so far by this thread have completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2): one MMA over the shared memory of two SMs
// The pair is a 2-CTA cluster; the even-ranked CTA (the leader) issues the MMAs with M = 256: rows 0-127 of A and of the
// accumulator live in the leader, rows 128-255 in its peer, and each CTA supplies half of the N rows of B.  Shared-memory
// descriptors and TMEM addresses are the leader's own; the hardware applies the same offsets in the peer.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {          // arrive on a (possibly remote) mbarrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {      // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// arrives on the mbarrier at the same offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void mma_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {   // single MUFU.EX2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------ packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMNMX3)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {     // FFMA2: two fp32 FMAs per issue slot
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {                  // FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// 2^x on the FMA / ALU pipes instead of the MUFU (16 results per clock and SM): round-to-nearest split x = n + f, |f| <= 0.5,
// degree-4 polynomial of 2^f (relative error 4e-5, far below bf16's 2^-9), n added to the exponent field.  x <= ~100.
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
    float lo, hi;
    unpack2(x2, lo, hi);
    x2 = pack2(fmaxf(lo, -125.f), fmaxf(hi, -125.f));
    const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
    const uint64_t t = fadd2(x2, magic);                    // n in the low mantissa bits
    const uint64_t n = fadd2(t, nmagic);
    float nl, nh;
    unpack2(n, nl, nh);
    const uint64_t f = fadd2(x2, pack2(-nl, -nh));
    uint64_t p = ffma2(f, pack2(0.0096181291f, 0.0096181291f), pack2(0.0555041087f, 0.0555041087f));
    p = ffma2(p, f, pack2(0.2402265070f, 0.2402265070f));
    p = ffma2(p, f, pack2(0.6931471806f, 0.6931471806f));
    p = ffma2(p, f, pack2(1.f, 1.f));
    float pl, ph, tl, th;
    unpack2(p, pl, ph);
    unpack2(t, tl, th);
    return pack2(__int_as_float(__float_as_int(pl) + (__float_as_int(tl) << 23)), __int_as_float(__float_as_int(ph) + (__float_as_int(th) << 23)));
}

__device__ __forceinline__ float fast_rcp(float x) {    // single MUFU.RCP (relative error 2^-23 class; no slow path)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_tanh(float x) {   // single MUFU.TANH, relative error <= 2^-11
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GELU for bf16 outputs: x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x q(x^2))), q a minimax fit of atanh(erf(x / sqrt 2)) / x
// on |x| <= 8 (fit error 2.5e-5 absolute, 20x below the tanh-GELU of the literature; x^2 is clamped so q stays positive).
// Together with the MUFU error the result is within 0.5 |x| 2^-11 of the erf form, i.e. under a quarter bf16 ulp.
__device__ __forceinline__ float gelu_fast(float x) {
    const float s = fminf(x * x, 64.f);
    float q = fmaf(s, -0.0003515175096f, 0.03700565079f);
    q = fmaf(q, s, 0.7975078786f);
    const float hx = 0.5f * x;
    return fmaf(hx, fast_tanh(x * q), hx);
}
// The same GELU on a PAIR of values in packed fp16 arithmetic (HMUL2 / HFMA2 / one MUFU.TANH.F16x2 for both): half the
// instructions per element.  fp16's 11-bit significand keeps the result within ~2^-11 of the fp32 evaluation, the same
// class as the MUFU.TANH error above; x^2 saturating to +inf for |x| > 255 is absorbed by the clamp.
__device__ __forceinline__ void gelu_fast2(float& a, float& b) {
    uint32_t x, s, q, t, hx, y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(b), "f"(a));
    asm("mul.f16x2 %0, %1, %1;" : "=r"(s) : "r"(x));
    asm("min.f16x2 %0, %1, %2;" : "=r"(s) : "r"(s), "r"(0x54005400u));                   // 64.0
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(s), "r"(0x8DC28DC2u), "r"(0x28BD28BDu));   // -0.0003515, 0.037006
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(q), "r"(s), "r"(0x3A613A61u));             // 0.797508
    asm("mul.f16x2 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(q));
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(q));
    asm("mul.f16x2 %0, %1, %2;" : "=r"(hx) : "r"(x), "r"(0x38003800u));                  // 0.5
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(hx), "r"(t));
    const __half2 h = *reinterpret_cast<const __half2*>(&y);
    a = __low2float(h);
    b = __high2float(h);
}
// TWICE the same GELU of a pair, as packed fp16 (low = a): 2 gelu(x) = x + x tanh(x q(x^2)) saves the halving; the consumer
// folds the factor 0.5 into its (fp16) weights.  5 packed-half FMA-pipe instructions and two MUFU.TANH per pair.
__device__ __forceinline__ uint32_t gelu2x_f16x2(float a, float b) {
    uint32_t x, s, q, t, y;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(b), "f"(a));
    asm("mul.f16x2 %0, %1, %1;" : "=r"(s) : "r"(x));
    asm("min.f16x2 %0, %1, %2;" : "=r"(s) : "r"(s), "r"(0x54005400u));                   // 64.0
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(s), "r"(0x8DC28DC2u), "r"(0x28BD28BDu));   // -0.0003515, 0.037006
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(q) : "r"(q), "r"(s), "r"(0x3A613A61u));             // 0.797508
    asm("mul.f16x2 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(q));
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(q));
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(x), "r"(t));
    return y;
}
// SiLU for bf16 outputs: x sigmoid(x) = 0.5 x (1 + tanh(x / 2)) exactly; one MUFU
__device__ __forceinline__ float silu_fast(float x) {
    const float hx = 0.5f * x;
    return fmaf(hx, fast_tanh(hx), hx);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// The same GELU of a pair in packed fp32x2 arithmetic (FMUL2 / FFMA2): 7 packed instructions, two FMNMX and two MUFU.TANH per
// pair -- no fp16 round trip (the packed-fp16 form costs 17 instructions with its conversions).
__device__ __forceinline__ uint64_t gelu_f32x2(uint64_t x2) {
    float s0, s1;
    unpack2(fmul2(x2, x2), s0, s1);
    const uint64_t s2 = pack2(fminf(s0, 64.f), fminf(s1, 64.f));
    uint64_t q2 = ffma2(s2, pack2(-0.0003515175096f, -0.0003515175096f), pack2(0.03700565079f, 0.03700565079f));
    q2 = ffma2(q2, s2, pack2(0.7975078786f, 0.7975078786f));
    float u0, u1;
    unpack2(fmul2(x2, q2), u0, u1);
    const uint64_t t2 = pack2(fast_tanh(u0), fast_tanh(u1));
    const uint64_t hx2 = fmul2(x2, pack2(0.5f, 0.5f));
    return ffma2(hx2, t2, hx2);
}
__device__ __forceinline__ uint32_t gelu_bf16x2_f32x2(uint64_t x2) {        // ... rounded to a packed bf16 pair (low = first)
    float y0, y1;
    unpack2(gelu_f32x2(x2), y0, y1);
    return pack_bf16(y0, y1);
}

}  // namespace tc
}  // namespace sodt
