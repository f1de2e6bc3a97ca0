// tcgen05 / TMEM primitives for the fused per-graph kernel (embed_tc_kernel.cuh), spelled once for the sm_100a build
// (inline PTX) and once for the host-compiler emulator of tests/emu (plain loops over a per-block TMEM array), so that
// the kernel's data layouts — swizzled operand planes, TMEM column maps, row / lane assignments — are checked on the
// CPU before a GPU minute is spent.  What the emulator cannot check is ordering (it executes an MMA synchronously at
// issue): fences, commits and mbarrier waits are written for the hardware and are no-ops / trivially satisfied there.
//
// Conventions (all M = 128, cta_group::1, kind::tf32, fp32 accumulate):
//   * an operand TILE is K-major with the 128-byte swizzle: row r occupies bytes [r*128, r*128+128) of its ATOM (32 fp32
//     values of K), 16-byte chunk c of row r stored at chunk c ^ (r & 7); atoms of one operand follow each other
//     (K = 64 -> two atoms); the atom base is 1024-byte aligned.  One MMA consumes a K STEP of 8 values: chunks
//     2*ks, 2*ks+1 of every row of an atom.
//   * a TMEM address is (lane << 16) | column; D[m][n] lives at lane m, column d_col + n; a TMEM-resident A operand holds
//     A[m][k] at lane m, column a_col + k.
#pragma once
#include "common.cuh"

namespace sgpr {
namespace tc {

__device__ __forceinline__ uint32_t sw128_float_offset(int r, int k) {      // float index of element (r, k), k < 32, in an atom
    return static_cast<uint32_t>(r) * 32u + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
}

#ifdef SGPR_EMU
// ------------------------------------------------------------------ emulator ------------------------------------------
inline uint32_t g_tmem[128][512];            // blocks run one after the other: one TMEM image is "the SM's"

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t b = __float_as_uint(x);
    b = (b + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float(b);
}
__device__ __forceinline__ float tf32_operand(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t alloc(uint32_t* slot, int) { if ((threadIdx.x & 31) == 0) *slot = 0; return 0; }
__device__ __forceinline__ void dealloc(uint32_t, int) {}
__device__ __forceinline__ void fence_before() {}
__device__ __forceinline__ void fence_after() {}
__device__ __forceinline__ void fence_proxy_async() { std::atomic_thread_fence(std::memory_order_seq_cst); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {          // count-1 barriers only (what the kernel uses)
    std::atomic_ref<uint64_t>(*bar).fetch_xor(1);
}
__device__ __forceinline__ void commit(uint64_t* bar) { mbar_arrive(bar); }
// D[m][j] (+)= sum_kk A[m][8 ks + kk] * B[j][8 ks + kk],  m < 128, j < n
__device__ __forceinline__ void mma_ss(uint32_t d, const float* a_atom, const float* b_atom, int ks, int n, bool accumulate) {
    const int dc = d & 0xffff;
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < n; ++j) {
            float acc = accumulate ? __uint_as_float(g_tmem[m][dc + j]) : 0.0f;
            for (int kk = 0; kk < 8; ++kk)
                acc = std::fmaf(tf32_operand(a_atom[sw128_float_offset(m, 8 * ks + kk)]), tf32_operand(b_atom[sw128_float_offset(j, 8 * ks + kk)]), acc);
            g_tmem[m][dc + j] = __float_as_uint(acc);
        }
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, const float* b_atom, int ks, int n, bool accumulate) {
    const int dc = d & 0xffff, ac = a_tmem & 0xffff;
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < n; ++j) {
            float acc = accumulate ? __uint_as_float(g_tmem[m][dc + j]) : 0.0f;
            for (int kk = 0; kk < 8; ++kk)
                acc = std::fmaf(tf32_operand(__uint_as_float(g_tmem[m][ac + kk])), tf32_operand(b_atom[sw128_float_offset(j, 8 * ks + kk)]), acc);
            g_tmem[m][dc + j] = __float_as_uint(acc);
        }
}
// operand handles: on the hardware a 64-bit shared-memory descriptor (advanced by plain adds), here the atom pointer + step
struct Operand { const float* atom; int ks; };
__device__ __forceinline__ Operand operand(const float* atom0) { return Operand{atom0, 0}; }
__device__ __forceinline__ Operand advance(Operand o, int ks) { return Operand{o.atom + (ks >> 2) * 2048, ks & 3}; }   // K step 0..7 over two atoms of 64 rows
__device__ __forceinline__ void mma_ss(uint32_t d, Operand a, Operand b, int n, bool accumulate) { mma_ss(d, a.atom, b.atom, a.ks, n, accumulate); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, Operand b, int n, bool accumulate) { mma_ts(d, a_tmem, b.atom, b.ks, n, accumulate); }
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xffff;
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(g_tmem[lane][col + i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&v)[16]) {
    const int lane = (taddr >> 16) + (threadIdx.x & 31), col = taddr & 0xffff;
    for (int i = 0; i < 16; ++i) g_tmem[lane][col + i] = __float_as_uint(v[i]);
}
__device__ __forceinline__ void wait_ld() {}
__device__ __forceinline__ void wait_st() {}
#else
// ------------------------------------------------------------------ sm_100a -------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t alloc(uint32_t* slot, int cols) {          // whole warp; returns nothing useful until a barrier
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    return 0;
}
__device__ __forceinline__ void dealloc(uint32_t base, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t atom_desc(const float* atom, int ks) {     // K-major, SWIZZLE_128B, SBO 1024, version 1
    const uint32_t addr = smem_u32(atom) + 32u * static_cast<uint32_t>(ks);
    return static_cast<uint64_t>((addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) {                        // M = 128, fp32 accumulate, both operands K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, const float* a_atom, const float* b_atom, int ks, int n, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(atom_desc(a_atom, ks)), "l"(atom_desc(b_atom, ks)), "r"(idesc_tf32(n)), "r"(accumulate ? 1u : 0u)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, const float* b_atom, int ks, int n, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a_tmem), "l"(atom_desc(b_atom, ks)), "r"(idesc_tf32(n)), "r"(accumulate ? 1u : 0u)
        : "memory");
}
// operand handles: the descriptor of K step 0 of the first atom, advanced by adds in 16-byte units (a K step is 32 bytes,
// the second atom of a 64-row plane starts 8 KB further)
struct Operand { uint64_t desc; };
__device__ __forceinline__ Operand operand(const float* atom0) { return Operand{atom_desc(atom0, 0)}; }
__device__ __forceinline__ Operand advance(Operand o, int ks) { return Operand{o.desc + static_cast<uint64_t>((ks >> 2) * 512 + (ks & 3) * 2)}; }
__device__ __forceinline__ void mma_ss(uint32_t d, Operand a, Operand b, int n, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a.desc), "l"(b.desc), "r"(idesc_tf32(n)), "r"(accumulate ? 1u : 0u)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, Operand b, int n, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a_tmem), "l"(b.desc), "r"(idesc_tf32(n)), "r"(accumulate ? 1u : 0u)
        : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#endif

}  // namespace tc
}  // namespace sgpr
