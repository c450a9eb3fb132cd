// kj_device.cu -- CUDA side of libkaijub200.so: index upload, the persistent classification kernel and the
// C ABI entry points kj_create / kj_classify / kj_classify_device / kj_destroy (include/kaiju_b200.h).
//
// Execution model: one warp classifies one read item at a time (kj_core.h); a persistent grid of
// (#SM x resident CTAs) pulls read indices from a global counter, so fast ("U" after the length gate)
// and slow (long matches, SEG trims) reads balance without a host-side scheduler -- the GPU replacement of
// the reference's ProducerConsumerQueue + N ConsumerThreads (kaiju.cpp:250-257, 288-396).
// There is no CPU fallback: without a GPU kj_create() fails with KJ_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "kj_core.h"
#include "kj_core_greedy.h"
#include "kj_host.h"

#define KJ_WARPS_PER_CTA 8
#ifndef KJ_MIN_BLOCKS
#define KJ_MIN_BLOCKS 4          // resident CTAs per SM the register allocation is tuned for (ncu: latency-bound, see profiles/)
#endif
#ifndef KJ_MIN_BLOCKS_GREEDY_SPLIT
#define KJ_MIN_BLOCKS_GREEDY_SPLIT 5   // front-end / search kernels of the two-kernel Greedy path: A/B round 2 (r2i) 18.14 vs 17.90 M pairs/s, e2e 17.07 vs 16.41
#endif
#ifndef KJ_MIN_BLOCKS_GREEDY
#define KJ_MIN_BLOCKS_GREEDY 4   // A/B round 2 (5 CTAs = 48 registers): 12.0 vs 12.3 M pairs/s -- more warps, but more spill traffic and more instruction-fetch stalls
#endif
#define KJ_KEPT_SMEM_FIXED 20      // = KJ_KEPT_SMEM of kj_host.cpp (checked at launch: the fixed profile is only used when the layouts agree)
#define KJ_CHUNK_READS (1u << 20)
#define KJ_CHUNK_BYTES (1ull << 28)  // and at most this many bases of one mate per chunk (long reads)

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { kj_err() = std::string(#call) + ": " + cudaGetErrorString(e_); return KJ_ERR_CUDA; } } while (0)

struct KjCtaShared { KjDevIndex ix; KjTables tb; };

// ---- bulk copy (the Blackwell/Hopper copy engine, "TMA" in its 1-D form) of a warp's claimed reads into shared memory: one elected lane
// arms the warp's mbarrier with the byte count and issues cp.async.bulk; the 32 lanes wait on the barrier's phase.  The translation
// passes then read the bases from shared memory instead of waiting on global loads pass by pass.
static __device__ __forceinline__ uint32_t kj_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void kj_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(kj_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void kj_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(kj_smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void kj_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(kj_smem_u32(dst)), "l"(src), "r"(bytes), "r"(kj_smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void kj_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n KJ_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra KJ_DONE;\n bra KJ_WAIT;\n KJ_DONE:\n}"
                 :: "r"(kj_smem_u32(bar)), "r"(parity) : "memory");
}

// GWS = false: the per-warp work space is carved out of shared memory (the compiler keeps every access in the shared
// address space); GWS = true (reads too long for that): the same carve-up in a global buffer, generic loads and stores.
// FIX = true: the work-space carve-up of the standard short-read case (mates up to 152 bases, -m 11) as compile-time constants: every offset
// becomes an immediate of the shared-memory instructions instead of a value that is kept in (or re-derived into) registers.
static __host__ __device__ __forceinline__ KjRunParams kj_fixed_profile(int mode) {
    KjRunParams p{}; p.mode = mode; p.m = 11; p.max_len = 152; p.max_frag = 152 / 3 + 1; p.item_cap = 128; p.kept_cap_smem = KJ_KEPT_SMEM_FIXED; p.stage = 0;
    return p;
}
// VB = true: the verbose outputs (id sets, accession sets, fragment strings) are compiled in; the kernels of the normal path carry none of it.
// ROLE: 0 = the whole item in this kernel; 1 = front end only (translation, fragments, ranked queue -> a record per item in `prep`); 2 = search only (from the records).
// Greedy runs as the pair 1 + 2 over sub-batches of the launch (kj_core.h: the search loop then shares the instruction cache with nothing it does not need).
template <int MODE, class IdxT, bool GWS, bool FIX, bool VB, int ROLE>
__global__ void __launch_bounds__(KJ_WARPS_PER_CTA * 32, MODE == 0 ? KJ_MIN_BLOCKS : ROLE != 0 ? KJ_MIN_BLOCKS_GREEDY_SPLIT : KJ_MIN_BLOCKS_GREEDY)
kj_classify_kernel(const KjDevIndex* __restrict__ g_ix, const __grid_constant__ KjRunParams rp, const __grid_constant__ KjSmemLayout lay,
                   const uint8_t* __restrict__ seq1, const uint64_t* __restrict__ off1,
                   const uint8_t* __restrict__ seq2, const uint64_t* __restrict__ off2,
                   uint64_t base1, uint64_t base2, uint64_t n_reads,
                   uint64_t* __restrict__ taxon_out, uint32_t* __restrict__ best_out, uint64_t* __restrict__ ids_out, uint8_t* __restrict__ nids_out, uint32_t* __restrict__ compact_out,
                   unsigned long long* __restrict__ counter, KjKept* __restrict__ spill, uint8_t* __restrict__ gscratch,
                   uint32_t gscratch_bytes, uint8_t* __restrict__ gws, unsigned long long* __restrict__ counts, uint32_t* __restrict__ err,
                   uint32_t* __restrict__ acc_out, uint8_t* __restrict__ nacc_out, char* __restrict__ frag_out, uint32_t frag_stride, uint32_t* __restrict__ frag_len_out,
                   uint8_t* __restrict__ prep, uint32_t prep_stride, uint64_t r_begin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    KjCtaShared* sh = (KjCtaShared*)smem_raw;
    {   // stage the index descriptor (C[] etc.) and the small tables once per CTA
        const uint32_t* src = (const uint32_t*)g_ix; uint32_t* dst = (uint32_t*)&sh->ix;
        KJ_ROLLED
        for (uint32_t i = threadIdx.x; i < sizeof(KjDevIndex) / 4; i += blockDim.x) dst[i] = src[i];
        const uint32_t* ts = (const uint32_t*)g_ix->tables; uint32_t* td = (uint32_t*)&sh->tb;
        KJ_ROLLED
        for (uint32_t i = threadIdx.x; i < sizeof(KjTables) / 4; i += blockDim.x) td[i] = ts[i];
    }
    __syncthreads();
    const int warp_in_cta = threadIdx.x >> 5;
    KjWarpCtx cx;
    cx.w.lane = threadIdx.x & 31;
    cx.ix = &sh->ix; cx.rp = &rp; cx.tb = &sh->tb;
    if (FIX) cx.L = kj_smem_layout(kj_fixed_profile(MODE));      // folded at compile time
    else cx.L = lay;                                 // computed on the host, read from the parameter bank
    const uint64_t gwarp = (uint64_t)blockIdx.x * KJ_WARPS_PER_CTA + (uint64_t)warp_in_cta;
    // per-warp work space: shared memory, or (reads too long for it) a slice of a global buffer that stays L1/L2-resident
    if (GWS) cx.smem = gws + gwarp * cx.L.total;
    else cx.smem = smem_raw + kj_align((uint32_t)sizeof(KjCtaShared), 16) + (uint32_t)warp_in_cta * cx.L.total;
    cx.spill = spill + gwarp * rp.scratch_entries;
    cx.gscratch = gscratch + gwarp * gscratch_bytes;
    cx.err = err; cx.text = nullptr; cx.text_cap = frag_stride; cx.text_len = 0; cx.want_acc = VB && acc_out != nullptr;
    const bool paired = seq2 != nullptr;
#ifdef KJ_STAGE
    const bool stage = !GWS && rp.stage != 0;
#else
    const bool stage = false;          // measured: -10 % (the larger shared-memory carve-out shrinks the L1 that serves the rank loads), see profiles/README.md
#endif
    uint64_t* mbar = (uint64_t*)(cx.smem + cx.L.mbar_off); uint8_t* stg = cx.smem + cx.L.stage_off; uint32_t phase = 0;
    if (stage) { if (cx.w.lane == 0) kj_mbar_init(mbar, 1); cx.w.sync(); }
    // work distribution: a warp claims KJ_CLAIM consecutive items per atomic and fetches their offsets with one coalesced load
    KJ_ROLLED
    for (;;) {
        unsigned long long r0 = 0;
        if (cx.w.lane == 0) r0 = r_begin + atomicAdd(counter, (unsigned long long)KJ_CLAIM);      // this launch covers items [r_begin, n_reads)
        r0 = cx.w.shfl64(r0, 0);
        if (r0 >= n_reads) break;
        const unsigned long long ri = r0 + (unsigned long long)cx.w.lane;
        uint64_t o1 = 0, o2 = 0;
        if (cx.w.lane <= KJ_CLAIM && ri <= n_reads) { o1 = off1[ri] - base1; if (paired) o2 = off2[ri] - base2; }
        uint32_t al1 = 0, al2 = 0; uint64_t f1 = 0, f2 = 0;
        if (stage) {
            // the claimed reads are contiguous in the packed arrays: one bulk copy per mate, source and size rounded to 16 bytes
            const int nk = (int)(n_reads - r0 < (unsigned long long)KJ_CLAIM ? n_reads - r0 : (unsigned long long)KJ_CLAIM);
            f1 = cx.w.shfl64(o1, 0); f2 = cx.w.shfl64(o2, 0);
            const uint64_t l1 = cx.w.shfl64(o1, nk), l2 = cx.w.shfl64(o2, nk);
            al1 = (uint32_t)((uintptr_t)(seq1 + f1) & 15u); al2 = paired ? (uint32_t)((uintptr_t)(seq2 + f2) & 15u) : 0u;
            const uint32_t b1 = (al1 + (uint32_t)(l1 - f1) + 15u) & ~15u, b2 = paired ? (al2 + (uint32_t)(l2 - f2) + 15u) & ~15u : 0u;
            if (cx.w.lane == 0) {
                kj_mbar_expect_tx(mbar, b1 + b2);
                if (b1) kj_bulk_g2s(stg, seq1 + f1 - al1, b1, mbar);
                if (b2) kj_bulk_g2s(stg + cx.L.stage_stride, seq2 + f2 - al2, b2, mbar);
            }
            kj_mbar_wait(mbar, phase); phase ^= 1u;
        }
        KJ_ROLLED
        for (int k = 0; k < KJ_CLAIM; k++) {
            const unsigned long long r = r0 + (unsigned long long)k;
            if (r >= n_reads) break;
            const uint64_t a0 = cx.w.shfl64(o1, k), a1 = cx.w.shfl64(o1, k + 1), b0 = cx.w.shfl64(o2, k), b1 = cx.w.shfl64(o2, k + 1);
            const uint8_t* p1 = stage ? stg + al1 + (uint32_t)(a0 - f1) : seq1 + a0;
            const uint8_t* p2 = !paired ? nullptr : stage ? stg + cx.L.stage_stride + al2 + (uint32_t)(b0 - f2) : seq2 + b0;
            uint32_t best = 0;
            if (VB && frag_out) { cx.text = frag_out + r * frag_stride; cx.text_len = 0; }
            uint32_t t = kj_classify_item<MODE, IdxT, ROLE>(cx, p1, (int)(a1 - a0), p2, (int)(b1 - b0), paired, best, ROLE ? prep + (size_t)(r - r_begin) * prep_stride : nullptr);
            if (ROLE == 1) { cx.w.sync(); continue; }
            const uint64_t id = t == KJ_TAX_BAD ? 0ull : sh->ix.tax_id[t];
            if (cx.w.lane == 0) {
                if (taxon_out) taxon_out[r] = id;
                if (compact_out) compact_out[r] = id ? t : KJ_TAX_BAD;
                if (best_out) best_out[r] = id ? best : 0u;
                // per-taxon read counts (kaiju2table's first pass), fused: a separate counting kernel behind a persistent grid would stall the chunk pipeline
                if (counts) atomicAdd(counts + (id ? t : sh->ix.n_tax), 1ull);
            }
            if (VB && ids_out) {   // column 5 of the reference's -v output: the match-id set in ascending order (std::set), classified reads only
                const uint32_t nids = id ? cx.nids : 0u; const uint32_t* ids = (const uint32_t*)(cx.smem + cx.L.ids_off);
                if ((uint32_t)cx.w.lane < nids) {
                    const uint64_t mine = sh->ix.tax_id[ids[cx.w.lane]]; uint32_t rank = 0;
                    KJ_ROLLED
                    for (uint32_t u = 0; u < nids; u++) rank += sh->ix.tax_id[ids[u]] < mine ? 1u : 0u;
                    ids_out[r * KJ_MAX_IDS + rank] = mine;
                }
                if (cx.w.lane == 0) nids_out[r] = (uint8_t)nids;
            }
            if (VB && acc_out) {   // column 6: accession ranks of the visited sequences, ascending (the reference's std::set<std::string> order)
                const uint32_t* accs = (const uint32_t*)(cx.smem + cx.L.accs_off); const uint32_t nacc = id ? accs[20] : 0u;
                if ((uint32_t)cx.w.lane < nacc) {
                    const uint32_t mine = accs[cx.w.lane]; uint32_t rank = 0;
                    KJ_ROLLED
                    for (uint32_t u = 0; u < nacc; u++) rank += accs[u] < mine ? 1u : 0u;
                    acc_out[r * KJ_MAX_MATCH_ACC + rank] = mine;
                }
                if (cx.w.lane == 0) nacc_out[r] = (uint8_t)nacc;
            }
            if (VB && frag_len_out && cx.w.lane == 0) frag_len_out[r] = id ? cx.text_len : 0u;
            cx.w.sync();
        }
    }
}

__global__ void kj_maxlen_kernel(const uint64_t* __restrict__ off, uint64_t n, unsigned int* __restrict__ out) {
    unsigned int m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { uint64_t l = off[i + 1] - off[i]; m = max(m, (unsigned int)min(l, (uint64_t)0xffffffffu)); }
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// Per-taxon read counts (the input of kaiju2table, src/kaiju2table.cpp:186-245): dense index of each result by binary search in the
// two ascending runs of tax_id (nodes.dmp ids | DB taxa absent from it); slot n_tax = unclassified.
__global__ void kj_count_kernel(const uint64_t* __restrict__ taxon, uint64_t n, const uint64_t* __restrict__ tax_id, uint32_t n_present, uint32_t n_tax,
                                unsigned long long* __restrict__ counts) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t id = taxon[r]; uint32_t slot = n_tax;
        if (id) {
            uint32_t lo = 0, hi = n_present;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (tax_id[mid] < id) lo = mid + 1; else hi = mid; }
            if (lo < n_present && tax_id[lo] == id) slot = lo;
            else { lo = n_present; hi = n_tax; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (tax_id[mid] < id) lo = mid + 1; else hi = mid; } if (lo < n_tax && tax_id[lo] == id) slot = lo; }
        }
        atomicAdd(counts + slot, 1ull);
    }
}
__global__ void kj_count_commit(unsigned long long* __restrict__ total, unsigned long long* __restrict__ pending, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { total[i] += pending[i]; pending[i] = 0; }
}

// ------------------------------------------------------------------------------------------------
struct KjFilesState; static void kj_files_state_free(KjFilesState* S);      // kj_ingest.h
struct kj_ctx {
    KjFilesState* files = nullptr;   // buffers of kj_classify_files, kept between calls
    int device = 0; kj_params params{}; int sm_count = 0;
    KjHostIndex H;                 // big arrays are released after upload; small ones stay
    KjDevIndex dix{};              // host copy of the descriptor (device pointers inside)
    KjDevIndex* d_ix = nullptr; KjTables* d_tables = nullptr;
    KjDevIndex* d_ix_mem = nullptr; void* d_kmer_mem = nullptr; int kmer_k_mem = 0;      // the MEM kernels' own descriptor: same index, 7-mer table (kj_create)
    void* d_rank = nullptr; void* d_letters = nullptr; void* d_sa_tax = nullptr; void* d_seq_tax = nullptr;
    void* d_sa_acc = nullptr; void* d_seq_acc = nullptr;
    uint32_t* d_acc[2] = {nullptr, nullptr}; uint8_t* d_nacc[2] = {nullptr, nullptr}; char* d_frag[2] = {nullptr, nullptr}; uint32_t* d_fraglen[2] = {nullptr, nullptr}; size_t d_v2_cap = 0, d_frag_stride = 0;
    void* d_tax_parent = nullptr; void* d_tax_depth = nullptr; void* d_tax_id = nullptr; void* d_lnfact = nullptr; void* d_kmer = nullptr;
    uint64_t index_bytes = 0; uint64_t n_sa = 0; double build_ms = 0.0;
    // run state
    unsigned long long* d_counter = nullptr; uint32_t* d_err = nullptr; unsigned int* d_maxlen = nullptr;
    KjKept* d_spill = nullptr; size_t spill_bytes = 0; uint8_t* d_gscratch = nullptr; size_t gscratch_bytes_total = 0;
    double* d_evbreaks = nullptr; uint32_t n_evbreaks = 0; uint64_t* d_quirk = nullptr;
    unsigned long long* d_counts = nullptr; unsigned long long* d_counts_pending = nullptr; uint32_t n_counts = 0, n_present = 0;   // per-taxon read counts (+1 slot: unclassified)
    uint32_t variant_boost = 1;    // Greedy variant-ring capacity multiplier, raised after an overflow (flag 4) so that a retry succeeds
    uint8_t* d_ws = nullptr; size_t ws_bytes = 0;
    uint8_t* d_prep = nullptr; size_t prep_bytes = 0;      // prepared-item records of the two-kernel Greedy path (two slots x two buffers)
    cudaStream_t fstream[2] = {nullptr, nullptr}; cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, ev_s[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // front-end stream + hand-over events per slot
    cudaStream_t stream[2] = {nullptr, nullptr}; cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // staging for kj_classify (host buffers)
    uint8_t* d_seq[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; size_t d_seq_cap[2][2] = {{0, 0}, {0, 0}};
    uint64_t* d_off[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; uint64_t* d_tax[2] = {nullptr, nullptr}; uint32_t* d_best[2] = {nullptr, nullptr};
    uint64_t* d_ids[2] = {nullptr, nullptr}; uint8_t* d_nids[2] = {nullptr, nullptr}; size_t d_ids_cap = 0;
    size_t d_reads_cap = 0;
    uint64_t launches = 0; double last_kernel_ms = 0.0;
    int grid = 0; size_t smem_bytes = 0; int cfg_mode = -1;
};

template <class T> static int upload(const std::vector<T>& v, void** d, uint64_t& total) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CK(cudaMalloc(d, bytes));
    if (!v.empty()) CK(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    total += bytes; return KJ_OK;
}

// Shared-memory work space while three CTAs still fit on an SM; beyond that (mates longer than ~280 bases) the same
// carve-up is addressed in a global buffer instead, which keeps the grid at full occupancy for any read length.
// Measured (MEM, kernel-only, M pairs/s, shared vs global): PE150 57.8 vs 45.8, PE250 29.4 (3 CTAs/SM) vs 28.2, PE350 14.3 (2 CTAs/SM) vs 18.3.
#define KJ_SMEM_WS_LIMIT (75u * 1024u)
// the fixed-profile kernels apply when the batch's carve-up is exactly the compiled-in one
static bool kj_use_fixed(const KjRunParams& rp) {
    if (rp.ws_global || rp.mode != 1 || getenv("KJ_NO_FIXED")) return false;      // A/B round 2: Greedy +17 % (14.35 vs 12.26 M pairs/s), MEM -2 % (62.4 vs 63.7): Greedy only
    const KjSmemLayout a = kj_smem_layout(rp), b = kj_smem_layout(kj_fixed_profile(rp.mode));
    return memcmp(&a, &b, sizeof a) == 0 && rp.max_len == 152 && rp.kept_cap_smem == KJ_KEPT_SMEM_FIXED;
}
static int configure_launch(kj_ctx* c, KjRunParams& rp, size_t& smem, int& grid, bool verbose) {
#ifdef KJ_STAGE
    rp.stage = getenv("KJ_NO_STAGE") ? 0u : 1u;            // build with -DKJ_STAGE: bulk-copy staging of the bases (A/B; not the default, see the kernel)
#else
    rp.stage = 0u;
#endif
    KjSmemLayout L = kj_smem_layout(rp);
    const size_t head = kj_align((uint32_t)sizeof(KjCtaShared), 16);
    smem = head + (size_t)KJ_WARPS_PER_CTA * L.total;
    size_t limit = KJ_SMEM_WS_LIMIT;
    if (const char* v = getenv("KJ_WS_LIMIT_KB")) { long x = atol(v); if (x >= 0 && x <= 227) limit = (size_t)x * 1024u; }    // tuning hook (A/B of the switch point)
    rp.ws_global = smem > limit ? 1u : 0u;
    if (rp.ws_global) { smem = head; rp.stage = 0; }
    const int cfg = rp.mode * 2 + (int)rp.ws_global + ((!verbose && kj_use_fixed(rp)) ? 4 : 0) + (verbose ? 8 : 0);
    // The max-dynamic-shared-memory attribute belongs to the kernel instantiation on the device, not to a context: several contexts
    // (or batches with different read lengths) share it, so it is only ever raised (process-wide table), never lowered.
    static std::mutex attr_mu; static size_t attr_set[64][32];
    const bool fixed = !verbose && kj_use_fixed(rp);
    const int inst = rp.mode * 4 + (c->H.wide ? 2 : 0) + (int)rp.ws_global + (fixed ? 8 : 0) + (verbose ? 16 : 0);
    bool raise = false;
    { std::lock_guard<std::mutex> lk(attr_mu); if (c->device < 64 && smem > attr_set[c->device][inst]) { attr_set[c->device][inst] = smem; raise = true; } else if (c->device >= 64) raise = true; }
    if (!raise && smem == c->smem_bytes && c->grid > 0 && c->cfg_mode == cfg) { grid = c->grid; return KJ_OK; }
    int per_sm = 0;
#define KJ_CFGR(M, T, G, F, V, R) { if (raise) CK(cudaFuncSetAttribute(kj_classify_kernel<M, T, G, F, V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kj_classify_kernel<M, T, G, F, V, R>, KJ_WARPS_PER_CTA * 32, smem)); }
#define KJ_CFG(M, T, G, F, V) KJ_CFGR(M, T, G, F, V, 0)
#define KJ_CFG2(M, T) { if (verbose) { if (rp.ws_global) KJ_CFG(M, T, true, false, true) else KJ_CFG(M, T, false, false, true) } \
                        else if (rp.ws_global) KJ_CFG(M, T, true, false, false) else if (fixed) KJ_CFG(M, T, false, true, false) else KJ_CFG(M, T, false, false, false) }
    // the Greedy front-end / search pair (launch()): the grid is the search kernel's, the front end needs no more than that
#define KJ_CFGS(T) { if (fixed) { KJ_CFGR(1, T, false, true, false, 1) KJ_CFGR(1, T, false, true, false, 2) } else { KJ_CFGR(1, T, false, false, false, 1) KJ_CFGR(1, T, false, false, false, 2) } }
    if (rp.mode == 0) { if (c->H.wide) KJ_CFG2(0, uint64_t) else KJ_CFG2(0, uint32_t) }
    else {
        if (c->H.wide) KJ_CFG2(1, uint64_t) else KJ_CFG2(1, uint32_t)
        if (!rp.ws_global && !verbose && !getenv("KJ_NO_SPLIT")) { if (c->H.wide) KJ_CFGS(uint64_t) else KJ_CFGS(uint32_t) }
    }
#undef KJ_CFGS
#undef KJ_CFG2
#undef KJ_CFG
#undef KJ_CFGR
    c->cfg_mode = cfg;
    if (per_sm < 1) { kj_err() = "kernel does not fit on an SM"; return KJ_ERR_UNSUPPORTED; }
    grid = c->sm_count * per_sm;             // persistent grid: a whole number of CTAs per SM
    return KJ_OK;
}

static int ensure_scratch(kj_ctx* c, const KjRunParams& rp, int grid) {
    size_t warps = (size_t)grid * KJ_WARPS_PER_CTA * 2;   // two pipeline slots may be in flight
    size_t need = warps * rp.scratch_entries * sizeof(KjKept);
    if (need > c->spill_bytes) { if (c->d_spill) cudaFree(c->d_spill); c->d_spill = nullptr; CK(cudaMalloc((void**)&c->d_spill, need)); c->spill_bytes = need; }
    size_t gneed = warps * (size_t)kj_greedy_scratch_bytes(rp);
    if (gneed > c->gscratch_bytes_total) { if (c->d_gscratch) cudaFree(c->d_gscratch); c->d_gscratch = nullptr; CK(cudaMalloc((void**)&c->d_gscratch, gneed)); c->gscratch_bytes_total = gneed; }
    size_t wneed = rp.ws_global ? warps * (size_t)kj_smem_layout(rp).total : 0;
    if (wneed > c->ws_bytes) { if (c->d_ws) cudaFree(c->d_ws); c->d_ws = nullptr; CK(cudaMalloc((void**)&c->d_ws, wneed)); c->ws_bytes = wneed; }
    return KJ_OK;
}

// E-value gate: break points of the minimal passing score (kj_build_evalue_breaks), rebuilt when the parameters change
static int upload_evalue_breaks(kj_ctx* c) {
    if (c->d_evbreaks) { cudaFree(c->d_evbreaks); c->d_evbreaks = nullptr; } c->n_evbreaks = 0;
    std::vector<double> br; int rc = kj_build_evalue_breaks(c->params, c->H.db_length, br); if (rc) return rc;
    if (br.empty()) return KJ_OK;
    CK(cudaMalloc((void**)&c->d_evbreaks, br.size() * sizeof(double)));
    CK(cudaMemcpy(c->d_evbreaks, br.data(), br.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->n_evbreaks = (uint32_t)br.size();
    return KJ_OK;
}

static thread_local bool kj_transient_ctx = false;      // the base context of kj_create_scaled lives for the construction only: no second k-mer table
static int new_ctx(kj_ctx** out, int device, const kj_params* params) {
    int rc = kj_check_params(*params); if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { kj_err() = "no CUDA device available (this library has no CPU fallback)"; return KJ_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { kj_err() = "device ordinal out of range"; return KJ_ERR_ARG; }
    CK(cudaSetDevice(device));
    kj_ctx* c = new kj_ctx(); c->device = device; c->params = *params;
    std::unique_ptr<kj_ctx, void (*)(kj_ctx*)> guard(c, kj_destroy);
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device)); c->sm_count = prop.multiProcessorCount;
    CK(cudaMalloc((void**)&c->d_err, sizeof(uint32_t))); CK(cudaMemset(c->d_err, 0, sizeof(uint32_t)));
    CK(cudaMalloc((void**)&c->d_quirk, sizeof c->H.quirk_d)); CK(cudaMemset(c->d_quirk, 0, sizeof c->H.quirk_d));
    *out = guard.release(); return KJ_OK;
}
// device descriptor from the context's device arrays + meta data
static int upload_descriptor(kj_ctx* c) {
    KjHostIndex& H = c->H; KjDevIndex& D = c->dix; memset(&D, 0, sizeof D);
    D.rank = (const uint64_t*)c->d_rank; D.nb = H.nb; D.letters = (const uint64_t*)c->d_letters; D.bwtlen = H.bwtlen; D.alen = H.alen;
    for (int a = 0; a <= H.alen; a++) D.C[a] = H.C[a];
    for (int a = 0; a < H.alen; a++) D.rank_base[a] = D.rank + (uint64_t)a * H.nb * kj_rank_words(H.wide);
    D.sa_acc = (const uint32_t*)c->d_sa_acc; D.seq_acc = (const uint32_t*)c->d_seq_acc;
    D.sa_tax = (const uint32_t*)c->d_sa_tax; D.seq_tax = (const uint32_t*)c->d_seq_tax; D.sa_check = H.sa_check; D.sa_exp = H.sa_exp; D.sa_bias = H.sa_bias;
    D.n_sa = c->n_sa; D.nseq = H.nseq;
    D.tax_parent = (const uint32_t*)c->d_tax_parent; D.tax_depth = (const uint32_t*)c->d_tax_depth; D.tax_id = (const uint64_t*)c->d_tax_id; D.n_tax = (uint32_t)H.tax_id.size();
    D.lnfact = (const double*)c->d_lnfact; D.n_lnfact = (int)H.lnfact.size(); D.kmer = H.kmer_k ? c->d_kmer : nullptr; D.kmer_k = H.kmer_k; D.wide = H.wide; D.tables = c->d_tables;
    D.quirk_lo = H.quirk_lo; D.mono = (H.quirk_lo == ~0ull && !getenv("KJ_NOMONO")) ? 1 : 0; D.quirk_d = c->d_quirk;      // KJ_NOMONO: developer hook (A/B of the chain bounds)
    if (!c->d_ix) CK(cudaMalloc((void**)&c->d_ix, sizeof(KjDevIndex)));
    CK(cudaMemcpy(c->d_ix, &D, sizeof(KjDevIndex), cudaMemcpyHostToDevice));
    if (c->d_kmer_mem && c->kmer_k_mem) {
        KjDevIndex M = D; M.kmer = c->d_kmer_mem; M.kmer_k = c->kmer_k_mem;
        if (!c->d_ix_mem) CK(cudaMalloc((void**)&c->d_ix_mem, sizeof(KjDevIndex)));
        CK(cudaMemcpy(c->d_ix_mem, &M, sizeof(KjDevIndex), cudaMemcpyHostToDevice));
    }
    return KJ_OK;
}
static int upload_small(kj_ctx* c, uint64_t& tot) {
    KjHostIndex& H = c->H; int rc;
    if ((rc = upload(H.tax_parent, &c->d_tax_parent, tot)) || (rc = upload(H.tax_depth, &c->d_tax_depth, tot)) || (rc = upload(H.tax_id, &c->d_tax_id, tot)) || (rc = upload(H.lnfact, &c->d_lnfact, tot))) return rc;
    CK(cudaMalloc((void**)&c->d_tables, sizeof(KjTables))); CK(cudaMemcpy(c->d_tables, &H.tables, sizeof(KjTables), cudaMemcpyHostToDevice));
    return KJ_OK;
}
static int finish_ctx(kj_ctx* c, uint64_t tot) {
    KjHostIndex& H = c->H; int rc;
    CK(cudaMemcpy(c->d_quirk, H.quirk_d, sizeof H.quirk_d, cudaMemcpyHostToDevice));
    if ((rc = upload_descriptor(c))) return rc;
    c->index_bytes = tot;
    CK(cudaMalloc((void**)&c->d_counter, 4 * sizeof(unsigned long long))); CK(cudaMalloc((void**)&c->d_maxlen, 2 * sizeof(unsigned int)));       // counters: [slot] classify / search, [2 + slot] front end
    for (int s = 0; s < 2; s++) {
        CK(cudaStreamCreateWithFlags(&c->stream[s], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->fstream[s], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming));
        for (int b = 0; b < 2; b++) { CK(cudaEventCreateWithFlags(&c->ev_f[s][b], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_s[s][b], cudaEventDisableTiming)); }
    }
    CK(cudaEventCreate(&c->ev_a)); CK(cudaEventCreate(&c->ev_b));
    if ((rc = upload_evalue_breaks(c))) return rc;
    c->n_counts = (uint32_t)H.tax_id.size() + 1u; c->n_present = H.n_present;
    CK(cudaMalloc((void**)&c->d_counts, (size_t)c->n_counts * 8)); CK(cudaMalloc((void**)&c->d_counts_pending, (size_t)c->n_counts * 8));
    CK(cudaMemset(c->d_counts, 0, (size_t)c->n_counts * 8)); CK(cudaMemset(c->d_counts_pending, 0, (size_t)c->n_counts * 8));
    return KJ_OK;
}

// kj_create_from_native (and kj_create with KJ_HOST_BUILD=1): `fill` produces the host arrays of the device layout, the rest uploads them
template <class Fill> static int create_ctx(kj_ctx** out, int device, const kj_params* params, Fill fill) {
    kj_ctx* c = nullptr; int rc = new_ctx(&c, device, params); if (rc) return rc;
    std::unique_ptr<kj_ctx, void (*)(kj_ctx*)> guard(c, kj_destroy);      // every early return below releases what was allocated so far
    rc = fill(c->H); if (rc) return rc;
    KjHostIndex& H = c->H; uint64_t tot = 0;
    // one guard entry: the reference's header counts one sampled row less than kaiju-mkbwt writes (suffixArray.c:160 vs 206-216), so the last
    // sampled row of an index has no entry; the reference reads past its array there, the device reads "no taxon"
    c->n_sa = H.sa_tax.size(); H.sa_tax.push_back(KJ_TAX_BAD);
    if (!H.seq_acc.empty()) { H.sa_acc.push_back(0xffffffffu); if ((rc = upload(H.sa_acc, &c->d_sa_acc, tot)) || (rc = upload(H.seq_acc, &c->d_seq_acc, tot))) return rc; }
    if ((rc = upload(H.rank, &c->d_rank, tot)) || (rc = upload(H.letters, &c->d_letters, tot)) || (rc = upload(H.sa_tax, &c->d_sa_tax, tot)) ||
        (rc = upload(H.seq_tax, &c->d_seq_tax, tot)) || (rc = upload_small(c, tot)) || (rc = (H.wide ? upload(H.kmer, &c->d_kmer, tot) : upload(H.kmer32, &c->d_kmer, tot)))) return rc;
    // host copies of the big arrays are no longer needed
    std::vector<uint64_t>().swap(H.rank); std::vector<uint64_t>().swap(H.letters); std::vector<uint32_t>().swap(H.sa_tax); std::vector<KjKmer>().swap(H.kmer); std::vector<KjKmer32>().swap(H.kmer32);
    if ((rc = finish_ctx(c, tot))) return rc;
    *out = guard.release(); return KJ_OK;
}

#include "kj_build.h"
// kj_create / kj_create_scaled: the large arrays are built on the device from the raw BWT bytes and suffix-array samples (kj_build.h)
static int create_ctx_device(kj_ctx** out, int device, const kj_params* params, const kj_index_view& v, const kj_taxonomy_view& t, uint32_t copies, const kj_ctx* base) {
    kj_ctx* c = nullptr; int rc = new_ctx(&c, device, params); if (rc) return rc;
    std::unique_ptr<kj_ctx, void (*)(kj_ctx*)> guard(c, kj_destroy);
    const auto t0 = std::chrono::steady_clock::now();
    uint8_t lcode[256]; rc = kj_build_host_meta(v, t, copies, c->H, lcode); if (rc) return rc;
    uint64_t tot = 0;
    if ((rc = upload_small(c, tot)) || (rc = kj_device_build(c, v, lcode, copies, base, tot))) return rc;
    c->H.kmer_k = 0;
    if ((rc = upload_descriptor(c))) return rc;
    { const char* ek = getenv("KJ_KMER_K"); if ((rc = kj_device_build_kmer(c, ek ? atoi(ek) : kj_default_kmer_k(c->H.bwtlen), tot))) return rc; }
    // MEM runs 4 % faster with the intervals of all 20^7 7-mers (10.2 GB below 2^32 rows): one look-up replaces the first LF step of every chain, the one
    // with all 32 lanes alive.  Greedy loses 2 % with it (its seeds rarely get that far), so it keeps the 6-mer table: two descriptors, one index.
    // Only where HBM is plentiful: narrow indexes, and the two level buffers of the construction (41 GB) must fit next to the index.
    if (!c->H.wide && c->H.kmer_k == 6 && !kj_transient_ctx && !getenv("KJ_KMER_K") && !getenv("KJ_NO_KMER7")) {
        size_t fr = 0, to = 0; CK(cudaMemGetInfo(&fr, &to));
        if ((double)fr > 1.28e9 * (2.0 * sizeof(KjKmer) + sizeof(KjKmer32)) + 16e9 && (rc = kj_device_build_kmer(c, 7, tot, &c->d_kmer_mem, &c->kmer_k_mem))) return rc;
    }
    std::vector<uint32_t>().swap(c->H.seq_tax);
    if ((rc = finish_ctx(c, tot))) return rc;
    CK(cudaDeviceSynchronize());
    c->build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = guard.release(); return KJ_OK;
}

extern "C" int kj_create(kj_ctx** out, int device, const kj_params* params, const kj_index_view* index, const kj_taxonomy_view* taxonomy) {
    if (!out || !params || !index || !taxonomy) { kj_err() = "kj_create: null argument"; return KJ_ERR_ARG; }
    if (getenv("KJ_HOST_BUILD")) return create_ctx(out, device, params, [&](KjHostIndex& H) { return kj_build_host_index(*index, *taxonomy, H); });   // developer hook: host transcoder + upload
    return create_ctx_device(out, device, params, *index, *taxonomy, 1, nullptr);
}
extern "C" int kj_create_scaled(kj_ctx** out, int device, const kj_params* params, const kj_index_view* index, const kj_taxonomy_view* taxonomy, uint32_t copies) {
    if (!out || !params || !index || !taxonomy) { kj_err() = "kj_create_scaled: null argument"; return KJ_ERR_ARG; }
    if (copies <= 1) return kj_create(out, device, params, index, taxonomy);
    kj_ctx* base = nullptr; kj_transient_ctx = true; int rc = create_ctx_device(&base, device, params, *index, *taxonomy, 1, nullptr); kj_transient_ctx = false; if (rc) return rc;
    rc = create_ctx_device(out, device, params, *index, *taxonomy, copies, base);
    kj_destroy(base);
    return rc;
}
extern "C" double kj_index_build_ms(const kj_ctx* c) { return c ? c->build_ms : 0.0; }
// test hook: checksums of the index arrays as they sit in HBM (rank, letters, sa_tax, seq_tax, kmer), to compare the device construction
// with the host transcoder array for array
extern "C" int kj_debug_index_checksums(kj_ctx* c, uint64_t out[8]) {
    if (!c || !out) return KJ_ERR_ARG; CK(cudaSetDevice(c->device)); CK(cudaDeviceSynchronize());
    const KjHostIndex& H = c->H; memset(out, 0, 64);
    const size_t sz[5] = {(size_t)H.alen * H.nb * kj_rank_words(H.wide) * 8, (size_t)(H.bwtlen / KJ_LETTERS_PER_WORD + 2) * 8, (size_t)c->n_sa * 4, (size_t)H.nseq * 4,
                          H.kmer_k ? (size_t)pow(20.0, H.kmer_k) * (H.wide ? sizeof(KjKmer) : sizeof(KjKmer32)) : 0};
    const void* ptr[5] = {c->d_rank, c->d_letters, c->d_sa_tax, c->d_seq_tax, c->d_kmer};
    for (int i = 0; i < 5; i++) { std::vector<uint8_t> h(sz[i]); if (sz[i]) CK(cudaMemcpy(h.data(), ptr[i], sz[i], cudaMemcpyDeviceToHost)); out[i] = kj_mix_bytes(0x6b616a75ull + i, h.data(), h.size()); }
    out[5] = H.bwtlen; out[6] = (uint64_t)H.wide; out[7] = c->n_sa;
    return KJ_OK;
}

// device-native index file: written once from the reference's .fmi + nodes.dmp, loaded without the transcode
extern "C" int kj_native_index_write(const kj_index_view* index, const kj_taxonomy_view* taxonomy, const char* path) {
    if (!index || !taxonomy || !path) { kj_err() = "kj_native_index_write: null argument"; return KJ_ERR_ARG; }
    KjHostIndex H; int rc = kj_build_host_index(*index, *taxonomy, H); if (rc) return rc;
    return kj_host_index_write(H, path);
}
extern "C" int kj_create_from_native(kj_ctx** out, int device, const kj_params* params, const char* path) {
    if (!out || !params || !path) { kj_err() = "kj_create_from_native: null argument"; return KJ_ERR_ARG; }
    return create_ctx(out, device, params, [&](KjHostIndex& H) { return kj_host_index_read(path, H); });
}

extern "C" int kj_set_params(kj_ctx* c, const kj_params* p) {
    if (!c || !p) { kj_err() = "kj_set_params: null argument"; return KJ_ERR_ARG; }
    int rc = kj_check_params(*p); if (rc) return rc;
    CK(cudaSetDevice(c->device));
    c->params = *p;
    // the record buffers of the two-kernel Greedy path (up to 32 GB) go back when the context leaves Greedy mode
    if (c->params.mode != 1 && c->d_prep) { CK(cudaDeviceSynchronize()); cudaFree(c->d_prep); c->d_prep = nullptr; c->prep_bytes = 0; }
    return upload_evalue_breaks(c);
}

extern "C" void kj_destroy(kj_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    kj_files_state_free(c->files); c->files = nullptr;
    if (c->d_ix_mem) cudaFree(c->d_ix_mem); if (c->d_kmer_mem) cudaFree(c->d_kmer_mem); if (c->d_prep) cudaFree(c->d_prep);
    void* ptrs[] = {c->d_rank, c->d_letters, c->d_sa_tax, c->d_seq_tax, c->d_tax_parent, c->d_tax_depth, c->d_tax_id, c->d_lnfact, c->d_kmer, c->d_tables, c->d_ix,
                    c->d_counter, c->d_err, c->d_maxlen, c->d_spill, c->d_gscratch, c->d_evbreaks, c->d_ws, c->d_counts, c->d_counts_pending, c->d_quirk, c->d_tax[0], c->d_tax[1], c->d_best[0], c->d_best[1],
                    c->d_seq[0][0], c->d_seq[0][1], c->d_seq[1][0], c->d_seq[1][1], c->d_off[0][0], c->d_off[0][1], c->d_off[1][0], c->d_off[1][1],
                    c->d_ids[0], c->d_ids[1], c->d_nids[0], c->d_nids[1], c->d_sa_acc, c->d_seq_acc, c->d_acc[0], c->d_acc[1], c->d_nacc[0], c->d_nacc[1],
                    c->d_frag[0], c->d_frag[1], c->d_fraglen[0], c->d_fraglen[1]};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int s = 0; s < 2; s++) { if (c->stream[s]) cudaStreamDestroy(c->stream[s]); if (c->fstream[s]) cudaStreamDestroy(c->fstream[s]); if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
                                  for (int b = 0; b < 2; b++) { if (c->ev_f[s][b]) cudaEventDestroy(c->ev_f[s][b]); if (c->ev_s[s][b]) cudaEventDestroy(c->ev_s[s][b]); } }
    if (c->ev_a) cudaEventDestroy(c->ev_a); if (c->ev_b) cudaEventDestroy(c->ev_b);
    delete c;
}

// one launch over reads [0,n) whose sequences/offsets are resident on the device
static int launch(kj_ctx* c, int slot, const uint8_t* d_seq1, const uint64_t* d_off1, const uint8_t* d_seq2, const uint64_t* d_off2, uint64_t base1, uint64_t base2,
                  uint64_t n, uint32_t max1, uint32_t max2, uint64_t* d_tax, uint32_t* d_best, cudaStream_t st, bool time_it,
                  uint64_t* d_ids = nullptr, uint8_t* d_nids = nullptr, unsigned long long* d_count_dst = nullptr, uint32_t* d_compact = nullptr,
                  uint32_t* d_acc = nullptr, uint8_t* d_nacc = nullptr, char* d_frag = nullptr, uint32_t frag_stride = 0, uint32_t* d_fraglen = nullptr) {
    if (c->params.input_is_protein) {
        if (d_seq2) { kj_err() = "protein input only supports one input (kaiju.cpp:201)"; return KJ_ERR_ARG; }
        if (max1 > KJ_MAX_PROTEIN_LEN) { kj_err() = "protein read longer than KJ_MAX_PROTEIN_LEN (5461 residues) is not supported"; return KJ_ERR_UNSUPPORTED; }
    } else if (max1 > KJ_MAX_READ_LEN || max2 > KJ_MAX_READ_LEN) { kj_err() = "read longer than KJ_MAX_READ_LEN (16383 bases) is not supported"; return KJ_ERR_UNSUPPORTED; }
    KjRunParams rp; kj_fill_run_params(c->params, std::max(max1, max2), rp);
    rp.ev_breaks = c->d_evbreaks; rp.n_ev_breaks = c->n_evbreaks;
    rp.variant_cap *= c->variant_boost;
    const bool verbose = d_ids || d_acc || d_frag;
    size_t smem; int grid; int rc = configure_launch(c, rp, smem, grid, verbose); if (rc) return rc;
    rc = ensure_scratch(c, rp, grid); if (rc) return rc;
    c->grid = grid; c->smem_bytes = smem;
    if (time_it) CK(cudaEventRecord(c->ev_a, st));
    const size_t warps = (size_t)grid * KJ_WARPS_PER_CTA;
    const KjSmemLayout lay = kj_smem_layout(rp);
    const bool fixed = !verbose && kj_use_fixed(rp);
    const KjDevIndex* dix = (rp.mode == 0 && c->d_ix_mem && rp.m >= (uint32_t)c->kmer_k_mem) ? c->d_ix_mem : c->d_ix;
    // Greedy with the work space in shared memory: front-end kernel + search kernel over sub-batches (the records of a sub-batch live in d_prep)
    const bool split = rp.mode == 1 && !rp.ws_global && !verbose && !getenv("KJ_NO_SPLIT");
    const uint32_t pstride = split ? kj_prep_stride(rp) : 0u; uint64_t sub = n;
    if (split) {
        // records of one sub-batch per buffer; two buffers per slot (the front end of sub-batch b+1 runs in the tail of the search of sub-batch b).
        // Up to 8 GB per buffer where HBM is plentiful: fewer, longer search launches (750 k vs 3 M pairs per launch: 17.9 vs 18.3 M pairs/s)
        uint64_t per_buf = c->prep_bytes / 4;
        if (per_buf < (8ull << 30) && per_buf < n * (uint64_t)pstride) {       // (re)allocate: what this launch needs, at least 1 GB, at most 8 GB or 1/16 of the free memory
            size_t fr = 0, to = 0; CK(cudaMemGetInfo(&fr, &to)); fr += c->prep_bytes;
            const uint64_t lim = std::max<uint64_t>(64ull << 20, std::min<uint64_t>(8ull << 30, fr / 16));
            const uint64_t want = std::min<uint64_t>(lim, std::max<uint64_t>(1ull << 30, n * (uint64_t)pstride + (n * (uint64_t)pstride) / 4));
            if (want > per_buf) {
                if (c->d_prep) cudaFree(c->d_prep); c->d_prep = nullptr; c->prep_bytes = 0;
                CK(cudaMalloc((void**)&c->d_prep, (size_t)want * 4)); c->prep_bytes = (size_t)want * 4; per_buf = want;
            }
        }
        sub = std::max<uint64_t>(1024, per_buf / pstride);
        if (const char* v = getenv("KJ_SPLIT_SUB")) { const long long x = atoll(v); if (x >= 1024 && (uint64_t)x < sub) sub = (uint64_t)x; }
        sub = std::min(sub, n);
    }
#define KJ_ARGS(B0, B1) dix, rp, lay, d_seq1, d_off1, d_seq2, d_off2, base1, base2, (uint64_t)(B1), d_tax, d_best, d_ids, d_nids, d_compact, \
            ctr, c->d_spill + (size_t)slot * warps * rp.scratch_entries, \
            c->d_gscratch + (size_t)slot * warps * kj_greedy_scratch_bytes(rp), kj_greedy_scratch_bytes(rp), \
            rp.ws_global ? c->d_ws + (size_t)slot * warps * kj_smem_layout(rp).total : nullptr, d_count_dst, c->d_err, d_acc, d_nacc, d_frag, frag_stride, d_fraglen, pbuf, pstride, (uint64_t)(B0)
#define KJ_LAUNCH3(M, T, G, F, V, R, B0, B1) kj_classify_kernel<M, T, G, F, V, R><<<kgrid, KJ_WARPS_PER_CTA * 32, smem, kst>>>(KJ_ARGS(B0, B1))
#define KJ_LAUNCH(M, T) if (verbose) { if (rp.ws_global) KJ_LAUNCH3(M, T, true, false, true, 0, 0, n); else KJ_LAUNCH3(M, T, false, false, true, 0, 0, n); } \
                        else if (rp.ws_global) KJ_LAUNCH3(M, T, true, false, false, 0, 0, n); else if (fixed) KJ_LAUNCH3(M, T, false, true, false, 0, 0, n); else KJ_LAUNCH3(M, T, false, false, false, 0, 0, n)
#define KJ_LAUNCH_SPLIT(T, R, B0, B1) if (fixed) KJ_LAUNCH3(1, T, false, true, false, R, B0, B1); else KJ_LAUNCH3(1, T, false, false, false, R, B0, B1)
    uint8_t* pbuf = nullptr; unsigned long long* ctr = c->d_counter + slot; cudaStream_t kst = st; int kgrid = grid;
    // (A/B r2l: a search grid that leaves one CTA slot per SM to the front end of the next sub-batch loses 9 % -- 32 instead of 40 search warps per SM cost
    // more than the hidden front end (8 % of the kernel time) gives back; KJ_SPLIT_LEAVE_SLOT keeps the experiment)
    const int per_sm = grid / c->sm_count; const int sgrid = (per_sm >= 3 && getenv("KJ_SPLIT_LEAVE_SLOT")) ? c->sm_count * (per_sm - 1) : grid;
    if (split) {
        // front end on the slot's own stream, search on the caller's: F(b) -> S(b) through ev_f, S(b) -> F(b+2) (same buffer) through ev_s
        cudaStream_t fs = c->fstream[slot]; uint8_t* base = c->d_prep + (size_t)slot * (c->prep_bytes / 2); uint64_t k = 0;
        CK(cudaEventRecord(c->ev_in[slot], st)); CK(cudaStreamWaitEvent(fs, c->ev_in[slot], 0));      // the inputs may have been produced on the caller's stream
        for (uint64_t b0 = 0; b0 < n; b0 += sub, k++) {
            const uint64_t b1 = std::min(n, b0 + sub); const int pb = (int)(k & 1);
            pbuf = base + (size_t)pb * (c->prep_bytes / 4);
            CK(cudaStreamWaitEvent(fs, c->ev_s[slot][pb], 0));                 // the search that last read this buffer (of this or an earlier launch; no-op if none)
            ctr = c->d_counter + 2 + slot; kst = fs; kgrid = grid;
            CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), fs));
            if (c->H.wide) KJ_LAUNCH_SPLIT(uint64_t, 1, b0, b1); else KJ_LAUNCH_SPLIT(uint32_t, 1, b0, b1);
            CK(cudaEventRecord(c->ev_f[slot][pb], fs));
            ctr = c->d_counter + slot; kst = st; kgrid = sgrid;
            CK(cudaStreamWaitEvent(st, c->ev_f[slot][pb], 0));
            CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
            if (c->H.wide) KJ_LAUNCH_SPLIT(uint64_t, 2, b0, b1); else KJ_LAUNCH_SPLIT(uint32_t, 2, b0, b1);
            CK(cudaEventRecord(c->ev_s[slot][pb], st));
            c->launches += 2;
        }
        c->launches--;          // (the common tail below counts one)
    } else {
        CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
        if (rp.mode == 0) { if (c->H.wide) KJ_LAUNCH(0, uint64_t); else KJ_LAUNCH(0, uint32_t); }
        else { if (c->H.wide) KJ_LAUNCH(1, uint64_t); else KJ_LAUNCH(1, uint32_t); }
    }
#undef KJ_LAUNCH
#undef KJ_LAUNCH3
#undef KJ_LAUNCH_SPLIT
#undef KJ_ARGS
    CK(cudaGetLastError());
    if (time_it) CK(cudaEventRecord(c->ev_b, st));
    c->launches++;
    return KJ_OK;
}

static int count_taxa(kj_ctx* c, const uint64_t* d_tax, uint64_t n, unsigned long long* d_dst, cudaStream_t st) {
    if (!n) return KJ_OK;
    kj_count_kernel<<<c->sm_count * 4, 256, 0, st>>>(d_tax, n, c->dix.tax_id, c->n_present, c->n_counts - 1u, d_dst);
    CK(cudaGetLastError()); c->launches++;
    return KJ_OK;
}

static int check_err_flag(kj_ctx* c) {
    uint32_t e = 0; CK(cudaMemcpy(&e, c->d_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e) {
        CK(cudaMemset(c->d_err, 0, sizeof e));
        // flag 4 = the Greedy variant ring of some read was full: the next launch gets a ring 4x as large (the reference's heap is unbounded)
        if ((e & 4u) && c->variant_boost < 256u) c->variant_boost *= 4u;
        char b[160]; snprintf(b, sizeof b, "per-read work queue overflow on the device (flags 0x%x)%s", e, (e & 4u) ? "; the variant ring was enlarged, call again" : (e & 128u) ? "; the fragment strings of a read exceed frag_stride" : ""); kj_err() = b;
        return KJ_ERR_OVERFLOW;
    }
    return KJ_OK;
}

extern "C" int kj_classify_device2(kj_ctx* c, const char* d_seq1, const uint64_t* d_off1, const char* d_seq2, const uint64_t* d_off2, uint64_t n,
                                   uint32_t max_len1, uint32_t max_len2, uint64_t* d_tax, uint32_t* d_best, uint32_t* d_compact, void* cuda_stream) {
    if (!c || !d_seq1 || !d_off1 || (!d_tax && !d_compact) || (d_seq2 && !d_off2)) { kj_err() = "kj_classify_device: null argument"; return KJ_ERR_ARG; }
    if (n == 0) return KJ_OK;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (max_len1 == 0 || (d_seq2 && max_len2 == 0)) {
        CK(cudaMemsetAsync(c->d_maxlen, 0, 2 * sizeof(unsigned int), st));
        kj_maxlen_kernel<<<256, 256, 0, st>>>(d_off1, n, c->d_maxlen); c->launches++;
        if (d_seq2) { kj_maxlen_kernel<<<256, 256, 0, st>>>(d_off2, n, c->d_maxlen + 1); c->launches++; }
        unsigned int h[2] = {0, 0}; CK(cudaMemcpyAsync(h, c->d_maxlen, sizeof h, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        max_len1 = h[0]; max_len2 = h[1];
    }
    return launch(c, 0, (const uint8_t*)d_seq1, d_off1, (const uint8_t*)d_seq2, d_off2, 0, 0, n, max_len1, max_len2, d_tax, d_best, st, true, nullptr, nullptr, nullptr, d_compact);
}

extern "C" int kj_classify_device(kj_ctx* c, const char* d_seq1, const uint64_t* d_off1, const char* d_seq2, const uint64_t* d_off2, uint64_t n,
                                  uint32_t max_len1, uint32_t max_len2, uint64_t* d_tax, uint32_t* d_best, void* cuda_stream) {
    if (!d_tax) { kj_err() = "kj_classify_device: null argument"; return KJ_ERR_ARG; }
    return kj_classify_device2(c, d_seq1, d_off1, d_seq2, d_off2, n, max_len1, max_len2, d_tax, d_best, nullptr, cuda_stream);
}

static int ensure_staging(kj_ctx* c, int slot, size_t bytes1, size_t bytes2, size_t reads) {
    size_t need[2] = {bytes1, bytes2};
    for (int m = 0; m < 2; m++) if (need[m] > c->d_seq_cap[slot][m]) {
        if (c->d_seq[slot][m]) cudaFree(c->d_seq[slot][m]); c->d_seq[slot][m] = nullptr;
        size_t cap = need[m] + need[m] / 8 + 4096; CK(cudaMalloc((void**)&c->d_seq[slot][m], cap)); c->d_seq_cap[slot][m] = cap;
    }
    if (reads > c->d_reads_cap) {
        for (int s = 0; s < 2; s++) {
            for (int m = 0; m < 2; m++) { if (c->d_off[s][m]) cudaFree(c->d_off[s][m]); c->d_off[s][m] = nullptr; CK(cudaMalloc((void**)&c->d_off[s][m], (reads + 1) * sizeof(uint64_t))); }
            if (c->d_tax[s]) cudaFree(c->d_tax[s]); if (c->d_best[s]) cudaFree(c->d_best[s]); c->d_tax[s] = nullptr; c->d_best[s] = nullptr;
            CK(cudaMalloc((void**)&c->d_tax[s], reads * sizeof(uint64_t))); CK(cudaMalloc((void**)&c->d_best[s], reads * sizeof(uint32_t)));
        }
        c->d_reads_cap = reads;
    }
    return KJ_OK;
}

static int classify_host(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                         uint64_t* taxon_out, uint32_t* best_out, uint64_t* ids_out, uint8_t* nids_out, uint32_t* d_compact = nullptr,
                         uint32_t* acc_out = nullptr, uint8_t* nacc_out = nullptr, char* frag_out = nullptr, uint32_t frag_stride = 0, uint32_t* frag_len_out = nullptr) {
    if (!c || !seq1 || !off1 || !taxon_out || (seq2 && !off2) || ((ids_out == nullptr) != (nids_out == nullptr))) { kj_err() = "kj_classify: null argument"; return KJ_ERR_ARG; }
    if (n == 0) return KJ_OK;
    CK(cudaSetDevice(c->device));
    const bool paired = seq2 != nullptr;
    // batch-wide length bounds (fixes the shared-memory carve-up for all chunks)
    uint32_t max1 = 0, max2 = 0;
    {
        unsigned nthr = std::max(1u, std::min(64u, std::thread::hardware_concurrency())); if (n < 65536) nthr = 1; std::vector<uint32_t> m1(nthr, 0), m2(nthr, 0); std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; t++) th.emplace_back([&, t] { uint64_t a = n * t / nthr, b = n * (t + 1) / nthr; uint32_t x = 0, y = 0;
            for (uint64_t i = a; i < b; i++) { uint64_t l = off1[i + 1] - off1[i]; if (l > 0xffffffffull) l = 0xffffffffull; x = std::max(x, (uint32_t)l);
                                               if (paired) { uint64_t k = off2[i + 1] - off2[i]; if (k > 0xffffffffull) k = 0xffffffffull; y = std::max(y, (uint32_t)k); } }
            m1[t] = x; m2[t] = y; });
        for (auto& x : th) x.join();
        for (unsigned t = 0; t < nthr; t++) { max1 = std::max(max1, m1[t]); max2 = std::max(max2, m2[t]); }
    }
    CK(cudaMemset(c->d_counts_pending, 0, (size_t)c->n_counts * 8));                // counts of this call: committed only if the whole call succeeds
    uint64_t chunk_reads = KJ_CHUNK_READS;
    if (const char* v = getenv("KJ_CHUNK_READS")) { long x = atol(v); if (x >= 1024 && x <= (1 << 24)) chunk_reads = (uint64_t)x; }       // tuning hook
    int rc = ensure_staging(c, 0, 0, 0, std::min<uint64_t>(n, chunk_reads)); if (rc) return rc;
    if (ids_out && c->d_ids_cap < c->d_reads_cap) {
        for (int s = 0; s < 2; s++) {
            if (c->d_ids[s]) cudaFree(c->d_ids[s]); if (c->d_nids[s]) cudaFree(c->d_nids[s]); c->d_ids[s] = nullptr; c->d_nids[s] = nullptr;
            CK(cudaMalloc((void**)&c->d_ids[s], c->d_reads_cap * KJ_MAX_IDS * sizeof(uint64_t))); CK(cudaMalloc((void**)&c->d_nids[s], c->d_reads_cap));
        }
        c->d_ids_cap = c->d_reads_cap;
    }
    if (acc_out || frag_out) {
        if (acc_out && !c->d_sa_acc) { kj_err() = "kj_classify_verbose2: the context was created without kj_index_view.seq_accession"; return KJ_ERR_UNSUPPORTED; }
        if ((acc_out && !nacc_out) || (frag_out && (!frag_len_out || frag_stride < 16))) { kj_err() = "kj_classify_verbose2: null argument"; return KJ_ERR_ARG; }
        if (c->d_v2_cap < c->d_reads_cap || c->d_frag_stride < frag_stride) {
            for (int s = 0; s < 2; s++) {
                if (c->d_acc[s]) cudaFree(c->d_acc[s]); if (c->d_nacc[s]) cudaFree(c->d_nacc[s]); if (c->d_frag[s]) cudaFree(c->d_frag[s]); if (c->d_fraglen[s]) cudaFree(c->d_fraglen[s]);
                c->d_acc[s] = nullptr; c->d_nacc[s] = nullptr; c->d_frag[s] = nullptr; c->d_fraglen[s] = nullptr;
                CK(cudaMalloc((void**)&c->d_acc[s], c->d_reads_cap * KJ_MAX_MATCH_ACC * 4)); CK(cudaMalloc((void**)&c->d_nacc[s], c->d_reads_cap));
                CK(cudaMalloc((void**)&c->d_frag[s], c->d_reads_cap * (size_t)frag_stride)); CK(cudaMalloc((void**)&c->d_fraglen[s], c->d_reads_cap * 4));
            }
            c->d_v2_cap = c->d_reads_cap; c->d_frag_stride = frag_stride;
        }
    }
    // software pipeline over chunks: H2D + kernel + D2H of chunk k on stream k&1 overlap with chunk k+1
    const bool trace = getenv("KJ_TRACE") != nullptr;            // developer hook: per-chunk timeline on stderr
    struct Tr { cudaEvent_t e[4]; double host_ms; uint64_t cnt; }; std::vector<Tr> tr; cudaEvent_t tr0 = nullptr; const auto th0 = std::chrono::steady_clock::now();
    if (trace) { cudaEventCreate(&tr0); cudaEventRecord(tr0, c->stream[0]); }
    for (uint64_t start = 0, k = 0, cnt = 0; start < n; start += cnt, k++) {
        const int s = (int)(k & 1); cudaStream_t st = c->stream[s];
        cnt = std::min<uint64_t>(chunk_reads, n - start);
        // long reads: bound the bases per chunk as well (at least one read)
        {
            const uint64_t* e1p = std::upper_bound(off1 + start + 1, off1 + start + cnt + 1, off1[start] + KJ_CHUNK_BYTES);
            uint64_t lim = std::max<uint64_t>(1, (uint64_t)(e1p - (off1 + start + 1)));
            if (paired) { const uint64_t* e2p = std::upper_bound(off2 + start + 1, off2 + start + cnt + 1, off2[start] + KJ_CHUNK_BYTES); lim = std::min(lim, std::max<uint64_t>(1, (uint64_t)(e2p - (off2 + start + 1)))); }
            cnt = std::min(cnt, lim);
        }
        CK(cudaStreamSynchronize(st));                      // slot s free again (its previous D2H has landed)
        if (trace) { Tr t; for (auto& e : t.e) cudaEventCreate(&e); t.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count(); t.cnt = 0; tr.push_back(t); cudaEventRecord(tr.back().e[0], st); }
        const uint64_t b1 = off1[start], e1 = off1[start + cnt], b2 = paired ? off2[start] : 0, e2 = paired ? off2[start + cnt] : 0;
        rc = ensure_staging(c, s, (size_t)(e1 - b1), (size_t)(e2 - b2), c->d_reads_cap); if (rc) return rc;
        CK(cudaMemcpyAsync(c->d_seq[s][0], seq1 + b1, (size_t)(e1 - b1), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->d_off[s][0], off1 + start, (cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if (paired) {
            CK(cudaMemcpyAsync(c->d_seq[s][1], seq2 + b2, (size_t)(e2 - b2), cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(c->d_off[s][1], off2 + start, (cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        }
        if (trace) { tr.back().cnt = cnt; cudaEventRecord(tr.back().e[1], st); }
        rc = launch(c, s, c->d_seq[s][0], c->d_off[s][0], paired ? c->d_seq[s][1] : nullptr, paired ? c->d_off[s][1] : nullptr, b1, b2, cnt, max1, max2,
                    c->d_tax[s], best_out ? c->d_best[s] : nullptr, st, true, ids_out ? c->d_ids[s] : nullptr, ids_out ? c->d_nids[s] : nullptr, c->d_counts_pending, d_compact ? d_compact + start : nullptr,
                    acc_out ? c->d_acc[s] : nullptr, acc_out ? c->d_nacc[s] : nullptr, frag_out ? c->d_frag[s] : nullptr, frag_stride, frag_out ? c->d_fraglen[s] : nullptr);
        if (rc) return rc;
        if (trace) cudaEventRecord(tr.back().e[2], st);
        CK(cudaMemcpyAsync(taxon_out + start, c->d_tax[s], cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        if (best_out) CK(cudaMemcpyAsync(best_out + start, c->d_best[s], cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (ids_out) {
            CK(cudaMemcpyAsync(ids_out + start * KJ_MAX_IDS, c->d_ids[s], cnt * KJ_MAX_IDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(nids_out + start, c->d_nids[s], cnt, cudaMemcpyDeviceToHost, st));
        }
        if (acc_out) {
            CK(cudaMemcpyAsync(acc_out + start * KJ_MAX_MATCH_ACC, c->d_acc[s], cnt * KJ_MAX_MATCH_ACC * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(nacc_out + start, c->d_nacc[s], cnt, cudaMemcpyDeviceToHost, st));
        }
        if (frag_out) {
            CK(cudaMemcpyAsync(frag_out + start * (size_t)frag_stride, c->d_frag[s], cnt * (size_t)frag_stride, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(frag_len_out + start, c->d_fraglen[s], cnt * 4, cudaMemcpyDeviceToHost, st));
        }
    }
    if (trace && !tr.empty()) cudaEventRecord(tr.back().e[3], c->stream[(tr.size() - 1) & 1]);
    CK(cudaStreamSynchronize(c->stream[0])); CK(cudaStreamSynchronize(c->stream[1]));
    if (trace) {
        fprintf(stderr, "KJ_TRACE chunk reads host_issue_ms h2d_start h2d_end kernel_end (ms since call start, device)\n");
        for (size_t k = 0; k < tr.size(); k++) { float a = 0, b = 0, d = 0; cudaEventElapsedTime(&a, tr0, tr[k].e[0]); cudaEventElapsedTime(&b, tr0, tr[k].e[1]); cudaEventElapsedTime(&d, tr0, tr[k].e[2]);
            fprintf(stderr, "KJ_TRACE %zu %llu %.2f %.2f %.2f %.2f\n", k, (unsigned long long)tr[k].cnt, tr[k].host_ms, a, b, d); }
        fprintf(stderr, "KJ_TRACE end host %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count());
    }
    rc = check_err_flag(c);
    // the per-taxon counts of this call become visible only if the whole call succeeded (a repeated call must not count twice)
    if (rc) { CK(cudaMemset(c->d_counts_pending, 0, (size_t)c->n_counts * 8)); return rc; }
    kj_count_commit<<<c->sm_count, 256, 0, c->stream[0]>>>(c->d_counts, c->d_counts_pending, c->n_counts); c->launches++;
    CK(cudaStreamSynchronize(c->stream[0]));
    return KJ_OK;
}

// A full Greedy variant ring (flag 4) enlarges the ring for the next launch: repeat the call until it fits (bounded).
static int classify_host_retry(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                               uint64_t* taxon_out, uint32_t* best_out, uint64_t* ids_out, uint8_t* nids_out, uint32_t* d_compact = nullptr,
                               uint32_t* acc_out = nullptr, uint8_t* nacc_out = nullptr, char* frag_out = nullptr, uint32_t frag_stride = 0, uint32_t* frag_len_out = nullptr) {
    for (;;) {
        const uint32_t boost = c ? c->variant_boost : 0;
        int rc = classify_host(c, seq1, off1, seq2, off2, n, taxon_out, best_out, ids_out, nids_out, d_compact, acc_out, nacc_out, frag_out, frag_stride, frag_len_out);
        if (rc != KJ_ERR_OVERFLOW || !c || c->variant_boost == boost) return rc;
    }
}

extern "C" int kj_classify(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                           uint64_t* taxon_out, uint32_t* best_out) {
    return classify_host_retry(c, seq1, off1, seq2, off2, n, taxon_out, best_out, nullptr, nullptr);
}
extern "C" int kj_classify2(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                            uint64_t* taxon_out, uint32_t* best_out, uint32_t* d_compact_out) {
    return classify_host_retry(c, seq1, off1, seq2, off2, n, taxon_out, best_out, nullptr, nullptr, d_compact_out);
}
// In-process multi-GPU (the drop-in counterpart of the reference's `-z N` consumer threads, kaiju.cpp:250-257): contiguous shards of the batch,
// one host thread per context, results written straight into the caller's arrays.  No exchange between the GPUs: reads are independent.
extern "C" int kj_classify_multi(kj_ctx** ctxs, int n_ctx, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                                 uint64_t* taxon_out, uint32_t* best_out) {
    if (!ctxs || n_ctx < 1 || !seq1 || !off1 || !taxon_out || (seq2 && !off2)) { kj_err() = "kj_classify_multi: null argument"; return KJ_ERR_ARG; }
    for (int i = 0; i < n_ctx; i++) if (!ctxs[i]) { kj_err() = "kj_classify_multi: null context"; return KJ_ERR_ARG; }
    if (n_ctx == 1) return kj_classify(ctxs[0], seq1, off1, seq2, off2, n, taxon_out, best_out);
    std::vector<int> rc((size_t)n_ctx, KJ_OK); std::vector<std::string> msg((size_t)n_ctx); std::vector<std::thread> th;
    for (int i = 0; i < n_ctx; i++) th.emplace_back([&, i] {
        const uint64_t lo = n * (uint64_t)i / (uint64_t)n_ctx, hi = n * (uint64_t)(i + 1) / (uint64_t)n_ctx;
        if (hi > lo) rc[(size_t)i] = kj_classify(ctxs[i], seq1, off1 + lo, seq2, seq2 ? off2 + lo : nullptr, hi - lo, taxon_out + lo, best_out ? best_out + lo : nullptr);
        if (rc[(size_t)i]) msg[(size_t)i] = kj_err();
    });
    for (auto& x : th) x.join();
    for (int i = 0; i < n_ctx; i++) if (rc[(size_t)i]) { kj_err() = "device " + std::to_string(ctxs[i]->device) + ": " + msg[(size_t)i]; return rc[(size_t)i]; }
    return KJ_OK;
}
extern "C" int kj_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
extern "C" int kj_classify_verbose(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n,
                                   uint64_t* taxon_out, uint32_t* best_out, uint64_t* ids_out, uint8_t* nids_out) {
    if (!ids_out || !nids_out) { kj_err() = "kj_classify_verbose: null argument"; return KJ_ERR_ARG; }
    return classify_host_retry(c, seq1, off1, seq2, off2, n, taxon_out, best_out, ids_out, nids_out);
}

extern "C" int kj_classify_verbose2(kj_ctx* c, const char* seq1, const uint64_t* off1, const char* seq2, const uint64_t* off2, uint64_t n, uint64_t* taxon_out, uint32_t* best_out,
                                    uint64_t* ids_out, uint8_t* nids_out, uint32_t* acc_out, uint8_t* nacc_out, char* frag_out, uint32_t frag_stride, uint32_t* frag_len_out) {
    if (!ids_out || !nids_out || !best_out) { kj_err() = "kj_classify_verbose2: null argument"; return KJ_ERR_ARG; }
    return classify_host_retry(c, seq1, off1, seq2, off2, n, taxon_out, best_out, ids_out, nids_out, nullptr, acc_out, nacc_out, frag_out, frag_stride, frag_len_out);
}
extern "C" uint64_t kj_kernel_launches(const kj_ctx* c) { return c ? c->launches : 0; }
extern "C" uint64_t kj_index_bytes(const kj_ctx* c) { return c ? c->index_bytes : 0; }
extern "C" double kj_last_kernel_ms(const kj_ctx* c) {
    if (!c) return 0.0;
    float ms = 0.f; if (cudaEventSynchronize(c->ev_b) != cudaSuccess) return 0.0;
    if (cudaEventElapsedTime(&ms, c->ev_a, c->ev_b) != cudaSuccess) return 0.0;
    return (double)ms;
}
extern "C" int kj_counts_reset(kj_ctx* c) {
    if (!c) return KJ_ERR_ARG; CK(cudaSetDevice(c->device));
    CK(cudaMemset(c->d_counts, 0, (size_t)c->n_counts * 8)); return KJ_OK;
}
extern "C" uint64_t kj_counts_size(const kj_ctx* c) { return c ? c->n_counts : 0; }
extern "C" void* kj_counts_device_ptr(kj_ctx* c) { return c ? (void*)c->d_counts : nullptr; }
extern "C" int kj_counts_add_device(kj_ctx* c, const uint64_t* d_taxon, uint64_t n, void* cuda_stream) {
    if (!c || (!d_taxon && n)) { kj_err() = "kj_counts_add_device: null argument"; return KJ_ERR_ARG; }
    CK(cudaSetDevice(c->device)); return count_taxa(c, d_taxon, n, c->d_counts, (cudaStream_t)cuda_stream);
}
extern "C" int kj_counts_get(kj_ctx* c, uint64_t* taxon_ids_out, uint64_t* counts_out) {
    if (!c || !counts_out) { kj_err() = "kj_counts_get: null argument"; return KJ_ERR_ARG; }
    CK(cudaSetDevice(c->device)); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(counts_out, c->d_counts, (size_t)c->n_counts * 8, cudaMemcpyDeviceToHost));
    if (taxon_ids_out) { for (uint32_t i = 0; i + 1 < c->n_counts; i++) taxon_ids_out[i] = c->H.tax_id[i]; taxon_ids_out[c->n_counts - 1] = 0; }
    return KJ_OK;
}
extern "C" int kj_counts_table(kj_ctx* c, const char* nodes_dmp, const char* names_dmp, const char* label, const kj_table_opts* opts, const char* out_path, int append) {
    if (!c) { kj_err() = "kj_counts_table: null argument"; return KJ_ERR_ARG; }
    std::vector<uint64_t> ids(c->n_counts), cnt(c->n_counts);
    int rc = kj_counts_get(c, ids.data(), cnt.data()); if (rc) return rc;
    return kj_table_write(ids.data(), cnt.data(), c->n_counts, nodes_dmp, names_dmp, label, opts, out_path, append);
}
extern "C" int kj_check_errors(kj_ctx* c) { if (!c) return KJ_ERR_ARG; cudaSetDevice(c->device); return check_err_flag(c); }
extern "C" int kj_launch_geometry(const kj_ctx* c, int* grid, int* block, int* smem) { if (!c) return KJ_ERR_ARG; if (grid) *grid = c->grid; if (block) *block = KJ_WARPS_PER_CTA * 32; if (smem) *smem = (int)c->smem_bytes; return KJ_OK; }

#include "kj_ingest.h"
