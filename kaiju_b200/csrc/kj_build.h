// kj_build.h -- index construction on the device: from the reference's in-memory index (BWT byte codes + sampled suffix
// array, the output of kaiju-mkfmi / readFMI: mkfmi.c:63-78, fmicommon.h:104-184) straight to the device layout of kj_layout.h.
//
// The host only uploads the raw bytes; the one-hot rank records, the packed letters, the taxon-reduced suffix array and the k-mer
// interval table are produced by the kernels below (the host transcoder of kj_host.cpp stays as the writer of device-native index
// files and as the CPU reference these kernels are tested against, array for array).
//
// `copies` > 1 builds the index of the collection in which every sequence occurs `copies` times in a row (kj_create_scaled).  That
// index follows from the base index without a suffix sort: identical suffixes are ordered by sequence number, so every suffix-array
// row of the base index becomes `copies` consecutive rows, BWT'[K r + c] = BWT[r], SA'[K r + c] = (K seq + c, pos) -- pinned against
// kaiju-mkbwt/-mkfmi run on the K-fold FASTA (tests/test_scaled_index.py).  It is how a refseq_ref-scale index (2.7e10 rows) is
// brought into HBM on a box that cannot run the reference's index builder at that size within a benchmark.
// Included by kj_device.cu (one translation unit).
#pragma once

#define KJ_BLD_THREADS 256
struct KjBuildLcode { uint8_t v[256]; };

// ---- pass 1: letter counts per tile of 256 rank blocks
template <int RB>
__global__ void __launch_bounds__(KJ_BLD_THREADS) kj_bld_count(const uint8_t* __restrict__ bwt, const __grid_constant__ KjBuildLcode lc, uint64_t n, uint32_t rep, int alen,
                                                                uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t tot[KJ_MAX_ALEN];
    __shared__ uint8_t lcs[256];
    __shared__ uint8_t cnt[KJ_MAX_ALEN * KJ_BLD_THREADS];             // per-thread counters (a thread sees RB <= 192 rows)
    if (threadIdx.x < KJ_MAX_ALEN) tot[threadIdx.x] = 0;
    lcs[threadIdx.x] = lc.v[threadIdx.x];
    for (int a = 0; a < KJ_MAX_ALEN; a++) cnt[a * KJ_BLD_THREADS + threadIdx.x] = 0;
    __syncthreads();
    const uint64_t tile_rows = (uint64_t)KJ_BLD_THREADS * RB, start = (uint64_t)blockIdx.x * tile_rows;
    for (uint32_t i = threadIdx.x; i < tile_rows; i += KJ_BLD_THREADS) {
        const uint64_t row = start + i;
        if (row < n) cnt[(uint32_t)lcs[bwt[rep == 1 ? row : row / rep]] * KJ_BLD_THREADS + threadIdx.x]++;
    }
    for (int a = 0; a < alen; a++) {
        uint32_t v = cnt[a * KJ_BLD_THREADS + threadIdx.x];
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&tot[a], v);
    }
    __syncthreads();
    if (threadIdx.x < KJ_MAX_ALEN) tile_counts[(uint64_t)blockIdx.x * KJ_MAX_ALEN + threadIdx.x] = threadIdx.x < (uint32_t)alen ? tot[threadIdx.x] : 0u;
}
// exclusive scan of the tile counts per letter (one warp per letter); totals[a] = number of rows holding letter a
__global__ void kj_bld_scan_tiles(const uint32_t* __restrict__ tile_counts, uint64_t ntiles, uint64_t* __restrict__ tile_prefix, uint64_t* __restrict__ totals) {
    const uint32_t a = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t carry = 0;
    for (uint64_t t0 = 0; t0 < ntiles; t0 += 32) {
        const uint64_t t = t0 + lane; const uint64_t v = t < ntiles ? tile_counts[t * KJ_MAX_ALEN + a] : 0ull; uint64_t x = v;
        for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += o; }
        if (t < ntiles) tile_prefix[t * KJ_MAX_ALEN + a] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) totals[a] = carry;
}

// ---- pass 2: rank records.  One thread per rank block (RB rows); the tile's letters are staged through shared memory (coalesced reads),
// each thread then builds, letter by letter, the one-hot words of its block with byte-wise SIMD compares; a block-wide scan of the
// per-block letter counts gives the header counts.
template <int RB, int RW>
__global__ void __launch_bounds__(KJ_BLD_THREADS) kj_bld_records(const uint8_t* __restrict__ bwt, const __grid_constant__ KjBuildLcode lc, uint64_t n, uint32_t rep, int alen, uint64_t nb,
                                                                  const uint64_t* __restrict__ tile_prefix, const uint64_t* __restrict__ Cdev, uint64_t* __restrict__ rank) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    constexpr int STR = RB + 4;                                          // row stride of a block's letters: odd number of words -> conflict-free word reads
    uint8_t* let = sm_raw;                                               // [256][STR]
    uint32_t* wsum = (uint32_t*)(sm_raw + (size_t)KJ_BLD_THREADS * STR); // [8] warp totals
    __shared__ uint8_t lcs[256];
    lcs[threadIdx.x] = lc.v[threadIdx.x];
    __syncthreads();
    const uint64_t tile_rows = (uint64_t)KJ_BLD_THREADS * RB, start = (uint64_t)blockIdx.x * tile_rows;
    for (uint32_t i = threadIdx.x; i < tile_rows; i += KJ_BLD_THREADS) {
        const uint64_t row = start + i;
        let[(i / RB) * STR + (i % RB)] = row < n ? lcs[bwt[rep == 1 ? row : row / rep]] : (uint8_t)0xff;
    }
    __syncthreads();
    const uint64_t b = (uint64_t)blockIdx.x * KJ_BLD_THREADS + threadIdx.x;
    const uint32_t* mine = (const uint32_t*)(let + (size_t)threadIdx.x * STR);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = 0; a < alen; a++) {
        uint64_t w[RB / 64]; uint32_t c = 0;
        const uint32_t a4 = (uint32_t)a * 0x01010101u;
        #pragma unroll
        for (int q = 0; q < RB / 64; q++) {
            uint64_t bits = 0;
            #pragma unroll
            for (int x = 0; x < 16; x++) {
                const uint32_t m = __vcmpeq4(mine[q * 16 + x], a4) & 0x01010101u;
                bits |= (uint64_t)((m * 0x01020408u) >> 24) << (4 * x);
            }
            w[q] = bits; c += (uint32_t)__popcll(bits);
        }
        // exclusive scan of c over the 256 blocks of the tile
        uint32_t x = c;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += o; }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        uint32_t base = 0;
        for (uint32_t k = 0; k < wid; k++) base += wsum[k];
        __syncthreads();
        if (b < nb) {
            uint64_t hdr = Cdev[a] + tile_prefix[(uint64_t)blockIdx.x * KJ_MAX_ALEN + a] + (uint64_t)(base + x - c);
            uint64_t* rec = rank + ((uint64_t)a * nb + b) * RW;
            if (RW == 2) { *(ulonglong2*)rec = make_ulonglong2(hdr, w[0]); }
            else {
                const uint64_t p1 = (uint64_t)__popcll(w[0]), p2 = p1 + (uint64_t)__popcll(w[RB / 64 > 1 ? 1 : 0]);
                hdr = (hdr & KJ_CNT_MASK) | (p1 << KJ_P1_SHIFT) | (p2 << KJ_P2_SHIFT);
                *(ulonglong2*)rec = make_ulonglong2(hdr, w[0]);
                *(ulonglong2*)(rec + 2) = make_ulonglong2(w[RB / 64 > 1 ? 1 : 0], w[RB / 64 > 2 ? 2 : 0]);
            }
        }
    }
}
// packed letters: 12 per 64-bit word
__global__ void kj_bld_letters(const uint8_t* __restrict__ bwt, const __grid_constant__ KjBuildLcode lc, uint64_t n, uint32_t rep, uint64_t nwords, uint64_t* __restrict__ letters) {
    __shared__ uint8_t lcs[256];
    lcs[threadIdx.x] = lc.v[threadIdx.x];
    __syncthreads();
    for (uint64_t wd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wd < nwords; wd += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = 0;
        for (int t = 0; t < KJ_LETTERS_PER_WORD; t++) { const uint64_t row = wd * KJ_LETTERS_PER_WORD + t; if (row < n) v |= (uint64_t)lcs[bwt[rep == 1 ? row : row / rep]] << (5 * t); }
        letters[wd] = v;
    }
}
// sampled suffix array (big-endian packed (seq,pos), suffixArray.h:37-51) -> compact taxon of the sequence
__global__ void kj_bld_sa_tax(const uint8_t* __restrict__ sa, uint64_t n_entries, uint64_t first, int nbytes, int pbits, uint32_t nseq, const uint32_t* __restrict__ seq_tax,
                              uint32_t* __restrict__ sa_tax, uint32_t* __restrict__ err) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* p = sa + e * (uint64_t)nbytes; uint64_t val = 0;
        for (int b = 0; b < nbytes; b++) val = (val << 8) + p[b];
        const uint64_t seq = val >> pbits;
        if (seq >= nseq) { atomicOr(err, 64u); continue; }
        sa_tax[first + e] = seq_tax[seq];
    }
}
// scaled index: the sampled row k' = (e + bias') << exp of the K-fold collection is copy k' % K of base row k' / K
template <class IdxT>
__global__ void kj_bld_sa_tax_scaled(const KjDevIndex* __restrict__ base_ix, uint64_t n_entries, int64_t bias, int exp, uint64_t n_rows, uint32_t rep, uint32_t* __restrict__ sa_tax) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = (uint64_t)((int64_t)e + bias) << exp;
        sa_tax[e] = k < n_rows ? kj_sa_taxon<IdxT>(*base_ix, k / rep) : KJ_TAX_BAD;
    }
}
__global__ void kj_bld_repeat_u32(const uint32_t* __restrict__ src, uint64_t n_out, uint32_t rep, uint32_t* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i / rep];
}
// FMindex with the reference's checkpoint quirk applied (host_rank of kj_host.cpp)
template <class IdxT>
static __device__ __forceinline__ uint64_t kj_bld_rank(const KjDevIndex& ix, uint32_t c, uint64_t k) {
    uint64_t v = (uint64_t)kj_rank<IdxT>(ix, c, (IdxT)k);
    if (k >= ix.quirk_lo) v -= ix.quirk_d[c];
    return v;
}
template <class IdxT>
__global__ void kj_bld_quirk(const KjDevIndex* __restrict__ ix, uint64_t k, uint64_t* __restrict__ quirk_d) {
    const uint32_t a = threadIdx.x;
    if (a < (uint32_t)ix->alen) quirk_d[a] = (uint64_t)kj_rank<IdxT>(*ix, a, (IdxT)k) - ix->C[a];
}
// one level of the k-mer interval table: nxt[i*20 + a] = interval of cur[i] extended to the left by letter a+1
template <class IdxT>
__global__ void kj_bld_kmer_level(const KjDevIndex* __restrict__ ix, const KjKmer* __restrict__ cur, uint64_t n_cur, KjKmer* __restrict__ nxt) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_cur * 20u; t += (uint64_t)gridDim.x * blockDim.x) {
        const KjKmer iv = cur[t / 20u]; const uint32_t a = (uint32_t)(t % 20u); KjKmer o; o.lo = 0; o.hi = 0;
        if (iv.lo < iv.hi) { o.lo = kj_bld_rank<IdxT>(*ix, a + 1u, iv.lo); o.hi = kj_bld_rank<IdxT>(*ix, a + 1u, iv.hi); if (o.lo >= o.hi) { o.lo = 0; o.hi = 0; } }
        nxt[t] = o;
    }
}
__global__ void kj_bld_kmer_narrow(const KjKmer* __restrict__ src, uint64_t n, KjKmer32* __restrict__ dst) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) { dst[t].lo = (uint32_t)src[t].lo; dst[t].hi = (uint32_t)src[t].hi; }
}

// chunked upload of a pageable host array
static int kj_bld_upload(void* d, const void* h, size_t bytes) {
    const size_t CH = (size_t)1 << 30;
    for (size_t o = 0; o < bytes; o += CH) CK(cudaMemcpy((char*)d + o, (const char*)h + o, std::min(CH, bytes - o), cudaMemcpyHostToDevice));
    return KJ_OK;
}

// Builds the large arrays of the context on the device.  c->H holds the meta data (kj_build_host_meta); `base` (scaled build only) is a
// finished context over the same index with copies = 1, whose device index resolves base suffix-array rows to taxa.
static int kj_device_build(kj_ctx* c, const kj_index_view& v, const uint8_t lcode[256], uint32_t rep, const kj_ctx* base, uint64_t& tot) {
    KjHostIndex& H = c->H; const int alen = H.alen; const uint64_t n = H.bwtlen, nb = H.nb; const int wide = H.wide;
    const uint32_t RB = kj_rank_rows(wide), RW = kj_rank_words(wide);
    KjBuildLcode lc; memcpy(lc.v, lcode, 256);
    const int grid_big = c->sm_count * 16;
    // ---- BWT bytes to the device (freed again below)
    uint8_t* d_bwt = nullptr; CK(cudaMalloc((void**)&d_bwt, (size_t)v.bwtlen));
    struct Free { void* p; ~Free() { if (p) cudaFree(p); } } free_bwt{d_bwt};
    int rc = kj_bld_upload(d_bwt, v.bwt, (size_t)v.bwtlen); if (rc) return rc;
    // ---- letter counts per tile, C[]
    const uint64_t ntiles = (nb + KJ_BLD_THREADS - 1) / KJ_BLD_THREADS;
    if (ntiles >= (1ull << 31)) { kj_err() = "index too large"; return KJ_ERR_UNSUPPORTED; }
    uint32_t* d_tc = nullptr; uint64_t* d_tp = nullptr; uint64_t* d_small = nullptr;
    CK(cudaMalloc((void**)&d_tc, ntiles * KJ_MAX_ALEN * 4)); Free f1{d_tc};
    CK(cudaMalloc((void**)&d_tp, ntiles * KJ_MAX_ALEN * 8)); Free f2{d_tp};
    CK(cudaMalloc((void**)&d_small, 3 * KJ_MAX_ALEN * 8 + 64)); Free f3{d_small};
    uint64_t* d_tot = d_small; uint64_t* d_C = d_small + KJ_MAX_ALEN + 1;
    if (wide) kj_bld_count<KJ_RANK_ROWS_WIDE><<<(unsigned)ntiles, KJ_BLD_THREADS>>>(d_bwt, lc, n, rep, alen, d_tc);
    else kj_bld_count<KJ_RANK_ROWS_NARROW><<<(unsigned)ntiles, KJ_BLD_THREADS>>>(d_bwt, lc, n, rep, alen, d_tc);
    kj_bld_scan_tiles<<<1, 32 * KJ_MAX_ALEN>>>(d_tc, ntiles, d_tp, d_tot);
    CK(cudaGetLastError());
    uint64_t tots[KJ_MAX_ALEN]; CK(cudaMemcpy(tots, d_tot, sizeof tots, cudaMemcpyDeviceToHost));
    H.C[0] = 0; for (int a = 0; a < alen; a++) H.C[a + 1] = H.C[a] + tots[a];
    if (H.C[alen] != n) { kj_err() = "letter counts do not add up"; return KJ_ERR_IO; }
    CK(cudaMemcpy(d_C, H.C, sizeof(uint64_t) * (size_t)(alen + 1), cudaMemcpyHostToDevice));
    // ---- rank records
    const size_t rank_bytes = (size_t)alen * nb * RW * 8;
    CK(cudaMalloc(&c->d_rank, rank_bytes)); tot += rank_bytes;
    {
        const size_t smem = (size_t)KJ_BLD_THREADS * (RB + 4) + 64;
        if (wide) {
            CK(cudaFuncSetAttribute(kj_bld_records<KJ_RANK_ROWS_WIDE, KJ_RANK_WORDS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kj_bld_records<KJ_RANK_ROWS_WIDE, KJ_RANK_WORDS_WIDE><<<(unsigned)ntiles, KJ_BLD_THREADS, smem>>>(d_bwt, lc, n, rep, alen, nb, d_tp, d_C, (uint64_t*)c->d_rank);
        } else kj_bld_records<KJ_RANK_ROWS_NARROW, KJ_RANK_WORDS_NARROW><<<(unsigned)ntiles, KJ_BLD_THREADS, smem>>>(d_bwt, lc, n, rep, alen, nb, d_tp, d_C, (uint64_t*)c->d_rank);
        CK(cudaGetLastError());
    }
    // ---- packed letters
    const uint64_t nwords = n / KJ_LETTERS_PER_WORD + 2;
    CK(cudaMalloc(&c->d_letters, nwords * 8)); tot += nwords * 8;
    kj_bld_letters<<<grid_big, 256>>>(d_bwt, lc, n, rep, nwords, (uint64_t*)c->d_letters);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    cudaFree(d_bwt); free_bwt.p = nullptr; cudaFree(d_tc); f1.p = nullptr; cudaFree(d_tp); f2.p = nullptr;
    c->launches += 4;
    // ---- sequence -> taxon, sampled suffix array -> taxon
    { size_t b = std::max<size_t>(H.seq_tax.size() * 4, 16); CK(cudaMalloc(&c->d_seq_tax, b)); tot += b; if (!H.seq_tax.empty()) CK(cudaMemcpy(c->d_seq_tax, H.seq_tax.data(), H.seq_tax.size() * 4, cudaMemcpyHostToDevice)); }
    if (!H.seq_acc.empty()) { CK(cudaMalloc(&c->d_seq_acc, H.seq_acc.size() * 4)); tot += H.seq_acc.size() * 4; CK(cudaMemcpy(c->d_seq_acc, H.seq_acc.data(), H.seq_acc.size() * 4, cudaMemcpyHostToDevice)); }
    uint64_t n_sa;
    if (rep == 1) {
        n_sa = (uint64_t)v.ncheck;
        CK(cudaMalloc(&c->d_sa_tax, (n_sa + 1) * 4)); tot += (n_sa + 1) * 4;
        CK(cudaMemset((uint32_t*)c->d_sa_tax + n_sa, 0xff, 4));       // guard entry (see create_ctx): the last sampled row has no entry in a reference-built index
        if (c->d_seq_acc) { CK(cudaMalloc(&c->d_sa_acc, (n_sa + 1) * 4)); tot += (n_sa + 1) * 4; CK(cudaMemset((uint32_t*)c->d_sa_acc + n_sa, 0xff, 4)); }
        const uint64_t CHE = (uint64_t)1 << 26;                        // entries per upload chunk
        uint8_t* d_sa = nullptr; CK(cudaMalloc((void**)&d_sa, (size_t)std::min<uint64_t>(CHE, std::max<uint64_t>(n_sa, 1)) * (size_t)v.nbytes)); Free f4{d_sa};
        for (uint64_t e0 = 0; e0 < n_sa; e0 += CHE) {
            const uint64_t m = std::min(CHE, n_sa - e0);
            CK(cudaMemcpy(d_sa, v.sa + e0 * (uint64_t)v.nbytes, (size_t)m * (size_t)v.nbytes, cudaMemcpyHostToDevice));
            kj_bld_sa_tax<<<grid_big, 256>>>(d_sa, m, e0, v.nbytes, v.pbits, (uint32_t)v.nseq, (const uint32_t*)c->d_seq_tax, (uint32_t*)c->d_sa_tax, c->d_err);
            if (c->d_sa_acc) kj_bld_sa_tax<<<grid_big, 256>>>(d_sa, m, e0, v.nbytes, v.pbits, (uint32_t)v.nseq, (const uint32_t*)c->d_seq_acc, (uint32_t*)c->d_sa_acc, c->d_err);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); c->launches++;
        }
        uint32_t e = 0; CK(cudaMemcpy(&e, c->d_err, 4, cudaMemcpyDeviceToHost));
        if (e & 64u) { CK(cudaMemset(c->d_err, 0, 4)); kj_err() = "corrupt suffix array (sequence number out of range)"; return KJ_ERR_IO; }
    } else {
        const int64_t last = (int64_t)((n - 1) >> H.sa_exp) - H.sa_bias;  // entry of the last sampled row
        n_sa = last >= 0 ? (uint64_t)last + 1 : 0;
        CK(cudaMalloc(&c->d_sa_tax, std::max<size_t>(n_sa * 4, 16))); tot += std::max<size_t>(n_sa * 4, 16);
        if (base->H.wide) kj_bld_sa_tax_scaled<uint64_t><<<grid_big, 256>>>(base->d_ix, n_sa, H.sa_bias, H.sa_exp, n, rep, (uint32_t*)c->d_sa_tax);
        else kj_bld_sa_tax_scaled<uint32_t><<<grid_big, 256>>>(base->d_ix, n_sa, H.sa_bias, H.sa_exp, n, rep, (uint32_t*)c->d_sa_tax);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); c->launches++;
    }
    c->n_sa = n_sa;
    return KJ_OK;
}

// quirk constants and the k-mer table need rank queries on the finished records: run after the device descriptor exists
// d_dst/k_out: where the table and its k go (default: the context's table for both modes); the second, larger table of the MEM kernels passes its own
static int kj_device_build_kmer(kj_ctx* c, int k, uint64_t& tot, void** d_dst = nullptr, int* k_out = nullptr) {
    KjHostIndex& H = c->H; const int wide = H.wide;
    if (!d_dst) { d_dst = &c->d_kmer; k_out = &H.kmer_k; }
    if (H.quirk_lo != ~0ull && d_dst == &c->d_kmer) {
        if (wide) kj_bld_quirk<uint64_t><<<1, 32>>>(c->d_ix, H.bwtlen - 65536ull, c->d_quirk); else kj_bld_quirk<uint32_t><<<1, 32>>>(c->d_ix, H.bwtlen - 65536ull, c->d_quirk);
        CK(cudaGetLastError()); CK(cudaMemcpy(H.quirk_d, c->d_quirk, sizeof H.quirk_d, cudaMemcpyDeviceToHost)); c->launches++;
    }
    *k_out = 0;
    if (k < 2 || k > 7 || H.alen != 21) return KJ_OK;
    uint64_t n_final = 1; for (int d = 0; d < k; d++) n_final *= 20;
    { size_t fr = 0, to = 0; CK(cudaMemGetInfo(&fr, &to));      // two level buffers + the final table next to the index: a smaller k when that does not fit
      while (k > 2 && (double)n_final * (2.0 * sizeof(KjKmer) + sizeof(KjKmer32)) > 0.8 * (double)fr) { k--; n_final /= 20; } }
    KjKmer* d_a = nullptr; KjKmer* d_b = nullptr;
    CK(cudaMalloc((void**)&d_a, n_final * sizeof(KjKmer))); struct Free { void* p; ~Free() { if (p) cudaFree(p); } } fa{d_a};
    CK(cudaMalloc((void**)&d_b, n_final * sizeof(KjKmer))); Free fb{d_b};
    std::vector<KjKmer> first(20); for (uint32_t a = 0; a < 20; a++) { first[a].lo = H.C[a + 1]; first[a].hi = H.C[a + 2]; }
    CK(cudaMemcpy(d_a, first.data(), 20 * sizeof(KjKmer), cudaMemcpyHostToDevice));
    uint64_t n_cur = 20;
    for (int d = 1; d < k; d++) {
        const unsigned g = (unsigned)std::min<uint64_t>((n_cur * 20 + 255) / 256, (uint64_t)c->sm_count * 32);
        if (wide) kj_bld_kmer_level<uint64_t><<<g, 256>>>(c->d_ix, d_a, n_cur, d_b); else kj_bld_kmer_level<uint32_t><<<g, 256>>>(c->d_ix, d_a, n_cur, d_b);
        CK(cudaGetLastError()); c->launches++;
        std::swap(d_a, d_b); fa.p = d_a; fb.p = d_b; n_cur *= 20;
    }
    if (wide) { *d_dst = d_a; fa.p = nullptr; tot += n_final * sizeof(KjKmer); }
    else {
        CK(cudaMalloc(d_dst, n_final * sizeof(KjKmer32))); tot += n_final * sizeof(KjKmer32);
        kj_bld_kmer_narrow<<<c->sm_count * 8, 256>>>(d_a, n_final, (KjKmer32*)*d_dst); CK(cudaGetLastError()); c->launches++;
    }
    CK(cudaDeviceSynchronize());
    *k_out = k;
    return KJ_OK;
}
