// kj_ingest.h -- FASTA/FASTQ text -> packed reads -> classification -> output text, all on the device.
//
// Replaces the reader loop of the reference's front end (src/kaiju.cpp:288-394: file-type detection by the first
// character, name = header line without its first character cut at the first of " /\t\r", FASTQ = 4-line records,
// FASTA = header + all lines up to the next '>' line, strip() of non-letters, util.cpp:25-33) and the output
// formatting of ConsumerThread::doWork (ConsumerThread.cpp:724-739) for whole chunks of text:
//   newline positions (count + scan + scatter) -> per-line classification (header / sequence / ignored), letters and
//   trimmed-name lengths (one warp per line) -> scans over lines -> packed sequences + offsets + names blob ->
//   kj_classify_kernel -> per-record output line lengths -> scan -> formatted "C\tname\ttaxid..." text.
// The host only moves bytes: file -> pinned buffer -> device, device -> pinned buffer -> file (kj_classify_files).
// Included by kj_device.cu (one translation unit: it uses kj_ctx and launch()).
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>
#include <condition_variable>
#include <deque>
#include <mutex>

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan of uint32 (in place); total -> *total.  Tile = 256 threads x 8 elements.
// ------------------------------------------------------------------------------------------------
#define KJ_SCAN_TILE 2048u
static __device__ __forceinline__ uint32_t kj_block_excl_scan(uint32_t v, uint32_t* smem_warp, uint32_t& block_total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += o; }
    if (lane == 31) smem_warp[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? smem_warp[lane] : 0u, y = w;
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, y, d); if (lane >= (uint32_t)d) y += o; }
        smem_warp[32 + lane] = y - w;                       // exclusive warp bases
        if (lane == 31) smem_warp[64] = y;                  // block total
    }
    __syncthreads();
    block_total = smem_warp[64];
    const uint32_t r = smem_warp[32 + wid] + x - v;
    __syncthreads();
    return r;
}
__global__ void kj_scan_reduce(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sw[80];
    const uint64_t base = (uint64_t)blockIdx.x * KJ_SCAN_TILE + (uint64_t)threadIdx.x * 8u; uint32_t s = 0;
    for (int k = 0; k < 8; k++) if (base + k < n) s += in[base + k];
    uint32_t tot; (void)kj_block_excl_scan(s, sw, tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void kj_scan_blocks(uint32_t* __restrict__ block_sums, uint32_t nblocks, uint32_t* __restrict__ total) {   // one block of 1024 threads
    __shared__ uint32_t sw[80];
    uint32_t carry = 0;
    for (uint32_t b = 0; b < nblocks; b += blockDim.x) {
        const uint32_t i = b + threadIdx.x; const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t tot; const uint32_t e = kj_block_excl_scan(v, sw, tot);
        if (i < nblocks) block_sums[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void kj_scan_apply(uint32_t* __restrict__ data, uint64_t n, const uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sw[80];
    const uint64_t base = (uint64_t)blockIdx.x * KJ_SCAN_TILE + (uint64_t)threadIdx.x * 8u;
    uint32_t v[8], s = 0;
    for (int k = 0; k < 8; k++) { v[k] = base + k < n ? data[base + k] : 0u; s += v[k]; }
    uint32_t tot; uint32_t e = kj_block_excl_scan(s, sw, tot) + block_sums[blockIdx.x];
    for (int k = 0; k < 8; k++) { if (base + k < n) data[base + k] = e; e += v[k]; }
}

// ------------------------------------------------------------------------------------------------
// text -> lines
// ------------------------------------------------------------------------------------------------
#define KJ_NL_TILE 4096u            // bytes per block: 256 threads x 16 bytes
static __device__ __forceinline__ uint32_t kj_nl_mask16(const char* __restrict__ text, uint64_t n, uint64_t p) {   // bit k set <=> text[p+k] == '\n'
    uint32_t m = 0;
    if (p + 16 <= n && ((uintptr_t)(text + p) & 15u) == 0) {
        const uint4 q = *(const uint4*)(text + p); const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        #pragma unroll
        for (int k = 0; k < 16; k++) if (((w[k >> 2] >> (8 * (k & 3))) & 0xffu) == '\n') m |= 1u << k;
    } else for (int k = 0; k < 16; k++) if (p + k < n && text[p + k] == '\n') m |= 1u << k;
    return m;
}
__global__ void kj_nl_count(const char* __restrict__ text, uint64_t n, uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t sw[80];
    const uint64_t p = (uint64_t)blockIdx.x * KJ_NL_TILE + (uint64_t)threadIdx.x * 16u;
    const uint32_t c = p < n ? (uint32_t)__popc(kj_nl_mask16(text, n, p)) : 0u;
    uint32_t tot; (void)kj_block_excl_scan(c, sw, tot);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}
// line_start[0] = 0, line_start[k] = position after the k-th newline
__global__ void kj_nl_scatter(const char* __restrict__ text, uint64_t n, const uint32_t* __restrict__ tile_base, uint64_t* __restrict__ line_start) {
    __shared__ uint32_t sw[80];
    const uint64_t p = (uint64_t)blockIdx.x * KJ_NL_TILE + (uint64_t)threadIdx.x * 16u;
    uint32_t m = p < n ? kj_nl_mask16(text, n, p) : 0u;
    uint32_t tot; uint32_t e = kj_block_excl_scan((uint32_t)__popc(m), sw, tot) + tile_base[blockIdx.x];
    while (m) { const int k = __ffs((int)m) - 1; m &= m - 1; line_start[++e] = p + (uint64_t)k + 1u; }
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
}

struct KjParseDims { uint32_t fastq; uint32_t n_lines; };
static __device__ __forceinline__ bool kj_is_letter(uint32_t c) { const uint32_t u = c & 0xDFu; return u >= 'A' && u <= 'Z'; }    // util.cpp:21-23
// one warp per complete line: cnt = letters of a sequence line, hdr = 1 for a header line, nlen = trimmed name length
__global__ void kj_line_info(const char* __restrict__ text, const uint64_t* __restrict__ line_start, KjParseDims d,
                             uint32_t* __restrict__ cnt, uint32_t* __restrict__ hdr, uint32_t* __restrict__ nlen, uint32_t* __restrict__ err) {
    const uint32_t lane = threadIdx.x & 31; const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i <= d.n_lines; i += nwarps) {
        if (i == d.n_lines) { if (lane == 0) { cnt[i] = 0; hdr[i] = 0; nlen[i] = 0; } continue; }     // virtual line: the scans deliver the totals here
        const uint64_t s = line_start[i], e = line_start[i + 1] - 1;                                  // [s, e) without the newline
        bool is_hdr, is_seq;
        if (d.fastq) { is_hdr = (i & 3u) == 0; is_seq = (i & 3u) == 1; if (is_hdr && lane == 0 && e > s && text[s] != '@') atomicOr(err, 16u); }
        else { is_hdr = e > s && text[s] == '>'; is_seq = !is_hdr; }
        uint32_t c = 0, nl = 0;
        if (is_seq) {
            for (uint64_t p = s + lane; p < e; p += 32) c += kj_is_letter((uint8_t)text[p]) ? 1u : 0u;
            for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
        } else if (is_hdr) {
            // name = line without its first character, cut at the first of " /\t\r" (kaiju.cpp:279, 303-307)
            nl = e > s ? (uint32_t)(e - (s + 1)) : 0u;
            for (uint64_t p0 = s + 1; p0 < e; p0 += 32) {
                const uint64_t p = p0 + lane; const uint32_t ch = p < e ? (uint8_t)text[p] : 0u;
                const uint32_t stop = __ballot_sync(0xffffffffu, p < e && (ch == ' ' || ch == '/' || ch == '\t' || ch == '\r'));
                if (stop) { nl = (uint32_t)(p0 - (s + 1)) + (uint32_t)(__ffs((int)stop) - 1); break; }
            }
        }
        if (lane == 0) { cnt[i] = c; hdr[i] = is_hdr ? 1u : 0u; nlen[i] = nl; }
    }
}
// after the scans: S = letters before line i, R = headers before line i, NS = name bytes before line i
__global__ void kj_line_emit(const char* __restrict__ text, const uint64_t* __restrict__ line_start, KjParseDims d,
                             const uint32_t* __restrict__ S, const uint32_t* __restrict__ R, const uint32_t* __restrict__ NS,
                             char* __restrict__ seq, uint64_t* __restrict__ off, char* __restrict__ names, uint32_t* __restrict__ name_off, uint64_t* __restrict__ rec_pos, uint64_t nbytes) {
    const uint32_t lane = threadIdx.x & 31; const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i <= d.n_lines; i += nwarps) {
        if (i == d.n_lines) { if (lane == 0) { const uint32_t r = R[i]; off[r] = S[i]; name_off[r] = NS[i]; rec_pos[r] = line_start[i]; } continue; }
        const uint64_t s = line_start[i], e = line_start[i + 1] - 1;
        const bool is_hdr = d.fastq ? (i & 3u) == 0 : (e > s && text[s] == '>');
        const bool is_seq = d.fastq ? (i & 3u) == 1 : !is_hdr;
        if (is_hdr) {
            const uint32_t r = R[i], nl = NS[i + 1] - NS[i];
            if (lane == 0) { off[r] = S[i]; name_off[r] = NS[i]; rec_pos[r] = s; }
            for (uint32_t k = lane; k < nl; k += 32) names[NS[i] + k] = text[s + 1 + k];
        } else if (is_seq) {
            uint64_t dst = S[i];
            for (uint64_t p0 = s; p0 < e; p0 += 32) {
                const uint64_t p = p0 + lane; const char ch = p < e ? text[p] : 0; const bool l = p < e && kj_is_letter((uint8_t)ch);
                const uint32_t m = __ballot_sync(0xffffffffu, l);
                if (l) seq[dst + (uint32_t)__popc(m & ((1u << lane) - 1u))] = ch;
                dst += (uint32_t)__popc(m);
            }
        }
    }
    (void)nbytes;
}
// paired input: names of both files must be identical record by record (kaiju.cpp:359-362, 377-380)
__global__ void kj_names_equal(const char* __restrict__ na, const uint32_t* __restrict__ oa, const char* __restrict__ nb, const uint32_t* __restrict__ ob, uint64_t n, uint32_t* __restrict__ err) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t la = oa[r + 1] - oa[r], lb = ob[r + 1] - ob[r]; bool same = la == lb;
        for (uint32_t k = 0; same && k < la; k++) same = na[oa[r] + k] == nb[ob[r] + k];
        if (!same) atomicOr(err, 32u);
    }
}

// ------------------------------------------------------------------------------------------------
// output lines (ConsumerThread.cpp:724-739): "C\t<name>\t<taxid>[\t<best>\t<id,id,...,>]\n" / "U\t<name>\t0\n"
// ------------------------------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t kj_dec_len(uint64_t v) { uint32_t l = 1; while (v >= 10) { v /= 10; l++; } return l; }
static __device__ __forceinline__ char* kj_dec_write(char* p, uint64_t v) { const uint32_t l = kj_dec_len(v); for (uint32_t k = l; k-- > 0;) { p[k] = (char)('0' + (uint32_t)(v % 10)); v /= 10; } return p + l; }
__global__ void kj_fmt_len(const uint64_t* __restrict__ tax, const uint32_t* __restrict__ best, const uint64_t* __restrict__ ids, const uint8_t* __restrict__ nids,
                           const uint32_t* __restrict__ name_off, uint64_t n, int verbose, uint32_t* __restrict__ len, unsigned long long* __restrict__ n_classified) {
    uint32_t ccount = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (uint64_t)gridDim.x * blockDim.x) {
        if (r == n) { len[r] = 0; continue; }
        const uint32_t nl = name_off[r + 1] - name_off[r]; const uint64_t t = tax[r]; uint32_t l;
        if (!t) l = 2 + nl + 3;
        else {
            ccount++;
            l = 2 + nl + 1 + kj_dec_len(t) + 1;
            if (verbose) { l += 1 + kj_dec_len(best[r]) + 1; for (uint32_t k = 0; k < nids[r]; k++) l += kj_dec_len(ids[r * KJ_MAX_IDS + k]) + 1; }
        }
        len[r] = l;
    }
    for (int m = 16; m > 0; m >>= 1) ccount += __shfl_xor_sync(0xffffffffu, ccount, m);
    if ((threadIdx.x & 31) == 0 && ccount) atomicAdd(n_classified, (unsigned long long)ccount);
}
__global__ void kj_fmt_write(const uint64_t* __restrict__ tax, const uint32_t* __restrict__ best, const uint64_t* __restrict__ ids, const uint8_t* __restrict__ nids,
                             const char* __restrict__ names, const uint32_t* __restrict__ name_off, uint64_t n, int verbose, const uint32_t* __restrict__ pos, char* __restrict__ out) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        char* p = out + pos[r]; const uint32_t nl = name_off[r + 1] - name_off[r]; const uint64_t t = tax[r];
        *p++ = t ? 'C' : 'U'; *p++ = '\t';
        for (uint32_t k = 0; k < nl; k++) *p++ = names[name_off[r] + k];
        *p++ = '\t';
        if (!t) { *p++ = '0'; *p++ = '\n'; continue; }
        p = kj_dec_write(p, t);
        if (verbose) { *p++ = '\t'; p = kj_dec_write(p, best[r]); *p++ = '\t'; for (uint32_t k = 0; k < nids[r]; k++) { p = kj_dec_write(p, ids[r * KJ_MAX_IDS + k]); *p++ = ','; } }
        *p++ = '\n';
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct KjDevBuf {   // grow-only device buffer
    void* p = nullptr; size_t cap = 0;
    int need(size_t bytes) { if (bytes <= cap) return KJ_OK; if (p) cudaFree(p); p = nullptr; cap = 0; size_t c = bytes + bytes / 4 + 4096; CK(cudaMalloc(&p, c)); cap = c; return KJ_OK; }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
static int kj_scan_u32(uint32_t* d, uint64_t n, KjDevBuf& tmp, uint32_t* d_total, cudaStream_t st) {
    const uint32_t nb = (uint32_t)((n + KJ_SCAN_TILE - 1) / KJ_SCAN_TILE);
    int rc = tmp.need((size_t)(nb + 1) * sizeof(uint32_t)); if (rc) return rc;
    kj_scan_reduce<<<nb, 256, 0, st>>>(d, n, tmp.as<uint32_t>());
    kj_scan_blocks<<<1, 1024, 0, st>>>(tmp.as<uint32_t>(), nb, d_total);
    kj_scan_apply<<<nb, 256, 0, st>>>(d, n, tmp.as<uint32_t>());
    CK(cudaGetLastError());
    return KJ_OK;
}

struct KjParsed {   // one side (file) of a chunk after parsing
    KjDevBuf text[2]; int cur = 0; uint64_t nbytes = 0;         // device text (ping-pong for the carry), valid bytes
    KjDevBuf line_start, cnt, hdr, nlen, tiles, scan_tmp, seq, off, names, name_off, rec_pos, totals;
    int fastq = -1;                                              // file type, fixed by the first byte of the file
    uint64_t n_lines = 0, n_rec = 0, consumed = 0; bool eof = false;
    void release() { for (KjDevBuf* b : {&text[0], &text[1], &line_start, &cnt, &hdr, &nlen, &tiles, &scan_tmp, &seq, &off, &names, &name_off, &rec_pos, &totals}) b->release(); }
};

// Parse the complete records in P.text[P.cur][0, P.nbytes).  Sets P.n_rec and P.consumed (bytes covered by those records).
static int kj_parse_side(kj_ctx* c, KjParsed& P, const std::string& fname, cudaStream_t st) {
    P.n_rec = 0; P.consumed = 0; P.n_lines = 0;
    if (P.nbytes == 0) return KJ_OK;
    if (P.nbytes >= (1ull << 31)) { kj_err() = "kj_classify_files: a single record larger than 2 GB"; return KJ_ERR_UNSUPPORTED; }
    const char* text = P.text[P.cur].as<char>();
    if (P.fastq < 0) {
        char first = 0; CK(cudaMemcpyAsync(&first, text, 1, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        if (first == '@') P.fastq = 1; else if (first == '>') P.fastq = 0;
        else { kj_err() = "Auto-detection of file type for file " + fname + " failed."; return KJ_ERR_IO; }                 // kaiju.cpp:296-299
    }
    int rc;
    const uint32_t ntiles = (uint32_t)((P.nbytes + KJ_NL_TILE - 1) / KJ_NL_TILE);
    if ((rc = P.tiles.need((size_t)(ntiles + 1) * 4)) || (rc = P.totals.need(64))) return rc;
    uint32_t* tot = P.totals.as<uint32_t>();
    kj_nl_count<<<ntiles, 256, 0, st>>>(text, P.nbytes, P.tiles.as<uint32_t>());
    if ((rc = kj_scan_u32(P.tiles.as<uint32_t>(), ntiles, P.scan_tmp, tot + 0, st))) return rc;
    uint32_t nl = 0; CK(cudaMemcpyAsync(&nl, tot + 0, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    P.n_lines = nl; c->launches += 4;
    if (nl == 0) return KJ_OK;                                                      // not even one complete line yet
    const size_t L1 = (size_t)nl + 1;
    if ((rc = P.line_start.need((L1 + 1) * 8)) || (rc = P.cnt.need((L1 + 1) * 4)) || (rc = P.hdr.need((L1 + 1) * 4)) || (rc = P.nlen.need((L1 + 1) * 4))) return rc;
    kj_nl_scatter<<<ntiles, 256, 0, st>>>(text, P.nbytes, P.tiles.as<uint32_t>(), P.line_start.as<uint64_t>());
    KjParseDims d; d.fastq = (uint32_t)P.fastq; d.n_lines = nl;
    const int blocks = c->sm_count * 8;
    kj_line_info<<<blocks, 256, 0, st>>>(text, P.line_start.as<uint64_t>(), d, P.cnt.as<uint32_t>(), P.hdr.as<uint32_t>(), P.nlen.as<uint32_t>(), c->d_err);
    if ((rc = kj_scan_u32(P.cnt.as<uint32_t>(), L1, P.scan_tmp, tot + 1, st)) || (rc = kj_scan_u32(P.hdr.as<uint32_t>(), L1, P.scan_tmp, tot + 2, st)) ||
        (rc = kj_scan_u32(P.nlen.as<uint32_t>(), L1, P.scan_tmp, tot + 3, st))) return rc;
    uint32_t h[4]; CK(cudaMemcpyAsync(h, tot, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    const uint64_t letters = h[1], headers = h[2], name_bytes = h[3];
    if ((rc = P.seq.need(letters + 64)) || (rc = P.off.need((headers + 2) * 8)) || (rc = P.names.need(name_bytes + 64)) || (rc = P.name_off.need((headers + 2) * 4)) ||
        (rc = P.rec_pos.need((headers + 2) * 8))) return rc;
    kj_line_emit<<<blocks, 256, 0, st>>>(text, P.line_start.as<uint64_t>(), d, P.cnt.as<uint32_t>(), P.hdr.as<uint32_t>(), P.nlen.as<uint32_t>(),
                                        P.seq.as<char>(), P.off.as<uint64_t>(), P.names.as<char>(), P.name_off.as<uint32_t>(), P.rec_pos.as<uint64_t>(), P.nbytes);
    CK(cudaGetLastError()); c->launches += 12;
    // complete records: FASTQ = whole groups of four lines; FASTA = every header that is followed by another header (or by the end of the file)
    if (P.fastq) { P.n_rec = nl / 4; }
    else { P.n_rec = P.eof ? headers : (headers ? headers - 1 : 0); }
    return KJ_OK;
}
// bytes of the text covered by the first n records (n <= n_rec)
static int kj_parse_consumed(KjParsed& P, uint64_t n, uint64_t headers_total_hint, cudaStream_t st, uint64_t& consumed) {
    (void)headers_total_hint;
    uint64_t v = 0; CK(cudaMemcpyAsync(&v, P.rec_pos.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    consumed = v; return KJ_OK;
}

struct KjChunk { char* p = nullptr; size_t n = 0; bool eof = false; std::string error; };
struct KjFileReader {   // per input file: gz (zlib, one thread) or plain (a pool of pread() threads) -> pinned chunks
    gzFile fp = nullptr; int fd = -1; uint64_t file_off = 0; std::string path; size_t chunk; std::vector<char*> pool; std::deque<KjChunk> ready; std::deque<char*> free_;
    std::mutex mu; std::condition_variable cv; std::thread th; bool stop = false;
    // plain files: NT worker threads copy 1 MB slices of the current chunk out of the page cache in parallel (one copy stream per thread;
    // a single thread moves ~1-3 GB/s, the device pipeline wants > 20 GB/s)
    static constexpr size_t SLICE = 1u << 20;
    std::vector<std::thread> workers; std::mutex wmu; std::condition_variable wcv, wdone; char* wbuf = nullptr; size_t wnext = 0, wslices = 0, wleft = 0; uint64_t wgen = 0; bool wstop = false, werr = false;
    std::vector<size_t> wgot;
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(wmu);
            wcv.wait(lk, [&] { return wstop || (wgen != seen && wnext < wslices); });
            if (wstop) return;
            const uint64_t gen = wgen;
            while (wgen == gen && wnext < wslices) {
                const size_t sl = wnext++; char* b = wbuf; const uint64_t base = file_off;
                lk.unlock();
                const size_t o = sl * SLICE, want = std::min(SLICE, chunk - o); size_t g = 0; bool bad = false;
                while (g < want) { const ssize_t r = ::pread(fd, b + o + g, want - g, (off_t)(base + o + g)); if (r < 0) { bad = true; break; } if (r == 0) break; g += (size_t)r; }
                lk.lock();
                wgot[sl] = g; if (bad) werr = true;
                if (--wleft == 0) wdone.notify_all();
            }
            seen = gen;
        }
    }
    int open(const std::string& p, size_t chunk_bytes, int nbuf) {
        path = p; chunk = chunk_bytes;
        fd = ::open(p.c_str(), O_RDONLY);
        if (fd < 0) { kj_err() = "Could not open file " + p; return KJ_ERR_IO; }
        unsigned char magic[2] = {0, 0}; const ssize_t got = ::pread(fd, magic, 2, 0);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {             // gzip: inflate through zlib; everything else is read as it is
            ::close(fd); fd = -1; fp = gzopen(p.c_str(), "rb");
            if (!fp) { kj_err() = "Could not open file " + p; return KJ_ERR_IO; }
            gzbuffer(fp, 1 << 20);
        } else {
            posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
            unsigned nt = std::max(2u, std::min(16u, std::thread::hardware_concurrency() / 4u));
            if (const char* v = getenv("KJ_IO_THREADS")) { const long x = atol(v); if (x >= 1 && x <= 64) nt = (unsigned)x; }      // tuning hook
            wgot.assign((chunk + SLICE - 1) / SLICE, 0);
            for (unsigned t = 0; t < nt; t++) workers.emplace_back([this] { worker(); });
        }
        for (int i = 0; i < nbuf; i++) { char* b = nullptr; if (cudaMallocHost((void**)&b, chunk) != cudaSuccess) { kj_err() = "cudaMallocHost failed"; return KJ_ERR_NOMEM; } pool.push_back(b); free_.push_back(b); }
        th = std::thread([this] { run(); });
        return KJ_OK;
    }
    void run() {
        for (;;) {
            char* b;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || !free_.empty(); }); if (stop) return; b = free_.front(); free_.pop_front(); }
            KjChunk ck; ck.p = b; size_t got = 0;
            if (fd >= 0) {
                const size_t ns = (chunk + SLICE - 1) / SLICE;
                {
                    std::unique_lock<std::mutex> lk(wmu);
                    wbuf = b; wnext = 0; wslices = ns; wleft = ns; werr = false; wgen++;
                    wcv.notify_all();
                    wdone.wait(lk, [&] { return wleft == 0; });
                    wslices = 0;
                }
                if (werr) ck.error = "read error in file " + path;
                else for (size_t sl = 0; sl < ns; sl++) {
                    got += wgot[sl];
                    if (wgot[sl] < std::min(SLICE, chunk - sl * SLICE)) { ck.eof = true; break; }          // short slice = end of file (later slices read nothing)
                }
                if (got == chunk && !ck.eof && ck.error.empty()) { char probe; if (::pread(fd, &probe, 1, (off_t)(file_off + chunk)) == 0) ck.eof = true; }
                file_off += got;
            } else while (got < chunk) {
                const ssize_t r = (ssize_t)gzread(fp, b + got, (unsigned)std::min<size_t>(chunk - got, 1u << 30));
                if (r < 0) { ck.error = "read error in file " + path; break; }
                if (r == 0) { ck.eof = true; break; }
                got += (size_t)r;
            }
            ck.n = got;
            const bool last = ck.eof || !ck.error.empty();
            { std::lock_guard<std::mutex> lk(mu); ready.push_back(ck); }
            cv.notify_all();
            if (last) return;
        }
    }
    KjChunk next() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !ready.empty(); }); KjChunk ck = ready.front(); ready.pop_front(); return ck; }
    void give_back(char* b) { { std::lock_guard<std::mutex> lk(mu); free_.push_back(b); } cv.notify_all(); }
    void close() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_all();
        if (th.joinable()) th.join();
        { std::lock_guard<std::mutex> lk(wmu); wstop = true; } wcv.notify_all();
        for (auto& w : workers) if (w.joinable()) w.join();
        workers.clear();
        if (fp) gzclose(fp); fp = nullptr;
        if (fd >= 0) ::close(fd); fd = -1;
        for (char* b : pool) cudaFreeHost(b); pool.clear();
    }
};
struct KjWriter {   // ordered output: pinned buffers filled by D2H copies, written by one thread
    FILE* out = nullptr; bool own = false; std::vector<char*> pool; size_t cap = 0; std::deque<std::pair<char*, size_t>> ready; std::deque<char*> free_;
    std::mutex mu; std::condition_variable cv; std::thread th; bool done = false, failed = false;
    int open(const char* path, size_t cap_bytes, int nbuf) {
        if (path && *path) { out = fopen(path, "w"); own = true; if (!out) { kj_err() = std::string("Could not open file ") + path + " for writing"; return KJ_ERR_IO; } } else out = stdout;
        cap = cap_bytes;
        for (int i = 0; i < nbuf; i++) { char* b = nullptr; if (cudaMallocHost((void**)&b, cap) != cudaSuccess) { kj_err() = "cudaMallocHost failed"; return KJ_ERR_NOMEM; } pool.push_back(b); free_.push_back(b); }
        th = std::thread([this] {
            for (;;) {
                std::pair<char*, size_t> job;
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done || !ready.empty(); }); if (ready.empty()) return; job = ready.front(); ready.pop_front(); }
                if (job.second && fwrite(job.first, 1, job.second, out) != job.second) failed = true;
                { std::lock_guard<std::mutex> lk(mu); free_.push_back(job.first); } cv.notify_all();
            }
        });
        return KJ_OK;
    }
    char* get() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !free_.empty(); }); char* b = free_.front(); free_.pop_front(); return b; }
    void put(char* b, size_t n) { { std::lock_guard<std::mutex> lk(mu); ready.push_back({b, n}); } cv.notify_all(); }
    int close() {
        { std::lock_guard<std::mutex> lk(mu); done = true; } cv.notify_all();
        if (th.joinable()) th.join();
        if (out) { fflush(out); if (own) fclose(out); } out = nullptr;
        for (char* b : pool) cudaFreeHost(b); pool.clear();
        if (failed) { kj_err() = "write error on the output file"; return KJ_ERR_IO; }
        return KJ_OK;
    }
};

struct KjPrefetch { KjDevBuf stage[2]; int cur = 0; cudaEvent_t done = nullptr; bool pending = false, eof = false; char* host = nullptr; size_t n = 0; char last = 0; };
struct KjFilesState {
    KjParsed side[2]; KjFileReader rd[2]; KjWriter wr; KjPrefetch pf[2]; int nfiles = 1; bool rd_open[2] = {false, false}, wr_open = false;
    KjDevBuf tax, best, ids, nids, len, out, scan_tmp, totals;
    void cleanup() {
        for (int f = 0; f < 2; f++) {
            if (pf[f].pending && pf[f].done) cudaEventSynchronize(pf[f].done);
            if (rd_open[f]) rd[f].close(); side[f].release(); pf[f].stage[0].release(); pf[f].stage[1].release(); if (pf[f].done) cudaEventDestroy(pf[f].done);
        }
        if (wr_open) wr.close();
        for (KjDevBuf* b : {&tax, &best, &ids, &nids, &len, &out, &scan_tmp, &totals}) b->release();
    }
};

// next chunk of file f: pinned buffer -> device staging buffer, asynchronously on the context's second stream
static int kj_prefetch(kj_ctx* c, KjFilesState& S, int f) {
    KjPrefetch& F = S.pf[f];
    KjChunk ck = S.rd[f].next();
    if (!ck.error.empty()) { kj_err() = ck.error; return KJ_ERR_IO; }
    F.cur ^= 1; int rc = F.stage[F.cur].need(ck.n + 16); if (rc) return rc;
    if (!F.done) CK(cudaEventCreateWithFlags(&F.done, cudaEventDisableTiming));
    if (ck.n) CK(cudaMemcpyAsync(F.stage[F.cur].p, ck.p, ck.n, cudaMemcpyHostToDevice, c->stream[1]));
    CK(cudaEventRecord(F.done, c->stream[1]));
    F.pending = true; F.eof = ck.eof; F.host = ck.p; F.n = ck.n; F.last = ck.n ? ck.p[ck.n - 1] : 0;
    return KJ_OK;
}

static int kj_classify_files_impl(kj_ctx* c, KjFilesState& S, const char* in1, const char* in2, const char* out_path, int verbose, uint64_t* n_reads_out, uint64_t* n_class_out) {
    // developer hook KJ_FILES_TRACE: where the wall time of the file pipeline goes (ms per stage, summed over the chunks)
    const bool ftrace = getenv("KJ_FILES_TRACE") != nullptr; double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}; uint64_t nchunks = 0;
    auto now = [] { return std::chrono::steady_clock::now(); }; auto t_last = now();
    auto lap = [&](int k) { if (!ftrace) return; cudaStreamSynchronize(c->stream[0]); const auto t = now(); tm[k] += std::chrono::duration<double, std::milli>(t - t_last).count(); t_last = t; };
    size_t chunk = 64u << 20;
    if (const char* v = getenv("KJ_INGEST_CHUNK")) { long x = atol(v); if (x >= 256 && x <= (1l << 30)) chunk = (size_t)x; }       // test hook: many small chunks
    const bool paired = in2 && *in2; S.nfiles = paired ? 2 : 1;
    const std::string fn[2] = {in1, paired ? in2 : ""};
    cudaStream_t st = c->stream[0];
    int rc;
    for (int f = 0; f < S.nfiles; f++) { if ((rc = S.rd[f].open(fn[f], chunk, 3))) return rc; S.rd_open[f] = true; }
    size_t out_cap = 16u << 20;
    if ((rc = S.wr.open(out_path, out_cap, 3))) return rc; S.wr_open = true;
    if ((rc = S.totals.need(64))) return rc;
    uint64_t n_reads = 0; unsigned long long n_class = 0;
    CK(cudaMemsetAsync(S.totals.p, 0, 64, st));
    CK(cudaMemsetAsync(c->d_counts_pending, 0, (size_t)c->n_counts * 8, st));
    unsigned long long* d_nclass = (unsigned long long*)((char*)S.totals.p + 16);
    for (int f = 0; f < S.nfiles; f++) if ((rc = kj_prefetch(c, S, f))) return rc;
    for (;;) {
        nchunks++; lap(7);
        // 1. top up both sides: carry (already at the front of the device text) + the chunk that was prefetched into the staging buffer
        for (int f = 0; f < S.nfiles; f++) {
            KjParsed& P = S.side[f]; KjPrefetch& F = S.pf[f];
            if (!F.pending) continue;
            CK(cudaEventSynchronize(F.done));                          // H2D finished: the pinned chunk goes back to the reader
            S.rd[f].give_back(F.host); F.pending = false;
            if ((P.nbytes + F.n + 1) > P.text[P.cur].cap) {            // grow: move the carry into the larger buffer
                KjDevBuf nb; if ((rc = nb.need(P.nbytes + F.n + chunk + 1))) return rc;
                if (P.nbytes) CK(cudaMemcpyAsync(nb.p, P.text[P.cur].p, P.nbytes, cudaMemcpyDeviceToDevice, st));
                CK(cudaStreamSynchronize(st)); P.text[P.cur].release(); P.text[P.cur] = nb;
            }
            if (F.n) CK(cudaMemcpyAsync(P.text[P.cur].as<char>() + P.nbytes, F.stage[F.cur].p, F.n, cudaMemcpyDeviceToDevice, st));
            bool need_nl = false;
            if (F.eof) {
                P.eof = true; const uint64_t tot = P.nbytes + F.n;
                if (tot) { char last = F.last; if (!F.n) { CK(cudaMemcpyAsync(&last, P.text[P.cur].as<char>() + tot - 1, 1, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); } need_nl = last != '\n'; }
            }
            P.nbytes += F.n;
            if (need_nl) { CK(cudaMemsetAsync(P.text[P.cur].as<char>() + P.nbytes, '\n', 1, st)); P.nbytes += 1; }      // the last line of a file may lack its newline
        }
        lap(0);
        // ... and start the host-to-device copy of the following chunks on the second stream: it overlaps with the kernels below
        for (int f = 0; f < S.nfiles; f++) if (!S.side[f].eof && (rc = kj_prefetch(c, S, f))) return rc;
        lap(1);
        // 2. parse
        for (int f = 0; f < S.nfiles; f++) if ((rc = kj_parse_side(c, S.side[f], fn[f], st))) return rc;
        uint64_t n = S.side[0].n_rec; if (paired) n = std::min(n, S.side[1].n_rec);
        const bool all_eof = S.side[0].eof && (!paired || S.side[1].eof);
        if (paired && all_eof && S.side[0].n_rec > S.side[1].n_rec) { kj_err() = "File " + fn[0] + " contains more reads then file " + fn[1]; return KJ_ERR_IO; }   // kaiju.cpp:337-340
        lap(2);
        // 3. classify + format the first n records
        if (n) {
            if (n >= (1ull << 31)) { kj_err() = "kj_classify_files: chunk with too many records"; return KJ_ERR_UNSUPPORTED; }
            if (paired) kj_names_equal<<<c->sm_count * 4, 256, 0, st>>>(S.side[0].names.as<char>(), S.side[0].name_off.as<uint32_t>(), S.side[1].names.as<char>(), S.side[1].name_off.as<uint32_t>(), n, c->d_err);
            if ((rc = S.tax.need(n * 8)) || (rc = S.best.need(n * 4)) || (rc = S.len.need((n + 2) * 4))) return rc;
            if (verbose && ((rc = S.ids.need(n * KJ_MAX_IDS * 8)) || (rc = S.nids.need(n)))) return rc;
            for (;;) {     // repeated only when the Greedy variant ring had to grow
                CK(cudaMemsetAsync(c->d_maxlen, 0, 2 * sizeof(unsigned int), st));
                kj_maxlen_kernel<<<256, 256, 0, st>>>(S.side[0].off.as<uint64_t>(), n, c->d_maxlen);
                if (paired) kj_maxlen_kernel<<<256, 256, 0, st>>>(S.side[1].off.as<uint64_t>(), n, c->d_maxlen + 1);
                unsigned int h[2] = {0, 0}; CK(cudaMemcpyAsync(h, c->d_maxlen, sizeof h, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
                rc = launch(c, 0, S.side[0].seq.as<uint8_t>(), S.side[0].off.as<uint64_t>(), paired ? S.side[1].seq.as<uint8_t>() : nullptr, paired ? S.side[1].off.as<uint64_t>() : nullptr, 0, 0, n, h[0], h[1],
                            S.tax.as<uint64_t>(), S.best.as<uint32_t>(), st, false, verbose ? S.ids.as<uint64_t>() : nullptr, verbose ? S.nids.as<uint8_t>() : nullptr, c->d_counts_pending);
                if (rc) return rc;
                CK(cudaStreamSynchronize(st));
                uint32_t e = 0; CK(cudaMemcpy(&e, c->d_err, sizeof e, cudaMemcpyDeviceToHost));
                if (e & 16u) { CK(cudaMemset(c->d_err, 0, 4)); kj_err() = "malformed FASTQ record (a header line does not start with '@') in the input"; return KJ_ERR_IO; }
                if (e & 32u) { CK(cudaMemset(c->d_err, 0, 4)); kj_err() = "Read names are not identical between the two input files. Probably reads are not in the same order in both files."; return KJ_ERR_IO; }
                const uint32_t boost = c->variant_boost;
                rc = check_err_flag(c);
                if (rc) CK(cudaMemsetAsync(c->d_counts_pending, 0, (size_t)c->n_counts * 8, st));     // a failed launch does not count
                if (rc == KJ_ERR_OVERFLOW && c->variant_boost != boost) continue;
                if (rc) return rc;
                break;
            }
            lap(3);
            kj_count_commit<<<c->sm_count, 256, 0, st>>>(c->d_counts, c->d_counts_pending, c->n_counts); c->launches++;     // this chunk succeeded: its reads join the per-taxon counts
            kj_fmt_len<<<c->sm_count * 4, 256, 0, st>>>(S.tax.as<uint64_t>(), S.best.as<uint32_t>(), S.ids.as<uint64_t>(), S.nids.as<uint8_t>(), S.side[0].name_off.as<uint32_t>(), n, verbose, S.len.as<uint32_t>(), d_nclass);
            if ((rc = kj_scan_u32(S.len.as<uint32_t>(), n + 1, S.scan_tmp, (uint32_t*)S.totals.p, st))) return rc;
            uint32_t out_bytes = 0; CK(cudaMemcpyAsync(&out_bytes, S.totals.p, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
            if ((rc = S.out.need(out_bytes + 64))) return rc;
            kj_fmt_write<<<c->sm_count * 4, 256, 0, st>>>(S.tax.as<uint64_t>(), S.best.as<uint32_t>(), S.ids.as<uint64_t>(), S.nids.as<uint8_t>(), S.side[0].names.as<char>(), S.side[0].name_off.as<uint32_t>(), n, verbose,
                                                       S.len.as<uint32_t>(), S.out.as<char>());
            CK(cudaGetLastError()); c->launches += 6;
            lap(4);
            for (size_t o = 0; o < out_bytes; o += out_cap) {
                const size_t m = std::min<size_t>(out_cap, out_bytes - o); char* hb = S.wr.get();
                CK(cudaMemcpyAsync(hb, S.out.as<char>() + o, m, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
                S.wr.put(hb, m);
            }
            n_reads += n; lap(5);
        }
        // 4. carry the unconsumed tail to the front of the other text buffer
        for (int f = 0; f < S.nfiles; f++) {
            KjParsed& P = S.side[f]; uint64_t consumed = 0;
            if (P.n_lines) { if ((rc = kj_parse_consumed(P, n, 0, st, consumed))) return rc; }
            const uint64_t tail = P.nbytes - consumed; const int other = P.cur ^ 1;
            if ((rc = P.text[other].need(tail + chunk + 1))) return rc;
            if (tail) CK(cudaMemcpyAsync(P.text[other].p, P.text[P.cur].as<char>() + consumed, tail, cudaMemcpyDeviceToDevice, st));
            P.cur = other; P.nbytes = tail;
        }
        if (all_eof) {
            if (paired && S.side[1].n_rec > n) fprintf(stderr, "Warning: File %s has more reads then file %s\n", fn[1].c_str(), fn[0].c_str());        // kaiju.cpp:400-404
            break;
        }
        // file 1 is exhausted (end of file, nothing carried over): the reference's loop ends here, whatever file 2 still holds (kaiju.cpp:288, 396-404)
        if (paired && S.side[0].eof && S.side[0].nbytes == 0) {
            if (S.side[1].nbytes > 0 || !S.side[1].eof) fprintf(stderr, "Warning: File %s has more reads then file %s\n", fn[1].c_str(), fn[0].c_str());
            break;
        }
    }
    CK(cudaMemcpyAsync(&n_class, d_nclass, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    if (ftrace) fprintf(stderr, "KJ_FILES_TRACE chunks %llu  wait-chunk+h2d %.1f  prefetch-issue(reader wait) %.1f  parse %.1f  classify %.1f  format %.1f  d2h+writer %.1f  carry %.1f ms\n", (unsigned long long)nchunks, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[7]);
    if (n_reads_out) *n_reads_out = n_reads; if (n_class_out) *n_class_out = n_class;
    return KJ_OK;
}

extern "C" int kj_classify_files(kj_ctx* c, const char* in1, const char* in2, const char* out_path, int verbose, uint64_t* n_reads, uint64_t* n_classified) {
    if (!c || !in1 || !*in1) { kj_err() = "kj_classify_files: null argument"; return KJ_ERR_ARG; }
    if (c->params.input_is_protein && in2 && *in2) { kj_err() = "Protein input only supports one input file."; return KJ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    KjFilesState* S = new KjFilesState();
    int rc = kj_classify_files_impl(c, *S, in1, in2, out_path, verbose, n_reads, n_classified);
    S->cleanup();
    delete S;
    return rc;
}
