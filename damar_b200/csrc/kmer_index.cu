// kmer_index.cu -- stage 0 (block residency) and stage 1 (k-mer tuple extraction + sort).
//
// Replaces, on the device:
//   Read_All_Sequences' unpacked layout -> 2 bit/base packed words     (db/DB.c:1547-1608)
//   complement_DB                         -> revcomp_packed / mirror_mask (daligner.c:511-628)
//   tuple_thread                          -> extract_tuples             (filter.c:458-547)
//   lex_sort + filler removal + -t        -> radix_sort_pairs + suppress (filter.c:328-435, 855-939)
//
// A k-mer entry is (code, gpos): gpos indexes the reference's unpacked `bases` array and names
// the k-mer's LAST base, so (read, rpos) = (r, gpos - boff[r]).  Sorting stably on code with
// entries generated in gpos order reproduces the reference's (code, read, rpos) order.
#include <algorithm>
#include <mutex>
#include <thread>
#include <type_traits>

#include <memory>

#include "common.cuh"

namespace damar {

DeviceBlock::~DeviceBlock() {
    for (int i = 0; i < 2; i++) {
        if (packed[i]) dev_free(packed[i] - PACK_FRONT_PAD);
        if (mask_anno[i]) dev_free(mask_anno[i]);
        if (mask_data[i]) dev_free(mask_data[i]);
    }
    if (boff) dev_free(boff);
    if (rlen) dev_free(rlen);
    if (rtab) dev_free(rtab);
}
DeviceIndex::~DeviceIndex() {
    if (codes) dev_free(codes);
    if (gpos) dev_free(gpos);
}

namespace {

// ---- stage 0 kernels ----

// 16 bases per 32-bit word, first base in the top two bits (big-endian within the word) so
// that a window read with funnel shifts has its first base most significant (filter.c:501-505)
__global__ void pack_bases(const uint8_t *__restrict__ bytes, int64_t nbytes, uint32_t *__restrict__ words, int64_t nwords) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    int64_t p = w * 16;
    uint32_t out = 0;
    if (p + 16 <= nbytes) {
        uint4 v = *reinterpret_cast<const uint4 *>(bytes + p);
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int b = 0; b < 4; b++) out = (out << 2) | ((x[j] >> (8 * b)) & 3u);
    } else {
        for (int i = 0; i < 16; i++) {
            uint32_t c = (p + i < nbytes) ? bytes[p + i] : 0u;
            out = (out << 2) | (c & 3u);
        }
    }
    words[w] = out;
}

__device__ __forceinline__ uint32_t base_at(const uint32_t *__restrict__ words, int64_t p) {
    return (words[p >> 4] >> (30 - 2 * (int)(p & 15))) & 3u;
}

// upper_bound(boff, p) - 1 : the read whose slot range [boff[r], boff[r+1]) holds p
__device__ __forceinline__ int read_of(const uint32_t *__restrict__ boff, int nreads, uint32_t p) {
    int lo = 0, hi = nreads;   // invariant boff[lo] <= p < boff[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (boff[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// rtab[i] = read whose slot range holds gpos i << RTAB_SHIFT: turns the per-entry binary search over boff of the
// merge kernels (seed_merge.cu) into one table load plus a step or two
__global__ void build_rtab(const uint32_t *__restrict__ boff, int nreads, int64_t nbases, uint32_t *__restrict__ rtab, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = i << RTAB_SHIFT;
    rtab[i] = (uint32_t)read_of(boff, nreads, (uint32_t)(p < nbases ? p : nbases - 1));
}

// reverse-complement every read in packed form (daligner.c:511-527): one thread per output word
__global__ void revcomp_packed(const uint32_t *__restrict__ fwd, uint32_t *__restrict__ rc, int64_t nwords,
                               const uint32_t *__restrict__ boff, const int32_t *__restrict__ rlen, int nreads,
                               int64_t nbases) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    int64_t p0 = w * 16;
    uint32_t out = 0;
    if (p0 < nbases) {
        int r = read_of(boff, nreads, (uint32_t)p0);
        for (int i = 0; i < 16; i++) {
            int64_t p = p0 + i;
            uint32_t c = 0;
            if (p < nbases) {
                while (p >= boff[r + 1]) r++;
                int j = (int)(p - boff[r]);
                if (j < rlen[r]) c = 3u - base_at(fwd, (int64_t)boff[r] + (rlen[r] - 1 - j));
            }
            out = (out << 2) | c;
        }
    }
    rc[w] = out;
}

// mirror mask intervals per read (daligner.c:596-624): one thread per read
__global__ void mirror_mask(const int64_t *__restrict__ anno, const int32_t *__restrict__ data,
                            int32_t *__restrict__ out, const int32_t *__restrict__ rlen, int nreads) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    int64_t lo = anno[r], hi = anno[r + 1];
    int len = rlen[r];
    for (int64_t j = lo; j < hi; j++) out[lo + (hi - 1 - j)] = len - data[j];
}

// ---- stage 1 kernels ----

// per read: number of admissible k-mers; for masked blocks also the per-gap exclusive prefix
// (gap g of read r lives at gap_pref[anno[r]/2 + r + g])
__global__ void count_tuples(const int32_t *__restrict__ rlen, int nreads, int K, const int64_t *__restrict__ anno,
                             const int32_t *__restrict__ data, uint32_t *__restrict__ per_read,
                             uint32_t *__restrict__ gap_pref) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    int len = rlen[r];
    if (anno == nullptr) {
        per_read[r] = (len >= K) ? (uint32_t)(len - K + 1) : 0u;
        return;
    }
    int64_t b = anno[r], f = anno[r + 1];
    int64_t gbase = b / 2 + r;
    uint32_t tot = 0;
    int g = 0;
    for (int64_t a = b; a <= f; a += 2, g++) {       // gaps between intervals, filter.c:487-513
        int p = (a == b) ? 0 : data[a - 1];
        int q = (a == f) ? len : data[a];
        gap_pref[gbase + g] = tot;
        if (p + K <= q) tot += (uint32_t)(q - (p + K - 1));
    }
    per_read[r] = tot;
}

// one thread per slot of `bases`; emits (code, gpos) at its compacted index
__global__ void __launch_bounds__(256) extract_tuples(const uint32_t *__restrict__ words, const uint32_t *__restrict__ boff,
                                                     const int32_t *__restrict__ rlen, int nreads, int64_t nbases, int K,
                                                     const int64_t *__restrict__ anno, const int32_t *__restrict__ data,
                                                     const uint32_t *__restrict__ kbase, const uint32_t *__restrict__ gap_pref,
                                                     uint64_t *__restrict__ codes, uint32_t *__restrict__ gpos, int gbits) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t warp_p0 = p - (threadIdx.x & 31);
    if (warp_p0 >= nbases) return;
    int r = read_of(boff, nreads, (uint32_t)warp_p0);     // same address in every lane: broadcast loads
    if (p >= nbases) return;
    while ((uint32_t)p >= boff[r + 1]) r++;
    const int rpos = (int)(p - boff[r]);
    const int len = rlen[r];
    if (rpos >= len || rpos < K - 1) return;
    uint32_t idx;
    if (anno == nullptr) {
        idx = kbase[r] + (uint32_t)(rpos - (K - 1));
    } else {
        // locate the gap [pp,q) holding rpos: j = #intervals with beg <= rpos
        const int64_t b = anno[r], f = anno[r + 1];
        int lo = 0, hi = (int)((f - b) >> 1);             // intervals [lo,hi)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (data[b + 2 * mid] <= rpos) lo = mid + 1; else hi = mid;
        }
        const int j = lo;
        const int pp = (j > 0) ? data[b + 2 * j - 1] : 0;
        if (rpos < pp) return;                            // inside interval j-1
        if (rpos - pp < K - 1) return;
        idx = kbase[r] + gap_pref[b / 2 + r + j] + (uint32_t)(rpos - pp - (K - 1));
    }
    // 64-bit big-endian window starting at the k-mer's first base
    const int64_t s = p - (K - 1);
    const int64_t w = s >> 4;
    const int o = (int)(s & 15) * 2;
    const uint32_t x0 = words[w], x1 = words[w + 1], x2 = words[w + 2];
    const uint32_t hi32 = __funnelshift_l(x1, x0, o);
    const uint32_t lo32 = __funnelshift_l(x2, x1, o);
    const uint64_t win = ((uint64_t)hi32 << 32) | lo32;
    const uint64_t code = win >> (64 - 2 * K);
    if (gpos) { codes[idx] = code; gpos[idx] = (uint32_t)p; }
    else codes[idx] = (code << gbits) | (uint64_t)p;           // one word per k-mer
}

// -b: variable-length k-mers whose summed -log4 base frequency reaches K "bases" of information
// (biased_tuple_thread, filter.c:549-688).  The window's left edge depends on where the previous window of the
// same stretch ended, so a stretch (a whole read, or a gap between mask intervals) is walked by ONE thread, first
// to count, then to emit.  A masked stretch that reaches q while still growing emits at p == q with the code of
// the bases before it (filter.c:598-599 leaves only the inner loop) -- kept, the unmasked variant leaves the read.
struct BiasTable { int lb0, lb1, lb2, lb3, lognorm, logthresh; };
__device__ __forceinline__ int bias_of(const BiasTable &t, uint32_t x) { return (x & 2u) ? ((x & 1u) ? t.lb3 : t.lb2) : ((x & 1u) ? t.lb1 : t.lb0); }

template <bool EMIT>
__device__ uint32_t biased_stretch(const uint32_t *__restrict__ words, int64_t g0, int p0, int q, bool masked, int K, uint64_t kmask,
                                   const BiasTable &bt, uint64_t *__restrict__ codes, uint32_t *__restrict__ gpos, int gbits, uint32_t n) {
    uint64_t c = 0;
    int a = 0, k = 1, p = p0;
    uint32_t cnt = 0;
    while (p < q) {
        uint32_t x = base_at(words, g0 + p);
        a += bias_of(bt, x); c = (c << 2) | x;
        while (a < bt.lognorm && k < K) {
            if (++p >= q) {
                if (!masked) return cnt;
                break;
            }
            x = base_at(words, g0 + p);
            k++; a += bias_of(bt, x); c = (c << 2) | x;
        }
        for (;;) {
            const int u = a - bias_of(bt, base_at(words, g0 + p - k + 1));
            if (u < bt.lognorm) break;
            a = u; k--;
        }
        if (a > bt.logthresh) {
            if (EMIT) {
                const uint64_t code = (c << (2 * (K - k))) & kmask;
                if (gpos) { codes[n + cnt] = code; gpos[n + cnt] = (uint32_t)(g0 + p); }
                else codes[n + cnt] = (code << gbits) | (uint64_t)(g0 + p);
            }
            cnt++;
        }
        p++;
        a -= bias_of(bt, base_at(words, g0 + p - k));
    }
    return cnt;
}

// one thread per read; EMIT=false fills per_read, EMIT=true writes from kbase[r] on
template <bool EMIT>
__global__ void biased_tuples(const uint32_t *__restrict__ words, const uint32_t *__restrict__ boff, const int32_t *__restrict__ rlen,
                              int nreads, int K, uint64_t kmask, const int64_t *__restrict__ anno, const int32_t *__restrict__ data,
                              BiasTable bt, uint32_t *__restrict__ per_read, const uint32_t *__restrict__ kbase,
                              uint64_t *__restrict__ codes, uint32_t *__restrict__ gpos, int gbits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const int len = rlen[r];
    const int64_t g0 = boff[r];
    uint32_t n = EMIT ? kbase[r] : 0u, n0 = n;
    if (anno == nullptr) {
        n += biased_stretch<EMIT>(words, g0, 0, len, false, K, kmask, bt, codes, gpos, gbits, n);
    } else {
        const int64_t b = anno[r], f = anno[r + 1];
        for (int64_t j = b; j <= f; j += 2) {
            const int p = (j == b) ? 0 : data[j - 1];
            const int q = (j == f) ? len : data[j];
            if (p + K <= q) n += biased_stretch<EMIT>(words, g0, p, q, true, K, kmask, bt, codes, gpos, gbits, n);
        }
    }
    if (!EMIT) per_read[r] = n - n0;
}

// -t suppression (compsize_thread / compress_thread, filter.c:700-751): keep[i] = 1 iff the
// run of equal codes containing i is shorter than `too_frequent`.  A thread only needs to look
// `too_frequent` entries to either side.
// cshift = gbits for packed entries (the code is what is left above), 0 for (code, gpos) pairs
__global__ void mark_suppressed(const uint64_t *__restrict__ codes, int cshift, int64_t n, int too_frequent, uint32_t *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t c = codes[i] >> cshift;
    int run = 1;
    for (int64_t j = i - 1; j >= 0 && run < too_frequent && (codes[j] >> cshift) == c; j--) run++;
    for (int64_t j = i + 1; j < n && run < too_frequent && (codes[j] >> cshift) == c; j++) run++;
    keep[i] = (run < too_frequent) ? 1u : 0u;
}
__global__ void compact_kept(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ gpos, const uint32_t *__restrict__ keep,
                             const uint64_t *__restrict__ off, int64_t n, uint64_t *__restrict__ ocodes, uint32_t *__restrict__ ogpos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    ocodes[off[i]] = codes[i];
    if (gpos) ogpos[off[i]] = gpos[i];
}

}  // namespace

// packed arrays carry PACK_FRONT_PAD zero words in front so that backward 32-base windows
// (reverse wave) may start before base 0
static uint32_t *alloc_packed(int64_t nwords, cudaStream_t st) {
    uint32_t *p = nullptr;
    p = (std::remove_reference<decltype(p)>::type)dev_alloc_bytes(sizeof(uint32_t) * (nwords + PACK_FRONT_PAD));
    CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(uint32_t) * PACK_FRONT_PAD, st));
    return p + PACK_FRONT_PAD;
}

// ------------------------------------------------------------------------------------------------
DeviceBlock *upload_block(const damar_block_desc *d) {
    Context &c = ctx();
    Timer tm(c.stream);
    std::unique_ptr<DeviceBlock> holder(new DeviceBlock());        // freed with its device arrays if an allocation below throws
    DeviceBlock *b = holder.get();
    b->nreads = d->nreads; b->maxlen = d->maxlen; b->totlen = d->totlen; b->ufirst = d->ufirst; b->part = d->part;
    b->sizeof_db = d->sizeof_db;
    b->as_is = (d->flags & DAMAR_BLOCK_AS_IS) != 0;
    memcpy(b->freq, d->freq, sizeof(b->freq));
    b->nbases = d->boff[d->nreads];
    if (b->nbases > 0x7fffffffll) fatal("DB blocks are greater than 2Gbp!");   // filter.c:794-798
    b->h_boff.assign(d->boff, d->boff + d->nreads + 1);
    b->h_rlen.assign(d->rlen, d->rlen + d->nreads);

    std::vector<uint32_t> boff32(d->nreads + 1);
    for (int i = 0; i <= d->nreads; i++) boff32[i] = (uint32_t)d->boff[i];
    b->boff = (std::remove_reference<decltype(b->boff)>::type)dev_alloc_bytes(sizeof(uint32_t) * (d->nreads + 1));
    b->rlen = (std::remove_reference<decltype(b->rlen)>::type)dev_alloc_bytes(sizeof(int32_t) * (d->nreads + 1));
    CUDA_CHECK(cudaMemcpyAsync(b->boff, boff32.data(), sizeof(uint32_t) * (d->nreads + 1), cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(b->rlen, d->rlen, sizeof(int32_t) * d->nreads, cudaMemcpyHostToDevice, c.stream));

    // bases: ship the unpacked bytes once, pack on the device, drop the bytes
    const int64_t nbytes = b->nbases;
    b->npacked_words = (nbytes + 15) / 16 + 8;            // slack words so window reads never fault
    uint8_t *dbytes = nullptr;
    dbytes = (std::remove_reference<decltype(dbytes)>::type)dev_alloc_bytes((size_t)nbytes + 64);
    {
        // The caller's buffer is ordinary pageable memory: copied through two pinned 8 MB buffers so that the host copy of
        // one chunk overlaps the DMA of the previous one (a plain cudaMemcpy from pageable memory ran at 1.2 GB/s here)
        WorkerScratch &ws = scratch();
        const size_t CH = 8u << 20;
        for (int i = 0; i < 2; i++) {
            ws.h_up[i].reserve(CH);
            if (!ws.up_ev[i]) CUDA_CHECK(cudaEventCreateWithFlags(&ws.up_ev[i], cudaEventDisableTiming | (blocking_sync() ? cudaEventBlockingSync : 0)));
        }
        int k = 0;
        for (size_t off = 0; off < (size_t)nbytes; off += CH, k ^= 1) {
            const size_t len = std::min(CH, (size_t)nbytes - off);
            if (off >= 2 * CH) CUDA_CHECK(cudaEventSynchronize(ws.up_ev[k]));
            {                                       // four host threads per chunk: one core copies at 3-4 GB/s only
                const int NT = 4;
                std::thread th[NT - 1];
                const size_t part = (len + NT - 1) / NT;
                uint8_t *dst = ws.h_up[k].p; const uint8_t *src = d->bases + off;
                for (int t = 1; t < NT; t++) {
                    const size_t lo = std::min(len, part * t), hi = std::min(len, part * (t + 1));
                    th[t - 1] = std::thread([=] { if (hi > lo) memcpy(dst + lo, src + lo, hi - lo); });
                }
                memcpy(dst, src, std::min(len, part));
                for (int t = 1; t < NT; t++) th[t - 1].join();
            }
            CUDA_CHECK(cudaMemcpyAsync(dbytes + off, ws.h_up[k].p, len, cudaMemcpyHostToDevice, c.stream));
            CUDA_CHECK(cudaEventRecord(ws.up_ev[k], c.stream));
        }
    }
    b->packed[0] = alloc_packed(b->npacked_words, c.stream);
    {
        int64_t nw = b->npacked_words;
        pack_bases<<<(unsigned)((nw + 255) / 256), 256, 0, c.stream>>>(dbytes, nbytes, b->packed[0], nw);
        KERNEL_CHECK(); count_launch();
    }
    {
        const int64_t nt = (b->nbases >> RTAB_SHIFT) + 1;
        b->rtab = (std::remove_reference<decltype(b->rtab)>::type)dev_alloc_bytes(sizeof(uint32_t) * nt);
        build_rtab<<<(unsigned)((nt + 255) / 256), 256, 0, c.stream>>>(b->boff, b->nreads, b->nbases, b->rtab, nt);
        KERNEL_CHECK(); count_launch();
    }
    if (d->mask_anno != nullptr) {
        b->has_mask = true;
        b->mask_len = d->mask_anno[d->nreads];
        b->mask_anno[0] = (std::remove_reference<decltype(b->mask_anno[0])>::type)dev_alloc_bytes(sizeof(int64_t) * (d->nreads + 1));
        b->mask_data[0] = (std::remove_reference<decltype(b->mask_data[0])>::type)dev_alloc_bytes(sizeof(int32_t) * (b->mask_len + 1));
        CUDA_CHECK(cudaMemcpyAsync(b->mask_anno[0], d->mask_anno, sizeof(int64_t) * (d->nreads + 1), cudaMemcpyHostToDevice, c.stream));
        if (b->mask_len > 0)
            CUDA_CHECK(cudaMemcpyAsync(b->mask_data[0], d->mask_data, sizeof(int32_t) * b->mask_len, cudaMemcpyHostToDevice, c.stream));
    }
    CUDA_CHECK(stream_wait(c.stream));
    dev_free(dbytes);
    c.tm.upload_ms = tm.stop();
    return holder.release();
}

// ---- device-to-device block transfer (damar_block_view, include/damar_b200.h) ----
void block_view(DeviceBlock *b, damar_block_view *v) {
    memset(v, 0, sizeof(*v));
    v->nreads = b->nreads; v->maxlen = b->maxlen; v->ufirst = b->ufirst; v->part = b->part;
    v->totlen = b->totlen; v->nbases = b->nbases; v->sizeof_db = b->sizeof_db;
    memcpy(v->freq, b->freq, sizeof(v->freq));
    v->npacked_words = b->npacked_words;
    v->packed = b->packed[0] - PACK_FRONT_PAD; v->boff = b->boff; v->rlen = b->rlen;
    v->mask_len = b->mask_len;
    v->mask_anno = b->has_mask ? b->mask_anno[0] : nullptr;
    v->mask_data = b->has_mask ? b->mask_data[0] : nullptr;
    v->flags = b->as_is ? DAMAR_BLOCK_AS_IS : 0;
}
DeviceBlock *block_from_view(const damar_block_view *v) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    std::unique_ptr<DeviceBlock> holder(new DeviceBlock());
    DeviceBlock *b = holder.get();
    b->nreads = v->nreads; b->maxlen = v->maxlen; b->ufirst = v->ufirst; b->part = v->part;
    b->totlen = v->totlen; b->nbases = v->nbases; b->sizeof_db = v->sizeof_db;
    memcpy(b->freq, v->freq, sizeof(b->freq));
    b->as_is = (v->flags & DAMAR_BLOCK_AS_IS) != 0;
    b->npacked_words = v->npacked_words;
    uint32_t *pk = (uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (b->npacked_words + PACK_FRONT_PAD));
    CUDA_CHECK(cudaMemcpyAsync(pk, v->packed, sizeof(uint32_t) * (b->npacked_words + PACK_FRONT_PAD), cudaMemcpyDeviceToDevice, st));
    b->packed[0] = pk + PACK_FRONT_PAD;
    b->boff = (uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (b->nreads + 1));
    b->rlen = (int32_t *)dev_alloc_bytes(sizeof(int32_t) * (b->nreads + 1));
    CUDA_CHECK(cudaMemcpyAsync(b->boff, v->boff, sizeof(uint32_t) * (b->nreads + 1), cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(b->rlen, v->rlen, sizeof(int32_t) * b->nreads, cudaMemcpyDeviceToDevice, st));
    std::vector<uint32_t> hb((size_t)b->nreads + 1);
    b->h_rlen.resize((size_t)b->nreads);
    CUDA_CHECK(cudaMemcpyAsync(hb.data(), b->boff, sizeof(uint32_t) * (b->nreads + 1), cudaMemcpyDeviceToHost, st));
    if (b->nreads) CUDA_CHECK(cudaMemcpyAsync(b->h_rlen.data(), b->rlen, sizeof(int32_t) * b->nreads, cudaMemcpyDeviceToHost, st));
    {
        const int64_t nt = (b->nbases >> RTAB_SHIFT) + 1;
        b->rtab = (uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * nt);
        build_rtab<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(b->boff, b->nreads, b->nbases, b->rtab, nt);
        KERNEL_CHECK(); count_launch();
    }
    if (v->mask_anno != nullptr) {
        b->has_mask = true; b->mask_len = v->mask_len;
        b->mask_anno[0] = (int64_t *)dev_alloc_bytes(sizeof(int64_t) * (b->nreads + 1));
        b->mask_data[0] = (int32_t *)dev_alloc_bytes(sizeof(int32_t) * (b->mask_len + 1));
        CUDA_CHECK(cudaMemcpyAsync(b->mask_anno[0], v->mask_anno, sizeof(int64_t) * (b->nreads + 1), cudaMemcpyDeviceToDevice, st));
        if (b->mask_len > 0) CUDA_CHECK(cudaMemcpyAsync(b->mask_data[0], v->mask_data, sizeof(int32_t) * b->mask_len, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_CHECK(stream_wait(st));
    b->h_boff.assign(hb.begin(), hb.end());
    return holder.release();
}

void ensure_revcomp(DeviceBlock *b) {
    static std::mutex mu;                        // several workers may want the same block's complement at once
    std::lock_guard<std::mutex> lock(mu);
    if (b->packed[1]) return;
    Context &c = ctx();
    if (b->nreads == 0) { b->packed[1] = alloc_packed(b->npacked_words, c.stream); CUDA_CHECK(cudaMemsetAsync(b->packed[1], 0, sizeof(uint32_t) * b->npacked_words, c.stream)); CUDA_CHECK(stream_wait(c.stream)); return; }
    b->packed[1] = alloc_packed(b->npacked_words, c.stream);
    int64_t nw = b->npacked_words;
    revcomp_packed<<<(unsigned)((nw + 255) / 256), 256, 0, c.stream>>>(b->packed[0], b->packed[1], nw, b->boff, b->rlen,
                                                                      b->nreads, b->nbases);
    KERNEL_CHECK(); count_launch();
    if (b->has_mask) {
        b->mask_data[1] = (std::remove_reference<decltype(b->mask_data[1])>::type)dev_alloc_bytes(sizeof(int32_t) * (b->mask_len + 1));
        b->mask_anno[1] = nullptr;   // offsets are identical; reuse [0]
        mirror_mask<<<(b->nreads + 255) / 256, 256, 0, c.stream>>>(b->mask_anno[0], b->mask_data[0], b->mask_data[1], b->rlen, b->nreads);
        KERNEL_CHECK(); count_launch();
    }
    CUDA_CHECK(stream_wait(c.stream));  // other workers' streams may read it as soon as the lock drops
}

DeviceIndex *build_index(DeviceBlock *b, int comp) {
    Context &c = ctx();
    const FilterParams &fp = c.fp;
    if (!fp.set) fatal("Sort_Kmers called before Set_Filter_Params");
    const int K = fp.kmer;
    cudaStream_t st = c.stream;
    comp = comp ? 1 : 0;
    if (b->nreads == 0) {                                 // an empty block has no k-mers (Sort_Kmers returns NULL, filter.c:990-993)
        DeviceIndex *e = new DeviceIndex();
        e->blk = b; e->comp = comp; e->len = 0;
        c.tm.extract_ms = c.tm.ksort_ms = 0; c.tm.kmers = 0;
        return e;
    }
    if (comp) ensure_revcomp(b);

    Timer t_ex(st);
    const int64_t *anno = b->has_mask ? b->mask_anno[0] : nullptr;
    const int32_t *data = b->has_mask ? b->mask_data[comp] : nullptr;

    DevBuf<uint32_t> per_read, gap_pref;
    DevBuf<uint64_t> total;
    per_read.reserve((size_t)b->nreads + 1);
    total.reserve(1);
    if (b->has_mask && !fp.biased) gap_pref.reserve((size_t)(b->mask_len / 2 + b->nreads + 1));
    BiasTable bt{};
    if (fp.biased) {
        int lb[4];
        bias_log_base(b->freq, lb);              // latched at the first Sort_Kmers after Set_Filter_Params (filter.c:774)
        bt = BiasTable{lb[0], lb[1], lb[2], lb[3], 10000 * K, 10000 * (K - 2)};   // MAX_BIAS = 2 (filter.c:69)
        biased_tuples<false><<<(b->nreads + 63) / 64, 64, 0, st>>>(b->packed[comp], b->boff, b->rlen, b->nreads, K, fp.kmask, anno, data,
                                                                    bt, per_read.p, nullptr, nullptr, nullptr, 0);
    } else {
        count_tuples<<<(b->nreads + 255) / 256, 256, 0, st>>>(b->rlen, b->nreads, K, anno, data, per_read.p, gap_pref.p);
    }
    KERNEL_CHECK(); count_launch();
    // kbase[r] = exclusive prefix of per-read counts (32-bit is enough: kmers < 2^31, filter.c:762)
    uint32_t total32 = 0;
    {
        DevBuf<uint32_t> &tot32 = scratch().tot32_index; tot32.reserve(1);
        exclusive_scan_u32_inplace(per_read.p, b->nreads, tot32.p, st);
        CUDA_CHECK(cudaMemcpyAsync(&total32, tot32.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
    }
    int64_t kmers = total32;
    std::unique_ptr<DeviceIndex> iholder(new DeviceIndex());
    DeviceIndex *idx = iholder.get();
    idx->blk = b; idx->comp = comp; idx->len = 0;
    if (kmers <= 0) { c.tm.extract_ms = t_ex.stop(); c.tm.ksort_ms = 0; c.tm.kmers = 0; return iholder.release(); }

    // one 64-bit word per k-mer (code << gbits | gpos) whenever that fits: 8-byte records through the sort and the merge
    int gbits = 1;
    while ((1ll << gbits) < b->nbases + 1) gbits++;
    static const bool pack_ok = !(getenv("DAMAR_PACK") && atoi(getenv("DAMAR_PACK")) == 0);
    const bool packed = pack_ok && 2 * K + gbits <= 64;
    uint64_t *k0, *k1; uint32_t *g0 = nullptr, *g1 = nullptr;
    FreeOnUnwind raw;
    k0 = raw.add((uint64_t *)dev_alloc_bytes(sizeof(uint64_t) * (kmers + 2)));
    k1 = raw.add((uint64_t *)dev_alloc_bytes(sizeof(uint64_t) * (kmers + 2)));
    if (!packed) {
        g0 = raw.add((uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (kmers + 4)));
        g1 = raw.add((uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (kmers + 4)));
    }
    if (fp.biased)
        biased_tuples<true><<<(b->nreads + 63) / 64, 64, 0, st>>>(b->packed[comp], b->boff, b->rlen, b->nreads, K, fp.kmask, anno, data,
                                                                   bt, nullptr, per_read.p, k0, g0, gbits);
    else
        extract_tuples<<<(unsigned)((b->nbases + 255) / 256), 256, 0, st>>>(b->packed[comp], b->boff, b->rlen, b->nreads, b->nbases, K,
                                                                            anno, data, per_read.p, gap_pref.p, k0, g0, gbits);
    KERNEL_CHECK(); count_launch();
    c.tm.extract_ms = t_ex.stop();

    Timer t_sort(st);
    if ((packed ? radix_sort(k0, g0, k1, g1, kmers, gbits, gbits + 2 * K, st) : radix_sort(k0, g0, k1, g1, kmers, 0, 2 * K, st)) != 0)
        fatal("Sort_Kmers: %s", damar_last_error());

    if (fp.suppress > 0) {                                // -t
        DevBuf<uint32_t> keep; DevBuf<uint64_t> off;
        keep.reserve((size_t)kmers); off.reserve((size_t)kmers);
        mark_suppressed<<<(unsigned)((kmers + 255) / 256), 256, 0, st>>>(k0, packed ? gbits : 0, kmers, fp.suppress, keep.p);
        KERNEL_CHECK();
        exclusive_scan_u32_to_u64(keep.p, off.p, kmers, total.p, st);
        compact_kept<<<(unsigned)((kmers + 255) / 256), 256, 0, st>>>(k0, g0, keep.p, off.p, kmers, k1, g1);
        KERNEL_CHECK(); count_launch(2);
        uint64_t kept = 0;
        CUDA_CHECK(cudaMemcpyAsync(&kept, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        kmers = (int64_t)kept;
        std::swap(k0, k1); std::swap(g0, g1);
    }
    c.tm.ksort_ms = t_sort.stop();
    raw.dismiss();
    dev_free(k1); if (g1) dev_free(g1);
    idx->codes = k0; idx->gpos = g0; idx->gbits = packed ? gbits : 0; idx->len = kmers;
    c.tm.kmers = kmers;
    return iholder.release();
}

}  // namespace damar
