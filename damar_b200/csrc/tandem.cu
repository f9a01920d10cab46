// tandem.cu -- datander's seed stage (reference scrub/tandem.c): self-comparison of every read for tandem repeats.
//
// The reference sorts all k-mers of a block by code (Sort_Kmers, tandem.c:428), rewrites every entry's code into the
// distance to the previous entry with the same k-mer IN THE SAME READ (count_thread, :556-590; 0 = none), sorts the list
// back into (read, rpos) order (:1281-1300) and lets report_thread walk each read's positions (:895).  Here:
//   self_links   one thread per entry of the code-sorted list (the stage-1 index of kmer_index.cu, entries in position
//                order inside a run): an entry whose predecessor has the same code and lies in the same read writes
//                link[gpos] = gpos - gpos_prev -- a scatter by position replaces the second sort
//   link_hits    the non-zero links in position order ARE the seed hits of the pair (read, read) in apos order:
//                compacted into the hit format of seed_merge.cu (bread | aread | apos | bpos), already sorted
// so stages 3 and 4 (pair_align.cu, tandem mode: panels over the read's own positions, every read, reach = 0, fusion
// without bridging) run unchanged on them.  Positions follow tandem.c:412: apos = index of the k-mer's last base + 1.
// One deviation, undefined in the reference: count_thread never rewrites the first entry of a thread's range (:588 zeroes
// the NEXT run's head), so report_thread would use a raw k-mer code as a distance there; such an entry is a run head,
// "no link", which is what it gets here (and in the oracle).
#include <type_traits>

#include "common.cuh"

namespace damar {
namespace {

__device__ __forceinline__ int read_of_tab(const uint32_t *__restrict__ boff, const uint32_t *__restrict__ rtab, uint32_t p) {
    int r = (int)rtab[p >> RTAB_SHIFT];
    while (boff[r + 1] <= p) r++;
    return r;
}

__global__ void self_links(const uint64_t *__restrict__ ent, const uint32_t *__restrict__ gpos, int gbits, int64_t n,
                           const uint32_t *__restrict__ boff, const uint32_t *__restrict__ rtab, uint32_t *__restrict__ link) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    const uint64_t e = ent[i], f = ent[i - 1];
    const uint64_t ce = gpos ? e : (e >> gbits), cf = gpos ? f : (f >> gbits);
    if (ce != cf) return;
    const uint32_t pe = gpos ? gpos[i] : (uint32_t)(e & ((1ull << gbits) - 1ull));
    const uint32_t pf = gpos ? gpos[i - 1] : (uint32_t)(f & ((1ull << gbits) - 1ull));
    const int r = read_of_tab(boff, rtab, pe);
    if (pf >= boff[r]) link[pe] = pe - pf;                 // same read: runs are in position order
}

__global__ void link_flags(const uint32_t *__restrict__ link, int64_t n, uint32_t *__restrict__ flag) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flag[p] = link[p] != 0u;
}

__global__ void link_hits(const uint32_t *__restrict__ link, const uint64_t *__restrict__ off, int64_t n,
                          const uint32_t *__restrict__ boff, const uint32_t *__restrict__ rtab, int pbits, int abits, int qbits,
                          uint64_t *__restrict__ keys, uint32_t *__restrict__ bpos) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t d = link[p];
    if (d == 0) return;
    const int r = read_of_tab(boff, rtab, (uint32_t)p);
    const uint32_t apos = (uint32_t)p - boff[r] + 1u;       // rpos of tandem.c:412
    const uint64_t key = ((((uint64_t)r << abits) | (uint64_t)r) << pbits) | apos;
    const uint64_t o = off[p];
    if (bpos) { keys[o] = key; bpos[o] = apos - d; }
    else keys[o] = (key << qbits) | (uint64_t)(apos - d);
}

int bits_for(int64_t maxval_exclusive) {
    int b = 1;
    while ((1ll << b) < maxval_exclusive) b++;
    return b;
}

}  // namespace

int self_hits(const DeviceIndex *idx, SeedHits &out) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    const DeviceBlock *b = idx->blk;
    out.release();
    out.limit = INT32_MAX;
    Timer t_merge(st);
    const int64_t nb = b->nbases + 1;
    DevBuf<uint32_t> link, flag; DevBuf<uint64_t> off, total;
    link.reserve((size_t)nb); flag.reserve((size_t)nb); off.reserve((size_t)nb); total.reserve(1);
    CUDA_CHECK(cudaMemsetAsync(link.p, 0, sizeof(uint32_t) * (size_t)nb, st));
    self_links<<<(unsigned)((idx->len + 255) / 256), 256, 0, st>>>(idx->codes, idx->gpos, idx->gbits, idx->len, b->boff, b->rtab, link.p);
    KERNEL_CHECK();
    link_flags<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(link.p, nb, flag.p);
    KERNEL_CHECK();
    exclusive_scan_u32_to_u64(flag.p, off.p, nb, total.p, st);
    uint64_t nhits = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nhits, total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    count_launch(2);
    if (c.opt.verbose) { printf("   Hit count = %lld\n", (long long)nhits); fflush(stdout); }
    if (nhits == 0) { c.tm.merge_ms = t_merge.stop(); return 0; }

    out.pbits = bits_for((int64_t)b->maxlen + 2);          // apos runs up to the read length
    out.abits = bits_for(b->nreads);
    out.bbits = out.abits;
    out.qbits = out.pbits;
    const int kbits = out.pbits + out.abits + out.bbits;
    if (kbits > 64) { set_error("seed-hit key needs %d bits (> 64): block too large for this build", kbits); return 1; }
    static const bool pack_ok = !(getenv("DAMAR_PACK") && atoi(getenv("DAMAR_PACK")) == 0);
    const bool packed = pack_ok && kbits + out.qbits <= 64;
    FreeOnUnwind raw;
    uint64_t *k0 = raw.add((uint64_t *)dev_alloc_bytes(sizeof(uint64_t) * (nhits + 2)));
    uint32_t *v0 = packed ? nullptr : raw.add((uint32_t *)dev_alloc_bytes(sizeof(uint32_t) * (nhits + 4)));
    raw.dismiss();                                          // nothing below allocates
    link_hits<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(link.p, off.p, nb, b->boff, b->rtab, out.pbits, out.abits, out.qbits, k0, v0);
    KERNEL_CHECK(); count_launch();
    c.tm.merge_ms = t_merge.stop();
    c.tm.hsort_ms = 0;                                      // generated in (read, apos) order: nothing to sort
    out.keys = k0; out.bpos = v0; out.n = (int64_t)nhits;
    return 0;
}

}  // namespace damar
