// trace_pts.cu -- Compute_Trace_PTS in its cost form: the exact edit distance of every trace-point segment.
//
// Replaces, for the numbers it produces, Compute_Trace_PTS + iter_np (reference dalign/align.c:5577-5692,
// 4892-5230): the consumers of a .las file (LAshow utils/LAshow.c:261, LAcorrect, LAstitch ...) re-align every
// segment between two trace points -- A[ab, ae) with ae on the trace grid, against B[bb, bb + b_i) -- and leave
// the sum of the segment distances in path->diffs.  iter_np counts a substitution, an insertion and a deletion
// as one difference each (it returns D + |M - N| with steps toward the end diagonal free, steps away from it 2
// and mismatches 1), i.e. the unit-cost edit distance, which does not depend on its tie-breaking; the edit
// script it also writes does, and is not produced here.
//
// One warp per record, one lane per segment (the B start of a segment is a warp prefix sum of the b_i), each
// lane running the furthest-reaching recurrence over its own diagonals with 32-base packed compares:
//     F[e][k] = slide(max(F[e-1][k] + 1, F[e-1][k-1] + 1, F[e-1][k+1])),   k = x - y,  done when F[e][M-N] = M.
// Doubles as the on-device validator of the wave kernel's output: every emitted (d_i, b_i) is realisable, and
// optimal unless the record went through bridging (tests/test_gpu_parity.py).
#include <vector>

#include "common.cuh"
#include "trace_script_core.cuh"
#include "wave_core.cuh"

namespace damar {
namespace {

constexpr int TP_EMAX = 1023;                 // largest segment distance handled (8-bit traces hold 255 at most)

struct TpArgs {
    const uint32_t *aw, *bw[2];               // packed A reads; B reads forward / reverse-complemented
    const uint32_t *aboff, *bboff; const int32_t *arlen, *brlen;
    int afirst, bfirst, anreads, bnreads;
    const damar_overlap *ovl; const uint64_t *toff; const void *trace; int tbytes; int64_t novl; int TS;
    int32_t *rec; uint16_t *seg; int *err;
};

__device__ __forceinline__ int trace_at(const TpArgs &P, uint64_t i) {
    return P.tbytes == 1 ? (int)((const uint8_t *)P.trace)[i] : (int)((const uint16_t *)P.trace)[i];
}

// unit-cost edit distance of A[0, M) (base index ax) and B[0, N) (base index bx); -1 beyond TP_EMAX
__device__ int segment_distance(const uint32_t *__restrict__ aw, uint32_t ax, int M, const uint32_t *__restrict__ bw, uint32_t bx, int N,
                                int16_t *F0, int16_t *F1) {
    const int del = M - N;
    // e = 0
    {
        const int lim = M < N ? M : N;
        const int x = wc::slide(aw, ax, bw, bx, lim);
        if (x == M && del == 0) return 0;
        F0[TP_EMAX + 1] = (int16_t)x;
    }
    int16_t *prev = F0, *cur = F1;
    for (int e = 1; e <= TP_EMAX; e++) {
        const int klo = (-e < -N) ? -N : -e, khi = (e > M) ? M : e;
        for (int k = klo; k <= khi; k++) {
            // predecessors live on diagonals within [-(e-1), e-1] and inside the box
            const int plo = (-(e - 1) < -N) ? -N : -(e - 1), phi = ((e - 1) > M) ? M : (e - 1);
            int x = -1;
            if (k >= plo && k <= phi) x = prev[TP_EMAX + 1 + k] + 1;                       // substitution
            if (k - 1 >= plo && k - 1 <= phi) { const int v = prev[TP_EMAX + 1 + k - 1] + 1; if (v > x) x = v; }   // A base alone
            if (k + 1 >= plo && k + 1 <= phi) { const int v = prev[TP_EMAX + 1 + k + 1]; if (v > x) x = v; }       // B base alone
            // clamp to the box: x <= M, y = x - k <= N, and y >= 0
            if (x > M) x = M;
            if (x - k > N) x = N + k;
            if (x < k) x = k;
            if (x < 0) x = 0;
            const int y = x - k;
            const int lim = (M - x < N - y) ? M - x : N - y;
            if (lim > 0) x += wc::slide(aw, ax + (uint32_t)x, bw, bx + (uint32_t)y, lim);
            cur[TP_EMAX + 1 + k] = (int16_t)x;
        }
        if (del >= klo && del <= khi && cur[TP_EMAX + 1 + del] >= M) return e;
        int16_t *t = prev; prev = cur; cur = t;
    }
    return -1;
}

__global__ void __launch_bounds__(128) trace_point_diffs(TpArgs P, int16_t *__restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int16_t *F0 = scratch + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2 * (2 * TP_EMAX + 3);
    int16_t *F1 = F0 + (2 * TP_EMAX + 3);
    for (int64_t r = gw; r < P.novl; r += nw) {
        const damar_overlap o = P.ovl[r];
        const int ar = o.aread - P.afirst, br = o.bread - P.bfirst;
        if (ar < 0 || ar >= P.anreads || br < 0 || br >= P.bnreads) { if (lane == 0) { atomicExch(P.err, 1); P.rec[r] = -1; } continue; }
        const int comp = (int)(o.flags & 1u);
        const uint32_t *bw = P.bw[comp];
        const uint32_t abase = P.aboff[ar], bbase = P.bboff[br];
        const int alen = P.arlen[ar], blen = P.brlen[br];
        const int nseg = o.tlen / 2;
        const uint64_t t0 = P.toff[r];
        int total = 0, bb0 = o.bbpos;
        bool bad = (o.aepos > alen || o.bepos > blen || o.abpos < 0 || o.bbpos < 0);      // TP_Error, align.c:5667
        if (nseg == 0) bad = true;                                                        // -T records carry no trace points
        for (int s0 = 0; s0 < nseg; s0 += 32) {
            const int s = s0 + lane;
            const bool v = s < nseg;
            const int bi = v ? trace_at(P, t0 + 2 * (uint64_t)s + 1) : 0;
            int inc = bi;                                                                 // inclusive prefix of the b_i
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            const int bb = bb0 + inc - bi;
            int dist = 0;
            if (v && !bad) {
                const int grid0 = (o.abpos / P.TS) * P.TS;
                const int ab = (s == 0) ? o.abpos : grid0 + s * P.TS;
                const int ae = (s == nseg - 1) ? o.aepos : grid0 + (s + 1) * P.TS;
                const int be = (s == nseg - 1) ? o.bepos : bb + bi;
                if (ae > alen || be > blen || ae < ab || be < bb) bad = true;
                else {
                    dist = segment_distance(P.aw, abase + (uint32_t)ab, ae - ab, bw, bbase + (uint32_t)bb, be - bb, F0, F1);
                    if (dist < 0) { bad = true; dist = 0; }
                    if (P.seg) P.seg[t0 / 2 + (uint64_t)s] = (uint16_t)dist;
                }
            }
            bad = __any_sync(0xffffffffu, bad);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, d);
            total += dist;
            bb0 += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) { P.rec[r] = bad ? -1 : total; if (bad) atomicExch(P.err, 1); }
    }
}

// ---- the edit script (Compute_Trace_PTS proper, align.c:5577-5692; iter_np in GREEDIEST mode via trace_script_core.cuh) ----
struct SegArgs {
    TpArgs P;
    const uint16_t *segcost;                  // exact distance of every segment (trace_point_diffs)
    int32_t *seg_ab, *seg_bb, *seg_m, *seg_n, *seg_dmax, *seg_rec;
    uint32_t *need, *cap;                     // scratch ints / output slots per segment
    const int32_t *dcap;                      // Compute_Trace_MID pieces: waves the scratch was sized for (their distance is not known beforehand)
    int32_t *cost_out;                        // ... and where their distance goes
};
// one warp per record: the segment table (align.c:5650-5680) and what every segment needs
__global__ void __launch_bounds__(128) script_segments(SegArgs S) {
    const TpArgs &P = S.P;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int64_t r = gw; r < P.novl; r += nw) {
        const damar_overlap o = P.ovl[r];
        const int nseg = o.tlen / 2;
        const uint64_t t0 = P.toff[r];
        const bool bad = P.rec[r] < 0;
        int dmax = 0;
        for (int s = lane; s < nseg; s += 32) { const int d = trace_at(P, t0 + 2 * (uint64_t)s); if (d > dmax) dmax = d; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, dmax, d); if (t > dmax) dmax = t; }
        int bb0 = o.bbpos;
        for (int s0 = 0; s0 < nseg; s0 += 32) {
            const int s = s0 + lane;
            const bool v = s < nseg;
            const int bi = v ? trace_at(P, t0 + 2 * (uint64_t)s + 1) : 0;
            int inc = bi;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            if (v) {
                const uint64_t g = t0 / 2 + (uint64_t)s;
                const int bb = bb0 + inc - bi;
                const int grid0 = (o.abpos / P.TS) * P.TS;
                const int ab = (s == 0) ? o.abpos : grid0 + s * P.TS;
                const int ae = (s == nseg - 1) ? o.aepos : grid0 + (s + 1) * P.TS;
                const int be = (s == nseg - 1) ? o.bepos : bb + bi;
                const int M = ae - ab, N = be - bb, del = M - N, c = S.segcost[g];
                const int D = c - (del < 0 ? -del : del);
                S.seg_ab[g] = ab; S.seg_bb[g] = bb; S.seg_m[g] = M; S.seg_n[g] = N; S.seg_dmax[g] = dmax; S.seg_rec[g] = (int32_t)r;
                S.need[g] = bad || D < 0 ? 0u : (uint32_t)ts::scratch_ints(D, del);
                S.cap[g] = bad ? 0u : (uint32_t)c;
            }
            bb0 += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
}
// base of one read at an index relative to a segment start; the read's terminators are 4, anything further out never matches
struct PackedRel {
    const uint32_t *words; uint32_t base; int pos, len; int oob;
    __device__ int operator()(int i) const {
        const int q = pos + i;
        if (q < 0 || q >= len) return (q == -1 || q == len) ? 4 : oob;
        const uint32_t p = base + (uint32_t)q;
        return (int)((words[p >> 4] >> (30 - 2 * (int)(p & 15u))) & 3u);
    }
};
// one thread per segment [g0, g1): iter_np with its tie-breaks
__global__ void __launch_bounds__(128) script_emit(SegArgs S, uint64_t g0, uint64_t g1, const uint64_t *__restrict__ need_off, uint64_t need_base,
                                                  const uint64_t *__restrict__ cap_off, int *__restrict__ scratch, int32_t *__restrict__ raw,
                                                  uint32_t *__restrict__ count) {
    const TpArgs &P = S.P;
    const uint64_t g = g0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g1) return;
    count[g] = 0;
    if (S.need[g] == 0) return;                            // a rejected record, or a slot no live record owns
    const damar_overlap o = P.ovl[S.seg_rec[g]];
    const int ar = o.aread - P.afirst, br = o.bread - P.bfirst;
    const int ab = S.seg_ab[g], bb = S.seg_bb[g], M = S.seg_m[g], N = S.seg_n[g], del = M - N;
    const int D = S.dcap ? S.dcap[g] : (int)S.segcost[g] - (del < 0 ? -del : del);
    const PackedRel A{P.aw, P.aboff[ar], ab, P.arlen[ar], 5}, B{P.bw[o.flags & 1u], P.bboff[br], bb, P.brlen[br], 6};
    int cost = 0;
    const int n = ts::np_script(A, M, B, N, S.seg_dmax[g], D, ab, bb, scratch + (need_off[g] - need_base), raw + cap_off[g], (int)S.cap[g], &cost);
    if (n < 0 || (!S.dcap && cost != (int)S.segcost[g])) { atomicExch(P.err, 2); return; }
    if (S.cost_out) S.cost_out[g] = cost;
    count[g] = (uint32_t)n;
}

// ---- Compute_Trace_MID (align.c:5694-5830): iter_np between the MIDDLES of successive trace-point segments ----
// one thread per trace-point segment [g0, g1): middle_np, the middle as absolute read coordinates
__global__ void __launch_bounds__(128) mid_points(SegArgs S, uint64_t g0, uint64_t g1, const uint64_t *__restrict__ need_off, uint64_t need_base,
                                                 int *__restrict__ scratch, int32_t *__restrict__ mid_a, int32_t *__restrict__ mid_b) {
    const TpArgs &P = S.P;
    const uint64_t g = g0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g1 || S.need[g] == 0) return;
    const damar_overlap o = P.ovl[S.seg_rec[g]];
    const int ar = o.aread - P.afirst, br = o.bread - P.bfirst;
    const int ab = S.seg_ab[g], bb = S.seg_bb[g], M = S.seg_m[g], N = S.seg_n[g], del = M - N;
    const int D = (int)S.segcost[g] - (del < 0 ? -del : del);
    const PackedRel A{P.aw, P.aboff[ar], ab, P.arlen[ar], 5}, B{P.bw[o.flags & 1u], P.bboff[br], bb, P.brlen[br], 6};
    int ma = 0, mb = 0;
    if (ts::np_middle(A, M, B, N, S.seg_dmax[g], D, scratch + (need_off[g] - need_base), &ma, &mb)) { atomicExch(P.err, 2); return; }
    mid_a[g] = ab + ma; mid_b[g] = bb + mb;
}
// pieces per record (segments + 1; 0 for a rejected record), to be scanned into the records' first piece slots
__global__ void mid_piece_counts(TpArgs P, uint32_t *__restrict__ cnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.novl) return;
    cnt[r] = (P.rec[r] < 0 || P.ovl[r].tlen < 2) ? 0u : (uint32_t)(P.ovl[r].tlen / 2 + 1);
}
// one thread per piece slot: piece q of record r runs from the middle of segment q - 1 (or the start of the alignment) to
// the middle of segment q (or its end).  Its distance is at most that of the two half segments it is made of, which sizes
// the scratch; more than dmax waves is the reference's "Bad alignment between trace points".
__global__ void mid_pieces(SegArgs S, SegArgs Q, const uint64_t *__restrict__ pbase, uint64_t npieces, const int32_t *__restrict__ mid_a,
                           const int32_t *__restrict__ mid_b, int32_t *__restrict__ dcap) {
    const TpArgs &P = S.P;
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npieces) return;
    int64_t lo = 0, hi = P.novl - 1;                           // the record that owns slot q: last r with pbase[r] <= q
    while (lo < hi) { const int64_t m = (lo + hi + 1) >> 1; if (pbase[m] <= q) lo = m; else hi = m - 1; }
    const int64_t r = lo;
    const damar_overlap o = P.ovl[r];
    const int nseg = o.tlen / 2, i = (int)(q - pbase[r]);
    const uint64_t gb = P.toff[r] / 2;
    const int a0 = i == 0 ? o.abpos : mid_a[gb + i - 1], b0 = i == 0 ? o.bbpos : mid_b[gb + i - 1];
    const int a1 = i == nseg ? o.aepos : mid_a[gb + i], b1 = i == nseg ? o.bepos : mid_b[gb + i];
    const int M = a1 - a0, N = b1 - b0, del = M - N, ad = del < 0 ? -del : del;
    const int c = (i > 0 ? (int)S.segcost[gb + i - 1] : 0) + (i < nseg ? (int)S.segcost[gb + i] : 0);
    const int dmax = S.seg_dmax[gb];
    int D = c - ad; if (D < 0) D = 0; if (D > dmax) D = dmax;
    Q.seg_ab[q] = a0; Q.seg_bb[q] = b0; Q.seg_m[q] = M; Q.seg_n[q] = N; Q.seg_dmax[q] = dmax; Q.seg_rec[q] = (int32_t)r;
    dcap[q] = D;
    Q.need[q] = (uint32_t)ts::scratch_ints(D, del);
    Q.cap[q] = (uint32_t)c;
}
// path->diffs as Compute_Trace_MID leaves it: the distances of the pieces, the one that ends at the last middle twice
// (`d += iter_np(...); diffs += d;`, align.c:5814-5817)
__global__ void mid_record_diffs(TpArgs P, const uint64_t *__restrict__ pbase, const int32_t *__restrict__ cost, int32_t *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.novl) return;
    if (P.rec[r] < 0 || P.ovl[r].tlen < 2) { out[r] = -1; return; }
    const int nseg = P.ovl[r].tlen / 2;
    int t = 0;
    for (int i = 0; i <= nseg; i++) t += cost[pbase[r] + i];
    out[r] = t + cost[pbase[r] + nseg - 1];
}
__global__ void mid_record_offsets(const uint64_t *__restrict__ pbase, int64_t novl, const uint64_t *__restrict__ out_off, uint64_t npieces, uint64_t total,
                                   int64_t *__restrict__ rec_off) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > novl) return;
    const uint64_t q = (r == novl) ? npieces : pbase[r];
    rec_off[r] = (int64_t)(q >= npieces ? total : out_off[q]);
}
__global__ void script_compact(const int32_t *__restrict__ raw, const uint64_t *__restrict__ cap_off, const uint32_t *__restrict__ count,
                               const uint64_t *__restrict__ out_off, uint64_t nseg, int32_t *__restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nseg) return;
    const uint32_t n = count[g];
    const int32_t *src = raw + cap_off[g];
    int32_t *dst = out + out_off[g];
    for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
}
__global__ void script_record_offsets(const uint64_t *__restrict__ toff, int64_t novl, const uint64_t *__restrict__ out_off, uint64_t nseg, uint64_t total,
                                      int64_t *__restrict__ rec_off) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > novl) return;
    const uint64_t g = (r == novl) ? nseg : toff[r] / 2;
    rec_off[r] = (int64_t)(g >= nseg ? total : out_off[g]);
}

}  // namespace

// rec_diffs[novl]: what Compute_Trace_PTS leaves in path->diffs (-1: trace point out of bounds, align.c:5667);
// seg_diffs (optional): one value per segment, segment i of record r at trace_off[r] / 2 + i.  Returns 0, or 1 if any
// record was out of bounds.
int trace_point_diffs_host(DeviceBlock *ab, DeviceBlock *bb, const damar_overlap *ovl, const uint64_t *toff, const void *trace,
                           int tbytes, int64_t novl, int64_t ntrace, int TS, int32_t *rec_diffs, uint16_t *seg_diffs,
                           int64_t *script_off, int32_t **script_out, int64_t *script_len, int flavour) {
    Context &c = ctx();
    cudaStream_t st = c.stream;
    if (novl <= 0) return 0;
    bool any_comp = false;
    for (int64_t i = 0; i < novl; i++) if (ovl[i].flags & 1u) { any_comp = true; break; }
    if (any_comp) ensure_revcomp(bb);
    DevBuf<damar_overlap> d_ovl; DevBuf<uint64_t> d_toff; DevBuf<uint8_t> d_trace; DevBuf<int32_t> d_rec; DevBuf<uint16_t> d_seg; DevBuf<int> d_err;
    DevBuf<int16_t> d_scr;
    d_ovl.reserve((size_t)novl); d_toff.reserve((size_t)novl); d_trace.reserve((size_t)ntrace * tbytes + 16); d_rec.reserve((size_t)novl);
    d_err.reserve(1);
    if (seg_diffs || script_off) d_seg.reserve((size_t)ntrace / 2 + 1);
    CUDA_CHECK(cudaMemcpyAsync(d_ovl.p, ovl, sizeof(damar_overlap) * novl, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(d_toff.p, toff, sizeof(uint64_t) * novl, cudaMemcpyHostToDevice, st));
    if (ntrace > 0) CUDA_CHECK(cudaMemcpyAsync(d_trace.p, trace, (size_t)ntrace * tbytes, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
    const int blocks = NUM_SMS_B200 * 4, threads = 128;
    d_scr.reserve((size_t)blocks * threads * 2 * (2 * TP_EMAX + 3));
    TpArgs P{};
    P.aw = ab->packed[0]; P.bw[0] = bb->packed[0]; P.bw[1] = any_comp ? bb->packed[1] : bb->packed[0];
    P.aboff = ab->boff; P.bboff = bb->boff; P.arlen = ab->rlen; P.brlen = bb->rlen;
    P.afirst = ab->ufirst; P.bfirst = bb->ufirst; P.anreads = ab->nreads; P.bnreads = bb->nreads;
    P.ovl = d_ovl.p; P.toff = d_toff.p; P.trace = d_trace.p; P.tbytes = tbytes; P.novl = novl; P.TS = TS;
    P.rec = d_rec.p; P.seg = (seg_diffs || script_off) ? d_seg.p : nullptr; P.err = d_err.p;
    trace_point_diffs<<<blocks, threads, 0, st>>>(P, d_scr.p);
    KERNEL_CHECK(); count_launch();
    int err = 0;
    CUDA_CHECK(cudaMemcpyAsync(rec_diffs, d_rec.p, sizeof(int32_t) * novl, cudaMemcpyDeviceToHost, st));
    if (seg_diffs && ntrace > 0) CUDA_CHECK(cudaMemcpyAsync(seg_diffs, d_seg.p, sizeof(uint16_t) * (size_t)(ntrace / 2), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    if (script_off == nullptr) return err;

    // ---- the edit scripts: segment table, per-segment scratch / output slots by prefix sums, one thread per segment ----
    *script_out = nullptr; *script_len = 0;
    const uint64_t nseg = (uint64_t)ntrace / 2;
    if (nseg == 0) { for (int64_t r = 0; r <= novl; r++) script_off[r] = 0; return err; }
    DevBuf<int32_t> tab; DevBuf<uint32_t> need, cap, count; DevBuf<uint64_t> need_off, cap_off, out_off, tot;
    tab.reserve(6 * (size_t)nseg); need.reserve((size_t)nseg + 1); cap.reserve((size_t)nseg + 1); count.reserve((size_t)nseg + 1);
    need_off.reserve((size_t)nseg + 1); cap_off.reserve((size_t)nseg + 1); out_off.reserve((size_t)nseg + 1); tot.reserve(8);
    CUDA_CHECK(cudaMemsetAsync(need.p, 0, sizeof(uint32_t) * nseg, st));
    CUDA_CHECK(cudaMemsetAsync(cap.p, 0, sizeof(uint32_t) * nseg, st));
    SegArgs S{};
    S.P = P; S.segcost = d_seg.p;
    S.seg_ab = tab.p; S.seg_bb = tab.p + nseg; S.seg_m = tab.p + 2 * nseg; S.seg_n = tab.p + 3 * nseg; S.seg_dmax = tab.p + 4 * nseg; S.seg_rec = tab.p + 5 * nseg;
    S.need = need.p; S.cap = cap.p;
    script_segments<<<blocks, threads, 0, st>>>(S);
    KERNEL_CHECK();
    exclusive_scan_u32_to_u64(need.p, need_off.p, (int64_t)nseg, tot.p, st);
    exclusive_scan_u32_to_u64(cap.p, cap_off.p, (int64_t)nseg, tot.p + 1, st);
    uint64_t htot[2];
    CUDA_CHECK(cudaMemcpyAsync(htot, tot.p, sizeof(htot), cudaMemcpyDeviceToHost, st));
    std::vector<uint64_t> h_need_off((size_t)nseg);
    CUDA_CHECK(cudaMemcpyAsync(h_need_off.data(), need_off.p, sizeof(uint64_t) * nseg, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    const uint64_t budget = (uint64_t)1 << 28;                     // ints of scratch per launch (1 GB)
    DevBuf<int> scr;
    // runs fn(g0, g1, scratch base) over batches of slots whose scratch fits the budget
    auto batches = [&](uint64_t n, const std::vector<uint64_t> &off, uint64_t tot_need, auto fn) {
        scr.reserve((size_t)std::min<uint64_t>(tot_need + 4, budget + ((uint64_t)1 << 24)));
        for (uint64_t g0 = 0; g0 < n;) {
            uint64_t lo = g0 + 1, hi = n;                          // largest g1 with off[g1] - off[g0] <= budget
            while (lo < hi) { const uint64_t m = (lo + hi + 1) >> 1; if (off[m - 1] - off[g0] <= budget || m == g0 + 1) lo = m; else hi = m - 1; }
            const uint64_t g1 = lo;
            const uint64_t span = ((g1 < n) ? off[g1] : tot_need) - off[g0];
            if (span + 4 > scr.n) scr.reserve((size_t)span + 4);
            fn(g0, g1, off[g0]);
            g0 = g1;
        }
    };
    // what the scripts are cut into: the trace-point segments themselves (PTS), or the pieces between their middles (MID)
    SegArgs E = S;
    uint64_t nslot = nseg, cap_total = htot[1];
    DevBuf<int32_t> mid, ptab, dcap, pcost; DevBuf<uint32_t> pneed, pcap, pcnt; DevBuf<uint64_t> pbase, pneed_off;
    std::vector<uint64_t> h_pneed_off;
    const uint64_t *slot_need_off = need_off.p;
    const std::vector<uint64_t> *h_slot_need_off = &h_need_off;
    uint64_t slot_need_total = htot[0];
    if (flavour == 1) {
        mid.reserve(2 * (size_t)nseg);
        batches(nseg, h_need_off, htot[0], [&](uint64_t g0, uint64_t g1, uint64_t base) {
            mid_points<<<(unsigned)((g1 - g0 + 127) / 128), 128, 0, st>>>(S, g0, g1, need_off.p, base, scr.p, mid.p, mid.p + nseg);
            KERNEL_CHECK(); count_launch();
        });
        pcnt.reserve((size_t)novl + 1); pbase.reserve((size_t)novl + 1);
        mid_piece_counts<<<(unsigned)((novl + 255) / 256), 256, 0, st>>>(P, pcnt.p);
        KERNEL_CHECK();
        exclusive_scan_u32_to_u64(pcnt.p, pbase.p, novl, tot.p + 3, st);
        uint64_t npieces = 0;
        CUDA_CHECK(cudaMemcpyAsync(&npieces, tot.p + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        if (err == 2) { set_error("Bad alignment between trace points (Compute_Trace_MID)"); return 2; }
        nslot = npieces;
        if (nslot == 0) { for (int64_t r = 0; r <= novl; r++) script_off[r] = 0; for (int64_t r = 0; r < novl; r++) rec_diffs[r] = -1; return err; }
        ptab.reserve(6 * (size_t)nslot); dcap.reserve((size_t)nslot); pcost.reserve((size_t)nslot);
        pneed.reserve((size_t)nslot + 1); pcap.reserve((size_t)nslot + 1); pneed_off.reserve((size_t)nslot + 1);
        E.seg_ab = ptab.p; E.seg_bb = ptab.p + nslot; E.seg_m = ptab.p + 2 * nslot; E.seg_n = ptab.p + 3 * nslot; E.seg_dmax = ptab.p + 4 * nslot;
        E.seg_rec = ptab.p + 5 * nslot; E.need = pneed.p; E.cap = pcap.p; E.dcap = dcap.p; E.cost_out = pcost.p;
        mid_pieces<<<(unsigned)((nslot + 127) / 128), 128, 0, st>>>(S, E, pbase.p, nslot, mid.p, mid.p + nseg, dcap.p);
        KERNEL_CHECK(); count_launch(3);
        count.reserve((size_t)nslot + 1); cap_off.reserve((size_t)nslot + 1); out_off.reserve((size_t)nslot + 1);
        exclusive_scan_u32_to_u64(pneed.p, pneed_off.p, (int64_t)nslot, tot.p, st);
        exclusive_scan_u32_to_u64(pcap.p, cap_off.p, (int64_t)nslot, tot.p + 1, st);
        CUDA_CHECK(cudaMemcpyAsync(htot, tot.p, sizeof(htot), cudaMemcpyDeviceToHost, st));
        h_pneed_off.resize((size_t)nslot);
        CUDA_CHECK(cudaMemcpyAsync(h_pneed_off.data(), pneed_off.p, sizeof(uint64_t) * nslot, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(stream_wait(st));
        slot_need_off = pneed_off.p; h_slot_need_off = &h_pneed_off; slot_need_total = htot[0]; cap_total = htot[1];
    }
    DevBuf<int32_t> raw; raw.reserve((size_t)cap_total + 4);
    batches(nslot, *h_slot_need_off, slot_need_total, [&](uint64_t g0, uint64_t g1, uint64_t base) {
        script_emit<<<(unsigned)((g1 - g0 + 127) / 128), 128, 0, st>>>(E, g0, g1, slot_need_off, base, cap_off.p, scr.p, raw.p, count.p);
        KERNEL_CHECK(); count_launch();
    });
    exclusive_scan_u32_to_u64(count.p, out_off.p, (int64_t)nslot, tot.p + 2, st);
    uint64_t total = 0;
    CUDA_CHECK(cudaMemcpyAsync(&total, tot.p + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    if (err == 2) {
        set_error(flavour == 1 ? "Bad alignment between trace points (Compute_Trace_MID)" : "damar_trace_point_scripts: a segment's script disagrees with its distance");
        return 2;
    }
    DevBuf<int32_t> outb; outb.reserve((size_t)total + 4);
    DevBuf<int64_t> roff; roff.reserve((size_t)novl + 2);
    script_compact<<<(unsigned)((nslot + 255) / 256), 256, 0, st>>>(raw.p, cap_off.p, count.p, out_off.p, nslot, outb.p);
    KERNEL_CHECK();
    if (flavour == 1) {
        mid_record_offsets<<<(unsigned)((novl + 1 + 255) / 256), 256, 0, st>>>(pbase.p, novl, out_off.p, nslot, total, roff.p);
        KERNEL_CHECK();
        mid_record_diffs<<<(unsigned)((novl + 255) / 256), 256, 0, st>>>(P, pbase.p, pcost.p, d_rec.p);
        KERNEL_CHECK(); count_launch();
        CUDA_CHECK(cudaMemcpyAsync(rec_diffs, d_rec.p, sizeof(int32_t) * novl, cudaMemcpyDeviceToHost, st));
    } else {
        script_record_offsets<<<(unsigned)((novl + 1 + 255) / 256), 256, 0, st>>>(d_toff.p, novl, out_off.p, nseg, total, roff.p);
        KERNEL_CHECK();
    }
    count_launch(3);
    int32_t *hs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total + 1));
    if (hs == nullptr) { set_error("out of host memory for %llu script values", (unsigned long long)total); return 2; }
    if (total) CUDA_CHECK(cudaMemcpyAsync(hs, outb.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(script_off, roff.p, sizeof(int64_t) * (novl + 1), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(stream_wait(st));
    *script_out = hs; *script_len = (int64_t)total;
    return err;
}

}  // namespace damar
