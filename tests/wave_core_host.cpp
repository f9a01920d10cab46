// wave_core_host.cpp -- TEST INFRASTRUCTURE: runs damar_b200/csrc/wave_core.cuh (the scalar wave-extension
// code the CUDA kernels execute one task per thread) on the CPU, so that tests/test_wave_core.py can
// compare it with the oracle's Local_Alignment without a GPU.  Never linked into the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../damar_b200/csrc/wave_core.cuh"

using namespace damar::wc;

namespace {

struct HostPool {                      // copied by value into push_raw like the device pool: shares the vector
    std::vector<Cell> *v;
    uint32_t take(int n) { uint32_t b = (uint32_t)v->size(); v->resize(v->size() + n); return b; }
    void store(uint32_t id, const Cell &c) { (*v)[id] = c; }
};

// 2 bit/base, 16 bases per word, first base in the top bits; `pad` slack words at the end
std::vector<uint32_t> pack(const std::vector<uint8_t> &s) {
    std::vector<uint32_t> w(s.size() / 16 + 10, 0u);
    for (size_t i = 0; i < s.size(); i++) w[i >> 4] |= (uint32_t)(s[i] & 3) << (30 - 2 * (int)(i & 15));
    return w;
}

struct HostTabs {
    int16_t t7[128], t8[256];
    TrimTabs tabs;
    explicit HostTabs(int dscore) {
        for (int x = 0; x < 128; x++) t7[x] = (int16_t)min_suffix_score(x, 7, dscore);
        for (int x = 0; x < 256; x++) t8[x] = (int16_t)min_suffix_score(x, 8, dscore);
        tabs = TrimTabs{t7, t8, dscore};
    }
};

int g_nway = 2;

template <int RING>
int run_pass(const TaskIn &t, const TrimTabs &tabs, HostPool &pool, TaskOut &out) {
    uint32_t store[FRONT_WORDS * RING];
    memset(store, 0xee, sizeof(store));
    Front<RING, 1> F{store};
    Lane L;
    setup(L, t, F, pool);
    while (L.status == ST_RUN) {
        if (L.k < L.low) transition(L, F);
        if (L.status == ST_RUN && L.k >= L.low) {
            if (g_nway == 2) cellsN<2>(L, F, tabs, pool);
            else if (g_nway == 4) cellsN<4>(L, F, tabs, pool);
            else cellsN<3>(L, F, tabs, pool);
        }
    }
    finish(L, out);
    return L.status;
}

}  // namespace

extern "C" {

// the 768-byte decomposition of the trim tables against the full 2 x 32768 tables of align.c:286-318
int wc_check_trim_tables(int dscore, const int16_t *score_tab, const int16_t *table_tab) {
    HostTabs ht(dscore);
    int bad = 0;
    for (int x = 0; x < 32768; x++)
        if (table15(ht.tabs, x) != table_tab[x] || score15(ht.tabs, x) != score_tab[x]) bad++;
    // the fused 30-bit test on a pseudo-random sample of windows (plus garbage above bit 29)
    uint64_t r = 88172645463325252ull;
    for (int i = 0; i < 2000000; i++) {
        r ^= r << 13; r ^= r >> 7; r ^= r << 17;
        const int lo = (int)(r & 0x7fff), hi = (int)((r >> 15) & 0x7fff);
        const bool want = table_tab[lo] >= 0 && table_tab[hi] + score_tab[lo] >= 0;
        if (trim_ok(ht.tabs, r) != want) bad++;
    }
    return bad;
}
void wc_set_nway(int n) { g_nway = n; }

// Local_Alignment for daligner's call shape on unpacked sequences (values 0..3).  `bseq` is the B read in the
// orientation that is aligned (already complemented when comp).  Returns 0, or the failing pass status.
// aout/bout = {abpos, bbpos, aepos, bepos, diffs, tlen}; traces are written from index 0.
int wc_local_alignment(const uint8_t *aseq, int alen, const uint8_t *bseq, int blen, int comp, int selfie, int diag, int anti,
                       int TS, int path_ave, int dscore, int ring,
                       int32_t *aout, uint16_t *atrace, int32_t *bout, uint16_t *btrace, int32_t *stats) {
    std::vector<uint8_t> a(aseq, aseq + alen), b(bseq, bseq + blen), arc(alen), brc(blen);
    for (int i = 0; i < alen; i++) arc[i] = 3 - a[alen - 1 - i];
    for (int i = 0; i < blen; i++) brc[i] = 3 - b[blen - 1 - i];
    // put the reads at a non-zero, non-aligned base offset like reads inside a block
    const uint32_t abase = 37, bbase = 5;
    auto place = [](const std::vector<uint8_t> &s, uint32_t base) {
        std::vector<uint8_t> v(base, 1);
        v.insert(v.end(), s.begin(), s.end());
        v.push_back(2);
        return pack(v);
    };
    std::vector<uint32_t> aw = place(a, abase), awr = place(arc, abase), bw = place(b, bbase), bwr = place(brc, bbase);

    TaskIn t{};
    t.abase = abase; t.bbase = bbase; t.alen = alen; t.blen = blen; t.diag = diag; t.anti = anti;
    t.aoff = 0; t.boff = comp ? blen % TS : 0;
    t.minp = (selfie && diag >= 0) ? 1 : -IMAX; t.maxp = (selfie && diag <= 0) ? -1 : IMAX;
    t.TS = TS; t.path_ave = path_ave;
    HostTabs ht(dscore);

    std::vector<Cell> cellvec;
    HostPool pool{&cellvec};
    TaskOut fo{}, ro{};
    t.rev = 0; t.aw = aw.data(); t.bw = bw.data();
    int st = (ring == 64) ? run_pass<64>(t, ht.tabs, pool, fo) : run_pass<32>(t, ht.tabs, pool, fo);
    if (st != ST_DONE) return st;
    t.rev = 1; t.aw = awr.data(); t.bw = bwr.data();
    st = (ring == 64) ? run_pass<64>(t, ht.tabs, pool, ro) : run_pass<32>(t, ht.tabs, pool, ro);
    if (st != ST_DONE) return st + 100;
    if (stats) { stats[0] = fo.waves + ro.waves; stats[1] = fo.cells + ro.cells; stats[2] = (int32_t)cellvec.size(); }

    const int maxtp = 2 * ((alen > blen ? alen : blen) / TS + 2) + 2;
    std::vector<uint16_t> tbuf(8 * maxtp, 0);
    uint16_t *atr = tbuf.data() + 2 * maxtp, *btr = tbuf.data() + 6 * maxtp;
    const int flenA = emit_chain_fwd<true>(cellvec.data(), fo.ha, true, anti, fo.ex, fo.ey, fo.ediff, atr);
    const int flenB = emit_chain_fwd<true>(cellvec.data(), fo.hb, false, anti, fo.ex, fo.ey, fo.ediff, btr);
    const int rroot = reverse_chain(cellvec.data(), ro.ha);
    const int preA = emit_chain_rev<true>(cellvec.data(), rroot, true, ro.ex, ro.ey, ro.ediff, TS, t.aoff, flenA, atr);
    const int preB = emit_chain_rev<true>(cellvec.data(), rroot, false, ro.ex, ro.ey, ro.ediff, TS, t.boff, flenB, btr);
    const int atlen = flenA + preA, btlen = flenB + preB;
    uint16_t *at = atr - preA, *bt = btr - preB;
    const int abpos = ro.ex, bbpos = ro.ey, aepos = fo.ex, bepos = fo.ey, diffs = fo.ediff + ro.ediff;
    aout[0] = abpos; aout[1] = bbpos; aout[2] = aepos; aout[3] = bepos; aout[4] = diffs; aout[5] = atlen;
    memcpy(atrace, at, sizeof(uint16_t) * atlen);
    if (comp) {                                              // align.c:2033-2055
        bout[0] = blen - bepos; bout[1] = alen - aepos; bout[2] = blen - bbpos; bout[3] = alen - abpos;
        const int np = btlen >> 1;
        for (int i = 0; i < np / 2; i++) {
            const int j = np - 1 - i;
            uint16_t d0 = bt[2 * i], d1 = bt[2 * i + 1];
            bt[2 * i] = bt[2 * j]; bt[2 * i + 1] = bt[2 * j + 1]; bt[2 * j] = d0; bt[2 * j + 1] = d1;
        }
    } else { bout[0] = bbpos; bout[1] = abpos; bout[2] = bepos; bout[3] = aepos; }
    bout[4] = diffs; bout[5] = btlen;
    memcpy(btrace, bt, sizeof(uint16_t) * btlen);
    return 0;
}

}  // extern "C"
