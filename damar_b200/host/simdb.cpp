// simdb -- synthetic read generator + Dazzler DB writer (host tool of damar_b200).
//
// Produces the DB that `simulator | FA2db -x<min> | DBsplit -s<size>` of the reference would
// produce, without the FASTA detour:
//   * read model of db/simulator.c:100-340 (random genome with base bias, log-normal read lengths
//     drawn through the same 60000-entry half-normal table, error budget len*e split into
//     ins .73333 / del .2 / sub .06667, placed by the same sequential urn scheme, 50% strand flip,
//     one drand48 stream seeded with -r);
//   * FA2db's on-disk layout: 88-byte HITS_DB header + 32-byte HITS_READ per read (`.idx`),
//     2 bit/base MSB-first byte-aligned reads (`.bps`), text stub (`.db`)
//     (db/FA2db.c:600-632, 1100-1145; db/DB.h:319-325, 361-404);
//   * DBsplit's partition rule: a block closes when its bases reach size*10^6 (db/DBsplit.c:196-227).
//
// usage: simdb <out_dir> <root> <genome_mbp> [-c cov] [-b bias] [-r seed] [-m mean] [-s sdev]
//              [-x shortest] [-e error] [-X fa2db_min_len] [-S block_size_mbp]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static const double INS_RATE = .73333, IDL_RATE = .93333, FLIP_RATE = .5;
static const int UNORM_LEN = 60000;
static const double UNORM_MAX = 6.0;
static std::vector<double> unorm_table(UNORM_LEN + 1);
static double unorm_scale;

static void init_unorm() {                                  // simulator.c:152-177
    double del, sum = 0, x;
    unorm_scale = del = UNORM_MAX / UNORM_LEN;
    for (int i = 0; i < UNORM_LEN; i++) { x = i * del; unorm_table[i] = sum; sum += exp(-.5 * x * x) * del; }
    unorm_table[UNORM_LEN] = sum;
    sum *= 2.;
    for (int i = 0; i < UNORM_LEN; i++) unorm_table[i] /= sum;
    unorm_table[UNORM_LEN] = 1.;
}
static double sample_unorm(double x) {                      // simulator.c:179-226
    double y = (x >= .5) ? x - .5 : .5 - x;
    int l = 0, r = UNORM_LEN;
    while (l < r) { int m = (l + r) >> 1; if (y < unorm_table[m]) r = m; else l = m + 1; }
    const int f = r;
    y = (f - (unorm_table[f] - y) / (unorm_table[f] - unorm_table[f - 1])) * unorm_scale;
    return (x < .5) ? -y : y;
}

#pragma pack(push, 1)
struct HitsDbDisk {          // db/DB.h:361-389, 88 bytes on LP64
    int32_t ureads; float freq[4]; int32_t maxlen; int64_t totlen; int32_t nreads, part, ufirst; int32_t pad0;
    uint64_t path; int32_t loaded; int32_t pad1; uint64_t bases, reads, tracks;
};
struct HitsReadDisk {        // db/DB.h:319-325, 32 bytes
    int32_t rlen; int32_t pad0; int64_t boff; int64_t coff; int32_t flags; int32_t pad1;
};
#pragma pack(pop)
static_assert(sizeof(HitsDbDisk) == 88, "HITS_DB layout");
static_assert(sizeof(HitsReadDisk) == 32, "HITS_READ layout");

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: simdb <out_dir> <root> <genome_mbp> [-c -b -r -m -s -x -e -X -S]\n"); return 1; }
    const std::string dir = argv[1], root = argv[2];
    const int GENOME = (int)(atof(argv[3]) * 1000000.);
    double COVERAGE = 20., BIAS = .5, ERROR = .15; int SEED = 1, RMEAN = 10000, RSDEV = 2000, RSHORT = 4000;
    int MINLEN = 1000; double BSIZE = 200.;
    for (int i = 4; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] != '-' || !a[1]) { fprintf(stderr, "simdb: bad argument %s\n", a); return 1; }
        const char *v = a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : "");
        switch (a[1]) {
            case 'c': COVERAGE = atof(v); break;  case 'b': BIAS = atof(v); break;   case 'r': SEED = atoi(v); break;
            case 'm': RMEAN = atoi(v); break;     case 's': RSDEV = atoi(v); break;  case 'x': RSHORT = atoi(v); break;
            case 'e': ERROR = atof(v); break;     case 'X': MINLEN = atoi(v); break; case 'S': BSIZE = atof(v); break;
            default: fprintf(stderr, "simdb: unknown option %s\n", a); return 1;
        }
    }
    if (GENOME < RSHORT) { fprintf(stderr, "Genome length is less than shortest read length !\n"); return 1; }

    // genome (simulator.c:100-128)
    std::vector<char> genome((size_t)GENOME + 1);
    {
        const double PRA = BIAS / 2., PRC = (1. - BIAS) / 2. + PRA, PRG = (1. - BIAS) / 2. + PRC;
        srand48(SEED);
        for (int i = 0; i < GENOME; i++) {
            const double x = drand48();
            genome[i] = (x < PRA) ? 0 : (x < PRC) ? 1 : (x < PRG) ? 2 : 3;
        }
        genome[GENOME] = 4;
    }

    const std::string hidden = dir + "/." + root;
    FILE *bps = fopen((hidden + ".bps").c_str(), "wb"), *idx = fopen((hidden + ".idx").c_str(), "wb");
    if (!bps || !idx) { fprintf(stderr, "simdb: cannot create %s.{bps,idx}\n", hidden.c_str()); return 1; }
    HitsDbDisk db; memset(&db, 0, sizeof(db));
    fwrite(&db, sizeof(db), 1, idx);

    double nsdev = (1. * RSDEV) / RMEAN;
    nsdev = log(1. + nsdev * nsdev);
    const double nmean = log(1. * RMEAN) - .5 * nsdev;
    nsdev = sqrt(nsdev);
    init_unorm();

    std::vector<char> rbuf; std::vector<unsigned char> packed;
    std::vector<int32_t> rlens;
    int64_t totlen = 0, kept_tot = 0, offset = 0, count[4] = {0, 0, 0, 0};
    const int64_t totbp = (int64_t)(COVERAGE * GENOME);
    int maxlen = 0;
    while (totlen < totbp) {                                // simulator.c:262-335
        int len = (int)exp(nmean + nsdev * sample_unorm(drand48()));
        if (len > GENOME) len = GENOME;
        if (len < RSHORT) continue;
        int sdl = (int)(len * ERROR), ins = 0, del = 0;
        for (int j = 0; j < sdl; j++) { const double x = drand48(); if (x < INS_RATE) ins++; else if (x < IDL_RATE) del++; }
        sdl -= ins;
        const int elen = len + (ins - del);
        const int rbeg = (int)(drand48() * ((GENOME - len) + .9999999));
        if ((int)rbuf.size() < elen + 8) rbuf.resize((size_t)(1.2 * elen) + 1024);
        char *t = rbuf.data(); const char *s = genome.data() + rbeg;
        while ((len + 1) * drand48() < ins) { *t++ = (char)(4. * drand48()); ins--; }
        for (; len > 0; len--) {
            if (len * drand48() >= sdl) *t++ = *s;
            else if (sdl * drand48() >= del) { double x = 3. * drand48(); if (x >= *s) x += 1.; *t++ = (char)x; sdl--; }
            else { del--; sdl--; }
            s++;
            while (len * drand48() < ins) { *t++ = (char)(4. * drand48()); ins--; }
        }
        if (drand48() >= FLIP_RATE) {                       // reverse complement
            char *a = rbuf.data(), *b = rbuf.data() + (elen - 1);
            while (a <= b) { const char c = *a; *a = (char)(3 - *b); *b = (char)(3 - c); a++; b--; }
        }
        totlen += elen;
        if (elen < MINLEN) continue;                        // FA2db -x
        // FA2db addReadToDB: 2 bit/base MSB first, byte aligned
        const int clen = (elen + 3) >> 2;
        packed.assign((size_t)clen, 0);
        for (int i = 0; i < elen; i++) { const int c = rbuf[i]; count[c]++; packed[i >> 2] |= (unsigned char)(c << (6 - 2 * (i & 3))); }
        fwrite(packed.data(), 1, (size_t)clen, bps);
        HitsReadDisk hr; memset(&hr, 0, sizeof(hr));
        hr.rlen = elen; hr.boff = offset; hr.coff = -1; hr.flags = 0x0800;   // DB_BEST
        fwrite(&hr, sizeof(hr), 1, idx);
        offset += clen; kept_tot += elen; rlens.push_back(elen);
        if (elen > maxlen) maxlen = elen;
    }
    db.ureads = (int32_t)rlens.size();
    for (int c = 0; c < 4; c++) db.freq[c] = (float)((1. * count[c]) / kept_tot);
    db.maxlen = maxlen; db.totlen = kept_tot; db.nreads = db.ureads;
    rewind(idx); fwrite(&db, sizeof(db), 1, idx);
    fclose(idx); fclose(bps);

    // stub with DBsplit partition (db/DBsplit.c:196-227; formats db/DB.h:400-404)
    FILE *stub = fopen((dir + "/" + root + ".db").c_str(), "w");
    if (!stub) { fprintf(stderr, "simdb: cannot create stub\n"); return 1; }
    fprintf(stub, "files = %9d\n", 1);
    fprintf(stub, "  %9d %s %s\n", db.ureads, "sim", "Sim");
    std::vector<int> firsts; firsts.push_back(0);
    {
        const int64_t size = (int64_t)(BSIZE * 1000000ll);
        int64_t tl = 0; int ireads = 0;
        for (int i = 0; i < db.ureads; i++) {
            ireads++; tl += rlens[i];
            if (tl >= size) { firsts.push_back(i + 1); tl = 0; ireads = 0; }
        }
        if (ireads > 0) firsts.push_back(db.ureads);
    }
    fprintf(stub, "blocks = %9d\n", (int)firsts.size() - 1);
    fprintf(stub, "size = %9lld\n", (long long)BSIZE);
    for (int f : firsts) fprintf(stub, " %9d\n", f);
    fclose(stub);
    printf("%d reads, %lld bases, maxlen %d, %d blocks\n", db.ureads, (long long)kept_tot, maxlen, (int)firsts.size() - 1);
    return 0;
}
