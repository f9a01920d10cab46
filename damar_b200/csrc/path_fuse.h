// path_fuse.h -- host-side multi-alignment post-pass (Handle_Redundancies, reference dalign/filter.c:1804-2077)
#pragma once
#include <cstdint>
#include <vector>

namespace damar {

struct FusePath {                     // Path (align.h:125-132) with its uncompressed trace
    int abpos = 0, bbpos = 0, aepos = 0, bepos = 0, diffs = 0;
    std::vector<uint16_t> trace;
};

// the two reads of the pair in the orientation that was aligned (B complemented for comp), fetched only when a
// bridge needs an exact realignment
struct FuseSeqs {
    virtual ~FuseSeqs() = default;
    virtual void fetch() const = 0;
    virtual const uint8_t *aseq() const = 0;
    virtual const uint8_t *bseq() const = 0;
    virtual int alen() const = 0;
    virtual int blen() const = 0;
};

// Returns the number of surviving alignments (both lists compacted).  bm may be null.
int handle_redundancies(std::vector<FusePath> &am, std::vector<FusePath> *bm, int comp, int ts, const FuseSeqs &seqs);

}  // namespace damar
