// Insert-size model of the alignment path, as the reference estimates it before the run:
//   processBAM::estimateInsertSize               mapper/processBAM.cpp:1071-1181
//   processBAM::calculateInsertSizeFromHistogram mapper/processBAM.cpp:991-1069
//   alignerBase::alignedReadPair_strandsValid / _pairsDistancesUnderlyingSequences  mapper/aligner/alignerBase.cpp:213-245, 290-329
//   verboseSeedChain::alignment_{begin,end}_originalSequenceAnchors               mapper/reads/verboseSeedChain.h:231-283
// The sample (which records of which pairs) comes from the BAM ingest (bam_reader.h: BamBatch::Sample); the alignments of its primary records
// come from the GPU chain kernels; this file is the arithmetic in between: strand rule, distances along the underlying sequences both
// alignments are anchored in, the weighted histogram, and (weighted median, max(|median - p20|, |median - p80|)).
#pragma once
#include "prg_graph.h"
#include <cstdint>
#include <vector>

namespace hlala {

struct InsertSizeEstimate { double mean = 0, sd = 0; int64_t used = 0, skipped = 0; };

// first_level / last_level / reverse: per read (2 per pair; first mate first) of the aligned primary records; -1 levels = nothing aligned.
// loaded_contigs: indices into g's contigs whose translation the reference has loaded when it computes the distances.
// Throws std::runtime_error where the reference would fail an assertion (no usable pair).
InsertSizeEstimate estimate_insert_size(const FlatGraph& g, int64_t n_pairs, const int32_t* first_level, const int32_t* last_level, const uint8_t* reverse,
                                        const int32_t* loaded_contigs, int32_t n_loaded);

} // namespace hlala
