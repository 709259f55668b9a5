// dn_batch_by_size: native length-bucketed batching of utterances under a padded-token budget.
// Host-side C++ (no CUDA).  Same contract and results as the reference's Cython
// fairseq/data/data_utils_fast.pyx:20-101 (batch_by_size_vec), which `--max-tokens` training uses
// (SURVEY.md §2.2, §8f-1); the normalization runner uses it to cut length-sorted utterances into batches.
//
// Walk the (length-sorted) utterances keeping a *committed* batch [start, ends[cur]) and a *tail* after it.
// The tail is absorbed whenever batch+tail is still within budget and its size is a multiple of bsz_mult (or
// smaller than it); when batch+tail overflows, the committed batch is closed and the tail starts the next one
// (closing the tail-without-the-newcomer as well when the tail alone overflows).
#include <stdint.h>

#include "../../include/diffnorm_b200.h"

extern "C" int64_t dn_batch_by_size(const int64_t* num_tokens, int64_t n, int64_t max_tokens, int64_t max_sentences,
                                    int32_t bsz_mult, int64_t* batch_ends) {
    if (!num_tokens || !batch_ends || n < 0 || bsz_mult < 1) return DN_EINVAL;
    if (n == 0) return 0;
    if (max_tokens > 0)
        for (int64_t i = 0; i < n; ++i)
            if (num_tokens[i] > max_tokens) return DN_EINVAL;  // the reference asserts the same (:29-31)
    int64_t cur = 0;        // index of the running (not yet closed) batch end in batch_ends
    int64_t start = 0;      // first utterance of the running batch
    int64_t batch_max = 0;  // longest utterance in the committed part
    int64_t tail_max = 0;   // longest utterance in the tail
    batch_ends[0] = 0;
    for (int64_t pos = 0; pos < n; ++pos) {
        const int64_t len = num_tokens[pos];
        if (len > tail_max) tail_max = len;
        const int64_t end = pos + 1;
        int64_t cand_max = batch_max > tail_max ? batch_max : tail_max;
        const int64_t sentences = end - start;
        const bool overflow = (max_sentences > 0 && sentences > max_sentences) ||
                              (max_tokens > 0 && sentences * cand_max > max_tokens);
        const bool fits_mult = sentences < bsz_mult || sentences % bsz_mult == 0;
        if (overflow) {
            const int64_t tail_tokens = tail_max * (end - batch_ends[cur]);
            if (max_tokens > 0 && tail_tokens > max_tokens) {  // the tail alone is too big: close it before `pos`
                batch_ends[++cur] = pos;
                tail_max = len;
            }
            start = batch_ends[cur];
            ++cur;
            cand_max = tail_max;
        }
        if (overflow || fits_mult) {
            batch_ends[cur] = end;
            batch_max = cand_max;
            tail_max = 0;
        }
    }
    if (batch_ends[cur] != n) batch_ends[++cur] = n;  // the unabsorbed tail rides with the last batch split
    return cur + 1;
}
